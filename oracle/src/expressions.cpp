// ORACLE — TEST INFRASTRUCTURE ONLY. See expressions.h for the reference file:line map.
#include "expressions.h"

#include <algorithm>
#include <limits>
#include <map>
#include <variant>

namespace oracle {

namespace {

#define CHECK_QUERY(condition, message)         \
   if (!(condition)) {                          \
      throw IllegalQueryException(message);     \
   }

const SequenceColumn& requireColumn(const Table& table, const std::string& name) {
   const SequenceColumn* column = table.findColumn(name);
   // validateSequenceName, query_parse_sequence_name.h:10-20
   CHECK_QUERY(column != nullptr, "Database does not contain the Sequence with name: '" + name + "'");
   return *column;
}

AmbiguityMode invertMode(AmbiguityMode mode) {
   if (mode == AmbiguityMode::UPPER_BOUND) {
      return AmbiguityMode::LOWER_BOUND;
   }
   if (mode == AmbiguityMode::LOWER_BOUND) {
      return AmbiguityMode::UPPER_BOUND;
   }
   return mode;
}

bool containsSymbol(const std::vector<Symbol>& symbols, Symbol symbol) {
   return std::find(symbols.begin(), symbols.end(), symbol) != symbols.end();
}

// ---------- literals ----------

class BoolLiteral : public Expression {
  public:
   bool value;
   explicit BoolLiteral(bool value) : value(value) {}
   std::string toString() const override { return value ? "true" : "false"; }
   ExprPtr rewrite(const Table&, AmbiguityMode) const override {
      return std::make_shared<BoolLiteral>(value);
   }
   std::unique_ptr<Operator> compile(const Table& table) const override {
      if (value) {
         return std::make_unique<Full>(table.row_layout);
      }
      return std::make_unique<Empty>(table.row_layout);
   }
};

// ---------- SymbolInSet ----------

class SymbolInSet : public Expression {
  public:
   std::string column;
   uint32_t position_idx;
   std::vector<Symbol> symbols;
   SymbolInSet(std::string column, uint32_t position_idx, std::vector<Symbol> symbols)
       : column(std::move(column)),
         position_idx(position_idx),
         symbols(std::move(symbols)) {}
   std::string toString() const override {
      return "(" + column + ":symbol at position " + std::to_string(position_idx + 1) + " in set)";
   }
   ExprPtr rewrite(const Table&, AmbiguityMode) const override {
      throw QueryCompilationException(
         "Cannot rewrite SymbolInSet - this expression should only be created during query rewrites "
         "and not directly used"
      );
   }
   std::unique_ptr<Operator> compile(const Table& table) const override {
      const auto& sequence_column = requireColumn(table, column);
      return compileSymbolInSet(sequence_column, position_idx, symbols, table.row_layout);
   }
};

// pre-rewrite form of a raw (sym-in ...) so that it survives computeFilter's rewrite pass
class RawSymbolInSet : public Expression {
  public:
   std::shared_ptr<SymbolInSet> inner;
   explicit RawSymbolInSet(std::shared_ptr<SymbolInSet> inner) : inner(std::move(inner)) {}
   std::string toString() const override { return inner->toString(); }
   ExprPtr rewrite(const Table&, AmbiguityMode) const override { return inner; }
   std::unique_ptr<Operator> compile(const Table& table) const override { return inner->compile(table); }
};

// ---------- SymbolEquals (symbol_equals.cpp:65-100) ----------

class SymbolEquals : public Expression {
  public:
   std::string column;
   uint32_t position_idx;
   std::optional<Symbol> value;  // nullopt = '.'
   SymbolEquals(std::string column, uint32_t position_idx, std::optional<Symbol> value)
       : column(std::move(column)),
         position_idx(position_idx),
         value(value) {}
   std::string toString() const override { return column + ":" + std::to_string(position_idx + 1); }
   ExprPtr rewrite(const Table& table, AmbiguityMode mode) const override {
      const auto& sequence_column = requireColumn(table, column);
      CHECK_QUERY(
         position_idx < sequence_column.reference_sequence.size(),
         "SymbolEquals<" + sequence_column.alphabet->symbol_name + "> position is out of bounds " +
            std::to_string(position_idx + 1) + " > " +
            std::to_string(sequence_column.reference_sequence.size())
      );
      const Symbol symbol = value.value_or(sequence_column.reference_sequence.at(position_idx));
      if (mode == AmbiguityMode::UPPER_BOUND) {
         return std::make_shared<SymbolInSet>(
            column, position_idx, sequence_column.alphabet->ambiguity_symbols.at(symbol)
         );
      }
      return std::make_shared<SymbolInSet>(column, position_idx, std::vector<Symbol>{symbol});
   }
   std::unique_ptr<Operator> compile(const Table&) const override {
      throw QueryCompilationException("SymbolEquals should have been rewritten before compilation");
   }
};

// ---------- HasMutation (has_mutation.cpp:34-67) ----------

class HasMutation : public Expression {
  public:
   std::string column;
   uint32_t position_idx;
   HasMutation(std::string column, uint32_t position_idx)
       : column(std::move(column)),
         position_idx(position_idx) {}
   std::string toString() const override { return column + ":" + std::to_string(position_idx); }
   ExprPtr rewrite(const Table& table, AmbiguityMode mode) const override {
      const auto& sequence_column = requireColumn(table, column);
      const Alphabet& alphabet = *sequence_column.alphabet;
      CHECK_QUERY(
         position_idx < sequence_column.reference_sequence.size(),
         "Has" + alphabet.symbol_name + "Mutation position is out of bounds " +
            std::to_string(position_idx + 1) + " > " +
            std::to_string(sequence_column.reference_sequence.size())
      );
      const Symbol ref_symbol = sequence_column.reference_sequence.at(position_idx);
      std::vector<Symbol> symbols;
      for (uint32_t symbol = 0; symbol < alphabet.count; ++symbol) {
         symbols.push_back(static_cast<Symbol>(symbol));
      }
      if (mode == AmbiguityMode::UPPER_BOUND) {
         std::erase(symbols, ref_symbol);
      } else {
         for (const Symbol symbol : alphabet.ambiguity_symbols.at(ref_symbol)) {
            std::erase(symbols, symbol);
         }
      }
      return std::make_shared<SymbolInSet>(column, position_idx, std::move(symbols));
   }
   std::unique_ptr<Operator> compile(const Table&) const override {
      throw QueryCompilationException("HasMutation expression must be eliminated in query rewrite phase");
   }
};

// ---------- Negation / Maybe / Exact ----------

class Negation : public Expression {
  public:
   ExprPtr child;
   explicit Negation(ExprPtr child) : child(std::move(child)) {}
   std::string toString() const override { return "!(" + child->toString() + ")"; }
   ExprPtr rewrite(const Table& table, AmbiguityMode mode) const override {
      return std::make_shared<Negation>(child->rewrite(table, invertMode(mode)));
   }
   std::unique_ptr<Operator> compile(const Table& table) const override {
      return Operator::negate(child->compile(table));
   }
};

class Maybe : public Expression {
  public:
   ExprPtr child;
   explicit Maybe(ExprPtr child) : child(std::move(child)) {}
   std::string toString() const override { return "Maybe (" + child->toString() + ")"; }
   ExprPtr rewrite(const Table& table, AmbiguityMode) const override {
      return child->rewrite(table, AmbiguityMode::UPPER_BOUND);
   }
   std::unique_ptr<Operator> compile(const Table&) const override {
      throw QueryCompilationException("Maybe expression must be elimitated in query rewrite phase");
   }
};

class Exact : public Expression {
  public:
   ExprPtr child;
   explicit Exact(ExprPtr child) : child(std::move(child)) {}
   std::string toString() const override { return "Exact (" + child->toString() + ")"; }
   ExprPtr rewrite(const Table& table, AmbiguityMode) const override {
      return child->rewrite(table, AmbiguityMode::LOWER_BOUND);
   }
   std::unique_ptr<Operator> compile(const Table&) const override {
      throw QueryCompilationException("Exact expression must be elimitated in query rewrite phase");
   }
};

// ---------- And (and.cpp:91-219) ----------

void appendOperators(OperatorVector& from, OperatorVector& to) {
   for (auto& op : from) {
      to.push_back(std::move(op));
   }
}

class And : public Expression {
  public:
   ExpressionVector children;
   explicit And(ExpressionVector children) : children(std::move(children)) {}
   std::string toString() const override {
      std::string res = "And(";
      for (const auto& child : children) {
         res += child->toString() + " & ";
      }
      return res + ")";
   }
   ExprPtr rewrite(const Table& table, AmbiguityMode mode) const override {
      ExpressionVector rewritten;
      rewritten.reserve(children.size());
      for (const auto& child : children) {
         rewritten.push_back(child->rewrite(table, mode));
      }
      return std::make_shared<And>(std::move(rewritten));
   }
   std::unique_ptr<Operator> compile(const Table& table) const override {
      OperatorVector unprocessed;
      for (const auto& child : children) {
         unprocessed.push_back(child->compile(table));
      }
      OperatorVector non_negated;
      OperatorVector negated;
      PredicateVector predicates;
      bool found_empty = false;
      while (!unprocessed.empty()) {
         auto child = std::move(unprocessed.back());
         unprocessed.pop_back();
         if (child->type() == FULL) {
            continue;
         }
         if (child->type() == EMPTY) {
            found_empty = true;
            break;
         }
         if (child->type() == INTERSECTION) {
            auto* intersection_child = static_cast<Intersection*>(child.get());
            appendOperators(intersection_child->children, non_negated);
            appendOperators(intersection_child->negated_children, negated);
         } else if (child->type() == COMPLEMENT) {
            negated.emplace_back(Operator::negate(std::move(child)));
         } else if (child->type() == SELECTION) {
            auto* selection_child = static_cast<Selection*>(child.get());
            for (auto& predicate : selection_child->predicates) {
               predicates.push_back(std::move(predicate));
            }
            if (selection_child->child_operator.has_value()) {
               unprocessed.emplace_back(std::move(selection_child->child_operator.value()));
            }
         } else {
            non_negated.push_back(std::move(child));
         }
      }
      if (found_empty) {
         non_negated.clear();
         negated.clear();
         predicates.clear();
         non_negated.emplace_back(std::make_unique<Empty>(table.row_layout));
      }
      if (non_negated.empty() && negated.empty()) {
         if (predicates.empty()) {
            return std::make_unique<Full>(table.row_layout);
         }
         return std::make_unique<Selection>(std::nullopt, std::move(predicates), table.row_layout);
      }
      std::unique_ptr<Operator> index_arithmetic_operator;
      if (non_negated.size() == 1 && negated.empty()) {
         index_arithmetic_operator = std::move(non_negated[0]);
      } else if (negated.size() == 1 && non_negated.empty()) {
         index_arithmetic_operator =
            std::make_unique<Complement>(std::move(negated[0]), table.row_layout);
      } else if (non_negated.empty()) {
         auto union_ret = std::make_unique<Union>(std::move(negated), table.row_layout);
         index_arithmetic_operator =
            std::make_unique<Complement>(std::move(union_ret), table.row_layout);
      } else {
         index_arithmetic_operator = std::make_unique<Intersection>(
            std::move(non_negated), std::move(negated), table.row_layout
         );
      }
      if (predicates.empty()) {
         return index_arithmetic_operator;
      }
      return std::make_unique<Selection>(
         std::optional<std::unique_ptr<Operator>>{std::move(index_arithmetic_operator)},
         std::move(predicates),
         table.row_layout
      );
   }
};

// ---------- Or (or.cpp:34-242) ----------

class Or : public Expression {
  public:
   ExpressionVector children;
   explicit Or(ExpressionVector children) : children(std::move(children)) {}
   std::string toString() const override {
      std::string res = "Or(";
      for (const auto& child : children) {
         res += child->toString() + " | ";
      }
      return res + ")";
   }
   ExprPtr rewrite(const Table& table, AmbiguityMode mode) const override {
      // collectChildren, or.cpp:47-68
      std::vector<const Expression*> collected;
      std::vector<const Expression*> queue;
      for (const auto& direct_child : children) {
         queue.push_back(direct_child.get());
      }
      while (!queue.empty()) {
         const auto* current = queue.back();
         queue.pop_back();
         if (const auto* or_child = dynamic_cast<const Or*>(current)) {
            for (const auto& child : or_child->children) {
               queue.push_back(child.get());
            }
         } else {
            collected.push_back(current);
         }
      }
      ExpressionVector rewritten;
      for (const auto* child : collected) {
         rewritten.push_back(child->rewrite(table, mode));
      }
      // algebraicSimplification, or.cpp:70-95
      ExpressionVector non_trivial;
      bool constant_true = false;
      while (!rewritten.empty()) {
         auto child = std::move(rewritten.back());
         rewritten.pop_back();
         if (const auto* literal = dynamic_cast<const BoolLiteral*>(child.get())) {
            if (literal->value) {
               constant_true = true;
               break;
            }
            continue;
         }
         if (const auto* or_child = dynamic_cast<const Or*>(child.get())) {
            for (const auto& grandchild : or_child->children) {
               rewritten.push_back(grandchild);
            }
         } else {
            non_trivial.push_back(std::move(child));
         }
      }
      if (constant_true) {
         non_trivial.clear();
         non_trivial.push_back(std::make_shared<BoolLiteral>(true));
      }
      // rewriteSymbolInSetExpressions, or.cpp:97-124 (Nucleotide pass, then AminoAcid pass; a
      // column has exactly one alphabet, so one keyed pass per alphabet is equivalent)
      for (int pass = 0; pass < 2; ++pass) {
         ExpressionVector new_children;
         std::map<std::pair<std::string, uint32_t>, std::vector<Symbol>> merged;
         for (auto& child : non_trivial) {
            const auto* in_set = dynamic_cast<const SymbolInSet*>(child.get());
            const SequenceColumn* column = in_set != nullptr ? table.findColumn(in_set->column) : nullptr;
            const bool alphabet_matches =
               column != nullptr &&
               (column->alphabet == (pass == 0 ? &Alphabet::nucleotide() : &Alphabet::aminoAcid()));
            if (in_set != nullptr && alphabet_matches) {
               auto& symbols_so_far = merged[{in_set->column, in_set->position_idx}];
               symbols_so_far.insert(symbols_so_far.end(), in_set->symbols.begin(), in_set->symbols.end());
            } else {
               new_children.push_back(std::move(child));
            }
         }
         for (auto& [column_and_position, symbols] : merged) {
            new_children.push_back(std::make_shared<SymbolInSet>(
               column_and_position.first, column_and_position.second, std::move(symbols)
            ));
         }
         non_trivial = std::move(new_children);
      }
      if (non_trivial.size() == 1) {
         return non_trivial[0];
      }
      return std::make_shared<Or>(std::move(non_trivial));
   }
   std::unique_ptr<Operator> compile(const Table& table) const override {
      OperatorVector all_child_operators;
      for (const auto& child : children) {
         all_child_operators.push_back(child->compile(table));
      }
      OperatorVector filtered;
      for (auto& child : all_child_operators) {
         if (child->type() == EMPTY) {
            continue;
         }
         if (child->type() == FULL) {
            return std::make_unique<Full>(table.row_layout);
         }
         if (child->type() == UNION) {
            auto* or_child = static_cast<Union*>(child.get());
            appendOperators(or_child->children, filtered);
         } else {
            filtered.push_back(std::move(child));
         }
      }
      if (filtered.empty()) {
         return std::make_unique<Empty>(table.row_layout);
      }
      if (filtered.size() == 1) {
         return std::move(filtered[0]);
      }
      if (std::any_of(filtered.begin(), filtered.end(), [](const auto& child) {
             return child->type() == COMPLEMENT;
          })) {
         return Complement::fromDeMorgan(std::move(filtered), table.row_layout);
      }
      return std::make_unique<Union>(std::move(filtered), table.row_layout);
   }
};

// ---------- NOf (nof.cpp:33-277) ----------

std::unique_ptr<Operator> nofToOperator(
   const int updated_number_of_matchers,
   OperatorVector&& non_negated,
   OperatorVector&& negated,
   bool match_exactly,
   const RowLayout& row_layout
) {
   const int child_operator_count = static_cast<int>(non_negated.size() + negated.size());
   // handleTrivialCases, nof.cpp:33-84
   if (updated_number_of_matchers > child_operator_count) {
      return std::make_unique<Empty>(row_layout);
   }
   if (updated_number_of_matchers < 0) {
      if (match_exactly) {
         return std::make_unique<Empty>(row_layout);
      }
      return std::make_unique<Full>(row_layout);
   }
   if (updated_number_of_matchers == 0) {
      if (!match_exactly) {
         return std::make_unique<Full>(row_layout);
      }
      if (child_operator_count == 0) {
         return std::make_unique<Full>(row_layout);
      }
      if (child_operator_count == 1) {
         if (non_negated.empty()) {
            return std::move(negated[0]);
         }
         return std::make_unique<Complement>(std::move(non_negated[0]), row_layout);
      }
      if (negated.empty()) {
         auto union_ret = std::make_unique<Union>(std::move(non_negated), row_layout);
         return std::make_unique<Complement>(std::move(union_ret), row_layout);
      }
      return std::make_unique<Intersection>(std::move(negated), std::move(non_negated), row_layout);
   }
   if (updated_number_of_matchers == 1 && child_operator_count == 1) {
      if (negated.empty()) {
         return std::move(non_negated[0]);
      }
      return std::make_unique<Complement>(std::move(negated[0]), row_layout);
   }
   // handleAndCase, nof.cpp:86-98
   if (updated_number_of_matchers == child_operator_count) {
      if (non_negated.empty()) {
         auto union_ret = std::make_unique<Union>(std::move(negated), row_layout);
         return std::make_unique<Complement>(std::move(union_ret), row_layout);
      }
      return std::make_unique<Intersection>(std::move(non_negated), std::move(negated), row_layout);
   }
   // handleOrCase, nof.cpp:100-114
   if (updated_number_of_matchers == 1 && !match_exactly) {
      if (negated.empty()) {
         return std::make_unique<Union>(std::move(non_negated), row_layout);
      }
      auto intersection_ret =
         std::make_unique<Intersection>(std::move(negated), std::move(non_negated), row_layout);
      return std::make_unique<Complement>(std::move(intersection_ret), row_layout);
   }
   return std::make_unique<Threshold>(
      std::move(non_negated),
      std::move(negated),
      static_cast<uint32_t>(updated_number_of_matchers),
      match_exactly,
      row_layout
   );
}

class NOf : public Expression {
  public:
   ExpressionVector children;
   int number_of_matchers;
   bool match_exactly;
   NOf(ExpressionVector children, int number_of_matchers, bool match_exactly)
       : children(std::move(children)),
         number_of_matchers(number_of_matchers),
         match_exactly(match_exactly) {}
   std::string toString() const override {
      return std::string(match_exactly ? "[exactly-" : "[") + std::to_string(number_of_matchers) +
             "-of:" + std::to_string(children.size()) + " children]";
   }
   ExpressionVector rewriteChildren(const Table& table, AmbiguityMode mode) const {
      ExpressionVector rewritten;
      rewritten.reserve(children.size());
      for (const auto& child : children) {
         rewritten.push_back(child->rewrite(table, mode));
      }
      return rewritten;
   }
   ExprPtr rewrite(const Table& table, AmbiguityMode mode) const override {
      if (mode != AmbiguityMode::NONE && match_exactly &&
          std::cmp_less(number_of_matchers, children.size())) {
         // rewriteToNonExact, nof.cpp:219-238
         auto at_least_k = std::make_shared<NOf>(rewriteChildren(table, mode), number_of_matchers, false);
         auto at_least_k_plus_one =
            std::make_shared<NOf>(rewriteChildren(table, mode), number_of_matchers + 1, false);
         ExpressionVector and_children;
         and_children.push_back(std::move(at_least_k));
         and_children.push_back(std::make_shared<Negation>(std::move(at_least_k_plus_one)));
         return std::make_shared<And>(std::move(and_children));
      }
      return std::make_shared<NOf>(rewriteChildren(table, mode), number_of_matchers, match_exactly);
   }
   std::unique_ptr<Operator> compile(const Table& table) const override {
      // mapChildExpressions, nof.cpp:184-215
      OperatorVector non_negated;
      OperatorVector negated;
      int updated_number_of_matchers = number_of_matchers;
      for (const auto& child_expression : children) {
         auto child_operator = child_expression->compile(table);
         if (child_operator->type() == EMPTY) {
            continue;
         }
         if (child_operator->type() == FULL) {
            updated_number_of_matchers--;
            continue;
         }
         if (child_operator->type() == COMPLEMENT) {
            negated.emplace_back(Operator::negate(std::move(child_operator)));
            continue;
         }
         non_negated.push_back(std::move(child_operator));
      }
      if (updated_number_of_matchers < 0) {
         if (match_exactly) {
            return std::make_unique<Empty>(table.row_layout);
         }
         return std::make_unique<Full>(table.row_layout);
      }
      return nofToOperator(
         updated_number_of_matchers,
         std::move(non_negated),
         std::move(negated),
         match_exactly,
         table.row_layout
      );
   }
};

// ---------- MutationProfile (mutation_profile.cpp:60-257) ----------

struct ProfileQuerySequence {
   std::string sequence;
};
struct ProfileRow {
   uint32_t global_row_id;
};
struct ProfileMutations {
   std::vector<std::pair<uint32_t, Symbol>> mutations;
};

class MutationProfile : public Expression {
  public:
   std::string column;
   uint32_t distance;
   std::variant<ProfileQuerySequence, ProfileRow, ProfileMutations> input;
   MutationProfile(
      std::string column,
      uint32_t distance,
      std::variant<ProfileQuerySequence, ProfileRow, ProfileMutations> input
   )
       : column(std::move(column)),
         distance(distance),
         input(std::move(input)) {}
   std::string toString() const override {
      return "MutationProfile(" + column + ":distance=" + std::to_string(distance) + ")";
   }
   ExprPtr rewrite(const Table& table, AmbiguityMode) const override {
      const auto& sequence_column = requireColumn(table, column);
      const Alphabet& alphabet = *sequence_column.alphabet;
      const size_t ref_len = sequence_column.reference_sequence.size();
      std::vector<Symbol> profile;
      if (const auto* query = std::get_if<ProfileQuerySequence>(&input)) {
         CHECK_QUERY(
            query->sequence.size() == ref_len,
            "querySequence length " + std::to_string(query->sequence.size()) +
               " does not match the reference sequence length " + std::to_string(ref_len) + " for " +
               alphabet.symbol_name + " MutationProfile"
         );
         for (char character : query->sequence) {
            const auto symbol = alphabet.charToSymbol(character);
            CHECK_QUERY(
               symbol.has_value(),
               "Invalid " + alphabet.symbol_name + " symbol '" + std::string(1, character) +
                  "' in querySequence for MutationProfile"
            );
            profile.push_back(symbol.value());
         }
      } else if (const auto* row = std::get_if<ProfileRow>(&input)) {
         // reconstructSequenceAtRow, mutation_profile.cpp:81-107: local reference, overwritten by
         // the row's stored diffs, then by the missing symbol outside coverage / at N positions
         const uint32_t row_id = row->global_row_id;
         profile = sequence_column.getLocalReference();
         for (const auto& [key, container] : sequence_column.vertical_sequence_index.vertical_bitmaps) {
            if (key.v_index == (row_id >> 16) && container.contains(static_cast<uint16_t>(row_id & 0xFFFF))) {
               profile.at(key.position) = key.symbol;
            }
         }
         const auto [start, end] = sequence_column.horizontal_coverage_index.coverageRange(row_id);
         for (uint32_t position = 0; position < ref_len; ++position) {
            if (position < start || position >= end) {
               profile[position] = alphabet.missing;
            }
         }
         auto iter = sequence_column.horizontal_coverage_index.horizontal_bitmaps.find(row_id);
         if (iter != sequence_column.horizontal_coverage_index.horizontal_bitmaps.end()) {
            iter->second.forEach([&](uint32_t position) { profile.at(position) = alphabet.missing; });
         }
      } else {
         profile = sequence_column.reference_sequence;
         for (const auto& [position_idx, symbol] : std::get<ProfileMutations>(input).mutations) {
            CHECK_QUERY(
               position_idx < ref_len,
               alphabet.symbol_name + " MutationProfile mutation position " +
                  std::to_string(position_idx + 1) + " is out of bounds (reference length " +
                  std::to_string(ref_len) + ")"
            );
            profile[position_idx] = symbol;
         }
      }
      ExpressionVector difference_children;
      for (size_t pos = 0; pos < profile.size(); ++pos) {
         const Symbol profile_symbol = profile[pos];
         if (profile_symbol == alphabet.missing) {
            continue;
         }
         const auto& compatible_symbols = alphabet.ambiguity_symbols.at(profile_symbol);
         std::vector<Symbol> difference_symbols;
         for (uint32_t symbol = 0; symbol < alphabet.count; ++symbol) {
            if (!containsSymbol(compatible_symbols, static_cast<Symbol>(symbol))) {
               difference_symbols.push_back(static_cast<Symbol>(symbol));
            }
         }
         if (difference_symbols.empty()) {
            continue;
         }
         difference_children.push_back(std::make_shared<SymbolInSet>(
            column, static_cast<uint32_t>(pos), std::move(difference_symbols)
         ));
      }
      auto at_least_distance_plus_one = std::make_shared<NOf>(
         std::move(difference_children), static_cast<int>(distance) + 1, false
      );
      return std::make_shared<Negation>(std::move(at_least_distance_plus_one));
   }
   std::unique_ptr<Operator> compile(const Table&) const override {
      throw QueryCompilationException(
         "MutationProfile expression must be eliminated in the query rewrite phase"
      );
   }
};

// ---------- boundary leaves ----------

class BitmapLeaf : public Expression {  // lineage_filter.cpp:77-100 and similar index lookups
  public:
   std::string name;
   explicit BitmapLeaf(std::string name) : name(std::move(name)) {}
   std::string toString() const override { return "bitmap:" + name; }
   ExprPtr rewrite(const Table&, AmbiguityMode) const override { return shared_from_this(); }
   std::unique_ptr<Operator> compile(const Table& table) const override {
      auto iter = table.named_bitmaps.find(name);
      if (iter == table.named_bitmaps.end()) {
         return std::make_unique<Empty>(table.row_layout);
      }
      return std::make_unique<IndexScan>(CowBitmap{&iter->second}, table.row_layout);
   }
};

class RangesLeaf : public Expression {  // date_between.cpp:75-79 on a sorted column
  public:
   std::vector<RangeSelection::Range> ranges;
   explicit RangesLeaf(std::vector<RangeSelection::Range> ranges) : ranges(std::move(ranges)) {}
   std::string toString() const override { return "ranges"; }
   ExprPtr rewrite(const Table&, AmbiguityMode) const override { return shared_from_this(); }
   std::unique_ptr<Operator> compile(const Table& table) const override {
      auto copy = ranges;
      return std::make_unique<RangeSelection>(std::move(copy), table.row_layout);
   }
};

// `column = 'value'` on an unindexed string column: Equals::rewrite -> StringInSet (equals.cpp:124-156), compiled to a
// Selection with the StringInSet predicate (string_in_set.cpp:73-83)
class StringEqualsExpr : public Expression {
  public:
   std::string column;
   std::string value;
   StringEqualsExpr(std::string column, std::string value) : column(std::move(column)), value(std::move(value)) {}
   std::string toString() const override { return column + " = '" + value + "'"; }
   ExprPtr rewrite(const Table&, AmbiguityMode) const override { return shared_from_this(); }
   std::unique_ptr<Operator> compile(const Table& table) const override {
      for (const StringValueColumn& candidate : table.string_columns) {
         if (candidate.name == column) {
            return std::make_unique<Selection>(
               std::make_unique<StringInSetPredicate>(&candidate, &table.chunk_begins, true, std::vector<std::string>{value}), table.row_layout
            );
         }
      }
      throw IllegalQueryException("The database does not contain the column '" + column + "'");
   }
};

// date_between.cpp:61-134
class DateBetweenExpr : public Expression {
  public:
   std::string column;
   std::optional<int32_t> date_from;
   std::optional<int32_t> date_to;
   DateBetweenExpr(std::string column, std::optional<int32_t> date_from, std::optional<int32_t> date_to)
       : column(std::move(column)), date_from(date_from), date_to(date_to) {}
   std::string toString() const override { return "[Date-between " + column + "]"; }
   ExprPtr rewrite(const Table&, AmbiguityMode) const override { return shared_from_this(); }
   std::unique_ptr<Operator> compile(const Table& table) const override {
      const DateValueColumn* date_column = nullptr;
      for (const DateValueColumn& candidate : table.date_columns) {
         if (candidate.name == column) {
            date_column = &candidate;
         }
      }
      if (date_column == nullptr) {
         throw IllegalQueryException("The database does not contain the column '" + column + "'");
      }
      const int32_t from = date_from.value_or(std::numeric_limits<int32_t>::min());
      if (date_column->is_sorted) {  // computeRangesOfSortedColumn :94-134
         std::vector<RangeSelection::Range> ranges;
         for (size_t chunk_idx = 0; chunk_idx < table.row_layout.numChunks(); ++chunk_idx) {
            const int32_t* begin = date_column->values.data() + table.chunk_begins[chunk_idx];
            const size_t chunk_size = table.row_layout.chunk_sizes[chunk_idx];
            const int32_t* end = begin + chunk_size;
            const auto lower_index = static_cast<size_t>(std::lower_bound(begin, end, from) - begin);
            const auto upper_index = date_to.has_value() ? static_cast<size_t>(std::upper_bound(begin, end, *date_to) - begin) : chunk_size;
            const auto chunk = static_cast<uint32_t>(chunk_idx);
            const uint32_t start_row = lower_index == chunk_size ? (chunk + 1) << 16 : (chunk << 16) | static_cast<uint32_t>(lower_index);
            const uint32_t end_row = upper_index == chunk_size ? (chunk + 1) << 16 : (chunk << 16) | static_cast<uint32_t>(upper_index);
            ranges.push_back({start_row, end_row});
         }
         return std::make_unique<RangeSelection>(std::move(ranges), table.row_layout);
      }
      PredicateVector predicates;
      predicates.push_back(std::make_unique<DateCompare>(date_column, &table.chunk_begins, DateCompare::Comparator::HIGHER_OR_EQUALS, from));
      predicates.push_back(std::make_unique<DateCompare>(
         date_column, &table.chunk_begins, DateCompare::Comparator::LESS_OR_EQUALS, date_to.value_or(std::numeric_limits<int32_t>::max())
      ));
      return std::make_unique<Selection>(std::nullopt, std::move(predicates), table.row_layout);
   }
};

// ---------- physical forms (operator-level known-answer tests) ----------

class IdsLeaf : public Expression {
  public:
   std::vector<uint32_t> ids;
   explicit IdsLeaf(std::vector<uint32_t> ids) : ids(std::move(ids)) {}
   std::string toString() const override { return "ids"; }
   ExprPtr rewrite(const Table&, AmbiguityMode) const override { return shared_from_this(); }
   std::unique_ptr<Operator> compile(const Table& table) const override {
      return std::make_unique<IndexScan>(
         CowBitmap{Roaring::fromIds(ids.data(), ids.size())}, table.row_layout
      );
   }
};

class CoveredLeaf : public Expression {
  public:
   std::string column;
   uint32_t position_idx;
   bool covered;
   CoveredLeaf(std::string column, uint32_t position_idx, bool covered)
       : column(std::move(column)),
         position_idx(position_idx),
         covered(covered) {}
   std::string toString() const override { return "covered"; }
   ExprPtr rewrite(const Table&, AmbiguityMode) const override { return shared_from_this(); }
   std::unique_ptr<Operator> compile(const Table& table) const override {
      const auto& sequence_column = requireColumn(table, column);
      return std::make_unique<Selection>(
         std::make_unique<IsInCoveredRegion>(
            &sequence_column.horizontal_coverage_index,
            position_idx,
            covered ? IsInCoveredRegion::Comparator::IS_COVERED
                    : IsInCoveredRegion::Comparator::IS_NOT_COVERED
         ),
         table.row_layout
      );
   }
};

class PhysicalOp : public Expression {
  public:
   enum Kind { AND, OR, NOT, THRESHOLD } kind;
   ExpressionVector first;
   ExpressionVector second;
   uint32_t number_of_matchers = 0;
   bool match_exactly = false;
   std::string toString() const override { return "physical"; }
   ExprPtr rewrite(const Table& table, AmbiguityMode mode) const override {
      auto copy = std::make_shared<PhysicalOp>(*this);
      for (auto& child : copy->first) {
         child = child->rewrite(table, mode);
      }
      for (auto& child : copy->second) {
         child = child->rewrite(table, mode);
      }
      return copy;
   }
   std::unique_ptr<Operator> compile(const Table& table) const override {
      OperatorVector first_ops;
      OperatorVector second_ops;
      for (const auto& child : first) {
         first_ops.push_back(child->compile(table));
      }
      for (const auto& child : second) {
         second_ops.push_back(child->compile(table));
      }
      switch (kind) {
         case AND:
            return std::make_unique<Intersection>(
               std::move(first_ops), std::move(second_ops), table.row_layout
            );
         case OR:
            return std::make_unique<Union>(std::move(first_ops), table.row_layout);
         case NOT:
            return std::make_unique<Complement>(std::move(first_ops.at(0)), table.row_layout);
         case THRESHOLD:
            return std::make_unique<Threshold>(
               std::move(first_ops),
               std::move(second_ops),
               number_of_matchers,
               match_exactly,
               table.row_layout
            );
      }
      throw std::runtime_error("unreachable");
   }
};

// ---------- s-expression reader ----------

struct SNode {
   bool is_atom = false;
   std::string atom;
   std::vector<SNode> items;
};

class SReader {
   const std::string& text;
   size_t pos = 0;

   void skipSpace() {
      while (pos < text.size() && std::isspace(static_cast<unsigned char>(text[pos])) != 0) {
         ++pos;
      }
   }

  public:
   explicit SReader(const std::string& text) : text(text) {}
   SNode read() {
      skipSpace();
      if (pos >= text.size()) {
         throw IllegalQueryException("filter expression ended unexpectedly");
      }
      SNode node;
      if (text[pos] == '(') {
         ++pos;
         while (true) {
            skipSpace();
            if (pos >= text.size()) {
               throw IllegalQueryException("filter expression: missing ')'");
            }
            if (text[pos] == ')') {
               ++pos;
               return node;
            }
            node.items.push_back(read());
         }
      }
      if (text[pos] == ')') {
         throw IllegalQueryException("filter expression: unexpected ')'");
      }
      node.is_atom = true;
      if (text[pos] == '"') {
         ++pos;
         while (pos < text.size() && text[pos] != '"') {
            node.atom.push_back(text[pos++]);
         }
         if (pos >= text.size()) {
            throw IllegalQueryException("filter expression: unterminated string");
         }
         ++pos;
         return node;
      }
      while (pos < text.size() && std::isspace(static_cast<unsigned char>(text[pos])) == 0 &&
             text[pos] != '(' && text[pos] != ')') {
         node.atom.push_back(text[pos++]);
      }
      return node;
   }
   bool atEnd() {
      skipSpace();
      return pos >= text.size();
   }
};

const std::string& atomOf(const SNode& node) {
   if (!node.is_atom) {
      throw IllegalQueryException("filter expression: expected an atom");
   }
   return node.atom;
}

uint64_t numberOf(const SNode& node) {
   const std::string& atom = atomOf(node);
   if (atom.empty() || !std::all_of(atom.begin(), atom.end(), [](char c) { return c >= '0' && c <= '9'; })) {
      throw IllegalQueryException("filter expression: expected a non-negative integer, got '" + atom + "'");
   }
   return std::stoull(atom);
}

uint32_t positionOf(const SNode& node) {
   const uint64_t position = numberOf(node);
   // ast_to_query.cpp: positions are 1-indexed in the query language
   CHECK_QUERY(position != 0, "The field 'position' is 1-indexed. Value of 0 not allowed.");
   return static_cast<uint32_t>(position - 1);
}

ExprPtr build(const SNode& node);

ExpressionVector buildAll(const SNode& list, size_t from = 0) {
   if (list.is_atom) {
      throw IllegalQueryException("filter expression: expected a list");
   }
   ExpressionVector result;
   for (size_t i = from; i < list.items.size(); ++i) {
      result.push_back(build(list.items[i]));
   }
   return result;
}

// Symbols are resolved lazily (the alphabet depends on the column), so the symbol-bearing forms
// keep the raw character and resolve it against both alphabets' shared rule: the table decides.
class DeferredSymbolExpression : public Expression {
  public:
   enum Kind { SYM_EQ, SYM_IN, PROFILE_MUTS } kind;
   std::string column;
   uint32_t position_idx = 0;
   std::string chars;  // SYM_EQ: one char or "."; SYM_IN: set
   uint32_t distance = 0;
   std::vector<std::pair<uint32_t, char>> mutations;
   std::string toString() const override { return "deferred"; }
   ExprPtr resolve(const Table& table) const {
      const auto& sequence_column = requireColumn(table, column);
      const Alphabet& alphabet = *sequence_column.alphabet;
      auto toSymbol = [&](char character) {
         const auto symbol = alphabet.charToSymbol(character);
         CHECK_QUERY(
            symbol.has_value(),
            "Invalid " + alphabet.symbol_name + " symbol '" + std::string(1, character) + "'"
         );
         return symbol.value();
      };
      if (kind == SYM_EQ) {
         if (chars == ".") {
            return std::make_shared<SymbolEquals>(column, position_idx, std::nullopt);
         }
         CHECK_QUERY(chars.size() == 1, "symbol must be a single character");
         return std::make_shared<SymbolEquals>(column, position_idx, toSymbol(chars[0]));
      }
      if (kind == SYM_IN) {
         std::vector<Symbol> symbols;
         for (char character : chars) {
            symbols.push_back(toSymbol(character));
         }
         return std::make_shared<RawSymbolInSet>(
            std::make_shared<SymbolInSet>(column, position_idx, std::move(symbols))
         );
      }
      ProfileMutations profile_mutations;
      for (const auto& [position, character] : mutations) {
         profile_mutations.mutations.emplace_back(position, toSymbol(character));
      }
      return std::make_shared<MutationProfile>(column, distance, std::move(profile_mutations));
   }
   ExprPtr rewrite(const Table& table, AmbiguityMode mode) const override {
      return resolve(table)->rewrite(table, mode);
   }
   std::unique_ptr<Operator> compile(const Table& table) const override {
      return resolve(table)->compile(table);
   }
};

ExprPtr build(const SNode& node) {
   if (node.is_atom || node.items.empty()) {
      throw IllegalQueryException("filter expression: expected a non-empty list");
   }
   const std::string& head = atomOf(node.items[0]);
   const auto& items = node.items;
   auto arity = [&](size_t count) {
      if (items.size() != count + 1) {
         throw IllegalQueryException("filter expression: wrong number of arguments for " + head);
      }
   };
   if (head == "true" || head == "false") {
      arity(0);
      return std::make_shared<BoolLiteral>(head == "true");
   }
   if (head == "sym-eq") {
      arity(3);
      auto expression = std::make_shared<DeferredSymbolExpression>();
      expression->kind = DeferredSymbolExpression::SYM_EQ;
      expression->column = atomOf(items[1]);
      expression->position_idx = positionOf(items[2]);
      expression->chars = atomOf(items[3]);
      return expression;
   }
   if (head == "sym-in") {
      arity(3);
      auto expression = std::make_shared<DeferredSymbolExpression>();
      expression->kind = DeferredSymbolExpression::SYM_IN;
      expression->column = atomOf(items[1]);
      expression->position_idx = positionOf(items[2]);
      expression->chars = atomOf(items[3]);
      return expression;
   }
   if (head == "has-mut") {
      arity(2);
      return std::make_shared<HasMutation>(atomOf(items[1]), positionOf(items[2]));
   }
   if (head == "and") {
      return std::make_shared<And>(buildAll(node, 1));
   }
   if (head == "or") {
      return std::make_shared<Or>(buildAll(node, 1));
   }
   if (head == "not") {
      arity(1);
      return std::make_shared<Negation>(build(items[1]));
   }
   if (head == "maybe") {
      arity(1);
      return std::make_shared<Maybe>(build(items[1]));
   }
   if (head == "exact") {
      arity(1);
      return std::make_shared<Exact>(build(items[1]));
   }
   if (head == "n-of") {
      if (items.size() < 3) {
         throw IllegalQueryException("filter expression: n-of needs K and EXACT");
      }
      return std::make_shared<NOf>(
         buildAll(node, 3), static_cast<int>(numberOf(items[1])), numberOf(items[2]) != 0
      );
   }
   if (head == "profile") {
      if (items.size() < 4) {
         throw IllegalQueryException("filter expression: profile needs COL DIST KIND ..");
      }
      const std::string& column = atomOf(items[1]);
      const auto distance = static_cast<uint32_t>(numberOf(items[2]));
      const std::string& kind = atomOf(items[3]);
      if (kind == "seq") {
         arity(4);
         return std::make_shared<MutationProfile>(column, distance, ProfileQuerySequence{atomOf(items[4])});
      }
      if (kind == "row") {
         arity(4);
         return std::make_shared<MutationProfile>(
            column, distance, ProfileRow{static_cast<uint32_t>(numberOf(items[4]))}
         );
      }
      if (kind == "muts") {
         if ((items.size() - 4) % 2 != 0) {
            throw IllegalQueryException("filter expression: profile muts needs POS SYM pairs");
         }
         auto expression = std::make_shared<DeferredSymbolExpression>();
         expression->kind = DeferredSymbolExpression::PROFILE_MUTS;
         expression->column = column;
         expression->distance = distance;
         for (size_t i = 4; i + 1 < items.size(); i += 2) {
            const std::string& symbol = atomOf(items[i + 1]);
            CHECK_QUERY(symbol.size() == 1, "symbol must be a single character");
            expression->mutations.emplace_back(positionOf(items[i]), symbol[0]);
         }
         return expression;
      }
      throw IllegalQueryException("filter expression: unknown profile kind " + kind);
   }
   if (head == "str-eq") {
      arity(2);
      return std::make_shared<StringEqualsExpr>(atomOf(items[1]), atomOf(items[2]));
   }
   if (head == "date-between") {
      arity(3);
      auto bound = [&](const auto& item) -> std::optional<int32_t> {
         const std::string bound_text = atomOf(item);
         if (bound_text == "*") {
            return std::nullopt;
         }
         return static_cast<int32_t>(std::stol(bound_text));
      };
      return std::make_shared<DateBetweenExpr>(atomOf(items[1]), bound(items[2]), bound(items[3]));
   }
   if (head == "bitmap") {
      arity(1);
      return std::make_shared<BitmapLeaf>(atomOf(items[1]));
   }
   if (head == "ranges") {
      if ((items.size() - 1) % 2 != 0) {
         throw IllegalQueryException("filter expression: ranges needs START END pairs");
      }
      std::vector<RangeSelection::Range> ranges;
      for (size_t i = 1; i + 1 < items.size(); i += 2) {
         ranges.push_back(
            {static_cast<uint32_t>(numberOf(items[i])), static_cast<uint32_t>(numberOf(items[i + 1]))}
         );
      }
      return std::make_shared<RangesLeaf>(std::move(ranges));
   }
   if (head == "ids") {
      std::vector<uint32_t> ids;
      for (size_t i = 1; i < items.size(); ++i) {
         ids.push_back(static_cast<uint32_t>(numberOf(items[i])));
      }
      return std::make_shared<IdsLeaf>(std::move(ids));
   }
   if (head == "covered" || head == "not-covered") {
      arity(2);
      return std::make_shared<CoveredLeaf>(atomOf(items[1]), positionOf(items[2]), head == "covered");
   }
   if (head == "op-and" || head == "op-threshold") {
      auto expression = std::make_shared<PhysicalOp>();
      size_t lists_from = 1;
      if (head == "op-threshold") {
         expression->kind = PhysicalOp::THRESHOLD;
         if (items.size() != 5) {
            throw IllegalQueryException("filter expression: op-threshold K EXACT (pos..) (neg..)");
         }
         expression->number_of_matchers = static_cast<uint32_t>(numberOf(items[1]));
         expression->match_exactly = numberOf(items[2]) != 0;
         lists_from = 3;
      } else {
         expression->kind = PhysicalOp::AND;
         arity(2);
      }
      expression->first = buildAll(items[lists_from]);
      expression->second = buildAll(items[lists_from + 1]);
      return expression;
   }
   if (head == "op-or") {
      auto expression = std::make_shared<PhysicalOp>();
      expression->kind = PhysicalOp::OR;
      expression->first = buildAll(node, 1);
      return expression;
   }
   if (head == "op-not") {
      arity(1);
      auto expression = std::make_shared<PhysicalOp>();
      expression->kind = PhysicalOp::NOT;
      expression->first.push_back(build(items[1]));
      return expression;
   }
   throw IllegalQueryException("filter expression: unknown form '" + head + "'");
}

}  // namespace

ExprPtr parseExpression(const std::string& text) {
   SReader reader(text);
   const SNode node = reader.read();
   if (!reader.atEnd()) {
      throw IllegalQueryException("filter expression: trailing input");
   }
   return build(node);
}

CowBitmap computeFilter(const Expression& filter, const Table& table) {
   auto rewritten_filter = filter.rewrite(table, AmbiguityMode::NONE);
   auto compiled_filter = rewritten_filter->compile(table);
   return compiled_filter->evaluate();
}

// ---------- compileSymbolInSet (symbol_in_set.cpp:67-264) ----------

namespace {

std::unique_ptr<Operator> makeDifference(
   std::unique_ptr<Operator> left,
   std::unique_ptr<Operator> right,
   const RowLayout& row_layout
) {
   OperatorVector non_negated_operators;
   non_negated_operators.push_back(std::move(left));
   OperatorVector negated_operators;
   negated_operators.push_back(std::move(right));
   return std::make_unique<Intersection>(
      std::move(non_negated_operators), std::move(negated_operators), row_layout
   );
}

std::unique_ptr<Operator> excludeNullSequences(
   std::unique_ptr<Operator> operator_,
   const SequenceColumn& sequence_column,
   const RowLayout& row_layout
) {
   if (sequence_column.null_bitmap.isEmpty()) {
      return operator_;
   }
   return makeDifference(
      std::move(operator_),
      std::make_unique<IndexScan>(CowBitmap{&sequence_column.null_bitmap}, row_layout),
      row_layout
   );
}

std::vector<Symbol> negateSymbols(
   const Alphabet& alphabet,
   const std::vector<Symbol>& symbols,
   std::optional<Symbol> excluded
) {
   std::vector<Symbol> result;
   for (uint32_t symbol = 0; symbol < alphabet.count; ++symbol) {
      if (excluded.has_value() && symbol == excluded.value()) {
         continue;
      }
      if (!containsSymbol(symbols, static_cast<Symbol>(symbol))) {
         result.push_back(static_cast<Symbol>(symbol));
      }
   }
   return result;
}

}  // namespace

std::unique_ptr<Operator> compileSymbolInSet(
   const SequenceColumn& sequence_column,
   uint32_t position_idx,
   const std::vector<Symbol>& symbols,
   const RowLayout& row_layout
) {
   const Alphabet& alphabet = *sequence_column.alphabet;
   CHECK_QUERY(
      position_idx < sequence_column.reference_sequence.size(),
      "SymbolInSet<" + alphabet.symbol_name + "> position is out of bounds " +
         std::to_string(position_idx + 1) + " > " +
         std::to_string(sequence_column.reference_sequence.size())
   );
   const Symbol local_reference_symbol = sequence_column.getLocalReferencePosition(position_idx);
   const bool includes_reference = containsSymbol(symbols, local_reference_symbol);
   const bool includes_missing_symbol = containsSymbol(symbols, alphabet.missing);
   const auto& index = sequence_column.vertical_sequence_index;

   if (includes_reference && includes_missing_symbol) {
      auto bitmap = CowBitmap::fromContainerViews(
         index.getMatchingContainerViews(position_idx, negateSymbols(alphabet, symbols, std::nullopt))
      );
      return excludeNullSequences(
         std::make_unique<Complement>(
            std::make_unique<IndexScan>(std::move(bitmap), row_layout), row_layout
         ),
         sequence_column,
         row_layout
      );
   }
   if (includes_missing_symbol) {
      auto bitmap = CowBitmap::fromContainerViews(index.getMatchingContainerViews(position_idx, symbols));
      OperatorVector operators_for_union;
      operators_for_union.push_back(std::make_unique<Selection>(
         std::make_unique<IsInCoveredRegion>(
            &sequence_column.horizontal_coverage_index,
            position_idx,
            IsInCoveredRegion::Comparator::IS_NOT_COVERED
         ),
         row_layout
      ));
      operators_for_union.push_back(std::make_unique<IndexScan>(std::move(bitmap), row_layout));
      return excludeNullSequences(
         std::make_unique<Union>(std::move(operators_for_union), row_layout),
         sequence_column,
         row_layout
      );
   }
   if (includes_reference) {
      auto bitmap = CowBitmap::fromContainerViews(index.getMatchingContainerViews(
         position_idx, negateSymbols(alphabet, symbols, alphabet.missing)
      ));
      return makeDifference(
         std::make_unique<Selection>(
            std::make_unique<IsInCoveredRegion>(
               &sequence_column.horizontal_coverage_index,
               position_idx,
               IsInCoveredRegion::Comparator::IS_COVERED
            ),
            row_layout
         ),
         std::make_unique<IndexScan>(std::move(bitmap), row_layout),
         row_layout
      );
   }
   auto bitmap = CowBitmap::fromContainerViews(index.getMatchingContainerViews(position_idx, symbols));
   return std::make_unique<IndexScan>(std::move(bitmap), row_layout);
}

}  // namespace oracle
