#include "expressions.h"

#include <algorithm>
#include <cctype>
#include <map>
#include <string_view>
#include <utility>

#include "roaring_writer.h"

namespace silo_host {

namespace {

// CHECK_SILO_QUERY (query_engine/illegal_query_exception.h:8): the message is only built when the check
// fails -- a MutationProfile runs one check per genome position
#define CHECK_QUERY(condition, message)            \
   do {                                            \
      if (!(condition)) [[unlikely]] {             \
         throw IllegalQueryException(message);     \
      }                                            \
   } while (false)

const SequenceColumnInfo& requireColumn(const Table& table, const std::string& name) {
   const SequenceColumnInfo* column = table.findColumn(name);
   // validateSequenceName, query_engine/query_parse_sequence_name.h:10-20
   CHECK_QUERY(column != nullptr, "Database does not contain the Sequence with name: '" + name + "'");
   return *column;
}

uint32_t maskOf(const std::vector<Symbol>& symbols) {
   uint32_t mask = 0;
   for (Symbol symbol : symbols) {
      mask |= 1u << symbol;
   }
   return mask;
}

uint32_t allSymbolsMask(const Alphabet& alphabet) {
   return alphabet.count() == 32 ? 0xFFFFFFFFu : (1u << alphabet.count()) - 1u;
}

Symbol toSymbol(const Alphabet& alphabet, char character) {
   const auto symbol = alphabet.charToSymbol(character);
   CHECK_QUERY(
      symbol.has_value(), "Invalid " + alphabet.symbol_name + " symbol '" + std::string(1, character) + "'"
   );
   return symbol.value();
}

}  // namespace

AmbiguityMode invertMode(AmbiguityMode mode) {
   switch (mode) {
      case AmbiguityMode::UPPER_BOUND:
         return AmbiguityMode::LOWER_BOUND;
      case AmbiguityMode::LOWER_BOUND:
         return AmbiguityMode::UPPER_BOUND;
      case AmbiguityMode::NONE:
         return AmbiguityMode::NONE;
   }
   return mode;
}

// ---- literals --------------------------------------------------------------------------------

ExpressionPtr BoolLiteral::rewrite(const Table&, AmbiguityMode) const {
   return std::make_shared<BoolLiteral>(value);
}

std::unique_ptr<Operator> BoolLiteral::compile(const Table&) const {
   if (value) {
      return std::make_unique<Full>();
   }
   return std::make_unique<Empty>();
}

// ---- SymbolInSet -----------------------------------------------------------------------------

std::string SymbolInSet::toString() const {  // symbol_in_set.cpp:37-51
   std::string symbols_string;
   for (uint32_t symbol = 0; symbol < 32; ++symbol) {
      if (((symbols.mask >> symbol) & 1u) == 0) {
         continue;
      }
      if (!symbols_string.empty()) {
         symbols_string += ", ";
      }
      symbols_string += alphabet != nullptr ? std::string(1, alphabet->symbolToChar(static_cast<Symbol>(symbol)))
                                            : "#" + std::to_string(symbol);
   }
   return "(" + column + ":symbol at position " + std::to_string(position_idx + 1) + " in {" + symbols_string + "})";
}

ExpressionPtr SymbolInSet::rewrite(const Table&, AmbiguityMode) const {
   throw QueryCompilationException(
      "Cannot rewrite SymbolInSet - this expression should only be created during query rewrites "
      "and not directly used"
   );
}

std::unique_ptr<Operator> SymbolInSet::compile(const Table& table) const {
   return compileSymbolInSet(requireColumn(table, column), position_idx, symbols);
}

std::unique_ptr<Operator> compileSymbolInSet(
   const SequenceColumnInfo& sequence_column,
   uint32_t position_idx,
   SymbolSet symbols
) {
   const Alphabet& alphabet = *sequence_column.alphabet;
   CHECK_QUERY(
      position_idx < sequence_column.reference_sequence.size(),
      "SymbolInSet<" + alphabet.symbol_name + "> position is out of bounds " +
         std::to_string(position_idx + 1) + " > " + std::to_string(sequence_column.reference_sequence.size())
   );
   const int column = sequence_column.device_column;
   const uint32_t requested = symbols.mask;
   const uint32_t reference_bit = 1u << sequence_column.local_reference.at(position_idx);
   const uint32_t missing_bit = 1u << alphabet.missing;
   const bool includes_reference = (requested & reference_bit) != 0;
   const bool includes_missing = (requested & missing_bit) != 0;

   // rows whose sequence is null are stored with an empty covered region; the compilations that
   // start from "not covered" must drop them again (symbol_in_set.cpp:80-98)
   auto excludeNullSequences = [&](std::unique_ptr<Operator> op) -> std::unique_ptr<Operator> {
      if (!sequence_column.has_null_rows) {
         return op;
      }
      OperatorVector keep;
      keep.push_back(std::move(op));
      OperatorVector drop;
      drop.push_back(IndexScan::overNulls(column));
      return std::make_unique<Intersection>(std::move(keep), std::move(drop));
   };

   if (includes_reference && includes_missing) {
      // everything except the rows that hold one of the symbols that were NOT asked for
      const uint32_t others = allSymbolsMask(alphabet) & ~requested;
      return excludeNullSequences(std::make_unique<Complement>(IndexScan::overSymbols(column, position_idx, others)));
   }
   if (includes_missing) {
      OperatorVector alternatives;
      alternatives.push_back(std::make_unique<Selection>(CoveragePredicate{column, position_idx, false}));
      alternatives.push_back(IndexScan::overSymbols(column, position_idx, requested));
      return excludeNullSequences(std::make_unique<Union>(std::move(alternatives)));
   }
   if (includes_reference) {
      // covered rows minus the ones that hold a non-missing symbol outside the request
      const uint32_t others = allSymbolsMask(alphabet) & ~requested & ~missing_bit;
      OperatorVector keep;
      keep.push_back(std::make_unique<Selection>(CoveragePredicate{column, position_idx, true}));
      OperatorVector drop;
      drop.push_back(IndexScan::overSymbols(column, position_idx, others));
      return std::make_unique<Intersection>(std::move(keep), std::move(drop));
   }
   return IndexScan::overSymbols(column, position_idx, requested);
}

// ---- SymbolEquals / HasMutation --------------------------------------------------------------

std::string SymbolEquals::toString() const {
   return column + ":" + std::to_string(position_idx + 1) + std::string(1, symbol.value_or('.'));
}

ExpressionPtr SymbolEquals::rewrite(const Table& table, AmbiguityMode mode) const {
   const auto& sequence_column = requireColumn(table, column);
   const Alphabet& alphabet = *sequence_column.alphabet;
   CHECK_QUERY(
      position_idx < sequence_column.reference_sequence.size(),
      "SymbolEquals<" + alphabet.symbol_name + "> position is out of bounds " +
         std::to_string(position_idx + 1) + " > " + std::to_string(sequence_column.reference_sequence.size())
   );
   const Symbol wanted = symbol.has_value() ? toSymbol(alphabet, symbol.value())
                                            : sequence_column.reference_sequence.at(position_idx);
   if (mode == AmbiguityMode::UPPER_BOUND) {
      return std::make_shared<SymbolInSet>(column, position_idx, alphabet.ambiguity_symbols.at(wanted), &alphabet);
   }
   return std::make_shared<SymbolInSet>(column, position_idx, std::vector<Symbol>{wanted}, &alphabet);
}

std::unique_ptr<Operator> SymbolEquals::compile(const Table&) const {
   throw QueryCompilationException("SymbolEquals should have been rewritten before compilation");
}

std::string HasMutation::toString() const {
   return column + ":" + std::to_string(position_idx);
}

ExpressionPtr HasMutation::rewrite(const Table& table, AmbiguityMode mode) const {
   const auto& sequence_column = requireColumn(table, column);
   const Alphabet& alphabet = *sequence_column.alphabet;
   CHECK_QUERY(
      position_idx < sequence_column.reference_sequence.size(),
      "Has" + alphabet.symbol_name + "Mutation position is out of bounds " + std::to_string(position_idx + 1) +
         " > " + std::to_string(sequence_column.reference_sequence.size())
   );
   const Symbol reference_symbol = sequence_column.reference_sequence.at(position_idx);
   uint32_t mask = allSymbolsMask(alphabet);
   if (mode == AmbiguityMode::UPPER_BOUND) {
      mask &= ~(1u << reference_symbol);
   } else {
      mask &= ~maskOf(alphabet.ambiguity_symbols.at(reference_symbol));
   }
   return std::make_shared<SymbolInSet>(column, position_idx, SymbolSet(mask), &alphabet);
}

std::unique_ptr<Operator> HasMutation::compile(const Table&) const {
   throw QueryCompilationException("HasMutation expression must be eliminated in query rewrite phase");
}

// ---- Negation / Maybe / Exact ----------------------------------------------------------------

ExpressionPtr Negation::rewrite(const Table& table, AmbiguityMode mode) const {
   return std::make_shared<Negation>(child->rewrite(table, invertMode(mode)));
}
std::unique_ptr<Operator> Negation::compile(const Table& table) const {
   return Operator::negate(child->compile(table));
}
ExpressionPtr Maybe::rewrite(const Table& table, AmbiguityMode) const {
   return child->rewrite(table, AmbiguityMode::UPPER_BOUND);
}
std::unique_ptr<Operator> Maybe::compile(const Table&) const {
   throw QueryCompilationException("Maybe expression must be elimitated in query rewrite phase");
}
ExpressionPtr Exact::rewrite(const Table& table, AmbiguityMode) const {
   return child->rewrite(table, AmbiguityMode::LOWER_BOUND);
}
std::unique_ptr<Operator> Exact::compile(const Table&) const {
   throw QueryCompilationException("Exact expression must be elimitated in query rewrite phase");
}

// ---- And -------------------------------------------------------------------------------------

std::string And::toString() const {  // and.cpp:32-37
   return "And(" + joinWithLimit(children, " & ") + ")";
}

ExpressionPtr And::rewrite(const Table& table, AmbiguityMode mode) const {
   ExpressionVector rewritten;
   rewritten.reserve(children.size());
   for (const auto& child : children) {
      rewritten.push_back(child->rewrite(table, mode));
   }
   return std::make_shared<And>(std::move(rewritten));
}

std::unique_ptr<Operator> And::compile(const Table& table) const {
   // and.cpp:91-143: flatten nested intersections, un-negate complements, dissolve selections into a
   // predicate list wrapped around the index-arithmetic child
   OperatorVector pending;
   for (const auto& child : children) {
      pending.push_back(child->compile(table));
   }
   OperatorVector non_negated;
   OperatorVector negated;
   std::vector<CoveragePredicate> predicates;
   while (!pending.empty()) {
      auto child = std::move(pending.back());
      pending.pop_back();
      switch (child->type()) {
         case FULL:
            break;
         case EMPTY:
            return std::make_unique<Empty>();  // and.cpp:150-166 with the shortcut of :103-108
         case INTERSECTION: {
            auto* nested = static_cast<Intersection*>(child.get());
            for (auto& grandchild : nested->children) {
               non_negated.push_back(std::move(grandchild));
            }
            for (auto& grandchild : nested->negated_children) {
               negated.push_back(std::move(grandchild));
            }
            break;
         }
         case COMPLEMENT:
            negated.push_back(Operator::negate(std::move(child)));
            break;
         case SELECTION: {
            auto* selection = static_cast<Selection*>(child.get());
            predicates.insert(predicates.end(), selection->predicates.begin(), selection->predicates.end());
            if (selection->child_operator.has_value()) {
               pending.push_back(std::move(selection->child_operator.value()));
            }
            break;
         }
         default:
            non_negated.push_back(std::move(child));
      }
   }
   if (non_negated.empty() && negated.empty()) {
      if (predicates.empty()) {
         return std::make_unique<Full>();
      }
      return std::make_unique<Selection>(std::nullopt, std::move(predicates));
   }
   std::unique_ptr<Operator> index_arithmetic;
   if (non_negated.size() == 1 && negated.empty()) {
      index_arithmetic = std::move(non_negated[0]);
   } else if (negated.size() == 1 && non_negated.empty()) {
      index_arithmetic = std::make_unique<Complement>(std::move(negated[0]));
   } else if (non_negated.empty()) {
      index_arithmetic = std::make_unique<Complement>(std::make_unique<Union>(std::move(negated)));
   } else {
      index_arithmetic = std::make_unique<Intersection>(std::move(non_negated), std::move(negated));
   }
   if (predicates.empty()) {
      return index_arithmetic;
   }
   return std::make_unique<Selection>(
      std::optional<std::unique_ptr<Operator>>{std::move(index_arithmetic)}, std::move(predicates)
   );
}

// ---- Or --------------------------------------------------------------------------------------

std::string Or::toString() const {  // or.cpp:27-32
   return "Or(" + joinWithLimit(children, " | ") + ")";
}

ExpressionPtr Or::rewrite(const Table& table, AmbiguityMode mode) const {
   // or.cpp:47-68: flatten nested disjunctions before rewriting
   std::vector<const ScalarExpression*> flat;
   std::vector<const ScalarExpression*> stack;
   for (const auto& child : children) {
      stack.push_back(child.get());
   }
   while (!stack.empty()) {
      const ScalarExpression* current = stack.back();
      stack.pop_back();
      if (const auto* nested = dynamic_cast<const Or*>(current)) {
         for (const auto& child : nested->children) {
            stack.push_back(child.get());
         }
      } else {
         flat.push_back(current);
      }
   }
   ExpressionVector rewritten;
   for (const ScalarExpression* child : flat) {
      rewritten.push_back(child->rewrite(table, mode));
   }
   // or.cpp:70-95: drop constant false, short-circuit on constant true, flatten again
   ExpressionVector simplified;
   while (!rewritten.empty()) {
      ExpressionPtr child = std::move(rewritten.back());
      rewritten.pop_back();
      if (const auto* literal = dynamic_cast<const BoolLiteral*>(child.get())) {
         if (literal->value) {
            simplified.clear();
            simplified.push_back(std::make_shared<BoolLiteral>(true));
            rewritten.clear();
            break;
         }
         continue;
      }
      if (const auto* nested = dynamic_cast<const Or*>(child.get())) {
         rewritten.insert(rewritten.end(), nested->children.begin(), nested->children.end());
      } else {
         simplified.push_back(std::move(child));
      }
   }
   // or.cpp:97-124: SymbolInSet children on the same (column, position) merge into one set
   ExpressionVector merged_children;
   std::map<std::pair<std::string, uint32_t>, SymbolSet> merged;
   std::vector<std::pair<std::string, uint32_t>> merge_order_nucleotide;
   std::vector<std::pair<std::string, uint32_t>> merge_order_amino_acid;
   for (auto& child : simplified) {
      if (const auto* in_set = dynamic_cast<const SymbolInSet*>(child.get())) {
         const auto key = std::make_pair(in_set->column, in_set->position_idx);
         merged[key].mask |= in_set->symbols.mask;
      } else {
         merged_children.push_back(std::move(child));
      }
   }
   // the reference runs the nucleotide pass before the amino-acid pass; each appends its merged
   // sets in map order
   for (int pass = 0; pass < 2; ++pass) {
      for (auto& [key, symbols] : merged) {
         const SequenceColumnInfo* column = table.findColumn(key.first);
         const bool is_nucleotide = column != nullptr && column->alphabet == &Alphabet::nucleotide();
         if ((pass == 0) == is_nucleotide) {
            merged_children.push_back(std::make_shared<SymbolInSet>(
               key.first, key.second, symbols, column != nullptr ? column->alphabet : nullptr
            ));
         }
      }
   }
   if (merged_children.size() == 1) {
      return merged_children[0];
   }
   return std::make_shared<Or>(std::move(merged_children));
}

std::unique_ptr<Operator> Or::compile(const Table& table) const {
   OperatorVector kept;
   for (const auto& child_expression : children) {
      auto child = child_expression->compile(table);
      if (child->type() == EMPTY) {
         continue;
      }
      if (child->type() == FULL) {
         return std::make_unique<Full>();
      }
      if (child->type() == UNION) {
         for (auto& grandchild : static_cast<Union*>(child.get())->children) {
            kept.push_back(std::move(grandchild));
         }
      } else {
         kept.push_back(std::move(child));
      }
   }
   if (kept.empty()) {
      return std::make_unique<Empty>();
   }
   if (kept.size() == 1) {
      return std::move(kept[0]);
   }
   const bool any_complement =
      std::any_of(kept.begin(), kept.end(), [](const auto& child) { return child->type() == COMPLEMENT; });
   if (any_complement) {
      return Complement::fromDeMorgan(std::move(kept));
   }
   return std::make_unique<Union>(std::move(kept));
}

// ---- NOf -------------------------------------------------------------------------------------

std::string NOf::toString() const {  // nof.cpp:161-171
   std::string joined = joinWithLimit(children);
   if (span != nullptr) {  // joinWithLimit over children followed by the span's children
      constexpr size_t LIMIT = 10;
      const size_t total = children.size() + span->size();
      joined.clear();
      size_t printed = 0;
      for (; printed < std::min(children.size(), LIMIT); ++printed) {
         joined += (printed > 0 ? ", " : "") + children[printed]->toString();
      }
      for (size_t i = 0; printed < std::min(total, LIMIT); ++printed, ++i) {
         joined += (printed > 0 ? ", " : "") +
                   SymbolInSet(span->column, span->positions[i], SymbolSet(span->masks[i]), span->alphabet).toString();
      }
      if (total > printed) {
         joined += ", ... (" + std::to_string(total - printed) + " more)";
      }
   }
   return std::string(match_exactly ? "[exactly-" : "[") + std::to_string(number_of_matchers) + "-of:" + joined + "]";
}

ExpressionPtr NOf::rewrite(const Table& table, AmbiguityMode mode) const {
   if (span != nullptr) {  // (its children are SymbolInSets: symbol_in_set.cpp rewrite throws)
      throw QueryCompilationException(
         "Cannot rewrite SymbolInSet - this expression should only be created during query rewrites "
         "and not directly used"
      );
   }
   auto rewriteChildren = [&]() {
      ExpressionVector rewritten;
      rewritten.reserve(children.size());
      for (const auto& child : children) {
         rewritten.push_back(child->rewrite(table, mode));
      }
      return rewritten;
   };
   // an exact count cannot carry an ambiguity mode: exactly-k = at-least-k and not at-least-(k+1)
   // (nof.cpp:219-248)
   if (mode != AmbiguityMode::NONE && match_exactly && std::cmp_less(number_of_matchers, children.size())) {
      ExpressionVector both;
      both.push_back(std::make_shared<NOf>(rewriteChildren(), number_of_matchers, false));
      both.push_back(std::make_shared<Negation>(std::make_shared<NOf>(rewriteChildren(), number_of_matchers + 1, false)));
      return std::make_shared<And>(std::move(both));
   }
   return std::make_shared<NOf>(rewriteChildren(), number_of_matchers, match_exactly);
}

std::unique_ptr<Operator> NOf::compile(const Table& table) const {
   // nof.cpp:184-215: Empty children can never match, Full children always do, complements are
   // kept as negated children
   OperatorVector non_negated;
   OperatorVector negated;
   non_negated.reserve(children.size());
   int k = number_of_matchers;
   auto addChild = [&](std::unique_ptr<Operator> child) {
      if (child->type() == EMPTY) {
         return;
      }
      if (child->type() == FULL) {
         --k;
      } else if (child->type() == COMPLEMENT) {
         negated.push_back(Operator::negate(std::move(child)));
      } else {
         non_negated.push_back(std::move(child));
      }
   };
   for (const auto& child_expression : children) {
      addChild(child_expression->compile(table));
   }
   // The compact children: what compileSymbolInSet would build for each (symbol_in_set.cpp:231-264), kept compact
   // when that is an IndexScan or Selection[IsCovered] minus IndexScan (never Empty / Full / Complement); a symbol
   // set that holds the missing symbol is compiled the ordinary way.
   std::shared_ptr<SymbolScanSpan> scan_span;
   if (span != nullptr) {
      const auto& sequence_column = requireColumn(table, span->column);
      const Alphabet& alphabet = *sequence_column.alphabet;
      scan_span = std::make_shared<SymbolScanSpan>();
      scan_span->device_column = sequence_column.device_column;
      scan_span->column_name = span->column;
      scan_span->all_symbols_mask = allSymbolsMask(alphabet);
      scan_span->missing_bit = 1u << alphabet.missing;
      scan_span->leaves.reserve(span->size());
      const size_t reference_length = sequence_column.reference_sequence.size();
      for (size_t i = 0; i < span->size(); ++i) {
         const uint32_t position_idx = span->positions[i];
         const uint32_t requested = span->masks[i];
         if (position_idx >= reference_length || (requested & scan_span->missing_bit) != 0) {
            addChild(compileSymbolInSet(sequence_column, position_idx, SymbolSet(requested)));  // (throws when out of bounds)
            continue;
         }
         const bool includes_reference = ((requested >> sequence_column.local_reference[position_idx]) & 1u) != 0;
         scan_span->leaves.push_back({position_idx, requested, includes_reference});
      }
      if (scan_span->leaves.empty()) {
         scan_span = nullptr;
      }
   }
   const size_t span_children = scan_span != nullptr ? scan_span->size() : 0;
   const int n = static_cast<int>(non_negated.size() + negated.size() + span_children);
   // only the wide forms (Threshold, Union) take the compact children; every other outcome gets operators
   const bool general_threshold = k > 1 && k < n;
   const bool wide_union = k == 1 && n > 1 && !match_exactly && negated.empty();
   if (scan_span != nullptr && !(general_threshold || wide_union) && k <= n && k >= 0) {
      for (size_t i = 0; i < scan_span->size(); ++i) {
         non_negated.push_back(scan_span->materialise(i));
      }
      scan_span = nullptr;
   }
   // nof.cpp:33-84 trivial cases
   if (k > n) {
      return std::make_unique<Empty>();
   }
   if (k < 0) {
      if (match_exactly) {
         return std::make_unique<Empty>();
      }
      return std::make_unique<Full>();
   }
   if (k == 0) {
      if (!match_exactly || n == 0) {
         return std::make_unique<Full>();
      }
      if (n == 1) {
         if (non_negated.empty()) {
            return std::move(negated[0]);
         }
         return std::make_unique<Complement>(std::move(non_negated[0]));
      }
      if (negated.empty()) {
         return std::make_unique<Complement>(std::make_unique<Union>(std::move(non_negated)));
      }
      return std::make_unique<Intersection>(std::move(negated), std::move(non_negated));
   }
   if (k == 1 && n == 1) {
      if (negated.empty()) {
         return std::move(non_negated[0]);
      }
      return std::make_unique<Complement>(std::move(negated[0]));
   }
   if (k == n) {  // nof.cpp:86-98: all must match
      if (non_negated.empty()) {
         return std::make_unique<Complement>(std::make_unique<Union>(std::move(negated)));
      }
      return std::make_unique<Intersection>(std::move(non_negated), std::move(negated));
   }
   if (k == 1 && !match_exactly) {  // nof.cpp:100-114: any may match
      if (negated.empty()) {
         return std::make_unique<Union>(std::move(non_negated), std::move(scan_span));
      }
      return std::make_unique<Complement>(std::make_unique<Intersection>(std::move(negated), std::move(non_negated)));
   }
   return std::make_unique<Threshold>(std::move(non_negated), std::move(negated), static_cast<uint32_t>(k), match_exactly, std::move(scan_span));
}

// ---- MutationProfile -------------------------------------------------------------------------

std::string MutationProfile::toString() const {  // mutation_profile.cpp:39-55
   std::string input_string;
   if (const auto* query = std::get_if<QuerySequence>(&input)) {
      input_string = "querySequence=" + query->sequence.substr(0, 20) + "...";
   } else {
      input_string = "mutations(count=" + std::to_string(std::get<Mutations>(input).mutations.size()) + ")";
   }
   return "MutationProfile(" + column + ":distance=" + std::to_string(distance) + "," + input_string + ")";
}

ExpressionPtr MutationProfile::rewrite(const Table& table, AmbiguityMode) const {
   const auto& sequence_column = requireColumn(table, column);
   const Alphabet& alphabet = *sequence_column.alphabet;
   const size_t reference_length = sequence_column.reference_sequence.size();
   std::vector<Symbol> profile;
   if (const auto* query = std::get_if<QuerySequence>(&input)) {
      CHECK_QUERY(
         query->sequence.size() == reference_length,
         "querySequence length " + std::to_string(query->sequence.size()) +
            " does not match the reference sequence length " + std::to_string(reference_length) + " for " +
            alphabet.symbol_name + " MutationProfile"
      );
      profile.reserve(reference_length);
      for (char character : query->sequence) {
         const auto symbol = alphabet.charToSymbol(character);
         CHECK_QUERY(
            symbol.has_value(),
            "Invalid " + alphabet.symbol_name + " symbol '" + std::string(1, character) +
               "' in querySequence for MutationProfile"
         );
         profile.push_back(symbol.value());
      }
   } else {
      profile = sequence_column.reference_sequence;
      for (const auto& [position, character] : std::get<Mutations>(input).mutations) {
         CHECK_QUERY(
            position < reference_length,
            alphabet.symbol_name + " MutationProfile mutation position " + std::to_string(position + 1) +
               " is out of bounds (reference length " + std::to_string(reference_length) + ")"
         );
         profile[position] = toSymbol(alphabet, character);
      }
   }
   // one "definitely different" child per position: the symbols that are NOT compatible with the
   // profile symbol (mutation_profile.cpp:222-247); the filter is "fewer than distance+1 differ"
   // (the children in compact form: one (position, symbol set) pair each, see SymbolInSetSpan)
   auto differences = std::make_shared<SymbolInSetSpan>();
   differences->column = column;
   differences->alphabet = &alphabet;
   differences->positions.reserve(profile.size());
   differences->masks.reserve(profile.size());
   const uint32_t all_symbols = allSymbolsMask(alphabet);
   std::vector<uint32_t> incompatible_with(alphabet.ambiguity_symbols.size());
   for (size_t symbol = 0; symbol < incompatible_with.size(); ++symbol) {
      incompatible_with[symbol] = all_symbols & ~maskOf(alphabet.ambiguity_symbols[symbol]);
   }
   for (size_t position = 0; position < profile.size(); ++position) {
      if (profile[position] == alphabet.missing) {
         continue;
      }
      const uint32_t incompatible = incompatible_with.at(profile[position]);
      if (incompatible == 0) {
         continue;
      }
      differences->positions.push_back(static_cast<uint32_t>(position));
      differences->masks.push_back(incompatible);
   }
   return std::make_shared<Negation>(
      std::make_shared<NOf>(ExpressionVector{}, static_cast<int>(distance) + 1, false, std::move(differences))
   );
}

std::unique_ptr<Operator> MutationProfile::compile(const Table&) const {
   throw QueryCompilationException("MutationProfile expression must be eliminated in the query rewrite phase");
}

// ---- predicates over value columns -------------------------------------------------------------

std::string StringEquals::toString() const {
   return column + " = '" + value + "'";
}

ExpressionPtr StringEquals::rewrite(const Table&, AmbiguityMode) const {
   return shared_from_this();
}

std::unique_ptr<Operator> StringEquals::compile(const Table& table) const {
   const ValueColumnInfo* found = table.findValueColumn(column);
   CHECK_QUERY(found != nullptr, "The database does not contain the column '" + column + "'");
   CHECK_QUERY(found->type == ValueColumnInfo::Type::STRING, "The column '" + column + "' is not of type string");
   CoveragePredicate predicate{};
   predicate.kind = CoveragePredicate::COMPARE;
   predicate.value_column = found->device_column;
   predicate.comparator = SILO_CMP_IN_SET;
   const auto id = found->dictionary.find(value);
   if (id != found->dictionary.end()) {
      predicate.set.push_back(id->second);  // (a value the column does not hold leaves the set empty: no row matches)
   }
   predicate.display = "$string " + column + " IN ['" + value + "']";
   return std::make_unique<Selection>(predicate);
}

std::string DateBetween::toString() const {
   return "[Date-between " + column + " " + (date_from.has_value() ? std::to_string(*date_from) : "unbounded") + " and " +
          (date_to.has_value() ? std::to_string(*date_to) : "unbounded") + "]";
}

ExpressionPtr DateBetween::rewrite(const Table&, AmbiguityMode) const {
   return shared_from_this();
}

std::unique_ptr<Operator> DateBetween::compile(const Table& table) const {
   const ValueColumnInfo* found = table.findValueColumn(column);
   CHECK_QUERY(found != nullptr, "The database does not contain the column '" + column + "'");
   CHECK_QUERY(found->type == ValueColumnInfo::Type::DATE, "The column '" + column + "' is not of type date");
   const int32_t from = date_from.value_or(std::numeric_limits<int32_t>::min());
   if (found->sorted) {
      // computeRangesOfSortedColumn (date_between.cpp:94-134): binary searches inside every chunk, one range per chunk
      std::vector<RangeSelection::Range> ranges;
      size_t chunk_begin = 0;
      for (size_t chunk = 0; chunk < table.row_layout.numChunks(); ++chunk) {
         const size_t chunk_size = table.row_layout.chunk_sizes[chunk];
         const int32_t* begin = found->dates.data() + chunk_begin;
         const int32_t* end = begin + chunk_size;
         const auto lower_index = static_cast<size_t>(std::lower_bound(begin, end, from) - begin);
         const auto upper_index = date_to.has_value() ? static_cast<size_t>(std::upper_bound(begin, end, *date_to) - begin) : chunk_size;
         const auto global_chunk = static_cast<uint32_t>(table.row_layout.first_chunk + chunk);
         const uint32_t start_row = lower_index == chunk_size ? (global_chunk + 1) << 16 : (global_chunk << 16) | static_cast<uint32_t>(lower_index);
         const uint32_t end_row = upper_index == chunk_size ? (global_chunk + 1) << 16 : (global_chunk << 16) | static_cast<uint32_t>(upper_index);
         ranges.push_back({start_row, end_row});
         chunk_begin += chunk_size;
      }
      const uint32_t layout_begin = table.row_layout.first_chunk << 16;
      const uint32_t layout_end = (table.row_layout.first_chunk + static_cast<uint32_t>(table.row_layout.numChunks())) << 16;
      return std::make_unique<RangeSelection>(std::move(ranges), layout_begin, layout_end);
   }
   std::vector<CoveragePredicate> predicates(2);
   for (CoveragePredicate& predicate : predicates) {
      predicate.kind = CoveragePredicate::COMPARE;
      predicate.value_column = found->device_column;
      predicate.is_signed = true;
   }
   predicates[0].comparator = SILO_CMP_HIGHER_OR_EQUALS;
   predicates[0].value = static_cast<uint32_t>(from);
   predicates[0].display = "$date " + column + " >= " + std::to_string(from);
   const int32_t to = date_to.value_or(std::numeric_limits<int32_t>::max());
   predicates[1].comparator = SILO_CMP_LESS_OR_EQUALS;
   predicates[1].value = static_cast<uint32_t>(to);
   predicates[1].display = "$date " + column + " <= " + std::to_string(to);
   return std::make_unique<Selection>(std::nullopt, std::move(predicates));
}

// ---- boundary leaves -------------------------------------------------------------------------

std::unique_ptr<Operator> BitmapFilter::compile(const Table& table) const {
   auto iter = table.named_bitmaps.find(name);
   if (iter == table.named_bitmaps.end()) {
      return std::make_unique<Empty>();  // unknown value, lineage_filter.cpp:93-95
   }
   if (iter->second.resident) {
      return IndexScan::overIndexBitmap(iter->second.device_id);
   }
   return IndexScan::overBitmap(&iter->second.bytes);
}

std::unique_ptr<Operator> RowRanges::compile(const Table& table) const {
   auto copy = ranges;
   const uint32_t begin = table.row_layout.first_chunk << 16;
   const uint32_t end = (table.row_layout.first_chunk + static_cast<uint32_t>(table.row_layout.numChunks())) << 16;
   return std::make_unique<RangeSelection>(std::move(copy), begin, end);
}

DeviceBitmap computeFilter(const ScalarExpression& filter, const Table& table) {
   const double begin = nowMicroseconds();
   const ExpressionPtr rewritten = filter.rewrite(table, AmbiguityMode::NONE);
   const std::unique_ptr<Operator> compiled = rewritten->compile(table);
   lastQueryProfile().compile_us = nowMicroseconds() - begin;
   return compiled->evaluate(table);
}

// ---- harness notation ------------------------------------------------------------------------

namespace {

// physical forms, used by the operator-level known-answer tests only
struct IdsLeaf : ScalarExpression {
   std::vector<uint8_t> bytes;
   explicit IdsLeaf(std::vector<uint32_t> ids) {
      std::sort(ids.begin(), ids.end());
      ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
      bytes = writePortableRoaring(ids);
   }
   std::string toString() const override { return "ids"; }
   ExpressionPtr rewrite(const Table&, AmbiguityMode) const override { return shared_from_this(); }
   std::unique_ptr<Operator> compile(const Table&) const override { return IndexScan::overBitmap(&bytes); }
};

struct CoveredLeaf : ScalarExpression {
   std::string column;
   uint32_t position_idx;
   bool covered;
   CoveredLeaf(std::string column, uint32_t position_idx, bool covered)
       : column(std::move(column)),
         position_idx(position_idx),
         covered(covered) {}
   std::string toString() const override { return "covered"; }
   ExpressionPtr rewrite(const Table&, AmbiguityMode) const override { return shared_from_this(); }
   std::unique_ptr<Operator> compile(const Table& table) const override {
      const auto& sequence_column = requireColumn(table, column);
      CHECK_QUERY(position_idx < sequence_column.reference_sequence.size(), "position is out of bounds");
      return std::make_unique<Selection>(CoveragePredicate{sequence_column.device_column, position_idx, covered});
   }
};

struct RawSymbolInSet : ScalarExpression {
   std::string column;
   uint32_t position_idx;
   std::string chars;
   std::string toString() const override { return "sym-in"; }
   ExpressionPtr rewrite(const Table& table, AmbiguityMode) const override {
      const auto& sequence_column = requireColumn(table, column);
      std::vector<Symbol> symbols;
      for (char character : chars) {
         symbols.push_back(toSymbol(*sequence_column.alphabet, character));
      }
      return std::make_shared<SymbolInSet>(column, position_idx, std::move(symbols), sequence_column.alphabet);
   }
   std::unique_ptr<Operator> compile(const Table& table) const override {
      return rewrite(table, AmbiguityMode::NONE)->compile(table);
   }
};

struct PhysicalOperator : ScalarExpression {
   enum Kind { AND, OR, NOT, THRESHOLD } kind = AND;
   ExpressionVector first;
   ExpressionVector second;
   uint32_t number_of_matchers = 0;
   bool match_exactly = false;
   std::string toString() const override { return "physical"; }
   ExpressionPtr rewrite(const Table& table, AmbiguityMode mode) const override {
      auto copy = std::make_shared<PhysicalOperator>(*this);
      for (auto& child : copy->first) {
         child = child->rewrite(table, mode);
      }
      for (auto& child : copy->second) {
         child = child->rewrite(table, mode);
      }
      return copy;
   }
   std::unique_ptr<Operator> compile(const Table& table) const override {
      OperatorVector first_operators;
      OperatorVector second_operators;
      for (const auto& child : first) {
         first_operators.push_back(child->compile(table));
      }
      for (const auto& child : second) {
         second_operators.push_back(child->compile(table));
      }
      switch (kind) {
         case AND:
            return std::make_unique<Intersection>(std::move(first_operators), std::move(second_operators));
         case OR:
            return std::make_unique<Union>(std::move(first_operators));
         case NOT:
            return std::make_unique<Complement>(std::move(first_operators.at(0)));
         case THRESHOLD:
            return std::make_unique<Threshold>(
               std::move(first_operators), std::move(second_operators), number_of_matchers, match_exactly
            );
      }
      throw std::logic_error("unreachable");
   }
};

struct Node {
   bool is_atom = false;
   std::string_view atom;  // a view into the expression text (no allocation per token)
   std::vector<Node> items;
   // the arguments of a "(ranges ...)" / "(ids ...)" list: read as numbers straight from the text (one Node
   // per number made a config-2 date filter of 153 ranges the slowest part of reading the expression)
   std::vector<uint64_t> numbers;
};

class Reader {
   const std::string& text;
   size_t cursor = 0;

   // (the "C" locale's white space, inline: std::isspace is a call per character, and a date filter over 153
   // chunks is 2.4 KB of text)
   static bool isSpace(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

   void skipSpace() {
      while (cursor < text.size() && isSpace(text[cursor])) {
         ++cursor;
      }
   }

  public:
   explicit Reader(const std::string& text) : text(text) {}

   Node next() {
      skipSpace();
      CHECK_QUERY(cursor < text.size(), "filter expression ended unexpectedly");
      Node node;
      const char head = text[cursor];
      CHECK_QUERY(head != ')', "filter expression: unexpected ')'");
      if (head == '(') {
         ++cursor;
         node.items.reserve(16);
         for (;;) {
            skipSpace();
            CHECK_QUERY(cursor < text.size(), "filter expression: missing ')'");
            if (text[cursor] == ')') {
               ++cursor;
               return node;
            }
            node.items.push_back(next());
            if (node.items.size() == 1 && node.items[0].is_atom && (node.items[0].atom == "ranges" || node.items[0].atom == "ids")) {
               readNumbers(node.numbers);
               return node;
            }
         }
      }
      node.is_atom = true;
      if (head == '"') {
         const size_t close = text.find('"', cursor + 1);
         CHECK_QUERY(close != std::string::npos, "filter expression: unterminated string");
         node.atom = std::string_view(text).substr(cursor + 1, close - cursor - 1);
         cursor = close + 1;
         return node;
      }
      const size_t begin = cursor;
      while (cursor < text.size() && !isSpace(text[cursor]) && text[cursor] != '(' && text[cursor] != ')') {
         ++cursor;
      }
      node.atom = std::string_view(text).substr(begin, cursor - begin);
      return node;
   }

   // the rest of a list of non-negative integers, up to and including its ')'
   void readNumbers(std::vector<uint64_t>& numbers) {
      numbers.reserve(64);
      for (;;) {
         skipSpace();
         CHECK_QUERY(cursor < text.size(), "filter expression: missing ')'");
         if (text[cursor] == ')') {
            ++cursor;
            return;
         }
         CHECK_QUERY(text[cursor] != '(', "filter expression: expected an atom");
         const size_t begin = cursor;
         uint64_t value = 0;
         bool valid = true;
         while (cursor < text.size() && !isSpace(text[cursor]) && text[cursor] != '(' && text[cursor] != ')') {
            const char c = text[cursor++];
            valid = valid && c >= '0' && c <= '9';
            value = value * 10 + static_cast<uint64_t>(c - '0');
         }
         if (!valid || cursor - begin > 19) {
            throw IllegalQueryException(
               "filter expression: expected a non-negative integer, got '" + text.substr(begin, cursor - begin) + "'"
            );
         }
         numbers.push_back(value);
      }
   }

   bool exhausted() {
      skipSpace();
      return cursor >= text.size();
   }
};

std::string atom(const Node& node) {
   CHECK_QUERY(node.is_atom, "filter expression: expected an atom");
   return std::string(node.atom);
}

uint64_t number(const Node& node) {
   CHECK_QUERY(node.is_atom, "filter expression: expected an atom");
   const std::string_view text = node.atom;
   uint64_t value = 0;
   bool valid = !text.empty() && text.size() <= 19;
   for (const char c : text) {
      valid = valid && c >= '0' && c <= '9';
      value = value * 10 + static_cast<uint64_t>(c - '0');
   }
   if (!valid) {
      throw IllegalQueryException("filter expression: expected a non-negative integer, got '" + std::string(text) + "'");
   }
   return value;
}

// numbers that end up in 32-bit fields: a value that does not fit is a user error, never silently another number
// (position 4294967297 must not pass the bounds check as position 1)
uint32_t narrow32(uint64_t value, const char* what) {
   if (value > UINT32_MAX) {
      throw IllegalQueryException(std::string("filter expression: ") + what + " " + std::to_string(value) + " does not fit in 32 bits");
   }
   return static_cast<uint32_t>(value);
}
int narrowInt(uint64_t value, const char* what) {
   if (value > static_cast<uint64_t>(INT32_MAX)) {
      throw IllegalQueryException(std::string("filter expression: ") + what + " " + std::to_string(value) + " is out of range");
   }
   return static_cast<int>(value);
}

uint32_t position(const Node& node) {
   const uint64_t one_based = number(node);
   CHECK_QUERY(one_based != 0, "The field 'position' is 1-indexed. Value of 0 not allowed.");
   return narrow32(one_based - 1, "position");
}

ExpressionPtr build(const Node& node);

ExpressionVector buildList(const Node& list, size_t from = 0) {
   CHECK_QUERY(!list.is_atom, "filter expression: expected a list");
   ExpressionVector result;
   for (size_t i = from; i < list.items.size(); ++i) {
      result.push_back(build(list.items[i]));
   }
   return result;
}

ExpressionPtr build(const Node& node) {
   CHECK_QUERY(!node.is_atom && !node.items.empty(), "filter expression: expected a non-empty list");
   const auto& items = node.items;
   const std::string& head = atom(items[0]);
   auto arity = [&](size_t count) {
      CHECK_QUERY(items.size() == count + 1, "filter expression: wrong number of arguments for " + head);
   };
   if (head == "true" || head == "false") {
      arity(0);
      return std::make_shared<BoolLiteral>(head == "true");
   }
   if (head == "sym-eq") {
      arity(3);
      const std::string& symbol = atom(items[3]);
      CHECK_QUERY(symbol.size() == 1, "symbol must be a single character");
      return std::make_shared<SymbolEquals>(
         atom(items[1]), position(items[2]), symbol == "." ? std::nullopt : std::optional<char>(symbol[0])
      );
   }
   if (head == "sym-in") {
      arity(3);
      auto expression = std::make_shared<RawSymbolInSet>();
      expression->column = atom(items[1]);
      expression->position_idx = position(items[2]);
      expression->chars = atom(items[3]);
      return expression;
   }
   if (head == "has-mut") {
      arity(2);
      return std::make_shared<HasMutation>(atom(items[1]), position(items[2]));
   }
   if (head == "and") {
      return std::make_shared<And>(buildList(node, 1));
   }
   if (head == "or") {
      return std::make_shared<Or>(buildList(node, 1));
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
      CHECK_QUERY(items.size() >= 3, "filter expression: n-of needs K and EXACT");
      return std::make_shared<NOf>(buildList(node, 3), narrowInt(number(items[1]), "numberOfMatchers"), number(items[2]) != 0);
   }
   if (head == "profile") {
      CHECK_QUERY(items.size() >= 4, "filter expression: profile needs COL DIST KIND ..");
      const std::string& column = atom(items[1]);
      const auto distance = static_cast<uint32_t>(narrowInt(number(items[2]), "distance"));
      const std::string& kind = atom(items[3]);
      if (kind == "seq") {
         arity(4);
         return std::make_shared<MutationProfile>(column, distance, MutationProfile::QuerySequence{atom(items[4])});
      }
      if (kind == "muts") {
         CHECK_QUERY((items.size() - 4) % 2 == 0, "filter expression: profile muts needs POS SYM pairs");
         MutationProfile::Mutations mutations;
         for (size_t i = 4; i + 1 < items.size(); i += 2) {
            const std::string& symbol = atom(items[i + 1]);
            CHECK_QUERY(symbol.size() == 1, "symbol must be a single character");
            mutations.mutations.emplace_back(position(items[i]), symbol[0]);
         }
         return std::make_shared<MutationProfile>(column, distance, std::move(mutations));
      }
      // `row` (sequenceId, mutation_profile.cpp:109-158) needs the primary-key column and a row
      // reconstruction, both owned by the unchanged host engine: it hands over `seq` instead
      throw IllegalQueryException("filter expression: unsupported profile kind " + kind);
   }
   if (head == "str-eq") {  // (str-eq <column> <value>)
      arity(2);
      return std::make_shared<StringEquals>(atom(items[1]), atom(items[2]));
   }
   if (head == "date-between") {  // (date-between <column> <from day | *> <to day | *>)
      arity(3);
      auto bound = [&](const Node& node) -> std::optional<int32_t> {
         const std::string bound_text = atom(node);
         if (bound_text == "*") {
            return std::nullopt;
         }
         try {
            return static_cast<int32_t>(std::stol(bound_text));
         } catch (const std::exception&) {
            throw IllegalQueryException("filter expression: date-between bounds are day numbers or *, got '" + bound_text + "'");
         }
      };
      return std::make_shared<DateBetween>(atom(items[1]), bound(items[2]), bound(items[3]));
   }
   if (head == "bitmap") {
      arity(1);
      return std::make_shared<BitmapFilter>(atom(items[1]));
   }
   if (head == "ranges") {
      const std::vector<uint64_t>& numbers = node.numbers;
      CHECK_QUERY(numbers.size() % 2 == 0, "filter expression: ranges needs START END pairs");
      std::vector<RangeSelection::Range> ranges;
      ranges.reserve(numbers.size() / 2);
      for (size_t i = 0; i + 1 < numbers.size(); i += 2) {
         ranges.push_back({narrow32(numbers[i], "row id"), narrow32(numbers[i + 1], "row id")});
      }
      return std::make_shared<RowRanges>(std::move(ranges));
   }
   if (head == "ids") {
      std::vector<uint32_t> ids;
      ids.reserve(node.numbers.size());
      for (const uint64_t id : node.numbers) {
         ids.push_back(narrow32(id, "row id"));
      }
      return std::make_shared<IdsLeaf>(std::move(ids));
   }
   if (head == "covered" || head == "not-covered") {
      arity(2);
      return std::make_shared<CoveredLeaf>(atom(items[1]), position(items[2]), head == "covered");
   }
   if (head == "op-and" || head == "op-threshold") {
      auto expression = std::make_shared<PhysicalOperator>();
      size_t lists_from = 1;
      if (head == "op-threshold") {
         CHECK_QUERY(items.size() == 5, "filter expression: op-threshold K EXACT (pos..) (neg..)");
         expression->kind = PhysicalOperator::THRESHOLD;
         expression->number_of_matchers = narrow32(number(items[1]), "numberOfMatchers");
         expression->match_exactly = number(items[2]) != 0;
         lists_from = 3;
      } else {
         arity(2);
         expression->kind = PhysicalOperator::AND;
      }
      expression->first = buildList(items[lists_from]);
      expression->second = buildList(items[lists_from + 1]);
      return expression;
   }
   if (head == "op-or") {
      auto expression = std::make_shared<PhysicalOperator>();
      expression->kind = PhysicalOperator::OR;
      expression->first = buildList(node, 1);
      return expression;
   }
   if (head == "op-not") {
      arity(1);
      auto expression = std::make_shared<PhysicalOperator>();
      expression->kind = PhysicalOperator::NOT;
      expression->first.push_back(build(items[1]));
      return expression;
   }
   throw IllegalQueryException("filter expression: unknown form '" + head + "'");
}

}  // namespace

ExpressionPtr parseFilterExpression(const std::string& text) {
   Reader reader(text);
   const Node node = reader.next();
   CHECK_QUERY(reader.exhausted(), "filter expression: trailing input");
   return build(node);
}

}  // namespace silo_host
