// Logical filter expressions of the host layer: the rewrite(table, AmbiguityMode) -> compile(table)
// contract of /root/reference/src/rhydb/query_engine/scalar_expressions/scalar_expression.h:24,81-90
// for the sequence filters on this path. compile() yields the host layer's Operator tree
// (operators.h), whose leaves name device-resident data instead of holding container views.
//   SymbolEquals symbol_equals.cpp:65-100   | SymbolInSet symbol_in_set.cpp:129-263
//   HasMutation has_mutation.cpp:34-67      | NOf nof.cpp:33-277
//   And and.cpp:91-219 | Or or.cpp:34-242   | Negation negation.cpp:27-34
//   Maybe maybe.cpp:28-33 | Exact exact.cpp:28-33 | MutationProfile mutation_profile.cpp:198-257
//   BoolLiteral literal.cpp:101-121
// Boundary leaves (their predicates are evaluated by the unchanged host engine):
//   BitmapFilter  = LineageFilter::compile lineage_filter.cpp:77-100 (IndexScan over a ready bitmap)
//   RowRanges     = DateBetween::compile on a sorted column, date_between.cpp:75-79,94-134
#pragma once
#include <memory>
#include <string>
#include <variant>
#include <vector>

#include "operators.h"
#include "table.h"

namespace silo_host {

enum class AmbiguityMode : uint8_t { UPPER_BOUND, LOWER_BOUND, NONE };
AmbiguityMode invertMode(AmbiguityMode mode);  // scalar_expression.cpp:7-15

class ScalarExpression;
using ExpressionPtr = std::shared_ptr<const ScalarExpression>;
using ExpressionVector = std::vector<ExpressionPtr>;

class ScalarExpression : public std::enable_shared_from_this<ScalarExpression> {
  public:
   virtual ~ScalarExpression() = default;
   [[nodiscard]] virtual std::string toString() const = 0;
   [[nodiscard]] virtual ExpressionPtr rewrite(const Table& table, AmbiguityMode mode) const = 0;
   [[nodiscard]] virtual std::unique_ptr<Operator> compile(const Table& table) const = 0;
};

struct BoolLiteral : ScalarExpression {
   bool value;
   explicit BoolLiteral(bool value) : value(value) {}
   std::string toString() const override { return value ? "true" : "false"; }
   ExpressionPtr rewrite(const Table&, AmbiguityMode) const override;
   std::unique_ptr<Operator> compile(const Table& table) const override;
};

// A set of symbol ids as a bit mask (both alphabets have <= 32 symbols); the reference keeps a
// std::vector<Symbol> (symbol_in_set.h). No allocation per SymbolInSet: a MutationProfile creates one
// per genome position.
struct SymbolSet {
   uint32_t mask = 0;
   SymbolSet() = default;
   explicit SymbolSet(uint32_t mask) : mask(mask) {}
   SymbolSet(const std::vector<Symbol>& symbols) {  // NOLINT(google-explicit-constructor)
      for (Symbol symbol : symbols) {
         mask |= 1u << symbol;
      }
   }
   [[nodiscard]] size_t size() const { return static_cast<size_t>(__builtin_popcount(mask)); }
};

struct SymbolInSet : ScalarExpression {
   std::string column;
   uint32_t position_idx;
   SymbolSet symbols;
   const Alphabet* alphabet;  // SymbolType of the reference's template; names the symbols in toString()
   SymbolInSet(std::string column, uint32_t position_idx, SymbolSet symbols, const Alphabet* alphabet = nullptr)
       : column(std::move(column)),
         position_idx(position_idx),
         symbols(symbols),
         alphabet(alphabet) {}
   std::string toString() const override;
   ExpressionPtr rewrite(const Table&, AmbiguityMode) const override;
   std::unique_ptr<Operator> compile(const Table& table) const override;
};

struct SymbolEquals : ScalarExpression {
   std::string column;
   uint32_t position_idx;
   std::optional<char> symbol;  // nullopt = '.', the global reference symbol
   SymbolEquals(std::string column, uint32_t position_idx, std::optional<char> symbol)
       : column(std::move(column)),
         position_idx(position_idx),
         symbol(symbol) {}
   std::string toString() const override;
   ExpressionPtr rewrite(const Table& table, AmbiguityMode mode) const override;
   std::unique_ptr<Operator> compile(const Table& table) const override;
};

struct HasMutation : ScalarExpression {
   std::string column;
   uint32_t position_idx;
   HasMutation(std::string column, uint32_t position_idx)
       : column(std::move(column)),
         position_idx(position_idx) {}
   std::string toString() const override;
   ExpressionPtr rewrite(const Table& table, AmbiguityMode mode) const override;
   std::unique_ptr<Operator> compile(const Table& table) const override;
};

struct Negation : ScalarExpression {
   ExpressionPtr child;
   explicit Negation(ExpressionPtr child) : child(std::move(child)) {}
   std::string toString() const override { return "!(" + child->toString() + ")"; }
   ExpressionPtr rewrite(const Table& table, AmbiguityMode mode) const override;
   std::unique_ptr<Operator> compile(const Table& table) const override;
};

struct Maybe : ScalarExpression {
   ExpressionPtr child;
   explicit Maybe(ExpressionPtr child) : child(std::move(child)) {}
   std::string toString() const override { return "Maybe (" + child->toString() + ")"; }
   ExpressionPtr rewrite(const Table& table, AmbiguityMode mode) const override;
   std::unique_ptr<Operator> compile(const Table& table) const override;
};

struct Exact : ScalarExpression {
   ExpressionPtr child;
   explicit Exact(ExpressionPtr child) : child(std::move(child)) {}
   std::string toString() const override { return "Exact (" + child->toString() + ")"; }
   ExpressionPtr rewrite(const Table& table, AmbiguityMode mode) const override;
   std::unique_ptr<Operator> compile(const Table& table) const override;
};

struct And : ScalarExpression {
   ExpressionVector children;
   explicit And(ExpressionVector children) : children(std::move(children)) {}
   std::string toString() const override;
   ExpressionPtr rewrite(const Table& table, AmbiguityMode mode) const override;
   std::unique_ptr<Operator> compile(const Table& table) const override;
};

struct Or : ScalarExpression {
   ExpressionVector children;
   explicit Or(ExpressionVector children) : children(std::move(children)) {}
   std::string toString() const override;
   ExpressionPtr rewrite(const Table& table, AmbiguityMode mode) const override;
   std::unique_ptr<Operator> compile(const Table& table) const override;
};

// Many SymbolInSet children of one NOf in compact form: (position, symbol set) pairs over one sequence column, in
// child order. MutationProfile::rewrite produces one child per genome position (mutation_profile.cpp:222-247);
// as ~30,000 expression objects they were half of the host time of such a query. They behave exactly like the
// SymbolInSet children they stand for (toString, compile).
struct SymbolInSetSpan {
   std::string column;
   const Alphabet* alphabet = nullptr;
   std::vector<uint32_t> positions;
   std::vector<uint32_t> masks;
   [[nodiscard]] size_t size() const { return positions.size(); }
};

struct NOf : ScalarExpression {
   ExpressionVector children;
   int number_of_matchers;
   bool match_exactly;
   std::shared_ptr<const SymbolInSetSpan> span;  // further children, behind `children`
   NOf(ExpressionVector children, int number_of_matchers, bool match_exactly, std::shared_ptr<const SymbolInSetSpan> span = nullptr)
       : children(std::move(children)),
         number_of_matchers(number_of_matchers),
         match_exactly(match_exactly),
         span(std::move(span)) {}
   std::string toString() const override;
   ExpressionPtr rewrite(const Table& table, AmbiguityMode mode) const override;
   std::unique_ptr<Operator> compile(const Table& table) const override;
};

// `column = 'value'` on a string column without an index (equals.cpp:124-156): rewritten to a StringInSet, compiled to a
// Selection with that predicate. The device compares dictionary ids.
struct StringEquals : ScalarExpression {
   std::string column;
   std::string value;
   StringEquals(std::string column, std::string value) : column(std::move(column)), value(std::move(value)) {}
   std::string toString() const override;
   ExpressionPtr rewrite(const Table& table, AmbiguityMode mode) const override;
   std::unique_ptr<Operator> compile(const Table& table) const override;
};

// date_between.cpp:61-134: a RangeSelection with one range per chunk on a sorted column, else a Selection with the two
// comparisons. from / to: days since the epoch (Date32), nullopt = open end.
struct DateBetween : ScalarExpression {
   std::string column;
   std::optional<int32_t> date_from;
   std::optional<int32_t> date_to;
   DateBetween(std::string column, std::optional<int32_t> date_from, std::optional<int32_t> date_to)
       : column(std::move(column)), date_from(date_from), date_to(date_to) {}
   std::string toString() const override;
   ExpressionPtr rewrite(const Table& table, AmbiguityMode mode) const override;
   std::unique_ptr<Operator> compile(const Table& table) const override;
};

struct MutationProfile : ScalarExpression {
   struct QuerySequence {
      std::string sequence;
   };
   struct Mutations {
      std::vector<std::pair<uint32_t, char>> mutations;  // (position_idx, symbol)
   };
   std::string column;
   uint32_t distance;
   std::variant<QuerySequence, Mutations> input;
   MutationProfile(std::string column, uint32_t distance, std::variant<QuerySequence, Mutations> input)
       : column(std::move(column)),
         distance(distance),
         input(std::move(input)) {}
   std::string toString() const override;
   ExpressionPtr rewrite(const Table& table, AmbiguityMode mode) const override;
   std::unique_ptr<Operator> compile(const Table& table) const override;
};

struct BitmapFilter : ScalarExpression {
   std::string name;
   explicit BitmapFilter(std::string name) : name(std::move(name)) {}
   std::string toString() const override { return "bitmap:" + name; }
   ExpressionPtr rewrite(const Table&, AmbiguityMode) const override { return shared_from_this(); }
   std::unique_ptr<Operator> compile(const Table& table) const override;
};

struct RowRanges : ScalarExpression {
   std::vector<RangeSelection::Range> ranges;
   explicit RowRanges(std::vector<RangeSelection::Range> ranges) : ranges(std::move(ranges)) {}
   std::string toString() const override { return "ranges"; }
   ExpressionPtr rewrite(const Table&, AmbiguityMode) const override { return shared_from_this(); }
   std::unique_ptr<Operator> compile(const Table& table) const override;
};

// symbol_in_set.cpp:231-264
std::unique_ptr<Operator> compileSymbolInSet(
   const SequenceColumnInfo& sequence_column,
   uint32_t position_idx,
   SymbolSet symbols
);

// operators/compute_filter.cpp:14-21
DeviceBitmap computeFilter(const ScalarExpression& filter, const Table& table);

// harness notation shared with the oracle (see oracle/src/expressions.h for the grammar)
ExpressionPtr parseFilterExpression(const std::string& text);

}  // namespace silo_host
