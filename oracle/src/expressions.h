// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// Restates the logical filter layer under
// /root/reference/src/rhydb/query_engine/scalar_expressions/:
//   scalar_expression.h:24,81-90 (AmbiguityMode, rewrite/compile), scalar_expression.cpp:7-15
//   symbol_equals.cpp:65-100 | symbol_in_set.cpp:129-263 | has_mutation.cpp:34-67
//   nof.cpp:33-150,184-277   | and.cpp:91-219           | or.cpp:34-242
//   negation.cpp:27-34       | maybe.cpp:28-33          | exact.cpp:28-33
//   mutation_profile.cpp:60-257 | literal.cpp:101-121 (BoolLiteral)
//   lineage_filter.cpp:77-100 (-> IndexScan over a ready-made bitmap)
//   date_between.cpp:61-134   (-> RangeSelection on a sorted column)
// and operators/compute_filter.cpp:14-21 (computeFilter = rewrite(NONE) -> compile -> evaluate).
//
// Filters are written in a small s-expression notation (the SaneQL front-end is out of scope):
//   (true) (false)
//   (sym-eq COL POS1 SYM|.)   (has-mut COL POS1)   (sym-in COL POS1 SYMS)
//   (and e..) (or e..) (not e) (maybe e) (exact e) (n-of K EXACT01 e..)
//   (profile COL DIST seq SEQ) | (profile COL DIST muts POS1 SYM ..) | (profile COL DIST row ROWID)
//   (bitmap NAME)  (ranges START END ..)            -- boundary leaves, global sparse row ids
//   physical forms used by the operator-level known-answer tests:
//   (ids v..) (op-and (e..) (e..)) (op-or e..) (op-not e) (op-threshold K EXACT01 (e..) (e..))
//   (covered COL POS1) (not-covered COL POS1)
// POS1 is 1-based like the query language; everything else is 0-based.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "operators.h"
#include "storage.h"

namespace oracle {

enum class AmbiguityMode : uint8_t { UPPER_BOUND, LOWER_BOUND, NONE };

class Expression;
using ExprPtr = std::shared_ptr<const Expression>;
using ExpressionVector = std::vector<ExprPtr>;

class Expression : public std::enable_shared_from_this<Expression> {
  public:
   virtual ~Expression() = default;
   [[nodiscard]] virtual std::string toString() const = 0;
   [[nodiscard]] virtual ExprPtr rewrite(const Table& table, AmbiguityMode mode) const = 0;
   [[nodiscard]] virtual std::unique_ptr<Operator> compile(const Table& table) const = 0;
};

ExprPtr parseExpression(const std::string& text);

// compute_filter.cpp:14-21
CowBitmap computeFilter(const Expression& filter, const Table& table);

// symbol_in_set.cpp:231-264
std::unique_ptr<Operator> compileSymbolInSet(
   const SequenceColumn& sequence_column,
   uint32_t position_idx,
   const std::vector<Symbol>& symbols,
   const RowLayout& row_layout
);

}  // namespace oracle
