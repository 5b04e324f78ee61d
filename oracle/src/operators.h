// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// Restates the physical filter operators under
// /root/reference/src/rhydb/query_engine/filter/operators/:
//   operator.h:11-39, operator.cpp:19-67 (negate dispatch)
//   index_scan.cpp:43-50 | intersection.cpp:19-110 | union.cpp:34-47 | complement.cpp:23-60
//   threshold.cpp:19-143 | selection.cpp:94-150 (+ is_in_covered_region.cpp:31-82)
//   range_selection.cpp:54-111 | full.cpp:26-33 | empty.cpp:25-31
#pragma once
#include <memory>
#include <optional>
#include <string>
#include <vector>

#include "cow_bitmap.h"
#include "storage.h"

namespace oracle {

enum OperatorType : uint8_t {
   EMPTY,
   FULL,
   INDEX_SCAN,
   INTERSECTION,
   COMPLEMENT,
   RANGE_SELECTION,
   SELECTION,
   THRESHOLD,
   UNION,
   BITMAP_PRODUCER
};

class Operator {
  public:
   virtual ~Operator() = default;
   [[nodiscard]] virtual OperatorType type() const = 0;
   [[nodiscard]] virtual CowBitmap evaluate() const = 0;
   [[nodiscard]] virtual std::string toString() const = 0;
   static std::unique_ptr<Operator> negate(std::unique_ptr<Operator>&& some_operator);
};
using OperatorVector = std::vector<std::unique_ptr<Operator>>;

class Empty : public Operator {
  public:
   RowLayout row_layout;
   explicit Empty(RowLayout row_layout) : row_layout(std::move(row_layout)) {}
   OperatorType type() const override { return EMPTY; }
   CowBitmap evaluate() const override { return {}; }
   std::string toString() const override { return "Empty"; }
};

class Full : public Operator {
  public:
   RowLayout row_layout;
   explicit Full(RowLayout row_layout) : row_layout(std::move(row_layout)) {}
   OperatorType type() const override { return FULL; }
   CowBitmap evaluate() const override { return CowBitmap{row_layout.fullBitmap()}; }
   std::string toString() const override { return "Full"; }
};

class IndexScan : public Operator {
  public:
   CowBitmap bitmap;
   RowLayout row_layout;
   IndexScan(CowBitmap bitmap, RowLayout row_layout)
       : bitmap(std::move(bitmap)),
         row_layout(std::move(row_layout)) {}
   OperatorType type() const override { return INDEX_SCAN; }
   CowBitmap evaluate() const override { return bitmap; }
   std::string toString() const override {
      return "IndexScan(Cardinality: " + std::to_string(bitmap.cardinality()) + ")";
   }
};

class Intersection : public Operator {
  public:
   OperatorVector children;
   OperatorVector negated_children;
   RowLayout row_layout;
   Intersection(OperatorVector&& children, OperatorVector&& negated_children, RowLayout row_layout);
   OperatorType type() const override { return INTERSECTION; }
   CowBitmap evaluate() const override;
   std::string toString() const override;
};

class Union : public Operator {
  public:
   OperatorVector children;
   RowLayout row_layout;
   Union(OperatorVector&& children, RowLayout row_layout)
       : children(std::move(children)),
         row_layout(std::move(row_layout)) {}
   OperatorType type() const override { return UNION; }
   CowBitmap evaluate() const override;
   std::string toString() const override;
};

class Complement : public Operator {
  public:
   std::unique_ptr<Operator> child;
   RowLayout row_layout;
   Complement(std::unique_ptr<Operator> child, RowLayout row_layout)
       : child(std::move(child)),
         row_layout(std::move(row_layout)) {}
   static std::unique_ptr<Complement> fromDeMorgan(OperatorVector disjunction, RowLayout row_layout);
   OperatorType type() const override { return COMPLEMENT; }
   CowBitmap evaluate() const override;
   std::string toString() const override { return "!" + child->toString(); }
};

class Threshold : public Operator {
  public:
   OperatorVector non_negated_children;
   OperatorVector negated_children;
   uint32_t number_of_matchers;
   bool match_exactly;
   RowLayout row_layout;
   Threshold(
      OperatorVector&& non_negated_children,
      OperatorVector&& negated_children,
      uint32_t number_of_matchers,
      bool match_exactly,
      RowLayout row_layout
   );
   OperatorType type() const override { return THRESHOLD; }
   CowBitmap evaluate() const override;
   std::string toString() const override;
};

class RangeSelection : public Operator {
  public:
   struct Range {
      uint32_t start;  // global sparse row id (chunk << 16 | row); end may be (numChunks << 16)
      uint32_t end;
   };
   std::vector<Range> ranges;
   RowLayout row_layout;
   RangeSelection(std::vector<Range>&& ranges, RowLayout row_layout)
       : ranges(std::move(ranges)),
         row_layout(std::move(row_layout)) {}
   OperatorType type() const override { return RANGE_SELECTION; }
   CowBitmap evaluate() const override;
   std::string toString() const override { return "RangeSelection"; }
};

// selection.h:30-47
class Predicate {
  public:
   virtual ~Predicate() = default;
   [[nodiscard]] virtual std::string toString() const = 0;
   [[nodiscard]] virtual bool match(uint32_t global_row_id) const = 0;
   [[nodiscard]] virtual Roaring makeBitmap(const RowLayout& row_layout) const;
   [[nodiscard]] virtual double estimateSelectivity(uint32_t /*row_count*/) const { return 0.5; }
   [[nodiscard]] virtual std::unique_ptr<Predicate> negate() const = 0;
};
using PredicateVector = std::vector<std::unique_ptr<Predicate>>;

class IsInCoveredRegion : public Predicate {
  public:
   enum class Comparator : uint8_t { IS_COVERED, IS_NOT_COVERED };
   const HorizontalCoverageIndex* horizontal_coverage_index;
   uint32_t position_idx;
   Comparator comparator;
   IsInCoveredRegion(const HorizontalCoverageIndex* index, uint32_t position_idx, Comparator comparator)
       : horizontal_coverage_index(index),
         position_idx(position_idx),
         comparator(comparator) {}
   std::string toString() const override;
   [[nodiscard]] bool isCovered(uint32_t row_id) const;
   bool match(uint32_t global_row_id) const override;
   Roaring makeBitmap(const RowLayout& row_layout) const override;
   double estimateSelectivity(uint32_t /*row_count*/) const override { return 0.1; }
   std::unique_ptr<Predicate> negate() const override;
};

// CompareToValueSelection<Date32Column> (selection.h:76-166): a null row gives with_nulls
class DateCompare : public Predicate {
  public:
   enum class Comparator : uint8_t { EQUALS, LESS, HIGHER, LESS_OR_EQUALS, HIGHER_OR_EQUALS, NOT_EQUALS };
   const DateValueColumn* column;
   const std::vector<size_t>* chunk_begin;  // dense index of every chunk's first row
   Comparator comparator;
   int32_t value;
   bool with_nulls;
   DateCompare(const DateValueColumn* column, const std::vector<size_t>* chunk_begin, Comparator comparator, int32_t value, bool with_nulls = false)
       : column(column), chunk_begin(chunk_begin), comparator(comparator), value(value), with_nulls(with_nulls) {}
   std::string toString() const override;
   bool match(uint32_t global_row_id) const override;
   std::unique_ptr<Predicate> negate() const override;
};

// filter/operators/string_in_set.cpp:41-58
class StringInSetPredicate : public Predicate {
  public:
   const StringValueColumn* column;
   const std::vector<size_t>* chunk_begin;
   bool in;  // IN / NOT_IN
   std::vector<std::string> values;
   StringInSetPredicate(const StringValueColumn* column, const std::vector<size_t>* chunk_begin, bool in, std::vector<std::string> values)
       : column(column), chunk_begin(chunk_begin), in(in), values(std::move(values)) {}
   std::string toString() const override;
   bool match(uint32_t global_row_id) const override;
   std::unique_ptr<Predicate> negate() const override;
};

class Selection : public Operator {
  public:
   std::optional<std::unique_ptr<Operator>> child_operator;
   PredicateVector predicates;
   RowLayout row_layout;
   Selection(
      std::optional<std::unique_ptr<Operator>> child_operator,
      PredicateVector&& predicates,
      RowLayout row_layout
   );
   Selection(std::unique_ptr<Predicate> predicate, RowLayout row_layout);
   OperatorType type() const override { return SELECTION; }
   CowBitmap evaluate() const override;
   std::string toString() const override;
};

}  // namespace oracle
