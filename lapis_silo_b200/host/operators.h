// Physical filter operators of the host layer: same names, constructor checks and negate() rules as
// /root/reference/src/rhydb/query_engine/filter/operators/ (operator.h:11-39, operator.cpp:19-67),
// but evaluate() lowers the tree to a flat filter program (include/silo_b200.h) and runs it on the
// device instead of folding roaring containers on the CPU. The result type is a device-resident
// dense bitset (DeviceBitmap) where the reference returns a CopyOnWriteBitmap
// (query_engine/copy_on_write_bitmap.h:28-147).
#pragma once
#include <memory>
#include <optional>
#include <string>
#include <vector>

#include "table.h"

namespace silo_host {

// common/string_utils.h:35-57: at most `limit` items, then "<delimiter>... (<n> more)"
template <typename T>
std::string joinWithLimit(const std::vector<T>& items, const std::string& delimiter = ", ", size_t limit = 10) {
   std::string res;
   const size_t items_to_print = items.size() < limit ? items.size() : limit;
   for (size_t i = 0; i < items_to_print; ++i) {
      if (i > 0) {
         res += delimiter;
      }
      res += items[i]->toString();
   }
   if (items.size() > items_to_print) {
      res += delimiter + "... (" + std::to_string(items.size() - items_to_print) + " more)";
   }
   return res;
}

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

// Owns a silo_gpu_filter. Consumers that stay on the device (Mutations, count) use the handle;
// the others download it (silo_gpu_filter_download) and wrap it as a roaring bitmap.
class DeviceBitmap {
   std::shared_ptr<silo_gpu_filter> handle;
   uint64_t cached_cardinality = 0;

  public:
   DeviceBitmap() = default;
   DeviceBitmap(silo_gpu_filter* filter, uint64_t cardinality);
   [[nodiscard]] uint64_t cardinality() const { return cached_cardinality; }
   [[nodiscard]] bool isEmpty() const { return cached_cardinality == 0; }
   [[nodiscard]] const silo_gpu_filter* get() const { return handle.get(); }
   // dense words, 1024 per chunk (bit r of chunk c = row ((first_chunk + c) << 16) | r)
   [[nodiscard]] std::vector<uint64_t> toWords(size_t n_chunks) const;
};

class ProgramBuilder {
  public:
   const Table* table = nullptr;  // for per-column facts the lowering needs (genome length)
   std::vector<silo_filter_instr> instrs;
   std::vector<uint8_t> blob;
   std::vector<silo_roaring_bytes> bitmaps;
   bool inside_counter_program = false;  // between THR_BEGIN and THR_END: counter programs do not nest

   void emit(uint8_t opcode, uint8_t flags = 0, uint16_t column = 0, uint32_t a = 0, uint64_t b = 0);
   uint64_t addBlob(const void* data, size_t bytes, size_t alignment);
   uint32_t addBitmap(const std::vector<uint8_t>& portable_roaring_bytes);
};

class Operator {
  public:
   virtual ~Operator() = default;
   [[nodiscard]] virtual OperatorType type() const = 0;
   [[nodiscard]] virtual std::string toString() const = 0;
   // appends instructions that leave exactly one more tile on the program's stack
   virtual void lower(ProgramBuilder& program) const = 0;
   [[nodiscard]] DeviceBitmap evaluate(const Table& table) const;
   // the flat program of this tree; its pointers refer to `builder`, which must outlive it
   [[nodiscard]] silo_filter_program lowerProgram(const Table& table, ProgramBuilder& builder) const;
   static std::unique_ptr<Operator> negate(std::unique_ptr<Operator>&& some_operator);
};
using OperatorVector = std::vector<std::unique_ptr<Operator>>;

class Empty : public Operator {
  public:
   OperatorType type() const override { return EMPTY; }
   std::string toString() const override { return "Empty"; }
   void lower(ProgramBuilder& program) const override;
};

class Full : public Operator {
  public:
   OperatorType type() const override { return FULL; }
   std::string toString() const override { return "Full"; }
   void lower(ProgramBuilder& program) const override;
};

// IndexScan keeps its provenance instead of container views: the union of stored containers
// (symbol_in_set.cpp:216-228), a foreign roaring bitmap (lineage_filter.cpp:96-99) or the column's
// null_bitmap (symbol_in_set.cpp:86-92).
class IndexScan : public Operator {
  public:
   enum class Source : uint8_t { SYMBOLS, BITMAP, INDEX_BITMAP, NULLS };
   Source source;
   int device_column = 0;
   uint32_t position_idx = 0;
   uint32_t symbol_mask = 0;
   const std::vector<uint8_t>* bitmap_bytes = nullptr;
   uint32_t index_bitmap_id = 0;  // silo_gpu_bitmap_register id (Source::INDEX_BITMAP)

   static std::unique_ptr<IndexScan> overSymbols(int device_column, uint32_t position_idx, uint32_t symbol_mask);
   static std::unique_ptr<IndexScan> overBitmap(const std::vector<uint8_t>* portable_roaring_bytes);
   static std::unique_ptr<IndexScan> overIndexBitmap(uint32_t device_id);
   static std::unique_ptr<IndexScan> overNulls(int device_column);
   OperatorType type() const override { return INDEX_SCAN; }
   std::string toString() const override;
   void lower(ProgramBuilder& program) const override;
};

// Many leaf children of one wide Union / Threshold in compact form: per position of ONE sequence column, what
// compileSymbolInSet (symbol_in_set.cpp:231-264) would have built for a symbol set that does not hold the missing
// symbol -- an IndexScan (`adds`) or Selection[IsCovered] minus IndexScan (`covered_positions`, `subs`). A
// MutationProfile has one such child per genome position; building ~30,000 operator objects per query was most of
// the host time of such a query. The operators themselves are only materialised for toString().
struct SymbolScanSpan {
   int device_column = 0;
   std::string column_name;
   struct Leaf {
      uint32_t position_idx;
      uint32_t mask;             // the requested symbols
      bool includes_reference;   // of the column's local reference at this position
   };
   std::vector<Leaf> leaves;     // in child order
   uint32_t all_symbols_mask = 0;
   uint32_t missing_bit = 0;
   [[nodiscard]] size_t size() const { return leaves.size(); }
   [[nodiscard]] std::unique_ptr<Operator> materialise(size_t index) const;
   [[nodiscard]] std::string joinedStrings(const std::string& delimiter, size_t already_printed, size_t limit = 10) const;
};

class Intersection : public Operator {
  public:
   OperatorVector children;
   OperatorVector negated_children;
   Intersection(OperatorVector&& children, OperatorVector&& negated_children);  // intersection.cpp:19-41
   OperatorType type() const override { return INTERSECTION; }
   std::string toString() const override;
   void lower(ProgramBuilder& program) const override;
};

class Union : public Operator {
  public:
   OperatorVector children;
   std::shared_ptr<const SymbolScanSpan> span;  // further children, behind `children`
   explicit Union(OperatorVector&& children, std::shared_ptr<const SymbolScanSpan> span = nullptr)
       : children(std::move(children)),
         span(std::move(span)) {}
   OperatorType type() const override { return UNION; }
   std::string toString() const override;
   void lower(ProgramBuilder& program) const override;
};

class Complement : public Operator {
  public:
   std::unique_ptr<Operator> child;
   explicit Complement(std::unique_ptr<Operator> child) : child(std::move(child)) {}
   static std::unique_ptr<Complement> fromDeMorgan(OperatorVector disjunction);  // complement.cpp:23-41
   OperatorType type() const override { return COMPLEMENT; }
   std::string toString() const override { return "!" + child->toString(); }
   void lower(ProgramBuilder& program) const override;
};

class Threshold : public Operator {
  public:
   OperatorVector non_negated_children;
   OperatorVector negated_children;
   uint32_t number_of_matchers;
   bool match_exactly;
   std::shared_ptr<const SymbolScanSpan> span;  // further non-negated children, behind non_negated_children
   Threshold(
      OperatorVector&& non_negated_children,
      OperatorVector&& negated_children,
      uint32_t number_of_matchers,
      bool match_exactly,
      std::shared_ptr<const SymbolScanSpan> span = nullptr
   );  // threshold.cpp:19-41
   OperatorType type() const override { return THRESHOLD; }
   std::string toString() const override;
   void lower(ProgramBuilder& program) const override;
};

class RangeSelection : public Operator {
  public:
   struct Range {
      uint32_t start;  // global sparse row ids, [start, end)
      uint32_t end;
   };
   std::vector<Range> ranges;
   uint32_t begin_of_layout;  // first_chunk << 16 (*row_layout.begin(), range_selection.cpp:97)
   uint32_t end_of_layout;    // (first_chunk + numChunks) << 16, range_selection.cpp:103-106
   RangeSelection(std::vector<Range>&& ranges, uint32_t begin_of_layout, uint32_t end_of_layout)
       : ranges(std::move(ranges)),
         begin_of_layout(begin_of_layout),
         end_of_layout(end_of_layout) {}
   OperatorType type() const override { return RANGE_SELECTION; }
   std::string toString() const override { return "RangeSelection"; }
   void lower(ProgramBuilder& program) const override;
};

// A Predicate of Selection (selection.h:24-43): IsInCoveredRegion over a sequence column (is_in_covered_region.cpp:31-62)
// or, kind == COMPARE, a comparison over a value column -- CompareToValueSelection (selection.h:76-166) and StringInSet
// (equals.cpp:124-156). On the device both are leaves of the filter program (PUSH_COVERED / PUSH_COMPARE), so the
// reference's "evaluate the child, then match the predicates row by row on the host" stage (selection.cpp:94-141) is one
// program: child AND predicate_1 AND ...
struct CoveragePredicate {
   int device_column;
   uint32_t position_idx;
   bool is_covered;  // IS_COVERED / IS_NOT_COVERED
   enum Kind : uint8_t { COVERAGE, COMPARE } kind = COVERAGE;
   // COMPARE
   int value_column = -1;       // device index of the value column
   uint8_t comparator = 0;      // silo_comparator
   bool is_signed = false;      // Date32 / integers
   bool with_nulls = false;     // what a null row gives (selection.h:113-115)
   uint32_t value = 0;
   std::vector<uint32_t> set = {};   // IN_SET: ascending dictionary ids
   std::string display = {};         // Predicate::toString()
   [[nodiscard]] CoveragePredicate negated() const;  // Predicate::negate()
};

class Selection : public Operator {
  public:
   std::optional<std::unique_ptr<Operator>> child_operator;
   std::vector<CoveragePredicate> predicates;
   Selection(std::optional<std::unique_ptr<Operator>> child_operator, std::vector<CoveragePredicate> predicates)
       : child_operator(std::move(child_operator)),
         predicates(std::move(predicates)) {}
   explicit Selection(CoveragePredicate predicate) { predicates.push_back(predicate); }
   OperatorType type() const override { return SELECTION; }
   std::string toString() const override;
   void lower(ProgramBuilder& program) const override;
};

}  // namespace silo_host
