#include "operators.h"

#include <cstdio>

#include <algorithm>
#include <cstring>
#include <map>

namespace silo_host {

// ---- DeviceBitmap ------------------------------------------------------------------------------

DeviceBitmap::DeviceBitmap(silo_gpu_filter* filter, uint64_t cardinality)
    : handle(filter, [](silo_gpu_filter* owned) { silo_gpu_filter_free(owned); }),
      cached_cardinality(cardinality) {}

std::vector<uint64_t> DeviceBitmap::toWords(size_t n_chunks) const {
   std::vector<uint64_t> words(n_chunks * 1024);
   throwOnDeviceError(silo_gpu_filter_download(handle.get(), words.data()));
   return words;
}

// ---- ProgramBuilder ----------------------------------------------------------------------------

void ProgramBuilder::emit(uint8_t opcode, uint8_t flags, uint16_t column, uint32_t a, uint64_t b) {
   silo_filter_instr instr{};
   instr.opcode = opcode;
   instr.flags = flags;
   instr.column = column;
   instr.a = a;
   instr.b = b;
   instrs.push_back(instr);
}

uint64_t ProgramBuilder::addBlob(const void* data, size_t bytes, size_t alignment) {
   blob.resize((blob.size() + alignment - 1) / alignment * alignment, 0);
   const uint64_t offset = blob.size();
   const auto* src = static_cast<const uint8_t*>(data);
   blob.insert(blob.end(), src, src + bytes);
   return offset;
}

uint32_t ProgramBuilder::addBitmap(const std::vector<uint8_t>& portable_roaring_bytes) {
   for (size_t i = 0; i < bitmaps.size(); ++i) {
      if (bitmaps[i].data == portable_roaring_bytes.data()) {
         return static_cast<uint32_t>(i);
      }
   }
   bitmaps.push_back(silo_roaring_bytes{portable_roaring_bytes.data(), portable_roaring_bytes.size()});
   return static_cast<uint32_t>(bitmaps.size() - 1);
}

// ---- Operator ----------------------------------------------------------------------------------

silo_filter_program Operator::lowerProgram(const Table& table, ProgramBuilder& builder) const {
   builder.table = &table;
   lower(builder);
   silo_filter_program program{};
   program.struct_size = sizeof(silo_filter_program);
   program.n_instrs = static_cast<uint32_t>(builder.instrs.size());
   program.instrs = builder.instrs.data();
   program.blob = builder.blob.data();
   program.blob_bytes = builder.blob.size();
   program.n_bitmaps = static_cast<uint32_t>(builder.bitmaps.size());
   program.bitmaps = builder.bitmaps.data();
   return program;
}

DeviceBitmap Operator::evaluate(const Table& table) const {
   const double lower_begin = nowMicroseconds();
   ProgramBuilder builder;
   const silo_filter_program program = lowerProgram(table, builder);
   silo_gpu_filter* filter = nullptr;
   uint64_t cardinality = 0;
   const double eval_begin = nowMicroseconds();
   throwOnDeviceError(silo_gpu_filter_eval(table.deviceTable(), &program, &filter, &cardinality));
   lastQueryProfile().compile_us += eval_begin - lower_begin;
   lastQueryProfile().filter_us = nowMicroseconds() - eval_begin;
   return DeviceBitmap{filter, cardinality};
}

namespace {
template <typename T>
std::unique_ptr<T> downcast(std::unique_ptr<Operator>&& op) {
   return std::unique_ptr<T>(static_cast<T*>(op.release()));
}

constexpr size_t WIDE_UNION_MIN_CHILDREN = 24;  // from here on a Union is lowered as "at least 1 of n"

void lowerCounting(
   ProgramBuilder& program,
   const OperatorVector& non_negated_children,
   const OperatorVector& negated_children,
   uint32_t number_of_matchers,
   bool match_exactly,
   const SymbolScanSpan* span = nullptr
);

// whether lowering the tree emits a counter program (THR_BEGIN .. THR_END); those do not nest
bool containsThreshold(const Operator& node) {
   auto any = [](const OperatorVector& nodes) {
      return std::any_of(nodes.begin(), nodes.end(), [](const auto& child) { return containsThreshold(*child); });
   };
   switch (node.type()) {
      case THRESHOLD:
         return true;
      case UNION: {
         const auto& children = static_cast<const Union&>(node).children;
         return static_cast<const Union&>(node).span != nullptr || children.size() >= WIDE_UNION_MIN_CHILDREN || any(children);
      }
      case INTERSECTION: {
         const auto& intersection = static_cast<const Intersection&>(node);
         return any(intersection.children) || any(intersection.negated_children);
      }
      case COMPLEMENT:
         return containsThreshold(*static_cast<const Complement&>(node).child);
      case SELECTION: {
         const auto& child = static_cast<const Selection&>(node).child_operator;
         return child.has_value() && containsThreshold(**child);
      }
      default:
         return false;
   }
}
}  // namespace

std::unique_ptr<Operator> Operator::negate(std::unique_ptr<Operator>&& some_operator) {
   switch (some_operator->type()) {
      case EMPTY:
         return std::make_unique<Full>();  // empty.cpp:29-31
      case FULL:
         return std::make_unique<Empty>();  // full.cpp:31-33
      case COMPLEMENT:  // complement.cpp:58-60
         return std::move(downcast<Complement>(std::move(some_operator))->child);
      case RANGE_SELECTION: {  // range_selection.cpp:89-111
         auto range_selection = downcast<RangeSelection>(std::move(some_operator));
         std::vector<RangeSelection::Range> new_ranges;
         uint32_t last_end = range_selection->begin_of_layout;
         if (range_selection->begin_of_layout != range_selection->end_of_layout) {
            for (const auto& current : range_selection->ranges) {
               if (last_end != current.start) {
                  new_ranges.push_back({last_end, current.start});
               }
               last_end = current.end;
            }
            if (last_end != range_selection->end_of_layout) {
               new_ranges.push_back({last_end, range_selection->end_of_layout});
            }
         }
         return std::make_unique<RangeSelection>(
            std::move(new_ranges), range_selection->begin_of_layout, range_selection->end_of_layout
         );
      }
      case SELECTION: {  // selection.cpp:143-151
         auto* selection = static_cast<Selection*>(some_operator.get());
         if (!selection->child_operator.has_value() && selection->predicates.size() == 1 &&
             (selection->predicates[0].kind == CoveragePredicate::COVERAGE || selection->predicates[0].comparator < 6)) {
            return std::make_unique<Selection>(selection->predicates[0].negated());
         }
         return std::make_unique<Complement>(std::move(some_operator));
      }
      case INDEX_SCAN:
      case INTERSECTION:
      case THRESHOLD:
      case UNION:
      case BITMAP_PRODUCER:
         return std::make_unique<Complement>(std::move(some_operator));
   }
   throw std::logic_error("unreachable operator type");
}

CoveragePredicate CoveragePredicate::negated() const {
   CoveragePredicate result = *this;
   if (kind == COVERAGE) {
      result.is_covered = !is_covered;  // is_in_covered_region.cpp:69-73
      return result;
   }
   // selection.h:133-166: the opposite comparator, and what a null row gives flips too
   static constexpr uint8_t OPPOSITE[6] = {SILO_CMP_NOT_EQUALS, SILO_CMP_EQUALS, SILO_CMP_HIGHER_OR_EQUALS, SILO_CMP_HIGHER,
                                           SILO_CMP_LESS_OR_EQUALS, SILO_CMP_LESS};
   if (comparator >= 6) {
      throw std::logic_error("a set / range predicate is negated by a Complement around its Selection");
   }
   static const char* const NAMES[6] = {"=", "!=", "<", "<=", ">", ">="};
   const std::string from = std::string(" ") + NAMES[comparator] + " ";
   const std::string to = std::string(" ") + NAMES[OPPOSITE[comparator]] + " ";
   const size_t at = result.display.find(from);
   if (at != std::string::npos) {
      result.display.replace(at, from.size(), to);
   }
   result.comparator = OPPOSITE[comparator];
   result.with_nulls = !with_nulls;
   return result;
}

void Empty::lower(ProgramBuilder& program) const {
   program.emit(SILO_OP_PUSH_EMPTY);
}

void Full::lower(ProgramBuilder& program) const {
   program.emit(SILO_OP_PUSH_FULL);
}

// ---- IndexScan ---------------------------------------------------------------------------------

std::unique_ptr<IndexScan> IndexScan::overSymbols(int device_column, uint32_t position_idx, uint32_t symbol_mask) {
   auto scan = std::make_unique<IndexScan>();
   scan->source = Source::SYMBOLS;
   scan->device_column = device_column;
   scan->position_idx = position_idx;
   scan->symbol_mask = symbol_mask;
   return scan;
}

std::unique_ptr<IndexScan> IndexScan::overBitmap(const std::vector<uint8_t>* portable_roaring_bytes) {
   auto scan = std::make_unique<IndexScan>();
   scan->source = Source::BITMAP;
   scan->bitmap_bytes = portable_roaring_bytes;
   return scan;
}

std::unique_ptr<IndexScan> IndexScan::overIndexBitmap(uint32_t device_id) {
   auto scan = std::make_unique<IndexScan>();
   scan->source = Source::INDEX_BITMAP;
   scan->index_bitmap_id = device_id;
   return scan;
}

std::unique_ptr<IndexScan> IndexScan::overNulls(int device_column) {
   auto scan = std::make_unique<IndexScan>();
   scan->source = Source::NULLS;
   scan->device_column = device_column;
   return scan;
}

std::string IndexScan::toString() const {
   switch (source) {
      case Source::SYMBOLS:
      {
         // (the reference prints its logical equivalent and the bitmap's cardinality, index_scan.cpp:31-37;
         // here the bitmap is a view of device-resident containers, so the scan is named by what it reads)
         char mask_hex[16];
         std::snprintf(mask_hex, sizeof(mask_hex), "0x%x", symbol_mask);
         return "IndexScan(column " + std::to_string(device_column) + ", position " +
                std::to_string(position_idx + 1) + ", symbols " + mask_hex + ")";
      }
      case Source::BITMAP:
         return "IndexScan(bitmap, " + std::to_string(bitmap_bytes->size()) + " bytes)";
      case Source::INDEX_BITMAP:
         return "IndexScan(device-resident index bitmap " + std::to_string(index_bitmap_id) + ")";
      case Source::NULLS:
         return "IndexScan(nulls of column " + std::to_string(device_column) + ")";
   }
   return "IndexScan";
}

void IndexScan::lower(ProgramBuilder& program) const {
   switch (source) {
      case Source::SYMBOLS:
         if (symbol_mask == 0) {
            program.emit(SILO_OP_PUSH_EMPTY);
         } else {
            program.emit(SILO_OP_PUSH_SYMBOLS, 0, static_cast<uint16_t>(device_column), position_idx, symbol_mask);
         }
         break;
      case Source::BITMAP:
         program.emit(SILO_OP_PUSH_BITMAP, 0, 0, program.addBitmap(*bitmap_bytes), 0);
         break;
      case Source::INDEX_BITMAP:
         program.emit(SILO_OP_PUSH_INDEX_BITMAP, 0, 0, index_bitmap_id, 0);
         break;
      case Source::NULLS:
         program.emit(SILO_OP_PUSH_NULLS, 0, static_cast<uint16_t>(device_column));
         break;
   }
}

// ---- Intersection / Union / Complement ------------------------------------------------------------

Intersection::Intersection(OperatorVector&& children_, OperatorVector&& negated_children_)
    : children(std::move(children_)),
      negated_children(std::move(negated_children_)) {
   if (children.empty()) {
      throw QueryCompilationException(
         "Compilation bug: Intersection without non-negated children is not allowed. "
         "Should be compiled as a union."
      );
   }
   if (children.size() + negated_children.size() < 2) {
      throw QueryCompilationException("Compilation bug: Intersection needs at least two children.");
   }
}

std::string Intersection::toString() const {  // intersection.cpp:45-53
   return "Intersection(non_negated: (" + joinWithLimit(children) + ") negated: (" + joinWithLimit(negated_children) + ") )";
}

void Intersection::lower(ProgramBuilder& program) const {
   // The reference orders operands by cardinality to keep roaring intermediates small
   // (intersection.cpp:73-87); on dense tiles the order is irrelevant.
   children[0]->lower(program);
   for (size_t i = 1; i < children.size(); ++i) {
      children[i]->lower(program);
      program.emit(SILO_OP_AND);
   }
   for (const auto& child : negated_children) {
      child->lower(program);
      program.emit(SILO_OP_ANDNOT);
   }
}

std::unique_ptr<Operator> SymbolScanSpan::materialise(size_t index) const {
   const Leaf& leaf = leaves.at(index);
   if (!leaf.includes_reference) {
      return IndexScan::overSymbols(device_column, leaf.position_idx, leaf.mask);
   }
   OperatorVector keep;
   keep.push_back(std::make_unique<Selection>(CoveragePredicate{device_column, leaf.position_idx, true}));
   OperatorVector drop;
   drop.push_back(IndexScan::overSymbols(device_column, leaf.position_idx, all_symbols_mask & ~leaf.mask & ~missing_bit));
   return std::make_unique<Intersection>(std::move(keep), std::move(drop));
}

// what joinWithLimit prints for these children when `already_printed` items of the same list came before them
std::string SymbolScanSpan::joinedStrings(const std::string& delimiter, size_t already_printed, size_t limit) const {
   std::string res;
   const size_t room = already_printed < limit ? limit - already_printed : 0;
   const size_t items_to_print = std::min(room, leaves.size());
   for (size_t i = 0; i < items_to_print; ++i) {
      if (already_printed + i > 0) {
         res += delimiter;
      }
      res += materialise(i)->toString();
   }
   return res;
}

namespace {
// joinWithLimit over `children` followed by the children of `span`
std::string joinWithSpan(const OperatorVector& children, const SymbolScanSpan* span, const std::string& delimiter = ", ", size_t limit = 10) {
   if (span == nullptr) {
      return joinWithLimit(children, delimiter, limit);
   }
   const size_t total = children.size() + span->size();
   std::string res;
   const size_t own = std::min(children.size(), limit);
   for (size_t i = 0; i < own; ++i) {
      if (i > 0) {
         res += delimiter;
      }
      res += children[i]->toString();
   }
   res += span->joinedStrings(delimiter, own, limit);
   const size_t printed = std::min(total, limit);
   if (total > printed) {
      res += delimiter + "... (" + std::to_string(total - printed) + " more)";
   }
   return res;
}
}  // namespace

std::string Union::toString() const {  // union.cpp:23-28
   return "(" + joinWithSpan(children, span.get(), " | ") + ")";
}

void Union::lower(ProgramBuilder& program) const {
   if (span != nullptr && !program.inside_counter_program) {
      static const OperatorVector NO_NEGATED;
      lowerCounting(program, children, NO_NEGATED, 1, false, span.get());
      return;
   }
   if (span != nullptr) {  // inside another counter program: a plain chain of ORs over the materialised leaves
      bool have_tile = false;
      for (const auto& child : children) {
         child->lower(program);
         if (have_tile) {
            program.emit(SILO_OP_OR);
         }
         have_tile = true;
      }
      for (size_t i = 0; i < span->size(); ++i) {
         span->materialise(i)->lower(program);
         if (have_tile) {
            program.emit(SILO_OP_OR);
         }
         have_tile = true;
      }
      if (!have_tile) {
         program.emit(SILO_OP_PUSH_EMPTY);
      }
      return;
   }
   if (children.empty()) {
      program.emit(SILO_OP_PUSH_EMPTY);
      return;
   }
   // A wide union (NOf with one matcher becomes an Or, nof.cpp; a MutationProfile of distance 0 has one
   // child per genome position) is "at least 1 of n": one sweep of the chunk's containers into the row
   // counters instead of n searches + ORs. Counter programs do not nest, so only when no child is one.
   const bool any_threshold_child =
      std::any_of(children.begin(), children.end(), [](const auto& child) { return containsThreshold(*child); });
   if (children.size() >= WIDE_UNION_MIN_CHILDREN && children.size() <= 65535 && !any_threshold_child &&
       !program.inside_counter_program) {
      static const OperatorVector NO_CHILDREN;
      lowerCounting(program, children, NO_CHILDREN, 1, false);
      return;
   }
   children[0]->lower(program);
   for (size_t i = 1; i < children.size(); ++i) {
      children[i]->lower(program);
      program.emit(SILO_OP_OR);
   }
}

std::unique_ptr<Complement> Complement::fromDeMorgan(OperatorVector disjunction) {
   OperatorVector non_negated_child_operators;
   OperatorVector negated_child_operators;
   for (auto& disjunction_child : disjunction) {
      if (disjunction_child->type() == COMPLEMENT) {
         negated_child_operators.emplace_back(Operator::negate(std::move(disjunction_child)));
      } else {
         non_negated_child_operators.push_back(std::move(disjunction_child));
      }
   }
   // a disjunction with negated members becomes  !(negated... & !non_negated...)
   auto intersection = std::make_unique<Intersection>(
      std::move(negated_child_operators), std::move(non_negated_child_operators)
   );
   return std::make_unique<Complement>(std::move(intersection));
}

void Complement::lower(ProgramBuilder& program) const {
   child->lower(program);
   program.emit(SILO_OP_NOT);
}

// ---- RangeSelection / Selection ------------------------------------------------------------------

void RangeSelection::lower(ProgramBuilder& program) const {
   std::vector<uint32_t> flat;
   flat.reserve(ranges.size() * 2);
   for (const auto& range : ranges) {
      flat.push_back(range.start);
      flat.push_back(range.end);
   }
   const uint64_t offset = program.addBlob(flat.data(), flat.size() * sizeof(uint32_t), 8);  // (the kernel loads {start, end} pairs)
   program.emit(SILO_OP_PUSH_RANGES, 0, 0, static_cast<uint32_t>(ranges.size()), offset);
}

std::string Selection::toString() const {  // selection.cpp:75-88, is_in_covered_region.cpp:25-29
   std::string res = "Select[";
   for (size_t i = 0; i < predicates.size(); ++i) {
      if (predicates[i].kind == CoveragePredicate::COMPARE) {
         res += std::string(i > 0 ? "," : "") + predicates[i].display;
         continue;
      }
      res += std::string(i > 0 ? "," : "") + (predicates[i].is_covered ? "" : "!") + "IsInCoveredRegion(" +
             std::to_string(predicates[i].position_idx) + ")";
   }
   res += "](";
   if (child_operator.has_value()) {
      res += child_operator.value()->toString();
   }
   return res + ")";
}

void Selection::lower(ProgramBuilder& program) const {
   // selection.cpp:94-141 picks between row-wise matching and makeBitmap by cardinality; both give
   // child AND predicate_1 AND ... which is what is emitted here.
   bool have_tile = false;
   if (child_operator.has_value()) {
      child_operator.value()->lower(program);
      have_tile = true;
   }
   for (const auto& predicate : predicates) {
      if (predicate.kind == CoveragePredicate::COMPARE) {
         const uint8_t flags = static_cast<uint8_t>(predicate.comparator | (predicate.is_signed ? SILO_CMP_SIGNED : 0) | (predicate.with_nulls ? SILO_CMP_WITH_NULLS : 0));
         if (predicate.comparator == SILO_CMP_IN_SET) {
            const uint64_t offset = program.addBlob(predicate.set.data(), predicate.set.size() * sizeof(uint32_t), 4);
            program.emit(SILO_OP_PUSH_COMPARE, flags, static_cast<uint16_t>(predicate.value_column), static_cast<uint32_t>(predicate.set.size()), offset);
         } else {
            program.emit(SILO_OP_PUSH_COMPARE, flags, static_cast<uint16_t>(predicate.value_column), predicate.value, 0);
         }
         if (have_tile) {
            program.emit(SILO_OP_AND);
         }
         have_tile = true;
         continue;
      }
      program.emit(
         SILO_OP_PUSH_COVERED,
         predicate.is_covered ? 0 : 1,
         static_cast<uint16_t>(predicate.device_column),
         predicate.position_idx
      );
      if (have_tile) {
         program.emit(SILO_OP_AND);
      }
      have_tile = true;
   }
}

// ---- Threshold -----------------------------------------------------------------------------------

Threshold::Threshold(
   OperatorVector&& non_negated_children_,
   OperatorVector&& negated_children_,
   uint32_t number_of_matchers,
   bool match_exactly,
   std::shared_ptr<const SymbolScanSpan> span_
)
    : non_negated_children(std::move(non_negated_children_)),
      negated_children(std::move(negated_children_)),
      number_of_matchers(number_of_matchers),
      match_exactly(match_exactly),
      span(std::move(span_)) {
   if (number_of_matchers >= non_negated_children.size() + negated_children.size() + (span != nullptr ? span->size() : 0)) {
      throw QueryCompilationException(
         "Compilation Error: number_of_matchers must be less than the number of children of a "
         "threshold expression"
      );
   }
   if (number_of_matchers == 0) {
      throw QueryCompilationException("Compilation Error: number_of_matchers must be greater than zero");
   }
}

std::string Threshold::toString() const {  // threshold.cpp:45-58
   return std::string("Threshold(") + (match_exactly ? "=" : ">=") + std::to_string(number_of_matchers) + "-of " +
          "non_negated: (" + joinWithSpan(non_negated_children, span.get()) + ") negated: (" + joinWithLimit(negated_children) + ") )";
}

namespace {

// Recognises  Selection[IsCovered(col,p)] minus IndexScan(symbols at (col,p))  — what
// compileWithReference builds (symbol_in_set.cpp:179-206). Every stored symbol is inside the covered
// region, so the child contributes  covered(p) - [row holds one of the scanned symbols].
bool matchCoveredMinusSymbols(const Operator& op, CoveragePredicate& predicate, uint32_t& mask) {
   if (op.type() != INTERSECTION) {
      return false;
   }
   const auto& intersection = static_cast<const Intersection&>(op);
   if (intersection.children.size() != 1 || intersection.negated_children.size() != 1) {
      return false;
   }
   const Operator& left = *intersection.children[0];
   const Operator& right = *intersection.negated_children[0];
   if (left.type() != SELECTION || right.type() != INDEX_SCAN) {
      return false;
   }
   const auto& selection = static_cast<const Selection&>(left);
   const auto& scan = static_cast<const IndexScan&>(right);
   if (selection.child_operator.has_value() || selection.predicates.size() != 1 ||
       selection.predicates[0].kind != CoveragePredicate::COVERAGE || !selection.predicates[0].is_covered || scan.source != IndexScan::Source::SYMBOLS ||
       scan.device_column != selection.predicates[0].device_column ||
       scan.position_idx != selection.predicates[0].position_idx) {
      return false;
   }
   predicate = selection.predicates[0];
   mask = scan.symbol_mask;
   return true;
}

constexpr size_t PROFILE_MIN_LEAVES = 24;  // below this, per-position binary searches are cheaper

struct FusedLeaves {
   std::vector<std::pair<uint32_t, uint32_t>> adds;  // (position, mask)
   std::vector<std::pair<uint32_t, uint32_t>> subs;
   std::vector<uint32_t> covered_positions;
};

// "at least / exactly k of the children contain the row" as a counter program. Shared by Threshold
// and by wide Unions (k = 1).
void lowerCounting(
   ProgramBuilder& program,
   const OperatorVector& non_negated_children,
   const OperatorVector& negated_children,
   uint32_t number_of_matchers,
   bool match_exactly,
   const SymbolScanSpan* span
) {
   // Threshold::evaluate (threshold.cpp:64-138) is a DP over k roaring bitmaps whose result is
   // "at least / exactly k of the children contain the row". The device keeps one u16 counter per
   // row instead; leaves over the vertical index are scattered into the counters without ever
   // materialising a tile.
   std::map<int, FusedLeaves> fused;  // by device column
   std::vector<const Operator*> generic;
   for (const auto& child : non_negated_children) {
      CoveragePredicate predicate{};
      uint32_t mask = 0;
      if (child->type() == INDEX_SCAN &&
          static_cast<const IndexScan&>(*child).source == IndexScan::Source::SYMBOLS) {
         const auto& scan = static_cast<const IndexScan&>(*child);
         if (scan.symbol_mask != 0) {
            fused[scan.device_column].adds.emplace_back(scan.position_idx, scan.symbol_mask);
         }
      } else if (matchCoveredMinusSymbols(*child, predicate, mask)) {
         FusedLeaves& leaves = fused[predicate.device_column];
         leaves.covered_positions.push_back(predicate.position_idx);
         if (mask != 0) {
            leaves.subs.emplace_back(predicate.position_idx, mask);
         }
      } else {
         generic.push_back(child.get());
      }
   }
   if (span != nullptr) {  // the same classification, straight from the compact leaves
      FusedLeaves& leaves = fused[span->device_column];
      leaves.adds.reserve(leaves.adds.size() + span->size());
      for (const SymbolScanSpan::Leaf& leaf : span->leaves) {
         if (!leaf.includes_reference) {
            if (leaf.mask != 0) {
               leaves.adds.emplace_back(leaf.position_idx, leaf.mask);
            }
         } else {
            leaves.covered_positions.push_back(leaf.position_idx);
            const uint32_t others = span->all_symbols_mask & ~leaf.mask & ~span->missing_bit;
            if (others != 0) {
               leaves.subs.emplace_back(leaf.position_idx, others);
            }
         }
      }
   }
   uint64_t bias = 0;
   for (const auto& [column, leaves] : fused) {
      bias += leaves.subs.size();
   }
   // Counter programs do not nest on the device (one counter tile per CTA), but the reference's Threshold takes any
   // operator as a child (threshold.cpp:64-138) -- an NOf inside an NOf, a MutationProfile inside an NOf, a wide Union:
   // such children are evaluated to tiles BEFORE this program's THR_BEGIN (their own counter programs run to their
   // THR_END first) and are added from the stack right behind it, last pushed first.
   std::vector<std::pair<const Operator*, bool>> plain;   // (child, negated)
   std::vector<bool> nested_negated;
   for (const Operator* child : generic) {
      if (containsThreshold(*child)) {
         child->lower(program);
         nested_negated.push_back(false);
      } else {
         plain.emplace_back(child, false);
      }
   }
   for (const auto& child : negated_children) {
      if (containsThreshold(*child)) {
         child->lower(program);
         nested_negated.push_back(true);
      } else {
         plain.emplace_back(child.get(), true);
      }
   }
   program.emit(SILO_OP_THR_BEGIN, match_exactly ? 1 : 0, 0, number_of_matchers, bias);
   for (auto negated = nested_negated.rbegin(); negated != nested_negated.rend(); ++negated) {
      program.emit(SILO_OP_THR_ADD, *negated ? 1 : 0);
   }
   const bool was_inside = program.inside_counter_program;
   program.inside_counter_program = true;  // (a narrow Union below stays a chain of ORs)
   for (const auto& [child, negated] : plain) {
      child->lower(program);
      program.emit(SILO_OP_THR_ADD, negated ? 1 : 0);
   }
   program.inside_counter_program = was_inside;
   for (auto& [column, leaves] : fused) {
      const auto device_column = static_cast<uint16_t>(column);
      std::vector<std::pair<uint32_t, uint32_t>> single_adds;
      std::vector<std::pair<uint32_t, uint32_t>> single_subs;
      if (leaves.adds.size() + leaves.subs.size() >= PROFILE_MIN_LEAVES) {
         // one streaming pass over the chunk's containers; the table holds at most one add mask and
         // one disjoint sub mask per position, anything else falls back to single leaves
         const size_t genome_length = program.table->columns.at(static_cast<size_t>(column)).reference_sequence.size();
         std::vector<uint32_t> table(2 * genome_length, 0);
         for (const auto& [position, mask] : leaves.adds) {
            if (table[2 * position] == 0 && (table[2 * position + 1] & mask) == 0) {
               table[2 * position] = mask;
            } else {
               single_adds.emplace_back(position, mask);
            }
         }
         for (const auto& [position, mask] : leaves.subs) {
            if (table[2 * position + 1] == 0 && (table[2 * position] & mask) == 0) {
               table[2 * position + 1] = mask;
            } else {
               single_subs.emplace_back(position, mask);
            }
         }
         const uint64_t offset = program.addBlob(table.data(), table.size() * sizeof(uint32_t), 8);
         program.emit(SILO_OP_THR_PROFILE, 0, device_column, 0, offset);
      } else {
         single_adds = leaves.adds;
         single_subs = leaves.subs;
      }
      for (const auto& [position, mask] : single_adds) {
         program.emit(SILO_OP_THR_ADD_SYMBOLS, 0, device_column, position, mask);
      }
      for (const auto& [position, mask] : single_subs) {
         program.emit(SILO_OP_THR_ADD_SYMBOLS, 1, device_column, position, mask);
      }
      // covered positions: strictly ascending lists; a position used by several children goes into
      // further lists
      std::vector<uint32_t> remaining = leaves.covered_positions;
      std::sort(remaining.begin(), remaining.end());
      while (!remaining.empty()) {
         std::vector<uint32_t> unique_positions;
         std::vector<uint32_t> duplicates;
         for (uint32_t position : remaining) {
            if (!unique_positions.empty() && unique_positions.back() == position) {
               duplicates.push_back(position);
            } else {
               unique_positions.push_back(position);
            }
         }
         const uint64_t offset =
            program.addBlob(unique_positions.data(), unique_positions.size() * sizeof(uint32_t), 4);
         program.emit(
            SILO_OP_THR_ADD_COVERED, 0, device_column, static_cast<uint32_t>(unique_positions.size()), offset
         );
         remaining = std::move(duplicates);
      }
   }
   program.emit(SILO_OP_THR_END);
}

}  // namespace

void Threshold::lower(ProgramBuilder& program) const {
   lowerCounting(program, non_negated_children, negated_children, number_of_matchers, match_exactly, span.get());
}

}  // namespace silo_host
