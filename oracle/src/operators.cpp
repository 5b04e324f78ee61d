// ORACLE — TEST INFRASTRUCTURE ONLY. See operators.h for the reference file:line map.
#include "operators.h"

#include <algorithm>

namespace oracle {

namespace {
template <typename T>
std::unique_ptr<T> downcast(std::unique_ptr<Operator>&& op) {
   return std::unique_ptr<T>(static_cast<T*>(op.release()));
}
}  // namespace

// operator.cpp:19-67
std::unique_ptr<Operator> Operator::negate(std::unique_ptr<Operator>&& some_operator) {
   switch (some_operator->type()) {
      case EMPTY: {
         auto empty = downcast<Empty>(std::move(some_operator));
         return std::make_unique<Full>(std::move(empty->row_layout));
      }
      case FULL: {
         auto full = downcast<Full>(std::move(some_operator));
         return std::make_unique<Empty>(std::move(full->row_layout));
      }
      case INDEX_SCAN: {
         auto row_layout = static_cast<IndexScan*>(some_operator.get())->row_layout;
         return std::make_unique<Complement>(std::move(some_operator), std::move(row_layout));
      }
      case INTERSECTION: {
         auto row_layout = static_cast<Intersection*>(some_operator.get())->row_layout;
         return std::make_unique<Complement>(std::move(some_operator), std::move(row_layout));
      }
      case COMPLEMENT: {
         auto complement = downcast<Complement>(std::move(some_operator));
         return std::move(complement->child);
      }
      case RANGE_SELECTION: {
         // range_selection.cpp:89-111
         auto range_selection = downcast<RangeSelection>(std::move(some_operator));
         std::vector<RangeSelection::Range> new_ranges;
         if (range_selection->row_layout.numChunks() == 0) {
            return std::make_unique<RangeSelection>(
               std::move(new_ranges), std::move(range_selection->row_layout)
            );
         }
         uint32_t last_end = 0;  // *row_layout.begin(): chunks are never empty
         for (const auto& current : range_selection->ranges) {
            if (last_end != current.start) {
               new_ranges.push_back({last_end, current.start});
            }
            last_end = current.end;
         }
         const auto ranges_end = static_cast<uint32_t>(range_selection->row_layout.numChunks()) << 16;
         if (last_end != ranges_end) {
            new_ranges.push_back({last_end, ranges_end});
         }
         return std::make_unique<RangeSelection>(
            std::move(new_ranges), std::move(range_selection->row_layout)
         );
      }
      case SELECTION: {
         // selection.cpp:143-151
         auto* selection = static_cast<Selection*>(some_operator.get());
         auto row_layout = selection->row_layout;
         if (!selection->child_operator.has_value() && selection->predicates.size() == 1) {
            return std::make_unique<Selection>(
               selection->predicates.at(0)->negate(), std::move(row_layout)
            );
         }
         return std::make_unique<Complement>(std::move(some_operator), std::move(row_layout));
      }
      case THRESHOLD: {
         auto row_layout = static_cast<Threshold*>(some_operator.get())->row_layout;
         return std::make_unique<Complement>(std::move(some_operator), std::move(row_layout));
      }
      case UNION: {
         auto row_layout = static_cast<Union*>(some_operator.get())->row_layout;
         return std::make_unique<Complement>(std::move(some_operator), std::move(row_layout));
      }
      case BITMAP_PRODUCER:
         break;
   }
   throw std::runtime_error("unreachable operator type");
}

// ---- Intersection (intersection.cpp:19-110) ----

Intersection::Intersection(
   OperatorVector&& children_,
   OperatorVector&& negated_children_,
   RowLayout row_layout
)
    : children(std::move(children_)),
      negated_children(std::move(negated_children_)),
      row_layout(std::move(row_layout)) {
   if (this->children.empty()) {
      throw QueryCompilationException(
         "Compilation bug: Intersection without non-negated children is not allowed. "
         "Should be compiled as a union."
      );
   }
   if (this->children.size() + this->negated_children.size() < 2) {
      throw QueryCompilationException("Compilation bug: Intersection needs at least two children.");
   }
}

CowBitmap Intersection::evaluate() const {
   std::vector<CowBitmap> children_bm;
   children_bm.reserve(children.size());
   for (const auto& child : children) {
      children_bm.push_back(child->evaluate());
   }
   std::vector<CowBitmap> negated_children_bm;
   negated_children_bm.reserve(negated_children.size());
   for (const auto& child : negated_children) {
      negated_children_bm.push_back(child->evaluate());
   }
   std::sort(children_bm.begin(), children_bm.end(), [](const CowBitmap& a, const CowBitmap& b) {
      return a.cardinality() < b.cardinality();
   });
   std::sort(
      negated_children_bm.begin(),
      negated_children_bm.end(),
      [](const CowBitmap& a, const CowBitmap& b) { return a.cardinality() > b.cardinality(); }
   );
   CowBitmap result = std::move(children_bm[0]);
   for (size_t i = 1; i < children_bm.size(); i++) {
      result &= children_bm[i];
   }
   for (auto& neg_bm : negated_children_bm) {
      result -= neg_bm;
   }
   return result;
}

std::string Intersection::toString() const {
   std::string res = "Intersection(non_negated: (";
   for (const auto& child : children) {
      res += child->toString() + ", ";
   }
   res += ") negated: (";
   for (const auto& child : negated_children) {
      res += child->toString() + ", ";
   }
   return res + "))";
}

// ---- Union (union.cpp:34-42) ----

CowBitmap Union::evaluate() const {
   std::vector<CowBitmap> child_res;
   child_res.reserve(children.size());
   for (const auto& child : children) {
      child_res.push_back(child->evaluate());
   }
   return CowBitmap::fastUnion(child_res);
}

std::string Union::toString() const {
   std::string res = "(";
   for (const auto& child : children) {
      res += child->toString() + " | ";
   }
   return res + ")";
}

// ---- Complement (complement.cpp:23-56) ----

std::unique_ptr<Complement> Complement::fromDeMorgan(OperatorVector disjunction, RowLayout row_layout) {
   OperatorVector non_negated_child_operators;
   OperatorVector negated_child_operators;
   for (auto& disjunction_child : disjunction) {
      if (disjunction_child->type() == COMPLEMENT) {
         negated_child_operators.emplace_back(Operator::negate(std::move(disjunction_child)));
      } else {
         non_negated_child_operators.push_back(std::move(disjunction_child));
      }
   }
   auto intersection = std::make_unique<Intersection>(
      std::move(negated_child_operators), std::move(non_negated_child_operators), row_layout
   );
   return std::make_unique<Complement>(std::move(intersection), std::move(row_layout));
}

CowBitmap Complement::evaluate() const {
   Roaring result = child->evaluate().toRoaring();
   row_layout.complementInPlace(result);
   return CowBitmap{std::move(result)};
}

// ---- Threshold (threshold.cpp:19-138) ----

Threshold::Threshold(
   OperatorVector&& non_negated_children_,
   OperatorVector&& negated_children_,
   uint32_t number_of_matchers,
   bool match_exactly,
   RowLayout row_layout
)
    : non_negated_children(std::move(non_negated_children_)),
      negated_children(std::move(negated_children_)),
      number_of_matchers(number_of_matchers),
      match_exactly(match_exactly),
      row_layout(std::move(row_layout)) {
   if (number_of_matchers >= this->non_negated_children.size() + this->negated_children.size()) {
      throw QueryCompilationException(
         "Compilation Error: number_of_matchers must be less than the number of children of a "
         "threshold expression"
      );
   }
   if (number_of_matchers == 0) {
      throw QueryCompilationException(
         "Compilation Error: number_of_matchers must be greater than zero"
      );
   }
}

CowBitmap Threshold::evaluate() const {
   const uint32_t dp_table_size = match_exactly ? number_of_matchers + 1 : number_of_matchers;
   std::vector<Roaring> bitmaps(dp_table_size);
   if (non_negated_children.empty()) {
      bitmaps[0] = negated_children[0]->evaluate().toRoaring();
      row_layout.complementInPlace(bitmaps[0]);
   } else {
      bitmaps[0] = non_negated_children[0]->evaluate().toRoaring();
   }
   const int max_table_index = static_cast<int>(dp_table_size - 1);
   const int non_negated_child_count = static_cast<int>(non_negated_children.size());
   const int negated_child_count = static_cast<int>(negated_children.size());
   const int n = static_cast<int>(number_of_matchers);
   const int k = non_negated_child_count + negated_child_count;

   for (int i = 1; i < non_negated_child_count; ++i) {
      const Roaring bitmap = non_negated_children[static_cast<size_t>(i)]->evaluate().toRoaring();
      for (int j = std::min(max_table_index, i); j > std::max(0, n - k + i - 1); --j) {
         bitmaps[static_cast<size_t>(j)] |= bitmaps[static_cast<size_t>(j - 1)] & bitmap;
      }
      if (k - i > n - 1) {
         bitmaps[0] |= bitmap;
      }
   }
   const int took_first_offset = non_negated_children.empty() ? 1 : 0;
   for (int local_i = took_first_offset; local_i < negated_child_count; ++local_i) {
      Roaring bitmap = negated_children[static_cast<size_t>(local_i)]->evaluate().toRoaring();
      const int i = local_i + non_negated_child_count;
      for (int j = std::min(max_table_index, i); j > std::max(0, n - k + i - 1); --j) {
         bitmaps[static_cast<size_t>(j)] |= bitmaps[static_cast<size_t>(j - 1)] - bitmap;
      }
      if (k - i > n - 1) {
         row_layout.complementInPlace(bitmap);
         bitmaps[0] |= bitmap;
      }
   }
   if (match_exactly) {
      bitmaps[number_of_matchers - 1] -= bitmaps[number_of_matchers];
      return CowBitmap(std::move(bitmaps[number_of_matchers - 1]));
   }
   return CowBitmap(std::move(bitmaps.back()));
}

std::string Threshold::toString() const {
   return std::string("Threshold(") + (match_exactly ? "=" : ">=") +
          std::to_string(number_of_matchers) + "-of " +
          std::to_string(non_negated_children.size()) + " non_negated, " +
          std::to_string(negated_children.size()) + " negated)";
}

// ---- RangeSelection (range_selection.cpp:54-87) ----

CowBitmap RangeSelection::evaluate() const {
   Roaring result_bitmap;
   for (const auto& [start, end] : ranges) {
      const uint32_t start_chunk = start >> 16;
      const uint32_t end_chunk = end >> 16;
      if (start_chunk == end_chunk) {
         result_bitmap.addRange(start, end);
      } else {
         const uint32_t end_of_start_chunk =
            (start_chunk << 16) + row_layout.chunkSize(static_cast<uint16_t>(start_chunk));
         if (start != end_of_start_chunk) {
            result_bitmap.addRange(start, end_of_start_chunk);
         }
         for (uint32_t chunk_id = start_chunk + 1; chunk_id < end_chunk; chunk_id++) {
            const uint32_t chunk_start = chunk_id << 16;
            result_bitmap.addRange(
               chunk_start, chunk_start + row_layout.chunkSize(static_cast<uint16_t>(chunk_id))
            );
         }
         const uint32_t start_of_end_chunk = end_chunk << 16;
         if (end != start_of_end_chunk) {
            result_bitmap.addRange(start_of_end_chunk, end);
         }
      }
   }
   return CowBitmap{std::move(result_bitmap)};
}

// ---- Predicates / Selection ----

Roaring Predicate::makeBitmap(const RowLayout& row_layout) const {
   Roaring result;
   for (size_t chunk_id = 0; chunk_id < row_layout.numChunks(); ++chunk_id) {
      for (uint32_t row = 0; row < row_layout.chunk_sizes[chunk_id]; ++row) {
         const uint32_t global = (static_cast<uint32_t>(chunk_id) << 16) | row;
         if (match(global)) {
            result.add(global);
         }
      }
   }
   return result;
}

std::string IsInCoveredRegion::toString() const {
   return std::string(comparator == Comparator::IS_COVERED ? "" : "!") + "IsInCoveredRegion(" +
          std::to_string(position_idx) + ")";
}

bool IsInCoveredRegion::isCovered(uint32_t row_id) const {
   const auto [start, end] = horizontal_coverage_index->coverageRange(row_id);
   if (position_idx < start || position_idx >= end) {
      return false;
   }
   if (auto row_bitmap = horizontal_coverage_index->horizontal_bitmaps.find(row_id);
       row_bitmap != horizontal_coverage_index->horizontal_bitmaps.end()) {
      return !row_bitmap->second.contains(position_idx);
   }
   return true;
}

bool IsInCoveredRegion::match(uint32_t global_row_id) const {
   return isCovered(global_row_id) == (comparator == Comparator::IS_COVERED);
}

Roaring IsInCoveredRegion::makeBitmap(const RowLayout& row_layout) const {
   Roaring coverage_bitmap = horizontal_coverage_index->getCoverageBitmapForPosition(position_idx);
   if (comparator == Comparator::IS_NOT_COVERED) {
      row_layout.complementInPlace(coverage_bitmap);
   }
   return coverage_bitmap;
}

std::unique_ptr<Predicate> IsInCoveredRegion::negate() const {
   return std::make_unique<IsInCoveredRegion>(
      horizontal_coverage_index,
      position_idx,
      comparator == Comparator::IS_COVERED ? Comparator::IS_NOT_COVERED : Comparator::IS_COVERED
   );
}

std::string DateCompare::toString() const {
   static const char* const NAMES[] = {"=", "<", ">", "<=", ">=", "!="};
   return "$date " + column->name + " " + NAMES[static_cast<int>(comparator)] + " " + std::to_string(value);
}

bool DateCompare::match(uint32_t global_row_id) const {
   const size_t row = (*chunk_begin)[global_row_id >> 16] + (global_row_id & 0xFFFFu);
   if (column->is_null[row]) {
      return with_nulls;
   }
   const int32_t stored = column->values[row];
   switch (comparator) {
      case Comparator::EQUALS: return stored == value;
      case Comparator::NOT_EQUALS: return stored != value;
      case Comparator::LESS: return stored < value;
      case Comparator::HIGHER_OR_EQUALS: return stored >= value;
      case Comparator::HIGHER: return stored > value;
      case Comparator::LESS_OR_EQUALS: return stored <= value;
   }
   return false;
}

std::unique_ptr<Predicate> DateCompare::negate() const {
   Comparator opposite = Comparator::EQUALS;
   switch (comparator) {
      case Comparator::EQUALS: opposite = Comparator::NOT_EQUALS; break;
      case Comparator::NOT_EQUALS: opposite = Comparator::EQUALS; break;
      case Comparator::LESS: opposite = Comparator::HIGHER_OR_EQUALS; break;
      case Comparator::HIGHER_OR_EQUALS: opposite = Comparator::LESS; break;
      case Comparator::HIGHER: opposite = Comparator::LESS_OR_EQUALS; break;
      case Comparator::LESS_OR_EQUALS: opposite = Comparator::HIGHER; break;
   }
   return std::make_unique<DateCompare>(column, chunk_begin, opposite, value, !with_nulls);
}

std::string StringInSetPredicate::toString() const {
   std::string joined;
   for (const std::string& value : values) {
      joined += (joined.empty() ? "'" : ",'") + value + "'";
   }
   return "$string " + column->name + (in ? " IN [" : " NOT IN [") + joined + "]";
}

bool StringInSetPredicate::match(uint32_t global_row_id) const {
   const size_t row = (*chunk_begin)[global_row_id >> 16] + (global_row_id & 0xFFFFu);
   const bool in_set = std::find(values.begin(), values.end(), column->values[row]) != values.end();
   return in ? in_set : !in_set;
}

std::unique_ptr<Predicate> StringInSetPredicate::negate() const {
   return std::make_unique<StringInSetPredicate>(column, chunk_begin, !in, values);
}

Selection::Selection(
   std::optional<std::unique_ptr<Operator>> child_operator,
   PredicateVector&& predicates_,
   RowLayout row_layout_
)
    : child_operator(std::move(child_operator)),
      predicates(std::move(predicates_)),
      row_layout(std::move(row_layout_)) {
   const auto row_count = this->row_layout.numRows();
   std::sort(this->predicates.begin(), this->predicates.end(), [row_count](const auto& left, const auto& right) {
      return left->estimateSelectivity(row_count) < right->estimateSelectivity(row_count);
   });
}

Selection::Selection(std::unique_ptr<Predicate> predicate, RowLayout row_layout)
    : row_layout(std::move(row_layout)) {
   predicates.emplace_back(std::move(predicate));
}

CowBitmap Selection::evaluate() const {
   auto matchesAll = [&](size_t first, uint32_t row_id) {
      for (size_t i = first; i < predicates.size(); ++i) {
         if (!predicates[i]->match(row_id)) {
            return false;
         }
      }
      return true;
   };
   CowBitmap candidates;
   if (child_operator.has_value()) {
      CowBitmap child_bitmap = (*child_operator)->evaluate();
      if (child_bitmap.cardinality() <= row_layout.numRows() / 10) {
         Roaring result;
         for (size_t idx = 0; idx < child_bitmap.size(); ++idx) {
            const uint32_t base = static_cast<uint32_t>(child_bitmap.keyAt(idx)) << 16;
            child_bitmap.containerAt(idx).forEach([&](uint16_t row_in_chunk) {
               if (matchesAll(0, base | row_in_chunk)) {
                  result.add(base | row_in_chunk);
               }
            });
         }
         return CowBitmap{std::move(result)};
      }
      candidates = std::move(child_bitmap);
      candidates &= CowBitmap{predicates.front()->makeBitmap(row_layout)};
   } else {
      candidates = CowBitmap{predicates.front()->makeBitmap(row_layout)};
   }
   if (predicates.size() == 1) {
      return candidates;
   }
   Roaring result;
   for (size_t idx = 0; idx < candidates.size(); ++idx) {
      const uint32_t base = static_cast<uint32_t>(candidates.keyAt(idx)) << 16;
      candidates.containerAt(idx).forEach([&](uint16_t row_in_chunk) {
         if (matchesAll(1, base | row_in_chunk)) {
            result.add(base | row_in_chunk);
         }
      });
   }
   return CowBitmap{std::move(result)};
}

std::string Selection::toString() const {
   std::string res = "Select[";
   for (const auto& predicate : predicates) {
      res += predicate->toString() + ",";
   }
   res += "](";
   if (child_operator.has_value()) {
      res += child_operator.value()->toString();
   }
   return res + ")";
}

}  // namespace oracle
