// ORACLE — TEST INFRASTRUCTURE ONLY. See aggregation.h.
#include "aggregation.h"

#include <algorithm>

#include "operators.h"

namespace oracle {

namespace {

struct GroupCombination {  // bitmap_aggregation_node.cpp:38-41
   std::vector<size_t> group_indices;
   uint64_t count;
};

// bitmap_aggregation_node.cpp:53-92
GroupBitmaps buildSymbolBitmaps(
   const SequenceColumn& column,
   uint32_t position_idx,
   const RowLayout& row_layout,
   const CowBitmap& filter_bitmap
) {
   GroupBitmaps result;
   for (Symbol symbol = 0; symbol < column.alphabet->count; ++symbol) {  // SymbolType::SYMBOLS, in enum order
      const auto compiled = compileSymbolInSet(column, position_idx, std::vector<Symbol>{symbol}, row_layout);
      CowBitmap bitmap = compiled->evaluate();
      bitmap &= filter_bitmap;
      if (!bitmap.isEmpty()) {
         result.emplace_back(std::string(1, column.alphabet->symbolToChar(symbol)), CowBitmap{bitmap.toRoaring()});
      }
   }
   CowBitmap intersection = filter_bitmap & CowBitmap{&column.null_bitmap};
   if (!intersection.isEmpty()) {
      result.emplace_back(std::nullopt, CowBitmap{intersection.toRoaring()});
   }
   return result;
}

// bitmap_aggregation_node.cpp:224-249
GroupBitmaps buildIndexedGroups(const Table& table, const GroupingDimension& dimension, const CowBitmap& filter_bitmap) {
   GroupBitmaps result;
   for (const auto& [value, bitmap_name] : dimension.value_bitmaps) {
      const auto found = table.named_bitmaps.find(bitmap_name);
      if (found == table.named_bitmaps.end()) {
         throw IllegalQueryException("unknown bitmap " + bitmap_name);
      }
      CowBitmap group = filter_bitmap & CowBitmap{&found->second};
      if (group.isEmpty()) {
         continue;
      }
      result.emplace_back(value, std::move(group));
   }
   std::sort(result.begin(), result.end(), [](const auto& lhs, const auto& rhs) { return lhs.first < rhs.first; });
   if (!dimension.null_bitmap.empty()) {
      const auto found = table.named_bitmaps.find(dimension.null_bitmap);
      if (found == table.named_bitmaps.end()) {
         throw IllegalQueryException("unknown bitmap " + dimension.null_bitmap);
      }
      CowBitmap null_group = CowBitmap{&found->second} & filter_bitmap;
      if (!null_group.isEmpty()) {
         result.emplace_back(std::nullopt, std::move(null_group));
      }
   }
   return result;
}

// bitmap_aggregation_node.cpp:99-124
void partition(
   const CowBitmap& current,
   size_t depth,
   const std::vector<GroupBitmaps>& group_bitmaps_per_dimension,
   std::vector<size_t>& accumulated_indices,
   std::vector<GroupCombination>& combinations
) {
   if (depth == group_bitmaps_per_dimension.size()) {
      combinations.push_back(GroupCombination{accumulated_indices, current.cardinality()});
      return;
   }
   const auto& dimension = group_bitmaps_per_dimension[depth];
   for (size_t group_index = 0; group_index < dimension.size(); ++group_index) {
      CowBitmap intersection = current & dimension[group_index].second;
      if (intersection.isEmpty()) {
         continue;
      }
      accumulated_indices.push_back(group_index);
      partition(intersection, depth + 1, group_bitmaps_per_dimension, accumulated_indices, combinations);
      accumulated_indices.pop_back();
   }
}

}  // namespace

std::vector<Combination> bitmapAggregation(
   const Table& table,
   const Expression& filter,
   const std::vector<GroupingDimension>& dimensions
) {
   const CowBitmap filter_bitmap = computeFilter(filter, table);
   std::vector<GroupBitmaps> group_bitmaps_per_dimension;
   group_bitmaps_per_dimension.reserve(dimensions.size());
   for (const auto& dimension : dimensions) {
      if (dimension.is_sequence_position) {
         const SequenceColumn* column = table.findColumn(dimension.column);
         if (column == nullptr) {
            throw IllegalQueryException("Database does not contain the Sequence with name: '" + dimension.column + "'");
         }
         group_bitmaps_per_dimension.push_back(buildSymbolBitmaps(*column, dimension.position_idx, table.row_layout, filter_bitmap));
      } else {
         group_bitmaps_per_dimension.push_back(buildIndexedGroups(table, dimension, filter_bitmap));
      }
   }
   std::vector<GroupCombination> combinations;  // computeCombinations, .cpp:130-139
   std::vector<size_t> accumulated_indices;
   partition(filter_bitmap, 0, group_bitmaps_per_dimension, accumulated_indices, combinations);

   std::vector<Combination> result;  // buildBatch, .cpp:146-160
   result.reserve(combinations.size());
   for (const auto& combination : combinations) {
      Combination row;
      for (size_t i = 0; i < dimensions.size(); ++i) {
         row.values.push_back(group_bitmaps_per_dimension[i][combination.group_indices[i]].first);
      }
      row.count = combination.count;
      result.push_back(std::move(row));
   }
   return result;
}

}  // namespace oracle
