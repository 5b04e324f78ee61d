// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the reference's BitmapAggregationNode
// (/root/reference/src/rhydb/query_engine/operators/bitmap_aggregation_node.{h,cpp}): per-dimension
// group bitmaps restricted to the filter, then the recursive depth-first partition.
#pragma once
#include <optional>
#include <string>
#include <utility>
#include <vector>

#include "cow_bitmap.h"
#include "expressions.h"
#include "storage.h"

namespace oracle {

// bitmap_aggregation_node.h: GroupBitmaps = one (value or null, bitmap) per group of a dimension
using GroupBitmaps = std::vector<std::pair<std::optional<std::string>, CowBitmap>>;

// SequencePositionDimension (.cpp:177-205) or IndexedColumnDimension (.cpp:217-249); the inverted
// index of the (out-of-scope) dictionary-encoded column arrives as named bitmaps of the table
struct GroupingDimension {
   bool is_sequence_position = true;
   std::string column;         // sequence position
   uint32_t position_idx = 0;  // sequence position, 0-based
   std::vector<std::pair<std::string, std::string>> value_bitmaps;  // indexed column: (value, bitmap name)
   std::string null_bitmap;                                         // indexed column: name, or empty
};

struct Combination {
   std::vector<std::optional<std::string>> values;  // one per dimension
   uint64_t count = 0;
};

// addToExecPlan (.cpp:304-356) up to the materialised combinations (buildBatch only formats them)
std::vector<Combination> bitmapAggregation(
   const Table& table,
   const Expression& filter,
   const std::vector<GroupingDimension>& dimensions
);

}  // namespace oracle
