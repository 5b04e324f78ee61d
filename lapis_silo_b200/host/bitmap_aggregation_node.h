// BitmapAggregationNode of the host layer: mutation co-occurrence and groupBy over sequence
// positions / indexed columns.
//
// Same interface as /root/reference/src/rhydb/query_engine/operators/bitmap_aggregation_node.{h,cpp}:
//   SequencePositionDimension (.cpp:177-205), IndexedColumnDimension (.cpp:217-249),
//   BitmapAggregationNode::addToExecPlan (.cpp:304-356)
// but buildGroups + computeCombinations (the |alphabet| SymbolInSet bitmaps per position and the
// recursive partition) are ONE device call, silo_gpu_query_combinations; this file keeps what the
// reference does on the host before and after: computeFilter's rewrite/compile, the position bound
// check of compileSymbolInSet (symbol_in_set.cpp:238-244), the sorting of an indexed column's value
// groups, and the materialisation of the combinations (buildBatch, .cpp:146-160).
#pragma once
#include <optional>
#include <string>
#include <utility>
#include <variant>
#include <vector>

#include "expressions.h"
#include "table.h"

namespace silo_host {

struct SequencePositionDimension {
   std::string column;
   uint32_t position_idx = 0;  // 0-based
   std::string output_name;
};

// The inverted index of a DictionaryEncodedColumn (out of scope here) arrives as named bitmaps of the
// table: one per dictionary value, plus the column's null bitmap.
struct IndexedColumnDimension {
   std::vector<std::pair<std::string, std::string>> value_bitmaps;  // (value, bitmap name)
   std::optional<std::string> null_bitmap;
   std::string output_name;
};

using GroupingDimension = std::variant<SequencePositionDimension, IndexedColumnDimension>;

struct CombinationRow {
   std::vector<std::optional<std::string>> values;  // one per dimension; nullopt = the null group
   int64_t count = 0;
};

class BitmapAggregationNode {
  public:
   const Table& table;
   ExpressionPtr filter;
   std::vector<GroupingDimension> dimensions;

   BitmapAggregationNode(const Table& table, ExpressionPtr filter, std::vector<GroupingDimension> dimensions)
       : table(table),
         filter(std::move(filter)),
         dimensions(std::move(dimensions)) {}

   // the combinations in the reference's depth-first output order
   [[nodiscard]] std::vector<CombinationRow> execute() const;

   // Row-partitioned tables (BASELINE.json configs[4]): every rank runs executeShard() on the table of
   // its shard -- the same dimensions everywhere, so that the keys mean the same --, the (key, count) lists travel to
   // one rank by whatever transport the ranks share (they are a few thousand entries), and that rank merges them and
   // materialises the rows. Counts of disjoint row sets are plain addends: merge == sum per key.
   struct ShardCombinations {
      std::vector<silo_combination> entries;  // ordered by key, count > 0
      uint64_t cardinality = 0;               // rows of the shard under the filter
   };
   [[nodiscard]] ShardCombinations executeShard() const;
   [[nodiscard]] static ShardCombinations mergeShards(const std::vector<ShardCombinations>& shards);
   // keys -> one value (or null) per dimension, in key order == the reference's output order
   [[nodiscard]] std::vector<CombinationRow> materialise(const ShardCombinations& combinations) const;
};

}  // namespace silo_host
