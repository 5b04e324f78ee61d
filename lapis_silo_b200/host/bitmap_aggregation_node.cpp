#include "bitmap_aggregation_node.h"

#include <algorithm>

#include "operators.h"

namespace silo_host {

namespace {

constexpr uint32_t POSITION_CODE_BITS = 5;  // include/silo_b200.h: key layout of silo_gpu_query_combinations
constexpr uint32_t BITMAP_CODE_BITS = 8;

// a named bitmap that is not device resident yet is registered for the duration of the query
struct TemporaryRegistrations {
   silo_gpu_table* device;
   std::vector<uint32_t> ids;
   ~TemporaryRegistrations() {
      for (uint32_t id : ids) {
         silo_gpu_bitmap_unregister(device, id);
      }
   }
};

}  // namespace

namespace {

struct DimensionDecoder {
   uint32_t bits = 0;
   const Alphabet* alphabet = nullptr;   // sequence position
   std::vector<std::string> values;      // indexed column, in output order
};

// value groups are emitted in sorted order (.cpp:241-243), the null group last
std::vector<std::pair<std::string, std::string>> sortedValueGroups(const IndexedColumnDimension& indexed) {
   std::vector<std::pair<std::string, std::string>> sorted = indexed.value_bitmaps;
   std::sort(sorted.begin(), sorted.end(), [](const auto& lhs, const auto& rhs) { return lhs.first < rhs.first; });
   return sorted;
}

std::vector<DimensionDecoder> decodersOf(const Table& table, const std::vector<GroupingDimension>& dimensions) {
   std::vector<DimensionDecoder> decoders(dimensions.size());
   for (size_t d = 0; d < dimensions.size(); ++d) {
      if (const auto* position = std::get_if<SequencePositionDimension>(&dimensions[d])) {
         const SequenceColumnInfo* column = table.findColumn(position->column);
         if (column == nullptr) {
            throw IllegalQueryException("Database does not contain the Sequence with name: '" + position->column + "'");
         }
         decoders[d].bits = POSITION_CODE_BITS;
         decoders[d].alphabet = column->alphabet;
      } else {
         for (const auto& [value, bitmap_name] : sortedValueGroups(std::get<IndexedColumnDimension>(dimensions[d]))) {
            decoders[d].values.push_back(value);
         }
         decoders[d].bits = BITMAP_CODE_BITS;
      }
   }
   return decoders;
}

}  // namespace

std::vector<CombinationRow> BitmapAggregationNode::execute() const {
   return materialise(executeShard());
}

BitmapAggregationNode::ShardCombinations BitmapAggregationNode::mergeShards(const std::vector<ShardCombinations>& shards) {
   ShardCombinations merged;
   size_t total = 0;
   for (const ShardCombinations& shard : shards) {
      merged.cardinality += shard.cardinality;
      total += shard.entries.size();
   }
   merged.entries.reserve(total);
   for (const ShardCombinations& shard : shards) {
      merged.entries.insert(merged.entries.end(), shard.entries.begin(), shard.entries.end());
   }
   std::stable_sort(merged.entries.begin(), merged.entries.end(), [](const silo_combination& lhs, const silo_combination& rhs) { return lhs.key < rhs.key; });
   size_t out = 0;
   for (size_t i = 0; i < merged.entries.size(); ++i) {
      if (out > 0 && merged.entries[out - 1].key == merged.entries[i].key) {
         merged.entries[out - 1].count += merged.entries[i].count;
      } else {
         merged.entries[out++] = merged.entries[i];
      }
   }
   merged.entries.resize(out);
   return merged;
}

std::vector<CombinationRow> BitmapAggregationNode::materialise(const ShardCombinations& combinations) const {
   // buildBatch (.cpp:146-160): one value (or null) per dimension and the count
   std::vector<CombinationRow> rows;
   if (dimensions.empty()) {
      // partition() at depth 0 == dimensions.size(): one combination holding the filter's cardinality
      rows.push_back(CombinationRow{{}, static_cast<int64_t>(combinations.cardinality)});
      return rows;
   }
   const std::vector<DimensionDecoder> decoders = decodersOf(table, dimensions);
   rows.reserve(combinations.entries.size());
   for (const silo_combination& entry : combinations.entries) {
      CombinationRow row;
      row.values.resize(dimensions.size());
      uint64_t key = entry.key;
      for (size_t d = dimensions.size(); d-- > 0;) {  // dimension 0 sits in the most significant bits
         const uint32_t code = static_cast<uint32_t>(key & ((1ULL << decoders[d].bits) - 1));
         key >>= decoders[d].bits;
         if (decoders[d].alphabet != nullptr) {
            if (code < decoders[d].alphabet->count()) {
               row.values[d] = std::string(1, decoders[d].alphabet->symbolToChar(static_cast<Symbol>(code)));
            }
         } else if (code < decoders[d].values.size()) {
            row.values[d] = decoders[d].values[code];
         }
      }
      row.count = static_cast<int64_t>(entry.count);
      rows.push_back(std::move(row));
   }
   return rows;
}

BitmapAggregationNode::ShardCombinations BitmapAggregationNode::executeShard() const {
   const ExpressionPtr rewritten = filter->rewrite(table, AmbiguityMode::NONE);  // computeFilter, compute_filter.cpp:14-21
   const std::unique_ptr<Operator> compiled = rewritten->compile(table);
   ProgramBuilder builder;
   const silo_filter_program program = compiled->lowerProgram(table, builder);

   TemporaryRegistrations temporary{table.deviceTable(), {}};
   auto deviceId = [&](const std::string& name) -> uint32_t {
      const auto found = table.named_bitmaps.find(name);
      if (found == table.named_bitmaps.end()) {
         throw IllegalQueryException("unknown bitmap " + name);
      }
      if (found->second.resident) {
         return found->second.device_id;
      }
      uint32_t id = 0;
      throwOnDeviceError(silo_gpu_bitmap_register(table.deviceTable(), found->second.bytes.data(), found->second.bytes.size(), &id));
      temporary.ids.push_back(id);
      return id;
   };

   std::vector<silo_group_dimension> device_dimensions(dimensions.size());
   std::vector<std::vector<uint32_t>> group_ids(dimensions.size());
   for (size_t d = 0; d < dimensions.size(); ++d) {
      silo_group_dimension& out = device_dimensions[d];
      out = silo_group_dimension{};
      if (const auto* position = std::get_if<SequencePositionDimension>(&dimensions[d])) {
         const SequenceColumnInfo* column = table.findColumn(position->column);
         if (column == nullptr) {
            throw IllegalQueryException("Database does not contain the Sequence with name: '" + position->column + "'");
         }
         // CHECK_SILO_QUERY of compileSymbolInSet, symbol_in_set.cpp:238-244
         if (position->position_idx >= column->reference_sequence.size()) {
            throw IllegalQueryException(
               "SymbolInSet<" + column->alphabet->symbol_name + "> position is out of bounds " + std::to_string(position->position_idx + 1) +
               " > " + std::to_string(column->reference_sequence.size())
            );
         }
         out.kind = SILO_DIM_SEQUENCE_POSITION;
         out.column = column->device_column;
         out.position = position->position_idx;
      } else {
         const auto& indexed = std::get<IndexedColumnDimension>(dimensions[d]);
         for (const auto& [value, bitmap_name] : sortedValueGroups(indexed)) {
            group_ids[d].push_back(deviceId(bitmap_name));
         }
         out.kind = SILO_DIM_INDEX_BITMAPS;
         out.n_groups = static_cast<uint32_t>(group_ids[d].size());
         out.bitmap_ids = group_ids[d].data();
         out.null_bitmap_id = indexed.null_bitmap.has_value() ? deviceId(indexed.null_bitmap.value()) : UINT32_MAX;
      }
   }

   const silo_combination* combinations = nullptr;
   uint64_t n_combinations = 0;
   uint64_t cardinality = 0;
   throwOnDeviceError(silo_gpu_query_combinations(
      table.deviceTable(), &program, nullptr, device_dimensions.data(), static_cast<uint32_t>(device_dimensions.size()), &combinations,
      &n_combinations, &cardinality
   ));

   ShardCombinations result;
   result.entries.assign(combinations, combinations + n_combinations);
   result.cardinality = cardinality;
   return result;
}

}  // namespace silo_host
