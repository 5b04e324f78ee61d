// The Mutations / AminoAcidMutations action and the count sink of the host layer.
//
// Same split as /root/reference/src/rhydb/query_engine/operators/mutations_node.cpp:
//   calculateMutationsPerPosition :268-288  -> one call into the device (silo_gpu_mutation_counts)
//   addMutationsToOutput :290-366           -> unchanged host arithmetic (uint32 totals, the
//        threshold  ceil(double(total) * minProportion) - 1  and the double proportion), so the
//        emitted rows are bit-identical
// and operators/count_filter_node.cpp:35-71 (count = filter cardinality).
#pragma once
#include <string>
#include <vector>

#include "expressions.h"
#include "operators.h"
#include "table.h"

namespace silo_host {

// SymbolMap<SymbolType, std::vector<uint32_t>>: counts[symbol * genome_length + position]
// A view: `values` points into a page-locked buffer of the table's pool (kept alive by `owner`) that
// the device wrote directly, or into caller memory (all-reduced counts of the multi-GPU scheduler).
struct SymbolCounts {
   uint32_t n_symbols = 0;
   uint32_t genome_length = 0;
   const uint32_t* values = nullptr;
   std::shared_ptr<uint32_t> owner;
   [[nodiscard]] size_t size() const { return static_cast<size_t>(n_symbols) * genome_length; }
   [[nodiscard]] uint32_t at(Symbol symbol, uint32_t position) const {
      return values[static_cast<size_t>(symbol) * genome_length + position];
   }
};

SymbolCounts calculateMutationsPerPosition(
   const Table& table,
   const SequenceColumnInfo& sequence_column,
   const DeviceBitmap& bitmap_filter,
   uint64_t sequence_count_in_column,
   // fetch only the rows of Alphabet::valid_mutation_symbols (all that addMutationsToOutput reads);
   // the other rows of the result are then unspecified
   bool valid_mutation_symbols_only = false
);

struct MutationRow {
   char mutation_from;
   char mutation_to;
   int32_t position;  // 1-based
   std::string sequence_name;
   double proportion;
   int32_t count;
   int32_t coverage;
};

// the per-position part of addMutationsToOutput (:307-363), also used on all-reduced counts
void appendMutationRows(
   const SequenceColumnInfo& sequence_column,
   const SymbolCounts& counts,
   double min_proportion,
   std::vector<MutationRow>& out
);

class MutationsNode {
  public:
   const Table& table;
   ExpressionPtr filter;
   std::vector<std::string> sequence_columns;
   double min_proportion;

   MutationsNode(const Table& table, ExpressionPtr filter, std::vector<std::string> sequence_columns, double min_proportion)
       : table(table),
         filter(std::move(filter)),
         sequence_columns(std::move(sequence_columns)),
         min_proportion(min_proportion) {}

   // addToExecPlan + producer (:372-428): the filter is evaluated once, then every column
   [[nodiscard]] std::vector<MutationRow> execute() const;

   // The same query on a row-partitioned table (one process per GPU), in two halves with the
   // scheduler's all-reduce of the counts in between. Every rank: compile the filter against its shard and
   // leave the shard's counts of the (single) sequence column in device memory, nothing synchronised.
   void enqueueShardCounts(void* d_counts, void* cuda_stream) const;
   // One rank, after the counts of all ranks were summed on the same stream: the output pass over the summed
   // counts; returns the rows and the number of this shard's rows that passed the filter.
   [[nodiscard]] std::vector<MutationRow> collectRows(const void* d_summed_counts, void* cuda_stream, uint64_t* shard_cardinality) const;

   // The same query through the shard group of the table (createShardGroup / connectShardGroup): every rank enqueues,
   // the ranks' rows travel to rank 0 inside the finalize kernels; rank 0 collects the rows of the whole table and
   // the number of rows of ALL shards that passed the filter.
   void enqueueSharded(void* cuda_stream) const;
   // rank 0's two halves in one device call (one replayed graph, one synchronisation; the table's own stream)
   [[nodiscard]] std::vector<MutationRow> executeShardedRoot(void* d_summed_counts, uint64_t* cardinality) const;
   [[nodiscard]] std::vector<MutationRow> collectSharded(void* d_summed_counts, void* cuda_stream, uint64_t* cardinality) const;
};

// The partition scheduler of a row-partitioned table: this process holds one shard. joinShardGroup is
// called once per table and column with every rank's handle exchanged in between (any transport); afterwards
// MutationsNode::enqueueSharded on every rank and ::collectSharded on rank 0 answer one Mutations query with no
// collective kernel and one host synchronisation, on rank 0 only (include/silo_b200.h, silo_gpu_shard_group_*).
std::vector<uint8_t> createShardGroup(const Table& table, const std::string& sequence_column, int rank, int world);
void connectShardGroup(const Table& table, const std::vector<uint8_t>& handles_of_all_ranks);

// CountFilterNode: `filter(...).groupBy({count:=count()})`
uint64_t countFilter(const Table& table, const ScalarExpression& filter);

}  // namespace silo_host
