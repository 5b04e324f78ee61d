#include "mutations_node.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace silo_host {

SymbolCounts calculateMutationsPerPosition(
   const Table& table,
   const SequenceColumnInfo& sequence_column,
   const DeviceBitmap& bitmap_filter,
   uint64_t sequence_count_in_column,
   bool valid_mutation_symbols_only
) {
   uint64_t symbol_mask = ~0ULL;
   if (valid_mutation_symbols_only) {
      symbol_mask = 0;
      for (const Symbol symbol : sequence_column.alphabet->valid_mutation_symbols) {
         symbol_mask |= 1ULL << symbol;
      }
   }
   SymbolCounts counts;
   counts.n_symbols = sequence_column.alphabet->count();
   counts.genome_length = static_cast<uint32_t>(sequence_column.reference_sequence.size());
   counts.owner = table.acquireCountsBuffer(counts.size());
   counts.values = counts.owner.get();
   const uint64_t filter_cardinality = bitmap_filter.cardinality();
   if (filter_cardinality == sequence_count_in_column) {
      // addMutationCountsForFullBitmaps (:239-266): stored cardinalities only, no intersections
      throwOnDeviceError(silo_gpu_mutation_counts_symbols(
         table.deviceTable(), sequence_column.device_column, nullptr, symbol_mask, counts.owner.get()
      ));
   } else if (filter_cardinality > 0) {
      // addMutationCountsForMixedBitmaps (:205-237)
      throwOnDeviceError(silo_gpu_mutation_counts_symbols(
         table.deviceTable(), sequence_column.device_column, bitmap_filter.get(), symbol_mask, counts.owner.get()
      ));
   } else {
      std::memset(counts.owner.get(), 0, counts.size() * sizeof(uint32_t));
   }
   return counts;
}

void appendMutationRows(
   const SequenceColumnInfo& sequence_column,
   const SymbolCounts& counts,
   double min_proportion,
   std::vector<MutationRow>& out
) {
   const Alphabet& alphabet = *sequence_column.alphabet;
   const uint32_t genome_length = counts.genome_length;
   // Same arithmetic as the reference, position by position; only the memory order differs: the
   // totals are summed symbol row by symbol row (sequential streams over the freshly DMA-written
   // buffer, auto-vectorised) before the per-position pass.
   thread_local std::vector<uint32_t> scratch;  // (raw pointers below: no TLS lookup inside the loops)
   scratch.assign(2ULL * genome_length, 0);
   uint32_t* __restrict__ const totals = scratch.data();
   uint32_t* __restrict__ const any_other = scratch.data() + genome_length;  // OR of the valid non-reference symbols' counts
   const Symbol* __restrict__ const reference_sequence = sequence_column.reference_sequence.data();
   for (const Symbol symbol : alphabet.valid_mutation_symbols) {
      const uint32_t* __restrict__ const row = counts.values + static_cast<size_t>(symbol) * genome_length;
      for (uint32_t pos = 0; pos < genome_length; ++pos) {
         const uint32_t value = row[pos];
         const uint32_t keep = static_cast<uint32_t>(reference_sequence[pos] == symbol) - 1u;  // 0 for the reference symbol
         totals[pos] += value;
         any_other[pos] |= value & keep;
      }
   }
   for (uint32_t pos = 0; pos < genome_length; ++pos) {
      const uint32_t total = totals[pos];
      // no valid non-reference symbol was counted here (the common case): `count > threshold_count`
      // cannot hold for a zero count, so no row is emitted
      if (total == 0 || any_other[pos] == 0) {
         continue;
      }
      const uint32_t threshold_count =
         min_proportion == 0 ? 0 : static_cast<uint32_t>(std::ceil(static_cast<double>(total) * min_proportion) - 1);
      const Symbol symbol_in_reference_genome = reference_sequence[pos];
      for (const Symbol symbol : alphabet.valid_mutation_symbols) {
         if (symbol == symbol_in_reference_genome) {
            continue;
         }
         const uint32_t count = counts.at(symbol, pos);
         if (count > threshold_count) {
            MutationRow row;
            row.mutation_from = alphabet.symbolToChar(symbol_in_reference_genome);
            row.mutation_to = alphabet.symbolToChar(symbol);
            row.position = static_cast<int32_t>(pos + 1);
            row.sequence_name = sequence_column.name;
            row.proportion = static_cast<double>(count) / static_cast<double>(total);
            row.count = static_cast<int32_t>(count);
            row.coverage = static_cast<int32_t>(total);
            out.push_back(std::move(row));
         }
      }
   }
}

namespace {

uint64_t validSymbolMask(const Alphabet& alphabet) {
   uint64_t mask = 0;
   for (const Symbol symbol : alphabet.valid_mutation_symbols) {
      mask |= 1ULL << symbol;
   }
   return mask;
}

// rows from the tuples of the device's output pass (silo_gpu_query_mutation_hits): the same fields
// addMutationsToOutput emits (:326-360), proportion computed here as double(count) / double(total)
void appendRowsFromHits(
   const SequenceColumnInfo& sequence_column,
   const silo_mutation_hit* hits,
   uint64_t n_hits,
   std::vector<MutationRow>& out
) {
   const Alphabet& alphabet = *sequence_column.alphabet;
   std::vector<silo_mutation_hit> reordered;
   if (!std::is_sorted(alphabet.valid_mutation_symbols.begin(), alphabet.valid_mutation_symbols.end())) {
      // the device orders by (position, symbol id); the reference walks VALID_MUTATION_SYMBOLS in its own order
      std::vector<uint32_t> rank(alphabet.count(), 0);
      for (size_t i = 0; i < alphabet.valid_mutation_symbols.size(); ++i) {
         rank[alphabet.valid_mutation_symbols[i]] = static_cast<uint32_t>(i);
      }
      reordered.assign(hits, hits + n_hits);
      std::stable_sort(reordered.begin(), reordered.end(), [&](const silo_mutation_hit& a, const silo_mutation_hit& b) {
         return a.position != b.position ? a.position < b.position : rank[a.symbol] < rank[b.symbol];
      });
      hits = reordered.data();
   }
   out.reserve(out.size() + n_hits);
   for (uint64_t i = 0; i < n_hits; ++i) {
      const silo_mutation_hit& hit = hits[i];
      MutationRow row;
      row.mutation_from = alphabet.symbolToChar(sequence_column.reference_sequence[hit.position]);
      row.mutation_to = alphabet.symbolToChar(static_cast<Symbol>(hit.symbol));
      row.position = static_cast<int32_t>(hit.position + 1);
      row.sequence_name = sequence_column.name;
      row.proportion = static_cast<double>(hit.count) / static_cast<double>(hit.total);
      row.count = static_cast<int32_t>(hit.count);
      row.coverage = static_cast<int32_t>(hit.total);
      out.push_back(std::move(row));
   }
}

}  // namespace

std::vector<MutationRow> MutationsNode::execute() const {
   lastQueryProfile().counts_us = 0;
   lastQueryProfile().threshold_us = 0;
   std::vector<MutationRow> rows;
   if (sequence_columns.size() == 1) {
      // One sequence column (the common query): filter program, counts and the output pass go to the
      // device in ONE call with one synchronisation (silo_gpu_query_mutation_hits); neither the filter
      // nor the counts leave the device, only the emitted (position, symbol, count, total) tuples.
      const SequenceColumnInfo* column = table.findColumn(sequence_columns.front());
      if (column == nullptr) {
         throw IllegalQueryException("Database does not contain the Sequence with name: '" + sequence_columns.front() + "'");
      }
      const double compile_begin = nowMicroseconds();
      const ExpressionPtr rewritten = filter->rewrite(table, AmbiguityMode::NONE);  // computeFilter, compute_filter.cpp:14-21
      const std::unique_ptr<Operator> compiled = rewritten->compile(table);
      ProgramBuilder builder;
      const silo_filter_program program = compiled->lowerProgram(table, builder);
      const double counts_begin = nowMicroseconds();
      const silo_mutation_hit* hits = nullptr;
      uint64_t n_hits = 0;
      uint64_t cardinality = 0;
      throwOnDeviceError(silo_gpu_query_mutation_hits(
         table.deviceTable(), &program, nullptr, column->device_column, validSymbolMask(*column->alphabet), min_proportion, &hits, &n_hits,
         &cardinality
      ));
      const double threshold_begin = nowMicroseconds();
      appendRowsFromHits(*column, hits, n_hits, rows);
      QueryProfile& profile = lastQueryProfile();
      profile.compile_us = counts_begin - compile_begin;
      profile.filter_us = 0;
      profile.counts_us = threshold_begin - counts_begin;
      profile.threshold_us = nowMicroseconds() - threshold_begin;
      return rows;
   }
   // Several sequence columns (AminoAcidMutations over all genes): the filter program is evaluated ONCE and the columns
   // follow on the device in one call with one synchronisation (silo_gpu_query_mutation_hits_columns).
   std::vector<const SequenceColumnInfo*> columns;
   for (const std::string& name : sequence_columns) {
      const SequenceColumnInfo* column = table.findColumn(name);
      if (column == nullptr) {
         throw IllegalQueryException("Database does not contain the Sequence with name: '" + name + "'");
      }
      columns.push_back(column);
   }
   if (columns.empty()) {
      return rows;
   }
   const double compile_begin = nowMicroseconds();
   const ExpressionPtr rewritten = filter->rewrite(table, AmbiguityMode::NONE);  // computeFilter, compute_filter.cpp:14-21
   const std::unique_ptr<Operator> compiled = rewritten->compile(table);
   ProgramBuilder builder;
   const silo_filter_program program = compiled->lowerProgram(table, builder);
   const double counts_begin = nowMicroseconds();
   std::vector<silo_column_hits> requests(columns.size());
   for (size_t c = 0; c < columns.size(); ++c) {
      requests[c].column = columns[c]->device_column;
      requests[c].valid_symbol_mask = validSymbolMask(*columns[c]->alphabet);
   }
   uint64_t cardinality = 0;
   throwOnDeviceError(silo_gpu_query_mutation_hits_columns(
      table.deviceTable(), &program, requests.data(), static_cast<uint32_t>(requests.size()), min_proportion, &cardinality
   ));
   const double threshold_begin = nowMicroseconds();
   for (size_t c = 0; c < columns.size(); ++c) {
      appendRowsFromHits(*columns[c], requests[c].hits, requests[c].n_hits, rows);
   }
   QueryProfile& profile = lastQueryProfile();
   profile.compile_us = counts_begin - compile_begin;
   profile.filter_us = 0;
   profile.counts_us = threshold_begin - counts_begin;
   profile.threshold_us = nowMicroseconds() - threshold_begin;
   return rows;
}

namespace {
const SequenceColumnInfo& singleColumn(const MutationsNode& node) {
   if (node.sequence_columns.size() != 1) {
      throw IllegalQueryException("the sharded Mutations query takes exactly one sequence column");
   }
   const SequenceColumnInfo* column = node.table.findColumn(node.sequence_columns.front());
   if (column == nullptr) {
      throw IllegalQueryException("Database does not contain the Sequence with name: '" + node.sequence_columns.front() + "'");
   }
   return *column;
}
}  // namespace

void MutationsNode::enqueueShardCounts(void* d_counts, void* cuda_stream) const {
   const SequenceColumnInfo& column = singleColumn(*this);
   const ExpressionPtr rewritten = filter->rewrite(table, AmbiguityMode::NONE);  // computeFilter, compute_filter.cpp:14-21
   const std::unique_ptr<Operator> compiled = rewritten->compile(table);
   ProgramBuilder builder;
   const silo_filter_program program = compiled->lowerProgram(table, builder);
   throwOnDeviceError(silo_gpu_query_mutation_counts_async(table.deviceTable(), &program, column.device_column, d_counts, cuda_stream));
}

std::vector<MutationRow> MutationsNode::collectRows(const void* d_summed_counts, void* cuda_stream, uint64_t* shard_cardinality) const {
   const SequenceColumnInfo& column = singleColumn(*this);
   const silo_mutation_hit* hits = nullptr;
   uint64_t n_hits = 0;
   throwOnDeviceError(silo_gpu_mutation_hits_from_counts(
      table.deviceTable(), column.device_column, d_summed_counts, validSymbolMask(*column.alphabet), min_proportion, cuda_stream, &hits, &n_hits,
      shard_cardinality
   ));
   std::vector<MutationRow> rows;
   appendRowsFromHits(column, hits, n_hits, rows);
   return rows;
}

void MutationsNode::enqueueSharded(void* cuda_stream) const {
   const ExpressionPtr rewritten = filter->rewrite(table, AmbiguityMode::NONE);  // computeFilter, compute_filter.cpp:14-21
   const std::unique_ptr<Operator> compiled = rewritten->compile(table);
   ProgramBuilder builder;
   const silo_filter_program program = compiled->lowerProgram(table, builder);
   throwOnDeviceError(silo_gpu_sharded_query_enqueue(table.deviceTable(), &program, cuda_stream));
}

std::vector<MutationRow> MutationsNode::executeShardedRoot(void* d_summed_counts, uint64_t* cardinality) const {
   const SequenceColumnInfo& column = singleColumn(*this);
   const ExpressionPtr rewritten = filter->rewrite(table, AmbiguityMode::NONE);
   const std::unique_ptr<Operator> compiled = rewritten->compile(table);
   ProgramBuilder builder;
   const silo_filter_program program = compiled->lowerProgram(table, builder);
   const silo_mutation_hit* hits = nullptr;
   uint64_t n_hits = 0;
   throwOnDeviceError(silo_gpu_sharded_query_hits(table.deviceTable(), &program, min_proportion, d_summed_counts, &hits, &n_hits, cardinality));
   std::vector<MutationRow> rows;
   appendRowsFromHits(column, hits, n_hits, rows);
   return rows;
}

std::vector<MutationRow> MutationsNode::collectSharded(void* d_summed_counts, void* cuda_stream, uint64_t* cardinality) const {
   const SequenceColumnInfo& column = singleColumn(*this);
   const silo_mutation_hit* hits = nullptr;
   uint64_t n_hits = 0;
   throwOnDeviceError(silo_gpu_sharded_collect(table.deviceTable(), min_proportion, d_summed_counts, cuda_stream, &hits, &n_hits, cardinality));
   std::vector<MutationRow> rows;
   appendRowsFromHits(column, hits, n_hits, rows);
   return rows;
}

std::vector<uint8_t> createShardGroup(const Table& table, const std::string& sequence_column, int rank, int world) {
   const SequenceColumnInfo* found = table.findColumn(sequence_column);
   if (found == nullptr) {
      throw IllegalQueryException("Database does not contain the Sequence with name: '" + sequence_column + "'");
   }
   const SequenceColumnInfo& column = *found;
   std::vector<uint8_t> handle(SILO_SHARD_HANDLE_BYTES, 0);
   throwOnDeviceError(silo_gpu_shard_group_init(table.deviceTable(), column.device_column, validSymbolMask(*column.alphabet), rank, world, handle.data()));
   return handle;
}

void connectShardGroup(const Table& table, const std::vector<uint8_t>& handles_of_all_ranks) {
   throwOnDeviceError(silo_gpu_shard_group_connect(table.deviceTable(), handles_of_all_ranks.data()));
}

uint64_t countFilter(const Table& table, const ScalarExpression& filter) {
   // computeFilter (compute_filter.cpp:14-21) + cardinality (count_filter_node.cpp:40-41) as one device call
   const ExpressionPtr rewritten = filter.rewrite(table, AmbiguityMode::NONE);
   const std::unique_ptr<Operator> compiled = rewritten->compile(table);
   ProgramBuilder builder;
   const silo_filter_program program = compiled->lowerProgram(table, builder);
   uint64_t cardinality = 0;
   throwOnDeviceError(silo_gpu_query_count(table.deviceTable(), &program, &cardinality));
   return cardinality;
}

}  // namespace silo_host
