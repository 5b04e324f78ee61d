#include "mutations_node.h"

#include <cmath>
#include <cstring>

namespace silo_host {

SymbolCounts calculateMutationsPerPosition(
   const Table& table,
   const SequenceColumnInfo& sequence_column,
   const DeviceBitmap& bitmap_filter,
   uint64_t sequence_count_in_column
) {
   SymbolCounts counts;
   counts.n_symbols = sequence_column.alphabet->count();
   counts.genome_length = static_cast<uint32_t>(sequence_column.reference_sequence.size());
   counts.owner = table.acquireCountsBuffer(counts.size());
   counts.values = counts.owner.get();
   const uint64_t filter_cardinality = bitmap_filter.cardinality();
   if (filter_cardinality == sequence_count_in_column) {
      // addMutationCountsForFullBitmaps (:239-266): stored cardinalities only, no intersections
      throwOnDeviceError(
         silo_gpu_mutation_counts(table.device, sequence_column.device_column, nullptr, counts.owner.get())
      );
   } else if (filter_cardinality > 0) {
      // addMutationCountsForMixedBitmaps (:205-237)
      throwOnDeviceError(silo_gpu_mutation_counts(
         table.device, sequence_column.device_column, bitmap_filter.get(), counts.owner.get()
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
   for (uint32_t pos = 0; pos < counts.genome_length; ++pos) {
      uint32_t total = 0;
      for (const Symbol symbol : alphabet.valid_mutation_symbols) {
         total += counts.at(symbol, pos);
      }
      if (total == 0) {
         continue;
      }
      const uint32_t threshold_count =
         min_proportion == 0 ? 0 : static_cast<uint32_t>(std::ceil(static_cast<double>(total) * min_proportion) - 1);
      const Symbol symbol_in_reference_genome = sequence_column.reference_sequence.at(pos);
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

std::vector<MutationRow> MutationsNode::execute() const {
   const DeviceBitmap bitmap_filter = computeFilter(*filter, table);
   std::vector<MutationRow> rows;
   for (const std::string& name : sequence_columns) {
      const SequenceColumnInfo* column = table.findColumn(name);
      if (column == nullptr) {
         throw IllegalQueryException("Database does not contain the Sequence with name: '" + name + "'");
      }
      const SymbolCounts counts =
         calculateMutationsPerPosition(table, *column, bitmap_filter, table.row_layout.numRows());
      appendMutationRows(*column, counts, min_proportion, rows);
   }
   return rows;
}

uint64_t countFilter(const Table& table, const ScalarExpression& filter) {
   return computeFilter(filter, table).cardinality();
}

}  // namespace silo_host
