// ORACLE — TEST INFRASTRUCTURE ONLY. See mutations.h for the reference file:line map.
#include "mutations.h"

#include <cmath>
#include <map>

namespace oracle {

namespace {

void initializeCountsWithSequenceCount(std::vector<uint32_t>& count_per_position, uint32_t sequence_count) {
   for (auto& count : count_per_position) {
      count += sequence_count;
   }
}

void subtractCumulativeNsFromPositions(
   std::vector<uint32_t>& count_per_position,
   uint32_t sequence_length,
   const std::vector<size_t>& cumulative_starts,
   const std::vector<size_t>& cumulative_ends
) {
   size_t running_total_start_n_offset = cumulative_starts.at(sequence_length);
   size_t start_position_iter = sequence_length - 1;
   while (true) {
      count_per_position.at(start_position_iter) -= static_cast<uint32_t>(running_total_start_n_offset);
      running_total_start_n_offset += cumulative_starts.at(start_position_iter);
      if (start_position_iter == 0) {
         break;
      }
      start_position_iter -= 1;
   }
   size_t running_total_end_n_offset = cumulative_ends.at(0);
   size_t end_position_iter = 0;
   while (true) {
      count_per_position.at(end_position_iter) -= static_cast<uint32_t>(running_total_end_n_offset);
      running_total_end_n_offset += cumulative_ends.at(end_position_iter + 1);
      if (end_position_iter == sequence_length - 1) {
         break;
      }
      end_position_iter += 1;
   }
}

void addMutationCountsForMixedBitmaps(
   const SequenceColumn& sequence_column,
   const CowBitmap& bitmap_filter,
   MutationCounts& counts
) {
   const auto local_reference = sequence_column.getLocalReference();
   const size_t sequence_length = local_reference.size();
   std::vector<uint32_t> count_per_local_reference_position(sequence_length);
   const Roaring filter_bitmap = bitmap_filter.toRoaring();

   initializeCountsWithSequenceCount(
      count_per_local_reference_position, static_cast<uint32_t>(filter_bitmap.cardinality())
   );
   {  // subtractFilteredNCounts :111-136
      const auto& coverage_index = sequence_column.horizontal_coverage_index;
      const auto& horizontal_bitmaps = coverage_index.horizontal_bitmaps;
      std::vector<size_t> cumulative_starts(sequence_length + 1);
      std::vector<size_t> cumulative_ends(sequence_length + 1);
      filter_bitmap.forEach([&](uint32_t idx) {
         auto iter = horizontal_bitmaps.find(idx);
         if (iter != horizontal_bitmaps.end()) {
            iter->second.forEach([&](uint32_t position_idx) {
               count_per_local_reference_position[position_idx] -= 1;
            });
         }
         auto [start, end] = coverage_index.coverageRange(idx);
         cumulative_starts.at(start) += 1;
         cumulative_ends.at(end) += 1;
      });
      subtractCumulativeNsFromPositions(
         count_per_local_reference_position,
         static_cast<uint32_t>(sequence_length),
         cumulative_starts,
         cumulative_ends
      );
   }
   {  // countActualFilteredMutations :153-189
      std::map<size_t, const Container*> filter_containers;
      for (size_t idx = 0; idx < filter_bitmap.keys.size(); ++idx) {
         filter_containers[filter_bitmap.keys[idx]] = &filter_bitmap.containers[idx];
      }
      for (const auto& [key, sequence_diff] : sequence_column.vertical_sequence_index.vertical_bitmaps) {
         auto iter = filter_containers.find(key.v_index);
         if (iter != filter_containers.end()) {
            const uint32_t contained_count = containerAndCardinality(*iter->second, sequence_diff);
            counts[key.symbol][key.position] += contained_count;
            count_per_local_reference_position[key.position] -= contained_count;
         }
      }
   }
   for (size_t position_idx = 0; position_idx < sequence_length; ++position_idx) {
      counts[local_reference.at(position_idx)][position_idx] +=
         count_per_local_reference_position[position_idx];
   }
}

void addMutationCountsForFullBitmaps(const SequenceColumn& sequence_column, MutationCounts& counts) {
   const auto local_reference = sequence_column.getLocalReference();
   const size_t sequence_length = local_reference.size();
   std::vector<uint32_t> count_per_local_reference_position(sequence_length);
   initializeCountsWithSequenceCount(count_per_local_reference_position, sequence_column.sequence_count);
   const auto& coverage_index = sequence_column.horizontal_coverage_index;
   for (const auto& [row_id, n_bitmap] : coverage_index.horizontal_bitmaps) {
      n_bitmap.forEach([&](uint32_t position_idx) {
         count_per_local_reference_position[position_idx] -= 1;
      });
   }
   {  // subtractStartAndEndNCounts :92-109
      std::vector<size_t> cumulative_starts(sequence_length + 1);
      std::vector<size_t> cumulative_ends(sequence_length + 1);
      for (const auto& chunk : coverage_index.start_end) {
         for (const auto& [start, end] : chunk) {
            cumulative_starts.at(start) += 1;
            cumulative_ends.at(end) += 1;
         }
      }
      subtractCumulativeNsFromPositions(
         count_per_local_reference_position,
         static_cast<uint32_t>(sequence_length),
         cumulative_starts,
         cumulative_ends
      );
   }
   for (const auto& [key, sequence_diff] : sequence_column.vertical_sequence_index.vertical_bitmaps) {
      counts[key.symbol][key.position] += sequence_diff.card;
      count_per_local_reference_position[key.position] -= sequence_diff.card;
   }
   for (size_t position_idx = 0; position_idx < sequence_length; ++position_idx) {
      counts[local_reference.at(position_idx)][position_idx] +=
         count_per_local_reference_position[position_idx];
   }
}

}  // namespace

MutationCounts calculateMutationsPerPosition(
   const SequenceColumn& sequence_column,
   const CowBitmap& bitmap_filter,
   uint64_t sequence_count_in_column
) {
   const size_t sequence_length = sequence_column.reference_sequence.size();
   MutationCounts counts(sequence_column.alphabet->count, std::vector<uint32_t>(sequence_length, 0));
   const uint64_t filter_cardinality = bitmap_filter.cardinality();
   if (filter_cardinality == sequence_count_in_column) {
      addMutationCountsForFullBitmaps(sequence_column, counts);
   } else if (filter_cardinality > 0) {
      addMutationCountsForMixedBitmaps(sequence_column, bitmap_filter, counts);
   }
   return counts;
}

std::vector<MutationRow> mutationRowsFromCounts(
   const SequenceColumn& sequence_column,
   const MutationCounts& counts,
   double min_proportion
) {
   const Alphabet& alphabet = *sequence_column.alphabet;
   const auto sequence_length = static_cast<uint32_t>(sequence_column.reference_sequence.size());
   std::vector<MutationRow> rows;
   for (uint32_t pos = 0; pos < sequence_length; ++pos) {
      uint32_t total = 0;
      for (const Symbol symbol : alphabet.valid_mutation_symbols) {
         total += counts.at(symbol)[pos];
      }
      if (total == 0) {
         continue;
      }
      const auto threshold_count =
         min_proportion == 0
            ? 0
            : static_cast<uint32_t>(std::ceil(static_cast<double>(total) * min_proportion) - 1);
      const Symbol symbol_in_reference_genome = sequence_column.reference_sequence.at(pos);
      for (const Symbol symbol : alphabet.valid_mutation_symbols) {
         if (symbol_in_reference_genome != symbol) {
            const uint32_t count = counts.at(symbol)[pos];
            if (count > threshold_count) {
               const double proportion = static_cast<double>(count) / static_cast<double>(total);
               rows.push_back(MutationRow{
                  .mutation_from = alphabet.symbolToChar(symbol_in_reference_genome),
                  .mutation_to = alphabet.symbolToChar(symbol),
                  .position = static_cast<int32_t>(pos + 1),
                  .sequence_name = sequence_column.name,
                  .proportion = proportion,
                  .count = static_cast<int32_t>(count),
                  .coverage = static_cast<int32_t>(total)
               });
            }
         }
      }
   }
   return rows;
}

}  // namespace oracle
