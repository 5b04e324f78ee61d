// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// Restates the Mutations / AminoAcidMutations action,
// /root/reference/src/rhydb/query_engine/operators/mutations_node.cpp:
//   initializeCountsWithSequenceCount :39-49 | subtractHorizontalBitmapCounts :51-61
//   subtractCumulativeNsFromPositions :63-90 | subtractStartAndEndNCounts :92-109
//   subtractFilteredNCounts :111-136         | countActualMutations :138-151
//   countActualFilteredMutations :153-189    | accumulateFinalCounts :191-203
//   addMutationCountsForMixedBitmaps :205-237 | addMutationCountsForFullBitmaps :239-266
//   calculateMutationsPerPosition :268-288   | addMutationsToOutput :290-366
// and count_filter_node.cpp:35-71 (count = filter.cardinality()).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "cow_bitmap.h"
#include "storage.h"

namespace oracle {

// counts[symbol][position], SymbolMap<SymbolType, std::vector<uint32_t>>
using MutationCounts = std::vector<std::vector<uint32_t>>;

MutationCounts calculateMutationsPerPosition(
   const SequenceColumn& sequence_column,
   const CowBitmap& bitmap_filter,
   uint64_t sequence_count_in_column
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

// the per-position thresholding of addMutationsToOutput (:307-363) on ready-made counts
std::vector<MutationRow> mutationRowsFromCounts(
   const SequenceColumn& sequence_column,
   const MutationCounts& counts,
   double min_proportion
);

}  // namespace oracle
