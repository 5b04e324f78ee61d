// Synthetic inputs of the reference's own benchmarks, produced on the product side so that bench.py
// never touches the oracle: the evolution model of
// /root/reference/performance/sequence_generator.h:113-185 (SequenceTreeGenerator; same std::mt19937
// call sequence, including the quirk that mutateBase never draws 'T') and the "row i =
// evolved[i % |evolved|]" table of writeFullSequenceNdjson (:367-384).
//
// 10 M x 29,903 nt as strings would be 300 GB, so the column is built DIRECTLY in the S1 upload
// format: per chunk and per (position, symbol) the rows come from the periodic row -> sequence map.
// The result is what the reference's ingest + finalize() would store (diffs against the adapted
// local reference, run-optimised containers); tests compare it with the oracle's string ingest.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/silo_b200.h"
#include "table.h"

namespace silo_host {

struct EvolvedTree {
   std::vector<std::string> sequences;
   std::vector<uint32_t> parent;      // parent[0] == 0
   std::vector<uint32_t> generation;  // generation[0] == 0
};

EvolvedTree generateEvolvedSequences(
   const std::string& reference,
   uint64_t seed = 42,
   double mutation_rate = 0.001,
   double death_rate = 0.1,
   size_t generations = 5,
   size_t children_per_node = 3,
   // the symbols mutateBase draws from; default: the reference generator's first four nucleotide symbols
   const std::string& replacement_symbols = "-ACG"
);

// uniformly random A/C/G/T reference (the real SARS-CoV-2 genome is reference data we do not ship)
std::string randomNucleotideReference(size_t length, uint64_t seed);
// uniformly random reference over the twenty standard amino acids (SURVEY.md 8(d) input 4: translation-free genes)
std::string randomAminoAcidReference(size_t length, uint64_t seed);

// The table of performance/co_occurrence_benchmark.cpp (sequence_generator.h:487-526): a uniformly random A/C/G/T
// reference of 100 nt (std::mt19937{42}) and `count` sequences, each the reference with Binomial(length, rate) point
// substitutions at uniformly drawn positions (std::mt19937{1234}, one stream over all rows in id order: the count, then
// per substitution the base and the position -- the right operand of the assignment is evaluated first).
std::string coOccurrenceReference(size_t length = 100);
std::vector<std::string> coOccurrenceSequences(const std::string& reference, size_t count = 2'000'000, double rate = 0.1);

// A column in the upload format that owns its buffers; `desc` points into them.
struct PackedColumn {
   silo_column_desc desc{};
   std::vector<uint8_t> local_reference;
   std::vector<silo_container_desc> containers;
   std::vector<uint8_t> payload;
   std::vector<uint32_t> start_end;
   PackedColumn() = default;
   PackedColumn(const PackedColumn&) = delete;
   PackedColumn& operator=(const PackedColumn&) = delete;
};

// A SHARD of the table with `total_rows` rows where row i holds sequences[i % sequences.size()]
// (full length, offset 0): the chunks first_chunk + k * chunk_stride, k < n_chunks. The local
// reference is adapted over ALL total_rows rows (sequence_column.cpp:158-212), so every shard sees
// the same one.
//   chunk_stride == 1: a contiguous range; containers, row ids and ranges keep their GLOBAL chunk ids
//                      (the shard's table is created with first_chunk).
//   chunk_stride  > 1: an interleaved shard (chunk c belongs to rank c % stride, which balances any
//                      filter on a sorted column); it is a table of its own with LOCAL chunk ids
//                      0..n_chunks-1 (created with first_chunk = 0). Counts and cardinalities are
//                      addends, so nothing has to be translated back.
void buildCycledColumn(
   const Alphabet& alphabet,
   const std::string& reference,
   const std::vector<std::string>& sequences,
   uint64_t total_rows,
   uint32_t first_chunk,
   uint32_t n_chunks,
   unsigned threads,
   PackedColumn& out,
   uint32_t chunk_stride = 1
);
// The short-read table of performance/sequence_generator.h:189-325 (ShortReadGenerator, uniform whole-genome tiling):
// read i covers [offset_i, offset_i + read_length) with offset_i = i * (L - read_length + 1) / count and carries that
// window of evolved[seq_dist(rng)] (std::mt19937(seed + 1000), one draw per read in id order). Built directly in the S1
// upload format for the chunks [first_chunk, first_chunk + n_chunks) (global chunk ids): coverage (start, end) per row,
// diffs against the local reference adapted over ALL reads (sequence_column.cpp:158-212).
struct ShortReads {
   uint64_t count = 0;
   uint32_t read_length = 0;
   std::vector<uint32_t> sequence_of_read;  // the draws, in id order
   [[nodiscard]] uint32_t offsetOf(uint64_t read_id, size_t genome_length) const {
      return static_cast<uint32_t>((read_id * (genome_length - read_length + 1)) / count);
   }
};
ShortReads drawShortReads(size_t n_sequences, uint64_t count, uint32_t read_length, uint64_t seed = 42);
void buildShortReadColumn(
   const Alphabet& alphabet,
   const std::string& reference,
   const std::vector<std::string>& sequences,
   const ShortReads& reads,
   uint32_t first_chunk,
   uint32_t n_chunks,
   unsigned threads,
   PackedColumn& out
);

// chunk sizes of such a shard
std::vector<uint32_t> shardChunkSizes(uint64_t total_rows, uint32_t first_chunk, uint32_t n_chunks, uint32_t chunk_stride);

// chunk sizes of a dense table with total_rows rows: all 65536 except possibly the last
std::vector<uint32_t> denseChunkSizes(uint64_t total_rows);

// metadata stand-ins for the config-2 filter (SURVEY.md §8d input 2)
// rows whose sequence descends from `ancestor` (inclusive), restricted to the shard, as ascending ids
std::vector<uint32_t> lineageRowIds(
   const EvolvedTree& tree,
   uint32_t ancestor,
   uint64_t total_rows,
   uint32_t first_chunk,
   uint32_t n_chunks,
   uint32_t chunk_stride = 1
);
// DateBetween on the sorted synthetic date column date(i) = day0 + (i * span_days) / total_rows:
// one RangeSelection::Range per chunk of the shard (date_between.cpp:94-134), flattened {start,end}
std::vector<uint32_t> sortedDateRanges(
   uint64_t total_rows,
   uint32_t span_days,
   uint32_t from_day,
   uint32_t to_day_inclusive,
   uint32_t first_chunk,
   uint32_t n_chunks,
   uint32_t chunk_stride = 1
);

// Multi-GPU partition scheduler: contiguous chunk ranges, balanced by weight (payload bytes or rows).
// Returns n_ranks + 1 boundaries.
std::vector<uint32_t> partitionChunks(const std::vector<uint64_t>& chunk_weights, uint32_t n_ranks);

}  // namespace silo_host
