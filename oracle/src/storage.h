// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the reference's sequence storage (all paths under /root/reference/src/rhydb):
//   common/nucleotide_symbols.h:23-200, nucleotide_symbols.cpp:10-70   -> Alphabet (nucleotide)
//   common/aa_symbols.h:23-300, aa_symbols.cpp:13-72                    -> Alphabet (amino acid)
//   common/aligned_sequence.cpp:20-122                                  -> extractCoverageAndMutations
//   storage/column/row_id.h:16-37, row_layout.h:27-55, row_layout.cpp:9-23 -> RowLayout
//   storage/column/vertical_sequence_index.h:22-44, .cpp:17-227         -> VerticalSequenceIndex
//   storage/column/horizontal_coverage_index.h:21-98, .cpp:17-73        -> HorizontalCoverageIndex
//   storage/column/sequence_column.h:58-171, .cpp:100-268,312-341       -> SequenceColumn(+Builder)
//   storage/table.cpp:75-94                                             -> Table::bulkInsert/finalize
#pragma once
#include <array>
#include <cstdint>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

#include "container.h"
#include "roaring.h"

namespace oracle {

// user-facing query validation error (illegal_query_exception.h:8, CHECK_SILO_QUERY)
struct IllegalQueryException : std::runtime_error {
   using std::runtime_error::runtime_error;
};
// query_compilation_exception.h
struct QueryCompilationException : std::runtime_error {
   using std::runtime_error::runtime_error;
};
// append/append_exception.h
struct AppendException : std::runtime_error {
   using std::runtime_error::runtime_error;
};

using Symbol = uint8_t;

struct Alphabet {
   std::string symbol_name;             // "Nucleotide" / "AminoAcid"
   uint32_t count = 0;                  // 16 / 28
   std::string chars;                   // symbolToChar by id
   Symbol missing = 0;                  // N / X
   std::vector<Symbol> valid_mutation_symbols;
   std::vector<std::vector<Symbol>> codes_for;
   std::vector<std::vector<Symbol>> ambiguity_symbols;  // derived, nucleotide_symbols.cpp:47-70
   std::array<int16_t, 256> char_to_symbol{};            // -1 = illegal

   [[nodiscard]] std::optional<Symbol> charToSymbol(char character) const {
      const int16_t symbol = char_to_symbol[static_cast<unsigned char>(character)];
      return symbol < 0 ? std::nullopt : std::optional<Symbol>(static_cast<Symbol>(symbol));
   }
   [[nodiscard]] char symbolToChar(Symbol symbol) const { return chars[symbol]; }

   static const Alphabet& nucleotide();
   static const Alphabet& aminoAcid();
};

constexpr uint32_t COLUMN_CHUNK_SIZE = 1U << 16;

struct RowLayout {
   std::vector<uint32_t> chunk_sizes;
   size_t num_rows = 0;

   void appendChunk(uint32_t chunk_size);
   [[nodiscard]] size_t numChunks() const { return chunk_sizes.size(); }
   [[nodiscard]] uint32_t chunkSize(uint16_t chunk_id) const { return chunk_sizes.at(chunk_id); }
   [[nodiscard]] uint32_t numRows() const { return static_cast<uint32_t>(num_rows); }
   [[nodiscard]] Roaring fullBitmap() const;           // row_layout.cpp:9-16
   void complementInPlace(Roaring& bitmap) const;      // row_layout.cpp:18-23
};

struct Coverage {
   uint32_t start = 0;
   uint32_t end = 0;
   std::vector<uint32_t> missing_positions;
};
struct CoverageAndMutations {
   Coverage coverage;
   std::vector<std::pair<uint32_t, Symbol>> mutations;
};
// aligned_sequence.cpp:20-122. Throws AppendException-compatible std::string via `error`.
std::optional<CoverageAndMutations> extractCoverageAndMutationsFromSequence(
   const Alphabet& alphabet,
   std::string_view sequence,
   size_t offset,
   std::string_view reference,
   bool reference_is_missing_somewhere,
   std::string& error
);

struct SequenceDiffKey {
   uint32_t position;
   uint16_t v_index;
   Symbol symbol;
   auto operator<=>(const SequenceDiffKey&) const = default;
};

struct VerticalSequenceIndex {
   std::map<SequenceDiffKey, Container> vertical_bitmaps;
   using const_iterator = std::map<SequenceDiffKey, Container>::const_iterator;

   void addSymbolsToPositions(
      uint32_t position_idx,
      const std::vector<std::vector<uint32_t>>& ids_per_symbol
   );  // .cpp:17-38
   [[nodiscard]] std::pair<const_iterator, const_iterator> getRangeForPosition(uint32_t position_idx
   ) const;  // .cpp:43-55
   [[nodiscard]] std::optional<Symbol> findBetterLocalReferenceSymbol(
      const Alphabet& alphabet,
      uint32_t position_idx,
      Symbol current_local_reference_symbol,
      uint64_t coverage_cardinality
   ) const;  // .cpp:57-116
   std::optional<Symbol> adaptLocalReference(
      const Alphabet& alphabet,
      const Roaring& coverage_bitmap,
      uint32_t position_idx,
      Symbol current_local_reference_symbol
   );  // .cpp:119-164
   [[nodiscard]] Roaring getMatchingContainersAsBitmap(
      uint32_t position_idx,
      const std::vector<Symbol>& symbols
   ) const;  // .cpp:178-205
   [[nodiscard]] std::vector<std::pair<uint16_t, const Container*>> getMatchingContainerViews(
      uint32_t position_idx,
      const std::vector<Symbol>& symbols
   ) const;  // .cpp:208-227
};

struct HorizontalCoverageIndex {
   std::map<uint32_t, Roaring> horizontal_bitmaps;
   std::vector<std::vector<std::pair<uint32_t, uint32_t>>> start_end;
   std::vector<std::pair<uint32_t, uint32_t>> batch_start_ends;

   void insertCoverage(uint16_t chunk_id, uint16_t row_in_chunk, const Coverage& coverage);  // .cpp:17-46
   [[nodiscard]] std::vector<uint64_t> computeCoverageCardinalities(size_t genome_length) const;  // .cpp:53-80
   [[nodiscard]] std::pair<uint32_t, uint32_t> coverageRange(uint32_t global_row_id) const {
      return start_end.at(global_row_id >> 16).at(global_row_id & 0xFFFF);
   }
   [[nodiscard]] Roaring getCoverageBitmapForPosition(uint32_t position) const;  // .h:57-98, BatchSize=1
};

struct BufferedSequence {
   bool is_null = false;
   Coverage coverage;
   std::vector<std::pair<uint32_t, Symbol>> mutations;
};

struct SequenceColumn {
   const Alphabet* alphabet = nullptr;
   std::string name;
   std::vector<Symbol> reference_sequence;  // metadata->reference_sequence (global reference)
   std::string local_reference_sequence_string;
   VerticalSequenceIndex vertical_sequence_index;
   HorizontalCoverageIndex horizontal_coverage_index;
   Roaring null_bitmap;
   uint32_t sequence_count = 0;
   uint16_t num_chunks = 0;
   std::vector<std::vector<std::vector<uint32_t>>> mutation_buffer;  // [position][symbol] -> ids

   // builder state (SequenceColumnBuilder, sequence_column.h:178-246)
   std::vector<BufferedSequence> buffer;

   SequenceColumn(const Alphabet& alphabet, std::string name, const std::string& reference);

   [[nodiscard]] size_t genomeLength() const { return reference_sequence.size(); }
   [[nodiscard]] std::vector<Symbol> getLocalReference() const;
   [[nodiscard]] Symbol getLocalReferencePosition(size_t position) const;

   void insert(std::string_view sequence, uint32_t offset);  // .cpp:312-341
   void insertNull() { buffer.push_back(BufferedSequence{.is_null = true, .coverage = {}, .mutations = {}}); }
   void appendChunk();  // .cpp:100-120 (consumes `buffer`)
   void finalize();     // .cpp:158-212
   void flushBuffer();  // .cpp:260-268
};

// Metadata columns the Selection predicates read: an unindexed StringColumn (string_column.h) and a Date32Column
// (date32_column.h:36-44, .cpp:17-35: sorted = appended in non-decreasing order and without nulls; a null is stored as 0).
// One value per row in layout order.
struct StringValueColumn {
   std::string name;
   std::vector<std::string> values;  // "" for a null row
   std::vector<bool> is_null;
};
struct DateValueColumn {
   std::string name;
   std::vector<int32_t> values;
   std::vector<bool> is_null;
   bool is_sorted = true;
};

struct Table {
   RowLayout row_layout;
   std::vector<std::unique_ptr<SequenceColumn>> columns;  // insertion order
   std::vector<StringValueColumn> string_columns;
   std::vector<DateValueColumn> date_columns;
   std::vector<size_t> chunk_begins;  // dense index of every chunk's first row (set when a value column is added)
   void computeChunkBegins() {
      chunk_begins.assign(row_layout.chunk_sizes.size() + 1, 0);
      for (size_t c = 0; c < row_layout.chunk_sizes.size(); ++c) {
         chunk_begins[c + 1] = chunk_begins[c] + row_layout.chunk_sizes[c];
      }
   }
   // dense index (layout order) of a global row id
   [[nodiscard]] size_t denseRow(uint32_t global_row_id) const {
      size_t begin = 0;
      const uint32_t chunk = global_row_id >> 16;
      for (uint32_t c = 0; c < chunk; ++c) {
         begin += row_layout.chunk_sizes[c];
      }
      return begin + (global_row_id & 0xFFFFu);
   }
   // stand-ins for indexes owned by out-of-scope columns (lineage index, dictionary index,
   // lineage_filter.cpp:77-100): ready-made roaring bitmaps that arrive via IndexScan
   std::map<std::string, Roaring> named_bitmaps;
   size_t buffered_rows = 0;

   SequenceColumn& addColumn(const Alphabet& alphabet, const std::string& name, const std::string& reference);
   [[nodiscard]] SequenceColumn* findColumn(const std::string& name) const;
   // rows are appended across all columns in lockstep; values[i] for columns[i] (nullopt = null)
   void appendRow(const std::vector<std::optional<std::pair<std::string_view, uint32_t>>>& values);
   void flushChunk();  // Table::bulkInsert, table.cpp:75-85
   void finalize();    // table.cpp:87-94 (flushes a partial chunk first, table_inserter.cpp:335-356)
};

}  // namespace oracle
