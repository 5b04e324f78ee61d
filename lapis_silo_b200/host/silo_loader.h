// `.silo` -> device loader: reads the sequence columns of a serialised table straight into the S1
// upload format (silo_column_desc), without rebuilding the reference's std::map / Roaring objects on
// the host and without boost.
//
// The file is what Table::serializeData writes (/root/reference/src/rhydb/storage/table.h:35-42): a
// boost::archive::binary_oarchive (library version 20) of the ColumnGroup's column maps in a fixed
// order (storage/column_group.h:28-61). This reader restates, for the sequence columns only,
//   storage/column/sequence_column.h:86-96            the members of a SequenceColumn in save order
//   storage/column/vertical_sequence_index.h:31-38,110-112   map<{u32 position, u16 v_index, Symbol}, RoaringContainer>
//   roaring_util/roaring_container.h:104-157          cardinality (u32), typecode (u8), container_write bytes as a string
//   storage/column/horizontal_coverage_index.h:109-113       map<u32 row, Roaring>, start_end, batch_start_ends
//   roaring_util/roaring_serialize.h:15-46            size_t size + the portable roaring bytes
// and the boost layout rules those rely on (see silo_loader.cpp). The metadata columns in front of the
// sequence columns are not parsed (out of scope: they reach the device as ready-made bitmaps); every
// sequence column is located by its length-prefixed local reference. The container payloads are
// already in the byte format S1 takes, so they are copied once, back to back, into the payload slab.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "table.h"

namespace silo_host {

struct ArchiveFormatError : std::runtime_error {
   using std::runtime_error::runtime_error;
};

// What the database schema (database_schema.silo, not parsed here) says about one sequence column.
struct ArchiveColumnSpec {
   std::string name;
   const Alphabet* alphabet = nullptr;
   std::string reference;  // the global reference genome (reference_genomes.json)
};

// One sequence column of the archive in the upload format; `desc` points into the vectors.
struct LoadedSequenceColumn {
   std::string name;
   const Alphabet* alphabet = nullptr;
   std::string reference;
   std::vector<uint8_t> local_reference;  // symbol ids
   std::vector<silo_container_desc> containers;
   std::vector<uint8_t> payload;
   std::vector<uint32_t> start_end;          // {start, end} per row, chunks back to back
   std::vector<uint32_t> chunk_sizes;        // rows per chunk (start_end[chunk].size())
   std::vector<uint32_t> batch_start_ends;   // {start, end} per chunk
   std::vector<uint32_t> missing_row_ids;
   std::vector<uint64_t> missing_offsets;
   std::vector<uint32_t> missing_runs;
   std::vector<uint32_t> null_row_ids;
   uint32_t sequence_count = 0;
   uint64_t vertical_bitmaps_size = 0;    // SequenceColumnInfo, sequence_column.h:35-39
   uint64_t horizontal_bitmaps_size = 0;
   uint16_t num_chunks = 0;
   uint64_t n_insertion_positions = 0;  // insertion_index.h:91 (read through, not kept)
   bool tail_parsed = false;  // set once the column was read to its end
   silo_column_desc desc{};
};

// Decodes a bitmap in the portable Roaring format (array, bitset and run containers; cookies 12346 /
// 12347) into ascending {first, end_exclusive} runs, appended to `runs`. Returns the cardinality.
uint64_t portableRoaringToRuns(const uint8_t* bytes, uint64_t size, std::vector<uint32_t>& runs);

struct ArchiveReadOptions {
   // Class types the metadata columns in front of the sequence columns already registered with the
   // archive (their first object then carries no class info). Every table has a primary key column
   // with a null bitmap; the pair type comes with a lineage index. -1 = detect from the bytes.
   int roaring_seen = -1;
   int row_bitmap_pair_seen = -1;
};

// Parses the sequence columns named by `specs` (in the archive's order: nucleotide columns by name,
// then amino-acid columns by name -- std::map order, column_group.h:49-57).
std::vector<std::unique_ptr<LoadedSequenceColumn>> readSequenceColumns(
   const uint8_t* data,
   uint64_t size,
   const std::vector<ArchiveColumnSpec>& specs,
   const ArchiveReadOptions& options = {}
);

// The part of a column that belongs to the chunks [first_chunk, first_chunk + n_chunks): what one rank of
// the row-partitioned table uploads (v_index and row ids stay global).
std::unique_ptr<LoadedSequenceColumn> shardOf(const LoadedSequenceColumn& column, uint32_t first_chunk, uint32_t n_chunks);

// S1 for a saved database: row layout from the first column's coverage index, every column uploaded
// through silo_gpu_column_upload (Table::addSequenceColumn). n_chunks == UINT32_MAX: up to the last chunk.
std::unique_ptr<Table> loadTableFromArchive(
   silo_gpu_ctx* ctx,
   const uint8_t* data,
   uint64_t size,
   const std::vector<ArchiveColumnSpec>& specs,
   const ArchiveReadOptions& options = {},
   uint32_t first_chunk = 0,
   uint32_t n_chunks = UINT32_MAX
);

}  // namespace silo_host
