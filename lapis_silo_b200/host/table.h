// Host-side view of a table whose sequence columns live in the device pool.
//
// Mirrors what the reference's query compiler reads from rhydb::storage::Table
// (/root/reference/src/rhydb/storage/table.h) on this path -- and nothing else:
//   row_layout (storage/column/row_layout.h:27-55), per sequence column the global reference
//   (sequence_column.h:47-56 metadata->reference_sequence), the adapted local reference
//   (:104, getLocalReferencePosition :135-138) and whether null_bitmap is empty (:111).
// Containers, coverage and N runs are NOT kept on the host: they are resident in HBM.
#pragma once
#include <array>
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/silo_b200.h"

namespace silo_host {

// user-facing validation error: query_engine/illegal_query_exception.h:8 (CHECK_SILO_QUERY)
struct IllegalQueryException : std::runtime_error {
   using std::runtime_error::runtime_error;
};
// query_engine/query_compilation_exception.h
struct QueryCompilationException : std::runtime_error {
   using std::runtime_error::runtime_error;
};
// failures reported by libsilo_b200.so
struct DeviceError : std::runtime_error {
   int status;
   DeviceError(int status, const std::string& message) : std::runtime_error(message), status(status) {}
};

using Symbol = uint8_t;

// common/nucleotide_symbols.h:23-200 and common/aa_symbols.h:23-300, as data
class Alphabet {
  public:
   std::string symbol_name;
   std::string chars;  // symbolToChar, indexed by symbol id
   Symbol missing = 0;
   std::vector<Symbol> valid_mutation_symbols;
   std::vector<uint32_t> codes_for;          // bit set of concrete/ambiguous symbols each id codes for
   std::vector<std::vector<Symbol>> ambiguity_symbols;
   std::array<int8_t, 256> from_char{};

   [[nodiscard]] uint32_t count() const { return static_cast<uint32_t>(chars.size()); }
   [[nodiscard]] std::optional<Symbol> charToSymbol(char character) const {
      const int8_t id = from_char[static_cast<unsigned char>(character)];
      return id < 0 ? std::nullopt : std::optional<Symbol>(static_cast<Symbol>(id));
   }
   [[nodiscard]] char symbolToChar(Symbol symbol) const { return chars[symbol]; }

   static const Alphabet& nucleotide();
   static const Alphabet& aminoAcid();
};

struct RowLayout {
   uint32_t first_chunk = 0;  // global id of chunk 0 of this shard
   std::vector<uint32_t> chunk_sizes;
   [[nodiscard]] uint64_t numRows() const {
      uint64_t total = 0;
      for (uint32_t size : chunk_sizes) {
         total += size;
      }
      return total;
   }
   [[nodiscard]] size_t numChunks() const { return chunk_sizes.size(); }
};

struct SequenceColumnInfo {
   std::string name;
   const Alphabet* alphabet = nullptr;
   std::vector<Symbol> reference_sequence;  // global reference
   std::vector<Symbol> local_reference;     // adapted, sequence_column.cpp:158-212
   bool has_null_rows = false;
   int device_column = -1;  // index returned by silo_gpu_column_upload
};

// A metadata column the filter's Selection predicates read (string_column.h, date32 columns of column_group.h:73-74):
// 32-bit values per row, resident on the device (silo_gpu_value_column_upload). A string column is stored as
// dictionary ids (the dictionary stays here: an equality test looks the literal up once per query); a Date32 column as
// its day numbers, with the host copy kept for the sorted fast path of DateBetween (date_between.cpp:75-79).
struct ValueColumnInfo {
   enum class Type : uint8_t { STRING, DATE } type = Type::STRING;
   std::string name;
   int device_column = -1;                         // index returned by silo_gpu_value_column_upload (-1: host-only table)
   std::map<std::string, uint32_t> dictionary;     // STRING: value -> id
   std::vector<int32_t> dates;                     // DATE: the rows of the shard in layout order
   bool sorted = false;                            // DATE: Date32Column::isSorted()
   bool has_nulls = false;
};

class Table {
  public:
   RowLayout row_layout;
   std::vector<SequenceColumnInfo> columns;
   std::vector<ValueColumnInfo> value_columns;
   // values: dictionary ids in layout order; null_row_ids: ascending global row ids
   void addStringColumn(const std::string& name, const std::vector<std::string>& dictionary, const uint32_t* ids, const std::vector<uint32_t>& null_row_ids);
   void addDateColumn(const std::string& name, const int32_t* days, const std::vector<uint32_t>& null_row_ids);
   [[nodiscard]] const ValueColumnInfo* findValueColumn(const std::string& name) const;
   // stand-ins for indexes owned by out-of-scope columns (LineageIndex, dictionary index): ready-made
   // roaring bitmaps in the portable format, as the reference would hand them over
   // (lineage_filter.cpp:96-99, roaring_serialize.h:15-30)
   // A `resident` one was made device resident once with silo_gpu_bitmap_register (static indexes:
   // the S1 hook registers them next to the columns); the others travel with every program.
   struct NamedBitmap {
      std::vector<uint8_t> bytes;
      bool resident = false;
      uint32_t device_id = 0;
   };
   std::map<std::string, NamedBitmap> named_bitmaps;
   void registerBitmap(const std::string& name, const uint8_t* bytes, uint64_t size, bool resident);

   // Page-locked result buffers (silo_gpu_host_alloc), recycled between queries: the device writes
   // the u32[n_symbols][genome_length] counts straight into them. Thread safe.
   [[nodiscard]] std::shared_ptr<uint32_t> acquireCountsBuffer(size_t n_values) const;

   // ctx == nullptr makes a HOST-ONLY table: the front half of the query compiler (parse -> rewrite ->
   // compile -> lower to a filter program) runs against its metadata, so the lowering can be tested and
   // timed where there is no GPU; nothing is uploaded and every device entry fails loudly (deviceTable()).
   silo_gpu_ctx* ctx = nullptr;
   silo_gpu_table* device = nullptr;
   [[nodiscard]] silo_gpu_table* deviceTable() const {
      if (device == nullptr) {
         throw DeviceError(SILO_E_NO_DEVICE, "host-only table (created without a device context): there is no CPU fallback for device work");
      }
      return device;
   }

  private:
   struct PinnedPool {
      std::mutex mutex;
      std::vector<std::pair<uint32_t*, size_t>> free_buffers;  // {buffer, capacity in values}
      silo_gpu_ctx* ctx = nullptr;
      ~PinnedPool();
   };
   std::shared_ptr<PinnedPool> pinned_pool;  // outlives the table while results are in flight

  public:

   Table(silo_gpu_ctx* ctx, RowLayout layout);
   ~Table();
   Table(const Table&) = delete;
   Table& operator=(const Table&) = delete;

   // S1: uploads the column (silo_gpu_column_upload) and records the host-side metadata
   int addSequenceColumn(
      const std::string& name,
      const Alphabet& alphabet,
      const std::string& global_reference,
      const silo_column_desc& column
   );
   [[nodiscard]] const SequenceColumnInfo* findColumn(const std::string& name) const;
};

void throwOnDeviceError(int status);

// Wall-clock phases of the last query run by the calling thread (microseconds); what
// profiles/e2e_breakdown.py prints. Costs a handful of steady_clock reads per query.
struct QueryProfile {
   double parse_us = 0;       // harness notation -> expression tree
   double compile_us = 0;     // rewrite + compile + lower to a filter program
   double filter_us = 0;      // silo_gpu_filter_eval: staging, H2D, kernel, cardinality D2H
   double counts_us = 0;      // silo_gpu_mutation_counts: kernels + D2H of the counts
   double threshold_us = 0;   // addMutationsToOutput arithmetic on the host
};
QueryProfile& lastQueryProfile();
double nowMicroseconds();

}  // namespace silo_host
