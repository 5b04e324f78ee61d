#include "silo_loader.h"

#include <cstring>
#include <set>

// boost binary archive facts this reader relies on (boost 1.85, observed on the reference's own
// testBaseData/siloSerializedState and restated in oracle/silo_archive.py, which the tests compare
// this reader with):
//   * header: u64 length + "serialization::archive", u16 library version (20), the sizes of int,
//     long, float, double as 4 bytes, u32 0x00000001 (endianness probe)
//   * the FIRST object of a class type is preceded by 1 byte tracking level + 4 bytes class version
//     (5 zero bytes for every class on this path); later objects of the type carry nothing
//   * std::string and collection sizes are u64; std::map = [size u64, item_version u32, items];
//     std::unordered_map = [size u64, bucket_count u64, item_version u32, items];
//     std::vector<std::pair<u32,u32>> = [size u64, raw bytes] (bitwise serialisable)
//   * enums are saved as 32-bit ints (the Symbol of a SequenceDiffKey takes 4 bytes)

namespace silo_host {

namespace {

constexpr uint32_t ROARING_COOKIE_NO_RUNS = 12346;  // RoaringFormatSpec: SERIAL_COOKIE_NO_RUNCONTAINER
constexpr uint32_t ROARING_COOKIE_RUNS = 12347;     // SERIAL_COOKIE (low 16 bits), n_containers - 1 above
constexpr uint32_t NO_OFFSET_THRESHOLD = 4;
constexpr uint32_t ARRAY_MAX_CARDINALITY = 4096;

enum class Seen { UNKNOWN, NO, YES };

struct ArchiveState {
   std::set<std::string> seen_classes;
   Seen roaring = Seen::UNKNOWN;
   Seen row_bitmap_pair = Seen::UNKNOWN;
};

class Cursor {
  public:
   const uint8_t* data;
   uint64_t size;
   uint64_t position;
   ArchiveState* state;

   const uint8_t* take(uint64_t n) {
      if (n > size - position) {
         throw ArchiveFormatError("archive truncated at byte " + std::to_string(position));
      }
      const uint8_t* out = data + position;
      position += n;
      return out;
   }
   template <typename T>
   T scalar() {
      T value;
      std::memcpy(&value, take(sizeof(T)), sizeof(T));
      return value;
   }
   uint8_t u8() { return scalar<uint8_t>(); }
   uint16_t u16() { return scalar<uint16_t>(); }
   uint32_t u32() { return scalar<uint32_t>(); }
   uint64_t u64() { return scalar<uint64_t>(); }
   // collection / string size that must fit into what is left of the archive at `unit` bytes per item
   uint64_t count(uint64_t unit) {
      const uint64_t n = u64();
      if (unit != 0 && n > (size - position) / unit) {
         throw ArchiveFormatError("implausible collection size " + std::to_string(n) + " at byte " + std::to_string(position - 8));
      }
      return n;
   }
   void preamble(const char* what) {
      static const uint8_t zeros[5] = {0, 0, 0, 0, 0};
      if (std::memcmp(take(5), zeros, 5) != 0) {
         throw ArchiveFormatError(std::string("unexpected class info in front of the first ") + what + " at byte " + std::to_string(position - 5));
      }
   }
   void classInfo(const std::string& name) {
      if (state->seen_classes.insert(name).second) {
         preamble(name.c_str());
      }
   }
   // does a serialised roaring::Roaring (u64 size + portable bytes, no class info) start at `at`?
   [[nodiscard]] bool roaringStartsAt(uint64_t at) const {
      if (at > size || size - at < 16) {
         return false;
      }
      uint64_t bytes;
      uint32_t cookie;
      std::memcpy(&bytes, data + at, 8);
      std::memcpy(&cookie, data + at + 8, 4);
      return bytes >= 8 && bytes <= size - at - 8 && (cookie == ROARING_COOKIE_NO_RUNS || (cookie & 0xFFFF) == ROARING_COOKIE_RUNS);
   }
   // roaring_serialize.h:15-46; returns the portable bytes
   std::pair<const uint8_t*, uint64_t> roaring() {
      if (state->roaring == Seen::UNKNOWN) {
         state->roaring = roaringStartsAt(position) ? Seen::YES : Seen::NO;
      }
      if (state->roaring == Seen::NO) {
         preamble("roaring::Roaring");
         state->roaring = Seen::YES;
      }
      const uint64_t bytes = count(1);
      return {take(bytes), bytes};
   }
   // one item of horizontal_bitmaps: std::pair<const u32, roaring::Roaring>
   uint32_t rowBitmapPairKey() {
      if (state->row_bitmap_pair == Seen::UNKNOWN) {
         // [pair class info?] u32 row [Roaring class info?] u64 size, cookie: take the first reading that
         // holds (the two one-preamble readings coincide when the row is 0 as well)
         auto zerosAt = [&](uint64_t at) {
            static const uint8_t zeros[5] = {0, 0, 0, 0, 0};
            return at <= size && size - at >= 5 && std::memcmp(data + at, zeros, 5) == 0;
         };
         const bool roaring_may_be_seen = state->roaring != Seen::NO;
         const bool roaring_may_be_first = state->roaring != Seen::YES;
         if (roaring_may_be_seen && roaringStartsAt(position + 4)) {
            state->row_bitmap_pair = Seen::YES;
            state->roaring = Seen::YES;
         } else if (roaring_may_be_seen && zerosAt(position) && roaringStartsAt(position + 9)) {
            state->row_bitmap_pair = Seen::NO;
            state->roaring = Seen::YES;
         } else if (roaring_may_be_first && zerosAt(position + 4) && roaringStartsAt(position + 9)) {
            state->row_bitmap_pair = Seen::YES;
            state->roaring = Seen::NO;
         } else {
            state->row_bitmap_pair = Seen::NO;
         }
      }
      if (state->row_bitmap_pair == Seen::NO) {
         preamble("pair<u32,Roaring>");
         state->row_bitmap_pair = Seen::YES;
      }
      return u32();
   }
};

uint64_t parseHeader(Cursor& cursor) {
   static const char SIGNATURE[] = "serialization::archive";
   const uint64_t length = cursor.count(1);
   if (length != sizeof(SIGNATURE) - 1 || std::memcmp(cursor.take(length), SIGNATURE, length) != 0) {
      throw ArchiveFormatError("not a boost binary archive");
   }
   const uint16_t version = cursor.u16();
   const uint8_t* sizes = cursor.take(4);
   const uint32_t endian = cursor.u32();
   if (version != 20 || sizes[0] != 4 || sizes[1] != 8 || sizes[2] != 4 || sizes[3] != 8 || endian != 1) {
      throw ArchiveFormatError("unsupported archive flavour (library version " + std::to_string(version) + ")");
   }
   return cursor.position;
}

void pairVector(Cursor& cursor, std::vector<uint32_t>& out) {
   cursor.classInfo("vector<pair<u32,u32>>");
   const uint64_t n = cursor.count(8);
   const uint8_t* bytes = cursor.take(n * 8);
   const size_t before = out.size();
   out.resize(before + 2 * n);
   if (n != 0) {
      std::memcpy(out.data() + before, bytes, n * 8);
   }
}

uint32_t expectedPayloadBytes(uint8_t typecode, uint32_t cardinality, const uint8_t* payload, uint64_t payload_bytes) {
   // roaring_container.h:104-116 / container_write: 1 bitset, 2 array, 3 run
   switch (typecode) {
      case 1:
         return 8192;
      case 2:
         return 2 * cardinality;
      case 3: {
         if (payload_bytes < 2) {
            throw ArchiveFormatError("run container without a run count");
         }
         uint16_t n_runs;
         std::memcpy(&n_runs, payload, 2);
         return 2 + 4 * static_cast<uint32_t>(n_runs);
      }
      default:
         throw ArchiveFormatError("unknown container typecode " + std::to_string(typecode));
   }
}

// Parses the column at cursor.position to its end.
void parseSequenceColumn(Cursor& cursor, const ArchiveColumnSpec& spec, LoadedSequenceColumn& column) {
   const Alphabet& alphabet = *spec.alphabet;
   const std::string sym = alphabet.symbol_name;
   cursor.classInfo("SequenceColumn<" + sym + ">");
   // sequence_column.h:88: local_reference_sequence_string
   const uint64_t length = cursor.count(1);
   if (length != spec.reference.size()) {
      throw ArchiveFormatError("local reference of column " + spec.name + " has length " + std::to_string(length));
   }
   const uint8_t* local_reference = cursor.take(length);
   column.local_reference.resize(length);
   for (uint64_t i = 0; i < length; ++i) {
      const auto symbol = alphabet.charToSymbol(static_cast<char>(local_reference[i]));
      if (!symbol.has_value()) {
         throw ArchiveFormatError("illegal character in the local reference of column " + spec.name);
      }
      column.local_reference[i] = symbol.value();
   }
   // vertical_sequence_index.h:110-112
   cursor.classInfo("VerticalSequenceIndex<" + sym + ">");
   cursor.classInfo("map<SequenceDiffKey<" + sym + ">,RoaringContainer>");
   const uint64_t n_containers = cursor.count(10 + 5 + 8);
   cursor.u32();  // item_version
   column.containers.reserve(n_containers);
   uint64_t previous_key = 0;
   for (uint64_t i = 0; i < n_containers; ++i) {
      cursor.classInfo("pair<SequenceDiffKey<" + sym + ">,RoaringContainer>");
      cursor.classInfo("SequenceDiffKey<" + sym + ">");
      silo_container_desc container{};
      container.position = cursor.u32();
      container.v_index = cursor.u16();
      const uint32_t symbol = cursor.u32();
      if (container.position >= length || symbol >= alphabet.count()) {
         throw ArchiveFormatError("container key out of range in column " + spec.name);
      }
      container.symbol = static_cast<uint8_t>(symbol);
      // std::map order: (position, v_index, symbol), vertical_sequence_index.h:27-29
      const uint64_t key = (static_cast<uint64_t>(container.position) << 24) | (static_cast<uint64_t>(container.v_index) << 8) | symbol;
      if (i != 0 && key <= previous_key) {
         throw ArchiveFormatError("container keys of column " + spec.name + " are not ascending");
      }
      previous_key = key;
      cursor.classInfo("RoaringContainer");
      container.cardinality = cursor.u32();
      container.typecode = cursor.u8();
      const uint64_t payload_bytes = cursor.count(1);
      const uint8_t* payload = cursor.take(payload_bytes);
      if (container.cardinality == 0 || container.cardinality > 65536 ||
          payload_bytes != expectedPayloadBytes(container.typecode, container.cardinality, payload, payload_bytes)) {
         throw ArchiveFormatError("container payload of column " + spec.name + " does not match its typecode / cardinality");
      }
      container.payload_bytes = static_cast<uint32_t>(payload_bytes);
      container.payload_offset = column.payload.size();
      column.payload.insert(column.payload.end(), payload, payload + payload_bytes);
      column.containers.push_back(container);
   }
   // horizontal_coverage_index.h:109-113
   cursor.classInfo("HorizontalCoverageIndex");
   cursor.classInfo("map<u32,Roaring>");
   const uint64_t n_missing = cursor.count(4 + 8 + 8);
   cursor.u32();
   column.missing_offsets.assign(1, 0);
   uint32_t previous_row = 0;
   for (uint64_t i = 0; i < n_missing; ++i) {
      const uint32_t row = cursor.rowBitmapPairKey();
      if (i != 0 && row <= previous_row) {
         throw ArchiveFormatError("rows of the missing-symbol bitmaps are not ascending");
      }
      previous_row = row;
      const auto [bytes, size] = cursor.roaring();
      const size_t runs_before = column.missing_runs.size();
      portableRoaringToRuns(bytes, size, column.missing_runs);
      for (size_t run = runs_before; run < column.missing_runs.size(); run += 2) {
         if (column.missing_runs[run + 1] > length) {
            throw ArchiveFormatError("missing-symbol bitmap reaches beyond the genome");
         }
      }
      column.missing_row_ids.push_back(row);
      column.missing_offsets.push_back(column.missing_runs.size() / 2);
   }
   cursor.classInfo("vector<vector<pair<u32,u32>>>");
   const uint64_t n_chunks = cursor.count(8);
   cursor.u32();
   if (n_chunks >= 65535) {
      throw ArchiveFormatError("more chunks than a row id can address (row_layout.h:44)");
   }
   for (uint64_t chunk = 0; chunk < n_chunks; ++chunk) {
      const size_t before = column.start_end.size();
      pairVector(cursor, column.start_end);
      const size_t rows = (column.start_end.size() - before) / 2;
      if (rows > 65536) {
         throw ArchiveFormatError("chunk with more than 2^16 rows");
      }
      column.chunk_sizes.push_back(static_cast<uint32_t>(rows));
   }
   pairVector(cursor, column.batch_start_ends);
   if (column.batch_start_ends.size() != 2 * n_chunks) {
      throw ArchiveFormatError("batch_start_ends does not have one entry per chunk");
   }
   for (size_t row = 0; row < column.start_end.size(); row += 2) {
      if (column.start_end[row] > column.start_end[row + 1] || column.start_end[row + 1] > length) {
         throw ArchiveFormatError("coverage range outside the genome in column " + spec.name);
      }
   }
   // insertion_index.h:84-100: insertion_positions (position -> InsertionPosition {insertions,
   // three_mer_index}, :28-63) and collected_insertions. The query path does not use them; they are
   // read through because sequence_count, the null bitmap and num_chunks lie behind them.
   cursor.classInfo("InsertionIndex<" + sym + ">");
   cursor.classInfo("unordered_map<u32,InsertionPosition<" + sym + ">>");
   const uint64_t n_positions = cursor.count(4 + 8 + 4 + 8 + 8 + 4);
   cursor.u64();  // bucket_count
   cursor.u32();  // item_version
   column.n_insertion_positions = n_positions;
   for (uint64_t i = 0; i < n_positions; ++i) {
      cursor.classInfo("pair<u32,InsertionPosition<" + sym + ">>");
      if (cursor.u32() > length) {
         throw ArchiveFormatError("insertion position beyond the genome in column " + spec.name);
      }
      cursor.classInfo("InsertionPosition<" + sym + ">");
      cursor.classInfo("vector<Insertion>");
      const uint64_t n_insertions = cursor.count(8 + 8 + 8);
      cursor.u32();
      for (uint64_t insertion = 0; insertion < n_insertions; ++insertion) {
         cursor.classInfo("Insertion");
         cursor.take(cursor.count(1));  // value
         cursor.roaring();              // row_ids
      }
      cursor.classInfo("unordered_map<ThreeMer<" + sym + ">,InsertionIds>");
      const uint64_t n_three_mers = cursor.count(8 + 12 + 8);
      cursor.u64();
      cursor.u32();
      for (uint64_t three_mer = 0; three_mer < n_three_mers; ++three_mer) {
         cursor.classInfo("pair<ThreeMer<" + sym + ">,InsertionIds>");
         cursor.classInfo("ThreeMer<" + sym + ">");  // std::array<Symbol, 3>: element count + the enums as ints
         if (cursor.u64() != 3) {
            throw ArchiveFormatError("a three-mer that does not have three symbols in column " + spec.name);
         }
         for (int symbol = 0; symbol < 3; ++symbol) {
            if (cursor.u32() >= alphabet.count()) {
               throw ArchiveFormatError("three-mer symbol out of range in column " + spec.name);
            }
         }
         cursor.take(4 * cursor.count(4));  // InsertionIds = std::vector<uint32_t>: size + raw values
      }
   }
   cursor.classInfo("unordered_map<u32,unordered_map<string,Roaring>>");
   const uint64_t n_collected = cursor.count(4 + 8 + 8 + 4);
   cursor.u64();
   cursor.u32();
   for (uint64_t i = 0; i < n_collected; ++i) {  // empty once buildIndex has run; same layout rules
      cursor.classInfo("pair<u32,unordered_map<string,Roaring>>");
      cursor.u32();
      cursor.classInfo("unordered_map<string,Roaring>");
      const uint64_t n_values = cursor.count(8 + 8 + 8);
      cursor.u64();
      cursor.u32();
      for (uint64_t value = 0; value < n_values; ++value) {
         cursor.classInfo("pair<string,Roaring>");
         cursor.take(cursor.count(1));
         cursor.roaring();
      }
   }
   // sequence_column.h:35-39,92-95
   cursor.classInfo("SequenceColumnInfo");
   column.sequence_count = cursor.u32();
   column.vertical_bitmaps_size = cursor.u64();
   column.horizontal_bitmaps_size = cursor.u64();
   if (cursor.u32() != column.sequence_count) {
      throw ArchiveFormatError("sequence_count of column " + spec.name + " does not match its info block");
   }
   const auto [null_bytes, null_size] = cursor.roaring();
   std::vector<uint32_t> null_runs;
   // (checked BEFORE the runs are expanded row by row: a crafted megabyte of full run containers would otherwise
   // ask for 2^32 row ids)
   if (portableRoaringToRuns(null_bytes, null_size, null_runs) > column.sequence_count) {
      throw ArchiveFormatError("null bitmap of column " + spec.name + " holds more rows than the column");
   }
   for (size_t run = 0; run < null_runs.size(); run += 2) {
      for (uint32_t row = null_runs[run]; row != null_runs[run + 1]; ++row) {
         column.null_row_ids.push_back(row);
      }
   }
   column.num_chunks = cursor.u16();
   uint64_t rows = 0;
   for (uint32_t chunk_rows : column.chunk_sizes) {
      rows += chunk_rows;
   }
   if (rows != column.sequence_count || column.num_chunks != n_chunks) {
      throw ArchiveFormatError("coverage index of column " + spec.name + " does not match sequence_count / num_chunks");
   }
   column.tail_parsed = true;
}

void fillDescriptor(LoadedSequenceColumn& column) {
   silo_column_desc& desc = column.desc;
   desc = silo_column_desc{};
   desc.struct_size = sizeof(silo_column_desc);
   desc.n_symbols = column.alphabet->count();
   desc.genome_length = static_cast<uint32_t>(column.local_reference.size());
   desc.missing_symbol = column.alphabet->missing;
   desc.local_reference = column.local_reference.data();
   desc.n_containers = column.containers.size();
   desc.containers = column.containers.data();
   desc.payload = column.payload.data();
   desc.payload_bytes = column.payload.size();
   desc.start_end = column.start_end.data();
   desc.n_rows_with_missing = column.missing_row_ids.size();
   desc.missing_row_ids = column.missing_row_ids.data();
   desc.missing_offsets = column.missing_offsets.data();
   desc.missing_runs = column.missing_runs.data();
   desc.n_null_rows = column.null_row_ids.size();
   desc.null_row_ids = column.null_row_ids.data();
}

// next position >= from where a length-prefixed string of spec's length made of the alphabet's
// characters starts (the local reference differs from the global one wherever a position's majority
// symbol does, sequence_column.cpp:158-212, so the reference itself cannot be searched for)
uint64_t findLocalReference(const uint8_t* data, uint64_t size, uint64_t from, const ArchiveColumnSpec& spec) {
   const uint64_t length = spec.reference.size();
   uint8_t prefix[8];
   std::memcpy(prefix, &length, 8);
   for (uint64_t at = from; at + 8 + length <= size; ++at) {
      if (std::memcmp(data + at, prefix, 8) != 0) {
         continue;
      }
      bool legal = true;
      for (uint64_t i = 0; i < length && legal; ++i) {
         legal = spec.alphabet->charToSymbol(static_cast<char>(data[at + 8 + i])).has_value();
      }
      if (legal) {
         return at;
      }
   }
   return UINT64_MAX;
}

}  // namespace

uint64_t portableRoaringToRuns(const uint8_t* bytes, uint64_t size, std::vector<uint32_t>& runs) {
   uint64_t at = 0;
   auto need = [&](uint64_t n) {
      if (n > size - at) {
         throw ArchiveFormatError("roaring bitmap truncated");
      }
   };
   auto read16 = [&](uint64_t where) {
      uint16_t value;
      std::memcpy(&value, bytes + where, 2);
      return value;
   };
   if (size < 4) {
      throw ArchiveFormatError("roaring bitmap truncated");
   }
   uint32_t cookie;
   std::memcpy(&cookie, bytes, 4);
   at = 4;
   uint32_t n_containers = 0;
   const uint8_t* run_flags = nullptr;
   if ((cookie & 0xFFFF) == ROARING_COOKIE_RUNS) {
      n_containers = (cookie >> 16) + 1;
      const uint64_t flag_bytes = (n_containers + 7) / 8;
      need(flag_bytes);
      run_flags = bytes + at;
      at += flag_bytes;
   } else if (cookie == ROARING_COOKIE_NO_RUNS) {
      need(4);
      std::memcpy(&n_containers, bytes + at, 4);
      at += 4;
      if (n_containers > 65536) {
         throw ArchiveFormatError("roaring bitmap with more than 2^16 containers");
      }
   } else {
      throw ArchiveFormatError("not a portable roaring bitmap (cookie " + std::to_string(cookie) + ")");
   }
   need(4ULL * n_containers);
   const uint64_t keys_at = at;
   at += 4ULL * n_containers;
   if (run_flags == nullptr || n_containers >= NO_OFFSET_THRESHOLD) {
      need(4ULL * n_containers);  // offset header, not needed for a sequential read
      at += 4ULL * n_containers;
   }
   uint64_t total = 0;
   const size_t first_run = runs.size();
   auto emit = [&](uint32_t first, uint32_t end_exclusive) {
      if (runs.size() > first_run && first < runs.back()) {
         throw ArchiveFormatError("roaring bitmap values are not ascending");
      }
      if (runs.size() > first_run && runs.back() == first) {
         runs.back() = end_exclusive;
      } else {
         runs.push_back(first);
         runs.push_back(end_exclusive);
      }
      total += end_exclusive - first;
   };
   int64_t previous_key = -1;
   for (uint32_t c = 0; c < n_containers; ++c) {
      const uint32_t key = read16(keys_at + 4ULL * c);
      // (key 0xFFFF would make `high + 65536` wrap to 0 below, and no row id lives there: a table has fewer than
      // 65,535 chunks, row_layout.h:44)
      if (static_cast<int64_t>(key) <= previous_key || key == 0xFFFF) {
         throw ArchiveFormatError("roaring bitmap container keys are not ascending / out of range");
      }
      previous_key = key;
      const uint32_t high = key << 16;
      const uint32_t cardinality = static_cast<uint32_t>(read16(keys_at + 4ULL * c + 2)) + 1;
      const bool is_run = run_flags != nullptr && ((run_flags[c / 8] >> (c % 8)) & 1) != 0;
      if (is_run) {
         need(2);
         const uint32_t n_runs = read16(at);
         at += 2;
         need(4ULL * n_runs);
         for (uint32_t r = 0; r < n_runs; ++r) {
            const uint32_t start = read16(at + 4ULL * r);
            const uint32_t length_minus_one = read16(at + 4ULL * r + 2);
            if (start + length_minus_one > 0xFFFF) {
               throw ArchiveFormatError("run reaches beyond its container");
            }
            emit(high | start, (high | start) + length_minus_one + 1);
         }
         at += 4ULL * n_runs;
      } else if (cardinality <= ARRAY_MAX_CARDINALITY) {
         need(2ULL * cardinality);
         uint32_t run_first = read16(at);
         uint32_t previous = run_first;
         for (uint32_t i = 1; i < cardinality; ++i) {
            const uint32_t value = read16(at + 2ULL * i);
            if (value <= previous) {
               throw ArchiveFormatError("array container values are not ascending");
            }
            if (value != previous + 1) {
               emit(high | run_first, (high | previous) + 1);
               run_first = value;
            }
            previous = value;
         }
         emit(high | run_first, (high | previous) + 1);
         at += 2ULL * cardinality;
      } else {
         need(8192);
         int64_t open = -1;
         for (uint32_t word_index = 0; word_index < 1024; ++word_index) {
            uint64_t word;
            std::memcpy(&word, bytes + at + 8ULL * word_index, 8);
            if (word == 0 && open < 0) {
               continue;
            }
            for (uint32_t bit = 0; bit < 64; ++bit) {
               const bool set = ((word >> bit) & 1) != 0;
               const uint32_t value = word_index * 64 + bit;
               if (set && open < 0) {
                  open = value;
               } else if (!set && open >= 0) {
                  emit(high | static_cast<uint32_t>(open), high | value);
                  open = -1;
               }
            }
         }
         if (open >= 0) {
            emit(high | static_cast<uint32_t>(open), high + 65536);
         }
         at += 8192;
      }
   }
   return total;
}

std::vector<std::unique_ptr<LoadedSequenceColumn>> readSequenceColumns(
   const uint8_t* data,
   uint64_t size,
   const std::vector<ArchiveColumnSpec>& specs,
   const ArchiveReadOptions& options
) {
   ArchiveState state;
   auto toSeen = [](int flag) { return flag < 0 ? Seen::UNKNOWN : (flag != 0 ? Seen::YES : Seen::NO); };
   state.roaring = toSeen(options.roaring_seen);
   state.row_bitmap_pair = toSeen(options.row_bitmap_pair_seen);
   Cursor header{data, size, 0, &state};
   uint64_t search_from = parseHeader(header);
   bool previous_complete = false;  // the previous column was read to its end: the next one starts right there

   std::vector<std::unique_ptr<LoadedSequenceColumn>> out;
   for (const ArchiveColumnSpec& spec : specs) {
      if (spec.alphabet == nullptr) {
         throw std::invalid_argument("column spec without an alphabet");
      }
      const bool first_of_alphabet = state.seen_classes.count("SequenceColumn<" + spec.alphabet->symbol_name + ">") == 0;
      const uint64_t class_info_bytes = first_of_alphabet ? 5 : 0;
      std::unique_ptr<LoadedSequenceColumn> parsed;
      std::string last_error = "no length-prefixed " + spec.alphabet->symbol_name + " string of length " + std::to_string(spec.reference.size());
      uint64_t candidate_from = search_from + class_info_bytes;
      while (parsed == nullptr) {
         const uint64_t at = previous_complete ? candidate_from : findLocalReference(data, size, candidate_from, spec);
         if (at == UINT64_MAX || at > size) {
            throw ArchiveFormatError("sequence column " + spec.name + " not found in the archive: " + last_error);
         }
         ArchiveState attempt_state = state;
         Cursor cursor{data, size, at - class_info_bytes, &attempt_state};
         auto column = std::make_unique<LoadedSequenceColumn>();
         column->name = spec.name;
         column->alphabet = spec.alphabet;
         column->reference = spec.reference;
         try {
            parseSequenceColumn(cursor, spec, *column);
         } catch (const ArchiveFormatError& error) {
            if (previous_complete) {
               throw;
            }
            // a metadata string that looks like a local reference: keep searching
            last_error = error.what();
            candidate_from = at + 1;
            continue;
         }
         state = attempt_state;
         search_from = cursor.position;
         previous_complete = column->tail_parsed;
         parsed = std::move(column);
      }
      fillDescriptor(*parsed);
      out.push_back(std::move(parsed));
   }
   return out;
}

std::unique_ptr<LoadedSequenceColumn> shardOf(const LoadedSequenceColumn& column, uint32_t first_chunk, uint32_t n_chunks) {
   const uint64_t total_chunks = column.chunk_sizes.size();
   if (first_chunk > total_chunks || n_chunks > total_chunks - first_chunk) {
      throw std::invalid_argument("chunk range outside the column");
   }
   const uint32_t chunk_end = first_chunk + n_chunks;
   auto shard = std::make_unique<LoadedSequenceColumn>();
   shard->name = column.name;
   shard->alphabet = column.alphabet;
   shard->reference = column.reference;
   shard->local_reference = column.local_reference;  // global: identical on every shard (sequence_column.h:104)
   for (const silo_container_desc& container : column.containers) {
      if (container.v_index < first_chunk || container.v_index >= chunk_end) {
         continue;
      }
      silo_container_desc copy = container;
      copy.payload_offset = shard->payload.size();
      shard->payload.insert(
         shard->payload.end(),
         column.payload.begin() + static_cast<std::ptrdiff_t>(container.payload_offset),
         column.payload.begin() + static_cast<std::ptrdiff_t>(container.payload_offset + container.payload_bytes)
      );
      shard->containers.push_back(copy);
   }
   uint64_t first_row = 0;
   for (uint32_t chunk = 0; chunk < first_chunk; ++chunk) {
      first_row += column.chunk_sizes[chunk];
   }
   uint64_t rows = 0;
   for (uint32_t chunk = first_chunk; chunk < chunk_end; ++chunk) {
      rows += column.chunk_sizes[chunk];
      shard->chunk_sizes.push_back(column.chunk_sizes[chunk]);
      shard->batch_start_ends.push_back(column.batch_start_ends[2 * chunk]);
      shard->batch_start_ends.push_back(column.batch_start_ends[2 * chunk + 1]);
   }
   shard->start_end.assign(
      column.start_end.begin() + static_cast<std::ptrdiff_t>(2 * first_row),
      column.start_end.begin() + static_cast<std::ptrdiff_t>(2 * (first_row + rows))
   );
   shard->missing_offsets.assign(1, 0);
   for (size_t i = 0; i < column.missing_row_ids.size(); ++i) {
      const uint32_t chunk = column.missing_row_ids[i] >> 16;
      if (chunk < first_chunk || chunk >= chunk_end) {
         continue;
      }
      shard->missing_row_ids.push_back(column.missing_row_ids[i]);
      shard->missing_runs.insert(
         shard->missing_runs.end(),
         column.missing_runs.begin() + static_cast<std::ptrdiff_t>(2 * column.missing_offsets[i]),
         column.missing_runs.begin() + static_cast<std::ptrdiff_t>(2 * column.missing_offsets[i + 1])
      );
      shard->missing_offsets.push_back(shard->missing_runs.size() / 2);
   }
   for (uint32_t row : column.null_row_ids) {
      if ((row >> 16) >= first_chunk && (row >> 16) < chunk_end) {
         shard->null_row_ids.push_back(row);
      }
   }
   shard->sequence_count = static_cast<uint32_t>(rows);
   shard->num_chunks = static_cast<uint16_t>(n_chunks);
   shard->tail_parsed = column.tail_parsed;
   fillDescriptor(*shard);
   return shard;
}

std::unique_ptr<Table> loadTableFromArchive(
   silo_gpu_ctx* ctx,
   const uint8_t* data,
   uint64_t size,
   const std::vector<ArchiveColumnSpec>& specs,
   const ArchiveReadOptions& options,
   uint32_t first_chunk,
   uint32_t n_chunks
) {
   const auto columns = readSequenceColumns(data, size, specs, options);
   if (columns.empty()) {
      throw std::invalid_argument("no sequence columns requested");
   }
   const std::vector<uint32_t>& all_chunks = columns.front()->chunk_sizes;
   if (first_chunk > all_chunks.size()) {
      throw std::invalid_argument("first_chunk beyond the table");
   }
   if (n_chunks == UINT32_MAX) {
      n_chunks = static_cast<uint32_t>(all_chunks.size()) - first_chunk;
   }
   if (n_chunks > all_chunks.size() - first_chunk) {
      throw std::invalid_argument("chunk range outside the table");
   }
   RowLayout layout;
   layout.first_chunk = first_chunk;
   layout.chunk_sizes.assign(all_chunks.begin() + first_chunk, all_chunks.begin() + first_chunk + n_chunks);
   auto table = std::make_unique<Table>(ctx, layout);
   const bool whole = first_chunk == 0 && n_chunks == all_chunks.size();
   for (const auto& column : columns) {
      if (column->chunk_sizes != all_chunks) {
         throw ArchiveFormatError("column " + column->name + " does not share the table's row layout");
      }
      if (whole) {
         table->addSequenceColumn(column->name, *column->alphabet, column->reference, column->desc);
      } else {
         const auto shard = shardOf(*column, first_chunk, n_chunks);
         table->addSequenceColumn(shard->name, *shard->alphabet, shard->reference, shard->desc);
      }
   }
   return table;
}

}  // namespace silo_host
