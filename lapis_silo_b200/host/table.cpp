#include "table.h"

#include <chrono>

#include <algorithm>

namespace silo_host {

namespace {

Alphabet makeAlphabet(
   const std::string& symbol_name,
   const std::string& chars,
   char missing_char,
   const std::string& valid_chars,
   const std::vector<std::pair<char, std::string>>& ambiguity_codes,
   const std::string& extra_aliases  // pairs "xy": character x is read as symbol y
) {
   Alphabet alphabet;
   alphabet.symbol_name = symbol_name;
   alphabet.chars = chars;
   alphabet.from_char.fill(-1);
   for (size_t id = 0; id < chars.size(); ++id) {
      const char upper = chars[id];
      alphabet.from_char[static_cast<unsigned char>(upper)] = static_cast<int8_t>(id);
      if (upper >= 'A' && upper <= 'Z') {
         alphabet.from_char[static_cast<unsigned char>(upper + ('a' - 'A'))] = static_cast<int8_t>(id);
      }
   }
   for (size_t i = 0; i + 1 < extra_aliases.size(); i += 2) {
      alphabet.from_char[static_cast<unsigned char>(extra_aliases[i])] =
         alphabet.from_char[static_cast<unsigned char>(extra_aliases[i + 1])];
   }
   alphabet.missing = alphabet.charToSymbol(missing_char).value();
   for (char c : valid_chars) {
      alphabet.valid_mutation_symbols.push_back(alphabet.charToSymbol(c).value());
   }
   // CODES_FOR: a concrete symbol codes for itself, an ambiguity code for its listed members, the
   // missing symbol for everything (nucleotide_symbols.cpp:10-45, aa_symbols.cpp:13-46)
   alphabet.codes_for.assign(chars.size(), 0);
   for (size_t id = 0; id < chars.size(); ++id) {
      alphabet.codes_for[id] = 1u << id;
   }
   for (const auto& [code, members] : ambiguity_codes) {
      uint32_t bits = 0;
      for (char member : members) {
         bits |= 1u << alphabet.charToSymbol(member).value();
      }
      alphabet.codes_for[alphabet.charToSymbol(code).value()] = bits;
   }
   alphabet.codes_for[alphabet.missing] = (chars.size() == 32 ? 0xFFFFFFFFu : (1u << chars.size()) - 1u);
   // AMBIGUITY_SYMBOLS[s] = every y whose CODES_FOR is a superset of CODES_FOR[s]
   alphabet.ambiguity_symbols.resize(chars.size());
   for (size_t s = 0; s < chars.size(); ++s) {
      for (size_t y = 0; y < chars.size(); ++y) {
         if ((alphabet.codes_for[s] & ~alphabet.codes_for[y]) == 0) {
            alphabet.ambiguity_symbols[s].push_back(static_cast<Symbol>(y));
         }
      }
   }
   return alphabet;
}

}  // namespace

const Alphabet& Alphabet::nucleotide() {
   static const Alphabet instance = makeAlphabet(
      "Nucleotide",
      "-ACGTRYSWKMBDHVN",
      'N',
      "-ACGT",
      {{'R', "AG"}, {'Y', "CT"}, {'S', "GC"}, {'W', "AT"}, {'K', "GT"}, {'M', "AC"},
       {'B', "CGT"}, {'D', "AGT"}, {'H', "ACT"}, {'V', "ACG"}},
      "UTut"
   );
   return instance;
}

const Alphabet& Alphabet::aminoAcid() {
   static const Alphabet instance = makeAlphabet(
      "AminoAcid",
      "-ACDEFGHIKLMNOPQRSTUVWYBJZ*X",
      'X',
      "-ACDEFGHIKLMNOPQRSTUVWY*",
      {{'B', "DN"}, {'J', "LI"}, {'Z', "QE"}},
      ""
   );
   return instance;
}

QueryProfile& lastQueryProfile() {
   thread_local QueryProfile profile;
   return profile;
}

double nowMicroseconds() {
   return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void throwOnDeviceError(int status) {
   if (status < 0) {
      throw DeviceError(status, silo_gpu_last_error());
   }
}

Table::Table(silo_gpu_ctx* ctx, RowLayout layout)
    : row_layout(std::move(layout)),
      ctx(ctx) {
   pinned_pool = std::make_shared<PinnedPool>();
   pinned_pool->ctx = ctx;
   if (ctx == nullptr) {
      return;  // host-only table, see table.h
   }
   throwOnDeviceError(silo_gpu_table_create(
      ctx,
      row_layout.first_chunk,
      row_layout.chunk_sizes.data(),
      static_cast<uint32_t>(row_layout.chunk_sizes.size()),
      &device
   ));
}

Table::~Table() {
   if (device != nullptr) {
      silo_gpu_table_free(device);
   }
}

Table::PinnedPool::~PinnedPool() {
   for (auto& [buffer, capacity] : free_buffers) {
      silo_gpu_host_free(ctx, buffer);
   }
}

std::shared_ptr<uint32_t> Table::acquireCountsBuffer(size_t n_values) const {
   PinnedPool& pool = *pinned_pool;
   uint32_t* buffer = nullptr;
   size_t capacity = 0;
   {
      std::lock_guard<std::mutex> lock(pool.mutex);
      for (size_t i = 0; i < pool.free_buffers.size(); ++i) {
         if (pool.free_buffers[i].second >= n_values) {
            std::tie(buffer, capacity) = pool.free_buffers[i];
            pool.free_buffers.erase(pool.free_buffers.begin() + static_cast<std::ptrdiff_t>(i));
            break;
         }
      }
   }
   if (buffer == nullptr) {
      capacity = n_values;
      buffer = static_cast<uint32_t*>(silo_gpu_host_alloc(ctx, capacity * sizeof(uint32_t)));
      if (buffer == nullptr) {
         throw DeviceError(SILO_E_OUT_OF_MEMORY, silo_gpu_last_error());
      }
   }
   std::shared_ptr<PinnedPool> keep_alive = pinned_pool;
   return std::shared_ptr<uint32_t>(buffer, [keep_alive, capacity](uint32_t* released) {
      std::lock_guard<std::mutex> lock(keep_alive->mutex);
      keep_alive->free_buffers.emplace_back(released, capacity);
   });
}

void Table::registerBitmap(const std::string& name, const uint8_t* bytes, uint64_t size, bool resident) {
   auto existing = named_bitmaps.find(name);
   if (existing != named_bitmaps.end() && existing->second.resident) {
      throwOnDeviceError(silo_gpu_bitmap_unregister(deviceTable(), existing->second.device_id));
      named_bitmaps.erase(existing);
   }
   NamedBitmap bitmap;
   bitmap.bytes.assign(bytes, bytes + size);
   bitmap.resident = resident;
   if (resident) {
      throwOnDeviceError(silo_gpu_bitmap_register(deviceTable(), bitmap.bytes.data(), bitmap.bytes.size(), &bitmap.device_id));
   }
   named_bitmaps[name] = std::move(bitmap);
}

int Table::addSequenceColumn(
   const std::string& name,
   const Alphabet& alphabet,
   const std::string& global_reference,
   const silo_column_desc& column
) {
   if (findColumn(name) != nullptr) {
      throw std::invalid_argument("duplicate column name " + name);
   }
   if (column.n_symbols != alphabet.count() || column.genome_length != global_reference.size()) {
      throw std::invalid_argument("column descriptor does not match the alphabet / reference length");
   }
   SequenceColumnInfo info;
   info.name = name;
   info.alphabet = &alphabet;
   for (char character : global_reference) {
      const auto symbol = alphabet.charToSymbol(character);
      if (!symbol.has_value()) {
         throw std::invalid_argument("illegal character in the reference sequence");
      }
      info.reference_sequence.push_back(symbol.value());
   }
   info.local_reference.assign(column.local_reference, column.local_reference + column.genome_length);
   info.has_null_rows = column.n_null_rows > 0;
   if (device == nullptr) {  // host-only table: metadata only
      info.device_column = static_cast<int>(columns.size());
      columns.push_back(std::move(info));
      return columns.back().device_column;
   }
   const int device_column = silo_gpu_column_upload(device, &column);
   throwOnDeviceError(device_column);
   info.device_column = device_column;
   // the output pass of the Mutations action runs on the device and names the reference genome's symbols
   throwOnDeviceError(silo_gpu_column_set_reference(device, device_column, info.reference_sequence.data()));
   columns.push_back(std::move(info));
   return device_column;
}

namespace {
int uploadValueColumn(const Table& table, const uint32_t* values, const std::vector<uint32_t>& null_row_ids) {
   if (table.device == nullptr) {
      return -1;
   }
   const int index = silo_gpu_value_column_upload(table.device, values, null_row_ids.empty() ? nullptr : null_row_ids.data(), null_row_ids.size());
   if (index < 0) {
      throwOnDeviceError(index);
   }
   return index;
}
}  // namespace

void Table::addStringColumn(const std::string& name, const std::vector<std::string>& dictionary, const uint32_t* ids, const std::vector<uint32_t>& null_row_ids) {
   ValueColumnInfo column;
   column.type = ValueColumnInfo::Type::STRING;
   column.name = name;
   for (size_t id = 0; id < dictionary.size(); ++id) {
      column.dictionary.emplace(dictionary[id], static_cast<uint32_t>(id));
   }
   column.has_nulls = !null_row_ids.empty();
   column.device_column = uploadValueColumn(*this, ids, null_row_ids);
   value_columns.push_back(std::move(column));
}

void Table::addDateColumn(const std::string& name, const int32_t* days, const std::vector<uint32_t>& null_row_ids) {
   ValueColumnInfo column;
   column.type = ValueColumnInfo::Type::DATE;
   column.name = name;
   column.dates.assign(days, days + row_layout.numRows());
   column.has_nulls = !null_row_ids.empty();
   // Date32Column::isSorted: ascending over the whole column and no nulls
   column.sorted = !column.has_nulls && std::is_sorted(column.dates.begin(), column.dates.end());
   column.device_column = uploadValueColumn(*this, reinterpret_cast<const uint32_t*>(days), null_row_ids);
   value_columns.push_back(std::move(column));
}

const ValueColumnInfo* Table::findValueColumn(const std::string& name) const {
   for (const ValueColumnInfo& column : value_columns) {
      if (column.name == name) {
         return &column;
      }
   }
   return nullptr;
}

const SequenceColumnInfo* Table::findColumn(const std::string& name) const {
   for (const auto& column : columns) {
      if (column.name == name) {
         return &column;
      }
   }
   return nullptr;
}

}  // namespace silo_host
