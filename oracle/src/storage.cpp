// ORACLE — TEST INFRASTRUCTURE ONLY. See storage.h for the reference file:line map.
#include "storage.h"

#include <algorithm>
#include <bit>
#include <cstring>

namespace oracle {

namespace {

Alphabet makeAlphabet(
   std::string symbol_name,
   std::string chars,
   Symbol missing,
   std::vector<Symbol> valid,
   std::vector<std::vector<Symbol>> codes_for,
   const std::vector<std::pair<char, Symbol>>& extra_chars
) {
   Alphabet alphabet;
   alphabet.symbol_name = std::move(symbol_name);
   alphabet.count = static_cast<uint32_t>(chars.size());
   alphabet.chars = std::move(chars);
   alphabet.missing = missing;
   alphabet.valid_mutation_symbols = std::move(valid);
   alphabet.codes_for = std::move(codes_for);
   alphabet.char_to_symbol.fill(-1);
   for (uint32_t symbol = 0; symbol < alphabet.count; ++symbol) {
      const char upper = alphabet.chars[symbol];
      alphabet.char_to_symbol[static_cast<unsigned char>(upper)] = static_cast<int16_t>(symbol);
      if (upper >= 'A' && upper <= 'Z') {
         alphabet.char_to_symbol[static_cast<unsigned char>(upper - 'A' + 'a')] =
            static_cast<int16_t>(symbol);
      }
   }
   for (const auto& [character, symbol] : extra_chars) {
      alphabet.char_to_symbol[static_cast<unsigned char>(character)] = static_cast<int16_t>(symbol);
   }
   // AMBIGUITY_SYMBOLS[S] = {Y : CODES_FOR[S] subset of CODES_FOR[Y]}  (nucleotide_symbols.cpp:47-70)
   alphabet.ambiguity_symbols.resize(alphabet.count);
   for (uint32_t symbol = 0; symbol < alphabet.count; ++symbol) {
      for (uint32_t candidate = 0; candidate < alphabet.count; ++candidate) {
         const auto& codes_symbol = alphabet.codes_for[symbol];
         const auto& codes_candidate = alphabet.codes_for[candidate];
         const bool is_superset = std::ranges::all_of(codes_symbol, [&](Symbol coded) {
            return std::ranges::find(codes_candidate, coded) != codes_candidate.end();
         });
         if (is_superset) {
            alphabet.ambiguity_symbols[symbol].push_back(static_cast<Symbol>(candidate));
         }
      }
   }
   return alphabet;
}

std::vector<Symbol> iota(uint32_t count) {
   std::vector<Symbol> all(count);
   for (uint32_t i = 0; i < count; ++i) {
      all[i] = static_cast<Symbol>(i);
   }
   return all;
}

}  // namespace

const Alphabet& Alphabet::nucleotide() {
   // ids: - A C G T R Y S W K M B D H V N   (nucleotide_symbols.h:23-40)
   enum : Symbol { GAP, A, C, G, T, R, Y, S, W, K, M, B, D, H, V, N };
   static const Alphabet instance = makeAlphabet(
      "Nucleotide",
      "-ACGTRYSWKMBDHVN",
      N,
      {GAP, A, C, G, T},
      {{GAP}, {A}, {C}, {G}, {T}, {A, G}, {C, T}, {G, C}, {A, T}, {G, T}, {A, C}, {C, G, T},
       {A, G, T}, {A, C, T}, {A, C, G}, iota(16)},
      {{'U', T}, {'u', T}}
   );
   return instance;
}

const Alphabet& Alphabet::aminoAcid() {
   // ids: - A C D E F G H I K L M N O P Q R S T U V W Y B J Z * X   (aa_symbols.h:23-52)
   enum : Symbol {
      GAP, A, C, D, E, F, G, H, I, K, L, M, N, O, P, Q, R, S, T, U, V, W, Y, B, J, Z, STOP, X
   };
   std::vector<std::vector<Symbol>> codes_for;
   for (Symbol symbol = 0; symbol <= Y; ++symbol) {
      codes_for.push_back({symbol});
   }
   codes_for.push_back({D, N});   // B
   codes_for.push_back({L, I});   // J
   codes_for.push_back({Q, E});   // Z
   codes_for.push_back({STOP});   // *
   codes_for.push_back(iota(28));  // X
   static const Alphabet instance = makeAlphabet(
      "AminoAcid",
      "-ACDEFGHIKLMNOPQRSTUVWYBJZ*X",
      X,
      {GAP, A, C, D, E, F, G, H, I, K, L, M, N, O, P, Q, R, S, T, U, V, W, Y, STOP},
      std::move(codes_for),
      {}
   );
   return instance;
}

void RowLayout::appendChunk(uint32_t chunk_size) {
   if (chunk_size == 0 || chunk_size > COLUMN_CHUNK_SIZE) {
      throw std::runtime_error("RowLayout: chunk size must be in [1, 65536]");
   }
   if (chunk_sizes.size() >= UINT16_MAX) {
      throw std::runtime_error("RowLayout: too many chunks");
   }
   chunk_sizes.push_back(chunk_size);
   num_rows += chunk_size;
}

Roaring RowLayout::fullBitmap() const {
   Roaring result;
   for (size_t chunk_id = 0; chunk_id < chunk_sizes.size(); ++chunk_id) {
      const uint64_t start = static_cast<uint64_t>(chunk_id) << 16;
      result.addRange(start, start + chunk_sizes[chunk_id]);
   }
   return result;
}

void RowLayout::complementInPlace(Roaring& bitmap) const {
   for (size_t chunk_id = 0; chunk_id < chunk_sizes.size(); ++chunk_id) {
      const uint64_t start = static_cast<uint64_t>(chunk_id) << 16;
      bitmap.flip(start, start + chunk_sizes[chunk_id]);
   }
}

std::optional<CoverageAndMutations> extractCoverageAndMutationsFromSequence(
   const Alphabet& alphabet,
   std::string_view sequence,
   size_t offset,
   std::string_view reference,
   bool reference_is_missing_somewhere,
   std::string& error
) {
   CoverageAndMutations result;
   Coverage& coverage = result.coverage;
   const char* const sequence_data = sequence.data();
   const char* const reference_data = reference.data() + offset;
   const size_t length = sequence.size();

   auto process_one = [&](size_t char_in_sequence) -> bool {
      const auto position_idx = static_cast<uint32_t>(char_in_sequence + offset);
      const char character = sequence_data[char_in_sequence];
      const auto symbol = alphabet.charToSymbol(character);
      if (!symbol.has_value()) {
         error = "illegal character '" + std::string(1, character) + "' at position " +
                 std::to_string(position_idx) + " in the input sequence";
         return false;
      }
      if (symbol.value() == alphabet.missing) {
         coverage.missing_positions.push_back(position_idx);
      } else if (symbol != alphabet.charToSymbol(reference_data[char_in_sequence])) {
         result.mutations.emplace_back(position_idx, symbol.value());
      }
      return true;
   };

   size_t char_in_sequence = 0;
   if (!reference_is_missing_somewhere) {
      static_assert(std::endian::native == std::endian::little);
      constexpr size_t WORD_SIZE = sizeof(uint64_t);
      for (; char_in_sequence + WORD_SIZE <= length; char_in_sequence += WORD_SIZE) {
         uint64_t sequence_word;
         uint64_t reference_word;
         std::memcpy(&sequence_word, sequence_data + char_in_sequence, WORD_SIZE);
         std::memcpy(&reference_word, reference_data + char_in_sequence, WORD_SIZE);
         uint64_t differing_bytes = sequence_word ^ reference_word;
         while (differing_bytes != 0) {
            const size_t byte_in_word = static_cast<size_t>(std::countr_zero(differing_bytes)) / 8;
            if (!process_one(char_in_sequence + byte_in_word)) {
               return std::nullopt;
            }
            differing_bytes &= ~(static_cast<uint64_t>(0xFF) << (byte_in_word * 8));
         }
      }
   }
   for (; char_in_sequence < length; ++char_in_sequence) {
      if (reference_is_missing_somewhere ||
          sequence_data[char_in_sequence] != reference_data[char_in_sequence]) {
         if (!process_one(char_in_sequence)) {
            return std::nullopt;
         }
      }
   }

   const auto& missing_positions = coverage.missing_positions;
   size_t leading_missing = 0;
   while (leading_missing < missing_positions.size() &&
          missing_positions[leading_missing] == offset + leading_missing) {
      ++leading_missing;
   }
   size_t trailing_missing = 0;
   while (trailing_missing < missing_positions.size() - leading_missing &&
          missing_positions[missing_positions.size() - 1 - trailing_missing] ==
             offset + length - 1 - trailing_missing) {
      ++trailing_missing;
   }
   if (leading_missing + trailing_missing == length) {
      coverage.missing_positions.clear();
   } else {
      coverage.start = static_cast<uint32_t>(offset + leading_missing);
      coverage.end = static_cast<uint32_t>(offset + length - trailing_missing);
      coverage.missing_positions.erase(
         coverage.missing_positions.end() - static_cast<ptrdiff_t>(trailing_missing),
         coverage.missing_positions.end()
      );
      coverage.missing_positions.erase(
         coverage.missing_positions.begin(),
         coverage.missing_positions.begin() + static_cast<ptrdiff_t>(leading_missing)
      );
   }
   return result;
}

// ---- VerticalSequenceIndex ----

void VerticalSequenceIndex::addSymbolsToPositions(
   uint32_t position_idx,
   const std::vector<std::vector<uint32_t>>& ids_per_symbol
) {
   for (size_t symbol = 0; symbol < ids_per_symbol.size(); ++symbol) {
      const auto& sorted_ids = ids_per_symbol[symbol];
      // splitIdsIntoBatches, .cpp:301-327
      size_t i = 0;
      while (i < sorted_ids.size()) {
         const auto upper_bits = static_cast<uint16_t>(sorted_ids[i] >> 16);
         size_t j = i;
         while (j < sorted_ids.size() && (sorted_ids[j] >> 16) == upper_bits) {
            ++j;
         }
         const SequenceDiffKey key{position_idx, upper_bits, static_cast<Symbol>(symbol)};
         auto iter = vertical_bitmaps.find(key);
         if (iter == vertical_bitmaps.end()) {
            iter = vertical_bitmaps.emplace(key, Container::withCapacity(static_cast<int32_t>(j - i)))
                      .first;
         }
         for (size_t k = i; k < j; ++k) {
            iter->second.add(static_cast<uint16_t>(sorted_ids[k] & 0xFFFF));
         }
         i = j;
      }
   }
}

std::pair<VerticalSequenceIndex::const_iterator, VerticalSequenceIndex::const_iterator>
VerticalSequenceIndex::getRangeForPosition(uint32_t position_idx) const {
   return {
      vertical_bitmaps.lower_bound(SequenceDiffKey{position_idx, 0, 0}),
      vertical_bitmaps.lower_bound(SequenceDiffKey{position_idx + 1, 0, 0})
   };
}

std::optional<Symbol> VerticalSequenceIndex::findBetterLocalReferenceSymbol(
   const Alphabet& alphabet,
   uint32_t position_idx,
   Symbol current_local_reference_symbol,
   uint64_t coverage_cardinality
) const {
   auto [start, end] = getRangeForPosition(position_idx);
   // computeSymbolCountsForPosition, .cpp:57-76 (uint32 arithmetic)
   std::vector<uint32_t> symbol_counts(alphabet.count, 0);
   symbol_counts[current_local_reference_symbol] = static_cast<uint32_t>(coverage_cardinality);
   for (auto it = start; it != end; ++it) {
      symbol_counts[it->first.symbol] += it->second.card;
      symbol_counts[current_local_reference_symbol] -= it->second.card;
   }
   // getSymbolWithHighestCount, .cpp:78-96
   Symbol best_symbol = current_local_reference_symbol;
   uint32_t best_count = symbol_counts[current_local_reference_symbol];
   for (uint32_t symbol = 0; symbol < alphabet.count; ++symbol) {
      if (symbol == current_local_reference_symbol) {
         continue;
      }
      if (symbol_counts[symbol] > best_count) {
         best_symbol = static_cast<Symbol>(symbol);
         best_count = symbol_counts[symbol];
      }
   }
   if (best_symbol == current_local_reference_symbol) {
      return std::nullopt;
   }
   return best_symbol;
}

std::optional<Symbol> VerticalSequenceIndex::adaptLocalReference(
   const Alphabet& alphabet,
   const Roaring& coverage_bitmap,
   uint32_t position_idx,
   Symbol current_local_reference_symbol
) {
   const auto best_symbol = findBetterLocalReferenceSymbol(
      alphabet, position_idx, current_local_reference_symbol, coverage_bitmap.cardinality()
   );
   if (!best_symbol.has_value()) {
      return std::nullopt;
   }
   const Symbol new_reference_symbol = best_symbol.value();
   std::vector<Symbol> all_symbols(alphabet.count);
   for (uint32_t i = 0; i < alphabet.count; ++i) {
      all_symbols[i] = static_cast<Symbol>(i);
   }
   Roaring old_reference_bitmap = coverage_bitmap;
   old_reference_bitmap -= getMatchingContainersAsBitmap(position_idx, all_symbols);

   for (size_t idx = 0; idx < old_reference_bitmap.keys.size(); ++idx) {
      const SequenceDiffKey key{
         position_idx, old_reference_bitmap.keys[idx], current_local_reference_symbol
      };
      vertical_bitmaps.insert({key, old_reference_bitmap.containers[idx]});
   }
   auto [start, end] = getRangeForPosition(position_idx);
   std::vector<uint16_t> v_indices_to_remove;
   for (auto it = start; it != end; ++it) {
      if (it->first.symbol == new_reference_symbol) {
         v_indices_to_remove.push_back(it->first.v_index);
      }
   }
   for (auto v_index : v_indices_to_remove) {
      vertical_bitmaps.erase(SequenceDiffKey{position_idx, v_index, new_reference_symbol});
   }
   return new_reference_symbol;
}

Roaring VerticalSequenceIndex::getMatchingContainersAsBitmap(
   uint32_t position_idx,
   const std::vector<Symbol>& symbols
) const {
   auto [start, end] = getRangeForPosition(position_idx);
   // BitmapBuilderByContainer (bitmap_builder.cpp:5-55): containers arrive by increasing v_index
   Roaring result;
   for (auto it = start; it != end; ++it) {
      if (std::find(symbols.begin(), symbols.end(), it->first.symbol) == symbols.end()) {
         continue;
      }
      if (!result.keys.empty() && result.keys.back() == it->first.v_index) {
         result.containers.back() = containerOr(result.containers.back(), it->second);
      } else {
         result.keys.push_back(it->first.v_index);
         result.containers.push_back(it->second);
      }
   }
   return result;
}

std::vector<std::pair<uint16_t, const Container*>> VerticalSequenceIndex::getMatchingContainerViews(
   uint32_t position_idx,
   const std::vector<Symbol>& symbols
) const {
   auto [start, end] = getRangeForPosition(position_idx);
   std::vector<std::pair<uint16_t, const Container*>> result;
   for (auto it = start; it != end; ++it) {
      if (std::find(symbols.begin(), symbols.end(), it->first.symbol) == symbols.end()) {
         continue;
      }
      result.emplace_back(it->first.v_index, &it->second);
   }
   return result;
}

// ---- HorizontalCoverageIndex ----

void HorizontalCoverageIndex::insertCoverage(
   uint16_t chunk_id,
   uint16_t row_in_chunk,
   const Coverage& coverage
) {
   if (chunk_id == batch_start_ends.size()) {
      batch_start_ends.emplace_back(UINT32_MAX, 0);
   }
   if (chunk_id == start_end.size()) {
      start_end.emplace_back();
   }
   if (chunk_id != start_end.size() - 1 || row_in_chunk != start_end.at(chunk_id).size()) {
      throw std::runtime_error("coverage must be inserted in ascending row order");
   }
   start_end.at(chunk_id).emplace_back(coverage.start, coverage.end);
   auto& [batch_start, batch_end] = batch_start_ends.back();
   batch_start = std::min(batch_start, coverage.start);
   batch_end = std::max(batch_end, coverage.end);

   Roaring horizontal_bitmap =
      Roaring::fromIds(coverage.missing_positions.data(), coverage.missing_positions.size());
   horizontal_bitmap.removeRange(0, coverage.start);
   horizontal_bitmap.removeRange(coverage.end, UINT32_MAX);
   horizontal_bitmap.runOptimize();
   if (horizontal_bitmap.cardinality() > 0) {
      horizontal_bitmaps.emplace(
         (static_cast<uint32_t>(chunk_id) << 16) | row_in_chunk, std::move(horizontal_bitmap)
      );
   }
}

std::vector<uint64_t> HorizontalCoverageIndex::computeCoverageCardinalities(size_t genome_length
) const {
   std::vector<int64_t> coverage_changes(genome_length + 1, 0);
   for (const auto& chunk : start_end) {
      for (const auto& [start, end] : chunk) {
         coverage_changes[start] += 1;
         coverage_changes[end] -= 1;
      }
   }
   for (const auto& [row_id, missing_positions] : horizontal_bitmaps) {
      missing_positions.forEach([&](uint32_t position_idx) {
         coverage_changes[position_idx] -= 1;
         coverage_changes[position_idx + 1] += 1;
      });
   }
   std::vector<uint64_t> cardinalities(genome_length);
   uint64_t cardinality = 0;
   for (size_t position_idx = 0; position_idx < genome_length; ++position_idx) {
      cardinality += static_cast<uint64_t>(coverage_changes[position_idx]);
      cardinalities[position_idx] = static_cast<uint32_t>(cardinality);
   }
   return cardinalities;
}

Roaring HorizontalCoverageIndex::getCoverageBitmapForPosition(uint32_t position) const {
   const uint32_t range_start = position;
   const uint32_t range_end = position + 1;
   Roaring result;
   // BitmapBuilderByRange (bitmap_builder.cpp:57-78): consecutive ids are added as ranges
   uint32_t current_range_start = 0;
   uint32_t current_range_end = 0;
   auto flush = [&]() {
      if (current_range_start < current_range_end) {
         result.addRange(current_range_start, current_range_end);
      }
   };
   for (size_t chunk_id = 0; chunk_id < start_end.size(); ++chunk_id) {
      auto [batch_start, batch_end] = batch_start_ends.at(chunk_id);
      if (batch_end <= range_start || batch_start >= range_end) {
         continue;
      }
      const uint32_t base_row_id = static_cast<uint32_t>(chunk_id) << 16;
      const auto& chunk = start_end[chunk_id];
      for (size_t row_in_chunk = 0; row_in_chunk < chunk.size(); ++row_in_chunk) {
         const uint32_t row_id = base_row_id | static_cast<uint32_t>(row_in_chunk);
         auto [coverage_start, coverage_end] = chunk[row_in_chunk];
         if (std::max(range_start, coverage_start) < std::min(range_end, coverage_end)) {
            if (row_id == current_range_end) {
               current_range_end++;
            } else {
               flush();
               current_range_start = row_id;
               current_range_end = row_id + 1;
            }
         }
      }
   }
   flush();
   for (const auto& [sequence_idx, bitmap] : horizontal_bitmaps) {
      if (bitmap.contains(position)) {
         result.remove(sequence_idx);
      }
   }
   return result;
}

// ---- SequenceColumn ----

SequenceColumn::SequenceColumn(const Alphabet& alphabet, std::string name, const std::string& reference)
    : alphabet(&alphabet),
      name(std::move(name)) {
   if (reference.empty()) {
      throw std::runtime_error("reference sequence must not be empty");
   }
   for (char character : reference) {
      const auto symbol = alphabet.charToSymbol(character);
      if (!symbol.has_value()) {
         throw std::runtime_error("illegal character in reference sequence");
      }
      reference_sequence.push_back(symbol.value());
      local_reference_sequence_string.push_back(alphabet.symbolToChar(symbol.value()));
   }
   mutation_buffer.assign(
      reference_sequence.size(), std::vector<std::vector<uint32_t>>(alphabet.count)
   );
}

std::vector<Symbol> SequenceColumn::getLocalReference() const {
   std::vector<Symbol> local_reference;
   local_reference.reserve(local_reference_sequence_string.size());
   for (const char character : local_reference_sequence_string) {
      local_reference.push_back(alphabet->charToSymbol(character).value());
   }
   return local_reference;
}

Symbol SequenceColumn::getLocalReferencePosition(size_t position) const {
   return alphabet->charToSymbol(local_reference_sequence_string.at(position)).value();
}

void SequenceColumn::insert(std::string_view sequence, uint32_t offset) {
   const size_t genome_length = local_reference_sequence_string.size();
   if (sequence.size() + offset > genome_length) {
      throw AppendException(
         "the sequence '" + std::string(sequence) + "' which was inserted with an offset " +
         std::to_string(offset) + " is larger than the length of the reference genome: " +
         std::to_string(genome_length)
      );
   }
   const bool reference_contains_missing =
      local_reference_sequence_string.find(alphabet->symbolToChar(alphabet->missing)) !=
      std::string::npos;
   std::string error;
   auto coverage_mutations = extractCoverageAndMutationsFromSequence(
      *alphabet, sequence, offset, local_reference_sequence_string, reference_contains_missing, error
   );
   if (!coverage_mutations.has_value()) {
      throw AppendException(error);
   }
   buffer.push_back(BufferedSequence{
      .is_null = false,
      .coverage = std::move(coverage_mutations->coverage),
      .mutations = std::move(coverage_mutations->mutations)
   });
}

void SequenceColumn::appendChunk() {
   for (size_t row_in_chunk = 0; row_in_chunk < buffer.size(); ++row_in_chunk) {
      const auto& buffered = buffer[row_in_chunk];
      const uint32_t global = (static_cast<uint32_t>(num_chunks) << 16) | static_cast<uint32_t>(row_in_chunk);
      sequence_count++;
      if (buffered.is_null) {
         null_bitmap.add(global);
         horizontal_coverage_index.insertCoverage(
            num_chunks, static_cast<uint16_t>(row_in_chunk), Coverage{}
         );
      } else {
         horizontal_coverage_index.insertCoverage(
            num_chunks, static_cast<uint16_t>(row_in_chunk), buffered.coverage
         );
         for (const auto& [position_idx, symbol] : buffered.mutations) {
            mutation_buffer.at(position_idx)[symbol].push_back(global);
         }
      }
   }
   num_chunks++;
   buffer.clear();
   flushBuffer();
}

void SequenceColumn::flushBuffer() {
   for (size_t position_idx = 0; position_idx != mutation_buffer.size(); ++position_idx) {
      vertical_sequence_index.addSymbolsToPositions(
         static_cast<uint32_t>(position_idx), mutation_buffer[position_idx]
      );
      for (auto& ids : mutation_buffer[position_idx]) {
         ids.clear();
      }
   }
}

void SequenceColumn::finalize() {
   flushBuffer();
   const size_t genome_length = genomeLength();
   const std::vector<uint64_t> coverage_cardinalities =
      horizontal_coverage_index.computeCoverageCardinalities(genome_length);
   for (uint32_t position_idx = 0; position_idx < genome_length; ++position_idx) {
      const Symbol current_reference_symbol = getLocalReferencePosition(position_idx);
      if (!vertical_sequence_index
              .findBetterLocalReferenceSymbol(
                 *alphabet, position_idx, current_reference_symbol, coverage_cardinalities.at(position_idx)
              )
              .has_value()) {
         continue;
      }
      const Roaring coverage_bitmap =
         horizontal_coverage_index.getCoverageBitmapForPosition(position_idx);
      const auto new_reference_symbol = vertical_sequence_index.adaptLocalReference(
         *alphabet, coverage_bitmap, position_idx, current_reference_symbol
      );
      local_reference_sequence_string.at(position_idx) =
         alphabet->symbolToChar(new_reference_symbol.value());
   }
   // optimizeBitmaps, .cpp:253-258
   for (auto& [key, sequence_diff] : vertical_sequence_index.vertical_bitmaps) {
      sequence_diff.runOptimize();
   }
}

// ---- Table ----

SequenceColumn& Table::addColumn(
   const Alphabet& alphabet,
   const std::string& name,
   const std::string& reference
) {
   if (row_layout.numChunks() != 0 || buffered_rows != 0) {
      throw std::runtime_error("columns must be added before rows");
   }
   columns.push_back(std::make_unique<SequenceColumn>(alphabet, name, reference));
   return *columns.back();
}

SequenceColumn* Table::findColumn(const std::string& name) const {
   for (const auto& column : columns) {
      if (column->name == name) {
         return column.get();
      }
   }
   return nullptr;
}

void Table::appendRow(const std::vector<std::optional<std::pair<std::string_view, uint32_t>>>& values
) {
   if (values.size() != columns.size()) {
      throw std::runtime_error("appendRow: one value per column required");
   }
   for (size_t i = 0; i < columns.size(); ++i) {
      if (values[i].has_value()) {
         columns[i]->insert(values[i]->first, values[i]->second);
      } else {
         columns[i]->insertNull();
      }
   }
   buffered_rows++;
   if (buffered_rows == COLUMN_CHUNK_SIZE) {
      flushChunk();
   }
}

void Table::flushChunk() {
   if (buffered_rows == 0) {
      return;
   }
   row_layout.appendChunk(static_cast<uint32_t>(buffered_rows));
   for (auto& column : columns) {
      column->appendChunk();
   }
   buffered_rows = 0;
}

void Table::finalize() {
   flushChunk();
   for (auto& column : columns) {
      column->finalize();
   }
}

}  // namespace oracle
