#include "synthetic.h"

#include <algorithm>
#include <cstring>
#include <map>
#include <random>
#include <thread>

namespace silo_host {

EvolvedTree generateEvolvedSequences(
   const std::string& reference,
   uint64_t seed,
   double mutation_rate,
   double death_rate,
   size_t generations,
   size_t children_per_node,
   const std::string& replacement_symbols
) {
   // RNG call order per child: survives -> binomial(#mutations) -> per mutation: position, then
   // base draws until the base changes. mutateBase picks among the first four nucleotide symbols
   // ('-', 'A', 'C', 'G'), as the reference does; an amino-acid gene draws from its valid mutation symbols.
   const std::string& FIRST_FOUR_SYMBOLS = replacement_symbols;
   std::mt19937 rng(seed);
   EvolvedTree tree;
   tree.sequences.push_back(reference);
   tree.parent.push_back(0);
   tree.generation.push_back(0);
   std::vector<uint32_t> frontier = {0};
   std::bernoulli_distribution survives(1.0 - death_rate);
   for (size_t generation = 0; generation < generations; ++generation) {
      std::vector<uint32_t> next_frontier;
      for (uint32_t parent : frontier) {
         for (size_t child = 0; child < children_per_node; ++child) {
            if (!survives(rng)) {
               continue;
            }
            std::string mutated = tree.sequences[parent];
            std::binomial_distribution<size_t> mutation_count(mutated.size(), mutation_rate);
            const size_t n_mutations = mutation_count(rng);
            std::uniform_int_distribution<size_t> position_distribution(0, mutated.size() - 1);
            for (size_t i = 0; i < n_mutations; ++i) {
               const size_t position = position_distribution(rng);
               std::uniform_int_distribution<size_t> base_distribution(0, replacement_symbols.size() - 1);
               char replacement;
               do {
                  replacement = FIRST_FOUR_SYMBOLS[base_distribution(rng)];
               } while (replacement == mutated[position]);
               mutated[position] = replacement;
            }
            tree.sequences.push_back(std::move(mutated));
            tree.parent.push_back(parent);
            tree.generation.push_back(static_cast<uint32_t>(generation + 1));
            next_frontier.push_back(static_cast<uint32_t>(tree.sequences.size() - 1));
         }
      }
      if (next_frontier.empty()) {
         next_frontier.push_back(static_cast<uint32_t>(tree.sequences.size() - 1));
      }
      frontier = std::move(next_frontier);
   }
   return tree;
}

std::string randomNucleotideReference(size_t length, uint64_t seed) {
   static constexpr char BASES[4] = {'A', 'C', 'G', 'T'};
   std::mt19937 rng(seed);
   std::uniform_int_distribution<size_t> base_distribution(0, 3);
   std::string reference(length, 'A');
   for (char& base : reference) {
      base = BASES[base_distribution(rng)];
   }
   return reference;
}

std::string coOccurrenceReference(size_t length) {
   return randomNucleotideReference(length, 42);  // makeCoOccurrenceReference: the same draws
}

std::vector<std::string> coOccurrenceSequences(const std::string& reference, size_t count, double rate) {
   static constexpr char BASES[4] = {'A', 'C', 'G', 'T'};
   std::mt19937 rng{1234};
   std::uniform_int_distribution<size_t> base_distribution(0, 3);
   std::binomial_distribution<size_t> mutation_count(reference.size(), rate);
   std::uniform_int_distribution<size_t> position_distribution(0, reference.size() - 1);
   std::vector<std::string> sequences(count, reference);
   for (std::string& sequence : sequences) {
      const size_t mutations = mutation_count(rng);
      for (size_t i = 0; i < mutations; ++i) {
         const char base = BASES[base_distribution(rng)];  // `sequence[pos_dist(rng)] = bases.at(base_dist(rng))`: right side first
         sequence[position_distribution(rng)] = base;
      }
   }
   return sequences;
}

std::string randomAminoAcidReference(size_t length, uint64_t seed) {
   static constexpr char RESIDUES[] = "ACDEFGHIKLMNPQRSTVWY";
   std::mt19937 rng(seed);
   std::uniform_int_distribution<size_t> residue_distribution(0, 19);
   std::string reference(length, 'A');
   for (char& residue : reference) {
      residue = RESIDUES[residue_distribution(rng)];
   }
   return reference;
}

std::vector<uint32_t> denseChunkSizes(uint64_t total_rows) {
   std::vector<uint32_t> sizes(total_rows / 65536, 65536);
   if (total_rows % 65536 != 0) {
      sizes.push_back(static_cast<uint32_t>(total_rows % 65536));
   }
   return sizes;
}

namespace {

struct DiffGroup {
   uint32_t position;
   Symbol symbol;
   std::vector<uint32_t> members;  // ascending indices into `sequences`
};

// What convert_run_optimize leaves behind ([external] CRoaring): the run form if it serialises
// strictly smaller, otherwise an array up to 4096 values, otherwise a bitset. Appends the
// container_write bytes to `payload`.
void encodeContainer(
   const uint64_t* words,
   uint32_t position,
   uint16_t v_index,
   Symbol symbol,
   std::vector<silo_container_desc>& descs,
   std::vector<uint8_t>& payload
) {
   uint32_t cardinality = 0;
   uint32_t n_runs = 0;
   uint64_t carry = 0;
   for (size_t w = 0; w < 1024; ++w) {
      const uint64_t word = words[w];
      cardinality += static_cast<uint32_t>(__builtin_popcountll(word));
      n_runs += static_cast<uint32_t>(__builtin_popcountll(word & ~((word << 1) | carry)));
      carry = word >> 63;
   }
   if (cardinality == 0) {
      return;
   }
   const size_t size_as_run = 2 + 4 * static_cast<size_t>(n_runs);
   const bool array_fits = cardinality <= 4096;
   const size_t size_non_run = array_fits ? 2 * static_cast<size_t>(cardinality) + 2 : 8192;
   silo_container_desc desc{};
   desc.position = position;
   desc.v_index = v_index;
   desc.symbol = symbol;
   desc.cardinality = cardinality;
   desc.payload_offset = payload.size();
   if (size_as_run < size_non_run) {
      desc.typecode = 3;
      desc.payload_bytes = static_cast<uint32_t>(size_as_run);
      payload.resize(payload.size() + size_as_run);
      uint8_t* dst = payload.data() + desc.payload_offset;
      const auto header = static_cast<uint16_t>(n_runs);
      std::memcpy(dst, &header, 2);
      dst += 2;
      int32_t run_start = -1;
      int32_t previous = -2;
      auto flush = [&]() {
         const uint16_t pair[2] = {static_cast<uint16_t>(run_start), static_cast<uint16_t>(previous - run_start)};
         std::memcpy(dst, pair, 4);
         dst += 4;
      };
      for (size_t w = 0; w < 1024; ++w) {
         uint64_t word = words[w];
         while (word != 0) {
            const auto value = static_cast<int32_t>(w * 64 + static_cast<size_t>(__builtin_ctzll(word)));
            // consume the whole stretch of consecutive ones inside this word at once
            if (value != previous + 1) {
               if (run_start >= 0) {
                  flush();
               }
               run_start = value;
            }
            previous = value;
            word &= word - 1;
         }
      }
      flush();
   } else if (array_fits) {
      desc.typecode = 2;
      desc.payload_bytes = 2 * cardinality;
      payload.resize(payload.size() + desc.payload_bytes);
      auto* dst = reinterpret_cast<uint16_t*>(payload.data() + desc.payload_offset);
      for (size_t w = 0; w < 1024; ++w) {
         uint64_t word = words[w];
         while (word != 0) {
            const auto value = static_cast<uint16_t>(w * 64 + static_cast<size_t>(__builtin_ctzll(word)));
            std::memcpy(dst++, &value, 2);
            word &= word - 1;
         }
      }
   } else {
      desc.typecode = 1;
      desc.payload_bytes = 8192;
      payload.resize(payload.size() + 8192);
      std::memcpy(payload.data() + desc.payload_offset, words, 8192);
   }
   descs.push_back(desc);
}

}  // namespace

void buildCycledColumn(
   const Alphabet& alphabet,
   const std::string& reference,
   const std::vector<std::string>& sequences,
   uint64_t total_rows,
   uint32_t first_chunk,
   uint32_t n_chunks,
   unsigned threads,
   PackedColumn& out,
   uint32_t chunk_stride
) {
   const size_t genome_length = reference.size();
   const size_t n_sequences = sequences.size();
   if (n_sequences == 0 || genome_length == 0) {
      throw std::invalid_argument("buildCycledColumn: empty input");
   }
   const std::vector<uint32_t> all_chunk_sizes = denseChunkSizes(total_rows);
   if (chunk_stride == 0 ||
       (n_chunks > 0 && static_cast<uint64_t>(first_chunk) + static_cast<uint64_t>(n_chunks - 1) * chunk_stride >= all_chunk_sizes.size())) {
      throw std::invalid_argument("buildCycledColumn: shard exceeds the table");
   }
   // the id a chunk carries in the output: global for a contiguous shard, local for an interleaved one
   auto globalChunk = [&](uint32_t local_chunk) { return first_chunk + local_chunk * chunk_stride; };
   auto emittedChunk = [&](uint32_t local_chunk) { return chunk_stride == 1 ? first_chunk + local_chunk : local_chunk; };
   std::vector<Symbol> reference_symbols(genome_length);
   for (size_t p = 0; p < genome_length; ++p) {
      const auto symbol = alphabet.charToSymbol(reference[p]);
      if (!symbol.has_value() || symbol.value() == alphabet.missing) {
         throw std::invalid_argument("buildCycledColumn: the reference must not hold missing / illegal symbols");
      }
      reference_symbols[p] = symbol.value();
   }
   // rows per sequence index over the WHOLE table
   std::vector<uint64_t> rows_of(n_sequences);
   for (size_t e = 0; e < n_sequences; ++e) {
      rows_of[e] = total_rows / n_sequences + (e < total_rows % n_sequences ? 1 : 0);
   }
   // diffs against the initial local reference (= the global one), grouped by (position, symbol)
   std::map<std::pair<uint32_t, Symbol>, std::vector<uint32_t>> diffs;
   for (size_t e = 0; e < n_sequences; ++e) {
      const std::string& sequence = sequences[e];
      if (sequence.size() != genome_length) {
         throw std::invalid_argument("buildCycledColumn: sequences must be full length");
      }
      for (size_t p = 0; p < genome_length; ++p) {
         if (sequence[p] == reference[p]) {
            continue;
         }
         const auto symbol = alphabet.charToSymbol(sequence[p]);
         if (!symbol.has_value() || symbol.value() == alphabet.missing) {
            throw std::invalid_argument("buildCycledColumn: sequences must not hold missing / illegal symbols");
         }
         if (symbol.value() != reference_symbols[p]) {
            diffs[{static_cast<uint32_t>(p), symbol.value()}].push_back(static_cast<uint32_t>(e));
         }
      }
   }
   // local reference adaptation, one finalize() over the whole table
   // (vertical_sequence_index.cpp:57-164): the first symbol (in SYMBOLS order) whose count strictly
   // exceeds the running best replaces the reference symbol; its containers vanish and the old
   // reference symbol gets containers for every covered row that held it
   out.local_reference.assign(reference_symbols.begin(), reference_symbols.end());
   std::vector<DiffGroup> groups;
   auto iter = diffs.begin();
   while (iter != diffs.end()) {
      const uint32_t position = iter->first.first;
      auto position_end = iter;
      std::vector<uint64_t> symbol_rows(alphabet.count(), 0);
      uint64_t differing_rows = 0;
      while (position_end != diffs.end() && position_end->first.first == position) {
         for (uint32_t e : position_end->second) {
            symbol_rows[position_end->first.second] += rows_of[e];
            differing_rows += rows_of[e];
         }
         ++position_end;
      }
      const Symbol current = reference_symbols[position];
      symbol_rows[current] = total_rows - differing_rows;
      Symbol best = current;
      uint64_t best_rows = symbol_rows[current];
      for (uint32_t symbol = 0; symbol < alphabet.count(); ++symbol) {
         if (symbol != current && symbol_rows[symbol] > best_rows) {
            best = static_cast<Symbol>(symbol);
            best_rows = symbol_rows[symbol];
         }
      }
      out.local_reference[position] = best;
      std::vector<bool> differs(n_sequences, false);
      for (auto group = iter; group != position_end; ++group) {
         for (uint32_t e : group->second) {
            differs[e] = true;
         }
         if (group->first.second != best) {
            groups.push_back(DiffGroup{position, group->first.second, group->second});
         }
      }
      if (best != current) {
         DiffGroup old_reference{position, current, {}};
         for (uint32_t e = 0; e < n_sequences; ++e) {
            if (!differs[e]) {
               old_reference.members.push_back(e);
            }
         }
         if (!old_reference.members.empty()) {
            groups.push_back(std::move(old_reference));
         }
      }
      iter = position_end;
   }
   std::sort(groups.begin(), groups.end(), [](const DiffGroup& a, const DiffGroup& b) {
      return a.position != b.position ? a.position < b.position : a.symbol < b.symbol;
   });

   // containers, chunk by chunk
   std::vector<std::vector<silo_container_desc>> chunk_descs(n_chunks);
   std::vector<std::vector<uint8_t>> chunk_payload(n_chunks);
   auto buildChunk = [&](uint32_t local_chunk) {
      const uint32_t global_chunk = globalChunk(local_chunk);
      const uint32_t emitted_chunk = emittedChunk(local_chunk);
      const uint32_t chunk_size = all_chunk_sizes[global_chunk];
      const uint64_t base_row = static_cast<uint64_t>(global_chunk) * 65536;
      const auto phase = static_cast<uint32_t>(base_row % n_sequences);
      auto& descs = chunk_descs[local_chunk];
      auto& payload = chunk_payload[local_chunk];
      std::vector<uint64_t> words(1024);
      for (const DiffGroup& group : groups) {
         std::fill(words.begin(), words.end(), 0);
         for (uint32_t e : group.members) {
            // first row r of the chunk with (base_row + r) % n_sequences == e
            const uint32_t first = (e + static_cast<uint32_t>(n_sequences) - phase) % static_cast<uint32_t>(n_sequences);
            for (uint32_t row = first; row < chunk_size; row += static_cast<uint32_t>(n_sequences)) {
               words[row >> 6] |= uint64_t{1} << (row & 63);
            }
         }
         encodeContainer(words.data(), group.position, static_cast<uint16_t>(emitted_chunk), group.symbol, descs, payload);
      }
   };
   threads = std::max(1u, std::min(threads, n_chunks == 0 ? 1u : n_chunks));
   std::vector<std::thread> workers;
   for (unsigned t = 0; t < threads; ++t) {
      workers.emplace_back([&, t]() {
         for (uint32_t chunk = t; chunk < n_chunks; chunk += threads) {
            buildChunk(chunk);
         }
      });
   }
   for (auto& worker : workers) {
      worker.join();
   }
   uint64_t total_payload = 0;
   uint64_t total_descs = 0;
   for (uint32_t chunk = 0; chunk < n_chunks; ++chunk) {
      total_payload += chunk_payload[chunk].size();
      total_descs += chunk_descs[chunk].size();
   }
   out.containers.clear();
   out.containers.reserve(total_descs);
   out.payload.clear();
   out.payload.reserve(total_payload);
   uint64_t shard_rows = 0;
   for (uint32_t chunk = 0; chunk < n_chunks; ++chunk) {
      const uint64_t payload_base = out.payload.size();
      for (silo_container_desc desc : chunk_descs[chunk]) {
         desc.payload_offset += payload_base;
         out.containers.push_back(desc);
      }
      out.payload.insert(out.payload.end(), chunk_payload[chunk].begin(), chunk_payload[chunk].end());
      std::vector<silo_container_desc>().swap(chunk_descs[chunk]);
      std::vector<uint8_t>().swap(chunk_payload[chunk]);
      shard_rows += all_chunk_sizes[globalChunk(chunk)];
   }
   out.start_end.resize(2 * shard_rows);
   for (uint64_t row = 0; row < shard_rows; ++row) {
      out.start_end[2 * row] = 0;
      out.start_end[2 * row + 1] = static_cast<uint32_t>(genome_length);
   }
   silo_column_desc& desc = out.desc;
   desc = silo_column_desc{};
   desc.struct_size = sizeof(silo_column_desc);
   desc.n_symbols = alphabet.count();
   desc.genome_length = static_cast<uint32_t>(genome_length);
   desc.missing_symbol = alphabet.missing;
   desc.local_reference = out.local_reference.data();
   desc.n_containers = out.containers.size();
   desc.containers = out.containers.data();
   desc.payload = out.payload.data();
   desc.payload_bytes = out.payload.size();
   desc.start_end = out.start_end.data();
}

ShortReads drawShortReads(size_t n_sequences, uint64_t count, uint32_t read_length, uint64_t seed) {
   ShortReads reads;
   reads.count = count;
   reads.read_length = read_length;
   reads.sequence_of_read.resize(count);
   std::mt19937 rng;
   rng.seed(seed + 1000);
   std::uniform_int_distribution<size_t> seq_dist(0, n_sequences - 1);
   for (uint64_t read = 0; read < count; ++read) {
      reads.sequence_of_read[read] = static_cast<uint32_t>(seq_dist(rng));
   }
   return reads;
}

void buildShortReadColumn(
   const Alphabet& alphabet,
   const std::string& reference,
   const std::vector<std::string>& sequences,
   const ShortReads& reads,
   uint32_t first_chunk,
   uint32_t n_chunks,
   unsigned threads,
   PackedColumn& out
) {
   const size_t genome_length = reference.size();
   if (sequences.empty() || reads.read_length == 0 || reads.read_length > genome_length || reads.count == 0) {
      throw std::invalid_argument("buildShortReadColumn: empty input or read_length exceeds the reference");
   }
   const std::vector<uint32_t> all_chunk_sizes = denseChunkSizes(reads.count);
   if (static_cast<uint64_t>(first_chunk) + n_chunks > all_chunk_sizes.size()) {
      throw std::invalid_argument("buildShortReadColumn: shard exceeds the table");
   }
   const uint32_t n_symbols = alphabet.count();
   // the sequences as symbol ids
   std::vector<std::vector<Symbol>> symbols(sequences.size());
   for (size_t e = 0; e < sequences.size(); ++e) {
      if (sequences[e].size() != genome_length) {
         throw std::invalid_argument("buildShortReadColumn: sequences must be full length");
      }
      symbols[e].resize(genome_length);
      for (size_t p = 0; p < genome_length; ++p) {
         const auto symbol = alphabet.charToSymbol(sequences[e][p]);
         if (!symbol.has_value() || symbol.value() == alphabet.missing) {
            throw std::invalid_argument("buildShortReadColumn: sequences must not hold missing / illegal symbols");
         }
         symbols[e][p] = symbol.value();
      }
   }
   std::vector<Symbol> reference_symbols(genome_length);
   for (size_t p = 0; p < genome_length; ++p) {
      reference_symbols[p] = alphabet.charToSymbol(reference[p]).value();
   }
   threads = std::max(1u, threads);
   // symbol counts per position over ALL reads -> the adapted local reference (one finalize() over the whole table)
   std::vector<std::vector<uint32_t>> partial(threads, std::vector<uint32_t>(genome_length * n_symbols, 0));
   {
      std::vector<std::thread> workers;
      for (unsigned t = 0; t < threads; ++t) {
         workers.emplace_back([&, t]() {
            std::vector<uint32_t>& counts = partial[t];
            const uint64_t begin = reads.count * t / threads;
            const uint64_t end = reads.count * (t + 1) / threads;
            for (uint64_t read = begin; read < end; ++read) {
               const uint32_t offset = reads.offsetOf(read, genome_length);
               const Symbol* source = symbols[reads.sequence_of_read[read]].data();
               for (uint32_t k = 0; k < reads.read_length; ++k) {
                  counts[static_cast<size_t>(offset + k) * n_symbols + source[offset + k]]++;
               }
            }
         });
      }
      for (std::thread& worker : workers) {
         worker.join();
      }
   }
   out.local_reference.assign(reference_symbols.begin(), reference_symbols.end());
   for (size_t p = 0; p < genome_length; ++p) {
      // vertical_sequence_index.cpp:57-116: the first symbol (SYMBOLS order) whose count strictly exceeds the running best
      const Symbol current = reference_symbols[p];
      uint64_t best_rows = 0;
      for (unsigned t = 0; t < threads; ++t) {
         best_rows += partial[t][p * n_symbols + current];
      }
      Symbol best = current;
      for (uint32_t symbol = 0; symbol < n_symbols; ++symbol) {
         uint64_t rows = 0;
         for (unsigned t = 0; t < threads; ++t) {
            rows += partial[t][p * n_symbols + symbol];
         }
         if (symbol != current && rows > best_rows) {
            best = static_cast<Symbol>(symbol);
            best_rows = rows;
         }
      }
      out.local_reference[p] = best;
   }
   partial.clear();

   // containers and coverage, chunk by chunk
   std::vector<std::vector<silo_container_desc>> chunk_descs(n_chunks);
   std::vector<std::vector<uint8_t>> chunk_payload(n_chunks);
   uint64_t shard_rows = 0;
   std::vector<uint64_t> chunk_row_begin(n_chunks + 1, 0);
   for (uint32_t local_chunk = 0; local_chunk < n_chunks; ++local_chunk) {
      shard_rows += all_chunk_sizes[first_chunk + local_chunk];
      chunk_row_begin[local_chunk + 1] = shard_rows;
   }
   out.start_end.assign(2 * shard_rows, 0);
   auto buildChunk = [&](uint32_t local_chunk) {
      const uint32_t global_chunk = first_chunk + local_chunk;
      const uint32_t chunk_size = all_chunk_sizes[global_chunk];
      const uint64_t base_read = static_cast<uint64_t>(global_chunk) * 65536;
      const uint32_t window_begin = reads.offsetOf(base_read, genome_length);
      const uint32_t window_end = reads.offsetOf(base_read + chunk_size - 1, genome_length) + reads.read_length;
      const uint32_t window = window_end - window_begin;
      // one bitmap per (position, symbol) that occurs in the chunk, allocated on first use
      std::vector<std::vector<uint64_t>> bitmaps(static_cast<size_t>(window) * n_symbols);
      for (uint32_t row = 0; row < chunk_size; ++row) {
         const uint64_t read = base_read + row;
         const uint32_t offset = reads.offsetOf(read, genome_length);
         out.start_end[2 * (chunk_row_begin[local_chunk] + row)] = offset;
         out.start_end[2 * (chunk_row_begin[local_chunk] + row) + 1] = offset + reads.read_length;
         const Symbol* source = symbols[reads.sequence_of_read[read]].data();
         for (uint32_t k = 0; k < reads.read_length; ++k) {
            const uint32_t position = offset + k;
            const Symbol symbol = source[position];
            if (symbol == out.local_reference[position]) {
               continue;
            }
            std::vector<uint64_t>& words = bitmaps[static_cast<size_t>(position - window_begin) * n_symbols + symbol];
            if (words.empty()) {
               words.assign(1024, 0);
            }
            words[row >> 6] |= uint64_t{1} << (row & 63);
         }
      }
      for (uint32_t position = window_begin; position < window_end; ++position) {
         for (uint32_t symbol = 0; symbol < n_symbols; ++symbol) {
            const std::vector<uint64_t>& words = bitmaps[static_cast<size_t>(position - window_begin) * n_symbols + symbol];
            if (!words.empty()) {
               encodeContainer(words.data(), position, static_cast<uint16_t>(global_chunk), static_cast<Symbol>(symbol), chunk_descs[local_chunk], chunk_payload[local_chunk]);
            }
         }
      }
   };
   {
      std::vector<std::thread> workers;
      const unsigned n_workers = std::max(1u, std::min(threads, n_chunks == 0 ? 1u : n_chunks));
      for (unsigned t = 0; t < n_workers; ++t) {
         workers.emplace_back([&, t]() {
            for (uint32_t local_chunk = t; local_chunk < n_chunks; local_chunk += n_workers) {
               buildChunk(local_chunk);
            }
         });
      }
      for (std::thread& worker : workers) {
         worker.join();
      }
   }
   out.containers.clear();
   out.payload.clear();
   for (uint32_t local_chunk = 0; local_chunk < n_chunks; ++local_chunk) {
      const uint64_t payload_base = out.payload.size();
      out.payload.insert(out.payload.end(), chunk_payload[local_chunk].begin(), chunk_payload[local_chunk].end());
      for (silo_container_desc desc : chunk_descs[local_chunk]) {
         desc.payload_offset += payload_base;
         out.containers.push_back(desc);
      }
   }
   out.desc = silo_column_desc{};
   out.desc.struct_size = sizeof(silo_column_desc);
   out.desc.n_symbols = n_symbols;
   out.desc.genome_length = static_cast<uint32_t>(genome_length);
   out.desc.missing_symbol = alphabet.missing;
   out.desc.local_reference = out.local_reference.data();
   out.desc.n_containers = out.containers.size();
   out.desc.containers = out.containers.data();
   out.desc.payload = out.payload.data();
   out.desc.payload_bytes = out.payload.size();
   out.desc.start_end = out.start_end.data();
}

std::vector<uint32_t> shardChunkSizes(uint64_t total_rows, uint32_t first_chunk, uint32_t n_chunks, uint32_t chunk_stride) {
   const std::vector<uint32_t> all = denseChunkSizes(total_rows);
   std::vector<uint32_t> sizes;
   for (uint32_t local_chunk = 0; local_chunk < n_chunks; ++local_chunk) {
      sizes.push_back(all.at(first_chunk + static_cast<size_t>(local_chunk) * chunk_stride));
   }
   return sizes;
}

std::vector<uint32_t> lineageRowIds(
   const EvolvedTree& tree,
   uint32_t ancestor,
   uint64_t total_rows,
   uint32_t first_chunk,
   uint32_t n_chunks,
   uint32_t chunk_stride
) {
   const size_t n_sequences = tree.sequences.size();
   std::vector<bool> in_lineage(n_sequences, false);
   in_lineage.at(ancestor) = true;
   for (size_t e = ancestor + 1; e < n_sequences; ++e) {  // parents precede their children
      in_lineage[e] = in_lineage[tree.parent[e]];
   }
   const std::vector<uint32_t> sizes = denseChunkSizes(total_rows);
   std::vector<uint32_t> ids;
   for (uint32_t local_chunk = 0; local_chunk < n_chunks; ++local_chunk) {
      const uint32_t chunk = first_chunk + local_chunk * chunk_stride;
      const uint32_t emitted_chunk = chunk_stride == 1 ? chunk : local_chunk;
      const uint64_t base_row = static_cast<uint64_t>(chunk) * 65536;
      for (uint32_t row = 0; row < sizes.at(chunk); ++row) {
         if (in_lineage[(base_row + row) % n_sequences]) {
            ids.push_back((emitted_chunk << 16) | row);
         }
      }
   }
   return ids;
}

std::vector<uint32_t> sortedDateRanges(
   uint64_t total_rows,
   uint32_t span_days,
   uint32_t from_day,
   uint32_t to_day_inclusive,
   uint32_t first_chunk,
   uint32_t n_chunks,
   uint32_t chunk_stride
) {
   // day(i) = floor(i * span_days / total_rows). lower_bound(from) and upper_bound(to) over rows:
   auto firstRowWithDayAtLeast = [&](uint64_t day) -> uint64_t {
      // smallest i with i * span >= day * total  <=>  i >= ceil(day * total / span)
      const unsigned __int128 numerator = static_cast<unsigned __int128>(day) * total_rows;
      const auto row = static_cast<uint64_t>((numerator + span_days - 1) / span_days);
      return std::min<uint64_t>(row, total_rows);
   };
   const uint64_t lower_row = firstRowWithDayAtLeast(from_day);
   const uint64_t upper_row = firstRowWithDayAtLeast(static_cast<uint64_t>(to_day_inclusive) + 1);
   const std::vector<uint32_t> sizes = denseChunkSizes(total_rows);
   std::vector<uint32_t> flat;
   for (uint32_t local_chunk = 0; local_chunk < n_chunks; ++local_chunk) {
      const uint32_t global_chunk = first_chunk + local_chunk * chunk_stride;
      const uint32_t chunk = chunk_stride == 1 ? global_chunk : local_chunk;  // the id the shard's table uses
      const uint64_t base_row = static_cast<uint64_t>(global_chunk) * 65536;
      const uint32_t size = sizes.at(global_chunk);
      auto clampToChunk = [&](uint64_t row) -> uint32_t {
         if (row <= base_row) {
            return 0;
         }
         return static_cast<uint32_t>(std::min<uint64_t>(row - base_row, size));
      };
      const uint32_t lower = clampToChunk(lower_row);
      const uint32_t upper = clampToChunk(upper_row);
      // an index equal to the chunk size is expressed as row 0 of the next chunk (date_between.cpp:113-127)
      flat.push_back(lower == size ? (chunk + 1) << 16 : (chunk << 16) | lower);
      flat.push_back(upper == size ? (chunk + 1) << 16 : (chunk << 16) | upper);
   }
   return flat;
}

std::vector<uint32_t> partitionChunks(const std::vector<uint64_t>& chunk_weights, uint32_t n_ranks) {
   const auto n_chunks = static_cast<uint32_t>(chunk_weights.size());
   std::vector<uint64_t> prefix(n_chunks + 1, 0);
   for (uint32_t chunk = 0; chunk < n_chunks; ++chunk) {
      prefix[chunk + 1] = prefix[chunk] + chunk_weights[chunk];
   }
   const uint64_t total = prefix[n_chunks];
   std::vector<uint32_t> boundaries(n_ranks + 1, n_chunks);
   boundaries[0] = 0;
   for (uint32_t rank = 1; rank < n_ranks; ++rank) {
      // the cut whose prefix weight is nearest to rank/n_ranks of the total ...
      const auto target = static_cast<uint64_t>(static_cast<unsigned __int128>(total) * rank / n_ranks);
      auto cut = static_cast<uint32_t>(std::lower_bound(prefix.begin(), prefix.end(), target) - prefix.begin());
      if (cut > 0 && target - prefix[cut - 1] < prefix[cut] - target) {
         --cut;
      }
      // ... kept monotone, and -- when there are enough chunks -- leaving every rank at least one
      uint32_t lowest = boundaries[rank - 1];
      uint32_t highest = n_chunks;
      if (n_chunks >= n_ranks) {
         lowest += 1;
         highest = n_chunks - (n_ranks - rank);
      }
      boundaries[rank] = std::min(std::max(cut, lowest), highest);
   }
   return boundaries;
}

}  // namespace silo_host
