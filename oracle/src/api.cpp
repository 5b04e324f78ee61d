// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// C entry points of liboracle.so, bound by oracle/oracle.py (ctypes). Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/silo_b200.h"
#include "aggregation.h"
#include "cow_bitmap.h"
#include "expressions.h"
#include "generator.h"
#include "mutations.h"
#include "operators.h"
#include "storage.h"

using namespace oracle;

namespace {
thread_local std::string g_last_error;

template <typename Fn>
int guarded(Fn&& fn) {
   try {
      fn();
      return 0;
   } catch (const IllegalQueryException& error) {
      g_last_error = std::string("IllegalQueryException: ") + error.what();
      return -2;
   } catch (const QueryCompilationException& error) {
      g_last_error = std::string("QueryCompilationException: ") + error.what();
      return -3;
   } catch (const AppendException& error) {
      g_last_error = std::string("AppendException: ") + error.what();
      return -4;
   } catch (const std::exception& error) {
      g_last_error = error.what();
      return -1;
   }
}

struct Export {
   silo_column_desc desc{};
   std::vector<uint8_t> local_reference;
   std::vector<silo_container_desc> containers;
   std::vector<uint8_t> payload;
   std::vector<uint32_t> start_end;
   std::vector<uint32_t> missing_row_ids;
   std::vector<uint64_t> missing_offsets;
   std::vector<uint32_t> missing_runs;
   std::vector<uint32_t> null_row_ids;
};

struct FilterResult {
   CowBitmap bitmap;
   std::vector<uint32_t> ids;
};

struct MutationRows {
   std::vector<MutationRow> rows;
};
}  // namespace

extern "C" {

const char* orc_last_error() {
   return g_last_error.c_str();
}

void* orc_table_new() {
   return new Table();
}
void orc_table_free(void* table) {
   delete static_cast<Table*>(table);
}

// alphabet: 0 nucleotide, 1 amino acid
int orc_table_add_column(void* table, const char* name, int alphabet, const char* reference) {
   return guarded([&] {
      static_cast<Table*>(table)->addColumn(
         alphabet == 0 ? Alphabet::nucleotide() : Alphabet::aminoAcid(), name, reference
      );
   });
}

// layout-only tables (operator-level known-answer tests): RowLayout::of(...)
int orc_table_set_layout(void* table, const uint32_t* chunk_sizes, uint32_t n_chunks) {
   return guarded([&] {
      auto* t = static_cast<Table*>(table);
      if (!t->columns.empty() || t->row_layout.numChunks() != 0) {
         throw std::runtime_error("set_layout needs an empty table");
      }
      for (uint32_t i = 0; i < n_chunks; ++i) {
         t->row_layout.appendChunk(chunk_sizes[i]);
      }
   });
}

// sequences[i] / offsets[i] per column; sequences[i] == NULL -> null value
int orc_table_append_row(void* table, const char* const* sequences, const uint32_t* offsets) {
   return guarded([&] {
      auto* t = static_cast<Table*>(table);
      std::vector<std::optional<std::pair<std::string_view, uint32_t>>> values;
      for (size_t i = 0; i < t->columns.size(); ++i) {
         if (sequences[i] == nullptr) {
            values.emplace_back(std::nullopt);
         } else {
            values.emplace_back(std::make_pair(std::string_view(sequences[i]), offsets[i]));
         }
      }
      t->appendRow(values);
   });
}

// single-column bulk append: row i = sequences[i % n_sequences] at offsets[i % n_sequences]
int orc_table_append_cycled(
   void* table,
   const char* const* sequences,
   const uint32_t* offsets,
   uint64_t n_sequences,
   uint64_t n_rows
) {
   return guarded([&] {
      auto* t = static_cast<Table*>(table);
      if (t->columns.size() != 1) {
         throw std::runtime_error("append_cycled needs a single-column table");
      }
      std::vector<std::string_view> views;
      for (uint64_t i = 0; i < n_sequences; ++i) {
         views.emplace_back(sequences[i]);
      }
      for (uint64_t row = 0; row < n_rows; ++row) {
         const uint64_t pick = row % n_sequences;
         t->appendRow({std::make_pair(views[pick], offsets == nullptr ? 0U : offsets[pick])});
      }
   });
}

int orc_table_flush_chunk(void* table) {
   return guarded([&] { static_cast<Table*>(table)->flushChunk(); });
}
int orc_table_finalize(void* table) {
   return guarded([&] { static_cast<Table*>(table)->finalize(); });
}

int orc_table_register_bitmap(void* table, const char* name, const uint32_t* ids, uint64_t count) {
   return guarded([&] {
      static_cast<Table*>(table)->named_bitmaps[name] = Roaring::fromIds(ids, count);
   });
}

// portable-format bytes of a registered bitmap (what a HOST_BITMAP leaf receives)
int64_t orc_table_bitmap_bytes(void* table, const char* name, uint8_t* out, uint64_t capacity) {
   int64_t size = -1;
   guarded([&] {
      const auto bytes = static_cast<Table*>(table)->named_bitmaps.at(name).write();
      size = static_cast<int64_t>(bytes.size());
      if (out != nullptr && capacity >= bytes.size()) {
         std::memcpy(out, bytes.data(), bytes.size());
      }
   });
   return size;
}

uint32_t orc_table_num_rows(void* table) {
   return static_cast<Table*>(table)->row_layout.numRows();
}
uint32_t orc_table_num_chunks(void* table) {
   return static_cast<uint32_t>(static_cast<Table*>(table)->row_layout.numChunks());
}
int orc_table_chunk_sizes(void* table, uint32_t* out) {
   const auto& sizes = static_cast<Table*>(table)->row_layout.chunk_sizes;
   std::copy(sizes.begin(), sizes.end(), out);
   return 0;
}

int orc_column_local_reference(void* table, const char* column, char* out) {
   return guarded([&] {
      const auto* col = static_cast<Table*>(table)->findColumn(column);
      if (col == nullptr) {
         throw std::runtime_error("no such column");
      }
      std::memcpy(out, col->local_reference_sequence_string.data(), col->local_reference_sequence_string.size());
   });
}

int64_t orc_column_num_containers(void* table, const char* column) {
   const auto* col = static_cast<Table*>(table)->findColumn(column);
   return col == nullptr ? -1 : static_cast<int64_t>(col->vertical_sequence_index.vertical_bitmaps.size());
}

// ---- filters ----

void* orc_filter_eval(void* table, const char* expression) {
   FilterResult* result = nullptr;
   guarded([&] {
      const auto* t = static_cast<Table*>(table);
      const auto parsed = parseExpression(expression);
      auto owned = std::make_unique<FilterResult>();
      owned->bitmap = computeFilter(*parsed, *t);
      owned->ids = owned->bitmap.toRoaring().toVector();
      result = owned.release();
   });
   return result;
}
void orc_filter_free(void* filter) {
   delete static_cast<FilterResult*>(filter);
}
uint64_t orc_filter_cardinality(void* filter) {
   return static_cast<FilterResult*>(filter)->bitmap.cardinality();
}
const uint32_t* orc_filter_ids(void* filter) {
   return static_cast<FilterResult*>(filter)->ids.data();
}
// dense words in the layout silo_gpu_filter_download uses: words[(chunk - first_chunk)*1024 + r/64]
int orc_filter_words(void* filter, uint32_t first_chunk, uint32_t n_chunks, uint64_t* words) {
   return guarded([&] {
      std::memset(words, 0, static_cast<size_t>(n_chunks) * 1024 * 8);
      for (uint32_t id : static_cast<FilterResult*>(filter)->ids) {
         const uint32_t chunk = id >> 16;
         if (chunk < first_chunk || chunk >= first_chunk + n_chunks) {
            continue;
         }
         words[static_cast<size_t>(chunk - first_chunk) * 1024 + ((id & 0xFFFF) >> 6)] |=
            UINT64_C(1) << (id & 63);
      }
   });
}

// ---- Mutations action ----

// counts[symbol * genome_length + position]; filter == NULL -> all rows (Full)
int orc_mutation_counts(void* table, const char* column, void* filter, uint32_t* counts) {
   return guarded([&] {
      const auto* t = static_cast<Table*>(table);
      const auto* col = t->findColumn(column);
      if (col == nullptr) {
         throw std::runtime_error("no such column");
      }
      CowBitmap full;
      const CowBitmap* bitmap = nullptr;
      if (filter == nullptr) {
         full = CowBitmap{t->row_layout.fullBitmap()};
         bitmap = &full;
      } else {
         bitmap = &static_cast<FilterResult*>(filter)->bitmap;
      }
      const auto result = calculateMutationsPerPosition(*col, *bitmap, t->row_layout.numRows());
      const size_t length = col->genomeLength();
      for (size_t symbol = 0; symbol < result.size(); ++symbol) {
         std::memcpy(counts + symbol * length, result[symbol].data(), length * 4);
      }
   });
}

void* orc_mutation_rows(void* table, const char* column, const uint32_t* counts, double min_proportion) {
   MutationRows* result = nullptr;
   guarded([&] {
      const auto* col = static_cast<Table*>(table)->findColumn(column);
      if (col == nullptr) {
         throw std::runtime_error("no such column");
      }
      const size_t length = col->genomeLength();
      MutationCounts by_symbol(col->alphabet->count);
      for (size_t symbol = 0; symbol < by_symbol.size(); ++symbol) {
         by_symbol[symbol].assign(counts + symbol * length, counts + (symbol + 1) * length);
      }
      auto owned = std::make_unique<MutationRows>();
      owned->rows = mutationRowsFromCounts(*col, by_symbol, min_proportion);
      result = owned.release();
   });
   return result;
}
void orc_mutation_rows_free(void* rows) {
   delete static_cast<MutationRows*>(rows);
}
uint64_t orc_mutation_rows_size(void* rows) {
   return static_cast<MutationRows*>(rows)->rows.size();
}
int orc_mutation_rows_get(
   void* rows,
   uint64_t index,
   char* from,
   char* to,
   int32_t* position,
   double* proportion,
   int32_t* count,
   int32_t* coverage
) {
   const auto& row = static_cast<MutationRows*>(rows)->rows.at(index);
   *from = row.mutation_from;
   *to = row.mutation_to;
   *position = row.position;
   *proportion = row.proportion;
   *count = row.count;
   *coverage = row.coverage;
   return 0;
}

// ---- the whole Mutations query, natively (bench.py's CPU baseline legs) ----

// computeFilter -> calculateMutationsPerPosition -> thresholding, on the calling thread, with
// parsing/compiling inside (timer placement of performance/nof_sequence_filter.cpp:43-52).
// Returns the number of output rows, -1 on error.
int64_t orc_mutations_query(void* table, const char* expression, const char* column, double min_proportion, uint64_t* cardinality_out) {
   int64_t n_rows = -1;
   guarded([&] {
      const auto* t = static_cast<Table*>(table);
      const auto* col = t->findColumn(column);
      if (col == nullptr) {
         throw std::runtime_error("no such column");
      }
      const auto parsed = parseExpression(expression);
      const CowBitmap bitmap = computeFilter(*parsed, *t);
      const auto counts = calculateMutationsPerPosition(*col, bitmap, t->row_layout.numRows());
      const auto rows = mutationRowsFromCounts(*col, counts, min_proportion);
      if (cardinality_out != nullptr) {
         *cardinality_out = bitmap.cardinality();
      }
      n_rows = static_cast<int64_t>(rows.size());
   });
   return n_rows;
}

// `threads` workers, each running independent queries back to back (the reference serves one
// request per worker thread, api/api.cpp:39-51, and has no intra-query parallelism) until `seconds`
// have passed or max_queries_per_thread ran.
// Metadata columns for the Selection predicates: one value per row in layout order (after the layout is known).
// values[i] == nullptr / is_null[i] != 0: a null row.
int orc_table_add_string_column(void* table, const char* name, const char* const* values, uint64_t n_rows) {
   return guarded([&] {
      auto* t = static_cast<Table*>(table);
      if (n_rows != t->row_layout.numRows()) {
         throw std::runtime_error("orc_table_add_string_column: one value per row of the layout");
      }
      StringValueColumn column;
      column.name = name;
      for (uint64_t i = 0; i < n_rows; ++i) {
         column.is_null.push_back(values[i] == nullptr);
         column.values.emplace_back(values[i] == nullptr ? "" : values[i]);
      }
      t->computeChunkBegins();
      t->string_columns.push_back(std::move(column));
   });
}

// the same from dictionary ids (large tables: no per-row C strings)
int orc_table_add_string_column_ids(void* table, const char* name, const char* const* dictionary, uint32_t n_values, const uint32_t* ids, uint64_t n_rows) {
   return guarded([&] {
      auto* t = static_cast<Table*>(table);
      if (n_rows != t->row_layout.numRows()) {
         throw std::runtime_error("orc_table_add_string_column_ids: one value per row of the layout");
      }
      StringValueColumn column;
      column.name = name;
      column.is_null.assign(n_rows, false);
      column.values.reserve(n_rows);
      for (uint64_t i = 0; i < n_rows; ++i) {
         if (ids[i] >= n_values) {
            throw std::runtime_error("orc_table_add_string_column_ids: id out of range");
         }
         column.values.emplace_back(dictionary[ids[i]]);
      }
      t->computeChunkBegins();
      t->string_columns.push_back(std::move(column));
   });
}

int orc_table_add_date_column(void* table, const char* name, const int32_t* days, const uint8_t* is_null, uint64_t n_rows) {
   return guarded([&] {
      auto* t = static_cast<Table*>(table);
      if (n_rows != t->row_layout.numRows()) {
         throw std::runtime_error("orc_table_add_date_column: one value per row of the layout");
      }
      DateValueColumn column;
      column.name = name;
      std::optional<int32_t> last;
      for (uint64_t i = 0; i < n_rows; ++i) {  // Date32Column::appendChunk, date32_column.cpp:17-35
         const bool null = is_null != nullptr && is_null[i] != 0;
         column.is_null.push_back(null);
         column.values.push_back(null ? 0 : days[i]);
         if (null) {
            column.is_sorted = false;
         } else {
            if (last.has_value() && days[i] < *last) {
               column.is_sorted = false;
            }
            last = days[i];
         }
      }
      t->computeChunkBegins();
      t->date_columns.push_back(std::move(column));
   });
}

int orc_mutations_query_bench(
   void* table,
   const char* expression,
   const char* column,
   double min_proportion,
   uint32_t threads,
   double seconds,
   uint64_t max_queries_per_thread,
   uint64_t* queries_out,
   double* elapsed_out,
   uint64_t* cardinality_out
) {
   return guarded([&] {
      if (threads == 0) {
         throw std::runtime_error("threads must be > 0");
      }
      std::vector<uint64_t> done(threads, 0);
      std::vector<int> failed(threads, 0);
      const auto started = std::chrono::steady_clock::now();
      auto worker = [&](uint32_t index) {
         while (true) {
            uint64_t cardinality = 0;
            if (orc_mutations_query(table, expression, column, min_proportion, &cardinality) < 0) {
               failed[index] = 1;
               return;
            }
            if (index == 0 && cardinality_out != nullptr) {
               *cardinality_out = cardinality;
            }
            ++done[index];
            const double elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - started).count();
            if (elapsed >= seconds || done[index] >= max_queries_per_thread) {
               return;
            }
         }
      };
      std::vector<std::thread> pool;
      for (uint32_t index = 1; index < threads; ++index) {
         pool.emplace_back(worker, index);
      }
      worker(0);
      for (std::thread& thread : pool) {
         thread.join();
      }
      *elapsed_out = std::chrono::duration<double>(std::chrono::steady_clock::now() - started).count();
      uint64_t total = 0;
      for (uint32_t index = 0; index < threads; ++index) {
         if (failed[index] != 0) {
            throw std::runtime_error("a query failed");
         }
         total += done[index];
      }
      *queries_out = total;
   });
}

// ---- S1 interchange: export a column as the upload format / import one ----

void* orc_column_export(void* table, const char* column, uint32_t first_chunk, uint32_t n_chunks) {
   Export* result = nullptr;
   guarded([&] {
      const auto* t = static_cast<Table*>(table);
      const auto* col = t->findColumn(column);
      if (col == nullptr) {
         throw std::runtime_error("no such column");
      }
      auto owned = std::make_unique<Export>();
      Export& e = *owned;
      const uint32_t chunk_end = first_chunk + n_chunks;
      for (Symbol symbol : col->getLocalReference()) {
         e.local_reference.push_back(symbol);
      }
      for (const auto& [key, container] : col->vertical_sequence_index.vertical_bitmaps) {
         if (key.v_index < first_chunk || key.v_index >= chunk_end) {
            continue;
         }
         silo_container_desc d{};
         d.position = key.position;
         d.v_index = key.v_index;
         d.symbol = key.symbol;
         d.typecode = container.type;
         d.cardinality = container.card;
         d.payload_bytes = static_cast<uint32_t>(container.sizeInBytes());
         d.payload_offset = e.payload.size();
         e.payload.resize(e.payload.size() + d.payload_bytes);
         container.write(e.payload.data() + d.payload_offset);
         e.containers.push_back(d);
      }
      for (uint32_t chunk = first_chunk; chunk < chunk_end; ++chunk) {
         for (const auto& [start, end] : col->horizontal_coverage_index.start_end.at(chunk)) {
            e.start_end.push_back(start);
            e.start_end.push_back(end);
         }
      }
      e.missing_offsets.push_back(0);
      for (const auto& [row_id, bitmap] : col->horizontal_coverage_index.horizontal_bitmaps) {
         if ((row_id >> 16) < first_chunk || (row_id >> 16) >= chunk_end) {
            continue;
         }
         e.missing_row_ids.push_back(row_id);
         uint32_t run_start = 0;
         uint32_t prev = 0;
         bool open = false;
         bitmap.forEach([&](uint32_t position) {
            if (open && position == prev + 1) {
               prev = position;
               return;
            }
            if (open) {
               e.missing_runs.push_back(run_start);
               e.missing_runs.push_back(prev + 1);
            }
            run_start = position;
            prev = position;
            open = true;
         });
         if (open) {
            e.missing_runs.push_back(run_start);
            e.missing_runs.push_back(prev + 1);
         }
         e.missing_offsets.push_back(e.missing_runs.size() / 2);
      }
      col->null_bitmap.forEach([&](uint32_t row_id) {
         if ((row_id >> 16) >= first_chunk && (row_id >> 16) < chunk_end) {
            e.null_row_ids.push_back(row_id);
         }
      });
      e.desc.struct_size = sizeof(silo_column_desc);
      e.desc.n_symbols = col->alphabet->count;
      e.desc.genome_length = static_cast<uint32_t>(col->genomeLength());
      e.desc.missing_symbol = col->alphabet->missing;
      e.desc.local_reference = e.local_reference.data();
      e.desc.n_containers = e.containers.size();
      e.desc.containers = e.containers.data();
      e.desc.payload = e.payload.data();
      e.desc.payload_bytes = e.payload.size();
      e.desc.start_end = e.start_end.data();
      e.desc.n_rows_with_missing = e.missing_row_ids.size();
      e.desc.missing_row_ids = e.missing_row_ids.data();
      e.desc.missing_offsets = e.missing_offsets.data();
      e.desc.missing_runs = e.missing_runs.data();
      e.desc.n_null_rows = e.null_row_ids.size();
      e.desc.null_row_ids = e.null_row_ids.data();
      result = owned.release();
   });
   return result;
}
const silo_column_desc* orc_export_desc(void* exported) {
   return &static_cast<Export*>(exported)->desc;
}
void orc_export_free(void* exported) {
   delete static_cast<Export*>(exported);
}

// Builds a column from the upload format (whole table: first_chunk must be 0). The table's row
// layout must already be set (orc_table_set_layout) or match a previously imported column.
int orc_table_import_column(
   void* table,
   const char* name,
   int alphabet_id,
   const char* reference,
   const silo_column_desc* desc
) {
   return guarded([&] {
      auto* t = static_cast<Table*>(table);
      const Alphabet& alphabet = alphabet_id == 0 ? Alphabet::nucleotide() : Alphabet::aminoAcid();
      auto column = std::make_unique<SequenceColumn>(alphabet, name, reference);
      if (desc->genome_length != column->genomeLength() || desc->n_symbols != alphabet.count) {
         throw std::runtime_error("import: alphabet / genome length mismatch");
      }
      column->mutation_buffer.clear();
      for (uint32_t position = 0; position < desc->genome_length; ++position) {
         column->local_reference_sequence_string[position] =
            alphabet.symbolToChar(desc->local_reference[position]);
      }
      for (uint64_t i = 0; i < desc->n_containers; ++i) {
         const auto& d = desc->containers[i];
         column->vertical_sequence_index.vertical_bitmaps.emplace(
            SequenceDiffKey{d.position, d.v_index, d.symbol},
            Container::read(d.typecode, d.cardinality, desc->payload + d.payload_offset, d.payload_bytes)
         );
      }
      auto& coverage = column->horizontal_coverage_index;
      size_t row_cursor = 0;
      for (size_t chunk = 0; chunk < t->row_layout.numChunks(); ++chunk) {
         coverage.start_end.emplace_back();
         coverage.batch_start_ends.emplace_back(UINT32_MAX, 0);
         for (uint32_t row = 0; row < t->row_layout.chunk_sizes[chunk]; ++row) {
            const uint32_t start = desc->start_end[2 * row_cursor];
            const uint32_t end = desc->start_end[2 * row_cursor + 1];
            coverage.start_end.back().emplace_back(start, end);
            coverage.batch_start_ends.back().first = std::min(coverage.batch_start_ends.back().first, start);
            coverage.batch_start_ends.back().second = std::max(coverage.batch_start_ends.back().second, end);
            ++row_cursor;
         }
      }
      for (uint64_t i = 0; i < desc->n_rows_with_missing; ++i) {
         Roaring bitmap;
         for (uint64_t run = desc->missing_offsets[i]; run < desc->missing_offsets[i + 1]; ++run) {
            bitmap.addRange(desc->missing_runs[2 * run], desc->missing_runs[2 * run + 1]);
         }
         bitmap.runOptimize();
         coverage.horizontal_bitmaps.emplace(desc->missing_row_ids[i], std::move(bitmap));
      }
      column->null_bitmap = Roaring::fromIds(desc->null_row_ids, desc->n_null_rows);
      column->sequence_count = t->row_layout.numRows();
      column->num_chunks = static_cast<uint16_t>(t->row_layout.numChunks());
      t->columns.push_back(std::move(column));
   });
}

// ---- generators (performance/sequence_generator.h) ----

// Returns the number of evolved sequences; each written NUL-terminated back to back into `out`
// (capacity bytes) when out != NULL. parents (optional): index of each sequence's parent.
int64_t orc_gen_evolved(
   const char* reference,
   uint64_t seed,
   double mutation_rate,
   double death_rate,
   uint64_t generations,
   uint64_t children,
   char* out,
   uint64_t capacity,
   uint64_t* parents
) {
   int64_t count = -1;
   guarded([&] {
      const std::string ref(reference);
      SequenceTreeGenerator generator(ref, seed, mutation_rate, death_rate, generations, children);
      std::vector<size_t> parent_indices;
      const auto evolved = generator.generateEvolvedSequences(&parent_indices);
      count = static_cast<int64_t>(evolved.size());
      if (out != nullptr) {
         const uint64_t needed = evolved.size() * (ref.size() + 1);
         if (capacity < needed) {
            throw std::runtime_error("gen_evolved: buffer too small");
         }
         char* cursor = out;
         for (const auto& sequence : evolved) {
            std::memcpy(cursor, sequence.c_str(), sequence.size() + 1);
            cursor += sequence.size() + 1;
         }
      }
      if (parents != nullptr) {
         for (size_t i = 0; i < parent_indices.size(); ++i) {
            parents[i] = parent_indices[i];
         }
      }
   });
   return count;
}

// writeFullSequenceNdjson :367-384 — row i = evolved[i % |evolved|], single nucleotide column "main"
int orc_gen_full_sequence_table(void* table, const char* reference, uint64_t count, uint64_t generations) {
   return guarded([&] {
      auto* t = static_cast<Table*>(table);
      const std::string ref(reference);
      t->addColumn(Alphabet::nucleotide(), "main", ref);
      SequenceTreeGenerator generator(ref, 42, 0.001, 0.1, generations, 3);
      const auto evolved = generator.generateEvolvedSequences();
      for (uint64_t i = 0; i < count; ++i) {
         t->appendRow({std::make_pair(std::string_view(evolved[i % evolved.size()]), 0U)});
      }
      t->finalize();
   });
}

// writeNRunSequenceNdjson :392-428 (tree 12 generations, mt19937(7) N runs), `count` rows
int orc_gen_nrun_table(void* table, const char* reference, uint64_t count, uint64_t generations) {
   return guarded([&] {
      auto* t = static_cast<Table*>(table);
      const std::string ref(reference);
      t->addColumn(Alphabet::nucleotide(), "main", ref);
      SequenceTreeGenerator generator(ref, 42, 0.001, 0.1, generations, 3);
      auto evolved = generator.generateEvolvedSequences();
      std::mt19937 rng(7);
      std::uniform_int_distribution<size_t> pick(0, evolved.size() - 1);
      std::uniform_int_distribution<size_t> head_n(0, 300);
      std::uniform_int_distribution<size_t> tail_n(0, 300);
      std::uniform_int_distribution<size_t> internal_runs(0, 5);
      std::uniform_int_distribution<size_t> run_len(1, 100);
      std::uniform_int_distribution<size_t> pos_dist(0, ref.size() - 200);
      for (uint64_t i = 0; i < count; ++i) {
         std::string sequence = evolved.at(pick(rng));
         const size_t head = std::min(head_n(rng), sequence.size());
         const size_t tail = std::min(tail_n(rng), sequence.size());
         for (size_t j = 0; j < head; ++j) {
            sequence[j] = 'N';
         }
         for (size_t j = 0; j < tail; ++j) {
            sequence[sequence.size() - 1 - j] = 'N';
         }
         const size_t runs = internal_runs(rng);
         for (size_t r = 0; r < runs; ++r) {
            const size_t start = pos_dist(rng);
            const size_t length = run_len(rng);
            const size_t end = std::min(start + length, sequence.size());
            for (size_t j = start; j < end; ++j) {
               sequence[j] = 'N';
            }
         }
         t->appendRow({std::make_pair(std::string_view(sequence), 0U)});
      }
      t->finalize();
   });
}

// writeMutationBenchmarkNdjson :444-466, scaled by `batch_rows` (1000 in the reference)
int orc_gen_mutation_benchmark_table(void* table, uint64_t batch_rows) {
   return guarded([&] {
      auto* t = static_cast<Table*>(table);
      t->addColumn(Alphabet::nucleotide(), "main", buildMutationBenchmarkReference());
      const std::string read = "ACGT";
      auto emitBatches = [&](size_t batches, uint32_t offset) {
         for (size_t batch = 0; batch < batches; ++batch) {
            for (uint64_t i = 0; i < batch_rows; ++i) {
               t->appendRow({std::make_pair(std::string_view(read), offset)});
            }
         }
      };
      emitBatches(1000, 0);
      emitBatches(1000, 4);
      emitBatches(100, 99);
      for (uint32_t i = 0; i < 100; ++i) {
         emitBatches(1, 100 + i);
      }
      emitBatches(1000, 2000);
      t->finalize();
   });
}


// ---- container-level hook for the known-answer tests (roaring_container.test.cpp) ----
// Builds containers the way the reference's tests do (withCapacity(max(n,1)) + add each value, in the
// given order), optionally run-optimises them, optionally round-trips them through
// container_write/read, then applies op: 0 identity(a) 1 and 2 andnot 3 or 4 and_cardinality.
int orc_container_op(
   const uint16_t* a, uint32_t na, int optimize_a,
   const uint16_t* b, uint32_t nb, int optimize_b,
   int op, int roundtrip,
   uint16_t* out, uint32_t* n_out, uint8_t* type_a, uint8_t* type_out
) {
   return guarded([&] {
      auto make = [&](const uint16_t* values, uint32_t count, int optimize) {
         Container c = Container::withCapacity(static_cast<int32_t>(std::max<uint32_t>(count, 1)));
         for (uint32_t i = 0; i < count; ++i) {
            c.add(values[i]);
         }
         if (optimize != 0) {
            c.runOptimize();
         }
         if (roundtrip != 0) {
            std::vector<uint8_t> bytes(c.sizeInBytes());
            c.write(bytes.data());
            c = Container::read(c.type, c.card, bytes.data(), bytes.size());
         }
         return c;
      };
      const Container ca = make(a, na, optimize_a);
      const Container cb = make(b, nb, optimize_b);
      *type_a = ca.type;
      Container result;
      if (op == 4) {
         *n_out = containerAndCardinality(ca, cb);
         *type_out = 0;
         return;
      }
      result = op == 0 ? ca : op == 1 ? containerAnd(ca, cb) : op == 2 ? containerAndNot(ca, cb) : containerOr(ca, cb);
      uint32_t n = 0;
      result.forEach([&](uint16_t value) { out[n++] = value; });
      if (n != result.card) {
         throw std::runtime_error("container cardinality bookkeeping is inconsistent");
      }
      *n_out = n;
      *type_out = result.type;
   });
}

// roaring portable format round trip: ids -> Roaring (run-optimised if asked) -> write -> read -> ids
int64_t orc_roaring_roundtrip(const uint32_t* ids, uint64_t count, int optimize, uint32_t* out, uint8_t* bytes_out, uint64_t capacity) {
   int64_t size = -1;
   guarded([&] {
      Roaring bitmap = Roaring::fromIds(ids, count);
      if (optimize != 0) {
         bitmap.runOptimize();
      }
      const auto bytes = bitmap.write();
      const Roaring back = Roaring::read(bytes.data(), bytes.size());
      const auto values = back.toVector();
      std::copy(values.begin(), values.end(), out);
      if (bytes_out != nullptr && capacity >= bytes.size()) {
         std::memcpy(bytes_out, bytes.data(), bytes.size());
      }
      size = static_cast<int64_t>(bytes.size());
   });
   return size;
}

// aligned_sequence.cpp:20-122 hook: returns 0 and fills outputs, or -4 with the error message
int orc_extract(
   int alphabet_id, const char* sequence, uint32_t offset, const char* reference,
   uint32_t* start, uint32_t* end,
   uint32_t* missing, uint32_t* n_missing,
   uint32_t* mutation_positions, uint8_t* mutation_symbols, uint32_t* n_mutations
) {
   return guarded([&] {
      const Alphabet& alphabet = alphabet_id == 0 ? Alphabet::nucleotide() : Alphabet::aminoAcid();
      const std::string_view ref(reference);
      const bool reference_missing = ref.find(alphabet.symbolToChar(alphabet.missing)) != std::string_view::npos;
      std::string error;
      const auto result = extractCoverageAndMutationsFromSequence(alphabet, sequence, offset, ref, reference_missing, error);
      if (!result.has_value()) {
         throw AppendException(error);
      }
      *start = result->coverage.start;
      *end = result->coverage.end;
      *n_missing = static_cast<uint32_t>(result->coverage.missing_positions.size());
      std::copy(result->coverage.missing_positions.begin(), result->coverage.missing_positions.end(), missing);
      *n_mutations = static_cast<uint32_t>(result->mutations.size());
      for (size_t i = 0; i < result->mutations.size(); ++i) {
         mutation_positions[i] = result->mutations[i].first;
         mutation_symbols[i] = result->mutations[i].second;
      }
   });
}

// stored diff containers of one position, for the vertical-index tests:
// out rows: (v_index, symbol, typecode, cardinality); returns count
int64_t orc_column_containers_at(void* table, const char* column, uint32_t position, uint32_t* out, uint64_t capacity) {
   int64_t count = -1;
   guarded([&] {
      const auto* col = static_cast<Table*>(table)->findColumn(column);
      if (col == nullptr) {
         throw std::runtime_error("no such column");
      }
      auto [start, end] = col->vertical_sequence_index.getRangeForPosition(position);
      uint64_t n = 0;
      for (auto it = start; it != end; ++it) {
         if (n < capacity) {
            out[4 * n] = it->first.v_index;
            out[4 * n + 1] = it->first.symbol;
            out[4 * n + 2] = it->second.type;
            out[4 * n + 3] = it->second.card;
         }
         ++n;
      }
      count = static_cast<int64_t>(n);
   });
   return count;
}

// monotonic seconds, for the cpu_baseline leg
// ---- BitmapAggregationNode ----
// dimensions: ';'-separated, each "p:<column>:<0-based position>" (SequencePositionDimension) or
// "b:<value>=<bitmap name>,...|<null bitmap name or empty>" (IndexedColumnDimension over named bitmaps).
// Writes one line per combination: the values (\N = null) and the count, tab-separated. Returns the
// number of bytes the whole result needs (the caller retries with a larger buffer), < 0 on error.
int64_t orc_bitmap_aggregation(void* table, const char* expression, const char* dimensions, char* out, uint64_t capacity) {
   int64_t needed = -1;
   const int status = guarded([&] {
      const auto* t = static_cast<Table*>(table);
      const auto parsed = parseExpression(expression != nullptr ? expression : "(true)");
      std::vector<GroupingDimension> dims;
      const std::string spec = dimensions;
      size_t begin = 0;
      while (begin <= spec.size() && !spec.empty()) {
         size_t end = spec.find(';', begin);
         if (end == std::string::npos) {
            end = spec.size();
         }
         const std::string item = spec.substr(begin, end - begin);
         GroupingDimension dimension;
         if (item.rfind("p:", 0) == 0) {
            const size_t colon = item.rfind(':');
            dimension.is_sequence_position = true;
            dimension.column = item.substr(2, colon - 2);
            dimension.position_idx = static_cast<uint32_t>(std::stoul(item.substr(colon + 1)));
         } else if (item.rfind("b:", 0) == 0) {
            dimension.is_sequence_position = false;
            const size_t bar = item.rfind('|');
            dimension.null_bitmap = item.substr(bar + 1);
            const std::string groups = item.substr(2, bar - 2);
            size_t group_begin = 0;
            while (group_begin < groups.size()) {
               size_t group_end = groups.find(',', group_begin);
               if (group_end == std::string::npos) {
                  group_end = groups.size();
               }
               const std::string group = groups.substr(group_begin, group_end - group_begin);
               const size_t equals = group.find('=');
               dimension.value_bitmaps.emplace_back(group.substr(0, equals), group.substr(equals + 1));
               group_begin = group_end + 1;
            }
         } else {
            throw std::runtime_error("bad dimension spec: " + item);
         }
         dims.push_back(std::move(dimension));
         begin = end + 1;
      }
      std::string text;
      for (const Combination& combination : bitmapAggregation(*t, *parsed, dims)) {
         for (const auto& value : combination.values) {
            text += value.has_value() ? value.value() : std::string("\\N");
            text += '\t';
         }
         text += std::to_string(combination.count);
         text += '\n';
      }
      needed = static_cast<int64_t>(text.size());
      if (text.size() <= capacity) {
         std::memcpy(out, text.data(), text.size());
      }
   });
   return status == 0 ? needed : status;
}

double orc_now_seconds() {
   return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // extern "C"
