// C entry points of the host layer for the Python harness (include/silo_b200_host.h).
#include <algorithm>
#include <cstring>
#include <memory>
#include <sstream>
#include <string>

#include "../../include/silo_b200_host.h"
#include "expressions.h"
#include "bitmap_aggregation_node.h"
#include "mutations_node.h"
#include "operators.h"
#include "roaring_writer.h"
#include "silo_loader.h"
#include "synthetic.h"
#include "table.h"

using namespace silo_host;

struct silo_host_table {
   std::unique_ptr<Table> table;
};

struct silo_host_filter {
   DeviceBitmap bitmap;
   size_t n_chunks = 0;
};

struct silo_host_prepared {
   // keeps everything the uploaded program points into alive
   ExpressionPtr expression;
   ExpressionPtr rewritten;
   std::unique_ptr<Operator> compiled;
   silo_gpu_program* program = nullptr;
   silo_gpu_filter* filter = nullptr;
};

struct silo_host_rows {
   std::vector<MutationRow> rows;
   std::vector<std::string> names;  // distinct sequence names in order of first appearance
   std::vector<uint32_t> name_ids;
   void indexNames() {
      for (const MutationRow& row : rows) {
         uint32_t id = 0;
         while (id < names.size() && names[id] != row.sequence_name) {
            ++id;
         }
         if (id == names.size()) {
            names.push_back(row.sequence_name);
         }
         name_ids.push_back(id);
      }
   }
};

struct silo_host_archive {
   std::vector<std::unique_ptr<LoadedSequenceColumn>> columns;
   std::vector<std::unique_ptr<LoadedSequenceColumn>> shards;  // handed out by silo_host_archive_column_shard
};

struct silo_host_synthetic {
   std::string reference;
   const Alphabet* alphabet = &Alphabet::nucleotide();
   EvolvedTree tree;
   ShortReads reads;  // silo_host_synthetic_draw_short_reads
   std::unique_ptr<PackedColumn> column;
};

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
   } catch (const DeviceError& error) {
      g_last_error = "DeviceError[" + std::to_string(error.status) + "]: " + error.what();
      return -5;
   } catch (const std::exception& error) {
      g_last_error = error.what();
      return -1;
   }
}

const SequenceColumnInfo& columnOf(const Table& table, const char* name) {
   const SequenceColumnInfo* column = table.findColumn(name);
   if (column == nullptr) {
      throw IllegalQueryException(std::string("Database does not contain the Sequence with name: '") + name + "'");
   }
   return *column;
}

ExpressionPtr parseOrTrue(const char* expression) {
   if (expression == nullptr) {
      return std::make_shared<BoolLiteral>(true);
   }
   return parseFilterExpression(expression);
}

std::vector<ArchiveColumnSpec> archiveSpecs(const char* const* names, const int* alphabets, const char* const* references, uint32_t n_columns) {
   std::vector<ArchiveColumnSpec> specs(n_columns);
   for (uint32_t i = 0; i < n_columns; ++i) {
      specs[i].name = names[i];
      specs[i].alphabet = alphabets[i] == 0 ? &Alphabet::nucleotide() : &Alphabet::aminoAcid();
      specs[i].reference = references[i];
   }
   return specs;
}

int copyText(const std::string& text, char* out, uint64_t capacity) {
   if (out == nullptr || capacity < text.size() + 1) {
      g_last_error = "output buffer too small: need " + std::to_string(text.size() + 1) + " bytes";
      return -1;
   }
   std::memcpy(out, text.c_str(), text.size() + 1);
   return 0;
}
}  // namespace

extern "C" {

const char* silo_host_last_error(void) {
   return g_last_error.c_str();
}

silo_host_table* silo_host_table_create(silo_gpu_ctx* ctx, uint32_t first_chunk, const uint32_t* chunk_sizes, uint32_t n_chunks) {
   silo_host_table* result = nullptr;
   guarded([&] {
      RowLayout layout;
      layout.first_chunk = first_chunk;
      layout.chunk_sizes.assign(chunk_sizes, chunk_sizes + n_chunks);
      auto owned = std::make_unique<silo_host_table>();
      owned->table = std::make_unique<Table>(ctx, std::move(layout));
      result = owned.release();
   });
   return result;
}

void silo_host_table_free(silo_host_table* table) {
   delete table;
}

int silo_host_table_add_column(silo_host_table* table, const char* name, int alphabet, const char* reference, const silo_column_desc* column) {
   return guarded([&] {
      table->table->addSequenceColumn(
         name, alphabet == 0 ? Alphabet::nucleotide() : Alphabet::aminoAcid(), reference, *column
      );
   });
}

int silo_host_table_register_bitmap(silo_host_table* table, const char* name, const uint8_t* bytes, uint64_t size, int resident) {
   return guarded([&] { table->table->registerBitmap(name, bytes, size, resident != 0); });
}

silo_gpu_table* silo_host_table_device(silo_host_table* table) {
   return table->table->device;
}

uint64_t silo_host_table_num_rows(const silo_host_table* table) {
   return table->table->row_layout.numRows();
}

int silo_host_table_add_string_column(
   silo_host_table* table, const char* name, const char* const* dictionary, uint32_t n_values, const uint32_t* ids, const uint32_t* null_row_ids, uint64_t n_null_rows
) {
   return guarded([&] {
      std::vector<std::string> values(dictionary, dictionary + n_values);
      table->table->addStringColumn(name, values, ids, std::vector<uint32_t>(null_row_ids, null_row_ids + n_null_rows));
   });
}

int silo_host_table_add_date_column(silo_host_table* table, const char* name, const int32_t* days, const uint32_t* null_row_ids, uint64_t n_null_rows) {
   return guarded([&] { table->table->addDateColumn(name, days, std::vector<uint32_t>(null_row_ids, null_row_ids + n_null_rows)); });
}

silo_host_filter* silo_host_filter_eval(silo_host_table* table, const char* expression) {
   silo_host_filter* result = nullptr;
   guarded([&] {
      const ExpressionPtr parsed = parseFilterExpression(expression);
      auto owned = std::make_unique<silo_host_filter>();
      owned->bitmap = computeFilter(*parsed, *table->table);
      owned->n_chunks = table->table->row_layout.numChunks();
      result = owned.release();
   });
   return result;
}

int silo_host_count(silo_host_table* table, const char* expression, uint64_t* count) {
   return guarded([&] {
      const ExpressionPtr parsed = parseOrTrue(expression);
      *count = countFilter(*table->table, *parsed);
   });
}

void silo_host_filter_free(silo_host_filter* filter) {
   delete filter;
}

uint64_t silo_host_filter_cardinality(const silo_host_filter* filter) {
   return filter->bitmap.cardinality();
}

const silo_gpu_filter* silo_host_filter_device(const silo_host_filter* filter) {
   return filter->bitmap.get();
}

int silo_host_filter_words(const silo_host_filter* filter, uint64_t* words) {
   return guarded([&] {
      const std::vector<uint64_t> downloaded = filter->bitmap.toWords(filter->n_chunks);
      std::memcpy(words, downloaded.data(), downloaded.size() * sizeof(uint64_t));
   });
}

silo_host_prepared* silo_host_filter_prepare(silo_host_table* table, const char* expression) {
   silo_host_prepared* result = nullptr;
   guarded([&] {
      auto owned = std::make_unique<silo_host_prepared>();
      owned->expression = parseFilterExpression(expression);
      owned->rewritten = owned->expression->rewrite(*table->table, AmbiguityMode::NONE);
      owned->compiled = owned->rewritten->compile(*table->table);
      ProgramBuilder builder;
      builder.table = table->table.get();
      owned->compiled->lower(builder);
      silo_filter_program program{};
      program.struct_size = sizeof(silo_filter_program);
      program.n_instrs = static_cast<uint32_t>(builder.instrs.size());
      program.instrs = builder.instrs.data();
      program.blob = builder.blob.data();
      program.blob_bytes = builder.blob.size();
      program.n_bitmaps = static_cast<uint32_t>(builder.bitmaps.size());
      program.bitmaps = builder.bitmaps.data();
      throwOnDeviceError(silo_gpu_program_prepare(table->table->deviceTable(), &program, &owned->program, &owned->filter));
      result = owned.release();
   });
   return result;
}

int silo_host_prepared_run_async(silo_host_prepared* prepared, void* cuda_stream) {
   return guarded([&] { throwOnDeviceError(silo_gpu_program_run_async(prepared->program, cuda_stream)); });
}

int silo_host_prepared_run_counts_async(silo_host_prepared* prepared, int column_index, void* d_counts, void* cuda_stream) {
   return guarded([&] { throwOnDeviceError(silo_gpu_program_run_counts_async(prepared->program, column_index, d_counts, cuda_stream)); });
}

int silo_host_prepared_run_sharded_async(silo_host_prepared* prepared, void* cuda_stream) {
   return guarded([&] { throwOnDeviceError(silo_gpu_program_run_sharded_async(prepared->program, cuda_stream)); });
}

int silo_host_prepared_run_sharded_collect_async(silo_host_prepared* prepared, void* d_summed_counts, void* cuda_stream) {
   return guarded([&] { throwOnDeviceError(silo_gpu_program_run_sharded_collect_async(prepared->program, d_summed_counts, cuda_stream)); });
}

int silo_host_sharded_collect_async(silo_host_table* table, void* d_summed_counts, void* cuda_stream) {
   return guarded([&] { throwOnDeviceError(silo_gpu_sharded_collect_async(table->table->deviceTable(), d_summed_counts, cuda_stream)); });
}

const silo_gpu_filter* silo_host_prepared_filter(const silo_host_prepared* prepared) {
   return prepared->filter;
}

uint64_t silo_host_prepared_staged_bytes(const silo_host_prepared* prepared) {
   return silo_gpu_program_device_bytes(prepared->program);
}

void silo_host_prepared_free(silo_host_prepared* prepared) {
   if (prepared == nullptr) {
      return;
   }
   silo_gpu_program_free(prepared->program);
   silo_gpu_filter_free(prepared->filter);
   delete prepared;
}

int silo_host_filter_explain(silo_host_table* table, const char* expression, char* out, uint64_t capacity) {
   std::string text;
   const int status = guarded([&] {
      const ExpressionPtr parsed = parseFilterExpression(expression);
      const ExpressionPtr rewritten = parsed->rewrite(*table->table, AmbiguityMode::NONE);
      const std::unique_ptr<Operator> compiled = rewritten->compile(*table->table);
      ProgramBuilder builder;
      builder.table = table->table.get();
      compiled->lower(builder);
      std::ostringstream stream;
      stream << "operator: " << compiled->toString() << "\n";
      for (const silo_filter_instr& instr : builder.instrs) {
         stream << "op=" << static_cast<int>(instr.opcode) << " flags=" << static_cast<int>(instr.flags)
                << " column=" << instr.column << " a=" << instr.a << " b=" << instr.b << "\n";
      }
      text = stream.str();
   });
   if (status != 0) {
      return status;
   }
   return copyText(text, out, capacity);
}

int silo_host_filter_to_string(silo_host_table* table, const char* expression, char* out, uint64_t capacity) {
   std::string text;
   const int status = guarded([&] {
      const ExpressionPtr parsed = parseFilterExpression(expression);
      const ExpressionPtr rewritten = parsed->rewrite(*table->table, AmbiguityMode::NONE);
      const std::unique_ptr<Operator> compiled = rewritten->compile(*table->table);
      text = parsed->toString() + "\n" + rewritten->toString() + "\n" + compiled->toString() + "\n";
   });
   if (status != 0) {
      return status;
   }
   return copyText(text, out, capacity);
}

int64_t silo_host_filter_program_bitmap(silo_host_table* table, const char* expression, uint32_t index, uint8_t* out, uint64_t capacity) {
   int64_t size = -1;
   guarded([&] {
      const ExpressionPtr parsed = parseFilterExpression(expression);
      const ExpressionPtr rewritten = parsed->rewrite(*table->table, AmbiguityMode::NONE);
      const std::unique_ptr<Operator> compiled = rewritten->compile(*table->table);
      ProgramBuilder builder;
      builder.table = table->table.get();
      compiled->lower(builder);
      const silo_roaring_bytes& bitmap = builder.bitmaps.at(index);
      if (bitmap.size > capacity) {
         throw std::invalid_argument("bitmap buffer too small: need " + std::to_string(bitmap.size));
      }
      std::memcpy(out, bitmap.data, bitmap.size);
      size = static_cast<int64_t>(bitmap.size);
   });
   return size;
}

int silo_host_filter_lower_timed(silo_host_table* table, const char* expression, double phase_us[4], uint64_t sizes[3], uint64_t* digest) {
   return guarded([&] {
      const double t0 = nowMicroseconds();
      const ExpressionPtr parsed = parseFilterExpression(expression);
      const double t1 = nowMicroseconds();
      const ExpressionPtr rewritten = parsed->rewrite(*table->table, AmbiguityMode::NONE);
      const double t2 = nowMicroseconds();
      const std::unique_ptr<Operator> compiled = rewritten->compile(*table->table);
      const double t3 = nowMicroseconds();
      ProgramBuilder builder;
      builder.table = table->table.get();
      compiled->lower(builder);
      const double t4 = nowMicroseconds();
      phase_us[0] = t1 - t0;
      phase_us[1] = t2 - t1;
      phase_us[2] = t3 - t2;
      phase_us[3] = t4 - t3;
      sizes[0] = builder.instrs.size();
      sizes[1] = builder.blob.size();
      sizes[2] = builder.bitmaps.size();
      uint64_t hash = 14695981039346656037ULL;
      auto mix = [&](uint64_t value, int bytes) {
         for (int i = 0; i < bytes; ++i) {
            hash = (hash ^ ((value >> (8 * i)) & 0xFF)) * 1099511628211ULL;
         }
      };
      for (const silo_filter_instr& instr : builder.instrs) {
         mix(instr.opcode, 1);
         mix(instr.flags, 1);
         mix(instr.column, 2);
         mix(instr.a, 4);
         mix(instr.b, 8);
      }
      for (uint8_t byte : builder.blob) {
         mix(byte, 1);
      }
      *digest = hash;
   });
}

int silo_host_mutation_counts(silo_host_table* table, const char* column, const silo_host_filter* filter, uint32_t* counts) {
   return guarded([&] {
      const Table& t = *table->table;
      const SequenceColumnInfo& info = columnOf(t, column);
      SymbolCounts result;
      if (filter == nullptr) {
         const DeviceBitmap everything = computeFilter(BoolLiteral(true), t);
         result = calculateMutationsPerPosition(t, info, everything, t.row_layout.numRows());
      } else {
         result = calculateMutationsPerPosition(t, info, filter->bitmap, t.row_layout.numRows());
      }
      std::memcpy(counts, result.values, result.size() * sizeof(uint32_t));
   });
}

silo_host_rows* silo_host_mutations(silo_host_table* table, const char* expression, const char* const* columns, uint32_t n_columns, double min_proportion) {
   silo_host_rows* result = nullptr;
   guarded([&] {
      std::vector<std::string> names(columns, columns + n_columns);
      const double parse_begin = nowMicroseconds();
      ExpressionPtr parsed = parseOrTrue(expression);
      lastQueryProfile().parse_us = nowMicroseconds() - parse_begin;
      const MutationsNode node(*table->table, std::move(parsed), std::move(names), min_proportion);
      auto owned = std::make_unique<silo_host_rows>();
      owned->rows = node.execute();
      owned->indexNames();
      result = owned.release();
   });
   return result;
}

namespace {

uint64_t alignEight(uint64_t value) {
   return (value + 7) / 8 * 8;
}

uint64_t packedBytes(const silo_host_rows& rows) {
   const uint64_t n = rows.rows.size();
   uint64_t bytes = alignEight(8 * n) + 4 * alignEight(4 * n) + 2 * alignEight(n);
   for (const std::string& name : rows.names) {
      bytes += name.size() + 1;
   }
   return bytes;
}

void packRows(const silo_host_rows& rows, uint8_t* out) {
   const uint64_t n = rows.rows.size();
   auto* proportion = reinterpret_cast<double*>(out);
   auto* position = reinterpret_cast<int32_t*>(out + alignEight(8 * n));
   auto* name_ids = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(position) + alignEight(4 * n));
   auto* count = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(name_ids) + alignEight(4 * n));
   auto* coverage = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(count) + alignEight(4 * n));
   char* from = reinterpret_cast<char*>(coverage) + alignEight(4 * n);
   char* to = from + alignEight(n);
   for (uint64_t i = 0; i < n; ++i) {
      const MutationRow& row = rows.rows[i];
      proportion[i] = row.proportion;
      position[i] = row.position;
      name_ids[i] = rows.name_ids[i];
      count[i] = row.count;
      coverage[i] = row.coverage;
      from[i] = row.mutation_from;
      to[i] = row.mutation_to;
   }
   char* names = to + alignEight(n);
   for (const std::string& name : rows.names) {
      std::memcpy(names, name.c_str(), name.size() + 1);
      names += name.size() + 1;
   }
}

thread_local std::unique_ptr<silo_host_rows> g_pending_rows;  // a packed result that did not fit the caller's buffer

}  // namespace

int silo_host_mutations_packed(
   silo_host_table* table,
   const char* expression,
   const char* const* columns,
   uint32_t n_columns,
   double min_proportion,
   void* buffer,
   uint64_t capacity,
   uint64_t* n_rows,
   uint32_t* n_names,
   uint64_t* needed_bytes
) {
   return guarded([&] {
      g_pending_rows.reset();
      std::vector<std::string> names(columns, columns + n_columns);
      const double parse_begin = nowMicroseconds();
      ExpressionPtr parsed = parseOrTrue(expression);
      lastQueryProfile().parse_us = nowMicroseconds() - parse_begin;
      const MutationsNode node(*table->table, std::move(parsed), std::move(names), min_proportion);
      auto owned = std::make_unique<silo_host_rows>();
      owned->rows = node.execute();
      owned->indexNames();
      *n_rows = owned->rows.size();
      *n_names = static_cast<uint32_t>(owned->names.size());
      *needed_bytes = packedBytes(*owned);
      if (*needed_bytes <= capacity) {
         packRows(*owned, static_cast<uint8_t*>(buffer));
      } else {
         g_pending_rows = std::move(owned);
      }
   });
}

int silo_host_mutations_enqueue(silo_host_table* table, const char* expression, const char* column, void* d_counts, void* cuda_stream) {
   return guarded([&] {
      const double parse_begin = nowMicroseconds();
      ExpressionPtr parsed = parseOrTrue(expression);
      lastQueryProfile().parse_us = nowMicroseconds() - parse_begin;
      const MutationsNode node(*table->table, std::move(parsed), {std::string(column)}, 0.0);
      node.enqueueShardCounts(d_counts, cuda_stream);
   });
}

int silo_host_mutations_collect_packed(
   silo_host_table* table,
   const char* column,
   double min_proportion,
   const void* d_summed_counts,
   void* cuda_stream,
   void* buffer,
   uint64_t capacity,
   uint64_t* n_rows,
   uint32_t* n_names,
   uint64_t* needed_bytes,
   uint64_t* shard_cardinality
) {
   return guarded([&] {
      g_pending_rows.reset();
      const MutationsNode node(*table->table, nullptr, {std::string(column)}, min_proportion);
      auto owned = std::make_unique<silo_host_rows>();
      owned->rows = node.collectRows(d_summed_counts, cuda_stream, shard_cardinality);
      owned->indexNames();
      *n_rows = owned->rows.size();
      *n_names = static_cast<uint32_t>(owned->names.size());
      *needed_bytes = packedBytes(*owned);
      if (*needed_bytes <= capacity) {
         packRows(*owned, static_cast<uint8_t*>(buffer));
      } else {
         g_pending_rows = std::move(owned);
      }
   });
}

int silo_host_shard_group_create(silo_host_table* table, const char* column, int rank, int world, void* handle_out) {
   return guarded([&] {
      const std::vector<uint8_t> handle = createShardGroup(*table->table, column, rank, world);
      std::memcpy(handle_out, handle.data(), handle.size());
   });
}

int silo_host_shard_group_connect(silo_host_table* table, const void* handles, uint64_t handles_bytes) {
   return guarded([&] {
      const auto* bytes = static_cast<const uint8_t*>(handles);
      connectShardGroup(*table->table, std::vector<uint8_t>(bytes, bytes + handles_bytes));
   });
}

int silo_host_sharded_enqueue(silo_host_table* table, const char* expression, const char* column, void* cuda_stream) {
   return guarded([&] {
      const MutationsNode node(*table->table, parseOrTrue(expression), {std::string(column)}, 0.0);
      node.enqueueSharded(cuda_stream);
   });
}

int silo_host_sharded_query_packed(
   silo_host_table* table,
   const char* expression,
   const char* column,
   double min_proportion,
   void* d_summed_counts,
   void* buffer,
   uint64_t capacity,
   uint64_t* n_rows,
   uint32_t* n_names,
   uint64_t* needed_bytes,
   uint64_t* cardinality
) {
   return guarded([&] {
      g_pending_rows.reset();
      const MutationsNode node(*table->table, parseOrTrue(expression), {std::string(column)}, min_proportion);
      auto owned = std::make_unique<silo_host_rows>();
      owned->rows = node.executeShardedRoot(d_summed_counts, cardinality);
      owned->indexNames();
      *n_rows = owned->rows.size();
      *n_names = static_cast<uint32_t>(owned->names.size());
      *needed_bytes = packedBytes(*owned);
      if (*needed_bytes <= capacity) {
         packRows(*owned, static_cast<uint8_t*>(buffer));
      } else {
         g_pending_rows = std::move(owned);
      }
   });
}

int silo_host_sharded_collect_packed(
   silo_host_table* table,
   const char* column,
   double min_proportion,
   void* d_summed_counts,
   void* cuda_stream,
   void* buffer,
   uint64_t capacity,
   uint64_t* n_rows,
   uint32_t* n_names,
   uint64_t* needed_bytes,
   uint64_t* cardinality
) {
   return guarded([&] {
      g_pending_rows.reset();
      const MutationsNode node(*table->table, nullptr, {std::string(column)}, min_proportion);
      auto owned = std::make_unique<silo_host_rows>();
      owned->rows = node.collectSharded(d_summed_counts, cuda_stream, cardinality);
      owned->indexNames();
      *n_rows = owned->rows.size();
      *n_names = static_cast<uint32_t>(owned->names.size());
      *needed_bytes = packedBytes(*owned);
      if (*needed_bytes <= capacity) {
         packRows(*owned, static_cast<uint8_t*>(buffer));
      } else {
         g_pending_rows = std::move(owned);
      }
   });
}

int silo_host_packed_fetch(void* buffer, uint64_t capacity) {
   return guarded([&] {
      if (g_pending_rows == nullptr || packedBytes(*g_pending_rows) > capacity) {
         throw std::invalid_argument("silo_host_packed_fetch: no pending result of that size on this thread");
      }
      packRows(*g_pending_rows, static_cast<uint8_t*>(buffer));
      g_pending_rows.reset();
   });
}

silo_host_rows* silo_host_mutation_rows_from_counts(silo_host_table* table, const char* column, const uint32_t* counts, double min_proportion) {
   silo_host_rows* result = nullptr;
   guarded([&] {
      const SequenceColumnInfo& info = columnOf(*table->table, column);
      SymbolCounts symbol_counts;
      symbol_counts.n_symbols = info.alphabet->count();
      symbol_counts.genome_length = static_cast<uint32_t>(info.reference_sequence.size());
      symbol_counts.values = counts;  // a view: the caller's (all-reduced) buffer is only read
      auto owned = std::make_unique<silo_host_rows>();
      appendMutationRows(info, symbol_counts, min_proportion, owned->rows);
      owned->indexNames();
      result = owned.release();
   });
   return result;
}

// dimensions: ';'-separated, each "p:<column>:<0-based position>" (SequencePositionDimension) or
// "b:<value>=<bitmap name>,...|<null bitmap name or empty>" (IndexedColumnDimension over named bitmaps).
// out: one line per combination, the values (\N = null) and the count, tab-separated.
static std::vector<GroupingDimension> parseDimensionSpec(const char* dimensions) {
   std::vector<GroupingDimension> dims;
   const std::string spec = dimensions;
   size_t begin = 0;
   while (!spec.empty() && begin <= spec.size()) {
      size_t end = spec.find(';', begin);
      if (end == std::string::npos) {
         end = spec.size();
      }
      const std::string item = spec.substr(begin, end - begin);
      if (item.rfind("p:", 0) == 0) {
         const size_t colon = item.rfind(':');
         SequencePositionDimension dimension;
         dimension.column = item.substr(2, colon - 2);
         dimension.position_idx = static_cast<uint32_t>(std::stoul(item.substr(colon + 1)));
         dims.emplace_back(std::move(dimension));
      } else if (item.rfind("b:", 0) == 0) {
         IndexedColumnDimension dimension;
         const size_t bar = item.rfind('|');
         if (bar + 1 < item.size()) {
            dimension.null_bitmap = item.substr(bar + 1);
         }
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
         dims.emplace_back(std::move(dimension));
      } else {
         throw std::invalid_argument("bad dimension spec: " + item);
      }
      begin = end + 1;
   }
   return dims;
}

static std::string combinationRowsText(const std::vector<CombinationRow>& rows) {
   std::string text;
   for (const CombinationRow& row : rows) {
      for (const auto& value : row.values) {
         text += value.has_value() ? value.value() : std::string("\\N");
         text += '\t';
      }
      text += std::to_string(row.count);
      text += '\n';
   }
   return text;
}

int silo_host_bitmap_aggregation(silo_host_table* table, const char* expression, const char* dimensions, char* out, uint64_t capacity) {
   std::string text;
   const int status = guarded([&] {
      const BitmapAggregationNode node(*table->table, parseOrTrue(expression), parseDimensionSpec(dimensions));
      text = combinationRowsText(node.execute());
   });
   if (status != 0) {
      return status;
   }
   return copyText(text, out, capacity);
}

// the rows as arrays instead of text: codes_out[row * n_dims + d] = the symbol character of a sequence-position dimension
// / the index of the value in the dimension's sorted value list of an indexed dimension; 0 / 255 = the null group
static void packCombinationRows(
   const std::vector<GroupingDimension>& dims, const std::vector<CombinationRow>& rows, uint8_t* codes_out, uint64_t* counts_out, uint64_t capacity_rows,
   uint64_t* n_rows
) {
   std::vector<std::vector<std::string>> sorted_values(dims.size());
   for (size_t d = 0; d < dims.size(); ++d) {
      if (const auto* indexed = std::get_if<IndexedColumnDimension>(&dims[d])) {
         for (const auto& [value, bitmap_name] : indexed->value_bitmaps) {
            sorted_values[d].push_back(value);
         }
         std::sort(sorted_values[d].begin(), sorted_values[d].end());
      }
   }
   *n_rows = rows.size();
   for (uint64_t r = 0; r < rows.size() && r < capacity_rows; ++r) {
      for (size_t d = 0; d < dims.size(); ++d) {
         const auto& value = rows[r].values[d];
         uint8_t code;
         if (std::holds_alternative<SequencePositionDimension>(dims[d])) {
            code = value.has_value() ? static_cast<uint8_t>(value.value()[0]) : 0;
         } else {
            code = 255;
            if (value.has_value()) {
               const auto found = std::lower_bound(sorted_values[d].begin(), sorted_values[d].end(), value.value());
               code = static_cast<uint8_t>(found - sorted_values[d].begin());
            }
         }
         codes_out[r * dims.size() + d] = code;
      }
      counts_out[r] = static_cast<uint64_t>(rows[r].count);
   }
}

int silo_host_bitmap_aggregation_packed(
   silo_host_table* table, const char* expression, const char* dimensions, uint8_t* codes_out, uint64_t* counts_out, uint64_t capacity_rows, uint64_t* n_rows
) {
   return guarded([&] {
      const std::vector<GroupingDimension> dims = parseDimensionSpec(dimensions);
      const BitmapAggregationNode node(*table->table, parseOrTrue(expression), dims);
      packCombinationRows(dims, node.execute(), codes_out, counts_out, capacity_rows, n_rows);
   });
}

int silo_host_bitmap_aggregation_merge_packed(
   silo_host_table* table, const char* dimensions, const uint64_t* pairs, const uint64_t* entries_per_shard, const uint64_t* cardinalities, uint32_t n_shards,
   uint8_t* codes_out, uint64_t* counts_out, uint64_t capacity_rows, uint64_t* n_rows
) {
   return guarded([&] {
      const std::vector<GroupingDimension> dims = parseDimensionSpec(dimensions);
      const BitmapAggregationNode node(*table->table, parseOrTrue(nullptr), dims);
      std::vector<BitmapAggregationNode::ShardCombinations> shards(n_shards);
      uint64_t at = 0;
      for (uint32_t shard = 0; shard < n_shards; ++shard) {
         shards[shard].cardinality = cardinalities[shard];
         shards[shard].entries.resize(entries_per_shard[shard]);
         for (uint64_t i = 0; i < entries_per_shard[shard]; ++i, ++at) {
            shards[shard].entries[i] = silo_combination{pairs[2 * at], pairs[2 * at + 1]};
         }
      }
      packCombinationRows(dims, node.materialise(BitmapAggregationNode::mergeShards(shards)), codes_out, counts_out, capacity_rows, n_rows);
   });
}

int silo_host_bitmap_aggregation_shard(
   silo_host_table* table, const char* expression, const char* dimensions, uint64_t* pairs_out, uint64_t capacity_entries, uint64_t* n_entries, uint64_t* cardinality
) {
   return guarded([&] {
      const BitmapAggregationNode node(*table->table, parseOrTrue(expression), parseDimensionSpec(dimensions));
      const BitmapAggregationNode::ShardCombinations shard = node.executeShard();
      *n_entries = shard.entries.size();
      *cardinality = shard.cardinality;
      for (uint64_t i = 0; i < shard.entries.size() && i < capacity_entries; ++i) {
         pairs_out[2 * i] = shard.entries[i].key;
         pairs_out[2 * i + 1] = shard.entries[i].count;
      }
   });
}

int silo_host_bitmap_aggregation_merge(
   silo_host_table* table, const char* dimensions, const uint64_t* pairs, const uint64_t* entries_per_shard, const uint64_t* cardinalities, uint32_t n_shards,
   char* out, uint64_t capacity
) {
   std::string text;
   const int status = guarded([&] {
      const BitmapAggregationNode node(*table->table, parseOrTrue(nullptr), parseDimensionSpec(dimensions));
      std::vector<BitmapAggregationNode::ShardCombinations> shards(n_shards);
      uint64_t at = 0;
      for (uint32_t shard = 0; shard < n_shards; ++shard) {
         shards[shard].cardinality = cardinalities[shard];
         shards[shard].entries.resize(entries_per_shard[shard]);
         for (uint64_t i = 0; i < entries_per_shard[shard]; ++i, ++at) {
            shards[shard].entries[i] = silo_combination{pairs[2 * at], pairs[2 * at + 1]};
         }
      }
      text = combinationRowsText(node.materialise(BitmapAggregationNode::mergeShards(shards)));
   });
   if (status != 0) {
      return status;
   }
   return copyText(text, out, capacity);
}

void silo_host_last_query_profile(double* out) {
   const QueryProfile& profile = lastQueryProfile();
   out[0] = profile.parse_us;
   out[1] = profile.compile_us;
   out[2] = profile.filter_us;
   out[3] = profile.counts_us;
   out[4] = profile.threshold_us;
}

void silo_host_rows_free(silo_host_rows* rows) {
   delete rows;
}

uint64_t silo_host_rows_size(const silo_host_rows* rows) {
   return rows->rows.size();
}

int silo_host_rows_export(const silo_host_rows* rows, char* from, char* to, int32_t* position, uint32_t* name_ids, double* proportion, int32_t* count, int32_t* coverage) {
   return guarded([&] {
      for (size_t i = 0; i < rows->rows.size(); ++i) {
         const MutationRow& row = rows->rows[i];
         from[i] = row.mutation_from;
         to[i] = row.mutation_to;
         position[i] = row.position;
         name_ids[i] = rows->name_ids[i];
         proportion[i] = row.proportion;
         count[i] = row.count;
         coverage[i] = row.coverage;
      }
   });
}

uint32_t silo_host_rows_num_names(const silo_host_rows* rows) {
   return static_cast<uint32_t>(rows->names.size());
}

const char* silo_host_rows_name(const silo_host_rows* rows, uint32_t name_id) {
   return name_id < rows->names.size() ? rows->names[name_id].c_str() : nullptr;
}

int silo_host_rows_get(const silo_host_rows* rows, uint64_t index, char* from, char* to, int32_t* position, const char** sequence_name, double* proportion, int32_t* count, int32_t* coverage) {
   return guarded([&] {
      const MutationRow& row = rows->rows.at(index);
      *from = row.mutation_from;
      *to = row.mutation_to;
      *position = row.position;
      *sequence_name = row.sequence_name.c_str();
      *proportion = row.proportion;
      *count = row.count;
      *coverage = row.coverage;
   });
}

// ---- synthetic ---------------------------------------------------------------------------------

silo_host_synthetic* silo_host_synthetic_create(uint32_t genome_length, uint64_t reference_seed, uint32_t generations) {
   silo_host_synthetic* result = nullptr;
   guarded([&] {
      auto owned = std::make_unique<silo_host_synthetic>();
      owned->reference = randomNucleotideReference(genome_length, reference_seed);
      owned->tree = generateEvolvedSequences(owned->reference, 42, 0.001, 0.1, generations, 3);
      result = owned.release();
   });
   return result;
}

silo_host_synthetic* silo_host_synthetic_create_gene(uint32_t gene_length, uint64_t reference_seed, uint64_t tree_seed, double mutation_rate, uint32_t generations) {
   silo_host_synthetic* result = nullptr;
   guarded([&] {
      auto owned = std::make_unique<silo_host_synthetic>();
      owned->alphabet = &Alphabet::aminoAcid();
      owned->reference = randomAminoAcidReference(gene_length, reference_seed);
      std::string valid;
      for (const Symbol symbol : owned->alphabet->valid_mutation_symbols) {
         valid.push_back(owned->alphabet->symbolToChar(symbol));
      }
      owned->tree = generateEvolvedSequences(owned->reference, tree_seed, mutation_rate, 0.1, generations, 3, valid);
      result = owned.release();
   });
   return result;
}

silo_host_synthetic* silo_host_synthetic_create_co_occurrence(uint64_t n_sequences) {
   silo_host_synthetic* result = nullptr;
   guarded([&] {
      auto owned = std::make_unique<silo_host_synthetic>();
      owned->reference = coOccurrenceReference();
      owned->tree.sequences = coOccurrenceSequences(owned->reference, n_sequences);
      owned->tree.parent.assign(n_sequences, 0);
      owned->tree.generation.assign(n_sequences, 0);
      result = owned.release();
   });
   return result;
}

void silo_host_synthetic_free(silo_host_synthetic* synthetic) {
   delete synthetic;
}

uint32_t silo_host_synthetic_num_sequences(const silo_host_synthetic* synthetic) {
   return static_cast<uint32_t>(synthetic->tree.sequences.size());
}
const char* silo_host_synthetic_reference(const silo_host_synthetic* synthetic) {
   return synthetic->reference.c_str();
}
const char* silo_host_synthetic_sequence(const silo_host_synthetic* synthetic, uint32_t index) {
   return index < synthetic->tree.sequences.size() ? synthetic->tree.sequences[index].c_str() : nullptr;
}
uint32_t silo_host_synthetic_parent(const silo_host_synthetic* synthetic, uint32_t index) {
   return synthetic->tree.parent.at(index);
}
uint32_t silo_host_synthetic_generation(const silo_host_synthetic* synthetic, uint32_t index) {
   return synthetic->tree.generation.at(index);
}

int silo_host_synthetic_build_column(silo_host_synthetic* synthetic, uint64_t total_rows, uint32_t first_chunk, uint32_t n_chunks, uint32_t threads, const silo_column_desc** out, uint32_t chunk_stride) {
   return guarded([&] {
      synthetic->column = std::make_unique<PackedColumn>();
      buildCycledColumn(
         *synthetic->alphabet, synthetic->reference, synthetic->tree.sequences, total_rows, first_chunk, n_chunks,
         threads, *synthetic->column, chunk_stride
      );
      *out = &synthetic->column->desc;
   });
}

int silo_host_synthetic_draw_short_reads(silo_host_synthetic* synthetic, uint64_t count, uint32_t read_length, uint32_t* sequence_of_read_out) {
   return guarded([&] {
      synthetic->reads = drawShortReads(synthetic->tree.sequences.size(), count, read_length);
      if (sequence_of_read_out != nullptr) {
         std::memcpy(sequence_of_read_out, synthetic->reads.sequence_of_read.data(), count * sizeof(uint32_t));
      }
   });
}

int silo_host_synthetic_build_short_read_column(silo_host_synthetic* synthetic, uint32_t first_chunk, uint32_t n_chunks, uint32_t threads, const silo_column_desc** out) {
   return guarded([&] {
      synthetic->column = std::make_unique<PackedColumn>();
      buildShortReadColumn(*synthetic->alphabet, synthetic->reference, synthetic->tree.sequences, synthetic->reads, first_chunk, n_chunks, threads, *synthetic->column);
      *out = &synthetic->column->desc;
   });
}

void silo_host_synthetic_release_column(silo_host_synthetic* synthetic) {
   synthetic->column.reset();
}

int64_t silo_host_synthetic_lineage_bitmap(const silo_host_synthetic* synthetic, uint32_t ancestor, uint64_t total_rows, uint32_t first_chunk, uint32_t n_chunks, uint8_t* out, uint64_t capacity, uint32_t chunk_stride) {
   int64_t size = -1;
   guarded([&] {
      const std::vector<uint8_t> bytes =
         writePortableRoaring(lineageRowIds(synthetic->tree, ancestor, total_rows, first_chunk, n_chunks, chunk_stride));
      size = static_cast<int64_t>(bytes.size());
      if (out != nullptr && capacity >= bytes.size()) {
         std::memcpy(out, bytes.data(), bytes.size());
      }
   });
   return size;
}

int silo_host_synthetic_date_ranges(uint64_t total_rows, uint32_t span_days, uint32_t from_day, uint32_t to_day_inclusive, uint32_t first_chunk, uint32_t n_chunks, char* out, uint64_t capacity, uint32_t chunk_stride) {
   std::string text;
   const int status = guarded([&] {
      const std::vector<uint32_t> flat = sortedDateRanges(total_rows, span_days, from_day, to_day_inclusive, first_chunk, n_chunks, chunk_stride);
      std::ostringstream stream;
      stream << "(ranges";
      for (uint32_t value : flat) {
         stream << ' ' << value;
      }
      stream << ')';
      text = stream.str();
   });
   if (status != 0) {
      return status;
   }
   return copyText(text, out, capacity);
}

int silo_host_partition_chunks(const uint64_t* chunk_weights, uint32_t n_chunks, uint32_t n_ranks, uint32_t* boundaries) {
   return guarded([&] {
      if (n_ranks == 0) {
         throw std::invalid_argument("n_ranks must be positive");
      }
      const std::vector<uint32_t> result =
         partitionChunks(std::vector<uint64_t>(chunk_weights, chunk_weights + n_chunks), n_ranks);
      std::copy(result.begin(), result.end(), boundaries);
   });
}

silo_host_archive* silo_host_archive_read(const uint8_t* bytes, uint64_t size, const char* const* names, const int* alphabets,
                                          const char* const* references, uint32_t n_columns) {
   silo_host_archive* result = nullptr;
   guarded([&] {
      auto owned = std::make_unique<silo_host_archive>();
      owned->columns = readSequenceColumns(bytes, size, archiveSpecs(names, alphabets, references, n_columns));
      result = owned.release();
   });
   return result;
}

void silo_host_archive_free(silo_host_archive* archive) {
   delete archive;
}

const silo_column_desc* silo_host_archive_column(const silo_host_archive* archive, uint32_t index) {
   return index < archive->columns.size() ? &archive->columns[index]->desc : nullptr;
}

int silo_host_archive_column_info(const silo_host_archive* archive, uint32_t index, uint64_t info[6]) {
   return guarded([&] {
      const LoadedSequenceColumn& column = *archive->columns.at(index);
      info[0] = column.chunk_sizes.size();
      info[1] = column.sequence_count;
      info[2] = column.n_insertion_positions;
      info[3] = column.vertical_bitmaps_size;
      info[4] = column.horizontal_bitmaps_size;
      info[5] = column.num_chunks;
   });
}

int silo_host_archive_chunk_sizes(const silo_host_archive* archive, uint32_t index, uint32_t* chunk_sizes, uint32_t capacity) {
   return guarded([&] {
      const LoadedSequenceColumn& column = *archive->columns.at(index);
      if (capacity < column.chunk_sizes.size()) {
         throw std::invalid_argument("chunk_sizes buffer too small");
      }
      std::copy(column.chunk_sizes.begin(), column.chunk_sizes.end(), chunk_sizes);
   });
}

const silo_column_desc* silo_host_archive_column_shard(silo_host_archive* archive, uint32_t index, uint32_t first_chunk, uint32_t n_chunks) {
   const silo_column_desc* result = nullptr;
   guarded([&] {
      archive->shards.push_back(shardOf(*archive->columns.at(index), first_chunk, n_chunks));
      result = &archive->shards.back()->desc;
   });
   return result;
}

silo_host_table* silo_host_table_load_archive(silo_gpu_ctx* ctx, const uint8_t* bytes, uint64_t size, const char* const* names,
                                              const int* alphabets, const char* const* references, uint32_t n_columns,
                                              uint32_t first_chunk, uint32_t n_chunks) {
   silo_host_table* result = nullptr;
   guarded([&] {
      auto owned = std::make_unique<silo_host_table>();
      owned->table = loadTableFromArchive(ctx, bytes, size, archiveSpecs(names, alphabets, references, n_columns), {}, first_chunk, n_chunks);
      result = owned.release();
   });
   return result;
}

int64_t silo_host_roaring_runs(const uint8_t* bytes, uint64_t size, uint32_t* runs, uint64_t capacity_runs) {
   int64_t n_runs = -1;
   guarded([&] {
      std::vector<uint32_t> decoded;
      portableRoaringToRuns(bytes, size, decoded);
      if (decoded.size() / 2 > capacity_runs) {
         throw std::invalid_argument("runs buffer too small: need " + std::to_string(decoded.size() / 2));
      }
      std::copy(decoded.begin(), decoded.end(), runs);
      n_runs = static_cast<int64_t>(decoded.size() / 2);
   });
   return n_runs;
}

}  // extern "C"
