/*
 * silo_b200_host.h — C entry points of the HOST layer (libsilo_b200_host.so), for the Python
 * harness (tests/, bench.py) only.
 *
 * The host layer is the C++ code that, in an integrated build, lives inside the reference's query
 * engine: it mirrors the reference's ScalarExpression -> Operator compilation
 * (/root/reference/src/rhydb/query_engine/scalar_expressions/, filter/operators/) and the
 * MutationsNode / CountFilterNode sinks (query_engine/operators/), and talks to the device only
 * through include/silo_b200.h. A maintainer of the reference would not use THIS header: it exists
 * because the reference's front-end (SaneQL, Planner, Arrow sinks) cannot be built here, so tests
 * and bench.py need some way to drive the host layer. See INTEGRATION.md.
 */
#ifndef SILO_B200_HOST_H
#define SILO_B200_HOST_H

#include <stdint.h>

#include "silo_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct silo_host_table silo_host_table;
typedef struct silo_host_filter silo_host_filter;
typedef struct silo_host_rows silo_host_rows;

/* thread-local message of the last failed silo_host_* call; prefixed with the exception class
 * ("IllegalQueryException: ", "QueryCompilationException: ", "DeviceError[<status>]: ") */
const char* silo_host_last_error(void);

/* ctx == NULL creates a HOST-ONLY table: columns keep their metadata only (nothing is uploaded), the front half
 * of the query compiler works (silo_host_filter_explain, silo_host_filter_lower_timed) and every call that
 * needs the device fails with DeviceError[SILO_E_NO_DEVICE] -- there is no CPU evaluation path. */
silo_host_table* silo_host_table_create(silo_gpu_ctx* ctx, uint32_t first_chunk, const uint32_t* chunk_sizes, uint32_t n_chunks);
void silo_host_table_free(silo_host_table* table);
/* alphabet: 0 nucleotide, 1 amino acid. Uploads through silo_gpu_column_upload. */
int silo_host_table_add_column(silo_host_table* table, const char* name, int alphabet, const char* reference, const silo_column_desc* column);
/* ready-made roaring bitmap (portable format) standing in for a lineage / dictionary index */
/* resident != 0: made device resident once (silo_gpu_bitmap_register), programs refer to it by id;
 * resident == 0: the bytes travel with every program that uses it (PUSH_BITMAP). */
int silo_host_table_register_bitmap(silo_host_table* table, const char* name, const uint8_t* bytes, uint64_t size, int resident);
silo_gpu_table* silo_host_table_device(silo_host_table* table);
uint64_t silo_host_table_num_rows(const silo_host_table* table);

/* computeFilter: parse (harness s-expression notation) -> rewrite(NONE) -> compile -> evaluate */
silo_host_filter* silo_host_filter_eval(silo_host_table* table, const char* expression);
void silo_host_filter_free(silo_host_filter* filter);
uint64_t silo_host_filter_cardinality(const silo_host_filter* filter);
const silo_gpu_filter* silo_host_filter_device(const silo_host_filter* filter);
int silo_host_filter_words(const silo_host_filter* filter, uint64_t* words /* 1024 x n_chunks */);
/* the lowered program as text, one instruction per line (debugging / tests of the lowering) */
int silo_host_filter_explain(silo_host_table* table, const char* expression, char* out, uint64_t capacity);

/* toString() of the expression as parsed, after rewrite(NONE) and of the compiled operator tree, one per line
 * (the reference's own formats: and.cpp:32, or.cpp:27, nof.cpp:161, symbol_in_set.cpp:37, threshold.cpp:45 ...) */
int silo_host_filter_to_string(silo_host_table* table, const char* expression, char* out, uint64_t capacity);
/* the bytes (portable roaring) of bitmap `index` travelling with the lowered program of `expression`; returns the
 * size, or a negative status (buffer too small included) */
int64_t silo_host_filter_program_bitmap(silo_host_table* table, const char* expression, uint32_t index, uint8_t* out, uint64_t capacity);
/* parse -> rewrite(NONE) -> compile -> lower without running anything: microseconds of the four phases,
 * sizes[3] = {instructions, blob bytes, bitmaps travelling with the program}, and a 64-bit FNV-1a digest of the
 * lowered program (instruction fields + blob) so that two lowerings can be compared without the text form */
int silo_host_filter_lower_timed(silo_host_table* table, const char* expression, double phase_us[4], uint64_t sizes[3], uint64_t* digest);

/* BitmapAggregationNode (operators/bitmap_aggregation_node.cpp:304-356) through the host layer.
 * dimensions: ';'-separated, each "p:<column>:<0-based position>" (SequencePositionDimension) or
 * "b:<value>=<bitmap name>,...|<null bitmap name or empty>" (IndexedColumnDimension; the inverted index
 * of the dictionary column arrives as named bitmaps). out: one line per combination in the reference's
 * depth-first order: the values (\\N = null) and the count, tab-separated. */
int silo_host_bitmap_aggregation(silo_host_table* table, const char* expression, const char* dimensions, char* out, uint64_t capacity);
/* The same rows as arrays (no text to format and parse): codes_out[row * n_dims + d] = the symbol character of a
 * sequence-position dimension, or the index of the value in the dimension's sorted value list; 0 resp. 255 = null.
 * *n_rows may exceed capacity_rows (call again with room). _merge_packed: the collecting rank's half, see below. */
int silo_host_bitmap_aggregation_packed(
   silo_host_table* table, const char* expression, const char* dimensions, uint8_t* codes_out, uint64_t* counts_out, uint64_t capacity_rows, uint64_t* n_rows
);
int silo_host_bitmap_aggregation_merge_packed(
   silo_host_table* table, const char* dimensions, const uint64_t* pairs, const uint64_t* entries_per_shard, const uint64_t* cardinalities, uint32_t n_shards,
   uint8_t* codes_out, uint64_t* counts_out, uint64_t capacity_rows, uint64_t* n_rows
);
/* Row-partitioned tables: BitmapAggregationNode::executeShard on every rank -- pairs_out receives (key, count) x
 * min(*n_entries, capacity_entries), ordered by key --, then ::mergeShards + ::materialise on one rank over the ranks'
 * lists laid end to end (entries_per_shard[s] pairs of shard s, its filter cardinality in cardinalities[s]); the text
 * is that of silo_host_bitmap_aggregation on the whole table. The same `dimensions` on every rank. */
int silo_host_bitmap_aggregation_shard(
   silo_host_table* table, const char* expression, const char* dimensions, uint64_t* pairs_out, uint64_t capacity_entries, uint64_t* n_entries, uint64_t* cardinality
);
int silo_host_bitmap_aggregation_merge(
   silo_host_table* table, const char* dimensions, const uint64_t* pairs, const uint64_t* entries_per_shard, const uint64_t* cardinalities, uint32_t n_shards,
   char* out, uint64_t capacity
);

/* the same filter with device-resident inputs: compile + lower + upload once (silo_gpu_program_prepare),
 * then each run only enqueues the kernel on `cuda_stream` (silo_gpu_program_run_async) */
typedef struct silo_host_prepared silo_host_prepared;
silo_host_prepared* silo_host_filter_prepare(silo_host_table* table, const char* expression);
int silo_host_prepared_run_async(silo_host_prepared* prepared, void* cuda_stream);
/* filter + Mutations counts of one column (index in add order) into d_counts, enqueued only (silo_gpu_program_run_counts_async) */
int silo_host_prepared_run_counts_async(silo_host_prepared* prepared, int column_index, void* d_counts, void* cuda_stream);
const silo_gpu_filter* silo_host_prepared_filter(const silo_host_prepared* prepared);
uint64_t silo_host_prepared_staged_bytes(const silo_host_prepared* prepared);
void silo_host_prepared_free(silo_host_prepared* prepared);

/* calculateMutationsPerPosition incl. the full / mixed / empty dispatch (mutations_node.cpp:279-286);
 * filter == NULL means the filter `true`. counts[n_symbols * genome_length]. */
int silo_host_mutation_counts(silo_host_table* table, const char* column, const silo_host_filter* filter, uint32_t* counts);
/* MutationsNode: filter expression (NULL = true) + columns + minProportion -> output rows */
silo_host_rows* silo_host_mutations(silo_host_table* table, const char* expression, const char* const* columns, uint32_t n_columns, double min_proportion);
/* The same query with the result as ONE record batch in a caller buffer (what the reference hands to
 * its Arrow sink, mutations_node.cpp:404-428), so that a binding needs a single call per query. Layout
 * for n rows, every section starting at a multiple of 8 bytes:
 *   f64 proportion[n] | i32 position[n] | u32 sequence_name_id[n] | i32 count[n] | i32 coverage[n] |
 *   char mutation_from[n] | char mutation_to[n] | the distinct sequence names, each NUL-terminated
 * *n_rows / *n_names / *needed_bytes are always set; when needed_bytes > capacity nothing is written
 * and the result is kept for silo_host_packed_fetch (same thread, before the next query). */
int silo_host_mutations_packed(silo_host_table* table, const char* expression, const char* const* columns, uint32_t n_columns, double min_proportion,
                               void* buffer, uint64_t capacity, uint64_t* n_rows, uint32_t* n_names, uint64_t* needed_bytes);
int silo_host_packed_fetch(void* buffer, uint64_t capacity);
/* The Mutations query on a row-partitioned table in two halves (MutationsNode::enqueueShardCounts /
 * collectRows): every rank parses + compiles the query against its shard and enqueues program upload, filter
 * and counts on `cuda_stream` (d_counts: n_symbols * genome_length u32 in device memory), the scheduler
 * all-reduces d_counts on that stream, then ONE rank collects the rows (record batch as above) from the
 * summed counts. The caller synchronises nothing in between. */
int silo_host_mutations_enqueue(silo_host_table* table, const char* expression, const char* column, void* d_counts, void* cuda_stream);
int silo_host_mutations_collect_packed(silo_host_table* table, const char* column, double min_proportion, const void* d_summed_counts, void* cuda_stream,
                                       void* buffer, uint64_t capacity, uint64_t* n_rows, uint32_t* n_names, uint64_t* needed_bytes,
                                       uint64_t* shard_cardinality);
/* CountFilterNode: `filter(...).groupBy({count:=count()})` as one device call (silo_gpu_query_count) */
int silo_host_count(silo_host_table* table, const char* expression, uint64_t* count);
/* Metadata columns for the Selection predicates (string equality on an unindexed string column, equals.cpp:124-156;
 * DateBetween, date_between.cpp:61-134): a string column as its dictionary + one id per row (layout order), a Date32
 * column as its day numbers; null_row_ids ascending global row ids (may be NULL with n_null_rows 0). The values become
 * device resident (silo_gpu_value_column_upload); harness notation: (str-eq COLUMN VALUE), (date-between COLUMN FROM TO)
 * with day numbers or * for an open end. */
int silo_host_table_add_string_column(silo_host_table* table, const char* name, const char* const* dictionary, uint32_t n_values,
                                      const uint32_t* ids, const uint32_t* null_row_ids, uint64_t n_null_rows);
int silo_host_table_add_date_column(silo_host_table* table, const char* name, const int32_t* days, const uint32_t* null_row_ids, uint64_t n_null_rows);
/* The same query through the table's shard group (MutationsNode::enqueueSharded / collectSharded; silo_gpu_shard_group_*
 * of include/silo_b200.h): _create on every rank writes this rank's handle (SILO_SHARD_HANDLE_BYTES), _connect takes the
 * handles of all ranks in rank order; then every rank enqueues, rank 0 collects (rows of the whole table; cardinality =
 * rows of all shards that passed the filter). d_summed_counts may be NULL. */
int silo_host_shard_group_create(silo_host_table* table, const char* column, int rank, int world, void* handle_out);
int silo_host_shard_group_connect(silo_host_table* table, const void* handles, uint64_t handles_bytes);
int silo_host_sharded_enqueue(silo_host_table* table, const char* expression, const char* column, void* cuda_stream);
int silo_host_sharded_collect_packed(silo_host_table* table, const char* column, double min_proportion, void* d_summed_counts, void* cuda_stream,
                                     void* buffer, uint64_t capacity, uint64_t* n_rows, uint32_t* n_names, uint64_t* needed_bytes,
                                     uint64_t* cardinality);
/* rank 0: enqueue + collect as one device call (MutationsNode::executeShardedRoot) */
int silo_host_sharded_query_packed(silo_host_table* table, const char* expression, const char* column, double min_proportion, void* d_summed_counts,
                                   void* buffer, uint64_t capacity, uint64_t* n_rows, uint32_t* n_names, uint64_t* needed_bytes, uint64_t* cardinality);
/* device-resident pipeline of the same: a prepared program on every rank, the root's collect without output pass */
int silo_host_prepared_run_sharded_async(silo_host_prepared* prepared, void* cuda_stream);
/* the root: the same with the collect inside the finalize kernel (silo_gpu_program_run_sharded_collect_async) */
int silo_host_prepared_run_sharded_collect_async(silo_host_prepared* prepared, void* d_summed_counts, void* cuda_stream);
int silo_host_sharded_collect_async(silo_host_table* table, void* d_summed_counts, void* cuda_stream);
/* thresholding only, on counts the caller summed over shards (multi-GPU) */
silo_host_rows* silo_host_mutation_rows_from_counts(silo_host_table* table, const char* column, const uint32_t* counts, double min_proportion);
/* microseconds of the calling thread's last silo_host_mutations call:
 * {parse, rewrite+compile+lower, filter_eval, mutation_counts, thresholding} */
void silo_host_last_query_profile(double out[5]);
void silo_host_rows_free(silo_host_rows* rows);
uint64_t silo_host_rows_size(const silo_host_rows* rows);
/* all rows at once (struct-of-arrays, each array silo_host_rows_size long); name_ids index
 * silo_host_rows_name */
int silo_host_rows_export(const silo_host_rows* rows, char* from, char* to, int32_t* position, uint32_t* name_ids, double* proportion, int32_t* count, int32_t* coverage);
uint32_t silo_host_rows_num_names(const silo_host_rows* rows);
const char* silo_host_rows_name(const silo_host_rows* rows, uint32_t name_id);
int silo_host_rows_get(const silo_host_rows* rows, uint64_t index, char* from, char* to, int32_t* position, const char** sequence_name, double* proportion, int32_t* count, int32_t* coverage);

/* ---- `.silo` -> device loader (host/silo_loader.h): the sequence columns of a table file written by
 * Table::serializeData (storage/table.h:35-42), read straight into the S1 upload format */

typedef struct silo_host_archive silo_host_archive;
/* columns in the archive's order (nucleotide columns by name, then amino-acid columns by name);
 * alphabets: 0 nucleotide, 1 amino acid; references: the global reference genomes. Host only, no device call. */
silo_host_archive* silo_host_archive_read(const uint8_t* bytes, uint64_t size, const char* const* names, const int* alphabets,
                                          const char* const* references, uint32_t n_columns);
void silo_host_archive_free(silo_host_archive* archive);
/* the column in the upload format; valid until the archive object is freed */
const silo_column_desc* silo_host_archive_column(const silo_host_archive* archive, uint32_t index);
/* info[6] = {n_chunks, sequence_count, positions with insertions (the insertion index is read through, not kept),
 * vertical_bitmaps_size, horizontal_bitmaps_size, num_chunks member} */
int silo_host_archive_column_info(const silo_host_archive* archive, uint32_t index, uint64_t info[6]);
int silo_host_archive_chunk_sizes(const silo_host_archive* archive, uint32_t index, uint32_t* chunk_sizes, uint32_t capacity);
/* the chunks [first_chunk, first_chunk + n_chunks) of the column (one rank's part of the row-partitioned table;
 * v_index and row ids stay global); valid until the archive object is freed */
const silo_column_desc* silo_host_archive_column_shard(silo_host_archive* archive, uint32_t index, uint32_t first_chunk, uint32_t n_chunks);
/* S1 for a saved database: creates the table (row layout from the coverage index) and uploads every column;
 * a rank of the row-partitioned table passes its chunk range (n_chunks = UINT32_MAX: up to the last chunk).
 * ctx == NULL gives a host-only table (metadata, no upload). */
silo_host_table* silo_host_table_load_archive(silo_gpu_ctx* ctx, const uint8_t* bytes, uint64_t size, const char* const* names,
                                              const int* alphabets, const char* const* references, uint32_t n_columns,
                                              uint32_t first_chunk, uint32_t n_chunks);
/* ascending {first, end_exclusive} runs of a portable roaring bitmap; returns the number of runs or -1 */
int64_t silo_host_roaring_runs(const uint8_t* bytes, uint64_t size, uint32_t* runs, uint64_t capacity_runs);

/* ---- synthetic benchmark inputs (performance/sequence_generator.h restated on the product side) */

typedef struct silo_host_synthetic silo_host_synthetic;
/* Evolution tree over a seeded random reference of `genome_length` nt (tree seed 42, mutation rate
 * 0.001, death rate 0.1, 3 children: SequenceTreeGenerator defaults) */
silo_host_synthetic* silo_host_synthetic_create(uint32_t genome_length, uint64_t reference_seed, uint32_t generations);
/* The short-read table of performance/sequence_generator.h:189-325 (uniform tiling): _draw makes the per-read sequence
 * draws (std::mt19937(42 + 1000), returned through sequence_of_read_out[count] when not NULL; read i starts at
 * i * (L - read_length + 1) / count), _build_short_read_column packs the chunks [first_chunk, first_chunk + n_chunks) of
 * the table of `count` reads in the S1 upload format (valid until silo_host_synthetic_release_column / the next build). */
int silo_host_synthetic_draw_short_reads(silo_host_synthetic* synthetic, uint64_t count, uint32_t read_length, uint32_t* sequence_of_read_out);
int silo_host_synthetic_build_short_read_column(silo_host_synthetic* synthetic, uint32_t first_chunk, uint32_t n_chunks, uint32_t threads, const silo_column_desc** out);
/* The same model for one amino-acid gene (SURVEY.md 8(d) input 4): random reference over the twenty standard residues,
 * mutations drawn from the alphabet's valid mutation symbols, its own tree seed and mutation rate. */
silo_host_synthetic* silo_host_synthetic_create_gene(uint32_t gene_length, uint64_t reference_seed, uint64_t tree_seed, double mutation_rate, uint32_t generations);
/* the table of performance/co_occurrence_benchmark.cpp: a 100-nt random reference, n_sequences rows with
 * Binomial(100, 0.1) point substitutions each (sequence_generator.h:487-526); build_column cycles over them */
silo_host_synthetic* silo_host_synthetic_create_co_occurrence(uint64_t n_sequences);
void silo_host_synthetic_free(silo_host_synthetic* synthetic);
uint32_t silo_host_synthetic_num_sequences(const silo_host_synthetic* synthetic);
const char* silo_host_synthetic_reference(const silo_host_synthetic* synthetic);
const char* silo_host_synthetic_sequence(const silo_host_synthetic* synthetic, uint32_t index);
uint32_t silo_host_synthetic_parent(const silo_host_synthetic* synthetic, uint32_t index);
uint32_t silo_host_synthetic_generation(const silo_host_synthetic* synthetic, uint32_t index);
/* Builds the shard [first_chunk, first_chunk + n_chunks) of the table "row i = sequence[i % E]"
 * with total_rows rows directly in the upload format; *out stays valid until the synthetic object
 * is freed or the next build. */
int silo_host_synthetic_build_column(silo_host_synthetic* synthetic, uint64_t total_rows, uint32_t first_chunk, uint32_t n_chunks, uint32_t threads, const silo_column_desc** out, uint32_t chunk_stride);
/* releases the host copy of the last built column (after it was uploaded) */
void silo_host_synthetic_release_column(silo_host_synthetic* synthetic);
/* lineage stand-in: portable roaring bytes of the shard's rows descending from `ancestor` */
int64_t silo_host_synthetic_lineage_bitmap(const silo_host_synthetic* synthetic, uint32_t ancestor, uint64_t total_rows, uint32_t first_chunk, uint32_t n_chunks, uint8_t* out, uint64_t capacity, uint32_t chunk_stride);
/* date stand-in: the "(ranges ...)" expression text of DateBetween on the sorted synthetic date column */
int silo_host_synthetic_date_ranges(uint64_t total_rows, uint32_t span_days, uint32_t from_day, uint32_t to_day_inclusive, uint32_t first_chunk, uint32_t n_chunks, char* out, uint64_t capacity, uint32_t chunk_stride);
/* partition scheduler: boundaries[n_ranks + 1] of contiguous chunk ranges balanced by weight */
int silo_host_partition_chunks(const uint64_t* chunk_weights, uint32_t n_chunks, uint32_t n_ranks, uint32_t* boundaries);

#ifdef __cplusplus
}
#endif
#endif /* SILO_B200_HOST_H */
