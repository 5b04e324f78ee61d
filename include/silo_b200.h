/*
 * silo_b200.h — C ABI of libsilo_b200.so: the B200 (sm_100a) drop-in for the bitmap filter +
 * Mutations hot path of RhyDB/SILO (GenSpectrum/LAPIS-SILO v0.13.3).
 *
 * The reference has no FFI seam on this path; the boundary replaces the BODIES of three call sites
 * (paths relative to /root/reference/src/rhydb), leaving every signature above them untouched:
 *
 *   S1 upload  storage/table.cpp:87-94 (Table::finalize) and :227-235 (Table::loadData):
 *              walk SequenceColumn::{vertical_sequence_index.vertical_bitmaps,
 *              horizontal_coverage_index, null_bitmap, local_reference_sequence_string}
 *              (storage/column/sequence_column.h:104-114) + Table::row_layout
 *              -> silo_gpu_table_create / silo_gpu_column_upload
 *   S2 filter  query_engine/operators/compute_filter.cpp:14-21 (and database.cpp:275-278):
 *              after rewrite()+compile(), the Operator tree (filter/operators/operator.h:11-39) is
 *              lowered to a flat filter program -> silo_gpu_filter_eval
 *   S3 action  query_engine/operators/mutations_node.cpp:268-288 (calculateMutationsPerPosition)
 *              -> silo_gpu_mutation_counts; addMutationsToOutput (:290-366) stays on the host.
 *              query_engine/operators/count_filter_node.cpp:35-71 -> silo_gpu_filter_cardinality
 *
 * Plain C: opaque handles, plain pointers and sizes, caller-owned outputs. Every function returns
 * 0 on success or a negative silo_status; silo_gpu_last_error() returns the thread-local message.
 * There is NO CPU fallback: without a CUDA device every entry point fails with SILO_E_NO_DEVICE.
 *
 * Row ids are the reference's sparse global ids (storage/column/row_id.h:16-37):
 *   id = (chunk_id << 16) | row_in_chunk, chunk k owns [k<<16, (k<<16)+chunk_size(k)).
 * A table handle may hold a SHARD of the chunks (multi-GPU: one process per GPU, contiguous chunk
 * ranges); `first_chunk` is the global chunk id of the shard's first chunk.
 */
#ifndef SILO_B200_H
#define SILO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
   SILO_OK = 0,
   SILO_E_INVALID_ARGUMENT = -1,
   SILO_E_NO_DEVICE = -2,
   SILO_E_CUDA = -3,
   SILO_E_OUT_OF_MEMORY = -4,
   SILO_E_BAD_PROGRAM = -5,
   SILO_E_OUT_OF_LAYOUT = -6, /* A leaf bitmap holds ids outside the row layout (row_layout.h:49-52 states the
                                 precondition the reference relies on). silo_gpu_filter_eval keeps such ids the way the
                                 reference's bitmaps do -- And / Or carry them, Not flips inside the layout only
                                 (row_layout.cpp:18-23), a Threshold counts them when a child holds them
                                 (threshold.test.cpp:249-311) --, provided they fall into a chunk of the table (ids in
                                 chunks the table does not have are dropped). Every entry that reads per-row data by the
                                 filter's rows (the Mutations counts, the fused queries, the aggregation) refuses such a
                                 filter with this status instead of reading rows that do not exist. */
   SILO_E_UNSUPPORTED = -7
} silo_status;

typedef struct silo_gpu_ctx silo_gpu_ctx;       /* one per process+device: streams, scratch pools */
typedef struct silo_gpu_table silo_gpu_table;   /* row layout of one table (shard) + its columns */
typedef struct silo_gpu_filter silo_gpu_filter; /* device-resident dense filter result */

const char* silo_gpu_last_error(void);
/* "libsilo_b200 <version> sm_100a" */
const char* silo_gpu_version(void);

int silo_gpu_init(int device_ordinal, silo_gpu_ctx** out);
void silo_gpu_shutdown(silo_gpu_ctx* ctx);

/* Page-locked host memory for caller-allocated outputs. silo_gpu_mutation_counts and
 * silo_gpu_filter_download accept any host pointer; into a buffer obtained here the DMA engine
 * writes directly (no staging copy). A long-lived caller keeps a few of these and reuses them. */
void* silo_gpu_host_alloc(silo_gpu_ctx* ctx, uint64_t bytes);
void silo_gpu_host_free(silo_gpu_ctx* ctx, void* ptr);

/* ---- S1: upload ------------------------------------------------------------------------------ */

/* One stored diff container: key of VerticalSequenceIndex::SequenceDiffKey
 * (storage/column/vertical_sequence_index.h:22-39) + the RoaringContainer bookkeeping
 * (roaring_util/roaring_container.h:24-26). payload = exactly what
 * roaring::internal::container_write emits (roaring_container.h:104-116):
 *   typecode 1 bitset: 1024 x u64 LE | 2 array: cardinality x u16 ascending
 *   | 3 run: u16 n_runs, then n_runs x {u16 start, u16 length_minus_1} */
typedef struct {
   uint32_t position;
   uint16_t v_index; /* GLOBAL chunk id */
   uint8_t symbol;
   uint8_t typecode;
   uint32_t cardinality;
   uint32_t payload_bytes;
   uint64_t payload_offset; /* into silo_column_desc.payload */
} silo_container_desc;

typedef struct {
   uint32_t struct_size; /* sizeof(silo_column_desc), for ABI evolution */
   uint32_t n_symbols;   /* 16 nucleotide (nucleotide_symbols.h:42), 28 amino acid (aa_symbols.h:57) */
   uint32_t genome_length;
   uint32_t missing_symbol;        /* Nucleotide::N = 15, AminoAcid::X = 27 */
   const uint8_t* local_reference; /* [genome_length] symbol ids, sequence_column.h:104 */
   /* vertical index, in std::map order (position, v_index, symbol); only chunks of this shard */
   uint64_t n_containers;
   const silo_container_desc* containers;
   const uint8_t* payload;
   uint64_t payload_bytes;
   /* horizontal coverage index (horizontal_coverage_index.h:25-35) */
   const uint32_t* start_end; /* {start,end} per row, rows of the shard's chunks back to back */
   uint64_t n_rows_with_missing;     /* horizontal_bitmaps.size() */
   const uint32_t* missing_row_ids;  /* ascending global row ids */
   const uint64_t* missing_offsets;  /* [n_rows_with_missing + 1], in runs */
   const uint32_t* missing_runs;     /* {first, end_exclusive} runs of the row's N positions */
   uint64_t n_null_rows; /* null_bitmap, sequence_column.h:111 */
   const uint32_t* null_row_ids;
} silo_column_desc;

int silo_gpu_table_create(
   silo_gpu_ctx* ctx,
   uint32_t first_chunk,
   const uint32_t* chunk_sizes, /* RowLayout::chunk_sizes of the shard, row_layout.h:28 */
   uint32_t n_chunks,
   silo_gpu_table** out
);
void silo_gpu_table_free(silo_gpu_table* table);
/* Copies everything; the caller keeps ownership of the inputs. Returns the column index (>= 0). */
int silo_gpu_column_upload(silo_gpu_table* table, const silo_column_desc* column);
/* bytes of HBM held by the table's pools */
uint64_t silo_gpu_table_device_bytes(const silo_gpu_table* table);
/* Tuning knobs of a table handle (no reference counterpart; defaults suit production):
 *   "sweep_min_pieces"  a THR_PROFILE instruction over a column with at least this many container pieces is counted
 *                       by the whole-column sweep kernel in front of the interpreter (default 65536; 0 = always,
 *                       UINT64_MAX = never: the interpreter walks the containers chunk by chunk). */
int silo_gpu_table_set_option(silo_gpu_table* table, const char* name, uint64_t value);

/* ---- S2: filter program ---------------------------------------------------------------------- */

/* Stack machine evaluated independently per chunk over dense 64 Ki-row tiles (1024 x u64).
 * Leaves push a tile, operators pop/push. Mapping from the reference's operators:
 *   Empty/Full (empty.cpp:25, full.cpp:26)            PUSH_EMPTY / PUSH_FULL
 *   IndexScan over vertical-index views (symbol_in_set.cpp:216-228)   PUSH_SYMBOLS
 *   IndexScan over a foreign roaring (lineage_filter.cpp:96-99, null_bitmap) PUSH_BITMAP / PUSH_INDEX_BITMAP / PUSH_NULLS
 *   Selection(IsInCoveredRegion) (is_in_covered_region.cpp:53-62)     PUSH_COVERED (flag: negated)
 *   RangeSelection (range_selection.cpp:54-87)        PUSH_RANGES
 *   Intersection (intersection.cpp:59-105)            chain of AND / ANDNOT
 *   Union (union.cpp:34-42)                           chain of OR
 *   Complement (complement.cpp:51-56)                 NOT  (flip inside [0, chunk_size) only)
 *   Threshold (threshold.cpp:64-138)                  THR_BEGIN .. THR_ADD* .. THR_END */
typedef enum {
   SILO_OP_PUSH_EMPTY = 1,
   SILO_OP_PUSH_FULL = 2,
   SILO_OP_PUSH_SYMBOLS = 3, /* column, a = position, b = symbol bit mask */
   SILO_OP_PUSH_COVERED = 4, /* column, a = position, flags&1: NOT covered (within the layout) */
   SILO_OP_PUSH_NULLS = 5,   /* column */
   SILO_OP_PUSH_BITMAP = 6,  /* a = index into silo_filter_program.bitmaps */
   SILO_OP_PUSH_RANGES = 7,  /* a = number of ranges, b = byte offset (a multiple of 8) into blob of {u32 start,u32 end} */
   SILO_OP_PUSH_INDEX_BITMAP = 8, /* a = id returned by silo_gpu_bitmap_register */
   SILO_OP_PUSH_COMPARE = 9, /* A Selection predicate over a value column (silo_gpu_value_column_upload), evaluated for
                                every row of the layout: column = value column index; flags[2:0] = comparator
                                (silo_comparator); flags bit 3: compare as signed 32-bit (Date32, integers), else
                                unsigned (dictionary ids); flags bit 4: null rows match (with_nulls of
                                CompareToValueSelection, selection.h:76-166). a = the value; BETWEEN: a <= v <= (u32) b;
                                IN_SET: a = number of values, b = blob offset of the ascending u32 values. */
   SILO_OP_AND = 16,         /* pop y, pop x, push x & y */
   SILO_OP_ANDNOT = 17,      /* pop y, pop x, push x & ~y */
   SILO_OP_OR = 18,          /* pop y, pop x, push x | y */
   SILO_OP_NOT = 19,         /* top ^= layout mask */
   SILO_OP_THR_BEGIN = 32,   /* a = number_of_matchers, flags&1: match_exactly */
   SILO_OP_THR_ADD = 33,     /* pop child tile; flags&1: negated child (counts rows NOT in it) */
   SILO_OP_THR_ADD_SYMBOLS = 34, /* column, a = position, b = mask; flags&1: subtract instead */
   SILO_OP_THR_ADD_COVERED = 35, /* column, a = n positions, b = blob offset of sorted u32 positions:
                                    +1 per listed position the row covers */
   SILO_OP_THR_PROFILE = 36, /* column, b = blob offset of u32[2*genome_length]: per position
                                {add_mask, sub_mask}; every stored container (pos,sym) adds +1 to its
                                rows if sym in add_mask, -1 if in sub_mask. One streaming pass. */
   SILO_OP_THR_END = 37      /* push (count >= k) or (count == k), restricted to the layout */
} silo_filter_opcode;

typedef enum {
   SILO_CMP_EQUALS = 0,
   SILO_CMP_NOT_EQUALS = 1,
   SILO_CMP_LESS = 2,
   SILO_CMP_LESS_OR_EQUALS = 3,
   SILO_CMP_HIGHER = 4,
   SILO_CMP_HIGHER_OR_EQUALS = 5,
   SILO_CMP_BETWEEN = 6, /* inclusive on both ends */
   SILO_CMP_IN_SET = 7
} silo_comparator;
#define SILO_CMP_SIGNED 8
#define SILO_CMP_WITH_NULLS 16

typedef struct {
   uint8_t opcode;
   uint8_t flags;
   uint16_t column;
   uint32_t a;
   uint64_t b;
} silo_filter_instr;

/* A ready-made roaring bitmap entering the program (lineage / dictionary index, host-evaluated
 * Selection result ...), in the portable Roaring format that Roaring::write emits
 * (roaring_util/roaring_serialize.h:15-30), global row ids. */
typedef struct {
   const uint8_t* data;
   uint64_t size;
} silo_roaring_bytes;

typedef struct {
   uint32_t struct_size;
   uint32_t n_instrs;
   const silo_filter_instr* instrs;
   const uint8_t* blob;
   uint64_t blob_bytes;
   uint32_t n_bitmaps;
   const silo_roaring_bytes* bitmaps;
} silo_filter_program;

/* A 32-bit value column of the table's rows, in the order of the row layout (chunk after chunk): the dictionary ids of
 * a string column, the days of a Date32 column, integers. null_row_ids: ascending row ids ((chunk << 16) | row, global
 * chunk ids) of the null rows, or NULL. Returns the value column's index (>= 0) or a negative status. Replaces, for the
 * predicates of Selection (filter/operators/selection.cpp:94-141: row-at-a-time match() or makeBitmap over all rows,
 * string_in_set / CompareToValueSelection selection.h:76-166, date_between.cpp:61-93), the host-side scan of the column:
 * SILO_OP_PUSH_COMPARE evaluates the predicate for all rows of a chunk inside the filter program. */
int silo_gpu_value_column_upload(silo_gpu_table* table, const uint32_t* values, const uint32_t* null_row_ids, uint64_t n_null_rows);

/* Makes a STATIC index bitmap device resident (the lineage index of lineage_index.h:18-22, a
 * dictionary index, the per-value bitmaps of an indexed string column ...; portable Roaring bytes,
 * global row ids), so that programs refer to it with PUSH_INDEX_BITMAP instead of uploading the bytes
 * with every query. Such indexes are immutable until the next Table::finalize(), which rebuilds the
 * table handle anyway (S1). Do not unregister an id while prepared programs that use it exist. */
int silo_gpu_bitmap_register(silo_gpu_table* table, const uint8_t* data, uint64_t size, uint32_t* id_out);
int silo_gpu_bitmap_unregister(silo_gpu_table* table, uint32_t id);

/* Evaluates the program over every chunk of the table. cardinality may be NULL. */
int silo_gpu_filter_eval(
   silo_gpu_table* table,
   const silo_filter_program* program,
   silo_gpu_filter** out,
   uint64_t* cardinality
);
/* CountFilterNode (count_filter_node.cpp:35-71): only the number of rows that pass. One call, one synchronisation; queries
 * of one shape (the same instruction sequence) replay a captured CUDA graph. */
int silo_gpu_query_count(silo_gpu_table* table, const silo_filter_program* program, uint64_t* cardinality);
/* The same evaluation with DEVICE-RESIDENT inputs, for callers that run one filter many times or must
 * not touch the host in the hot loop: _prepare validates the program and uploads instructions, blob
 * and bitmaps once; _run only enqueues the kernel on `cuda_stream` (cudaStream_t, NULL = the
 * table's stream) into the filter handle created by _prepare, without synchronising. */
typedef struct silo_gpu_program silo_gpu_program;
int silo_gpu_program_prepare(silo_gpu_table* table, const silo_filter_program* program, silo_gpu_program** out, silo_gpu_filter** filter_out);
int silo_gpu_program_run_async(silo_gpu_program* prepared, void* cuda_stream);
/* The prepared program and the Mutations counts of its filter for one column, enqueued on `cuda_stream`
 * without synchronising: silo_gpu_program_run_async + silo_gpu_mutation_counts_async with one launch less
 * (the interpreter also zeroes d_counts and builds the container kernel's work list). */
int silo_gpu_program_run_counts_async(silo_gpu_program* prepared, int column, void* d_counts, void* cuda_stream);
uint64_t silo_gpu_program_device_bytes(const silo_gpu_program* prepared); /* bytes uploaded by _prepare */
void silo_gpu_program_free(silo_gpu_program* prepared); /* does not free the filter handle */

/* Wraps caller-provided dense words: bit r of words[c*1024 + r/64] = row ((first_chunk+c)<<16)|r. */
int silo_gpu_filter_from_words(silo_gpu_table* table, const uint64_t* words, silo_gpu_filter** out);
int silo_gpu_filter_cardinality(const silo_gpu_filter* filter, uint64_t* cardinality);
/* words[1024 * n_chunks] */
int silo_gpu_filter_download(const silo_gpu_filter* filter, uint64_t* words);
void silo_gpu_filter_free(silo_gpu_filter* filter);

/* ---- S3: Mutations action -------------------------------------------------------------------- */

/* counts[symbol * genome_length + position] = SymbolMap<Sym, vector<u32>> of
 * calculateMutationsPerPosition (mutations_node.cpp:268-288), for the rows of this shard.
 * filter == NULL means "all rows" (the cardinality == numRows path, :280-281). uint32 arithmetic is
 * modular exactly like the reference's count_per_local_reference_position. Shards are plain
 * addends: the multi-GPU result is the element-wise sum over ranks. */
int silo_gpu_mutation_counts(
   silo_gpu_table* table,
   int column,
   const silo_gpu_filter* filter,
   uint32_t* counts
);

/* Same layout, but only the rows counts[symbol][*] of the symbols in symbol_mask (bit s = symbol s)
 * are copied to the host; the other rows of `counts` are left untouched. addMutationsToOutput
 * (mutations_node.cpp:292-366) reads only SymbolType::VALID_MUTATION_SYMBOLS (5 of 16 nucleotide
 * symbols), so the D2H copy shrinks from 1.9 MB to 0.6 MB for a SARS-CoV-2 genome. */
int silo_gpu_mutation_counts_symbols(
   silo_gpu_table* table,
   int column,
   const silo_gpu_filter* filter,
   uint64_t symbol_mask,
   uint32_t* counts
);

/* Filter program + Mutations counts in ONE call with ONE host synchronisation (silo_gpu_filter_eval
 * followed by silo_gpu_mutation_counts_symbols costs two): what MutationsNode does for a query with a
 * single sequence column. The filter exists only inside the call; *cardinality (may be NULL) receives
 * its cardinality. A program that is just PUSH_FULL takes the stored-cardinality path
 * (mutations_node.cpp:280-281); any other program takes the intersecting path, which yields the same
 * counts for every filter. */
int silo_gpu_query_mutation_counts(
   silo_gpu_table* table,
   const silo_filter_program* program,
   int column,
   uint64_t symbol_mask,
   uint32_t* counts,
   uint64_t* cardinality
);

/* ---- S3': the action's output pass on the device ------------------------------------------------
 * addMutationsToOutput (mutations_node.cpp:292-366) reads the counts of VALID_MUTATION_SYMBOLS and
 * emits one row per (position, symbol != reference genome symbol) whose count exceeds
 *     threshold_count = min_proportion == 0 ? 0 : uint32(ceil(double(total) * min_proportion) - 1)
 * with total = sum of the valid symbols' counts at the position. That test is IEEE double arithmetic
 * and is evaluated by the finalize kernel with the same operations, so only the emitted
 * (position, symbol, count, total) tuples cross PCIe -- a few KB instead of the count rows -- and the
 * host neither scans nor thresholds. The host computes proportion = double(count) / double(total)
 * exactly as the reference does. Needs the reference genome's symbols on the device
 * (silo_gpu_column_set_reference, once per column). */
typedef struct {
   uint32_t position; /* 0-based */
   uint32_t symbol;
   uint32_t count;
   uint32_t total; /* the "coverage" output field */
} silo_mutation_hit;

/* reference_symbols[genome_length]: metadata->reference_sequence (sequence_column.h:47-56) as symbol ids */
int silo_gpu_column_set_reference(silo_gpu_table* table, int column, const uint8_t* reference_symbols);

/* Filter + counts + output pass in ONE call with ONE host synchronisation. program != NULL: evaluated
 * inside the call (filter is ignored); else filter (NULL = all rows). *hits points at n_hits tuples, ordered by
 * (position, symbol id), in storage of the CALLING THREAD, valid until that thread's next call that returns tuples
 * (the calls on one table are serialised by the library and may come from any number of host threads). cardinality may be NULL (and is only set when a program was given). The full counts stay
 * in device memory (they are not copied to the host by this call). */
int silo_gpu_query_mutation_hits(
   silo_gpu_table* table,
   const silo_filter_program* program,
   const silo_gpu_filter* filter,
   int column,
   uint64_t valid_symbol_mask,
   double min_proportion,
   const silo_mutation_hit** hits,
   uint64_t* n_hits,
   uint64_t* cardinality
);

/* The same for SEVERAL sequence columns under one filter (AminoAcidMutations over all genes: the producer of
 * mutations_node.cpp:372-428 loops the columns of one query): the program is evaluated once, every column's counts and
 * output pass follow on the stream, ONE synchronisation. columns[c].hits / n_hits are set on return (storage of the
 * calling thread, valid until its next call that returns tuples); valid_symbol_mask is clipped to the column's alphabet. */
typedef struct {
   int column;
   uint64_t valid_symbol_mask;
   const silo_mutation_hit* hits; /* out */
   uint64_t n_hits;               /* out */
} silo_column_hits;
int silo_gpu_query_mutation_hits_columns(
   silo_gpu_table* table,
   const silo_filter_program* program,
   silo_column_hits* columns,
   uint32_t n_columns,
   double min_proportion,
   uint64_t* cardinality
);

/* Same, but leaves the counts in device memory (d_counts: n_symbols*genome_length u32) and only
 * enqueues on `cuda_stream` (a cudaStream_t; NULL = the table's own stream) without synchronising,
 * so that a collective (ncclAllReduce on the same stream) can follow with no host round trip. */
int silo_gpu_mutation_counts_async(
   silo_gpu_table* table,
   int column,
   const silo_gpu_filter* filter,
   void* d_counts,
   void* cuda_stream
);

/* The two halves of a Mutations query on a ROW-PARTITIONED table (one process per GPU, each
 * holding a shard of the chunks): every rank evaluates the filter program on its shard and leaves its
 * counts in device memory without synchronising (program H2D + filter + counts on `cuda_stream`); the
 * scheduler sums the counts of the ranks on the same stream (ncclAllReduce: counts are plain addends,
 * mutations_node.cpp:154-203); one rank then runs the output pass of addMutationsToOutput
 * (mutations_node.cpp:292-366) over the summed counts and gets only the emitted tuples back, in
 * page-locked memory owned by the table, ordered by (position, symbol id), valid until the next call.
 * One query in flight per table: the staged program lives in per-table buffers until the stream has
 * been synchronised (silo_gpu_mutation_hits_from_counts does). *shard_cardinality (may be NULL): rows of
 * THIS shard that passed the filter of the preceding silo_gpu_query_mutation_counts_async. */
int silo_gpu_query_mutation_counts_async(
   silo_gpu_table* table,
   const silo_filter_program* program,
   int column,
   void* d_counts,
   void* cuda_stream
);
int silo_gpu_mutation_hits_from_counts(
   silo_gpu_table* table,
   int column,
   const void* d_counts,
   uint64_t valid_symbol_mask,
   double min_proportion,
   void* cuda_stream,
   const silo_mutation_hit** hits,
   uint64_t* n_hits,
   uint64_t* shard_cardinality
);

/* ---- Row-partitioned tables: the multi-GPU scheduler's device side -----------------------------
 * One process per GPU; every rank uploads the chunks of its shard as a table of its own. The per-rank counts of the
 * Mutations action are plain addends (the local reference is the same on all shards) and only one rank -- the root,
 * rank 0 -- needs their sum, for the output pass. No collective kernel runs: the finalize kernel of every rank stores
 * its rows of the valid mutation symbols straight into a gather area in the ROOT's memory over NVLink (CUDA IPC peer
 * mapping; 8-byte words tagged with the query number, no fence, no counter), and the root's own finalize kernel adds
 * the ranks' rows to its counts and runs addMutationsToOutput (mutations_node.cpp:307-363) on the device. Replaces, for a row-partitioned table, the loop over partitions that the reference's
 * MutationsNode producer would run on one host (mutations_node.cpp:372-428).
 *
 *   every rank:  silo_gpu_shard_group_init(table, column, valid mask, rank, world, my_handle)
 *                -- exchange the handles (SILO_SHARD_HANDLE_BYTES each) by any means: MPI, torch.distributed, a file --
 *                silo_gpu_shard_group_connect(table, all_handles /+ world x SILO_SHARD_HANDLE_BYTES, by rank +/)
 *   per query:   every rank, in the same order:  silo_gpu_sharded_query_enqueue(table, program, stream)   (no sync)
 *                the root:                       silo_gpu_sharded_collect(...)                            (one sync)
 * Ranks other than the root never synchronise with the host; they may run up to 3 queries ahead of the root (a rank
 * whose gather slot has not been handed back waits inside its finalize kernel). All device-side waits are bounded
 * (~2 s): a rank that never delivers becomes SILO_E_CUDA at the root, not a hung GPU. One sharded query stream per
 * table; the table's other entry points stay usable between sharded queries. */
#define SILO_SHARD_HANDLE_BYTES 128
int silo_gpu_shard_group_init(silo_gpu_table* table, int column, uint64_t valid_symbol_mask, int rank, int world, void* handle_out);
int silo_gpu_shard_group_connect(silo_gpu_table* table, const void* handles);
/* every rank: filter program + counts of the group's column + this rank's rows to the root; enqueue only */
int silo_gpu_sharded_query_enqueue(silo_gpu_table* table, const silo_filter_program* program, void* cuda_stream);
/* the root, once per enqueued query and in the same order: waits (on the device) for all ranks, sums, runs the output
 * pass; hits / n_hits / cardinality (the sum of the ranks' filter cardinalities) as silo_gpu_query_mutation_hits.
 * d_summed_counts: NULL, or n_symbols*genome_length u32 in device memory that receive the summed rows of the valid
 * symbols (the other rows are left alone). */
int silo_gpu_sharded_collect(
   silo_gpu_table* table,
   double min_proportion,
   void* d_summed_counts,
   void* cuda_stream,
   const silo_mutation_hit** hits,
   uint64_t* n_hits,
   uint64_t* cardinality
);
/* the root's two calls in one (one launch -- a replayed CUDA graph per query shape --, one synchronisation), on the
 * table's own stream; the other ranks call silo_gpu_sharded_query_enqueue. The collect happens INSIDE the root's
 * finalize kernel here: it waits for the other ranks' rows, adds the root's own counts (never stored to the gather
 * area) and runs the output pass over the sums -- one kernel and one pass over the rows less than enqueue + collect.
 * Every earlier query of the group must have been collected (else SILO_E_CUDA, and the group has to be re-created). */
int silo_gpu_sharded_query_hits(
   silo_gpu_table* table,
   const silo_filter_program* program,
   double min_proportion,
   void* d_summed_counts,
   const silo_mutation_hit** hits,
   uint64_t* n_hits,
   uint64_t* cardinality
);
/* every rank: like silo_gpu_sharded_query_enqueue for a program that silo_gpu_program_prepare made device resident
 * (nothing is uploaded; replayable inside a captured CUDA graph: slot and generation live in device memory) */
int silo_gpu_program_run_sharded_async(silo_gpu_program* prepared, void* cuda_stream);
/* the root only: the same with the collect inside -- the root's finalize kernel waits for the other ranks' rows, adds
 * its own counts and writes the sums of the valid symbols' rows to d_summed_counts ([n_symbols][genome_length] u32,
 * may be NULL). One kernel fewer than run_sharded_async + sharded_collect_async and no store of the root's own rows;
 * every earlier query of the group must have been collected. */
int silo_gpu_program_run_sharded_collect_async(silo_gpu_program* prepared, void* d_summed_counts, void* cuda_stream);
/* the same without the output pass and without synchronising (device-resident pipelines) */
int silo_gpu_sharded_collect_async(silo_gpu_table* table, void* d_summed_counts, void* cuda_stream);
void silo_gpu_shard_group_free(silo_gpu_table* table);

/* ---- BitmapAggregationNode: co-occurrence / groupBy over sequence positions and indexed columns ---
 * Replaces buildGroups + computeCombinations of operators/bitmap_aggregation_node.cpp:53-139,224-249
 * (reached from BitmapAggregationNode::addToExecPlan :304-356). The groups of a dimension are disjoint,
 * so the reference's recursive partition is a GROUP BY over per-row group codes:
 *   sequence position: code = symbol id the row carries at the position (stored container -> that
 *     symbol; covered and not N -> local reference symbol; otherwise the missing symbol), or n_symbols
 *     of the column when the row's sequence is null (the null group, :80-90);
 *   index bitmaps: code = index g of the value bitmap that holds the row, n_groups for the null bitmap;
 *     a row in none of them is in no combination.
 * key = the codes packed with dimension 0 in the most significant bits: 5 bits per sequence-position
 * dimension, 8 bits per index-bitmap dimension, 63 bits in total at most; ascending key order is the
 * reference's depth-first output order.
 * LIMITS the reference does not have (SILO_E_INVALID_ARGUMENT beyond them; the caller keeps its CPU path for such
 * queries): at most 254 value groups per index-bitmap dimension (one code byte per row and dimension), 63 key bits
 * = e.g. 12 sequence positions, or 7 index dimensions, or 6 positions + 4 index dimensions. The value bitmaps of a
 * dimension must be DISJOINT (they are, for a dictionary index): a row held by two of them is counted once here,
 * in the first, but once per bitmap by the reference's partition().
 * Row-partitioned tables: counts of disjoint row sets add up, so every rank runs the call on its shard with the
 * same dimensions and one rank sums the (key, count) lists per key (host/bitmap_aggregation_node.h mergeShards). */
enum { SILO_DIM_SEQUENCE_POSITION = 0, SILO_DIM_INDEX_BITMAPS = 1 };
typedef struct {
   uint32_t kind;
   int32_t column;             /* sequence position: column index */
   uint32_t position;          /* sequence position: 0-based */
   uint32_t n_groups;          /* index bitmaps: number of value bitmaps (<= 254), in output order */
   const uint32_t* bitmap_ids; /* index bitmaps: ids from silo_gpu_bitmap_register */
   uint32_t null_bitmap_id;    /* index bitmaps: id of the null bitmap, or UINT32_MAX */
} silo_group_dimension;
typedef struct {
   uint64_t key;
   uint64_t count;
} silo_combination;
/* program != NULL: the filter program is evaluated inside the call; else `filter` (NULL = all rows).
 * *combinations points at n_combinations entries (count > 0 each) ordered by key, valid until the next
 * call of this function on the calling thread. One host synchronisation unless more than 2047
 * combinations come back or the hash table has to grow. */
int silo_gpu_query_combinations(
   silo_gpu_table* table,
   const silo_filter_program* program,
   const silo_gpu_filter* filter,
   const silo_group_dimension* dimensions,
   uint32_t n_dimensions,
   const silo_combination** combinations,
   uint64_t* n_combinations,
   uint64_t* cardinality
);

/* ---- measurement hooks (bench.py / profiles; not needed by the reference) --------------------- */

typedef struct {
   uint64_t containers;          /* containers of the chunks touched by the last mutation_counts call */
   uint64_t algorithmic_bytes;   /* whole call: descriptors + payloads + filter tiles + 8 B/row of the
                                    touched chunks + the counts written */
   uint64_t counts_kernel_bytes; /* dominant kernel only: descriptors + payloads + filter tiles */
   uint64_t kernel_launches;     /* kernels launched through this table since it was created */
   float last_counts_kernel_ms;  /* CUDA-event duration of the dominant (container AND) kernel ... */
   float last_total_ms;          /* ... and of the whole mutation_counts enqueue: averages over */
   uint64_t timed_calls;         /* ... this many calls since the previous silo_gpu_get_stats */
} silo_gpu_stats;
/* Synchronises the stream of the last mutation_counts call, reads the per-call CUDA events recorded on
 * the launching stream (a ring of 256 calls) and resets the averaging window. Byte counts describe
 * the last call. */
int silo_gpu_get_stats(silo_gpu_table* table, silo_gpu_stats* out);
/* The threshold sweep kernel (a THR_PROFILE over a large column): mean CUDA-event duration of its launches since the
 * previous call (at most the last 64, launches inside a stream capture excluded) and its algorithmic bytes per launch
 * (16 B descriptor + reference-format payload of every container of the column). Synchronises. */
int silo_gpu_get_sweep_stats(silo_gpu_table* table, float* mean_kernel_ms, uint64_t* algorithmic_bytes, uint64_t* timed_calls);

#ifdef __cplusplus
}
#endif
#endif /* SILO_B200_H */
