// Internal declarations of libsilo_b200.so (sm_100a only). Public surface: include/silo_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/silo_b200.h"

namespace silo {

// ---------------------------------------------------------------------------------------------
// HBM layout of one sequence column (see DESIGN.md "Data layout in HBM")
// ---------------------------------------------------------------------------------------------

constexpr uint32_t TILE_WORDS = 1024;        // one chunk's dense filter tile: 1024 x u64 = 8 KiB
constexpr uint32_t TILE_BYTES = TILE_WORDS * 8;
// Geometry of the container kernel (mutations.cu) and, with it, of the segments the upload cuts a chunk into:
// `warps` consumer warps take `pieces` pieces each from one ring stage, so a segment holds at most
// warps * pieces pieces (of <= 1 KiB) of ONE kind; the ring has `stages` stages. Process-wide (SILO_K1_VARIANT picks
// one of the compiled variants, for measurements); a column remembers the capacity it was cut for.
struct K1Geometry {
   uint32_t warps;
   uint32_t pieces;
   uint32_t stages;
   uint32_t variant;
   uint32_t segmentPieces() const { return warps * pieces; }
   uint32_t segmentPayloadBytes() const { return warps * pieces * 1024u; }
   uint32_t inlinePieces() const { return warps * pieces * 32u; }  // descriptor-only pieces: one per LANE
};
const K1Geometry& k1Geometry();

constexpr uint32_t TYPE_BITSET = 1;  // CRoaring typecodes, roaring_container.h:131-145
constexpr uint32_t TYPE_ARRAY = 2;
constexpr uint32_t TYPE_RUN = 3;

// Device-side piece kinds (DevContainer::packed[28:26]). The payload of a stored container is
// RE-ENCODED at upload into the form the container kernel decodes with the fewest instructions;
// host bitmaps entering a filter program keep the CRoaring payloads (RAW kinds).
constexpr uint32_t KIND_INLINE = 0;     // array of 1..2 values stored in DevContainer::aux, no payload
constexpr uint32_t KIND_BITSET = 1;     // aux = first_word | (n_words << 16); n_words x u64
constexpr uint32_t KIND_ARRAY_T = 2;    // lane-transposed u16 values, see below
constexpr uint32_t KIND_RUNS_W = 3;     // runs cut at 32-row word boundaries, one u32 entry each; aux = entries
constexpr uint32_t KIND_WORDRANGE = 4;  // whole 32-row words [wa, wb): u32 entries wa | wb << 16; aux = entries
constexpr uint32_t KIND_RAW_ARRAY = 5;  // cardinality sorted u16 values (CRoaring array container)
constexpr uint32_t KIND_RAW_RUN = 6;    // aux pairs {start, length-1} (CRoaring run container)
constexpr uint32_t PIECE_BYTES = 1024;  // max payload of one piece of a stored container
// A ring stage of the container kernel (= a segment) holds pieces of one CLASS: one kind, and either all with two
// payload regions (a full first one) or all with at most one.
__host__ __device__ constexpr uint32_t stageClass(uint32_t kind, bool two_regions) {
   return kind * 2u + (two_regions ? 1u : 0u);
}

// Both array and run pieces are laid out in REGIONS of up to 512 bytes that a warp reads with ONE
// 128-bit shared-memory load per lane (a piece is at most two regions, so a warp can pull a whole
// piece into registers, hand the stage back to the producer, and only then do the lookups).
//
// KIND_ARRAY_T: n = cardinality <= 512 values; region r holds count = min(256, n - 256 r) of them in
// P = ceil(count / 8) lanes: lane L's 16 bytes are the values with in-region index L + P*j, j = 0..7
// (two per 32-bit word, low half first; indices >= count are padding). The 32 simultaneous tile
// lookups of one j therefore belong to P CONSECUTIVE sorted values (mostly distinct banks).
// A stored value is row ^ 31 (the low five bits hold 31 - (row & 31), the left shift that moves the
// row's bit of its tile word to bit 31).
constexpr uint32_t ARRAY_REGION_VALUES = 256;
constexpr uint32_t ARRAY_VALUE_FLIP = 31;
__host__ __device__ inline uint32_t arrayRegionLanes(uint32_t count) {
   return (count + 7) / 8;
}
__host__ __device__ inline uint32_t arrayPieceBytes(uint32_t n) {
   return n <= ARRAY_REGION_VALUES ? arrayRegionLanes(n) * 16u : 512u + arrayRegionLanes(n - ARRAY_REGION_VALUES) * 16u;
}

// KIND_RUNS_W entry: [31:20] index of the tile's 32-bit word (2048 = the zero pad word: a padding
// entry) | [19:10] zero | [9:5] (32 - length) & 31 | [4:0] first bit; 1 <= length <= 32 - first bit.
// |run AND tile| = popc((tile32[word] >> first) << (32 - length)). n = aux <= 256 entries; region r
// holds count = min(128, n - 128 r) of them in P = ceil(count / 4) lanes: lane L's 16 bytes are the
// entries L + P*j, j = 0..3, padded with padding entries.
constexpr uint32_t RUNS_REGION_ENTRIES = 128;
constexpr uint32_t RUNS_PAD_ENTRY = 2048u << 20;
__host__ __device__ inline uint32_t runsRegionLanes(uint32_t count) {
   return (count + 3) / 4;
}
__host__ __device__ inline uint32_t runsPieceBytes(uint32_t n) {
   return n <= RUNS_REGION_ENTRIES ? runsRegionLanes(n) * 16u : 512u + runsRegionLanes(n - RUNS_REGION_ENTRIES) * 16u;
}
__host__ __device__ inline uint32_t runEntry(uint32_t word, uint32_t first_bit, uint32_t length) {
   return (word << 20) | (((32u - length) & 31u) << 5) | first_bit;
}
__host__ __device__ inline uint32_t runEntryMask(uint32_t entry) {  // the rows of the entry inside its word
   const uint32_t length = 32u - ((entry >> 5) & 31u);
   return (length == 32u ? 0xFFFFFFFFu : ((1u << length) - 1u)) << (entry & 31u);
}

// 16-byte device descriptor of one piece of a stored diff container (chunk-major order).
struct __align__(16) DevContainer {
   uint32_t position;
   uint32_t offset4;  // payload offset from the slab start, in 4-byte units
   uint32_t packed;   // [15:0] rows-1 | [20:16] symbol | [25:21] local-reference symbol | [28:26] kind
   uint32_t aux;      // see the kinds

   __host__ __device__ uint32_t cardinality() const { return (packed & 0xFFFFu) + 1u; }
   __host__ __device__ uint32_t symbol() const { return (packed >> 16) & 0x1Fu; }
   __host__ __device__ uint32_t refSymbol() const { return (packed >> 21) & 0x1Fu; }
   __host__ __device__ uint32_t type() const { return (packed >> 26) & 0x7u; }
   __host__ __device__ uint32_t firstWord() const { return aux & 0xFFFFu; }
   __host__ __device__ uint32_t wordCount() const { return aux >> 16; }
   __host__ __device__ static uint32_t pack(uint32_t rows, uint32_t symbol, uint32_t ref_symbol, uint32_t kind) {
      return (rows - 1u) | (symbol << 16) | (ref_symbol << 21) | (kind << 26);
   }
};
static_assert(sizeof(DevContainer) == 16);

// A segment = pieces of ONE chunk and ONE kind as one block of the slab, [descriptors | payloads], moved into a ring
// stage of the container kernel by ONE 1-D bulk (TMA) copy. The descriptors of the block are copies of the chunk-major
// ones the filter interpreter searches, except that `position` is replaced by the piece's index into the counts
// array, symbol * genome_length + position (what the kernel adds its result to).
struct __align__(16) DevSegment {
   uint32_t payload_offset16;  // block start from the slab start, in 16-byte units (the slab is <= 16 GiB, see DevContainer::offset4)
   uint32_t bytes_and_kind;    // [23:0] block bytes: 16 per descriptor + payloads, a multiple of 16 | [31:24] the stageClass() of its pieces
   uint32_t reserved;
   uint32_t chunk_and_count;   // [15:0] local chunk index | [31:16] number of descriptors

   __host__ __device__ uint32_t chunk() const { return chunk_and_count & 0xFFFFu; }
   __host__ __device__ uint32_t descCount() const { return chunk_and_count >> 16; }
   __host__ __device__ uint32_t blockBytes() const { return bytes_and_kind & 0xFFFFFFu; }
   __host__ __device__ uint32_t kind() const { return bytes_and_kind >> 24; }
   __host__ __device__ uint64_t payloadOffset() const { return static_cast<uint64_t>(payload_offset16) << 4; }
};
static_assert(sizeof(DevSegment) == 16);

struct DevColumn {
   uint32_t n_symbols;
   uint32_t genome_length;
   uint32_t missing_symbol;
   uint32_t n_chunks;
   uint64_t n_containers;
   uint32_t n_segments;
   const uint8_t* local_reference;      // [genome_length]
   const DevContainer* containers;      // sorted (chunk, position, symbol)
   const uint8_t* payload;              // slab
   const uint32_t* chunk_desc_begin;    // [n_chunks + 1]
   const DevSegment* segments;          // [n_segments]
   const uint32_t* chunk_seg_begin;     // [n_chunks + 1]
   const uint2* start_end;              // rows of the shard back to back
   const uint32_t* chunk_row_begin;     // [n_chunks + 1] into start_end
   const uint32_t* chunk_missing_begin; // [n_chunks + 1] into missing_row / missing_offsets
   const uint16_t* missing_row;         // row_in_chunk of each row with N positions
   const uint64_t* missing_offsets;     // [n_rows_with_missing + 1] into missing_runs
   const uint2* missing_runs;           // {first, end_exclusive}
   const uint64_t* null_words;          // [n_chunks * 1024] or nullptr when the column has no nulls
   const uint8_t* global_reference;     // [genome_length] reference genome symbols, or nullptr (silo_gpu_column_set_reference)
};

// ---------------------------------------------------------------------------------------------
// host-side objects behind the opaque handles
// ---------------------------------------------------------------------------------------------

struct Stats {
   uint64_t containers = 0;
   uint64_t algorithmic_bytes = 0;
   uint64_t counts_kernel_bytes = 0;
   uint64_t kernel_launches = 0;
   float last_counts_kernel_ms = 0;
   float last_total_ms = 0;
   uint64_t timed_calls = 0;
};

struct HostColumn {
   DevColumn dev{};
   std::vector<void*> allocations;
   // per-chunk byte totals for the algorithmic-bytes accounting
   std::vector<uint64_t> chunk_desc_payload_bytes;  // 16 B/descriptor + payload bytes
   std::vector<uint64_t> chunk_missing_rows;
   std::vector<uint64_t> chunk_containers;
   uint64_t device_bytes = 0;
   uint32_t segment_pieces = 0;  // the K1Geometry capacity the segments were cut for
   // Threshold sweep (filter_eval.cu thresholdSweepKernel): the column's pieces cut into one contiguous range per
   // CTA, balanced by bytes; sweep_flushes[c] = CTAs whose range holds pieces of chunk c
   uint32_t sweep_ctas = 0;
   uint32_t* d_sweep_split = nullptr;    // [sweep_ctas + 1] piece indices
   uint32_t* d_sweep_flushes = nullptr;  // [n_chunks]
   uint32_t sweep_max_flushes = 0;
};

}  // namespace silo

namespace silo {
// The difference array is followed by the sums of its 256-element blocks (accumulated by the same
// kernel), so that the finalize kernel gets the prefix in front of a block from <= 117 numbers
// instead of re-summing up to 29,903.
constexpr uint32_t DIFF_BLOCK = 256;
__host__ __device__ inline uint32_t diffPadded(uint32_t genome_length) {
   return (genome_length + 1 + DIFF_BLOCK - 1) / DIFF_BLOCK * DIFF_BLOCK;
}
__host__ __device__ inline uint32_t diffWords(uint32_t genome_length) {  // whole scratch array
   return diffPadded(genome_length) + diffPadded(genome_length) / DIFF_BLOCK;
}

}  // namespace silo

namespace silo {
struct ShardGroup;  // mutations.cu: this table as one rank of a row-partitioned table
struct DevValueColumn {
   const uint32_t* values;      // rows of the shard back to back (chunk-major)
   const uint64_t* null_words;  // [n_chunks * 1024] or nullptr
};
}

struct silo_gpu_ctx {
   int device = 0;
   int sm_count = 0;
   cudaStream_t stream = nullptr;
};

struct silo_gpu_table {
   silo_gpu_ctx* ctx = nullptr;
   uint32_t first_chunk = 0;
   uint32_t n_chunks = 0;
   std::vector<uint32_t> chunk_sizes;
   uint64_t n_rows = 0;
   uint32_t* d_chunk_sizes = nullptr;
   std::vector<silo::HostColumn*> columns;
   // per-query scratch (calls on one table are serialised by `mutex`; tables are independent)
   std::mutex mutex;
   uint32_t* d_work_state = nullptr;   // [0] number of work items, [1] the container kernel's claim counter,
                                       // [2] finalize blocks done; all zero between queries (reset by the finalize kernel)
   silo::DevSegment* d_work_items = nullptr;  // [max n_segments over the columns]: segments of the active chunks
   uint32_t work_items_capacity = 0;
   // coverage difference array + block totals; all-zero between queries (the finalize kernel clears what
   // it has read), so the coverage kernel depends on nothing but the filter
   uint32_t* d_coverage_diff = nullptr;  // [diffWords(max genome_length)]
   uint32_t coverage_diff_capacity = 0;
   uint32_t* d_counts = nullptr;  // staging for the synchronous API
   uint64_t counts_capacity = 0;
   uint32_t* h_counts_pinned = nullptr;
   // output pass on the device (silo_gpu_query_mutation_hits): [0] = {number of hits, 0, 0, 0}, then the tuples
   silo_mutation_hit* h_hits_pinned = nullptr;
   uint64_t hits_capacity = 0;  // tuples, without the header
   // The fused query calls (silo_gpu_query_*) run on persistent buffers -- a device copy of the staged program
   // and one filter -- so that the launch sequence of a query SHAPE can be replayed as a CUDA graph:
   // what differs between two queries of one shape is only the content of the staging buffer.
   uint8_t* d_staging_fixed = nullptr;  // device copy of h_staging_pinned (same capacity)
   silo_gpu_filter* query_filter = nullptr;
   struct CachedGraph {
      std::string key;
      cudaGraphExec_t exec = nullptr;
   };
   std::vector<CachedGraph> query_graphs;  // a few shapes, replaced round-robin
   size_t next_graph_slot = 0;
   std::string last_query_key;             // a shape is captured the second time in a row it is seen
   unsigned long long* h_scalars_pinned = nullptr;  // [4]: a query's filter cardinality ([0]) and error flag ([2]) land here
   uint8_t* h_staging_pinned = nullptr;  // program upload staging (grow-only)
   size_t staging_capacity = 0;
   cudaEvent_t ev_free_fence = nullptr;  // orders stream-ordered frees after foreign-stream users
   // recorded behind the H2D copy of a staged query that the caller does not synchronise (the _async and sharded
   // entries): the next staging waits for it before it overwrites the pinned buffer
   cudaEvent_t ev_staging_copied = nullptr;
   bool staging_copy_pending = false;
   // the coverage kernel runs beside the container kernel on an auxiliary stream (fork / join)
   cudaStream_t aux_stream = nullptr;
   cudaEvent_t ev_fork = nullptr;
   cudaEvent_t ev_join = nullptr;
   uint32_t* d_chunk_popcount_full = nullptr;  // popcounts of the "all rows" filter (= chunk sizes)
   uint64_t* d_full_words = nullptr;           // layout mask tiles [n_chunks * 1024]
   // ring of CUDA-event pairs recorded on the launching stream around every mutation_counts
   // enqueue (whole call) and around its dominant kernel; silo_gpu_get_stats averages and resets
   static constexpr int EVENT_RING = 256;
   cudaEvent_t ev_begin[EVENT_RING] = {}, ev_k1_begin[EVENT_RING] = {}, ev_k1_end[EVENT_RING] = {}, ev_end[EVENT_RING] = {};
   uint64_t timed_calls = 0;  // since the last silo_gpu_get_stats
   // what the last mutation_counts enqueue ran on / with (read back by silo_gpu_get_stats)
   cudaStream_t last_stream = nullptr;
   int last_column = -1;
   const uint32_t* last_popcounts = nullptr;
   bool last_was_full = false;
   silo::Stats stats;
   uint64_t device_bytes = 0;
   // index bitmaps made device resident by silo_gpu_bitmap_register (one block each:
   // [descriptors | payload]); id = index, a freed slot keeps d_block == nullptr
   struct RegisteredBitmap {
      uint8_t* d_block = nullptr;
      uint32_t n_containers = 0;
      uint64_t payload_offset = 0;  // of the payload inside the block
      uint64_t bytes = 0;
   };
   std::vector<RegisteredBitmap> registered;
   // per-row counters of a Threshold profile pass computed by thresholdSweepKernel in front of the interpreter:
   // [n_chunks][32768] packed u16 pairs; used by programs over columns with at least sweep_min_pieces pieces
   uint32_t* d_sweep_counters = nullptr;
   uint64_t sweep_min_pieces = 1u << 16;
   // measurement: CUDA events around the last launches of the sweep kernel (outside stream captures)
   static constexpr int SWEEP_EVENT_RING = 64;
   cudaEvent_t ev_sweep_begin[SWEEP_EVENT_RING] = {}, ev_sweep_end[SWEEP_EVENT_RING] = {};
   uint64_t sweep_timed_calls = 0;
   uint64_t sweep_algorithmic_bytes = 0;  // descriptors + payloads (reference format) of the swept column
   cudaStream_t sweep_stream = nullptr;
   silo::ShardGroup* shard = nullptr;
   // value columns (silo_gpu_value_column_upload) for SILO_OP_PUSH_COMPARE
   std::vector<silo::DevValueColumn> value_columns;
   uint32_t* d_chunk_row_begin = nullptr;  // [n_chunks + 1]
};

struct silo_gpu_filter {
   silo_gpu_table* table = nullptr;
   // one stream-ordered allocation: [words | per-chunk popcounts | cardinality | error flag]
   uint64_t* d_words = nullptr;           // [n_chunks * 1024]
   uint32_t* d_chunk_popcount = nullptr;  // [n_chunks]
   unsigned long long* d_cardinality = nullptr;
   uint32_t* d_error_flag = nullptr;
   // silo_gpu_filter_eval: a leaf bitmap held row ids outside the row layout. The filter carries them like the
   // reference's bitmaps do; the consumers that index per-row data by the filter's rows refuse it.
   bool out_of_layout = false;
};

namespace silo {

void setLastError(const std::string& message);

// filter_eval.cu, for the fused query call of mutations.cu. The caller holds table->mutex. Stages and
// launches the program on `stream` WITHOUT synchronising; *d_staging_out must be freed
// (cudaFreeAsync on the same stream) after the kernel, the filter with releaseFilterLocked.
silo_gpu_filter* evalProgramAsync(silo_gpu_table* table, const silo_filter_program* program, cudaStream_t stream, uint8_t** d_staging_out);
void releaseFilterLocked(silo_gpu_filter* filter);

// The same for the fused query calls, split into a host part and an enqueue part so that the enqueue can be
// captured into a CUDA graph: stageQueryLocked validates the program and writes everything the kernel reads
// into table->h_staging_pinned (device addresses inside refer to table->d_staging_fixed);
// enqueueStagedQuery issues the H2D copy and the interpreter kernel into table->query_filter.
struct StagedQuery {
   alignas(16) unsigned char params[192];  // the interpreter's kernel parameters (filter_eval.cu EvalParams)
   uint64_t staged_bytes = 0;
   uint32_t shared_bytes = 0;
};
// prepare_column >= 0: the interpreter also zeroes prepare_counts and builds the container kernel's work list for that
// column (the caller then passes prepared = true to enqueueMutationCounts)
void stageQueryLocked(silo_gpu_table* table, const silo_filter_program* program, StagedQuery* out, int prepare_column = -1, uint32_t* prepare_counts = nullptr);
void enqueueStagedQuery(silo_gpu_table* table, const StagedQuery& staged, cudaStream_t stream, bool scalars_are_zero = true);
void dropQueryGraphsLocked(silo_gpu_table* table);
void freeShardGroup(silo_gpu_table* table);  // mutations.cu
int shardGroupColumnLocked(const silo_gpu_table* table);
void enqueuePreparedShardedLocked(silo_gpu_table* table, const silo_gpu_filter* filter, cudaStream_t stream, bool collect_here = false, void* d_summed_counts = nullptr);
// mutations.cu: coverage + container + finalize kernels for a filter whose interpreter launch already zeroed
// d_counts and built the work list (caller holds table->mutex); records the per-call timing events
void enqueuePreparedCountsLocked(silo_gpu_table* table, int column, const silo_gpu_filter* filter, uint32_t* d_counts, cudaStream_t stream);

// NVTX range around the enqueue of a kernel, named after the evobench scope of the reference code it replaces
// (EVOBENCH_SCOPE in mutations_node.cpp:43-197, threshold.cpp:65, intersection.cpp:60, selection.cpp:95): a timeline of
// this library lines up with one of the reference. Costs a few nanoseconds when no profiler is attached.
struct NvtxRange {
   explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
   ~NvtxRange() { nvtxRangePop(); }
   NvtxRange(const NvtxRange&) = delete;
   NvtxRange& operator=(const NvtxRange&) = delete;
};

struct ApiError : std::runtime_error {
   int status;
   ApiError(int status, const std::string& message) : std::runtime_error(message), status(status) {}
};

#define SILO_CUDA_CHECK(expr)                                                                      \
   do {                                                                                            \
      cudaError_t silo_cuda_status = (expr);                                                       \
      if (silo_cuda_status != cudaSuccess) {                                                       \
         throw ::silo::ApiError(                                                                   \
            silo_cuda_status == cudaErrorMemoryAllocation ? SILO_E_OUT_OF_MEMORY : SILO_E_CUDA,    \
            std::string(#expr) + ": " + cudaGetErrorString(silo_cuda_status)                       \
         );                                                                                        \
      }                                                                                            \
   } while (0)

template <typename Fn>
int guarded(Fn&& fn) {
   // a non-sticky error left behind by an earlier, unrelated runtime call on this thread (a free on a stream that
   // was destroyed meanwhile, a pointer query on plain host memory ...) must not be blamed on this call's launches
   cudaGetLastError();
   try {
      fn();
      return SILO_OK;
   } catch (const ApiError& error) {
      setLastError(error.what());
      return error.status;
   } catch (const std::exception& error) {
      setLastError(error.what());
      return SILO_E_INVALID_ARGUMENT;
   }
}

inline void require(bool condition, const char* message) {
   if (!condition) {
      throw ApiError(SILO_E_INVALID_ARGUMENT, message);
   }
}

// Programmatic dependent launch (sm_90+): the kernels of a query form a chain on one stream -- filter program,
// container kernel, finalize kernel, the next query's filter program ... -- and each runs for microseconds, so the
// launch latency and the cold start of a kernel (first instruction fetches, parameter and index loads that do not
// depend on the predecessor) are a visible share of a step. A kernel launched with launchDependent() may start while
// its predecessor drains; everything that reads or writes what an earlier kernel of the stream touches comes after
// gridDependencyWait() (a no-op when the kernel was launched plainly). Measured on the bench step (B200): 74.7 -> 74.1 us
// at N = 1, 81.7 -> 80.6 us at N = 2 -- the kernels' independent prologues are short --, so the attribute is OFF unless
// SILO_PDL=1 asks for it: half a microsecond does not pay for a second launch path in production.
__device__ __forceinline__ void gridDependencyWait() {
   asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ void gridDependencyLaunch() {
   asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
inline bool dependentLaunchEnabled() {
   static const bool enabled = [] {
      const char* flag = std::getenv("SILO_PDL");
      return flag != nullptr && flag[0] == '1';
   }();
   return enabled;
}
template <typename... Params, typename... Args>
inline cudaError_t launchDependent(void (*kernel)(Params...), dim3 grid, dim3 block, size_t shared_bytes, cudaStream_t stream, Args&&... args) {
   cudaLaunchConfig_t config{};
   config.gridDim = grid;
   config.blockDim = block;
   config.dynamicSmemBytes = shared_bytes;
   config.stream = stream;
   cudaLaunchAttribute attribute{};
   attribute.id = cudaLaunchAttributeProgrammaticStreamSerialization;
   attribute.val.programmaticStreamSerializationAllowed = 1;
   config.attrs = &attribute;
   config.numAttrs = dependentLaunchEnabled() ? 1 : 0;
   return cudaLaunchKernelEx(&config, kernel, Params(std::forward<Args>(args))...);
}

// SILO_QUERY_GRAPHS=0 keeps the fused query calls on plain launches (debugging aid)
inline bool queryGraphsEnabled() {
   static const bool enabled = [] {
      const char* flag = std::getenv("SILO_QUERY_GRAPHS");
      return flag == nullptr || flag[0] != '0';
   }();
   return enabled;
}


// The launch sequence of a fused query touches persistent buffers only, so it is the same for every query of one
// SHAPE (`key`: every kernel parameter and copy size the enqueue bakes into nodes) and can be replayed as a CUDA graph.
// A shape is captured the second time in a row it is seen; a few graphs are kept per table. Returns the graph to
// launch, or nullptr (first sight, or graphs disabled): the caller then enqueues plainly.
template <typename Enqueue>
inline cudaGraphExec_t queryGraphFor(silo_gpu_table* table, std::string key, cudaStream_t stream, Enqueue&& enqueueAll) {
   if (!queryGraphsEnabled()) {
      return nullptr;
   }
   cudaGraphExec_t replay = nullptr;
   for (const silo_gpu_table::CachedGraph& cached : table->query_graphs) {
      if (cached.key == key) {
         replay = cached.exec;
         break;
      }
   }
   if (replay == nullptr && key == table->last_query_key) {
      cudaGraph_t graph = nullptr;
      SILO_CUDA_CHECK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
      try {
         enqueueAll();
      } catch (...) {
         cudaStreamEndCapture(stream, &graph);
         if (graph != nullptr) {
            cudaGraphDestroy(graph);
         }
         cudaGetLastError();
         throw;
      }
      SILO_CUDA_CHECK(cudaStreamEndCapture(stream, &graph));
      cudaGraphExec_t exec = nullptr;
      const cudaError_t instantiated = cudaGraphInstantiate(&exec, graph, 0);
      cudaGraphDestroy(graph);
      SILO_CUDA_CHECK(instantiated);
      constexpr size_t MAX_QUERY_GRAPHS = 8;
      if (table->query_graphs.size() < MAX_QUERY_GRAPHS) {
         table->query_graphs.push_back({key, exec});
      } else {
         silo_gpu_table::CachedGraph& slot = table->query_graphs[table->next_graph_slot++ % MAX_QUERY_GRAPHS];
         cudaGraphExecDestroy(slot.exec);
         slot = {key, exec};
      }
      replay = exec;
   }
   table->last_query_key = std::move(key);
   return replay;
}


template <typename T>
T* deviceAlloc(size_t count, uint64_t* accounting = nullptr) {
   void* ptr = nullptr;
   const size_t bytes = (count == 0 ? 1 : count) * sizeof(T);
   SILO_CUDA_CHECK(cudaMalloc(&ptr, bytes));
   if (accounting != nullptr) {
      *accounting += bytes;
   }
   return static_cast<T*>(ptr);
}

// per-query scratch comes from the device's stream-ordered pool (cudaMallocAsync with an unlimited
// release threshold, set in silo_gpu_init): after warm-up an allocation costs microseconds
template <typename T>
T* poolAlloc(size_t count, cudaStream_t stream) {
   void* ptr = nullptr;
   SILO_CUDA_CHECK(cudaMallocAsync(&ptr, (count == 0 ? 1 : count) * sizeof(T), stream));
   return static_cast<T*>(ptr);
}

template <typename T>
T* deviceUpload(const std::vector<T>& host, cudaStream_t stream, uint64_t* accounting = nullptr) {
   T* ptr = deviceAlloc<T>(host.size(), accounting);
   if (!host.empty()) {
      SILO_CUDA_CHECK(cudaMemcpyAsync(ptr, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice, stream));
   }
   return ptr;
}

// ---- device helpers shared by the kernels ---------------------------------------------------

#ifdef __CUDACC__

// zeroes words[0 .. n_words) with n_threads threads (128-bit stores when the array is 16-byte aligned)
__device__ __forceinline__ void zeroCountWords(uint32_t* words, uint32_t n_words, uint32_t thread, uint32_t n_threads) {
   if ((reinterpret_cast<uintptr_t>(words) & 15u) == 0) {
      uint4* vectors = reinterpret_cast<uint4*>(words);
      for (uint32_t i = thread; i < n_words / 4; i += n_threads) {
         vectors[i] = make_uint4(0u, 0u, 0u, 0u);
      }
      for (uint32_t i = (n_words & ~3u) + thread; i < n_words; i += n_threads) {
         words[i] = 0;
      }
   } else {
      for (uint32_t i = thread; i < n_words; i += n_threads) {
         words[i] = 0;
      }
   }
}

__device__ __forceinline__ uint32_t smemAddr(const void* ptr) {
   return static_cast<uint32_t>(__cvta_generic_to_shared(ptr));
}

__device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t count) {
   asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count));
}

__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes) {
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes)
                : "memory");
}

__device__ __forceinline__ void mbarArrive(uint64_t* bar) {
   asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smemAddr(bar)) : "memory");
}

__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity) {
   // try_wait suspends the warp in hardware (up to the time hint) instead of burning issue slots
   asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smemAddr(bar)),
      "r"(parity),
      "r"(0x989680u)
      : "memory"
   );
}

// 1-D bulk (TMA) copy global -> shared, completion signalled on `bar` (SASS: UBLKCP).
// dst, src 16-byte aligned; bytes a multiple of 16.
__device__ __forceinline__ void bulkLoad(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
   asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
         smemAddr(dst)
      ),
      "l"(src),
      "r"(bytes),
      "r"(smemAddr(bar))
      : "memory"
   );
}

__device__ __forceinline__ void fenceBarrierInit() {
   asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// mask of the rows [0, chunk_size) inside 64-bit word `word_index` of a tile
__device__ __forceinline__ uint64_t layoutWord(uint32_t chunk_size, uint32_t word_index) {
   const uint32_t first = word_index * 64;
   if (chunk_size >= first + 64) {
      return ~0ULL;
   }
   if (chunk_size <= first) {
      return 0ULL;
   }
   return (~0ULL) >> (64 - (chunk_size - first));
}

#endif  // __CUDACC__

}  // namespace silo
