// S3: the Mutations / AminoAcidMutations action on device.
//
// Replaces calculateMutationsPerPosition and its helpers
// (/root/reference/src/rhydb/query_engine/operators/mutations_node.cpp:39-288):
//   countActualFilteredMutations :153-189  -> containerAndCountKernel (K1): every stored container
//        of an active chunk AND the chunk's dense filter tile, popcount, one RED per container
//   countActualMutations :138-151 (full filter) -> containerCardinalityKernel (descriptor-only)
//   subtractFilteredNCounts :111-136, subtractStartAndEndNCounts :92-109,
//   subtractHorizontalBitmapCounts :51-61 -> coverageDiffKernel (K6): a single u32 difference array
//        D[start]++ / D[end]-- / N runs D[a]-- D[b]++
//   subtractCumulativeNsFromPositions :63-90 + accumulateFinalCounts :191-203
//        -> finalizeCountsKernel: prefix sum of D = rows covering p; the local-reference symbol's
//           count is  covered(p) - sum of the other symbols' counts  (uint32 modular, as the reference)
#include <algorithm>
#include <cstdlib>
#include <utility>
#include <cstring>

#include <chrono>
#include <memory>

#include <unistd.h>

#include "common.cuh"

namespace silo {

namespace {

// ---------------------------------------------------------------------------------------------
// prepare: zero the outputs, list the segments of the chunks that hold at least one filtered row
// ---------------------------------------------------------------------------------------------

// work_state[0] = number of work items, work_state[1] = claim counter of containerAndCountKernel; both
// are zero when a query starts (finalizeCountsKernel resets them).
constexpr int PREP_THREADS = 256;

// One launch in front of the container kernel instead of two memsets and a scan: every CTA zeroes a
// slice of the counts, and CTA c
// -- when chunk c holds a filtered row -- reserves room in the work list with one atomic and copies
// the chunk's segment records there. The order of the chunks in the list is whatever the atomics
// decide; the counts are sums, so the result does not depend on it, and the segments of one chunk
// stay contiguous (the container kernel reloads the filter tile only when the chunk changes).
__global__ void __launch_bounds__(PREP_THREADS) prepareQueryKernel(
   DevColumn column,
   const uint32_t* __restrict__ chunk_popcount,  // nullptr: no work list (full filter)
   uint32_t* __restrict__ work_state,
   DevSegment* __restrict__ work_items,
   uint32_t* __restrict__ counts,
   uint32_t counts_words
) {
   __shared__ uint32_t list_base;
   uint32_t first_segment = 0;
   uint32_t n_segments = 0;
   if (chunk_popcount != nullptr && blockIdx.x < column.n_chunks) {
      // three independent loads in flight
      const uint32_t popcount = chunk_popcount[blockIdx.x];
      first_segment = column.chunk_seg_begin[blockIdx.x];
      const uint32_t end_segment = column.chunk_seg_begin[blockIdx.x + 1];
      n_segments = popcount != 0 ? end_segment - first_segment : 0u;
      if (threadIdx.x == 0 && n_segments != 0) {
         list_base = atomicAdd(&work_state[0], n_segments);
      }
   }
   const uint32_t thread = blockIdx.x * PREP_THREADS + threadIdx.x;
   const uint32_t n_threads = gridDim.x * PREP_THREADS;
   zeroCountWords(counts, counts_words, thread, n_threads);
   __syncthreads();
   if (n_segments != 0) {
      const uint4* source = reinterpret_cast<const uint4*>(column.segments + first_segment);
      uint4* target = reinterpret_cast<uint4*>(work_items + list_base);
      for (uint32_t i = threadIdx.x; i < n_segments; i += PREP_THREADS) {
         target[i] = source[i];
      }
   }
}

// ---------------------------------------------------------------------------------------------
// K1: fused container AND filter-tile + popcount
//
// Persistent CTAs, ONE per SM. Warp 0 is the producer (one thread): it claims work items (segments) from a
// grid-wide counter and streams each segment's block [descriptors | payloads] into a ring of shared-memory
// stages with one 1-D bulk (TMA) copy that completes on an mbarrier. A stage holds pieces of ONE kind, at most
// W * P of them: each of the W consumer warps takes P pieces per stage visit. A warp pulls its pieces into
// REGISTERS, hands the stage back to the producer, and only then ANDs them with the chunk's filter tile held in
// shared memory, issuing one RED per piece with a non-zero count. The fixed cost of a stage visit (barrier wait,
// control word, ring bookkeeping, kind dispatch, two barrier arrivals) is paid once per P pieces: with one piece
// per visit it was 45 % of the kernel's instructions (profiles/r1_summary.md).
//
// Everything the consumers touch is addressed with 32-bit shared-memory addresses kept in registers
// (inline PTX loads): the kernel is bound by instruction issue, and the generic-pointer arithmetic of plain
// C++ costs as many instructions per piece as the lookups.
// ---------------------------------------------------------------------------------------------

constexpr uint32_t K1_STOP = 0xFFFFFFFFu;  // control.desc_count marker: no more work
constexpr uint32_t TILE32_WORDS = 2 * TILE_WORDS;
constexpr uint32_t TILE_BUFFER_BYTES = (TILE32_WORDS + 4) * 4;  // [2048] = the zero pad word

// per-stage control block: the first four words are written by the producer before it arms `full`
struct __align__(16) K1Control {
   uint32_t desc_count;  // K1_STOP: no more work
   uint32_t base4;       // slab offset (4-byte units) of the stage's first byte
   uint32_t flags;       // K1_TILE_SLOT
   uint32_t kind;        // stageClass() of every piece of the stage
   uint64_t full;   // the stage's bulk copies have landed
   uint64_t empty;  // every consumer warp has pulled its pieces of the stage into registers
   uint64_t done;   // ... and has finished the tile lookups for them
   uint64_t pad1;
};
static_assert(sizeof(K1Control) == 48);
constexpr uint32_t K1_CTRL_FULL = 16;
constexpr uint32_t K1_CTRL_EMPTY = 24;
constexpr uint32_t K1_CTRL_DONE = 32;
constexpr uint32_t K1_TILE_SLOT = 2;   // which of the two filter-tile buffers the stage reads

// dynamic shared memory: STAGES stage buffers of stageBytes(W, P), the control blocks, and a pad -- a warp pulls two
// whole 512-byte regions whatever the size of its piece, so the reads behind a short piece at the end of the last
// stage must stay inside the CTA's shared-memory window
__host__ __device__ constexpr uint32_t k1StageBytes(uint32_t warps, uint32_t pieces) {
   return warps * pieces * (1024u + 16u);
}
__host__ __device__ constexpr uint32_t k1DynamicBytes(uint32_t warps, uint32_t pieces, uint32_t stages) {
   return stages * k1StageBytes(warps, pieces) + stages * static_cast<uint32_t>(sizeof(K1Control)) + 1024u;
}

__device__ __forceinline__ uint32_t warpSum(uint32_t value) {
   return __reduce_add_sync(0xFFFFFFFFu, value);
}

// ---- shared memory by 32-bit address ----------------------------------------------------------
// Reads of stage data and of the filter tile. Not volatile: every address derives from words that were
// loaded after the stage's full barrier was observed (a volatile asm with a memory clobber).
__device__ __forceinline__ uint32_t lds32(uint32_t address) {
   uint32_t value;
   asm("ld.shared.u32 %0, [%1];" : "=r"(value) : "r"(address));
   return value;
}
__device__ __forceinline__ uint4 lds128(uint32_t address) {
   uint4 value;
   asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                : "=r"(value.x), "=r"(value.y), "=r"(value.z), "=r"(value.w)
                : "r"(address)
                : "memory");
   return value;
}
__device__ __forceinline__ void sts128(uint32_t address, const uint4& value) {
   asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(address), "r"(value.x), "r"(value.y), "r"(value.z), "r"(value.w)
                : "memory");
}
__device__ __forceinline__ void mbarWaitAt(uint32_t address, uint32_t parity) {
   asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(address),
      "r"(parity),
      "r"(0x989680u)
      : "memory"
   );
}
__device__ __forceinline__ void mbarArriveAt(uint32_t address) {
   asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(address) : "memory");
}
// lane 0 arrives, without a branch (the divergence bookkeeping of `if (lane == 0)` costs the consumer
// loop more issue slots than the arrive itself)
__device__ __forceinline__ void mbarArriveLane0(uint32_t address, uint32_t lane) {
   asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.eq.u32 p, %1, 0;\n"
      "@p mbarrier.arrive.shared::cta.b64 _, [%0];\n"
      "}\n" ::"r"(address),
      "r"(lane)
      : "memory"
   );
}
// lane 0 adds `value` to target[index] if value != 0, without a branch
__device__ __forceinline__ void redAddLane0(uint32_t* target, uint32_t index, uint32_t value, uint32_t lane) {
   asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .u64 a;\n"
      "setp.eq.u32 p, %3, 0;\n"
      "setp.ne.and.u32 p, %2, 0, p;\n"
      "mad.wide.u32 a, %1, 4, %0;\n"
      "@p red.global.add.u32 [a], %2;\n"
      "}\n" ::"l"(target),
      "r"(index), "r"(value), "r"(lane)
      : "memory"
   );
}
__device__ __forceinline__ void mbarExpectTxAt(uint32_t address, uint32_t bytes) {
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(address), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulkLoadAt(uint32_t destination, const void* source, uint32_t bytes, uint32_t barrier) {
   asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(destination),
                "l"(source), "r"(bytes), "r"(barrier)
                : "memory");
}

// The tile lookups are spelled for the ALU pipe (LOP3, SHF, LEA.HI: one warp instruction per two cycles and
// scheduler). mad.hi / mul.hi -- which round 1 used to fold a shift and an addition into one instruction "on the FMA
// pipe" -- are IMAD.HI in SASS and execute on the XU pipe of sm_100 at ONE warp instruction per EIGHT cycles: with 2.5
// of them per array value that pipe was 90 % busy and bounded the kernel (profiles/r2_summary.md). `Multipliers` is
// kept as an empty tag so that the call signatures stay.
struct Multipliers {};

// Two stored array values packed in one word -> the tile words of their rows, each shifted so that
// the row's bit sits in bit 31. A stored value is row ^ 31: its low five bits are 31 - (row & 31), the
// left shift that takes the row's bit to the top (funnel shifts use the low five bits of the amount).
__device__ __forceinline__ void pairTopBits(uint32_t tile_address, const Multipliers&, uint32_t pair, uint32_t& lo_top, uint32_t& hi_top) {
   const uint32_t lo_word = lds32(tile_address + ((pair & 0x0000FFE0u) >> 3));   // LOP3 + LEA.HI
   const uint32_t hi_word = lds32(tile_address + ((pair & 0xFFE00000u) >> 19));
   lo_top = __funnelshift_l(0u, lo_word, pair);        // lo_word << (pair & 31)
   hi_top = __funnelshift_l(0u, hi_word, pair >> 16);  // hi_word << ((pair >> 16) & 31)
}
__device__ __forceinline__ uint32_t addTopBit(uint32_t value, const Multipliers&, uint32_t accumulator) {
   return accumulator + (value >> 31);  // LEA.HI
}

// one KIND_RUNS_W entry -> rows of the run that are set in the tile
__device__ __forceinline__ uint32_t runEntryCount(uint32_t tile_address, const Multipliers&, uint32_t entry) {
   // tile32[entry >> 20]; the mask only matters for junk (lanes behind a partial region, descriptor slots without a
   // piece): it keeps their addresses aligned and inside the CTA's shared memory
   const uint32_t word = lds32(tile_address + ((entry & 0xFFF00000u) >> 18));
   const uint32_t from_first = __funnelshift_r(word, 0u, entry);        // word >> (entry & 31)
   return __popc(__funnelshift_l(0u, from_first, entry >> 5));          // << (32 - length)
}

// lookups of one array region held in registers: `count` values in ceil(count / 8) lanes. FULL: count == 256.
// In a partial region the slots behind the last value repeat it (pool.cu encodeArrayPiece). Every slot is looked
// up -- no test per slot --, and the repeats are taken out again: slot 7 of the last lane always holds the
// region's last value, so that lane subtracts (8 lanes - count) times its bit. Lanes behind the region hold
// other bytes of the stage; their lookups stay inside the tile (11-bit word index) and are dropped. Straight-line
// code: the regions of all the pieces of a stage visit form one basic block of independent lookups.
template <bool FULL>
__device__ __forceinline__ uint32_t arrayRegionCount(
   uint32_t tile_address,
   const Multipliers& k,
   const uint4& eight,
   uint32_t count,
   uint32_t lane
) {
   uint32_t t0, t1, t2, t3, t4, t5, t6, t7;
   pairTopBits(tile_address, k, eight.x, t0, t1);
   pairTopBits(tile_address, k, eight.y, t2, t3);
   pairTopBits(tile_address, k, eight.z, t4, t5);
   pairTopBits(tile_address, k, eight.w, t6, t7);
   // two accumulation chains
   const uint32_t even = addTopBit(t6, k, addTopBit(t4, k, addTopBit(t2, k, addTopBit(t0, k, 0u))));
   const uint32_t odd = addTopBit(t7, k, addTopBit(t5, k, addTopBit(t3, k, addTopBit(t1, k, 0u))));
   uint32_t local = even + odd;
   if (!FULL) {
      const uint32_t lanes = arrayRegionLanes(count);
      local -= lane + 1 == lanes ? (8u * lanes - count) * (t7 >> 31) : 0u;
      local = lane < lanes ? local : 0u;
   }
   return local;
}

// one run region held in registers (padding entries count nothing; lanes behind the region hold junk whose
// lookups stay inside the CTA's shared memory and are dropped)
template <bool FULL>
__device__ __forceinline__ uint32_t runsRegionCount(
   uint32_t tile_address,
   const Multipliers& k,
   const uint4& four,
   uint32_t count,
   uint32_t lane
) {
   const uint32_t local = runEntryCount(tile_address, k, four.x) + runEntryCount(tile_address, k, four.y) +
                          runEntryCount(tile_address, k, four.z) + runEntryCount(tile_address, k, four.w);
   return FULL || lane < runsRegionLanes(count) ? local : 0u;
}

// |piece AND tile| per lane for a piece whose (at most) two 512-byte regions sit in registers: lane L holds
// bytes [16 L, 16 L + 16) of each region. desc = the block descriptor's four words. KIND is the stage's kind;
// TWO: every piece of the stage has a full first region and a second one (pool.cu sorts them that way).
template <uint32_t KIND, bool TWO>
__device__ __forceinline__ uint32_t pieceLaneCount(
   const uint4& desc,
   const uint4& first,
   const uint4& second,
   uint32_t tile_address,  // shared: 2049 words ([2048] = 0)
   const Multipliers& k,
   uint32_t lane,
   uint32_t lane16
) {
   if (KIND == KIND_ARRAY_T) {
      const uint32_t n = (desc.z & 0xFFFFu) + 1u;
      if (TWO) {
         return arrayRegionCount<true>(tile_address, k, first, ARRAY_REGION_VALUES, lane) +
                arrayRegionCount<false>(tile_address, k, second, n - ARRAY_REGION_VALUES, lane);
      }
      return arrayRegionCount<false>(tile_address, k, first, n, lane);
   }
   if (KIND == KIND_RUNS_W) {
      const uint32_t n = desc.w;
      if (TWO) {
         return runsRegionCount<true>(tile_address, k, first, RUNS_REGION_ENTRIES, lane) +
                runsRegionCount<false>(tile_address, k, second, n - RUNS_REGION_ENTRIES, lane);
      }
      return runsRegionCount<false>(tile_address, k, first, n, lane);
   }
   // KIND_BITSET: 128 words: 64 vectors, two per lane (the mask keeps the window of a descriptor slot that holds no
   // piece -- junk -- inside the tile and aligned)
   const uint32_t window = tile_address + (desc.w & 0x380u) * 8u + lane16;
   const uint4 a = lds128(window);
   const uint4 b = lds128(window + 512);
   return __popc(first.x & a.x) + __popc(first.y & a.y) + __popc(first.z & a.z) + __popc(first.w & a.w) +
          __popc(second.x & b.x) + __popc(second.y & b.y) + __popc(second.z & b.z) + __popc(second.w & b.w);
}

// One stage visit of a consumer warp for the kinds whose pieces are pulled into registers: P pieces per warp,
// straight-line code (a piece the stage does not hold is read from offset 0 and counts nothing).
template <int W, int P, uint32_t KIND, bool TWO, int MODE>
__device__ __forceinline__ void consumeRegisterStage(
   const uint4& meta,
   uint32_t stage_address,
   uint32_t my_control,
   uint32_t tile,
   const Multipliers& k,
   uint32_t lane,
   uint32_t lane16,
   uint32_t cwarp,
   uint32_t* counts
) {
   uint4 desc[P];
   uint4 first[P];
   uint4 second[P];
#pragma unroll
   for (int p = 0; p < P; ++p) {
      const uint32_t index = p * W + cwarp;
      desc[p] = lds128(stage_address + index * 16);  // {counts index, offset4, packed, aux}: the block starts with its descriptors
      // (reads past a short piece stay inside the CTA's shared memory; those lanes are ignored)
      const uint32_t payload_address = stage_address + (index < meta.x ? (desc[p].y - meta.y) << 2 : 0u) + lane16;
      first[p] = lds128(payload_address);
      if (TWO) {
         second[p] = lds128(payload_address + 512);
      } else {
         second[p] = make_uint4(0u, 0u, 0u, 0u);
      }
   }
   __syncwarp();
   mbarArriveLane0(my_control + K1_CTRL_EMPTY, lane);  // the stage can be refilled while the lookups run
   uint32_t local[P];
#pragma unroll
   for (int p = 0; p < P; ++p) {
      if (MODE == 0) {
         local[p] = pieceLaneCount<KIND, TWO>(desc[p], first[p], second[p], tile, k, lane, lane16);
      } else {  // profiling: touch the payload only
         local[p] = (first[p].x ^ first[p].y ^ first[p].z ^ first[p].w ^ second[p].x ^ second[p].y ^ second[p].z ^ second[p].w) == 0x12345678u ? 1u : 0u;
      }
   }
#pragma unroll
   for (int p = 0; p < P; ++p) {
      const uint32_t total = warpSum(local[p]);
      redAddLane0(counts, desc[p].x, p * W + cwarp < meta.x ? total : 0u, lane);
   }
}

// KIND_WORDRANGE: the whole 32-row words [wa, wb) in the middle of long runs: the warp popcounts the
// tile words of each range directly (a range of n words costs n/32 iterations per lane, less than
// the lookups of an array piece of the same row count). The payload is read from the stage.
__device__ __forceinline__ uint32_t wordRangeCount(uint32_t entries, uint32_t payload_address, uint32_t tile_address, uint32_t lane) {
   uint32_t local = 0;
   for (uint32_t r = 0; r < entries; ++r) {
      const uint32_t range = lds32(payload_address + 4 * r);
      for (uint32_t word = (range & 0xFFFFu) + lane; word < (range >> 16); word += 32) {
         local += __popc(lds32(tile_address + 4 * word));
      }
   }
   return warpSum(local);
}

// MODE (profiling aid, SILO_K1_STREAM_ONLY): 0 = the product; 1 = consumers skip the intersection
// (what the bulk-copy pipeline alone can stream).
template <int W, int P, int STAGES, int MAXREG, int MODE>
__global__ void __launch_bounds__((W + 1) * 32, 1) __maxnreg__(MAXREG) containerAndCountKernel(
   DevColumn column,
   const uint64_t* __restrict__ filter_words,   // [n_chunks * 1024]
   uint32_t* __restrict__ work_state,           // [0] number of work items, [1] grid-wide claim counter (zero at launch)
   const DevSegment* __restrict__ work_items,   // the segment record of every work item (prepareQueryKernel)
   uint32_t* __restrict__ counts,               // [n_symbols * genome_length]
   uint32_t tail_factor,                        // the last tail_factor * gridDim.x work items are claimed just in time
   uint32_t* __restrict__ debug_times           // nullptr, or [4 * gridDim.x]: per CTA {start, producer end, consumers' end (ns), stages}
) {
   // Two tile buffers: the producer loads the next chunk's tile while stages of the current chunk
   // are still being consumed, so a chunk switch does not drain the pipeline.
   __shared__ __align__(16) uint32_t tile_buffers[2][TILE32_WORDS + 4];  // [2048] = zero pad word
   extern __shared__ __align__(128) uint8_t smem_raw[];
   constexpr uint32_t STAGE_BYTES = k1StageBytes(W, P);
   constexpr uint32_t CONTROL_OFFSET = STAGE_BYTES * STAGES;
   constexpr uint32_t CONTROL_BYTES = static_cast<uint32_t>(sizeof(K1Control));

   const uint32_t warp = threadIdx.x >> 5;
   const uint32_t lane = threadIdx.x & 31;
   const uint32_t ring_address = smemAddr(smem_raw);
   const uint32_t control_address = ring_address + CONTROL_OFFSET;
   const uint32_t tile_address0 = smemAddr(tile_buffers[0]);

   if (threadIdx.x == 0) {
      K1Control* control = reinterpret_cast<K1Control*>(smem_raw + CONTROL_OFFSET);
      for (int s = 0; s < STAGES; ++s) {
         mbarInit(&control[s].full, 1);
         mbarInit(&control[s].empty, W);
         mbarInit(&control[s].done, W);
      }
      fenceBarrierInit();
   }
   if (threadIdx.x < 8) {
      tile_buffers[threadIdx.x >> 2][TILE32_WORDS + (threadIdx.x & 3)] = 0;
   }
   __syncthreads();
   // everything above is this CTA's own shared memory; the work list, the filter tiles and the zeroed counts come from
   // the kernels in front
   gridDependencyWait();
   gridDependencyLaunch();

   if (warp == 0) {
      // ---------------- producer: ONE thread ------------------------------------------------------
      // Lane 0 alone claims work items (segments) from the grid-wide counter, fetches their 16-byte records and
      // issues one bulk copy per ring stage (a warp-wide producer queues its shuffles, ballots and barrier tests
      // behind the consumers' shared-memory loads, ~40 cycles per instruction; a single thread needs no
      // cross-lane instruction at all). The first two items of every CTA are fixed (no atomic in front of the
      // first copies). A stage holds ~60 KiB, i.e. ~2 us of consumer work, and the ring keeps two such stages
      // waiting behind the one being consumed: the latency of a claim (atomic + record load) hides behind them.
      if (lane != 0) {
         return;
      }
      if (debug_times != nullptr) {
         asm volatile("mov.u32 %0, %%globaltimer_lo;" : "=r"(debug_times[4 * blockIdx.x]));
      }
      uint32_t* const work_counter = work_state + 1;
      const uint32_t static_items = 2 * gridDim.x;
      const uint32_t total = work_state[0];
      // the last tail_factor items per CTA are claimed just in time (see below)
      const uint32_t tail_items = tail_factor * gridDim.x;
      const uint32_t tail_begin = total > tail_items ? total - tail_items : 0u;
      const uint4* const records = reinterpret_cast<const uint4*>(work_items);
      auto load = [&](uint32_t index) {
         // (records behind the end of the list are stale or uninitialised and never used)
         return index < total ? records[index] : make_uint4(0u, 0u, 0u, 0u);
      };
      auto claim = [&]() { return atomicAdd(work_counter, 1u) + static_items; };
      // The item being issued, ONE item ahead with its record fetched, and one claim in flight behind it: neither the
      // atomic nor the record load is waited for in the iteration that issues it. Near the end of the list nothing is
      // held ahead: an item is claimed only once the stage it goes into is free. What a CTA owns beyond its ring when
      // the list runs out is what makes the CTAs finish at different times (measured with two items held ahead
      // everywhere: the producers ran out of work over 6 us of a 55 us kernel; with this rule 2.5 us).
      uint32_t current_index = blockIdx.x;
      uint4 current = load(current_index);
      bool have_current = true;
      uint32_t next_index = gridDim.x + blockIdx.x;
      uint4 next = load(next_index);
      bool have_next = true;
      uint32_t ahead_index = 0;
      bool have_ahead = false;
      if (next_index < tail_begin) {
         ahead_index = claim();
         have_ahead = true;
      }
      uint32_t tile_chunk = 0xFFFFFFFFu;  // chunk whose tile the latest stage reads
      uint32_t tile_slot = 1;             // ... and the buffer it sits in
      uint32_t tile_first_stage = 0;      // first stage that reads it
      uint32_t it = 0;                    // stages issued so far
      uint32_t stage = 0;                 // it % STAGES
      uint32_t round = 0;                 // it / STAGES
      for (;;) {
         const uint32_t my_control = control_address + stage * CONTROL_BYTES;
         const uint32_t my_ring = ring_address + stage * STAGE_BYTES;
         if (round > 0) {
            mbarWaitAt(my_control + K1_CTRL_EMPTY, (round - 1) & 1u);  // => every stage <= it - STAGES is pulled
         }
         if (!have_current) {
            current_index = claim();
            current = load(current_index);
         }
         if (current_index >= total) {
            break;  // (the stage at `stage` is free: the stop marker goes there)
         }
         // ---- one ring stage: {payload_offset16, block bytes | class << 24, -, chunk | pieces << 16} ----
         const uint32_t chunk = current.w & 0xFFFFu;
         const bool new_tile = chunk != tile_chunk;
         if (new_tile) {
            // The new tile goes into the OTHER buffer, last read by the lookups of the stages before
            // tile_first_stage. A warp hands a stage back BEFORE it does the lookups, so the stage's empty
            // barrier says nothing about the tile: wait for the done barriers. (Stages below it - STAGES
            // are implied: a warp pulls stage q + STAGES only after it has finished stage q, and the
            // empty wait above covered it - STAGES.)
            for (uint32_t prev = it >= STAGES ? it - STAGES : 0; prev < tile_first_stage; ++prev) {
               mbarWaitAt(control_address + (prev % STAGES) * CONTROL_BYTES + K1_CTRL_DONE, (prev / STAGES) & 1u);
            }
            tile_slot ^= 1u;
            tile_first_stage = it;
            tile_chunk = chunk;
         }
         const uint32_t block_bytes = current.y & 0xFFFFFFu;
         sts128(my_control, make_uint4(current.w >> 16, current.x << 2, tile_slot != 0 ? K1_TILE_SLOT : 0u, current.y >> 24));
         mbarExpectTxAt(my_control + K1_CTRL_FULL, block_bytes + (new_tile ? TILE_BYTES : 0u));
         if (new_tile) {
            bulkLoadAt(tile_address0 + tile_slot * TILE_BUFFER_BYTES, filter_words + static_cast<size_t>(chunk) * TILE_WORDS, TILE_BYTES, my_control + K1_CTRL_FULL);
         }
         // the segment's block [descriptors | payloads] in one copy
         bulkLoadAt(my_ring, column.payload + (static_cast<uint64_t>(current.x) << 4), block_bytes, my_control + K1_CTRL_FULL);
         ++it;
         const bool wrap = stage == STAGES - 1;
         stage = wrap ? 0u : stage + 1u;
         round += wrap ? 1u : 0u;
         // ---- the next item ----
         have_current = have_next;
         current = next;
         current_index = next_index;
         have_next = have_ahead;
         if (have_ahead) {
            next_index = ahead_index;  // (the atomic was issued one stage ago)
            next = load(next_index);
            have_ahead = false;
            if (next_index < tail_begin) {
               ahead_index = claim();
               have_ahead = true;
            }
         }
      }
      if (debug_times != nullptr) {
         asm volatile("mov.u32 %0, %%globaltimer_lo;" : "=r"(debug_times[4 * blockIdx.x + 1]));
         debug_times[4 * blockIdx.x + 3] = it;
      }
      // tell the consumers that nothing follows (the loop left with the stage at `stage` free)
      const uint32_t stop_control = control_address + stage * CONTROL_BYTES;
      sts128(stop_control, make_uint4(K1_STOP, 0u, 0u, 0u));
      mbarArriveAt(stop_control + K1_CTRL_FULL);
      return;
   }

   // ---------------- consumers: W warps ------------------------------------------------------------
   // Warp w takes the pieces w, w + W, ... of the stage (all of one kind). It pulls them (descriptor + at most two
   // 512-byte regions each) into registers, hands the stage back to the producer at once, and only then does the
   // tile lookups: the ring's stages are in flight again while the arithmetic runs.
   const uint32_t cwarp = warp - 1;
   const uint32_t lane16 = lane * 16;
   const Multipliers k{};
   // base addresses as opaque register values (the compiler would otherwise re-derive the shared
   // window from special registers at every use)
   uint32_t ring_base = ring_address;
   uint32_t control_base = control_address;
   uint32_t tile_base = tile_address0;
   asm volatile("" : "+r"(ring_base), "+r"(control_base), "+r"(tile_base));
   uint32_t stage = 0;  // ring position and phase of the next visit
   uint32_t parity = 0;
   for (;;) {
      const uint32_t stage_address = ring_base + stage * STAGE_BYTES;
      const uint32_t my_control = control_base + stage * CONTROL_BYTES;
      mbarWaitAt(my_control + K1_CTRL_FULL, parity);
      const uint4 meta = lds128(my_control);  // {pieces, base4, flags, kind}
      const uint32_t desc_count = meta.x;
      if (desc_count == K1_STOP) {
         if (debug_times != nullptr && cwarp == 0 && lane == 0) {
            asm volatile("mov.u32 %0, %%globaltimer_lo;" : "=r"(debug_times[4 * blockIdx.x + 2]));
         }
         break;
      }
      // next ring position (no division: the stage count need not be a power of two)
      const bool wrap = stage == STAGES - 1;
      stage = wrap ? 0u : stage + 1u;
      parity ^= wrap ? 1u : 0u;
      const uint32_t tile = tile_base + ((meta.z & K1_TILE_SLOT) != 0 ? TILE_BUFFER_BYTES : 0u);
      const uint32_t kind = meta.w;
      if (kind == stageClass(KIND_INLINE, false)) {
         // descriptor-only pieces (arrays of one or two values): one piece per LANE, no warp reduction
         uint4 desc[P];
#pragma unroll
         for (int p = 0; p < P; ++p) {
            const uint32_t index = (p * W + cwarp) * 32 + lane;
            desc[p] = make_uint4(0u, 0u, 0u, 0u);
            if (index < desc_count) {
               desc[p] = lds128(stage_address + index * 16);
            }
         }
         __syncwarp();
         mbarArriveLane0(my_control + K1_CTRL_EMPTY, lane);
#pragma unroll
         for (int p = 0; p < P; ++p) {
            const uint32_t index = (p * W + cwarp) * 32 + lane;
            if (index < desc_count && MODE == 0) {
               const uint32_t first_value = desc[p].w & 0xFFFFu;
               uint32_t count = (lds32(tile + ((first_value >> 5) << 2)) >> (first_value & 31u)) & 1u;
               if ((desc[p].z & 0xFFFFu) != 0) {
                  const uint32_t second_value = desc[p].w >> 16;
                  count += (lds32(tile + ((second_value >> 5) << 2)) >> (second_value & 31u)) & 1u;
               }
               if (count != 0) {
                  atomicAdd(&counts[desc[p].x], count);
               }
            }
         }
         __syncwarp();
      } else if ((kind >> 1) == KIND_WORDRANGE) {  // rare; reads the stage while it works
#pragma unroll 1
         for (int p = 0; p < P; ++p) {
            const uint32_t index = p * W + cwarp;
            if (index < desc_count) {
               const uint4 desc = lds128(stage_address + index * 16);
               const uint32_t count = wordRangeCount(desc.w, stage_address + ((desc.y - meta.y) << 2), tile, lane);
               redAddLane0(counts, desc.x, MODE == 0 ? count : 0u, lane);
            }
         }
         __syncwarp();
         mbarArriveLane0(my_control + K1_CTRL_EMPTY, lane);
      } else if (kind == stageClass(KIND_ARRAY_T, true)) {
         consumeRegisterStage<W, P, KIND_ARRAY_T, true, MODE>(meta, stage_address, my_control, tile, k, lane, lane16, cwarp, counts);
      } else if (kind == stageClass(KIND_ARRAY_T, false)) {
         consumeRegisterStage<W, P, KIND_ARRAY_T, false, MODE>(meta, stage_address, my_control, tile, k, lane, lane16, cwarp, counts);
      } else if (kind == stageClass(KIND_RUNS_W, true)) {
         consumeRegisterStage<W, P, KIND_RUNS_W, true, MODE>(meta, stage_address, my_control, tile, k, lane, lane16, cwarp, counts);
      } else if (kind == stageClass(KIND_RUNS_W, false)) {
         consumeRegisterStage<W, P, KIND_RUNS_W, false, MODE>(meta, stage_address, my_control, tile, k, lane, lane16, cwarp, counts);
      } else {
         consumeRegisterStage<W, P, KIND_BITSET, true, MODE>(meta, stage_address, my_control, tile, k, lane, lane16, cwarp, counts);
      }
      mbarArriveLane0(my_control + K1_CTRL_DONE, lane);
   }
}

// ---------------------------------------------------------------------------------------------
// full filter (cardinality == numRows, mutations_node.cpp:239-266): stored cardinalities only
// ---------------------------------------------------------------------------------------------

__global__ void containerCardinalityKernel(DevColumn column, uint32_t* __restrict__ counts) {
   const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
   for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < column.n_containers; i += stride) {
      const DevContainer desc = column.containers[i];
      atomicAdd(&counts[desc.symbol() * column.genome_length + desc.position], desc.cardinality());
   }
}

// ---------------------------------------------------------------------------------------------
// K6: coverage of the filtered rows as one difference array
// ---------------------------------------------------------------------------------------------

constexpr int K6_THREADS = 256;
constexpr int K6_SLICES = 32;  // CTAs per chunk
constexpr uint32_t K6_ROWS_PER_WARP = 65536 / K6_SLICES / (K6_THREADS / 32);  // 256 rows = 8 filter words
constexpr uint32_t K6_UNROLL = 4;

struct DiffArrays {
   uint32_t* diff;          // [genome_length + 1]
   uint32_t* block_totals;  // [diffPadded / 256]
   __device__ __forceinline__ void add(uint32_t key, uint32_t amount) const {
      atomicAdd(&diff[key], amount);
      atomicAdd(&block_totals[key / DIFF_BLOCK], amount);
   }
};

struct PendingAdd {
   uint32_t key;
   uint32_t count;
};

__device__ __forceinline__ void flushPending(PendingAdd& pending, const DiffArrays& out, uint32_t lane, bool negate) {
   if (pending.count != 0 && lane == 0) {
      out.add(pending.key, negate ? 0u - pending.count : pending.count);
   }
   pending.count = 0;
}

// adds `+1` (or -1) at diff[key] for every active lane, aggregating equal keys: a warp-uniform key
// is accumulated in registers across iterations (the common case: full-length genomes, sorted reads)
__device__ __forceinline__ void aggregateAdd(
   PendingAdd& pending,
   const DiffArrays& out,
   uint32_t key,
   bool active,
   uint32_t active_mask,
   uint32_t lane,
   bool negate
) {
   const uint32_t leader = __ffs(active_mask) - 1;
   const uint32_t leader_key = __shfl_sync(0xFFFFFFFFu, key, leader);
   const uint32_t same_mask = __ballot_sync(0xFFFFFFFFu, active && key == leader_key);
   if (same_mask == active_mask) {
      if (pending.count != 0 && pending.key != leader_key) {
         flushPending(pending, out, lane, negate);
      }
      pending.key = leader_key;
      pending.count += __popc(active_mask);
      return;
   }
   // mixed keys in this warp iteration: one RED per distinct key
   const uint32_t peers = __match_any_sync(0xFFFFFFFFu, active ? key : 0xFFFFFFFFu);
   if (active && lane == static_cast<uint32_t>(__ffs(peers) - 1)) {
      const uint32_t amount = __popc(peers);
      out.add(key, negate ? 0u - amount : amount);
   }
}

__global__ void __launch_bounds__(K6_THREADS) coverageDiffKernel(
   DevColumn column,
   const uint64_t* __restrict__ filter_words,
   const uint32_t* __restrict__ chunk_popcount,
   uint32_t* __restrict__ diff_scratch  // [diffWords(genome_length)], zeroed
) {
   const uint32_t chunk = blockIdx.x / K6_SLICES;
   const uint32_t slice = blockIdx.x % K6_SLICES;
   if (chunk_popcount[chunk] == 0) {
      return;
   }
   const DiffArrays out{diff_scratch, diff_scratch + diffPadded(column.genome_length)};
   const uint32_t lane = threadIdx.x & 31;
   const uint32_t warp = threadIdx.x >> 5;
   const uint32_t* tile32 = reinterpret_cast<const uint32_t*>(filter_words + static_cast<size_t>(chunk) * TILE_WORDS);
   const uint2* rows = column.start_end + column.chunk_row_begin[chunk];

   // A filter evaluated inside the same query may hold ids outside the row layout (the query then fails with
   // SILO_E_OUT_OF_LAYOUT once its error flag comes back): such bits are masked here, there is no row data behind them.
   // The warp's eight filter words are fetched at once, and the (start, end) loads of K6_UNROLL words are in
   // flight together: the kernel is bound by global-memory latency, not bandwidth.
   const uint32_t chunk_rows = column.chunk_row_begin[chunk + 1] - column.chunk_row_begin[chunk];
   const uint32_t warp_first = (slice * (K6_THREADS / 32) + warp) * K6_ROWS_PER_WARP;
   uint32_t my_word = lane < K6_ROWS_PER_WARP / 32 ? tile32[(warp_first >> 5) + lane] : 0u;
   const uint32_t word_first = warp_first + lane * 32;
   if (word_first + 32 > chunk_rows) {
      my_word &= word_first >= chunk_rows ? 0u : (1u << (chunk_rows - word_first)) - 1u;
   }
   PendingAdd pending_start{0, 0};
   PendingAdd pending_end{0, 0};
   for (uint32_t group = 0; group < K6_ROWS_PER_WARP / 32; group += K6_UNROLL) {
      uint32_t bits[K6_UNROLL];
      uint2 range[K6_UNROLL];
#pragma unroll
      for (uint32_t j = 0; j < K6_UNROLL; ++j) {
         bits[j] = __shfl_sync(0xFFFFFFFFu, my_word, group + j);
         range[j] = make_uint2(0, 0);
         if (((bits[j] >> lane) & 1u) != 0) {
            range[j] = rows[warp_first + (group + j) * 32 + lane];
         }
      }
#pragma unroll
      for (uint32_t j = 0; j < K6_UNROLL; ++j) {
         if (bits[j] != 0) {
            const bool active = ((bits[j] >> lane) & 1u) != 0;
            aggregateAdd(pending_start, out, range[j].x, active, bits[j], lane, false);
            aggregateAdd(pending_end, out, range[j].y, active, bits[j], lane, true);
         }
      }
   }
   flushPending(pending_start, out, lane, false);
   flushPending(pending_end, out, lane, true);

   // N positions inside the covered range, stored as runs per row
   const uint32_t missing_begin = column.chunk_missing_begin[chunk];
   const uint32_t missing_end = column.chunk_missing_begin[chunk + 1];
   for (uint32_t i = missing_begin + slice * K6_THREADS + threadIdx.x; i < missing_end; i += K6_SLICES * K6_THREADS) {
      const uint32_t row = column.missing_row[i];
      if (((tile32[row >> 5] >> (row & 31)) & 1u) == 0) {
         continue;
      }
      for (uint64_t run = column.missing_offsets[i]; run < column.missing_offsets[i + 1]; ++run) {
         const uint2 r = column.missing_runs[run];
         out.add(r.x, 0u - 1u);
         out.add(r.y, 1u);
      }
   }
}

// ---------------------------------------------------------------------------------------------
// finalize: covered(p) = prefix sum of diff; counts[local_ref[p]][p] = covered(p) - sum(others)
// ---------------------------------------------------------------------------------------------

// One thread per position. The prefix in front of a 256-position block comes from the block totals
// that coverageDiffKernel accumulated next to the difference array.
constexpr int FIN_THREADS = DIFF_BLOCK;

// The output pass of addMutationsToOutput (mutations_node.cpp:307-363) for one position, on request:
// which (position, symbol) rows the action emits. `hits` is PAGE-LOCKED HOST memory: the kernel stores the
// tuples (from hits[1]) straight into it over PCIe -- a few hundred 16-byte posted writes --, and the last
// block to finish writes the header hits[0] = {number of tuples, the filter's error flag, the filter's
// cardinality (low, high word)}, so a fused query needs no device-to-host copy at all. The running tuple
// count lives in work_state[3] (zero between queries).
// ---------------------------------------------------------------------------------------------
// Row-partitioned tables: one process per GPU, every rank holds the chunks of its shard. The per-rank
// counts are plain addends, and only ONE rank (the root, rank 0) needs the sums. Instead of a collective kernel that
// competes with the container kernel for SMs, the finalize kernel of every rank STORES its rows of the valid mutation
// symbols straight into the root's gather area over NVLink (peer memory mapped with CUDA IPC) as tagged words, and the
// root's own finalize kernel (FIN_COLLECT; or shardCollectKernel when the root collects in a call of its own) adds the
// `world` arrays and runs the output pass. Flow control: a gather slot is reused every SHARD_SLOTS queries; a rank writes
// slot s for its query q only after the root released the slot's previous use (the collecting kernel's last block stores
// the generation into every rank's `released[s]`), so ranks may run at most SHARD_SLOTS - 1 queries ahead of the root. Every wait in a kernel is bounded (SHARD_SPIN_LIMIT): a missing peer becomes an error
// flag in the result, not a hung GPU.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t SHARD_MAX_WORLD = 16;
constexpr uint32_t SHARD_SLOTS = 4;
constexpr long long SHARD_SPIN_LIMIT = 4000000000LL;  // cycles (~2 s)

struct ShardBlockHeader {  // the start of every rank's exported block
   uint32_t released[SHARD_SLOTS];  // generation of the slot's last consumed use; written by the root's collecting kernel
   // root only, tagged words like the rows (see storeTagged): the ranks' filter cardinalities (two halves) and error flags
   uint2 cardinality_low[SHARD_SLOTS][SHARD_MAX_WORLD];
   uint2 cardinality_high[SHARD_SLOTS][SHARD_MAX_WORLD];
   uint2 error[SHARD_SLOTS][SHARD_MAX_WORLD];
   uint32_t collect_error;  // root only: a block of the collecting kernel gave up waiting (cleared by its last block)
   // Device-side query counters (local use): which slot and generation a launch works on is read from here, not from
   // kernel parameters, so that a captured CUDA graph of sharded queries can be replayed.
   uint32_t queries_pushed;     // sharded queries this rank's finalize kernels have completed
   uint32_t queries_collected;  // root: queries its collecting kernels have completed
};
constexpr size_t SHARD_HEADER_BYTES = (sizeof(ShardBlockHeader) + 255) / 256 * 256;

struct ShardPush {  // what the finalize kernel of a sharded query needs (all zero: not sharded)
   uint8_t* root_block = nullptr;           // the root's block (header + gather area [slot][rank][n_valid][genome_length]) as mapped here
   ShardBlockHeader* own_header = nullptr;  // this rank's block: released[], queries_pushed
   uint32_t rank = 0;
   uint32_t world = 0;
   uint32_t n_valid = 0;
   uint32_t use_fixed_cardinality = 0;
   uint64_t valid_mask = 0;
   unsigned long long fixed_cardinality = 0;          // filter == all rows
   unsigned long long* filter_scalars = nullptr;      // else: {cardinality, -, error flag} of the query's filter; reset here
   // root only, "collect here": the finalize kernel itself waits for the other ranks' rows, adds this rank's counts from
   // its registers, runs the output pass over the sums and hands the slot back -- no store of the root's own rows, no
   // second kernel behind it (peers == nullptr: the root's rows go to the slot like everybody else's and
   // shardCollectKernel sums them later)
   ShardBlockHeader* const* peers = nullptr;  // [world] every rank's block as mapped on the root
   uint32_t* summed_out = nullptr;            // optional: [n_symbols][genome_length], the rows of the valid symbols are written
   unsigned long long* debug_times = nullptr; // SILO_SHARD_DEBUG: [blocks][4] globaltimer: start, wait over, rows done, scalars sent
};

__device__ __forceinline__ void shardDebugStamp(const ShardPush& push, uint32_t index) {
   if (push.debug_times != nullptr) {  // (called by one thread of the block)
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      push.debug_times[4 * blockIdx.x + index] = now;
   }
}

__device__ __forceinline__ uint32_t loadAcquireSystem(const uint32_t* address) {
   uint32_t value;
   asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(value) : "l"(address) : "memory");
   return value;
}
__device__ __forceinline__ void storeReleaseSystem(uint32_t* address, uint32_t value) {
   asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(address), "r"(value) : "memory");
}
// spins until *address >= target (wrap-safe); false after SHARD_SPIN_LIMIT cycles
__device__ __forceinline__ bool waitForAtLeast(const uint32_t* address, uint32_t target) {
   const long long begin = clock64();
   while (static_cast<int32_t>(*reinterpret_cast<const volatile uint32_t*>(address) - target) < 0) {
      if (clock64() - begin > SHARD_SPIN_LIMIT) {
         return false;
      }
      __nanosleep(100);
   }
   return static_cast<int32_t>(loadAcquireSystem(address) - target) >= 0;
}

// The gather area and the per-rank scalars are made of TAGGED words {value, tag}, tag = the group's query number + 1,
// written with ONE 8-byte store each: a reader that sees the tag of its query has the value. No fence and no arrival
// counter on the writer's side (a system-scope fence behind stores over NVLink is a round trip of ~2-3 us on the
// critical path of every rank), no wait-for-everybody on the reader's side: the root polls the words it needs.
__device__ __forceinline__ void storeTagged(uint2* address, uint32_t value, uint32_t tag) {
   asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(address), "r"(value), "r"(tag) : "memory");
}
__device__ __forceinline__ uint2 loadTagged(const uint2* address) {
   uint2 word;
   asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(word.x), "=r"(word.y) : "l"(address) : "memory");
   return word;
}
// re-reads the word until it carries `tag` (the caller's first load did not); gives up after SHARD_SPIN_LIMIT cycles and
// returns the word as it is then. (Not inlined: the callers keep dozens of words in registers.)
__device__ __noinline__ uint2 pollTagged(const uint2* address, uint32_t tag) {
   const long long begin = clock64();
   uint2 word = loadTagged(address);
   while (word.y != tag && clock64() - begin <= SHARD_SPIN_LIMIT) {
      __nanosleep(200);
      word = loadTagged(address);
   }
   return word;
}

// the root's last block, one whole warp: the filter cardinalities and error flags that block 0 of every rank's finalize
// kernel stored, one rank per lane (the three loads of all ranks in flight together), summed / ORed over the warp
__device__ __forceinline__ void warpSumRankScalars(
   ShardBlockHeader* root_header, uint32_t slot, uint32_t world, uint32_t tag, uint32_t lane, unsigned long long& cardinality, uint32_t& error
) {
   unsigned long long mine = 0;
   uint32_t my_error = 0;
   if (lane < world) {
      uint2 low = loadTagged(&root_header->cardinality_low[slot][lane]);
      uint2 high = loadTagged(&root_header->cardinality_high[slot][lane]);
      uint2 flag = loadTagged(&root_header->error[slot][lane]);
      if (low.y != tag || high.y != tag || flag.y != tag) {
         low = pollTagged(&root_header->cardinality_low[slot][lane], tag);
         high = pollTagged(&root_header->cardinality_high[slot][lane], tag);
         flag = pollTagged(&root_header->error[slot][lane], tag);
      }
      if (low.y != tag || high.y != tag || flag.y != tag) {
         my_error = 2u;
      } else {
         mine = low.x | (static_cast<unsigned long long>(high.x) << 32);
         my_error = flag.x;
      }
   }
   for (int offset = 16; offset > 0; offset >>= 1) {
      mine += __shfl_xor_sync(0xFFFFFFFFu, mine, offset);
   }
   cardinality += mine;
   error |= __reduce_or_sync(0xFFFFFFFFu, my_error);
}

struct HitRequest {
   silo_mutation_hit* hits = nullptr;
   uint32_t capacity = 0;
   unsigned long long* filter_scalars = nullptr;  // {cardinality, -, error flag (u32), -} of a filter evaluated inside
                                                  // the call: reported in the header and zeroed for the next query
   uint64_t valid_mask = 0;  // SymbolType::VALID_MUTATION_SYMBOLS
   double min_proportion = 0;
   uint32_t keep_scalars = 0;  // report the filter's scalars but leave them for the next column of the same query
};

// Grid: diffPadded(genome_length) / 256 blocks (the difference array has genome_length + 1 entries) of
// FIN_GROUPS x 256 threads: thread (g, i) works on position i of the block and on the symbols g, g + 4, g + 8, ...
// (at most eight loads per thread, all in flight together).
//
// The kernel runs once per query, right behind a container kernel that has pushed everything else out of the
// instruction caches: what it costs is mostly the FETCH of its own code (an earlier version that unrolled the
// 32 possible symbols per thread, plus 56 loads of the other ranks' rows, was 6,152 instructions = 98 KB and took
// ~8 us for work that needs two memory round trips). So: parallelism over threads instead of unrolled code, one
// instantiation per mode, loops kept as loops.
//
// It leaves the difference array, its block totals and the work-list state all-zero for the next query: every
// thread clears the element it read, the last block to finish clears the totals.
constexpr int FIN_GROUPS = 4;
constexpr int FIN_BLOCK_THREADS = FIN_GROUPS * FIN_THREADS;
constexpr int FIN_SYMBOLS_PER_THREAD = 32 / FIN_GROUPS;
enum FinalizeMode : int {
   FIN_COUNTS = 0,   // the counts only
   FIN_OUTPUT = 1,   // + the output pass (addMutationsToOutput) over this table's counts
   FIN_PUSH = 2,     // sharded query: + this rank's rows of the valid symbols into the root's gather area
   FIN_COLLECT = 3,  // sharded query, root: + the other ranks' rows added, the output pass over the sums
};

template <int MODE>
__global__ void __launch_bounds__(FIN_BLOCK_THREADS) finalizeCountsKernel(
   DevColumn column,
   uint32_t* __restrict__ diff_scratch,
   uint32_t* __restrict__ counts,
   uint32_t* __restrict__ work_state,
   HitRequest request,
   ShardPush push
) {
   constexpr bool SHARDED = MODE == FIN_PUSH || MODE == FIN_COLLECT;
   constexpr bool USE_ROWS = MODE == FIN_OUTPUT || MODE == FIN_COLLECT;  // the valid symbols' (summed) counts per position in shared memory
   __shared__ uint32_t warp_totals[FIN_THREADS / 32];
   __shared__ uint32_t block_offset;
   __shared__ uint32_t slot_is_free;  // sharded query: the wait of this block succeeded (else: timed out)
   __shared__ uint32_t part_others[FIN_GROUPS][FIN_THREADS];
   __shared__ uint32_t reference_counts[FIN_THREADS];
   __shared__ uint32_t row_sum[USE_ROWS ? 32 : 1][FIN_THREADS];  // [row r = the r-th valid symbol][position]
   const uint32_t genome_length = column.genome_length;
   uint32_t* diff = diff_scratch;
   uint32_t* block_totals = diff_scratch + diffPadded(genome_length);
   const uint32_t i = threadIdx.x % FIN_THREADS;
   const uint32_t group = threadIdx.x / FIN_THREADS;  // (uniform per warp)
   const uint32_t lane = threadIdx.x & 31;
   const uint32_t warp = i >> 5;
   const uint32_t p = blockIdx.x * FIN_THREADS + i;
   const bool in_range = p < genome_length;
   const uint64_t valid_mask = SHARDED ? push.valid_mask : request.valid_mask;
   const uint32_t n_valid = __popcll(valid_mask);
   // Sharded query: every thread reads the group's query counter itself (the gather slot and the tag of this query; no
   // barrier between that and the loads that need it). The block's one wait is started first, by the last thread, so that
   // it overlaps everything below: a rank that stores its rows waits until the root has released the slot's previous use
   // (a root that collects here has nothing to wait for: it polls the other ranks' rows). Every block reads the counter
   // before it counts itself in at the end; the last block increments it after that.
   const uint32_t query_index = SHARDED ? *reinterpret_cast<volatile uint32_t*>(&push.own_header->queries_pushed) : 0u;
   const uint32_t shard_slot = query_index % SHARD_SLOTS;
   const uint32_t tag = query_index + 1;
   if (SHARDED && threadIdx.x == FIN_BLOCK_THREADS - 1) {
      shardDebugStamp(push, 0);
      if (MODE == FIN_COLLECT) {
         // every earlier query of the group must have been collected: the slots are summed in order
         slot_is_free = *reinterpret_cast<volatile uint32_t*>(&reinterpret_cast<ShardBlockHeader*>(push.root_block)->queries_collected) == query_index ? 1u : 0u;
      } else {
         slot_is_free = waitForAtLeast(&push.own_header->released[shard_slot], query_index / SHARD_SLOTS) ? 1u : 0u;
      }
      shardDebugStamp(push, 1);
   }
   const uint32_t reference_symbol = in_range ? column.local_reference[p] : 0u;  // (static data: may be read before the wait)
   // the query counter and the released flag above are written by EARLIER finalize kernels of this stream and by the
   // root (transitively complete / another device); everything below is written by the kernels right in front
   gridDependencyWait();
   gridDependencyLaunch();
   // every global load up front
   const uint32_t mine = group == 0 ? diff[p] : 0u;  // (the array is padded to whole blocks)
   uint32_t values[FIN_SYMBOLS_PER_THREAD];
#pragma unroll
   for (uint32_t k = 0; k < FIN_SYMBOLS_PER_THREAD; ++k) {
      const uint32_t symbol = group + k * FIN_GROUPS;
      // (the reference symbol's row too, dropped below: the loads must not wait for local_reference[p])
      values[k] = in_range && symbol < column.n_symbols ? counts[symbol * genome_length + p] : 0u;
   }
   bool peer_timed_out = false;
   if (MODE == FIN_COLLECT && in_range) {
      // The other ranks' tagged rows, beside the loads above: thread group g sums the rows g, g + 4 (one pass), g + 8,
      // g + 12 (the next), ... of this position over the ranks, up to 2 x 8 loads in flight; a word that has not landed
      // yet is polled. (Loops, not unrolled items: see above.)
      const size_t rank_stride = static_cast<size_t>(n_valid) * genome_length;
      const uint2* const slot_rows = reinterpret_cast<const uint2*>(push.root_block + SHARD_HEADER_BYTES) +
                                     static_cast<size_t>(shard_slot) * push.world * rank_stride + p;
      constexpr uint32_t IN_FLIGHT = 8;
#pragma unroll 1
      for (uint32_t row = group; row < n_valid; row += 2 * FIN_GROUPS) {
         const bool second = row + FIN_GROUPS < n_valid;
         uint32_t sums[2] = {0u, 0u};
#pragma unroll 1
         for (uint32_t first_rank = 1; first_rank < push.world; first_rank += IN_FLIGHT) {
            const uint2* const base = slot_rows + first_rank * rank_stride + static_cast<size_t>(row) * genome_length;
            uint2 words[2][IN_FLIGHT];
#pragma unroll
            for (uint32_t u = 0; u < IN_FLIGHT; ++u) {
               const bool rank_exists = first_rank + u < push.world;
               words[0][u] = rank_exists ? loadTagged(base + u * rank_stride) : make_uint2(0u, tag);
               words[1][u] = rank_exists && second ? loadTagged(base + u * rank_stride + static_cast<size_t>(FIN_GROUPS) * genome_length) : make_uint2(0u, tag);
            }
            uint32_t landed = 0;  // stays zero if every word carries the tag
            uint32_t partial[2] = {0u, 0u};
#pragma unroll
            for (uint32_t u = 0; u < IN_FLIGHT; ++u) {
               landed |= (words[0][u].y ^ tag) | (words[1][u].y ^ tag);
               partial[0] += words[0][u].x;
               partial[1] += words[1][u].x;
            }
            if (landed != 0) {  // some rank is behind: the words again, one by one, each polled until it is there
               partial[0] = 0;
               partial[1] = 0;
#pragma unroll 1
               for (uint32_t u = 0; u < IN_FLIGHT && first_rank + u < push.world; ++u) {
                  const uint2 word = pollTagged(base + u * rank_stride, tag);
                  peer_timed_out |= word.y != tag;
                  partial[0] += word.x;
                  if (second) {
                     const uint2 other = pollTagged(base + u * rank_stride + static_cast<size_t>(FIN_GROUPS) * genome_length, tag);
                     peer_timed_out |= other.y != tag;
                     partial[1] += other.x;
                  }
               }
            }
            sums[0] += partial[0];
            sums[1] += partial[1];
         }
         row_sum[row][i] = sums[0];  // (this rank's own count is added behind the barriers below; one thread per row and position)
         if (second) {
            row_sum[row + FIN_GROUPS][i] = sums[1];
         }
      }
   }
   if (threadIdx.x < 32) {
      uint32_t partial = 0;
#pragma unroll 1
      for (uint32_t block = lane; block < blockIdx.x; block += 32) {
         partial += block_totals[block];
      }
      partial = __reduce_add_sync(0xFFFFFFFFu, partial);
      if (lane == 0) {
         block_offset = partial;
      }
   }
   if (group == 0) {
      diff[p] = 0;
   }
   uint32_t others = 0;
#pragma unroll
   for (uint32_t k = 0; k < FIN_SYMBOLS_PER_THREAD; ++k) {
      others += group + k * FIN_GROUPS != reference_symbol ? values[k] : 0u;
   }
   part_others[group][i] = others;
   // inclusive scan of this block's 256 elements
   uint32_t inclusive = mine;
   if (group == 0) {
      for (int offset = 1; offset < 32; offset <<= 1) {
         const uint32_t other = __shfl_up_sync(0xFFFFFFFFu, inclusive, offset);
         if (lane >= static_cast<uint32_t>(offset)) {
            inclusive += other;
         }
      }
      if (lane == 31) {
         warp_totals[warp] = inclusive;
      }
   }
   __syncthreads();  // (also publishes the sharded query's wait)
   const bool shard_timed_out = SHARDED && (slot_is_free == 0 || peer_timed_out);
   if (group == 0) {
      uint32_t covered = block_offset + inclusive;
#pragma unroll 1
      for (uint32_t w = 0; w < warp; ++w) {
         covered += warp_totals[w];
      }
      uint32_t all_others = 0;
#pragma unroll
      for (uint32_t g = 0; g < FIN_GROUPS; ++g) {
         all_others += part_others[g][i];
      }
      const uint32_t reference_count = covered - all_others;
      reference_counts[i] = reference_count;
      if (in_range) {
         counts[reference_symbol * genome_length + p] = reference_count;
      }
   }
   if (MODE != FIN_COUNTS) {
      __syncthreads();
      // this thread's valid symbols: into the root's gather area (coalesced 8-byte stores over NVLink, every word tagged
      // with the query: fire and forget, nothing here waits for them to land), or into the rows in shared memory
      const uint32_t reference_count = reference_counts[i];
      if (in_range && !(MODE == FIN_PUSH && shard_timed_out)) {
         uint2* const root_rows = MODE == FIN_PUSH ? reinterpret_cast<uint2*>(push.root_block + SHARD_HEADER_BYTES) +
                                                        (static_cast<size_t>(shard_slot) * push.world + push.rank) * n_valid * genome_length + p
                                                   : nullptr;
#pragma unroll
         for (uint32_t k = 0; k < FIN_SYMBOLS_PER_THREAD; ++k) {
            const uint32_t symbol = group + k * FIN_GROUPS;
            if (((valid_mask >> symbol) & 1ULL) != 0) {
               const uint32_t row = __popcll(valid_mask & ((1ULL << symbol) - 1));
               const uint32_t value = symbol == reference_symbol ? reference_count : values[k];
               if (MODE == FIN_PUSH) {
                  storeTagged(root_rows + static_cast<size_t>(row) * genome_length, value, tag);
               } else if (MODE == FIN_COLLECT) {
                  row_sum[row][i] += value;  // (the other ranks' sum was stored before the barriers above; one thread per row and position)
               } else {
                  row_sum[row][i] = value;
               }
            }
         }
      }
   }
   if (USE_ROWS) {
      __syncthreads();
      // The output pass of addMutationsToOutput over the (summed) counts of one position, on request: which
      // (position, symbol) rows the action emits. `hits` is PAGE-LOCKED HOST memory: the tuples are stored straight
      // into it over PCIe (a few hundred 16-byte posted writes); the last block writes the header hits[0].
      if (group == 0 && in_range && !shard_timed_out) {
         const uint32_t genome_symbol = request.hits != nullptr ? column.global_reference[p] : 0u;
         uint32_t total = 0;
         uint32_t candidates = 0;  // OR of the counts that could be emitted (valid, not the reference genome's symbol)
         uint32_t row = 0;
#pragma unroll 1
         for (uint32_t symbol = 0; symbol < column.n_symbols; ++symbol) {
            if (((valid_mask >> symbol) & 1ULL) != 0) {
               const uint32_t sum = row_sum[row++][i];
               total += sum;
               candidates |= symbol != genome_symbol ? sum : 0u;
               if (MODE == FIN_COLLECT && push.summed_out != nullptr) {
                  push.summed_out[symbol * genome_length + p] = sum;
               }
            }
         }
         // (`count > threshold_count` cannot hold for a zero count)
         if (request.hits != nullptr && total != 0 && candidates != 0) {
            // ceil(double(total) * min_proportion) - 1, the reference's operations in the reference's order
            const uint32_t threshold_count =
               request.min_proportion == 0 ? 0u : static_cast<uint32_t>(ceil(__dmul_rn(static_cast<double>(total), request.min_proportion)) - 1.0);
            row = 0;
#pragma unroll 1
            for (uint32_t symbol = 0; symbol < column.n_symbols; ++symbol) {
               if (((valid_mask >> symbol) & 1ULL) == 0) {
                  continue;
               }
               const uint32_t sum = row_sum[row++][i];
               if (symbol != genome_symbol && sum > threshold_count) {
                  const uint32_t index = atomicAdd(&work_state[3], 1u);
                  if (index < request.capacity) {
                     request.hits[1 + index] = silo_mutation_hit{p, symbol, sum, total};
                  }
               }
            }
         }
      }
   }
   if (MODE == FIN_COLLECT && shard_timed_out) {  // (rare: a rank never delivered, or the queries are out of order)
      atomicOr(&reinterpret_cast<ShardBlockHeader*>(push.root_block)->collect_error, 2u);
   }
   // This block has consumed its block totals (block_offset went into `covered`) and appended its tuples:
   // the last block to get here clears the totals and the work-list state and writes the header (its first warp).
   __syncthreads();
   if (threadIdx.x < 32) {
      uint32_t is_last = 0;
      if (lane == 0) {
         if (SHARDED) {
            shardDebugStamp(push, 2);
            if (blockIdx.x == 0) {  // this rank's filter cardinality and error flag for the root
               ShardBlockHeader* const root_header = reinterpret_cast<ShardBlockHeader*>(push.root_block);
               unsigned long long cardinality = push.fixed_cardinality;
               uint32_t error = MODE == FIN_PUSH && shard_timed_out ? 2u : 0u;
               if (push.use_fixed_cardinality == 0) {
                  cardinality = *reinterpret_cast<volatile unsigned long long*>(&push.filter_scalars[0]);
                  error |= *reinterpret_cast<volatile uint32_t*>(&push.filter_scalars[2]);
               }
               storeTagged(&root_header->cardinality_low[shard_slot][push.rank], static_cast<uint32_t>(cardinality), tag);
               storeTagged(&root_header->cardinality_high[shard_slot][push.rank], static_cast<uint32_t>(cardinality >> 32), tag);
               storeTagged(&root_header->error[shard_slot][push.rank], error, tag);
            }
            shardDebugStamp(push, 3);
         }
         __threadfence();
         is_last = atomicAdd(&work_state[2], 1u) == gridDim.x - 1 ? 1u : 0u;
      }
      is_last = __shfl_sync(0xFFFFFFFFu, is_last, 0);
      if (is_last != 0) {
         __threadfence();
#pragma unroll 1
         for (uint32_t block = lane; block < gridDim.x; block += 32) {
            block_totals[block] = 0;
         }
         if (MODE == FIN_COLLECT) {
            // the whole group's result: every rank's cardinality and error flag, the slot back to every rank, the query
            // counted as collected
            ShardBlockHeader* const root_header = reinterpret_cast<ShardBlockHeader*>(push.root_block);
            unsigned long long cardinality = 0;
            uint32_t error = lane == 31 ? *reinterpret_cast<volatile uint32_t*>(&root_header->collect_error) : 0u;
            warpSumRankScalars(root_header, shard_slot, push.world, tag, lane, cardinality, error);
            if (lane == 0) {
               root_header->collect_error = 0;
               if (request.hits != nullptr) {
                  request.hits[0] = silo_mutation_hit{
                     *reinterpret_cast<volatile uint32_t*>(&work_state[3]), error, static_cast<uint32_t>(cardinality), static_cast<uint32_t>(cardinality >> 32)};
               }
               root_header->queries_collected = tag;
            }
            if (lane < push.world) {
               *reinterpret_cast<volatile uint32_t*>(&push.peers[lane]->released[shard_slot]) = (tag - 1) / SHARD_SLOTS + 1;
            }
         } else if (MODE == FIN_OUTPUT && lane == 0) {
            silo_mutation_hit header{*reinterpret_cast<volatile uint32_t*>(&work_state[3]), 0u, 0u, 0u};
            if (request.filter_scalars != nullptr) {
               const unsigned long long cardinality = *reinterpret_cast<volatile unsigned long long*>(&request.filter_scalars[0]);
               header.symbol = *reinterpret_cast<volatile uint32_t*>(&request.filter_scalars[2]);
               header.count = static_cast<uint32_t>(cardinality);
               header.total = static_cast<uint32_t>(cardinality >> 32);
               if (request.keep_scalars == 0) {
                  request.filter_scalars[0] = 0;
                  request.filter_scalars[2] = 0;
               }
            }
            request.hits[0] = header;
         }
         if (lane == 0) {
            if (SHARDED) {  // every block has read the query counter and the filter's scalars
               if (push.use_fixed_cardinality == 0) {
                  push.filter_scalars[0] = 0;
                  push.filter_scalars[2] = 0;
               }
               push.own_header->queries_pushed = tag;
            }
            work_state[0] = 0;  // the work list and its claim counter are empty between queries
            work_state[1] = 0;
            work_state[2] = 0;
            work_state[3] = 0;
         }
      }
   }
}

// The root's half of a sharded query when it is a call of its own (silo_gpu_sharded_collect*): reads the tagged rows of
// all `world` ranks from the slot (the root's own finalize kernel stored its rows there like everybody else's), polling
// the words that have not landed yet, sums them per position and valid symbol, runs the output pass of
// addMutationsToOutput (mutations_node.cpp:307-363) over the sums (as mutationHitsKernel does), and hands the slot back
// to every rank.
struct ShardCollect {
   ShardBlockHeader* header = nullptr;   // the root's own block: header, then the gather area [slot][world][n_valid][genome_length] of tagged words
   ShardBlockHeader* const* peers = nullptr;  // [world] every rank's block as mapped on the root
   uint32_t world = 0;
   uint32_t n_valid = 0;
   uint32_t* summed_out = nullptr;  // optional: [n_symbols][genome_length], the rows of the valid symbols are written
};

constexpr int COLLECT_THREADS = 128;
__global__ void __launch_bounds__(COLLECT_THREADS, 4) shardCollectKernel(DevColumn column, ShardCollect collect, uint32_t* __restrict__ work_state, HitRequest request) {
   const uint32_t genome_length = column.genome_length;
   const uint32_t p = blockIdx.x * COLLECT_THREADS + threadIdx.x;
   // (read before this block counts itself in below; the last block increments it after that)
   const uint32_t query_index = *reinterpret_cast<volatile uint32_t*>(&collect.header->queries_collected);
   const uint32_t slot = query_index % SHARD_SLOTS;
   const uint32_t generation = query_index / SHARD_SLOTS + 1;
   const uint32_t tag = query_index + 1;
   const uint2* const rows = reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(collect.header) + SHARD_HEADER_BYTES) +
                             static_cast<size_t>(slot) * collect.world * collect.n_valid * genome_length;
   bool timed_out = false;
   if (p < genome_length) {
      const uint32_t genome_symbol = request.hits != nullptr ? column.global_reference[p] : 0u;
      const size_t rank_stride = static_cast<size_t>(collect.n_valid) * genome_length;
      // (no per-symbol array: the few positions that emit rows sum their symbols a second time)
      auto sumOf = [&](uint32_t row) {  // (the loads of all ranks in flight together: one trip to L2 per row)
         const uint2* const base = rows + row * genome_length + p;
         uint2 words[SHARD_MAX_WORLD];
#pragma unroll
         for (uint32_t rank = 0; rank < SHARD_MAX_WORLD; ++rank) {
            words[rank] = rank < collect.world ? loadTagged(base + rank * rank_stride) : make_uint2(0u, tag);
         }
         uint32_t sum = 0;
#pragma unroll
         for (uint32_t rank = 0; rank < SHARD_MAX_WORLD; ++rank) {
            uint2 word = words[rank];
            if (word.y != tag) {
               word = pollTagged(base + rank * rank_stride, tag);
               timed_out |= word.y != tag;
            }
            sum += word.x;
         }
         return sum;
      };
      uint32_t total = 0;
      uint32_t candidates = 0;
      uint32_t row = 0;
      for (uint32_t symbol = 0; symbol < column.n_symbols; ++symbol) {
         if (((request.valid_mask >> symbol) & 1ULL) != 0) {
            const uint32_t sum = sumOf(row);
            total += sum;
            candidates |= symbol != genome_symbol ? sum : 0u;
            if (collect.summed_out != nullptr) {
               collect.summed_out[symbol * genome_length + p] = sum;
            }
            ++row;
         }
      }
      if (request.hits != nullptr && total != 0 && candidates != 0) {
         const uint32_t threshold_count =
            request.min_proportion == 0 ? 0u : static_cast<uint32_t>(ceil(__dmul_rn(static_cast<double>(total), request.min_proportion)) - 1.0);
         row = 0;
         for (uint32_t symbol = 0; symbol < column.n_symbols; ++symbol) {
            if (((request.valid_mask >> symbol) & 1ULL) == 0) {
               continue;
            }
            const uint32_t this_row = row++;
            if (symbol == genome_symbol) {
               continue;
            }
            const uint32_t sum = sumOf(this_row);
            if (sum > threshold_count) {
               const uint32_t index = atomicAdd(&work_state[3], 1u);
               if (index < request.capacity) {
                  request.hits[1 + index] = silo_mutation_hit{p, symbol, sum, total};
               }
            }
         }
      }
   }
   if (timed_out) {  // (rare: a rank never delivered)
      atomicOr(&collect.header->collect_error, 2u);
   }
   __syncthreads();
   if (threadIdx.x < 32) {
      const uint32_t lane = threadIdx.x;
      uint32_t is_last = 0;
      if (lane == 0) {
         __threadfence();
         is_last = atomicAdd(&work_state[2], 1u) == gridDim.x - 1 ? 1u : 0u;
      }
      is_last = __shfl_sync(0xFFFFFFFFu, is_last, 0);
      if (is_last != 0) {
         __threadfence();
         unsigned long long cardinality = 0;
         uint32_t error = lane == 31 ? *reinterpret_cast<volatile uint32_t*>(&collect.header->collect_error) : 0u;
         warpSumRankScalars(collect.header, slot, collect.world, tag, lane, cardinality, error);
         if (lane == 0) {
            collect.header->collect_error = 0;
            if (request.hits != nullptr) {
               request.hits[0] = silo_mutation_hit{
                  *reinterpret_cast<volatile uint32_t*>(&work_state[3]), error, static_cast<uint32_t>(cardinality), static_cast<uint32_t>(cardinality >> 32)};
            }
            work_state[2] = 0;
            work_state[3] = 0;
            collect.header->queries_collected = query_index + 1;
         }
         // every block has read the slot (its loads have returned): hand it back to the ranks. Nothing written here has
         // to be visible to them first, so plain system-scope stores, no fence.
         if (lane < collect.world) {
            *reinterpret_cast<volatile uint32_t*>(&collect.peers[lane]->released[slot]) = generation;
         }
      }
   }
}

// The output pass alone, over counts that are already final -- the multi-GPU scheduler all-reduces the
// per-rank counts first (counts are plain addends over row partitions) and then asks one rank for the
// rows: addMutationsToOutput (mutations_node.cpp:307-363) for one position per thread, header written by
// the last block as in finalizeCountsKernel.
__global__ void __launch_bounds__(FIN_THREADS) mutationHitsKernel(
   DevColumn column,
   const uint32_t* __restrict__ counts,
   uint32_t* __restrict__ work_state,
   HitRequest request
) {
   const uint32_t genome_length = column.genome_length;
   const uint32_t p = blockIdx.x * FIN_THREADS + threadIdx.x;
   if (p < genome_length) {
      const uint32_t genome_symbol = column.global_reference[p];
      uint32_t total = 0;
      uint32_t candidates = 0;
      for (uint32_t symbol = 0; symbol < column.n_symbols; ++symbol) {
         if (((request.valid_mask >> symbol) & 1ULL) != 0) {
            const uint32_t value = counts[symbol * genome_length + p];
            total += value;
            candidates |= symbol != genome_symbol ? value : 0u;
         }
      }
      if (total != 0 && candidates != 0) {
         const uint32_t threshold_count =
            request.min_proportion == 0 ? 0u : static_cast<uint32_t>(ceil(__dmul_rn(static_cast<double>(total), request.min_proportion)) - 1.0);
         for (uint32_t symbol = 0; symbol < column.n_symbols; ++symbol) {
            if (((request.valid_mask >> symbol) & 1ULL) == 0 || symbol == genome_symbol) {
               continue;
            }
            const uint32_t count = counts[symbol * genome_length + p];
            if (count > threshold_count) {
               const uint32_t index = atomicAdd(&work_state[3], 1u);
               if (index < request.capacity) {
                  request.hits[1 + index] = silo_mutation_hit{p, symbol, count, total};
               }
            }
         }
      }
   }
   __syncthreads();
   if (threadIdx.x == 0) {
      __threadfence();
      const uint32_t finished = atomicAdd(&work_state[2], 1u);
      if (finished == gridDim.x - 1) {
         __threadfence();
         silo_mutation_hit header{*reinterpret_cast<volatile uint32_t*>(&work_state[3]), 0u, 0u, 0u};
         if (request.filter_scalars != nullptr) {
            const unsigned long long cardinality = *reinterpret_cast<volatile unsigned long long*>(&request.filter_scalars[0]);
            header.symbol = *reinterpret_cast<volatile uint32_t*>(&request.filter_scalars[2]);
            header.count = static_cast<uint32_t>(cardinality);
            header.total = static_cast<uint32_t>(cardinality >> 32);
            request.filter_scalars[0] = 0;
            request.filter_scalars[2] = 0;
         }
         request.hits[0] = header;
         work_state[2] = 0;
         work_state[3] = 0;
      }
   }
}

// ---------------------------------------------------------------------------------------------

// The compiled geometries of the container kernel: {consumer warps, pieces per warp and stage visit, ring stages}.
// Variant 5 is the product (31 x 2 x 3 at 56 registers: as fast as the 64-register build of variant 0, and a 256-thread
// block of the coverage kernel or of the shard group's collect kernel still fits beside it on the SM); the others
// are kept for measurements (SILO_K1_VARIANT, read once per process).
constexpr K1Geometry K1_VARIANTS[] = {{31, 2, 3, 0}, {23, 2, 4, 1}, {20, 3, 3, 2}, {15, 4, 3, 3}, {31, 1, 6, 4}, {31, 2, 3, 5}};

template <int W, int P, int STAGES, int MAXREG, int MODE>
void launchContainerVariant(
   int blocks, cudaStream_t stream, const DevColumn& column, const uint64_t* words, uint32_t* work_state, const DevSegment* work_items,
   uint32_t* counts, uint32_t tail_factor, uint32_t* debug_times
) {
   static bool attribute_set = false;
   constexpr uint32_t dynamic_bytes = k1DynamicBytes(W, P, STAGES);
   static_assert(dynamic_bytes + 2 * TILE_BUFFER_BYTES <= 227 * 1024, "the ring does not fit the SM's shared memory");
   if (!attribute_set) {
      SILO_CUDA_CHECK(cudaFuncSetAttribute(containerAndCountKernel<W, P, STAGES, MAXREG, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dynamic_bytes)));
      attribute_set = true;
   }
   SILO_CUDA_CHECK(launchDependent(
      containerAndCountKernel<W, P, STAGES, MAXREG, MODE>, dim3(blocks), dim3((W + 1) * 32), dynamic_bytes, stream, column, words, work_state, work_items, counts,
      tail_factor, debug_times
   ));
}

template <typename... Args>
void launchContainerKernel(const K1Geometry& geometry, int mode, Args&&... args) {
#define SILO_K1_CASE(V, W, P, S, R)                                        \
   case V:                                                                 \
      if (mode == 1) {                                                     \
         launchContainerVariant<W, P, S, R, 1>(std::forward<Args>(args)...); \
      } else {                                                             \
         launchContainerVariant<W, P, S, R, 0>(std::forward<Args>(args)...); \
      }                                                                    \
      break;
   switch (geometry.variant) {
      SILO_K1_CASE(1, 23, 2, 4, 64)
      SILO_K1_CASE(2, 20, 3, 3, 80)
      SILO_K1_CASE(3, 15, 4, 3, 96)
      SILO_K1_CASE(4, 31, 1, 6, 56)
      SILO_K1_CASE(5, 31, 2, 3, 56)
      default:
         SILO_K1_CASE(0, 31, 2, 3, 64)
   }
#undef SILO_K1_CASE
}

void enqueueMutationCounts(
   silo_gpu_table* table,
   int column_index,
   const silo_gpu_filter* filter,
   uint32_t* d_counts,
   cudaStream_t stream,
   const HitRequest* request = nullptr,
   bool timed = false,    // record the per-call CUDA events that silo_gpu_get_stats reads (measurement only)
   bool prepared = false, // the filter interpreter already zeroed d_counts and built the work list (fused query)
   const ShardPush* push = nullptr  // sharded query: the finalize kernel also sends this rank's rows to the root
) {
   require(table != nullptr, "mutation_counts: table is NULL");
   require(column_index >= 0 && static_cast<size_t>(column_index) < table->columns.size(), "mutation_counts: bad column index");
   require(filter == nullptr || filter->table == table, "mutation_counts: filter belongs to another table");
   if (filter != nullptr && filter->out_of_layout) {
      throw ApiError(SILO_E_OUT_OF_LAYOUT, "the filter holds row ids outside the row layout: the action has no row data for them");
   }
   require(d_counts != nullptr, "mutation_counts: counts is NULL");
   const HostColumn& host = *table->columns[static_cast<size_t>(column_index)];
   const DevColumn& column = host.dev;
   const uint32_t n_chunks = table->n_chunks;
   const size_t counts_bytes = static_cast<size_t>(column.n_symbols) * column.genome_length * sizeof(uint32_t);

   table->last_stream = stream;
   table->last_column = column_index;
   table->last_popcounts = filter != nullptr ? filter->d_chunk_popcount : table->d_chunk_popcount_full;
   table->last_was_full = filter == nullptr;
   // The synchronous query calls do not pay for the four event records; the _async entry (what bench.py and
   // the probes time) does. (Inside a stream capture the records are capture bookkeeping only; silo_gpu_get_stats
   // then describes the last EAGER calls.)
   const uint64_t slot = table->timed_calls % silo_gpu_table::EVENT_RING;
   if (timed) {
      table->timed_calls++;
   }
   cudaEvent_t ev_begin = table->ev_begin[slot];
   cudaEvent_t ev_k1_begin = table->ev_k1_begin[slot];
   cudaEvent_t ev_k1_end = table->ev_k1_end[slot];
   cudaEvent_t ev_end = table->ev_end[slot];
   auto recordTiming = [&](cudaEvent_t event) {
      if (timed) {
         SILO_CUDA_CHECK(cudaEventRecord(event, stream));
      }
   };
   recordTiming(ev_begin);
   if (n_chunks == 0) {
      SILO_CUDA_CHECK(cudaMemsetAsync(d_counts, 0, counts_bytes, stream));
      if (request != nullptr) {
         std::memset(request->hits, 0, sizeof(silo_mutation_hit));  // (host memory; nothing on the stream writes it)
      }
      recordTiming(ev_k1_begin);
      recordTiming(ev_k1_end);
      recordTiming(ev_end);
      return;
   }
   const uint64_t* words = filter != nullptr ? filter->d_words : table->d_full_words;
   const uint32_t* popcounts = filter != nullptr ? filter->d_chunk_popcount : table->d_chunk_popcount_full;
   uint32_t* const diff = table->d_coverage_diff;  // all-zero: the finalize kernel of the query before cleared it

   // fork: the coverage kernel needs nothing but the filter, so it runs on the auxiliary stream beside
   // the prepare kernel and the first microseconds of the container kernel
   SILO_CUDA_CHECK(cudaEventRecord(table->ev_fork, stream));
   SILO_CUDA_CHECK(cudaStreamWaitEvent(table->aux_stream, table->ev_fork, 0));
   {
      const NvtxRange range("Mutations: subtractFilteredNCounts + subtractStartAndEndNCounts [coverageDiffKernel]");
      coverageDiffKernel<<<n_chunks * K6_SLICES, K6_THREADS, 0, table->aux_stream>>>(column, words, popcounts, diff);
   }
   SILO_CUDA_CHECK(cudaGetLastError());
   SILO_CUDA_CHECK(cudaEventRecord(table->ev_join, table->aux_stream));

   if (!prepared) {
      const int prepare_blocks = static_cast<int>(std::max<uint32_t>(n_chunks, static_cast<uint32_t>(table->ctx->sm_count)));
      prepareQueryKernel<<<prepare_blocks, PREP_THREADS, 0, stream>>>(
         column, filter != nullptr ? popcounts : nullptr, table->d_work_state, table->d_work_items, d_counts,
         static_cast<uint32_t>(counts_bytes / sizeof(uint32_t))
      );
      SILO_CUDA_CHECK(cudaGetLastError());
      table->stats.kernel_launches += 1;
   }

   recordTiming(ev_k1_begin);
   if (filter == nullptr) {
      if (column.n_containers > 0) {
         const int blocks = static_cast<int>(std::min<uint64_t>((column.n_containers + 255) / 256, static_cast<uint64_t>(table->ctx->sm_count) * 8));
         const NvtxRange range("Mutations: countActualMutations [containerCardinalityKernel]");
         containerCardinalityKernel<<<blocks, 256, 0, stream>>>(column, d_counts);
         SILO_CUDA_CHECK(cudaGetLastError());
         table->stats.kernel_launches++;
      }
   } else if (column.n_segments > 0) {
      const K1Geometry& geometry = k1Geometry();
      require(host.segment_pieces == geometry.segmentPieces(), "mutation_counts: the column was uploaded for another container-kernel geometry");
      static const int stream_only = [] {
         const char* flag = std::getenv("SILO_K1_STREAM_ONLY");
         return flag != nullptr && flag[0] == '1' ? 1 : 0;
      }();
      static const uint32_t tail_factor = [] {
         const char* flag = std::getenv("SILO_K1_TAIL");
         return flag != nullptr ? static_cast<uint32_t>(std::max(0, std::atoi(flag))) : 3u;
      }();
      const int blocks = static_cast<int>(std::min<uint32_t>(column.n_segments, static_cast<uint32_t>(table->ctx->sm_count)));
      // SILO_K1_DEBUG=1 (measurement aid, synchronises): when every CTA's producer and consumers ran out of work
      static const bool debug = std::getenv("SILO_K1_DEBUG") != nullptr;
      static uint32_t* d_debug_times = nullptr;
      if (debug && d_debug_times == nullptr) {
         SILO_CUDA_CHECK(cudaMalloc(&d_debug_times, 4 * sizeof(uint32_t) * 1024));
      }
      {
         const NvtxRange range("Mutations: countActualFilteredMutations [containerAndCountKernel]");
         launchContainerKernel(geometry, stream_only, blocks, stream, column, words, table->d_work_state, table->d_work_items, d_counts, tail_factor, debug ? d_debug_times : nullptr);
      }
      SILO_CUDA_CHECK(cudaGetLastError());
      table->stats.kernel_launches++;
      cudaStreamCaptureStatus capture_status = cudaStreamCaptureStatusNone;
      if (debug && cudaStreamIsCapturing(stream, &capture_status) == cudaSuccess && capture_status == cudaStreamCaptureStatusNone) {
         std::vector<uint32_t> times(4 * static_cast<size_t>(blocks));
         SILO_CUDA_CHECK(cudaMemcpyAsync(times.data(), d_debug_times, times.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
         SILO_CUDA_CHECK(cudaStreamSynchronize(stream));
         uint32_t first_start = times[0];
         for (int b = 0; b < blocks; ++b) {
            first_start = static_cast<int32_t>(times[4 * b] - first_start) < 0 ? times[4 * b] : first_start;
         }
         std::vector<uint32_t> starts, producer_ends, consumer_ends, stages;
         for (int b = 0; b < blocks; ++b) {
            starts.push_back(times[4 * b] - first_start);
            producer_ends.push_back(times[4 * b + 1] - first_start);
            consumer_ends.push_back(times[4 * b + 2] - first_start);
            stages.push_back(times[4 * b + 3]);
         }
         auto summary = [](std::vector<uint32_t>& values) {
            std::sort(values.begin(), values.end());
            char text[96];
            std::snprintf(text, sizeof(text), "min %u p10 %u median %u p90 %u max %u", values.front(), values[values.size() / 10], values[values.size() / 2],
                          values[values.size() * 9 / 10], values.back());
            return std::string(text);
         };
         std::fprintf(stderr, "[silo k1 debug] %d CTAs (ns from the first start): start %s | producer end %s | consumers end %s | stages %s\n", blocks,
                      summary(starts).c_str(), summary(producer_ends).c_str(), summary(consumer_ends).c_str(), summary(stages).c_str());
      }
   }
   recordTiming(ev_k1_end);
   SILO_CUDA_CHECK(cudaStreamWaitEvent(stream, table->ev_join, 0));
   const NvtxRange finalize_range("Mutations: subtractCumulativeNsFromPositions + accumulateFinalCounts [finalizeCountsKernel]");
   {
      const uint32_t blocks = diffPadded(column.genome_length) / FIN_THREADS;
      const HitRequest hit_request = request != nullptr ? *request : HitRequest{};
      const ShardPush shard_push = push != nullptr ? *push : ShardPush{};
      require(column.n_symbols <= 32, "mutation_counts: alphabets of more than 32 symbols are not supported");
      if (shard_push.peers != nullptr) {
         SILO_CUDA_CHECK(launchDependent(finalizeCountsKernel<FIN_COLLECT>, dim3(blocks), dim3(FIN_BLOCK_THREADS), 0, stream, column, diff, d_counts, table->d_work_state, hit_request, shard_push));
      } else if (shard_push.root_block != nullptr) {
         SILO_CUDA_CHECK(launchDependent(finalizeCountsKernel<FIN_PUSH>, dim3(blocks), dim3(FIN_BLOCK_THREADS), 0, stream, column, diff, d_counts, table->d_work_state, hit_request, shard_push));
      } else if (hit_request.hits != nullptr) {
         SILO_CUDA_CHECK(launchDependent(finalizeCountsKernel<FIN_OUTPUT>, dim3(blocks), dim3(FIN_BLOCK_THREADS), 0, stream, column, diff, d_counts, table->d_work_state, hit_request, shard_push));
      } else {
         SILO_CUDA_CHECK(launchDependent(finalizeCountsKernel<FIN_COUNTS>, dim3(blocks), dim3(FIN_BLOCK_THREADS), 0, stream, column, diff, d_counts, table->d_work_state, hit_request, shard_push));
      }
   }
   SILO_CUDA_CHECK(cudaGetLastError());
   table->stats.kernel_launches += 2;
   recordTiming(ev_end);
}

}  // namespace

const K1Geometry& k1Geometry() {
   static const K1Geometry geometry = [] {
      const char* flag = std::getenv("SILO_K1_VARIANT");
      const int variant = flag != nullptr ? std::atoi(flag) : 5;
      constexpr int n_variants = static_cast<int>(sizeof(K1_VARIANTS) / sizeof(K1_VARIANTS[0]));
      return K1_VARIANTS[variant >= 0 && variant < n_variants ? variant : 5];
   }();
   return geometry;
}

void enqueuePreparedCountsLocked(silo_gpu_table* table, int column, const silo_gpu_filter* filter, uint32_t* d_counts, cudaStream_t stream) {
   enqueueMutationCounts(table, column, filter, d_counts, stream, nullptr, true, true);
}

int shardGroupColumnLocked(const silo_gpu_table* table);

}  // namespace silo

using namespace silo;

extern "C" {

int silo_gpu_mutation_counts_async(
   silo_gpu_table* table,
   int column,
   const silo_gpu_filter* filter,
   void* d_counts,
   void* cuda_stream
) {
   return guarded([&] {
      require(table != nullptr, "mutation_counts: table is NULL");
      std::lock_guard<std::mutex> lock(table->mutex);
      SILO_CUDA_CHECK(cudaSetDevice(table->ctx->device));
      cudaStream_t stream = cuda_stream != nullptr ? static_cast<cudaStream_t>(cuda_stream) : table->ctx->stream;
      enqueueMutationCounts(table, column, filter, static_cast<uint32_t*>(d_counts), stream, nullptr, true);
   });
}

// the synchronous calls: [evaluate a program,] enqueue, copy the rows of the wanted symbols to the
// host, synchronise ONCE. With a program the filter lives only inside the call.
static void mutationCountsToHost(
   silo_gpu_table* table,
   int column,
   const silo_gpu_filter* filter,
   const silo_filter_program* program,
   uint64_t symbol_mask,
   uint32_t* counts,
   uint64_t* cardinality_out
) {
   require(table != nullptr && counts != nullptr, "mutation_counts: NULL argument");
   std::lock_guard<std::mutex> lock(table->mutex);
   SILO_CUDA_CHECK(cudaSetDevice(table->ctx->device));
   cudaStream_t stream = table->ctx->stream;
   require(column >= 0 && static_cast<size_t>(column) < table->columns.size(), "mutation_counts: bad column index");
   uint8_t* d_staging = nullptr;
   silo_gpu_filter* own_filter = nullptr;
   unsigned long long host_cardinality = 0;
   uint32_t host_error = 0;
   bool scalars_pending = false;
   try {
      if (program != nullptr) {
         // a program that is just PUSH_FULL is the `cardinality == numRows` path (stored cardinalities)
         const bool trivially_full = program->n_instrs == 1 && program->instrs != nullptr && program->instrs[0].opcode == SILO_OP_PUSH_FULL;
         if (trivially_full) {
            host_cardinality = table->n_rows;
         } else {
            own_filter = evalProgramAsync(table, program, stream, &d_staging);
            filter = own_filter;
            // cardinality and error flag sit 16 bytes apart (allocFilter): one copy into page-locked memory, read
            // after the call's single synchronise (a copy into pageable memory would block the host right here)
            SILO_CUDA_CHECK(cudaMemcpyAsync(table->h_scalars_pinned, own_filter->d_cardinality, 32, cudaMemcpyDeviceToHost, stream));
            scalars_pending = true;
         }
      }
      enqueueMutationCounts(table, column, filter, table->d_counts, stream);
      const HostColumn& host = *table->columns[static_cast<size_t>(column)];
      const size_t row_values = host.dev.genome_length;
      // page-locked destination (silo_gpu_host_alloc): the copy engine writes it directly
      cudaPointerAttributes attributes{};
      const bool pinned_destination =
         cudaPointerGetAttributes(&attributes, counts) == cudaSuccess && attributes.type == cudaMemoryTypeHost;
      cudaGetLastError();
      uint32_t* destination = pinned_destination ? counts : table->h_counts_pinned;
      // one copy per maximal range of consecutive wanted symbols
      std::vector<std::pair<uint32_t, uint32_t>> ranges;
      for (uint32_t symbol = 0; symbol < host.dev.n_symbols;) {
         if (((symbol_mask >> symbol) & 1ULL) == 0) {
            ++symbol;
            continue;
         }
         uint32_t end = symbol;
         while (end < host.dev.n_symbols && ((symbol_mask >> end) & 1ULL) != 0) {
            ++end;
         }
         ranges.emplace_back(symbol, end);
         symbol = end;
      }
      for (const auto& [first, end] : ranges) {
         SILO_CUDA_CHECK(cudaMemcpyAsync(
            destination + first * row_values, table->d_counts + first * row_values, (end - first) * row_values * sizeof(uint32_t),
            cudaMemcpyDeviceToHost, stream
         ));
      }
      if (d_staging != nullptr) {
         SILO_CUDA_CHECK(cudaFreeAsync(d_staging, stream));
         d_staging = nullptr;
      }
      SILO_CUDA_CHECK(cudaStreamSynchronize(stream));
      if (scalars_pending) {
         host_cardinality = table->h_scalars_pinned[0];
         host_error = static_cast<uint32_t>(table->h_scalars_pinned[2]);
      }
      if (!pinned_destination) {
         for (const auto& [first, end] : ranges) {
            std::memcpy(counts + first * row_values, table->h_counts_pinned + first * row_values, (end - first) * row_values * sizeof(uint32_t));
         }
      }
   } catch (...) {
      if (d_staging != nullptr) {
         cudaFreeAsync(d_staging, stream);
      }
      cudaStreamSynchronize(stream);
      releaseFilterLocked(own_filter);
      throw;
   }
   releaseFilterLocked(own_filter);
   if (host_error != 0) {
      throw ApiError(SILO_E_OUT_OF_LAYOUT, "a leaf bitmap holds row ids outside the row layout");
   }
   if (cardinality_out != nullptr) {
      *cardinality_out = host_cardinality;
   }
}

int silo_gpu_column_set_reference(silo_gpu_table* table, int column, const uint8_t* reference_symbols) {
   return guarded([&] {
      require(table != nullptr && reference_symbols != nullptr, "silo_gpu_column_set_reference: NULL argument");
      std::lock_guard<std::mutex> lock(table->mutex);
      require(column >= 0 && static_cast<size_t>(column) < table->columns.size(), "silo_gpu_column_set_reference: bad column index");
      SILO_CUDA_CHECK(cudaSetDevice(table->ctx->device));
      HostColumn& host = *table->columns[static_cast<size_t>(column)];
      for (uint32_t p = 0; p < host.dev.genome_length; ++p) {
         require(reference_symbols[p] < host.dev.n_symbols, "silo_gpu_column_set_reference: symbol id out of range");
      }
      uint8_t* d_reference = const_cast<uint8_t*>(host.dev.global_reference);
      if (d_reference == nullptr) {
         d_reference = deviceAlloc<uint8_t>(host.dev.genome_length, &table->device_bytes);
         host.allocations.push_back(d_reference);
         host.dev.global_reference = d_reference;
      }
      SILO_CUDA_CHECK(cudaMemcpyAsync(d_reference, reference_symbols, host.dev.genome_length, cudaMemcpyHostToDevice, table->ctx->stream));
      SILO_CUDA_CHECK(cudaStreamSynchronize(table->ctx->stream));
   });
}

// SILO_QUERY_TRACE=1: where a fused query call spends its host time (printed every 64 calls)
struct QueryTrace {
   static bool enabled() {
      static const bool on = std::getenv("SILO_QUERY_TRACE") != nullptr;
      return on;
   }
   std::chrono::steady_clock::time_point begin = std::chrono::steady_clock::now();
   void mark(int phase) {
      if (!enabled()) {
         return;
      }
      static double sums[5] = {0, 0, 0, 0, 0};  // stage, enqueue, wait, results (3), entry: lock + device + buffers (4)
      static uint64_t calls = 0;
      const auto now = std::chrono::steady_clock::now();
      sums[phase] += std::chrono::duration<double, std::micro>(now - begin).count();
      begin = now;
      if (phase == 3 && ++calls % 64 == 0) {
         std::fprintf(stderr, "[silo query trace] entry %.1f us, stage %.1f us, enqueue %.1f us, wait %.1f us, results %.1f us (mean of 64)\n", sums[4] / 64,
                      sums[0] / 64, sums[1] / 64, sums[2] / 64, sums[3] / 64);
         sums[0] = sums[1] = sums[2] = sums[3] = sums[4] = 0;
      }
   }
};

// page-locked tuple buffer for the worst case: every valid symbol but the reference genome's at every position
static void ensureHitsTuples(silo_gpu_table* table, uint64_t needed, cudaStream_t stream);
static void ensureHitsCapacity(silo_gpu_table* table, const HostColumn& host, uint64_t valid_symbol_mask, cudaStream_t stream) {
   ensureHitsTuples(table, static_cast<uint64_t>(__builtin_popcountll(valid_symbol_mask)) * host.dev.genome_length, stream);
}
static void ensureHitsTuples(silo_gpu_table* table, uint64_t needed, cudaStream_t stream) {
   if (needed > table->hits_capacity || table->h_hits_pinned == nullptr) {
      SILO_CUDA_CHECK(cudaStreamSynchronize(stream));
      dropQueryGraphsLocked(table);  // their kernel nodes write the old buffer
      if (table->h_hits_pinned != nullptr) {
         cudaFreeHost(table->h_hits_pinned);
         table->h_hits_pinned = nullptr;
      }
      SILO_CUDA_CHECK(cudaMallocHost(&table->h_hits_pinned, (needed + 1) * sizeof(silo_mutation_hit)));
      table->hits_capacity = needed;
   }
}

// The tuples a query returns: copied out of the table's page-locked buffer (which the next query on the table -- maybe from
// another host thread, as soon as the table's mutex is released -- overwrites) into storage of the CALLING THREAD, valid until
// that thread's next call that returns tuples. `slot`: a call that returns several tuple lists keeps them all.
// The kernel appends in whatever order its threads get there; the caller gets (position, symbol id) order. A few hundred
// tuples: a comparison sort spends ~10 us of a 120 us query on mispredicted branches (measured: "results 12.0 us" in
// SILO_QUERY_TRACE), so the order comes from an LSD radix sort of indices, one byte of (position << 6 | symbol) per pass.
static const silo_mutation_hit* publishHits(const silo_mutation_hit* first, uint64_t count, size_t slot = 0) {
   thread_local std::vector<std::vector<silo_mutation_hit>> published;
   thread_local std::vector<uint64_t> keys;
   thread_local std::vector<uint32_t> order;
   thread_local std::vector<uint32_t> next_order;
   if (published.size() <= slot) {
      published.resize(slot + 1);
   }
   std::vector<silo_mutation_hit>& out = published[slot];
   out.resize(count);
   keys.resize(count);
   order.resize(count);
   next_order.resize(count);
   uint64_t all_bits = 0;
   for (uint64_t i = 0; i < count; ++i) {
      keys[i] = (static_cast<uint64_t>(first[i].position) << 6) | (first[i].symbol & 63u);
      all_bits |= keys[i];
      order[i] = static_cast<uint32_t>(i);
   }
   for (uint32_t shift = 0; shift < 64 && (all_bits >> shift) != 0; shift += 8) {
      uint32_t starts[257] = {0};
      for (uint64_t i = 0; i < count; ++i) {
         ++starts[((keys[i] >> shift) & 255u) + 1];
      }
      for (uint32_t digit = 0; digit < 256; ++digit) {
         starts[digit + 1] += starts[digit];
      }
      for (uint64_t i = 0; i < count; ++i) {  // (stable: the order of the previous pass survives inside a digit)
         const uint32_t index = order[i];
         next_order[starts[(keys[index] >> shift) & 255u]++] = index;
      }
      order.swap(next_order);
   }
   for (uint64_t i = 0; i < count; ++i) {
      out[i] = first[order[i]];
   }
   return out.data();
}

int silo_gpu_query_mutation_counts_async(
   silo_gpu_table* table,
   const silo_filter_program* program,
   int column,
   void* d_counts,
   void* cuda_stream
) {
   return guarded([&] {
      require(table != nullptr && program != nullptr && d_counts != nullptr, "silo_gpu_query_mutation_counts_async: NULL argument");
      std::lock_guard<std::mutex> lock(table->mutex);
      SILO_CUDA_CHECK(cudaSetDevice(table->ctx->device));
      cudaStream_t stream = cuda_stream != nullptr ? static_cast<cudaStream_t>(cuda_stream) : table->ctx->stream;
      const bool trivially_full = program->n_instrs == 1 && program->instrs != nullptr && program->instrs[0].opcode == SILO_OP_PUSH_FULL;
      if (trivially_full) {
         enqueueMutationCounts(table, column, nullptr, static_cast<uint32_t*>(d_counts), stream);
         return;
      }
      StagedQuery staged;
      require(column >= 0 && static_cast<size_t>(column) < table->columns.size(), "silo_gpu_query_mutation_counts_async: bad column index");
      stageQueryLocked(table, program, &staged, column, static_cast<uint32_t*>(d_counts));
      enqueueStagedQuery(table, staged, stream, false);  // (nothing guarantees that a hits call reset the scalars)
      enqueueMutationCounts(table, column, table->query_filter, static_cast<uint32_t*>(d_counts), stream, nullptr, false, true);
   });
}

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
) {
   return guarded([&] {
      require(table != nullptr && d_counts != nullptr && hits != nullptr && n_hits != nullptr, "silo_gpu_mutation_hits_from_counts: NULL argument");
      std::lock_guard<std::mutex> lock(table->mutex);
      SILO_CUDA_CHECK(cudaSetDevice(table->ctx->device));
      cudaStream_t stream = cuda_stream != nullptr ? static_cast<cudaStream_t>(cuda_stream) : table->ctx->stream;
      require(column >= 0 && static_cast<size_t>(column) < table->columns.size(), "silo_gpu_mutation_hits_from_counts: bad column index");
      const HostColumn& host = *table->columns[static_cast<size_t>(column)];
      require(host.dev.global_reference != nullptr, "silo_gpu_mutation_hits_from_counts: call silo_gpu_column_set_reference first");
      if (host.dev.n_symbols < 64) {
         valid_symbol_mask &= (1ULL << host.dev.n_symbols) - 1;
      }
      ensureHitsCapacity(table, host, valid_symbol_mask, stream);
      HitRequest request;
      request.hits = table->h_hits_pinned;
      request.capacity = static_cast<uint32_t>(table->hits_capacity);
      request.valid_mask = valid_symbol_mask;
      request.min_proportion = min_proportion;
      // the filter of a preceding silo_gpu_query_mutation_counts_async: its cardinality and error flag ride along
      request.filter_scalars = table->query_filter != nullptr ? table->query_filter->d_cardinality : nullptr;
      mutationHitsKernel<<<(host.dev.genome_length + FIN_THREADS - 1) / FIN_THREADS, FIN_THREADS, 0, stream>>>(
         host.dev, static_cast<const uint32_t*>(d_counts), table->d_work_state, request
      );
      SILO_CUDA_CHECK(cudaGetLastError());
      table->stats.kernel_launches++;
      SILO_CUDA_CHECK(cudaStreamSynchronize(stream));
      const silo_mutation_hit header = table->h_hits_pinned[0];
      if (header.symbol != 0) {
         throw ApiError(SILO_E_OUT_OF_LAYOUT, "a leaf bitmap holds row ids outside the row layout");
      }
      const uint64_t count = std::min<uint64_t>(header.position, table->hits_capacity);
      *hits = publishHits(table->h_hits_pinned + 1, count);
      *n_hits = count;
      if (shard_cardinality != nullptr) {
         *shard_cardinality = header.count | (static_cast<unsigned long long>(header.total) << 32);
      }
   });
}

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
) {
   return guarded([&] {
      QueryTrace trace;
      require(table != nullptr && hits != nullptr && n_hits != nullptr, "silo_gpu_query_mutation_hits: NULL argument");
      std::lock_guard<std::mutex> lock(table->mutex);
      SILO_CUDA_CHECK(cudaSetDevice(table->ctx->device));
      cudaStream_t stream = table->ctx->stream;
      require(column >= 0 && static_cast<size_t>(column) < table->columns.size(), "silo_gpu_query_mutation_hits: bad column index");
      const HostColumn& host = *table->columns[static_cast<size_t>(column)];
      require(host.dev.global_reference != nullptr, "silo_gpu_query_mutation_hits: call silo_gpu_column_set_reference first");
      if (host.dev.n_symbols < 64) {
         valid_symbol_mask &= (1ULL << host.dev.n_symbols) - 1;
      }
      ensureHitsCapacity(table, host, valid_symbol_mask, stream);
      HitRequest request;
      request.hits = table->h_hits_pinned;  // page-locked host memory, written by the finalize kernel
      request.capacity = static_cast<uint32_t>(table->hits_capacity);
      request.valid_mask = valid_symbol_mask;
      request.min_proportion = min_proportion;

      unsigned long long host_cardinality = 0;
      uint32_t host_error = 0;
      uint64_t count = 0;
      const bool trivially_full =
         program != nullptr && program->n_instrs == 1 && program->instrs != nullptr && program->instrs[0].opcode == SILO_OP_PUSH_FULL;
      if (trivially_full) {
         host_cardinality = table->n_rows;
         filter = nullptr;
      }
      const bool own_program = program != nullptr && !trivially_full;
      StagedQuery staged;
      trace.mark(4);
      if (own_program) {
         // host work only: the pinned staging buffer now holds this query (the interpreter also prepares the counts kernels)
         stageQueryLocked(table, program, &staged, column, table->d_counts);
         filter = table->query_filter;
         request.filter_scalars = filter->d_cardinality;  // reported in the header, zeroed again by the finalize kernel
      }
      // everything the query puts on the stream; with a program it touches persistent buffers only, so the
      // sequence is the same for every query of one shape and can be replayed as a graph
      auto enqueueAll = [&]() {
         if (own_program) {
            enqueueStagedQuery(table, staged, stream);
         }
         enqueueMutationCounts(table, column, filter, table->d_counts, stream, &request, false, own_program);
      };
      try {
         cudaGraphExec_t replay = nullptr;
         if (own_program) {
            // the shape of the query: every kernel parameter and copy size that enqueueAll bakes into nodes
            std::string key(reinterpret_cast<const char*>(staged.params), sizeof(staged.params));
            const uint64_t scalars[] = {staged.staged_bytes, staged.shared_bytes, static_cast<uint64_t>(column), valid_symbol_mask,
                                        reinterpret_cast<uint64_t>(table->h_hits_pinned), table->hits_capacity,
                                        reinterpret_cast<uint64_t>(host.dev.containers), host.dev.n_segments,
                                        reinterpret_cast<uint64_t>(host.dev.global_reference), reinterpret_cast<uint64_t>(table->d_counts),
                                        reinterpret_cast<uint64_t>(table->d_work_items), reinterpret_cast<uint64_t>(table->d_coverage_diff)};
            key.append(reinterpret_cast<const char*>(scalars), sizeof(scalars));
            key.append(reinterpret_cast<const char*>(&min_proportion), sizeof(min_proportion));
            replay = queryGraphFor(table, std::move(key), stream, enqueueAll);
         }
         trace.mark(0);  // staged (host work)
         if (replay != nullptr) {
            SILO_CUDA_CHECK(cudaGraphLaunch(replay, stream));
            table->stats.kernel_launches += 4;  // interpreter (+ prepare), coverage, container, finalize
         } else {
            enqueueAll();
         }
         trace.mark(1);  // on the stream
         SILO_CUDA_CHECK(cudaStreamSynchronize(stream));
         trace.mark(2);  // device done
         const silo_mutation_hit header = table->h_hits_pinned[0];
         count = std::min<uint64_t>(header.position, table->hits_capacity);
         if (own_program) {
            host_error = header.symbol;
            host_cardinality = header.count | (static_cast<unsigned long long>(header.total) << 32);
         }
      } catch (...) {
         cudaStreamSynchronize(stream);
         if (own_program) {  // the finalize kernel may not have run: leave the persistent filter's scalars zero
            cudaMemsetAsync(table->query_filter->d_cardinality, 0, 32, stream);
            cudaMemsetAsync(table->d_work_state, 0, 4 * sizeof(uint32_t), stream);
            cudaStreamSynchronize(stream);
         }
         throw;
      }
      if (host_error != 0) {
         throw ApiError(SILO_E_OUT_OF_LAYOUT, "a leaf bitmap holds row ids outside the row layout");
      }
      // the kernel appends in whatever order its threads get there: (position, symbol id) order
      silo_mutation_hit* const first = table->h_hits_pinned + 1;
      *hits = publishHits(first, count);
      trace.mark(3);
      *n_hits = count;
      if (cardinality != nullptr && program != nullptr) {
         *cardinality = host_cardinality;
      }
   });
}

// ---- row-partitioned tables: the scheduler's device side (see ShardBlockHeader above) ----------------------

}  // extern "C"

namespace silo {

struct ShardGroup {
   int rank = 0;
   int world = 1;
   int column = -1;
   uint64_t valid_mask = 0;
   uint32_t n_valid = 0;
   uint32_t genome_length = 0;
   uint8_t* d_local = nullptr;  // this rank's exported block: header (+ the gather area on the root)
   size_t local_bytes = 0;
   uint8_t* d_root = nullptr;   // the root's block as mapped here (== d_local on the root)
   std::vector<uint8_t*> peer_blocks;     // root: every rank's block as mapped here
   std::vector<bool> opened_with_ipc;     // which of d_root / peer_blocks must be closed with cudaIpcCloseMemHandle
   ShardBlockHeader** d_peer_table = nullptr;  // root: device copy of peer_blocks
   uint32_t* d_collect_state = nullptr;        // root: the collect kernel's block counter and tuple counter
   unsigned long long* d_debug_times = nullptr;  // SILO_SHARD_DEBUG=1: the last finalize kernel's stamps, printed when the group is freed
   uint32_t debug_blocks = 0;
   uint64_t queries_enqueued = 0;  // the same on every rank: queries are issued in the same order everywhere
   uint64_t queries_collected = 0; // root
   size_t rowsBytes() const { return static_cast<size_t>(n_valid) * genome_length * sizeof(uint2); }  // tagged words
};

struct ShardHandle {  // SILO_SHARD_HANDLE_BYTES
   uint64_t magic;
   uint64_t process_id;
   uint64_t pointer;   // valid inside the exporting process (several shards of one table in one process: tests)
   uint64_t bytes;
   cudaIpcMemHandle_t ipc;
};
static_assert(sizeof(ShardHandle) <= SILO_SHARD_HANDLE_BYTES);
constexpr uint64_t SHARD_HANDLE_MAGIC = 0x53494C4F53484152ULL;

ShardPush shardPushOf(const ShardGroup& group) {
   ShardPush push;
   push.root_block = group.d_root;
   push.own_header = reinterpret_cast<ShardBlockHeader*>(group.d_local);
   push.rank = static_cast<uint32_t>(group.rank);
   push.world = static_cast<uint32_t>(group.world);
   push.n_valid = group.n_valid;
   push.valid_mask = group.valid_mask;
   push.debug_times = group.d_debug_times;
   return push;
}

int shardGroupColumnLocked(const silo_gpu_table* table) {
   require(table->shard != nullptr && table->shard->d_root != nullptr, "the table is not connected to a shard group");
   return table->shard->column;
}

// a prepared program's filter (its interpreter launch zeroed table->d_counts and built the work list): counts of the
// group's column + this rank's rows to the root
void enqueuePreparedShardedLocked(silo_gpu_table* table, const silo_gpu_filter* filter, cudaStream_t stream, bool collect_here, void* d_summed_counts) {
   ShardGroup* group = table->shard;
   ShardPush push = shardPushOf(*group);
   push.filter_scalars = filter->d_cardinality;
   if (collect_here) {
      require(group->rank == 0 && group->d_peer_table != nullptr, "sharded collect: only the connected root (rank 0) collects");
      push.peers = group->d_peer_table;
      push.summed_out = static_cast<uint32_t*>(d_summed_counts);
      group->queries_collected++;
   }
   enqueueMutationCounts(table, group->column, filter, table->d_counts, stream, nullptr, true, true, &push);
   group->queries_enqueued++;
}

void freeShardGroup(silo_gpu_table* table) {
   ShardGroup* group = table->shard;
   if (group == nullptr) {
      return;
   }
   cudaStreamSynchronize(table->ctx->stream);
   if (group->d_debug_times != nullptr) {
      cudaDeviceSynchronize();
      std::vector<unsigned long long> times(4 * static_cast<size_t>(group->debug_blocks));
      cudaMemcpy(times.data(), group->d_debug_times, times.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
      unsigned long long first = ~0ULL;
      for (uint32_t b = 0; b < group->debug_blocks; ++b) {
         first = std::min(first, times[4 * b]);
      }
      unsigned long long maxima[4] = {0, 0, 0, 0};
      std::vector<unsigned long long> waits, rows, fences;
      for (uint32_t b = 0; b < group->debug_blocks; ++b) {
         for (int k = 0; k < 4; ++k) {
            maxima[k] = std::max(maxima[k], times[4 * b + k] - first);
         }
         waits.push_back(times[4 * b + 1] - times[4 * b]);
         rows.push_back(times[4 * b + 2] - times[4 * b + 1]);
         fences.push_back(times[4 * b + 3] - times[4 * b + 2]);
      }
      auto median = [](std::vector<unsigned long long>& v) { std::sort(v.begin(), v.end()); return v[v.size() / 2]; };
      std::fprintf(stderr, "[silo shard debug] rank %d, last finalize kernel, ns from its first block's stamp: last start %llu, last wait over %llu, last rows done %llu, "
                   "last scalars sent %llu | per block medians: wait %llu, rows %llu, scalars %llu\n",
                   group->rank, maxima[0], maxima[1], maxima[2], maxima[3], median(waits), median(rows), median(fences));
      cudaFree(group->d_debug_times);
   }
   for (size_t r = 0; r < group->peer_blocks.size(); ++r) {
      if (group->opened_with_ipc[r] && group->peer_blocks[r] != nullptr) {
         cudaIpcCloseMemHandle(group->peer_blocks[r]);
      }
   }
   cudaFree(group->d_peer_table);
   cudaFree(group->d_collect_state);
   cudaFree(group->d_local);
   delete group;
   table->shard = nullptr;
}

}  // namespace silo

static uint64_t currentProcessId() {
   return static_cast<uint64_t>(getpid());
}

extern "C" {

int silo_gpu_shard_group_init(silo_gpu_table* table, int column, uint64_t valid_symbol_mask, int rank, int world, void* handle_out) {
   return guarded([&] {
      require(table != nullptr && handle_out != nullptr, "silo_gpu_shard_group_init: NULL argument");
      require(world >= 1 && world <= static_cast<int>(SHARD_MAX_WORLD) && rank >= 0 && rank < world, "silo_gpu_shard_group_init: bad rank / world");
      std::lock_guard<std::mutex> lock(table->mutex);
      SILO_CUDA_CHECK(cudaSetDevice(table->ctx->device));
      require(column >= 0 && static_cast<size_t>(column) < table->columns.size(), "silo_gpu_shard_group_init: bad column index");
      freeShardGroup(table);
      const HostColumn& host = *table->columns[static_cast<size_t>(column)];
      if (host.dev.n_symbols < 64) {
         valid_symbol_mask &= (1ULL << host.dev.n_symbols) - 1;
      }
      require(valid_symbol_mask != 0, "silo_gpu_shard_group_init: no valid symbol");
      auto group = std::make_unique<ShardGroup>();
      group->rank = rank;
      group->world = world;
      group->column = column;
      group->valid_mask = valid_symbol_mask;
      group->n_valid = static_cast<uint32_t>(__builtin_popcountll(valid_symbol_mask));
      group->genome_length = host.dev.genome_length;
      if (std::getenv("SILO_SHARD_DEBUG") != nullptr) {
         group->debug_blocks = diffPadded(host.dev.genome_length) / FIN_THREADS;
         SILO_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&group->d_debug_times), 4 * sizeof(unsigned long long) * group->debug_blocks));
         SILO_CUDA_CHECK(cudaMemset(group->d_debug_times, 0, 4 * sizeof(unsigned long long) * group->debug_blocks));
      }
      group->local_bytes = SHARD_HEADER_BYTES + (rank == 0 ? static_cast<size_t>(SHARD_SLOTS) * world * group->rowsBytes() : 0);
      SILO_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&group->d_local), group->local_bytes));
      SILO_CUDA_CHECK(cudaMemset(group->d_local, 0, group->local_bytes));
      ShardHandle handle{};
      handle.magic = SHARD_HANDLE_MAGIC;
      handle.process_id = currentProcessId();
      handle.pointer = reinterpret_cast<uint64_t>(group->d_local);
      handle.bytes = group->local_bytes;
      SILO_CUDA_CHECK(cudaIpcGetMemHandle(&handle.ipc, group->d_local));
      std::memset(handle_out, 0, SILO_SHARD_HANDLE_BYTES);
      std::memcpy(handle_out, &handle, sizeof(handle));
      table->shard = group.release();
   });
}

int silo_gpu_shard_group_connect(silo_gpu_table* table, const void* handles) {
   return guarded([&] {
      require(table != nullptr && handles != nullptr, "silo_gpu_shard_group_connect: NULL argument");
      std::lock_guard<std::mutex> lock(table->mutex);
      ShardGroup* group = table->shard;
      require(group != nullptr, "silo_gpu_shard_group_connect: call silo_gpu_shard_group_init first");
      SILO_CUDA_CHECK(cudaSetDevice(table->ctx->device));
      auto open = [&](int of_rank, bool* with_ipc) -> uint8_t* {
         ShardHandle handle{};
         std::memcpy(&handle, static_cast<const uint8_t*>(handles) + static_cast<size_t>(of_rank) * SILO_SHARD_HANDLE_BYTES, sizeof(handle));
         require(handle.magic == SHARD_HANDLE_MAGIC, "silo_gpu_shard_group_connect: not a shard handle");
         *with_ipc = false;
         if (of_rank == group->rank) {
            return group->d_local;
         }
         if (handle.process_id == currentProcessId()) {
            // another shard of the same process (tests, several GPUs driven by one process): plain pointer, peer access
            cudaPointerAttributes attributes{};
            SILO_CUDA_CHECK(cudaPointerGetAttributes(&attributes, reinterpret_cast<void*>(handle.pointer)));
            if (attributes.device != table->ctx->device) {
               const cudaError_t status = cudaDeviceEnablePeerAccess(attributes.device, 0);
               if (status != cudaSuccess && status != cudaErrorPeerAccessAlreadyEnabled) {
                  SILO_CUDA_CHECK(status);
               }
               cudaGetLastError();
            }
            return reinterpret_cast<uint8_t*>(handle.pointer);
         }
         void* mapped = nullptr;
         SILO_CUDA_CHECK(cudaIpcOpenMemHandle(&mapped, handle.ipc, cudaIpcMemLazyEnablePeerAccess));
         *with_ipc = true;
         return static_cast<uint8_t*>(mapped);
      };
      group->peer_blocks.assign(static_cast<size_t>(group->world), nullptr);
      group->opened_with_ipc.assign(static_cast<size_t>(group->world), false);
      if (group->rank == 0) {
         for (int r = 0; r < group->world; ++r) {
            bool with_ipc = false;
            group->peer_blocks[static_cast<size_t>(r)] = open(r, &with_ipc);
            group->opened_with_ipc[static_cast<size_t>(r)] = with_ipc;
         }
         group->d_root = group->d_local;
         SILO_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&group->d_peer_table), sizeof(ShardBlockHeader*) * static_cast<size_t>(group->world)));
         SILO_CUDA_CHECK(cudaMemcpy(group->d_peer_table, group->peer_blocks.data(), sizeof(ShardBlockHeader*) * static_cast<size_t>(group->world), cudaMemcpyHostToDevice));
         SILO_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&group->d_collect_state), 4 * sizeof(uint32_t)));
         SILO_CUDA_CHECK(cudaMemset(group->d_collect_state, 0, 4 * sizeof(uint32_t)));
      } else {
         bool with_ipc = false;
         group->peer_blocks[0] = open(0, &with_ipc);
         group->opened_with_ipc[0] = with_ipc;
         group->d_root = group->peer_blocks[0];
      }
   });
}

// every rank's half of a sharded query; with_collect (root): the finalize kernel collects and runs the output pass, in
// the same graph. Returns whether hits were requested and produced on the stream.
static void enqueueShardedQuery(
   silo_gpu_table* table,
   const silo_filter_program* program,
   cudaStream_t stream,
   bool with_collect,
   double min_proportion,
   void* d_summed_counts
);

int silo_gpu_sharded_query_enqueue(silo_gpu_table* table, const silo_filter_program* program, void* cuda_stream) {
   return guarded([&] {
      require(table != nullptr && program != nullptr, "silo_gpu_sharded_query_enqueue: NULL argument");
      std::lock_guard<std::mutex> lock(table->mutex);
      SILO_CUDA_CHECK(cudaSetDevice(table->ctx->device));
      cudaStream_t stream = cuda_stream != nullptr ? static_cast<cudaStream_t>(cuda_stream) : table->ctx->stream;
      enqueueShardedQuery(table, program, stream, false, 0.0, nullptr);
   });
}

static void enqueueShardCollect(silo_gpu_table* table, double min_proportion, bool with_hits, void* d_summed_counts, cudaStream_t stream, bool check_pending = false) {
   ShardGroup* group = table->shard;
   require(group != nullptr && group->rank == 0 && group->d_peer_table != nullptr, "sharded collect: only the connected root (rank 0) collects");
   // (the host-side counters do not see replays of captured graphs: the synchronous collect checks them, the _async one does not)
   require(!check_pending || group->queries_collected < group->queries_enqueued, "sharded collect: no enqueued query is waiting");
   const HostColumn& host = *table->columns[static_cast<size_t>(group->column)];
   HitRequest request;
   request.valid_mask = group->valid_mask;
   request.min_proportion = min_proportion;
   if (with_hits) {
      require(host.dev.global_reference != nullptr, "sharded collect: call silo_gpu_column_set_reference first");
      ensureHitsCapacity(table, host, group->valid_mask, stream);
      request.hits = table->h_hits_pinned;
      request.capacity = static_cast<uint32_t>(table->hits_capacity);
   }
   ShardCollect collect;
   collect.header = reinterpret_cast<ShardBlockHeader*>(group->d_local);
   collect.peers = group->d_peer_table;
   collect.world = static_cast<uint32_t>(group->world);
   collect.n_valid = group->n_valid;
   collect.summed_out = static_cast<uint32_t*>(d_summed_counts);
   shardCollectKernel<<<(host.dev.genome_length + COLLECT_THREADS - 1) / COLLECT_THREADS, COLLECT_THREADS, 0, stream>>>(host.dev, collect, group->d_collect_state, request);
   SILO_CUDA_CHECK(cudaGetLastError());
   table->stats.kernel_launches++;
   group->queries_collected++;
}

static void enqueueShardedQuery(
   silo_gpu_table* table,
   const silo_filter_program* program,
   cudaStream_t stream,
   bool with_collect,
   double min_proportion,
   void* d_summed_counts
) {
   ShardGroup* group = table->shard;
   require(group != nullptr && group->d_root != nullptr, "sharded query: the table is not connected to a shard group");
   require(table->n_chunks > 0, "sharded query: a shard must hold at least one chunk");
   const HostColumn& host = *table->columns[static_cast<size_t>(group->column)];
   if (with_collect) {
      require(host.dev.global_reference != nullptr, "sharded query: call silo_gpu_column_set_reference first");
      ensureHitsCapacity(table, host, group->valid_mask, stream);
   }
   ShardPush push = shardPushOf(*group);
   HitRequest request;
   if (with_collect) {  // the root's finalize kernel sums the ranks and runs the output pass itself
      require(group->rank == 0 && group->d_peer_table != nullptr, "sharded collect: only the connected root (rank 0) collects");
      push.peers = group->d_peer_table;
      push.summed_out = static_cast<uint32_t*>(d_summed_counts);
      request.valid_mask = group->valid_mask;
      request.min_proportion = min_proportion;
      request.hits = table->h_hits_pinned;
      request.capacity = static_cast<uint32_t>(table->hits_capacity);
   }
   const HitRequest* const request_or_null = with_collect ? &request : nullptr;
   const bool trivially_full = program->n_instrs == 1 && program->instrs != nullptr && program->instrs[0].opcode == SILO_OP_PUSH_FULL;
   StagedQuery staged;
   if (trivially_full) {
      push.use_fixed_cardinality = 1;
      push.fixed_cardinality = table->n_rows;
   } else {
      stageQueryLocked(table, program, &staged, group->column, table->d_counts);
      push.filter_scalars = table->query_filter->d_cardinality;
   }
   auto enqueueAll = [&]() {
      if (trivially_full) {
         enqueueMutationCounts(table, group->column, nullptr, table->d_counts, stream, request_or_null, false, false, &push);
      } else {
         enqueueStagedQuery(table, staged, stream);
         enqueueMutationCounts(table, group->column, table->query_filter, table->d_counts, stream, request_or_null, false, true, &push);
      }
   };
   cudaGraphExec_t replay = nullptr;
   if (!trivially_full) {
      // the shape of the query (slot and generation live in device memory, not in kernel parameters)
      std::string key(reinterpret_cast<const char*>(staged.params), sizeof(staged.params));
      const uint64_t scalars[] = {0x5348415244ULL, staged.staged_bytes, staged.shared_bytes, static_cast<uint64_t>(group->column), group->valid_mask,
                                  reinterpret_cast<uint64_t>(table->h_hits_pinned), table->hits_capacity, with_collect ? 1ULL : 0ULL,
                                  reinterpret_cast<uint64_t>(host.dev.containers), host.dev.n_segments,
                                  reinterpret_cast<uint64_t>(host.dev.global_reference), reinterpret_cast<uint64_t>(table->d_counts),
                                  reinterpret_cast<uint64_t>(table->d_work_items), reinterpret_cast<uint64_t>(table->d_coverage_diff),
                                  reinterpret_cast<uint64_t>(group->d_root), reinterpret_cast<uint64_t>(group->d_local),
                                  reinterpret_cast<uint64_t>(d_summed_counts)};
      key.append(reinterpret_cast<const char*>(scalars), sizeof(scalars));
      key.append(reinterpret_cast<const char*>(&min_proportion), sizeof(min_proportion));
      replay = queryGraphFor(table, std::move(key), stream, enqueueAll);
   }
   if (replay != nullptr) {
      SILO_CUDA_CHECK(cudaGraphLaunch(replay, stream));
      table->stats.kernel_launches += 4;
   } else {
      enqueueAll();
   }
   group->queries_enqueued++;
   if (with_collect) {  // (host-side bookkeeping of the synchronous collect's "is a query waiting" check)
      group->queries_collected++;
   }
}

// reads the result of a collect with output pass that has completed on the stream
static void readShardedHits(silo_gpu_table* table, const silo_mutation_hit** hits, uint64_t* n_hits, uint64_t* cardinality) {
   const silo_mutation_hit header = table->h_hits_pinned[0];
   if ((header.symbol & 2u) != 0) {
      throw ApiError(SILO_E_CUDA, "sharded query: a rank of the shard group did not deliver its counts in time (or an earlier query of the group is still uncollected)");
   }
   if (header.symbol != 0) {
      throw ApiError(SILO_E_OUT_OF_LAYOUT, "a leaf bitmap holds row ids outside the row layout");
   }
   const uint64_t count = std::min<uint64_t>(header.position, table->hits_capacity);
   *hits = publishHits(table->h_hits_pinned + 1, count);
   *n_hits = count;
   if (cardinality != nullptr) {
      *cardinality = header.count | (static_cast<unsigned long long>(header.total) << 32);
   }
}

int silo_gpu_sharded_query_hits(
   silo_gpu_table* table,
   const silo_filter_program* program,
   double min_proportion,
   void* d_summed_counts,
   const silo_mutation_hit** hits,
   uint64_t* n_hits,
   uint64_t* cardinality
) {
   return guarded([&] {
      require(table != nullptr && program != nullptr && hits != nullptr && n_hits != nullptr, "silo_gpu_sharded_query_hits: NULL argument");
      std::lock_guard<std::mutex> lock(table->mutex);
      SILO_CUDA_CHECK(cudaSetDevice(table->ctx->device));
      cudaStream_t stream = table->ctx->stream;
      require(table->shard != nullptr && table->shard->rank == 0, "silo_gpu_sharded_query_hits: only the root (rank 0) collects");
      enqueueShardedQuery(table, program, stream, true, min_proportion, d_summed_counts);
      SILO_CUDA_CHECK(cudaStreamSynchronize(stream));
      readShardedHits(table, hits, n_hits, cardinality);
   });
}

int silo_gpu_sharded_collect_async(silo_gpu_table* table, void* d_summed_counts, void* cuda_stream) {
   return guarded([&] {
      require(table != nullptr, "silo_gpu_sharded_collect_async: NULL argument");
      std::lock_guard<std::mutex> lock(table->mutex);
      SILO_CUDA_CHECK(cudaSetDevice(table->ctx->device));
      cudaStream_t stream = cuda_stream != nullptr ? static_cast<cudaStream_t>(cuda_stream) : table->ctx->stream;
      enqueueShardCollect(table, 0.0, false, d_summed_counts, stream);
   });
}

int silo_gpu_sharded_collect(
   silo_gpu_table* table,
   double min_proportion,
   void* d_summed_counts,
   void* cuda_stream,
   const silo_mutation_hit** hits,
   uint64_t* n_hits,
   uint64_t* cardinality
) {
   return guarded([&] {
      require(table != nullptr && hits != nullptr && n_hits != nullptr, "silo_gpu_sharded_collect: NULL argument");
      std::lock_guard<std::mutex> lock(table->mutex);
      SILO_CUDA_CHECK(cudaSetDevice(table->ctx->device));
      cudaStream_t stream = cuda_stream != nullptr ? static_cast<cudaStream_t>(cuda_stream) : table->ctx->stream;
      enqueueShardCollect(table, min_proportion, true, d_summed_counts, stream, true);
      SILO_CUDA_CHECK(cudaStreamSynchronize(stream));
      readShardedHits(table, hits, n_hits, cardinality);
   });
}

void silo_gpu_shard_group_free(silo_gpu_table* table) {
   if (table == nullptr) {
      return;
   }
   std::lock_guard<std::mutex> lock(table->mutex);
   cudaSetDevice(table->ctx->device);
   freeShardGroup(table);
}

// Several sequence columns under ONE filter (AminoAcidMutations over all genes; the producer of mutations_node.cpp:372-428
// loops the columns of one query): the program is evaluated once, then every column runs work list, coverage, container
// and finalize kernels back to back on the stream, each writing its tuples into its own region of the page-locked
// buffer; one synchronisation for the whole query, and the whole sequence is one replayed CUDA graph per query shape.
int silo_gpu_query_mutation_hits_columns(
   silo_gpu_table* table,
   const silo_filter_program* program,
   silo_column_hits* columns,
   uint32_t n_columns,
   double min_proportion,
   uint64_t* cardinality
) {
   return guarded([&] {
      require(table != nullptr && program != nullptr && columns != nullptr && n_columns >= 1, "silo_gpu_query_mutation_hits_columns: NULL argument");
      std::lock_guard<std::mutex> lock(table->mutex);
      SILO_CUDA_CHECK(cudaSetDevice(table->ctx->device));
      cudaStream_t stream = table->ctx->stream;
      // one region per column: [header][tuples...], sized for the worst case
      std::vector<uint64_t> region_begin(n_columns + 1, 0);
      for (uint32_t c = 0; c < n_columns; ++c) {
         require(columns[c].column >= 0 && static_cast<size_t>(columns[c].column) < table->columns.size(), "silo_gpu_query_mutation_hits_columns: bad column index");
         const HostColumn& host = *table->columns[static_cast<size_t>(columns[c].column)];
         require(host.dev.global_reference != nullptr, "silo_gpu_query_mutation_hits_columns: call silo_gpu_column_set_reference first");
         if (host.dev.n_symbols < 64) {
            columns[c].valid_symbol_mask &= (1ULL << host.dev.n_symbols) - 1;
         }
         region_begin[c + 1] = region_begin[c] + 1 + static_cast<uint64_t>(__builtin_popcountll(columns[c].valid_symbol_mask)) * host.dev.genome_length;
      }
      ensureHitsTuples(table, region_begin[n_columns], stream);
      const bool trivially_full = program->n_instrs == 1 && program->instrs != nullptr && program->instrs[0].opcode == SILO_OP_PUSH_FULL;
      StagedQuery staged;
      if (!trivially_full) {
         stageQueryLocked(table, program, &staged, columns[0].column, table->d_counts);
      }
      auto enqueueAll = [&]() {
         if (!trivially_full) {
            enqueueStagedQuery(table, staged, stream);
         }
         for (uint32_t c = 0; c < n_columns; ++c) {
            const HostColumn& host = *table->columns[static_cast<size_t>(columns[c].column)];
            HitRequest request;
            request.hits = table->h_hits_pinned + region_begin[c];
            request.capacity = static_cast<uint32_t>(region_begin[c + 1] - region_begin[c] - 1);
            request.valid_mask = columns[c].valid_symbol_mask;
            request.min_proportion = min_proportion;
            (void) host;
            if (trivially_full) {
               enqueueMutationCounts(table, columns[c].column, nullptr, table->d_counts, stream, &request);
            } else {
               request.filter_scalars = table->query_filter->d_cardinality;
               request.keep_scalars = c + 1 < n_columns ? 1u : 0u;
               enqueueMutationCounts(table, columns[c].column, table->query_filter, table->d_counts, stream, &request, false, c == 0);
            }
         }
      };
      try {
         cudaGraphExec_t replay = nullptr;
         if (!trivially_full) {
            std::string key(reinterpret_cast<const char*>(staged.params), sizeof(staged.params));
            std::vector<uint64_t> scalars = {0x434F4C53ULL, staged.staged_bytes, staged.shared_bytes, n_columns,
                                             reinterpret_cast<uint64_t>(table->h_hits_pinned), table->hits_capacity,
                                             reinterpret_cast<uint64_t>(table->d_counts), reinterpret_cast<uint64_t>(table->d_work_items),
                                             reinterpret_cast<uint64_t>(table->d_coverage_diff)};
            for (uint32_t c = 0; c < n_columns; ++c) {
               const HostColumn& host = *table->columns[static_cast<size_t>(columns[c].column)];
               scalars.insert(scalars.end(), {static_cast<uint64_t>(columns[c].column), columns[c].valid_symbol_mask,
                                              reinterpret_cast<uint64_t>(host.dev.containers), host.dev.n_segments,
                                              reinterpret_cast<uint64_t>(host.dev.global_reference)});
            }
            key.append(reinterpret_cast<const char*>(scalars.data()), scalars.size() * sizeof(uint64_t));
            key.append(reinterpret_cast<const char*>(&min_proportion), sizeof(min_proportion));
            replay = queryGraphFor(table, std::move(key), stream, enqueueAll);
         }
         if (replay != nullptr) {
            SILO_CUDA_CHECK(cudaGraphLaunch(replay, stream));
            table->stats.kernel_launches += 1 + 4ULL * n_columns;
         } else {
            enqueueAll();
         }
         SILO_CUDA_CHECK(cudaStreamSynchronize(stream));
      } catch (...) {
         cudaStreamSynchronize(stream);
         if (!trivially_full) {
            cudaMemsetAsync(table->query_filter->d_cardinality, 0, 32, stream);
            cudaMemsetAsync(table->d_work_state, 0, 4 * sizeof(uint32_t), stream);
            cudaStreamSynchronize(stream);
         }
         throw;
      }
      unsigned long long host_cardinality = trivially_full ? table->n_rows : 0;
      for (uint32_t c = 0; c < n_columns; ++c) {
         silo_mutation_hit* const region = table->h_hits_pinned + region_begin[c];
         const silo_mutation_hit header = region[0];
         if (!trivially_full) {
            if (header.symbol != 0) {
               throw ApiError(SILO_E_OUT_OF_LAYOUT, "a leaf bitmap holds row ids outside the row layout");
            }
            host_cardinality = header.count | (static_cast<unsigned long long>(header.total) << 32);
         }
         const uint64_t count = std::min<uint64_t>(header.position, region_begin[c + 1] - region_begin[c] - 1);
         columns[c].hits = publishHits(region + 1, count, c);
         columns[c].n_hits = count;
      }
      if (cardinality != nullptr) {
         *cardinality = host_cardinality;
      }
   });
}

int silo_gpu_query_mutation_counts(
   silo_gpu_table* table,
   const silo_filter_program* program,
   int column,
   uint64_t symbol_mask,
   uint32_t* counts,
   uint64_t* cardinality
) {
   return guarded([&] {
      require(program != nullptr, "silo_gpu_query_mutation_counts: program is NULL");
      mutationCountsToHost(table, column, nullptr, program, symbol_mask, counts, cardinality);
   });
}

int silo_gpu_mutation_counts(
   silo_gpu_table* table,
   int column,
   const silo_gpu_filter* filter,
   uint32_t* counts
) {
   return guarded([&] { mutationCountsToHost(table, column, filter, nullptr, ~0ULL, counts, nullptr); });
}

int silo_gpu_mutation_counts_symbols(
   silo_gpu_table* table,
   int column,
   const silo_gpu_filter* filter,
   uint64_t symbol_mask,
   uint32_t* counts
) {
   return guarded([&] { mutationCountsToHost(table, column, filter, nullptr, symbol_mask, counts, nullptr); });
}

}  // extern "C"
