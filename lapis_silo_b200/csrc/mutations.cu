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
#include <cstring>

#include <chrono>

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
// Persistent CTAs, ONE per SM (1,024 threads, 56 registers: a 256-thread block of the coverage kernel still
// fits beside it). Two CTAs of 17 warps per SM do not run at the same speed -- the warp schedulers favour one
// of them, half of the CTAs visited ~66 stages in the time the other half visited ~45 -- and when the work
// list ran out the slow ones still owned what they had claimed ahead: ~10 us of a 72 us kernel. One CTA per
// SM has no such pair. Warp 0 is the producer (one thread): it claims batches of work items (segments)
// from a grid-wide counter and streams each segment's block [descriptors | payloads] into a ring of
// shared-memory stages with one 1-D bulk (TMA) copy that completes on an mbarrier. The 31 consumer
// warps take one piece (<= 1 KiB of payload) each: a warp pulls its piece into REGISTERS, hands the
// stage back to the producer, and only then ANDs the piece with the chunk's filter tile held in
// shared memory, issuing one RED per piece with a non-zero count.
//
// Everything the consumers touch is addressed with 32-bit shared-memory addresses kept in registers
// (inline PTX loads): the kernel is bound by instruction issue and by the 16-lane ALU pipe, and the
// generic-pointer arithmetic of plain C++ costs as many instructions per piece as the lookups.
// ---------------------------------------------------------------------------------------------

#ifndef SILO_K1_STAGES
#define SILO_K1_STAGES 6
#endif
constexpr int K1_STAGES = SILO_K1_STAGES;
constexpr int K1_CONSUMER_WARPS = 31;
constexpr int K1_CONSUMER_THREADS = K1_CONSUMER_WARPS * 32;
constexpr int K1_THREADS = K1_CONSUMER_THREADS + 32;  // warp 0 = bulk-copy producer; 1,024 threads: one CTA per SM
constexpr uint32_t K1_BATCH_DEFAULT = 4;               // work items claimed per atomicAdd (<= 32)
constexpr uint32_t K1_STOP = 0xFFFFFFFFu;              // meta.x marker: no more work
constexpr uint32_t TILE32_WORDS = 2 * TILE_WORDS;
constexpr uint32_t TILE_BUFFER_BYTES = (TILE32_WORDS + 4) * 4;  // [2048] = the zero pad word

struct __align__(16) K1Stage {  // one segment block: n descriptors, then their payloads (<= this size)
   uint8_t payload[SEG_PAYLOAD_BYTES];
   DevContainer descs[SEG_MAX_DESCS];
};

// per-stage control block: meta is written by the producer before it arms `full`
struct __align__(16) K1Control {
   uint32_t desc_count;  // K1_STOP: no more work
   uint32_t base4;       // slab offset (4-byte units) of the stage's payload[0]
   uint32_t flags;       // K1_NEW_TILE | K1_TILE_SLOT
   uint32_t pad0;
   uint64_t full;   // the stage's bulk copies have landed
   uint64_t empty;  // every consumer warp has pulled its pieces of the stage into registers
   uint64_t done;   // ... and has finished the tile lookups for them
   uint64_t pad1;
};
static_assert(sizeof(K1Control) == 48);
constexpr uint32_t K1_CTRL_FULL = 16;
constexpr uint32_t K1_CTRL_EMPTY = 24;
constexpr uint32_t K1_CTRL_DONE = 32;
constexpr uint32_t K1_NEW_TILE = 1;    // the stage's bulk copies also (re)loaded the filter tile
constexpr uint32_t K1_TILE_SLOT = 2;   // which of the two filter-tile buffers the stage reads

struct __align__(16) K1Dynamic {
   K1Stage stages[K1_STAGES];
   K1Control control[K1_STAGES];
   // a warp pulls two whole 512-byte regions whatever the size of its piece: the reads behind a short piece at the
   // end of the last stage must stay inside the CTA's shared-memory window
   uint8_t overrun_pad[1024 - K1_STAGES * sizeof(K1Control) % 1024];
};
constexpr uint32_t K1_CONTROL_OFFSET = sizeof(K1Stage) * K1_STAGES;

__device__ __forceinline__ uint32_t warpSum(uint32_t value) {
   return __reduce_add_sync(0xFFFFFFFFu, value);
}

// ---- shared memory by 32-bit address ----------------------------------------------------------
// Reads of stage data and of the filter tile. Not volatile: every address derives from words that were
// loaded after the stage's full barrier was observed (a volatile asm with a memory clobber).
#ifndef SILO_K1_PROBE
#define SILO_K1_PROBE 0  // 1: tile lookups all hit one word (no bank conflicts); 2: no tile lookups at all
#endif
__device__ __forceinline__ uint32_t lds32(uint32_t address) {
   uint32_t value;
#if SILO_K1_PROBE == 1
   asm("ld.shared.u32 %0, [%1];" : "=r"(value) : "r"(address & 0x3u));
#elif SILO_K1_PROBE == 2
   value = address;
#else
   asm("ld.shared.u32 %0, [%1];" : "=r"(value) : "r"(address));
#endif
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
#ifndef SILO_WAIT_HINT_NS
#define SILO_WAIT_HINT_NS 0x989680u
#endif
__device__ __forceinline__ void mbarWaitAt(uint32_t address, uint32_t parity) {
#if SILO_WAIT_HINT_NS == 0
   asm volatile(  // spin on the phase test
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(address),
      "r"(parity)
      : "memory"
   );
#else
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
      "r"(SILO_WAIT_HINT_NS)
      : "memory"
   );
#endif
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
// lane 0 adds `value` to *target if value != 0, without a branch
__device__ __forceinline__ void redAddLane0(uint32_t* target, uint32_t value, uint32_t lane) {
   asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.eq.u32 p, %2, 0;\n"
      "setp.ne.and.u32 p, %1, 0, p;\n"
      "@p red.global.add.u32 [%0], %1;\n"
      "}\n" ::"l"(target),
      "r"(value), "r"(lane)
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

// The consumers are bound by instruction issue and the ALU pipe, so every bit test is spelled with
// the fewest instructions the ISA offers: a byte offset into the tile is one IMAD.HI with
// accumulate on the FMA pipe (x * 2^k >> 32, plus the tile's shared-memory address), shifts take
// their amount from the low five bits of a register (funnel shifts in wrap mode), and the
// multipliers live in registers the compiler cannot see through (otherwise it turns the
// multiplications back into shift + add on the ALU pipe).
struct Multipliers {
   uint32_t one;         //  x * 1 + acc: an addition on the FMA pipe
   uint32_t two;         //  x * 2 >> 32 = x >> 31
   uint32_t two_pow_13;  // (x & 0xFFE00000) * 2^13 >> 32 = (x >> 21) * 4
   uint32_t two_pow_14;  //  x * 2^14 >> 32 = x >> 18
   uint32_t two_pow_16;  //  x * 2^16 >> 32 = x >> 16
   uint32_t two_pow_27;  //  x * 2^27 >> 32 = x >> 5
   uint32_t two_pow_29;  // (x & 0xFFE0) * 2^29 >> 32 = ((x & 0xFFFF) >> 5) * 4
};

__device__ __forceinline__ uint32_t madHi(uint32_t a, uint32_t b, uint32_t c) {
   uint32_t d;
   asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
   return d;
}
__device__ __forceinline__ uint32_t madLo(uint32_t a, uint32_t b, uint32_t c) {
   uint32_t d;
   asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
   return d;
}
__device__ __forceinline__ uint32_t mulHi(uint32_t a, uint32_t b) {
   uint32_t d;
   asm("mul.hi.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
   return d;
}

// Two stored array values packed in one word -> the tile words of their rows, each shifted so that
// the row's bit sits in bit 31. A stored value is row ^ 31: its low five bits are 31 - (row & 31), the
// left shift that takes the row's bit to the top (funnel shifts use the low five bits of the amount).
// The caller accumulates bit 31 with one IMAD.HI (x * 2 >> 32 + acc) on the FMA pipe, so a lookup
// costs the 16-lane ALU pipe only the address mask and the shift.
__device__ __forceinline__ void pairTopBits(uint32_t tile_address, const Multipliers& k, uint32_t pair, uint32_t& lo_top, uint32_t& hi_top) {
   const uint32_t lo_word = lds32(madHi(pair & 0x0000FFE0u, k.two_pow_29, tile_address));
   const uint32_t hi_word = lds32(madHi(pair & 0xFFE00000u, k.two_pow_13, tile_address));
   lo_top = __funnelshift_l(0u, lo_word, pair);                        // lo_word << (pair & 31)
   hi_top = __funnelshift_l(0u, hi_word, mulHi(pair, k.two_pow_16));   // hi_word << ((pair >> 16) & 31)
}
__device__ __forceinline__ uint32_t addTopBit(uint32_t value, const Multipliers& k, uint32_t accumulator) {
   return madHi(value, k.two, accumulator);  // (value >> 31) + accumulator
}

// one KIND_RUNS_W entry -> rows of the run that are set in the tile
__device__ __forceinline__ uint32_t runEntryCount(uint32_t tile_address, const Multipliers& k, uint32_t entry) {
   const uint32_t word = lds32(madHi(entry, k.two_pow_14, tile_address));       // tile32[entry >> 20]
   const uint32_t from_first = __funnelshift_r(word, 0u, entry);                 // word >> (entry & 31)
   return __popc(__funnelshift_l(0u, from_first, mulHi(entry, k.two_pow_27)));   // << (32 - length)
}

// lookups of one array region held in registers (count values in P = ceil(count/8) lanes)
__device__ __forceinline__ uint32_t arrayRegionCount(
   uint32_t tile_address,
   const Multipliers& k,
   const uint4& eight,
   uint32_t count,
   uint32_t lane
) {
   // `count` values in P = ceil(count / 8) lanes; in a partial region the slots behind the last value
   // repeat it (pool.cu encodeArrayPiece). Every slot is looked up -- no test per slot --, and the
   // repeats are taken out again: slot 7 of lane P-1 always holds the region's last value, so that
   // lane subtracts (8 P - count) times its bit. Lanes >= P hold other bytes of the stage; their
   // lookups stay inside the tile (11-bit word index) and are dropped.
   uint32_t t0, t1, t2, t3, t4, t5, t6, t7;
   pairTopBits(tile_address, k, eight.x, t0, t1);
   pairTopBits(tile_address, k, eight.y, t2, t3);
   pairTopBits(tile_address, k, eight.z, t4, t5);
   pairTopBits(tile_address, k, eight.w, t6, t7);
   // two accumulation chains
   const uint32_t even = addTopBit(t6, k, addTopBit(t4, k, addTopBit(t2, k, addTopBit(t0, k, 0u))));
   const uint32_t odd = addTopBit(t7, k, addTopBit(t5, k, addTopBit(t3, k, addTopBit(t1, k, 0u))));
   uint32_t local = even + odd;
   if (count != ARRAY_REGION_VALUES) {
      const uint32_t lanes = arrayRegionLanes(count);
      if (lane + 1 == lanes) {
         local -= (8u * lanes - count) * (t7 >> 31);
      }
      local = lane < lanes ? local : 0u;
   }
   return local;
}

// one run region held in registers (padding entries count nothing; lanes beyond the region hold junk)
__device__ __forceinline__ uint32_t runsRegionCount(
   uint32_t tile_address,
   const Multipliers& k,
   const uint4& four,
   uint32_t count,
   uint32_t lane
) {
   if (lane >= runsRegionLanes(count)) {
      return 0;
   }
   // (sums on the FMA pipe: x * 1 + acc)
   const uint32_t first_two = madLo(runEntryCount(tile_address, k, four.y), k.one, runEntryCount(tile_address, k, four.x));
   const uint32_t last_two = madLo(runEntryCount(tile_address, k, four.w), k.one, runEntryCount(tile_address, k, four.z));
   return madLo(first_two, k.one, last_two);
}

// |piece AND tile| for a piece whose (at most) two 512-byte regions sit in registers: lane L holds
// bytes [16 L, 16 L + 16) of each region. desc = the descriptor's four words.
__device__ __forceinline__ uint32_t pieceFromRegisters(
   const uint4& desc,
   uint32_t kind,
   const uint4& first,
   const uint4& second,
   uint32_t tile_address,  // shared: 2049 words ([2048] = 0)
   const Multipliers& k,
   uint32_t lane,
   uint32_t lane16
) {
   uint32_t local = 0;
   if (kind == KIND_ARRAY_T) {
      const uint32_t n = (desc.z & 0xFFFFu) + 1u;
      local = arrayRegionCount(tile_address, k, first, min(n, ARRAY_REGION_VALUES), lane);
      if (n > ARRAY_REGION_VALUES) {
         local += arrayRegionCount(tile_address, k, second, n - ARRAY_REGION_VALUES, lane);
      }
   } else if (kind == KIND_RUNS_W) {
      const uint32_t n = desc.w;
      local = runsRegionCount(tile_address, k, first, min(n, RUNS_REGION_ENTRIES), lane);
      if (n > RUNS_REGION_ENTRIES) {
         local += runsRegionCount(tile_address, k, second, n - RUNS_REGION_ENTRIES, lane);
      }
   } else if (kind == KIND_BITSET) {  // 128 words: 64 vectors, two per lane
      const uint32_t window = tile_address + (desc.w & 0xFFFFu) * 8u + lane16;
      const uint4 a = lds128(window);
      const uint4 b = lds128(window + 512);
      local = __popc(first.x & a.x) + __popc(first.y & a.y) + __popc(first.z & a.z) + __popc(first.w & a.w) +
              __popc(second.x & b.x) + __popc(second.y & b.y) + __popc(second.z & b.z) + __popc(second.w & b.w);
   } else if (kind == KIND_INLINE) {  // one or two values inside the descriptor
      if (lane <= (desc.z & 0xFFFFu)) {
         const uint32_t value = (desc.w >> (16 * lane)) & 0xFFFFu;
         local = (lds32(tile_address + ((value >> 5) << 2)) >> (value & 31u)) & 1u;
      }
   }
   return warpSum(local);
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
// (what the bulk-copy pipeline alone can stream); 2 = no atomics; 3 = touch the payload only.
template <int MODE>
__global__ void __maxnreg__(56) containerAndCountKernel(
   DevColumn column,
   const uint64_t* __restrict__ filter_words,   // [n_chunks * 1024]
   uint32_t* __restrict__ work_state,           // [0] number of work items, [1] grid-wide claim counter (zero at launch)
   const DevSegment* __restrict__ work_items,   // the segment record of every work item (prepareQueryKernel)
   uint32_t* __restrict__ counts,               // [n_symbols * genome_length]
   uint32_t claim_batch,                        // work items per claim
   uint32_t tail_batches                        // the last tail_batches * gridDim.x * claim_batch items are claimed one by one
) {
   // Two tile buffers: the producer loads the next chunk's tile while stages of the current chunk
   // are still being consumed, so a chunk switch does not drain the pipeline.
   __shared__ __align__(16) uint32_t tile_buffers[2][TILE32_WORDS + 4];  // [2048] = zero pad word
   extern __shared__ __align__(128) uint8_t smem_raw[];
   K1Dynamic& sh = *reinterpret_cast<K1Dynamic*>(smem_raw);

   const uint32_t warp = threadIdx.x >> 5;
   const uint32_t lane = threadIdx.x & 31;
   const uint32_t ring_address = smemAddr(smem_raw);
   const uint32_t control_address = ring_address + K1_CONTROL_OFFSET;
   const uint32_t tile_address0 = smemAddr(tile_buffers[0]);

   if (threadIdx.x == 0) {
      for (int s = 0; s < K1_STAGES; ++s) {
         mbarInit(&sh.control[s].full, 1);
         mbarInit(&sh.control[s].empty, K1_CONSUMER_WARPS);
         mbarInit(&sh.control[s].done, K1_CONSUMER_WARPS);
      }
      fenceBarrierInit();
   }
   if (threadIdx.x < 8) {
      tile_buffers[threadIdx.x >> 2][TILE32_WORDS + (threadIdx.x & 3)] = 0;
   }
   __syncthreads();

   if (warp == 0) {
      // ---------------- producer: ONE thread ------------------------------------------------------
      // Lane 0 alone claims batches of work items (segments) from the grid-wide counter, fetches their
      // 16-byte records and issues one bulk copy per ring stage. A warp-wide producer (lane j preparing
      // stage j, the copies issued lane after lane) spent 0.95 us per stage whatever the stage size, the
      // batch size or the number of copies: its shuffles, ballots, warp syncs and barrier tests all queue
      // behind the consumers' shared-memory loads, ~40 cycles per instruction, and the consumers waited for
      // data 16 % of their time while the producer waited for a free stage only 12 % of its. A single
      // thread needs no cross-lane instruction at all.
      // Pipeline: the records of batch k + 2 are loaded right after the copies of batch k were issued and
      // are used after those of batch k + 1; three claims are in flight (an atomic has three batches to
      // come back). The first two batches of every CTA are fixed (no atomic in front of the first
      // copies); near the end of the list claims shrink to single items, so that the last CTAs to finish
      // are a few stages, not a few batches, behind the others.
      if (lane != 0) {
         return;
      }
      constexpr uint32_t HELD = 4;  // records of one batch held in registers
      const uint32_t batch_items = min(claim_batch, HELD);
      uint32_t* const work_counter = work_state + 1;
      const uint32_t static_items = 2 * gridDim.x * batch_items;
      const uint32_t capacity = column.n_segments;
      const uint32_t total = work_state[0];
      const uint32_t tail_items = tail_batches * gridDim.x * batch_items;
      const uint32_t tail_begin = total > tail_items ? total - tail_items : 0;
      const uint4* const records = reinterpret_cast<const uint4*>(work_items);
      // (records behind the end of the list are stale or uninitialised and never used: `n` below cuts them off)
      auto load = [&](uint3 (&into)[HELD], uint32_t first, uint32_t size) {
#pragma unroll
         for (uint32_t j = 0; j < HELD; ++j) {
            if (j < size && first + j < capacity) {
               const uint4 record = records[first + j];
               into[j] = make_uint3(record.x, record.y, record.w);  // (desc_begin is unused)
            }
         }
      };
      long long producer_waited = 0;  // MODE 4: cycles spent waiting for a free stage
      const long long producer_begin = MODE == 4 ? clock64() : 0;
      uint3 current[HELD];
      uint3 next[HELD];
      uint32_t current_first = blockIdx.x * batch_items;
      uint32_t current_size = batch_items;
      uint32_t next_first = (gridDim.x + blockIdx.x) * batch_items;
      uint32_t next_size = batch_items;
      load(current, current_first, current_size);
      load(next, next_first, next_size);
      uint32_t claim_a = atomicAdd(work_counter, batch_items) + static_items;
      uint32_t claim_b = atomicAdd(work_counter, batch_items) + static_items;
      uint32_t claim_c = atomicAdd(work_counter, batch_items) + static_items;
      uint32_t size_a = batch_items;
      uint32_t size_b = batch_items;
      uint32_t size_c = batch_items;
      uint32_t tile_chunk = 0xFFFFFFFFu;  // chunk whose tile the latest stage reads
      uint32_t tile_slot = 1;             // ... and the buffer it sits in
      uint32_t tile_first_stage = 0;      // first stage that reads it
      uint32_t it = 0;                    // stages issued so far
      uint32_t stage = 0;                 // it % K1_STAGES
      uint32_t round = 0;                 // it / K1_STAGES
      // one ring stage: {payload_offset16, block bytes, chunk | pieces << 16} of the DevSegment
      auto issue = [&](const uint3& record) {
         const uint32_t my_control = control_address + stage * static_cast<uint32_t>(sizeof(K1Control));
         const uint32_t my_ring = ring_address + stage * static_cast<uint32_t>(sizeof(K1Stage));
         if (round > 0) {
            const long long wait_begin = MODE == 4 ? clock64() : 0;
            mbarWaitAt(my_control + K1_CTRL_EMPTY, (round - 1) & 1u);  // => every stage <= it - K1_STAGES is pulled
            if (MODE == 4) {
               producer_waited += clock64() - wait_begin;
            }
         }
         const uint32_t chunk = record.z & 0xFFFFu;
         const bool new_tile = chunk != tile_chunk;
         if (new_tile) {
            // The new tile goes into the OTHER buffer, last read by the lookups of the stages before
            // tile_first_stage. A warp hands a stage back BEFORE it does the lookups, so the stage's empty
            // barrier says nothing about the tile: wait for the done barriers. (Stages below it - K1_STAGES
            // are implied: a warp pulls stage q + K1_STAGES only after it has finished stage q, and the
            // empty wait above covered it - K1_STAGES.)
            for (uint32_t prev = it >= K1_STAGES ? it - K1_STAGES : 0; prev < tile_first_stage; ++prev) {
               mbarWaitAt(control_address + (prev % K1_STAGES) * static_cast<uint32_t>(sizeof(K1Control)) + K1_CTRL_DONE, (prev / K1_STAGES) & 1u);
            }
            tile_slot ^= 1u;
            tile_first_stage = it;
            tile_chunk = chunk;
         }
         sts128(my_control, make_uint4(record.z >> 16, record.x << 2, (new_tile ? K1_NEW_TILE : 0u) | (tile_slot != 0 ? K1_TILE_SLOT : 0u), 0u));
         mbarExpectTxAt(my_control + K1_CTRL_FULL, record.y + (new_tile ? TILE_BYTES : 0u));
         if (new_tile) {
            bulkLoadAt(tile_address0 + tile_slot * TILE_BUFFER_BYTES, filter_words + static_cast<size_t>(chunk) * TILE_WORDS, TILE_BYTES, my_control + K1_CTRL_FULL);
         }
         // the segment's block [descriptors | payloads] in one copy
         bulkLoadAt(my_ring, column.payload + (static_cast<uint64_t>(record.x) << 4), record.y, my_control + K1_CTRL_FULL);
         ++it;
         const bool wrap = stage == K1_STAGES - 1;
         stage = wrap ? 0u : stage + 1u;
         round += wrap ? 1u : 0u;
      };
      while (current_first < total) {
         const uint32_t n = min(current_size, total - current_first);
#pragma unroll
         for (uint32_t j = 0; j < HELD; ++j) {
            if (j < n) {
               issue(current[j]);
            }
         }
#pragma unroll
         for (uint32_t j = 0; j < HELD; ++j) {
            current[j] = next[j];
         }
         current_first = next_first;
         current_size = next_size;
         next_first = claim_a;  // (waits for the atomic issued three batches ago)
         next_size = size_a;
         load(next, next_first, next_size);
         claim_a = claim_b;
         size_a = size_b;
         claim_b = claim_c;
         size_b = size_c;
         size_c = next_first >= tail_begin ? 1u : batch_items;
         claim_c = atomicAdd(work_counter, size_c) + static_items;
      }
      if (MODE == 4) {
         const uint32_t row = 15 * column.genome_length;
         // per CTA (profiles/k1_wait_probe.py): wall clock (ns, low word) at which this producer ran out of work,
         // the SM it ran on, the stages it issued
         uint32_t now;
         uint32_t smid;
         asm volatile("mov.u32 %0, %%globaltimer_lo;" : "=r"(now));
         asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
         counts[row + 1024 + blockIdx.x] = now;
         counts[row + 2048 + blockIdx.x] = smid;
         counts[row + 3072 + blockIdx.x] = it;
         atomicAdd(&counts[row + 5], static_cast<uint32_t>(producer_waited >> 6));
         atomicAdd(&counts[row + 6], static_cast<uint32_t>((clock64() - producer_begin) >> 6));
         atomicAdd(&counts[row + 7], 1u);
      }
      // tell the consumers that nothing follows
      const uint32_t stop_control = control_address + stage * static_cast<uint32_t>(sizeof(K1Control));
      if (round > 0) {
         mbarWaitAt(stop_control + K1_CTRL_EMPTY, (round - 1) & 1u);
      }
      sts128(stop_control, make_uint4(K1_STOP, 0u, 0u, 0u));
      mbarArriveAt(stop_control + K1_CTRL_FULL);
      return;
   }

   // ---------------- consumers: 16 warps ----------------------------------------------------------
   // A stage holds at most one piece per consumer warp (SEG_MAX_DESCS == K1_CONSUMER_WARPS), handed out
   // round-robin. A warp pulls its piece (descriptor + at most two 512-byte regions) into registers,
   // hands the stage back to the producer at once, and only then does the tile lookups: the ring's
   // stages are in flight again while the arithmetic runs. The loop body is written for the fewest
   // instructions per stage visit -- the kernel is bound by instruction issue.
   static_assert(SEG_MAX_DESCS == K1_CONSUMER_WARPS, "one piece per consumer warp and stage");
   const uint32_t cwarp = warp - 1;
   const uint32_t lane16 = lane * 16;
   const uint32_t genome_length = column.genome_length;
   // opaque to the compiler (gridDim.y is 1): keeps the IMAD.HI forms, see Multipliers
   Multipliers k;
   k.one = gridDim.y;
   k.two = gridDim.y << 1;
   k.two_pow_13 = gridDim.y << 13;
   k.two_pow_14 = gridDim.y << 14;
   k.two_pow_16 = gridDim.y << 16;
   k.two_pow_27 = gridDim.y << 27;
   k.two_pow_29 = gridDim.y << 29;
   // base addresses as opaque register values (the compiler would otherwise re-derive the shared
   // window from special registers at every use)
   uint32_t ring_base = ring_address;
   uint32_t control_base = control_address;
   uint32_t tile_base = tile_address0;
   asm volatile("" : "+r"(ring_base), "+r"(control_base), "+r"(tile_base));
   uint32_t rotation = cwarp;  // piece index of this warp in the current stage, in [0, K1_CONSUMER_WARPS)
   uint32_t stage = 0;         // ring position and phase of the next visit
   uint32_t parity = 0;
   uint32_t visit = 0;         // MODE 4 only
   // MODE 4 (profiling): cycles this warp spent waiting for data, reported through the counts array
   long long probe_begin = 0;
   long long probe_waited = 0;
   long long probe_first = 0;
   if (MODE == 4) {
      probe_begin = clock64();
   }
   for (;;) {
      const uint32_t stage_address = ring_base + stage * static_cast<uint32_t>(sizeof(K1Stage));
      const uint32_t my_control = control_base + stage * static_cast<uint32_t>(sizeof(K1Control));
      const long long wait_begin = MODE == 4 ? clock64() : 0;
      mbarWaitAt(my_control + K1_CTRL_FULL, parity);
      if (MODE == 4) {
         const long long now = clock64();
         if (visit == 0) {
            probe_first = now - probe_begin;
         } else {
            probe_waited += now - wait_begin;
         }
         ++visit;
      }
      const uint4 meta = lds128(my_control);
      const uint32_t desc_count = meta.x;
      if (desc_count == K1_STOP) {
         if (MODE == 4 && lane == 0) {
            const uint32_t row = 15 * genome_length;  // a symbol row the finalize kernel never writes
            atomicAdd(&counts[row + 0], static_cast<uint32_t>((clock64() - probe_begin) >> 6));  // total
            atomicAdd(&counts[row + 1], static_cast<uint32_t>(probe_waited >> 6));               // waiting for data, after the first stage
            atomicAdd(&counts[row + 2], static_cast<uint32_t>(probe_first >> 6));                // until the first stage landed
            atomicAdd(&counts[row + 3], visit - 1);
            atomicAdd(&counts[row + 4], 1u);
            if (cwarp == 0) {  // per CTA: wall clock at which its consumers saw the end of the work
               uint32_t now;
               asm volatile("mov.u32 %0, %%globaltimer_lo;" : "=r"(now));
               counts[row + 4096 + blockIdx.x] = now;
            }
         }
         break;
      }
      // next ring position (no division: the stage count is not a power of two)
      const bool wrap = stage == K1_STAGES - 1;
      stage = wrap ? 0u : stage + 1u;
      parity ^= wrap ? 1u : 0u;
      const uint32_t slot_offset = (meta.z & K1_TILE_SLOT) != 0 ? TILE_BUFFER_BYTES : 0u;
      // the rotation continues where the previous stage stopped, so that a stage with fewer than 16
      // pieces does not always leave the same warps idle
      const uint32_t index = rotation;  // in [0, K1_CONSUMER_WARPS)
      rotation = rotation >= desc_count ? rotation - desc_count : rotation + K1_CONSUMER_WARPS - desc_count;
      // Lane 0's barrier arrives and the RED are predicated instructions (mbarArriveLane0,
      // redAddLane0), not branches: a warp without a piece in this stage runs the same tail with a
      // count of zero.
      uint4 desc = make_uint4(0u, 0u, 0u, 0u);
      uint32_t count = 0;
      if (index < desc_count && MODE != 1) {
         desc = lds128(stage_address + index * 16);  // {position, offset4, packed, aux}: the block starts with its descriptors
         const uint32_t payload_address = stage_address + ((desc.y - meta.y) << 2);
         const uint32_t kind = (desc.z >> 26) & 7u;
         if (kind == KIND_WORDRANGE) {  // rare; reads the stage while it works
            count = wordRangeCount(desc.w, payload_address, tile_base + slot_offset, lane);
            __syncwarp();
            mbarArriveLane0(my_control + K1_CTRL_EMPTY, lane);
         } else {
            // (reads past a short piece stay inside the stage buffer; those lanes are ignored)
            const uint4 first = lds128(payload_address + lane16);
            const uint4 second = lds128(payload_address + 512 + lane16);
            __syncwarp();
            mbarArriveLane0(my_control + K1_CTRL_EMPTY, lane);  // the stage can be refilled while the lookups run
            if (MODE == 3) {  // profiling: touch the payload only
               const uint32_t local = first.x ^ first.y ^ first.z ^ first.w ^ second.x ^ second.y ^ second.z ^ second.w;
               count = warpSum(local) == 0x12345678u ? 1u : 0u;
            } else {
               count = pieceFromRegisters(desc, kind, first, second, tile_base + slot_offset, k, lane, lane16);
            }
         }
      } else {
         __syncwarp();
         mbarArriveLane0(my_control + K1_CTRL_EMPTY, lane);
      }
      if (MODE != 2) {
         redAddLane0(&counts[((desc.z >> 16) & 0x1Fu) * genome_length + desc.x], count, lane);
      } else if (count == 0xFFFFFFFFu) {
         counts[0] = 1;
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

   // Filter bits are confined to the row layout, so a set bit is always an existing row. The warp's
   // eight filter words are fetched at once, and the (start, end) loads of K6_UNROLL words are in
   // flight together: the kernel is bound by global-memory latency, not bandwidth.
   const uint32_t warp_first = (slice * (K6_THREADS / 32) + warp) * K6_ROWS_PER_WARP;
   const uint32_t my_word = lane < K6_ROWS_PER_WARP / 32 ? tile32[(warp_first >> 5) + lane] : 0u;
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
struct HitRequest {
   silo_mutation_hit* hits = nullptr;
   uint32_t capacity = 0;
   unsigned long long* filter_scalars = nullptr;  // {cardinality, -, error flag (u32), -} of a filter evaluated inside
                                                  // the call: reported in the header and zeroed for the next query
   uint64_t valid_mask = 0;  // SymbolType::VALID_MUTATION_SYMBOLS
   double min_proportion = 0;
};

// Grid: diffPadded(genome_length) / 256 blocks (the difference array has genome_length + 1 entries).
// The kernel leaves the difference array, its block totals and the work-list state all-zero for the
// next query: every thread clears the element it read, the last block to finish clears the totals.
__global__ void __launch_bounds__(FIN_THREADS) finalizeCountsKernel(
   DevColumn column,
   uint32_t* __restrict__ diff_scratch,
   uint32_t* __restrict__ counts,
   uint32_t* __restrict__ work_state,
   HitRequest request
) {
   __shared__ uint32_t warp_totals[FIN_THREADS / 32];
   __shared__ uint32_t block_offset;
   const uint32_t genome_length = column.genome_length;
   uint32_t* diff = diff_scratch;
   uint32_t* block_totals = diff_scratch + diffPadded(genome_length);
   const uint32_t block_first = blockIdx.x * FIN_THREADS;
   const uint32_t lane = threadIdx.x & 31;
   const uint32_t warp = threadIdx.x >> 5;
   const uint32_t p = block_first + threadIdx.x;
   // issue every global load up front: the kernel is one latency chain otherwise
   const uint32_t mine = diff[p];  // (the array is padded to whole blocks)
   const uint32_t reference_symbol = p < genome_length ? column.local_reference[p] : 0u;
   const bool output_pass = request.hits != nullptr && p < genome_length;
   const uint32_t genome_symbol = output_pass ? column.global_reference[p] : 0u;
   uint32_t others = 0;
   uint32_t valid_others = 0;  // the same sum over the valid mutation symbols only
   uint32_t candidates = 0;    // OR of the counts that could be emitted (valid, not the reference genome's symbol)
   if (p < genome_length) {
      for (uint32_t symbol = 0; symbol < column.n_symbols; ++symbol) {
         const uint32_t value = symbol != reference_symbol ? counts[symbol * genome_length + p] : 0u;
         others += value;
         const uint32_t valid_value = ((request.valid_mask >> symbol) & 1ULL) != 0 ? value : 0u;
         valid_others += valid_value;
         candidates |= symbol != genome_symbol ? valid_value : 0u;
      }
   }
   if (warp == 0) {
      uint32_t partial = 0;
      for (uint32_t block = lane; block < blockIdx.x; block += 32) {
         partial += block_totals[block];
      }
      partial = __reduce_add_sync(0xFFFFFFFFu, partial);
      if (lane == 0) {
         block_offset = partial;
      }
   }
   diff[p] = 0;
   // inclusive scan of this block's 256 elements
   uint32_t inclusive = mine;
   for (int offset = 1; offset < 32; offset <<= 1) {
      const uint32_t other = __shfl_up_sync(0xFFFFFFFFu, inclusive, offset);
      if (lane >= static_cast<uint32_t>(offset)) {
         inclusive += other;
      }
   }
   if (lane == 31) {
      warp_totals[warp] = inclusive;
   }
   __syncthreads();
   uint32_t covered = block_offset + inclusive;
   for (uint32_t w = 0; w < warp; ++w) {
      covered += warp_totals[w];
   }
   const uint32_t reference_count = covered - others;
   if (p < genome_length) {
      counts[reference_symbol * genome_length + p] = reference_count;
   }
   if (output_pass) {
      const bool reference_is_valid = ((request.valid_mask >> reference_symbol) & 1ULL) != 0;
      const uint32_t total = valid_others + (reference_is_valid ? reference_count : 0u);
      if (reference_is_valid && reference_symbol != genome_symbol) {
         candidates |= reference_count;
      }
      // (`count > threshold_count` cannot hold for a zero count)
      if (total != 0 && candidates != 0) {
         // ceil(double(total) * min_proportion) - 1, the reference's operations in the reference's order
         const uint32_t threshold_count =
            request.min_proportion == 0 ? 0u : static_cast<uint32_t>(ceil(__dmul_rn(static_cast<double>(total), request.min_proportion)) - 1.0);
         for (uint32_t symbol = 0; symbol < column.n_symbols; ++symbol) {
            if (((request.valid_mask >> symbol) & 1ULL) == 0 || symbol == genome_symbol) {
               continue;
            }
            const uint32_t count = symbol == reference_symbol ? reference_count : counts[symbol * genome_length + p];
            if (count > threshold_count) {
               const uint32_t index = atomicAdd(&work_state[3], 1u);
               if (index < request.capacity) {
                  request.hits[1 + index] = silo_mutation_hit{p, symbol, count, total};
               }
            }
         }
      }
   }
   // This block has consumed its block totals (block_offset went into `covered`) and appended its tuples:
   // the last block to get here clears the totals and the work-list state and writes the header.
   __syncthreads();
   if (threadIdx.x == 0) {
      __threadfence();
      const uint32_t finished = atomicAdd(&work_state[2], 1u);
      if (finished == gridDim.x - 1) {
         __threadfence();
         for (uint32_t block = 0; block < gridDim.x; ++block) {
            block_totals[block] = 0;
         }
         if (request.hits != nullptr) {
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
         }
         work_state[0] = 0;  // the work list and its claim counter are empty between queries
         work_state[1] = 0;
         work_state[2] = 0;
         work_state[3] = 0;
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

void enqueueMutationCounts(
   silo_gpu_table* table,
   int column_index,
   const silo_gpu_filter* filter,
   uint32_t* d_counts,
   cudaStream_t stream,
   const HitRequest* request = nullptr,
   bool timed = false,    // record the per-call CUDA events that silo_gpu_get_stats reads (measurement only)
   bool prepared = false  // the filter interpreter already zeroed d_counts and built the work list (fused query)
) {
   require(table != nullptr, "mutation_counts: table is NULL");
   require(column_index >= 0 && static_cast<size_t>(column_index) < table->columns.size(), "mutation_counts: bad column index");
   require(filter == nullptr || filter->table == table, "mutation_counts: filter belongs to another table");
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
   coverageDiffKernel<<<n_chunks * K6_SLICES, K6_THREADS, 0, table->aux_stream>>>(column, words, popcounts, diff);
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
         containerCardinalityKernel<<<blocks, 256, 0, stream>>>(column, d_counts);
         SILO_CUDA_CHECK(cudaGetLastError());
         table->stats.kernel_launches++;
      }
   } else if (column.n_segments > 0) {
      static bool attribute_set = false;
      static int stream_only = 0;
      static uint32_t claim_batch = K1_BATCH_DEFAULT;
      static uint32_t tail_batches = 4;
      if (!attribute_set) {
         SILO_CUDA_CHECK(cudaFuncSetAttribute(containerAndCountKernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(K1Dynamic))));
         SILO_CUDA_CHECK(cudaFuncSetAttribute(containerAndCountKernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(K1Dynamic))));
         SILO_CUDA_CHECK(cudaFuncSetAttribute(containerAndCountKernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(K1Dynamic))));
         SILO_CUDA_CHECK(cudaFuncSetAttribute(containerAndCountKernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(K1Dynamic))));
         SILO_CUDA_CHECK(cudaFuncSetAttribute(containerAndCountKernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(K1Dynamic))));
         const char* flag = std::getenv("SILO_K1_STREAM_ONLY");
         stream_only = flag != nullptr ? flag[0] - '0' : 0;
         const char* batch_flag = std::getenv("SILO_K1_BATCH");
         if (batch_flag != nullptr) {
            claim_batch = static_cast<uint32_t>(std::min(32, std::max(1, std::atoi(batch_flag))));
         }
         const char* tail_flag = std::getenv("SILO_K1_TAIL");
         if (tail_flag != nullptr) {
            tail_batches = static_cast<uint32_t>(std::max(0, std::atoi(tail_flag)));
         }
         attribute_set = true;
      }
      const int blocks = static_cast<int>(std::min<uint32_t>(column.n_segments, static_cast<uint32_t>(table->ctx->sm_count)));
#define SILO_LAUNCH_K1(MODE)                                                                         \
   containerAndCountKernel<MODE><<<blocks, K1_THREADS, sizeof(K1Dynamic), stream>>>(                 \
      column, words, table->d_work_state, table->d_work_items, d_counts, claim_batch, tail_batches   \
   )
      switch (stream_only) {
         case 1: SILO_LAUNCH_K1(1); break;
         case 2: SILO_LAUNCH_K1(2); break;
         case 3: SILO_LAUNCH_K1(3); break;
         case 4: SILO_LAUNCH_K1(4); break;
         default: SILO_LAUNCH_K1(0); break;
      }
#undef SILO_LAUNCH_K1
      SILO_CUDA_CHECK(cudaGetLastError());
      table->stats.kernel_launches++;
   }
   recordTiming(ev_k1_end);
   SILO_CUDA_CHECK(cudaStreamWaitEvent(stream, table->ev_join, 0));
   finalizeCountsKernel<<<diffPadded(column.genome_length) / FIN_THREADS, FIN_THREADS, 0, stream>>>(
      column, diff, d_counts, table->d_work_state, request != nullptr ? *request : HitRequest{}
   );
   SILO_CUDA_CHECK(cudaGetLastError());
   table->stats.kernel_launches += 2;
   recordTiming(ev_end);
}

}  // namespace

void enqueuePreparedCountsLocked(silo_gpu_table* table, int column, const silo_gpu_filter* filter, uint32_t* d_counts, cudaStream_t stream) {
   enqueueMutationCounts(table, column, filter, d_counts, stream, nullptr, true, true);
}

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
      static double sums[3] = {0, 0, 0};
      static uint64_t calls = 0;
      const auto now = std::chrono::steady_clock::now();
      sums[phase] += std::chrono::duration<double, std::micro>(now - begin).count();
      begin = now;
      if (phase == 2 && ++calls % 64 == 0) {
         std::fprintf(stderr, "[silo query trace] stage %.1f us, enqueue %.1f us, wait %.1f us (mean of 64)\n", sums[0] / 64, sums[1] / 64, sums[2] / 64);
         sums[0] = sums[1] = sums[2] = 0;
      }
   }
};

// SILO_QUERY_GRAPHS=0 keeps the fused query calls on plain launches (debugging aid)
static bool queryGraphsEnabled() {
   static const bool enabled = [] {
      const char* flag = std::getenv("SILO_QUERY_GRAPHS");
      return flag == nullptr || flag[0] != '0';
   }();
   return enabled;
}

// page-locked tuple buffer for the worst case: every valid symbol but the reference genome's at every position
static void ensureHitsCapacity(silo_gpu_table* table, const HostColumn& host, uint64_t valid_symbol_mask, cudaStream_t stream) {
   const uint64_t needed = static_cast<uint64_t>(__builtin_popcountll(valid_symbol_mask)) * host.dev.genome_length;
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

static void sortHits(silo_mutation_hit* first, uint64_t count) {
   // the kernel appends in whatever order its threads get there: (position, symbol id) order
   std::sort(first, first + count, [](const silo_mutation_hit& a, const silo_mutation_hit& b) {
      return a.position != b.position ? a.position < b.position : a.symbol < b.symbol;
   });
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
      sortHits(table->h_hits_pinned + 1, count);
      *hits = table->h_hits_pinned + 1;
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
      QueryTrace trace;
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
         if (own_program && queryGraphsEnabled()) {
            // the shape of the query: every kernel parameter and copy size that enqueueAll bakes into nodes
            std::string key(reinterpret_cast<const char*>(staged.params), sizeof(staged.params));
            const uint64_t scalars[] = {staged.staged_bytes, staged.shared_bytes, static_cast<uint64_t>(column), valid_symbol_mask,
                                        reinterpret_cast<uint64_t>(table->h_hits_pinned), table->hits_capacity,
                                        reinterpret_cast<uint64_t>(host.dev.containers), host.dev.n_segments,
                                        reinterpret_cast<uint64_t>(host.dev.global_reference), reinterpret_cast<uint64_t>(table->d_counts),
                                        reinterpret_cast<uint64_t>(table->d_work_items), reinterpret_cast<uint64_t>(table->d_coverage_diff)};
            key.append(reinterpret_cast<const char*>(scalars), sizeof(scalars));
            key.append(reinterpret_cast<const char*>(&min_proportion), sizeof(min_proportion));
            for (const silo_gpu_table::CachedGraph& cached : table->query_graphs) {
               if (cached.key == key) {
                  replay = cached.exec;
                  break;
               }
            }
            if (replay == nullptr && key == table->last_query_key) {  // second time in a row: worth a graph
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
      sortHits(first, count);
      *hits = first;
      *n_hits = count;
      if (cardinality != nullptr && program != nullptr) {
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
