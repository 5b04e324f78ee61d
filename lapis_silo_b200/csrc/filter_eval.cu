// S2: filter-program evaluation on device.
//
// Replaces Operator::evaluate() of the compiled filter tree
// (/root/reference/src/rhydb/query_engine/filter/operators/*.cpp, reached from
// operators/compute_filter.cpp:14-21). The reference folds roaring containers operator by operator
// on one thread; here ONE CTA per 2^16-row chunk interprets the whole flat program over dense
// 8 KiB tiles held in shared memory, so intermediate results never touch HBM:
//   IndexScan views (symbol_in_set.cpp:216-228)         -> orContainerIntoTile (array/bitset/run decode)
//   IsInCoveredRegion::makeBitmap (is_in_covered_region.cpp:53-62,
//        horizontal_coverage_index.h:57-98)              -> coveredTile (start/end scan + N runs)
//   RangeSelection::evaluate (range_selection.cpp:54-87) -> rangesTile
//   Intersection / Union / Complement (intersection.cpp:59-105, union.cpp:34-42,
//        complement.cpp:51-56 + row_layout.cpp:18-23)    -> word-wise AND / ANDNOT / OR / XOR-with-layout
//   Threshold::evaluate (threshold.cpp:64-138)           -> per-row u16 counters in shared memory
//        (the DP over k roaring bitmaps computes exactly "at least / exactly k children match")
#include <algorithm>
#include <cstring>
#include <memory>

#include "common.cuh"
#include "eval_device.cuh"

namespace silo {

namespace {

constexpr int STACK_DEPTH = 6;

struct DevBitmap {
   const DevContainer* containers;  // `position` holds the GLOBAL chunk key, ascending
   const uint8_t* payload;          // base the descriptors' offsets refer to
   uint32_t n_containers;
   uint32_t pad;
};

struct EvalParams {
   const silo_filter_instr* instrs;
   uint32_t n_instrs;
   uint32_t n_chunks;
   const DevColumn* columns;
   const uint8_t* blob;
   const DevBitmap* bitmaps;
   const uint32_t* chunk_sizes;
   const DevValueColumn* value_columns;  // SILO_OP_PUSH_COMPARE
   const uint32_t* chunk_row_begin;
   uint32_t first_chunk;
   uint32_t stack_depth;    // tiles the program needs at most
   uint32_t has_threshold;  // whether the counter tile is needed
   uint32_t pad;
   uint64_t* out_words;
   uint32_t* out_popcount;
   unsigned long long* out_cardinality;
   uint32_t* error_flag;
   // Fused query only (nullptr otherwise): the interpreter also does what prepareQueryKernel does for the counts
   // kernels of ONE column -- zero the counts, and CTA c appends the segment records of chunk c to the container
   // kernel's work list when the chunk holds a filtered row -- so the query has one launch less.
   const DevSegment* prepare_segments;
   const uint32_t* prepare_chunk_seg_begin;
   uint32_t* prepare_work_state;
   DevSegment* prepare_work_items;
   uint32_t* prepare_counts;
   uint32_t prepare_counts_words;
   // Threshold sweep (thresholdSweepKernel, launched in front of the interpreter by launchProgram): the THR_PROFILE at
   // instruction sweep_pc finds its per-row counts in sweep_counters instead of walking the containers itself.
   uint32_t sweep_pc;               // NO_SWEEP: none
   const uint32_t* sweep_counters;  // [n_chunks][32768] packed u16 pairs, every lane holds count + flushes * sweep_bias
   const uint32_t* sweep_flushes;   // [n_chunks]
   uint32_t sweep_bias;
   int32_t sweep_column;            // (host side of the launch)
   uint64_t sweep_table_offset;     // ... the instruction's table inside the blob
};
constexpr uint32_t NO_SWEEP = 0xFFFFFFFFu;

// Dynamic shared memory of the interpreter: [small | stack_depth tiles | 128 KiB of counters if the
// program holds a Threshold]. A boolean-only program needs 8 KiB per stack level, so several CTAs
// share an SM; only Threshold programs take the whole SM.
constexpr uint32_t RANGE_HITS = 62;
constexpr uint32_t CACHED_INSTRS = 64;
struct EvalSmall {
   uint32_t range[2];
   uint32_t reduce[EVAL_WARPS];
   uint32_t n_hits;
   uint32_t pad;
   uint2 hits[RANGE_HITS];                  // PUSH_RANGES: the ranges that overlap this chunk
   silo_filter_instr instrs[CACHED_INSTRS];  // the head of the program, fetched once per CTA
};
constexpr size_t COUNTER_BYTES = 65536 * sizeof(uint16_t);

struct EvalShared {
   EvalSmall* small;
   uint64_t (*stack)[TILE_WORDS];
   uint32_t* counters32;  // 65536 x u16 per-row match counters (Threshold)
   uint32_t* range;
   uint32_t* reduce;
};

__host__ __device__ inline size_t evalSharedBytes(uint32_t stack_depth, bool has_threshold) {
   return sizeof(EvalSmall) + static_cast<size_t>(stack_depth) * TILE_BYTES + (has_threshold ? COUNTER_BYTES : 0);
}

__global__ void __launch_bounds__(EVAL_THREADS, 2) evalProgramKernel(EvalParams p) {
   extern __shared__ __align__(128) uint8_t smem_raw[];
   EvalShared sh;
   sh.small = reinterpret_cast<EvalSmall*>(smem_raw);
   sh.stack = reinterpret_cast<uint64_t(*)[TILE_WORDS]>(smem_raw + sizeof(EvalSmall));
   sh.counters32 = reinterpret_cast<uint32_t*>(smem_raw + sizeof(EvalSmall) + static_cast<size_t>(p.stack_depth) * TILE_BYTES);
   sh.range = sh.small->range;
   sh.reduce = sh.small->reduce;
   const uint32_t chunk = blockIdx.x;
   const uint32_t tid = threadIdx.x;
   const uint32_t lane = tid & 31;
   const uint32_t warp = tid >> 5;
   const uint32_t chunk_size = p.chunk_sizes[chunk];
   const uint64_t layout_word = layoutWord(chunk_size, tid);
   const uint32_t chunk_base = (p.first_chunk + chunk) << 16;

   int sp = 0;  // number of tiles on the stack
   uint32_t thr_target = 0;
   bool thr_exact = false;
   // Rows outside the row layout that a threshold's child tiles hold (this thread's word of them). The reference's
   // Threshold (threshold.cpp:64-138) complements negated children inside the layout only (row_layout.cpp:18-23), so
   // such an id takes part in the count exactly when some child holds it (threshold.test.cpp:249-311).
   uint64_t thr_outside = 0;

   if (tid < min(p.n_instrs, CACHED_INSTRS) * 4) {  // 16-byte instructions as 4 words each
      reinterpret_cast<uint32_t*>(sh.small->instrs)[tid] = reinterpret_cast<const uint32_t*>(p.instrs)[tid];
   }
   // the chunk's segment records for the work list of the counts kernels (static column data): fetched now, so that
   // the epilogue is one atomic and one store instead of a chain of three dependent global loads
   uint32_t first_segment = 0;
   uint32_t chunk_segments = 0;
   uint4 my_segment = make_uint4(0u, 0u, 0u, 0u);
   if (p.prepare_work_items != nullptr) {  // (kernel parameter: uniform)
      first_segment = p.prepare_chunk_seg_begin[chunk];
      chunk_segments = p.prepare_chunk_seg_begin[chunk + 1] - first_segment;
      if (tid < chunk_segments) {
         my_segment = reinterpret_cast<const uint4*>(p.prepare_segments + first_segment)[tid];
      }
   }
   // (the program was staged by a copy in front of the launch, or long ago: it is there. What follows reads and writes
   // what the previous query's kernels still use: tiles, counts, work list, scalars.)
   gridDependencyWait();
   gridDependencyLaunch();
   __syncthreads();

   for (uint32_t pc = 0; pc < p.n_instrs; ++pc) {
      const silo_filter_instr ins = pc < CACHED_INSTRS ? sh.small->instrs[pc] : p.instrs[pc];
      switch (ins.opcode) {
         case SILO_OP_PUSH_EMPTY:
            sh.stack[sp++][tid] = 0;
            break;
         case SILO_OP_PUSH_FULL:
            sh.stack[sp++][tid] = layout_word;
            break;
         case SILO_OP_PUSH_SYMBOLS: {
            const DevColumn& column = p.columns[ins.column];
            uint64_t* tile = sh.stack[sp++];
            tile[tid] = 0;
            findPositionRange(sh.range, column, chunk, ins.a);  // includes a __syncthreads
            const uint32_t lo = sh.range[0];
            const uint32_t hi = sh.range[1];
            for (uint32_t i = lo; i < hi; ++i) {
               const DevContainer desc = column.containers[i];
               if (((ins.b >> desc.symbol()) & 1ULL) != 0) {
                  orContainerIntoTile(tile, column.payload, desc);
               }
               __syncthreads();  // the next container may mix plain and atomic updates of the same word
            }
            break;
         }
         case SILO_OP_PUSH_COVERED: {
            const DevColumn& column = p.columns[ins.column];
            uint32_t* tile32 = reinterpret_cast<uint32_t*>(sh.stack[sp++]);
            const uint2* rows = column.start_end + column.chunk_row_begin[chunk];
            const uint32_t position = ins.a;
            constexpr uint32_t UNROLL = 4;  // (four loads in flight per thread instead of a chain of 64 memory latencies)
            for (uint32_t base = 0; base < 65536; base += EVAL_THREADS * UNROLL) {
               uint2 ranges[UNROLL];
#pragma unroll
               for (uint32_t u = 0; u < UNROLL; ++u) {
                  const uint32_t row = base + u * EVAL_THREADS + tid;
                  ranges[u] = row < chunk_size ? rows[row] : make_uint2(0u, 0u);
               }
#pragma unroll
               for (uint32_t u = 0; u < UNROLL; ++u) {
                  const uint32_t row = base + u * EVAL_THREADS + tid;
                  const bool covered = ranges[u].x <= position && position < ranges[u].y;
                  const uint32_t bits = __ballot_sync(0xFFFFFFFFu, covered);
                  if (lane == 0) {
                     tile32[row >> 5] = bits;
                  }
               }
            }
            __syncthreads();
            const uint32_t missing_begin = column.chunk_missing_begin[chunk];
            const uint32_t missing_end = column.chunk_missing_begin[chunk + 1];
            for (uint32_t i = missing_begin + tid; i < missing_end; i += EVAL_THREADS) {
               if (rowMissingAt(column, i, position)) {
                  const uint32_t row = column.missing_row[i];
                  atomicAnd(&tile32[row >> 5], ~(1u << (row & 31)));
               }
            }
            __syncthreads();
            if ((ins.flags & 1) != 0) {
               sh.stack[sp - 1][tid] ^= layout_word;
            }
            break;
         }
         case SILO_OP_PUSH_NULLS: {
            const DevColumn& column = p.columns[ins.column];
            sh.stack[sp++][tid] =
               column.null_words != nullptr ? column.null_words[static_cast<size_t>(chunk) * TILE_WORDS + tid] : 0ULL;
            break;
         }
         case SILO_OP_PUSH_BITMAP: {
            const DevBitmap bitmap = p.bitmaps[ins.a];
            uint64_t* tile = sh.stack[sp++];
            tile[tid] = 0;
            if (tid < 32) {
               const uint32_t key = p.first_chunk + chunk;
               const uint32_t lo = warpLowerBound(0, bitmap.n_containers, key, tid, [&](uint32_t index) {
                  return bitmap.containers[index].position;
               });
               if (tid == 0) {
                  sh.range[0] = lo;
                  sh.range[1] = (lo < bitmap.n_containers && bitmap.containers[lo].position == key) ? 1u : 0u;
               }
            }
            __syncthreads();
            if (sh.range[1] != 0) {
               const DevContainer desc = bitmap.containers[sh.range[0]];
               orContainerIntoTile(tile, bitmap.payload, desc);
               __syncthreads();
               if ((tile[tid] & ~layout_word) != 0) {
                  atomicOr(p.error_flag, 1u);  // ids outside the row layout
               }
            }
            break;
         }
         case SILO_OP_PUSH_COMPARE: {
            // A Selection predicate over a value column for every row of the chunk (4 bytes per row, coalesced): what
            // Predicate::makeBitmap computes on the host (selection.h:33-43), or match() row by row on the candidates
            // of the child (selection.cpp:105-118) -- both are "child AND predicate".
            const DevValueColumn column = p.value_columns[ins.column];
            const uint32_t* values = column.values + p.chunk_row_begin[chunk];
            uint32_t* tile32 = reinterpret_cast<uint32_t*>(sh.stack[sp++]);
            const uint32_t comparator = ins.flags & 7u;
            const bool is_signed = (ins.flags & SILO_CMP_SIGNED) != 0;
            const uint32_t bias = is_signed ? 0x80000000u : 0u;  // signed order as unsigned order
            const uint32_t wanted = ins.a ^ bias;
            const uint32_t upper = static_cast<uint32_t>(ins.b) ^ bias;
            const uint32_t* set = reinterpret_cast<const uint32_t*>(p.blob + ins.b);
            // (four loads in flight per thread: as one load per iteration the scan was a chain of 64 memory latencies)
            constexpr uint32_t UNROLL = 4;
            for (uint32_t base = 0; base < 65536; base += EVAL_THREADS * UNROLL) {
               uint32_t raws[UNROLL];
#pragma unroll
               for (uint32_t u = 0; u < UNROLL; ++u) {
                  const uint32_t row = base + u * EVAL_THREADS + tid;
                  raws[u] = row < chunk_size ? values[row] : 0u;
               }
#pragma unroll
               for (uint32_t u = 0; u < UNROLL; ++u) {
                  const uint32_t row = base + u * EVAL_THREADS + tid;
                  const uint32_t raw = raws[u];
                  const uint32_t value = raw ^ bias;
                  bool match = false;
                  switch (comparator) {
                     case SILO_CMP_EQUALS: match = value == wanted; break;
                     case SILO_CMP_NOT_EQUALS: match = value != wanted; break;
                     case SILO_CMP_LESS: match = value < wanted; break;
                     case SILO_CMP_LESS_OR_EQUALS: match = value <= wanted; break;
                     case SILO_CMP_HIGHER: match = value > wanted; break;
                     case SILO_CMP_HIGHER_OR_EQUALS: match = value >= wanted; break;
                     case SILO_CMP_BETWEEN: match = value >= wanted && value <= upper; break;
                     default: {  // SILO_CMP_IN_SET
                        const uint32_t at = lowerBound(set, ins.a, raw);
                        match = at < ins.a && set[at] == raw;
                     }
                  }
                  const uint32_t bits = __ballot_sync(0xFFFFFFFFu, match && row < chunk_size);
                  if (lane == 0) {
                     tile32[row >> 5] = bits;
                  }
               }
            }
            __syncthreads();
            if (column.null_words != nullptr) {  // CompareToValueSelection::match: a null row gives with_nulls
               const uint64_t nulls = column.null_words[static_cast<size_t>(chunk) * TILE_WORDS + tid];
               uint64_t word = sh.stack[sp - 1][tid];
               word = (ins.flags & SILO_CMP_WITH_NULLS) != 0 ? (word | nulls) : (word & ~nulls);
               sh.stack[sp - 1][tid] = word & layout_word;
            }
            break;
         }
         case SILO_OP_PUSH_RANGES: {
            // phase 1: every thread tests ranges tid, tid + 1024, ... against the chunk and appends the
            // overlapping ones (usually zero to two: RangeSelection holds at most one per chunk) to a
            // shared list; phase 2: every thread ORs the listed ranges into its own word.
            const uint2* ranges = reinterpret_cast<const uint2*>(p.blob + ins.b);
            if (tid == 0) {
               sh.small->n_hits = 0;
            }
            __syncthreads();
            const uint64_t chunk_end = static_cast<uint64_t>(chunk_base) + 65536;
            uint32_t spilled = 0;
            for (uint32_t i = tid; i < ins.a; i += EVAL_THREADS) {
               const uint2 range = ranges[i];
               if (range.x < chunk_end && range.y > chunk_base && range.x < range.y) {
                  const uint32_t slot = atomicAdd(&sh.small->n_hits, 1u);
                  if (slot < RANGE_HITS) {
                     sh.small->hits[slot] = range;
                  } else {
                     spilled = 1;
                  }
               }
            }
            const bool any_spilled = __syncthreads_or(static_cast<int>(spilled)) != 0;
            const uint32_t word_first = chunk_base + tid * 64;
            auto orRange = [&](uint64_t word, uint2 range) {
               const uint32_t start = max(range.x, word_first);
               const uint64_t end = min(static_cast<uint64_t>(range.y), static_cast<uint64_t>(word_first) + 64);
               if (start < end) {
                  const uint32_t first_bit = start - word_first;
                  const uint32_t last_bit = static_cast<uint32_t>(end - 1 - word_first);
                  word |= (~0ULL << first_bit) & (~0ULL >> (63 - last_bit));
               }
               return word;
            };
            uint64_t word = 0;
            if (any_spilled) {  // more overlapping ranges than the list holds: the plain loop
               for (uint32_t i = 0; i < ins.a; ++i) {
                  word = orRange(word, ranges[i]);
               }
            } else {
               const uint32_t n_hits = sh.small->n_hits;
               for (uint32_t i = 0; i < n_hits; ++i) {
                  word = orRange(word, sh.small->hits[i]);
               }
            }
            sh.stack[sp++][tid] = word & layout_word;
            break;
         }
         case SILO_OP_AND:
            sh.stack[sp - 2][tid] &= sh.stack[sp - 1][tid];
            --sp;
            break;
         case SILO_OP_ANDNOT:
            sh.stack[sp - 2][tid] &= ~sh.stack[sp - 1][tid];
            --sp;
            break;
         case SILO_OP_OR:
            sh.stack[sp - 2][tid] |= sh.stack[sp - 1][tid];
            --sp;
            break;
         case SILO_OP_NOT:
            sh.stack[sp - 1][tid] ^= layout_word;
            break;
         case SILO_OP_THR_BEGIN: {
            const uint32_t bias = static_cast<uint32_t>(ins.b & 0xFFFFu);
            thr_target = ins.a + bias;
            thr_exact = (ins.flags & 1) != 0;
            thr_outside = 0;
            for (uint32_t i = tid; i < 32768; i += EVAL_THREADS) {
               sh.counters32[i] = bias | (bias << 16);
            }
            break;
         }
         case SILO_OP_THR_ADD: {
            uint64_t word = sh.stack[--sp][tid];
            thr_outside |= word & ~layout_word;
            if ((ins.flags & 1) != 0) {
               word = ~word;  // (rows outside the layout that no child holds are dropped at THR_END)
            }
            uint16_t* counters16 = reinterpret_cast<uint16_t*>(sh.counters32);
            while (word != 0) {
               const uint32_t bit = static_cast<uint32_t>(__ffsll(static_cast<long long>(word)) - 1);
               counters16[tid * 64 + bit] += 1;  // rows 64*tid.. are owned by this thread in this op
               word &= word - 1;
            }
            break;
         }
         case SILO_OP_THR_ADD_SYMBOLS: {
            const DevColumn& column = p.columns[ins.column];
            findPositionRange(sh.range, column, chunk, ins.a);
            const uint32_t lo = sh.range[0];
            const uint32_t hi = sh.range[1];
            for (uint32_t i = lo + warp; i < hi; i += EVAL_WARPS) {
               const DevContainer desc = column.containers[i];
               if (((ins.b >> desc.symbol()) & 1ULL) != 0) {
                  addContainerToCounters(sh.counters32, column.payload, desc, (ins.flags & 1) != 0, lane);
               }
            }
            break;
         }
         case SILO_OP_THR_ADD_COVERED: {
            const DevColumn& column = p.columns[ins.column];
            const uint32_t* positions = reinterpret_cast<const uint32_t*>(p.blob + ins.b);
            const uint32_t n_positions = ins.a;
            const uint2* rows = column.start_end + column.chunk_row_begin[chunk];
            uint16_t* counters16 = reinterpret_cast<uint16_t*>(sh.counters32);
            for (uint32_t row = tid; row < chunk_size; row += EVAL_THREADS) {
               const uint2 range = rows[row];
               const uint32_t inside =
                  lowerBound(positions, n_positions, range.y) - lowerBound(positions, n_positions, range.x);
               counters16[row] += static_cast<uint16_t>(inside);
            }
            __syncthreads();
            const uint32_t missing_begin = column.chunk_missing_begin[chunk];
            const uint32_t missing_end = column.chunk_missing_begin[chunk + 1];
            for (uint32_t i = missing_begin + tid; i < missing_end; i += EVAL_THREADS) {
               uint32_t hidden = 0;
               for (uint64_t run = column.missing_offsets[i]; run < column.missing_offsets[i + 1]; ++run) {
                  const uint2 r = column.missing_runs[run];
                  hidden += lowerBound(positions, n_positions, r.y) - lowerBound(positions, n_positions, r.x);
               }
               counters16[column.missing_row[i]] -= static_cast<uint16_t>(hidden);
            }
            break;
         }
         case SILO_OP_THR_PROFILE: {
            if (pc == p.sweep_pc) {
               // the sweep kernel already counted: every lane of its array holds count + bias per CTA that flushed
               // into this chunk (no lane underflows: count + bias >= 0 within every flush)
               const uint32_t surplus = p.sweep_flushes[chunk] * p.sweep_bias;
               const uint32_t* counted = p.sweep_counters + static_cast<size_t>(chunk) * 32768;
               for (uint32_t i = tid; i < 32768; i += EVAL_THREADS) {
                  sh.counters32[i] += counted[i] - (surplus | (surplus << 16));
               }
               break;
            }
            const DevColumn& column = p.columns[ins.column];
            const uint2* table = reinterpret_cast<const uint2*>(p.blob + ins.b);  // {add_mask, sub_mask}
            const uint32_t lo = column.chunk_desc_begin[chunk];
            const uint32_t hi = column.chunk_desc_begin[chunk + 1];
            for (uint32_t i = lo + warp; i < hi; i += EVAL_WARPS) {
               const DevContainer desc = column.containers[i];
               const uint2 masks = table[desc.position];
               const uint32_t bit = 1u << desc.symbol();
               if ((masks.x & bit) != 0) {
                  addContainerToCounters(sh.counters32, column.payload, desc, false, lane);
               } else if ((masks.y & bit) != 0) {
                  addContainerToCounters(sh.counters32, column.payload, desc, true, lane);
               }
            }
            break;
         }
         case SILO_OP_THR_END: {
            const uint16_t* counters16 = reinterpret_cast<const uint16_t*>(sh.counters32);
            uint64_t word = 0;
            for (uint32_t bit = 0; bit < 64; ++bit) {
               // rotate the start so that the 32 lanes of a warp hit 32 different banks
               const uint32_t b = (bit + lane * 2) & 63;
               const uint32_t count = counters16[tid * 64 + b];
               const bool hit = thr_exact ? count == thr_target : count >= thr_target;
               word |= static_cast<uint64_t>(hit) << b;
            }
            sh.stack[sp++][tid] = word & (layout_word | thr_outside);
            break;
         }
         default:
            break;
      }
      __syncthreads();
   }

   const uint64_t result = sp > 0 ? sh.stack[sp - 1][tid] : 0ULL;
   p.out_words[static_cast<size_t>(chunk) * TILE_WORDS + tid] = result;
   uint32_t total = __reduce_add_sync(0xFFFFFFFFu, static_cast<uint32_t>(__popcll(result)));
   if (lane == 0) {
      sh.reduce[warp] = total;
   }
   if (p.prepare_counts != nullptr) {
      zeroCountWords(p.prepare_counts, p.prepare_counts_words, chunk * EVAL_THREADS + tid, gridDim.x * EVAL_THREADS);
   }
   __syncthreads();
   if (warp == 0) {
      total = __reduce_add_sync(0xFFFFFFFFu, sh.reduce[lane]);
      if (lane == 0) {
         p.out_popcount[chunk] = total;
         atomicAdd(p.out_cardinality, static_cast<unsigned long long>(total));
         if (p.prepare_work_items != nullptr) {
            const uint32_t n_segments = total != 0 ? chunk_segments : 0u;
            sh.range[0] = n_segments != 0 ? atomicAdd(&p.prepare_work_state[0], n_segments) : 0u;
            sh.range[1] = n_segments;
         }
      }
   }
   if (p.prepare_work_items != nullptr) {  // (kernel parameter: uniform)
      __syncthreads();
      const uint32_t n_segments = sh.range[1];
      uint4* target = reinterpret_cast<uint4*>(p.prepare_work_items + sh.range[0]);
      if (tid < n_segments) {
         target[tid] = my_segment;
      }
      const uint4* source = reinterpret_cast<const uint4*>(p.prepare_segments + first_segment);
      for (uint32_t i = tid + EVAL_THREADS; i < n_segments; i += EVAL_THREADS) {  // (a chunk with more than 1,024 segments)
         target[i] = source[i];
      }
   }
}

// ---------------------------------------------------------------------------------------------
// Threshold sweep: the per-row counts of ONE THR_PROFILE instruction for the whole column.
//
// Replaces the O(n k) whole-bitmap DP of Threshold::evaluate (threshold.cpp:64-138) for the children that are index
// scans over the vertical index -- for a MutationProfile (mutation_profile.cpp:198-257) that is every stored container
// of the column. The interpreter's own THR_PROFILE walks a chunk's containers inside the chunk's CTA: one CTA per
// chunk is 153 CTAs of 130 KiB shared memory on 148 SMs (two waves), and every container costs a chain of dependent
// global loads. Here the column's pieces are cut into ONE contiguous range per SM, equal bytes (pool.cu), each warp
// keeps the loads of its next piece in flight while it bumps the current one, and a CTA flushes its shared-memory
// counters into the chunk's global array when its range leaves the chunk.
// Algorithmic bytes: 16 B descriptor + payload of every container + 128 KiB of counters per chunk (written here,
// read by the interpreter).
// ---------------------------------------------------------------------------------------------
constexpr int SWEEP_THREADS = 1024;
constexpr int SWEEP_WARPS = SWEEP_THREADS / 32;

struct SweepPiece {
   DevContainer desc;
   uint4 first;
   uint4 second;
   uint32_t action;  // 0: skip, 1: add, 2: subtract
};

__device__ __forceinline__ SweepPiece loadSweepPiece(const DevColumn& column, const uint2* table, uint32_t index, uint32_t lane) {
   SweepPiece piece;
   piece.desc = column.containers[index];  // one 16-byte line for the warp
   const uint4* payload = reinterpret_cast<const uint4*>(column.payload + (static_cast<size_t>(piece.desc.offset4) << 2));
   const uint32_t kind = piece.desc.type();
   const uint32_t bytes = kind == KIND_ARRAY_T ? arrayPieceBytes(piece.desc.cardinality())
                          : kind == KIND_RUNS_W ? runsPieceBytes(piece.desc.aux)
                          : kind == KIND_BITSET ? 1024u
                                                : 0u;
   piece.first = lane * 16 < bytes ? payload[lane] : make_uint4(0u, 0u, 0u, 0u);
   piece.second = 512 + lane * 16 < bytes ? payload[32 + lane] : make_uint4(0u, 0u, 0u, 0u);
   const uint2 masks = table[piece.desc.position];
   const uint32_t bit = 1u << piece.desc.symbol();
   piece.action = (masks.x & bit) != 0 ? 1u : (masks.y & bit) != 0 ? 2u : 0u;
   return piece;
}

__global__ void __launch_bounds__(SWEEP_THREADS, 1) thresholdSweepKernel(
   DevColumn column,
   const uint2* __restrict__ table,       // [genome_length] {add_mask, sub_mask}
   const uint32_t* __restrict__ split,    // [gridDim.x + 1] piece indices
   uint32_t bias,                         // every counter lane starts at bias (>= the subtractions a row can see)
   uint32_t* __restrict__ out             // [n_chunks][32768], zeroed
) {
   extern __shared__ __align__(16) uint32_t sweep_counters[];  // 65536 x u16
   __shared__ uint32_t first_chunk_of_range;
   const uint32_t begin = split[blockIdx.x];
   const uint32_t end = split[blockIdx.x + 1];
   if (begin >= end) {
      return;
   }
   const uint32_t lane = threadIdx.x & 31;
   const uint32_t warp = threadIdx.x >> 5;
   if (threadIdx.x == 0) {  // the chunk that holds piece `begin`: last chunk whose first piece is <= begin
      uint32_t lo = 0;
      uint32_t hi = column.n_chunks;
      while (hi - lo > 1) {
         const uint32_t mid = (lo + hi) >> 1;
         if (column.chunk_desc_begin[mid] <= begin) {
            lo = mid;
         } else {
            hi = mid;
         }
      }
      first_chunk_of_range = lo;
   }
   __syncthreads();
   uint32_t chunk = first_chunk_of_range;
   uint32_t cursor = begin;
   while (cursor < end) {
      const uint32_t chunk_end = min(end, column.chunk_desc_begin[chunk + 1]);
      if (chunk_end > cursor) {
         for (uint32_t i = threadIdx.x; i < 32768; i += SWEEP_THREADS) {
            sweep_counters[i] = bias | (bias << 16);
         }
         __syncthreads();
         // warp w takes the pieces cursor + w, + 32, ...; the next piece's loads are in flight during the bumps
         uint32_t index = cursor + warp;
         SweepPiece next{};
         if (index < chunk_end) {
            next = loadSweepPiece(column, table, index, lane);
         }
         while (index < chunk_end) {
            const SweepPiece piece = next;
            const uint32_t following = index + SWEEP_WARPS;
            if (following < chunk_end) {
               next = loadSweepPiece(column, table, following, lane);
            }
            if (piece.action != 0) {
               bumpPieceFromRegisters(
                  sweep_counters, piece.desc, piece.first, piece.second, column.payload + (static_cast<size_t>(piece.desc.offset4) << 2),
                  piece.action == 2, lane
               );
            }
            index = following;
         }
         __syncthreads();
         uint32_t* target = out + static_cast<size_t>(chunk) * 32768;
         for (uint32_t i = threadIdx.x; i < 32768; i += SWEEP_THREADS) {
            const uint32_t value = sweep_counters[i];
            if (value != 0) {
               atomicAdd(&target[i], value);
            }
         }
         __syncthreads();
      }
      cursor = chunk_end;
      ++chunk;
   }
}

__global__ void popcountTilesKernel(
   const uint64_t* __restrict__ words,
   uint32_t* __restrict__ out_popcount,
   unsigned long long* __restrict__ out_cardinality
) {
   __shared__ uint32_t reduce[32];
   const uint32_t chunk = blockIdx.x;
   const uint32_t lane = threadIdx.x & 31;
   const uint32_t warp = threadIdx.x >> 5;
   uint32_t total = __reduce_add_sync(
      0xFFFFFFFFu, static_cast<uint32_t>(__popcll(words[static_cast<size_t>(chunk) * TILE_WORDS + threadIdx.x]))
   );
   if (lane == 0) {
      reduce[warp] = total;
   }
   __syncthreads();
   if (warp == 0) {
      total = __reduce_add_sync(0xFFFFFFFFu, reduce[lane]);
      if (lane == 0) {
         out_popcount[chunk] = total;
         atomicAdd(out_cardinality, static_cast<unsigned long long>(total));
      }
   }
}

// ---- host side ---------------------------------------------------------------------------------

// Portable Roaring format (RoaringFormatSpec; what roaring::Roaring::write emits,
// roaring_util/roaring_serialize.h:15-30) -> descriptors + a 16-byte aligned payload slab.
// Two passes over the same code: with descs_out == nullptr only the sizes are measured, so that the
// second pass can write straight into pinned staging memory (no intermediate copies).
struct RoaringSizes {
   uint32_t n_containers = 0;
   uint64_t payload_bytes = 0;  // multiple of 16
};

uint32_t readU16(const uint8_t* data) {
   uint16_t value;
   std::memcpy(&value, data, 2);
   return value;
}
uint32_t readU32(const uint8_t* data) {
   uint32_t value;
   std::memcpy(&value, data, 4);
   return value;
}

RoaringSizes parseRoaring(const uint8_t* data, uint64_t size, DevContainer* descs_out, uint8_t* payload_out) {
   constexpr uint32_t SERIAL_COOKIE_NO_RUNCONTAINER = 12346;
   constexpr uint32_t SERIAL_COOKIE = 12347;
   constexpr uint32_t NO_OFFSET_THRESHOLD = 4;
   auto bad = [](const char* what) { throw ApiError(SILO_E_BAD_PROGRAM, std::string("roaring bitmap: ") + what); };
   if (data == nullptr || size < 4) {
      bad("too short");
   }
   uint64_t pos = 0;
   const uint32_t cookie = readU32(data);
   pos += 4;
   uint32_t n = 0;
   const uint8_t* run_flags = nullptr;
   if ((cookie & 0xFFFF) == SERIAL_COOKIE) {
      n = (cookie >> 16) + 1;
      if (pos + (n + 7) / 8 > size) {
         bad("truncated run flags");
      }
      run_flags = data + pos;
      pos += (n + 7) / 8;
   } else if (cookie == SERIAL_COOKIE_NO_RUNCONTAINER) {
      if (pos + 4 > size) {
         bad("truncated header");
      }
      n = readU32(data + pos);
      pos += 4;
   } else {
      bad("bad cookie");
   }
   if (n > 65536 || pos + 4ULL * n > size) {
      bad("truncated key table");
   }
   const uint8_t* keys = data + pos;
   pos += 4ULL * n;
   if (run_flags == nullptr || n >= NO_OFFSET_THRESHOLD) {
      pos += 4ULL * n;
   }
   RoaringSizes sizes;
   sizes.n_containers = n;
   for (uint32_t i = 0; i < n; ++i) {
      const uint32_t key = readU16(keys + 4 * i);
      const uint32_t cardinality = readU16(keys + 4 * i + 2) + 1;
      if (i > 0 && readU16(keys + 4 * (i - 1)) >= key) {
         bad("keys not ascending");
      }
      const bool is_run = run_flags != nullptr && ((run_flags[i / 8] >> (i % 8)) & 1) != 0;
      DevContainer desc{};
      desc.position = key;
      uint64_t payload_bytes = 0;
      uint32_t type = 0;
      const uint8_t* src = nullptr;
      if (is_run) {
         if (pos + 2 > size) {
            bad("truncated run container");
         }
         desc.aux = readU16(data + pos);
         payload_bytes = 4ULL * desc.aux;
         src = data + pos + 2;
         pos += 2 + payload_bytes;
         type = KIND_RAW_RUN;
      } else if (cardinality <= 4096) {
         payload_bytes = 2ULL * cardinality;
         src = data + pos;
         pos += payload_bytes;
         type = KIND_RAW_ARRAY;
      } else {
         payload_bytes = 8192;
         src = data + pos;
         pos += payload_bytes;
         type = KIND_BITSET;
         desc.aux = TILE_WORDS << 16;  // first word 0, all 1024 words
      }
      if (pos > size) {
         bad("truncated container payload");
      }
      if (descs_out != nullptr) {
         desc.offset4 = static_cast<uint32_t>(sizes.payload_bytes / 4);
         desc.packed = DevContainer::pack(cardinality, 0, 0, type);
         std::memcpy(payload_out + sizes.payload_bytes, src, payload_bytes);
         const uint64_t padded = (payload_bytes + 15) / 16 * 16;
         std::memset(payload_out + sizes.payload_bytes + payload_bytes, 0, padded - payload_bytes);
         descs_out[i] = desc;
      }
      sizes.payload_bytes += (payload_bytes + 15) / 16 * 16;
   }
   return sizes;
}

void validateProgram(const silo_gpu_table* table, const silo_filter_program* program, uint32_t* max_depth_out, bool* has_threshold_out) {
   auto bad = [](const std::string& what) { throw ApiError(SILO_E_BAD_PROGRAM, "filter program: " + what); };
   if (program->struct_size != sizeof(silo_filter_program)) {
      bad("struct_size mismatch");
   }
   if (program->n_instrs == 0 || program->instrs == nullptr) {
      bad("empty program");
   }
   int depth = 0;
   int max_depth = 1;
   bool any_threshold = false;
   bool in_threshold = false;
   int threshold_base = 0;
   uint64_t adds = 0;
   uint64_t bias = 0;
   uint64_t target = 0;
   auto needColumn = [&](const silo_filter_instr& ins) -> const DevColumn& {
      if (ins.column >= table->columns.size()) {
         bad("column index out of range");
      }
      return table->columns[ins.column]->dev;
   };
   auto needBlob = [&](uint64_t offset, uint64_t bytes) {
      if (offset % 4 != 0 || offset + bytes > program->blob_bytes || (bytes > 0 && program->blob == nullptr)) {
         bad("blob reference out of range or misaligned");
      }
   };
   for (uint32_t pc = 0; pc < program->n_instrs; ++pc) {
      const silo_filter_instr& ins = program->instrs[pc];
      switch (ins.opcode) {
         case SILO_OP_PUSH_EMPTY:
         case SILO_OP_PUSH_FULL:
            ++depth;
            break;
         case SILO_OP_PUSH_SYMBOLS:
         case SILO_OP_PUSH_COVERED: {
            const DevColumn& column = needColumn(ins);
            if (ins.a >= column.genome_length) {
               bad("position out of range");
            }
            ++depth;
            break;
         }
         case SILO_OP_PUSH_NULLS:
            needColumn(ins);
            ++depth;
            break;
         case SILO_OP_PUSH_BITMAP:
            if (ins.a >= program->n_bitmaps) {
               bad("bitmap index out of range");
            }
            ++depth;
            break;
         case SILO_OP_PUSH_INDEX_BITMAP:
            if (ins.a >= table->registered.size() || table->registered[ins.a].d_block == nullptr) {
               bad("index bitmap id is not registered");
            }
            ++depth;
            break;
         case SILO_OP_PUSH_COMPARE:
            if (ins.column >= table->value_columns.size()) {
               bad("value column index out of range");
            }
            if ((ins.flags & 7u) == SILO_CMP_IN_SET) {
               needBlob(ins.b, 4ULL * ins.a);
               const uint32_t* set = reinterpret_cast<const uint32_t*>(program->blob + ins.b);
               for (uint32_t i = 1; i < ins.a; ++i) {
                  if (set[i - 1] >= set[i]) {
                     bad("PUSH_COMPARE set values must be strictly ascending");
                  }
               }
            }
            ++depth;
            break;
         case SILO_OP_PUSH_RANGES:
            needBlob(ins.b, 8ULL * ins.a);
            if (ins.b % 8 != 0) {
               bad("PUSH_RANGES pairs must be 8-byte aligned in the blob");
            }
            for (uint32_t i = 0; i < ins.a; ++i) {
               const uint32_t* range = reinterpret_cast<const uint32_t*>(program->blob + ins.b) + 2 * i;
               if (range[0] > range[1]) {
                  bad("range start > end");
               }
            }
            ++depth;
            break;
         case SILO_OP_AND:
         case SILO_OP_ANDNOT:
         case SILO_OP_OR:
            if (depth - (in_threshold ? threshold_base : 0) < 2) {
               bad("binary operator needs two operands");
            }
            --depth;
            break;
         case SILO_OP_NOT:
            if (depth - (in_threshold ? threshold_base : 0) < 1) {
               bad("NOT needs an operand");
            }
            break;
         case SILO_OP_THR_BEGIN:
            if (in_threshold) {
               bad("nested thresholds are not supported by the shared-memory counter tile");
            }
            in_threshold = true;
            any_threshold = true;
            threshold_base = depth;
            adds = 0;
            bias = ins.b & 0xFFFF;
            target = ins.a;
            break;
         case SILO_OP_THR_ADD:
            // the child tile may have been pushed in front of THR_BEGIN (a child that is a counter program itself is
            // evaluated first: counter programs do not nest); the threshold's base sinks with it
            if (!in_threshold || depth < 1) {
               bad("THR_ADD needs a child tile inside a threshold");
            }
            --depth;
            threshold_base = std::min(threshold_base, depth);
            ++adds;
            break;
         case SILO_OP_THR_ADD_SYMBOLS: {
            const DevColumn& column = needColumn(ins);
            if (!in_threshold || ins.a >= column.genome_length) {
               bad("THR_ADD_SYMBOLS outside a threshold or position out of range");
            }
            ++adds;
            break;
         }
         case SILO_OP_THR_ADD_COVERED: {
            const DevColumn& column = needColumn(ins);
            if (!in_threshold) {
               bad("THR_ADD_COVERED outside a threshold");
            }
            needBlob(ins.b, 4ULL * ins.a);
            const uint32_t* positions = reinterpret_cast<const uint32_t*>(program->blob + ins.b);
            for (uint32_t i = 0; i < ins.a; ++i) {
               if (positions[i] >= column.genome_length || (i > 0 && positions[i - 1] >= positions[i])) {
                  bad("THR_ADD_COVERED positions must be strictly ascending and in range");
               }
            }
            adds += ins.a;
            break;
         }
         case SILO_OP_THR_PROFILE: {
            const DevColumn& column = needColumn(ins);
            if (!in_threshold) {
               bad("THR_PROFILE outside a threshold");
            }
            if (ins.b % 8 != 0) {
               bad("THR_PROFILE table must be 8-byte aligned");
            }
            needBlob(ins.b, 8ULL * column.genome_length);
            adds += column.genome_length;
            break;
         }
         case SILO_OP_THR_END:
            if (!in_threshold || depth != threshold_base) {
               bad("THR_END with children left on the stack");
            }
            if (bias + adds > 65535 || target + bias > 65535) {
               throw ApiError(SILO_E_UNSUPPORTED, "threshold needs more than 65535 counts per row");
            }
            in_threshold = false;
            ++depth;
            break;
         default:
            bad("unknown opcode");
      }
      max_depth = std::max(max_depth, depth);
      if (depth > STACK_DEPTH) {
         throw ApiError(SILO_E_UNSUPPORTED, "filter program needs a deeper tile stack than the kernel provides");
      }
   }
   *max_depth_out = static_cast<uint32_t>(max_depth);
   *has_threshold_out = any_threshold;
   if (in_threshold || depth != 1) {
      bad("program must leave exactly one tile on the stack");
   }
}

silo_gpu_filter* allocFilter(silo_gpu_table* table) {
   auto filter = std::make_unique<silo_gpu_filter>();
   filter->table = table;
   const size_t words_bytes = static_cast<size_t>(table->n_chunks) * TILE_BYTES;
   const size_t popcount_bytes = (static_cast<size_t>(table->n_chunks) * sizeof(uint32_t) + 15) / 16 * 16;
   uint8_t* base = poolAlloc<uint8_t>(words_bytes + popcount_bytes + 32, table->ctx->stream);
   filter->d_words = reinterpret_cast<uint64_t*>(base);
   filter->d_chunk_popcount = reinterpret_cast<uint32_t*>(base + words_bytes);
   filter->d_cardinality = reinterpret_cast<unsigned long long*>(base + words_bytes + popcount_bytes);
   filter->d_error_flag = reinterpret_cast<uint32_t*>(base + words_bytes + popcount_bytes + 16);
   return filter.release();
}

// stream-ordered free on the table's stream, ordered after whatever foreign stream used the memory last
void poolFreeAfterUsers(silo_gpu_table* table, void* ptr) {
   if (ptr == nullptr) {
      return;
   }
   cudaStream_t stream = table->ctx->stream;
   if (table->last_stream != nullptr && table->last_stream != stream) {
      cudaEventRecord(table->ev_free_fence, table->last_stream);
      cudaStreamWaitEvent(stream, table->ev_free_fence, 0);
   }
   cudaFreeAsync(ptr, stream);
}

// caller holds table->mutex
void freeFilterLocked(silo_gpu_filter* filter) {
   if (filter == nullptr) {
      return;
   }
   poolFreeAfterUsers(filter->table, filter->d_words);
   delete filter;
}

}  // namespace

}  // namespace silo

using namespace silo;

extern "C" {

void silo_gpu_filter_free(silo_gpu_filter* filter) {
   if (filter == nullptr) {
      return;
   }
   cudaSetDevice(filter->table->ctx->device);
   std::lock_guard<std::mutex> lock(filter->table->mutex);
   freeFilterLocked(filter);
}

// staged program: everything the kernel reads, written straight into pinned memory and uploaded with
// one H2D copy: [instrs | columns | bitmap table | blob | per-bitmap descriptors | payloads | pad]
static void stageProgram(
   silo_gpu_table* table,
   const silo_filter_program* program,
   cudaStream_t stream,
   uint8_t** d_staging_out,
   uint64_t* staged_bytes_out,
   EvalParams* params_out,
   bool persistent = false  // device side = table->d_staging_fixed, the H2D copy is left to the caller
) {
   uint32_t stack_depth = 1;
   bool has_threshold = false;
   validateProgram(table, program, &stack_depth, &has_threshold);

   std::vector<RoaringSizes> sizes(program->n_bitmaps);
   uint64_t total_descs = 0;
   uint64_t total_payload = 0;
   for (uint32_t i = 0; i < program->n_bitmaps; ++i) {
      sizes[i] = parseRoaring(program->bitmaps[i].data, program->bitmaps[i].size, nullptr, nullptr);
      total_descs += sizes[i].n_containers;
      total_payload += sizes[i].payload_bytes;
   }
   // registered (device-resident) index bitmaps referenced by the program get a slot of the
   // program's own bitmap table behind the uploaded ones
   std::vector<uint32_t> registered_used;
   for (uint32_t pc = 0; pc < program->n_instrs; ++pc) {
      if (program->instrs[pc].opcode == SILO_OP_PUSH_INDEX_BITMAP) {
         registered_used.push_back(program->instrs[pc].a);
      }
   }
   std::sort(registered_used.begin(), registered_used.end());
   registered_used.erase(std::unique(registered_used.begin(), registered_used.end()), registered_used.end());

   size_t cursor = 0;
   auto reserve = [&](size_t bytes) {
      cursor = (cursor + 15) / 16 * 16;
      const size_t offset = cursor;
      cursor += bytes;
      return offset;
   };
   const size_t n_table = program->n_bitmaps + registered_used.size();
   const size_t off_instrs = reserve(sizeof(silo_filter_instr) * program->n_instrs);
   const size_t off_columns = reserve(sizeof(DevColumn) * table->columns.size());
   const size_t off_value_columns = reserve(sizeof(DevValueColumn) * table->value_columns.size());
   const size_t off_bitmaps = reserve(sizeof(DevBitmap) * n_table);
   const size_t off_blob = reserve(program->blob_bytes);
   const size_t off_descs = reserve(sizeof(DevContainer) * total_descs);
   const size_t off_payload = reserve(total_payload + 16);  // the kernel reads whole 16-byte vectors
   const size_t staging_bytes = (cursor + 15) / 16 * 16;

   // The pinned buffer is reused by the next call on this table. The synchronous callers synchronise the stream before
   // they return; behind the copy of a call that does not (the _async and sharded entries) an event is recorded, and
   // the next staging waits for it here -- the host runs at most one query ahead of the copies.
   if (table->staging_copy_pending) {
      SILO_CUDA_CHECK(cudaEventSynchronize(table->ev_staging_copied));
      table->staging_copy_pending = false;
   }
   if (staging_bytes > table->staging_capacity) {
      dropQueryGraphsLocked(table);  // their copy nodes read the old buffers
      SILO_CUDA_CHECK(cudaStreamSynchronize(table->ctx->stream));
      if (table->h_staging_pinned != nullptr) {
         cudaFreeHost(table->h_staging_pinned);
         table->h_staging_pinned = nullptr;
         table->staging_capacity = 0;
      }
      if (table->d_staging_fixed != nullptr) {
         cudaFree(table->d_staging_fixed);
         table->d_staging_fixed = nullptr;
      }
      const size_t capacity = std::max<size_t>(staging_bytes * 2, 1 << 20);
      const cudaError_t pinned_status = cudaMallocHost(reinterpret_cast<void**>(&table->h_staging_pinned), capacity);
      if (pinned_status != cudaSuccess) {
         throw ApiError(SILO_E_OUT_OF_MEMORY, std::string("pinned staging: ") + cudaGetErrorString(pinned_status));
      }
      table->staging_capacity = capacity;
   }
   uint8_t* staging = table->h_staging_pinned;
   if (persistent && table->d_staging_fixed == nullptr) {
      table->d_staging_fixed = deviceAlloc<uint8_t>(table->staging_capacity, &table->device_bytes);
   }
   uint8_t* d_staging = persistent ? table->d_staging_fixed : poolAlloc<uint8_t>(staging_bytes, stream);

   auto* instrs = reinterpret_cast<silo_filter_instr*>(staging + off_instrs);
   std::memcpy(instrs, program->instrs, sizeof(silo_filter_instr) * program->n_instrs);
   for (uint32_t pc = 0; pc < program->n_instrs; ++pc) {
      if (instrs[pc].opcode == SILO_OP_PUSH_INDEX_BITMAP) {
         const auto slot = std::lower_bound(registered_used.begin(), registered_used.end(), instrs[pc].a);
         instrs[pc].opcode = SILO_OP_PUSH_BITMAP;
         instrs[pc].a = program->n_bitmaps + static_cast<uint32_t>(slot - registered_used.begin());
      }
   }
   auto* columns = reinterpret_cast<DevColumn*>(staging + off_columns);
   for (size_t i = 0; i < table->columns.size(); ++i) {
      columns[i] = table->columns[i]->dev;
   }
   if (!table->value_columns.empty()) {
      std::memcpy(staging + off_value_columns, table->value_columns.data(), sizeof(DevValueColumn) * table->value_columns.size());
   }
   if (program->blob_bytes > 0) {
      std::memcpy(staging + off_blob, program->blob, program->blob_bytes);
   }
   auto* bitmaps = reinterpret_cast<DevBitmap*>(staging + off_bitmaps);
   size_t desc_cursor = 0;
   size_t payload_cursor = 0;
   for (uint32_t i = 0; i < program->n_bitmaps; ++i) {
      parseRoaring(
         program->bitmaps[i].data, program->bitmaps[i].size,
         reinterpret_cast<DevContainer*>(staging + off_descs) + desc_cursor, staging + off_payload + payload_cursor
      );
      bitmaps[i].containers = reinterpret_cast<const DevContainer*>(d_staging + off_descs) + desc_cursor;
      bitmaps[i].payload = d_staging + off_payload + payload_cursor;
      bitmaps[i].n_containers = sizes[i].n_containers;
      bitmaps[i].pad = 0;
      desc_cursor += sizes[i].n_containers;
      payload_cursor += sizes[i].payload_bytes;
   }
   std::memset(staging + off_payload + payload_cursor, 0, 16);
   for (size_t k = 0; k < registered_used.size(); ++k) {
      const silo_gpu_table::RegisteredBitmap& registered = table->registered[registered_used[k]];
      DevBitmap& bitmap = bitmaps[program->n_bitmaps + k];
      bitmap.containers = reinterpret_cast<const DevContainer*>(registered.d_block);
      bitmap.payload = registered.d_block + registered.payload_offset;
      bitmap.n_containers = registered.n_containers;
      bitmap.pad = 0;
   }
   if (!persistent) {
      const cudaError_t status = cudaMemcpyAsync(d_staging, staging, staging_bytes, cudaMemcpyHostToDevice, stream);
      if (status != cudaSuccess) {
         cudaFreeAsync(d_staging, stream);
         throw ApiError(SILO_E_CUDA, std::string("program upload failed: ") + cudaGetErrorString(status));
      }
   }
   EvalParams params{};
   params.instrs = reinterpret_cast<const silo_filter_instr*>(d_staging + off_instrs);
   params.n_instrs = program->n_instrs;
   params.n_chunks = table->n_chunks;
   params.columns = reinterpret_cast<const DevColumn*>(d_staging + off_columns);
   params.blob = d_staging + off_blob;
   params.bitmaps = bitmaps == nullptr ? nullptr : reinterpret_cast<const DevBitmap*>(d_staging + off_bitmaps);
   params.chunk_sizes = table->d_chunk_sizes;
   params.value_columns = reinterpret_cast<const DevValueColumn*>(d_staging + off_value_columns);
   params.chunk_row_begin = table->d_chunk_row_begin;
   params.first_chunk = table->first_chunk;
   params.stack_depth = stack_depth;
   params.has_threshold = has_threshold ? 1u : 0u;
   // The first THR_PROFILE over a large column is counted by the sweep kernel in front of the interpreter.
   params.sweep_pc = NO_SWEEP;
   for (uint32_t pc = 0; pc < program->n_instrs; ++pc) {
      const silo_filter_instr& ins = program->instrs[pc];
      if (ins.opcode != SILO_OP_THR_PROFILE) {
         continue;
      }
      const HostColumn& host = *table->columns[ins.column];
      if (host.dev.n_containers < table->sweep_min_pieces || table->n_chunks == 0) {
         continue;
      }
      // lanes of the sweep start at `bias` = the number of positions with a subtracting mask (all a row can lose)
      const uint32_t* masks = reinterpret_cast<const uint32_t*>(program->blob + ins.b);
      uint32_t bias = 0;
      for (uint32_t position = 0; position < host.dev.genome_length; ++position) {
         bias += masks[2 * position + 1] != 0 ? 1u : 0u;
      }
      // validateProgram bounded the program's own lanes by 65535; the interpreter adds the sweep's surplus on top
      // before it takes it off again
      if (static_cast<uint64_t>(host.sweep_max_flushes + 1) * bias + host.dev.genome_length + 2ULL * bias > 65535) {
         continue;
      }
      if (table->d_sweep_counters == nullptr) {
         table->d_sweep_counters = deviceAlloc<uint32_t>(static_cast<size_t>(table->n_chunks) * 32768, &table->device_bytes);
      }
      params.sweep_pc = pc;
      params.sweep_counters = table->d_sweep_counters;
      params.sweep_flushes = host.d_sweep_flushes;
      params.sweep_bias = bias;
      params.sweep_column = ins.column;
      params.sweep_table_offset = ins.b;
      break;
   }
   *d_staging_out = d_staging;
   *staged_bytes_out = staging_bytes;
   *params_out = params;
}

static void launchProgram(silo_gpu_table* table, EvalParams params, silo_gpu_filter* filter, cudaStream_t stream, bool scalars_are_zero = false) {
   params.out_words = filter->d_words;
   params.out_popcount = filter->d_chunk_popcount;
   params.out_cardinality = filter->d_cardinality;
   params.error_flag = filter->d_error_flag;
   if (!scalars_are_zero) {
      SILO_CUDA_CHECK(cudaMemsetAsync(filter->d_cardinality, 0, 32, stream));  // cardinality and, 16 bytes on, the error flag (allocFilter)
   }
   if (table->n_chunks == 0) {
      return;
   }
   static bool attribute_set = false;
   if (!attribute_set) {
      SILO_CUDA_CHECK(cudaFuncSetAttribute(
         evalProgramKernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
         static_cast<int>(evalSharedBytes(STACK_DEPTH, true))
      ));
      SILO_CUDA_CHECK(cudaFuncSetAttribute(thresholdSweepKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(COUNTER_BYTES)));
      attribute_set = true;
   }
   if (params.sweep_pc != NO_SWEEP) {
      const HostColumn& host = *table->columns[static_cast<size_t>(params.sweep_column)];
      SILO_CUDA_CHECK(cudaMemsetAsync(table->d_sweep_counters, 0, static_cast<size_t>(table->n_chunks) * COUNTER_BYTES, stream));
      cudaStreamCaptureStatus capture_status = cudaStreamCaptureStatusNone;
      const bool timed = cudaStreamIsCapturing(stream, &capture_status) == cudaSuccess && capture_status == cudaStreamCaptureStatusNone;
      const int slot = static_cast<int>(table->sweep_timed_calls % silo_gpu_table::SWEEP_EVENT_RING);
      if (timed) {
         if (table->ev_sweep_begin[slot] == nullptr) {
            SILO_CUDA_CHECK(cudaEventCreate(&table->ev_sweep_begin[slot]));
            SILO_CUDA_CHECK(cudaEventCreate(&table->ev_sweep_end[slot]));
         }
         SILO_CUDA_CHECK(cudaEventRecord(table->ev_sweep_begin[slot], stream));
      }
      const NvtxRange sweep_range("Threshold: evaluate [thresholdSweepKernel]");
      thresholdSweepKernel<<<host.sweep_ctas, SWEEP_THREADS, COUNTER_BYTES, stream>>>(
         host.dev, reinterpret_cast<const uint2*>(params.blob + params.sweep_table_offset), host.d_sweep_split, params.sweep_bias,
         table->d_sweep_counters
      );
      SILO_CUDA_CHECK(cudaGetLastError());
      table->stats.kernel_launches++;
      if (timed) {
         SILO_CUDA_CHECK(cudaEventRecord(table->ev_sweep_end[slot], stream));
         table->sweep_timed_calls++;
         table->sweep_stream = stream;
         table->sweep_algorithmic_bytes = 0;
         for (uint64_t bytes : host.chunk_desc_payload_bytes) {
            table->sweep_algorithmic_bytes += bytes;
         }
      }
   }
   const size_t shared_bytes = evalSharedBytes(params.stack_depth, params.has_threshold != 0);
   const NvtxRange eval_range("computeFilter: Intersection / Union / Threshold / Selection evaluate [evalProgramKernel]");
   SILO_CUDA_CHECK(launchDependent(evalProgramKernel, dim3(table->n_chunks), dim3(EVAL_THREADS), shared_bytes, stream, params));
   SILO_CUDA_CHECK(cudaGetLastError());
   table->stats.kernel_launches++;
}

}  // extern "C"

namespace silo {

silo_gpu_filter* evalProgramAsync(silo_gpu_table* table, const silo_filter_program* program, cudaStream_t stream, uint8_t** d_staging_out) {
   uint64_t staged_bytes = 0;
   EvalParams params{};
   stageProgram(table, program, stream, d_staging_out, &staged_bytes, &params);
   silo_gpu_filter* filter = nullptr;
   try {
      filter = allocFilter(table);
      launchProgram(table, params, filter, stream);
   } catch (...) {
      cudaFreeAsync(*d_staging_out, stream);
      *d_staging_out = nullptr;
      cudaStreamSynchronize(stream);
      freeFilterLocked(filter);
      throw;
   }
   return filter;
}

void releaseFilterLocked(silo_gpu_filter* filter) {
   freeFilterLocked(filter);
}

void dropQueryGraphsLocked(silo_gpu_table* table) {
   for (silo_gpu_table::CachedGraph& cached : table->query_graphs) {
      if (cached.exec != nullptr) {
         cudaGraphExecDestroy(cached.exec);
      }
   }
   table->query_graphs.clear();
   table->last_query_key.clear();
}

void stageQueryLocked(silo_gpu_table* table, const silo_filter_program* program, StagedQuery* out, int prepare_column, uint32_t* prepare_counts) {
   static_assert(sizeof(EvalParams) <= sizeof(out->params));
   if (table->query_filter == nullptr) {  // persistent: plain device memory, not the stream-ordered pool
      auto filter = std::make_unique<silo_gpu_filter>();
      filter->table = table;
      const size_t words_bytes = static_cast<size_t>(table->n_chunks) * TILE_BYTES;
      const size_t popcount_bytes = (static_cast<size_t>(table->n_chunks) * sizeof(uint32_t) + 15) / 16 * 16;
      uint8_t* base = deviceAlloc<uint8_t>(words_bytes + popcount_bytes + 32, &table->device_bytes);
      filter->d_words = reinterpret_cast<uint64_t*>(base);
      filter->d_chunk_popcount = reinterpret_cast<uint32_t*>(base + words_bytes);
      filter->d_cardinality = reinterpret_cast<unsigned long long*>(base + words_bytes + popcount_bytes);
      filter->d_error_flag = reinterpret_cast<uint32_t*>(base + words_bytes + popcount_bytes + 16);
      // cardinality and error flag are zero between queries: the finalize kernel of a fused query resets them
      SILO_CUDA_CHECK(cudaMemsetAsync(filter->d_cardinality, 0, 32, table->ctx->stream));
      table->query_filter = filter.release();
   }
   uint8_t* d_staging = nullptr;
   EvalParams params{};
   stageProgram(table, program, table->ctx->stream, &d_staging, &out->staged_bytes, &params, true);
   params.out_words = table->query_filter->d_words;
   params.out_popcount = table->query_filter->d_chunk_popcount;
   params.out_cardinality = table->query_filter->d_cardinality;
   params.error_flag = table->query_filter->d_error_flag;
   if (prepare_column >= 0 && table->n_chunks > 0) {
      const HostColumn& host = *table->columns[static_cast<size_t>(prepare_column)];
      if (host.dev.n_segments > 0) {
         params.prepare_segments = host.dev.segments;
         params.prepare_chunk_seg_begin = host.dev.chunk_seg_begin;
         params.prepare_work_state = table->d_work_state;
         params.prepare_work_items = table->d_work_items;
      }
      params.prepare_counts = prepare_counts;
      params.prepare_counts_words = host.dev.n_symbols * host.dev.genome_length;
   }
   std::memset(out->params, 0, sizeof(out->params));
   std::memcpy(out->params, &params, sizeof(params));
   out->shared_bytes = static_cast<uint32_t>(evalSharedBytes(params.stack_depth, params.has_threshold != 0));
}

void enqueueStagedQuery(silo_gpu_table* table, const StagedQuery& staged, cudaStream_t stream, bool scalars_are_zero) {
   EvalParams params{};
   std::memcpy(&params, staged.params, sizeof(params));
   SILO_CUDA_CHECK(cudaMemcpyAsync(table->d_staging_fixed, table->h_staging_pinned, staged.staged_bytes, cudaMemcpyHostToDevice, stream));
   cudaStreamCaptureStatus capture_status = cudaStreamCaptureStatusNone;
   if (cudaStreamIsCapturing(stream, &capture_status) == cudaSuccess && capture_status == cudaStreamCaptureStatusNone) {
      SILO_CUDA_CHECK(cudaEventRecord(table->ev_staging_copied, stream));
      table->staging_copy_pending = true;
   }
   launchProgram(table, params, table->query_filter, stream, scalars_are_zero);
}

}  // namespace silo

struct silo_gpu_program {
   silo_gpu_table* table = nullptr;
   uint8_t* d_staging = nullptr;
   uint64_t staged_bytes = 0;
   silo::EvalParams* params = nullptr;  // heap copy (EvalParams lives in an unnamed namespace)
   silo_gpu_filter* filter = nullptr;
};

extern "C" {

int silo_gpu_filter_eval(
   silo_gpu_table* table,
   const silo_filter_program* program,
   silo_gpu_filter** out,
   uint64_t* cardinality
) {
   return guarded([&] {
      require(table != nullptr && program != nullptr && out != nullptr, "silo_gpu_filter_eval: NULL argument");
      std::lock_guard<std::mutex> lock(table->mutex);
      SILO_CUDA_CHECK(cudaSetDevice(table->ctx->device));
      cudaStream_t stream = table->ctx->stream;
      uint8_t* d_staging = nullptr;
      uint64_t staged_bytes = 0;
      EvalParams params{};
      stageProgram(table, program, stream, &d_staging, &staged_bytes, &params);
      std::unique_ptr<silo_gpu_filter, void (*)(silo_gpu_filter*)> filter(nullptr, freeFilterLocked);
      try {
         filter.reset(allocFilter(table));
         launchProgram(table, params, filter.get(), stream);
         unsigned long long host_cardinality = 0;
         uint32_t host_error = 0;
         SILO_CUDA_CHECK(cudaMemcpyAsync(&host_cardinality, filter->d_cardinality, sizeof(host_cardinality), cudaMemcpyDeviceToHost, stream));
         SILO_CUDA_CHECK(cudaMemcpyAsync(&host_error, filter->d_error_flag, sizeof(host_error), cudaMemcpyDeviceToHost, stream));
         SILO_CUDA_CHECK(cudaFreeAsync(d_staging, stream));
         d_staging = nullptr;
         SILO_CUDA_CHECK(cudaStreamSynchronize(stream));
         filter->out_of_layout = host_error != 0;  // kept, like the reference's bitmaps keep such ids (see silo_b200.h)
         if (cardinality != nullptr) {
            *cardinality = host_cardinality;
         }
      } catch (...) {
         if (d_staging != nullptr) {
            cudaFreeAsync(d_staging, stream);
         }
         cudaStreamSynchronize(stream);
         throw;
      }
      *out = filter.release();
   });
}

// CountFilterNode (count_filter_node.cpp:35-71): the filter's cardinality only. The program runs on the table's persistent
// query buffers, so a query SHAPE (the same instructions with other positions / values) is a replayed CUDA graph: staging
// copy, interpreter, the 32 bytes of {cardinality, error flag} into page-locked memory, scalars reset; one synchronisation.
int silo_gpu_query_count(silo_gpu_table* table, const silo_filter_program* program, uint64_t* cardinality) {
   return guarded([&] {
      require(table != nullptr && program != nullptr && cardinality != nullptr, "silo_gpu_query_count: NULL argument");
      std::lock_guard<std::mutex> lock(table->mutex);
      SILO_CUDA_CHECK(cudaSetDevice(table->ctx->device));
      cudaStream_t stream = table->ctx->stream;
      StagedQuery staged;
      stageQueryLocked(table, program, &staged);
      auto enqueueAll = [&]() {
         enqueueStagedQuery(table, staged, stream);
         SILO_CUDA_CHECK(cudaMemcpyAsync(table->h_scalars_pinned, table->query_filter->d_cardinality, 32, cudaMemcpyDeviceToHost, stream));
         SILO_CUDA_CHECK(cudaMemsetAsync(table->query_filter->d_cardinality, 0, 32, stream));  // zero between queries
      };
      std::string key(reinterpret_cast<const char*>(staged.params), sizeof(staged.params));
      const uint64_t scalars[] = {0x434F554E54ULL, staged.staged_bytes, staged.shared_bytes};
      key.append(reinterpret_cast<const char*>(scalars), sizeof(scalars));
      try {
         cudaGraphExec_t replay = queryGraphFor(table, std::move(key), stream, enqueueAll);
         if (replay != nullptr) {
            SILO_CUDA_CHECK(cudaGraphLaunch(replay, stream));
            table->stats.kernel_launches += 1;
         } else {
            enqueueAll();
         }
         SILO_CUDA_CHECK(cudaStreamSynchronize(stream));
      } catch (...) {
         cudaStreamSynchronize(stream);
         cudaMemsetAsync(table->query_filter->d_cardinality, 0, 32, stream);
         cudaStreamSynchronize(stream);
         throw;
      }
      *cardinality = table->h_scalars_pinned[0];  // (rows outside the layout that a leaf bitmap held are counted, as by the reference)
   });
}

int silo_gpu_program_prepare(silo_gpu_table* table, const silo_filter_program* program, silo_gpu_program** out, silo_gpu_filter** filter_out) {
   return guarded([&] {
      require(table != nullptr && program != nullptr && out != nullptr && filter_out != nullptr, "silo_gpu_program_prepare: NULL argument");
      std::lock_guard<std::mutex> lock(table->mutex);
      SILO_CUDA_CHECK(cudaSetDevice(table->ctx->device));
      auto prepared = std::make_unique<silo_gpu_program>();
      prepared->table = table;
      EvalParams params{};
      stageProgram(table, program, table->ctx->stream, &prepared->d_staging, &prepared->staged_bytes, &params);
      try {
         prepared->params = new EvalParams(params);
         prepared->filter = allocFilter(table);
         SILO_CUDA_CHECK(cudaStreamSynchronize(table->ctx->stream));  // the upload left pinned memory
      } catch (...) {
         cudaFreeAsync(prepared->d_staging, table->ctx->stream);
         cudaStreamSynchronize(table->ctx->stream);
         delete prepared->params;
         throw;
      }
      *filter_out = prepared->filter;
      *out = prepared.release();
   });
}

int silo_gpu_program_run_async(silo_gpu_program* prepared, void* cuda_stream) {
   return guarded([&] {
      require(prepared != nullptr, "silo_gpu_program_run_async: NULL argument");
      silo_gpu_table* table = prepared->table;
      std::lock_guard<std::mutex> lock(table->mutex);
      SILO_CUDA_CHECK(cudaSetDevice(table->ctx->device));
      cudaStream_t stream = cuda_stream != nullptr ? static_cast<cudaStream_t>(cuda_stream) : table->ctx->stream;
      launchProgram(table, *prepared->params, prepared->filter, stream);
   });
}

int silo_gpu_program_run_counts_async(silo_gpu_program* prepared, int column, void* d_counts, void* cuda_stream) {
   return guarded([&] {
      require(prepared != nullptr && d_counts != nullptr, "silo_gpu_program_run_counts_async: NULL argument");
      silo_gpu_table* table = prepared->table;
      std::lock_guard<std::mutex> lock(table->mutex);
      SILO_CUDA_CHECK(cudaSetDevice(table->ctx->device));
      require(column >= 0 && static_cast<size_t>(column) < table->columns.size(), "silo_gpu_program_run_counts_async: bad column index");
      cudaStream_t stream = cuda_stream != nullptr ? static_cast<cudaStream_t>(cuda_stream) : table->ctx->stream;
      EvalParams params = *prepared->params;
      const HostColumn& host = *table->columns[static_cast<size_t>(column)];
      if (table->n_chunks > 0) {
         if (host.dev.n_segments > 0) {
            params.prepare_segments = host.dev.segments;
            params.prepare_chunk_seg_begin = host.dev.chunk_seg_begin;
            params.prepare_work_state = table->d_work_state;
            params.prepare_work_items = table->d_work_items;
         }
         params.prepare_counts = static_cast<uint32_t*>(d_counts);
         params.prepare_counts_words = host.dev.n_symbols * host.dev.genome_length;
      }
      launchProgram(table, params, prepared->filter, stream);
      enqueuePreparedCountsLocked(table, column, prepared->filter, static_cast<uint32_t*>(d_counts), stream);
   });
}

static int runPreparedSharded(silo_gpu_program* prepared, void* cuda_stream, bool collect_here, void* d_summed_counts) {
   return guarded([&] {
      require(prepared != nullptr, "silo_gpu_program_run_sharded_async: NULL argument");
      silo_gpu_table* table = prepared->table;
      std::lock_guard<std::mutex> lock(table->mutex);
      SILO_CUDA_CHECK(cudaSetDevice(table->ctx->device));
      const int column = shardGroupColumnLocked(table);
      require(table->n_chunks > 0, "silo_gpu_program_run_sharded_async: a shard must hold at least one chunk");
      cudaStream_t stream = cuda_stream != nullptr ? static_cast<cudaStream_t>(cuda_stream) : table->ctx->stream;
      EvalParams params = *prepared->params;
      const HostColumn& host = *table->columns[static_cast<size_t>(column)];
      if (host.dev.n_segments > 0) {
         params.prepare_segments = host.dev.segments;
         params.prepare_chunk_seg_begin = host.dev.chunk_seg_begin;
         params.prepare_work_state = table->d_work_state;
         params.prepare_work_items = table->d_work_items;
      }
      params.prepare_counts = table->d_counts;
      params.prepare_counts_words = host.dev.n_symbols * host.dev.genome_length;
      launchProgram(table, params, prepared->filter, stream);
      enqueuePreparedShardedLocked(table, prepared->filter, stream, collect_here, d_summed_counts);
   });
}

int silo_gpu_program_run_sharded_async(silo_gpu_program* prepared, void* cuda_stream) {
   return runPreparedSharded(prepared, cuda_stream, false, nullptr);
}

int silo_gpu_program_run_sharded_collect_async(silo_gpu_program* prepared, void* d_summed_counts, void* cuda_stream) {
   return runPreparedSharded(prepared, cuda_stream, true, d_summed_counts);
}

uint64_t silo_gpu_program_device_bytes(const silo_gpu_program* prepared) {
   return prepared == nullptr ? 0 : prepared->staged_bytes;
}

void silo_gpu_program_free(silo_gpu_program* prepared) {
   if (prepared == nullptr) {
      return;
   }
   cudaSetDevice(prepared->table->ctx->device);
   {
      std::lock_guard<std::mutex> lock(prepared->table->mutex);
      poolFreeAfterUsers(prepared->table, prepared->d_staging);
   }
   delete prepared->params;
   delete prepared;
}

int silo_gpu_get_sweep_stats(silo_gpu_table* table, float* mean_kernel_ms, uint64_t* algorithmic_bytes, uint64_t* timed_calls) {
   return guarded([&] {
      require(table != nullptr && mean_kernel_ms != nullptr && algorithmic_bytes != nullptr && timed_calls != nullptr, "silo_gpu_get_sweep_stats: NULL argument");
      std::lock_guard<std::mutex> lock(table->mutex);
      SILO_CUDA_CHECK(cudaSetDevice(table->ctx->device));
      if (table->sweep_stream != nullptr) {
         SILO_CUDA_CHECK(cudaStreamSynchronize(table->sweep_stream));
      }
      const uint64_t n = std::min<uint64_t>(table->sweep_timed_calls, silo_gpu_table::SWEEP_EVENT_RING);
      double total_ms = 0;
      uint64_t valid = 0;
      for (uint64_t back = 0; back < n; ++back) {
         const uint64_t slot = (table->sweep_timed_calls - 1 - back) % silo_gpu_table::SWEEP_EVENT_RING;
         float ms = 0;
         if (cudaEventElapsedTime(&ms, table->ev_sweep_begin[slot], table->ev_sweep_end[slot]) == cudaSuccess) {
            total_ms += ms;
            ++valid;
         } else {
            cudaGetLastError();
         }
      }
      *mean_kernel_ms = valid > 0 ? static_cast<float>(total_ms / static_cast<double>(valid)) : 0.0f;
      *algorithmic_bytes = table->sweep_algorithmic_bytes;
      *timed_calls = valid;
      table->sweep_timed_calls = 0;
   });
}

int silo_gpu_bitmap_register(silo_gpu_table* table, const uint8_t* data, uint64_t size, uint32_t* id_out) {
   return guarded([&] {
      require(table != nullptr && data != nullptr && id_out != nullptr, "silo_gpu_bitmap_register: NULL argument");
      std::lock_guard<std::mutex> lock(table->mutex);
      SILO_CUDA_CHECK(cudaSetDevice(table->ctx->device));
      const RoaringSizes sizes = parseRoaring(data, size, nullptr, nullptr);
      silo_gpu_table::RegisteredBitmap registered;
      registered.n_containers = sizes.n_containers;
      registered.payload_offset = (sizeof(DevContainer) * static_cast<uint64_t>(sizes.n_containers) + 15) / 16 * 16;
      registered.bytes = registered.payload_offset + sizes.payload_bytes + 16;
      std::vector<uint8_t> block(registered.bytes, 0);
      parseRoaring(data, size, reinterpret_cast<DevContainer*>(block.data()), block.data() + registered.payload_offset);
      // ids outside this shard's chunk range are ignored by the interpreter (it looks containers up
      // by chunk key); ids inside a chunk but beyond its size are rejected when a program runs
      SILO_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&registered.d_block), registered.bytes));
      const cudaError_t status = cudaMemcpy(registered.d_block, block.data(), registered.bytes, cudaMemcpyHostToDevice);
      if (status != cudaSuccess) {
         cudaFree(registered.d_block);
         SILO_CUDA_CHECK(status);
      }
      table->device_bytes += registered.bytes;
      size_t slot = 0;
      while (slot < table->registered.size() && table->registered[slot].d_block != nullptr) {
         ++slot;
      }
      if (slot == table->registered.size()) {
         table->registered.emplace_back();
      }
      table->registered[slot] = registered;
      *id_out = static_cast<uint32_t>(slot);
   });
}

int silo_gpu_bitmap_unregister(silo_gpu_table* table, uint32_t id) {
   return guarded([&] {
      require(table != nullptr, "silo_gpu_bitmap_unregister: NULL argument");
      std::lock_guard<std::mutex> lock(table->mutex);
      require(id < table->registered.size() && table->registered[id].d_block != nullptr, "silo_gpu_bitmap_unregister: unknown id");
      SILO_CUDA_CHECK(cudaSetDevice(table->ctx->device));
      // programs in flight may still read the block
      SILO_CUDA_CHECK(cudaStreamSynchronize(table->ctx->stream));
      if (table->last_stream != nullptr) {
         SILO_CUDA_CHECK(cudaStreamSynchronize(table->last_stream));
      }
      cudaFree(table->registered[id].d_block);
      table->device_bytes -= table->registered[id].bytes;
      table->registered[id] = silo_gpu_table::RegisteredBitmap{};
   });
}

int silo_gpu_filter_from_words(silo_gpu_table* table, const uint64_t* words, silo_gpu_filter** out) {
   return guarded([&] {
      require(table != nullptr && out != nullptr, "silo_gpu_filter_from_words: NULL argument");
      require(table->n_chunks == 0 || words != nullptr, "silo_gpu_filter_from_words: words is NULL");
      std::lock_guard<std::mutex> lock(table->mutex);
      SILO_CUDA_CHECK(cudaSetDevice(table->ctx->device));
      cudaStream_t stream = table->ctx->stream;
      for (uint32_t chunk = 0; chunk < table->n_chunks; ++chunk) {
         const uint32_t size = table->chunk_sizes[chunk];
         for (uint32_t w = 0; w < TILE_WORDS; ++w) {
            const uint64_t layout = size >= (w + 1) * 64 ? ~0ULL : (size <= w * 64 ? 0ULL : (~0ULL >> (64 - (size - w * 64))));
            if ((words[static_cast<size_t>(chunk) * TILE_WORDS + w] & ~layout) != 0) {
               throw ApiError(SILO_E_OUT_OF_LAYOUT, "filter words hold rows outside the row layout");
            }
         }
      }
      std::unique_ptr<silo_gpu_filter, void (*)(silo_gpu_filter*)> filter(allocFilter(table), freeFilterLocked);
      SILO_CUDA_CHECK(cudaMemsetAsync(filter->d_cardinality, 0, sizeof(unsigned long long), stream));
      if (table->n_chunks > 0) {
         SILO_CUDA_CHECK(cudaMemcpyAsync(
            filter->d_words, words, static_cast<size_t>(table->n_chunks) * TILE_BYTES, cudaMemcpyHostToDevice, stream
         ));
         popcountTilesKernel<<<table->n_chunks, EVAL_THREADS, 0, stream>>>(
            filter->d_words, filter->d_chunk_popcount, filter->d_cardinality
         );
         SILO_CUDA_CHECK(cudaGetLastError());
         table->stats.kernel_launches++;
      }
      SILO_CUDA_CHECK(cudaStreamSynchronize(stream));
      *out = filter.release();
   });
}

int silo_gpu_filter_cardinality(const silo_gpu_filter* filter, uint64_t* cardinality) {
   return guarded([&] {
      require(filter != nullptr && cardinality != nullptr, "silo_gpu_filter_cardinality: NULL argument");
      SILO_CUDA_CHECK(cudaSetDevice(filter->table->ctx->device));
      unsigned long long value = 0;
      SILO_CUDA_CHECK(cudaMemcpyAsync(&value, filter->d_cardinality, sizeof(value), cudaMemcpyDeviceToHost, filter->table->ctx->stream));
      SILO_CUDA_CHECK(cudaStreamSynchronize(filter->table->ctx->stream));
      *cardinality = value;
   });
}

int silo_gpu_filter_download(const silo_gpu_filter* filter, uint64_t* words) {
   return guarded([&] {
      require(filter != nullptr && (words != nullptr || filter->table->n_chunks == 0), "silo_gpu_filter_download: NULL argument");
      SILO_CUDA_CHECK(cudaSetDevice(filter->table->ctx->device));
      if (filter->table->n_chunks > 0) {
         SILO_CUDA_CHECK(cudaMemcpyAsync(
            words, filter->d_words, static_cast<size_t>(filter->table->n_chunks) * TILE_BYTES, cudaMemcpyDeviceToHost,
            filter->table->ctx->stream
         ));
      }
      SILO_CUDA_CHECK(cudaStreamSynchronize(filter->table->ctx->stream));
   });
}

}  // extern "C"
