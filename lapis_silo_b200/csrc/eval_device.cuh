// Device helpers shared by the kernels that decode stored containers into dense 64 Ki-row tiles: the
// filter-program interpreter (filter_eval.cu) and the co-occurrence kernels (aggregation.cu).
// All of them are called by whole CTAs of EVAL_THREADS threads unless stated otherwise.
#pragma once
#include "common.cuh"

namespace silo {

constexpr int EVAL_THREADS = 1024;
constexpr int EVAL_WARPS = EVAL_THREADS / 32;

__device__ __forceinline__ void orBits32(uint32_t* tile32, uint32_t first, uint32_t last /*inclusive*/) {
   const uint32_t fw = first >> 5;
   const uint32_t lw = last >> 5;
   const uint32_t head = 0xFFFFFFFFu << (first & 31);
   const uint32_t tail = 0xFFFFFFFFu >> (31 - (last & 31));
   if (fw == lw) {
      atomicOr(&tile32[fw], head & tail);
      return;
   }
   atomicOr(&tile32[fw], head);
   for (uint32_t w = fw + 1; w < lw; ++w) {
      atomicOr(&tile32[w], 0xFFFFFFFFu);
   }
   atomicOr(&tile32[lw], tail);
}

// Decoders of the column's device piece formats (common.cuh) for the interpreter. `slot` runs over
// the stored slots of a piece in memory order; the order of the rows does not matter here.
__device__ __forceinline__ uint32_t arrayPieceSlots(uint32_t n) {
   return arrayPieceBytes(n) >> 1;
}
// whether stored u16 slot `slot` of a KIND_ARRAY_T piece of n values is a value (not padding)
__device__ __forceinline__ bool arraySlotValid(uint32_t slot, uint32_t n) {
   const uint32_t region = slot >> 8;  // 256 slots = 512 bytes per region
   const uint32_t count = min(ARRAY_REGION_VALUES, n - region * ARRAY_REGION_VALUES);
   const uint32_t in_region = slot & 255u;  // lane = in_region / 8, j = in_region % 8
   return (in_region >> 3) + arrayRegionLanes(count) * (in_region & 7u) < count;
}
__device__ __forceinline__ uint32_t runsPieceSlots(uint32_t n) {  // u32 slots incl. padding entries
   return runsPieceBytes(n) >> 2;
}

// tile |= piece (whole CTA). `slab` is the payload slab the descriptor's offset4 refers to. Pieces of
// a column are <= 1 KiB; containers of a host bitmap come whole in the CRoaring layouts (arrays up to
// 4096 values, 1024-word bitsets, any number of runs) and take the strided loops.
__device__ inline void orContainerIntoTile(uint64_t* tile, const uint8_t* slab, const DevContainer& desc) {
   const uint8_t* payload = slab + (static_cast<size_t>(desc.offset4) << 2);
   const uint32_t kind = desc.type();
   uint32_t* tile32 = reinterpret_cast<uint32_t*>(tile);
   if (kind == KIND_BITSET) {
      const uint64_t* words = reinterpret_cast<const uint64_t*>(payload);
      const uint32_t first = desc.firstWord();
      for (uint32_t w = threadIdx.x; w < desc.wordCount(); w += EVAL_THREADS) {
         tile[first + w] |= words[w];  // callers separate pieces by __syncthreads, so no other writer
      }
   } else if (kind == KIND_ARRAY_T) {
      const uint16_t* values = reinterpret_cast<const uint16_t*>(payload);
      const uint32_t n = desc.cardinality();
      for (uint32_t slot = threadIdx.x; slot < arrayPieceSlots(n); slot += EVAL_THREADS) {
         if (arraySlotValid(slot, n)) {
            const uint32_t value = values[slot] ^ ARRAY_VALUE_FLIP;
            atomicOr(&tile32[value >> 5], 1u << (value & 31));
         }
      }
   } else if (kind == KIND_RUNS_W) {
      const uint32_t* entries = reinterpret_cast<const uint32_t*>(payload);
      for (uint32_t i = threadIdx.x; i < runsPieceSlots(desc.aux); i += EVAL_THREADS) {
         const uint32_t entry = entries[i];
         if (entry < RUNS_PAD_ENTRY) {
            atomicOr(&tile32[entry >> 20], runEntryMask(entry));
         }
      }
   } else if (kind == KIND_WORDRANGE) {
      const uint32_t* ranges = reinterpret_cast<const uint32_t*>(payload);
      for (uint32_t r = 0; r < desc.aux; ++r) {
         const uint32_t range = ranges[r];
         for (uint32_t w = (range & 0xFFFFu) + threadIdx.x; w < (range >> 16); w += EVAL_THREADS) {
            atomicOr(&tile32[w], 0xFFFFFFFFu);
         }
      }
   } else if (kind == KIND_RAW_ARRAY) {
      const uint16_t* values = reinterpret_cast<const uint16_t*>(payload);
      const uint32_t cardinality = desc.cardinality();
      for (uint32_t i = threadIdx.x; i < cardinality; i += EVAL_THREADS) {
         const uint32_t value = values[i];
         atomicOr(&tile32[value >> 5], 1u << (value & 31));
      }
   } else if (kind == KIND_RAW_RUN) {
      const uint32_t* runs = reinterpret_cast<const uint32_t*>(payload);
      for (uint32_t i = threadIdx.x; i < desc.aux; i += EVAL_THREADS) {
         const uint32_t run = runs[i];
         const uint32_t first = run & 0xFFFFu;
         orBits32(tile32, first, first + (run >> 16));
      }
   } else if (threadIdx.x < desc.cardinality()) {  // KIND_INLINE
      const uint32_t value = (desc.aux >> (16 * threadIdx.x)) & 0xFFFFu;
      atomicOr(&tile32[value >> 5], 1u << (value & 31));
   }
}

// counters[row] += delta for every row of the piece (one warp). delta is +1 or -1 applied to a
// u16 lane of a packed u32; the host-chosen bias keeps every lane inside [0, 65535].
__device__ inline void addContainerToCounters(
   uint32_t* counters32,
   const uint8_t* slab,
   const DevContainer& desc,
   bool subtract,
   uint32_t lane
) {
   const uint8_t* payload = slab + (static_cast<size_t>(desc.offset4) << 2);
   const uint32_t kind = desc.type();
   auto bump = [&](uint32_t row) {
      const uint32_t unit = 1u << ((row & 1u) << 4);
      atomicAdd(&counters32[row >> 1], subtract ? 0u - unit : unit);
   };
   if (kind == KIND_ARRAY_T) {
      // one 128-bit load per lane and region, like the container kernel (common.cuh: lane L of a region
      // with `count` values in P lanes holds the values L + P*j)
      const uint32_t n = desc.cardinality();
      for (uint32_t first = 0; first < n; first += ARRAY_REGION_VALUES) {
         const uint32_t count = min(ARRAY_REGION_VALUES, n - first);
         const uint32_t lanes = arrayRegionLanes(count);
         if (lane < lanes) {
            const uint4 eight = reinterpret_cast<const uint4*>(payload + (first / ARRAY_REGION_VALUES) * 512u)[lane];
            const uint32_t words[4] = {eight.x, eight.y, eight.z, eight.w};
#pragma unroll
            for (uint32_t j = 0; j < 8; ++j) {
               if (lane + lanes * j < count) {
                  bump(((words[j >> 1] >> (16 * (j & 1))) & 0xFFFFu) ^ ARRAY_VALUE_FLIP);
               }
            }
         }
      }
   } else if (kind == KIND_RUNS_W) {
      const uint32_t n = desc.aux;
      for (uint32_t first = 0; first < n; first += RUNS_REGION_ENTRIES) {
         const uint32_t count = min(RUNS_REGION_ENTRIES, n - first);
         if (lane < runsRegionLanes(count)) {
            const uint4 four = reinterpret_cast<const uint4*>(payload + (first / RUNS_REGION_ENTRIES) * 512u)[lane];
            const uint32_t entries[4] = {four.x, four.y, four.z, four.w};
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j) {
               if (entries[j] < RUNS_PAD_ENTRY) {  // a run inside one word: consecutive rows, no bit scan needed
                  const uint32_t first_row = (entries[j] >> 20) * 32u + (entries[j] & 31u);
                  const uint32_t length = 32u - ((entries[j] >> 5) & 31u);
                  for (uint32_t row = first_row; row < first_row + length; ++row) {
                     bump(row);
                  }
               }
            }
         }
      }
   } else if (kind == KIND_WORDRANGE) {
      const uint32_t* ranges = reinterpret_cast<const uint32_t*>(payload);
      for (uint32_t r = 0; r < desc.aux; ++r) {
         const uint32_t range = ranges[r];
         for (uint32_t row = (range & 0xFFFFu) * 32u + lane; row < (range >> 16) * 32u; row += 32) {
            bump(row);
         }
      }
   } else if (kind == KIND_BITSET) {
      const uint64_t* words = reinterpret_cast<const uint64_t*>(payload);
      const uint32_t first = desc.firstWord();
      for (uint32_t w = lane; w < desc.wordCount(); w += 32) {
         uint64_t word = words[w];
         while (word != 0) {
            bump((first + w) * 64 + static_cast<uint32_t>(__ffsll(static_cast<long long>(word)) - 1));
            word &= word - 1;
         }
      }
   } else if (kind == KIND_RAW_ARRAY) {
      const uint16_t* values = reinterpret_cast<const uint16_t*>(payload);
      const uint32_t cardinality = desc.cardinality();
      for (uint32_t i = lane; i < cardinality; i += 32) {
         bump(values[i]);
      }
   } else if (kind == KIND_RAW_RUN) {
      const uint32_t* runs = reinterpret_cast<const uint32_t*>(payload);
      for (uint32_t r = 0; r < desc.aux; ++r) {
         const uint32_t run = runs[r];
         const uint32_t first = run & 0xFFFFu;
         const uint32_t last = first + (run >> 16);
         for (uint32_t row = first + lane; row <= last; row += 32) {
            bump(row);
         }
      }
   } else if (lane < desc.cardinality()) {  // KIND_INLINE
      bump((desc.aux >> (16 * lane)) & 0xFFFFu);
   }
}

// counters[row] +/-= 1 for every row of a column piece whose (at most) two 512-byte payload regions one warp holds
// in registers (lane L: bytes [16 L, 16 L + 16) of each region), the way the container kernel pulls them; WORDRANGE
// pieces are read from `payload`. Used by the threshold sweep, which keeps the loads of the next piece in flight
// while it bumps the current one.
__device__ __forceinline__ void bumpPieceFromRegisters(
   uint32_t* counters32,
   const DevContainer& desc,
   const uint4& first,
   const uint4& second,
   const uint8_t* payload,
   bool subtract,
   uint32_t lane
) {
   const uint32_t kind = desc.type();
   const uint32_t low_unit = subtract ? 0xFFFFFFFFu : 1u;          // -1 / +1 on the low u16 lane (no borrow: the lanes are biased)
   const uint32_t high_unit = subtract ? 0xFFFF0000u : 0x00010000u;
   auto bump = [&](uint32_t row) { atomicAdd(&counters32[row >> 1], (row & 1u) != 0 ? high_unit : low_unit); };
   // rows 2t and 2t + 1 of a 32-row word as one packed add
   auto bumpPairs = [&](uint32_t word_index, uint32_t mask) {
      while (mask != 0) {
         const uint32_t t = static_cast<uint32_t>(__ffs(static_cast<int>(mask)) - 1) >> 1;
         const uint32_t bits = (mask >> (2 * t)) & 3u;
         atomicAdd(&counters32[word_index * 16 + t], ((bits & 1u) != 0 ? low_unit : 0u) + ((bits & 2u) != 0 ? high_unit : 0u));
         mask &= ~(3u << (2 * t));
      }
   };
   // consecutive rows [first, first + length) inside one 32-row word: whole pairs with one packed add each
   auto bumpRun = [&](uint32_t entry) {
      const uint32_t first = entry & 31u;
      const uint32_t last = first + (32u - ((entry >> 5) & 31u)) - 1u;
      uint32_t* const pairs = counters32 + (entry >> 20) * 16;
      uint32_t t = first >> 1;
      const uint32_t t_last = last >> 1;
      if ((first & 1u) != 0) {  // the run starts on the high lane of its pair
         atomicAdd(&pairs[t], high_unit);
         ++t;
      }
      const bool low_lane_end = (last & 1u) == 0;  // ... and ends on the low lane of its last pair
      const uint32_t full_end = low_lane_end ? t_last : t_last + 1;
      for (; t < full_end; ++t) {
         atomicAdd(&pairs[t], low_unit + high_unit);
      }
      if (low_lane_end) {
         atomicAdd(&pairs[t_last], low_unit);
      }
   };
   if (kind == KIND_ARRAY_T) {
      const uint32_t n = desc.cardinality();
#pragma unroll
      for (uint32_t region = 0; region < 2; ++region) {
         if (region * ARRAY_REGION_VALUES < n) {
            const uint32_t count = min(ARRAY_REGION_VALUES, n - region * ARRAY_REGION_VALUES);
            const uint32_t lanes = arrayRegionLanes(count);
            const uint4& eight = region == 0 ? first : second;
            const uint32_t words[4] = {eight.x, eight.y, eight.z, eight.w};
#pragma unroll
            for (uint32_t j = 0; j < 8; ++j) {
               if (lane < lanes && lane + lanes * j < count) {
                  bump(((words[j >> 1] >> (16 * (j & 1))) & 0xFFFFu) ^ ARRAY_VALUE_FLIP);
               }
            }
         }
      }
   } else if (kind == KIND_RUNS_W) {
      const uint32_t n = desc.aux;
#pragma unroll
      for (uint32_t region = 0; region < 2; ++region) {
         if (region * RUNS_REGION_ENTRIES < n) {
            const uint32_t count = min(RUNS_REGION_ENTRIES, n - region * RUNS_REGION_ENTRIES);
            const uint4& four = region == 0 ? first : second;
            const uint32_t entries[4] = {four.x, four.y, four.z, four.w};
            if (lane < runsRegionLanes(count)) {
#pragma unroll
               for (uint32_t j = 0; j < 4; ++j) {
                  if (entries[j] < RUNS_PAD_ENTRY) {
                     bumpRun(entries[j]);
                  }
               }
            }
         }
      }
   } else if (kind == KIND_BITSET) {  // 128 u64 words: lane L holds words 2L, 2L+1 of each half
      const uint32_t first_word32 = desc.firstWord() * 2;
      bumpPairs(first_word32 + 4 * lane + 0, first.x);
      bumpPairs(first_word32 + 4 * lane + 1, first.y);
      bumpPairs(first_word32 + 4 * lane + 2, first.z);
      bumpPairs(first_word32 + 4 * lane + 3, first.w);
      bumpPairs(first_word32 + 128 + 4 * lane + 0, second.x);
      bumpPairs(first_word32 + 128 + 4 * lane + 1, second.y);
      bumpPairs(first_word32 + 128 + 4 * lane + 2, second.z);
      bumpPairs(first_word32 + 128 + 4 * lane + 3, second.w);
   } else if (kind == KIND_WORDRANGE) {
      const uint32_t* ranges = reinterpret_cast<const uint32_t*>(payload);
      for (uint32_t r = 0; r < desc.aux; ++r) {
         const uint32_t range = ranges[r];
         for (uint32_t pair = (range & 0xFFFFu) * 16u + lane; pair < (range >> 16) * 16u; pair += 32) {
            atomicAdd(&counters32[pair], low_unit + high_unit);
         }
      }
   } else if (lane < desc.cardinality()) {  // KIND_INLINE
      bump((desc.aux >> (16 * lane)) & 0xFFFFu);
   }
}

// First index in [lo, hi) whose key is >= target, searched by ONE converged warp with 32 probes per
// round (a 33-ary search: a chunk's ~4k descriptors take 3 rounds of global-memory latency instead
// of 12). The result is warp-uniform.
template <typename KeyAt>
__device__ __forceinline__ uint32_t warpLowerBound(uint32_t lo, uint32_t hi, uint32_t target, uint32_t lane, KeyAt key_at) {
   while (hi - lo > 32) {
      const uint32_t span = hi - lo;
      const uint32_t probe = lo + static_cast<uint32_t>(static_cast<uint64_t>(lane + 1) * span / 33);  // in (lo, hi)
      const uint32_t n_less = __popc(__ballot_sync(0xFFFFFFFFu, key_at(probe) < target));  // monotone in the lane
      const uint32_t below = __shfl_sync(0xFFFFFFFFu, probe, n_less == 0 ? 0 : n_less - 1);
      const uint32_t above = __shfl_sync(0xFFFFFFFFu, probe, n_less == 32 ? 31 : n_less);
      lo = n_less == 0 ? lo : below + 1;
      hi = n_less == 32 ? hi : above;
   }
   const uint32_t index = lo + lane;
   const bool less = index < hi && key_at(index) < target;
   return lo + __popc(__ballot_sync(0xFFFFFFFFu, less));
}

// [lo, hi) = descriptors of `chunk` at `position` (warp 0 searches, result broadcast via smem)
__device__ inline void findPositionRange(uint32_t* range, const DevColumn& column, uint32_t chunk, uint32_t position) {
   if (threadIdx.x < 32) {
      const uint32_t begin = column.chunk_desc_begin[chunk];
      const uint32_t end = column.chunk_desc_begin[chunk + 1];
      auto key_at = [&](uint32_t index) { return column.containers[index].position; };
      const uint32_t lo = warpLowerBound(begin, end, position, threadIdx.x, key_at);
      const uint32_t stop = warpLowerBound(lo, end, position + 1, threadIdx.x, key_at);
      if (threadIdx.x == 0) {
         range[0] = lo;
         range[1] = stop;
      }
   }
   __syncthreads();
}

__device__ __forceinline__ uint32_t lowerBound(const uint32_t* sorted, uint32_t count, uint32_t value) {
   uint32_t lo = 0;
   uint32_t hi = count;
   while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (sorted[mid] < value) {
         lo = mid + 1;
      } else {
         hi = mid;
      }
   }
   return lo;
}

__device__ __forceinline__ bool rowMissingAt(const DevColumn& column, uint32_t missing_index, uint32_t position) {
   uint64_t lo = column.missing_offsets[missing_index];
   uint64_t hi = column.missing_offsets[missing_index + 1];
   while (lo < hi) {  // first run with end_exclusive > position
      const uint64_t mid = (lo + hi) >> 1;
      if (column.missing_runs[mid].y <= position) {
         lo = mid + 1;
      } else {
         hi = mid;
      }
   }
   return lo < column.missing_offsets[missing_index + 1] && column.missing_runs[lo].x <= position;
}

}  // namespace silo
