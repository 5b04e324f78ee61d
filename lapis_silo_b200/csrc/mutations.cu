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
#include <cstring>

#include "common.cuh"

namespace silo {

namespace {

// ---------------------------------------------------------------------------------------------
// work list: segments of the chunks that hold at least one filtered row
// ---------------------------------------------------------------------------------------------

__global__ void buildWorkListKernel(
   const uint32_t* __restrict__ chunk_popcount,
   const uint32_t* __restrict__ chunk_seg_begin,
   uint32_t n_chunks,
   uint32_t* __restrict__ work_prefix  // [n_chunks + 1]; work_prefix[n_chunks] = total
) {
   __shared__ uint32_t warp_totals[32];
   __shared__ uint32_t carry;
   if (threadIdx.x == 0) {
      carry = 0;
   }
   __syncthreads();
   for (uint32_t base = 0; base < n_chunks; base += blockDim.x) {
      const uint32_t chunk = base + threadIdx.x;
      uint32_t value = 0;
      if (chunk < n_chunks && chunk_popcount[chunk] != 0) {
         value = chunk_seg_begin[chunk + 1] - chunk_seg_begin[chunk];
      }
      uint32_t inclusive = value;
      for (int offset = 1; offset < 32; offset <<= 1) {
         const uint32_t other = __shfl_up_sync(0xFFFFFFFFu, inclusive, offset);
         if ((threadIdx.x & 31) >= offset) {
            inclusive += other;
         }
      }
      if ((threadIdx.x & 31) == 31) {
         warp_totals[threadIdx.x >> 5] = inclusive;
      }
      __syncthreads();
      uint32_t warp_offset = 0;
      for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) {
         warp_offset += warp_totals[w];
      }
      const uint32_t block_carry = carry;
      if (chunk < n_chunks) {
         work_prefix[chunk] = block_carry + warp_offset + inclusive - value;
      }
      __syncthreads();
      if (threadIdx.x == blockDim.x - 1) {
         carry = block_carry + warp_offset + inclusive;
      }
      __syncthreads();
   }
   if (threadIdx.x == 0) {
      work_prefix[n_chunks] = carry;
   }
}

// ---------------------------------------------------------------------------------------------
// K1: fused container AND filter-tile + popcount
// ---------------------------------------------------------------------------------------------

constexpr int K1_STAGES = 4;
constexpr int K1_CONSUMER_WARPS = 16;
constexpr int K1_THREADS = (K1_CONSUMER_WARPS + 1) * 32;  // warp 0 = bulk-copy producer

struct __align__(16) K1Stage {
   uint8_t payload[SEG_PAYLOAD_BYTES];
   DevContainer descs[SEG_MAX_DESCS];
};

struct __align__(16) K1Shared {
   uint64_t tile[TILE_WORDS];       // dense filter tile of the current chunk
   uint16_t rank[TILE_WORDS];       // exclusive popcount prefix per 64-bit word (run containers)
   uint32_t warp_sums[K1_CONSUMER_WARPS];
   K1Stage stages[K1_STAGES];
   uint64_t full_bar[K1_STAGES];
   uint64_t empty_bar[K1_STAGES];
   // per-stage meta written by the producer before it arms full_bar
   uint32_t meta_desc_count[K1_STAGES];
   uint32_t meta_base4[K1_STAGES];  // slab offset (4-byte units) of the stage's payload[0]
   uint32_t meta_new_tile[K1_STAGES];
   uint32_t next_container[K1_STAGES];  // dynamic container scheduling inside a stage
};

__device__ __forceinline__ uint32_t warpSum(uint32_t value) {
   return __reduce_add_sync(0xFFFFFFFFu, value);
}

// |container AND tile| for one container whose payload sits in shared memory; whole warp cooperates.
__device__ __forceinline__ uint32_t andCardinality(
   const DevContainer& desc,
   const uint8_t* payload,  // shared
   const uint64_t* tile,    // shared
   const uint16_t* rank,    // shared
   uint32_t lane
) {
   const uint32_t type = desc.type();
   uint32_t local = 0;
   if (type == TYPE_ARRAY) {
      const uint32_t cardinality = desc.cardinality();
      const uint32_t* tile32 = reinterpret_cast<const uint32_t*>(tile);
      const uint32_t* pairs = reinterpret_cast<const uint32_t*>(payload);
      const uint32_t n_pairs = cardinality >> 1;
      for (uint32_t i = lane; i < n_pairs; i += 32) {
         const uint32_t two = pairs[i];
         const uint32_t lo = two & 0xFFFFu;
         const uint32_t hi = two >> 16;
         local += (tile32[lo >> 5] >> (lo & 31)) & 1u;
         local += (tile32[hi >> 5] >> (hi & 31)) & 1u;
      }
      if ((cardinality & 1u) != 0 && lane == 0) {
         const uint32_t last = reinterpret_cast<const uint16_t*>(payload)[cardinality - 1];
         local += (tile32[last >> 5] >> (last & 31)) & 1u;
      }
   } else if (type == TYPE_RUN) {
      const uint32_t* runs = reinterpret_cast<const uint32_t*>(payload);
      for (uint32_t i = lane; i < desc.n_runs; i += 32) {
         const uint32_t run = runs[i];
         const uint32_t first = run & 0xFFFFu;
         const uint32_t last = first + (run >> 16);  // inclusive
         const uint32_t fw = first >> 6;
         const uint32_t lw = last >> 6;
         const uint64_t head = ~0ULL << (first & 63);
         const uint64_t tail = ~0ULL >> (63 - (last & 63));
         if (fw == lw) {
            local += __popcll(tile[fw] & head & tail);
         } else {
            // rows of the words strictly between come from the rank table
            local += __popcll(tile[fw] & head) + __popcll(tile[lw] & tail) +
                     (static_cast<uint32_t>(rank[lw]) - static_cast<uint32_t>(rank[fw]) -
                      static_cast<uint32_t>(__popcll(tile[fw])));
         }
      }
   } else {
      const uint4* words = reinterpret_cast<const uint4*>(payload);
      const uint4* tile4 = reinterpret_cast<const uint4*>(tile);
#pragma unroll 4
      for (uint32_t i = lane; i < TILE_WORDS / 2; i += 32) {
         const uint4 a = words[i];
         const uint4 b = tile4[i];
         local += __popc(a.x & b.x) + __popc(a.y & b.y) + __popc(a.z & b.z) + __popc(a.w & b.w);
      }
   }
   return warpSum(local);
}

__global__ void __launch_bounds__(K1_THREADS, 2) containerAndCountKernel(
   DevColumn column,
   const uint64_t* __restrict__ filter_words,   // [n_chunks * 1024]
   const uint32_t* __restrict__ work_prefix,    // [n_chunks + 1]
   uint32_t* __restrict__ counts                // [n_symbols * genome_length]
) {
   extern __shared__ __align__(128) uint8_t smem_raw[];
   K1Shared& sh = *reinterpret_cast<K1Shared*>(smem_raw);

   const uint32_t warp = threadIdx.x >> 5;
   const uint32_t lane = threadIdx.x & 31;
   const uint32_t total = work_prefix[column.n_chunks];
   // contiguous, balanced slice of the work list for this CTA
   const uint32_t begin = static_cast<uint32_t>(static_cast<uint64_t>(total) * blockIdx.x / gridDim.x);
   const uint32_t end = static_cast<uint32_t>(static_cast<uint64_t>(total) * (blockIdx.x + 1) / gridDim.x);
   if (begin >= end) {
      return;
   }

   if (threadIdx.x == 0) {
      for (int s = 0; s < K1_STAGES; ++s) {
         mbarInit(&sh.full_bar[s], 1);
         mbarInit(&sh.empty_bar[s], K1_CONSUMER_WARPS);
      }
      fenceBarrierInit();
   }
   __syncthreads();

   if (warp == 0) {
      // ---------------- producer: one elected lane streams segments with bulk copies --------------
      if (lane == 0) {
         uint32_t chunk = 0;
         // locate the chunk of the first work item
         {
            uint32_t lo = 0;
            uint32_t hi = column.n_chunks;  // work_prefix[lo] <= begin < work_prefix[hi]
            while (hi - lo > 1) {
               const uint32_t mid = (lo + hi) >> 1;
               if (work_prefix[mid] <= begin) {
                  lo = mid;
               } else {
                  hi = mid;
               }
            }
            chunk = lo;
         }
         uint32_t current_tile_chunk = 0xFFFFFFFFu;
         for (uint32_t item = begin; item < end; ++item) {
            while (work_prefix[chunk + 1] <= item) {
               ++chunk;
            }
            const uint32_t it = item - begin;
            const uint32_t stage = it % K1_STAGES;
            const uint32_t round = it / K1_STAGES;
            const bool new_tile = chunk != current_tile_chunk;
            if (new_tile) {
               // the single filter tile is shared by all stages: drain the pipeline before replacing it
               for (uint32_t back = 1; back < K1_STAGES && back <= it; ++back) {
                  const uint32_t prev = it - back;
                  mbarWait(&sh.empty_bar[prev % K1_STAGES], (prev / K1_STAGES) & 1u);
               }
               current_tile_chunk = chunk;
            }
            if (round > 0) {
               mbarWait(&sh.empty_bar[stage], (round - 1) & 1u);
            }
            const DevSegment segment = column.segments[column.chunk_seg_begin[chunk] + (item - work_prefix[chunk])];
            sh.meta_desc_count[stage] = segment.desc_count;
            sh.meta_base4[stage] = static_cast<uint32_t>(segment.payload_offset >> 2);
            sh.meta_new_tile[stage] = new_tile ? 1u : 0u;
            sh.next_container[stage] = 0;
            const uint32_t desc_bytes = segment.desc_count * static_cast<uint32_t>(sizeof(DevContainer));
            mbarExpectTx(
               &sh.full_bar[stage], desc_bytes + segment.payload_bytes + (new_tile ? TILE_BYTES : 0u)
            );
            if (new_tile) {
               bulkLoad(sh.tile, filter_words + static_cast<size_t>(chunk) * TILE_WORDS, TILE_BYTES, &sh.full_bar[stage]);
            }
            bulkLoad(sh.stages[stage].descs, column.containers + segment.desc_begin, desc_bytes, &sh.full_bar[stage]);
            bulkLoad(sh.stages[stage].payload, column.payload + segment.payload_offset, segment.payload_bytes, &sh.full_bar[stage]);
         }
      }
      return;
   }

   // ---------------- consumers: 16 warps, one container per warp at a time ----------------------
   const uint32_t cwarp = warp - 1;
   const uint32_t cthread = threadIdx.x - 32;  // 0..511
   const uint32_t genome_length = column.genome_length;
   for (uint32_t item = begin; item < end; ++item) {
      const uint32_t it = item - begin;
      const uint32_t stage = it % K1_STAGES;
      mbarWait(&sh.full_bar[stage], (it / K1_STAGES) & 1u);
      if (sh.meta_new_tile[stage] != 0) {
         // rebuild the per-word exclusive rank table of the freshly loaded tile (512 threads x 2 words)
         const uint32_t p0 = __popcll(sh.tile[2 * cthread]);
         const uint32_t p1 = __popcll(sh.tile[2 * cthread + 1]);
         uint32_t inclusive = p0 + p1;
         for (int offset = 1; offset < 32; offset <<= 1) {
            const uint32_t other = __shfl_up_sync(0xFFFFFFFFu, inclusive, offset);
            if (lane >= static_cast<uint32_t>(offset)) {
               inclusive += other;
            }
         }
         if (lane == 31) {
            sh.warp_sums[cwarp] = inclusive;
         }
         asm volatile("bar.sync 1, %0;" ::"n"(K1_CONSUMER_WARPS * 32) : "memory");
         uint32_t warp_offset = 0;
         for (uint32_t w = 0; w < cwarp; ++w) {
            warp_offset += sh.warp_sums[w];
         }
         const uint32_t exclusive = warp_offset + inclusive - (p0 + p1);
         sh.rank[2 * cthread] = static_cast<uint16_t>(exclusive);
         sh.rank[2 * cthread + 1] = static_cast<uint16_t>(exclusive + p0);
         asm volatile("bar.sync 1, %0;" ::"n"(K1_CONSUMER_WARPS * 32) : "memory");
      }
      const uint32_t desc_count = sh.meta_desc_count[stage];
      const uint32_t base4 = sh.meta_base4[stage];
      const K1Stage& st = sh.stages[stage];
      while (true) {
         uint32_t index = 0;
         if (lane == 0) {
            index = atomicAdd(&sh.next_container[stage], 1u);
         }
         index = __shfl_sync(0xFFFFFFFFu, index, 0);
         if (index >= desc_count) {
            break;
         }
         const DevContainer desc = st.descs[index];
         const uint8_t* payload = st.payload + (static_cast<size_t>(desc.offset4 - base4) << 2);
         const uint32_t count = andCardinality(desc, payload, sh.tile, sh.rank, lane);
         if (lane == 0 && count != 0) {
            atomicAdd(&counts[desc.symbol() * genome_length + desc.position], count);
         }
      }
      __syncwarp();
      if (lane == 0) {
         mbarArrive(&sh.empty_bar[stage]);
      }
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
constexpr int K6_SLICES = 4;  // CTAs per chunk

struct PendingAdd {
   uint32_t key;
   uint32_t count;
};

__device__ __forceinline__ void flushPending(PendingAdd& pending, uint32_t* diff, uint32_t lane, bool negate) {
   if (pending.count != 0 && lane == 0) {
      atomicAdd(&diff[pending.key], negate ? 0u - pending.count : pending.count);
   }
   pending.count = 0;
}

// adds `+1` (or -1) at diff[key] for every active lane, aggregating equal keys: a warp-uniform key
// is accumulated in registers across iterations (the common case: full-length genomes, sorted reads)
__device__ __forceinline__ void aggregateAdd(
   PendingAdd& pending,
   uint32_t* diff,
   uint32_t key,
   bool active,
   uint32_t lane,
   bool negate
) {
   const uint32_t active_mask = __ballot_sync(0xFFFFFFFFu, active);
   if (active_mask == 0) {
      return;
   }
   const uint32_t leader = __ffs(active_mask) - 1;
   const uint32_t leader_key = __shfl_sync(0xFFFFFFFFu, key, leader);
   const uint32_t same_mask = __ballot_sync(0xFFFFFFFFu, active && key == leader_key);
   if (same_mask == active_mask) {
      if (pending.count != 0 && pending.key != leader_key) {
         flushPending(pending, diff, lane, negate);
      }
      pending.key = leader_key;
      pending.count += __popc(active_mask);
      return;
   }
   // mixed keys in this warp iteration: one RED per distinct key
   const uint32_t peers = __match_any_sync(0xFFFFFFFFu, active ? key : 0xFFFFFFFFu);
   if (active && lane == static_cast<uint32_t>(__ffs(peers) - 1)) {
      const uint32_t amount = __popc(peers);
      atomicAdd(&diff[key], negate ? 0u - amount : amount);
   }
}

__global__ void __launch_bounds__(K6_THREADS) coverageDiffKernel(
   DevColumn column,
   const uint64_t* __restrict__ filter_words,
   const uint32_t* __restrict__ chunk_popcount,
   const uint32_t* __restrict__ chunk_sizes,
   uint32_t* __restrict__ diff  // [genome_length + 1]
) {
   const uint32_t chunk = blockIdx.x / K6_SLICES;
   const uint32_t slice = blockIdx.x % K6_SLICES;
   if (chunk_popcount[chunk] == 0) {
      return;
   }
   const uint32_t lane = threadIdx.x & 31;
   const uint32_t warp = threadIdx.x >> 5;
   const uint32_t chunk_size = chunk_sizes[chunk];
   const uint32_t* tile32 = reinterpret_cast<const uint32_t*>(filter_words + static_cast<size_t>(chunk) * TILE_WORDS);
   const uint2* rows = column.start_end + column.chunk_row_begin[chunk];

   constexpr uint32_t ROWS_PER_SLICE = 65536 / K6_SLICES;
   constexpr uint32_t WARPS = K6_THREADS / 32;
   constexpr uint32_t ROWS_PER_WARP = ROWS_PER_SLICE / WARPS;
   const uint32_t warp_first = slice * ROWS_PER_SLICE + warp * ROWS_PER_WARP;
   PendingAdd pending_start{0, 0};
   PendingAdd pending_end{0, 0};
   for (uint32_t base = warp_first; base < warp_first + ROWS_PER_WARP && base < chunk_size; base += 32) {
      const uint32_t bits = tile32[base >> 5];
      if (bits == 0) {
         continue;
      }
      const uint32_t row = base + lane;
      const bool active = ((bits >> lane) & 1u) != 0 && row < chunk_size;
      uint2 range = make_uint2(0, 0);
      if (active) {
         range = rows[row];
      }
      aggregateAdd(pending_start, diff, range.x, active, lane, false);
      aggregateAdd(pending_end, diff, range.y, active, lane, true);
   }
   flushPending(pending_start, diff, lane, false);
   flushPending(pending_end, diff, lane, true);

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
         atomicAdd(&diff[r.x], 0u - 1u);
         atomicAdd(&diff[r.y], 1u);
      }
   }
}

// ---------------------------------------------------------------------------------------------
// finalize: covered(p) = prefix sum of diff; counts[local_ref[p]][p] = covered(p) - sum(others)
// ---------------------------------------------------------------------------------------------

constexpr int FIN_THREADS = 1024;

__global__ void __launch_bounds__(FIN_THREADS) finalizeCountsKernel(
   DevColumn column,
   const uint32_t* __restrict__ diff,
   uint32_t* __restrict__ counts
) {
   __shared__ uint32_t warp_totals[32];
   const uint32_t genome_length = column.genome_length;
   const uint32_t per_thread = (genome_length + FIN_THREADS - 1) / FIN_THREADS;
   const uint32_t first = threadIdx.x * per_thread;
   const uint32_t last = min(first + per_thread, genome_length);
   uint32_t local = 0;
   for (uint32_t p = first; p < last; ++p) {
      local += diff[p];
   }
   uint32_t inclusive = local;
   const uint32_t lane = threadIdx.x & 31;
   for (int offset = 1; offset < 32; offset <<= 1) {
      const uint32_t other = __shfl_up_sync(0xFFFFFFFFu, inclusive, offset);
      if (lane >= static_cast<uint32_t>(offset)) {
         inclusive += other;
      }
   }
   if (lane == 31) {
      warp_totals[threadIdx.x >> 5] = inclusive;
   }
   __syncthreads();
   uint32_t running = inclusive - local;
   for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) {
      running += warp_totals[w];
   }
   for (uint32_t p = first; p < last; ++p) {
      running += diff[p];  // rows of the filter that cover position p
      const uint32_t reference_symbol = column.local_reference[p];
      uint32_t others = 0;
      for (uint32_t symbol = 0; symbol < column.n_symbols; ++symbol) {
         if (symbol != reference_symbol) {
            others += counts[symbol * genome_length + p];
         }
      }
      counts[reference_symbol * genome_length + p] = running - others;
   }
}

// ---------------------------------------------------------------------------------------------

void enqueueMutationCounts(
   silo_gpu_table* table,
   int column_index,
   const silo_gpu_filter* filter,
   uint32_t* d_counts,
   cudaStream_t stream
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
   const uint64_t slot = table->timed_calls % silo_gpu_table::EVENT_RING;
   table->timed_calls++;
   cudaEvent_t ev_begin = table->ev_begin[slot];
   cudaEvent_t ev_k1_begin = table->ev_k1_begin[slot];
   cudaEvent_t ev_k1_end = table->ev_k1_end[slot];
   cudaEvent_t ev_end = table->ev_end[slot];
   SILO_CUDA_CHECK(cudaEventRecord(ev_begin, stream));
   SILO_CUDA_CHECK(cudaMemsetAsync(d_counts, 0, counts_bytes, stream));
   SILO_CUDA_CHECK(cudaMemsetAsync(table->d_coverage_diff, 0, (column.genome_length + 1) * sizeof(uint32_t), stream));
   if (n_chunks == 0) {
      SILO_CUDA_CHECK(cudaEventRecord(ev_k1_begin, stream));
      SILO_CUDA_CHECK(cudaEventRecord(ev_k1_end, stream));
      SILO_CUDA_CHECK(cudaEventRecord(ev_end, stream));
      return;
   }
   const uint64_t* words = filter != nullptr ? filter->d_words : table->d_full_words;
   const uint32_t* popcounts = filter != nullptr ? filter->d_chunk_popcount : table->d_chunk_popcount_full;

   if (filter == nullptr) {
      SILO_CUDA_CHECK(cudaEventRecord(ev_k1_begin, stream));
      if (column.n_containers > 0) {
         const int blocks = static_cast<int>(std::min<uint64_t>((column.n_containers + 255) / 256, static_cast<uint64_t>(table->ctx->sm_count) * 8));
         containerCardinalityKernel<<<blocks, 256, 0, stream>>>(column, d_counts);
         SILO_CUDA_CHECK(cudaGetLastError());
         table->stats.kernel_launches++;
      }
      SILO_CUDA_CHECK(cudaEventRecord(ev_k1_end, stream));
   } else {
      buildWorkListKernel<<<1, 1024, 0, stream>>>(popcounts, column.chunk_seg_begin, n_chunks, table->d_work_prefix);
      SILO_CUDA_CHECK(cudaGetLastError());
      table->stats.kernel_launches++;
      SILO_CUDA_CHECK(cudaEventRecord(ev_k1_begin, stream));
      if (column.n_segments > 0) {
         static bool attribute_set = false;
         if (!attribute_set) {
            SILO_CUDA_CHECK(cudaFuncSetAttribute(
               containerAndCountKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(K1Shared))
            ));
            attribute_set = true;
         }
         const int blocks = static_cast<int>(std::min<uint32_t>(column.n_segments, static_cast<uint32_t>(table->ctx->sm_count) * 2));
         containerAndCountKernel<<<blocks, K1_THREADS, sizeof(K1Shared), stream>>>(column, words, table->d_work_prefix, d_counts);
         SILO_CUDA_CHECK(cudaGetLastError());
         table->stats.kernel_launches++;
      }
      SILO_CUDA_CHECK(cudaEventRecord(ev_k1_end, stream));
   }
   coverageDiffKernel<<<n_chunks * K6_SLICES, K6_THREADS, 0, stream>>>(
      column, words, popcounts, table->d_chunk_sizes, table->d_coverage_diff
   );
   SILO_CUDA_CHECK(cudaGetLastError());
   finalizeCountsKernel<<<1, FIN_THREADS, 0, stream>>>(column, table->d_coverage_diff, d_counts);
   SILO_CUDA_CHECK(cudaGetLastError());
   table->stats.kernel_launches += 2;
   SILO_CUDA_CHECK(cudaEventRecord(ev_end, stream));
}

}  // namespace

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
      enqueueMutationCounts(table, column, filter, static_cast<uint32_t*>(d_counts), stream);
   });
}

int silo_gpu_mutation_counts(
   silo_gpu_table* table,
   int column,
   const silo_gpu_filter* filter,
   uint32_t* counts
) {
   return guarded([&] {
      require(table != nullptr && counts != nullptr, "mutation_counts: NULL argument");
      std::lock_guard<std::mutex> lock(table->mutex);
      SILO_CUDA_CHECK(cudaSetDevice(table->ctx->device));
      cudaStream_t stream = table->ctx->stream;
      enqueueMutationCounts(table, column, filter, table->d_counts, stream);
      const HostColumn& host = *table->columns[static_cast<size_t>(column)];
      const size_t counts_bytes = static_cast<size_t>(host.dev.n_symbols) * host.dev.genome_length * sizeof(uint32_t);
      SILO_CUDA_CHECK(cudaMemcpyAsync(table->h_counts_pinned, table->d_counts, counts_bytes, cudaMemcpyDeviceToHost, stream));
      SILO_CUDA_CHECK(cudaStreamSynchronize(stream));
      std::memcpy(counts, table->h_counts_pinned, counts_bytes);
   });
}

}  // extern "C"
