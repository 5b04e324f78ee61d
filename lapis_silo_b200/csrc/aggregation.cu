// BitmapAggregationNode (mutation co-occurrence / groupBy over sequence positions and indexed
// columns) on device.
//
// Replaces /root/reference/src/rhydb/query_engine/operators/bitmap_aggregation_node.cpp:
//   buildSymbolBitmaps :53-92       one SymbolInSet bitmap per symbol of the alphabet + the null group,
//                                   each ANDed with the filter
//   IndexedColumnDimension::buildGroups :224-249   one bitmap per dictionary value + the null group
//   partition / computeCombinations :99-139        recursive AND of the group bitmaps, depth first,
//                                   one (group index per dimension, cardinality) per non-empty leaf
//
// The groups of one dimension are disjoint (a row carries exactly one symbol at a position, or its
// sequence is null; a row has one dictionary value, or null), so the recursion computes a GROUP BY
// over per-row tuples. Instead of |alphabet| x dimensions whole-table bitmaps and their pairwise
// intersections, the device
//   1. decodes, per (chunk, dimension), the one-byte group CODE of every row into shared memory and
//      writes it out as a byte plane (positionCodesKernel / bitmapCodesKernel): the stored containers
//      at that position say which rows carry a non-reference symbol, the coverage index says which rows
//      carry the reference symbol and which the missing symbol (symbol_in_set.cpp:129-228),
//   2. packs the codes of every filtered row into one integer key and counts equal keys in a hash
//      table (combinationCountKernel: warp-aggregated, then per-CTA shared-memory table, then global),
//   3. compacts the table; the host orders the keys, which IS the reference's depth-first order
//      (dimension 0 in the most significant bits, symbols in SYMBOLS order, null last).
#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "eval_device.cuh"

namespace silo {

namespace {

constexpr uint32_t CODE_NONE = 0xFF;  // the row is in no group of the dimension: it is in no combination
constexpr uint32_t MAX_DIMS = 12;
constexpr uint32_t POSITION_CODE_BITS = 5;  // symbol ids < 28, null = n_symbols <= 28
constexpr uint32_t BITMAP_CODE_BITS = 8;    // <= 254 value groups, null = n_groups
constexpr size_t CODES_SHARED_BYTES = 65536 + TILE_BYTES + 16;

struct DevBitmapRef {  // a registered roaring bitmap (filter_eval.cu: descriptors keyed by the GLOBAL chunk id)
   const DevContainer* containers;
   const uint8_t* payload;
   uint32_t n_containers;
   uint32_t code;
};

struct PositionDims {
   uint32_t n;
   uint32_t plane[MAX_DIMS];     // index of the dimension's byte plane
   uint32_t position[MAX_DIMS];
   DevColumn column[MAX_DIMS];
};

struct BitmapDims {
   uint32_t n;
   uint32_t plane[MAX_DIMS];
   uint32_t group_begin[MAX_DIMS + 1];  // into `groups`
   const DevBitmapRef* groups;
};

// the rows whose bit is set in `word` (word index `tid` of a tile) get `value`
__device__ __forceinline__ void scatterCode(uint8_t* code, uint64_t word, uint32_t tid, uint32_t value) {
   while (word != 0) {
      const uint32_t bit = static_cast<uint32_t>(__ffsll(static_cast<long long>(word)) - 1);
      code[tid * 64 + bit] = static_cast<uint8_t>(value);
      word &= word - 1;
   }
}

__device__ __forceinline__ void writePlane(uint8_t* plane, const uint8_t* code, uint32_t tid) {
   uint4* out = reinterpret_cast<uint4*>(plane);
   const uint4* in = reinterpret_cast<const uint4*>(code);
   for (uint32_t i = tid; i < 65536 / 16; i += EVAL_THREADS) {
      out[i] = in[i];
   }
}

// grid (n_chunks, dims.n). code(row) at a sequence position, in the order compileSymbolInSet decides it:
//   the row is in the stored container of a symbol s (never the local reference symbol)   -> s
//   else position covered by the row and not one of its N positions                          -> local reference symbol
//   else the row's sequence is null                                                          -> n_symbols (null group)
//   else                                                                                     -> missing symbol
__global__ void __launch_bounds__(EVAL_THREADS) positionCodesKernel(
   PositionDims dims,
   const uint32_t* __restrict__ chunk_sizes,
   uint8_t* __restrict__ codes,  // [planes][n_chunks * 65536]
   uint32_t n_chunks
) {
   extern __shared__ __align__(128) uint8_t smem_raw[];
   uint8_t* code = smem_raw;
   uint64_t* tile = reinterpret_cast<uint64_t*>(smem_raw + 65536);
   uint32_t* range = reinterpret_cast<uint32_t*>(smem_raw + 65536 + TILE_BYTES);
   const uint32_t chunk = blockIdx.x;
   const uint32_t tid = threadIdx.x;
   const DevColumn& column = dims.column[blockIdx.y];
   const uint32_t position = dims.position[blockIdx.y];
   const uint32_t chunk_size = chunk_sizes[chunk];
   const uint32_t local_reference = column.local_reference[position];
   const uint32_t missing_symbol = column.missing_symbol;
   const uint2* rows = column.start_end + column.chunk_row_begin[chunk];

   for (uint32_t row = tid; row < 65536; row += EVAL_THREADS) {
      uint32_t value = CODE_NONE;
      if (row < chunk_size) {
         const uint2 covered = rows[row];
         value = covered.x <= position && position < covered.y ? local_reference : missing_symbol;
      }
      code[row] = static_cast<uint8_t>(value);
   }
   __syncthreads();
   const uint32_t missing_begin = column.chunk_missing_begin[chunk];
   const uint32_t missing_end = column.chunk_missing_begin[chunk + 1];
   for (uint32_t i = missing_begin + tid; i < missing_end; i += EVAL_THREADS) {
      if (rowMissingAt(column, i, position)) {
         code[column.missing_row[i]] = static_cast<uint8_t>(missing_symbol);
      }
   }
   __syncthreads();
   if (column.null_words != nullptr) {
      scatterCode(code, column.null_words[static_cast<size_t>(chunk) * TILE_WORDS + tid] & layoutWord(chunk_size, tid), tid, column.n_symbols);
   }
   findPositionRange(range, column, chunk, position);  // includes a __syncthreads
   const uint32_t lo = range[0];
   const uint32_t hi = range[1];
   for (uint32_t i = lo; i < hi; ++i) {
      const DevContainer desc = column.containers[i];
      tile[tid] = 0;
      __syncthreads();
      orContainerIntoTile(tile, column.payload, desc);
      __syncthreads();
      scatterCode(code, tile[tid], tid, desc.symbol());
   }
   __syncthreads();
   writePlane(codes + (static_cast<size_t>(dims.plane[blockIdx.y]) * n_chunks + chunk) * 65536, code, tid);
}

// grid (n_chunks, dims.n). code(row) = index of the group bitmap that holds the row (the null bitmap is
// the last group), CODE_NONE when none does.
__global__ void __launch_bounds__(EVAL_THREADS) bitmapCodesKernel(
   BitmapDims dims,
   const uint32_t* __restrict__ chunk_sizes,
   uint32_t first_chunk,
   uint8_t* __restrict__ codes,
   uint32_t n_chunks
) {
   extern __shared__ __align__(128) uint8_t smem_raw[];
   uint8_t* code = smem_raw;
   uint64_t* tile = reinterpret_cast<uint64_t*>(smem_raw + 65536);
   uint32_t* range = reinterpret_cast<uint32_t*>(smem_raw + 65536 + TILE_BYTES);
   const uint32_t chunk = blockIdx.x;
   const uint32_t tid = threadIdx.x;
   const uint64_t layout_word = layoutWord(chunk_sizes[chunk], tid);
   for (uint32_t row = tid; row < 65536; row += EVAL_THREADS) {
      code[row] = static_cast<uint8_t>(CODE_NONE);
   }
   const uint32_t key = first_chunk + chunk;
   for (uint32_t g = dims.group_begin[blockIdx.y]; g < dims.group_begin[blockIdx.y + 1]; ++g) {
      const DevBitmapRef bitmap = dims.groups[g];
      __syncthreads();
      if (tid < 32) {
         const uint32_t at = warpLowerBound(0, bitmap.n_containers, key, tid, [&](uint32_t index) {
            return bitmap.containers[index].position;
         });
         if (tid == 0) {
            range[0] = at;
            range[1] = (at < bitmap.n_containers && bitmap.containers[at].position == key) ? 1u : 0u;
         }
      }
      tile[tid] = 0;
      __syncthreads();
      if (range[1] != 0) {
         orContainerIntoTile(tile, bitmap.payload, bitmap.containers[range[0]]);
         __syncthreads();
         scatterCode(code, tile[tid] & layout_word, tid, bitmap.code);
      }
   }
   __syncthreads();
   writePlane(codes + (static_cast<size_t>(dims.plane[blockIdx.y]) * n_chunks + chunk) * 65536, code, tid);
}

// ---- counting -----------------------------------------------------------------------------------

constexpr uint64_t EMPTY_KEY = ~0ULL;

struct KeyLayout {
   uint32_t n_dims;
   uint32_t shift[MAX_DIMS];
};

struct CombinationTable {
   unsigned long long* keys;    // [capacity], EMPTY_KEY = free
   unsigned long long* counts;  // [capacity]
   uint32_t capacity;           // power of two
   uint32_t* state;             // [0] entries, [1] overflow flag
};

__device__ __forceinline__ uint32_t hashKey(uint64_t key) {
   key ^= key >> 33;
   key *= 0xff51afd7ed558ccdULL;
   key ^= key >> 33;
   return static_cast<uint32_t>(key);
}

__device__ void globalInsert(const CombinationTable& table, uint64_t key, uint64_t amount) {
   uint32_t slot = hashKey(key) & (table.capacity - 1);
   for (uint32_t probe = 0; probe < 4096 && probe < table.capacity; ++probe) {
      const unsigned long long seen = atomicCAS(&table.keys[slot], EMPTY_KEY, key);
      if (seen == EMPTY_KEY) {
         atomicAdd(&table.state[0], 1u);
      }
      if (seen == EMPTY_KEY || seen == key) {
         atomicAdd(&table.counts[slot], static_cast<unsigned long long>(amount));
         return;
      }
      slot = (slot + 1) & (table.capacity - 1);
   }
   atomicOr(&table.state[1], 1u);
}

constexpr int COUNT_THREADS = 256;
constexpr uint32_t COUNT_ROWS_PER_CTA = 8192;  // 8 CTAs per chunk
constexpr uint32_t LOCAL_SLOTS = 1024;

__global__ void __launch_bounds__(COUNT_THREADS) combinationCountKernel(
   const uint64_t* __restrict__ filter_words,
   const uint32_t* __restrict__ chunk_popcount,
   const uint8_t* __restrict__ codes,
   uint32_t n_chunks,
   KeyLayout layout,
   CombinationTable table
) {
   __shared__ unsigned long long local_keys[LOCAL_SLOTS];
   __shared__ uint32_t local_counts[LOCAL_SLOTS];
   const uint32_t chunk = blockIdx.x / (65536 / COUNT_ROWS_PER_CTA);
   if (chunk_popcount[chunk] == 0) {
      return;
   }
   const uint32_t first_row = (blockIdx.x % (65536 / COUNT_ROWS_PER_CTA)) * COUNT_ROWS_PER_CTA;
   const uint32_t tid = threadIdx.x;
   const uint32_t lane = tid & 31;
   for (uint32_t slot = tid; slot < LOCAL_SLOTS; slot += COUNT_THREADS) {
      local_keys[slot] = EMPTY_KEY;
      local_counts[slot] = 0;
   }
   __syncthreads();
   const uint32_t* tile32 = reinterpret_cast<const uint32_t*>(filter_words + static_cast<size_t>(chunk) * TILE_WORDS);
   const size_t plane_stride = static_cast<size_t>(n_chunks) * 65536;
   const uint8_t* chunk_codes = codes + static_cast<size_t>(chunk) * 65536;
   for (uint32_t base = first_row; base < first_row + COUNT_ROWS_PER_CTA; base += COUNT_THREADS) {
      const uint32_t row = base + tid;  // a warp covers exactly one 32-bit filter word
      const uint32_t word = tile32[row >> 5];
      if (word == 0) {
         continue;  // warp-uniform
      }
      bool active = ((word >> lane) & 1u) != 0;
      uint64_t key = 0;
      if (active) {
         for (uint32_t d = 0; d < layout.n_dims; ++d) {
            const uint32_t value = chunk_codes[d * plane_stride + row];
            active = active && value != CODE_NONE;
            key |= static_cast<uint64_t>(value) << layout.shift[d];
         }
      }
      // equal keys of the warp are counted once
      const uint32_t peers = __match_any_sync(0xFFFFFFFFu, active ? key : EMPTY_KEY);
      if (!active || lane != static_cast<uint32_t>(__ffs(peers) - 1)) {
         continue;
      }
      const uint32_t amount = __popc(peers);
      uint32_t slot = hashKey(key) & (LOCAL_SLOTS - 1);
      bool placed = false;
      for (uint32_t probe = 0; probe < 16; ++probe) {
         const unsigned long long seen = atomicCAS(&local_keys[slot], EMPTY_KEY, key);
         if (seen == EMPTY_KEY || seen == key) {
            atomicAdd(&local_counts[slot], amount);
            placed = true;
            break;
         }
         slot = (slot + 1) & (LOCAL_SLOTS - 1);
      }
      if (!placed) {
         globalInsert(table, key, amount);
      }
   }
   __syncthreads();
   for (uint32_t slot = tid; slot < LOCAL_SLOTS; slot += COUNT_THREADS) {
      if (local_keys[slot] != EMPTY_KEY) {
         globalInsert(table, local_keys[slot], local_counts[slot]);
      }
   }
}

// out[0] = {number of combinations, 0}; combinations from out[1], in table order
// Second version of the count kernel, the default (SILO_COOC_KERNEL=1 selects the first; measurement of the first: 254 us
// for 10 M rows x 6 dimensions, 61 MB of DRAM traffic -- 0.04 of the bandwidth roofline, bound by per-row work and by the
// flushes of 1,224 CTAs; with this one the co-occurrence query of bench.py --workload cooc went 0.542 -> 0.491 ms):
//  * a fixed grid of resident blocks that loop over the (chunk, 1,024-row) units, so a block's shared-memory table is
//    flushed to the global table ONCE;
//  * four consecutive rows per thread: one 32-bit load per dimension instead of four byte loads;
//  * the block's most frequent key (sampled from its first rows: in sequence data most rows carry the reference at
//    every position) is counted in a register, without any atomic.
constexpr int COUNT2_THREADS = 256;
constexpr uint32_t COUNT2_UNIT_ROWS = COUNT2_THREADS * 4;
constexpr uint32_t COUNT2_UNITS_PER_CHUNK = 65536 / COUNT2_UNIT_ROWS;

__device__ __forceinline__ void localInsert(
   unsigned long long* local_keys, uint32_t* local_counts, const CombinationTable& table, uint64_t key, uint32_t amount
) {
   uint32_t slot = hashKey(key) & (LOCAL_SLOTS - 1);
   for (uint32_t probe = 0; probe < 16; ++probe) {
      const unsigned long long seen = atomicCAS(&local_keys[slot], EMPTY_KEY, key);
      if (seen == EMPTY_KEY || seen == key) {
         atomicAdd(&local_counts[slot], amount);
         return;
      }
      slot = (slot + 1) & (LOCAL_SLOTS - 1);
   }
   globalInsert(table, key, amount);
}

__global__ void __launch_bounds__(COUNT2_THREADS) combinationCountKernelV2(
   const uint64_t* __restrict__ filter_words,
   const uint32_t* __restrict__ chunk_popcount,
   const uint8_t* __restrict__ codes,
   uint32_t n_chunks,
   KeyLayout layout,
   CombinationTable table
) {
   __shared__ unsigned long long local_keys[LOCAL_SLOTS];
   __shared__ uint32_t local_counts[LOCAL_SLOTS];
   __shared__ unsigned long long hot_key_shared;
   __shared__ unsigned long long hot_total;
   const uint32_t tid = threadIdx.x;
   const uint32_t lane = tid & 31;
   for (uint32_t slot = tid; slot < LOCAL_SLOTS; slot += COUNT2_THREADS) {
      local_keys[slot] = EMPTY_KEY;
      local_counts[slot] = 0;
   }
   if (tid == 0) {
      hot_key_shared = EMPTY_KEY;
      hot_total = 0;
   }
   __syncthreads();
   const size_t plane_stride = static_cast<size_t>(n_chunks) * 65536;
   const uint32_t n_units = n_chunks * COUNT2_UNITS_PER_CHUNK;
   // the four rows of this thread in a unit: their filter bits and keys (EMPTY_KEY: not in the filter / in no group)
   auto loadKeys = [&](uint32_t unit, uint64_t (&keys)[4]) {
      const uint32_t chunk = unit / COUNT2_UNITS_PER_CHUNK;
      const uint32_t row = (unit % COUNT2_UNITS_PER_CHUNK) * COUNT2_UNIT_ROWS + tid * 4;
      const uint32_t* tile32 = reinterpret_cast<const uint32_t*>(filter_words + static_cast<size_t>(chunk) * TILE_WORDS);
      const uint32_t bits = (tile32[row >> 5] >> (row & 31)) & 0xFu;
#pragma unroll
      for (uint32_t j = 0; j < 4; ++j) {
         keys[j] = ((bits >> j) & 1u) != 0 ? 0ULL : EMPTY_KEY;
      }
      if (bits == 0) {
         return;
      }
      const uint8_t* base = codes + static_cast<size_t>(chunk) * 65536 + row;
      for (uint32_t d = 0; d < layout.n_dims; ++d) {
         const uint32_t four = *reinterpret_cast<const uint32_t*>(base + d * plane_stride);
#pragma unroll
         for (uint32_t j = 0; j < 4; ++j) {
            const uint32_t value = (four >> (8 * j)) & 0xFFu;
            if (keys[j] != EMPTY_KEY) {
               keys[j] = value == CODE_NONE ? EMPTY_KEY : keys[j] | (static_cast<uint64_t>(value) << layout.shift[d]);
            }
         }
      }
   };
   uint32_t first_unit = blockIdx.x;
   while (first_unit < n_units && chunk_popcount[first_unit / COUNT2_UNITS_PER_CHUNK] == 0) {
      first_unit += gridDim.x;
   }
   if (first_unit >= n_units) {
      return;  // (block-uniform)
   }
   if (tid < 32) {  // the most frequent key among the first unit's first 128 rows
      uint64_t keys[4];
      loadKeys(first_unit, keys);
      uint64_t candidate = EMPTY_KEY;
#pragma unroll
      for (uint32_t j = 0; j < 4; ++j) {
         candidate = candidate == EMPTY_KEY ? keys[j] : candidate;
      }
      const uint32_t peers = __match_any_sync(0xFFFFFFFFu, candidate);
      const uint32_t votes = candidate == EMPTY_KEY ? 0u : static_cast<uint32_t>(__popc(peers));
      const uint32_t best = __reduce_max_sync(0xFFFFFFFFu, (votes << 5) | (31u - lane));
      const uint64_t winner = __shfl_sync(0xFFFFFFFFu, candidate, 31 - (best & 31u));
      if (lane == 0) {
         hot_key_shared = (best >> 5) != 0 ? winner : EMPTY_KEY;
      }
   }
   __syncthreads();
   const uint64_t hot_key = hot_key_shared;
   uint32_t hot_count = 0;
   for (uint32_t unit = first_unit; unit < n_units; unit += gridDim.x) {
      if (chunk_popcount[unit / COUNT2_UNITS_PER_CHUNK] == 0) {
         continue;  // (block-uniform)
      }
      uint64_t keys[4];
      loadKeys(unit, keys);
#pragma unroll
      for (uint32_t j = 0; j < 4; ++j) {
         if (keys[j] == EMPTY_KEY) {
            continue;
         }
         if (keys[j] == hot_key) {
            ++hot_count;
         } else {
            localInsert(local_keys, local_counts, table, keys[j], 1u);
         }
      }
   }
   hot_count = __reduce_add_sync(0xFFFFFFFFu, hot_count);
   if (lane == 0 && hot_count != 0) {
      atomicAdd(&hot_total, static_cast<unsigned long long>(hot_count));
   }
   __syncthreads();
   if (tid == 0 && hot_total != 0) {
      globalInsert(table, hot_key, hot_total);
   }
   for (uint32_t slot = tid; slot < LOCAL_SLOTS; slot += COUNT2_THREADS) {
      if (local_keys[slot] != EMPTY_KEY) {
         globalInsert(table, local_keys[slot], local_counts[slot]);
      }
   }
}

inline int combinationKernelVersion() {
   static const int version = [] {
      const char* flag = std::getenv("SILO_COOC_KERNEL");  // (1: the first version, kept for comparison)
      return flag != nullptr && flag[0] == '1' ? 1 : 2;
   }();
   return version;
}

__global__ void compactCombinationsKernel(CombinationTable table, silo_combination* __restrict__ out, uint32_t out_capacity) {
   const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
   if (slot >= table.capacity || table.keys[slot] == EMPTY_KEY) {
      return;
   }
   const unsigned long long index = atomicAdd(reinterpret_cast<unsigned long long*>(&out[0].key), 1ULL);
   if (index < out_capacity) {
      out[1 + index] = silo_combination{table.keys[slot], table.counts[slot]};
   }
}

}  // namespace

}  // namespace silo

using namespace silo;

extern "C" {

int silo_gpu_query_combinations(
   silo_gpu_table* table,
   const silo_filter_program* program,
   const silo_gpu_filter* filter,
   const silo_group_dimension* dimensions,
   uint32_t n_dimensions,
   const silo_combination** combinations,
   uint64_t* n_combinations,
   uint64_t* cardinality
) {
   return guarded([&] {
      require(table != nullptr && combinations != nullptr && n_combinations != nullptr, "silo_gpu_query_combinations: NULL argument");
      require(n_dimensions == 0 || dimensions != nullptr, "silo_gpu_query_combinations: dimensions is NULL");
      require(n_dimensions <= MAX_DIMS, "silo_gpu_query_combinations: at most 12 grouping dimensions");
      std::lock_guard<std::mutex> lock(table->mutex);
      SILO_CUDA_CHECK(cudaSetDevice(table->ctx->device));
      cudaStream_t stream = table->ctx->stream;
      const uint32_t n_chunks = table->n_chunks;

      // ---- key layout: dimension 0 in the most significant bits ----
      PositionDims position_dims{};
      BitmapDims bitmap_dims{};
      KeyLayout layout{};
      layout.n_dims = n_dimensions;
      std::vector<DevBitmapRef> groups;
      uint32_t total_bits = 0;
      std::vector<uint32_t> bits(n_dimensions);
      for (uint32_t d = 0; d < n_dimensions; ++d) {
         const silo_group_dimension& dimension = dimensions[d];
         if (dimension.kind == SILO_DIM_SEQUENCE_POSITION) {
            require(dimension.column >= 0 && static_cast<size_t>(dimension.column) < table->columns.size(), "silo_gpu_query_combinations: bad column index");
            const DevColumn& column = table->columns[static_cast<size_t>(dimension.column)]->dev;
            require(dimension.position < column.genome_length, "silo_gpu_query_combinations: position is out of bounds");
            position_dims.plane[position_dims.n] = d;
            position_dims.position[position_dims.n] = dimension.position;
            position_dims.column[position_dims.n] = column;
            position_dims.n++;
            bits[d] = POSITION_CODE_BITS;
         } else if (dimension.kind == SILO_DIM_INDEX_BITMAPS) {
            require(dimension.n_groups <= 254, "silo_gpu_query_combinations: at most 254 value groups per dimension");
            require(dimension.n_groups == 0 || dimension.bitmap_ids != nullptr, "silo_gpu_query_combinations: bitmap_ids is NULL");
            bitmap_dims.plane[bitmap_dims.n] = d;
            bitmap_dims.group_begin[bitmap_dims.n] = static_cast<uint32_t>(groups.size());
            auto addGroup = [&](uint32_t id, uint32_t code) {
               require(id < table->registered.size() && table->registered[id].d_block != nullptr, "silo_gpu_query_combinations: unknown registered bitmap id");
               const silo_gpu_table::RegisteredBitmap& registered = table->registered[id];
               groups.push_back(DevBitmapRef{
                  reinterpret_cast<const DevContainer*>(registered.d_block), registered.d_block + registered.payload_offset, registered.n_containers, code
               });
            };
            for (uint32_t g = 0; g < dimension.n_groups; ++g) {
               addGroup(dimension.bitmap_ids[g], g);
            }
            if (dimension.null_bitmap_id != UINT32_MAX) {
               addGroup(dimension.null_bitmap_id, dimension.n_groups);
            }
            bitmap_dims.n++;
            bitmap_dims.group_begin[bitmap_dims.n] = static_cast<uint32_t>(groups.size());
            bits[d] = BITMAP_CODE_BITS;
         } else {
            throw ApiError(SILO_E_INVALID_ARGUMENT, "silo_gpu_query_combinations: unknown dimension kind");
         }
         total_bits += bits[d];
      }
      require(total_bits <= 63, "silo_gpu_query_combinations: the grouping key does not fit 63 bits");
      for (uint32_t d = 0, below = total_bits; d < n_dimensions; ++d) {
         below -= bits[d];
         layout.shift[d] = below;
      }

      uint8_t* d_staging = nullptr;
      silo_gpu_filter* own_filter = nullptr;
      uint8_t* d_codes = nullptr;
      DevBitmapRef* d_groups = nullptr;
      unsigned long long* d_table = nullptr;
      uint32_t* d_state = nullptr;
      silo_combination* d_out = nullptr;
      auto release = [&]() {
         for (void* pointer : {static_cast<void*>(d_staging), static_cast<void*>(d_codes), static_cast<void*>(d_groups),
                               static_cast<void*>(d_table), static_cast<void*>(d_state), static_cast<void*>(d_out)}) {
            if (pointer != nullptr) {
               cudaFreeAsync(pointer, stream);
            }
         }
         d_staging = nullptr;
         d_codes = nullptr;
         d_groups = nullptr;
         d_table = nullptr;
         d_state = nullptr;
         d_out = nullptr;
      };
      unsigned long long host_cardinality = 0;
      uint32_t host_error = 0;
      thread_local std::vector<silo_combination> result;
      result.clear();
      try {
         // ---- the filter ----
         const uint64_t* words = table->d_full_words;
         const uint32_t* popcounts = table->d_chunk_popcount_full;
         if (program != nullptr) {
            own_filter = evalProgramAsync(table, program, stream, &d_staging);
            filter = own_filter;
            SILO_CUDA_CHECK(cudaMemcpyAsync(&host_cardinality, own_filter->d_cardinality, sizeof(host_cardinality), cudaMemcpyDeviceToHost, stream));
            SILO_CUDA_CHECK(cudaMemcpyAsync(&host_error, own_filter->d_error_flag, sizeof(host_error), cudaMemcpyDeviceToHost, stream));
         }
         if (filter != nullptr) {
            require(filter->table == table, "silo_gpu_query_combinations: filter belongs to another table");
            if (filter->out_of_layout) {
               throw ApiError(SILO_E_OUT_OF_LAYOUT, "the filter holds row ids outside the row layout: the aggregation has no row data for them");
            }
            words = filter->d_words;
            popcounts = filter->d_chunk_popcount;
         }
         if (n_chunks > 0) {
            // ---- 1. code planes ----
            static bool attributes_set = false;
            if (!attributes_set) {
               SILO_CUDA_CHECK(cudaFuncSetAttribute(positionCodesKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(CODES_SHARED_BYTES)));
               SILO_CUDA_CHECK(cudaFuncSetAttribute(bitmapCodesKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(CODES_SHARED_BYTES)));
               attributes_set = true;
            }
            d_codes = poolAlloc<uint8_t>(static_cast<size_t>(std::max(n_dimensions, 1u)) * n_chunks * 65536, stream);
            if (position_dims.n > 0) {
               positionCodesKernel<<<dim3(n_chunks, position_dims.n), EVAL_THREADS, CODES_SHARED_BYTES, stream>>>(
                  position_dims, table->d_chunk_sizes, d_codes, n_chunks
               );
               SILO_CUDA_CHECK(cudaGetLastError());
               table->stats.kernel_launches++;
            }
            if (bitmap_dims.n > 0) {
               d_groups = poolAlloc<DevBitmapRef>(groups.size(), stream);
               // (pageable source: the runtime stages it before the call returns)
               SILO_CUDA_CHECK(cudaMemcpyAsync(d_groups, groups.data(), groups.size() * sizeof(DevBitmapRef), cudaMemcpyHostToDevice, stream));
               bitmap_dims.groups = d_groups;
               bitmapCodesKernel<<<dim3(n_chunks, bitmap_dims.n), EVAL_THREADS, CODES_SHARED_BYTES, stream>>>(
                  bitmap_dims, table->d_chunk_sizes, table->first_chunk, d_codes, n_chunks
               );
               SILO_CUDA_CHECK(cudaGetLastError());
               table->stats.kernel_launches++;
            }
            // ---- 2. + 3. count, compact; a table that turns out too small is rebuilt eight times larger ----
            d_state = poolAlloc<uint32_t>(2, stream);
            for (uint32_t capacity = 1u << 16;; capacity <<= 3) {
               d_table = poolAlloc<unsigned long long>(2ULL * capacity, stream);
               d_out = poolAlloc<silo_combination>(static_cast<size_t>(capacity) + 1, stream);
               SILO_CUDA_CHECK(cudaMemsetAsync(d_table, 0xFF, sizeof(unsigned long long) * capacity, stream));
               SILO_CUDA_CHECK(cudaMemsetAsync(d_table + capacity, 0, sizeof(unsigned long long) * capacity, stream));
               SILO_CUDA_CHECK(cudaMemsetAsync(d_state, 0, 2 * sizeof(uint32_t), stream));
               SILO_CUDA_CHECK(cudaMemsetAsync(d_out, 0, sizeof(silo_combination), stream));
               const CombinationTable combination_table{d_table, d_table + capacity, capacity, d_state};
               if (combinationKernelVersion() == 2) {
                  const uint32_t units = n_chunks * COUNT2_UNITS_PER_CHUNK;
                  const uint32_t blocks = std::min<uint32_t>(units, static_cast<uint32_t>(table->ctx->sm_count) * 4u);
                  combinationCountKernelV2<<<blocks, COUNT2_THREADS, 0, stream>>>(words, popcounts, d_codes, n_chunks, layout, combination_table);
               } else {
                  combinationCountKernel<<<n_chunks * (65536 / COUNT_ROWS_PER_CTA), COUNT_THREADS, 0, stream>>>(
                     words, popcounts, d_codes, n_chunks, layout, combination_table
                  );
               }
               SILO_CUDA_CHECK(cudaGetLastError());
               compactCombinationsKernel<<<(capacity + 255) / 256, 256, 0, stream>>>(combination_table, d_out, capacity);
               SILO_CUDA_CHECK(cudaGetLastError());
               table->stats.kernel_launches += 2;
               uint32_t host_state[2] = {0, 0};
               SILO_CUDA_CHECK(cudaMemcpyAsync(host_state, d_state, sizeof(host_state), cudaMemcpyDeviceToHost, stream));
               // the header and the first combinations in one copy; the rest only if there are more
               constexpr size_t FIRST_COPY = 2048;
               result.resize(FIRST_COPY);
               SILO_CUDA_CHECK(cudaMemcpyAsync(result.data(), d_out, FIRST_COPY * sizeof(silo_combination), cudaMemcpyDeviceToHost, stream));
               SILO_CUDA_CHECK(cudaStreamSynchronize(stream));
               const bool too_small = host_state[1] != 0 || host_state[0] > capacity / 2;
               if (!too_small) {
                  const uint64_t count = result[0].key;
                  result.resize(count + 1);
                  if (count + 1 > FIRST_COPY) {
                     SILO_CUDA_CHECK(cudaMemcpy(
                        result.data() + FIRST_COPY, d_out + FIRST_COPY, (count + 1 - FIRST_COPY) * sizeof(silo_combination), cudaMemcpyDeviceToHost
                     ));
                  }
                  break;
               }
               require(capacity < (1u << 28), "silo_gpu_query_combinations: more than 2^27 distinct combinations");
               cudaFreeAsync(d_table, stream);
               cudaFreeAsync(d_out, stream);
               d_table = nullptr;
               d_out = nullptr;
            }
         } else {
            SILO_CUDA_CHECK(cudaStreamSynchronize(stream));
            result.assign(1, silo_combination{0, 0});
         }
      } catch (...) {
         release();
         cudaStreamSynchronize(stream);
         releaseFilterLocked(own_filter);
         throw;
      }
      release();
      SILO_CUDA_CHECK(cudaStreamSynchronize(stream));
      releaseFilterLocked(own_filter);
      if (host_error != 0) {
         throw ApiError(SILO_E_OUT_OF_LAYOUT, "a leaf bitmap holds row ids outside the row layout");
      }
      // ascending keys = the reference's depth-first order (partition(), bitmap_aggregation_node.cpp:99-124)
      std::sort(result.begin() + 1, result.end(), [](const silo_combination& a, const silo_combination& b) { return a.key < b.key; });
      *combinations = result.data() + 1;
      *n_combinations = result.size() - 1;
      if (cardinality != nullptr && program != nullptr) {
         *cardinality = host_cardinality;
      }
   });
}

}  // extern "C"
