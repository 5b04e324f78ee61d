// S1: context, table layout and the device-resident container pool of a sequence column.
//
// Replaces the host-memory layout of rhydb::storage::column::SequenceColumn
// (/root/reference/src/rhydb/storage/column/sequence_column.h:104-114): the std::map of
// individually malloc'ed roaring containers (vertical_sequence_index.h:44) becomes ONE payload slab
// plus a 16-byte descriptor array in chunk-major order, cut into <=16 KiB segments that a CTA pulls
// into shared memory with 1-D bulk (TMA) copies. See DESIGN.md "Data layout in HBM".
#include <algorithm>
#include <cstring>
#include <memory>
#include <numeric>

#include "common.cuh"

namespace silo {

namespace {
thread_local std::string g_last_error;
}

void setLastError(const std::string& message) {
   g_last_error = message;
}

namespace {

__global__ void fillLayoutTiles(uint64_t* words, const uint32_t* chunk_sizes, uint32_t n_chunks) {
   const uint32_t chunk = blockIdx.x;
   if (chunk >= n_chunks) {
      return;
   }
   for (uint32_t w = threadIdx.x; w < TILE_WORDS; w += blockDim.x) {
      words[static_cast<size_t>(chunk) * TILE_WORDS + w] = layoutWord(chunk_sizes[chunk], w);
   }
}

uint32_t alignUp(uint32_t value, uint32_t alignment) {
   return (value + alignment - 1) / alignment * alignment;
}

}  // namespace

}  // namespace silo

using namespace silo;

// ---- re-encoding of stored containers into the device piece formats (common.cuh) --------------

// n sorted u16 values -> KIND_ARRAY_T payload
static void encodeArrayPiece(const uint8_t* sorted_values, uint32_t n, std::vector<uint8_t>& out) {
   out.assign(arrayPieceBytes(n), 0);
   auto valueAt = [&](uint32_t index) {
      uint16_t value = 0;
      std::memcpy(&value, sorted_values + 2ULL * std::min(index, n - 1), 2);  // padding repeats the last value
      return static_cast<uint16_t>(value ^ ARRAY_VALUE_FLIP);
   };
   for (uint32_t first = 0; first < n; first += ARRAY_REGION_VALUES) {
      const uint32_t count = std::min(ARRAY_REGION_VALUES, n - first);
      const uint32_t lanes = arrayRegionLanes(count);
      auto* slots = reinterpret_cast<uint16_t*>(out.data() + (first / ARRAY_REGION_VALUES) * 512u);
      for (uint32_t lane = 0; lane < lanes; ++lane) {
         for (uint32_t j = 0; j < 8; ++j) {
            slots[lane * 8 + j] = valueAt(first + std::min(lane + lanes * j, count - 1));
         }
      }
   }
}

// n KIND_RUNS_W entries (ascending) -> region order with padding entries
static void encodeRunsPiece(const uint32_t* entries, uint32_t n, std::vector<uint8_t>& out) {
   out.assign(runsPieceBytes(n), 0);
   for (uint32_t first = 0; first < n; first += RUNS_REGION_ENTRIES) {
      const uint32_t count = std::min(RUNS_REGION_ENTRIES, n - first);
      const uint32_t lanes = runsRegionLanes(count);
      auto* slots = reinterpret_cast<uint32_t*>(out.data() + (first / RUNS_REGION_ENTRIES) * 512u);
      for (uint32_t lane = 0; lane < lanes; ++lane) {
         for (uint32_t j = 0; j < 4; ++j) {
            const uint32_t index = lane + lanes * j;
            slots[lane * 4 + j] = index < count ? entries[first + index] : RUNS_PAD_ENTRY;
         }
      }
   }
}

// CRoaring run pairs {start, length-1} -> runs confined to one 32-row word each (KIND_RUNS_W entries)
// plus ranges of whole words for the long runs (KIND_WORDRANGE entries, popcounted straight from the tile)
static void splitRuns(
   const uint8_t* pairs,
   uint32_t n_runs,
   std::vector<uint32_t>& word_entries,
   std::vector<uint32_t>& word_ranges,
   uint64_t* covered_rows
) {
   constexpr uint32_t MAX_INLINE_FULL_WORDS = 3;  // shorter stretches of whole words stay plain entries
   word_entries.clear();
   word_ranges.clear();
   uint64_t covered = 0;
   uint32_t previous_end = 0;
   for (uint32_t r = 0; r < n_runs; ++r) {
      uint16_t start = 0;
      uint16_t length_minus_one = 0;
      std::memcpy(&start, pairs + 4ULL * r, 2);
      std::memcpy(&length_minus_one, pairs + 4ULL * r + 2, 2);
      const uint32_t first = start;
      const uint32_t last = first + length_minus_one;  // inclusive
      require(last <= 65535 && (r == 0 || first >= previous_end), "run container runs must be ascending and inside the chunk");
      previous_end = last + 1;
      covered += length_minus_one + 1u;
      const uint32_t first_word = first >> 5;
      const uint32_t last_word = last >> 5;
      if (first_word == last_word) {
         word_entries.push_back(runEntry(first_word, first & 31u, last - first + 1));
         continue;
      }
      word_entries.push_back(runEntry(first_word, first & 31u, 32u - (first & 31u)));
      const uint32_t whole_first = first_word + 1;
      const uint32_t whole_end = last_word;  // exclusive
      if (whole_end - whole_first <= MAX_INLINE_FULL_WORDS) {
         for (uint32_t word = whole_first; word < whole_end; ++word) {
            word_entries.push_back(runEntry(word, 0, 32));
         }
      } else {
         word_ranges.push_back(whole_first | (whole_end << 16));
      }
      word_entries.push_back(runEntry(last_word, 0, (last & 31u) + 1));
   }
   *covered_rows = covered;
}

extern "C" {

const char* silo_gpu_last_error(void) {
   return g_last_error.c_str();
}

const char* silo_gpu_version(void) {
   return "libsilo_b200 0.1.0 sm_100a";
}

int silo_gpu_init(int device_ordinal, silo_gpu_ctx** out) {
   return guarded([&] {
      require(out != nullptr, "silo_gpu_init: out is NULL");
      int device_count = 0;
      const cudaError_t status = cudaGetDeviceCount(&device_count);
      if (status != cudaSuccess || device_count == 0) {
         throw ApiError(
            SILO_E_NO_DEVICE,
            std::string("no CUDA device available (there is no CPU fallback): ") +
               cudaGetErrorString(status)
         );
      }
      require(device_ordinal >= 0 && device_ordinal < device_count, "silo_gpu_init: bad device ordinal");
      SILO_CUDA_CHECK(cudaSetDevice(device_ordinal));
      cudaDeviceProp prop{};
      SILO_CUDA_CHECK(cudaGetDeviceProperties(&prop, device_ordinal));
      if (prop.major < 10) {
         throw ApiError(
            SILO_E_NO_DEVICE,
            std::string("device '") + prop.name + "' is not sm_100-class; this library ships sm_100a code only"
         );
      }
      cudaMemPool_t pool = nullptr;
      SILO_CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, device_ordinal));
      uint64_t keep_everything = UINT64_MAX;
      SILO_CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep_everything));
      auto ctx = std::make_unique<silo_gpu_ctx>();
      ctx->device = device_ordinal;
      ctx->sm_count = prop.multiProcessorCount;
      SILO_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
      *out = ctx.release();
   });
}

void silo_gpu_shutdown(silo_gpu_ctx* ctx) {
   if (ctx == nullptr) {
      return;
   }
   cudaSetDevice(ctx->device);
   if (ctx->stream != nullptr) {
      cudaStreamDestroy(ctx->stream);
   }
   delete ctx;
}

void* silo_gpu_host_alloc(silo_gpu_ctx* ctx, uint64_t bytes) {
   void* ptr = nullptr;
   const int status = guarded([&] {
      require(ctx != nullptr, "silo_gpu_host_alloc: NULL context");
      SILO_CUDA_CHECK(cudaSetDevice(ctx->device));
      SILO_CUDA_CHECK(cudaMallocHost(&ptr, std::max<uint64_t>(bytes, 1)));
   });
   return status == SILO_OK ? ptr : nullptr;
}

void silo_gpu_host_free(silo_gpu_ctx* ctx, void* ptr) {
   if (ctx != nullptr && ptr != nullptr) {
      cudaSetDevice(ctx->device);
      cudaFreeHost(ptr);
   }
}

int silo_gpu_table_create(
   silo_gpu_ctx* ctx,
   uint32_t first_chunk,
   const uint32_t* chunk_sizes,
   uint32_t n_chunks,
   silo_gpu_table** out
) {
   return guarded([&] {
      require(ctx != nullptr && out != nullptr, "silo_gpu_table_create: NULL argument");
      require(n_chunks == 0 || chunk_sizes != nullptr, "silo_gpu_table_create: chunk_sizes is NULL");
      // row_layout.h:40-45: chunks are never empty, hold <= 2^16 rows, and there are < 65535 of them
      require(static_cast<uint64_t>(first_chunk) + n_chunks < 65535, "too many chunks for 16-bit chunk ids");
      SILO_CUDA_CHECK(cudaSetDevice(ctx->device));
      auto table = std::make_unique<silo_gpu_table>();
      table->ctx = ctx;
      table->first_chunk = first_chunk;
      table->n_chunks = n_chunks;
      table->chunk_sizes.assign(chunk_sizes, chunk_sizes + n_chunks);
      for (uint32_t size : table->chunk_sizes) {
         require(size >= 1 && size <= 65536, "chunk sizes must be in [1, 65536]");
         table->n_rows += size;
      }
      table->d_chunk_sizes = deviceUpload(table->chunk_sizes, ctx->stream, &table->device_bytes);
      table->d_chunk_popcount_full = deviceUpload(table->chunk_sizes, ctx->stream, &table->device_bytes);
      table->d_work_state = deviceAlloc<uint32_t>(4, &table->device_bytes);
      SILO_CUDA_CHECK(cudaMemsetAsync(table->d_work_state, 0, 4 * sizeof(uint32_t), ctx->stream));
      table->d_full_words =
         deviceAlloc<uint64_t>(static_cast<size_t>(n_chunks) * TILE_WORDS, &table->device_bytes);
      if (n_chunks > 0) {
         fillLayoutTiles<<<n_chunks, 256, 0, ctx->stream>>>(table->d_full_words, table->d_chunk_sizes, n_chunks);
         SILO_CUDA_CHECK(cudaGetLastError());
         table->stats.kernel_launches++;
      }
      for (int i = 0; i < silo_gpu_table::EVENT_RING; ++i) {
         SILO_CUDA_CHECK(cudaEventCreate(&table->ev_begin[i]));
         SILO_CUDA_CHECK(cudaEventCreate(&table->ev_k1_begin[i]));
         SILO_CUDA_CHECK(cudaEventCreate(&table->ev_k1_end[i]));
         SILO_CUDA_CHECK(cudaEventCreate(&table->ev_end[i]));
      }
      SILO_CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&table->h_scalars_pinned), 4 * sizeof(unsigned long long)));
      SILO_CUDA_CHECK(cudaEventCreateWithFlags(&table->ev_free_fence, cudaEventDisableTiming));
      SILO_CUDA_CHECK(cudaEventCreateWithFlags(&table->ev_staging_copied, cudaEventDisableTiming));
      SILO_CUDA_CHECK(cudaEventCreateWithFlags(&table->ev_fork, cudaEventDisableTiming));
      SILO_CUDA_CHECK(cudaEventCreateWithFlags(&table->ev_join, cudaEventDisableTiming));
      SILO_CUDA_CHECK(cudaStreamCreateWithFlags(&table->aux_stream, cudaStreamNonBlocking));
      SILO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
      *out = table.release();
   });
}

void silo_gpu_table_free(silo_gpu_table* table) {
   if (table == nullptr) {
      return;
   }
   cudaSetDevice(table->ctx->device);
   cudaStreamSynchronize(table->ctx->stream);
   for (HostColumn* column : table->columns) {
      for (void* allocation : column->allocations) {
         cudaFree(allocation);
      }
      delete column;
   }
   for (const silo_gpu_table::RegisteredBitmap& registered : table->registered) {
      cudaFree(registered.d_block);
   }
   for (const DevValueColumn& column : table->value_columns) {
      cudaFree(const_cast<uint32_t*>(column.values));
      cudaFree(const_cast<uint64_t*>(column.null_words));
   }
   cudaFree(table->d_chunk_row_begin);
   cudaFree(table->d_chunk_sizes);
   cudaFree(table->d_chunk_popcount_full);
   cudaFree(table->d_work_state);
   cudaFree(table->d_work_items);
   cudaFree(table->d_full_words);
   cudaFree(table->d_coverage_diff);
   cudaFree(table->d_counts);
   dropQueryGraphsLocked(table);
   cudaFree(table->d_staging_fixed);
   cudaFree(table->d_sweep_counters);
   for (int i = 0; i < silo_gpu_table::SWEEP_EVENT_RING; ++i) {
      for (cudaEvent_t event : {table->ev_sweep_begin[i], table->ev_sweep_end[i]}) {
         if (event != nullptr) {
            cudaEventDestroy(event);
         }
      }
   }
   freeShardGroup(table);
   if (table->query_filter != nullptr) {
      cudaFree(table->query_filter->d_words);
      delete table->query_filter;
   }
   if (table->h_hits_pinned != nullptr) {
      cudaFreeHost(table->h_hits_pinned);
   }
   if (table->h_counts_pinned != nullptr) {
      cudaFreeHost(table->h_counts_pinned);
   }
   if (table->h_scalars_pinned != nullptr) {
      cudaFreeHost(table->h_scalars_pinned);
   }
   if (table->h_staging_pinned != nullptr) {
      cudaFreeHost(table->h_staging_pinned);
   }
   if (table->aux_stream != nullptr) {
      cudaStreamSynchronize(table->aux_stream);
      cudaStreamDestroy(table->aux_stream);
   }
   for (cudaEvent_t event : {table->ev_fork, table->ev_join}) {
      if (event != nullptr) {
         cudaEventDestroy(event);
      }
   }
   if (table->ev_free_fence != nullptr) {
      cudaEventDestroy(table->ev_free_fence);
   }
   if (table->ev_staging_copied != nullptr) {
      cudaEventDestroy(table->ev_staging_copied);
   }
   for (int i = 0; i < silo_gpu_table::EVENT_RING; ++i) {
      for (cudaEvent_t event : {table->ev_begin[i], table->ev_k1_begin[i], table->ev_k1_end[i], table->ev_end[i]}) {
         if (event != nullptr) {
            cudaEventDestroy(event);
         }
      }
   }
   delete table;
}

uint64_t silo_gpu_table_device_bytes(const silo_gpu_table* table) {
   if (table == nullptr) {
      return 0;
   }
   uint64_t total = table->device_bytes;
   for (const HostColumn* column : table->columns) {
      total += column->device_bytes;
   }
   return total;
}

int silo_gpu_value_column_upload(silo_gpu_table* table, const uint32_t* values, const uint32_t* null_row_ids, uint64_t n_null_rows) {
   int index = -1;
   const int status = guarded([&] {
      require(table != nullptr && (values != nullptr || table->n_rows == 0), "silo_gpu_value_column_upload: NULL argument");
      require(n_null_rows == 0 || null_row_ids != nullptr, "silo_gpu_value_column_upload: null_row_ids is NULL");
      std::lock_guard<std::mutex> lock(table->mutex);
      SILO_CUDA_CHECK(cudaSetDevice(table->ctx->device));
      cudaStream_t stream = table->ctx->stream;
      if (table->d_chunk_row_begin == nullptr) {
         std::vector<uint32_t> chunk_row_begin(table->n_chunks + 1, 0);
         for (uint32_t chunk = 0; chunk < table->n_chunks; ++chunk) {
            chunk_row_begin[chunk + 1] = chunk_row_begin[chunk] + table->chunk_sizes[chunk];
         }
         table->d_chunk_row_begin = deviceUpload(chunk_row_begin, stream, &table->device_bytes);
      }
      DevValueColumn column{};
      uint32_t* d_values = deviceAlloc<uint32_t>(table->n_rows, &table->device_bytes);
      if (table->n_rows > 0) {
         SILO_CUDA_CHECK(cudaMemcpyAsync(d_values, values, table->n_rows * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
      }
      column.values = d_values;
      if (n_null_rows > 0) {
         std::vector<uint64_t> null_words(static_cast<size_t>(table->n_chunks) * TILE_WORDS, 0);
         for (uint64_t i = 0; i < n_null_rows; ++i) {
            const uint32_t row_id = null_row_ids[i];
            const uint32_t global_chunk = row_id >> 16;
            require(global_chunk >= table->first_chunk && global_chunk < table->first_chunk + table->n_chunks, "null row outside the shard");
            const uint32_t local_chunk = global_chunk - table->first_chunk;
            require((row_id & 0xFFFF) < table->chunk_sizes[local_chunk], "null row outside the row layout");
            null_words[static_cast<size_t>(local_chunk) * TILE_WORDS + ((row_id & 0xFFFF) >> 6)] |= 1ULL << (row_id & 63);
         }
         column.null_words = deviceUpload(null_words, stream, &table->device_bytes);
      }
      SILO_CUDA_CHECK(cudaStreamSynchronize(stream));
      dropQueryGraphsLocked(table);
      table->value_columns.push_back(column);
      index = static_cast<int>(table->value_columns.size()) - 1;
   });
   return status == SILO_OK ? index : status;
}

int silo_gpu_table_set_option(silo_gpu_table* table, const char* name, uint64_t value) {
   return guarded([&] {
      require(table != nullptr && name != nullptr, "silo_gpu_table_set_option: NULL argument");
      std::lock_guard<std::mutex> lock(table->mutex);
      if (std::strcmp(name, "sweep_min_pieces") == 0) {
         table->sweep_min_pieces = value;
         dropQueryGraphsLocked(table);
      } else {
         throw ApiError(SILO_E_INVALID_ARGUMENT, std::string("silo_gpu_table_set_option: unknown option '") + name + "'");
      }
   });
}

int silo_gpu_column_upload(silo_gpu_table* table, const silo_column_desc* in) {
   int column_index = -1;
   const int status = guarded([&] {
      require(table != nullptr && in != nullptr, "silo_gpu_column_upload: NULL argument");
      require(in->struct_size == sizeof(silo_column_desc), "silo_column_desc.struct_size mismatch");
      require(in->n_symbols >= 2 && in->n_symbols <= 32, "n_symbols must be in [2, 32]");
      require(in->genome_length > 0, "genome_length must be positive");
      require(in->missing_symbol < in->n_symbols, "missing_symbol out of range");
      require(in->local_reference != nullptr, "local_reference is NULL");
      require(table->n_rows == 0 || in->start_end != nullptr, "start_end is NULL");
      std::lock_guard<std::mutex> lock(table->mutex);
      SILO_CUDA_CHECK(cudaSetDevice(table->ctx->device));
      cudaStream_t stream = table->ctx->stream;
      const uint32_t n_chunks = table->n_chunks;
      const uint32_t first_chunk = table->first_chunk;

      for (uint32_t position = 0; position < in->genome_length; ++position) {
         require(in->local_reference[position] < in->n_symbols, "local_reference holds an invalid symbol id");
      }

      // ---- containers: validate, re-order chunk-major, cut into segments, re-lay the payload ----
      std::vector<uint64_t> order(in->n_containers);
      std::iota(order.begin(), order.end(), 0);
      for (uint64_t i = 0; i < in->n_containers; ++i) {
         const silo_container_desc& c = in->containers[i];
         require(c.v_index >= first_chunk && c.v_index < first_chunk + n_chunks, "container v_index outside the shard");
         require(c.position < in->genome_length, "container position >= genome_length");
         require(c.symbol < in->n_symbols, "container symbol out of range");
         require(c.cardinality >= 1 && c.cardinality <= 65536, "container cardinality must be in [1, 65536]");
         require(c.payload_offset + c.payload_bytes <= in->payload_bytes, "container payload out of bounds");
         if (c.typecode == TYPE_BITSET) {
            require(c.payload_bytes == 8192, "bitset payload must be 8192 bytes");
         } else if (c.typecode == TYPE_ARRAY) {
            require(c.payload_bytes == 2 * c.cardinality && c.cardinality <= 65536, "array payload must be 2*cardinality bytes");
         } else if (c.typecode == TYPE_RUN) {
            require(c.payload_bytes >= 2, "run payload too short");
            uint16_t n_runs = 0;
            std::memcpy(&n_runs, in->payload + c.payload_offset, 2);
            require(c.payload_bytes == 2 + 4u * n_runs, "run payload must be 2+4*n_runs bytes");
         } else {
            throw ApiError(SILO_E_INVALID_ARGUMENT, "unknown roaring container typecode");
         }
      }
      std::sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) {
         const auto& ca = in->containers[a];
         const auto& cb = in->containers[b];
         if (ca.v_index != cb.v_index) {
            return ca.v_index < cb.v_index;
         }
         if (ca.position != cb.position) {
            return ca.position < cb.position;
         }
         return ca.symbol < cb.symbol;
      });
      for (uint64_t i = 1; i < order.size(); ++i) {
         const auto& prev = in->containers[order[i - 1]];
         const auto& cur = in->containers[order[i]];
         require(
            !(prev.v_index == cur.v_index && prev.position == cur.position && prev.symbol == cur.symbol),
            "duplicate (position, v_index, symbol) container key"
         );
      }

      auto column = std::make_unique<HostColumn>();
      const K1Geometry& geometry = k1Geometry();
      column->segment_pieces = geometry.segmentPieces();
      require(static_cast<uint64_t>(in->n_symbols) * in->genome_length <= UINT32_MAX, "counts array exceeds 32-bit indices");
      // Every stored container is cut into PIECES of at most PIECE_BYTES of payload, each with its own
      // 16-byte descriptor: counts are additive, so a piece is an independent unit of work and a warp
      // never sits on a multi-kilobyte container while its CTA waits. Arrays of one or two values
      // carry their values inside the descriptor and own no payload at all.
      std::vector<DevContainer> descs;
      descs.reserve(in->n_containers + in->payload_bytes / PIECE_BYTES + 16);
      std::vector<uint32_t> desc_weights;  // bytes a sweep over the piece reads
      desc_weights.reserve(descs.capacity());
      std::vector<DevSegment> segments;
      std::vector<uint32_t> chunk_desc_begin(n_chunks + 1, 0);
      std::vector<uint32_t> chunk_seg_begin(n_chunks + 1, 0);
      std::vector<uint8_t> slab;
      slab.reserve(in->payload_bytes + in->payload_bytes / 16 + 4096);
      column->chunk_desc_payload_bytes.assign(n_chunks, 0);
      column->chunk_containers.assign(n_chunks, 0);

      struct PendingPiece {
         DevContainer desc;
         size_t byte_offset;
         uint32_t bytes;
         uint32_t cost_class;
      };
      std::vector<PendingPiece> chunk_pieces;
      std::vector<uint8_t> chunk_bytes;
      std::vector<uint32_t> piece_order;
      std::vector<uint8_t> scratch;
      std::vector<uint32_t> word_entries;
      std::vector<uint32_t> word_ranges;
      uint64_t cursor = 0;
      for (uint32_t chunk = 0; chunk < n_chunks; ++chunk) {
         chunk_desc_begin[chunk] = static_cast<uint32_t>(descs.size());
         chunk_seg_begin[chunk] = static_cast<uint32_t>(segments.size());
         // A segment is ONE block of the slab: [its descriptors, 16 bytes each | the payloads they point at],
         // so that the container kernel's producer moves it into a ring stage with a single bulk copy.
         std::vector<uint32_t> segment_pieces;   // indices into chunk_pieces
         std::vector<uint32_t> segment_offsets;  // payload offset of each piece behind the descriptors
         uint32_t segment_payload = 0;
         uint32_t segment_kind = 0;
         auto closeSegment = [&]() {
            if (segment_pieces.empty()) {
               return;
            }
            slab.resize((slab.size() + 15) / 16 * 16, 0);
            const uint64_t block_start = slab.size();
            const uint64_t payload_start = block_start + sizeof(DevContainer) * segment_pieces.size();
            require((payload_start + segment_payload) / 4 <= UINT32_MAX, "column payload exceeds the 16 GiB per-shard addressing limit");
            slab.resize(payload_start + segment_payload, 0);
            for (size_t i = 0; i < segment_pieces.size(); ++i) {
               PendingPiece& piece = chunk_pieces[segment_pieces[i]];
               piece.desc.offset4 = static_cast<uint32_t>((payload_start + segment_offsets[i]) / 4);
               // the block's copy of the descriptor names the piece's slot of the counts array instead of its position
               DevContainer block_desc = piece.desc;
               block_desc.position = piece.desc.symbol() * in->genome_length + piece.desc.position;
               std::memcpy(slab.data() + block_start + sizeof(DevContainer) * i, &block_desc, sizeof(DevContainer));
               if (piece.bytes > 0) {
                  std::memcpy(slab.data() + payload_start + segment_offsets[i], chunk_bytes.data() + piece.byte_offset, piece.bytes);
               }
            }
            DevSegment segment{};
            segment.payload_offset16 = static_cast<uint32_t>(block_start / 16);
            segment.bytes_and_kind = static_cast<uint32_t>(payload_start + segment_payload - block_start) | (segment_kind << 24);
            segment.reserved = 0;
            segment.chunk_and_count = chunk | (static_cast<uint32_t>(segment_pieces.size()) << 16);
            segments.push_back(segment);
            segment_pieces.clear();
            segment_offsets.clear();
            segment_payload = 0;
         };
         // The pieces of the chunk are first collected in (position, symbol) order -- the order of `descs`, which
         // the filter interpreter searches -- and then laid out for the container kernel (placePiece below).
         chunk_pieces.clear();
         chunk_bytes.clear();
         auto emitPiece = [&](const silo_container_desc& c, uint32_t kind, uint32_t cardinality, uint32_t aux,
                              const uint8_t* src, uint32_t bytes) {
            PendingPiece piece{};
            piece.desc.position = c.position;
            piece.desc.packed = DevContainer::pack(cardinality, c.symbol, in->local_reference[c.position], kind);
            piece.desc.aux = aux;
            piece.byte_offset = chunk_bytes.size();
            piece.bytes = bytes;
            // container-kernel order: by kind (a ring stage holds pieces of ONE kind), two-region pieces first
            piece.cost_class = stageClass(kind, kind == KIND_BITSET || bytes > 512);
            if (bytes > 0) {
               chunk_bytes.insert(chunk_bytes.end(), src, src + bytes);
            }
            chunk_pieces.push_back(piece);
         };
         // appends one piece to the open segment (container-kernel order), closing it first when it is full
         auto placePiece = [&](uint32_t index) {
            const uint32_t padded = (chunk_pieces[index].bytes + 15) / 16 * 16;
            const uint32_t stage_class = chunk_pieces[index].cost_class;
            const uint32_t capacity = chunk_pieces[index].desc.type() == KIND_INLINE ? geometry.inlinePieces() : geometry.segmentPieces();
            if (segment_pieces.size() == capacity || segment_payload + padded > geometry.segmentPayloadBytes() ||
                (!segment_pieces.empty() && stage_class != segment_kind)) {
               closeSegment();
            }
            segment_kind = stage_class;
            segment_pieces.push_back(index);
            segment_offsets.push_back(segment_payload);
            segment_payload += padded;
         };
         while (cursor < order.size() && in->containers[order[cursor]].v_index == first_chunk + chunk) {
            const silo_container_desc& c = in->containers[order[cursor]];
            const uint8_t* src = in->payload + c.payload_offset;
            if (c.typecode == TYPE_ARRAY) {
               if (c.cardinality <= 2) {
                  uint16_t values[2] = {0, 0};
                  std::memcpy(values, src, 2 * c.cardinality);
                  emitPiece(c, KIND_INLINE, c.cardinality, values[0] | (static_cast<uint32_t>(values[1]) << 16), nullptr, 0);
               } else {
                  constexpr uint32_t VALUES_PER_PIECE = PIECE_BYTES / 2;
                  for (uint32_t first = 0; first < c.cardinality; first += VALUES_PER_PIECE) {
                     const uint32_t count = std::min(VALUES_PER_PIECE, c.cardinality - first);
                     encodeArrayPiece(reinterpret_cast<const uint8_t*>(src + 2ULL * first), count, scratch);
                     emitPiece(c, KIND_ARRAY_T, count, 0, scratch.data(), static_cast<uint32_t>(scratch.size()));
                  }
               }
            } else if (c.typecode == TYPE_RUN) {
               uint16_t n_runs = 0;
               std::memcpy(&n_runs, src, 2);
               uint64_t covered = 0;
               splitRuns(src + 2, n_runs, word_entries, word_ranges, &covered);
               require(covered == c.cardinality, "run container cardinality does not match its runs");
               constexpr uint32_t ENTRIES_PER_PIECE = PIECE_BYTES / 4;
               for (size_t first = 0; first < word_entries.size(); first += ENTRIES_PER_PIECE) {
                  const auto count = static_cast<uint32_t>(std::min<size_t>(ENTRIES_PER_PIECE, word_entries.size() - first));
                  uint32_t rows = 0;
                  for (uint32_t i = 0; i < count; ++i) {
                     rows += static_cast<uint32_t>(__builtin_popcount(runEntryMask(word_entries[first + i])));
                  }
                  encodeRunsPiece(word_entries.data() + first, count, scratch);
                  emitPiece(c, KIND_RUNS_W, rows, count, scratch.data(), static_cast<uint32_t>(scratch.size()));
               }
               for (size_t first = 0; first < word_ranges.size();) {
                  // a piece's row count must fit the descriptor's 16 bits: close it before 65536 rows
                  uint32_t rows = 0;
                  size_t end = first;
                  while (end < word_ranges.size() && end - first < ENTRIES_PER_PIECE) {
                     const uint32_t range_rows = 32u * ((word_ranges[end] >> 16) - (word_ranges[end] & 0xFFFFu));
                     if (rows + range_rows > 65536u) {
                        break;
                     }
                     rows += range_rows;
                     ++end;
                  }
                  scratch.assign(
                     reinterpret_cast<const uint8_t*>(word_ranges.data() + first),
                     reinterpret_cast<const uint8_t*>(word_ranges.data() + end)
                  );
                  emitPiece(c, KIND_WORDRANGE, rows, static_cast<uint32_t>(end - first), scratch.data(), static_cast<uint32_t>(scratch.size()));
                  first = end;
               }
            } else {
               constexpr uint32_t WORDS_PER_PIECE = PIECE_BYTES / 8;
               uint64_t covered = 0;
               for (uint32_t word = 0; word < TILE_WORDS; word += WORDS_PER_PIECE) {
                  uint32_t cardinality = 0;
                  for (uint32_t w = word; w < word + WORDS_PER_PIECE; ++w) {
                     uint64_t value = 0;
                     std::memcpy(&value, src + 8ULL * w, 8);
                     cardinality += static_cast<uint32_t>(__builtin_popcountll(value));
                  }
                  covered += cardinality;
                  if (cardinality != 0) {
                     emitPiece(c, KIND_BITSET, cardinality, word | (WORDS_PER_PIECE << 16), src + 8ULL * word, PIECE_BYTES);
                  }
               }
               require(covered == c.cardinality, "bitset container cardinality does not match its bits");
            }
            column->chunk_desc_payload_bytes[chunk] += sizeof(DevContainer) + c.payload_bytes;
            column->chunk_containers[chunk]++;
            ++cursor;
         }
         // Container-kernel order: a ring stage holds pieces of ONE kind (the kernel branches on the kind once per
         // stage visit, not once per piece) and of equal cost, so that the stage keeps its consumer warps busy for
         // the same time. Counts are sums: the order is free.
         piece_order.resize(chunk_pieces.size());
         std::iota(piece_order.begin(), piece_order.end(), 0u);
         std::stable_sort(piece_order.begin(), piece_order.end(), [&](uint32_t a, uint32_t b) {
            return chunk_pieces[a].cost_class > chunk_pieces[b].cost_class;
         });
         for (const uint32_t index : piece_order) {
            placePiece(index);
         }
         closeSegment();
         for (const PendingPiece& piece : chunk_pieces) {
            descs.push_back(piece.desc);  // (position, symbol) order, same payload offsets
            desc_weights.push_back(static_cast<uint32_t>(sizeof(DevContainer)) + piece.bytes);
         }
      }
      require(descs.size() <= UINT32_MAX, "too many container pieces for 32-bit descriptor indices");
      chunk_desc_begin[n_chunks] = static_cast<uint32_t>(descs.size());
      chunk_seg_begin[n_chunks] = static_cast<uint32_t>(segments.size());
      slab.resize((slab.size() + 15) / 16 * 16 + 16, 0);

      // ---- threshold sweep: one contiguous range of pieces per CTA (one CTA per SM), equal bytes ----
      column->sweep_ctas = static_cast<uint32_t>(std::max(1, table->ctx->sm_count));
      std::vector<uint32_t> sweep_split(column->sweep_ctas + 1, static_cast<uint32_t>(descs.size()));
      std::vector<uint32_t> sweep_flushes(n_chunks, 0);
      {
         uint64_t total_weight = 0;
         for (uint32_t weight : desc_weights) {
            total_weight += weight;
         }
         uint64_t running = 0;
         uint32_t next_cta = 0;
         for (size_t i = 0; i < descs.size(); ++i) {
            while (next_cta < column->sweep_ctas && running * column->sweep_ctas >= static_cast<uint64_t>(next_cta) * total_weight) {
               sweep_split[next_cta++] = static_cast<uint32_t>(i);
            }
            running += desc_weights[i];
         }
         sweep_split[0] = 0;
         for (uint32_t cta = 0; cta < column->sweep_ctas; ++cta) {
            for (uint32_t chunk = 0; chunk < n_chunks; ++chunk) {
               const uint32_t lo = std::max(sweep_split[cta], chunk_desc_begin[chunk]);
               const uint32_t hi = std::min(sweep_split[cta + 1], chunk_desc_begin[chunk + 1]);
               if (lo < hi) {
                  sweep_flushes[chunk]++;
               }
            }
         }
         for (uint32_t count : sweep_flushes) {
            column->sweep_max_flushes = std::max(column->sweep_max_flushes, count);
         }
      }

      // ---- coverage ----
      std::vector<uint32_t> chunk_row_begin(n_chunks + 1, 0);
      for (uint32_t chunk = 0; chunk < n_chunks; ++chunk) {
         chunk_row_begin[chunk + 1] = chunk_row_begin[chunk] + table->chunk_sizes[chunk];
      }
      std::vector<uint2> start_end(table->n_rows);
      for (uint64_t row = 0; row < table->n_rows; ++row) {
         start_end[row] = make_uint2(in->start_end[2 * row], in->start_end[2 * row + 1]);
         require(start_end[row].x <= start_end[row].y && start_end[row].y <= in->genome_length, "coverage range out of bounds");
      }
      std::vector<uint32_t> chunk_missing_begin(n_chunks + 1, 0);
      std::vector<uint16_t> missing_row(in->n_rows_with_missing);
      std::vector<uint64_t> missing_offsets(in->n_rows_with_missing + 1, 0);
      column->chunk_missing_rows.assign(n_chunks, 0);
      uint64_t total_runs = 0;
      {
         uint32_t chunk = 0;
         for (uint64_t i = 0; i < in->n_rows_with_missing; ++i) {
            const uint32_t row_id = in->missing_row_ids[i];
            require(i == 0 || in->missing_row_ids[i - 1] < row_id, "missing_row_ids must be strictly ascending");
            const uint32_t global_chunk = row_id >> 16;
            require(global_chunk >= first_chunk && global_chunk < first_chunk + n_chunks, "missing row outside the shard");
            const uint32_t local_chunk = global_chunk - first_chunk;
            require((row_id & 0xFFFF) < table->chunk_sizes[local_chunk], "missing row outside the row layout");
            while (chunk < local_chunk) {
               chunk_missing_begin[++chunk] = static_cast<uint32_t>(i);
            }
            missing_row[i] = static_cast<uint16_t>(row_id & 0xFFFF);
            require(in->missing_offsets[i] <= in->missing_offsets[i + 1], "missing_offsets must be ascending");
            missing_offsets[i] = in->missing_offsets[i];
            column->chunk_missing_rows[local_chunk]++;
         }
         while (chunk < n_chunks) {
            chunk_missing_begin[++chunk] = static_cast<uint32_t>(in->n_rows_with_missing);
         }
         if (in->n_rows_with_missing > 0) {
            require(in->missing_offsets[0] == 0, "missing_offsets[0] must be 0");
            total_runs = in->missing_offsets[in->n_rows_with_missing];
            missing_offsets[in->n_rows_with_missing] = total_runs;
         }
      }
      std::vector<uint2> missing_runs(total_runs);
      for (uint64_t run = 0; run < total_runs; ++run) {
         missing_runs[run] = make_uint2(in->missing_runs[2 * run], in->missing_runs[2 * run + 1]);
         require(missing_runs[run].x < missing_runs[run].y && missing_runs[run].y <= in->genome_length, "missing run out of bounds");
      }
      // nulls as dense tiles (only when present)
      std::vector<uint64_t> null_words;
      if (in->n_null_rows > 0) {
         null_words.assign(static_cast<size_t>(n_chunks) * TILE_WORDS, 0);
         for (uint64_t i = 0; i < in->n_null_rows; ++i) {
            const uint32_t row_id = in->null_row_ids[i];
            const uint32_t global_chunk = row_id >> 16;
            require(global_chunk >= first_chunk && global_chunk < first_chunk + n_chunks, "null row outside the shard");
            const uint32_t local_chunk = global_chunk - first_chunk;
            require((row_id & 0xFFFF) < table->chunk_sizes[local_chunk], "null row outside the row layout");
            null_words[static_cast<size_t>(local_chunk) * TILE_WORDS + ((row_id & 0xFFFF) >> 6)] |= 1ULL << (row_id & 63);
         }
      }

      // ---- upload ----
      uint64_t* acc = &column->device_bytes;
      auto track = [&](auto* ptr) {
         column->allocations.push_back(const_cast<void*>(static_cast<const void*>(ptr)));
         return ptr;
      };
      DevColumn& dev = column->dev;
      dev.n_symbols = in->n_symbols;
      dev.genome_length = in->genome_length;
      dev.missing_symbol = in->missing_symbol;
      dev.n_chunks = n_chunks;
      dev.n_containers = descs.size();  // pieces
      dev.n_segments = static_cast<uint32_t>(segments.size());
      std::vector<uint8_t> local_reference(in->local_reference, in->local_reference + in->genome_length);
      dev.local_reference = track(deviceUpload(local_reference, stream, acc));
      dev.containers = track(deviceUpload(descs, stream, acc));
      dev.payload = track(deviceUpload(slab, stream, acc));
      dev.chunk_desc_begin = track(deviceUpload(chunk_desc_begin, stream, acc));
      dev.segments = track(deviceUpload(segments, stream, acc));
      dev.chunk_seg_begin = track(deviceUpload(chunk_seg_begin, stream, acc));
      dev.start_end = track(deviceUpload(start_end, stream, acc));
      dev.chunk_row_begin = track(deviceUpload(chunk_row_begin, stream, acc));
      dev.chunk_missing_begin = track(deviceUpload(chunk_missing_begin, stream, acc));
      dev.missing_row = track(deviceUpload(missing_row, stream, acc));
      dev.missing_offsets = track(deviceUpload(missing_offsets, stream, acc));
      dev.missing_runs = track(deviceUpload(missing_runs, stream, acc));
      dev.null_words = null_words.empty() ? nullptr : track(deviceUpload(null_words, stream, acc));
      column->d_sweep_split = track(deviceUpload(sweep_split, stream, acc));
      column->d_sweep_flushes = track(deviceUpload(sweep_flushes, stream, acc));
      SILO_CUDA_CHECK(cudaStreamSynchronize(stream));

      if (dev.n_segments > table->work_items_capacity) {
         cudaFree(table->d_work_items);
         table->d_work_items = deviceAlloc<DevSegment>(dev.n_segments, &table->device_bytes);
         table->work_items_capacity = dev.n_segments;
      }
      if (diffWords(in->genome_length) > table->coverage_diff_capacity) {
         // (all-zero between queries, see silo_gpu_table; a fresh one starts all-zero too)
         SILO_CUDA_CHECK(cudaStreamSynchronize(table->ctx->stream));
         cudaFree(table->d_coverage_diff);
         table->d_coverage_diff = deviceAlloc<uint32_t>(diffWords(in->genome_length), &table->device_bytes);
         SILO_CUDA_CHECK(cudaMemset(table->d_coverage_diff, 0, diffWords(in->genome_length) * sizeof(uint32_t)));
         table->coverage_diff_capacity = diffWords(in->genome_length);
      }
      const uint64_t counts_elems = static_cast<uint64_t>(in->n_symbols) * in->genome_length;
      if (counts_elems > table->counts_capacity) {
         cudaFree(table->d_counts);
         if (table->h_counts_pinned != nullptr) {
            cudaFreeHost(table->h_counts_pinned);
         }
         table->d_counts = deviceAlloc<uint32_t>(counts_elems, &table->device_bytes);
         SILO_CUDA_CHECK(cudaMallocHost(&table->h_counts_pinned, counts_elems * sizeof(uint32_t)));
         table->counts_capacity = counts_elems;
      }
      dropQueryGraphsLocked(table);  // the staged column table and the scratch buffers may have changed
      table->columns.push_back(column.release());
      column_index = static_cast<int>(table->columns.size()) - 1;
   });
   return status == SILO_OK ? column_index : status;
}

int silo_gpu_get_stats(silo_gpu_table* table, silo_gpu_stats* out) {
   return guarded([&] {
      require(table != nullptr && out != nullptr, "silo_gpu_get_stats: NULL argument");
      std::lock_guard<std::mutex> lock(table->mutex);
      Stats& stats = table->stats;
      if (table->last_column >= 0) {
         SILO_CUDA_CHECK(cudaSetDevice(table->ctx->device));
         const HostColumn& host = *table->columns[static_cast<size_t>(table->last_column)];
         std::vector<uint32_t> popcounts(table->n_chunks);
         if (table->n_chunks > 0) {
            SILO_CUDA_CHECK(cudaMemcpyAsync(
               popcounts.data(), table->last_popcounts, popcounts.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost,
               table->last_stream
            ));
         }
         SILO_CUDA_CHECK(cudaStreamSynchronize(table->last_stream));
         stats.containers = 0;
         stats.counts_kernel_bytes = 0;
         stats.algorithmic_bytes = static_cast<uint64_t>(host.dev.n_symbols) * host.dev.genome_length * sizeof(uint32_t);
         for (uint32_t chunk = 0; chunk < table->n_chunks; ++chunk) {
            if (popcounts[chunk] == 0) {
               continue;
            }
            stats.containers += host.chunk_containers[chunk];
            // descriptors (+ payloads and the filter tile unless the filter is full)
            stats.counts_kernel_bytes += table->last_was_full
                                            ? host.chunk_containers[chunk] * sizeof(DevContainer)
                                            : host.chunk_desc_payload_bytes[chunk] + TILE_BYTES;
            // + 8 B per row of the chunk (start, end) + the missing-row index of the chunk
            stats.algorithmic_bytes += 8ULL * table->chunk_sizes[chunk] + 2ULL * host.chunk_missing_rows[chunk];
         }
         stats.algorithmic_bytes += stats.counts_kernel_bytes;
         // average over the calls recorded since the previous get_stats (at most the ring size)
         const uint64_t n_timed = std::min<uint64_t>(table->timed_calls, silo_gpu_table::EVENT_RING);
         double kernel_ms = 0;
         double total_ms = 0;
         uint64_t n_valid = 0;
         for (uint64_t back = 0; back < n_timed; ++back) {
            const uint64_t slot = (table->timed_calls - 1 - back) % silo_gpu_table::EVENT_RING;
            float k1_ms = 0;
            float call_ms = 0;
            // a call enqueued inside a stream capture (CUDA graph) leaves its events unrecorded: skipped
            if (cudaEventElapsedTime(&k1_ms, table->ev_k1_begin[slot], table->ev_k1_end[slot]) != cudaSuccess ||
                cudaEventElapsedTime(&call_ms, table->ev_begin[slot], table->ev_end[slot]) != cudaSuccess) {
               cudaGetLastError();
               continue;
            }
            kernel_ms += k1_ms;
            total_ms += call_ms;
            ++n_valid;
         }
         if (n_valid > 0) {
            stats.last_counts_kernel_ms = static_cast<float>(kernel_ms / static_cast<double>(n_valid));
            stats.last_total_ms = static_cast<float>(total_ms / static_cast<double>(n_valid));
         }
         stats.timed_calls = n_valid;
         table->timed_calls = 0;
      }
      out->containers = stats.containers;
      out->algorithmic_bytes = stats.algorithmic_bytes;
      out->counts_kernel_bytes = stats.counts_kernel_bytes;
      out->kernel_launches = stats.kernel_launches;
      out->last_counts_kernel_ms = stats.last_counts_kernel_ms;
      out->last_total_ms = stats.last_total_ms;
      out->timed_calls = stats.timed_calls;
   });
}

}  // extern "C"
