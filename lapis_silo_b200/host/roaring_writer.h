// Minimal writer of the portable Roaring format (the byte format roaring::Roaring::write emits and
// a PUSH_BITMAP leaf accepts: roaring_util/roaring_serialize.h:15-30, RoaringFormatSpec). Only used
// where this repo has to play the part of the unchanged host engine that would hand over a ready
// bitmap (synthetic lineage index in the generator, literal id sets in the test notation). Emits
// array (<= 4096 values) and bitset containers, no run containers.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace silo_host {

inline std::vector<uint8_t> writePortableRoaring(const std::vector<uint32_t>& sorted_unique_ids) {
   struct Block {
      uint16_t key;
      size_t begin;
      size_t end;
   };
   std::vector<Block> blocks;
   for (size_t i = 0; i < sorted_unique_ids.size();) {
      const auto key = static_cast<uint16_t>(sorted_unique_ids[i] >> 16);
      size_t j = i;
      while (j < sorted_unique_ids.size() && (sorted_unique_ids[j] >> 16) == key) {
         ++j;
      }
      blocks.push_back({key, i, j});
      i = j;
   }
   std::vector<uint8_t> out;
   auto put = [&](const void* data, size_t bytes) {
      const auto* src = static_cast<const uint8_t*>(data);
      out.insert(out.end(), src, src + bytes);
   };
   const uint32_t cookie = 12346;  // SERIAL_COOKIE_NO_RUNCONTAINER
   const auto size = static_cast<uint32_t>(blocks.size());
   put(&cookie, 4);
   put(&size, 4);
   for (const Block& block : blocks) {
      const auto cardinality_minus_one = static_cast<uint16_t>(block.end - block.begin - 1);
      put(&block.key, 2);
      put(&cardinality_minus_one, 2);
   }
   auto offset = static_cast<uint32_t>(out.size() + 4 * blocks.size());
   for (const Block& block : blocks) {
      put(&offset, 4);
      const size_t cardinality = block.end - block.begin;
      offset += cardinality <= 4096 ? static_cast<uint32_t>(2 * cardinality) : 8192u;
   }
   for (const Block& block : blocks) {
      const size_t cardinality = block.end - block.begin;
      if (cardinality <= 4096) {
         for (size_t i = block.begin; i < block.end; ++i) {
            const auto low = static_cast<uint16_t>(sorted_unique_ids[i] & 0xFFFF);
            put(&low, 2);
         }
      } else {
         std::vector<uint64_t> words(1024, 0);
         for (size_t i = block.begin; i < block.end; ++i) {
            const uint32_t low = sorted_unique_ids[i] & 0xFFFF;
            words[low >> 6] |= uint64_t{1} << (low & 63);
         }
         put(words.data(), 8192);
      }
   }
   return out;
}

}  // namespace silo_host
