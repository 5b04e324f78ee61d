// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into or called from the product path.
//
// CPU restatement of the 2^16-value roaring *container* that RhyDB/SILO's hot path is built on.
// The reference delegates this arithmetic to CRoaring 4.5.0 (`roaring::internal::container_*`,
// pinned in /root/reference/conanfile.py:22, NOT vendored in the tree), reached through
//   src/rhydb/roaring_util/roaring_container.h:23-158 (RoaringContainer: add / |= / runOptimize /
//   serialize) and roaring_container.cpp:7-128 (withCapacity, &, -, |).
// What is restated here is CRoaring's *published* semantics: three container kinds
//   BITSET (typecode 1, 1024 x u64), ARRAY (typecode 2, <= 4096 sorted u16), RUN (typecode 3,
//   {start, length-1} pairs), set algebra on [0, 2^16), and the `container_write` byte layout
//   (roaring_container.h:104-157). Container *kind* selection only affects bytes, never results.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace oracle {

constexpr uint8_t BITSET_CONTAINER_TYPE = 1;
constexpr uint8_t ARRAY_CONTAINER_TYPE = 2;
constexpr uint8_t RUN_CONTAINER_TYPE = 3;
constexpr int32_t DEFAULT_MAX_SIZE = 4096;  // roaring_container.cpp:11
constexpr size_t BITSET_WORDS = 1024;

struct Container {
   uint8_t type = ARRAY_CONTAINER_TYPE;
   uint32_t card = 0;
   // ARRAY: sorted distinct values. RUN: flattened {start, length-1} pairs, sorted, disjoint,
   // non-adjacent.
   std::vector<uint16_t> vals;
   // BITSET: exactly 1024 words.
   std::vector<uint64_t> words;

   // roaring_container.cpp:7-23
   static Container withCapacity(int32_t capacity);
   static Container fromWords(const uint64_t* src);  // array if card <= 4096 else bitset
   static Container fromRange(uint32_t begin, uint32_t end);  // [begin, end) as one run
   static Container fromSorted(const uint16_t* values, size_t count);

   // roaring_container.cpp:38-48 (container_add; the reference only adds absent values)
   void add(uint16_t value);
   [[nodiscard]] bool contains(uint16_t value) const;
   [[nodiscard]] bool empty() const { return card == 0; }
   [[nodiscard]] uint32_t numRuns() const;
   // CRoaring container_size_in_bytes == number of bytes container_write emits
   [[nodiscard]] size_t sizeInBytes() const;
   // roaring_container.cpp:53-62 (convert_run_optimize + shrink)
   void runOptimize();
   void toWords(uint64_t* dst) const;  // dst[1024], overwritten
   void write(uint8_t* dst) const;     // container_write layout
   static Container read(uint8_t typecode, uint32_t cardinality, const uint8_t* src, size_t len);

   template <typename Fn>
   void forEach(Fn&& fn) const {
      if (type == ARRAY_CONTAINER_TYPE) {
         for (uint16_t value : vals) {
            fn(value);
         }
      } else if (type == RUN_CONTAINER_TYPE) {
         for (size_t i = 0; i + 1 < vals.size(); i += 2) {
            const uint32_t start = vals[i];
            const uint32_t last = start + vals[i + 1];
            for (uint32_t value = start; value <= last; ++value) {
               fn(static_cast<uint16_t>(value));
            }
         }
      } else {
         for (size_t w = 0; w < BITSET_WORDS; ++w) {
            uint64_t word = words[w];
            while (word != 0) {
               fn(static_cast<uint16_t>(w * 64 + static_cast<size_t>(__builtin_ctzll(word))));
               word &= word - 1;
            }
         }
      }
   }
};

// container_and_cardinality — the Mutations hot-loop primitive (mutations_node.cpp:177-182).
uint32_t containerAndCardinality(const Container& lhs, const Container& rhs);
// container_and / container_andnot / container_or (copy_on_write_bitmap.cpp:159,194,231)
Container containerAnd(const Container& lhs, const Container& rhs);
Container containerAndNot(const Container& lhs, const Container& rhs);
Container containerOr(const Container& lhs, const Container& rhs);
// complement of `c` within [begin, end) only; values outside are kept (Roaring::flip semantics,
// row_layout.cpp:18-23)
Container containerFlipRange(const Container& c, uint32_t begin, uint32_t end);

}  // namespace oracle
