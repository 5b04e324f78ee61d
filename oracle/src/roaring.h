// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// Whole-bitmap `roaring::Roaring` restatement ([external] CRoaring 4.5.0, conanfile.py:22): a sorted
// list of (high-16 key, container). Used by the reference for horizontal N bitmaps
// (horizontal_coverage_index.h:25), null_bitmap (sequence_column.h:111), the Threshold DP
// (threshold.cpp:73-137), Complement (complement.cpp:51-56) and RowLayout (row_layout.cpp:9-23).
// The portable serialisation restated here is the public RoaringFormatSpec
// (roaring_serialize.h:15-46 uses Roaring::write / readSafe).
#pragma once
#include <cstdint>
#include <vector>

#include "container.h"

namespace oracle {

class Roaring {
  public:
   std::vector<uint16_t> keys;
   std::vector<Container> containers;

   Roaring() = default;
   static Roaring fromIds(const uint32_t* ids, size_t count);  // any order, duplicates allowed

   void add(uint32_t value);
   void addRange(uint64_t begin, uint64_t end);  // [begin, end)
   void removeRange(uint64_t begin, uint64_t end);
   void remove(uint32_t value);
   void flip(uint64_t begin, uint64_t end);  // [begin, end)
   [[nodiscard]] bool contains(uint32_t value) const;
   [[nodiscard]] uint64_t cardinality() const;
   [[nodiscard]] bool isEmpty() const { return keys.empty(); }
   [[nodiscard]] uint32_t minimum() const;
   void runOptimize();

   Roaring& operator|=(const Roaring& other);
   Roaring& operator&=(const Roaring& other);
   Roaring& operator-=(const Roaring& other);
   [[nodiscard]] Roaring operator&(const Roaring& other) const;
   [[nodiscard]] Roaring operator-(const Roaring& other) const;
   [[nodiscard]] Roaring operator|(const Roaring& other) const;
   bool operator==(const Roaring& other) const;

   [[nodiscard]] std::vector<uint32_t> toVector() const;
   template <typename Fn>
   void forEach(Fn&& fn) const {
      for (size_t i = 0; i < keys.size(); ++i) {
         const uint32_t base = static_cast<uint32_t>(keys[i]) << 16;
         containers[i].forEach([&](uint16_t low) { fn(base | low); });
      }
   }

   // Portable format (RoaringFormatSpec). `write` returns the bytes.
   [[nodiscard]] std::vector<uint8_t> write() const;
   static Roaring read(const uint8_t* data, size_t len);

   [[nodiscard]] int findKey(uint16_t key) const;  // index or -1
   Container& getOrCreate(uint16_t key);
   void dropEmpty();
};

}  // namespace oracle
