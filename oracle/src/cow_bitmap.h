// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// Restates rhydb::query_engine::CopyOnWriteBitmap
// (/root/reference/src/rhydb/query_engine/copy_on_write_bitmap.h:28-147, .cpp:20-364): the return
// type of every filter Operator::evaluate(). Sorted 2^16 keys, each container either a non-owning
// view into an index or a privately owned result.
#pragma once
#include <cstdint>
#include <memory>
#include <utility>
#include <vector>

#include "container.h"
#include "roaring.h"

namespace oracle {

class CowBitmap {
   struct Slot {
      const Container* view = nullptr;   // non-owning (copy_on_write_bitmap.h:31-32)
      std::unique_ptr<Container> owned;  // owning

      Slot() = default;
      explicit Slot(const Container* view) : view(view) {}
      explicit Slot(Container&& container)
          : owned(std::make_unique<Container>(std::move(container))) {}
      Slot(const Slot& other)  // copyContainer, .cpp:45-58: views stay views, owned is cloned
          : view(other.view),
            owned(other.owned ? std::make_unique<Container>(*other.owned) : nullptr) {}
      Slot& operator=(const Slot& other) {
         if (this != &other) {
            view = other.view;
            owned = other.owned ? std::make_unique<Container>(*other.owned) : nullptr;
         }
         return *this;
      }
      Slot(Slot&&) noexcept = default;
      Slot& operator=(Slot&&) noexcept = default;
      [[nodiscard]] const Container& get() const { return owned ? *owned : *view; }
   };

   std::vector<uint16_t> keys;
   std::vector<Slot> slots;

   void pushIfNonEmpty(uint16_t key, Container&& container);  // .cpp:20-35

  public:
   CowBitmap() = default;
   explicit CowBitmap(const Roaring* bitmap);  // views, .cpp:60-73
   explicit CowBitmap(Roaring&& bitmap);       // owns, .cpp:75-92

   [[nodiscard]] uint64_t cardinality() const;  // .cpp:108-114
   [[nodiscard]] bool isEmpty() const { return keys.empty(); }
   [[nodiscard]] uint64_t andCardinality(const CowBitmap& other) const;  // .cpp:120-143

   CowBitmap& operator&=(const CowBitmap& other);  // .cpp:145-176
   CowBitmap& operator-=(const CowBitmap& other);  // .cpp:178-211
   CowBitmap& operator|=(const CowBitmap& other);  // .cpp:213-248
   [[nodiscard]] CowBitmap operator&(const CowBitmap& other) const;
   [[nodiscard]] CowBitmap operator-(const CowBitmap& other) const;

   static CowBitmap fastUnion(const std::vector<CowBitmap>& bitmaps);  // .cpp:286-297
   static CowBitmap fromContainerViews(std::vector<std::pair<uint16_t, const Container*>> views
   );                                         // .cpp:299-349
   [[nodiscard]] Roaring toRoaring() const;  // .cpp:352-364

   [[nodiscard]] size_t size() const { return keys.size(); }
   [[nodiscard]] uint16_t keyAt(size_t idx) const { return keys[idx]; }
   [[nodiscard]] const Container& containerAt(size_t idx) const { return slots[idx].get(); }
};

}  // namespace oracle
