// ORACLE — TEST INFRASTRUCTURE ONLY. See cow_bitmap.h for provenance.
#include "cow_bitmap.h"

#include <algorithm>

namespace oracle {

void CowBitmap::pushIfNonEmpty(uint16_t key, Container&& container) {
   if (container.empty()) {
      return;
   }
   keys.push_back(key);
   slots.emplace_back(std::move(container));
}

CowBitmap::CowBitmap(const Roaring* bitmap) {
   keys = bitmap->keys;
   slots.reserve(bitmap->containers.size());
   for (const auto& container : bitmap->containers) {
      slots.emplace_back(&container);
   }
}

CowBitmap::CowBitmap(Roaring&& bitmap) {
   keys = std::move(bitmap.keys);
   slots.reserve(bitmap.containers.size());
   for (auto& container : bitmap.containers) {
      slots.emplace_back(std::move(container));
   }
   bitmap.keys.clear();
   bitmap.containers.clear();
}

uint64_t CowBitmap::cardinality() const {
   uint64_t total = 0;
   for (const auto& slot : slots) {
      total += slot.get().card;
   }
   return total;
}

uint64_t CowBitmap::andCardinality(const CowBitmap& other) const {
   uint64_t total = 0;
   size_t left = 0;
   size_t right = 0;
   while (left < keys.size() && right < other.keys.size()) {
      if (keys[left] < other.keys[right]) {
         ++left;
      } else if (keys[left] > other.keys[right]) {
         ++right;
      } else {
         total += containerAndCardinality(slots[left].get(), other.slots[right].get());
         ++left;
         ++right;
      }
   }
   return total;
}

CowBitmap& CowBitmap::operator&=(const CowBitmap& other) {
   CowBitmap result;
   size_t left = 0;
   size_t right = 0;
   while (left < keys.size() && right < other.keys.size()) {
      if (keys[left] < other.keys[right]) {
         ++left;
      } else if (keys[left] > other.keys[right]) {
         ++right;
      } else {
         result.pushIfNonEmpty(keys[left], containerAnd(slots[left].get(), other.slots[right].get()));
         ++left;
         ++right;
      }
   }
   *this = std::move(result);
   return *this;
}

CowBitmap& CowBitmap::operator-=(const CowBitmap& other) {
   CowBitmap result;
   size_t left = 0;
   size_t right = 0;
   while (left < keys.size()) {
      if (right >= other.keys.size() || keys[left] < other.keys[right]) {
         result.keys.push_back(keys[left]);
         result.slots.push_back(std::move(slots[left]));
         ++left;
      } else if (keys[left] > other.keys[right]) {
         ++right;
      } else {
         result.pushIfNonEmpty(
            keys[left], containerAndNot(slots[left].get(), other.slots[right].get())
         );
         ++left;
         ++right;
      }
   }
   *this = std::move(result);
   return *this;
}

CowBitmap& CowBitmap::operator|=(const CowBitmap& other) {
   CowBitmap result;
   size_t left = 0;
   size_t right = 0;
   while (left < keys.size() || right < other.keys.size()) {
      if (right >= other.keys.size() || (left < keys.size() && keys[left] < other.keys[right])) {
         result.keys.push_back(keys[left]);
         result.slots.push_back(std::move(slots[left]));
         ++left;
      } else if (left >= keys.size() || keys[left] > other.keys[right]) {
         result.keys.push_back(other.keys[right]);
         result.slots.push_back(other.slots[right]);  // copyContainer
         ++right;
      } else {
         result.pushIfNonEmpty(keys[left], containerOr(slots[left].get(), other.slots[right].get()));
         ++left;
         ++right;
      }
   }
   *this = std::move(result);
   return *this;
}

CowBitmap CowBitmap::operator&(const CowBitmap& other) const {
   CowBitmap result;
   size_t left = 0;
   size_t right = 0;
   while (left < keys.size() && right < other.keys.size()) {
      if (keys[left] < other.keys[right]) {
         ++left;
      } else if (keys[left] > other.keys[right]) {
         ++right;
      } else {
         result.pushIfNonEmpty(keys[left], containerAnd(slots[left].get(), other.slots[right].get()));
         ++left;
         ++right;
      }
   }
   return result;
}

CowBitmap CowBitmap::operator-(const CowBitmap& other) const {
   CowBitmap result = *this;
   result -= other;
   return result;
}

CowBitmap CowBitmap::fastUnion(const std::vector<CowBitmap>& bitmaps) {
   if (bitmaps.empty()) {
      return CowBitmap{};
   }
   CowBitmap result = bitmaps.front();
   for (size_t i = 1; i < bitmaps.size(); ++i) {
      result |= bitmaps[i];
   }
   return result;
}

CowBitmap CowBitmap::fromContainerViews(std::vector<std::pair<uint16_t, const Container*>> views) {
   std::erase_if(views, [](const auto& view) { return view.second->empty(); });
   std::stable_sort(views.begin(), views.end(), [](const auto& lhs, const auto& rhs) {
      return lhs.first < rhs.first;
   });
   CowBitmap result;
   size_t idx = 0;
   while (idx < views.size()) {
      const uint16_t key = views[idx].first;
      size_t group_end = idx + 1;
      while (group_end < views.size() && views[group_end].first == key) {
         ++group_end;
      }
      if (group_end - idx == 1) {
         result.keys.push_back(key);
         result.slots.emplace_back(views[idx].second);
      } else {
         Container accumulator = *views[idx].second;
         for (size_t i = idx + 1; i < group_end; ++i) {
            accumulator = containerOr(accumulator, *views[i].second);
         }
         result.pushIfNonEmpty(key, std::move(accumulator));
      }
      idx = group_end;
   }
   return result;
}

Roaring CowBitmap::toRoaring() const {
   Roaring result;
   result.keys = keys;
   result.containers.reserve(slots.size());
   for (const auto& slot : slots) {
      result.containers.push_back(slot.get());  // clone
   }
   return result;
}

}  // namespace oracle
