// ORACLE — TEST INFRASTRUCTURE ONLY. See roaring.h for provenance.
#include "roaring.h"

#include <algorithm>
#include <stdexcept>

namespace oracle {

namespace {
constexpr uint32_t SERIAL_COOKIE_NO_RUNCONTAINER = 12346;
constexpr uint32_t SERIAL_COOKIE = 12347;
constexpr uint32_t NO_OFFSET_THRESHOLD = 4;

template <typename Op>
Roaring mergeOp(const Roaring& lhs, const Roaring& rhs, bool keep_left_only, bool keep_right_only, Op op) {
   Roaring out;
   size_t i = 0;
   size_t j = 0;
   while (i < lhs.keys.size() || j < rhs.keys.size()) {
      if (j >= rhs.keys.size() || (i < lhs.keys.size() && lhs.keys[i] < rhs.keys[j])) {
         if (keep_left_only) {
            out.keys.push_back(lhs.keys[i]);
            out.containers.push_back(lhs.containers[i]);
         }
         ++i;
      } else if (i >= lhs.keys.size() || lhs.keys[i] > rhs.keys[j]) {
         if (keep_right_only) {
            out.keys.push_back(rhs.keys[j]);
            out.containers.push_back(rhs.containers[j]);
         }
         ++j;
      } else {
         Container result = op(lhs.containers[i], rhs.containers[j]);
         if (!result.empty()) {
            out.keys.push_back(lhs.keys[i]);
            out.containers.push_back(std::move(result));
         }
         ++i;
         ++j;
      }
   }
   return out;
}
}  // namespace

int Roaring::findKey(uint16_t key) const {
   auto iter = std::lower_bound(keys.begin(), keys.end(), key);
   if (iter == keys.end() || *iter != key) {
      return -1;
   }
   return static_cast<int>(iter - keys.begin());
}

Container& Roaring::getOrCreate(uint16_t key) {
   auto iter = std::lower_bound(keys.begin(), keys.end(), key);
   const auto idx = static_cast<size_t>(iter - keys.begin());
   if (iter == keys.end() || *iter != key) {
      keys.insert(iter, key);
      containers.insert(containers.begin() + static_cast<ptrdiff_t>(idx), Container{});
   }
   return containers[idx];
}

void Roaring::dropEmpty() {
   size_t out = 0;
   for (size_t i = 0; i < keys.size(); ++i) {
      if (!containers[i].empty()) {
         if (out != i) {
            keys[out] = keys[i];
            containers[out] = std::move(containers[i]);
         }
         ++out;
      }
   }
   keys.resize(out);
   containers.resize(out);
}

Roaring Roaring::fromIds(const uint32_t* ids, size_t count) {
   std::vector<uint32_t> sorted(ids, ids + count);
   std::sort(sorted.begin(), sorted.end());
   sorted.erase(std::unique(sorted.begin(), sorted.end()), sorted.end());
   Roaring result;
   size_t i = 0;
   while (i < sorted.size()) {
      const auto key = static_cast<uint16_t>(sorted[i] >> 16);
      std::vector<uint16_t> lows;
      while (i < sorted.size() && (sorted[i] >> 16) == key) {
         lows.push_back(static_cast<uint16_t>(sorted[i] & 0xFFFF));
         ++i;
      }
      result.keys.push_back(key);
      result.containers.push_back(Container::fromSorted(lows.data(), lows.size()));
   }
   return result;
}

void Roaring::add(uint32_t value) {
   getOrCreate(static_cast<uint16_t>(value >> 16)).add(static_cast<uint16_t>(value & 0xFFFF));
}

void Roaring::addRange(uint64_t begin, uint64_t end) {
   while (begin < end) {
      const auto key = static_cast<uint16_t>(begin >> 16);
      const uint64_t block_end = std::min<uint64_t>(end, (static_cast<uint64_t>(key) + 1) << 16);
      const Container range = Container::fromRange(
         static_cast<uint32_t>(begin & 0xFFFF), static_cast<uint32_t>(block_end - (begin & ~UINT64_C(0xFFFF)))
      );
      Container& target = getOrCreate(key);
      target = target.empty() ? range : containerOr(target, range);
      if (target.type != RUN_CONTAINER_TYPE) {
         target.runOptimize();
      }
      begin = block_end;
   }
}

void Roaring::removeRange(uint64_t begin, uint64_t end) {
   end = std::min<uint64_t>(end, UINT64_C(1) << 32);
   while (begin < end) {
      const auto key = static_cast<uint16_t>(begin >> 16);
      const uint64_t block_end = std::min<uint64_t>(end, (static_cast<uint64_t>(key) + 1) << 16);
      const int idx = findKey(key);
      if (idx >= 0) {
         const Container range = Container::fromRange(
            static_cast<uint32_t>(begin & 0xFFFF),
            static_cast<uint32_t>(block_end - (begin & ~UINT64_C(0xFFFF)))
         );
         containers[static_cast<size_t>(idx)] = containerAndNot(containers[static_cast<size_t>(idx)], range);
      }
      // jump over key gaps quickly
      if (idx < 0) {
         auto iter = std::lower_bound(keys.begin(), keys.end(), key);
         if (iter == keys.end()) {
            break;
         }
         begin = std::max<uint64_t>(block_end, static_cast<uint64_t>(*iter) << 16);
      } else {
         begin = block_end;
      }
   }
   dropEmpty();
}

void Roaring::remove(uint32_t value) {
   const int idx = findKey(static_cast<uint16_t>(value >> 16));
   if (idx < 0) {
      return;
   }
   Container& target = containers[static_cast<size_t>(idx)];
   if (!target.contains(static_cast<uint16_t>(value & 0xFFFF))) {
      return;
   }
   const auto low = static_cast<uint16_t>(value & 0xFFFF);
   target = containerAndNot(target, Container::fromSorted(&low, 1));
   dropEmpty();
}

void Roaring::flip(uint64_t begin, uint64_t end) {
   while (begin < end) {
      const auto key = static_cast<uint16_t>(begin >> 16);
      const uint64_t block_end = std::min<uint64_t>(end, (static_cast<uint64_t>(key) + 1) << 16);
      Container& target = getOrCreate(key);
      target = containerFlipRange(
         target,
         static_cast<uint32_t>(begin & 0xFFFF),
         static_cast<uint32_t>(block_end - (begin & ~UINT64_C(0xFFFF)))
      );
      begin = block_end;
   }
   dropEmpty();
}

bool Roaring::contains(uint32_t value) const {
   const int idx = findKey(static_cast<uint16_t>(value >> 16));
   return idx >= 0 && containers[static_cast<size_t>(idx)].contains(static_cast<uint16_t>(value & 0xFFFF));
}

uint64_t Roaring::cardinality() const {
   uint64_t total = 0;
   for (const auto& container : containers) {
      total += container.card;
   }
   return total;
}

uint32_t Roaring::minimum() const {
   if (keys.empty()) {
      return UINT32_MAX;
   }
   uint32_t low = 0;
   bool found = false;
   containers[0].forEach([&](uint16_t value) {
      if (!found) {
         low = value;
         found = true;
      }
   });
   return (static_cast<uint32_t>(keys[0]) << 16) | low;
}

void Roaring::runOptimize() {
   for (auto& container : containers) {
      container.runOptimize();
   }
}

Roaring& Roaring::operator|=(const Roaring& other) {
   *this = mergeOp(*this, other, true, true, containerOr);
   return *this;
}
Roaring& Roaring::operator&=(const Roaring& other) {
   *this = mergeOp(*this, other, false, false, containerAnd);
   return *this;
}
Roaring& Roaring::operator-=(const Roaring& other) {
   *this = mergeOp(*this, other, true, false, containerAndNot);
   return *this;
}
Roaring Roaring::operator&(const Roaring& other) const {
   return mergeOp(*this, other, false, false, containerAnd);
}
Roaring Roaring::operator-(const Roaring& other) const {
   return mergeOp(*this, other, true, false, containerAndNot);
}
Roaring Roaring::operator|(const Roaring& other) const {
   return mergeOp(*this, other, true, true, containerOr);
}

bool Roaring::operator==(const Roaring& other) const {
   return toVector() == other.toVector();
}

std::vector<uint32_t> Roaring::toVector() const {
   std::vector<uint32_t> out;
   out.reserve(cardinality());
   forEach([&](uint32_t value) { out.push_back(value); });
   return out;
}

std::vector<uint8_t> Roaring::write() const {
   const auto size = static_cast<uint32_t>(keys.size());
   bool has_run = false;
   for (const auto& container : containers) {
      has_run = has_run || container.type == RUN_CONTAINER_TYPE;
   }
   std::vector<uint8_t> out;
   auto put = [&](const void* src, size_t len) {
      const auto* bytes = static_cast<const uint8_t*>(src);
      out.insert(out.end(), bytes, bytes + len);
   };
   if (has_run) {
      const uint32_t cookie = SERIAL_COOKIE | ((size - 1) << 16);
      put(&cookie, 4);
      std::vector<uint8_t> run_flags((size + 7) / 8, 0);
      for (uint32_t i = 0; i < size; ++i) {
         if (containers[i].type == RUN_CONTAINER_TYPE) {
            run_flags[i / 8] |= static_cast<uint8_t>(1U << (i % 8));
         }
      }
      put(run_flags.data(), run_flags.size());
   } else {
      put(&SERIAL_COOKIE_NO_RUNCONTAINER, 4);
      put(&size, 4);
   }
   // bitset containers with <= 4096 values are written as arrays: the reader infers the kind
   // from the cardinality.
   auto writtenAsArray = [](const Container& c) {
      return c.type == ARRAY_CONTAINER_TYPE ||
             (c.type == BITSET_CONTAINER_TYPE && c.card <= static_cast<uint32_t>(DEFAULT_MAX_SIZE));
   };
   for (uint32_t i = 0; i < size; ++i) {
      const uint16_t key = keys[i];
      const auto card_minus_one = static_cast<uint16_t>(containers[i].card - 1);
      put(&key, 2);
      put(&card_minus_one, 2);
   }
   if (!has_run || size >= NO_OFFSET_THRESHOLD) {
      auto offset = static_cast<uint32_t>(out.size() + 4 * static_cast<size_t>(size));
      for (uint32_t i = 0; i < size; ++i) {
         put(&offset, 4);
         const Container& c = containers[i];
         if (c.type == RUN_CONTAINER_TYPE) {
            offset += static_cast<uint32_t>(2 + 2 * c.vals.size());
         } else if (writtenAsArray(c)) {
            offset += 2 * c.card;
         } else {
            offset += BITSET_WORDS * 8;
         }
      }
   }
   for (uint32_t i = 0; i < size; ++i) {
      const Container& c = containers[i];
      if (c.type == BITSET_CONTAINER_TYPE && writtenAsArray(c)) {
         std::vector<uint16_t> values;
         c.forEach([&](uint16_t value) { values.push_back(value); });
         put(values.data(), values.size() * 2);
      } else {
         std::vector<uint8_t> buffer(c.sizeInBytes());
         c.write(buffer.data());
         put(buffer.data(), buffer.size());
      }
   }
   return out;
}

Roaring Roaring::read(const uint8_t* data, size_t len) {
   size_t pos = 0;
   auto get = [&](void* dst, size_t count) {
      if (pos + count > len) {
         throw std::runtime_error("roaring portable format: ran out of bytes");
      }
      std::memcpy(dst, data + pos, count);
      pos += count;
   };
   uint32_t cookie = 0;
   get(&cookie, 4);
   uint32_t size = 0;
   std::vector<uint8_t> run_flags;
   bool has_run = false;
   if ((cookie & 0xFFFF) == SERIAL_COOKIE) {
      has_run = true;
      size = (cookie >> 16) + 1;
      run_flags.resize((size + 7) / 8);
      get(run_flags.data(), run_flags.size());
   } else if (cookie == SERIAL_COOKIE_NO_RUNCONTAINER) {
      get(&size, 4);
   } else {
      throw std::runtime_error("roaring portable format: bad cookie");
   }
   if (size > 65536) {
      throw std::runtime_error("roaring portable format: too many containers");
   }
   Roaring result;
   std::vector<uint32_t> cards(size);
   for (uint32_t i = 0; i < size; ++i) {
      uint16_t key = 0;
      uint16_t card_minus_one = 0;
      get(&key, 2);
      get(&card_minus_one, 2);
      result.keys.push_back(key);
      cards[i] = static_cast<uint32_t>(card_minus_one) + 1;
   }
   if (!has_run || size >= NO_OFFSET_THRESHOLD) {
      pos += 4 * static_cast<size_t>(size);
   }
   for (uint32_t i = 0; i < size; ++i) {
      const bool is_run = has_run && ((run_flags[i / 8] >> (i % 8)) & 1U) != 0;
      if (is_run) {
         uint16_t n_runs = 0;
         if (pos + 2 > len) {
            throw std::runtime_error("roaring portable format: ran out of bytes");
         }
         std::memcpy(&n_runs, data + pos, 2);
         const size_t bytes = 2 + 4 * static_cast<size_t>(n_runs);
         if (pos + bytes > len) {
            throw std::runtime_error("roaring portable format: ran out of bytes");
         }
         result.containers.push_back(Container::read(RUN_CONTAINER_TYPE, cards[i], data + pos, bytes));
         pos += bytes;
      } else if (cards[i] <= static_cast<uint32_t>(DEFAULT_MAX_SIZE)) {
         const size_t bytes = 2 * static_cast<size_t>(cards[i]);
         if (pos + bytes > len) {
            throw std::runtime_error("roaring portable format: ran out of bytes");
         }
         result.containers.push_back(Container::read(ARRAY_CONTAINER_TYPE, cards[i], data + pos, bytes));
         pos += bytes;
      } else {
         const size_t bytes = BITSET_WORDS * 8;
         if (pos + bytes > len) {
            throw std::runtime_error("roaring portable format: ran out of bytes");
         }
         result.containers.push_back(Container::read(BITSET_CONTAINER_TYPE, cards[i], data + pos, bytes));
         pos += bytes;
      }
   }
   return result;
}

}  // namespace oracle
