// ORACLE — TEST INFRASTRUCTURE ONLY. See container.h for provenance.
#include "container.h"

#include <algorithm>
#include <stdexcept>

namespace oracle {

namespace {

inline uint32_t popcountWords(const uint64_t* words) {
   uint32_t total = 0;
   for (size_t i = 0; i < BITSET_WORDS; ++i) {
      total += static_cast<uint32_t>(__builtin_popcountll(words[i]));
   }
   return total;
}

// number of set bits of `words` in [start, start + lenminusone]  (bitset_lenrange_cardinality)
inline uint32_t lenrangeCardinality(const uint64_t* words, uint32_t start, uint32_t lenminusone) {
   const uint32_t last = start + lenminusone;  // inclusive
   const uint32_t firstword = start / 64;
   const uint32_t endword = last / 64;
   const uint64_t head_mask = ~UINT64_C(0) << (start % 64);
   const uint64_t tail_mask = ~UINT64_C(0) >> (63 - last % 64);
   if (firstword == endword) {
      return static_cast<uint32_t>(__builtin_popcountll(words[firstword] & head_mask & tail_mask));
   }
   uint32_t answer = static_cast<uint32_t>(__builtin_popcountll(words[firstword] & head_mask));
   for (uint32_t i = firstword + 1; i < endword; i++) {
      answer += static_cast<uint32_t>(__builtin_popcountll(words[i]));
   }
   answer += static_cast<uint32_t>(__builtin_popcountll(words[endword] & tail_mask));
   return answer;
}

inline void setRange(uint64_t* words, uint32_t begin, uint32_t end) {  // [begin, end)
   if (begin >= end) {
      return;
   }
   const uint32_t last = end - 1;
   const uint32_t firstword = begin / 64;
   const uint32_t endword = last / 64;
   const uint64_t head_mask = ~UINT64_C(0) << (begin % 64);
   const uint64_t tail_mask = ~UINT64_C(0) >> (63 - last % 64);
   if (firstword == endword) {
      words[firstword] |= head_mask & tail_mask;
      return;
   }
   words[firstword] |= head_mask;
   for (uint32_t i = firstword + 1; i < endword; i++) {
      words[i] = ~UINT64_C(0);
   }
   words[endword] |= tail_mask;
}

uint32_t arrayArrayAndCardinality(const std::vector<uint16_t>& a, const std::vector<uint16_t>& b) {
   // CRoaring switches to a galloping ("skewed") intersection at a 64x size ratio.
   const std::vector<uint16_t>& small = a.size() <= b.size() ? a : b;
   const std::vector<uint16_t>& large = a.size() <= b.size() ? b : a;
   if (small.empty()) {
      return 0;
   }
   uint32_t count = 0;
   if (small.size() * 64 < large.size()) {
      auto from = large.begin();
      for (uint16_t value : small) {
         from = std::lower_bound(from, large.end(), value);
         if (from == large.end()) {
            break;
         }
         if (*from == value) {
            ++count;
         }
      }
      return count;
   }
   size_t i = 0;
   size_t j = 0;
   while (i < small.size() && j < large.size()) {
      if (small[i] < large[j]) {
         ++i;
      } else if (small[i] > large[j]) {
         ++j;
      } else {
         ++count;
         ++i;
         ++j;
      }
   }
   return count;
}

uint32_t arrayRunAndCardinality(const std::vector<uint16_t>& arr, const std::vector<uint16_t>& runs) {
   uint32_t count = 0;
   size_t run = 0;
   const size_t n_runs = runs.size() / 2;
   for (uint16_t value : arr) {
      while (run < n_runs && static_cast<uint32_t>(runs[2 * run]) + runs[2 * run + 1] < value) {
         ++run;
      }
      if (run == n_runs) {
         break;
      }
      if (value >= runs[2 * run]) {
         ++count;
      }
   }
   return count;
}

uint32_t runRunAndCardinality(const std::vector<uint16_t>& a, const std::vector<uint16_t>& b) {
   uint32_t count = 0;
   size_t i = 0;
   size_t j = 0;
   const size_t na = a.size() / 2;
   const size_t nb = b.size() / 2;
   while (i < na && j < nb) {
      const uint32_t a_start = a[2 * i];
      const uint32_t a_end = a_start + a[2 * i + 1] + 1;
      const uint32_t b_start = b[2 * j];
      const uint32_t b_end = b_start + b[2 * j + 1] + 1;
      const uint32_t lo = std::max(a_start, b_start);
      const uint32_t hi = std::min(a_end, b_end);
      if (lo < hi) {
         count += hi - lo;
      }
      if (a_end <= b_end) {
         ++i;
      } else {
         ++j;
      }
   }
   return count;
}

}  // namespace

Container Container::withCapacity(int32_t capacity) {
   Container result;
   if (capacity <= DEFAULT_MAX_SIZE) {
      result.type = ARRAY_CONTAINER_TYPE;
      result.vals.reserve(static_cast<size_t>(std::max(capacity, 0)));
   } else {
      result.type = BITSET_CONTAINER_TYPE;
      result.words.assign(BITSET_WORDS, 0);
   }
   return result;
}

Container Container::fromWords(const uint64_t* src) {
   Container result;
   result.card = popcountWords(src);
   if (result.card <= static_cast<uint32_t>(DEFAULT_MAX_SIZE)) {
      result.type = ARRAY_CONTAINER_TYPE;
      result.vals.reserve(result.card);
      for (size_t w = 0; w < BITSET_WORDS; ++w) {
         uint64_t word = src[w];
         while (word != 0) {
            result.vals.push_back(static_cast<uint16_t>(w * 64 + static_cast<size_t>(__builtin_ctzll(word)))
            );
            word &= word - 1;
         }
      }
   } else {
      result.type = BITSET_CONTAINER_TYPE;
      result.words.assign(src, src + BITSET_WORDS);
   }
   return result;
}

Container Container::fromRange(uint32_t begin, uint32_t end) {
   Container result;
   if (begin >= end) {
      return result;
   }
   result.type = RUN_CONTAINER_TYPE;
   result.card = end - begin;
   result.vals = {static_cast<uint16_t>(begin), static_cast<uint16_t>(end - begin - 1)};
   return result;
}

Container Container::fromSorted(const uint16_t* values, size_t count) {
   Container result = withCapacity(static_cast<int32_t>(count));
   if (result.type == ARRAY_CONTAINER_TYPE) {
      result.vals.assign(values, values + count);
   } else {
      for (size_t i = 0; i < count; ++i) {
         result.words[values[i] >> 6] |= UINT64_C(1) << (values[i] & 63);
      }
   }
   result.card = static_cast<uint32_t>(count);
   return result;
}

void Container::add(uint16_t value) {
   if (type == ARRAY_CONTAINER_TYPE) {
      if (vals.empty() || vals.back() < value) {
         if (card >= static_cast<uint32_t>(DEFAULT_MAX_SIZE)) {
            // array_container_try_add refuses beyond DEFAULT_MAX_SIZE -> convert to bitset
            words.assign(BITSET_WORDS, 0);
            for (uint16_t existing : vals) {
               words[existing >> 6] |= UINT64_C(1) << (existing & 63);
            }
            vals.clear();
            vals.shrink_to_fit();
            type = BITSET_CONTAINER_TYPE;
            words[value >> 6] |= UINT64_C(1) << (value & 63);
            card += 1;
            return;
         }
         vals.push_back(value);
         card += 1;
         return;
      }
      auto iter = std::lower_bound(vals.begin(), vals.end(), value);
      if (iter != vals.end() && *iter == value) {
         return;
      }
      if (card >= static_cast<uint32_t>(DEFAULT_MAX_SIZE)) {
         words.assign(BITSET_WORDS, 0);
         for (uint16_t existing : vals) {
            words[existing >> 6] |= UINT64_C(1) << (existing & 63);
         }
         vals.clear();
         type = BITSET_CONTAINER_TYPE;
         words[value >> 6] |= UINT64_C(1) << (value & 63);
         card += 1;
         return;
      }
      vals.insert(iter, value);
      card += 1;
      return;
   }
   if (type == BITSET_CONTAINER_TYPE) {
      const uint64_t bit = UINT64_C(1) << (value & 63);
      if ((words[value >> 6] & bit) == 0) {
         words[value >> 6] |= bit;
         card += 1;
      }
      return;
   }
   // RUN: rare on the ingest path; go through the dense form
   uint64_t tmp[BITSET_WORDS];
   toWords(tmp);
   tmp[value >> 6] |= UINT64_C(1) << (value & 63);
   *this = fromWords(tmp);
}

bool Container::contains(uint16_t value) const {
   if (type == ARRAY_CONTAINER_TYPE) {
      return std::binary_search(vals.begin(), vals.end(), value);
   }
   if (type == BITSET_CONTAINER_TYPE) {
      return ((words[value >> 6] >> (value & 63)) & 1U) != 0;
   }
   for (size_t i = 0; i + 1 < vals.size(); i += 2) {
      if (value < vals[i]) {
         return false;
      }
      if (static_cast<uint32_t>(value) <= static_cast<uint32_t>(vals[i]) + vals[i + 1]) {
         return true;
      }
   }
   return false;
}

uint32_t Container::numRuns() const {
   if (type == RUN_CONTAINER_TYPE) {
      return static_cast<uint32_t>(vals.size() / 2);
   }
   if (type == ARRAY_CONTAINER_TYPE) {
      uint32_t runs = 0;
      int32_t prev = -2;
      for (uint16_t value : vals) {
         if (static_cast<int32_t>(value) != prev + 1) {
            ++runs;
         }
         prev = value;
      }
      return runs;
   }
   uint32_t runs = 0;
   uint64_t carry = 0;  // bit 63 of the previous word
   for (size_t w = 0; w < BITSET_WORDS; ++w) {
      const uint64_t word = words[w];
      // run starts: bit set and previous bit clear
      runs += static_cast<uint32_t>(__builtin_popcountll(word & ~((word << 1) | carry)));
      carry = word >> 63;
   }
   return runs;
}

size_t Container::sizeInBytes() const {
   if (type == BITSET_CONTAINER_TYPE) {
      return BITSET_WORDS * 8;
   }
   if (type == ARRAY_CONTAINER_TYPE) {
      return static_cast<size_t>(card) * 2;
   }
   return 2 + 2 * vals.size();
}

void Container::runOptimize() {
   // convert_run_optimize [external: CRoaring]: a non-run container becomes a run container iff
   // the run form serialises strictly smaller; a run container falls back to array/bitset iff
   // that is strictly smaller than the run form.
   const uint32_t n_runs = numRuns();
   const size_t size_as_run = 2 + 4 * static_cast<size_t>(n_runs);
   const size_t size_as_array = static_cast<size_t>(card) * 2 + 2;
   const size_t size_as_bitset = BITSET_WORDS * 8;
   if (type == RUN_CONTAINER_TYPE) {
      const size_t min_non_run = std::min(size_as_array, size_as_bitset);
      if (size_as_run <= min_non_run) {
         vals.shrink_to_fit();
         return;
      }
      uint64_t tmp[BITSET_WORDS];
      toWords(tmp);
      *this = fromWords(tmp);
      return;
   }
   const size_t current = type == ARRAY_CONTAINER_TYPE ? size_as_array : size_as_bitset;
   if (size_as_run >= current) {
      vals.shrink_to_fit();
      return;
   }
   std::vector<uint16_t> runs;
   runs.reserve(2 * static_cast<size_t>(n_runs));
   int32_t run_start = -1;
   int32_t prev = -2;
   forEach([&](uint16_t value) {
      if (static_cast<int32_t>(value) != prev + 1) {
         if (run_start >= 0) {
            runs.push_back(static_cast<uint16_t>(run_start));
            runs.push_back(static_cast<uint16_t>(prev - run_start));
         }
         run_start = value;
      }
      prev = value;
   });
   if (run_start >= 0) {
      runs.push_back(static_cast<uint16_t>(run_start));
      runs.push_back(static_cast<uint16_t>(prev - run_start));
   }
   type = RUN_CONTAINER_TYPE;
   vals = std::move(runs);
   words.clear();
   words.shrink_to_fit();
}

void Container::toWords(uint64_t* dst) const {
   if (type == BITSET_CONTAINER_TYPE) {
      std::memcpy(dst, words.data(), BITSET_WORDS * 8);
      return;
   }
   std::memset(dst, 0, BITSET_WORDS * 8);
   if (type == ARRAY_CONTAINER_TYPE) {
      for (uint16_t value : vals) {
         dst[value >> 6] |= UINT64_C(1) << (value & 63);
      }
      return;
   }
   for (size_t i = 0; i + 1 < vals.size(); i += 2) {
      setRange(dst, vals[i], static_cast<uint32_t>(vals[i]) + vals[i + 1] + 1);
   }
}

void Container::write(uint8_t* dst) const {
   if (type == BITSET_CONTAINER_TYPE) {
      std::memcpy(dst, words.data(), BITSET_WORDS * 8);
   } else if (type == ARRAY_CONTAINER_TYPE) {
      std::memcpy(dst, vals.data(), vals.size() * 2);
   } else {
      const auto n_runs = static_cast<uint16_t>(vals.size() / 2);
      std::memcpy(dst, &n_runs, 2);
      std::memcpy(dst + 2, vals.data(), vals.size() * 2);
   }
}

Container Container::read(uint8_t typecode, uint32_t cardinality, const uint8_t* src, size_t len) {
   Container result;
   result.type = typecode;
   result.card = cardinality;
   if (typecode == BITSET_CONTAINER_TYPE) {
      if (len != BITSET_WORDS * 8) {
         throw std::runtime_error("bitset container payload must be 8192 bytes");
      }
      result.words.resize(BITSET_WORDS);
      std::memcpy(result.words.data(), src, len);
   } else if (typecode == ARRAY_CONTAINER_TYPE) {
      if (len != static_cast<size_t>(cardinality) * 2) {
         throw std::runtime_error("array container payload must be 2*cardinality bytes");
      }
      result.vals.resize(cardinality);
      std::memcpy(result.vals.data(), src, len);
   } else if (typecode == RUN_CONTAINER_TYPE) {
      uint16_t n_runs = 0;
      if (len < 2) {
         throw std::runtime_error("run container payload too short");
      }
      std::memcpy(&n_runs, src, 2);
      if (len != 2 + 4 * static_cast<size_t>(n_runs)) {
         throw std::runtime_error("run container payload must be 2+4*n_runs bytes");
      }
      result.vals.resize(2 * static_cast<size_t>(n_runs));
      std::memcpy(result.vals.data(), src + 2, 4 * static_cast<size_t>(n_runs));
   } else {
      throw std::runtime_error("unknown roaring container typecode");
   }
   return result;
}

uint32_t containerAndCardinality(const Container& lhs, const Container& rhs) {
   const Container* a = &lhs;
   const Container* b = &rhs;
   if (a->type > b->type) {
      std::swap(a, b);
   }
   // now a->type <= b->type with BITSET(1) < ARRAY(2) < RUN(3)
   if (a->type == BITSET_CONTAINER_TYPE) {
      if (b->type == BITSET_CONTAINER_TYPE) {
         uint32_t total = 0;
         for (size_t i = 0; i < BITSET_WORDS; ++i) {
            total += static_cast<uint32_t>(__builtin_popcountll(a->words[i] & b->words[i]));
         }
         return total;
      }
      if (b->type == ARRAY_CONTAINER_TYPE) {
         uint32_t total = 0;
         for (uint16_t value : b->vals) {
            total += static_cast<uint32_t>((a->words[value >> 6] >> (value & 63)) & 1U);
         }
         return total;
      }
      uint32_t total = 0;
      for (size_t i = 0; i + 1 < b->vals.size(); i += 2) {
         total += lenrangeCardinality(a->words.data(), b->vals[i], b->vals[i + 1]);
      }
      return total;
   }
   if (a->type == ARRAY_CONTAINER_TYPE) {
      if (b->type == ARRAY_CONTAINER_TYPE) {
         return arrayArrayAndCardinality(a->vals, b->vals);
      }
      return arrayRunAndCardinality(a->vals, b->vals);
   }
   return runRunAndCardinality(a->vals, b->vals);
}

namespace {
Container arrayArrayOp(const Container& lhs, const Container& rhs, int op /*0 and,1 andnot,2 or*/) {
   std::vector<uint16_t> out;
   const auto& a = lhs.vals;
   const auto& b = rhs.vals;
   size_t i = 0;
   size_t j = 0;
   while (i < a.size() && j < b.size()) {
      if (a[i] < b[j]) {
         if (op != 0) {
            out.push_back(a[i]);
         }
         ++i;
      } else if (a[i] > b[j]) {
         if (op == 2) {
            out.push_back(b[j]);
         }
         ++j;
      } else {
         if (op != 1) {
            out.push_back(a[i]);
         }
         ++i;
         ++j;
      }
   }
   if (op != 0) {
      out.insert(out.end(), a.begin() + static_cast<ptrdiff_t>(i), a.end());
   }
   if (op == 2) {
      out.insert(out.end(), b.begin() + static_cast<ptrdiff_t>(j), b.end());
   }
   return Container::fromSorted(out.data(), out.size());
}

Container denseOp(const Container& lhs, const Container& rhs, int op) {
   uint64_t a[BITSET_WORDS];
   uint64_t b[BITSET_WORDS];
   lhs.toWords(a);
   rhs.toWords(b);
   for (size_t i = 0; i < BITSET_WORDS; ++i) {
      a[i] = op == 0 ? (a[i] & b[i]) : (op == 1 ? (a[i] & ~b[i]) : (a[i] | b[i]));
   }
   return Container::fromWords(a);
}
}  // namespace

Container containerAnd(const Container& lhs, const Container& rhs) {
   if (lhs.type == ARRAY_CONTAINER_TYPE && rhs.type == ARRAY_CONTAINER_TYPE) {
      return arrayArrayOp(lhs, rhs, 0);
   }
   if (lhs.type == ARRAY_CONTAINER_TYPE || rhs.type == ARRAY_CONTAINER_TYPE) {
      const Container& arr = lhs.type == ARRAY_CONTAINER_TYPE ? lhs : rhs;
      const Container& other = lhs.type == ARRAY_CONTAINER_TYPE ? rhs : lhs;
      std::vector<uint16_t> out;
      for (uint16_t value : arr.vals) {
         if (other.contains(value)) {
            out.push_back(value);
         }
      }
      return Container::fromSorted(out.data(), out.size());
   }
   return denseOp(lhs, rhs, 0);
}

Container containerAndNot(const Container& lhs, const Container& rhs) {
   if (lhs.type == ARRAY_CONTAINER_TYPE && rhs.type == ARRAY_CONTAINER_TYPE) {
      return arrayArrayOp(lhs, rhs, 1);
   }
   if (lhs.type == ARRAY_CONTAINER_TYPE) {
      std::vector<uint16_t> out;
      for (uint16_t value : lhs.vals) {
         if (!rhs.contains(value)) {
            out.push_back(value);
         }
      }
      return Container::fromSorted(out.data(), out.size());
   }
   return denseOp(lhs, rhs, 1);
}

Container containerOr(const Container& lhs, const Container& rhs) {
   if (lhs.type == ARRAY_CONTAINER_TYPE && rhs.type == ARRAY_CONTAINER_TYPE &&
       lhs.card + rhs.card <= static_cast<uint32_t>(DEFAULT_MAX_SIZE)) {
      return arrayArrayOp(lhs, rhs, 2);
   }
   return denseOp(lhs, rhs, 2);
}

Container containerFlipRange(const Container& c, uint32_t begin, uint32_t end) {
   uint64_t tmp[BITSET_WORDS];
   c.toWords(tmp);
   uint64_t mask[BITSET_WORDS];
   std::memset(mask, 0, sizeof(mask));
   setRange(mask, begin, end);
   for (size_t i = 0; i < BITSET_WORDS; ++i) {
      tmp[i] ^= mask[i];
   }
   Container result = Container::fromWords(tmp);
   result.runOptimize();  // roaring's negation returns the most compact form
   return result;
}

}  // namespace oracle
