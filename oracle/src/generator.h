// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// Restates the synthetic-data generators of /root/reference/performance/sequence_generator.h so
// that the oracle can be fed the reference benchmarks' own inputs (same std::mt19937 seeds and
// libstdc++ distributions => same streams under g++ 13):
//   SequenceTreeGenerator :113-185 | writeFullSequenceNdjson :367-384
//   writeNRunSequenceNdjson :392-428 | buildMutationBenchmarkReference / writeMutationBenchmarkNdjson
//   :432-466 | makeCoOccurrenceReference / writeCoOccurrenceNdjson :487-526
#pragma once
#include <random>
#include <string>
#include <string_view>
#include <vector>

namespace oracle {

class SequenceTreeGenerator {
   std::mt19937 rng;
   const std::string& reference;
   double mutation_rate;
   double death_rate;
   size_t generations;
   size_t children_per_node;

   char mutateBase(char base) {
      static constexpr char SYMBOLS[4] = {'-', 'A', 'C', 'G'};  // Nucleotide::SYMBOLS.at(0..3): never T
      std::uniform_int_distribution<size_t> dist(0, 3);
      char new_base;
      do {
         new_base = SYMBOLS[dist(rng)];
      } while (new_base == base);
      return new_base;
   }

   std::string mutateSequence(std::string_view sequence) {
      std::string mutated{sequence};
      std::binomial_distribution<size_t> num_mutations_dist(sequence.size(), mutation_rate);
      const size_t num_mutations = num_mutations_dist(rng);
      std::uniform_int_distribution<size_t> pos_dist(0, sequence.size() - 1);
      for (size_t i = 0; i < num_mutations; ++i) {
         const size_t pos = pos_dist(rng);
         mutated[pos] = mutateBase(mutated[pos]);
      }
      return mutated;
   }

  public:
   SequenceTreeGenerator(
      const std::string& ref,
      uint64_t seed = 42,
      double mut_rate = 0.001,
      double death = 0.1,
      size_t gens = 5,
      size_t children = 3
   )
       : rng(seed),
         reference(ref),
         mutation_rate(mut_rate),
         death_rate(death),
         generations(gens),
         children_per_node(children) {}

   // parent[i] = index of the sequence i was mutated from (parent[0] = 0)
   std::vector<std::string> generateEvolvedSequences(std::vector<size_t>* parents = nullptr) {
      std::vector<std::string> all_sequences = {reference};
      if (parents != nullptr) {
         parents->assign(1, 0);
      }
      std::vector<size_t> current_gen = {0};
      std::bernoulli_distribution survives(1.0 - death_rate);
      for (size_t gen = 0; gen < generations; ++gen) {
         std::vector<size_t> next_gen;
         for (size_t seq_index : current_gen) {
            for (size_t child = 0; child < children_per_node; ++child) {
               if (survives(rng)) {
                  all_sequences.push_back(mutateSequence(all_sequences.at(seq_index)));
                  if (parents != nullptr) {
                     parents->push_back(seq_index);
                  }
                  next_gen.push_back(all_sequences.size() - 1);
               }
            }
         }
         if (next_gen.empty()) {
            next_gen.push_back(all_sequences.size() - 1);
         }
         current_gen = std::move(next_gen);
      }
      return all_sequences;
   }
};

inline std::string buildMutationBenchmarkReference() {
   std::string reference;
   for (size_t i = 0; i < 1000; ++i) {
      reference += "ACGT";
   }
   return reference;
}

inline std::string makeCoOccurrenceReference() {
   constexpr char bases[4] = {'A', 'C', 'G', 'T'};
   std::mt19937 rng{42};
   std::uniform_int_distribution<size_t> base_dist(0, 3);
   std::string reference(100, 'A');
   for (char& base : reference) {
      base = bases[base_dist(rng)];
   }
   return reference;
}

}  // namespace oracle
