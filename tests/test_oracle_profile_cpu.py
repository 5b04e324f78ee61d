"""The oracle's Threshold DP (threshold.cpp:64-138 restated) against a brute-force Hamming count on synthetic tree
data: pins the checker that tests/test_gpu_parity.py::test_baseline_sizes_equal_the_oracle uses for the profile
distances the DP cannot reach at 10 M rows in reasonable time (O(n k) whole-bitmap passes). Run once at full size
in the build container (10 M rows, d = 0 / 5 / 50: identical row sets; 9 s / 59 s / 541 s of oracle time)."""
import numpy as np

from lapis_silo_b200 import host_api
from oracle import oracle as O


def test_profile_threshold_dp_equals_hamming_brute_force():
    total_rows, length = 200_000, 29903
    synthetic = host_api.Synthetic(genome_length=length, reference_seed=20200101, generations=5)
    sizes = host_api.dense_chunk_sizes(total_rows)
    table = O.Table()
    table.set_layout(*sizes)
    table.import_column("main", O.NUCLEOTIDE, synthetic.reference, synthetic.build_column(total_rows, 0, len(sizes), 4))
    synthetic.release_column()
    n = synthetic.num_sequences
    sequences = np.array([np.frombuffer(synthetic.sequence(e).encode(), dtype=np.uint8) for e in range(n)])
    query_index = n - 1
    distances = (sequences != sequences[query_index]).sum(axis=1)
    row_sequence = np.arange(total_rows) % n
    for distance in (0, 5, 35, 70):
        got = table.filter(f"(profile main {distance} seq {synthetic.sequence(query_index)})")
        want = np.flatnonzero((distances <= distance)[row_sequence]).astype(np.uint32)
        assert got.cardinality == len(want) > 0
        np.testing.assert_array_equal(got.ids(), want)
