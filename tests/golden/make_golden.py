"""Generates tests/golden/synthetic_small.json with the ORACLE (the CPU restatement of the reference
algorithm, pinned by tests/test_oracle_kat.py against the reference's own unit-test vectors).

The reference itself can be neither built nor imported in this image, so these
vectors are oracle outputs on a seeded synthetic table, not reference outputs; they freeze the
behaviour that the known-answer tests pinned, so that a later change of the oracle, of the
generators or of the device path shows up as a diff of a committed file.

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from oracle import oracle as O  # noqa: E402

GENOME_LENGTH = 400
ROWS = 2 * 65536 + 4321
GENERATIONS = 5
SEED = 20260101

EXPRESSIONS = [
    "(true)",
    "(bitmap lineage)",
    "(not (bitmap lineage))",
    "(and (ranges 1000 60000 65536 131072 131072 133000) (bitmap lineage))",
    "(has-mut main 17)",
    "(or (sym-eq main 5 A) (sym-eq main 5 C) (sym-eq main 5 G) (sym-eq main 5 T))",
    "(n-of 2 0 (has-mut main 3) (has-mut main 40) (has-mut main 77) (sym-eq main 100 -))",
    "(n-of 1 1 (has-mut main 3) (has-mut main 40) (has-mut main 77))",
    "(profile main 3 muts)",
    "(maybe (sym-eq main 9 R))",
]


def build():
    rng = np.random.default_rng(SEED)
    reference = "".join("ACGT"[int(i)] for i in rng.integers(0, 4, GENOME_LENGTH))
    table = O.full_sequence_table(reference, ROWS, GENERATIONS)
    evolved, parents = O.gen_evolved(reference, seed=42, generations=GENERATIONS)
    n = len(evolved)
    generation = [0] * n
    for e in range(1, n):
        generation[e] = generation[parents[e]] + 1
    ancestor = next(e for e in range(n) if generation[e] == 2)
    in_lineage = np.zeros(n, dtype=bool)
    in_lineage[ancestor] = True
    for e in range(ancestor + 1, n):
        in_lineage[e] = in_lineage[parents[e]]
    table.register_bitmap("lineage", np.flatnonzero(in_lineage[np.arange(ROWS) % n]).tolist())
    return reference, table


def digest(array: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(array).tobytes()).hexdigest()


def main():
    reference, table = build()
    cases = []
    for expression in EXPRESSIONS:
        flt = table.filter(expression)
        counts = table.mutation_counts("main", flt)
        cases.append({
            "expression": expression,
            "cardinality": int(flt.cardinality),
            "ids_sha256": digest(flt.ids().astype(np.uint32)),
            "first_ids": [int(v) for v in flt.ids()[:8]],
            "counts_sha256": digest(counts.astype(np.uint32)),
            "counts_column_sums_head": [int(v) for v in counts.sum(axis=0)[:6]],
            "rows_min_proportion_0.05": table.mutation_rows("main", counts, 0.05)[:12],
            "n_rows_min_proportion_0.05": len(table.mutation_rows("main", counts, 0.05)),
            "n_rows_min_proportion_0": len(table.mutation_rows("main", counts, 0.0)),
        })
    out = {
        "generator": "tests/golden/make_golden.py (oracle outputs; see its docstring)",
        "genome_length": GENOME_LENGTH, "rows": ROWS, "generations": GENERATIONS, "reference_seed": SEED,
        "reference_sha256": hashlib.sha256(reference.encode()).hexdigest(),
        "local_reference_sha256": hashlib.sha256(table.local_reference("main").encode()).hexdigest(),
        "num_containers": int(table.num_containers("main")),
        "cases": cases,
    }
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "synthetic_small.json")
    with open(path, "w") as handle:
        json.dump(out, handle, indent=1)
    print("wrote", path, "with", len(cases), "cases")


if __name__ == "__main__":
    main()
