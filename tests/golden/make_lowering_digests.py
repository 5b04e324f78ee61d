"""Golden digests of LOWERED filter programs (FNV-1a over the instruction fields and the blob, silo_host_filter_lower_timed)
for a fixed list of expressions on a host-only table: a refactor of the host mirror (expressions / operators / lowering)
must not change what the device is given. The digests were produced by the build whose GPU parity suite was green on
B200 (commit f446028 and HEAD at the time gave identical values); regenerate only together with a GPU run:
    python tests/golden/make_lowering_digests.py"""
import ctypes as C
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from lapis_silo_b200 import abi, host_api as H  # noqa: E402

NUC = "-ACGTRYSWKMBDHVN"
REFERENCE = "ACGTACGTACGTACGTACGTACGTACGTAC"
LOCAL = "ACGTACGTTCGTACGAACGTACGTACGTCC"  # three positions where the local reference differs from the global one
LINEAGE = bytes.fromhex("3a300000010000000000010010000000" + "01000500")  # portable roaring {1, 5}


def table(null_rows: bool):
    t = H.HostTable(None, [65536, 65536, 65536, 700])
    d = abi.ColumnDesc()
    d.struct_size = C.sizeof(abi.ColumnDesc)
    d.n_symbols, d.genome_length, d.missing_symbol = 16, len(REFERENCE), 15
    ids = (C.c_uint8 * len(LOCAL))(*[NUC.index(c) for c in LOCAL])
    d.local_reference = C.cast(ids, C.POINTER(C.c_uint8))
    nulls = (C.c_uint32 * 1)(3)
    if null_rows:
        d.n_null_rows = 1
        d.null_row_ids = C.cast(nulls, C.POINTER(C.c_uint32))
    t.add_column("c", H.NUCLEOTIDE, REFERENCE, C.pointer(d))
    t._keep = (d, ids, nulls)
    t.register_bitmap("lineage", LINEAGE, False)
    return t


EXPRESSIONS = [
    "(true)", "(false)", "(bitmap lineage)", "(not (bitmap lineage))", "(bitmap nosuch)", "(has-mut c 7)", "(not (has-mut c 9))",
    "(and (bitmap lineage) (not (sym-eq c 12 N)))", "(ranges 3 90 196608 197000)", "(profile c 4 muts)", "(profile c 2 muts 3 T 9 N 15 R)",
    "(profile c 0 seq " + REFERENCE + ")", "(profile c 3 seq " + LOCAL.replace("A", "R", 2) + ")",
    "(sym-eq c 9 .)", "(sym-eq c 9 T)", "(sym-eq c 3 -)", "(maybe (sym-eq c 4 R))", "(exact (sym-eq c 4 R))", "(maybe (has-mut c 5))",
    "(exact (has-mut c 5))", "(or (sym-eq c 3 A) (sym-eq c 3 G) (sym-eq c 4 T))",
    "(or (has-mut c 1) (and (sym-eq c 2 C) (not (sym-eq c 16 A))))",
    "(n-of 2 0 (sym-eq c 1 C) (sym-eq c 2 A) (not (sym-eq c 3 T)) (bitmap lineage))", "(n-of 1 1 (has-mut c 1) (has-mut c 2) (sym-eq c 9 N))",
    "(n-of 3 0 (has-mut c 1) (has-mut c 2) (has-mut c 3))", "(n-of 0 1 (has-mut c 1) (not (has-mut c 2)))", "(n-of 1 0 (true) (has-mut c 2))",
    "(not (or (sym-eq c 9 N) (ranges 0 10)))", "(and (ranges 5 100) (has-mut c 20) (not (sym-eq c 21 .)) (sym-eq c 22 N))",
    "(and (maybe (profile c 1 muts 2 Y)) (not (exact (n-of 2 0 (sym-eq c 1 R) (sym-eq c 2 C) (has-mut c 30)))))",
]


def lowered():
    out = []
    for null_rows in (False, True):
        t = table(null_rows)
        for expression in EXPRESSIONS:
            r = t.lower_timed(expression)
            out.append({"null_rows": null_rows, "expression": expression, "n_instrs": r["n_instrs"], "blob_bytes": r["blob_bytes"],
                        "n_bitmaps": r["n_bitmaps"], "digest": str(r["digest"])})
        t.close()
    return out


if __name__ == "__main__":
    with open(os.path.join(HERE, "lowering_digests.json"), "w") as handle:
        json.dump(lowered(), handle, indent=1)
    print("wrote", 2 * len(EXPRESSIONS), "digests")
