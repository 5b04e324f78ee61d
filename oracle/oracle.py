"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/liboracle.so, the CPU restatement of the reference's bitmap filter +
Mutations path (see oracle/README.md and the file:line maps in oracle/src/*.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this module. The product package (lapis_silo_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Iterable, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    """Compile liboracle.so (g++ only; seconds)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.run(["make", "-s", "-C", _HERE, "-j8"], check=True)
    return _LIB_PATH


class ContainerDesc(C.Structure):  # silo_container_desc, include/silo_b200.h
    _fields_ = [
        ("position", C.c_uint32),
        ("v_index", C.c_uint16),
        ("symbol", C.c_uint8),
        ("typecode", C.c_uint8),
        ("cardinality", C.c_uint32),
        ("payload_bytes", C.c_uint32),
        ("payload_offset", C.c_uint64),
    ]


class ColumnDesc(C.Structure):  # silo_column_desc, include/silo_b200.h
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("n_symbols", C.c_uint32),
        ("genome_length", C.c_uint32),
        ("missing_symbol", C.c_uint32),
        ("local_reference", C.POINTER(C.c_uint8)),
        ("n_containers", C.c_uint64),
        ("containers", C.POINTER(ContainerDesc)),
        ("payload", C.POINTER(C.c_uint8)),
        ("payload_bytes", C.c_uint64),
        ("start_end", C.POINTER(C.c_uint32)),
        ("n_rows_with_missing", C.c_uint64),
        ("missing_row_ids", C.POINTER(C.c_uint32)),
        ("missing_offsets", C.POINTER(C.c_uint64)),
        ("missing_runs", C.POINTER(C.c_uint32)),
        ("n_null_rows", C.c_uint64),
        ("null_row_ids", C.POINTER(C.c_uint32)),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        L = _lib
        L.orc_last_error.restype = C.c_char_p
        L.orc_table_new.restype = C.c_void_p
        L.orc_table_free.argtypes = [C.c_void_p]
        L.orc_table_add_column.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_char_p]
        L.orc_table_set_layout.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.c_uint32]
        L.orc_table_append_row.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_uint32)]
        L.orc_table_append_cycled.argtypes = [
            C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_uint32), C.c_uint64, C.c_uint64]
        L.orc_table_flush_chunk.argtypes = [C.c_void_p]
        L.orc_table_finalize.argtypes = [C.c_void_p]
        L.orc_table_register_bitmap.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_uint32), C.c_uint64]
        L.orc_table_bitmap_bytes.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_uint8), C.c_uint64]
        L.orc_table_bitmap_bytes.restype = C.c_int64
        L.orc_table_num_rows.argtypes = [C.c_void_p]
        L.orc_table_num_rows.restype = C.c_uint32
        L.orc_table_num_chunks.argtypes = [C.c_void_p]
        L.orc_table_num_chunks.restype = C.c_uint32
        L.orc_table_chunk_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
        L.orc_column_local_reference.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.orc_column_num_containers.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_column_num_containers.restype = C.c_int64
        L.orc_filter_eval.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_filter_eval.restype = C.c_void_p
        L.orc_filter_free.argtypes = [C.c_void_p]
        L.orc_filter_cardinality.argtypes = [C.c_void_p]
        L.orc_filter_cardinality.restype = C.c_uint64
        L.orc_filter_ids.argtypes = [C.c_void_p]
        L.orc_filter_ids.restype = C.POINTER(C.c_uint32)
        L.orc_filter_words.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]
        L.orc_mutation_counts.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_uint32)]
        L.orc_mutation_rows.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_uint32), C.c_double]
        L.orc_mutation_rows.restype = C.c_void_p
        L.orc_mutation_rows_free.argtypes = [C.c_void_p]
        L.orc_mutation_rows_size.argtypes = [C.c_void_p]
        L.orc_mutation_rows_size.restype = C.c_uint64
        L.orc_mutation_rows_get.argtypes = [
            C.c_void_p, C.c_uint64, C.POINTER(C.c_char), C.POINTER(C.c_char), C.POINTER(C.c_int32),
            C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.orc_column_export.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_uint32]
        L.orc_column_export.restype = C.c_void_p
        L.orc_export_desc.argtypes = [C.c_void_p]
        L.orc_export_desc.restype = C.POINTER(ColumnDesc)
        L.orc_export_free.argtypes = [C.c_void_p]
        L.orc_table_import_column.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_char_p, C.c_void_p]
        L.orc_gen_evolved.argtypes = [
            C.c_char_p, C.c_uint64, C.c_double, C.c_double, C.c_uint64, C.c_uint64,
            C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.orc_gen_evolved.restype = C.c_int64
        L.orc_gen_full_sequence_table.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_uint64]
        L.orc_gen_nrun_table.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_uint64]
        L.orc_gen_mutation_benchmark_table.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_table_add_string_column.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_uint64]
        L.orc_table_add_string_column_ids.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_uint32, C.c_void_p, C.c_uint64]
        L.orc_table_add_date_column.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_now_seconds.restype = C.c_double
        L.orc_mutations_query.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_double, C.POINTER(C.c_uint64)]
        L.orc_mutations_query.restype = C.c_int64
        L.orc_bitmap_aggregation.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint64]
        L.orc_bitmap_aggregation.restype = C.c_int64
        L.orc_mutations_query_bench.argtypes = [
            C.c_void_p, C.c_char_p, C.c_char_p, C.c_double, C.c_uint32, C.c_double, C.c_uint64,
            C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
        L.orc_container_op.argtypes = [
            C.POINTER(C.c_uint16), C.c_uint32, C.c_int, C.POINTER(C.c_uint16), C.c_uint32, C.c_int,
            C.c_int, C.c_int, C.POINTER(C.c_uint16), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8),
            C.POINTER(C.c_uint8)]
        L.orc_roaring_roundtrip.argtypes = [
            C.POINTER(C.c_uint32), C.c_uint64, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint8), C.c_uint64]
        L.orc_roaring_roundtrip.restype = C.c_int64
        L.orc_extract.argtypes = [
            C.c_int, C.c_char_p, C.c_uint32, C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
            C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8),
            C.POINTER(C.c_uint32)]
        L.orc_column_containers_at.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint64]
        L.orc_column_containers_at.restype = C.c_int64
    return _lib


class OracleError(RuntimeError):
    pass


def _check(rc: int) -> None:
    if rc != 0:
        raise OracleError(lib().orc_last_error().decode())


NUCLEOTIDE = 0
AMINO_ACID = 1
NUC_SYMBOLS = "-ACGTRYSWKMBDHVN"
AA_SYMBOLS = "-ACDEFGHIKLMNOPQRSTUVWYBJZ*X"


class Export:
    """A column in the S1 upload format; `.desc` is a POINTER(ColumnDesc) valid until close()."""

    def __init__(self, handle):
        self._h = handle
        self.desc = lib().orc_export_desc(handle)

    def close(self):
        if self._h:
            lib().orc_export_free(self._h)
            self._h = None

    def __del__(self):
        self.close()


class Filter:
    def __init__(self, handle):
        self._h = handle

    @property
    def cardinality(self) -> int:
        return int(lib().orc_filter_cardinality(self._h))

    def ids(self) -> np.ndarray:
        n = self.cardinality
        if n == 0:
            return np.zeros(0, dtype=np.uint32)
        return np.ctypeslib.as_array(lib().orc_filter_ids(self._h), shape=(n,)).copy()

    def words(self, first_chunk: int, n_chunks: int) -> np.ndarray:
        out = np.zeros(n_chunks * 1024, dtype=np.uint64)
        _check(lib().orc_filter_words(self._h, first_chunk, n_chunks, out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return out

    def close(self):
        if self._h:
            lib().orc_filter_free(self._h)
            self._h = None

    def __del__(self):
        self.close()


class Table:
    def __init__(self):
        self._h = lib().orc_table_new()
        self.columns: list[tuple[str, int, str]] = []

    def close(self):
        if self._h:
            lib().orc_table_free(self._h)
            self._h = None

    def __del__(self):
        self.close()

    # ---- building ----
    def add_column(self, name: str, alphabet: int, reference: str) -> None:
        _check(lib().orc_table_add_column(self._h, name.encode(), alphabet, reference.encode()))
        self.columns.append((name, alphabet, reference))

    def set_layout(self, *chunk_sizes: int) -> None:
        arr = (C.c_uint32 * len(chunk_sizes))(*chunk_sizes)
        _check(lib().orc_table_set_layout(self._h, arr, len(chunk_sizes)))

    def append_row(self, values: Sequence[Optional[str | tuple[str, int]]]) -> None:
        """One value per column: None (null), "SEQ" or ("SEQ", offset)."""
        seqs = (C.c_char_p * len(values))()
        offs = (C.c_uint32 * len(values))()
        for i, value in enumerate(values):
            if value is None:
                seqs[i] = None
            elif isinstance(value, tuple):
                seqs[i] = value[0].encode()
                offs[i] = value[1]
            else:
                seqs[i] = value.encode()
        _check(lib().orc_table_append_row(self._h, seqs, offs))

    def append_cycled(self, sequences: Sequence[str], n_rows: int, offsets: Optional[Sequence[int]] = None) -> None:
        seqs = (C.c_char_p * len(sequences))(*[s.encode() for s in sequences])
        offs = (C.c_uint32 * len(sequences))(*offsets) if offsets is not None else None
        _check(lib().orc_table_append_cycled(self._h, seqs, offs, len(sequences), n_rows))

    def flush_chunk(self) -> None:
        _check(lib().orc_table_flush_chunk(self._h))

    def finalize(self) -> None:
        _check(lib().orc_table_finalize(self._h))

    def add_string_column(self, name: str, values: Sequence[Optional[str]]) -> None:
        """An unindexed string column: one value per row in layout order (None = null); after the layout is known."""
        arr = (C.c_char_p * max(len(values), 1))(*[v.encode() if v is not None else None for v in values])
        _check(lib().orc_table_add_string_column(self._h, name.encode(), arr, len(values)))

    def add_string_column_ids(self, name: str, dictionary: Sequence[str], ids) -> None:
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        names = (C.c_char_p * max(len(dictionary), 1))(*[v.encode() for v in dictionary])
        _check(lib().orc_table_add_string_column_ids(self._h, name.encode(), names, len(dictionary), ids.ctypes.data, ids.size))

    def add_date_column(self, name: str, days) -> None:
        """A Date32 column: days since the epoch per row in layout order (None = null)."""
        if isinstance(days, np.ndarray):
            values, nulls = np.ascontiguousarray(days, dtype=np.int32), None
        else:
            values = np.array([d if d is not None else 0 for d in days], dtype=np.int32)
            nulls = np.array([d is None for d in days], dtype=np.uint8)
        _check(lib().orc_table_add_date_column(self._h, name.encode(), values.ctypes.data, nulls.ctypes.data if nulls is not None else None, values.size))

    def register_bitmap(self, name: str, ids: Iterable[int]) -> None:
        arr = np.ascontiguousarray(np.fromiter(ids, dtype=np.uint32))
        _check(lib().orc_table_register_bitmap(
            self._h, name.encode(), arr.ctypes.data_as(C.POINTER(C.c_uint32)), arr.size))

    def bitmap_bytes(self, name: str) -> bytes:
        n = lib().orc_table_bitmap_bytes(self._h, name.encode(), None, 0)
        if n < 0:
            raise OracleError(lib().orc_last_error().decode())
        buf = (C.c_uint8 * n)()
        lib().orc_table_bitmap_bytes(self._h, name.encode(), buf, n)
        return bytes(buf)

    # ---- inspection ----
    @property
    def num_rows(self) -> int:
        return int(lib().orc_table_num_rows(self._h))

    @property
    def chunk_sizes(self) -> list[int]:
        n = int(lib().orc_table_num_chunks(self._h))
        arr = (C.c_uint32 * n)()
        lib().orc_table_chunk_sizes(self._h, arr)
        return list(arr)

    def local_reference(self, column: str) -> str:
        length = len(next(ref for name, _, ref in self.columns if name == column))
        buf = C.create_string_buffer(length)
        _check(lib().orc_column_local_reference(self._h, column.encode(), buf))
        return buf.raw.decode()

    def num_containers(self, column: str) -> int:
        return int(lib().orc_column_num_containers(self._h, column.encode()))

    # ---- the path ----
    def filter(self, expression: str) -> Filter:
        handle = lib().orc_filter_eval(self._h, expression.encode())
        if not handle:
            raise OracleError(lib().orc_last_error().decode())
        return Filter(handle)

    def mutation_counts(self, column: str, flt: Optional[Filter] = None) -> np.ndarray:
        name, alphabet, reference = next(c for c in self.columns if c[0] == column)
        n_symbols = 16 if alphabet == NUCLEOTIDE else 28
        out = np.zeros(n_symbols * len(reference), dtype=np.uint32)
        _check(lib().orc_mutation_counts(
            self._h, column.encode(), flt._h if flt is not None else None,
            out.ctypes.data_as(C.POINTER(C.c_uint32))))
        return out.reshape(n_symbols, len(reference))

    def mutation_rows(self, column: str, counts: np.ndarray, min_proportion: float) -> list[dict]:
        counts = np.ascontiguousarray(counts, dtype=np.uint32)
        handle = lib().orc_mutation_rows(
            self._h, column.encode(), counts.ctypes.data_as(C.POINTER(C.c_uint32)), min_proportion)
        if not handle:
            raise OracleError(lib().orc_last_error().decode())
        rows = []
        try:
            for i in range(lib().orc_mutation_rows_size(handle)):
                frm, to = C.c_char(), C.c_char()
                pos, cnt, cov = C.c_int32(), C.c_int32(), C.c_int32()
                prop = C.c_double()
                lib().orc_mutation_rows_get(handle, i, C.byref(frm), C.byref(to), C.byref(pos),
                                            C.byref(prop), C.byref(cnt), C.byref(cov))
                rows.append({
                    "mutationFrom": frm.value.decode(), "mutationTo": to.value.decode(),
                    "position": pos.value, "sequenceName": column, "proportion": prop.value,
                    "count": cnt.value, "coverage": cov.value,
                })
        finally:
            lib().orc_mutation_rows_free(handle)
        return rows

    def mutations_query(self, column: str, expression: str, min_proportion: float) -> tuple[int, int]:
        """The whole query in one native call; returns (output rows, filter cardinality)."""
        cardinality = C.c_uint64()
        n_rows = lib().orc_mutations_query(self._h, expression.encode(), column.encode(), min_proportion, C.byref(cardinality))
        if n_rows < 0:
            raise OracleError(lib().orc_last_error().decode())
        return int(n_rows), int(cardinality.value)

    def mutations_query_bench(self, column: str, expression: str, min_proportion: float, threads: int,
                              seconds: float, max_queries_per_thread: int = 1 << 62) -> tuple[int, float, int]:
        """`threads` native workers running independent queries back to back; returns
        (queries completed, elapsed seconds, filter cardinality)."""
        queries, elapsed, cardinality = C.c_uint64(), C.c_double(), C.c_uint64()
        _check(lib().orc_mutations_query_bench(
            self._h, expression.encode(), column.encode(), min_proportion, threads, seconds, max_queries_per_thread,
            C.byref(queries), C.byref(elapsed), C.byref(cardinality)))
        return int(queries.value), float(elapsed.value), int(cardinality.value)

    def mutations(self, column: str, expression: Optional[str], min_proportion: float) -> list[dict]:
        flt = self.filter(expression) if expression is not None else None
        return self.mutation_rows(column, self.mutation_counts(column, flt), min_proportion)

    # ---- S1 interchange ----
    def bitmap_aggregation(self, dimensions: Sequence, expression: Optional[str] = None) -> list[tuple]:
        """BitmapAggregationNode (bitmap_aggregation_node.cpp): dimensions are ("position", column, position0)
        or ("bitmaps", [(value, bitmap name), ...], null bitmap name or None). Returns the combinations in
        the reference's output order as (value-or-None per dimension ..., count) tuples."""
        parts = []
        for dim in dimensions:
            if dim[0] == "position":
                parts.append(f"p:{dim[1]}:{int(dim[2])}")
            else:
                parts.append("b:" + ",".join(f"{value}={name}" for value, name in dim[1]) + "|" + (dim[2] or ""))
        spec = ";".join(parts).encode()
        text_expression = expression.encode() if expression else None
        capacity = 1 << 16
        while True:
            buffer = C.create_string_buffer(capacity)
            needed = lib().orc_bitmap_aggregation(self._h, text_expression, spec, buffer, capacity)
            if needed < 0:
                _check(int(needed))
            if needed <= capacity:
                break
            capacity = int(needed)
        rows = []
        for line in buffer.raw[:needed].decode().splitlines():
            fields = line.split("\t")
            rows.append(tuple(None if f == "\\N" else f for f in fields[:-1]) + (int(fields[-1]),))
        return rows

    def export_column(self, column: str, first_chunk: int = 0, n_chunks: Optional[int] = None) -> Export:
        if n_chunks is None:
            n_chunks = len(self.chunk_sizes) - first_chunk
        handle = lib().orc_column_export(self._h, column.encode(), first_chunk, n_chunks)
        if not handle:
            raise OracleError(lib().orc_last_error().decode())
        return Export(handle)

    def import_column(self, name: str, alphabet: int, reference: str, desc_ptr) -> None:
        _check(lib().orc_table_import_column(
            self._h, name.encode(), alphabet, reference.encode(), C.cast(desc_ptr, C.c_void_p)))
        self.columns.append((name, alphabet, reference))


def gen_evolved(reference: str, seed: int = 42, mutation_rate: float = 0.001, death_rate: float = 0.1,
                generations: int = 5, children: int = 3) -> tuple[list[str], list[int]]:
    """SequenceTreeGenerator::generateEvolvedSequences (performance/sequence_generator.h:113-185)."""
    ref = reference.encode()
    n = lib().orc_gen_evolved(ref, seed, mutation_rate, death_rate, generations, children, None, 0, None)
    if n < 0:
        raise OracleError(lib().orc_last_error().decode())
    cap = n * (len(ref) + 1)
    buf = C.create_string_buffer(cap)
    parents = (C.c_uint64 * n)()
    lib().orc_gen_evolved(ref, seed, mutation_rate, death_rate, generations, children, buf, cap, parents)
    raw = buf.raw
    step = len(ref) + 1
    return [raw[i * step:i * step + len(ref)].decode() for i in range(n)], list(parents)


def full_sequence_table(reference: str, count: int, generations: int = 5) -> Table:
    table = Table()
    _check(lib().orc_gen_full_sequence_table(table._h, reference.encode(), count, generations))
    table.columns.append(("main", NUCLEOTIDE, reference))
    return table


def nrun_table(reference: str, count: int, generations: int = 12) -> Table:
    table = Table()
    _check(lib().orc_gen_nrun_table(table._h, reference.encode(), count, generations))
    table.columns.append(("main", NUCLEOTIDE, reference))
    return table


def mutation_benchmark_table(batch_rows: int = 1000) -> Table:
    table = Table()
    _check(lib().orc_gen_mutation_benchmark_table(table._h, batch_rows))
    table.columns.append(("main", NUCLEOTIDE, "ACGT" * 1000))
    return table


def now_seconds() -> float:
    return float(lib().orc_now_seconds())


CONTAINER_OPS = {"identity": 0, "and": 1, "andnot": 2, "or": 3, "and_cardinality": 4}


def container_op(a: Sequence[int], b: Sequence[int] = (), op: str = "identity", optimize_a: bool = False,
                 optimize_b: bool = False, roundtrip: bool = False):
    """Container-level hook. Returns (values | cardinality, typecode_of_a, typecode_of_result)."""
    arr_a = (C.c_uint16 * max(len(a), 1))(*a)
    arr_b = (C.c_uint16 * max(len(b), 1))(*b)
    out = (C.c_uint16 * 65536)()
    n_out = C.c_uint32()
    type_a, type_out = C.c_uint8(), C.c_uint8()
    _check(lib().orc_container_op(arr_a, len(a), int(optimize_a), arr_b, len(b), int(optimize_b),
                                  CONTAINER_OPS[op], int(roundtrip), out, C.byref(n_out),
                                  C.byref(type_a), C.byref(type_out)))
    if op == "and_cardinality":
        return n_out.value, type_a.value, 0
    return list(out[:n_out.value]), type_a.value, type_out.value


def roaring_roundtrip(ids: Sequence[int], optimize: bool = False) -> tuple[list[int], bytes]:
    """ids -> Roaring -> portable bytes -> Roaring -> ids."""
    arr = np.ascontiguousarray(np.asarray(list(ids), dtype=np.uint32))
    n_unique = len(set(int(i) for i in ids))
    out = np.zeros(max(n_unique, 1), dtype=np.uint32)
    cap = 16 + 16 * (len(arr) + 1) + 8192 * (len({int(i) >> 16 for i in ids}) + 1)
    buf = (C.c_uint8 * cap)()
    n = lib().orc_roaring_roundtrip(arr.ctypes.data_as(C.POINTER(C.c_uint32)), arr.size, int(optimize),
                                    out.ctypes.data_as(C.POINTER(C.c_uint32)), buf, cap)
    if n < 0:
        raise OracleError(lib().orc_last_error().decode())
    return [int(v) for v in out[:n_unique]], bytes(buf[:n])


def extract(sequence: str, offset: int, reference: str, alphabet: int = NUCLEOTIDE):
    """extractCoverageAndMutationsFromSequence -> (start, end, missing, [(pos, symbol_char)])."""
    n = max(len(sequence), 1)
    start, end, n_missing, n_mut = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    missing = (C.c_uint32 * n)()
    mut_pos = (C.c_uint32 * n)()
    mut_sym = (C.c_uint8 * n)()
    _check(lib().orc_extract(alphabet, sequence.encode(), offset, reference.encode(), C.byref(start),
                             C.byref(end), missing, C.byref(n_missing), mut_pos, mut_sym, C.byref(n_mut)))
    chars = NUC_SYMBOLS if alphabet == NUCLEOTIDE else AA_SYMBOLS
    return (start.value, end.value, list(missing[:n_missing.value]),
            [(mut_pos[i], chars[mut_sym[i]]) for i in range(n_mut.value)])


def containers_at(table: Table, column: str, position0: int) -> list[tuple[int, str, int, int]]:
    """Stored diff containers at a 0-based position: (v_index, symbol_char, typecode, cardinality)."""
    alphabet = next(a for name, a, _ in table.columns if name == column)
    chars = NUC_SYMBOLS if alphabet == NUCLEOTIDE else AA_SYMBOLS
    n = lib().orc_column_containers_at(table._h, column.encode(), position0, None, 0)
    if n < 0:
        raise OracleError(lib().orc_last_error().decode())
    out = (C.c_uint32 * (4 * max(n, 1)))()
    lib().orc_column_containers_at(table._h, column.encode(), position0, out, n)
    return [(out[4 * i], chars[out[4 * i + 1]], out[4 * i + 2], out[4 * i + 3]) for i in range(n)]
