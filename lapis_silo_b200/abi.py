"""ctypes view of include/silo_b200.h (libsilo_b200.so) — the drop-in C ABI itself.

This is plumbing for the Python harness (tests, bench.py); the product is the shared library.
There is no fallback: a missing library or a missing CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
import subprocess
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsilo_b200.so")

SILO_OK = 0
SILO_E_INVALID_ARGUMENT = -1
SILO_E_NO_DEVICE = -2
SILO_E_CUDA = -3
SILO_E_OUT_OF_MEMORY = -4
SILO_E_BAD_PROGRAM = -5
SILO_E_OUT_OF_LAYOUT = -6
SILO_E_UNSUPPORTED = -7

# silo_filter_opcode
OP_PUSH_EMPTY, OP_PUSH_FULL, OP_PUSH_SYMBOLS, OP_PUSH_COVERED = 1, 2, 3, 4
OP_PUSH_NULLS, OP_PUSH_BITMAP, OP_PUSH_RANGES = 5, 6, 7
OP_AND, OP_ANDNOT, OP_OR, OP_NOT = 16, 17, 18, 19
OP_PUSH_INDEX_BITMAP = 8
OP_THR_BEGIN, OP_THR_ADD, OP_THR_ADD_SYMBOLS, OP_THR_ADD_COVERED, OP_THR_PROFILE, OP_THR_END = 32, 33, 34, 35, 36, 37


class ContainerDesc(C.Structure):
    _fields_ = [
        ("position", C.c_uint32),
        ("v_index", C.c_uint16),
        ("symbol", C.c_uint8),
        ("typecode", C.c_uint8),
        ("cardinality", C.c_uint32),
        ("payload_bytes", C.c_uint32),
        ("payload_offset", C.c_uint64),
    ]


class ColumnDesc(C.Structure):
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


class FilterInstr(C.Structure):
    _fields_ = [
        ("opcode", C.c_uint8),
        ("flags", C.c_uint8),
        ("column", C.c_uint16),
        ("a", C.c_uint32),
        ("b", C.c_uint64),
    ]


class RoaringBytes(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_uint8)), ("size", C.c_uint64)]


class FilterProgram(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("n_instrs", C.c_uint32),
        ("instrs", C.POINTER(FilterInstr)),
        ("blob", C.POINTER(C.c_uint8)),
        ("blob_bytes", C.c_uint64),
        ("n_bitmaps", C.c_uint32),
        ("bitmaps", C.POINTER(RoaringBytes)),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("containers", C.c_uint64),
        ("algorithmic_bytes", C.c_uint64),
        ("counts_kernel_bytes", C.c_uint64),
        ("kernel_launches", C.c_uint64),
        ("last_counts_kernel_ms", C.c_float),
        ("last_total_ms", C.c_float),
        ("timed_calls", C.c_uint64),
    ]


EXPORTED_SYMBOLS = [
    "silo_gpu_last_error", "silo_gpu_version", "silo_gpu_init", "silo_gpu_shutdown",
    "silo_gpu_table_create", "silo_gpu_table_free", "silo_gpu_column_upload",
    "silo_gpu_table_device_bytes", "silo_gpu_table_set_option", "silo_gpu_filter_eval", "silo_gpu_program_prepare",
    "silo_gpu_program_run_async", "silo_gpu_program_run_counts_async", "silo_gpu_program_device_bytes", "silo_gpu_program_free",
    "silo_gpu_host_alloc", "silo_gpu_host_free",
    "silo_gpu_filter_from_words", "silo_gpu_bitmap_register", "silo_gpu_bitmap_unregister",
    "silo_gpu_filter_cardinality", "silo_gpu_filter_download", "silo_gpu_filter_free",
    "silo_gpu_mutation_counts", "silo_gpu_mutation_counts_symbols", "silo_gpu_query_mutation_counts", "silo_gpu_mutation_counts_async", "silo_gpu_get_stats",
    "silo_gpu_column_set_reference", "silo_gpu_query_mutation_hits", "silo_gpu_query_combinations",
    "silo_gpu_query_mutation_counts_async", "silo_gpu_mutation_hits_from_counts",
    "silo_gpu_shard_group_init", "silo_gpu_shard_group_connect", "silo_gpu_sharded_query_enqueue", "silo_gpu_sharded_collect",
    "silo_gpu_sharded_collect_async", "silo_gpu_shard_group_free", "silo_gpu_program_run_sharded_async", "silo_gpu_program_run_sharded_collect_async", "silo_gpu_sharded_query_hits", "silo_gpu_get_sweep_stats", "silo_gpu_query_mutation_hits_columns", "silo_gpu_value_column_upload", "silo_gpu_query_count",
]


def build(force: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a (see Makefile); cross-compiles without a GPU."""
    if force:
        subprocess.run(["make", "-s", "-C", _HERE, "clean"], check=True)
    subprocess.run(["make", "-s", "-C", _HERE, "-j4"], check=True)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.silo_gpu_last_error.restype = C.c_char_p
        L.silo_gpu_version.restype = C.c_char_p
        L.silo_gpu_init.argtypes = [C.c_int, C.POINTER(vp)]
        L.silo_gpu_shutdown.argtypes = [vp]
        L.silo_gpu_shutdown.restype = None
        L.silo_gpu_table_create.argtypes = [vp, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(vp)]
        L.silo_gpu_table_free.argtypes = [vp]
        L.silo_gpu_table_free.restype = None
        L.silo_gpu_column_upload.argtypes = [vp, vp]
        L.silo_gpu_table_set_option.argtypes = [vp, C.c_char_p, C.c_uint64]
        L.silo_gpu_get_sweep_stats.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.silo_gpu_table_device_bytes.argtypes = [vp]
        L.silo_gpu_table_device_bytes.restype = C.c_uint64
        L.silo_gpu_filter_eval.argtypes = [vp, C.POINTER(FilterProgram), C.POINTER(vp), C.POINTER(C.c_uint64)]
        L.silo_gpu_program_prepare.argtypes = [vp, C.POINTER(FilterProgram), C.POINTER(vp), C.POINTER(vp)]
        L.silo_gpu_program_run_async.argtypes = [vp, vp]
        L.silo_gpu_program_run_counts_async.argtypes = [vp, C.c_int, vp, vp]
        L.silo_gpu_program_device_bytes.argtypes = [vp]
        L.silo_gpu_program_device_bytes.restype = C.c_uint64
        L.silo_gpu_program_free.argtypes = [vp]
        L.silo_gpu_program_free.restype = None
        L.silo_gpu_filter_from_words.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(vp)]
        L.silo_gpu_host_alloc.argtypes = [vp, C.c_uint64]
        L.silo_gpu_host_alloc.restype = vp
        L.silo_gpu_host_free.argtypes = [vp, vp]
        L.silo_gpu_host_free.restype = None
        L.silo_gpu_bitmap_register.argtypes = [vp, C.c_char_p, C.c_uint64, C.POINTER(C.c_uint32)]
        L.silo_gpu_bitmap_unregister.argtypes = [vp, C.c_uint32]
        L.silo_gpu_filter_cardinality.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.silo_gpu_filter_download.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.silo_gpu_filter_free.argtypes = [vp]
        L.silo_gpu_filter_free.restype = None
        L.silo_gpu_mutation_counts.argtypes = [vp, C.c_int, vp, C.POINTER(C.c_uint32)]
        L.silo_gpu_mutation_counts_symbols.argtypes = [vp, C.c_int, vp, C.c_uint64, C.POINTER(C.c_uint32)]
        L.silo_gpu_mutation_counts_async.argtypes = [vp, C.c_int, vp, vp, vp]
        L.silo_gpu_get_stats.argtypes = [vp, C.POINTER(Stats)]
        _lib = L
    return _lib


class SiloGpuError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"[{status}] {message}")
        self.status = status


def check(status: int) -> int:
    if status < 0:
        raise SiloGpuError(status, lib().silo_gpu_last_error().decode())
    return status


class Context:
    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        check(lib().silo_gpu_init(device, C.byref(self._h)))
        self.device = device

    def close(self):
        if self._h:
            lib().silo_gpu_shutdown(self._h)
            self._h = C.c_void_p()


class Filter:
    def __init__(self, table: "Table", handle: C.c_void_p, cardinality: int | None = None):
        self.table = table
        self._h = handle
        self._cardinality = cardinality
        table._children.add(self)  # a table frees its filters before its pools

    @property
    def cardinality(self) -> int:
        if self._cardinality is None:
            value = C.c_uint64()
            check(lib().silo_gpu_filter_cardinality(self._h, C.byref(value)))
            self._cardinality = int(value.value)
        return self._cardinality

    def words(self) -> np.ndarray:
        out = np.zeros(self.table.n_chunks * 1024, dtype=np.uint64)
        check(lib().silo_gpu_filter_download(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return out

    def ids(self) -> np.ndarray:
        """Global sparse row ids, ascending."""
        bits = np.unpackbits(self.words().view(np.uint8), bitorder="little")
        local = np.flatnonzero(bits).astype(np.uint64)
        return (local + (np.uint64(self.table.first_chunk) << np.uint64(16))).astype(np.uint32)

    def close(self):
        if self._h:
            lib().silo_gpu_filter_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        if sys is None or sys.is_finalizing():  # (the CUDA runtime's own exit handlers may have run: the process frees everything anyway)
            return
        try:
            self.close()
        except Exception:
            pass


class Table:
    def __init__(self, ctx: Context, chunk_sizes, first_chunk: int = 0):
        self.ctx = ctx
        self.chunk_sizes = [int(s) for s in chunk_sizes]
        self.n_chunks = len(self.chunk_sizes)
        self.first_chunk = first_chunk
        self.columns: list[tuple[int, int]] = []  # (n_symbols, genome_length)
        self._children = weakref.WeakSet()
        self._h = C.c_void_p()
        arr = (C.c_uint32 * max(self.n_chunks, 1))(*self.chunk_sizes)
        check(lib().silo_gpu_table_create(ctx._h, first_chunk, arr, self.n_chunks, C.byref(self._h)))

    def upload_column(self, desc_ptr) -> int:
        """desc_ptr: POINTER(ColumnDesc) or address of a silo_column_desc."""
        desc = C.cast(desc_ptr, C.POINTER(ColumnDesc)).contents
        index = check(lib().silo_gpu_column_upload(self._h, C.cast(desc_ptr, C.c_void_p)))
        self.columns.append((int(desc.n_symbols), int(desc.genome_length)))
        return index

    @property
    def device_bytes(self) -> int:
        return int(lib().silo_gpu_table_device_bytes(self._h))

    def filter_eval(self, instrs: list[tuple], blob: bytes = b"", bitmaps: list[bytes] = ()) -> Filter:
        """instrs: (opcode, flags, column, a, b) tuples."""
        arr = (FilterInstr * max(len(instrs), 1))()
        for i, (opcode, flags, column, a, b) in enumerate(instrs):
            arr[i] = FilterInstr(opcode, flags, column, a, b)
        blob_buf = (C.c_uint8 * max(len(blob), 1)).from_buffer_copy(blob.ljust(1, b"\0"))
        keep = [(C.c_uint8 * max(len(raw), 1)).from_buffer_copy(raw.ljust(1, b"\0")) for raw in bitmaps]
        bm = (RoaringBytes * max(len(bitmaps), 1))()
        for i, raw in enumerate(bitmaps):
            bm[i] = RoaringBytes(C.cast(keep[i], C.POINTER(C.c_uint8)), len(raw))
        program = FilterProgram(C.sizeof(FilterProgram), len(instrs), arr, C.cast(blob_buf, C.POINTER(C.c_uint8)),
                                len(blob), len(bitmaps), bm)
        handle = C.c_void_p()
        cardinality = C.c_uint64()
        check(lib().silo_gpu_filter_eval(self._h, C.byref(program), C.byref(handle), C.byref(cardinality)))
        return Filter(self, handle, int(cardinality.value))

    def register_bitmap(self, portable_roaring_bytes: bytes) -> int:
        out = C.c_uint32()
        check(lib().silo_gpu_bitmap_register(self._h, portable_roaring_bytes, len(portable_roaring_bytes), C.byref(out)))
        return out.value

    def unregister_bitmap(self, bitmap_id: int) -> None:
        check(lib().silo_gpu_bitmap_unregister(self._h, bitmap_id))

    def filter_from_words(self, words: np.ndarray) -> Filter:
        words = np.ascontiguousarray(words, dtype=np.uint64)
        assert words.size == self.n_chunks * 1024
        handle = C.c_void_p()
        check(lib().silo_gpu_filter_from_words(self._h, words.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(handle)))
        return Filter(self, handle)

    def mutation_counts(self, column: int, flt: Filter | None = None) -> np.ndarray:
        n_symbols, genome_length = self.columns[column]
        out = np.zeros(n_symbols * genome_length, dtype=np.uint32)
        check(lib().silo_gpu_mutation_counts(
            self._h, column, flt._h if flt is not None else None, out.ctypes.data_as(C.POINTER(C.c_uint32))))
        return out.reshape(n_symbols, genome_length)

    def mutation_counts_async(self, column: int, flt: Filter | None, d_counts_ptr: int, stream_ptr: int) -> None:
        check(lib().silo_gpu_mutation_counts_async(
            self._h, column, flt._h if flt is not None else None, C.c_void_p(d_counts_ptr), C.c_void_p(stream_ptr)))

    def stats(self) -> Stats:
        out = Stats()
        check(lib().silo_gpu_get_stats(self._h, C.byref(out)))
        return out

    def close(self):
        if self._h:
            for child in list(self._children):
                child.close()
            lib().silo_gpu_table_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        if sys is None or sys.is_finalizing():  # (the CUDA runtime's own exit handlers may have run: the process frees everything anyway)
            return
        try:
            self.close()
        except Exception:
            pass
