"""ctypes view of include/silo_b200_host.h (libsilo_b200_host.so): the C++ host layer that mirrors
the reference's expression -> operator compilation and the Mutations / count sinks.

Harness only (tests, bench.py). No compute happens in Python and nothing here falls back to a CPU
path: every filter and every count goes host layer -> C ABI -> sm_100a kernels.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
import weakref
from typing import Optional, Sequence

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "libsilo_b200_host.so")

HOST_EXPORTED_SYMBOLS = [
    "silo_host_last_error", "silo_host_table_create", "silo_host_table_free", "silo_host_table_add_column",
    "silo_host_last_query_profile", "silo_host_table_register_bitmap", "silo_host_table_device", "silo_host_table_num_rows",
    "silo_host_filter_eval", "silo_host_filter_free", "silo_host_filter_cardinality",
    "silo_host_filter_device", "silo_host_filter_words", "silo_host_filter_explain", "silo_host_bitmap_aggregation", "silo_host_bitmap_aggregation_shard", "silo_host_bitmap_aggregation_merge", "silo_host_bitmap_aggregation_packed", "silo_host_bitmap_aggregation_merge_packed",
    "silo_host_filter_prepare", "silo_host_prepared_run_async", "silo_host_prepared_run_counts_async", "silo_host_prepared_filter",
    "silo_host_prepared_staged_bytes", "silo_host_prepared_free",
    "silo_host_mutation_counts", "silo_host_mutations", "silo_host_mutation_rows_from_counts",
    "silo_host_mutations_packed", "silo_host_packed_fetch", "silo_host_mutations_enqueue", "silo_host_mutations_collect_packed",
    "silo_host_rows_free", "silo_host_rows_size", "silo_host_rows_get", "silo_host_rows_export",
    "silo_host_rows_num_names", "silo_host_rows_name",
    "silo_host_synthetic_create", "silo_host_synthetic_free", "silo_host_synthetic_num_sequences",
    "silo_host_synthetic_reference", "silo_host_synthetic_sequence", "silo_host_synthetic_parent",
    "silo_host_synthetic_generation", "silo_host_synthetic_build_column",
    "silo_host_synthetic_release_column", "silo_host_synthetic_lineage_bitmap",
    "silo_host_synthetic_date_ranges", "silo_host_partition_chunks",
    "silo_host_filter_lower_timed", "silo_host_filter_to_string", "silo_host_filter_program_bitmap",
    "silo_host_archive_read", "silo_host_archive_free", "silo_host_archive_column", "silo_host_archive_column_info",
    "silo_host_archive_chunk_sizes", "silo_host_archive_column_shard", "silo_host_table_load_archive", "silo_host_roaring_runs",
    "silo_host_count", "silo_host_synthetic_draw_short_reads", "silo_host_synthetic_build_short_read_column", "silo_host_table_add_string_column", "silo_host_table_add_date_column", "silo_host_synthetic_create_gene", "silo_host_synthetic_create_co_occurrence", "silo_host_shard_group_create", "silo_host_shard_group_connect", "silo_host_sharded_enqueue", "silo_host_sharded_collect_packed",
    "silo_host_prepared_run_sharded_async", "silo_host_prepared_run_sharded_collect_async", "silo_host_sharded_collect_async", "silo_host_sharded_query_packed",
]
SHARD_HANDLE_BYTES = 128  # SILO_SHARD_HANDLE_BYTES

NUCLEOTIDE = 0
AMINO_ACID = 1

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        abi.lib()  # libsilo_b200.so first (the host library links against it)
        if not os.path.exists(HOST_LIB_PATH):
            raise RuntimeError(f"{HOST_LIB_PATH} is missing: run __graft_entry__.build()")
        L = C.CDLL(HOST_LIB_PATH)
        vp = C.c_void_p
        L.silo_host_last_error.restype = C.c_char_p
        L.silo_host_table_create.argtypes = [vp, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32]
        L.silo_host_table_create.restype = vp
        L.silo_host_table_free.argtypes = [vp]
        L.silo_host_table_free.restype = None
        L.silo_host_table_add_column.argtypes = [vp, C.c_char_p, C.c_int, C.c_char_p, vp]
        L.silo_host_table_register_bitmap.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_uint64, C.c_int]
        L.silo_host_table_add_string_column.argtypes = [vp, C.c_char_p, C.POINTER(C.c_char_p), C.c_uint32, vp, vp, C.c_uint64]
        L.silo_host_table_add_date_column.argtypes = [vp, C.c_char_p, vp, vp, C.c_uint64]
        L.silo_host_table_device.argtypes = [vp]
        L.silo_host_table_device.restype = vp
        L.silo_host_table_num_rows.argtypes = [vp]
        L.silo_host_table_num_rows.restype = C.c_uint64
        L.silo_host_count.argtypes = [vp, C.c_char_p, C.POINTER(C.c_uint64)]
        L.silo_host_filter_eval.argtypes = [vp, C.c_char_p]
        L.silo_host_filter_eval.restype = vp
        L.silo_host_filter_free.argtypes = [vp]
        L.silo_host_filter_free.restype = None
        L.silo_host_filter_cardinality.argtypes = [vp]
        L.silo_host_filter_cardinality.restype = C.c_uint64
        L.silo_host_filter_device.argtypes = [vp]
        L.silo_host_filter_device.restype = vp
        L.silo_host_filter_words.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.silo_host_filter_explain.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_uint64]
        L.silo_host_bitmap_aggregation.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint64]
        L.silo_host_bitmap_aggregation_shard.argtypes = [vp, C.c_char_p, C.c_char_p, vp, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.silo_host_bitmap_aggregation_merge.argtypes = [vp, C.c_char_p, vp, vp, vp, C.c_uint32, C.c_char_p, C.c_uint64]
        L.silo_host_bitmap_aggregation_packed.argtypes = [vp, C.c_char_p, C.c_char_p, vp, vp, C.c_uint64, C.POINTER(C.c_uint64)]
        L.silo_host_bitmap_aggregation_merge_packed.argtypes = [vp, C.c_char_p, vp, vp, vp, C.c_uint32, vp, vp, C.c_uint64, C.POINTER(C.c_uint64)]
        L.silo_host_filter_prepare.argtypes = [vp, C.c_char_p]
        L.silo_host_filter_prepare.restype = vp
        L.silo_host_prepared_run_async.argtypes = [vp, vp]
        L.silo_host_prepared_run_counts_async.argtypes = [vp, C.c_int, vp, vp]
        L.silo_host_prepared_filter.argtypes = [vp]
        L.silo_host_prepared_filter.restype = vp
        L.silo_host_prepared_staged_bytes.argtypes = [vp]
        L.silo_host_prepared_staged_bytes.restype = C.c_uint64
        L.silo_host_prepared_free.argtypes = [vp]
        L.silo_host_prepared_free.restype = None
        L.silo_host_mutation_counts.argtypes = [vp, C.c_char_p, vp, C.POINTER(C.c_uint32)]
        L.silo_host_mutations.argtypes = [vp, C.c_char_p, C.POINTER(C.c_char_p), C.c_uint32, C.c_double]
        L.silo_host_mutations.restype = vp
        L.silo_host_mutations_packed.argtypes = [
            vp, C.c_char_p, C.POINTER(C.c_char_p), C.c_uint32, C.c_double, vp, C.c_uint64,
            C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
        L.silo_host_packed_fetch.argtypes = [vp, C.c_uint64]
        L.silo_host_mutations_enqueue.argtypes = [vp, C.c_char_p, C.c_char_p, vp, vp]
        L.silo_host_mutations_collect_packed.argtypes = [
            vp, C.c_char_p, C.c_double, vp, vp, vp, C.c_uint64,
            C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.silo_host_shard_group_create.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, C.c_char_p]
        L.silo_host_shard_group_connect.argtypes = [vp, C.c_char_p, C.c_uint64]
        L.silo_host_sharded_enqueue.argtypes = [vp, C.c_char_p, C.c_char_p, vp]
        L.silo_host_sharded_collect_packed.argtypes = [
            vp, C.c_char_p, C.c_double, vp, vp, vp, C.c_uint64,
            C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.silo_host_sharded_query_packed.argtypes = [
            vp, C.c_char_p, C.c_char_p, C.c_double, vp, vp, C.c_uint64,
            C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.silo_host_prepared_run_sharded_async.argtypes = [vp, vp]
        L.silo_host_prepared_run_sharded_collect_async.argtypes = [vp, vp, vp]
        L.silo_host_sharded_collect_async.argtypes = [vp, vp, vp]
        L.silo_host_mutation_rows_from_counts.argtypes = [vp, C.c_char_p, C.POINTER(C.c_uint32), C.c_double]
        L.silo_host_mutation_rows_from_counts.restype = vp
        L.silo_host_rows_free.argtypes = [vp]
        L.silo_host_rows_free.restype = None
        L.silo_host_rows_size.argtypes = [vp]
        L.silo_host_rows_size.restype = C.c_uint64
        L.silo_host_rows_get.argtypes = [
            vp, C.c_uint64, C.POINTER(C.c_char), C.POINTER(C.c_char), C.POINTER(C.c_int32),
            C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.silo_host_rows_export.argtypes = [
            vp, C.c_char_p, C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_double),
            C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.silo_host_rows_num_names.argtypes = [vp]
        L.silo_host_rows_num_names.restype = C.c_uint32
        L.silo_host_rows_name.argtypes = [vp, C.c_uint32]
        L.silo_host_rows_name.restype = C.c_char_p
        L.silo_host_synthetic_create.argtypes = [C.c_uint32, C.c_uint64, C.c_uint32]
        L.silo_host_synthetic_create.restype = vp
        L.silo_host_synthetic_create_gene.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, C.c_double, C.c_uint32]
        L.silo_host_synthetic_create_gene.restype = vp
        L.silo_host_synthetic_create_co_occurrence.argtypes = [C.c_uint64]
        L.silo_host_synthetic_create_co_occurrence.restype = vp
        L.silo_host_synthetic_free.argtypes = [vp]
        L.silo_host_synthetic_free.restype = None
        L.silo_host_synthetic_num_sequences.argtypes = [vp]
        L.silo_host_synthetic_num_sequences.restype = C.c_uint32
        L.silo_host_synthetic_reference.argtypes = [vp]
        L.silo_host_synthetic_reference.restype = C.c_char_p
        L.silo_host_synthetic_sequence.argtypes = [vp, C.c_uint32]
        L.silo_host_synthetic_sequence.restype = C.c_char_p
        L.silo_host_synthetic_parent.argtypes = [vp, C.c_uint32]
        L.silo_host_synthetic_parent.restype = C.c_uint32
        L.silo_host_synthetic_generation.argtypes = [vp, C.c_uint32]
        L.silo_host_synthetic_generation.restype = C.c_uint32
        L.silo_host_synthetic_build_column.argtypes = [vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(vp), C.c_uint32]
        L.silo_host_synthetic_draw_short_reads.argtypes = [vp, C.c_uint64, C.c_uint32, vp]
        L.silo_host_synthetic_build_short_read_column.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(vp)]
        L.silo_host_synthetic_release_column.argtypes = [vp]
        L.silo_host_synthetic_release_column.restype = None
        L.silo_host_synthetic_lineage_bitmap.argtypes = [vp, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_char_p, C.c_uint64, C.c_uint32]
        L.silo_host_synthetic_lineage_bitmap.restype = C.c_int64
        L.silo_host_synthetic_date_ranges.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_char_p, C.c_uint64, C.c_uint32]
        L.silo_host_partition_chunks.argtypes = [C.POINTER(C.c_uint64), C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
        L.silo_host_filter_program_bitmap.argtypes = [vp, C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint64]
        L.silo_host_filter_program_bitmap.restype = C.c_int64
        L.silo_host_filter_to_string.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_uint64]
        L.silo_host_filter_lower_timed.argtypes = [vp, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        strings, ints = C.POINTER(C.c_char_p), C.POINTER(C.c_int)
        L.silo_host_archive_read.argtypes = [C.c_char_p, C.c_uint64, strings, ints, strings, C.c_uint32]
        L.silo_host_archive_read.restype = vp
        L.silo_host_archive_free.argtypes = [vp]
        L.silo_host_archive_free.restype = None
        L.silo_host_archive_column.argtypes = [vp, C.c_uint32]
        L.silo_host_archive_column.restype = C.POINTER(abi.ColumnDesc)
        L.silo_host_archive_column_info.argtypes = [vp, C.c_uint32, C.POINTER(C.c_uint64)]
        L.silo_host_archive_chunk_sizes.argtypes = [vp, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32]
        L.silo_host_archive_column_shard.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32]
        L.silo_host_archive_column_shard.restype = C.POINTER(abi.ColumnDesc)
        L.silo_host_table_load_archive.argtypes = [vp, C.c_char_p, C.c_uint64, strings, ints, strings, C.c_uint32, C.c_uint32, C.c_uint32]
        L.silo_host_table_load_archive.restype = vp
        L.silo_host_roaring_runs.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(C.c_uint32), C.c_uint64]
        L.silo_host_roaring_runs.restype = C.c_int64
        _lib = L
    return _lib


class HostError(RuntimeError):
    pass


def _check(status: int) -> None:
    if status != 0:
        raise HostError(lib().silo_host_last_error().decode())


def _columns(handle) -> dict:
    """The result set as columns (one bulk export call): what the reference hands to its Arrow sink as
    a record batch (mutations_node.cpp:404-428). Keys: mutationFrom / mutationTo (bytes, one char per
    row), position, sequenceNameId (+ sequenceNames), proportion, count, coverage (numpy arrays)."""
    if not handle:
        raise HostError(lib().silo_host_last_error().decode())
    try:
        n = int(lib().silo_host_rows_size(handle))
        frm, to = C.create_string_buffer(max(n, 1)), C.create_string_buffer(max(n, 1))
        position = np.empty(n, dtype=np.int32)
        name_ids = np.empty(n, dtype=np.uint32)
        proportion = np.empty(n, dtype=np.float64)
        count = np.empty(n, dtype=np.int32)
        coverage = np.empty(n, dtype=np.int32)
        if n > 0:
            _check(lib().silo_host_rows_export(
                handle, frm, to, position.ctypes.data_as(C.POINTER(C.c_int32)),
                name_ids.ctypes.data_as(C.POINTER(C.c_uint32)), proportion.ctypes.data_as(C.POINTER(C.c_double)),
                count.ctypes.data_as(C.POINTER(C.c_int32)), coverage.ctypes.data_as(C.POINTER(C.c_int32))))
        names = [lib().silo_host_rows_name(handle, i).decode() for i in range(lib().silo_host_rows_num_names(handle))]
        return {"mutationFrom": frm.raw[:n], "mutationTo": to.raw[:n], "position": position, "sequenceNameId": name_ids,
                "sequenceNames": names, "proportion": proportion, "count": count, "coverage": coverage}
    finally:
        lib().silo_host_rows_free(handle)


def _unpack_record_batch(batch: np.ndarray, n: int, n_names: int) -> dict:
    """Views over the record batch of silo_host_mutations_packed (include/silo_b200_host.h)."""
    def pad(size):
        return (size + 7) // 8 * 8
    cursor = 0
    out = {}
    for key, dtype, width in (("proportion", np.float64, 8), ("position", np.int32, 4), ("sequenceNameId", np.uint32, 4),
                              ("count", np.int32, 4), ("coverage", np.int32, 4)):
        out[key] = batch[cursor:cursor + width * n].view(dtype)
        cursor += pad(width * n)
    out["mutationFrom"] = batch[cursor:cursor + n].tobytes()
    cursor += pad(n)
    out["mutationTo"] = batch[cursor:cursor + n].tobytes()
    cursor += pad(n)
    out["sequenceNames"] = batch[cursor:].tobytes().decode().split("\0")[:n_names]
    return out


def rows_from_columns(columns: dict) -> list[dict]:
    """The same result set as the list of row dicts the oracle binding returns."""
    frm_text, to_text = columns["mutationFrom"].decode("latin-1"), columns["mutationTo"].decode("latin-1")
    names = columns["sequenceNames"]
    return [{
        "mutationFrom": f, "mutationTo": t, "position": p, "sequenceName": names[s], "proportion": q,
        "count": c, "coverage": v,
    } for f, t, p, s, q, c, v in zip(frm_text, to_text, columns["position"].tolist(), columns["sequenceNameId"].tolist(),
                                     columns["proportion"].tolist(), columns["count"].tolist(), columns["coverage"].tolist())]


def _rows(handle) -> list[dict]:
    return rows_from_columns(_columns(handle))


class HostFilter:
    def __init__(self, table: "HostTable", handle):
        self.table = table
        self._h = handle
        table._children.add(self)  # a table closes its filters before it frees the device pools

    @property
    def cardinality(self) -> int:
        return int(lib().silo_host_filter_cardinality(self._h))

    @property
    def device_handle(self) -> int:
        return lib().silo_host_filter_device(self._h)

    def words(self) -> np.ndarray:
        out = np.zeros(self.table.n_chunks * 1024, dtype=np.uint64)
        _check(lib().silo_host_filter_words(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return out

    def ids(self) -> np.ndarray:
        bits = np.unpackbits(self.words().view(np.uint8), bitorder="little")
        local = np.flatnonzero(bits).astype(np.uint64)
        return (local + (np.uint64(self.table.first_chunk) << np.uint64(16))).astype(np.uint32)

    def close(self):
        if self._h:
            lib().silo_host_filter_free(self._h)
            self._h = None

    def __del__(self):
        if sys is None or sys.is_finalizing():  # (the CUDA runtime's own exit handlers may have run: the process frees everything anyway)
            return
        try:
            self.close()
        except Exception:
            pass


class PreparedFilter:
    """A filter whose program is resident on the device; run_async() only enqueues its kernel."""

    def __init__(self, table: "HostTable", handle):
        self.table = table
        self._h = handle
        table._children.add(self)

    def run_async(self, stream_ptr: int) -> None:
        _check(lib().silo_host_prepared_run_async(self._h, C.c_void_p(stream_ptr)))

    def run_counts_async(self, column_index: int, d_counts_ptr: int, stream_ptr: int) -> None:
        """The filter and the Mutations counts of one column, enqueued only (one launch less than run_async +
        HostTable.mutation_counts_async: the interpreter also prepares the counts kernels)."""
        _check(lib().silo_host_prepared_run_counts_async(self._h, column_index, C.c_void_p(d_counts_ptr), C.c_void_p(stream_ptr)))

    def run_sharded_async(self, stream_ptr: int) -> None:
        """The filter and the counts of the shard group's column, this rank's rows sent to rank 0; enqueue only and
        replayable inside a CUDA graph (silo_gpu_program_run_sharded_async)."""
        _check(lib().silo_host_prepared_run_sharded_async(self._h, C.c_void_p(stream_ptr)))

    def run_sharded_collect_async(self, stream_ptr: int, d_summed_counts_ptr: int = 0) -> None:
        """Rank 0: run_sharded_async with the collect inside -- the finalize kernel waits for the other ranks' rows and
        writes the sums (silo_gpu_program_run_sharded_collect_async)."""
        _check(lib().silo_host_prepared_run_sharded_collect_async(self._h, C.c_void_p(d_summed_counts_ptr), C.c_void_p(stream_ptr)))

    @property
    def device_handle(self) -> int:
        return lib().silo_host_prepared_filter(self._h)

    @property
    def staged_bytes(self) -> int:
        return int(lib().silo_host_prepared_staged_bytes(self._h))

    def cardinality(self) -> int:
        value = C.c_uint64()
        abi.check(abi.lib().silo_gpu_filter_cardinality(self.device_handle, C.byref(value)))
        return int(value.value)

    def close(self):
        if self._h:
            lib().silo_host_prepared_free(self._h)
            self._h = None

    def __del__(self):
        if sys is None or sys.is_finalizing():  # (the CUDA runtime's own exit handlers may have run: the process frees everything anyway)
            return
        try:
            self.close()
        except Exception:
            pass


def _archive_specs(columns):
    """columns: (name, NUCLEOTIDE | AMINO_ACID, reference) in the archive's order"""
    n = len(columns)
    names = (C.c_char_p * n)(*[name.encode() for name, _, _ in columns])
    alphabets = (C.c_int * n)(*[alphabet for _, alphabet, _ in columns])
    references = (C.c_char_p * n)(*[reference.encode() for _, _, reference in columns])
    return names, alphabets, references, n


def roaring_runs(portable_roaring_bytes: bytes) -> list[tuple[int, int]]:
    """ascending (first, end_exclusive) runs of a portable roaring bitmap, decoded by the host layer"""
    capacity = 1 << 16
    while True:
        runs = (C.c_uint32 * (2 * capacity))()
        n = lib().silo_host_roaring_runs(portable_roaring_bytes, len(portable_roaring_bytes), runs, capacity)
        if n >= 0:
            return [(runs[2 * i], runs[2 * i + 1]) for i in range(n)]
        message = lib().silo_host_last_error().decode()
        if "too small" not in message:
            raise HostError(message)
        capacity *= 16


class Archive:
    """The sequence columns of a `.silo` table file in the S1 upload format (host/silo_loader.h). Host only."""

    def __init__(self, data: bytes, columns):
        self._data = data
        self.specs = list(columns)
        names, alphabets, references, n = _archive_specs(self.specs)
        self._h = lib().silo_host_archive_read(data, len(data), names, alphabets, references, n)
        if not self._h:
            raise HostError(lib().silo_host_last_error().decode())

    def desc(self, index: int):
        """ctypes pointer to the column's silo_column_desc (valid until close())"""
        pointer = lib().silo_host_archive_column(self._h, index)
        if not pointer:
            raise IndexError(index)
        return pointer

    def shard_desc(self, index: int, first_chunk: int, n_chunks: int):
        """the column restricted to the chunks [first_chunk, first_chunk + n_chunks) (global v_index / row ids)"""
        pointer = lib().silo_host_archive_column_shard(self._h, index, first_chunk, n_chunks)
        if not pointer:
            raise HostError(lib().silo_host_last_error().decode())
        return pointer

    def info(self, index: int) -> dict:
        values = (C.c_uint64 * 6)()
        _check(lib().silo_host_archive_column_info(self._h, index, values))
        keys = ("n_chunks", "sequence_count", "n_insertion_positions", "vertical_bitmaps_size", "horizontal_bitmaps_size", "num_chunks")
        return dict(zip(keys, (int(v) for v in values)))

    def chunk_sizes(self, index: int) -> list[int]:
        n = self.info(index)["n_chunks"]
        out = (C.c_uint32 * max(n, 1))()
        _check(lib().silo_host_archive_chunk_sizes(self._h, index, out, n))
        return [int(out[i]) for i in range(n)]

    def close(self):
        if self._h:
            lib().silo_host_archive_free(self._h)
            self._h = None

    def __del__(self):
        if sys is None or sys.is_finalizing():  # (the CUDA runtime's own exit handlers may have run: the process frees everything anyway)
            return
        try:
            self.close()
        except Exception:
            pass


def _dimension_spec(dimensions: Sequence) -> bytes:
    parts = []
    for dim in dimensions:
        if dim[0] == "position":
            parts.append(f"p:{dim[1]}:{int(dim[2])}")
        else:
            parts.append("b:" + ",".join(f"{value}={name}" for value, name in dim[1]) + "|" + (dim[2] or ""))
    return ";".join(parts).encode()


def combination_rows_from_columns(codes: np.ndarray, counts: np.ndarray) -> list[tuple]:
    """The tuple rows of bitmap_aggregation() from the arrays of bitmap_aggregation_columns(), for dimensions that are all
    sequence positions (a code is the symbol's character, 0 = null)."""
    return [tuple(chr(c) if c else None for c in row) + (int(count),) for row, count in zip(codes.tolist(), counts.tolist())]


def _combination_rows(text: str) -> list[tuple]:
    rows = []
    for line in text.splitlines():
        fields = line.split("\t")
        rows.append(tuple(None if f == "\\N" else f for f in fields[:-1]) + (int(fields[-1]),))
    return rows


class HostTable:
    """rhydb::storage::Table as the query compiler sees it, with its sequence columns in HBM."""

    @classmethod
    def from_archive(cls, ctx: Optional[abi.Context], data: bytes, columns, first_chunk: int = 0, n_chunks: Optional[int] = None) -> "HostTable":
        """S1 for a saved database (silo_host_table_load_archive): row layout and columns come from the `.silo` bytes;
        a rank of the row-partitioned table passes its chunk range. ctx None: host-only table."""
        columns = list(columns)
        archive = Archive(data, columns)
        try:
            chunk_sizes = archive.chunk_sizes(0)
        finally:
            archive.close()
        if n_chunks is None:
            n_chunks = len(chunk_sizes) - first_chunk
        names, alphabets, references, n = _archive_specs(columns)
        handle = lib().silo_host_table_load_archive(ctx._h if ctx is not None else None, data, len(data), names, alphabets, references, n,
                                                    first_chunk, n_chunks)
        if not handle:
            raise HostError(lib().silo_host_last_error().decode())
        table = cls.__new__(cls)
        table._init_fields(ctx, chunk_sizes[first_chunk:first_chunk + n_chunks], first_chunk)
        table._h = handle
        for name, alphabet, reference in columns:
            table.columns[name] = (16 if alphabet == NUCLEOTIDE else 28, len(reference))
        return table

    def _init_fields(self, ctx: abi.Context, chunk_sizes: Sequence[int], first_chunk: int) -> None:
        self.ctx = ctx
        self.chunk_sizes = [int(s) for s in chunk_sizes]
        self.n_chunks = len(self.chunk_sizes)
        self.first_chunk = first_chunk
        self.columns: dict[str, tuple[int, int]] = {}
        self._children = weakref.WeakSet()  # live HostFilter / PreparedFilter objects of this table
        # mutations_columns: record-batch buffer (grown on demand) and the call's out-parameters
        self._packed = np.empty(1 << 16, dtype=np.uint8)
        self._packed_out = (C.c_uint64(), C.c_uint32(), C.c_uint64())
        self._name_arrays: dict = {}
        self._h = None

    def __init__(self, ctx: abi.Context, chunk_sizes: Sequence[int], first_chunk: int = 0):
        self._init_fields(ctx, chunk_sizes, first_chunk)
        arr = (C.c_uint32 * max(self.n_chunks, 1))(*self.chunk_sizes)
        # ctx None: a host-only table (metadata, query compiler front half; device work fails loudly)
        self._h = lib().silo_host_table_create(ctx._h if ctx is not None else None, first_chunk, arr, self.n_chunks)
        if not self._h:
            raise HostError(lib().silo_host_last_error().decode())

    def add_column(self, name: str, alphabet: int, reference: str, desc_ptr) -> None:
        _check(lib().silo_host_table_add_column(
            self._h, name.encode(), alphabet, reference.encode(), C.cast(desc_ptr, C.c_void_p)))
        self.columns[name] = (16 if alphabet == NUCLEOTIDE else 28, len(reference))

    def add_string_column(self, name: str, values: Sequence[Optional[str]]) -> None:
        """A string column without an index (one value per row in layout order, None = null): dictionary ids on the device."""
        dictionary = sorted({v for v in values if v is not None})
        index = {v: i for i, v in enumerate(dictionary)}
        ids = np.array([index[v] if v is not None else 0 for v in values], dtype=np.uint32)
        nulls = self._row_ids(np.flatnonzero(np.array([v is None for v in values], dtype=bool)))
        names = (C.c_char_p * max(len(dictionary), 1))(*[v.encode() for v in dictionary])
        _check(lib().silo_host_table_add_string_column(self._h, name.encode(), names, len(dictionary), ids.ctypes.data, nulls.ctypes.data, nulls.size))

    def add_string_column_ids(self, name: str, dictionary: Sequence[str], ids: np.ndarray) -> None:
        """the same from ready-made dictionary ids (no nulls)"""
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        names = (C.c_char_p * max(len(dictionary), 1))(*[v.encode() for v in dictionary])
        _check(lib().silo_host_table_add_string_column(self._h, name.encode(), names, len(dictionary), ids.ctypes.data, None, 0))

    def add_date_column(self, name: str, days: Sequence[Optional[int]]) -> None:
        """A Date32 column (days since the epoch per row in layout order, None = null)."""
        if isinstance(days, np.ndarray):
            values, nulls = np.ascontiguousarray(days, dtype=np.int32), np.zeros(0, dtype=np.uint32)
        else:
            values = np.array([d if d is not None else 0 for d in days], dtype=np.int32)
            nulls = self._row_ids(np.flatnonzero(np.array([d is None for d in days], dtype=bool)))
        _check(lib().silo_host_table_add_date_column(self._h, name.encode(), values.ctypes.data, nulls.ctypes.data if nulls.size else None, nulls.size))

    def _row_ids(self, dense_rows: np.ndarray) -> np.ndarray:
        """dense row numbers (layout order) -> global row ids (chunk << 16 | row)"""
        starts = np.concatenate([[0], np.cumsum(self.chunk_sizes)])
        chunks = np.searchsorted(starts, dense_rows, side="right") - 1
        return (((chunks + self.first_chunk).astype(np.uint64) << np.uint64(16)) | (dense_rows - starts[chunks]).astype(np.uint64)).astype(np.uint32)

    def to_strings(self, expression: str) -> tuple[str, str, str]:
        """toString() of the parsed expression, the rewritten expression and the compiled operator tree"""
        buf = C.create_string_buffer(1 << 22)
        _check(lib().silo_host_filter_to_string(self._h, expression.encode(), buf, len(buf)))
        parsed, rewritten, compiled = buf.value.decode().split("\n")[:3]
        return parsed, rewritten, compiled

    def program_bitmap(self, expression: str, index: int = 0) -> bytes:
        """portable roaring bytes of bitmap `index` of the expression's lowered program"""
        buf = C.create_string_buffer(1 << 22)
        size = lib().silo_host_filter_program_bitmap(self._h, expression.encode(), index, buf, len(buf))
        if size < 0:
            raise HostError(lib().silo_host_last_error().decode())
        return buf.raw[:size]

    def lower_timed(self, expression: str) -> dict:
        """parse -> rewrite -> compile -> lower only: phase times (us), program sizes and a digest of the program"""
        phases, sizes, digest = (C.c_double * 4)(), (C.c_uint64 * 3)(), C.c_uint64()
        _check(lib().silo_host_filter_lower_timed(self._h, expression.encode(), phases, sizes, C.byref(digest)))
        return {"parse_us": phases[0], "rewrite_us": phases[1], "compile_us": phases[2], "lower_us": phases[3],
                "n_instrs": int(sizes[0]), "blob_bytes": int(sizes[1]), "n_bitmaps": int(sizes[2]), "digest": int(digest.value)}

    def register_bitmap(self, name: str, portable_roaring_bytes: bytes, resident: bool = True) -> None:
        """resident: a static index bitmap, uploaded once (silo_gpu_bitmap_register); otherwise the
        bytes travel with every program that uses the bitmap (PUSH_BITMAP)."""
        _check(lib().silo_host_table_register_bitmap(
            self._h, name.encode(), portable_roaring_bytes, len(portable_roaring_bytes), 1 if resident else 0))

    @property
    def num_rows(self) -> int:
        return int(lib().silo_host_table_num_rows(self._h))

    @property
    def device_table(self) -> int:
        return lib().silo_host_table_device(self._h)

    def filter(self, expression: str) -> HostFilter:
        handle = lib().silo_host_filter_eval(self._h, expression.encode())
        if not handle:
            raise HostError(lib().silo_host_last_error().decode())
        return HostFilter(self, handle)

    def count(self, expression: Optional[str]) -> int:
        """CountFilterNode: the number of rows that pass, one device call (silo_gpu_query_count)"""
        out = C.c_uint64()
        _check(lib().silo_host_count(self._h, expression.encode() if expression is not None else None, C.byref(out)))
        return int(out.value)

    def prepare(self, expression: str) -> PreparedFilter:
        handle = lib().silo_host_filter_prepare(self._h, expression.encode())
        if not handle:
            raise HostError(lib().silo_host_last_error().decode())
        return PreparedFilter(self, handle)

    def explain(self, expression: str) -> str:
        buf = C.create_string_buffer(1 << 22)
        _check(lib().silo_host_filter_explain(self._h, expression.encode(), buf, len(buf)))
        return buf.value.decode()

    def _text_buffer(self):
        # (allocated once: ctypes zero-fills a fresh buffer, a millisecond for 16 MB)
        if getattr(self, "_text", None) is None:
            self._text = C.create_string_buffer(1 << 24)
        return self._text

    def bitmap_aggregation(self, dimensions: Sequence, expression: Optional[str] = None) -> list[tuple]:
        """BitmapAggregationNode (co-occurrence / groupBy): dimensions are ("position", column, position0) or
        ("bitmaps", [(value, bitmap name), ...], null bitmap name or None). Returns the combinations in the
        reference's output order as (value-or-None per dimension ..., count) tuples."""
        buf = self._text_buffer()
        _check(lib().silo_host_bitmap_aggregation(
            self._h, expression.encode() if expression else None, _dimension_spec(dimensions), buf, len(buf)))
        return _combination_rows(buf.value.decode())

    def _packed_buffers(self, n_dims: int, rows: int):
        have = getattr(self, "_aggregation_buffers", None)
        if have is None or have[0].shape[0] < rows or have[0].shape[1] != max(1, n_dims):
            have = self._aggregation_buffers = (np.empty((rows, max(1, n_dims)), dtype=np.uint8), np.empty(rows, dtype=np.uint64))
        return have

    def bitmap_aggregation_columns(self, dimensions: Sequence, expression: Optional[str] = None) -> tuple[np.ndarray, np.ndarray]:
        """bitmap_aggregation() with the result as arrays: codes [n, n_dims] uint8 (a sequence position's symbol character
        or an indexed dimension's index into its sorted values; 0 / 255 = null) and counts [n] uint64 -- one call, no text."""
        spec = _dimension_spec(dimensions)
        n = C.c_uint64()
        codes, counts = self._packed_buffers(len(dimensions), 4096)
        for _ in range(2):
            _check(lib().silo_host_bitmap_aggregation_packed(
                self._h, expression.encode() if expression else None, spec, codes.ctypes.data, counts.ctypes.data, len(counts), n))
            if n.value <= len(counts):
                break
            codes, counts = self._packed_buffers(len(dimensions), int(n.value))
        return codes[:n.value, :len(dimensions)].copy(), counts[:n.value].copy()

    def bitmap_aggregation_merge_columns(self, dimensions: Sequence, shards: Sequence[tuple[np.ndarray, int]]) -> tuple[np.ndarray, np.ndarray]:
        """bitmap_aggregation_merge() with the result as arrays (see bitmap_aggregation_columns)."""
        pairs = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.uint64).reshape(-1, 2) for p, _ in shards]))
        sizes = np.array([len(p) for p, _ in shards], dtype=np.uint64)
        cardinalities = np.array([c for _, c in shards], dtype=np.uint64)
        n = C.c_uint64()
        codes, counts = self._packed_buffers(len(dimensions), max(4096, len(pairs) + 1))
        _check(lib().silo_host_bitmap_aggregation_merge_packed(
            self._h, _dimension_spec(dimensions), pairs.ctypes.data, sizes.ctypes.data, cardinalities.ctypes.data, len(shards),
            codes.ctypes.data, counts.ctypes.data, len(counts), n))
        return codes[:n.value, :len(dimensions)].copy(), counts[:n.value].copy()

    def bitmap_aggregation_shard(self, dimensions: Sequence, expression: Optional[str] = None) -> tuple[np.ndarray, int]:
        """One rank's half of the aggregation over a row-partitioned table (BitmapAggregationNode::executeShard): the
        (key, count) pairs of this shard ordered by key, [n, 2] uint64, and the shard's filter cardinality."""
        pairs = np.empty((4096, 2), dtype=np.uint64)
        n, cardinality = C.c_uint64(), C.c_uint64()
        for _ in range(2):
            _check(lib().silo_host_bitmap_aggregation_shard(
                self._h, expression.encode() if expression else None, _dimension_spec(dimensions), pairs.ctypes.data, len(pairs), n, cardinality))
            if n.value <= len(pairs):
                break
            pairs = np.empty((n.value, 2), dtype=np.uint64)
        return pairs[:n.value].copy(), int(cardinality.value)

    def bitmap_aggregation_merge(self, dimensions: Sequence, shards: Sequence[tuple[np.ndarray, int]]) -> list[tuple]:
        """The collecting rank: the ranks' (pairs, cardinality) results of bitmap_aggregation_shard summed per key and
        materialised (BitmapAggregationNode::mergeShards + ::materialise); same rows as bitmap_aggregation on the whole table."""
        pairs = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.uint64).reshape(-1, 2) for p, _ in shards]))
        sizes = np.array([len(p) for p, _ in shards], dtype=np.uint64)
        cardinalities = np.array([c for _, c in shards], dtype=np.uint64)
        buf = self._text_buffer()
        _check(lib().silo_host_bitmap_aggregation_merge(
            self._h, _dimension_spec(dimensions), pairs.ctypes.data, sizes.ctypes.data, cardinalities.ctypes.data, len(shards), buf, len(buf)))
        return _combination_rows(buf.value.decode())

    def mutation_counts(self, column: str, flt: Optional[HostFilter] = None) -> np.ndarray:
        n_symbols, genome_length = self.columns[column]
        out = np.zeros(n_symbols * genome_length, dtype=np.uint32)
        _check(lib().silo_host_mutation_counts(
            self._h, column.encode(), flt._h if flt is not None else None, out.ctypes.data_as(C.POINTER(C.c_uint32))))
        return out.reshape(n_symbols, genome_length)

    def mutations(self, columns: Sequence[str], expression: Optional[str], min_proportion: float) -> list[dict]:
        names = (C.c_char_p * len(columns))(*[c.encode() for c in columns])
        handle = lib().silo_host_mutations(
            self._h, expression.encode() if expression is not None else None, names, len(columns), min_proportion)
        return _rows(handle)

    def mutations_columns(self, columns: Sequence[str], expression: Optional[str], min_proportion: float) -> dict:
        """mutations() with the result as columns (numpy arrays over one record-batch buffer) instead of a
        list of row dicts: ONE call into the host library per query (silo_host_mutations_packed)."""
        names = self._name_arrays.get(tuple(columns))
        if names is None:
            names = self._name_arrays[tuple(columns)] = (C.c_char_p * len(columns))(*[c.encode() for c in columns])
        n_rows, n_names, needed = self._packed_out
        status = lib().silo_host_mutations_packed(
            self._h, expression.encode() if expression is not None else None, names, len(columns), min_proportion,
            self._packed.ctypes.data, self._packed.nbytes, n_rows, n_names, needed)
        _check(status)
        if needed.value > self._packed.nbytes:
            self._packed = np.empty(int(needed.value) * 2, dtype=np.uint8)
            _check(lib().silo_host_packed_fetch(self._packed.ctypes.data, self._packed.nbytes))
        return _unpack_record_batch(self._packed[:needed.value].copy(), int(n_rows.value), int(n_names.value))

    def mutations_enqueue(self, column: str, expression: Optional[str], d_counts_ptr: int, stream_ptr: int) -> None:
        """First half of the Mutations query on a row-partitioned table: parse + compile against this shard,
        then program upload, filter and counts enqueued on the stream (nothing synchronised). The scheduler
        all-reduces the counts on the same stream next."""
        _check(lib().silo_host_mutations_enqueue(
            self._h, expression.encode() if expression is not None else None, column.encode(), d_counts_ptr, stream_ptr))

    def mutations_collect(self, column: str, min_proportion: float, d_summed_counts_ptr: int, stream_ptr: int) -> tuple[dict, int]:
        """Second half, on one rank: the output pass over the summed counts on the device, rows back as one
        record batch. Returns (columns, number of this shard's rows that passed the filter)."""
        n_rows, n_names, needed = self._packed_out
        cardinality = C.c_uint64()
        _check(lib().silo_host_mutations_collect_packed(
            self._h, column.encode(), min_proportion, d_summed_counts_ptr, stream_ptr,
            self._packed.ctypes.data, self._packed.nbytes, n_rows, n_names, needed, cardinality))
        if needed.value > self._packed.nbytes:
            self._packed = np.empty(int(needed.value) * 2, dtype=np.uint8)
            _check(lib().silo_host_packed_fetch(self._packed.ctypes.data, self._packed.nbytes))
        return _unpack_record_batch(self._packed[:needed.value].copy(), int(n_rows.value), int(n_names.value)), int(cardinality.value)

    def shard_group_create(self, column: str, rank: int, world: int) -> bytes:
        """This table as rank `rank` of a row-partitioned table of `world` shards: returns this rank's handle
        (SHARD_HANDLE_BYTES), to be exchanged with all ranks (silo_gpu_shard_group_init)."""
        handle = C.create_string_buffer(SHARD_HANDLE_BYTES)
        _check(lib().silo_host_shard_group_create(self._h, column.encode(), rank, world, handle))
        return handle.raw

    def shard_group_connect(self, handles_by_rank: Sequence[bytes]) -> None:
        joined = b"".join(handles_by_rank)
        _check(lib().silo_host_shard_group_connect(self._h, joined, len(joined)))

    def sharded_enqueue(self, column: str, expression: Optional[str], stream_ptr: int) -> None:
        """Every rank: parse + compile against this shard, filter + counts + this rank's rows to rank 0; enqueue only."""
        _check(lib().silo_host_sharded_enqueue(self._h, expression.encode() if expression is not None else None, column.encode(), stream_ptr))

    def sharded_collect(self, column: str, min_proportion: float, stream_ptr: int, d_summed_counts_ptr: int = 0) -> tuple[dict, int]:
        """Rank 0: the rows of the whole table (columns as mutations_columns) and the filter cardinality over all shards."""
        n_rows, n_names, needed = self._packed_out
        cardinality = C.c_uint64()
        _check(lib().silo_host_sharded_collect_packed(
            self._h, column.encode(), min_proportion, d_summed_counts_ptr or None, stream_ptr,
            self._packed.ctypes.data, self._packed.nbytes, n_rows, n_names, needed, cardinality))
        if needed.value > self._packed.nbytes:
            self._packed = np.empty(int(needed.value) * 2, dtype=np.uint8)
            _check(lib().silo_host_packed_fetch(self._packed.ctypes.data, self._packed.nbytes))
        return _unpack_record_batch(self._packed[:needed.value].copy(), int(n_rows.value), int(n_names.value)), int(cardinality.value)

    def sharded_query(self, column: str, expression: Optional[str], min_proportion: float, d_summed_counts_ptr: int = 0) -> tuple[dict, int]:
        """Rank 0: sharded_enqueue + sharded_collect as ONE device call on the table's own stream (a replayed graph)."""
        n_rows, n_names, needed = self._packed_out
        cardinality = C.c_uint64()
        _check(lib().silo_host_sharded_query_packed(
            self._h, expression.encode() if expression is not None else None, column.encode(), min_proportion, d_summed_counts_ptr or None,
            self._packed.ctypes.data, self._packed.nbytes, n_rows, n_names, needed, cardinality))
        if needed.value > self._packed.nbytes:
            self._packed = np.empty(int(needed.value) * 2, dtype=np.uint8)
            _check(lib().silo_host_packed_fetch(self._packed.ctypes.data, self._packed.nbytes))
        return _unpack_record_batch(self._packed[:needed.value].copy(), int(n_rows.value), int(n_names.value)), int(cardinality.value)

    def sharded_collect_async(self, stream_ptr: int, d_summed_counts_ptr: int = 0) -> None:
        """Rank 0, device-resident pipeline: wait for all ranks on the device, sum, hand the slot back; enqueue only."""
        _check(lib().silo_host_sharded_collect_async(self._h, d_summed_counts_ptr or None, stream_ptr))

    def mutation_columns_from_counts(self, column: str, counts: np.ndarray, min_proportion: float) -> dict:
        counts = np.ascontiguousarray(counts, dtype=np.uint32)
        handle = lib().silo_host_mutation_rows_from_counts(
            self._h, column.encode(), counts.ctypes.data_as(C.POINTER(C.c_uint32)), min_proportion)
        return _columns(handle)

    def mutation_rows_from_counts(self, column: str, counts: np.ndarray, min_proportion: float) -> list[dict]:
        counts = np.ascontiguousarray(counts, dtype=np.uint32)
        handle = lib().silo_host_mutation_rows_from_counts(
            self._h, column.encode(), counts.ctypes.data_as(C.POINTER(C.c_uint32)), min_proportion)
        return _rows(handle)

    @staticmethod
    def last_query_profile() -> dict:
        """Microseconds per phase of this thread's last mutations() call."""
        out = (C.c_double * 5)()
        lib().silo_host_last_query_profile(out)
        return dict(zip(("parse_us", "compile_us", "filter_us", "counts_us", "threshold_us"), out))

    def set_option(self, name: str, value: int) -> None:
        """silo_gpu_table_set_option on the device table (e.g. "sweep_min_pieces")"""
        abi.check(abi.lib().silo_gpu_table_set_option(self.device_table, name.encode(), value))

    def sweep_stats(self) -> tuple[float, int, int]:
        """(mean ms per launch, algorithmic bytes per launch, launches timed) of the threshold sweep kernel"""
        ms, nbytes, calls = C.c_float(), C.c_uint64(), C.c_uint64()
        abi.check(abi.lib().silo_gpu_get_sweep_stats(self.device_table, C.byref(ms), C.byref(nbytes), C.byref(calls)))
        return float(ms.value), int(nbytes.value), int(calls.value)

    def stats(self) -> abi.Stats:
        out = abi.Stats()
        abi.check(abi.lib().silo_gpu_get_stats(self.device_table, C.byref(out)))
        return out

    def mutation_counts_async(self, column_index: int, flt, d_counts_ptr: int, stream_ptr: int) -> None:
        """flt: HostFilter | PreparedFilter | None (all rows). Enqueues only; no host synchronisation."""
        abi.check(abi.lib().silo_gpu_mutation_counts_async(
            self.device_table, column_index, flt.device_handle if flt is not None else None,
            C.c_void_p(d_counts_ptr), C.c_void_p(stream_ptr)))

    def close(self):
        if self._h:
            for child in list(self._children):
                child.close()
            lib().silo_host_table_free(self._h)
            self._h = None

    def __del__(self):
        if sys is None or sys.is_finalizing():  # (the CUDA runtime's own exit handlers may have run: the process frees everything anyway)
            return
        try:
            self.close()
        except Exception:
            pass


class Synthetic:
    """performance/sequence_generator.h restated on the product side (host/synthetic.h)."""

    def __init__(self, genome_length: int = 29903, reference_seed: int = 1, generations: int = 5, gene: bool = False,
                 tree_seed: int = 42, mutation_rate: float = 0.003, co_occurrence_sequences: int = 0):
        """gene: an amino-acid gene (valid-symbol mutations, its own tree seed and rate) instead of a nucleotide genome;
        co_occurrence_sequences > 0: the table of performance/co_occurrence_benchmark.cpp instead (100-nt reference,
        that many independently mutated sequences; the other arguments are ignored)"""
        if co_occurrence_sequences > 0:
            self._h = lib().silo_host_synthetic_create_co_occurrence(co_occurrence_sequences)
            genome_length = 100
        elif gene:
            self._h = lib().silo_host_synthetic_create_gene(genome_length, reference_seed, tree_seed, mutation_rate, generations)
        else:
            self._h = lib().silo_host_synthetic_create(genome_length, reference_seed, generations)
        if not self._h:
            raise HostError(lib().silo_host_last_error().decode())
        self.genome_length = genome_length

    @property
    def reference(self) -> str:
        return lib().silo_host_synthetic_reference(self._h).decode()

    @property
    def num_sequences(self) -> int:
        return int(lib().silo_host_synthetic_num_sequences(self._h))

    def sequence(self, index: int) -> str:
        return lib().silo_host_synthetic_sequence(self._h, index).decode()

    def parent(self, index: int) -> int:
        return int(lib().silo_host_synthetic_parent(self._h, index))

    def generation(self, index: int) -> int:
        return int(lib().silo_host_synthetic_generation(self._h, index))

    def build_column(self, total_rows: int, first_chunk: int, n_chunks: int, threads: int = 8, stride: int = 1):
        """Returns POINTER(ColumnDesc), valid until release_column()/the next build. The shard holds the
        chunks first_chunk + k*stride; stride > 1 (an interleaved shard) uses shard-local chunk ids."""
        out = C.c_void_p()
        _check(lib().silo_host_synthetic_build_column(self._h, total_rows, first_chunk, n_chunks, threads, C.byref(out), stride))
        return C.cast(out, C.POINTER(abi.ColumnDesc))

    def draw_short_reads(self, count: int, read_length: int) -> np.ndarray:
        """ShortReadGenerator (uniform tiling): the evolved-sequence index of every read, in id order; read i starts at
        i * (L - read_length + 1) // count. Must precede build_short_read_column."""
        out = np.empty(count, dtype=np.uint32)
        _check(lib().silo_host_synthetic_draw_short_reads(self._h, count, read_length, out.ctypes.data))
        return out

    def build_short_read_column(self, first_chunk: int, n_chunks: int, threads: int = 8):
        out = C.c_void_p()
        _check(lib().silo_host_synthetic_build_short_read_column(self._h, first_chunk, n_chunks, threads, C.byref(out)))
        return C.cast(out, C.POINTER(abi.ColumnDesc))

    def release_column(self) -> None:
        lib().silo_host_synthetic_release_column(self._h)

    def lineage_bitmap(self, ancestor: int, total_rows: int, first_chunk: int, n_chunks: int, stride: int = 1) -> bytes:
        n = lib().silo_host_synthetic_lineage_bitmap(self._h, ancestor, total_rows, first_chunk, n_chunks, None, 0, stride)
        if n < 0:
            raise HostError(lib().silo_host_last_error().decode())
        buf = C.create_string_buffer(int(n))
        lib().silo_host_synthetic_lineage_bitmap(self._h, ancestor, total_rows, first_chunk, n_chunks, buf, n, stride)
        return buf.raw

    def close(self):
        if self._h:
            lib().silo_host_synthetic_free(self._h)
            self._h = None

    def __del__(self):
        if sys is None or sys.is_finalizing():  # (the CUDA runtime's own exit handlers may have run: the process frees everything anyway)
            return
        try:
            self.close()
        except Exception:
            pass


def date_ranges_expression(total_rows: int, span_days: int, from_day: int, to_day_inclusive: int,
                           first_chunk: int, n_chunks: int, stride: int = 1) -> str:
    buf = C.create_string_buffer(64 + 24 * max(n_chunks, 1))
    _check(lib().silo_host_synthetic_date_ranges(
        total_rows, span_days, from_day, to_day_inclusive, first_chunk, n_chunks, buf, len(buf), stride))
    return buf.value.decode()


def interleaved_shard(n_chunks_total: int, n_ranks: int, rank: int) -> tuple[int, int, int]:
    """(first_chunk, n_chunks, stride) of rank's shard when chunk c belongs to rank c % n_ranks: any
    filter on a sorted column (a date range selects a contiguous row range) then loads every rank
    alike, which contiguous chunk ranges do not."""
    count = (n_chunks_total - rank + n_ranks - 1) // n_ranks if rank < n_chunks_total else 0
    return rank, count, n_ranks


def shard_chunk_sizes(total_rows: int, first_chunk: int, n_chunks: int, stride: int = 1) -> list[int]:
    sizes = dense_chunk_sizes(total_rows)
    return [sizes[first_chunk + k * stride] for k in range(n_chunks)]


def dense_chunk_sizes(total_rows: int) -> list[int]:
    sizes = [65536] * (total_rows // 65536)
    if total_rows % 65536:
        sizes.append(total_rows % 65536)
    return sizes


def partition_chunks(chunk_weights: Sequence[int], n_ranks: int) -> list[int]:
    weights = (C.c_uint64 * max(len(chunk_weights), 1))(*chunk_weights)
    out = (C.c_uint32 * (n_ranks + 1))()
    _check(lib().silo_host_partition_chunks(weights, len(chunk_weights), n_ranks, out))
    return list(out)
