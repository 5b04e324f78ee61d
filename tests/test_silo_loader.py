"""The `.silo` -> device loader of the host layer (lapis_silo_b200/host/silo_loader.cpp).

  * (where /root/reference is mounted) the C++ reader, run on the reference's OWN serialised state
    (testBaseData/siloSerializedState, database.test.cpp:100-116), gives the committed extract
    tests/golden/silo_state_unit_test_dummy.json; the test-side archive writer reproduces that file's bytes;
  * everywhere: archives written from the golden extract and from oracle-built tables (several chunks, null rows,
    N runs, array / bitset / run containers) come back from the C++ reader field by field;
  * the host's portable-roaring decoder against the oracle's roaring serialiser;
  * GPU: a table loaded from archive bytes (silo_host_table_load_archive -> S1) answers filters and the Mutations
    action like the oracle built from the rows."""
import ctypes as C
import json
import os
import struct

import numpy as np
import pytest

import silo_archive_writer as W

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "silo_state_unit_test_dummy.json")) as handle:
    STATE = json.load(handle)
STATE_FILE = "/root/reference/testBaseData/siloSerializedState/1785915539/default.silo"


def specs_of_state():
    from lapis_silo_b200 import host_api as H
    return [(c["name"], H.NUCLEOTIDE if c["alphabet"] == "Nucleotide" else H.AMINO_ACID, c["reference"]) for c in STATE["columns"]]


def state_columns_for_writer():
    columns = []
    for c in STATE["columns"]:
        column = {
            "alphabet": c["alphabet"], "local_reference": c["local_reference"],
            "containers": [dict(k, payload=bytes.fromhex(k["payload_hex"])) for k in c["containers"]],
            "missing_bitmaps": {int(row): bytes.fromhex(blob) for row, blob in c["missing_bitmaps"].items()},
            "start_end": [[tuple(pair) for pair in chunk] for chunk in c["start_end"]],
            "batch_start_ends": [tuple(pair) for pair in c["batch_start_ends"]],
        }
        column.update(sequence_count=c["sequence_count"], vertical_bitmaps_size=c["vertical_bitmaps_size"],
                      horizontal_bitmaps_size=c["horizontal_bitmaps_size"], null_bitmap=bytes.fromhex(c["null_bitmap_hex"]),
                      num_chunks=c["num_chunks"], insertion_bucket_counts=c["insertion_bucket_counts"],
                      insertion_positions=[dict(p, insertions=[(value, bytes.fromhex(blob)) for value, blob in p["insertions"]])
                                           for p in c["insertion_positions"]])  # column E has one (4:EPE)
        columns.append(column)
    return columns


def desc_as_values(desc, n_rows):
    """a silo_column_desc as plain Python values"""
    containers = []
    for i in range(desc.n_containers):
        c = desc.containers[i]
        containers.append({"position": c.position, "v_index": c.v_index, "symbol": c.symbol, "cardinality": c.cardinality,
                           "typecode": c.typecode, "payload_hex": C.string_at(C.addressof(desc.payload.contents) + c.payload_offset, c.payload_bytes).hex()})
    start_end = [(desc.start_end[2 * r], desc.start_end[2 * r + 1]) for r in range(n_rows)]
    missing = {}
    for i in range(desc.n_rows_with_missing):
        missing[int(desc.missing_row_ids[i])] = [(desc.missing_runs[2 * run], desc.missing_runs[2 * run + 1])
                                                 for run in range(desc.missing_offsets[i], desc.missing_offsets[i + 1])]
    return {
        "n_symbols": desc.n_symbols, "genome_length": desc.genome_length, "missing_symbol": desc.missing_symbol,
        "local_reference": [desc.local_reference[p] for p in range(desc.genome_length)],
        "containers": containers, "start_end": start_end, "missing": missing,
        "null_rows": [int(desc.null_row_ids[i]) for i in range(desc.n_null_rows)],
    }


def runs_of(values):
    runs = []
    for v in values:
        if runs and runs[-1][1] == v:
            runs[-1][1] = v + 1
        else:
            runs.append([v, v + 1])
    return [tuple(r) for r in runs]


def roaring_values(blob: bytes) -> list[int]:
    """portable roaring format without run containers (cookie 12346), enough for the fixture's bitmaps"""
    cookie, n = struct.unpack_from("<II", blob, 0)
    assert cookie == 12346
    keys = [struct.unpack_from("<HH", blob, 8 + 4 * i) for i in range(n)]
    cursor = 8 + 8 * n
    out = []
    for key, cardinality_minus_one in keys:
        count = cardinality_minus_one + 1
        assert count <= 4096
        out += [(key << 16) | v for v in struct.unpack_from(f"<{count}H", blob, cursor)]
        cursor += 2 * count
    return out


def check_state_archive(data: bytes):
    from lapis_silo_b200 import host_api as H
    from oracle import oracle as O
    archive = H.Archive(data, specs_of_state())
    for index, column in enumerate(STATE["columns"]):
        chars = O.NUC_SYMBOLS if column["alphabet"] == "Nucleotide" else O.AA_SYMBOLS
        info = archive.info(index)
        assert archive.chunk_sizes(index) == [len(chunk) for chunk in column["start_end"]]
        got = desc_as_values(archive.desc(index).contents, sum(archive.chunk_sizes(index)))
        assert got["containers"] == column["containers"], column["name"]
        assert "".join(chars[s] for s in got["local_reference"]) == column["local_reference"]
        assert (got["n_symbols"], got["genome_length"], got["missing_symbol"]) == (len(chars), len(column["reference"]), len(chars) - 1)
        assert got["start_end"] == [tuple(pair) for chunk in column["start_end"] for pair in chunk]
        assert got["missing"] == {int(row): runs_of(roaring_values(bytes.fromhex(blob))) for row, blob in column["missing_bitmaps"].items()}
        assert info["n_insertion_positions"] == len(column["insertion_positions"])
        assert (info["sequence_count"], info["vertical_bitmaps_size"], info["horizontal_bitmaps_size"], info["num_chunks"]) == \
               (column["sequence_count"], column["vertical_bitmaps_size"], column["horizontal_bitmaps_size"], column["num_chunks"])
        assert got["null_rows"] == roaring_values(bytes.fromhex(column["null_bitmap_hex"]))
    archive.close()


@pytest.mark.skipif(not os.path.exists(STATE_FILE), reason="the reference tree is not mounted here")
def test_reader_on_the_reference_serialised_state():
    check_state_archive(open(STATE_FILE, "rb").read())


@pytest.mark.skipif(not os.path.exists(STATE_FILE), reason="the reference tree is not mounted here")
def test_writer_reproduces_the_reference_bytes():
    """pins the test-side writer: fed the extract of the reference's file it must emit exactly the bytes the reference
    wrote for the four sequence columns (E with its insertion index included)"""
    data = open(STATE_FILE, "rb").read()
    seen = {"roaring::Roaring", "pair<u32,Roaring>"}
    written = b"".join(W.sequence_column_bytes(column, seen) for column in state_columns_for_writer())
    at = data.find(written)
    assert at > 0, "the sequence columns as written by the test writer are not in the reference's file"
    assert written[:5] == bytes(5) and written[5:13] == struct.pack("<Q", 8)  # class info of the first sequence column, then "main"'s local reference


@pytest.mark.parametrize("roaring_seen,pair_seen", [(True, True), (False, False), (True, False), (False, True)])
def test_reader_on_the_rewritten_state(roaring_seen, pair_seen):
    metadata = b"\x05\0\0\0\0\0\0\0hello" + bytes(37) + struct.pack("<Q", 4) + b"ACGT" + b"\xff" * 9  # decoys that look like local references
    check_state_archive(W.write_archive(state_columns_for_writer(), metadata, roaring_seen, pair_seen))


def oracle_column_for_writer(table, name, alphabet_name, chars):
    from oracle import oracle as O
    export = table.export_column(name)
    values = desc_as_values(export.desc.contents, table.num_rows)
    export.close()
    chunk_sizes = table.chunk_sizes
    start_end, row = [], 0
    for size in chunk_sizes:
        start_end.append(values["start_end"][row:row + size])
        row += size
    missing = {}
    for index, (row_id, runs) in enumerate(values["missing"].items()):
        positions = [p for first, end in runs for p in range(first, end)]
        missing[row_id] = O.roaring_roundtrip(positions, optimize=(index % 2 == 0))[1]
    column = {
        "alphabet": alphabet_name, "local_reference": "".join(chars[s] for s in values["local_reference"]),
        "containers": [dict(c, payload=bytes.fromhex(c["payload_hex"])) for c in values["containers"]],
        "missing_bitmaps": missing, "start_end": start_end,
        "batch_start_ends": [(min((s for s, e in chunk), default=0), max((e for s, e in chunk), default=0)) for chunk in start_end],
        "sequence_count": table.num_rows, "vertical_bitmaps_size": 123, "horizontal_bitmaps_size": 456,
        "null_bitmap": O.roaring_roundtrip(values["null_rows"], optimize=True)[1] if values["null_rows"] else W.EMPTY_ROARING,
        "num_chunks": len(chunk_sizes),
    }
    return column, values


def random_archive(seed, alphabet_id):
    from oracle import oracle as O
    from test_gpu_kernels import random_table
    table = random_table(seed, 900, 60, flushes=(99, 130, 131), alphabet=alphabet_id)
    alphabet_name, chars = ("Nucleotide", O.NUC_SYMBOLS) if alphabet_id == O.NUCLEOTIDE else ("AminoAcid", O.AA_SYMBOLS)
    column, values = oracle_column_for_writer(table, "c", alphabet_name, chars)
    reference = table.columns[0][2]
    # an insertion index with several positions, insertions and three-mers (class info only in front of the first of each)
    rows = O.roaring_roundtrip([1, 5, 70000], optimize=False)[1]
    column["insertion_positions"] = [
        {"position": p, "insertions": [("ACGT", rows), ("TTTTTTTTTTTTTTTTTTTT", O.roaring_roundtrip(list(range(300)), optimize=True)[1])],
         "three_mers": [([1, 2, 3], [0]), ([2, 3, 4], [0, 1]), ([4, 4, 4], [1])], "three_mer_buckets": 13} for p in (3, 17, 60)]
    column["insertion_bucket_counts"] = (5, 1)
    return table, W.write_archive([column], b"\x01\x02" * 50), values, [("c", alphabet_id, reference)]


@pytest.mark.parametrize("seed,alphabet_id", [(301, 0), (302, 1)])
def test_reader_on_oracle_built_tables(seed, alphabet_id):
    from lapis_silo_b200 import host_api as H
    table, data, want, specs = random_archive(seed, alphabet_id)
    assert len(table.chunk_sizes) == 4 and want["null_rows"] and want["missing"]
    assert {c["typecode"] for c in want["containers"]} >= {2}
    archive = H.Archive(data, specs)
    assert archive.chunk_sizes(0) == table.chunk_sizes
    assert archive.info(0) == {"n_chunks": 4, "sequence_count": table.num_rows, "n_insertion_positions": 3, "vertical_bitmaps_size": 123,
                               "horizontal_bitmaps_size": 456, "num_chunks": 4}
    assert desc_as_values(archive.desc(0).contents, table.num_rows) == want
    archive.close()


@pytest.mark.parametrize("seed,alphabet_id", [(301, 0), (302, 1)])
def test_shards_of_an_archive_equal_the_oracle_shard_exports(seed, alphabet_id):
    """one rank's part of the row-partitioned table: chunk range of the archive's column == the
    oracle's export of the same chunk range (global v_index and row ids)"""
    from lapis_silo_b200 import host_api as H
    table, data, _, specs = random_archive(seed, alphabet_id)
    archive = H.Archive(data, specs)
    for first_chunk, n_chunks in ((0, 4), (0, 1), (1, 2), (3, 1), (2, 0), (4, 0)):
        rows = sum(table.chunk_sizes[first_chunk:first_chunk + n_chunks])
        export = table.export_column("c", first_chunk, n_chunks)
        want = desc_as_values(export.desc.contents, rows)
        export.close()
        assert desc_as_values(archive.shard_desc(0, first_chunk, n_chunks).contents, rows) == want, (first_chunk, n_chunks)
    with pytest.raises(H.HostError, match="chunk range outside the column"):
        archive.shard_desc(0, 3, 2)
    archive.close()
    # a host-only table over a shard: layout and metadata come from the archive, the query compiler front half works
    shard = H.HostTable.from_archive(None, data, specs, first_chunk=1, n_chunks=2)
    assert shard.first_chunk == 1 and shard.chunk_sizes == table.chunk_sizes[1:3] and shard.num_rows == sum(table.chunk_sizes[1:3])
    assert "op=3" in shard.explain("(sym-eq c 3 -)")
    with pytest.raises(H.HostError, match=r"DeviceError\[-2\]"):
        shard.filter("(true)")
    shard.close()


def test_roaring_decoder_against_the_oracle_serialiser():
    from lapis_silo_b200 import host_api as H
    from oracle import oracle as O
    rng = np.random.default_rng(7)
    cases = [
        [], [0], [65535, 65536], list(range(10, 5000)), list(range(0, 65536)), list(range(65000, 140000)),
        sorted({int(v) for v in rng.integers(0, 1 << 16, 9000)}),                       # bitset container
        sorted({int(v) for v in rng.integers(0, 1 << 20, 3000)}),                       # 16 array containers
        [k << 16 | v for k in range(6) for v in range(100, 400)],                       # >= 4 run containers: offset header
        sorted({int(v) for v in rng.integers(0, 1 << 18, 2000)} | set(range(70000, 75000)) | set(range(3 << 16, (3 << 16) + 65536))),
    ]
    for values in cases:
        for optimize in (False, True):
            back, blob = O.roaring_roundtrip(values, optimize=optimize)
            assert back == values
            assert H.roaring_runs(blob) == runs_of(values)
    with pytest.raises(H.HostError):
        H.roaring_runs(O.roaring_roundtrip(list(range(100)), optimize=False)[1][:-3])
    with pytest.raises(H.HostError):
        H.roaring_runs(b"\x01\x02\x03\x04\x05\x06\x07\x08")


def test_reader_rejects_what_it_cannot_read():
    from lapis_silo_b200 import host_api as H
    good = W.write_archive(state_columns_for_writer())
    specs = specs_of_state()
    with pytest.raises(H.HostError, match="not a boost binary archive"):
        H.Archive(b"\x16\0\0\0\0\0\0\0serialization::archivX" + good[30:], specs)
    with pytest.raises(H.HostError, match="unsupported archive flavour"):
        H.Archive(good[:30] + b"\x13" + good[31:], specs)
    with pytest.raises(H.HostError):
        H.Archive(good[:len(good) - 7], specs)  # the last column's tail is cut off
    with pytest.raises(H.HostError, match="not found"):
        H.Archive(good, [("main", H.NUCLEOTIDE, "ACGTACGTACGTACGTACGTA")])
    # a container whose payload does not match its cardinality
    columns = state_columns_for_writer()
    columns[0]["containers"][0]["cardinality"] += 1
    with pytest.raises(H.HostError):
        H.Archive(W.write_archive(columns), specs)


@pytest.fixture(scope="module")
def ctx():
    from lapis_silo_b200 import abi
    context = abi.Context(0)
    yield context
    context.close()


@pytest.mark.gpu
@pytest.mark.parametrize("seed,alphabet_id", [(301, 0), (302, 1)])
def test_device_table_loaded_from_archive_bytes(ctx, seed, alphabet_id):
    from lapis_silo_b200 import host_api as H
    oracle_table, data, _, specs = random_archive(seed, alphabet_id)
    device_table = H.HostTable.from_archive(ctx, data, specs)
    assert device_table.num_rows == oracle_table.num_rows and device_table.chunk_sizes == oracle_table.chunk_sizes
    missing = "N" if alphabet_id == 0 else "X"
    for expression in (None, "(true)", "(has-mut c 7)", f"(sym-eq c 12 {missing})", f"(not (sym-eq c 12 {missing}))", "(sym-eq c 3 .)",
                       "(and (has-mut c 5) (not (has-mut c 9)))", "(ranges 3 90 196608 197000)"):
        if expression is not None:
            np.testing.assert_array_equal(device_table.filter(expression).ids(), oracle_table.filter(expression).ids())
        for min_proportion in (0.0, 0.05, 0.5):
            assert device_table.mutations(["c"], expression, min_proportion) == oracle_table.mutations("c", expression, min_proportion), expression
    np.testing.assert_array_equal(device_table.mutation_counts("c"), oracle_table.mutation_counts("c"))
    device_table.close()


@pytest.mark.gpu
def test_device_shards_loaded_from_archive_sum_to_the_whole(ctx):
    """three ranks' chunk ranges loaded from the same archive bytes on one device: filter cardinalities and counts are
    plain addends and sum to the oracle's whole-table result"""
    from lapis_silo_b200 import host_api as H
    oracle_table, data, _, specs = random_archive(301, 0)
    bounds = H.partition_chunks([1] * len(oracle_table.chunk_sizes), 3)
    shards = [H.HostTable.from_archive(ctx, data, specs, first_chunk=a, n_chunks=b - a) for a, b in zip(bounds, bounds[1:])]
    assert sum(s.num_rows for s in shards) == oracle_table.num_rows
    for expression in (None, "(has-mut c 7)", "(not (sym-eq c 12 N))"):
        want_filter = oracle_table.filter(expression) if expression else None
        total = np.zeros_like(oracle_table.mutation_counts("c"), dtype=np.uint64)
        ids = []
        for shard in shards:
            flt = shard.filter(expression) if expression else None
            total += shard.mutation_counts("c", flt).astype(np.uint64)
            if flt is not None:
                ids += [int(v) for v in flt.ids()]
        np.testing.assert_array_equal(total.astype(np.uint32), oracle_table.mutation_counts("c", want_filter))
        if expression:
            assert ids == [int(v) for v in want_filter.ids()]
    for shard in shards:
        shard.close()


@pytest.mark.gpu
def test_device_table_loaded_from_the_reference_state(ctx):
    """the reference's own containers through the loader (the file itself where the reference tree is mounted, else
    the archive rewritten from the committed extract), all four columns, E with its insertion index read through"""
    from lapis_silo_b200 import host_api as H
    from test_silo_state import oracle_table
    data = open(STATE_FILE, "rb").read() if os.path.exists(STATE_FILE) else W.write_archive(state_columns_for_writer())
    oracle = oracle_table()
    device_table = H.HostTable.from_archive(ctx, data, specs_of_state())
    assert sorted(device_table.columns) == ["E", "M", "main", "testSecondSequence"]
    for name in device_table.columns:
        for expression in (None, f"(has-mut {name} 2)", f"(not (has-mut {name} 2))", f"(sym-eq {name} 2 A)", f"(sym-eq {name} 4 .)"):
            if expression is not None:
                np.testing.assert_array_equal(device_table.filter(expression).ids(), oracle.filter(expression).ids())
            for min_proportion in (0.0, 0.05, 0.3):
                assert device_table.mutations([name], expression, min_proportion) == oracle.mutations(name, expression, min_proportion), (name, expression)
    device_table.close()


def test_reader_survives_damaged_input():
    """truncations, flipped bytes and absurd sizes anywhere in an archive / a roaring bitmap are either rejected with an
    error or parsed -- never a crash, and whatever the roaring decoder returns is a list of ascending, disjoint runs"""
    from lapis_silo_b200 import host_api as H
    from oracle import oracle as O
    good = W.write_archive(state_columns_for_writer(), b"\x05\0\0\0\0\0\0\0hello")
    specs = specs_of_state()
    rng = np.random.default_rng(0)
    outcomes = {"parsed": 0, "rejected": 0}
    for trial in range(3000):
        data = bytearray(good)
        kind = trial % 4
        if kind == 0:
            data = data[:int(rng.integers(0, len(data)))]
        elif kind == 1:
            for _ in range(int(rng.integers(1, 4))):
                data[int(rng.integers(30, len(data)))] = int(rng.integers(0, 256))
        elif kind == 2:
            at = int(rng.integers(30, len(data) - 8))
            data[at:at + 8] = int(rng.integers(0, 1 << 63)).to_bytes(8, "little")
        else:
            at = int(rng.integers(30, len(data) - 4))
            data[at:at + 4] = b"\xff\xff\xff\xff"
        try:
            H.Archive(bytes(data), specs).close()
            outcomes["parsed"] += 1
        except H.HostError:
            outcomes["rejected"] += 1
    assert outcomes["rejected"] > 1000 and outcomes["parsed"] > 0
    blobs = [O.roaring_roundtrip(sorted({int(v) for v in rng.integers(0, 1 << 18, 3000)} | set(range(70000, 80000))), optimize=o)[1]
             for o in (False, True)]
    for trial in range(3000):
        data = bytearray(blobs[trial % 2])
        if trial % 3 == 0:
            data = data[:int(rng.integers(0, len(data)))]
        elif trial % 3 == 1:
            for _ in range(int(rng.integers(1, 4))):
                data[int(rng.integers(0, min(len(data), 64)))] = int(rng.integers(0, 256))
        else:
            at = int(rng.integers(0, len(data) - 4))
            data[at:at + 4] = int(rng.integers(0, 1 << 32)).to_bytes(4, "little")
        try:
            runs = H.roaring_runs(bytes(data))
        except H.HostError:
            continue
        assert all(first < end for first, end in runs)
        assert all(a[1] < b[0] for a, b in zip(runs, runs[1:]))


def test_roaring_decoder_rejects_wrapping_keys():
    """a container with key 0xFFFF would make `key << 16 + 65536` wrap to 0 in 32 bits (runs with end < start that the
    ascending check then lets pass); keys must ascend strictly"""
    import struct
    from lapis_silo_b200 import host_api as H

    def no_runs(containers):  # SERIAL_COOKIE_NO_RUNCONTAINER, array containers only
        out = struct.pack("<II", 12346, len(containers))
        for key, values in containers:
            out += struct.pack("<HH", key, len(values) - 1)
        offset = len(out) + 4 * len(containers)
        for _, values in containers:
            out += struct.pack("<I", offset)
            offset += 2 * len(values)
        for _, values in containers:
            out += b"".join(struct.pack("<H", v) for v in values)
        return out

    assert H.roaring_runs(no_runs([(0, [1, 2, 3]), (2, [7])])) == [(1, 4), (2 * 65536 + 7, 2 * 65536 + 8)]
    for bad in ([(0xFFFF, [65535])], [(3, [1]), (3, [2])], [(5, [1]), (2, [9])]):
        with pytest.raises(H.HostError):
            H.roaring_runs(no_runs(bad))
    # run containers: SERIAL_COOKIE with the run flag of the only container set, one run covering the whole container
    full_last = struct.pack("<I", 12347 | (0 << 16)) + b"\x01" + struct.pack("<HH", 0xFFFF, 65535) + struct.pack("<HHH", 1, 0, 65535)
    with pytest.raises(H.HostError):
        H.roaring_runs(full_last)
