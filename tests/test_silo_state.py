"""The storage stage pinned on REFERENCE-PRODUCED bytes: tests/golden/silo_state_unit_test_dummy.json holds the
sequence columns of testBaseData/siloSerializedState/1785915539/default.silo -- the state the reference itself
serialised from unitTestDummyDataset/input.ndjson and reloads in database.test.cpp:100-116 -- extracted by
oracle/silo_archive.py (tests/golden/make_silo_state_golden.py), plus the five input rows.

  * the oracle, fed the same five rows, must build exactly the reference's containers (key, typecode,
    cardinality, payload bytes), coverage ranges, missing-symbol bitmaps and local references;
  * (where /root/reference is mounted) the archive reader must still reproduce the committed JSON;
  * GPU: the reference's containers, uploaded as they are through S1, must give the oracle's Mutations rows."""
import ctypes as C
import json
import os
import struct

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "silo_state_unit_test_dummy.json")) as handle:
    STATE = json.load(handle)
STATE_FILE = "/root/reference/testBaseData/siloSerializedState/1785915539/default.silo"


def roaring_values(blob: bytes) -> list[int]:
    """portable roaring format without run containers (cookie 12346), enough for the fixture's bitmaps"""
    cookie, n = struct.unpack_from("<II", blob, 0)
    assert cookie == 12346
    keys = [struct.unpack_from("<HH", blob, 8 + 4 * i) for i in range(n)]
    cursor = 8 + 4 * n + 4 * n  # descriptive header + offset header
    out = []
    for key, cardinality_minus_one in keys:
        count = cardinality_minus_one + 1
        assert count <= 4096
        out += [(key << 16) | v for v in struct.unpack_from(f"<{count}H", blob, cursor)]
        cursor += 2 * count
    return out


def oracle_table():
    from oracle import oracle as O
    table = O.Table()
    for column in STATE["columns"]:
        table.add_column(column["name"], O.NUCLEOTIDE if column["alphabet"] == "Nucleotide" else O.AMINO_ACID, column["reference"])
    for row in STATE["rows"]:
        table.append_row([row[column["name"]] for column in STATE["columns"]])
    table.finalize()
    return table


def exported(table, name):
    """the oracle's column in the S1 upload format, as plain Python values"""
    export = table.export_column(name)
    desc = export.desc.contents
    containers = []
    for i in range(desc.n_containers):
        c = desc.containers[i]
        payload = bytes(desc.payload[c.payload_offset + k] for k in range(c.payload_bytes))
        containers.append({"position": c.position, "v_index": c.v_index, "symbol": c.symbol, "cardinality": c.cardinality,
                           "typecode": c.typecode, "payload_hex": payload.hex()})
    n_rows = table.num_rows
    start_end = [(desc.start_end[2 * r], desc.start_end[2 * r + 1]) for r in range(n_rows)]
    missing = {}
    for i in range(desc.n_rows_with_missing):
        positions = []
        for run in range(desc.missing_offsets[i], desc.missing_offsets[i + 1]):
            positions += list(range(desc.missing_runs[2 * run], desc.missing_runs[2 * run + 1]))
        missing[int(desc.missing_row_ids[i])] = positions
    local_reference = bytes(desc.local_reference[p] for p in range(desc.genome_length))
    export.close()
    return containers, start_end, missing, local_reference


def test_oracle_builds_the_reference_serialised_state():
    from oracle import oracle as O
    table = oracle_table()
    assert table.num_rows == 5
    for column in STATE["columns"]:
        chars = O.NUC_SYMBOLS if column["alphabet"] == "Nucleotide" else O.AA_SYMBOLS
        containers, start_end, missing, local_reference = exported(table, column["name"])
        # vertical_sequence_index.h:22-44 in map order (position, v_index, symbol); container_write bytes
        assert containers == column["containers"], column["name"]
        # horizontal_coverage_index.h:25-35
        assert [start_end] == [[tuple(pair) for pair in chunk] for chunk in column["start_end"]], column["name"]
        assert missing == {int(row): roaring_values(bytes.fromhex(blob)) for row, blob in column["missing_bitmaps"].items()}, column["name"]
        assert "".join(chars[s] for s in local_reference) == column["local_reference"]
        assert table.local_reference(column["name"]) == column["local_reference"]
    # database.test.cpp:100-116
    main = next(c for c in STATE["columns"] if c["name"] == "main")
    assert main["sequence_count"] == 5 and main["horizontal_bitmaps_size"] == 9


@pytest.mark.skipif(not os.path.exists(STATE_FILE), reason="the reference tree is not mounted here")
def test_archive_reader_reproduces_the_committed_extract():
    from oracle import silo_archive
    columns = [(c["name"], c["alphabet"], c["reference"].encode()) for c in STATE["columns"]]
    parsed = silo_archive.read_sequence_columns(STATE_FILE, columns)
    for column in STATE["columns"]:
        got = parsed[column["name"]]
        assert got.local_reference.decode() == column["local_reference"]
        assert [{"position": k[0], "v_index": k[1], "symbol": k[2], "cardinality": n, "typecode": t, "payload_hex": p.hex()}
                for k, n, t, p in got.containers] == column["containers"]
        assert [[list(pair) for pair in chunk] for chunk in got.start_end] == column["start_end"]
        assert {str(row): blob.hex() for row, blob in got.missing_bitmaps.items()} == column["missing_bitmaps"]
        assert got.sequence_count == column["sequence_count"]
        assert got.horizontal_bitmaps_size == column["horizontal_bitmaps_size"]


def test_archive_reader_rejects_other_files(tmp_path):
    from oracle import silo_archive
    bad = tmp_path / "not_an_archive.silo"
    bad.write_bytes(b"\x16\0\0\0\0\0\0\0serialization::archive\x13\0\x04\x08\x04\x08\x01\0\0\0")
    with pytest.raises(ValueError):
        silo_archive.read_sequence_columns(str(bad), [("main", "Nucleotide", b"ACGT")])


@pytest.mark.gpu
def test_device_path_on_the_reference_containers():
    """S1 fed the reference's own container bytes (not the oracle's export), then filters and the Mutations
    action against the oracle built from the input rows."""
    from lapis_silo_b200 import abi, host_api
    from oracle import oracle as O
    oracle = oracle_table()
    ctx = abi.Context(0)
    device = host_api.HostTable(ctx, [5])
    keep = []
    for column in STATE["columns"]:
        nucleotide = column["alphabet"] == "Nucleotide"
        chars = O.NUC_SYMBOLS if nucleotide else O.AA_SYMBOLS
        length = len(column["reference"])
        payload = b"".join(bytes.fromhex(c["payload_hex"]) for c in column["containers"])
        descs = (abi.ContainerDesc * max(len(column["containers"]), 1))()
        offset = 0
        for i, c in enumerate(column["containers"]):
            size = len(c["payload_hex"]) // 2
            descs[i] = abi.ContainerDesc(c["position"], c["v_index"], c["symbol"], c["typecode"], c["cardinality"], size, offset)
            offset += size
        payload_buffer = (C.c_uint8 * max(len(payload), 1)).from_buffer_copy(payload or b"\0")
        local_reference = (C.c_uint8 * length)(*[chars.index(ch) for ch in column["local_reference"]])
        flat = [v for pair in column["start_end"][0] for v in pair]
        start_end = (C.c_uint32 * len(flat))(*flat)
        rows = sorted(int(r) for r in column["missing_bitmaps"])
        runs, offsets = [], [0]
        for row in rows:
            for position in roaring_values(bytes.fromhex(column["missing_bitmaps"][str(row)])):
                runs += [position, position + 1]
            offsets.append(len(runs) // 2)
        row_ids = (C.c_uint32 * max(len(rows), 1))(*rows)
        missing_offsets = (C.c_uint64 * len(offsets))(*offsets)
        missing_runs = (C.c_uint32 * max(len(runs), 1))(*runs)
        desc = abi.ColumnDesc()
        desc.struct_size = C.sizeof(abi.ColumnDesc)
        desc.n_symbols = len(chars)
        desc.genome_length = length
        desc.missing_symbol = len(chars) - 1  # N / X: the last symbol of both alphabets
        desc.local_reference = C.cast(local_reference, C.POINTER(C.c_uint8))
        desc.n_containers = len(column["containers"])
        desc.containers = C.cast(descs, C.POINTER(abi.ContainerDesc))
        desc.payload = C.cast(payload_buffer, C.POINTER(C.c_uint8))
        desc.payload_bytes = len(payload)
        desc.start_end = C.cast(start_end, C.POINTER(C.c_uint32))
        desc.n_rows_with_missing = len(rows)
        desc.missing_row_ids = C.cast(row_ids, C.POINTER(C.c_uint32))
        desc.missing_offsets = C.cast(missing_offsets, C.POINTER(C.c_uint64))
        desc.missing_runs = C.cast(missing_runs, C.POINTER(C.c_uint32))
        desc.n_null_rows = 0
        keep.append((descs, payload_buffer, local_reference, start_end, row_ids, missing_offsets, missing_runs, desc))
        device.add_column(column["name"], host_api.NUCLEOTIDE if nucleotide else host_api.AMINO_ACID, column["reference"], C.pointer(desc))
    for column in STATE["columns"]:
        name = column["name"]
        for expression in (None, "(true)", f"(has-mut {name} 2)", f"(not (has-mut {name} 2))", f"(sym-eq {name} 2 A)", f"(sym-eq {name} 4 .)"):
            if expression is not None:
                np.testing.assert_array_equal(device.filter(expression).ids(), oracle.filter(expression).ids())
            for min_proportion in (0.0, 0.05, 0.3):
                assert device.mutations([name], expression, min_proportion) == oracle.mutations(name, expression, min_proportion), (name, expression)
            np.testing.assert_array_equal(device.mutation_counts(name), oracle.mutation_counts(name))
    device.close()
    ctx.close()
