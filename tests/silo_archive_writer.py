"""TEST INFRASTRUCTURE: writes the sequence-column section of a `.silo` table file (boost binary archive, library
version 20) in the layout oracle/silo_archive.py documents, so that the product's C++ reader
(lapis_silo_b200/host/silo_loader.cpp) can be driven with tables of any shape where /root/reference is not
mounted. The writer is itself pinned on reference-produced bytes: tests/test_silo_loader.py checks that, fed the
columns of the reference's own serialised state, it reproduces that file's bytes for every column it covers.

Restates (relative to /root/reference/src/rhydb/): storage/column/sequence_column.h:86-96,
vertical_sequence_index.h:31-38,110-112, roaring_util/roaring_container.h:104-157,
horizontal_coverage_index.h:109-113, roaring_util/roaring_serialize.h:15-46, insertion_index.h:84-100."""
from __future__ import annotations

import struct

HEADER = struct.pack("<Q", 22) + b"serialization::archive" + struct.pack("<H", 20) + bytes([4, 8, 4, 8]) + struct.pack("<I", 1)
EMPTY_ROARING = struct.pack("<II", 12346, 0)


class _Out:
    def __init__(self, seen: set):
        self.parts: list[bytes] = []
        self.seen = seen

    def raw(self, data: bytes): self.parts.append(data)
    def u8(self, v): self.raw(struct.pack("<B", v))
    def u16(self, v): self.raw(struct.pack("<H", v))
    def u32(self, v): self.raw(struct.pack("<I", v))
    def u64(self, v): self.raw(struct.pack("<Q", v))

    def string(self, data: bytes):
        self.u64(len(data))
        self.raw(data)

    def class_info(self, name: str):
        if name not in self.seen:
            self.seen.add(name)
            self.raw(b"\0\0\0\0\0")

    def roaring(self, data: bytes):
        self.class_info("roaring::Roaring")
        self.string(data)

    def pair_vector(self, pairs):
        self.class_info("vector<pair<u32,u32>>")
        self.u64(len(pairs))
        for first, second in pairs:
            self.raw(struct.pack("<II", first, second))


def sequence_column_bytes(column: dict, seen: set) -> bytes:
    """column: alphabet ("Nucleotide" | "AminoAcid"), local_reference (str), containers [{position, v_index, symbol,
    cardinality, typecode, payload (bytes)}] in map order, missing_bitmaps {row: portable roaring bytes},
    start_end [[(start, end)] per chunk], batch_start_ends [(start, end)], sequence_count, vertical_bitmaps_size,
    horizontal_bitmaps_size, null_bitmap (portable roaring bytes), num_chunks; optional insertion_positions
    [{position, insertions [(value, portable roaring bytes)], three_mers [([3 symbol ids], [insertion ids])],
    three_mer_buckets}] (the reader has to read through them) and insertion_bucket_counts (the two hash tables'
    bucket counts)."""
    alphabet = column["alphabet"]
    out = _Out(seen)
    out.class_info(f"SequenceColumn<{alphabet}>")
    out.string(column["local_reference"].encode())
    out.class_info(f"VerticalSequenceIndex<{alphabet}>")
    out.class_info(f"map<SequenceDiffKey<{alphabet}>,RoaringContainer>")
    out.u64(len(column["containers"]))
    out.u32(0)
    for container in column["containers"]:
        out.class_info(f"pair<SequenceDiffKey<{alphabet}>,RoaringContainer>")
        out.class_info(f"SequenceDiffKey<{alphabet}>")
        out.u32(container["position"])
        out.u16(container["v_index"])
        out.u32(container["symbol"])
        out.class_info("RoaringContainer")
        out.u32(container["cardinality"])
        out.u8(container["typecode"])
        out.string(container["payload"])
    out.class_info("HorizontalCoverageIndex")
    out.class_info("map<u32,Roaring>")
    out.u64(len(column["missing_bitmaps"]))
    out.u32(0)
    for row in sorted(column["missing_bitmaps"]):
        out.class_info("pair<u32,Roaring>")
        out.u32(row)
        out.roaring(column["missing_bitmaps"][row])
    out.class_info("vector<vector<pair<u32,u32>>>")
    out.u64(len(column["start_end"]))
    out.u32(0)
    for chunk in column["start_end"]:
        out.pair_vector(chunk)
    out.pair_vector(column["batch_start_ends"])
    out.class_info(f"InsertionIndex<{alphabet}>")
    out.class_info(f"unordered_map<u32,InsertionPosition<{alphabet}>>")
    # boost saves bucket_count even for an empty table; the value is whatever the hash table held at save time
    # (the reference's file has 2 / 13 for insertion_positions and 1 / 13 for collected_insertions), the reader ignores it
    buckets = column.get("insertion_bucket_counts", (1, 1))
    positions = column.get("insertion_positions", [])
    out.u64(len(positions))
    out.u64(buckets[0])
    out.u32(0)
    for entry in positions:  # insertion_index.h:28-63
        out.class_info(f"pair<u32,InsertionPosition<{alphabet}>>")
        out.u32(entry["position"])
        out.class_info(f"InsertionPosition<{alphabet}>")
        out.class_info("vector<Insertion>")
        out.u64(len(entry["insertions"]))
        out.u32(0)
        for value, row_ids in entry["insertions"]:
            out.class_info("Insertion")
            out.string(value.encode())
            out.roaring(row_ids)
        out.class_info(f"unordered_map<ThreeMer<{alphabet}>,InsertionIds>")
        out.u64(len(entry["three_mers"]))
        out.u64(entry.get("three_mer_buckets", 13))
        out.u32(0)
        for symbols, ids in entry["three_mers"]:
            out.class_info(f"pair<ThreeMer<{alphabet}>,InsertionIds>")
            out.class_info(f"ThreeMer<{alphabet}>")
            out.u64(3)
            for symbol in symbols:
                out.u32(symbol)
            out.u64(len(ids))
            for insertion_id in ids:
                out.u32(insertion_id)
    out.class_info("unordered_map<u32,unordered_map<string,Roaring>>")
    out.u64(0)
    out.u64(buckets[1])
    out.u32(0)
    out.class_info("SequenceColumnInfo")
    out.u32(column["sequence_count"])
    out.u64(column["vertical_bitmaps_size"])
    out.u64(column["horizontal_bitmaps_size"])
    out.u32(column["sequence_count"])
    out.roaring(column["null_bitmap"])
    out.u16(column["num_chunks"])
    return b"".join(out.parts)


def write_archive(columns: list[dict], metadata: bytes = b"", roaring_seen: bool = True, pair_seen: bool = True) -> bytes:
    """header + `metadata` (stand-in for the metadata columns in front, not parsed by the reader) + the sequence columns.
    roaring_seen / pair_seen: whether the metadata columns already registered roaring::Roaring / pair<u32,Roaring>."""
    seen: set = set()
    if roaring_seen:
        seen.add("roaring::Roaring")
    if pair_seen:
        seen.add("pair<u32,Roaring>")
    return HEADER + metadata + b"".join(sequence_column_bytes(column, seen) for column in columns)
