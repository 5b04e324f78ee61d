"""Reader for the sequence-column section of a RhyDB/SILO `.silo` table file (a boost::archive::binary_oarchive,
library version 20) -- TEST INFRASTRUCTURE: it pins the oracle's storage stage on bytes the REFERENCE ITSELF
produced (testBaseData/siloSerializedState/<ts>/default.silo is the serialised state of
testBaseData/unitTestDummyDataset/input.ndjson, database.test.cpp:100-116). Nothing in the product imports it.

What is restated (all relative to /root/reference/src/rhydb/):
  storage/table.h:35-42            Table::serializeData: columns, sequence_count, row_layout
  storage/column_group.h:28-61     the column maps in a fixed order; sequence columns come after every
                                   metadata column and before the zstd-compressed (unaligned) string columns
  storage/column/sequence_column.h:86-96   local_reference_sequence_string, vertical_sequence_index,
                                   horizontal_coverage_index, insertion_index, sequence_column_info,
                                   sequence_count, null_bitmap, num_chunks
  storage/column/vertical_sequence_index.h:31-38,110-112   map<{u32 position, u16 v_index, Symbol}, RoaringContainer>
  roaring_util/roaring_container.h:104-157 cardinality (u32), typecode (u8), container_write bytes as a string
  storage/column/horizontal_coverage_index.h:109-113       map<u32 row, Roaring>, start_end, batch_start_ends
  roaring_util/roaring_serialize.h:15-46   size_t size + the portable roaring bytes

boost binary archive facts the reader relies on (observed on the fixture, consistent with boost 1.85):
  * header: u64 length + "serialization::archive", u16 library version (20), 4 type-size bytes, u32 0x00000001
  * the FIRST time an object of a class type is saved the archive holds 1 byte tracking + 4 bytes class
    version in front of it (5 zero bytes here); later objects of the same type hold nothing
  * std::string / collection sizes are u64; std::map = [size u64, item_version u32, items];
    std::unordered_map = [size u64, bucket_count u64, item_version u32, items];
    vector<pair<u32,u32>> is stored as size + raw bytes (bitwise serialisable)
  * enums are saved as 32-bit ints (the Symbol of a SequenceDiffKey takes 4 bytes)
  * std::array<Symbol, 3> = [element count u64, the enums]; std::vector<uint32_t> = [size u64, raw values]
The metadata columns in front of the sequence columns are not parsed: the first sequence column is located by
its length-prefixed local reference, and the class types those columns already registered are listed in
PRE_SEEN. The insertion index (insertion_index.h:28-100) is read through -- the query path does not use it, but
the column's sequence_count, null bitmap and num_chunks lie behind it."""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

PRE_SEEN = {"roaring::Roaring", "pair<u32,Roaring>"}  # registered by the metadata columns (null bitmaps, lineage index)


class _Cursor:
    def __init__(self, data: bytes, position: int, seen: set):
        self.data, self.position, self.seen = data, position, seen

    def take(self, n: int) -> bytes:
        if self.position + n > len(self.data):
            raise ValueError("archive truncated")
        out = self.data[self.position:self.position + n]
        self.position += n
        return out

    def u8(self): return self.take(1)[0]
    def u16(self): return struct.unpack("<H", self.take(2))[0]
    def u32(self): return struct.unpack("<I", self.take(4))[0]
    def u64(self): return struct.unpack("<Q", self.take(8))[0]
    def string(self): return self.take(self.u64())

    def class_info(self, name: str) -> None:
        """tracking byte + class version in front of the first object of a class type"""
        if name not in self.seen:
            self.seen.add(name)
            preamble = self.take(5)
            if preamble != b"\0\0\0\0\0":
                raise ValueError(f"unexpected class info {preamble.hex()} for {name} at {self.position - 5}")

    def roaring(self) -> bytes:
        self.class_info("roaring::Roaring")
        return self.take(self.u64())


@dataclass
class SequenceColumn:
    local_reference: bytes
    containers: list = field(default_factory=list)   # ((position, v_index, symbol id), cardinality, typecode, payload bytes)
    missing_bitmaps: dict = field(default_factory=dict)  # row -> portable roaring bytes (positions with the missing symbol)
    start_end: list = field(default_factory=list)     # per chunk: [(start, end)] per row
    batch_start_ends: list = field(default_factory=list)
    sequence_count: int | None = None
    vertical_bitmaps_size: int | None = None
    horizontal_bitmaps_size: int | None = None
    null_bitmap: bytes | None = None
    num_chunks: int | None = None
    insertion_positions: list = field(default_factory=list)  # [{position, insertions [(value, roaring bytes)], three_mers [([3 symbol ids], [insertion ids])], three_mer_buckets}]
    insertion_bucket_counts: list = field(default_factory=list)  # bucket counts of the two hash tables (a runtime artefact boost saves)


def parse_header(data: bytes) -> int:
    cursor = _Cursor(data, 0, set())
    if cursor.string() != b"serialization::archive":
        raise ValueError("not a boost binary archive")
    version = cursor.u16()
    sizes = cursor.take(4)
    endian = cursor.u32()
    if version != 20 or sizes != bytes([4, 8, 4, 8]) or endian != 1:
        raise ValueError(f"unsupported archive flavour: version {version}, sizes {sizes.hex()}, endian {endian}")
    return cursor.position


def _pair_vector(cursor: _Cursor) -> list:
    cursor.class_info("vector<pair<u32,u32>>")
    return [(cursor.u32(), cursor.u32()) for _ in range(cursor.u64())]


def _sequence_column(cursor: _Cursor, alphabet: str) -> SequenceColumn:
    cursor.class_info(f"SequenceColumn<{alphabet}>")
    column = SequenceColumn(local_reference=cursor.string())
    # vertical_sequence_index.h:110-112
    cursor.class_info(f"VerticalSequenceIndex<{alphabet}>")
    cursor.class_info(f"map<SequenceDiffKey<{alphabet}>,RoaringContainer>")
    n, _item_version = cursor.u64(), cursor.u32()
    for _ in range(n):
        cursor.class_info(f"pair<SequenceDiffKey<{alphabet}>,RoaringContainer>")
        cursor.class_info(f"SequenceDiffKey<{alphabet}>")
        key = (cursor.u32(), cursor.u16(), cursor.u32())  # position, v_index, symbol (an enum: saved as int)
        cursor.class_info("RoaringContainer")
        cardinality, typecode = cursor.u32(), cursor.u8()
        column.containers.append((key, cardinality, typecode, cursor.string()))
    # horizontal_coverage_index.h:109-113
    cursor.class_info("HorizontalCoverageIndex")
    cursor.class_info("map<u32,Roaring>")
    n, _item_version = cursor.u64(), cursor.u32()
    for _ in range(n):
        cursor.class_info("pair<u32,Roaring>")
        row = cursor.u32()
        column.missing_bitmaps[row] = cursor.roaring()
    cursor.class_info("vector<vector<pair<u32,u32>>>")
    n, _item_version = cursor.u64(), cursor.u32()
    column.start_end = [_pair_vector(cursor) for _ in range(n)]
    column.batch_start_ends = _pair_vector(cursor)
    # insertion_index.h:84-100: insertion_positions (position -> InsertionPosition {insertions, three_mer_index},
    # :28-63) and collected_insertions; the query path does not need them, they are read to get past them
    cursor.class_info(f"InsertionIndex<{alphabet}>")
    cursor.class_info(f"unordered_map<u32,InsertionPosition<{alphabet}>>")
    n_positions, buckets, _item_version = cursor.u64(), cursor.u64(), cursor.u32()
    column.insertion_bucket_counts = [buckets]
    for _ in range(n_positions):
        cursor.class_info(f"pair<u32,InsertionPosition<{alphabet}>>")
        position = cursor.u32()
        cursor.class_info(f"InsertionPosition<{alphabet}>")
        cursor.class_info("vector<Insertion>")
        n_insertions, _item_version = cursor.u64(), cursor.u32()
        insertions = []
        for _ in range(n_insertions):
            cursor.class_info("Insertion")
            value = cursor.string()
            insertions.append((value.decode(), cursor.roaring()))
        cursor.class_info(f"unordered_map<ThreeMer<{alphabet}>,InsertionIds>")
        n_three_mers, three_mer_buckets, _item_version = cursor.u64(), cursor.u64(), cursor.u32()
        three_mers = []
        for _ in range(n_three_mers):
            cursor.class_info(f"pair<ThreeMer<{alphabet}>,InsertionIds>")
            cursor.class_info(f"ThreeMer<{alphabet}>")  # std::array<Symbol, 3>: element count + the enums as ints
            if cursor.u64() != 3:
                raise ValueError("a three-mer that does not have three symbols")
            symbols = [cursor.u32() for _ in range(3)]
            ids = [cursor.u32() for _ in range(cursor.u64())]  # std::vector<uint32_t>: size + raw values
            three_mers.append((symbols, ids))
        column.insertion_positions.append({"position": position, "insertions": insertions, "three_mers": three_mers,
                                           "three_mer_buckets": three_mer_buckets})
    cursor.class_info("unordered_map<u32,unordered_map<string,Roaring>>")
    n_collected, buckets, _item_version = cursor.u64(), cursor.u64(), cursor.u32()
    column.insertion_bucket_counts.append(buckets)
    for _ in range(n_collected):  # (empty after buildIndex; layout by the same rules, not seen in a fixture)
        cursor.class_info("pair<u32,unordered_map<string,Roaring>>")
        cursor.u32()
        cursor.class_info("unordered_map<string,Roaring>")
        n_values, _buckets, _item_version = cursor.u64(), cursor.u64(), cursor.u32()
        for _ in range(n_values):
            cursor.class_info("pair<string,Roaring>")
            cursor.string()
            cursor.roaring()
    # sequence_column.h:35-39,92-95
    cursor.class_info("SequenceColumnInfo")
    column.sequence_count, column.vertical_bitmaps_size, column.horizontal_bitmaps_size = cursor.u32(), cursor.u64(), cursor.u64()
    if cursor.u32() != column.sequence_count:
        raise ValueError("sequence_count mismatch")
    column.null_bitmap = cursor.roaring()
    column.num_chunks = cursor.u16()
    return column


def read_sequence_columns(path: str, columns: list[tuple[str, str, bytes]]) -> dict[str, SequenceColumn]:
    """columns: (name, "Nucleotide" | "AminoAcid", reference sequence) in the archive's order: nucleotide columns by
    name, then amino-acid columns by name (std::map order, column_group.h:49-57). Returns name -> SequenceColumn."""
    data = open(path, "rb").read()
    parse_header(data)
    # locate every column by its length-prefixed local reference (identical to the reference here: no
    # position of the fixture has a majority symbol that differs from it), in order
    starts = []
    search_from = 0
    for name, _alphabet, reference in columns:
        needle = struct.pack("<Q", len(reference)) + reference
        at = data.find(needle, search_from)
        if at < 0:
            raise ValueError(f"local reference of column {name} not found")
        starts.append(at)
        search_from = at + len(needle)
    seen = set(PRE_SEEN)
    out = {}
    cursor = _Cursor(data, 0, seen)
    for index, (name, alphabet, _reference) in enumerate(columns):
        first_of_alphabet = f"SequenceColumn<{alphabet}>" not in seen
        # the class info of the first column of an alphabet sits in front of the located string
        cursor.position = starts[index] - (5 if first_of_alphabet else 0)
        out[name] = _sequence_column(cursor, alphabet)
        if index + 1 < len(columns):
            first_of_next = f"SequenceColumn<{columns[index + 1][1]}>" not in seen
            if cursor.position != starts[index + 1] - (5 if first_of_next else 0):
                raise ValueError(f"column {name} ends at {cursor.position}, the next one starts at {starts[index + 1]}")
    return out
