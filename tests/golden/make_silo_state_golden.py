"""Extracts the sequence columns of the reference's OWN serialised state
(/root/reference/testBaseData/siloSerializedState/1785915539/default.silo, written by the reference from
testBaseData/unitTestDummyDataset/input.ndjson; database.test.cpp:100-116 loads it) into
tests/golden/silo_state_unit_test_dummy.json with oracle/silo_archive.py, together with the five input rows.
Run here (the container that has /root/reference); the JSON is what the tests read everywhere."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import silo_archive  # noqa: E402

REFERENCE = "/root/reference/testBaseData"
STATE = f"{REFERENCE}/siloSerializedState/1785915539/default.silo"


def main():
    genomes = json.load(open(f"{REFERENCE}/unitTestDummyDataset/reference_genomes.json"))
    columns = [(g["name"], "Nucleotide", g["sequence"].encode()) for g in sorted(genomes["nucleotideSequences"], key=lambda g: g["name"])]
    columns += [(g["name"], "AminoAcid", g["sequence"].encode()) for g in sorted(genomes["genes"], key=lambda g: g["name"])]
    parsed = silo_archive.read_sequence_columns(STATE, columns)
    rows = [json.loads(line) for line in open(f"{REFERENCE}/unitTestDummyDataset/input.ndjson")]
    out = {
        "source": "testBaseData/siloSerializedState/1785915539/default.silo (reference-produced) + unitTestDummyDataset/input.ndjson",
        "columns": [],
        "rows": [{name: (row[name]["sequence"] if row[name] is not None else None) for name, _, _ in columns} for row in rows],
    }
    for name, alphabet, reference in columns:
        column = parsed[name]
        out["columns"].append({
            "name": name, "alphabet": alphabet, "reference": reference.decode(),
            "local_reference": column.local_reference.decode(),
            "containers": [{"position": key[0], "v_index": key[1], "symbol": key[2], "cardinality": cardinality,
                            "typecode": typecode, "payload_hex": payload.hex()}
                           for key, cardinality, typecode, payload in column.containers],
            "missing_bitmaps": {str(row): blob.hex() for row, blob in column.missing_bitmaps.items()},
            "start_end": column.start_end, "batch_start_ends": column.batch_start_ends,
            "sequence_count": column.sequence_count, "vertical_bitmaps_size": column.vertical_bitmaps_size,
            "horizontal_bitmaps_size": column.horizontal_bitmaps_size,
            "null_bitmap_hex": column.null_bitmap.hex() if column.null_bitmap is not None else None,
            "num_chunks": column.num_chunks,
            # insertion_index.h:28-100 (not on the query path; the loader has to read through it)
            "insertion_positions": [{"position": p["position"], "three_mer_buckets": p["three_mer_buckets"],
                                     "insertions": [[value, blob.hex()] for value, blob in p["insertions"]],
                                     "three_mers": [[symbols, ids] for symbols, ids in p["three_mers"]]}
                                    for p in column.insertion_positions],
            "insertion_bucket_counts": column.insertion_bucket_counts,
        })
    with open(os.path.join(HERE, "silo_state_unit_test_dummy.json"), "w") as handle:
        json.dump(out, handle, indent=1)
    print("wrote", len(out["columns"]), "columns,", len(out["rows"]), "rows")


if __name__ == "__main__":
    main()
