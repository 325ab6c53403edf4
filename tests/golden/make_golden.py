"""Generates the golden fixtures of tests/golden/ (run once, outputs committed).

The reference (Scala/JVM) cannot run in this image, so the vectors are produced by the two independent
implementations that pin the oracle: payloads are built with google.protobuf (upb) / hand-assembled wire bytes,
framed with the pure-Python bitwise CRC-32C of oracle/pyref.py, and the expected rows come from pyref's
restatement of M/TFRecordDeserializer.scala over upb message objects -- NOT from the C oracle or the CUDA path.
Files: <name>.tfrecord (framed bytes) + golden.json (schema, record type, expected rows or error)."""
import base64
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import numpy as np  # noqa: E402

import cases as CS  # noqa: E402
from spark_tfrecord_b200.sqltypes import ArrayType  # noqa: E402


def enc(v):
    if isinstance(v, list):
        return [enc(x) for x in v]
    if isinstance(v, bytes):
        return {"b64": base64.b64encode(v).decode()}
    if isinstance(v, str):
        return {"b64": base64.b64encode(v.encode()).decode()}
    if isinstance(v, (np.floating, float)):
        f = np.float32(v) if isinstance(v, np.float32) else np.float64(v)
        return {"f": float(f), "bits": int(f.view(np.uint32 if f.dtype == np.float32 else np.uint64))}
    if isinstance(v, np.integer):
        return int(v)
    return v


def type_str(dt):
    return f"array<{type_str(dt.elementType)}>" if isinstance(dt, ArrayType) else type(dt).__name__


def main():
    index = []
    picked = [c for c in CS.all_cases() if c.rows is not None or c.error is not None]
    for c in picked:
        fn = c.name + ".tfrecord"
        with open(os.path.join(HERE, fn), "wb") as f:
            f.write(c.data())
        index.append({"name": c.name, "file": fn, "record_type": c.record_type, "flags": c.flags, "is_final": c.is_final,
                      "schema": [{"name": fl.name, "type": type_str(fl.dataType), "nullable": fl.nullable} for fl in c.schema],
                      "rows": enc(c.rows) if c.rows is not None else None, "error": c.error, "error_row": c.error_row,
                      "error_field": c.error_field, "rows_before_error": enc(c.rows_before_error) if c.rows_before_error is not None else None})
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(index, f, indent=0)
    print(len(index), "fixtures,", sum(os.path.getsize(os.path.join(HERE, i["file"])) for i in index), "bytes")


if __name__ == "__main__":
    main()
