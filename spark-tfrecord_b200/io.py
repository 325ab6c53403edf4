"""Host-side mirror of the reference interface for the hot path, by name.

    reference (Scala)                                   here
    ------------------------------------------------    ---------------------------------------------
    TFRecordFileReader.readFile(conf, options, file,    TFRecordFileReader.readFile(conf, options, file, schema)
        schema): Iterator[InternalRow]                      -> iterator of row tuples
        (M/TFRecordFileReader.scala:16-83)
    new TFRecordDeserializer(schema).deserializeExample TFRecordDeserializer(schema).deserializeExample(bytes)
        (M/TFRecordDeserializer.scala:21-61)
    new TFRecordSerializer(schema).serializeExample     TFRecordSerializer(schema).serializeExample(row) -> bytes
        (M/TFRecordSerializer.scala:20-60)
    new TFRecordOutputWriter(path, options, schema,     TFRecordOutputWriter(path, options, dataSchema, context)
        context).write(row)/close()                         .write(row) / .close()
        (M/TFRecordOutputWriter.scala:12-44)
    DefaultSource (M/DefaultSource.scala:23-143)        DefaultSource: shortName/isSplitable/buildReader/prepareWrite

Everything that touches record bytes goes through the C ABI (libtfrgpu.so): there is no Python or CPU
implementation of the path in this package.  Rows are tuples of Python values (None, int, float, str,
bytes, list, list of lists) standing in for Catalyst's InternalRow; message arguments are serialized
protobuf bytes (there are no org.tensorflow.example classes here)."""
from __future__ import annotations

import os
import struct
from typing import Dict, Iterator, List, Optional, Sequence

import numpy as np

from . import _native
from ._cabi import TFR_F_DEFAULT, TFR_E_BATCH_TOO_LARGE as A_TFR_E_BATCH_TOO_LARGE, columns_from_rows
from .sqltypes import RECORD_TYPES, StructType, byte_array_schema

M = "src/main/scala/com/linkedin/spark/datasources/tfrecord/"


def _record_type(options: Optional[Dict[str, str]]) -> int:
    rt = (options or {}).get("recordType", "Example")          # M/TFRecordFileReader.scala:22
    if rt not in RECORD_TYPES:                                  # :78-79
        raise _native.IllegalArgumentException(-3, f"Unsupported recordType {rt}: recordType can be ByteArray, Example or SequenceExample")
    return RECORD_TYPES[rt]


# ---- stream compression: host-side, around the same GPU kernels -------------------------------------------
# The reference hands `codec` to Hadoop (M/DefaultSource.scala:94-102: a codec class name) and CodecStreams compresses the
# whole output stream; on read, Hadoop picks the codec from the file extension.  The compressed bytes are a container
# around the framed records, so they are (de)compressed on the host and the framed bytes go through the C ABI unchanged.
_CODECS = {   # name -> (file extension, Hadoop class)
    "gzip": (".gz", "org.apache.hadoop.io.compress.GzipCodec"),
    "deflate": (".deflate", "org.apache.hadoop.io.compress.DefaultCodec"),
    "bzip2": (".bz2", "org.apache.hadoop.io.compress.BZip2Codec"),
}


def _codec_name(codec: str) -> Optional[str]:
    """option value (Hadoop class name, or its short name) -> one of _CODECS; '' -> None"""
    if not codec:
        return None
    for name, (_, cls) in _CODECS.items():
        if codec == cls or codec.lower() == name or codec.lower() == cls.rsplit(".", 1)[1].lower():
            return name
    raise _native.IllegalArgumentException(-3, f"codec {codec}: only GzipCodec, DefaultCodec (deflate) and BZip2Codec are available on this host")


def _codec_of_path(path: str) -> Optional[str]:
    for name, (ext, _) in _CODECS.items():
        if path.endswith(ext):
            return name
    return None


class _DeflateReader:
    """zlib-format stream (Hadoop DefaultCodec, '.deflate') as a file-like object with read(n)"""

    def __init__(self, f):
        import zlib
        self._f, self._z, self._buf, self._eof = f, zlib.decompressobj(), b"", False

    def read(self, n: int = -1) -> bytes:
        while not self._eof and (n < 0 or len(self._buf) < n):
            raw = self._f.read(1 << 20)
            if not raw:
                self._buf += self._z.flush()
                self._eof = True
                break
            self._buf += self._z.decompress(raw)
        if n < 0:
            out, self._buf = self._buf, b""
        else:
            out, self._buf = self._buf[:n], self._buf[n:]
        return out

    def close(self):
        self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class _DeflateWriter:
    def __init__(self, f):
        import zlib
        self._f, self._z = f, zlib.compressobj()

    def write(self, b: bytes):
        self._f.write(self._z.compress(b))

    def close(self):
        self._f.write(self._z.flush())
        self._f.close()


def _open_read(path: str):
    """file object yielding the FRAMED bytes of `path` (decompressed when the extension names a codec)"""
    codec = _codec_of_path(path)
    if codec == "gzip":
        import gzip
        return gzip.open(path, "rb")
    if codec == "bzip2":
        import bz2
        return bz2.open(path, "rb")
    if codec == "deflate":
        return _DeflateReader(open(path, "rb"))
    return open(path, "rb")


def _open_write(path: str, codec: Optional[str]):
    if codec == "gzip":
        import gzip
        return gzip.open(path, "wb")
    if codec == "bzip2":
        import bz2
        return bz2.open(path, "wb")
    if codec == "deflate":
        return _DeflateWriter(open(path, "wb"))
    return open(path, "wb")


def _rows_of(batch: "_native.Batch") -> List[tuple]:
    cols = batch.to_host()
    return [tuple(c.get(r) for c in cols) for r in range(batch.n_rows)]


class TFRecordDeserializer:
    """One record at a time, like the reference class; each call is one tfr_decode of a single frame (CRC
    check off: the payload never went through TFRecordReader here)."""

    def __init__(self, dataSchema: StructType, device: int = 0):
        self.schema = dataSchema
        self.device = device
        self._dec: Dict[int, _native.Decoder] = {}

    def _decoder(self, rt: int) -> "_native.Decoder":
        if rt not in self._dec:
            self._dec[rt] = _native.Decoder(self.schema, rt, self.device, flags=0)
        return self._dec[rt]

    def _one(self, payload: bytes, rt: int) -> tuple:
        frame = struct.pack("<QI", len(payload), 0) + bytes(payload) + b"\0\0\0\0"
        batch, _ = self._decoder(rt).decode(frame)
        try:
            batch.raise_if_error()
            return _rows_of(batch)[0]
        finally:
            batch.release()

    def deserializeByteArray(self, byteArray: bytes) -> tuple:
        return self._one(byteArray, 2)

    def deserializeExample(self, example: bytes) -> tuple:
        return self._one(example, 0)

    def deserializeSequenceExample(self, sequenceExample: bytes) -> tuple:
        return self._one(sequenceExample, 1)

    def close(self):
        for d in self._dec.values():
            d.close()
        self._dec = {}


class TFRecordSerializer:
    """Constructor validates the types like the reference's (featureConverters are built eagerly,
    M/TFRecordSerializer.scala:14 -> RuntimeException for unsupported types)."""

    def __init__(self, dataSchema: StructType, device: int = 0):
        self.schema = dataSchema
        self.device = device
        self._enc: Dict[int, _native.Encoder] = {}
        _native.Schema(dataSchema, 1).close()        # type validation only

    def _encoder(self, rt: int) -> "_native.Encoder":
        if rt not in self._enc:
            schema = byte_array_schema() if rt == 2 else self.schema
            self._enc[rt] = _native.Encoder(schema, rt, self.device)
        return self._enc[rt]

    def _one(self, row: Sequence, rt: int) -> bytes:
        schema = byte_array_schema() if rt == 2 else self.schema
        framed = self._encoder(rt).encode(columns_from_rows(schema, [tuple(row)], rt))
        return framed[12:-4]

    def serializeByteArray(self, row: Sequence) -> bytes:
        return self._one(row, 2)

    def serializeExample(self, row: Sequence) -> bytes:
        return self._one(row, 0)

    def serializeSequenceExample(self, row: Sequence) -> bytes:
        return self._one(row, 1)

    def close(self):
        for e in self._enc.values():
            e.close()
        self._enc = {}


class PartitionedFile:
    def __init__(self, filePath: str, start: int = 0, length: Optional[int] = None):
        self.filePath = filePath
        self.start = start
        self.length = os.path.getsize(filePath) if length is None else length

    def toPath(self):
        return self.filePath


def _stream_blocks(f, remaining: int, block: int, stage, process):
    """The block loop shared by readFile and inferSchema: reads `f` in blocks of about `block` bytes into stage(nbytes) (a
    writable uint8 array), calls process(buffer, nbytes, is_final) -> consumed bytes and carries the unconsumed tail (a
    partial record) into the next block.  Yields after every block so that the caller can drain rows in between."""
    carry = b""
    while True:
        want = min(max(block - len(carry), block // 2), remaining)   # a carried record larger than the block still makes progress
        chunk = f.read(want) if want > 0 else b""
        remaining -= len(chunk)
        final = remaining == 0 or len(chunk) < want
        nbytes = len(carry) + len(chunk)
        st = stage(max(nbytes, 1))
        if carry:
            st[: len(carry)] = np.frombuffer(carry, dtype=np.uint8)
        if chunk:
            st[len(carry): nbytes] = np.frombuffer(chunk, dtype=np.uint8)
        used = process(st, nbytes, final)
        yield
        carry = st[used:nbytes].tobytes()
        if final:
            return


class TFRecordFileReader:
    BLOCK_BYTES = 256 << 20

    @staticmethod
    def readFile(conf, options: Dict[str, str], file: PartitionedFile, schema: StructType, device: int = 0,
                 block_bytes: Optional[int] = None) -> Iterator[tuple]:
        """Stages the file in blocks into the decoder's pinned staging slots and decodes each block on the GPU
        (tfr_decode_submit); where a block ends -- the carry into the next one -- is known as soon as its frame index has
        run (tfr_batch_consumed), so block k+1 is read and submitted while block k decodes and block k-1's rows are
        handed out.  Rows before a bad record are yielded, then the exception the reference would throw is raised."""
        rt = _record_type(options)
        block = block_bytes or TFRecordFileReader.BLOCK_BYTES
        dec = _native.Decoder(schema, rt, device, TFR_F_DEFAULT)

        def gen():
            todo = []
            try:
                compressed = _codec_of_path(file.toPath()) is not None
                with _open_read(file.toPath()) as f:
                    if not compressed:
                        f.seek(file.start)
                    remaining = (1 << 62) if compressed else file.length          # a compressed file is read to its end
                    n_slots = dec.num_staging_slots()
                    turn = [0]

                    def stage(nbytes):
                        st = dec.staging_slot(turn[0] % n_slots, nbytes)
                        turn[0] += 1
                        return st

                    def process(st, nbytes, final):
                        batch = dec.submit(st, is_final=final, nbytes=nbytes)
                        todo.append(batch)
                        return batch.consumed()

                    def drain(batch):
                        try:
                            for row in _rows_of(batch):
                                yield row
                            batch.raise_if_error()
                        finally:
                            batch.release()

                    for _ in _stream_blocks(f, remaining, block, stage, process):
                        while len(todo) > 1:                 # the block before the one just submitted
                            yield from drain(todo.pop(0))
                    while todo:
                        yield from drain(todo.pop(0))
            finally:
                for batch in todo:
                    batch.release()
                dec.close()

        return gen()


def _row_bytes(row) -> int:
    """rough size of a buffered row's values (what decides when the writer flushes)"""
    n = 0
    for v in row:
        if v is None:
            continue
        if isinstance(v, (bytes, bytearray, str)):
            n += len(v) + 8
        elif isinstance(v, (list, tuple)):
            n += 8 + sum((len(x) + 8) if isinstance(x, (bytes, bytearray, str)) else (8 * len(x) + 8 if isinstance(x, (list, tuple)) else 8) for x in v)
        else:
            n += 8
    return n + 16


class TFRecordOutputWriter:
    FLUSH_ROWS = 1 << 16
    FLUSH_BYTES = 256 << 20        # large rows (images, long byte strings) flush by size: one tfr_encode call frames < 2 GiB

    def __init__(self, path: str, options: Dict[str, str], dataSchema: StructType, context=None, device: int = 0):
        self.path = path
        self.recordType = _record_type(options)                 # validated up front; the reference throws at the first write
        self.schema = byte_array_schema() if self.recordType == 2 else dataSchema
        self._enc = _native.Encoder(self.schema, self.recordType, device)
        self._rows: List[tuple] = []
        self._bytes = 0
        self._out = _open_write(path, _codec_name((options or {}).get("codec", "")))   # CodecStreams.createOutputStream (:19)

    def write(self, row: Sequence) -> None:
        row = tuple(row)
        self._rows.append(row)
        self._bytes += _row_bytes(row)
        if len(self._rows) >= self.FLUSH_ROWS or self._bytes >= self.FLUSH_BYTES:
            self._flush()

    def _encode_rows(self, rows: List[tuple]) -> None:
        try:
            self._out.write(self._enc.encode(columns_from_rows(self.schema, rows, self.recordType)))
        except _native.TfrError as e:
            if e.code != A_TFR_E_BATCH_TOO_LARGE or len(rows) < 2:
                raise
            half = len(rows) // 2                                # the size estimate was too low: frame the rows in two calls
            self._encode_rows(rows[:half])
            self._encode_rows(rows[half:])

    def _flush(self):
        if self._rows:
            rows, self._rows, self._bytes = self._rows, [], 0
            self._encode_rows(rows)

    def close(self) -> None:
        try:
            self._flush()
        finally:
            self._out.close()
            self._enc.close()


class DefaultSource:
    """The FileFormat surface that stays (M/DefaultSource.scala:23-143): names and meanings only."""

    def shortName(self) -> str:
        return "tfrecord"

    def isSplitable(self, *a, **k) -> bool:
        return False                                             # :26-29; splitting happens inside the native side

    def inferSchema(self, options: Dict[str, str], files: Sequence[str], device: int = 0, dist=None):
        """M/DefaultSource.scala:31-39,48-70: the first non-empty file is scanned (the reference scans it twice);
        ByteArray has the fixed one-column schema.  With `dist`, every rank scans its shard of the files and the maps
        are merged with one all-reduce (sharding.allreduce_schema)."""
        from .sharding import allreduce_schema, codes_to_struct, shard_lpt
        rt = _record_type(options)
        if rt == 2:
            return byte_array_schema()
        todo = [f for f in files if os.path.getsize(f) > 0]
        if dist is None or not dist.is_initialized():
            todo = todo[:1]
        else:
            mine = shard_lpt([os.path.getsize(f) for f in todo], dist.get_world_size())[dist.get_rank()]
            todo = [todo[i] for i in mine]
        inf = _native.Infer(rt, device)
        block = TFRecordFileReader.BLOCK_BYTES
        buf = [np.empty(0, dtype=np.uint8)]

        def stage(nbytes):
            if len(buf[0]) < nbytes:
                buf[0] = np.empty(nbytes + nbytes // 8, dtype=np.uint8)
            return buf[0]

        try:
            for f in todo:                       # streamed in blocks like readFile: files of any size, no whole-file copy
                with _open_read(f) as fh:
                    remaining = (1 << 62) if _codec_of_path(f) is not None else os.path.getsize(f)
                    for _ in _stream_blocks(fh, remaining, block, stage, lambda st, nb, final: inf.update_block(st, final, nb)):
                        pass
            local = inf.result()
        finally:
            inf.close()
        return codes_to_struct(allreduce_schema(local, dist, f"cuda:{device}" if dist is not None and dist.is_initialized() and dist.get_backend() == "nccl" else None))

    def buildReader(self, dataSchema: StructType, requiredSchema: StructType, options: Dict[str, str], device: int = 0):
        """-> PartitionedFile => Iterator[row] (filters are accepted and ignored, :123)"""
        return lambda file: TFRecordFileReader.readFile(None, options, file, requiredSchema, device)

    def prepareWrite(self, options: Dict[str, str], dataSchema: StructType):
        codec = _codec_name((options or {}).get("codec", ""))             # :94-102: the option turns output compression on

        class _Factory:
            def newInstance(self_inner, path, schema, context=None):
                return TFRecordOutputWriter(path, options, schema, context)

            def getFileExtension(self_inner, context=None):                 # :110-112
                return ".tfrecord" + (_CODECS[codec][0] if codec else "")

        return _Factory()

    # convenience used by the tests: spark.read.format("tfrecord").schema(s).load(p) / df.write...save(p)
    def load(self, path: str, schema: StructType, options: Optional[Dict[str, str]] = None, device: int = 0) -> List[tuple]:
        files = [path] if os.path.isfile(path) else sorted(os.path.join(path, f) for f in os.listdir(path)
                                                           if not f.startswith(("_", ".")))
        reader = self.buildReader(schema, schema, options or {}, device)
        rows: List[tuple] = []
        for f in files:
            rows.extend(reader(PartitionedFile(f)))
        return rows

    def save(self, path: str, schema: StructType, rows: Sequence[Sequence], options: Optional[Dict[str, str]] = None) -> None:
        os.makedirs(path, exist_ok=True)
        factory = self.prepareWrite(options or {}, schema)
        w = factory.newInstance(os.path.join(path, "part-00000" + factory.getFileExtension()), schema)
        for r in rows:
            w.write(r)
        w.close()
        open(os.path.join(path, "_SUCCESS"), "wb").close()
