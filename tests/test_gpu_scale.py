"""Parity at BENCH scale and through the pipelined (speculative) decode: the configuration bench.py times is the one
compared here, bit for bit, with the CPU oracle.

  * a 1 GiB / 621 k-record configs[1] batch decoded three times on one decoder: the 2nd and 3rd decode run in the steady
    state (uniform-shape speculation, rows counted on the device, no host synchronisation) and are compared with the oracle
    over the whole batch (the oracle is threaded over record-aligned slices);
  * the pipelined API (tfr_decode_submit) with several batches in flight, batches released before they resolve, and every way
    the speculation can fail (a corrupt record, a shape change, more records than provisioned, a record larger than the
    tile slot): each must give exactly the oracle's result;
  * configs[2] (encode, 250 k rows, byte diff) and configs[3] (SequenceExample, 100 k records) at scale;
  * a device buffer whose allocation ends exactly at data + nbytes.

Reference semantics: M/TFRecordFileReader.scala:49-81, M/TFRecordDeserializer.scala:21-35, M/TFRecordSerializer.scala:20-35."""
import ctypes
import threading

import numpy as np
import pytest

from util import assert_columns_equal, slice_columns, record_offsets
from spark_tfrecord_b200 import _cabi as A
from spark_tfrecord_b200.sqltypes import *  # noqa

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def native():
    from spark_tfrecord_b200 import _native
    _native.lib()
    return _native


def oracle_check_slices(oracle, data: np.ndarray, schema, got, record_type=0, n_slices=16, what=""):
    """threads: oracle.decode over record-aligned slices of `data`; each slice's columns == the same rows of `got`"""
    offs = record_offsets(data)
    n = len(offs) - 1
    cuts = [int(round(i * n / n_slices)) for i in range(n_slices + 1)]
    errs = []

    def work(i):
        r0, r1 = cuts[i], cuts[i + 1]
        if r1 <= r0:
            return
        try:
            want = oracle.decode(data[offs[r0]:offs[r1]], schema, record_type)
            assert want.info["error_code"] == 0 and want.n_rows == r1 - r0, want.info
            assert_columns_equal(slice_columns(got, r0, r1), want.columns, schema.names, f"{what} rows [{r0},{r1})")
        except BaseException as e:      # noqa: BLE001
            errs.append(e)

    ths = [threading.Thread(target=work, args=(i,)) for i in range(n_slices)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    if errs:
        raise errs[0]
    return n


@pytest.fixture(scope="module")
def cfg2_gib(oracle):
    """configs[1] at bench scale: 621,378 records = 1 GiB of framed bytes, written by the oracle's writer"""
    from oracle.corpus import cfg2_columns
    n = (1024 << 20) // 1728
    sch, cols = cfg2_columns(n, seed=2024)
    data, rc, _ = oracle.encode(cols, sch)
    assert rc == 0
    return sch, cols, np.frombuffer(data, dtype=np.uint8)


def test_cfg2_one_gib_steady_state_is_bit_exact(native, oracle, cfg2_gib):
    import torch
    sch, cols, data = cfg2_gib
    n = cols[0].n_rows
    d_data = torch.from_numpy(data.copy()).cuda()
    dec = native.Decoder(sch)
    try:
        for it in range(3):
            batch, used = dec.decode(d_data)
            assert used == len(data) and batch.info["error_code"] == 0 and batch.n_rows == n, batch.info
            if it >= 1:
                got = batch.to_host()
                assert_columns_equal(got, cols, sch.names, f"decode #{it + 1} vs source")
                assert oracle_check_slices(oracle, data, sch, got, what=f"decode #{it + 1} vs oracle") == n
            batch.release()
        st = dec.stats()
        # decode #1 learns the shapes (count mode); #2 and #3 are the benchmarked mode
        assert st["speculative_submits"] == 2 and st["speculative_redone"] == 0 and st["count_mode_batches"] == 1, st
    finally:
        dec.close()


def _encode(oracle, sch, cols, record_type=0):
    data, rc, _ = oracle.encode(cols, sch, record_type)
    assert rc == 0
    return np.frombuffer(data, dtype=np.uint8)


def _check_batch(oracle, batch, data, sch, record_type=0, is_final=True, what=""):
    want = oracle.decode(data, sch, record_type, is_final=is_final)
    info = batch.info
    for k in ("error_code", "error_row", "error_field", "n_rows", "consumed_bytes"):
        assert info[k] == want.info[k], (what, k, info, want.info)
    assert_columns_equal(batch.to_host(), want.columns, sch.names, what)


def test_pipelined_submits_match_oracle(native, oracle):
    """several batches in flight (more than the decoder has lanes), different sizes, one released before it resolves"""
    import torch
    from oracle.corpus import cfg2_columns
    sch, _ = cfg2_columns(1, seed=1)
    datas = [_encode(oracle, sch, cfg2_columns(nr, seed=100 + i)[1]) for i, nr in enumerate([20000, 20000, 17000, 23000, 20011, 19999, 20000, 21000])]
    dev = [torch.from_numpy(d.copy()).cuda() for d in datas]
    dec = native.Decoder(sch)
    try:
        b0, _ = dec.decode(dev[0])             # learns the shapes
        b0.release()
        inflight = []
        for i in range(1, len(datas)):
            inflight.append((i, dec.submit(dev[i])))
            if i == 3:
                inflight.pop()[1].release()    # dropped while pending: its lane must still be recycled correctly
        for i, b in inflight:
            _check_batch(oracle, b, datas[i], sch, what=f"pipelined batch {i}")
            b.release()
        st = dec.stats()
        assert st["speculative_submits"] == len(datas) - 1 and st["speculative_redone"] == 0, st
        # host input through the staging slots: the same pipeline with H2D copies on the copy stream
        inflight = []
        for i in range(1, 7):
            slot = i % dec.num_staging_slots()
            for j, b in [x for x in inflight if x[0] % dec.num_staging_slots() == slot]:
                _check_batch(oracle, b, datas[j], sch, what=f"staged batch {j}")
                b.release()
                inflight.remove((j, b))
            st_buf = dec.staging_slot(slot, len(datas[i]))
            st_buf[: len(datas[i])] = datas[i]
            b = dec.submit(st_buf, nbytes=len(datas[i]))
            b.to_host_async()
            inflight.append((i, b))
        for j, b in inflight:
            _check_batch(oracle, b, datas[j], sch, what=f"staged batch {j}")
            b.release()
        assert dec.stats()["speculative_redone"] == 0
    finally:
        dec.close()


def _flip(data: np.ndarray, pos: int) -> np.ndarray:
    out = data.copy()
    out[pos] ^= 0x10
    return out


def test_speculation_failures_are_redone_exactly(native, oracle):
    """every flag the single-pass kernel can raise leads to the oracle's result: CRC error in the middle, a length-header
    flip, a shape change, a non-canonical record, more records than provisioned, a record larger than the slot"""
    from oracle import pyref
    from oracle.corpus import cfg2_columns
    sch, cols = cfg2_columns(12000, seed=5)
    good = _encode(oracle, sch, cols)
    offs = record_offsets(good)
    dec = native.Decoder(sch)
    try:
        for _ in range(2):
            b, _ = dec.decode(good); b.release()
        assert dec.stats()["speculative_submits"] == 1

        def steady():
            """the decoder is (back) in the pipelined mode: a clean batch is submitted speculatively and not redone"""
            s0 = dec.stats()
            b = dec.submit(good); _check_batch(oracle, b, good, sch, what="steady state"); b.release()
            s1 = dec.stats()
            assert s1["speculative_submits"] == s0["speculative_submits"] + 1 and s1["speculative_redone"] == s0["speculative_redone"], (s0, s1)

        def redone_exactly(data, what, code=None):
            s0 = dec.stats()
            b = dec.submit(data); _check_batch(oracle, b, data, sch, what=what)
            if code is not None:
                assert b.info["error_code"] == code
            b.release()
            s1 = dec.stats()
            assert s1["speculative_submits"] == s0["speculative_submits"] + 1 and s1["speculative_redone"] == s0["speculative_redone"] + 1, (what, s0, s1)

        # payload bit flip in record 7000 -> Data crc32 checking failed at row 7000, rows before delivered
        redone_exactly(_flip(good, int(offs[7000]) + 40), "payload flip", A.TFR_E_CRC_DATA)
        steady()
        # a flip inside a list-length byte looks like a shape change to the kernel: still a CRC error, shapes kept
        for pos in range(14, 40):
            redone_exactly(_flip(good, int(offs[5000]) + pos), f"payload flip at +{pos}", A.TFR_E_CRC_DATA)
        steady()
        # length flip in record 300 (low byte of the length): length CRC error at that row
        redone_exactly(_flip(good, int(offs[300])), "length flip", A.TFR_E_CRC_LENGTH)
        steady()
        redone = dec.stats()["speculative_redone"]
        # shape change: FloatList[7] in one column of a second corpus -> count mode, then the new shapes are learned
        sch2, cols2 = cfg2_columns(9000, seed=6, float_len=7)
        other = _encode(oracle, sch2, cols2)
        for k in range(3):
            b = dec.submit(other); _check_batch(oracle, b, other, sch, what=f"other shapes #{k}"); b.release()
        st = dec.stats()
        assert st["shapes_learned"] == 2 and st["speculative_redone"] == redone + 1, st
        redone += 1
        good = other
        # a non-canonical record (unpacked floats) among canonical ones: general path for the batch, identical rows
        k = 4000
        rec = bytes(other[record_offsets(other)[k] + 12: record_offsets(other)[k + 1] - 4])
        odd = pyref.frame_fast(rec + b"\x0a\x00")           # `features` field repeated (empty): merges, not canonical
        o2 = record_offsets(other)
        mixed = np.concatenate([other[: o2[k]], np.frombuffer(odd, np.uint8), other[o2[k + 1]:]])
        redone_exactly(mixed, "non-canonical record")
        steady()
        # one record much larger than every slot seen so far (a long extra feature the schema ignores)
        def ld(tag, payload):
            return bytes([tag]) + pyref.varint(len(payload)) + payload
        entry = ld(0x0A, ld(0x0A, b"zz") + ld(0x12, ld(0x0A, ld(0x0A, bytes(3000)))))
        rec = bytes(other[o2[10] + 12: o2[11] - 4])
        assert rec[0] == 0x0A and rec[1] & 0x80
        rec2 = ld(0x0A, rec[3:] + entry)
        wide = np.concatenate([other[: o2[10]], np.frombuffer(pyref.frame_fast(rec2), np.uint8), other[o2[11]:]])
        redone_exactly(wide, "record larger than the slot")
        b = dec.submit(wide); _check_batch(oracle, b, wide, sch, what="record larger than the slot, again"); b.release()
    finally:
        dec.close()


def test_more_records_than_provisioned(native, oracle):
    """the output capacity of a pipelined batch comes from the previous batch's record size: a block with four times as many
    (smaller) records overflows it, is flagged on the device and redone"""
    from spark_tfrecord_b200._cabi import HostColumn
    sch = StructType([StructField(f"c{i}", LongType()) for i in range(8)])

    def corpus(n, present, seed):
        rng = np.random.default_rng(seed)
        cols = []
        for i in range(8):
            bits = np.full(n, i < present)
            v = rng.integers(-2**40, 2**40, n, dtype=np.int64)
            v[~bits] = 0
            cols.append(HostColumn(A.TFR_T_INT64, 0, n, np.packbits(bits, bitorder="little"), [], v))
        return _encode(oracle, sch, cols)

    full = corpus(40000, 8, 1)
    thin = corpus(190000, 1, 2)
    assert len(record_offsets(thin)) - 1 > 1.2 * (len(thin) / (len(full) / 40000))
    dec = native.Decoder(sch)
    try:
        for _ in range(2):
            b, _ = dec.decode(full); b.release()
        assert dec.stats()["speculative_submits"] == 1
        b = dec.submit(thin); _check_batch(oracle, b, thin, sch, what="thin records"); b.release()
        assert dec.stats()["speculative_redone"] == 1
        b = dec.submit(thin); _check_batch(oracle, b, thin, sch, what="thin records, capacity relearned"); b.release()
        b = dec.submit(full); _check_batch(oracle, b, full, sch, what="back to full records"); b.release()
        assert dec.stats()["speculative_redone"] == 1
    finally:
        dec.close()


def test_nonfinal_block_with_corrupt_length_is_an_error_not_a_tail(native, oracle):
    """a bit-flipped length that points past the block must be 'Length header crc32 checking failed' at that record, not a
    partial record to carry over (the unverified chain of the fast path sees only the length)"""
    from oracle.corpus import cfg2_columns
    sch, cols = cfg2_columns(3000, seed=12)
    good = _encode(oracle, sch, cols)
    offs = record_offsets(good)
    bad = good.copy()
    bad[int(offs[2990]) + 2] ^= 0x40          # length + 4 MiB: runs past the end of the block
    for final in (False, True):
        dec = native.Decoder(sch)
        try:
            for _ in range(2):
                b, _ = dec.decode(good); b.release()
            b = dec.submit(bad, is_final=final)
            _check_batch(oracle, b, bad, sch, is_final=final, what=f"corrupt length, is_final={final}")
            assert b.info["error_code"] == A.TFR_E_CRC_LENGTH and b.info["error_row"] == 2990
            b.release()
            # first decode of a fresh decoder (count mode, unverified chain) as well
            d2 = native.Decoder(sch)
            b, used = d2.decode(bad, is_final=final)
            _check_batch(oracle, b, bad, sch, is_final=final, what="fresh decoder")
            b.release(); d2.close()
        finally:
            dec.close()


def test_cfg3_encode_250k_rows_byte_identical(native, oracle):
    from oracle.corpus import cfg2_columns
    sch, cols = cfg2_columns(250_000, seed=31)
    want, rc, _ = oracle.encode(cols, sch)
    assert rc == 0
    enc = native.Encoder(sch)
    try:
        got = enc.encode(cols)
        assert len(got) == len(want)
        assert got == want, "GPU encoder bytes differ from the reference writer restatement"
        got2 = enc.encode(cols)                # second call: slot sizes learned from the first
        assert got2 == want
    finally:
        enc.close()


def test_cfg4_sequence_example_100k(native, oracle):
    from oracle.corpus import cfg4_columns
    sch, cols = cfg4_columns(100_000, seed=77, mean_steps=64)
    data = _encode(oracle, sch, cols, TFR_RT_SEQUENCE_EXAMPLE)
    dec = native.Decoder(sch, TFR_RT_SEQUENCE_EXAMPLE)
    try:
        for it in range(2):
            batch, used = dec.decode(data)
            assert used == len(data) and batch.info["error_code"] == 0 and batch.n_rows == 100_000
            got = batch.to_host()
            assert_columns_equal(got, cols, sch.names, f"cfg4 decode #{it + 1} vs source")
            if it == 1:
                oracle_check_slices(oracle, data, sch, got, record_type=TFR_RT_SEQUENCE_EXAMPLE, n_slices=8, what="cfg4 vs oracle")
            batch.release()
    finally:
        dec.close()


def test_device_buffer_ending_exactly_at_nbytes(native, oracle):
    """data_on_device input needs no padding: cudaMalloc(nbytes) exactly, nbytes not a multiple of 16, so the 16-byte group
    that holds the last record's tail crosses the end of the allocation.  Correctness is checked here; the absence of any
    out-of-bounds read is what compute-sanitizer verifies on this test (profiles/r2_memcheck_exact_end.log)."""
    import torch
    from oracle.corpus import cfg2_columns
    torch.cuda.init()
    rt = ctypes.CDLL("libcudart.so.12")
    rt.cudaMalloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]
    rt.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    rt.cudaFree.argtypes = [ctypes.c_void_p]
    sch, cols = cfg2_columns(4099, seed=77)
    full = _encode(oracle, sch, cols)
    offs = record_offsets(full)
    dec = native.Decoder(sch)
    ptrs = []
    tested = 0
    try:
        for n_rec in (4099, 4098, 4097, 4096, 4095, 4090, 33, 1):
            data = full[: offs[n_rec]]
            nb = len(data)
            if nb % 16 == 0:
                continue                      # nothing would cross the end
            p = ctypes.c_void_p()
            assert rt.cudaMalloc(ctypes.byref(p), nb) == 0
            ptrs.append(p)
            assert p.value % 16 == 0
            assert rt.cudaMemcpy(p, data.ctypes.data, nb, 1) == 0
            want = oracle.decode(data, sch)
            for it in range(2):
                batch, used = dec.decode((p.value, nb, 1))
                assert used == nb and batch.info["error_code"] == 0 and batch.n_rows == n_rec, batch.info
                assert_columns_equal(batch.to_host(), want.columns, sch.names, f"exact-end buffer, {n_rec} records, decode #{it + 1}")
                batch.release()
            tested += 1
        assert tested >= 4
    finally:
        dec.close()
        for p in ptrs:
            rt.cudaFree(p)


def test_ragged_columns_one_pass_matches_oracle(native, oracle):
    """ragged columns in the pipelined mode: the tile kernel finishes them itself (tile-local prefix + look-back across
    tiles) -- scalar strings / binaries, lists of every element type, lists of strings, nulls, empty lists, many tiles"""
    from oracle.corpus import mixed_columns
    sch, _ = mixed_columns(10, seed=1)
    datas = [_encode(oracle, sch, mixed_columns(n, seed=40 + i)[1]) for i, n in enumerate([30000, 30000, 30011, 25000, 33000, 31, 30000])]
    dec = native.Decoder(sch)
    try:
        b, _ = dec.decode(datas[0]); _check_batch(oracle, b, datas[0], sch, what="learning batch"); b.release()
        inflight = [(i, dec.submit(datas[i])) for i in range(1, 4)]
        for i, b in inflight:
            _check_batch(oracle, b, datas[i], sch, what=f"ragged batch {i}"); b.release()
        for i in range(4, len(datas)):
            b = dec.submit(datas[i]); _check_batch(oracle, b, datas[i], sch, what=f"ragged batch {i}"); b.release()
        st = dec.stats()
        assert st["speculative_submits"] >= 5 and st["speculative_redone"] <= 1, st      # (the 31-record batch may be too small to speculate on)
        assert st["count_mode_batches"] <= 2, st
    finally:
        dec.close()


def test_ragged_bytes_cfg2_steady_state(native, oracle):
    """configs[1] with BytesList values of 0..40 bytes: float lists stay uniform, bytes columns are ragged; 300 k records"""
    import importlib.util, os, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod2", os.path.join(root, "bench.py"))
    bm = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(bm)
    finally:
        sys.argv = argv
    sch, cols = bm.cfg2_schema_and_columns(300_000, seed=777, ragged_bytes=True)
    data = _encode(oracle, sch, cols)
    dec = native.Decoder(sch)
    try:
        for it in range(3):
            b = dec.submit(data)
            assert b.info["error_code"] == 0 and b.n_rows == 300_000
            got = b.to_host()
            assert_columns_equal(got, cols, sch.names, f"ragged cfg2 decode #{it + 1} vs source")
            if it == 2:
                oracle_check_slices(oracle, data, sch, got, n_slices=8, what="ragged cfg2 vs oracle")
            b.release()
        st = dec.stats()
        assert st["speculative_submits"] == 2 and st["speculative_redone"] == 0 and st["count_mode_batches"] == 1, st
        # a batch whose strings are much longer than the capacities learned: flagged (capacity), redone, relearned
        sch2, cols2 = bm.cfg2_schema_and_columns(20_000, seed=778, ragged_bytes=True)
        from spark_tfrecord_b200._cabi import HostColumn
        rng = np.random.default_rng(5)
        for i in range(48, 64):
            lens = rng.integers(100, 200, 20_000)
            offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
            cols2[i] = HostColumn(A.TFR_T_BINARY, 0, 20_000, cols2[i].validity, [offs], rng.integers(0, 256, int(offs[-1]), dtype=np.uint8))
        longer = _encode(oracle, sch2, cols2)
        b = dec.submit(longer); _check_batch(oracle, b, longer, sch, what="longer strings than provisioned"); b.release()
        assert dec.stats()["speculative_redone"] == 1
        b = dec.submit(longer); _check_batch(oracle, b, longer, sch, what="longer strings, capacities relearned"); b.release()
        assert dec.stats()["speculative_redone"] == 1
    finally:
        dec.close()


def test_malformed_utf8_in_ragged_string_columns_is_not_a_cliff(native, oracle):
    """StringType cells with malformed UTF-8 (Java re-encodes them with U+FFFD, M/TFRecordDeserializer.scala:91,215) in ragged
    columns are handled inside the single-pass kernel: same bytes as the oracle, no redo of the batch"""
    from oracle.corpus import mixed_columns
    from spark_tfrecord_b200._cabi import HostColumn
    sch, cols = mixed_columns(20000, seed=9)
    names = sch.names
    rng = np.random.default_rng(3)
    bad_seqs = [b"\xff", b"\xc3", b"\xe2\x82", b"\xed\xa0\x80", b"\xf0\x9f\x98", b"\xc0\xaf", b"ok\x80ok", b"\xf5\x80\x80\x80"]

    def poison(col, every):
        data = col.values.copy()
        so = col.offsets[-1]
        k = 0
        for i in range(0, len(so) - 1, every):
            a, b = int(so[i]), int(so[i + 1])
            seq = bad_seqs[k % len(bad_seqs)]; k += 1
            if b - a >= len(seq):
                data[a:a + len(seq)] = np.frombuffer(seq, np.uint8)
        return HostColumn(col.elem_type, col.depth, col.n_rows, col.validity, col.offsets, data)

    clean = _encode(oracle, sch, cols)
    cols2 = list(cols)
    cols2[names.index("s")] = poison(cols[names.index("s")], 97)
    cols2[names.index("as")] = poison(cols[names.index("as")], 53)
    dirty = _encode(oracle, sch, cols2)
    want = oracle.decode(dirty, sch)
    assert not np.array_equal(want.columns[names.index("s")].values, cols2[names.index("s")].values)      # something was re-encoded
    dec = native.Decoder(sch)
    try:
        b, _ = dec.decode(clean); b.release()
        s0 = dec.stats()
        b = dec.submit(dirty); _check_batch(oracle, b, dirty, sch, what="malformed UTF-8 in ragged string columns"); b.release()
        s1 = dec.stats()
        # the default kernel met the first malformed string and handed the batch to its transcoding instantiation: one more
        # single-pass run, not the general path
        assert s1["transcode_reruns"] == 1 and s1["speculative_redone"] == s0["speculative_redone"] and s1["general_path_batches"] == 0, (s0, s1)
        b = dec.submit(dirty); _check_batch(oracle, b, dirty, sch, what="malformed UTF-8, transcoding kernel already active"); b.release()
        s2 = dec.stats()
        assert s2["transcode_reruns"] == 1 and s2["speculative_submits"] == s1["speculative_submits"] + 1 and s2["speculative_redone"] == s1["speculative_redone"], (s1, s2)
    finally:
        dec.close()


def _frame_payloads(payloads):
    from oracle import pyref
    return np.frombuffer(b"".join(pyref.frame_fast(p) for p in payloads), dtype=np.uint8)


def _bytes_rows_of(batch):
    c = batch.to_host()[0]
    o = c.offsets[0]
    v = c.values.view(np.uint8)
    return c, [v[o[i]:o[i + 1]].tobytes() for i in range(len(o) - 1)]


def test_bytearray_rows_single_pass_and_pipelined(native, oracle):
    """recordType=ByteArray through decode_bytes_kernel: the synchronising decode, the pipelined submit (device buffers at every
    alignment, host staging), a CRC error and a record larger than the slot in the middle of the pipeline -- rows, offsets,
    validity and the error position must be the oracle's (M/TFRecordDeserializer.scala:17-19)"""
    import torch
    rng = np.random.default_rng(2024)
    sch = byte_array_schema()

    def make(n, lo, hi, seed):
        r = np.random.default_rng(seed)
        sizes = r.integers(lo, hi, n)
        sizes[:: 97] = 0                                            # empty payloads are rows too
        blob = r.integers(0, 256, int(sizes.sum()), dtype=np.uint8).tobytes()
        pos = np.concatenate([[0], np.cumsum(sizes)])
        return [blob[pos[i]:pos[i + 1]] for i in range(n)]

    batches = [make(n, lo, hi, 50 + i) for i, (n, lo, hi) in enumerate([(30000, 0, 2000), (30011, 0, 2000), (29000, 0, 2040), (31000, 0, 1990), (30000, 900, 1100), (1, 5, 6)])]
    datas = [_frame_payloads(p) for p in batches]
    dec = native.Decoder(sch, TFR_RT_BYTE_ARRAY)
    try:
        # synchronising decode
        b, used = dec.decode(torch.from_numpy(datas[0].copy()).cuda())
        assert used == len(datas[0]) and b.info["error_code"] == 0 and b.n_rows == len(batches[0]), b.info
        c, rows = _bytes_rows_of(b)
        assert rows == batches[0] and c.null_count == 0 and np.all(np.unpackbits(c.validity, bitorder="little")[: b.n_rows] == 1)
        want = oracle.decode(datas[0], sch, 2)
        assert_columns_equal(b.to_host(), want.columns, ["byteArray"], "ByteArray rows, first decode")
        b.release()
        assert dec.stats()["general_path_batches"] == 0, dec.stats()
        # pipelined, device buffers starting at every alignment mod 16
        inflight = []
        keep = []
        for i in range(1, len(datas)):
            t = torch.empty(len(datas[i]) + 16, dtype=torch.uint8, device="cuda")
            view = t[i % 16: i % 16 + len(datas[i])]
            view.copy_(torch.from_numpy(datas[i].copy()))
            keep.append(t)
            inflight.append((i, dec.submit(view)))
        for i, b in inflight:
            assert b.info["error_code"] == 0 and b.info["consumed_bytes"] == len(datas[i]) and b.n_rows == len(batches[i]), (i, b.info)
            assert _bytes_rows_of(b)[1] == batches[i], f"pipelined ByteArray batch {i}"
            b.release()
        st = dec.stats()
        assert st["speculative_submits"] >= len(datas) - 2 and st["general_path_batches"] == 0, st
        # host input through the staging slots + asynchronous copy-out
        inflight = []
        for i in range(1, 5):
            slot = i % dec.num_staging_slots()
            for j, b in [x for x in inflight if x[0] % dec.num_staging_slots() == slot]:
                assert _bytes_rows_of(b)[1] == batches[j]
                b.release(); inflight.remove((j, b))
            sb = dec.staging_slot(slot, len(datas[i]))
            sb[: len(datas[i])] = datas[i]
            b = dec.submit(sb, nbytes=len(datas[i]))
            b.to_host_async()
            inflight.append((i, b))
        for j, b in inflight:
            assert _bytes_rows_of(b)[1] == batches[j], f"staged ByteArray batch {j}"
            b.release()
        # a flipped payload bit in the middle of a pipelined batch: the error is reported at that record, rows before it stand
        k = 12345
        offs = record_offsets(datas[1])
        bad = datas[1].copy(); bad[offs[k] + 12 + len(batches[1][k]) // 2] ^= 4
        if len(batches[1][k]) == 0:
            bad[offs[k] + 12] ^= 4               # (the CRC field itself)
        b = dec.submit(torch.from_numpy(bad).cuda())
        assert b.info["error_code"] == A.TFR_E_CRC_DATA and b.info["error_row"] == k and b.n_rows == k, b.info
        assert _bytes_rows_of(b)[1] == batches[1][:k]
        b.release()
        # a flipped length-CRC bit
        bad = datas[2].copy(); offs2 = record_offsets(datas[2]); bad[offs2[777] + 9] ^= 1
        b = dec.submit(torch.from_numpy(bad).cuda())
        assert b.info["error_code"] == A.TFR_E_CRC_LENGTH and b.info["error_row"] == 777 and b.n_rows == 777, b.info
        b.release()
        # a record much larger than anything seen so far (slot overflow -> redone), then one too large for any tile (general kernels)
        for big in (5000, 300000):
            rows = list(batches[3][:5000]); rows[2500] = rng.integers(0, 256, big, dtype=np.uint8).tobytes()
            d = _frame_payloads(rows)
            b = dec.submit(torch.from_numpy(d.copy()).cuda())
            assert b.info["error_code"] == 0 and b.n_rows == len(rows), b.info
            assert _bytes_rows_of(b)[1] == rows, f"ByteArray batch with a {big}-byte record"
            b.release()
        # and back to the pipeline afterwards
        b = dec.submit(torch.from_numpy(datas[4].copy()).cuda())
        assert _bytes_rows_of(b)[1] == batches[4]
        b.release()
    finally:
        dec.close()


def test_bytearray_exact_end_buffer(native, oracle):
    """ByteArray rows from a device allocation that ends exactly at data + nbytes (no padding, see the Example variant above)"""
    import torch
    torch.cuda.init()
    rt = ctypes.CDLL("libcudart.so.12")
    rt.cudaMalloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]
    rt.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    rt.cudaFree.argtypes = [ctypes.c_void_p]
    rng = np.random.default_rng(5)
    sch = byte_array_schema()
    dec = native.Decoder(sch, TFR_RT_BYTE_ARRAY)
    ptrs = []
    try:
        for n_rec, last in ((4000, 1001), (4001, 7), (33, 1), (1, 3), (2, 0)):
            rows = [rng.integers(0, 256, int(s), dtype=np.uint8).tobytes() for s in list(rng.integers(0, 1500, n_rec - 1)) + [last]]
            data = _frame_payloads(rows)
            nb = len(data)
            p = ctypes.c_void_p()
            assert rt.cudaMalloc(ctypes.byref(p), nb) == 0
            ptrs.append(p)
            assert rt.cudaMemcpy(p, data.ctypes.data, nb, 1) == 0
            for it in range(2):
                b, used = dec.decode((p.value, nb, 1))
                assert used == nb and b.info["error_code"] == 0 and b.n_rows == n_rec, b.info
                assert _bytes_rows_of(b)[1] == rows, f"exact-end ByteArray buffer, {n_rec} records, decode #{it + 1}"
                b.release()
    finally:
        dec.close()
        for p in ptrs:
            rt.cudaFree(p)


def test_consumed_is_known_before_the_rows(native, oracle):
    """tfr_batch_consumed: a streaming reader cuts block t+1 from block t's consumed count right after submitting it, with block t
    still in flight (M/TFRecordFileReader.scala:49-61 reads record after record; a block boundary is wherever the last complete
    record ends).  Every block's early count must equal its final consumed_bytes and the rows of all blocks together must be
    the file's rows; a corrupt record ends the stream with the oracle's error at the oracle's row."""
    import torch
    from oracle.corpus import cfg2_columns
    sch, cols = cfg2_columns(60000, seed=31)
    data = _encode(oracle, sch, cols)
    offs = record_offsets(data)
    d_data = torch.from_numpy(data.copy()).cuda()
    dec = native.Decoder(sch)
    try:
        for block in (7 << 20, 3_333_333):
            pos, row0, inflight = 0, 0, []
            while pos < len(data):
                take = min(block, len(data) - pos)
                final = pos + take == len(data)
                b = dec.submit((d_data.data_ptr() + pos, take, 1), is_final=final)
                used = b.consumed()                          # before anything else is asked of the batch
                want_used = int(offs[np.searchsorted(offs, pos + take, side="right") - 1]) - pos
                assert used == want_used, (pos, take, used, want_used)
                inflight.append((b, pos, used))
                pos += used
                if len(inflight) == 3:
                    ob, p0, u0 = inflight.pop(0)
                    assert ob.info["error_code"] == 0 and ob.info["consumed_bytes"] == u0, ob.info
                    r1 = row0 + ob.n_rows
                    assert_columns_equal(ob.to_host(), slice_columns(cols, row0, r1), sch.names, f"block at {p0}")
                    row0 = r1
                    ob.release()
            for ob, p0, u0 in inflight:
                assert ob.info["error_code"] == 0 and ob.info["consumed_bytes"] == u0, ob.info
                r1 = row0 + ob.n_rows
                assert_columns_equal(ob.to_host(), slice_columns(cols, row0, r1), sch.names, f"block at {p0}")
                row0 = r1
                ob.release()
            assert row0 == 60000
        assert dec.stats()["speculative_submits"] > 10
        # a flipped payload byte in record 40000: the block that holds it reports the error; its final consumed count stops in front of the record
        bad = data.copy(); bad[offs[40000] + 100] ^= 1
        d_bad = torch.from_numpy(bad).cuda()
        b = dec.submit((d_bad.data_ptr() + int(offs[39000]), int(offs[41000] - offs[39000]), 1), is_final=True)
        early = b.consumed()
        assert early == int(offs[41000] - offs[39000])
        assert b.info["error_code"] == A.TFR_E_CRC_DATA and b.info["error_row"] == 1000 and b.info["consumed_bytes"] == int(offs[40000] - offs[39000]), b.info
        assert b.consumed() == b.info["consumed_bytes"]          # resolved: the final word
        b.release()
    finally:
        dec.close()


def test_bytearray_encode_single_pass(native, oracle):
    """recordType=ByteArray rows -> framed records through encode_bytes_kernel: byte-identical to the oracle writer for host
    columns, for device columns whose values pointer is misaligned and whose offsets do not start at 0 (a sliced Arrow array),
    and through the general kernels when a row is too large for a tile (serializeByteArray M/TFRecordSerializer.scala:16-18)"""
    import torch
    from spark_tfrecord_b200._cabi import tfr_column
    rng = np.random.default_rng(77)
    sch = byte_array_schema()
    for n, hi in ((1, 5), (33, 40), (50000, 2000), (20000, 300)):
        sizes = rng.integers(0, hi, n)
        sizes[:: 53] = 0
        rows = [(rng.integers(0, 256, int(s), dtype=np.uint8).tobytes(),) for s in sizes]
        cols = A.columns_from_rows(sch, rows)
        want, rc, _ = oracle.encode(cols, sch, TFR_RT_BYTE_ARRAY)
        assert rc == 0
        enc = native.Encoder(sch, TFR_RT_BYTE_ARRAY, 0)
        try:
            for it in range(2):
                assert bytes(enc.encode(cols)) == bytes(want), f"{n} ByteArray rows, call {it + 1}"
            # the same rows as a slice of a larger device-resident Arrow array: 7 junk bytes in front, values at an odd address
            c = cols[0]
            junk = rng.integers(0, 256, 7, dtype=np.uint8)
            vals = torch.from_numpy(np.concatenate([np.zeros(5, np.uint8), junk, c.values.view(np.uint8)])).cuda()
            offs = torch.from_numpy((c.offsets[0].astype(np.int64) + 7).astype(np.int32)).cuda()
            valid = torch.from_numpy(c.validity).cuda()
            t = tfr_column()
            hc = c.to_ctypes()
            for f, _ in tfr_column._fields_:
                setattr(t, f, getattr(hc, f))
            t.validity = valid.data_ptr(); t.offsets[0] = offs.data_ptr(); t.values = vals.data_ptr() + 5
            enc.encode_columns([t], True)
            assert bytes(enc.result_host()) == bytes(want), f"{n} ByteArray rows, sliced device column"
        finally:
            enc.close()
    # one row larger than any tile: the general kernels take the batch
    rows = [(rng.integers(0, 256, int(s), dtype=np.uint8).tobytes(),) for s in list(rng.integers(0, 500, 100)) + [300000] + list(rng.integers(0, 500, 100))]
    cols = A.columns_from_rows(sch, rows)
    want, rc, _ = oracle.encode(cols, sch, TFR_RT_BYTE_ARRAY)
    enc = native.Encoder(sch, TFR_RT_BYTE_ARRAY, 0)
    try:
        assert bytes(enc.encode(cols)) == bytes(want)
    finally:
        enc.close()
