"""Span-file reader (syzgydb_b200/csrc/spanfile.cu) against the Python restatement of the
reference's writer/reader (oracle/spanfile.py).  The parse tests need no GPU; the load test does."""
import json
import os
import zlib

import numpy as np
import pytest

import syzgydb_b200 as szg
from oracle import pyoracle as o
from oracle import spanfile as sfo
from tests.common import assert_results_match


def _build_collection(n, dims, bits, metric, seed, updates=0, removes=0, meta=True):
    """A span file written the way the reference writes it: header, n documents, then updates (re-written
    records move and leave FREE spans behind) and removals."""
    w = sfo.SpanFileWriter()
    w.write_header("fixture", metric, dims, bits)
    codes = o.synth_rows(seed, 0, n, dims, bits)
    ids = (np.arange(n, dtype=np.uint64) * 3 + 1)
    rng = np.random.default_rng(seed)
    metas = {}
    for i in range(n):
        m = json.dumps({"bucket": int(ids[i] % 10), "pad": "x" * int(rng.integers(0, 40))}).encode() if meta else b""
        metas[int(ids[i])] = m
        w.add_document(int(ids[i]), codes[i].tobytes(), m)
    live = {int(ids[i]): codes[i].copy() for i in range(n)}
    for j in rng.choice(n, size=updates, replace=False) if updates else []:
        new = o.synth_rows(seed + 1000 + int(j), 0, 1, dims, bits)[0]
        m = json.dumps({"bucket": 99, "pad": "y" * int(rng.integers(0, 200))}).encode()
        w.add_document(int(ids[j]), new.tobytes(), m)
        live[int(ids[j])] = new
        metas[int(ids[j])] = m
    for j in rng.choice(n, size=removes, replace=False) if removes else []:
        w.remove_document(int(ids[j]))
        live.pop(int(ids[j]))
        metas.pop(int(ids[j]))
    return w, live, metas


def test_7code_and_crc_known_answers():
    # write7Code's thresholds are off by one (spanfile.go:568-571): 126 -> 1 byte, 127 -> 2 bytes
    assert sfo.write_7code(0) == b"\x00" and sfo.write_7code(126) == b"\x7e"
    assert sfo.write_7code(127) == b"\x80\x7f" and sfo.write_7code(128) == b"\x81\x00"
    assert sfo.write_7code(0x3ffe) == b"\xff\x7e" and sfo.write_7code(0x3fff) == b"\x80\xff\x7f"
    for n in [0, 1, 126, 127, 128, 16382, 16383, 16384, 2 ** 21, 2 ** 28 - 1, 2 ** 35, 2 ** 63, 2 ** 64 - 1]:
        enc = sfo.write_7code(n)
        assert len(enc) == min(sfo.length_of_7code(n), 9)
        assert sfo.read_7code(enc + b"\x55", 0) == (n & ((1 << 63) - 1) if len(enc) == 9 and n >= 2 ** 63 else n, len(enc))
    assert sfo.crc32_ieee(b"123456789") == 0xCBF43926  # the CRC-32/IEEE check value
    # the minimal span OpenFile writes into an empty file (spanfile.go:216-238): 15 bytes
    w = sfo.SpanFileWriter()
    first = w.tobytes()[:15]
    assert first[:4] == b"SPAN" and int.from_bytes(first[4:8], "big") == 15 and first[8:11] == b"\x00\x00\x00"
    assert int.from_bytes(first[11:15], "big") == zlib.crc32(first[:11])


def test_writer_layout_follows_the_reference(tmp_path):
    w, live, metas = _build_collection(50, 8, 8, 1, 5, updates=10, removes=5)
    data = w.tobytes()
    # the header replaced the initial span: the file starts with that span marked FREE (spanfile.go:459-472)
    assert data[:4] == b"FREE" and int.from_bytes(data[4:8], "big") == 15
    assert len(data) % 1 == 0 and len(data) >= 4096  # grown by max(4096, need, 5 %) (spanfile.go:485)
    index, stats = sfo.scan_file(data)
    assert stats["corrupt"] == 0 and stats["free"] >= 1
    assert set(index) == {b""} | {str(i).encode() for i in live}
    opts, recs = sfo.live_records(data)
    assert opts == {"name": "fixture", "distance_method": 1, "dimension_count": 8, "quantization": 8}
    assert [r[0] for r in recs] == sorted(live, key=str)


@pytest.mark.parametrize("bits,dims,n,updates,removes", [(8, 24, 300, 40, 25), (4, 33, 120, 0, 0), (16, 10, 200, 60, 60),
                                                         (32, 7, 64, 5, 0), (64, 3, 90, 30, 10)])
def test_reader_matches_the_restated_scan(tmp_path, bits, dims, n, updates, removes):
    w, live, metas = _build_collection(n, dims, bits, 0, 100 + bits, updates=updates, removes=removes)
    path = os.path.join(tmp_path, "c.dat")
    with open(path, "wb") as f:
        f.write(w.tobytes())
    index, stats = sfo.scan_file(w.tobytes())
    with szg.SpanFile(path) as sf:
        info = sf.info()
        assert info["has_header"] == 1 and info["name"] == "fixture"
        assert (info["distance_method"], info["dimension_count"], info["quantization"]) == (0, dims, bits)
        assert info["records"] == len(live) and info["spans_corrupt"] == 0 and info["foreign_records"] == 0
        assert info["spans_active"] == stats["active"] and info["spans_free"] == stats["free"]
        assert info["next_sequence"] == stats["highest_seq"] + 1 and info["file_bytes"] == len(w.tobytes())
        assert sf.ids().tolist() == sorted(live, key=str)  # IterateSortedRecords order
        for i, code in live.items():
            vec, meta = sf.record(i)
            assert vec == code.tobytes() and meta == metas[i]
        with pytest.raises(KeyError):
            sf.record(2 ** 40)


def test_corrupt_spans_superseded_versions_and_foreign_ids(tmp_path):
    w, live, metas = _build_collection(40, 6, 8, 0, 9, meta=False)
    w.write_record(b"not-a-number", [(0, b"{}"), (1, b"\x00" * 6)])   # skipped by the reload loop (collection.go:299-302)
    w.write_record(b"007", [(0, b""), (1, b"\x01" * 6)])              # non-canonical spelling: unreachable by getDocument
    data = bytearray(w.tobytes())
    index, _ = sfo.scan_file(bytes(data))
    # flip a byte inside document 4's vector: its checksum fails, the span is skipped, the record disappears
    victim = index[b"4"]
    data[victim.offset + victim.length - 6] ^= 0xFF
    # append an OLDER version of document 7 (lower sequence number) after everything else: it must lose
    tail = sfo.serialize_span(0, b"7", [(0, b"old"), (1, b"\xEE" * 6)])
    tail += sfo.crc32_ieee(bytes(tail)).to_bytes(4, "big")
    end = len(data)
    while data[end - 1] == 0:
        end -= 1
    # the zero tail starts right after the last span / free marker; find it by scanning
    off = 0
    while off + 15 <= len(data) and int.from_bytes(data[off:off + 4], "big") != 0:
        off += int.from_bytes(data[off + 4:off + 8], "big")
    data[off:off + len(tail)] = tail
    path = os.path.join(tmp_path, "d.dat")
    with open(path, "wb") as f:
        f.write(data)
    ref_index, ref_stats = sfo.scan_file(bytes(data))
    assert ref_stats["corrupt"] == 1 and b"4" not in ref_index and ref_index[b"7"].stream(0) != b"old"
    with szg.SpanFile(path) as sf:
        info = sf.info()
        assert info["spans_corrupt"] == 1 and info["foreign_records"] == 2
        assert info["records"] == len(live) - 1
        assert 4 not in sf.ids().tolist()
        assert sf.record(7)[0] == live[7].tobytes()


def test_open_errors(tmp_path):
    with pytest.raises(szg.SzgError):
        szg.SpanFile(os.path.join(tmp_path, "missing.dat"))
    bad = os.path.join(tmp_path, "bad.dat")
    with open(bad, "wb") as f:
        f.write(b"NOPE" + b"\x00" * 60)
    with pytest.raises(szg.SzgError, match="invalid magic number"):  # spanfile.go:250-254
        szg.SpanFile(bad)
    zero = os.path.join(tmp_path, "zero.dat")
    with open(zero, "wb") as f:
        f.write(b"SPAN" + (0).to_bytes(4, "big") + b"\x00" * 30)
    with pytest.raises(szg.SzgError, match="length is 0"):           # spanfile.go:319, 352
        szg.SpanFile(zero)
    empty = os.path.join(tmp_path, "empty.dat")
    open(empty, "wb").close()
    with szg.SpanFile(empty) as sf:
        assert sf.info()["records"] == 0 and sf.info()["has_header"] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("bits,metric,dims,n", [(8, szg.COSINE, 96, 5000), (4, szg.EUCLIDEAN, 31, 3000), (64, szg.COSINE, 12, 2000)])
def test_load_span_file_into_the_mirror_and_search(tmp_path, bits, metric, dims, n):
    w, live, metas = _build_collection(n, dims, bits, metric, 70 + bits, updates=n // 10, removes=n // 20)
    path = os.path.join(tmp_path, "c.dat")
    with open(path, "wb") as f:
        f.write(w.tobytes())
    ids = np.array(sorted(live), dtype=np.uint64)
    codes = np.stack([live[int(i)] for i in ids])
    queries = o.synth_queries(3, 0, 4, dims)
    with szg.SpanFile(path) as sf, sf.open_index() as ix:
        assert ix.count() == len(live)
        gi, gd, gn, scanned = ix.search_topk(queries, 10)
        assert scanned == len(live)
        for qi, q in enumerate(queries):
            ri, rd, _ = o.search_exact(codes, ids, dims, bits, metric, q, k=10)
            if np.isnan(rd).any():
                continue
            assert_results_match(gi[qi, :gn[qi]], gd[qi, :gn[qi]], ri, rd, None, f"spanfile b{bits} q{qi}")
        # metadata of the winners comes from the mapping (the shim fills SearchResult.Metadata this way)
        assert sf.record(int(gi[0, 0]))[1] == metas[int(gi[0, 0])]
        # a mirror of the wrong shape is refused
        with szg.Index(dims + 1, bits, metric) as other:
            with pytest.raises(szg.SzgError):
                sf.load_into(other)


# ------------------------------------------------------------------ the reference's own span-file tests, restated end to end
def test_reference_checksum_verification(tmp_path):
    """spanfile_test.go:66-97 (TestChecksumVerification): WriteRecord("record1", {1: "Hello"}), flip byte offset+9 of the span,
    reading the record must fail on the checksum.  The library's reader meets such a span when it opens the file
    (scanFile, spanfile.go:313-329: the span is skipped and counted); the same file with a document id shows the
    record gone from szg_spanfile_record / szg_spanfile_ids, like ReadRecord's error."""
    for rid in (b"record1", b"1"):
        w = sfo.SpanFileWriter()
        w.write_record(rid, [(1, b"Hello")])
        clean = w.tobytes()
        path = os.path.join(tmp_path, "ok_%s.dat" % rid.decode())
        with open(path, "wb") as f:
            f.write(clean)
        with szg.SpanFile(path) as sf:
            assert sf.info()["spans_corrupt"] == 0 and sf.info()["spans_active"] == 2   # the "" span OpenFile writes + the record
            if rid == b"1":
                assert sf.record(1) == (b"Hello", None) and sf.ids().tolist() == [1]
        data = bytearray(clean)
        offset = w.index[rid]
        data[offset + 9] ^= 0xFF                                       # spanfile_test.go:84
        ref_index, ref_stats = sfo.scan_file(bytes(data))              # the restated verifyChecksum (spanfile.go:841-849) rejects it
        assert rid not in ref_index and ref_stats["corrupt"] == 1
        path = os.path.join(tmp_path, "bad_%s.dat" % rid.decode())
        with open(path, "wb") as f:
            f.write(data)
        with szg.SpanFile(path) as sf:
            info = sf.info()
            assert info["spans_corrupt"] == 1 and info["spans_active"] == 1 and info["records"] == 0
            if rid == b"1":
                assert sf.ids().size == 0
                with pytest.raises(KeyError):                          # ReadRecord: error (spanfile.go:513-519)
                    sf.record(1)


def test_reference_sequence_number_wraparound(tmp_path):
    """spanfile_test.go:117-134 (TestSequenceNumberWraparound): with sequenceNumber = 0xFFFFFFFF a WriteRecord succeeds and
    the counter wraps to 0.  On disk that record carries sequence 0xFFFFFFFF (a 5-byte 7-code); the reader must parse it,
    report next_sequence = 0 like scanFile's highest + 1 in uint32 (spanfile.go:355), and -- the reference's rule, kept bug
    for bug -- let it win over any later write of the same record, whose sequence number (0, 1, ...) is lower (337-341)."""
    w = sfo.SpanFileWriter()
    w.write_header("wrap", 0, 2, 8)
    w.seq = 0xFFFFFFFF
    w.add_document(5, b"\x01\x02", b'{"v": 1}')
    assert w.seq == 0                                                  # the reference's assertion
    path = os.path.join(tmp_path, "wrap.dat")
    with open(path, "wb") as f:
        f.write(w.tobytes())
    with szg.SpanFile(path) as sf:
        info = sf.info()
        assert info["next_sequence"] == 0 and info["records"] == 1 and info["spans_corrupt"] == 0
        assert sf.record(5) == (b"\x01\x02", b'{"v": 1}')
    # a second version written after the wrap (sequence 0) next to the first one (e.g. after a crash before the old span was
    # freed): both are valid spans; scanFile keeps the HIGHER sequence number, i.e. the older write
    tail = sfo.serialize_span(0, b"5", [(0, b'{"v": 2}'), (1, b"\x09\x09")])
    tail += sfo.crc32_ieee(bytes(tail)).to_bytes(4, "big")
    data = bytearray(w.tobytes())
    off = 0
    while off + 8 <= len(data) and int.from_bytes(data[off:off + 4], "big") != 0:
        off += int.from_bytes(data[off + 4:off + 8], "big")
    data[off:off + len(tail)] = tail
    ref_index, _ = sfo.scan_file(bytes(data))
    assert ref_index[b"5"].stream(1) == b"\x01\x02"
    path2 = os.path.join(tmp_path, "wrap2.dat")
    with open(path2, "wb") as f:
        f.write(data)
    with szg.SpanFile(path2) as sf:
        assert sf.record(5) == (b"\x01\x02", b'{"v": 1}') and sf.info()["spans_active"] >= 3
