"""oracle/spanfile.py -- TEST INFRASTRUCTURE: a pure-Python restatement of the reference's span-file
format (spanfile.go, freemap.go), used to generate .dat fixtures and to check the C++ reader
(syzgydb_b200/csrc/spanfile.cpp).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
import it.

Parity pinning: the reference ships no .dat files and Go cannot run here, so this restatement is pinned only
by the cited lines ("parity unpinned by the reference's own fixtures"); the writer reproduces the reference's
non-canonical 7-code thresholds (spanfile.go:568-625), the padding rule (spanfile.go:426-434), the free-span
markers (446-452), the first-fit free map (freemap.go:94-121) and the growth policy (477-492) so that the
fixtures have the byte layout a Go-written collection has.
"""
from __future__ import annotations

import json
import zlib

ACTIVE_MAGIC = 0x5350414E  # 'SPAN'  spanfile.go:57
FREE_MAGIC = 0x46524545    # 'FREE'  spanfile.go:58
MIN_SPAN_LENGTH = 15       # spanfile.go:61


# ------------------------------------------------------------------ 7-code (spanfile.go:568-661)
_THRESHOLDS = [0x7f, 0x3fff, 0x1fffff, 0xfffffff, 0x7ffffffff, 0x3ffffffffff, 0x1ffffffffffff, 0xffffffffffffff]


def length_of_7code(n: int) -> int:
    """lengthOf7Code (spanfile.go:638-661): thresholds are off by one (127 takes 2 bytes)."""
    for i, t in enumerate(_THRESHOLDS):
        if n < t:
            return i + 1
    return 9 if n < 0x7fffffffffffffff else 10


def write_7code(n: int) -> bytes:
    """write7Code (spanfile.go:568-625): big-endian base 128, high bit = more; at most 9 groups are ever written."""
    groups = 9
    for i, t in enumerate(_THRESHOLDS):
        if n < t:
            groups = i + 1
            break
    out = bytearray()
    for g in range(groups - 1, -1, -1):
        b = (n >> (7 * g)) & 0x7f
        out.append(b | (0x80 if g else 0))
    return bytes(out)


def read_7code(buf: bytes, at: int):
    """read7Code (spanfile.go:627-636) -> (value, new offset); raises on a truncated code."""
    result = 0
    while at < len(buf):
        d = buf[at]
        result = ((result << 7) | (d & 0x7f)) & 0xFFFFFFFFFFFFFFFF
        at += 1
        if not d & 0x80:
            return result, at
    raise ValueError("buffer too short to read unsigned value")


def crc32_ieee(data: bytes) -> int:
    return zlib.crc32(data) & 0xFFFFFFFF  # hash/crc32 ChecksumIEEE (spanfile.go:836-838)


# ------------------------------------------------------------------ serializeSpan (spanfile.go:679-728)
def serialize_span(seq: int, record_id: bytes, streams) -> bytearray:
    """Everything but the checksum; the length field already counts the 4 checksum bytes."""
    length = 4 + 4 + length_of_7code(seq) + length_of_7code(len(record_id)) + len(record_id) + 1 + 4
    for _, data in streams:
        length += 1 + length_of_7code(len(data)) + len(data)
    buf = bytearray()
    buf += ACTIVE_MAGIC.to_bytes(4, "big")
    buf += (length & 0xFFFFFFFF).to_bytes(4, "big")
    buf += write_7code(seq)
    buf += write_7code(len(record_id))
    buf += record_id
    buf.append(len(streams) & 0xFF)
    for sid, data in streams:
        buf.append(sid)
        buf += write_7code(len(data))
        buf += data
    return buf


# ------------------------------------------------------------------ freeMap (freemap.go)
class FreeMap:
    def __init__(self):
        self.spaces = []  # [start, length], sorted by start after every mark_free

    def mark_free(self, start: int, length: int):  # freemap.go:61-91
        if length <= 0:
            return
        self.spaces.append([start, length])
        self.spaces.sort(key=lambda s: s[0])
        merged = []
        for s in self.spaces:
            if not merged or merged[-1][0] + merged[-1][1] < s[0]:
                merged.append(list(s))
            else:
                merged[-1][1] = s[0] + s[1] - merged[-1][0]
        self.spaces = merged

    def mark_used(self, start: int, length: int):  # freemap.go:12-49
        if length <= 0:
            return
        for i, s in enumerate(self.spaces):
            if s[0] <= start and start + length <= s[0] + s[1]:
                if start == s[0]:
                    s[0] += length
                    s[1] -= length
                elif start + length == s[0] + s[1]:
                    s[1] -= length
                else:
                    self.spaces.append([start + length, s[0] + s[1] - (start + length)])
                    s[1] = start - s[0]
                if s[1] == 0:
                    del self.spaces[i]
                break

    def get_free_range(self, length: int):  # freemap.go:94-121, first fit
        for i, s in enumerate(self.spaces):
            if s[1] >= length:
                start, had = s[0], s[1]
                s[0] += length
                s[1] -= length
                if s[1] == 0:
                    del self.spaces[i]
                return start, had - length
        return None


# ------------------------------------------------------------------ reader: scanFile (spanfile.go:282-357) + parseSpan (730-818)
class ParsedSpan:
    __slots__ = ("offset", "length", "seq", "record_id", "streams")

    def __init__(self, offset, length, seq, record_id, streams):
        self.offset, self.length, self.seq, self.record_id, self.streams = offset, length, seq, record_id, streams

    def stream(self, sid: int):
        """getStream (spanfile.go:67-118): the first stream with that id."""
        for i, d in self.streams:
            if i == sid:
                return d
        return None


def parse_span(data: bytes, offset: int, length: int) -> ParsedSpan:
    span = data[offset:offset + length]
    if len(span) < MIN_SPAN_LENGTH:
        raise ValueError("data too short to be a valid span")
    if crc32_ieee(span[:-4]) != int.from_bytes(span[-4:], "big"):
        raise ValueError("checksum failed")
    at = 8
    seq, at = read_7code(span, at)
    seq &= 0xFFFFFFFF
    idlen, at = read_7code(span, at)
    rid = bytes(span[at:at + idlen])
    at += idlen
    nstreams = span[at]
    at += 1
    streams = []
    for _ in range(nstreams):
        if at >= len(span):
            raise ValueError("data too short to contain all streams")
        sid = span[at]
        at += 1
        slen, at = read_7code(span, at)
        if at + slen > len(span):
            raise ValueError("data too short for stream data")
        streams.append((sid, bytes(span[at:at + slen])))
        at += slen
    if at + 4 > len(span):
        raise ValueError("data too short for checksum")
    return ParsedSpan(offset, length, seq, rid, streams)


def scan_file(data: bytes):
    """scanFile: returns (index: record id bytes -> ParsedSpan of the highest sequence number (first seen wins a
    tie), stats dict).  Corrupt spans are skipped, a zero magic ends the file, a zero length is an error."""
    offset, size = 0, len(data)
    index, stats = {}, {"active": 0, "free": 0, "corrupt": 0, "highest_seq": 0}
    while offset < size:
        if offset + MIN_SPAN_LENGTH > size:
            break
        magic = int.from_bytes(data[offset:offset + 4], "big")
        if magic == 0:
            offset = size
            break
        length = int.from_bytes(data[offset + 4:offset + 8], "big")
        if offset + length > size:
            break
        if magic == ACTIVE_MAGIC:
            try:
                sp = parse_span(data, offset, length)
            except (ValueError, IndexError):
                stats["corrupt"] += 1
                if length == 0:
                    raise ValueError("length is 0; can't continue")
                offset += length
                continue
            stats["active"] += 1
            stats["highest_seq"] = max(stats["highest_seq"], sp.seq)
            old = index.get(sp.record_id)
            if old is None or sp.seq > old.seq:
                index[sp.record_id] = sp
        elif magic == FREE_MAGIC:
            stats["free"] += 1
        if length == 0:
            raise ValueError("length is 0; can't continue")
        offset += length
    return index, stats


def live_records(data: bytes):
    """What NewCollection's reload loop sees (collection.go:298-311): numeric decimal ids only, in
    IterateSortedRecords order (sort.Strings).  -> (header options dict, [(id, vector bytes, metadata bytes)])"""
    index, _ = scan_file(data)
    header = index.get(b"")
    opts = json.loads(header.stream(0)) if header is not None and header.stream(0) is not None else None
    out = []
    for rid in sorted(k for k in index if k != b""):
        try:
            s = rid.decode("ascii")
            if not s.isdigit():  # strconv.ParseUint(recordID, 10, 64): no sign, no spaces
                continue
            v = int(s)
            if v >= 1 << 64:
                continue
        except UnicodeDecodeError:
            continue
        sp = index[rid]
        out.append((v, sp.stream(1), sp.stream(0)))
    return opts, out


# ------------------------------------------------------------------ writer: OpenFile / WriteRecord / RemoveRecord
class SpanFileWriter:
    """In-memory image of a span file written the way the reference writes it."""

    def __init__(self):
        self.data = bytearray()
        self.index = {}
        self.free = FreeMap()
        # OpenFile on an empty file (spanfile.go:216-238): a minimal span with id "" and no streams, then scanFile
        first = serialize_span(0, b"", [])
        first += crc32_ieee(bytes(first)).to_bytes(4, "big")
        self.data += first
        self.index[b""] = 0
        self.seq = 1  # scanFile: highest + 1

    def _span_length(self, offset: int) -> int:
        return int.from_bytes(self.data[offset + 4:offset + 8], "big")

    def _allocate(self, size: int):  # allocateSpan (spanfile.go:477-497)
        got = self.free.get_free_range(size)
        if got is not None:
            return got
        cur = len(self.data)
        expand = max(4096, size, int(cur * 0.05))
        self.data += bytes(expand)
        self.free.mark_free(cur + size, expand - size)
        return cur, expand - size

    def write_record(self, record_id: bytes, streams):  # WriteRecord (spanfile.go:398-475)
        seq = self.seq
        self.seq = (self.seq + 1) & 0xFFFFFFFF
        buf = serialize_span(seq, record_id, streams)
        offset, remaining = self._allocate(len(buf) + 4)
        if 0 < remaining < MIN_SPAN_LENGTH:
            self.free.mark_used(offset + len(buf) + 4, remaining)
            buf += bytes(remaining)
            buf[4:8] = (len(buf) + 4).to_bytes(4, "big")
        buf += crc32_ieee(bytes(buf)).to_bytes(4, "big")
        if remaining >= MIN_SPAN_LENGTH:
            buf += FREE_MAGIC.to_bytes(4, "big") + (remaining & 0xFFFFFFFF).to_bytes(4, "big")
        self.data[offset:offset + len(buf)] = buf
        old = self.index.get(record_id)
        if old is not None:
            ln = self._span_length(old)
            self.data[old:old + 4] = FREE_MAGIC.to_bytes(4, "big")
            self.free.mark_free(old, ln)
        self.index[record_id] = offset

    def remove_record(self, record_id: bytes):  # RemoveRecord (spanfile.go:365-396)
        offset = self.index.pop(record_id)
        ln = self._span_length(offset)
        self.data[offset:offset + 4] = FREE_MAGIC.to_bytes(4, "big")
        self.free.mark_free(offset, ln)

    # -- the collection layer above it (collection.go:258-271, 446-453)
    def write_header(self, name: str, distance_method: int, dims: int, quant: int):
        opts = {"name": name, "distance_method": distance_method, "dimension_count": dims, "quantization": quant}
        self.write_record(b"", [(0, json.dumps(opts, separators=(",", ":")).encode())])

    def add_document(self, doc_id: int, vector_bytes: bytes, metadata: bytes = b""):
        self.write_record(str(int(doc_id)).encode(), [(0, bytes(metadata)), (1, bytes(vector_bytes))])

    def remove_document(self, doc_id: int):
        self.remove_record(str(int(doc_id)).encode())

    def tobytes(self) -> bytes:
        return bytes(self.data)
