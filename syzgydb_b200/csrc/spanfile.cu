// spanfile.cu -- direct span-file reader: a collection's .dat file -> the GPU mirror, without going
// through the reference's per-record Go loop.  Host code only.
//
// Restates, for reading: the span grammar (spanfile.go:1-22), scanFile (282-357: corrupt spans are
// skipped, a zero magic ends the file, the highest sequence number per record id wins and the first seen
// wins a tie), parseSpan (730-818), read7Code (627-636; the writer's thresholds are non-canonical, so no
// canonical-length assumption is made), verifyChecksum (841-849, CRC32-IEEE over the span minus its last
// 4 bytes), getStream (67-118: the FIRST stream with the wanted id), the header record "" whose stream 0 is
// the CollectionOptions JSON (collection.go:31-47, 241-252) and the reload loop of NewCollection
// (collection.go:298-311: only ids that strconv.ParseUint accepts are documents).
//
// What is different from the reference by design: the file is mapped read-only and never written; spans
// are verified and parsed by all host threads (phase 2) between a sequential walk of the span headers
// (phase 1) and a sequential, file-ordered merge (phase 3) that keeps the reference's tie rule; stream 1 of
// every live document is gathered into pinned staging in large batches and handed to szg_upsert.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cerrno>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/syzgy_b200.h"

namespace szg {
int set_error(int code, const char *msg);
void index_geometry(const szg_index *h, int *dim, int *quant, int *metric, uint32_t *rowbytes);
} // namespace szg

namespace {

constexpr uint32_t kActiveMagic = 0x5350414E; // 'SPAN'  spanfile.go:57
constexpr uint32_t kFreeMagic = 0x46524545;   // 'FREE'  spanfile.go:58
constexpr size_t kMinSpanLength = 15;         // spanfile.go:61

int failf(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    return szg::set_error(code, buf);
}

inline uint32_t be32(const uint8_t *p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }

// CRC32-IEEE (hash/crc32 ChecksumIEEE: reflected 0xEDB88320, init/xorout 0xFFFFFFFF), slicing-by-8
struct CrcTables {
    uint32_t t[8][256];
    CrcTables() {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u)));
            t[0][i] = c;
        }
        for (uint32_t i = 0; i < 256; ++i)
            for (int j = 1; j < 8; ++j) t[j][i] = (t[j - 1][i] >> 8) ^ t[0][t[j - 1][i] & 0xFF];
    }
};
const CrcTables &crc_tables() {
    static const CrcTables T;
    return T;
}
uint32_t crc32_ieee(const uint8_t *p, size_t n) {
    const CrcTables &T = crc_tables();
    uint32_t c = 0xFFFFFFFFu;
    while (n && (reinterpret_cast<uintptr_t>(p) & 7)) { c = (c >> 8) ^ T.t[0][(c ^ *p++) & 0xFF]; --n; }
    while (n >= 8) {
        uint64_t w;
        memcpy(&w, p, 8);
        w ^= c;
        c = T.t[7][w & 0xFF] ^ T.t[6][(w >> 8) & 0xFF] ^ T.t[5][(w >> 16) & 0xFF] ^ T.t[4][(w >> 24) & 0xFF] ^
            T.t[3][(w >> 32) & 0xFF] ^ T.t[2][(w >> 40) & 0xFF] ^ T.t[1][(w >> 48) & 0xFF] ^ T.t[0][(w >> 56) & 0xFF];
        p += 8;
        n -= 8;
    }
    while (n--) c = (c >> 8) ^ T.t[0][(c ^ *p++) & 0xFF];
    return ~c;
}

// read7Code (spanfile.go:627-636): big-endian base 128, high bit = more
bool read7(const uint8_t *b, size_t len, size_t *at, uint64_t *out) {
    uint64_t r = 0;
    for (size_t i = *at; i < len; ++i) {
        const uint64_t d = b[i];
        r = (r << 7) | (d & 0x7f);
        if (!(d & 0x80)) { *at = i + 1; *out = r; return true; }
    }
    return false;
}

struct SpanRef { uint64_t off; uint32_t len; };

struct Parsed {
    bool ok = false;
    uint32_t seq = 0;
    uint64_t id_off = 0;   // record id bytes (absolute offsets into the mapping)
    uint32_t id_len = 0;
    uint64_t vec_off = 0, meta_off = 0;
    uint32_t vec_len = 0, meta_len = 0;
    bool has_vec = false, has_meta = false;
};

// parseSpan (spanfile.go:730-818) + verifyChecksum; records the first stream 0 and the first stream 1 (getStream)
void parse_span(const uint8_t *map, const SpanRef &s, Parsed *out) {
    const uint8_t *d = map + s.off;
    const size_t len = s.len;
    if (len < kMinSpanLength) return;
    if (crc32_ieee(d, len - 4) != be32(d + len - 4)) return;
    size_t at = 8;
    uint64_t seq, idlen;
    if (!read7(d, len, &at, &seq) || !read7(d, len, &at, &idlen)) return;
    if (idlen > len || at + idlen >= len) return; // data[at : at+idlength] and the stream count byte must exist
    out->seq = (uint32_t)seq;
    out->id_off = s.off + at;
    out->id_len = (uint32_t)idlen;
    at += idlen;
    const unsigned nstreams = d[at++];
    for (unsigned i = 0; i < nstreams; ++i) {
        if (at >= len) return; // "data too short to contain all streams"
        const uint8_t sid = d[at++];
        uint64_t slen;
        if (!read7(d, len, &at, &slen)) return;
        if (slen > len || at + slen > len) return; // "data too short for stream data"
        if (sid == 0 && !out->has_meta) { out->has_meta = true; out->meta_off = s.off + at; out->meta_len = (uint32_t)slen; }
        if (sid == 1 && !out->has_vec) { out->has_vec = true; out->vec_off = s.off + at; out->vec_len = (uint32_t)slen; }
        at += slen;
    }
    if (at + 4 > len) return; // "data too short for checksum"
    out->ok = true;
}

// strconv.ParseUint(s, 10, 64) restricted to what the reference itself writes (fmt.Sprintf("%d", id)):
// canonical decimal, no sign, no leading zeros (a non-canonical spelling could never be found again by
// getDocument, which looks the canonical string up: collection.go:470-475)
bool parse_doc_id(const uint8_t *p, uint32_t n, uint64_t *out) {
    if (n == 0 || n > 20) return false;
    if (n > 1 && p[0] == '0') return false;
    unsigned __int128 v = 0;
    for (uint32_t i = 0; i < n; ++i) {
        if (p[i] < '0' || p[i] > '9') return false;
        v = v * 10 + (p[i] - '0');
    }
    if (v > (unsigned __int128)UINT64_MAX) return false;
    *out = (uint64_t)v;
    return true;
}

// the four CollectionOptions fields of the header JSON (collection.go:31-47); encoding/json output is compact,
// but any whitespace after the colon is tolerated
bool json_int(const std::string &js, const char *key, long long *out) {
    const std::string k = std::string("\"") + key + "\"";
    size_t p = js.find(k);
    if (p == std::string::npos) return false;
    p = js.find(':', p + k.size());
    if (p == std::string::npos) return false;
    ++p;
    while (p < js.size() && (js[p] == ' ' || js[p] == '\t' || js[p] == '\n' || js[p] == '\r')) ++p;
    char *end = nullptr;
    const long long v = strtoll(js.c_str() + p, &end, 10);
    if (end == js.c_str() + p) return false;
    *out = v;
    return true;
}
bool json_string(const std::string &js, const char *key, std::string *out) {
    const std::string k = std::string("\"") + key + "\"";
    size_t p = js.find(k);
    if (p == std::string::npos) return false;
    p = js.find(':', p + k.size());
    if (p == std::string::npos) return false;
    p = js.find('"', p);
    if (p == std::string::npos) return false;
    std::string r;
    for (++p; p < js.size() && js[p] != '"'; ++p) {
        if (js[p] == '\\' && p + 1 < js.size()) ++p;
        r.push_back(js[p]);
    }
    *out = r;
    return true;
}

unsigned worker_count() {
    unsigned n = std::thread::hardware_concurrency();
    if (const char *e = getenv("SZG_HOST_THREADS")) n = (unsigned)atoi(e);
    return std::max(1u, std::min(n, 32u));
}

template <typename F>
void parallel_for(size_t n, size_t grain, F f) {
    const unsigned nt = (unsigned)std::min<size_t>(worker_count(), (n + grain - 1) / std::max<size_t>(grain, 1));
    if (nt <= 1) { f(0, n); return; }
    std::atomic<size_t> next{0};
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
        th.emplace_back([&] {
            for (;;) {
                const size_t b = next.fetch_add(grain);
                if (b >= n) return;
                f(b, std::min(n, b + grain));
            }
        });
    for (auto &x : th) x.join();
}

struct Doc {
    uint64_t id;
    uint64_t vec_off, meta_off;
    uint32_t vec_len, meta_len;
    uint32_t seq;
    uint8_t has_vec, has_meta;
};

// sort.Strings over the decimal spellings (spanfile.go:540-560) without building strings: left-align the digits
// (id * 10^(20 - digits) fits 128 bits); equal padded values differ only by trailing zeros, the shorter sorts first
struct LexKey {
    uint64_t hi, lo;
    uint32_t digits, index;
};
LexKey lex_key(uint64_t id, uint32_t index) {
    uint32_t digits = 1;
    for (uint64_t v = id; v >= 10; v /= 10) ++digits;
    unsigned __int128 p = id;
    for (uint32_t i = digits; i < 20; ++i) p *= 10;
    return {(uint64_t)(p >> 64), (uint64_t)p, digits, index};
}

} // namespace

struct szg_spanfile {
    int fd = -1;
    const uint8_t *map = nullptr;
    size_t size = 0;
    std::vector<Doc> by_num;     // live documents in ascending numeric id order (binary-searched by szg_spanfile_record)
    std::vector<uint32_t> lex;   // indices into by_num in lexicographic decimal-id order (IterateSortedRecords)
    szg_spanfile_info info;
};

extern "C" {

int szg_spanfile_open(const char *path, szg_spanfile **out) {
    if (!path || !out) return failf(SZG_EINVAL, "null argument");
    *out = nullptr;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return failf(SZG_ENOTFOUND, "cannot open %s: %s", path, strerror(errno));
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return failf(SZG_EINTERNAL, "fstat(%s): %s", path, strerror(errno)); }
    szg_spanfile *sf = new (std::nothrow) szg_spanfile();
    if (!sf) { close(fd); return failf(SZG_ENOMEM, "out of memory"); }
    memset(&sf->info, 0, sizeof sf->info);
    sf->fd = fd;
    sf->size = (size_t)st.st_size;
    if (sf->size) {
        void *m = mmap(nullptr, sf->size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
        if (m == MAP_FAILED) { close(fd); delete sf; return failf(SZG_EINTERNAL, "mmap(%s): %s", path, strerror(errno)); }
        sf->map = static_cast<const uint8_t *>(m);
        madvise(m, sf->size, MADV_SEQUENTIAL);
        // OpenFile (spanfile.go:241-256): a non-empty file must start with a span
        if (sf->size >= 4) {
            const uint32_t magic = be32(sf->map);
            if (magic != kActiveMagic && magic != kFreeMagic) {
                szg_spanfile_close(sf);
                return failf(SZG_EINVAL, "invalid magic number: %x", magic);
            }
        }
    }
    const bool timing = getenv("SZG_SPANFILE_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    const auto t0 = now();
    const uint8_t *map = sf->map;
    const size_t size = sf->size;
    // ---- phase 1: walk the span headers (scanFile's loop without the parsing)
    std::vector<SpanRef> active;
    size_t offset = 0;
    uint64_t nfree = 0, free_bytes = 0;
    while (offset < size) {
        if (offset + kMinSpanLength > size) break;
        const uint32_t magic = be32(map + offset);
        if (magic == 0) { free_bytes += size - offset; offset = size; break; } // the rest of the file is free space
        const uint32_t length = be32(map + offset + 4);
        if (offset + (size_t)length > size) break;
        if (length == 0) { szg_spanfile_close(sf); return failf(SZG_EINVAL, "length is 0; can't continue (offset %zu)", offset); }
        if (magic == kActiveMagic) active.push_back({(uint64_t)offset, length});
        else if (magic == kFreeMagic) { ++nfree; free_bytes += length; }
        offset += length;
    }
    free_bytes += size - offset;
    const auto t1 = now();
    // ---- phase 2: checksum + parse, all host threads
    std::vector<Parsed> parsed(active.size());
    parallel_for(active.size(), 4096, [&](size_t b, size_t e) {
        for (size_t i = b; i < e; ++i) parse_span(map, active[i], &parsed[i]);
    });
    const auto t2 = now();
    // ---- phase 3: file-ordered merge: highest sequence number per id, first seen wins a tie (spanfile.go:337-341)
    // sort-based (no hash maps): candidates ordered by (id, file position); per id the highest sequence number wins
    // and, walking in file order, the first seen wins a tie
    uint64_t ncorrupt = 0, nforeign = 0;
    uint32_t highest = 0;
    long header = -1;
    struct Cand { uint64_t id; uint32_t span; };
    std::vector<Cand> cands;
    cands.reserve(parsed.size());
    std::vector<std::string> foreign; // ids that are not documents: only counted (distinct spellings)
    for (size_t i = 0; i < parsed.size(); ++i) {
        const Parsed &p = parsed[i];
        if (!p.ok) { ++ncorrupt; continue; }
        highest = std::max(highest, p.seq);
        if (p.id_len == 0) {
            if (header < 0 || p.seq > parsed[(size_t)header].seq) header = (long)i;
            continue;
        }
        uint64_t id;
        if (!parse_doc_id(map + p.id_off, p.id_len, &id)) {
            foreign.emplace_back(reinterpret_cast<const char *>(map + p.id_off), p.id_len);
            continue;
        }
        cands.push_back({id, (uint32_t)i});
    }
    std::sort(foreign.begin(), foreign.end());
    nforeign = (uint64_t)(std::unique(foreign.begin(), foreign.end()) - foreign.begin());
    bool ordered = true; // a freshly written collection is already in ascending id order
    for (size_t i = 1; i < cands.size() && ordered; ++i) ordered = cands[i - 1].id < cands[i].id;
    if (!ordered)
        std::sort(cands.begin(), cands.end(), [](const Cand &a, const Cand &b) { return a.id != b.id ? a.id < b.id : a.span < b.span; });
    sf->by_num.clear();
    sf->by_num.reserve(cands.size());
    for (size_t i = 0; i < cands.size();) {
        size_t j = i, bestj = i;
        for (; j < cands.size() && cands[j].id == cands[i].id; ++j)
            if (parsed[cands[j].span].seq > parsed[cands[bestj].span].seq) bestj = j;
        const Parsed &p = parsed[cands[bestj].span];
        Doc d;
        d.id = cands[i].id; d.vec_off = p.vec_off; d.meta_off = p.meta_off; d.vec_len = p.vec_len; d.meta_len = p.meta_len;
        d.seq = p.seq; d.has_vec = p.has_vec; d.has_meta = p.has_meta;
        sf->by_num.push_back(d);
        i = j;
    }
    // IterateSortedRecords order: indices into by_num sorted by the decimal spelling
    {
        std::vector<LexKey> keys(sf->by_num.size());
        parallel_for(keys.size(), 65536, [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; ++i) keys[i] = lex_key(sf->by_num[i].id, (uint32_t)i);
        });
        std::sort(keys.begin(), keys.end(), [](const LexKey &a, const LexKey &b) {
            return a.hi != b.hi ? a.hi < b.hi : a.lo != b.lo ? a.lo < b.lo : a.digits < b.digits;
        });
        sf->lex.resize(keys.size());
        for (size_t i = 0; i < keys.size(); ++i) sf->lex[i] = keys[i].index;
    }
    if (timing)
        fprintf(stderr, "szg_spanfile_open: walk %.1f ms, crc+parse %.1f ms (%u threads), merge+sort %.1f ms\n", ms(t0, t1), ms(t1, t2),
                worker_count(), ms(t2, now()));
    szg_spanfile_info &in = sf->info;
    in.file_bytes = size;
    in.records = sf->by_num.size();
    in.spans_active = active.size() - ncorrupt;
    in.spans_free = nfree;
    in.spans_corrupt = ncorrupt;
    in.free_bytes = free_bytes;
    in.foreign_records = nforeign;
    in.next_sequence = highest + 1; // scanFile: db.sequenceNumber = highestSeqNum + 1
    in.has_header = 0;
    in.distance_method = -1; in.dimension_count = 0; in.quantization = 0;
    if (header >= 0 && parsed[(size_t)header].has_meta) {
        const Parsed &h = parsed[(size_t)header];
        const std::string js(reinterpret_cast<const char *>(map + h.meta_off), h.meta_len);
        long long dm = 0, dc = 0, q = 0;
        std::string name;
        if (json_int(js, "distance_method", &dm) && json_int(js, "dimension_count", &dc) && json_int(js, "quantization", &q)) {
            in.has_header = 1;
            in.distance_method = (int32_t)dm;
            in.dimension_count = (int32_t)dc;
            in.quantization = (int32_t)(q == 0 ? 64 : q); // collection.go:254-256
            if (json_string(js, "name", &name)) snprintf(in.name, sizeof in.name, "%s", name.c_str());
        }
    }
    *out = sf;
    return SZG_OK;
}

int szg_spanfile_close(szg_spanfile *sf) {
    if (!sf) return SZG_OK;
    if (sf->map) munmap(const_cast<uint8_t *>(sf->map), sf->size);
    if (sf->fd >= 0) close(sf->fd);
    delete sf;
    return SZG_OK;
}

int szg_spanfile_get_info(const szg_spanfile *sf, szg_spanfile_info *out) {
    if (!sf || !out) return failf(SZG_EINVAL, "null argument");
    *out = sf->info;
    return SZG_OK;
}

int szg_spanfile_ids(const szg_spanfile *sf, uint64_t *out_ids, uint64_t cap, uint64_t *n) {
    if (!sf || !n) return failf(SZG_EINVAL, "null argument");
    *n = sf->lex.size();
    if (out_ids)
        for (uint64_t i = 0; i < std::min<uint64_t>(cap, sf->lex.size()); ++i) out_ids[i] = sf->by_num[sf->lex[i]].id;
    return SZG_OK;
}

int szg_spanfile_record(const szg_spanfile *sf, uint64_t id, const uint8_t **vector, uint64_t *vector_len, const uint8_t **metadata,
                        uint64_t *metadata_len) {
    if (!sf) return failf(SZG_EINVAL, "null argument");
    auto it = std::lower_bound(sf->by_num.begin(), sf->by_num.end(), id, [](const Doc &a, uint64_t v) { return a.id < v; });
    if (it == sf->by_num.end() || it->id != id) return failf(SZG_ENOTFOUND, "record not found"); // spanfile.go:516
    const Doc &d = *it;
    if (vector) *vector = d.has_vec ? sf->map + d.vec_off : nullptr;
    if (vector_len) *vector_len = d.has_vec ? d.vec_len : 0;
    if (metadata) *metadata = d.has_meta ? sf->map + d.meta_off : nullptr;
    if (metadata_len) *metadata_len = d.has_meta ? d.meta_len : 0;
    return SZG_OK;
}

int szg_spanfile_load(const szg_spanfile *sf, szg_index *h, uint64_t *loaded) {
    if (!sf || !h) return failf(SZG_EINVAL, "null argument");
    int dim, quant, metric;
    uint32_t rowbytes;
    szg::index_geometry(h, &dim, &quant, &metric, &rowbytes);
    if (sf->info.has_header && (sf->info.dimension_count != dim || sf->info.quantization != quant))
        return failf(SZG_EINVAL, "collection file is %d x %d-bit, the mirror is %d x %d-bit", sf->info.dimension_count,
                     sf->info.quantization, dim, quant);
    const size_t n = sf->by_num.size();
    // decodeDocument panics when stream 1 is missing and decodeVector reads dimension-many values out of it
    // (collection.go:752-757, 768-794): a record whose stream 1 is not exactly one row is an error here
    for (size_t i = 0; i < n; ++i)
        if (!sf->by_num[i].has_vec || sf->by_num[i].vec_len != rowbytes)
            return failf(SZG_EINVAL, "record %llu: vector stream has %u bytes, a row has %u", (unsigned long long)sf->by_num[i].id,
                         sf->by_num[i].has_vec ? sf->by_num[i].vec_len : 0u, rowbytes);
    uint64_t have = 0;
    int rc = szg_count(h, &have);
    if (rc || (rc = szg_reserve(h, have + n))) return rc;
    const size_t batch = std::max<size_t>(1, std::min<size_t>(n, ((size_t)64 << 20) / rowbytes));
    std::vector<uint8_t> stage(batch * rowbytes);
    std::vector<uint64_t> ids(batch);
    for (size_t b = 0; b < n; b += batch) {
        const size_t m = std::min(batch, n - b);
        parallel_for(m, 8192, [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; ++i) {
                const Doc &d = sf->by_num[sf->lex[b + i]]; // upload in scan order: slot order = lexicographic id order
                ids[i] = d.id;
                memcpy(stage.data() + i * rowbytes, sf->map + d.vec_off, rowbytes);
            }
        });
        if ((rc = szg_upsert(h, ids.data(), stage.data(), m))) return rc;
    }
    if (loaded) *loaded = n;
    return SZG_OK;
}

} // extern "C"
