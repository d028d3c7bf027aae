/*
 * syzgy_oracle.c -- CPU restatement of the search hot path of smhanov/syzgydb.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under syzgydb_b200/ may include, link or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, as the checker / the timed CPU arm.
 *
 * PARITY PINNING: the reference is Go and no Go toolchain exists in this image, so
 * the reference itself cannot be run here.  The only numeric known-answer the
 * reference's own tests hold for this path is
 *   euclideanDistance([1,2,3],[4,5,6]) == 5.196152422706632  (collection_test.go:12-21)
 * plus behavioural properties (collection_test.go:283-382, 549-612).  Those are
 * checked in tests/test_oracle.py.  Everything else (codec values, angular
 * distance values, tie order, NaN handling, LSH replay) is "parity unpinned": it is
 * pinned only by this line-by-line restatement of the cited Go lines.  Go's math.Acos
 * (standard library, not under /root/reference) is restated from its published Cephes
 * algorithm below and cross-checked against libm (tests/test_oracle.py).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (see oracle/Makefile).
 * -ffp-contract=off mirrors Go/amd64, which never fuses x*y+z.
 *
 * Every function cites the reference file:line it restates (paths relative to the
 * reference repository root).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_EUCLIDEAN 0 /* collection.go:186-189 */
#define ORC_COSINE 1

/* signals, collection.go:19-24 */
#define ORC_STOP_SEARCH 0
#define ORC_POINT_ACCEPTED 1
#define ORC_POINT_CHECKED 2
#define ORC_POINT_IGNORED 3

/* ------------------------------------------------------------------ codec */

/* quantization.go:5-23.  math.Round == C round (half away from zero). */
uint64_t orc_quantize(double value, int bits) {
    if (bits == 32) {
        float f = (float)value;
        uint32_t u;
        memcpy(&u, &f, 4);
        return (uint64_t)u;
    }
    if (bits == 64) {
        uint64_t u;
        memcpy(&u, &value, 8);
        return u;
    }
    if (value < -1) value = -1;
    else if (value > 1) value = 1;
    int64_t maxInt = ((int64_t)1 << bits) - 1;
    double q = (value + 1) / 2 * (double)maxInt;
    return (uint64_t)round(q);
}

/* quantization.go:25-36 */
double orc_dequantize(uint64_t value, int bits) {
    if (bits == 32) {
        uint32_t u = (uint32_t)value;
        float f;
        memcpy(&f, &u, 4);
        return (double)f;
    }
    if (bits == 64) {
        double d;
        memcpy(&d, &value, 8);
        return d;
    }
    int64_t maxInt = ((int64_t)1 << bits) - 1;
    return ((double)value / (double)maxInt) * 2 - 1;
}

/* collection.go:796-811; returns -1 where the reference panics */
int64_t orc_vector_size(int bits, int64_t dims) {
    switch (bits) {
    case 4: return (dims + 1) / 2;
    case 8: return dims;
    case 16: return dims * 2;
    case 32: return dims * 4;
    case 64: return dims * 8;
    }
    return -1;
}

/* collection.go:713-744 (4-bit: even index -> high nibble; 16/32/64 big-endian) */
void orc_encode(const double *vec, int64_t dims, int bits, uint8_t *data) {
    int64_t size = orc_vector_size(bits, dims);
    memset(data, 0, (size_t)size);
    for (int64_t i = 0; i < dims; i++) {
        uint64_t q = orc_quantize(vec[i], bits);
        switch (bits) {
        case 4:
            if (i % 2 == 0) data[i / 2] = (uint8_t)(q << 4);
            else data[i / 2] |= (uint8_t)(q & 0x0F);
            break;
        case 8: data[i] = (uint8_t)q; break;
        case 16:
            data[i * 2] = (uint8_t)(q >> 8);
            data[i * 2 + 1] = (uint8_t)q;
            break;
        case 32:
            for (int b = 0; b < 4; b++) data[i * 4 + b] = (uint8_t)(q >> (24 - 8 * b));
            break;
        case 64:
            for (int b = 0; b < 8; b++) data[i * 8 + b] = (uint8_t)(q >> (56 - 8 * b));
            break;
        }
    }
}

/* collection.go:768-794 */
void orc_decode(const uint8_t *data, int64_t dims, int bits, double *vec) {
    for (int64_t i = 0; i < dims; i++) {
        uint64_t q = 0;
        switch (bits) {
        case 4:
            if (i % 2 == 0) q = (uint64_t)(data[i / 2] >> 4);
            else q = (uint64_t)(data[i / 2] & 0x0F);
            break;
        case 8: q = data[i]; break;
        case 16: q = ((uint64_t)data[i * 2] << 8) | data[i * 2 + 1]; break;
        case 32:
            for (int b = 0; b < 4; b++) q = (q << 8) | data[i * 4 + b];
            break;
        case 64:
            for (int b = 0; b < 8; b++) q = (q << 8) | data[i * 8 + b];
            break;
        }
        vec[i] = orc_dequantize(q, bits);
    }
}

/* -------------------------------------------------------------- distances */

/* collection.go:812-819 */
double orc_euclidean(const double *a, const double *b, int64_t n) {
    double sum = 0.0;
    for (int64_t i = 0; i < n; i++) {
        double diff = a[i] - b[i];
        sum += diff * diff;
    }
    return sqrt(sum);
}

/* ---- Go's math.Acos, restated (the Go standard library is a dependency that is not under /root/reference:
 * go 1.21 per go.mod:3, src/math/asin.go and src/math/atan.go, pure Go on amd64).  Published algorithm (Cephes):
 *   Acos(x) = Pi/2 - Asin(x)
 *   Asin(x): x == 0 -> x; work on |x|; |x| > 1 -> NaN; t = Sqrt(1 - x*x);
 *            |x| > 0.7 ? Pi/2 - satan(t/|x|) : satan(|x|/t); restore the sign
 *   satan(x): x <= 0.66 -> xatan(x); x > tan(3pi/8) -> Pi/2 - xatan(1/x) + Morebits;
 *             else Pi/4 + xatan((x-1)/(x+1)) + 0.5*Morebits
 *   xatan(x): z = x*x; z = z * P(z)/Q(z) (degree 4 / degree 5, coefficients below); x*z + x
 * The coefficients are the Cephes atan.c ones, written down from memory (the Go source is not available here to diff
 * against): tests/test_oracle.py checks the restatement against libm atan/acos over dense sweeps -- a wrong
 * coefficient would show as an error of many ulp, the observed maximum is reported there.  orc_angular and the
 * cosine hyperplane distance use this function, so the oracle follows the reference's arithmetic to the last
 * library call; orc_libm_acos_mode(1) switches back to libm for comparison. */
static double go_xatan(double x) {
    const double P0 = -8.750608600031904122785e-01, P1 = -1.615753718733365076637e+01, P2 = -7.500855792314704667340e+01,
                 P3 = -1.228866684490136173410e+02, P4 = -6.485021904942025371773e+01;
    const double Q0 = +2.485846490142306297962e+01, Q1 = +1.650270098316988542046e+02, Q2 = +4.328810604912902668951e+02,
                 Q3 = +4.853903996359136964868e+02, Q4 = +1.945506571482613964425e+02;
    double z = x * x;
    z = z * ((((P0 * z + P1) * z + P2) * z + P3) * z + P4) / (((((z + Q0) * z + Q1) * z + Q2) * z + Q3) * z + Q4);
    z = x * z + x;
    return z;
}
static double go_satan(double x) {
    const double Morebits = 6.123233995736765886130e-17; /* pi/2 = PIO2 + Morebits */
    const double Tan3pio8 = 2.41421356237309504880;      /* tan(3*pi/8) */
    if (x <= 0.66) return go_xatan(x);
    if (x > Tan3pio8) return M_PI / 2 - go_xatan(1 / x) + Morebits;
    return M_PI / 4 + go_xatan((x - 1) / (x + 1)) + 0.5 * Morebits;
}
double orc_go_atan(double x) { /* math.Atan */
    if (x == 0) return x;
    if (x > 0) return go_satan(x);
    return -go_satan(-x);
}
double orc_go_asin(double x) { /* math.Asin */
    if (x == 0) return x;
    int sign = 0;
    if (x < 0) { x = -x; sign = 1; }
    if (x > 1) return NAN;
    double temp = sqrt(1 - x * x);
    if (x > 0.7) temp = M_PI / 2 - go_satan(temp / x);
    else temp = go_satan(x / temp);
    if (sign) temp = -temp;
    return temp;
}
double orc_go_acos(double x) { return M_PI / 2 - orc_go_asin(x); } /* math.Acos; NaN for |x| > 1 and for NaN */

static int g_libm_acos = 0;
void orc_libm_acos_mode(int on) { g_libm_acos = on; }
static double ref_acos(double x) { return g_libm_acos ? acos(x) : orc_go_acos(x); }

/* collection.go:821-832.  math.Acos(x > 1) = NaN. */
double orc_angular(const double *a, const double *b, int64_t n) {
    double dot = 0.0, m1 = 0.0, m2 = 0.0;
    for (int64_t i = 0; i < n; i++) {
        dot += a[i] * b[i];
        m1 += a[i] * a[i];
        m2 += b[i] * b[i];
    }
    if (m1 == 0 || m2 == 0) return 1.0;
    return ref_acos(dot / (sqrt(m1) * sqrt(m2))) / M_PI;
}

double orc_distance(int metric, const double *a, const double *b, int64_t n) {
    return metric == ORC_EUCLIDEAN ? orc_euclidean(a, b, n) : orc_angular(a, b, n);
}

/* ------------------------------------------ Go container/heap, max by priority */

typedef struct {
    uint64_t id;
    double dist; /* SearchResult.Distance == resultItem.Priority, collection.go:599-603 */
} orc_item;

typedef struct {
    orc_item *a;
    int64_t n, cap;
} orc_heap;

/* resultPriorityQueue.Less, collection.go:545-547 */
static int heap_less(const orc_heap *h, int64_t i, int64_t j) { return h->a[i].dist > h->a[j].dist; }
static void heap_swap(orc_heap *h, int64_t i, int64_t j) {
    orc_item t = h->a[i];
    h->a[i] = h->a[j];
    h->a[j] = t;
}
/* container/heap.up (Go 1.21 src/container/heap/heap.go) */
static void heap_up(orc_heap *h, int64_t j) {
    for (;;) {
        int64_t i = (j - 1) / 2;
        if (i == j || !heap_less(h, j, i)) break;
        heap_swap(h, i, j);
        j = i;
    }
}
/* container/heap.down */
static void heap_down(orc_heap *h, int64_t i0, int64_t n) {
    int64_t i = i0;
    for (;;) {
        int64_t j1 = 2 * i + 1;
        if (j1 >= n || j1 < 0) break;
        int64_t j = j1;
        int64_t j2 = j1 + 1;
        if (j2 < n && heap_less(h, j2, j1)) j = j2;
        if (!heap_less(h, j, i)) break;
        heap_swap(h, i, j);
        i = j;
    }
}
static void heap_push(orc_heap *h, orc_item it) {
    if (h->n == h->cap) {
        h->cap = h->cap ? h->cap * 2 : 16;
        h->a = (orc_item *)realloc(h->a, (size_t)h->cap * sizeof(orc_item));
    }
    h->a[h->n++] = it;
    heap_up(h, h->n - 1);
}
static orc_item heap_pop(orc_heap *h) {
    int64_t n = h->n - 1;
    heap_swap(h, 0, n);
    heap_down(h, 0, n);
    h->n = n;
    return h->a[n];
}

/* --------------------------------------------------- the `consider` closure */

typedef struct {
    /* collection view: row-major records, exactly stream 1 of each span */
    const uint8_t *codes;
    const uint64_t *ids; /* ids[row] */
    int64_t nrows;
    int64_t dims;
    int bits;
    int metric;
    int64_t rowbytes;
    /* search args, collection.go:140-158 */
    const double *query;
    int64_t k;
    double radius;
    const uint8_t *pass; /* pass[row] != 0 <=> Filter(id, metadata) true; NULL = no filter */
    /* state */
    orc_heap heap;
    int64_t points_searched;
    double *scratch; /* decoded row: decodeVector's make([]float64) */
    int faithful;    /* 1: every record goes through getDocument's whole path (orc_get_document below) */
    const struct orc_spans *spans; /* faithful: the span-file image of the rows */
} orc_search;

/* --------------------------------------------------- getDocument, the faithful CPU variant
 * What the reference does per record BEFORE the arithmetic: collection.go:470-484 formats the id (fmt.Sprintf "%d"),
 * SpanFile.ReadRecord (spanfile.go:513-519) looks the string up in map[string]uint64, parseSpan (730-818) walks magic,
 * length, the 7-code sequence number, the record id and the streams, and verifyChecksum (841-849) runs CRC-32/IEEE over
 * the WHOLE span on every read; decodeVector (768-794) then allocates the []float64.  orc_spans is an in-memory image of
 * those spans for a row-major code matrix (built once, outside any timed region), with a string-keyed hash index.
 * CRC: slicing-by-8 here; Go's hash/crc32 uses PCLMULQDQ on amd64 and is faster -- the figure this variant produces is a
 * restatement's, reported beside the lean one, never instead of it. */
static uint32_t g_crc_tab[8][256];
static int g_crc_ready = 0;
static void crc_init(void) {
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        g_crc_tab[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; i++)
        for (int t = 1; t < 8; t++) g_crc_tab[t][i] = (g_crc_tab[t - 1][i] >> 8) ^ g_crc_tab[0][g_crc_tab[t - 1][i] & 0xFF];
    g_crc_ready = 1;
}
static uint32_t crc32_ieee(const uint8_t *p, size_t n) {
    uint32_t c = 0xFFFFFFFFu;
    while (n >= 8) {
        uint32_t a, b;
        memcpy(&a, p, 4);
        memcpy(&b, p + 4, 4);
        a ^= c;
        c = g_crc_tab[7][a & 0xFF] ^ g_crc_tab[6][(a >> 8) & 0xFF] ^ g_crc_tab[5][(a >> 16) & 0xFF] ^ g_crc_tab[4][a >> 24] ^
            g_crc_tab[3][b & 0xFF] ^ g_crc_tab[2][(b >> 8) & 0xFF] ^ g_crc_tab[1][(b >> 16) & 0xFF] ^ g_crc_tab[0][b >> 24];
        p += 8;
        n -= 8;
    }
    while (n--) c = g_crc_tab[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
    return ~c;
}
uint32_t orc_crc32(const uint8_t *p, int64_t n) { /* check value: "123456789" -> 0xCBF43926 */
    if (!g_crc_ready) crc_init();
    return crc32_ieee(p, (size_t)n);
}

/* write7Code, spanfile.go:568-625 (thresholds off by one: n < 0x7f takes 1 byte, so 127 takes 2) */
static size_t write7(uint8_t *out, uint64_t v) {
    int len = v < 0x7f ? 1 : v < 0x3fff ? 2 : v < 0x1fffff ? 3 : v < 0xfffffff ? 4 : 5;
    for (int i = 0; i < len; i++) out[i] = (uint8_t)(((v >> (7 * (len - 1 - i))) & 0x7f) | (i + 1 < len ? 0x80 : 0));
    return (size_t)len;
}
/* read7Code, spanfile.go:627-636 */
static uint64_t read7(const uint8_t *p, size_t *at) {
    uint64_t v = 0;
    for (;;) {
        uint8_t b = p[(*at)++];
        v = (v << 7) | (b & 0x7f);
        if (!(b & 0x80)) return v;
    }
}

typedef struct orc_spans {
    uint8_t *buf;
    size_t len;
    int64_t nrows;
    uint64_t *slot_off;  /* open addressing: offset + 1 of the span whose record id hashes here (0 = empty) */
    size_t nslots;       /* power of two */
} orc_spans;

static uint64_t str_hash(const char *s, size_t n) { /* FNV-1a: a string hash like the one Go's map runs on the key */
    uint64_t h = 1469598103934665603ULL;
    for (size_t i = 0; i < n; i++) h = (h ^ (uint8_t)s[i]) * 1099511628211ULL;
    return h;
}

/* serializeSpan + WriteRecord (spanfile.go:679-728, 398-475) for every row: streams {0: metadata, 1: vector bytes}
 * (collection.go:446-449), sequence numbers 1.., the CRC at the end.  meta_len bytes of metadata per record. */
orc_spans *orc_spans_build(const uint8_t *codes, const uint64_t *ids, int64_t nrows, int64_t rowbytes, int64_t meta_len) {
    if (!g_crc_ready) crc_init();
    orc_spans *sp = (orc_spans *)calloc(1, sizeof *sp);
    sp->nrows = nrows;
    size_t cap = (size_t)nrows * ((size_t)rowbytes + (size_t)meta_len + 48) + 64;
    sp->buf = (uint8_t *)malloc(cap);
    sp->nslots = 16;
    while (sp->nslots < (size_t)nrows * 2) sp->nslots <<= 1;
    sp->slot_off = (uint64_t *)calloc(sp->nslots, sizeof(uint64_t));
    size_t at = 0;
    for (int64_t r = 0; r < nrows; r++) {
        char key[24];
        int klen = snprintf(key, sizeof key, "%llu", (unsigned long long)ids[r]);
        const size_t start = at;
        uint8_t *b = sp->buf;
        memcpy(b + at, "SPAN", 4);
        at += 8; /* length patched below */
        at += write7(b + at, (uint64_t)r + 1);
        at += write7(b + at, (uint64_t)klen);
        memcpy(b + at, key, (size_t)klen);
        at += (size_t)klen;
        b[at++] = 2;
        b[at++] = 0;
        at += write7(b + at, (uint64_t)meta_len);
        memset(b + at, '{', (size_t)meta_len);
        at += (size_t)meta_len;
        b[at++] = 1;
        at += write7(b + at, (uint64_t)rowbytes);
        memcpy(b + at, codes + r * rowbytes, (size_t)rowbytes);
        at += (size_t)rowbytes;
        const uint32_t total = (uint32_t)(at - start + 4);
        b[start + 4] = (uint8_t)(total >> 24); b[start + 5] = (uint8_t)(total >> 16);
        b[start + 6] = (uint8_t)(total >> 8); b[start + 7] = (uint8_t)total;
        const uint32_t crc = crc32_ieee(b + start, at - start);
        b[at++] = (uint8_t)(crc >> 24); b[at++] = (uint8_t)(crc >> 16); b[at++] = (uint8_t)(crc >> 8); b[at++] = (uint8_t)crc;
        size_t h = (size_t)str_hash(key, (size_t)klen) & (sp->nslots - 1);
        while (sp->slot_off[h]) h = (h + 1) & (sp->nslots - 1);
        sp->slot_off[h] = (uint64_t)start + 1;
    }
    sp->len = at;
    return sp;
}
void orc_spans_free(orc_spans *sp) {
    if (!sp) return;
    free(sp->buf);
    free(sp->slot_off);
    free(sp);
}

/* getDocument (collection.go:470-484): id -> decimal string -> index lookup -> parseSpan + verifyChecksum -> stream 1.
 * Returns the vector bytes inside the image (NULL: record not found or checksum mismatch). */
static const uint8_t *orc_get_document(const orc_spans *sp, uint64_t id, int64_t *veclen) {
    char key[24];
    const int klen = snprintf(key, sizeof key, "%llu", (unsigned long long)id); /* fmt.Sprintf("%d", id) */
    size_t h = (size_t)str_hash(key, (size_t)klen) & (sp->nslots - 1);
    for (;; h = (h + 1) & (sp->nslots - 1)) {
        if (!sp->slot_off[h]) return NULL;
        const uint8_t *b = sp->buf + (sp->slot_off[h] - 1);
        /* parseSpan, spanfile.go:730-818 */
        if (memcmp(b, "SPAN", 4) != 0) return NULL;
        const uint32_t length = ((uint32_t)b[4] << 24) | ((uint32_t)b[5] << 16) | ((uint32_t)b[6] << 8) | b[7];
        size_t at = 8;
        (void)read7(b, &at); /* sequence number */
        const uint64_t idlen = read7(b, &at);
        if (idlen != (uint64_t)klen || memcmp(b + at, key, (size_t)klen) != 0) continue; /* another key in this bucket chain */
        at += (size_t)idlen;
        const int nstreams = b[at++];
        const uint8_t *vec = NULL;
        for (int s = 0; s < nstreams; s++) {
            const int sid = b[at++];
            const uint64_t slen = read7(b, &at);
            if (sid == 1) { vec = b + at; *veclen = (int64_t)slen; }
            at += (size_t)slen;
        }
        /* verifyChecksum, spanfile.go:841-849: over everything but the last 4 bytes, on every read */
        const uint32_t want = ((uint32_t)b[length - 4] << 24) | ((uint32_t)b[length - 3] << 16) | ((uint32_t)b[length - 2] << 8) | b[length - 1];
        if (crc32_ieee(b, length - 4) != want) return NULL;
        return vec;
    }
}

/* collection.go:583-629.  `row` is the record already resolved (getDocument found it);
 * row < 0 stands for "record not found" (StopSearch, 585-587). */
static int orc_consider(orc_search *s, int64_t row, double *radius) {
    if (row < 0) return ORC_STOP_SEARCH;
    double *vec = s->scratch;
    const uint8_t *bytes = s->codes + row * s->rowbytes;
    if (s->faithful) {
        if (s->spans) { /* the record as getDocument reaches it: string key, index, parseSpan, CRC over the span */
            int64_t vl = 0;
            bytes = orc_get_document(s->spans, s->ids[row], &vl);
            if (!bytes || vl != s->rowbytes) return ORC_STOP_SEARCH; /* 585-587 */
        }
        vec = (double *)malloc((size_t)s->dims * sizeof(double)); /* decodeVector's make([]float64, dims) */
    }
    orc_decode(bytes, s->dims, s->bits, vec); /* 470-484 */
    s->points_searched++;                                             /* 589 */
    int signal = ORC_POINT_CHECKED;
    if (s->pass && !s->pass[row]) { /* 592-594 */
        signal = ORC_POINT_IGNORED;
        goto done;
    }
    {
        double distance = orc_distance(s->metric, s->query, vec, s->dims); /* 596 */
        orc_item it = {s->ids[row], distance};
        if (s->radius > 0 && distance <= s->radius) { /* 598-603 */
            heap_push(&s->heap, it);
            signal = ORC_POINT_ACCEPTED;
        } else if (s->radius > 0) { /* 604-605 */
            signal = ORC_POINT_CHECKED;
        } else if (s->k > 0) { /* 606-620 */
            if (s->heap.n <= s->k) {
                if (s->heap.n < s->k || s->heap.a[0].dist > distance) {
                    heap_push(&s->heap, it);
                    if (s->heap.n > s->k) heap_pop(&s->heap);
                    *radius = s->heap.a[0].dist;
                    signal = ORC_POINT_ACCEPTED;
                }
            }
        } else if (s->k == 0 && s->radius == 0) { /* 621-627 (unreachable from Search: list mode) */
            heap_push(&s->heap, it);
            signal = ORC_POINT_ACCEPTED;
        }
    }
done:
    if (s->faithful) free(vec);
    return signal;
}

/* collection.go:693-697: pop back to front => ascending distance */
static int64_t orc_drain(orc_search *s, uint64_t *out_ids, double *out_dist, int64_t out_cap) {
    int64_t n = s->heap.n;
    for (int64_t i = n - 1; i >= 0; i--) {
        orc_item it = heap_pop(&s->heap);
        if (i < out_cap) {
            out_ids[i] = it.id;
            out_dist[i] = it.dist;
        }
    }
    free(s->heap.a);
    s->heap.a = NULL;
    return n;
}

/* ---------------------------------------------------------------- scan order */

static int lexcmp_u64(uint64_t a, uint64_t b) {
    char sa[24], sb[24];
    snprintf(sa, sizeof sa, "%llu", (unsigned long long)a);
    snprintf(sb, sizeof sb, "%llu", (unsigned long long)b);
    return strcmp(sa, sb);
}
static const uint64_t *g_sort_ids;
static int perm_cmp(const void *pa, const void *pb) {
    int64_t a = *(const int64_t *)pa, b = *(const int64_t *)pb;
    return lexcmp_u64(g_sort_ids[a], g_sort_ids[b]);
}
/* spanfile.go:540-560 IterateSortedRecords: sort.Strings over decimal ids ("10" < "2").
 * This is the scan order whenever the RNG is seeded (spanfile.go:522-524); it is the
 * canonical deterministic order of the oracle.  Not thread-safe. */
void orc_lex_order(const uint64_t *ids, int64_t n, int64_t *perm) {
    for (int64_t i = 0; i < n; i++) perm[i] = i;
    g_sort_ids = ids;
    qsort(perm, (size_t)n, sizeof(int64_t), perm_cmp);
}

/* ------------------------------------------------------------- exact search */

/*
 * Search with Precision=="exact": collection.go:569-711, branch 672-684.
 * order: scan order as row indices (NULL = rows as given).  Returns the number of
 * results (may exceed out_cap in radius mode; only out_cap are written).
 * *percent_searched follows 700-710 (nrows == numRecords).
 */
static int64_t search_exact_impl(const uint8_t *codes, const uint64_t *ids, int64_t nrows, int64_t dims,
                         int bits, int metric, const double *query, int64_t k, double radius,
                         const uint8_t *pass, const int64_t *order, int faithful, const orc_spans *spans,
                         uint64_t *out_ids, double *out_dist, int64_t out_cap,
                         double *percent_searched);
int64_t orc_search_exact(const uint8_t *codes, const uint64_t *ids, int64_t nrows, int64_t dims,
                         int bits, int metric, const double *query, int64_t k, double radius,
                         const uint8_t *pass, const int64_t *order, int faithful,
                         uint64_t *out_ids, double *out_dist, int64_t out_cap,
                         double *percent_searched) {
    return search_exact_impl(codes, ids, nrows, dims, bits, metric, query, k, radius, pass, order, faithful, NULL, out_ids, out_dist,
                             out_cap, percent_searched);
}
/* the faithful variant: every record through orc_get_document over the span image `spans` of the same rows */
int64_t orc_search_exact_spans(const orc_spans *spans, const uint8_t *codes, const uint64_t *ids, int64_t nrows, int64_t dims,
                               int bits, int metric, const double *query, int64_t k, double radius, const uint8_t *pass,
                               const int64_t *order, uint64_t *out_ids, double *out_dist, int64_t out_cap,
                               double *percent_searched) {
    return search_exact_impl(codes, ids, nrows, dims, bits, metric, query, k, radius, pass, order, 1, spans, out_ids, out_dist, out_cap,
                             percent_searched);
}
static int64_t search_exact_impl(const uint8_t *codes, const uint64_t *ids, int64_t nrows, int64_t dims,
                         int bits, int metric, const double *query, int64_t k, double radius,
                         const uint8_t *pass, const int64_t *order, int faithful, const orc_spans *spans,
                         uint64_t *out_ids, double *out_dist, int64_t out_cap,
                         double *percent_searched) {
    orc_search s;
    memset(&s, 0, sizeof s);
    s.spans = spans;
    s.codes = codes; s.ids = ids; s.nrows = nrows; s.dims = dims; s.bits = bits;
    s.metric = metric; s.rowbytes = orc_vector_size(bits, dims);
    s.query = query; s.k = k; s.radius = radius; s.pass = pass; s.faithful = faithful;
    s.scratch = (double *)malloc((size_t)(dims > 0 ? dims : 1) * sizeof(double));
    int64_t nout = 0;
    if (!(radius == 0 && k == 0)) { /* list mode (633-668) is not on the hot path */
        for (int64_t i = 0; i < nrows; i++) {
            double r = 1.7976931348623157e308; /* math.MaxFloat64, 679 */
            orc_consider(&s, order ? order[i] : i, &r);
        }
        nout = orc_drain(&s, out_ids, out_dist, out_cap);
    }
    if (percent_searched)
        *percent_searched = nrows == 0 ? 0.0 : (double)s.points_searched / (double)nrows * 100;
    free(s.scratch);
    return nout;
}

/*
 * Replay of `consider` over an explicit visit sequence of row indices (the ids an
 * index fed to the callback), with no tree logic: used to check GPU rescoring
 * (SURVEY.md appendix B-13).  visit[i] < 0 => StopSearch.
 */
int64_t orc_replay(const uint8_t *codes, const uint64_t *ids, int64_t nrows, int64_t dims, int bits,
                   int metric, const double *query, int64_t k, double radius, const uint8_t *pass,
                   const int64_t *visit, int64_t nvisit, uint64_t *out_ids, double *out_dist,
                   int64_t out_cap, int64_t *points_searched) {
    orc_search s;
    memset(&s, 0, sizeof s);
    s.codes = codes; s.ids = ids; s.nrows = nrows; s.dims = dims; s.bits = bits;
    s.metric = metric; s.rowbytes = orc_vector_size(bits, dims);
    s.query = query; s.k = k; s.radius = radius; s.pass = pass;
    s.scratch = (double *)malloc((size_t)(dims > 0 ? dims : 1) * sizeof(double));
    double r = radius > 0 ? radius : 1.7976931348623157e308;
    for (int64_t i = 0; i < nvisit; i++)
        if (orc_consider(&s, visit[i], &r) == ORC_STOP_SEARCH) break;
    if (points_searched) *points_searched = s.points_searched;
    int64_t n = orc_drain(&s, out_ids, out_dist, out_cap);
    free(s.scratch);
    return n;
}

/* distances of a list of rows to the query (what szg_rescore must reproduce) */
void orc_row_distances(const uint8_t *codes, int64_t dims, int bits, int metric, const double *query,
                       const int64_t *rows, int64_t n, double *out) {
    double *vec = (double *)malloc((size_t)(dims > 0 ? dims : 1) * sizeof(double));
    int64_t rb = orc_vector_size(bits, dims);
    for (int64_t i = 0; i < n; i++) {
        orc_decode(codes + rows[i] * rb, dims, bits, vec);
        out[i] = orc_distance(metric, query, vec, dims);
    }
    free(vec);
}

/* --------------------------------------------------- deterministic randomness */
/*
 * The reference draws from Go math/rand (settings.go:42-76), whose stream cannot be
 * reproduced outside Go, and shares one rand.Rand between 5 goroutines
 * (lshtree.go:101-114), so its tree shape is not reproducible even when seeded.
 * The oracle and the product therefore share THIS generator (splitmix64 counter
 * hash); LSH parity is defined per visit sequence, never per seed.
 */
uint64_t orc_mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
uint64_t orc_rand_u64(uint64_t seed, uint64_t ctr) {
    return orc_mix64(seed + 0x9E3779B97F4A7C15ULL * (ctr + 1));
}
double orc_rand_unit(uint64_t seed, uint64_t ctr) { /* uniform [0,1) */
    return (double)(orc_rand_u64(seed, ctr) >> 11) * (1.0 / 9007199254740992.0);
}

/*
 * Synthetic collection rows, identical in the CUDA generator
 * (syzgydb_b200/csrc/synth.cuh).  4/8/16-bit: the row is the first rowbytes bytes of
 * the little-endian u64 words rand(seed, row*wpr + j), wpr = ceil(rowbytes/8)
 * (uniform over the full code range).  32/64-bit: element e is
 * v = unit(seed, row*dims + e)*2 - 1, stored as quantize(v) big-endian.
 */
void orc_synth_rows(uint64_t seed, int64_t row0, int64_t nrows, int64_t dims, int bits, uint8_t *out) {
    int64_t rb = orc_vector_size(bits, dims);
    if (bits <= 16) {
        int64_t wpr = (rb + 7) / 8;
        for (int64_t r = 0; r < nrows; r++) {
            uint8_t *dst = out + r * rb;
            for (int64_t j = 0; j < wpr; j++) {
                uint64_t w = orc_rand_u64(seed, (uint64_t)((row0 + r) * wpr + j));
                for (int b = 0; b < 8 && j * 8 + b < rb; b++) dst[j * 8 + b] = (uint8_t)(w >> (8 * b));
            }
            if (bits == 4 && (dims & 1)) dst[rb - 1] &= 0xF0; /* odd d: last low nibble is 0 (encodeDocument) */
        }
    } else {
        int eb = bits / 8;
        for (int64_t r = 0; r < nrows; r++)
            for (int64_t e = 0; e < dims; e++) {
                double v = orc_rand_unit(seed, (uint64_t)((row0 + r) * dims + e)) * 2 - 1;
                uint64_t q = orc_quantize(v, bits);
                uint8_t *dst = out + r * rb + e * eb;
                for (int b = 0; b < eb; b++) dst[b] = (uint8_t)(q >> (8 * (eb - 1 - b)));
            }
    }
}

/* synthetic queries: q[qi][i] = unit(seed, qi*dims+i)*2-1, never copied from the rows */
void orc_synth_queries(uint64_t seed, int64_t q0, int64_t nq, int64_t dims, double *out) {
    for (int64_t q = 0; q < nq; q++)
        for (int64_t i = 0; i < dims; i++)
            out[q * dims + i] = orc_rand_unit(seed, (uint64_t)((q0 + q) * dims + i)) * 2 - 1;
}

/* ------------------------------------------------------------------ LSH tree */

typedef struct orc_node {
    double *normal;
    double b;
    double radius; /* tracked, never read by search (lshtree.go:46-52) */
    struct orc_node *left, *right;
    int64_t *rows; /* node.ids, as row indices */
    int64_t nrows, cap;
} orc_node;

typedef struct {
    orc_node **roots;
    int ntrees;
    int threshold;
    /* collection view for getDocument inside split */
    const uint8_t *codes;
    int64_t dims;
    int bits;
    int metric;
    int64_t rowbytes;
    uint64_t seed, ctr; /* sequential draws from orc_rand_u64(seed, ctr++) */
} orc_lsh;

static uint64_t lsh_next(orc_lsh *t) { return orc_rand_u64(t->seed, t->ctr++); }
static int64_t lsh_intn(orc_lsh *t, int64_t n) { return (int64_t)(lsh_next(t) % (uint64_t)n); }
static double lsh_unit(orc_lsh *t) { return (double)(lsh_next(t) >> 11) * (1.0 / 9007199254740992.0); }
/* stands in for rand.NormFloat64 (Box-Muller, one draw pair per value) */
static double lsh_norm(orc_lsh *t) {
    double u1 = lsh_unit(t), u2 = lsh_unit(t);
    if (u1 < 1e-300) u1 = 1e-300;
    return sqrt(-2.0 * log(u1)) * cos(2.0 * M_PI * u2);
}

/* lshtree.go:136-145 */
static double dot_product(const double *a, const double *b, int64_t n) {
    double dot = 0.0;
    for (int64_t i = 0; i < n; i++) dot += a[i] * b[i];
    return dot;
}
/* lshtree.go:30-36 */
static double vector_length(const double *v, int64_t n) {
    double sum = 0.0;
    for (int64_t i = 0; i < n; i++) sum += v[i] * v[i];
    return sqrt(sum);
}
/* lshtree.go:59-77 */
static double distance_to_hyperplane(int method, const double *v, double length, const double *normal,
                                     double b, int64_t n, int *right) {
    double dist = dot_product(v, normal, n) - b;
    *right = 0;
    if (method == ORC_EUCLIDEAN) {
        if (dist > 0) *right = 1;
        else dist = -dist;
        return dist;
    }
    dist = ref_acos(dist / length) / M_PI;
    if (dist > 0.5) {
        *right = 1;
        dist = 1 - dist;
    }
    return dist;
}

static orc_node *node_new(void) { return (orc_node *)calloc(1, sizeof(orc_node)); }
static void node_append(orc_node *nd, int64_t row) {
    if (nd->nrows == nd->cap) {
        nd->cap = nd->cap ? nd->cap * 2 : 8;
        nd->rows = (int64_t *)realloc(nd->rows, (size_t)nd->cap * sizeof(int64_t));
    }
    nd->rows[nd->nrows++] = row;
}
static int node_is_leaf(const orc_node *nd) { return nd->left == NULL; } /* lshtree.go:55-57 */

/* lshtree.go:172-248 */
static orc_node *lsh_split(orc_lsh *t, orc_node *node) {
    int64_t d = t->dims;
    int64_t i1 = lsh_intn(t, node->nrows), i2;
    do { i2 = lsh_intn(t, node->nrows); } while (i2 == i1);
    double *v1 = (double *)malloc((size_t)d * 8), *v2 = (double *)malloc((size_t)d * 8);
    orc_decode(t->codes + node->rows[i1] * t->rowbytes, d, t->bits, v1);
    orc_decode(t->codes + node->rows[i2] * t->rowbytes, d, t->bits, v2);
    int about_equal = 1; /* lshtree.go:158-170, tolerance 1e-9 */
    for (int64_t i = 0; i < d; i++)
        if (fabs(v1[i] - v2[i]) > 1e-9) { about_equal = 0; break; }
    if (about_equal) { free(v1); free(v2); return node; }
    double *mid = v1; /* midpoint, lshtree.go:147-156 */
    for (int64_t i = 0; i < d; i++) mid[i] = (v1[i] + v2[i]) / 2;
    /* randomNormalizedVector, lshtree.go:38-44 + normalizeVector 10-28 */
    double *normal = (double *)malloc((size_t)d * 8);
    for (int64_t i = 0; i < d; i++) normal[i] = lsh_norm(t);
    double norm = 0.0;
    for (int64_t i = 0; i < d; i++) norm += normal[i] * normal[i];
    if (norm != 0) {
        norm = sqrt(norm);
        for (int64_t i = 0; i < d; i++) normal[i] = normal[i] / norm;
    }
    double b = 0.0;
    if (t->metric == ORC_EUCLIDEAN) b = sqrt(dot_product(mid, mid, d)); /* 207 */
    orc_node *l = node_new(), *r = node_new();
    double radius = 0.0;
    for (int64_t i = 0; i < node->nrows; i++) {
        orc_decode(t->codes + node->rows[i] * t->rowbytes, d, t->bits, v2);
        double length = vector_length(v2, d);
        int right;
        double dist = distance_to_hyperplane(t->metric, v2, length, normal, b, d, &right);
        radius = fmax(radius, dist);
        node_append(right ? r : l, node->rows[i]);
    }
    free(v1); free(v2);
    if (l->nrows == 0 || r->nrows == 0) { /* 236-238 */
        free(l->rows); free(r->rows); free(l); free(r); free(normal);
        return node;
    }
    orc_node *inner = node_new();
    inner->normal = normal; inner->b = b; inner->radius = radius;
    inner->left = l; inner->right = r;
    free(node->rows); free(node);
    return inner;
}

/* lshtree.go:116-134 */
static orc_node *lsh_insert(orc_lsh *t, orc_node *node, int64_t row, const double *vec, double length) {
    if (node_is_leaf(node)) {
        node_append(node, row);
        if (node->nrows > t->threshold) node = lsh_split(t, node);
        return node;
    }
    int right;
    double dist = distance_to_hyperplane(t->metric, vec, length, node->normal, node->b, t->dims, &right);
    node->radius = fmax(node->radius, dist);
    if (!right) node->left = lsh_insert(t, node->left, row, vec, length);
    else node->right = lsh_insert(t, node->right, row, vec, length);
    return node;
}

/* newLSHTree(c, 100, 5), collection.go:292, lshtree.go:88-99 */
orc_lsh *orc_lsh_new(const uint8_t *codes, int64_t dims, int bits, int metric, int threshold,
                     int ntrees, uint64_t seed) {
    orc_lsh *t = (orc_lsh *)calloc(1, sizeof(orc_lsh));
    t->roots = (orc_node **)calloc((size_t)ntrees, sizeof(orc_node *));
    for (int i = 0; i < ntrees; i++) t->roots[i] = node_new();
    t->ntrees = ntrees; t->threshold = threshold; t->codes = codes; t->dims = dims;
    t->bits = bits; t->metric = metric; t->rowbytes = orc_vector_size(bits, dims);
    t->seed = seed; t->ctr = 0;
    return t;
}

/* lshtree.go:101-114; the 5 goroutines are run as trees 0..4 in order (one legal
 * interleaving).  vec is the UNQUANTIZED vector AddDocument received (collection.go:456)
 * or the decoded one on reload (collection.go:305-306). */
void orc_lsh_add(orc_lsh *t, int64_t row, const double *vec) {
    double length = vector_length(vec, t->dims);
    for (int i = 0; i < t->ntrees; i++) t->roots[i] = lsh_insert(t, t->roots[i], row, vec, length);
}

static void node_free(orc_node *nd) {
    if (!nd) return;
    node_free(nd->left); node_free(nd->right);
    free(nd->normal); free(nd->rows); free(nd);
}
void orc_lsh_free(orc_lsh *t) {
    for (int i = 0; i < t->ntrees; i++) node_free(t->roots[i]);
    free(t->roots); free(t);
}

/* nodePriorityQueue (lshtree.go:353-381) on Go container/heap */
typedef struct { orc_node *node; double prio; } nq_item;
typedef struct { nq_item *a; int64_t n, cap; } nq_heap;
static int nq_less(nq_heap *h, int64_t i, int64_t j) { return h->a[i].prio > h->a[j].prio; }
static void nq_swap(nq_heap *h, int64_t i, int64_t j) { nq_item t = h->a[i]; h->a[i] = h->a[j]; h->a[j] = t; }
static void nq_push(nq_heap *h, orc_node *node, double prio) {
    if (h->n == h->cap) {
        h->cap = h->cap ? h->cap * 2 : 16;
        h->a = (nq_item *)realloc(h->a, (size_t)h->cap * sizeof(nq_item));
    }
    h->a[h->n].node = node; h->a[h->n].prio = prio;
    int64_t j = h->n++;
    for (;;) {
        int64_t i = (j - 1) / 2;
        if (i == j || !nq_less(h, j, i)) break;
        nq_swap(h, i, j);
        j = i;
    }
}
static nq_item nq_pop(nq_heap *h) {
    int64_t n = h->n - 1;
    nq_swap(h, 0, n);
    int64_t i = 0;
    for (;;) {
        int64_t j1 = 2 * i + 1;
        if (j1 >= n || j1 < 0) break;
        int64_t j = j1, j2 = j1 + 1;
        if (j2 < n && nq_less(h, j2, j1)) j = j2;
        if (!nq_less(h, j, i)) break;
        nq_swap(h, i, j);
        i = j;
    }
    h->n = n;
    return h->a[n];
}

/*
 * Search with Precision != "exact": collection.go:685-691 -> lshtree.go:283-351 with
 * `consider` as the callback.  visit_out (optional, cap visit_cap) receives the row
 * indices in the order they were fed to the callback; *nvisit their count.
 */
int64_t orc_search_lsh(orc_lsh *t, const uint64_t *ids, int64_t nrows, const double *query, int64_t k,
                       double radius_arg, const uint8_t *pass, uint64_t *out_ids, double *out_dist,
                       int64_t out_cap, double *percent_searched, int64_t *visit_out,
                       int64_t visit_cap, int64_t *nvisit) {
    orc_search s;
    memset(&s, 0, sizeof s);
    s.codes = t->codes; s.ids = ids; s.nrows = nrows; s.dims = t->dims; s.bits = t->bits;
    s.metric = t->metric; s.rowbytes = t->rowbytes; s.query = query; s.k = k;
    s.radius = radius_arg; s.pass = pass;
    s.scratch = (double *)malloc((size_t)(t->dims > 0 ? t->dims : 1) * sizeof(double));
    int64_t nv = 0;
    if (!(radius_arg == 0 && k == 0)) {
        double radius = radius_arg > 0 ? radius_arg : 1.7976931348623157e308; /* 686-689 */
        double length = vector_length(query, t->dims);
        uint8_t *visited = (uint8_t *)calloc((size_t)(nrows > 0 ? nrows : 1), 1);
        const int search_k = 200; /* lshtree.go:286 */
        int k_counter = 0, point_accepted = 0, stop = 0;
        nq_heap pq = {0};
        for (int i = 0; i < t->ntrees; i++) nq_push(&pq, t->roots[i], 0);
        while (pq.n > 0 && !stop) {
            nq_item item = nq_pop(&pq);
            orc_node *node = item.node;
            if (item.prio < 0 && -item.prio > radius && node_is_leaf(node)) continue; /* 304-309 */
            if (k_counter >= search_k) break;                                       /* 311-313 */
            if (node_is_leaf(node)) {
                for (int64_t i = 0; i < node->nrows; i++) {
                    int64_t row = node->rows[i];
                    if (visited[row]) continue;
                    visited[row] = 1;
                    if (visit_out && nv < visit_cap) visit_out[nv] = row;
                    nv++;
                    int signal = orc_consider(&s, row, &radius);
                    if (signal == ORC_STOP_SEARCH) { stop = 1; break; }
                    if (signal == ORC_POINT_ACCEPTED) { k_counter = 0; point_accepted = 1; }
                    else if (signal == ORC_POINT_CHECKED) { if (point_accepted) k_counter++; }
                }
            } else {
                int right;
                double dist = distance_to_hyperplane(t->metric, query, length, node->normal, node->b,
                                                     t->dims, &right);
                if (right) { nq_push(&pq, node->right, dist); nq_push(&pq, node->left, -dist); }
                else { nq_push(&pq, node->left, dist); nq_push(&pq, node->right, -dist); }
            }
        }
        free(pq.a); free(visited);
    }
    if (nvisit) *nvisit = nv;
    if (percent_searched)
        *percent_searched = nrows == 0 ? 0.0 : (double)s.points_searched / (double)nrows * 100;
    int64_t n = orc_drain(&s, out_ids, out_dist, out_cap);
    free(s.scratch);
    return n;
}
