// collection.cpp -- see collection.hpp.  Host logic only; every distance is computed on the GPU through
// the C ABI (include/syzgy_b200.h).  No CPU distance code lives here (lshtree.go's hyperplane
// arithmetic, which is tree navigation rather than the search hot path, stays on the host as in the
// reference).
#include "collection.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <random>
#include <shared_mutex>
#include <stdexcept>
#include <thread>
#include <unordered_map>
#include <unordered_set>

#include "../../include/syzgy_b200.h"

namespace syzgydb {

// ------------------------------------------------------------------------------------ codec
uint64_t quantize(double value, int bits) { // quantization.go:5-23
    if (bits == 32) {
        float f = (float)value;
        uint32_t u;
        std::memcpy(&u, &f, 4);
        return u;
    }
    if (bits == 64) {
        uint64_t u;
        std::memcpy(&u, &value, 8);
        return u;
    }
    if (value < -1) value = -1;
    else if (value > 1) value = 1;
    const int64_t maxInt = ((int64_t)1 << bits) - 1;
    return (uint64_t)std::round((value + 1) / 2 * (double)maxInt); // math.Round: half away from zero
}

double dequantize(uint64_t value, int bits) { // quantization.go:25-36
    if (bits == 32) {
        uint32_t u = (uint32_t)value;
        float f;
        std::memcpy(&f, &u, 4);
        return (double)f;
    }
    if (bits == 64) {
        double d;
        std::memcpy(&d, &value, 8);
        return d;
    }
    const int64_t maxInt = ((int64_t)1 << bits) - 1;
    return ((double)value / (double)maxInt) * 2 - 1;
}

int getVectorSize(int quantization, int dimensions) { // collection.go:796-811
    switch (quantization) {
    case 4: return (dimensions + 1) / 2;
    case 8: return dimensions;
    case 16: return dimensions * 2;
    case 32: return dimensions * 4;
    case 64: return dimensions * 8;
    }
    throw std::invalid_argument("Unsupported quantization level");
}

std::vector<uint8_t> encodeDocument(const std::vector<double> &vector, int quantization) { // collection.go:713-744
    const int dims = (int)vector.size();
    std::vector<uint8_t> data((size_t)getVectorSize(quantization, dims), 0);
    size_t off = 0;
    for (int i = 0; i < dims; ++i) {
        const uint64_t q = quantize(vector[(size_t)i], quantization);
        switch (quantization) {
        case 4:
            if (i % 2 == 0) data[off] = (uint8_t)(q << 4); // even index: high nibble
            else { data[off] |= (uint8_t)(q & 0x0F); ++off; }
            break;
        case 8: data[off++] = (uint8_t)q; break;
        default: {
            const int nb = quantization / 8;
            for (int b = 0; b < nb; ++b) data[off++] = (uint8_t)(q >> (8 * (nb - 1 - b))); // big-endian
        }
        }
    }
    return data;
}

std::vector<double> decodeVector(const uint8_t *data, int dimensions, int quantization) { // collection.go:768-794
    std::vector<double> v((size_t)dimensions);
    size_t off = 0;
    for (int i = 0; i < dimensions; ++i) {
        uint64_t q = 0;
        switch (quantization) {
        case 4:
            if (i % 2 == 0) q = data[off] >> 4;
            else { q = data[off] & 0x0F; ++off; }
            break;
        case 8: q = data[off++]; break;
        default: {
            const int nb = quantization / 8;
            for (int b = 0; b < nb; ++b) q = (q << 8) | data[off++];
        }
        }
        v[(size_t)i] = dequantize(q, quantization);
    }
    return v;
}

namespace {

[[noreturn]] void gpu_fail(const char *what) { // the reference panics on storage faults (collection.go:667, 683)
    throw std::runtime_error(std::string("syzgy_b200 ") + what + ": " + szg_last_error());
}
#define GPU(call, what) do { if ((call) != SZG_OK) gpu_fail(what); } while (0)

// sort.Strings order of decimal ids (spanfile.go:540-560)
bool lex_less(uint64_t a, uint64_t b) { return std::to_string(a) < std::to_string(b); }

// ----------------------------------------------------------------- container/heap, restated
// Go's container/heap over a slice with a user Less: Push = append + up, Pop = swap(0, n-1) +
// down(0, n-1) + remove last.  Needed to reproduce pop order among equal priorities.
template <typename T, typename Less>
struct GoHeap {
    std::vector<T> a;
    Less less;
    void up(size_t j) {
        while (j > 0) {
            size_t i = (j - 1) / 2;
            if (i == j || !less(a[j], a[i])) break;
            std::swap(a[i], a[j]);
            j = i;
        }
    }
    void down(size_t i0, size_t n) {
        size_t i = i0;
        for (;;) {
            size_t j1 = 2 * i + 1;
            if (j1 >= n) break;
            size_t j = j1, j2 = j1 + 1;
            if (j2 < n && less(a[j2], a[j1])) j = j2;
            if (!less(a[j], a[i])) break;
            std::swap(a[i], a[j]);
            i = j;
        }
    }
    void push(T x) { a.push_back(std::move(x)); up(a.size() - 1); }
    T pop() {
        const size_t n = a.size() - 1;
        std::swap(a[0], a[n]);
        down(0, n);
        T x = std::move(a.back());
        a.pop_back();
        return x;
    }
    size_t size() const { return a.size(); }
};

struct ResultItem { uint64_t id; double priority; };
struct ResultLess { bool operator()(const ResultItem &x, const ResultItem &y) const { return x.priority > y.priority; } }; // collection.go:545-547

// ----------------------------------------------------------------------------- lshtree.go
struct LshNode { // lshtree.go:46-52 (radius is tracked by the reference but never read by search)
    std::vector<double> normal;
    double b = 0;
    std::unique_ptr<LshNode> left, right;
    std::vector<uint64_t> ids;
    bool isLeaf() const { return !left; } // lshtree.go:55-57
};

double dotProduct(const std::vector<double> &a, const std::vector<double> &b) {
    double dot = 0.0;
    for (size_t i = 0; i < a.size(); ++i) dot += a[i] * b[i];
    return dot;
}
double vectorLength(const std::vector<double> &v) { return std::sqrt(dotProduct(v, v)); } // lshtree.go:30-36

// Go's math.Acos (standard library, Cephes algorithm: src/math/asin.go, atan.go), restated so that the hyperplane side
// decisions of the LSH trees follow the reference to the last bit (same restatement as oracle/syzgy_oracle.c and
// csrc/exact.cuh; compiled with -ffp-contract=off like the rest of this file).
static double goXatan(double x) {
    const double P0 = -8.750608600031904122785e-01, P1 = -1.615753718733365076637e+01, P2 = -7.500855792314704667340e+01,
                 P3 = -1.228866684490136173410e+02, P4 = -6.485021904942025371773e+01;
    const double Q0 = +2.485846490142306297962e+01, Q1 = +1.650270098316988542046e+02, Q2 = +4.328810604912902668951e+02,
                 Q3 = +4.853903996359136964868e+02, Q4 = +1.945506571482613964425e+02;
    double z = x * x;
    z = z * ((((P0 * z + P1) * z + P2) * z + P3) * z + P4) / (((((z + Q0) * z + Q1) * z + Q2) * z + Q3) * z + Q4);
    return x * z + x;
}
static double goSatan(double x) {
    const double Morebits = 6.123233995736765886130e-17, Tan3pio8 = 2.41421356237309504880;
    if (x <= 0.66) return goXatan(x);
    if (x > Tan3pio8) return M_PI / 2 - goXatan(1 / x) + Morebits;
    return M_PI / 4 + goXatan((x - 1) / (x + 1)) + 0.5 * Morebits;
}
static double goAcos(double x) {
    double as;
    if (x == 0) as = x;
    else {
        const bool sign = x < 0;
        const double ax = sign ? -x : x;
        if (!(ax <= 1)) return std::nan("");
        double t = std::sqrt(1 - ax * ax);
        t = ax > 0.7 ? M_PI / 2 - goSatan(t / ax) : goSatan(ax / t);
        as = sign ? -t : t;
    }
    return M_PI / 2 - as;
}

// lshtree.go:59-77
void distanceToHyperplane(int method, const std::vector<double> &v, double length, const std::vector<double> &normal,
                          double b, double *dist, bool *right) {
    double d = dotProduct(v, normal) - b;
    *right = false;
    if (method == Euclidean) {
        if (d > 0) *right = true;
        else d = -d;
        *dist = d;
        return;
    }
    d = goAcos(d / length) / M_PI; // lshtree.go:71 math.Acos
    if (d > 0.5) { *right = true; d = 1 - d; }
    *dist = d;
}

struct NodeItem { LshNode *node; double priority; };
struct NodeLess { bool operator()(const NodeItem &x, const NodeItem &y) const { return x.priority > y.priority; } }; // lshtree.go:362-364

} // namespace

// ================================================================================ Collection
struct Collection::Impl {
    CollectionOptions opt;
    szg_index *gpu = nullptr;
    int rowbytes = 0;
    struct Rec { std::string meta; std::vector<uint8_t> codes; };
    std::unordered_map<uint64_t, Rec> store; // stands in for the span file (id -> {stream 0, stream 1})
    mutable std::shared_mutex mu;
    // lsh tree (newLSHTree(c, 100, 5), collection.go:292)
    std::vector<std::unique_ptr<LshNode>> roots;
    int threshold = 100;
    // one random source per tree: the reference inserts into its 5 trees on 5 goroutines (lshtree.go:101-114) -- sharing one
    // rand.Rand without a lock, which is why its trees are not reproducible; here every tree owns its stream, so the bulk
    // path can build the trees on 5 threads and still be deterministic for a given Seed
    struct TreeRng {
        std::mt19937_64 rng;
        std::normal_distribution<double> gauss{0.0, 1.0};
    };
    std::vector<TreeRng> trng;
    // test hooks
    // test hooks (LastVisitSequence / LastRescoreBatches): Search holds only the shared lock, so they are per calling thread
    static thread_local std::vector<uint64_t> last_visit;
    static thread_local int last_batches;

    std::vector<double> docVector(uint64_t id) const { // getDocument's decode (collection.go:470-484)
        auto it = store.find(id);
        if (it == store.end()) throw std::runtime_error("error getting document");
        return decodeVector(it->second.codes.data(), opt.DimensionCount, opt.Quantization);
    }

    std::vector<double> randomNormalizedVector(int dim, TreeRng &tr) { // lshtree.go:38-44, 10-28
        std::vector<double> v((size_t)dim);
        double norm = 0;
        for (auto &x : v) { x = tr.gauss(tr.rng); norm += x * x; }
        if (norm == 0) return v;
        norm = std::sqrt(norm);
        for (auto &x : v) x /= norm;
        return v;
    }

    std::unique_ptr<LshNode> split(std::unique_ptr<LshNode> node, TreeRng &tr) { // lshtree.go:172-248
        const size_t n = node->ids.size();
        const size_t i1 = (size_t)(tr.rng() % n);
        size_t i2;
        do { i2 = (size_t)(tr.rng() % n); } while (i2 == i1);
        const std::vector<double> v1 = docVector(node->ids[i1]), v2 = docVector(node->ids[i2]);
        bool same = true; // aboutEqual, tolerance 1e-9 (lshtree.go:158-170)
        for (size_t i = 0; i < v1.size(); ++i)
            if (std::fabs(v1[i] - v2[i]) > 1e-9) { same = false; break; }
        if (same) return node;
        std::vector<double> mid(v1.size());
        for (size_t i = 0; i < v1.size(); ++i) mid[i] = (v1[i] + v2[i]) / 2;
        std::vector<double> normal = randomNormalizedVector((int)mid.size(), tr);
        double b = 0;
        if (opt.DistanceMethod == Euclidean) b = std::sqrt(dotProduct(mid, mid));
        std::vector<uint64_t> leftIDs, rightIDs;
        for (uint64_t id : node->ids) {
            const std::vector<double> v = docVector(id);
            double dist;
            bool right;
            distanceToHyperplane(opt.DistanceMethod, v, vectorLength(v), normal, b, &dist, &right);
            (right ? rightIDs : leftIDs).push_back(id);
        }
        if (leftIDs.empty() || rightIDs.empty()) return node;
        auto parent = std::make_unique<LshNode>();
        parent->normal = std::move(normal);
        parent->b = b;
        parent->left = std::make_unique<LshNode>();
        parent->left->ids = std::move(leftIDs);
        parent->right = std::make_unique<LshNode>();
        parent->right->ids = std::move(rightIDs);
        return parent;
    }

    std::unique_ptr<LshNode> insert(std::unique_ptr<LshNode> node, uint64_t id, const std::vector<double> &v,
                                    double length, TreeRng &tr) { // lshtree.go:116-134
        if (node->isLeaf()) {
            node->ids.push_back(id);
            if ((int)node->ids.size() > threshold) node = split(std::move(node), tr);
            return node;
        }
        double dist;
        bool right;
        distanceToHyperplane(opt.DistanceMethod, v, length, node->normal, node->b, &dist, &right);
        if (!right) node->left = insert(std::move(node->left), id, v, length, tr);
        else node->right = insert(std::move(node->right), id, v, length, tr);
        return node;
    }

    std::unique_ptr<LshNode> remove(std::unique_ptr<LshNode> node, uint64_t id, const std::vector<double> &v,
                                    double length) { // lshtree.go:257-281
        if (!node) return node;
        if (node->isLeaf()) {
            auto it = std::find(node->ids.begin(), node->ids.end(), id);
            if (it != node->ids.end()) node->ids.erase(it);
            if (node->ids.empty()) return nullptr;
            return node;
        }
        double dist;
        bool right;
        distanceToHyperplane(opt.DistanceMethod, v, length, node->normal, node->b, &dist, &right);
        if (!right) node->left = remove(std::move(node->left), id, v, length);
        else node->right = remove(std::move(node->right), id, v, length);
        return node;
    }

    void addPoint(uint64_t id, const std::vector<double> &v) { // lshtree.go:101-114 (sequential here)
        const double length = vectorLength(v);
        for (size_t t = 0; t < roots.size(); ++t) roots[t] = insert(std::move(roots[t]), id, v, length, trng[t]);
    }
    // the reload loop of NewCollection (collection.go:298-311) for many documents: one thread per tree, like the
    // reference's goroutine per tree; the store is only read
    void addPoints(const std::vector<uint64_t> &ids, const std::vector<std::vector<double>> &vectors) {
        if (ids.size() < 4096) {
            for (size_t i = 0; i < ids.size(); ++i) addPoint(ids[i], vectors[i]);
            return;
        }
        std::vector<double> length(ids.size());
        for (size_t i = 0; i < ids.size(); ++i) length[i] = vectorLength(vectors[i]);
        std::vector<std::thread> th;
        for (size_t t = 0; t < roots.size(); ++t)
            th.emplace_back([&, t]() {
                for (size_t i = 0; i < ids.size(); ++i) roots[t] = insert(std::move(roots[t]), ids[i], vectors[i], length[i], trng[t]);
            });
        for (auto &x : th) x.join();
    }
    void removePoint(uint64_t id, const std::vector<double> &v) { // lshtree.go:250-255
        const double length = vectorLength(v);
        for (auto &root : roots) {
            root = remove(std::move(root), id, v, length);
            if (!root) root = std::make_unique<LshNode>();
        }
    }

    int buildMask(const FilterFn &filter) { // FilterFn outcome -> GPU bitmask (collection.go:592-594 applied per record)
        std::vector<uint64_t> ids;
        std::vector<uint8_t> pass;
        ids.reserve(store.size());
        pass.reserve(store.size());
        for (const auto &kv : store) {
            ids.push_back(kv.first);
            pass.push_back(filter(kv.first, kv.second.meta) ? 1 : 0);
        }
        int mask = -1;
        GPU(szg_mask_create(gpu, ids.data(), pass.data(), ids.size(), &mask), "mask_create");
        return mask;
    }

    SearchResults search(SearchArgs &args);
    SearchResults searchIndex(const SearchArgs &args);
};

Collection::Collection(const CollectionOptions &options) : p_(new Impl) {
    p_->opt = options;
    if (p_->opt.Quantization == 0) p_->opt.Quantization = 64; // collection.go:254-256
    p_->rowbytes = getVectorSize(p_->opt.Quantization, p_->opt.DimensionCount);
    if (p_->opt.DistanceMethod != Euclidean && p_->opt.DistanceMethod != Cosine)
        throw std::invalid_argument("Unsupported distance method"); // collection.go:281-282 panics
    if (szg_create(p_->opt.DimensionCount, p_->opt.Quantization, p_->opt.DistanceMethod, p_->opt.Device, &p_->gpu) != SZG_OK)
        throw std::runtime_error(std::string("syzgy_b200 create: ") + szg_last_error());
    p_->trng.resize(5);
    for (int i = 0; i < 5; ++i) {
        p_->trng[i].rng.seed(options.Seed * 0x9E3779B97F4A7C15ull + (uint64_t)i);
        p_->roots.push_back(std::make_unique<LshNode>());
    }
}

std::unique_ptr<Collection> Collection::Open(const std::string &path, int device, uint64_t seed) {
    szg_spanfile *sf = nullptr;
    if (szg_spanfile_open(path.c_str(), &sf) != SZG_OK) throw std::runtime_error(std::string("failed to open file: ") + szg_last_error());
    struct Closer { szg_spanfile *f; ~Closer() { szg_spanfile_close(f); } } closer{sf};
    szg_spanfile_info info;
    szg_spanfile_get_info(sf, &info);
    if (!info.has_header) throw std::runtime_error("failed to read header: record not found"); // collection.go:243-246
    CollectionOptions o;
    o.Name = info.name;
    o.DistanceMethod = info.distance_method;
    o.DimensionCount = info.dimension_count;
    o.Quantization = info.quantization;
    o.Device = device;
    o.Seed = seed;
    std::unique_ptr<Collection> c(new Collection(o));
    Impl &p = *c->p_;
    uint64_t loaded = 0;
    GPU(szg_spanfile_load(sf, p.gpu, &loaded), "spanfile_load");
    std::vector<uint64_t> ids((size_t)info.records);
    uint64_t n = 0;
    GPU(szg_spanfile_ids(sf, ids.data(), ids.size(), &n), "spanfile_ids");
    for (uint64_t id : ids) { // IterateSortedRecords order (spanfile.go:540-560)
        const uint8_t *vec = nullptr, *meta = nullptr;
        uint64_t vl = 0, ml = 0;
        GPU(szg_spanfile_record(sf, id, &vec, &vl, &meta, &ml), "spanfile_record");
        Impl::Rec rec{std::string(reinterpret_cast<const char *>(meta), (size_t)ml), std::vector<uint8_t>(vec, vec + vl)};
        const std::vector<double> v = decodeVector(rec.codes.data(), o.DimensionCount, o.Quantization);
        p.store[id] = std::move(rec);
        p.addPoint(id, v); // collection.go:304-305: the decoded vector
    }
    return c;
}

std::vector<SearchResults> Collection::SearchBatch(const std::vector<SearchArgs> &args) {
    std::vector<SearchResults> out(args.size());
    std::vector<size_t> batch; // exact, K > 0, no radius, no filter, one common K
    int K = 0;
    for (size_t i = 0; i < args.size(); ++i) {
        const SearchArgs &a = args[i];
        const bool fits = a.Precision == "exact" && a.K > 0 && a.K <= (int)SZG_MAX_K && a.Radius <= 0 && !a.Filter &&
                          (int)a.Vector.size() == p_->opt.DimensionCount && (K == 0 || a.K == K);
        if (fits) { K = a.K; batch.push_back(i); }
    }
    if (batch.size() >= 2) {
        std::shared_lock<std::shared_mutex> lk(p_->mu);
        const size_t nq = batch.size(), d = (size_t)p_->opt.DimensionCount;
        std::vector<double> flat(nq * d), dist(nq * (size_t)K);
        std::vector<uint64_t> ids(nq * (size_t)K);
        std::vector<uint32_t> cnt(nq);
        for (size_t j = 0; j < nq; ++j) std::copy(args[batch[j]].Vector.begin(), args[batch[j]].Vector.end(), flat.begin() + j * d);
        uint64_t scanned = 0;
        GPU(szg_search_batch(p_->gpu, flat.data(), (uint32_t)nq, (uint32_t)K, -1, SZG_F_DEFAULT, ids.data(), dist.data(), cnt.data(),
                             &scanned), "search_batch");
        const size_t numRecords = p_->store.size();
        for (size_t j = 0; j < nq; ++j) {
            SearchResults &r = out[batch[j]];
            for (uint32_t e = 0; e < cnt[j]; ++e) {
                const uint64_t id = ids[j * (size_t)K + e];
                r.Results.push_back(SearchResult{id, p_->store.at(id).meta, dist[j * (size_t)K + e]});
            }
            r.PercentSearched = numRecords == 0 ? 0.0 : (double)scanned / (double)numRecords * 100;
        }
    } else {
        batch.clear();
    }
    size_t b = 0;
    for (size_t i = 0; i < args.size(); ++i) {
        if (b < batch.size() && batch[b] == i) { ++b; continue; }
        out[i] = Search(args[i]);
    }
    return out;
}

Collection::~Collection() { Close(); }

void Collection::Close() {
    std::unique_lock<std::shared_mutex> lk(p_->mu);
    if (p_->gpu) { szg_destroy(p_->gpu); p_->gpu = nullptr; }
}

const CollectionOptions &Collection::Options() const { return p_->opt; }
thread_local std::vector<uint64_t> Collection::Impl::last_visit;
thread_local int Collection::Impl::last_batches = 0;
const std::vector<uint64_t> &Collection::LastVisitSequence() const { return Impl::last_visit; }
int Collection::LastRescoreBatches() const { return Impl::last_batches; }

void Collection::AddDocument(uint64_t id, const std::vector<double> &vector, const std::string &metadata) {
    std::unique_lock<std::shared_mutex> lk(p_->mu);
    if ((int)vector.size() != p_->opt.DimensionCount)
        throw std::invalid_argument("vector size does not match the expected number of dimensions"); // 431-434 panics
    Impl::Rec rec{metadata, encodeDocument(vector, p_->opt.Quantization)};
    GPU(szg_upsert(p_->gpu, &id, rec.codes.data(), 1), "upsert"); // right after WriteRecord (446-453)
    p_->store[id] = std::move(rec);
    p_->addPoint(id, vector); // 456: the raw vector, not the decoded one
}

void Collection::AddDocuments(const std::vector<uint64_t> &ids, const std::vector<std::vector<double>> &vectors,
                              const std::vector<std::string> &metadata) {
    std::unique_lock<std::shared_mutex> lk(p_->mu);
    if (ids.size() != vectors.size() || (!metadata.empty() && metadata.size() != ids.size()))
        throw std::invalid_argument("ids, vectors and metadata must have the same length");
    // bulk ingest: encodeDocument runs on the device (szg_encode); the bytes come back for the record store
    const size_t d = (size_t)p_->opt.DimensionCount;
    std::vector<double> flat(d * ids.size());
    for (size_t i = 0; i < ids.size(); ++i) {
        if (vectors[i].size() != d)
            throw std::invalid_argument("vector size does not match the expected number of dimensions");
        std::memcpy(flat.data() + i * d, vectors[i].data(), d * sizeof(double));
    }
    std::vector<uint8_t> all((size_t)p_->rowbytes * ids.size());
    GPU(szg_encode(p_->gpu, ids.data(), flat.data(), ids.size(), all.data(), 1), "encode");
    for (size_t i = 0; i < ids.size(); ++i) {
        const uint8_t *row = all.data() + i * (size_t)p_->rowbytes;
        p_->store[ids[i]] = Impl::Rec{metadata.empty() ? std::string() : metadata[i], std::vector<uint8_t>(row, row + p_->rowbytes)};
    }
    p_->addPoints(ids, vectors);
}

bool Collection::GetDocument(uint64_t id, Document *out) const {
    std::shared_lock<std::shared_mutex> lk(p_->mu);
    auto it = p_->store.find(id);
    if (it == p_->store.end()) return false;
    if (out) {
        out->ID = id;
        out->Metadata = it->second.meta;
        out->Vector = decodeVector(it->second.codes.data(), p_->opt.DimensionCount, p_->opt.Quantization);
    }
    return true;
}

bool Collection::UpdateDocument(uint64_t id, const std::string &metadata) { // vector unchanged: the mirror is untouched
    std::unique_lock<std::shared_mutex> lk(p_->mu);
    auto it = p_->store.find(id);
    if (it == p_->store.end()) return false;
    it->second.meta = metadata;
    return true;
}

bool Collection::removeDocument(uint64_t id) {
    std::unique_lock<std::shared_mutex> lk(p_->mu);
    auto it = p_->store.find(id);
    if (it == p_->store.end()) return false;
    p_->removePoint(id, decodeVector(it->second.codes.data(), p_->opt.DimensionCount, p_->opt.Quantization));
    GPU(szg_remove(p_->gpu, &id, 1, nullptr), "remove");
    p_->store.erase(it);
    return true;
}

int Collection::GetDocumentCount() const {
    std::shared_lock<std::shared_mutex> lk(p_->mu);
    return (int)p_->store.size();
}

SearchResults Collection::Search(SearchArgs args) {
    std::shared_lock<std::shared_mutex> lk(p_->mu); // RLock, collection.go:570
    return p_->search(args);
}

SearchResults Collection::Impl::search(SearchArgs &args) {
    if (args.Precision.empty()) args.Precision = "medium"; // 573-575
    SearchResults ret;
    const size_t numRecords = store.size();
    size_t pointsSearched = 0;

    if (args.Radius == 0 && args.K == 0) { // list mode, 633-668
        std::vector<uint64_t> ids;
        ids.reserve(numRecords);
        for (const auto &kv : store) ids.push_back(kv.first);
        std::sort(ids.begin(), ids.end(), lex_less); // IterateSortedRecords
        for (uint64_t id : ids) {
            const std::string &meta = store.at(id).meta;
            if (args.Filter && !args.Filter(id, meta)) continue;
            ++pointsSearched;
            if (args.Offset > 0 && (int)pointsSearched <= args.Offset) continue;
            ret.Results.push_back(SearchResult{id, meta, 0.0});
            if (args.Limit > 0 && (int)ret.Results.size() >= args.Limit) break;
        }
    } else {
        if ((int)args.Vector.size() != opt.DimensionCount) // appendix B-12: the reference would read out of bounds
            throw std::invalid_argument("query dimension does not match the collection");
        if (args.Precision == "exact") { // 672-684 -> one GPU scan
            int mask = -1;
            if (args.Filter) mask = buildMask(args.Filter);
            std::vector<uint64_t> ids;
            std::vector<double> dist;
            uint64_t scanned = 0;
            if (args.Radius > 0) { // Radius overrides K, inclusive (598-605)
                szg_result *r = nullptr;
                GPU(szg_search_radius(gpu, args.Vector.data(), args.Radius, mask, SZG_F_DEFAULT, &r, &scanned), "search_radius");
                uint64_t n = 0;
                szg_result_count(r, &n);
                ids.resize(n);
                dist.resize(n);
                if (n) szg_result_fetch(r, 0, n, ids.data(), dist.data());
                szg_result_free(r);
            } else {
                if (args.K > (int)SZG_MAX_K) throw std::invalid_argument("K exceeds SZG_MAX_K on the GPU path");
                ids.resize((size_t)args.K);
                dist.resize((size_t)args.K);
                uint32_t n = 0;
                GPU(szg_search_topk(gpu, args.Vector.data(), 1, (uint32_t)args.K, mask, SZG_F_DEFAULT, ids.data(), dist.data(),
                                    &n, &scanned), "search_topk");
                ids.resize(n);
                dist.resize(n);
            }
            if (mask >= 0) szg_mask_destroy(gpu, mask);
            pointsSearched = scanned; // filtered rows count as searched (589 precedes 592)
            for (size_t i = 0; i < ids.size(); ++i) ret.Results.push_back(SearchResult{ids[i], store.at(ids[i]).meta, dist[i]});
        } else {
            SearchResults r = searchIndex(args);
            ret.Results = std::move(r.Results);
            pointsSearched = (size_t)r.PercentSearched; // carries the count, converted below
        }
    }
    ret.PercentSearched = numRecords == 0 ? 0.0 : (double)pointsSearched / (double)numRecords * 100; // 700-710
    return ret;
}

// lshTree.search (lshtree.go:283-351) driving `consider` (collection.go:583-629), with the distances
// computed on the GPU.  Only leaves are ever pruned and inner nodes always expand, so the order in
// which nodes pop does not depend on distances: the host pops ahead, gathers the ids of upcoming
// leaves, rescoring them in one szg_rescore call, then replays the reference's loop over them.
SearchResults Collection::Impl::searchIndex(const SearchArgs &args) {
    const std::vector<double> &vector = args.Vector;
    const double length = vectorLength(vector);
    double radius = args.Radius > 0 ? args.Radius : 1.7976931348623157e308; // 686-689
    std::unordered_set<uint64_t> visited;
    const int search_k = 200;
    int k_counter = 0;
    bool pointAccepted = false;
    size_t pointsSearched = 0;
    GoHeap<ResultItem, ResultLess> results;
    GoHeap<NodeItem, NodeLess> pq;
    for (auto &root : roots) pq.push(NodeItem{root.get(), 0.0}); // 295-297
    last_visit.clear();
    last_batches = 0;

    std::unordered_map<uint64_t, double> cache; // id -> GPU distance
    struct Leaf { LshNode *node; double priority; };
    std::vector<Leaf> ahead;
    size_t ahead_pos = 0;
    const size_t kBatchIds = 2048;
    bool stop = false;

    auto consider = [&](uint64_t id, double d, double *rad) -> int { // collection.go:583-629
        if (d == SZG_MISSING_DISTANCE) return StopSearch;             // getDocument failed (585-587)
        auto it = store.find(id);
        if (it == store.end()) return StopSearch;
        ++pointsSearched;
        if (args.Filter && !args.Filter(id, it->second.meta)) return PointIgnored;
        if (args.Radius > 0 && d <= args.Radius) {
            results.push(ResultItem{id, d});
            return PointAccepted;
        } else if (args.Radius > 0) {
            return PointChecked;
        } else if (args.K > 0) {
            if ((int)results.size() <= args.K) {
                if ((int)results.size() < args.K || results.a[0].priority > d) {
                    results.push(ResultItem{id, d});
                    if ((int)results.size() > args.K) results.pop();
                    *rad = results.a[0].priority;
                    return PointAccepted;
                }
            }
        }
        return PointChecked;
    };

    while (!stop) {
        // ---- pop ahead: leaves in the reference's pop order, until a batch of unseen ids is gathered
        if (ahead_pos == ahead.size()) {
            ahead.clear();
            ahead_pos = 0;
            std::vector<uint64_t> need;
            std::unordered_set<uint64_t> in_batch;
            while (pq.size() > 0 && need.size() < kBatchIds) {
                NodeItem item = pq.pop();
                LshNode *node = item.node;
                if (node->isLeaf()) {
                    ahead.push_back(Leaf{node, item.priority});
                    for (uint64_t id : node->ids)
                        if (!cache.count(id) && in_batch.insert(id).second) need.push_back(id);
                } else { // 337-348
                    double dist;
                    bool right;
                    distanceToHyperplane(opt.DistanceMethod, vector, length, node->normal, node->b, &dist, &right);
                    if (right) {
                        pq.push(NodeItem{node->right.get(), dist});
                        pq.push(NodeItem{node->left.get(), -dist});
                    } else {
                        pq.push(NodeItem{node->left.get(), dist});
                        pq.push(NodeItem{node->right.get(), -dist});
                    }
                }
            }
            if (ahead.empty()) break; // queue exhausted
            if (!need.empty()) {
                std::vector<double> dist(need.size());
                GPU(szg_rescore(gpu, vector.data(), need.data(), need.size(), dist.data()), "rescore");
                ++last_batches;
                for (size_t i = 0; i < need.size(); ++i) cache.emplace(need[i], dist[i]);
            }
        }
        // ---- replay lshtree.go:299-336 over the next leaf
        const Leaf lf = ahead[ahead_pos++];
        if (lf.priority < 0 && -lf.priority > radius) continue; // far side of the hyperplane, beyond the radius (304-309)
        if (k_counter >= search_k) break;                        // 311-313
        for (uint64_t id : lf.node->ids) {
            if (visited.count(id)) continue;
            visited.insert(id);
            last_visit.push_back(id);
            const int signal = consider(id, cache.at(id), &radius);
            if (signal == StopSearch) { stop = true; break; }
            if (signal == PointAccepted) { k_counter = 0; pointAccepted = true; }
            else if (signal == PointChecked) { if (pointAccepted) ++k_counter; }
        }
    }

    SearchResults out;
    out.Results.resize(results.size()); // 693-697: pop back to front => ascending
    for (size_t i = out.Results.size(); i-- > 0;) {
        ResultItem it = results.pop();
        out.Results[i] = SearchResult{it.id, store.at(it.id).meta, it.priority};
    }
    out.PercentSearched = (double)pointsSearched; // the caller converts the count
    return out;
}

} // namespace syzgydb
