// test_collection.cpp -- the reference's own Collection tests (collection_test.go, rest_test.go list mode)
// restated against the C++ host mirror (syzgydb_b200/host), plus parity of the GPU-backed results with
// the CPU oracle.  `--cpu` runs only the codec checks (no GPU needed).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <random>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "collection.hpp"

extern "C" { // oracle/syzgy_oracle.c (test infrastructure)
uint64_t orc_quantize(double value, int bits);
double orc_dequantize(uint64_t value, int bits);
void orc_encode(const double *vec, int64_t dims, int bits, uint8_t *data);
void orc_decode(const uint8_t *data, int64_t dims, int bits, double *vec);
int64_t orc_search_exact(const uint8_t *codes, const uint64_t *ids, int64_t nrows, int64_t dims, int bits, int metric,
                         const double *query, int64_t k, double radius, const uint8_t *pass, const int64_t *order,
                         int faithful, uint64_t *out_ids, double *out_dist, int64_t out_cap, double *percent_searched);
int64_t orc_replay(const uint8_t *codes, const uint64_t *ids, int64_t nrows, int64_t dims, int bits, int metric,
                   const double *query, int64_t k, double radius, const uint8_t *pass, const int64_t *visit,
                   int64_t nvisit, uint64_t *out_ids, double *out_dist, int64_t out_cap, int64_t *points_searched);
void orc_lex_order(const uint64_t *ids, int64_t n, int64_t *perm);
}

using namespace syzgydb;

static int g_fail = 0;
#define CHECK(cond, ...)                                                   \
    do {                                                                   \
        if (!(cond)) {                                                     \
            std::printf("  FAIL %s:%d: ", __FILE__, __LINE__);             \
            std::printf(__VA_ARGS__);                                      \
            std::printf("\n");                                             \
            ++g_fail;                                                      \
            return;                                                        \
        }                                                                  \
    } while (0)

static bool close_rel(double a, double b, double rtol) { return std::fabs(a - b) <= rtol * std::fmax(std::fabs(b), 1e-300) + 1e-300; }

// ------------------------------------------------------------------ codec vs oracle (CPU)
static void TestCodec() {
    std::mt19937_64 rng(3);
    std::uniform_real_distribution<double> u(-1.6, 1.6);
    const int bitsv[] = {4, 8, 16, 32, 64};
    for (int bits : bitsv) {
        for (int i = 0; i < 20000; ++i) {
            double v = u(rng);
            CHECK(quantize(v, bits) == orc_quantize(v, bits), "quantize(%a, %d)", v, bits);
            uint64_t q = quantize(v, bits);
            double a = dequantize(q, bits), b = orc_dequantize(q, bits);
            CHECK(std::memcmp(&a, &b, 8) == 0, "dequantize(%llu, %d)", (unsigned long long)q, bits);
        }
        for (int dims : {1, 7, 33}) {
            std::vector<double> vec((size_t)dims);
            for (auto &x : vec) x = u(rng);
            std::vector<uint8_t> mine = encodeDocument(vec, bits), ref(mine.size());
            orc_encode(vec.data(), dims, bits, ref.data());
            CHECK(mine == ref, "encodeDocument bits=%d dims=%d", bits, dims);
            std::vector<double> dm = decodeVector(mine.data(), dims, bits), dr((size_t)dims);
            orc_decode(ref.data(), dims, bits, dr.data());
            CHECK(std::memcmp(dm.data(), dr.data(), (size_t)dims * 8) == 0, "decodeVector bits=%d dims=%d", bits, dims);
        }
    }
    CHECK(getVectorSize(4, 7) == 4 && getVectorSize(16, 7) == 14, "getVectorSize");
    bool threw = false;
    try { getVectorSize(12, 7); } catch (const std::invalid_argument &) { threw = true; }
    CHECK(threw, "unsupported quantization must throw (the reference panics, collection.go:809)");
}

// ------------------------------------------------------------------ collection_test.go:12-21
static void TestEuclideanDistance() {
    CollectionOptions o; o.Name = "kat"; o.DistanceMethod = Euclidean; o.DimensionCount = 3; o.Quantization = 64;
    Collection c(o);
    c.AddDocument(1, {4, 5, 6}, "");
    SearchArgs a; a.Vector = {1, 2, 3}; a.K = 1; a.Precision = "exact";
    SearchResults r = c.Search(a);
    CHECK(r.Results.size() == 1 && r.Results[0].Distance == 5.196152422706632, "expected 5.196152422706632, got %.17g",
          r.Results.empty() ? -1.0 : r.Results[0].Distance);
}

// ------------------------------------------------------------------ collection_test.go:283-382
static void TestCollectionSearch() {
    CollectionOptions o; o.Name = "search"; o.DistanceMethod = Euclidean; o.DimensionCount = 2; o.Quantization = 64;
    Collection c(o);
    SearchArgs a; a.Vector = {50, 50}; a.K = 5;
    CHECK(c.Search(a).Results.empty(), "empty collection must give no results");
    std::mt19937_64 rng(7);
    std::uniform_real_distribution<double> u(0, 100);
    for (uint64_t i = 0; i < 10; ++i) c.AddDocument(i, {u(rng), u(rng)}, "metadata");
    a.K = 3;
    SearchResults r = c.Search(a);
    CHECK(r.Results.size() <= 3 && !r.Results.empty(), "expected at most 3 results, got %zu", r.Results.size());
    SearchArgs ra; ra.Vector = {50, 50}; ra.Radius = 30; // radius overrides K
    ra.K = 1;
    r = c.Search(ra);
    for (const auto &x : r.Results) CHECK(x.Distance <= 30.0, "radius result beyond the radius: %g", x.Distance);
    SearchArgs fa; fa.Vector = {50, 50}; fa.K = 5;
    fa.Filter = [](uint64_t id, const std::string &) { return id % 2 == 0; };
    for (const char *prec : {"", "exact"}) {
        fa.Precision = prec;
        r = c.Search(fa);
        CHECK(!r.Results.empty(), "filter search returned nothing");
        for (const auto &x : r.Results) CHECK(x.ID % 2 == 0, "expected only even ids, got %llu", (unsigned long long)x.ID);
    }
}

// ------------------------------------------------------------------ collection_test.go:549-612
static void TestExhaustiveSearch() {
    CollectionOptions o; o.Name = "exh"; o.DistanceMethod = Euclidean; o.DimensionCount = 3; o.Quantization = 64;
    Collection c(o);
    c.AddDocument(1, {1, 2, 3}, "doc1");
    c.AddDocument(2, {4, 5, 6}, "doc2");
    c.AddDocument(3, {7, 8, 9}, "doc3");
    SearchArgs a; a.Vector = {1, 2, 3}; a.K = 3; a.Precision = "exact";
    SearchResults r = c.Search(a);
    CHECK(r.Results.size() == 3, "expected 3 results, got %zu", r.Results.size());
    CHECK(r.Results[0].ID == 1 && r.Results[1].ID == 2 && r.Results[2].ID == 3, "expected ids 1,2,3 in order");
    CHECK(r.Results[1].Metadata == "doc2", "metadata must come back with the result");
    CHECK(r.PercentSearched == 100.0, "PercentSearched %g", r.PercentSearched);
}

// ------------------------------------------------------------------ collection_test.go:614-667
static void TestVectorSearchWith4BitQuantization() {
    CollectionOptions o; o.Name = "q4"; o.DistanceMethod = Euclidean; o.DimensionCount = 4; o.Quantization = 4;
    Collection c(o);
    std::mt19937_64 rng(11);
    std::uniform_real_distribution<double> u(-1, 1);
    for (uint64_t i = 0; i < 10; ++i) c.AddDocument(i, {u(rng), u(rng), u(rng), u(rng)}, "m");
    SearchArgs a; a.Vector = {0.1, -0.2, 0.3, 0.4}; a.K = 5;
    SearchResults r = c.Search(a);
    CHECK(!r.Results.empty() && r.Results.size() <= 5, "4-bit search returned %zu results", r.Results.size());
    for (size_t i = 1; i < r.Results.size(); ++i) CHECK(r.Results[i - 1].Distance <= r.Results[i].Distance, "not ascending");
}

// ------------------------------------------------------------------ collection_test.go:145-281, 384-534 (CRUD side)
static void TestDocumentRoundTripUpdateRemove() {
    CollectionOptions o; o.Name = "crud"; o.DistanceMethod = Cosine; o.DimensionCount = 5; o.Quantization = 64;
    Collection c(o);
    std::vector<double> v = {0.25, -3.5, 1e-9, 7.0, 2.0};
    c.AddDocument(42, v, "original");
    Document d;
    CHECK(c.GetDocument(42, &d) && d.Vector == v && d.Metadata == "original", "fp64 vectors must round-trip exactly");
    CHECK(c.UpdateDocument(42, "updated") && c.GetDocument(42, &d) && d.Metadata == "updated" && d.Vector == v, "update");
    CHECK(!c.UpdateDocument(43, "x") && !c.GetDocument(43, nullptr), "missing id");
    c.AddDocument(7, {1, 0, 0, 0, 0}, "seven");
    CHECK(c.GetDocumentCount() == 2, "count");
    SearchArgs a; a.Vector = v; a.K = 2; a.Precision = "exact";
    CHECK(c.Search(a).Results.size() == 2, "two docs");
    CHECK(c.removeDocument(42) && !c.removeDocument(42) && c.GetDocumentCount() == 1, "remove");
    SearchResults r = c.Search(a);
    CHECK(r.Results.size() == 1 && r.Results[0].ID == 7, "removed document must not be returned");
    a.Precision = "";
    r = c.Search(a);
    CHECK(r.Results.size() == 1 && r.Results[0].ID == 7, "removed document must not be returned by the index path");
    bool threw = false;
    try { c.AddDocument(9, {1, 2}, ""); } catch (const std::invalid_argument &) { threw = true; }
    CHECK(threw, "dimension mismatch must throw (the reference panics, collection.go:431-434)");
}

// ------------------------------------------------------------------ list mode, collection.go:633-668 / rest_test.go:72-148
static void TestListModePagination() {
    CollectionOptions o; o.Name = "list"; o.DistanceMethod = Euclidean; o.DimensionCount = 2; o.Quantization = 8;
    Collection c(o);
    for (uint64_t id : {2, 10, 1, 33, 3, 100}) c.AddDocument(id, {0.1, 0.2}, "m" + std::to_string(id));
    SearchArgs a; // K == 0 && Radius == 0
    SearchResults r = c.Search(a);
    std::vector<uint64_t> want = {1, 10, 100, 2, 3, 33}; // sort.Strings order
    CHECK(r.Results.size() == want.size(), "list size");
    for (size_t i = 0; i < want.size(); ++i) CHECK(r.Results[i].ID == want[i], "list order at %zu", i);
    a.Offset = 2; a.Limit = 3;
    r = c.Search(a);
    CHECK(r.Results.size() == 3 && r.Results[0].ID == 100 && r.Results[2].ID == 3, "offset/limit");
    a.Offset = 0; a.Limit = 0;
    a.Filter = [](uint64_t id, const std::string &) { return id >= 10; };
    r = c.Search(a);
    CHECK(r.Results.size() == 3 && r.PercentSearched == 50.0, "filter in list mode: %zu results, %g%%", r.Results.size(), r.PercentSearched);
}

// ------------------------------------------------------------------ collection_test.go:23-103 + oracle parity (appendix B-13)
struct Corpus {
    std::vector<uint64_t> ids;
    std::vector<std::vector<double>> vecs;
    std::vector<uint8_t> codes; // row-major stream-1 bytes
    std::unordered_map<uint64_t, int64_t> row_of;
};
static Corpus make_corpus(int n, int dims, int bits, uint64_t seed, bool gaussian) {
    Corpus c;
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> u(-1, 1);
    std::normal_distribution<double> g(0, 1);
    const int rb = getVectorSize(bits, dims);
    c.codes.resize((size_t)n * rb);
    for (int i = 0; i < n; ++i) {
        std::vector<double> v((size_t)dims);
        for (auto &x : v) x = gaussian ? g(rng) : u(rng);
        uint64_t id = (uint64_t)i * 3 + 1;
        c.ids.push_back(id);
        std::vector<uint8_t> e = encodeDocument(v, bits);
        std::memcpy(c.codes.data() + (size_t)i * rb, e.data(), (size_t)rb);
        c.row_of[id] = i;
        c.vecs.push_back(std::move(v));
    }
    return c;
}

struct LshVariant { int k; double radius; bool filter; };
static void lsh_queries(const char *name, Collection &c, const Corpus &cp, int n, int dims, int bits, int metric, int k, double radius,
                        bool filter);
// one collection, one or more (k, radius, filter) variants of the same three queries
static void lsh_cases(const char *name, int n, int dims, int bits, int metric, const std::vector<LshVariant> &variants) {
    Corpus cp = make_corpus(n, dims, bits, 1234 + n + dims, metric == Cosine);
    CollectionOptions o; o.Name = name; o.DistanceMethod = metric; o.DimensionCount = dims; o.Quantization = bits; o.Seed = 99;
    Collection c(o);
    std::vector<std::string> meta;
    for (uint64_t id : cp.ids) meta.push_back("{\"bucket\": " + std::to_string(id % 10) + "}");
    c.AddDocuments(cp.ids, cp.vecs, meta);
    std::vector<std::vector<double>>().swap(cp.vecs); // the collection keeps the encoded rows; the oracle needs cp.codes only
    for (const LshVariant &v : variants) {
        const int before = g_fail;
        lsh_queries(name, c, cp, n, dims, bits, metric, v.k, v.radius, v.filter);
        if (variants.size() > 1) std::printf("%s %s k=%d radius=%g filter=%d\n", g_fail == before ? "PASS" : "FAILED", name, v.k, v.radius, (int)v.filter);
    }
}
static void lsh_case(const char *name, int n, int dims, int bits, int metric, int k, double radius, bool filter) {
    lsh_cases(name, n, dims, bits, metric, {LshVariant{k, radius, filter}});
}
static void lsh_queries(const char *name, Collection &c, const Corpus &cp, int n, int dims, int bits, int metric, int k, double radius,
                        bool filter) {
    std::mt19937_64 rng(5);
    std::uniform_real_distribution<double> u(-1, 1);
    std::vector<uint8_t> pass;
    FilterFn fn;
    if (filter) {
        fn = [](uint64_t id, const std::string &) { return id % 10 < 3; }; // bucket < 3: 30 % density (cfg3)
        for (uint64_t id : cp.ids) pass.push_back(id % 10 < 3);
    }
    for (int qi = 0; qi < 3; ++qi) {
        SearchArgs a;
        a.Vector.resize((size_t)dims);
        for (auto &x : a.Vector) x = u(rng);
        a.K = k; a.Radius = radius; a.Filter = fn;
        SearchResults lsh = c.Search(a); // Precision "" -> medium -> the tree
        const std::vector<uint64_t> &visit_ids = c.LastVisitSequence();
        CHECK(lsh.PercentSearched < 100.0 && lsh.PercentSearched > 0.0, "%s: PercentSearched %g", name, lsh.PercentSearched);
        CHECK(std::fabs(lsh.PercentSearched - 100.0 * (double)visit_ids.size() / n) < 1e-9, "%s: PercentSearched vs visits", name);
        // oracle: `consider` replayed over the same visit sequence with CPU distances
        std::vector<int64_t> visit;
        for (uint64_t id : visit_ids) visit.push_back(cp.row_of.at(id));
        std::vector<uint64_t> oi(visit.size() + 1);
        std::vector<double> od(visit.size() + 1);
        int64_t ps = 0;
        int64_t m = orc_replay(cp.codes.data(), cp.ids.data(), n, dims, bits, metric, a.Vector.data(), k, radius,
                               filter ? pass.data() : nullptr, visit.data(), (int64_t)visit.size(), oi.data(), od.data(),
                               (int64_t)oi.size(), &ps);
        CHECK((size_t)m == lsh.Results.size(), "%s q%d: %zu results, oracle replay has %lld", name, qi, lsh.Results.size(), (long long)m);
        for (int64_t i = 0; i < m; ++i) {
            CHECK(lsh.Results[(size_t)i].ID == oi[(size_t)i], "%s q%d rank %lld: id %llu vs oracle %llu", name, qi, (long long)i,
                  (unsigned long long)lsh.Results[(size_t)i].ID, (unsigned long long)oi[(size_t)i]);
            CHECK(close_rel(lsh.Results[(size_t)i].Distance, od[(size_t)i], 1e-13), "%s q%d rank %lld: distance %.17g vs %.17g", name, qi,
                  (long long)i, lsh.Results[(size_t)i].Distance, od[(size_t)i]);
        }
        CHECK(c.LastRescoreBatches() >= 1 && c.LastRescoreBatches() <= 1 + (int)visit_ids.size() / 1024 + 2, "%s: %d rescoring batches for %zu visits",
              name, c.LastRescoreBatches(), visit_ids.size());
        // exact search: same count for k-mode (collection_test.go:88-91), ids/dist equal to the oracle's scan
        a.Precision = "exact";
        SearchResults ex = c.Search(a);
        std::vector<int64_t> order((size_t)n);
        orc_lex_order(cp.ids.data(), n, order.data());
        std::vector<uint64_t> ei((size_t)n);
        std::vector<double> ed((size_t)n);
        double pct = 0;
        int64_t em = orc_search_exact(cp.codes.data(), cp.ids.data(), n, dims, bits, metric, a.Vector.data(), k, radius,
                                      filter ? pass.data() : nullptr, order.data(), 0, ei.data(), ed.data(), n, &pct);
        CHECK((size_t)em == ex.Results.size() && ex.PercentSearched == 100.0, "%s q%d exact: %zu results vs oracle %lld", name, qi,
              ex.Results.size(), (long long)em);
        if (radius == 0) {
            CHECK(lsh.Results.size() == ex.Results.size(), "%s: LSH and exact result counts differ", name);
            for (int64_t i = 0; i < em; ++i) {
                CHECK(close_rel(ex.Results[(size_t)i].Distance, ed[(size_t)i], 1e-12), "%s exact rank %lld distance", name, (long long)i);
                CHECK(ex.Results[(size_t)i].ID == ei[(size_t)i] || close_rel(ex.Results[(size_t)i].Distance, ed[(size_t)i], 1e-5),
                      "%s exact rank %lld id", name, (long long)i);
            }
            for (size_t i = 0; i < lsh.Results.size(); ++i) // an LSH result can never beat the exact one at the same rank
                CHECK(lsh.Results[i].Distance >= ex.Results[i].Distance * (1 - 1e-12), "%s: LSH better than exact at rank %zu", name, i);
        }
    }
}

static void TestSearchExactVsLSH() { lsh_case("lsh_cos_f64", 20000, 3, 64, Cosine, 10, 0, false); }
static void TestLSHQuantized() {
    lsh_case("lsh_euc_q8", 12000, 24, 8, Euclidean, 10, 0, false);
    lsh_case("lsh_cos_q4", 8000, 16, 4, Cosine, 5, 0, true);
    lsh_case("lsh_euc_q16", 6000, 10, 16, Euclidean, 20, 0, false);
}
static void TestLSHRadiusWithFilter() { lsh_case("lsh_cos_f64_radius_filter", 30000, 48, 64, Cosine, 0, 0.46, true); } // cfg3 shape, small

// --open <file.dat> <k> <q_0,0> <q_0,1> ... : opens a collection file written by the reference (or its restated writer),
// prints its options and document count, then the exact top-k of the queries found on the command line (one
// line per result: query index, id, distance as a hex float, metadata) answered as ONE SearchBatch and, for the
// first query, once more through Search.  tests/test_host_cpp.py compares the output with the oracle.
static int open_mode(int argc, char **argv) {
    const std::string path = argv[2];
    const int k = std::atoi(argv[3]);
    auto c = Collection::Open(path);
    const CollectionOptions &o = c->Options();
    std::printf("OPTIONS %s %d %d %d\n", o.Name.c_str(), o.DistanceMethod, o.DimensionCount, o.Quantization);
    std::printf("COUNT %d\n", c->GetDocumentCount());
    std::vector<SearchArgs> batch;
    for (int at = 4; at + o.DimensionCount <= argc; at += o.DimensionCount) {
        SearchArgs a;
        a.K = k;
        a.Precision = "exact";
        for (int i = 0; i < o.DimensionCount; ++i) a.Vector.push_back(std::strtod(argv[at + i], nullptr));
        batch.push_back(a);
    }
    const auto res = c->SearchBatch(batch);
    for (size_t q = 0; q < res.size(); ++q)
        for (const auto &r : res[q].Results)
            std::printf("RESULT %zu %llu %a %s\n", q, (unsigned long long)r.ID, r.Distance, r.Metadata.c_str());
    const SearchResults one = c->Search(batch[0]);
    for (const auto &r : one.Results) std::printf("SINGLE 0 %llu %a %s\n", (unsigned long long)r.ID, r.Distance, r.Metadata.c_str());
    std::printf("PERCENT %.6f\n", res[0].PercentSearched);
    // the medium-precision path over the rebuilt LSH trees still answers
    SearchArgs m = batch[0];
    m.Precision = "";
    const SearchResults lsh = c->Search(m);
    std::printf("LSH %zu %.6f\n", lsh.Results.size(), lsh.PercentSearched);
    return 0;
}

// --cfg3-bench [rows]: BASELINE.json configs[2] timed through the host mirror -- what Precision "medium" (LSH-tree walk on the
// host, candidates re-scored on the GPU in speculative batches, `consider` replayed) costs next to Precision "exact" (one GPU
// scan) on the same collection and queries.  One JSON line; bench.py embeds it (is a device-side tree worth it?).
static int cfg3_bench(int n) {
    using clk = std::chrono::steady_clock;
    auto ms = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    const int dims = 384, nq = 20;
    Corpus cp = make_corpus(n, dims, 64, 1234 + n + dims, true);
    CollectionOptions o; o.Name = "cfg3"; o.DistanceMethod = Cosine; o.DimensionCount = dims; o.Quantization = 64; o.Seed = 99;
    Collection c(o);
    std::vector<std::string> meta;
    for (uint64_t id : cp.ids) meta.push_back("{\"bucket\": " + std::to_string(id % 10) + "}");
    auto t0 = clk::now();
    c.AddDocuments(cp.ids, cp.vecs, meta);
    const double build_ms = ms(t0, clk::now());
    std::vector<std::vector<double>>().swap(cp.vecs);
    std::mt19937_64 rng(5);
    std::normal_distribution<double> g(0, 1);
    std::vector<SearchArgs> qs((size_t)nq);
    for (auto &a : qs) { a.Vector.resize((size_t)dims); for (auto &x : a.Vector) x = g(rng); a.K = 10; }
    for (int w = 0; w < 2; ++w) { SearchArgs a = qs[(size_t)w]; c.Search(a); a.Precision = "exact"; c.Search(a); }
    double med_ms = 0, pct = 0, batches = 0, ex_ms = 0, agree = 0;
    for (auto &a : qs) {
        SearchArgs m = a;
        auto t1 = clk::now();
        SearchResults r = c.Search(m);
        med_ms += ms(t1, clk::now());
        pct += r.PercentSearched;
        batches += c.LastRescoreBatches();
        SearchArgs e = a;
        e.Precision = "exact";
        t1 = clk::now();
        SearchResults x = c.Search(e);
        ex_ms += ms(t1, clk::now());
        std::unordered_set<uint64_t> truth;
        for (auto &h : x.Results) truth.insert(h.ID);
        int hit = 0;
        for (auto &h : r.Results) hit += (int)truth.count(h.ID);
        agree += x.Results.empty() ? 1.0 : (double)hit / (double)x.Results.size();
    }
    std::printf("{\"rows\": %d, \"dims\": %d, \"queries\": %d, \"k\": 10, \"tree_build_ms_5_threads\": %.1f, "
                "\"medium_ms_per_query\": %.3f, \"medium_percent_searched\": %.3f, \"medium_rescore_batches_per_query\": %.2f, "
                "\"medium_recall_at_10_vs_exact\": %.3f, \"exact_ms_per_query\": %.3f}\n",
                n, dims, nq, build_ms, med_ms / nq, pct / nq, batches / nq, agree / nq, ex_ms / nq);
    return 0;
}

int main(int argc, char **argv) {
    if (argc > 1 && std::string(argv[1]) == "--cfg3-bench") {
        try {
            return cfg3_bench(argc > 2 ? std::atoi(argv[2]) : 1000000);
        } catch (const std::exception &e) {
            std::printf("{\"error\": \"%s\"}\n", e.what());
            return 2;
        }
    }
    if (argc > 1 && std::string(argv[1]) == "--cfg3") {
        // BASELINE.json configs[2] at its stated size (default 1 M x 384 float64, cosine): LSH-tree candidate walk on the
        // host, candidates re-scored on the GPU, radius 0.46 with the `bucket < 3` filter, then top-k -- both against the
        // oracle's replay of `consider` over the same visit sequence and against its exact scan (lsh_case above)
        const int n = argc > 2 ? std::atoi(argv[2]) : 1000000;
        try {
            lsh_cases("cfg3", n, 384, 64, Cosine, {LshVariant{0, 0.46, true}, LshVariant{10, 0, false}});
        } catch (const std::exception &e) {
            std::printf("ERROR %s\n", e.what());
            return 2;
        }
        return g_fail ? 1 : 0;
    }
    if (argc > 4 && std::string(argv[1]) == "--open") {
        try {
            return open_mode(argc, argv);
        } catch (const std::exception &e) {
            std::printf("ERROR %s\n", e.what());
            return 2;
        }
    }
    const bool cpu_only = argc > 1 && std::string(argv[1]) == "--cpu";
    struct T { const char *name; void (*fn)(); bool gpu; };
    const T tests[] = {
        {"TestCodec", TestCodec, false},
        {"TestEuclideanDistance", TestEuclideanDistance, true},
        {"TestCollectionSearch", TestCollectionSearch, true},
        {"TestExhaustiveSearch", TestExhaustiveSearch, true},
        {"TestVectorSearchWith4BitQuantization", TestVectorSearchWith4BitQuantization, true},
        {"TestDocumentRoundTripUpdateRemove", TestDocumentRoundTripUpdateRemove, true},
        {"TestListModePagination", TestListModePagination, true},
        {"TestSearchExactVsLSH", TestSearchExactVsLSH, true},
        {"TestLSHQuantized", TestLSHQuantized, true},
        {"TestLSHRadiusWithFilter", TestLSHRadiusWithFilter, true},
    };
    int ran = 0;
    for (const T &t : tests) {
        if (cpu_only && t.gpu) continue;
        const int before = g_fail;
        std::printf("RUN  %s\n", t.name);
        try {
            t.fn();
        } catch (const std::exception &e) {
            std::printf("  FAIL exception: %s\n", e.what());
            ++g_fail;
        }
        std::printf("%s %s\n", g_fail == before ? "PASS" : "FAILED", t.name);
        ++ran;
    }
    if (cpu_only) { // without a GPU the mirror must fail loudly, never fall back
        bool threw = false;
        try {
            CollectionOptions o; o.DimensionCount = 2;
            Collection c(o);
        } catch (const std::runtime_error &e) { threw = std::string(e.what()).find("no CPU fallback") != std::string::npos; }
        std::printf("%s NoGpuFailsLoudly\n", threw ? "PASS" : "SKIP(gpu present)");
    }
    std::printf("%d tests, %d failures\n", ran, g_fail);
    return g_fail ? 1 : 0;
}
