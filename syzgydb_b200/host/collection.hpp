// collection.hpp -- C++ host mirror of the reference's Collection surface for the search hot path.
//
// The reference is Go and this image has no Go toolchain, so the host side above the C ABI
// (include/syzgy_b200.h) is written in C++ with the reference's names, argument meaning and
// error behaviour (collection.go:31-48 CollectionOptions, 98-158 Document/SearchResult(s)/SearchArgs,
// 184 FilterFn, 186-189 Euclidean/Cosine, 427-521 Add/Get/Update/removeDocument, 569-711 Search).
// What is NOT mirrored: the span file (records live in an in-memory map here; storage is out of
// scope), stats, REST, the filter language (FilterFn is any callable).
//
// Search paths:
//   Precision == "exact"  -> one szg_search_topk / szg_search_radius call (the GPU scan)
//   anything else         -> the LSH tree (lshtree.go restated: 5 random-hyperplane trees, leaf
//                            threshold 100) is traversed on the host; leaf id lists are rescored on
//                            the GPU in speculative batches (szg_rescore) and `consider`
//                            (collection.go:583-629) + k_counter/prune (lshtree.go:299-350) are
//                            replayed over the returned distances, stopping exactly where the
//                            reference would.
//   K == 0 && Radius == 0 -> list mode (collection.go:633-668), host only.
#pragma once
#include <cstdint>
#include <functional>
#include <memory>
#include <string>
#include <vector>

namespace syzgydb {

enum DistanceMethod : int { Euclidean = 0, Cosine = 1 }; // collection.go:186-189

// searchCallback signals, collection.go:19-24
enum Signal : int { StopSearch = 0, PointAccepted = 1, PointChecked = 2, PointIgnored = 3 };

struct CollectionOptions { // collection.go:31-48
    std::string Name;
    int DistanceMethod = Euclidean;
    int DimensionCount = 0;
    int Quantization = 64; // 4, 8, 16, 32, 64; 0 means 64 (collection.go:254-256)
    // additions of this build (no reference counterpart)
    int Device = 0;      // CUDA device ordinal of the mirror
    uint64_t Seed = 1;   // seed of the LSH tree's random source (the reference's tree is racy/non-reproducible)
};

struct Document { // collection.go:101-110
    uint64_t ID = 0;
    std::vector<double> Vector;
    std::string Metadata;
};

struct SearchResult { // collection.go:115-124
    uint64_t ID = 0;
    std::string Metadata;
    double Distance = 0;
};

struct SearchResults { // collection.go:129-135
    std::vector<SearchResult> Results;
    double PercentSearched = 0;
};

using FilterFn = std::function<bool(uint64_t id, const std::string &metadata)>; // collection.go:184

struct SearchArgs { // collection.go:140-158
    std::vector<double> Vector;
    FilterFn Filter;
    int K = 0;
    double Radius = 0;
    int Offset = 0;
    int Limit = 0;
    std::string Precision; // "" -> "medium"; only "exact" scans everything (collection.go:573-575, 672)
};

// codec, quantization.go:5-36 and collection.go:713-811
uint64_t quantize(double value, int bits);
double dequantize(uint64_t value, int bits);
int getVectorSize(int quantization, int dimensions); // throws on an unsupported level (the reference panics)
std::vector<uint8_t> encodeDocument(const std::vector<double> &vector, int quantization);
std::vector<double> decodeVector(const uint8_t *data, int dimensions, int quantization);

class Collection {
public:
    explicit Collection(const CollectionOptions &options); // throws std::runtime_error without a B200
    // NewCollection on an existing collection file (collection.go:241-252, 298-311): options from the header record,
    // stream 1 of every live document bulk-loaded into the mirror by the library's span-file reader (szg_spanfile_*,
    // the file is read, never written), metadata copied out of the mapping, the LSH trees rebuilt from the decoded
    // vectors in IterateSortedRecords order like the reference's reload loop.
    static std::unique_ptr<Collection> Open(const std::string &path, int device = 0, uint64_t seed = 1);
    ~Collection();
    Collection(const Collection &) = delete;
    Collection &operator=(const Collection &) = delete;

    const CollectionOptions &Options() const;
    void AddDocument(uint64_t id, const std::vector<double> &vector, const std::string &metadata); // collection.go:427-457
    // bulk form of the reload loop in NewCollection (collection.go:298-311): one GPU upsert for all rows
    void AddDocuments(const std::vector<uint64_t> &ids, const std::vector<std::vector<double>> &vectors,
                      const std::vector<std::string> &metadata);
    bool GetDocument(uint64_t id, Document *out) const;           // collection.go:463-484 (false = not found)
    bool UpdateDocument(uint64_t id, const std::string &metadata); // collection.go:490-509
    bool removeDocument(uint64_t id);                              // collection.go:511-521
    int GetDocumentCount() const;                                  // collection.go:54-61
    SearchResults Search(SearchArgs args);                         // collection.go:569-711
    // The results of Search(args[i]) for every i.  Exact top-k queries without a filter (all with the same K) go to the
    // GPU as ONE batch (szg_search_batch: the tensor-core contraction for 8/16-bit collections); everything else is
    // answered one by one.  No reference counterpart: there the calls would run concurrently under the RLock (570).
    std::vector<SearchResults> SearchBatch(const std::vector<SearchArgs> &args);
    void Close();                                                  // collection.go:408-421

    // test hooks: the ids the calling thread's last index-driven Search fed to `consider`, in order, and how many GPU
    // rescoring batches it took
    const std::vector<uint64_t> &LastVisitSequence() const;
    int LastRescoreBatches() const;

private:
    struct Impl;
    std::unique_ptr<Impl> p_;
};

} // namespace syzgydb
