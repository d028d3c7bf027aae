// Concurrent single-query szg_search_topk calls from T native threads (what the cgo shim's goroutines do):
// aggregate QPS with and without SZG_OPT_COMBINE.  Build: see tools/Makefile-less line in profiles/README.md:
//   g++ -O2 -std=c++17 -o tools/concurrent_bench tools/concurrent_bench.cpp -Iinclude -Lsyzgydb_b200 -lsyzgy_b200 \
//       -Wl,-rpath,'$ORIGIN/../syzgydb_b200' -lpthread
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>

#include "syzgy_b200.h"

int main(int argc, char **argv) {
    const uint64_t rows = argc > 1 ? strtoull(argv[1], nullptr, 10) : 10000000ull;
    const int dims = argc > 2 ? atoi(argv[2]) : 768;
    szg_index *h = nullptr;
    if (szg_create(dims, 8, SZG_COSINE, 0, &h)) { fprintf(stderr, "create: %s\n", szg_last_error()); return 1; }
    if (szg_fill_synthetic(h, 7, 0, rows)) { fprintf(stderr, "fill: %s\n", szg_last_error()); return 1; }
    std::mt19937_64 rng(1);
    std::uniform_real_distribution<double> u(-1.0, 1.0);
    const int NQ = 256;
    std::vector<double> qs((size_t)NQ * dims);
    for (auto &x : qs) x = u(rng);
    {
        uint64_t ids[10]; double dist[10]; uint32_t n;
        szg_search_topk(h, qs.data(), 1, 10, -1, 0, ids, dist, &n, nullptr);
    }
    for (int combine = 0; combine <= 1; ++combine) {
        szg_set_option(h, SZG_OPT_COMBINE, combine);
        for (int T : {1, 4, 16, 64, 128}) {
            const int per = combine ? (T >= 16 ? 64 : 128) : (T >= 16 ? 8 : 64);
            std::atomic<int> errors{0};
            std::vector<std::thread> th;
            const auto t0 = std::chrono::steady_clock::now();
            for (int t = 0; t < T; ++t)
                th.emplace_back([&, t] {
                    uint64_t ids[10]; double dist[10]; uint32_t n;
                    for (int r = 0; r < per; ++r)
                        if (szg_search_topk(h, qs.data() + (size_t)((t * per + r) % NQ) * dims, 1, 10, -1, 0, ids, dist, &n, nullptr) || n != 10)
                            errors++;
                });
            for (auto &x : th) x.join();
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            szg_stats st;
            szg_get_stats(h, &st);
            printf("{\"rows\": %llu, \"dims\": %d, \"combine\": %d, \"threads\": %d, \"calls\": %d, \"qps\": %.1f, \"ms_per_call\": %.3f, "
                   "\"combined_queries_total\": %llu, \"errors\": %d}\n",
                   (unsigned long long)rows, dims, combine, T, T * per, T * per / dt, dt / per * 1e3,
                   (unsigned long long)st.combined_queries, errors.load());
            fflush(stdout);
        }
    }
    szg_destroy(h);
    return 0;
}
