// compare_all_histories STRAIN_DIRECTORY NUM_SPLINE_POINTS — GPU drop-in for the reference command
// line clustering/compare_all_histories.cc:31-87 (same argv and stdout format).
//
// The reference source is stale against its own header (it does not compile: SURVEY.md §8c); this
// follows its evident intent (the test-side build of the reference patches it the same way): print
// "Reading: '<f>'" / "Ignoring: '<f>'" per directory entry, then "<A> vs <B>:<L2>" for every pair
// i <= j in batch order (self pairs included), then the three wall-clock lines in whole seconds.
// <A>/<B> are the bare directory entry names. Every distance is computed on the GPU by the exact
// direct-difference path (threshold = +inf makes every finite distance an "edge").
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <iostream>
#include <limits>
#include <string>
#include <vector>
#include "strain2spline_b200.h"
#include "cli_common.h"

int main(int argc, char **argv)
{
    if (argc != 3) {
        fprintf(stderr, "Usage: ./compare_all_histories STRAIN_DIRECTORY NUM_SPLINE_POINTS\n");
        return 1;
    }
    const std::string dir = argv[1];
    const uint32_t spline_points = (uint32_t)atoi(argv[2]);

    std::vector<std::string> entries;
    cli::list_directory(dir, entries);

    const time_t t_read0 = time(NULL);
    std::vector<MatHistPredict::Strain6D *> hist;
    std::vector<std::string> names;
    for (size_t e = 0; e < entries.size(); e++) {
        if (!cli::is_strain_file(entries[e])) {
            std::cout << "Ignoring: '" << entries[e] << "'\n";
            continue;
        }
        std::cout << "Reading: '" << entries[e] << "'\n";
        MatHistPredict::Strain6D *h = new MatHistPredict::Strain6D();
        h->from_file((dir + entries[e]).c_str());
        hist.push_back(h);
        names.push_back(entries[e]);
    }
    const time_t t_read1 = time(NULL);

    const time_t t_spl0 = time(NULL);
    MatHistPredict::splinify_batch(hist, spline_points);
    const time_t t_spl1 = time(NULL);

    const time_t t_cmp0 = time(NULL);
    const size_t n = hist.size();
    if (n) {
        // all finite distances in one dense GPU pass, sorted by (i,j)
        const uint32_t K = (uint32_t)hist[0]->get_spline()->size();
        std::vector<double> rows(n * (size_t)K);
        for (size_t i = 0; i < n; i++) {
            std::vector<double> *sp = hist[i]->get_spline();
            for (uint32_t k = 0; k < K; k++) rows[i * (size_t)K + k] = (*sp)[k];
        }
        scema_ctx *ctx = MatHistPredict::b200::context();
        uint64_t m = 0;
        MatHistPredict::b200::check(scema_set_spline(ctx, rows.data(), 0, n, K, NULL), "compare");
        MatHistPredict::b200::check(scema_compare(ctx, std::numeric_limits<double>::infinity(), SCEMA_PAIRS_EXACT, 0, 1, &m), "compare");
        std::vector<uint32_t> a(m), b(m);
        std::vector<double> d(m);
        MatHistPredict::b200::check(scema_get_edges(ctx, a.data(), b.data(), d.data(), m), "compare");
        uint64_t e = 0;
        for (size_t i = 0; i < n; i++) {
            for (size_t j = i; j < n; j++) {
                double L2;
                if (e < m && a[e] == i && b[e] == j) {
                    L2 = d[e++];
                } else {
                    // self pairs, and pairs whose distance is inf/NaN (never below +inf): host loop
                    L2 = MatHistPredict::compare_L2_norm(hist[i], hist[j]);
                }
                std::cout << names[i] << " vs " << names[j] << ":" << L2 << "\n";
            }
        }
    }
    const time_t t_cmp1 = time(NULL);

    std::cout << "Read time: " << double(t_read1 - t_read0) << "\n";
    std::cout << "Spline time: " << double(t_spl1 - t_spl0) << "\n";
    std::cout << "Compare time: " << double(t_cmp1 - t_cmp0) << "\n";
    return 0;
}
