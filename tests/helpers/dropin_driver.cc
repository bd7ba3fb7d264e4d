// A caller written ONLY against the MatHistPredict API of SCEMa's headers/strain2spline.h, shaped like
// the in-process caller (reference headers/FE_problem.h:1091-1103 update_strain_quadrature_point_history,
// :1167-1191 spline_building, :1197-1270 spline_comparison) and like clustering/mpi_comparison_test.cc.
// It is compiled twice, unchanged:
//   * against the reference header (oracle/Makefile -> oracle/_ref/dropin_driver_ref), and
//   * against the drop-in header scema_b200/host/strain2spline_b200.h + libscema_hist.so (the test),
// and both binaries must print the same bytes and write the same files.
//
//   dropin_driver OUT_DIR N_QP N_STEPS SPLINE_POINTS THRESHOLD SEED
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <string>
#include <vector>
#include <stdint.h>
#ifdef DROPIN_REFERENCE
#include <mpi.h>  // oracle/mpi_shim/mpi.h (single rank)
#include "strain2spline.h"
#else
#include "strain2spline_b200.h"
#endif

static uint64_t mix(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static double unit(uint64_t seed, uint64_t a, uint64_t b, uint64_t c)  // [0,1), exact in double
{
    return (double)(mix(mix(mix(seed ^ a) ^ b) ^ c) >> 11) * (1.0 / 9007199254740992.0);
}

int main(int argc, char **argv)
{
    if (argc != 7) {
        fprintf(stderr, "Usage: dropin_driver OUT_DIR N_QP N_STEPS SPLINE_POINTS THRESHOLD SEED\n");
        return 1;
    }
    const std::string out = argv[1];
    const uint32_t n_qp = (uint32_t)atoi(argv[2]), n_steps = (uint32_t)atoi(argv[3]), P = (uint32_t)atoi(argv[4]);
    const double thr = atof(argv[5]);
    const uint64_t seed = (uint64_t)atoll(argv[6]);

    std::vector<MatHistPredict::Strain6D> qp(n_qp);  // embedded by value, like PointHistory::hist_strain (FE.h:98)
    for (uint32_t q = 0; q < n_qp; q++) qp[q].set_ID(3 * q + 1);

    for (uint32_t t = 1; t <= n_steps; t++) {
        // one strain sample per quadrature point per timestep; groups of 8 points share a loading path
        for (uint32_t q = 0; q < n_qp; q++) {
            const uint32_t grp = q / 8;
            double s[6];
            for (int c = 0; c < 6; c++) {
                const double amp = (unit(seed, grp, c, 1) - 0.5) * 1e-3;
                const double jit = (unit(seed, q, c, 2) - 0.5) * 4e-8;
                const double noise = (unit(seed, q, c, 1000 + t) - 0.5) * 1e-10;
                s[c] = (amp + jit) * t + noise;
            }
            if (t % 2) qp[q].add_current_strain(s[0], s[1], s[2], s[3], s[4], s[5]);
            else qp[q].add_current_strain(s[0], s[1], s[2], s[3], s[4], s[5], 1.0, 2.0, 3.0, 4.0, 5.0, 6.0);
            qp[q].set_most_recent_ID_to_get_results_from(qp[q].get_ID_to_get_results_from());
            qp[q].set_ID_to_get_results_from(qp[q].get_ID());
        }
        if (t < 3 || (t != 3 && t != n_steps && t % 5)) continue;

        // spline_building: every point
        for (uint32_t q = 0; q < n_qp; q++) qp[q].splinify(P);
        // spline_comparison: only the flagged points
        std::vector<MatHistPredict::Strain6D *> flagged;
        for (uint32_t q = 0; q < n_qp; q++)
            if (unit(seed, q, t, 7) < 0.7) flagged.push_back(&qp[q]);
        MatHistPredict::compare_histories_with_all_ranks(flagged, thr, MPI_COMM_WORLD);
        std::cout << "timestep " << t << ": " << flagged.size() << " flagged\n";
        for (size_t i = 0; i < flagged.size(); i++) {
            char name[1024];
            snprintf(name, sizeof name, "%s/t%u.last.%u.similar_hist", out.c_str(), t, flagged[i]->get_ID());
            flagged[i]->most_similar_histories_to_file(name);
            if (i < 5) flagged[i]->print_most_similar_histories();
            // DROPIN_LEGACY=1: the "theory-checking" outputs too — the full comparison lists (FE_problem.h:1237-1238)
            // and the legacy nearest neighbour; =2: the nearest neighbour only
            const char *legacy = getenv("DROPIN_LEGACY");
            if (legacy && atoi(legacy) == 1) {
                snprintf(name, sizeof name, "%s/t%u.last.%u.all_similar_hist", out.c_str(), t, flagged[i]->get_ID());
                flagged[i]->all_similar_histories_to_file(name);
            }
            if (legacy && atoi(legacy) >= 1 && i < 40) {
                std::cout.precision(17);
                std::cout << "nearest of " << flagged[i]->get_ID() << ": " << flagged[i]->get_most_similar_history_ID() << " at "
                          << flagged[i]->get_most_similar_history_diff() << "\n";
                std::cout.precision(6);
            }
        }
        // a few splines and one direct distance, in full precision
        std::cout.precision(17);
        std::vector<double> *sp = qp[t % n_qp].get_spline();
        std::cout << "spline of point " << t % n_qp << ":";
        for (size_t k = 0; k < sp->size(); k++) std::cout << ' ' << (*sp)[k];
        std::cout << "\nL2(0,1) = " << MatHistPredict::compare_L2_norm(&qp[0], &qp[1]) << "\n";
        std::cout.precision(6);
        qp[2].print();
    }
    if (n_steps < 3) qp[0].splinify(P);  // too few samples: message + exit(1) (strain2spline.h:142-148)
    // mapping round trip (read_coarsegrain_dependency_mapping + run_new_md)
    const std::string map_file = out + "/mapping.csv";
    {
        FILE *f = fopen(map_file.c_str(), "w");
        for (uint32_t i = 0; i < 3 * n_qp + 1; i++) fprintf(f, "%u %u\n", i, i % 4 ? i : (i > 8 ? i - 3 : i));
        fclose(f);
    }
    uint32_t new_md = 0;
    for (uint32_t q = 0; q < n_qp; q++) {
        qp[q].read_coarsegrain_dependency_mapping(map_file.c_str());
        new_md += qp[q].run_new_md();
    }
    std::cout << "run_new_md: " << new_md << " of " << n_qp << ", point 1 takes results from "
              << qp[1].get_ID_to_get_results_from() << " (was " << qp[1].get_most_recent_ID_to_get_results_from() << ")\n";
#ifndef DROPIN_REFERENCE
    if (getenv("DROPIN_STATS")) {  // how the histories reached the GPU (stderr: not part of the compared output)
        const MatHistPredict::b200::StoreState &st = MatHistPredict::b200::store_state();
        fprintf(stderr, "store: rebuilds=%llu appended_steps=%llu h2d_bytes=%llu flat_uploads=%llu\n", (unsigned long long)st.rebuilds,
                (unsigned long long)st.appended_steps, (unsigned long long)st.h2d_bytes, (unsigned long long)st.flat_uploads);
    }
#endif
    return 0;
}
