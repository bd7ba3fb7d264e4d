// Multi-rank caller of MatHistPredict::compare_histories_with_all_ranks, ranks = threads (tests/helpers/mpi_threads).
// Compiled twice: -DDROPIN_REFERENCE against the reference's own header (the real ring, strain2spline.h:546-614) and
// against scema_b200/host/strain2spline_b200.h (gather to rank 0, GPU, scatter back in ring order). Prints, rank by
// rank and history by history, the result lists exactly as most_similar_histories_to_file would write them.
//   multirank_driver N_RANKS N_HISTORIES SPLINE_POINTS THRESHOLD SEED
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <unistd.h>
#include <sstream>
#include <string>
#include <vector>
#include <stdint.h>
#include <mpi.h>
#ifdef DROPIN_REFERENCE
#include "strain2spline.h"
#else
#include "strain2spline_b200.h"
#endif

static uint64_t mix(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static double u01(uint64_t a, uint64_t b, uint64_t s) { return (double)(mix(mix(s + 0x9e3779b97f4a7c15ull * (a + 1)) + b) >> 11) * 1.1102230246251565e-16; }

int main(int argc, char **argv)
{
    if (argc != 6) { fprintf(stderr, "usage: %s N_RANKS N_HISTORIES SPLINE_POINTS THRESHOLD SEED\n", argv[0]); return 2; }
    const int n_ranks = atoi(argv[1]);
    const uint32_t n = (uint32_t)atoi(argv[2]), P = (uint32_t)atoi(argv[3]);
    const double thr = atof(argv[4]);
    const uint64_t seed = (uint64_t)atoll(argv[5]);
    std::vector<std::string> out(n_ranks);
    mpi_threads::run(n_ranks, [&](int rank) {
        // history q lives on rank q % n_ranks (uneven when n is not a multiple); groups of 4 consecutive histories are
        // near copies of each other, so partners sit on the same rank and on other ranks
        std::vector<MatHistPredict::Strain6D *> mine;
        for (uint32_t q = 0; q < n; q++) {
            if ((int)(q % (uint32_t)n_ranks) != rank) continue;
            MatHistPredict::Strain6D *h = new MatHistPredict::Strain6D();
            const uint32_t g = q / 4, L = 5 + (uint32_t)(mix(seed + g) % 9);
            for (uint32_t s = 0; s < L; s++) {
                double v[6];
                const double t = (double)s / (double)(L - 1);
                for (int c = 0; c < 6; c++)
                    v[c] = 5e-3 * (2.0 * u01(g, c, seed) - 1.0) * t + 4e-7 * (2.0 * u01(q, 10 + c, seed) - 1.0) * t * t;
                h->add_current_strain(v[0], v[1], v[2], v[3], v[4], v[5]);
            }
            h->set_ID(1000 + 7 * q);
            h->splinify(P);
            mine.push_back(h);
        }
        MatHistPredict::compare_histories_with_all_ranks(mine, thr, MPI_COMM_WORLD);
        std::ostringstream os;
        for (size_t i = 0; i < mine.size(); i++) {
            // capture print_most_similar_histories' format through the file writer's twin: "<ID> <other> <diff>\n"
            char name[64];
            snprintf(name, sizeof name, "/tmp/multirank_%d_%d_%zu.txt", (int)getpid(), rank, i);
            mine[i]->most_similar_histories_to_file(name);
            FILE *f = fopen(name, "r");
            char line[256];
            os << "history " << mine[i]->get_ID() << " on rank " << rank << "\n";
            while (f && fgets(line, sizeof line, f)) os << line;
            if (f) fclose(f);
            remove(name);
        }
        out[rank] = os.str();
    });
    for (int r = 0; r < n_ranks; r++) fputs(out[r].c_str(), stdout);
    return 0;
}
