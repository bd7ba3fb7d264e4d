// mpi_comparison_test STRAIN_DIRECTORY NUM_SPLINE_POINTS THRESH — GPU drop-in for the reference
// command line clustering/mpi_comparison_test.cc:46-108 (same argv, same stdout, same result files).
//
// Reads every `strain_<ID>` file of STRAIN_DIRECTORY (which must end in '/': the reference
// concatenates directory and name, :81), resamples each history onto NUM_SPLINE_POINTS points,
// compares all pairs and writes, for every history, `__results/ID_<ID>.txt` with one line
// "<ID> <otherID> <diff>" per history closer than THRESH. The `__results/` directory must exist.
// With MPI (build with -DSCEMA_B200_WITH_MPI and an MPI compiler) files are dealt round-robin to
// the ranks exactly like the reference (:79) and rank 0 drives the GPU.
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>
#ifdef SCEMA_B200_WITH_MPI
#include <mpi.h>
#endif
#include "strain2spline_b200.h"
#include "cli_common.h"

int main(int argc, char **argv)
{
    if (argc != 4) {
        fprintf(stderr, "Usage: ./mpi_comparison_test STRAIN_DIRECTORY NUM_SPLINE_POINTS THRESH\n");
        return 1;
    }
    const std::string dir = argv[1];
    const uint32_t spline_points = (uint32_t)atoi(argv[2]);
    const double threshold = atof(argv[3]);

    int rank = 0, n_ranks = 1;
    MPI_Comm comm = MPI_COMM_WORLD;
#ifdef SCEMA_B200_WITH_MPI
    MPI_Init(NULL, NULL);
    MPI_Comm_rank(comm, &rank);
    MPI_Comm_size(comm, &n_ranks);
#endif

    std::vector<std::string> entries;
    cli::list_directory(dir, entries);

    std::vector<MatHistPredict::Strain6D *> mine;
    int seen = 0;
    for (size_t e = 0; e < entries.size(); e++) {
        if (!cli::is_strain_file(entries[e])) {
            std::cout << "Ignoring: '" << entries[e] << "'\n";
            continue;
        }
        if (seen++ % n_ranks != rank) continue;
        MatHistPredict::Strain6D *h = new MatHistPredict::Strain6D();
        h->from_file((dir + entries[e]).c_str());
        h->splinify(spline_points);  // deferred: one batched K1 launch inside the compare below
        h->set_ID(cli::id_from_name(entries[e]));
        mine.push_back(h);
    }

    MatHistPredict::compare_histories_with_all_ranks(mine, threshold, comm);

    for (size_t i = 0; i < mine.size(); i++)
        mine[i]->most_similar_histories_to_file(("__results/ID_" + std::to_string(mine[i]->get_ID()) + ".txt").c_str());

#ifdef SCEMA_B200_WITH_MPI
    MPI_Finalize();
#endif
    return 0;
}
