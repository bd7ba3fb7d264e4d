// mpi_comparison_test STRAIN_DIRECTORY NUM_SPLINE_POINTS THRESH — GPU drop-in for the reference
// command line clustering/mpi_comparison_test.cc:46-108 (same argv, same stdout, same result files).
//
// Reads every `strain_<ID>` file of STRAIN_DIRECTORY (which must end in '/': the reference
// concatenates directory and name, :81), resamples each history onto NUM_SPLINE_POINTS points,
// compares all pairs and writes, for every history, `__results/ID_<ID>.txt` with one line
// "<ID> <otherID> <diff>" per history closer than THRESH. The `__results/` directory must exist.
// Single process (the default build): the strain files are parsed concurrently by the batch
// reader (include/scema_ingest.h) into one ragged batch that goes straight through the C ABI —
// K1 resample, K2 all-pairs, K3 compaction, then one result file per history.
// With MPI (build with -DSCEMA_B200_WITH_MPI and an MPI compiler) files are dealt round-robin to
// the ranks exactly like the reference (:79), each rank reads its share through the drop-in
// Strain6D class and rank 0 drives the GPU.
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>
#ifdef SCEMA_B200_WITH_MPI
#include <mpi.h>
#endif
#include "strain2spline_b200.h"
#include "scema_ingest.h"
#include "cli_common.h"

#ifndef SCEMA_B200_WITH_MPI
// reference error behaviour: message on stderr, exit(1)
static void die(const std::string &msg)
{
    fprintf(stderr, "%s\n", msg.c_str());
    exit(1);
}
#endif

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

#ifndef SCEMA_B200_WITH_MPI
    (void)rank; (void)n_ranks; (void)comm;
    std::vector<std::string> paths;
    std::vector<uint32_t> ids;
    for (size_t e = 0; e < entries.size(); e++) {
        if (!cli::is_strain_file(entries[e])) {
            std::cout << "Ignoring: '" << entries[e] << "'\n";
            continue;
        }
        paths.push_back(dir + entries[e]);
        ids.push_back(cli::id_from_name(entries[e]));
    }
    std::vector<const char *> cpaths(paths.size());
    for (size_t i = 0; i < paths.size(); i++) cpaths[i] = paths[i].c_str();
    scema_batch *batch = NULL;
    if (scema_batch_read_files(cpaths.data(), ids.data(), paths.size(), 0, &batch) != SCEMA_OK) die(scema_ingest_last_error());
    const uint64_t n = scema_batch_count(batch);
    const uint64_t *off = scema_batch_offsets(batch);
    for (uint64_t i = 0; i < n; i++) {  // Strain6D::splinify's checks (strain2spline.h:142-148), in batch order
        if (off[i + 1] == off[i])
            die("Error: Nothing to splinify! No strain data has been read in yet. Please use .from_file() or .add_current_strain() first.");
        if (off[i + 1] - off[i] < 3) die("Error: Not enough strain steps added. Need at least 3 points for splinify().");
    }
    if (n) {
        scema_ctx *ctx = MatHistPredict::b200::context();
        uint64_t n_edges = 0;
        if (MatHistPredict::b200::multi() && spline_points > 0) {
            // SCEMA_B200_DEVICES lists several GPUs: the batch is sharded over them inside the library (scema_multi_cluster);
            // the result files below come from the first context, which holds the merged list
            MatHistPredict::b200::check_multi(scema_multi_cluster(MatHistPredict::b200::multi(), scema_batch_steps(batch), off,
                                                                  scema_batch_ids(batch), n, spline_points, threshold, SCEMA_PAIRS_TC,
                                                                  &n_edges),
                                              "compare_histories_with_all_ranks");
            if (scema_write_similar_hist(ctx, "__results/ID_%u.txt") != SCEMA_OK) die(scema_last_error(ctx));
            scema_batch_free(batch);
            return 0;
        }
        MatHistPredict::b200::check(scema_set_histories_from_batch(ctx, batch), "from_file");
        if (spline_points == 0) {
            // the reference builds empty spline vectors, every distance is sqrt(0) = 0
            std::vector<double> none(1, 0.0);
            MatHistPredict::b200::check(scema_set_spline(ctx, none.data(), 0, n, 0, scema_batch_ids(batch)), "splinify");
        } else {
            MatHistPredict::b200::check(scema_resample(ctx, spline_points), "splinify");
        }
        MatHistPredict::b200::check(scema_compare(ctx, threshold, SCEMA_PAIRS_TC, 0, 1, &n_edges), "compare_histories_with_all_ranks");
        // most_similar_histories_to_file per history (:99-103); fails like the reference's ofstream
        // when __results/ does not exist
        if (scema_write_similar_hist(ctx, "__results/ID_%u.txt") != SCEMA_OK) die(scema_last_error(ctx));
    }
    scema_batch_free(batch);
    return 0;
#else
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

    MPI_Finalize();
    return 0;
#endif
}
