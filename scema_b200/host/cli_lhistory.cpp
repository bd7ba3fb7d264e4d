// lhistory_to_strain OUT_DIRECTORY [--tensor strain|updstrain|stress] pr_0.lhistory.csv [pr_1.lhistory.csv ...]
//
// Converts the FE solver's per-rank history logs (FEProblem::output_lhistory, reference
// headers/FE_problem.h:1985-2045) into the strain_<qpid> text files that the clustering command
// lines read (Strain6D::from_file, headers/strain2spline.h:112-134): one file per quadrature
// point, one line per logged timestep, components reordered from the log's 00,01,02,11,12,22 to
// xx yy zz xy xz yz. Gives real dealammps runs as input to mpi_comparison_test / compare_all_histories.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "scema_ingest.h"

int main(int argc, char **argv)
{
    std::string tensor = "strain";
    std::vector<const char *> files;
    const char *out_dir = NULL;
    for (int i = 1; i < argc; i++) {
        if (strcmp(argv[i], "--tensor") == 0 && i + 1 < argc) tensor = argv[++i];
        else if (!out_dir) out_dir = argv[i];
        else files.push_back(argv[i]);
    }
    if (!out_dir || files.empty()) {
        fprintf(stderr, "Usage: ./lhistory_to_strain OUT_DIRECTORY [--tensor strain|updstrain|stress] LHISTORY_CSV...\n");
        return 1;
    }
    scema_batch *b = NULL;
    if (scema_batch_from_lhistory(files.data(), files.size(), tensor.c_str(), &b) != SCEMA_OK ||
        scema_batch_write_strain_files(b, out_dir) != SCEMA_OK) {
        fprintf(stderr, "%s\n", scema_ingest_last_error());
        return 1;
    }
    printf("%llu histories, %llu steps\n", (unsigned long long)scema_batch_count(b), (unsigned long long)scema_batch_total_steps(b));
    scema_batch_free(b);
    return 0;
}
