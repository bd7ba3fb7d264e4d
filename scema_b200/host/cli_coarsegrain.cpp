// coarsegrain_dependency_network INPUT_FOLDER OUT_MAPPING_CSV NUMBER_OF_GPS — native equivalent of
// clustering/coarsegrain_dependency_network.py (same argv, same mapping.csv, same three summary
// lines, :92-94). SCEMa keeps calling the Python script unchanged; this exists because the script is
// O(V^2 log V) and becomes the wall beyond ~1e5 histories.
#include <cstdio>
#include <cstdlib>
#include "scema_hist.h"

int main(int argc, char **argv)
{
    if (argc != 4) {
        fprintf(stderr, "Usage: coarsegrain_dependency_network.py [input_folder] [out_mapping.csv] [number_of_gps]\n");
        return 1;
    }
    uint64_t iterations = 0, files = 0, removed = 0;
    int rc = scema_reduce_dir(argv[1], argv[2], (uint32_t)atoi(argv[3]), &iterations, &files, &removed);
    if (rc != SCEMA_OK) {
        fprintf(stderr, "coarsegrain_dependency_network: failed (code %d)\n", rc);
        return 1;
    }
    printf("              Converged in %llu iterations\n", (unsigned long long)iterations);
    printf("              Number of gauss points to be udpated:  %llu\n", (unsigned long long)files);
    printf("              Number of simulations required:  %lld\n", (long long)files - (long long)removed);
    return 0;
}
