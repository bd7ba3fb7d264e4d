/* TEST INFRASTRUCTURE ONLY — single-rank stand-in for <mpi.h>.
 *
 * The reference header (headers/strain2spline.h:513-614) uses MPI_Comm, MPI_Send, MPI_Recv,
 * MPI_Comm_rank and MPI_Comm_size without including <mpi.h>, and this container has no MPI.
 * With one rank only the `target_rank == this_rank` branch of compare_histories_with_all_ranks
 * runs (strain2spline.h:576,601), so Send/Recv are never reached; they trap if they are.
 * This file is ours (not reference code); it exists only so oracle/_ref can be compiled from
 * the unmodified reference sources.
 */
#ifndef SCEMA_ORACLE_MPI_SHIM_H
#define SCEMA_ORACLE_MPI_SHIM_H

#include <cstdio>
#include <cstdlib>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef struct { int unused; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_UNSIGNED 1
#define MPI_DOUBLE 2

static inline int MPI_Init(int *, char ***) { return 0; }
static inline int MPI_Finalize(void) { return 0; }
static inline int MPI_Comm_rank(MPI_Comm, int *rank) { *rank = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm, int *size) { *size = 1; return 0; }
static inline int MPI_Barrier(MPI_Comm) { return 0; }

static inline int MPI_Send(const void *, int, MPI_Datatype, int, int, MPI_Comm)
{
    fprintf(stderr, "mpi_shim: MPI_Send reached in a single-rank build\n");
    abort();
}
static inline int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *)
{
    fprintf(stderr, "mpi_shim: MPI_Recv reached in a single-rank build\n");
    abort();
}

#endif
