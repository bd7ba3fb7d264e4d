/* TEST INFRASTRUCTURE ONLY — thread-backed stand-in for <mpi.h>: the "ranks" of MPI_COMM_WORLD are threads of one
 * process (mpi_threads::run(n_ranks, fn)). Enough of MPI to run BOTH multi-rank code paths unchanged in a container
 * without MPI:
 *   - the reference's ring (headers/strain2spline.h:513-614): MPI_Send / MPI_Recv / MPI_Comm_rank / MPI_Comm_size;
 *   - the drop-in header's gather-to-rank-0 branch (scema_b200/host/strain2spline_b200.h): MPI_Gather(v), MPI_Scatter(v).
 * Sends are buffered (the reference issues all its blocking sends before its receives, which only works with eager
 * delivery); messages match on (source, destination, tag) in FIFO order; a receive takes at most `count` elements.
 * This file is ours; no reference code. */
#ifndef SCEMA_TEST_MPI_THREADS_H
#define SCEMA_TEST_MPI_THREADS_H

#include <condition_variable>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <mutex>
#include <thread>
#include <tuple>
#include <vector>

#define MPI_VERSION 3
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef struct { int MPI_SOURCE, MPI_TAG, count_bytes; } MPI_Status;
#define MPI_COMM_WORLD 0
#define MPI_UNSIGNED 4
#define MPI_INT 5
#define MPI_DOUBLE 8
#define MPI_SUCCESS 0

namespace mpi_threads {

struct World {
    int size = 1;
    std::mutex m;
    std::condition_variable cv;
    std::map<std::tuple<int, int, int>, std::deque<std::vector<char> > > box;  // (src, dst, tag) -> messages
};
inline World &world() { static World w; return w; }
inline int &my_rank() { static thread_local int r = 0; return r; }
inline size_t type_size(MPI_Datatype t) { return t == MPI_DOUBLE ? 8 : 4; }

inline void run(int n_ranks, const std::function<void(int)> &fn)
{
    World &w = world();
    w.size = n_ranks;
    w.box.clear();
    std::vector<std::thread> th;
    for (int r = 0; r < n_ranks; r++)
        th.emplace_back([r, &fn] { my_rank() = r; fn(r); });
    for (size_t i = 0; i < th.size(); i++) th[i].join();
}

}  // namespace mpi_threads

static inline int MPI_Init(int *, char ***) { return 0; }
static inline int MPI_Finalize(void) { return 0; }
static inline int MPI_Comm_rank(MPI_Comm, int *rank) { *rank = mpi_threads::my_rank(); return 0; }
static inline int MPI_Comm_size(MPI_Comm, int *size) { *size = mpi_threads::world().size; return 0; }

static inline int MPI_Send(const void *buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm)
{
    mpi_threads::World &w = mpi_threads::world();
    std::vector<char> msg((size_t)count * mpi_threads::type_size(t));
    if (!msg.empty()) memcpy(msg.data(), buf, msg.size());
    {
        std::lock_guard<std::mutex> lk(w.m);
        w.box[std::make_tuple(mpi_threads::my_rank(), dest, tag)].push_back(msg);
    }
    w.cv.notify_all();
    return 0;
}

static inline int MPI_Recv(void *buf, int count, MPI_Datatype t, int source, int tag, MPI_Comm, MPI_Status *st)
{
    mpi_threads::World &w = mpi_threads::world();
    std::unique_lock<std::mutex> lk(w.m);
    std::deque<std::vector<char> > &q = w.box[std::make_tuple(source, mpi_threads::my_rank(), tag)];
    w.cv.wait(lk, [&] { return !q.empty(); });
    std::vector<char> msg = q.front();
    q.pop_front();
    const size_t take = std::min(msg.size(), (size_t)count * mpi_threads::type_size(t));
    if (take) memcpy(buf, msg.data(), take);
    if (st) { st->MPI_SOURCE = source; st->MPI_TAG = tag; st->count_bytes = (int)take; }
    return 0;
}

static inline int MPI_Barrier(MPI_Comm c)
{
    int r, n, token = 0;
    MPI_Comm_rank(c, &r);
    MPI_Comm_size(c, &n);
    if (r == 0) {
        for (int q = 1; q < n; q++) MPI_Recv(&token, 1, MPI_INT, q, -7, c, NULL);
        for (int q = 1; q < n; q++) MPI_Send(&token, 1, MPI_INT, q, -8, c);
    } else {
        MPI_Send(&token, 1, MPI_INT, 0, -7, c);
        MPI_Recv(&token, 1, MPI_INT, 0, -8, c, NULL);
    }
    return 0;
}

static inline int MPI_Gatherv(const void *sbuf, int scount, MPI_Datatype st, void *rbuf, const int *rcounts, const int *displs,
                              MPI_Datatype rt, int root, MPI_Comm c)
{
    int r, n;
    MPI_Comm_rank(c, &r);
    MPI_Comm_size(c, &n);
    if (r != root) return MPI_Send(sbuf, scount, st, root, -3, c);
    for (int q = 0; q < n; q++) {
        char *dst = (char *)rbuf + (size_t)displs[q] * mpi_threads::type_size(rt);
        if (q == root) { if (scount) memcpy(dst, sbuf, (size_t)scount * mpi_threads::type_size(st)); }
        else MPI_Recv(dst, rcounts[q], rt, q, -3, c, NULL);
    }
    return 0;
}

static inline int MPI_Gather(const void *sbuf, int scount, MPI_Datatype st, void *rbuf, int rcount, MPI_Datatype rt, int root, MPI_Comm c)
{
    int n;
    MPI_Comm_size(c, &n);
    std::vector<int> counts(n, rcount), displs(n);
    for (int q = 0; q < n; q++) displs[q] = q * rcount;
    return MPI_Gatherv(sbuf, scount, st, rbuf, counts.data(), displs.data(), rt, root, c);
}

static inline int MPI_Scatterv(const void *sbuf, const int *scounts, const int *displs, MPI_Datatype st, void *rbuf, int rcount,
                               MPI_Datatype rt, int root, MPI_Comm c)
{
    int r, n;
    MPI_Comm_rank(c, &r);
    MPI_Comm_size(c, &n);
    if (r != root) return MPI_Recv(rbuf, rcount, rt, root, -4, c, NULL);
    for (int q = 0; q < n; q++) {
        const char *src = (const char *)sbuf + (size_t)displs[q] * mpi_threads::type_size(st);
        if (q == root) { if (scounts[q]) memcpy(rbuf, src, (size_t)scounts[q] * mpi_threads::type_size(st)); }
        else MPI_Send(src, scounts[q], st, q, -4, c);
    }
    return 0;
}

static inline int MPI_Scatter(const void *sbuf, int scount, MPI_Datatype st, void *rbuf, int rcount, MPI_Datatype rt, int root, MPI_Comm c)
{
    int n;
    MPI_Comm_size(c, &n);
    std::vector<int> counts(n, scount), displs(n);
    for (int q = 0; q < n; q++) displs[q] = q * scount;
    return MPI_Scatterv(sbuf, counts.data(), displs.data(), st, rbuf, rcount, rt, root, c);
}

#endif
