// TEST INFRASTRUCTURE ONLY — C-ABI harness around the UNMODIFIED reference header.
//
// This file is ours; it contains no reference code. It is compiled by oracle/Makefile with
//   -include limits -I oracle/mpi_shim -I /root/reference/headers
// so that `#include "strain2spline.h"` resolves to the reference's own
// headers/strain2spline.h (+ headers/spline.h) where they lie under /root/reference.
// The result (oracle/_ref/libscema_ref.so) is the parity pin for oracle/hist_oracle.c and,
// via bench.py --impl reference / cpu_baseline, the timed CPU baseline ("kind": "reference").
// Nothing on the product path may load it.
//
// Every function below is a thin call into reference symbols:
//   Strain6D::add_current_strain   strain2spline.h:75-86
//   Strain6D::splinify             strain2spline.h:140-180   (-> tk::spline, spline.h:284-396)
//   Strain6D::get_spline           strain2spline.h:214-221
//   compare_L2_norm(double*,...)   strain2spline.h:469-484
//   compare_histories_with_all_ranks  strain2spline.h:546-614 (single rank via mpi_shim)
//   most_similar_histories_to_file strain2spline.h:301-314
//   Strain6D::from_file            strain2spline.h:112-134
#include <limits>
#include <iostream>
#include <sstream>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <stdint.h>
#include <mpi.h>            // oracle/mpi_shim/mpi.h
#include <fstream>
#include <cmath>
#include <cstdio>
#include <cstdlib>
// ref_from_file() below needs to read back what Strain6D::from_file stored; the six input vectors
// are private and have no accessor, so the header is compiled with `private` spelled `public`
// (every standard header it uses is already included above; the reference source is untouched).
#define private public
#include "strain2spline.h"  // the reference header, unmodified
#undef private

#ifdef _OPENMP
#include <omp.h>
#endif

using MatHistPredict::Strain6D;

extern "C" {

int ref_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// steps: [L][6] (xx yy zz xy xz yz per step). out: [6*P], reference interleaved order p*6+c.
void ref_splinify(const double *steps, uint32_t L, uint32_t P, double *out)
{
    Strain6D h;
    for (uint32_t n = 0; n < L; n++) {
        const double *s = steps + 6 * (size_t)n;
        h.add_current_strain(s[0], s[1], s[2], s[3], s[4], s[5]);
    }
    h.splinify(P);
    std::vector<double> *sp = h.get_spline();
    memcpy(out, sp->data(), sizeof(double) * sp->size());
}

// Ragged batch: offsets[N+1] in steps. out: [N][6P]. nthreads<=0 -> all.
void ref_splinify_batch(const double *steps, const uint64_t *offsets, uint64_t N, uint32_t P,
                        double *out, int nthreads)
{
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads)
#endif
    for (int64_t i = 0; i < (int64_t)N; i++) {
        ref_splinify(steps + 6 * offsets[i], (uint32_t)(offsets[i + 1] - offsets[i]), P,
                     out + (size_t)i * 6 * P);
    }
}

double ref_compare_l2(const double *a, const double *b, uint32_t K)
{
    return MatHistPredict::compare_L2_norm(const_cast<double *>(a), const_cast<double *>(b), K, K);
}

// All pairs i<j with i in [row_begin,row_end) through the reference's compare_L2_norm, strict
// `diff < thr` as choose_most_similar_history (strain2spline.h:272). Edges are appended to
// (ei, ej, ed) up to cap and finally sorted by (i,j). Returns the number of edges found (may
// exceed cap; then only the first cap of some order were kept and the caller must retry).
// pairs_out receives the number of pair comparisons executed.
uint64_t ref_all_pairs(const double *rows, uint64_t N, uint32_t K, double thr, uint64_t row_begin,
                       uint64_t row_end, int nthreads, uint32_t *ei, uint32_t *ej, double *ed,
                       uint64_t cap, uint64_t *pairs_out)
{
    if (row_end > N) row_end = N;
    uint64_t total = 0, pairs = 0;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    std::vector<std::vector<uint32_t> > li(nthreads), lj(nthreads);
    std::vector<std::vector<double> > ld(nthreads);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 8) num_threads(nthreads) reduction(+ : pairs)
#endif
    for (int64_t i = (int64_t)row_begin; i < (int64_t)row_end; i++) {
#ifdef _OPENMP
        int t = omp_get_thread_num();
#else
        int t = 0;
#endif
        double *a = const_cast<double *>(rows + (size_t)i * K);
        for (uint64_t j = i + 1; j < N; j++) {
            double *b = const_cast<double *>(rows + (size_t)j * K);
            double d = MatHistPredict::compare_L2_norm(a, b, K, K);
            if (d < thr) {
                li[t].push_back((uint32_t)i);
                lj[t].push_back((uint32_t)j);
                ld[t].push_back(d);
            }
        }
        pairs += N - 1 - i;
    }
    struct E { uint32_t i, j; double d; };
    std::vector<E> all;
    for (int t = 0; t < nthreads; t++)
        for (size_t e = 0; e < li[t].size(); e++) all.push_back(E{li[t][e], lj[t][e], ld[t][e]});
    std::sort(all.begin(), all.end(),
              [](const E &x, const E &y) { return x.i != y.i ? x.i < y.i : x.j < y.j; });
    total = all.size();
    if (ei && ej && ed)
        for (uint64_t e = 0; e < total && e < cap; e++) {
            ei[e] = all[e].i; ej[e] = all[e].j; ed[e] = all[e].d;
        }
    if (pairs_out) *pairs_out = pairs;
    return total;
}

// The as-called production path from raw histories, exactly as mpi_comparison_test.cc:71-103 /
// FE_problem.h:1229-1235 drive it: build Strain6D objects, splinify, set_ID,
// compare_histories_with_all_ranks (single rank), then one most_similar_histories_to_file per
// history. fname_pattern is a printf pattern with one %u (the ID), e.g. "dir/last.%u.similar_hist".
// If spline_out != NULL it receives [N][6P]. Pass fname_pattern == NULL to skip the files.
void ref_pipeline(const double *steps, const uint64_t *offsets, const uint32_t *ids, uint64_t N,
                  uint32_t P, double thr, const char *fname_pattern, double *spline_out)
{
    std::vector<Strain6D *> hist;
    for (uint64_t i = 0; i < N; i++) {
        Strain6D *h = new Strain6D();
        for (uint64_t n = offsets[i]; n < offsets[i + 1]; n++) {
            const double *s = steps + 6 * n;
            h->add_current_strain(s[0], s[1], s[2], s[3], s[4], s[5]);
        }
        h->splinify(P);
        h->set_ID(ids[i]);
        if (spline_out)
            memcpy(spline_out + (size_t)i * 6 * P, h->get_spline()->data(), sizeof(double) * 6 * P);
        hist.push_back(h);
    }
    MatHistPredict::compare_histories_with_all_ranks(hist, thr, MPI_COMM_WORLD);
    if (fname_pattern) {
        char buf[4096];
        for (uint64_t i = 0; i < N; i++) {
            snprintf(buf, sizeof buf, fname_pattern, hist[i]->get_ID());
            hist[i]->most_similar_histories_to_file(buf);
        }
    }
    for (uint64_t i = 0; i < N; i++) delete hist[i];
}

// Strain6D::from_file (strain2spline.h:112-134) on one file; the parsed steps come back as
// [L][6] (xx yy zz xy xz yz). Returns the number of steps read (may exceed cap_steps; then only
// the first cap_steps were copied).
uint64_t ref_from_file(const char *path, double *out, uint64_t cap_steps)
{
    Strain6D h;
    h.from_file(path);
    const uint64_t L = h.in_XX.size();
    for (uint64_t n = 0; n < L && n < cap_steps; n++) {
        out[6 * n + 0] = h.in_XX[n]; out[6 * n + 1] = h.in_YY[n]; out[6 * n + 2] = h.in_ZZ[n];
        out[6 * n + 3] = h.in_XY[n]; out[6 * n + 4] = h.in_XZ[n]; out[6 * n + 5] = h.in_YZ[n];
    }
    return L;
}

// Default-ostream formatting of a double, as every reference writer uses it
// (strain2spline.h:310, compare_all_histories.cc:73). Returns the length written.
int ref_format_double(double v, char *out, int cap)
{
    std::ostringstream os;
    os << v;
    std::string s = os.str();
    int n = (int)std::min<size_t>(s.size(), (size_t)cap - 1);
    memcpy(out, s.data(), n);
    out[n] = 0;
    return n;
}

}  // extern "C"
