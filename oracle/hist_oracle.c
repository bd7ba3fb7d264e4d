/* TEST INFRASTRUCTURE ONLY — CPU restatement (plain C) of SCEMa's MD-redundancy clustering path.
 *
 * This is the parity oracle for the CUDA path. Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it; the product
 * (scema_b200/libscema_hist.so) never links, loads or calls it.
 *
 * PARITY PIN: the reference ships no tests or golden vectors for this path.
 * The pin is the reference header itself, compiled unmodified into oracle/_ref/libscema_ref.so
 * (oracle/Makefile, oracle/ref_harness.cc); tests/test_oracle_vs_ref.py checks this file
 * bit-for-bit against it whenever oracle/_ref exists, and tests/golden/ holds vectors generated
 * from it by tests/golden/make_golden.py (committed, so the check also runs where
 * /root/reference is absent).
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp (no FMA contraction: the reference's Makefile builds
 * plain x86-64, clustering/Makefile:4,7; contraction changes spline samples by 1-3 ulp).
 *
 * All file:line citations are relative to /root/reference/.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * Natural cubic spline through (i/(L-1), y_i), sampled at p/(P-1).
 * Follows Strain6D::splinify (headers/strain2spline.h:140-180) which calls, per component,
 * tk::spline::set_points (headers/spline.h:284-373) -> band_matrix::lu_solve (:252-263)
 * -> lu_decompose (:187-220), l_solve (:222-235), r_solve (:237-250), and then
 * tk::spline::operator() (:375-396) at each sample. The ORDER of floating-point operations
 * below is the reference's; do not re-associate.
 *
 * work must hold 9*L doubles. y has stride ystride (6 for the [L][6] layout). out has stride
 * ostride (6 for the interleaved p*6+c layout, strain2spline.h:170-177).
 * ------------------------------------------------------------------------------------------ */
static void spline_component(const double *y, size_t ystride, uint32_t L, uint32_t P, double *out,
                             size_t ostride, double *work)
{
    const double third = 1.0 / 3.0, twothird = 2.0 / 3.0; /* spline.h:303-305 constants */
    double *x = work, *lo = x + L, *di = lo + L, *up = di + L, *sd = up + L, *rhs = sd + L,
           *z = rhs + L, *b = z + L, *spare = b + L;
    (void)spare;
    const int n = (int)L;

    /* knots: strain2spline.h:156-159 */
    for (int i = 0; i < n; i++) x[i] = (double)i / (double)(L - 1);

    /* system rows: spline.h:302-307; natural boundary rows :309-313, :323-327 */
    for (int i = 1; i < n - 1; i++) {
        lo[i] = third * (x[i] - x[i - 1]);
        di[i] = twothird * (x[i + 1] - x[i - 1]);
        up[i] = third * (x[i + 1] - x[i]);
        rhs[i] = (y[(size_t)(i + 1) * ystride] - y[(size_t)i * ystride]) / (x[i + 1] - x[i]) -
                 (y[(size_t)i * ystride] - y[(size_t)(i - 1) * ystride]) / (x[i] - x[i - 1]);
    }
    di[0] = 2.0; up[0] = 0.0; lo[0] = 0.0; rhs[0] = 0.0;
    di[n - 1] = 2.0; lo[n - 1] = 0.0; up[n - 1] = 0.0; rhs[n - 1] = 0.0;

    /* preconditioning: spline.h:195-204 (row i scaled by 1/diag, diag forced to 1) */
    for (int i = 0; i < n; i++) {
        sd[i] = 1.0 / di[i];
        if (i > 0) lo[i] *= sd[i];
        if (i < n - 1) up[i] *= sd[i];
        di[i] = 1.0;
    }
    /* elimination: spline.h:207-219 (one sub-diagonal, one super-diagonal) */
    for (int k = 0; k < n - 1; k++) {
        double xx = -lo[k + 1] / di[k];
        lo[k + 1] = -xx;
        di[k + 1] = di[k + 1] + xx * up[k];
    }
    /* forward substitution: spline.h:228-233 */
    for (int i = 0; i < n; i++) {
        double sum = 0;
        if (i > 0) sum += lo[i] * z[i - 1];
        z[i] = (rhs[i] * sd[i]) - sum;
    }
    /* back substitution: spline.h:243-248 */
    for (int i = n - 1; i >= 0; i--) {
        double sum = 0;
        if (i < n - 1) sum += up[i] * b[i + 1];
        b[i] = (z[i] - sum) / di[i];
    }

    /* sampling: strain2spline.h:170-177 calling spline.h:375-396 */
    for (uint32_t p = 0; p < P; p++) {
        double t = (double)p / (double)(P - 1);
        /* std::lower_bound: first knot with !(x < t) (spline.h:380), then idx = max(it-1, 0) */
        int it = 0;
        while (it < n && x[it] < t) it++;
        int idx = it - 1 > 0 ? it - 1 : 0;
        double h = t - x[idx];
        double v;
        if (t < x[0]) {
            /* left extrapolation (spline.h:385-387): unreachable for t in [0,1]; kept for NaN
             * fidelity only (comparisons with NaN are false, so NaN falls through below). */
            double c0 = (y[ystride] - y[0]) / (x[1] - x[0]) - third * (2.0 * b[0] + b[1]) * (x[1] - x[0]);
            v = (b[0] * h + c0) * h + y[0];
        } else if (t > x[n - 1]) {
            v = NAN; /* right extrapolation: unreachable, t <= 1 == x[n-1] */
        } else {
            /* coefficients of interval idx: spline.h:345-349 (idx <= n-2 always since t<=x[n-1]) */
            int i = idx < n - 1 ? idx : n - 2;
            double a_i = third * (b[i + 1] - b[i]) / (x[i + 1] - x[i]);
            double c_i = (y[(size_t)(i + 1) * ystride] - y[(size_t)i * ystride]) / (x[i + 1] - x[i]) -
                         third * (2.0 * b[i] + b[i + 1]) * (x[i + 1] - x[i]);
            v = ((a_i * h + b[i]) * h + c_i) * h + y[(size_t)i * ystride]; /* spline.h:393 */
        }
        out[(size_t)p * ostride] = v;
    }
}

/* steps: [L][6]; out: [6P] interleaved p*6+c. Returns 0, or 1 if L<3 (the reference exits,
 * strain2spline.h:142-148) or P==0. */
int oracle_splinify(const double *steps, uint32_t L, uint32_t P, double *out)
{
    if (L < 3 || P == 0) return 1;
    double *work = (double *)malloc(sizeof(double) * 9 * (size_t)L);
    if (!work) return 2;
    for (int c = 0; c < 6; c++) spline_component(steps + c, 6, L, P, out + c, 6, work);
    free(work);
    return 0;
}

int oracle_splinify_batch(const double *steps, const uint64_t *offsets, uint64_t N, uint32_t P,
                          double *out, int nthreads)
{
    int err = 0;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads) reduction(| : err)
#endif
    for (int64_t i = 0; i < (int64_t)N; i++) {
        err |= oracle_splinify(steps + 6 * offsets[i], (uint32_t)(offsets[i + 1] - offsets[i]), P,
                               out + (size_t)i * 6 * P);
    }
    return err;
}

/* compare_L2_norm, strain2spline.h:469-484: sequential k, multiply then add, IEEE sqrt. */
double oracle_compare_l2(const double *a, const double *b, uint32_t K)
{
    double sum = 0;
    for (uint32_t k = 0; k < K; k++) {
        double diff = a[k] - b[k];
        sum += diff * diff;
    }
    return sqrt(sum);
}

typedef struct { uint32_t i, j; double d; } oracle_edge;

static int edge_cmp(const void *pa, const void *pb)
{
    const oracle_edge *a = (const oracle_edge *)pa, *b = (const oracle_edge *)pb;
    if (a->i != b->i) return a->i < b->i ? -1 : 1;
    if (a->j != b->j) return a->j < b->j ? -1 : 1;
    return 0;
}

/* All unordered pairs i<j, i in [row_begin,row_end): the local double loop of
 * compare_histories_with_all_ranks (strain2spline.h:603-611) with the strict threshold of
 * choose_most_similar_history (:272). Output sorted by (i,j). Returns #edges found (may exceed
 * cap; only min(found,cap) are stored). */
uint64_t oracle_all_pairs(const double *rows, uint64_t N, uint32_t K, double thr, uint64_t row_begin,
                          uint64_t row_end, int nthreads, uint32_t *ei, uint32_t *ej, double *ed,
                          uint64_t cap, uint64_t *pairs_out)
{
    if (row_end > N) row_end = N;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    oracle_edge **buf = (oracle_edge **)calloc(nthreads, sizeof(*buf));
    uint64_t *cnt = (uint64_t *)calloc(nthreads, sizeof(*cnt));
    uint64_t *capv = (uint64_t *)calloc(nthreads, sizeof(*capv));
    uint64_t pairs = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 8) num_threads(nthreads) reduction(+ : pairs)
#endif
    for (int64_t i = (int64_t)row_begin; i < (int64_t)row_end; i++) {
#ifdef _OPENMP
        int t = omp_get_thread_num();
#else
        int t = 0;
#endif
        const double *a = rows + (size_t)i * K;
        for (uint64_t j = (uint64_t)i + 1; j < N; j++) {
            double d = oracle_compare_l2(a, rows + (size_t)j * K, K);
            if (d < thr) {
                if (cnt[t] == capv[t]) {
                    capv[t] = capv[t] ? 2 * capv[t] : 1024;
                    buf[t] = (oracle_edge *)realloc(buf[t], capv[t] * sizeof(oracle_edge));
                }
                buf[t][cnt[t]].i = (uint32_t)i;
                buf[t][cnt[t]].j = (uint32_t)j;
                buf[t][cnt[t]].d = d;
                cnt[t]++;
            }
        }
        pairs += N - 1 - (uint64_t)i;
    }
    uint64_t total = 0;
    for (int t = 0; t < nthreads; t++) total += cnt[t];
    oracle_edge *all = (oracle_edge *)malloc((total ? total : 1) * sizeof(oracle_edge));
    uint64_t o = 0;
    for (int t = 0; t < nthreads; t++) {
        if (cnt[t]) memcpy(all + o, buf[t], cnt[t] * sizeof(oracle_edge));
        o += cnt[t];
        free(buf[t]);
    }
    qsort(all, total, sizeof(oracle_edge), edge_cmp);
    if (ei && ej && ed)
        for (uint64_t e = 0; e < total && e < cap; e++) {
            ei[e] = all[e].i; ej[e] = all[e].j; ed[e] = all[e].d;
        }
    free(all); free(buf); free(cnt); free(capv);
    if (pairs_out) *pairs_out = pairs;
    return total;
}

/* Re-check a list of candidate pairs by direct differences (used to verify every emitted edge at
 * sizes where the full N^2 oracle is infeasible). Returns the number of entries
 * whose recomputed distance differs in bits from ed[] or fails `d < thr`. */
uint64_t oracle_check_edges(const double *rows, uint32_t K, double thr, const uint32_t *ei,
                            const uint32_t *ej, const double *ed, uint64_t n, int nthreads)
{
    uint64_t bad = 0;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(nthreads) reduction(+ : bad)
#endif
    for (int64_t e = 0; e < (int64_t)n; e++) {
        double d = oracle_compare_l2(rows + (size_t)ei[e] * K, rows + (size_t)ej[e] * K, K);
        if (!(d < thr) || memcmp(&d, &ed[e], sizeof d) != 0) bad++;
    }
    return bad;
}

/* Default std::ostream formatting of a double (precision 6, general) as used by
 * most_similar_histories_to_file (strain2spline.h:310) == printf("%g"). */
int oracle_format_double(double v, char *out, int cap)
{
    return snprintf(out, (size_t)cap, "%g", v);
}

/* Per-history similarity files, strain2spline.h:301-314 as driven by FE_problem.h:1232-1235 and
 * mpi_comparison_test.cc:99-103: one file per compared history (created even when empty), one
 * line "<ID> <otherID> <diff>\n" per partner. Single-rank partner order (strain2spline.h:603-611):
 * history at index k lists partners in ascending index. Edges must be sorted by (i,j), i<j.
 * fname_pattern holds one %u. Returns 0 or the number of files that could not be opened. */
int oracle_write_similar_files(const uint32_t *ids, uint64_t N, const uint32_t *ei,
                               const uint32_t *ej, const double *ed, uint64_t n_edges,
                               const char *fname_pattern)
{
    /* CSR over directed copies, keeping ascending partner order */
    uint64_t *deg = (uint64_t *)calloc(N + 1, sizeof(uint64_t));
    for (uint64_t e = 0; e < n_edges; e++) { deg[ei[e] + 1]++; deg[ej[e] + 1]++; }
    for (uint64_t i = 0; i < N; i++) deg[i + 1] += deg[i];
    uint64_t *fill = (uint64_t *)malloc((N ? N : 1) * sizeof(uint64_t));
    memcpy(fill, deg, N * sizeof(uint64_t));
    uint32_t *other = (uint32_t *)malloc((2 * n_edges + 1) * sizeof(uint32_t));
    double *dist = (double *)malloc((2 * n_edges + 1) * sizeof(double));
    /* pass 1: partners with smaller index (edges (i,k), sorted by i for fixed k because the
     * list is sorted by (i,j)); pass 2: partners with larger index */
    for (uint64_t e = 0; e < n_edges; e++) { other[fill[ej[e]]] = ei[e]; dist[fill[ej[e]]++] = ed[e]; }
    for (uint64_t e = 0; e < n_edges; e++) { other[fill[ei[e]]] = ej[e]; dist[fill[ei[e]]++] = ed[e]; }
    int failed = 0;
    char name[4096], num[64];
    for (uint64_t k = 0; k < N; k++) {
        snprintf(name, sizeof name, fname_pattern, ids[k]);
        FILE *f = fopen(name, "w");
        if (!f) { failed++; continue; }
        for (uint64_t q = deg[k]; q < deg[k + 1]; q++) {
            oracle_format_double(dist[q], num, sizeof num);
            fprintf(f, "%u %u %s\n", ids[k], ids[other[q]], num);
        }
        fclose(f);
    }
    free(deg); free(fill); free(other); free(dist);
    return failed;
}
