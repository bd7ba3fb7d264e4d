// K1 — batched ragged natural-cubic-spline fit + resample (sm_100a).
//
// Replaces, for a whole batch of quadrature-point histories, Strain6D::splinify
// (reference headers/strain2spline.h:140-180), i.e. per component tk::spline::set_points
// (headers/spline.h:284-373) -> band_matrix::lu_decompose/l_solve/r_solve (:187-250) followed by
// tk::spline::operator() (:375-396) at spline_points equally spaced abscissae.
//
// Bit-exactness: every floating-point operation below is issued in the reference's order with
// explicit round-to-nearest intrinsics (__dadd_rn/__dsub_rn/__dmul_rn/__ddiv_rn are never
// contracted into FMAs), including the `0.0 + x` of the solver's `sum` accumulators, which
// matters for the sign of zero.
//
// Structure (DESIGN.md "K1"):
//  * The tridiagonal matrix depends only on the history length L (knots are i/(L-1)), so its
//    preconditioned LU factors are tabulated once per distinct L (k_build_tables), together with
//    the sample -> interval map for the current spline_points.
//  * k_resample_staged: one warp per group of five histories. The TMA engine copies each
//    history's contiguous [L][6] block into the warp's shared-memory slab (cp.async.bulk +
//    mbarrier: coalesced, no register staging); lane = (history-in-group, component) then runs
//    the two sequential sweeps of its own chain out of shared memory (30 of 32 lanes busy),
//    overwriting y by z and b in place; finally the whole warp evaluates the 5*6*P samples and
//    stores them in the reference's p*6+c order (one contiguous 48*P-byte row per history).
//  * Histories are processed in length classes (one launch each) so the slab, and with it the
//    number of resident warps per SM, is sized for the class and not for the longest history.
//  * k_resample_global: fallback for histories longer than a slab (L > 800).
#include "common.cuh"
#include <algorithm>

namespace scema {

// table for one L, every array padded to an even length Lp (so each starts 16-byte aligned and the
// five arrays the sweeps need are one contiguous block for a bulk copy):
//   x[Lp] | hd[Lp] sd[Lp] lo[Lp] up[Lp] di[Lp] | ht[Pp] idx[Pp] | FW[Lp][4] | BW[Lp][4]
// FW[i] = {1/hd_i, hd_i, sd_i, lo_i} and BW[i] = {up_i, di_i, 1/di_i, 0} are what one forward / one
// backward step of k_resample_stream needs, packed so that each step is two 16-byte loads; the
// correctly rounded reciprocals feed the exact division div_tab() below.
__host__ __device__ inline uint32_t pad2(uint32_t v) { return v + (v & 1u); }
__host__ __device__ inline uint64_t table_fw_offset(uint32_t L, uint32_t P) { return 6ull * pad2(L) + 2ull * pad2(P); }
__host__ __device__ inline uint64_t table_doubles(uint32_t L, uint32_t P) { return table_fw_offset(L, P) + 8ull * pad2(L); }

__global__ void k_build_tables(const uint32_t *__restrict__ lens, const uint64_t *__restrict__ offs,
                               uint32_t n_tables, uint32_t P, double *__restrict__ tables)
{
    uint32_t ti = blockIdx.x * blockDim.x + threadIdx.x;
    if (ti >= n_tables) return;
    const uint32_t L = lens[ti];
    const int n = (int)L;
    const uint32_t Lp = pad2(L);
    double *x = tables + offs[ti], *hd = x + Lp, *sd = hd + Lp, *lo = sd + Lp, *up = lo + Lp, *di = up + Lp,
           *ht = di + Lp, *ix = ht + pad2(P);
    if (Lp != L) x[L] = hd[L] = sd[L] = lo[L] = up[L] = di[L] = 0.0;
    const double third = 1.0 / 3.0, twothird = 2.0 / 3.0;  // spline.h:303-305
    for (int i = 0; i < n; i++) x[i] = __ddiv_rn((double)i, (double)(L - 1));  // strain2spline.h:157
    for (int i = 0; i < n; i++) hd[i] = i < n - 1 ? __dsub_rn(x[i + 1], x[i]) : 0.0;
    // rows, spline.h:302-305 and natural boundary rows :309-313, :323-327
    for (int i = 1; i < n - 1; i++) {
        lo[i] = __dmul_rn(third, __dsub_rn(x[i], x[i - 1]));
        di[i] = __dmul_rn(twothird, __dsub_rn(x[i + 1], x[i - 1]));
        up[i] = __dmul_rn(third, __dsub_rn(x[i + 1], x[i]));
    }
    di[0] = 2.0; up[0] = 0.0; lo[0] = 0.0;
    di[n - 1] = 2.0; lo[n - 1] = 0.0; up[n - 1] = 0.0;
    // preconditioning, spline.h:195-204
    for (int i = 0; i < n; i++) {
        sd[i] = __ddiv_rn(1.0, di[i]);
        if (i > 0) lo[i] = __dmul_rn(lo[i], sd[i]);
        if (i < n - 1) up[i] = __dmul_rn(up[i], sd[i]);
        di[i] = 1.0;
    }
    // elimination, spline.h:207-219
    for (int k = 0; k < n - 1; k++) {
        double xx = __ddiv_rn(-lo[k + 1], di[k]);
        lo[k + 1] = -xx;
        di[k + 1] = __dadd_rn(di[k + 1], __dmul_rn(xx, up[k]));
    }
    // sample -> interval map, strain2spline.h:171 and spline.h:380-383
    for (uint32_t p = 0; p < P; p++) {
        double t = __ddiv_rn((double)p, (double)(P - 1));
        int it = 0;
        while (it < n && x[it] < t) it++;  // std::lower_bound on the rounded knots
        int idx = it - 1 > 0 ? it - 1 : 0;
        if (idx > n - 2) idx = n - 2;      // t <= x[n-1] always, so this never binds
        ht[p] = __dsub_rn(t, x[idx]);
        ix[p] = (double)idx;
    }
    double *fw = tables + offs[ti] + table_fw_offset(L, P), *bw = fw + 4ull * Lp;
    for (uint32_t i = 0; i < Lp; i++) {
        const bool in = i < L;
        fw[4 * i + 0] = in && i < L - 1 ? __ddiv_rn(1.0, hd[i]) : 0.0;
        fw[4 * i + 1] = in ? hd[i] : 0.0;
        fw[4 * i + 2] = in ? sd[i] : 0.0;
        fw[4 * i + 3] = in ? lo[i] : 0.0;
        bw[4 * i + 0] = in ? up[i] : 0.0;
        bw[4 * i + 1] = in ? di[i] : 1.0;
        bw[4 * i + 2] = in ? __ddiv_rn(1.0, di[i]) : 1.0;
        bw[4 * i + 3] = 0.0;
    }
}

constexpr int GROUP = 5;  // histories per warp (5*6 = 30 chain lanes)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- staged kernel: each history's [L][6] block is brought into the warp's shared-memory slab by
// one 1-D bulk copy of the TMA engine (cp.async.bulk, completion on an mbarrier); the sweeps run out
// of shared memory and overwrite y_i by z_i and then b_i in place. History hh of the group starts at
// a slab offset == 6*hh (mod 16 doubles), which makes the lock-step accesses of the 30 chain lanes
// (stride 6 doubles per step) shared-memory bank-conflict free.
__global__ void __launch_bounds__(32) k_resample_staged(const double *__restrict__ steps,
                                                        const uint64_t *__restrict__ offsets,
                                                        const uint32_t *__restrict__ order, uint64_t first, uint64_t count,
                                                        const int64_t *__restrict__ table_index,
                                                        const double *__restrict__ tables, uint32_t P,
                                                        double *__restrict__ out, uint32_t cap)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    double *tslot = reinterpret_cast<double *>(smem_raw + 16);  // hd sd lo up di of one length
    double *slab = tslot + 5 * (size_t)pad2(cap);
    const int lane = threadIdx.x;
    const uint32_t K = 6 * P;
    const double third = 1.0 / 3.0;
    const uint64_t n_groups = (count + GROUP - 1) / GROUP;
    int slot_len = -1;  // length whose factor table currently sits in tslot
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t phase = 0;

    for (uint64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const uint64_t q0 = first + grp * GROUP;
        const uint32_t n_here = (uint32_t)((count - grp * GROUP) < (uint64_t)GROUP ? (count - grp * GROUP) : (uint64_t)GROUP);
        // group geometry, computed redundantly by every lane
        uint64_t h_idx[GROUP], h_off[GROUP];
        int h_len[GROUP];
        uint32_t h_base[GROUP];
        uint32_t total_bytes = 0, next = 0;
#pragma unroll
        for (int j = 0; j < GROUP; j++) {
            h_idx[j] = 0; h_off[j] = 0; h_len[j] = 0; h_base[j] = 0;
            if ((uint32_t)j < n_here) {
                h_idx[j] = order ? (uint64_t)order[q0 + j] : q0 + j;
                h_off[j] = offsets[h_idx[j]];
                h_len[j] = (int)(offsets[h_idx[j] + 1] - h_off[j]);
                uint32_t want = (6u * j) & 15u;
                next += (want + 16u - (next & 15u)) & 15u;
                h_base[j] = next;
                next += 6u * h_len[j];
                total_bytes += 48u * h_len[j];
            }
        }
        // the group is sorted by length, so one table usually serves all five histories; it stays in
        // the slot across groups until the length changes
        const bool load_table = h_len[0] != slot_len;
        const uint32_t table_bytes = 5u * pad2((uint32_t)h_len[0]) * 8u;
        if (load_table) { total_bytes += table_bytes; slot_len = h_len[0]; }
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(total_bytes) : "memory");
        __syncwarp();
        if (lane == GROUP && load_table)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(tslot)),
                         "l"(tables + table_index[h_len[0]] + pad2((uint32_t)h_len[0])), "r"(table_bytes), "r"(smem_u32(bar))
                         : "memory");
#pragma unroll
        for (int j = 0; j < GROUP; j++)
            if (lane == j && (uint32_t)j < n_here)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(slab + h_base[j])),
                             "l"(steps + h_off[j] * 6), "r"(48u * h_len[j]), "r"(smem_u32(bar))
                             : "memory");
        {
            uint32_t ok = 0;
            while (!ok)
                asm volatile(
                    "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                    : "=r"(ok)
                    : "r"(smem_u32(bar)), "r"(phase)
                    : "memory");
            phase ^= 1;
        }

        const int hh = lane / 6, c = lane - hh * 6;
        int L = 0;
        uint32_t base = 0;
#pragma unroll
        for (int j = 0; j < GROUP; j++)
            if (j == hh) { L = h_len[j]; base = h_base[j]; }
        const bool chain = lane < GROUP * 6 && (uint32_t)hh < n_here;
        if (chain) {
            // factor table: the shared-memory slot when this history has the slot's length, else global
            const uint32_t Lp = pad2((uint32_t)L);
            const double *hd = L == slot_len ? tslot : tables + table_index[L] + Lp;
            const double *sd = hd + Lp, *lo = sd + Lp, *up = lo + Lp, *di = up + Lp;
            double *ys = slab + base + c;
            // ---- forward substitution fused with the right-hand side (spline.h:306, :228-233)
            double y1 = ys[0], y2 = ys[6];
            double s_prev = __ddiv_rn(__dsub_rn(y2, y1), hd[0]);
            double z_prev = __dsub_rn(__dmul_rn(0.0, sd[0]), 0.0);  // row 0: rhs = 0, empty sum
            ys[0] = z_prev;
            y1 = y2;
#pragma unroll 4
            for (int i = 1; i < L - 1; i++) {
                y2 = ys[(i + 1) * 6];
                double s_cur = __ddiv_rn(__dsub_rn(y2, y1), hd[i]);
                double r = __dmul_rn(__dsub_rn(s_cur, s_prev), sd[i]);
                double sum = __dadd_rn(0.0, __dmul_rn(lo[i], z_prev));
                z_prev = __dsub_rn(r, sum);
                ys[i * 6] = z_prev;
                s_prev = s_cur;
                y1 = y2;
            }
            {
                double r = __dmul_rn(0.0, sd[L - 1]);  // row L-1: rhs = 0
                double sum = __dadd_rn(0.0, __dmul_rn(lo[L - 1], z_prev));
                z_prev = __dsub_rn(r, sum);
            }
            // ---- back substitution (spline.h:243-248); b overwrites z
            double b_next = __ddiv_rn(__dsub_rn(z_prev, 0.0), di[L - 1]);
            ys[(L - 1) * 6] = b_next;
#pragma unroll 4
            for (int i = L - 2; i >= 0; i--) {
                double sum = __dadd_rn(0.0, __dmul_rn(up[i], b_next));
                b_next = __ddiv_rn(__dsub_rn(ys[i * 6], sum), di[i]);
                ys[i * 6] = b_next;
            }
        }
        __syncwarp();

        // ---- evaluation at the P sample points, all 32 lanes (spline.h:345-349, :393)
        for (uint32_t o = lane; o < n_here * K; o += 32) {
            const uint32_t eh = o / K, k = o - eh * K, p = k / 6, ec = k - p * 6;
            uint64_t eidx = 0, eoff = 0;
            int eL = 0;
            uint32_t ebase = 0;
#pragma unroll
            for (int j = 0; j < GROUP; j++)
                if ((uint32_t)j == eh) { eidx = h_idx[j]; eoff = h_off[j]; eL = h_len[j]; ebase = h_base[j]; }
            const double *etab = tables + table_index[eL];
            const double *ehd = etab + pad2((uint32_t)eL), *eht = etab + 6 * (size_t)pad2((uint32_t)eL), *eix = eht + pad2(P);
            const int idx = (int)__ldg(eix + p);
            const double hstep = __ldg(eht + p), hdv = __ldg(ehd + idx);
            const double *ey = steps + eoff * 6 + ec;
            const double ya = __ldg(ey + (size_t)idx * 6), yb = __ldg(ey + (size_t)(idx + 1) * 6);
            const double b0 = slab[ebase + idx * 6 + ec], b1 = slab[ebase + (idx + 1) * 6 + ec];
            const double a_i = __ddiv_rn(__dmul_rn(third, __dsub_rn(b1, b0)), hdv);
            const double c_i = __dsub_rn(__ddiv_rn(__dsub_rn(yb, ya), hdv),
                                         __dmul_rn(__dmul_rn(third, __dadd_rn(__dmul_rn(2.0, b0), b1)), hdv));
            double v = __dadd_rn(__dmul_rn(a_i, hstep), b0);
            v = __dadd_rn(__dmul_rn(v, hstep), c_i);
            v = __dadd_rn(__dmul_rn(v, hstep), ya);
            out[eidx * K + k] = v;
        }
        // the slab was written through the generic proxy; order that before the next bulk copy
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
    }
}


// ---- streamed kernel (K1 v4). One warp per group of five histories, lane = (history, component)
// chain as above, but nothing is staged in shared memory, so the number of resident chains is not
// capped by the slab size: y is read straight from global memory through an 8-deep register ring
// (each 48-byte step of a history is one sector pair, every line is used by 2.7 consecutive
// steps), z goes to a warp-private scratch column block zs[i][32] (coalesced 256-byte rows that
// are re-read by the same warp ~L steps later, i.e. out of L2), and the samples are evaluated
// inside the backward sweep at the moment b_idx and b_idx+1 exist, so b is never stored.
//
// div_tab(a, b, rb) == __ddiv_rn(a, b) for a table divisor b with rb = RN(1/b): two
// Newton-style FMA corrections of q = a*rb. After the first, q is a faithful rounding of a/b;
// Markstein's theorem then makes q + (a - b*q)*rb round to RN(a/b). Guarded to numerators whose
// exponent keeps every intermediate normal; zeros, subnormals, huge values, inf and NaN take the
// IEEE division. 5 dependent FP64 ops (~40 cycles) instead of ~110 for the division sequence.
// Checked against the true quotient for 6e8 (a,b) pairs by tests/test_fastdiv.py.
__device__ __forceinline__ double div_tab(double a, double b, double rb)
{
    const uint32_t e = ((uint32_t)__double2hiint(a) >> 20) & 0x7ffu;
    if (e - 128u < 1792u) {
        double q = __dmul_rn(a, rb);
        double r = __fma_rn(-b, q, a);
        q = __fma_rn(r, rb, q);
        r = __fma_rn(-b, q, a);
        return __fma_rn(r, rb, q);
    }
    return __ddiv_rn(a, b);
}

constexpr int RS_WARPS = 4;  // warps per CTA (independent of each other)
constexpr int RING = 8;      // steps per unrolled block
constexpr int DEPTH = 16;    // slots of the per-lane shared-memory prefetch rings (y and z)
constexpr size_t RS_SMEM = (size_t)RS_WARPS * DEPTH * 32 * sizeof(double);

// The two long-latency streams of a chain — y on the way up, z on the way down — are prefetched
// DEPTH-2 steps ahead with 8-byte cp.async copies into a per-lane shared-memory ring (one commit
// group per step, cp.async.wait_group DEPTH-2 before the slot is read). They deliberately do NOT
// go through registers: a warp has six scoreboards, and with eight register-ring loads in flight
// the compiler had to put the short-latency factor-table loads on the same scoreboards as the
// DRAM-latency ring loads, so every step waited a full memory round trip
// (profiles/r01_ncu_resample_v6_c3.txt: ~1600 cycles per step). cp.async completion is counted
// per group instead, which leaves the scoreboards to the table loads.
__device__ __forceinline__ void cp_async8(double *dst_smem, const double *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// table entries of the coming step, loaded into the ping-pong variable they are consumed from
__device__ __forceinline__ void ld_tab(double2 &d, const double2 *p)
{
    asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(d.x), "=d"(d.y) : "l"(p));
}

// One forward step i = i0 + U (spline.h:306, :228-233). On entry: s_cur = s_i, s_prev = s_{i-1},
// z_prev = z_{i-1}, y_hi = y_{i+1}, E1 = {sd_i, lo_i}, E0 = {1/hd_{i+1}, hd_{i+1}}. Loads the same
// two entries for step i+1 into N1/N0, takes y_{i+2} from ring slot (i+2)%DEPTH and starts the copy
// of y_{i+1+DEPTH} into the slot read one step earlier.
#define K1_FWD_STEP(U, E1, E0, N1, N0)                                                              \
    if (i0 + (U) < L - 1) {                                                                         \
        const int i = i0 + (U);                                                                     \
        ld_tab(N1, FWb + 2 * ((U) + 1) + 1);                                                        \
        ld_tab(N0, FWb + 2 * ((U) + 2));                                                            \
        cp_async_wait<DEPTH - 2>();                                                                 \
        const double y_nx = ry[((i + 2) & (DEPTH - 1)) * 32];                                       \
        if (i + 1 + DEPTH < L) cp_async8(ry + ((i + 1) & (DEPTH - 1)) * 32, yb + (size_t)((U) + 1 + DEPTH) * ys); \
        cp_async_commit();                                                                          \
        const double r = __dmul_rn(__dsub_rn(s_cur, s_prev), E1.x);                                 \
        const double sum = __dadd_rn(0.0, __dmul_rn(E1.y, z_prev));                                 \
        z_prev = __dsub_rn(r, sum);                                                                 \
        __stcg(zb + (size_t)(U) * 32, z_prev);                                                      \
        s_prev = s_cur;                                                                             \
        if (i + 1 < L - 1) s_cur = div_tab(__dsub_rn(y_nx, y_hi), E0.y, E0.x);                      \
        y_hi = y_nx;                                                                                \
    }

// One backward step i = i0 - U (spline.h:243-248) plus the samples of interval i (spline.h:345-349,
// :393). On entry: b_next = b_{i+1}, W0 = {up_i, di_i}, W1 = {1/di_i, -}. Loads the entries of step
// i-1 into V0/V1, takes z_i from ring slot i%DEPTH and starts the copy of z_{i+1-DEPTH} into the
// slot read one step earlier.
#define K1_BWD_STEP(U, W0, W1, V0, V1)                                                              \
    if (i0 - (U) >= 0) {                                                                            \
        const int i = i0 - (U);                                                                     \
        ld_tab(V0, BWb - 2 * ((U) + 1));                                                            \
        ld_tab(V1, BWb - 2 * ((U) + 1) + 1);                                                        \
        cp_async_wait<DEPTH - 2>();                                                                 \
        const double zi = rz[(i & (DEPTH - 1)) * 32];                                               \
        if (i + 1 - DEPTH >= 0) cp_async8(rz + ((i + 1) & (DEPTH - 1)) * 32, zb - (size_t)((U) - 1 + DEPTH) * 32); \
        cp_async_commit();                                                                          \
        const double sum = __dadd_rn(0.0, __dmul_rn(W0.x, b_next));                                 \
        const double b_i = div_tab(__dsub_rn(zi, sum), W0.y, W1.x);                                 \
        if (i == nxt) {                                                                             \
            const double2 f0 = __ldg(FW + 2 * i);                                                   \
            const double hdv = f0.y;                                                                \
            const double a_i = div_tab(__dmul_rn(third, __dsub_rn(b_next, b_i)), hdv, f0.x);        \
            const double c_i =                                                                      \
                __dsub_rn(div_tab(__dsub_rn(y_b, y_a), hdv, f0.x),                                  \
                          __dmul_rn(__dmul_rn(third, __dadd_rn(__dmul_rn(2.0, b_i), b_next)), hdv)); \
            do {                                                                                    \
                const double hstep = __ldg(ht + p);                                                 \
                double v = __dadd_rn(__dmul_rn(a_i, hstep), b_i);                                   \
                v = __dadd_rn(__dmul_rn(v, hstep), c_i);                                            \
                v = __dadd_rn(__dmul_rn(v, hstep), y_a);                                            \
                orow[(size_t)p * 6] = v;                                                            \
                p--;                                                                                \
                nxt = p >= 0 ? (int)__ldg(ix + p) : -1;                                             \
            } while (nxt == i);                                                                     \
            if (nxt >= 0) { y_a = __ldg(y + (size_t)nxt * ys); y_b = __ldg(y + (size_t)(nxt + 1) * ys); } \
        }                                                                                           \
        b_next = b_i;                                                                               \
    }

__global__ void __launch_bounds__(32 * RS_WARPS) k_resample_stream(const double *__restrict__ steps,
                                                                  const uint64_t *__restrict__ offsets,
                                                                  const uint32_t *__restrict__ order, uint64_t first,
                                                                  uint64_t count, const int64_t *__restrict__ table_index,
                                                                  const double *__restrict__ tables, uint32_t P,
                                                                  double *__restrict__ out, double *__restrict__ zscratch,
                                                                  uint32_t cap, uint64_t ys, uint32_t uniform_L)
{
    // ys = distance (in doubles) between consecutive steps of one history: 6 for the ragged batch
    // ([L][6] blocks, history h starts at offsets[h]), n*6 for the time-major history store
    // ([step][n][6], history h starts at h, every history uniform_L steps long).
    extern __shared__ __align__(16) double rings[];
    const int lane = threadIdx.x & 31;
    // this lane's ring, ry[slot * 32]: y on the way up, then (all y copies have landed by then) z on
    // the way down. Keeping it to 4 KB per warp leaves most of the SM's 256 KB to L1, where the factor
    // tables live.
    double *ry = rings + (size_t)(threadIdx.x >> 5) * (DEPTH * 32) + lane;
    double *rz = ry;
    const uint64_t wid = (uint64_t)blockIdx.x * RS_WARPS + (threadIdx.x >> 5);
    const uint64_t n_warps = (uint64_t)gridDim.x * RS_WARPS;
    double *__restrict__ zs = zscratch + wid * cap * 32 + lane;
    const uint32_t K = 6 * P, Pp = pad2(P);
    const double third = 1.0 / 3.0;
    const uint64_t n_groups = (count + GROUP - 1) / GROUP;
    const int hh = lane / 6, c = lane - hh * 6;

    for (uint64_t grp = wid; grp < n_groups; grp += n_warps) {
        const uint64_t left = count - grp * GROUP;
        if (lane >= GROUP * 6 || (uint64_t)hh >= left) continue;  // no warp-level primitive below
        const uint64_t q = first + grp * GROUP + hh;
        const uint64_t h = order ? (uint64_t)order[q] : q;
        const uint64_t off = uniform_L ? h : offsets[h];
        const int L = uniform_L ? (int)uniform_L : (int)(offsets[h + 1] - off);
        const uint32_t Lp = pad2((uint32_t)L);
        const double *tab = tables + table_index[L];
        const double *ht = tab + 6ull * Lp, *ix = ht + Pp;
        const double2 *FW = reinterpret_cast<const double2 *>(ht + 2ull * Pp);
        const double2 *BW = FW + 2ull * Lp;
        const double *y = steps + off * 6 + c;

        // whole history -> L2 now (one bulk prefetch per history); the ring copies then hit L2
        if (c == 0 && !uniform_L)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(steps + off * 6), "r"(48u * (uint32_t)L) : "memory");

        // ---- forward substitution fused with the right-hand side. Software pipeline: the slope
        // s_{i+1} and the table entries of step i+1 are produced while the z chain of step i runs.
        // Table reads one entry past the end of FW / before the start of BW stay inside this
        // table's own allocation and their values are never consumed.
#pragma unroll
        for (int m = 3; m <= DEPTH + 1; m++) {  // y_3 .. y_{DEPTH+1}: DEPTH-1 groups
            if (m < L) cp_async8(ry + (m & (DEPTH - 1)) * 32, y + (size_t)m * ys);
            cp_async_commit();
        }
        double s_prev, s_cur, z_prev, y_hi;
        double2 a1, a0, b1, b0;  // ping-pong: {sd, lo} of the coming step and {1/hd, hd} of the one after
        {
            const double y_0 = __ldg(y), y_1 = __ldg(y + ys), y_2 = __ldg(y + 2 * ys);  // L >= 3
            const double2 f00 = __ldg(FW), f01 = __ldg(FW + 1), f10 = __ldg(FW + 2);
            ld_tab(a1, FW + 3);  // {sd_1, lo_1}
            ld_tab(a0, FW + 4);  // {1/hd_2, hd_2}
            s_prev = div_tab(__dsub_rn(y_1, y_0), f00.y, f00.x);  // s_0
            s_cur = div_tab(__dsub_rn(y_2, y_1), f10.y, f10.x);   // s_1
            z_prev = __dsub_rn(__dmul_rn(0.0, f01.x), 0.0);       // row 0: rhs = 0, empty sum
            __stcg(zs, z_prev);
            y_hi = y_2;
        }
        for (int i0 = 1; i0 < L - 1; i0 += RING) {
            const double2 *FWb = FW + 2 * i0;
            const double *yb = y + (size_t)i0 * ys;
            double *zb = zs + (size_t)i0 * 32;
            K1_FWD_STEP(0, a1, a0, b1, b0)
            K1_FWD_STEP(1, b1, b0, a1, a0)
            K1_FWD_STEP(2, a1, a0, b1, b0)
            K1_FWD_STEP(3, b1, b0, a1, a0)
            K1_FWD_STEP(4, a1, a0, b1, b0)
            K1_FWD_STEP(5, b1, b0, a1, a0)
            K1_FWD_STEP(6, a1, a0, b1, b0)
            K1_FWD_STEP(7, b1, b0, a1, a0)
        }
        {
            const double2 f1 = __ldg(FW + 2 * (L - 1) + 1);  // {sd, lo} of row L-1
            const double r = __dmul_rn(0.0, f1.x);           // rhs = 0
            const double sum = __dadd_rn(0.0, __dmul_rn(f1.y, z_prev));
            z_prev = __dsub_rn(r, sum);
        }

        // ---- back substitution with the samples evaluated on the way: sample p lives in interval
        // ix[p], non-increasing as p falls, so b is never stored
        cp_async_wait<0>();  // the ring changes hands: no y copy may still be in flight
#pragma unroll
        for (int u = 0; u < DEPTH - 1; u++) {  // z_{L-2} .. z_{L-DEPTH}: DEPTH-1 groups
            const int m = L - 2 - u;
            if (m >= 0) cp_async8(rz + (m & (DEPTH - 1)) * 32, zs + (size_t)m * 32);
            cp_async_commit();
        }
        double b_next;
        {
            const double2 w0 = __ldg(BW + 2 * (L - 1)), w1 = __ldg(BW + 2 * (L - 1) + 1);
            b_next = div_tab(__dsub_rn(z_prev, 0.0), w0.y, w1.x);
        }
        int p = (int)P - 1;
        int nxt = (int)__ldg(ix + p);
        double y_a = __ldg(y + (size_t)nxt * ys), y_b = __ldg(y + (size_t)(nxt + 1) * ys);
        double *orow = out + h * K + c;
        ld_tab(a0, BW + 2 * (L - 2));      // {up, di} of step L-2
        ld_tab(a1, BW + 2 * (L - 2) + 1);  // {1/di, -}
        for (int i0 = L - 2; i0 >= 0; i0 -= RING) {
            const double2 *BWb = BW + 2 * i0;
            const double *zb = zs + (size_t)i0 * 32;
            K1_BWD_STEP(0, a0, a1, b0, b1)
            K1_BWD_STEP(1, b0, b1, a0, a1)
            K1_BWD_STEP(2, a0, a1, b0, b1)
            K1_BWD_STEP(3, b0, b1, a0, a1)
            K1_BWD_STEP(4, a0, a1, b0, b1)
            K1_BWD_STEP(5, b0, b1, a0, a1)
            K1_BWD_STEP(6, a0, a1, b0, b1)
            K1_BWD_STEP(7, b0, b1, a0, a1)
        }
        cp_async_wait<0>();  // nothing of this group may land in the rings after the next group starts
    }
}
#undef K1_FWD_STEP
#undef K1_BWD_STEP

// ---- fallback for histories too long for a shared-memory slab: the sweeps read y from global
// memory and keep z/b in a global scratch buffer of the input's shape.
__global__ void __launch_bounds__(32) k_resample_global(const double *__restrict__ steps,
                                                        const uint64_t *__restrict__ offsets,
                                                        const uint32_t *__restrict__ order, uint64_t first, uint64_t count,
                                                        const int64_t *__restrict__ table_index,
                                                        const double *__restrict__ tables, uint32_t P,
                                                        double *__restrict__ out, double *__restrict__ zglobal)
{
    const int lane = threadIdx.x;
    const uint32_t K = 6 * P;
    const double third = 1.0 / 3.0;
    const uint64_t n_groups = (count + GROUP - 1) / GROUP;
    for (uint64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const uint64_t q0 = first + grp * GROUP;
        const uint32_t n_here = (uint32_t)((count - grp * GROUP) < (uint64_t)GROUP ? (count - grp * GROUP) : (uint64_t)GROUP);
        const int hh = lane / 6, c = lane - hh * 6;
        const bool chain = lane < GROUP * 6 && (uint32_t)hh < n_here;
        if (chain) {
            const uint64_t h = order ? (uint64_t)order[q0 + hh] : q0 + hh;
            const uint64_t off = offsets[h];
            const int L = (int)(offsets[h + 1] - off);
            const double *tab = tables + table_index[L];
            const double *y = steps + off * 6 + c;
            const uint32_t Lp = pad2((uint32_t)L);
            const double *hd = tab + Lp, *sd = hd + Lp, *lo = sd + Lp, *up = lo + Lp, *di = up + Lp;
            double *zg = zglobal + off * 6 + c;
            double y1 = __ldg(y), y2 = __ldg(y + 6);
            double s_prev = __ddiv_rn(__dsub_rn(y2, y1), __ldg(hd));
            double z_prev = __dsub_rn(__dmul_rn(0.0, __ldg(sd)), 0.0);
            zg[0] = z_prev;
            y1 = y2;
#pragma unroll 4
            for (int i = 1; i < L - 1; i++) {
                y2 = __ldg(y + (size_t)(i + 1) * 6);
                double s_cur = __ddiv_rn(__dsub_rn(y2, y1), __ldg(hd + i));
                double r = __dmul_rn(__dsub_rn(s_cur, s_prev), __ldg(sd + i));
                double sum = __dadd_rn(0.0, __dmul_rn(__ldg(lo + i), z_prev));
                z_prev = __dsub_rn(r, sum);
                zg[(size_t)i * 6] = z_prev;
                s_prev = s_cur;
                y1 = y2;
            }
            {
                double r = __dmul_rn(0.0, __ldg(sd + L - 1));
                double sum = __dadd_rn(0.0, __dmul_rn(__ldg(lo + L - 1), z_prev));
                z_prev = __dsub_rn(r, sum);
            }
            double b_next = __ddiv_rn(__dsub_rn(z_prev, 0.0), __ldg(di + L - 1));
            zg[(size_t)(L - 1) * 6] = b_next;
#pragma unroll 4
            for (int i = L - 2; i >= 0; i--) {
                double sum = __dadd_rn(0.0, __dmul_rn(__ldg(up + i), b_next));
                b_next = __ddiv_rn(__dsub_rn(zg[(size_t)i * 6], sum), __ldg(di + i));
                zg[(size_t)i * 6] = b_next;
            }
        }
        __threadfence_block();
        __syncwarp();
        for (uint32_t o = lane; o < n_here * K; o += 32) {
            const uint32_t eh = o / K, k = o - eh * K, p = k / 6, ec = k - p * 6;
            const uint64_t eidx = order ? (uint64_t)order[q0 + eh] : q0 + eh;
            const uint64_t eoff = offsets[eidx];
            const int eL = (int)(offsets[eidx + 1] - eoff);
            const double *etab = tables + table_index[eL];
            const double *ehd = etab + pad2((uint32_t)eL), *eht = etab + 6 * (size_t)pad2((uint32_t)eL), *eix = eht + pad2(P);
            const int idx = (int)__ldg(eix + p);
            const double hstep = __ldg(eht + p), hdv = __ldg(ehd + idx);
            const double *ey = steps + eoff * 6 + ec;
            const double ya = __ldg(ey + (size_t)idx * 6), yb = __ldg(ey + (size_t)(idx + 1) * 6);
            const double *g = zglobal + eoff * 6 + ec;
            const double b0 = g[(size_t)idx * 6], b1 = g[(size_t)(idx + 1) * 6];
            const double a_i = __ddiv_rn(__dmul_rn(third, __dsub_rn(b1, b0)), hdv);
            const double c_i = __dsub_rn(__ddiv_rn(__dsub_rn(yb, ya), hdv),
                                         __dmul_rn(__dmul_rn(third, __dadd_rn(__dmul_rn(2.0, b0), b1)), hdv));
            double v = __dadd_rn(__dmul_rn(a_i, hstep), b0);
            v = __dadd_rn(__dmul_rn(v, hstep), c_i);
            v = __dadd_rn(__dmul_rn(v, hstep), ya);
            out[eidx * K + k] = v;
        }
        __syncwarp();
    }
}

static int ensure_tables_present(scema_ctx *ctx, uint32_t P, const std::vector<uint8_t> &present, uint32_t max_len);

static int ensure_tables(scema_ctx *ctx, uint32_t P)
{
    // distinct lengths of the current batch
    std::vector<uint8_t> present((size_t)ctx->max_len + 1, 0);
    for (uint64_t i = 0; i < ctx->hn; i++) present[ctx->h_offsets[i + 1] - ctx->h_offsets[i]] = 1;
    return ensure_tables_present(ctx, P, present, ctx->max_len);
}

static int ensure_tables_present(scema_ctx *ctx, uint32_t P, const std::vector<uint8_t> &present, uint32_t max_len)
{
    if (ctx->table_P != P) { ctx->table_off.clear(); ctx->tables_used = 0; ctx->table_P = P; }
    std::vector<uint32_t> new_lens;
    std::vector<uint64_t> new_offs;
    uint64_t used = ctx->tables_used;
    for (uint32_t L = 3; L <= max_len; L++) {
        if (!present[L] || ctx->table_off.count(L)) continue;
        new_lens.push_back(L);
        new_offs.push_back(used);
        used += table_doubles(L, P);
    }
    bool index_stale = ctx->table_index_len < max_len + 1;
    if (new_lens.empty() && !index_stale) return SCEMA_OK;

    if (used * sizeof(double) > ctx->d_tables.bytes) {
        // grow: keep old tables by copying
        scema::DevBuf nb;
        SCEMA_CUDA(ctx, nb.reserve(used * sizeof(double) * 2));
        if (ctx->tables_used)
            SCEMA_CUDA(ctx, cudaMemcpyAsync(nb.p, ctx->d_tables.p, ctx->tables_used * sizeof(double),
                                            cudaMemcpyDeviceToDevice, ctx->stream));
        SCEMA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->d_tables.release();
        ctx->d_tables = nb;
    }
    if (!new_lens.empty()) {
        scema::DevBuf dl, dofs;
        SCEMA_CUDA(ctx, dl.reserve(new_lens.size() * sizeof(uint32_t)));
        SCEMA_CUDA(ctx, dofs.reserve(new_offs.size() * sizeof(uint64_t)));
        SCEMA_CUDA(ctx, cudaMemcpyAsync(dl.p, new_lens.data(), new_lens.size() * sizeof(uint32_t),
                                        cudaMemcpyHostToDevice, ctx->stream));
        SCEMA_CUDA(ctx, cudaMemcpyAsync(dofs.p, new_offs.data(), new_offs.size() * sizeof(uint64_t),
                                        cudaMemcpyHostToDevice, ctx->stream));
        uint32_t nt = (uint32_t)new_lens.size();
        k_build_tables<<<(nt + 31) / 32, 32, 0, ctx->stream>>>(dl.as<uint32_t>(), dofs.as<uint64_t>(), nt, P,
                                                               ctx->d_tables.as<double>());
        ctx->launches++;
        SCEMA_CUDA(ctx, cudaGetLastError());
        SCEMA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        dl.release(); dofs.release();
        for (size_t i = 0; i < new_lens.size(); i++) ctx->table_off[new_lens[i]] = new_offs[i];
        ctx->tables_used = used;
    }
    // L -> offset index
    const uint32_t index_len = std::max<uint32_t>(max_len + 1, ctx->table_index_len);
    std::vector<int64_t> index((size_t)index_len, -1);
    for (auto &kv : ctx->table_off)
        if (kv.first < index_len) index[kv.first] = (int64_t)kv.second;
    SCEMA_CUDA(ctx, ctx->d_table_index.reserve(index.size() * sizeof(int64_t)));
    SCEMA_CUDA(ctx, cudaMemcpyAsync(ctx->d_table_index.p, index.data(), index.size() * sizeof(int64_t),
                                    cudaMemcpyHostToDevice, ctx->stream));
    SCEMA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->table_index_len = index_len;
    return SCEMA_OK;
}

int resample_run(scema_ctx *ctx, uint32_t P)
{
    if (!ctx->have_histories) return fail(ctx, SCEMA_ERR_STATE, "resample: no histories set");
    if (P == 0) return fail(ctx, SCEMA_ERR_INVALID, "resample: spline_points must be >= 1");
    if (ctx->hn && ctx->min_len < 3)
        return fail(ctx, SCEMA_ERR_INVALID,
                    "Not enough strain steps added. Need at least 3 points for splinify().");
    const uint32_t K = 6 * P;
    SCEMA_CUDA(ctx, ctx->spline_own.reserve((size_t)(ctx->hn ? ctx->hn : 1) * K * sizeof(double)));
    ctx->d_spline = ctx->spline_own.as<double>();
    ctx->n = ctx->hn;
    ctx->ids = ctx->hist_ids;
    ctx->K = K;
    ctx->spline_version++;
    ctx->have_spline = true;
    ctx->have_edges = false;
    if (ctx->hn == 0) return SCEMA_OK;

    int rc = ensure_tables(ctx, P);
    if (rc) return rc;

    // Length classes: every class gets its own launch, with the warp-private z scratch (and, for
    // the staged variant, the shared-memory slab) sized for the class. Inside a class the work
    // order is sorted by exact length (stable counting sort), so the five histories of a warp
    // group almost always share one factor table and every warp of a launch gets a similar mix.
    // SCEMA_K1=staged selects the shared-memory-slab kernel (kept for A/B measurements).
    static const char *k1_env = getenv("SCEMA_K1");
    const bool staged = k1_env && strcmp(k1_env, "staged") == 0;
    static const uint32_t caps_staged[] = {16, 32, 48, 64, 96, 128, 192, 256, 384, 512, 800};
    static const uint32_t caps_stream[] = {64, 256, 2048, 16384, 131072};
    const uint32_t *caps = staged ? caps_staged : caps_stream;
    const int NCLS = (int)(staged ? sizeof(caps_staged) / sizeof(uint32_t) : sizeof(caps_stream) / sizeof(uint32_t)) + 1;
    constexpr int MAXCLS = 12;  // last class: global-scratch fallback
    auto cls_of = [&](uint64_t L) { int k = 0; while (k < NCLS - 1 && L > caps[k]) k++; return k; };
    uint64_t cls_count[MAXCLS] = {};
    for (uint64_t i = 0; i < ctx->hn; i++) cls_count[cls_of(ctx->h_offsets[i + 1] - ctx->h_offsets[i])]++;
    uint64_t cls_first[MAXCLS + 1] = {};
    for (int k = 0; k < NCLS; k++) cls_first[k + 1] = cls_first[k] + cls_count[k];
    const uint32_t *d_order = nullptr;
    if (ctx->min_len != ctx->max_len) {
        if (ctx->order_version != ctx->histories_version) {
            std::vector<uint64_t> len_first((size_t)ctx->max_len + 2, 0);
            for (uint64_t i = 0; i < ctx->hn; i++) len_first[ctx->h_offsets[i + 1] - ctx->h_offsets[i] + 1]++;
            for (uint32_t L = 0; L <= ctx->max_len; L++) len_first[L + 1] += len_first[L];
            std::vector<uint32_t> order(ctx->hn);
            for (uint64_t i = 0; i < ctx->hn; i++) order[len_first[ctx->h_offsets[i + 1] - ctx->h_offsets[i]]++] = (uint32_t)i;
            SCEMA_CUDA(ctx, ctx->d_order.reserve(ctx->hn * sizeof(uint32_t)));
            SCEMA_CUDA(ctx, cudaMemcpyAsync(ctx->d_order.p, order.data(), ctx->hn * sizeof(uint32_t), cudaMemcpyHostToDevice,
                                            ctx->stream));
            SCEMA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            ctx->order_version = ctx->histories_version;
        }
        d_order = ctx->d_order.as<uint32_t>();
    }
    auto smem_for = [&](uint32_t cap) { return (size_t)16 + 5 * (size_t)pad2(cap) * 8 + ((size_t)GROUP * 6 * cap + 64) * sizeof(double); };
    // streamed kernel: resident warps per SM and the scratch they need (one [cap][32] block per warp)
    static const char *wps_env = getenv("SCEMA_K1_WPS");
    const int wps = wps_env && atoi(wps_env) > 0 ? atoi(wps_env) : 24;
    const uint64_t scratch_budget = 1ull << 30;
    uint64_t stream_warps[MAXCLS] = {};
    uint64_t scratch_need = 0;
    if (cls_count[NCLS - 1]) scratch_need = (uint64_t)ctx->total_steps * 6 * sizeof(double);
    if (!staged)
        for (int k = 0; k < NCLS - 1; k++) {
            if (!cls_count[k]) continue;
            const uint32_t cap = std::min<uint32_t>(caps[k], ctx->max_len);
            uint64_t w = std::min<uint64_t>((uint64_t)ctx->sm_count * wps, (cls_count[k] + GROUP - 1) / GROUP);
            w = std::min<uint64_t>(w, std::max<uint64_t>(RS_WARPS, scratch_budget / ((uint64_t)cap * 256)));
            w = (w + RS_WARPS - 1) / RS_WARPS * RS_WARPS;
            stream_warps[k] = w;
            scratch_need = std::max<uint64_t>(scratch_need, w * cap * 256);
        }
    if (scratch_need) SCEMA_CUDA(ctx, ctx->zscratch.reserve(scratch_need));

    t_begin(ctx, SCEMA_T_RESAMPLE);
    for (int k = 0; k < NCLS; k++) {
        if (!cls_count[k]) continue;
        const uint64_t n_groups = (cls_count[k] + GROUP - 1) / GROUP;
        if (k < NCLS - 1 && !staged) {
            const uint32_t cap = std::min<uint32_t>(caps[k], ctx->max_len);
            k_resample_stream<<<(unsigned)(stream_warps[k] / RS_WARPS), 32 * RS_WARPS, RS_SMEM, ctx->stream>>>(
                ctx->d_steps, ctx->d_offsets.as<uint64_t>(), d_order, cls_first[k], cls_count[k],
                ctx->d_table_index.as<int64_t>(), ctx->d_tables.as<double>(), P, ctx->spline_own.as<double>(),
                ctx->zscratch.as<double>(), cap, 6, 0);
        } else if (k < NCLS - 1) {
            const size_t slab = smem_for(caps[k]);
            int per_sm = (int)(ctx->smem_optin / (slab + 1024));
            per_sm = per_sm > 32 ? 32 : (per_sm < 1 ? 1 : per_sm);
            uint64_t grid = std::min<uint64_t>((uint64_t)ctx->sm_count * per_sm, n_groups);
            SCEMA_CUDA(ctx, cudaFuncSetAttribute(k_resample_staged, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)smem_for(caps[NCLS - 2])));
            k_resample_staged<<<(unsigned)grid, 32, slab, ctx->stream>>>(
                ctx->d_steps, ctx->d_offsets.as<uint64_t>(), d_order, cls_first[k], cls_count[k],
                ctx->d_table_index.as<int64_t>(), ctx->d_tables.as<double>(), P, ctx->spline_own.as<double>(), caps[k]);
        } else {
            uint64_t grid = std::min<uint64_t>((uint64_t)ctx->sm_count * 32, n_groups);
            k_resample_global<<<(unsigned)grid, 32, 0, ctx->stream>>>(
                ctx->d_steps, ctx->d_offsets.as<uint64_t>(), d_order, cls_first[k], cls_count[k],
                ctx->d_table_index.as<int64_t>(), ctx->d_tables.as<double>(), P, ctx->spline_own.as<double>(),
                ctx->zscratch.as<double>());
        }
        ctx->launches++;
    }
    t_end(ctx, SCEMA_T_RESAMPLE);
    SCEMA_CUDA(ctx, cudaGetLastError());
    return SCEMA_OK;
}

// ------------------------------------------------------------------------------------------------
// Device-resident incremental history store: the in-process caller appends one
// strain sample per quadrature point per timestep (FEProblem::update_strain_quadrature_point_history,
// reference headers/FE_problem.h:1091-1103 -> Strain6D::add_current_strain, strain2spline.h:75-86)
// and re-fits every spline each step (spline_building, FE_problem.h:1167-1191). The store keeps the
// histories on the device in time-major order [step][n][6]: an append is one contiguous 48*n-byte
// copy, all histories have the same length (one factor table, no sorting), and in K1 the 30 chain
// lanes of a warp read 240 contiguous bytes per step.
// ------------------------------------------------------------------------------------------------
int store_reset(scema_ctx *ctx, uint64_t n, const uint32_t *ids, uint32_t capacity_steps)
{
    if (n >= (1ull << 32)) return fail(ctx, SCEMA_ERR_INVALID, "store_reset: more than 2^32-1 histories");
    ctx->store_n = n;
    ctx->store_steps = 0;
    ctx->store_cap = std::max<uint32_t>(capacity_steps, 8);
    ctx->store_ids.resize(n);
    for (uint64_t i = 0; i < n; i++) ctx->store_ids[i] = ids ? ids[i] : (uint32_t)i;
    SCEMA_CUDA(ctx, ctx->d_store.reserve(std::max<uint64_t>(n, 1) * 6 * sizeof(double) * ctx->store_cap));
    ctx->have_store = true;
    return SCEMA_OK;
}

int store_append(scema_ctx *ctx, const double *strain, int on_device)
{
    if (!ctx->have_store) return fail(ctx, SCEMA_ERR_STATE, "store_append: no store (call scema_store_reset)");
    if (ctx->store_n && !strain) return fail(ctx, SCEMA_ERR_INVALID, "store_append: null pointer");
    const size_t row = (size_t)ctx->store_n * 6 * sizeof(double);
    if (ctx->store_steps == ctx->store_cap) {  // grow: the filled prefix is contiguous
        scema::DevBuf nb;
        const uint32_t cap = ctx->store_cap * 2;
        SCEMA_CUDA(ctx, nb.reserve(std::max<size_t>(row, 48) * cap));
        SCEMA_CUDA(ctx, cudaMemcpyAsync(nb.p, ctx->d_store.p, row * ctx->store_steps, cudaMemcpyDeviceToDevice, ctx->stream));
        SCEMA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->d_store.release();
        ctx->d_store = nb;
        ctx->store_cap = cap;
    }
    if (row)
        SCEMA_CUDA(ctx, cudaMemcpyAsync(static_cast<char *>(ctx->d_store.p) + row * ctx->store_steps, strain, row,
                                        on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
    if (!on_device) SCEMA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the caller's buffer may be pageable / reused
    ctx->store_steps++;
    return SCEMA_OK;
}

int store_resample(scema_ctx *ctx, uint32_t P)
{
    if (!ctx->have_store) return fail(ctx, SCEMA_ERR_STATE, "store_resample: no store (call scema_store_reset)");
    if (P == 0) return fail(ctx, SCEMA_ERR_INVALID, "resample: spline_points must be >= 1");
    const uint32_t L = ctx->store_steps;
    if (ctx->store_n && L == 0)
        return fail(ctx, SCEMA_ERR_INVALID,
                    "Nothing to splinify! No strain data has been read in yet. Please use .from_file() or .add_current_strain() first.");
    if (ctx->store_n && L < 3)
        return fail(ctx, SCEMA_ERR_INVALID, "Not enough strain steps added. Need at least 3 points for splinify().");
    const uint64_t n = ctx->store_n;
    const uint32_t K = 6 * P;
    SCEMA_CUDA(ctx, ctx->spline_own.reserve((size_t)(n ? n : 1) * K * sizeof(double)));
    ctx->d_spline = ctx->spline_own.as<double>();
    ctx->n = n;
    ctx->ids = ctx->store_ids;
    ctx->K = K;
    ctx->spline_version++;
    ctx->have_spline = true;
    ctx->have_edges = false;
    ctx->ev_used[SCEMA_T_RESAMPLE] = false;
    if (n == 0) return SCEMA_OK;
    std::vector<uint8_t> present((size_t)L + 1, 0);
    present[L] = 1;
    int rc = ensure_tables_present(ctx, P, present, L);
    if (rc) return rc;
    static const char *wps_env = getenv("SCEMA_K1_WPS");
    const int wps = wps_env && atoi(wps_env) > 0 ? atoi(wps_env) : 24;
    const uint64_t n_groups = (n + GROUP - 1) / GROUP;
    uint64_t w = std::min<uint64_t>((uint64_t)ctx->sm_count * wps, n_groups);
    w = std::min<uint64_t>(w, std::max<uint64_t>(RS_WARPS, (1ull << 30) / ((uint64_t)L * 256)));
    w = (w + RS_WARPS - 1) / RS_WARPS * RS_WARPS;
    SCEMA_CUDA(ctx, ctx->zscratch.reserve(w * L * 256));
    t_begin(ctx, SCEMA_T_RESAMPLE);
    k_resample_stream<<<(unsigned)(w / RS_WARPS), 32 * RS_WARPS, RS_SMEM, ctx->stream>>>(
        ctx->d_store.as<double>(), nullptr, nullptr, 0, n, ctx->d_table_index.as<int64_t>(), ctx->d_tables.as<double>(), P,
        ctx->spline_own.as<double>(), ctx->zscratch.as<double>(), L, n * 6, L);
    ctx->launches++;
    t_end(ctx, SCEMA_T_RESAMPLE);
    SCEMA_CUDA(ctx, cudaGetLastError());
    return SCEMA_OK;
}

// Restrict the current spline matrix to a subset of its rows (the flagged quadrature points that
// spline_comparison collects, FE_problem.h:1202-1224): gather into a compact matrix that becomes the
// current one, IDs carried along.
__global__ void k_gather_rows(const double *__restrict__ src, const uint32_t *__restrict__ rows, uint64_t m, uint32_t K,
                              double *__restrict__ dst)
{
    const uint64_t total = m * K;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = e / K, k = e - r * K;
        dst[e] = src[(uint64_t)rows[r] * K + k];
    }
}

int select_rows(scema_ctx *ctx, const uint32_t *rows, uint64_t m)
{
    if (!ctx->have_spline) return fail(ctx, SCEMA_ERR_STATE, "Spline is not up to date.");
    if (m && !rows) return fail(ctx, SCEMA_ERR_INVALID, "select_rows: null pointer");
    for (uint64_t i = 0; i < m; i++)
        if (rows[i] >= ctx->n) return fail(ctx, SCEMA_ERR_INVALID, "select_rows: row index out of range");
    const uint32_t K = ctx->K;
    SCEMA_CUDA(ctx, ctx->d_select.reserve(std::max<uint64_t>(m, 1) * sizeof(uint32_t)));
    SCEMA_CUDA(ctx, ctx->spline_sel.reserve(std::max<uint64_t>(m * K, 1) * sizeof(double)));
    if (m) {
        SCEMA_CUDA(ctx, cudaMemcpyAsync(ctx->d_select.p, rows, m * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        if (K) {
            const unsigned grid = (unsigned)std::min<uint64_t>((m * K + 255) / 256, (uint64_t)ctx->sm_count * 16);
            k_gather_rows<<<grid, 256, 0, ctx->stream>>>(ctx->d_spline, ctx->d_select.as<uint32_t>(), m, K,
                                                         ctx->spline_sel.as<double>());
            ctx->launches++;
            SCEMA_CUDA(ctx, cudaGetLastError());
        }
        SCEMA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `rows` is the caller's buffer
    }
    std::vector<uint32_t> ids(m);
    for (uint64_t i = 0; i < m; i++) ids[i] = ctx->ids[rows[i]];
    ctx->ids.swap(ids);
    ctx->d_spline = ctx->spline_sel.as<double>();
    ctx->n = m;
    ctx->spline_version++;
    ctx->have_edges = false;
    return SCEMA_OK;
}

}  // namespace scema
