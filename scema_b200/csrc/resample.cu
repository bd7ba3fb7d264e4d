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
//  * k_resample_stream: one warp per group of five same-length histories, lane = (history,
//    component) chain; y and z stream through cp.async rings, the factor table of the length being
//    worked on sits in shared memory (one copy per CTA), the samples are evaluated inside the
//    backward sweep. Work is handed out in chunks of 16 groups of one length, longest first.
//  * k_resample_global: fallback for histories longer than 131072 steps.
#include "common.cuh"
#include <algorithm>
#include "resample_pair.cuh"

// Defaults of the K1 tuning (k1_tuning below), from the grid of tools/k1_probe.py --scan on one B200
// (profiles/r02_k1_scan.txt): the two-chain kernel; 12 resident warps per SM for the ragged batch (more chains in flight
// push the z rows out of L2 before they are read back: 20 warps cost 8-12 %), 20 for the time-major store; whole-history
// L2 prefetch for histories of at most 64 steps, windowed prefetch above; evict_last on the store's z rows.
#ifndef K1_DEFAULT_KERNEL
#define K1_DEFAULT_KERNEL 1
#endif
#ifndef K1_PAIR_WPS_RAGGED
#define K1_PAIR_WPS_RAGGED 12
#endif
#ifndef K1_PAIR_WPS_STORE
#define K1_PAIR_WPS_STORE 20
#endif
constexpr uint32_t K1_FLAGS_AUTO = 0xffffffffu;  // per launch: by layout and length class (k1_flags_for)

namespace scema {


// ---- streamed kernel. One warp per group of five histories of the SAME length, lane = (history,
// component) chain; 30 of 32 lanes busy. Nothing is staged in bulk, so the number of resident chains
// is not capped by a slab: y is read straight from global memory and z goes to a warp-private
// scratch column block zs[i][32] (coalesced 256-byte rows, re-read by the same warp ~L steps later,
// mostly out of L2); the samples are evaluated inside the backward sweep at the moment b_idx and
// b_idx+1 exist, so b is never stored.
//
// div_tab(a, b, rb) == __ddiv_rn(a, b) for a table divisor b with rb = RN(1/b): two
// Newton-style FMA corrections of q = a*rb. After the first, q is a faithful rounding of a/b;
// Markstein's theorem then makes q + (a - b*q)*rb round to RN(a/b). Guarded to numerators whose
// exponent keeps every intermediate normal; zeros, subnormals, huge values, inf and NaN take the
// IEEE division. 5 dependent FP64 ops (~40 cycles) instead of ~110 for the division sequence.
// Checked against the true quotient for 4e8 (a,b) pairs by tests/test_fastdiv.py.
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

// The same quotient WITHOUT a branch, for the streamed kernel: the guard only ACCUMULATES (`ok` goes false when a numerator
// is outside the range the two corrections are proven for: subnormal, huge, inf, NaN), and a chain whose flag is false at
// the end is redone by resample_chain_slow with IEEE divisions. ncu (profiles/r02_ncu_resample_c3.txt): with a fast / slow
// branch at each of the 36 division sites, control flow (BRA, BSSY, BSYNC, ISETP) was 20 % of the executed instructions —
// as much as the FP64 arithmetic — and the inlined IEEE sequences doubled the size of the hot loops. A zero numerator is
// exact in the first product (a * rb = +-0 with the sign of a / b) and stays in range.
__device__ __forceinline__ double div_fast(double a, double b, double rb, bool &ok)
{
    const uint32_t e = ((uint32_t)__double2hiint(a) >> 20) & 0x7ffu;
    const double q0 = __dmul_rn(a, rb);
    double r = __fma_rn(-b, q0, a);
    double q = __fma_rn(r, rb, q0);
    r = __fma_rn(-b, q, a);
    q = __fma_rn(r, rb, q);
    const bool zero = a == 0.0;
    ok = ok && (zero || e - 128u < 1792u);
    return zero ? q0 : q;
}

constexpr int RING = 8;        // steps per unrolled block
constexpr int DEPTH = 16;      // slots of the per-lane shared-memory prefetch ring (y, then z)
constexpr size_t RS_RING_BYTES = (size_t)RS_WARPS * DEPTH * 32 * sizeof(double);


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

// Table entries of the coming step, loaded into the ping-pong variable they are consumed from
// (inline PTX with the destination tied to that variable; left to the compiler, each prefetch went to
// a temporary that was copied at the end of the SAME step — a move that waits for the load).
// STAB: the table of the length being worked on sits in shared memory.
template <bool STAB>
__device__ __forceinline__ void ld_tab(double2 &d, const double2 *p)
{
    if (STAB)
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(d.x), "=d"(d.y) : "r"((uint32_t)__cvta_generic_to_shared(p)));
    else
        asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(d.x), "=d"(d.y) : "l"(p));
}
template <bool STAB>
__device__ __forceinline__ double tab_f64(const double *p) { return STAB ? *p : __ldg(p); }
template <bool STAB>
__device__ __forceinline__ double2 tab_f64x2(const double2 *p) { return STAB ? *p : __ldg(p); }

// One forward step i = i0 + U (spline.h:306, :228-233). On entry: s_cur = s_i, s_prev = s_{i-1},
// z_prev = z_{i-1}, y_hi = y_{i+1}, E1 = {sd_i, lo_i}, E0 = {1/hd_{i+1}, hd_{i+1}}. Loads the same
// two entries for step i+1 into N1/N0, takes y_{i+2} from ring slot (i+2)%DEPTH and starts the copy
// of y_{i+1+DEPTH} into the slot read one step earlier.
#define K1_FWD_STEP(U, E1, E0, N1, N0)                                                              \
    if (i0 + (U) < L - 1) {                                                                         \
        const int i = i0 + (U);                                                                     \
        ld_tab<STAB>(N1, FWb + 2 * ((U) + 1) + 1);                                                  \
        ld_tab<STAB>(N0, FWb + 2 * ((U) + 2));                                                      \
        cp_async_wait<DEPTH - 2>();                                                                 \
        const double y_nx = ry[((i + 2) & (DEPTH - 1)) * 32];                                       \
        if (i + 1 + DEPTH < L) cp_async8(ry + ((i + 1) & (DEPTH - 1)) * 32, yb + (size_t)((U) + 1 + DEPTH) * ys); \
        cp_async_commit();                                                                          \
        const double r = __dmul_rn(__dsub_rn(s_cur, s_prev), E1.x);                                 \
        const double sum = __dadd_rn(0.0, __dmul_rn(E1.y, z_prev));                                 \
        z_prev = __dsub_rn(r, sum);                                                                 \
        __stcg(zb + (size_t)(U) * 32, z_prev);                                                      \
        s_prev = s_cur;                                                                             \
        if (i + 1 < L - 1) s_cur = div_fast(__dsub_rn(y_nx, y_hi), E0.y, E0.x, ok);                 \
        y_hi = y_nx;                                                                                \
    }

// One backward step i = i0 - U (spline.h:243-248) plus the samples of interval i (spline.h:345-349,
// :393). On entry: b_next = b_{i+1}, W0 = {up_i, di_i}, W1 = {1/di_i, -}. Loads the entries of step
// i-1 into V0/V1, takes z_i from ring slot i%DEPTH and starts the copy of z_{i+1-DEPTH} into the
// slot read one step earlier.
#define K1_BWD_STEP(U, W0, W1, V0, V1)                                                              \
    if (i0 - (U) >= 0) {                                                                            \
        const int i = i0 - (U);                                                                     \
        ld_tab<STAB>(V0, BWb - 2 * ((U) + 1));                                                      \
        ld_tab<STAB>(V1, BWb - 2 * ((U) + 1) + 1);                                                  \
        cp_async_wait<DEPTH - 2>();                                                                 \
        const double zi = rz[(i & (DEPTH - 1)) * 32];                                               \
        if (i + 1 - DEPTH >= 0) cp_async8(rz + ((i + 1) & (DEPTH - 1)) * 32, zb - (size_t)((U) - 1 + DEPTH) * 32); \
        cp_async_commit();                                                                          \
        const double sum = __dadd_rn(0.0, __dmul_rn(W0.x, b_next));                                 \
        const double b_i = div_fast(__dsub_rn(zi, sum), W0.y, W1.x, ok);                            \
        if (i == nxt) {                                                                             \
            const double2 f0 = tab_f64x2<STAB>(FW + 2 * i);                                         \
            const double hdv = f0.y;                                                                \
            const double a_i = div_fast(__dmul_rn(third, __dsub_rn(b_next, b_i)), hdv, f0.x, ok);   \
            const double c_i =                                                                      \
                __dsub_rn(div_fast(__dsub_rn(y_b, y_a), hdv, f0.x, ok),                             \
                          __dmul_rn(__dmul_rn(third, __dadd_rn(__dmul_rn(2.0, b_i), b_next)), hdv)); \
            do {                                                                                    \
                const double hstep = tab_f64<STAB>(ht + p);                                         \
                double v = __dadd_rn(__dmul_rn(a_i, hstep), b_i);                                   \
                v = __dadd_rn(__dmul_rn(v, hstep), c_i);                                            \
                v = __dadd_rn(__dmul_rn(v, hstep), y_a);                                            \
                orow[(size_t)p * 6] = v;                                                            \
                p--;                                                                                \
                nxt = p >= 0 ? (int)tab_f64<STAB>(ix + p) : -1;                                     \
            } while (nxt == i);                                                                     \
            if (nxt >= 0) { y_a = __ldg(y + (size_t)nxt * ys); y_b = __ldg(y + (size_t)(nxt + 1) * ys); } \
        }                                                                                           \
        b_next = b_i;                                                                               \
    }

// ys = distance (in doubles) between consecutive steps of one history: 6 for the ragged batch
// ([L][6] blocks, history h starts at offsets[h]; `order` lists the histories group by group, five
// slots per group, 0xffffffff = empty slot), n*6 for the time-major history store ([step][n][6],
// history h starts at h, every history uniform_L steps long, order == nullptr, groups in index order).
// chunks == nullptr: the chunk list is implicit (CHUNK_GROUPS consecutive groups of length uniform_L).
template <bool STAB>
__global__ void __launch_bounds__(32 * RS_WARPS, 6) k_resample_stream(const double *__restrict__ steps,
                                                                  const uint64_t *__restrict__ offsets,
                                                                  const uint32_t *__restrict__ order, uint64_t n_hist,
                                                                  const K1Chunk *__restrict__ chunks, uint32_t n_chunks,
                                                                  unsigned int *__restrict__ chunk_counter,
                                                                  const int64_t *__restrict__ table_index,
                                                                  const double *__restrict__ tables, uint32_t P,
                                                                  double *__restrict__ out, double *__restrict__ zscratch,
                                                                  uint32_t cap, uint64_t ys, uint32_t uniform_L)
{
    extern __shared__ __align__(16) double rs_smem[];
    __shared__ uint32_t s_chunk;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // this lane's ring, ry[slot * 32]: y on the way up, then (all y copies have landed by then) z on
    // the way down; 4 KB per warp
    double *ry = rs_smem + (size_t)warp * (DEPTH * 32) + lane;
    double *rz = ry;
    double *stab = rs_smem + (size_t)RS_WARPS * DEPTH * 32;  // [ht | ix | FW | BW] of the current length
    double *__restrict__ zs = zscratch + ((uint64_t)blockIdx.x * RS_WARPS + warp) * cap * 32 + lane;
    const uint32_t K = 6 * P, Pp = pad2(P);
    const double third = 1.0 / 3.0;
    const int hh = lane / 6, c = lane - hh * 6;
    int loaded_L = -1;

    while (true) {
        __syncthreads();  // every warp is done with the previous chunk and its table
        if (threadIdx.x == 0) s_chunk = atomicAdd(chunk_counter, 1u);
        __syncthreads();
        const uint32_t ci = s_chunk;
        if (ci >= n_chunks) break;
        K1Chunk ch;
        if (chunks) {
            ch = chunks[ci];
        } else {
            const uint64_t n_groups = (n_hist + GROUP - 1) / GROUP;
            ch.first_group = ci * CHUNK_GROUPS;
            ch.n_groups = (uint32_t)(n_groups - ch.first_group < (uint64_t)CHUNK_GROUPS ? n_groups - ch.first_group : (uint64_t)CHUNK_GROUPS);
            ch.L = uniform_L;
        }
        const int L = (int)ch.L;
        const uint32_t Lp = pad2((uint32_t)L);
        const double *tab = tables + table_index[L];
        if (STAB && L != loaded_L) {
            const double2 *src = reinterpret_cast<const double2 *>(tab + 6ull * Lp);
            double2 *dst = reinterpret_cast<double2 *>(stab);
            const uint32_t n2 = (uint32_t)(rs_table_doubles((uint32_t)L, P) / 2);
            for (uint32_t q = threadIdx.x; q < n2; q += 32 * RS_WARPS) dst[q] = __ldg(src + q);
            loaded_L = L;
            __syncthreads();
        }
        const double *ht = STAB ? stab : tab + 6ull * Lp, *ix = ht + Pp;
        const double2 *FW = reinterpret_cast<const double2 *>(ht + 2ull * Pp);
        const double2 *BW = FW + 2ull * Lp;

        for (uint32_t g = ch.first_group + warp; g < ch.first_group + ch.n_groups; g += RS_WARPS) {
            uint64_t h = ~0ull;
            if (lane < GROUP * 6) {
                if (order) { const uint32_t o = order[(uint64_t)g * GROUP + hh]; if (o != 0xffffffffu) h = o; }
                else { const uint64_t o = (uint64_t)g * GROUP + hh; if (o < n_hist) h = o; }
            }
            if (h == ~0ull) continue;  // idle lane; no warp-level primitive below
            const uint64_t off = uniform_L ? h : offsets[h];
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
            bool ok = true;          // every numerator of this chain stayed in the range of the branch-free division
            double2 a1, a0, b1, b0;  // ping-pong: {sd, lo} of the coming step and {1/hd, hd} of the one after
            {
                const double y_0 = __ldg(y), y_1 = __ldg(y + ys), y_2 = __ldg(y + 2 * ys);  // L >= 3
                const double2 f00 = tab_f64x2<STAB>(FW), f01 = tab_f64x2<STAB>(FW + 1), f10 = tab_f64x2<STAB>(FW + 2);
                ld_tab<STAB>(a1, FW + 3);  // {sd_1, lo_1}
                ld_tab<STAB>(a0, FW + 4);  // {1/hd_2, hd_2}
                s_prev = div_fast(__dsub_rn(y_1, y_0), f00.y, f00.x, ok);  // s_0
                s_cur = div_fast(__dsub_rn(y_2, y_1), f10.y, f10.x, ok);   // s_1
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
                const double2 f1 = tab_f64x2<STAB>(FW + 2 * (L - 1) + 1);  // {sd, lo} of row L-1
                const double r = __dmul_rn(0.0, f1.x);                     // rhs = 0
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
                const double2 w0 = tab_f64x2<STAB>(BW + 2 * (L - 1)), w1 = tab_f64x2<STAB>(BW + 2 * (L - 1) + 1);
                b_next = div_fast(__dsub_rn(z_prev, 0.0), w0.y, w1.x, ok);
            }
            int p = (int)P - 1;
            int nxt = (int)tab_f64<STAB>(ix + p);
            double y_a = __ldg(y + (size_t)nxt * ys), y_b = __ldg(y + (size_t)(nxt + 1) * ys);
            double *orow = out + h * K + c;
            ld_tab<STAB>(a0, BW + 2 * (L - 2));      // {up, di} of step L-2
            ld_tab<STAB>(a1, BW + 2 * (L - 2) + 1);  // {1/di, -}
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
            cp_async_wait<0>();  // nothing of this group may land in the ring after the next group starts
            if (!ok) resample_chain_slow(y, ys, L, tab, P, zs, 32, out + h * K + c);  // rare: subnormal / huge / non-finite numerators
        }
    }
}
#undef K1_FWD_STEP
#undef K1_BWD_STEP

// ---- fallback for histories longer than the last length class: the sweeps read y from global
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
        bool chain = lane < GROUP * 6 && (uint32_t)hh < n_here;
        if (chain && order && order[q0 + hh] == 0xffffffffu) chain = false;  // empty slot of a padded group
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
            if (order && order[q0 + eh] == 0xffffffffu) continue;
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
    // nothing to do when these histories already have their tables for this P (the scan below is O(n) on the host,
    // with the GPU idle: a re-fit of unchanged histories should not pay it)
    if (ctx->tables_for_version == ctx->histories_version && ctx->table_P == P && ctx->tables_used) return SCEMA_OK;
    // distinct lengths of the current batch
    std::vector<uint8_t> present((size_t)ctx->max_len + 1, 0);
    for (uint64_t i = 0; i < ctx->hn; i++) present[ctx->h_offsets[i + 1] - ctx->h_offsets[i]] = 1;
    const int rc = ensure_tables_present(ctx, P, present, ctx->max_len);
    if (!rc) ctx->tables_for_version = ctx->histories_version;
    return rc;
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

// Work plan of the ragged batch (rebuilt when the histories change): histories sorted by length,
// cut into groups of five of the SAME length (the last group of a length may have empty slots),
// groups cut into chunks of CHUNK_GROUPS, chunks ordered longest first inside each length class so
// that the dynamic hand-out ends with the short ones. One launch per class sizes the z scratch.
static const uint32_t K1_CAPS[] = {64, SMEM_TAB_MAX_L, 2048, 16384, 131072};
constexpr int K1_NCLS = sizeof(K1_CAPS) / sizeof(K1_CAPS[0]) + 1;  // last class: global-scratch fallback

// The plan covers one or more contiguous RANGES of histories; every range has its own groups, chunks
// and per-class launch set, so a range can be resampled as soon as its raw steps are on the device
// (the host-buffer pipeline of scema_cluster) while the default path is the single range [0, n).
static int build_plan(scema_ctx *ctx, const std::vector<uint64_t> &bounds)
{
    if (ctx->order_version == ctx->histories_version && ctx->plan_bounds == bounds) return SCEMA_OK;
    const size_t n_ranges = bounds.size() - 1;
    auto cls_of = [&](uint32_t L) { int k = 0; while (k < K1_NCLS - 1 && L > K1_CAPS[k]) k++; return k; };
    // scratch kept between calls (per thread; a context is driven by one thread at a time): a fresh 4 MB vector per
    // call costs more in page faults than the plan costs to compute
    static thread_local std::vector<uint32_t> order;
    static thread_local std::vector<K1Chunk> chunks;
    static thread_local std::vector<uint64_t> len_count, group_first, fill;
    order.clear();
    chunks.clear();
    ctx->plan_chunk_begin.assign(n_ranges * (K1_NCLS + 1), 0);
    ctx->plan_groups.assign(n_ranges * K1_NCLS, 0);
    ctx->plan_first_slot.assign(n_ranges * K1_NCLS, 0);
    ctx->plan_max_len.assign(n_ranges, 0);
    ctx->plan_steps.assign(n_ranges, 0);
    for (size_t r = 0; r < n_ranges; r++) {
        const uint64_t h0 = bounds[r], h1 = bounds[r + 1];
        uint32_t max_len = 0;
        for (uint64_t i = h0; i < h1; i++) max_len = std::max<uint32_t>(max_len, (uint32_t)(ctx->h_offsets[i + 1] - ctx->h_offsets[i]));
        ctx->plan_max_len[r] = max_len;
        ctx->plan_steps[r] = ctx->h_offsets[h1] - ctx->h_offsets[h0];
        len_count.assign((size_t)max_len + 2, 0);
        for (uint64_t i = h0; i < h1; i++) len_count[ctx->h_offsets[i + 1] - ctx->h_offsets[i]]++;
        // group slots per length, in ascending length; slot numbers are global over all ranges
        const uint64_t base = order.size() / GROUP;
        group_first.assign((size_t)max_len + 2, base);
        for (uint32_t L = 0; L <= max_len; L++) group_first[L + 1] = group_first[L] + (len_count[L] + GROUP - 1) / GROUP;
        const uint64_t n_groups = group_first[max_len + 1] - base;
        order.resize((base + n_groups) * GROUP, 0xffffffffu);
        fill.assign((size_t)max_len + 1, 0);
        for (uint64_t i = h0; i < h1; i++) {
            const uint64_t L = ctx->h_offsets[i + 1] - ctx->h_offsets[i];
            order[group_first[L] * GROUP + fill[L]++] = (uint32_t)i;
        }
        for (int k = 0; k < K1_NCLS; k++) {
            ctx->plan_chunk_begin[r * (K1_NCLS + 1) + k] = (uint32_t)chunks.size();
            bool first = true;
            for (uint32_t L = max_len; L >= 3; L--) {  // longest first
                if (cls_of(L) != k || !len_count[L]) continue;
                const uint64_t g0 = group_first[L], g1 = group_first[L + 1];
                uint64_t &fs = ctx->plan_first_slot[r * K1_NCLS + k];
                if (first) { fs = g0; first = false; }
                fs = std::min<uint64_t>(fs, g0);
                ctx->plan_groups[r * K1_NCLS + k] += g1 - g0;
                for (uint64_t g = g0; g < g1; g += CHUNK_GROUPS)
                    chunks.push_back(K1Chunk{(uint32_t)g, (uint32_t)std::min<uint64_t>(CHUNK_GROUPS, g1 - g), L, 0});
            }
        }
        ctx->plan_chunk_begin[r * (K1_NCLS + 1) + K1_NCLS] = (uint32_t)chunks.size();
    }
    if (order.size() >= (1ull << 32)) return fail(ctx, SCEMA_ERR_INVALID, "resample: too many histories");
    if (order.empty()) order.push_back(0xffffffffu);
    SCEMA_CUDA(ctx, ctx->d_order.reserve(order.size() * sizeof(uint32_t)));
    SCEMA_CUDA(ctx, ctx->d_chunks.reserve(std::max<size_t>(chunks.size(), 1) * sizeof(K1Chunk)));
    SCEMA_CUDA(ctx, cudaMemcpyAsync(ctx->d_order.p, order.data(), order.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    if (!chunks.empty())
        SCEMA_CUDA(ctx, cudaMemcpyAsync(ctx->d_chunks.p, chunks.data(), chunks.size() * sizeof(K1Chunk), cudaMemcpyHostToDevice,
                                        ctx->stream));
    SCEMA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // order / chunks are stack-lifetime host buffers
    ctx->order_version = ctx->histories_version;
    ctx->plan_bounds = bounds;
    return SCEMA_OK;
}

// Which streamed kernel runs and how: process-wide settings, read once from the environment and adjustable through
// scema_k1_tune (a measurement hook: tools/k1_probe.py scans them in one process).
//   kernel  1 = k_resample_pair (two chains per lane, resample_pair.cuh), 0 = the first-generation k_resample_stream;
//           SCEMA_K1_KERNEL=pair|stream
//   wps     resident warps per SM a launch is sized for, for the ragged batch and for the history store
//           (0 = the kernel's default); SCEMA_K1_WPS (both)
//   flags   memory-system behaviour of k_resample_pair (PR_* in resample_pair.cuh); SCEMA_K1_FLAGS; unset = per launch
//           (k1_flags_for)
// Both kernels take the same plan, tables and arguments; a warp's z scratch is cap rows of 32 (stream) or 64 (pair) doubles.
struct K1Tuning {
    int kernel, wps_ragged, wps_store;
    uint32_t flags;
};
static K1Tuning &k1_tuning()
{
    static K1Tuning t = [] {
        K1Tuning v{K1_DEFAULT_KERNEL, 0, 0, K1_FLAGS_AUTO};
        if (const char *e = getenv("SCEMA_K1_KERNEL")) v.kernel = !strcmp(e, "pair") ? 1 : !strcmp(e, "stream") ? 0 : v.kernel;
        if (const char *e = getenv("SCEMA_K1_WPS")) v.wps_ragged = v.wps_store = atoi(e) > 0 ? atoi(e) : 0;
        if (const char *e = getenv("SCEMA_K1_FLAGS")) v.flags = (uint32_t)strtoul(e, nullptr, 0);
        return v;
    }();
    return t;
}
int k1_tune(int kernel, int wps_ragged, int wps_store, int flags)
{
    K1Tuning &t = k1_tuning();
    if (kernel == 0 || kernel == 1) t.kernel = kernel;
    if (wps_ragged >= 0) t.wps_ragged = wps_ragged;
    if (wps_store >= 0) t.wps_store = wps_store;
    if (flags >= 0) t.flags = (uint32_t)flags;
    if (flags == -2) t.flags = K1_FLAGS_AUTO;
    return SCEMA_OK;
}
static bool k1_pair() { return k1_tuning().kernel == 1; }
static uint32_t k1_flags_for(uint32_t cap, bool store)
{
    const uint32_t f = k1_tuning().flags;
    if (f != K1_FLAGS_AUTO) return f;
    if (store) return PR_Z_EVICT_LAST;
    return cap <= 64 ? PR_PF_WHOLE : PR_PF_WINDOWS;
}
static size_t k1_row_bytes() { return k1_pair() ? PR_ROW * sizeof(double) : 32 * sizeof(double); }

// resident warps for a launch over n_groups groups of at most cap steps, and the z scratch they need
static uint64_t stream_warps_for(const scema_ctx *ctx, uint64_t n_groups, uint32_t cap, bool store)
{
    const K1Tuning &t = k1_tuning();
    int wps = store ? t.wps_store : t.wps_ragged;
    if (wps <= 0) wps = k1_pair() ? (store ? K1_PAIR_WPS_STORE : K1_PAIR_WPS_RAGGED) : 24;
    const uint64_t units = k1_pair() ? (n_groups + 1) / 2 : n_groups;  // what one warp takes at a time
    uint64_t w = std::min<uint64_t>((uint64_t)ctx->sm_count * wps, units);
    w = std::min<uint64_t>(w, std::max<uint64_t>(RS_WARPS, (1ull << 30) / ((uint64_t)cap * k1_row_bytes())));
    return (w + RS_WARPS - 1) / RS_WARPS * RS_WARPS;
}

static int launch_stream(scema_ctx *ctx, const double *steps, const uint64_t *offsets, const uint32_t *order, uint64_t n_hist,
                         const K1Chunk *chunks, uint32_t n_chunks, unsigned int *counter, uint32_t P, uint64_t warps, uint32_t cap,
                         uint64_t ys, uint32_t uniform_L)
{
    const bool stab = cap <= SMEM_TAB_MAX_L;
    const bool pair = k1_pair();
    const size_t smem = (pair ? PR_RING_BYTES : RS_RING_BYTES) + (stab ? rs_table_doubles(cap, P) * sizeof(double) : 0);
    if (smem > ctx->smem_optin) return fail(ctx, SCEMA_ERR_INVALID, "resample: spline_points too large for the shared-memory table");
    const unsigned grid = (unsigned)(warps / RS_WARPS);
    if (pair) {
        auto kern = stab ? k_resample_pair<true> : k_resample_pair<false>;
        SCEMA_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, 32 * RS_WARPS, smem, ctx->stream>>>(steps, offsets, order, n_hist, chunks, n_chunks, counter,
                                                        ctx->d_table_index.as<int64_t>(), ctx->d_tables.as<double>(), P,
                                                        ctx->spline_own.as<double>(), ctx->zscratch.as<double>(), cap, ys, uniform_L,
                                                        k1_flags_for(cap, uniform_L != 0));
    } else {
        auto kern = stab ? k_resample_stream<true> : k_resample_stream<false>;
        SCEMA_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, 32 * RS_WARPS, smem, ctx->stream>>>(steps, offsets, order, n_hist, chunks, n_chunks, counter,
                                                        ctx->d_table_index.as<int64_t>(), ctx->d_tables.as<double>(), P,
                                                        ctx->spline_own.as<double>(), ctx->zscratch.as<double>(), cap, ys, uniform_L);
    }
    ctx->launches++;
    return SCEMA_OK;
}

// Everything a resample needs before the first kernel: output matrix, factor tables, plan over the
// given ranges of histories (bounds[0] = 0 < ... < bounds.back() = n), scratch, zeroed chunk counters.
int resample_prepare(scema_ctx *ctx, uint32_t P, const std::vector<uint64_t> &bounds)
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
    ctx->ids_lazy = ctx->hist_ids_lazy;
    if (!ctx->hist_ids_lazy) ctx->ids = ctx->hist_ids;
    ctx->K = K;
    ctx->spline_version++;
    ctx->have_spline = true;
    ctx->have_edges = false;
    if (ctx->hn == 0) return SCEMA_OK;

    int rc = ensure_tables(ctx, P);
    if (rc) return rc;
    rc = build_plan(ctx, bounds);
    if (rc) return rc;
    const size_t n_ranges = bounds.size() - 1;
    uint64_t scratch_need = 0;
    for (size_t r = 0; r < n_ranges; r++)
        for (int k = 0; k < K1_NCLS; k++) {
            if (!ctx->plan_groups[r * K1_NCLS + k]) continue;
            if (k == K1_NCLS - 1) { scratch_need = std::max<uint64_t>(scratch_need, (uint64_t)ctx->total_steps * 6 * sizeof(double)); continue; }
            const uint32_t cap = std::min<uint32_t>(K1_CAPS[k], ctx->plan_max_len[r]);
            scratch_need = std::max<uint64_t>(scratch_need, stream_warps_for(ctx, ctx->plan_groups[r * K1_NCLS + k], cap, false) * cap * k1_row_bytes());
        }
    if (scratch_need) SCEMA_CUDA(ctx, ctx->zscratch.reserve(scratch_need));
    SCEMA_CUDA(ctx, ctx->d_chunk_counters.reserve(n_ranges * K1_NCLS * sizeof(unsigned int)));
    SCEMA_CUDA(ctx, cudaMemsetAsync(ctx->d_chunk_counters.p, 0, n_ranges * K1_NCLS * sizeof(unsigned int), ctx->stream));
    return SCEMA_OK;
}

// K1 launches of range r of the prepared plan.
int resample_launch_range(scema_ctx *ctx, uint32_t P, size_t r)
{
    int rc;
    for (int k = 0; k < K1_NCLS; k++) {
        const uint64_t n_groups = ctx->plan_groups[r * K1_NCLS + k];
        if (!n_groups) continue;
        if (k < K1_NCLS - 1) {
            const uint32_t cap = std::min<uint32_t>(K1_CAPS[k], ctx->plan_max_len[r]);
            const uint32_t cb = ctx->plan_chunk_begin[r * (K1_NCLS + 1) + k], ce = ctx->plan_chunk_begin[r * (K1_NCLS + 1) + k + 1];
            rc = launch_stream(ctx, ctx->d_steps, ctx->d_offsets.as<uint64_t>(), ctx->d_order.as<uint32_t>(), ctx->hn,
                               ctx->d_chunks.as<K1Chunk>() + cb, ce - cb, ctx->d_chunk_counters.as<unsigned int>() + r * K1_NCLS + k, P,
                               stream_warps_for(ctx, n_groups, cap, false), cap, 6, 0);
            if (rc) return rc;
        } else {
            // histories longer than the last class: sweeps through a global scratch of the input's shape
            uint64_t grid = std::min<uint64_t>((uint64_t)ctx->sm_count * 32, n_groups);
            k_resample_global<<<(unsigned)grid, 32, 0, ctx->stream>>>(
                ctx->d_steps, ctx->d_offsets.as<uint64_t>(), ctx->d_order.as<uint32_t>(), ctx->plan_first_slot[r * K1_NCLS + k] * GROUP,
                n_groups * GROUP, ctx->d_table_index.as<int64_t>(), ctx->d_tables.as<double>(), P,
                ctx->spline_own.as<double>(), ctx->zscratch.as<double>());
            ctx->launches++;
        }
    }
    SCEMA_CUDA(ctx, cudaGetLastError());
    return SCEMA_OK;
}

int resample_run(scema_ctx *ctx, uint32_t P)
{
    const std::vector<uint64_t> bounds = {0, ctx->hn};
    int rc = resample_prepare(ctx, P, bounds);
    if (rc || ctx->hn == 0) return rc;
    t_begin(ctx, SCEMA_T_RESAMPLE);
    rc = resample_launch_range(ctx, P, 0);
    if (rc) return rc;
    t_end(ctx, SCEMA_T_RESAMPLE);
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
    ctx->ids_lazy = false;
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
    if (L > K1_CAPS[K1_NCLS - 2]) return fail(ctx, SCEMA_ERR_INVALID, "store_resample: more than 131072 steps per history");
    const uint64_t n_groups = (n + GROUP - 1) / GROUP;
    const uint64_t w = stream_warps_for(ctx, n_groups, L, true);
    SCEMA_CUDA(ctx, ctx->zscratch.reserve(w * L * k1_row_bytes()));
    SCEMA_CUDA(ctx, ctx->d_chunk_counters.reserve(K1_NCLS * sizeof(unsigned int)));
    SCEMA_CUDA(ctx, cudaMemsetAsync(ctx->d_chunk_counters.p, 0, sizeof(unsigned int), ctx->stream));
    t_begin(ctx, SCEMA_T_RESAMPLE);
    rc = launch_stream(ctx, ctx->d_store.as<double>(), nullptr, nullptr, n, nullptr,
                       (uint32_t)((n_groups + CHUNK_GROUPS - 1) / CHUNK_GROUPS), ctx->d_chunk_counters.as<unsigned int>(), P, w, L,
                       n * 6, L);
    if (rc) return rc;
    ctx->launches--;  // counted below
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
    // never gather in place, never grow the buffer that is being read: a selection of a selection (rows of the
    // current matrix, duplicates allowed, so m may exceed n) goes to the other of two buffers
    scema::DevBuf &dst = ctx->spline_sel[ctx->d_spline == ctx->spline_sel[0].as<double>() ? 1 : 0];
    SCEMA_CUDA(ctx, dst.reserve(std::max<uint64_t>(m * K, 1) * sizeof(double)));
    if (m) {
        SCEMA_CUDA(ctx, cudaMemcpyAsync(ctx->d_select.p, rows, m * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        if (K) {
            const unsigned grid = (unsigned)std::min<uint64_t>((m * K + 255) / 256, (uint64_t)ctx->sm_count * 16);
            k_gather_rows<<<grid, 256, 0, ctx->stream>>>(ctx->d_spline, ctx->d_select.as<uint32_t>(), m, K, dst.as<double>());
            ctx->launches++;
            SCEMA_CUDA(ctx, cudaGetLastError());
        }
        SCEMA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `rows` is the caller's buffer
    }
    std::vector<uint32_t> ids(m);
    const std::vector<uint32_t> &cur = ids_of(ctx);
    for (uint64_t i = 0; i < m; i++) ids[i] = cur[rows[i]];
    ctx->ids.swap(ids);
    ctx->ids_lazy = false;
    ctx->d_spline = dst.as<double>();
    ctx->n = m;
    ctx->spline_version++;
    ctx->have_edges = false;
    return SCEMA_OK;
}

}  // namespace scema
