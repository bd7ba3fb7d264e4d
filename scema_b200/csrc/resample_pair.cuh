// K1, second generation: k_resample_pair — TWO chains per lane, small loops.
//
// Same arithmetic, operation for operation, as k_resample_stream (resample.cu) and therefore as
// Strain6D::splinify (reference headers/strain2spline.h:140-180) over tk::spline::set_points
// (headers/spline.h:284-373), band_matrix::l_solve / r_solve (:222-250) and spline::operator() (:375-396).
// What changes is the mapping (profiles/r02_ncu_resample_c3.txt: k_resample_stream issues ~100 warp
// instructions per step of one chain of which 21 are FP64, its two 8-step loops are 7 and 15 KB of
// code — 1.3 stall cycles per issued instruction are "no instruction" — and a warp has ONE dependent
// FP64 recurrence in flight per lane):
//  * a warp takes a PAIR of groups of five same-length histories; lane = (slot, component) runs the
//    chain of its slot in group A and in group B. The two recurrences are independent, so every
//    fixed-latency dependency has a second instruction to hide behind, and everything that depends
//    only on the step — factor-table entries, the cp.async bookkeeping, ring and scratch addresses,
//    the loop itself, the sample schedule — is issued once for two chains;
//  * the sweeps are rolled loops of two steps (the ping-pong of the table registers) without
//    per-step guards: the forward loop covers the steps that compute a next slope, the last row
//    that does not is peeled; ~4 KB of hot code in total;
//  * the range guard of the exact division costs four integer instructions: one running maximum of
//    a shifted exponent per chain, zero numerators exempt, the sign of a zero quotient restored by
//    one LOP3 (see div_acc).
//
// What the measurement showed (profiles/r02_k1_experiments.txt, r02_k1_scan.txt): 1.7x fewer executed
// instructions than k_resample_stream and, at the same number of chains in flight, the same time —
// K1 is bound by the round trip of z (as large as the input) through L2 / HBM and by memory latency,
// not by issue. The launch is therefore sized for the chains whose z rows L2 can hold (12 warps per SM
// for the ragged batch) and told per length class how y should reach L2 (`flags` below).
//
// This header is also compiled for the HOST by tests/helpers/k1_emul.cpp (K1_EMULATE): the kernel
// source below runs there thread by thread, with cp.async modelled in its two extreme legal timings
// (every copy lands at once / only when a wait_group forces it), against the CPU oracle — the
// indexing of the rings, the peeled steps and the pairing are checked without a GPU
// (tests/test_k1_emul.py).
#pragma once
#include <cstdint>

#ifndef K1_EMULATE
#define K1_HD __host__ __device__
#define K1_DEV __device__ __forceinline__
#define K1_GLOBAL __global__
#define K1_RESTRICT __restrict__
#define K1_SHARED_DECL(arr, chunkvar) extern __shared__ __align__(16) double arr[]; __shared__ uint32_t chunkvar;
#endif

namespace scema {

// table for one L, every array padded to an even length Lp (so each starts 16-byte aligned and the
// five arrays the sweeps need are one contiguous block for a bulk copy):
//   x[Lp] | hd[Lp] sd[Lp] lo[Lp] up[Lp] di[Lp] | ht[Pp] idx[Pp] | FW[Lp][4] | BW[Lp][4]
// FW[i] = {1/hd_i, hd_i, sd_i, lo_i} and BW[i] = {up_i, di_i, 1/di_i, 0} are what one forward / one
// backward step of the streamed kernels needs, packed so that each step is two 16-byte loads; the
// correctly rounded reciprocals feed the exact division (div_tab / div_fast / div_acc).
K1_HD inline uint32_t pad2(uint32_t v) { return v + (v & 1u); }
K1_HD inline uint64_t table_fw_offset(uint32_t L, uint32_t P) { return 6ull * pad2(L) + 2ull * pad2(P); }
K1_HD inline uint64_t table_doubles(uint32_t L, uint32_t P) { return table_fw_offset(L, P) + 8ull * pad2(L); }
K1_HD inline size_t rs_table_doubles(uint32_t L, uint32_t P) { return 8ull * pad2(L) + 2ull * pad2(P); }  // ht | ix | FW | BW

constexpr int GROUP = 5;          // histories per group (5 * 6 = 30 chain lanes)
constexpr int RS_WARPS = 4;       // warps per CTA
constexpr int CHUNK_GROUPS = 16;  // groups (of one length) handed to a CTA at a time
constexpr uint32_t SMEM_TAB_MAX_L = 256;  // longer histories read the factor table from global memory

struct K1Chunk {
    uint32_t first_group, n_groups, L, pad;
};

K1_GLOBAL void k_build_tables(const uint32_t *K1_RESTRICT lens, const uint64_t *K1_RESTRICT offs, uint32_t n_tables, uint32_t P,
                              double *K1_RESTRICT tables)
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

// One chain (history, component) with IEEE divisions throughout — the arithmetic of k_resample_global — for the chains
// whose numerators left the range of the branch-free division. z, then b, live in the private scratch column zs[i * zstride].
#ifndef K1_EMULATE
__device__ __noinline__
#endif
void resample_chain_slow(const double *K1_RESTRICT y, uint64_t ys, int L, const double *K1_RESTRICT tab, uint32_t P,
                         double *K1_RESTRICT zs, uint32_t zstride, double *K1_RESTRICT orow)
{
    const uint32_t Lp = pad2((uint32_t)L);
    const double *hd = tab + Lp, *sd = hd + Lp, *lo = sd + Lp, *up = lo + Lp, *di = up + Lp;
    const double *ht = tab + 6ull * Lp, *ix = ht + pad2(P);
    const double third = 1.0 / 3.0;
    double y1 = __ldg(y), y2 = __ldg(y + ys);
    double s_prev = __ddiv_rn(__dsub_rn(y2, y1), __ldg(hd));
    double z_prev = __dsub_rn(__dmul_rn(0.0, __ldg(sd)), 0.0);
    zs[0] = z_prev;
    y1 = y2;
    for (int i = 1; i < L - 1; i++) {
        y2 = __ldg(y + (size_t)(i + 1) * ys);
        const double s_cur = __ddiv_rn(__dsub_rn(y2, y1), __ldg(hd + i));
        const double r = __dmul_rn(__dsub_rn(s_cur, s_prev), __ldg(sd + i));
        const double sum = __dadd_rn(0.0, __dmul_rn(__ldg(lo + i), z_prev));
        z_prev = __dsub_rn(r, sum);
        zs[(size_t)i * zstride] = z_prev;
        s_prev = s_cur;
        y1 = y2;
    }
    {
        const double r = __dmul_rn(0.0, __ldg(sd + L - 1));
        const double sum = __dadd_rn(0.0, __dmul_rn(__ldg(lo + L - 1), z_prev));
        z_prev = __dsub_rn(r, sum);
    }
    double b_next = __ddiv_rn(__dsub_rn(z_prev, 0.0), __ldg(di + L - 1));
    zs[(size_t)(L - 1) * zstride] = b_next;
    for (int i = L - 2; i >= 0; i--) {
        const double sum = __dadd_rn(0.0, __dmul_rn(__ldg(up + i), b_next));
        b_next = __ddiv_rn(__dsub_rn(zs[(size_t)i * zstride], sum), __ldg(di + i));
        zs[(size_t)i * zstride] = b_next;
    }
    for (uint32_t p = 0; p < P; p++) {
        const int idx = (int)__ldg(ix + p);
        const double hstep = __ldg(ht + p), hdv = __ldg(hd + idx);
        const double ya = __ldg(y + (size_t)idx * ys), yb = __ldg(y + (size_t)(idx + 1) * ys);
        const double b0 = zs[(size_t)idx * zstride], b1 = zs[(size_t)(idx + 1) * zstride];
        const double a_i = __ddiv_rn(__dmul_rn(third, __dsub_rn(b1, b0)), hdv);
        const double c_i = __dsub_rn(__ddiv_rn(__dsub_rn(yb, ya), hdv), __dmul_rn(__dmul_rn(third, __dadd_rn(__dmul_rn(2.0, b0), b1)), hdv));
        double v = __dadd_rn(__dmul_rn(a_i, hstep), b0);
        v = __dadd_rn(__dmul_rn(v, hstep), c_i);
        v = __dadd_rn(__dmul_rn(v, hstep), ya);
        orow[(size_t)p * 6] = v;
    }
}

// ---- primitives of the pair kernel (the emulation provides its own under K1_EMULATE)
#ifndef PR_DEPTH_SLOTS
#define PR_DEPTH_SLOTS 8
#endif
constexpr int PR_DEPTH = PR_DEPTH_SLOTS;  // ring slots per chain (a power of two): y (then z) of step m sits in slot m % PR_DEPTH
constexpr int PR_ROW = 64;         // doubles per ring slot and per scratch row of a warp: [chain A: 32 lanes | chain B: 32 lanes]
constexpr uint32_t PR_ACC_LIMIT = 0xE0000000u;  // div_acc: every numerator in range <=> running maximum below this
constexpr size_t PR_RING_BYTES = (size_t)RS_WARPS * PR_DEPTH * PR_ROW * sizeof(double);

#ifndef K1_EMULATE
// 8-byte asynchronous copy global -> shared at smem address `dst` (+ a constant), one commit group per step.
template <int OFF>
K1_DEV void pr_cp_async8(uint32_t dst, const double *src)
{
    asm volatile("cp.async.ca.shared.global [%0+%1], [%2], 8;" ::"r"(dst), "n"(OFF), "l"(src) : "memory");
}
// the same with an L2 eviction policy (pr_policy) for the source line
template <int OFF>
K1_DEV void pr_cp_async8_hint(uint32_t dst, const double *src, uint64_t policy)
{
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0+%1], [%2], 8, %3;" ::"r"(dst), "n"(OFF), "l"(src), "l"(policy) : "memory");
}
// L2 eviction policy of the y stream: evict_first (read once: it should not displace the z rows that come back) or normal
K1_DEV uint64_t pr_policy(bool evict_first)
{
    uint64_t pf, pn;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pf));
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pn));
    return evict_first ? pf : pn;
}
// L2 eviction policy of the z stores: evict_last (written once, read back once, soon: keep it in L2 over the y stream) or normal
K1_DEV uint64_t pr_policy_z(bool evict_last)
{
    uint64_t pl, pn;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pl));
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pn));
    return evict_last ? pl : pn;
}
K1_DEV void pr_store_z(double *p, double v, uint64_t policy)
{
    asm volatile("st.global.cg.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(policy) : "memory");
}
K1_DEV void pr_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
K1_DEV void pr_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
template <int OFF>
K1_DEV double pr_ring_read(uint32_t a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a), "n"(OFF) : "memory");
    return v;
}
K1_DEV uint32_t pr_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
K1_DEV void pr_prefetch_l2(const double *p, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// Cursor over {double2} table entries: a 32-bit shared-memory address (STAB: the table of the length being worked on
// sits in shared memory) or a global pointer; ld<E>() reads entry E relative to the cursor into the register pair it is
// consumed from (destination tied: left to the compiler, each prefetch went to a temporary that was copied at the end
// of the same step — a move that waits for the load).
template <bool STAB>
struct TabCursor;
template <>
struct TabCursor<true> {
    uint32_t a;
    K1_DEV void set(const double2 *p) { a = pr_smem_addr(p); }
    template <int E>
    K1_DEV void ld(double2 &d) const
    {
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+%3];" : "=d"(d.x), "=d"(d.y) : "r"(a), "n"(E * 16));
    }
    K1_DEV void advance(int entries) { a += entries * 16; }
};
template <>
struct TabCursor<false> {
    const double2 *p;
    K1_DEV void set(const double2 *q) { p = q; }
    template <int E>
    K1_DEV void ld(double2 &d) const
    {
        asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2+%3];" : "=d"(d.x), "=d"(d.y) : "l"(p), "n"(E * 16));
    }
    K1_DEV void advance(int entries) { p += entries; }
};
// magnitude of q with the sign of s: one LOP3 on the high word, (q.hi & ~0x80000000) | (s.hi & 0x80000000)
K1_DEV double pr_copysign(double q, double s)
{
    int qh = __double2hiint(q);
    asm("lop3.b32 %0, %0, 0x80000000, %1, 0xB8;" : "+r"(qh) : "r"(__double2hiint(s)));
    return __hiloint2double(qh, __double2loint(q));
}
template <bool STAB>
K1_DEV double pr_tab_f64(const double *p) { return STAB ? *p : __ldg(p); }
template <bool STAB>
K1_DEV double2 pr_tab_f64x2(const double2 *p) { return STAB ? *p : __ldg(p); }
#endif

// a / b for a table divisor b with rb = RN(1/b), WITHOUT a branch: q = a*rb plus two FMA corrections (after the first, q is a
// faithful rounding of a/b; Markstein's theorem then makes q + (a - b*q)*rb round to RN(a/b)) — proven for numerators whose
// exponent field lies in [128, 1920) (every intermediate stays normal) and checked against the true quotient for 4e8 pairs by
// tests/test_fastdiv.py. The guard only ACCUMULATES: acc is the running maximum of (exponent field - 128) taken as an
// unsigned number with the high mantissa bits below it, so "every numerator so far was in range" <=> acc < PR_ACC_LIMIT
// (subnormal and tiny numerators wrap to >= 0xF0000000, huge ones, inf and NaN are >= 0xE0000000). A numerator that is
// exactly +-0 is exempt: all five operations are then exact zeros and only the SIGN can differ from a/b (fma(+0, rb, -0) = +0);
// the first product q0 = a*rb always has the sign of a/b, for zero as for every in-range numerator, so it is copied onto the
// result. A chain whose acc ends >= PR_ACC_LIMIT is redone by resample_chain_slow with IEEE divisions.
K1_DEV double div_acc(double a, double b, double rb, uint32_t &acc)
{
    const uint32_t hi = (uint32_t)__double2hiint(a), lo = (uint32_t)__double2loint(a);
    const double q0 = __dmul_rn(a, rb);
    double r = __fma_rn(-b, q0, a);
    double q = __fma_rn(r, rb, q0);
    r = __fma_rn(-b, q, a);
    q = __fma_rn(r, rb, q);
    const uint32_t t = (hi << 1) - 0x10000000u;
    if (((hi & 0x7fffffffu) | lo) != 0u) acc = acc > t ? acc : t;
    return pr_copysign(q, q0);
}

// One FULL forward step i (spline.h:306, :228-233) of both chains: z_i from s_i, s_{i-1}, z_{i-1}, then the next slope
// s_{i+1} = (y_{i+2} - y_{i+1}) / hd_{i+1}. On entry E1 = {sd_i, lo_i}, E0 = {1/hd_{i+1}, hd_{i+1}}; the same two entries of
// step i+1 are loaded into N1/N0. y_{i+2} comes from ring slot (i+2) % PR_DEPTH (its copy is the oldest of the PR_DEPTH-1
// groups in flight); the slot read one step earlier is refilled with y_{i+1+PR_DEPTH}.
#define PR_FWD_STEP(E1, E0, N1, N0, PF)                                                               \
    {                                                                                                 \
        fwc.template ld<3>(N1);                                                                       \
        fwc.template ld<4>(N0);                                                                       \
        pr_cp_async_wait<PR_DEPTH - 2>();                                                             \
        const uint32_t rd = ring + (uint32_t)((i + 2) & (PR_DEPTH - 1)) * (PR_ROW * 8);               \
        const double ynA = pr_ring_read<0>(rd), ynB = pr_ring_read<256>(rd);                          \
        if (i + 1 + PR_DEPTH < L) {                                                                   \
            const uint32_t wr = ring + (uint32_t)((i + 1) & (PR_DEPTH - 1)) * (PR_ROW * 8);           \
            pr_cp_async8_hint<0>(wr, ypA, ypol);                                                      \
            pr_cp_async8_hint<256>(wr, ypB, ypol);                                                    \
        }                                                                                             \
        pr_cp_async_commit();                                                                         \
        ypA += ys;                                                                                    \
        ypB += ys;                                                                                    \
        if (PF && pf_windows && (i & (PR_PF_WINDOW - 1)) == 0 && i + PR_PF_WINDOW < L) {              \
            /* steps [i+W, i+2W) of both histories -> L2, by the lane of component 0 */               \
            const int w0 = i + PR_PF_WINDOW, wn = L - w0 < PR_PF_WINDOW ? L - w0 : PR_PF_WINDOW;      \
            pr_prefetch_l2(yA + (size_t)w0 * 6, 48u * (uint32_t)wn);                                  \
            if (two) pr_prefetch_l2(yB + (size_t)w0 * 6, 48u * (uint32_t)wn);                         \
        }                                                                                             \
        {                                                                                             \
            const double r = __dmul_rn(__dsub_rn(sA_cur, sA_prev), E1.x);                             \
            const double sum = __dadd_rn(0.0, __dmul_rn(E1.y, zA));                                   \
            zA = __dsub_rn(r, sum);                                                                   \
            pr_store_z(zp, zA, zpol);                                                                 \
            sA_prev = sA_cur;                                                                         \
            sA_cur = div_acc(__dsub_rn(ynA, yA_hi), E0.y, E0.x, accA);                                \
            yA_hi = ynA;                                                                              \
        }                                                                                             \
        {                                                                                             \
            const double r = __dmul_rn(__dsub_rn(sB_cur, sB_prev), E1.x);                             \
            const double sum = __dadd_rn(0.0, __dmul_rn(E1.y, zB));                                   \
            zB = __dsub_rn(r, sum);                                                                   \
            pr_store_z(zp + 32, zB, zpol);                                                            \
            sB_prev = sB_cur;                                                                         \
            sB_cur = div_acc(__dsub_rn(ynB, yB_hi), E0.y, E0.x, accB);                                \
            yB_hi = ynB;                                                                              \
        }                                                                                             \
        zp += PR_ROW;                                                                                 \
        fwc.advance(2);                                                                               \
        i++;                                                                                          \
    }

// One backward step i (spline.h:243-248) of both chains plus, when a sample lives in interval i, the samples of that
// interval (spline.h:345-349, :393). On entry bX_next = b_{i+1}, W0 = {up_i, di_i}, W1 = {1/di_i, -}; the entries of step
// i-1 are loaded into V0/V1 (for i = 0 these are the last entries of FW, inside this table's allocation, never consumed).
// z_i comes from ring slot i % PR_DEPTH; the slot read one step earlier is refilled with z_{i+1-PR_DEPTH}.
#define PR_BWD_STEP(W0, W1, V0, V1)                                                                   \
    {                                                                                                 \
        bwc.template ld<-2>(V0);                                                                      \
        bwc.template ld<-1>(V1);                                                                      \
        pr_cp_async_wait<PR_DEPTH - 2>();                                                             \
        const uint32_t rd = ring + (uint32_t)(i & (PR_DEPTH - 1)) * (PR_ROW * 8);                     \
        const double ziA = pr_ring_read<0>(rd), ziB = pr_ring_read<256>(rd);                          \
        if (i + 1 - PR_DEPTH >= 0) {                                                                  \
            const uint32_t wr = ring + (uint32_t)((i + 1) & (PR_DEPTH - 1)) * (PR_ROW * 8);           \
            pr_cp_async8<0>(wr, zq);                                                                  \
            pr_cp_async8<256>(wr, zq + 32);                                                           \
        }                                                                                             \
        pr_cp_async_commit();                                                                         \
        zq -= PR_ROW;                                                                                 \
        const double bA = div_acc(__dsub_rn(ziA, __dadd_rn(0.0, __dmul_rn(W0.x, bA_next))), W0.y, W1.x, accA); \
        const double bB = div_acc(__dsub_rn(ziB, __dadd_rn(0.0, __dmul_rn(W0.x, bB_next))), W0.y, W1.x, accB); \
        if (i == nxt) {                                                                               \
            const double2 f0 = pr_tab_f64x2<STAB>(FW + 2 * i);                                        \
            const double hdv = f0.y;                                                                  \
            const double aA = div_acc(__dmul_rn(third, __dsub_rn(bA_next, bA)), hdv, f0.x, accA);     \
            const double cA = __dsub_rn(div_acc(__dsub_rn(yA_b, yA_a), hdv, f0.x, accA),              \
                                        __dmul_rn(__dmul_rn(third, __dadd_rn(__dmul_rn(2.0, bA), bA_next)), hdv)); \
            const double aB = div_acc(__dmul_rn(third, __dsub_rn(bB_next, bB)), hdv, f0.x, accB);     \
            const double cB = __dsub_rn(div_acc(__dsub_rn(yB_b, yB_a), hdv, f0.x, accB),              \
                                        __dmul_rn(__dmul_rn(third, __dadd_rn(__dmul_rn(2.0, bB), bB_next)), hdv)); \
            do {                                                                                      \
                const double hstep = pr_tab_f64<STAB>(ht + p);                                        \
                double vA = __dadd_rn(__dmul_rn(aA, hstep), bA);                                      \
                double vB = __dadd_rn(__dmul_rn(aB, hstep), bB);                                      \
                vA = __dadd_rn(__dmul_rn(vA, hstep), cA);                                             \
                vB = __dadd_rn(__dmul_rn(vB, hstep), cB);                                             \
                vA = __dadd_rn(__dmul_rn(vA, hstep), yA_a);                                           \
                vB = __dadd_rn(__dmul_rn(vB, hstep), yB_a);                                           \
                orowA[(size_t)p * 6] = vA;                                                            \
                orowB[(size_t)p * 6] = vB;                                                            \
                p--;                                                                                  \
                nxt = p >= 0 ? (int)pr_tab_f64<STAB>(ix + p) : -1;                                    \
            } while (nxt == i);                                                                       \
            if (nxt >= 0) {                                                                           \
                yA_a = __ldg(yA + (size_t)nxt * ys);                                                  \
                yA_b = __ldg(yA + (size_t)(nxt + 1) * ys);                                            \
                yB_a = __ldg(yB + (size_t)nxt * ys);                                                  \
                yB_b = __ldg(yB + (size_t)(nxt + 1) * ys);                                            \
            }                                                                                         \
        }                                                                                             \
        bA_next = bA;                                                                                 \
        bB_next = bB;                                                                                 \
        bwc.advance(-2);                                                                              \
        i--;                                                                                          \
    }

#ifndef PR_MIN_CTAS
#define PR_MIN_CTAS 5
#endif

// `flags` of k_resample_pair (memory-system behaviour only; the arithmetic never depends on them):
//   bits 0-1  how y reaches L2 in the ragged layout: 0 = one bulk prefetch of the whole history when its chains start,
//             1 = nothing ahead of the ring copies, 2 = windows of PR_PF_WINDOW steps, one window ahead of the ring
//   bit 2     the ring copies of y carry the L2 evict_first policy
//   bit 3     the z stores carry the L2 evict_last policy
// (discard.global.L2 of the z rows after their read-back was tried in round 1 — DRAM writes fell, the time did not — and is
// not here: one lane would drop a line that holds sixteen lanes' values, which needs the warp to be in lockstep.)
constexpr uint32_t PR_PF_MASK = 3u, PR_PF_WHOLE = 0u, PR_PF_NONE = 1u, PR_PF_WINDOWS = 2u, PR_Y_EVICT_FIRST = 4u, PR_Z_EVICT_LAST = 8u;
constexpr int PR_PF_WINDOW = 16;  // steps per prefetch window (768 bytes of one history)

// ys = distance (in doubles) between consecutive steps of one history: 6 for the ragged batch ([L][6] blocks, history h
// starts at offsets[h]; `order` lists the histories group by group, five slots per group, 0xffffffff = empty slot), n*6
// for the time-major history store ([step][n][6], history h starts at h, every history uniform_L steps long, order ==
// nullptr, groups in index order). chunks == nullptr: the chunk list is implicit (CHUNK_GROUPS consecutive groups of
// length uniform_L). zscratch: gridDim.x * RS_WARPS * cap rows of PR_ROW doubles.
template <bool STAB>
K1_GLOBAL void
#ifndef K1_EMULATE
__launch_bounds__(32 * RS_WARPS, PR_MIN_CTAS)
#endif
k_resample_pair(const double *K1_RESTRICT steps, const uint64_t *K1_RESTRICT offsets, const uint32_t *K1_RESTRICT order, uint64_t n_hist,
                const K1Chunk *K1_RESTRICT chunks, uint32_t n_chunks, unsigned int *K1_RESTRICT chunk_counter,
                const int64_t *K1_RESTRICT table_index, const double *K1_RESTRICT tables, uint32_t P, double *K1_RESTRICT out,
                double *K1_RESTRICT zscratch, uint32_t cap, uint64_t ys, uint32_t uniform_L, uint32_t flags)
{
    K1_SHARED_DECL(rs_smem, s_chunk)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // this lane's ring: slot s of chain A at ring + s * 512 bytes, of chain B 256 bytes further; y on the way up, then (all y
    // copies have landed by then) z on the way down; PR_DEPTH * 512 bytes per warp
    const uint32_t ring = pr_smem_addr(rs_smem + (size_t)warp * (PR_DEPTH * PR_ROW) + lane);
    double *stab = rs_smem + (size_t)RS_WARPS * PR_DEPTH * PR_ROW;  // [ht | ix | FW | BW] of the current length
    double *K1_RESTRICT zs = zscratch + ((uint64_t)blockIdx.x * RS_WARPS + warp) * cap * PR_ROW + lane;
    const uint32_t K = 6 * P, Pp = pad2(P);
    const double third = 1.0 / 3.0;
    const int hh = lane / 6, c = lane - hh * 6;
    const uint64_t ypol = pr_policy((flags & PR_Y_EVICT_FIRST) != 0), zpol = pr_policy_z((flags & PR_Z_EVICT_LAST) != 0);
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

        // a warp takes the groups of the chunk two at a time: chain A = group gA, chain B = group gA + 1
        for (uint32_t gA = ch.first_group + 2 * warp; gA < ch.first_group + ch.n_groups; gA += 2 * RS_WARPS) {
            uint64_t hA = ~0ull, hB = ~0ull;
            if (lane < GROUP * 6) {
                const bool haveB = gA + 1 < ch.first_group + ch.n_groups;
                if (order) {
                    const uint32_t oA = order[(uint64_t)gA * GROUP + hh];
                    if (oA != 0xffffffffu) hA = oA;
                    if (haveB) { const uint32_t oB = order[((uint64_t)gA + 1) * GROUP + hh]; if (oB != 0xffffffffu) hB = oB; }
                } else {
                    const uint64_t oA = (uint64_t)gA * GROUP + hh, oB = oA + GROUP;
                    if (oA < n_hist) hA = oA;
                    if (haveB && oB < n_hist) hB = oB;
                }
            }
            if (hA == ~0ull && hB == ~0ull) continue;  // idle lane; no warp-level primitive below
            // an empty slot on one side runs the other side's chain a second time (same bits to the same addresses)
            const bool twoA = hA != ~0ull, twoB = hB != ~0ull;
            if (!twoA) hA = hB;
            if (!twoB) hB = hA;
            const uint64_t offA = uniform_L ? hA : offsets[hA], offB = uniform_L ? hB : offsets[hB];
            const double *yA = steps + offA * 6 + c, *yB = steps + offB * 6 + c;

            // ragged layout: the history (or its first two windows) -> L2 now, one bulk prefetch per history by the lane
            // of component 0; the ring copies then hit L2
            const bool two = twoA && twoB;
            const bool pf_windows = c == 0 && !uniform_L && (flags & PR_PF_MASK) == PR_PF_WINDOWS;
            if (c == 0 && !uniform_L && (flags & PR_PF_MASK) != PR_PF_NONE) {
                const uint32_t nb = 48u * (uint32_t)(pf_windows && L > 2 * PR_PF_WINDOW ? 2 * PR_PF_WINDOW : L);
                pr_prefetch_l2(steps + offA * 6, nb);
                if (two) pr_prefetch_l2(steps + offB * 6, nb);
            }

            // ---- forward substitution fused with the right-hand side. Software pipeline: the slope s_{i+1} and the
            // table entries of step i+1 are produced while the z recurrences of step i run.
#pragma unroll
            for (int m = 3; m <= PR_DEPTH + 1; m++) {  // y_3 .. y_{PR_DEPTH+1}: PR_DEPTH-1 groups
                if (m < L) {
                    const uint32_t wr = ring + (uint32_t)(m & (PR_DEPTH - 1)) * (PR_ROW * 8);
                    pr_cp_async8_hint<0>(wr, yA + (size_t)m * ys, ypol);
                    pr_cp_async8_hint<256>(wr, yB + (size_t)m * ys, ypol);
                }
                pr_cp_async_commit();
            }
            double sA_prev, sA_cur, zA, yA_hi, sB_prev, sB_cur, zB, yB_hi;
            uint32_t accA = 0, accB = 0;  // div_acc: running maxima; < PR_ACC_LIMIT <=> every numerator of the chain was in range
            double2 a1, a0, b1, b0;       // ping-pong: {sd, lo} of the coming step and {1/hd, hd} of the one after
            TabCursor<STAB> fwc;
            fwc.set(FW + 2);              // cursor = FW + 2 i, i = 1
            {
                const double2 f00 = pr_tab_f64x2<STAB>(FW), f01 = pr_tab_f64x2<STAB>(FW + 1), f10 = pr_tab_f64x2<STAB>(FW + 2);
                fwc.template ld<1>(a1);  // {sd_1, lo_1}
                fwc.template ld<2>(a0);  // {1/hd_2, hd_2}
                const double yA0 = __ldg(yA), yA1 = __ldg(yA + ys), yA2 = __ldg(yA + 2 * ys);  // L >= 3
                const double yB0 = __ldg(yB), yB1 = __ldg(yB + ys), yB2 = __ldg(yB + 2 * ys);
                sA_prev = div_acc(__dsub_rn(yA1, yA0), f00.y, f00.x, accA);  // s_0
                sB_prev = div_acc(__dsub_rn(yB1, yB0), f00.y, f00.x, accB);
                sA_cur = div_acc(__dsub_rn(yA2, yA1), f10.y, f10.x, accA);   // s_1
                sB_cur = div_acc(__dsub_rn(yB2, yB1), f10.y, f10.x, accB);
                zA = zB = __dsub_rn(__dmul_rn(0.0, f01.x), 0.0);             // row 0: rhs = 0, empty sum
                pr_store_z(zs, zA, zpol);
                pr_store_z(zs + 32, zB, zpol);
                yA_hi = yA2;
                yB_hi = yB2;
            }
            int i = 1;
            {
                const double *ypA = yA + (size_t)(2 + PR_DEPTH) * ys, *ypB = yB + (size_t)(2 + PR_DEPTH) * ys;  // y_{i+1+PR_DEPTH}
                double *zp = zs + PR_ROW;                                                                       // row i
                // full steps i = 1 .. L-3 (each also produces s_{i+1}), two per trip
                while (i + 1 <= L - 3) {  // i odd in the first step, even in the second: windows start on even steps
                    PR_FWD_STEP(a1, a0, b1, b0, false)
                    PR_FWD_STEP(b1, b0, a1, a0, true)
                }
                if (i <= L - 3) PR_FWD_STEP(a1, a0, b1, b0, false)
                // i == L-2: the last interior row has no next slope
                {
                    const double2 e = pr_tab_f64x2<STAB>(FW + 2 * i + 1);  // {sd_{L-2}, lo_{L-2}}
                    zA = __dsub_rn(__dmul_rn(__dsub_rn(sA_cur, sA_prev), e.x), __dadd_rn(0.0, __dmul_rn(e.y, zA)));
                    zB = __dsub_rn(__dmul_rn(__dsub_rn(sB_cur, sB_prev), e.x), __dadd_rn(0.0, __dmul_rn(e.y, zB)));
                    pr_store_z(zp, zA, zpol);
                    pr_store_z(zp + 32, zB, zpol);
                }
            }
            {
                const double2 f1 = pr_tab_f64x2<STAB>(FW + 2 * (L - 1) + 1);  // {sd, lo} of row L-1
                const double r = __dmul_rn(0.0, f1.x);                        // rhs = 0
                zA = __dsub_rn(r, __dadd_rn(0.0, __dmul_rn(f1.y, zA)));
                zB = __dsub_rn(r, __dadd_rn(0.0, __dmul_rn(f1.y, zB)));
            }

            // ---- back substitution with the samples evaluated on the way: sample p lives in interval ix[p],
            // non-increasing as p falls, so b is never stored
            pr_cp_async_wait<0>();  // the ring changes hands: no y copy may still be in flight
#pragma unroll
            for (int u = 0; u < PR_DEPTH - 1; u++) {  // z_{L-2} .. z_{L-PR_DEPTH}: PR_DEPTH-1 groups
                const int m = L - 2 - u;
                if (m >= 0) {
                    const uint32_t wr = ring + (uint32_t)(m & (PR_DEPTH - 1)) * (PR_ROW * 8);
                    pr_cp_async8<0>(wr, zs + (size_t)m * PR_ROW);
                    pr_cp_async8<256>(wr, zs + (size_t)m * PR_ROW + 32);
                }
                pr_cp_async_commit();
            }
            double bA_next, bB_next;
            {
                const double2 w0 = pr_tab_f64x2<STAB>(BW + 2 * (L - 1)), w1 = pr_tab_f64x2<STAB>(BW + 2 * (L - 1) + 1);
                bA_next = div_acc(__dsub_rn(zA, 0.0), w0.y, w1.x, accA);
                bB_next = div_acc(__dsub_rn(zB, 0.0), w0.y, w1.x, accB);
            }
            int p = (int)P - 1;
            int nxt = (int)pr_tab_f64<STAB>(ix + p);
            double yA_a = __ldg(yA + (size_t)nxt * ys), yA_b = __ldg(yA + (size_t)(nxt + 1) * ys);
            double yB_a = __ldg(yB + (size_t)nxt * ys), yB_b = __ldg(yB + (size_t)(nxt + 1) * ys);
            double *orowA = out + hA * K + c, *orowB = out + hB * K + c;
            TabCursor<STAB> bwc;
            i = L - 2;
            bwc.set(BW + 2 * i);     // cursor = BW + 2 i
            bwc.template ld<0>(a0);  // {up, di} of step L-2
            bwc.template ld<1>(a1);  // {1/di, -}
            {
                const double *zq = zs + (size_t)(i + 1 - PR_DEPTH) * PR_ROW;  // row i+1-PR_DEPTH (only dereferenced when >= 0)
                while (i >= 1) {
                    PR_BWD_STEP(a0, a1, b0, b1)
                    PR_BWD_STEP(b0, b1, a0, a1)
                }
                if (i == 0) PR_BWD_STEP(a0, a1, b0, b1)
            }
            pr_cp_async_wait<0>();  // nothing of this pair may land in the ring after the next pair starts
            // rare: subnormal / huge / non-finite numerators
            if (accA >= PR_ACC_LIMIT) resample_chain_slow(yA, ys, L, tab, P, zs, PR_ROW, orowA);
            if (accB >= PR_ACC_LIMIT && two) resample_chain_slow(yB, ys, L, tab, P, zs + 32, PR_ROW, orowB);
        }
    }
}
#undef PR_FWD_STEP
#undef PR_BWD_STEP

}  // namespace scema
