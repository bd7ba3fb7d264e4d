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
//  * k_resample: one warp per group of five histories; lane = (history-in-group, component)
//    runs the two sequential sweeps of its own chain (30 of 32 lanes busy), z/b live in a
//    per-warp shared-memory slab laid out [step][lane] (conflict-free); the whole warp then
//    evaluates the 5*6*P samples and stores them coalesced in the reference's p*6+c order.
#include "common.cuh"

namespace scema {

// table for one L: x[L] hd[L] sd[L] lo[L] up[L] di[L] ht[P] idx[P]
__host__ __device__ inline uint64_t table_doubles(uint32_t L, uint32_t P) { return 6ull * L + 2ull * P; }

__global__ void k_build_tables(const uint32_t *__restrict__ lens, const uint64_t *__restrict__ offs,
                               uint32_t n_tables, uint32_t P, double *__restrict__ tables)
{
    uint32_t ti = blockIdx.x * blockDim.x + threadIdx.x;
    if (ti >= n_tables) return;
    const uint32_t L = lens[ti];
    const int n = (int)L;
    double *x = tables + offs[ti], *hd = x + L, *sd = hd + L, *lo = sd + L, *up = lo + L, *di = up + L,
           *ht = di + L, *ix = ht + P;
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
}

constexpr int GROUP = 5;  // histories per warp (5*6 = 30 chain lanes)

template <bool SMEM_Z>
__global__ void __launch_bounds__(32) k_resample(const double *__restrict__ steps,
                                                 const uint64_t *__restrict__ offsets,
                                                 const int64_t *__restrict__ table_index,
                                                 const double *__restrict__ tables, uint32_t P, uint64_t n,
                                                 double *__restrict__ out, double *__restrict__ zglobal)
{
    extern __shared__ double zs[];  // [Lmax][32] when SMEM_Z
    const int lane = threadIdx.x;
    const uint32_t K = 6 * P;
    const double third = 1.0 / 3.0;
    const uint64_t n_groups = (n + GROUP - 1) / GROUP;

    for (uint64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int hh = lane / 6, c = lane - hh * 6;
        const uint64_t h = grp * GROUP + hh;
        const bool chain = lane < GROUP * 6 && h < n;
        uint64_t off = 0;
        int L = 0;
        const double *tab = tables;
        if (chain) {
            off = offsets[h];
            L = (int)(offsets[h + 1] - off);
            tab = tables + table_index[L];
        }
        const double *y = steps + off * 6 + c;
        const double *hd = tab + L, *sd = hd + L, *lo = sd + L, *up = lo + L, *di = up + L;
        double *zg = SMEM_Z ? nullptr : zglobal + off * 6 + c;

#define ZAT(i) (SMEM_Z ? zs[(i) * 32 + lane] : zg[(size_t)(i) * 6])

        if (chain) {
            // ---- forward substitution fused with the right-hand side (spline.h:306, :228-233)
            double y1 = __ldg(y), y2 = __ldg(y + 6);
            double s_prev = __ddiv_rn(__dsub_rn(y2, y1), __ldg(hd));
            double z_prev = __dsub_rn(__dmul_rn(0.0, __ldg(sd)), 0.0);  // row 0: rhs = 0, empty sum
            ZAT(0) = z_prev;
            y1 = y2;
#pragma unroll 4
            for (int i = 1; i < L - 1; i++) {
                y2 = __ldg(y + (size_t)(i + 1) * 6);
                double s_cur = __ddiv_rn(__dsub_rn(y2, y1), __ldg(hd + i));
                double r = __dmul_rn(__dsub_rn(s_cur, s_prev), __ldg(sd + i));
                double sum = __dadd_rn(0.0, __dmul_rn(__ldg(lo + i), z_prev));
                z_prev = __dsub_rn(r, sum);
                ZAT(i) = z_prev;
                s_prev = s_cur;
                y1 = y2;
            }
            {
                double r = __dmul_rn(0.0, __ldg(sd + L - 1));  // row L-1: rhs = 0
                double sum = __dadd_rn(0.0, __dmul_rn(__ldg(lo + L - 1), z_prev));
                z_prev = __dsub_rn(r, sum);
            }
            // ---- back substitution (spline.h:243-248); b overwrites z
            double b_next = __ddiv_rn(__dsub_rn(z_prev, 0.0), __ldg(di + L - 1));
            ZAT(L - 1) = b_next;
#pragma unroll 4
            for (int i = L - 2; i >= 0; i--) {
                double sum = __dadd_rn(0.0, __dmul_rn(__ldg(up + i), b_next));
                b_next = __ddiv_rn(__dsub_rn(ZAT(i), sum), __ldg(di + i));
                ZAT(i) = b_next;
            }
        }
        if (!SMEM_Z) __threadfence_block();
        __syncwarp();

        // ---- evaluation at the P sample points, all 32 lanes (spline.h:345-349, :393)
        const uint64_t h0 = grp * GROUP;
        const uint32_t n_here = (uint32_t)((n - h0) < (uint64_t)GROUP ? (n - h0) : (uint64_t)GROUP);
        for (uint32_t o = lane; o < n_here * K; o += 32) {
            const uint32_t eh = o / K, k = o - eh * K, p = k / 6, ec = k - p * 6;
            const uint64_t eoff = offsets[h0 + eh];
            const int eL = (int)(offsets[h0 + eh + 1] - eoff);
            const double *etab = tables + table_index[eL];
            const double *ehd = etab + eL, *eht = etab + 6 * (size_t)eL, *eix = eht + P;
            const int idx = (int)__ldg(eix + p);
            const double hstep = __ldg(eht + p), hdv = __ldg(ehd + idx);
            const double *ey = steps + eoff * 6 + ec;
            const double ya = __ldg(ey + (size_t)idx * 6), yb = __ldg(ey + (size_t)(idx + 1) * 6);
            const int cl = eh * 6 + ec;
            double b0, b1;
            if (SMEM_Z) { b0 = zs[idx * 32 + cl]; b1 = zs[(idx + 1) * 32 + cl]; }
            else { const double *g = zglobal + eoff * 6 + ec; b0 = g[(size_t)idx * 6]; b1 = g[(size_t)(idx + 1) * 6]; }
            const double a_i = __ddiv_rn(__dmul_rn(third, __dsub_rn(b1, b0)), hdv);
            const double c_i = __dsub_rn(__ddiv_rn(__dsub_rn(yb, ya), hdv),
                                         __dmul_rn(__dmul_rn(third, __dadd_rn(__dmul_rn(2.0, b0), b1)), hdv));
            double v = __dadd_rn(__dmul_rn(a_i, hstep), b0);
            v = __dadd_rn(__dmul_rn(v, hstep), c_i);
            v = __dadd_rn(__dmul_rn(v, hstep), ya);
            out[h0 * K + o] = v;
        }
        __syncwarp();
#undef ZAT
    }
}

static int ensure_tables(scema_ctx *ctx, uint32_t P)
{
    // distinct lengths of the current batch
    std::vector<uint8_t> present((size_t)ctx->max_len + 1, 0);
    for (uint64_t i = 0; i < ctx->n; i++) present[ctx->h_offsets[i + 1] - ctx->h_offsets[i]] = 1;
    if (ctx->table_P != P) { ctx->table_off.clear(); ctx->tables_used = 0; ctx->table_P = P; }
    std::vector<uint32_t> new_lens;
    std::vector<uint64_t> new_offs;
    uint64_t used = ctx->tables_used;
    for (uint32_t L = 3; L <= ctx->max_len; L++) {
        if (!present[L] || ctx->table_off.count(L)) continue;
        new_lens.push_back(L);
        new_offs.push_back(used);
        used += table_doubles(L, P);
    }
    bool index_stale = ctx->table_index_len < ctx->max_len + 1;
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
    std::vector<int64_t> index((size_t)ctx->max_len + 1, -1);
    for (auto &kv : ctx->table_off)
        if (kv.first <= ctx->max_len) index[kv.first] = (int64_t)kv.second;
    SCEMA_CUDA(ctx, ctx->d_table_index.reserve(index.size() * sizeof(int64_t)));
    SCEMA_CUDA(ctx, cudaMemcpyAsync(ctx->d_table_index.p, index.data(), index.size() * sizeof(int64_t),
                                    cudaMemcpyHostToDevice, ctx->stream));
    SCEMA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->table_index_len = ctx->max_len + 1;
    return SCEMA_OK;
}

int resample_run(scema_ctx *ctx, uint32_t P)
{
    if (!ctx->have_histories) return fail(ctx, SCEMA_ERR_STATE, "resample: no histories set");
    if (P == 0) return fail(ctx, SCEMA_ERR_INVALID, "resample: spline_points must be >= 1");
    if (ctx->n && ctx->min_len < 3)
        return fail(ctx, SCEMA_ERR_INVALID,
                    "Not enough strain steps added. Need at least 3 points for splinify().");
    const uint32_t K = 6 * P;
    SCEMA_CUDA(ctx, ctx->spline_own.reserve((size_t)(ctx->n ? ctx->n : 1) * K * sizeof(double)));
    ctx->d_spline = ctx->spline_own.as<double>();
    ctx->K = K;
    ctx->spline_version++;
    ctx->have_spline = true;
    ctx->have_edges = false;
    if (ctx->n == 0) return SCEMA_OK;

    int rc = ensure_tables(ctx, P);
    if (rc) return rc;

    const uint64_t n_groups = (ctx->n + GROUP - 1) / GROUP;
    const size_t slab = (size_t)ctx->max_len * 32 * sizeof(double);
    t_begin(ctx, SCEMA_T_RESAMPLE);
    if (slab <= ctx->smem_optin) {
        int per_sm = (int)(ctx->smem_optin / (slab + 1024));
        if (per_sm > 32) per_sm = 32;
        if (per_sm < 1) per_sm = 1;
        uint64_t grid = (uint64_t)ctx->sm_count * per_sm;
        if (grid > n_groups) grid = n_groups;
        SCEMA_CUDA(ctx, cudaFuncSetAttribute(k_resample<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)slab));
        k_resample<true><<<(unsigned)grid, 32, slab, ctx->stream>>>(
            ctx->d_steps, ctx->d_offsets.as<uint64_t>(), ctx->d_table_index.as<int64_t>(),
            ctx->d_tables.as<double>(), P, ctx->n, ctx->spline_own.as<double>(), nullptr);
    } else {
        // very long histories: z/b sweep buffers in global memory (same [step][6] shape as the input)
        SCEMA_CUDA(ctx, ctx->zscratch.reserve((size_t)ctx->total_steps * 6 * sizeof(double)));
        uint64_t grid = (uint64_t)ctx->sm_count * 32;
        if (grid > n_groups) grid = n_groups;
        k_resample<false><<<(unsigned)grid, 32, 0, ctx->stream>>>(
            ctx->d_steps, ctx->d_offsets.as<uint64_t>(), ctx->d_table_index.as<int64_t>(),
            ctx->d_tables.as<double>(), P, ctx->n, ctx->spline_own.as<double>(), ctx->zscratch.as<double>());
    }
    ctx->launches++;
    t_end(ctx, SCEMA_T_RESAMPLE);
    SCEMA_CUDA(ctx, cudaGetLastError());
    return SCEMA_OK;
}

}  // namespace scema
