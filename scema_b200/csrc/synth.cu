// Deterministic synthetic histories (include/scema_synth.h). Bit-identical twin: scema_b200/synth.py.
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/scema_hist.h"
#include "../../include/scema_synth.h"

namespace {

__host__ __device__ inline uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__host__ __device__ inline uint64_t hash4(uint64_t seed, uint64_t a, uint64_t b, uint64_t stream)
{
    uint64_t z = mix64(seed + 0x9e3779b97f4a7c15ull * (a + 1));
    z = mix64(z + b);
    return mix64(z + stream);
}
__host__ __device__ inline double u01(uint64_t x) { return (double)(x >> 11) * 1.1102230246251565e-16; }

__host__ __device__ inline uint32_t cluster_len(uint64_t seed, uint64_t q, uint32_t lmin, uint32_t lmax)
{
    return lmin + (uint32_t)(hash4(seed, q, 0, 7) % (uint64_t)(lmax - lmin + 1));
}

// value of component c of history i (cluster q) at abscissa t; every operation individually rounded
__device__ inline double synth_value(uint64_t seed, uint64_t i, uint64_t q, int c, double t, double amp, double pert)
{
    const double A = __dmul_rn(amp, __dsub_rn(__dmul_rn(2.0, u01(hash4(seed, q, c, 1))), 1.0));
    const double beta = __dsub_rn(u01(hash4(seed, q, c, 2)), 0.5);
    const double delta = __dmul_rn(pert, __dsub_rn(__dmul_rn(2.0, u01(hash4(seed, i, c, 3))), 1.0));
    const double centre = __dmul_rn(A, __dadd_rn(t, __dmul_rn(beta, __dmul_rn(t, t))));
    return __dadd_rn(centre, __dmul_rn(delta, t));
}

// "smooth" model (the dogbone emulation): every history follows the SAME stretch path, scaled by a
// per-group factor: zz = amp t (1 + delta_q), xx = yy = -0.3 zz, shear_c = zz sigma_qc with delta_q in spread (-1,1),
// sigma_qc in 1e-2 (-1,1); the members of a group (the symmetric quadrature points) differ by the same jitter as in
// the cluster model. Row norms are all ~ 2 amp while distances between groups are ~ spread amp: a filter whose guard
// band scales with the row norms keeps every pair unless the common path is subtracted first.
__device__ inline double synth_value_smooth(uint64_t seed, uint64_t i, uint64_t q, int c, double t, double amp, double pert,
                                            double spread)
{
    const double dq = __dmul_rn(spread, __dsub_rn(__dmul_rn(2.0, u01(hash4(seed, q, 0, 11))), 1.0));
    const double zz = __dmul_rn(__dmul_rn(amp, t), __dadd_rn(1.0, dq));
    double centre;
    if (c == 2) centre = zz;
    else if (c < 2) centre = __dmul_rn(-0.3, zz);
    else centre = __dmul_rn(zz, __dmul_rn(0.01, __dsub_rn(__dmul_rn(2.0, u01(hash4(seed, q, c, 12))), 1.0)));
    const double delta = __dmul_rn(pert, __dsub_rn(__dmul_rn(2.0, u01(hash4(seed, i, c, 3))), 1.0));
    return __dadd_rn(centre, __dmul_rn(delta, t));
}

__device__ inline double synth_any(int model, uint64_t seed, uint64_t i, uint64_t q, int c, double t, double amp, double pert,
                                   double spread)
{
    return model == 1 ? synth_value_smooth(seed, i, q, c, t, amp, pert, spread) : synth_value(seed, i, q, c, t, amp, pert);
}

__global__ void k_synth_hist(int model, double spread, uint64_t seed, uint64_t first, uint64_t n, uint32_t cs, double amp, double pert,
                             const uint64_t *__restrict__ offsets, double *__restrict__ steps)
{
    // one warp per history, lanes stride the steps
    const uint64_t li = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (li >= n) return;
    const int lane = threadIdx.x & 31;
    const uint64_t off = offsets[li];
    const uint32_t L = (uint32_t)(offsets[li + 1] - off);
    const uint64_t i = first + li;
    const uint64_t q = i / cs;
    for (uint32_t e = lane; e < L * 6; e += 32) {
        const uint32_t s = e / 6;
        const int c = (int)(e - s * 6);
        const double t = __ddiv_rn((double)s, (double)(L - 1));
        steps[off * 6 + e] = synth_any(model, seed, i, q, c, t, amp, pert, spread);
    }
}

__global__ void k_synth_rows(int model, double spread, uint64_t seed, uint64_t first, uint64_t n, uint32_t cs, uint32_t P, double amp, double pert,
                             double *__restrict__ rows)
{
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t K = 6 * P;
    if (idx >= n * K) return;
    const uint64_t li = idx / K, i = first + li;
    const uint32_t k = (uint32_t)(idx - li * K), p = k / 6;
    const int c = (int)(k - p * 6);
    const double t = __ddiv_rn((double)p, (double)(P - 1));
    rows[idx] = synth_any(model, seed, i, i / cs, c, t, amp, pert, spread);
}

}  // namespace

extern "C" {

int scema_synth_offsets(uint64_t seed, uint64_t first, uint64_t n, uint32_t cluster_size, uint32_t len_min,
                        uint32_t len_max, uint64_t *offsets_host)
{
    if (!offsets_host || cluster_size == 0 || len_min < 2 || len_max < len_min) return SCEMA_ERR_INVALID;
    offsets_host[0] = 0;
    for (uint64_t i = 0; i < n; i++) offsets_host[i + 1] = offsets_host[i] + cluster_len(seed, (first + i) / cluster_size, len_min, len_max);
    return SCEMA_OK;
}

int scema_synth_histories_model_device(int model, double spread, uint64_t seed, uint64_t first, uint64_t n, uint32_t cluster_size,
                                       double amp, double pert, const uint64_t *d_offsets, double *d_steps, void *stream)
{
    if (cluster_size == 0 || model < 0 || model > 1) return SCEMA_ERR_INVALID;
    if (n == 0) return SCEMA_OK;
    const uint64_t blocks = (n * 32 + 255) / 256;
    k_synth_hist<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(model, spread, seed, first, n, cluster_size, amp, pert, d_offsets,
                                                                     d_steps);
    return cudaGetLastError() == cudaSuccess ? SCEMA_OK : SCEMA_ERR_CUDA;
}

int scema_synth_histories_device(uint64_t seed, uint64_t first, uint64_t n, uint32_t cluster_size, double amp, double pert,
                                 const uint64_t *d_offsets, double *d_steps, void *stream)
{
    return scema_synth_histories_model_device(0, 0.0, seed, first, n, cluster_size, amp, pert, d_offsets, d_steps, stream);
}

int scema_synth_rows_model_device(int model, double spread, uint64_t seed, uint64_t first, uint64_t n, uint32_t cluster_size,
                                  uint32_t spline_points, double amp, double pert, double *d_rows, void *stream)
{
    if (cluster_size == 0 || spline_points < 2 || model < 0 || model > 1) return SCEMA_ERR_INVALID;
    if (n == 0) return SCEMA_OK;
    const uint64_t total = n * 6 * spline_points;
    k_synth_rows<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(model, spread, seed, first, n, cluster_size,
                                                                                   spline_points, amp, pert, d_rows);
    return cudaGetLastError() == cudaSuccess ? SCEMA_OK : SCEMA_ERR_CUDA;
}

int scema_synth_rows_device(uint64_t seed, uint64_t first, uint64_t n, uint32_t cluster_size, uint32_t spline_points, double amp,
                            double pert, double *d_rows, void *stream)
{
    return scema_synth_rows_model_device(0, 0.0, seed, first, n, cluster_size, spline_points, amp, pert, d_rows, stream);
}

}  // extern "C"
