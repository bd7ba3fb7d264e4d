/* scema_synth.h — deterministic synthetic strain histories for benchmarks and parity tests.
 *
 * Not part of the reference's interface: SCEMa gets its histories from the FE solver
 * (FE_problem.h:1091-1098). This generator exists because the benchmark configurations
 * (BASELINE.json configs 2-5) are synthetic. It is counter-based (splitmix64 finaliser of
 * (seed, a, b, stream)) and uses only individually rounded +, *, / so that
 * scema_b200/synth.py (numpy) reproduces every double bit-for-bit on the host.
 *
 * Cluster model: histories come in clusters of `cluster_size` consecutive
 * indices. A cluster has a common length and a common smooth centre path
 *   centre_c(t) = A_c * (t + beta_c * t*t),  A_c in amp*(-1,1), beta_c in (-0.5,0.5),  t = step/(L-1)
 * and member i adds delta_ic * t with delta_ic in pert*(-1,1). Members of one cluster end up a
 * distance of order pert from each other after resampling; different clusters are ~amp apart.
 */
#ifndef SCEMA_SYNTH_H
#define SCEMA_SYNTH_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* All functions generate histories first .. first+n-1 of the infinite synthetic population, so a
 * rank can generate just its own shard. Offsets are relative to the shard (offsets[0] == 0).
 * Host: offsets[n+1] (in steps) for cluster-wise lengths uniform in [len_min, len_max]. */
int scema_synth_offsets(uint64_t seed, uint64_t first, uint64_t n, uint32_t cluster_size, uint32_t len_min,
                        uint32_t len_max, uint64_t *offsets_host);
/* Device: fill d_steps [offsets[n]][6] given device offsets. stream: cudaStream_t as void*. */
int scema_synth_histories_device(uint64_t seed, uint64_t first, uint64_t n, uint32_t cluster_size, double amp, double pert,
                                 const uint64_t *d_offsets, double *d_steps, void *stream);
/* Device: already-resampled rows d_rows [n][6*spline_points] in the reference's p*6+c order. */
int scema_synth_rows_device(uint64_t seed, uint64_t first, uint64_t n, uint32_t cluster_size, uint32_t spline_points, double amp,
                            double pert, double *d_rows, void *stream);


/* Same with a choice of population model. model 0: the cluster model above (spread unused). model 1, "smooth"
 * (emulation of the reference's dogbone stretch, input_configurations/inputs_dogbone_cuboid.json:
 * every quadrature point follows nearly the same path): zz(t) = amp t (1 + delta_q), xx = yy = -0.3 zz,
 * shear_c = zz sigma_qc, delta_q in spread (-1,1) and sigma_qc in 0.01 (-1,1) per group q of cluster_size symmetric
 * points, members jittered by pert (-1,1) t per component as in model 0. All row norms are ~ 2 amp, group
 * distances ~ spread amp. */
int scema_synth_histories_model_device(int model, double spread, uint64_t seed, uint64_t first, uint64_t n, uint32_t cluster_size,
                                       double amp, double pert, const uint64_t *d_offsets, double *d_steps, void *stream);
int scema_synth_rows_model_device(int model, double spread, uint64_t seed, uint64_t first, uint64_t n, uint32_t cluster_size,
                                  uint32_t spline_points, double amp, double pert, double *d_rows, void *stream);

#ifdef __cplusplus
}
#endif
#endif
