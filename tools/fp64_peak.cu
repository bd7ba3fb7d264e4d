// FP64 issue-rate microbenchmark for B200 (sm_100a): the roofline denominators that
// MEASURED_PEAKS.json does not carry ("FP64_peak must be measured on the box").
//   - DFMA: register-resident independent FMA chains (vector FP64 pipe)
//   - DMMA: mma.sync.aligned.{m8n8k4,m16n8k4,m16n8k8,m16n8k16}.f64 (FP64 tensor path)
//   - dependent-chain latencies of DADD/DMUL/DFMA/DDIV (they bound the K1 tridiagonal sweeps)
// Prints one JSON object. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(2); } } while (0)

constexpr int ITERS = 4096;

__global__ void __launch_bounds__(256) k_dfma(double *out, double a, double b)
{
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += acc[i];
    if (s == 123.456) out[0] = s;
}

template <int SHAPE>  // 0: m8n8k4, 1: m16n8k4, 2: m16n8k8, 3: m16n8k16
__global__ void __launch_bounds__(256) k_dmma(double *out, double av, double bv)
{
    constexpr int NACC = 8;  // independent accumulator tiles per warp
    double c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0; }
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = av + threadIdx.x * 1e-12 + i;
#pragma unroll
    for (int i = 0; i < 4; i++) b[i] = bv + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            if (SHAPE == 0) {
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a[0]), "d"(b[0]));
            } else if (SHAPE == 1) {
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                             : "d"(a[0]), "d"(a[1]), "d"(b[0]));
            } else if (SHAPE == 2) {
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
            } else {
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                               "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456) out[0] = s;
}


// Do the FP64 tensor path (DMMA) and the vector FP64 pipe (DFMA) share execution resources? MODE 0:
// odd warps issue DMMA, even warps DFMA; MODE 1: every warp interleaves ND DMMAs with NF DFMAs per
// iteration (compile-time counts, fully unrolled). Independent pipes would give dmma + dfma.
template <int MODE, int ND, int NF>
__global__ void __launch_bounds__(256) k_mixed(double *out, double a, double b)
{
    double c[ND > 0 ? ND : 1][2], acc[NF > 0 ? NF : 1];
#pragma unroll
    for (int i = 0; i < ND; i++) c[i][0] = c[i][1] = 0.0;
#pragma unroll
    for (int i = 0; i < NF; i++) acc[i] = threadIdx.x * 1e-9 + i;
    const bool tensor_warp = MODE == 1 || ((threadIdx.x >> 5) & 1);
    const bool vector_warp = MODE == 1 || !((threadIdx.x >> 5) & 1);
    if (MODE == 1) {
        for (int it = 0; it < ITERS; it++) {
#pragma unroll
            for (int i = 0; i < (ND > NF ? ND : NF); i++) {
                if (i < ND)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                 : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
                if (i < NF) acc[i] = fma(acc[i], a, b);
            }
        }
    } else if (tensor_warp) {
        for (int it = 0; it < ITERS; it++) {
#pragma unroll
            for (int i = 0; i < ND; i++)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
    } else if (vector_warp) {
        for (int it = 0; it < ITERS; it++) {
#pragma unroll
            for (int i = 0; i < NF; i++) acc[i] = fma(acc[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ND; i++) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < NF; i++) s += acc[i];
    if (s == 123.456) out[0] = s;
}

// DMMA issue rate at the occupancy of the K2 filter: `blocks` CTAs of 256 threads with a large
// dynamic shared-memory request (one CTA per SM -> 2 warps per scheduler), NACC independent
// accumulator fragments per warp (the filter has 32).
template <int NACC>
__global__ void __launch_bounds__(256, 1) k_dmma_occ(double *out, double av, double bv)
{
    extern __shared__ double pad[];
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = 0.0;
    const double a = av + threadIdx.x * 1e-12, b = bv;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s + pad[0];
}

// one warp, one thread active: dependent chain latency in cycles per op
template <int OP>
__global__ void k_lat(double *out, long long *cyc, double a, double b)
{
    double x = a;
    long long t0 = clock64();
#pragma unroll 16
    for (int it = 0; it < 4096; it++) {
        if (OP == 0) x = __dadd_rn(x, b);
        else if (OP == 1) x = __dmul_rn(x, b);
        else if (OP == 2) x = fma(x, b, a);
        else x = b / x + a;  // IEEE division + add
    }
    long long t1 = clock64();
    out[1] = x;
    cyc[OP] = t1 - t0;
}

static float time_kernel(void (*launch)(), int reps)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

static double *g_out;
static int g_blocks;
static void l_dfma() { k_dfma<<<g_blocks, 256>>>(g_out, 1.0000001, 1e-9); }
static void l_dmma0() { k_dmma<0><<<g_blocks, 256>>>(g_out, 1.0, 1e-9); }
static void l_dmma1() { k_dmma<1><<<g_blocks, 256>>>(g_out, 1.0, 1e-9); }
static void l_dmma2() { k_dmma<2><<<g_blocks, 256>>>(g_out, 1.0, 1e-9); }
static void l_dmma3() { k_dmma<3><<<g_blocks, 256>>>(g_out, 1.0, 1e-9); }
template <int MODE, int ND, int NF> static void l_mix() { k_mixed<MODE, ND, NF><<<g_blocks, 256>>>(g_out, 1.0000001, 1e-9); }

int main()
{
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    g_blocks = sms * 8;  // 8 CTAs x 256 threads = full occupancy
    CK(cudaMalloc(&g_out, 64));
    long long *cyc; CK(cudaMalloc(&cyc, 64));
    double threads = (double)g_blocks * 256, warps = threads / 32;
    float t;
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
    t = time_kernel(l_dfma, 5);
    printf(", \"dfma_tflops\": %.3f", threads * ITERS * 16 * 2 / (t * 1e-3) / 1e12);
    const double fl[4] = {8 * 8 * 4 * 2.0, 16 * 8 * 4 * 2.0, 16 * 8 * 8 * 2.0, 16 * 8 * 16 * 2.0};
    void (*ls[4])() = {l_dmma0, l_dmma1, l_dmma2, l_dmma3};
    const char *nm[4] = {"dmma_m8n8k4_tflops", "dmma_m16n8k4_tflops", "dmma_m16n8k8_tflops", "dmma_m16n8k16_tflops"};
    for (int s = 0; s < 4; s++) {
        t = time_kernel(ls[s], 5);
        printf(", \"%s\": %.3f", nm[s], warps * ITERS * 8 * fl[s] / (t * 1e-3) / 1e12);
    }

    // mixed issue: flops of both kinds over the common time
    {
        struct Cfg { int mode, nd, nf; void (*fn)(); };
        const Cfg cfg[] = {
            {0, 8, 0, l_mix<0, 8, 0>}, {0, 0, 16, l_mix<0, 0, 16>}, {0, 8, 16, l_mix<0, 8, 16>}, {0, 8, 4, l_mix<0, 8, 4>},
            {1, 8, 0, l_mix<1, 8, 0>}, {1, 0, 16, l_mix<1, 0, 16>}, {1, 8, 16, l_mix<1, 8, 16>}, {1, 8, 8, l_mix<1, 8, 8>},
            {1, 8, 4, l_mix<1, 8, 4>}, {1, 8, 2, l_mix<1, 8, 2>}, {1, 8, 1, l_mix<1, 8, 1>}, {1, 4, 16, l_mix<1, 4, 16>},
        };
        printf(", \"mixed\": [");
        for (unsigned q = 0; q < sizeof(cfg) / sizeof(cfg[0]); q++) {
            t = time_kernel(cfg[q].fn, 5);
            const double wt = cfg[q].mode == 0 ? warps / 2 : warps, tt = cfg[q].mode == 0 ? threads / 2 : threads;
            const double f_t = wt * ITERS * cfg[q].nd * 512.0, f_v = tt * ITERS * cfg[q].nf * 2.0;
            printf("%s{\"mode\": \"%s\", \"dmma_per_iter\": %d, \"dfma_per_iter\": %d, \"ms\": %.3f, \"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f, \"sum\": %.2f}",
                   q ? ", " : "", cfg[q].mode == 0 ? "split_warps" : "interleaved", cfg[q].nd, cfg[q].nf, t,
                   f_t / (t * 1e-3) / 1e12, f_v / (t * 1e-3) / 1e12, (f_t + f_v) / (t * 1e-3) / 1e12);
        }
        printf("]");
    }
    {
        // one CTA per SM (180 KB of dynamic shared memory), 8 warps -> 2 per scheduler
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaFuncSetAttribute(k_dmma_occ<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 180 * 1024));
        CK(cudaFuncSetAttribute(k_dmma_occ<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 180 * 1024));
        printf(", \"dmma_2warps_per_scheduler\": {");
        for (int v = 0; v < 2; v++) {
            float best = 1e30f;
            for (int r = 0; r < 6; r++) {
                CK(cudaEventRecord(e0));
                if (v == 0) k_dmma_occ<8><<<sms, 256, 180 * 1024>>>(g_out, 1.0, 1e-9);
                else k_dmma_occ<32><<<sms, 256, 180 * 1024>>>(g_out, 1.0, 1e-9);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (r && ms < best) best = ms;
            }
            const int nacc = v == 0 ? 8 : 32;
            printf("%s\"nacc%d_tflops\": %.3f", v ? ", " : "", nacc, (double)sms * 8 * ITERS * nacc * 512.0 / (best * 1e-3) / 1e12);
        }
        printf("}");
    }
    k_lat<0><<<1, 1>>>(g_out, cyc, 1.0, 1e-9);
    k_lat<1><<<1, 1>>>(g_out, cyc, 1.0, 1.0000001);
    k_lat<2><<<1, 1>>>(g_out, cyc, 1.0, 0.999);
    k_lat<3><<<1, 1>>>(g_out, cyc, 1.0, 0.999);
    CK(cudaDeviceSynchronize());
    long long h[4]; CK(cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost));
    printf(", \"lat_cycles\": {\"dadd\": %.2f, \"dmul\": %.2f, \"dfma\": %.2f, \"ddiv_plus_dadd\": %.2f}",
           h[0] / 4096.0, h[1] / 4096.0, h[2] / 4096.0, h[3] / 4096.0);
    int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    printf(", \"sm_clock_khz_max\": %d}\n", clk);
    return 0;
}
