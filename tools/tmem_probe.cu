// Stand-alone probe of the TMEM read-out rate on B200 (sm_100a): how many bytes per clock and SM can
// tcgen05.ld move out of tensor memory, for which shape, from how many warps?
//
// Why: the one-slice tcgen05 filter (scema_b200/csrc/pairs_tc.cu, k_filter_tc) has to read every pair's
// fp32 accumulator once. Round 1 inferred a ceiling of 128 B/clk/SM from the kernel's own speed; this probe
// measures the rate in isolation: W warps (W = 4, 8, 16; warp w reads the lane quarter w % 4) issue
// nothing but tcgen05.ld against a resident 128 x 512-column allocation, for the shapes 32x32b.x{16,32,64,128}
// and 16x256b.x{8,16,32}, waiting (tcgen05.wait::ld) after every load or with two loads in flight per wait, and — the
// `consume` variants — AND-reducing the loaded registers like the filter's epilogue does.
// One CTA per SM, clock64() around the loop, median over the CTAs; prints one JSON object.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo tools/tmem_probe.cu -o tools/tmem_probe
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <string>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(2); } } while (0)

// register name lists: eight groups a..h of 16 registers declared inside each asm block
#define L16(p) p "0," p "1," p "2," p "3," p "4," p "5," p "6," p "7," p "8," p "9," p "10," p "11," p "12," p "13," p "14," p "15"
#define REGS16 L16("a")
#define REGS32 L16("a") "," L16("b")
#define REGS64 L16("a") "," L16("b") "," L16("c") "," L16("d")
#define REGS128 L16("a") "," L16("b") "," L16("c") "," L16("d") "," L16("e") "," L16("f") "," L16("g") "," L16("h")
#define DECL ".reg .b32 a<16>, b<16>, c<16>, d<16>, e<16>, f<16>, g<16>, h<16>;\n"
// AND-reduce of one 16-register group into %0 (three-input LOP3 would halve it; and.b32 is what nvcc emits before ptxas fuses)
#define AND16(p) "and.b32 %0, %0, " p "0; and.b32 %0, %0, " p "1; and.b32 %0, %0, " p "2; and.b32 %0, %0, " p "3;" \
                 "and.b32 %0, %0, " p "4; and.b32 %0, %0, " p "5; and.b32 %0, %0, " p "6; and.b32 %0, %0, " p "7;" \
                 "and.b32 %0, %0, " p "8; and.b32 %0, %0, " p "9; and.b32 %0, %0, " p "10; and.b32 %0, %0, " p "11;" \
                 "and.b32 %0, %0, " p "12; and.b32 %0, %0, " p "13; and.b32 %0, %0, " p "14; and.b32 %0, %0, " p "15;\n"
#define TOUCH(p) "and.b32 %0, %0, " p "0; and.b32 %0, %0, " p "15;\n"

enum Shape { S32x32_16, S32x32_32, S32x32_64, S32x32_128, S16x256_8, S16x256_16, S16x256_32, N_SHAPES };
static const char *shape_name[N_SHAPES] = {"32x32b.x16", "32x32b.x32", "32x32b.x64", "32x32b.x128", "16x256b.x8", "16x256b.x16", "16x256b.x32"};
// bytes one warp-wide instruction moves
static const int shape_bytes[N_SHAPES] = {32 * 16 * 4, 32 * 32 * 4, 32 * 64 * 4, 32 * 128 * 4, 16 * 32 * 8, 16 * 32 * 16, 16 * 32 * 32};
// TMEM columns one instruction spans
static const int shape_cols[N_SHAPES] = {16, 32, 64, 128, 8 * 8, 16 * 8, 32 * 8};

#define LD_VARIANT(NAME, SHAPESTR, REGLIST, BODY)                                                      \
    __device__ __forceinline__ void NAME(uint32_t t, uint32_t &acc)                                     \
    {                                                                                                   \
        asm volatile("{\n" DECL "tcgen05.ld.sync.aligned." SHAPESTR ".b32 {" REGLIST "}, [%1];\n" BODY "}" \
                     : "+r"(acc) : "r"(t) : "memory");                                                  \
    }
#define WAITLD "tcgen05.wait::ld.sync.aligned;\n"
// w = wait after the load, n = no wait (the caller waits later), t = touch two registers, c = AND-reduce all
LD_VARIANT(ld_32_16_w, "32x32b.x16", REGS16, WAITLD TOUCH("a"))
LD_VARIANT(ld_32_16_c, "32x32b.x16", REGS16, WAITLD AND16("a"))
LD_VARIANT(ld_32_32_w, "32x32b.x32", REGS32, WAITLD TOUCH("b"))
LD_VARIANT(ld_32_32_c, "32x32b.x32", REGS32, WAITLD AND16("a") AND16("b"))
LD_VARIANT(ld_32_64_w, "32x32b.x64", REGS64, WAITLD TOUCH("d"))
LD_VARIANT(ld_32_64_c, "32x32b.x64", REGS64, WAITLD AND16("a") AND16("b") AND16("c") AND16("d"))
LD_VARIANT(ld_32_128_w, "32x32b.x128", REGS128, WAITLD TOUCH("h"))
LD_VARIANT(ld_32_128_c, "32x32b.x128", REGS128, WAITLD AND16("a") AND16("b") AND16("c") AND16("d") AND16("e") AND16("f") AND16("g") AND16("h"))
LD_VARIANT(ld_16_8_w, "16x256b.x8", REGS32, WAITLD TOUCH("b"))
LD_VARIANT(ld_16_16_w, "16x256b.x16", REGS64, WAITLD TOUCH("d"))
LD_VARIANT(ld_16_32_w, "16x256b.x32", REGS128, WAITLD TOUCH("h"))
LD_VARIANT(ld_16_32_c, "16x256b.x32", REGS128, WAITLD AND16("a") AND16("b") AND16("c") AND16("d") AND16("e") AND16("f") AND16("g") AND16("h"))

// two loads in flight, one wait (registers of both touched afterwards, so neither can be dropped)
#define LD2_VARIANT(NAME, SHAPESTR, REGS_A, REGS_B, TA, TB)                                                       \
    __device__ __forceinline__ void NAME(uint32_t t, uint32_t t2, uint32_t &acc)                                   \
    {                                                                                                              \
        asm volatile("{\n" DECL "tcgen05.ld.sync.aligned." SHAPESTR ".b32 {" REGS_A "}, [%1];\n"                   \
                     "tcgen05.ld.sync.aligned." SHAPESTR ".b32 {" REGS_B "}, [%2];\n" WAITLD TOUCH(TA) TOUCH(TB) "}" \
                     : "+r"(acc) : "r"(t), "r"(t2) : "memory");                                                    \
    }
#define REGS16B L16("b")
#define REGS32B L16("c") "," L16("d")
#define REGS64B L16("e") "," L16("f") "," L16("g") "," L16("h")
LD2_VARIANT(ld2_32_16, "32x32b.x16", REGS16, REGS16B, "a", "b")
LD2_VARIANT(ld2_32_32, "32x32b.x32", REGS32, REGS32B, "b", "d")
LD2_VARIANT(ld2_32_64, "32x32b.x64", REGS64, REGS64B, "d", "h")
LD2_VARIANT(ld2_16_8, "16x256b.x8", REGS32, REGS32B, "b", "d")
LD2_VARIANT(ld2_16_16, "16x256b.x16", REGS64, REGS64B, "d", "h")

// MODE: 0 = wait after every load, 1 = two loads in flight per wait (ld_pair), 2 = wait + AND-reduce everything (epilogue-like)
template <int S>
__device__ __forceinline__ void ld_pair(uint32_t t, uint32_t t2, uint32_t &acc)
{
    if (S == S32x32_16) ld2_32_16(t, t2, acc);
    if (S == S32x32_32) ld2_32_32(t, t2, acc);
    if (S == S32x32_64) ld2_32_64(t, t2, acc);
    if (S == S16x256_8) ld2_16_8(t, t2, acc);
    if (S == S16x256_16) ld2_16_16(t, t2, acc);
}
template <int S, int MODE>
__device__ __forceinline__ void ld_dispatch(uint32_t t, uint32_t &acc)
{
    if (MODE == 0) {
        if (S == S32x32_16) ld_32_16_w(t, acc);
        if (S == S32x32_32) ld_32_32_w(t, acc);
        if (S == S32x32_64) ld_32_64_w(t, acc);
        if (S == S32x32_128) ld_32_128_w(t, acc);
        if (S == S16x256_8) ld_16_8_w(t, acc);
        if (S == S16x256_16) ld_16_16_w(t, acc);
        if (S == S16x256_32) ld_16_32_w(t, acc);
    } else {
        if (S == S32x32_16) ld_32_16_c(t, acc);
        if (S == S32x32_32) ld_32_32_c(t, acc);
        if (S == S32x32_64) ld_32_64_c(t, acc);
        if (S == S32x32_128) ld_32_128_c(t, acc);
        if (S == S16x256_32) ld_16_32_c(t, acc);
    }
}

template <int S, int MODE, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_probe(int iters, int span_cols, long long *cycles, uint32_t *sink)
{
    __shared__ uint32_t s_tmem;
    const uint32_t warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = s_tmem + (((warp & 3u) * 32u) << 16);
    // every warp walks the 512 columns in steps of the shape's span; warps sharing a lane quarter start half a
    // window apart so they do not read the same columns at the same moment
    const int n_pos = 512 / span_cols;
    int pos = (int)((warp >> 2) * (n_pos / 2 ? n_pos / 2 : 1)) % n_pos;
    uint32_t acc = 0xffffffffu;
    __syncthreads();
    const long long t0 = clock64();
    if (MODE == 1) {
        for (int it = 0; it < iters; it += 2) {
            const int pos2 = pos + 1 == n_pos ? 0 : pos + 1;
            ld_pair<S>(base + (uint32_t)(pos * span_cols), base + (uint32_t)(pos2 * span_cols), acc);
            pos = pos2 + 1 == n_pos ? 0 : pos2 + 1;
        }
    } else {
        for (int it = 0; it < iters; it++) {
            ld_dispatch<S, MODE>(base + (uint32_t)(pos * span_cols), acc);
            pos = pos + 1 == n_pos ? 0 : pos + 1;
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(s_tmem) : "memory");
}

template <int S, int MODE, int MAXT>
static double run(int warps, int sms, long long *d_cyc, uint32_t *d_sink, double *ms_out)
{
    const int iters = 4096;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    std::vector<long long> h(sms);
    double best = 0.0;
    float best_ms = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        CK(cudaEventRecord(e0));
        k_probe<S, MODE, MAXT><<<sms, warps * 32>>>(iters, shape_cols[S], d_cyc, d_sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        CK(cudaMemcpy(h.data(), d_cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
        std::sort(h.begin(), h.end());
        const double cyc = (double)h[sms / 2];
        const double bpc = (double)warps * iters * shape_bytes[S] / cyc;
        if (rep > 0 && bpc > best) best = bpc;
        if (rep > 0 && ms < best_ms) best_ms = ms;
    }
    *ms_out = best_ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return best;
}

int main()
{
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    long long *d_cyc;
    uint32_t *d_sink;
    CK(cudaMalloc(&d_cyc, sms * sizeof(long long)));
    CK(cudaMalloc(&d_sink, 64));
    printf("{\"device\": \"%s\", \"sms\": %d, \"unit\": \"bytes per clock per SM (median CTA, clock64 around 4096 loads per warp)\", \"results\": [\n", p.name, sms);
    bool first = true;
    auto emit = [&](const char *shape, const char *mode, int warps, double bpc, double ms) {
        printf("%s  {\"shape\": \"%s\", \"mode\": \"%s\", \"warps\": %d, \"bytes_per_clk_sm\": %.1f, \"kernel_ms\": %.3f}", first ? "" : ",\n", shape, mode, warps, bpc, ms);
        first = false;
    };
    const int wl[3] = {4, 8, 16};
    const char *modes[3] = {"wait_each", "two_loads_per_wait", "wait_each+and_all"};
    // 16 warps leave 128 registers per thread: the 128-register loads only run with 4 and 8 warps
#define RUN(S, M)                                                                  \
    for (int wi = 0; wi < 3; wi++) {                                               \
        double ms;                                                                 \
        constexpr bool wide = (S == S32x32_128 || S == S16x256_32);                \
        if (wide && wl[wi] > 8) continue;                                          \
        double b;                                                                  \
        if constexpr (wide) b = run<S, M, 256>(wl[wi], sms, d_cyc, d_sink, &ms);   \
        else b = run<S, M, 512>(wl[wi], sms, d_cyc, d_sink, &ms);                  \
        emit(shape_name[S], modes[M], wl[wi], b, ms);                              \
    }
    RUN(S32x32_16, 0) RUN(S32x32_32, 0) RUN(S32x32_64, 0) RUN(S32x32_128, 0)
    RUN(S16x256_8, 0) RUN(S16x256_16, 0) RUN(S16x256_32, 0)
    RUN(S32x32_16, 1) RUN(S32x32_32, 1) RUN(S32x32_64, 1) RUN(S16x256_8, 1) RUN(S16x256_16, 1)
    RUN(S32x32_16, 2) RUN(S32x32_32, 2) RUN(S32x32_64, 2) RUN(S32x32_128, 2) RUN(S16x256_32, 2)
    printf("\n]}\n");
    return 0;
}
