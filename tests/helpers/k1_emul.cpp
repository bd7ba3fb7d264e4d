// Host emulation of k_resample_pair (scema_b200/csrc/resample_pair.cuh): the KERNEL SOURCE is compiled for the CPU
// and run thread by thread, so that the indexing of the prefetch rings, the peeled first / last steps, the pairing of
// the groups, the chunk hand-out and the slow-path hand-over are checked against the CPU oracle without a GPU
// (tests/test_k1_emul.py). Test infrastructure only: nothing under scema_b200/ links or loads this.
//
// What is modelled:
//  * a CTA = 128 host threads with a real barrier for __syncthreads(); several CTAs run concurrently and share the
//    chunk counter (atomicAdd);
//  * shared memory = one byte array per CTA, addressed by 32-bit offsets exactly as the kernel's ld.shared / cp.async
//    operands are; it starts out filled with signalling garbage, so a slot read before its copy landed shows;
//  * cp.async in three timings: 0 = every copy lands when it is issued (the earliest legal moment: exposes a ring slot
//    overwritten before its last read), 1 = a copy lands only when a wait_group forces it, reading its SOURCE only then
//    (the latest legal moment: exposes a slot read before its copy was waited for, and a source overwritten while a copy
//    is in flight), 2 = a pseudo-random mix of the two;
//  * __d*_rn / __fma_rn = the host's IEEE operations (built with -ffp-contract=off).
// What is not: scheduling, scoreboards, memory-model subtleties between warps (the kernel has no inter-warp data flow
// besides the factor table behind __syncthreads()).
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <thread>
#include <vector>

#define K1_EMULATE
#define K1_HD
#define K1_DEV static inline
#define K1_GLOBAL static
#define K1_RESTRICT

struct double2 {
    double x, y;
};
struct EmuIdx {
    unsigned x = 0;
};
struct EmuCta {
    std::vector<unsigned char> smem;
    uint32_t s_chunk = 0;
    std::barrier<> *bar = nullptr;
};
struct EmuCopy {
    uint32_t dst;
    const double *src;
};
struct EmuThread {
    EmuCta *cta = nullptr;
    std::deque<std::vector<EmuCopy>> groups;  // committed, not yet landed (oldest first)
    std::vector<EmuCopy> open;                // issued since the last commit
    int mode = 0;
    uint64_t rng = 0x9E3779B97F4A7C15ull;
    uint64_t copies = 0, late = 0;
};
static thread_local EmuIdx threadIdx, blockIdx, blockDim;
static thread_local EmuThread emu;

// every global-memory address the kernel touches must lie inside one of the launch's buffers; every shared-memory address
// a lane reads or fills inside its own warp's ring (tables: inside the table area)
struct EmuRange {
    const unsigned char *b, *e;
};
static std::vector<EmuRange> g_ranges;
static std::atomic<uint64_t> g_oob{0};
static inline void emu_check_global(const void *p, size_t bytes)
{
    const unsigned char *q = (const unsigned char *)p;
    for (const EmuRange &r : g_ranges)
        if (q >= r.b && q + bytes <= r.e) return;
    g_oob++;
}
static uint32_t g_ring_bytes_per_warp = 4096;  // PR_DEPTH * PR_ROW * 8, set before a launch
static inline void emu_check_ring(uint32_t a)
{
    const uint32_t warp = threadIdx.x >> 5;
    if (a < warp * g_ring_bytes_per_warp || a + 8 > (warp + 1) * g_ring_bytes_per_warp) g_oob++;
}

static inline void __syncthreads() { emu.cta->bar->arrive_and_wait(); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <class T>
static inline T __ldg(const T *p) { emu_check_global(p, sizeof(T)); return *p; }
static inline void __stcg(double *p, double v) { emu_check_global(p, 8); *p = v; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline int __double2hiint(double d) { uint64_t u; memcpy(&u, &d, 8); return (int)(uint32_t)(u >> 32); }
static inline int __double2loint(double d) { uint64_t u; memcpy(&u, &d, 8); return (int)(uint32_t)u; }
static inline double __hiloint2double(int hi, int lo)
{
    const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double d;
    memcpy(&d, &u, 8);
    return d;
}
#define K1_SHARED_DECL(arr, chunkvar)                                    \
    double *arr = reinterpret_cast<double *>(emu.cta->smem.data());      \
    uint32_t &chunkvar = emu.cta->s_chunk;

namespace scema {
static inline void emu_land(const EmuCopy &c)
{
    emu_check_global(c.src, 8);
    emu_check_ring(c.dst);
    memcpy(emu.cta->smem.data() + c.dst, c.src, 8);
}
static inline bool emu_coin()
{
    emu.rng ^= emu.rng << 13; emu.rng ^= emu.rng >> 7; emu.rng ^= emu.rng << 17;
    return (emu.rng >> 33) & 1;
}
template <int OFF>
static inline void pr_cp_async8(uint32_t dst, const double *src)
{
    emu.copies++;
    if (emu.mode == 0 || (emu.mode == 2 && emu_coin())) emu_land(EmuCopy{dst + (uint32_t)OFF, src});
    else emu.open.push_back(EmuCopy{dst + (uint32_t)OFF, src});
}
template <int OFF>
static inline void pr_cp_async8_hint(uint32_t dst, const double *src, uint64_t) { pr_cp_async8<OFF>(dst, src); }
static inline uint64_t pr_policy(bool) { return 0; }
static inline uint64_t pr_policy_z(bool) { return 0; }
static inline void pr_store_z(double *p, double v, uint64_t) { emu_check_global(p, 8); *p = v; }
static inline void pr_cp_async_commit()
{
    emu.groups.push_back(std::move(emu.open));
    emu.open.clear();
}
template <int N>
static inline void pr_cp_async_wait()
{
    while ((int)emu.groups.size() > N) {
        for (const EmuCopy &c : emu.groups.front()) { emu_land(c); emu.late++; }
        emu.groups.pop_front();
    }
    if (emu.mode == 2)  // some of the younger groups may have landed as well, oldest first
        while (!emu.groups.empty() && emu_coin()) {
            for (const EmuCopy &c : emu.groups.front()) emu_land(c);
            emu.groups.pop_front();
        }
}
template <int OFF>
static inline double pr_ring_read(uint32_t a)
{
    double v;
    emu_check_ring(a + OFF);
    memcpy(&v, emu.cta->smem.data() + a + OFF, 8);
    return v;
}
static inline uint32_t pr_smem_addr(const void *p) { return (uint32_t)((const unsigned char *)p - emu.cta->smem.data()); }
static inline void pr_prefetch_l2(const double *, uint32_t) {}
template <bool STAB>
struct TabCursor {
    const double2 *p;
    void set(const double2 *q) { p = q; }
    template <int E>
    void ld(double2 &d) const
    {
        if (STAB) {
            const unsigned char *q = (const unsigned char *)(p + E), *b = emu.cta->smem.data();
            if (q < b + 4 * g_ring_bytes_per_warp || q + 16 > b + emu.cta->smem.size()) g_oob++;
        } else
            emu_check_global(p + E, 16);
        d = p[E];
    }
    void advance(int entries) { p += entries; }
};
static inline double pr_copysign(double q, double s)
{
    const uint32_t qh = ((uint32_t)__double2hiint(q) & 0x7fffffffu) | ((uint32_t)__double2hiint(s) & 0x80000000u);
    return __hiloint2double((int)qh, __double2loint(q));
}
template <bool STAB>
static inline double pr_tab_f64(const double *p) { return *p; }
template <bool STAB>
static inline double2 pr_tab_f64x2(const double2 *p) { return *p; }
}  // namespace scema

#include "../../scema_b200/csrc/resample_pair.cuh"

using namespace scema;

namespace {

struct Launch {
    bool stab;
    const double *steps;
    const uint64_t *offsets;
    const uint32_t *order;
    uint64_t n_hist;
    const K1Chunk *chunks;
    uint32_t n_chunks;
    const int64_t *table_index;
    const double *tables;
    uint32_t P;
    double *out;
    uint32_t cap;
    uint64_t ys;
    uint32_t uniform_L;
    uint32_t flags;
};

// one launch: n_ctas CTAs of 128 threads, all resident at once
void run_launch(const Launch &a, int n_ctas, int mode, uint64_t *stats)
{
    unsigned counter = 0;
    g_ring_bytes_per_warp = PR_DEPTH * PR_ROW * 8;
    const size_t smem = PR_RING_BYTES + (a.stab ? rs_table_doubles(a.cap, a.P) * sizeof(double) : 0);
    const size_t zn = (size_t)n_ctas * RS_WARPS * a.cap * PR_ROW;
    std::vector<double> zraw(zn + 16);
    double *zscratch = zraw.data();
    while ((uintptr_t)zscratch % 128) zscratch++;  // cudaMalloc alignment: a scratch row is four whole 128-byte lines
    for (size_t q = 0; q < zn; q++) zscratch[q] = std::nan("");
    g_ranges.push_back(EmuRange{(const unsigned char *)zscratch, (const unsigned char *)(zscratch + zn)});
    std::vector<std::unique_ptr<EmuCta>> ctas;
    std::vector<std::unique_ptr<std::barrier<>>> bars;
    for (int b = 0; b < n_ctas; b++) {
        ctas.emplace_back(new EmuCta);
        bars.emplace_back(new std::barrier<>(32 * RS_WARPS));
        ctas[b]->smem.assign(smem, 0xFB);  // 0xFBFB... as a double is a huge negative number: shows in any result it reaches
        ctas[b]->bar = bars[b].get();
    }
    std::atomic<uint64_t> copies{0}, late{0};
    std::vector<std::thread> th;
    for (int b = 0; b < n_ctas; b++)
        for (int t = 0; t < 32 * RS_WARPS; t++)
            th.emplace_back([&, b, t] {
                threadIdx.x = (unsigned)t;
                blockIdx.x = (unsigned)b;
                blockDim.x = 32 * RS_WARPS;
                emu = EmuThread();
                emu.cta = ctas[b].get();
                emu.mode = mode;
                emu.rng ^= (uint64_t)(b * 131 + t + 1) * 0xD1B54A32D192ED03ull;
                if (a.stab)
                    k_resample_pair<true>(a.steps, a.offsets, a.order, a.n_hist, a.chunks, a.n_chunks, &counter, a.table_index, a.tables, a.P,
                                          a.out, zscratch, a.cap, a.ys, a.uniform_L, a.flags);
                else
                    k_resample_pair<false>(a.steps, a.offsets, a.order, a.n_hist, a.chunks, a.n_chunks, &counter, a.table_index, a.tables, a.P,
                                           a.out, zscratch, a.cap, a.ys, a.uniform_L, a.flags);
                copies += emu.copies;
                late += emu.late;
            });
    for (auto &t : th) t.join();
    g_ranges.pop_back();
    if (stats) { stats[0] += copies; stats[1] += late; stats[2] = g_oob; }
}

// factor tables of the given lengths (k_build_tables, one emulated thread per table)
void build_tables(const std::vector<uint32_t> &lens, uint32_t P, std::vector<double> &tables, std::vector<int64_t> &index)
{
    uint32_t max_len = 0;
    for (uint32_t L : lens) max_len = std::max(max_len, L);
    index.assign((size_t)max_len + 1, -1);
    std::vector<uint64_t> offs;
    uint64_t used = 0;
    for (uint32_t L : lens) { offs.push_back(used); index[L] = (int64_t)used; used += table_doubles(L, P); }
    tables.assign(used, std::nan(""));
    blockDim.x = 32;
    blockIdx.x = 0;
    for (size_t i = 0; i < lens.size(); i++) {
        threadIdx.x = 0;
        k_build_tables(lens.data() + i, offs.data() + i, 1, P, tables.data());
    }
}

}  // namespace

// Ragged batch: steps [sum L][6], offsets[n+1]; out [n][6P]. The plan is the library's (resample.cu build_plan):
// histories sorted by length, groups of five of one length, chunks of 16 groups, one launch per length class, longest
// first inside a class. force_global_table: run every class with the factor table read from global memory.
// stats[0] = cp.async copies issued, stats[1] = copies that landed only because a wait_group forced them,
// stats[2] = accesses outside the launch's buffers / the lane's own ring (must be 0).
extern "C" int k1_emul_ragged(const double *steps, const uint64_t *offsets, uint64_t n, uint32_t P, double *out, int mode, int n_ctas,
                              int force_global_table, uint32_t flags, uint64_t *stats)
{
    static const uint32_t caps[] = {64, SMEM_TAB_MAX_L, 2048, 16384, 131072};
    std::map<uint32_t, std::vector<uint32_t>> by_len;
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t L = offsets[i + 1] - offsets[i];
        if (L < 3 || L > 131072) return 1;
        by_len[(uint32_t)L].push_back((uint32_t)i);
    }
    std::vector<uint32_t> lens;
    for (auto &kv : by_len) lens.push_back(kv.first);
    std::vector<double> tables;
    std::vector<int64_t> index;
    build_tables(lens, P, tables, index);
    std::vector<uint32_t> order;
    std::map<uint32_t, std::pair<uint64_t, uint64_t>> groups_of;  // L -> [g0, g1)
    for (auto &kv : by_len) {
        const uint64_t g0 = order.size() / GROUP;
        for (uint32_t h : kv.second) order.push_back(h);
        while (order.size() % GROUP) order.push_back(0xffffffffu);
        groups_of[kv.first] = {g0, order.size() / GROUP};
    }
    g_ranges.clear();
    g_oob = 0;
    g_ranges.push_back(EmuRange{(const unsigned char *)steps, (const unsigned char *)(steps + offsets[n] * 6)});
    g_ranges.push_back(EmuRange{(const unsigned char *)tables.data(), (const unsigned char *)(tables.data() + tables.size())});
    uint32_t lo = 0;
    for (uint32_t cap_k : caps) {
        std::vector<K1Chunk> chunks;
        uint32_t max_len = 0;
        for (auto it = groups_of.rbegin(); it != groups_of.rend(); ++it) {
            const uint32_t L = it->first;
            if (L <= lo || L > cap_k) continue;
            max_len = std::max(max_len, L);
            for (uint64_t g = it->second.first; g < it->second.second; g += CHUNK_GROUPS)
                chunks.push_back(K1Chunk{(uint32_t)g, (uint32_t)std::min<uint64_t>(CHUNK_GROUPS, it->second.second - g), L, 0});
        }
        lo = cap_k;
        if (chunks.empty()) continue;
        const uint32_t cap = std::min(cap_k, max_len);
        Launch a{!force_global_table && cap <= SMEM_TAB_MAX_L, steps, offsets, order.data(), n, chunks.data(), (uint32_t)chunks.size(),
                 index.data(), tables.data(), P, out, cap, 6, 0, flags};
        run_launch(a, n_ctas, mode, stats);
    }
    return 0;
}

// History store: steps time-major [L][n][6], every history L steps long; out [n][6P].
extern "C" int k1_emul_store(const double *steps, uint64_t n, uint32_t L, uint32_t P, double *out, int mode, int n_ctas, uint32_t flags,
                             uint64_t *stats)
{
    if (L < 3) return 1;
    std::vector<double> tables;
    std::vector<int64_t> index;
    build_tables({L}, P, tables, index);
    g_ranges.clear();
    g_oob = 0;
    g_ranges.push_back(EmuRange{(const unsigned char *)steps, (const unsigned char *)(steps + (uint64_t)L * n * 6)});
    g_ranges.push_back(EmuRange{(const unsigned char *)tables.data(), (const unsigned char *)(tables.data() + tables.size())});
    const uint64_t n_groups = (n + GROUP - 1) / GROUP;
    Launch a{L <= SMEM_TAB_MAX_L, steps, nullptr, nullptr, n, nullptr, (uint32_t)((n_groups + CHUNK_GROUPS - 1) / CHUNK_GROUPS),
             index.data(), tables.data(), P, out, L, n * 6, L, flags};
    run_launch(a, n_ctas, mode, stats);
    return 0;
}
