// Static tile schedule of the tcgen05 filter (k_filter_tc, pairs_tc.cu). Plain C++ so that the host-side test
// tests/test_tc_schedule.py can enumerate it exhaustively (tests/helpers/tc_sched_check.cc).
#pragma once
#include <cstdint>
#ifdef __CUDACC__
#define SCEMA_TC_HD __host__ __device__ __forceinline__
#else
#define SCEMA_TC_HD inline
#endif

namespace scema {
namespace tc {

struct SchedArgs {
    uint32_t NT;        // column tiles (256 rows each) of the whole problem
    uint32_t I0, I1;    // row range of the launch, in 256-row tiles
    uint32_t C0, C1;    // column range of the launch, in 256-row tiles
    uint32_t strip_len;
    uint32_t shard, n_shards;
};

// Identical in every role of the kernel: work item = (row tile I of 128*CG rows, the column tiles of one strip
// at or right of the diagonal and inside the column range). Items are numbered strip by strip so that the
// clusters working at the same time walk the same strip of B (shared through L2); item g belongs to shard
// g % n_shards, and inside a shard unit u (CTA or CTA pair) takes the shard's items u, u + n_units, ...
template <int CG>
struct Sched {
    static constexpr uint32_t RPC = 2 / CG;  // row tiles per column tile
    uint32_t S, NT, C0, R0, R1, ns, s, shard, n_shards;  // NT: end of the column range
    uint64_t cum, k, stride;
    static SCEMA_TC_HD uint32_t mn(uint32_t a, uint32_t b) { return a < b ? a : b; }
    static SCEMA_TC_HD uint32_t mx(uint32_t a, uint32_t b) { return a > b ? a : b; }
    SCEMA_TC_HD void init(const SchedArgs &a, uint32_t unit, uint32_t n_units)
    {
        S = a.strip_len; NT = mn(a.NT, a.C1); C0 = a.C0; R0 = a.I0 * RPC; R1 = a.I1 * RPC;
        ns = C0 < NT ? (NT + S - 1) / S : 0u;  // an empty column range has no items
        s = mx(a.I0, a.C0) / S; cum = 0; k = unit; stride = n_units;
        shard = a.shard; n_shards = a.n_shards;
    }
    SCEMA_TC_HD uint32_t count(uint32_t strip) const
    {
        const uint32_t ce = mn((strip + 1) * S, NT) * RPC;
        const uint32_t e = mn(R1, ce);
        return e > R0 ? e - R0 : 0u;
    }
    SCEMA_TC_HD bool next(uint32_t &I, uint32_t &J0, uint32_t &J1)
    {
        const uint64_t g = k * n_shards + shard;
        while (s < ns && g >= cum + count(s)) { cum += count(s); s++; }
        if (s >= ns) return false;
        I = R0 + (uint32_t)(g - cum);
        J1 = mn((s + 1) * S, NT);
        J0 = mx(mx(I / RPC, s * S), C0);
        k += stride;
        return true;
    }
};

// Band schedule (rows sorted by norm, SCEMA_NORM_BAND): row tile I only meets the column tiles
// [I / RPC, jend[I]) — beyond that every pair is farther apart than the threshold by the triangle
// inequality | |a| - |b| | <= d(a, b). item_start[I] = number of items (strips of <= S tiles) of the
// row tiles before I; item g = (row tile found by bisection, strip g - item_start[I] of it).
template <int CG>
struct BandSched {
    static constexpr uint32_t RPC = 2 / CG;
    const uint32_t *item_start, *jend;
    uint32_t S, n_row_tiles, shard, n_shards;
    uint64_t k, stride;
    SCEMA_TC_HD void init(const uint32_t *item_start_, const uint32_t *jend_, uint32_t n_row_tiles_, uint32_t strip_len,
                          uint32_t shard_, uint32_t n_shards_, uint32_t unit, uint32_t n_units)
    {
        item_start = item_start_; jend = jend_; n_row_tiles = n_row_tiles_; S = strip_len;
        shard = shard_; n_shards = n_shards_; k = unit; stride = n_units;
    }
    SCEMA_TC_HD bool next(uint32_t &I, uint32_t &J0, uint32_t &J1)
    {
        const uint64_t g = k * n_shards + shard;
        if (g >= item_start[n_row_tiles]) return false;
        uint32_t lo = 0, hi = n_row_tiles;  // item_start[lo] <= g < item_start[hi]
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (item_start[mid] <= g) lo = mid; else hi = mid;
        }
        I = lo;
        J0 = I / RPC + (uint32_t)(g - item_start[I]) * S;
        J1 = J0 + S < jend[I] ? J0 + S : jend[I];
        k += stride;
        return true;
    }
};

}  // namespace tc
}  // namespace scema
