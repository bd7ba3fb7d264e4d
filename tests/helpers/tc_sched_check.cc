// Exhaustive host-side check of the tcgen05 filter's schedules (scema_b200/csrc/tc_sched.h; dense and norm-band): over all
// shards and units, every (row tile, column tile) of the launch's row and column range at or right of the diagonal
// is visited exactly once, nothing else is, and every item is non-empty. Prints "ok <cases>" or the first failure.
#include "../../scema_b200/csrc/tc_sched.h"
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace scema::tc;

static uint64_t rng_state = 88172645463325252ull;
static uint32_t rnd(uint32_t n) { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state % n); }

template <int CG>
static int check(const SchedArgs &a, uint32_t n_units)
{
    const uint32_t RPC = 2 / CG, rows = a.NT * RPC;
    std::vector<int> seen((size_t)rows * a.NT, 0);
    for (uint32_t shard = 0; shard < a.n_shards; shard++)
        for (uint32_t u = 0; u < n_units; u++) {
            SchedArgs b = a;
            b.shard = shard;
            Sched<CG> sc;
            sc.init(b, u, n_units);
            uint32_t I, J0, J1;
            while (sc.next(I, J0, J1)) {
                if (J0 >= J1 || I >= rows || J1 > a.NT) { printf("bad item I=%u J=[%u,%u)\n", I, J0, J1); return 1; }
                if (J1 - J0 > a.strip_len) { printf("item longer than a strip\n"); return 1; }
                for (uint32_t J = J0; J < J1; J++) seen[(size_t)I * a.NT + J]++;
            }
        }
    const uint32_t c1 = a.C1 < a.NT ? a.C1 : a.NT;
    for (uint32_t I = 0; I < rows; I++)
        for (uint32_t J = 0; J < a.NT; J++) {
            const bool want = I >= a.I0 * RPC && I < a.I1 * RPC && J >= a.C0 && J < c1 && J >= I / RPC;
            if (seen[(size_t)I * a.NT + J] != (want ? 1 : 0)) {
                printf("CG=%d NT=%u I=[%u,%u) C=[%u,%u) S=%u shards=%u units=%u: tile (%u,%u) visited %d times, want %d\n", CG, a.NT,
                       a.I0, a.I1, a.C0, a.C1, a.strip_len, a.n_shards, n_units, I, J, seen[(size_t)I * a.NT + J], want ? 1 : 0);
                return 1;
            }
        }
    return 0;
}

// BandSched: row tile I meets the column tiles [I / RPC, jend[I]); strips of S tiles, exclusive prefix in item_start
template <int CG>
static int check_band(uint32_t n_col, uint32_t S, uint32_t n_shards, uint32_t n_units)
{
    const uint32_t RPC = 2 / CG, rows = n_col * RPC;
    std::vector<uint32_t> jend(rows), item_start(rows + 1);
    uint32_t acc = 0;
    for (uint32_t I = 0; I < rows; I++) {
        jend[I] = I / RPC + 1 + rnd(n_col - I / RPC);  // at least the diagonal tile, at most everything to the right
        if (I && jend[I] < jend[I - 1]) jend[I] = jend[I - 1] > I / RPC ? jend[I - 1] : jend[I];  // bands of sorted norms never shrink
        item_start[I] = acc;
        acc += (jend[I] - I / RPC + S - 1) / S;
    }
    item_start[rows] = acc;
    std::vector<int> seen((size_t)rows * n_col, 0);
    for (uint32_t shard = 0; shard < n_shards; shard++)
        for (uint32_t u = 0; u < n_units; u++) {
            BandSched<CG> sc;
            sc.init(item_start.data(), jend.data(), rows, S, shard, n_shards, u, n_units);
            uint32_t I, J0, J1;
            while (sc.next(I, J0, J1)) {
                if (J0 >= J1 || I >= rows || J1 > n_col || J1 - J0 > S) { printf("band: bad item I=%u J=[%u,%u)\n", I, J0, J1); return 1; }
                for (uint32_t J = J0; J < J1; J++) seen[(size_t)I * n_col + J]++;
            }
        }
    for (uint32_t I = 0; I < rows; I++)
        for (uint32_t J = 0; J < n_col; J++) {
            const int want = J >= I / RPC && J < jend[I] ? 1 : 0;
            if (seen[(size_t)I * n_col + J] != want) {
                printf("band CG=%d cols=%u S=%u shards=%u units=%u: tile (%u,%u) visited %d times, want %d\n", CG, n_col, S, n_shards,
                       n_units, I, J, seen[(size_t)I * n_col + J], want);
                return 1;
            }
        }
    return 0;
}

int main(int argc, char **argv)
{
    const int cases = argc > 1 ? atoi(argv[1]) : 3000;
    for (int t = 0; t < cases; t++) {
        SchedArgs a;
        a.NT = 1 + rnd(t % 7 == 0 ? 90 : 24);
        a.I0 = rnd(a.NT + 1);
        a.I1 = a.I0 + rnd(a.NT - a.I0 + 2);          // may exceed NT by one: tc_launch clamps, so do we
        if (a.I1 > a.NT) a.I1 = a.NT;
        a.C0 = rnd(3) == 0 ? rnd(a.NT + 1) : 0;
        a.C1 = rnd(3) == 0 ? a.C0 + rnd(a.NT - a.C0 + 1) : 0xffffffffu;
        a.strip_len = 1 + rnd(t % 5 == 0 ? 40 : 6);
        a.n_shards = 1 + rnd(t % 3 == 0 ? 8 : 2);
        a.shard = 0;
        const uint32_t n_units = 1 + rnd(t % 4 == 0 ? 148 : 5);
        if (check<1>(a, n_units) || check<2>(a, n_units)) return 1;
        if (t % 4 == 0 && (check_band<1>(a.NT, a.strip_len, a.n_shards, n_units) || check_band<2>(a.NT, a.strip_len, a.n_shards, n_units))) return 1;
    }
    printf("ok %d\n", cases);
    return 0;
}
