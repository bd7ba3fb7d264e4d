// CPU check of b200::ring_sort (scema_b200/host/strain2spline_b200.h) against a literal simulation of the reference's
// ring loops (headers/strain2spline.h:571-611) on a random neighbour relation: for ring step i rank r receives the
// histories of rank r - i in their sender's order and appends a record to each of ITS histories that is a neighbour;
// step 0 is the local a < b double loop recording on both ends. No GPU, no library: only the header's host logic.
#include <cstdio>
#include <cstdlib>
#include <utility>
#include <vector>
#include <stdint.h>
#include "strain2spline_b200.h"

static uint64_t rng_state = 88172645463325252ull;
static uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

int main()
{
    int checked = 0;
    for (int R = 1; R <= 5; R++)
        for (int trial = 0; trial < 40; trial++) {
            // local vectors of uneven size; batch = rank-major concatenation
            std::vector<std::vector<uint32_t> > local(R);
            std::vector<int> ring_rank;
            uint32_t n = 0;
            for (int r = 0; r < R; r++) {
                const int cnt = (int)(rnd() % 7);
                for (int k = 0; k < cnt; k++) { local[r].push_back(n++); ring_rank.push_back(r); }
            }
            std::vector<std::vector<char> > nb(n, std::vector<char>(n, 0));
            for (uint32_t a = 0; a < n; a++)
                for (uint32_t b = a + 1; b < n; b++) nb[a][b] = nb[b][a] = (rnd() % 3 == 0);
            // literal ring
            std::vector<std::vector<uint32_t> > want(n);
            for (int r = 0; r < R; r++)
                for (int i = 0; i < R; i++) {
                    const int from = ((r - i) % R + R) % R;
                    if (from == r) {
                        for (size_t a = 0; a < local[r].size(); a++)
                            for (size_t b = a + 1; b < local[r].size(); b++)
                                if (nb[local[r][a]][local[r][b]]) { want[local[r][a]].push_back(local[r][b]); want[local[r][b]].push_back(local[r][a]); }
                    } else {
                        for (size_t s = 0; s < local[from].size(); s++)
                            for (size_t h = 0; h < local[r].size(); h++)
                                if (nb[local[r][h]][local[from][s]]) want[local[r][h]].push_back(local[from][s]);
                    }
                }
            for (uint32_t h = 0; h < n; h++) {
                std::vector<std::pair<uint32_t, double> > li;
                for (uint32_t j = 0; j < n; j++)
                    if (nb[h][j]) li.push_back(std::make_pair(j, 0.5 * j));
                MatHistPredict::b200::ring_sort(li, ring_rank[h], ring_rank, R);
                if (li.size() != want[h].size()) { printf("FAIL size R=%d h=%u\n", R, h); return 1; }
                for (size_t q = 0; q < li.size(); q++)
                    if (li[q].first != want[h][q] || li[q].second != 0.5 * want[h][q]) { printf("FAIL order R=%d h=%u q=%zu\n", R, h, q); return 1; }
                checked++;
            }
        }
    printf("ring order ok: %d histories\n", checked);
    return 0;
}
