// Host-side ends of the path: the per-history similarity files and the native graph reduction.
//
//  * write_similar_hist: Strain6D::most_similar_histories_to_file (reference
//    headers/strain2spline.h:301-314) for every history of the batch, as FE_problem.h:1232-1235 and
//    clustering/mpi_comparison_test.cc:99-103 call it.
//  * reduce_graph_calls / scema_reduce_dir: clustering/coarsegrain_dependency_network.py:24-94 —
//    greedy removal of the maximum-degree node with its neighbours; ties go to the node that
//    entered the networkx node dict last (stable sort + [-1], :20-21). The script re-sorts all
//    nodes every iteration (O(V^2 log V)); here the graph is a counting-sorted CSR adjacency and
//    the greedy choice runs connected component by connected component (a removal only changes
//    degrees inside its own component, so the mapping is the same): a scan per pick for the usual
//    small clusters, degree buckets of lazy max-heaps for large components.
#include "common.cuh"
#include <algorithm>
#include <dirent.h>
#include <queue>
#include <string>
#include <sys/types.h>
#include <unordered_map>

namespace scema {

int write_similar_hist(scema_ctx *c, const char *pattern)
{
    const uint64_t m = c->n_edges, n = c->n;
    // the pattern goes to snprintf as the format: exactly one conversion, and that one %u ("%%" is a literal percent sign)
    {
        int n_u = 0;
        for (const char *q = pattern; *q; q++) {
            if (*q != '%') continue;
            if (q[1] == '%') { q++; continue; }
            if (q[1] == 'u') { n_u++; q++; continue; }
            return fail(c, SCEMA_ERR_INVALID, "write_similar_hist: the file name pattern may only contain one %u");
        }
        if (n_u != 1) return fail(c, SCEMA_ERR_INVALID, "write_similar_hist: the file name pattern needs exactly one %u");
    }
    std::vector<uint32_t> a(m), b(m);
    std::vector<double> d(m);
    int rc = scema_get_edges(c, a.data(), b.data(), d.data(), m);
    if (rc) return rc;
    // history k lists partners in ascending batch index: those below k first (edges (i,k), which
    // are ascending in i for fixed k because the list is sorted by (a,b)), then those above k.
    // This is the order in which the single-rank loop strain2spline.h:603-611 appends them.
    std::vector<uint64_t> start(n + 2, 0);
    for (uint64_t e = 0; e < m; e++) { start[a[e] + 1]++; start[b[e] + 1]++; }
    for (uint64_t i = 0; i < n; i++) start[i + 1] += start[i];
    std::vector<uint64_t> fill(start.begin(), start.begin() + n);
    std::vector<uint32_t> other(2 * m);
    std::vector<double> dist(2 * m);
    for (uint64_t e = 0; e < m; e++) { uint64_t q = fill[b[e]]++; other[q] = a[e]; dist[q] = d[e]; }
    for (uint64_t e = 0; e < m; e++) { uint64_t q = fill[a[e]]++; other[q] = b[e]; dist[q] = d[e]; }
    const std::vector<uint32_t> &idv = ids_of(c);
    char name[4096];
    std::string buf;
    char line[96];
    for (uint64_t k = 0; k < n; k++) {
        snprintf(name, sizeof name, pattern, idv[k]);
        FILE *f = fopen(name, "w");
        if (!f) return fail(c, SCEMA_ERR_IO, std::string("Could not open ") + name + " for writing.");
        buf.clear();
        for (uint64_t q = start[k]; q < start[k + 1]; q++) {
            // default ostream formatting of a double == %g (6 significant digits)
            int len = snprintf(line, sizeof line, "%u %u %g\n", idv[k], idv[other[q]], dist[q]);
            buf.append(line, len);
        }
        if (!buf.empty() && fwrite(buf.data(), 1, buf.size(), f) != buf.size()) {
            fclose(f);
            return fail(c, SCEMA_ERR_IO, std::string("short write to ") + name);
        }
        fclose(f);
    }
    return SCEMA_OK;
}

// add_edge(eu[k], ev[k]) call sequence -> mapping. Returns 0, or SCEMA_ERR_MAPPING if an ID is
// >= num_gps (the script's IndexError).
int reduce_graph_calls(const uint32_t *eu, const uint32_t *ev, uint64_t n_calls, uint32_t num_gps, uint32_t *mapping,
                       uint64_t *iterations, uint64_t *neighbours_removed)
{
    for (uint32_t i = 0; i < num_gps; i++) mapping[i] = i;
    // node dict order = first appearance, cell1 before cell2
    std::vector<int64_t> pos(num_gps, -1);
    std::vector<uint32_t> order;
    for (uint64_t e = 0; e < n_calls; e++) {
        if (eu[e] >= num_gps || ev[e] >= num_gps) return SCEMA_ERR_MAPPING;
        if (pos[eu[e]] < 0) { pos[eu[e]] = (int64_t)order.size(); order.push_back(eu[e]); }
        if (pos[ev[e]] < 0) { pos[ev[e]] = (int64_t)order.size(); order.push_back(ev[e]); }
    }
    // deduplicated adjacency (both directions) in CSR form: counting sort by source node, then every list is compacted
    // in place with a last-seen marker (the similarity files list every edge from both endpoints, so most arcs arrive
    // twice). The order inside a list does not matter below.
    std::vector<uint64_t> start((size_t)num_gps + 1, 0);
    for (uint64_t e = 0; e < n_calls; e++) { start[(size_t)eu[e] + 1]++; start[(size_t)ev[e] + 1]++; }
    for (uint32_t i = 0; i < num_gps; i++) start[i + 1] += start[i];
    std::vector<uint32_t> arcs(2 * n_calls);
    {
        std::vector<uint64_t> fill(start.begin(), start.end() - 1);
        for (uint64_t e = 0; e < n_calls; e++) { arcs[fill[eu[e]]++] = ev[e]; arcs[fill[ev[e]]++] = eu[e]; }
    }
    std::vector<int64_t> deg(num_gps, 0);
    {
        std::vector<uint32_t> last_seen(num_gps, 0xffffffffu);  // node i marks its neighbours with i (IDs are < num_gps <= 2^32 - 1)
        uint64_t w = 0;
        for (uint32_t i = 0; i < num_gps; i++) {
            const uint64_t b0 = start[i], b1 = start[i + 1];
            start[i] = w;  // w <= b0: compaction in place
            bool self = false;
            for (uint64_t q = b0; q < b1; q++) {
                const uint32_t v = arcs[q];
                if (last_seen[v] == i) continue;
                last_seen[v] = i;
                arcs[w++] = v;
                self |= v == i;
            }
            deg[i] = (int64_t)(w - start[i]) + (self ? 1 : 0);  // a self loop counts twice in networkx
        }
        start[num_gps] = w;
    }
    // The script's choice = the alive node with the largest (degree, insertion position); it maps itself and its alive
    // neighbours to itself and removes them. A removal only changes degrees inside the connected component it happens in,
    // so the picks inside one component — and with them the mapping — are the same whether the components are worked on
    // interleaved (the script's global order) or one after the other. Component by component:
    //  * up to SMALL nodes (the usual case: one cluster of similar histories): every pick is a scan of the component's
    //    node list for the largest (degree, position) — no priority structure at all, everything cache-resident;
    //  * larger: one bucket per degree (degrees only fall, so the largest degree present never rises), each a max-heap
    //    of insertion positions with lazy deletion (a node enters the bucket of a degree exactly once, when its degree
    //    becomes that value; an entry is stale when the node is gone or its degree has moved on).
    constexpr size_t SMALL = 96;
    std::vector<uint8_t> alive(num_gps, 0), seen(num_gps, 0);
    for (uint32_t v : order) alive[v] = 1;
    uint64_t iters = 0, removed = 0;
    std::vector<uint32_t> comp, batch;
    std::vector<std::vector<uint32_t>> bucket;
    // removes `best` with its alive neighbours; on_drop(w) is called for every alive node whose degree fell
    auto take = [&](uint32_t best, auto &&on_drop) {
        mapping[best] = best;
        batch.clear();
        batch.push_back(best);
        for (uint64_t q = start[best]; q < start[best + 1]; q++) {
            const uint32_t w = arcs[q];
            if (!alive[w]) continue;
            mapping[w] = best;
            removed++;
            if (w != best) batch.push_back(w);
        }
        for (uint32_t rnode : batch) alive[rnode] = 0;
        for (uint32_t rnode : batch)
            for (uint64_t q = start[rnode]; q < start[rnode + 1]; q++) {
                const uint32_t w = arcs[q];
                if (alive[w]) { deg[w]--; on_drop(w); }
            }
        iters++;
        return batch.size();
    };
    for (uint32_t root : order) {
        if (seen[root]) continue;
        // the component of `root` (depth-first over the adjacency)
        comp.clear();
        comp.push_back(root);
        seen[root] = 1;
        for (size_t h = 0; h < comp.size(); h++)
            for (uint64_t q = start[comp[h]]; q < start[comp[h] + 1]; q++) {
                const uint32_t w = arcs[q];
                if (!seen[w]) { seen[w] = 1; comp.push_back(w); }
            }
        size_t remaining = comp.size();
        if (comp.size() <= SMALL) {
            while (remaining > 0) {
                uint32_t best = 0;
                int64_t bd = -1, bp = -1;
                for (uint32_t v : comp)
                    if (alive[v] && (deg[v] > bd || (deg[v] == bd && pos[v] > bp))) { best = v; bd = deg[v]; bp = pos[v]; }
                remaining -= take(best, [](uint32_t) {});
            }
            continue;
        }
        int64_t max_deg = 0;
        for (uint32_t v : comp) max_deg = std::max(max_deg, deg[v]);
        if (bucket.size() < (size_t)max_deg + 1) bucket.resize((size_t)max_deg + 1);
        for (int64_t d = 0; d <= max_deg; d++) bucket[d].clear();
        for (uint32_t v : comp) bucket[deg[v]].push_back((uint32_t)pos[v]);
        for (int64_t d = 0; d <= max_deg; d++) std::make_heap(bucket[d].begin(), bucket[d].end());
        int64_t d = max_deg;
        while (remaining > 0) {
            uint32_t best;
            while (true) {
                while (bucket[d].empty()) d--;  // some alive node of the component always sits in a bucket <= d
                std::pop_heap(bucket[d].begin(), bucket[d].end());
                best = order[bucket[d].back()];
                bucket[d].pop_back();
                if (alive[best] && deg[best] == d) break;
            }
            remaining -= take(best, [&](uint32_t w) {
                std::vector<uint32_t> &bk = bucket[deg[w]];
                bk.push_back((uint32_t)pos[w]);
                std::push_heap(bk.begin(), bk.end());
            });
        }
    }
    if (iterations) *iterations = iters;
    if (neighbours_removed) *neighbours_removed = removed;
    return SCEMA_OK;
}

}  // namespace scema

extern "C" int scema_reduce_calls(const uint32_t *cell1, const uint32_t *cell2, uint64_t n_calls, uint32_t num_gps,
                                  uint32_t *mapping_host, uint64_t *iterations, uint64_t *neighbours_removed)
{
    if ((n_calls && (!cell1 || !cell2)) || (num_gps && !mapping_host)) return SCEMA_ERR_INVALID;
    return scema::reduce_graph_calls(cell1, cell2, n_calls, num_gps, mapping_host, iterations, neighbours_removed);
}

extern "C" int scema_reduce_dir(const char *input_folder, const char *out_mapping_csv, uint32_t num_gps,
                                uint64_t *iterations, uint64_t *files_read, uint64_t *neighbours_removed)
{
    if (!input_folder || !out_mapping_csv) return SCEMA_ERR_INVALID;
    DIR *dirp = opendir(input_folder);
    if (!dirp) return SCEMA_ERR_IO;
    // glob(input_folder + "/last.*.similar_hist") (coarsegrain_dependency_network.py:48): directory
    // order, no sorting; fnmatch '*' may match the empty string and dots.
    std::vector<std::string> names;
    const std::string pre = "last.", suf = ".similar_hist";
    while (struct dirent *dp = readdir(dirp)) {
        std::string nm = dp->d_name;
        if (nm.size() >= pre.size() + suf.size() && nm.compare(0, pre.size(), pre) == 0 &&
            nm.compare(nm.size() - suf.size(), suf.size(), suf) == 0)
            names.push_back(nm);
    }
    closedir(dirp);
    std::vector<uint32_t> eu, ev;
    for (const std::string &nm : names) {
        std::string path = std::string(input_folder) + "/" + nm;
        FILE *f = fopen(path.c_str(), "r");
        if (!f) return SCEMA_ERR_IO;
        char line[512];
        while (fgets(line, sizeof line, f)) {
            long long c1, c2;
            double dist;
            char extra;
            // `cell1, cell2, dist = line.split()` needs exactly three tokens (:53)
            if (sscanf(line, "%lld %lld %lf %c", &c1, &c2, &dist, &extra) != 3) { fclose(f); return SCEMA_ERR_INVALID; }
            if (dist == 0.0) { fclose(f); return SCEMA_ERR_MAPPING; }  // 1.0/dist raises (:57)
            if (c1 < 0 || c2 < 0 || c1 >= (long long)num_gps || c2 >= (long long)num_gps) { fclose(f); return SCEMA_ERR_MAPPING; }
            eu.push_back((uint32_t)c1);
            ev.push_back((uint32_t)c2);
        }
        fclose(f);
    }
    std::vector<uint32_t> mapping(num_gps);
    uint64_t it = 0, nr = 0;
    int rc = scema::reduce_graph_calls(eu.data(), ev.data(), eu.size(), num_gps, mapping.data(), &it, &nr);
    if (rc) return rc;
    FILE *o = fopen(out_mapping_csv, "w");
    if (!o) return SCEMA_ERR_IO;
    for (uint32_t i = 0; i < num_gps; i++) fprintf(o, "%u %u\n", i, mapping[i]);
    fclose(o);
    if (iterations) *iterations = it;
    if (files_read) *files_read = names.size();
    if (neighbours_removed) *neighbours_removed = nr;
    return SCEMA_OK;
}
