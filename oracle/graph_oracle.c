/* TEST INFRASTRUCTURE ONLY — literal CPU restatement of the greedy graph reduction in
 * clustering/coarsegrain_dependency_network.py:24-94 (networkx semantics spelled out).
 * Pinned against the real script (run from oracle/_ref or /root/reference) by
 * tests/test_graph_reduce.py and by the fixtures under tests/golden/.
 *
 * Semantics restated (file:line relative to /root/reference/clustering/):
 *  - G.add_edge(cell1, cell2) for every line, in call order (coarsegrain_dependency_network.py:51-57);
 *    a new node enters the node dict when first seen, cell1 before cell2; repeated edges are
 *    no-ops for the structure. dist == 0 raises ZeroDivisionError (:57) -> we return 3.
 *  - mapping = identity over range(num_gps) (:60); an ID >= num_gps raises IndexError -> return 2.
 *  - every iteration (:66-85): degree dict in node-dict order (insertion order of the surviving
 *    nodes), sorted by degree with Python's stable sort, LAST element taken (:20-21): the
 *    maximum degree, ties to the node latest in dict order. It and all its neighbours map to it
 *    and are removed.
 * Deliberately O(iterations * V): this is the checker, not the product.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint32_t u, v; } arc;

static int arc_cmp(const void *pa, const void *pb)
{
    const arc *a = (const arc *)pa, *b = (const arc *)pb;
    if (a->u != b->u) return a->u < b->u ? -1 : 1;
    if (a->v != b->v) return a->v < b->v ? -1 : 1;
    return 0;
}

int oracle_reduce_graph(const uint32_t *eu, const uint32_t *ev, const double *edist, uint64_t n_calls,
                        uint32_t num_gps, uint32_t *mapping, uint64_t *iterations_out,
                        uint64_t *neighbours_removed_out)
{
    for (uint32_t i = 0; i < num_gps; i++) mapping[i] = i;
    /* node dict order */
    uint32_t *order = (uint32_t *)malloc((2 * n_calls + 1) * sizeof(uint32_t));
    uint8_t *seen = (uint8_t *)calloc((size_t)num_gps + 1, 1);
    uint64_t n_nodes = 0;
    int rc = 0;
    for (uint64_t e = 0; e < n_calls && !rc; e++) {
        if (edist && edist[e] == 0.0) rc = 3;
        /* note: the script inserts nodes before it can fail on mapping[]; the failure is fatal
         * either way, so the order of the two checks is immaterial */
        if (eu[e] >= num_gps || ev[e] >= num_gps) { rc = 2; break; }
        if (!seen[eu[e]]) { seen[eu[e]] = 1; order[n_nodes++] = eu[e]; }
        if (!seen[ev[e]]) { seen[ev[e]] = 1; order[n_nodes++] = ev[e]; }
    }
    if (rc) { free(order); free(seen); return rc; }

    /* adjacency (deduplicated, both directions) */
    arc *arcs = (arc *)malloc((2 * n_calls + 1) * sizeof(arc));
    uint64_t na = 0;
    for (uint64_t e = 0; e < n_calls; e++) {
        arcs[na].u = eu[e]; arcs[na++].v = ev[e];
        arcs[na].u = ev[e]; arcs[na++].v = eu[e];
    }
    qsort(arcs, na, sizeof(arc), arc_cmp);
    uint64_t nu = 0;
    for (uint64_t a = 0; a < na; a++)
        if (nu == 0 || arc_cmp(&arcs[a], &arcs[nu - 1]) != 0) arcs[nu++] = arcs[a];
    na = nu;
    uint64_t *start = (uint64_t *)calloc((size_t)num_gps + 2, sizeof(uint64_t));
    for (uint64_t a = 0; a < na; a++) start[arcs[a].u + 1]++;
    for (uint32_t i = 0; i < num_gps; i++) start[i + 1] += start[i];
    int64_t *deg = (int64_t *)calloc((size_t)num_gps + 1, sizeof(int64_t));
    for (uint32_t i = 0; i < num_gps; i++) {
        deg[i] = (int64_t)(start[i + 1] - start[i]);
        /* a self loop counts twice in networkx degree */
        for (uint64_t a = start[i]; a < start[i + 1]; a++) if (arcs[a].v == i) deg[i]++;
    }
    uint8_t *alive = seen; /* reuse: 1 = in graph */
    uint64_t remaining = n_nodes, iterations = 0, neighbours_removed = 0;
    uint32_t *batch = (uint32_t *)malloc((n_nodes + 1) * sizeof(uint32_t));
    while (remaining > 0) {
        int64_t best_deg = -1; uint32_t best = 0;
        for (uint64_t q = 0; q < n_nodes; q++) {
            uint32_t v = order[q];
            if (alive[v] && deg[v] >= best_deg) { best_deg = deg[v]; best = v; }
        }
        mapping[best] = best;
        uint64_t nb = 0;
        batch[nb++] = best;
        for (uint64_t a = start[best]; a < start[best + 1]; a++) {
            uint32_t w = arcs[a].v;
            if (!alive[w]) continue;
            /* nx.all_neighbors yields a self-looped node as its own neighbour */
            mapping[w] = best;
            neighbours_removed++;
            if (w != best) batch[nb++] = w;
        }
        for (uint64_t q = 0; q < nb; q++) alive[batch[q]] = 0;
        for (uint64_t q = 0; q < nb; q++) {
            uint32_t r = batch[q];
            for (uint64_t a = start[r]; a < start[r + 1]; a++)
                if (alive[arcs[a].v]) deg[arcs[a].v]--;
        }
        remaining -= nb;
        iterations++;
    }
    if (iterations_out) *iterations_out = iterations;
    if (neighbours_removed_out) *neighbours_removed_out = neighbours_removed;
    free(order); free(seen); free(arcs); free(start); free(deg); free(batch);
    return 0;
}
