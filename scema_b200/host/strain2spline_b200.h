// strain2spline_b200.h — drop-in for SCEMa's headers/strain2spline.h on top of libscema_hist.so.
//
// Same namespace, class, method and free-function names, argument meaning and error behaviour
// (message on stderr/stdout + exit(1)) as the reference header, so FE_problem.h
// (spline_building :1167-1191, spline_comparison :1197-1270), dealammps.cc scheduling and the
// two clustering command lines compile against it unchanged:
//     #include "strain2spline_b200.h"      // instead of "strain2spline.h"
// What is different is WHERE the work happens. The reference fits six tk::spline objects per
// history on the CPU and walks the N^2 pair loop per rank (strain2spline.h:140-180, :546-614);
// here a Strain6D only stores its samples, splinify() is deferred, and
// compare_histories_with_all_ranks() ships the whole batch through the C ABI
// (include/scema_hist.h): K1 resample, K2 all-pairs, K3 compaction on the GPU. Results land in
// the same per-object lists, in the reference's order, with bit-identical doubles.
//
// Deliberate differences from strain2spline.h (everything else is byte-compatible, tests/test_dropin_header.py):
//   * all_similar_histories (the O(N^2) "theory-checking only" lists behind all_similar_histories_to_file,
//     strain2spline.h:271,316-330, written at FE_problem.h:1237-1238) stay EMPTY unless SCEMA_B200_ALL_SIMILAR=1
//     (then a dense exact pass fills them, in the reference's order, single- and multi-rank);
//   * get_most_similar_history_ID / _diff (the legacy nearest neighbour, :277-289) then report the nearest history
//     among those BELOW the threshold ({UINT32_MAX, inf} if none) instead of the global nearest; SCEMA_B200_NEAREST=1
//     (or SCEMA_B200_ALL_SIMILAR=1) restores the reference's global nearest neighbour with its lowest-ID tie-break
//     (scema_nearest: exact distances to all other histories, no O(N^2) storage). Nothing in SCEMa reads either.
//   * Strain6DReceiver / send_strain6D_mpi / receive_strain6D_mpi / modulo_neg exist for name compatibility; the
//     collective below does not use them (no ring).
//
// MPI: when <mpi.h> has been included before this header (MPI_VERSION defined) the collective
// gathers every rank's histories to rank 0 of `comm`, which owns the GPU, and scatters the
// per-history results back in the reference's ring order (local partners first, then those of
// rank-1, rank-2, ... — strain2spline.h:571-599). Define SCEMA_B200_NO_MPI to build without MPI
// (single rank; MPI_Comm becomes a placeholder type).
#ifndef MATHISTPREDICT_STRAIN2SPLINE_B200_H
#define MATHISTPREDICT_STRAIN2SPLINE_B200_H

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <fstream>
#include <iostream>
#include <limits>
#include <string>
#include <utility>
#include <vector>
#include <stdint.h>

#include "scema_hist.h"

#if !defined(MPI_VERSION)
#if !defined(SCEMA_B200_NO_MPI)
#define SCEMA_B200_NO_MPI 1
#endif
typedef int MPI_Comm;
#ifndef MPI_COMM_WORLD
#define MPI_COMM_WORLD 0
#endif
#endif

namespace MatHistPredict {

typedef struct {
    uint32_t ID;
    double diff;
} HISTORY_ID_DIFF_PAIR;

class Strain6D;

namespace b200 {

inline void die(const std::string &msg)
{
    fprintf(stderr, "%s\n", msg.c_str());
    exit(1);
}

// The GPUs of this process: $SCEMA_B200_DEVICES="0,1,2,3" shards every comparison over those GPUs (scema_multi_*: one
// host thread per GPU and NCCL inside the library; the caller stays single-threaded), otherwise one GPU
// ($SCEMA_B200_DEVICE, default 0). Returns NULL when a single device is configured.
inline scema_multi *multi()
{
    static scema_multi *m = NULL;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char *e = getenv("SCEMA_B200_DEVICES");
        std::vector<int> devs;
        if (e)
            for (const char *p = e; *p;) {
                char *end = NULL;
                long v = strtol(p, &end, 10);
                if (end == p) break;
                devs.push_back((int)v);
                p = (*end == ',') ? end + 1 : end;
            }
        if (devs.size() > 1) {
            int rc = scema_multi_create(&m, devs.data(), (int)devs.size());
            if (rc != SCEMA_OK) die("scema_b200: cannot create contexts on the GPUs listed in SCEMA_B200_DEVICES");
        }
    }
    return m;
}

// The context that holds the results (the only one, or the first of the device set).
inline scema_ctx *context()
{
    static scema_ctx *ctx = NULL;
    if (!ctx) {
        if (multi()) {
            ctx = scema_multi_context(multi(), 0);
        } else {
            const char *d = getenv("SCEMA_B200_DEVICE");
            int rc = scema_create(&ctx, d ? atoi(d) : 0, NULL);
            if (rc != SCEMA_OK) die("scema_b200: cannot create a GPU context (no CUDA device? there is no CPU fallback)");
        }
    }
    return ctx;
}

inline void check(int rc, const char *what)
{
    if (rc != SCEMA_OK) die(std::string(what) + ": " + scema_last_error(context()));
}

inline void check_multi(int rc, const char *what)
{
    if (rc != SCEMA_OK) die(std::string(what) + ": " + scema_multi_last_error(multi()));
}

// $SCEMA_B200_ALL_SIMILAR=1 also fills the reference's "theory-checking only" full comparison
// lists (all_similar_histories, strain2spline.h:271,431-432) and the legacy nearest neighbour
// (:277-289) from a dense pass; off by default because it is O(N^2) output.
inline bool keep_all_similar()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("SCEMA_B200_ALL_SIMILAR"); v = (e && atoi(e) != 0) ? 1 : 0; }
    return v == 1;
}

// $SCEMA_B200_NEAREST=1: the legacy global nearest neighbour without the O(N^2) lists (scema_nearest).
inline bool keep_nearest()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("SCEMA_B200_NEAREST"); v = (e && atoi(e) != 0) ? 1 : 0; }
    return v == 1;
}

void resolve_splines(std::vector<Strain6D *> &pending);

// ---- device-resident history store behind the in-process call pattern. FE_problem.h appends one sample to EVERY
// quadrature point per timestep (:1091-1103), calls splinify() on every point (spline_building, :1167-1191) and compares
// the flagged subset (:1202-1229). The objects that asked for a (deferred) splinify() since the last comparison are
// remembered here; when they all have the same length and spline point count — the production case — their histories are
// kept on the GPU (scema_store_*, time-major) and a timestep only uploads the samples added since the last comparison
// (48 bytes per point and step) instead of every history again. The fit of all points and the selection of the flagged
// rows then happen on the device (scema_store_resample, scema_select_rows). Anything irregular (different lengths, a
// history replaced by from_file, points created or destroyed) rebuilds the store once or falls back to the flat upload.
// SCEMA_B200_STORE=0 switches it off.
struct StoreState {
    std::vector<Strain6D *> registry;     // objects with a pending splinify(), in call order (NULL: destroyed since)
    std::vector<uint32_t> ids;            // IDs of the store's rows
    std::vector<uint64_t> epochs;         // data epoch of every row's object when it was stored
    uint32_t steps;                       // samples per point in the store
    bool valid;
    uint64_t rebuilds, appended_steps, h2d_bytes, flat_uploads;
    StoreState() : steps(0), valid(false), rebuilds(0), appended_steps(0), h2d_bytes(0), flat_uploads(0) {}
};
// per thread: SCEMa drives this header from one thread per process; a test that runs several "ranks" as threads of one
// process (tests/helpers/mpi_threads) must not share the registry between them
inline StoreState &store_state() { static thread_local StoreState st; return st; }
inline bool store_enabled()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("SCEMA_B200_STORE"); v = (e && atoi(e) == 0) ? 0 : 1; }
    return v == 1;
}

// Partner order of the reference's ring (strain2spline.h:571-599). `li` lists the partners of one history of rank
// `my_rank` by batch index, ascending, where the batch is the rank-major concatenation of every rank's local vector
// and ring_rank[j] is the rank history j lives on. The ring visits the own rank first (the local a < b loop: partners
// in ascending local index), then the histories received from rank-1, rank-2, ... each in its sender's order — a
// stable pass by ring step keeps exactly that. Pure host logic (tests/helpers/ring_order_check.cc).
inline void ring_sort(std::vector<std::pair<uint32_t, double> > &li, int my_rank, const std::vector<int> &ring_rank, int n_ranks)
{
    if (n_ranks <= 1) return;
    std::vector<std::pair<uint32_t, double> > sorted;
    sorted.reserve(li.size());
    for (int step = 0; step < n_ranks; step++)
        for (size_t q = 0; q < li.size(); q++)
            if (((my_rank - ring_rank[li[q].first]) % n_ranks + n_ranks) % n_ranks == step) sorted.push_back(li[q]);
    li.swap(sorted);
}

}  // namespace b200

class Strain6D {
public:
    Strain6D()
    {
        up_to_date = false;
        pending = false;
        num_steps_added = 0;
        num_spline_points_per_component = 0;
        ID = std::numeric_limits<uint32_t>::max();
        ID_is_set = false;
        most_similar_history.ID = 0;
        most_similar_history.diff = 0;
        ID_to_get_results_from = std::numeric_limits<uint32_t>::max();
        most_recent_ID_to_get_results_from = std::numeric_limits<uint32_t>::max();
        for (int i = 0; i < 6; i++) stress[i] = 0.0;
        reg_slot = -1;
        data_epoch = next_epoch();
    }
    // a copy is a new object as far as the registry of pending fits is concerned (PointHistory holds a Strain6D by value)
    Strain6D(const Strain6D &o) { reg_slot = -1; copy_from(o); }
    Strain6D &operator=(const Strain6D &o)
    {
        if (this != &o) { unregister(); copy_from(o); }
        return *this;
    }
    ~Strain6D() { unregister(); }

    void set_ID(uint32_t id) { ID = id; ID_is_set = true; }

    void add_current_strain(double xx, double yy, double zz, double xy, double xz, double yz)
    {
        up_to_date = false;
        pending = false;
        const double s[6] = {xx, yy, zz, xy, xz, yz};
        steps.insert(steps.end(), s, s + 6);
        num_steps_added++;
    }

    void add_current_strain(double xx, double yy, double zz, double xy, double xz, double yz, double sxx, double syy,
                            double szz, double sxy, double sxz, double syz)
    {
        add_current_strain(xx, yy, zz, xy, xz, yz);
        stress[0] = sxx; stress[1] = syy; stress[2] = szz; stress[3] = sxy; stress[4] = sxz; stress[5] = syz;
    }

    // One line per step: xx yy zz xy xz yz; reading stops at the first token that does not parse.
    void from_file(const char *in_fname)
    {
        up_to_date = false;
        pending = false;
        std::ifstream in(in_fname);
        if (in.fail()) {
            fprintf(stderr, "Could not open %s for reading.\n", in_fname);
            exit(1);
        }
        data_epoch = next_epoch();  // not an append of one sample per timestep: a device copy of this history is stale
        double s[6];
        while (in >> s[0] >> s[1] >> s[2] >> s[3] >> s[4] >> s[5]) {
            steps.insert(steps.end(), s, s + 6);
            num_steps_added++;
        }
    }

    // Deferred: validates like the reference, remembers the request; the fit itself runs on the GPU
    // for the whole batch in compare_histories_with_all_ranks / splinify_batch, or for this object
    // alone the first time get_spline() needs it.
    void splinify(uint32_t n_points)
    {
        if (num_steps_added == 0) {
            fprintf(stderr, "Error: Nothing to splinify! No strain data has been read in yet. Please use .from_file() or .add_current_strain() first.\n");
            exit(1);
        } else if (num_steps_added < 3) {
            fprintf(stderr, "Error: Not enough strain steps added. Need at least 3 points for splinify().\n");
            exit(1);
        }
        num_spline_points_per_component = n_points;
        spline.clear();
        pending = true;
        up_to_date = true;
        if (reg_slot < 0 && b200::store_enabled()) {  // remembered until the next comparison (b200::StoreState)
            b200::StoreState &st = b200::store_state();
            reg_slot = (long)st.registry.size();
            st.registry.push_back(this);
        }
    }

    void print()
    {
        if (!up_to_date) std::cout << "Warning: spline is not up to date (run splinify() to rebuild spline)\n";
        materialise();
        for (uint32_t n = 0; n + 5 < spline.size(); n += 6)
            std::cout << spline[n] << ' ' << spline[n + 1] << ' ' << spline[n + 2] << ' ' << spline[n + 3] << ' '
                      << spline[n + 4] << ' ' << spline[n + 5] << '\n';
    }

    // The reference writes the first component six times per line (strain2spline.h:207); kept.
    void spline_to_file(char *out_fname)
    {
        if (!up_to_date) std::cout << "Warning: spline is not up to date (run splinify() to rebuild spline)\n";
        materialise();
        std::ofstream out(out_fname);
        if (out.fail()) {
            fprintf(stderr, "Could not open %s for writing.\n", out_fname);
            exit(1);
        }
        for (uint32_t n = 0; n + 5 < spline.size(); n += 6)
            out << spline[n] << ' ' << spline[n] << ' ' << spline[n] << ' ' << spline[n] << ' ' << spline[n] << ' '
                << spline[n] << '\n';
    }

    std::vector<double> *get_spline()
    {
        if (!up_to_date) {
            std::cout << "Spline is not up to date.\n";
            exit(1);
        }
        materialise();
        return &spline;
    }

    uint32_t get_ID()
    {
        if (!ID_is_set) {
            fprintf(stderr, "Error: history ID is unset. Please use set_ID().\n");
            exit(1);
        }
        return ID;
    }

    uint32_t get_num_spline_points_per_component() { return num_spline_points_per_component; }
    uint32_t get_most_similar_history_ID() { return most_similar_history.ID; }
    double get_most_similar_history_diff() { return most_similar_history.diff; }

    void clear_most_similar_history()
    {
        most_similar_history.ID = std::numeric_limits<uint32_t>::max();
        most_similar_history.diff = std::numeric_limits<double>::infinity();
        most_similar_histories.clear();
        all_similar_histories.clear();
    }

    void choose_most_similar_history(double candidate_diff, uint32_t candidate_ID, double threshold)
    {
        HISTORY_ID_DIFF_PAIR hp;
        hp.diff = candidate_diff;
        hp.ID = candidate_ID;
        all_similar_histories.push_back(hp);
        if (candidate_diff < threshold) most_similar_histories.push_back(hp);
        note_nearest(candidate_diff, candidate_ID);
    }

    void print_most_similar_histories()
    {
        for (size_t i = 0; i < most_similar_histories.size(); i++)
            std::cout << ID << " " << most_similar_histories[i].ID << " " << most_similar_histories[i].diff << "\n";
    }

    void most_similar_histories_to_file(const char *out_fname) { dump(most_similar_histories, out_fname); }
    void all_similar_histories_to_file(const char *out_fname) { dump(all_similar_histories, out_fname); }

    bool run_new_md() { return ID_to_get_results_from == ID; }

    // mapping.csv of coarsegrain_dependency_network.py: line ID holds "<ID> <source ID>".
    void read_coarsegrain_dependency_mapping(const char *in_fname)
    {
        std::ifstream in(in_fname);
        if (in.fail()) {
            fprintf(stderr, "Could not open %s for reading.\n", in_fname);
            exit(1);
        }
        std::string skip;
        for (uint32_t i = 0; i < ID; i++) std::getline(in, skip);
        uint32_t id_from = 0, id_to = 0;
        in >> id_from >> id_to;
        if (id_from != ID) {
            fprintf(stderr, "ID in mapping file (%u) does not match cell ID (%u)\n", id_from, ID);
            exit(1);
        }
        ID_to_get_results_from = id_to;
    }

    void set_ID_to_get_results_from(uint32_t id) { ID_to_get_results_from = id; }
    void set_most_recent_ID_to_get_results_from(uint32_t id) { most_recent_ID_to_get_results_from = id; }
    uint32_t get_ID_to_get_results_from() { return ID_to_get_results_from; }
    uint32_t get_most_recent_ID_to_get_results_from() { return most_recent_ID_to_get_results_from; }

    // ---- batch plumbing used by this header's free functions (not part of the reference API)
    uint64_t b200_epoch() const { return data_epoch; }
    void b200_left_registry() { reg_slot = -1; }
    bool b200_pending() const { return pending; }
    bool b200_up_to_date() const { return up_to_date; }
    uint32_t b200_num_steps() const { return num_steps_added; }
    const std::vector<double> &b200_steps() const { return steps; }
    void b200_install_spline(const double *row, uint32_t k) { spline.assign(row, row + k); pending = false; }
    void b200_push_partner(uint32_t other_id, double diff, bool below_threshold)
    {
        HISTORY_ID_DIFF_PAIR hp;
        hp.ID = other_id;
        hp.diff = diff;
        if (below_threshold) most_similar_histories.push_back(hp);
    }
    void b200_push_all(uint32_t other_id, double diff)
    {
        HISTORY_ID_DIFF_PAIR hp;
        hp.ID = other_id;
        hp.diff = diff;
        all_similar_histories.push_back(hp);
        note_nearest(diff, other_id);
    }
    void b200_note_nearest(double diff, uint32_t other_id) { note_nearest(diff, other_id); }
    void b200_set_nearest(uint32_t other_id, double diff) { most_similar_history.ID = other_id; most_similar_history.diff = diff; }
    void b200_push_all_raw(uint32_t other_id, double diff)
    {
        HISTORY_ID_DIFF_PAIR hp;
        hp.ID = other_id;
        hp.diff = diff;
        all_similar_histories.push_back(hp);
    }
    const std::vector<HISTORY_ID_DIFF_PAIR> &b200_most_similar() const { return most_similar_histories; }
    const std::vector<HISTORY_ID_DIFF_PAIR> &b200_all_similar() const { return all_similar_histories; }

private:
    static uint64_t next_epoch() { static uint64_t e = 0; return ++e; }
    void unregister()
    {
        if (reg_slot >= 0) {
            b200::StoreState &st = b200::store_state();
            if ((size_t)reg_slot < st.registry.size() && st.registry[reg_slot] == this) st.registry[reg_slot] = NULL;
            reg_slot = -1;
        }
    }
    void copy_from(const Strain6D &o)
    {
        up_to_date = o.up_to_date; pending = o.pending; num_steps_added = o.num_steps_added; steps = o.steps;
        for (int i = 0; i < 6; i++) stress[i] = o.stress[i];
        ID = o.ID; ID_is_set = o.ID_is_set; num_spline_points_per_component = o.num_spline_points_per_component;
        spline = o.spline; most_similar_history = o.most_similar_history; most_similar_histories = o.most_similar_histories;
        all_similar_histories = o.all_similar_histories; ID_to_get_results_from = o.ID_to_get_results_from;
        most_recent_ID_to_get_results_from = o.most_recent_ID_to_get_results_from;
        data_epoch = next_epoch();
    }
    void materialise()
    {
        if (!pending) return;
        std::vector<Strain6D *> one(1, this);
        b200::resolve_splines(one);
    }

    // legacy single nearest neighbour, lowest ID wins ties (strain2spline.h:277-289)
    void note_nearest(double candidate_diff, uint32_t candidate_ID)
    {
        if (candidate_diff <= most_similar_history.diff) {
            if (candidate_diff == most_similar_history.diff && candidate_ID > most_similar_history.ID) return;
            most_similar_history.ID = candidate_ID;
            most_similar_history.diff = candidate_diff;
        }
    }

    void dump(const std::vector<HISTORY_ID_DIFF_PAIR> &list, const char *out_fname)
    {
        std::ofstream out(out_fname);
        if (out.fail()) {
            fprintf(stderr, "Could not open %s for writing.\n", out_fname);
            exit(1);
        }
        for (size_t i = 0; i < list.size(); i++) out << ID << " " << list[i].ID << " " << list[i].diff << "\n";
    }

    long reg_slot;         // position in b200::StoreState::registry while a deferred splinify() is pending there, else -1
    uint64_t data_epoch;   // changes whenever the history is anything but appended to
    bool up_to_date, pending;
    uint32_t num_steps_added;
    std::vector<double> steps;  // [num_steps_added][6]: xx yy zz xy xz yz
    double stress[6];
    uint32_t ID;
    bool ID_is_set;
    uint32_t num_spline_points_per_component;
    std::vector<double> spline;  // [P][6], filled when the deferred fit has run
    HISTORY_ID_DIFF_PAIR most_similar_history;
    std::vector<HISTORY_ID_DIFF_PAIR> most_similar_histories;
    std::vector<HISTORY_ID_DIFF_PAIR> all_similar_histories;
    uint32_t ID_to_get_results_from;
    uint32_t most_recent_ID_to_get_results_from;
};

namespace b200 {

// Run the deferred fits of `pending` (grouped by spline point count) through K1 and install the rows.
inline void resolve_splines(std::vector<Strain6D *> &pending)
{
    std::vector<uint32_t> point_counts;
    for (size_t i = 0; i < pending.size(); i++) {
        if (!pending[i]->b200_pending()) continue;
        uint32_t P = pending[i]->get_num_spline_points_per_component();
        bool seen = false;
        for (size_t q = 0; q < point_counts.size(); q++) seen |= point_counts[q] == P;
        if (!seen) point_counts.push_back(P);
    }
    scema_ctx *ctx = context();
    for (size_t q = 0; q < point_counts.size(); q++) {
        const uint32_t P = point_counts[q];
        std::vector<Strain6D *> grp;
        std::vector<uint64_t> offsets(1, 0);
        std::vector<double> flat;
        for (size_t i = 0; i < pending.size(); i++) {
            Strain6D *h = pending[i];
            if (!h->b200_pending() || h->get_num_spline_points_per_component() != P) continue;
            grp.push_back(h);
            flat.insert(flat.end(), h->b200_steps().begin(), h->b200_steps().end());
            offsets.push_back(offsets.back() + h->b200_num_steps());
        }
        if (P == 0) {  // the reference yields an empty spline vector
            for (size_t i = 0; i < grp.size(); i++) grp[i]->b200_install_spline(NULL, 0);
            continue;
        }
        check(scema_set_histories(ctx, flat.data(), 0, offsets.data(), NULL, grp.size()), "splinify");
        check(scema_resample(ctx, P), "splinify");
        std::vector<double> rows((size_t)grp.size() * 6 * P);
        check(scema_get_spline(ctx, rows.data()), "splinify");
        for (size_t i = 0; i < grp.size(); i++) grp[i]->b200_install_spline(rows.data() + i * 6 * (size_t)P, 6 * P);
    }
}

// The store path of compare_batch (see StoreState). True when the comparison ran from the device-resident store; false
// when this batch does not fit the pattern (the caller then uploads the flat batch as before).
inline bool store_compare(std::vector<Strain6D *> &hist, const std::vector<uint32_t> &ids, uint32_t P0, double threshold)
{
    if (!store_enabled()) return false;
    StoreState &st = store_state();
    // the points that asked for a fit since the last comparison, still alive and still waiting for it
    std::vector<Strain6D *> all;
    all.reserve(st.registry.size());
    for (size_t i = 0; i < st.registry.size(); i++) {
        Strain6D *h = st.registry[i];
        if (!h) continue;
        h->b200_left_registry();
        if (h->b200_pending()) all.push_back(h);
    }
    st.registry.clear();
    const size_t N = all.size();
    // the in-process pattern is "fit every point, compare the flagged subset"; a one-off batch in which every fitted point
    // is also compared (the command lines, the proxies of the multi-rank branch) is cheaper as one flat upload
    if (N <= hist.size() || N == 0) return false;
    const uint32_t L = all[0]->b200_num_steps();
    if (L < 3) return false;
    for (size_t i = 0; i < N; i++)
        if (all[i]->b200_num_steps() != L || all[i]->get_num_spline_points_per_component() != P0) return false;
    // rows of the flagged points: position in `all` by address (both lists are in the caller's order in production)
    std::vector<uint32_t> rows(hist.size());
    {
        size_t cursor = 0;
        for (size_t i = 0; i < hist.size(); i++) {
            size_t tries = 0;
            while (tries < N && all[cursor] != hist[i]) { cursor = cursor + 1 == N ? 0 : cursor + 1; tries++; }
            if (all[cursor] != hist[i]) return false;   // a flagged point that never asked for a fit through splinify()
            rows[i] = (uint32_t)cursor;
        }
    }
    scema_ctx *ctx = context();
    bool same = st.valid && st.ids.size() == N && L >= st.steps;
    for (size_t i = 0; same && i < N; i++) same = st.ids[i] == all[i]->get_ID() && st.epochs[i] == all[i]->b200_epoch();
    if ((same ? L - st.steps : L) > 4096) return false;   // that many single-step uploads: the flat copy is the better catch-up
    if (!same) {
        st.ids.resize(N);
        st.epochs.resize(N);
        for (size_t i = 0; i < N; i++) { st.ids[i] = all[i]->get_ID(); st.epochs[i] = all[i]->b200_epoch(); }
        check(scema_store_reset(ctx, N, st.ids.data(), L + 64), "add_current_strain");
        st.steps = 0;
        st.valid = true;
        st.rebuilds++;
    }
    std::vector<double> sample(N * 6);
    for (uint32_t s = st.steps; s < L; s++) {   // one 48 N-byte upload per new timestep
        for (size_t i = 0; i < N; i++) {
            const double *src = all[i]->b200_steps().data() + (size_t)s * 6;
            for (int c = 0; c < 6; c++) sample[i * 6 + c] = src[c];
        }
        check(scema_store_append(ctx, sample.data(), 0), "add_current_strain");
        st.appended_steps++;
        st.h2d_bytes += N * 48;
    }
    st.steps = L;
    check(scema_store_resample(ctx, P0), "splinify");
    check(scema_select_rows(ctx, rows.data(), rows.size()), "compare_histories_with_all_ranks");
    uint64_t m = 0;
    check(scema_compare(ctx, threshold, SCEMA_PAIRS_TC, 0, 1, &m), "compare_histories_with_all_ranks");
    (void)ids;
    return true;
}

// All-pairs on the GPU for one (already rank-merged) batch; fills every object's result lists.
// ring_rank[i] / n_ranks describe which MPI rank history i lives on, to reproduce the partner
// order of the reference's ring when more than one rank takes part.
inline void compare_batch(std::vector<Strain6D *> &hist, double threshold, const std::vector<int> &ring_rank, int n_ranks)
{
    const size_t n = hist.size();
    for (size_t i = 0; i < n; i++) hist[i]->clear_most_similar_history();
    if (n == 0) return;
    scema_ctx *ctx = context();
    std::vector<uint32_t> ids(n);
    for (size_t i = 0; i < n; i++) {
        if (!hist[i]->b200_up_to_date()) {  // get_spline() of a stale object (strain2spline.h:214-221)
            std::cout << "Spline is not up to date.\n";
            exit(1);
        }
        ids[i] = hist[i]->get_ID();
    }
    // Usual case (FE_problem.h:1187 then :1229, mpi_comparison_test.cc:83 then :96): every history still
    // carries its deferred splinify(P) with the same P. The raw histories then go to the GPU once,
    // K1 writes the spline matrix where K2 reads it, and nothing comes back but the edges; an object's
    // own spline vector is only materialised if somebody asks for it (get_spline, print, ...).
    bool direct = true, clustered = false;
    const uint32_t P0 = hist[0]->get_num_spline_points_per_component();
    for (size_t i = 0; i < n; i++) direct = direct && hist[i]->b200_pending() && hist[i]->get_num_spline_points_per_component() == P0;
    if (direct && P0 > 0 && !keep_all_similar() && !multi() && store_compare(hist, ids, P0, threshold)) {
        clustered = true;   // histories kept on the device, only the new samples uploaded
    } else if (direct && P0 > 0) {
        store_state().flat_uploads++;
        std::vector<uint64_t> offsets(1, 0);
        std::vector<double> flat;
        for (size_t i = 0; i < n; i++) {
            flat.insert(flat.end(), hist[i]->b200_steps().begin(), hist[i]->b200_steps().end());
            offsets.push_back(offsets.back() + hist[i]->b200_num_steps());
        }
        // one call: for large batches the library pipelines copy, resample and compare range by range (scema_cluster)
        const bool dense0 = keep_all_similar();
        uint64_t m0 = 0;
        const double thr0 = dense0 ? std::numeric_limits<double>::infinity() : threshold;
        const int var0 = dense0 ? SCEMA_PAIRS_EXACT : SCEMA_PAIRS_TC;
        if (multi())
            check_multi(scema_multi_cluster(multi(), flat.data(), offsets.data(), ids.data(), n, P0, thr0, var0, &m0),
                        "compare_histories_with_all_ranks");
        else
            check(scema_cluster(ctx, flat.data(), offsets.data(), ids.data(), n, P0, thr0, var0, &m0), "compare_histories_with_all_ranks");
        clustered = true;
    } else {
        resolve_splines(hist);
        const uint32_t K = (uint32_t)hist[0]->get_spline()->size();
        std::vector<double> rows(n * (size_t)K);
        for (size_t i = 0; i < n; i++) {
            std::vector<double> *sp = hist[i]->get_spline();
            if (sp->size() != K) {
                fprintf(stderr, "Error in compare_L2_norm(): given strain6D objects have different numbers of spline points (%u and %u)\n",
                        K, (uint32_t)sp->size());
                exit(1);
            }
            if (K) memcpy(rows.data() + i * (size_t)K, sp->data(), K * sizeof(double));
        }
        if (multi()) {
            uint64_t m1 = 0;
            const bool dense1 = keep_all_similar();
            check_multi(scema_multi_compare_rows(multi(), rows.data(), n, K, ids.data(),
                                                 dense1 ? std::numeric_limits<double>::infinity() : threshold,
                                                 dense1 ? SCEMA_PAIRS_EXACT : SCEMA_PAIRS_TC, &m1),
                        "compare_histories_with_all_ranks");
            clustered = true;
        } else {
            check(scema_set_spline(ctx, rows.data(), 0, n, K, ids.data()), "compare_histories_with_all_ranks");
        }
    }
    const bool dense = keep_all_similar();
    uint64_t m = 0;
    if (!clustered)
        check(scema_compare(ctx, dense ? std::numeric_limits<double>::infinity() : threshold, dense ? SCEMA_PAIRS_EXACT : SCEMA_PAIRS_TC, 0, 1, &m),
              "compare_histories_with_all_ranks");
    check(scema_edges_device(ctx, NULL, NULL, NULL, &m), "compare_histories_with_all_ranks");
    std::vector<uint32_t> a(m), b(m);
    std::vector<double> d(m);
    check(scema_get_edges(ctx, a.data(), b.data(), d.data(), m), "compare_histories_with_all_ranks");

    // Edges are sorted by (a,b), so one pass appends to every history its partners in ascending batch
    // index: those below it first, then those above — the order of the single-rank loop
    // strain2spline.h:603-611.
    std::vector<std::vector<std::pair<uint32_t, double> > > lists(n);
    for (uint64_t e = 0; e < m; e++) {
        lists[a[e]].push_back(std::make_pair(b[e], d[e]));
        lists[b[e]].push_back(std::make_pair(a[e], d[e]));
    }
    for (size_t i = 0; i < n; i++) {
        std::vector<std::pair<uint32_t, double> > &li = lists[i];
        if (n_ranks > 1) ring_sort(li, ring_rank[i], ring_rank, n_ranks);
        for (size_t q = 0; q < li.size(); q++) {
            const bool keep = li[q].second < threshold;
            if (dense) hist[i]->b200_push_all(ids[li[q].first], li[q].second);
            hist[i]->b200_push_partner(ids[li[q].first], li[q].second, keep);
            if (!dense && keep) hist[i]->b200_note_nearest(li[q].second, ids[li[q].first]);
        }
    }
    if (!dense && keep_nearest()) {
        // the reference's global nearest neighbour (lowest ID on ties), from exact distances to ALL other histories
        std::vector<uint32_t> nid(n);
        std::vector<double> nd(n);
        check(scema_nearest(ctx, nid.data(), nd.data()), "compare_histories_with_all_ranks");
        for (size_t i = 0; i < n; i++) hist[i]->b200_set_nearest(nid[i], nd[i]);
    }
}

}  // namespace b200

// Not in the reference: resolve every deferred splinify() of the batch in one K1 launch.
inline void splinify_batch(std::vector<Strain6D *> &histories, uint32_t num_spline_points_per_component)
{
    for (size_t i = 0; i < histories.size(); i++) histories[i]->splinify(num_spline_points_per_component);
    b200::resolve_splines(histories);
}

inline double compare_L2_norm(double *a, double *b, uint32_t num_points_a, uint32_t num_points_b)
{
    if (num_points_a != num_points_b) {
        fprintf(stderr, "Error in compare_L2_norm(): given strain6D objects have different numbers of spline points (%u and %u)\n",
                num_points_a, num_points_b);
        exit(1);
    }
    // single pair on the host: not worth a kernel launch; same operation order as the GPU exact path
    double sum = 0;
    for (uint32_t i = 0; i < num_points_a; i++) {
        double diff = a[i] - b[i];
        sum += diff * diff;
    }
    return sqrt(sum);
}

inline double compare_L2_norm(Strain6D *a, Strain6D *b)
{
    std::vector<double> *sa = a->get_spline(), *sb = b->get_spline();
    return compare_L2_norm(sa->data(), sb->data(), (uint32_t)sa->size(), (uint32_t)sb->size());
}

// ---- names the reference header also exports (strain2spline.h:445-464, :499-539). The collective below gathers the
// histories instead of passing them round a ring, so it uses none of these; they behave like the reference's.
class Strain6DReceiver {
public:
    Strain6DReceiver(uint32_t in_max_buf_size)
    {
        max_buf_size = in_max_buf_size;
        spline = new double[max_buf_size];
        recv_count = 0;
        ID = 0;
    }
    ~Strain6DReceiver() { delete[] spline; }
    double *spline;
    uint32_t recv_count, max_buf_size, ID;
};

inline double compare_L2_norm(Strain6D *a, Strain6DReceiver *b)
{
    std::vector<double> *sa = a->get_spline();
    return compare_L2_norm(sa->data(), b->spline, (uint32_t)sa->size(), b->recv_count);
}

inline int32_t modulo_neg(int32_t x, int32_t n) { return ((x % n + n) % n); }

#if !defined(SCEMA_B200_NO_MPI)
inline void send_strain6D_mpi(Strain6D *in_s6D, int32_t target_rank, int32_t this_rank, MPI_Comm comm)
{
    std::vector<double> *strain = in_s6D->get_spline();
    uint32_t ID = in_s6D->get_ID();
    int32_t num_doubles_to_send = (int32_t)strain->size();
    MPI_Send(&num_doubles_to_send, 1, MPI_UNSIGNED, target_rank, this_rank, comm);
    MPI_Send(strain->data(), num_doubles_to_send, MPI_DOUBLE, target_rank, this_rank, comm);
    MPI_Send(&ID, 1, MPI_UNSIGNED, target_rank, this_rank, comm);
}

inline void receive_strain6D_mpi(Strain6DReceiver *recv, int32_t from_rank, MPI_Comm comm)
{
    MPI_Status status;
    MPI_Recv(&(recv->recv_count), 1, MPI_UNSIGNED, from_rank, from_rank, comm, &status);
    MPI_Recv(recv->spline, (int)recv->max_buf_size, MPI_DOUBLE, from_rank, from_rank, comm, &status);
    MPI_Recv(&(recv->ID), 1, MPI_UNSIGNED, from_rank, from_rank, comm, &status);
}
#endif

// Collective over `comm` exactly like the reference: every rank passes its local histories and
// returns with each of them holding the list of all other histories (on any rank) closer than
// `threshold`.
inline void compare_histories_with_all_ranks(std::vector<Strain6D *> &histories, double threshold, MPI_Comm comm)
{
#if defined(SCEMA_B200_NO_MPI)
    (void)comm;
    std::vector<int> ring(histories.size(), 0);
    b200::compare_batch(histories, threshold, ring, 1);
#else
    int rank = 0, n_ranks = 1;
    MPI_Comm_rank(comm, &rank);
    MPI_Comm_size(comm, &n_ranks);
    if (n_ranks == 1) {
        std::vector<int> ring(histories.size(), 0);
        b200::compare_batch(histories, threshold, ring, 1);
        return;
    }
    // ---- gather raw histories on rank 0 (the GPU owner)
    const int n_local = (int)histories.size();
    std::vector<int> counts(n_ranks, 0);
    MPI_Gather(&n_local, 1, MPI_INT, counts.data(), 1, MPI_INT, 0, comm);
    std::vector<unsigned> meta_local(3 * (size_t)n_local);  // steps, spline points, ID
    std::vector<double> flat_local;
    for (int i = 0; i < n_local; i++) {
        if (!histories[i]->b200_up_to_date()) { std::cout << "Spline is not up to date.\n"; exit(1); }
        meta_local[3 * i] = histories[i]->b200_num_steps();
        meta_local[3 * i + 1] = histories[i]->get_num_spline_points_per_component();
        meta_local[3 * i + 2] = histories[i]->get_ID();
        flat_local.insert(flat_local.end(), histories[i]->b200_steps().begin(), histories[i]->b200_steps().end());
    }
    std::vector<int> mcounts(n_ranks), mdispl(n_ranks), dcounts(n_ranks), ddispl(n_ranks);
    int n_total = 0;
    if (rank == 0)
        for (int r = 0; r < n_ranks; r++) { mcounts[r] = 3 * counts[r]; mdispl[r] = 3 * n_total; n_total += counts[r]; }
    std::vector<unsigned> meta(rank == 0 ? 3 * (size_t)n_total : 1);
    MPI_Gatherv(meta_local.data(), 3 * n_local, MPI_UNSIGNED, meta.data(), mcounts.data(), mdispl.data(), MPI_UNSIGNED, 0, comm);
    const int n_doubles_local = (int)flat_local.size();
    MPI_Gather(&n_doubles_local, 1, MPI_INT, dcounts.data(), 1, MPI_INT, 0, comm);
    long total_doubles = 0;
    if (rank == 0)
        for (int r = 0; r < n_ranks; r++) { ddispl[r] = (int)total_doubles; total_doubles += dcounts[r]; }
    std::vector<double> flat(rank == 0 ? (size_t)total_doubles + 1 : 1);
    MPI_Gatherv(flat_local.data(), n_doubles_local, MPI_DOUBLE, flat.data(), dcounts.data(), ddispl.data(), MPI_DOUBLE, 0, comm);

    // ---- rank 0: rebuild proxies, run the batch, serialise the per-history results
    std::vector<int> rcounts(n_ranks, 0), rdispl(n_ranks, 0);   // result doubles per rank
    std::vector<double> packed;                                 // per history: count, then (ID, diff)*
    if (rank == 0) {
        std::vector<Strain6D> proxy(n_total);
        std::vector<Strain6D *> all(n_total);
        std::vector<int> ring(n_total);
        size_t cursor = 0;
        int idx = 0;
        for (int r = 0; r < n_ranks; r++)
            for (int i = 0; i < counts[r]; i++, idx++) {
                const unsigned L = meta[3 * idx], P = meta[3 * idx + 1], id = meta[3 * idx + 2];
                for (unsigned s = 0; s < L; s++, cursor += 6)
                    proxy[idx].add_current_strain(flat[cursor], flat[cursor + 1], flat[cursor + 2], flat[cursor + 3],
                                                  flat[cursor + 4], flat[cursor + 5]);
                proxy[idx].splinify(P);
                proxy[idx].set_ID(id);
                all[idx] = &proxy[idx];
                ring[idx] = r;
            }
        b200::compare_batch(all, threshold, ring, n_ranks);
        idx = 0;
        for (int r = 0; r < n_ranks; r++) {
            rdispl[r] = (int)packed.size();
            for (int i = 0; i < counts[r]; i++, idx++) {
                // per history: partners below the threshold, the full comparison list (dense mode), the nearest neighbour
                const std::vector<HISTORY_ID_DIFF_PAIR> &ms = proxy[idx].b200_most_similar();
                packed.push_back((double)ms.size());
                for (size_t q = 0; q < ms.size(); q++) { packed.push_back((double)ms[q].ID); packed.push_back(ms[q].diff); }
                const std::vector<HISTORY_ID_DIFF_PAIR> &as = proxy[idx].b200_all_similar();
                packed.push_back((double)as.size());
                for (size_t q = 0; q < as.size(); q++) { packed.push_back((double)as[q].ID); packed.push_back(as[q].diff); }
                packed.push_back((double)proxy[idx].get_most_similar_history_ID());
                packed.push_back(proxy[idx].get_most_similar_history_diff());
            }
            rcounts[r] = (int)packed.size() - rdispl[r];
        }
    }
    int my_doubles = 0;
    MPI_Scatter(rcounts.data(), 1, MPI_INT, &my_doubles, 1, MPI_INT, 0, comm);
    std::vector<double> mine((size_t)my_doubles + 1);
    MPI_Scatterv(packed.data(), rcounts.data(), rdispl.data(), MPI_DOUBLE, mine.data(), my_doubles, MPI_DOUBLE, 0, comm);
    size_t c = 0;
    for (int i = 0; i < n_local; i++) {
        histories[i]->clear_most_similar_history();
        const size_t cnt = (size_t)mine[c++];
        for (size_t q = 0; q < cnt; q++, c += 2) histories[i]->b200_push_partner((uint32_t)mine[c], mine[c + 1], true);
        const size_t cnt_all = (size_t)mine[c++];
        for (size_t q = 0; q < cnt_all; q++, c += 2) histories[i]->b200_push_all_raw((uint32_t)mine[c], mine[c + 1]);
        histories[i]->b200_set_nearest((uint32_t)mine[c], mine[c + 1]);
        c += 2;
    }
#endif
}

}  // namespace MatHistPredict
#endif /* MATHISTPREDICT_STRAIN2SPLINE_B200_H */
