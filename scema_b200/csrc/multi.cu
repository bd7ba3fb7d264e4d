// Multi-GPU driver behind the C ABI (include/scema_hist.h, scema_multi_*): the N x N pair matrix tile-sharded over the
// GPUs of one box from ONE process — one host thread and one context per GPU, NCCL over NVLink / NVSwitch for the two
// exchanges of the path. Replaces the MPI ring of compare_histories_with_all_ranks (reference
// headers/strain2spline.h:546-614: every rank sends all its histories to every other rank, R - 1 ring steps of blocking
// point-to-point messages per history):
//   1. every GPU receives its contiguous share of the raw histories over its own PCIe link and resamples it (K1);
//   2. ONE all-gather of the resampled row blocks (grouped ncclBroadcast, in place, uneven shares allowed) gives every
//      GPU the full [n][K] matrix — the path's only real exchange;
//   3. every GPU filters / recomputes its share of the pair-matrix tiles (shard r of G, no communication). The filters
//      split the matrix differently, so when one shard reports SCEMA_ERR_DENSE all of them repeat with the next filter;
//   4. ncclAllGather of the 8-byte edge counts -> exclusive offsets on every GPU; the shards' sorted lists are sent to
//      GPU 0 (grouped ncclSend / ncclRecv at those offsets), which puts the union in canonical (a, b) order.
// Results (edge list, per-history files, graph reduction) are then served by GPU 0's context exactly as for one GPU.
// NCCL is loaded at run time (dlopen "libnccl.so.2": the copy a host application such as PyTorch has already loaded,
// else the system one), so the single-GPU library has no NCCL dependency.
#include "common.cuh"
#include <nccl.h>
#include <dlfcn.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>

using namespace scema;

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    bool load(std::string &err)
    {
        if (handle) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) { err = std::string("cannot load NCCL: ") + dlerror(); return false; }
        auto sym = [&](const char *s) { return dlsym(handle, s); };
        CommInitAll = (decltype(CommInitAll))sym("ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
        GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
        Broadcast = (decltype(Broadcast))sym("ncclBroadcast");
        AllGather = (decltype(AllGather))sym("ncclAllGather");
        Send = (decltype(Send))sym("ncclSend");
        Recv = (decltype(Recv))sym("ncclRecv");
        GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
        if (!CommInitAll || !CommDestroy || !GetErrorString || !Broadcast || !AllGather || !Send || !Recv || !GroupStart || !GroupEnd) {
            err = "NCCL library lacks a required symbol";
            return false;
        }
        return true;
    }
};

// reusable barrier for the G worker threads of one call
struct Barrier {
    std::mutex m;
    std::condition_variable cv;
    int n = 0, waiting = 0;
    uint64_t gen = 0;
    void wait()
    {
        std::unique_lock<std::mutex> lk(m);
        const uint64_t g = gen;
        if (++waiting == n) { waiting = 0; gen++; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
};

}  // namespace

struct scema_multi {
    int G = 0;
    std::vector<int> devices;
    std::vector<scema_ctx *> ctx;
    std::vector<ncclComm_t> comms;
    NcclApi nccl;
    std::string err;
    std::vector<DevBuf> full;        // per GPU: the whole spline matrix [n][K]
    std::vector<DevBuf> d_counts;    // per GPU: edge counts of all shards (ncclAllGather), 8 bytes each
    std::vector<uint64_t *> h_counts;  // pinned mirror per GPU
    DevBuf merged_key, merged_val;   // GPU 0: union of the shards' lists before the canonical sort
    std::vector<uint64_t> bounds;    // history range of every GPU
    std::vector<uint64_t> shard_edges, shard_offsets;
    std::vector<int> rc;
    std::vector<std::string> rc_msg;
    uint64_t n = 0, n_edges = 0;
    uint32_t K = 0;
    int variant_used = -1;
    bool have_edges = false;
    double last_ms[4] = {0, 0, 0, 0};  // wall clock of the phases of the last call: ingest+K1, all-gather, compare, merge
    Barrier bar;
};

namespace {

// Balanced contiguous shares whose inner boundaries are multiples of 2048 rows (whole scheduling panels and operand blocks).
void shard_bounds(uint64_t n, int G, std::vector<uint64_t> &b)
{
    b.assign(G + 1, 0);
    const uint64_t panel = (uint64_t)PANEL_ROWBLOCKS * TILE;
    for (int r = 1; r < G; r++) {
        uint64_t x = n * (uint64_t)r / (uint64_t)G;
        x = (x + panel / 2) / panel * panel;
        b[r] = std::min<uint64_t>(std::max<uint64_t>(x, b[r - 1]), n);
    }
    b[G] = n;
}

struct Call {
    scema_multi *m;
    const double *steps;        // host: raw histories (cluster) ...
    const uint64_t *offsets;
    const double *rows;         // ... or already-resampled rows (compare_rows)
    const uint32_t *ids;
    uint64_t n;
    uint32_t P, K;
    double thr;
    int variant;
    std::atomic<int> any_dense{0}, any_fail{0};
    std::vector<std::chrono::steady_clock::time_point> t;
};

#define MCHECK(r, call)                                                                       \
    do {                                                                                      \
        int rc__ = (call);                                                                    \
        if (rc__) { m->rc[r] = rc__; m->rc_msg[r] = scema_last_error(c); c_.any_fail = 1; }   \
    } while (0)
#define NCHECK(r, call)                                                                                    \
    do {                                                                                                   \
        ncclResult_t nr__ = (call);                                                                        \
        if (nr__ != ncclSuccess && !m->rc[r]) {                                                            \
            m->rc[r] = SCEMA_ERR_CUDA; m->rc_msg[r] = std::string(#call) + ": " + m->nccl.GetErrorString(nr__); c_.any_fail = 1; \
        }                                                                                                  \
    } while (0)
#define CCHECK(r, call)                                                                                    \
    do {                                                                                                   \
        cudaError_t ce__ = (call);                                                                         \
        if (ce__ != cudaSuccess && !m->rc[r]) {                                                            \
            m->rc[r] = ce__ == cudaErrorMemoryAllocation ? SCEMA_ERR_NOMEM : SCEMA_ERR_CUDA;               \
            m->rc_msg[r] = std::string(#call) + ": " + cudaGetErrorString(ce__); c_.any_fail = 1;          \
        }                                                                                                  \
    } while (0)

void worker(Call &c_, int r)
{
    scema_multi *m = c_.m;
    scema_ctx *c = m->ctx[r];
    const int G = m->G;
    const uint64_t n = c_.n, b = m->bounds[r], e = m->bounds[r + 1];
    const uint32_t K = c_.K;
    cudaSetDevice(m->devices[r]);
    cudaStream_t st = c->stream;
    auto now = [] { return std::chrono::steady_clock::now(); };

    // ---- 1. own share: ingest + K1 (or the caller's rows), straight into its place in the full matrix
    CCHECK(r, m->full[r].reserve(std::max<uint64_t>(n * K, 1) * sizeof(double)));
    double *full = m->full[r].as<double>();
    if (!c_.any_fail && e > b) {
        if (c_.steps) {
            std::vector<uint64_t> off(e - b + 1);
            for (uint64_t i = b; i <= e; i++) off[i - b] = c_.offsets[i] - c_.offsets[b];
            MCHECK(r, scema_set_histories(c, c_.steps + c_.offsets[b] * 6, 0, off.data(), c_.ids ? c_.ids + b : nullptr, e - b));
            if (!m->rc[r]) MCHECK(r, scema_resample(c, c_.P));
            if (!m->rc[r]) CCHECK(r, cudaMemcpyAsync(full + b * K, c->d_spline, (e - b) * K * sizeof(double), cudaMemcpyDeviceToDevice, st));
        } else {
            CCHECK(r, cudaMemcpyAsync(full + b * K, c_.rows + b * K, (e - b) * K * sizeof(double), cudaMemcpyHostToDevice, st));
        }
    }
    if (r == 0) { cudaStreamSynchronize(st); c_.t[1] = now(); }
    m->bar.wait();  // nobody enters a collective after a failure elsewhere
    if (c_.any_fail) return;

    // ---- 2. all-gather of the row blocks, in place: one broadcast per owner, fused into one NCCL group
    NCHECK(r, m->nccl.GroupStart());
    for (int root = 0; root < G; root++) {
        const uint64_t rb = m->bounds[root], re = m->bounds[root + 1];
        if (re > rb) NCHECK(r, m->nccl.Broadcast(full + rb * K, full + rb * K, (re - rb) * K, ncclDouble, root, m->comms[r], st));
    }
    NCHECK(r, m->nccl.GroupEnd());
    if (r == 0) { cudaStreamSynchronize(st); c_.t[2] = now(); }
    if (!m->rc[r]) MCHECK(r, scema_set_spline(c, full, 1, n, K, c_.ids));

    // ---- 3. this GPU's share of the pair matrix; all shards change filter together
    static const int next_variant[4] = {SCEMA_PAIRS_EXACT, SCEMA_PAIRS_EXACT, -1, SCEMA_PAIRS_DMMA};
    int variant = c_.variant;
    uint64_t ne = 0;
    while (true) {
        m->bar.wait();
        if (c_.any_fail) return;
        int rc = scema_compare(c, c_.thr, variant, (uint32_t)r, (uint32_t)G, &ne);
        if (rc == SCEMA_ERR_DENSE && next_variant[variant] >= 0) c_.any_dense = 1;
        else if (rc) { m->rc[r] = rc; m->rc_msg[r] = scema_last_error(c); c_.any_fail = 1; }
        m->bar.wait();
        if (c_.any_fail) return;
        if (!c_.any_dense) break;
        m->bar.wait();          // everybody has read the flag
        if (r == 0) c_.any_dense = 0;
        variant = next_variant[variant];
    }
    if (r == 0) m->variant_used = variant;
    m->shard_edges[r] = ne;

    // ---- 4. edge counts -> offsets on every GPU (ncclAllGather), then the shards' lists to GPU 0
    CCHECK(r, m->d_counts[r].reserve((size_t)(G + 1) * sizeof(uint64_t)));
    if (!m->rc[r]) {
        uint64_t *dc = m->d_counts[r].as<uint64_t>();
        CCHECK(r, cudaMemcpyAsync(dc + G, &m->shard_edges[r], sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        NCHECK(r, m->nccl.AllGather(dc + G, dc, 1, ncclUint64, m->comms[r], st));
        CCHECK(r, cudaMemcpyAsync(m->h_counts[r], dc, (size_t)G * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        CCHECK(r, cudaStreamSynchronize(st));
    }
    if (r == 0) c_.t[3] = now();
    m->bar.wait();
    if (c_.any_fail) return;
    uint64_t total = 0;
    for (int q = 0; q < G; q++) total += m->h_counts[r][q];
    if (r == 0) {
        for (int q = 0; q < G; q++) m->shard_offsets[q] = q ? m->shard_offsets[q - 1] + m->h_counts[0][q - 1] : 0;
        CCHECK(r, m->merged_key.reserve(std::max<uint64_t>(total, 1) * sizeof(uint64_t)));
        CCHECK(r, m->merged_val.reserve(std::max<uint64_t>(total, 1) * sizeof(double)));
    }
    m->bar.wait();
    if (c_.any_fail) return;
    const uint64_t *keys = nullptr;
    const double *vals = nullptr;
    uint32_t shift = 0;
    uint64_t mine = 0;
    MCHECK(r, scema_edges_device(c, &keys, &vals, &shift, &mine));
    m->bar.wait();
    if (c_.any_fail) return;
    NCHECK(r, m->nccl.GroupStart());
    if (r == 0) {
        if (mine) {
            CCHECK(r, cudaMemcpyAsync(m->merged_key.p, keys, mine * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
            CCHECK(r, cudaMemcpyAsync(m->merged_val.p, vals, mine * sizeof(double), cudaMemcpyDeviceToDevice, st));
        }
        for (int q = 1; q < G; q++) {
            const uint64_t cnt = m->h_counts[0][q];
            if (!cnt) continue;
            NCHECK(r, m->nccl.Recv(m->merged_key.as<uint64_t>() + m->shard_offsets[q], cnt, ncclUint64, q, m->comms[0], st));
            NCHECK(r, m->nccl.Recv(m->merged_val.as<double>() + m->shard_offsets[q], cnt, ncclDouble, q, m->comms[0], st));
        }
    } else if (mine) {
        NCHECK(r, m->nccl.Send(keys, mine, ncclUint64, 0, m->comms[r], st));
        NCHECK(r, m->nccl.Send(vals, mine, ncclDouble, 0, m->comms[r], st));
    }
    NCHECK(r, m->nccl.GroupEnd());
    if (r == 0 && !m->rc[0]) {
        int rc = edges_adopt(c, m->merged_key.as<uint64_t>(), m->merged_val.as<double>(), total);
        if (rc) { m->rc[0] = rc; m->rc_msg[0] = scema_last_error(c); c_.any_fail = 1; }
        m->n_edges = total;
    }
    CCHECK(r, cudaStreamSynchronize(st));
    if (r == 0) c_.t[4] = now();
    m->bar.wait();
}

int run_call(scema_multi *m, Call &c_)
{
    m->err.clear();
    m->have_edges = false;
    m->n = c_.n;
    m->K = c_.K;
    shard_bounds(c_.n, m->G, m->bounds);
    std::fill(m->rc.begin(), m->rc.end(), 0);
    for (auto &s : m->rc_msg) s.clear();
    c_.t.assign(5, std::chrono::steady_clock::now());
    std::vector<std::thread> th;
    for (int r = 1; r < m->G; r++) th.emplace_back(worker, std::ref(c_), r);
    worker(c_, 0);
    for (auto &t : th) t.join();
    for (int r = 0; r < m->G; r++)
        if (m->rc[r]) { m->err = "GPU " + std::to_string(m->devices[r]) + ": " + m->rc_msg[r]; return m->rc[r]; }
    for (int k = 0; k < 4; k++) m->last_ms[k] = std::chrono::duration<double, std::milli>(c_.t[k + 1] - c_.t[k]).count();
    m->have_edges = true;
    return SCEMA_OK;
}

}  // namespace

extern "C" {

int scema_multi_create(scema_multi **out, const int *devices, int n_devices)
{
    if (!out || n_devices < 1 || n_devices > 64) return SCEMA_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { cudaGetLastError(); return SCEMA_ERR_CUDA; }
    scema_multi *m = new (std::nothrow) scema_multi();
    if (!m) return SCEMA_ERR_NOMEM;
    m->G = n_devices;
    for (int r = 0; r < n_devices; r++) {
        const int d = devices ? devices[r] : r;
        if (d < 0 || d >= count || std::find(m->devices.begin(), m->devices.end(), d) != m->devices.end()) { delete m; return SCEMA_ERR_INVALID; }
        m->devices.push_back(d);
    }
    m->ctx.assign(n_devices, nullptr);
    m->full.resize(n_devices);
    m->d_counts.resize(n_devices);
    m->h_counts.assign(n_devices, nullptr);
    m->rc.assign(n_devices, 0);
    m->rc_msg.assign(n_devices, "");
    m->shard_edges.assign(n_devices, 0);
    m->shard_offsets.assign(n_devices, 0);
    m->bar.n = n_devices;
    int rc = SCEMA_OK;
    for (int r = 0; r < n_devices && !rc; r++) {
        rc = scema_create(&m->ctx[r], m->devices[r], nullptr);
        if (!rc && cudaMallocHost(&m->h_counts[r], (size_t)(n_devices + 1) * sizeof(uint64_t)) != cudaSuccess) rc = SCEMA_ERR_NOMEM;
    }
    if (!rc && n_devices > 1) {
        if (!m->nccl.load(m->err)) rc = SCEMA_ERR_CUDA;
        else {
            m->comms.assign(n_devices, nullptr);
            ncclResult_t nr = m->nccl.CommInitAll(m->comms.data(), n_devices, m->devices.data());
            if (nr != ncclSuccess) { m->comms.clear(); rc = SCEMA_ERR_CUDA; }
        }
    }
    if (rc) { scema_multi_destroy(m); return rc; }
    *out = m;
    return SCEMA_OK;
}

void scema_multi_destroy(scema_multi *m)
{
    if (!m) return;
    for (int r = 0; r < m->G; r++) {
        if (r < (int)m->devices.size()) cudaSetDevice(m->devices[r]);
        if (r < (int)m->comms.size() && m->comms[r]) m->nccl.CommDestroy(m->comms[r]);
        if (r < (int)m->full.size()) { m->full[r].release(); m->d_counts[r].release(); }
        if (r < (int)m->h_counts.size() && m->h_counts[r]) cudaFreeHost(m->h_counts[r]);
        if (r == 0) { m->merged_key.release(); m->merged_val.release(); }
        if (r < (int)m->ctx.size() && m->ctx[r]) scema_destroy(m->ctx[r]);
    }
    delete m;
}

const char *scema_multi_last_error(const scema_multi *m) { return m ? m->err.c_str() : "null handle"; }
int scema_multi_devices(const scema_multi *m) { return m ? m->G : 0; }
scema_ctx *scema_multi_context(scema_multi *m, int rank) { return (m && rank >= 0 && rank < m->G) ? m->ctx[rank] : nullptr; }

int scema_multi_cluster(scema_multi *m, const double *steps, const uint64_t *offsets, const uint32_t *ids, uint64_t n,
                        uint32_t spline_points, double threshold, int variant, uint64_t *n_edges)
{
    if (!m) return SCEMA_ERR_INVALID;
    if (n && (!steps || !offsets)) { m->err = "cluster: null pointer"; return SCEMA_ERR_INVALID; }
    if (variant < 0 || variant > 3 || spline_points == 0) { m->err = "cluster: bad variant or spline_points"; return SCEMA_ERR_INVALID; }
    if (m->G == 1) {
        int rc = scema_cluster(m->ctx[0], steps, offsets, ids, n, spline_points, threshold, variant, n_edges);
        if (rc) m->err = scema_last_error(m->ctx[0]);
        m->have_edges = !rc;
        m->n = n; m->K = 6 * spline_points; m->variant_used = variant;
        if (!rc) { uint64_t e = 0; scema_edges_device(m->ctx[0], nullptr, nullptr, nullptr, &e); m->n_edges = e; m->shard_edges[0] = e; }
        return rc;
    }
    for (uint64_t i = 0; i < n; i++)
        if (offsets[i + 1] < offsets[i] || offsets[i + 1] - offsets[i] < 3) {
            m->err = offsets[i + 1] < offsets[i] ? "set_histories: offsets not monotone" : "Not enough strain steps added. Need at least 3 points for splinify().";
            return SCEMA_ERR_INVALID;
        }
    Call c;
    c.m = m; c.steps = steps; c.offsets = offsets; c.rows = nullptr; c.ids = ids; c.n = n; c.P = spline_points; c.K = 6 * spline_points;
    c.thr = threshold; c.variant = variant;
    int rc = run_call(m, c);
    if (!rc && n_edges) *n_edges = m->n_edges;
    return rc;
}

int scema_multi_compare_rows(scema_multi *m, const double *rows, uint64_t n, uint32_t k, const uint32_t *ids, double threshold,
                             int variant, uint64_t *n_edges)
{
    if (!m) return SCEMA_ERR_INVALID;
    if (n && k && !rows) { m->err = "compare_rows: null pointer"; return SCEMA_ERR_INVALID; }
    if (variant < 0 || variant > 3) { m->err = "compare_rows: bad variant"; return SCEMA_ERR_INVALID; }
    if (m->G == 1) {
        int rc = scema_set_spline(m->ctx[0], rows, 0, n, k, ids);
        if (!rc) rc = scema_compare(m->ctx[0], threshold, variant, 0, 1, n_edges);
        if (rc) m->err = scema_last_error(m->ctx[0]);
        m->have_edges = !rc;
        m->n = n; m->K = k; m->variant_used = variant;
        if (!rc) { uint64_t e = 0; scema_edges_device(m->ctx[0], nullptr, nullptr, nullptr, &e); m->n_edges = e; m->shard_edges[0] = e; }
        return rc;
    }
    Call c;
    c.m = m; c.steps = nullptr; c.offsets = nullptr; c.rows = rows; c.ids = ids; c.n = n; c.P = 0; c.K = k; c.thr = threshold; c.variant = variant;
    int rc = run_call(m, c);
    if (!rc && n_edges) *n_edges = m->n_edges;
    return rc;
}

int scema_multi_shard_edges(scema_multi *m, uint64_t *counts, uint64_t *offsets)
{
    if (!m || !m->have_edges) return SCEMA_ERR_STATE;
    for (int r = 0; r < m->G; r++) {
        if (counts) counts[r] = m->shard_edges[r];
        if (offsets) offsets[r] = m->G > 1 ? m->shard_offsets[r] : 0;
    }
    return SCEMA_OK;
}

int scema_multi_last_ms(scema_multi *m, double ms[4], int *variant_used)
{
    if (!m) return SCEMA_ERR_INVALID;
    for (int k = 0; k < 4; k++) ms[k] = m->last_ms[k];
    if (variant_used) *variant_used = m->variant_used;
    return SCEMA_OK;
}

}  // extern "C"
