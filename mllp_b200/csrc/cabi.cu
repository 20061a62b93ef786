// cabi.cu -- the C ABI of include/mllp_b200.h for the single-instance path.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mllp_b200.h"
#include "lp_format.h"
#include "pdhg_host.h"
#include "pdhg_kernels.cuh"

using namespace mllp;

namespace {
thread_local std::string g_err;

int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char* what)
{
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return (int)e;
}
int env_int(const char* name, int dflt)
{
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}
}  // namespace

namespace mllp {
void set_last_error(const std::string& msg) { g_err = msg; }  // used by batch_kernels.cu
static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}

#define CUDA_OK(call)                                           \
    do {                                                        \
        cudaError_t e_ = (call);                                \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call);     \
    } while (0)
#define RC_OK(call)                                                                   \
    do {                                                                              \
        int rc_ = (call);                                                             \
        if (rc_ != 0) {                                                               \
            if (rc_ < 1000) return cuda_fail((cudaError_t)rc_, #call);                \
            return rc_;                                                               \
        }                                                                             \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

struct mllp_lp {
    int device = 0;
    int m = 0, n = 0;
    int64_t nnz = 0;
    uint32_t flags = 0;
    bool bounds = false;
    int G = 0, threads = 0;
    size_t dyn_smem = 0;          // dynamic shared memory of the persistent kernels
    DevLP d{};
    std::vector<void*> allocs;
    std::vector<void*> mat_allocs;    // device arrays of the two tiled matrices (replaced by the tuning rounds)
    std::vector<void*>* sink = &allocs;
    double tune_ns[2] = {0.0, 0.0};   // measured ns / iteration before and after the tuning rounds
    double geom_ns[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // geometry search: ns / iteration of grid, cluster 16 / 8 / 4, one CTA, broadcast cluster 16 / 8 / 4 / 1 (0 = not tried)
    int tune_rounds = 0;
    BlockPlan* blocks = nullptr;      // block-angular structure (blocks.cu), or null
    bool use_blocks = false;          // the parity kernel runs on it (chosen by timing both)
    double blocks_ns[2] = {0.0, 0.0}; // ns / iteration of the grid kernel and of the block kernel
    int32_t* d_orderX = nullptr;  // internal position k holds original column order[k]
    int32_t* d_orderY = nullptr;
    double* tmp_n = nullptr;
    double* tmp_m = nullptr;
    double* tmp_n2 = nullptr;
    double* d_b = nullptr;        // writable aliases of d.b / d.c
    double* d_c = nullptr;
    double* u_x = nullptr;        // user-order staging for the *_host entry
    double* u_y = nullptr;
    double* u_b = nullptr;
    double* u_c = nullptr;
    double* d_box[4] = {nullptr, nullptr, nullptr, nullptr};   // lb, ub (ni), ylo, yhi (mi) in internal order (general form)
    double* d_dr = nullptr;       // preconditioned handle: row / column scaling in internal order (d.dr / d.dc)
    double* d_dc = nullptr;
    double* d_scal = nullptr;     // MLLP_NUM_SCALARS
    double* d_norm2 = nullptr;
    int64_t info[16] = {0};
    cudaGraphExec_t graph = nullptr;  // graph mode: GRAPH_UNROLL iterations
    // row partition over GPUs (nranks > 1): this rank owns a slice of the rows of A (its entries of y); the A' phase is
    // replicated (every rank holds all of A' and of x), so the internal y has nranks * Ly entries (equal padded slices)
    // and ONE exchange per iteration moves the slices: tagged words through peer mailboxes inside the persistent
    // kernel, or one NCCL all-gather between the two launches of an iteration
    int rank = 0, nranks = 1;
    int mi = 0, ni = 0;               // internal vector lengths (mi padded)
    int Ly = 0;                       // slice length of y
    void* comm = nullptr;             // ncclComm_t
    // in-kernel exchange over NVLink peer memory (mllp_rowpart_ipc_export / _import)
    PeerInfo peers{};
    unsigned long long* d_mail = nullptr;   // this rank's mailbox: 2 buffers x mi tagged 16-byte words, written by the peers
    unsigned* d_err = nullptr;
    unsigned long long join_epoch = 0;   // last tag used by the polled split-row join
    bool p2p_ready = false;
    unsigned long long xseq = 0;      // exchanges done on this handle (tag / buffer of the next one; same on all ranks)
    McState mcs{};                    // NVSwitch multicast mailbox (mllp_rowpart_mc_*), replaces d_mail + peer pointers
    std::vector<void*> ipc_opened;
};
constexpr int GRAPH_UNROLL = 32;

namespace {

template <class T>
int dev_alloc(mllp_lp* lp, T** out, size_t count)
{
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (count ? count : 1) * sizeof(T));
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    lp->sink->push_back(p);
    *out = (T*)p;
    return 0;
}
template <class T>
int dev_upload(mllp_lp* lp, T** out, const T* host, size_t count)
{
    RC_OK(dev_alloc(lp, out, count));
    if (count) CUDA_OK(cudaMemcpy(*out, host, count * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}
template <class T>
int dev_zeros(mllp_lp* lp, T** out, size_t count)
{
    RC_OK(dev_alloc(lp, out, count));
    CUDA_OK(cudaMemset(*out, 0, (count ? count : 1) * sizeof(T)));
    return 0;
}

struct SinkGuard {   // route the allocations of a scope into lp->mat_allocs
    mllp_lp* lp;
    explicit SinkGuard(mllp_lp* l) : lp(l) { lp->sink = &lp->mat_allocs; }
    ~SinkGuard() { lp->sink = &lp->allocs; }
};

int upload_mat(mllp_lp* lp, const HostMat& H, DevMat& D)
{
    SinkGuard sg(lp);
    double* vals; int32_t* idx; Tile* tiles; uint32_t* cb; uint32_t* csb; SplitRow* sp;
    LocalSplit* lsp; uint32_t* clb; uint32_t* cns;
    RC_OK(dev_upload(lp, &vals, H.vals.data(), H.vals.size()));
    RC_OK(dev_upload(lp, &idx, H.idx.data(), H.idx.size()));
    RC_OK(dev_upload(lp, &tiles, H.tiles.data(), H.tiles.size()));
    RC_OK(dev_upload(lp, &cb, H.cta_begin.data(), H.cta_begin.size()));
    RC_OK(dev_upload(lp, &csb, H.cta_step_begin.data(), H.cta_step_begin.size()));
    RC_OK(dev_upload(lp, &sp, H.splits.data(), H.splits.size()));
    RC_OK(dev_upload(lp, &lsp, H.lsplits.data(), H.lsplits.size()));
    RC_OK(dev_upload(lp, &clb, H.cta_lsplit_begin.data(), H.cta_lsplit_begin.size()));
    RC_OK(dev_upload(lp, &cns, H.cta_nsplit.data(), H.cta_nsplit.size()));
    D.lsplits = lsp; D.cta_lsplit_begin = clb; D.cta_nsplit = cns;
    RC_OK(dev_zeros(lp, &D.partials, (size_t)H.num_partials + 1));
    RC_OK(dev_zeros(lp, &D.slots, 2 * (size_t)H.num_partials + 2));
    RC_OK(dev_zeros(lp, &D.counters, H.splits.size()));
    D.vals = reinterpret_cast<const double2*>(vals);
    D.idx = reinterpret_cast<const int2*>(idx);
    D.tiles = tiles;
    D.cta_begin = cb;
    D.cta_step_begin = csb;
    D.splits = sp;
    D.nrows = H.nrows;
    D.ncols = H.ncols;
    return 0;
}

int build_graph(mllp_lp* lp)
{
    cudaStream_t cs;
    CUDA_OK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    cudaGraph_t g = nullptr;
    CUDA_OK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    int rc = 0;
    for (int i = 0; i < GRAPH_UNROLL && rc == 0; ++i) {
        rc = launch_primal(lp->d, lp->bounds, lp->G, lp->threads, cs);
        if (rc == 0) rc = launch_dual(lp->d, lp->bounds, lp->G, lp->threads, cs);
    }
    cudaError_t e = cudaStreamEndCapture(cs, &g);
    if (rc != 0 || e != cudaSuccess) {
        cudaStreamDestroy(cs);
        return rc != 0 ? cuda_fail((cudaError_t)rc, "graph capture") : cuda_fail(e, "cudaStreamEndCapture");
    }
    e = cudaGraphInstantiate(&lp->graph, g, 0);
    cudaGraphDestroy(g);
    cudaStreamDestroy(cs);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGraphInstantiate");
    return 0;
}

// ---- NCCL, bound at run time (the library is usually already loaded by torch.distributed) ----
struct NcclId { char internal[128]; };   // ncclUniqueId
typedef int (*nccl_get_unique_id_t)(NcclId*);
typedef int (*nccl_comm_init_rank_t)(void**, int, NcclId, int);
typedef int (*nccl_all_gather_t)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*nccl_all_reduce_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_comm_destroy_t)(void*);
typedef const char* (*nccl_get_error_string_t)(int);
struct NcclApi {
    void* lib = nullptr;
    nccl_get_unique_id_t get_unique_id = nullptr;
    nccl_comm_init_rank_t comm_init_rank = nullptr;
    nccl_all_gather_t all_gather = nullptr;
    nccl_all_reduce_t all_reduce = nullptr;
    nccl_comm_destroy_t comm_destroy = nullptr;
    nccl_get_error_string_t get_error_string = nullptr;
};
constexpr int NCCL_FLOAT64 = 8, NCCL_SUM = 0;

NcclApi* nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char* names[] = {getenv("MLLP_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            if (!nm || !*nm) continue;
            api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (api.lib) {
            api.get_unique_id = (nccl_get_unique_id_t)dlsym(api.lib, "ncclGetUniqueId");
            api.comm_init_rank = (nccl_comm_init_rank_t)dlsym(api.lib, "ncclCommInitRank");
            api.all_gather = (nccl_all_gather_t)dlsym(api.lib, "ncclAllGather");
            api.all_reduce = (nccl_all_reduce_t)dlsym(api.lib, "ncclAllReduce");
            api.comm_destroy = (nccl_comm_destroy_t)dlsym(api.lib, "ncclCommDestroy");
            api.get_error_string = (nccl_get_error_string_t)dlsym(api.lib, "ncclGetErrorString");
        }
    }
    if (!api.lib || !api.get_unique_id || !api.comm_init_rank || !api.all_gather || !api.all_reduce || !api.comm_destroy)
        return nullptr;
    return &api;
}
int nccl_fail(int rc, const char* what)
{
    NcclApi* a = nccl_api();
    g_err = std::string(what) + ": NCCL error " + std::to_string(rc) + (a && a->get_error_string ? std::string(" (") + a->get_error_string(rc) + ")" : "");
    return 2000 + rc;
}
#define NCCL_OK(call)                                     \
    do {                                                  \
        int r_ = (call);                                  \
        if (r_ != 0) return nccl_fail(r_, #call);         \
    } while (0)


// Shared-memory residency of the matrix slices (persistent kernels): sizes lp->dyn_smem and the per-CTA
// caps of resident warp-steps from the built images.
static int configure_residency(mllp_lp* lp, const HostMat& HA, const HostMat& HAT, const cudaDeviceProp& prop, int bpsm,
                               bool resident)
{
    const size_t desc_bytes = 16 * ((size_t)HA.max_cta_tiles + (size_t)HAT.max_cta_tiles);
    const bool bcast = lp->d.sync_mode == SYNC_BCAST;
    // Own entries of the CTA's rows (y, b / x, c) resident in shared memory (parity kernel): only when (nearly) the whole
    // matrix share stays resident next to them (measured: ken-18 -2 %; on osa-60, where they would push matrix steps
    // out of shared memory, +2 %).  SYNC_BCAST: always, with the anchors (6 arrays) and the two full vector copies.
    bool own = env_int("MLLP_OWN", 1) != 0 && resident && lp->nranks == 1 && !(lp->flags & MLLP_F_GRAPH_MODE);
    const size_t rows_a = (size_t)((HA.max_cta_rows + 1) & ~1), rows_at = (size_t)((HAT.max_cta_rows + 1) & ~1);
    size_t own_bytes = 16 * (rows_a + rows_at);
    if (bcast) {
        own = true;
        own_bytes = 24 * (rows_a + rows_at) + 8 * ((size_t)((lp->m + 1) & ~1) + (size_t)((lp->n + 1) & ~1));
    } else {
        const size_t res_cap = (size_t)env_int("MLLP_RES_KB", 140) * 1024;
        const size_t all = 768 * ((size_t)HA.max_cta_steps + (size_t)HAT.max_cta_steps);
        if (env_int("MLLP_OWN", 1) < 2 && all + own_bytes > res_cap + res_cap / 7) own = false;   // MLLP_OWN=2 forces it (dev knob)
    }
    if (!own) own_bytes = 0;
    lp->d.own_rows_A = own ? (uint32_t)rows_a : 0u;
    lp->d.own_rows_AT = own ? (uint32_t)rows_at : 0u;
    lp->dyn_smem = desc_bytes + own_bytes;
    lp->d.res_steps_A = 0; lp->d.res_steps_AT = 0;
    if (resident) {
        const size_t per_cta = (size_t)prop.sharedMemPerMultiprocessor / (size_t)(lp->d.sync_mode == SYNC_GRID ? bpsm : 1);
        size_t budget = std::min<size_t>(per_cta - 1024, (size_t)prop.sharedMemPerBlockOptin);
        budget -= std::min<size_t>(budget, 6144);  // static shared memory of the kernels + slack
        if (bcast && budget < desc_bytes + own_bytes)
            return fail(MLLP_E_STATE, "mllp_lp_create: the vector copies of the broadcast geometry do not fit in shared memory");
        budget -= std::min<size_t>(budget, desc_bytes + own_bytes);
        // The gathered vectors are read through L1 (the grid barrier invalidates it), so part of
        // the SM's 228 KB stays L1: cap the shared-memory share of the matrix.  (SYNC_BCAST gathers from shared memory.)
        if (!bcast) {
            const size_t res_cap = (size_t)env_int("MLLP_RES_KB", 140) * 1024;
            budget = std::min<size_t>(budget, res_cap > own_bytes ? res_cap - own_bytes : 0);
        }
        // what is left keeps (a prefix of) each CTA's share of A' and A resident
        const size_t cap = budget / 768;
        size_t a = (size_t)HA.max_cta_steps, at = (size_t)HAT.max_cta_steps;
        if (a + at > cap) {   // A' first: its tiles are the short, latency-dominated ones
            at = std::min(at, cap);
            a = std::min(a, cap - at);
        }
        const int lim = env_int("MLLP_RES_STEPS", -1);  // dev knob: cap the resident steps
        if (lim >= 0) { a = std::min<size_t>(a, (size_t)lim); at = std::min<size_t>(at, (size_t)lim); }
        lp->d.res_steps_A = (uint32_t)a; lp->d.res_steps_AT = (uint32_t)at;
        lp->dyn_smem = desc_bytes + own_bytes + 768 * (a + at);
    }
    if (lp->dyn_smem > 48 * 1024 - 4096) RC_OK(persistent_set_smem(lp->bounds, lp->dyn_smem));
    if (lp->d.sync_mode == SYNC_CLUSTER || bcast) {
        if (persistent_cluster_fits(lp->G, lp->threads, lp->bounds, lp->dyn_smem) != 1)
            return fail(MLLP_E_STATE, "mllp_lp_create: a cluster of this many CTAs cannot be resident");
        return 0;
    }
    if (persistent_max_blocks_per_sm(lp->threads, lp->bounds, lp->dyn_smem) < (lp->d.sync_mode == SYNC_CTA ? 1 : bpsm))
        return fail(MLLP_E_STATE, "mllp_lp_create: persistent grid does not fit with the chosen shared memory");
    return 0;
}

static void free_mats(mllp_lp* lp)
{
    for (void* p : lp->mat_allocs) cudaFree(p);
    lp->mat_allocs.clear();
}

// `iters` traced parity iterations on the current internal state: per CTA and iteration
// [A' phase done, barrier released, A phase done, barrier released] (globaltimer ns).
static int run_trace(mllp_lp* lp, double tau, double sigma, int iters, std::vector<unsigned long long>& out)
{
    unsigned long long* d_tr = nullptr;
    const size_t cnt = (size_t)iters * lp->G * 4;
    out.assign(cnt, 0ull);
    CUDA_OK(cudaMalloc(&d_tr, cnt * sizeof(unsigned long long)));
    int rc = (int)cudaMemset(d_tr, 0, cnt * sizeof(unsigned long long));
    lp->d.join_base = lp->join_epoch;
    lp->join_epoch += 2ull * (unsigned long long)iters + 2ull;
    DevLP d = lp->d;
    d.trace = d_tr;
    if (rc == 0) rc = launch_pdhg_persistent(d, lp->bounds, lp->G, lp->threads, lp->dyn_smem, tau, sigma, iters, 0);
    if (rc == 0) rc = (int)cudaDeviceSynchronize();
    if (rc == 0) rc = (int)cudaMemcpy(out.data(), d_tr, cnt * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(d_tr);
    if (rc != 0) return cuda_fail((cudaError_t)rc, "traced persistent run");
    return 0;
}

// Mean work time of every CTA in the A' phase and in the A phase (ns) and the mean iteration time.
struct PhaseTimes {
    std::vector<double> at, a;
    double iter_ns = 0.0;
};
static int measure_phases(mllp_lp* lp, PhaseTimes& T)
{
    const int iters = 48, skip = 8, G = lp->G;
    std::vector<unsigned long long> tr;
    RC_OK(run_trace(lp, 1.0, 1.0, iters, tr));
    T.at.assign((size_t)G, 0.0);
    T.a.assign((size_t)G, 0.0);
    auto at = [&](int it, int g, int k) { return tr[((size_t)it * G + g) * 4 + k]; };
    for (int g = 0; g < G; ++g) {
        double sa = 0, sat = 0;
        for (int it = skip; it < iters; ++it) {
            sat += (double)(long long)(at(it, g, 0) - at(it - 1, g, 3));
            sa += (double)(long long)(at(it, g, 2) - at(it, g, 1));
        }
        T.at[g] = sat / (iters - skip);
        T.a[g] = sa / (iters - skip);
    }
    unsigned long long t0 = 0, t1 = 0;
    for (int g = 0; g < G; ++g) {
        t0 = std::max(t0, at(skip - 1, g, 3));
        t1 = std::max(t1, at(iters - 1, g, 3));
    }
    T.iter_ns = (double)(long long)(t1 - t0) / (iters - skip);
    return 0;
}

// One launch of `iters` parity iterations on the handle's internal state: the block kernel when the LP is block-angular
// and it measured faster, else the persistent kernel of the chosen geometry.
static int launch_parity(mllp_lp* lp, double tau, double sigma, int iters, cudaStream_t s)
{
    lp->d.join_base = lp->join_epoch;
    lp->join_epoch += 2ull * (unsigned long long)iters + 2ull;
    if (lp->use_blocks && lp->blocks)
        return blocks_run(lp->blocks, lp->d.x, lp->d.y, lp->d.b, lp->d.c, tau, sigma, iters, lp->d.join_base, s);
    return launch_pdhg_persistent(lp->d, lp->bounds, lp->G, lp->threads, lp->dyn_smem, tau, sigma, iters, s);
}

// ns per parity iteration as currently configured (CUDA events around one launch of `iters` iterations on the handle's
// internal state, after a short warm-up launch).
static int time_parity(mllp_lp* lp, int iters, double& ns_per_iter)
{
    cudaEvent_t e0, e1;
    CUDA_OK(cudaEventCreate(&e0));
    CUDA_OK(cudaEventCreate(&e1));
    int rc = launch_parity(lp, 1.0, 1.0, 16, 0);
    if (rc == 0) rc = (int)cudaEventRecord(e0, 0);
    if (rc == 0) rc = launch_parity(lp, 1.0, 1.0, iters, 0);
    if (rc == 0) rc = (int)cudaEventRecord(e1, 0);
    if (rc == 0) rc = (int)cudaEventSynchronize(e1);
    float ms = 0.f;
    if (rc == 0) rc = (int)cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (rc != 0) return cuda_fail((cudaError_t)rc, "timed parity run");
    ns_per_iter = (double)ms * 1e6 / (double)iters;
    return 0;
}

// One feedback step of the dealing from measured per-CTA phase times (see DealFeedback).
static void update_feedback(const HostMat& H, const std::vector<double>& t, bool contiguous, DealFeedback& fb)
{
    const size_t G = t.size();
    double mean = 0;
    for (double v : t) mean += v;
    mean /= (double)G;
    if (!(mean > 0)) return;
    // fixed cost of a phase (start-up after the barrier, first round trip): half of the fastest CTA's time
    double t0 = mean;
    for (double v : t) t0 = std::min(t0, v);
    t0 *= 0.5;
    if (contiguous) {
        if (fb.tile_w.size() != H.reg_cta.size()) fb.tile_w.assign(H.reg_cta.size(), 1.0);
        if (fb.cta_f.size() != G) fb.cta_f.assign(G, 1.0);
        std::vector<double> f(G);
        for (size_t g = 0; g < G; ++g) f[g] = std::min(2.0, std::max(0.5, (t[g] - t0) / (mean - t0)));
        for (size_t k = 0; k < H.reg_cta.size(); ++k) fb.tile_w[k] *= f[H.reg_cta[k]];
        for (size_t g = 0; g < G; ++g) fb.cta_f[g] *= f[g];
    } else {
        if (fb.cta_bias.size() != G) fb.cta_bias.assign(G, 0.0);
        double load = 0;
        for (double v : H.cta_load) load += v;
        if (!(load > 0)) return;
        const double unit_ns = (mean - t0) * (double)G / load;   // ns per load unit
        for (size_t g = 0; g < G; ++g) fb.cta_bias[g] += 0.8 * (t[g] - mean) / unit_ns;
    }
}

}  // namespace

extern "C" {

const char* mllp_last_error(void) { return g_err.c_str(); }
int mllp_version(void) { return 200; }
long long mllp_launch_count(void) { return mllp::g_launches.load(std::memory_order_relaxed); }

int mllp_device_info(int device, int64_t* out3)
{
    if (!out3) return fail(MLLP_E_INVALID, "mllp_device_info: null output");
    cudaDeviceProp p;
    CUDA_OK(cudaGetDeviceProperties(&p, device));
    out3[0] = p.multiProcessorCount;
    out3[1] = p.l2CacheSize;
    out3[2] = (int64_t)p.sharedMemPerBlockOptin;
    return 0;
}

static int create_impl(int32_t m, int32_t n, int64_t nnz, const int32_t* h_indptr, const int32_t* h_indices,
                       const double* h_values_in, const double* h_lb_in, const double* h_ub_in, const double* h_ylo,
                       const double* h_yhi, int device, uint32_t flags, int rank, int nranks, const unsigned char* uid,
                       mllp_lp_t* out)
{
    if (!out) return fail(MLLP_E_INVALID, "mllp_lp_create: null output handle");
    *out = nullptr;
    const double* h_values = h_values_in;
    const double *h_lb = h_lb_in, *h_ub = h_ub_in;
    if (m < 0 || n < 0 || nnz < 0 || !h_indptr || (nnz > 0 && (!h_indices || !h_values)))
        return fail(MLLP_E_INVALID, "mllp_lp_create: bad shape or null CSR arrays");
    if ((flags & MLLP_F_PRECONDITION) && nranks > 1)
        return fail(MLLP_E_STATE, "mllp_lp_create_rowpart: MLLP_F_PRECONDITION is not available on a row-partitioned handle");
    if (h_indptr[0] != 0 || (int64_t)h_indptr[m] != nnz)
        return fail(MLLP_E_INVALID, "mllp_lp_create: indptr[0] must be 0 and indptr[m] == nnz");
    for (int i = 0; i < m; ++i)
        if (h_indptr[i + 1] < h_indptr[i]) return fail(MLLP_E_INVALID, "mllp_lp_create: indptr not monotone");
    for (int64_t k = 0; k < nnz; ++k)
        if (h_indices[k] < 0 || h_indices[k] >= n) return fail(MLLP_E_INVALID, "mllp_lp_create: column index out of range");
    if ((h_lb == nullptr) != (h_ub == nullptr) || (h_ylo == nullptr) != (h_yhi == nullptr))
        return fail(MLLP_E_INVALID, "mllp_lp_create: lb/ub (ylo/yhi) must be given together");

    int ndev = 0;
    CUDA_OK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(MLLP_E_INVALID, "mllp_lp_create: no such CUDA device");
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(MLLP_E_STATE, "mllp_lp_create: this library is built for sm_100a (B200) only");

    mllp_lp* lp = new (std::nothrow) mllp_lp();
    if (!lp) return fail(MLLP_E_NOMEM, "mllp_lp_create: out of host memory");
    lp->device = device;
    lp->m = m; lp->n = n; lp->nnz = nnz;
    lp->bounds = (h_lb != nullptr) || (h_ylo != nullptr);
    lp->rank = rank; lp->nranks = nranks;
    lp->flags = flags;

    // launch geometry of the persistent grid.  Resident mode (default): one 1024-thread CTA per
    // SM whose share of A and A' lives in shared memory; streaming mode: 2 x 512.
    const bool resident = !(flags & (MLLP_F_NO_SMEM_RESIDENT | MLLP_F_GRAPH_MODE));
    lp->threads = env_int("MLLP_THREADS", 1024);
    if (lp->threads < 32 || lp->threads > 1024 || (lp->threads & 31)) lp->threads = 1024;
    int bpsm = persistent_max_blocks_per_sm(lp->threads, lp->bounds, 0);
    if (bpsm <= 0) { delete lp; return fail(MLLP_E_STATE, "mllp_lp_create: persistent kernel cannot be resident"); }
    const int want = env_int("MLLP_CTAS_PER_SM", 1);
    if (want > 0 && want < bpsm) bpsm = want;
    lp->G = prop.multiProcessorCount * bpsm;

    BuildParams bp;
    bp.num_ctas = lp->G;
    bp.pref_steps = env_int("MLLP_PREF_STEPS", 4);
    // Two decisions from the share of the nonzeros that sits in very long rows (> 512 entries) of A:
    //  * chunk length of split rows: 8 steps (512 entries) unless there is enough long-row work to give
    //    every warp of the grid its own 4-step chunk (then the chunk round is one gather group deep);
    //  * dealing of the regular tiles: contiguous runs per CTA (neighbouring rows share gathered sectors,
    //    which one SM's L1 then serves: ken-18 -8 %, pds-20 -1 %, pilot87 -2 %) unless the long rows
    //    dominate (osa-60: 83 % of nnz), where the least-loaded dealing balances better (+6 % otherwise).
    int64_t heavy = 0;
    for (int i = 0; i < m; ++i) {
        const int64_t len = h_indptr[i + 1] - h_indptr[i];
        if (len > 512) heavy += len;
    }
    if (bp.pref_steps < 1) bp.pref_steps = 1;
    bp.cluster = env_int("MLLP_CLUSTER", 1) != 0;
    bp.cluster_rounds = env_int("MLLP_CLUSTER_ROUNDS", 3);
    bp.contiguous = env_int("MLLP_CONTIGUOUS", 2 * heavy < nnz ? 1 : 0) != 0;
    auto set_grid = [&](int G) {   // the parameters that depend on the number of CTAs
        lp->G = G;
        bp.num_ctas = G;
        const int64_t one_round = (int64_t)G * (lp->threads / 32) * 256;
        bp.max_steps = env_int("MLLP_MAX_STEPS", 2 * heavy >= one_round ? 4 : 8);
        if (bp.max_steps < bp.pref_steps) bp.max_steps = bp.pref_steps;
        if (bp.max_steps > 1024) bp.max_steps = 1024;
    };
    const int G_grid = lp->G;
    set_grid(G_grid);

    int rc = 0;
    try {
        std::vector<int32_t> tptr, tind;
        std::vector<double> tval;
        csr_transpose(m, n, h_indptr, h_indices, h_values, tptr, tind, tval);
        // MLLP_F_PRECONDITION: Ruiz + Pock-Chambolle on the device; from here on the matrix is Dr A Dc and the box is
        // l / dc, u / dc.  The caller keeps speaking the ORIGINAL LP: vectors are scaled at the boundary
        // (load_problem / store_solution) and the KKT scalars are evaluated on the original LP (Eval*Op).
        std::vector<double> scaled_vals, h_dr, h_dc, lb_s, ub_s;
        if (flags & MLLP_F_PRECONDITION) {
            scaled_vals.assign(h_values_in, h_values_in + nnz);
            h_dr.assign((size_t)m, 1.0); h_dc.assign((size_t)n, 1.0);
            DeviceGuard pg(device);
            int prc = precondition_device(m, n, nnz, h_indptr, h_indices, scaled_vals.data(), tptr.data(), tind.data(), tval.data(),
                                          env_int("MLLP_RUIZ_ITERS", 10), h_dr.data(), h_dc.data());
            if (prc != 0) { delete lp; return prc < 1000 ? cuda_fail((cudaError_t)prc, "mllp_lp_create: preconditioning") : prc; }
            h_values = scaled_vals.data();
            csr_transpose(m, n, h_indptr, h_indices, h_values, tptr, tind, tval);
            if (h_lb) {
                lb_s.resize((size_t)n); ub_s.resize((size_t)n);
                for (int j = 0; j < n; ++j) { lb_s[j] = h_lb_in[j] / h_dc[j]; ub_s[j] = h_ub_in[j] / h_dc[j]; }
                h_lb = lb_s.data(); h_ub = ub_s.data();
            }
        }
        std::vector<int32_t> orderY, posY, orderX, posX;
        plan_orders(m, n, h_indptr, h_indices, tptr.data(), tind.data(), bp, orderY, posY, orderX, posX);
        HostMat HA, HAT;
        int mi = m, ni = n;
        if (nranks == 1) {
            build_host_mat(m, n, h_indptr, h_indices, h_values, orderY, posX, bp, HA);
            build_host_mat(n, m, tptr.data(), tind.data(), tval.data(), orderX, posY, bp, HAT);
        } else {
            // Row partition: rank p owns the rows of A (entries of y) that LPT-by-nnz gives it; A' is built whole on every
            // rank (replicated A' phase).  The internal order of y is [rank 0's rows | rank 1's | ...], each slice padded
            // to the same even length (one in-place all-gather moves it in the NCCL variant); x keeps the single-GPU order.
            const BuildParams bpA = effective_params(m, h_indptr, bp), bpAT = effective_params(n, tptr.data(), bp);
            RowPartition PY;
            partition_rows(m, h_indptr, orderY, nranks, PY);
            lp->Ly = PY.L;
            posY = PY.pos;
            const std::vector<int32_t>& mineY = PY.lists[rank];
            mi = lp->Ly * nranks;
            build_host_mat((int)mineY.size(), ni, h_indptr, h_indices, h_values, mineY, posX, bpA, HA, (uint32_t)(rank * lp->Ly));
            build_host_mat(n, mi, tptr.data(), tind.data(), tval.data(), orderX, posY, bpAT, HAT);
            orderY = PY.order_pad;
            lp->peers.rank = rank; lp->peers.nranks = nranks; lp->peers.Ly = lp->Ly; lp->peers.mi = mi;
            for (int q = 0; q < nranks && q < MAX_RANKS; ++q) lp->peers.cnt[q] = (int)PY.lists[q].size();
        }
        lp->mi = mi; lp->ni = ni;
        auto permuted_pad = [](const double* src, const std::vector<int32_t>& order, double fill) {
            std::vector<double> outv(order.size());
            for (size_t k = 0; k < order.size(); ++k) outv[k] = order[k] >= 0 ? src[order[k]] : fill;
            return outv;
        };

        auto upload_scales = [&]() -> int {   // dr / dc in the CURRENT internal order (re-done when a geometry re-orders)
            if (h_dr.empty()) return 0;
            const std::vector<double> pr = permuted_pad(h_dr.data(), orderY, 1.0), pc = permuted_pad(h_dc.data(), orderX, 1.0);
            CUDA_OK(cudaMemcpy(lp->d_dr, pr.data(), pr.size() * sizeof(double), cudaMemcpyHostToDevice));
            CUDA_OK(cudaMemcpy(lp->d_dc, pc.data(), pc.size() * sizeof(double), cudaMemcpyHostToDevice));
            return 0;
        };
        auto upload_boxes = [&]() -> int {   // general form: every box array is materialised (missing ones get +-inf / 0), in the
                                             // CURRENT internal order (re-done when a geometry re-orders the rows / columns)
            std::vector<double> lb(orderX.size(), 0.0), ub(orderX.size(), INFINITY), ylo(orderY.size(), -INFINITY), yhi(orderY.size(), INFINITY);
            if (h_lb) { lb = permuted_pad(h_lb, orderX, 0.0); ub = permuted_pad(h_ub, orderX, 0.0); }
            if (h_ylo) { ylo = permuted_pad(h_ylo, orderY, 0.0); yhi = permuted_pad(h_yhi, orderY, 0.0); }
            const std::vector<double>* src[4] = {&lb, &ub, &ylo, &yhi};
            for (int q = 0; q < 4; ++q)
                CUDA_OK(cudaMemcpy(lp->d_box[q], src[q]->data(), src[q]->size() * sizeof(double), cudaMemcpyHostToDevice));
            return 0;
        };
        auto body = [&]() -> int {
            RC_OK(upload_mat(lp, HA, lp->d.A));
            RC_OK(upload_mat(lp, HAT, lp->d.AT));
            RC_OK(dev_upload(lp, &lp->d_orderX, orderX.data(), orderX.size()));
            RC_OK(dev_upload(lp, &lp->d_orderY, orderY.data(), orderY.size()));
            lp->d.m = mi; lp->d.n = ni;
            RC_OK(dev_zeros(lp, &lp->d_b, (size_t)mi));
            RC_OK(dev_zeros(lp, &lp->d_c, (size_t)ni));
            lp->d.b = lp->d_b; lp->d.c = lp->d_c;
            if (lp->bounds) {
                for (double** q : {&lp->d_box[0], &lp->d_box[1]}) RC_OK(dev_zeros(lp, q, (size_t)ni));
                for (double** q : {&lp->d_box[2], &lp->d_box[3]}) RC_OK(dev_zeros(lp, q, (size_t)mi));
                lp->d.lb = lp->d_box[0]; lp->d.ub = lp->d_box[1]; lp->d.ylo = lp->d_box[2]; lp->d.yhi = lp->d_box[3];
                RC_OK(upload_boxes());
            }
            RC_OK(dev_zeros(lp, &lp->d.x, (size_t)ni + 2));
            RC_OK(dev_zeros(lp, &lp->d.y, (size_t)mi + 2));
            RC_OK(dev_zeros(lp, &lp->d.xbar, (size_t)ni + 2));
            RC_OK(dev_zeros(lp, &lp->d.x0, (size_t)ni));
            RC_OK(dev_zeros(lp, &lp->d.y0, (size_t)mi));
            RC_OK(dev_zeros(lp, &lp->d.red, (size_t)RED_BUFFERS * lp->G * NRED));
            RC_OK(dev_zeros(lp, &lp->d.barrier, 4));
            RC_OK(dev_zeros(lp, &lp->d.ctrl, CTRL_SIZE));
            RC_OK(dev_zeros(lp, &lp->tmp_n, (size_t)ni));
            RC_OK(dev_zeros(lp, &lp->tmp_n2, (size_t)ni));
            RC_OK(dev_zeros(lp, &lp->tmp_m, (size_t)mi));
            RC_OK(dev_zeros(lp, &lp->u_x, (size_t)n));
            RC_OK(dev_zeros(lp, &lp->u_y, (size_t)m));
            RC_OK(dev_zeros(lp, &lp->u_b, (size_t)m));
            RC_OK(dev_zeros(lp, &lp->u_c, (size_t)n));
            if (!h_dr.empty()) {
                RC_OK(dev_zeros(lp, &lp->d_dr, (size_t)mi));
                RC_OK(dev_zeros(lp, &lp->d_dc, (size_t)ni));
                lp->d.dr = lp->d_dr; lp->d.dc = lp->d_dc;
                RC_OK(upload_scales());
            }
            RC_OK(dev_zeros(lp, &lp->d_scal, MLLP_NUM_SCALARS));
            RC_OK(dev_zeros(lp, &lp->d_norm2, 2 + 256));
            if (flags & MLLP_F_GRAPH_MODE) RC_OK(build_graph(lp));
            if (nranks > 1) {
                RC_OK(dev_zeros(lp, &lp->d_mail, (size_t)4 * mi + 2));
                RC_OK(dev_zeros(lp, &lp->d_err, 4));
                NcclApi* api = nccl_api();
                if (!api) return fail(MLLP_E_STATE, "mllp_lp_create_rowpart: libnccl.so.2 could not be loaded");
                NcclId id;
                memcpy(id.internal, uid, sizeof(id.internal));
                NCCL_OK(api->comm_init_rank(&lp->comm, nranks, id, rank));
            }
            RC_OK(configure_residency(lp, HA, HAT, prop, bpsm, resident));
            return 0;
        };
        rc = body();

        // Geometry of the persistent kernels (single GPU).  The grid barrier costs ~0.9 us and there are two per
        // iteration; an LP whose phases are shorter than that runs faster on ONE thread-block cluster (<= 16 SMs,
        // hardware cluster barrier) or on one CTA.  Candidates are built and timed (48 traced iterations on the zero
        // state, like the tuning rounds), the fastest is kept.  MLLP_GEOM forces one: 0 grid, 1 one CTA, 2..16 cluster.
        // The format (row order, dealing) depends on the CTA count, the arithmetic of a row does not: results of
        // different geometries agree to rounding of the split rows' chunk sums.
        auto build_geom = [&](int G, int mode) -> int {
            set_grid(G);
            lp->d.sync_mode = mode;
            plan_orders(m, n, h_indptr, h_indices, tptr.data(), tind.data(), bp, orderY, posY, orderX, posX);
            HostMat nA, nAT;
            build_host_mat(m, n, h_indptr, h_indices, h_values, orderY, posX, bp, nA);
            build_host_mat(n, m, tptr.data(), tind.data(), tval.data(), orderX, posY, bp, nAT);
            HA = std::move(nA); HAT = std::move(nAT);
            free_mats(lp);
            RC_OK(upload_mat(lp, HA, lp->d.A));
            RC_OK(upload_mat(lp, HAT, lp->d.AT));
            CUDA_OK(cudaMemcpy(lp->d_orderX, orderX.data(), orderX.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
            CUDA_OK(cudaMemcpy(lp->d_orderY, orderY.data(), orderY.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
            RC_OK(upload_scales());
            if (lp->bounds) RC_OK(upload_boxes());
            RC_OK(configure_residency(lp, HA, HAT, prop, bpsm, resident));
            return 0;
        };
        const int tune = env_int("MLLP_TUNE", (flags & MLLP_F_NO_TUNE) ? 0 : 8);
        if (rc == 0 && nranks == 1 && resident && m > 0 && n > 0) {
            // candidate code: 0 grid, 1 one CTA, 2..16 cluster of c CTAs, 100 + c the same cluster with broadcast copies
            auto mode_of = [](int c) { return c == 0 ? SYNC_GRID : c == 1 ? SYNC_CTA : c >= 100 ? SYNC_BCAST : SYNC_CLUSTER; };
            auto ctas_of = [&](int c) { return c == 0 ? G_grid : c >= 100 ? std::min(c - 100, 16) : std::min(c, 16); };
            auto select = [&]() -> int {
                const int forced = env_int("MLLP_GEOM", -1);
                if (forced == 0) return 0;
                if (forced > 0) return build_geom(ctas_of(forced), mode_of(forced));
                if (tune <= 0 || nnz > (int64_t)env_int("MLLP_GEOM_MAX_NNZ", 200000)) return 0;
                PhaseTimes T;
                RC_OK(measure_phases(lp, T));
                double best = T.iter_ns;
                int best_c = 0, cur = 0;
                lp->geom_ns[0] = T.iter_ns;
                const int cands[8] = {16, 8, 4, 1, 116, 108, 104, 101};
                for (int q = 0; q < 8; ++q) {
                    const int c = cands[q];
                    if ((c == 1 || c == 101) && nnz > 30000) continue;
                    if (c > 1 && c != 101 && persistent_cluster_fits(ctas_of(c), lp->threads, lp->bounds, 0) != 1) continue;
                    if (build_geom(ctas_of(c), mode_of(c)) != 0) { cur = -1; continue; }   // does not fit: skip the candidate
                    cur = c;
                    RC_OK(measure_phases(lp, T));
                    lp->geom_ns[q + 1] = T.iter_ns;
                    if (T.iter_ns < 0.97 * best) { best = T.iter_ns; best_c = c; }
                }
                if (cur != best_c) RC_OK(build_geom(ctas_of(best_c), mode_of(best_c)));
                return 0;
            };
            rc = select();
        }

        // Tuning rounds (single GPU, persistent kernel): trace a few iterations on the zero state, feed the
        // per-CTA phase times back into the dealing of the regular tiles, rebuild, keep the fastest build.
        // The dealing does not change the summation order inside a row, so results are unaffected.
        const bool worth = (int64_t)HA.tiles.size() + (int64_t)HAT.tiles.size() >= 8 * (int64_t)lp->G;
        if (rc == 0 && tune > 0 && nranks == 1 && !(flags & MLLP_F_GRAPH_MODE) && worth) {
            auto tuning = [&]() -> int {
                PhaseTimes T;
                RC_OK(measure_phases(lp, T));
                lp->tune_ns[0] = lp->tune_ns[1] = T.iter_ns;
                DealFeedback fbA, fbAT;
                HostMat bestA, bestAT, curA, curAT;   // best = fastest build so far (if not the first), cur = uploaded
                bool have_best = false, cur_is_first = true, uploaded_is_best = true;
                int stale = 0;   // rounds since the last improvement: stop after 3
                for (int r = 0; r < tune && stale < 3; ++r) {
                    update_feedback(cur_is_first ? HA : curA, T.a, bp.contiguous, fbA);
                    update_feedback(cur_is_first ? HAT : curAT, T.at, bp.contiguous, fbAT);
                    HostMat nA, nAT;
                    build_host_mat(m, n, h_indptr, h_indices, h_values, orderY, posX, bp, nA, 0, &fbA);
                    build_host_mat(n, m, tptr.data(), tind.data(), tval.data(), orderX, posY, bp, nAT, 0, &fbAT);
                    free_mats(lp);
                    RC_OK(upload_mat(lp, nA, lp->d.A));
                    RC_OK(upload_mat(lp, nAT, lp->d.AT));
                    RC_OK(configure_residency(lp, nA, nAT, prop, bpsm, resident));
                    RC_OK(measure_phases(lp, T));
                    ++lp->tune_rounds;
                    curA = std::move(nA); curAT = std::move(nAT);
                    cur_is_first = false;
                    if (T.iter_ns < 0.995 * lp->tune_ns[1]) {
                        lp->tune_ns[1] = T.iter_ns;
                        bestA = curA; bestAT = curAT;
                        have_best = true; uploaded_is_best = true;
                        stale = 0;
                    } else {
                        uploaded_is_best = false;
                        ++stale;
                    }
                }
                if (have_best) { HA = std::move(bestA); HAT = std::move(bestAT); }
                if (!uploaded_is_best) {
                    free_mats(lp);
                    RC_OK(upload_mat(lp, HA, lp->d.A));
                    RC_OK(upload_mat(lp, HAT, lp->d.AT));
                    RC_OK(configure_residency(lp, HA, HAT, prop, bpsm, resident));
                }
                return 0;
            };
            rc = tuning();
        }

        // Block-angular LPs (ken-18: 475 independent blocks + 151 linking rows): the block kernel of blocks.cu keeps every
        // CTA's blocks in shared memory and needs no grid barrier.  Built when the structure is there (single GPU,
        // cooperative grid), timed against the grid kernel on the zero state, kept when faster.
        // MLLP_BLOCKS = 0 / 1 disables / forces it.
        if (rc == 0 && nranks == 1 && resident && lp->d.sync_mode == SYNC_GRID && m > 0 && n > 0 &&
            env_int("MLLP_BLOCKS", tune > 0 ? -1 : 0) != 0) {
            auto pick = [&]() -> int {
                RC_OK(blocks_create(m, n, h_indptr, h_indices, h_values, posX.data(), posY.data(), lp->d.lb, lp->d.ub, lp->d.ylo, lp->d.yhi,
                                    device, prop.multiProcessorCount, &lp->blocks));
                if (!lp->blocks) return 0;
                if (env_int("MLLP_BLOCKS", -1) == 1) { lp->use_blocks = true; return 0; }
                double t_grid = 0, t_blk = 0;
                lp->use_blocks = false;
                RC_OK(time_parity(lp, 200, t_grid));
                lp->use_blocks = true;
                // the block kernel's work per CTA is small: fewer threads mean cheaper CTA barriers, more threads fewer rounds
                int best_threads = 1024;
                if (getenv("MLLP_BLOCKS_THREADS") == nullptr) {
                    for (int th : {1024, 512}) {
                        double t = 0;
                        blocks_set_threads(lp->blocks, th);
                        RC_OK(time_parity(lp, 200, t));
                        if (t_blk == 0 || t < t_blk) { t_blk = t; best_threads = th; }
                    }
                    blocks_set_threads(lp->blocks, best_threads);
                } else {
                    RC_OK(time_parity(lp, 200, t_blk));
                }
                lp->blocks_ns[0] = t_grid; lp->blocks_ns[1] = t_blk;
                lp->use_blocks = t_blk < 0.97 * t_grid;
                return 0;
            };
            rc = pick();
        }

        int64_t* I = lp->info;
        I[0] = m; I[1] = n; I[2] = nranks > 1 ? HA.nnz_emitted : nnz;
        I[3] = (int64_t)HA.tiles.size(); I[4] = (int64_t)HAT.tiles.size();
        I[5] = (int64_t)HA.total_steps * 64; I[6] = (int64_t)HAT.total_steps * 64;
        I[7] = (int64_t)HA.splits.size(); I[8] = (int64_t)HAT.splits.size();
        I[9] = lp->G; I[10] = lp->threads; I[11] = (int64_t)lp->dyn_smem;
        I[12] = 24 * nnz + 36 * (int64_t)m + 44 * (int64_t)n + 8 + (h_lb ? 16 * (int64_t)n : 0) + (h_ylo ? 16 * (int64_t)m : 0);
        I[13] = lp->d.res_steps_A; I[14] = lp->d.res_steps_AT; I[15] = bpsm;
    } catch (const std::bad_alloc&) {
        rc = fail(MLLP_E_NOMEM, "mllp_lp_create: out of host memory");
    }
    if (rc != 0) {
        const std::string keep = g_err;
        mllp_lp_destroy(lp);
        g_err = keep;
        return rc;
    }
    *out = lp;
    return 0;
}

int mllp_lp_create(int32_t m, int32_t n, int64_t nnz, const int32_t* h_indptr, const int32_t* h_indices,
                   const double* h_values, const double* h_lb, const double* h_ub, const double* h_ylo,
                   const double* h_yhi, int device, uint32_t flags, mllp_lp_t* out)
{
    return create_impl(m, n, nnz, h_indptr, h_indices, h_values, h_lb, h_ub, h_ylo, h_yhi, device, flags, 0, 1, nullptr, out);
}

int mllp_nccl_unique_id(unsigned char* out128)
{
    if (!out128) return fail(MLLP_E_INVALID, "mllp_nccl_unique_id: null output");
    NcclApi* api = nccl_api();
    if (!api) return fail(MLLP_E_STATE, "mllp_nccl_unique_id: libnccl.so.2 could not be loaded");
    NcclId id;
    NCCL_OK(api->get_unique_id(&id));
    memcpy(out128, id.internal, sizeof(id.internal));
    return 0;
}

int mllp_lp_create_rowpart(int32_t m, int32_t n, int64_t nnz, const int32_t* h_indptr, const int32_t* h_indices,
                           const double* h_values, const double* h_lb, const double* h_ub, const double* h_ylo,
                           const double* h_yhi, int device, uint32_t flags, int32_t rank, int32_t nranks,
                           const unsigned char* uid128, mllp_lp_t* out)
{
    if (nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !uid128))
        return fail(MLLP_E_INVALID, "mllp_lp_create_rowpart: bad rank / nranks / unique id");
    return create_impl(m, n, nnz, h_indptr, h_indices, h_values, h_lb, h_ub, h_ylo, h_yhi, device, flags, rank, nranks,
                       uid128, out);
}

int mllp_rowpart_ipc_export(mllp_lp_t lp, unsigned char* out64)
{
    if (!lp || !out64 || lp->nranks < 2) return fail(MLLP_E_INVALID, "mllp_rowpart_ipc_export: not a row-partitioned handle");
    DeviceGuard guard(lp->device);
    cudaIpcMemHandle_t h;
    CUDA_OK(cudaIpcGetMemHandle(&h, lp->d_mail));
    static_assert(sizeof(cudaIpcMemHandle_t) == MLLP_IPC_HANDLE_BYTES, "IPC handle size");
    memcpy(out64, &h, sizeof(h));
    return 0;
}

int mllp_rowpart_ipc_import(mllp_lp_t lp, const unsigned char* all)
{
    if (!lp || !all || lp->nranks < 2) return fail(MLLP_E_INVALID, "mllp_rowpart_ipc_import: not a row-partitioned handle");
    if (lp->nranks > MAX_RANKS) return fail(MLLP_E_STATE, "mllp_rowpart_ipc_import: at most 8 ranks");
    DeviceGuard guard(lp->device);
    PeerInfo& P = lp->peers;
    P.err = lp->d_err;
    P.backoff_ns = env_int("MLLP_MAIL_BACKOFF", 40);       // measured on 4 GPUs: 40 / 100..400 / 250..2000 ns make no difference
    P.backoff_max_ns = env_int("MLLP_MAIL_BACKOFF_MAX", 40);
    P.st_mode = env_int("MLLP_MAIL_ST", 0);                // dev knob; relaxed.sys / weak / .cg / volatile stores: no difference either
    P.mc = nullptr;
    for (int q = 0; q < lp->nranks; ++q) {
        if (q == lp->rank) {
            P.mail[q] = lp->d_mail;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, all + (size_t)q * MLLP_IPC_HANDLE_BYTES, sizeof(h));
        void* ptr = nullptr;
        CUDA_OK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        lp->ipc_opened.push_back(ptr);
        P.mail[q] = (unsigned long long*)ptr;
    }
    if (lp->dyn_smem > 48 * 1024 - 4096) RC_OK(rowpart_set_smem(lp->bounds, lp->dyn_smem));
    lp->p2p_ready = true;
    return 0;
}

// NVSwitch multicast for the in-kernel exchange (see include/mllp_b200.h)
static size_t mailbox_bytes(const mllp_lp* lp) { return sizeof(unsigned long long) * ((size_t)4 * lp->mi + 2); }

int mllp_rowpart_mc_supported(mllp_lp_t lp, int32_t* out)
{
    if (!lp || !out || lp->nranks < 2) return fail(MLLP_E_INVALID, "mllp_rowpart_mc_supported: not a row-partitioned handle");
    DeviceGuard guard(lp->device);
    int sup = 0;
    RC_OK(mc_supported(lp->device, &sup));
    *out = sup;
    return 0;
}

int mllp_rowpart_mc_create(mllp_lp_t lp, int32_t* out_fd)
{
    if (!lp || !out_fd || lp->nranks < 2 || lp->rank != 0) return fail(MLLP_E_INVALID, "mllp_rowpart_mc_create: rank 0 of a row-partitioned handle only");
    DeviceGuard guard(lp->device);
    int fd = -1;
    RC_OK(mc_create(&lp->mcs, lp->nranks, mailbox_bytes(lp), &fd));
    *out_fd = fd;
    return 0;
}

int mllp_rowpart_mc_attach(mllp_lp_t lp, int32_t fd)
{
    if (!lp || lp->nranks < 2) return fail(MLLP_E_INVALID, "mllp_rowpart_mc_attach: not a row-partitioned handle");
    DeviceGuard guard(lp->device);
    if (lp->rank != 0) RC_OK(mc_import(&lp->mcs, lp->nranks, mailbox_bytes(lp), fd));
    if (!lp->mcs.have_mc) return fail(MLLP_E_STATE, "mllp_rowpart_mc_attach: no multicast object (mllp_rowpart_mc_create first on rank 0)");
    RC_OK(mc_add_device(&lp->mcs, lp->device));
    return 0;
}

int mllp_rowpart_mc_bind(mllp_lp_t lp)
{
    if (!lp || lp->nranks < 2 || !lp->mcs.have_mc) return fail(MLLP_E_INVALID, "mllp_rowpart_mc_bind: attach first");
    DeviceGuard guard(lp->device);
    RC_OK(mc_bind_map(&lp->mcs));
    PeerInfo& P = lp->peers;
    P.err = lp->d_err;
    P.backoff_ns = env_int("MLLP_MAIL_BACKOFF", 40);
    P.backoff_max_ns = env_int("MLLP_MAIL_BACKOFF_MAX", 40);
    P.st_mode = 0;
    for (int q = 0; q < MAX_RANKS; ++q) P.mail[q] = nullptr;
    P.mail[lp->rank] = (unsigned long long*)lp->mcs.uc;     // the local polls read this rank's own (unicast) mapping
    P.mc = (unsigned long long*)lp->mcs.mcva;
    if (lp->dyn_smem > 48 * 1024 - 4096) RC_OK(rowpart_set_smem(lp->bounds, lp->dyn_smem));
    lp->p2p_ready = true;
    return 0;
}

int mllp_graph_edges(int32_t m, int64_t nnz, const int32_t* d_indptr, const int32_t* d_indices, const double* d_values,
                     int64_t* d_edge_index, float* d_edge_attr, void* stream)
{
    if (m < 0 || nnz < 0 || !d_indptr || (nnz > 0 && (!d_indices || !d_values || !d_edge_index || !d_edge_attr)))
        return fail(MLLP_E_INVALID, "mllp_graph_edges: bad argument");
    RC_OK(launch_graph_edges(m, nnz, d_indptr, d_indices, d_values, (long long*)d_edge_index, d_edge_attr, (cudaStream_t)stream));
    return 0;
}

int mllp_rowpart_error(mllp_lp_t lp, int32_t* out_flag)
{
    if (!lp || !out_flag) return fail(MLLP_E_INVALID, "mllp_rowpart_error: null argument");
    *out_flag = 0;
    if (!lp->d_err) return 0;
    DeviceGuard guard(lp->device);
    unsigned v = 0;
    CUDA_OK(cudaMemcpy(&v, lp->d_err, sizeof(v), cudaMemcpyDeviceToHost));
    *out_flag = (int32_t)v;
    return 0;
}

int mllp_lp_destroy(mllp_lp_t lp)
{
    if (!lp) return 0;
    DeviceGuard guard(lp->device);
    for (void* p : lp->ipc_opened) cudaIpcCloseMemHandle(p);
    mc_destroy(&lp->mcs);
    if (lp->comm) { NcclApi* api = nccl_api(); if (api) api->comm_destroy(lp->comm); }
    if (lp->graph) cudaGraphExecDestroy(lp->graph);
    blocks_destroy(lp->blocks);
    for (void* p : lp->allocs) cudaFree(p);
    for (void* p : lp->mat_allocs) cudaFree(p);
    delete lp;
    return 0;
}

int mllp_lp_info(mllp_lp_t lp, int64_t* out16)
{
    if (!lp || !out16) return fail(MLLP_E_INVALID, "mllp_lp_info: null argument");
    memcpy(out16, lp->info, sizeof(lp->info));
    return 0;
}

int mllp_lp_tune_info(mllp_lp_t lp, double* out4)
{
    if (!lp || !out4) return fail(MLLP_E_INVALID, "mllp_lp_tune_info: null argument");
    out4[0] = lp->tune_ns[0]; out4[1] = lp->tune_ns[1]; out4[2] = (double)lp->tune_rounds; out4[3] = 0.0;
    return 0;
}

int mllp_lp_blocks_info(mllp_lp_t lp, double* out8)
{
    if (!lp || !out8) return fail(MLLP_E_INVALID, "mllp_lp_blocks_info: null argument");
    int64_t b[4];
    blocks_info(lp->blocks, b);
    out8[0] = lp->use_blocks ? 1.0 : 0.0; out8[1] = (double)b[0]; out8[2] = (double)b[1]; out8[3] = (double)b[2];
    out8[4] = (double)(b[3] & 0xffffffffll); out8[5] = lp->blocks_ns[0]; out8[6] = lp->blocks_ns[1]; out8[7] = lp->blocks ? 1.0 : 0.0;
    return 0;
}

int mllp_lp_geometry(mllp_lp_t lp, double* out12)
{
    if (!lp || !out12) return fail(MLLP_E_INVALID, "mllp_lp_geometry: null argument");
    out12[0] = (double)lp->d.sync_mode; out12[1] = (double)lp->G;
    for (int k = 0; k < 9; ++k) out12[2 + k] = lp->geom_ns[k];
    out12[11] = 0.0;
    return 0;
}

int mllp_lp_scaling(mllp_lp_t lp, double* d_dr, double* d_dc, void* stream)
{
    if (!lp || !d_dr || !d_dc) return fail(MLLP_E_INVALID, "mllp_lp_scaling: null argument");
    DeviceGuard guard(lp->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (!lp->d_dr) {
        RC_OK(launch_fill(d_dr, 1.0, lp->m, s));
        RC_OK(launch_fill(d_dc, 1.0, lp->n, s));
        return 0;
    }
    RC_OK(launch_scatter(d_dr, lp->d_dr, lp->d_orderY, lp->mi, s));
    RC_OK(launch_scatter(d_dc, lp->d_dc, lp->d_orderX, lp->ni, s));
    return 0;
}

int mllp_spmv(mllp_lp_t lp, int trans, const double* d_in, double* d_out, void* stream)
{
    if (!lp || !d_in || !d_out) return fail(MLLP_E_INVALID, "mllp_spmv: null argument");
    if (lp->nranks > 1) return fail(MLLP_E_STATE, "mllp_spmv: not available on a row-partitioned handle");
    DeviceGuard guard(lp->device);
    cudaStream_t s = (cudaStream_t)stream;
    // products with the ORIGINAL matrix also on a preconditioned handle: A = Dr^-1 (Dr A Dc) Dc^-1
    if (!trans) {
        RC_OK(launch_gather_scaled(lp->tmp_n, d_in, lp->d_orderX, lp->d_dc, 1, lp->n, s));
        RC_OK(launch_spmv(lp->d.A, lp->tmp_n, lp->tmp_m, lp->G, lp->threads, s));
        RC_OK(launch_scatter_scaled(d_out, lp->tmp_m, lp->d_orderY, lp->d_dr, 1, lp->m, s));
    } else {
        RC_OK(launch_gather_scaled(lp->tmp_m, d_in, lp->d_orderY, lp->d_dr, 1, lp->m, s));
        RC_OK(launch_spmv(lp->d.AT, lp->tmp_m, lp->tmp_n, lp->G, lp->threads, s));
        RC_OK(launch_scatter_scaled(d_out, lp->tmp_n, lp->d_orderX, lp->d_dc, 1, lp->n, s));
    }
    return 0;
}

int mllp_estimate_norm(mllp_lp_t lp, int iters, double* h_sigma_max, void* stream)
{
    if (!lp || !h_sigma_max || iters < 1) return fail(MLLP_E_INVALID, "mllp_estimate_norm: bad argument");
    if (lp->nranks > 1) return fail(MLLP_E_STATE, "mllp_estimate_norm: not available on a row-partitioned handle");
    DeviceGuard guard(lp->device);
    cudaStream_t s = (cudaStream_t)stream;
    // v (tmp_n) = 1/sqrt(n); order does not matter for a constant vector
    RC_OK(launch_fill(lp->tmp_n, lp->n > 0 ? 1.0 / sqrt((double)lp->n) : 0.0, lp->n, s));
    for (int it = 0; it < iters; ++it) {
        RC_OK(launch_spmv(lp->d.A, lp->tmp_n, lp->tmp_m, lp->G, lp->threads, s));
        RC_OK(launch_spmv(lp->d.AT, lp->tmp_m, lp->tmp_n2, lp->G, lp->threads, s));
        RC_OK(launch_sumsq(lp->tmp_n2, lp->n, lp->d_norm2, lp->d_norm2 + 2, s));
        RC_OK(launch_scale_by_invnorm(lp->tmp_n, lp->tmp_n2, lp->d_norm2, lp->n, s));
    }
    double nz2 = 0.0;
    CUDA_OK(cudaMemcpyAsync(&nz2, lp->d_norm2, sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaStreamSynchronize(s));
    *h_sigma_max = sqrt(sqrt(nz2));
    return 0;
}

static int load_problem(mllp_lp* lp, const double* d_x, const double* d_y, const double* d_b, const double* d_c,
                        cudaStream_t s)
{
    // preconditioned handle: x~ = x / dc, y~ = y / dr, b~ = dr b, c~ = dc c (the caller's vectors are the original LP's)
    RC_OK(launch_gather_scaled(lp->d.x, d_x, lp->d_orderX, lp->d_dc, 1, lp->ni, s));
    RC_OK(launch_gather_scaled(lp->d.y, d_y, lp->d_orderY, lp->d_dr, 1, lp->mi, s));
    RC_OK(launch_gather_scaled(lp->d_b, d_b, lp->d_orderY, lp->d_dr, 0, lp->mi, s));
    RC_OK(launch_gather_scaled(lp->d_c, d_c, lp->d_orderX, lp->d_dc, 0, lp->ni, s));
    return 0;
}
static int store_solution(mllp_lp* lp, double* d_x, double* d_y, cudaStream_t s)
{
    RC_OK(launch_scatter_scaled(d_x, lp->d.x, lp->d_orderX, lp->d_dc, 0, lp->ni, s));   // x = dc x~
    RC_OK(launch_scatter_scaled(d_y, lp->d.y, lp->d_orderY, lp->d_dr, 0, lp->mi, s));   // y = dr y~
    return 0;
}

int mllp_pdhg_run(mllp_lp_t lp, double* d_x, double* d_y, const double* d_b, const double* d_c, double tau,
                  double sigma, int32_t num_iters, double* d_scalars, void* stream)
{
    if (!lp || !d_x || !d_y || !d_b || !d_c || num_iters < 0)
        return fail(MLLP_E_INVALID, "mllp_pdhg_run: null argument or negative iteration count");
    DeviceGuard guard(lp->device);
    cudaStream_t s = (cudaStream_t)stream;
    RC_OK(load_problem(lp, d_x, d_y, d_b, d_c, s));
    if (lp->nranks > 1) {
        // row partition: whole-x update (replicated) | y-slice update | exchange of the y slices, all on `s`
        NcclApi* api = nccl_api();
        const double ts[2] = {tau, sigma};
        CUDA_OK(cudaMemcpyAsync(lp->d.ctrl, ts, sizeof(ts), cudaMemcpyHostToDevice, s));
        const bool p2p = lp->p2p_ready && env_int("MLLP_ROWPART_NCCL", 0) == 0;
        if (p2p && num_iters > 0) {
            // all iterations in ONE cooperative launch per rank; exchange = tagged words through peer mailboxes
            lp->d.join_base = lp->join_epoch;
            lp->join_epoch += 2ull * (unsigned long long)num_iters + 2ull;
            RC_OK(launch_pdhg_rowpart(lp->d, lp->peers, lp->bounds, lp->G, lp->threads, lp->dyn_smem, tau, sigma, num_iters,
                                      lp->xseq, s));
            lp->xseq += (unsigned long long)num_iters;
        }
        for (int it = 0; it < (p2p ? 0 : num_iters); ++it) {
            RC_OK(launch_primal(lp->d, lp->bounds, lp->G, lp->threads, s));
            RC_OK(launch_dual(lp->d, lp->bounds, lp->G, lp->threads, s));
            NCCL_OK(api->all_gather(lp->d.y + (size_t)lp->rank * lp->Ly, lp->d.y, (size_t)lp->Ly, NCCL_FLOAT64, lp->comm, s));
        }
        if (d_scalars) {
            // the A' side sums are complete on every rank (replicated phase); the A side sums are per rank
            RC_OK(launch_eval_partial(lp->d, lp->bounds, lp->G, lp->threads, s));
            double* red = lp->d.red + (size_t)RED_EVALD * lp->G * NRED;
            NCCL_OK(api->all_reduce(red, red, (size_t)lp->G * NRED, NCCL_FLOAT64, NCCL_SUM, lp->comm, s));
            RC_OK(launch_eval_finalize(lp->d, lp->G, d_scalars, (double)num_iters, s));
        }
        RC_OK(store_solution(lp, d_x, d_y, s));
        return 0;
    }
    if (lp->flags & MLLP_F_GRAPH_MODE) {
        const double ts[2] = {tau, sigma};
        CUDA_OK(cudaMemcpyAsync(lp->d.ctrl, ts, sizeof(ts), cudaMemcpyHostToDevice, s));
        int left = num_iters;
        for (; left >= GRAPH_UNROLL; left -= GRAPH_UNROLL) { count_launch(2 * GRAPH_UNROLL); CUDA_OK(cudaGraphLaunch(lp->graph, s)); }
        for (; left > 0; --left) {
            RC_OK(launch_primal(lp->d, lp->bounds, lp->G, lp->threads, s));
            RC_OK(launch_dual(lp->d, lp->bounds, lp->G, lp->threads, s));
        }
    } else if (num_iters > 0) {
        RC_OK(launch_parity(lp, tau, sigma, num_iters, s));
    }
    if (d_scalars) RC_OK(launch_eval(lp->d, lp->bounds, lp->G, lp->threads, d_scalars, (double)num_iters, s));
    RC_OK(store_solution(lp, d_x, d_y, s));
    return 0;
}

// Dev tool (not in the public header): run `iters` parity iterations on the current internal
// state with barrier tracing; h_out receives iters*G*4 timestamps (ns): per CTA
// [A' phase done, barrier released, A phase done, barrier released].
int mllp_debug_trace(mllp_lp_t lp, double tau, double sigma, int32_t iters, unsigned long long* h_out)
{
    if (!lp || !h_out || iters < 1) return fail(MLLP_E_INVALID, "mllp_debug_trace: bad argument");
    DeviceGuard guard(lp->device);
    {
        std::vector<unsigned long long> tr;
        RC_OK(run_trace(lp, tau, sigma, iters, tr));
        memcpy(h_out, tr.data(), tr.size() * sizeof(unsigned long long));
        return 0;
    }
}

// Dev tool (not in the public header), collective: `iters` traced iterations of the row-partitioned kernel on the current
// internal state; h_out receives iters*G*6 timestamps (ns) per CTA: [A' phase done, barrier released, A phase done,
// unpack done, barrier released, unused].
int mllp_debug_trace_rowpart(mllp_lp_t lp, double tau, double sigma, int32_t iters, unsigned long long* h_out)
{
    if (!lp || !h_out || iters < 1 || lp->nranks < 2 || !lp->p2p_ready) return fail(MLLP_E_INVALID, "mllp_debug_trace_rowpart: bad argument");
    DeviceGuard guard(lp->device);
    unsigned long long* d_tr = nullptr;
    const size_t cnt = (size_t)iters * lp->G * 6;
    CUDA_OK(cudaMalloc(&d_tr, cnt * sizeof(unsigned long long)));
    int rc = (int)cudaMemset(d_tr, 0, cnt * sizeof(unsigned long long));
    lp->d.join_base = lp->join_epoch;
    lp->join_epoch += 2ull * (unsigned long long)iters + 2ull;
    DevLP d = lp->d;
    d.trace = d_tr;
    if (rc == 0) rc = launch_pdhg_rowpart(d, lp->peers, lp->bounds, lp->G, lp->threads, lp->dyn_smem, tau, sigma, iters, lp->xseq, 0);
    lp->xseq += (unsigned long long)iters;
    if (rc == 0) rc = (int)cudaDeviceSynchronize();
    if (rc == 0) rc = (int)cudaMemcpy(h_out, d_tr, cnt * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(d_tr);
    if (rc != 0) return cuda_fail((cudaError_t)rc, "traced row-partitioned run");
    return 0;
}

int mllp_pdhg_run_host(mllp_lp_t lp, double* h_x, double* h_y, const double* h_b, const double* h_c, double tau,
                       double sigma, int32_t num_iters, double* h_scalars, void* stream)
{
    if (!lp || !h_x || !h_y || !h_b || !h_c) return fail(MLLP_E_INVALID, "mllp_pdhg_run_host: null argument");
    DeviceGuard guard(lp->device);
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_OK(cudaMemcpyAsync(lp->u_x, h_x, sizeof(double) * lp->n, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaMemcpyAsync(lp->u_y, h_y, sizeof(double) * lp->m, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaMemcpyAsync(lp->u_b, h_b, sizeof(double) * lp->m, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaMemcpyAsync(lp->u_c, h_c, sizeof(double) * lp->n, cudaMemcpyHostToDevice, s));
    RC_OK(mllp_pdhg_run(lp, lp->u_x, lp->u_y, lp->u_b, lp->u_c, tau, sigma, num_iters,
                        h_scalars ? lp->d_scal : nullptr, stream));
    CUDA_OK(cudaMemcpyAsync(h_x, lp->u_x, sizeof(double) * lp->n, cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaMemcpyAsync(h_y, lp->u_y, sizeof(double) * lp->m, cudaMemcpyDeviceToHost, s));
    if (h_scalars)
        CUDA_OK(cudaMemcpyAsync(h_scalars, lp->d_scal, sizeof(double) * MLLP_NUM_SCALARS, cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaStreamSynchronize(s));
    return 0;
}

int mllp_pdhg_solve(mllp_lp_t lp, double* d_x, double* d_y, const double* d_b, const double* d_c, double eta,
                    double w0, int32_t max_iters, int32_t check_every, double tol, double* d_scalars, void* stream)
{
    if (!lp || !d_x || !d_y || !d_b || !d_c || !d_scalars || max_iters < 0 || check_every < 1 || !(w0 >= 0.0) ||
        !(eta > 0.0))
        return fail(MLLP_E_INVALID, "mllp_pdhg_solve: bad argument");
    if (lp->nranks > 1) return fail(MLLP_E_STATE, "mllp_pdhg_solve: not available on a row-partitioned handle");
    if (lp->flags & MLLP_F_GRAPH_MODE) return fail(MLLP_E_STATE, "mllp_pdhg_solve: needs the persistent kernel (handle was created with MLLP_F_GRAPH_MODE)");
    DeviceGuard guard(lp->device);
    cudaStream_t s = (cudaStream_t)stream;
    RC_OK(load_problem(lp, d_x, d_y, d_b, d_c, s));
    // scalars of the starting point (also what is returned when max_iters == 0)
    RC_OK(launch_eval(lp->d, lp->bounds, lp->G, lp->threads, d_scalars, 0.0, s));
    lp->d.join_base = lp->join_epoch;
    lp->join_epoch += 2ull * (unsigned long long)max_iters + 2ull;
    const double* w0_dev = nullptr;
    if (w0 == 0.0 && max_iters > 0) {
        // the PDLP default initial primal weight ||c~|| / ||b~|| of the LP the handle iterates on (scaled when preconditioned):
        // squared norms by the fixed-order two-stage sum, read by the kernel (no host synchronisation)
        RC_OK(launch_sumsq(lp->d_b, lp->mi, lp->d_norm2, lp->d_norm2 + 2, s));
        RC_OK(launch_sumsq(lp->d_c, lp->ni, lp->d_norm2 + 1, lp->d_norm2 + 2, s));
        w0_dev = lp->d_norm2;
    }
    if (max_iters > 0)
        RC_OK(launch_solve_persistent(lp->d, lp->bounds, lp->G, lp->threads, lp->dyn_smem, eta, w0 > 0.0 ? w0 : 1.0, max_iters,
                                      check_every, tol, d_scalars, w0_dev, s));
    RC_OK(store_solution(lp, d_x, d_y, s));
    return 0;
}

}  // extern "C"
