// batch_kernels.cu -- batched multi-instance mode: many small LPs in one launch.
//
// One CTA owns one LP instance at a time (grid-stride over instances): the instance's
// iterates (x, xbar, y) and data (b, c) live in the CTA's shared memory in internal order, the
// matrix (tiled format built for a one-CTA grid) is read through the read-only L1 path and,
// with no grid barrier, stays cached across iterations.  Phases are separated by
// __syncthreads() only; every instance has its own step sizes, restart state and termination
// flag.  With `shared_matrix` all instances use one matrix image (the perturbed-b/c workload of
// BASELINE.json configs[4]).
//
// Mirrors the per-instance loop of the reference driver, linear_program_experiment.py:123
// (`for name, constrs, constr_weights, coefs, rhs, basis_opt in train_dataset:`); the iteration
// itself is the frozen spec of oracle/pdhg_oracle.c (the reference has none).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mllp_b200.h"
#include "lp_format.h"
#include "pdhg_kernels.cuh"

using namespace mllp;

namespace mllp {

struct BatchInst {          // device-side descriptor of one instance
    DevMat A, AT;
    const int32_t* orderX;  // internal position k holds original column orderX[k]
    const int32_t* orderY;
    int m, n;
    long long x_off, y_off; // offsets into the concatenated user vectors
};

// shared-memory vectors of one instance
struct BatchSmem {
    double *x, *xbar, *y, *b, *c, *x0, *y0;
    double* red;            // 32 * NRED scratch
    double* bc;             // 16 broadcast slots
};

__device__ __forceinline__ BatchSmem carve(unsigned char* dsm, int m, int n, bool anchors)
{
    BatchSmem S;
    double* p = reinterpret_cast<double*>(dsm);
    S.red = p; p += 32 * NRED;
    S.bc = p; p += 16;
    S.x = p; p += n; S.xbar = p; p += n; S.c = p; p += n;
    S.y = p; p += m; S.b = p; p += m;
    S.x0 = anchors ? p : nullptr; if (anchors) p += n;
    S.y0 = anchors ? p : nullptr;
    return S;
}

__device__ __forceinline__ DevLP smem_lp(const BatchInst& I, const BatchSmem& S)
{
    DevLP lp{};
    lp.A = I.A; lp.AT = I.AT; lp.m = I.m; lp.n = I.n;
    lp.b = S.b; lp.c = S.c; lp.x = S.x; lp.y = S.y; lp.xbar = S.xbar; lp.x0 = S.x0; lp.y0 = S.y0;
    return lp;
}

// Sum acc[0..N) over the CTA (fixed order); result in every thread.
template <int N>
__device__ __forceinline__ void cta_allreduce(double* acc, const BatchSmem& S)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        double v = acc[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        if (lane == 0) S.red[warp * NRED + k] = v;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            double v = lane < nwarps ? S.red[lane * NRED + k] : 0.0;
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
            if (lane == 0) S.bc[k] = v;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < N; ++k) acc[k] = S.bc[k];
    __syncthreads();
}

__device__ __forceinline__ void load_instance(const BatchInst& I, const BatchSmem& S, const double* x, const double* y,
                                              const double* b, const double* c)
{
    for (int k = threadIdx.x; k < I.n; k += blockDim.x) {
        const int j = __ldg(I.orderX + k);
        S.x[k] = x[I.x_off + j];
        S.c[k] = c[I.x_off + j];
        S.xbar[k] = 0.0;
    }
    for (int k = threadIdx.x; k < I.m; k += blockDim.x) {
        const int i = __ldg(I.orderY + k);
        S.y[k] = y[I.y_off + i];
        S.b[k] = b[I.y_off + i];
    }
    __syncthreads();
}
__device__ __forceinline__ void store_instance(const BatchInst& I, const BatchSmem& S, double* x, double* y)
{
    for (int k = threadIdx.x; k < I.n; k += blockDim.x) x[I.x_off + __ldg(I.orderX + k)] = S.x[k];
    for (int k = threadIdx.x; k < I.m; k += blockDim.x) y[I.y_off + __ldg(I.orderY + k)] = S.y[k];
    __syncthreads();
}

// KKT scalars of (S.x, S.y) -> s[0..9] in every thread; also ||x-x0||^2, ||y-y0||^2 in dd[0..1]
__device__ __forceinline__ void batch_kkt(const DevLP& lp, const MatView& VA, const MatView& VAT, const BatchSmem& S,
                                          double* s, double* dd)
{
    double ap[NRED], ad[NRED];
#pragma unroll
    for (int k = 0; k < NRED; ++k) { ap[k] = 0.0; ad[k] = 0.0; }
    {
        EvalPrimalOp<false, SmemMem> op{lp};
        run_phase(lp.AT, VAT, op, ap);
    }
    {
        EvalDualOp<false, SmemMem> op{lp};
        run_phase(lp.A, VA, op, ad);
    }
    cta_allreduce<6>(ap, S);
    cta_allreduce<5>(ad, S);
    const double pobj = ap[0], dobj = ad[0] + ap[1];
    s[0] = pobj; s[1] = dobj; s[2] = sqrt(ad[1]); s[3] = sqrt(ap[2]);
    s[4] = sqrt(ad[2]); s[5] = sqrt(ap[3]); s[6] = sqrt(ap[4]); s[7] = sqrt(ad[3]);
    const double gap = fabs(pobj - dobj);
    double e = s[2] / (1.0 + s[4]);
    e = fmax(e, s[3] / (1.0 + s[5]));
    e = fmax(e, gap / (1.0 + fabs(pobj) + fabs(dobj)));
    s[8] = e; s[9] = gap;
    dd[0] = ap[5]; dd[1] = ad[4];
}

// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1)
k_batch_run(const BatchInst* __restrict__ insts, int count, int shared, double* x, double* y, const double* b,
            const double* c, const double* tau, const double* sigma, int iters, double* scalars)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    for (int inst = blockIdx.x; inst < count; inst += gridDim.x) {
        BatchInst I = insts[shared ? 0 : inst];
        if (shared) { I.x_off = (long long)inst * I.n; I.y_off = (long long)inst * I.m; }
        const BatchSmem S = carve(dsm, I.m, I.n, true);
        load_instance(I, S, x, y, b, c);
        for (int k = threadIdx.x; k < I.n; k += blockDim.x) S.x0[k] = 0.0;   // anchors unused here; keep eval finite
        for (int k = threadIdx.x; k < I.m; k += blockDim.x) S.y0[k] = 0.0;
        const DevLP lp = smem_lp(I, S);
        const MatView VA = global_view(I.A, 0u), VAT = global_view(I.AT, 0u);
        PrimalOp<false, SmemMem> pop{lp, __ldg(tau + inst)};
        DualOp<false, SmemMem> dop{lp, __ldg(sigma + inst)};
        double acc[NRED];
        __syncthreads();
        for (int it = 0; it < iters; ++it) {
            run_phase(lp.AT, VAT, pop, acc);
            __syncthreads();
            run_phase(lp.A, VA, dop, acc);
            __syncthreads();
        }
        if (scalars) {
            double s[10], dd[2];
            batch_kkt(lp, VA, VAT, S, s, dd);
            if (threadIdx.x == 0) {
                double* o = scalars + (size_t)inst * MLLP_NUM_SCALARS;
                for (int k = 0; k < 10; ++k) o[k] = s[k];
                o[10] = (double)iters; o[11] = 0.0; o[12] = 0.0; o[13] = 1.0; o[14] = 0.0; o[15] = 0.0;
            }
        }
        store_instance(I, S, x, y);
    }
}

// Solve mode per CTA: same control flow as k_solve_persistent / oracle_pdhg_solve, with
// __syncthreads() in place of the grid barrier.
__global__ void __launch_bounds__(1024, 1)
k_batch_solve(const BatchInst* __restrict__ insts, int count, int shared, double* x, double* y, const double* b,
              const double* c, const double* eta_arr, double w0, int max_iters, int check_every, double tol,
              double* scalars, int* next_inst)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    __shared__ int s_inst;
    // instances converge after very different iteration counts: CTAs pull the next instance from a
    // device counter instead of a static stride
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_inst = atomicAdd(next_inst, 1);
        __syncthreads();
        const int inst = s_inst;
        if (inst >= count) break;
        BatchInst I = insts[shared ? 0 : inst];
        if (shared) { I.x_off = (long long)inst * I.n; I.y_off = (long long)inst * I.m; }
        const BatchSmem S = carve(dsm, I.m, I.n, true);
        load_instance(I, S, x, y, b, c);
        for (int k = threadIdx.x; k < I.n; k += blockDim.x) S.x0[k] = S.x[k];
        for (int k = threadIdx.x; k < I.m; k += blockDim.x) S.y0[k] = S.y[k];
        __syncthreads();
        const DevLP lp = smem_lp(I, S);
        const MatView VA = global_view(I.A, 0u), VAT = global_view(I.AT, 0u);
        const double eta = __ldg(eta_arr + inst);
        double w = w0, fpe_restart = -1.0, fpe_prev = INFINITY, fpe = 0.0;
        int k = 0, it = 0, restarts = 0, converged = 0;
        double kk[10], dd[2];
        batch_kkt(lp, VA, VAT, S, kk, dd);
        while (it < max_iters) {
            const double tau = eta / w, sigma = eta * w;
            const double lam = (double)(k + 1) / (double)(k + 2);
            const bool check = ((it + 1) % check_every == 0) || (it + 1 == max_iters);
            const bool need_fpe = check || fpe_restart < 0.0;
            double a2[2];
            {
                PrimalHalpernOp<false, SmemMem> op{lp, tau, lam};
                double acc[NRED];
                acc[0] = 0.0;
                run_phase(lp.AT, VAT, op, acc);
                a2[0] = acc[0];
            }
            __syncthreads();
            {
                DualHalpernOp<false, SmemMem> op{lp, sigma, lam};
                double acc[NRED];
                acc[0] = 0.0;
                run_phase(lp.A, VA, op, acc);
                a2[1] = acc[0];
            }
            __syncthreads();
            ++it; ++k;
            if (need_fpe) {
                cta_allreduce<2>(a2, S);
                fpe = sqrt(w * a2[0] + a2[1] / w);
                if (fpe_restart < 0.0) fpe_restart = fpe;
            }
            if (check) {
                batch_kkt(lp, VA, VAT, S, kk, dd);
                if (kk[8] <= tol) { converged = 1; break; }
                const bool do_restart = (fpe <= 0.2 * fpe_restart) || (fpe <= 0.8 * fpe_restart && fpe > fpe_prev) ||
                                        ((double)k >= 0.36 * (double)it);
                fpe_prev = fpe;
                if (do_restart) {
                    const double ddx = sqrt(dd[0]), ddy = sqrt(dd[1]);
                    if (ddx > 1e-10 && ddy > 1e-10) w = exp(0.5 * log(ddy / ddx) + 0.5 * log(w));
                    for (int q = threadIdx.x; q < I.n; q += blockDim.x) S.x0[q] = S.x[q];
                    for (int q = threadIdx.x; q < I.m; q += blockDim.x) S.y0[q] = S.y[q];
                    __syncthreads();
                    k = 0; fpe_restart = -1.0; fpe_prev = INFINITY;
                    ++restarts;
                }
            }
        }
        if (threadIdx.x == 0) {
            double* o = scalars + (size_t)inst * MLLP_NUM_SCALARS;
            for (int q = 0; q < 10; ++q) o[q] = kk[q];
            o[10] = (double)it; o[11] = (double)restarts; o[12] = (double)converged; o[13] = w; o[14] = fpe; o[15] = 0.0;
        }
        store_instance(I, S, x, y);
    }
}

// sigma_max(A) per instance by power iteration (same recurrence as oracle_power_iteration).
__global__ void __launch_bounds__(1024, 1)
k_batch_norm(const BatchInst* __restrict__ insts, int count, int shared, int iters, double* sigma_out)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    for (int inst = blockIdx.x; inst < count; inst += gridDim.x) {
        const BatchInst I = insts[shared ? 0 : inst];
        const BatchSmem S = carve(dsm, I.m, I.n, true);
        // v = S.x, w = S.y, z = S.xbar
        for (int k = threadIdx.x; k < I.n; k += blockDim.x) S.x[k] = rsqrt((double)I.n);
        __syncthreads();
        const MatView VA = global_view(I.A, 0u), VAT = global_view(I.AT, 0u);
        double lam = 0.0;
        double acc[NRED];
        for (int it = 0; it < iters; ++it) {
            { SpmvOp<SmemMem> op{S.x, S.y}; run_phase(I.A, VA, op, acc); }
            __syncthreads();
            { SpmvOp<SmemMem> op{S.y, S.xbar}; run_phase(I.AT, VAT, op, acc); }
            __syncthreads();
            double nz2[1] = {0.0};
            for (int k = threadIdx.x; k < I.n; k += blockDim.x) nz2[0] += S.xbar[k] * S.xbar[k];
            cta_allreduce<1>(nz2, S);
            const double nz = sqrt(nz2[0]);
            lam = nz;
            if (nz == 0.0) break;
            for (int k = threadIdx.x; k < I.n; k += blockDim.x) S.x[k] = S.xbar[k] / nz;
            __syncthreads();
        }
        if (threadIdx.x == 0) sigma_out[inst] = sqrt(lam);
        __syncthreads();
        if (shared) {   // one matrix: every instance gets the same value
            if (threadIdx.x == 0) for (int q = 1; q < count; ++q) sigma_out[q] = sqrt(lam);
            break;
        }
    }
}

}  // namespace mllp

// ---------------------------------------------------------------------------------------
// host side
namespace mllp {
void set_last_error(const std::string& msg);  // cabi.cu: the message mllp_last_error() returns
}
namespace {
int bfail(int code, const std::string& msg) { mllp::set_last_error(msg); return code; }
}  // namespace

struct mllp_batch {
    int device = 0;
    int count = 0, shared = 0;
    int threads = 0, grid = 0;
    size_t dyn_smem = 0;
    int64_t sum_m = 0, sum_n = 0, sum_nnz = 0;
    int max_m = 0, max_n = 0;
    BatchInst* d_insts = nullptr;
    std::vector<void*> allocs;
    int* d_next = nullptr;            // work counter of k_batch_solve
    int64_t info[16] = {0};
};

namespace {

struct DevGuard {
    int prev = -1;
    explicit DevGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

template <class T>
cudaError_t up(mllp_batch* bt, T** out, const std::vector<T>& h)
{
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(1, h.size()) * sizeof(T));
    if (e != cudaSuccess) return e;
    bt->allocs.push_back(p);
    if (!h.empty()) e = cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    *out = (T*)p;
    return e;
}

// All instances' images are concatenated into a few big arrays; DevMat pointers are offsets into them.
struct Pools {
    std::vector<double> vals;
    std::vector<int32_t> idx;
    std::vector<Tile> tiles;
    std::vector<uint32_t> u32;       // cta_begin, cta_step_begin, cta_lsplit_begin, cta_nsplit
    std::vector<SplitRow> splits;
    std::vector<LocalSplit> lsplits;
    std::vector<int32_t> order;
};
struct MatOff { size_t vals, idx, tiles, cb, csb, clb, cns, splits, lsplits; int nrows, ncols; };

MatOff append(Pools& P, const HostMat& H)
{
    MatOff o;
    o.nrows = H.nrows; o.ncols = H.ncols;
    o.vals = P.vals.size(); P.vals.insert(P.vals.end(), H.vals.begin(), H.vals.end());
    o.idx = P.idx.size(); P.idx.insert(P.idx.end(), H.idx.begin(), H.idx.end());
    // keep 16 B alignment of the tile pool entries (16 B each) -- vals/idx are multiples of 64 entries
    o.tiles = P.tiles.size(); P.tiles.insert(P.tiles.end(), H.tiles.begin(), H.tiles.end());
    o.cb = P.u32.size(); P.u32.insert(P.u32.end(), H.cta_begin.begin(), H.cta_begin.end());
    o.csb = P.u32.size(); P.u32.insert(P.u32.end(), H.cta_step_begin.begin(), H.cta_step_begin.end());
    o.clb = P.u32.size(); P.u32.insert(P.u32.end(), H.cta_lsplit_begin.begin(), H.cta_lsplit_begin.end());
    o.cns = P.u32.size(); P.u32.insert(P.u32.end(), H.cta_nsplit.begin(), H.cta_nsplit.end());
    o.splits = P.splits.size(); P.splits.insert(P.splits.end(), H.splits.begin(), H.splits.end());
    o.lsplits = P.lsplits.size(); P.lsplits.insert(P.lsplits.end(), H.lsplits.begin(), H.lsplits.end());
    return o;
}

}  // namespace

extern "C" {

int mllp_batch_create(int32_t count, int32_t shared_matrix, const int32_t* h_m, const int32_t* h_n,
                      const int64_t* h_indptr_off, const int64_t* h_nnz_off, const int32_t* h_indptr,
                      const int32_t* h_indices, const double* h_values, int device, uint32_t flags, mllp_batch_t* out)
{
    (void)flags;
    if (!out) return bfail(MLLP_E_INVALID, "mllp_batch_create: null output handle");
    *out = nullptr;
    if (count < 1 || !h_m || !h_n || !h_indptr || !h_indptr_off || !h_nnz_off)
        return bfail(MLLP_E_INVALID, "mllp_batch_create: bad count or null arrays");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess) return bfail((int)e, std::string("mllp_batch_create: cudaGetDeviceCount: ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return bfail(MLLP_E_INVALID, "mllp_batch_create: no such CUDA device");
    DevGuard guard(device);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return bfail((int)e, "mllp_batch_create: cudaGetDeviceProperties failed");
    if (prop.major != 10) return bfail(MLLP_E_STATE, "mllp_batch_create: this library is built for sm_100a (B200) only");

    mllp_batch* bt = new (std::nothrow) mllp_batch();
    if (!bt) return bfail(MLLP_E_NOMEM, "mllp_batch_create: out of host memory");
    bt->device = device; bt->count = count; bt->shared = shared_matrix ? 1 : 0;
    const int nmat = bt->shared ? 1 : count;

    BuildParams bp;
    bp.num_ctas = 1;           // one CTA walks the whole instance
    bp.pref_steps = 2;         // small LPs: favour lanes over steps (latency, not throughput)
    bp.max_steps = 4;
    const char* ev = getenv("MLLP_BATCH_PREF_STEPS");
    if (ev && *ev) bp.pref_steps = std::max(1, atoi(ev));
    if (bp.max_steps < bp.pref_steps) bp.max_steps = bp.pref_steps;

    int rc = 0;
    try {
        Pools P;
        std::vector<MatOff> offA((size_t)nmat), offAT((size_t)nmat);
        std::vector<size_t> offOX((size_t)nmat), offOY((size_t)nmat);
        int max_tiles = 0;
        for (int k = 0; k < nmat && rc == 0; ++k) {
            const int m = h_m[k], n = h_n[k];
            const int32_t* ip = h_indptr + h_indptr_off[k];
            const int32_t* ii = h_indices ? h_indices + h_nnz_off[k] : nullptr;
            const double* vv = h_values ? h_values + h_nnz_off[k] : nullptr;
            if (m < 0 || n < 0 || ip[0] != 0) { rc = bfail(MLLP_E_INVALID, "mllp_batch_create: bad instance shape / indptr"); break; }
            const int64_t nnz = ip[m];
            if (nnz > 0 && (!ii || !vv)) { rc = bfail(MLLP_E_INVALID, "mllp_batch_create: null indices/values"); break; }
            for (int64_t q = 0; q < nnz; ++q)
                if (ii[q] < 0 || ii[q] >= n) { rc = bfail(MLLP_E_INVALID, "mllp_batch_create: column index out of range"); break; }
            if (rc) break;
            std::vector<int32_t> tptr, tind;
            std::vector<double> tval;
            csr_transpose(m, n, ip, ii, vv, tptr, tind, tval);
            std::vector<int32_t> orderY, posY, orderX, posX;
            plan_orders(m, n, ip, ii, tptr.data(), tind.data(), bp, orderY, posY, orderX, posX);
            HostMat HA, HAT;
            build_host_mat(m, n, ip, ii, vv, orderY, posX, bp, HA);
            build_host_mat(n, m, tptr.data(), tind.data(), tval.data(), orderX, posY, bp, HAT);
            offA[k] = append(P, HA);
            offAT[k] = append(P, HAT);
            offOX[k] = P.order.size(); P.order.insert(P.order.end(), orderX.begin(), orderX.end());
            offOY[k] = P.order.size(); P.order.insert(P.order.end(), orderY.begin(), orderY.end());
            max_tiles = std::max<int>(max_tiles, (int)std::max(HA.tiles.size(), HAT.tiles.size()));
            bt->max_m = std::max(bt->max_m, m); bt->max_n = std::max(bt->max_n, n);
            bt->sum_nnz += nnz;
        }
        if (rc == 0) {
            for (int k = 0; k < count; ++k) {
                bt->sum_m += h_m[bt->shared ? 0 : k];
                bt->sum_n += h_n[bt->shared ? 0 : k];
            }
            double* d_vals; int32_t* d_idx; Tile* d_tiles; uint32_t* d_u32; SplitRow* d_splits; LocalSplit* d_ls; int32_t* d_order;
            double* d_dummy_partials; unsigned* d_dummy_counters;
            auto ck = [&](cudaError_t ce, const char* what) { if (ce != cudaSuccess && rc == 0) rc = bfail((int)ce, std::string(what) + ": " + cudaGetErrorString(ce)); };
            ck(up(bt, &d_vals, P.vals), "upload vals");
            ck(up(bt, &d_idx, P.idx), "upload idx");
            ck(up(bt, &d_tiles, P.tiles), "upload tiles");
            ck(up(bt, &d_u32, P.u32), "upload tables");
            ck(up(bt, &d_splits, P.splits), "upload splits");
            ck(up(bt, &d_ls, P.lsplits), "upload local splits");
            ck(up(bt, &d_order, P.order), "upload orders");
            ck(up(bt, &d_dummy_partials, std::vector<double>(P.splits.size() + 1, 0.0)), "alloc partials");
            ck(up(bt, &d_dummy_counters, std::vector<unsigned>(P.splits.size() + 1, 0u)), "alloc counters");
            ck(up(bt, &bt->d_next, std::vector<int>(4, 0)), "alloc work counter");
            if (rc == 0) {
                auto dev_mat = [&](const MatOff& o) {
                    DevMat D{};
                    D.vals = reinterpret_cast<const double2*>(d_vals + o.vals);
                    D.idx = reinterpret_cast<const int2*>(d_idx + o.idx);
                    D.tiles = d_tiles + o.tiles;
                    D.cta_begin = d_u32 + o.cb; D.cta_step_begin = d_u32 + o.csb;
                    D.cta_lsplit_begin = d_u32 + o.clb; D.cta_nsplit = d_u32 + o.cns;
                    D.splits = d_splits + o.splits; D.lsplits = d_ls + o.lsplits;
                    D.partials = d_dummy_partials + o.splits; D.slots = nullptr; D.counters = d_dummy_counters + o.splits;
                    D.nrows = o.nrows; D.ncols = o.ncols;
                    return D;
                };
                std::vector<BatchInst> insts((size_t)nmat);
                long long xo = 0, yo = 0;
                for (int k = 0; k < nmat; ++k) {
                    BatchInst& I = insts[k];
                    I.A = dev_mat(offA[k]); I.AT = dev_mat(offAT[k]);
                    I.orderX = d_order + offOX[k]; I.orderY = d_order + offOY[k];
                    I.m = h_m[k]; I.n = h_n[k];
                    I.x_off = xo; I.y_off = yo;
                    xo += h_n[k]; yo += h_m[k];
                }
                ck(up(bt, &bt->d_insts, insts), "upload instance table");
            }
            // launch geometry: warps ~ tiles per phase (latency bound), shared memory for the vectors
            const size_t vec_bytes = 8 * ((size_t)32 * NRED + 16 + 4 * (size_t)bt->max_n + 3 * (size_t)bt->max_m) + 32;
            bt->dyn_smem = (vec_bytes + 15) & ~(size_t)15;
            const size_t smem_cap = (size_t)prop.sharedMemPerBlockOptin - 4096;
            if (rc == 0 && bt->dyn_smem > smem_cap)
                rc = bfail(MLLP_E_STATE, "mllp_batch_create: an instance does not fit in shared memory (8(4n+3m) bytes needed); use mllp_lp_create for it");
            int threads = 128;
            while (threads < 1024 && threads / 32 < max_tiles) threads *= 2;
            const char* tv = getenv("MLLP_BATCH_THREADS");
            if (tv && *tv) threads = std::max(32, std::min(1024, atoi(tv) & ~31));
            bt->threads = threads;
            if (rc == 0) {
                const void* fns[3] = {(const void*)k_batch_run, (const void*)k_batch_solve, (const void*)k_batch_norm};
                int nb = 1 << 30;
                for (const void* fn : fns) {
                    if (bt->dyn_smem > 40 * 1024)
                        ck(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bt->dyn_smem), "cudaFuncSetAttribute");
                    int b = 0;
                    ck(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, fn, threads, bt->dyn_smem), "occupancy");
                    nb = std::min(nb, b);
                }
                if (rc == 0 && nb < 1) rc = bfail(MLLP_E_STATE, "mllp_batch_create: kernel does not fit on an SM");
                if (rc == 0) bt->grid = std::min<int64_t>(count, (int64_t)prop.multiProcessorCount * nb);
            }
            int64_t* I = bt->info;
            I[0] = count; I[1] = bt->sum_m; I[2] = bt->sum_n; I[3] = bt->sum_nnz; I[4] = bt->grid; I[5] = bt->threads;
            I[6] = (int64_t)bt->dyn_smem;
            // algorithmic bytes per batch iteration: per instance 36 m + 44 n, plus the matrix (24 nnz) once per
            // instance (separate matrices) or once per batch (shared matrix)
            I[7] = 36 * bt->sum_m + 44 * bt->sum_n + 24 * bt->sum_nnz;
            I[8] = 1;
        }
    } catch (const std::bad_alloc&) {
        rc = bfail(MLLP_E_NOMEM, "mllp_batch_create: out of host memory");
    }
    if (rc != 0) {
        mllp_batch_destroy(bt);
        return rc;
    }
    *out = bt;
    return 0;
}

int mllp_batch_destroy(mllp_batch_t bt)
{
    if (!bt) return 0;
    DevGuard guard(bt->device);
    for (void* p : bt->allocs) cudaFree(p);
    delete bt;
    return 0;
}

int mllp_batch_info(mllp_batch_t bt, int64_t* out16)
{
    if (!bt || !out16) return bfail(MLLP_E_INVALID, "mllp_batch_info: null argument");
    memcpy(out16, bt->info, sizeof(bt->info));
    return 0;
}

int mllp_batch_estimate_norm(mllp_batch_t bt, int32_t iters, double* d_sigma_max, void* stream)
{
    if (!bt || !d_sigma_max || iters < 1) return bfail(MLLP_E_INVALID, "mllp_batch_estimate_norm: bad argument");
    DevGuard guard(bt->device);
    const int grid = bt->shared ? 1 : bt->grid;
    k_batch_norm<<<grid, bt->threads, bt->dyn_smem, (cudaStream_t)stream>>>(bt->d_insts, bt->count, bt->shared, iters, d_sigma_max);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return bfail((int)e, std::string("mllp_batch_estimate_norm: ") + cudaGetErrorString(e));
    return 0;
}

int mllp_batch_run(mllp_batch_t bt, double* d_x, double* d_y, const double* d_b, const double* d_c, const double* d_tau,
                   const double* d_sigma, int32_t num_iters, double* d_scalars, void* stream)
{
    if (!bt || !d_x || !d_y || !d_b || !d_c || !d_tau || !d_sigma || num_iters < 0)
        return bfail(MLLP_E_INVALID, "mllp_batch_run: null argument or negative iteration count");
    DevGuard guard(bt->device);
    k_batch_run<<<bt->grid, bt->threads, bt->dyn_smem, (cudaStream_t)stream>>>(bt->d_insts, bt->count, bt->shared, d_x, d_y, d_b,
                                                                              d_c, d_tau, d_sigma, num_iters, d_scalars);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return bfail((int)e, std::string("mllp_batch_run: ") + cudaGetErrorString(e));
    return 0;
}

int mllp_batch_solve(mllp_batch_t bt, double* d_x, double* d_y, const double* d_b, const double* d_c, const double* d_eta,
                     double w0, int32_t max_iters, int32_t check_every, double tol, double* d_scalars, void* stream)
{
    if (!bt || !d_x || !d_y || !d_b || !d_c || !d_eta || !d_scalars || max_iters < 0 || check_every < 1 || !(w0 > 0.0))
        return bfail(MLLP_E_INVALID, "mllp_batch_solve: bad argument");
    DevGuard guard(bt->device);
    cudaError_t e0 = cudaMemsetAsync(bt->d_next, 0, sizeof(int), (cudaStream_t)stream);
    if (e0 != cudaSuccess) return bfail((int)e0, std::string("mllp_batch_solve: ") + cudaGetErrorString(e0));
    k_batch_solve<<<bt->grid, bt->threads, bt->dyn_smem, (cudaStream_t)stream>>>(bt->d_insts, bt->count, bt->shared, d_x, d_y, d_b,
                                                                                d_c, d_eta, w0, max_iters, check_every, tol,
                                                                                d_scalars, bt->d_next);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return bfail((int)e, std::string("mllp_batch_solve: ") + cudaGetErrorString(e));
    return 0;
}

}  // extern "C"
