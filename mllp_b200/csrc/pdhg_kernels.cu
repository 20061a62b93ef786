// pdhg_kernels.cu -- kernels and launchers of the single-instance PDHG path (sm_100a).
#include "pdhg_host.h"
#include "pdhg_kernels.cuh"

namespace mllp {

// ---------------------------------------------------------------------------------------
// small utility kernels (boundary reordering, not on the per-iteration path)
__global__ void k_gather(double* __restrict__ dst, const double* __restrict__ src,
                         const int32_t* __restrict__ order, int n)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
        dst[k] = order[k] >= 0 ? src[order[k]] : 0.0;   // -1: padding of a row-partitioned slice
}
__global__ void k_scatter(double* __restrict__ dst, const double* __restrict__ src,
                          const int32_t* __restrict__ order, int n)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
        if (order[k] >= 0) dst[order[k]] = src[k];
}
// the same with a diagonal scaling in internal order (preconditioned handles: the caller's vectors are the ORIGINAL LP's)
__global__ void k_gather_scaled(double* __restrict__ dst, const double* __restrict__ src, const int32_t* __restrict__ order,
                                const double* __restrict__ s, int div, int n)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
        dst[k] = order[k] >= 0 ? (div ? src[order[k]] / s[k] : src[order[k]] * s[k]) : 0.0;
}
__global__ void k_scatter_scaled(double* __restrict__ dst, const double* __restrict__ src, const int32_t* __restrict__ order,
                                 const double* __restrict__ s, int div, int n)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
        if (order[k] >= 0) dst[order[k]] = div ? src[k] / s[k] : src[k] * s[k];
}
__global__ void k_fill(double* dst, double v, int n)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) dst[k] = v;
}
// out[0] = sum v^2 in a fixed order: 148 block partials (grid-stride, block tree), then one block.
constexpr int SUMSQ_BLOCKS = 148;
__device__ __forceinline__ double block_sum(double s, double* sm)
{
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    s = 0.0;
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    }
    return s;   // valid in thread 0
}
__global__ void k_sumsq_partial(const double* __restrict__ v, int n, double* __restrict__ part)
{
    __shared__ double sm[32];
    double s = 0.0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) s += v[k] * v[k];
    s = block_sum(s, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
}
__global__ void k_sumsq_final(const double* __restrict__ part, int nparts, double* out)
{
    __shared__ double sm[32];
    double s = threadIdx.x < nparts ? part[threadIdx.x] : 0.0;
    s = block_sum(s, sm);
    if (threadIdx.x == 0) out[0] = s;
}
// dst = src / sqrt(sqrt-free norm2[0])   (dst = src * rsqrt(norm2))
__global__ void k_scale_by_invnorm(double* __restrict__ dst, const double* __restrict__ src,
                                   const double* norm2, int n)
{
    const double nz = sqrt(norm2[0]);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
        dst[k] = nz > 0.0 ? src[k] / nz : src[k];
}

// CSR -> COO edge list of the LP's bipartite graph: one warp per row, lanes stride the row.
// edge_index is [2][nnz] int64 (row 0: variable = column id, row 1: constraint = row id), edge_attr
// float32, in CSR nonzero order (the order the reference's Python loop produces).
__global__ void k_graph_edges(int m, long long nnz, const int32_t* __restrict__ indptr,
                              const int32_t* __restrict__ indices, const double* __restrict__ values,
                              long long* __restrict__ edge_index, float* __restrict__ edge_attr)
{
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < m; row += warps) {
        const int32_t a = indptr[row], b = indptr[row + 1];
        for (int32_t k = a + lane; k < b; k += 32) {
            edge_index[k] = (long long)indices[k];
            edge_index[nnz + k] = (long long)row;
            edge_attr[k] = (float)values[k];
        }
    }
}

int launch_graph_edges(int m, long long nnz, const int32_t* indptr, const int32_t* indices, const double* values,
                       long long* edge_index, float* edge_attr, cudaStream_t s)
{
    if (m <= 0 || nnz <= 0) return 0;
    const int blocks = (m + 7) / 8 > 2368 ? 2368 : (m + 7) / 8;
    count_launch(1);
    k_graph_edges<<<blocks, 256, 0, s>>>(m, nnz, indptr, indices, values, edge_index, edge_attr);
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// phase kernels (graph mode, evaluation, unit SpMV).  Grid = the persistent grid G: the tile
// to CTA assignment is fixed at build time.
__global__ void __launch_bounds__(1024, 1) k_spmv(DevMat M, const double* in, double* out)
{
    SpmvOp<> op{in, out};
    double acc[NRED];
    run_phase(M, global_view(M), op, acc);
}

template <bool BOUNDS>
__global__ void __launch_bounds__(1024, 1) k_primal(DevLP lp)
{
    PrimalOp<BOUNDS> op{lp, __ldcg(lp.ctrl + CTRL_TAU)};
    const MatView VAT = global_view(lp.AT);
    double acc[NRED];
    run_phase(lp.AT, VAT, op, acc);
}
template <bool BOUNDS>
__global__ void __launch_bounds__(1024, 1) k_dual(DevLP lp)
{
    DualOp<BOUNDS> op{lp, __ldcg(lp.ctrl + CTRL_SIGMA)};
    const MatView VA = global_view(lp.A);
    double acc[NRED];
    run_phase(lp.A, VA, op, acc);
}

template <bool BOUNDS>
__device__ __forceinline__ void eval_phases(const DevLP& lp, const MatView& VA, const MatView& VAT, double* red_p,
                                            double* red_d, double* smem)
{
    {
        EvalPrimalOp<BOUNDS> op{lp};
        double acc[NRED];
#pragma unroll
        for (int k = 0; k < NRED; ++k) acc[k] = 0.0;
        run_phase(lp.AT, VAT, op, acc);
        cta_reduce_store<7>(acc, red_p + (size_t)blockIdx.x * NRED, smem);
    }
    {
        EvalDualOp<BOUNDS> op{lp};
        double acc[NRED];
#pragma unroll
        for (int k = 0; k < NRED; ++k) acc[k] = 0.0;
        run_phase(lp.A, VA, op, acc);
        cta_reduce_store<6>(acc, red_d + (size_t)blockIdx.x * NRED, smem);
    }
}

template <bool BOUNDS>
__global__ void __launch_bounds__(1024, 1) k_eval(DevLP lp, int G)
{
    __shared__ double smem[32 * NRED];
    eval_phases<BOUNDS>(lp, global_view(lp.A), global_view(lp.AT), lp.red + (size_t)RED_EVALP * G * NRED,
                        lp.red + (size_t)RED_EVALD * G * NRED, smem);
}

// Turn the summed eval partials into the public scalars (calling warp, all lanes).
__device__ __forceinline__ void kkt_from_sums(const double* red_p, const double* red_d, int G, double* s /*[10]*/)
{
    const double pobj = grid_sum(red_p, G, 0);
    const double dbnd = grid_sum(red_p, G, 1);
    const double dr2 = grid_sum(red_p, G, 2) + grid_sum(red_d, G, 5);   // + distance of y from its cone
    const double nc2 = grid_sum(red_p, G, 3);
    const double nx2 = grid_sum(red_p, G, 4);
    const double by = grid_sum(red_d, G, 0);
    const double pr2 = grid_sum(red_d, G, 1) + grid_sum(red_p, G, 6);   // + distance of x from its box
    const double nb2 = grid_sum(red_d, G, 2);
    const double ny2 = grid_sum(red_d, G, 3);
    const double dobj = by + dbnd;
    s[0] = pobj; s[1] = dobj; s[2] = sqrt(pr2); s[3] = sqrt(dr2);
    s[4] = sqrt(nb2); s[5] = sqrt(nc2); s[6] = sqrt(nx2); s[7] = sqrt(ny2);
    const double gap = fabs(pobj - dobj);
    double e = s[2] / (1.0 + s[4]);
    e = fmax(e, s[3] / (1.0 + s[5]));
    e = fmax(e, gap / (1.0 + fabs(pobj) + fabs(dobj)));
    s[8] = e; s[9] = gap;
}

__global__ void k_eval_finalize(DevLP lp, int G, double* out, double iters)
{
    double s[10];
    kkt_from_sums(lp.red + (size_t)RED_EVALP * G * NRED, lp.red + (size_t)RED_EVALD * G * NRED, G, s);
    if (threadIdx.x == 0) {
        for (int k = 0; k < 10; ++k) out[k] = s[k];
        out[10] = iters; out[11] = 0.0; out[12] = 0.0; out[13] = 1.0; out[14] = 0.0; out[15] = 0.0;
    }
}

// ---------------------------------------------------------------------------------------
// persistent cooperative kernel, parity mode: `iters` full iterations in one launch,
// two grid barriers per iteration, no host involvement.
// Shared-memory layout of the persistent kernels (dynamic part):
//   [ A: tile descriptors, resident vals, resident idx | A': same ]
struct PersistentSmem {
    MatView VA, VAT;
    unsigned long long tag;   // tag of the last polled split-row join (one per phase call)
    double *ys, *bs, *xs, *cs; // own entries of the CTA's regular rows (parity kernel), or null
    uint32_t used;             // bytes of the dynamic shared memory taken by the two matrix views
};

// CTA-local row slots of the regular tiles: written into the .split field of the shared-memory copy of
// the descriptors (it is -1 there for regular tiles); one thread, once per launch.
__device__ __forceinline__ void assign_row_slots(const MatView& V)
{
    Tile* d = const_cast<Tile*>(V.desc);
    int run = 0;
    for (uint32_t t = V.nsplit; t < V.ntiles; ++t) {
        d[t].split = run;
        run += d[t].nrows;
    }
}
// own[slot] <-> vec[row] for all regular rows of this CTA
template <bool TO_SMEM>
__device__ __forceinline__ void copy_own(const MatView& V, double* vec, double* own)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (uint32_t t = V.nsplit + warp; t < V.ntiles; t += nwarps) {
        const int4 raw = *reinterpret_cast<const int4*>(V.desc + t);
        const int nrows = (raw.z >> 24) & 0xff;
        if (lane < nrows) {
            if (TO_SMEM) own[raw.w + lane] = __ldcg(vec + raw.y + lane);
            else vec[raw.y + lane] = own[raw.w + lane];
        }
    }
}

__device__ __forceinline__ void persistent_setup(const DevLP& lp, unsigned char* dsm, PersistentSmem& P, bool own = false)
{
    unsigned char* base = dsm;
    P.tag = lp.join_base;
    const uint32_t used = resident_view(lp.A, lp.res_steps_A, base, P.VA);
    const uint32_t used2 = resident_view(lp.AT, lp.res_steps_AT, base + used, P.VAT);
    P.ys = P.bs = P.xs = P.cs = nullptr;
    P.used = used + used2;
    __syncthreads();
    if (own && lp.own_rows_A + lp.own_rows_AT > 0) {
        P.ys = reinterpret_cast<double*>(base + used + used2);
        P.bs = P.ys + lp.own_rows_A;
        P.xs = P.bs + lp.own_rows_A;
        P.cs = P.xs + lp.own_rows_AT;
        if (threadIdx.x == 0) assign_row_slots(P.VA);
        if (threadIdx.x == 32) assign_row_slots(P.VAT);
        __syncthreads();
        copy_own<true>(P.VA, lp.y, P.ys);
        copy_own<true>(P.VA, const_cast<double*>(lp.b), P.bs);
        copy_own<true>(P.VAT, lp.x, P.xs);
        copy_own<true>(P.VAT, const_cast<double*>(lp.c), P.cs);
        __syncthreads();
    }
}

// A' phase (gathers y) and A phase (gathers xbar).
template <class Op>
__device__ __forceinline__ void phase_AT(const DevLP& lp, PersistentSmem& P, const Op& op, double* acc)
{
    run_phase<true>(lp.AT, P.VAT, op, acc, ++P.tag);
}
template <class Op>
__device__ __forceinline__ void phase_A(const DevLP& lp, PersistentSmem& P, const Op& op, double* acc)
{
    run_phase<true>(lp.A, P.VA, op, acc, ++P.tag);
}

template <bool BOUNDS>
__global__ void __launch_bounds__(1024, 1) k_pdhg_persistent(DevLP lp, double tau, double sigma, int iters)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    PersistentSmem P;
    persistent_setup(lp, dsm, P, true);
    unsigned target = 0;
    double acc[NRED];
    if (P.xs) {
        // own entries (x, c / y, b) of the CTA's rows stay in shared memory across the iterations
        PrimalResOp<BOUNDS> pop{lp, tau, P.xs, P.cs};
        DualResOp<BOUNDS> dop{lp, sigma, P.ys, P.bs};
        for (int it = 0; it < iters; ++it) {
            unsigned long long* tr = lp.trace ? lp.trace + ((size_t)it * gridDim.x + blockIdx.x) * 4 : nullptr;
            phase_AT(lp, P, pop, acc);
            sync_all(lp, target, tr);
            phase_A(lp, P, dop, acc);
            sync_all(lp, target, tr ? tr + 2 : nullptr);
        }
        copy_own<false>(P.VAT, lp.x, P.xs);   // y was published every iteration
        return;
    }
    PrimalOp<BOUNDS> pop{lp, tau};
    DualOp<BOUNDS> dop{lp, sigma};
    for (int it = 0; it < iters; ++it) {
        unsigned long long* tr = lp.trace ? lp.trace + ((size_t)it * gridDim.x + blockIdx.x) * 4 : nullptr;
        phase_AT(lp, P, pop, acc);
        sync_all(lp, target, tr);
        phase_A(lp, P, dop, acc);
        sync_all(lp, target, tr ? tr + 2 : nullptr);
    }
}

// Row-partitioned persistent kernel (see PeerInfo): replicated A' phase, local grid barrier, A phase on this rank's rows
// with the new duals also mailed to the peers, unpack of the peers' duals, local grid barrier.  One cross-GPU exchange
// per iteration, no NCCL call and no kernel launch inside the loop.  `seq` = exchanges done on this handle so far (the
// same on every rank: the calls are collective); it selects the mailbox buffer and makes the tags unique.
template <bool BOUNDS>
__global__ void __launch_bounds__(1024, 1)
k_pdhg_rowpart(DevLP lp, PeerInfo pi, double tau, double sigma, int iters, unsigned long long seq)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    PersistentSmem P;
    persistent_setup(lp, dsm, P);
    unsigned target = 0;
    PrimalOp<BOUNDS> pop{lp, tau};
    double acc[NRED];
    for (int it = 0; it < iters; ++it) {
        // dev trace, 6 slots per (iteration, CTA): A' phase done, barrier released, A phase done, unpack done, barrier released
        unsigned long long* tr = lp.trace ? lp.trace + ((size_t)it * gridDim.x + blockIdx.x) * 6 : nullptr;
        const unsigned long long s = seq + (unsigned long long)it + 1ull;
        const size_t buf = (size_t)(s & 1ull) * 2 * (size_t)pi.mi;
        phase_AT(lp, P, pop, acc);
        grid_barrier(lp.barrier, target, tr);
        DualMailOp<BOUNDS> dop{lp, pi, sigma, s, buf};
        phase_A(lp, P, dop, acc);
        if (tr) {
            __syncthreads();
            if (threadIdx.x == 0) tr[2] = global_ns();
        }
        unpack_mail(lp, pi, s, buf);
        grid_barrier(lp.barrier, target, tr ? tr + 3 : nullptr);
    }
}

// persistent cooperative kernel, solve mode (reflected restarted Halpern PDHG).
// Control state is replicated: every CTA derives it from the same global partial sums with
// the same arithmetic, so all CTAs take identical branches.
template <bool BOUNDS>
__global__ void __launch_bounds__(1024, 1) k_solve_persistent(DevLP lp, int G, double eta, double w0, int max_iters,
                                                              int check_every, double tol, double* out, const double* w0_dev)
{
    // w0_dev != null: the PDLP default ||c~||_2 / ||b~||_2 from the squared norms {||b~||^2, ||c~||^2} left there by the host side
    if (w0_dev) {
        const double nb2 = __ldcg(w0_dev), nc2 = __ldcg(w0_dev + 1);
        w0 = (nb2 > 0.0 && nc2 > 0.0) ? sqrt(nc2 / nb2) : 1.0;
    }
    __shared__ double smem[32 * NRED];
    __shared__ double bc[16];
    extern __shared__ __align__(16) unsigned char dsm[];
    PersistentSmem P;
    persistent_setup(lp, dsm, P);
    const MatView& VA = P.VA;
    const MatView& VAT = P.VAT;
    unsigned target = 0;
    const size_t RS = (size_t)G * NRED;  // one reduction buffer
    // x0 = x, y0 = y
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < lp.n; k += gridDim.x * blockDim.x) lp.x0[k] = __ldcg(lp.x + k);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < lp.m; k += gridDim.x * blockDim.x) lp.y0[k] = __ldcg(lp.y + k);
    sync_all(lp, target);

    double w = w0, fpe_restart = -1.0, fpe_prev = INFINITY, fpe = 0.0;
    int k = 0, it = 0, restarts = 0, converged = 0;
    double kk[10];
    for (int q = 0; q < 10; ++q) kk[q] = 0.0;

    while (it < max_iters) {
        const double tau = eta / w, sigma = eta * w;
        const double lam = (double)(k + 1) / (double)(k + 2);
        const bool check = ((it + 1) % check_every == 0) || (it + 1 == max_iters);
        const bool need_fpe = check || fpe_restart < 0.0;
        double* redp = lp.red + (size_t)((it & 1) * 4 + RED_STEPP) * RS;
        double* redd = lp.red + (size_t)((it & 1) * 4 + RED_STEPD) * RS;
        {
            PrimalHalpernOp<BOUNDS> op{lp, tau, lam};
            double acc[NRED];
            acc[0] = 0.0;
            phase_AT(lp, P, op, acc);
            if (need_fpe) cta_reduce_store<1>(acc, redp + (size_t)blockIdx.x * NRED, smem);
        }
        sync_all(lp, target);
        {
            DualHalpernOp<BOUNDS> op{lp, sigma, lam};
            double acc[NRED];
            acc[0] = 0.0;
            phase_A(lp, P, op, acc);
            if (need_fpe) cta_reduce_store<1>(acc, redd + (size_t)blockIdx.x * NRED, smem);
        }
        if (check) {
            // KKT at the new iterate needs the complete x and y
            sync_all(lp, target);
            eval_phases<BOUNDS>(lp, VA, VAT, lp.red + (size_t)((it & 1) * 4 + RED_EVALP) * RS,
                                lp.red + (size_t)((it & 1) * 4 + RED_EVALD) * RS, smem);
        }
        sync_all(lp, target);
        ++it; ++k;
        if (need_fpe) {
            if (threadIdx.x < 32) {
                const double dx2 = grid_sum(redp, G, 0), dy2 = grid_sum(redd, G, 0);
                if (threadIdx.x == 0) bc[0] = sqrt(w * dx2 + dy2 / w);
            }
            __syncthreads();
            fpe = bc[0];
            __syncthreads();
            if (fpe_restart < 0.0) fpe_restart = fpe;
        }
        if (check) {
            const double* ep = lp.red + (size_t)(((it - 1) & 1) * 4 + RED_EVALP) * RS;
            const double* ed = lp.red + (size_t)(((it - 1) & 1) * 4 + RED_EVALD) * RS;
            if (threadIdx.x < 32) {
                double s[10];
                kkt_from_sums(ep, ed, G, s);
                const double ddx2 = grid_sum(ep, G, 5), ddy2 = grid_sum(ed, G, 4);
                if (threadIdx.x == 0) {
                    for (int q = 0; q < 10; ++q) bc[q] = s[q];
                    bc[10] = ddx2; bc[11] = ddy2;
                }
            }
            __syncthreads();
            for (int q = 0; q < 10; ++q) kk[q] = bc[q];
            const double ddx = sqrt(bc[10]), ddy = sqrt(bc[11]);
            __syncthreads();
            if (kk[8] <= tol) { converged = 1; break; }
            const bool do_restart = (fpe <= 0.2 * fpe_restart) || (fpe <= 0.8 * fpe_restart && fpe > fpe_prev) ||
                                    ((double)k >= 0.36 * (double)it);
            fpe_prev = fpe;
            if (do_restart) {
                if (ddx > 1e-10 && ddy > 1e-10) w = exp(0.5 * log(ddy / ddx) + 0.5 * log(w));
                for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < lp.n; q += gridDim.x * blockDim.x) lp.x0[q] = __ldcg(lp.x + q);
                for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < lp.m; q += gridDim.x * blockDim.x) lp.y0[q] = __ldcg(lp.y + q);
                sync_all(lp, target);
                k = 0; fpe_restart = -1.0; fpe_prev = INFINITY;
                ++restarts;
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int q = 0; q < 10; ++q) out[q] = kk[q];
        out[10] = (double)it; out[11] = (double)restarts; out[12] = (double)converged;
        out[13] = w; out[14] = fpe; out[15] = 0.0;
    }
}

// ---------------------------------------------------------------------------------------
// SYNC_BCAST kernels: the grid is ONE thread-block cluster.  Dynamic shared memory:
//   ycopy[m~] | xbcopy[n~] | ys | bs | y0s  (own_rows_A each) | xs | cs | x0s  (own_rows_AT each) | matrix views
struct ClusterSmem {
    double *ys, *bs, *y0s, *xs, *cs, *x0s, *ycopy, *xbcopy;
};

__device__ __forceinline__ void cluster_barrier()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void cluster_setup(const DevLP& lp, unsigned char* dsm, PersistentSmem& P, ClusterSmem& C, bool anchors)
{
    // the copies come first: their shared-memory addresses must be the same in every CTA of the cluster (the matrix
    // views that follow have CTA-dependent sizes)
    double* q = reinterpret_cast<double*>(dsm);
    const uint32_t ra = lp.own_rows_A, rat = lp.own_rows_AT;
    C.ycopy = q; q += (lp.m + 1) & ~1;
    C.xbcopy = q; q += (lp.n + 1) & ~1;
    C.ys = q; q += ra; C.bs = q; q += ra; C.y0s = q; q += ra;
    C.xs = q; q += rat; C.cs = q; q += rat; C.x0s = q; q += rat;
    persistent_setup(lp, reinterpret_cast<unsigned char*>(q), P, false);
    if (threadIdx.x == 0) assign_row_slots(P.VA);
    if (threadIdx.x == 32) assign_row_slots(P.VAT);
    __syncthreads();
    copy_own<true>(P.VA, lp.y, C.ys);
    copy_own<true>(P.VA, const_cast<double*>(lp.b), C.bs);
    copy_own<true>(P.VAT, lp.x, C.xs);
    copy_own<true>(P.VAT, const_cast<double*>(lp.c), C.cs);
    if (anchors) {
        copy_own<true>(P.VA, lp.y0, C.y0s);
        copy_own<true>(P.VAT, lp.x0, C.x0s);
    }
    for (int k = threadIdx.x; k < lp.m; k += blockDim.x) C.ycopy[k] = __ldcg(lp.y + k);
    // every CTA of the cluster has started and initialised its copies before anyone stores into them
    cluster_barrier();
}

template <bool BOUNDS>
__global__ void __launch_bounds__(1024, 1) k_pdhg_cluster(DevLP lp, double tau, double sigma, int iters)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    PersistentSmem P;
    ClusterSmem C;
    cluster_setup(lp, dsm, P, C, false);
    const int nctas = (int)gridDim.x;
    PrimalBcastOp<BOUNDS> pop{lp, tau, C.xs, C.cs, C.ycopy, Bcast{smem_u32(C.xbcopy), nctas}};
    DualBcastOp<BOUNDS> dop{lp, sigma, C.ys, C.bs, C.xbcopy, Bcast{smem_u32(C.ycopy), nctas}};
    unsigned target = 0;
    double acc[NRED];
    for (int it = 0; it < iters; ++it) {
        unsigned long long* tr = lp.trace ? lp.trace + ((size_t)it * gridDim.x + blockIdx.x) * 4 : nullptr;
        phase_AT(lp, P, pop, acc);
        sync_all(lp, target, tr);
        phase_A(lp, P, dop, acc);
        sync_all(lp, target, tr ? tr + 2 : nullptr);
    }
    copy_own<false>(P.VAT, lp.x, C.xs);
    copy_own<false>(P.VA, lp.y, C.ys);
}

// Solve mode in the SYNC_BCAST geometry: k_solve_persistent's control flow; the iterates reach global memory only
// for the KKT checks (every check_every iterations), which run on the global-memory evaluation path.
template <bool BOUNDS>
__global__ void __launch_bounds__(1024, 1) k_solve_cluster(DevLP lp, int G, double eta, double w0, int max_iters,
                                                           int check_every, double tol, double* out, const double* w0_dev)
{
    if (w0_dev) {
        const double nb2 = __ldcg(w0_dev), nc2 = __ldcg(w0_dev + 1);
        w0 = (nb2 > 0.0 && nc2 > 0.0) ? sqrt(nc2 / nb2) : 1.0;
    }
    __shared__ double smem[32 * NRED];
    __shared__ double bc[16];
    extern __shared__ __align__(16) unsigned char dsm[];
    PersistentSmem P;
    ClusterSmem C;
    // x0 = x, y0 = y (global), then the slots are filled from global memory
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < lp.n; k += gridDim.x * blockDim.x) lp.x0[k] = __ldcg(lp.x + k);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < lp.m; k += gridDim.x * blockDim.x) lp.y0[k] = __ldcg(lp.y + k);
    cluster_barrier();
    cluster_setup(lp, dsm, P, C, true);
    const MatView& VA = P.VA;
    const MatView& VAT = P.VAT;
    const int nctas = (int)gridDim.x;
    const Bcast xb{smem_u32(C.xbcopy), nctas}, yb{smem_u32(C.ycopy), nctas};
    unsigned target = 0;
    const size_t RS = (size_t)G * NRED;

    double w = w0, fpe_restart = -1.0, fpe_prev = INFINITY, fpe = 0.0;
    int k = 0, it = 0, restarts = 0, converged = 0;
    double kk[10];
    for (int q = 0; q < 10; ++q) kk[q] = 0.0;

    while (it < max_iters) {
        const double tau = eta / w, sigma = eta * w;
        const double lam = (double)(k + 1) / (double)(k + 2);
        const bool check = ((it + 1) % check_every == 0) || (it + 1 == max_iters);
        const bool need_fpe = check || fpe_restart < 0.0;
        double* redp = lp.red + (size_t)((it & 1) * 4 + RED_STEPP) * RS;
        double* redd = lp.red + (size_t)((it & 1) * 4 + RED_STEPD) * RS;
        {
            PrimalHalpernBcastOp<BOUNDS> op{lp, tau, lam, C.xs, C.cs, C.x0s, C.ycopy, xb};
            double acc[NRED];
            acc[0] = 0.0;
            phase_AT(lp, P, op, acc);
            if (need_fpe) cta_reduce_store<1>(acc, redp + (size_t)blockIdx.x * NRED, smem);
        }
        sync_all(lp, target);
        {
            DualHalpernBcastOp<BOUNDS> op{lp, sigma, lam, C.ys, C.bs, C.y0s, C.xbcopy, yb};
            double acc[NRED];
            acc[0] = 0.0;
            phase_A(lp, P, op, acc);
            if (need_fpe) cta_reduce_store<1>(acc, redd + (size_t)blockIdx.x * NRED, smem);
        }
        if (check) {
            // KKT at the new iterate: publish x and y, then the global-memory evaluation path
            copy_own<false>(VAT, lp.x, C.xs);
            copy_own<false>(VA, lp.y, C.ys);
            sync_all(lp, target);
            eval_phases<BOUNDS>(lp, VA, VAT, lp.red + (size_t)((it & 1) * 4 + RED_EVALP) * RS,
                                lp.red + (size_t)((it & 1) * 4 + RED_EVALD) * RS, smem);
        }
        sync_all(lp, target);
        ++it; ++k;
        if (need_fpe) {
            if (threadIdx.x < 32) {
                const double dx2 = grid_sum(redp, G, 0), dy2 = grid_sum(redd, G, 0);
                if (threadIdx.x == 0) bc[0] = sqrt(w * dx2 + dy2 / w);
            }
            __syncthreads();
            fpe = bc[0];
            __syncthreads();
            if (fpe_restart < 0.0) fpe_restart = fpe;
        }
        if (check) {
            const double* ep = lp.red + (size_t)(((it - 1) & 1) * 4 + RED_EVALP) * RS;
            const double* ed = lp.red + (size_t)(((it - 1) & 1) * 4 + RED_EVALD) * RS;
            if (threadIdx.x < 32) {
                double s[10];
                kkt_from_sums(ep, ed, G, s);
                const double ddx2 = grid_sum(ep, G, 5), ddy2 = grid_sum(ed, G, 4);
                if (threadIdx.x == 0) {
                    for (int q = 0; q < 10; ++q) bc[q] = s[q];
                    bc[10] = ddx2; bc[11] = ddy2;
                }
            }
            __syncthreads();
            for (int q = 0; q < 10; ++q) kk[q] = bc[q];
            const double ddx = sqrt(bc[10]), ddy = sqrt(bc[11]);
            __syncthreads();
            if (kk[8] <= tol) { converged = 1; break; }
            const bool do_restart = (fpe <= 0.2 * fpe_restart) || (fpe <= 0.8 * fpe_restart && fpe > fpe_prev) ||
                                    ((double)k >= 0.36 * (double)it);
            fpe_prev = fpe;
            if (do_restart) {
                if (ddx > 1e-10 && ddy > 1e-10) w = exp(0.5 * log(ddy / ddx) + 0.5 * log(w));
                // x and y were published for the check: anchors in global memory and in the slots
                for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < lp.n; q += gridDim.x * blockDim.x) lp.x0[q] = __ldcg(lp.x + q);
                for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < lp.m; q += gridDim.x * blockDim.x) lp.y0[q] = __ldcg(lp.y + q);
                for (uint32_t q = threadIdx.x; q < lp.own_rows_AT; q += blockDim.x) C.x0s[q] = C.xs[q];
                for (uint32_t q = threadIdx.x; q < lp.own_rows_A; q += blockDim.x) C.y0s[q] = C.ys[q];
                sync_all(lp, target);
                k = 0; fpe_restart = -1.0; fpe_prev = INFINITY;
                ++restarts;
            }
        }
    }
    copy_own<false>(VAT, lp.x, C.xs);
    copy_own<false>(VA, lp.y, C.ys);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int q = 0; q < 10; ++q) out[q] = kk[q];
        out[10] = (double)it; out[11] = (double)restarts; out[12] = (double)converged;
        out[13] = w; out[14] = fpe; out[15] = 0.0;
    }
}

// ---------------------------------------------------------------------------------------
// host launchers
#define CK(call)                                  \
    do {                                          \
        cudaError_t e_ = (call);                  \
        if (e_ != cudaSuccess) return (int)e_;    \
    } while (0)

static inline int blocks_for(int n) { return n <= 0 ? 1 : (n + 255) / 256 > 1184 ? 1184 : (n + 255) / 256; }

int launch_gather(double* dst, const double* src, const int32_t* order, int n, cudaStream_t s)
{
    if (n <= 0) return 0;
    count_launch(1);
    k_gather<<<blocks_for(n), 256, 0, s>>>(dst, src, order, n);
    return (int)cudaGetLastError();
}
int launch_scatter(double* dst, const double* src, const int32_t* order, int n, cudaStream_t s)
{
    if (n <= 0) return 0;
    count_launch(1);
    k_scatter<<<blocks_for(n), 256, 0, s>>>(dst, src, order, n);
    return (int)cudaGetLastError();
}
int launch_gather_scaled(double* dst, const double* src, const int32_t* order, const double* sc, int div, int n, cudaStream_t s)
{
    if (!sc) return launch_gather(dst, src, order, n, s);
    if (n <= 0) return 0;
    count_launch(1);
    k_gather_scaled<<<blocks_for(n), 256, 0, s>>>(dst, src, order, sc, div, n);
    return (int)cudaGetLastError();
}
int launch_scatter_scaled(double* dst, const double* src, const int32_t* order, const double* sc, int div, int n, cudaStream_t s)
{
    if (!sc) return launch_scatter(dst, src, order, n, s);
    if (n <= 0) return 0;
    count_launch(1);
    k_scatter_scaled<<<blocks_for(n), 256, 0, s>>>(dst, src, order, sc, div, n);
    return (int)cudaGetLastError();
}
int launch_fill(double* dst, double v, int n, cudaStream_t s)
{
    if (n <= 0) return 0;
    count_launch(1);
    k_fill<<<blocks_for(n), 256, 0, s>>>(dst, v, n);
    return (int)cudaGetLastError();
}
int launch_sumsq(const double* v, int n, double* out, double* scratch /* >= SUMSQ_BLOCKS doubles */, cudaStream_t s)
{
    count_launch(2);
    k_sumsq_partial<<<SUMSQ_BLOCKS, 256, 0, s>>>(v, n, scratch);
    k_sumsq_final<<<1, 256, 0, s>>>(scratch, SUMSQ_BLOCKS, out);
    return (int)cudaGetLastError();
}
int launch_scale_by_invnorm(double* dst, const double* src, const double* norm2, int n, cudaStream_t s)
{
    if (n <= 0) return 0;
    count_launch(1);
    k_scale_by_invnorm<<<blocks_for(n), 256, 0, s>>>(dst, src, norm2, n);
    return (int)cudaGetLastError();
}

int launch_spmv(const DevMat& M, const double* in, double* out, int G, int threads, cudaStream_t s)
{
    count_launch(1);
    k_spmv<<<G, threads, 0, s>>>(M, in, out);
    return (int)cudaGetLastError();
}

int launch_primal(const DevLP& lp, bool bounds, int G, int threads, cudaStream_t s)
{
    count_launch(1);
    if (bounds) k_primal<true><<<G, threads, 0, s>>>(lp);
    else k_primal<false><<<G, threads, 0, s>>>(lp);
    return (int)cudaGetLastError();
}
int launch_dual(const DevLP& lp, bool bounds, int G, int threads, cudaStream_t s)
{
    count_launch(1);
    if (bounds) k_dual<true><<<G, threads, 0, s>>>(lp);
    else k_dual<false><<<G, threads, 0, s>>>(lp);
    return (int)cudaGetLastError();
}
int launch_eval_partial(const DevLP& lp, bool bounds, int G, int threads, cudaStream_t s)
{
    count_launch(1);
    if (bounds) k_eval<true><<<G, threads, 0, s>>>(lp, G);
    else k_eval<false><<<G, threads, 0, s>>>(lp, G);
    return (int)cudaGetLastError();
}
int launch_eval_finalize(const DevLP& lp, int G, double* out, double iters, cudaStream_t s)
{
    count_launch(1);
    k_eval_finalize<<<1, 32, 0, s>>>(lp, G, out, iters);
    return (int)cudaGetLastError();
}
int launch_eval(const DevLP& lp, bool bounds, int G, int threads, double* out, double iters, cudaStream_t s)
{
    count_launch(2);
    if (bounds) k_eval<true><<<G, threads, 0, s>>>(lp, G);
    else k_eval<false><<<G, threads, 0, s>>>(lp, G);
    CK(cudaGetLastError());
    k_eval_finalize<<<1, 32, 0, s>>>(lp, G, out, iters);
    return (int)cudaGetLastError();
}

static const void* persistent_fn(bool solve, bool bounds, bool bcast = false)
{
    if (bcast) {
        if (solve) return bounds ? (const void*)k_solve_cluster<true> : (const void*)k_solve_cluster<false>;
        return bounds ? (const void*)k_pdhg_cluster<true> : (const void*)k_pdhg_cluster<false>;
    }
    if (solve) return bounds ? (const void*)k_solve_persistent<true> : (const void*)k_solve_persistent<false>;
    return bounds ? (const void*)k_pdhg_persistent<true> : (const void*)k_pdhg_persistent<false>;
}

int persistent_set_smem(bool bounds, size_t dyn_smem)
{
    for (int k = 0; k < 4; ++k) {
        cudaError_t e = cudaFuncSetAttribute(persistent_fn(k & 1, bounds, k >> 1), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem);
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}

int persistent_max_blocks_per_sm(int threads, bool bounds, size_t dyn_smem)
{
    int best = 1 << 30;
    for (int solve = 0; solve < 2; ++solve) {
        int nb = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, persistent_fn(solve, bounds), threads, dyn_smem);
        if (e != cudaSuccess) return -(int)e;
        if (nb < best) best = nb;
    }
    return best;
}

// One launch of a persistent kernel in the geometry lp.sync_mode names: cooperative grid, ONE cluster of G CTAs, or
// one CTA.
static int launch_persistent_fn(const void* fn, int sync_mode, int G, int threads, size_t dyn_smem, void** args, cudaStream_t s)
{
    count_launch(1);
    if (sync_mode == SYNC_GRID) {
        CK(cudaLaunchCooperativeKernel(fn, dim3(G), dim3(threads), args, dyn_smem, s));
        return 0;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(G);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = dyn_smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = sync_mode == SYNC_CTA ? 1 : G;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CK(cudaLaunchKernelExC(&cfg, fn, args));
    return 0;
}

// Can ONE cluster of `ctas` CTAs of the persistent kernels (both modes) be resident with this much shared memory?
// (Clusters above 8 CTAs are non-portable and need the opt-in attribute.)  Returns 1 / 0, or a negative CUDA error.
int persistent_cluster_fits(int ctas, int threads, bool bounds, size_t dyn_smem)
{
    for (int k = 0; k < 4; ++k) {
        const void* fn = persistent_fn(k & 1, bounds, k >> 1);
        if (ctas > 8) {
            cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            if (e != cudaSuccess) { cudaGetLastError(); return 0; }
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(ctas);
        cfg.blockDim = dim3(threads);
        cfg.dynamicSmemBytes = dyn_smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = ctas;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        int n = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, fn, &cfg);
        if (e != cudaSuccess) { cudaGetLastError(); return 0; }
        if (n < 1) return 0;
    }
    return 1;
}

int launch_pdhg_persistent(const DevLP& lp, bool bounds, int G, int threads, size_t dyn_smem, double tau, double sigma,
                           int iters, cudaStream_t s)
{
    if (lp.sync_mode == SYNC_GRID) CK(cudaMemsetAsync(lp.barrier, 0, sizeof(unsigned), s));
    DevLP lpv = lp;
    void* args[] = {&lpv, &tau, &sigma, &iters};
    return launch_persistent_fn(persistent_fn(false, bounds, lp.sync_mode == SYNC_BCAST), lp.sync_mode, G, threads, dyn_smem, args, s);
}

static const void* rowpart_fn(bool bounds)
{
    return bounds ? (const void*)k_pdhg_rowpart<true> : (const void*)k_pdhg_rowpart<false>;
}
int rowpart_set_smem(bool bounds, size_t dyn_smem)
{
    return (int)cudaFuncSetAttribute(rowpart_fn(bounds), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem);
}

int launch_pdhg_rowpart(const DevLP& lp, const PeerInfo& pi, bool bounds, int G, int threads, size_t dyn_smem,
                        double tau, double sigma, int iters, unsigned long long seq, cudaStream_t s)
{
    CK(cudaMemsetAsync(lp.barrier, 0, sizeof(unsigned), s));
    DevLP lpv = lp;
    PeerInfo piv = pi;
    void* args[] = {&lpv, &piv, &tau, &sigma, &iters, &seq};
    count_launch(1);
    CK(cudaLaunchCooperativeKernel(rowpart_fn(bounds), dim3(G), dim3(threads), args, dyn_smem, s));
    return 0;
}

int launch_solve_persistent(const DevLP& lp, bool bounds, int G, int threads, size_t dyn_smem, double eta, double w0,
                            int max_iters, int check_every, double tol, double* out, const double* w0_dev, cudaStream_t s)
{
    if (lp.sync_mode == SYNC_GRID) CK(cudaMemsetAsync(lp.barrier, 0, sizeof(unsigned), s));
    DevLP lpv = lp;
    void* args[] = {&lpv, &G, &eta, &w0, &max_iters, &check_every, &tol, &out, &w0_dev};
    return launch_persistent_fn(persistent_fn(true, bounds, lp.sync_mode == SYNC_BCAST), lp.sync_mode, G, threads, dyn_smem, args, s);
}

}  // namespace mllp
