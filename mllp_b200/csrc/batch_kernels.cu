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
    const double* dr;       // preconditioned batch (MLLP_F_PRECONDITION): the matrix held is Dr A Dc; internal order; else null
    const double* dc;
    int m, n;
    long long x_off, y_off; // offsets into the concatenated user vectors
    // row-per-lane images of A / A' (lp_format.h: HostEll) for the warp-per-instance solve kernel, or null
    const int32_t* eidxA; const double* evalA; const uint32_t* eoffA;
    const int32_t* eidxT; const double* evalT; const uint32_t* eoffT;
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
    lp.dr = I.dr; lp.dc = I.dc;
    return lp;
}

// Where the matrix lives for the CTA: with `use_res` the tile descriptors and a prefix of the warp-steps of
// A' and A are copied into shared memory behind the vectors (byte offset mat_off) -- once per CTA for a shared
// matrix, once per instance otherwise -- and the rest streams through L1 / L2; without it everything streams.
struct BatchGeom {
    uint32_t mat_off;
    uint32_t res_A, res_AT;
    int use_res;
};
__device__ __forceinline__ void batch_views(const BatchInst& I, const BatchGeom& g, unsigned char* dsm, MatView& VA, MatView& VAT)
{
    if (g.use_res) {
        __syncthreads();   // the previous instance's tiles are no longer read
        unsigned char* mb = dsm + g.mat_off;
        const uint32_t used = resident_view(I.AT, g.res_AT, mb, VAT, 0u);
        resident_view(I.A, g.res_A, mb + used, VA, 0u);
        __syncthreads();
    } else {
        VA = global_view(I.A, 0u);
        VAT = global_view(I.AT, 0u);
    }
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
    // preconditioned batch: the caller's vectors are those of the ORIGINAL LP (x~ = x / dc, c~ = dc c, y~ = y / dr, b~ = dr b)
    for (int k = threadIdx.x; k < I.n; k += blockDim.x) {
        const int j = __ldg(I.orderX + k);
        const double sc = I.dc ? __ldg(I.dc + k) : 1.0;
        S.x[k] = x[I.x_off + j] / sc;
        S.c[k] = c[I.x_off + j] * sc;
        S.xbar[k] = 0.0;
    }
    for (int k = threadIdx.x; k < I.m; k += blockDim.x) {
        const int i = __ldg(I.orderY + k);
        const double sc = I.dr ? __ldg(I.dr + k) : 1.0;
        S.y[k] = y[I.y_off + i] / sc;
        S.b[k] = b[I.y_off + i] * sc;
    }
    __syncthreads();
}
__device__ __forceinline__ void store_instance(const BatchInst& I, const BatchSmem& S, double* x, double* y)
{
    for (int k = threadIdx.x; k < I.n; k += blockDim.x) x[I.x_off + __ldg(I.orderX + k)] = S.x[k] * (I.dc ? __ldg(I.dc + k) : 1.0);
    for (int k = threadIdx.x; k < I.m; k += blockDim.x) y[I.y_off + __ldg(I.orderY + k)] = S.y[k] * (I.dr ? __ldg(I.dr + k) : 1.0);
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
    cta_allreduce<7>(ap, S);
    cta_allreduce<6>(ad, S);
    const double pobj = ap[0], dobj = ad[0] + ap[1];
    s[0] = pobj; s[1] = dobj; s[2] = sqrt(ad[1] + ap[6]); s[3] = sqrt(ap[2] + ad[5]);
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
k_batch_run(const BatchInst* __restrict__ insts, int count, int shared, BatchGeom geom, double* x, double* y, const double* b,
            const double* c, const double* tau, const double* sigma, int iters, double* scalars)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    MatView VA, VAT;
    bool have_views = false;
    for (int inst = blockIdx.x; inst < count; inst += gridDim.x) {
        BatchInst I = insts[shared ? 0 : inst];
        if (shared) { I.x_off = (long long)inst * I.n; I.y_off = (long long)inst * I.m; }
        const BatchSmem S = carve(dsm, I.m, I.n, true);
        load_instance(I, S, x, y, b, c);
        for (int k = threadIdx.x; k < I.n; k += blockDim.x) S.x0[k] = 0.0;   // anchors unused here; keep eval finite
        for (int k = threadIdx.x; k < I.m; k += blockDim.x) S.y0[k] = 0.0;
        const DevLP lp = smem_lp(I, S);
        if (!have_views || !shared) { batch_views(I, geom, dsm, VA, VAT); have_views = true; }
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
k_batch_solve(const BatchInst* __restrict__ insts, int count, int shared, BatchGeom geom, double* x, double* y,
              const double* b, const double* c, const double* eta_arr, double w0, int max_iters, int check_every, double tol,
              double* scalars, int* next_inst)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    __shared__ int s_inst;
    __shared__ double s_par[2][3];   // [buffer][tau, sigma, lam]: computed by one thread, one iteration ahead (see k_batch_solve_r)
    MatView VA, VAT;
    bool have_views = false;
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
        const double eta = __ldg(eta_arr + inst);
        int pb = 0;
        double w_init = w0;
        if (!(w0 > 0.0)) {   // the PDLP default ||c~|| / ||b~|| of this instance (scaled data, as loaded)
            double nn[2] = {0.0, 0.0};
            for (int q = threadIdx.x; q < I.m; q += blockDim.x) nn[0] += S.b[q] * S.b[q];
            for (int q = threadIdx.x; q < I.n; q += blockDim.x) nn[1] += S.c[q] * S.c[q];
            cta_allreduce<2>(nn, S);
            w_init = (nn[0] > 0.0 && nn[1] > 0.0) ? sqrt(nn[1] / nn[0]) : 1.0;
        }
        if (threadIdx.x == 0) { s_par[0][0] = eta / w_init; s_par[0][1] = eta * w_init; s_par[0][2] = 0.5; }
        __syncthreads();
        const DevLP lp = smem_lp(I, S);
        if (!have_views || !shared) { batch_views(I, geom, dsm, VA, VAT); have_views = true; }
        double w = w_init, fpe_restart = -1.0, fpe_prev = INFINITY, fpe = 0.0;
        int k = 0, it = 0, restarts = 0, converged = 0;
        double kk[10], dd[2];
        batch_kkt(lp, VA, VAT, S, kk, dd);
        while (it < max_iters) {
            const double tau = s_par[pb][0], sigma = s_par[pb][1], lam = s_par[pb][2];
            if (threadIdx.x == 0) {   // the next iteration's parameters if nothing restarts
                s_par[pb ^ 1][0] = tau; s_par[pb ^ 1][1] = sigma;
                s_par[pb ^ 1][2] = (double)(k + 2) / (double)(k + 3);
            }
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
            pb ^= 1;
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
                    if (threadIdx.x == 0) { s_par[pb][0] = eta / w; s_par[pb][1] = eta * w; s_par[pb][2] = 0.5; }
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

// ---------------------------------------------------------------------------------------
// Shared matrix, R instances per CTA ("multi-RHS"): the R instances' vectors are interleaved in shared memory
// (entry k of instance r at [k * R + r]), every matrix step (values + indices) is fetched ONCE and applied to all
// R instances, so the matrix stream from L2 / shared memory, the descriptor decoding and the address arithmetic
// are paid once per R instance-iterations.  Per instance the summation order is exactly that of the one-instance
// walker (run_phase), so the iterates are bitwise the same for every R.
template <int R>
struct SmemR {
    double *x, *xbar, *c, *y, *b, *x0, *y0;
    double* red;    // 32 * NRED
    double* bc;     // 16 * R
    double* spart;  // SPLIT_SLOTS * R
};
template <int R>
__device__ __forceinline__ SmemR<R> carve_r(unsigned char* dsm, int m, int n, bool anchors)
{
    SmemR<R> S;
    double* p = reinterpret_cast<double*>(dsm);
    S.red = p; p += 32 * NRED;
    S.bc = p; p += 16 * R;
    S.spart = p; p += SPLIT_SLOTS * R;
    S.x = p; p += (size_t)n * R; S.xbar = p; p += (size_t)n * R; S.c = p; p += (size_t)n * R;
    S.y = p; p += (size_t)m * R; S.b = p; p += (size_t)m * R;
    S.x0 = anchors ? p : nullptr; if (anchors) p += (size_t)n * R;
    S.y0 = anchors ? p : nullptr;
    return S;
}

template <int R>
__device__ __forceinline__ void ld_r(const double* __restrict__ p, double* g)
{
    if (R % 2 == 0) {
#pragma unroll
        for (int r = 0; r < R; r += 2) {
            const double2 t = *reinterpret_cast<const double2*>(p + r);
            g[r] = t.x; g[r + 1] = t.y;
        }
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r) g[r] = p[r];
    }
}

template <int R, bool RES>
__device__ __forceinline__ void tile_dot_r_at(const double2* __restrict__ vp, const int2* __restrict__ ip,
                                              const double* __restrict__ vec, int nsteps, double* dot)
{
    if (RES) {
        __builtin_assume(__isShared(vp));
        __builtin_assume(__isShared(ip));
    } else {
        __builtin_assume(__isGlobal(vp));
        __builtin_assume(__isGlobal(ip));
    }
    __builtin_assume(__isShared(vec));
#pragma unroll 2
    for (int s = 0; s < nsteps; ++s) {
        const int2 j = ip[s * 32];
        const double2 v = vp[s * 32];
        double g0[R], g1[R];
        ld_r<R>(vec + (size_t)j.x * R, g0);
        ld_r<R>(vec + (size_t)j.y * R, g1);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            dot[r] = fma(v.x, g0[r], dot[r]);
            dot[r] = fma(v.y, g1[r], dot[r]);
        }
    }
}
template <int R>
__device__ __forceinline__ void tile_dot_r(const MatView& V, const double* __restrict__ vec, uint32_t off, int nsteps, int lane,
                                           double* dot)
{
    const double2* vp;
    const int2* ip;
    if (tile_ptrs(V, off, nsteps, lane, vp, ip)) tile_dot_r_at<R, true>(vp, ip, vec, nsteps, dot);
    else tile_dot_r_at<R, false>(vp, ip, vec, nsteps, dot);
}

// Walker for a one-CTA grid (every split row is joined inside the CTA).  Op: vec(), row(i, dot[R], acc).
template <int R, class Op>
__device__ __forceinline__ void run_phase_r(const DevMat& M, const MatView& V, const Op& op, double* spart, double* acc)
{
    const double* __restrict__ vec = op.vec();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    if (V.nsplit > 0) {
        for (uint32_t t = warp; t < V.nsplit; t += nwarps) {
            const int4 raw = *reinterpret_cast<const int4*>(V.desc + t);
            double dot[R];
#pragma unroll
            for (int r = 0; r < R; ++r) dot[r] = 0.0;
            tile_dot_r<R>(V, vec, (uint32_t)raw.x, raw.z & 0xffff, lane, dot);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                for (int o = 16; o > 0; o >>= 1) dot[r] += __shfl_xor_sync(FULL, dot[r], o);
                if (lane == 0) spart[raw.w * R + r] = dot[r];
            }
        }
        __syncthreads();
        for (uint32_t li = warp; li < V.nls; li += nwarps) {
            const int4 ls = __ldg(reinterpret_cast<const int4*>(M.lsplits + V.ls0 + li));
            const int4 sr = __ldg(reinterpret_cast<const int4*>(M.splits + ls.x));
            const int first = ls.z & 0xffff, count = (ls.z >> 16) & 0xffff;
            double p[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                p[r] = 0.0;
                for (int k = lane; k < count; k += 32) p[r] += spart[(first + k) * R + r];
                for (int o = 16; o > 0; o >>= 1) p[r] += __shfl_xor_sync(FULL, p[r], o);
            }
            if (lane == 0) op.row(sr.x, p, acc);
        }
    }
    for (uint32_t t = V.nsplit + warp; t < V.ntiles; t += nwarps) {
        const int4 raw = *reinterpret_cast<const int4*>(V.desc + t);
        const int nsteps = raw.z & 0xffff;
        const int logL = (raw.z >> 16) & 0xff;
        const int nrows = (raw.z >> 24) & 0xff;
        const int L = 1 << logL;
        const int rr = lane >> logL;
        const bool owner = ((lane & (L - 1)) == 0) && (rr < nrows);
        double dot[R];
#pragma unroll
        for (int r = 0; r < R; ++r) dot[r] = 0.0;
        tile_dot_r<R>(V, vec, (uint32_t)raw.x, nsteps, lane, dot);
#pragma unroll
        for (int r = 0; r < R; ++r)
            for (int o = L >> 1; o > 0; o >>= 1) dot[r] += __shfl_xor_sync(FULL, dot[r], o);
        if (owner) op.row(raw.y + rr, dot, acc);
    }
}

// R-wide row updates (same arithmetic as PrimalOp / DualOp / *HalpernOp / Eval*Op, instance r at [i * R + r])
template <int R>
struct PrimalR {
    const SmemR<R>& S;
    const double* tau;      // [R]
    __device__ __forceinline__ const double* vec() const { return S.y; }
    __device__ __forceinline__ void row(int i, const double* dot, double*) const
    {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const double xx = S.x[(size_t)i * R + r];
            const double g = S.c[(size_t)i * R + r] - dot[r];
            const double xn = fmax(xx - tau[r] * g, 0.0);
            S.xbar[(size_t)i * R + r] = 2.0 * xn - xx;
            S.x[(size_t)i * R + r] = xn;
        }
    }
};
template <int R>
struct DualR {
    const SmemR<R>& S;
    const double* sigma;
    __device__ __forceinline__ const double* vec() const { return S.xbar; }
    __device__ __forceinline__ void row(int i, const double* dot, double*) const
    {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const double yy = S.y[(size_t)i * R + r];
            S.y[(size_t)i * R + r] = yy + sigma[r] * (S.b[(size_t)i * R + r] - dot[r]);
        }
    }
};
// solve mode; `active[r]` = 0 freezes instance r (it has converged; the group runs on for the others)
template <int R>
struct PrimalHalpernR {
    const SmemR<R>& S;
    const double *tau, *lam;
    const int* active;
    __device__ __forceinline__ const double* vec() const { return S.y; }
    __device__ __forceinline__ void row(int i, const double* dot, double* acc) const
    {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (!active[r]) continue;
            const double xx = S.x[(size_t)i * R + r];
            const double g = S.c[(size_t)i * R + r] - dot[r];
            const double xn = fmax(xx - tau[r] * g, 0.0);
            const double d = xn - xx;
            acc[r] += d * d;
            const double xb = 2.0 * xn - xx;
            S.xbar[(size_t)i * R + r] = xb;
            S.x[(size_t)i * R + r] = lam[r] * xb + (1.0 - lam[r]) * S.x0[(size_t)i * R + r];
        }
    }
};
template <int R>
struct DualHalpernR {
    const SmemR<R>& S;
    const double *sigma, *lam;
    const int* active;
    __device__ __forceinline__ const double* vec() const { return S.xbar; }
    __device__ __forceinline__ void row(int i, const double* dot, double* acc) const
    {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (!active[r]) continue;
            const double yy = S.y[(size_t)i * R + r];
            const double yn = yy + sigma[r] * (S.b[(size_t)i * R + r] - dot[r]);
            const double d = yn - yy;
            acc[r] += d * d;
            S.y[(size_t)i * R + r] = lam[r] * (2.0 * yn - yy) + (1.0 - lam[r]) * S.y0[(size_t)i * R + r];
        }
    }
};
// Same sums as EvalPrimalOp / EvalDualOp (ORIGINAL LP on a preconditioned batch: dc / dr are the matrix's scaling
// vectors in internal order, or null).
// acc[r * 7 + k]: 0 pobj, 1 dobj bound terms (0 here: l = 0, u = inf), 2 dual residual^2, 3 ||c||^2, 4 ||x||^2,
//                 5 ||x~ - x~0||^2 (scaled), 6 distance^2 of x from x >= 0
template <int R>
struct EvalPrimalR {
    const SmemR<R>& S;
    const double* dc;
    __device__ __forceinline__ const double* vec() const { return S.y; }
    __device__ __forceinline__ void row(int i, const double* dot, double* acc) const
    {
        const double sc = dc ? __ldg(dc + i) : 1.0;
        const double inv = 1.0 / sc;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const double cc = S.c[(size_t)i * R + r], xx = S.x[(size_t)i * R + r], x0 = S.x0 ? S.x0[(size_t)i * R + r] : 0.0;
            const double rc = cc - dot[r];
            const double rn = rc < 0.0 ? rc : 0.0;
            const double xv = xx < 0.0 ? xx : 0.0;
            acc[r * 7 + 0] += cc * xx;
            acc[r * 7 + 2] += (rn * rn) * (inv * inv);
            acc[r * 7 + 3] += (cc * inv) * (cc * inv);
            acc[r * 7 + 4] += (xx * sc) * (xx * sc);
            acc[r * 7 + 5] += (xx - x0) * (xx - x0);
            acc[r * 7 + 6] += (xv * sc) * (xv * sc);
        }
    }
};
// acc[r * 7 + k]: 0 b'y, 1 primal residual^2, 2 ||b||^2, 3 ||y||^2, 4 ||y~ - y~0||^2 (scaled)
template <int R>
struct EvalDualR {
    const SmemR<R>& S;
    const double* dr;
    __device__ __forceinline__ const double* vec() const { return S.x; }
    __device__ __forceinline__ void row(int i, const double* dot, double* acc) const
    {
        const double sc = dr ? __ldg(dr + i) : 1.0;
        const double inv = 1.0 / sc;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const double bb = S.b[(size_t)i * R + r], yy = S.y[(size_t)i * R + r], y0 = S.y0 ? S.y0[(size_t)i * R + r] : 0.0;
            const double res = dot[r] - bb;
            acc[r * 7 + 0] += bb * yy;
            acc[r * 7 + 1] += (res * inv) * (res * inv);
            acc[r * 7 + 2] += (bb * inv) * (bb * inv);
            acc[r * 7 + 3] += (yy * sc) * (yy * sc);
            acc[r * 7 + 4] += (yy - y0) * (yy - y0);
        }
    }
};

// CTA-wide sums of acc[0..N) (fixed order), result in every thread; N may exceed NRED (done in slices)
template <int N, int R>
__device__ __forceinline__ void cta_allreduce_r(double* acc, const SmemR<R>& S)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int k0 = 0; k0 < N; k0 += NRED) {
        const int kn = N - k0 < NRED ? N - k0 : NRED;
        for (int k = 0; k < kn; ++k) {
            double v = acc[k0 + k];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
            if (lane == 0) S.red[warp * NRED + k] = v;
        }
        __syncthreads();
        if (warp == 0) {
            for (int k = 0; k < kn; ++k) {
                double v = lane < nwarps ? S.red[lane * NRED + k] : 0.0;
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
                if (lane == 0) S.bc[k] = v;
            }
        }
        __syncthreads();
        for (int k = 0; k < kn; ++k) acc[k0 + k] = S.bc[k];
        __syncthreads();
    }
}

// KKT scalars of the R instances: s[r * 10 + q], dd[r * 2 + q]
template <int R>
__device__ __forceinline__ void batch_kkt_r(const BatchInst& I, const MatView& VA, const MatView& VAT, const SmemR<R>& S,
                                            double* s, double* dd)
{
    double ap[R * 7], ad[R * 7];
#pragma unroll
    for (int k = 0; k < R * 7; ++k) { ap[k] = 0.0; ad[k] = 0.0; }
    { EvalPrimalR<R> op{S, I.dc}; run_phase_r<R>(I.AT, VAT, op, S.spart, ap); }
    __syncthreads();
    { EvalDualR<R> op{S, I.dr}; run_phase_r<R>(I.A, VA, op, S.spart, ad); }
    cta_allreduce_r<R * 7, R>(ap, S);
    cta_allreduce_r<R * 7, R>(ad, S);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const double* p = ap + r * 7;
        const double* d = ad + r * 7;
        double* o = s + r * 10;
        const double pobj = p[0], dobj = d[0] + p[1];
        o[0] = pobj; o[1] = dobj; o[2] = sqrt(d[1] + p[6]); o[3] = sqrt(p[2]);
        o[4] = sqrt(d[2]); o[5] = sqrt(p[3]); o[6] = sqrt(p[4]); o[7] = sqrt(d[3]);
        const double gap = fabs(pobj - dobj);
        double e = o[2] / (1.0 + o[4]);
        e = fmax(e, o[3] / (1.0 + o[5]));
        e = fmax(e, gap / (1.0 + fabs(pobj) + fabs(dobj)));
        o[8] = e; o[9] = gap;
        dd[r * 2 + 0] = p[5]; dd[r * 2 + 1] = d[4];
    }
}

template <int R>
__device__ __forceinline__ void load_group(const BatchInst& I, const SmemR<R>& S, int inst0, int count, const double* x,
                                           const double* y, const double* b, const double* c)
{
    for (int k = threadIdx.x; k < I.n; k += blockDim.x) {
        const int j = __ldg(I.orderX + k);
        const double sc = I.dc ? __ldg(I.dc + k) : 1.0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const bool ok = inst0 + r < count;
            const size_t at = (size_t)(inst0 + r) * I.n + j;
            S.x[(size_t)k * R + r] = ok ? x[at] / sc : 0.0;
            S.c[(size_t)k * R + r] = ok ? c[at] * sc : 0.0;
            S.xbar[(size_t)k * R + r] = 0.0;
        }
    }
    for (int k = threadIdx.x; k < I.m; k += blockDim.x) {
        const int i = __ldg(I.orderY + k);
        const double sc = I.dr ? __ldg(I.dr + k) : 1.0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const bool ok = inst0 + r < count;
            const size_t at = (size_t)(inst0 + r) * I.m + i;
            S.y[(size_t)k * R + r] = ok ? y[at] / sc : 0.0;
            S.b[(size_t)k * R + r] = ok ? b[at] * sc : 0.0;
        }
    }
    __syncthreads();
}
template <int R>
__device__ __forceinline__ void store_group(const BatchInst& I, const SmemR<R>& S, int inst0, int count, double* x, double* y)
{
    for (int k = threadIdx.x; k < I.n; k += blockDim.x) {
        const int j = __ldg(I.orderX + k);
        const double sc = I.dc ? __ldg(I.dc + k) : 1.0;
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (inst0 + r < count) x[(size_t)(inst0 + r) * I.n + j] = S.x[(size_t)k * R + r] * sc;
    }
    for (int k = threadIdx.x; k < I.m; k += blockDim.x) {
        const int i = __ldg(I.orderY + k);
        const double sc = I.dr ? __ldg(I.dr + k) : 1.0;
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (inst0 + r < count) y[(size_t)(inst0 + r) * I.m + i] = S.y[(size_t)k * R + r] * sc;
    }
    __syncthreads();
}

template <int R>
__global__ void __launch_bounds__(1024, 1)
k_batch_run_r(const BatchInst* __restrict__ insts, int count, BatchGeom geom, double* x, double* y, const double* b,
              const double* c, const double* tau, const double* sigma, int iters, double* scalars)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    const BatchInst I = insts[0];
    const SmemR<R> S = carve_r<R>(dsm, I.m, I.n, false);
    MatView VA, VAT;
    batch_views(I, geom, dsm, VA, VAT);
    const int ngroups = (count + R - 1) / R;
    for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const int inst0 = grp * R;
        load_group<R>(I, S, inst0, count, x, y, b, c);
        double ta[R], si[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            ta[r] = inst0 + r < count ? __ldg(tau + inst0 + r) : 0.0;
            si[r] = inst0 + r < count ? __ldg(sigma + inst0 + r) : 0.0;
        }
        PrimalR<R> pop{S, ta};
        DualR<R> dop{S, si};
        for (int it = 0; it < iters; ++it) {
            run_phase_r<R>(I.AT, VAT, pop, S.spart, nullptr);
            __syncthreads();
            run_phase_r<R>(I.A, VA, dop, S.spart, nullptr);
            __syncthreads();
        }
        if (scalars) {
            double s[R * 10], dd[R * 2];
            batch_kkt_r<R>(I, VA, VAT, S, s, dd);
            if (threadIdx.x == 0) {
                for (int r = 0; r < R; ++r) {
                    if (inst0 + r >= count) break;
                    double* o = scalars + (size_t)(inst0 + r) * MLLP_NUM_SCALARS;
                    for (int k = 0; k < 10; ++k) o[k] = s[r * 10 + k];
                    o[10] = (double)iters; o[11] = 0.0; o[12] = 0.0; o[13] = 1.0; o[14] = 0.0; o[15] = 0.0;
                }
            }
        }
        store_group<R>(I, S, inst0, count, x, y);
    }
}

// one slot of the interleaved vectors <-> one instance of the batch (anchors = the loaded point)
template <int R>
__device__ __forceinline__ void load_slot(const BatchInst& I, const SmemR<R>& S, int r, int inst, const double* x,
                                          const double* y, const double* b, const double* c)
{
    for (int k = threadIdx.x; k < I.n; k += blockDim.x) {
        const size_t at = (size_t)inst * I.n + __ldg(I.orderX + k);
        const double sc = I.dc ? __ldg(I.dc + k) : 1.0;
        const double xv = x[at] / sc;
        S.x[(size_t)k * R + r] = xv; S.x0[(size_t)k * R + r] = xv;
        S.c[(size_t)k * R + r] = c[at] * sc;
        S.xbar[(size_t)k * R + r] = 0.0;
    }
    for (int k = threadIdx.x; k < I.m; k += blockDim.x) {
        const size_t at = (size_t)inst * I.m + __ldg(I.orderY + k);
        const double sc = I.dr ? __ldg(I.dr + k) : 1.0;
        const double yv = y[at] / sc;
        S.y[(size_t)k * R + r] = yv; S.y0[(size_t)k * R + r] = yv;
        S.b[(size_t)k * R + r] = b[at] * sc;
    }
}
template <int R>
__device__ __forceinline__ void store_slot(const BatchInst& I, const SmemR<R>& S, int r, int inst, double* x, double* y)
{
    for (int k = threadIdx.x; k < I.n; k += blockDim.x)
        x[(size_t)inst * I.n + __ldg(I.orderX + k)] = S.x[(size_t)k * R + r] * (I.dc ? __ldg(I.dc + k) : 1.0);
    for (int k = threadIdx.x; k < I.m; k += blockDim.x)
        y[(size_t)inst * I.m + __ldg(I.orderY + k)] = S.y[(size_t)k * R + r] * (I.dr ? __ldg(I.dr + k) : 1.0);
}

// Solve mode for a shared matrix, R instances per CTA: every slot runs its own instance with its own primal
// weight, Halpern counter, restart state, iteration count and termination (the rules of k_batch_solve /
// oracle_pdhg_solve); the matrix steps are shared by the R slots.  A slot whose instance has finished stores it
// and pulls the next instance from the device counter at once, so slow instances do not hold the others back.
template <int R>
__global__ void __launch_bounds__(1024, 1)
k_batch_solve_r(const BatchInst* __restrict__ insts, int count, BatchGeom geom, double* x, double* y, const double* b,
                const double* c, const double* eta_arr, double w0, int max_iters, int check_every, double tol,
                double* scalars, int* next_inst)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    __shared__ int s_next[R];
    // step parameters of the slots, double buffered by iteration parity: [buffer][tau, sigma, lam][slot].  The two
    // fp64 divisions per slot and iteration (eta / w, (k+1)/(k+2)) are done by R threads one iteration ahead instead
    // of by every thread (they cost more than a small LP's share of the SpMV per thread).
    __shared__ double s_par[2][3][R];
    __shared__ int s_act[R];
    const BatchInst I = insts[0];
    const SmemR<R> S = carve_r<R>(dsm, I.m, I.n, true);
    MatView VA, VAT;
    batch_views(I, geom, dsm, VA, VAT);

    double eta[R], w[R], fpe_restart[R], fpe_prev[R], fpe[R];
    int k[R], it[R], restarts[R], active[R], inst[R];
    int pb = 0;   // buffer of s_par the next iteration reads
    // (re)fill the slots flagged in `want`: uniform control flow, the state is replicated in every thread
    auto fill = [&](const bool* want) {
        __syncthreads();
        if (threadIdx.x == 0)
            for (int r = 0; r < R; ++r) s_next[r] = want[r] ? atomicAdd(next_inst, 1) : -1;
        __syncthreads();
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (!want[r]) continue;
            const int nx = s_next[r];
            active[r] = nx < count;
            inst[r] = nx;
            eta[r] = 1.0; w[r] = w0 > 0.0 ? w0 : 1.0; fpe_restart[r] = -1.0; fpe_prev[r] = INFINITY; fpe[r] = 0.0;
            k[r] = 0; it[r] = 0; restarts[r] = 0;
            if (threadIdx.x == 0) s_act[r] = active[r];
            if (active[r]) {
                eta[r] = __ldg(eta_arr + nx);
                if (threadIdx.x == 0) { s_par[pb][0][r] = eta[r] / w[r]; s_par[pb][1][r] = eta[r] * w[r]; s_par[pb][2][r] = 0.5; }
                load_slot<R>(I, S, r, nx, x, y, b, c);
            } else {
                // idle slot: zero vectors keep its (masked) lanes finite
                for (int q = threadIdx.x; q < I.n; q += blockDim.x) {
                    S.x[(size_t)q * R + r] = 0.0; S.xbar[(size_t)q * R + r] = 0.0; S.c[(size_t)q * R + r] = 0.0; S.x0[(size_t)q * R + r] = 0.0;
                }
                for (int q = threadIdx.x; q < I.m; q += blockDim.x) {
                    S.y[(size_t)q * R + r] = 0.0; S.b[(size_t)q * R + r] = 0.0; S.y0[(size_t)q * R + r] = 0.0;
                }
            }
        }
        __syncthreads();
        if (!(w0 > 0.0)) {   // the PDLP default ||c~|| / ||b~|| of every slot that was just filled (uniform control flow)
            double nn[2 * R];
#pragma unroll
            for (int q = 0; q < 2 * R; ++q) nn[q] = 0.0;
            for (int q = threadIdx.x; q < I.m; q += blockDim.x)
#pragma unroll
                for (int r = 0; r < R; ++r) nn[2 * r] += S.b[(size_t)q * R + r] * S.b[(size_t)q * R + r];
            for (int q = threadIdx.x; q < I.n; q += blockDim.x)
#pragma unroll
                for (int r = 0; r < R; ++r) nn[2 * r + 1] += S.c[(size_t)q * R + r] * S.c[(size_t)q * R + r];
            cta_allreduce_r<2 * R, R>(nn, S);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (!want[r] || !active[r]) continue;
                w[r] = (nn[2 * r] > 0.0 && nn[2 * r + 1] > 0.0) ? sqrt(nn[2 * r + 1] / nn[2 * r]) : 1.0;
                if (threadIdx.x == 0) { s_par[pb][0][r] = eta[r] / w[r]; s_par[pb][1][r] = eta[r] * w[r]; }
            }
            __syncthreads();
        }
    };
    {
        bool all[R];
#pragma unroll
        for (int r = 0; r < R; ++r) { all[r] = true; active[r] = 0; }
        fill(all);
    }
    for (;;) {
        bool any = false;
#pragma unroll
        for (int r = 0; r < R; ++r) any |= (active[r] != 0);
        if (!any) break;
        const double* tau = &s_par[pb][0][0];
        const double* sigma = &s_par[pb][1][0];
        const double* lam = &s_par[pb][2][0];
        bool check[R];
        bool need = false, any_check = false;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            check[r] = active[r] && (((it[r] + 1) % check_every == 0) || (it[r] + 1 >= max_iters));
            any_check |= check[r];
            need |= check[r] || (active[r] && fpe_restart[r] < 0.0);
            if (threadIdx.x == r) {   // the next iteration's parameters if nothing restarts (rewritten below if it does)
                const int kn = k[r] + active[r];
                s_par[pb ^ 1][0][r] = tau[r]; s_par[pb ^ 1][1][r] = sigma[r];
                s_par[pb ^ 1][2][r] = (double)(kn + 1) / (double)(kn + 2);
            }
        }
        double a2[2 * R];
#pragma unroll
        for (int q = 0; q < 2 * R; ++q) a2[q] = 0.0;
        { PrimalHalpernR<R> op{S, tau, lam, s_act}; run_phase_r<R>(I.AT, VAT, op, S.spart, a2); }
        __syncthreads();
        { DualHalpernR<R> op{S, sigma, lam, s_act}; run_phase_r<R>(I.A, VA, op, S.spart, a2 + R); }
        __syncthreads();
        pb ^= 1;
#pragma unroll
        for (int r = 0; r < R; ++r) { it[r] += active[r]; k[r] += active[r]; }
        if (need) {
            cta_allreduce_r<2 * R, R>(a2, S);
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (active[r] && (check[r] || fpe_restart[r] < 0.0)) {
                    fpe[r] = sqrt(w[r] * a2[r] + a2[R + r] / w[r]);
                    if (fpe_restart[r] < 0.0) fpe_restart[r] = fpe[r];
                }
        }
        if (any_check) {
            double kk[R * 10], dd[R * 2];
            batch_kkt_r<R>(I, VA, VAT, S, kk, dd);
            bool rs[R], done[R];
            bool any_rs = false, any_done = false;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                rs[r] = false; done[r] = false;
                if (!check[r]) continue;
                const bool conv = kk[r * 10 + 8] <= tol;
                if (conv || it[r] >= max_iters) {
                    done[r] = true; any_done = true;
                    if (threadIdx.x == 0) {
                        double* o = scalars + (size_t)inst[r] * MLLP_NUM_SCALARS;
                        for (int q = 0; q < 10; ++q) o[q] = kk[r * 10 + q];
                        o[10] = (double)it[r]; o[11] = (double)restarts[r]; o[12] = conv ? 1.0 : 0.0; o[13] = w[r];
                        o[14] = fpe[r]; o[15] = 0.0;
                    }
                    continue;
                }
                const bool do_restart = (fpe[r] <= 0.2 * fpe_restart[r]) ||
                                        (fpe[r] <= 0.8 * fpe_restart[r] && fpe[r] > fpe_prev[r]) ||
                                        ((double)k[r] >= 0.36 * (double)it[r]);
                fpe_prev[r] = fpe[r];
                if (do_restart) {
                    const double ddx = sqrt(dd[r * 2]), ddy = sqrt(dd[r * 2 + 1]);
                    if (ddx > 1e-10 && ddy > 1e-10) w[r] = exp(0.5 * log(ddy / ddx) + 0.5 * log(w[r]));
                    rs[r] = true; any_rs = true;
                    k[r] = 0; fpe_restart[r] = -1.0; fpe_prev[r] = INFINITY;
                    ++restarts[r];
                    if (threadIdx.x == 0) { s_par[pb][0][r] = eta[r] / w[r]; s_par[pb][1][r] = eta[r] * w[r]; s_par[pb][2][r] = 0.5; }
                }
            }
            if (any_rs) {
                for (int q = threadIdx.x; q < I.n; q += blockDim.x)
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if (rs[r]) S.x0[(size_t)q * R + r] = S.x[(size_t)q * R + r];
                for (int q = threadIdx.x; q < I.m; q += blockDim.x)
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if (rs[r]) S.y0[(size_t)q * R + r] = S.y[(size_t)q * R + r];
                __syncthreads();
            }
            if (any_done) {
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (done[r]) store_slot<R>(I, S, r, inst[r], x, y);
                fill(done);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Small LPs in large batches: ONE WARP per instance.
//
// The CTA-per-instance kernels above spend most of their issue slots on what a CTA needs to cooperate (tile strides over
// the warps, __syncthreads between the phases, block-wide reductions through shared memory): measured on 4096 x sc105
// (340 nonzeros, a handful of tiles per phase) ~3 000 warp instructions per LP iteration, and the shared memory of a CTA
// (split-row scratch, reduction scratch) caps the LPs in flight at 12 per SM.  Here a warp owns an instance: it walks
// ALL tiles of A' and of A itself (two at a time), the phases are separated by __syncwarp(), every reduction is a
// butterfly, and shared memory holds nothing but the instance's seven vectors, so ~30 instances are in flight per SM.
// Chosen by the host (mllp_batch_create) for SOLVE mode on shared-matrix batches of >= WARP_MIN_COUNT instances without split
// rows whose vectors fit WARP_MIN_PER_SM times into an SM's shared memory (MLLP_BATCH_WARP=1 forces both kernels for any
// batch without split rows: with one matrix per instance, and in parity mode, they measured slower, see there).  Per row the summation order is that of run_phase (same tiles, same
// lanes), so parity-mode iterates are bitwise those of the CTA kernels; the sums over rows (KKT scalars, fixed-point
// error) are added in another order (lane-wise, then the butterfly), still fixed.
constexpr int WARP_LPS = 4;            // warps (instances) per CTA
constexpr int WARP_MIN_COUNT = 1024;   // smaller batches are latency-bound: the CTA kernels serve them
constexpr int WARP_MIN_PER_SM = 16;

struct WarpSmem {
    double *x, *xbar, *c, *y, *b, *x0, *y0;
};
__host__ __device__ inline size_t warp_lp_bytes(int m, int n) { return (8 * (4 * (size_t)n + 3 * (size_t)m) + 15) & ~(size_t)15; }
__device__ __forceinline__ WarpSmem carve_warp(unsigned char* base, int m, int n)
{
    WarpSmem S;
    double* p = reinterpret_cast<double*>(base);
    S.x = p; p += n; S.xbar = p; p += n; S.c = p; p += n; S.x0 = p; p += n;
    S.y = p; p += m; S.b = p; p += m; S.y0 = p;
    return S;
}
__device__ __forceinline__ DevLP warp_lp(const BatchInst& I, const WarpSmem& S)
{
    DevLP lp{};
    lp.A = I.A; lp.AT = I.AT; lp.m = I.m; lp.n = I.n;
    lp.b = S.b; lp.c = S.c; lp.x = S.x; lp.y = S.y; lp.xbar = S.xbar; lp.x0 = S.x0; lp.y0 = S.y0;
    lp.dr = I.dr; lp.dc = I.dc;
    return lp;
}

// all tiles of one matrix by one warp, two tiles in flight (the regular-tile loop of run_phase with one warp)
template <class Op>
__device__ __forceinline__ void warp_tiles(const MatView& V, const Op& op, double* acc)
{
    const double* __restrict__ vec = op.vec();
    const int lane = threadIdx.x & 31;
    for (uint32_t t = 0; t < V.ntiles; t += 2) {
        const int4 raw = __ldg(reinterpret_cast<const int4*>(V.desc + t));
        const int nsteps = raw.z & 0xffff, logL = (raw.z >> 16) & 0xff, nrows = (raw.z >> 24) & 0xff;
        const int L = 1 << logL, rr = lane >> logL;
        const bool owner = ((lane & (L - 1)) == 0) && (rr < nrows);
        const int r = raw.y + rr;
        typename Op::Pre pre{};
        if (owner) pre = op.prefetch(r);
        if (t + 1 < V.ntiles) {
            const int4 raw2 = __ldg(reinterpret_cast<const int4*>(V.desc + t + 1));
            const int nsteps2 = raw2.z & 0xffff, logL2 = (raw2.z >> 16) & 0xff, nrows2 = (raw2.z >> 24) & 0xff;
            const int L2 = 1 << logL2, rr2 = lane >> logL2;
            const bool owner2 = ((lane & (L2 - 1)) == 0) && (rr2 < nrows2);
            const int r2 = raw2.y + rr2;
            typename Op::Pre pre2{};
            if (owner2) pre2 = op.prefetch(r2);
            double dot, dot2;
            if (nsteps == nsteps2 && nsteps >= 1 && nsteps <= 3) {
                const double2* vpa = V.gvals + (size_t)(uint32_t)raw.x * 32 + lane;
                const int2* ipa = V.gidx + (size_t)(uint32_t)raw.x * 32 + lane;
                const double2* vpb = V.gvals + (size_t)(uint32_t)raw2.x * 32 + lane;
                const int2* ipb = V.gidx + (size_t)(uint32_t)raw2.x * 32 + lane;
                if (nsteps == 1) pair_dot<typename Op::Mem, 1>(vpa, ipa, vpb, ipb, vec, dot, dot2);
                else if (nsteps == 2) pair_dot<typename Op::Mem, 2>(vpa, ipa, vpb, ipb, vec, dot, dot2);
                else pair_dot<typename Op::Mem, 3>(vpa, ipa, vpb, ipb, vec, dot, dot2);
            } else {
                dot = tile_dot<typename Op::Mem>(V, vec, (uint32_t)raw.x, nsteps, lane);
                dot2 = tile_dot<typename Op::Mem>(V, vec, (uint32_t)raw2.x, nsteps2, lane);
            }
            for (int o = L >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(FULL, dot, o);
            for (int o = L2 >> 1; o > 0; o >>= 1) dot2 += __shfl_xor_sync(FULL, dot2, o);
            if (owner) op.row(r, dot, pre, acc);
            if (owner2) op.row(r2, dot2, pre2, acc);
        } else {
            double dot = tile_dot<typename Op::Mem>(V, vec, (uint32_t)raw.x, nsteps, lane);
            for (int o = L >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(FULL, dot, o);
            if (owner) op.row(r, dot, pre, acc);
        }
    }
    __syncwarp();   // the rows this phase wrote are gathered by the next one
}

// The same phase on the row-per-lane image: groups of 32 rows, lane = row, the group's slots one after the other (coalesced
// 128 / 256-byte loads of indices / values, one gather and one DFMA per entry, no descriptor decoding, no shuffles, no owner
// logic).  ncu of the tile walk on 4096 x sc105: 1 357 warp instructions per LP iteration with DFMA 6.8 % of them; this walk
// needs ~7 per slot and sc105 has 31 slots per iteration.  Per row the entries are added in CSR order (the tile walk splits a
// row over lanes), so the iterates differ from the tile walk's in rounding; the kernels below use it in SOLVE mode.
template <class Op>
__device__ __forceinline__ void warp_ell(int nrows, const int32_t* __restrict__ eidx, const double* __restrict__ eval,
                                         const uint32_t* __restrict__ eoff, const Op& op, double* acc)
{
    const double* __restrict__ vec = op.vec();
    const int lane = threadIdx.x & 31;
    const int ngroups = (nrows + 31) >> 5;
    uint32_t s0 = __ldg(eoff);
    for (int g = 0; g < ngroups; ++g) {
        const uint32_t s1 = __ldg(eoff + g + 1);
        const int r = (g << 5) + lane;
        const bool live = r < nrows;
        typename Op::Pre pre{};
        if (live) pre = op.prefetch(r);
        double dot = 0.0;
        uint32_t s = s0;
        for (; s + 1 < s1; s += 2) {   // two slots in flight
            const int32_t ja = __ldg(eidx + (size_t)s * 32 + lane), jb = __ldg(eidx + (size_t)(s + 1) * 32 + lane);
            const double va = __ldg(eval + (size_t)s * 32 + lane), vb = __ldg(eval + (size_t)(s + 1) * 32 + lane);
            dot = fma(va, Op::Mem::gather(vec + ja), dot);
            dot = fma(vb, Op::Mem::gather(vec + jb), dot);
        }
        if (s < s1) {
            const int32_t ja = __ldg(eidx + (size_t)s * 32 + lane);
            const double va = __ldg(eval + (size_t)s * 32 + lane);
            dot = fma(va, Op::Mem::gather(vec + ja), dot);
        }
        if (live) op.row(r, dot, pre, acc);
        s0 = s1;
    }
    __syncwarp();   // the rows this phase wrote are gathered by the next one
}

template <int N>
__device__ __forceinline__ void warp_allreduce(double* acc)
{
#pragma unroll
    for (int k = 0; k < N; ++k)
        for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(FULL, acc[k], o);
}

__device__ __forceinline__ void warp_load_instance(const BatchInst& I, const WarpSmem& S, const double* x, const double* y,
                                                   const double* b, const double* c)
{
    const int lane = threadIdx.x & 31;
    for (int k = lane; k < I.n; k += 32) {
        const int j = __ldg(I.orderX + k);
        const double sc = I.dc ? __ldg(I.dc + k) : 1.0;
        const double xv = x[I.x_off + j] / sc;
        S.x[k] = xv; S.x0[k] = xv;
        S.c[k] = c[I.x_off + j] * sc;
        S.xbar[k] = 0.0;
    }
    for (int k = lane; k < I.m; k += 32) {
        const int i = __ldg(I.orderY + k);
        const double sc = I.dr ? __ldg(I.dr + k) : 1.0;
        const double yv = y[I.y_off + i] / sc;
        S.y[k] = yv; S.y0[k] = yv;
        S.b[k] = b[I.y_off + i] * sc;
    }
    __syncwarp();
}
__device__ __forceinline__ void warp_store_instance(const BatchInst& I, const WarpSmem& S, double* x, double* y)
{
    const int lane = threadIdx.x & 31;
    for (int k = lane; k < I.n; k += 32) x[I.x_off + __ldg(I.orderX + k)] = S.x[k] * (I.dc ? __ldg(I.dc + k) : 1.0);
    for (int k = lane; k < I.m; k += 32) y[I.y_off + __ldg(I.orderY + k)] = S.y[k] * (I.dr ? __ldg(I.dr + k) : 1.0);
    __syncwarp();
}

// KKT scalars of (S.x, S.y) -> s[0..9] in every lane; also ||x-x0||^2, ||y-y0||^2 in dd[0..1]
// (`E` = the instance when its row-per-lane images are to be walked, else null: the tile walk)
__device__ __forceinline__ void warp_kkt(const DevLP& lp, const MatView& VA, const MatView& VAT, double* s, double* dd,
                                         const BatchInst* E = nullptr)
{
    double ap[NRED], ad[NRED];
#pragma unroll
    for (int k = 0; k < NRED; ++k) { ap[k] = 0.0; ad[k] = 0.0; }
    if (E) {
        { EvalPrimalOp<false, SmemMem> op{lp}; warp_ell(lp.n, E->eidxT, E->evalT, E->eoffT, op, ap); }
        { EvalDualOp<false, SmemMem> op{lp}; warp_ell(lp.m, E->eidxA, E->evalA, E->eoffA, op, ad); }
    } else {
        { EvalPrimalOp<false, SmemMem> op{lp}; warp_tiles(VAT, op, ap); }
        { EvalDualOp<false, SmemMem> op{lp}; warp_tiles(VA, op, ad); }
    }
    warp_allreduce<7>(ap);
    warp_allreduce<6>(ad);
    const double pobj = ap[0], dobj = ad[0] + ap[1];
    s[0] = pobj; s[1] = dobj; s[2] = sqrt(ad[1] + ap[6]); s[3] = sqrt(ap[2] + ad[5]);
    s[4] = sqrt(ad[2]); s[5] = sqrt(ap[3]); s[6] = sqrt(ap[4]); s[7] = sqrt(ad[3]);
    const double gap = fabs(pobj - dobj);
    double e = s[2] / (1.0 + s[4]);
    e = fmax(e, s[3] / (1.0 + s[5]));
    e = fmax(e, gap / (1.0 + fabs(pobj) + fabs(dobj)));
    s[8] = e; s[9] = gap;
    dd[0] = ap[5]; dd[1] = ad[4];
}

__global__ void __launch_bounds__(32 * WARP_LPS, 6)
k_batch_run_warp(const BatchInst* __restrict__ insts, int count, int shared, uint32_t lp_bytes, double* x, double* y, const double* b,
                 const double* c, const double* tau, const double* sigma, int iters, double* scalars)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned char* mine = dsm + (size_t)wid * lp_bytes;
    const int nw = gridDim.x * WARP_LPS;
    for (int inst = blockIdx.x * WARP_LPS + wid; inst < count; inst += nw) {
        BatchInst I = insts[shared ? 0 : inst];
        if (shared) { I.x_off = (long long)inst * I.n; I.y_off = (long long)inst * I.m; }
        const WarpSmem S = carve_warp(mine, I.m, I.n);
        warp_load_instance(I, S, x, y, b, c);
        for (int k = lane; k < I.n; k += 32) S.x0[k] = 0.0;   // anchors unused here; keep the evaluation finite
        for (int k = lane; k < I.m; k += 32) S.y0[k] = 0.0;
        __syncwarp();
        const DevLP lp = warp_lp(I, S);
        const MatView VA = global_view(I.A, 0u), VAT = global_view(I.AT, 0u);
        PrimalOp<false, SmemMem> pop{lp, __ldg(tau + inst)};
        DualOp<false, SmemMem> dop{lp, __ldg(sigma + inst)};
        double acc[NRED];
        for (int it = 0; it < iters; ++it) {
            warp_tiles(VAT, pop, acc);
            warp_tiles(VA, dop, acc);
        }
        if (scalars) {
            double s[10], dd[2];
            warp_kkt(lp, VA, VAT, s, dd);
            if (lane == 0) {
                double* o = scalars + (size_t)inst * MLLP_NUM_SCALARS;
                for (int k = 0; k < 10; ++k) o[k] = s[k];
                o[10] = (double)iters; o[11] = 0.0; o[12] = 0.0; o[13] = 1.0; o[14] = 0.0; o[15] = 0.0;
            }
        }
        warp_store_instance(I, S, x, y);
    }
}

// Solve mode per warp: the control flow of k_batch_solve / oracle_pdhg_solve; all lanes hold the same scalars.
__global__ void __launch_bounds__(32 * WARP_LPS, 6)
k_batch_solve_warp(const BatchInst* __restrict__ insts, int count, int shared, uint32_t lp_bytes, double* x, double* y,
                   const double* b, const double* c, const double* eta_arr, double w0, int max_iters, int check_every, double tol,
                   double* scalars, int* next_inst)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned char* mine = dsm + (size_t)wid * lp_bytes;
    for (;;) {
        // instances converge after very different iteration counts: warps pull the next one from a device counter
        int inst = 0;
        if (lane == 0) inst = atomicAdd(next_inst, 1);
        inst = __shfl_sync(FULL, inst, 0);
        if (inst >= count) break;
        BatchInst I = insts[shared ? 0 : inst];
        if (shared) { I.x_off = (long long)inst * I.n; I.y_off = (long long)inst * I.m; }
        const WarpSmem S = carve_warp(mine, I.m, I.n);
        warp_load_instance(I, S, x, y, b, c);
        const double eta = __ldg(eta_arr + inst);
        double w = w0;
        if (!(w0 > 0.0)) {   // the PDLP default ||c~|| / ||b~|| of this instance (scaled data, as loaded)
            double nn[2] = {0.0, 0.0};
            for (int q = lane; q < I.m; q += 32) nn[0] += S.b[q] * S.b[q];
            for (int q = lane; q < I.n; q += 32) nn[1] += S.c[q] * S.c[q];
            warp_allreduce<2>(nn);
            w = (nn[0] > 0.0 && nn[1] > 0.0) ? sqrt(nn[1] / nn[0]) : 1.0;
        }
        const DevLP lp = warp_lp(I, S);
        const MatView VA = global_view(I.A, 0u), VAT = global_view(I.AT, 0u);
        const BatchInst* E = (I.eoffA && I.eoffT) ? &I : nullptr;   // row-per-lane images present: walk those
        double tau = eta / w, sigma = eta * w;
        double fpe_restart = -1.0, fpe_prev = INFINITY, fpe = 0.0;
        int k = 0, it = 0, restarts = 0, converged = 0;
        double kk[10], dd[2];
        warp_kkt(lp, VA, VAT, kk, dd, E);
        while (it < max_iters) {
            const double lam = (double)(k + 1) / (double)(k + 2);
            const bool check = ((it + 1) % check_every == 0) || (it + 1 == max_iters);
            const bool need_fpe = check || fpe_restart < 0.0;
            double a2[2];
            {
                PrimalHalpernOp<false, SmemMem> op{lp, tau, lam};
                double acc[NRED];
                acc[0] = 0.0;
                if (E) warp_ell(I.n, I.eidxT, I.evalT, I.eoffT, op, acc);
                else warp_tiles(VAT, op, acc);
                a2[0] = acc[0];
            }
            {
                DualHalpernOp<false, SmemMem> op{lp, sigma, lam};
                double acc[NRED];
                acc[0] = 0.0;
                if (E) warp_ell(I.m, I.eidxA, I.evalA, I.eoffA, op, acc);
                else warp_tiles(VA, op, acc);
                a2[1] = acc[0];
            }
            ++it; ++k;
            if (need_fpe) {
                warp_allreduce<2>(a2);
                fpe = sqrt(w * a2[0] + a2[1] / w);
                if (fpe_restart < 0.0) fpe_restart = fpe;
            }
            if (check) {
                warp_kkt(lp, VA, VAT, kk, dd, E);
                if (kk[8] <= tol) { converged = 1; break; }
                const bool do_restart = (fpe <= 0.2 * fpe_restart) || (fpe <= 0.8 * fpe_restart && fpe > fpe_prev) ||
                                        ((double)k >= 0.36 * (double)it);
                fpe_prev = fpe;
                if (do_restart) {
                    const double ddx = sqrt(dd[0]), ddy = sqrt(dd[1]);
                    if (ddx > 1e-10 && ddy > 1e-10) w = exp(0.5 * log(ddy / ddx) + 0.5 * log(w));
                    tau = eta / w; sigma = eta * w;
                    for (int q = lane; q < I.n; q += 32) S.x0[q] = S.x[q];
                    for (int q = lane; q < I.m; q += 32) S.y0[q] = S.y[q];
                    __syncwarp();
                    k = 0; fpe_restart = -1.0; fpe_prev = INFINITY;
                    ++restarts;
                }
            }
        }
        if (lane == 0) {
            double* o = scalars + (size_t)inst * MLLP_NUM_SCALARS;
            for (int q = 0; q < 10; ++q) o[q] = kk[q];
            o[10] = (double)it; o[11] = (double)restarts; o[12] = (double)converged; o[13] = w; o[14] = fpe; o[15] = 0.0;
        }
        warp_store_instance(I, S, x, y);
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
void count_launch(int n);                       // cabi.cu: launch statistics (mllp_launch_count)
int precondition_device(int m, int n, long long nnz, const int* h_ptr, const int* h_ind, double* h_values, const int* h_tptr,
                        const int* h_tind, const double* h_tval, int ruiz_iters, double* h_dr, double* h_dc);   // scaling.cu
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
    // shared matrix: instances per CTA (multi-RHS) and launch geometry of the parity / solve kernels
    int R_run = 1, R_solve = 1;
    int grid_run = 0, grid_solve = 0;
    size_t smem_run = 0, smem_solve = 0;
    BatchGeom g_run{}, g_solve{};
    // one warp per instance (k_batch_run_warp / k_batch_solve_warp): large batches of small LPs without split rows
    int use_warp = 0, use_warp_run = 0, grid_warp = 0;
    uint32_t warp_lp_bytes = 0;
    size_t smem_warp = 0;
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
    std::vector<double> scale;       // preconditioned batch: dr | dc of every matrix, internal order
    std::vector<int32_t> ell_idx;    // row-per-lane images (HostEll) of A | A' of every small matrix
    std::vector<double> ell_val;
    std::vector<uint32_t> ell_off;
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
    const bool precondition = (flags & MLLP_F_PRECONDITION) != 0;
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
    // Lanes per row: a batch that fills the GPU several times over is bound by INSTRUCTION ISSUE (ncu, 4096 x 25fv47:
    // issue slots 46 % busy, DFMA 6 % of the instructions, the rest is tile decoding, shuffle trees and row updates), so rows
    // get few lanes and up to 6 steps -- fewer, longer tiles: 152 -> 105 us per batch iteration (measured 2 / 3 / 4 / 6 / 8
    // steps: 152 / 129 / 126 / 105 / 109 us).  A small batch is latency bound: 2 steps, wide rows.
    bp.pref_steps = count >= 4 * prop.multiProcessorCount ? 6 : 2;
    bp.max_steps = 4;
    const char* ev = getenv("MLLP_BATCH_PREF_STEPS");
    if (ev && *ev) bp.pref_steps = std::max(1, atoi(ev));
    if (bp.max_steps < bp.pref_steps) bp.max_steps = bp.pref_steps;

    int rc = 0;
    try {
        Pools P;
        std::vector<MatOff> offA((size_t)nmat), offAT((size_t)nmat);
        std::vector<size_t> offOX((size_t)nmat), offOY((size_t)nmat), offDR((size_t)nmat), offDC((size_t)nmat);
        // row-per-lane images: offsets of the entries (idx / val) and of the group table, A and A'; NO_ELL: not built
        constexpr size_t NO_ELL = ~(size_t)0;
        std::vector<size_t> offEAe((size_t)nmat, NO_ELL), offEAo((size_t)nmat, NO_ELL), offETe((size_t)nmat, NO_ELL), offETo((size_t)nmat, NO_ELL);
        const size_t warp_cap = (size_t)prop.sharedMemPerBlockOptin - 4096;
        int max_tiles = 0, max_tiles_A = 0, max_tiles_AT = 0, max_steps_A = 0, max_steps_AT = 0;
        bool any_split = false;
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
            // MLLP_F_PRECONDITION: Ruiz + Pock-Chambolle of every distinct matrix on the device; the kernels scale the caller's
            // (original) vectors when they load / store an instance and evaluate the KKT scalars on the original LP
            std::vector<double> scaled, h_dr, h_dc;
            if (precondition) {
                scaled.assign(vv, vv + nnz);
                h_dr.assign((size_t)m, 1.0); h_dc.assign((size_t)n, 1.0);
                const char* rz = getenv("MLLP_RUIZ_ITERS");
                const int prc = precondition_device(m, n, nnz, ip, ii, scaled.data(), tptr.data(), tind.data(), tval.data(),
                                                    rz && *rz ? atoi(rz) : 10, h_dr.data(), h_dc.data());
                if (prc != 0) { rc = prc; break; }
                vv = scaled.data();
                csr_transpose(m, n, ip, ii, vv, tptr, tind, tval);
            }
            std::vector<int32_t> orderY, posY, orderX, posX;
            plan_orders(m, n, ip, ii, tptr.data(), tind.data(), bp, orderY, posY, orderX, posX);
            if (precondition) {
                offDR[k] = P.scale.size();
                for (int q = 0; q < m; ++q) P.scale.push_back(h_dr[orderY[q]]);
                offDC[k] = P.scale.size();
                for (int q = 0; q < n; ++q) P.scale.push_back(h_dc[orderX[q]]);
            }
            HostMat HA, HAT;
            build_host_mat(m, n, ip, ii, vv, orderY, posX, bp, HA);
            build_host_mat(n, m, tptr.data(), tind.data(), tval.data(), orderX, posY, bp, HAT);
            offA[k] = append(P, HA);
            offAT[k] = append(P, HAT);
            if (warp_lp_bytes(m, n) * WARP_LPS <= warp_cap) {   // small enough for the warp-per-instance kernels
                HostEll EA, ET;
                build_host_ell(m, ip, ii, vv, orderY, posX, EA);
                build_host_ell(n, tptr.data(), tind.data(), tval.data(), orderX, posY, ET);
                offEAe[k] = P.ell_idx.size(); offEAo[k] = P.ell_off.size();
                P.ell_idx.insert(P.ell_idx.end(), EA.idx.begin(), EA.idx.end());
                P.ell_val.insert(P.ell_val.end(), EA.val.begin(), EA.val.end());
                P.ell_off.insert(P.ell_off.end(), EA.off.begin(), EA.off.end());
                offETe[k] = P.ell_idx.size(); offETo[k] = P.ell_off.size();
                P.ell_idx.insert(P.ell_idx.end(), ET.idx.begin(), ET.idx.end());
                P.ell_val.insert(P.ell_val.end(), ET.val.begin(), ET.val.end());
                P.ell_off.insert(P.ell_off.end(), ET.off.begin(), ET.off.end());
            }
            offOX[k] = P.order.size(); P.order.insert(P.order.end(), orderX.begin(), orderX.end());
            offOY[k] = P.order.size(); P.order.insert(P.order.end(), orderY.begin(), orderY.end());
            max_tiles = std::max<int>(max_tiles, (int)std::max(HA.tiles.size(), HAT.tiles.size()));
            any_split = any_split || !HA.splits.empty() || !HAT.splits.empty() || HA.cta_nsplit[0] != 0 || HAT.cta_nsplit[0] != 0;
            max_tiles_A = std::max<int>(max_tiles_A, (int)HA.tiles.size());
            max_tiles_AT = std::max<int>(max_tiles_AT, (int)HAT.tiles.size());
            max_steps_A = std::max<int>(max_steps_A, (int)HA.total_steps);
            max_steps_AT = std::max<int>(max_steps_AT, (int)HAT.total_steps);
            bt->max_m = std::max(bt->max_m, m); bt->max_n = std::max(bt->max_n, n);
            bt->sum_nnz += nnz;
        }
        if (rc == 0) {
            for (int k = 0; k < count; ++k) {
                bt->sum_m += h_m[bt->shared ? 0 : k];
                bt->sum_n += h_n[bt->shared ? 0 : k];
            }
            double* d_vals; int32_t* d_idx; Tile* d_tiles; uint32_t* d_u32; SplitRow* d_splits; LocalSplit* d_ls; int32_t* d_order;
            double* d_dummy_partials; unsigned* d_dummy_counters; double* d_scale = nullptr;
            auto ck = [&](cudaError_t ce, const char* what) { if (ce != cudaSuccess && rc == 0) rc = bfail((int)ce, std::string(what) + ": " + cudaGetErrorString(ce)); };
            ck(up(bt, &d_vals, P.vals), "upload vals");
            ck(up(bt, &d_idx, P.idx), "upload idx");
            ck(up(bt, &d_tiles, P.tiles), "upload tiles");
            ck(up(bt, &d_u32, P.u32), "upload tables");
            ck(up(bt, &d_splits, P.splits), "upload splits");
            ck(up(bt, &d_ls, P.lsplits), "upload local splits");
            ck(up(bt, &d_order, P.order), "upload orders");
            if (precondition) ck(up(bt, &d_scale, P.scale), "upload scaling vectors");
            int32_t* d_eidx = nullptr; double* d_eval = nullptr; uint32_t* d_eoff = nullptr;
            if (!P.ell_off.empty()) {
                ck(up(bt, &d_eidx, P.ell_idx), "upload row-per-lane indices");
                ck(up(bt, &d_eval, P.ell_val), "upload row-per-lane values");
                ck(up(bt, &d_eoff, P.ell_off), "upload row-per-lane group table");
            }
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
                    I.dr = precondition ? d_scale + offDR[k] : nullptr;
                    I.dc = precondition ? d_scale + offDC[k] : nullptr;
                    I.m = h_m[k]; I.n = h_n[k];
                    I.x_off = xo; I.y_off = yo;
                    const bool ell = offEAe[k] != NO_ELL && d_eoff;
                    I.eidxA = ell ? d_eidx + offEAe[k] : nullptr; I.evalA = ell ? d_eval + offEAe[k] : nullptr;
                    I.eoffA = ell ? d_eoff + offEAo[k] : nullptr;
                    I.eidxT = ell ? d_eidx + offETe[k] : nullptr; I.evalT = ell ? d_eval + offETe[k] : nullptr;
                    I.eoffT = ell ? d_eoff + offETo[k] : nullptr;
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

            // Shared matrix: R instances per CTA with interleaved vectors (multi-RHS), as many as fit in shared memory
            // while every SM still gets a group.  What is left of the shared memory keeps the tile descriptors and a
            // prefix of the matrix steps resident (A' first); also used for one big matrix per instance.
            auto bytes_r = [&](int R, bool anchors) {
                const size_t per = anchors ? 4 * (size_t)bt->max_n + 3 * (size_t)bt->max_m : 3 * (size_t)bt->max_n + 2 * (size_t)bt->max_m;
                return (8 * ((size_t)32 * NRED + 16 * (size_t)R + (size_t)SPLIT_SLOTS * R + per * R) + 15) & ~(size_t)15;
            };
            const size_t desc_bytes = 16 * ((size_t)max_tiles_A + (size_t)max_tiles_AT);
            const size_t mat_bytes = desc_bytes + 768 * ((size_t)max_steps_A + (size_t)max_steps_AT);
            auto pick_R = [&](bool anchors) {
                int R = 1;
                if (bt->shared) {
                    // measured (scripts/batch_bench.py): R = 2 brings 1.5x (25fv47) to 2.1x (sc105) in LP-iterations/s,
                    // R = 3, 4 add nothing (the interleaved gathers are bound by shared-memory bank conflicts)
                    for (int cand : {2})
                        if (bytes_r(cand, anchors) + desc_bytes <= smem_cap && (int64_t)count >= (int64_t)cand * prop.multiProcessorCount) R = cand;
                    const char* rv = getenv("MLLP_BATCH_R");
                    if (rv && *rv) {
                        const int want = atoi(rv);
                        if (want >= 1 && want <= 4 && (want == 1 || bytes_r(want, anchors) + desc_bytes <= smem_cap)) R = want;
                    }
                }
                return R;
            };
            auto geometry = [&](int R, bool anchors, size_t& smem, BatchGeom& g) {
                const size_t vec = R == 1 ? bt->dyn_smem : bytes_r(R, anchors);
                smem = vec;
                g = BatchGeom{0u, 0u, 0u, 0};
                const char* rv = getenv("MLLP_BATCH_RES");
                // default: resident only if the WHOLE matrix fits behind the vectors (a partly resident matrix leaves too
                // little L1 for the streamed rest: 25fv47, R = 1: -10 %)
                const bool forced = rv && *rv && atoi(rv) != 0;
                const bool want_res = rv && *rv ? forced : (bt->shared && vec + mat_bytes <= smem_cap);
                if (want_res && vec + desc_bytes <= smem_cap) {
                    size_t steps = (smem_cap - vec - desc_bytes) / 768;
                    g.use_res = 1;
                    g.mat_off = (uint32_t)vec;
                    g.res_AT = (uint32_t)std::min<size_t>(steps, (size_t)max_steps_AT);
                    steps -= g.res_AT;
                    g.res_A = (uint32_t)std::min<size_t>(steps, (size_t)max_steps_A);
                    smem = vec + desc_bytes + 768 * ((size_t)g.res_A + g.res_AT);
                }
            };
            if (rc == 0) {
                bt->R_run = pick_R(false);
                bt->R_solve = pick_R(true);
                geometry(bt->R_run, false, bt->smem_run, bt->g_run);
                geometry(bt->R_solve, true, bt->smem_solve, bt->g_solve);
                auto run_fn = [](int R) -> const void* {
                    return R == 4 ? (const void*)k_batch_run_r<4> : R == 3 ? (const void*)k_batch_run_r<3> : R == 2 ? (const void*)k_batch_run_r<2> : (const void*)k_batch_run;
                };
                auto solve_fn = [](int R) -> const void* {
                    return R == 4 ? (const void*)k_batch_solve_r<4> : R == 3 ? (const void*)k_batch_solve_r<3> : R == 2 ? (const void*)k_batch_solve_r<2> : (const void*)k_batch_solve;
                };
                auto prepare = [&](const void* fn, size_t smem, int64_t units, int& grid) {
                    if (smem > 40 * 1024)
                        ck(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute");
                    int b = 0;
                    const bool warp_fn = fn == (const void*)k_batch_run_warp || fn == (const void*)k_batch_solve_warp;
                    ck(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, fn, warp_fn ? 32 * WARP_LPS : threads, smem), "occupancy");
                    if (rc == 0 && b < 1) rc = bfail(MLLP_E_STATE, "mllp_batch_create: kernel does not fit on an SM");
                    grid = (int)std::min<int64_t>(units, (int64_t)prop.multiProcessorCount * std::max(b, 1));
                };
                prepare(run_fn(bt->R_run), bt->smem_run, (count + bt->R_run - 1) / bt->R_run, bt->grid_run);
                prepare(solve_fn(bt->R_solve), bt->smem_solve, (count + bt->R_solve - 1) / bt->R_solve, bt->grid_solve);
                if (bt->R_solve > 1 && bt->smem_solve > 40 * 1024)   // max_iters == 0 runs the one-instance kernel with this geometry
                    ck(cudaFuncSetAttribute((const void*)k_batch_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bt->smem_solve), "cudaFuncSetAttribute");
                int gn = 0;
                prepare((const void*)k_batch_norm, bt->dyn_smem, count, gn);
                bt->grid = gn;
                // one warp per instance: large batches of small LPs (see k_batch_solve_warp)
                const size_t lpb = warp_lp_bytes(bt->max_m, bt->max_n);
                const size_t sm_smem = (size_t)prop.sharedMemPerMultiprocessor;
                // measured (scripts/small_batch_bench.py, profiles/r02_batch_experiments.md): 4096 x sc105 sharing one matrix
                // +14 % LPs/s; 3072 LPs with their OWN matrices 0.4x -- a lone warp needs 11.7 us per iteration against the
                // CTA's 1.6, and the few instances that run to the iteration cap then set the time: shared matrices only
                bool warp = bt->shared && !any_split && count >= WARP_MIN_COUNT && lpb * WARP_MIN_PER_SM + 1024 * (WARP_MIN_PER_SM / WARP_LPS) <= sm_smem &&
                            lpb * WARP_LPS <= smem_cap;
                // ... and in SOLVE mode only: at a fixed iteration count (scripts/batch_bench.py sc105 4096) the solve loop runs at
                // 4.3e8 LP-iterations/s on warps against 3.7e8 on CTAs (R = 2), the parity loop at 5.0e8 against 7.1e8 (there
                // the R = 2 kernel with its resident matrix has nothing else to do): mllp_batch_run keeps the CTA kernels
                bool warp_run = false;
                const char* wv = getenv("MLLP_BATCH_WARP");
                if (wv && *wv) warp = warp_run = atoi(wv) != 0 && !any_split && lpb * WARP_LPS <= smem_cap;
                if (warp) {
                    bt->use_warp = 1;
                    bt->use_warp_run = warp_run ? 1 : 0;
                    bt->warp_lp_bytes = (uint32_t)lpb;
                    bt->smem_warp = lpb * WARP_LPS;
                    int gw = 0, gw2 = 0;
                    prepare((const void*)k_batch_run_warp, bt->smem_warp, (count + WARP_LPS - 1) / WARP_LPS, gw);
                    prepare((const void*)k_batch_solve_warp, bt->smem_warp, (count + WARP_LPS - 1) / WARP_LPS, gw2);
                    bt->grid_warp = std::min(gw, gw2);
                }
            }
            int64_t* I = bt->info;
            I[0] = count; I[1] = bt->sum_m; I[2] = bt->sum_n; I[3] = bt->sum_nnz; I[4] = bt->grid_run; I[5] = bt->threads;
            I[6] = (int64_t)bt->smem_run;
            // algorithmic bytes per batch iteration: per instance 36 m + 44 n, plus the matrix (24 nnz) once per
            // instance (separate matrices) or once per batch (shared matrix)
            I[7] = 36 * bt->sum_m + 44 * bt->sum_n + 24 * bt->sum_nnz;
            I[8] = bt->R_run; I[9] = bt->R_solve; I[10] = bt->g_run.res_A; I[11] = bt->g_run.res_AT;
            I[12] = bt->g_solve.res_A; I[13] = bt->g_solve.res_AT; I[14] = (int64_t)bt->smem_solve; I[15] = bt->grid_solve;
            if (bt->use_warp) {   // solve mode with one warp per instance: WARP_LPS instances per CTA, nothing resident
                I[9] = WARP_LPS; I[12] = I[13] = 0; I[14] = (int64_t)bt->smem_warp; I[15] = bt->grid_warp;
            }
            if (bt->use_warp_run) {
                I[4] = bt->grid_warp; I[5] = 32 * WARP_LPS; I[6] = (int64_t)bt->smem_warp; I[8] = WARP_LPS; I[10] = I[11] = 0;
            }
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
    mllp::count_launch(1);
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
    cudaStream_t st = (cudaStream_t)stream;
    mllp::count_launch(1);
    if (bt->use_warp_run) {
        k_batch_run_warp<<<bt->grid_warp, 32 * WARP_LPS, bt->smem_warp, st>>>(bt->d_insts, bt->count, bt->shared, bt->warp_lp_bytes, d_x, d_y,
                                                                            d_b, d_c, d_tau, d_sigma, num_iters, d_scalars);
        cudaError_t ew = cudaGetLastError();
        if (ew != cudaSuccess) return bfail((int)ew, std::string("mllp_batch_run: ") + cudaGetErrorString(ew));
        return 0;
    }
#define MLLP_RUN_R(RR) k_batch_run_r<RR><<<bt->grid_run, bt->threads, bt->smem_run, st>>>(bt->d_insts, bt->count, bt->g_run, d_x, d_y, d_b, d_c, d_tau, d_sigma, num_iters, d_scalars)
    switch (bt->R_run) {
        case 4: MLLP_RUN_R(4); break;
        case 3: MLLP_RUN_R(3); break;
        case 2: MLLP_RUN_R(2); break;
        default:
            k_batch_run<<<bt->grid_run, bt->threads, bt->smem_run, st>>>(bt->d_insts, bt->count, bt->shared, bt->g_run, d_x, d_y, d_b,
                                                                       d_c, d_tau, d_sigma, num_iters, d_scalars);
    }
#undef MLLP_RUN_R
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return bfail((int)e, std::string("mllp_batch_run: ") + cudaGetErrorString(e));
    return 0;
}

int mllp_batch_solve(mllp_batch_t bt, double* d_x, double* d_y, const double* d_b, const double* d_c, const double* d_eta,
                     double w0, int32_t max_iters, int32_t check_every, double tol, double* d_scalars, void* stream)
{
    if (!bt || !d_x || !d_y || !d_b || !d_c || !d_eta || !d_scalars || max_iters < 0 || check_every < 1 || !(w0 >= 0.0))
        return bfail(MLLP_E_INVALID, "mllp_batch_solve: bad argument");
    DevGuard guard(bt->device);
    cudaError_t e0 = cudaMemsetAsync(bt->d_next, 0, sizeof(int), (cudaStream_t)stream);
    if (e0 != cudaSuccess) return bfail((int)e0, std::string("mllp_batch_solve: ") + cudaGetErrorString(e0));
    cudaStream_t st = (cudaStream_t)stream;
    mllp::count_launch(1);
    if (bt->use_warp) {
        k_batch_solve_warp<<<bt->grid_warp, 32 * WARP_LPS, bt->smem_warp, st>>>(bt->d_insts, bt->count, bt->shared, bt->warp_lp_bytes, d_x,
                                                                              d_y, d_b, d_c, d_eta, w0, max_iters, check_every, tol,
                                                                              d_scalars, bt->d_next);
        cudaError_t ew = cudaGetLastError();
        if (ew != cudaSuccess) return bfail((int)ew, std::string("mllp_batch_solve: ") + cudaGetErrorString(ew));
        return 0;
    }
#define MLLP_SOLVE_R(RR) k_batch_solve_r<RR><<<bt->grid_solve, bt->threads, bt->smem_solve, st>>>(bt->d_insts, bt->count, bt->g_solve, d_x, d_y, d_b, d_c, d_eta, w0, max_iters, check_every, tol, d_scalars, bt->d_next)
    switch (max_iters > 0 ? bt->R_solve : 1) {   // max_iters == 0 (scalars of the starting point): one-instance kernel
        case 4: MLLP_SOLVE_R(4); break;
        case 3: MLLP_SOLVE_R(3); break;
        case 2: MLLP_SOLVE_R(2); break;
        default:
            k_batch_solve<<<bt->grid_solve, bt->threads, bt->smem_solve, st>>>(bt->d_insts, bt->count, bt->shared, bt->g_solve, d_x, d_y,
                                                                             d_b, d_c, d_eta, w0, max_iters, check_every, tol,
                                                                             d_scalars, bt->d_next);
    }
#undef MLLP_SOLVE_R
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return bfail((int)e, std::string("mllp_batch_solve: ") + cudaGetErrorString(e));
    return 0;
}

}  // extern "C"
