// pdhg_kernels.cuh -- device side of the fused primal-dual LP iteration (sm_100a).
//
// One generic warp-tile SpMV walker (run_phase) drives every kernel; what happens to a
// finished row dot product is decided by an Op:
//   PrimalOp : g = c - A'y ; x+ = clip(x - tau g) ; xbar = 2x+ - x          
//   DualOp   : y+ = clip(y + sigma (b - A xbar))                            (row a7)
//   Halpern variants of both (solve mode, row a9), Eval ops (KKT scalars, row a8), SpmvOp.
// The reference has no counterpart of these; the spec is
// oracle/pdhg_oracle.c.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "lp_format.h"

namespace mllp {

constexpr int NRED = 8;            // reduction slots per phase
constexpr unsigned FULL = 0xffffffffu;
// kinds of per-CTA reduction buffers
constexpr int RED_STEPP = 0, RED_STEPD = 1, RED_EVALP = 2, RED_EVALD = 3, RED_BUFFERS = 8;
// control block slots (device doubles)
constexpr int CTRL_TAU = 0, CTRL_SIGMA = 1, CTRL_SIZE = 16;

struct DevMat {
    const double2* vals;           // [total_steps][32]
    const int2* idx;               // [total_steps][32]
    const Tile* tiles;
    const uint32_t* cta_begin;     // [G+1]
    const uint32_t* cta_step_begin;// [G+1]
    const SplitRow* splits;
    double* partials;
    unsigned* counters;
    int nrows, ncols;
};

// Everything below is in INTERNAL row/column order.
struct DevLP {
    DevMat A, AT;
    int m, n;
    const double* b;
    const double* c;
    const double* lb;   // null => 0
    const double* ub;   // null => +inf
    const double* ylo;  // null => -inf
    const double* yhi;  // null => +inf
    double* x;
    double* y;
    double* xbar;
    double* x0;         // Halpern anchors
    double* y0;
    double* red;        // [2 parity][4 kinds][G][NRED] per-CTA partial sums (RED_* below)
    unsigned* barrier;  // grid barrier counter
    double* ctrl;       // solve-mode control block (device): see SolveCtrl
};

__device__ __forceinline__ double ld_mut(const double* p) { return __ldcg(p); }   // L2-coherent
__device__ __forceinline__ double ld_ro(const double* p) { return __ldg(p); }     // read-only path

// ---------------------------------------------------------------------------------------
// Grid barrier for the persistent cooperative kernel: monotonic counter, one arrival per
// CTA (release), acquire-poll.  `target` lives in thread 0's register.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while ((int)(v - target) < 0);
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------
// Ops.  Interface:
//   const double* vec()                     gather vector (mutable => read through L2)
//   Pre  prefetch(int r)                    issue the row's own vector loads early
//   void row(int r, double dot, Pre)        consume a finished dot product
//   void row_late(int r, double dot)        same, without prefetch (split rows)
// Each op may accumulate into acc[NRED] (thread-local), reduced per CTA at the phase end.

template <bool BOUNDS>
struct PrimalOp {
    const DevLP& lp;
    double tau;
    struct Pre { double c, x; };
    __device__ __forceinline__ const double* vec() const { return lp.y; }
    __device__ __forceinline__ Pre prefetch(int r) const { return {ld_ro(lp.c + r), ld_mut(lp.x + r)}; }
    __device__ __forceinline__ void row(int r, double dot, const Pre& p, double*) const
    {
        const double g = p.c - dot;
        double xn = p.x - tau * g;
        if (BOUNDS) {
            xn = fmin(fmax(xn, ld_ro(lp.lb + r)), ld_ro(lp.ub + r));
        } else {
            xn = fmax(xn, 0.0);
        }
        lp.xbar[r] = 2.0 * xn - p.x;
        lp.x[r] = xn;
    }
};

template <bool BOUNDS>
struct DualOp {
    const DevLP& lp;
    double sigma;
    struct Pre { double b, y; };
    __device__ __forceinline__ const double* vec() const { return lp.xbar; }
    __device__ __forceinline__ Pre prefetch(int r) const { return {ld_ro(lp.b + r), ld_mut(lp.y + r)}; }
    __device__ __forceinline__ void row(int r, double dot, const Pre& p, double*) const
    {
        double yn = p.y + sigma * (p.b - dot);
        if (BOUNDS) yn = fmin(fmax(yn, ld_ro(lp.ylo + r)), ld_ro(lp.yhi + r));
        lp.y[r] = yn;
    }
};

// Solve mode (reflected Halpern).  acc[0] accumulates ||x' - x||^2 (resp. y).
template <bool BOUNDS>
struct PrimalHalpernOp {
    const DevLP& lp;
    double tau, lam;
    struct Pre { double c, x, x0; };
    __device__ __forceinline__ const double* vec() const { return lp.y; }
    __device__ __forceinline__ Pre prefetch(int r) const
    {
        return {ld_ro(lp.c + r), ld_mut(lp.x + r), ld_mut(lp.x0 + r)};
    }
    __device__ __forceinline__ void row(int r, double dot, const Pre& p, double* acc) const
    {
        const double g = p.c - dot;
        double xn = p.x - tau * g;
        if (BOUNDS) {
            xn = fmin(fmax(xn, ld_ro(lp.lb + r)), ld_ro(lp.ub + r));
        } else {
            xn = fmax(xn, 0.0);
        }
        const double d = xn - p.x;
        acc[0] += d * d;
        const double xb = 2.0 * xn - p.x;
        lp.xbar[r] = xb;
        lp.x[r] = lam * xb + (1.0 - lam) * p.x0;
    }
};

template <bool BOUNDS>
struct DualHalpernOp {
    const DevLP& lp;
    double sigma, lam;
    struct Pre { double b, y, y0; };
    __device__ __forceinline__ const double* vec() const { return lp.xbar; }
    __device__ __forceinline__ Pre prefetch(int r) const
    {
        return {ld_ro(lp.b + r), ld_mut(lp.y + r), ld_mut(lp.y0 + r)};
    }
    __device__ __forceinline__ void row(int r, double dot, const Pre& p, double* acc) const
    {
        double yn = p.y + sigma * (p.b - dot);
        if (BOUNDS) yn = fmin(fmax(yn, ld_ro(lp.ylo + r)), ld_ro(lp.yhi + r));
        const double d = yn - p.y;
        acc[0] += d * d;
        lp.y[r] = lam * (2.0 * yn - p.y) + (1.0 - lam) * p.y0;
    }
};

// KKT scalars, A' side: r = c - A'y.
// acc: 0 pobj, 1 dobj bound terms, 2 dual residual^2, 3 ||c||^2, 4 ||x||^2, 5 ||x - x0||^2
template <bool BOUNDS>
struct EvalPrimalOp {
    const DevLP& lp;
    struct Pre { double c, x, x0; };
    __device__ __forceinline__ const double* vec() const { return lp.y; }
    __device__ __forceinline__ Pre prefetch(int r) const
    {
        return {ld_ro(lp.c + r), ld_mut(lp.x + r), ld_mut(lp.x0 + r)};
    }
    __device__ __forceinline__ void row(int r, double dot, const Pre& p, double* acc) const
    {
        const double rc = p.c - dot;
        const double rp = rc > 0.0 ? rc : 0.0, rn = rc < 0.0 ? rc : 0.0;
        double lo = 0.0, hi = INFINITY;
        if (BOUNDS) { lo = ld_ro(lp.lb + r); hi = ld_ro(lp.ub + r); }
        double viol = 0.0, dob = 0.0;
        if (isinf(hi)) viol += rn * rn; else dob += hi * rn;
        if (isinf(lo)) viol += rp * rp; else dob += lo * rp;
        acc[0] += p.c * p.x;
        acc[1] += dob;
        acc[2] += viol;
        acc[3] += p.c * p.c;
        acc[4] += p.x * p.x;
        acc[5] += (p.x - p.x0) * (p.x - p.x0);
    }
};

// KKT scalars, A side: res = Ax - b.
// acc: 0 b'y, 1 primal residual^2, 2 ||b||^2, 3 ||y||^2, 4 ||y - y0||^2
template <bool BOUNDS>
struct EvalDualOp {
    const DevLP& lp;
    struct Pre { double b, y, y0; };
    __device__ __forceinline__ const double* vec() const { return lp.x; }
    __device__ __forceinline__ Pre prefetch(int r) const
    {
        return {ld_ro(lp.b + r), ld_mut(lp.y + r), ld_mut(lp.y0 + r)};
    }
    __device__ __forceinline__ void row(int r, double dot, const Pre& p, double* acc) const
    {
        double res = dot - p.b;
        if (BOUNDS) {
            const double lo = ld_ro(lp.ylo + r), hi = ld_ro(lp.yhi + r);
            if (res > 0.0 && isinf(hi) && lo == 0.0) res = 0.0;
            if (res < 0.0 && isinf(lo) && hi == 0.0) res = 0.0;
        }
        acc[0] += p.b * p.y;
        acc[1] += res * res;
        acc[2] += p.b * p.b;
        acc[3] += p.y * p.y;
        acc[4] += (p.y - p.y0) * (p.y - p.y0);
    }
};

struct SpmvOp {
    const double* in;
    double* out;
    struct Pre {};
    __device__ __forceinline__ const double* vec() const { return in; }
    __device__ __forceinline__ Pre prefetch(int) const { return {}; }
    __device__ __forceinline__ void row(int r, double dot, const Pre&, double*) const { out[r] = dot; }
};

// ---------------------------------------------------------------------------------------
// The tile walker.  Every warp of the CTA takes tiles warp, warp+nwarps, ... of the CTA's
// range.  `vals`/`idx` may point to global memory or to the CTA's shared-memory copy; in
// both cases they are indexed by (step - step_base).
template <class Op>
__device__ __forceinline__ void run_phase(const DevMat& M, const Op& op, double* acc,
                                          const double2* __restrict__ vals,
                                          const int2* __restrict__ idx, uint32_t step_base)
{
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    const uint32_t t0 = __ldg(M.cta_begin + blockIdx.x);
    const uint32_t t1 = __ldg(M.cta_begin + blockIdx.x + 1);
    const double* __restrict__ vec = op.vec();

    for (uint32_t t = t0 + warp; t < t1; t += nwarps) {
        const int4 raw = __ldg(reinterpret_cast<const int4*>(M.tiles + t));
        const uint32_t off = (uint32_t)raw.x - step_base;
        const uint32_t row_base = (uint32_t)raw.y;
        const int nsteps = raw.z & 0xffff;
        const int logL = (raw.z >> 16) & 0xff;
        const int nrows = (raw.z >> 24) & 0xff;
        const int split = raw.w;
        const int L = 1 << logL;
        const int rr = lane >> logL;
        const bool owner = ((lane & (L - 1)) == 0) && (rr < nrows) && (split < 0);
        const int r = (int)row_base + rr;

        typename Op::Pre pre{};
        if (owner) pre = op.prefetch(r);

        const double2* vp = vals + (size_t)off * 32 + lane;
        const int2* ip = idx + (size_t)off * 32 + lane;
        double dot = 0.0;
        for (int s0 = 0; s0 < nsteps; s0 += 4) {
            double2 v[4];
            int2 j[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (s0 + u < nsteps) {
                    v[u] = vp[(s0 + u) * 32];
                    j[u] = ip[(s0 + u) * 32];
                } else {
                    v[u] = make_double2(0.0, 0.0);
                    j[u] = make_int2(0, 0);
                }
            }
            double g[8];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (s0 + u < nsteps) {
                    g[2 * u] = ld_mut(vec + j[u].x);
                    g[2 * u + 1] = ld_mut(vec + j[u].y);
                } else {
                    g[2 * u] = 0.0;
                    g[2 * u + 1] = 0.0;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                dot = fma(v[u].x, g[2 * u], dot);
                dot = fma(v[u].y, g[2 * u + 1], dot);
            }
        }
        for (int o = L >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(FULL, dot, o);

        if (split < 0) {
            if (owner) op.row(r, dot, pre, acc);
        } else {
            // chunk of a split row: park the partial, last arrival sums them in index order
            const int4 sraw = __ldg(reinterpret_cast<const int4*>(M.splits + split));
            const uint32_t srow = (uint32_t)sraw.x, first = (uint32_t)sraw.y, nch = (uint32_t)sraw.z;
            unsigned last = 0;
            if (lane == 0) {
                __stcg(M.partials + first + row_base, dot);
                __threadfence();
                const unsigned old = atomicAdd(M.counters + split, 1u);
                last = (old == nch - 1);
            }
            last = __shfl_sync(FULL, last, 0);
            if (last) {
                __threadfence();
                double s = 0.0;
                for (uint32_t k = lane; k < nch; k += 32) s += __ldcg(M.partials + first + k);
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
                if (lane == 0) {
                    M.counters[split] = 0;
                    op.row((int)srow, s, op.prefetch((int)srow), acc);
                }
            }
        }
    }
}

// Sum acc[0..NUSED) over the CTA in a fixed order and store to out[0..NUSED).
template <int NUSED>
__device__ __forceinline__ void cta_reduce_store(double* acc, double* out, double* smem /*[32*NRED]*/)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < NUSED; ++k) {
        double v = acc[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        if (lane == 0) smem[warp * NRED + k] = v;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NUSED; ++k) {
            double v = lane < nwarps ? smem[lane * NRED + k] : 0.0;
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
            if (lane == 0) __stcg(out + k, v);
        }
    }
    __syncthreads();
}

// Sum the per-CTA partials red[g][k], g = 0..G-1, in a fixed order; result in every lane of
// the calling warp.  (G <= a few hundred.)
__device__ __forceinline__ double grid_sum(const double* red, int G, int k)
{
    const int lane = threadIdx.x & 31;
    double v = 0.0;
    for (int g = lane; g < G; g += 32) v += __ldcg(red + (size_t)g * NRED + k);
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

}  // namespace mllp
