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
#include <type_traits>

#include "lp_format.h"

namespace mllp {

constexpr int NRED = 8;            // reduction slots per phase
constexpr unsigned FULL = 0xffffffffu;
#ifndef MLLP_BARRIER_BACKOFF
#define MLLP_BARRIER_BACKOFF 40
#endif
// kinds of per-CTA reduction buffers
constexpr int RED_STEPP = 0, RED_STEPD = 1, RED_EVALP = 2, RED_EVALD = 3, RED_BUFFERS = 8;
// control block slots (device doubles)
constexpr int CTRL_TAU = 0, CTRL_SIGMA = 1, CTRL_ERR = 15, CTRL_SIZE = 16;

struct DevMat {
    const double2* vals;           // [total_steps][32]
    const int2* idx;               // [total_steps][32]
    const Tile* tiles;
    const uint32_t* cta_begin;     // [G+1]
    const uint32_t* cta_step_begin;// [G+1]
    const SplitRow* splits;
    const LocalSplit* lsplits;     // CTA-major
    const uint32_t* cta_lsplit_begin; // [G+1]
    const uint32_t* cta_nsplit;    // [G] leading split-chunk tiles of each CTA
    double* partials;              // one slot per (CTA, split row)
    double* slots;                 // 16 B per (CTA, split row): {partial, partial ^ tag} of the polled join
    unsigned* counters;            // one per split row
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
    // Preconditioned handles (MLLP_F_PRECONDITION): the matrix held is Dr A Dc and the iteration runs on the scaled LP
    // (x~ = x / dc, y~ = y / dr, b~ = dr b, c~ = dc c, box / dc); the KKT scalars are evaluated on the ORIGINAL LP.
    const double* dr;   // null => 1
    const double* dc;   // null => 1
    double* x;
    double* y;
    double* xbar;
    double* x0;         // Halpern anchors
    double* y0;
    double* red;        // [2 parity][4 kinds][G][NRED] per-CTA partial sums (RED_* below)
    unsigned* barrier;  // grid barrier counter
    uint32_t res_steps_A;   // per-CTA cap of shared-memory resident warp-steps of A / A'
    uint32_t res_steps_AT;
    uint32_t own_rows_A;    // parity kernel: per-CTA capacity of shared-memory resident own entries (y, b / x, c);
    uint32_t own_rows_AT;   // both 0 = own entries stay in global memory
    double* ctrl;       // control block (device doubles): CTRL_* slots
    unsigned long long join_base;  // tags of the polled split-row join start above this value
    unsigned long long* trace;  // dev tool: [iter][cta][4] barrier timestamps, or null
    int sync_mode;      // how the CTAs of the persistent kernels meet between phases: SYNC_* below
};
// SYNC_GRID: cooperative grid of one CTA per SM, counter barrier in global memory (~0.9 us).  SYNC_CLUSTER: the
// whole grid is ONE thread-block cluster (<= 16 CTAs) and meets on the hardware cluster barrier (~0.2 us): small and
// mid-size LPs, whose phases are shorter than a grid barrier.  SYNC_CTA: one CTA, __syncthreads().
// SYNC_BCAST: one cluster whose CTAs each keep a full copy of the two gathered vectors (y, xbar) in shared memory; a row
// update stores its new entry into every CTA's copy over distributed shared memory, gathers are local LDS and the
// rows' own entries (x, c, x0 / y, b, y0) stay in shared-memory slots: no global-memory traffic inside the iteration.
constexpr int SYNC_GRID = 0, SYNC_CLUSTER = 1, SYNC_CTA = 2, SYNC_BCAST = 3;

// Memory policies of the row ops.  GlobalMem: vectors live in global memory and are shared by
// all CTAs (single-instance path).  SmemMem: vectors are the CTA's own shared-memory copies
// (batched path: one LP per CTA), reached through generic pointers.
struct GlobalMem {
    static __device__ __forceinline__ double ld_mut(const double* p) { return __ldcg(p); }  // L2-coherent
    static __device__ __forceinline__ double ld_ro(const double* p) { return __ldg(p); }    // read-only path
    static __device__ __forceinline__ double gather(const double* p) { return __ldca(p); }  // L1, see tile_dot
};
// ClusterMem: the gathered vector is this CTA's shared-memory copy (SYNC_BCAST); own entries without a slot (rows of
// split chunks) stay in global memory.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
struct ClusterMem {
    static __device__ __forceinline__ double ld_mut(const double* p) { return __ldcg(p); }
    static __device__ __forceinline__ double ld_ro(const double* p) { return __ldg(p); }
    static __device__ __forceinline__ double gather(const double* p)
    {
        double v;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(smem_u32(p)) : "memory");
        return v;
    }
};
// One vector entry stored into the copies of all CTAs of the cluster (own copy included).
struct Bcast {
    uint32_t base;   // shared-memory address of the copy (the layout is the same in every CTA)
    int nctas;
    __device__ __forceinline__ void put(int r, double v) const
    {
        const uint32_t a = base + 8u * (uint32_t)r;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            if (q < nctas) {
                uint32_t ra;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(q));   // volatile: never speculated for q >= nctas
                asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(ra), "d"(v) : "memory");
            }
        }
    }
};
struct SmemMem {
    static __device__ __forceinline__ double ld_mut(const double* p) { return *p; }
    static __device__ __forceinline__ double ld_ro(const double* p) { return *p; }
    static __device__ __forceinline__ double gather(const double* p) { return *p; }
};

// Tagged 16-byte words {bits(value), bits(value) ^ tag}: value and validity travel in ONE 128-bit access, so no fence
// orders them -- a word whose halves do not xor to the expected tag is stale or torn and is simply read again.  (The
// polled join of split rows below and the linking rows of blocks.cu cross CTAs this way.)
__device__ __forceinline__ void ld_tagged(const unsigned long long* p, unsigned long long& a, unsigned long long& b)
{
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void st_tagged(unsigned long long* p, double v, unsigned long long tag)
{
    const unsigned long long a = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(a ^ tag) : "memory");
}

// ---------------------------------------------------------------------------------------
// Grid barrier for the persistent cooperative kernel: monotonic counter, one arrival per
// CTA (release), acquire-poll.  `target` lives in thread 0's register.
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// `trace` (dev tool, may be null): thread 0 stores the time at which the whole CTA had arrived
// and the time at which the barrier released it.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& target, unsigned long long* trace = nullptr)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        if (trace) trace[0] = global_ns();
        target += gridDim.x;
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
        unsigned v;
        for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            if ((int)(v - target) >= 0) break;
            __nanosleep(MLLP_BARRIER_BACKOFF);  // keeps the pollers from crowding out the arrivals
        }
        if (trace) trace[1] = global_ns();
    }
    __syncthreads();
}

// Phase boundary of the persistent kernels (see SYNC_*).  The cluster barrier's release / acquire pair orders the
// phase's global-memory writes before the other CTAs' reads and invalidates L1 like the acquire poll of the grid
// barrier does, so the L1-cached gathers stay valid; inside one CTA __syncthreads() gives the same guarantee.
__device__ __forceinline__ void sync_all(const DevLP& lp, unsigned& target, unsigned long long* trace = nullptr)
{
    if (lp.sync_mode == SYNC_GRID) {
        grid_barrier(lp.barrier, target, trace);
        return;
    }
    if (trace) {
        __syncthreads();
        if (threadIdx.x == 0) trace[0] = global_ns();
    }
    if (lp.sync_mode != SYNC_CTA && gridDim.x > 1) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
        __syncthreads();
    }
    if (trace && threadIdx.x == 0) trace[1] = global_ns();
}

// ---------------------------------------------------------------------------------------
// Ops.  Interface:
//   const double* vec()                     gather vector (mutable => read through L2)
//   Pre  prefetch(int r)                    issue the row's own vector loads early
//   void row(int r, double dot, Pre)        consume a finished dot product
//   void row_late(int r, double dot)        same, without prefetch (split rows)
// Each op may accumulate into acc[NRED] (thread-local), reduced per CTA at the phase end.

template <bool BOUNDS, class MEM = GlobalMem>
struct PrimalOp {
    using Mem = MEM;
    const DevLP& lp;
    double tau;
    struct Pre { double c, x; };
    __device__ __forceinline__ const double* vec() const { return lp.y; }
    __device__ __forceinline__ Pre prefetch(int r) const { return {MEM::ld_ro(lp.c + r), MEM::ld_mut(lp.x + r)}; }
    __device__ __forceinline__ void row(int r, double dot, const Pre& p, double*) const
    {
        const double g = p.c - dot;
        double xn = p.x - tau * g;
        if (BOUNDS) {
            xn = fmin(fmax(xn, MEM::ld_ro(lp.lb + r)), MEM::ld_ro(lp.ub + r));
        } else {
            xn = fmax(xn, 0.0);
        }
        lp.xbar[r] = 2.0 * xn - p.x;
        lp.x[r] = xn;
    }
};

template <bool BOUNDS, class MEM = GlobalMem>
struct DualOp {
    using Mem = MEM;
    const DevLP& lp;
    double sigma;
    struct Pre { double b, y; };
    __device__ __forceinline__ const double* vec() const { return lp.xbar; }
    __device__ __forceinline__ Pre prefetch(int r) const { return {MEM::ld_ro(lp.b + r), MEM::ld_mut(lp.y + r)}; }
    __device__ __forceinline__ void row(int r, double dot, const Pre& p, double*) const
    {
        double yn = p.y + sigma * (p.b - dot);
        if (BOUNDS) yn = fmin(fmax(yn, MEM::ld_ro(lp.ylo + r)), MEM::ld_ro(lp.yhi + r));
        lp.y[r] = yn;
    }
};

// Ops whose rows' OWN vector entries (x, c / y, b) live in the CTA's shared memory for the whole persistent
// kernel (`slot` = CTA-local row slot, -1 for the rows of split chunks, which stay in global memory).  The
// tile -> CTA assignment is fixed, so x and c are read from L2 once per launch instead of once per iteration and
// x is written back once at the end; only the gathered vectors (xbar, y) go to global memory every iteration.
template <class T, class = void>
struct op_slots : std::false_type {};
template <class T>
struct op_slots<T, std::void_t<decltype(T::kSlots)>> : std::bool_constant<T::kSlots> {};

template <bool BOUNDS>
struct PrimalResOp {
    using Mem = GlobalMem;
    static constexpr bool kSlots = true;
    const DevLP& lp;
    double tau;
    double* xs;          // shared: own x
    const double* cs;    // shared: own c
    struct Pre { double c, x; };
    __device__ __forceinline__ const double* vec() const { return lp.y; }
    __device__ __forceinline__ Pre prefetch(int r, int slot) const
    {
        if (slot >= 0) return {cs[slot], xs[slot]};
        return {Mem::ld_ro(lp.c + r), Mem::ld_mut(lp.x + r)};
    }
    __device__ __forceinline__ void row(int r, int slot, double dot, const Pre& p, double*) const
    {
        const double g = p.c - dot;
        double xn = p.x - tau * g;
        if (BOUNDS) {
            xn = fmin(fmax(xn, Mem::ld_ro(lp.lb + r)), Mem::ld_ro(lp.ub + r));
        } else {
            xn = fmax(xn, 0.0);
        }
        lp.xbar[r] = 2.0 * xn - p.x;
        if (slot >= 0) xs[slot] = xn; else lp.x[r] = xn;
    }
};

template <bool BOUNDS>
struct DualResOp {
    using Mem = GlobalMem;
    static constexpr bool kSlots = true;
    const DevLP& lp;
    double sigma;
    double* ys;          // shared: own y
    const double* bs;    // shared: own b
    struct Pre { double b, y; };
    __device__ __forceinline__ const double* vec() const { return lp.xbar; }
    __device__ __forceinline__ Pre prefetch(int r, int slot) const
    {
        if (slot >= 0) return {bs[slot], ys[slot]};
        return {Mem::ld_ro(lp.b + r), Mem::ld_mut(lp.y + r)};
    }
    __device__ __forceinline__ void row(int r, int slot, double dot, const Pre& p, double*) const
    {
        double yn = p.y + sigma * (p.b - dot);
        if (BOUNDS) yn = fmin(fmax(yn, Mem::ld_ro(lp.ylo + r)), Mem::ld_ro(lp.yhi + r));
        lp.y[r] = yn;    // y is the gather vector of the A' phase: always published
        if (slot >= 0) ys[slot] = yn;
    }
};

// SYNC_BCAST ops: own entries in shared-memory slots (global memory for rows without a slot), the new entry of the
// gathered vector goes to every CTA's copy.
template <bool BOUNDS>
struct PrimalBcastOp {
    using Mem = ClusterMem;
    static constexpr bool kSlots = true;
    const DevLP& lp;
    double tau;
    double* xs;
    const double* cs;
    const double* ycopy;
    Bcast xb;
    struct Pre { double c, x; };
    __device__ __forceinline__ const double* vec() const { return ycopy; }
    __device__ __forceinline__ Pre prefetch(int r, int slot) const
    {
        if (slot >= 0) return {cs[slot], xs[slot]};
        return {Mem::ld_ro(lp.c + r), Mem::ld_mut(lp.x + r)};
    }
    __device__ __forceinline__ void row(int r, int slot, double dot, const Pre& p, double*) const
    {
        const double g = p.c - dot;
        double xn = p.x - tau * g;
        if (BOUNDS) xn = fmin(fmax(xn, Mem::ld_ro(lp.lb + r)), Mem::ld_ro(lp.ub + r));
        else xn = fmax(xn, 0.0);
        xb.put(r, 2.0 * xn - p.x);
        if (slot >= 0) xs[slot] = xn; else lp.x[r] = xn;
    }
};

template <bool BOUNDS>
struct DualBcastOp {
    using Mem = ClusterMem;
    static constexpr bool kSlots = true;
    const DevLP& lp;
    double sigma;
    double* ys;
    const double* bs;
    const double* xbcopy;
    Bcast yb;
    struct Pre { double b, y; };
    __device__ __forceinline__ const double* vec() const { return xbcopy; }
    __device__ __forceinline__ Pre prefetch(int r, int slot) const
    {
        if (slot >= 0) return {bs[slot], ys[slot]};
        return {Mem::ld_ro(lp.b + r), Mem::ld_mut(lp.y + r)};
    }
    __device__ __forceinline__ void row(int r, int slot, double dot, const Pre& p, double*) const
    {
        double yn = p.y + sigma * (p.b - dot);
        if (BOUNDS) yn = fmin(fmax(yn, Mem::ld_ro(lp.ylo + r)), Mem::ld_ro(lp.yhi + r));
        yb.put(r, yn);
        if (slot >= 0) ys[slot] = yn; else lp.y[r] = yn;
    }
};

template <bool BOUNDS>
struct PrimalHalpernBcastOp {
    using Mem = ClusterMem;
    static constexpr bool kSlots = true;
    const DevLP& lp;
    double tau, lam;
    double* xs;
    const double* cs;
    const double* x0s;
    const double* ycopy;
    Bcast xb;
    struct Pre { double c, x, x0; };
    __device__ __forceinline__ const double* vec() const { return ycopy; }
    __device__ __forceinline__ Pre prefetch(int r, int slot) const
    {
        if (slot >= 0) return {cs[slot], xs[slot], x0s[slot]};
        return {Mem::ld_ro(lp.c + r), Mem::ld_mut(lp.x + r), Mem::ld_mut(lp.x0 + r)};
    }
    __device__ __forceinline__ void row(int r, int slot, double dot, const Pre& p, double* acc) const
    {
        const double g = p.c - dot;
        double xn = p.x - tau * g;
        if (BOUNDS) xn = fmin(fmax(xn, Mem::ld_ro(lp.lb + r)), Mem::ld_ro(lp.ub + r));
        else xn = fmax(xn, 0.0);
        const double d = xn - p.x;
        acc[0] += d * d;
        const double xbv = 2.0 * xn - p.x;
        xb.put(r, xbv);
        const double xh = lam * xbv + (1.0 - lam) * p.x0;
        if (slot >= 0) xs[slot] = xh; else lp.x[r] = xh;
    }
};

template <bool BOUNDS>
struct DualHalpernBcastOp {
    using Mem = ClusterMem;
    static constexpr bool kSlots = true;
    const DevLP& lp;
    double sigma, lam;
    double* ys;
    const double* bs;
    const double* y0s;
    const double* xbcopy;
    Bcast yb;
    struct Pre { double b, y, y0; };
    __device__ __forceinline__ const double* vec() const { return xbcopy; }
    __device__ __forceinline__ Pre prefetch(int r, int slot) const
    {
        if (slot >= 0) return {bs[slot], ys[slot], y0s[slot]};
        return {Mem::ld_ro(lp.b + r), Mem::ld_mut(lp.y + r), Mem::ld_mut(lp.y0 + r)};
    }
    __device__ __forceinline__ void row(int r, int slot, double dot, const Pre& p, double* acc) const
    {
        double yn = p.y + sigma * (p.b - dot);
        if (BOUNDS) yn = fmin(fmax(yn, Mem::ld_ro(lp.ylo + r)), Mem::ld_ro(lp.yhi + r));
        const double d = yn - p.y;
        acc[0] += d * d;
        const double yh = lam * (2.0 * yn - p.y) + (1.0 - lam) * p.y0;
        yb.put(r, yh);
        if (slot >= 0) ys[slot] = yh; else lp.y[r] = yh;
    }
};

template <class Op>
__device__ __forceinline__ typename Op::Pre op_prefetch(const Op& op, int r, int slot)
{
    if constexpr (op_slots<Op>::value) return op.prefetch(r, slot);
    else return op.prefetch(r);
}
template <class Op>
__device__ __forceinline__ void op_row(const Op& op, int r, int slot, double dot, const typename Op::Pre& p, double* acc)
{
    if constexpr (op_slots<Op>::value) op.row(r, slot, dot, p, acc);
    else op.row(r, dot, p, acc);
}

// Solve mode (reflected Halpern).  acc[0] accumulates ||x' - x||^2 (resp. y).
template <bool BOUNDS, class MEM = GlobalMem>
struct PrimalHalpernOp {
    using Mem = MEM;
    const DevLP& lp;
    double tau, lam;
    struct Pre { double c, x, x0; };
    __device__ __forceinline__ const double* vec() const { return lp.y; }
    __device__ __forceinline__ Pre prefetch(int r) const
    {
        return {MEM::ld_ro(lp.c + r), MEM::ld_mut(lp.x + r), MEM::ld_mut(lp.x0 + r)};
    }
    __device__ __forceinline__ void row(int r, double dot, const Pre& p, double* acc) const
    {
        const double g = p.c - dot;
        double xn = p.x - tau * g;
        if (BOUNDS) {
            xn = fmin(fmax(xn, MEM::ld_ro(lp.lb + r)), MEM::ld_ro(lp.ub + r));
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

template <bool BOUNDS, class MEM = GlobalMem>
struct DualHalpernOp {
    using Mem = MEM;
    const DevLP& lp;
    double sigma, lam;
    struct Pre { double b, y, y0; };
    __device__ __forceinline__ const double* vec() const { return lp.xbar; }
    __device__ __forceinline__ Pre prefetch(int r) const
    {
        return {MEM::ld_ro(lp.b + r), MEM::ld_mut(lp.y + r), MEM::ld_mut(lp.y0 + r)};
    }
    __device__ __forceinline__ void row(int r, double dot, const Pre& p, double* acc) const
    {
        double yn = p.y + sigma * (p.b - dot);
        if (BOUNDS) yn = fmin(fmax(yn, MEM::ld_ro(lp.ylo + r)), MEM::ld_ro(lp.yhi + r));
        const double d = yn - p.y;
        acc[0] += d * d;
        lp.y[r] = lam * (2.0 * yn - p.y) + (1.0 - lam) * p.y0;
    }
};

// KKT scalars, A' side: r = c - A'y.  All sums refer to the ORIGINAL LP: on a preconditioned handle the column's scale
// s = dc_j turns the scaled quantities back (r = r~ / s, x = s x~, c = c~ / s; c'x and the bound terms u r^- are invariant).
// acc: 0 pobj, 1 dobj bound terms, 2 dual residual^2, 3 ||c||^2, 4 ||x||^2, 5 ||x~ - x~0||^2 (scaled: feeds the primal
// weight), 6 distance^2 of x from its box (part of the primal residual: the Halpern combinations can leave the box)
template <bool BOUNDS, class MEM = GlobalMem>
struct EvalPrimalOp {
    using Mem = MEM;
    const DevLP& lp;
    struct Pre { double c, x, x0; };
    __device__ __forceinline__ const double* vec() const { return lp.y; }
    __device__ __forceinline__ Pre prefetch(int r) const
    {
        return {MEM::ld_ro(lp.c + r), MEM::ld_mut(lp.x + r), MEM::ld_mut(lp.x0 + r)};
    }
    __device__ __forceinline__ void row(int r, double dot, const Pre& p, double* acc) const
    {
        const double rc = p.c - dot;
        const double rp = rc > 0.0 ? rc : 0.0, rn = rc < 0.0 ? rc : 0.0;
        double lo = 0.0, hi = INFINITY;
        if (BOUNDS) { lo = MEM::ld_ro(lp.lb + r); hi = MEM::ld_ro(lp.ub + r); }
        double viol = 0.0, dob = 0.0;
        if (isinf(hi)) viol += rn * rn; else dob += hi * rn;
        if (isinf(lo)) viol += rp * rp; else dob += lo * rp;
        const double xv = p.x - fmin(fmax(p.x, lo), hi);
        const double s = lp.dc ? __ldg(lp.dc + r) : 1.0;
        const double inv = 1.0 / s;
        acc[0] += p.c * p.x;
        acc[1] += dob;
        acc[2] += viol * (inv * inv);
        acc[3] += (p.c * inv) * (p.c * inv);
        acc[4] += (p.x * s) * (p.x * s);
        acc[5] += (p.x - p.x0) * (p.x - p.x0);
        acc[6] += (xv * s) * (xv * s);
    }
};

// KKT scalars, A side: res = Ax - b (original LP: res = res~ / s, b = b~ / s, y = s y~ with s = dr_i).
// acc: 0 b'y, 1 primal residual^2, 2 ||b||^2, 3 ||y||^2, 4 ||y~ - y~0||^2 (scaled), 5 distance^2 of y from its cone
template <bool BOUNDS, class MEM = GlobalMem>
struct EvalDualOp {
    using Mem = MEM;
    const DevLP& lp;
    struct Pre { double b, y, y0; };
    __device__ __forceinline__ const double* vec() const { return lp.x; }
    __device__ __forceinline__ Pre prefetch(int r) const
    {
        return {MEM::ld_ro(lp.b + r), MEM::ld_mut(lp.y + r), MEM::ld_mut(lp.y0 + r)};
    }
    __device__ __forceinline__ void row(int r, double dot, const Pre& p, double* acc) const
    {
        double res = dot - p.b;
        double yv = 0.0;
        if (BOUNDS) {
            const double lo = MEM::ld_ro(lp.ylo + r), hi = MEM::ld_ro(lp.yhi + r);
            if (res > 0.0 && isinf(hi) && lo == 0.0) res = 0.0;
            if (res < 0.0 && isinf(lo) && hi == 0.0) res = 0.0;
            yv = p.y - fmin(fmax(p.y, lo), hi);
        }
        const double s = lp.dr ? __ldg(lp.dr + r) : 1.0;
        const double inv = 1.0 / s;
        acc[0] += p.b * p.y;
        acc[1] += (res * inv) * (res * inv);
        acc[2] += (p.b * inv) * (p.b * inv);
        acc[3] += (p.y * s) * (p.y * s);
        acc[4] += (p.y - p.y0) * (p.y - p.y0);
        acc[5] += (yv * s) * (yv * s);
    }
};

template <class MEM = GlobalMem>
struct SpmvOp {
    using Mem = MEM;
    const double* in;
    double* out;
    struct Pre {};
    __device__ __forceinline__ const double* vec() const { return in; }
    __device__ __forceinline__ Pre prefetch(int) const { return {}; }
    __device__ __forceinline__ void row(int r, double dot, const Pre&, double*) const { out[r] = dot; }
};

// ---------------------------------------------------------------------------------------
// Row partition of ONE large LP over GPUs (BASELINE.json configs[3]) with the exchange INSIDE the persistent kernel.
// Rank p owns a slice of the rows of A (its entries of y); the cheap A' phase is replicated (every rank updates the
// whole of x from the whole of y), so only y crosses GPUs: ONE exchange per iteration.  The exchange has no fence, no
// flag and no acknowledgement: a row update stores its new dual value into every peer's MAILBOX over NVLink as one
// tagged 16-byte word {bits(y), bits(y) ^ tag} (value and validity travel in one access; a word whose halves do not xor
// to this iteration's tag is stale or torn and is read again), and every rank unpacks the words of the other ranks'
// slices into its own y (coalesced polls of its local memory) before the local grid barrier that ends the iteration.
// Two mailbox buffers alternate: rank p overwrites a word of iteration k only in iteration k + 2, which it starts after
// it has received q's words of iteration k + 1, which q sent after unpacking iteration k.
constexpr int MAX_RANKS = 8;
struct PeerInfo {
    unsigned long long* mail[MAX_RANKS];   // every rank's mailbox [2][mi][2] (own entry = local pointer)
    unsigned long long* mc;                // multicast address of all mailboxes (one store reaches every rank), or null
    int cnt[MAX_RANKS];                    // valid rows at the head of every rank's slice of y
    unsigned* err;                         // local: set to 1 when a wait timed out
    int rank, nranks, Ly, mi;              // slice length, padded length of y (= nranks * Ly)
    int backoff_ns, backoff_max_ns;        // sleep between two looks at a word that has not arrived (doubles up to the cap)
    int st_mode;                           // dev knob: how a word is stored into a peer's mailbox (see st_mail)
};

__device__ __forceinline__ void st_mail(unsigned long long* p, double v, unsigned long long tag, int mode = 0)
{
    const unsigned long long a = (unsigned long long)__double_as_longlong(v);
    if (mode == 1) asm volatile("st.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(a ^ tag) : "memory");              // weak
    else if (mode == 2) asm volatile("st.global.cg.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(a ^ tag) : "memory");      // weak, L2 only
    else if (mode == 3) asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(a ^ tag) : "memory");
    else asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(a ^ tag) : "memory");
}
__device__ __forceinline__ void ld_mail(const unsigned long long* p, unsigned long long& a, unsigned long long& b)
{
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}

// Dual update of this rank's rows: y goes to the local vector and, tagged, to every peer's mailbox.
template <bool BOUNDS>
struct DualMailOp {
    using Mem = GlobalMem;
    const DevLP& lp;
    const PeerInfo& pi;
    double sigma;
    unsigned long long tag;
    size_t buf;            // word offset of this iteration's mailbox buffer
    struct Pre { double b, y; };
    __device__ __forceinline__ const double* vec() const { return lp.xbar; }
    __device__ __forceinline__ Pre prefetch(int r) const { return {Mem::ld_ro(lp.b + r), Mem::ld_mut(lp.y + r)}; }
    __device__ __forceinline__ void row(int r, double dot, const Pre& p, double*) const
    {
        double yn = p.y + sigma * (p.b - dot);
        if (BOUNDS) yn = fmin(fmax(yn, Mem::ld_ro(lp.ylo + r)), Mem::ld_ro(lp.yhi + r));
        lp.y[r] = yn;
        if (pi.mc) {   // NVSwitch multicast: ONE 16-byte store, replicated by the switch into every rank's mailbox
            const unsigned long long a = (unsigned long long)__double_as_longlong(yn), b = a ^ tag;
            asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(pi.mc + buf + 2 * (size_t)r),
                         "r"((unsigned)a), "r"((unsigned)(a >> 32)), "r"((unsigned)b), "r"((unsigned)(b >> 32)) : "memory");
            return;
        }
#pragma unroll
        for (int q = 0; q < MAX_RANKS; ++q)
            if (q < pi.nranks && q != pi.rank) st_mail(pi.mail[q] + buf + 2 * (size_t)r, yn, tag, pi.st_mode);
    }
};

// Unpack the peers' slices of y from this rank's mailbox (all threads of the grid, coalesced).  Waits give up after
// ~10 s and raise pi.err instead of hanging the GPU; once it is raised every later wait gives up after one more look,
// so the launch runs out quickly (with meaningless iterates) and the host reports the error (mllp_rowpart_error).
__device__ __forceinline__ void unpack_mail(const DevLP& lp, const PeerInfo& pi, unsigned long long tag, size_t buf)
{
    const unsigned long long* box = pi.mail[pi.rank] + buf;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    // ONE flat index space over the words of all peers, so that the whole grid shares them: with a loop over the peers
    // inside every thread the first cnt threads polled world - 1 words one after the other (a dependent L2 round trip each)
    // while the rest of the grid idled -- the wait grew with the rank count (ken-18: 2.5 / 5.2 / 9.9 us on 2 / 4 / 8 GPUs)
    int total = 0;
#pragma unroll
    for (int q = 0; q < MAX_RANKS; ++q)
        if (q < pi.nranks && q != pi.rank) total += pi.cnt[q];
    for (int f = gtid; f < total; f += gsz) {
        int k = f, q = 0;
#pragma unroll
        for (int p = 0; p < MAX_RANKS; ++p) {       // flat index -> (peer q, row k within its slice)
            const int c = (p < pi.nranks && p != pi.rank) ? pi.cnt[p] : 0;
            if (k < c) { q = p; break; }
            k -= c;
        }
        const int at = q * pi.Ly + k;
        const unsigned long long* w = box + 2 * (size_t)at;
        unsigned long long a, b;
        ld_mail(w, a, b);
        if ((a ^ b) != tag) {
            const unsigned long long t0 = global_ns();
            unsigned ns = (unsigned)pi.backoff_ns;
            for (;;) {
                __nanosleep(ns);
                if (ns < (unsigned)pi.backoff_max_ns) ns *= 2;
                ld_mail(w, a, b);
                if ((a ^ b) == tag) break;
                if (global_ns() - t0 > 10000000000ull || *(volatile unsigned*)pi.err != 0u) {
                    atomicExch(pi.err, 1u);
                    break;
                }
            }
        }
        lp.y[at] = __longlong_as_double((long long)a);
    }
}

// ---------------------------------------------------------------------------------------
// Per-CTA view of a matrix.  The CTA's tile descriptors and a prefix of its tiles' data can
// live in shared memory for the whole persistent kernel (the matrix never changes), the rest
// is streamed from global memory / L2.
struct MatView {
    const Tile* desc;        // this CTA's tiles, index 0 .. ntiles-1 (shared or global memory)
    const double2* rvals;    // resident copy of the first res_steps warp-steps (shared), or null
    const int2* ridx;
    const double2* gvals;    // global arrays, indexed by absolute warp-step
    const int2* gidx;
    uint32_t ntiles, res_steps, step0;   // step0 = first absolute warp-step of this CTA
    uint32_t nsplit;         // leading split-chunk tiles
    uint32_t ls0, nls;       // this CTA's LocalSplit range
};

__device__ __forceinline__ void view_common(const DevMat& M, MatView& V, uint32_t cta)
{
    V.nsplit = __ldg(M.cta_nsplit + cta);
    V.ls0 = __ldg(M.cta_lsplit_begin + cta);
    V.nls = __ldg(M.cta_lsplit_begin + cta + 1) - V.ls0;
}

// `cta` = which CTA's share of the format to walk (the batched path builds formats for a
// one-CTA grid and always walks share 0).
__device__ __forceinline__ MatView global_view(const DevMat& M, uint32_t cta)
{
    const uint32_t t0 = __ldg(M.cta_begin + cta), t1 = __ldg(M.cta_begin + cta + 1);
    MatView V;
    V.desc = M.tiles + t0;
    V.rvals = nullptr; V.ridx = nullptr;
    V.gvals = M.vals; V.gidx = M.idx;
    V.ntiles = t1 - t0; V.res_steps = 0;
    V.step0 = __ldg(M.cta_step_begin + cta);
    view_common(M, V, cta);
    return V;
}
__device__ __forceinline__ MatView global_view(const DevMat& M) { return global_view(M, blockIdx.x); }

// Copy this CTA's descriptors and its first `res_steps` warp-steps into shared memory.
// Layout at `base` (16 B aligned): desc[ntiles] | vals[res_steps*32] (double2) | idx[res_steps*32] (int2).
// Returns the number of bytes used (multiple of 16).
__device__ __forceinline__ uint32_t resident_view(const DevMat& M, uint32_t res_steps_cap, unsigned char* base,
                                                  MatView& V, uint32_t cta)
{
    const uint32_t t0 = __ldg(M.cta_begin + cta), t1 = __ldg(M.cta_begin + cta + 1);
    const uint32_t s0 = __ldg(M.cta_step_begin + cta), s1 = __ldg(M.cta_step_begin + cta + 1);
    const uint32_t nt = t1 - t0;
    const uint32_t rs = min(res_steps_cap, s1 - s0);
    int4* d_desc = reinterpret_cast<int4*>(base);
    int4* d_vals = d_desc + nt;
    int4* d_idx = d_vals + (size_t)rs * 32;
    const int4* g_desc = reinterpret_cast<const int4*>(M.tiles + t0);
    const int4* g_vals = reinterpret_cast<const int4*>(M.vals + (size_t)s0 * 32);
    const int4* g_idx = reinterpret_cast<const int4*>(M.idx + (size_t)s0 * 32);
    for (uint32_t k = threadIdx.x; k < nt; k += blockDim.x) d_desc[k] = __ldg(g_desc + k);
    for (uint32_t k = threadIdx.x; k < rs * 32; k += blockDim.x) d_vals[k] = __ldg(g_vals + k);
    for (uint32_t k = threadIdx.x; k < rs * 16; k += blockDim.x) d_idx[k] = __ldg(g_idx + k);
    V.desc = reinterpret_cast<const Tile*>(d_desc);
    V.rvals = reinterpret_cast<const double2*>(d_vals);
    V.ridx = reinterpret_cast<const int2*>(d_idx);
    V.gvals = M.vals; V.gidx = M.idx;
    V.ntiles = nt; V.res_steps = rs; V.step0 = s0;
    view_common(M, V, cta);
    return nt * 16u + rs * 768u;
}
__device__ __forceinline__ uint32_t resident_view(const DevMat& M, uint32_t res_steps_cap, unsigned char* base, MatView& V)
{
    return resident_view(M, res_steps_cap, base, V, blockIdx.x);
}

// ---------------------------------------------------------------------------------------
// The tile walker.
//
// One tile = `nsteps` warp-steps.  The common step counts (1..4) run as straight-line code
// with all gathers of the tile in flight at once; longer tiles (chunks of split rows) run in
// groups of 4 steps.  Gathers are L1-cached loads: the grid barrier's acquire invalidated this
// SM's L1 and the gathered vector is not written during the phase, so lines fetched now stay
// valid until the next barrier (hot entries -- the long rows' y -- are then served by L1).
template <class MEM, int NS>
__device__ __forceinline__ double steps_dot(const double2* __restrict__ vp, const int2* __restrict__ ip,
                                            const double* __restrict__ vec, double dot)
{
    int2 j[NS];
#pragma unroll
    for (int u = 0; u < NS; ++u) j[u] = ip[u * 32];
    double g[2 * NS];
#pragma unroll
    for (int u = 0; u < NS; ++u) {
        g[2 * u] = MEM::gather(vec + j[u].x);
        g[2 * u + 1] = MEM::gather(vec + j[u].y);
    }
#pragma unroll
    for (int u = 0; u < NS; ++u) {
        const double2 v = vp[u * 32];
        dot = fma(v.x, g[2 * u], dot);
        dot = fma(v.y, g[2 * u + 1], dot);
    }
    return dot;
}

// Two tiles of the same small step count at once: both tiles' gathers are in flight together, so a
// warp that owns two tiles of a phase pays one memory round trip instead of two.
template <class MEM, int NS>
__device__ __forceinline__ void pair_dot(const double2* __restrict__ vpa, const int2* __restrict__ ipa,
                                         const double2* __restrict__ vpb, const int2* __restrict__ ipb,
                                         const double* __restrict__ vec, double& da, double& db)
{
    int2 ja[NS], jb[NS];
#pragma unroll
    for (int u = 0; u < NS; ++u) { ja[u] = ipa[u * 32]; jb[u] = ipb[u * 32]; }
    double ga[2 * NS], gb[2 * NS];
#pragma unroll
    for (int u = 0; u < NS; ++u) {
        ga[2 * u] = MEM::gather(vec + ja[u].x); ga[2 * u + 1] = MEM::gather(vec + ja[u].y);
        gb[2 * u] = MEM::gather(vec + jb[u].x); gb[2 * u + 1] = MEM::gather(vec + jb[u].y);
    }
    da = 0.0; db = 0.0;
#pragma unroll
    for (int u = 0; u < NS; ++u) {
        const double2 va = vpa[u * 32], vb = vpb[u * 32];
        da = fma(va.x, ga[2 * u], da); da = fma(va.y, ga[2 * u + 1], da);
        db = fma(vb.x, gb[2 * u], db); db = fma(vb.y, gb[2 * u + 1], db);
    }
}

// pointers to a tile's values / indices for this lane (resident copy or global)
__device__ __forceinline__ bool tile_ptrs(const MatView& V, uint32_t off, int nsteps, int lane, const double2*& vp,
                                          const int2*& ip)
{
    const uint32_t loc = off - V.step0;
    const bool res = loc + (uint32_t)nsteps <= V.res_steps;
    vp = res ? V.rvals + (size_t)loc * 32 + lane : V.gvals + (size_t)off * 32 + lane;
    ip = res ? V.ridx + (size_t)loc * 32 + lane : V.gidx + (size_t)off * 32 + lane;
    return res;
}

// `RES`: the tile's values / indices are the CTA's shared-memory copy (tell the compiler, so the
// loads are LDS with immediate offsets instead of generic loads); otherwise global memory.
template <class MEM, bool RES>
__device__ __forceinline__ double tile_dot_at(const double2* __restrict__ vp, const int2* __restrict__ ip,
                                              const double* __restrict__ vec, int nsteps)
{
    if (RES) {
        __builtin_assume(__isShared(vp));
        __builtin_assume(__isShared(ip));
    } else {
        __builtin_assume(__isGlobal(vp));
        __builtin_assume(__isGlobal(ip));
    }
    double dot = 0.0;
    for (; nsteps > 4; nsteps -= 4, vp += 128, ip += 128) dot = steps_dot<MEM, 4>(vp, ip, vec, dot);
    switch (nsteps) {
        case 4: dot = steps_dot<MEM, 4>(vp, ip, vec, dot); break;
        case 3: dot = steps_dot<MEM, 3>(vp, ip, vec, dot); break;
        case 2: dot = steps_dot<MEM, 2>(vp, ip, vec, dot); break;
        case 1: dot = steps_dot<MEM, 1>(vp, ip, vec, dot); break;
        default: break;
    }
    return dot;
}

template <class MEM>
__device__ __forceinline__ double tile_dot(const MatView& V, const double* __restrict__ vec, uint32_t off, int nsteps,
                                           int lane)
{
    // whole tiles are resident or not (the resident region is a prefix of the CTA's steps)
    const uint32_t loc = off - V.step0;
    if (loc + (uint32_t)nsteps <= V.res_steps)
        return tile_dot_at<MEM, true>(V.rvals + (size_t)loc * 32 + lane, V.ridx + (size_t)loc * 32 + lane, vec, nsteps);
    return tile_dot_at<MEM, false>(V.gvals + (size_t)off * 32 + lane, V.gidx + (size_t)off * 32 + lane, vec, nsteps);
}

// One phase of one CTA.  Split-row chunks come first: each warp parks its chunk's partial in
// shared memory; after a CTA barrier warp 0 adds the CTA's chunks per row (fixed order),
// publishes one partial per (CTA, row), and the last CTA to arrive for a row sums the row's
// partials (fixed order) and applies the row update.  The other warps go straight on to the
// regular tiles (warp 0 takes the last tile of every round).
__device__ __forceinline__ double* split_scratch()
{
    __shared__ double s_part[SPLIT_SLOTS];  // one instance per kernel (not per instantiation of run_phase)
    return s_part;
}

// COOP (cooperative persistent kernels only: all CTAs are co-resident): rows that span several CTAs
// are joined WITHOUT fences or atomics -- every CTA stores {partial, partial ^ tag} as one 16 B word
// (value and validity travel together; a torn or stale word fails the xor test) and the row's
// finisher CTA polls the row's words until all carry this phase's `tag`, then sums them in the same
// fixed order as the atomic join.  Other kernels use the last-arrival join (fence + atomic counter).
template <bool COOP = false, class Op>
__device__ __forceinline__ void run_phase(const DevMat& M, const MatView& V, const Op& op, double* acc,
                                          unsigned long long tag = 0ull)
{
    const double* __restrict__ vec = op.vec();
    double* s_part = split_scratch();
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;

    if (V.nsplit > 0) {
        // a publishing warp fetches its first row's descriptors and the row's own vector entries NOW, so the
        // two dependent L2 round trips overlap the chunk tiles instead of following the CTA barrier
        int4 ls_first = make_int4(0, 0, 0, 0), sr_first = make_int4(0, 0, 0, 0);
        typename Op::Pre spre_first{};
        if ((uint32_t)warp < V.nls) {
            ls_first = __ldg(reinterpret_cast<const int4*>(M.lsplits + V.ls0 + warp));
            sr_first = __ldg(reinterpret_cast<const int4*>(M.splits + ls_first.x));
            if (lane == 0) spre_first = op_prefetch(op, sr_first.x, -1);
        }
        for (uint32_t t = warp; t < V.nsplit; t += nwarps) {
            const int4 raw = *reinterpret_cast<const int4*>(V.desc + t);
            double dot = tile_dot<typename Op::Mem>(V, vec, (uint32_t)raw.x, raw.z & 0xffff, lane);
            for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(FULL, dot, o);
            if (lane == 0) s_part[raw.w] = dot;
        }
        // every warp signals that its chunks are parked; only the publishing warps wait
        const uint32_t npub = min(V.nls, (uint32_t)nwarps);
        if ((uint32_t)warp < npub) {
            asm volatile("bar.sync 1, %0;" ::"r"(blockDim.x) : "memory");
        } else {
            asm volatile("bar.arrive 1, %0;" ::"r"(blockDim.x) : "memory");
        }
        // local split row w is published by warp w (round robin if there are more rows than warps)
        for (uint32_t li = warp; li < V.nls; li += nwarps) {
            const bool firstrow = (li == (uint32_t)warp);
            const int4 ls = firstrow ? ls_first : __ldg(reinterpret_cast<const int4*>(M.lsplits + V.ls0 + li));
            const int4 sr = firstrow ? sr_first : __ldg(reinterpret_cast<const int4*>(M.splits + ls.x));
            const int first = ls.z & 0xffff, count = (ls.z >> 16) & 0xffff;
            // the CTA's chunks of this row, summed in a fixed order (lane-strided, then butterfly)
            double p = 0.0;
            for (int k = lane; k < count; k += 32) p += s_part[first + k];
            for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(FULL, p, o);
            // the row's own vector entries are needed by whoever finishes the row: fetch them now,
            // off the critical path of the join
            typename Op::Pre spre = spre_first;
            if (lane == 0 && !firstrow) spre = op_prefetch(op, sr.x, -1);
            if (sr.z == 1) {   // the whole row lives in this CTA: no global join
                if (lane == 0) op_row(op, sr.x, -1, p, spre, acc);
                continue;
            }
            if (COOP) {
                if (lane == 0) {
                    const unsigned long long a = (unsigned long long)__double_as_longlong(p);
                    unsigned long long* w = reinterpret_cast<unsigned long long*>(M.slots) + 2 * (size_t)ls.y;
                    asm volatile("st.global.cg.v2.u64 [%0], {%1, %2};" ::"l"(w), "l"(a), "l"(a ^ tag) : "memory");
                }
                if (ls.w == 0) continue;               // not the finisher of this row
                const uint32_t np = (uint32_t)sr.z;
                const unsigned long long* w0 = reinterpret_cast<const unsigned long long*>(M.slots) + 2 * (size_t)(uint32_t)sr.y;
                double s = 0.0;
                const long long t0 = clock64();
                // PW words per lane are in flight together and only the stale ones are read again: a row whose partials come
                // from more than 32 CTAs (row partition: a rank's one long row is spread over the whole grid) costs one
                // polled round trip, not one per 32 partials.  Same summation order as before: per lane ascending k, then the
                // butterfly.
                constexpr int PW = 4;
                for (uint32_t k0 = 0; k0 < np; k0 += 32 * PW) {
                    unsigned long long qa[PW], qb[PW];
#pragma unroll
                    for (int u = 0; u < PW; ++u) {
                        const uint32_t k = k0 + 32u * u + lane;
                        qa[u] = 0ull; qb[u] = tag;                     // a word that does not exist counts as arrived, value 0
                        if (k < np)
                            asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(qa[u]), "=l"(qb[u]) : "l"(w0 + 2 * (size_t)k) : "memory");
                    }
                    for (;;) {
                        bool ok = true;
#pragma unroll
                        for (int u = 0; u < PW; ++u) ok &= (qa[u] ^ qb[u]) == tag;
                        if (__all_sync(FULL, ok)) break;
                        if (clock64() - t0 > 4000000000LL) {   // never hang: flag the error and move on
                            if (lane == 0) M.partials[0] = NAN, *reinterpret_cast<volatile double*>(M.slots) = NAN;
                            break;
                        }
#pragma unroll
                        for (int u = 0; u < PW; ++u) {
                            const uint32_t k = k0 + 32u * u + lane;
                            if ((qa[u] ^ qb[u]) != tag)
                                asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(qa[u]), "=l"(qb[u]) : "l"(w0 + 2 * (size_t)k) : "memory");
                        }
                    }
#pragma unroll
                    for (int u = 0; u < PW; ++u)
                        if (k0 + 32u * u < np) s += __longlong_as_double((long long)qa[u]);
                }
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
                if (lane == 0) op_row(op, sr.x, -1, s, spre, acc);
                continue;
            }
            unsigned last = 0;
            if (lane == 0) {
                __stcg(M.partials + ls.y, p);
                __threadfence();
                const unsigned old = atomicAdd(M.counters + ls.x, 1u);
                last = (old == (unsigned)sr.z - 1u);
            }
            last = __shfl_sync(FULL, last, 0);
            if (last) {
                __threadfence();
                const uint32_t np = (uint32_t)sr.z;
                const double* pp = M.partials + (uint32_t)sr.y;
                double s = 0.0;
                for (uint32_t k0 = 0; k0 < np; k0 += 32) {   // lane-strided, ascending: the order of the polled join
                    const uint32_t k = k0 + lane;
                    s += k < np ? __ldcg(pp + k) : 0.0;
                }
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
                if (lane == 0) {
                    M.counters[ls.x] = 0;
                    op_row(op, sr.x, -1, s, spre, acc);
                }
            }
        }
    }

    // regular tiles; the warps that published split rows are served last in every round
    const uint32_t busy = min(V.nls, (uint32_t)nwarps);
    const uint32_t wslot = ((uint32_t)warp + (uint32_t)nwarps - busy) % (uint32_t)nwarps;
    for (uint32_t t = V.nsplit + wslot; t < V.ntiles; t += 2 * nwarps) {
        const uint32_t t2 = t + nwarps;
        const int4 raw = *reinterpret_cast<const int4*>(V.desc + t);
        const int nsteps = raw.z & 0xffff;
        const int logL = (raw.z >> 16) & 0xff;
        const int nrows = (raw.z >> 24) & 0xff;
        const int L = 1 << logL;
        const int rr = lane >> logL;
        const bool owner = ((lane & (L - 1)) == 0) && (rr < nrows);
        const int r = raw.y + rr;
        const int slot = op_slots<Op>::value ? raw.w + rr : -1;   // resident views hold the CTA-local row slot in .split
        typename Op::Pre pre{};
        if (owner) pre = op_prefetch(op, r, slot);
        if (t2 < V.ntiles) {
            const int4 raw2 = *reinterpret_cast<const int4*>(V.desc + t2);
            const int nsteps2 = raw2.z & 0xffff;
            const int logL2 = (raw2.z >> 16) & 0xff;
            const int nrows2 = (raw2.z >> 24) & 0xff;
            const int L2 = 1 << logL2;
            const int rr2 = lane >> logL2;
            const bool owner2 = ((lane & (L2 - 1)) == 0) && (rr2 < nrows2);
            const int r2 = raw2.y + rr2;
            const int slot2 = op_slots<Op>::value ? raw2.w + rr2 : -1;
            typename Op::Pre pre2{};
            if (owner2) pre2 = op_prefetch(op, r2, slot2);
            double dot, dot2;
            if (nsteps == nsteps2 && nsteps >= 1 && nsteps <= 3) {
                const double2 *vpa, *vpb;
                const int2 *ipa, *ipb;
                const bool ra = tile_ptrs(V, (uint32_t)raw.x, nsteps, lane, vpa, ipa);
                const bool rb = tile_ptrs(V, (uint32_t)raw2.x, nsteps2, lane, vpb, ipb);
                if (ra && rb) {   // both in shared memory: LDS with immediate offsets
                    __builtin_assume(__isShared(vpa)); __builtin_assume(__isShared(ipa));
                    __builtin_assume(__isShared(vpb)); __builtin_assume(__isShared(ipb));
                    if (nsteps == 1) pair_dot<typename Op::Mem, 1>(vpa, ipa, vpb, ipb, vec, dot, dot2);
                    else if (nsteps == 2) pair_dot<typename Op::Mem, 2>(vpa, ipa, vpb, ipb, vec, dot, dot2);
                    else pair_dot<typename Op::Mem, 3>(vpa, ipa, vpb, ipb, vec, dot, dot2);
                } else {
                    if (nsteps == 1) pair_dot<typename Op::Mem, 1>(vpa, ipa, vpb, ipb, vec, dot, dot2);
                    else if (nsteps == 2) pair_dot<typename Op::Mem, 2>(vpa, ipa, vpb, ipb, vec, dot, dot2);
                    else pair_dot<typename Op::Mem, 3>(vpa, ipa, vpb, ipb, vec, dot, dot2);
                }
            } else {
                dot = tile_dot<typename Op::Mem>(V, vec, (uint32_t)raw.x, nsteps, lane);
                dot2 = tile_dot<typename Op::Mem>(V, vec, (uint32_t)raw2.x, nsteps2, lane);
            }
            for (int o = L >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(FULL, dot, o);
            for (int o = L2 >> 1; o > 0; o >>= 1) dot2 += __shfl_xor_sync(FULL, dot2, o);
            if (owner) op_row(op, r, slot, dot, pre, acc);
            if (owner2) op_row(op, r2, slot2, dot2, pre2, acc);
        } else {
            double dot = tile_dot<typename Op::Mem>(V, vec, (uint32_t)raw.x, nsteps, lane);
            for (int o = L >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(FULL, dot, o);
            if (owner) op_row(op, r, slot, dot, pre, acc);
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
