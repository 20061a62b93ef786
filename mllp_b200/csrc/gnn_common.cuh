// gnn_common.cuh -- device helpers shared by the forward (gnn_kernels.cu) and the backward (gnn_backward.cu) of the
// bipartite message passing: parameter-block layout, per-lane online-softmax state, destination-node prologue, edge
// loop, group merge, output channels.  See gnn_kernels.cu for the formulation.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <new>
#include <string>

#include "../../include/mllp_b200.h"

namespace mllp {
void set_last_error(const std::string& msg);
void count_launch(int n);   // cabi.cu: launch statistics (mllp_launch_count)

namespace {
constexpr int C = 16;              // channels
constexpr unsigned FULLM = 0xffffffffu;
constexpr int ITEM_FLOATS = 20;    // partial state of one item: m, l, pa, pad, acc[16]

// Parameter block of one conv (floats), see conv_offsets():
//   MQ[din][din] (input-major: qt[o] += x[i] MQ[i][o]) | vq[din] | wq[din] | sq | pad to 4 |
//   Wv'[din][16] | bv[16] | Ws'[din][16] | bs[16] | We[16]
template <int DIN>
struct Off {
    static constexpr int mq = 0, vq = DIN * DIN, wq = vq + DIN, sq = wq + DIN;
    static constexpr int wv = (sq + 1 + 3) & ~3, bv = wv + DIN * C, ws = bv + C, bs = ws + DIN * C, we = bs + C, total = we + C;
};

// One lane's online-softmax state over the edges it has seen: running max m, sum l of exp(s - m), the same
// weights times the edge attribute (pa) and times the source feature rows (acc).
template <int DIN>
struct State {
    float m, l, pa, acc[DIN];
};

template <int DIN>
__device__ __forceinline__ void load_row(const float* __restrict__ p, float* r)
{
    if constexpr (DIN % 4 == 0) {
        const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
        for (int k = 0; k < DIN / 4; ++k) {
            const float4 t = __ldg(q + k);
            r[4 * k] = t.x; r[4 * k + 1] = t.y; r[4 * k + 2] = t.z; r[4 * k + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int d = 0; d < DIN; ++d) r[d] = __ldg(p + d);
    }
}

template <int DIN>
__device__ __forceinline__ void state_init(State<DIN>& st)
{
    st.m = -INFINITY; st.l = 0.0f; st.pa = 0.0f;
#pragma unroll
    for (int d = 0; d < DIN; ++d) st.acc[d] = 0.0f;
}

// fold one edge (score s, attribute a, source row x) into the lane's state
template <int DIN>
__device__ __forceinline__ void fold(State<DIN>& st, float s, float a, const float* x)
{
    if (s > st.m) {   // new maximum: rescale what was accumulated (rare after the first edges)
        const float sc = __expf(st.m - s);   // exp(-inf) = 0 on the first edge
        st.l *= sc; st.pa *= sc;
#pragma unroll
        for (int d = 0; d < DIN; ++d) st.acc[d] *= sc;
        st.m = s;
    }
    const float p = __expf(s - st.m);
    st.l += p;
    st.pa = fmaf(p, a, st.pa);
#pragma unroll
    for (int d = 0; d < DIN; ++d) st.acc[d] = fmaf(p, x[d], st.acc[d]);
}

// qt = (Wk'(Wq x + bq)) / 4 and qe = (We . (Wq x + bq)) / 4 of a destination node with feature row x.  The S lanes of the
// node's group share the din x din product: lane gl forms din / S entries of qt, a round of shuffles hands them round.
template <int S, int DIN>
__device__ __forceinline__ void dst_prologue(const float* __restrict__ prm, const float* x, int gl, float* qt, float& qe)
{
    using O = Off<DIN>;
    qe = prm[O::sq];
    if constexpr (DIN % 4 == 0) {
#pragma unroll
        for (int d4 = 0; d4 < DIN / 4; ++d4) {
            const float4 w = *reinterpret_cast<const float4*>(prm + O::wq + 4 * d4);
            qe = fmaf(x[4 * d4], w.x, qe); qe = fmaf(x[4 * d4 + 1], w.y, qe);
            qe = fmaf(x[4 * d4 + 2], w.z, qe); qe = fmaf(x[4 * d4 + 3], w.w, qe);
        }
    } else {
#pragma unroll
        for (int d = 0; d < DIN; ++d) qe = fmaf(x[d], prm[O::wq + d], qe);
    }
    if constexpr (DIN >= 4 * S && DIN % 4 == 0) {
        constexpr int PER = DIN / S;          // entries of qt per lane (a multiple of 4)
        float mine[PER];
#pragma unroll
        for (int k = 0; k < PER; ++k) mine[k] = prm[O::vq + gl * PER + k];
#pragma unroll
        for (int i = 0; i < DIN; ++i) {
#pragma unroll
            for (int k4 = 0; k4 < PER / 4; ++k4) {
                const float4 w = *reinterpret_cast<const float4*>(prm + O::mq + i * DIN + gl * PER + 4 * k4);
                mine[4 * k4] = fmaf(x[i], w.x, mine[4 * k4]); mine[4 * k4 + 1] = fmaf(x[i], w.y, mine[4 * k4 + 1]);
                mine[4 * k4 + 2] = fmaf(x[i], w.z, mine[4 * k4 + 2]); mine[4 * k4 + 3] = fmaf(x[i], w.w, mine[4 * k4 + 3]);
            }
        }
        if constexpr (S == 1) {
#pragma unroll
            for (int d = 0; d < DIN; ++d) qt[d] = mine[d];
        } else {
            const int base = (threadIdx.x & 31) & ~(S - 1);
#pragma unroll
            for (int d = 0; d < DIN; ++d) qt[d] = __shfl_sync(FULLM, mine[d % PER], base + d / PER);
        }
    } else {
#pragma unroll
        for (int d = 0; d < DIN; ++d) qt[d] = prm[O::vq + d];
        if constexpr (DIN % 4 == 0) {
#pragma unroll
            for (int i = 0; i < DIN; ++i) {
#pragma unroll
                for (int o4 = 0; o4 < DIN / 4; ++o4) {
                    const float4 w = *reinterpret_cast<const float4*>(prm + O::mq + i * DIN + 4 * o4);
                    qt[4 * o4] = fmaf(x[i], w.x, qt[4 * o4]); qt[4 * o4 + 1] = fmaf(x[i], w.y, qt[4 * o4 + 1]);
                    qt[4 * o4 + 2] = fmaf(x[i], w.z, qt[4 * o4 + 2]); qt[4 * o4 + 3] = fmaf(x[i], w.w, qt[4 * o4 + 3]);
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < DIN; ++i)
#pragma unroll
                for (int o = 0; o < DIN; ++o) qt[o] = fmaf(x[i], prm[O::mq + i * DIN + o], qt[o]);
        }
    }
}

// edges [e0, e1) of one destination node, strided over the S lanes of its group (lane `gl` of the group); two
// edges per lane are in flight together
template <int S, int DIN>
__device__ __forceinline__ void edge_loop(const int32_t* __restrict__ indices, const double* __restrict__ values,
                                          const float* __restrict__ hsrc, int e0, int e1, int gl, const float* qt, float qe,
                                          State<DIN>& st)
{
    for (int e = e0 + gl; e < e1; e += 2 * S) {
        const int eb = e + S;
        const bool two = eb < e1;
        const int ja = __ldg(indices + e);
        const int jb = two ? __ldg(indices + eb) : ja;
        const float aa = (float)__ldg(values + e);     // edge_attr = float32(a_ij), as the reference casts it
        const float ab = two ? (float)__ldg(values + eb) : 0.0f;
        float xa[DIN], xb[DIN];
        load_row<DIN>(hsrc + (size_t)ja * DIN, xa);
        load_row<DIN>(hsrc + (size_t)jb * DIN, xb);
        float sa = aa * qe, sb = ab * qe;
#pragma unroll
        for (int d = 0; d < DIN; ++d) { sa = fmaf(qt[d], xa[d], sa); sb = fmaf(qt[d], xb[d], sb); }
        fold<DIN>(st, sa, aa, xa);
        if (two) fold<DIN>(st, sb, ab, xb);
    }
}

// Merge the states of the S lanes of a group (butterfly all-reduce: every lane ends with the group's m, l, pa, acc).
template <int S, int DIN>
__device__ __forceinline__ void merge_group(State<DIN>& st)
{
    float M = st.m;
#pragma unroll
    for (int o = S / 2; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(FULLM, M, o));
    const float sc = st.m == -INFINITY ? 0.0f : __expf(st.m - M);   // a lane that saw no edge contributes nothing
    st.l *= sc; st.pa *= sc;
#pragma unroll
    for (int d = 0; d < DIN; ++d) st.acc[d] *= sc;
    st.m = M;
#pragma unroll
    for (int o = S / 2; o > 0; o >>= 1) {
        st.l += __shfl_xor_sync(FULLM, st.l, o);
        st.pa += __shfl_xor_sync(FULLM, st.pa, o);
#pragma unroll
        for (int d = 0; d < DIN; ++d) st.acc[d] += __shfl_xor_sync(FULLM, st.acc[d], o);
    }
}

// channels c0 .. c0 + CNT of the layer's output for a destination node with feature row x and merged state (l, pa, acc):
//   out_c = (Wv acc)_c / l + bv_c [l > 0] + (pa / l) We_c + (Ws x)_c + bs_c
template <int DIN, int CNT>
__device__ __forceinline__ void out_channels(const float* __restrict__ prm, int c0, const float* x, const float* acc, float pa,
                                             float inv, bool any, int relu, float* o)
{
    using O = Off<DIN>;
    float v[CNT], sk[CNT];
#pragma unroll
    for (int k = 0; k < CNT; ++k) { v[k] = 0.0f; sk[k] = prm[O::bs + c0 + k]; }
#pragma unroll
    for (int d = 0; d < DIN; ++d) {
        if constexpr (CNT == 4) {
            const float4 wv = *reinterpret_cast<const float4*>(prm + O::wv + d * C + c0);
            const float4 ws = *reinterpret_cast<const float4*>(prm + O::ws + d * C + c0);
            v[0] = fmaf(acc[d], wv.x, v[0]); v[1] = fmaf(acc[d], wv.y, v[1]); v[2] = fmaf(acc[d], wv.z, v[2]); v[3] = fmaf(acc[d], wv.w, v[3]);
            sk[0] = fmaf(x[d], ws.x, sk[0]); sk[1] = fmaf(x[d], ws.y, sk[1]); sk[2] = fmaf(x[d], ws.z, sk[2]); sk[3] = fmaf(x[d], ws.w, sk[3]);
        } else if constexpr (CNT == 2) {
            const float2 wv = *reinterpret_cast<const float2*>(prm + O::wv + d * C + c0);
            const float2 ws = *reinterpret_cast<const float2*>(prm + O::ws + d * C + c0);
            v[0] = fmaf(acc[d], wv.x, v[0]); v[1] = fmaf(acc[d], wv.y, v[1]);
            sk[0] = fmaf(x[d], ws.x, sk[0]); sk[1] = fmaf(x[d], ws.y, sk[1]);
        } else {
#pragma unroll
            for (int k = 0; k < CNT; ++k) {
                v[k] = fmaf(acc[d], prm[O::wv + d * C + c0 + k], v[k]);
                sk[k] = fmaf(x[d], prm[O::ws + d * C + c0 + k], sk[k]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < CNT; ++k) {
        float r = fmaf(pa * inv, prm[O::we + c0 + k], v[k] * inv) + (any ? prm[O::bv + c0 + k] : 0.0f) + sk[k];
        if (relu) r = fmaxf(r, 0.0f);
        o[k] = r;
    }
}

// merged state of the items [t0, t1) of a cut row: lane k folds items t0 + k, t0 + k + 32, ... in order, then the 32
// lanes' states are merged by the butterfly (a fixed order); every lane ends with the row's state
template <int DIN>
__device__ __forceinline__ void merge_items(const float* __restrict__ scratch, int t0, int t1, int lane, State<DIN>& st)
{
    state_init<DIN>(st);
#pragma unroll 2
    for (int t = t0 + lane; t < t1; t += 32) {
        const float4* o = reinterpret_cast<const float4*>(scratch + (size_t)t * ITEM_FLOATS);
        const float4 h = o[0];   // m, l, pa
        float a2[DIN];
        if constexpr (DIN % 4 == 0) {
#pragma unroll
            for (int k = 0; k < DIN / 4; ++k) {
                const float4 q = o[1 + k];
                a2[4 * k] = q.x; a2[4 * k + 1] = q.y; a2[4 * k + 2] = q.z; a2[4 * k + 3] = q.w;
            }
        } else {
#pragma unroll
            for (int d = 0; d < DIN; ++d) a2[d] = scratch[(size_t)t * ITEM_FLOATS + 4 + d];
        }
        const float mn = fmaxf(st.m, h.x);
        const float s1 = st.m == -INFINITY ? 0.0f : __expf(st.m - mn), s2 = h.x == -INFINITY ? 0.0f : __expf(h.x - mn);
        st.l = st.l * s1 + h.y * s2;
        st.pa = st.pa * s1 + h.z * s2;
#pragma unroll
        for (int d = 0; d < DIN; ++d) st.acc[d] = st.acc[d] * s1 + a2[d] * s2;
        st.m = mn;
    }
    merge_group<32, DIN>(st);
}

// long rows: item t = (row, first edge, end edge); one warp per item, merged partial state -> scratch[t][20]
template <int DIN>
__global__ void __launch_bounds__(256) k_gnn_conv_items(int nitems, const int32_t* __restrict__ items, const int32_t* __restrict__ indices,
                                                        const double* __restrict__ values, const float* __restrict__ hdst,
                                                        const float* __restrict__ hsrc, const float* __restrict__ prm_g,
                                                        float* __restrict__ scratch)
{
    using O = Off<DIN>;
    __shared__ __align__(16) float prm[O::wv];   // the prologue's part of the block
    for (int k = threadIdx.x; k < O::wv; k += blockDim.x) prm[k] = prm_g[k];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < nitems; t += warps) {
        const int i = __ldg(items + 3 * t), e0 = __ldg(items + 3 * t + 1), e1 = __ldg(items + 3 * t + 2);
        float x[DIN], qt[DIN], qe;
        load_row<DIN>(hdst + (size_t)i * DIN, x);
        dst_prologue<32, DIN>(prm, x, lane, qt, qe);
        State<DIN> st;
        state_init<DIN>(st);
        edge_loop<32, DIN>(indices, values, hsrc, e0, e1, lane, qt, qe, st);
        merge_group<32, DIN>(st);
        if (lane == 0) {
            float4* o = reinterpret_cast<float4*>(scratch + (size_t)t * ITEM_FLOATS);
            o[0] = make_float4(st.m, st.l, st.pa, 0.0f);
            float a[C];
#pragma unroll
            for (int d = 0; d < C; ++d) a[d] = d < DIN ? st.acc[d < DIN ? d : 0] : 0.0f;
#pragma unroll
            for (int k = 0; k < 4; ++k) o[1 + k] = make_float4(a[4 * k], a[4 * k + 1], a[4 * k + 2], a[4 * k + 3]);
        }
    }
}


// host helpers
int gfail(int code, const std::string& msg) { set_last_error(msg); return code; }
int grid_for_warps(long long warps_needed)
{
    const long long blocks = (warps_needed + 7) / 8;
    return (int)(blocks < 1 ? 1 : blocks > 148 * 8 ? 148 * 8 : blocks);   // 8 CTAs of 256 threads per SM
}
int cuda_status(const char* what)
{
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return gfail((int)e, std::string(what) + ": " + cudaGetErrorString(e));
    return 0;
}

bool side_ok(const mllp_gnn_side* g)
{
    if (!g || g->nd < 0 || g->ns < 0 || !g->indptr) return false;
    if (g->group != 1 && g->group != 2 && g->group != 4 && g->group != 8 && g->group != 16 && g->group != 32) return false;
    if (g->chunk < 16) return false;
    if (g->nlong > 0 && (!g->long_rows || !g->long_first || !g->items || !g->scratch || g->nitems < g->nlong)) return false;
    return true;
}

}  // namespace
}  // namespace mllp

// A captured launch sequence (forward: mllp_gnn_plan_create, backward: mllp_gnn_backward_plan_create)
struct mllp_gnn_plan {
    cudaGraphExec_t exec = nullptr;
    int launches = 0;
};

namespace mllp {
namespace {
// Capture what `body(stream, second stream, events[8])` launches into a CUDA graph and instantiate it.
template <class F>
int capture_plan(const char* who, mllp_gnn_plan_t* out, F body)
{
    cudaStream_t s = nullptr, s2 = nullptr;
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaGraph_t graph = nullptr;
    mllp_gnn_plan* plan = new (std::nothrow) mllp_gnn_plan();
    if (!plan) return gfail(MLLP_E_NOMEM, std::string(who) + ": out of host memory");
    auto cleanup = [&]() {
        for (cudaEvent_t e : ev) if (e) cudaEventDestroy(e);
        if (graph) cudaGraphDestroy(graph);
        if (s2) cudaStreamDestroy(s2);
        if (s) cudaStreamDestroy(s);
    };
    cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    for (int k = 0; k < 8 && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) { cleanup(); delete plan; return gfail((int)e, std::string(who) + ": " + cudaGetErrorString(e)); }
    int rc = body(s, s2, ev);
    e = cudaStreamEndCapture(s, &graph);
    if (rc == 0 && e != cudaSuccess) rc = gfail((int)e, std::string(who) + ": capture: " + cudaGetErrorString(e));
    if (rc == 0) {
        size_t nodes = 0;
        cudaGraphGetNodes(graph, nullptr, &nodes);
        plan->launches = (int)nodes;
        e = cudaGraphInstantiate(&plan->exec, graph, 0);
        if (e != cudaSuccess) rc = gfail((int)e, std::string(who) + ": instantiate: " + cudaGetErrorString(e));
    }
    cleanup();
    if (rc != 0) { cudaGetLastError(); delete plan; return rc; }
    *out = plan;
    return 0;
}
}  // namespace
}  // namespace mllp
