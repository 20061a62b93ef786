// gnn_kernels.cu -- bipartite message passing over the LP's nonzeros, sm_100a.
//
// What it replaces: the forward pass of the reference's GNNModel (linear_program_methods.py:238-251): five
// torch_geometric TransformerConv layers (heads = 1, 16 channels, edge_dim = 1, root weight, bias; :199-204) that
// alternate constraint->variable ("w2s", along the rows of A') and variable->constraint ("s2w", along the rows of
// A) passes over the same sparsity as the PDHG products, ReLU between them and a final Linear(16, 1) (:215, :250).
// torch_geometric is not vendored (nor installed here); the layer is restated from its published definition:
//
//     q_i = Wq x_i + bq,  k_j = Wk x_j + bk,  v_j = Wv x_j + bv,  e_ij = We a_ij            (We has no bias)
//     alpha_ij = softmax_j( q_i . (k_j + e_ij) / sqrt(16) )   over the incoming edges j -> i of node i
//     out_i = sum_j alpha_ij (v_j + e_ij) + Ws x_i + bs
//
// fp32, as the reference (dtype=torch.float, :90-91, :100).  Structure:
//   * the four 16-wide linear maps of a layer are node-wise and run first (k_gnn_project*): {q | skip} rows for the
//     destination nodes, {k | v} rows for the source nodes, 128 B per node;
//   * the conv kernel gives every destination node (a CSR row) a group of S lanes (S = 4 .. 32, chosen from the
//     mean row length like the lanes-per-row of the LP format); a lane owns every S-th edge of the row: it gathers
//     the 64 B key row of the edge's source node, scores it against q, and folds the 64 B value row into its own
//     online-softmax state (running max, sum, 16 accumulators).  Nothing crosses lanes inside the edge loop and
//     nothing of size nnz is written; at the row end the group's states are merged (max butterfly, rescale,
//     reduce-scatter of the 16 channels), so every edge costs ~2 warp instructions instead of a warp-wide reduction;
//   * rows longer than `chunk` edges (osa-60 has rows of 173 366 edges) are cut into items (one warp each) whose
//     partial states are merged in a fixed order by a second kernel.
// HBM/L2-bound gather work (hidden = 16): no tensor cores.
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "../../include/mllp_b200.h"

namespace mllp {
void set_last_error(const std::string& msg);

namespace {
constexpr int C = 16;              // channels
constexpr unsigned FULLM = 0xffffffffu;
constexpr int ITEM_FLOATS = 20;    // partial state of one item: m, l, pa, pad, acc[16]

// One lane's online-softmax state over the edges it has seen: running max m, sum l of exp(s - m), the same
// weights times the edge attribute (pa) and times the value rows (acc).
struct State {
    float m, l, pa, acc[C];
};

__device__ __forceinline__ void load16(const float* __restrict__ p, float* r)
{
    const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float4 t = __ldg(q + k);
        r[4 * k] = t.x; r[4 * k + 1] = t.y; r[4 * k + 2] = t.z; r[4 * k + 3] = t.w;
    }
}

// fold one edge (score s, attribute a, value row v) into the lane's state
__device__ __forceinline__ void fold(State& st, float s, float a, const float* v)
{
    if (s > st.m) {   // new maximum: rescale what was accumulated (rare after the first edges)
        const float sc = __expf(st.m - s);   // exp(-inf) = 0 on the first edge
        st.l *= sc; st.pa *= sc;
#pragma unroll
        for (int c = 0; c < C; ++c) st.acc[c] *= sc;
        st.m = s;
    }
    const float p = __expf(s - st.m);
    st.l += p;
    st.pa = fmaf(p, a, st.pa);
#pragma unroll
    for (int c = 0; c < C; ++c) st.acc[c] = fmaf(p, v[c], st.acc[c]);
}

// edges [e0, e1) of one destination node, strided over the S lanes of its group (lane `gl` of the group); two
// edges per lane are in flight together
template <int S>
__device__ __forceinline__ void edge_loop(const int32_t* __restrict__ indices, const double* __restrict__ values,
                                          const float* __restrict__ kv, int e0, int e1, int gl, const float* q, float qe, State& st)
{
    for (int e = e0 + gl; e < e1; e += 2 * S) {
        const int eb = e + S;
        const bool two = eb < e1;
        const int ja = __ldg(indices + e);
        const int jb = two ? __ldg(indices + eb) : ja;
        const float aa = (float)__ldg(values + e);     // edge_attr = float32(a_ij), as the reference casts it
        const float ab = two ? (float)__ldg(values + eb) : 0.0f;
        float ka[C], kb[C], va[C], vb[C];
        load16(kv + (size_t)ja * 32, ka);
        load16(kv + (size_t)jb * 32, kb);
        load16(kv + (size_t)ja * 32 + C, va);
        load16(kv + (size_t)jb * 32 + C, vb);
        float sa = aa * qe, sb = ab * qe;
#pragma unroll
        for (int c = 0; c < C; ++c) { sa = fmaf(q[c], ka[c], sa); sb = fmaf(q[c], kb[c], sb); }
        fold(st, sa * 0.25f, aa, va);   // / sqrt(16)
        if (two) fold(st, sb * 0.25f, ab, vb);
    }
}

// Merge the states of the S lanes of a group.  On return lane `gl` holds, in acc[0 .. n), the group's sums of the
// n = max(16 / S, 1) channels starting at `cbase`; l and pa are the group's sums in every lane.
template <int S>
__device__ __forceinline__ int merge_group(State& st, int gl, int& cbase)
{
    float M = st.m;
#pragma unroll
    for (int o = S / 2; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(FULLM, M, o));
    const float sc = st.m == -INFINITY ? 0.0f : __expf(st.m - M);   // a lane that saw no edge contributes nothing
    st.l *= sc; st.pa *= sc;
#pragma unroll
    for (int c = 0; c < C; ++c) st.acc[c] *= sc;
    st.m = M;
#pragma unroll
    for (int o = S / 2; o > 0; o >>= 1) {
        st.l += __shfl_xor_sync(FULLM, st.l, o);
        st.pa += __shfl_xor_sync(FULLM, st.pa, o);
    }
    // reduce-scatter of the 16 channels: at every level a lane keeps one half of its channels and receives the
    // partner's sums of that half
    cbase = 0;
    int cnt = C;
#pragma unroll
    for (int o = S / 2; o > 0; o >>= 1) {
        if (cnt > 1) {
            const int half = cnt / 2;
            const bool upper = (gl & o) != 0;
#pragma unroll
            for (int k = 0; k < C / 2; ++k) {
                if (k < half) {
                    const float send = upper ? st.acc[k] : st.acc[k + half];
                    const float keep = upper ? st.acc[k + half] : st.acc[k];
                    st.acc[k] = keep + __shfl_xor_sync(FULLM, send, o);
                }
            }
            cbase += upper ? half : 0;
            cnt = half;
        } else {
            st.acc[0] += __shfl_xor_sync(FULLM, st.acc[0], o);   // S = 32: both lanes of a pair hold the same channel
        }
    }
    return cnt;
}

// out_i[c] = acc_c / l + (pa / l) We_c + skip_c  for this lane's channels
template <int S>
__device__ __forceinline__ void row_epilogue(float* __restrict__ hout, const float* __restrict__ qs, const float* we, int i, int gl,
                                             const State& st, int cbase, int cnt, int relu)
{
    if (S == 32 && (gl & 1)) return;   // the odd lane of a pair holds a copy
    const float inv = st.l > 0.0f ? 1.0f / st.l : 0.0f;   // a node without incoming edges keeps only the root term
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < cnt) {
            const int c = cbase + k;
            float o = fmaf(st.pa * inv, we[c], st.acc[k] * inv) + __ldg(qs + (size_t)i * 32 + C + c);
            if (relu) o = fmaxf(o, 0.0f);
            hout[(size_t)i * C + c] = o;
        }
    }
}

__device__ __forceinline__ void state_init(State& st)
{
    st.m = -INFINITY; st.l = 0.0f; st.pa = 0.0f;
#pragma unroll
    for (int c = 0; c < C; ++c) st.acc[c] = 0.0f;
}

// rows with at most `chunk` edges: S lanes per row
template <int S>
__global__ void __launch_bounds__(256) k_gnn_conv_rows(int nd, const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                                       const double* __restrict__ values, const float* __restrict__ qs,
                                                       const float* __restrict__ kv, const float* __restrict__ we_g,
                                                       float* __restrict__ hout, int chunk, int relu)
{
    __shared__ float we[C];
    if (threadIdx.x < C) we[threadIdx.x] = we_g[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, gl = lane & (S - 1);
    constexpr int RPW = 32 / S;   // rows per warp
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * RPW; base < nd; base += warps * RPW) {
        const int i = base + lane / S;   // warp-uniform trip count: the merges below are warp-wide
        int e0 = 0, e1 = 0;
        if (i < nd) { e0 = __ldg(indptr + i); e1 = __ldg(indptr + i + 1); }
        const bool live = i < nd && e1 - e0 <= chunk;   // long row: k_gnn_conv_items + k_gnn_conv_merge
        State st;
        state_init(st);
        if (live && e1 > e0) {
            float q[C];
            load16(qs + (size_t)i * 32, q);
            float qe = 0.0f;
#pragma unroll
            for (int c = 0; c < C; ++c) qe = fmaf(q[c], we[c], qe);
            edge_loop<S>(indices, values, kv, e0, e1, gl, q, qe, st);
        }
        int cbase;
        const int cnt = merge_group<S>(st, gl, cbase);
        if (live) row_epilogue<S>(hout, qs, we, i, gl, st, cbase, cnt, relu);
    }
}

// long rows: item t = (row, first edge, end edge); one warp per item, merged partial state -> scratch[t][20]
__global__ void __launch_bounds__(256) k_gnn_conv_items(int nitems, const int32_t* __restrict__ items, const int32_t* __restrict__ indices,
                                                        const double* __restrict__ values, const float* __restrict__ qs,
                                                        const float* __restrict__ kv, const float* __restrict__ we_g,
                                                        float* __restrict__ scratch)
{
    __shared__ float we[C];
    if (threadIdx.x < C) we[threadIdx.x] = we_g[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < nitems; t += warps) {
        const int i = __ldg(items + 3 * t), e0 = __ldg(items + 3 * t + 1), e1 = __ldg(items + 3 * t + 2);
        float q[C];
        load16(qs + (size_t)i * 32, q);
        float qe = 0.0f;
#pragma unroll
        for (int c = 0; c < C; ++c) qe = fmaf(q[c], we[c], qe);
        State st;
        state_init(st);
        edge_loop<32>(indices, values, kv, e0, e1, lane, q, qe, st);
        int cbase;
        merge_group<32>(st, lane, cbase);
        float* o = scratch + (size_t)t * ITEM_FLOATS;
        if (lane == 0) { o[0] = st.m; o[1] = st.l; o[2] = st.pa; o[3] = 0.0f; }
        if (!(lane & 1)) o[4 + cbase] = st.acc[0];
    }
}

// long rows: merge the partial states of row r's items [first[r], first[r+1]) in order (lane c < 16 owns channel c),
// then the epilogue
__global__ void __launch_bounds__(256) k_gnn_conv_merge(int nlong, const int32_t* __restrict__ long_rows,
                                                        const int32_t* __restrict__ first, const float* __restrict__ scratch,
                                                        const float* __restrict__ qs, const float* __restrict__ we_g,
                                                        float* __restrict__ hout, int relu)
{
    const int lane = threadIdx.x & 31, c = lane & 15;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < nlong; r += warps) {
        const int i = __ldg(long_rows + r);
        float m = -INFINITY, l = 0.0f, pa = 0.0f, acc = 0.0f;
        for (int t = __ldg(first + r); t < __ldg(first + r + 1); ++t) {
            const float* o = scratch + (size_t)t * ITEM_FLOATS;
            const float m2 = o[0], l2 = o[1], pa2 = o[2], a2 = o[4 + c];
            const float mn = fmaxf(m, m2);
            const float s1 = m == -INFINITY ? 0.0f : __expf(m - mn), s2 = m2 == -INFINITY ? 0.0f : __expf(m2 - mn);
            l = l * s1 + l2 * s2;
            pa = pa * s1 + pa2 * s2;
            acc = acc * s1 + a2 * s2;
            m = mn;
        }
        const float inv = l > 0.0f ? 1.0f / l : 0.0f;
        float o = fmaf(pa * inv, __ldg(we_g + c), acc * inv) + __ldg(qs + (size_t)i * 32 + C + c);
        if (relu) o = fmaxf(o, 0.0f);
        if (lane < 16) hout[(size_t)i * C + c] = o;
    }
}

// Node-wise linear maps: out[j][0..15] = W1 h_j + b1, out[j][16..31] = W2 h_j + b2, with
// params = W1'[din][16] | b1[16] | W2'[din][16] | b2[16] (W' = transposed weight: input-major).
// A node set can feed two pairs at once (its {q | skip} rows as destination of one conv and its {k | v} rows as
// source of the other conv of the layer), and both node sets of the graph share the launch.
// k_gnn_project<DIN> (the model's widths, 1 and 16): one THREAD per (node, pair) -- the node's row in registers, the 32
// outputs in registers, weights read as broadcast LDS.128 (one per 4 FMAs).  k_gnn_project_generic (any din <= 32, unit
// parity only): one warp per node, lane l < 16 channel l of the first map, lane l >= 16 channel l - 16 of the second.
// Both accumulate bias first, then inputs in ascending order: bitwise equal.
struct ProjJob {
    const float* h;        // [n][din]
    const float* pa;       // first pair of maps, or null
    float* oa;             // [n][32]
    const float* pb;       // second pair, or null
    float* ob;
    int n;
};

__global__ void __launch_bounds__(256) k_gnn_project_generic(ProjJob j0, ProjJob j1, int din)
{
    extern __shared__ float sp[];   // 4 parameter blocks of np floats
    const int np = 2 * din * C + 2 * C;
    const float* src[4] = {j0.pa, j0.pb, j1.pa, j1.pb};
    for (int b = 0; b < 4; ++b)
        if (src[b])
            for (int k = threadIdx.x; k < np; k += blockDim.x) sp[b * np + k] = src[b][k];
    __syncthreads();
    const int lane = threadIdx.x & 31, c = lane & 15;
    const int woff = lane < 16 ? 0 : din * C + C;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int total = j0.n + j1.n;
    // the node's feature row is fetched by ONE coalesced load (lane d holds h[d], din <= 32) and handed round by
    // shuffles; the next node's row is in flight while this one is multiplied
    int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    auto fetch = [&](int ww) -> float {
        if (ww >= total || lane >= din) return 0.0f;
        const bool sec = ww >= j0.n;
        return __ldg((sec ? j1.h : j0.h) + (size_t)(sec ? ww - j0.n : ww) * din + lane);
    };
    float hrow = fetch(w);
    for (; w < total; w += warps) {
        const float hnext = fetch(w + warps);
        const bool second = w >= j0.n;
        const int j = second ? w - j0.n : w;
        const float* pA = sp + (second ? 2 : 0) * np + woff;
        const float* pB = pA + np;
        const bool ha = (second ? j1.pa : j0.pa) != nullptr, hb = (second ? j1.pb : j0.pb) != nullptr;
        float oa = pA[din * C + c], ob = pB[din * C + c];
        for (int d = 0; d < din; ++d) {
            const float h = __shfl_sync(FULLM, hrow, d);
            oa = fmaf(h, pA[d * C + c], oa);
            ob = fmaf(h, pB[d * C + c], ob);
        }
        if (ha) (second ? j1.oa : j0.oa)[(size_t)j * 32 + lane] = oa;
        if (hb) (second ? j1.ob : j0.ob)[(size_t)j * 32 + lane] = ob;
        hrow = hnext;
    }
}

template <int DIN>
__global__ void __launch_bounds__(128) k_gnn_project(ProjJob j0, ProjJob j1)
{
    extern __shared__ __align__(16) float sp[];   // 4 parameter blocks of NP floats
    constexpr int NP = 2 * DIN * C + 2 * C;
    const float* src[4] = {j0.pa, j0.pb, j1.pa, j1.pb};
    for (int b = 0; b < 4; ++b)
        if (src[b])
            for (int k = threadIdx.x; k < NP; k += blockDim.x) sp[b * NP + k] = src[b][k];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int nb0 = (j0.n + 31) >> 5, nb1 = (j1.n + 31) >> 5;
    const int units = 2 * (nb0 + nb1);   // unit = (node set, block of 32 nodes, pair of maps): one warp, a node per lane
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < units; u += warps) {
        const bool second = u >= 2 * nb0;
        const int uu = second ? u - 2 * nb0 : u;
        const int pair = uu & 1;
        const int j = (uu >> 1) * 32 + lane;
        const float* prm = second ? (pair ? j1.pb : j1.pa) : (pair ? j0.pb : j0.pa);
        const int n = second ? j1.n : j0.n;
        if (!prm || j >= n) continue;
        float* out = (second ? (pair ? j1.ob : j1.oa) : (pair ? j0.ob : j0.oa)) + (size_t)j * 32;
        const float* hp = (second ? j1.h : j0.h) + (size_t)j * DIN;
        const float* w = sp + ((second ? 2 : 0) + pair) * NP;
        float h[DIN];
        if constexpr (DIN % 4 == 0) {
#pragma unroll
            for (int k = 0; k < DIN / 4; ++k) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(hp) + k);
                h[4 * k] = t.x; h[4 * k + 1] = t.y; h[4 * k + 2] = t.z; h[4 * k + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int d = 0; d < DIN; ++d) h[d] = __ldg(hp + d);
        }
        float acc[2 * C];
#pragma unroll
        for (int c = 0; c < C; ++c) { acc[c] = w[DIN * C + c]; acc[C + c] = w[2 * DIN * C + C + c]; }
#pragma unroll
        for (int d = 0; d < DIN; ++d) {
#pragma unroll
            for (int c4 = 0; c4 < C / 4; ++c4) {
                const float4 wa = *reinterpret_cast<const float4*>(w + d * C + 4 * c4);
                const float4 wb = *reinterpret_cast<const float4*>(w + DIN * C + C + d * C + 4 * c4);
                acc[4 * c4] = fmaf(h[d], wa.x, acc[4 * c4]); acc[4 * c4 + 1] = fmaf(h[d], wa.y, acc[4 * c4 + 1]);
                acc[4 * c4 + 2] = fmaf(h[d], wa.z, acc[4 * c4 + 2]); acc[4 * c4 + 3] = fmaf(h[d], wa.w, acc[4 * c4 + 3]);
                acc[C + 4 * c4] = fmaf(h[d], wb.x, acc[C + 4 * c4]); acc[C + 4 * c4 + 1] = fmaf(h[d], wb.y, acc[C + 4 * c4 + 1]);
                acc[C + 4 * c4 + 2] = fmaf(h[d], wb.z, acc[C + 4 * c4 + 2]); acc[C + 4 * c4 + 3] = fmaf(h[d], wb.w, acc[C + 4 * c4 + 3]);
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
            reinterpret_cast<float4*>(out)[k] = make_float4(acc[4 * k], acc[4 * k + 1], acc[4 * k + 2], acc[4 * k + 3]);
    }
}

// out[i] = w . h_i + b     (the model's final Linear(16, 1))
__global__ void __launch_bounds__(256) k_gnn_fc(int n, const float* __restrict__ h, const float* __restrict__ wb, float* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4* r = reinterpret_cast<const float4*>(h + (size_t)i * C);
    float o = __ldg(wb + C);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float4 v = __ldg(r + k);
        o = fmaf(v.x, __ldg(wb + 4 * k), o); o = fmaf(v.y, __ldg(wb + 4 * k + 1), o);
        o = fmaf(v.z, __ldg(wb + 4 * k + 2), o); o = fmaf(v.w, __ldg(wb + 4 * k + 3), o);
    }
    out[i] = o;
}

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

int launch_project(const ProjJob& j0, const ProjJob& j1, int din, cudaStream_t s)
{
    if (j0.n + j1.n <= 0) return 0;
    const size_t smem = 4 * (size_t)(2 * din * C + 2 * C) * sizeof(float);
    if (din == 1 || din == C) {
        const long long units = 2 * ((long long)((j0.n + 31) >> 5) + (long long)((j1.n + 31) >> 5));
        const long long blocks = (units + 3) / 4;
        const int grid = (int)(blocks < 1 ? 1 : blocks > 148 * 16 ? 148 * 16 : blocks);   // 16 CTAs of 128 threads per SM
        if (din == 1) k_gnn_project<1><<<grid, 128, smem, s>>>(j0, j1);
        else k_gnn_project<C><<<grid, 128, smem, s>>>(j0, j1);
    } else {
        k_gnn_project_generic<<<grid_for_warps(j0.n + j1.n), 256, smem, s>>>(j0, j1, din);
    }
    return cuda_status("mllp_gnn: projection");
}

bool side_ok(const mllp_gnn_side* g)
{
    if (!g || g->nd < 0 || g->ns < 0 || !g->indptr) return false;
    if (g->group != 4 && g->group != 8 && g->group != 16 && g->group != 32) return false;
    if (g->chunk < 32) return false;
    if (g->nlong > 0 && (!g->long_rows || !g->long_first || !g->items || !g->scratch || g->nitems < g->nlong)) return false;
    return true;
}

int launch_conv(const mllp_gnn_side& g, const float* qs, const float* kv, const float* we, float* hout, int relu, cudaStream_t s)
{
    if (g.nd == 0) return 0;
    const int grid = grid_for_warps(((long long)g.nd * g.group + 31) / 32);
    switch (g.group) {
        case 4: k_gnn_conv_rows<4><<<grid, 256, 0, s>>>(g.nd, g.indptr, g.indices, g.values, qs, kv, we, hout, g.chunk, relu); break;
        case 8: k_gnn_conv_rows<8><<<grid, 256, 0, s>>>(g.nd, g.indptr, g.indices, g.values, qs, kv, we, hout, g.chunk, relu); break;
        case 16: k_gnn_conv_rows<16><<<grid, 256, 0, s>>>(g.nd, g.indptr, g.indices, g.values, qs, kv, we, hout, g.chunk, relu); break;
        default: k_gnn_conv_rows<32><<<grid, 256, 0, s>>>(g.nd, g.indptr, g.indices, g.values, qs, kv, we, hout, g.chunk, relu); break;
    }
    if (g.nlong > 0) {
        k_gnn_conv_items<<<grid_for_warps(g.nitems), 256, 0, s>>>(g.nitems, g.items, g.indices, g.values, qs, kv, we, g.scratch);
        k_gnn_conv_merge<<<grid_for_warps(g.nlong), 256, 0, s>>>(g.nlong, g.long_rows, g.long_first, g.scratch, qs, we, hout, relu);
    }
    return cuda_status("mllp_gnn: conv");
}

// floats of one conv's parameter block: Wq'|bq|Ws'|bs | Wk'|bk|Wv'|bv | We
size_t conv_block(int din) { return 2 * (size_t)(2 * din * C + 2 * C) + C; }
}  // namespace
}  // namespace mllp

using namespace mllp;

extern "C" {

int mllp_gnn_project(int32_t n, const float* d_h, int32_t din, const float* d_params, float* d_out, void* stream)
{
    if (n < 0 || din < 1 || din > 32 || !d_h || !d_params || !d_out) return gfail(MLLP_E_INVALID, "mllp_gnn_project: bad argument (din must be 1 .. 32)");
    const ProjJob j0{d_h, d_params, d_out, nullptr, nullptr, n}, j1{nullptr, nullptr, nullptr, nullptr, nullptr, 0};
    return launch_project(j0, j1, din, (cudaStream_t)stream);
}

int mllp_gnn_conv(const mllp_gnn_side* side, const float* d_qs_dst, const float* d_kv_src, const float* d_we, float* d_hout,
                  int32_t relu, void* stream)
{
    if (!side_ok(side) || !d_qs_dst || !d_kv_src || !d_we || !d_hout) return gfail(MLLP_E_INVALID, "mllp_gnn_conv: bad argument");
    return launch_conv(*side, d_qs_dst, d_kv_src, d_we, d_hout, relu, (cudaStream_t)stream);
}

int mllp_gnn_fc(int32_t n, const float* d_h, const float* d_wb, float* d_out, void* stream)
{
    if (n < 0 || !d_h || !d_wb || !d_out) return gfail(MLLP_E_INVALID, "mllp_gnn_fc: bad argument");
    if (n == 0) return 0;
    k_gnn_fc<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, d_h, d_wb, d_out);
    return cuda_status("mllp_gnn_fc");
}

int64_t mllp_gnn_workspace_floats(int32_t n, int32_t m) { return 96 * ((int64_t)n + (int64_t)m) + 64; }

int mllp_gnn_forward(const mllp_gnn_side* to_var, const mllp_gnn_side* to_con, const float* d_x1, const float* d_x2,
                     const float* d_params, float* d_work, float* d_out, void* stream)
{
    if (!side_ok(to_var) || !side_ok(to_con) || !d_x1 || !d_x2 || !d_params || !d_work || !d_out)
        return gfail(MLLP_E_INVALID, "mllp_gnn_forward: bad argument");
    const int n = to_var->nd, m = to_con->nd;
    if (to_var->ns != m || to_con->ns != n) return gfail(MLLP_E_INVALID, "mllp_gnn_forward: the two sides do not describe one graph");
    cudaStream_t s = (cudaStream_t)stream;
    // workspace: two feature buffers and {q | skip}, {k | v} rows per node set
    float* h1[2] = {d_work, d_work + (size_t)16 * n};
    float* qs1 = d_work + (size_t)32 * n;
    float* kv1 = qs1 + (size_t)32 * n;
    float* base2 = d_work + (size_t)96 * n;
    float* h2[2] = {base2, base2 + (size_t)16 * m};
    float* qs2 = base2 + (size_t)32 * m;
    float* kv2 = qs2 + (size_t)32 * m;
    // parameter blocks: gconv1_w2s, gconv1_s2w (din 1), gconv2_w2s, gconv2_s2w, gconv3_w2s (din 16), fc
    const float* P[5];
    size_t off = 0;
    for (int k = 0; k < 5; ++k) { P[k] = d_params + off; off += conv_block(k < 2 ? 1 : C); }
    const float* fc = d_params + off;
    auto dstp = [&](int k) { return P[k]; };
    auto srcp = [&](int k) { return P[k] + (2 * (k < 2 ? 1 : C) * C + 2 * C); };
    auto wep = [&](int k) { return P[k] + 2 * (2 * (k < 2 ? 1 : C) * C + 2 * C); };

    const float* x1 = d_x1;
    const float* x2 = d_x2;
    int rc = 0;
    for (int layer = 0; layer < 2 && rc == 0; ++layer) {
        const int kw = 2 * layer, ks = 2 * layer + 1, din = layer == 0 ? 1 : C;
        // variables: destination of the w2s conv, source of the s2w conv; constraints: the other way round
        const ProjJob jv{x1, dstp(kw), qs1, srcp(ks), kv1, n}, jc{x2, dstp(ks), qs2, srcp(kw), kv2, m};
        rc = launch_project(jv, jc, din, s);
        if (rc == 0) rc = launch_conv(*to_var, qs1, kv2, wep(kw), h1[layer], 1, s);
        if (rc == 0) rc = launch_conv(*to_con, qs2, kv1, wep(ks), h2[layer], 1, s);
        x1 = h1[layer]; x2 = h2[layer];
    }
    if (rc == 0) {
        const ProjJob jv{x1, dstp(4), qs1, nullptr, nullptr, n}, jc{x2, nullptr, nullptr, srcp(4), kv2, m};
        rc = launch_project(jv, jc, C, s);
    }
    if (rc == 0) rc = launch_conv(*to_var, qs1, kv2, wep(4), h1[0], 1, s);
    if (rc == 0 && n > 0) {
        k_gnn_fc<<<(n + 255) / 256, 256, 0, s>>>(n, h1[0], fc, d_out);
        rc = cuda_status("mllp_gnn_forward: fc");
    }
    return rc;
}

}  // extern "C"
