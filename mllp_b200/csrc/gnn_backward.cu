// gnn_backward.cu -- backward pass of the bipartite message passing (gnn_kernels.cu), sm_100a.
//
// What it replaces: loss.backward() through the reference's GNNModel in its training loop
// (linear_program_experiment.py:115-157: model(graph) -> BCEWithLogitsLoss -> backward -> Adam step), i.e. autograd
// through five torch_geometric TransformerConv layers.  Like the forward it never materialises q, k, v or anything of
// size nnz: with the fused parameter block of a conv (gnn_common.cuh: MQ = Wq'Wk/4, vq, wq, sq, Wv', bv, Ws', bs, We)
//
//     s_ij = qt_i . x_j + a_ij qe_i,   alpha_ij = exp(s_ij - lse_i),   xbar_i = sum_j alpha_ij x_j,  abar_i = sum_j alpha_ij a_ij
//     out_i = Wv xbar_i + bv [i has edges] + abar_i We + Ws x_i + bs
//
// and g_i = d loss / d out_i (the ReLU mask is recomputed from out_i, activations are not stored twice):
//
//   destination pass (rows of the conv's own CSR, S lanes per row as in the forward; rows above `chunk` edges get a CTA):
//     softmax statistics again (online, one sweep), dxbar_i = Wv' g_i, dabar_i = We . g_i, D_i = dxbar_i . xbar_i + dabar_i abar_i,
//     second sweep: ds_ij = alpha_ij (dxbar_i . x_j + dabar_i a_ij - D_i),  dqt_i = sum_j ds_ij x_j,  dqe_i = sum_j ds_ij a_ij,
//     d x_i = Ws' g_i + MQ dqt_i + wq dqe_i;  leaves a 36-float record {qt_i, dxbar_i, qe_i, lse_i, dabar_i, D_i} per node;
//   source pass (rows of the TRANSPOSED structure = the other direction's CSR, which the graph holds anyway): every source
//     node j walks its edges, gathers the destination records and owns its gradient row
//     d x_j = sum_i alpha_ij dxbar_i + ds_ij qt_i  -- no scatter, no atomics, fixed summation order;
//   parameter gradients: sums over the nodes of outer products of per-node vectors ({x, xbar, abar, any, 1} x {dqt, dqe, g}),
//     per-CTA partials in registers, then a fixed-order sum over the CTAs (deterministic);
//   chain rule through the fused block back to the module's tensors (k_gnn_unpack_grads), the fused block itself is
//     formed on the device from the flat parameter vector (k_gnn_pack) so an optimiser step needs no host round trip.
//
// fp32 like the forward.  L2/HBM-bound gathers (hidden = 16): no tensor cores.
#include <mutex>
#include <unordered_map>

#include "gnn_common.cuh"

namespace mllp {
namespace {

// record of a destination node for the source pass
template <int DIN>
struct Rec {
    static constexpr int qt = 0, dxb = DIN, qe = 2 * DIN, lse = qe + 1, dab = qe + 2, dd = qe + 3, total = 2 * DIN + 4;
};
// per-node vectors for the parameter gradients
template <int DIN>
struct NR {
    static constexpr int g = 0, xbar = C, dqt = C + DIN, abar = C + 2 * DIN, dqe = abar + 1, any = abar + 2, total = (any + 1 + 3) & ~3;   // g first: float4 stores
};
// flat parameter vector of one conv, torch_geometric's registration order:
//   lin_key.weight[16][din] | lin_key.bias[16] | lin_query.weight | lin_query.bias | lin_value.weight | lin_value.bias |
//   lin_edge.weight[16] | lin_skip.weight | lin_skip.bias
template <int DIN>
struct Flat {
    static constexpr int kw = 0, kb = C * DIN, qw = kb + C, qb = qw + C * DIN, vw = qb + C, vb = vw + C * DIN, ew = vb + C,
                         sw = ew + C, sb = sw + C * DIN, total = sb + C;
};
constexpr int FLAT_TOTAL = 2 * Flat<1>::total + 4 * Flat<C>::total + C + 1;   // six convs (gconv3_s2w is unused) + fc
constexpr int PACKED_TOTAL = 2 * Off<1>::total + 3 * Off<C>::total + C + 1;
constexpr int PGRID = 148 * 4;   // CTAs of the parameter-gradient partial sums

__host__ __device__ inline int flat_offset(int conv) { return conv < 2 ? conv * Flat<1>::total : 2 * Flat<1>::total + (conv - 2) * Flat<C>::total; }
__host__ __device__ inline int packed_offset(int conv) { return conv < 2 ? conv * Off<1>::total : 2 * Off<1>::total + (conv - 2) * Off<C>::total; }

// ---------------------------------------------------------------------------------------------------------------
// fused parameter blocks from the flat vector, and the chain rule back
template <int DIN>
__device__ void pack_conv_dev(const float* __restrict__ f, float* __restrict__ p)
{
    using O = Off<DIN>;
    using F = Flat<DIN>;
    for (int e = threadIdx.x; e < O::total; e += blockDim.x) {
        float v = 0.0f;
        if (e < O::vq) {
            const int i = e / DIN, o = e % DIN;
            for (int c = 0; c < C; ++c) v = fmaf(f[F::qw + c * DIN + i], f[F::kw + c * DIN + o], v);
            v *= 0.25f;
        } else if (e < O::wq) {
            const int o = e - O::vq;
            for (int c = 0; c < C; ++c) v = fmaf(f[F::kw + c * DIN + o], f[F::qb + c], v);
            v *= 0.25f;
        } else if (e < O::sq) {
            const int i = e - O::wq;
            for (int c = 0; c < C; ++c) v = fmaf(f[F::qw + c * DIN + i], f[F::ew + c], v);
            v *= 0.25f;
        } else if (e == O::sq) {
            for (int c = 0; c < C; ++c) v = fmaf(f[F::ew + c], f[F::qb + c], v);
            v *= 0.25f;
        } else if (e < O::wv) {
            v = 0.0f;
        } else if (e < O::bv) {
            const int d = (e - O::wv) / C, c = (e - O::wv) % C;
            v = f[F::vw + c * DIN + d];
        } else if (e < O::ws) {
            v = f[F::vb + e - O::bv];
        } else if (e < O::bs) {
            const int d = (e - O::ws) / C, c = (e - O::ws) % C;
            v = f[F::sw + c * DIN + d];
        } else if (e < O::we) {
            v = f[F::sb + e - O::bs];
        } else {
            v = f[F::ew + e - O::we];
        }
        p[e] = v;
    }
}

__global__ void __launch_bounds__(256) k_gnn_pack(const float* __restrict__ flat, float* __restrict__ packed)
{
    const int b = blockIdx.x;
    if (b < 2) pack_conv_dev<1>(flat + flat_offset(b), packed + packed_offset(b));
    else if (b < 5) pack_conv_dev<C>(flat + flat_offset(b), packed + packed_offset(b));
    else if (threadIdx.x <= C) packed[packed_offset(5) + threadIdx.x] = flat[flat_offset(6) + threadIdx.x];
}

template <int DIN>
__device__ void unpack_conv_dev(const float* __restrict__ f, const float* __restrict__ dp, float* __restrict__ df)
{
    using O = Off<DIN>;
    using F = Flat<DIN>;
    for (int e = threadIdx.x; e < F::total; e += blockDim.x) {
        float v = 0.0f;
        if (e < F::kb) {            // d Wk[c][o] = (sum_i Wq[c][i] dMQ[i][o] + bq[c] dvq[o]) / 4
            const int c = e / DIN, o = e % DIN;
            v = f[F::qb + c] * dp[O::vq + o];
            for (int i = 0; i < DIN; ++i) v = fmaf(f[F::qw + c * DIN + i], dp[O::mq + i * DIN + o], v);
            v *= 0.25f;
        } else if (e < F::qw) {     // lin_key.bias shifts every score of a node by the same amount: no gradient
            v = 0.0f;
        } else if (e < F::qb) {     // d Wq[c][i] = (sum_o Wk[c][o] dMQ[i][o] + We[c] dwq[i]) / 4
            const int c = (e - F::qw) / DIN, i = (e - F::qw) % DIN;
            v = f[F::ew + c] * dp[O::wq + i];
            for (int o = 0; o < DIN; ++o) v = fmaf(f[F::kw + c * DIN + o], dp[O::mq + i * DIN + o], v);
            v *= 0.25f;
        } else if (e < F::vw) {     // d bq[c] = (sum_o Wk[c][o] dvq[o] + We[c] dsq) / 4
            const int c = e - F::qb;
            v = f[F::ew + c] * dp[O::sq];
            for (int o = 0; o < DIN; ++o) v = fmaf(f[F::kw + c * DIN + o], dp[O::vq + o], v);
            v *= 0.25f;
        } else if (e < F::vb) {
            const int c = (e - F::vw) / DIN, d = (e - F::vw) % DIN;
            v = dp[O::wv + d * C + c];
        } else if (e < F::ew) {
            v = dp[O::bv + e - F::vb];
        } else if (e < F::sw) {     // d We[c] = dWe[c] + (sum_i Wq[c][i] dwq[i] + bq[c] dsq) / 4
            const int c = e - F::ew;
            v = f[F::qb + c] * dp[O::sq];
            for (int i = 0; i < DIN; ++i) v = fmaf(f[F::qw + c * DIN + i], dp[O::wq + i], v);
            v = fmaf(v, 0.25f, dp[O::we + c]);
        } else if (e < F::sb) {
            const int c = (e - F::sw) / DIN, d = (e - F::sw) % DIN;
            v = dp[O::ws + d * C + c];
        } else {
            v = dp[O::bs + e - F::sb];
        }
        df[e] = v;
    }
}

__global__ void __launch_bounds__(256) k_gnn_unpack_grads(const float* __restrict__ flat, const float* __restrict__ dpacked,
                                                          float* __restrict__ dflat)
{
    const int b = blockIdx.x;
    if (b < 2) unpack_conv_dev<1>(flat + flat_offset(b), dpacked + packed_offset(b), dflat + flat_offset(b));
    else if (b < 5) unpack_conv_dev<C>(flat + flat_offset(b), dpacked + packed_offset(b), dflat + flat_offset(b));
    else if (b == 5) {
        for (int e = threadIdx.x; e < Flat<C>::total; e += blockDim.x) dflat[flat_offset(5) + e] = 0.0f;   // gconv3_s2w (:247)
    } else if (threadIdx.x <= C) dflat[flat_offset(6) + threadIdx.x] = dpacked[packed_offset(5) + threadIdx.x];
}

// ---------------------------------------------------------------------------------------------------------------
// shared pieces of the destination pass

// dxbar = Wv' g, dabar = We . g, dxs = Ws' g over the channels [c0, c0 + CNT) held by this lane
template <int DIN, int CNT>
__device__ __forceinline__ void g_products_xb(const float* prm, int c0, const float* g, float* dxb, float& dab)
{
    using O = Off<DIN>;
#pragma unroll
    for (int k = 0; k < CNT; ++k) {
        const float gc = g[k];
        dab = fmaf(prm[O::we + c0 + k], gc, dab);
#pragma unroll
        for (int d = 0; d < DIN; ++d) dxb[d] = fmaf(prm[O::wv + d * C + c0 + k], gc, dxb[d]);
    }
}
template <int DIN, int CNT>
__device__ __forceinline__ void g_products_xs(const float* prm, int c0, const float* g, float* dxs)
{
    using O = Off<DIN>;
#pragma unroll
    for (int k = 0; k < CNT; ++k) {
        const float gc = g[k];
#pragma unroll
        for (int d = 0; d < DIN; ++d) dxs[d] = fmaf(prm[O::ws + d * C + c0 + k], gc, dxs[d]);
    }
}
template <int DIN, int CNT>
__device__ __forceinline__ void g_products(const float* prm, int c0, const float* g, float* dxb, float* dxs, float& dab)
{
    g_products_xb<DIN, CNT>(prm, c0, g, dxb, dab);
    g_products_xs<DIN, CNT>(prm, c0, g, dxs);
}

// v[0..DIN) -> p[0..DIN): 128-bit stores when the row allows it (a lane writes its node's vectors: scattered rows)
template <int DIN>
__device__ __forceinline__ void store_vec(float* p, const float* v)
{
    if constexpr (DIN % 4 == 0) {
#pragma unroll
        for (int k = 0; k < DIN / 4; ++k) reinterpret_cast<float4*>(p)[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
    } else {
#pragma unroll
        for (int d = 0; d < DIN; ++d) p[d] = v[d];
    }
}

// one edge of the second sweep
template <int DIN>
__device__ __forceinline__ void edge_ds(float a, const float* xj, const float* qt, float qe, float lse, const float* dxb, float dab,
                                        float D, float* dqt, float& dqe)
{
    float s = a * qe, da = dab * a;
#pragma unroll
    for (int d = 0; d < DIN; ++d) { s = fmaf(qt[d], xj[d], s); da = fmaf(dxb[d], xj[d], da); }
    const float ds = __expf(s - lse) * (da - D);
    dqe = fmaf(ds, a, dqe);
#pragma unroll
    for (int d = 0; d < DIN; ++d) dqt[d] = fmaf(ds, xj[d], dqt[d]);
}

// d x_i = dxs + wq dqe + MQ dqt, written by the calling lane
template <int DIN>
__device__ __forceinline__ void store_dx_dst(const float* prm, const float* dxs, const float* dqt, float dqe, float* out)
{
    using O = Off<DIN>;
    float dx[DIN];
#pragma unroll
    for (int d = 0; d < DIN; ++d) dx[d] = fmaf(prm[O::wq + d], dqe, dxs[d]);
    if constexpr (DIN % 4 == 0) {
#pragma unroll
        for (int d = 0; d < DIN; ++d) {
#pragma unroll
            for (int o4 = 0; o4 < DIN / 4; ++o4) {
                const float4 w = *reinterpret_cast<const float4*>(prm + O::mq + d * DIN + 4 * o4);
                dx[d] = fmaf(w.x, dqt[4 * o4], dx[d]); dx[d] = fmaf(w.y, dqt[4 * o4 + 1], dx[d]);
                dx[d] = fmaf(w.z, dqt[4 * o4 + 2], dx[d]); dx[d] = fmaf(w.w, dqt[4 * o4 + 3], dx[d]);
            }
        }
#pragma unroll
        for (int k = 0; k < DIN / 4; ++k)
            reinterpret_cast<float4*>(out)[k] = make_float4(dx[4 * k], dx[4 * k + 1], dx[4 * k + 2], dx[4 * k + 3]);
    } else {
#pragma unroll
        for (int d = 0; d < DIN; ++d) {
#pragma unroll
            for (int o = 0; o < DIN; ++o) dx[d] = fmaf(prm[O::mq + d * DIN + o], dqt[o], dx[d]);
            out[d] = dx[d];
        }
    }
}

// the record of a destination node, in two parts: {qt, qe, lse} and {dxbar, dabar, D}
template <int DIN>
__device__ __forceinline__ void store_rec_q(float* r, const float* qt, float qe, float lse)
{
    using R = Rec<DIN>;
    if constexpr (DIN % 4 == 0) {
        float4* q = reinterpret_cast<float4*>(r);
#pragma unroll
        for (int k = 0; k < DIN / 4; ++k) q[k] = make_float4(qt[4 * k], qt[4 * k + 1], qt[4 * k + 2], qt[4 * k + 3]);
        *reinterpret_cast<float2*>(r + R::qe) = make_float2(qe, lse);
    } else {
#pragma unroll
        for (int d = 0; d < DIN; ++d) r[R::qt + d] = qt[d];
        r[R::qe] = qe; r[R::lse] = lse;
    }
}
template <int DIN>
__device__ __forceinline__ void store_rec_d(float* r, const float* dxb, float dab, float D)
{
    using R = Rec<DIN>;
    if constexpr (DIN % 4 == 0) {
        float4* q = reinterpret_cast<float4*>(r + R::dxb);
#pragma unroll
        for (int k = 0; k < DIN / 4; ++k) q[k] = make_float4(dxb[4 * k], dxb[4 * k + 1], dxb[4 * k + 2], dxb[4 * k + 3]);
        *reinterpret_cast<float2*>(r + R::dab) = make_float2(dab, D);
    } else {
#pragma unroll
        for (int d = 0; d < DIN; ++d) r[R::dxb + d] = dxb[d];
        r[R::dab] = dab; r[R::dd] = D;
    }
}
template <int DIN>
__device__ __forceinline__ void store_rec(float* r, const float* qt, const float* dxb, float qe, float lse, float dab, float D)
{
    store_rec_q<DIN>(r, qt, qe, lse);
    store_rec_d<DIN>(r, dxb, dab, D);
}

template <int DIN>
__device__ __forceinline__ void load_rec(const float* __restrict__ rec_i, float* qt, float* dxb, float& qe, float& lse, float& dab, float& D)
{
    using R = Rec<DIN>;
    load_row<DIN>(rec_i + R::qt, qt);
    load_row<DIN>(rec_i + R::dxb, dxb);
    if constexpr (DIN % 4 == 0) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(rec_i + R::qe));
        qe = t.x; lse = t.y; dab = t.z; D = t.w;
    } else {
        qe = __ldg(rec_i + R::qe); lse = __ldg(rec_i + R::lse); dab = __ldg(rec_i + R::dab); D = __ldg(rec_i + R::dd);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// destination pass, rows of at most `chunk` edges: S lanes per row, in two kernels (one kernel needed 255 registers and
// ran one CTA per SM; the gathers want occupancy):
//   k_gnn_bwd_dst_stats: the forward's row walk again: softmax statistics, pre-activation output -> g (ReLU mask applied), the
//                        query-side part of the node's record and its per-node vectors;
//   k_gnn_bwd_dst_dense: the dense maps dxbar = Wv' g, dabar, D (rest of the record), Ws' g parked in d x -- 16 lanes per node;
//   k_gnn_bwd_dst_sweep: second sweep of the row's edges from the record -> dqt, dqe, d x.
//   upstream gradient: gh[nd][16] (d loss / d relu(out)), or for the last conv dout[nd] and fcw[16] (the folded Linear(16, 1));
//   dxdst may be null (first layer: the inputs need no gradient); hfc[nd][16] = dout_i relu(out_i) for d fc.weight.
template <int S, int DIN>
__global__ void __launch_bounds__(256, 3) k_gnn_bwd_dst_stats(int nd, const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                                     const double* __restrict__ values, const float* __restrict__ hdst,
                                                     const float* __restrict__ hsrc, const float* __restrict__ prm_g,
                                                     const float* __restrict__ gh, const float* __restrict__ dout,
                                                     const float* __restrict__ fcw_g, int chunk, float* __restrict__ rec,
                                                     float* __restrict__ nr, float* __restrict__ hfc)
{
    using O = Off<DIN>;
    using N = NR<DIN>;
    __shared__ __align__(16) float prm[O::total + C];
    for (int k = threadIdx.x; k < O::total; k += blockDim.x) prm[k] = prm_g[k];
    if (fcw_g && threadIdx.x < C) prm[O::total + threadIdx.x] = fcw_g[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, gl = lane & (S - 1);
    constexpr int RPW = 32 / S;
    constexpr int CNT = S <= C ? C / S : 1;
    const int c0 = S <= C ? gl * CNT : (gl >> 1);
    const bool owner = !(S == 32 && (gl & 1));
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * RPW; base < nd; base += warps * RPW) {
        const int i = base + lane / S;
        int e0 = 0, e1 = 0;
        if (i < nd) { e0 = __ldg(indptr + i); e1 = __ldg(indptr + i + 1); }
        const bool live = i < nd && e1 - e0 <= chunk;
        if (!live) e0 = e1 = 0;
        float x[DIN], qt[DIN], qe;
#pragma unroll
        for (int d = 0; d < DIN; ++d) x[d] = 0.0f;
        if (live) load_row<DIN>(hdst + (size_t)i * DIN, x);
        dst_prologue<S, DIN>(prm, x, gl, qt, qe);
        State<DIN> st;
        state_init<DIN>(st);
        if (e1 > e0) edge_loop<S, DIN>(indices, values, hsrc, e0, e1, gl, qt, qe, st);
        merge_group<S, DIN>(st);
        const bool any = st.l > 0.0f;
        const float inv = any ? 1.0f / st.l : 0.0f;
        const float lse = any ? st.m + __logf(st.l) : 0.0f;
        const bool writer = live && gl == 0;
        float* rc = rec + (size_t)i * Rec<DIN>::total;
        // the phases below are ordered so that few vectors are live at a time (S = 1: one lane holds all 16 channels):
        // the query-side part of the record leaves the registers first
        if (writer) store_rec_q<DIN>(rc, qt, qe, lse);
        // xbar, abar in place of the unnormalised sums
#pragma unroll
        for (int d = 0; d < DIN; ++d) st.acc[d] *= inv;
        const float abar = st.pa * inv;
        if (writer) {
            float* r = nr + (size_t)i * N::total;
            store_vec<DIN>(r + N::xbar, st.acc);
            r[N::abar] = abar;
            r[N::any] = any ? 1.0f : 0.0f;
        }
        // this lane's channels: pre-activation output, ReLU mask, upstream gradient -> g (k_gnn_bwd_dst_dense goes on from it)
        const float di = (dout && live) ? __ldg(dout + i) : 0.0f;
        constexpr int STEP = CNT >= 4 ? 4 : CNT;
#pragma unroll
        for (int k0 = 0; k0 < CNT; k0 += STEP) {
            float o[STEP];
            out_channels<DIN, STEP>(prm, c0 + k0, x, st.acc, abar, 1.0f, any, 0, o);
            if (live && owner) {
                float gk[STEP];
#pragma unroll
                for (int k = 0; k < STEP; ++k) {
                    const int c = c0 + k0 + k;
                    const float up = gh ? __ldg(gh + (size_t)i * C + c) : di * prm[O::total + c];
                    gk[k] = o[k] > 0.0f ? up : 0.0f;
                    if (hfc) hfc[(size_t)i * C + c] = di * fmaxf(o[k], 0.0f);
                }
                float* gp = nr + (size_t)i * N::total + N::g + c0 + k0;
                if constexpr (STEP == 4) *reinterpret_cast<float4*>(gp) = make_float4(gk[0], gk[1], gk[2], gk[3]);
                else if constexpr (STEP == 2) *reinterpret_cast<float2*>(gp) = make_float2(gk[0], gk[1]);
                else gp[0] = gk[0];
            }
        }
    }
}

// The dense maps of the destination pass for the rows of at most `chunk` edges, 16 lanes per node (lane q holds channel q of
// g and produces entry q of dxbar = Wv' g and of Ws' g with its columns of the weights in registers):
//   record {dxbar, dabar = We . g, D = dxbar . xbar + dabar abar}, and Ws' g parked in the row's d x.
template <int DIN>
__global__ void __launch_bounds__(256) k_gnn_bwd_dst_dense(int nd, const int32_t* __restrict__ indptr, const float* __restrict__ prm,
                                                           int chunk, const float* __restrict__ nr, float* __restrict__ rec,
                                                           float* __restrict__ dxdst)
{
    using O = Off<DIN>;
    using N = NR<DIN>;
    using R = Rec<DIN>;
    const int lane = threadIdx.x & 31, q = lane & 15;
    const bool has_d = q < DIN;
    float wv[C], ws[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        wv[c] = has_d ? __ldg(prm + O::wv + q * C + c) : 0.0f;
        ws[c] = has_d ? __ldg(prm + O::ws + q * C + c) : 0.0f;
    }
    const float we = __ldg(prm + O::we + q);
    const int groups = (gridDim.x * blockDim.x) >> 4;
    // trip count uniform over the warp (the shuffles below are warp-wide)
    for (int base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 2; base < nd; base += groups) {
        const int i = base + (lane >> 4);
        bool live = i < nd;
        if (live) live = __ldg(indptr + i + 1) - __ldg(indptr + i) <= chunk;
        const float* p = nr + (size_t)(live ? i : 0) * N::total;
        const float gq = live ? p[N::g + q] : 0.0f;
        const float xbq = (live && has_d) ? p[N::xbar + q] : 0.0f;
        const float abar = live ? p[N::abar] : 0.0f;
        float dxb = 0.0f, dxs = 0.0f;
        float dab = we * gq;
        const int src0 = lane & 16;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float gc = __shfl_sync(FULLM, gq, src0 + c);
            dxb = fmaf(wv[c], gc, dxb);
            dxs = fmaf(ws[c], gc, dxs);
        }
        float D = dxb * xbq;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            dab += __shfl_xor_sync(FULLM, dab, o);
            D += __shfl_xor_sync(FULLM, D, o);
        }
        D = fmaf(dab, abar, D);
        if (live) {
            float* r = rec + (size_t)i * R::total;
            if (has_d) {
                r[R::dxb + q] = dxb;
                if (dxdst) dxdst[(size_t)i * DIN + q] = dxs;
            }
            if (q == 0) { r[R::dab] = dab; r[R::dd] = D; }
        }
    }
}

template <int S, int DIN>
__global__ void __launch_bounds__(256, 3) k_gnn_bwd_dst_sweep(int nd, const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                                           const double* __restrict__ values, const float* __restrict__ hsrc,
                                                           const float* __restrict__ prm_g, int chunk, const float* __restrict__ rec,
                                                           float* __restrict__ nr, float* __restrict__ dxdst)
{
    using O = Off<DIN>;
    using N = NR<DIN>;
    __shared__ __align__(16) float prm[O::wv];   // MQ, wq: the d x epilogue
    for (int k = threadIdx.x; k < O::wv; k += blockDim.x) prm[k] = prm_g[k];
    __syncthreads();
    const int lane = threadIdx.x & 31, gl = lane & (S - 1);
    constexpr int RPW = 32 / S;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * RPW; base < nd; base += warps * RPW) {
        const int i = base + lane / S;
        int e0 = 0, e1 = 0;
        if (i < nd) { e0 = __ldg(indptr + i); e1 = __ldg(indptr + i + 1); }
        const bool live = i < nd && e1 - e0 <= chunk;
        if (!live) e0 = e1 = 0;
        float dqt[DIN], dqe = 0.0f;
#pragma unroll
        for (int d = 0; d < DIN; ++d) dqt[d] = 0.0f;
        if (e1 > e0) {
            float qt[DIN], dxb[DIN], qe, lse, dab, D;
            load_rec<DIN>(rec + (size_t)i * Rec<DIN>::total, qt, dxb, qe, lse, dab, D);
            for (int e = e0 + gl; e < e1; e += 2 * S) {
                const int eb = e + S;
                const bool two = eb < e1;
                const int ja = __ldg(indices + e);
                const int jb = two ? __ldg(indices + eb) : ja;
                const float aa = (float)__ldg(values + e);
                const float ab = two ? (float)__ldg(values + eb) : 0.0f;
                float xa[DIN], xb[DIN];
                load_row<DIN>(hsrc + (size_t)ja * DIN, xa);
                load_row<DIN>(hsrc + (size_t)jb * DIN, xb);
                edge_ds<DIN>(aa, xa, qt, qe, lse, dxb, dab, D, dqt, dqe);
                if (two) edge_ds<DIN>(ab, xb, qt, qe, lse, dxb, dab, D, dqt, dqe);
            }
        }
#pragma unroll
        for (int o = S / 2; o > 0; o >>= 1) {
            dqe += __shfl_xor_sync(FULLM, dqe, o);
#pragma unroll
            for (int d = 0; d < DIN; ++d) dqt[d] += __shfl_xor_sync(FULLM, dqt[d], o);
        }
        if (live && gl == 0) {
            float* r = nr + (size_t)i * N::total;
            store_vec<DIN>(r + N::dqt, dqt);
            r[N::dqe] = dqe;
            if (dxdst) {
                float park[DIN];
#pragma unroll
                for (int d = 0; d < DIN; ++d) park[d] = dxdst[(size_t)i * DIN + d];
                store_dx_dst<DIN>(prm, park, dqt, dqe, dxdst + (size_t)i * DIN);
            }
        }
    }
}

// destination pass, rows above `chunk` edges (osa-60: rows of 173 366 edges).  The row is cut into the forward's items
// (one warp each, mllp_gnn_side.items):  k_gnn_conv_items leaves the items' partial softmax states in the scratch, then
//   k_gnn_bwd_dst_long_stats  (one warp per row) merges them in a fixed order, forms g, dxbar, dabar, D, writes the node's
//                             record and per-node vectors and parks Ws' g in the row's d x;
//   k_gnn_bwd_dst_items       (one warp per item) sweeps the item's edges for its part of dqt, dqe -> scratch;
//   k_gnn_bwd_dst_long_final  (one warp per row) adds the parts in a fixed order and finishes d x.
template <int DIN>
__global__ void __launch_bounds__(256) k_gnn_bwd_dst_long_stats(int nlong, const int32_t* __restrict__ long_rows,
                                                                const int32_t* __restrict__ first, const float* __restrict__ scratch,
                                                                const float* __restrict__ hdst, const float* __restrict__ prm_g,
                                                                const float* __restrict__ gh, const float* __restrict__ dout,
                                                                const float* __restrict__ fcw_g, float* __restrict__ rec,
                                                                float* __restrict__ nr, float* __restrict__ dxdst, float* __restrict__ hfc)
{
    using O = Off<DIN>;
    using N = NR<DIN>;
    __shared__ __align__(16) float prm[O::total + C];
    for (int k = threadIdx.x; k < O::total; k += blockDim.x) prm[k] = prm_g[k];
    if (fcw_g && threadIdx.x < C) prm[O::total + threadIdx.x] = fcw_g[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, c = lane & 15;
    const bool owner = lane < C;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < nlong; r += warps) {
        const int i = __ldg(long_rows + r);
        State<DIN> st;
        merge_items<DIN>(scratch, __ldg(first + r), __ldg(first + r + 1), lane, st);
        const bool any = st.l > 0.0f;
        const float inv = any ? 1.0f / st.l : 0.0f;
        const float lse = any ? st.m + __logf(st.l) : 0.0f;
        float x[DIN], qt[DIN], qe;
        load_row<DIN>(hdst + (size_t)i * DIN, x);
        dst_prologue<1, DIN>(prm, x, 0, qt, qe);
        const float di = dout ? __ldg(dout + i) : 0.0f;
        float o1[1];
        out_channels<DIN, 1>(prm, c, x, st.acc, st.pa, inv, any, 0, o1);
        float g[1] = {0.0f};
        if (owner) {
            const float up = gh ? __ldg(gh + (size_t)i * C + c) : di * prm[O::total + c];
            g[0] = o1[0] > 0.0f ? up : 0.0f;
            nr[(size_t)i * N::total + N::g + c] = g[0];
            if (hfc) hfc[(size_t)i * C + c] = di * fmaxf(o1[0], 0.0f);
        }
        float dxb[DIN], dxs[DIN], dab = 0.0f;
#pragma unroll
        for (int d = 0; d < DIN; ++d) { dxb[d] = 0.0f; dxs[d] = 0.0f; }
        g_products<DIN, 1>(prm, c, g, dxb, dxs, dab);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            dab += __shfl_xor_sync(FULLM, dab, o);
#pragma unroll
            for (int d = 0; d < DIN; ++d) {
                dxb[d] += __shfl_xor_sync(FULLM, dxb[d], o);
                dxs[d] += __shfl_xor_sync(FULLM, dxs[d], o);
            }
        }
        const float abar = st.pa * inv;
        float D = dab * abar;
#pragma unroll
        for (int d = 0; d < DIN; ++d) D = fmaf(dxb[d], st.acc[d] * inv, D);
        if (lane == 0) {
            float* q = nr + (size_t)i * N::total;
#pragma unroll
            for (int d = 0; d < DIN; ++d) q[N::xbar + d] = st.acc[d] * inv;
            q[N::abar] = abar;
            q[N::any] = any ? 1.0f : 0.0f;
            store_rec<DIN>(rec + (size_t)i * Rec<DIN>::total, qt, dxb, qe, lse, dab, D);
            if (dxdst) {
#pragma unroll
                for (int d = 0; d < DIN; ++d) dxdst[(size_t)i * DIN + d] = dxs[d];
            }
        }
    }
}

template <int DIN>
__global__ void __launch_bounds__(256) k_gnn_bwd_dst_items(int nitems, const int32_t* __restrict__ items, const int32_t* __restrict__ indices,
                                                           const double* __restrict__ values, const float* __restrict__ hsrc,
                                                           const float* __restrict__ rec, float* __restrict__ scratch)
{
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < nitems; t += warps) {
        const int i = __ldg(items + 3 * t), e0 = __ldg(items + 3 * t + 1), e1 = __ldg(items + 3 * t + 2);
        float qt[DIN], dxb[DIN], qe, lse, dab, D;
        load_rec<DIN>(rec + (size_t)i * Rec<DIN>::total, qt, dxb, qe, lse, dab, D);
        float dqt[DIN], dqe = 0.0f;
#pragma unroll
        for (int d = 0; d < DIN; ++d) dqt[d] = 0.0f;
        for (int e = e0 + lane; e < e1; e += 64) {
            const int eb = e + 32;
            const bool two = eb < e1;
            const int ja = __ldg(indices + e);
            const int jb = two ? __ldg(indices + eb) : ja;
            const float aa = (float)__ldg(values + e);
            const float ab = two ? (float)__ldg(values + eb) : 0.0f;
            float xa[DIN], xb[DIN];
            load_row<DIN>(hsrc + (size_t)ja * DIN, xa);
            load_row<DIN>(hsrc + (size_t)jb * DIN, xb);
            edge_ds<DIN>(aa, xa, qt, qe, lse, dxb, dab, D, dqt, dqe);
            if (two) edge_ds<DIN>(ab, xb, qt, qe, lse, dxb, dab, D, dqt, dqe);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            dqe += __shfl_xor_sync(FULLM, dqe, o);
#pragma unroll
            for (int d = 0; d < DIN; ++d) dqt[d] += __shfl_xor_sync(FULLM, dqt[d], o);
        }
        if (lane == 0) {
            float* o = scratch + (size_t)t * ITEM_FLOATS;
#pragma unroll
            for (int d = 0; d < DIN; ++d) o[d] = dqt[d];
            o[DIN] = dqe;
        }
    }
}

template <int DIN>
__global__ void __launch_bounds__(256) k_gnn_bwd_dst_long_final(int nlong, const int32_t* __restrict__ long_rows,
                                                                const int32_t* __restrict__ first, const float* __restrict__ scratch,
                                                                const float* __restrict__ prm_g, float* __restrict__ nr,
                                                                float* __restrict__ dxdst)
{
    using O = Off<DIN>;
    using N = NR<DIN>;
    __shared__ __align__(16) float prm[O::wv];
    for (int k = threadIdx.x; k < O::wv; k += blockDim.x) prm[k] = prm_g[k];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < nlong; r += warps) {
        const int i = __ldg(long_rows + r);
        const int t1 = __ldg(first + r + 1);
        float dqt[DIN], dqe = 0.0f;
#pragma unroll
        for (int d = 0; d < DIN; ++d) dqt[d] = 0.0f;
        for (int t = __ldg(first + r) + lane; t < t1; t += 32) {
            const float* o = scratch + (size_t)t * ITEM_FLOATS;
#pragma unroll
            for (int d = 0; d < DIN; ++d) dqt[d] += o[d];
            dqe += o[DIN];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            dqe += __shfl_xor_sync(FULLM, dqe, o);
#pragma unroll
            for (int d = 0; d < DIN; ++d) dqt[d] += __shfl_xor_sync(FULLM, dqt[d], o);
        }
        if (lane == 0) {
            float* q = nr + (size_t)i * N::total;
            store_vec<DIN>(q + N::dqt, dqt);
            q[N::dqe] = dqe;
            if (dxdst) {
                float dxs[DIN];
#pragma unroll
                for (int d = 0; d < DIN; ++d) dxs[d] = dxdst[(size_t)i * DIN + d];   // Ws' g, parked by the stats kernel
                store_dx_dst<DIN>(prm, dxs, dqt, dqe, dxdst + (size_t)i * DIN);
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------
// source pass along the transposed structure (rows = source nodes j of the conv, indices = destination nodes i):
//   d x_j (+)= sum_i alpha_ij dxbar_i + ds_ij qt_i
template <int DIN>
__device__ __forceinline__ void edge_src(float a, const float* __restrict__ rec_i, const float* xj, float* dx)
{
    using R = Rec<DIN>;
    float qt[DIN], dxb[DIN];
    load_row<DIN>(rec_i + R::qt, qt);
    load_row<DIN>(rec_i + R::dxb, dxb);
    float qe, lse, dab, D;
    if constexpr (DIN % 4 == 0) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(rec_i + R::qe));
        qe = t.x; lse = t.y; dab = t.z; D = t.w;
    } else {
        qe = __ldg(rec_i + R::qe); lse = __ldg(rec_i + R::lse); dab = __ldg(rec_i + R::dab); D = __ldg(rec_i + R::dd);
    }
    float s = a * qe, da = dab * a;
#pragma unroll
    for (int d = 0; d < DIN; ++d) { s = fmaf(qt[d], xj[d], s); da = fmaf(dxb[d], xj[d], da); }
    const float al = __expf(s - lse);
    const float ds = al * (da - D);
#pragma unroll
    for (int d = 0; d < DIN; ++d) dx[d] = fmaf(al, dxb[d], fmaf(ds, qt[d], dx[d]));
}

template <int DIN>
__device__ __forceinline__ void store_dx_src(float* out, const float* dx, int accumulate)
{
    static_assert(DIN % 4 == 0, "source pass: 16 channels");
    float4* q = reinterpret_cast<float4*>(out);
#pragma unroll
    for (int k = 0; k < DIN / 4; ++k) {
        float4 v = make_float4(dx[4 * k], dx[4 * k + 1], dx[4 * k + 2], dx[4 * k + 3]);
        if (accumulate) { const float4 t = q[k]; v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
        q[k] = v;
    }
}

template <int S, int DIN>
__global__ void __launch_bounds__(256) k_gnn_bwd_src(int ns, const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                                     const double* __restrict__ values, const float* __restrict__ hsrc,
                                                     const float* __restrict__ rec, int chunk, float* __restrict__ dxsrc, int accumulate)
{
    const int lane = threadIdx.x & 31, gl = lane & (S - 1);
    constexpr int RPW = 32 / S;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * RPW; base < ns; base += warps * RPW) {
        const int j = base + lane / S;
        int e0 = 0, e1 = 0;
        if (j < ns) { e0 = __ldg(indptr + j); e1 = __ldg(indptr + j + 1); }
        const bool live = j < ns && e1 - e0 <= chunk;
        if (!live) e0 = e1 = 0;
        float xj[DIN], dx[DIN];
#pragma unroll
        for (int d = 0; d < DIN; ++d) { xj[d] = 0.0f; dx[d] = 0.0f; }
        if (live) load_row<DIN>(hsrc + (size_t)j * DIN, xj);
        for (int e = e0 + gl; e < e1; e += 2 * S) {
            const int eb = e + S;
            const bool two = eb < e1;
            const int ia = __ldg(indices + e);
            const int ib = two ? __ldg(indices + eb) : ia;
            const float aa = (float)__ldg(values + e);
            const float ab = two ? (float)__ldg(values + eb) : 0.0f;
            edge_src<DIN>(aa, rec + (size_t)ia * Rec<DIN>::total, xj, dx);
            if (two) edge_src<DIN>(ab, rec + (size_t)ib * Rec<DIN>::total, xj, dx);
        }
#pragma unroll
        for (int o = S / 2; o > 0; o >>= 1) {
#pragma unroll
            for (int d = 0; d < DIN; ++d) dx[d] += __shfl_xor_sync(FULLM, dx[d], o);
        }
        if (live && gl == 0) store_dx_src<DIN>(dxsrc + (size_t)j * DIN, dx, accumulate);
    }
}

// source pass, rows of the transposed structure above its `chunk`: one warp per item, then one warp per row adds the items'
// parts in a fixed order
template <int DIN>
__global__ void __launch_bounds__(256) k_gnn_bwd_src_items(int nitems, const int32_t* __restrict__ items, const int32_t* __restrict__ indices,
                                                           const double* __restrict__ values, const float* __restrict__ hsrc,
                                                           const float* __restrict__ rec, float* __restrict__ scratch)
{
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < nitems; t += warps) {
        const int j = __ldg(items + 3 * t), e0 = __ldg(items + 3 * t + 1), e1 = __ldg(items + 3 * t + 2);
        float xj[DIN], dx[DIN];
        load_row<DIN>(hsrc + (size_t)j * DIN, xj);
#pragma unroll
        for (int d = 0; d < DIN; ++d) dx[d] = 0.0f;
        for (int e = e0 + lane; e < e1; e += 64) {
            const int eb = e + 32;
            const bool two = eb < e1;
            const int ia = __ldg(indices + e);
            const int ib = two ? __ldg(indices + eb) : ia;
            const float aa = (float)__ldg(values + e);
            const float ab = two ? (float)__ldg(values + eb) : 0.0f;
            edge_src<DIN>(aa, rec + (size_t)ia * Rec<DIN>::total, xj, dx);
            if (two) edge_src<DIN>(ab, rec + (size_t)ib * Rec<DIN>::total, xj, dx);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int d = 0; d < DIN; ++d) dx[d] += __shfl_xor_sync(FULLM, dx[d], o);
        }
        if (lane == 0) {
            float4* o = reinterpret_cast<float4*>(scratch + (size_t)t * ITEM_FLOATS);
#pragma unroll
            for (int k = 0; k < DIN / 4; ++k) o[k] = make_float4(dx[4 * k], dx[4 * k + 1], dx[4 * k + 2], dx[4 * k + 3]);
        }
    }
}

template <int DIN>
__global__ void __launch_bounds__(256) k_gnn_bwd_src_long_final(int nlong, const int32_t* __restrict__ long_rows,
                                                                const int32_t* __restrict__ first, const float* __restrict__ scratch,
                                                                float* __restrict__ dxsrc, int accumulate)
{
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < nlong; r += warps) {
        const int j = __ldg(long_rows + r);
        const int t1 = __ldg(first + r + 1);
        float dx[DIN];
#pragma unroll
        for (int d = 0; d < DIN; ++d) dx[d] = 0.0f;
        for (int t = __ldg(first + r) + lane; t < t1; t += 32) {
            const float4* o = reinterpret_cast<const float4*>(scratch + (size_t)t * ITEM_FLOATS);
#pragma unroll
            for (int k = 0; k < DIN / 4; ++k) {
                const float4 v = o[k];
                dx[4 * k] += v.x; dx[4 * k + 1] += v.y; dx[4 * k + 2] += v.z; dx[4 * k + 3] += v.w;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int d = 0; d < DIN; ++d) dx[d] += __shfl_xor_sync(FULLM, dx[d], o);
        }
        if (lane == 0) store_dx_src<DIN>(dxsrc + (size_t)j * DIN, dx, accumulate);
    }
}


// ---------------------------------------------------------------------------------------------------------------
// parameter gradients of one conv (layout of the fused block): every entry is sum_i U_i[a] V_i[b] with
//   U_i = {x_i[DIN], xbar_i[DIN], abar_i, any_i, 1},  V_i = {dqt_i[DIN], dqe_i, g_i[16]}
// A CTA walks tiles of 32 nodes: the nodes' rows {per-node vectors | x | 1} go to shared memory as float4 loads that are all
// issued before the previous tile is consumed; a thread owns up to 4 entries of the block and keeps them in registers.
template <int DIN>
__global__ void __launch_bounds__(256) k_gnn_param_partial(int nd, const float* __restrict__ hdst, const float* __restrict__ nr,
                                                           float* __restrict__ partial)
{
    using O = Off<DIN>;
    using N = NR<DIN>;
    constexpr int TN = 32, NT = N::total, NV = NT / 4, W = NT + DIN + 1, WS = (W + 3) & ~3;
    constexpr int COL_X = NT, COL_ONE = NT + DIN;
    constexpr int LOADS = (TN * NV + 255) / 256;
    static_assert(4 * 256 >= O::total, "four entries per thread cover the block");
    __shared__ __align__(16) float T[TN][WS];
    // a thread owns the four CONSECUTIVE entries e0 .. e0 + 3 of the block: in all blocks but wq / sq they share the U
    // operand and their V operands are four consecutive floats (one 128-bit shared-memory load per node)
    const int e0 = 4 * (int)threadIdx.x;
    int ua[4], vb[4];
    float acc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int e = e0 + k;
        int a = -1, b = 0;
        if (e < O::vq) { a = COL_X + e / DIN; b = N::dqt + e % DIN; }
        else if (e < O::wq) { a = COL_ONE; b = N::dqt + e - O::vq; }
        else if (e < O::sq) { a = COL_X + e - O::wq; b = N::dqe; }
        else if (e == O::sq) { a = COL_ONE; b = N::dqe; }
        else if (e < O::wv) { a = -1; }
        else if (e < O::bv) { a = N::xbar + (e - O::wv) / C; b = N::g + (e - O::wv) % C; }
        else if (e < O::ws) { a = N::any; b = N::g + e - O::bv; }
        else if (e < O::bs) { a = COL_X + (e - O::ws) / C; b = N::g + (e - O::ws) % C; }
        else if (e < O::we) { a = COL_ONE; b = N::g + e - O::bs; }
        else if (e < O::total) { a = N::abar; b = N::g + e - O::we; }
        ua[k] = a; vb[k] = b; acc[k] = 0.0f;
    }
    const bool fast = ua[0] >= 0 && ua[1] == ua[0] && ua[2] == ua[0] && ua[3] == ua[0] && (vb[0] & 3) == 0 && vb[1] == vb[0] + 1 &&
                      vb[2] == vb[0] + 2 && vb[3] == vb[0] + 3;
    const bool any_entry = ua[0] >= 0 || ua[1] >= 0 || ua[2] >= 0 || ua[3] >= 0;
    const int ntiles = (nd + TN - 1) / TN;
    float4 buf[LOADS];
    float4 xb = make_float4(0.f, 0.f, 0.f, 0.f);
    auto fetch = [&](int tile) {
#pragma unroll
        for (int u = 0; u < LOADS; ++u) {
            const int idx = threadIdx.x + 256 * u;
            const int r = idx / NV, q = idx % NV;
            const int node = tile * TN + r;
            buf[u] = (idx < TN * NV && node < nd) ? __ldg(reinterpret_cast<const float4*>(nr + (size_t)node * NT) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if constexpr (DIN % 4 == 0) {
            const int r = threadIdx.x / (DIN / 4), q = threadIdx.x % (DIN / 4);
            const int node = tile * TN + r;
            xb = (r < TN && node < nd) ? __ldg(reinterpret_cast<const float4*>(hdst + (size_t)node * DIN) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            const int node = tile * TN + (int)threadIdx.x;
            xb.x = (threadIdx.x < TN && node < nd) ? __ldg(hdst + (size_t)node * DIN) : 0.0f;
        }
    };
    int tile = blockIdx.x;
    if (tile < ntiles) fetch(tile);
    for (; tile < ntiles; tile += gridDim.x) {
        __syncthreads();   // the previous tile has been consumed
#pragma unroll
        for (int u = 0; u < LOADS; ++u) {
            const int idx = threadIdx.x + 256 * u;
            const int r = idx / NV, q = idx % NV;
            if (idx < TN * NV) *reinterpret_cast<float4*>(&T[r][4 * q]) = buf[u];
        }
        if constexpr (DIN % 4 == 0) {
            const int r = threadIdx.x / (DIN / 4), q = threadIdx.x % (DIN / 4);
            if (r < TN) *reinterpret_cast<float4*>(&T[r][COL_X + 4 * q]) = xb;
        } else {
            if (threadIdx.x < TN) T[threadIdx.x][COL_X] = xb.x;
        }
        if (threadIdx.x < TN) T[threadIdx.x][COL_ONE] = (tile * TN + (int)threadIdx.x < nd) ? 1.0f : 0.0f;
        __syncthreads();
        if (tile + (int)gridDim.x < ntiles) fetch(tile + gridDim.x);   // in flight while this tile is consumed
        if (fast) {
#pragma unroll 8
            for (int r = 0; r < TN; ++r) {
                const float u = T[r][ua[0]];
                const float4 v = *reinterpret_cast<const float4*>(&T[r][vb[0]]);
                acc[0] = fmaf(u, v.x, acc[0]); acc[1] = fmaf(u, v.y, acc[1]);
                acc[2] = fmaf(u, v.z, acc[2]); acc[3] = fmaf(u, v.w, acc[3]);
            }
        } else if (any_entry) {
#pragma unroll 4
            for (int r = 0; r < TN; ++r) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (ua[k] >= 0) acc[k] = fmaf(T[r][ua[k]], T[r][vb[k]], acc[k]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (e0 + k < O::total) partial[(size_t)blockIdx.x * O::total + e0 + k] = acc[k];
}

// out[e] = sum over the parts: 32 strided sub-sums per entry (four loads in flight each), added in a fixed order
__global__ void __launch_bounds__(256) k_gnn_sum_parts(int nparts, int width, const float* __restrict__ partial, float* __restrict__ out)
{
    __shared__ float sm[32][9];
    const int el = threadIdx.x & 7, pg = threadIdx.x >> 3;
    const int e = blockIdx.x * 8 + el;
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
    if (e < width) {
        int b = pg;
        for (; b + 96 < nparts; b += 128) {
            s0 += partial[(size_t)b * width + e];
            s1 += partial[(size_t)(b + 32) * width + e];
            s2 += partial[(size_t)(b + 64) * width + e];
            s3 += partial[(size_t)(b + 96) * width + e];
        }
        for (; b < nparts; b += 32) s0 += partial[(size_t)b * width + e];
    }
    sm[pg][el] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (pg == 0 && e < width) {
        float t = 0.0f;
#pragma unroll
        for (int q = 0; q < 32; ++q) t += sm[q][el];
        out[e] = t;
    }
}

// d fc.weight[c] = sum_i hfc[i][c], d fc.bias = sum_i dout[i]: per-CTA partials (17 floats)
__global__ void __launch_bounds__(256) k_gnn_fc_partial(int n, const float* __restrict__ hfc, const float* __restrict__ dout,
                                                        float* __restrict__ partial)
{
    __shared__ float sm[16][C + 1];
    const int r = threadIdx.x >> 4, c = threadIdx.x & 15;
    const int stride = gridDim.x * 16;
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f, b = 0.0f;
    int i = blockIdx.x * 16 + r;
    for (; i + 3 * stride < n; i += 4 * stride) {   // four independent loads in flight
        a0 += hfc[(size_t)i * C + c];
        a1 += hfc[(size_t)(i + stride) * C + c];
        a2 += hfc[(size_t)(i + 2 * stride) * C + c];
        a3 += hfc[(size_t)(i + 3 * stride) * C + c];
        if (c == 0) b += (dout[i] + dout[i + stride]) + (dout[i + 2 * stride] + dout[i + 3 * stride]);
    }
    for (; i < n; i += stride) {
        a0 += hfc[(size_t)i * C + c];
        if (c == 0) b += dout[i];
    }
    sm[r][c] = (a0 + a1) + (a2 + a3);
    if (c == 0) sm[r][C] = b;
    __syncthreads();
    if (threadIdx.x <= C) {
        float s = 0.0f;
        for (int q = 0; q < 16; ++q) s += sm[q][threadIdx.x];
        partial[(size_t)blockIdx.x * (C + 1) + threadIdx.x] = s;
    }
}


// ---------------------------------------------------------------------------------------------------------------
// host side
// one wave of resident CTAs of `kernel` (256 threads, no dynamic shared memory); the occupancy query is cached per kernel
template <class K>
int resident_grid(K kernel, long long want)
{
    static std::mutex mu;
    static std::unordered_map<const void*, int> cache;
    int per_sm = 0;
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find((const void*)kernel);
        if (it != cache.end()) per_sm = it->second;
    }
    if (per_sm == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
        std::lock_guard<std::mutex> lock(mu);
        cache[(const void*)kernel] = per_sm;
    }
    const long long cap = 148LL * per_sm;
    return (int)(want < 1 ? 1 : want > cap ? cap : want);
}

template <int DIN>
int launch_bwd_dst(const mllp_gnn_side& g, const float* hdst, const float* hsrc, const float* prm, const float* gh, const float* dout,
                   const float* fcw, float* rec, float* nr, float* dxdst, float* hfc, cudaStream_t s)
{
    const long long want = (((long long)g.nd * g.group + 31) / 32 + 7) / 8;
    count_launch(g.nlong > 0 ? 7 : 3);
#define MLLP_BWD_DST(SS)                                                                                                              \
    do {                                                                                                                              \
        k_gnn_bwd_dst_stats<SS, DIN><<<resident_grid(k_gnn_bwd_dst_stats<SS, DIN>, want), 256, 0, s>>>(                               \
            g.nd, g.indptr, g.indices, g.values, hdst, hsrc, prm, gh, dout, fcw, g.chunk, rec, nr, hfc);                              \
        k_gnn_bwd_dst_dense<DIN><<<resident_grid(k_gnn_bwd_dst_dense<DIN>, ((long long)g.nd + 15) / 16), 256, 0, s>>>(                \
            g.nd, g.indptr, prm, g.chunk, nr, rec, dxdst);                                                                            \
        k_gnn_bwd_dst_sweep<SS, DIN><<<resident_grid(k_gnn_bwd_dst_sweep<SS, DIN>, want), 256, 0, s>>>(                               \
            g.nd, g.indptr, g.indices, g.values, hsrc, prm, g.chunk, rec, nr, dxdst);                                                 \
    } while (0)
    switch (g.group) {
        case 1: MLLP_BWD_DST(1); break;
        case 2: MLLP_BWD_DST(2); break;
        case 4: MLLP_BWD_DST(4); break;
        case 8: MLLP_BWD_DST(8); break;
        case 16: MLLP_BWD_DST(16); break;
        default: MLLP_BWD_DST(32); break;
    }
#undef MLLP_BWD_DST
    if (g.nlong > 0) {   // cut rows: the forward's items
        k_gnn_conv_items<DIN><<<grid_for_warps(g.nitems), 256, 0, s>>>(g.nitems, g.items, g.indices, g.values, hdst, hsrc, prm, g.scratch);
        k_gnn_bwd_dst_long_stats<DIN><<<grid_for_warps(g.nlong), 256, 0, s>>>(g.nlong, g.long_rows, g.long_first, g.scratch, hdst, prm, gh, dout,
                                                                            fcw, rec, nr, dxdst, hfc);
        k_gnn_bwd_dst_items<DIN><<<grid_for_warps(g.nitems), 256, 0, s>>>(g.nitems, g.items, g.indices, g.values, hsrc, rec, g.scratch);
        k_gnn_bwd_dst_long_final<DIN><<<grid_for_warps(g.nlong), 256, 0, s>>>(g.nlong, g.long_rows, g.long_first, g.scratch, prm, nr, dxdst);
    }
    return cuda_status("mllp_gnn_backward: destination pass");
}

// `t` = the transposed structure: rows are this conv's source nodes
int launch_bwd_src(const mllp_gnn_side& t, const float* hsrc, const float* rec, float* dxsrc, int accumulate, cudaStream_t s)
{
    const long long want = (((long long)t.nd * t.group + 31) / 32 + 7) / 8;
    count_launch(t.nlong > 0 ? 3 : 1);
#define MLLP_BWD_SRC(SS)                                                                                                              \
    k_gnn_bwd_src<SS, C><<<resident_grid(k_gnn_bwd_src<SS, C>, want), 256, 0, s>>>(t.nd, t.indptr, t.indices, t.values, hsrc, rec, t.chunk, \
                                                                                   dxsrc, accumulate)
    switch (t.group) {
        case 1: MLLP_BWD_SRC(1); break;
        case 2: MLLP_BWD_SRC(2); break;
        case 4: MLLP_BWD_SRC(4); break;
        case 8: MLLP_BWD_SRC(8); break;
        case 16: MLLP_BWD_SRC(16); break;
        default: MLLP_BWD_SRC(32); break;
    }
#undef MLLP_BWD_SRC
    if (t.nlong > 0) {
        k_gnn_bwd_src_items<C><<<grid_for_warps(t.nitems), 256, 0, s>>>(t.nitems, t.items, t.indices, t.values, hsrc, rec, t.scratch);
        k_gnn_bwd_src_long_final<C><<<grid_for_warps(t.nlong), 256, 0, s>>>(t.nlong, t.long_rows, t.long_first, t.scratch, dxsrc, accumulate);
    }
    return cuda_status("mllp_gnn_backward: source pass");
}

template <int DIN>
int launch_param_grads(int nd, const float* hdst, const float* nr, float* partial, float* dpacked, cudaStream_t s)
{
    const int ntiles = (nd + 31) / 32;
    const int grid = ntiles < 1 ? 1 : ntiles > PGRID ? PGRID : ntiles;
    count_launch(2);
    k_gnn_param_partial<DIN><<<grid, 256, 0, s>>>(nd, hdst, nr, partial);
    k_gnn_sum_parts<<<(Off<DIN>::total + 7) / 8, 256, 0, s>>>(grid, Off<DIN>::total, partial, dpacked);
    return cuda_status("mllp_gnn_backward: parameter gradients");
}

struct BwdWork {
    float *d1b, *d1a, *d2b, *d2a, *rec1, *rec2, *nr[5], *hfc, *partial, *dpacked;   // nr[k]: per-node vectors of conv k
};
size_t align4(size_t v) { return (v + 3) & ~(size_t)3; }
size_t carve(BwdWork& w, float* base, size_t n, size_t m)
{
    size_t off = 0;
    auto take = [&](size_t cnt) { float* p = base ? base + off : nullptr; off += align4(cnt); return p; };
    w.d1b = take(16 * n); w.d1a = take(16 * n); w.d2b = take(16 * m); w.d2a = take(16 * m);
    w.rec1 = take(Rec<C>::total * n); w.rec2 = take(Rec<C>::total * m);
    // one set of per-node vectors per conv: their parameter-gradient sums may run beside the next conv's passes
    w.nr[0] = take(NR<1>::total * n); w.nr[1] = take(NR<1>::total * m);
    w.nr[2] = take(NR<C>::total * n); w.nr[3] = take(NR<C>::total * m); w.nr[4] = take(NR<C>::total * n);
    w.hfc = take(16 * n);
    w.partial = take((size_t)PGRID * Off<C>::total); w.dpacked = take(PACKED_TOTAL);
    return off;
}
}  // namespace
}  // namespace mllp

using namespace mllp;

extern "C" {

int64_t mllp_gnn_flat_param_floats(void) { return FLAT_TOTAL; }
int64_t mllp_gnn_packed_param_floats(void) { return PACKED_TOTAL; }

int mllp_gnn_pack_params(const float* d_flat, float* d_packed, void* stream)
{
    if (!d_flat || !d_packed) return gfail(MLLP_E_INVALID, "mllp_gnn_pack_params: null pointer");
    count_launch(1);
    k_gnn_pack<<<6, 256, 0, (cudaStream_t)stream>>>(d_flat, d_packed);
    return cuda_status("mllp_gnn_pack_params");
}

int64_t mllp_gnn_backward_workspace_floats(int32_t n, int32_t m)
{
    if (n < 0 || m < 0) return -1;
    BwdWork w;
    return (int64_t)carve(w, nullptr, (size_t)n, (size_t)m) + 16;
}

}  // extern "C"

// With a second stream `s2` and events ev[0..5] (plan capture) the parameter-gradient sums of every conv (and of fc) run on
// a parallel branch: they hang off the conv's destination pass and are only joined before the final unpack, so the chain
// of dependent launches is the passes alone.  Without `s2` everything runs on `s` in order (same results: same kernels).
static int backward_impl(const mllp_gnn_side* to_var, const mllp_gnn_side* to_con, const float* d_x1, const float* d_x2,
                         const float* d_flat, const float* d_packed, const float* d_work, float* d_bwork, const float* d_dout,
                         float* d_dflat, cudaStream_t s, const char* who, cudaStream_t s2 = nullptr, cudaEvent_t* ev = nullptr)
{
    if (!side_ok(to_var) || !side_ok(to_con) || !d_x1 || !d_x2 || !d_flat || !d_packed || !d_work || !d_bwork || !d_dout || !d_dflat)
        return gfail(MLLP_E_INVALID, std::string(who) + ": bad argument");
    if (to_var->ns != to_con->nd || to_con->ns != to_var->nd)
        return gfail(MLLP_E_INVALID, std::string(who) + ": the two sides do not describe one graph");
    const size_t n = (size_t)to_var->nd, m = (size_t)to_con->nd;
    BwdWork w;
    carve(w, d_bwork, n, m);
    // activations the forward left in its workspace (gnn_kernels.cu: forward_impl)
    const float* h1a = d_work;
    const float* h1b = d_work + 16 * n;
    const float* h2a = d_work + 32 * n;
    const float* h2b = d_work + 32 * n + 16 * m;
    const float* P[5];
    float* dP[5];
    for (int k = 0; k < 5; ++k) { P[k] = d_packed + packed_offset(k); dP[k] = w.dpacked + packed_offset(k); }
    const float* fcw = d_packed + packed_offset(5);
    float* dfc = w.dpacked + packed_offset(5);
    int rc = 0;
    if (n == 0 || m == 0) {
        if (cudaMemsetAsync(d_dflat, 0, sizeof(float) * FLAT_TOTAL, s) != cudaSuccess) return cuda_status("mllp_gnn_backward: memset");
        return 0;
    }
    int forks = 0;
    // the stream the parameter-gradient sums of the conv just finished on `s` go to
    auto branch = [&]() -> cudaStream_t {
        if (!s2) return s;
        if (cudaEventRecord(ev[forks], s) != cudaSuccess || cudaStreamWaitEvent(s2, ev[forks], 0) != cudaSuccess) rc = cuda_status("mllp_gnn_backward: fork");
        ++forks;
        return s2;
    };
    // layer 3: gconv3_w2s (destination = variables) with the folded Linear(16, 1)
    rc = launch_bwd_dst<C>(*to_var, h1b, h2b, P[4], nullptr, d_dout, fcw, w.rec1, w.nr[4], w.d1b, w.hfc, s);
    if (rc == 0) {
        cudaStream_t sp = branch();
        if (rc == 0) rc = launch_param_grads<C>((int)n, h1b, w.nr[4], w.partial, dP[4], sp);
        if (rc == 0) {
            const int grid = (int)((n + 15) / 16 > PGRID ? PGRID : (n + 15) / 16);
            count_launch(2);
            k_gnn_fc_partial<<<grid, 256, 0, sp>>>((int)n, w.hfc, d_dout, w.partial);
            k_gnn_sum_parts<<<(C + 1 + 7) / 8, 256, 0, sp>>>(grid, C + 1, w.partial, dfc);
            rc = cuda_status("mllp_gnn_backward: fc");
        }
    }
    if (rc == 0) rc = launch_bwd_src(*to_con, h2b, w.rec1, w.d2b, 0, s);
    // layer 2: gconv2_w2s (variables <- constraints) and gconv2_s2w (constraints <- variables)
    if (rc == 0) rc = launch_bwd_dst<C>(*to_var, h1a, h2a, P[2], w.d1b, nullptr, nullptr, w.rec1, w.nr[2], w.d1a, nullptr, s);
    if (rc == 0) { cudaStream_t sp = branch(); if (rc == 0) rc = launch_param_grads<C>((int)n, h1a, w.nr[2], w.partial, dP[2], sp); }
    if (rc == 0) rc = launch_bwd_dst<C>(*to_con, h2a, h1a, P[3], w.d2b, nullptr, nullptr, w.rec2, w.nr[3], w.d2a, nullptr, s);
    if (rc == 0) { cudaStream_t sp = branch(); if (rc == 0) rc = launch_param_grads<C>((int)m, h2a, w.nr[3], w.partial, dP[3], sp); }
    if (rc == 0) rc = launch_bwd_src(*to_con, h2a, w.rec1, w.d2a, 1, s);
    if (rc == 0) rc = launch_bwd_src(*to_var, h1a, w.rec2, w.d1a, 1, s);
    // layer 1: the inputs need no gradient
    if (rc == 0) rc = launch_bwd_dst<1>(*to_var, d_x1, d_x2, P[0], w.d1a, nullptr, nullptr, w.rec1, w.nr[0], nullptr, nullptr, s);
    if (rc == 0) { cudaStream_t sp = branch(); if (rc == 0) rc = launch_param_grads<1>((int)n, d_x1, w.nr[0], w.partial, dP[0], sp); }
    if (rc == 0) rc = launch_bwd_dst<1>(*to_con, d_x2, d_x1, P[1], w.d2a, nullptr, nullptr, w.rec2, w.nr[1], nullptr, nullptr, s);
    if (rc == 0) { cudaStream_t sp = branch(); if (rc == 0) rc = launch_param_grads<1>((int)m, d_x2, w.nr[1], w.partial, dP[1], sp); }
    if (rc != 0) return rc;
    if (s2) {   // join: the unpack needs every conv's sums
        if (cudaEventRecord(ev[forks], s2) != cudaSuccess || cudaStreamWaitEvent(s, ev[forks], 0) != cudaSuccess)
            return cuda_status("mllp_gnn_backward: join");
    }
    count_launch(1);
    k_gnn_unpack_grads<<<7, 256, 0, s>>>(d_flat, w.dpacked, d_dflat);
    return cuda_status("mllp_gnn_backward: unpack");
}


extern "C" {

int mllp_gnn_backward(const mllp_gnn_side* to_var, const mllp_gnn_side* to_con, const float* d_x1, const float* d_x2,
                      const float* d_flat, const float* d_packed, const float* d_work, float* d_bwork, const float* d_dout,
                      float* d_dflat, void* stream)
{
    return backward_impl(to_var, to_con, d_x1, d_x2, d_flat, d_packed, d_work, d_bwork, d_dout, d_dflat, (cudaStream_t)stream,
                         "mllp_gnn_backward");
}

int mllp_gnn_backward_plan_create(const mllp_gnn_side* to_var, const mllp_gnn_side* to_con, const float* d_x1, const float* d_x2,
                                  const float* d_flat, float* d_packed, const float* d_work, float* d_bwork, const float* d_dout,
                                  float* d_dflat, mllp_gnn_plan_t* out)
{
    if (!out) return gfail(MLLP_E_INVALID, "mllp_gnn_backward_plan_create: null output");
    *out = nullptr;
    return capture_plan("mllp_gnn_backward_plan_create", out, [&](cudaStream_t s, cudaStream_t s2, cudaEvent_t* ev) {
        int rc = mllp_gnn_pack_params(d_flat, d_packed, s);
        if (rc == 0) rc = backward_impl(to_var, to_con, d_x1, d_x2, d_flat, d_packed, d_work, d_bwork, d_dout, d_dflat, s,
                                        "mllp_gnn_backward_plan_create", s2, ev);
        return rc;
    });
}

}  // extern "C"
