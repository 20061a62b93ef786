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
    static constexpr int xbar = 0, dqt = DIN, g = 2 * DIN, abar = g + C, dqe = abar + 1, any = abar + 2, total = (any + 1 + 3) & ~3;
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
constexpr int PGRID = 148 * 2;   // CTAs of the parameter-gradient partial sums

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
__device__ __forceinline__ void g_products(const float* prm, int c0, const float* g, float* dxb, float* dxs, float& dab)
{
    using O = Off<DIN>;
#pragma unroll
    for (int k = 0; k < CNT; ++k) {
        const float gc = g[k];
        dab = fmaf(prm[O::we + c0 + k], gc, dab);
#pragma unroll
        for (int d = 0; d < DIN; ++d) {
            dxb[d] = fmaf(prm[O::wv + d * C + c0 + k], gc, dxb[d]);
            dxs[d] = fmaf(prm[O::ws + d * C + c0 + k], gc, dxs[d]);
        }
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

template <int DIN>
__device__ __forceinline__ void store_rec(float* r, const float* qt, const float* dxb, float qe, float lse, float dab, float D)
{
    using R = Rec<DIN>;
    if constexpr (DIN % 4 == 0) {
        float4* q = reinterpret_cast<float4*>(r);
#pragma unroll
        for (int k = 0; k < DIN / 4; ++k) q[k] = make_float4(qt[4 * k], qt[4 * k + 1], qt[4 * k + 2], qt[4 * k + 3]);
#pragma unroll
        for (int k = 0; k < DIN / 4; ++k) q[DIN / 4 + k] = make_float4(dxb[4 * k], dxb[4 * k + 1], dxb[4 * k + 2], dxb[4 * k + 3]);
        q[DIN / 2] = make_float4(qe, lse, dab, D);
    } else {
#pragma unroll
        for (int d = 0; d < DIN; ++d) { r[R::qt + d] = qt[d]; r[R::dxb + d] = dxb[d]; }
        r[R::qe] = qe; r[R::lse] = lse; r[R::dab] = dab; r[R::dd] = D;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// destination pass, rows of at most `chunk` edges: S lanes per row.
//   upstream gradient: gh[nd][16] (d loss / d relu(out)), or for the last conv dout[nd] and fcw[16] (the folded Linear(16, 1));
//   rec / dxdst may be null (first layer: the inputs need no gradient); hfc[nd][16] = dout_i relu(out_i) for d fc.weight.
template <int S, int DIN>
__global__ void __launch_bounds__(256, 1) k_gnn_bwd_dst(int nd, const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                                     const double* __restrict__ values, const float* __restrict__ hdst,
                                                     const float* __restrict__ hsrc, const float* __restrict__ prm_g,
                                                     const float* __restrict__ gh, const float* __restrict__ dout,
                                                     const float* __restrict__ fcw_g, int chunk, float* __restrict__ rec,
                                                     float* __restrict__ nr, float* __restrict__ dxdst, float* __restrict__ hfc)
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
        // this lane's channels: pre-activation output, ReLU mask, upstream gradient
        const float di = (dout && live) ? __ldg(dout + i) : 0.0f;
        float g[CNT];
        constexpr int STEP = CNT >= 4 ? 4 : CNT;
#pragma unroll
        for (int k0 = 0; k0 < CNT; k0 += STEP) {
            float o[STEP];
            out_channels<DIN, STEP>(prm, c0 + k0, x, st.acc, st.pa, inv, any, 0, o);
#pragma unroll
            for (int k = 0; k < STEP; ++k) {
                const int c = c0 + k0 + k;
                float up = 0.0f;
                if (live && owner) up = gh ? __ldg(gh + (size_t)i * C + c) : di * prm[O::total + c];
                g[k0 + k] = o[k] > 0.0f ? up : 0.0f;
                if (live && owner) {
                    nr[(size_t)i * N::total + N::g + c] = g[k0 + k];
                    if (hfc) hfc[(size_t)i * C + c] = di * fmaxf(o[k], 0.0f);
                }
            }
        }
        float dxb[DIN], dxs[DIN], dab = 0.0f;
#pragma unroll
        for (int d = 0; d < DIN; ++d) { dxb[d] = 0.0f; dxs[d] = 0.0f; }
        g_products<DIN, CNT>(prm, c0, g, dxb, dxs, dab);
#pragma unroll
        for (int o = S / 2; o > 0; o >>= 1) {
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
        if (live && gl == 0) {
            float* r = nr + (size_t)i * N::total;
#pragma unroll
            for (int d = 0; d < DIN; ++d) r[N::xbar + d] = st.acc[d] * inv;
            r[N::abar] = abar;
            r[N::any] = any ? 1.0f : 0.0f;
            if (rec) store_rec<DIN>(rec + (size_t)i * Rec<DIN>::total, qt, dxb, qe, lse, dab, D);
        }
        // second sweep over the row's edges
        float dqt[DIN], dqe = 0.0f;
#pragma unroll
        for (int d = 0; d < DIN; ++d) dqt[d] = 0.0f;
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
#pragma unroll
        for (int o = S / 2; o > 0; o >>= 1) {
            dqe += __shfl_xor_sync(FULLM, dqe, o);
#pragma unroll
            for (int d = 0; d < DIN; ++d) dqt[d] += __shfl_xor_sync(FULLM, dqt[d], o);
        }
        if (live && gl == 0) {
            float* r = nr + (size_t)i * N::total;
#pragma unroll
            for (int d = 0; d < DIN; ++d) r[N::dqt + d] = dqt[d];
            r[N::dqe] = dqe;
            if (dxdst) store_dx_dst<DIN>(prm, dxs, dqt, dqe, dxdst + (size_t)i * DIN);
        }
    }
}

// destination pass, rows above `chunk` edges: one CTA per row, the warps' partial states / sums combined in shared
// memory in a fixed order
template <int DIN>
__global__ void __launch_bounds__(256) k_gnn_bwd_dst_long(int nlong, const int32_t* __restrict__ long_rows, const int32_t* __restrict__ indptr,
                                                          const int32_t* __restrict__ indices, const double* __restrict__ values,
                                                          const float* __restrict__ hdst, const float* __restrict__ hsrc,
                                                          const float* __restrict__ prm_g, const float* __restrict__ gh,
                                                          const float* __restrict__ dout, const float* __restrict__ fcw_g,
                                                          float* __restrict__ rec, float* __restrict__ nr, float* __restrict__ dxdst,
                                                          float* __restrict__ hfc)
{
    using O = Off<DIN>;
    using N = NR<DIN>;
    __shared__ __align__(16) float prm[O::total + C];
    __shared__ float red[8][DIN + 4];
    __shared__ float gsh[C];
    for (int k = threadIdx.x; k < O::total; k += blockDim.x) prm[k] = prm_g[k];
    if (fcw_g && threadIdx.x < C) prm[O::total + threadIdx.x] = fcw_g[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = blockIdx.x; r < nlong; r += gridDim.x) {
        const int i = __ldg(long_rows + r);
        const int e0 = __ldg(indptr + i), e1 = __ldg(indptr + i + 1);
        float x[DIN], qt[DIN], qe;
        load_row<DIN>(hdst + (size_t)i * DIN, x);
        dst_prologue<1, DIN>(prm, x, 0, qt, qe);
        State<DIN> st;
        state_init<DIN>(st);
        for (int e = e0 + (int)threadIdx.x; e < e1; e += 256) {
            const int j = __ldg(indices + e);
            const float a = (float)__ldg(values + e);
            float xj[DIN];
            load_row<DIN>(hsrc + (size_t)j * DIN, xj);
            float s = a * qe;
#pragma unroll
            for (int d = 0; d < DIN; ++d) s = fmaf(qt[d], xj[d], s);
            fold<DIN>(st, s, a, xj);
        }
        merge_group<32, DIN>(st);
        if (lane == 0) {
            red[warp][0] = st.m; red[warp][1] = st.l; red[warp][2] = st.pa;
#pragma unroll
            for (int d = 0; d < DIN; ++d) red[warp][4 + d] = st.acc[d];
        }
        __syncthreads();
        state_init<DIN>(st);
        for (int w = 0; w < 8; ++w) {
            const float mw = red[w][0];
            const float mn = fmaxf(st.m, mw);
            const float s1 = st.m == -INFINITY ? 0.0f : __expf(st.m - mn), s2 = mw == -INFINITY ? 0.0f : __expf(mw - mn);
            st.l = st.l * s1 + red[w][1] * s2;
            st.pa = st.pa * s1 + red[w][2] * s2;
#pragma unroll
            for (int d = 0; d < DIN; ++d) st.acc[d] = st.acc[d] * s1 + red[w][4 + d] * s2;
            st.m = mn;
        }
        const bool any = st.l > 0.0f;
        const float inv = any ? 1.0f / st.l : 0.0f;
        const float lse = any ? st.m + __logf(st.l) : 0.0f;
        const float di = dout ? __ldg(dout + i) : 0.0f;
        if (threadIdx.x < C) {
            const int c = threadIdx.x;
            float o1[1];
            out_channels<DIN, 1>(prm, c, x, st.acc, st.pa, inv, any, 0, o1);
            const float up = gh ? __ldg(gh + (size_t)i * C + c) : di * prm[O::total + c];
            const float gc = o1[0] > 0.0f ? up : 0.0f;
            gsh[c] = gc;
            nr[(size_t)i * N::total + N::g + c] = gc;
            if (hfc) hfc[(size_t)i * C + c] = di * fmaxf(o1[0], 0.0f);
        }
        __syncthreads();   // gsh ready; red free again
        float g[C], dxb[DIN], dxs[DIN], dab = 0.0f;
#pragma unroll
        for (int c = 0; c < C; ++c) g[c] = gsh[c];
#pragma unroll
        for (int d = 0; d < DIN; ++d) { dxb[d] = 0.0f; dxs[d] = 0.0f; }
        g_products<DIN, C>(prm, 0, g, dxb, dxs, dab);
        const float abar = st.pa * inv;
        float D = dab * abar;
#pragma unroll
        for (int d = 0; d < DIN; ++d) D = fmaf(dxb[d], st.acc[d] * inv, D);
        if (threadIdx.x == 0) {
            float* q = nr + (size_t)i * N::total;
#pragma unroll
            for (int d = 0; d < DIN; ++d) q[N::xbar + d] = st.acc[d] * inv;
            q[N::abar] = abar;
            q[N::any] = any ? 1.0f : 0.0f;
            if (rec) store_rec<DIN>(rec + (size_t)i * Rec<DIN>::total, qt, dxb, qe, lse, dab, D);
        }
        float dqt[DIN], dqe = 0.0f;
#pragma unroll
        for (int d = 0; d < DIN; ++d) dqt[d] = 0.0f;
        for (int e = e0 + (int)threadIdx.x; e < e1; e += 256) {
            const int j = __ldg(indices + e);
            const float a = (float)__ldg(values + e);
            float xj[DIN];
            load_row<DIN>(hsrc + (size_t)j * DIN, xj);
            edge_ds<DIN>(a, xj, qt, qe, lse, dxb, dab, D, dqt, dqe);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            dqe += __shfl_xor_sync(FULLM, dqe, o);
#pragma unroll
            for (int d = 0; d < DIN; ++d) dqt[d] += __shfl_xor_sync(FULLM, dqt[d], o);
        }
        if (lane == 0) {
            red[warp][0] = dqe;
#pragma unroll
            for (int d = 0; d < DIN; ++d) red[warp][4 + d] = dqt[d];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            dqe = 0.0f;
#pragma unroll
            for (int d = 0; d < DIN; ++d) dqt[d] = 0.0f;
            for (int w = 0; w < 8; ++w) {
                dqe += red[w][0];
#pragma unroll
                for (int d = 0; d < DIN; ++d) dqt[d] += red[w][4 + d];
            }
            float* q = nr + (size_t)i * N::total;
#pragma unroll
            for (int d = 0; d < DIN; ++d) q[N::dqt + d] = dqt[d];
            q[N::dqe] = dqe;
            if (dxdst) store_dx_dst<DIN>(prm, dxs, dqt, dqe, dxdst + (size_t)i * DIN);
        }
        __syncthreads();
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

template <int DIN>
__global__ void __launch_bounds__(256) k_gnn_bwd_src_long(int nlong, const int32_t* __restrict__ long_rows, const int32_t* __restrict__ indptr,
                                                          const int32_t* __restrict__ indices, const double* __restrict__ values,
                                                          const float* __restrict__ hsrc, const float* __restrict__ rec,
                                                          float* __restrict__ dxsrc, int accumulate)
{
    __shared__ float red[8][DIN];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = blockIdx.x; r < nlong; r += gridDim.x) {
        const int j = __ldg(long_rows + r);
        const int e0 = __ldg(indptr + j), e1 = __ldg(indptr + j + 1);
        float xj[DIN], dx[DIN];
        load_row<DIN>(hsrc + (size_t)j * DIN, xj);
#pragma unroll
        for (int d = 0; d < DIN; ++d) dx[d] = 0.0f;
        for (int e = e0 + (int)threadIdx.x; e < e1; e += 256)
            edge_src<DIN>((float)__ldg(values + e), rec + (size_t)__ldg(indices + e) * Rec<DIN>::total, xj, dx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int d = 0; d < DIN; ++d) dx[d] += __shfl_xor_sync(FULLM, dx[d], o);
        }
        if (lane == 0) {
#pragma unroll
            for (int d = 0; d < DIN; ++d) red[warp][d] = dx[d];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
#pragma unroll
            for (int d = 0; d < DIN; ++d) dx[d] = 0.0f;
            for (int w = 0; w < 8; ++w) {
#pragma unroll
                for (int d = 0; d < DIN; ++d) dx[d] += red[w][d];
            }
            store_dx_src<DIN>(dxsrc + (size_t)j * DIN, dx, accumulate);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// parameter gradients of one conv (layout of the fused block): every entry is sum_i U_i[a] V_i[b] with
//   U_i = {x_i[DIN], xbar_i[DIN], abar_i, any_i, 1},  V_i = {dqt_i[DIN], dqe_i, g_i[16]}
template <int DIN>
__global__ void __launch_bounds__(256) k_gnn_param_partial(int nd, const float* __restrict__ hdst, const float* __restrict__ nr,
                                                           float* __restrict__ partial)
{
    using O = Off<DIN>;
    using N = NR<DIN>;
    constexpr int UW = 2 * DIN + 3, VW = DIN + 1 + C, TN = 32;
    constexpr int EPT = (O::total + 255) / 256;
    constexpr int U_ABAR = 2 * DIN, U_ANY = 2 * DIN + 1, U_ONE = 2 * DIN + 2, V_DQE = DIN, V_G = DIN + 1;
    __shared__ float U[TN][UW + 1], V[TN][VW + 1];
    int ua[EPT], vb[EPT];
    float acc[EPT];
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        const int e = threadIdx.x + 256 * k;
        int a = -1, b = 0;
        if (e < O::vq) { a = e / DIN; b = e % DIN; }
        else if (e < O::wq) { a = U_ONE; b = e - O::vq; }
        else if (e < O::sq) { a = e - O::wq; b = V_DQE; }
        else if (e == O::sq) { a = U_ONE; b = V_DQE; }
        else if (e < O::wv) { a = -1; }
        else if (e < O::bv) { a = DIN + (e - O::wv) / C; b = V_G + (e - O::wv) % C; }
        else if (e < O::ws) { a = U_ANY; b = V_G + e - O::bv; }
        else if (e < O::bs) { a = (e - O::ws) / C; b = V_G + (e - O::ws) % C; }
        else if (e < O::we) { a = U_ONE; b = V_G + e - O::bs; }
        else if (e < O::total) { a = U_ABAR; b = V_G + e - O::we; }
        ua[k] = a; vb[k] = b; acc[k] = 0.0f;
    }
    const int ntiles = (nd + TN - 1) / TN;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < TN * UW; idx += 256) {
            const int r = idx / UW, q = idx % UW;
            const int node = tile * TN + r;
            float v = 0.0f;
            if (node < nd) {
                const float* p = nr + (size_t)node * N::total;
                if (q < DIN) v = __ldg(hdst + (size_t)node * DIN + q);
                else if (q < 2 * DIN) v = p[N::xbar + q - DIN];
                else if (q == U_ABAR) v = p[N::abar];
                else if (q == U_ANY) v = p[N::any];
                else v = 1.0f;
            }
            U[r][q] = v;
        }
        for (int idx = threadIdx.x; idx < TN * VW; idx += 256) {
            const int r = idx / VW, q = idx % VW;
            const int node = tile * TN + r;
            float v = 0.0f;
            if (node < nd) {
                const float* p = nr + (size_t)node * N::total;
                v = q < DIN ? p[N::dqt + q] : q == V_DQE ? p[N::dqe] : p[N::g + q - V_G];
            }
            V[r][q] = v;
        }
        __syncthreads();
#pragma unroll 4
        for (int r = 0; r < TN; ++r) {
#pragma unroll
            for (int k = 0; k < EPT; ++k)
                if (ua[k] >= 0) acc[k] = fmaf(U[r][ua[k]], V[r][vb[k]], acc[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        const int e = threadIdx.x + 256 * k;
        if (e < O::total) partial[(size_t)blockIdx.x * O::total + e] = acc[k];
    }
}

// out[e] = sum over the parts, in order
__global__ void k_gnn_sum_parts(int nparts, int width, const float* __restrict__ partial, float* __restrict__ out)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= width) return;
    float s = 0.0f;
    for (int b = 0; b < nparts; ++b) s += partial[(size_t)b * width + e];
    out[e] = s;
}

// d fc.weight[c] = sum_i hfc[i][c], d fc.bias = sum_i dout[i]: per-CTA partials (17 floats)
__global__ void __launch_bounds__(256) k_gnn_fc_partial(int n, const float* __restrict__ hfc, const float* __restrict__ dout,
                                                        float* __restrict__ partial)
{
    __shared__ float sm[16][C + 1];
    const int r = threadIdx.x >> 4, c = threadIdx.x & 15;
    float a = 0.0f, b = 0.0f;
    for (int i = blockIdx.x * 16 + r; i < n; i += gridDim.x * 16) {
        a += hfc[(size_t)i * C + c];
        if (c == 0) b += dout[i];
    }
    sm[r][c] = a;
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
template <class K>
int resident_grid(K kernel, long long want)
{
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    const long long cap = 148LL * per_sm;
    return (int)(want < 1 ? 1 : want > cap ? cap : want);
}

template <int DIN>
int launch_bwd_dst(const mllp_gnn_side& g, const float* hdst, const float* hsrc, const float* prm, const float* gh, const float* dout,
                   const float* fcw, float* rec, float* nr, float* dxdst, float* hfc, cudaStream_t s)
{
    const long long want = (((long long)g.nd * g.group + 31) / 32 + 7) / 8;
    count_launch(g.nlong > 0 ? 2 : 1);
#define MLLP_BWD_DST(SS)                                                                                                              \
    k_gnn_bwd_dst<SS, DIN><<<resident_grid(k_gnn_bwd_dst<SS, DIN>, want), 256, 0, s>>>(g.nd, g.indptr, g.indices, g.values, hdst, hsrc, prm, \
                                                                                       gh, dout, fcw, g.chunk, rec, nr, dxdst, hfc)
    switch (g.group) {
        case 1: MLLP_BWD_DST(1); break;
        case 2: MLLP_BWD_DST(2); break;
        case 4: MLLP_BWD_DST(4); break;
        case 8: MLLP_BWD_DST(8); break;
        case 16: MLLP_BWD_DST(16); break;
        default: MLLP_BWD_DST(32); break;
    }
#undef MLLP_BWD_DST
    if (g.nlong > 0)
        k_gnn_bwd_dst_long<DIN><<<g.nlong < 148 * 4 ? g.nlong : 148 * 4, 256, 0, s>>>(g.nlong, g.long_rows, g.indptr, g.indices, g.values, hdst,
                                                                                      hsrc, prm, gh, dout, fcw, rec, nr, dxdst, hfc);
    return cuda_status("mllp_gnn_backward: destination pass");
}

// `t` = the transposed structure: rows are this conv's source nodes
int launch_bwd_src(const mllp_gnn_side& t, const float* hsrc, const float* rec, float* dxsrc, int accumulate, cudaStream_t s)
{
    const long long want = (((long long)t.nd * t.group + 31) / 32 + 7) / 8;
    count_launch(t.nlong > 0 ? 2 : 1);
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
    if (t.nlong > 0)
        k_gnn_bwd_src_long<C><<<t.nlong < 148 * 4 ? t.nlong : 148 * 4, 256, 0, s>>>(t.nlong, t.long_rows, t.indptr, t.indices, t.values, hsrc, rec,
                                                                                    dxsrc, accumulate);
    return cuda_status("mllp_gnn_backward: source pass");
}

template <int DIN>
int launch_param_grads(int nd, const float* hdst, const float* nr, float* partial, float* dpacked, cudaStream_t s)
{
    const int ntiles = (nd + 31) / 32;
    const int grid = ntiles < 1 ? 1 : ntiles > PGRID ? PGRID : ntiles;
    count_launch(2);
    k_gnn_param_partial<DIN><<<grid, 256, 0, s>>>(nd, hdst, nr, partial);
    k_gnn_sum_parts<<<(Off<DIN>::total + 255) / 256, 256, 0, s>>>(grid, Off<DIN>::total, partial, dpacked);
    return cuda_status("mllp_gnn_backward: parameter gradients");
}

struct BwdWork {
    float *d1b, *d1a, *d2b, *d2a, *rec1, *rec2, *nr, *hfc, *partial, *dpacked;
};
size_t align4(size_t v) { return (v + 3) & ~(size_t)3; }
size_t carve(BwdWork& w, float* base, size_t n, size_t m)
{
    const size_t mx = n > m ? n : m;
    size_t off = 0;
    auto take = [&](size_t cnt) { float* p = base ? base + off : nullptr; off += align4(cnt); return p; };
    w.d1b = take(16 * n); w.d1a = take(16 * n); w.d2b = take(16 * m); w.d2a = take(16 * m);
    w.rec1 = take(Rec<C>::total * n); w.rec2 = take(Rec<C>::total * m);
    w.nr = take(NR<C>::total * mx); w.hfc = take(16 * n);
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

int mllp_gnn_backward(const mllp_gnn_side* to_var, const mllp_gnn_side* to_con, const float* d_x1, const float* d_x2,
                      const float* d_flat, const float* d_packed, const float* d_work, float* d_bwork, const float* d_dout,
                      float* d_dflat, void* stream)
{
    if (!side_ok(to_var) || !side_ok(to_con) || !d_x1 || !d_x2 || !d_flat || !d_packed || !d_work || !d_bwork || !d_dout || !d_dflat)
        return gfail(MLLP_E_INVALID, "mllp_gnn_backward: bad argument");
    if (to_var->ns != to_con->nd || to_con->ns != to_var->nd)
        return gfail(MLLP_E_INVALID, "mllp_gnn_backward: the two sides do not describe one graph");
    cudaStream_t s = (cudaStream_t)stream;
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
    // layer 3: gconv3_w2s (destination = variables) with the folded Linear(16, 1)
    rc = launch_bwd_dst<C>(*to_var, h1b, h2b, P[4], nullptr, d_dout, fcw, w.rec1, w.nr, w.d1b, w.hfc, s);
    if (rc == 0) rc = launch_param_grads<C>((int)n, h1b, w.nr, w.partial, dP[4], s);
    if (rc == 0) {
        const int grid = (int)((n + 15) / 16 > PGRID ? PGRID : (n + 15) / 16);
        count_launch(2);
        k_gnn_fc_partial<<<grid, 256, 0, s>>>((int)n, w.hfc, d_dout, w.partial);
        k_gnn_sum_parts<<<1, 32, 0, s>>>(grid, C + 1, w.partial, dfc);
        rc = cuda_status("mllp_gnn_backward: fc");
    }
    if (rc == 0) rc = launch_bwd_src(*to_con, h2b, w.rec1, w.d2b, 0, s);
    // layer 2: gconv2_w2s (variables <- constraints) and gconv2_s2w (constraints <- variables)
    if (rc == 0) rc = launch_bwd_dst<C>(*to_var, h1a, h2a, P[2], w.d1b, nullptr, nullptr, w.rec1, w.nr, w.d1a, nullptr, s);
    if (rc == 0) rc = launch_param_grads<C>((int)n, h1a, w.nr, w.partial, dP[2], s);
    if (rc == 0) rc = launch_bwd_dst<C>(*to_con, h2a, h1a, P[3], w.d2b, nullptr, nullptr, w.rec2, w.nr, w.d2a, nullptr, s);
    if (rc == 0) rc = launch_param_grads<C>((int)m, h2a, w.nr, w.partial, dP[3], s);
    if (rc == 0) rc = launch_bwd_src(*to_con, h2a, w.rec1, w.d2a, 1, s);
    if (rc == 0) rc = launch_bwd_src(*to_var, h1a, w.rec2, w.d1a, 1, s);
    // layer 1: the inputs need no gradient
    if (rc == 0) rc = launch_bwd_dst<1>(*to_var, d_x1, d_x2, P[0], w.d1a, nullptr, nullptr, nullptr, w.nr, nullptr, nullptr, s);
    if (rc == 0) rc = launch_param_grads<1>((int)n, d_x1, w.nr, w.partial, dP[0], s);
    if (rc == 0) rc = launch_bwd_dst<1>(*to_con, d_x2, d_x1, P[1], w.d2a, nullptr, nullptr, nullptr, w.nr, nullptr, nullptr, s);
    if (rc == 0) rc = launch_param_grads<1>((int)m, d_x2, w.nr, w.partial, dP[1], s);
    if (rc != 0) return rc;
    count_launch(1);
    k_gnn_unpack_grads<<<7, 256, 0, s>>>(d_flat, w.dpacked, d_dflat);
    return cuda_status("mllp_gnn_backward: unpack");
}

}  // extern "C"
