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
// fp32, as the reference (dtype=torch.float, :90-91, :100).  The layer is evaluated WITHOUT materialising q, k or v:
//
//     q_i . k_j = (Wk' q_i) . x_j + q_i . bk        the second term is the same for all edges of i: it cancels in
//                                                   the softmax, so the score of an edge is  qt_i . x_j + a_ij qe_i
//                                                   with qt_i = Wk'(Wq x_i + bq) / 4,  qe_i = We . (Wq x_i + bq) / 4
//     sum_j alpha_ij v_j = Wv (sum_j alpha_ij x_j) + bv
//
// so an edge gathers ONE feature row of its source node (64 B for 16 channels, 4 B in the first layer, instead of a
// 128 B {k | v} row), costs din + din FMAs, and the four dense maps are applied once per destination node in the
// prologue / epilogue of its row (din x din and 2 x din x 16 / lanes FMAs): no projection kernels, nothing of size
// nodes x 32 or nnz is written.  The products Wq'Wk, Wk'bq, Wq'We are formed on the host (mllp_b200/gnn.py, float64,
// rounded once).  Structure:
//   * every destination node (a CSR row) has a group of S lanes (S = 1 .. 32, chosen from the row lengths like the
//     lanes-per-row of the LP format: one lane per row when rows hold a handful of edges, as A' of the large Netlib
//     instances does); a lane owns every S-th edge of the row and folds it into its own
//     online-softmax state (running max, sum, din accumulators), two edges in flight; at the row end the group's
//     states are merged by a butterfly all-reduce and every lane finishes 16 / S output channels;
//   * rows longer than `chunk` edges (osa-60 has rows of 173 366 edges, ken-18 151 rows of ~300 among 105 127 of ~3)
//     are cut into items (one warp each) whose partial states are merged in a fixed order by a second kernel;
//   * the final Linear(16, 1) is folded into the epilogue of the last conv.
// HBM/L2-bound gather work (hidden = 16): no tensor cores.
#include "gnn_common.cuh"

namespace mllp {
namespace {

// rows with at most `chunk` edges: S lanes per row.  hout (nd x 16) and / or, with `fc` (= w[16] | b), the folded final
// linear layer fc_out[i] = w . out_i + b.
template <int S, int DIN>
__global__ void __launch_bounds__(256, 3) k_gnn_conv_rows(int nd, const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                                       const double* __restrict__ values, const float* __restrict__ hdst,
                                                       const float* __restrict__ hsrc, const float* __restrict__ prm_g,
                                                       float* __restrict__ hout, int chunk, int relu,
                                                       const float* __restrict__ fc, float* __restrict__ fc_out)
{
    using O = Off<DIN>;
    __shared__ __align__(16) float prm[O::total + C + 4];
    for (int k = threadIdx.x; k < O::total; k += blockDim.x) prm[k] = prm_g[k];
    if (fc && threadIdx.x <= C) prm[O::total + threadIdx.x] = fc[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, gl = lane & (S - 1);
    constexpr int RPW = 32 / S;                 // rows per warp
    constexpr int CNT = S <= C ? C / S : 1;     // output channels per lane (S = 32: lane pairs share a channel)
    const int c0 = S <= C ? gl * CNT : (gl >> 1);
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * RPW; base < nd; base += warps * RPW) {
        const int i = base + lane / S;   // warp-uniform trip count: the merges below are warp-wide
        int e0 = 0, e1 = 0;
        if (i < nd) { e0 = __ldg(indptr + i); e1 = __ldg(indptr + i + 1); }
        const bool live = i < nd && e1 - e0 <= chunk;   // long row: k_gnn_conv_items + k_gnn_conv_merge
        State<DIN> st;
        state_init<DIN>(st);
        {
            float x[DIN], qt[DIN], qe;
#pragma unroll
            for (int d = 0; d < DIN; ++d) x[d] = 0.0f;
            if (live) load_row<DIN>(hdst + (size_t)i * DIN, x);
            dst_prologue<S, DIN>(prm, x, gl, qt, qe);   // (warp-wide: the group's lanes exchange their parts)
            if (live && e1 > e0) edge_loop<S, DIN>(indices, values, hsrc, e0, e1, gl, qt, qe, st);
        }
        merge_group<S, DIN>(st);
        const bool any = st.l > 0.0f;
        const float inv = any ? 1.0f / st.l : 0.0f;   // a node without incoming edges keeps only the root term
        float x[DIN];   // read again (L1) rather than kept in registers across the edge loop
#pragma unroll
        for (int d = 0; d < DIN; ++d) x[d] = 0.0f;
        if (live) load_row<DIN>(hdst + (size_t)i * DIN, x);
        float part = 0.0f;
        const bool writer = live && !(S == 32 && (gl & 1));
        constexpr int STEP = CNT >= 4 ? 4 : CNT;   // channels finished together (a float4 of weights per input channel)
#pragma unroll
        for (int k0 = 0; k0 < CNT; k0 += STEP) {
            float o[STEP];
            out_channels<DIN, STEP>(prm, c0 + k0, x, st.acc, st.pa, inv, any, relu, o);
#pragma unroll
            for (int k = 0; k < STEP; ++k) part = fmaf(o[k], prm[O::total + c0 + k0 + k], part);
            if (writer && hout) {
                if constexpr (STEP == 4) {
                    *reinterpret_cast<float4*>(hout + (size_t)i * C + c0 + k0) = make_float4(o[0], o[1], o[2], o[3]);
                } else {
#pragma unroll
                    for (int k = 0; k < STEP; ++k) hout[(size_t)i * C + c0 + k0 + k] = o[k];
                }
            }
        }
        if (fc) {   // warp-uniform
            if (S == 32 && (gl & 1)) part = 0.0f;
#pragma unroll
            for (int w = S / 2; w > 0; w >>= 1) part += __shfl_xor_sync(FULLM, part, w);
            if (live && gl == 0) fc_out[i] = part + prm[O::total + C];
        }
    }
}

// long rows: merge the partial states of row r's items [first[r], first[r+1]) -- lane k folds items k, k + 32, ... in
// order, then the 32 lanes' states are merged by the butterfly (a fixed order: osa-60's longest row has 340 items, one
// after the other they were a 140 us chain of dependent loads) -- then the epilogue (lane c < 16 finishes channel c)
template <int DIN>
__global__ void __launch_bounds__(256) k_gnn_conv_merge(int nlong, const int32_t* __restrict__ long_rows,
                                                        const int32_t* __restrict__ first, const float* __restrict__ scratch,
                                                        const float* __restrict__ hdst, const float* __restrict__ prm,
                                                        float* __restrict__ hout, int relu, const float* __restrict__ fc,
                                                        float* __restrict__ fc_out)
{
    const int lane = threadIdx.x & 31, c = lane & 15;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < nlong; r += warps) {
        const int i = __ldg(long_rows + r);
        const int t1 = __ldg(first + r + 1);
        State<DIN> st;
        merge_items<DIN>(scratch, __ldg(first + r), t1, lane, st);
        const bool any = st.l > 0.0f;
        const float inv = any ? 1.0f / st.l : 0.0f;
        float x[DIN], o1[1];
        load_row<DIN>(hdst + (size_t)i * DIN, x);
        out_channels<DIN, 1>(prm, c, x, st.acc, st.pa, inv, any, relu, o1);
        if (lane < 16 && hout) hout[(size_t)i * C + c] = o1[0];
        if (fc) {
            float part = lane < 16 ? o1[0] * __ldg(fc + c) : 0.0f;
#pragma unroll
            for (int w = 16; w > 0; w >>= 1) part += __shfl_xor_sync(FULLM, part, w);
            if (lane == 0) fc_out[i] = part + __ldg(fc + C);
        }
    }
}

template <int DIN>
int launch_conv_din(const mllp_gnn_side& g, const float* hdst, const float* hsrc, const float* prm, float* hout, int relu,
                    const float* fc, float* fc_out, cudaStream_t s)
{
    // one wave of resident CTAs (the kernel walks its rows with a grid stride)
    const long long want = (((long long)g.nd * g.group + 31) / 32 + 7) / 8;
    count_launch(g.nlong > 0 ? 3 : 1);
#define MLLP_CONV_ROWS(SS)                                                                                                   \
    do {                                                                                                                     \
        static int per_sm = 0;                                                                                               \
        if (per_sm == 0 && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gnn_conv_rows<SS, DIN>, 256, 0) != cudaSuccess || per_sm < 1)) \
            per_sm = 1;                                                                                                      \
        const long long cap = 148LL * per_sm;                                                                                \
        const int grid = (int)(want < 1 ? 1 : want > cap ? cap : want);                                                      \
        k_gnn_conv_rows<SS, DIN><<<grid, 256, 0, s>>>(g.nd, g.indptr, g.indices, g.values, hdst, hsrc, prm, hout, g.chunk, relu, fc, fc_out); \
    } while (0)
    switch (g.group) {
        case 1: MLLP_CONV_ROWS(1); break;
        case 2: MLLP_CONV_ROWS(2); break;
        case 4: MLLP_CONV_ROWS(4); break;
        case 8: MLLP_CONV_ROWS(8); break;
        case 16: MLLP_CONV_ROWS(16); break;
        default: MLLP_CONV_ROWS(32); break;
    }
#undef MLLP_CONV_ROWS
    if (g.nlong > 0) {
        k_gnn_conv_items<DIN><<<grid_for_warps(g.nitems), 256, 0, s>>>(g.nitems, g.items, g.indices, g.values, hdst, hsrc, prm, g.scratch);
        k_gnn_conv_merge<DIN><<<grid_for_warps(g.nlong), 256, 0, s>>>(g.nlong, g.long_rows, g.long_first, g.scratch, hdst, prm, hout,
                                                                     relu, fc, fc_out);
    }
    return cuda_status("mllp_gnn: conv");
}

int launch_conv(const mllp_gnn_side& g, int din, const float* hdst, const float* hsrc, const float* prm, float* hout, int relu,
                const float* fc, float* fc_out, cudaStream_t s)
{
    if (g.nd == 0) return 0;
    if (din == 1) return launch_conv_din<1>(g, hdst, hsrc, prm, hout, relu, fc, fc_out, s);
    return launch_conv_din<C>(g, hdst, hsrc, prm, hout, relu, fc, fc_out, s);
}
}  // namespace
}  // namespace mllp

using namespace mllp;

extern "C" {

int64_t mllp_gnn_conv_param_floats(int32_t din)
{
    return din == 1 ? Off<1>::total : din == C ? Off<C>::total : -1;
}

int mllp_gnn_conv(const mllp_gnn_side* side, int32_t din, const float* d_hdst, const float* d_hsrc, const float* d_params,
                  float* d_hout, int32_t relu, void* stream)
{
    if (!side_ok(side) || (din != 1 && din != C) || !d_hdst || !d_hsrc || !d_params || !d_hout)
        return gfail(MLLP_E_INVALID, "mllp_gnn_conv: bad argument (din must be 1 or 16)");
    return launch_conv(*side, din, d_hdst, d_hsrc, d_params, d_hout, relu, nullptr, nullptr, (cudaStream_t)stream);
}

int64_t mllp_gnn_workspace_floats(int32_t n, int32_t m) { return 32 * ((int64_t)n + (int64_t)m) + 64; }

// The launches of one forward.  With a second stream the two convs of a layer (independent: both read the layer's
// input features, :241-246) run side by side: fork / join by events (used under stream capture by the plan).
static int forward_impl(const mllp_gnn_side* to_var, const mllp_gnn_side* to_con, const float* d_x1, const float* d_x2,
                        const float* d_params, float* d_work, float* d_out, cudaStream_t s, cudaStream_t s2, cudaEvent_t* ev)
{
    const int n = to_var->nd, m = to_con->nd;
    // workspace: two feature buffers per node set (16 B aligned: n and m are multiplied by 16 floats)
    float* h1[2] = {d_work, d_work + (size_t)16 * n};
    float* h2[2] = {d_work + (size_t)32 * n, d_work + (size_t)32 * n + (size_t)16 * m};
    // parameter blocks: gconv1_w2s, gconv1_s2w (din 1), gconv2_w2s, gconv2_s2w, gconv3_w2s (din 16), fc
    const float* P[5];
    size_t off = 0;
    for (int k = 0; k < 5; ++k) { P[k] = d_params + off; off += (size_t)(k < 2 ? Off<1>::total : Off<C>::total); }
    const float* fc = d_params + off;
    const float* in1 = d_x1;
    const float* in2 = d_x2;
    int rc = 0;
    // variables are the destination of the w2s convs and the source of the s2w convs
    for (int layer = 0; layer < 2 && rc == 0; ++layer) {
        const int din = layer == 0 ? 1 : C;
        cudaStream_t sb = s;
        if (s2) {
            if (cudaEventRecord(ev[2 * layer], s) != cudaSuccess || cudaStreamWaitEvent(s2, ev[2 * layer], 0) != cudaSuccess)
                return cuda_status("mllp_gnn: fork");
            sb = s2;
        }
        rc = launch_conv(*to_var, din, in1, in2, P[2 * layer], h1[layer], 1, nullptr, nullptr, s);
        if (rc == 0) rc = launch_conv(*to_con, din, in2, in1, P[2 * layer + 1], h2[layer], 1, nullptr, nullptr, sb);
        if (s2 && rc == 0) {
            if (cudaEventRecord(ev[2 * layer + 1], s2) != cudaSuccess || cudaStreamWaitEvent(s, ev[2 * layer + 1], 0) != cudaSuccess)
                return cuda_status("mllp_gnn: join");
        }
        in1 = h1[layer]; in2 = h2[layer];
    }
    if (rc == 0) rc = launch_conv(*to_var, C, in1, in2, P[4], nullptr, 1, fc, d_out, s);
    return rc;
}

static int forward_args_ok(const mllp_gnn_side* to_var, const mllp_gnn_side* to_con, const float* d_x1, const float* d_x2,
                           const float* d_params, float* d_work, float* d_out, const char* who)
{
    if (!side_ok(to_var) || !side_ok(to_con) || !d_x1 || !d_x2 || !d_params || !d_work || !d_out)
        return gfail(MLLP_E_INVALID, std::string(who) + ": bad argument");
    if (to_var->ns != to_con->nd || to_con->ns != to_var->nd)
        return gfail(MLLP_E_INVALID, std::string(who) + ": the two sides do not describe one graph");
    return 0;
}

int mllp_gnn_forward(const mllp_gnn_side* to_var, const mllp_gnn_side* to_con, const float* d_x1, const float* d_x2,
                     const float* d_params, float* d_work, float* d_out, void* stream)
{
    const int rc = forward_args_ok(to_var, to_con, d_x1, d_x2, d_params, d_work, d_out, "mllp_gnn_forward");
    if (rc != 0) return rc;
    return forward_impl(to_var, to_con, d_x1, d_x2, d_params, d_work, d_out, (cudaStream_t)stream, nullptr, nullptr);
}

int mllp_gnn_plan_create(const mllp_gnn_side* to_var, const mllp_gnn_side* to_con, const float* d_x1, const float* d_x2,
                         const float* d_params, float* d_work, float* d_out, mllp_gnn_plan_t* out)
{
    if (!out) return gfail(MLLP_E_INVALID, "mllp_gnn_plan_create: null output");
    *out = nullptr;
    int rc = forward_args_ok(to_var, to_con, d_x1, d_x2, d_params, d_work, d_out, "mllp_gnn_plan_create");
    if (rc != 0) return rc;
    return capture_plan("mllp_gnn_plan_create", out, [&](cudaStream_t s, cudaStream_t s2, cudaEvent_t* ev) {
        return forward_impl(to_var, to_con, d_x1, d_x2, d_params, d_work, d_out, s, s2, ev);
    });
}

int mllp_gnn_train_plan_create(const mllp_gnn_side* to_var, const mllp_gnn_side* to_con, const float* d_x1, const float* d_x2,
                               const float* d_flat, float* d_packed, float* d_work, float* d_out, mllp_gnn_plan_t* out)
{
    if (!out) return gfail(MLLP_E_INVALID, "mllp_gnn_train_plan_create: null output");
    *out = nullptr;
    int rc = forward_args_ok(to_var, to_con, d_x1, d_x2, d_packed, d_work, d_out, "mllp_gnn_train_plan_create");
    if (rc != 0) return rc;
    if (!d_flat) return gfail(MLLP_E_INVALID, "mllp_gnn_train_plan_create: null parameters");
    return capture_plan("mllp_gnn_train_plan_create", out, [&](cudaStream_t s, cudaStream_t s2, cudaEvent_t* ev) {
        int r = mllp_gnn_pack_params(d_flat, d_packed, s);
        if (r == 0) r = forward_impl(to_var, to_con, d_x1, d_x2, d_packed, d_work, d_out, s, s2, ev);
        return r;
    });
}

int mllp_gnn_plan_run(mllp_gnn_plan_t plan, void* stream)
{
    if (!plan || !plan->exec) return gfail(MLLP_E_INVALID, "mllp_gnn_plan_run: null plan");
    mllp::count_launch(1);   // one graph launch (the kernels inside were counted when the plan was captured)
    const cudaError_t e = cudaGraphLaunch(plan->exec, (cudaStream_t)stream);
    if (e != cudaSuccess) return gfail((int)e, std::string("mllp_gnn_plan_run: ") + cudaGetErrorString(e));
    return 0;
}

int mllp_gnn_plan_destroy(mllp_gnn_plan_t plan)
{
    if (!plan) return 0;
    if (plan->exec) cudaGraphExecDestroy(plan->exec);
    delete plan;
    return 0;
}

}  // extern "C"
