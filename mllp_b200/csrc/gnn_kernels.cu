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
// fp32, as the reference (dtype=torch.float, :90-91, :100).  One warp owns one destination node (a CSR row):
// lanes 0-15 carry the key channels, lanes 16-31 the value channels, so ONE coalesced 128-byte load fetches a
// source node's {k, v}; the softmax is computed online (running max / sum), so every edge is visited once and
// nothing of size nnz is written.  Rows longer than a chunk (osa-60 has rows of 173 366 edges) are cut into
// chunks (one warp each) whose partial (max, sum, acc) triples are merged in a fixed order by a second kernel.
// HBM-bound gather work (hidden = 16): no tensor cores.
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "../../include/mllp_b200.h"

namespace mllp {
void set_last_error(const std::string& msg);

namespace {
constexpr int C = 16;              // channels
constexpr unsigned FULLM = 0xffffffffu;

// packed parameters of one TransformerConv as the kernels read them (floats):
//   conv: Wq[din][16] | bq[16] | Ws[din][16] | bs[16] | We[16]        (input-major: conflict-free, coalesced)
//   proj: Wk[din][16] | bk[16] | Wv[din][16] | bv[16]
struct Partial { float m, l, acc; };   // per lane: running max, running sum, accumulator of its channel

// q_c and skip_c of destination node i for this lane's channel, and qe = q . We (all lanes)
__device__ __forceinline__ void row_prologue(const float* __restrict__ hdst, int din, const float* __restrict__ sp, int i, int c,
                                             float& q, float& skip, float& qe)
{
    const float* Wq = sp;
    const float* bq = sp + din * C;
    const float* Ws = bq + C;
    const float* bs = Ws + din * C;
    const float* We = bs + C;
    q = bq[c];
    skip = bs[c];
    for (int d = 0; d < din; ++d) {
        const float h = __ldg(hdst + (size_t)i * din + d);
        q = fmaf(h, Wq[d * C + c], q);
        skip = fmaf(h, Ws[d * C + c], skip);
    }
    float t = q * We[c];
    for (int o = 8; o > 0; o >>= 1) t += __shfl_xor_sync(FULLM, t, o);   // both halves hold the same 16 channels
    qe = t;
}

// edges [e0, e1) of one destination node, online softmax.  Lanes < 16: key channel c, lanes >= 16: value channel c.
__device__ __forceinline__ Partial edge_loop(const int32_t* __restrict__ indices, const double* __restrict__ values,
                                             const float* __restrict__ kv, int e0, int e1, float q, float qe, float we, int lane)
{
    Partial P{-INFINITY, 0.0f, 0.0f};
    for (int base = e0; base < e1; base += 32) {
        const int e = base + lane;
        const int j_l = e < e1 ? __ldg(indices + e) : 0;
        const float a_l = e < e1 ? (float)__ldg(values + e) : 0.0f;   // edge_attr = float32(a_ij), as the reference casts it
        const int cnt = min(32, e1 - base);
        // the {k, v} rows of up to 4 edges are in flight together
        for (int u0 = 0; u0 < cnt; u0 += 4) {
            float x[4], a[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int src = __shfl_sync(FULLM, j_l, (u0 + u) & 31);
                a[u] = __shfl_sync(FULLM, a_l, (u0 + u) & 31);
                x[u] = u0 + u < cnt ? __ldg(kv + (size_t)src * 32 + lane) : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (u0 + u >= cnt) break;
                float t = lane < 16 ? q * x[u] : 0.0f;
                for (int o = 8; o > 0; o >>= 1) t += __shfl_xor_sync(FULLM, t, o);
                const float s = (__shfl_sync(FULLM, t, 0) + a[u] * qe) * 0.25f;   // / sqrt(16)
                const float mn = fmaxf(P.m, s);
                const float sc = expf(P.m - mn);   // exp(-inf) = 0 on the first edge
                const float p = expf(s - mn);
                P.l = P.l * sc + p;
                P.acc = P.acc * sc + p * (x[u] + a[u] * we);
                P.m = mn;
            }
        }
    }
    return P;
}

__device__ __forceinline__ void row_epilogue(float* __restrict__ hout, int i, int lane, const Partial& P, float skip, int relu)
{
    float o = (P.l > 0.0f ? P.acc / P.l : 0.0f) + skip;   // a node without incoming edges keeps only the root term
    if (relu) o = fmaxf(o, 0.0f);
    if (lane >= 16) hout[(size_t)i * C + (lane - 16)] = o;
}

// rows with at most `chunk` edges: one warp per row
__global__ void __launch_bounds__(256) k_gnn_conv_rows(int nd, const int32_t* __restrict__ indptr,
                                                       const int32_t* __restrict__ indices, const double* __restrict__ values,
                                                       const float* __restrict__ hdst, int din, const float* __restrict__ kv,
                                                       const float* __restrict__ params, float* __restrict__ hout, int chunk,
                                                       int relu)
{
    extern __shared__ float sp[];
    const int np = 2 * din * C + 3 * C;
    for (int k = threadIdx.x; k < np; k += blockDim.x) sp[k] = params[k];
    __syncthreads();
    const int lane = threadIdx.x & 31, c = lane & 15;
    const float we = sp[2 * din * C + 2 * C + c];
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < nd; i += warps) {
        const int e0 = __ldg(indptr + i), e1 = __ldg(indptr + i + 1);
        if (e1 - e0 > chunk) continue;   // long row: k_gnn_conv_items + k_gnn_conv_merge
        float q, skip, qe;
        row_prologue(hdst, din, sp, i, c, q, skip, qe);
        const Partial P = edge_loop(indices, values, kv, e0, e1, q, qe, we, lane);
        row_epilogue(hout, i, lane, P, skip, relu);
    }
}

// long rows: item t = (row, first edge, last edge); one warp per item, partial -> scratch[t][3][32]
__global__ void __launch_bounds__(256) k_gnn_conv_items(int nitems, const int32_t* __restrict__ items,
                                                        const int32_t* __restrict__ indices, const double* __restrict__ values,
                                                        const float* __restrict__ hdst, int din, const float* __restrict__ kv,
                                                        const float* __restrict__ params, float* __restrict__ scratch)
{
    extern __shared__ float sp[];
    const int np = 2 * din * C + 3 * C;
    for (int k = threadIdx.x; k < np; k += blockDim.x) sp[k] = params[k];
    __syncthreads();
    const int lane = threadIdx.x & 31, c = lane & 15;
    const float we = sp[2 * din * C + 2 * C + c];
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < nitems; t += warps) {
        const int i = __ldg(items + 3 * t), e0 = __ldg(items + 3 * t + 1), e1 = __ldg(items + 3 * t + 2);
        float q, skip, qe;
        row_prologue(hdst, din, sp, i, c, q, skip, qe);
        const Partial P = edge_loop(indices, values, kv, e0, e1, q, qe, we, lane);
        float* o = scratch + (size_t)t * 96;
        o[lane] = P.m; o[32 + lane] = P.l; o[64 + lane] = P.acc;
    }
}

// long rows: merge the partials of row r's items [first[r], first[r+1]) in order, then the epilogue
__global__ void __launch_bounds__(256) k_gnn_conv_merge(int nlong, const int32_t* __restrict__ long_rows,
                                                        const int32_t* __restrict__ first, const float* __restrict__ scratch,
                                                        const float* __restrict__ hdst, int din, const float* __restrict__ params,
                                                        float* __restrict__ hout, int relu)
{
    extern __shared__ float sp[];
    const int np = 2 * din * C + 3 * C;
    for (int k = threadIdx.x; k < np; k += blockDim.x) sp[k] = params[k];
    __syncthreads();
    const int lane = threadIdx.x & 31, c = lane & 15;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < nlong; r += warps) {
        const int i = __ldg(long_rows + r);
        float q, skip, qe;
        row_prologue(hdst, din, sp, i, c, q, skip, qe);
        Partial P{-INFINITY, 0.0f, 0.0f};
        for (int t = __ldg(first + r); t < __ldg(first + r + 1); ++t) {
            const float* o = scratch + (size_t)t * 96;
            const float m2 = o[lane], l2 = o[32 + lane], a2 = o[64 + lane];
            const float mn = fmaxf(P.m, m2);
            const float s1 = expf(P.m - mn), s2 = expf(m2 - mn);
            P.l = P.l * s1 + l2 * s2;
            P.acc = P.acc * s1 + a2 * s2;
            P.m = mn;
        }
        row_epilogue(hout, i, lane, P, skip, relu);
    }
}

// {k, v} rows of the source nodes for the next conv: kv[j][0..15] = Wk h_j + bk, kv[j][16..31] = Wv h_j + bv
__global__ void __launch_bounds__(256) k_gnn_project(int n, const float* __restrict__ h, int din, const float* __restrict__ params,
                                                     float* __restrict__ kv)
{
    extern __shared__ float sp[];
    const int np = 2 * din * C + 2 * C;
    for (int k = threadIdx.x; k < np; k += blockDim.x) sp[k] = params[k];
    __syncthreads();
    const int lane = threadIdx.x & 31, c = lane & 15;
    const float* W = lane < 16 ? sp : sp + din * C + C;
    const float* bias = W + din * C;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < n; j += warps) {
        float o = bias[c];
        for (int d = 0; d < din; ++d) o = fmaf(__ldg(h + (size_t)j * din + d), W[d * C + c], o);
        kv[(size_t)j * 32 + lane] = o;
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
}  // namespace
}  // namespace mllp

using namespace mllp;

extern "C" {

int mllp_gnn_project(int32_t n, const float* d_h, int32_t din, const float* d_params, float* d_kv, void* stream)
{
    if (n < 0 || din < 1 || din > 64 || !d_h || !d_params || !d_kv) return gfail(MLLP_E_INVALID, "mllp_gnn_project: bad argument");
    if (n == 0) return 0;
    k_gnn_project<<<grid_for_warps(n), 256, (2 * din * C + 2 * C) * sizeof(float), (cudaStream_t)stream>>>(n, d_h, din, d_params, d_kv);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return gfail((int)e, std::string("mllp_gnn_project: ") + cudaGetErrorString(e));
    return 0;
}

int mllp_gnn_conv(int32_t nd, const int32_t* d_indptr, const int32_t* d_indices, const double* d_values, const float* d_hdst,
                  int32_t din, const float* d_kv_src, const float* d_params, float* d_hout, int32_t relu, int32_t chunk,
                  int32_t nlong, const int32_t* d_long_rows, const int32_t* d_long_first, int32_t nitems, const int32_t* d_items,
                  float* d_scratch, void* stream)
{
    if (nd < 0 || din < 1 || din > 64 || !d_indptr || !d_hdst || !d_kv_src || !d_params || !d_hout || chunk < 32)
        return gfail(MLLP_E_INVALID, "mllp_gnn_conv: bad argument");
    if (nlong > 0 && (!d_long_rows || !d_long_first || !d_items || !d_scratch || nitems < nlong))
        return gfail(MLLP_E_INVALID, "mllp_gnn_conv: long-row tables missing");
    if (nd == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = (2 * din * C + 3 * C) * sizeof(float);
    k_gnn_conv_rows<<<grid_for_warps(nd), 256, smem, s>>>(nd, d_indptr, d_indices, d_values, d_hdst, din, d_kv_src, d_params, d_hout,
                                                          chunk, relu);
    if (nlong > 0) {
        k_gnn_conv_items<<<grid_for_warps(nitems), 256, smem, s>>>(nitems, d_items, d_indices, d_values, d_hdst, din, d_kv_src,
                                                                   d_params, d_scratch);
        k_gnn_conv_merge<<<grid_for_warps(nlong), 256, smem, s>>>(nlong, d_long_rows, d_long_first, d_scratch, d_hdst, din, d_params,
                                                                  d_hout, relu);
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return gfail((int)e, std::string("mllp_gnn_conv: ") + cudaGetErrorString(e));
    return 0;
}

int mllp_gnn_fc(int32_t n, const float* d_h, const float* d_wb, float* d_out, void* stream)
{
    if (n < 0 || !d_h || !d_wb || !d_out) return gfail(MLLP_E_INVALID, "mllp_gnn_fc: bad argument");
    if (n == 0) return 0;
    k_gnn_fc<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, d_h, d_wb, d_out);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return gfail((int)e, std::string("mllp_gnn_fc: ") + cudaGetErrorString(e));
    return 0;
}

}  // extern "C"
