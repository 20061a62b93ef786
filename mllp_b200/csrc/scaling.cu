// scaling.cu -- device-side scaling of LPs.
//
// (1) mllp_norm_scale: the rule behind the reference's `_norm` arrays (dataset/netlib_mps_norm/*, consumed at
//     linear_program_data.py:66-77; rule reverse-engineered from the data, SURVEY App. A.3, restated and pinned in
//     oracle/norm_rule.py): raw CSR + row senses -> standard form (one slack column per inequality row) with every row
//     scaled by 1 / ||row||_2, or by 5 / b_i when that would leave |b_i| above 5, and c / ||c||_2.
// (2) mllp_precondition: Ruiz equilibration + one Pock-Chambolle pass (the PDLP recipe) of a CSR matrix, in place on the
//     values, returning the row / column scaling vectors (used by mllp_lp_create with MLLP_F_PRECONDITION).
//
// All of it is streaming integer / fp64 work on a few MB: HBM-(L2-)bound, one pass per quantity, coalesced.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <string>

#include "../../include/mllp_b200.h"
#include "pdhg_host.h"

namespace mllp {
void set_last_error(const std::string& msg);

// ---- (1) the reference's `_norm` rule ------------------------------------------------------------------------------
// Exclusive prefix count of the rows that own a slack column (sense != 0): one CTA, chunks of 1024 rows.
__global__ void __launch_bounds__(1024, 1) k_slack_scan(int m, const signed char* __restrict__ sense, int* __restrict__ rank)
{
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < m; base += 1024) {
        const int i = base + threadIdx.x;
        const int f = (i < m && sense[i] != 0) ? 1 : 0;
        int v = f;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        if (lane == 31) warp_tot[warp] = v;
        __syncthreads();
        if (warp == 0) {
            int w = warp_tot[lane];
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_tot[lane] = w;   // inclusive over warps
        }
        __syncthreads();
        const int carry = carry_s;
        const int before = carry + (warp > 0 ? warp_tot[warp - 1] : 0) + v - f;
        if (i < m) rank[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_tot[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) rank[m] = carry_s;
}

// One warp per row.  The squares are added one by one in ascending column order (the order that reproduces the
// reference's numbers to the bit): the lanes load 32 consecutive entries (coalesced) and every lane replays the same
// left-to-right sum over the 32 shuffled squares; __dmul_rn / __dadd_rn keep the compiler from fusing them.
__global__ void __launch_bounds__(256) k_norm_rows(int m, int n, const int* __restrict__ indptr, const int* __restrict__ indices,
                                                   const double* __restrict__ values, const signed char* __restrict__ sense,
                                                   const double* __restrict__ rhs, const int* __restrict__ rank,
                                                   int* __restrict__ out_indptr, int* __restrict__ out_indices,
                                                   double* __restrict__ out_values, double* __restrict__ out_rhs,
                                                   double* __restrict__ row_scale)
{
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < m; i += warps) {
        const int a = indptr[i], e = indptr[i + 1];
        const int s = sense[i];
        double acc = 0.0;
        for (int k0 = a; k0 < e; k0 += 32) {
            const int k = k0 + lane;
            const double v = k < e ? values[k] : 0.0;
            const double sq = __dmul_rn(v, v);
            const int cnt = min(32, e - k0);
            for (int l = 0; l < cnt; ++l) acc = __dadd_rn(acc, __shfl_sync(0xffffffffu, sq, l));
        }
        if (s != 0) acc = __dadd_rn(acc, 1.0);
        const double r = sqrt(acc);
        const double b = rhs[i];
        const bool nonempty = r > 0.0;
        const bool divided = nonempty && (fabs(b) / r <= 5.0);
        const double d5 = 5.0 / b;
        const int o = a + rank[i];
        for (int k = a + lane; k < e; k += 32) {
            const double v = values[k];
            out_values[o + (k - a)] = divided ? v / r : (nonempty ? __dmul_rn(v, d5) : v);
            out_indices[o + (k - a)] = indices[k];
        }
        if (lane == 0) {
            if (s != 0) {
                const double v = s > 0 ? 1.0 : -1.0;
                out_values[o + (e - a)] = divided ? v / r : __dmul_rn(v, d5);
                out_indices[o + (e - a)] = n + rank[i];
            }
            out_rhs[i] = divided ? b / r : (nonempty ? __dmul_rn(b, d5) : b);
            row_scale[i] = divided ? 1.0 / r : (nonempty ? d5 : 1.0);
            out_indptr[i] = o;
            if (i == m - 1) out_indptr[m] = e + rank[m];
        }
    }
}

__global__ void k_norm_coefs_tail(double* out_coefs, int n, int nslack, const double* norm2, double* cnorm)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nslack; k += gridDim.x * blockDim.x) out_coefs[n + k] = 0.0;
    if (blockIdx.x == 0 && threadIdx.x == 0 && cnorm) cnorm[0] = sqrt(norm2[0]);
}

// ---- (2) Ruiz + Pock-Chambolle preconditioning ---------------------------------------------------------------------
// out[r] = max (SUM = false) or sum (SUM = true) over the entries of row r of |a| * srow[r] * scol[col]: one warp per row,
// lanes stride the row (coalesced), lane-strided partials then a butterfly -- a fixed order, no atomics.  The same kernel
// walks A (row statistics) and A' (column statistics), so nothing is accumulated across rows.
template <bool SUM>
__global__ void __launch_bounds__(256) k_scaled_row_stat(int nrows, const int* __restrict__ ptr, const int* __restrict__ ind,
                                                         const double* __restrict__ val, const double* __restrict__ srow,
                                                         const double* __restrict__ scol, double* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < nrows; r += warps) {
        const int a = ptr[r], e = ptr[r + 1];
        const double sr = srow[r];
        double acc = 0.0;
        for (int k = a + lane; k < e; k += 32) {
            const double v = (sr * fabs(val[k])) * scol[ind[k]];
            acc = SUM ? acc + v : fmax(acc, v);
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double t = __shfl_xor_sync(0xffffffffu, acc, o);
            acc = SUM ? acc + t : fmax(acc, t);
        }
        if (lane == 0) out[r] = acc;
    }
}
// s[k] /= sqrt(stat[k])  (stat = 0: an empty row / column keeps its scale)
__global__ void k_div_sqrt(double* __restrict__ s, const double* __restrict__ stat, int n)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const double t = sqrt(stat[k]);
        if (t > 0.0) s[k] = s[k] / t;
    }
}
// a_ij <- (dr_i * a_ij) * dc_j
__global__ void __launch_bounds__(256) k_scale_values(int nrows, const int* __restrict__ ptr, const int* __restrict__ ind,
                                                      double* __restrict__ val, const double* __restrict__ dr,
                                                      const double* __restrict__ dc)
{
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < nrows; r += warps) {
        const int a = ptr[r], e = ptr[r + 1];
        const double sr = dr[r];
        for (int k = a + lane; k < e; k += 32) val[k] = (sr * val[k]) * dc[ind[k]];
    }
}

static inline int row_blocks(int nrows) { return nrows <= 0 ? 1 : ((nrows + 7) / 8 > 148 * 8 ? 148 * 8 : (nrows + 7) / 8); }

// Diagonal preconditioning of A (m x n CSR; its transpose is passed as well so that column statistics are row walks):
// `ruiz_iters` rounds of Ruiz equilibration (rows and columns divided by the square root of their largest scaled
// magnitude, both from the same scaled matrix) and one Pock-Chambolle pass with alpha = 1 (square root of the scaled
// absolute row / column sums) -- the PDLP recipe, computed ON THE DEVICE.  On return h_values holds Dr A Dc and
// h_dr[m] / h_dc[n] the scaling vectors (x = Dc x~, y = Dr y~).  Synchronous (creation time, not the iteration path).
int precondition_device(int m, int n, long long nnz, const int* h_ptr, const int* h_ind, double* h_values, const int* h_tptr,
                        const int* h_tind, const double* h_tval, int ruiz_iters, double* h_dr, double* h_dc)
{
    int *ptr = nullptr, *ind = nullptr, *tptr = nullptr, *tind = nullptr;
    double *val = nullptr, *tval = nullptr, *dr = nullptr, *dc = nullptr, *sr = nullptr, *scn = nullptr;
    cudaError_t e = cudaSuccess;
    auto al = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes ? bytes : 8); };
    auto up = [&](void* d, const void* h, size_t bytes) { if (e == cudaSuccess && bytes) e = cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice); };
    const size_t z = (size_t)(nnz > 0 ? nnz : 0);
    al((void**)&ptr, sizeof(int) * ((size_t)m + 1)); al((void**)&ind, sizeof(int) * z); al((void**)&val, 8 * z);
    al((void**)&tptr, sizeof(int) * ((size_t)n + 1)); al((void**)&tind, sizeof(int) * z); al((void**)&tval, 8 * z);
    al((void**)&dr, 8 * (size_t)m); al((void**)&dc, 8 * (size_t)n); al((void**)&sr, 8 * (size_t)m); al((void**)&scn, 8 * (size_t)n);
    up(ptr, h_ptr, sizeof(int) * ((size_t)m + 1)); up(ind, h_ind, sizeof(int) * z); up(val, h_values, 8 * z);
    up(tptr, h_tptr, sizeof(int) * ((size_t)n + 1)); up(tind, h_tind, sizeof(int) * z); up(tval, h_tval, 8 * z);
    int rc = (int)e;
    if (rc == 0) rc = launch_fill(dr, 1.0, m, 0);
    if (rc == 0) rc = launch_fill(dc, 1.0, n, 0);
    for (int it = 0; it <= ruiz_iters && rc == 0; ++it) {
        const bool pc = it == ruiz_iters;   // the last round is the Pock-Chambolle pass (sums instead of maxima)
        count_launch(4);
        if (pc) {
            k_scaled_row_stat<true><<<row_blocks(m), 256>>>(m, ptr, ind, val, dr, dc, sr);
            k_scaled_row_stat<true><<<row_blocks(n), 256>>>(n, tptr, tind, tval, dc, dr, scn);
        } else {
            k_scaled_row_stat<false><<<row_blocks(m), 256>>>(m, ptr, ind, val, dr, dc, sr);
            k_scaled_row_stat<false><<<row_blocks(n), 256>>>(n, tptr, tind, tval, dc, dr, scn);
        }
        k_div_sqrt<<<row_blocks(m), 256>>>(dr, sr, m);
        k_div_sqrt<<<row_blocks(n), 256>>>(dc, scn, n);
        rc = (int)cudaGetLastError();
    }
    if (rc == 0) {
        count_launch(1);
        k_scale_values<<<row_blocks(m), 256>>>(m, ptr, ind, val, dr, dc);
        rc = (int)cudaGetLastError();
    }
    if (rc == 0) rc = (int)cudaMemcpy(h_values, val, 8 * z, cudaMemcpyDeviceToHost);
    if (rc == 0) rc = (int)cudaMemcpy(h_dr, dr, 8 * (size_t)m, cudaMemcpyDeviceToHost);
    if (rc == 0) rc = (int)cudaMemcpy(h_dc, dc, 8 * (size_t)n, cudaMemcpyDeviceToHost);
    for (void* p : {(void*)ptr, (void*)ind, (void*)val, (void*)tptr, (void*)tind, (void*)tval, (void*)dr, (void*)dc, (void*)sr, (void*)scn})
        cudaFree(p);
    if (rc != 0) set_last_error(std::string("preconditioning on the device: ") + cudaGetErrorString((cudaError_t)rc));
    return rc;
}

}  // namespace mllp

using namespace mllp;

extern "C" {

int64_t mllp_norm_scale_work_bytes(int32_t m) { return (int64_t)sizeof(int) * ((int64_t)m + 2) + 8 * (2 + 160); }

int mllp_norm_scale(int32_t m, int32_t n, int64_t nnz, int32_t nslack, const int32_t* d_indptr, const int32_t* d_indices,
                    const double* d_values, const int8_t* d_sense, const double* d_rhs, const double* d_coefs,
                    int32_t* d_out_indptr, int32_t* d_out_indices, double* d_out_values, double* d_out_rhs,
                    double* d_out_coefs, double* d_row_scale, double* d_cnorm, void* d_work, void* stream)
{
    if (m < 0 || n < 0 || nnz < 0 || nslack < 0 || nslack > m || !d_indptr || !d_sense || !d_rhs || !d_coefs || !d_out_indptr || !d_out_rhs || !d_out_coefs ||
        !d_row_scale || !d_work || (nnz > 0 && (!d_indices || !d_values || !d_out_indices || !d_out_values))) {
        set_last_error("mllp_norm_scale: null argument or bad shape");
        return MLLP_E_INVALID;
    }
    cudaStream_t s = (cudaStream_t)stream;
    // work: [m + 1] slack ranks (padded to 8 bytes) | norm2 (1 double) + the partial sums of the two-stage sum
    int* rank = (int*)d_work;
    double* norm2 = (double*)((char*)d_work + sizeof(int) * (((size_t)m + 2) & ~(size_t)1));
    count_launch(m > 0 ? 3 : 2);   // scan, rows, coefficient tail (the two-stage sum and the scaling count themselves)
    k_slack_scan<<<1, 1024, 0, s>>>(m, (const signed char*)d_sense, rank);
    if (m > 0) {
        const int blocks = (m + 7) / 8 > 148 * 8 ? 148 * 8 : (m + 7) / 8;
        k_norm_rows<<<blocks, 256, 0, s>>>(m, n, d_indptr, d_indices, d_values, (const signed char*)d_sense, d_rhs, rank,
                                           d_out_indptr, d_out_indices, d_out_values, d_out_rhs, d_row_scale);
    } else {
        cudaMemsetAsync(d_out_indptr, 0, sizeof(int), s);
    }
    int rc = launch_sumsq(d_coefs, n, norm2, norm2 + 2, s);
    if (rc == 0) rc = launch_scale_by_invnorm(d_out_coefs, d_coefs, norm2, n, s);
    if (rc == 0) {
        k_norm_coefs_tail<<<nslack > 0 ? (nslack + 255) / 256 : 1, 256, 0, s>>>(d_out_coefs, n, nslack, norm2, d_cnorm);   // slack costs are 0
        rc = (int)cudaGetLastError();
    }
    if (rc != 0) {
        set_last_error(std::string("mllp_norm_scale: ") + cudaGetErrorString((cudaError_t)rc));
        return rc;
    }
    return (int)cudaGetLastError();
}

}  // extern "C"
