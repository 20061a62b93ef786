// blocks.cu -- block-angular LPs (independent blocks coupled by a few linking rows) without grid barriers (sm_100a).
//
// Many large Netlib instances are block-angular: with its 151 linking rows (13.7 % of the nonzeros) set aside, the
// bipartite graph of ken-18 (multicommodity flow, 105 127 x 154 699) falls apart into 475 components of at most 801
// nodes.  The fused iteration of pdhg_kernels.cu pays two grid barriers (~0.95 us each) per iteration for such a
// matrix although almost no data has to cross SMs.  Here whole components are dealt to the CTAs of a cooperative
// grid: a CTA keeps its components' iterates (x, xbar, c, y, b) AND its share of the matrix (plain CSR by rows and by
// columns: a group holds ~2 000 nonzeros) in shared memory for the whole launch, one thread per row / column (four lanes
// for rows above four entries), and synchronises with __syncthreads() only.
// What does cross CTAs travels as tagged 16-byte words {value, value ^ tag} (tag = iteration number, so value and
// validity arrive in one 128-bit access and no fence is needed; ld_tagged / st_tagged in pdhg_kernels.cuh):
//   * every CTA publishes its part of every linking row's product  p[r][g] = sum_{j in g} a_rj xbar_j;
//   * the row's finisher (one warp of CTA r mod G) polls the G parts, sums them in a fixed order, updates the row's
//     dual value y_r and stores it into every CTA's mailbox;
//   * every CTA polls the nlink dual values in its mailbox into the tail of its y vector, where the column lists of
//     its A' point for their linking entries.
// Two polled L2 hops per iteration (scripts/blocks_trace.py) -- about the cost of the two grid barriers they replace -- but
// only the columns that occur in linking rows stand between them: they are updated first, the last warps then serve the
// linking rows while the other warps update the remaining columns and the block rows, all out of shared memory.  Both
// hops STORE contiguously (a CTA's nlink parts, a finisher's G copies of a dual) and poll scattered words: scattered
// 16-byte stores arrive over more than a microsecond.
// A word is overwritten only after every reader has consumed it: a CTA publishes its next parts only after it has
// read all dual values of this iteration, and a finisher publishes its next dual value only after it has read all
// parts of the next iteration.
//
// Same arithmetic as the frozen spec (oracle/pdhg_oracle.c); rows and columns are summed in CSR order (the oracle's),
// the linking rows per group and then over the groups, so iterates agree with the other kernels to rounding (1e-15),
// not bitwise.  Standard and general form (boxes on x, sign cones on y); parity mode only.
// The reference has no counterpart.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <numeric>
#include <queue>
#include <string>
#include <vector>

#include "../../include/mllp_b200.h"
#include "pdhg_host.h"
#include "pdhg_kernels.cuh"

namespace mllp {
void set_last_error(const std::string& msg);

struct BlockGroup {             // one CTA's share: the union of some components, local rows / columns sorted by length
    const int32_t* colidx;      // [wc][n] A' of the group, padded: entry slot e of local column k (original row order):
    const double* colval;       //   local row, or m + r for linking row r; padding = (0, 0.0)
    const int32_t* rowptr;      // [m + 1] the block rows
    const int32_t* rowidx;      //   local column
    const double* rowval;
    const int32_t* lptr;        // [nlink + 1] linking row r restricted to the group's columns
    const int32_t* lidx;        //   local column
    const double* lval;
    const int32_t* xpos;        // local column -> position in the handle's internal x / c vectors
    const int32_t* ypos;        // local row -> position in the handle's internal y / b vectors
    int m, n;
    int nlong;                  // leading block rows with more than 4 entries (4 lanes each)
    int wc;                     // entry slots per column (the group's longest column)
    int nlc;                    // leading columns that have an entry in a linking row
};

struct BlocksDev {
    const BlockGroup* groups;
    const int32_t* link_pos;        // [nlink] position of the linking row in the handle's internal y / b vectors
    unsigned long long* partial;    // [nlink][G] tagged words
    unsigned long long* ylink;      // [G][nlink] tagged words: every CTA's mailbox of the linking rows' dual values
    unsigned* abort_flag;
    int nlink, G;
    int max_m, max_n, max_col_nnz, max_row_nnz, max_link_nnz;   // shared memory is sized for the largest group
    const double *lb, *ub, *ylo, *yhi;   // general form: boxes in the handle's internal order (all four or none)
    unsigned long long* trace;      // dev tool (MLLP_BLOCKS_DEBUG & 2): [iter][cta][4] globaltimer stamps, else null
    int poll_gap;                   // ns between two reads of a polled word (MLLP_BLOCKS_POLL_NS)
    int dbg;                        // dev knob (MLLP_BLOCKS_DEBUG): bit 1 = timeline stamps; bit 0 (builds with -DMLLP_DEV only) = no
                                    // cross-CTA waits (timing of the local work only, results are wrong)
};

namespace {

// read a tagged word until it carries `tag`; never hangs the GPU (flags the launch after ~1 s and returns NaN).
// (One read in flight at a time: keeping four staggered reads in flight made the hop slower, 0.8 -> 1.6 us -- the extra
// polls get in the way of the stores that are waited for.)
__device__ __forceinline__ double poll_tagged(const unsigned long long* p, unsigned long long tag, unsigned* abort_flag, unsigned gap = 0)
{
    unsigned long long a, b;
    unsigned spins = 0;
    long long t0 = 0;
    for (;;) {
        ld_tagged(p, a, b);
        if ((a ^ b) == tag) return __longlong_as_double((long long)a);
        if (gap) __nanosleep(gap);
        if ((++spins & 1023u) == 0u) {
            if (t0 == 0) t0 = clock64();
            if (*(volatile unsigned*)abort_flag != 0u || clock64() - t0 > 2000000000LL) {
                *(volatile unsigned*)abort_flag = 1u;
                return __longlong_as_double(0x7ff8000000000000LL);
            }
        }
    }
}

// where CTA g's part of linking row r lives: CTA-major (a CTA's nlink parts are one contiguous, fully coalesced store;
// a finisher then reads nctas scattered words) or row-major (MLLP_PART_ROW_MAJOR: the other way round)
#ifdef MLLP_PART_ROW_MAJOR
#define PART_AT(r, g) ((size_t)(r) * nctas + (size_t)(g))
#else
#define PART_AT(r, g) ((size_t)(g) * nlink + (size_t)(r))
#endif
// where linking row r's dual for CTA g lives: row-major (the finisher's nctas copies are one contiguous store, a CTA
// reads nlink scattered words) or CTA-major (MLLP_MAIL_CTA_MAJOR: a CTA's mailbox is contiguous)
#ifdef MLLP_MAIL_CTA_MAJOR
#define MAIL_AT(r, g) ((size_t)(g) * nlink + (size_t)(r))
#else
#define MAIL_AT(r, g) ((size_t)(r) * nctas + (size_t)(g))
#endif
constexpr int MAX_FIN = 32;   // linking rows finished per CTA (one warp each)
constexpr int FIN_BATCH = 8;  // parts per lane in flight together (8 x 32 = 256 CTAs per round)

__host__ __device__ inline size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

template <bool BOUNDS>
__global__ void __launch_bounds__(1024, 1)
k_pdhg_blocks(BlocksDev B, double* gx, double* gy, const double* gb, const double* gc, double tau, double sigma, int iters,
              unsigned long long tag0)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    __shared__ double fin_y[MAX_FIN], fin_b[MAX_FIN], fin_lo[MAX_FIN], fin_hi[MAX_FIN];
    const BlockGroup G = B.groups[blockIdx.x];
    const int n = G.n, m = G.m, nlong = G.nlong, nlink = B.nlink, nctas = (int)gridDim.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int wc = G.wc;                                              // entry slots per column
    const int pw = min(max((nlink + 31) >> 5, 1), nwarps / 2);        // warps that serve the linking rows (a thread per row)
    const int rw = nwarps - pw;                                       // warps that walk the block
    const int nlc = G.nlc;                                            // leading columns with linking entries
    // shared memory: x | xbar | c | yy = (y of the block rows | duals of the linking rows) | b | colval | rowval | lval |
    //                rowptr | lptr | colidx | rowidx | lidx
    double* p = reinterpret_cast<double*>(dsm);
    double* sx = p; p += B.max_n;
    double* sxbar = p; p += B.max_n;
    double* sc = p; p += B.max_n;
    double* syy = p; p += B.max_m + nlink;
    double* sb = p; p += B.max_m;
    double *slb = nullptr, *sub = nullptr, *sylo = nullptr, *syhi = nullptr;
    if (BOUNDS) { slb = p; p += B.max_n; sub = p; p += B.max_n; sylo = p; p += B.max_m; syhi = p; p += B.max_m; }
    double* colval = p; p += B.max_col_nnz;
    double* rowval = p; p += B.max_row_nnz;
    double* lval = p; p += B.max_link_nnz;
    int* q = reinterpret_cast<int*>(p);
    int* rowptr = q; q += B.max_m + 1;
    int* lptr = q; q += nlink + 1;
    int* colidx = q; q += B.max_col_nnz;
    int* rowidx = q; q += B.max_row_nnz;
    int* lidx = q;

    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const int pos = __ldg(G.xpos + k);
        sx[k] = __ldcg(gx + pos);
        sc[k] = __ldg(gc + pos);
        sxbar[k] = 0.0;
        if (BOUNDS) { slb[k] = __ldg(B.lb + pos); sub[k] = __ldg(B.ub + pos); }
    }
    for (int k = threadIdx.x; k < m; k += blockDim.x) {
        const int pos = __ldg(G.ypos + k);
        syy[k] = __ldcg(gy + pos);
        sb[k] = __ldg(gb + pos);
        if (BOUNDS) { sylo[k] = __ldg(B.ylo + pos); syhi[k] = __ldg(B.yhi + pos); }
    }
    for (int r = threadIdx.x; r < nlink; r += blockDim.x) syy[m + r] = __ldcg(gy + __ldg(B.link_pos + r));
    for (int k = threadIdx.x; k <= m; k += blockDim.x) rowptr[k] = __ldg(G.rowptr + k);
    for (int k = threadIdx.x; k <= nlink; k += blockDim.x) lptr[k] = __ldg(G.lptr + k);
    const int nc = n * wc, nr = __ldg(G.rowptr + m), nl = __ldg(G.lptr + nlink);
    for (int e = threadIdx.x; e < nc; e += blockDim.x) { colidx[e] = __ldg(G.colidx + e); colval[e] = __ldg(G.colval + e); }
    for (int e = threadIdx.x; e < nr; e += blockDim.x) { rowidx[e] = __ldg(G.rowidx + e); rowval[e] = __ldg(G.rowval + e); }
    for (int e = threadIdx.x; e < nl; e += blockDim.x) { lidx[e] = __ldg(G.lidx + e); lval[e] = __ldg(G.lval + e); }
    // the linking rows this CTA finishes: row blockIdx.x + w * nctas is warp w's
    if (threadIdx.x < MAX_FIN) {
        const int r = (int)blockIdx.x + (int)threadIdx.x * nctas;
        if (r < nlink) {
            const int pos = __ldg(B.link_pos + r);
            fin_y[threadIdx.x] = __ldcg(gy + pos);
            fin_b[threadIdx.x] = __ldg(gb + pos);
            if (BOUNDS) { fin_lo[threadIdx.x] = __ldg(B.ylo + pos); fin_hi[threadIdx.x] = __ldg(B.yhi + pos); }
        }
    }
    __syncthreads();

    unsigned long long tag = tag0;
    for (int it = 0; it < iters; ++it) {
        ++tag;
        unsigned long long* tr = B.trace ? B.trace + ((size_t)it * gridDim.x + blockIdx.x) * 4 : nullptr;
        if (tr && threadIdx.x == 0) tr[0] = global_ns();   // iteration starts (all duals of the linking rows are here)
        // Column update: g = c - A'y (block rows and linking rows alike), x+ = max(x - tau g, 0), xbar = 2 x+ - x.  The
        // column lists are padded to the group's longest column (entry slot e of column k at [e * n + k]: conflict-free,
        // no per-column loop bounds).
        auto column = [&](int k) {
            double dot = 0.0;
            for (int e = 0; e < wc; ++e) dot = fma(colval[e * n + k], syy[colidx[e * n + k]], dot);
            const double g = sc[k] - dot, xk = sx[k];
            double xn = xk - tau * g;
            if (BOUNDS) xn = fmin(fmax(xn, slb[k]), sub[k]); else xn = fmax(xn, 0.0);
            sxbar[k] = 2.0 * xn - xk;
            sx[k] = xn;
        };
        // (1) the columns that appear in linking rows come first (they are sorted to the front): only they stand between the
        //     arrival of the linking rows' duals and the publication of this CTA's parts -- the critical path of the iteration
        for (int k = threadIdx.x; k < nlc; k += blockDim.x) column(k);
        __syncthreads();
        if (warp >= rw) {
            // (2a) the last pw warps serve the linking rows: this CTA's part of every row's product ...
            for (int r = (int)blockDim.x - 1 - (int)threadIdx.x; r < nlink; r += 32 * pw) {
                double s = 0.0;
                for (int e = lptr[r]; e < lptr[r + 1]; ++e) s = fma(lval[e], sxbar[lidx[e]], s);
                st_tagged(B.partial + 2 * PART_AT(r, blockIdx.x), s, tag);
            }
            if (tr && threadIdx.x == blockDim.x - 1) tr[1] = global_ns();   // this CTA's parts are published
            // ... the rows this CTA finishes: all parts, fixed order (lane-strided ascending, then butterfly); a lane's
            // words are all in flight together and only the stale ones are read again ...
            for (int w = nwarps - 1 - warp; w < MAX_FIN && !(B.dbg & 1); w += pw) {
                const int r = (int)blockIdx.x + w * nctas;
                if (r >= nlink) break;

                double s = 0.0;
                for (int k0 = 0; k0 < nctas; k0 += 32 * FIN_BATCH) {
                    unsigned long long a[FIN_BATCH], b[FIN_BATCH];
                    unsigned stale = 0u;
#pragma unroll
                    for (int u = 0; u < FIN_BATCH; ++u) {
                        a[u] = 0ull; b[u] = 0ull;
                        if (k0 + 32 * u + lane < nctas) stale |= 1u << u;
                    }
                    unsigned spins = 0;
                    long long t0 = 0;
                    while (stale) {
                        if (spins && B.poll_gap) __nanosleep((unsigned)B.poll_gap);
#pragma unroll
                        for (int u = 0; u < FIN_BATCH; ++u)
                            if (stale & (1u << u)) ld_tagged(B.partial + 2 * PART_AT(r, k0 + 32 * u + lane), a[u], b[u]);
#pragma unroll
                        for (int u = 0; u < FIN_BATCH; ++u)
                            if ((stale & (1u << u)) && (a[u] ^ b[u]) == tag) stale &= ~(1u << u);
                        if (stale && (++spins & 1023u) == 0u) {
                            if (t0 == 0) t0 = clock64();
                            if (*(volatile unsigned*)B.abort_flag != 0u || clock64() - t0 > 2000000000LL) {
                                *(volatile unsigned*)B.abort_flag = 1u;
                                a[0] = 0x7ff8000000000000ull;
                                stale = 0u;
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < FIN_BATCH; ++u)
                        if (k0 + 32 * u + lane < nctas) s += __longlong_as_double((long long)a[u]);
                }
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
                // the row's new dual value goes into every CTA's mailbox (one word per CTA and linking row: a CTA polls
                // only its own copy -- a single copy polled by the whole grid made its L2 lines a hot spot that also
                // delayed the store everybody was waiting for)
                double yf = 0.0;
                if (lane == 0) {
                    yf = fin_y[w] + sigma * (fin_b[w] - s);
                    if (BOUNDS) yf = fmin(fmax(yf, fin_lo[w]), fin_hi[w]);
                }
                const double yn = __shfl_sync(FULL, yf, 0);
                if (lane == 0) fin_y[w] = yn;
                for (int k = lane; k < nctas; k += 32) st_tagged(B.ylink + 2 * MAIL_AT(r, k), yn, tag);
                if (tr && w == 0 && lane == 0) tr[2] = global_ns();   // this CTA's first linking row is finished and sent
            }
            // ... and the new duals of all linking rows, from this CTA's mailbox
            for (int r = (int)blockDim.x - 1 - (int)threadIdx.x; r < nlink && !(B.dbg & 1); r += 32 * pw)
                syy[m + r] = poll_tagged(B.ylink + 2 * MAIL_AT(r, blockIdx.x), tag, B.abort_flag, (unsigned)B.poll_gap);
            if (tr && threadIdx.x == blockDim.x - 1) tr[3] = global_ns();   // all duals received (by the last warp)
        } else {
            // (2b) meanwhile the other warps update the remaining columns (none of them reads a linking row's dual) ...
            for (int k = nlc + (int)threadIdx.x; k < n; k += 32 * rw) column(k);
            asm volatile("bar.sync 1, %0;" ::"r"(32 * rw) : "memory");
            // ... and the block rows: y+ = y + sigma (b - A xbar).  Rows are sorted by length: the first nlong rows (more
            // than 4 entries) take 4 lanes each, lane q the entries q, q + 4, ...; the others one thread each.  Warps are
            // homogeneous (the lanes of long rows are padded to a warp).
            const int ulong = (4 * nlong + 31) & ~31, units = ulong + (m - nlong);
            for (int u0 = warp * 32; u0 < units; u0 += 32 * rw) {
                const int u = u0 + lane;
                if (u0 < ulong) {
                    const int k = u >> 2, qq = u & 3;
                    double dot = 0.0;
                    if (k < nlong)
                        for (int e = rowptr[k] + qq; e < rowptr[k + 1]; e += 4) dot = fma(rowval[e], sxbar[rowidx[e]], dot);
                    dot += __shfl_xor_sync(FULL, dot, 1);
                    dot += __shfl_xor_sync(FULL, dot, 2);
                    if (k < nlong && qq == 0) {
                        double yn = syy[k] + sigma * (sb[k] - dot);
                        if (BOUNDS) yn = fmin(fmax(yn, sylo[k]), syhi[k]);
                        syy[k] = yn;
                    }
                } else {
                    const int k = nlong + (u - ulong);
                    if (k < m) {
                        double dot = 0.0;
                        for (int e = rowptr[k]; e < rowptr[k + 1]; ++e) dot = fma(rowval[e], sxbar[rowidx[e]], dot);
                        double yn = syy[k] + sigma * (sb[k] - dot);
                        if (BOUNDS) yn = fmin(fmax(yn, sylo[k]), syhi[k]);
                        syy[k] = yn;
                    }
                }
            }
        }
        __syncthreads();
    }
    for (int k = threadIdx.x; k < n; k += blockDim.x) gx[__ldg(G.xpos + k)] = sx[k];
    for (int k = threadIdx.x; k < m; k += blockDim.x) gy[__ldg(G.ypos + k)] = syy[k];
    if (threadIdx.x < MAX_FIN) {
        const int r = (int)blockIdx.x + (int)threadIdx.x * nctas;
        if (r < nlink) gy[__ldg(B.link_pos + r)] = fin_y[threadIdx.x];
    }
}

struct UnionFind {
    std::vector<int32_t> parent;
    explicit UnionFind(size_t n) : parent(n) { std::iota(parent.begin(), parent.end(), 0); }
    int32_t find(int32_t a)
    {
        while (parent[a] != a) { parent[a] = parent[parent[a]]; a = parent[a]; }
        return a;
    }
    void unite(int32_t a, int32_t b)
    {
        a = find(a); b = find(b);
        if (a != b) parent[std::max(a, b)] = std::min(a, b);
    }
};

int env_i(const char* name, int dflt)
{
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

}  // namespace

struct BlockPlan {
    int device = 0;
    BlocksDev dev{};
    int threads = 1024;
    size_t smem = 0;
    const void* fn = nullptr;     // k_pdhg_blocks<false> (standard form) or <true> (boxes on x and y)
    std::vector<void*> allocs;
    int ncomp = 0, nlink = 0;
    int64_t link_nnz = 0;
};

template <class T, class F>
void up_vec(BlockPlan* bp, T** dst, const std::vector<T>& h, F&& fail_cuda)
{
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, std::max<size_t>(1, h.size()) * sizeof(T));
    if (e == cudaSuccess) {
        bp->allocs.push_back(q);
        if (!h.empty()) e = cudaMemcpy(q, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    }
    fail_cuda(e, "upload");
    *dst = reinterpret_cast<T*>(q);
}

void blocks_destroy(BlockPlan* bp)
{
    if (!bp) return;
    for (void* q : bp->allocs) cudaFree(q);
    delete bp;
}

int blocks_set_threads(BlockPlan* bp, int threads)
{
    if (!bp || threads < 64 || threads > 1024 || (threads & 31)) return -1;
    bp->threads = threads;
    return 0;
}

void blocks_info(const BlockPlan* bp, int64_t* out4)
{
    out4[0] = bp ? bp->ncomp : 0; out4[1] = bp ? bp->nlink : 0; out4[2] = bp ? bp->link_nnz : 0;
    out4[3] = bp ? (int64_t)bp->smem + ((int64_t)bp->threads << 32) : 0;   // shared memory | threads << 32
}

// Host image of the block structure: linking rows, groups of components, every group's lists (no CUDA involved).
struct HostGroups {
    int G = 0, nlink = 0, ncomp = 0;
    int64_t link_nnz = 0;
    std::vector<int32_t> link_rows;                  // original row ids, ascending
    std::vector<std::vector<int32_t>> rows, cols;    // per group: original ids in local order
    struct Img {
        int wc = 0, nlong = 0, nlc = 0;
        std::vector<int32_t> colidx, rowptr, rowidx, lptr, lidx;
        std::vector<double> colval, rowval, lval;
    };
    std::vector<Img> img;
    int max_m = 0, max_n = 0, max_c = 0, max_r = 0, max_l = 0;
};

// false: the matrix has no usable block structure
static bool build_groups(int m, int n, const int32_t* indptr, const int32_t* indices, const double* values, int G, HostGroups& H)
{
    const int64_t nnz = indptr[m];
    if (m < 1 || n < 1 || nnz < 1 || G < 2) return false;
    // 1. linking rows: far longer than the typical row
    const double mean = (double)nnz / (double)m;
    const int thr = env_i("MLLP_BLOCKS_ROW", (int)std::max(32.0, 8.0 * mean));
    std::vector<int32_t> link_id((size_t)m, -1);
    for (int i = 0; i < m; ++i)
        if (indptr[i + 1] - indptr[i] > thr) { link_id[i] = (int32_t)H.link_rows.size(); H.link_rows.push_back(i); H.link_nnz += indptr[i + 1] - indptr[i]; }
    const int nlink = H.nlink = (int)H.link_rows.size();
    if (nlink > MAX_FIN * G || 2 * H.link_nnz > nnz) return false;
    // 2. connected components of the rest (rows 0 .. m-1, columns m .. m+n-1)
    UnionFind uf((size_t)m + (size_t)n);
    for (int i = 0; i < m; ++i) {
        if (link_id[i] >= 0) continue;
        for (int32_t q = indptr[i]; q < indptr[i + 1]; ++q) uf.unite(i, m + indices[q]);
    }
    std::vector<int32_t> comp_of((size_t)m + n, -1);
    std::vector<int64_t> comp_cost;
    for (int v = 0; v < m + n; ++v) {
        if (v < m && link_id[v] >= 0) continue;
        const int32_t r = uf.find(v);
        if (comp_of[r] < 0) { comp_of[r] = (int32_t)comp_cost.size(); comp_cost.push_back(0); }
        comp_of[v] = comp_of[r];
        comp_cost[comp_of[v]] += 1 + (v < m ? indptr[v + 1] - indptr[v] : 0);
    }
    const int ncomp = H.ncomp = (int)comp_cost.size();
    if (ncomp < 2 * G) return false;   // not enough independent pieces to balance a grid
    // 3. components to groups: longest processing time first
    std::vector<int32_t> by_cost((size_t)ncomp);
    std::iota(by_cost.begin(), by_cost.end(), 0);
    std::stable_sort(by_cost.begin(), by_cost.end(), [&](int32_t a, int32_t b) { return comp_cost[a] > comp_cost[b]; });
    std::vector<int32_t> group_of((size_t)ncomp);
    {
        typedef std::pair<int64_t, int32_t> Load;   // (load, group): least loaded first, ties by group id
        std::priority_queue<Load, std::vector<Load>, std::greater<Load>> pq;
        for (int g = 0; g < G; ++g) pq.push(Load(0, g));
        for (int32_t c : by_cost) {
            Load l = pq.top(); pq.pop();
            group_of[c] = l.second;
            l.first += comp_cost[c];
            pq.push(l);
        }
    }
    // 4. local orders: rows sorted by length (threads of a warp then loop alike), columns with linking entries first
    std::vector<int32_t> col_len((size_t)n, 0);
    for (int64_t q = 0; q < nnz; ++q) ++col_len[indices[q]];
    std::vector<char> col_link((size_t)n, 0);
    for (int i : H.link_rows)
        for (int32_t q = indptr[i]; q < indptr[i + 1]; ++q) col_link[indices[q]] = 1;
    H.G = G;
    H.rows.assign((size_t)G, {}); H.cols.assign((size_t)G, {}); H.img.assign((size_t)G, {});
    auto& rows = H.rows; auto& cols = H.cols;
    for (int i = 0; i < m; ++i)
        if (link_id[i] < 0) rows[group_of[comp_of[i]]].push_back(i);
    for (int j = 0; j < n; ++j) cols[group_of[comp_of[m + j]]].push_back(j);
    std::vector<int32_t> local_row((size_t)m, -1), local_col((size_t)n, -1), col_group((size_t)n, 0);
    for (int g = 0; g < G; ++g) {
        std::stable_sort(rows[g].begin(), rows[g].end(), [&](int32_t a, int32_t b) { return indptr[a + 1] - indptr[a] > indptr[b + 1] - indptr[b]; });
        std::stable_sort(cols[g].begin(), cols[g].end(), [&](int32_t a, int32_t b) {
            return col_link[a] != col_link[b] ? col_link[a] > col_link[b] : col_len[a] > col_len[b];   // linking columns first
        });
        for (size_t k = 0; k < rows[g].size(); ++k) local_row[rows[g][k]] = (int32_t)k;
        for (size_t k = 0; k < cols[g].size(); ++k) { local_col[cols[g][k]] = (int32_t)k; col_group[cols[g][k]] = g; }
        H.max_m = std::max<int>(H.max_m, (int)rows[g].size());
        H.max_n = std::max<int>(H.max_n, (int)cols[g].size());
    }
    // 5. the groups' lists.  Transpose once: entries of every column in original row order
    std::vector<int64_t> tptr((size_t)n + 1, 0);
    for (int64_t q = 0; q < nnz; ++q) ++tptr[indices[q] + 1];
    for (int j = 0; j < n; ++j) tptr[j + 1] += tptr[j];
    std::vector<int32_t> trow((size_t)nnz);
    std::vector<double> tval((size_t)nnz);
    {
        std::vector<int64_t> cur(tptr.begin(), tptr.end() - 1);
        for (int i = 0; i < m; ++i)
            for (int32_t q = indptr[i]; q < indptr[i + 1]; ++q) { const int64_t at = cur[indices[q]]++; trow[at] = i; tval[at] = values[q]; }
    }
    for (int g = 0; g < G; ++g) {
        HostGroups::Img& I = H.img[g];
        const int mg = (int)rows[g].size(), ng = (int)cols[g].size();
        for (int j : cols[g]) { I.wc = std::max<int>(I.wc, (int)(tptr[j + 1] - tptr[j])); I.nlc += col_link[j] ? 1 : 0; }
        I.colidx.assign((size_t)I.wc * ng, 0);
        I.colval.assign((size_t)I.wc * ng, 0.0);
        for (int k = 0; k < ng; ++k) {
            const int j = cols[g][k];
            int e = 0;
            for (int64_t q = tptr[j]; q < tptr[j + 1]; ++q, ++e) {
                const int i = trow[q];
                I.colidx[(size_t)e * ng + k] = link_id[i] >= 0 ? mg + link_id[i] : local_row[i];
                I.colval[(size_t)e * ng + k] = tval[q];
            }
        }
        H.max_c = std::max(H.max_c, I.wc * ng);
        for (int i : rows[g]) {
            I.rowptr.push_back((int32_t)I.rowidx.size());
            I.nlong += (indptr[i + 1] - indptr[i] > 4) ? 1 : 0;   // rows are sorted by length
            for (int32_t q = indptr[i]; q < indptr[i + 1]; ++q) { I.rowidx.push_back(local_col[indices[q]]); I.rowval.push_back(values[q]); }
        }
        I.rowptr.push_back((int32_t)I.rowidx.size());
        H.max_r = std::max(H.max_r, (int)I.rowidx.size());
        I.lptr.reserve((size_t)nlink + 1);
    }
    // linking rows restricted to every group (column order of the original row)
    for (int r = 0; r < nlink; ++r) {
        for (int g = 0; g < G; ++g) H.img[g].lptr.push_back((int32_t)H.img[g].lidx.size());
        const int i = H.link_rows[r];
        for (int32_t q = indptr[i]; q < indptr[i + 1]; ++q) {
            const int j = indices[q], g = col_group[j];
            H.img[g].lidx.push_back(local_col[j]); H.img[g].lval.push_back(values[q]);
        }
    }
    for (int g = 0; g < G; ++g) {
        H.img[g].lptr.push_back((int32_t)H.img[g].lidx.size());
        H.max_l = std::max(H.max_l, (int)H.img[g].lidx.size());
    }
    return true;
}

static size_t blocks_smem_bytes(const HostGroups& H, bool bounds = false)
{
    return align16(8 * ((bounds ? 5 : 3) * (size_t)H.max_n + (bounds ? 4 : 2) * (size_t)H.max_m + (size_t)H.nlink + (size_t)H.max_c + (size_t)H.max_r + (size_t)H.max_l) +
                   4 * ((size_t)H.max_m + (size_t)H.nlink + 2 + (size_t)H.max_c + (size_t)H.max_r + (size_t)H.max_l));
}

// Returns 0 with *out == nullptr when the matrix has no usable block structure (not an error).
int blocks_create(int m, int n, const int32_t* indptr, const int32_t* indices, const double* values, const int32_t* posX,
                  const int32_t* posY, const double* d_lb, const double* d_ub, const double* d_ylo, const double* d_yhi, int device,
                  int G, BlockPlan** out)
{
    *out = nullptr;
    const bool bounds = d_lb != nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return 0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return 0;
    const size_t smem_cap = (size_t)prop.sharedMemPerBlockOptin - 2048;
    BlockPlan* bp = nullptr;
    int rc = 0;
    auto fail_cuda = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && rc == 0) { set_last_error(std::string("blocks_create: ") + what + ": " + cudaGetErrorString(e)); rc = (int)e; }
    };
    try {
        HostGroups H;
        if (!build_groups(m, n, indptr, indices, values, G, H)) return 0;
        const size_t smem = blocks_smem_bytes(H, bounds);
        if (smem > smem_cap) return 0;   // a group's iterates and matrix share do not fit in shared memory
        bp = new (std::nothrow) BlockPlan();
        if (!bp) { set_last_error("blocks_create: out of host memory"); return MLLP_E_NOMEM; }
        bp->device = device; bp->ncomp = H.ncomp; bp->nlink = H.nlink; bp->link_nnz = H.link_nnz;
        auto up = [&](auto** dst, const auto& h) { up_vec(bp, dst, h, fail_cuda); };
        // all groups' lists in a few pools
        struct GroupOff { size_t ce, rp, re, lp, le, xp, yp; };
        std::vector<GroupOff> GO((size_t)G);
        std::vector<int32_t> P_ci, P_rp, P_ri, P_lp, P_li, P_pos;
        std::vector<double> P_cv, P_rv, P_lv;
        for (int g = 0; g < G; ++g) {
            const HostGroups::Img& I = H.img[g];
            GroupOff& o = GO[g];
            o.ce = P_ci.size(); P_ci.insert(P_ci.end(), I.colidx.begin(), I.colidx.end()); P_cv.insert(P_cv.end(), I.colval.begin(), I.colval.end());
            o.rp = P_rp.size(); P_rp.insert(P_rp.end(), I.rowptr.begin(), I.rowptr.end());
            o.re = P_ri.size(); P_ri.insert(P_ri.end(), I.rowidx.begin(), I.rowidx.end()); P_rv.insert(P_rv.end(), I.rowval.begin(), I.rowval.end());
            o.lp = P_lp.size(); P_lp.insert(P_lp.end(), I.lptr.begin(), I.lptr.end());
            o.le = P_li.size(); P_li.insert(P_li.end(), I.lidx.begin(), I.lidx.end()); P_lv.insert(P_lv.end(), I.lval.begin(), I.lval.end());
            o.xp = P_pos.size(); for (int j : H.cols[g]) P_pos.push_back(posX[j]);
            o.yp = P_pos.size(); for (int i : H.rows[g]) P_pos.push_back(posY[i]);
        }
        int32_t *d_ci = nullptr, *d_rp = nullptr, *d_ri = nullptr, *d_lp = nullptr, *d_li = nullptr, *d_pos = nullptr, *d_link_pos = nullptr;
        double *d_cv = nullptr, *d_rv = nullptr, *d_lv = nullptr;
        unsigned long long *d_partial = nullptr, *d_ylink = nullptr;
        unsigned* d_abort = nullptr;
        BlockGroup* d_groups = nullptr;
        up(&d_ci, P_ci); up(&d_cv, P_cv); up(&d_rp, P_rp); up(&d_ri, P_ri); up(&d_rv, P_rv);
        up(&d_lp, P_lp); up(&d_li, P_li); up(&d_lv, P_lv); up(&d_pos, P_pos);
        std::vector<int32_t> link_pos((size_t)H.nlink);
        for (int r = 0; r < H.nlink; ++r) link_pos[r] = posY[H.link_rows[r]];
        up(&d_link_pos, link_pos);
        up(&d_partial, std::vector<unsigned long long>(2 * (size_t)std::max(H.nlink, 1) * G, 0ull));
        up(&d_ylink, std::vector<unsigned long long>(2 * (size_t)std::max(H.nlink, 1) * G, 0ull));
        up(&d_abort, std::vector<unsigned>(4, 0u));
        if (rc == 0) {
            std::vector<BlockGroup> groups((size_t)G);
            for (int g = 0; g < G; ++g) {
                BlockGroup& Q = groups[g];
                const GroupOff& o = GO[g];
                Q.colidx = d_ci + o.ce; Q.colval = d_cv + o.ce; Q.wc = H.img[g].wc;
                Q.rowptr = d_rp + o.rp; Q.rowidx = d_ri + o.re; Q.rowval = d_rv + o.re;
                Q.lptr = d_lp + o.lp; Q.lidx = d_li + o.le; Q.lval = d_lv + o.le;
                Q.xpos = d_pos + o.xp; Q.ypos = d_pos + o.yp;
                Q.m = (int)H.rows[g].size(); Q.n = (int)H.cols[g].size();
                Q.nlong = H.img[g].nlong; Q.nlc = H.img[g].nlc;
            }
            up(&d_groups, groups);
        }
        if (rc == 0) {
            BlocksDev& D = bp->dev;
            D.groups = d_groups; D.link_pos = d_link_pos; D.partial = d_partial; D.ylink = d_ylink; D.abort_flag = d_abort;
            D.nlink = H.nlink; D.G = G; D.max_m = H.max_m; D.max_n = H.max_n;
            D.max_col_nnz = H.max_c; D.max_row_nnz = H.max_r; D.max_link_nnz = H.max_l;
            D.lb = d_lb; D.ub = d_ub; D.ylo = d_ylo; D.yhi = d_yhi;
            D.trace = nullptr;
#ifdef MLLP_DEV
            D.dbg = env_i("MLLP_BLOCKS_DEBUG", 0);   // dev builds only (-DMLLP_DEV): bit 0 skips the cross-CTA waits (WRONG results)
#else
            D.dbg = env_i("MLLP_BLOCKS_DEBUG", 0) & 2;   // release builds: only the timeline stamps; nothing can skip work
#endif
            D.poll_gap = env_i("MLLP_BLOCKS_POLL_NS", 0);
            bp->smem = smem;
            bp->threads = env_i("MLLP_BLOCKS_THREADS", 1024);
            if (bp->threads < 64 || bp->threads > 1024 || (bp->threads & 31)) bp->threads = 1024;
            bp->fn = bounds ? (const void*)k_pdhg_blocks<true> : (const void*)k_pdhg_blocks<false>;
            fail_cuda(cudaFuncSetAttribute(bp->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bp->smem), "cudaFuncSetAttribute");
            int nb = 0;
            fail_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, bp->fn, bp->threads, bp->smem), "occupancy");
            if (rc == 0 && nb * prop.multiProcessorCount < G) rc = -1;   // the whole grid must be resident
        }
    } catch (const std::bad_alloc&) {
        set_last_error("blocks_create: out of host memory");
        rc = MLLP_E_NOMEM;
    }
    if (rc != 0) {
        blocks_destroy(bp);
        return rc == -1 ? 0 : rc;   // -1: does not fit -- not an error, the caller keeps the grid kernel
    }
    *out = bp;
    return 0;
}

// Host-only check of the block images (no GPU): every row and column is placed exactly once, and the products A xbar
// and A'y replayed from the groups' lists -- block rows per group, linking rows as the sum of the groups' parts, columns
// from the padded column lists over (block duals | linking duals) -- equal the plain CSR products.
// out8: [0] 1 if the matrix has a usable block structure for G groups, [1] components, [2] linking rows, [3] their
// nonzeros, [4] worst relative error of A xbar, [5] of A'y, [6] shared memory per CTA (bytes), [7] largest group's columns.
extern "C" int mllp_blocks_selfcheck(int32_t m, int32_t n, int64_t nnz, const int32_t* indptr, const int32_t* indices,
                                     const double* values, int32_t G, double* out8)
{
    if (!out8 || !indptr || m < 0 || n < 0 || nnz < 0 || (nnz > 0 && (!indices || !values)) || G < 2) {
        set_last_error("mllp_blocks_selfcheck: bad argument");
        return MLLP_E_INVALID;
    }
    for (int k = 0; k < 8; ++k) out8[k] = 0.0;
    try {
        HostGroups H;
        if (!build_groups(m, n, indptr, indices, values, G, H)) return 0;
        out8[0] = 1.0; out8[1] = H.ncomp; out8[2] = H.nlink; out8[3] = (double)H.link_nnz;
        out8[6] = (double)blocks_smem_bytes(H); out8[7] = H.max_n;
        std::vector<int> seen_row((size_t)m, 0), seen_col((size_t)n, 0);
        for (int i : H.link_rows) ++seen_row[i];
        for (int g = 0; g < G; ++g) {
            for (int i : H.rows[g]) ++seen_row[i];
            for (int j : H.cols[g]) ++seen_col[j];
        }
        for (int i = 0; i < m; ++i) if (seen_row[i] != 1) { set_last_error("mllp_blocks_selfcheck: a row is placed " + std::to_string(seen_row[i]) + " times"); return MLLP_E_STATE; }
        for (int j = 0; j < n; ++j) if (seen_col[j] != 1) { set_last_error("mllp_blocks_selfcheck: a column is placed " + std::to_string(seen_col[j]) + " times"); return MLLP_E_STATE; }
        // deterministic pseudo-random vectors
        std::vector<double> xb((size_t)n), y((size_t)m), ax((size_t)m, 0.0), aty((size_t)n, 0.0), ax2((size_t)m, 0.0), aty2((size_t)n, 0.0);
        unsigned long long st = 88172645463325252ull;
        auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double)(st >> 11) / 9007199254740992.0 - 0.5; };
        for (double& v : xb) v = rnd();
        for (double& v : y) v = rnd();
        for (int i = 0; i < m; ++i)
            for (int32_t q = indptr[i]; q < indptr[i + 1]; ++q) { ax[i] += values[q] * xb[indices[q]]; aty[indices[q]] += values[q] * y[i]; }
        for (int g = 0; g < G; ++g) {
            const HostGroups::Img& I = H.img[g];
            const int mg = (int)H.rows[g].size(), ng = (int)H.cols[g].size();
            std::vector<double> yy((size_t)mg + H.nlink);
            for (int k = 0; k < mg; ++k) yy[k] = y[H.rows[g][k]];
            for (int r = 0; r < H.nlink; ++r) yy[(size_t)mg + r] = y[H.link_rows[r]];
            for (int k = 0; k < ng; ++k) {
                double d = 0.0;
                for (int e = 0; e < I.wc; ++e) d += I.colval[(size_t)e * ng + k] * yy[I.colidx[(size_t)e * ng + k]];
                aty2[H.cols[g][k]] = d;
            }
            for (int k = 0; k < mg; ++k) {
                double d = 0.0;
                for (int32_t e = I.rowptr[k]; e < I.rowptr[k + 1]; ++e) d += I.rowval[e] * xb[H.cols[g][I.rowidx[e]]];
                ax2[H.rows[g][k]] = d;
            }
            for (int r = 0; r < H.nlink; ++r)
                for (int32_t e = I.lptr[r]; e < I.lptr[r + 1]; ++e) ax2[H.link_rows[r]] += I.lval[e] * xb[H.cols[g][I.lidx[e]]];
            if (I.nlc > ng || I.nlong > mg) { set_last_error("mllp_blocks_selfcheck: bad counts"); return MLLP_E_STATE; }
        }
        double ea = 0.0, et = 0.0;
        for (int i = 0; i < m; ++i) ea = std::max(ea, std::fabs(ax[i] - ax2[i]) / (1.0 + std::fabs(ax[i])));
        for (int j = 0; j < n; ++j) et = std::max(et, std::fabs(aty[j] - aty2[j]) / (1.0 + std::fabs(aty[j])));
        out8[4] = ea; out8[5] = et;
    } catch (const std::bad_alloc&) {
        set_last_error("mllp_blocks_selfcheck: out of host memory");
        return MLLP_E_NOMEM;
    }
    return 0;
}

// dev tool: stamps of the last traced launch (MLLP_BLOCKS_DEBUG & 2), [iters][G][4]
static std::vector<unsigned long long> g_blocks_trace;
static int g_blocks_trace_iters = 0, g_blocks_trace_G = 0;
extern "C" int mllp_debug_blocks_trace(unsigned long long* out, int64_t cap, int32_t* iters, int32_t* G)
{
    *iters = g_blocks_trace_iters; *G = g_blocks_trace_G;
    const size_t nwords = std::min<size_t>((size_t)cap, g_blocks_trace.size());
    if (out && nwords) memcpy(out, g_blocks_trace.data(), nwords * sizeof(unsigned long long));
    return 0;
}

int blocks_run(BlockPlan* bp, double* gx, double* gy, const double* gb, const double* gc, double tau, double sigma, int iters,
               unsigned long long tag0, cudaStream_t s)
{
    if ((bp->dev.dbg & 2) && iters > 0 && iters <= 256) {   // traced (synchronous) launch
        unsigned long long* d_tr = nullptr;
        const size_t cnt = (size_t)iters * bp->dev.G * 4;
        if (cudaMalloc(&d_tr, cnt * sizeof(unsigned long long)) != cudaSuccess) return (int)cudaGetLastError();
        cudaMemset(d_tr, 0, cnt * sizeof(unsigned long long));
        BlocksDev D = bp->dev;
        D.trace = d_tr;
        cudaMemsetAsync(D.abort_flag, 0, sizeof(unsigned), s);
        void* args[] = {&D, &gx, &gy, &gb, &gc, &tau, &sigma, &iters, &tag0};
        count_launch(1);
        cudaError_t e = cudaLaunchCooperativeKernel(bp->fn, dim3(D.G), dim3(bp->threads), args, bp->smem, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        g_blocks_trace.assign(cnt, 0ull);
        if (e == cudaSuccess) e = cudaMemcpy(g_blocks_trace.data(), d_tr, cnt * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        g_blocks_trace_iters = iters; g_blocks_trace_G = D.G;
        cudaFree(d_tr);
        return (int)e;
    }
    cudaError_t e = cudaMemsetAsync(bp->dev.abort_flag, 0, sizeof(unsigned), s);
    if (e != cudaSuccess) return (int)e;
    BlocksDev D = bp->dev;
    void* args[] = {&D, &gx, &gy, &gb, &gc, &tau, &sigma, &iters, &tag0};
    count_launch(1);
    e = cudaLaunchCooperativeKernel(bp->fn, dim3(D.G), dim3(bp->threads), args, bp->smem, s);
    return (int)e;
}

}  // namespace mllp
