// pdhg_host.h -- host-visible launchers of pdhg_kernels.cu (internal to the library).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace mllp {
struct DevMat;
struct DevLP;
struct PeerInfo;

// statistics: every kernel (or captured graph) launch the library issues is counted (mllp_launch_count)
void count_launch(int n);

int launch_gather(double* dst, const double* src, const int32_t* order, int n, cudaStream_t s);
int launch_scatter(double* dst, const double* src, const int32_t* order, int n, cudaStream_t s);
int launch_fill(double* dst, double v, int n, cudaStream_t s);
int launch_sumsq(const double* v, int n, double* out, double* scratch, cudaStream_t s);
int launch_scale_by_invnorm(double* dst, const double* src, const double* norm2, int n, cudaStream_t s);
int launch_graph_edges(int m, long long nnz, const int32_t* indptr, const int32_t* indices, const double* values,
                       long long* edge_index, float* edge_attr, cudaStream_t s);
int launch_spmv(const DevMat& M, const double* in, double* out, int G, int threads, cudaStream_t s);
int launch_primal(const DevLP& lp, bool bounds, int G, int threads, cudaStream_t s);
int launch_dual(const DevLP& lp, bool bounds, int G, int threads, cudaStream_t s);
int launch_eval_partial(const DevLP& lp, bool bounds, int G, int threads, cudaStream_t s);
int launch_eval_finalize(const DevLP& lp, int G, double* out, double iters, cudaStream_t s);
int launch_eval(const DevLP& lp, bool bounds, int G, int threads, double* out, double iters, cudaStream_t s);
int persistent_set_smem(bool bounds, size_t dyn_smem);
int persistent_max_blocks_per_sm(int threads, bool bounds, size_t dyn_smem);
int persistent_cluster_fits(int ctas, int threads, bool bounds, size_t dyn_smem);
int launch_pdhg_persistent(const DevLP& lp, bool bounds, int G, int threads, size_t dyn_smem, double tau, double sigma,
                           int iters, cudaStream_t s);
int rowpart_set_smem(bool bounds, size_t dyn_smem);
int launch_pdhg_rowpart(const DevLP& lp, const PeerInfo& pi, bool bounds, int G, int threads, size_t dyn_smem,
                        double tau, double sigma, int iters, unsigned long long seq, cudaStream_t s);
int launch_solve_persistent(const DevLP& lp, bool bounds, int G, int threads, size_t dyn_smem, double eta, double w0,
                            int max_iters, int check_every, double tol, double* out, const double* w0_dev, cudaStream_t s);

// scaling.cu: Ruiz + Pock-Chambolle preconditioning on the device (creation time); h_values becomes Dr A Dc
int precondition_device(int m, int n, long long nnz, const int* h_ptr, const int* h_ind, double* h_values, const int* h_tptr,
                        const int* h_tind, const double* h_tval, int ruiz_iters, double* h_dr, double* h_dc);
// dst[k] = src[order[k]] * s[k] (div = 0) or / s[k] (div = 1); dst[order[k]] = src[k] * s[k] or / s[k]; s = null: plain
int launch_gather_scaled(double* dst, const double* src, const int32_t* order, const double* s, int div, int n, cudaStream_t st);
int launch_scatter_scaled(double* dst, const double* src, const int32_t* order, const double* s, int div, int n, cudaStream_t st);

// multicast.cu: NVSwitch multicast mailbox of the row-partitioned exchange (driver API, bound at run time)
struct McState {
    unsigned long long mc = 0, mem = 0;      // CUmemGenericAllocationHandle of the multicast object / of this rank's mailbox
    unsigned long long uc = 0, mcva = 0;     // unicast and multicast mappings (CUdeviceptr)
    size_t size = 0, gran = 0;
    int device = 0;
    bool have_mc = false, have_mem = false, bound = false, uc_mapped = false, mc_mapped = false;
};
int mc_supported(int device, int* out);
int mc_create(McState* S, int nranks, size_t bytes, int* fd);       // rank 0
int mc_import(McState* S, int nranks, size_t bytes, int fd);        // the other ranks (fd already duplicated into this process)
int mc_add_device(McState* S, int device);                          // every rank
int mc_bind_map(McState* S);                                        // every rank, after all have added their device
void mc_destroy(McState* S);

// blocks.cu: block-angular LPs (components dealt to CTAs, linking rows through tagged words; no grid barrier)
struct BlockPlan;
int blocks_create(int m, int n, const int32_t* indptr, const int32_t* indices, const double* values, const int32_t* posX,
                  const int32_t* posY, const double* d_lb, const double* d_ub, const double* d_ylo, const double* d_yhi, int device,
                  int G, BlockPlan** out);   // *out == nullptr: no usable structure; d_lb .. d_yhi: internal-order boxes or null
int blocks_run(BlockPlan* bp, double* gx, double* gy, const double* gb, const double* gc, double tau, double sigma, int iters,
               unsigned long long tag0, cudaStream_t s);
void blocks_info(const BlockPlan* bp, int64_t* out4);   // components, linking rows, linking nonzeros, shared memory | threads << 32
int blocks_set_threads(BlockPlan* bp, int threads);
void blocks_destroy(BlockPlan* bp);
}  // namespace mllp
