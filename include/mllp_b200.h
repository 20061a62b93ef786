/*
 * mllp_b200.h -- C ABI of the B200-native primal-dual LP iteration for HAHHHD/mllp.
 *
 * The reference is pure Python and has no FFI; the boundary it offers is "a free function
 * in linear_program_methods.py called per instance with the loader's tuple" (SURVEY.md
 * section 8b).  Each entry point below names the reference interface it sits behind:
 *
 *   mllp_lp_create        <- the CSR arrays produced by linear_program_data.py:75-77
 *                            (scipy.sparse.load_npz -> .indptr/.indices/.data; float64 + int32)
 *                            in the argument order of
 *                            build_graph_from_weights_sets(constrs, constr_weights, rhs, coefs,
 *                            device), linear_program_methods.py:89
 *   mllp_pdhg_run / _host <- the "solve an LP, return (objective, solution)" convention of
 *                            ortools_max_covering / gurobi_max_covering,
 *                            linear_program_methods.py:477 / :542 (the reference's only LP solve
 *                            call sites; the iteration itself is new)
 *   mllp_pdhg_solve       <- same, run to a tolerance instead of a fixed iteration count
 *   mllp_spmv             <- the A-traversal of build_graph_from_weights_sets,
 *                            linear_program_methods.py:93-96 (exposed for unit parity)
 *   mllp_batch_*          <- the per-instance loop of linear_program_experiment.py:123
 *                            (for name, constrs, constr_weights, coefs, rhs, basis_opt in ...)
 *                            packed into one launch
 *
 * Conventions: plain pointers and sizes, no torch types.  Every function returns 0 on
 * success or a non-zero status (CUDA error code, or MLLP_E_*); mllp_last_error() returns the
 * message of the last failure on the calling thread.  "d_" pointers are device pointers owned
 * by the caller (e.g. torch tensors' data_ptr()), "h_" pointers are host pointers.  `stream` is
 * a cudaStream_t passed as void* (NULL = legacy default stream).  Calls are asynchronous on
 * that stream unless noted.  A handle is bound to one device and is not re-entrant.
 * All vectors are in the caller's (reference) row/column order; internal reordering is hidden.
 */
#ifndef MLLP_B200_H
#define MLLP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MLLP_E_INVALID 1001  /* bad argument */
#define MLLP_E_NOMEM 1002    /* host allocation failed */
#define MLLP_E_STATE 1003    /* handle in the wrong state / unsupported configuration */

/* number of doubles written by the *_run / *_solve calls into their `scalars` output */
#define MLLP_NUM_SCALARS 16
/* scalars[0] pobj=c'x  [1] dobj  [2] ||primal res||_2  [3] ||dual res||_2  [4] ||b||_2
 * [5] ||c||_2  [6] ||x||_2  [7] ||y||_2  [8] relative KKT error  [9] |pobj-dobj|
 * [10] iterations done  [11] restarts  [12] converged flag  [13] final primal weight
 * [14] last fixed-point error  [15] reserved
 * (same definitions as oracle_kkt in oracle/pdhg_oracle.c) */

typedef struct mllp_lp *mllp_lp_t;
typedef struct mllp_batch *mllp_batch_t;

/* creation flags */
#define MLLP_F_DEFAULT 0u
#define MLLP_F_NO_SMEM_RESIDENT 1u /* stream the matrix from L2/HBM even if it fits on-chip */
#define MLLP_F_GRAPH_MODE 2u       /* one kernel launch per half-iteration (CUDA graph) instead
                                      of the persistent cooperative kernel */
#define MLLP_F_NO_TUNE 4u          /* skip the tuning rounds of mllp_lp_create (measured re-dealing of the
                                      tiles to the CTAs; results never depend on it, only speed) */
#define MLLP_F_PRECONDITION 8u     /* diagonal preconditioning computed ON THE DEVICE at create time (10 rounds of Ruiz
                                      equilibration + one Pock-Chambolle pass, the PDLP recipe): the handle holds Dr A Dc and
                                      iterates on the scaled LP.  Nothing changes for the caller: b, c, x, y, the boxes and
                                      mllp_spmv are those of the ORIGINAL LP (scaled / un-scaled at the boundary) and the KKT
                                      scalars -- also the termination test of mllp_pdhg_solve -- are evaluated on the
                                      ORIGINAL LP.  mllp_estimate_norm returns ||Dr A Dc||_2 (what the step size needs).
                                      Also accepted by mllp_batch_create (every distinct matrix is preconditioned). */

const char *mllp_last_error(void);
int mllp_version(void);
/* Statistics: number of kernel (or captured-graph) launches this library has issued in the calling process so far --
 * the difference around a region is the count of the library's own launches inside it (bench.py: gpu_launches). */
long long mllp_launch_count(void);

/* Host-only self check of the format builder (no GPU needed): builds the tiled images of A
 * and A' for a grid of `num_ctas` CTAs and replays the kernel's tile walk on the CPU against
 * the plain CSR dot products.  out8[0] = worst relative row error, [1]/[2] tiles of A/A',
 * [3]/[4] padding factors, [5] split rows of A, [6]/[7] largest per-CTA step counts.
 * Returns 0 if every row is produced exactly once and the format is well formed. */
int mllp_format_selfcheck(int32_t m, int32_t n, int64_t nnz, const int32_t *indptr,
                          const int32_t *indices, const double *values, int32_t num_ctas,
                          int32_t pref_steps, int32_t max_steps, double *out8);

/* Host-only self check of the row-partitioned build: emulates all `nranks` ranks on the CPU
 * (tile walk of every rank's rows of A + exchange of the slices; the whole of A' against the padded y)
 * against the plain CSR products.  out4: [0] worst relative row error, [1] padded internal length of y,
 * [2] length of x, [3] largest nonzeros of A per rank over the mean. */
int mllp_rowpart_selfcheck(int32_t m, int32_t n, int64_t nnz, const int32_t *indptr,
                           const int32_t *indices, const double *values, int32_t num_ctas,
                           int32_t nranks, double *out4);

/* Host-only self check of the row-per-lane images the warp-per-instance batch kernels walk in solve mode (groups of 32
 * internal rows, lane = row, slot-major entries): builds them for A and A' in the batch builder's internal orders and replays
 * the lane walk on the CPU against the plain CSR products.  out4: [0] worst relative row error, [1] / [2] slots (of 32
 * entries) of A / A', [3] padding factor (entries stored / nonzeros). */
int mllp_ell_selfcheck(int32_t m, int32_t n, int64_t nnz, const int32_t *indptr, const int32_t *indices,
                       const double *values, double *out4);

/* Host-only statistic of the built format: distinct 128-byte lines touched by the warp-wide
 * gather instructions (the gather cost model).  out6: [0]/[1] total lines A / A', [2]/[3] the
 * largest per-CTA sum, [4]/[5] gather instructions.  `cluster` = cluster rows inside length
 * classes (the default of mllp_lp_create). */
int mllp_format_gather_lines(int32_t m, int32_t n, int64_t nnz, const int32_t *indptr,
                             const int32_t *indices, const double *values, int32_t num_ctas,
                             int32_t pref_steps, int32_t max_steps, int32_t cluster, double *out6);

/* Device facts the host side needs: out[0]=SM count, out[1]=L2 bytes, out[2]=max smem/CTA. */
int mllp_device_info(int device, int64_t *out3);

/*
 * Build the device-resident formats of A (m x n, CSR, host pointers) and A' on `device`.
 * h_lb / h_ub (n) may be NULL => l = 0, u = +inf (the dataset's standard form).
 * h_ylo / h_yhi (m) may be NULL => all rows are equalities (dual free).
 * Synchronous.  The input arrays are not referenced after return.
 */
int mllp_lp_create(int32_t m, int32_t n, int64_t nnz, const int32_t *h_indptr,
                   const int32_t *h_indices, const double *h_values, const double *h_lb,
                   const double *h_ub, const double *h_ylo, const double *h_yhi, int device,
                   uint32_t flags, mllp_lp_t *out);
int mllp_lp_destroy(mllp_lp_t lp);

/* Geometry of the built formats: out[0]=m, [1]=n, [2]=nnz, [3]=tiles(A), [4]=tiles(A'),
 * [5]=padded entries(A), [6]=padded entries(A'), [7]=split rows(A), [8]=split rows(A'),
 * [9]=grid CTAs, [10]=threads per CTA, [11]=dynamic shared memory per CTA (bytes),
 * [12]=algorithmic bytes per iteration (24 nnz + 36 m + 44 n + 8, +16 n with bounds,
 * +16 m with row senses), [13]/[14]=per-CTA cap of shared-memory resident warp-steps of A/A',
 * [15]=CTAs per SM. */
int mllp_lp_info(mllp_lp_t lp, int64_t *out16);

/* Tuning rounds of mllp_lp_create (single GPU, persistent kernel): a few traced iterations give every
 * CTA's time per phase, which is fed back into the dealing of the tiles; the fastest build is kept.
 * out4: [0] ns / iteration of the first build, [1] of the build kept, [2] rounds run, [3] reserved. */
int mllp_lp_tune_info(mllp_lp_t lp, double *out4);

/* Launch geometry of the persistent kernels chosen by mllp_lp_create.  A grid barrier costs ~0.9 us, twice per
 * iteration, so an LP whose phases are shorter than that runs on ONE thread-block cluster (<= 16 SMs, hardware
 * cluster barrier) or on one CTA instead of the cooperative grid of one CTA per SM.  In the "broadcast" cluster
 * geometry every CTA also keeps a full copy of the two gathered vectors in shared memory (a row update stores its
 * entry into all copies over distributed shared memory), so nothing inside the iteration touches global memory.
 * With tuning enabled the candidates are timed and the fastest is kept; environment MLLP_GEOM forces one
 * (0 grid, 1 one CTA, 2..16 cluster, 101..116 broadcast cluster of 1..16 CTAs; 101 = one CTA with all vectors in
 * shared memory).
 * out12: [0] 0 = cooperative grid, 1 = cluster, 2 = one CTA, 3 = broadcast cluster; [1] CTAs; [2..10] measured
 * ns / iteration of the grid, of a cluster of 16 / 8 / 4 CTAs, of one CTA and of a broadcast cluster of 16 / 8 / 4 / 1
 * CTAs (0 = not tried); [11] reserved. */
int mllp_lp_geometry(mllp_lp_t lp, double *out12);

/* Block-angular LPs (independent blocks coupled by a few long "linking" rows, e.g. the multicommodity-flow instances
 * ken-*): mllp_lp_create looks for the structure (rows more than 8x longer than the mean set aside, connected
 * components of the rest) and, when there are enough blocks to fill the grid, builds a second image of the LP in which
 * whole blocks are dealt to CTAs.  The parity kernel then keeps every CTA's blocks and iterates in shared memory, meets
 * with __syncthreads() only, and exchanges the linking rows' partial products and dual values as tagged 16-byte words
 * (two L2 hops per iteration instead of two grid barriers); standard and general form.  Both kernels are timed at create time, the faster is
 * used by mllp_pdhg_run (MLLP_BLOCKS=0 / 1 disables / forces it); iterates agree with the grid kernel to rounding.
 * out8: [0] 1 if mllp_pdhg_run uses the block kernel, [1] blocks (components), [2] linking rows, [3] their nonzeros,
 * [4] shared memory per CTA, [5] / [6] measured ns per iteration of the grid kernel / of the block kernel,
 * [7] 1 if the structure was found and the image built. */
int mllp_lp_blocks_info(mllp_lp_t lp, double *out8);

/* Host-only check of the block images (no GPU; used by the CPU tests): every row and column is placed exactly once,
 * and the products A xbar and A'y replayed from the groups' lists equal the plain CSR products.
 * out8: [0] 1 if the matrix has a usable block structure for G groups, [1] components, [2] linking rows, [3] their
 * nonzeros, [4] worst relative error of A xbar, [5] of A'y, [6] shared memory per CTA (bytes), [7] largest group's
 * columns. */
int mllp_blocks_selfcheck(int32_t m, int32_t n, int64_t nnz, const int32_t *indptr, const int32_t *indices,
                          const double *values, int32_t G, double *out8);

/* The diagonal preconditioner of a handle created with MLLP_F_PRECONDITION, in the caller's row / column order:
 * d_dr[m], d_dc[n] with the handle holding Dr A Dc (all ones on a handle without preconditioning). */
int mllp_lp_scaling(mllp_lp_t lp, double *d_dr, double *d_dc, void *stream);

/* d_out = A d_in (trans = 0; d_in has n, d_out m entries) or A' d_in (trans = 1). */
int mllp_spmv(mllp_lp_t lp, int trans, const double *d_in, double *d_out, void *stream);

/* sigma_max(A) by `iters` power-iteration steps on A'A from 1/sqrt(n); synchronous. */
int mllp_estimate_norm(mllp_lp_t lp, int iters, double *h_sigma_max, void *stream);

/*
 * Parity mode: `num_iters` fixed-step iterations
 *     g = c - A'y;  x+ = clip(x - tau g, l, u);  xbar = 2x+ - x;  y+ = clip(y + sigma(b - A xbar))
 * in place on d_x (n) and d_y (m).  If d_scalars != NULL the KKT scalars of the final
 * (x, y) are written there (MLLP_NUM_SCALARS doubles, device memory).  No host sync.
 */
int mllp_pdhg_run(mllp_lp_t lp, double *d_x, double *d_y, const double *d_b, const double *d_c,
                  double tau, double sigma, int32_t num_iters, double *d_scalars, void *stream);

/* Same, with HOST buffers: copies b, c, x, y in, runs, copies x, y and the scalars out,
 * and synchronises the stream before returning (the end-to-end entry the Python
 * pdhg_linear_program() uses for numpy inputs). */
int mllp_pdhg_run_host(mllp_lp_t lp, double *h_x, double *h_y, const double *h_b,
                       const double *h_c, double tau, double sigma, int32_t num_iters,
                       double *h_scalars, void *stream);

/*
 * Solve mode: reflected restarted Halpern PDHG with fixed step eta (tau = eta/w,
 * sigma = eta*w, w0 > 0 the initial primal weight; w0 = 0 selects the PDLP default ||c||_2 / ||b||_2 of the LP the
 * handle iterates on -- the scaled one when preconditioned --, computed on the device), KKT check and restart test every
 * `check_every` iterations, all decided on the device; stops at rel. KKT error <= tol or
 * max_iters.  Spec: oracle_pdhg_solve in oracle/pdhg_oracle.c.  No host sync.
 */
int mllp_pdhg_solve(mllp_lp_t lp, double *d_x, double *d_y, const double *d_b,
                    const double *d_c, double eta, double w0, int32_t max_iters,
                    int32_t check_every, double tol, double *d_scalars, void *stream);

/*
 * LP -> bipartite graph edges on the device, replacing the Python double loop of
 * build_graph_from_weights_sets (linear_program_methods.py:93-96): from DEVICE CSR arrays
 * (m rows) fills edge_index [2][nnz] int64 (row 0 = variable/column id, row 1 = constraint/row id,
 * CSR nonzero order) and edge_attr [nnz] float32 = a_ij.
 */
int mllp_graph_edges(int32_t m, int64_t nnz, const int32_t *d_indptr, const int32_t *d_indices,
                     const double *d_values, int64_t *d_edge_index, float *d_edge_attr, void *stream);

/*
 * The rule behind the reference's `_norm` arrays (dataset/netlib_mps_norm/<name>_{constrs.npz,rhs.npy,coefs.npy}, loaded at
 * linear_program_data.py:66-77; the generating script is not in the reference, the rule is restated and pinned against
 * the data in oracle/norm_rule.py): from the raw arrays (device CSR, m x n) and the row senses d_sense[m]
 * (0 = equality, +1 = "L": slack +1, -1 = "G": slack -1) build the standard form [A | S] -- one slack column per
 * inequality row, appended in row order, `nslack` of them (the caller counts the non-zero senses and sizes the outputs) --
 * scale row i by 1 / r_i, r_i = ||(a_i, slack_i)||_2 (squares added left to right in column order), or by 5 / b_i when
 * |b_i| / r_i > 5, and c by 1 / ||c||_2.  Outputs (device): CSR of m x (n + nslack) with nnz + nslack entries,
 * d_out_rhs[m], d_out_coefs[n + nslack], d_row_scale[m] (the factor of every row: y_raw = d_row_scale * y_norm) and
 * d_cnorm[1] = ||c||_2 (objective in the file's units = objective of the `_norm` LP x ||c||_2 + offset).  d_work:
 * mllp_norm_scale_work_bytes(m) bytes, 8-byte aligned.  Asynchronous on `stream`.
 */
int64_t mllp_norm_scale_work_bytes(int32_t m);
int mllp_norm_scale(int32_t m, int32_t n, int64_t nnz, int32_t nslack, const int32_t *d_indptr, const int32_t *d_indices,
                    const double *d_values, const int8_t *d_sense, const double *d_rhs, const double *d_coefs,
                    int32_t *d_out_indptr, int32_t *d_out_indices, double *d_out_values, double *d_out_rhs,
                    double *d_out_coefs, double *d_row_scale, double *d_cnorm, void *d_work, void *stream);

/*
 * Row partition of ONE large LP over `nranks` GPUs (one process per GPU; BASELINE.json
 * configs[3]: ken-18, osa-60, pds-20).  Every rank passes the whole matrix; rank p keeps the
 * rows of A assigned to it (balanced by nonzeros) -- its slice of y -- and ALL of A': the cheap
 * A' phase is replicated (every rank updates the whole of x from the whole of y), so only the y
 * slices cross GPUs, ONCE per iteration (the "all-gather once per iteration" of the north star; both A and
 * A' are stored, so no reduce-scatter is needed).  mllp_nccl_unique_id() is called on rank 0 and its 128
 * bytes are sent to the other ranks by the host (e.g. torch.distributed.broadcast) before the collective
 * mllp_lp_create_rowpart().  mllp_pdhg_run() on such a handle is a collective call: all ranks
 * pass the same full-length x, y, b, c (caller order) and the same num_iters, and all receive the full
 * result.  mllp_pdhg_solve / mllp_spmv / mllp_estimate_norm are not available on these handles.
 */
#define MLLP_NCCL_UNIQUE_ID_BYTES 128
int mllp_nccl_unique_id(unsigned char *out128);
int mllp_lp_create_rowpart(int32_t m, int32_t n, int64_t nnz, const int32_t *h_indptr,
                           const int32_t *h_indices, const double *h_values, const double *h_lb,
                           const double *h_ub, const double *h_ylo, const double *h_yhi, int device,
                           uint32_t flags, int32_t rank, int32_t nranks,
                           const unsigned char *uid128, mllp_lp_t *out);

/* In-kernel exchange for a row-partitioned handle: export the CUDA IPC handle of this rank's mailbox
 * (MLLP_IPC_HANDLE_BYTES), let the host all-gather them (nranks x 64 bytes, rank order) and import.
 * Afterwards mllp_pdhg_run() runs ALL iterations in one cooperative launch per rank: a row update stores
 * its new dual value straight into every peer's mailbox over NVLink as one tagged 16-byte word
 * {bits(y), bits(y) ^ tag} (value and validity in one access: no fence, no flag, no acknowledgement), and every
 * rank unpacks the peers' words into its own y before the local grid barrier that ends the iteration
 * (no NCCL call, no kernel launch inside the loop).  Waits time out (status via mllp_rowpart_error)
 * instead of hanging.  Without import the handle uses one NCCL all-gather per iteration between the
 * two launches of the iteration (environment MLLP_ROWPART_NCCL=1 forces that variant). */
#define MLLP_IPC_HANDLE_BYTES 64
int mllp_rowpart_ipc_export(mllp_lp_t lp, unsigned char *out64);
int mllp_rowpart_ipc_import(mllp_lp_t lp, const unsigned char *all_ranks);
/* 0 = no exchange wait has timed out on this rank (synchronises the device). */
int mllp_rowpart_error(mllp_lp_t lp, int32_t *out_flag);

/* NVSwitch multicast for the in-kernel exchange (instead of mllp_rowpart_ipc_*): the mailboxes of all ranks are bound into
 * ONE multicast object, so a row update issues a single `multimem.st` that the switch replicates into every GPU's mailbox
 * (with peer pointers every dual value is stored nranks - 1 times as a 16-byte NVLink write).  Collective protocol, driven by
 * the host (mllp_b200/distributed.py):
 *   1. every rank: mllp_rowpart_mc_supported (CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED); continue only if all say 1;
 *   2. rank 0: mllp_rowpart_mc_create -> a POSIX file descriptor of the multicast object; the other ranks duplicate it into
 *      their process (pidfd_getfd on rank 0's pid, or SCM_RIGHTS);
 *   3. every rank: mllp_rowpart_mc_attach(fd) (imports the object, adds this rank's device); barrier;
 *   4. every rank: mllp_rowpart_mc_bind (allocates the mailbox with the VMM API, binds it, maps it unicast for the local
 *      polls and multicast for the stores); barrier.  mllp_pdhg_run then uses the multicast exchange. */
int mllp_rowpart_mc_supported(mllp_lp_t lp, int32_t *out);
int mllp_rowpart_mc_create(mllp_lp_t lp, int32_t *out_fd);
int mllp_rowpart_mc_attach(mllp_lp_t lp, int32_t fd);
int mllp_rowpart_mc_bind(mllp_lp_t lp);

/*
 * Batched mode: `count` independent LPs in one launch (one CTA per LP, the whole LP held
 * in shared memory).  Instance k has shape m[k] x n[k]; its CSR arrays are the slices
 * [indptr_off[k], ...) of the concatenated host arrays, exactly as np.concatenate of the
 * loader's per-instance arrays.  Vectors are concatenated in instance order
 * (x: sum n[k], y: sum m[k]).  If `shared_matrix` != 0 only instance 0's matrix is given and
 * all `count` instances use it with their own b, c (the perturbed-instance workload).
 */
int mllp_batch_create(int32_t count, int32_t shared_matrix, const int32_t *h_m,
                      const int32_t *h_n, const int64_t *h_indptr_off,
                      const int64_t *h_nnz_off, const int32_t *h_indptr,
                      const int32_t *h_indices, const double *h_values, int device,
                      uint32_t flags, mllp_batch_t *out);
int mllp_batch_destroy(mllp_batch_t bt);
/* out[0]=count, [1]=sum m, [2]=sum n, [3]=sum nnz (shared: nnz of the one matrix), [4]=CTAs,
 * [5]=threads per CTA, [6]=dynamic smem bytes per CTA, [7]=algorithmic bytes per batch iteration,
 * [8]=instances per CTA of the parity kernel (shared matrix: R instances share every matrix step),
 * [9]=same for the solve kernel, [10]/[11]=shared-memory resident warp-steps of A / A' (parity kernel),
 * [12]/[13]=same (solve kernel), [14]=dynamic smem of the solve kernel, [15]=CTAs of the solve kernel. */
int mllp_batch_info(mllp_batch_t bt, int64_t *out16);
/* sigma_max(A_k) of every instance by power iteration (as mllp_estimate_norm) into the device
 * array d_sigma_max[count]; asynchronous on `stream`. */
int mllp_batch_estimate_norm(mllp_batch_t bt, int32_t iters, double *d_sigma_max, void *stream);
/* per-instance tau[k], sigma[k] (device arrays of `count`); scalars: count*MLLP_NUM_SCALARS */
int mllp_batch_run(mllp_batch_t bt, double *d_x, double *d_y, const double *d_b,
                   const double *d_c, const double *d_tau, const double *d_sigma,
                   int32_t num_iters, double *d_scalars, void *stream);
int mllp_batch_solve(mllp_batch_t bt, double *d_x, double *d_y, const double *d_b,
                     const double *d_c, const double *d_eta, double w0, int32_t max_iters,
                     int32_t check_every, double tol, double *d_scalars, void *stream);

/*
 * Bipartite message passing over the LP's nonzeros: the forward pass of the reference's GNNModel
 * (linear_program_methods.py:199-215, :238-251; five torch_geometric TransformerConv layers with heads = 1,
 * 16 channels, edge_dim = 1, root weight and bias, alternating along the rows of A' and of A, ReLU, Linear(16, 1)).
 * fp32 like the reference.  All pointers are device pointers.
 *
 * mllp_gnn_side: one direction of the graph as CSR of the DESTINATION side (rows = destination nodes: A' for the
 * constraint->variable "w2s" passes, A for the variable->constraint "s2w" passes); `values` are the fp64 coefficients
 * (cast to float per edge, as the reference's edge_attr).  `group` = lanes per destination row (1, 2, 4, 8, 16
 * or 32; about an eighth of the 90th-percentile row length, every lane walks its edges two at a time).  Rows with more
 * than `chunk` (>= 16) edges are listed in long_rows[nlong]; their pieces
 * are items[nitems][3] = (row, first edge, end edge), row r owning items long_first[r] .. long_first[r+1];
 * scratch holds 20 floats per item (16-byte aligned).
 */
typedef struct mllp_gnn_side {
    int32_t nd, ns, group, chunk;
    const int32_t *indptr;
    const int32_t *indices;
    const double *values;
    int32_t nlong, nitems;
    const int32_t *long_rows;
    const int32_t *long_first;
    const int32_t *items;
    float *scratch;
} mllp_gnn_side;

/* The layer is evaluated without materialising query / key / value rows (see mllp_b200/csrc/gnn_kernels.cu): an edge
 * gathers the feature row of its source node, the dense maps are applied once per destination node.  Parameter block
 * of one conv with din input channels (1 or 16), mllp_gnn_conv_param_floats(din) floats, formed on the host from the
 * module's weights (mllp_b200/gnn.py: pack_conv):
 *   MQ[din][din] = (Wq' Wk) / 4 | vq[din] = (Wk' bq) / 4 | wq[din] = (Wq' We) / 4 | sq = (We . bq) / 4 | pad to 4 floats |
 *   Wv'[din][16] | bv[16] | Ws'[din][16] | bs[16] | We[16]            (W' = transposed lin_*.weight, 1/4 = 1/sqrt(16))
 *
 * mllp_gnn_forward: the whole forward in one call (5 launches, + 2 per conv with cut rows): logit per variable into
 * d_out[n].  d_x1[n] = coefs, d_x2[m] = rhs as float32 (the reference's x1 / x2, :100-101).  d_params: the blocks of
 * gconv1_w2s, gconv1_s2w (din 1), gconv2_w2s, gconv2_s2w, gconv3_w2s (din 16), then fc.weight[16] | fc.bias (the final
 * Linear(16, 1) is folded into the last conv).  d_work: mllp_gnn_workspace_floats(n, m) floats, caller-owned, 16-byte
 * aligned.  Asynchronous on `stream`. */
int64_t mllp_gnn_conv_param_floats(int32_t din);
int64_t mllp_gnn_workspace_floats(int32_t n, int32_t m);
int mllp_gnn_forward(const mllp_gnn_side *to_var, const mllp_gnn_side *to_con, const float *d_x1, const float *d_x2,
                     const float *d_params, float *d_work, float *d_out, void *stream);

/* The same forward as a replayable plan: the launches are captured once into a CUDA graph (the two convs of a layer,
 * which are independent, on parallel branches), mllp_gnn_plan_run replays it on `stream` with one graph launch -- for the
 * small Netlib graphs the forward is launch-bound (afiro: ~10 launches of a few microseconds each).  The plan keeps the
 * POINTERS it was created with (sides, x1, x2, params, work, out): they must stay valid and in place, their contents may
 * change between runs.  Create on the device that owns the buffers. */
typedef struct mllp_gnn_plan *mllp_gnn_plan_t;
int mllp_gnn_plan_create(const mllp_gnn_side *to_var, const mllp_gnn_side *to_con, const float *d_x1, const float *d_x2,
                         const float *d_params, float *d_work, float *d_out, mllp_gnn_plan_t *out);
int mllp_gnn_plan_run(mllp_gnn_plan_t plan, void *stream);
int mllp_gnn_plan_destroy(mllp_gnn_plan_t plan);

/* One layer (exposed for unit parity): d_hout[nd][16] = [relu] TransformerConv(d_hsrc[ns][din] -> d_hdst[nd][din]) along
 * `side`, d_params = the conv's block as above. */
int mllp_gnn_conv(const mllp_gnn_side *side, int32_t din, const float *d_hdst, const float *d_hsrc, const float *d_params,
                  float *d_hout, int32_t relu, void *stream);

/* Backward pass of the same model: what loss.backward() does through GNNModel in the reference's training loop
 * (linear_program_experiment.py:115-157: model(graph) -> BCEWithLogitsLoss -> backward -> Adam step).
 *
 * The trainable parameters live in ONE flat float vector of mllp_gnn_flat_param_floats() entries: the six convs in the
 * module's order (gconv1_w2s, gconv1_s2w, gconv2_w2s, gconv2_s2w, gconv3_w2s, gconv3_s2w -- the last one is registered
 * but unused, linear_program_methods.py:247), each as torch_geometric registers its tensors
 *   lin_key.weight[16][din] | lin_key.bias[16] | lin_query.weight | lin_query.bias | lin_value.weight | lin_value.bias |
 *   lin_edge.weight[16] | lin_skip.weight | lin_skip.bias,
 * then fc.weight[16] | fc.bias.  mllp_gnn_pack_params forms, on the device, the parameter blocks mllp_gnn_forward takes
 * (mllp_gnn_packed_param_floats() floats) from the flat vector.
 *
 * mllp_gnn_backward: d_work is the workspace of the forward that was just run with d_packed on the same graph (it holds
 * the activations), d_bwork a scratch of mllp_gnn_backward_workspace_floats(n, m) floats (16-byte aligned), d_dout[n] =
 * d loss / d logit; writes d loss / d flat into d_dflat (all entries; gconv3_s2w and lin_key.bias get zeros).  No atomics:
 * the result is bitwise reproducible.  Asynchronous on `stream`. */
int64_t mllp_gnn_flat_param_floats(void);
int64_t mllp_gnn_packed_param_floats(void);
int mllp_gnn_pack_params(const float *d_flat, float *d_packed, void *stream);
int64_t mllp_gnn_backward_workspace_floats(int32_t n, int32_t m);
int mllp_gnn_backward(const mllp_gnn_side *to_var, const mllp_gnn_side *to_con, const float *d_x1, const float *d_x2,
                      const float *d_flat, const float *d_packed, const float *d_work, float *d_bwork,
                      const float *d_dout, float *d_dflat, void *stream);
/* The training step's two halves as replayable plans (CUDA graphs, run / destroyed with mllp_gnn_plan_run / _destroy; the
 * small Netlib graphs are launch-bound: the backward is 32 - 44 launches): mllp_gnn_train_plan_create = pack + forward into
 * d_out, mllp_gnn_backward_plan_create = pack + backward from d_dout into d_dflat.  The plans keep the POINTERS they were
 * created with; the contents (parameters, d_dout) may change between runs. */
int mllp_gnn_train_plan_create(const mllp_gnn_side *to_var, const mllp_gnn_side *to_con, const float *d_x1, const float *d_x2,
                               const float *d_flat, float *d_packed, float *d_work, float *d_out, mllp_gnn_plan_t *out);
int mllp_gnn_backward_plan_create(const mllp_gnn_side *to_var, const mllp_gnn_side *to_con, const float *d_x1,
                                  const float *d_x2, const float *d_flat, float *d_packed, const float *d_work,
                                  float *d_bwork, const float *d_dout, float *d_dflat, mllp_gnn_plan_t *out);

#ifdef __cplusplus
}
#endif
#endif /* MLLP_B200_H */
