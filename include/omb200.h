/*
 * omb200.h -- C ABI of libomb200.so: the B200 (sm_100a) snapshot-POD sparse-sensing hot path.
 *
 * The reference (OpenMEASURE v0.3.8) is pure Python and has NO FFI of its own: its hot path calls
 * numpy/scipy (OpenBLAS/LAPACK) directly from src/openmeasure/sparse_sensing.py.  Each entry point
 * below therefore cites the reference call site whose library routine it replaces; the Python
 * mirror of the reference classes (openmeasure_b200/sparse_sensing.py: ROM, SPR) binds them through
 * ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every pointer named d_* is a DEVICE pointer owned by the caller; h_* is a host pointer;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all calls are
 *     asynchronous on it unless stated, and never allocate device memory (workspaces are
 *     caller-provided, sized by the *_ws_bytes queries);
 *   - return value: 0 ok, <0 invalid argument, >0 cudaError_t; omb_last_error() gives the text;
 *   - snapshot layout: X is (F * n_c) x m, C-order FP64 (row = m contiguous snapshots), feature
 *     block f = rows [f*n_c, (f+1)*n_c)  (sparse_sensing.py:110);
 *   - basis layout: "tiled mode-major": the n candidate rows are cut into tiles of 128 and a tile
 *     stores its r modes back to back, Ut[(i/128)*r*128 + q*128 + i%128] = U_r[i, q]; the buffer
 *     holds ceil(n/128) tiles (padding rows are zero).  The candidates are the coalesced axis of
 *     every placement kernel and the trailing rows of a tile are one contiguous burst in HBM.
 */
#ifndef OMB200_H
#define OMB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ---------------------------------------------------------------------------- */
int omb_version(void);
const char* omb_last_error(void);
/* number of kernels this library has launched since load / since the last reset (bench.py's
 * gpu_launches claim is read from here, not estimated) */
int64_t omb_launch_count(void);
void omb_launch_count_reset(void);

/* scale_type codes for omb_finalize_scale (sparse_sensing.py:114-161) */
enum {
    OMB_SCALE_STD = 0, OMB_SCALE_NONE = 1, OMB_SCALE_PARETO = 2, OMB_SCALE_VAST = 3,
    OMB_SCALE_RANGE = 4, OMB_SCALE_LEVEL = 5, OMB_SCALE_MAX = 6, OMB_SCALE_VARIANCE = 7,
    OMB_SCALE_POISSON = 8, OMB_SCALE_L2NORM = 9
};

/* ---- measured FP64 tensor-pipe peak (no reference counterpart; bench.py's roofline denominator) --
 * Runs a register-only DMMA.8x8x4 stream for about ms_target milliseconds on every SM (synchronous:
 * returns after the kernel) and reports TFLOP/s and the measured duration.  d_scratch: at least
 * 2 * SMs * 256 doubles.  ~5 ms = burst figure, >= 300 ms = sustained under the power cap. */
int omb_fp64_peak(double ms_target, double* d_scratch, int64_t scratch_doubles, double* h_tflops,
                  double* h_ms, void* stream);

/* ---- synthetic snapshots (no reference counterpart; DESIGN.md "Synthetic workload") -------- */
int64_t omb_synth_ws_bytes(int64_t F, int64_t m, int64_t K);
int omb_synth_fill(double* d_X, int64_t F, int64_t n_cells, int64_t cell0, int64_t ncell_loc,
                   int64_t m, int64_t K, uint64_t seed, const double* d_amp, const double* d_dec,
                   double eps, void* d_ws, void* stream);

/* ---- K1: centring / scaling statistics (replaces np.average/np.std/np.max/np.min over the
 *      n_cells-row feature blocks, sparse_sensing.py:112-161) -------------------------------- */
/* Row means with numpy's pairwise tree (bit-exact): d_cnt[i] = mean(X[i, :]).  rows = F*n_c. */
int omb_row_means(const double* d_X, int64_t rows, int64_t m, double* d_cnt, void* stream);
/* Centred copy for the many-snapshot (m > 64) contraction kernels: d_X0c[i][j] = X[i][j] - cnt[i].
 * compute_means != 0: cnt[i] = np.average(X[i, :]) (bit-exact, written to d_cnt) from the same pass -- HBM sees
 * one read of X and one write of X0c; compute_means == 0: d_cnt is given (axis_cnt=None: block means).
 * Replaces the (X - X_cnt) temporary of sparse_sensing.py:169 for the tensor-core passes, which then run
 * without a single FP64 add in their inner loops (DADD shares the FP64 pipe with DMMA).  m even; X, X0c
 * 16-byte aligned. */
int omb_center_rows(const double* d_X, int64_t rows, int64_t m, int compute_means, double* d_cnt,
                    double* d_X0c, void* stream);
/* The same with a padded row pitch ld_out >= m (columns m .. ld_out-1 written as zeros) and any m: an odd snapshot
 * count gets an even pitch, so that the tensor-core kernels (16-byte aligned rows) serve it too -- a zero snapshot
 * changes neither the Gram of the first m columns nor the back-projection. */
int omb_center_rows_padded(const double* d_X, int64_t rows, int64_t m, int64_t ld_out, int compute_means,
                           double* d_cnt, double* d_X0c, void* stream);
/* Per-feature block reductions with numpy's pairwise tree over the n_c*m contiguous elements:
 *   mode 0: d_out[f*4 + {0,1,2}] = {sum, min, max}
 *   mode 1: d_out[f*4 + 3]       = sum((x - d_out[f*4+0]/mean_count)^2)   (np.std's second pass;
 *           mean_count = elements per block over ALL ranks, d_out[f*4+0] the global block sum)
 * d_ws: omb_block_stats_ws_bytes(F, n_c*m) bytes. */
int64_t omb_block_stats_ws_bytes(int64_t F, int64_t block_elems);
int omb_block_stats(const double* d_X, int64_t F, int64_t block_elems, int mode, int64_t mean_count,
                    double* d_out, void* d_ws, void* stream);
/* d_scl[f] from the block statistics (count = GLOBAL elements per block); when axis_cnt is
 * None (fill_cnt != 0) also fills d_cnt[f*n_c_loc .. ) with the block mean (sparse_sensing.py:112
 * with axis=None). */
int omb_finalize_scale(const double* d_stats, int64_t F, int64_t count, int scale_type,
                       double* d_scl, int fill_cnt, double* d_cnt, int64_t n_c_loc, void* stream);
/* X0 = (X - cnt) / scl materialised (sparse_sensing.py:169); only used when user code reads .X0 */
int omb_scale_rows(const double* d_X, int64_t F, int64_t n_c, int64_t m, const double* d_cnt,
                   const double* d_scl, double* d_X0, void* stream);

/* x = scl[f(i)] * x0 + cnt[i] for one length-n vector (replaces ROM.unscale_data's
 * cp.multiply(X_scl, x0) + X_cnt, sparse_sensing.py:235; two roundings like the reference) */
int omb_unscale(const double* d_x0, const double* d_cnt, const double* d_scl, int64_t n_c, int64_t n,
                double* d_out, void* stream);

/* ---- K3: POD contraction (replaces the tall-skinny part of np.linalg.svd, sparse_sensing.py:272)
 * Per-feature Gram of the centred rows: d_Gf[f] (m x m, row-major, full symmetric) =
 * sum_i (x_i - cnt_i)(x_i - cnt_i)^T over the rows of feature f.  d_cnt may be NULL (no centring).
 * Deterministic: fixed row split, fixed-order reduction, no atomics.  */
int64_t omb_gram_ws_bytes(int64_t F, int64_t n_c, int64_t m);
int omb_gram(const double* d_X, int64_t F, int64_t n_c, int64_t m, const double* d_cnt,
             double* d_Gf, void* d_ws, void* stream);
/* Row means (np.average(x, axis=1), sparse_sensing.py:112, bit-exact) written to d_cnt_out AND the
 * per-feature Grams of the rows centred by them, from ONE read of X (m <= 64; larger m runs
 * omb_row_means then omb_gram). */
int omb_gram_rowmeans(const double* d_X, int64_t F, int64_t n_c, int64_t m, double* d_cnt_out,
                      double* d_Gf, void* d_ws, void* stream);
/* G (m x m) = sum_f Gf[f] / scl[f]^2, fixed order. d_scl may be NULL (all ones). */
int omb_gram_combine(const double* d_Gf, int64_t F, int64_t m, const double* d_scl, double* d_G,
                     void* stream);

/* ---- S3: m x m symmetric eigensolve of the Gram matrix for few snapshots (m <= omb_eigh_max_m()):
 *      one-CTA cyclic Jacobi; replaces the small dense part of dgesdd (sparse_sensing.py:272).
 *      d_G row-major symmetric (upper triangle read), d_w eigenvalues DESCENDING, d_V[i*m + k] =
 *      component i of eigenvector k, d_info (may be NULL) = sweeps used. */
int omb_eigh_max_m(void);
int omb_eigh_jacobi(const double* d_G, int64_t m, double* d_w, double* d_V, int* d_info, void* stream);
/* sigma = sqrt(max(lambda, 0)) (m values) and W = V diag(1/sigma) (m x m, row-major like V) with a zero
 * column for every sigma <= rel_floor * sigma_1: the weights of the back-projection U = X0 W[:, :r]
 * (U[:, :r] of np.linalg.svd, sparse_sensing.py:272, :336). */
int omb_pod_weights(const double* d_w, const double* d_V, int64_t m, double rel_floor, double* d_S, double* d_W,
                    void* stream);

/* ---- K5: back-projection U_r = X0 * W, W = V_r Sigma_r^-1 (m x r row-major), written mode-major,
 *      with the initial QRCP column norms vn[i] = ||U_r[i,:]||_2 fused (replaces U = Q*U_R inside
 *      dgesdd and the dnrm2 initialisation of dgeqp3). d_vn may be NULL. -------------------- */
int omb_backproject(const double* d_X, int64_t F, int64_t n_c, int64_t m, const double* d_cnt,
                    const double* d_scl, const double* d_W, int64_t r, double* d_Ut, double* d_vn,
                    void* stream);

/* ---- K6: QR with column pivoting over the n candidate locations (replaces
 *      scipy.linalg.qr(self.Ur.T, pivoting=True) -> LAPACK dgeqp3/dlaqp2, sparse_sensing.py:739).
 * d_Ut   tiled mode-major basis (read only), d_vn initial norms (NULL -> computed here; else
 *        ceil(n/128)*128 doubles), d_work scratch of the same size as d_Ut for the trailing
 *        matrix, d_ws omb_qrcp_ws_bytes() bytes,
 * block  = steps (1..8) between trailing-matrix rewrites; 1 = LAPACK's unblocked dlaqp2
 *        arithmetic, bit for bit, for r <= 100,
 * outputs (device): d_piv[s] global row indices in selection order (+ index_base), d_rdiag[s] =
 * R[k,k], d_gap[s] = relative gap between the best and second-best candidate norm (degeneracy
 * meter).  optimal_placement(mask=...) zeroes the excluded rows of the basis beforehand, exactly
 * like the reference (:737-738).  One placement at a time per device. */
int64_t omb_qrcp_ws_bytes(int64_t n, int64_t r);
/* Lazy norm down-dates of the blocked schedule (block > 1; omb_qrcp and omb_qrcp_p2p).  dlaqp2's partial
 * norms only ever shrink, so a candidate whose norm at a block start is below the norm of the pivot that
 * is finally chosen cannot be that pivot: segments of 64 candidates whose largest norm is below
 * alpha * (pivot norm at the block start) sit the block's read-only passes out and take their down-dates
 * in the block-closing pass, which reads them anyway.  A pivot is accepted only if its norm reaches the
 * bound; otherwise the skipped segments that could beat it are brought up to date first (decided on the
 * device) -- the pivots are those of the eager schedule.  At a block boundary every column's norm is
 * recomputed exactly from the rows the block-closing pass writes (dlaqp2's recompute branch, taken
 * unconditionally), so eager and lazy runs hold bit-identical norms, exact ties included.
 * alpha in (0, 1) fixed, 0 = off, negative = automatic by problem size (the default, or $OMB_QR_LAZY);
 * returns the previous setting (-1 = automatic).  Process-wide.  Same LAPACK call site (:739). */
double omb_qrcp_set_lazy(double alpha);
/* Executed schedule of the last placement on this workspace (6 values): out[0] = (segment, row) visits
 * of the read-only passes (512 bytes each), out[1] = their segment visits (64 x 24 bytes of norms each),
 * out[2] = catch-up rounds, out[3] = 1 if the lazy scheme was on, out[4] = its alpha in parts per
 * million, out[5] = 0.  Synchronises the stream. */
int omb_qrcp_stats(const void* d_ws, int64_t n, int64_t* out, void* stream);
int omb_qrcp(const double* d_Ut, int64_t n, int64_t r, int64_t s, const double* d_vn, double* d_work,
             void* d_ws, int block, int64_t index_base, int64_t* d_piv, double* d_rdiag,
             double* d_gap, void* stream);

/* Multi-rank stepping interface (one process per GPU).  Rank g holds cells [cell0, cell0+n_c_loc) of
 * every feature (n = F*n_c_loc local rows); per pivot step the caller all-gathers one record of
 * omb_qrcp_record_doubles() doubles per rank (NCCL), everything else stays on the device:
 *     omb_qrcp_mr_start(...)                          norms, position maps, step-0 candidates
 *     for i in 0..s-1:  omb_qrcp_mr_local(i) -> d_rec ; all-gather -> d_recs ; omb_qrcp_mr_step(i)
 * Every rank takes the same decision (records are scanned in rank order, ties by LAPACK position)
 * and receives the same d_piv (GLOBAL row indices), d_rdiag, d_gap. */
int64_t omb_qrcp_record_doubles(void);
int omb_qrcp_mr_start(const double* d_Ut, int64_t n, int64_t r, int64_t s, const double* d_vn, void* d_ws,
                      int64_t n_c_loc, int64_t n_c, int64_t cell0, int rank, int world, void* stream);
int omb_qrcp_mr_local(const double* d_Ut, const double* d_work, int64_t n, int64_t r, void* d_ws,
                      int block, int64_t i, int64_t n_c_loc, int64_t n_c, int64_t cell0, int rank,
                      int world, double* d_rec, void* stream);
int omb_qrcp_mr_step(const double* d_Ut, double* d_work, int64_t n, int64_t r, int64_t s, void* d_ws,
                     int block, int64_t i, int64_t n_c_loc, int64_t n_c, int64_t cell0, int rank,
                     int world, const double* d_recs, int64_t* d_piv, double* d_rdiag, double* d_gap,
                     void* stream);

/* Same placement with the per-step exchange done by the kernels themselves over NVLink peer memory
 * (stores into every peer's symmetric buffer + a step flag; no NCCL call and no host round trip
 * inside the loop).  d_peers: device array of `world` buffer addresses (entry `rank` == d_mine), each
 * omb_qrcp_p2p_buffer_doubles(world) doubles, zero-filled once; epoch: strictly increasing per
 * call and identical on every rank.  After synchronising, a non-zero int64 at
 * d_mine + 2*world*record_doubles + 3*world means a peer never answered (10 s timeout). */
int64_t omb_qrcp_p2p_buffer_doubles(int world);
int64_t omb_qrcp_p2p_error_index(int world);   /* double index of the buffer's time-out flag */
int omb_qrcp_p2p(const double* d_Ut, int64_t n, int64_t r, int64_t s, const double* d_vn, double* d_work,
                 void* d_ws, int block, int64_t n_c_loc, int64_t n_c, int64_t cell0, int rank, int world,
                 const void* d_peers, double* d_mine, int64_t epoch, int64_t* d_piv, double* d_rdiag,
                 double* d_gap, void* stream);

/* ---- multi-GPU plumbing: all-gather of a small FP64 payload over NVLink peer memory (replaces an
 *      NCCL all-gather for the F*4 statistics, the m x m Gram and the sensor rows).
 *      d_peers: device array of `world` symmetric-buffer addresses (entry `rank` == d_mine), each
 *      omb_p2p_allgather_buffer_doubles(world, capacity) doubles, zero-filled once.  seq: strictly
 *      increasing per call, identical on every rank.  d_out: world x n, rank order. */
int64_t omb_p2p_allgather_buffer_doubles(int world, int64_t capacity);
int64_t omb_p2p_allgather_error_index(int world, int64_t capacity);
int omb_p2p_allgather(const double* d_src, int64_t n, double* d_out, const void* d_peers, double* d_mine,
                      int64_t capacity, int64_t seq, int rank, int world, void* stream);
/* All-reduce over the same buffers (shares the seq counter with omb_p2p_allgather): d_out[k] = the n
 * payload elements of all ranks combined IN RANK ORDER (identical bits on every rank).
 *   op 0: sum; op 1: groups of 4 = block statistics {sum, min, max, rank 0's}; op 2: element 3 of every
 *   group summed (np.std's second pass), the others rank 0's value (sparse_sensing.py:112-161 across ranks) */
int omb_p2p_allreduce(const double* d_src, int64_t n, double* d_out, const void* d_peers, double* d_mine,
                      int64_t capacity, int64_t seq, int rank, int world, int op, void* stream);

/* ---- GEM: greedy entropy-maximisation placement (SPR.gem, sparse_sensing.py:586-698; reached via
 *      optimal_placement(calc_type='gem') :745-751).  One streaming pass over the basis per sensor.
 *      omb_gem_variance: d_var[j] = np.var(Ur[j, :], ddof=1)                              (:621, :639)
 *      omb_gem_step:     argmax over live candidates of
 *                        coef^2 var[j] - Sigma_ya B Sigma_ay   (k chosen rows Z (k x r, centred and
 *                        scaled by coef), B = inverse covariance of the chosen rows, k x k; :670-681);
 *                        writes the winner's index (-1: none alive), value and UNSCALED basis row.
 *      omb_gem_exclude:  alive[j] &= ||xyz[j % n_c] - xyz[sensor % n_c]|| >= d_min         (:646-649) */
int64_t omb_gem_ws_bytes(void);
int omb_gem_max_sensors(void);
int omb_gem_variance(const double* d_Ut, int64_t n, int64_t r, double* d_var, void* stream);
int omb_gem_step(const double* d_Ut, int64_t n, int64_t r, double coef, int64_t k, const double* d_Z,
                 const double* d_B, const double* d_var, const unsigned char* d_alive, void* d_ws,
                 int64_t* d_idx, double* d_val, double* d_row, void* stream);
int omb_gem_exclude(const double* d_xyz, int64_t n_c, int64_t n, const int64_t* d_sensor, double d_min,
                    unsigned char* d_alive, void* stream);
/* the same exclusion around a point given by its coordinates (row-sharded runs: the chosen cell may be a peer's) */
int omb_gem_exclude_point(const double* d_xyz, int64_t n_c, int64_t n, double px, double py, double pz,
                          double d_min, unsigned char* d_alive, void* stream);

/* ---- weighted OLS predict, batched (replaces the per-vector pinv(W Theta) loop of SPR.predict for
 *      measurements with non-zero uncertainties, sparse_sensing.py:871-878): for vector b
 *        W = diag(1 / y0s[b]);  Ar[b] = argmin || W Theta a - W y0v[b] ||;  Asig[b] = | argmin || W Theta a - y0s[b] || |
 *      by Householder QR of W Theta (s x r, s >= r) in shared memory, one CTA per vector.  flag[b] = 1 (and no
 *      output) when min|R_kk| <= rank_tol * max|R_kk|: not full column rank, take the pseudo-inverse route.
 *      omb_wols_smem_bytes(s, r) must not exceed 227 KB. */
int64_t omb_wols_smem_bytes(int64_t s, int64_t r);
int omb_wols_predict(const double* d_Theta, int64_t s, int64_t r, const double* d_y0v, const double* d_y0s,
                     int64_t N, double rank_tol, double* d_Ar, double* d_Asig, int* d_flag, void* stream);

/* ---- general (non-one-hot) measurement / sampling matrices in CSR form: replaces
 *      Theta = C.dot(Ur) (sparse_sensing.py:797), C.dot(X_cnt) (:573) and the sampled scale / centre of
 *      unscale_data / reconstruct(sampling=) (:233, :365-368) for scipy CSR matrices such as the line-of-sight
 *      matrices of utils.camera.project (utils.py:318-469).  Never densified.  indices are LOCAL row indices
 *      (row-sharded runs pass the columns of C that fall on the rank's rows and sum the partial results in rank
 *      order).  d_Theta (s x r), d_cnt_s (s), d_scl_s (s): any may be NULL.  d_ws: omb_csr_ws_bytes(). */
int64_t omb_csr_ws_bytes(int64_t s, int64_t max_row_nnz, int64_t r);
int omb_csr_times_basis(const int64_t* d_indptr, const int64_t* d_indices, const double* d_data, int64_t s,
                        int64_t max_row_nnz, const double* d_Ut, int64_t n, int64_t r, const double* d_cnt,
                        const double* d_scl, int64_t n_c, double* d_Theta, double* d_cnt_s, double* d_scl_s,
                        void* d_ws, void* stream);

/* ---- K8/K9: train = row gather (replaces the dense C.dot(Ur), C.dot(X_cnt);
 *      sparse_sensing.py:797, :573).  d_Theta is s x r row-major, d_cnt_s may be NULL. -------- */
int omb_gather_rows(const double* d_Ut, int64_t r, const int64_t* d_piv, int64_t s,
                    double* d_Theta, const double* d_cnt, double* d_cnt_s, void* stream);
/* Ur (n x r, C-order) <- Ut, and back (for the .Ur attribute / fit(basis=...) / mask) */
/* d_dst (tiled, r_dst modes) = the first r_dst modes of d_src (tiled, r_src modes): drops the zero padding mode an
 * odd mode count is back-projected with (U[:, :r], sparse_sensing.py:336) */
int omb_copy_modes(const double* d_src, int64_t r_src, double* d_dst, int64_t r_dst, int64_t n, void* stream);
int omb_modes_to_rows(const double* d_Ut, int64_t n, int64_t r, double* d_Ur, void* stream);
int omb_rows_to_modes(const double* d_Ur, int64_t n, int64_t r, double* d_Ut, double* d_vn,
                      void* stream);

/* ---- K10: batched OLS predict (replaces the per-vector np.linalg.pinv + dot loop,
 *      sparse_sensing.py:865-878 with all sigma == 0):
 *      A (N x r) = ((Y - cnt_s) / scl_s) * PinvT,  Y N x s row-major, PinvT s x r row-major ---- */
int omb_ols_predict(const double* d_Y, const double* d_cnt_s, const double* d_scl_s,
                    const double* d_PinvT, int64_t N, int64_t s, int64_t r, double* d_A,
                    void* stream);

/* ---- K11: reconstruct rows [row0, row0+nrows) of X_rec = U_r A^T, unscaled in the epilogue
 *      out[i, k] = scl[f(i)] * acc + cnt[i]  (replaces Ur @ Ar.T + the per-column unscale_data,
 *      sparse_sensing.py:371-373, :235).  d_out is nrows x N row-major.  d_cnt/d_scl may be NULL. */
int omb_reconstruct(const double* d_Ut, int64_t n, int64_t r, const double* d_A, int64_t N,
                    const double* d_cnt, const double* d_scl, int64_t n_c, int64_t row0,
                    int64_t nrows, double* d_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OMB200_H */
