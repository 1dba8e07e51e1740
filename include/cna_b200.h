/*
 * cna_b200 — C-ABI of the B200-native CNA hot path (NAM construction + permutation association).
 *
 * The reference (immunogenomics/cna v0.2.3) is pure Python: it has no FFI of its own, its hot
 * arithmetic lives in scipy/numpy/OpenBLAS calls made from src/cna/tools/_nam.py,
 * _association.py and _stats.py.  Each entry point below replaces one of those call sites; the
 * citation after "replaces:" is the reference file:line whose arithmetic the kernel reproduces.
 * INTEGRATION.md shows the ctypes stub a reference maintainer would add at each site.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer (cudaMalloc / torch storage) unless its name ends in _h;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); every launch is
 *     asynchronous on that stream, no entry point synchronises unless documented;
 *   - matrices are row-major with an explicit leading dimension `ld*` counted in elements;
 *   - the diffusion state / NAM is CELLS x SAMPLES (the transpose of the reference's DataFrames),
 *     fp32, leading dimension a multiple of 8 floats (32-byte sectors), padding columns zero;
 *   - cell-axis shards (one process per GPU): a shard passes its own rows of the CSR (column
 *     indices stay global) and `row_offset` = global index of its first row; n_rows is the number
 *     of local rows.  Single-GPU callers pass row_offset = 0;
 *   - return value 0 = success, otherwise an error code; cna_last_error() returns the message of
 *     the last failing call on the calling thread.
 */
#ifndef CNA_B200_H
#define CNA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CNA_B200_ABI_VERSION 5

/* bit pattern (a signalling NaN) of vector entries cna_median_f64 ignores: padding of gathered shards */
#define CNA_MEDIAN_SKIP_BITS 0x7FF4DEADBEEF0001ull

enum {
    CNA_OK = 0,
    CNA_ERR_INVALID = 1,  /* bad argument (shape, alignment, unsupported size) */
    CNA_ERR_CUDA = 2      /* a CUDA runtime call or kernel launch failed */
};

int cna_abi_version(void);
const char *cna_last_error(void);
/* number of kernels launched by this library since load (bench.py reports it as gpu_launches) */
int64_t cna_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * kernel (i): random-walk diffusion over the kNN graph
 * ------------------------------------------------------------------------------------------ */

/* Column sums of the CSR adjacency.  replaces: _nam.py:28 `a.sum(axis=0)`.
 * colsum[n_rows] (fp64) must be zeroed by the caller; data is fp64 (is_f64 != 0) or fp32. */
int cna_graph_colsum(const int32_t *indptr, const int32_t *indices, const void *data, int is_f64,
                     int64_t n_rows, double *colsum, void *stream);

/* Fold the input-side normalisation into the edges: vals[e] = A[i,j] / (colsum[j] + w),
 * diag[i] = w / (colsum[i] + w).  replaces: _nam.py:28,33 `s/colsums[:,None]`, `self_weight*s/colsums`.
 * out_f64 selects double outputs (used by cna.tl.diffuse on user vectors) or float. */
int cna_graph_scale(const int32_t *indptr, const int32_t *indices, const void *data, int is_f64,
                    int64_t n_rows, const double *colsum, double self_weight, void *vals,
                    void *diag, int out_f64, int64_t row_offset, void *stream);

/* First diffusion step when the state is the one-hot sample indicator (never materialised):
 * out[i, c] = sum_{j: code[j]=c} vals[i,j] + diag[i]*[code[i]=c].
 * replaces: _nam.py:51 `pd.get_dummies` + the first iteration of _nam.py:33. */
int cna_diffuse_onehot(const int32_t *indptr, const int32_t *indices, const float *vals,
                       const float *diag, const int32_t *code, int64_t n_rows, int n_samples,
                       float *out, int64_t ld, int64_t row_offset, void *stream);

/* One diffusion step on a dense fp32 state: out = A'.in + diag*in  (CSR SpMM).
 * replaces: _nam.py:33 `a.dot(s/colsums[:,None]) + self_weight*s/colsums[:,None]`
 * (scipy _sparsetools.csr_matvecs).  n_cols <= ld, ld % 4 == 0, in != out. */
int cna_diffuse_step_f32(const int32_t *indptr, const int32_t *indices, const float *vals,
                         const float *diag, const float *in, float *out, int64_t n_rows,
                         int n_cols, int64_t ld, int64_t in_row_offset, void *stream);

/* Same, fp64 state with any number of columns (cna.tl.diffuse / diffuse_stepwise on user input). */
/* cna_diffuse_step_f32 that also emits the QC statistic of _nam.py:78-82 for every finished row:
 * kurt[i] = Pearson kurtosis across batches of the per-batch means of out[i, :] * inv_count (what
 * cna_batch_kurtosis computes in a separate pass).  col_batch [ld] int8 = batch of each sample column
 * (-1 for padding), inv_count [ld] (0 for padding), batch_inv [n_batches] = 1 / samples per batch,
 * 2 <= n_batches <= 8.  replaces: _nam.py:33 of the last step + :78-82. */
int cna_diffuse_step_f32_qc(const int32_t *indptr, const int32_t *indices, const float *vals,
                            const float *diag, const float *in, float *out, int64_t n_rows, int n_cols,
                            int64_t ld, int64_t in_row_offset, const int8_t *col_batch,
                            const double *inv_count, const double *batch_inv, int n_batches, double *kurt,
                            void *stream);
int cna_diffuse_step_f64(const int32_t *indptr, const int32_t *indices, const double *vals,
                         const double *diag, const double *in, double *out, int64_t n_rows,
                         int n_cols, int64_t ld, int64_t in_row_offset, void *stream);

/* The same step with the gathered rows staged in shared memory (csrc/diffuse_tiled.cu): bit-identical
 * results, a source row crosses L2 -> SM once per tile instead of once per edge.  The plan is made once per
 * resident graph (cna_b200/tl/_graph.py:TilePlan): consecutive output rows are cut into tiles of at most
 * `tile_rows` rows whose edges reference at most `tile_sources` distinct source rows
 * (cna_diffuse_tile_limits); tile_row[n_tiles + 1] = first row of each tile; usrc = the distinct source rows
 * (positions in `in`, whose row count is in_rows) of every tile, each list padded to a multiple of 4 with
 * valid rows; tile_u[n_tiles + 1] = offsets into usrc; epair[nnz] = per stored edge, in CSR order,
 * {uint32 128 * (position of its source in the tile's list), float weight}.
 * stage_mode 0: cp.async.bulk.tensor tile::gather4 (TMA), 1: cp.async.
 * replaces: _nam.py:33 (as cna_diffuse_step_f32). */
int cna_diffuse_tile_limits(int32_t *tile_rows, int32_t *tile_sources);
int cna_diffuse_step_f32_tiled(const int32_t *indptr, const void *epair, const float *diag, const float *in,
                               float *out, int64_t n_rows, int64_t in_rows, int n_cols, int64_t ld,
                               int64_t in_row_offset, const int32_t *tile_row, const int32_t *tile_u,
                               const int32_t *usrc, int n_tiles, int stage_mode, void *stream);

/* Per-cell excess kurtosis (biased, Fisher) across samples of s[i,:]*inv_count[:].
 * replaces: _nam.py:59 `st.kurtosis(s/C, axis=1)`.  kurt[n_rows] fp64 (NaN where scipy gives NaN). */
int cna_row_kurtosis(const float *s, int64_t ld, int64_t n_rows, int n_samples,
                     const double *inv_count, double *kurt, void *stream);

/* ------------------------------------------------------------------------------------------
 * kernel (ii): QC, sample selection, residualisation, standardisation — one HBM pass
 * ------------------------------------------------------------------------------------------ */

/* Per-cell Pearson kurtosis across the per-batch means of s[i,:]*inv_count[:].
 * replaces: _nam.py:78-82 `_batch_kurtosis` inside `_qc_nam` (_nam.py:85-99).
 * Samples are grouped by batch through seg_order[n_sel] (column ids into the state, grouped so
 * that batch b owns seg_order[seg_off[b]..seg_off[b+1])), seg_off[n_batches] = n_sel; n_batches >= 2. */
int cna_batch_kurtosis(const float *s, int64_t ld, int64_t n_rows, const double *inv_count,
                       const int32_t *seg_order, const int32_t *seg_off, int n_batches, int n_sel,
                       double *kurt, void *stream);

typedef struct cna_resid_args {
    /* input state (cells x all samples) and its per-sample scaling 1/C_n (_nam.py:73) */
    const float *s;
    int64_t ld_s;
    int64_t n_rows;
    const double *inv_count;
    /* sample selection in phenotype order (_association.py:178-181): column ids, length n */
    const int32_t *colmap;
    int n;
    /* QC keep mask from cna_batch_kurtosis + threshold (_nam.py:96), may be NULL (keep all) */
    const uint8_t *row_keep;
    /* residualisation (_nam.py:128-156): X <- X - (X Wt^T) C^T with C [n x r], Wt [r x n]; r may be 0 */
    const double *C;
    const double *Wt;
    int r;
    /* batches of the selected samples for the ridge stopping rule (_nam.py:150-155):
     * positions 0..n-1 grouped by batch; n_batches <= 1 disables the kurtosis output */
    const int32_t *seg_order;
    const int32_t *seg_off;
    int n_batches;
    /* standardised phenotype (ddof=0, _association.py:22), length n */
    const double *y;
    /* outputs */
    float *x_out;        /* [n_rows x ld_x] residualised + standardised NAM (_nam.py:159), zero rows for dropped cells */
    int64_t ld_x;
    double *kurt;        /* [n_rows] batch kurtosis after residualisation (NaN for dropped cells), may be NULL */
    double *ncorr;       /* [n_rows] neighbourhood coefficients (_association.py:77), 0 for dropped cells */
    uint8_t *row_valid;  /* [n_rows] row_keep && variance > 0 (_association.py:182-185) */
    /* optional fp16 hi/lo planes of x_out for the tensor-core kernels ([n_rows x ld16] each,
     * ld16 a multiple of 16 with n <= ld16 <= 32 * ceil(n / 32)); NULL to skip.  x_out itself may
     * be NULL when only the planes are wanted. */
    void *x16_hi;
    void *x16_lo;
    int64_t ld16;
    /* QC keep decision taken inside the pass (used when row_keep is NULL): keep row i iff
     * qc_kurt[i] < max(6, 2 * qc_median[0]) (_nam.py:94-96; qc_median is what cna_median_f64 left on the
     * device, so no host round trip separates the QC from the residualisation).  Both NULL: keep all. */
    const double *qc_kurt;
    const double *qc_median;
    /* Optional output (needs the batch segments and qc_kurt == row_keep == NULL): the QC statistic itself,
     * qc_out[i] = Pearson kurtosis across the per-batch means of the selected, scaled raw row i (_nam.py:78-82) —
     * a by-product of the pass when every sample of the state is selected.  The pass then keeps every row;
     * the caller takes the median of qc_out and blanks the rows that fail with cna_qc_fixup, which saves the
     * QC epilogue of the last diffusion step. */
    double *qc_out;
} cna_resid_args;

/* replaces: _association.py:178-185 (reindex, filter, zero-variance drop), _nam.py:122 (centre),
 * :128-156 (M.NAM as a rank-r update), :159 (ddof=1 standardise), _association.py:77 (ncorrs). */
int cna_resid_pass(const cna_resid_args *args, void *stream);

/* Late QC decision for rows produced by cna_resid_pass with qc_out: rows with !(qc[i] < max(6, 2 * median[0]))
 * (_nam.py:94-96) get zero operand planes / x rows, ncorr = 0, valid = 0, kurt = NaN — what the pass writes
 * for a dropped cell.  Any of x, x16_hi / x16_lo, kurt may be NULL.
 * replaces: _nam.py:96-99 `keep = kurtoses < threshold; NAM.iloc[:, keep]`. */
int cna_qc_fixup(const double *qc, const double *median, int64_t n_rows, float *x, int64_t ld_x, void *x16_hi,
                 void *x16_lo, int64_t ld16, double *kurt, double *ncorr, uint8_t *row_valid, void *stream);

/* ------------------------------------------------------------------------------------------
 * kernel (iii): Gram matrix of the standardised NAM
 * ------------------------------------------------------------------------------------------ */

/* gram[n x n] (fp64, row-major, zeroed by the caller) += X^T X for X [n_rows x ld_x] fp32.
 * replaces: _nam.py:105 `NAM.dot(NAM.T)` (OpenBLAS dgemm).  Only ld_x % 4 == 0 and columns >= n
 * being zero padding are required. */
int cna_gram(const float *x, int64_t ld_x, int64_t n_rows, int n, double *gram, void *stream);
/* the fp32 CUDA-core implementation of the same contract (cross-check for the tensor-core kernel) */
int cna_gram_simt(const float *x, int64_t ld_x, int64_t n_rows, int n, double *gram, void *stream);

/* out[n_rows x ld_out] (fp32) = X . B for B [ld_x x ld_b] fp32 (rows >= n zero).
 * replaces: _nam.py:106 `NAM.T.dot(U) / sqrt(svs)` (the caller pre-divides U's columns). */
int cna_right_multiply(const float *x, int64_t ld_x, int64_t n_rows, int n, const float *b,
                       int64_t ld_b, int n_out, float *out, int64_t ld_out, void *stream);

/* ------------------------------------------------------------------------------------------
 * permutation engine
 * ------------------------------------------------------------------------------------------ */

/* For every permutation k (one warp each): z = y[perm[k,:]], zc = (I - C.W) z, zc /= std_ddof1(zc),
 * ssered[k] = |zc|^2, beta = Ut[:kmax] zc, ssefull[k, a] = |zc - U[:, :ks[a]] beta[:ks[a]]|^2.
 * The first n_local conditioned phenotypes are also written for the neighbourhood-level null
 * (_association.py:94-97): as fp32 columns of ycond [ld_x rows x ld_y] (may be NULL) and/or as
 * TRANSPOSED fp16 hi/lo planes yt [n_local x ld16] (may be NULL; padding columns are the caller's
 * zeros) for cna_null_hist_tc.  kmax = 0 skips the PC regressions (Ut, ks, ssefull may be NULL):
 * the conditioned phenotypes do not depend on U, so the null GEMM can start before the SVD is done.
 * replaces: _stats.py:18 `Y[bix]`, _association.py:35-61 (`_reg`, `_stats`, `_minp_stats` up to the
 * F statistic) evaluated in a Python loop at _association.py:84.
 *   perm  [K x n] int32 (row k = sample indices of permutation k)
 *   C [n x r], W [r x n] (last ridge stage only, _nam.py:169), Ut [kmax x n], ks [nks] ascending. */
int cna_perm_stats(const double *y, const int32_t *perm, int64_t K, int n, const double *C,
                   const double *W, int r, const double *Ut, int kmax, const int32_t *ks, int nks,
                   double *ssered, double *ssefull, float *ycond, int64_t ld_y, int n_local,
                   void *yt_hi, void *yt_lo, int64_t ld16, void *stream);

/* For every permutation k: p[k, a] = F survival function (scipy.special.fdtrc) of
 * f = ((ssered - ssefull)/ks[a]) / (ssefull/n) with (ks[a], n - 1 - r - ks[a]) degrees of freedom,
 * minp[k] = nanmin_a p[k, a], argk[k] = its index (first minimum), r2[k] = 1 - ssefull[k, argk]/ssered[k].
 * fp64 continued fraction of the incomplete beta function; agrees with scipy to ~1e-13 relative.
 * replaces: _association.py:45-47 and :53-60 evaluated per permutation at :84. */
int cna_perm_minp(const double *ssered, const double *ssefull, int64_t K, const int32_t *ks, int nks, int n,
                  int r, double *minp, int32_t *argk, double *r2, void *stream);

/* ------------------------------------------------------------------------------------------
 * neighbourhood-level null: GEMM with a threshold-histogram epilogue
 * ------------------------------------------------------------------------------------------ */

/* hist[k, b] (uint32, [n_null x n_edges], zeroed by caller) += #{cells i : edges[b] <= z_ik^2 <
 * edges[b+1]} with z = X.ycond/n, last bin closed at +inf, values below edges[0] dropped.
 * replaces: _association.py:99 `abs(NAMresid.T.dot(ycond_)/n)` + _stats.py:52-54 `np.histogram`
 * per null column (the N x Nnull matrix is never materialised).  `edge0` is the host copy of
 * edges[0] (lets the kernel reject sub-threshold products without touching the edge table). */
int cna_null_hist(const float *x, int64_t ld_x, int64_t n_rows, int n, const float *ycond,
                  int64_t ld_y, int n_null, const double *edges, int n_edges, double edge0,
                  uint32_t *hist, void *stream);

/* ------------------------------------------------------------------------------------------
 * tensor-core (tcgen05 / TMEM / TMA) versions of the dense contractions.  Operands are the fp16
 * hi/lo planes produced by cna_split_f16 (or directly by cna_resid_pass): x = hi + lo to 2^-22.
 * ------------------------------------------------------------------------------------------ */

/* dst planes [dst_rows x ld_dst] fp16: hi = fp16(v), lo = fp16(v - hi) with v = src[r][c]
 * (transpose = 0) or src[c][r] (transpose = 1); everything outside the source extent is zero. */
int cna_split_f16(const float *src, int64_t ld_src, int64_t src_rows, int src_cols, int transpose,
                  void *hi, void *lo, int64_t ld_dst, int64_t dst_rows, void *stream);

/* Same contract as cna_gram (gram += X^T X), X given as fp16 planes [n_rows x ld16], n <= 512.
 * `workspace` holds per-SM fp64 partial Grams (cna_gram_tc_workspace(n) bytes, contents
 * irrelevant on entry).  replaces: _nam.py:105 `NAM.dot(NAM.T)`. */
int64_t cna_gram_tc_workspace(int n);
int cna_gram_tc(const void *xh, const void *xl, int64_t ld16, int64_t n_rows, int n, double *gram,
                void *workspace, int64_t workspace_bytes, void *stream);

/* Cap on the number of CTAs of the persistent tensor-core kernels (cna_null_hist_tc*, cna_right_multiply_tc):
 * they launch one CTA per SM; a caller that runs a single-CTA kernel (cna_sym_eig_top) on another stream at the
 * same time leaves it an SM by passing (SM count - 1).  0 = no cap.  Returns the previous cap.  Process-wide.
 * replaces: nothing in the reference (scheduling only). */
int cna_tc_max_ctas(int cap);

/* Leading k eigenpairs of the symmetric n x n matrix (G + G^T) / 2 (the Gram), n <= 512, entirely on the
 * device in one launch: Householder tridiagonalisation, multisection on Sturm counts, inverse iteration
 * (dstein's clustering rule) and back-transformation, all fp64.  w_out[k]: eigenvalues, descending;
 * ut_out[k x n]: row c = unit eigenvector of the c-th largest eigenvalue (sign arbitrary — the association
 * test only forms projectors U_k U_k^T); de_out: optional [2 n + 4], diagonal and off-diagonal of the
 * tridiagonal form, then the SM clocks spent in the four phases (diagnostics).  `workspace`: cna_sym_eig_workspace(n, k) bytes, contents irrelevant on entry.
 * replaces: _nam.py:105 `U, svs, UT = np.linalg.svd(NAM.dot(NAM.T))` for the columns the association
 * test reads (_association.py:35-42 `U[:, :k]`, k <= max(ks)); the full decomposition of
 * return_full=True stays with LAPACK on the host. */
int64_t cna_sym_eig_workspace(int n, int k);
int cna_sym_eig_top(const double *G, int64_t ldg, int n, int k, double *w_out, double *ut_out, double *de_out,
                    void *workspace, int64_t workspace_bytes, void *stream);

/* Same contract as cna_right_multiply: out = X . B, with B given TRANSPOSED as fp16 planes
 * bt [n_out x ld16_b] (row j = column j of B).  replaces: _nam.py:106. */
int cna_right_multiply_tc(const void *xh, const void *xl, int64_t ld16, int64_t n_rows, int n,
                          const void *bth, const void *btl, int64_t ld16_b, int n_out, float *out,
                          int64_t ld_out, void *stream);

/* The histogram of cna_null_hist SUMMED over the null columns: hist[b] (uint64, [n_edges], zeroed
 * by the caller) += #{(cell i, null k): edges[b] <= z_ik^2 < edges[b+1]}, with the conditioned
 * phenotypes given TRANSPOSED as fp16 planes yt [n_null x ld16_y].  The reference's FDR is
 * mean_k(tails[k, i] / ranks[i]) (_stats.py:79-80), which only depends on sum_k tails[k, i], so the
 * per-null table never needs to exist; counts stay in shared memory until the kernel ends.
 * replaces: _association.py:99 + _stats.py:52-59 + :79-80. */
int cna_null_hist_tc(const void *xh, const void *xl, int64_t ld16, int64_t n_rows, int n,
                     const void *yth, const void *ytl, int64_t ld16_y, int n_null, const double *edges,
                     int n_edges, double edge0, uint64_t *hist, void *stream);

/* Histograms of the observed coefficients against the same edges and the strict thresholds:
 * rank_hist[b] += #{i valid: edges[b] <= ncorr_i^2 < edges[b+1]} (last bin closed),
 * det_hist[b]  += #{i valid: thresholds[b] < |ncorr_i| <= thresholds[b+1]} (last bin open above),
 * maxabs[0] = max |ncorr_i| must be computed first with cna_absmax.
 * replaces: _stats.py:73 (tail_counts of the observed z), _association.py:105-108. */
int cna_obs_hist(const double *ncorr, const uint8_t *row_valid, int64_t n_rows, const double *edges,
                 const double *thresholds, int n_edges, uint32_t *rank_hist, uint32_t *det_hist,
                 void *stream);

/* The same histograms with the number of edges left on the device by cna_fdr_thresholds (n_edges is
 * then the capacity of the tables; n_edges_dev == NULL: exactly cna_obs_hist). */
int cna_obs_hist_dev(const double *ncorr, const uint8_t *row_valid, int64_t n_rows, const double *edges,
                     const double *thresholds, int n_edges, const int32_t *n_edges_dev, uint32_t *rank_hist,
                     uint32_t *det_hist, void *stream);

/* cna_null_hist_tc with the number of edges (and the rejection bound derived from edges[0]) read from
 * device memory: the null GEMM can be queued behind cna_fdr_thresholds without the host ever seeing
 * max |ncorr|.  n_edges = capacity; edge0 is ignored when n_edges_dev != NULL. */
int cna_null_hist_tc_dev(const void *xh, const void *xl, int64_t ld16, int64_t n_rows, int n,
                         const void *yth, const void *ytl, int64_t ld16_y, int n_null, const double *edges,
                         int n_edges, const int32_t *n_edges_dev, double edge0, uint64_t *hist, void *stream);

/* np.median of a float64 device vector, left on the device: out[0] = median over the entries with
 * valid[i] != 0 (valid may be NULL) whose bit pattern is not CNA_MEDIAN_SKIP_BITS; NaN when one of them
 * is NaN or none exists (numpy's semantics); out[1] = number of entries considered.  Radix select on the
 * sortable bit pattern (4 passes of 16 bits + one pass for the lower median of an even population);
 * workspace of cna_median_workspace() bytes, contents irrelevant on entry.
 * replaces: _nam.py:59, :94, :153 `np.median(...)` (and the host round trip a host median implies). */
int64_t cna_median_workspace(void);
int cna_median_f64(const double *v, const uint8_t *valid, int64_t n, double *out, void *workspace,
                   int64_t workspace_bytes, void *stream);

/* thresholds[0..T) = np.arange(m/4, m, m/400) with m = max(maxabs[0], 0.001), edges[i] = t_i^2 - 1e-8 -
 * 1e-5 t_i^2, n_thresholds[0] = T (<= cap; entries beyond T are zero), computed with numpy's own sequence
 * of float64 operations.  replaces: _association.py:101-102, _stats.py:51. */
int cna_fdr_thresholds(const double *maxabs, int cap, double *thresholds, double *edges, int32_t *n_thresholds,
                       void *stream);

/* out[0] = max_i |v_i| over valid rows (0 if none).  replaces: _association.py:101. out zeroed by caller. */
int cna_absmax(const double *v, const uint8_t *row_valid, int64_t n_rows, double *out, void *stream);

/* Per-cell FDR: coef[i] = ncorr[i] (NaN for dropped cells); fdr[i] = prefix_min_fdr[idx-1] with
 * idx = #{thresholds <= |ncorr_i|}, or 1 when idx == 0 or the cell was dropped.
 * replaces: _association.py:228-237 (the per-cell `Series.apply(min_fdr_for_corr)`). */
int cna_cell_fdr(const double *ncorr, const uint8_t *row_valid, int64_t n_rows,
                 const double *thresholds, const double *prefix_min_fdr, int n_thr, double *coef,
                 double *fdr, void *stream);

/* cna_cell_fdr with the number of thresholds read from device memory (n_thr = capacity of the tables). */
int cna_cell_fdr_dev(const double *ncorr, const uint8_t *row_valid, int64_t n_rows,
                     const double *thresholds, const double *prefix_min_fdr, int n_thr, const int32_t *n_thr_dev,
                     double *coef, double *fdr, void *stream);

/* The FDR table on the device: fdr[i] = (reverse cumulative sum of null_hist)[i] / (reverse cumulative sum
 * of rank_hist)[i] / n_null for i < n_thresholds[0] (the float64 operations of _stats.py:79-80 applied to
 * the histogram summed over the nulls), prefix_min_fdr = its running minimum skipping NaN (what the
 * per-cell lookup of _association.py:234 needs).  With it the whole chain null GEMM -> FDR table -> per-cell
 * FDR column runs without the host.  cap <= 1024.  replaces: _stats.py:57-59, :79-80, _association.py:234. */
int cna_fdr_table(const uint64_t *null_hist, const uint32_t *rank_hist, const int32_t *n_thresholds, int cap,
                  int n_null, double *fdr, double *prefix_min_fdr, void *stream);

/* ------------------------------------------------------------------------------------------
 * locality-restoring cell order (no reference counterpart: a property of the HBM layout)
 * ------------------------------------------------------------------------------------------ */

/* One breadth-first level of a Cuthill-McKee ordering.  frontier[0..n_front) are the nodes of the
 * current level in their final order (positions pos_base + i).  Every neighbour with level == -1 is
 * claimed (level <- next_level, appended to next[], *next_count incremented) and first_parent[v] <-
 * min(first_parent[v], position of the parent).  The caller sorts next[] by (first_parent, id). */
int cna_bfs_expand(const int32_t *indptr, const int32_t *indices, const int32_t *frontier, int n_front,
                   int pos_base, int next_level, int32_t *level, int32_t *first_parent, int32_t *next,
                   int32_t *next_count, void *stream);

/* Helpers of the per-level sort: keys[i] = (first_parent[next[i]] << 32) | next[i]; after sorting the
 * keys, the node ids (low words) are written to order_out[0..n) and to the next frontier. */
int cna_bfs_keys(const int32_t *next, int n, const int32_t *first_parent, int64_t *keys, void *stream);
int cna_bfs_place(const int64_t *sorted_keys, int n, int64_t *order_out, int32_t *frontier, void *stream);

/* Symmetric permutation of a CSR: row i of the result is row order[i] of the input with its column
 * ids renamed through inv (inv[order[i]] = i); edges keep their order inside a row, so downstream
 * sums are performed in the original order.  new_indptr = prefix sums of the permuted row lengths. */
int cna_permute_csr(const int32_t *indptr, const int32_t *indices, const void *data, int is_f64,
                    const int64_t *order, const int32_t *inv, const int32_t *new_indptr, int64_t n_rows,
                    int32_t *new_indices, void *new_data, void *stream);

/* ------------------------------------------------------------------------------------------
 * host-side permutation drawing (no GPU involved; runs on the caller's host threads)
 * ------------------------------------------------------------------------------------------ */

/* out[0..count) = the next `count` deviates of numpy's legacy global generator, i.e. what
 * np.random.randn(count) would return for the RandomState (MT19937 key[624], pos, has_gauss,
 * cached gauss) given by np.random.get_state(); the state is advanced in place so that
 * np.random.set_state() continues the reference's stream.  n_threads <= 0: all host threads (<= 16).
 * replaces: the np.random.randn calls at _stats.py:12 and :31 (bit-exact). */
int cna_host_randn(uint32_t *key, int *pos, int *has_gauss, double *gauss, int64_t count, double *out,
                   int n_threads);

/* Permutation index matrix out [num x ld_out] int32 (row k = permutation k): for every block b
 * (rows block_off[b]..block_off[b+1]) draws np.random.randn(rows_b, num) from the legacy stream and
 * takes argsort(axis=0).  With src_pos (the concatenated sample positions of the blocks):
 * out[k, src_pos[r0+t]] = src_pos[r0 + argsort_k[t]] (conditional_permutation, _stats.py:8-16);
 * with src_pos == NULL: out[k, r0+t] = argsort_k[t] (grouplevel_permutation, _stats.py:31).
 * replaces: _stats.py:11-16 and :31 (bit-exact permutations). */
int cna_host_perm_blocks(uint32_t *key, int *pos, int *has_gauss, double *gauss, int n_blocks,
                         const int32_t *block_off, const int32_t *src_pos, int64_t num, int32_t *out,
                         int64_t ld_out, int n_threads);

/* cna_host_perm_blocks on the device (perm_dev.cu): the MT19937 recurrence in one CTA, every attempt of the
 * polar rejection loop tested in parallel, accepted pairs numbered by a scan, per-column argsort by counting.
 * key_host / pos / has_gauss / gauss: numpy's legacy state BEFORE the draw; block_off_host [n_blocks + 1],
 * src_pos_host [n] or NULL: HOST tables as in cna_host_perm_blocks (packed into a page-locked staging
 * buffer of the library and sent in one asynchronous copy: the call never waits for the stream).
 * out: device int32 [num x ld_out].  state_out (device, 626 uint32): key after the draw, pos, 1 if the
 * stream was long enough; tail (device, 4 doubles): attempt index, r2 and x1 of the last accepted pair
 * (an odd count leaves f * x1 cached in numpy's state: the caller finishes it with the host libm, so the
 * state is bit-exact); ambiguous (device int): 1 when two keys of a column are closer than 1e-14
 * relative, i.e. when the last place of log() could decide an argsort — the caller then repeats the draw
 * with cna_host_perm_blocks.  Integer stream, acceptance tests and uniforms are exact; permutations are
 * exact unless `ambiguous`.  workspace: cna_perm_draw_workspace(n, num, n_blocks) bytes.
 * replaces: _stats.py:11-16 and :31 (`np.argsort(np.random.randn(rows, num), axis=0)`). */
int64_t cna_perm_draw_workspace(int64_t n, int64_t num, int n_blocks);
int cna_perm_draw_device(const uint32_t *key_host, int pos, int has_gauss, double gauss, int n_blocks,
                         const int32_t *block_off_host, const int32_t *src_pos_host, int64_t num, int32_t *out,
                         int64_t ld_out, uint32_t *state_out, double *tail, int *ambiguous, void *workspace,
                         int64_t workspace_bytes, void *stream);

/* Asynchronous form of cna_host_perm_blocks: the draw runs on a thread owned by the library and
 * every buffer (including the four state words) must stay alive until cna_host_perm_wait, which
 * joins the thread, frees the handle and returns the status.  cna_host_perm_done polls (1 = finished). */
void *cna_host_perm_blocks_async(uint32_t *key, int *pos, int *has_gauss, double *gauss, int n_blocks,
                                 const int32_t *block_off, const int32_t *src_pos, int64_t num,
                                 int32_t *out, int64_t ld_out, int n_threads);
int cna_host_perm_done(void *handle);
int cna_host_perm_wait(void *handle);

/* ------------------------------------------------------------------------------------------
 * host -> device upload of pageable buffers (upload_host.cpp)
 * ------------------------------------------------------------------------------------------ */

/* dst (device) <- src (host), `bytes` bytes, ordered on `stream`.  Page-locked sources take one
 * cudaMemcpyAsync; pageable sources are staged by `n_threads` (<= 0: 4) host threads through a ring of
 * page-locked 4 MB slots owned by the library, one asynchronous copy per chunk, so that staging and DMA
 * overlap (a pageable cudaMemcpyAsync stages on one thread at ~10 GB/s).  Returns once the last chunk
 * has left `src` (the semantics of a pageable cudaMemcpyAsync).  The _async form runs on a thread of its
 * own: `src` must stay untouched until cna_host_upload_wait, which joins it and returns the status.
 * replaces: nothing in the reference (it never leaves the host); serves the graph read at _nam.py:12-19. */
int cna_host_upload(void *dst, const void *src, int64_t bytes, void *stream, int n_threads);
void *cna_host_upload_async(void *dst, const void *src, int64_t bytes, void *stream, int n_threads);
int cna_host_upload_wait(void *handle);

/* ------------------------------------------------------------------------------------------
 * kNN graph construction (cna_b200.pp.neighbors and the data generator; not on the timed path)
 * ------------------------------------------------------------------------------------------ */

/* Exact brute-force k nearest neighbours (self excluded) of fp32 points [n x dim], dim in {4, 8, 16, 32, 64}
 * (pad with zeros), k <= 64: candidates stream through shared memory in tiles, nothing N x N is formed.
 * idx [n x k] int32 ascending by distance, dist2 [n x k] squared distances.
 * replaces: the kNN search of scanpy.pp.neighbors that produces the graph read at _nam.py:12-19. */
int cna_knn_bruteforce(const float *points, int64_t n, int dim, int k, int32_t *idx, float *dist2,
                       void *stream);
/* The same search for the queries [q0, q0 + nq) only (idx / dist2 are [nq x k]): lets the ranks of a
 * multi-GPU run split the queries of one data set between them. */
int cna_knn_bruteforce_range(const float *points, int64_t n, int dim, int k, int64_t q0, int64_t nq,
                             int32_t *idx, float *dist2, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CNA_B200_H */
