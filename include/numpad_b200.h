/* numpad_b200 -- C-ABI of the B200 (sm_100a) engine behind numpad's hot path.
 *
 * The reference (qiqi/numpad) is pure Python and has no FFI of its own; the
 * seams this library replaces are the places where all of its arithmetic
 * funnels into scipy:
 *
 *   S1  adstate.py:282-297   _add_ops / _multiply_ops   -> scipy csr_plus_csr,
 *                                                           csr_matmat, _mul_scalar
 *   S2  adstate.py:244-272   dia_jac.tocsr / csr_jac.tocsr -> coo_tocsr
 *   S3  adsolve.py:193-195   splinalg.spsolve(J, rhs)    -> SuperLU
 *       adsolve.py:78,141    splinalg.factorized(...)    -> SuperLU
 *   S4  admpisolve.py:146-188,190-453  matvec / _dot / _norm2 / _lgmres
 *
 * Conventions
 *   - every function returns int: 0 = ok, < 0 = error (npb_last_error()).
 *   - pointers are DEVICE pointers owned by the caller unless marked [host];
 *     the library never allocates or frees device memory and keeps no state
 *     besides an error string, a launch counter and one pinned staging buffer.
 *   - `stream` is a cudaStream_t passed as void*; ordering is by stream only.
 *   - CSR = indptr int32[nrows+1], indices int32[nnz], data double[nnz], rows
 *     sorted by column and duplicate-free ("canonical").
 *   - two-phase calls: *_symbolic fills indptr_out and writes
 *     stats[0] = nnz of the result, stats[1] = its longest row (int64, device);
 *     the caller reads stats, allocates, then calls *_numeric.
 *   - `npb_scales`: chain of pending scalar factors applied (in order, one
 *     rounding each, like the reference's repeated `number * csr`) when the
 *     data are consumed.
 */
#ifndef NUMPAD_B200_H
#define NUMPAD_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NPB_OK 0
#define NPB_ERR_CUDA (-1)
#define NPB_ERR_ARG (-2)
#define NPB_ERR_WORKSPACE (-3)
#define NPB_ERR_BREAKDOWN (-4)
#define NPB_ERR_NOCONV (-5)
#define NPB_MAX_SCALES 6

typedef struct npb_scales {
    int n;
    double s[NPB_MAX_SCALES];
} npb_scales;

/* ---- library ---------------------------------------------------------- */
const char* npb_last_error(void);                 /* [host] */
int npb_version(void);
int npb_device_check(void);
unsigned long long npb_launch_count(void);
void npb_launch_count_reset(void);
/* algorithmic bytes (each operand once) of the SpMV-shaped and vector kernels of the solve
 * phase launched since the last reset (bench.py: whole-Krylov-phase roofline) */
double npb_bytes_count(void);
void npb_bytes_count_reset(void);
int64_t npb_scan_workspace_elems(int64_t n);      /* int64 elements of scan_ws */

/* ---- S1: number * csr (scipy _mul_scalar): materialise a scale chain --- */
int npb_csr_apply_scales(int64_t nnz, const double* in, const npb_scales* sc,
                         double* out, void* stream);

/* ---- S1: csr * dia_jac  /  dia_jac * csr (csr_matmat with a diagonal) -- */
int npb_csr_scale_cols(int64_t nnz, const int32_t* indices, const double* in,
                       const npb_scales* sc, const double* d, double* out,
                       int64_t* zero_count, void* stream);
int npb_csr_scale_rows(int64_t nrows, const int32_t* indptr, const double* in,
                       const npb_scales* sc, const double* d, double* out,
                       int64_t* zero_count, void* stream);

/* ---- csr_matmat drops entries whose value is exactly 0.0: every product
 * that can create zeros reports in *zero_count (int64, device, may be NULL)
 * how many it wrote; npb_csr_prune_* compacts them away when that is > 0. */
int npb_csr_prune_symbolic(int64_t nrows, const int32_t* indptr, const double* data,
                           const npb_scales* sc, int32_t* counts, int32_t* indptr_out,
                           int64_t* scan_ws, int64_t* stats, void* stream);
int npb_csr_prune_numeric(int64_t nrows, const int32_t* indptr, const int32_t* indices,
                          const double* data, const npb_scales* sc,
                          const int32_t* indptr_out, int32_t* indices_out, double* out,
                          void* stream);

/* ---- S1: csr * selection (adjoint order).  map[c] = new column of column c
 * or -1 (dropped); w = optional weight per old column.  `_same`: map keeps
 * every column and is increasing -> pattern shared, only indices rewritten. */
int npb_csr_relabel_same(int64_t nnz, const int32_t* indices, const double* in,
                         const npb_scales* sc, const int32_t* map,
                         const double* w, int32_t* indices_out, double* out,
                         int64_t* zero_count, void* stream);
int npb_csr_relabel_symbolic(int64_t nrows, const int32_t* indptr,
                             const int32_t* indices, const int32_t* map,
                             int32_t* counts, int32_t* indptr_out,
                             int64_t* scan_ws, int64_t* stats, void* stream);
int npb_csr_relabel_numeric(int64_t nrows, const int32_t* indptr,
                            const int32_t* indices, const double* in,
                            const npb_scales* sc, const int32_t* map,
                            const double* w, const int32_t* indptr_out,
                            int32_t* indices_out, double* out,
                            int64_t* zero_count, void* stream);

/* ---- S1: selection * csr (tangent order): row k of the result is
 * w[k] * B[map[k], :] (empty when map[k] < 0). */
int npb_csr_gather_rows_symbolic(int64_t nrows_out, const int32_t* map,
                                 const int32_t* b_indptr, int32_t* counts,
                                 int32_t* indptr_out, int64_t* scan_ws,
                                 int64_t* stats, void* stream);
int npb_csr_gather_rows_numeric(int64_t nrows_out, const int32_t* map,
                                const double* w, const int32_t* b_indptr,
                                const int32_t* b_indices, const double* b_data,
                                const npb_scales* sc, const int32_t* indptr_out,
                                int32_t* indices_out, double* out,
                                int64_t* zero_count, void* stream);

/* ---- S1: csr + csr (scipy csr_plus_csr; prune != 0 drops exact zeros) -- */
int npb_spadd_symbolic(int64_t nrows, const int32_t* a_indptr,
                       const int32_t* a_indices, const double* a_data,
                       const npb_scales* a_sc, const int32_t* b_indptr,
                       const int32_t* b_indices, const double* b_data,
                       const npb_scales* b_sc, int prune, int32_t* counts,
                       int32_t* indptr_out, int64_t* scan_ws, int64_t* stats,
                       void* stream);
int npb_spadd_numeric(int64_t nrows, const int32_t* a_indptr,
                      const int32_t* a_indices, const double* a_data,
                      const npb_scales* a_sc, const int32_t* b_indptr,
                      const int32_t* b_indices, const double* b_data,
                      const npb_scales* b_sc, int prune,
                      const int32_t* indptr_out, int32_t* indices_out,
                      double* out, void* stream);
/* Replays of a recorded assembly plan (trace-once / replay of the tape, _dev.AssemblyPlan):
 * fills into a RECORDED row structure.  Rows are written inside their recorded slots only;
 * *mismatch (device, reset by the call) counts the rows whose entry count differs from the
 * recording -- the caller discards the plan when the total of an assembly is not 0. */
int npb_spadd_numeric_checked(int64_t nrows, const int32_t* a_indptr,
                              const int32_t* a_indices, const double* a_data,
                              const npb_scales* a_sc, const int32_t* b_indptr,
                              const int32_t* b_indices, const double* b_data,
                              const npb_scales* b_sc, int prune, const int32_t* indptr_out,
                              int32_t* indices_out, double* out, int64_t* mismatch,
                              void* stream);
int npb_csr_prune_numeric_checked(int64_t nrows, const int32_t* indptr, const int32_t* indices,
                                  const double* data, const npb_scales* sc,
                                  const int32_t* indptr_out, int32_t* indices_out, double* out,
                                  int64_t* mismatch, void* stream);

/* ---- S1: general csr * csr (scipy csr_matmat), expansion to COO; finish
 * with npb_coo_to_csr_* (prune = 1: csr_matmat never stores a zero). */
/* row-batched SpGEMM with shared-memory expansion / sort / run sums (csrc/spgemm.cu): the
 * fast path of the general product when no row of C has more than cap / 2 partial products.
 * bound: ub[m], ubptr[m + 1], stats = {sum of ub, max ub}; numeric: brow[batches + 1],
 * scratch tidx / tval of `total` entries, cnt / tstart[m]; the caller scans cnt into
 * c_indptr, allocates C and calls compact. */
int npb_spgemm_rows_cap(void);
int npb_spgemm_set_warp_rows(int on);   /* tuning / test: 0 = rows of <= 512 products stay on the CTA-batched kernel; returns old */
int64_t npb_spgemm_rows_batches(int64_t total, int64_t max_ub);
int npb_spgemm_rows_bound(int64_t m, const int32_t* a_indptr, const int32_t* a_indices,
                          const int32_t* b_indptr, int32_t* ub, int32_t* ubptr,
                          int64_t* scan_ws, int64_t* stats, void* stream);
int npb_spgemm_rows_numeric(int64_t m, int64_t total, int64_t max_ub, const int32_t* a_indptr,
                            const int32_t* a_indices, const double* a_data,
                            const npb_scales* a_scales, const int32_t* b_indptr,
                            const int32_t* b_indices, const double* b_data,
                            const npb_scales* b_scales, const int32_t* ubptr, int prune,
                            int32_t* brow, int32_t* tidx, double* tval, int32_t* cnt,
                            int32_t* tstart, void* stream);
int npb_spgemm_rows_indptr(int64_t m, const int32_t* cnt, int32_t* c_indptr, int64_t* scan_ws,
                           int64_t* stats, void* stream);
int npb_spgemm_rows_compact(int64_t m, const int32_t* cnt, const int32_t* tstart,
                            const int32_t* c_indptr, const int32_t* tidx, const double* tval,
                            int32_t* c_indices, double* c_data, void* stream);
/* A times a selection with a non-monotone column map (map[c] < 0 drops column c, optional
 * weights): per-row relabel + stable sort + duplicate sums in scratch (nnz(A) entries), then
 * npb_spgemm_rows_indptr on cnt and npb_spgemm_rows_compact */
int npb_csr_relabel_merge_rows(int64_t m, const int32_t* indptr, const int32_t* indices,
                               const double* data, const npb_scales* scales, const int32_t* map,
                               const double* weights, int prune, int32_t* tidx, double* tval,
                               int32_t* cnt, int32_t* tstart, void* stream);
/* columns relabelled by an injective map, every (short) row re-sorted: the column half of a
 * symmetric permutation P A P^T (shard.py); indptr is shared with the input */
int npb_csr_relabel_sort_rows(int64_t m, const int32_t* indptr, const int32_t* indices,
                              const double* data, const int32_t* map, int32_t* indices_out,
                              double* data_out, void* stream);
int npb_spgemm_expand_symbolic(int64_t a_nnz, const int32_t* a_indices,
                               const int32_t* b_indptr, int32_t* counts,
                               int32_t* offsets, int64_t* scan_ws,
                               int64_t* stats, void* stream);
int npb_spgemm_expand_numeric(int64_t a_nnz, const int32_t* a_rows,
                              const int32_t* a_indices, const double* a_data,
                              const npb_scales* a_sc, const int32_t* b_indptr,
                              const int32_t* b_indices, const double* b_data,
                              const npb_scales* b_sc, const int32_t* offsets,
                              int32_t* out_rows, int32_t* out_cols,
                              double* out_vals, void* stream);

/* ---- S2: block -> CSR (dia_jac.tocsr, csr_jac.tocsr, sp.eye seeds) ----- */
int npb_diag_to_csr(int64_t n, const double* d /* NULL: scale*I */, double scale,
                    int32_t* indptr, int32_t* indices, double* data, void* stream);
int npb_sel_to_csr_symbolic(int64_t nrows, const int32_t* map, int32_t* counts,
                            int32_t* indptr_out, int64_t* scan_ws,
                            int64_t* stats, void* stream);
int npb_sel_to_csr_numeric(int64_t nrows, const int32_t* map, const double* w,
                           const int32_t* indptr_out, int32_t* indices_out,
                           double* out, void* stream);
int64_t npb_coo_to_csr_workspace(int64_t nrows, int64_t nnz);   /* bytes */
int npb_coo_to_csr_symbolic(int64_t nrows, int64_t ncols, int64_t nnz, const int32_t* rows,
                            const int32_t* cols, const double* vals, int prune,
                            void* ws, int64_t ws_bytes, int32_t* indptr_out,
                            int64_t* stats, void* stream);
int npb_coo_to_csr_numeric(int64_t nrows, int64_t ncols, int64_t nnz, void* ws,
                           int32_t* indices_out, double* data_out, void* stream);
int npb_csr_expand_rows(int64_t nrows, const int32_t* indptr, int32_t* rows,
                        void* stream);
int npb_csr_max_row(int64_t nrows, const int32_t* indptr, int64_t* stats,
                    void* stream);

/* ---- S3/S4: SpMV and Krylov vector kernels ----------------------------- */
int npb_spmv(int64_t nrows, int64_t nnz, const int32_t* indptr,
             const int32_t* indices, const double* data, const double* x,
             double* y, void* stream);                       /* y = A x     */
int npb_spmv_variant(int64_t nrows, int64_t nnz, const int32_t* indptr,
                     const int32_t* indices, const double* data, const double* x,
                     double* y, int variant /* 1 stream, 2 vector, 3 stream+cp.async,
                     4 stream+128-bit loads, 5 persistent TMA pipeline */,
                     void* stream);
int npb_spmv_set_tma_geometry(int g);   /* tuning: shared-memory ring geometry -1 (auto) or 0..3; returns old */
int npb_spmv_set_tma_f32_three(int on); /* tuning: fp32 operators on 3 resident CTAs per SM (default 1); returns old */
int npb_spmv_set_default_tma(int on);   /* tuning: 0 keeps npb_spmv off the TMA kernel; returns old */
/* diagonal storage for stencil operators (entries on <= 8 constant offsets col - row; offsets on the
 * host): conversion (*bad = entries on other offsets) and SpMV with the same epilogue modes */
int npb_csr_to_dia(int64_t nrows, const int32_t* indptr, const int32_t* indices, const double* data,
                   int ndiag, const int32_t* offsets, float* vals, int64_t ld, int64_t* bad,
                   void* stream);
int npb_spmv_dia_set_dot_ctas(int ctas_per_sm);      /* tuning switch: grid of the fused-dot DIA kernels, returns old */
int npb_spmv_dia_set_rows(int rows_per_thread);      /* tuning switch (1 or 2), returns the old value */
int npb_spmv_dia(int64_t nrows, int64_t ncols, int ndiag, const int32_t* offsets, const float* vals,
                 int64_t ld, const double* x, int mode, const double* b, const double* dinv,
                 double omega, double* y, void* stream);
/* hybrid image: diagonals + an ELL part (width <= 8, column-major like the diagonals) for the
 * entries on none of the offsets -- operators most of whose entries are on a few diagonals (the
 * momentum block of the cavity: 5 of 9 per row).  npb_csr_dia_remainder gives the ELL width. */
int npb_csr_dia_remainder(int64_t nrows, const int32_t* indptr, const int32_t* indices, int ndiag,
                          const int32_t* offsets, int32_t* maxw, void* stream);
int npb_csr_to_hyb(int64_t nrows, int64_t ncols, const int32_t* indptr, const int32_t* indices,
                   const double* data, int ndiag, const int32_t* offsets, float* vals, int64_t ld,
                   int ell_w, int32_t* ell_idx, float* ell_val, int64_t* bad, void* stream);
int npb_spmv_hyb(int64_t nrows, int64_t ncols, int ndiag, const int32_t* offsets, const float* vals,
                 int64_t ld, int ell_w, const int32_t* ell_idx, const float* ell_val, const double* x,
                 int mode, const double* b, const double* dinv, double omega, double* y, void* stream);
/* fused kernels of the preconditioned CG on a diagonal-storage operator (pressure solves of the
 * LSC preconditioner): the Jacobi sweep that ends a multigrid cycle also returns b . y; the
 * direction update p <- z + beta p, q <- L z + beta q (= L p) and p . q are one kernel
 * (beta = *rz_new / *rz_old on the device, first != 0: beta = 0).  red: reduction scratch. */
int npb_spmv_dia_jacobi_dot(int64_t nrows, int64_t ncols, int ndiag, const int32_t* offsets,
                            const float* vals, int64_t ld, const double* x, const double* b,
                            const double* dinv, double omega, double* y, double* dot_out,
                            void* red, void* stream);
int npb_dia_cg_dir(int64_t nrows, int ndiag, const int32_t* offsets, const float* vals, int64_t ld,
                   const double* z, const double* rz_new, const double* rz_old, int first,
                   double* p, double* q, double* pq, void* red, void* stream);
/* same epilogues for an operator whose values are stored in fp32 (vectors and sums fp64) */
int npb_spmv_f32(int64_t nrows, int64_t nnz, const int32_t* indptr, const int32_t* indices,
                 const float* data32, const double* x, int mode, const double* b,
                 const double* dinv, double omega, double* y, void* stream);
int npb_spmv_residual(int64_t nrows, int64_t nnz, const int32_t* indptr,
                      const int32_t* indices, const double* data,
                      const double* x, const double* b, double* y,
                      void* stream);                         /* y = b - A x */
/* y = epilogue(A x), kernel variant forced (0 = auto).  mode 0: y = Ax, 1: y = b - Ax,
 * 2: y = b + Ax, 3: y = x + omega * dinv .* (b - Ax) (damped-Jacobi sweep, y != x) */
int npb_spmv_epilogue(int64_t nrows, int64_t nnz, const int32_t* indptr,
                      const int32_t* indices, const double* data, const double* x,
                      int mode, const double* b, const double* dinv, double omega,
                      double* y, int variant, void* stream);
int64_t npb_reduce_scratch_bytes(void);   /* zero-initialised by the caller */
int npb_dot_multi(int64_t n, int k, const double* V, int64_t ld,
                  const double* w, double* out, void* scratch, void* stream);
int npb_axpy_multi(int64_t n, int k, const double* V, int64_t ld,
                   const double* h, double sign, double* w, double* out_norm2,
                   void* scratch, void* stream);
/* Gram-Schmidt with delayed re-orthogonalisation (npb_fgmres): both products of a step in one
 * sweep over the basis, out0[j] = V_j . w0 and out1[j] = V_j . w1 (j < k); and both updates in
 * one sweep, u <- (u - sum_j c_j V_j) * inv_rho, w <- w - sum_j a_j V_j - b * u_old with
 * *out_norm2 = ||w||^2 afterwards; coef = [c (k), a (k), inv_rho, b] on the device. */
int npb_dot_multi2(int64_t n, int k, const double* V, int64_t ld,
                   const double* w0, const double* w1, double* out0, double* out1,
                   void* scratch, void* stream);
int npb_axpy_multi2(int64_t n, int k, const double* V, int64_t ld,
                    const double* coef, double* u, double* w, double* out_norm2,
                    void* scratch, void* stream);
int npb_axpby(int64_t n, double a, const double* x, double b, double* y,
              void* stream);

/* ---- S3: banded LU with partial pivoting (robust direct solve for small
 * systems, stands where the reference calls SuperLU) + host RCM ordering -- */
int npb_rcm_host(int64_t n, const int32_t* indptr /*[host]*/,
                 const int32_t* indices /*[host]*/, int32_t* perm /*[host]*/,
                 int32_t* bw /*[host] lower, upper*/);
int npb_band_fill(int64_t n, const int32_t* indptr, const int32_t* indices,
                  const double* data, const int32_t* inv, int kl, int ku, double* AB,
                  int ldab, void* stream);
int npb_band_lu_factor(int n, int kl, int ku, double* AB, int ldab, int32_t* ipiv,
                       int32_t* info, void* stream);
int npb_band_lu_solve(int n, int kl, int ku, const double* AB, int ldab,
                      const int32_t* ipiv, double* b, void* stream);

/* Block-tridiagonal direct solver over the level sets of the matrix graph (csrc/blocktri.cu;
 * replaces spsolve, adsolve.py:193-195, where preconditioned Krylov fails: convection-
 * dominated saddle points, compressible-flow Jacobians at the scripts' time steps).  The
 * caller owns the dense blocks; "window" = rows [r0, r1) x columns [c0, c1) of a CSR matrix. */
int npb_bfs_levels_host(int64_t n, const int32_t* indptr /*[host]*/, const int32_t* indices /*[host]*/,
                        const int32_t* indptr_t /*[host] pattern of A^T, may be NULL*/,
                        const int32_t* indices_t /*[host]*/, int32_t* level /*[host] out*/,
                        int32_t* nlevels /*[host] out*/);
/* two BFS balls (around a pseudo-peripheral node and around the node farthest from it) and one
 * middle level between them: level order for the twisted factorisation (elimination from both
 * ends); has_diag [host, may be NULL]: 1 where the diagonal entry is non-zero (seeds are moved
 * onto such nodes); NPB_ERR_ARG for disconnected / tiny graphs */
int npb_bfs_levels_two_sided_host(int64_t n, const int32_t* indptr /*[host]*/,
                                  const int32_t* indices /*[host]*/, const int32_t* indptr_t /*[host]*/,
                                  const int32_t* indices_t /*[host]*/,
                                  const unsigned char* has_diag /*[host]*/, int32_t* level /*[host] out*/,
                                  int32_t* nlevels /*[host] out*/, int32_t* middle /*[host] out*/);
int npb_bt_window_to_dense(const int32_t* indptr, const int32_t* indices, const double* data,
                           int64_t r0, int64_t r1, int64_t c0, int64_t c1, double alpha,
                           double* out, int64_t ld, void* stream);
int npb_bt_window_spmm(const int32_t* indptr, const int32_t* indices, const double* data,
                       int64_t r0, int64_t r1, int64_t c0, int64_t c1, const double* B,
                       int64_t ldb, int64_t ncols, double alpha, double beta, double* C,
                       int64_t ldc, void* stream);
int npb_bt_window_spmv(const int32_t* indptr, const int32_t* indices, const double* data,
                       int64_t r0, int64_t r1, int64_t c0, int64_t c1, const double* x,
                       double alpha, double beta, const double* y0 /* NULL: y */, double* y,
                       void* stream);
/* both sweeps of a solve in one call (x = A^-1 b, or A^-T b when trans); off [host]: K + 1 level
 * offsets, sinv [host]: K device pointers to the row-major inverses; M: the level at which the
 * two elimination chains of a twisted factorisation meet (K - 1: one chain from the top);
 * c / scratch: 2 * mmax and 2 * npb_bt_gemv_scratch_doubles(mmax) doubles; with graph_id > 0 the
 * launch sequence is captured after the first call (the two chains as two branches) and
 * replayed (buffers must not move) */
int npb_bt_solve(int K, int M, const int64_t* off, const double* const* sinv, const int32_t* indptr,
                 const int32_t* indices, const double* data, const double* bp, double* z,
                 double* c, double* scratch, int64_t mmax, int trans, int graph_id, void* stream);
int npb_bt_graph_release(int graph_id);
int64_t npb_bt_gemv_scratch_doubles(int64_t m);
int npb_bt_gemv(int64_t m, const double* M, int64_t ld, const double* x, double alpha,
                double beta, const double* y0, double* y, int trans, double* scratch,
                void* stream);

/* ---- S3: the linear solve (replaces spsolve / factorized / _lgmres) ---- */
typedef int (*npb_apply_fn)(void* ctx, const double* x, double* y, void* stream);
typedef int (*npb_allreduce_fn)(void* ctx, double* buf, int count, void* stream);

typedef struct npb_krylov_ops {          /* [host] */
    npb_apply_fn matvec;   void* matvec_ctx;      /* y = A x                 */
    npb_apply_fn precond;  void* precond_ctx;     /* y = M^-1 x, may be NULL */
    npb_allreduce_fn allreduce; void* allreduce_ctx; /* NULL on one rank     */
} npb_krylov_ops;

typedef struct npb_csr_view {            /* [host] ctx of npb_csr_matvec_cb */
    int64_t nrows, nnz;
    const int32_t* indptr;
    const int32_t* indices;
    const double* data;
} npb_csr_view;

typedef struct npb_krylov_info {         /* [host] */
    int32_t iters, restarts, status, pad;
    double bnorm, rnorm0, rnorm;
} npb_krylov_info;

int npb_csr_matvec_cb(void* ctx, const double* x, double* y, void* stream);
/* in-situ timing of npb_csr_matvec_cb launches (CUDA events on the launching stream):
 * npb_profile_matvec(1) starts / resets, (0) stops; _read synchronises and sums */
int npb_profile_matvec(int on);
int npb_profile_matvec_read(double* total_ms, int64_t* count);
int64_t npb_fgmres_workspace_bytes(int64_t n, int restart, int flexible);
int npb_fgmres(int64_t n, const npb_krylov_ops* ops, const double* b, double* x,
               int restart, int maxiter, double rtol, double atol, int flexible,
               void* work, int64_t work_bytes, npb_krylov_info* info,
               void* stream);

/* BiCGStab, right-preconditioned, same callbacks and stopping rule (true residual) as
 * npb_fgmres: short recurrences, 8 work vectors; for Jacobians with a full diagonal
 * (test/euler_kec.py, test/navierstokes.py at moderate CFL).  info->restarts counts the
 * confirmations of the recurrence residual against the operator. */
int64_t npb_bicgstab_workspace_bytes(int64_t n);
int npb_bicgstab(int64_t n, const npb_krylov_ops* ops, const double* b, double* x,
                 int maxiter, double rtol, double atol, void* work, int64_t work_bytes,
                 npb_krylov_info* info, void* stream);
/* Block-Jacobi preconditioner: the unknowns of block c are rows c * cs + k * ks, k < bs <= 8
 * (cs = 1, ks = nblk: the reference's variable-major state arrays w[k, i, j]; cs = bs, ks = 1:
 * interleaved).  Setup inverts the bs x bs diagonal blocks (dinv: bs * bs * nblk doubles,
 * *info [device] = singular blocks, replaced by the identity); apply is an npb_apply_fn. */
typedef struct npb_block_jacobi {        /* [host] ctx of npb_block_jacobi_apply_cb */
    int64_t nblk;
    int32_t bs, pad;
    int64_t cs, ks;
    const double* dinv;
} npb_block_jacobi;
int npb_block_jacobi_setup(int64_t nblk, int bs, int64_t cs, int64_t ks, const int32_t* indptr,
                           const int32_t* indices, const double* data, double* dinv,
                           int32_t* info, void* stream);
int npb_block_jacobi_apply_cb(void* ctx, const double* x, double* y, void* stream);

/* ---- S3: algebraic multigrid + least-squares-commutator preconditioner --
 * Setup pieces (device): diagonal extraction, MIS aggregation, dense inverse
 * of the coarsest operator.  Solve phase: npb_amg_vcycle, and the two
 * npb_apply_fn callbacks usable as `precond` of npb_fgmres. */
/* Halo of a row-sharded operator (multi-GPU, one process per GPU).  Local column numbering:
 * [0, n_own) = columns this rank owns, n_own + q * maxb + k = k-th boundary value of rank q.
 * Before the local SpMV every rank packs x[send_idx] into its slot of `recv` and one
 * all-gather over the library's NCCL communicator fills the other slots. */
typedef struct npb_halo {
    const void* dist;               /* handle of npb_dist_init; NULL: not sharded */
    int64_t n_own;
    const int32_t* send_idx;        /* [n_send] local indices of the values to publish */
    int32_t n_send, maxb;
    double* recv;                   /* [world * maxb] */
    /* direct-store path: offsets into the symmetric heap (npb_dist_heap_alloc) of 2 x world x
     * maxb doubles and of world 64-bit arrival counters; < 0: NCCL all-gather into recv */
    int64_t p2p_data, p2p_flags;
    unsigned long long* seq;        /* DEVICE exchange counter of this operator (zero-initialised by
                                       the caller; read and bumped by the kernels, so that launch
                                       sequences do not depend on it and can be graph-captured) */
} npb_halo;

typedef struct npb_csr_mat {
    int64_t nrows, ncols, nnz;
    const int32_t* indptr;
    const int32_t* indices;
    const double* data;
    npb_halo halo;
    const float* data32;            /* != NULL: the same values in fp32 (preconditioner operators; the
                                       SpMV then reads 8 instead of 12 bytes per entry) */
    const float* dia_vals;          /* != NULL: diagonal-storage image (fp32, ndiag x nrows, from
                                       npb_csr_to_dia) of a stencil operator whose entries sit on
                                       ndiag <= 8 constant offsets col - row; preferred over CSR */
    int32_t ndiag, dia_ld;          /* dia_ld: floats between consecutive diagonals (>= nrows) */
    int32_t dia_off[8];
    const int32_t* ell_idx;         /* hybrid image (ell_w > 0): the entries on none of the diagonals, */
    const float* ell_val;           /* slot k of row r at [k * dia_ld + r] */
    int32_t ell_w, pad2;
} npb_csr_mat;

typedef struct npb_amg_level {      /* [host] */
    npb_csr_mat A, P, R;            /* operator, prolongation, restriction */
    const double* dinv;             /* 1 / diag(A) */
    double *x, *x2, *b, *r;         /* work vectors of the level's size */
    int32_t gather, pad;            /* != 0: last row-sharded level; the levels below are  */
    const int64_t* offsets;         /* replicated: [host] world + 1 slice bounds of their rhs */
    int64_t g_data, g_flags;        /* direct-store gather: heap offsets of 2 x offsets[world]  */
    unsigned long long* g_seq;      /* doubles / world counters (< 0: NCCL); DEVICE exchange counter */
} npb_amg_level;

typedef struct npb_amg {            /* [host] */
    int32_t nlevels, nu_pre, nu_post, coarse_n;
    double omega;
    const double* coarse_inv;       /* dense inverse of the coarsest operator */
    npb_amg_level* levels;
    const void* dist;               /* npb_dist_init handle when levels are row-sharded */
    int32_t gamma, gamma_from;      /* gamma > 1: levels >= gamma_from are visited gamma times per
                                       cycle (W-cycle on the coarse part of the hierarchy) */
} npb_amg;

typedef struct npb_amg_precond {    /* [host] ctx of npb_amg_apply_cb */
    const npb_amg* amg;
    npb_csr_mat A;
    int32_t cycles, pad;
    double *r, *dx;
} npb_amg_precond;

typedef struct npb_lsc {            /* [host] ctx of npb_lsc_apply_cb */
    int64_t n, nv, np;
    const int32_t* vset;            /* rows with a diagonal entry  */
    const int32_t* pset;            /* rows with a zero diagonal   */
    npb_csr_mat A, G, D, Lm;        /* J[V,V], J[V,P], J[P,V], -D G */
    const double* a_dinv;
    const npb_amg* amg_a;
    const npb_amg* amg_l;
    int32_t cycles_a, cycles_l;     /* A: stationary V-cycles; L: V-cycle-preconditioned CG steps
                                     * (first solve in the low 16 bits, second in the high 16; 0: same) */
    double* wv[5];
    double* wp[8];
    double* scal;                   /* 4 device doubles */
    void* red;                      /* npb_reduce_scratch_bytes(), zeroed */
    int32_t a_gmres;                /* > 0: A-solve = that many Jacobi-GMRES steps */
    int32_t graph_id;               /* > 0 (unique per preconditioner): npb_lsc_apply_cb may replay its
                                       launch sequence as a CUDA graph; release with npb_lsc_graph_release */
    void* a_work;                   /* npb_fgmres_workspace_bytes(nv, a_gmres, 0) */
    int64_t a_work_bytes;
} npb_lsc;

/* ---- multi-GPU: the library's own NCCL communicator (row-sharded solve, SURVEY 8e;
 * stands where the reference calls MPI in admpisolve.py:146-188).  nccl_path may be
 * NULL / "" (the copy torch has loaded is found by soname). */
int npb_dist_unique_id(const char* nccl_path, char* id128);        /* rank 0; host broadcasts */
int npb_dist_init(const char* nccl_path, const char* id128, int rank, int world,
                  void** handle);                                  /* collective */
int npb_dist_free(void* handle);
/* symmetric heap for exchanges by direct stores into NVLink peer memory (same layout on
 * every rank).  The memory is the CALLER's (`mem`, `bytes` long, kept alive until
 * npb_dist_free): the library allocates no device memory.  ipc72 = CUDA IPC handle of the
 * allocation containing `mem` (64 bytes) + offset of `mem` in it (8); all_handles = world x
 * 72 bytes in rank order */
int npb_dist_heap_create(void* handle, void* mem, int64_t bytes, char* ipc72);
int npb_dist_heap_open(void* handle, const char* all_handles);         /* collective */
int64_t npb_dist_heap_alloc(void* handle, int64_t bytes);              /* offset or < 0 */
int npb_dist_heap_reset(void* handle, void* stream);    /* between barriers: drop all allocs */
int npb_dist_rank(void* handle);
int npb_dist_world(void* handle);
int npb_dist_allreduce_cb(void* ctx, double* buf, int count, void* stream); /* npb_allreduce_fn */
int npb_mat_matvec_cb(void* ctx, const double* x, double* y, void* stream); /* ctx = npb_csr_mat */
int npb_mat_halo_exchange(void* ctx, const double* x, void* stream);   /* the exchange alone */

int npb_csr_diagonal(int64_t nrows, const int32_t* indptr, const int32_t* indices,
                     const double* data, double* diag, void* stream);
int npb_csr_filter_symbolic(int64_t nrows, const int32_t* indptr, const int32_t* indices,
                            const double* data, double theta, int32_t* counts,
                            int32_t* indptr_out, int64_t* scan_ws, int64_t* stats,
                            void* stream);
int npb_csr_filter_numeric(int64_t nrows, const int32_t* indptr, const int32_t* indices,
                           const double* data, double theta, const int32_t* indptr_out,
                           int32_t* indices_out, double* out, void* stream);
/* prolongator truncation: keep |p| >= theta * row max, rescale to the row sum */
int npb_csr_truncate_symbolic(int64_t nrows, const int32_t* indptr, const int32_t* indices,
                              const double* data, double theta, int32_t* counts,
                              int32_t* indptr_out, int64_t* scan_ws, int64_t* stats,
                              void* stream);
int npb_csr_truncate_numeric(int64_t nrows, const int32_t* indptr, const int32_t* indices,
                             const double* data, double theta, const int32_t* indptr_out,
                             int32_t* indices_out, double* out, void* stream);
int npb_amg_aggregate(int64_t n, const int32_t* indptr, const int32_t* indices,
                      const double* data, double theta, int max_rounds, int8_t* state,
                      int8_t* state2, int32_t* flag, int32_t* root_id, int32_t* agg,
                      int32_t* undecided, int64_t* scan_ws, int64_t* stats,
                      void* stream);
/* distance-2 MIS aggregation (aggregates of ~7 nodes on a 5-point stencil): keys = 2 n
 * uint64 scratch, state / near int8[n], label / label2 / flag / agg int32[n],
 * root_id int32[n + 1], undecided device int32; stats[0] = number of aggregates */
int npb_amg_aggregate_mis2(int64_t n, const int32_t* indptr, const int32_t* indices,
                           const double* data, double theta, int max_rounds, int8_t* state,
                           int8_t* near, uint64_t* keys, int32_t* label, int32_t* label2,
                           int32_t* flag, int32_t* root_id, int32_t* agg, int32_t* undecided,
                           int64_t* scan_ws, int64_t* stats, void* stream);
int npb_dense_inverse(int n, const int32_t* indptr, const int32_t* indices,
                      const double* data, double* W, double* Minv, int32_t* info,
                      void* stream);
int npb_gather(int64_t n, const int32_t* idx, const double* src, double* dst,
               void* stream);
/* CG steps with device-resident scalars: p = z + (rz_new/rz_old) p ;
 * x += (rz/pq) p, r -= (rz/pq) q */
int npb_cg_dir(int64_t n, const double* rz_new, const double* rz_old, int first,
               const double* z, double* p, void* stream);
int npb_cg_update(int64_t n, const double* rz, const double* pq, const double* p,
                  const double* q, double* x, double* r, int first, void* stream);
typedef struct npb_diag_ctx { int64_t n; const double* d; } npb_diag_ctx;   /* [host] */
int npb_diag_apply_cb(void* ctx, const double* x, double* y, void* stream);  /* y = d .* x */
int npb_amg_vcycle(const npb_amg* h, const double* b, double* x, void* stream);
int npb_amg_apply_cb(void* ctx, const double* x, double* y, void* stream);
int npb_lsc_apply_cb(void* ctx, const double* r, double* z, void* stream);
/* CUDA-graph replay of the preconditioner's launch sequence (single GPU, multigrid momentum
 * solve): global switch (returns the old value) and release of a preconditioner's graph */
int npb_lsc_graphs_enable(int on);
int npb_lsc_graph_release(int graph_id);

#ifdef __cplusplus
}
#endif
#endif /* NUMPAD_B200_H */
