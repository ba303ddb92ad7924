/*
 * spuq_sg.h -- C ABI of libspuq_sg.so: the B200 (sm_100a) implementation of spuq's
 * stochastic-Galerkin operator application and the PCG solver around it.
 *
 * The reference (aleadev/spuq) is pure Python and has no FFI for this path; the boundary it
 * offers is its Python operator protocol.  Every entry point below therefore names the
 * reference *Python* interface it replaces (paths relative to the reference root).  The
 * reference-side binding (a ctypes stub) is shown in INTEGRATION.md and shipped as
 * spuq_b200/_abi.py.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only.  No exceptions cross the boundary: every
 *     function returns 0 on success or a negative SPUQ_E_* code; spuq_sg_last_error() returns
 *     a thread-local, human-readable message for the last failure.
 *   - The caller owns every device buffer it passes (PyTorch allocates them); the library owns
 *     only what hangs off its opaque handles.  "_dev" = device pointer, "_host" = host pointer.
 *   - Slabs are column-major N x L fp64: element (i, mu) of W lives at W[i + ldw * mu]; column
 *     mu is the spatial vector w[mu] with mu the mu-th index of Lambda in the reference's
 *     sorted order (MultiVector.active_indices(), application/egsz/multi_vector.py:97-98).
 *   - A handle is not thread-safe; different handles may be used from different threads.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 */
#ifndef SPUQ_SG_H
#define SPUQ_SG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPUQ_SG_ABI_VERSION 1

#if defined(__GNUC__)
#define SPUQ_API __attribute__((visibility("default")))
#else
#define SPUQ_API
#endif

/* error codes */
#define SPUQ_OK              0
#define SPUQ_E_INVALID      -1   /* bad argument                                   */
#define SPUQ_E_CUDA         -2   /* a CUDA runtime call failed (see last_error)    */
#define SPUQ_E_NOMEM        -3
#define SPUQ_E_UNSUPPORTED  -4   /* e.g. requested kernel path not applicable      */

/* PCG status words (spuq_sg_pcg*, *status_out); mapped by the Python side onto the
 * reference's four exception messages, application/egsz/pcg.py:33-36,42-45,53 */
#define SPUQ_PCG_RUNNING        -1
#define SPUQ_PCG_CONVERGED       0
#define SPUQ_PCG_PRECOND_NOT_PD  1   /* zeta < 0                     pcg.py:33-36 */
#define SPUQ_PCG_SINGULAR        2   /* <Av, v> == 0                 pcg.py:42-43 */
#define SPUQ_PCG_NOT_PD          3   /* <Av, v> < 0                  pcg.py:44-45 */
#define SPUQ_PCG_NOT_CONVERGED   4   /* loop exhausted               pcg.py:53    */

/* kernel-path selection for spuq_sg_plan_create(flags) */
#define SPUQ_PATH_AUTO     0   /* stencil path when the pattern allows it, else ELL, else CSR */
#define SPUQ_PATH_CSR      1   /* general CSR, one thread per (row, column)                   */
#define SPUQ_PATH_ELL      2   /* padded ELL copy of the values, coalesced                     */
#define SPUQ_PATH_STENCIL  3   /* 9-point-window diagonal storage: TMA-fed, accumulators in TMEM */
#define SPUQ_PATH_MASK     0xF

typedef struct spuq_sg_plan spuq_sg_plan;
typedef struct spuq_sg_precond spuq_sg_precond;

SPUQ_API const char* spuq_sg_last_error(void);
SPUQ_API int spuq_sg_abi_version(void);

/* ---------------------------------------------------------------------------------------------
 * K5: integer set-up of the active index set Lambda (host code, bit-exact with the reference).
 * ------------------------------------------------------------------------------------------- */

/* Sort L multi-indices (rows of `mi_host`, row-major L x M, non-negative) into the reference's
 * order: total degree first, then entries left to right, lower entry first
 * (Multiindex.cmp_by_order, math_utils/multiindex.py:42-53, used by sorted() in
 * MultiVector.active_indices(), application/egsz/multi_vector.py:97-98).
 * perm_out[k] = row of mi_host that comes k-th.  Duplicate rows are an error. */
SPUQ_API int spuq_sg_lambda_sort(const int32_t* mi_host, int32_t L, int32_t M, int32_t* perm_out);

/* Neighbour tables over a sorted Lambda (rows of mi_sorted_host):
 * nbr_up[m*L + k] = position of mu_k + e_m in Lambda, nbr_dn[m*L + k] = position of mu_k - e_m,
 * -1 when that index is not in Lambda (or would be negative).  Replaces the list scans
 * `mu.inc(m) in Lambda` / `mu.dec(m) in Lambda`, application/egsz/multi_operator2.py:72-79
 * (Multiindex.inc/dec, math_utils/multiindex.py:103-127). */
SPUQ_API int spuq_sg_lambda_neighbours(const int32_t* mi_sorted_host, int32_t L, int32_t M,
                              int32_t* nbr_up_out, int32_t* nbr_dn_out);

/* ---------------------------------------------------------------------------------------------
 * K1: the SG operator  y[mu] = A_0 w[mu] + sum_{m in [m_begin, m_end)} A_m ( beta_up[m,mu] w[mu+e_m]
 *                               - beta_0[m,mu] w[mu] + beta_dn[m,mu] w[mu-e_m] )
 * Replaces MultiOperator.__init__ / MultiOperator.apply,
 * application/egsz/multi_operator2.py:30-36 and :38-84 (multi_operator.py:42-112 for equal bases).
 * ------------------------------------------------------------------------------------------- */

/* Build a plan.
 *   n_rows, n_cols   shape of every spatial block (square for the SG operator; a rectangular
 *                    block is allowed when M == 0, used for the leaf operators of
 *                    linalg/operator.py:365-368,400-403 and linalg/scipy_operator.py:29-33)
 *   L, M             |Lambda| and the number of expansion terms (= maxm of multi_operator2.py:44-48)
 *   rowptr_dev[n_rows+1], colidx_dev[nnz]   ONE CSR pattern shared by all blocks (int32)
 *   vals_host[M+1]   host array of device pointers: vals_host[0] -> values of A_0 (mean),
 *                    vals_host[1+m] -> values of A_m, each nnz doubles on the shared pattern.
 *                    Entries for m outside [m_begin, m_end) may be NULL.
 *   nbr_up/nbr_dn    host, M x L int32, from spuq_sg_lambda_neighbours
 *   beta_up/dn/0     host, M x L fp64: beta[1], beta[-1], beta[0] of
 *                    rv_m.orth_polys.get_beta(mu[m]) (polyquad/polynomials.py:44-48) per (m, mu)
 *   m_begin, m_end   this plan applies only terms m in [m_begin, m_end) (M-sharding over GPUs)
 *   include_mean     non-zero: add the A_0 w[mu] term (exactly one shard does)
 *   flags            SPUQ_PATH_* in the low bits
 * The CSR arrays must stay valid while the plan lives on the CSR path; the ELL and stencil paths
 * keep a re-laid-out private copy of the values and do not touch vals again after creation. */
SPUQ_API int spuq_sg_plan_create(spuq_sg_plan** plan_out, int device,
                        int64_t n_rows, int64_t n_cols, int32_t L, int32_t M,
                        const int32_t* rowptr_dev, const int32_t* colidx_dev,
                        const double* const* vals_host,
                        const int32_t* nbr_up_host, const int32_t* nbr_dn_host,
                        const double* beta_up_host, const double* beta_dn_host,
                        const double* beta_0_host,
                        int32_t m_begin, int32_t m_end, int include_mean, int32_t flags);

SPUQ_API int spuq_sg_plan_destroy(spuq_sg_plan* plan);

/* Which path the plan runs (SPUQ_PATH_CSR / _ELL / _STENCIL) and a few numbers for reporting. */
SPUQ_API int spuq_sg_plan_info(const spuq_sg_plan* plan, int32_t* path_out, int64_t* n_terms_out,
                      int64_t* nnz_out, int64_t* algorithmic_bytes_out,
                      int64_t* algorithmic_flops_out);

/* Tunables of the stencil path (development sweeps, tools/sweep.py); 0 keeps the default.
 *   variant 0  automatic: the TMEM-accumulator kernel (4) whenever the grid width and the leading
 *              dimensions are even and the slabs 16-byte aligned, else the simple stencil kernel
 *   variant 1  k_apply_stencil_simple            (one thread per pair of rows)
 *   variant 2  k_apply_tiled  (cols_per_block = C, rows_per_thread = RY)   registers only
 *   variant 3  k_apply_tma    (same tiling, operands through TMA rings)
 *   variant 4  k_apply_tmem   (cols_per_block = K register sets of block matrices, 1..3;
 *              rows_per_thread = consumer groups per CTA, 1 or 2, or 3 = one fused
 *              group of eight warps on 64 x 16 tiles; accumulators in tensor memory) */
SPUQ_API int spuq_sg_plan_tune(spuq_sg_plan* plan, int32_t variant, int32_t cols_per_block,
                      int32_t rows_per_thread, int32_t ctas_per_sm);

/* Y = A W.  W_dev: n_cols x L (ldw >= n_cols), Y_dev: n_rows x L (ldy >= n_rows); must not alias.
 * One kernel launch (grid sized to a multiple of the SM count). */
SPUQ_API int spuq_sg_apply(spuq_sg_plan* plan, const double* W_dev, int64_t ldw,
                  double* Y_dev, int64_t ldy, void* stream);

/* The same product restricted to a part of the OUTPUT grid rows, for callers that overlap a halo exchange
 * with the bulk of the work (row-sharded slabs, spuq_b200.sharding.RowShardedApply): the first and the
 * last grid row of the local slab are halo rows filled from the neighbouring ranks.
 *   part 1  every output tile row whose stencil window reads neither halo row (may start before the
 *           exchange has finished)
 *   part 2  the remaining tile rows (after the exchange)
 *   part 0  all rows (= spuq_sg_apply)
 * Parts 1 and 2 together write every output row exactly once.  On kernel paths without a tile-row range
 * part 1 does nothing and part 2 computes everything. */
SPUQ_API int spuq_sg_apply_part(spuq_sg_plan* plan, const double* W_dev, int64_t ldw,
                       double* Y_dev, int64_t ldy, int32_t part, void* stream);

/* Number of kernels launched by the last spuq_sg_* call on this plan (for bench.py's
 * gpu_launches) and the name of the apply kernel. */
SPUQ_API int spuq_sg_plan_last_launches(const spuq_sg_plan* plan);
SPUQ_API const char* spuq_sg_plan_kernel_name(const spuq_sg_plan* plan);

/* ---------------------------------------------------------------------------------------------
 * K2 / K3: slab vector operations with device-resident scalars (no host round trip).
 * Replace MultiVector.__inner__/__iadd__/__isub__/__imul__,
 * application/egsz/multi_vector.py:126-152, and the Vector arithmetic of linalg/vector.py:103-133
 * as used by application/egsz/pcg.py:26-51.  All slabs contiguous, n = N*L elements.
 * ------------------------------------------------------------------------------------------- */

/* Scratch space (bytes) the reductions below need; caller allocates once per stream. */
SPUQ_API int64_t spuq_sg_reduce_scratch_bytes(void);

/* out_dev[0] = sum_i x[i]*y[i]; deterministic (fixed reduction tree, independent of timing). */
SPUQ_API int spuq_sg_dot(int64_t n, const double* x_dev, const double* y_dev, double* out_dev,
                void* scratch_dev, void* stream);
/* out_dev[0] = <x,y>, out_dev[1] = <x,x> in one pass (pcg.py:32 logs both). */
SPUQ_API int spuq_sg_dot2(int64_t n, const double* x_dev, const double* y_dev, double* out_dev,
                 void* scratch_dev, void* stream);
/* y += a*x  with a = alpha_host * (num_dev ? num_dev[0] : 1) / (den_dev ? den_dev[0] : 1) */
SPUQ_API int spuq_sg_axpy(int64_t n, double alpha_host, const double* num_dev, const double* den_dev,
                 const double* x_dev, double* y_dev, void* stream);
/* w += a*v ; rho -= a*z   with a = num_dev[0]/den_dev[0]        (pcg.py:47-48) */
SPUQ_API int spuq_sg_axpy2(int64_t n, const double* num_dev, const double* den_dev,
                  const double* v_dev, double* w_dev, const double* z_dev, double* rho_dev,
                  void* stream);
/* v = s + b*v  with b = num_dev[0]/den_dev[0]                    (pcg.py:51) */
SPUQ_API int spuq_sg_xpay(int64_t n, const double* num_dev, const double* den_dev,
                 const double* s_dev, double* v_dev, void* stream);
/* x *= alpha_host */
SPUQ_API int spuq_sg_scal(int64_t n, double alpha_host, double* x_dev, void* stream);
/* out[:, mu] = -b0[mu]*W[:, mu] + bu[mu]*W[:, up[mu]] + bd[mu]*W[:, dn[mu]]  for all L columns (absent
 * neighbour: index < 0; all five tables on the device, rows of the plan's neighbour / beta tables for
 * one term m).  The neighbour combination of residual_estimator2.py:111-131, same operation order. */
SPUQ_API int spuq_sg_neighbour_combine(int64_t n, int32_t L, const double* W_dev, int64_t ldw,
                              const int32_t* up_dev, const int32_t* dn_dev, const double* bu_dev,
                              const double* bd_dev, const double* b0_dev, double* out_dev, int64_t ldo,
                              void* stream);
/* out[s*ldo + i] = sum_mu coef_dev[s*L + mu] * W[i + ldw*mu]   for S coefficient vectors at once:
 * the value of the expansion at S parameter samples, coef[s][mu] = prod_m P_{mu_m}(y^s_m).
 * Replaces the sum over Lambda in compute_parametric_sample_solution,
 * application/egsz/sampling.py:59-75 (one pass over the slab per 8 samples). */
SPUQ_API int spuq_sg_combine(int64_t n, int32_t L, int32_t S, const double* W_dev, int64_t ldw,
                    const double* coef_dev, double* out_dev, int64_t ldo, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K4: preconditioners  s[mu] = P rho[mu]  applied to every column.
 * Replace PreconditioningOperator.apply, application/egsz/multi_operator2.py:151-165, and the
 * other preconditioners exercised by application/egsz/tests/test_pcg.py:12-44.
 * ------------------------------------------------------------------------------------------- */

/* P = scale * I   (MultiplicationOperator, linalg/operator.py:465-480) */
SPUQ_API int spuq_sg_precond_identity(spuq_sg_precond** out, double scale);
/* P = explicit sparse matrix applied to every column: wraps an M == 0 plan (not owned).
 * Covers DiagonalMatrixOperator(...).inverse() and MatrixOperator(inverse) of test_pcg.py. */
SPUQ_API int spuq_sg_precond_from_plan(spuq_sg_precond** out, spuq_sg_plan* plan);
/* P = A_0^{-1} for the mean block A_0 = 5-point stencil with CONSTANT coefficients on an
 * nx x ny grid (rows i = iy*nx + ix): diag `d`, x-coupling `ox`, y-coupling `oy` (so that
 * A_0 = tridiag_x(ox, d/2, ox) (+) tridiag_y(oy, d/2, oy)).  Exact to rounding: sine-transform
 * diagonalisation on strips of <= 32 grid rows held in shared memory plus an exact Schur-complement solve
 * on the separator rows (csrc/precond_strip.cuh).  Replaces the per-column sparse direct
 * solve of ScipySolveOperator.apply / FEniCSSolveOperator.apply (linalg/scipy_operator.py:45-49). */
SPUQ_API int spuq_sg_precond_mean_fd(spuq_sg_precond** out, int device, int32_t nx, int32_t ny,
                            double d, double ox, double oy);
/* P = A_0^{-1} for a GENERAL sparse mean block (variable coefficients, unstructured meshes): the caller
 * factorises once, Pr A_0 Pc = L U (scipy.sparse.linalg.splu, the solver behind the reference's
 * ScipySolveOperator.apply, linalg/scipy_operator.py:45-49), and hands over the factors as host CSR
 * arrays: the STRICT lower triangle of the unit-lower L, the STRICT upper triangle and the diagonal of U,
 * and the permutations (Pr b)[perm_r[i]] = b[i], (Pc w)[i] = w[perm_c[i]].  The device runs both
 * triangular solves for all slab columns in one launch (level-scheduled, one CTA per column group). */
SPUQ_API int spuq_sg_precond_sparse_lu(spuq_sg_precond** out, int device, int64_t n,
                              const int32_t* l_rowptr_host, const int32_t* l_colidx_host, const double* l_vals_host,
                              const int32_t* u_rowptr_host, const int32_t* u_colidx_host, const double* u_vals_host,
                              const double* u_diag_host, const int32_t* perm_r_host, const int32_t* perm_c_host);
/* The mean_fd solve in two halves, for callers that complete the y direction themselves (grid rows
 * sharded over GPUs: spuq_b200.sharding, partition method for the tridiagonal systems along y).
 *   forward: T = (tridiagonal solves along the handle's ny rows)(sine transform in x of R), left in the
 *            transformed domain; position c of a grid row holds mode order[c] (..._mode_order);
 *            work: spuq_sg_precond_work_bytes; R and T must not alias
 *   inverse: S = sine transform back of T; T and S must not alias
 *   spuq_sg_spike_correct: T[col][y][c] -= zl[col][c]*sl[y][c] + zr[col][c]*sr[y][c]  (L x ny x nx) */
SPUQ_API int spuq_sg_precond_mean_fd_forward(spuq_sg_precond* p, int64_t N, int32_t L, const double* R_dev,
                                    double* T_dev, void* work, void* stream);
SPUQ_API int spuq_sg_precond_mean_fd_inverse(spuq_sg_precond* p, int64_t N, int32_t L, const double* T_dev,
                                    double* S_dev, void* stream);
SPUQ_API int spuq_sg_precond_mean_fd_mode_order(const spuq_sg_precond* p, int32_t* order_host);
SPUQ_API int spuq_sg_spike_correct(int32_t L, int32_t ny, int32_t nx, double* T_dev, const double* zl_dev,
                          const double* zr_dev, const double* sl_dev, const double* sr_dev, void* stream);
/* The strip solver for a grid whose ROWS are sharded over ranks (spuq_b200.sharding, one rank per GPU).
 * `bounds_host[world + 1]` are the global row boundaries; the handle belongs to the band of rank `rank`.
 * The last row of every band but the last is a separator row, so the bands couple only through the
 * separator system, which is solved redundantly on every rank from the gathered right-hand sides:
 *     strip_forward   local strips, first / last rows of their solutions -> Gf, Gl [L][gf_ld][nx]
 *     (caller)        the first-strip rows of Gf travel to the previous rank (slot gf_ld - 1 there)
 *     strip_assemble  right-hand sides of the local separators -> rows first_sep .. of Z [q_glob][L][nx]
 *     (caller)        all-gather of the separator rows
 *     strip_schur     separator system of the whole grid, in place on Z (Z2: scratch of the same size)
 *     strip_final     local strips with the separator values on their right-hand side -> S
 * R and S point at the band's first row of column 0; ld = column stride in doubles (0: n_loc * nx, the band
 * stored on its own; larger: the band inside a halo-padded slab, no copy needed).  strip_final with
 * dot_out_dev != NULL also leaves <R, S> over the band in dot_out_dev[0] (same kernel, no extra slab pass).
 * All kernels honour spuq_sg_set_gate. */
SPUQ_API int spuq_sg_precond_mean_fd_rows(spuq_sg_precond** out, int device, int32_t nx, double d, double ox, double oy,
                                 int32_t world, int32_t rank, const int32_t* bounds_host);
SPUQ_API int spuq_sg_precond_strip_info(const spuq_sg_precond* p, int32_t* n_strips_loc, int32_t* q_loc, int32_t* q_glob,
                               int32_t* first_sep, int32_t* gf_ld);
SPUQ_API int spuq_sg_precond_strip_forward(spuq_sg_precond* p, int32_t L, const double* R_dev, int64_t ld, double* Gf_dev,
                                  double* Gl_dev, void* stream);
SPUQ_API int spuq_sg_precond_strip_assemble(spuq_sg_precond* p, int32_t L, const double* R_dev, int64_t ld, const double* Gf_dev,
                                   const double* Gl_dev, double* Z_dev, void* stream);
SPUQ_API int spuq_sg_precond_strip_schur(spuq_sg_precond* p, int32_t L, double* Z_dev, double* Z2_dev, void* stream);
SPUQ_API int spuq_sg_precond_strip_final(spuq_sg_precond* p, int32_t L, const double* R_dev, const double* U_dev, double* S_dev,
                                int64_t ld, double* dot_out_dev, void* stream);
SPUQ_API int spuq_sg_precond_destroy(spuq_sg_precond* p);
/* workspace (bytes) precond_apply needs for an N x L slab */
SPUQ_API int64_t spuq_sg_precond_work_bytes(const spuq_sg_precond* p, int64_t N, int32_t L);
/* kernels the last spuq_sg_precond_apply on this handle launched (bench.py's gpu_launches, the launch-count tests) */
SPUQ_API int spuq_sg_precond_last_launches(const spuq_sg_precond* p);
/* S = P R on N x L contiguous slabs */
SPUQ_API int spuq_sg_precond_apply(spuq_sg_precond* p, int64_t N, int32_t L, const double* R_dev,
                          double* S_dev, void* work_dev, void* stream);

/* ---------------------------------------------------------------------------------------------
 * PCG, application/egsz/pcg.py:15-53, all state on the device.
 * ------------------------------------------------------------------------------------------- */

/* Device words of PCG state, laid out in one small caller-owned buffer of
 * SPUQ_PCG_STATE_DOUBLES doubles: [0] zeta, [1] alpha, [2] zeta_new, [3] <rho,rho>,
 * [4] status (as double), [5] numit (as double), [6] eps^2. */
#define SPUQ_PCG_STATE_DOUBLES 16

/* One-thread kernels that evaluate the reference's branch conditions on the device.
 * pass = 1-based loop index i of pcg.py:31. */
SPUQ_API int spuq_sg_pcg_check_zeta(double* state_dev, int32_t pass, void* stream);   /* pcg.py:33-38 */
SPUQ_API int spuq_sg_pcg_check_alpha(double* state_dev, void* stream);                /* pcg.py:41-45 */
SPUQ_API int spuq_sg_pcg_rotate(double* state_dev, void* stream);                     /* zeta <- zeta_new */
/* check_zeta that also records zeta of the pass in hist_dev[pass - 1] (device, may be NULL) */
SPUQ_API int spuq_sg_pcg_check_zeta_hist(double* state_dev, int32_t pass, double* hist_dev, void* stream);
/* state[dst] = state[src] while the solve is running: takes an all-reduced scalar over from a staging slot
 * (slots 8..15 of the state are free for the caller) */
SPUQ_API int spuq_sg_pcg_accept(double* state_dev, int32_t src, int32_t dst, void* stream);

/* Peer-memory halo exchange (row-sharded slabs on the GPUs of one NVSwitch box): a rank copies its boundary
 * grid rows straight into the halo rows of its neighbours' slabs, which it has mapped through CUDA IPC,
 * with the copy engines (no kernel, no staging buffer), and signals with a sequence number written after
 * the copy on the same stream; the neighbour's stream waits for that number before its apply kernel.
 *   copy2d_async     cudaMemcpy2DAsync with cudaMemcpyDefault (either side may be peer memory)
 *   flag_wait_geq    the stream waits until *flag_dev >= value (one polling thread, ahead of the kernel)
 *   enable_peer_access  cudaDeviceEnablePeerAccess, idempotent */
/* The exchange itself: halo_push first publishes "ready to receive exchange `seq`" in the neighbours' flag words
 * (everything enqueued before it on this rank's stream -- the kernel that read the previous halo rows, any
 * update that swept over the slab since -- is complete), waits for the same word from them, stores this rank's
 * boundary rows into their slabs (peer pointers, NULL = no neighbour on that side) and publishes `seq` as
 * "written"; halo_wait holds the stream until both neighbours have published `seq` here.
 * flags: 4 x int64 per rank, [0] top halo written, [1] bottom halo written, [2] rank above ready to receive,
 * [3] rank below ready to receive; ticket_dev: one zeroed uint32. */
SPUQ_API int spuq_sg_halo_push(const double* W_dev, int32_t L, int32_t nyl, int32_t nx, double* up_slab, int32_t up_nyl,
                      double* dn_slab, int32_t dn_nyl, const int64_t* my_flags, int64_t* up_flags, int64_t* dn_flags,
                      int64_t seq, void* ticket_dev, void* stream);
SPUQ_API int spuq_sg_halo_wait(const int64_t* my_flags, int has_up, int has_dn, int64_t seq, void* stream);
SPUQ_API int spuq_sg_copy2d_async(void* dst, int64_t dpitch, const void* src, int64_t spitch, int64_t width_bytes,
                         int64_t height, void* stream);
SPUQ_API int spuq_sg_flag_wait_geq(const int64_t* flag_dev, int64_t value, void* stream);
SPUQ_API int spuq_sg_enable_peer_access(int device, int peer_device);
/* cudaIpcOpenMemHandle / cudaIpcCloseMemHandle with `device` current: maps an allocation another process
 * exported (64-byte cudaIpcMemHandle_t) so that kernels of `device` can access it.  ipc_export: the handle of
 * the allocation that contains ptr_dev (cudaMalloc'ed, e.g. a block of torch's caching allocator) and the
 * byte offset of ptr_dev inside it. */
SPUQ_API int spuq_sg_ipc_export(const void* ptr_dev, void* handle64_out_host, int64_t* offset_out_host);
SPUQ_API int spuq_sg_ipc_open(int device, const void* handle64_host, void** ptr_out);
SPUQ_API int spuq_sg_ipc_close(int device, void* ptr);

/* Conditional-execution scope for host-enqueued solver loops: while gate_dev is set for the calling thread,
 * every kernel launched through spuq_sg_apply / _apply_part / _dot* / _axpy* / _xpay / _precond_apply and the
 * strip phase calls is a no-op once *gate_dev != SPUQ_PCG_RUNNING (gate_dev = &state[4]).  NULL ends the scope. */
SPUQ_API int spuq_sg_set_gate(const double* gate_dev);

/* Whole solve on one GPU.  F_dev: right-hand side, W_dev: initial guess in, solution out (both
 * N x L contiguous).  work_dev: caller-owned, at least spuq_sg_pcg_work_bytes() bytes.
 * Returns (through host pointers) zeta, the pass index `numit` and the status; semantic of the
 * return triple as pcg.py:38.  For the statuses SINGULAR and NOT_PD zeta_out receives alpha = <Av, v>, the
 * value the reference prints in those messages (pcg.py:42-45).  `poll_every` passes are enqueued between two status read-backs
 * (no host round trip per iteration); zeta_hist_host (optional, maxiter doubles) receives zeta of
 * every executed pass for trajectory comparisons. */
SPUQ_API int64_t spuq_sg_pcg_work_bytes(const spuq_sg_plan* A, const spuq_sg_precond* P);
SPUQ_API int spuq_sg_pcg(spuq_sg_plan* A, spuq_sg_precond* P, const double* F_dev, double* W_dev,
                double eps, int32_t maxiter, int32_t poll_every, void* work_dev,
                double* zeta_out_host, int32_t* numit_out_host, int32_t* status_out_host,
                double* zeta_hist_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPUQ_SG_H */
