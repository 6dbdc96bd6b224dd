/* sfxm_b200.h - C ABI of the B200-native SplitFXM hot path.
 *
 * The reference (gpavanb1/SplitFXM v0.5.2) has no FFI for this path: its only native boundary is
 * Python -> Cython `def` wrappers (splitfxm/flux.pyx:42,302; splitfxm/derivatives.pyx:196-281) called from
 * splitfxm/equations/fv_transport.py:56-57 and fd_transport.py:46-50, driven by the pure-Python loops in
 * splitfxm/system.py:58-65 and splitfxm/simulation.py:150-202, 270-394.  The entry points below are what a
 * ctypes binding placed at those call sites would bind (INTEGRATION.md shows the stub); each one cites the
 * reference code it replaces.
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - the caller owns all memory; calls are asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - return value 0 = OK, non-zero = error (sfxm_last_error() gives the text; Python raises SFXM);
 *   - no global state except the last-error string.
 *
 * Data layout in HBM ("member-fastest"): an ensemble of `batch` independent members shares one mesh.
 *   flat vectors   y[ flat(p, j) * batch + m ]       p = interior cell, j = variable, m = member
 *                  flat(p, j) = var_base[j] + p * var_stride[j] + var_off[j]
 *                  (unsplit: base 0, stride nv, off j  == the reference's cell-major vector,
 *                   simulation.py:143; split: segment-major, simulation.py:137-140, 192-200)
 *   positions      x_int[nx]  (interior only; ghost positions are recomputed, bc.py:56-151)
 *   parameters     params[ k * param_stride + m ]   (param_stride = 0: one set shared by all members)
 *   Jacobian       J[ (((i * W + d) * nv + r) * nv + c) * batch + m ],  W = nbl + nbr + 1,
 *                  = d r_(i,r) / d U_((i - nbl + d) mod nx, c)  (block band with cyclic wrap slots)
 * With batch = 1 these are exactly the reference's flat vector and a block-banded matrix.
 */
#ifndef SFXM_B200_H
#define SFXM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFXM_MAX_NV 4
#define SFXM_MAX_NB 2
#define SFXM_MAX_PARAMS 4
#define SFXM_RK_MAX_STAGES 8

/* models (models/advection_diffusion.py:28-31, tutorials/sod_shock/euler1d.py:27-50,
 * tutorials/conservative_sod_shock/euler1d.py:28-63) */
enum { SFXM_MODEL_AD = 0, SFXM_MODEL_EULER_PRIM = 1, SFXM_MODEL_EULER_CONS = 2,
       /* custom Model._equation: device functors compiled from user CUDA source into a plugin library that
        * registers itself in slot (model - SFXM_MODEL_USER0), see sfxm_register_user_model */
       SFXM_MODEL_USER0 = 16 };
#define SFXM_MAX_USER_MODELS 16
/* schemes (schemes.py:4-14) */
enum { SFXM_FV_LAX_FRIEDRICHS = 0, SFXM_FV_CENTRAL = 1, SFXM_FV_UPWIND = 2, SFXM_FD_CENTRAL = 3,
       SFXM_FD_RIGHT_BIAS = 4 };
/* boundary types (bc.py:47-153); KEEP = variable/side has no entry in `bcs`, ghost keeps its value */
enum { SFXM_BC_PERIODIC = 0, SFXM_BC_OUTFLOW = 1, SFXM_BC_NEUMANN = 2, SFXM_BC_DIRICHLET = 3, SFXM_BC_KEEP = 4,
       /* slab decomposition: the ghost values come from the neighbouring rank (halo_left / halo_right below) */
       SFXM_BC_HALO = 5 };
/* which rule wrote the shared ghost positions last (bc.py:56,64 periodic; :73,87,114,127,141,151 mirror) */
enum { SFXM_GX_PERIODIC = 0, SFXM_GX_MIRROR = 1, SFXM_GX_KEEP = 2 };

typedef struct sfxm_problem {
  int32_t model, scheme;
  int32_t nv, nbl, nbr;
  int32_t n_params;
  int64_t nx;    /* interior cells */
  int64_t batch; /* ensemble members sharing the mesh */
  double xmin, xmax;
  /* per variable, per side */
  int32_t bc_ltype[SFXM_MAX_NV], bc_rtype[SFXM_MAX_NV];
  double bc_lval[SFXM_MAX_NV][SFXM_MAX_NB], bc_rval[SFXM_MAX_NV][SFXM_MAX_NB]; /* dirichlet/neumann value; KEEP: per ghost */
  int32_t gx_left, gx_right;                               /* SFXM_GX_* */
  double gx_keep_left[SFXM_MAX_NB], gx_keep_right[SFXM_MAX_NB];
  /* flat-vector layout (split permutation) */
  int64_t var_base[SFXM_MAX_NV];
  int32_t var_stride[SFXM_MAX_NV], var_off[SFXM_MAX_NV], var_seg[SFXM_MAX_NV];
  /* device buffers [(g * nv + j) * batch + m] with the neighbouring slab's boundary cells, ghost g = 0 nearest
   * (used by variables whose BC type is SFXM_BC_HALO; NULL otherwise) */
  const double *halo_left, *halo_right;
  /* fused halo exchange: when non-NULL, the explicit-Euler update (y_out of sfxm_residual) also stores this slab's
   * first nbr / last nbl updated cells straight into the LEFT neighbour's halo_right / RIGHT neighbour's halo_left
   * buffer - peer-mapped device memory of another GPU, written over NVLink by the same kernel (no pack, no NCCL
   * call); sfxm_halo_signal / sfxm_halo_wait order the steps. */
  double *peer_left_dst, *peer_right_dst;
} sfxm_problem_t;

int sfxm_abi_version(void);
const char *sfxm_last_error(void);

/* r = f(y).  Replaces Simulation.get_residuals_from_list (simulation.py:150-202): initialize_from_list
 * (:107-148) + apply_BC for every variable (bc.py:6-159) + System.residuals (system.py:28-70) +
 * FV/FDTransportEquation.residuals (fv_transport.py:27-59, fd_transport.py:26-51) + flux.pyx / derivatives.pyx
 * + the split-major re-ordering (:192-200), fused into one launch.
 * If y_out != NULL also writes y_out = y + dt * f(y) (the explicit Euler update, simulation.py:455-463). */
int sfxm_residual(const sfxm_problem_t *prob, const double *x_int, const double *params, int64_t param_stride,
                  const double *geom, const double *y, double *r, double *y_out, double dt, void *stream);

/* sfxm_residual for a PIECE of one mesh: batches below 32 of the coupled models (Euler) evaluate the residual with one warp
 * per 32 - (nbl + nbr) consecutive cells (fluxes once per cell, neighbours by shuffle); this entry point launches warps
 * [warp_lo, warp_hi) of the sfxm_residual_warp_count() warps only.  A slab-decomposed mesh sends its interior warps out
 * before the neighbours' halos have arrived and the two end warps - the only ones that read halo values and feed the
 * neighbours' - afterwards (splitfxm_b200/distributed.py); warp_lo = -1 launches exactly those two end warps together
 * (warp_hi is then ignored beyond being positive).  Other models / batch sizes: error. */
int64_t sfxm_residual_warp_count(const sfxm_problem_t *prob);
int sfxm_residual_warps(const sfxm_problem_t *prob, const double *x_int, const double *params, int64_t param_stride,
                        const double *y, double *r, double *y_out, double dt, int64_t warp_lo, int64_t warp_hi, void *stream);

/* Per-cell geometry table geom[SFXM_GEOM_ROWS * nx] (the divisors of the scheme, their reciprocals, 1/dx): it
 * depends on the mesh and the ghost-position rule only, so an ensemble computes it once per mesh instead of once
 * per member.  One record of SFXM_GEOM_ROWS doubles per cell (geom[i * SFXM_GEOM_ROWS + k]); the buffer must be 16-byte
 * aligned.  Optional everywhere: pass geom = NULL and each thread derives it from x_int. */
#define SFXM_GEOM_ROWS 10
int sfxm_geometry(const sfxm_problem_t *prob, const double *x_int, double *geom, void *stream);

/* Materialise the Domain view: x_all[nbl+nx+nbr], U_all[(nbl+nx+nbr) * nv * batch] (cell-major, unsplit,
 * member-fastest) with ghosts filled.  Replaces initialize_from_list + apply_BC (simulation.py:171-175). */
int sfxm_fill_domain(const sfxm_problem_t *prob, const double *x_int, const double *y, double *x_all, double *U_all,
                     void *stream);

/* Block-banded finite-difference Jacobian; all perturbations in one pass.  Replaces Simulation.jacobian
 * (simulation.py:270-394) incl. get_periodic_bcs / extend_band (bc.py:162-240).  Split mode writes
 * same-segment entries only (simulation.py:373-392); the others are left untouched (keep J zero-filled).  The same holds
 * for the off-diagonal block entries of SFXM_MODEL_AD (one scalar equation per variable): they are exact zeros in the
 * reference too and are never written - J must be zero-filled once by its owner.
 * r0 (optional) receives the unperturbed residual in flat order. */
int sfxm_jacobian(const sfxm_problem_t *prob, const double *x_int, const double *params, int64_t param_stride,
                  const double *geom, const double *y, double eps, double *J, double *r0, void *stream);

/* sfxm_jacobian restricted to the variable segments [seg_lo, seg_hi) of a split layout: only those segments' blocks
 * are evaluated and stored (the other entries of J are left untouched).  Split-Newton solves one segment at a time
 * (simulation.py:520-523); the reference re-assembles the whole matrix for every sub-solve and slices it.
 * r0 (optional): the rows of f(y) that belong to the requested segments are written; rows of other segments are
 * unspecified (some kernels write them, the ensemble kernel does not). */
int sfxm_jacobian_segments(const sfxm_problem_t *prob, const double *x_int, const double *params, int64_t param_stride,
                           const double *geom, const double *y, double eps, int32_t seg_lo, int32_t seg_hi, double *J,
                           double *r0, void *stream);

/* Compact copy of a split Jacobian for the trip to the host: J_packed[((i*W + d)*L + e)*batch + m] holds the L
 * same-segment (r, c) entries of every block in row-major (r, c) order - the only entries Simulation.jacobian stores
 * in split mode (simulation.py:373-392).  For SFXM_MODEL_AD (one scalar equation per variable) only the r == c entries
 * are kept: the others are exact zeros, which the reference's lil_matrix does not store either.  *n_live_host receives L;
 * with J == NULL the call only reports L. */
int sfxm_pack_jacobian(const sfxm_problem_t *prob, const double *J, double *J_packed, int32_t *n_live_host, void *stream);

/* Solve, for every member and every variable segment, (J_seg + shift_m I) s = -f  with the block-tridiagonal
 * (cyclic if wrap slots are non-zero) J from sfxm_jacobian; W must be 3.  `shift` is per member (NULL = 0):
 * the pseudo-transient term of Newton's method.  `active` (NULL = all) masks members.  `work` needs
 * sfxm_solve_workspace_bytes().  status[m] = 0 ok, 1 = singular pivot.  The linear algebra behind
 * splitnewton.newton / split_newton as called at simulation.py:515-523. */
int64_t sfxm_solve_workspace_bytes(const sfxm_problem_t *prob);
int sfxm_block_tridiag_solve(const sfxm_problem_t *prob, const double *J, const double *f, const double *shift,
                             const int32_t *active, int32_t seg_lo, int32_t seg_hi, double *s, double *work,
                             int32_t *status, void *stream);

/* Newton linear solve with the Jacobian assembled INSIDE the sweep (ensembles): (J_seg(y) + shift_m I) s = -f(y) for one
 * variable segment, where the thread that walks a member's cells evaluates the finite-difference block row of each cell
 * (sfxm_jacobian's arithmetic, simulation.py:320-354) in registers - J is never written to or read from HBM.  The
 * result is bit-identical to sfxm_jacobian_segments + sfxm_block_tridiag_solve.  r0 (optional) receives the segment's
 * rows of f(y).
 * work: sfxm_solve_workspace_bytes().  sfxm_newton_fused_supported() says whether kernels exist for the problem's
 * model / scheme (3-cell stencils of the built-in models); otherwise use the two calls above. */
int sfxm_newton_fused_supported(const sfxm_problem_t *prob);
int sfxm_newton_solve_fused(const sfxm_problem_t *prob, const double *x_int, const double *params, int64_t param_stride,
                            const double *geom, const double *y, double eps, const double *shift, const int32_t *active,
                            int32_t seg, double *s, double *r0, double *work, int32_t *status, void *stream);
/* The same solve with the Newton update folded into the back substitution: y += s for the active members (in place)
 * and norms2[m] = sum s^2, norms2[batch + m] = sum y_new^2 over the segment (the two sums of the convergence
 * criterion), accumulated by the one thread that owns member m - no separate update + reduction pass over (y, s). */
int sfxm_newton_step_fused(const sfxm_problem_t *prob, const double *x_int, const double *params, int64_t param_stride,
                           const double *geom, double *y, double eps, const double *shift, const int32_t *active,
                           int32_t seg, double *s, double *r0, double *work, int32_t *status, double *norms2,
                           void *stream);

/* Same contract as sfxm_block_tridiag_solve, by block parallel cyclic reduction: one thread per (cell, member)
 * and ceil(log2 nx) streaming passes instead of one thread per member - the solver for a single large mesh or
 * a small batch.  work: sfxm_pcr_workspace_bytes(). */
int64_t sfxm_pcr_workspace_bytes(const sfxm_problem_t *prob);
int sfxm_block_tridiag_solve_pcr(const sfxm_problem_t *prob, const double *J, const double *f, const double *shift,
                                 const int32_t *active, int32_t seg_lo, int32_t seg_hi, double *s, double *work,
                                 int32_t *status, void *stream);

/* The same solve for 4-cell stencils (FD RIGHT_BIAS, nb_left + nb_right = 3): cells are folded into two-cell
 * super-blocks, which makes the band block tridiagonal again, then PCR.  Periodic variables need an even nx. */
int64_t sfxm_band4_workspace_bytes(const sfxm_problem_t *prob);
int sfxm_block_band4_solve_pcr(const sfxm_problem_t *prob, const double *J, const double *f, const double *shift,
                               const int32_t *active, int32_t seg_lo, int32_t seg_hi, double *s, double *work,
                               int32_t *status, void *stream);

/* ---- Newton linear solve of ONE mesh decomposed into slabs, one slab per GPU (no reference equivalent: the
 * reference solves the whole mesh in one process, simulation.py:515-523).  J is this slab's block-tridiagonal
 * Jacobian whose wrap slots (cell 0, slot 0) / (cell nx-1, slot 2) hold the coupling blocks to the neighbouring
 * slabs (sfxm_jacobian fills them for SFXM_BC_HALO ghosts).
 * reduce: PCR with 1 + 2 s right-hand sides; leaves x_i = q_i - V_i xl_prev - W_i xf_next in `work` and copies the
 *         first/last rows out: iface[((e*s + r)*(1+2s) + c)*batch + m], e = 0 first cell, 1 last cell.
 * (all-gather iface, solve the 2 G s interface system on every rank: splitfxm_b200/spike.py)
 * fold:   s = q - V xl_prev - W xf_next with nbr[(e*s + c)*batch + m], e = 0 the left neighbour's last block
 *         unknown, e = 1 the right neighbour's first.  One segment per call; W must be 3. ---- */
int64_t sfxm_spike_workspace_bytes(const sfxm_problem_t *prob);
int sfxm_spike_reduce(const sfxm_problem_t *prob, const double *J, const double *f, const double *shift, int32_t seg,
                      double *work, double *iface, int32_t *status, void *stream);
int sfxm_spike_fold(const sfxm_problem_t *prob, int32_t seg, double *work, const double *nbr, const int32_t *active,
                    double *s, void *stream);
/* The interface system itself: aug[(m*n + r)*(n+1) + c] = [M | rhs] per member (n = 2 G s <= 76), z[m*n + r] the
 * solution; Gaussian elimination with partial pivoting, one CTA per member.  status[m] = 1 on a zero pivot. */
int sfxm_dense_solve(const double *aug, int32_t n, int64_t batch, double *z, int32_t *status, void *stream);

/* Per-member sums of squares over a flat-index range [lo, hi): out[m] = sum v[q*batch+m]^2, accumulated in a
 * fixed order (deterministic).  work: sfxm_sumsq_workspace_bytes(hi - lo, batch).  The norms of
 * splitnewton's convergence criterion. */
int64_t sfxm_sumsq_workspace_bytes(int64_t n, int64_t batch);
int sfxm_sumsq(const double *v, int64_t lo, int64_t hi, int64_t batch, double *out, double *work, void *stream);

/* The Newton update and both norms of the convergence criterion in ONE pass over (y, s): if update != 0,
 * y[q, m] += alpha[m] * s[q, m] for q in [lo, hi) and active members (alpha NULL = 1, active NULL = all); then
 * out2[m] = sum s^2 and out2[batch + m] = sum y_new^2 over [lo, hi), accumulated in the order of sfxm_sumsq (same bits).
 * work: 2 * sfxm_sumsq_workspace_bytes(hi - lo, batch). */
int sfxm_update_norms(double *y, const double *s, const double *alpha, const int32_t *active, int32_t update, int64_t lo,
                      int64_t hi, int64_t batch, double *out2, double *work, void *stream);

/* Per-member bookkeeping of the Newton / Split-Newton iteration (the loop the reference delegates to the third-party
 * splitnewton, simulation.py:515-523; specification: oracle/newton.py).  All arrays are device arrays [batch].
 * begin: (reset != 0: first crit[m] = inf, it[m] = 0, dt[m] = dt0 - the start of a loop;) run[m] = parent[m] (NULL = all)
 *        && crit[m] >= 1 && it[m] < maxiter; shift[m] = -1 / dt[m] (shift NULL: no pseudo-transient term); *any_flag = 1
 *        if any member runs (zeroed first).
 * end:   for running members crit[m] = sqrt(SS) / (abs_tol + rel_tol * sqrt(YY)) with SS = sum_{q < nsum} ss[q*batch + m]
 *        (YY likewise; nsum = 1: one segment's sums; an enclosing Split-Newton level sums the kept sums of its segments),
 *        it[m] += 1, keep_ss[m] = SS / keep_yy[m] = YY if given, and, if dt != NULL,
 *        dt[m] = min(dt[m] * sqrt(r_old_sq[m]) / sqrt(r_new_sq[m]), dtmax) (dtmax when the new residual is zero). */
int sfxm_newton_begin(int64_t batch, const int32_t *parent, double *crit, int64_t *it, double maxiter, double *dt, double dt0,
                      int32_t reset, int32_t *run, double *shift, int32_t *any_flag, void *stream);
int sfxm_newton_end(int64_t batch, const int32_t *run, const double *ss, const double *yy, int32_t nsum, double abs_tol,
                    double rel_tol, double *crit, int64_t *it, double *dt, const double *r_old_sq, const double *r_new_sq,
                    double dtmax, double *keep_ss, double *keep_yy, void *stream);

/* The whole Newton / Split-Newton loop nest of Simulation.steady_state (simulation.py:474-526; the reference hands it to
 * the third-party splitnewton: splitnewton.newton / split_newton at :515-523; specification oracle/newton.py) for the
 * plain configuration - full steps, no pseudo-transient term, no Armijo back-tracking, no bounds - driven from inside the
 * library: the calls above in the order of splitfxm_b200/newton.py (begin; Jacobian of the segment + block solve +
 * update/norms, or the fused sweep; end), so that no interpreter sits between the launches.  All buffers are the
 * caller's.  Loop levels: l = k for the Newton loop of segment k, l = nseg + k0 for the Split-Newton loop over segments
 * k0.. ; it / crit / run hold one [batch] row per level (2 * nseg rows), any_host one pinned int32 per level.
 * *result_level receives the level whose `it` row is the iteration count of the run. */
enum { SFXM_NEWTON_THOMAS = 0, SFXM_NEWTON_PCR = 1, SFXM_NEWTON_BAND4 = 2, SFXM_NEWTON_FUSED_STEP = 3 };
typedef struct sfxm_newton_plan {
  const sfxm_problem_t *prob;
  const double *x_int, *params;
  int64_t param_stride;
  const double *geom;
  double eps, abs_tol, rel_tol;
  int32_t nseg, solver;                       /* SFXM_NEWTON_* */
  int64_t row_lo[SFXM_MAX_NV], row_hi[SFXM_MAX_NV]; /* flat rows of segment k */
  int32_t solve_launches[SFXM_MAX_NV];        /* kernel launches of one solve of segment k (for the launch count only) */
  double *y, *s, *r, *J;                      /* state (in place), step, residual, banded Jacobian (unfused path) */
  double *work_solve, *work_norms, *norms2;   /* solver workspace, 2 * sfxm_sumsq_workspace_bytes, [2 * batch] */
  int32_t *status;
  int64_t *it;                                /* [2 * nseg, batch] */
  double *crit;                               /* [2 * nseg, batch] */
  int32_t *run;                               /* [2 * nseg, batch] */
  double *seg_ss, *seg_yy;                    /* [nseg, batch] */
  int32_t *any_dev;                           /* one device word */
  volatile int32_t *any_host;                 /* [2 * nseg] pinned host words */
} sfxm_newton_plan_t;
int sfxm_newton_run(const sfxm_newton_plan_t *plan, double maxiter, int32_t *result_level, int64_t *launches, void *stream);

/* y[q, m] += alpha[m] * s[q, m] for q in [lo, hi) and members with active[m] != 0. */
int sfxm_axpy_masked(double *y, const double *s, const double *alpha, const int32_t *active, int64_t lo, int64_t hi,
                     int64_t batch, void *stream);

/* ---- explicit Runge-Kutta vector kernels for Simulation.evolve (simulation.py:396-472, which calls
 * scipy.integrate.solve_ivp at :468-469; arithmetic order of scipy/integrate/_ivp/rk.py + common.py).
 * K holds the stages: K_s = K + s * stage_stride, each a flat [n * batch] vector; `total` = n * batch. ---- */

/* out = y + (sum_{s<nstage} K_s coef[s]) * h                     (rk.py rk_step) */
int sfxm_rk_lincomb(const double *y, const double *K, int64_t stage_stride, const double *coef_host, int32_t nstage,
                    double h, double *out, int64_t total, void *stream);
/* out = y_old + h * sum_k (sum_s K_s P[s][k]),  P_host[nstage][4]  (RkDenseOutput._call_impl at x = 1) */
int sfxm_rk_dense_end(const double *y_old, const double *K, int64_t stage_stride, const double *P_host, int32_t nstage,
                      double h, double *out, int64_t total, void *stream);
/* out[m] = sum_q ((a - a2) / (atol + |b| rtol))^2   (a2 may be NULL; select_initial_step's d0, d1, d2) */
int64_t sfxm_scaled_sumsq_workspace_bytes(int64_t n, int64_t batch);
int sfxm_scaled_diff_sumsq(const double *a, const double *a2, const double *b, double atol, double rtol, int64_t n,
                           int64_t batch, double *out, double *work, void *stream);
/* out[m] = sum_q (((sum_s K_s E[s]) * h) / (atol + max(|y|, |y_new|) rtol))^2   (_estimate_error_norm) */
int sfxm_rk_error_sumsq(const double *K, int64_t stage_stride, const double *E_host, int32_t nstage, double h,
                        const double *y, const double *y_new, double atol, double rtol, int64_t n, int64_t batch,
                        double *out, double *work, void *stream);

/* RK45 with scipy's step-size controller (rk.py:124-187) on the device: ONE attempted step is a fixed launch sequence -
 * sfxm_rk45_begin, six x (sfxm_rk_lincomb_dev + sfxm_residual), sfxm_rk_error_sumsq_dev, sfxm_rk45_control,
 * sfxm_rk45_commit - with no host decision inside, so it is captured once in a CUDA graph and replayed.
 *   ctl[8]  doubles: t, h_abs, h (this attempt), t_new, t_bound, max_step, sqrt(n_global), -
 *   flag[16] int32:  done, step_rejected, need_begin, accept, final, failed, nfev, accepted, attempts
 * begin: clamps h_abs for a new step, fails on a step below 10 ulp(t), clips the step to t_bound.  control: error norm
 * sqrt(err_sumsq) / ctl[6], accept (factor min(10, 0.9 err^-0.2), <= 1 after a rejection) or reject (max(0.2, ...)).
 * commit (accepted steps only): the final step writes the dense-output end value to `result`; y <- y_new, K_0 <- K_last.
 * After `done` every kernel of the sequence is a no-op (the right-hand sides still run, on scratch). */
int sfxm_rk45_begin(double *ctl, int32_t *flag, void *stream);
int sfxm_rk45_control(double *ctl, int32_t *flag, const double *err_sumsq, void *stream);
int sfxm_rk_lincomb_dev(const double *y, const double *K, int64_t stage_stride, const double *coef_host, int32_t nstage,
                        const double *ctl, const int32_t *flag, double *out, int64_t total, void *stream);
int sfxm_rk_error_sumsq_dev(const double *K, int64_t stage_stride, const double *E_host, int32_t nstage, const double *ctl,
                            const double *y, const double *y_new, double atol, double rtol, int64_t n, int64_t batch,
                            double *out, double *work, void *stream);
int sfxm_rk45_commit(double *y, const double *y_new, double *K, int64_t stage_stride, const double *P_host, int32_t nstage,
                     const double *ctl, const int32_t *flag, double *result, int64_t total, void *stream);

/* ---- adaptive mesh refinement (one mesh, batch = 1): Refiner.refine + perform_changes (refine.py:90-282).
 * x[n], U[n * nv] cell-major interior.  mark: reductions (min/max of values and slopes per variable,
 * refine.py:136-153), slope / curvature / grid-ratio / prune flags (:159-227), the adjacent-deletion rule (:231-233)
 * and the exclusive scans that give every surviving or new cell its output slot; counts_dev[3] = {new size,
 * inserted, deleted}.  apply: writes the refined mesh (kept cells + midpoints, refine.py:260-282).
 * flags: copies the decision flags out (loc[n], keep[n], gridpt[n] bytes; comp[SFXM_MAX_NV] ints) for the
 * reference's stdout report (refine.py:284-312).  `work` is shared by the three calls. ---- */
int64_t sfxm_refine_workspace_bytes(int64_t n);
int sfxm_refine_mark(const double *x, const double *U, int64_t n, int32_t nv, double ratio, double slope, double curve,
                     double prune, double min_range, double min_grid, void *work, int64_t *counts_dev, void *stream);
int sfxm_refine_flags(const void *work, int64_t n, uint8_t *loc_out, uint8_t *keep_out, uint8_t *gridpt_out,
                      int32_t *comp_out, void *stream);
int sfxm_refine_apply(const double *x, const double *U, int64_t n, int32_t nv, const void *work, double *x_out,
                      double *U_out, void *stream);
/* The whole refinement pass of ONE mesh in three launches (stats + zero-fill, mark, single-pass scan + apply with
 * decoupled look-back): x_out / U_out need capacity 2n - 1 cells; counts_dev[0] = new size, [1] = inserted,
 * [2] = deleted.  work: sfxm_refine_fused_workspace_bytes(n).  sfxm_refine_fused_flags copies the decision flags out
 * (as sfxm_refine_flags) for the reference's report (refine.py:284-312). */
int64_t sfxm_refine_fused_workspace_bytes(int64_t n);
int sfxm_refine_fused(const double *x, const double *U, int64_t n, int32_t nv, double ratio, double slope, double curve,
                      double prune, double min_range, double min_grid, void *work, double *x_out, double *U_out,
                      int64_t *counts_dev, void *stream);
int sfxm_refine_fused_flags(const void *work, int64_t n, uint8_t *loc_out, uint8_t *keep_out, uint8_t *gridpt_out,
                            int32_t *comp_out, void *stream);
/* Explicit flags instead of sfxm_refine_mark - Refiner.perform_changes(d, loc, keep), refine.py:239-282:
 * keep_in[t] = 0 deletes cell t, loc_in[t] = 1 inserts the midpoint after cell t (device byte arrays [n]); fills the
 * workspace (scans, counts_dev as for sfxm_refine_mark) so that sfxm_refine_apply builds the new mesh. */
int sfxm_refine_set_flags(const uint8_t *loc_in, const uint8_t *keep_in, int64_t n, void *work, int64_t *counts_dev,
                          void *stream);

/* The same refinement on a WINDOW of a slab-decomposed mesh (AMR over several GPUs).  x[ne], U[ne * nv] hold
 * the global cells g0 .. g0 + ne - 1 of a mesh of ng cells; [own_lo, own_hi) (window indices) are this rank's cells, the
 * rest is overlap copied from the neighbours (6 cells per side, fewer at a domain end).  Phases, with the two global
 * exchanges between them done by the caller:
 *   stats    -> stats_dev[16] = vmin[4] | vmax[4] | smin[4] | smax[4] over the own cells   (all-reduce min / max)
 *   mark     -> flags; last_run_dev[0] = global index of the last deletion-run start so far, -1 if none (all-gather;
 *               carry = max over the ranks before this one)
 *   finalize -> the adjacent-deletion rule with `carry_run_start`, output slots, counts_dev[3] for the own cells
 *   apply    -> x_out / U_out: this rank's refined cells (midpoints after own cells included)
 *   flags    -> decision flags of the own cells.   `work`: sfxm_refine_workspace_bytes(ne). */
int sfxm_refine_window_stats(const double *x, const double *U, int64_t ne, int32_t nv, int64_t g0, int64_t ng,
                             int64_t own_lo, int64_t own_hi, void *work, double *stats_dev, void *stream);
int sfxm_refine_window_mark(const double *x, const double *U, int64_t ne, int32_t nv, int64_t g0, int64_t ng,
                            int64_t own_lo, int64_t own_hi, double ratio, double slope, double curve, double prune,
                            double min_range, double min_grid, const double *stats_dev, void *work,
                            int64_t *last_run_dev, void *stream);
int sfxm_refine_window_finalize(int64_t ne, int64_t g0, int64_t ng, int64_t own_lo, int64_t own_hi,
                                int64_t carry_run_start, void *work, int64_t *counts_dev, void *stream);
int sfxm_refine_window_apply(const double *x, const double *U, int64_t ne, int32_t nv, int64_t g0, int64_t ng,
                             int64_t own_lo, int64_t own_hi, const void *work, double *x_out, double *U_out, void *stream);
int sfxm_refine_window_flags(const void *work, int64_t ne, int64_t own_lo, int64_t own_hi, uint8_t *loc_out,
                             uint8_t *keep_out, uint8_t *gridpt_out, int32_t *comp_out, void *stream);

/* Register the launcher of a user-model plugin (built by splitfxm_b200/custom.py from the user's device functor:
 * the same residual/Jacobian kernel templates instantiated for that functor).  launch_fn has the signature
 * int (*)(const void *launch) and is called by sfxm_residual / sfxm_jacobian when prob->model is
 * SFXM_MODEL_USER0 + slot.  The reference's equivalent is a Model whose _equation holds Python callables
 * F, D, S, dFdU (model.py:8, fv_transport.py:18, fd_transport.py:18). */
int sfxm_register_user_model(int32_t slot, void *launch_fn);

/* Step ordering for the fused halo exchange.  signal: after everything queued so far on `stream`, publish `value`
 * (system scope) in the two neighbours' flag words (either may be NULL).  wait: a one-thread kernel that spins until
 * both of this rank's flag words (either may be NULL) are >= value; it gives up after ~2 s and sets *timeout_flag.
 * Flags are step counters in peer-mapped memory. */
int sfxm_halo_signal(int64_t *left_peer_flag, int64_t *right_peer_flag, int64_t value, void *stream);
int sfxm_halo_wait(const int64_t *from_left_flag, const int64_t *from_right_flag, int64_t value, int32_t *timeout_flag,
                   void *stream);

/* Same, with the step counter kept in device memory (`counter`, owned by this rank): signal increments it and
 * publishes the new value, wait blocks until both neighbours' flags reach it.  No per-step kernel argument, so a
 * (wait, update, signal) sequence is captured once in a CUDA graph and replayed. */
int sfxm_halo_signal_counter(int64_t *counter, int64_t *left_peer_flag, int64_t *right_peer_flag, void *stream);
int sfxm_halo_wait_counter(const int64_t *counter, const int64_t *from_left_flag, const int64_t *from_right_flag,
                           int32_t *timeout_flag, void *stream);

/* ---- kernel-level entry points on ONE stencil: the cdef workers behind the reference's Python wrappers fluxes /
 * diffusion_fluxes (flux.pyx:42-124, 302-327) and Dx / D2x / dx / d2x (derivatives.pyx:196-281).  The wrappers
 * evaluate the user's Python callables on the stencil cells (host) and pass the values; arrays are device pointers,
 * [slots][nv] row-major. ---- */
int sfxm_fv_fluxes(int32_t scheme, int32_t nv, const double *F, const double *u, double sigma, const double *eig,
                   double *Fw, double *Fe, void *stream);
int sfxm_fv_diffusion_fluxes(int32_t nv, const double *D, const double *x, double *Dw, double *De, void *stream);
int sfxm_fd_derivative(int32_t scheme, int32_t order, int32_t nv, const double *F, const double *x, double *out,
                       void *stream);
/* One cell's residual from the stencil VALUES - the arithmetic of FVTransportEquation.residuals
 * (equations/fv_transport.py:49-58: rhs = (S - (1/dx)(Fe - Fw)) + (1/dx)(DFe - DFw)) and FDTransportEquation.residuals
 * (equations/fd_transport.py:46-50: rhs = (S - Dx(F)) + D2x(D)) in ONE launch.  F, D, u: [slots][nv]; x: [slots];
 * S: [nv]; sigma / eig as for sfxm_fv_fluxes. */
int sfxm_fv_cell_residual(int32_t scheme, int32_t nv, const double *F, const double *D, const double *u, const double *x,
                          const double *S, double sigma, const double *eig, double *out, void *stream);
int sfxm_fd_cell_residual(int32_t scheme, int32_t nv, const double *F, const double *D, const double *x, const double *S,
                          double *out, void *stream);

/* Device self-test: counts (into *mismatch_dev, a device uint64 the caller zeroes) operands for which the
 * five-FMA division by a known reciprocal used in the stencil kernels differs from the IEEE quotient.
 * mode 0 = random operands, mode 1 = adversarial (divisor mantissa all ones/zeros, half-way numerators). */
int sfxm_selftest_div(int64_t n, uint64_t seed, int32_t mode, uint64_t *mismatch_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SFXM_B200_H */
