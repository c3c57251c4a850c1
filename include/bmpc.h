/*
 * bmpc.h -- C ABI of the B200-native batched LinMPC / linear-MHE step (libbmpc.so).
 *
 * Drop-in boundary for ONE hot path of JuliaControl/ModelPredictiveControl.jl v2.11.0:
 * the per-period `moveinput!` of `LinMPC` (reference src/controller/execute.jl:59-80 =
 * initpred! :247-277 + linconstraint! src/controller/transcription.jl:811-848 +
 * optim_objective! execute.jl:466-505 + getinput! :536-546), executed for a batch of N
 * independent controller instances in one launch of hand-written sm_100a CUDA.
 *
 * The reference has no FFI of its own (100 % Julia).  The seam this ABI plugs into is the
 * one `ExplicitMPC` already uses (src/controller/explicitmpc.jl:198-209): the generic
 * functions initpred!/linconstraint!/optim_objective!/getinput! called by moveinput!
 * (execute.jl:75-79), with the state estimate supplied from outside exactly like
 * `LinMPC(ManualEstimator(model))` + `setstate!` (src/estimator/manual.jl:60-64,150-154).
 * INTEGRATION.md shows the Julia `ccall` binding; `modelpredictivecontrol.jl_b200/` holds
 * the ctypes mirror used by this repo's tests and bench.
 *
 * Conventions
 *   - all reals are IEEE fp64; matrices are COLUMN-MAJOR exactly as the Julia fields;
 *   - batch arrays are instance-major: element k of instance i is  arr[i*len + k];
 *   - "0" suffix = deviation from the operating point, as in the reference
 *     (mpc.estim.x̂0, mpc.lastu0, con.U0min = umin - Uop, ...: construct.jl:356-409);
 *   - nYhat = ny*Hp, nU = nu*Hp, nDU = nu*Hc, n = nDU + neps (Z̃ = [ΔU; ϵ], slack LAST,
 *     construct.jl:1000-1004);
 *   - +-Inf in a bound array means "no constraint" (i_b mask, transcription.jl:692-700); the
 *     finiteness pattern is shared by all instances of a handle and frozen after the first
 *     step (construct.jl:548-551);
 *   - functions return BMPC_OK or a negative error code and never throw; bmpc_last_error()
 *     returns a thread-local message.  Per-instance solver outcome is in `status`.
 *   - there is NO CPU fallback: every entry point that computes requires a CUDA device.
 */
#ifndef BMPC_H
#define BMPC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BMPC_OK 0
#define BMPC_ERR_ARG (-1)         /* bad dimension / null pointer / inconsistent pattern        */
#define BMPC_ERR_CUDA (-2)        /* CUDA runtime failure (message in bmpc_last_error)           */
#define BMPC_ERR_STATE (-3)       /* call order, or +-Inf pattern changed after the first step   */
#define BMPC_ERR_UNSUPPORTED (-4) /* feature outside the hot-path scope          */

/* per-instance solver status, mirrors the reference policy in optim_objective!
 * (execute.jl:482-503, general.jl:45-61) */
#define BMPC_STATUS_OPTIMAL 0         /* solved to tolerance                                    */
#define BMPC_STATUS_ITERATION_LIMIT 1 /* not converged, iterate kept (cf. @warn branch :490-496) */
#define BMPC_STATUS_INFEASIBLE 2      /* infeasible / numerical failure: Ztilde = shifted previous
                                         solution Z̃s, exactly like :499-500                     */

typedef struct bmpc_handle bmpc_handle;

/* Dimensions of a batch of identically-structured LinMPC controllers
 * (mirrors the allocation in the LinMPC inner constructor, src/controller/linmpc.jl:50-111). */
typedef struct {
    int32_t N;            /* controller instances in the batch                                  */
    int32_t nu, ny, nd;   /* manipulated inputs, outputs, measured disturbances of the LinModel */
    int32_t nxhat;        /* augmented state size (augment_model, estimator/construct.jl:305-323) */
    int32_t Hp, Hc;       /* prediction horizon; number of move blocks (length of nb)           */
    int32_t neps;         /* 1 if Cwt is finite (slack variable present), else 0                */
    int32_t shared_model; /* 1: model-dependent constants are given ONCE and shared by all N    */
    int32_t max_iter;     /* interior-point iteration cap (0 -> 50); replaces the Ts time limit */
    int32_t device;       /* CUDA device ordinal                                                */
    int32_t team;         /* threads per instance: 0 = auto, else 8/16/32/64/128/256           */
    double tol;           /* relative KKT tolerance (0 -> 1e-11)                                */
} bmpc_dims;

/* Softness (ECR) vectors, shared by all instances; NULL member -> reference default
 * (c_u = c_du = 0 hard, c_y = c_xhat = 1 soft: construct.jl:909-913). Ignored when neps = 0. */
typedef struct {
    const double *C_umin, *C_umax;   /* nU   */
    const double *C_dumin, *C_dumax; /* nDU  */
    const double *C_ymin, *C_ymax;   /* nYhat */
    const double *c_xmin, *c_xmax;   /* nxhat */
} bmpc_softness;

/* Inputs / outputs of one control period for the whole batch (= moveinput! arguments,
 * execute.jl:59-70).  Pointers are HOST pointers unless device_ptrs = 1. */
typedef struct {
    const double *xhat0;  /* N x nxhat   current state estimate, deviation (mpc.estim.x̂0)        */
    double *lastu0;       /* N x nu      in: u0(k-1); out: u0(k)  (mpc.lastu0, getinput! :544)   */
    const double *ry;     /* N x ny      output setpoint, repeated over Hp when Rhat_y is NULL   */
    const double *Rhat_y; /* N x nYhat   or NULL                                                 */
    const double *Rhat_u; /* N x nU      or NULL (= Uop, execute.jl:66)                          */
    const double *d0;     /* N x nd      measured disturbance deviation d0(k), NULL if nd = 0    */
    const double *Dhat0;  /* N x nd*Hp   predicted deviations D̂0, NULL -> repeat(d0, Hp)         */
    double *Ztilde;       /* N x n       in: previous solution (warm start / fallback); out: Z̃   */
    double *u;            /* N x nu      out: u(k) = Z̃[1:nu] + lastu0 + uop                      */
    double *J;            /* N           out: objective 1/2 Z̃'H̃Z̃ + q̃'Z̃ + r, or NULL             */
    int32_t *status;      /* N           out: BMPC_STATUS_*                                      */
    int32_t *iters;       /* N           out: interior-point iterations (0 = unconstrained exit) */
    int32_t device_ptrs;  /* 1: every pointer above is a device pointer on dims.device           */
    int32_t sync;         /* 1: block until the results are complete                             */
    int32_t resident;     /* host pointers only.  0: lastu0 and Ztilde are uploaded and downloaded every
                           * call.  1: they are STATE OF THE HANDLE, as mpc.lastu0 and mpc.Z̃ are fields of the
                           * reference controller (linmpc.jl:3-49) that moveinput! neither takes nor returns:
                           * nothing is uploaded; lastu0, Ztilde, J and iters may be NULL and are download-only
                           * when given.  The state is the one left by the previous call (zeros after
                           * bmpc_create).                                                         */
    int32_t host_mapped;  /* host pointers only.  1: the caller's arrays are page-locked, device-accessible host memory
                           * (cudaHostAlloc / cudaHostRegister under unified addressing): the step kernel reads the
                           * inputs from and writes u, J, status, iters to them directly over PCIe -- zero-copy, no
                           * staging copies or copy launches (lastu0 / Ztilde still follow `resident`).            */
    const double *y0m;    /* N x nym  measured outputs, deviation (ym - yop[i_ym]).  Used instead of xhat0 (which must
                           * then be NULL) after bmpc_set_estimator: the step kernel runs the observer's correction
                           * before the controller and its prediction after it (one launch per control period).   */
    const double *Yhat_s; /* N x nYhat or NULL  stochastic output predictions Ŷs = Ks x̂s + Ps ŷs of an InternalModel estimator
                           * (init_stochpred construct.jl:1254-1267, predictstoch! execute.jl:321-327): F starts from them.     */
    double *kkt;          /* out, N x 3 or NULL  relative KKT residuals of the returned iterate: primal ||Gx+s-h|| / (1+||h||),
                           * dual ||Hx+q+G'lam|| / (1+||q||+terms), complementarity s'lam / ((1+||q||)(1+||h||)).  Zeros for the
                           * unconstrained exit.  Tells a tol-level solve (1e-11) from an "acceptable" one (<= 1e-8).           */
} bmpc_step_io;

/* Diagnostics of the last step (= getinfo, execute.jl:145-198); any pointer may be NULL.
 * Host pointers. */
typedef struct {
    double *Yhat0;      /* N x nYhat  Ŷ0 = Ẽ Z̃ + F   (predict!, transcription.jl:1136-1145)      */
    double *U0;         /* N x nU     U0 = P̃u Z̃ + Tu lastu0                                      */
    double *xhat0end;   /* N x nxhat  x̂0(k+Hp) = ẽx̂ Z̃ + fx̂                                       */
    double *F;          /* N x nYhat  (initpred! output)                                          */
    double *qtilde;     /* N x n      (initpred! output, reference coordinates)                   */
    double *r;          /* N                                                                      */
} bmpc_info;

const char *bmpc_last_error(void);
int bmpc_version(void);

/* nb: move-blocking vector (move_blocking, construct.jl:629-660), length Hc, sum = Hp. */
int bmpc_create(bmpc_handle **out, const bmpc_dims *dims, const int32_t *nb);
int bmpc_destroy(bmpc_handle *h);
/* Use an existing CUDA stream (cudaStream_t passed as void*); NULL -> the handle's own. */
int bmpc_set_stream(bmpc_handle *h, void *stream);

/* Route A -- give the augmented model (estim.Â,B̂u,Ĉ,B̂d,D̂d, f̂op-x̂op) and the weights; the
 * prediction matrices and the Hessian are built ON THE DEVICE (init_predmat
 * transcription.jl:115-194 + init_quadprog construct.jl:837-845).  This is also the
 * batched `setmodel!` (execute.jl:621-790).  M_diag nYhat, N_diag nDU, L_diag nU.
 * Arrays are N x len, or 1 x len when dims.shared_model = 1. */
int bmpc_set_model(bmpc_handle *h, const double *Ahat, const double *Buhat, const double *Chat,
                   const double *Bdhat, const double *Ddhat, const double *fop_minus_xop,
                   const double *M_diag, const double *N_diag, const double *L_diag, double Cwt);

/* Route B -- the host (Julia) already holds the matrices: mpc.Ẽ (without the slack column:
 * nYhat x nDU), mpc.K, .V, .B, .G, .J, mpc.H̃ (n x n, lower triangle authoritative,
 * construct.jl:842) and the terminal matrices con.ẽx̂ (nxhat x nDU), kx̂, vx̂, bx̂, gx̂, jx̂
 * (linmpc.jl:31-38, construct.jl:128-134).  G/J/gx/jx may be NULL when nd = 0; the
 * terminal set may be NULL when no x̂ bound is ever finite. */
int bmpc_set_predmat(bmpc_handle *h, const double *E, const double *K, const double *V,
                     const double *B, const double *G, const double *J, const double *Htilde,
                     const double *ex, const double *kx, const double *vx, const double *bx,
                     const double *gx, const double *jx);

/* Weights needed per step for q̃ and r (ControllerWeights, construct.jl:45-93).
 * M: nYhat (diag) or nYhat x nYhat when M_dense = 1; L_diag: nU or NULL (= 0). */
int bmpc_set_weights(bmpc_handle *h, const double *M, int32_t M_dense, const double *L_diag);
/* The same with a DENSE input-setpoint weight: L is nU x nU (column-major, lower triangle authoritative, Hermitian as
 * the reference stores it) when L_dense = 1, else the diagonal (nU).  A dense Ñ_Hc only enters H̃ (route B).  Route B
 * only: on route A the weights are arguments of bmpc_set_model (BMPC_ERR_STATE). */
int bmpc_set_weights_dense(bmpc_handle *h, const double *M, int32_t M_dense, const double *L, int32_t L_dense);

/* Operating points uop (nu), yop (ny) per instance (Uop/Yop = repeat, linmpc.jl:90).
 * NULL = zeros. */
int bmpc_set_oppoints(bmpc_handle *h, const double *uop, const double *yop);

/* Bounds in deviation form, N x len each (NULL = all infinite) -- the state that
 * setconstraint! leaves in mpc.con (construct.jl:152-161). */
int bmpc_set_constraints(bmpc_handle *h, const double *U0min, const double *U0max,
                         const double *DUmin, const double *DUmax, const double *Y0min,
                         const double *Y0max, const double *xhat0min, const double *xhat0max,
                         const bmpc_softness *soft);

/* MultipleShooting transcription (LinMPC(...; transcription = MultipleShooting())).  For a LinModel the
 * MultipleShooting QP -- decision vector Z = [ΔU; X̂0], equality constraints ES Z + FS = 0 (init_predmat
 * transcription.jl:217-240, init_defectmat :373-414, linconstrainteq! :913-928) -- has the same optimum as the condensed
 * SingleShooting QP: the defect equations are exactly init_predmat's state recursion.  The CUDA path therefore always
 * solves the condensed problem; bmpc_get_states returns the X̂0 block (N x nxhat Hp: x̂0(k+1) ... x̂0(k+Hp), deviation
 * form) implied by the ΔU of the last step, so that a host can assemble Z̃ = [ΔU; X̂0; ε] in the reference's layout.
 * Needs the augmented model of route A (bmpc_set_model).  After an infeasible period (status 2) ΔU is the shifted
 * previous solution and X̂0 is the trajectory THAT sequence produces (the reference keeps the shifted previous states). */
int bmpc_get_states(bmpc_handle *h, double *X0);

/* Custom linear inequality constraints (LinMPC kwargs Wy, Wu, Wd, Wr: validate_custom_lincon
 * construct.jl:666-695, relaxW :1138-1160, linconstraint_custom! execute.jl:337-366):
 *     Wmin <= Wy [ŷ(k); Ŷ] + Wu [U; u(k+Hp-1)] + Wd [d(k); D̂] + Wr [r̂y(k); R̂y] <= Wmax       (nw rows x (Hp + 1) steps)
 * Wy nw x ny, Wu nw x nu, Wd nw x nd, Wr nw x ny (column-major, NM copies, NULL = zero matrix); Chat = estim.Ĉ (ny x nxhat) and
 * Ddhat = estim.D̂d (ny x nd) give ŷ(k) = Ĉ x̂0 + D̂d d0 + yop of the first block; dop (nd) turns the deviation-form d0 / D̂0
 * of bmpc_step into the absolute values the constraint is written in.  The rows' matrix Ew is built on the device; Fw
 * is rebuilt every period inside the step kernel.  nw = 0 removes them.  Call before bmpc_set_constraints. */
int bmpc_set_custom(bmpc_handle *h, int32_t nw, const double *Wy, const double *Wu, const double *Wd, const double *Wr,
                    const double *Chat, const double *Ddhat, const double *dop);
/* Wmin / Wmax: N x nw (Hp + 1), ABSOLUTE units (NULL = -Inf / +Inf); C_wmin / C_wmax: nw (Hp + 1) softness, shared
 * (NULL = 1, the reference default).  Compiled into the row tables by the NEXT bmpc_set_constraints call. */
int bmpc_set_custom_bounds(bmpc_handle *h, const double *Wmin, const double *Wmax, const double *C_wmin,
                           const double *C_wmax);

/* One control period for all N instances (= moveinput!). */
int bmpc_step(bmpc_handle *h, const bmpc_step_io *io);

/* getinfo quantities of the last step. */
int bmpc_getinfo(bmpc_handle *h, const bmpc_info *info);

/* Fused observer -- SURVEY 8f-1.  SteadyKalmanFilter in direct form, as the reference's default LinMPC estimator:
 *   preparestate!  x̂0 <- x̂0 + K̂ (y0m - Ĉm x̂0 - D̂dm d0)              (correct_estimate_obsv!, kalman.jl:284-296)
 *   updatestate!   x̂0 <- Â x̂0 + B̂u u0 + B̂d d0 + (f̂op - x̂op)        (predict_estimate_obsv!, kalman.jl:298-309)
 * run inside the step kernel around moveinput!, with x̂0 owned by the handle (estim.x̂0).  Matrices column-major,
 * N x len or 1 x len when dims.shared_model = 1; K̂ is the steady-state gain (nxhat x nym), Ĉm/D̂dm the measured rows.
 * Bdhat/Ddmhat may be NULL when nd = 0, fop_minus_xop may be NULL (= 0). */
int bmpc_set_estimator(bmpc_handle *h, const double *Ahat, const double *Buhat, const double *Bdhat,
                       const double *Cmhat, const double *Ddmhat, const double *Khat,
                       const double *fop_minus_xop, int32_t nym);
/* Time-varying KalmanFilter (reference src/estimator/kalman.jl:311-525) as the fused observer: after bmpc_set_estimator
 * (whose Khat may then be NULL) give the covariances cov.P̂_0, Q̂, R̂ (column-major, N x len or 1 x len when
 * dims.shared_model = 1).  Every bmpc_step with io.y0m then runs, around the step kernel,
 *   before:  K̂(k) = P̂ Ĉm' (Ĉm P̂ Ĉm' + R̂)^-1,  P̂ <- Hermitian((I - K̂ Ĉm) P̂, :L)   (correct_estimate_kf!, :1235-1268)
 *   after:   P̂ <- Hermitian(Â P̂ Â' + Q̂, :L)                                        (predict_estimate_kf!, :1270-1290)
 * in one small kernel each (one CTA per model; the recursion does not depend on the data), while the state update
 * x̂ += K̂ v̂, x̂ <- Â x̂ + ... stays inside the step kernel.  bmpc_get_cov reads P̂ back (N or 1 x nxhat x nxhat). */
int bmpc_set_estimator_cov(bmpc_handle *h, const double *P0, const double *Qhat, const double *Rhat);
int bmpc_get_cov(bmpc_handle *h, double *Phat);
/* setstate! / read-back of the observer state: x̂0 (prediction for the next period) and the corrected estimate the
 * last step used; either output may be NULL. */
int bmpc_set_state(bmpc_handle *h, const double *xhat0);
int bmpc_get_state(bmpc_handle *h, double *xhat0, double *xhat0_corrected);

/* Multi-GPU collection of the moves ("NCCL all-gather only to collect ΔŨ"), fused into the step:
 * peer_bufs[p] is rank p's gather buffer [world x N x n] doubles, mapped into THIS process (CUDA IPC / symmetric
 * memory, peer access over NVLink).  The step kernel's epilogue then stores every instance's Z̃ straight into slot
 * (rank, instance) of every peer's buffer, so no separate collective kernel runs; the caller only needs a
 * cross-rank barrier before reading the buffer.  world = 0 or peer_bufs = NULL switches it off.  world <= 8.
 * Every rank must hold the SAME number of instances N (row block `rank`); for uneven shards, or to
 * drop the per-period barrier and the 8-byte remote stores, use bmpc_set_gather_pull. */
int bmpc_set_gather(bmpc_handle *h, void *const *peer_bufs, int32_t world, int32_t rank);

/* The gather WITHOUT a cross-rank barrier and without sender-side traffic (pull protocol, uneven shards allowed).
 * peer_bufs[p] is rank p's slot buffer [slots x N_p x n] doubles (N_p = row_offsets[p+1] - row_offsets[p] rows: rank p's own
 * instances), peer_flags[p] its flag array [2 world] of uint64 (zero-initialised); both mapped into this process (CUDA IPC /
 * symmetric memory).  Period e = 1, 2, ... of this handle (bmpc_gather_epoch() after the step): the step kernel's epilogue
 * stores Z̃ into slot e % slots of THIS rank's buffer -- local, coalesced stores -- and its last CTA publishes e into entry
 * `rank` of THIS rank's flag array with st.release.sys: a launch never touches remote memory.  A reader calls
 * bmpc_gather_pull(h, e, dst, stream): a kernel on `stream` (NULL = the handle's; a side stream overlaps the next period and
 * is made to wait for this rank's own launch of period e) that, per rank p, polls peer_flags[p][p] >= e over NVLink
 * (ld.acquire.sys) and copies rank p's rows of period e into dst rows [row_offsets[p], row_offsets[p+1]) of the device array
 * [row_offsets[world] x n] (dst = NULL: wait only), then publishes its ack (entry world + rank of every peer's flag array).
 * Once a handle has pulled, its step kernels wait, before overwriting a slot, until every rank has acknowledged the period
 * that lived there (normally satisfied at once; every wait gives up after ~2 s: bmpc_gather_timed_out).
 * slots >= 2; readers that lag L periods need slots >= L + 2. */
int bmpc_set_gather_pull(bmpc_handle *h, void *const *peer_bufs, void *const *peer_flags, int32_t world, int32_t rank,
                         const int32_t *row_offsets, int32_t slots);
int64_t bmpc_gather_epoch(bmpc_handle *h);
int bmpc_gather_pull(bmpc_handle *h, int64_t epoch, double *dst, void *stream);
int bmpc_gather_timed_out(bmpc_handle *h); /* 1 if a wait gave up (synchronises the device) */

/* Launch geometry actually used: {team, teams_per_cta, grid, smem_bytes_per_cta,
 * pd_in_smem, n_rows_m, n_sparse_rows, n_dense_rows} -- for DESIGN.md / bench reporting. */
int bmpc_launch_info(bmpc_handle *h, int32_t out[8]);
/* Number of kernel launches issued by this handle since creation (bench `gpu_launches`). */
int64_t bmpc_launch_count(bmpc_handle *h);

/* ------------------------------------------------------------------------------------------------
 * Linear MovingHorizonEstimator (LinModel + SingleShooting, direct = true or false), batched.
 * Replaces preparestate!/correct_estimate! and updatestate!/update_estimate! of the reference
 * (src/estimator/mhe/execute.jl:44-84): the handle owns the data windows, the arrival state and
 * the arrival covariance of every instance (estim.Y0m/U0/D0/X̂0_old/x̂0arr_old/P̂arr_old/invP̄, Nk).
 * Z̃ = [ε; x̂0(k-Nk); Ŵ] (slack FIRST, src/estimator/mhe/construct.jl:1174-1178).
 * ------------------------------------------------------------------------------------------------ */
typedef struct bmhe_handle bmhe_handle;

typedef struct {
    int32_t N;            /* estimator instances                                                   */
    int32_t nu, nym, nd;  /* inputs, measured outputs, measured disturbances                       */
    int32_t nxhat;        /* augmented state size (<= 32)                                          */
    int32_t He;           /* estimation horizon                                                    */
    int32_t neps;         /* 1 if Cwt finite                                                       */
    int32_t direct;       /* 1: current form (reference default, p = 0); 0: prediction form (p = 1) */
    int32_t shared_model; /* 1: matrices given once                                                */
    int32_t max_iter;     /* 0 -> 50                                                               */
    int32_t device;
    int32_t reserved;
    double tol;           /* 0 -> 1e-11                                                            */
} bmhe_dims;

int bmhe_create(bmhe_handle **out, const bmhe_dims *dims);
int bmhe_destroy(bmhe_handle *h);
/* estim.Ẽ (without the slack column: nym*He x nxhat*(He+1)), .G, .J, .B and con.ẼX̂ (same), GX̂, JX̂, BX̂
 * (init_predmat_mhe, src/estimator/mhe/transcription.jl:151-260), column-major, NM copies. */
int bmhe_set_predmat(bmhe_handle *h, const double *E, const double *G, const double *J, const double *B,
                     const double *EX, const double *GX, const double *JX, const double *BX);
/* estim.Â, Ĉm and the covariances P̂_0, Q̂, R̂ of cov (estimator/construct.jl:60-119; R̂ diagonal or dense SPD); resets. */
int bmhe_set_cov(bmhe_handle *h, const double *Ahat, const double *Cmhat, const double *P0, const double *Qhat,
                 const double *Rhat, double Cwt);
/* setconstraint!(mhe; x̂min, x̂max, ŵmin, ŵmax, v̂min, v̂max) in deviation form, N x len (NULL = none);
 * softness c_x (2*nxhat: min then max), c_w (2*nxhat), c_v (2*nym), shared, NULL = hard (the default). */
int bmhe_set_constraints(bmhe_handle *h, const double *xmin, const double *xmax, const double *wmin,
                         const double *wmax, const double *vmin, const double *vmax, const double *c_x,
                         const double *c_w, const double *c_v);
/* init_estimate_cov!: empty windows, Nk = 0, P̄ = P̂_0, x̂0 = 0 (execute.jl:2-37). */
int bmhe_reset(bmhe_handle *h);
/* preparestate!: add y0m(k), d0(k) (and the stored u0(k-1)) to the windows, correct P̄ when the window
 * moves, rebuild H̃/q̃, solve, return x̂0(k) (N x nxhat).  Optional outputs may be NULL. */
int bmhe_correct(bmhe_handle *h, const double *y0m, const double *d0, double *xhat0, double *Ztilde, double *J,
                 int32_t *status, int32_t *iters, double *Vhat, double *X0);
/* updatestate! (direct = 1): store u0(k); when the window is full, P̄ <- Â P̄ Â' + Q̂ (update_cov!). */
int bmhe_update(bmhe_handle *h, const double *u0);
/* updatestate! (direct = 0, update_estimate! src/estimator/mhe/execute.jl:71-84): add u0(k), y0m(k), d0(k) to the
 * windows, rebuild H̃/q̃ with the p = 1 prediction matrices, solve, return x̂0(k+1) (N x nxhat); then, when the
 * window is full, update_cov! = KalmanFilter correction + prediction of P̄ (kalman.jl:520-525).  With direct = 0
 * bmhe_correct (preparestate!) only returns the current x̂0, as correct_estimate! is empty in that mode. */
int bmhe_update_solve(bmhe_handle *h, const double *u0, const double *y0m, const double *d0, double *xhat0,
                      double *Ztilde, double *J, int32_t *status, int32_t *iters, double *Vhat, double *X0);
/* Run the per-period calls on the caller's stream (NULL = the handle's own); with sync = 0 they return without
 * synchronising it and their array arguments may be DEVICE pointers (the copies are cudaMemcpyDefault): the form a
 * device-resident caller (or a CUDA-event timed loop) uses.  Default: own stream, sync = 1, host pointers. */
int bmhe_set_stream(bmhe_handle *h, void *stream, int32_t sync);
int64_t bmhe_launch_count(bmhe_handle *h);

#ifdef __cplusplus
}
#endif
#endif /* BMPC_H */
