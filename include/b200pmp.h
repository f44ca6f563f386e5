/*
 * b200pmp.h -- C ABI of libb200pmp.so: batched Pontryagin (PMP) backward-shooting rollouts on B200.
 *
 * This is the drop-in boundary for ONE hot path of mbjd/approximate_optimal_control.  The reference
 * has no FFI (it is a Python closure API over JAX pytrees); every entry point below names the
 * reference interface it replaces.  All citations are relative to the reference checkout.
 *
 *   pmp_solve_batch_cuda   <- jax.vmap(solve_backward)            pontryagin_utils.py:488-527 as
 *                             vmapped at levelsets.py:601,1254; and the missing value-reparameterised
 *                             solver called at experiment_orbits.py:117,182, rrt_sampler.py:64-70
 *   pmp_solve_host         <- the same call as a host-language binding makes it: the reference's state dict
 *                             leaves in HOST memory in, endpoints in HOST memory out (levelsets.py:595-601,1215)
 *   pmp_rhs_batch_cuda     <- jax.vmap(f_extended)                pontryagin_utils.py:418-482
 *   pmp_ustar_batch_cuda   <- jax.vmap(u_star_2d) / u_star_new    pontryagin_utils.py:11-216, :532-562
 *                             (batched at levelsets.py:604-628 find_min_l)
 *   pmp_forward_batch_cuda <- jax.vmap(forward_sim_lqr)           levelsets.py:818-836, and
 *                             closed_loop_eval_general              eval_utils.py:48-99 (linear costate law)
 *   pmp_forward_nn_batch_cuda <- jax.vmap(forward_sim_nn / forward_sim_nn_until_value)  levelsets.py:838-933,
 *                             closed_loop_eval_nn_ensemble          eval_utils.py:10-45
 *   pmp_fma_peak_cuda      <- (no reference counterpart) register-resident FMA chain used as the
 *                             roofline denominator of bench.py
 *
 * Conventions
 *   - plain pointers and sizes; no torch / CUDA types in signatures (cuda_stream is a cudaStream_t
 *     passed as void*; NULL = the legacy default stream).
 *   - every *_cuda pointer argument is a DEVICE pointer owned by the caller; nothing is allocated,
 *     nothing is freed, no hidden host<->device copy, no synchronisation: the call enqueues work on
 *     `cuda_stream` and returns.  The one exception is pmp_solve_host (host pointers, synchronous, owns its
 *     staging memory and streams).
 *   - return value 0 = ok, <0 = error (PMP_ERR_*); pmp_last_error() gives a thread-local message.
 *     Functions never throw and never abort.
 *   - real arrays are float (dtype 0) or double (dtype 1), chosen by pmp_opts_t.dtype.
 *
 * ODE state of one trajectory ("extended state"), NS = 2*nx + 2 scalars, in the order of the
 * reference's state dict {'x','t','v','vx'} (pontryagin_utils.py:454-458):
 *       row 0 .. nx-1        x        state
 *       row nx               t        time (integrated as a state with dt/dt = 1)
 *       row nx+1             v        value / cost-to-go
 *       row nx+2 .. 2nx+1    vx       costate  lambda = V_x
 * Batches are struct-of-arrays: y[row * n + i] is component `row` of trajectory i, so that a warp
 * reading 32 consecutive trajectories issues one coalesced 128-byte load per row.
 */
#ifndef B200PMP_H
#define B200PMP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMP_ABI_VERSION 5

/* systems (dynamics are compile-time functors; problem_params['system_name'] selects one) */
enum {
    PMP_SYS_FLATQUAD = 0,         /* flatquad_landing_experiment.py:269-365   nx=6 nu=2 */
    PMP_SYS_ORBITS = 1,           /* experiment_orbits.py:24-40,83-104        nx=2 nu=1 */
    PMP_SYS_LQR_STATEPENALTY = 2, /* experiment_lqr_statepenalty.py:23-38     nx=2 nu=1 */
    /* plug-in dynamics: a functor generated from the user's problem_params['f'], ['l'] (sympy derivatives in
     * place of jax autodiff, pontryagin_utils.py:35-38,438-439) and compiled, with the same kernels, into a
     * library of its own that exports this same ABI (approximate_optimal_control_b200/codegen.py).  Only that
     * library accepts this id; nx <= 16, nu in {1, 2}; params[] unused (constants are baked into the code). */
    PMP_SYS_GENERATED = 3
};

/* integrators */
enum {
    PMP_METHOD_RK4 = 0,   /* classic fixed-step RK4 (BASELINE.json config 0) */
    PMP_METHOD_TSIT5 = 1, /* diffrax.Tsit5, the only solver the reference uses (pontryagin_utils.py:515,522) */
    PMP_METHOD_DOPRI5 = 2 /* Dormand-Prince 5(4) (BASELINE.json config 1) */
};

enum { PMP_F32 = 0, PMP_F64 = 1 };

/* independent variable */
enum {
    PMP_INDEP_TIME = 0, /* solve_backward, pontryagin_utils.py:488-527 */
    PMP_INDEP_VALUE = 1 /* value-reparameterised solver (experiment_orbits.py:109-117) */
};

/* terminating events (diffrax.DiscreteTerminatingEvent, evaluated after every step attempt) */
enum {
    PMP_EVENT_NONE = 0,
    PMP_EVENT_VALUE_GE = 1, /* stop once v >= event_level (value level set, problem_params['V_max']) */
    PMP_EVENT_VXX_NORM = 2, /* stop once ||vxx||_F > event_level (algo_params['vxx_max_norm'],
                               pontryagin_utils.py:508-519); needs PMP_FLAG_VXX */
    PMP_EVENT_LQR_VALUE_LE = 3, /* pmp_forward_batch_cuda only: stop once 1/2 (x-x_eq)'P(x-x_eq) <= event_level
                                   (inside the value level set; LQR analogue of levelsets.py:897-915) */
    PMP_EVENT_NN_VALUE_UCB_LE = 4 /* pmp_forward_nn_batch_cuda only: stop once v_mean + n_sigma v_std <= event_level
                                     AND v_std <= sigma_max over the ensemble (levelsets.py:897-913) */
};

/* per-trajectory result codes, diffrax-compatible (levelsets.py:1218 tests result == 1) */
enum {
    PMP_RESULT_OK = 0,
    PMP_RESULT_EVENT = 1,
    PMP_RESULT_MAX_STEPS = 2,
    PMP_RESULT_DT_MIN = 3,
    PMP_RESULT_NAN = 4
};

/* pmp_opts_t.flags */
enum {
    /* nu = 2 systems: pick u* among the four box-edge minimisers exactly as u_star_2d does
     * (pontryagin_utils.py:88-158: evaluate H at every candidate, argmin, re-expand about the winner,
     * argmin again).  Default (flag clear): pick the edge minimiser whose KKT multiplier is
     * non-negative -- the same point in exact arithmetic (the box-constrained minimiser of a strictly
     * convex quadratic is the unique KKT point), better conditioned in floating point than comparing
     * H values, and ~2.5x fewer instructions.  pmp_ustar_batch_cuda always uses the literal form. */
    PMP_FLAG_USTAR_LITERAL = 1,
    /* algo_params['pontryagin_solver_vxx']: also propagate the value Hessian V_xx along the characteristic,
     * vxx' = gx + glam vxx - vxx fx - vxx flam vxx with (fx, flam, gx, glam) the Jacobians of the PMP right
     * hand side THROUGH u* (pontryagin_utils.py:460-467).  The state gains nx*nx rows after vx (row-major
     * vxx[i][j] at row 2nx+2 + i*nx + j), the error norm one more leaf.  The flatquad functor carries the
     * Jacobians in closed form; a generated system (PMP_SYS_GENERATED, nx <= 9) is differentiated
     * in forward mode, as the reference's jax.jacobian(pmp_rhs) handles any f, l (csrc/pmp_vxx_dual.cuh).  orbits /
     * LQR: PMP_ERR_UNSUPPORTED (u_star_new systems; the reference's vxx branch runs through u_star_2d). */
    PMP_FLAG_VXX = 2,
    /* with PMP_FLAG_VXX: the caller guarantees that every terminal vxx is symmetric (vxx[i][j] == vxx[j][i]
     * bit for bit; the reference starts from the LQR solution P, levelsets.py:595).  vxx^T obeys the same
     * Riccati equation, so the solution stays symmetric: the kernel then reads, integrates and stores only
     * the upper triangle (21 of 36 entries; off-diagonal entries count twice in the error norm and in
     * ||vxx||_F) and mirrors it on output.  Differs from the general path by rounding only. */
    PMP_FLAG_VXX_SYMMETRIC = 4,
    /* pmp_ustar_batch_cuda_flags only: u_star_2d(..., smooth=True) (pontryagin_utils.py:217-254), the softmax blend of the
     * five candidates (smoothing length 0.01) the reference keeps as a debugging aid; nu == 2 systems.  Not a solver option. */
    PMP_FLAG_USTAR_SMOOTH = 8
};

/* error codes */
enum {
    PMP_OK = 0,
    PMP_ERR_BAD_ARG = -1,
    PMP_ERR_UNSUPPORTED = -2,
    PMP_ERR_CUDA = -3,
    PMP_ERR_NO_DEVICE = -4
};

typedef struct {
    int32_t system_id;
    int32_t nx, nu;
    double u_lo[2], u_hi[2]; /* problem_params['U_interval'] */
    /* flatquad:  m, g, r, I, q_pos, q_ang, q_vel, q_omega   (q = 1/length_scale^2)
     * orbits:    w_dist (0.1), w_u (10)
     * lqr:       a (state-penalty steepness 0.01), r_u (1) */
    double params[16];
} pmp_problem_t;

typedef struct {
    int32_t method;      /* PMP_METHOD_*                                                  */
    int32_t dtype;       /* PMP_F32 / PMP_F64                                             */
    int32_t indep;       /* PMP_INDEP_*                                                   */
    int32_t adaptive;    /* 1: PIDController (I-controller) step control; 0: constant dt0 */
    int32_t max_steps;   /* step ATTEMPTS (accepted + rejected), diffrax max_steps        */
    int32_t force_dtmin; /* diffrax PIDController.force_dtmin (default 1)                 */
    int32_t event;       /* PMP_EVENT_*                                                   */
    int32_t save_dense;  /* 1: every accepted step is written to pmp_dense_t              */
    double t0, t1, dt0;  /* t1 < t0 integrates backward (pontryagin_utils.py:515: 0 -> -T, dt0=-0.1);
                            PMP_INDEP_VALUE: t0 is ignored (each trajectory starts at its own v),
                            t1 = V_max, dt0 > 0 in value units                             */
    double rtol, atol;   /* algo_params['pontryagin_solver_rtol'|'_atol']                  */
    double dtmin, dtmax; /* 0.0005, 0.5 hard-coded at pontryagin_utils.py:500-501; <=0 disables */
    double safety, factormin, factormax; /* diffrax defaults .9, .2, 10                   */
    double event_level;
    int32_t flags;       /* PMP_FLAG_* bit mask (0 = defaults)                            */
    int32_t reserved;    /* must be 0                                                     */
} pmp_opts_t;

/* Dense ("SaveAt(steps=True, t0=True)") output, the reference's Solution layout: slot 0 is the
 * terminal state, one slot per ACCEPTED step, unused slots are +inf (ts: -inf when integrating
 * backward, cf. visualiser.py:76-99).  S1 = max_steps + 1 slots per trajectory.  All five leaves are
 * required, 16-byte aligned, and n * S1 must be < 2^32. */
typedef struct {
    void *ts; /* [n][S1]     */
    void *x;  /* [n][S1][nx] */
    void *t;  /* [n][S1]     */
    void *v;  /* [n][S1]     */
    void *vx; /* [n][S1][nx] */
    void *vxx; /* [n][S1][nx][nx], required with PMP_FLAG_VXX, ignored (may be NULL) otherwise */
} pmp_dense_t;

int pmp_abi_version(void);
const char *pmp_last_error(void);
int pmp_state_rows(const pmp_problem_t *problem); /* NS = 2*nx+2 */
int pmp_state_rows_flags(const pmp_problem_t *problem, int32_t flags); /* + nx*nx with PMP_FLAG_VXX */

/* number of CUDA devices visible to this process (<0: error, 0: none) */
int pmp_device_count(void);

/* Scratch the solve needs for n trajectories (work-queue counter + the first-stage derivative
 * k1 = F(y_f) of every trajectory, [2nx+1 (+ nx*nx with PMP_FLAG_VXX)][n] reals).  The caller allocates it on the device. */
int64_t pmp_solve_workspace_bytes(const pmp_problem_t *problem, const pmp_opts_t *opts, int64_t n);

/* Integrate n trajectories from their terminal states y_f [NS][n].
 *   y_end      [NS][n]  last accepted state (== ys[num_accepted_steps], levelsets.py:1215)
 *   dense      may be NULL unless opts->save_dense
 *   n_accepted,n_steps,result  [n] int32 (any may be NULL)
 *   workspace  device scratch of >= pmp_solve_workspace_bytes() bytes, 256-byte aligned; its
 *              content is undefined afterwards; must not be shared by concurrent calls
 * Work is enqueued on cuda_stream of the CURRENT device; returns immediately. */
int pmp_solve_batch_cuda(const pmp_problem_t *problem, const pmp_opts_t *opts, int64_t n,
                         const void *y_f, void *y_end, const pmp_dense_t *dense,
                         int32_t *n_accepted, int32_t *n_steps, int32_t *result,
                         void *workspace, int64_t workspace_bytes, void *cuda_stream);

/* Host in -> host out.  The terminal states arrive as the leaves of the reference's state dict (pontryagin_utils.py:
 * 454-458), row-major in HOST memory: x_f [n][nx], t_f [n] (NULL: every trajectory starts at opts->t0), v_f [n],
 * vx_f [n][nx]; the endpoints (ys[num_accepted_steps], levelsets.py:1215) come back the same way in x_end, t_end,
 * v_end, vx_end, with n_accepted, n_steps, result [n] int32 (any output may be NULL).  The batch is cut into n_pieces
 * pieces (0 = choose: a small first piece, growing to n/8) whose host->device copies, integration and device->host copies overlap on the library's own
 * streams of the CURRENT device; the copies overlap only from/to page-locked memory (cudaHostAlloc / cudaHostRegister,
 * torch pin_memory), pageable buffers work but serialise.  Returns when every output is in host memory.  Device
 * staging (~(8nx+7) reals per trajectory of three pieces) is kept between calls; pmp_host_release() frees it.
 * The three counters cross the link as one packed word per trajectory (max_steps < 2^14) and are unfolded by the
 * calling thread while later pieces are in flight.  Concurrent calls on one device are serialised.  Endpoints only: save_dense and PMP_FLAG_VXX are refused.
 * Batches of >= 2^16 trajectories (time-parameterised, default flags) run as ONE persistent launch of the integrator that
 * consumes the pieces as the copy engine delivers them (csrc/pmp_host.cu); smaller ones as one launch per piece.
 * Results are bit-identical to pmp_solve_batch_cuda on the same states. */
int pmp_solve_host(const pmp_problem_t *problem, const pmp_opts_t *opts, int64_t n,
                   const void *x_f, const void *t_f, const void *v_f, const void *vx_f,
                   void *x_end, void *t_end, void *v_end, void *vx_end,
                   int32_t *n_accepted, int32_t *n_steps, int32_t *result, int32_t n_pieces);
/* The same with the terminal set of levelsets.py:580-601 (`solve_backward_lqr`, the function the reference vmaps over its
 * terminal samples `xfs`): only x_f [n][nx] crosses the link, the terminal costate and value are formed on the device,
 * vx_f = P_lqr (x_f - x_eq), v_f = 1/2 (x_f - x_eq)' vx_f, in the solve's precision (as jax does inside the vmap).
 * P_lqr [nx][nx] row-major and x_eq [nx] (NULL = 0) are HOST doubles; t_f as in pmp_solve_host (NULL: opts->t0). */
int pmp_solve_host_lqr(const pmp_problem_t *problem, const pmp_opts_t *opts, int64_t n,
                       const void *x_f, const void *t_f, const double *P_lqr, const double *x_eq,
                       void *x_end, void *t_end, void *v_end, void *vx_end,
                       int32_t *n_accepted, int32_t *n_steps, int32_t *result, int32_t n_pieces);
/* kernels + copies one pmp_solve_host call enqueues: *kernels, *copies (bench bookkeeping); returns the piece count */
int pmp_solve_host_plan(const pmp_problem_t *problem, const pmp_opts_t *opts, int64_t n, int32_t n_pieces,
                        int32_t *kernels, int32_t *copies);
int pmp_host_release(void);
/* Diagnostics: a plain copy kernel (16 bytes per thread and iteration, `blocks` CTAs of 256) between two unified-address
 * pointers; with a pinned host pointer on one side it measures what SM loads / stores move over the host link, next to
 * the copy engines (scripts/pcie_probe.py).  bytes must be a multiple of 16.  Asynchronous on cuda_stream. */
int pmp_host_probe_copy(void *dst, const void *src, int64_t bytes, int32_t blocks, void *cuda_stream);

/* ---- multi-GPU: one process per GPU, terminal samples sharded by index, no exchange during the solve, then ONE all-gather of
 * the endpoint records so that every rank holds the whole (x, t, v, lambda) dataset for the value fit.
 * NCCL behind plain C entries (resolved with dlopen at the first call: no link-time dependency), for hosts that are not
 * PyTorch.  Bootstrap as with NCCL itself: rank 0 calls pmp_comm_unique_id, ships the 128 bytes to the other ranks by any
 * means (MPI, a file, jax.distributed), every rank -- its CUDA device current -- calls pmp_comm_init.
 * pmp_allgather_endpoints: local [rows][n_local] (this rank's y_end) -> global [world][rows][n_local] on every rank,
 * rank-major; asynchronous on cuda_stream.  (PyTorch hosts have sharding.PipelinedGather, which hides the transfer behind
 * the next solve with peer-to-peer copies.) */
int pmp_comm_unique_id(void *id128 /* 128 bytes, host */);
int pmp_comm_init(void **comm, int32_t world, int32_t rank, const void *id128);
int pmp_comm_destroy(void *comm);
int pmp_allgather_endpoints(void *comm, const void *local, void *global, int64_t n_local, int32_t rows, int32_t dtype,
                            void *cuda_stream);

/* The same gather FUSED into the integrator (north star: "shard by terminal-sample index, with only an all-gather of endpoint
 * (x, lambda, v) datasets"; no counterpart in the single-process reference).  pmp_solve_batch_cuda for time-parameterised endpoint
 * solves (flags 0, save_dense 0) that, besides writing y_end [NS][n] locally, stores every finished work-queue chunk (32 consecutive
 * trajectories = 128 contiguous bytes per fp32 row) into peer_slots[0 .. n_peers): device pointers, valid on THIS device (peer
 * access / symmetric memory over NVLink), to this rank's [NS][n] block inside each peer's gather buffer.  The records travel while
 * the rest of the shard is integrated: when the kernel has finished, so has this rank's part of the gather -- no copy engine, no
 * NCCL kernel competing for SMs, no transfer left over after the last step.  The caller still tells the peers (a barrier / stream
 * event exchange after the launch) and must not let a peer reuse the buffer before every rank has read it.
 *   multicast_slot  NULL, or the address of the same block through an NVSwitch multicast (multimem) mapping of the gather
 *                   buffer: one multimem.st per row then reaches every GPU of the group and peer_slots is not used
 * Everything else as pmp_solve_batch_cuda. */
int pmp_solve_batch_push_cuda(const pmp_problem_t *problem, const pmp_opts_t *opts, int64_t n, const void *y_f, void *y_end,
                              int32_t *n_accepted, int32_t *n_steps, int32_t *result, void *workspace,
                              int64_t workspace_bytes, void *const *peer_slots, int32_t n_peers, void *multicast_slot,
                              void *cuda_stream);

/* Forward-mode sensitivity of a solve: endpoint AND its directional derivative along dy_f, d y_end / d y_f . dy_f, for n
 * trajectories (y_f, dy_f, y_end, dy_end [NS][n]).  Replaces jax.jacobian / jvp through diffrax.diffeqsolve
 * (experiment_orbits.py:186-197 differentiates the value-parameterised solve, indep = PMP_INDEP_VALUE, with respect to the
 * terminal state; the time-parameterised solve works the same way).  Dual numbers through the dynamics functor and the
 * Runge-Kutta steps; step control, clipping and argmin decisions on primal values, as in jax's jvp of the traced program;
 * t0 / t1 / v0 / v1 carry no tangent.  Works for every compiled and generated system.  Endpoints only, flags 0. */
int pmp_solve_jvp_cuda(const pmp_problem_t *problem, const pmp_opts_t *opts, int64_t n, const void *y_f, const void *dy_f,
                       void *y_end, void *dy_end, int32_t *n_accepted, int32_t *n_steps, int32_t *result,
                       void *cuda_stream);

/* ---- data selection after the solve: select_train_pts (levelsets.py:667-709) + data_normaliser statistics
 * (nn_utils.py:112-134) on the device-resident SaveAt(steps) buffers of pmp_solve_batch_cuda (save_dense).
 * Pass 1: counts[i] = number of saved nodes of trajectory i with v_lower <= v <= v_upper (and ||vxx||_F < vxx_max_norm when
 * vxx != NULL; nvxx = nx*nx).  n_accepted (optional) bounds the slots read to 0..n_accepted[i]; padding never passes the
 * test anyway.  v [n][s1], vxx [n][s1][nvxx]. */
int pmp_select_count_cuda(int32_t dtype, int64_t n, int32_t s1, const void *v, const void *vxx, int32_t nvxx,
                          const int32_t *n_accepted, double v_lower, double v_upper, double vxx_max_norm,
                          int32_t *counts, void *cuda_stream);
/* Pass 2: offsets [n] = exclusive prefix sum of counts (int64, device).  Writes the selected nodes of every leaf to flat
 * arrays in the order of the reference's boolean-mask gather (trajectory by trajectory, slot by slot): x_out [M][nx],
 * t_out, v_out [M], vx_out [M][nx] (, vxx_out [M][nvxx]); ADDS to stats [2 nx + 2] (device doubles, zeroed by the caller)
 * sum x[nx], sum x^2[nx], sum v, sum v^2 over the selected nodes, and ORs 1 into flags[0] if a selected node holds a NaN. */
int pmp_select_write_cuda(int32_t dtype, int64_t n, int32_t s1, int32_t nx, const void *x, const void *t, const void *v,
                          const void *vx, const void *vxx, int32_t nvxx, const int32_t *n_accepted,
                          const int64_t *offsets, double v_lower, double v_upper, double vxx_max_norm,
                          void *x_out, void *t_out, void *v_out, void *vx_out, void *vxx_out,
                          double *stats, int32_t *flags, void *cuda_stream);

/* number of kernels one pmp_solve_batch_cuda call launches with these options (bench bookkeeping) */
int pmp_solve_launch_count(const pmp_problem_t *problem, const pmp_opts_t *opts, int64_t n);

/* forward-time extended RHS f_extended(t, state) for n states: y [NS][n] -> dy [NS][n] */
int pmp_rhs_batch_cuda(const pmp_problem_t *problem, int32_t dtype, int64_t n, const void *y,
                       void *dy, void *cuda_stream);

/* Same with opts-style flags: PMP_FLAG_VXX adds the vxx rows (f_extended with
 * algo_params['pontryagin_solver_vxx'], pontryagin_utils.py:460-467): y, dy [NS + nx*nx][n]; flatquad only */
int pmp_rhs_batch_cuda_flags(const pmp_problem_t *problem, int32_t dtype, int32_t flags, int64_t n,
                             const void *y, void *dy, void *cuda_stream);

/* u* = argmin_u H(x,u,lambda) over the input box: x [nx][n], lam [nx][n] -> u [nu][n] */
int pmp_ustar_batch_cuda(const pmp_problem_t *problem, int32_t dtype, int64_t n, const void *x,
                         const void *lam, void *u, void *cuda_stream);

/* Forward closed-loop simulation x' = f(x, u*(x, lambda(x))) under the LINEAR costate law lambda(x) = P (x - x_eq)
 * (levelsets.py:818-836 forward_sim_lqr), with the accumulated running cost as an extra state, cost' = l(x, u*)
 * (eval_utils.py:48-99 closed_loop_eval_general).  opts: method, dtype, adaptive (0 = the constant step dt0 of
 * eval_utils), t0 < t1, dt0 > 0, rtol / atol / dtmin / dtmax / controller factors, max_steps, event (NONE or
 * PMP_EVENT_LQR_VALUE_LE), save_dense; indep must be PMP_INDEP_TIME and flags 0.  The error norm runs over the x leaf.
 *   x0      [nx][n]    initial states (device)
 *   P, x_eq HOST arrays of doubles, nx*nx row-major and nx (x_eq may be NULL = 0); read during the call
 *   z_end   [nx+1][n]  final state and accumulated cost (device)
 *   dense_ts [n][S1], dense_x [n][S1][nx], dense_cost [n][S1]: required iff opts->save_dense; slot 0 = x0, one slot
 *           per accepted step, +inf padded
 *   n_accepted, n_steps, result [n] int32 (any may be NULL); result codes as for pmp_solve_batch_cuda */
int pmp_forward_batch_cuda(const pmp_problem_t *problem, const pmp_opts_t *opts, int64_t n, const void *x0,
                           const double *P, const double *x_eq, void *z_end, void *dense_ts, void *dense_x,
                           void *dense_cost, int32_t *n_accepted, int32_t *n_steps, int32_t *result,
                           void *cuda_stream);

/* Ensemble of value-function networks as a costate law (levelsets.py:726-778,838-933; nn_utils.py:119-199): member e is
 * the flax MLP Dense(nx -> width) softplus, (n_hidden_layers - 1) x [Dense(width -> width) softplus], Dense(width -> 1)
 * on the normalised state, V_e(x) = v_std * MLP_e((x - x_mean) / x_std) + v_mean; lambda(x) = grad of mean_e V_e(x).
 * weights: DEVICE array of the dtype of opts, per member (flax kernels are [in][out], row-major):
 *     K1 [nx][width], b1 [width], (n_hidden_layers - 1) x { K [width][width], b [width] }, Kout [width], bout, 3 pad reals
 * 16-byte aligned; width in {32, 64}; n_hidden_layers in 1..4; n_nets in 1..16. */
typedef struct {
    int32_t n_nets, n_hidden_layers, width, reserved;
    const void *weights;
    double x_mean[16], x_std[16]; /* data_normaliser: first nx entries (identity: 0 and 1)   */
    double v_mean, v_std;         /* (identity: 0 and 1)                                      */
    double n_sigma, sigma_max;    /* PMP_EVENT_NN_VALUE_UCB_LE: 2 and 0.5 in the reference    */
} pmp_nn_costate_t;

/* pmp_forward_batch_cuda with the costate law of a value-network ensemble in place of P (x - x_eq):
 * jax.vmap(forward_sim_nn) / forward_sim_nn_until_value (levelsets.py:838-933) and, with a constant step,
 * closed_loop_eval_nn_ensemble (eval_utils.py:10-45).  event: NONE or PMP_EVENT_NN_VALUE_UCB_LE. */
int pmp_forward_nn_batch_cuda(const pmp_problem_t *problem, const pmp_opts_t *opts, const pmp_nn_costate_t *nn,
                              int64_t n, const void *x0, void *z_end, void *dense_ts, void *dense_x, void *dense_cost,
                              int32_t *n_accepted, int32_t *n_steps, int32_t *result, void *cuda_stream);

/* The ensemble itself at n states (v_meanstds / vx_meanstd, levelsets.py:789-815): x [nx][n] -> lam [nx][n] (gradient of
 * the ensemble mean), v_mean [n], v_std [n] (population std over the members); dtype selects float / double buffers and
 * weights; problem supplies nx. */
int pmp_nn_value_batch_cuda(const pmp_problem_t *problem, const pmp_nn_costate_t *nn, int32_t dtype, int64_t n,
                            const void *x, void *lam, void *v_mean, void *v_std, void *cuda_stream);

/* Dense output, diffrax Solution.evaluate (SaveAt(dense=True), pontryagin_utils.py:504; used at
 * experiment_orbits.py:183-184, misc.py:40-95).  For every query q: y0[:, q] is the saved node at the
 * START of the accepted step that contains the query (rows x, t, v, vx as everywhere), t1[q] the saved
 * time at the END of that step (the value, for PMP_INDEP_VALUE solves).  The step's stage derivatives are
 * recomputed with opts->method / dtype / indep / flags and the solver's interpolant is evaluated:
 *   mode 0: at the independent variable target[q]                      -> y_out [NS][nq]
 *   mode 1: at the point of the step where v == target[q] (bisection; time-parameterised solves only;
 *           NaN if the level is not bracketed by the step)            -> y_out [NS][nq] */
int pmp_interp_batch_cuda(const pmp_problem_t *problem, const pmp_opts_t *opts, int64_t nq,
                          const void *y0, const void *t1, const void *target, int32_t mode,
                          void *y_out, void *cuda_stream);

/* Same with the selection rule the integrator uses for `flags` (PMP_FLAG_USTAR_LITERAL set: identical
 * to pmp_ustar_batch_cuda; clear: KKT-multiplier selection).  For verification of the two rules.
 * flags == PMP_FLAG_USTAR_SMOOTH: u_star_2d(x, costate, problem_params, smooth=True). */
int pmp_ustar_batch_cuda_flags(const pmp_problem_t *problem, int32_t dtype, int32_t flags, int64_t n,
                               const void *x, const void *lam, void *u, void *cuda_stream);

/* Register-resident dependent-FMA chains on every SM (roofline denominator of bench.py): enqueues
 * one kernel that executes *flops_out floating point operations (written on the host, may be
 * NULL); the caller times it with events on cuda_stream.  dtype selects FFMA / DFMA.
 * variant 0: a = fma(a, b, c) with register operands; variant 1: multiplier/addend as
 * compile-time immediates (the form the integrator's tableau FMAs take); variant 2 (fp32 only; fp64 runs
 * variant 1): the same chains as packed pairs (FFMA2 / fma.rn.f32x2), half the issue slots per flop. */
int pmp_fma_peak_cuda(int32_t dtype, int32_t variant, int32_t iters, void *scratch /* >= 8 B device */,
                      double *flops_out, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* B200PMP_H */
