/* nmpc_b200 -- C ABI of the batched DDP/iLQR and FMPC engines (libnmpc_b200.so).
 *
 * This is the drop-in boundary for the hot path of isri-aist/NMPC.  The reference has no FFI of its
 * own: its boundary is the C++ template API nmpc_ddp::DDPSolver<StateDim, InputDim>
 * (nmpc_ddp/include/nmpc_ddp/DDPSolver.h:23-375) and nmpc_fmpc::FmpcSolver<StateDim, InputDim,
 * IneqDim> (nmpc_fmpc/include/nmpc_fmpc/FmpcSolver.h:22-424).  Each entry point below names the
 * reference member it stands in for; include/nmpc_ddp/DDPSolver.h and
 * include/nmpc_fmpc/FmpcSolver.h re-create the template API on top of these calls.
 *
 * Conventions
 *  - plain C types only; every function returns an nmpc_b200_status (0 = OK) and never throws;
 *    nmpc_b200_last_error() returns the message of the last failure on the calling thread.
 *  - all host- or device-side I/O arrays are instance-major ("one solver object after another"):
 *      x0[B][NX], u[B][N][NU], x[B][N+1][NX], cost_list[B][N+1], k[B][N][NU],
 *      K[B][N][NU*NX] (column-major NU x NX per step, like Eigen), trace[B][max_iter+1][9].
 *    Inside the engine the batch index is the fastest-varying one.
 *  - `on_device` != 0 means the pointers are device pointers on the handle's device; the call is
 *    then asynchronous on `stream` (a cudaStream_t passed as void*, NULL = the handle's own stream).
 *  - a handle is not re-entrant (like a DDPSolver object); distinct handles are independent.
 *  - there is no CPU fallback: creation fails with NMPC_B200_ERR_NO_DEVICE when no CUDA device
 *    is usable.
 */
#ifndef NMPC_B200_C_API_H
#define NMPC_B200_C_API_H

#include <stddef.h>

#ifdef __cplusplus
extern "C"
{
#endif

#define NMPC_B200_VERSION 100

  typedef enum
  {
    NMPC_B200_OK = 0,
    NMPC_B200_ERR_INVALID_ARGUMENT = 1, /* std::invalid_argument in the reference (DDPSolver.hpp:41-45) */
    NMPC_B200_ERR_RUNTIME = 2, /* std::runtime_error in the reference (DDPSolver.hpp:51-56, :393) */
    NMPC_B200_ERR_UNKNOWN_MODEL = 3,
    NMPC_B200_ERR_NO_DEVICE = 4,
    NMPC_B200_ERR_CUDA = 5,
    NMPC_B200_ERR_CAPACITY = 6,
    NMPC_B200_ERR_UNSUPPORTED = 7
  } nmpc_b200_status;

  /** Message of the last error raised on this thread ("" if none). */
  const char * nmpc_b200_last_error(void);
  int nmpc_b200_version(void);
  /** Number of usable CUDA devices (0 when the driver or a device is missing). */
  int nmpc_b200_device_count(void);

  /* ------------------------------------------------------------------ problem functors ---- */

  /** Dimensions of a registered problem functor (stands in for DDPProblem::stateDim()/inputDim(),
      DDPProblem.h:52-85, and FmpcProblem::ineqDim(), FmpcProblem.h:62-86).  `ng` is 0 for functors
      without inequality constraints. */
  int nmpc_b200_model_dims(const char * model, int * nx, int * nu, int * ng, int * n_params);
  /** Default flat parameter vector of a registered functor (n_params doubles). */
  int nmpc_b200_model_default_params(const char * model, double * params);
  /** Number of registered functors and their names (for diagnostics). */
  int nmpc_b200_model_count(void);
  const char * nmpc_b200_model_name(int index);

  /** Evaluate a functor ON THE DEVICE at `n` points (derivative checks, TestDDPCartPole.cpp:609-649).
      Inputs (host): t[n], x[n][NX], u[n][NU].  Outputs (host, any may be NULL): x_next[n][NX],
      running_cost[n], terminal_cost[n], Fx[n][NX*NX], Fu[n][NX*NU], Lx[n][NX], Lu[n][NU],
      Lxx[n][NX*NX], Luu[n][NU*NU], Lxu[n][NX*NU], Vx[n][NX], Vxx[n][NX*NX], g[n][NG], C[n][NG*NX],
      D[n][NG*NU]; matrices column-major. */
  int nmpc_b200_model_eval(const char * model,
                           const double * params,
                           int n_params,
                           int device,
                           int n,
                           const double * t,
                           const double * x,
                           const double * u,
                           double * x_next,
                           double * running_cost,
                           double * terminal_cost,
                           double * Fx,
                           double * Fu,
                           double * Lx,
                           double * Lu,
                           double * Lxx,
                           double * Luu,
                           double * Lxu,
                           double * Vx,
                           double * Vxx,
                           double * g,
                           double * C,
                           double * D);

  /* ----------------------------------------------------------------------------- DDP ---- */

  /** Mirror of DDPSolver::Configuration (DDPSolver.h:47-110), same defaults.  print_level has no
      meaning for a batch; use_state_eq_second_derivative is rejected exactly where the reference
      throws (DDPSolver.hpp:391-414).  ONE LIMIT the reference does not have: alpha_list holds at most 16 candidates
      (the reference's std::vector takes any length, default 11) -- the line-search kernels evaluate the candidates of
      an instance on the lanes of a half warp; a longer list is NMPC_B200_ERR_INVALID_ARGUMENT at set_config. */
  typedef struct
  {
    int horizon_steps; /* 100 */
    int max_iter; /* 500 */
    int reg_type; /* 1: Quu + lambda I, 2: Vxx + lambda I */
    int with_input_constraint; /* 0 */
    int n_alpha; /* 11 */
    int use_state_eq_second_derivative; /* 0; non-zero => NMPC_B200_ERR_RUNTIME at solve() */
    double initial_lambda; /* 1e-4 */
    double initial_dlambda; /* 1.0 */
    double lambda_factor; /* 1.6 */
    double lambda_min; /* 1e-6 */
    double lambda_max; /* 1e10 */
    double k_rel_norm_thre; /* 1e-4 */
    double lambda_thre; /* 1e-5 */
    double cost_update_ratio_thre; /* 0 */
    double cost_update_thre; /* 1e-7 */
    double alpha_list[16]; /* 10^linspace(0,-3,11) */
  } nmpc_b200_ddp_config;

  void nmpc_b200_ddp_config_default(nmpc_b200_ddp_config * cfg);

  typedef struct nmpc_b200_ddp nmpc_b200_ddp; /* opaque: a batch of DDPSolver objects on one GPU */

  /** DDPSolver::DDPSolver(problem) (DDPSolver.hpp:20-24) for `batch_capacity` instances that share one
      problem functor and parameter set.  `device` is a CUDA ordinal. */
  int nmpc_b200_ddp_create(const char * model,
                           const double * params,
                           int n_params,
                           const nmpc_b200_ddp_config * cfg,
                           int batch_capacity,
                           int device,
                           nmpc_b200_ddp ** out);
  int nmpc_b200_ddp_destroy(nmpc_b200_ddp * h);

  /** DDPSolver::config() (DDPSolver.h:258-267).  Changing horizon_steps or max_iter reallocates. */
  int nmpc_b200_ddp_set_config(nmpc_b200_ddp * h, const nmpc_b200_ddp_config * cfg);
  int nmpc_b200_ddp_get_config(const nmpc_b200_ddp * h, nmpc_b200_ddp_config * cfg);

  /** DDPSolver::setInputLimitsFunc (DDPSolver.h:282-285) for limits constant over the horizon:
      lower[NU], upper[NU] (host).  Call after the horizon is configured (the values are replicated per step). */
  int nmpc_b200_ddp_set_input_limits(nmpc_b200_ddp * h, const double * lower, const double * upper);
  /** DDPSolver::setInputLimitsFunc for limits that CHANGE along the horizon: the values of the reference's
      input_limits_func_(current_t + i dt) (DDPSolver.hpp:470) for i = 0 .. horizon_steps-1, lower / upper
      [n_steps][NU] (host).  To be given again before a solve with another current_t or horizon. */
  int nmpc_b200_ddp_set_input_limits_horizon(nmpc_b200_ddp * h, int n_steps, const double * lower, const double * upper);
  /** The same for the device-resident MPC loop (nmpc_b200_ddp_run_mpc), where the reference evaluates
      input_limits_func_ anew at every solve (DDPSolver.hpp:470): lower / upper [n_ticks][n_steps][NU] (host) =
      input_limits_func_(current_t + tick * tick_dt + i dt).  Needed only when the limits depend on time; run_mpc with
      n_ticks beyond the table (or without one) answers NMPC_B200_ERR_UNSUPPORTED for such limits. */
  int nmpc_b200_ddp_set_input_limits_mpc(nmpc_b200_ddp * h, int n_ticks, int n_steps, const double * lower, const double * upper);

  /** DDPSolver::solve(current_t, current_x, initial_u_list) (DDPSolver.hpp:27-141) for B <= capacity
      independent instances: x0[B][NX], u_init[B][N][NU].  n_u_steps must equal horizon_steps
      (else NMPC_B200_ERR_INVALID_ARGUMENT, as DDPSolver.hpp:41-45 throws). */
  int nmpc_b200_ddp_solve(nmpc_b200_ddp * h,
                          int B,
                          double current_t,
                          const double * x0,
                          const double * u_init,
                          int n_u_steps,
                          int on_device,
                          void * stream);

  typedef enum
  {
    NMPC_B200_DDP_X = 0, /* controlData().x_list     double [B][N+1][NX] */
    NMPC_B200_DDP_U = 1, /* controlData().u_list     double [B][N][NU] */
    NMPC_B200_DDP_COST_LIST = 2, /* controlData().cost_list  double [B][N+1] */
    NMPC_B200_DDP_K_FF = 3, /* k_list_                  double [B][N][NU] */
    NMPC_B200_DDP_K_FB = 4, /* K_list_                  double [B][N][NU*NX] */
    NMPC_B200_DDP_TRACE = 5, /* traceDataList()          double [B][max_iter+1][9]: iter cost lambda dlambda alpha
                                k_rel_norm cost_update_actual cost_update_expected cost_update_ratio */
    NMPC_B200_DDP_STATUS = 6, /* last procOnce retval     int [B]: 1 converged (solve() true), 0 max_iter, -1 failure */
    NMPC_B200_DDP_ITERS = 7, /* traceDataList().back().iter  int [B] */
    NMPC_B200_DDP_N_FORWARD = 8, /* forwardPass() calls      int [B] */
    NMPC_B200_DDP_N_BACKWARD = 9, /* backwardPass() calls     int [B] */
    NMPC_B200_DDP_COST = 10, /* cost_list.sum()          double [B] */
    NMPC_B200_DDP_U0 = 11, /* u_list[0]                double [B][NU] */
    NMPC_B200_DDP_N_TRACE = 12 /* traceDataList().size()   int [B] */
  } nmpc_b200_ddp_field;

  /** controlData()/traceDataList() accessors (DDPSolver.h:288-303): copies field `what` of the last
      solve into dst (dst_bytes must be at least the field size for the last B). */
  int nmpc_b200_ddp_get(nmpc_b200_ddp * h, int what, void * dst, size_t dst_bytes, int dst_on_device, void * stream);

  /** Wait for the handle's pending work. */
  int nmpc_b200_ddp_sync(nmpc_b200_ddp * h);

  /** computationDuration() (DDPSolver.h:219-247, :300-303) measured with CUDA events on the solve
      stream.  Enable before solve(); get() waits for the events.  ms[8] = {solve, setup (layout +
      initial rollout), opt, derivative, backward, forward, copy_in, copy_out}; launches[4] = number
      of launches of {rollout, derivative, backward, forward} kernels in the last solve. */
  int nmpc_b200_ddp_enable_timing(nmpc_b200_ddp * h, int enable);
  int nmpc_b200_ddp_get_durations(nmpc_b200_ddp * h, double * ms, int * launches);
  /** TraceData::duration_derivative / duration_backward / duration_forward of every trace entry of the last solve
      (DDPSolver.h:208-215; the last three columns of dumpTraceDataList(), DDPSolver.hpp:567-596), from the same stage
      events: ms[rows][4] = {derivative, backward, first line-search candidate, other candidates} in ms; row 0 is the
      iter-0 entry (initial rollout in column 2), duration_forward = columns 2 + 3.  The batch runs every stage as one
      launch, so the durations of an iteration are those of the whole batch.  *rows_filled = entries available. */
  int nmpc_b200_ddp_get_iteration_durations(nmpc_b200_ddp * h, double * ms, int rows, int * rows_filled);

  /** Kernel-selection knobs of a handle.  The engine picks, per stage, the kernel variant measured fastest for the
      problem size on the device it runs on (thresholds scale with the device's SM count; DESIGN.md lists the variants
      and the measurements); these calls read a knob or pin it, e.g. to reproduce an experiment or to re-tune for
      another part.  Results do not depend on the choice beyond rounding.  Keys: backward_lanes (0 | 1 | 2),
      backward_lanes_tiles_per_cta, backward_lanes_max_batch, backward_fused (0 | 1), backward_quad (-1 auto | 0 | 1),
      backward_quad_max_batch, backward_group_size (-1 auto | 1 | 4 or 16), backward_coop_max_batch, backward_wide
      (0 | 1), forward_lanes (-1 auto | 1 | 3 phased | 4 | 16), forward_phased_max_batch, forward_split (0 | 1),
      forward_split_max_batch, threads_per_block (-1 auto | 32 .. 128), solve_tile (0 | 1), solve_tile_max_batch.
      An unknown key is NMPC_B200_ERR_INVALID_ARGUMENT.  (Environment variables NMPC_B200_<KEY> preset a knob for
      every handle of a process: a developer switch for the experiment scripts under tools/.) */
  int nmpc_b200_ddp_set_tuning(nmpc_b200_ddp * h, const char * key, int value);
  int nmpc_b200_ddp_get_tuning(nmpc_b200_ddp * h, const char * key, int * value);

  /* ------------------------------------------------------------------- user functors ---- */

  /** Load a shared library that registers problem functors (include/nmpc_b200/plugin.h: a .cu file with
      NMPC_B200_REGISTER_DDP_MODEL / _FMPC_MODEL lines, built against this library).  The reference binds any
      std::shared_ptr<DDPProblem> at run time (DDPSolver.h:255, FmpcSolver.h:296); this is the device-side equivalent:
      after the call the plugin's functors can be named in nmpc_b200_ddp_create / nmpc_b200_fmpc_create.  Loading the
      same path twice is harmless.  NMPC_B200_ERR_RUNTIME with the loader's message when the library cannot be loaded. */
  int nmpc_b200_load_plugin(const char * path);

  /* ------------------------------------------------------------- several GPUs, one box ---- */

  /** A batch of DDPSolver objects SHARDED over several GPUs of one box from ONE process: a solver handle, a stream and
      a host worker thread per device.  Instances never interact (the reference runs one DDPSolver object per problem,
      DDPSolver.h:329-374), so instance b of a solve with B instances lives on shard s with
      begin(s) <= b < end(s), begin(s) = s * (B / n) + min(s, B % n): contiguous chunks, sizes differing by at most one,
      and no collective on the data path.  `devices` = CUDA ordinals (NULL: 0 .. n_devices-1; n_devices <= 0: every
      visible device; an ordinal may appear more than once: two shards then share that device, each with
      its own stream).  `total_capacity` = the largest B of any later solve. */
  typedef struct nmpc_b200_ddp_sharded nmpc_b200_ddp_sharded;
  int nmpc_b200_ddp_create_sharded(const char * model,
                                   const double * params,
                                   int n_params,
                                   const nmpc_b200_ddp_config * cfg,
                                   int total_capacity,
                                   const int * devices,
                                   int n_devices,
                                   nmpc_b200_ddp_sharded ** out);
  int nmpc_b200_ddp_sharded_destroy(nmpc_b200_ddp_sharded * h);
  int nmpc_b200_ddp_sharded_num_shards(const nmpc_b200_ddp_sharded * h);
  /** The single-GPU handle behind shard `shard` (owned by the sharded handle), e.g. for the timing entries. */
  nmpc_b200_ddp * nmpc_b200_ddp_sharded_shard(nmpc_b200_ddp_sharded * h, int shard);
  /** [begin, end) of shard `shard` for a solve with B instances; *device = its CUDA ordinal (any may be NULL). */
  int nmpc_b200_ddp_sharded_range(const nmpc_b200_ddp_sharded * h, int B, int shard, int * begin, int * end, int * device);
  /** DDPSolver::config() / setInputLimitsFunc for every shard. */
  int nmpc_b200_ddp_sharded_set_config(nmpc_b200_ddp_sharded * h, const nmpc_b200_ddp_config * cfg);
  int nmpc_b200_ddp_sharded_set_input_limits(nmpc_b200_ddp_sharded * h, const double * lower, const double * upper);
  /** DDPSolver::solve for B <= total_capacity instances from HOST arrays x0[B][NX], u_init[B][N][NU]: every shard copies
      its chunk in and solves on its own device, all shards at once; returns when all have finished. */
  int nmpc_b200_ddp_sharded_solve(nmpc_b200_ddp_sharded * h,
                                  int B,
                                  double current_t,
                                  const double * x0,
                                  const double * u_init,
                                  int n_u_steps);
  /** nmpc_b200_ddp_get over all shards: field `what` of the last solve, [B][...] in instance order.  dst_device < 0:
      dst is host memory.  dst_device >= 0: dst is memory of that CUDA device and every shard's gather kernel stores its
      rows straight into it (over NVLink for the shards on other devices; peer access is enabled on first use). */
  int nmpc_b200_ddp_sharded_get(nmpc_b200_ddp_sharded * h, int what, void * dst, size_t dst_bytes, int dst_device);

  /** Several PROCESSES (one per GPU, e.g. under torchrun / MPI): a device buffer every process of the box can store
      into.  The owner creates it (cudaMalloc, zero-filled) and hands the 64-byte handle to the other processes by any
      host channel; they open it and pass `ptr + offset` as a device destination to nmpc_b200_ddp_get /
      nmpc_b200_fmpc_get, whose gather kernel then writes its rows directly into the owner's memory -- the first-step
      controls of all shards land in one place without a collective.  peer_signal / peer_wait order it: after its
      stores a process signals flag[rank] <- value (release, system scope) on its stream; the owner's stream waits until
      n flags have reached `value`.  peer_wait gives up after timeout_ms (NMPC_B200_ERR_RUNTIME at the next
      nmpc_b200_peer_check) instead of hanging the device. */
  int nmpc_b200_peer_buffer_create(size_t bytes, int device, void ** ptr, unsigned char handle[64]);
  int nmpc_b200_peer_buffer_open(const unsigned char handle[64], int device, void ** ptr);
  int nmpc_b200_peer_buffer_close(void * ptr, int device);
  int nmpc_b200_peer_buffer_destroy(void * ptr, int device);
  int nmpc_b200_peer_signal(void * flag, unsigned long long value, int device, void * stream);
  int nmpc_b200_peer_wait(void * flags, int n_flags, unsigned long long value, int timeout_ms, int device, void * stream);
  /** Synchronises `stream` and reports whether a peer_wait on flags timed out since the last check (flags[n_flags] is
      the time-out word, so the buffer needs n_flags + 1 words). */
  int nmpc_b200_peer_check(void * flags, int n_flags, int device, void * stream);

  /* ------------------------------------------------------- receding-horizon (MPC) loop ---- */

  /** The MPC loops that call the solvers in the reference (TestDDPBipedal.cpp:243-268,
      TestDDPCartPole.cpp:313-343 + :388-396, TestFmpcOscillator.cpp:166-190), run for a whole batch ON THE DEVICE:
      tick after tick { solve; apply u_list[0] to the plant; warm-start the next solve } with no host round trip.
        plant = 0   current_x <- controlData().x_list[1]                      (TestDDPBipedal.cpp:264)
        plant = 1   current_x <- stateEq(t, current_x, u, sim_dt), n_substeps times per tick, u held
                    (TestDDPCartPole.cpp:330, TestFmpcOscillator.cpp:187); needs a functor with the 4-argument
                    stateEq overload, else NMPC_B200_ERR_UNSUPPORTED
        shift_inputs = 1   initial_u_list <- u_list[1:], last entry repeated  (TestDDPBipedal.cpp:265-267)
        shift_inputs = 0   initial_u_list <- u_list                           (TestDDPCartPole.cpp:395)
        clamp_u0 = 1       the applied input is u_list[0] clamped to the input limits (TestDDPCartPole.cpp:393-394;
                           needs nmpc_b200_ddp_set_input_limits, or _set_input_limits_mpc for limits that depend on time)
      current_t advances by tick_dt per tick. */
  typedef struct
  {
    int n_ticks;
    int plant;
    int shift_inputs;
    int clamp_u0;
    int n_substeps;
    int feedback; /* FMPC only: apply u_list[0] + coeffList().front().K (x_list[0] - current_x) at every plant sub-step
                     (TestFmpcCartPole.cpp:351-356) */
    double tick_dt;
    double sim_dt;
  } nmpc_b200_mpc_config;

  /** Run n_ticks of the loop from (current_t, x0[B][NX], u_init[B][N][NU]).  Logs (any may be NULL; host or device
      according to `on_device`): x_log[B][n_ticks+1][NX] = current_x at every tick and after the last one,
      u_log[B][n_ticks][NU] = the applied input, iters_log[B][n_ticks] = traceDataList().back().iter,
      status_log[B][n_ticks] = last procOnce retval.  Afterwards the handle holds the LAST solve (nmpc_b200_ddp_get). */
  int nmpc_b200_ddp_run_mpc(nmpc_b200_ddp * h,
                            int B,
                            double current_t,
                            const double * x0,
                            const double * u_init,
                            int n_u_steps,
                            const nmpc_b200_mpc_config * mpc,
                            double * x_log,
                            double * u_log,
                            int * iters_log,
                            int * status_log,
                            int on_device,
                            void * stream);

  /* ---------------------------------------------------------------------------- FMPC ---- */

  /** Mirror of FmpcSolver::Configuration (FmpcSolver.h:58-89). */
  typedef struct
  {
    int horizon_steps; /* 100 */
    int max_iter; /* 10 */
    int check_nan; /* 1 */
    int init_complementary_variable; /* 0 */
    int update_barrier_eps; /* 1 */
    int break_if_llt_fails; /* 0 */
    int enable_line_search; /* 0 */
    int merit_const_scale_from_lagrange_multipliers; /* 0 */
    double kkt_error_thre; /* 1e-4 */
    double initial_barrier_eps; /* barrier_eps_ on entry (FmpcSolver.h:413-414): 1e-4 */
  } nmpc_b200_fmpc_config;

  void nmpc_b200_fmpc_config_default(nmpc_b200_fmpc_config * cfg);

  typedef struct nmpc_b200_fmpc nmpc_b200_fmpc;

  int nmpc_b200_fmpc_create(const char * model,
                            const double * params,
                            int n_params,
                            const nmpc_b200_fmpc_config * cfg,
                            int batch_capacity,
                            int device,
                            nmpc_b200_fmpc ** out);
  int nmpc_b200_fmpc_destroy(nmpc_b200_fmpc * h);
  int nmpc_b200_fmpc_set_config(nmpc_b200_fmpc * h, const nmpc_b200_fmpc_config * cfg);

  /** FmpcSolver::solve(current_t, current_x, initial_variable) (FmpcSolver.hpp:158-257): x0[B][NX] and
      the initial Variable as five arrays x[B][N+1][NX], u[B][N][NU], lambda[B][N+1][NX], s[B][N][NG],
      nu[B][N][NG].  Negative s/nu => NMPC_B200_ERR_RUNTIME (checkVariable, FmpcSolver.hpp:348-361). */
  int nmpc_b200_fmpc_solve(nmpc_b200_fmpc * h,
                           int B,
                           double current_t,
                           const double * x0,
                           const double * x,
                           const double * u,
                           const double * lambda,
                           const double * s,
                           const double * nu,
                           int n_steps,
                           int on_device,
                           void * stream);

  typedef enum
  {
    NMPC_B200_FMPC_X = 0, /* variable().x_list       double [B][N+1][NX] */
    NMPC_B200_FMPC_U = 1, /* variable().u_list       double [B][N][NU] */
    NMPC_B200_FMPC_LAMBDA = 2, /* variable().lambda_list  double [B][N+1][NX] */
    NMPC_B200_FMPC_S = 3, /* variable().s_list       double [B][N][NG] */
    NMPC_B200_FMPC_NU = 4, /* variable().nu_list      double [B][N][NG] */
    NMPC_B200_FMPC_K_FF = 5, /* coeffList()[i].k        double [B][N][NU] */
    NMPC_B200_FMPC_K_FB = 6, /* coeffList()[i].K        double [B][N][NU*NX] */
    NMPC_B200_FMPC_TRACE = 7, /* traceDataList()         double [B][max_iter][5]: iter kkt_error barrier_eps
                                 alpha_s alpha_nu */
    NMPC_B200_FMPC_STATUS = 8, /* FmpcSolver::Status      int [B] (FmpcSolver.h:92-114) */
    NMPC_B200_FMPC_N_TRACE = 9, /* traceDataList().size()  int [B] */
    NMPC_B200_FMPC_U0 = 10 /* u_list[0]               double [B][NU] */
  } nmpc_b200_fmpc_field;

  /** The FMPC loops of the reference (TestFmpcOscillator.cpp:166-190, TestFmpcCartPole.cpp:344-357 + :405-412) on
      the device: every tick { solve(current_t, current_x, variable); u <- variable().u_list[0]; current_x <- plant;
      variable <- fmpc_solver->variable() }.  barrier_eps_ persists from tick to tick as the member does
      (FmpcSolver.h:413-414).  mpc->shift_inputs and mpc->clamp_u0 must be 0.  Logs (may be NULL): x_log[B][n_ticks+1][NX],
      u_log[B][n_ticks][NU] (u_list[0]), kkt_log[B][n_ticks] (traceDataList().back().kkt_error), status_log[B][n_ticks]. */
  int nmpc_b200_fmpc_run_mpc(nmpc_b200_fmpc * h,
                             int B,
                             double current_t,
                             const double * x0,
                             const double * x,
                             const double * u,
                             const double * lambda,
                             const double * s,
                             const double * nu,
                             int n_steps,
                             const nmpc_b200_mpc_config * mpc,
                             double * x_log,
                             double * u_log,
                             double * kkt_log,
                             int * status_log,
                             int on_device,
                             void * stream);

  int nmpc_b200_fmpc_get(nmpc_b200_fmpc * h, int what, void * dst, size_t dst_bytes, int dst_on_device, void * stream);
  int nmpc_b200_fmpc_sync(nmpc_b200_fmpc * h);
  int nmpc_b200_fmpc_enable_timing(nmpc_b200_fmpc * h, int enable);
  /** ms[8] = {solve, setup, opt, coeff, backward, forward, update, copy}; launches[4] = {coeff,
      backward, forward, update}. */
  int nmpc_b200_fmpc_get_durations(nmpc_b200_fmpc * h, double * ms, int * launches);

  /** nmpc_b200_ddp_create_sharded ("several GPUs, one box") for FmpcSolver: x0[B][NX] and the five Variable arrays as in nmpc_b200_fmpc_solve, all host memory; fields
      as in nmpc_b200_fmpc_field. */
  typedef struct nmpc_b200_fmpc_sharded nmpc_b200_fmpc_sharded;
  int nmpc_b200_fmpc_create_sharded(const char * model,
                                    const double * params,
                                    int n_params,
                                    const nmpc_b200_fmpc_config * cfg,
                                    int total_capacity,
                                    const int * devices,
                                    int n_devices,
                                    nmpc_b200_fmpc_sharded ** out);
  int nmpc_b200_fmpc_sharded_destroy(nmpc_b200_fmpc_sharded * h);
  int nmpc_b200_fmpc_sharded_num_shards(const nmpc_b200_fmpc_sharded * h);
  int nmpc_b200_fmpc_sharded_set_config(nmpc_b200_fmpc_sharded * h, const nmpc_b200_fmpc_config * cfg);
  int nmpc_b200_fmpc_sharded_solve(nmpc_b200_fmpc_sharded * h,
                                   int B,
                                   double current_t,
                                   const double * x0,
                                   const double * x,
                                   const double * u,
                                   const double * lambda,
                                   const double * s,
                                   const double * nu,
                                   int n_steps);
  int nmpc_b200_fmpc_sharded_get(nmpc_b200_fmpc_sharded * h, int what, void * dst, size_t dst_bytes, int dst_device);

#ifdef __cplusplus
}
#endif

#endif /* NMPC_B200_C_API_H */
