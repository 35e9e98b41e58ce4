/*
 * sepaihrd_b200.h -- C ABI of the B200-native batched SEPAIHRD likelihood evaluator.
 *
 * This is the drop-in boundary for ONE hot path of
 * adjo0043/Mathematical-Modeling-Of-Infectious-Diseases-V1:
 *
 *     parameters -> AgeSEPAIHRDModel RHS -> adaptive Dopri5 (Boost.Odeint semantics)
 *                -> daily incidence -> Poisson log-likelihood
 *
 * evaluated for B parameter vectors per call on one GPU.  The reference has no C ABI today:
 * its boundary is the C++ virtual interface
 *     double IObjectiveFunction::calculate(const Eigen::VectorXd&) const
 *         (reference include/sir_age_structured/interfaces/IObjectiveFunction.hpp:24)
 * implemented by SEPAIHRDObjectiveFunction::calculate
 *         (reference src/model/objectives/SEPAIHRDObjectiveFunction.cpp:62-235).
 * Every entry point below states which reference function(s) it replaces.  INTEGRATION.md
 * shows the adapter a reference maintainer would add on their side.
 *
 * Conventions
 *   - plain pointers and sizes only; no exceptions cross this boundary; every call returns a
 *     sepaihrd_rc (0 = OK).  sepaihrd_last_error() gives a thread-local message.
 *   - per-set failures are NOT call failures: the set's logL is -DBL_MAX
 *     (std::numeric_limits<double>::lowest(), SEPAIHRDObjectiveFunction.cpp:111,119,161,227) and
 *     the reason is in out_status[b] (SEPAIHRD_ST_* bits).
 *   - all floating point is IEEE binary64.
 *   - there is NO CPU fallback: if no CUDA device is usable, sepaihrd_create fails.
 *   - a ctx may be used from several host threads: every entry point takes the ctx's lock for its duration.  The
 *     `_device` variants only ENQUEUE on the ctx stream; ordering between threads that enqueue is the callers' business.
 *   - streams: evaluation / simulation launches of one ctx may be in flight on DIFFERENT streams (sepaihrd_set_stream between
 *     calls): every launch owns its work counter.  The host-pointer entry points, the posterior-predictive pass, the swarm
 *     and the sampler reuse per-ctx staging / work buffers and therefore assume ONE stream at a time; buffers that have to
 *     grow are replaced only after the whole device has drained.
 */
#ifndef SEPAIHRD_B200_H
#define SEPAIHRD_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEPAIHRD_ABI_VERSION 1

/* Number of compartments: S,E,P,A,I,H,ICU,R,D,CumH,CumICU
 * (reference include/model/ModelConstants.hpp:18). State layout is compartment-major:
 * state[c*n_ages + age]  (reference src/model/AgeSEPAIHRDModel.cpp:116-134). */
#define SEPAIHRD_NUM_COMPARTMENTS 11
#define SEPAIHRD_MAX_AGES 16
#define SEPAIHRD_MAX_SEGMENTS 32

/* ---- return codes ------------------------------------------------------------------------ */
typedef enum sepaihrd_rc {
    SEPAIHRD_OK = 0,
    SEPAIHRD_ERR_INVALID_ARGUMENT = 1, /* reference: InvalidParameterException */
    SEPAIHRD_ERR_NO_DEVICE = 2,        /* CUDA extension / device missing: fail loudly */
    SEPAIHRD_ERR_CUDA = 3,
    SEPAIHRD_ERR_UNSUPPORTED = 4,
    SEPAIHRD_ERR_OUT_OF_MEMORY = 5
} sepaihrd_rc;

/* ---- per-set status bits (out_status) ------------------------------------------------------ */
#define SEPAIHRD_ST_OK             0u
#define SEPAIHRD_ST_S_OVERFLOW     1u  /* sum of non-S compartments > N_i: calculate() returns lowest() (ObjectiveFunction.cpp:161) */
#define SEPAIHRD_ST_STEP_FAILURE   2u  /* >500 consecutive rejected steps: Boost failed_step_checker; reference throws SimulationException (Dopri5SolverStrategy.cpp:38-42) */
#define SEPAIHRD_ST_NONFINITE      4u  /* total logL NaN/Inf -> lowest() (ObjectiveFunction.cpp:227) */
#define SEPAIHRD_ST_INVALID_PARAM  8u  /* negative kappa: setCalibratableValues throws, caught at ObjectiveFunction.cpp:117-122 -> lowest() */

/* ---- parameter-slot layout ----------------------------------------------------------------
 * A "slot vector" is the flat image of reference struct SEPAIHRDParameters
 * (include/model/parameters/SEPAIHRDParameters.hpp) restricted to what the hot path reads.
 * For n = n_ages, nb = n_beta, nk = n_kappa the slots are, in order:
 *   beta_values[nb] | kappa_values[nk] (kappa_values[0] = fixed baseline kappa_1) |
 *   theta sigma gamma_p gamma_A gamma_I gamma_H gamma_ICU |
 *   a[n] h_infec[n] p[n] h[n] icu[n] d_H[n] d_ICU[n] d_community[n] |
 *   E0 P0 A0 I0 H0 ICU0 R0 D0 multipliers | seed_exposed runup_days | beta (scalar, quirk Q1)
 */
int32_t sepaihrd_slot_count(int32_t n_ages, int32_t n_beta, int32_t n_kappa);

/* Resolve a reference parameter name ("beta_3", "kappa_2", "h_infec_1", "gamma_ICU", ...) to
 * its slot index, with the same prefix-dispatch order as
 * SEPAIHRDParameterManager::updateModelParameters (src/model/parameters/SEPAIHRDParameterManager.cpp:197-267).
 * Returns -1 for an unknown name, -2 for a name that the reference rejects at construction
 * (kappa_1 / kappa_baseline with a fixed baseline, out-of-range index; .cpp:45-88). */
int32_t sepaihrd_slot_for_name(int32_t n_ages, int32_t n_beta, int32_t n_kappa, const char* name);

/* ---- problem description (everything that is constant across parameter sets) ------------- */
typedef struct sepaihrd_problem {
    int32_t abi_version;          /* = SEPAIHRD_ABI_VERSION */
    int32_t n_ages;               /* n: 1..16.  4 (Spain-2020) and 16 (synthetic variant) run natively; any other count runs
                                   * zero-padded to the next of the two (empty classes: population 0, no contacts, skipped
                                   * observations), with every input and output in the caller's n                          */
    int32_t n_times;              /* K output times, strictly increasing (Simulator.cpp:82-90)          */
    int32_t n_obs;                /* rows of each observation matrix; must equal K - runup_offset       */
    const double* times;          /* [K]   (src/model/main.cpp:244-253: integers -int(runup)..num_days-1) */
    const double* obs_hosp;       /* [n_obs*n] row-major (day, age): CalibrationData::getNewHospitalizations */
    const double* obs_icu;        /* [n_obs*n] getNewICU                                                */
    const double* obs_deaths;     /* [n_obs*n] getNewDeaths                                             */
    const double* population;     /* [n]  N                                                             */
    const double* contact_matrix; /* [n*n] COLUMN-major M(i,j) = data[j*n+i] (Eigen default; AgeSEPAIHRDModel.cpp:145,168) */
    int32_t n_beta;               /* length of beta_end_times / beta_values (element 0 = baseline period) */
    int32_t n_kappa;              /* length of kappa_end_times / kappa_values (element 0 = fixed baseline) */
    const double* beta_end_times; /* [n_beta]  (initial_guess.txt:6)                                    */
    const double* kappa_end_times;/* [n_kappa] (initial_guess.txt:7)                                    */
    const double* base_slots;     /* [sepaihrd_slot_count] values of every slot for non-calibrated parameters */
    const double* data_initial_state; /* [11*n] CalibrationData::getInitialSEPAIHRDState (multiplier mode, a6) */
    int32_t n_params;             /* P: length of each parameter vector handed to eval                  */
    int32_t constraint_mode;      /* 0 = OPTIMIZATION_CLAMP, 1 = MCMC_REFLECT (SEPAIHRDParameterManager.hpp:22-25) */
    const int32_t* param_slot;    /* [P] slot of calibrated parameter i (from sepaihrd_slot_for_name)   */
    const double* lower_bound;    /* [P] (param_bounds.txt); NaN = parameter has no bounds entry        */
    const double* upper_bound;    /* [P]                                                                */
    double abs_tol;               /* 1e-6 (main.cpp:260)                                                */
    double rel_tol;               /* 1e-6 (main.cpp:261)                                                */
    double dt_hint;               /* 1.0: objective builds its simulator with time_step 1.0 (ObjectiveFunction.cpp:113) */
} sepaihrd_problem;

typedef struct sepaihrd_ctx sepaihrd_ctx; /* opaque: owns device copies of the problem, a stream, scratch */

/* Arithmetic flavour of the kernels.
 *   SEPAIHRD_MATH_FAST   : FMA-contracted, likelihood accumulated per lane (default; production)
 *   SEPAIHRD_MATH_STRICT : unfused IEEE mul/add in the reference's source order, likelihood summed
 *                          row-by-row like calculateSingleLogLikelihood; bit-comparable with the
 *                          CPU oracle up to libm (log/pow) differences. */
#define SEPAIHRD_MATH_FAST   0
#define SEPAIHRD_MATH_STRICT 1
/* FAST arithmetic with the general kernel build even when every breakpoint sits on an output-grid point (the default then
 * picks the build without the mixed-segment attempt body): a verification switch, results are bit-identical to FAST. */
#define SEPAIHRD_MATH_FAST_GENERAL 2
/* EXPERIMENTAL builds only (-DSEPAIHRD_WITH_SPLIT; the shipped library answers SEPAIHRD_ERR_UNSUPPORTED): FAST arithmetic on
 * the warp-pair kernel (csrc/experiments/sepaihrd_split.cuh: every set split over an upstream warp S E P A I and a downstream
 * warp H ICU R D CumH CumICU).  Bit-identical to FAST but measured SLOWER at every batch size
 * (profiles/r02_split_kernel_experiment.txt), so it is not part of the product. */
#define SEPAIHRD_MATH_FAST_SPLIT 3

/* Replaces: construction of AgeSEPAIHRDModel + PiecewiseConstantNpiStrategy + SEPAIHRDParameterManager
 * + SEPAIHRDObjectiveFunction + AgeSEPAIHRDSimulator + Dopri5SolverStrategy
 * (src/model/SEPAIHRDModelCalibration.cpp:73-132; ObjectiveFunction.cpp:22-50).
 * device < 0 means "current CUDA device". The problem is deep-copied. */
sepaihrd_rc sepaihrd_create(const sepaihrd_problem* problem, int32_t device, sepaihrd_ctx** out_ctx);
void        sepaihrd_destroy(sepaihrd_ctx* ctx);

/* Replaces SEPAIHRDParameterManager::setConstraintMode (SEPAIHRDParameterManager.hpp:138). */
sepaihrd_rc sepaihrd_set_constraint_mode(sepaihrd_ctx* ctx, int32_t mode);
sepaihrd_rc sepaihrd_set_math_mode(sepaihrd_ctx* ctx, int32_t mode);
/* Work on a caller-provided CUDA stream (cudaStream_t cast to void*).  NULL is the CUDA legacy default
 * stream (what torch.cuda.current_stream().cuda_stream returns by default); SEPAIHRD_STREAM_OWN selects
 * the non-blocking stream the ctx created for itself (the initial setting). */
#define SEPAIHRD_STREAM_OWN ((void*)(intptr_t)-1)
sepaihrd_rc sepaihrd_set_stream(sepaihrd_ctx* ctx, void* cuda_stream);

/* Replaces B calls of SEPAIHRDObjectiveFunction::calculate (ObjectiveFunction.cpp:62-235) with a
 * NullSimulationCache (benchmark_main.cpp:229-238).  HOST buffers:
 *   params     [B][ld] row-major, ld >= P: one Eigen::VectorXd per row, unconstrained (constraints
 *              are applied on the device like updateModelParameters -> applyConstraints, .cpp:173)
 *   out_ll     [B]   log-likelihood or -DBL_MAX
 *   out_status [B]   SEPAIHRD_ST_* bits, may be NULL
 *   out_steps  [B][2] accepted / rejected Dopri5 step attempts, may be NULL (parity diagnostics)
 * Copies H2D, runs the fused kernel, copies D2H, synchronises.
 * Thread-safe like calculate() has to be (the reference calls it from OpenMP loops, ParticleSwarmOptimizer.cpp:368-424,
 * HillClimbingOptimizer.cpp:228-234), and concurrent calls of at most 4096 sets each are MERGED: a caller that finds no
 * launch in flight takes every request queued so far (its own included) to the device as one launch.  The reference's
 * optimizers, unchanged, therefore cost one launch per round of their threads, not one per calculate().
 * Requests of at most 4096 sets are latency, not throughput: one stream, one synchronisation, and pageable buffers (an
 * Eigen::VectorXd, a std::vector) staged through a page-locked buffer of the ctx, so page-locked caller buffers
 * (sepaihrd_alloc_pinned) save one small copy, no more.  Larger requests overlap their copies with the kernel in growing
 * chunks and want page-locked memory.  A launch of up to ~4 000 sets costs what one set costs (~0.5 ms): callers with a
 * sequential loop over calculate() should look ahead (host/optimizers.hpp: the one-chain sampler, the line search). */
sepaihrd_rc sepaihrd_eval_batch(sepaihrd_ctx* ctx, const double* params, int64_t B, int64_t ld,
                                double* out_ll, uint32_t* out_status, int32_t* out_steps);

/* Same, with DEVICE pointers and no synchronisation: work is enqueued on the ctx stream. */
sepaihrd_rc sepaihrd_eval_batch_device(sepaihrd_ctx* ctx, const double* d_params, int64_t B, int64_t ld,
                                       double* d_out_ll, uint32_t* d_out_status, int32_t* d_out_steps);

/* Trajectory output selector for sepaihrd_simulate_batch. */
#define SEPAIHRD_TRAJ_FULL      0  /* all 11*n state components per output time                     */
#define SEPAIHRD_TRAJ_OBSERVED  1  /* D, CumH, CumICU only (3*n per output time), the streams that
                                      SimulationResultProcessor::getCompartmentData extracts
                                      (ObjectiveFunction.cpp:172-174)                                */

/* Replaces B calls of AgeSEPAIHRDSimulator::run / Simulator::run (src/sir_age_structured/Simulator.cpp:60-150)
 * on models updated by updateModelParameters, with the initial-state rule of calculate()
 * (ObjectiveFunction.cpp:124-163).  HOST buffers.
 *   what       SEPAIHRD_TRAJ_*
 *   stride     keep every stride-th output time starting at index 0 (1 = all K times)
 *   out        [B][ceil(K/stride)][W] with W = 11*n or 3*n; rows of failed sets are NaN-filled
 *   out_status [B] may be NULL */
sepaihrd_rc sepaihrd_simulate_batch(sepaihrd_ctx* ctx, const double* params, int64_t B, int64_t ld,
                                    int32_t what, int32_t stride, double* out, uint32_t* out_status);
sepaihrd_rc sepaihrd_simulate_batch_device(sepaihrd_ctx* ctx, const double* d_params, int64_t B, int64_t ld,
                                           int32_t what, int32_t stride, double* d_out, uint32_t* d_out_status);

/* Replaces B calls of Simulator::run(initial_state, times) (Simulator.cpp:60-150) where the CALLER supplies the
 * initial state, integrated as given (no seeding / multiplier / S-remainder rule): what
 * PostCalibrationAnalyser and ResultAggregator do with one fixed initial state for every posterior draw
 * (PostCalibrationAnalyser.cpp:156,221; ResultAggregator.cpp:289 -- quirk Q9).
 *   initial_states [B][state_stride] compartment-major states, or ONE state shared by all sets when
 *                  state_stride == 0.  Other arguments as sepaihrd_simulate_batch. */
sepaihrd_rc sepaihrd_simulate_from_state(sepaihrd_ctx* ctx, const double* params, int64_t B, int64_t ld,
                                         const double* initial_states, int64_t state_stride, int32_t what, int32_t stride,
                                         double* out, uint32_t* out_status);
sepaihrd_rc sepaihrd_simulate_from_state_device(sepaihrd_ctx* ctx, const double* d_params, int64_t B, int64_t ld,
                                                const double* d_initial_states, int64_t state_stride, int32_t what,
                                                int32_t stride, double* d_out, uint32_t* d_out_status);

/* Replaces ResultAggregator::aggregatePosteriorPredictives (src/model/ResultAggregator.cpp:174-412): B posterior draws are
 * simulated from ONE fixed initial state (quirk Q9) and reduced on the device to quantiles of six series on the output
 * days t >= 0 (T of them): daily hospitalisations, ICU admissions, deaths (first differences, clamped at 0, .cpp:292-335)
 * and their running sums (.cpp:337-351).
 *   probs          [n_probs] in [0, 1]  (the reference uses 0.025, 0.05, 0.5, 0.95, 0.975, .cpp:233)
 *   out_quantiles  [6][T][n_ages][n_probs]; NaN where no valid draw exists
 *   out_valid_draws (optional) draws whose simulation succeeded (failed ones are skipped like .cpp:290)
 * Difference from the reference (documented in DESIGN.md): EXACT sample quantiles with linear interpolation between order
 * statistics (selected per column on the device; read off a full sort for more than 8 probabilities), not Boost's
 * order-dependent extended P-square estimate. */
sepaihrd_rc sepaihrd_posterior_predictive(sepaihrd_ctx* ctx, const double* params, int64_t B, int64_t ld,
                                          const double* initial_state, int32_t n_probs, const double* probs,
                                          double* out_quantiles, int64_t* out_valid_draws);

/* ---- device-resident particle swarm ---------------------------------------------------------------------------------
 * Replaces the per-iteration host work of ParticleSwarmOptimization for swarms scored by this evaluator
 * (src/model/optimizers/ParticleSwarmOptimizer.cpp: initializeSwarm :249-328, updateParticles :330-425, standardPSOUpdate
 * :576-618, personal / global best :301-303, :417-421, :149-156).  Positions, velocities and personal bests stay in HBM;
 * per iteration the caller supplies one 32-bit seed per particle of the WHOLE swarm (the reference draws them from its
 * master std::mt19937, :365-371) and the global best position, and reads back the shard's best (value, index, position).
 * The device runs std::mt19937 + std::uniform_real_distribution<double> and the update in unfused FP64, so the swarm
 * visits exactly the positions the host implementation visits.  Bounds are the ctx's parameter bounds (all finite).
 * This process owns particles [particle_offset, particle_offset + local_count) of a swarm of swarm_size.               */
typedef struct sepaihrd_swarm sepaihrd_swarm;
enum { SEPAIHRD_SWARM_POSITIONS = 0, SEPAIHRD_SWARM_VELOCITIES = 1, SEPAIHRD_SWARM_PERSONAL_BEST = 2,
       SEPAIHRD_SWARM_PERSONAL_BEST_VALUES = 3, SEPAIHRD_SWARM_FITNESS = 4 };
sepaihrd_rc sepaihrd_swarm_create(sepaihrd_ctx* ctx, int64_t swarm_size, int64_t particle_offset, int64_t local_count,
                                  sepaihrd_swarm** out);
void sepaihrd_swarm_destroy(sepaihrd_swarm* swarm);
/* seeds [swarm_size]; initial (or NULL): global particle 0 starts at clamp(initial) instead of a uniform draw (:283-288) */
sepaihrd_rc sepaihrd_swarm_init(sepaihrd_swarm* swarm, const uint32_t* seeds, const double* initial);
/* One objective launch over the shard's positions, personal-best update, arg-max of the personal bests (first maximum
 * wins).  out_best_local_index = -1 (and value -inf) for an empty shard.  out_best_position [P] may be NULL.            */
sepaihrd_rc sepaihrd_swarm_evaluate(sepaihrd_swarm* swarm, double* out_best_value, int64_t* out_best_local_index,
                                    double* out_best_position);
/* Velocity / position update with inertia omega and acceleration coefficients c1, c2 towards global_best [P].          */
sepaihrd_rc sepaihrd_swarm_step(sepaihrd_swarm* swarm, const uint32_t* seeds, double omega, double c1, double c2,
                                const double* global_best);
/* Copy one of the swarm's arrays to the host: [local][P] or [local] (SEPAIHRD_SWARM_*).                                */
sepaihrd_rc sepaihrd_swarm_read(sepaihrd_swarm* swarm, int32_t what, double* out);

/* Fully asynchronous form of the swarm iteration (no host synchronisation, no host<->device copy per iteration): the seeds of
 * every iteration are uploaded once, the global best stays on the device, and the per-rank best travels as ONE record
 * [value, global particle index, position[P]] (2 + P doubles) that the caller all-gathers between the ranks -- with
 * sepaihrd_exchange_all_gather below (NVLink peer memory) or any collective on device buffers (ncclAllGather).
 *   upload_seeds     seeds [n_sets][swarm_size]: set 0 is consumed by sepaihrd_swarm_init_async, set 1 + it by step `it`
 *   init_async       sepaihrd_swarm_init from seed set 0 (initial as for sepaihrd_swarm_init, host pointer or NULL)
 *   evaluate_async   objective launch + personal bests + the shard's best record, all enqueued on the ctx stream
 *   record_device    device pointer of this shard's record (valid until sepaihrd_swarm_destroy)
 *   adopt_global_best  arg-max over `n_records` records (device, [n_records][record_stride]; ties: the lowest particle index,
 *                    like a serial scan of the whole swarm) replaces the device-resident global best when STRICTLY better
 *                    (ParticleSwarmOptimizer.cpp:149-156); the resulting best value is stored in trace[trace_slot]
 *   step_async       sepaihrd_swarm_step with seed set 1 + iteration and the device-resident global best
 *   read_trace       copies trace[0 .. n) to the host (synchronises the ctx stream)
 *   read_global_best copies the device-resident global best (value, position[P]) to the host (synchronises)            */
#define SEPAIHRD_SWARM_TRACE_CAPACITY 4096
sepaihrd_rc sepaihrd_swarm_upload_seeds(sepaihrd_swarm* swarm, const uint32_t* seeds, int64_t n_sets);
sepaihrd_rc sepaihrd_swarm_init_async(sepaihrd_swarm* swarm, const double* initial);
sepaihrd_rc sepaihrd_swarm_evaluate_async(sepaihrd_swarm* swarm);
sepaihrd_rc sepaihrd_swarm_record_device(sepaihrd_swarm* swarm, const double** d_record, int32_t* record_doubles);
sepaihrd_rc sepaihrd_swarm_adopt_global_best(sepaihrd_swarm* swarm, const double* d_records, int32_t n_records,
                                             int64_t record_stride, int32_t trace_slot);
sepaihrd_rc sepaihrd_swarm_step_async(sepaihrd_swarm* swarm, int32_t iteration, double omega, double c1, double c2);
sepaihrd_rc sepaihrd_swarm_read_trace(sepaihrd_swarm* swarm, double* out, int32_t n);
sepaihrd_rc sepaihrd_swarm_read_global_best(sepaihrd_swarm* swarm, double* out_value, double* out_position);

/* ---- device-resident Metropolis-Hastings chains ----------------------------------------------------------------------
 * Replaces the per-iteration host work of MetropolisHastingsSampler (src/sir_age_structured/optimizers/
 * MetropolisHastingsSampler.cpp: generateProposal :91-102, the accept test :318-330, adaptGlobalScale :104-152) for
 * `local_count` of `n_chains` independent chains scored by this evaluator: chain states, generators, scales and accept
 * history stay in HBM; one iteration = propose kernel -> fused likelihood kernel -> accept kernel on the ctx stream, no
 * host work.  Chain c (GLOBAL index) draws from std::mt19937(std::seed_seq{seed, c}); arithmetic, generator and
 * distributions reproduce the host sampler of this repository bit for bit (host/optimizers.cpp; log / exp of both are
 * csrc/det_math.h), so a sharded device run makes the accept decisions of the single-process host run.
 * Scope: the fixed-kernel phase (iterations - 1 <= burn_in); the covariance adaptation after burn-in (:154-199) is the host
 * sampler's.  The ctx's constraint mode applies to the proposals (MCMC_REFLECT for the reference's sampler, :207-210).   */
#define SEPAIHRD_MH_MAX_PARAMS 128
typedef struct sepaihrd_mh sepaihrd_mh;
typedef struct sepaihrd_mh_settings {
    int32_t iterations;             /* mcmc_iterations: the loop runs t = 1 .. iterations - 1 (.cpp:283)            */
    int32_t burn_in;                /* must be >= iterations - 1 here                                              */
    int32_t adapt_scale;            /* 0 / 1: Robbins-Monro global scale (.cpp:104-152)                            */
    int32_t record_accepts;         /* 0 / 1: keep the accept decision of every (iteration, chain) for parity checks */
    double target_acceptance_rate;  /* 0.234                                                                       */
} sepaihrd_mh_settings;
enum { SEPAIHRD_MH_POSITIONS = 0, SEPAIHRD_MH_LOGPOST = 1, SEPAIHRD_MH_SCALES = 2, SEPAIHRD_MH_ACCEPTED_COUNTS = 3 /* int64 */,
       SEPAIHRD_MH_BEST_LOGPOST = 4, SEPAIHRD_MH_BEST_POSITIONS = 5, SEPAIHRD_MH_ACCEPT_MATRIX = 6 /* uint8 [iterations done][local] */,
       SEPAIHRD_MH_TRACE = 7 /* [iterations + 1] */, SEPAIHRD_MH_PROPOSALS = 8,
       SEPAIHRD_MH_FAULT = 9 /* uint32: nonzero if a proposal ever ran out of its 128 polar attempts (never in practice) */ };
sepaihrd_rc sepaihrd_mh_create(sepaihrd_ctx* ctx, int64_t n_chains, int64_t chain_offset, int64_t local_count,
                               const sepaihrd_mh_settings* settings, sepaihrd_mh** out);
void sepaihrd_mh_destroy(sepaihrd_mh* mh);
/* All chains start at initial [P] (its log-posterior is evaluated once on the device); chol_lower [P*P] column-major is the
 * lower Cholesky factor of the proposal covariance (MetropolisHastingsSampler.cpp:216-240).  Enqueues; host buffers are
 * consumed before the call returns. */
sepaihrd_rc sepaihrd_mh_begin(sepaihrd_mh* mh, uint32_t seed, const double* initial, const double* chol_lower);
/* Enqueue up to n_iterations iterations on the ctx stream (stops at settings.iterations). */
sepaihrd_rc sepaihrd_mh_iterate(sepaihrd_mh* mh, int32_t n_iterations);
/* The three phases of ONE iteration as separate calls (propose kernel / fused likelihood kernel / accept kernel; accept
 * advances the iteration index), for callers that time them with CUDA events between the calls. */
sepaihrd_rc sepaihrd_mh_propose(sepaihrd_mh* mh);
sepaihrd_rc sepaihrd_mh_evaluate(sepaihrd_mh* mh);
sepaihrd_rc sepaihrd_mh_accept(sepaihrd_mh* mh);
int32_t sepaihrd_mh_iteration(const sepaihrd_mh* mh);    /* next iteration index, 1-based */
/* Device pointer of the local chains' current log-posteriors [local_count]: the block a rank contributes to the
 * per-iteration all-gather (north_star: "gather log-likelihoods for the MCMC accept step"). */
sepaihrd_rc sepaihrd_mh_logpost_device(sepaihrd_mh* mh, const double** d_logpost);
/* trace[trace_slot] = max over the gathered log-posteriors of all chains; d_all_logpost is [world][block_stride] with rank
 * r's n_chains/world (+1 for the first n_chains % world ranks) values at the start of its block.  Enqueues. */
sepaihrd_rc sepaihrd_mh_note_gathered(sepaihrd_mh* mh, const double* d_all_logpost, int32_t world, int64_t block_stride,
                                      int32_t trace_slot);
/* Look-ahead windows: K iterations of every chain per likelihood launch.  A launch costs the same from 1 to ~4 000 parameter
 * sets, and while a chain rejects it does not move, so the proposals of its next K iterations are known in advance (the
 * generator's next normals; a rejected iteration has consumed its uniform, MetropolisHastingsSampler.cpp:323-329; the scale
 * after j more rejections, :104-152).  _propose draws them for every local chain from a copy of its generator, _evaluate scores
 * local_count x K proposals in ONE launch, _commit replays the sequential loop of every chain up to and including its first
 * accepted proposal and discards the rest.  Decisions, states, scales and generator positions are those of the
 * one-iteration-per-launch phases above; chains advance by different amounts, so each keeps its own iteration index, and a run
 * uses EITHER the windows OR the one-iteration phases.  1 <= K <= 64; the first window of a run fixes the largest K, unless
 * _reserve has allocated the window buffers for a larger one beforehand (it keeps the allocation out of the run).
 * _commit also fills the rank's record for the per-window exchange: [0, local_count) the chains' current log-posteriors,
 * [record_stride] the smallest next-iteration index among the local chains as a double (settings.iterations once all are
 * done); record_stride = local_count or local_count + 1 (the largest shard of the run; 0 = local_count).
 * _progress synchronises and returns that index (also what sepaihrd_mh_iteration reports afterwards). */
sepaihrd_rc sepaihrd_mh_window_reserve(sepaihrd_mh* mh, int32_t K);
sepaihrd_rc sepaihrd_mh_window_propose(sepaihrd_mh* mh, int32_t K);
sepaihrd_rc sepaihrd_mh_window_evaluate(sepaihrd_mh* mh);
sepaihrd_rc sepaihrd_mh_window_commit(sepaihrd_mh* mh, int64_t record_stride);
sepaihrd_rc sepaihrd_mh_window_record(sepaihrd_mh* mh, const double** d_record);
sepaihrd_rc sepaihrd_mh_window_progress(sepaihrd_mh* mh, int32_t* out_min_iteration);
/* Copy one of the sampler's arrays to the host (SEPAIHRD_MH_*); synchronises the ctx stream. */
sepaihrd_rc sepaihrd_mh_read(sepaihrd_mh* mh, int32_t what, void* out);

/* ---- small-record all-gather between the GPUs of one node over NVLink peer memory --------------------------------------
 * The callers of the hot path exchange KiB-scale records once per iteration: the ranks' log-likelihood blocks of a
 * multi-chain Metropolis-Hastings run (MetropolisHastingsSampler.cpp:312-330), the per-rank best of a particle swarm
 * (ParticleSwarmOptimizer.cpp:149-156, 417-421).  One process per GPU; every rank owns a mailbox in its HBM that its peers
 * map through CUDA IPC.  sepaihrd_exchange_all_gather is ONE kernel on the ctx stream: block p stores this rank's record into
 * peer p's mailbox (NVLink stores, then a system-scope release of a sequence flag), then waits (acquire, bounded spin) for
 * rank p's record in its own mailbox and copies it out -- no host round trip, no NCCL launch latency.
 *   create   allocates the mailbox (records of at most max_doubles doubles) and returns its 64-byte IPC handle
 *   connect  handles [world][SEPAIHRD_EXCHANGE_HANDLE_BYTES] of all ranks (gathered by the caller through any channel,
 *            e.g. torch.distributed.all_gather_object); rank r's own entry is ignored
 *   all_gather  d_src [count] -> d_dst [world][count] on every rank (device pointers); all ranks must make the same
 *            sequence of calls.  A peer that does not show up within timeout (default 10 s) raises the status flag
 *            instead of hanging the GPU.
 *   status   0 = fine; nonzero = a wait timed out (synchronises the ctx stream)                                          */
#define SEPAIHRD_EXCHANGE_HANDLE_BYTES 64
typedef struct sepaihrd_exchange sepaihrd_exchange;
sepaihrd_rc sepaihrd_exchange_create(sepaihrd_ctx* ctx, int32_t world, int32_t rank, int64_t max_doubles,
                                     sepaihrd_exchange** out, unsigned char* out_handle);
sepaihrd_rc sepaihrd_exchange_connect(sepaihrd_exchange* ex, const unsigned char* handles);
sepaihrd_rc sepaihrd_exchange_all_gather(sepaihrd_exchange* ex, const double* d_src, int64_t count, double* d_dst);
sepaihrd_rc sepaihrd_exchange_status(sepaihrd_exchange* ex, int32_t* out_status);
void sepaihrd_exchange_destroy(sepaihrd_exchange* ex);

/* ---- ordering pass in front of large launches (csrc/sepaihrd_order.cu) --------------------------------------------------
 * The kernel steps the 8 sets of a warp in lockstep, so a warp pays the largest number of step attempts among its sets on
 * every output day; on widely spread batches (uniform in the bounds: the swarm initialisation of ParticleSwarmOptimizer.cpp:291)
 * that idles 22 % of the lane-attempts.  With a fitted model, evaluations of >= 32,768 sets hand the sets to the warps in an
 * order that puts sets with alike predicted attempt profiles together (two linear predictors per set, a counting sort, an
 * index list): results are bit-identical, a 1M-set uniform-in-bounds launch takes 100 ms instead of 113 ms.
 *   sepaihrd_fit_ordering   fits the model on a pilot of 2048 rows spread over `params` (one extra launch of a profiling
 *                           instantiation of the kernel + a small least-squares problem on the host, ~15 ms, synchronous).
 *                           Call it once per distribution of parameter sets.  sepaihrd_eval_batch (host buffers) does so by
 *                           itself: at its first large batch, and again when a batch is centred or spread differently than
 *                           the pilot; the `_device` entry points never fit on their own (they only enqueue).
 *   sepaihrd_set_ordering   0 = never reorder, 1 = reorder when a model exists (default).
 * 4-age problems in FAST arithmetic; other configurations simply stay unordered.  The swarm's and the sampler's own
 * evaluations are never reordered (measured on the 65,536-particle swarm: no gain -- its particles are not spread like the
 * uniform batch after the first update -- and their batches change every iteration).  No reference counterpart. */
sepaihrd_rc sepaihrd_set_ordering(sepaihrd_ctx* ctx, int32_t mode);
sepaihrd_rc sepaihrd_fit_ordering(sepaihrd_ctx* ctx, const double* params, int64_t B, int64_t ld, int32_t params_on_device);
sepaihrd_rc sepaihrd_ordering_state(const sepaihrd_ctx* ctx, int32_t* out_fitted, int64_t* out_fits);

/* Exact sample quantiles (linear interpolation between order statistics, numpy's default) of n_cols device-resident columns of
 * B doubles each, columns contiguous: d_columns[col * B + draw].  NaNs are left out (a column of NaNs gives NaN).  This is the
 * second half of sepaihrd_posterior_predictive, exposed for callers that hold their own series on the device (replaces the
 * per-column accumulators of ResultAggregator.cpp:276-371 for any series, not only the six built-in ones).
 *   probs   host, [n_probs], each in [0, 1];  d_out  device, [n_cols][n_probs];  enqueued on the ctx stream.
 *   path    0: by size (a cluster of 4 or 8 thread blocks holds a column in shared memory up to ~208 k draws, one block per
 *              column sweeping global memory beyond); 1: force the one-block-per-column kernel; 2: force the cluster kernel. */
sepaihrd_rc sepaihrd_column_quantiles_device(sepaihrd_ctx* ctx, const double* d_columns, int64_t B, int64_t n_cols, int32_t n_probs,
                                             const double* probs, double* d_out, int32_t path);

/* The aggregation passes (sepaihrd_posterior_predictive) keep their device work buffers in the ctx and reuse them across
 * calls; this frees them (they are also freed by sepaihrd_destroy). */
sepaihrd_rc sepaihrd_release_scratch(sepaihrd_ctx* ctx);

/* Page-locked host memory for the host-buffer entry points (sepaihrd_eval_batch, sepaihrd_simulate_batch, ...): parameter
 * and result buffers allocated here move at full PCIe / C2C rate and let the call overlap its copies with the kernel; any
 * other host pointer works too, through the driver's staging copies.  Free with sepaihrd_free_pinned.               */
sepaihrd_rc sepaihrd_alloc_pinned(size_t bytes, void** out);
void sepaihrd_free_pinned(void* ptr);

/* Block until everything enqueued on the ctx stream has finished. */
sepaihrd_rc sepaihrd_synchronize(sepaihrd_ctx* ctx);

/* Counters since creation: kernel launches issued by this ctx, parameter sets evaluated. */
sepaihrd_rc sepaihrd_get_counters(const sepaihrd_ctx* ctx, int64_t* launches, int64_t* sets);
/* Request merging of sepaihrd_eval_batch since creation: launches that served more than one call, and the calls they served. */
sepaihrd_rc sepaihrd_get_merge_counters(const sepaihrd_ctx* ctx, int64_t* merged_launches, int64_t* merged_requests);

/* Device-side FP64 pipe microbenchmark (dependent DFMA chains on every SM): returns the measured
 * peak in FP64 instructions/s (x2 for FMA-counted FLOP/s).  Used as the roofline denominator,
 * because MEASURED_PEAKS.json has no FP64 entry. */
sepaihrd_rc sepaihrd_measure_fp64_peak(int32_t device, double* out_dfma_per_second);

const char* sepaihrd_last_error(void);
const char* sepaihrd_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SEPAIHRD_B200_H */
