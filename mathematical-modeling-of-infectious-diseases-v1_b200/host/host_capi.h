/*
 * host_capi.h -- C entry points of the host layer (libsepaihrd_host.so).
 *
 * The host layer is C++ (epidemic_host.hpp, optimizers.hpp: the reference's interface shapes over the
 * device C ABI).  These extern "C" wrappers exist for two callers that cannot include C++ headers:
 *   - the Python harness (tests/, bench.py, drivers.py), which steps the batched samplers and does the
 *     cross-GPU exchange with torch.distributed between the steps;
 *   - the parity tests, which run the SAME sampler code once against the device evaluator and once against
 *     the CPU oracle through a batch callback (the host library itself never links or loads the oracle).
 * Every function returns 0 on success; sepaihrd_host_last_error() gives the message of the last failure on
 * this thread (C++ exceptions never cross this boundary).
 */
#ifndef SEPAIHRD_HOST_CAPI_H
#define SEPAIHRD_HOST_CAPI_H

#include <stdint.h>

#include "sepaihrd_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

const char* sepaihrd_host_last_error(void);

/* OpenMP threads of the host-side sampler loops (proposal / particle updates).  torchrun exports OMP_NUM_THREADS=1 to
 * every rank; the drivers call this with (host cores / ranks per node).  n <= 0 leaves the setting alone; returns the
 * number of threads now in use. */
/* Directory for the Metropolis-Hastings trace files of whole runs (posterior_trace_checkpoint.csv, posterior_trace_final.csv,
 * posterior_trace.csv; MetropolisHastingsSampler.cpp:380-469).  NULL or "": back to the reference's rule (<project root>/data/
 * mcmc_samples inside a reference-style tree, nothing elsewhere).                                                             */
int32_t sepaihrd_host_set_trace_directory(const char* dir);
int32_t sepaihrd_host_set_threads(int32_t n);

/* B log-posteriors for B parameter rows ([B][ld] row-major).  Return non-zero to abort the run. */
typedef int32_t (*sepaihrd_host_batch_fn)(void* user, const double* params, int64_t B, int64_t ld, double* out);

/* ---- parameter manager over plain arrays (IParameterManager: sigmas, bounds, clamp / reflect) ------------- *
 * lower/upper: NaN = "no bounds entry" (SEPAIHRDParameterManager.cpp:337-343).  mode: 0 clamp, 1 reflect.     */
typedef struct sepaihrd_host_pm sepaihrd_host_pm;
int32_t sepaihrd_host_pm_create(int32_t n_params, const double* sigmas, const double* lower, const double* upper,
                                int32_t mode, sepaihrd_host_pm** out);
int32_t sepaihrd_host_pm_set_mode(sepaihrd_host_pm* pm, int32_t mode);
int32_t sepaihrd_host_pm_apply_constraints(const sepaihrd_host_pm* pm, const double* in, double* out);
void    sepaihrd_host_pm_destroy(sepaihrd_host_pm* pm);

/* ---- multi-chain Metropolis-Hastings (MetropolisHastingsSampler, step-wise) ---------------------------------- *
 * settings: the reference's keys (mcmc_iterations, burn_in, adaptation_period, thinning, regularization_epsilon,
 * target_acceptance_rate, adapt_scale, store_samples) plus n_chains, chain_offset, seed.                         */
typedef struct sepaihrd_host_mh sepaihrd_host_mh;
int32_t sepaihrd_host_mh_create(sepaihrd_host_pm* pm, int32_t n_settings, const char* const* keys, const double* values,
                                sepaihrd_host_mh** out);
int32_t sepaihrd_host_mh_set_initial_covariance(sepaihrd_host_mh* mh, const double* cov_colmajor, int32_t n);
int32_t sepaihrd_host_mh_begin(sepaihrd_host_mh* mh, const double* initial, const double* initial_logpost);
int32_t sepaihrd_host_mh_done(const sepaihrd_host_mh* mh);            /* 1 when all iterations have run */
int32_t sepaihrd_host_mh_iteration(const sepaihrd_host_mh* mh);
int32_t sepaihrd_host_mh_propose(sepaihrd_host_mh* mh, double* out_proposals /* [n_chains][P] */);
int32_t sepaihrd_host_mh_accept(sepaihrd_host_mh* mh, const double* proposed_logpost, uint8_t* out_accepted /* or NULL */);
int32_t sepaihrd_host_mh_state(const sepaihrd_host_mh* mh, double* out_x /* [n][P] */, double* out_logpost /* [n] */,
                               double* out_scale /* [n] */, int64_t* out_accepted /* [n] */);
int32_t sepaihrd_host_mh_best(const sepaihrd_host_mh* mh, double* out_x /* [P] */, double* out_value);
/* csrc/det_math.h on the host: the log / exp both Metropolis-Hastings samplers (host and device-resident) use */
double sepaihrd_host_det_log(double x);
double sepaihrd_host_det_exp(double x);
/* lower Cholesky factor [P*P] column-major of the start kernel all chains share until they adapt (after begin): what a
 * device-resident run of the same chains (sepaihrd_mh_begin) has to be given to reproduce them */
int32_t sepaihrd_host_mh_shared_cholesky(const sepaihrd_host_mh* mh, double* out_chol);
void    sepaihrd_host_mh_destroy(sepaihrd_host_mh* mh);

/* ---- particle swarm (ParticleSwarmOptimization) ----------------------------------------------------------------- *
 * settings: every key of the reference's pso_settings.txt (iterations, swarm_size, omega_start/end, c1_initial/final,
 * c2_initial/final, variant, topology, use_opposition_learning, use_adaptive_parameters, restart_threshold, max_stagnation,
 * quantum_beta, levy_alpha, ...) with the reference class's defaults, plus particle_offset, local_count, seed,
 * device_resident.  The step-wise calls (begin / tell / step, and their device forms) drive one contiguous shard of the
 * STANDARD / GLOBAL_BEST swarm (variant 0, topology 0, no opposition learning, no adaptive parameters).               */
typedef struct sepaihrd_host_pso sepaihrd_host_pso;
int32_t sepaihrd_host_pso_create(sepaihrd_host_pm* pm, int32_t n_settings, const char* const* keys, const double* values,
                                 sepaihrd_host_pso** out);
int32_t sepaihrd_host_pso_begin(sepaihrd_host_pso* pso, const double* initial_or_null);
int32_t sepaihrd_host_pso_local_count(const sepaihrd_host_pso* pso);
int32_t sepaihrd_host_pso_positions(const sepaihrd_host_pso* pso, double* out /* [local][P] */);
/* personal-best update; returns the shard's best personal best (value, local index, position) */
int32_t sepaihrd_host_pso_tell(sepaihrd_host_pso* pso, const double* fitness, double* out_best_value, int32_t* out_best_local,
                               double* out_best_position /* [P] */);
int32_t sepaihrd_host_pso_set_global_best(sepaihrd_host_pso* pso, double value, const double* position);
int32_t sepaihrd_host_pso_global_best(const sepaihrd_host_pso* pso, double* out_value, double* out_position);
int32_t sepaihrd_host_pso_step(sepaihrd_host_pso* pso, int32_t iter);
/* Device-resident form of the same swarm (sepaihrd_swarm_*): the shard's particles stay in HBM of `ctx`'s GPU; the same seeds
 * and arithmetic, hence the same positions, as begin / tell / step.  evaluate_device = objective launch + tell.  fetch copies
 * positions, velocities and personal bests back so that sepaihrd_host_pso_positions() can be read.                             */
int32_t sepaihrd_host_pso_begin_device(sepaihrd_host_pso* pso, const double* initial /* [P] or NULL */, sepaihrd_ctx* ctx);
int32_t sepaihrd_host_pso_evaluate_device(sepaihrd_host_pso* pso, double* out_best_value, int32_t* out_best_local_index,
                                          double* out_best_position /* [P] or NULL */);
int32_t sepaihrd_host_pso_step_device(sepaihrd_host_pso* pso, int32_t iter);
int32_t sepaihrd_host_pso_fetch(sepaihrd_host_pso* pso);
/* Whole run of ParticleSwarmOptimization::optimize on this handle against a batch callback: every variant / topology /
 * opposition / adaptation / restart / elitist-learning setting of the reference class (ParticleSwarmOptimizer.cpp:106-247).
 * out_stats[4] = objective evaluations, stagnation restarts, elitist-learning trials consumed, final swarm diversity.        */
int32_t sepaihrd_host_pso_run(sepaihrd_host_pso* pso, sepaihrd_host_batch_fn fn, void* user, const double* initial_or_null,
                              double* out_best /* [P] */, double* out_best_value, double* out_stats /* [4] or NULL */);
/* state of the swarm after sepaihrd_host_pso_run: personal-best values and current fitness, [swarm_size] each */
int32_t sepaihrd_host_pso_values(const sepaihrd_host_pso* pso, double* out_pbest_values, double* out_current_fitness);
/* getNeighbors (ParticleSwarmOptimizer.cpp:836-905) of the configured topology; returns the count, fills at most `cap`      */
int32_t sepaihrd_host_pso_neighbors(sepaihrd_host_pso* pso, int32_t particle, int32_t* out, int32_t cap);
void    sepaihrd_host_pso_destroy(sepaihrd_host_pso* pso);

/* ---- whole runs against a batch callback: "mh", "pso", "hill" or "nuts" (IOptimizationAlgorithm::optimize); "nuts" takes the
 * forward-difference gradient of the callback (P perturbed vectors per gradient, one batch) ------------------------------- */
int32_t sepaihrd_host_optimize(const char* algorithm, sepaihrd_host_pm* pm, int32_t n_settings, const char* const* keys,
                               const double* values, sepaihrd_host_batch_fn fn, void* user, const double* initial,
                               double* out_best /* [P] */, double* out_best_value, int64_t* out_n_evaluations);

/* ModelCalibrator::calibrate (two phases, covariance hand-off, batched re-scoring of the samples) against a batch callback:
 * phase 1 = "pso" | "hill", phase 2 = Metropolis-Hastings.  The parameter manager's mode is switched clamp -> reflect between
 * the phases exactly like ModelCalibrator.cpp:62-66, 88-92. */
int32_t sepaihrd_host_calibrate(const char* phase1, sepaihrd_host_pm* pm, int32_t n1, const char* const* keys1, const double* values1,
                                int32_t n2, const char* const* keys2, const double* values2, sepaihrd_host_batch_fn fn, void* user,
                                const double* initial, double* out_best /* [P] */, double* out_best_value, int64_t* out_n_samples,
                                double* out_phase1_value);

/* ---- the reference-shaped object graph over the device evaluator -------------------------------------------- *
 * Builds SEPAIHRDParameters -> PiecewiseConstantNpiStrategy -> AgeSEPAIHRDModel -> CalibrationData ->
 * SEPAIHRDModelCalibration (-> SEPAIHRDParameterManager, SEPAIHRDObjectiveFunction, AgeSEPAIHRDSimulator) from a
 * sepaihrd_problem plus the calibrated names and proposal sigmas.  Needs a CUDA device: no CPU fallback.          */
typedef struct sepaihrd_host_model sepaihrd_host_model;
int32_t sepaihrd_host_model_create(const sepaihrd_problem* problem, const char* const* param_names, const double* sigmas,
                                   sepaihrd_host_model** out);
/* SEPAIHRDObjectiveFunction::calculate / calculateBatch */
int32_t sepaihrd_host_model_calculate(sepaihrd_host_model* m, const double* params, double* out);
int32_t sepaihrd_host_model_calculate_batch(sepaihrd_host_model* m, const double* params, int64_t B, int64_t ld, double* out);
int32_t sepaihrd_host_model_set_constraint_mode(sepaihrd_host_model* m, int32_t mode);
/* SEPAIHRDParameterManager::getCurrentParameters / updateModelParameters (then getCurrentParameters again) */
int32_t sepaihrd_host_model_current_parameters(sepaihrd_host_model* m, double* out /* [P] */);
int32_t sepaihrd_host_model_update_parameters(sepaihrd_host_model* m, const double* params);
/* AgeSEPAIHRDSimulator::run(initial_state, times): out [K][11 n] */
int32_t sepaihrd_host_model_simulate(sepaihrd_host_model* m, const double* initial_state, const double* times, int32_t K, double* out);
/* SEPAIHRDModelCalibration::runPSOMCMC / runHillClimbingMCMC (phase1 = "pso" | "hill") */
int32_t sepaihrd_host_model_calibrate(sepaihrd_host_model* m, const char* phase1, int32_t n1, const char* const* keys1, const double* values1,
                                      int32_t n2, const char* const* keys2, const double* values2, double* out_best /* [P] */,
                                      double* out_best_value, int64_t* out_n_samples);
/* MetropolisHastingsSampler::optimize (MetropolisHastingsSampler.cpp:201-412) on the model's device objective, timed as a
 * whole: the reference's phase-2 sampler as shipped (one chain; setting "lookahead" != 1 evaluates the next K iterations'
 * proposals in one launch, see optimizers.hpp) or n_chains in lockstep.  out_last [P + 1]: final state of chain 0 and its
 * log-posterior; out_stats [6]: wall ms, parameter sets evaluated, kernel launches, acceptance rate, final scale, iterations. */
int32_t sepaihrd_host_model_metropolis(sepaihrd_host_model* m, int32_t n, const char* const* keys, const double* values, const double* initial,
                                       double* out_best /* [P] */, double* out_best_value, double* out_last, double* out_stats);
/* ResultAggregator::aggregatePosteriorPredictives over `samples` ([S][P]); out [6][T][n][5] in the order
 * lower_95, lower_90, median, upper_90, upper_95 (probabilities 0.025, 0.05, 0.5, 0.95, 0.975) */
int32_t sepaihrd_host_model_posterior_predictive(sepaihrd_host_model* m, const double* samples, int64_t S, int32_t num_samples_for_ppc,
                                                 uint32_t random_seed, const double* initial_state, double* out, int64_t* out_samples_used);
/* SEPAIHRDGradientObjectiveFunction::evaluate_with_gradient: the objective at params and its forward-difference gradient
 * (step epsilon * max(|x_i|, epsilon), epsilon <= 0: the reference's 1e-4), the P perturbed vectors as one device batch.     */
int32_t sepaihrd_host_model_gradient(sepaihrd_host_model* m, const double* params, double epsilon, double* out_value, double* out_grad /* [P] */);
/* AnalysisWriter (src/model/AnalysisWriter.cpp) from plain arrays, host only.
 * posterior predictive (.cpp:283-347): quantiles [6][T][n_ages][5] in the order lower_95, lower_90, median, upper_90, upper_95 (what
 * sepaihrd_host_model_posterior_predictive returns), observed [6][T][n_ages] or NULL -> 36 files <series>_<what>.csv in output_dir.
 * parameter posteriors (.cpp:201-281): samples [S][P] -> posterior_samples.csv, posterior_summary.csv.                          */
int32_t sepaihrd_host_write_posterior_predictive(const char* output_dir, int32_t T, int32_t n_ages, const double* time_points,
                                                 const double* quantiles, const double* observed_or_null);
int32_t sepaihrd_host_write_parameter_posteriors(const char* output_dir, const double* samples, int64_t S, int32_t P,
                                                 const char* const* names, int32_t burn_in, int32_t thinning);
/* The model is built with the cache that caches nothing (parity runs).  capacity > 0: the objective and the
 * calibration are rebuilt over a SimulationCache of that capacity (src/model/main.cpp:371 uses 1000); 0: back to none.
 * stats[4] = entries, getLikelihood calls, hits, storeLikelihood calls.                                                     */
int32_t sepaihrd_host_model_set_cache(sepaihrd_host_model* m, int64_t capacity);
int32_t sepaihrd_host_model_cache_stats(const sepaihrd_host_model* m, int64_t* out_stats /* [4] */);
void    sepaihrd_host_model_destroy(sepaihrd_host_model* m);

/* ---- SimulationCache (src/sir_age_structured/caching/SimulationCache.cpp) on its own, host only ---------------------- *
 * hash = computeHash (1e-8 quantisation, .cpp:35-52); get / store = the numeric-key getLikelihood / storeLikelihood
 * (.cpp:212-252); get_vector / set_vector = get / set (.cpp:106-150).  get returns 1 on a hit.                            */
typedef struct sepaihrd_host_cache sepaihrd_host_cache;
int32_t  sepaihrd_host_cache_create(int64_t capacity, sepaihrd_host_cache** out);
uint64_t sepaihrd_host_cache_hash(const sepaihrd_host_cache* c, const double* params, int32_t n);
int32_t  sepaihrd_host_cache_get(sepaihrd_host_cache* c, uint64_t key, double* out_value);
void     sepaihrd_host_cache_store(sepaihrd_host_cache* c, uint64_t key, double value);
int32_t  sepaihrd_host_cache_get_vector(sepaihrd_host_cache* c, const double* params, int32_t n, double* out_value);
void     sepaihrd_host_cache_set_vector(sepaihrd_host_cache* c, const double* params, int32_t n, double value);
/* evaluateThroughCache (what SEPAIHRDObjectiveFunction::calculateBatch does with its cache) against a batch callback: probe per
 * row, one callback for the distinct misses, store, repeats counted as hits.  status_of_row (or NULL): per-row status words,
 * indexed by the row's FIRST coordinate -- rows with a failure status other than NONFINITE are not stored (test hook).          */
int32_t  sepaihrd_host_cache_batch(sepaihrd_host_cache* c, int32_t n_params, const double* params, int64_t B, int64_t ld, double* out,
                                   sepaihrd_host_batch_fn fn, void* user, const uint32_t* status_of_row);
int64_t  sepaihrd_host_cache_size(const sepaihrd_host_cache* c);
void     sepaihrd_host_cache_clear(sepaihrd_host_cache* c);
void     sepaihrd_host_cache_stats(const sepaihrd_host_cache* c, int64_t* out_stats /* [3]: get calls, hits, store calls */);
void     sepaihrd_host_cache_destroy(sepaihrd_host_cache* c);

/* ---- the reference's on-disk formats (config_io.hpp: C++ readers / writer) ------------- *
 * Both return JSON text owned by the library (valid until the next call on this thread), or NULL with the message in
 * sepaihrd_host_last_error().  Non-finite numbers are written as null.
 *   kind = "parameters" (a = number of age classes), "bounds", "sigmas", "names", "settings", "matrix" (a rows, b cols),
 *          "data" (start / end date window; empty string = open)                                                       */
const char* sepaihrd_host_read_file_json(const char* kind, const char* filename, int32_t a, int32_t b, const char* start_date,
                                         const char* end_date);
/* loadReferenceProject(root, start, end): the problem main() assembles, with the keys of the Python Problem JSON
 * (format "sepaihrd_problem/1") plus "initial_state" (the state main() integrates from).                               */
const char* sepaihrd_host_project_json(const char* project_root, const char* start_date, const char* end_date, int32_t n_ages);
/* readSEPAIHRDParameters(in_file) -> saveCalibrationResults(out_file, ...): the writer's round trip */
int32_t sepaihrd_host_resave_parameters(const char* in_file, int32_t n_ages, const char* out_file, int32_t n_calibrated,
                                        const char* const* calibrated_names, double obj_value, const char* timestamp);

/* ---- post-calibration analysis (analysis.hpp: MetricsCalculator, ReproductionNumberCalculator, PostCalibrationAnalyser) *
 * Scalar metrics of one run, in this order: R0, overall_IFR, overall_attack_rate, peak_hospital_occupancy,
 * peak_ICU_occupancy, time_to_peak_hospital, time_to_peak_ICU, total_cumulative_deaths, max_Rt, min_Rt, final_Rt,
 * seroprevalence_at_target_day.  Per-age metrics: [4][n] = IFR, IHR, IICUR, attack rate.                                */
#define SEPAIHRD_HOST_NUM_METRICS 12
/* MetricsCalculator on a GIVEN trajectory ([K][11 n]) of the model described by `problem` (its base slots): host arithmetic
 * only, no device call.  out_rt / out_sero: [K] Rt and seroprevalence trajectories, or NULL.                            */
int32_t sepaihrd_host_metrics(const sepaihrd_problem* problem, const double* times, int32_t K, const double* trajectory,
                              const double* initial_state, double* out_scalars, double* out_age /* [4][n] or NULL */,
                              double* out_rt, double* out_sero);
/* PostCalibrationAnalyser, scenario step of generateFullReport: mean of the kept samples -> "baseline", "stricter_lockdown"
 * (first calibratable kappa x 0.9), "weaker_lockdown" (x 1.1), integrated as ONE device batch from `initial_state`.
 * out_scalars [3][12], out_age [3][4][n], out_kappa [3][n_kappa], out_trajectories [3][K][11 n] (each may be NULL);
 * csv_path non-empty: scenario_comparison.csv in the reference's format.                                                */
int32_t sepaihrd_host_model_scenarios(sepaihrd_host_model* m, const double* samples, int64_t S, int32_t burn_in, int32_t thinning,
                                      const double* initial_state, double* out_scalars, double* out_age, double* out_kappa,
                                      double* out_trajectories, const char* csv_path);
/* analyzeMCMCRunsInBatches without the files: one run per kept sample (one device batch); out_scalars [runs][12];
 * out_*_quantiles [5][K] in the order 0.025, 0.05, 0.5, 0.95, 0.975.                                                    */
int32_t sepaihrd_host_model_analyze_runs(sepaihrd_host_model* m, const double* samples, int64_t S, int32_t burn_in, int32_t thinning,
                                         const double* initial_state, double* out_scalars, double* out_rt_quantiles,
                                         double* out_sero_quantiles, int64_t* out_runs);

#ifdef __cplusplus
}
#endif
#endif /* SEPAIHRD_HOST_CAPI_H */
