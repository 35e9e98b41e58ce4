/*
 * sepaihrd_oracle.h -- CPU oracle for the SEPAIHRD Dopri5 + Poisson-likelihood path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it, and
 * there only as the checker / the timed CPU baseline.  The product library
 * (libsepaihrd_b200.so) never links or calls this code and has no CPU fallback.
 *
 * What it is: a dependency-free C++17 restatement of the reference's algorithm
 * (adjo0043/Mathematical-Modeling-Of-Infectious-Diseases-V1), function by function, each citing
 * the reference file:line it follows.  Built with g++ -O2 -ffp-contract=off so the arithmetic is
 * unfused IEEE binary64 in the reference's source order (the reference's default build has no
 * -march flag, CMakeLists.txt:10,27-29, hence no FMA).
 *
 * PARITY STATUS
 *   pinned   : Poisson log-likelihood (reference KAT tests/model/SEPAIHRDObjectivefunctionTest.cpp:688-752),
 *              data-derived initial state (tests/utils/GetCalibrationDataTests.cpp:163-227,296-344),
 *              clamp / reflect constraints (formula identity), the RHS against a closed-form
 *              restatement in numpy.
 *   UNPINNED : the adaptive Dopri5 controller.  It lives in Boost.Odeint (un-vendored, version
 *              un-pinned, CMakeLists.txt:34; call site src/sir_age_structured/solvers/Dopri5SolverStrategy.cpp:28-37),
 *              Boost is not installed in the build container and the reference's tests pin no
 *              trajectory or full log-likelihood value.  The controller below restates the published
 *              algorithm of boost/numeric/odeint (controlled_runge_kutta<runge_kutta_dopri5>,
 *              default_error_checker, default_step_adjuster, integrate_times, >= 1.60).  The only
 *              external anchor is SURVEY.md section 8c: an independent transcription made during the
 *              survey gives logL = 1.206869676728e+06, 441 accepted / 45 rejected steps, 2917 RHS
 *              calls for the shipped default parameters; this oracle reproduces those numbers
 *              (tests/test_oracle.py).  "parity unpinned" therefore applies to trajectories and
 *              full log-likelihoods.
 */
#ifndef SEPAIHRD_ORACLE_H
#define SEPAIHRD_ORACLE_H

#include "../include/sepaihrd_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* name -> slot, restated independently of the product (SEPAIHRDParameterManager.cpp:197-267). */
int32_t sepaihrd_oracle_slot_for_name(int32_t n_ages, int32_t n_beta, int32_t n_kappa, const char* name);

/* SEPAIHRDParameterManager::applyConstraints (.cpp:315-347) for one vector of P values. */
void sepaihrd_oracle_apply_constraints(const sepaihrd_problem* pb, int32_t mode, const double* in, double* out);

/* AgeSEPAIHRDModel::computeDerivatives (src/model/AgeSEPAIHRDModel.cpp:101-228) for one state,
 * with the model parameters given as a slot vector. */
void sepaihrd_oracle_rhs(const sepaihrd_problem* pb, const double* slots, const double* state, double t,
                         double* dxdt);

/* SEPAIHRDObjectiveFunction::calculateSingleLogLikelihood (.cpp:241-279), sequential rows. */
double sepaihrd_oracle_poisson_ll(const double* simulated, const double* observed, int32_t rows, int32_t cols);

/* CalibrationData::getInitialSEPAIHRDState (src/utils/GetCalibrationData.cpp:107-234).
 * cum_* are the first rows of the cumulative matrices. out: [11*n]. */
void sepaihrd_oracle_initial_state_from_data(int32_t n, const double* population,
                                             const double* cum_confirmed0, const double* cum_deaths0,
                                             const double* cum_hosp0, const double* cum_icu0,
                                             double sigma, double gamma_p, double gamma_a, double gamma_i,
                                             const double* p_asymptomatic, double* out);

/* One full evaluation (calculate(), .cpp:62-235, null cache) with optional diagnostics.
 *   out_traj          [K][11n] or NULL
 *   out_interval_steps[K-1][2] accepted / rejected attempts per output interval, or NULL
 *   out_counts        [3] accepted, rejected, rhs_calls, or NULL
 * Returns logL (or -DBL_MAX); *out_status gets SEPAIHRD_ST_* bits. */
double sepaihrd_oracle_eval_one(const sepaihrd_problem* pb, const double* params, uint32_t* out_status,
                                double* out_traj, int32_t* out_interval_steps, int64_t* out_counts);

/* Diagnostics: one evaluation with (t, dt, err) of every step attempt written to out_trace[cap][3]; returns the number of
 * attempts (which may exceed cap). */
int64_t sepaihrd_oracle_trace_one(const sepaihrd_problem* pb, const double* params, int64_t cap, double* out_trace, double* out_ll);

/* B evaluations, OpenMP over sets (schedule(dynamic)); nthreads <= 0 = all cores.
 * out_status / out_steps ([B][2]) may be NULL. Returns the number of threads used. */
int32_t sepaihrd_oracle_eval_batch(const sepaihrd_problem* pb, const double* params, int64_t B, int64_t ld,
                                   double* out_ll, uint32_t* out_status, int32_t* out_steps, int32_t nthreads);

/* B simulations (trajectories), same selectors as sepaihrd_simulate_batch. */
int32_t sepaihrd_oracle_simulate_batch(const sepaihrd_problem* pb, const double* params, int64_t B, int64_t ld,
                                       int32_t what, int32_t stride, double* out, uint32_t* out_status,
                                       int32_t nthreads);

/* B simulations from caller-supplied initial states (Simulator::run semantics; quirk Q9).
 * state_stride == 0: one state shared by all sets. */
int32_t sepaihrd_oracle_simulate_from_state(const sepaihrd_problem* pb, const double* params, int64_t B, int64_t ld,
                                            const double* initial_states, int64_t state_stride, int32_t what,
                                            int32_t stride, double* out, uint32_t* out_status, int32_t nthreads);

/* The reference benchmark's jitter recipe (sepaihrd_objective_benchmark_main.cpp:452-460):
 * candidate_i = base_i + sigma_i * N(0,1) from std::mt19937(seed) + std::normal_distribution,
 * then applyConstraints in the problem's mode. out: [B][P]. */
void sepaihrd_oracle_jitter_params(const sepaihrd_problem* pb, const double* base, const double* sigmas,
                                   uint32_t seed, int64_t B, double* out);

/* PSO initialisation recipe (ParticleSwarmOptimizer.cpp:291): lo + u*(hi-lo), u ~ U[0,1) from
 * std::mt19937(seed) + std::uniform_real_distribution. out: [B][P]. */
void sepaihrd_oracle_uniform_params(const sepaihrd_problem* pb, uint32_t seed, int64_t B, double* out);

#ifdef __cplusplus
}
#endif
#endif
