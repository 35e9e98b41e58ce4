// optimizers.hpp -- batched forms of the reference's calibrators: every iteration hands ALL
// chains / particles to IObjectiveFunction::calculateBatch, i.e. to one launch of the fused device kernel.
//
//   MetropolisHastingsSampler   src/sir_age_structured/optimizers/MetropolisHastingsSampler.cpp:201-412 (one chain)
//                               -> n_chains independent chains of exactly that algorithm, stepped in lockstep
//   ParticleSwarmOptimization   src/model/optimizers/ParticleSwarmOptimizer.cpp:10-948 (whole class)
//                               -> all five variants (STANDARD, QUANTUM, ADAPTIVE, LEVY_FLIGHT, HYBRID), all four topologies
//                                  (GLOBAL_BEST, LOCAL_BEST ring, VON_NEUMANN grid, RANDOM_DYNAMIC), opposition-based
//                                  initialisation, evolutionary-state parameter adaptation, stagnation restart, elitist
//                                  learning; every swarm evaluation is ONE batch.  The STANDARD / GLOBAL_BEST swarm with
//                                  linear schedules additionally has a step-wise (shardable) and a device-resident form
//   HillClimbingOptimizer       src/sir_age_structured/optimizers/HillClimbingOptimizer.cpp:38-109, 131-352
//   ModelCalibrator             src/sir_age_structured/ModelCalibrator.cpp:22-168 (two phases + covariance hand-off)
//   SEPAIHRDModelCalibration    src/model/SEPAIHRDModelCalibration.cpp:24-236
//
// The reference seeds every generator from std::random_device (MetropolisHastingsSampler.cpp:20-23,
// ParticleSwarmOptimizer.hpp:578); here a "seed" setting makes runs repeatable: chain c draws from
// std::mt19937(std::seed_seq{seed, c}) with c the GLOBAL chain index, so a run sharded over several GPUs
// (settings "chain_offset" / "particle_offset") visits exactly the states of the single-process run.
#pragma once

#include <deque>
#include <random>
#include <string>

#include "epidemic_host.hpp"

namespace epidemic {

class MetropolisHastingsSampler : public IOptimizationAlgorithm {
public:
    MetropolisHastingsSampler();
    void configure(const std::map<std::string, double>& settings) override;
    OptimizationResult optimize(const VectorXd& initialParameters, IObjectiveFunction& objectiveFunction,
                                IParameterManager& parameterManager) override;
    void setInitialCovariance(const MatrixXd& cov);
    // Trace files of optimize() (.cpp:380-382, 399-409, 414-469): `posterior_trace_checkpoint.csv` (last 5000 samples, every
    // report_interval), `posterior_trace_final.csv` and `posterior_trace.csv` (all samples) -- columns iter, log_posterior, one
    // per parameter, written when the settings write_checkpoints / write_trace are on (default, like the reference) AND a
    // directory is known: the one set here, else the process-wide default, else <project root>/data/mcmc_samples when the
    // working directory lies in a reference-style tree (data/, include/, src/ -- FileUtils::getProjectRoot).  Outside such a
    // tree and without a directory nothing is written (the reference would create ./data/mcmc_samples).
    void setOutputDirectory(const std::string& dir) { output_dir_ = dir; }
    static void setDefaultOutputDirectory(const std::string& dir);
    static void saveSamplesToCSV(const std::vector<VectorXd>& samples, const std::vector<double>& objectiveValues,
                                 const std::vector<std::string>& parameterNames, const std::string& filepath, size_t first = 0);

    // ---- step-wise form (used by optimize() and by the multi-GPU driver) --------------------------------------
    // begin(): chains start at `initial` ([P], shared) with log-posteriors `initial_logpost` ([n_chains]).
    void begin(const VectorXd& initial, const double* initial_logpost, IParameterManager& pm);
    int iteration() const { return t_; }                 // next iteration index (1-based like the reference loop)
    bool done() const { return t_ >= iterations_; }
    // propose(): adaptation step + proposal + constraints for every chain; out is [n_chains][P] row-major
    void propose(IParameterManager& pm, double* out);
    // accept(): Metropolis accept/reject from the proposals' log-posteriors; accepted_out (optional) gets 0/1
    void accept(const double* proposed_logpost, uint8_t* accepted_out = nullptr);
    int numChains() const { return n_chains_; }
    int numParams() const { return n_params_; }
    const double* currentPositions() const { return cur_x_.data(); }       // [n_chains][P]
    const double* currentLogPost() const { return cur_lp_.data(); }        // [n_chains]
    const MatrixXd& sharedCholesky() const { return shared_chol_; }         // lower factor of the start kernel (after begin)
    double globalScale(int chain) const { return chains_[static_cast<size_t>(chain)].global_scale; }
    long acceptedCount(int chain) const { return chains_[static_cast<size_t>(chain)].accepted; }
    OptimizationResult result(int upto_iteration = -1) const;
    static double safeValue(double v) { return (std::isnan(v) || std::isinf(v)) ? -1e18 : v; }   // safeEvaluate, .cpp:65-74

    // Look-ahead (setting "lookahead": 0 = sized from the recent acceptance rate, 1 = off, K = that many; up to 2048 chains).
    // The reference's shipped run is ONE chain of 100 000 sequential iterations (data/configuration/mcmc_settings.txt,
    // .cpp:283-384): one evaluation per device launch, and a launch costs ~0.5 ms whatever it holds.  Between two accepted
    // proposals the chain does not move, so the proposals of the next K iterations ARE known before any of them is evaluated:
    // iteration t + j proposes x + s_j L z_j with z_j the next normals of the generator (a rejected iteration has consumed its
    // uniform, .cpp:323-329) and s_j the scale after j more rejections (.cpp:104-152).  optimize() evaluates those K as one
    // batch, commits iterations up to and including the first accepted one and throws the rest away.  The chain, its
    // generator, scale, covariance and trace files are bit for bit those of the sequential run (tests/test_mh_lookahead.py).
    // With several chains every chain looks ahead on its own (the chains then run apart by a few iterations between launches;
    // a launch holds at most 4096 proposals, so the windows shrink as the chain count grows: 256 chains x 16, 2048 x 2).
    int lookahead() const { return lookahead_; }
    long speculatedEvaluations() const { return speculated_; }          // evaluations made / iterations they committed
    long committedIterations() const { return committed_; }

private:
    struct ScaleState {                   // what adaptGlobalScale reads and writes
        double log_scale = 0.0, global_scale = 1.0;
        std::deque<uint8_t> recent;
        int recent_sum = 0;
        int emergency_shrink_count = 0;
    };
    struct Chain : ScaleState {
        std::mt19937 gen;
        int t = 1;                        // next iteration of THIS chain (== iteration() in the lockstep loop; look-ahead lets chains run apart)
        long accepted = 0;
        bool own_kernel = false;          // false: shares the initial Cholesky factor
        MatrixXd cov, chol;
        VectorXd running_mean;
        std::vector<double> hist;         // chain_history_: [states][P], contiguous
        VectorXd hist_sum;                // sum of the history in order (the mean of the full recomputation)
        std::vector<double> centred;      // work buffer of recomputeFullCovariance
        double best_lp = -std::numeric_limits<double>::infinity();
        VectorXd best_x;
        std::vector<VectorXd> samples;
        std::vector<double> sample_lp;
    };
    void adaptGlobalScale(ScaleState& c, bool accepted, int step) const;   // .cpp:104-152
    void adaptKernel(Chain& c, int step) const;                          // .cpp:286-303
    bool acceptOne(Chain& c, int ci, double proposed_logpost);           // .cpp:310-367
    void drawProposal(std::mt19937& gen, const Chain& c, double scale, const double* x, IParameterManager& pm, double* out) const;   // .cpp:91-102, 308
    int windowLength(const Chain& c, int running_chains, int share) const;
    void runLookahead(IObjectiveFunction& f, IParameterManager& pm, const std::string& dir);
    void updateCovarianceRank1(Chain& c, int step) const;               // .cpp:154-168
    void recomputeFullCovariance(Chain& c) const;                       // .cpp:170-199
    void ownKernel(Chain& c) const;
    void pushHistory(Chain& c, const double* x) const;

    int iterations_ = 10000, burn_in_ = 1000, adaptation_period_ = 100, report_interval_ = 100, thinning_ = 1;
    bool write_checkpoints_ = true, write_trace_ = true;
    std::string output_dir_;
    std::string traceDirectory() const;
    void saveCheckpoint(const OptimizationResult& res, IParameterManager& pm, bool final, const std::string& dir) const;
    double regularization_epsilon_ = 1e-6, target_acceptance_rate_ = 0.234;
    bool adapt_scale_ = true, store_samples_ = true;
    int n_chains_ = 1;
    int lookahead_ = 0;
    static constexpr int LOOKAHEAD_SETS = 4096;          // proposals per launch: up to here a launch costs what one set costs
    long speculated_ = 0, committed_ = 0;
    double launch_seconds_ = 0.0, proposal_seconds_ = 0.0;     // running means: fixed part of an objective call; host arithmetic per proposal
    double row_seconds_ = 0.0, probe_rows_[2] = {0.0, 0.0}, probe_seconds_[2] = {0.0, 0.0};   // per-row part of an objective call, from the probe windows
    double commit_seconds_ = 0.0;                              // host time per committed iteration
    long calls_seen_ = 0;
    long chain_offset_ = 0;
    bool shared_diagonal_ = false;
    bool has_seed_ = false;
    unsigned seed_ = 0;
    bool hasInitialCovariance_ = false;
    MatrixXd initialCovariance_;
    // run state
    int n_params_ = 0, t_ = 1;
    bool keep_history_ = false;
    MatrixXd shared_cov_, shared_chol_;
    std::vector<Chain> chains_;
    std::vector<double> cur_x_, cur_lp_, prop_x_;
};

class ParticleSwarmOptimization : public IOptimizationAlgorithm {
public:
    void configure(const std::map<std::string, double>& settings) override;
    OptimizationResult optimize(const VectorXd& initialParameters, IObjectiveFunction& objectiveFunction,
                                IParameterManager& parameterManager) override;

    // ---- step-wise form: this process owns particles [particle_offset, particle_offset + local_count) --------
    // begin(): draws the initial swarm (uniform in bounds; particle 0 = clamped initialParameters); positions()
    // then holds what has to be evaluated.
    void begin(const VectorXd* initialParameters, IParameterManager& pm);
    const double* positions() const { return pos_.data(); }              // [local][P]
    int localCount() const { return local_; }
    int swarmSize() const { return swarm_size_; }
    int numParams() const { return n_; }
    int iterations() const { return iterations_; }
    // tell(): fitness of positions(); updates personal bests; returns the best (value, LOCAL index) of this shard
    std::pair<double, int> tell(const double* fitness);
    const double* personalBest(int local_index) const { return pbest_.data() + static_cast<size_t>(local_index) * n_; }
    // global best as agreed by all shards (single process: the shard's own best)
    void setGlobalBest(double value, const double* position);
    double globalBestValue() const { return gbest_value_; }
    const std::vector<double>& globalBestPosition() const { return gbest_; }
    // step(): velocity/position update of iteration `iter` (0-based) for every local particle
    void step(int iter);
    MatrixXd personalBestScatter(VectorXd& mean_out) const;               // sum of (pbest - mean)(pbest - mean)^T pieces for the covariance hand-off

    // ---- device-resident form (sepaihrd_swarm_*, csrc/sepaihrd_swarm.cu): the shard's particles live in HBM ----------
    // Same seeds, same generator, same unfused arithmetic as begin()/tell()/step(): the swarm visits exactly the same
    // positions, but an iteration moves one seed per particle and the global best instead of the whole swarm.
    // optimize() takes this path by itself when the objective is a SEPAIHRDObjectiveFunction (setting
    // "device_resident", default 1).  `ctx` must have been created with the parameter manager's bounds.
    void beginDevice(const VectorXd* initialParameters, IParameterManager& pm, sepaihrd_ctx* ctx);
    std::pair<double, int> evaluateDevice(double* best_position /* [P] or null */);   // objective launch + tell on the device
    void stepDevice(int iter);
    void fetchPersonalBests(bool with_positions = true);                  // device -> pbest_ / pbest_val_ (covariance hand-off) [+ pos_ / vel_]
    bool onDevice() const { return dev_swarm_ != nullptr; }
    ~ParticleSwarmOptimization() override;
    ParticleSwarmOptimization() = default;
    ParticleSwarmOptimization(const ParticleSwarmOptimization&) = delete;
    ParticleSwarmOptimization& operator=(const ParticleSwarmOptimization&) = delete;


    // ---- the reference's enums (ParticleSwarmOptimizer.hpp) and the state they drive ---------------------------------
    enum class PSOVariant { STANDARD = 0, QUANTUM = 1, ADAPTIVE = 2, LEVY_FLIGHT = 3, HYBRID = 4 };
    enum class TopologyType { GLOBAL_BEST = 0, LOCAL_BEST = 1, VON_NEUMANN = 2, RANDOM_DYNAMIC = 3 };
    enum class EvolutionaryState { EXPLORATION, EXPLOITATION, CONVERGENCE, JUMPING_OUT };
    PSOVariant variant() const { return variant_; }
    TopologyType topology() const { return topology_; }
    // true when the configured swarm is the STANDARD update on the GLOBAL_BEST topology with linear coefficient schedules
    // and no opposition-based initialisation: the form the step-wise and the device-resident interfaces implement
    bool isBasicSwarm() const;
    int restartCount() const { return restarts_; }                        // stagnation restarts of the last optimize()
    int elitistTrials() const { return els_trials_; }                     // elitist-learning trial evaluations of the last optimize()
    long evaluations() const { return evaluations_; }                     // objective evaluations of the last optimize()
    // getNeighbors (.cpp:836-905) for the configured topology; RANDOM_DYNAMIC draws from the master generator
    std::vector<int> getNeighbors(int particle_idx);
    // whole-swarm state of the last optimize() on the host path (tests, diagnostics)
    const std::vector<double>& personalBestValues() const { return pbest_val_; }
    const std::vector<double>& currentFitness() const { return cur_fit_; }
    double swarmDiversity() const;                                        // calculateSwarmDiversity (.cpp:679-703)

private:
    int iterations_ = 100, swarm_size_ = 30, report_interval_ = 10;
    double omega_start_ = 0.9, omega_end_ = 0.4, c1_initial_ = 2.5, c1_final_ = 0.5, c2_initial_ = 0.5, c2_final_ = 2.5;
    // defaults of the reference class (ParticleSwarmOptimizer.hpp:201-236)
    PSOVariant variant_ = PSOVariant::ADAPTIVE;
    TopologyType topology_ = TopologyType::GLOBAL_BEST;
    bool use_opposition_learning_ = true, use_parallel_ = false, use_adaptive_parameters_ = true, log_evolutionary_state_ = true;
    double diversity_threshold_ = 0.1, restart_threshold_ = 1e-6, quantum_beta_ = 1.0, levy_alpha_ = 1.5;
    int stagnation_counter_ = 0, max_stagnation_ = 50;
    std::uniform_real_distribution<> uniform_dist_{0.0, 1.0};
    std::normal_distribution<> normal_dist_{0.0, 1.0};
    std::vector<double> cur_fit_, success_rate_;
    std::vector<int> success_count_, total_updates_;
    int restarts_ = 0, els_trials_ = 0;
    long evaluations_ = 0;
    // the whole-swarm engine behind optimize() (host-resident arrays, one calculateBatch per swarm evaluation)
    void evaluateSwarm(IObjectiveFunction& f, int first, std::vector<double>& fitness);
    void initializeSwarmFull(const VectorXd* init, IObjectiveFunction& f, IParameterManager& pm);
    void oppositionBasedInitialization();
    void permuteSwarm(const std::vector<int>& order);
    void runHostLoop(int start_iter, double previous_gbest, IObjectiveFunction& f);
    void restartSwarm(IObjectiveFunction& f, int keep_best_count = 3);
    void updateParticles(int iter, IObjectiveFunction& f);
    double calculateEvolutionaryFactor() const;
    EvolutionaryState estimateEvolutionaryState() const;
    void adaptParameters(EvolutionaryState state, int iter, double& omega, double& c1, double& c2);
    void standardPSOUpdate(int i, const double* lbest, double omega, double c1, double c2, std::mt19937& rng);
    void quantumPSOUpdate(int i, const std::vector<double>& mean_best, int iter, std::mt19937& rng);
    void levyFlightUpdate(int i, double omega, double c1, double c2, std::mt19937& rng);
    double generateLevyNumber(std::mt19937& rng) const;
    void applyElitistLearningStrategy(int best, IObjectiveFunction& f);
    void rescanGlobalBest();
    OptimizationResult finish() const;
    void reportProgress(int iter, const char* where) const;
    long particle_offset_ = 0;
    int local_count_setting_ = -1;
    bool has_seed_ = false;
    unsigned seed_ = 0;
    // run state
    int n_ = 0, local_ = 0;
    std::mt19937 rng_;
    std::vector<double> lb_, ub_, pos_, vel_, pbest_, pbest_val_, gbest_;
    double gbest_value_ = -std::numeric_limits<double>::infinity();
    bool first_tell_ = true;
    bool device_resident_ = true;
    sepaihrd_swarm* dev_swarm_ = nullptr;
    void setupRun(IParameterManager& pm);
    void drawInitialSwarm(const VectorXd* initialParameters);
    std::vector<uint32_t> drawSeeds();
    void coefficients(int iter, double& omega, double& c1, double& c2) const;
};

// HillClimbingOptimizer   src/sir_age_structured/optimizers/HillClimbingOptimizer.cpp:131-352: a candidate cloud per
// iteration (ONE batch), then the robust line search (.cpp:38-109) whose <= 10 backtracking and <= 12 expansion
// candidates are evaluated speculatively as two batches and consumed in the reference's sequential order.
class HillClimbingOptimizer : public IOptimizationAlgorithm {
public:
    void configure(const std::map<std::string, double>& settings) override;
    OptimizationResult optimize(const VectorXd& initialParameters, IObjectiveFunction& objectiveFunction,
                                IParameterManager& parameterManager) override;

private:
    bool lineSearch(VectorXd& current, double& current_logL, const VectorXd& direction, IObjectiveFunction& func, IParameterManager& pm) const;
    int iterations_ = 2000, report_interval_ = 100, cloud_size_multiplier_ = 8, cloud_size_ = 0;
    bool has_seed_ = false;
    unsigned seed_ = 0;
};

// NUTSSampler   src/model/optimizers/NUTSSampler.cpp:16-426: the No-U-Turn sampler with dual-averaging step-size adaptation
// (Hoffman & Gelman 2014, algorithm 6 as the reference writes it: heuristic first step size, gradients rescaled to norm
// <= 1000, constraints applied inside the leapfrog step, slice variable in log space, DELTA_MAX = 1000).  It needs an
// IGradientObjectiveFunction; every gradient is one batch of P perturbed vectors on the device.  The reference evaluates the
// gradient at a leapfrog end point and again, at the same point, when the next leapfrog starts or the tree leaf is scored:
// the last (point, value, gradient) triple is remembered here, which changes no number.  Extra setting: `seed`.
class NUTSSampler : public IOptimizationAlgorithm {
public:
    NUTSSampler();
    void configure(const std::map<std::string, double>& settings) override;
    OptimizationResult optimize(const VectorXd& initialParameters, IObjectiveFunction& objectiveFunction,
                                IParameterManager& parameterManager) override;
    long gradientEvaluations() const { return gradient_evaluations_; }      // distinct points whose gradient was computed
    double finalEpsilon() const { return final_epsilon_; }
    const std::vector<int>& treeDepths() const { return tree_depths_; }

private:
    struct Tree {
        VectorXd theta_minus, theta_plus, r_minus, r_plus, theta_prime;
        int n_valid = 0;
        bool s = false;
        double alpha = 0.0;
        int n_alpha = 0;
    };
    double gradientAt(IGradientObjectiveFunction& objective, const VectorXd& theta, VectorXd& grad) const;
    double findReasonableEpsilon(IGradientObjectiveFunction& objective, const VectorXd& theta, IParameterManager& pm) const;
    void leapfrog(IGradientObjectiveFunction& objective, VectorXd& theta, VectorXd& r, double epsilon, IParameterManager& pm) const;
    void buildTree(IGradientObjectiveFunction& objective, const VectorXd& theta, const VectorXd& r, double log_u_slice, int v, int j,
                   double epsilon, double H0, IParameterManager& pm, Tree& tree) const;
    static bool checkNoUTurn(const VectorXd& theta_minus, const VectorXd& theta_plus, const VectorXd& r_minus, const VectorXd& r_plus);
    int num_iterations_ = 2000, adaptation_window_ = 500, max_tree_depth_ = 10;
    double delta_target_ = 0.8;
    static constexpr double DELTA_MAX = 1000.0;
    bool has_seed_ = false;
    unsigned seed_ = 0;
    mutable std::mt19937 rng_;
    // the last gradient evaluation (see above)
    mutable bool memo_valid_ = false;
    mutable VectorXd memo_theta_, memo_grad_;
    mutable double memo_value_ = 0.0;
    mutable long gradient_evaluations_ = 0;
    double final_epsilon_ = 0.0;
    std::vector<int> tree_depths_;
};

class ModelCalibrator {
public:
    static constexpr const char* PHASE1_NAME = "Phase1";
    static constexpr const char* PHASE2_NAME = "Phase2";
    ModelCalibrator(std::unique_ptr<IParameterManager> parameterManager, std::unique_ptr<IObjectiveFunction> objectiveFunction,
                    std::map<std::string, std::unique_ptr<IOptimizationAlgorithm>> algorithms);
    void calibrate(const std::map<std::string, double>& phase1_settings, const std::map<std::string, double>& phase2_settings);
    const VectorXd& getBestParameterVector() const { return best_params_vector_; }
    double getBestObjectiveValue() const { return best_objective_value_; }
    const std::vector<VectorXd>& getMCMCSamples() const { return phase2_result_.samples; }
    const std::vector<double>& getMCMCObjectiveValues() const { return mcmcObjectiveValues_; }
    const OptimizationResult& getPhase1Result() const { return phase1_result_; }
    const OptimizationResult& getPhase2Result() const { return phase2_result_; }
    IParameterManager& getParameterManager() { return *parameterManager_; }
    IObjectiveFunction& getObjectiveFunction() { return *objectiveFunction_; }

private:
    std::unique_ptr<IParameterManager> parameterManager_;
    std::unique_ptr<IObjectiveFunction> objectiveFunction_;
    std::map<std::string, std::unique_ptr<IOptimizationAlgorithm>> optimization_algorithms_;
    VectorXd best_params_vector_;
    double best_objective_value_ = -std::numeric_limits<double>::infinity();
    OptimizationResult phase1_result_, phase2_result_;
    std::vector<double> mcmcObjectiveValues_;
};

class SEPAIHRDModelCalibration {
public:
    SEPAIHRDModelCalibration(std::shared_ptr<AgeSEPAIHRDModel> model_ptr, const CalibrationData& calibration_data,
                             const std::vector<double>& time_points, const std::vector<std::string>& params_to_calibrate,
                             const std::map<std::string, double>& proposal_sigmas,
                             const std::map<std::string, std::pair<double, double>>& param_bounds,
                             std::shared_ptr<IOdeSolverStrategy> solver_strategy = std::make_shared<Dopri5SolverStrategy>(),
                             std::shared_ptr<ISimulationCache> cache = std::make_shared<SimulationCache>());
    SEPAIHRDParameterManager& getParameterManager();
    VectorXd getCurrentParameterValues();
    ModelCalibrator runPSOMCMC(const std::map<std::string, double>& phase1_settings, const std::map<std::string, double>& phase2_settings);
    ModelCalibrator runHillClimbingMCMC(const std::map<std::string, double>& phase1_settings, const std::map<std::string, double>& phase2_settings);
    ModelCalibrator runNUTS(const std::map<std::string, double>& nuts_settings);        // .cpp:210-236: NUTS as phase 2, no phase 1

private:
    ModelCalibrator setupCalibrator(std::map<std::string, std::unique_ptr<IOptimizationAlgorithm>> algorithms);
    std::shared_ptr<AgeSEPAIHRDModel> model_;
    const CalibrationData& observed_data_;       // by const reference, like the reference: the caller owns it
    std::vector<double> time_points_;
    std::vector<std::string> params_to_calibrate_;
    std::map<std::string, double> proposal_sigmas_;
    std::map<std::string, std::pair<double, double>> param_bounds_;
    std::shared_ptr<IOdeSolverStrategy> solver_strategy_;
    std::shared_ptr<ISimulationCache> cache_;
    VectorXd initial_state_cached_;
    std::unique_ptr<SEPAIHRDParameterManager> parameter_manager_;
};

}  // namespace epidemic
