// analysis.hpp -- post-calibration analysis over the device simulator: NPI scenario comparison and per-run essential
// metrics (BASELINE.json configs[4] "posterior-predictive + NPI scenario analysis").
//
// Reference classes mirrored (paths relative to the reference repository):
//   include/model/AnalysisTypes.hpp:14-39                 EssentialMetrics
//   src/model/ReproductionNumberCalculator.cpp:18-171     R0 / Rt from the next-generation matrix
//   src/model/MetricsCalculator.cpp:8-226                 essential metrics, Rt and seroprevalence trajectories
//   src/model/SimulationRunner.cpp:24-120                 parameters -> one simulation (NPI value alignment)
//   src/model/PostCalibrationAnalyser.cpp:92-140, 375-402 scenario definition (first calibratable kappa x0.9 / x1.1) and
//                                                         comparison; :300-330 sorted-sample trajectory quantiles
//   src/model/AnalysisWriter.cpp:439-477                  scenario_comparison.csv
//
// What is different by design: every simulation a call needs (baseline + all scenarios, or one run per posterior sample)
// is ONE batch on the device (AgeSEPAIHRDSimulator::runBatch -> sepaihrd_simulate_from_state); the metrics are reductions
// over the returned trajectories on the host cores.  The SimulationRunner cache of the reference is not reproduced
// (a hit may return the run of another parameter vector: hash of a lossy key, SimulationRunner.cpp:133-160).
#pragma once

#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "epidemic_host.hpp"

namespace epidemic {

struct EssentialMetrics {                                       // AnalysisTypes.hpp:14-39
    double R0 = 0.0, overall_IFR = 0.0, overall_attack_rate = 0.0;
    double peak_hospital_occupancy = 0.0, peak_ICU_occupancy = 0.0, time_to_peak_hospital = 0.0, time_to_peak_ICU = 0.0;
    double total_cumulative_deaths = 0.0;
    double max_Rt = 0.0, min_Rt = 1e6, final_Rt = 0.0;
    double seroprevalence_at_target_day = 0.0;                  // day 64 (ENE-COVID round 1)
    std::vector<double> age_specific_IFR, age_specific_IHR, age_specific_IICUR, age_specific_attack_rate;
    std::map<std::string, double> kappa_values;
};

// Spectral radius of the next-generation matrix F V^-1 over the (E, P, A, I) x age states.  Only the E rows of F are
// non-zero, so the non-zero spectrum is that of the n x n block  K(i, j) = T(i, j) (1/gamma_p + p_j/gamma_A +
// theta (1 - p_j)/(gamma_I + h_j)),  T(i, j) = beta kappa M(i, j) a_i h_infec_j X_i / N_j  (X = N for R0, S(t) for Rt);
// K is non-negative, its Perron root is found by power iteration.
class ReproductionNumberCalculator {
public:
    explicit ReproductionNumberCalculator(std::shared_ptr<AgeSEPAIHRDModel> model);
    double calculateR0() const;                                             // beta(0), kappa(0), X = N
    double calculateRt(const VectorXd& S_current, double time) const;       // beta(t), kappa(t), X = S(t)
    MatrixXd nextGenerationBlock(const VectorXd& X, double time, bool clamp_negative) const;
    static double spectralRadiusNonNegative(const MatrixXd& K);

private:
    std::shared_ptr<AgeSEPAIHRDModel> model_;
};

class MetricsCalculator {
public:
    EssentialMetrics calculateEssentialMetrics(const SimulationResult& sim_result, std::shared_ptr<AgeSEPAIHRDModel> model,
                                               const SEPAIHRDParameters& params, const VectorXd& initial_state,
                                               const std::vector<double>& time_points) const;
    std::vector<double> calculateRtTrajectory(const SimulationResult& sim_result, std::shared_ptr<AgeSEPAIHRDModel> model,
                                              const std::vector<double>& time_points) const;
    std::vector<double> calculateSeroprevalenceTrajectory(const SimulationResult& sim_result, const SEPAIHRDParameters& params,
                                                          const std::vector<double>& time_points) const;
};

// Parameters -> simulations.  runSimulations() integrates all parameter structs in one device batch.
class SimulationRunner {
public:
    SimulationRunner(std::shared_ptr<AgeSEPAIHRDModel> model_template, std::shared_ptr<IOdeSolverStrategy> solver,
                     double abs_error = 1.0e-6, double rel_error = 1.0e-6, double dt_hint = 1.0);
    ~SimulationRunner();
    SimulationResult runSimulation(const SEPAIHRDParameters& params, const VectorXd& initial_state, const std::vector<double>& time_points);
    std::vector<SimulationResult> runSimulations(const std::vector<SEPAIHRDParameters>& params, const VectorXd& initial_state,
                                                 const std::vector<double>& time_points);
    // A copy of the template carrying `params`, NPI values aligned as SimulationRunner.cpp:45-88 does
    std::shared_ptr<AgeSEPAIHRDModel> modelFor(const SEPAIHRDParameters& params) const;

private:
    std::shared_ptr<AgeSEPAIHRDModel> model_template_;
    std::shared_ptr<IOdeSolverStrategy> solver_;
    double abs_err_, rel_err_, dt_hint_;
    std::unique_ptr<AgeSEPAIHRDSimulator> simulator_;
    double sim_start_ = 0.0, sim_end_ = 0.0;
};

struct AggregatedTrajectory {                                   // per time point: sorted-sample quantiles
    std::vector<double> median, q025, q975, q05, q95;
};

class PostCalibrationAnalyser {
public:
    using NamedParameters = std::pair<std::string, SEPAIHRDParameters>;
    using NamedMetrics = std::pair<std::string, EssentialMetrics>;

    PostCalibrationAnalyser(std::shared_ptr<AgeSEPAIHRDModel> model_template, std::shared_ptr<IOdeSolverStrategy> solver,
                            const std::vector<double>& time_points, const VectorXd& initial_state,
                            double abs_error = 1.0e-6, double rel_error = 1.0e-6);

    // mean of samples[burn_in], samples[burn_in + thinning], ... (PostCalibrationAnalyser.cpp:96-103)
    static VectorXd meanOfSamples(const std::vector<VectorXd>& samples, int burn_in, int thinning, std::ptrdiff_t n_params);
    // "stricter_lockdown" / "weaker_lockdown": the first calibratable kappa x0.9 / x1.1 (.cpp:108-135)
    std::vector<NamedParameters> defineNpiScenarios(const SEPAIHRDParameters& baseline_params) const;
    // baseline + scenarios as ONE device batch, metrics per run, "baseline" first (.cpp:375-402)
    std::vector<NamedMetrics> performScenarioAnalysis(const SEPAIHRDParameters& baseline_params, const std::vector<NamedParameters>& scenarios,
                                                      std::vector<SimulationResult>* trajectories = nullptr);
    // generateFullReport step 4: mean parameters through the parameter manager, then the two NPI scenarios
    std::vector<NamedMetrics> scenarioAnalysisFromSamples(const std::vector<VectorXd>& samples, SEPAIHRDParameterManager& param_manager,
                                                          int burn_in, int thinning, std::vector<SimulationResult>* trajectories = nullptr);
    // analyzeMCMCRunsInBatches without the file output: one run per kept sample (one device batch), its metrics, and the
    // Rt / seroprevalence trajectories aggregated to quantiles
    struct McmcAnalysis {
        std::vector<EssentialMetrics> metrics;
        AggregatedTrajectory rt, seroprevalence;
    };
    McmcAnalysis analyzeMCMCRuns(const std::vector<VectorXd>& samples, SEPAIHRDParameterManager& param_manager, int burn_in, int thinning);

    static AggregatedTrajectory aggregateTrajectories(const std::vector<std::vector<double>>& trajectories, size_t num_timesteps);
    static void writeScenarioComparison(const std::string& filepath, const std::vector<NamedMetrics>& scenarios);   // AnalysisWriter.cpp:439-477

private:
    std::shared_ptr<AgeSEPAIHRDModel> model_template_;
    std::vector<double> time_points_;
    VectorXd initial_state_;
    SimulationRunner runner_;
    MetricsCalculator metrics_;
};

// AnalysisWriter (src/model/AnalysisWriter.cpp): the report files the reference's plotting scripts read.  The reference queues
// the writes on a worker thread; here they are written when the call is made (waitForCompletion() is kept and does nothing).
//   savePosteriorPredictiveData  .cpp:283-347   <series>_{median,lower90,upper90,lower95,upper95,observed}.csv for the six series
//   saveParameterPosteriors      .cpp:201-281   posterior_samples.csv, posterior_summary.csv
//   saveScenarioComparison       .cpp:439-477   scenario_comparison.csv
// Number formats are the reference's, stream state included: in the posterior-predictive files the time column of the FIRST row
// is printed with the stream's default format and, once a value has switched the stream to fixed / 6 digits, every later one in
// that format ("0", then "1.000000", ...).
class AnalysisWriter {
public:
    void savePosteriorPredictiveData(const std::string& output_dir, const PosteriorPredictiveData& ppd_data) const;
    void saveParameterPosteriors(const std::string& output_dir, const std::vector<VectorXd>& param_samples,
                                 const std::vector<std::string>& param_names, int burn_in, int thinning) const;
    void saveScenarioComparison(const std::string& filepath, const std::vector<PostCalibrationAnalyser::NamedMetrics>& scenarios) const {
        PostCalibrationAnalyser::writeScenarioComparison(filepath, scenarios);
    }
    void waitForCompletion() const {}
};

}  // namespace epidemic
