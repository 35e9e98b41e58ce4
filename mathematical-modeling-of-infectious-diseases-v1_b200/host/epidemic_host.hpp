// epidemic_host.hpp -- C++ host mirror of the reference's plugin interfaces for the SEPAIHRD hot path.
//
// Same names, argument meaning and error behaviour as the reference classes they stand in for, so that the
// reference's calibrators can be pointed at the B200 evaluator without source changes beyond the include
// path.  Everything numerical happens behind the C ABI of include/sepaihrd_b200.h; there is
// no CPU implementation of the right-hand side, the solver or the likelihood in this directory.
//
// Reference files mirrored (paths relative to the reference repository):
//   include/sir_age_structured/interfaces/{IObjectiveFunction,IParameterManager,IOdeSolverStrategy,
//       ISimulationCache,IOptimizationAlgorithm}.hpp
//   include/sir_age_structured/SimulationResult.hpp, include/exceptions/Exceptions.hpp
//   include/model/parameters/{SEPAIHRDParameters,SEPAIHRDParameterManager}.hpp
//   include/model/{AgeSEPAIHRDModel,AgeSEPAIHRDsimulator,PieceWiseConstantNPIStrategy,SEPAIHRDModelCalibration}.hpp
//   include/model/objectives/SEPAIHRDObjectiveFunction.hpp, include/utils/GetCalibrationData.hpp
#pragma once

#include <cstdint>
#include <atomic>
#include <functional>
#include <mutex>
#include <limits>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "linalg.hpp"
#include "sepaihrd_b200.h"

namespace epidemic {

// ---- exceptions (include/exceptions/Exceptions.hpp) ------------------------------------------------------
class ModelException : public std::runtime_error {
public:
    ModelException(const std::string& source, const std::string& message)
        : std::runtime_error("[" + source + "] " + message) {}
};
class InvalidParameterException : public ModelException { using ModelException::ModelException; };
class SimulationException : public ModelException { using ModelException::ModelException; };
class ModelConstructionException : public ModelException { using ModelException::ModelException; };
class InvalidResultException : public ModelException { using ModelException::ModelException; };
class OutOfRangeException : public ModelException { using ModelException::ModelException; };

// ---- SEPAIHRDParameters (include/model/parameters/SEPAIHRDParameters.hpp:12-115) ---------------------------
struct SEPAIHRDParameters {
    VectorXd N;
    MatrixXd M_baseline;
    double contact_matrix_scaling_factor = 1.0;
    double beta = std::numeric_limits<double>::quiet_NaN();   // quirk Q1: never read when a beta schedule exists
    std::vector<double> beta_end_times, beta_values;
    VectorXd a, h_infec;
    double theta = 0, sigma = 0, gamma_p = 0, gamma_A = 0, gamma_I = 0, gamma_H = 0, gamma_ICU = 0;
    VectorXd p, h, icu, d_H, d_ICU;
    std::vector<double> kappa_end_times, kappa_values;
    double E0_multiplier = 1, P0_multiplier = 1, A0_multiplier = 1, I0_multiplier = 1, H0_multiplier = 1,
           ICU0_multiplier = 1, R0_multiplier = 1, D0_multiplier = 1;
    double runup_days = 30.0, seed_exposed = 10.0;
    VectorXd d_community;
    bool validate() const;
};

// ---- NPI strategy (include/model/PieceWiseConstantNPIStrategy.hpp) ----------------------------------------
class INpiStrategy {
public:
    virtual ~INpiStrategy() = default;
    virtual double getReductionFactor(double time) const = 0;
    virtual const std::vector<double>& getEndTimes() const = 0;
    virtual std::vector<double> getValues() const = 0;
    virtual void setValues(const std::vector<double>& new_values) = 0;
    virtual std::shared_ptr<INpiStrategy> clone() const = 0;
};

class PiecewiseConstantNpiStrategy : public INpiStrategy {
public:
    PiecewiseConstantNpiStrategy(const std::vector<double>& npi_end_times_after_baseline,
                                 const std::vector<double>& npi_values_after_baseline,
                                 const std::map<std::string, std::pair<double, double>>& param_specific_bounds = {},
                                 double baseline_kappa = 1.0, double baseline_period_end_time = 0.0,
                                 bool fixed_baseline = true,
                                 const std::vector<std::string>& param_names_for_npi_values = {});
    double getReductionFactor(double time) const override;          // src/model/PieceWiseConstantNPIStrategy.cpp:86-127
    const std::vector<double>& getEndTimes() const override { return npi_period_end_times_; }
    std::vector<double> getValues() const override;
    void setValues(const std::vector<double>& new_values) override;
    std::shared_ptr<INpiStrategy> clone() const override { return std::make_shared<PiecewiseConstantNpiStrategy>(*this); }
    double getBaselineKappa() const { return baseline_kappa_value_; }
    double getBaselinePeriodEndTime() const { return baseline_period_end_time_; }
    bool isBaselineFixed() const { return is_baseline_fixed_; }
    size_t getNumCalibratableNpiParams() const;
    std::string getNpiParamName(int calibratable_idx) const;
    void setCalibratableValues(const std::vector<double>& calibratable_values);   // .cpp:228-262 (throws on negatives)
    std::vector<double> getCalibratableValues() const;
    double getLowerBoundForParamIndex(int calibratable_idx) const;
    double getUpperBoundForParamIndex(int calibratable_idx) const;

private:
    double baseline_kappa_value_;
    bool is_baseline_fixed_;
    double baseline_period_end_time_;
    std::vector<double> npi_period_end_times_, npi_kappa_values_;
    std::vector<std::string> npi_param_names_;
    std::map<std::string, std::pair<double, double>> param_bounds_map_;
};

// ---- AgeSEPAIHRDModel (include/model/AgeSEPAIHRDModel.hpp) -------------------------------------------------
// Holds the model parameters and the NPI schedule.  computeDerivatives lives on the device
// (csrc/sepaihrd_kernels.cuh, rhs<>); this class is what the parameter manager mutates and what the
// evaluator snapshots into the C-ABI problem description.
class AgeSEPAIHRDModel {
public:
    AgeSEPAIHRDModel(const SEPAIHRDParameters& params, std::shared_ptr<INpiStrategy> npi_strategy_ptr);
    std::shared_ptr<AgeSEPAIHRDModel> clone() const;
    int getNumAgeClasses() const { return static_cast<int>(params_.N.size()); }
    int getStateSize() const { return SEPAIHRD_NUM_COMPARTMENTS * getNumAgeClasses(); }
    std::vector<std::string> getStateNames() const;
    SEPAIHRDParameters getModelParameters() const;                      // src/model/AgeSEPAIHRDModel.cpp:294-323
    void setModelParameters(const SEPAIHRDParameters& params);           // .cpp:325-363
    std::shared_ptr<INpiStrategy> getNpiStrategy() const { return npi_strategy_; }
    const VectorXd& getPopulationSizes() const { return params_.N; }
    const MatrixXd& getContactMatrix() const { return params_.M_baseline; }
    double computeBeta(double time) const;                              // .cpp:366-368 + PiecewiseConstantParameterStrategy.cpp:37-74

    // Flat image of the parameters the hot path reads, in the slot order of sepaihrd_b200.h.
    std::vector<double> slotVector() const;
    std::vector<double> kappaEndTimes() const;      // [baseline end, NPI period ends...]
    int numKappa() const;

private:
    SEPAIHRDParameters params_;
    std::shared_ptr<INpiStrategy> npi_strategy_;
};

// ---- CalibrationData (include/utils/GetCalibrationData.hpp) ------------------------------------------------
// In-memory backend only (the reference's matrix constructor, src/utils/GetCalibrationData.cpp:24-89); the CSV
// reader is a Python-side fixture tool (config.py).  Matrices are (day, age).
class CalibrationData {
public:
    CalibrationData(const MatrixXd& new_hospitalizations, const MatrixXd& new_icu, const MatrixXd& new_deaths,
                    const VectorXd& population_by_age, const VectorXd& initial_sepaihrd_state = VectorXd());
    const MatrixXd& getNewHospitalizations() const { return new_hospitalizations_; }
    const MatrixXd& getNewICU() const { return new_icu_; }
    const MatrixXd& getNewDeaths() const { return new_deaths_; }
    const VectorXd& getPopulationByAgeGroup() const { return population_; }
    const VectorXd& getInitialSEPAIHRDState() const { return initial_state_; }
    int getNumDataPoints() const { return static_cast<int>(new_hospitalizations_.rows()); }
    int getNumAgeClasses() const { return static_cast<int>(population_.size()); }

private:
    MatrixXd new_hospitalizations_, new_icu_, new_deaths_;
    VectorXd population_, initial_state_;
};

// ---- SimulationResult (include/sir_age_structured/SimulationResult.hpp:18-35) ------------------------------
struct SimulationResult {
    std::vector<double> time_points;
    std::vector<std::vector<double>> solution;     // [time][state], compartment-major state
    std::vector<std::string> compartment_names;
    int num_age_classes = 0;
    bool isValid() const { return !time_points.empty() && time_points.size() == solution.size(); }
};

// ---- interfaces (include/sir_age_structured/interfaces/*.hpp) ---------------------------------------------
using state_type = std::vector<double>;

class IObjectiveFunction {
public:
    virtual ~IObjectiveFunction() = default;
    virtual double calculate(const VectorXd& parameters) const = 0;                       // IObjectiveFunction.hpp:24
    virtual const std::vector<std::string>& getParameterNames() const = 0;                 // :30
    // B200 addition: B evaluations in one call.  params is row-major [B][ld], ld >= parameter count.
    // The default loops over calculate(); SEPAIHRDObjectiveFunction overrides it with ONE kernel launch.
    virtual void calculateBatch(const double* params, int64_t B, int64_t ld, double* out) const;
};

// include/model/interfaces/IGradientObjectiveFunction.hpp: objective + gradient in one call (the NUTS sampler needs it)
class IGradientObjectiveFunction : public virtual IObjectiveFunction {
public:
    ~IGradientObjectiveFunction() override = default;
    virtual double evaluate_with_gradient(const VectorXd& params, VectorXd& grad) const = 0;
};

// Forward finite differences the way SEPAIHRDGradientObjectiveFunction::evaluate_with_gradient forms them
// (src/model/objectives/SEPAIHRDGradientObjectiveFunction.cpp:15-171): step eps_i = epsilon * max(|x_i|, epsilon),
// grad_i = (f(x + eps_i e_i) - f(x)) / eps_i when f(x + eps_i e_i) is finite, else 0.  `rows` receives the P perturbed vectors
// ([P][P] row-major) and is what a caller evaluates as ONE batch.
struct ForwardDifferences {
    static void perturb(const VectorXd& params, double epsilon, std::vector<double>& rows, std::vector<double>& steps);
    static void gradient(double f_center, const double* f_plus, const uint8_t* skip /* or null */, const std::vector<double>& steps, VectorXd& grad);
};

class IParameterManager {
public:
    virtual ~IParameterManager() = default;
    virtual VectorXd getCurrentParameters() const = 0;
    virtual void updateModelParameters(const VectorXd& parameters) = 0;
    virtual const std::vector<std::string>& getParameterNames() const = 0;
    virtual size_t getParameterCount() const = 0;
    virtual double getSigmaForParamIndex(int index) const = 0;
    virtual VectorXd applyConstraints(const VectorXd& parameters) const = 0;
    virtual int getIndexForParam(const std::string& name) const = 0;
    virtual double getLowerBoundForParamIndex(int idx) const = 0;
    virtual double getUpperBoundForParamIndex(int idx) const = 0;
};

class IOdeSolverStrategy {
public:
    virtual ~IOdeSolverStrategy() = default;
    virtual void integrate(const std::function<void(const state_type&, state_type&, double)>& system, state_type& initial_state,
                           const std::vector<double>& times, double dt_hint,
                           std::function<void(const state_type&, double)> observer, double abs_error,
                           double rel_error) const = 0;                                      // IOdeSolverStrategy.hpp:35-42
};

// Tag for "adaptive Dopri5 with Boost.Odeint's controller": the only stepper the device kernels implement.
// AgeSEPAIHRDSimulator recognises it and integrates on the GPU; integrate() on an arbitrary std::function system
// cannot run on the device and throws SimulationException instead of silently falling back to the CPU.
class Dopri5SolverStrategy : public IOdeSolverStrategy {
public:
    void integrate(const std::function<void(const state_type&, state_type&, double)>&, state_type&, const std::vector<double>&,
                   double, std::function<void(const state_type&, double)>, double, double) const override;
};

class ISimulationCache {
public:
    virtual ~ISimulationCache() = default;
    virtual std::optional<double> get(const VectorXd& parameters) = 0;
    virtual void set(const VectorXd& parameters, double result) = 0;
    virtual void clear() = 0;
    virtual size_t size() const = 0;
    virtual std::string createCacheKey(const VectorXd& parameters) const = 0;
    virtual bool getLikelihood(const std::string& key, double& value) = 0;
    virtual void storeLikelihood(const std::string& key, double value) = 0;
};

// The cache that caches nothing (pattern: sepaihrd_objective_benchmark_main.cpp:229-238).  Parity runs use it: a hit of the
// reference's SimulationCache may return the likelihood of ANOTHER vector (1e-8 quantisation / hash collision, SURVEY.md
// quirk Q6).
class NullSimulationCache : public ISimulationCache {
public:
    std::optional<double> get(const VectorXd&) override { return std::nullopt; }
    void set(const VectorXd&, double) override {}
    void clear() override {}
    size_t size() const override { return 0; }
    std::string createCacheKey(const VectorXd&) const override { return {}; }
    bool getLikelihood(const std::string&, double&) override { return false; }
    void storeLikelihood(const std::string&, double) override {}
};

// SimulationCache (src/sir_age_structured/caching/SimulationCache.cpp): parameter vector -> log-likelihood, keyed by a 64-bit
// hash of the vector quantised at 1e-8 (computeHash, .cpp:35-52), open addressing with linear probing over `max_size` slots,
// least-frequently-used eviction with least-recently-used tie-break by a linear scan (.cpp:74-104), one mutex, hit counters.
// SEPAIHRDObjectiveFunction consults it in calculate() and in calculateBatch() for batches that fit into it.
class SimulationCache : public ISimulationCache {
public:
    explicit SimulationCache(size_t max_size = 1000);
    std::optional<double> get(const VectorXd& parameters) override;
    void set(const VectorXd& parameters, double result) override;
    void clear() override;
    size_t size() const override;
    std::string createCacheKey(const VectorXd& parameters) const override;      // the hash as a decimal string
    size_t computeHash(const VectorXd& parameters) const;
    size_t computeHash(const double* parameters, std::ptrdiff_t count) const;
    bool getLikelihood(const std::string& key, double& value) override;
    void storeLikelihood(const std::string& key, double value) override;
    bool getLikelihood(size_t key, double& value);                              // fast path: numeric key
    void storeLikelihood(size_t key, double value);
    size_t capacity() const { return capacity_; }
    size_t getLikelihoodCalls() const { return get_calls_.load(std::memory_order_relaxed); }
    size_t getLikelihoodHits() const { return get_hits_.load(std::memory_order_relaxed); }
    size_t storeLikelihoodCalls() const { return store_calls_.load(std::memory_order_relaxed); }
    void resetLikelihoodStats() { get_calls_ = 0; get_hits_ = 0; store_calls_ = 0; }

private:
    static constexpr size_t NOT_FOUND = static_cast<size_t>(-1);
    size_t findIndex(size_t key) const;
    size_t evict();
    bool lookupLocked(size_t key, double& value);
    void storeLocked(size_t key, double value);
    size_t capacity_, count_ = 0;
    uint32_t current_tick_ = 0;
    std::vector<size_t> keys_;
    std::vector<double> values_;
    std::vector<uint32_t> frequencies_, timestamps_;
    std::vector<uint8_t> occupied_;
    mutable std::mutex mutex_;
    mutable std::atomic<size_t> get_calls_{0}, get_hits_{0}, store_calls_{0};
};

// B calls of calculate() as far as `cache` is concerned (probe per vector, ONE evaluate() call for the distinct misses, store
// what calculate() stores: rows whose status word is 0 or SEPAIHRD_ST_NONFINITE); batches larger than a SimulationCache's
// capacity and the NullSimulationCache go to evaluate() whole.  evaluate(rows, M, ld, values, status).
void evaluateThroughCache(ISimulationCache& cache, std::ptrdiff_t P, const double* params, int64_t B, int64_t ld, double* out,
                          const std::function<void(const double*, int64_t, int64_t, double*, uint32_t*)>& evaluate);

struct OptimizationResult {                                          // IOptimizationAlgorithm.hpp:18-27
    VectorXd bestParameters;
    double bestObjectiveValue = -std::numeric_limits<double>::infinity();
    std::vector<VectorXd> samples;
    std::vector<double> sampleObjectiveValues;
    std::map<std::string, double> additionalStats;
    MatrixXd finalCovariance;
};

class IOptimizationAlgorithm {
public:
    virtual ~IOptimizationAlgorithm() = default;
    virtual OptimizationResult optimize(const VectorXd& initialParameters, IObjectiveFunction& objectiveFunction,
                                        IParameterManager& parameterManager) = 0;           // :44-49
    virtual void configure(const std::map<std::string, double>& settings) = 0;              // :53
};

// ---- SEPAIHRDParameterManager (include/model/parameters/SEPAIHRDParameterManager.hpp) ---------------------
enum class ConstraintMode { OPTIMIZATION_CLAMP, MCMC_REFLECT };

class SEPAIHRDParameterManager : public IParameterManager {
public:
    SEPAIHRDParameterManager(std::shared_ptr<AgeSEPAIHRDModel> model, const std::vector<std::string>& params_to_calibrate,
                             const std::map<std::string, double>& proposal_sigmas,
                             const std::map<std::string, std::pair<double, double>>& param_bounds);
    VectorXd getCurrentParameters() const override;                                          // .cpp:91-158
    void updateModelParameters(const VectorXd& parameters) override;                         // .cpp:160-162
    void updateModelParameters(const VectorXd& parameters, std::shared_ptr<AgeSEPAIHRDModel> target_model);   // .cpp:164-287
    const std::vector<std::string>& getParameterNames() const override { return param_names_; }
    size_t getParameterCount() const override { return param_names_.size(); }
    double getSigmaForParamIndex(int index) const override;
    VectorXd applyConstraints(const VectorXd& parameters) const override;                    // .cpp:315-347
    int getIndexForParam(const std::string& name) const override;
    double getLowerBoundForParamIndex(int idx) const override;
    double getUpperBoundForParamIndex(int idx) const override;
    const std::map<std::string, double>& getProposalSigmas() const { return proposal_sigmas_; }
    const std::map<std::string, std::pair<double, double>>& getParamBounds() const { return param_bounds_; }
    void setConstraintMode(ConstraintMode mode);
    ConstraintMode getConstraintMode() const { return mode_; }
    std::shared_ptr<AgeSEPAIHRDModel> getModel() const { return model_; }
    // slot (sepaihrd_b200.h layout) of every calibrated parameter, resolved once at construction
    const std::vector<int32_t>& slots() const { return slots_; }
    // observers of the constraint mode (the device context of an objective follows the manager's mode)
    int onConstraintModeChange(std::function<void(ConstraintMode)> cb) { mode_listeners_.push_back(std::move(cb)); return static_cast<int>(mode_listeners_.size()) - 1; }
    void removeConstraintModeListener(int id) { if (id >= 0 && static_cast<size_t>(id) < mode_listeners_.size()) mode_listeners_[static_cast<size_t>(id)] = nullptr; }
    static double reflectBound(double value, double min_b, double max_b);                    // .cpp:302-313

private:
    std::shared_ptr<AgeSEPAIHRDModel> model_;
    ConstraintMode mode_ = ConstraintMode::OPTIMIZATION_CLAMP;
    std::vector<std::string> param_names_;
    std::map<std::string, double> proposal_sigmas_;
    std::map<std::string, std::pair<double, double>> param_bounds_;
    std::vector<double> lower_, upper_;       // param_bounds_ by parameter index
    std::vector<int32_t> slots_;
    std::vector<std::function<void(ConstraintMode)>> mode_listeners_;
};

// ---- device context: RAII over sepaihrd_ctx ---------------------------------------------------------------
// Snapshots (model, NPI schedule, data, time grid, calibrated names + bounds, tolerances) into a
// sepaihrd_problem and creates the device evaluator.  Throws ModelConstructionException when the CUDA library
// reports no usable device: there is no CPU fallback.
class DeviceContext {
public:
    DeviceContext(const AgeSEPAIHRDModel& model, const std::vector<double>& time_points, const MatrixXd* obs_hosp,
                  const MatrixXd* obs_icu, const MatrixXd* obs_deaths, const VectorXd& data_initial_state,
                  const std::vector<int32_t>& param_slots, const std::vector<double>& lower, const std::vector<double>& upper,
                  ConstraintMode mode, double abs_error, double rel_error, double dt_hint, int device = -1);
    ~DeviceContext();
    DeviceContext(const DeviceContext&) = delete;
    DeviceContext& operator=(const DeviceContext&) = delete;
    sepaihrd_ctx* get() const { return ctx_; }
    int numParams() const { return n_params_; }
    int numTimes() const { return n_times_; }
    int stateSize() const { return state_size_; }
    void setConstraintMode(ConstraintMode mode);

private:
    sepaihrd_ctx* ctx_ = nullptr;
    int n_params_ = 0, n_times_ = 0, state_size_ = 0;
};

// ---- AgeSEPAIHRDSimulator (include/model/AgeSEPAIHRDsimulator.hpp:33-39, Simulator.cpp:60-150) -------------
class AgeSEPAIHRDSimulator {
public:
    AgeSEPAIHRDSimulator(std::shared_ptr<AgeSEPAIHRDModel> model, std::shared_ptr<IOdeSolverStrategy> solver_strategy,
                         double start_time, double end_time, double time_step, double abs_error = 1.0e-6,
                         double rel_error = 1.0e-6);
    ~AgeSEPAIHRDSimulator();
    SimulationResult run(const VectorXd& initial_state, const std::vector<double>& output_time_points);
    // B200 addition: the same run for B parameterisations of the model at once (posterior-predictive draws):
    // slot_rows is [B][slot_count] in the layout of sepaihrd_b200.h (AgeSEPAIHRDModel::slotVector()).
    // out is [B][times][11 n]; returns per-draw status words.
    std::vector<uint32_t> runBatch(const VectorXd& initial_state, const std::vector<double>& output_time_points,
                                   const double* slot_rows, int64_t B, double* out);
    std::shared_ptr<AgeSEPAIHRDModel> getModel() const { return model_; }

private:
    void ensureContext(const std::vector<double>& times);
    std::shared_ptr<AgeSEPAIHRDModel> model_;
    std::shared_ptr<IOdeSolverStrategy> solver_;
    double start_time_, end_time_, time_step_, abs_err_, rel_err_;
    std::unique_ptr<DeviceContext> dev_;
    std::vector<double> dev_times_;
};

// ---- SEPAIHRDObjectiveFunction (include/model/objectives/SEPAIHRDObjectiveFunction.hpp:40-57) --------------
class SEPAIHRDObjectiveFunction : public virtual IObjectiveFunction {
public:
    SEPAIHRDObjectiveFunction(std::shared_ptr<AgeSEPAIHRDModel> model, IParameterManager& parameterManager,
                              ISimulationCache& cache, const CalibrationData& calibration_data,
                              const std::vector<double>& time_points, const VectorXd& initial_state,
                              std::shared_ptr<IOdeSolverStrategy> solver_strategy, double abs_error = 1.0e-6,
                              double rel_error = 1.0e-6);
    ~SEPAIHRDObjectiveFunction() override;
    double calculate(const VectorXd& parameters) const override;                             // .cpp:62-235
    // B calls of calculate() in order.  The cache is consulted like calculate() does (.cpp:63-77, :227-234) when the batch fits
    // into it (B <= capacity: a larger batch would evict its own entries; it goes to the device whole); a key that repeats
    // inside the batch is evaluated once.  Failures that calculate() returns before its store (invalid parameters, S overflow,
    // failed integration) are not stored here either.
    void calculateBatch(const double* params, int64_t B, int64_t ld, double* out) const override;
    // batch form with per-set status words (SEPAIHRD_ST_*) and optional accepted/rejected step counts; never cached
    void calculateBatch(const double* params, int64_t B, int64_t ld, double* out, uint32_t* status, int32_t* steps) const;
    const std::vector<std::string>& getParameterNames() const override;
    DeviceContext& device() const { return *dev_; }

protected:
    void evaluateRows(const double* params, int64_t B, int64_t ld, double* out, uint32_t* status, int32_t* steps) const;
    IParameterManager& parameterManager_;
    ISimulationCache& cache_;
    std::shared_ptr<AgeSEPAIHRDModel> model_;
    std::unique_ptr<DeviceContext> dev_;
    int mode_listener_id_ = -1;
};

// SEPAIHRDGradientObjectiveFunction (include/model/objectives/SEPAIHRDGradientObjectiveFunction.hpp): the reference loops over
// the P parameters with OpenMP, one simulation each; here the P perturbed vectors are ONE device batch (SURVEY.md section 2,
// row 15: "P+1 batch").  Like the reference, the centre value goes through calculate() (and its cache) and the perturbed
// vectors are constrained by CLAMPING whatever the manager's current mode is (the reference builds a fresh
// SEPAIHRDParameterManager per component, whose mode is the default OPTIMIZATION_CLAMP, .cpp:40-43).
// Deliberate difference: the reference builds the perturbed runs' initial state from the multipliers only
// and ignores run-up seeding, so under the shipped configuration its differences compare two different initial conditions;
// here every evaluation uses calculate()'s initial-state rule.
class SEPAIHRDGradientObjectiveFunction : public virtual SEPAIHRDObjectiveFunction, public IGradientObjectiveFunction {
public:
    SEPAIHRDGradientObjectiveFunction(std::shared_ptr<AgeSEPAIHRDModel> model, IParameterManager& parameterManager, ISimulationCache& cache,
                                      const CalibrationData& calibration_data, const std::vector<double>& time_points,
                                      const VectorXd& initial_state, std::shared_ptr<IOdeSolverStrategy> solver_strategy,
                                      double abs_error = 1.0e-6, double rel_error = 1.0e-6)
        : SEPAIHRDObjectiveFunction(std::move(model), parameterManager, cache, calibration_data, time_points, initial_state,
                                    std::move(solver_strategy), abs_error, rel_error) {}
    double epsilon_ = 1e-4;
    double evaluate_with_gradient(const VectorXd& params, VectorXd& grad) const override;
    long gradientBatches() const { return gradient_batches_; }

private:
    mutable long gradient_batches_ = 0;
};

// ---- posterior-predictive aggregation (include/model/ResultAggregator.hpp, PostCalibrationAnalyser) ---------------
struct PosteriorPredictiveData {
    struct IncidenceData { MatrixXd median, lower_90, upper_90, lower_95, upper_95, observed; };    // (day, age)
    std::vector<double> time_points;                       // the output days t >= 0
    IncidenceData daily_hospitalizations, daily_icu_admissions, daily_deaths;
    IncidenceData cumulative_hospitalizations, cumulative_icu_admissions, cumulative_deaths;
    int64_t samples_used = 0;
};

class ResultAggregator {
public:
    // ResultAggregator::aggregatePosteriorPredictives (src/model/ResultAggregator.cpp:174-412) without the runner /
    // metrics-calculator indirection: the selected samples go to the device in ONE call (sepaihrd_posterior_predictive).
    // num_samples_for_ppc > 0 and < samples: that many draws WITH replacement from std::mt19937(random_seed) +
    // uniform_int_distribution, like the reference (.cpp:262-272); otherwise every sample once.  Quantiles are exact
    // sample quantiles, not Boost's P-square estimates (known deviations).
    PosteriorPredictiveData aggregatePosteriorPredictives(const std::vector<VectorXd>& param_samples, SEPAIHRDParameterManager& param_manager,
                                                          int num_samples_for_ppc, const std::vector<double>& time_points,
                                                          const VectorXd& initial_state, const CalibrationData& observed_data,
                                                          std::shared_ptr<AgeSEPAIHRDModel> model_template, unsigned int random_seed = 0) const;
};

}  // namespace epidemic
