// analysis.cpp -- NPI scenario comparison and per-run essential metrics over the device simulator.  See analysis.hpp.
#include "analysis.hpp"

#include <algorithm>
#include <cmath>
#include <fstream>
#include <iomanip>
#include <limits>

namespace epidemic {

namespace {

constexpr int kE = 1, kP = 2, kA = 3, kI = 4, kH = 5, kICU = 6, kR = 7, kD = 8, kCumH = 9, kCumICU = 10;

double block_sum(const std::vector<double>& state, int comp, int n) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += state[static_cast<size_t>(comp * n + i)];
    return s;
}

// q-quantile of a sorted sample by linear interpolation at position q (m - 1)  (PostCalibrationAnalyser.cpp:312-320)
double sorted_quantile(const std::vector<double>& v, double q) {
    const double pos = q * static_cast<double>(v.size() - 1);
    const size_t idx = static_cast<size_t>(pos);
    const double frac = pos - static_cast<double>(idx);
    if (idx + 1 < v.size()) return v[idx] * (1.0 - frac) + v[idx + 1] * frac;
    return v[idx];
}

}  // namespace

// ---- reproduction numbers -------------------------------------------------------------------------------------
ReproductionNumberCalculator::ReproductionNumberCalculator(std::shared_ptr<AgeSEPAIHRDModel> model) : model_(std::move(model)) {
    if (!model_) throw InvalidParameterException("ReproductionNumberCalculator", "Model pointer cannot be null.");
}

MatrixXd ReproductionNumberCalculator::nextGenerationBlock(const VectorXd& X, double time, bool clamp_negative) const {
    const SEPAIHRDParameters p = model_->getModelParameters();
    const int n = model_->getNumAgeClasses();
    if (X.size() != n) throw InvalidParameterException("ReproductionNumberCalculator", "S_current vector size mismatch.");
    const double kappa = model_->getNpiStrategy()->getReductionFactor(time);
    if (kappa < 0) throw ModelException("ReproductionNumberCalculator", "NPI reduction factor cannot be negative.");
    const double beta = model_->computeBeta(time);
    MatrixXd K = MatrixXd::Zero(n, n);
    for (int j = 0; j < n; ++j) {
        if (p.N(j) < 1e-9) continue;
        // expected infectious "weight" of one newly exposed individual of class j: time in P, in A, and theta x time in I
        const double w = 1.0 / p.gamma_p + p.p(j) / p.gamma_A + p.theta * (1.0 - p.p(j)) / (p.gamma_I + p.h(j));
        for (int i = 0; i < n; ++i) {
            double T = beta * kappa * p.M_baseline(i, j) * p.a(i) * p.h_infec(j) * (X(i) / p.N(j));
            if (clamp_negative) T = std::max(0.0, T);
            K(i, j) = T * w;
        }
    }
    return K;
}

double ReproductionNumberCalculator::spectralRadiusNonNegative(const MatrixXd& K) {
    const std::ptrdiff_t n = K.rows();
    VectorXd v = VectorXd::Constant(n, 1.0 / static_cast<double>(n));
    double lambda = 0.0;
    for (int it = 0; it < 20000; ++it) {
        const VectorXd w = K * v;
        const double norm = w.sum();                  // entries are non-negative: the 1-norm
        if (!(norm > 0.0) || !std::isfinite(norm)) return std::isfinite(norm) ? 0.0 : norm;
        VectorXd next = w;
        next /= norm;
        double change = 0.0;
        for (std::ptrdiff_t i = 0; i < n; ++i) change = std::max(change, std::fabs(next(i) - v(i)));
        v = next;
        const bool settled = std::fabs(norm - lambda) <= 4e-16 * norm && change <= 1e-15;
        lambda = norm;
        if (settled) break;
    }
    return lambda;
}

double ReproductionNumberCalculator::calculateR0() const {
    return spectralRadiusNonNegative(nextGenerationBlock(model_->getPopulationSizes(), 0.0, false));
}
double ReproductionNumberCalculator::calculateRt(const VectorXd& S_current, double time) const {
    return spectralRadiusNonNegative(nextGenerationBlock(S_current, time, true));
}

// ---- metrics ----------------------------------------------------------------------------------------------------
EssentialMetrics MetricsCalculator::calculateEssentialMetrics(const SimulationResult& sim, std::shared_ptr<AgeSEPAIHRDModel> model,
                                                              const SEPAIHRDParameters& params, const VectorXd& initial_state,
                                                              const std::vector<double>& time_points) const {
    EssentialMetrics m;
    const int n = static_cast<int>(params.N.size());
    m.age_specific_IFR.assign(static_cast<size_t>(n), 0.0);
    m.age_specific_IHR.assign(static_cast<size_t>(n), 0.0);
    m.age_specific_IICUR.assign(static_cast<size_t>(n), 0.0);
    m.age_specific_attack_rate.assign(static_cast<size_t>(n), 0.0);
    if (!sim.isValid()) return m;

    ReproductionNumberCalculator rn(model);
    m.R0 = rn.calculateR0();

    // everyone who is not susceptible at the start has been infected already
    std::vector<double> cum_inf(static_cast<size_t>(n), 0.0);
    for (int i = 0; i < n; ++i)
        for (int c = kE; c <= kR; ++c) cum_inf[static_cast<size_t>(i)] += initial_state(c * n + i);
    const double total_population = params.N.sum();

    size_t target_idx = 0;                      // grid point closest to day 64 (first one on ties)
    double best = std::numeric_limits<double>::max();
    for (size_t t = 0; t < time_points.size(); ++t) {
        const double diff = std::fabs(time_points[t] - 64.0);
        if (diff < best) { best = diff; target_idx = t; }
    }

    VectorXd S(n);
    std::vector<double> load(static_cast<size_t>(n));
    for (size_t t = 0; t < time_points.size(); ++t) {
        const std::vector<double>& x = sim.solution[t];
        const double time_t = time_points[t];
        const double dt = (t > 0) ? time_t - time_points[t - 1] : 1.0;
        for (int i = 0; i < n; ++i) S(i) = x[static_cast<size_t>(i)];

        const double Rt = rn.calculateRt(S, time_t);
        m.max_Rt = std::max(m.max_Rt, Rt);
        m.min_Rt = std::min(m.min_Rt, Rt);
        if (t + 1 == time_points.size()) m.final_Rt = Rt;

        const double total_H = block_sum(x, kH, n), total_ICU = block_sum(x, kICU, n);
        if (total_H > m.peak_hospital_occupancy) { m.peak_hospital_occupancy = total_H; m.time_to_peak_hospital = time_t; }
        if (total_ICU > m.peak_ICU_occupancy) { m.peak_ICU_occupancy = total_ICU; m.time_to_peak_ICU = time_t; }

        // new infections of the interval by a forward-Euler sum of lambda S dt.  The reference multiplies by the SCALAR
        // params.beta here (MetricsCalculator.cpp:113), which its own configuration never sets once a beta schedule exists
        // (quirk Q1: an uninitialised double); a non-finite scalar falls back to the schedule value beta(t).
        const double kappa_t = model->getNpiStrategy()->getReductionFactor(time_t);
        const double beta_t = std::isfinite(params.beta) ? params.beta : model->computeBeta(time_t);
        for (int j = 0; j < n; ++j) {
            const size_t k = static_cast<size_t>(j);
            load[k] = (params.N(j) > 1e-9) ? (x[static_cast<size_t>(kP * n + j)] + x[static_cast<size_t>(kA * n + j)] +
                                              params.theta * x[static_cast<size_t>(kI * n + j)]) / params.N(j)
                                           : 0.0;
        }
        const double scale = beta_t * kappa_t;
        for (int i = 0; i < n; ++i) {
            double lambda = 0.0;
            for (int j = 0; j < n; ++j) lambda += (scale * params.M_baseline(i, j)) * load[static_cast<size_t>(j)];
            cum_inf[static_cast<size_t>(i)] += lambda * S(i) * dt;
        }
        if (t == target_idx) m.seroprevalence_at_target_day = (total_population - S.sum()) / total_population;
    }

    const std::vector<double>& last = sim.solution.back();
    double deaths = 0.0, infections = 0.0;
    for (int i = 0; i < n; ++i) {
        const size_t k = static_cast<size_t>(i);
        const double d = last[static_cast<size_t>(kD * n + i)] - initial_state(kD * n + i);
        const double ch = last[static_cast<size_t>(kCumH * n + i)] - initial_state(kCumH * n + i);
        const double cu = last[static_cast<size_t>(kCumICU * n + i)] - initial_state(kCumICU * n + i);
        deaths += d;
        infections += cum_inf[k];
        m.age_specific_attack_rate[k] = (params.N(i) > 0) ? cum_inf[k] / params.N(i) : 0.0;
        if (cum_inf[k] > 1.0) {                 // ratios only where at least one infection backs them, clamped to [0, 1]
            m.age_specific_IFR[k] = std::max(0.0, std::min(d / cum_inf[k], 1.0));
            m.age_specific_IHR[k] = std::max(0.0, std::min(ch / cum_inf[k], 1.0));
            m.age_specific_IICUR[k] = std::max(0.0, std::min(cu / cum_inf[k], 1.0));
        }
    }
    m.total_cumulative_deaths = deaths;
    m.overall_attack_rate = infections / total_population;
    m.overall_IFR = (infections > 1e-9) ? deaths / infections : 0.0;
    for (size_t i = 0; i < params.kappa_values.size(); ++i) m.kappa_values["kappa_" + std::to_string(i + 1)] = params.kappa_values[i];
    return m;
}

std::vector<double> MetricsCalculator::calculateRtTrajectory(const SimulationResult& sim, std::shared_ptr<AgeSEPAIHRDModel> model,
                                                             const std::vector<double>& time_points) const {
    std::vector<double> rt;
    if (!sim.isValid()) return rt;
    ReproductionNumberCalculator rn(model);
    const int n = model->getNumAgeClasses();
    VectorXd S(n);
    for (size_t t = 0; t < time_points.size(); ++t) {
        for (int i = 0; i < n; ++i) S(i) = sim.solution[t][static_cast<size_t>(i)];
        rt.push_back(rn.calculateRt(S, time_points[t]));
    }
    return rt;
}

std::vector<double> MetricsCalculator::calculateSeroprevalenceTrajectory(const SimulationResult& sim, const SEPAIHRDParameters& params,
                                                                         const std::vector<double>& time_points) const {
    std::vector<double> sero;
    if (!sim.isValid()) return sero;
    const int n = static_cast<int>(params.N.size());
    const double total = params.N.sum();
    for (size_t t = 0; t < time_points.size(); ++t) sero.push_back((total - block_sum(sim.solution[t], 0, n)) / total);
    return sero;
}

// ---- runner -------------------------------------------------------------------------------------------------------
SimulationRunner::SimulationRunner(std::shared_ptr<AgeSEPAIHRDModel> model_template, std::shared_ptr<IOdeSolverStrategy> solver,
                                   double abs_error, double rel_error, double dt_hint)
    : model_template_(std::move(model_template)), solver_(std::move(solver)), abs_err_(abs_error), rel_err_(rel_error), dt_hint_(dt_hint) {
    if (!model_template_) throw InvalidParameterException("SimulationRunner", "Model template cannot be null");
    if (!solver_) throw InvalidParameterException("SimulationRunner", "Solver strategy cannot be null");
}
SimulationRunner::~SimulationRunner() = default;

std::shared_ptr<AgeSEPAIHRDModel> SimulationRunner::modelFor(const SEPAIHRDParameters& params) const {
    auto npi = model_template_->getNpiStrategy()->clone();
    if (auto pw = std::dynamic_pointer_cast<PiecewiseConstantNpiStrategy>(npi)) {
        const size_t expected = pw->getNumCalibratableNpiParams();
        const std::vector<double>& kv = params.kappa_values;
        std::vector<double> calibratable;
        if (expected > 0) {
            if (kv.size() == expected) calibratable = kv;
            else if (pw->isBaselineFixed() && kv.size() == expected + 1) calibratable.assign(kv.begin() + 1, kv.end());
            else if (kv.size() >= expected) calibratable.assign(kv.end() - static_cast<std::ptrdiff_t>(expected), kv.end());
            if (calibratable.size() == expected) pw->setCalibratableValues(calibratable);     // throws on a negative kappa
        }
    }
    return std::make_shared<AgeSEPAIHRDModel>(params, npi);
}

std::vector<SimulationResult> SimulationRunner::runSimulations(const std::vector<SEPAIHRDParameters>& params, const VectorXd& initial_state,
                                                               const std::vector<double>& times) {
    if (params.empty()) return {};
    if (times.empty()) throw InvalidParameterException("SimulationRunner", "Time points cannot be empty");
    if (!simulator_ || times.front() != sim_start_ || times.back() != sim_end_) {      // one simulator per time window
        simulator_ = std::make_unique<AgeSEPAIHRDSimulator>(model_template_, solver_, times.front(), times.back(), dt_hint_, abs_err_, rel_err_);
        sim_start_ = times.front();
        sim_end_ = times.back();
    }
    std::vector<double> rows;
    size_t width = 0;
    for (const SEPAIHRDParameters& p : params) {
        const std::vector<double> row = modelFor(p)->slotVector();
        if (width == 0) width = row.size();
        if (row.size() != width) throw InvalidParameterException("SimulationRunner", "parameter sets of one batch must share the schedule lengths of the template");
        rows.insert(rows.end(), row.begin(), row.end());
    }
    const size_t B = params.size(), K = times.size(), W = static_cast<size_t>(model_template_->getStateSize());
    std::vector<double> out(B * K * W);
    const std::vector<uint32_t> status = simulator_->runBatch(initial_state, times, rows.data(), static_cast<int64_t>(B), out.data());
    std::vector<SimulationResult> results(B);
    for (size_t b = 0; b < B; ++b) {
        if (status[b] & SEPAIHRD_ST_STEP_FAILURE) throw SimulationException("SimulationRunner", "ODE integration failed: too many consecutive rejected steps.");
        SimulationResult& r = results[b];
        r.time_points = times;
        r.num_age_classes = model_template_->getNumAgeClasses();
        r.compartment_names = {"S", "E", "P", "A", "I", "H", "ICU", "R", "D", "CumH", "CumICU"};
        r.solution.resize(K);
        for (size_t k = 0; k < K; ++k) {
            const double* row = out.data() + (b * K + k) * W;
            r.solution[k].assign(row, row + W);
        }
    }
    return results;
}

SimulationResult SimulationRunner::runSimulation(const SEPAIHRDParameters& params, const VectorXd& initial_state, const std::vector<double>& times) {
    return runSimulations({params}, initial_state, times).front();
}

// ---- analyser -------------------------------------------------------------------------------------------------------
PostCalibrationAnalyser::PostCalibrationAnalyser(std::shared_ptr<AgeSEPAIHRDModel> model_template, std::shared_ptr<IOdeSolverStrategy> solver,
                                                 const std::vector<double>& time_points, const VectorXd& initial_state, double abs_error,
                                                 double rel_error)
    : model_template_(model_template), time_points_(time_points), initial_state_(initial_state),
      runner_(model_template, std::move(solver), abs_error, rel_error) {
    if (time_points_.empty()) throw InvalidParameterException("PostCalibrationAnalyser", "Time points cannot be empty");
    if (initial_state_.size() == 0) throw InvalidParameterException("PostCalibrationAnalyser", "Initial state cannot be empty");
    if (initial_state_.size() != model_template_->getStateSize())
        throw InvalidParameterException("PostCalibrationAnalyser", "Initial state size does not match model state size");
}

VectorXd PostCalibrationAnalyser::meanOfSamples(const std::vector<VectorXd>& samples, int burn_in, int thinning, std::ptrdiff_t n_params) {
    VectorXd mean = VectorXd::Zero(n_params);
    int count = 0;
    for (size_t i = static_cast<size_t>(std::max(burn_in, 0)); i < samples.size(); i += static_cast<size_t>(std::max(thinning, 1))) {
        mean += samples[i];
        ++count;
    }
    if (count > 0) mean /= static_cast<double>(count);
    return mean;
}

std::vector<PostCalibrationAnalyser::NamedParameters> PostCalibrationAnalyser::defineNpiScenarios(const SEPAIHRDParameters& baseline) const {
    size_t first_modifiable = 1;                // kappa_1 is the fixed baseline period
    if (auto pw = std::dynamic_pointer_cast<PiecewiseConstantNpiStrategy>(model_template_->getNpiStrategy()))
        if (!pw->isBaselineFixed()) first_modifiable = 0;
    std::vector<NamedParameters> scenarios;
    if (baseline.kappa_values.size() > first_modifiable) {
        const std::pair<const char*, double> variants[] = {{"stricter_lockdown", 0.9}, {"weaker_lockdown", 1.1}};
        for (const auto& [name, factor] : variants) {
            SEPAIHRDParameters p = baseline;
            p.kappa_values[first_modifiable] *= factor;
            scenarios.emplace_back(name, std::move(p));
        }
    }
    return scenarios;
}

std::vector<PostCalibrationAnalyser::NamedMetrics> PostCalibrationAnalyser::performScenarioAnalysis(
    const SEPAIHRDParameters& baseline_params, const std::vector<NamedParameters>& scenarios, std::vector<SimulationResult>* trajectories) {
    std::vector<std::string> names = {"baseline"};
    std::vector<SEPAIHRDParameters> params = {baseline_params};
    for (const auto& [name, p] : scenarios) { names.push_back(name); params.push_back(p); }
    std::vector<SimulationResult> runs = runner_.runSimulations(params, initial_state_, time_points_);
    std::vector<NamedMetrics> out;
    for (size_t r = 0; r < runs.size(); ++r) {
        // like analyzeSingleRunLightweight (.cpp:142-165): the metrics model carries the run's parameters but a CLONE OF THE
        // TEMPLATE'S NPI schedule, i.e. R0 / Rt / the infection sum see the baseline kappa values even for a scenario run
        auto metrics_model = std::make_shared<AgeSEPAIHRDModel>(params[r], model_template_->getNpiStrategy()->clone());
        out.emplace_back(names[r], metrics_.calculateEssentialMetrics(runs[r], metrics_model, params[r], initial_state_, time_points_));
    }
    if (trajectories) *trajectories = std::move(runs);
    return out;
}

std::vector<PostCalibrationAnalyser::NamedMetrics> PostCalibrationAnalyser::scenarioAnalysisFromSamples(
    const std::vector<VectorXd>& samples, SEPAIHRDParameterManager& pm, int burn_in, int thinning, std::vector<SimulationResult>* trajectories) {
    if (samples.empty()) return {};
    const VectorXd mean = meanOfSamples(samples, burn_in, thinning, static_cast<std::ptrdiff_t>(pm.getParameterCount()));
    pm.updateModelParameters(mean, model_template_);
    const SEPAIHRDParameters baseline = model_template_->getModelParameters();
    return performScenarioAnalysis(baseline, defineNpiScenarios(baseline), trajectories);
}

PostCalibrationAnalyser::McmcAnalysis PostCalibrationAnalyser::analyzeMCMCRuns(const std::vector<VectorXd>& samples, SEPAIHRDParameterManager& pm,
                                                                               int burn_in, int thinning) {
    McmcAnalysis res;
    std::vector<SEPAIHRDParameters> params;
    std::vector<std::shared_ptr<AgeSEPAIHRDModel>> models;
    for (size_t i = static_cast<size_t>(std::max(burn_in, 0)); i < samples.size(); i += static_cast<size_t>(std::max(thinning, 1))) {
        pm.updateModelParameters(samples[i], model_template_);
        params.push_back(model_template_->getModelParameters());
        models.push_back(model_template_->clone());          // here the NPI values ARE the sample's (.cpp:198-222)
    }
    if (params.empty()) return res;
    const std::vector<SimulationResult> runs = runner_.runSimulations(params, initial_state_, time_points_);
    const size_t R = runs.size();
    res.metrics.resize(R);
    std::vector<std::vector<double>> rt(R), sero(R);
#pragma omp parallel for schedule(dynamic)
    for (size_t r = 0; r < R; ++r) {
        res.metrics[r] = metrics_.calculateEssentialMetrics(runs[r], models[r], params[r], initial_state_, time_points_);
        rt[r] = metrics_.calculateRtTrajectory(runs[r], models[r], time_points_);
        sero[r] = metrics_.calculateSeroprevalenceTrajectory(runs[r], params[r], time_points_);
    }
    res.rt = aggregateTrajectories(rt, time_points_.size());
    res.seroprevalence = aggregateTrajectories(sero, time_points_.size());
    return res;
}

AggregatedTrajectory PostCalibrationAnalyser::aggregateTrajectories(const std::vector<std::vector<double>>& trajectories, size_t num_timesteps) {
    AggregatedTrajectory agg;
    std::vector<double> column;
    for (size_t t = 0; t < num_timesteps; ++t) {
        column.clear();
        for (const auto& traj : trajectories)
            if (t < traj.size()) column.push_back(traj[t]);
        if (column.empty()) break;
        std::sort(column.begin(), column.end());
        agg.median.push_back(sorted_quantile(column, 0.5));
        agg.q025.push_back(sorted_quantile(column, 0.025));
        agg.q975.push_back(sorted_quantile(column, 0.975));
        agg.q05.push_back(sorted_quantile(column, 0.05));
        agg.q95.push_back(sorted_quantile(column, 0.95));
    }
    return agg;
}

void PostCalibrationAnalyser::writeScenarioComparison(const std::string& filepath, const std::vector<NamedMetrics>& scenarios) {
    std::ofstream file(filepath);
    if (!file.is_open()) throw ModelException("AnalysisWriter", "Failed to open: " + filepath);
    file << "scenario,R0,overall_IFR,overall_attack_rate,peak_hospital,peak_ICU,time_to_peak_hospital,time_to_peak_ICU,total_deaths,seroprevalence_day64";
    if (!scenarios.empty())
        for (const auto& kv : scenarios[0].second.kappa_values) file << ',' << kv.first;
    file << '\n';
    for (const auto& [name, m] : scenarios) {
        file << name << ',' << m.R0 << ',' << m.overall_IFR << ',' << m.overall_attack_rate << ',' << m.peak_hospital_occupancy << ','
             << m.peak_ICU_occupancy << ',' << m.time_to_peak_hospital << ',' << m.time_to_peak_ICU << ',' << m.total_cumulative_deaths << ','
             << m.seroprevalence_at_target_day;
        for (const auto& kv : m.kappa_values) file << ',' << kv.second;
        file << '\n';
    }
}

// ---- AnalysisWriter -----------------------------------------------------------------------------------------------------
void AnalysisWriter::savePosteriorPredictiveData(const std::string& output_dir, const PosteriorPredictiveData& ppd) const {
    std::ptrdiff_t n_ages = 0;                                        // .cpp:288-299: from the data, else the default 4
    if (ppd.daily_hospitalizations.median.cols() > 0) n_ages = ppd.daily_hospitalizations.median.cols();
    else if (ppd.daily_deaths.median.cols() > 0) n_ages = ppd.daily_deaths.median.cols();
    else n_ages = 4;
    auto save_matrix = [&](const MatrixXd& m, const std::string& base, const char* suffix) {
        std::ofstream file(output_dir + "/" + base + "_" + suffix + ".csv");
        if (!file.is_open()) return;                                  // the reference logs and goes on
        const std::ptrdiff_t cols = std::min(n_ages, m.cols());
        file << "time";
        for (std::ptrdiff_t a = 0; a < cols; ++a) file << ",age_" << a;
        file << "\n";
        for (size_t t = 0; t < ppd.time_points.size(); ++t) {
            file << ppd.time_points[t];
            for (std::ptrdiff_t a = 0; a < cols; ++a) file << "," << std::fixed << std::setprecision(6) << m(static_cast<std::ptrdiff_t>(t), a);
            file << "\n";
        }
    };
    auto save_series = [&](const PosteriorPredictiveData::IncidenceData& d, const std::string& base) {
        save_matrix(d.median, base, "median");
        save_matrix(d.lower_90, base, "lower90");
        save_matrix(d.upper_90, base, "upper90");
        save_matrix(d.lower_95, base, "lower95");
        save_matrix(d.upper_95, base, "upper95");
        save_matrix(d.observed, base, "observed");
    };
    save_series(ppd.daily_hospitalizations, "daily_hospitalizations");
    save_series(ppd.daily_icu_admissions, "daily_icu_admissions");
    save_series(ppd.daily_deaths, "daily_deaths");
    save_series(ppd.cumulative_hospitalizations, "cumulative_hospitalizations");
    save_series(ppd.cumulative_icu_admissions, "cumulative_icu_admissions");
    save_series(ppd.cumulative_deaths, "cumulative_deaths");
}

void AnalysisWriter::saveParameterPosteriors(const std::string& output_dir, const std::vector<VectorXd>& samples,
                                             const std::vector<std::string>& names, int burn_in, int thinning) const {
    const size_t first = static_cast<size_t>(std::max(burn_in, 0)), step = static_cast<size_t>(std::max(thinning, 1));
    {
        std::ofstream sfile(output_dir + "/posterior_samples.csv");
        if (!sfile.is_open()) return;
        sfile << "sample_index";
        for (const auto& nm : names) sfile << "," << nm;
        sfile << "\n";
        int saved = 0;
        for (size_t i = first; i < samples.size(); i += step) {
            sfile << saved++;
            for (std::ptrdiff_t j = 0; j < samples[i].size(); ++j) sfile << "," << std::scientific << std::setprecision(8) << samples[i](j);
            sfile << "\n";
        }
    }
    std::ofstream sum(output_dir + "/posterior_summary.csv");
    if (!sum.is_open()) return;
    sum << "parameter,mean,median,std_dev,lower_95_ci,upper_95_ci\n";
    sum << std::fixed << std::setprecision(8);
    for (size_t p = 0; p < names.size(); ++p) {
        std::vector<double> values;
        for (size_t i = first; i < samples.size(); i += step)
            if (static_cast<std::ptrdiff_t>(p) < samples[i].size()) values.push_back(samples[i](static_cast<std::ptrdiff_t>(p)));
        if (values.empty()) continue;
        std::sort(values.begin(), values.end());
        double total = 0.0;
        for (double v : values) total += v;                           // std::accumulate over the SORTED values (.cpp:258)
        const double mean = total / static_cast<double>(values.size());
        const double median = values[values.size() / 2];              // upper median, no interpolation (.cpp:259)
        const double q025 = values[static_cast<size_t>(0.025 * static_cast<double>(values.size()))];
        const double q975 = values[static_cast<size_t>(0.975 * static_cast<double>(values.size()))];
        double ss = 0.0;
        for (double v : values) ss += (v - mean) * (v - mean);
        const double sd = std::sqrt(ss / static_cast<double>(values.size()));   // population form
        sum << names[p] << "," << mean << "," << median << "," << sd << "," << q025 << "," << q975 << "\n";
    }
}

}  // namespace epidemic
