// host_capi.cpp -- extern "C" wrappers over the C++ host layer.  See host_capi.h.
#include "host_capi.h"
#include "../csrc/det_math.h"

#include <chrono>
#include <cmath>
#include <omp.h>
#include <cstring>
#include <string>

#include "analysis.hpp"
#include "optimizers.hpp"

using namespace epidemic;

namespace {

thread_local std::string g_err;

template <class F>
int32_t guarded(F&& f) {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
    } catch (...) {
        g_err = "unknown C++ exception";
    }
    return 1;
}

std::map<std::string, double> settings_map(int32_t n, const char* const* keys, const double* values) {
    std::map<std::string, double> m;
    for (int32_t i = 0; i < n; ++i) m[keys[i]] = values[i];
    return m;
}

// IParameterManager over plain arrays: what the samplers need (sigmas, bounds, clamp / reflect) without a model.
class ArrayParameterManager : public IParameterManager {
public:
    ArrayParameterManager(int n, const double* sigmas, const double* lo, const double* hi, int mode)
        : sig_(sigmas, sigmas + n), lo_(lo, lo + n), hi_(hi, hi + n), mode_(mode) {
        for (int i = 0; i < n; ++i) names_.push_back("p" + std::to_string(i));
    }
    VectorXd getCurrentParameters() const override { return VectorXd::Zero(static_cast<std::ptrdiff_t>(sig_.size())); }
    void updateModelParameters(const VectorXd&) override {}
    const std::vector<std::string>& getParameterNames() const override { return names_; }
    size_t getParameterCount() const override { return sig_.size(); }
    double getSigmaForParamIndex(int i) const override { return sig_.at(static_cast<size_t>(i)); }
    VectorXd applyConstraints(const VectorXd& p) const override {          // SEPAIHRDParameterManager.cpp:315-347
        if (static_cast<size_t>(p.size()) != sig_.size()) throw InvalidParameterException("applyConstraints", "Parameter vector size mismatch.");
        VectorXd out = p;
        for (size_t i = 0; i < sig_.size(); ++i) {
            const auto k = static_cast<std::ptrdiff_t>(i);
            if (!std::isnan(lo_[i])) {
                double lo = lo_[i], hi = hi_[i];
                if (lo > hi) std::swap(lo, hi);
                out(k) = (mode_ == 0) ? std::min(std::max(p(k), lo), hi) : SEPAIHRDParameterManager::reflectBound(p(k), lo, hi);
            } else {
                out(k) = (mode_ == 0) ? std::max(0.0, p(k)) : std::abs(p(k));
            }
        }
        return out;
    }
    int getIndexForParam(const std::string& name) const override {
        for (size_t i = 0; i < names_.size(); ++i) if (names_[i] == name) return static_cast<int>(i);
        return -1;
    }
    double getLowerBoundForParamIndex(int i) const override { return lo_.at(static_cast<size_t>(i)); }
    double getUpperBoundForParamIndex(int i) const override { return hi_.at(static_cast<size_t>(i)); }
    void setMode(int m) { mode_ = m; }

private:
    std::vector<double> sig_, lo_, hi_;
    int mode_;
    std::vector<std::string> names_;
};

// IObjectiveFunction over a C batch callback.
class CallbackObjective : public virtual IObjectiveFunction {
public:
    CallbackObjective(sepaihrd_host_batch_fn fn, void* user, const std::vector<std::string>& names) : fn_(fn), user_(user), names_(names) {}
    double calculate(const VectorXd& p) const override {
        double out = 0.0;
        calculateBatch(p.data(), 1, p.size(), &out);
        return out;
    }
    void calculateBatch(const double* params, int64_t B, int64_t ld, double* out) const override {
        n_evals += B;
        if (fn_(user_, params, B, ld, out) != 0) throw SimulationException("CallbackObjective", "the batch callback reported a failure");
    }
    const std::vector<std::string>& getParameterNames() const override { return names_; }
    mutable int64_t n_evals = 0;

private:
    sepaihrd_host_batch_fn fn_;
    void* user_;
    std::vector<std::string> names_;
};

// the callback objective with the forward-difference gradient of SEPAIHRDGradientObjectiveFunction: centre value, then the P
// perturbed vectors as ONE batch
class CallbackGradientObjective : public CallbackObjective, public IGradientObjectiveFunction {
public:
    using CallbackObjective::CallbackObjective;
    double epsilon_ = 1e-4;
    double evaluate_with_gradient(const VectorXd& params, VectorXd& grad) const override {
        const double f_center = calculate(params);
        grad.resize(params.size());
        if (!std::isfinite(f_center)) { grad.setZero(); return f_center; }
        std::vector<double> rows, steps, f_plus(static_cast<size_t>(params.size()));
        ForwardDifferences::perturb(params, epsilon_, rows, steps);
        calculateBatch(rows.data(), params.size(), params.size(), f_plus.data());
        ForwardDifferences::gradient(f_center, f_plus.data(), nullptr, steps, grad);
        return f_center;
    }
};

}  // namespace

struct sepaihrd_host_pm { ArrayParameterManager pm; };
struct sepaihrd_host_mh { MetropolisHastingsSampler s; sepaihrd_host_pm* pm; };
struct sepaihrd_host_pso { ParticleSwarmOptimization s; sepaihrd_host_pm* pm; };

struct sepaihrd_host_model {
    std::shared_ptr<AgeSEPAIHRDModel> model;
    std::unique_ptr<CalibrationData> data;
    std::vector<double> times;
    std::unique_ptr<SEPAIHRDModelCalibration> calibration;
    std::unique_ptr<SEPAIHRDParameterManager> pm;
    std::shared_ptr<ISimulationCache> cache;
    std::unique_ptr<SEPAIHRDObjectiveFunction> objective;
    std::unique_ptr<AgeSEPAIHRDSimulator> simulator;
    double abs_tol = 1e-6, rel_tol = 1e-6;
    std::vector<std::string> pnames;
    std::map<std::string, double> sig;
    std::map<std::string, std::pair<double, double>> bounds;
    // the calibration and the objective hold the cache: (re)built whenever the cache changes
    void build(std::shared_ptr<ISimulationCache> c) {
        objective.reset(); calibration.reset();
        cache = std::move(c);
        calibration = std::make_unique<SEPAIHRDModelCalibration>(model, *data, times, pnames, sig, bounds, std::make_shared<Dopri5SolverStrategy>(), cache);
        objective = std::make_unique<SEPAIHRDObjectiveFunction>(model, *pm, *cache, *data, times, data->getInitialSEPAIHRDState(),
                                                                std::make_shared<Dopri5SolverStrategy>(), abs_tol, rel_tol);
    }
};
struct sepaihrd_host_cache { SimulationCache c; explicit sepaihrd_host_cache(size_t n) : c(n) {} };

// SEPAIHRDParameters -> PiecewiseConstantNpiStrategy -> AgeSEPAIHRDModel from the flat problem description (host only: no
// device is touched).  createNpiStrategy (src/model/main.cpp:81-130): element 0 of the kappa schedule is the fixed baseline.
static std::shared_ptr<AgeSEPAIHRDModel> model_from_problem(const sepaihrd_problem* pb, const std::map<std::string, std::pair<double, double>>& bounds,
                                                     SEPAIHRDParameters* params_out = nullptr) {
    const int n = pb->n_ages, nb = pb->n_beta, nk = pb->n_kappa;
    const double* s = pb->base_slots;
    const int scal0 = nb + nk, age0 = scal0 + 7, mult0 = age0 + 8 * n;
    SEPAIHRDParameters p;
    p.N = VectorXd::FromPointer(pb->population, n);
    p.M_baseline = MatrixXd(n, n);
    std::copy(pb->contact_matrix, pb->contact_matrix + n * n, p.M_baseline.data());      // column-major both sides
    p.beta_end_times.assign(pb->beta_end_times, pb->beta_end_times + nb);
    p.beta_values.assign(s, s + nb);
    p.kappa_end_times.assign(pb->kappa_end_times, pb->kappa_end_times + nk);
    p.kappa_values.assign(s + nb, s + nb + nk);
    p.theta = s[scal0]; p.sigma = s[scal0 + 1]; p.gamma_p = s[scal0 + 2]; p.gamma_A = s[scal0 + 3];
    p.gamma_I = s[scal0 + 4]; p.gamma_H = s[scal0 + 5]; p.gamma_ICU = s[scal0 + 6];
    VectorXd* blocks[8] = {&p.a, &p.h_infec, &p.p, &p.h, &p.icu, &p.d_H, &p.d_ICU, &p.d_community};
    for (int b = 0; b < 8; ++b) *blocks[b] = VectorXd::FromPointer(s + age0 + b * n, n);
    p.E0_multiplier = s[mult0]; p.P0_multiplier = s[mult0 + 1]; p.A0_multiplier = s[mult0 + 2]; p.I0_multiplier = s[mult0 + 3];
    p.H0_multiplier = s[mult0 + 4]; p.ICU0_multiplier = s[mult0 + 5]; p.R0_multiplier = s[mult0 + 6]; p.D0_multiplier = s[mult0 + 7];
    p.seed_exposed = s[mult0 + 8]; p.runup_days = s[mult0 + 9]; p.beta = s[mult0 + 10];
    std::vector<std::string> knames;
    std::map<std::string, std::pair<double, double>> kbounds;
    for (int k = 1; k < nk; ++k) {
        knames.push_back("kappa_" + std::to_string(k + 1));
        auto it = bounds.find(knames.back());
        if (it != bounds.end()) kbounds[knames.back()] = it->second;
    }
    auto npi = std::make_shared<PiecewiseConstantNpiStrategy>(
        std::vector<double>(p.kappa_end_times.begin() + 1, p.kappa_end_times.end()),
        std::vector<double>(p.kappa_values.begin() + 1, p.kappa_values.end()), kbounds, p.kappa_values.at(0), p.kappa_end_times.at(0), true, knames);
    if (params_out) *params_out = p;
    return std::make_shared<AgeSEPAIHRDModel>(p, npi);
}

extern "C" {

const char* sepaihrd_host_last_error(void) { return g_err.c_str(); }

int32_t sepaihrd_host_set_trace_directory(const char* dir) {
    return guarded([&] { MetropolisHastingsSampler::setDefaultOutputDirectory(dir ? dir : ""); });
}
int32_t sepaihrd_host_set_threads(int32_t n) {
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
}

// ---- parameter manager ---------------------------------------------------------------------------------------
int32_t sepaihrd_host_pm_create(int32_t n, const double* sigmas, const double* lo, const double* hi, int32_t mode, sepaihrd_host_pm** out) {
    return guarded([&] {
        if (n <= 0 || !sigmas || !lo || !hi || !out) throw InvalidParameterException("sepaihrd_host_pm_create", "bad argument");
        *out = new sepaihrd_host_pm{ArrayParameterManager(n, sigmas, lo, hi, mode)};
    });
}
int32_t sepaihrd_host_pm_set_mode(sepaihrd_host_pm* pm, int32_t mode) { return guarded([&] { pm->pm.setMode(mode); }); }
int32_t sepaihrd_host_pm_apply_constraints(const sepaihrd_host_pm* pm, const double* in, double* out) {
    return guarded([&] {
        const auto n = static_cast<std::ptrdiff_t>(pm->pm.getParameterCount());
        const VectorXd r = pm->pm.applyConstraints(VectorXd::FromPointer(in, n));
        std::copy(r.data(), r.data() + n, out);
    });
}
void sepaihrd_host_pm_destroy(sepaihrd_host_pm* pm) { delete pm; }

// ---- Metropolis-Hastings ------------------------------------------------------------------------------------
int32_t sepaihrd_host_mh_create(sepaihrd_host_pm* pm, int32_t n, const char* const* keys, const double* values, sepaihrd_host_mh** out) {
    return guarded([&] {
        if (!pm || !out) throw InvalidParameterException("sepaihrd_host_mh_create", "bad argument");
        auto* h = new sepaihrd_host_mh{MetropolisHastingsSampler(), pm};
        h->s.configure(settings_map(n, keys, values));
        *out = h;
    });
}
int32_t sepaihrd_host_mh_set_initial_covariance(sepaihrd_host_mh* mh, const double* cov, int32_t n) {
    return guarded([&] {
        MatrixXd m(n, n);
        std::copy(cov, cov + static_cast<size_t>(n) * n, m.data());
        mh->s.setInitialCovariance(m);
    });
}
int32_t sepaihrd_host_mh_begin(sepaihrd_host_mh* mh, const double* initial, const double* lp) {
    return guarded([&] { mh->s.begin(VectorXd::FromPointer(initial, static_cast<std::ptrdiff_t>(mh->pm->pm.getParameterCount())), lp, mh->pm->pm); });
}
int32_t sepaihrd_host_mh_done(const sepaihrd_host_mh* mh) { return mh->s.done() ? 1 : 0; }
int32_t sepaihrd_host_mh_iteration(const sepaihrd_host_mh* mh) { return mh->s.iteration(); }
int32_t sepaihrd_host_mh_propose(sepaihrd_host_mh* mh, double* out) { return guarded([&] { mh->s.propose(mh->pm->pm, out); }); }
int32_t sepaihrd_host_mh_accept(sepaihrd_host_mh* mh, const double* lp, uint8_t* acc) { return guarded([&] { mh->s.accept(lp, acc); }); }
int32_t sepaihrd_host_mh_state(const sepaihrd_host_mh* mh, double* x, double* lp, double* scale, int64_t* accepted) {
    return guarded([&] {
        const int n = mh->s.numChains(), P = mh->s.numParams();
        if (x) std::copy(mh->s.currentPositions(), mh->s.currentPositions() + static_cast<size_t>(n) * P, x);
        if (lp) std::copy(mh->s.currentLogPost(), mh->s.currentLogPost() + n, lp);
        for (int c = 0; c < n; ++c) {
            if (scale) scale[c] = mh->s.globalScale(c);
            if (accepted) accepted[c] = mh->s.acceptedCount(c);
        }
    });
}
int32_t sepaihrd_host_mh_best(const sepaihrd_host_mh* mh, double* x, double* value) {
    return guarded([&] {
        const OptimizationResult r = mh->s.result();
        if (x) std::copy(r.bestParameters.data(), r.bestParameters.data() + r.bestParameters.size(), x);
        if (value) *value = r.bestObjectiveValue;
    });
}
double sepaihrd_host_det_log(double x) { return detm::log(x); }
double sepaihrd_host_det_exp(double x) { return detm::exp(x); }
int32_t sepaihrd_host_mh_shared_cholesky(const sepaihrd_host_mh* mh, double* out) {
    return guarded([&] {
        const MatrixXd& L = mh->s.sharedCholesky();
        std::copy(L.data(), L.data() + L.rows() * L.cols(), out);
    });
}
void sepaihrd_host_mh_destroy(sepaihrd_host_mh* mh) { delete mh; }

// ---- particle swarm -----------------------------------------------------------------------------------------
int32_t sepaihrd_host_pso_create(sepaihrd_host_pm* pm, int32_t n, const char* const* keys, const double* values, sepaihrd_host_pso** out) {
    return guarded([&] {
        if (!pm || !out) throw InvalidParameterException("sepaihrd_host_pso_create", "bad argument");
        auto* h = new sepaihrd_host_pso{ParticleSwarmOptimization(), pm};
        h->s.configure(settings_map(n, keys, values));
        *out = h;
    });
}
int32_t sepaihrd_host_pso_begin(sepaihrd_host_pso* pso, const double* init) {
    return guarded([&] {
        if (init) {
            const VectorXd v = VectorXd::FromPointer(init, static_cast<std::ptrdiff_t>(pso->pm->pm.getParameterCount()));
            pso->s.begin(&v, pso->pm->pm);
        } else {
            pso->s.begin(nullptr, pso->pm->pm);
        }
    });
}
int32_t sepaihrd_host_pso_local_count(const sepaihrd_host_pso* pso) { return pso->s.localCount(); }
int32_t sepaihrd_host_pso_positions(const sepaihrd_host_pso* pso, double* out) {
    return guarded([&] { std::copy(pso->s.positions(), pso->s.positions() + static_cast<size_t>(pso->s.localCount()) * pso->s.numParams(), out); });
}
int32_t sepaihrd_host_pso_tell(sepaihrd_host_pso* pso, const double* fitness, double* best_value, int32_t* best_local, double* best_pos) {
    return guarded([&] {
        const auto b = pso->s.tell(fitness);
        if (best_value) *best_value = b.first;
        if (best_local) *best_local = b.second;
        if (best_pos && b.second >= 0) std::copy(pso->s.personalBest(b.second), pso->s.personalBest(b.second) + pso->s.numParams(), best_pos);
    });
}
int32_t sepaihrd_host_pso_set_global_best(sepaihrd_host_pso* pso, double value, const double* pos) { return guarded([&] { pso->s.setGlobalBest(value, pos); }); }
int32_t sepaihrd_host_pso_global_best(const sepaihrd_host_pso* pso, double* value, double* pos) {
    return guarded([&] {
        if (value) *value = pso->s.globalBestValue();
        if (pos) std::copy(pso->s.globalBestPosition().begin(), pso->s.globalBestPosition().end(), pos);
    });
}
int32_t sepaihrd_host_pso_step(sepaihrd_host_pso* pso, int32_t iter) { return guarded([&] { pso->s.step(iter); }); }
int32_t sepaihrd_host_pso_begin_device(sepaihrd_host_pso* pso, const double* init, sepaihrd_ctx* ctx) {
    return guarded([&] {
        if (init) {
            const VectorXd v = VectorXd::FromPointer(init, static_cast<std::ptrdiff_t>(pso->pm->pm.getParameterCount()));
            pso->s.beginDevice(&v, pso->pm->pm, ctx);
        } else {
            pso->s.beginDevice(nullptr, pso->pm->pm, ctx);
        }
    });
}
int32_t sepaihrd_host_pso_evaluate_device(sepaihrd_host_pso* pso, double* best_value, int32_t* best_local, double* best_pos) {
    return guarded([&] {
        const auto b = pso->s.evaluateDevice(best_pos);
        if (best_value) *best_value = b.first;
        if (best_local) *best_local = b.second;
    });
}
int32_t sepaihrd_host_pso_step_device(sepaihrd_host_pso* pso, int32_t iter) { return guarded([&] { pso->s.stepDevice(iter); }); }
int32_t sepaihrd_host_pso_fetch(sepaihrd_host_pso* pso) { return guarded([&] { pso->s.fetchPersonalBests(); }); }
int32_t sepaihrd_host_pso_run(sepaihrd_host_pso* pso, sepaihrd_host_batch_fn fn, void* user, const double* init, double* out_best,
                              double* out_value, double* out_stats) {
    return guarded([&] {
        if (!pso || !fn) throw InvalidParameterException("sepaihrd_host_pso_run", "bad argument");
        CallbackObjective f(fn, user, pso->pm->pm.getParameterNames());
        const auto P = static_cast<std::ptrdiff_t>(pso->pm->pm.getParameterCount());
        const VectorXd x0 = init ? VectorXd::FromPointer(init, P) : VectorXd();
        const OptimizationResult r = pso->s.optimize(x0, f, pso->pm->pm);
        if (out_best) std::copy(r.bestParameters.data(), r.bestParameters.data() + P, out_best);
        if (out_value) *out_value = r.bestObjectiveValue;
        if (out_stats) {
            out_stats[0] = static_cast<double>(pso->s.evaluations()); out_stats[1] = pso->s.restartCount();
            out_stats[2] = pso->s.elitistTrials(); out_stats[3] = pso->s.swarmDiversity();
        }
    });
}
int32_t sepaihrd_host_pso_values(const sepaihrd_host_pso* pso, double* out_pbest, double* out_fitness) {
    return guarded([&] {
        if (out_pbest) std::copy(pso->s.personalBestValues().begin(), pso->s.personalBestValues().end(), out_pbest);
        if (out_fitness) std::copy(pso->s.currentFitness().begin(), pso->s.currentFitness().end(), out_fitness);
    });
}
int32_t sepaihrd_host_pso_neighbors(sepaihrd_host_pso* pso, int32_t particle, int32_t* out, int32_t cap) {
    int32_t count = -1;
    const int32_t rc = guarded([&] {
        if (!pso || particle < 0 || particle >= pso->s.swarmSize()) throw InvalidParameterException("sepaihrd_host_pso_neighbors", "particle outside the swarm");
        const std::vector<int> nb = pso->s.getNeighbors(particle);
        count = static_cast<int32_t>(nb.size());
        for (int32_t i = 0; i < count && i < cap; ++i) out[i] = nb[static_cast<size_t>(i)];
    });
    return rc == 0 ? count : -1;
}
void sepaihrd_host_pso_destroy(sepaihrd_host_pso* pso) { delete pso; }

// ---- whole runs against a callback --------------------------------------------------------------------------
int32_t sepaihrd_host_optimize(const char* algorithm, sepaihrd_host_pm* pm, int32_t n, const char* const* keys, const double* values,
                               sepaihrd_host_batch_fn fn, void* user, const double* initial, double* out_best, double* out_value,
                               int64_t* out_evals) {
    return guarded([&] {
        const std::string which = algorithm ? algorithm : "";
        std::unique_ptr<IOptimizationAlgorithm> algo;
        if (which == "mh") algo = std::make_unique<MetropolisHastingsSampler>();
        else if (which == "pso") algo = std::make_unique<ParticleSwarmOptimization>();
        else if (which == "hill") algo = std::make_unique<HillClimbingOptimizer>();
        else if (which == "nuts") algo = std::make_unique<NUTSSampler>();
        else throw InvalidParameterException("sepaihrd_host_optimize", "algorithm must be mh, pso, hill or nuts");
        algo->configure(settings_map(n, keys, values));
        CallbackGradientObjective f(fn, user, pm->pm.getParameterNames());
        const auto P = static_cast<std::ptrdiff_t>(pm->pm.getParameterCount());
        const OptimizationResult r = algo->optimize(VectorXd::FromPointer(initial, P), f, pm->pm);
        if (out_best) std::copy(r.bestParameters.data(), r.bestParameters.data() + P, out_best);
        if (out_value) *out_value = r.bestObjectiveValue;
        if (out_evals) *out_evals = f.n_evals;
    });
}

int32_t sepaihrd_host_calibrate(const char* phase1, sepaihrd_host_pm* pm, int32_t n1, const char* const* k1, const double* v1, int32_t n2,
                                const char* const* k2, const double* v2, sepaihrd_host_batch_fn fn, void* user, const double* initial,
                                double* out_best, double* out_value, int64_t* out_samples, double* out_phase1) {
    return guarded([&] {
        // ModelCalibrator owns its parameter manager and objective: hand it copies that forward to the caller's objects
        struct ForwardPM : IParameterManager {
            ArrayParameterManager& p; VectorXd current;
            ForwardPM(ArrayParameterManager& p_, const VectorXd& c) : p(p_), current(c) {}
            VectorXd getCurrentParameters() const override { return current; }
            void updateModelParameters(const VectorXd& v) override { current = v; }
            const std::vector<std::string>& getParameterNames() const override { return p.getParameterNames(); }
            size_t getParameterCount() const override { return p.getParameterCount(); }
            double getSigmaForParamIndex(int i) const override { return p.getSigmaForParamIndex(i); }
            VectorXd applyConstraints(const VectorXd& v) const override { return p.applyConstraints(v); }
            int getIndexForParam(const std::string& n) const override { return p.getIndexForParam(n); }
            double getLowerBoundForParamIndex(int i) const override { return p.getLowerBoundForParamIndex(i); }
            double getUpperBoundForParamIndex(int i) const override { return p.getUpperBoundForParamIndex(i); }
        };
        const auto P = static_cast<std::ptrdiff_t>(pm->pm.getParameterCount());
        std::map<std::string, std::unique_ptr<IOptimizationAlgorithm>> algos;
        const std::string which = phase1 ? phase1 : "pso";
        if (which == "hill") algos[ModelCalibrator::PHASE1_NAME] = std::make_unique<HillClimbingOptimizer>();
        else if (which == "pso") algos[ModelCalibrator::PHASE1_NAME] = std::make_unique<ParticleSwarmOptimization>();
        else throw InvalidParameterException("sepaihrd_host_calibrate", "phase1 must be pso or hill");
        // the sampler of phase 2 flips the SEPAIHRD manager to reflection itself; an array manager is flipped here
        struct ReflectingMH : MetropolisHastingsSampler {
            ArrayParameterManager& p;
            explicit ReflectingMH(ArrayParameterManager& p_) : p(p_) {}
            OptimizationResult optimize(const VectorXd& x0, IObjectiveFunction& f, IParameterManager& m) override {
                p.setMode(1);
                return MetropolisHastingsSampler::optimize(x0, f, m);
            }
        };
        algos[ModelCalibrator::PHASE2_NAME] = std::make_unique<ReflectingMH>(pm->pm);
        pm->pm.setMode(0);
        auto objective = std::make_unique<CallbackObjective>(fn, user, pm->pm.getParameterNames());
        ModelCalibrator c(std::make_unique<ForwardPM>(pm->pm, VectorXd::FromPointer(initial, P)), std::move(objective), std::move(algos));
        c.calibrate(settings_map(n1, k1, v1), settings_map(n2, k2, v2));
        const VectorXd& b = c.getBestParameterVector();
        if (out_best) std::copy(b.data(), b.data() + P, out_best);
        if (out_value) *out_value = c.getBestObjectiveValue();
        if (out_samples) *out_samples = static_cast<int64_t>(c.getMCMCObjectiveValues().size());
        if (out_phase1) *out_phase1 = c.getPhase1Result().bestObjectiveValue;
    });
}

// ---- reference-shaped object graph --------------------------------------------------------------------------
int32_t sepaihrd_host_model_create(const sepaihrd_problem* pb, const char* const* names, const double* sigmas, sepaihrd_host_model** out) {
    return guarded([&] {
        if (!pb || !names || !sigmas || !out) throw InvalidParameterException("sepaihrd_host_model_create", "bad argument");
        const int n = pb->n_ages;
        std::map<std::string, double> sig;
        std::map<std::string, std::pair<double, double>> bounds;
        std::vector<std::string> pnames;
        for (int i = 0; i < pb->n_params; ++i) {
            pnames.emplace_back(names[i]);
            sig[names[i]] = sigmas[i];
            bounds[names[i]] = {pb->lower_bound[i], pb->upper_bound[i]};
        }
        SEPAIHRDParameters p;
        auto h = std::make_unique<sepaihrd_host_model>();
        h->model = model_from_problem(pb, bounds, &p);
        auto obs = [&](const double* src) {
            MatrixXd m(pb->n_obs, n);
            for (int r = 0; r < pb->n_obs; ++r) for (int a = 0; a < n; ++a) m(r, a) = src[r * n + a];
            return m;
        };
        h->data = std::make_unique<CalibrationData>(obs(pb->obs_hosp), obs(pb->obs_icu), obs(pb->obs_deaths), p.N,
                                                    VectorXd::FromPointer(pb->data_initial_state, SEPAIHRD_NUM_COMPARTMENTS * n));
        h->times.assign(pb->times, pb->times + pb->n_times);
        h->abs_tol = pb->abs_tol; h->rel_tol = pb->rel_tol;
        h->pm = std::make_unique<SEPAIHRDParameterManager>(h->model, pnames, sig, bounds);
        h->pm->setConstraintMode(pb->constraint_mode == 1 ? ConstraintMode::MCMC_REFLECT : ConstraintMode::OPTIMIZATION_CLAMP);
        h->pnames = pnames; h->sig = sig; h->bounds = bounds;
        h->build(std::make_shared<NullSimulationCache>());         // parity runs: no cache (quirk Q6); sepaihrd_host_model_set_cache switches
        *out = h.release();
    });
}

int32_t sepaihrd_host_model_calculate(sepaihrd_host_model* m, const double* params, double* out) {
    return guarded([&] { *out = m->objective->calculate(VectorXd::FromPointer(params, static_cast<std::ptrdiff_t>(m->pm->getParameterCount()))); });
}
int32_t sepaihrd_host_model_calculate_batch(sepaihrd_host_model* m, const double* params, int64_t B, int64_t ld, double* out) {
    return guarded([&] { m->objective->calculateBatch(params, B, ld, out); });
}
int32_t sepaihrd_host_model_set_constraint_mode(sepaihrd_host_model* m, int32_t mode) {
    return guarded([&] { m->pm->setConstraintMode(mode == 1 ? ConstraintMode::MCMC_REFLECT : ConstraintMode::OPTIMIZATION_CLAMP); });
}
int32_t sepaihrd_host_model_current_parameters(sepaihrd_host_model* m, double* out) {
    return guarded([&] { const VectorXd v = m->pm->getCurrentParameters(); std::copy(v.data(), v.data() + v.size(), out); });
}
int32_t sepaihrd_host_model_update_parameters(sepaihrd_host_model* m, const double* params) {
    return guarded([&] { m->pm->updateModelParameters(VectorXd::FromPointer(params, static_cast<std::ptrdiff_t>(m->pm->getParameterCount()))); });
}
int32_t sepaihrd_host_model_simulate(sepaihrd_host_model* m, const double* init, const double* times, int32_t K, double* out) {
    return guarded([&] {
        const std::vector<double> t(times, times + K);
        if (!m->simulator)
            m->simulator = std::make_unique<AgeSEPAIHRDSimulator>(m->model, std::make_shared<Dopri5SolverStrategy>(), t.front(), t.back(), 1.0, m->abs_tol, m->rel_tol);
        const SimulationResult r = m->simulator->run(VectorXd::FromPointer(init, m->model->getStateSize()), t);
        const size_t W = static_cast<size_t>(m->model->getStateSize());
        for (size_t k = 0; k < r.solution.size(); ++k) std::copy(r.solution[k].begin(), r.solution[k].end(), out + k * W);
    });
}
int32_t sepaihrd_host_model_calibrate(sepaihrd_host_model* m, const char* phase1, int32_t n1, const char* const* k1, const double* v1, int32_t n2,
                                      const char* const* k2, const double* v2, double* out_best, double* out_value, int64_t* out_samples) {
    return guarded([&] {
        const std::string which = phase1 ? phase1 : "pso";
        if (which != "hill" && which != "pso" && which != "nuts") throw InvalidParameterException("sepaihrd_host_model_calibrate", "phase1 must be pso, hill or nuts");
        ModelCalibrator c = (which == "hill")   ? m->calibration->runHillClimbingMCMC(settings_map(n1, k1, v1), settings_map(n2, k2, v2))
                            : (which == "nuts") ? m->calibration->runNUTS(settings_map(n2, k2, v2))          // single phase: the second settings map
                                                : m->calibration->runPSOMCMC(settings_map(n1, k1, v1), settings_map(n2, k2, v2));
        const VectorXd& b = c.getBestParameterVector();
        if (out_best) std::copy(b.data(), b.data() + b.size(), out_best);
        if (out_value) *out_value = c.getBestObjectiveValue();
        if (out_samples) *out_samples = static_cast<int64_t>(c.getMCMCSamples().size());
    });
}
int32_t sepaihrd_host_model_metropolis(sepaihrd_host_model* m, int32_t n, const char* const* keys, const double* values, const double* initial,
                                       double* out_best, double* out_value, double* out_last, double* out_stats) {
    return guarded([&] {
        const auto P = static_cast<std::ptrdiff_t>(m->pm->getParameterCount());
        MetropolisHastingsSampler mh;
        mh.configure(settings_map(n, keys, values));
        int64_t l0 = 0, s0 = 0, l1 = 0, s1 = 0;
        sepaihrd_get_counters(m->objective->device().get(), &l0, &s0);
        const auto t0 = std::chrono::steady_clock::now();
        const OptimizationResult r = mh.optimize(VectorXd::FromPointer(initial, P), *m->objective, *m->pm);
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        sepaihrd_get_counters(m->objective->device().get(), &l1, &s1);
        if (out_best) std::copy(r.bestParameters.data(), r.bestParameters.data() + P, out_best);
        if (out_value) *out_value = r.bestObjectiveValue;
        if (out_last) { std::copy(mh.currentPositions(), mh.currentPositions() + P, out_last); out_last[P] = mh.currentLogPost()[0]; }
        if (out_stats) {
            out_stats[0] = ms;
            out_stats[1] = static_cast<double>(s1 - s0);                               // parameter sets the device evaluated
            out_stats[2] = static_cast<double>(l1 - l0);                               // kernel launches
            out_stats[3] = r.additionalStats.at("acceptance_rate");
            out_stats[4] = r.additionalStats.at("final_scale");
            out_stats[5] = static_cast<double>(mh.iteration());
        }
    });
}
int32_t sepaihrd_host_model_posterior_predictive(sepaihrd_host_model* m, const double* samples, int64_t S, int32_t num_samples, uint32_t seed,
                                                 const double* init, double* out, int64_t* out_used) {
    return guarded([&] {
        const auto P = static_cast<std::ptrdiff_t>(m->pm->getParameterCount());
        std::vector<VectorXd> sv;
        for (int64_t i = 0; i < S; ++i) sv.push_back(VectorXd::FromPointer(samples + i * P, P));
        const PosteriorPredictiveData d = ResultAggregator().aggregatePosteriorPredictives(
            sv, *m->pm, num_samples, m->times, VectorXd::FromPointer(init, m->model->getStateSize()), *m->data, m->model, seed);
        const PosteriorPredictiveData::IncidenceData* series[6] = {&d.daily_hospitalizations, &d.daily_icu_admissions, &d.daily_deaths,
                                                                   &d.cumulative_hospitalizations, &d.cumulative_icu_admissions, &d.cumulative_deaths};
        const std::ptrdiff_t T = static_cast<std::ptrdiff_t>(d.time_points.size());
        const int n = m->model->getNumAgeClasses();
        for (int s = 0; s < 6; ++s) {
            const MatrixXd* q[5] = {&series[s]->lower_95, &series[s]->lower_90, &series[s]->median, &series[s]->upper_90, &series[s]->upper_95};
            for (std::ptrdiff_t t = 0; t < T; ++t)
                for (int a = 0; a < n; ++a)
                    for (int k = 0; k < 5; ++k) out[((s * T + t) * n + a) * 5 + k] = (*q[k])(t, a);
        }
        if (out_used) *out_used = d.samples_used;
    });
}
int32_t sepaihrd_host_model_gradient(sepaihrd_host_model* m, const double* params, double epsilon, double* out_value, double* out_grad) {
    return guarded([&] {
        if (!m || !params || !out_grad) throw InvalidParameterException("sepaihrd_host_model_gradient", "bad argument");
        SEPAIHRDGradientObjectiveFunction g(m->model, *m->pm, *m->cache, *m->data, m->times, m->data->getInitialSEPAIHRDState(),
                                            std::make_shared<Dopri5SolverStrategy>(), m->abs_tol, m->rel_tol);
        if (epsilon > 0) g.epsilon_ = epsilon;
        const auto P = static_cast<std::ptrdiff_t>(m->pm->getParameterCount());
        VectorXd grad;
        const double v = g.evaluate_with_gradient(VectorXd::FromPointer(params, P), grad);
        if (out_value) *out_value = v;
        std::copy(grad.data(), grad.data() + P, out_grad);
    });
}
// ---- AnalysisWriter files from plain arrays (host only) ---------------------------------------------------------------------
int32_t sepaihrd_host_write_posterior_predictive(const char* output_dir, int32_t T, int32_t n_ages, const double* time_points,
                                                 const double* quantiles, const double* observed) {
    return guarded([&] {
        if (!output_dir || !time_points || !quantiles || T < 0 || n_ages < 0) throw InvalidParameterException("sepaihrd_host_write_posterior_predictive", "bad argument");
        PosteriorPredictiveData d;
        d.time_points.assign(time_points, time_points + T);
        PosteriorPredictiveData::IncidenceData* series[6] = {&d.daily_hospitalizations, &d.daily_icu_admissions, &d.daily_deaths,
                                                             &d.cumulative_hospitalizations, &d.cumulative_icu_admissions, &d.cumulative_deaths};
        for (int s = 0; s < 6; ++s) {
            MatrixXd* q[5] = {&series[s]->lower_95, &series[s]->lower_90, &series[s]->median, &series[s]->upper_90, &series[s]->upper_95};
            for (int k = 0; k < 5; ++k) {
                *q[k] = MatrixXd::Zero(T, n_ages);
                for (int t = 0; t < T; ++t) for (int a = 0; a < n_ages; ++a) (*q[k])(t, a) = quantiles[(((size_t)s * T + t) * n_ages + a) * 5 + k];
            }
            if (observed) {
                series[s]->observed = MatrixXd::Zero(T, n_ages);
                for (int t = 0; t < T; ++t) for (int a = 0; a < n_ages; ++a) series[s]->observed(t, a) = observed[((size_t)s * T + t) * n_ages + a];
            }
        }
        AnalysisWriter().savePosteriorPredictiveData(output_dir, d);
    });
}
int32_t sepaihrd_host_write_parameter_posteriors(const char* output_dir, const double* samples, int64_t S, int32_t P, const char* const* names,
                                                 int32_t burn_in, int32_t thinning) {
    return guarded([&] {
        if (!output_dir || (S > 0 && !samples) || !names || P < 0) throw InvalidParameterException("sepaihrd_host_write_parameter_posteriors", "bad argument");
        std::vector<VectorXd> sv;
        for (int64_t i = 0; i < S; ++i) sv.push_back(VectorXd::FromPointer(samples + i * P, P));
        std::vector<std::string> nm;
        for (int32_t j = 0; j < P; ++j) nm.emplace_back(names[j]);
        AnalysisWriter().saveParameterPosteriors(output_dir, sv, nm, burn_in, thinning);
    });
}
int32_t sepaihrd_host_model_set_cache(sepaihrd_host_model* m, int64_t capacity) {
    return guarded([&] {
        if (!m || capacity < 0) throw InvalidParameterException("sepaihrd_host_model_set_cache", "bad argument");
        if (capacity == 0) m->build(std::make_shared<NullSimulationCache>());
        else m->build(std::make_shared<SimulationCache>(static_cast<size_t>(capacity)));
    });
}
int32_t sepaihrd_host_model_cache_stats(const sepaihrd_host_model* m, int64_t* out) {
    return guarded([&] {
        if (!m || !out) throw InvalidParameterException("sepaihrd_host_model_cache_stats", "bad argument");
        const auto* sc = dynamic_cast<const SimulationCache*>(m->cache.get());
        out[0] = static_cast<int64_t>(m->cache->size());
        out[1] = sc ? static_cast<int64_t>(sc->getLikelihoodCalls()) : 0;
        out[2] = sc ? static_cast<int64_t>(sc->getLikelihoodHits()) : 0;
        out[3] = sc ? static_cast<int64_t>(sc->storeLikelihoodCalls()) : 0;
    });
}
void sepaihrd_host_model_destroy(sepaihrd_host_model* m) { delete m; }

// ---- SimulationCache on its own (host only) -----------------------------------------------------------------------------
int32_t sepaihrd_host_cache_create(int64_t capacity, sepaihrd_host_cache** out) {
    return guarded([&] {
        if (!out || capacity < 0) throw InvalidParameterException("sepaihrd_host_cache_create", "bad argument");
        *out = new sepaihrd_host_cache(static_cast<size_t>(capacity));      // capacity 0: "SimulationCache: max_size must be > 0."
    });
}
uint64_t sepaihrd_host_cache_hash(const sepaihrd_host_cache* c, const double* params, int32_t n) {
    return static_cast<uint64_t>(c->c.computeHash(params, n));
}
int32_t sepaihrd_host_cache_get(sepaihrd_host_cache* c, uint64_t key, double* out_value) {
    double v = 0.0;
    const bool hit = c->c.getLikelihood(static_cast<size_t>(key), v);
    if (hit && out_value) *out_value = v;
    return hit ? 1 : 0;
}
void sepaihrd_host_cache_store(sepaihrd_host_cache* c, uint64_t key, double value) { c->c.storeLikelihood(static_cast<size_t>(key), value); }
int32_t sepaihrd_host_cache_get_vector(sepaihrd_host_cache* c, const double* params, int32_t n, double* out_value) {
    const auto v = c->c.get(VectorXd::FromPointer(params, n));
    if (v && out_value) *out_value = *v;
    return v ? 1 : 0;
}
void sepaihrd_host_cache_set_vector(sepaihrd_host_cache* c, const double* params, int32_t n, double value) { c->c.set(VectorXd::FromPointer(params, n), value); }
int32_t sepaihrd_host_cache_batch(sepaihrd_host_cache* c, int32_t n_params, const double* params, int64_t B, int64_t ld, double* out,
                                  sepaihrd_host_batch_fn fn, void* user, const uint32_t* status_of_row) {
    return guarded([&] {
        if (!c || !fn || !params || !out) throw InvalidParameterException("sepaihrd_host_cache_batch", "bad argument");
        evaluateThroughCache(c->c, n_params, params, B, ld, out, [&](const double* rows, int64_t M, int64_t rld, double* vals, uint32_t* st) {
            if (fn(user, rows, M, rld, vals) != 0) throw SimulationException("sepaihrd_host_cache_batch", "the batch callback reported a failure");
            // test hook: a status word per ORIGINAL row, looked up by the row's first coordinate used as an index
            for (int64_t j = 0; j < M; ++j) st[j] = status_of_row ? status_of_row[static_cast<int64_t>(rows[j * rld])] : 0u;
        });
    });
}
int64_t sepaihrd_host_cache_size(const sepaihrd_host_cache* c) { return static_cast<int64_t>(c->c.size()); }
void sepaihrd_host_cache_clear(sepaihrd_host_cache* c) { c->c.clear(); }
void sepaihrd_host_cache_stats(const sepaihrd_host_cache* c, int64_t* out) {
    out[0] = static_cast<int64_t>(c->c.getLikelihoodCalls()); out[1] = static_cast<int64_t>(c->c.getLikelihoodHits());
    out[2] = static_cast<int64_t>(c->c.storeLikelihoodCalls());
}
void sepaihrd_host_cache_destroy(sepaihrd_host_cache* c) { delete c; }

}  // extern "C"

// ---- the reference's on-disk formats ----------------------------------------------------------------------------
#include "config_io.hpp"

#include <cstdio>

namespace {

thread_local std::string g_json;

struct Json {
    std::string s;
    static std::string num(double v) {
        if (!std::isfinite(v)) return "null";
        char buf[40];
        std::snprintf(buf, sizeof buf, "%.17g", v);
        return buf;
    }
    static std::string str(const std::string& v) {
        std::string o = "\"";
        for (char c : v) {
            if (c == '"' || c == '\\') { o += '\\'; o += c; }
            else if (static_cast<unsigned char>(c) < 0x20) { char b[8]; std::snprintf(b, sizeof b, "\\u%04x", c); o += b; }
            else o += c;
        }
        return o + "\"";
    }
    template <class It> static std::string nums(It b, It e) {
        std::string o = "[";
        for (It i = b; i != e; ++i) { if (i != b) o += ','; o += num(*i); }
        return o + "]";
    }
    static std::string nums(const VectorXd& v) { return nums(v.data(), v.data() + v.size()); }
    static std::string nums(const std::vector<double>& v) { return nums(v.begin(), v.end()); }
    static std::string rowmajor(const MatrixXd& m) {
        std::vector<double> f;
        for (std::ptrdiff_t i = 0; i < m.rows(); ++i) for (std::ptrdiff_t j = 0; j < m.cols(); ++j) f.push_back(m(i, j));
        return nums(f);
    }
    static std::string strs(const std::vector<std::string>& v) {
        std::string o = "[";
        for (size_t i = 0; i < v.size(); ++i) { if (i) o += ','; o += str(v[i]); }
        return o + "]";
    }
    void field(const char* key, const std::string& value) { s += (s.empty() ? "{" : ",") + str(key) + ":" + value; }
    std::string done() { return s.empty() ? "{}" : s + "}"; }
};

std::string parameters_json(const SEPAIHRDParameters& p) {
    Json j;
    j.field("beta", Json::num(p.beta)); j.field("theta", Json::num(p.theta)); j.field("sigma", Json::num(p.sigma));
    j.field("gamma_p", Json::num(p.gamma_p)); j.field("gamma_A", Json::num(p.gamma_A)); j.field("gamma_I", Json::num(p.gamma_I));
    j.field("gamma_H", Json::num(p.gamma_H)); j.field("gamma_ICU", Json::num(p.gamma_ICU));
    j.field("E0_multiplier", Json::num(p.E0_multiplier)); j.field("P0_multiplier", Json::num(p.P0_multiplier));
    j.field("A0_multiplier", Json::num(p.A0_multiplier)); j.field("I0_multiplier", Json::num(p.I0_multiplier));
    j.field("H0_multiplier", Json::num(p.H0_multiplier)); j.field("ICU0_multiplier", Json::num(p.ICU0_multiplier));
    j.field("R0_multiplier", Json::num(p.R0_multiplier)); j.field("D0_multiplier", Json::num(p.D0_multiplier));
    j.field("runup_days", Json::num(p.runup_days)); j.field("seed_exposed", Json::num(p.seed_exposed));
    j.field("a", Json::nums(p.a)); j.field("h_infec", Json::nums(p.h_infec)); j.field("p", Json::nums(p.p)); j.field("h", Json::nums(p.h));
    j.field("icu", Json::nums(p.icu)); j.field("d_H", Json::nums(p.d_H)); j.field("d_ICU", Json::nums(p.d_ICU));
    j.field("d_community", Json::nums(p.d_community));
    j.field("beta_end_times", Json::nums(p.beta_end_times)); j.field("beta_values", Json::nums(p.beta_values));
    j.field("kappa_end_times", Json::nums(p.kappa_end_times)); j.field("kappa_values", Json::nums(p.kappa_values));
    return j.done();
}

}  // namespace

extern "C" const char* sepaihrd_host_read_file_json(const char* kind, const char* filename, int32_t a, int32_t b, const char* start_date,
                                                    const char* end_date) {
    const int32_t rc = guarded([&] {
        const std::string k = kind ? kind : "", f = filename ? filename : "";
        Json j;
        if (k == "parameters") {
            g_json = parameters_json(readSEPAIHRDParameters(f, a));
            return;
        } else if (k == "bounds") {
            for (const auto& [name, lh] : readParamBounds(f)) j.field(name.c_str(), "[" + Json::num(lh.first) + "," + Json::num(lh.second) + "]");
        } else if (k == "sigmas") {
            for (const auto& [name, v] : readProposalSigmas(f)) j.field(name.c_str(), Json::num(v));
        } else if (k == "settings") {
            for (const auto& [name, v] : readMetropolisHastingsSettings(f)) j.field(name.c_str(), Json::num(v));
        } else if (k == "names") {
            j.field("names", Json::strs(readParamsToCalibrate(f)));
        } else if (k == "matrix") {
            j.field("rowmajor", Json::rowmajor(readMatrixFromCSV(f, a, b)));
        } else if (k == "data") {
            const CalibrationDataFile d(f, start_date ? start_date : "", end_date ? end_date : "");
            j.field("dates", Json::strs(d.getDates()));
            j.field("population", Json::nums(d.getPopulationByAgeGroup()));
            j.field("new_confirmed", Json::rowmajor(d.getNewConfirmedCases()));
            j.field("new_hospitalizations", Json::rowmajor(d.getNewHospitalizations()));
            j.field("new_icu", Json::rowmajor(d.getNewICU()));
            j.field("new_deaths", Json::rowmajor(d.getNewDeaths()));
            j.field("cumulative_confirmed", Json::rowmajor(d.getCumulativeConfirmedCases()));
            j.field("cumulative_deaths", Json::rowmajor(d.getCumulativeDeaths()));
            j.field("cumulative_hospitalizations", Json::rowmajor(d.getCumulativeHospitalizations()));
            j.field("cumulative_icu", Json::rowmajor(d.getCumulativeICU()));
        } else {
            throw InvalidParameterException("sepaihrd_host_read_file_json", "unknown kind: " + k);
        }
        g_json = j.done();
    });
    return rc == 0 ? g_json.c_str() : nullptr;
}

extern "C" const char* sepaihrd_host_project_json(const char* project_root, const char* start_date, const char* end_date, int32_t n_ages) {
    const int32_t rc = guarded([&] {
        const ReferenceProject prj = loadReferenceProject(project_root ? project_root : "", start_date ? start_date : "",
                                                          end_date ? end_date : "", n_ages);
        SEPAIHRDParameterManager pm(prj.model, prj.params_to_calibrate, prj.proposal_sigmas, prj.param_bounds);   // validates names / sigmas / bounds
        std::vector<double> lo, hi, sg;
        for (size_t i = 0; i < prj.params_to_calibrate.size(); ++i) {
            lo.push_back(pm.getLowerBoundForParamIndex(static_cast<int>(i)));
            hi.push_back(pm.getUpperBoundForParamIndex(static_cast<int>(i)));
            sg.push_back(pm.getSigmaForParamIndex(static_cast<int>(i)));
        }
        Json j;
        j.field("format", Json::str("sepaihrd_problem/1"));
        j.field("n_ages", std::to_string(n_ages));
        j.field("times", Json::nums(prj.time_points));
        j.field("obs_hosp", Json::rowmajor(prj.data->getNewHospitalizations()));
        j.field("obs_icu", Json::rowmajor(prj.data->getNewICU()));
        j.field("obs_deaths", Json::rowmajor(prj.data->getNewDeaths()));
        j.field("population", Json::nums(prj.data->getPopulationByAgeGroup()));
        j.field("contact_matrix_rowmajor", Json::rowmajor(prj.model->getContactMatrix()));
        j.field("beta_end_times", Json::nums(prj.params.beta_end_times));
        j.field("kappa_end_times", Json::nums(prj.model->kappaEndTimes()));
        j.field("base_slots", Json::nums(prj.model->slotVector()));
        j.field("data_initial_state", Json::nums(prj.data_initial_state));
        j.field("initial_state", Json::nums(prj.initial_state));
        j.field("param_names", Json::strs(prj.params_to_calibrate));
        j.field("lower_bound", Json::nums(lo));
        j.field("upper_bound", Json::nums(hi));
        j.field("sigmas", Json::nums(sg));
        j.field("constraint_mode", "0");
        j.field("abs_tol", Json::num(prj.abs_error));
        j.field("rel_tol", Json::num(prj.rel_error));
        j.field("dt_hint", Json::num(prj.dt_hint));
        g_json = j.done();
    });
    return rc == 0 ? g_json.c_str() : nullptr;
}

extern "C" int32_t sepaihrd_host_resave_parameters(const char* in_file, int32_t n_ages, const char* out_file, int32_t n_calibrated,
                                                   const char* const* calibrated_names, double obj_value, const char* timestamp) {
    return guarded([&] {
        const SEPAIHRDParameters p = readSEPAIHRDParameters(in_file ? in_file : "", n_ages);
        std::vector<std::string> names;
        for (int32_t i = 0; i < n_calibrated; ++i) names.emplace_back(calibrated_names[i]);
        saveCalibrationResults(out_file ? out_file : "", p, names, obj_value, timestamp ? timestamp : "");
    });
}


// ---- post-calibration analysis: essential metrics and NPI scenarios -------------------------------------------------
namespace {

void store_metrics(const EssentialMetrics& m, int n, int nk, double* scalars, double* age, double* kappa) {
    if (scalars) {
        const double v[SEPAIHRD_HOST_NUM_METRICS] = {m.R0, m.overall_IFR, m.overall_attack_rate, m.peak_hospital_occupancy, m.peak_ICU_occupancy,
                                                     m.time_to_peak_hospital, m.time_to_peak_ICU, m.total_cumulative_deaths, m.max_Rt, m.min_Rt,
                                                     m.final_Rt, m.seroprevalence_at_target_day};
        std::copy(v, v + SEPAIHRD_HOST_NUM_METRICS, scalars);
    }
    if (age) {
        const std::vector<double>* blocks[4] = {&m.age_specific_IFR, &m.age_specific_IHR, &m.age_specific_IICUR, &m.age_specific_attack_rate};
        for (int b = 0; b < 4; ++b) std::copy(blocks[b]->begin(), blocks[b]->begin() + n, age + b * n);
    }
    if (kappa)
        for (int k = 0; k < nk; ++k) {
            const auto it = m.kappa_values.find("kappa_" + std::to_string(k + 1));
            kappa[k] = (it != m.kappa_values.end()) ? it->second : std::nan("");
        }
}

}  // namespace

extern "C" int32_t sepaihrd_host_metrics(const sepaihrd_problem* pb, const double* times, int32_t K, const double* trajectory,
                                         const double* initial_state, double* out_scalars, double* out_age, double* out_rt, double* out_sero) {
    return guarded([&] {
        if (!pb || !times || K <= 0 || !trajectory || !initial_state) throw InvalidParameterException("sepaihrd_host_metrics", "bad argument");
        SEPAIHRDParameters p;
        auto model = model_from_problem(pb, {}, &p);
        const int n = pb->n_ages, W = SEPAIHRD_NUM_COMPARTMENTS * n;
        SimulationResult sim;
        sim.time_points.assign(times, times + K);
        sim.num_age_classes = n;
        sim.solution.resize(static_cast<size_t>(K));
        for (int k = 0; k < K; ++k) sim.solution[static_cast<size_t>(k)].assign(trajectory + static_cast<size_t>(k) * W, trajectory + static_cast<size_t>(k + 1) * W);
        const VectorXd x0 = VectorXd::FromPointer(initial_state, W);
        const SEPAIHRDParameters params = model->getModelParameters();
        MetricsCalculator mc;
        store_metrics(mc.calculateEssentialMetrics(sim, model, params, x0, sim.time_points), n, 0, out_scalars, out_age, nullptr);
        if (out_rt) { const auto rt = mc.calculateRtTrajectory(sim, model, sim.time_points); std::copy(rt.begin(), rt.end(), out_rt); }
        if (out_sero) { const auto se = mc.calculateSeroprevalenceTrajectory(sim, params, sim.time_points); std::copy(se.begin(), se.end(), out_sero); }
    });
}

extern "C" int32_t sepaihrd_host_model_scenarios(sepaihrd_host_model* m, const double* samples, int64_t S, int32_t burn_in, int32_t thinning,
                                                 const double* initial_state, double* out_scalars, double* out_age, double* out_kappa,
                                                 double* out_trajectories, const char* csv_path) {
    return guarded([&] {
        if (!m || !samples || S <= 0 || !initial_state) throw InvalidParameterException("sepaihrd_host_model_scenarios", "bad argument");
        const auto P = static_cast<std::ptrdiff_t>(m->pm->getParameterCount());
        const int n = m->model->getNumAgeClasses(), nk = m->model->numKappa(), W = m->model->getStateSize();
        std::vector<VectorXd> smp;
        for (int64_t i = 0; i < S; ++i) smp.push_back(VectorXd::FromPointer(samples + i * P, P));
        PostCalibrationAnalyser analyser(m->model, std::make_shared<Dopri5SolverStrategy>(), m->times, VectorXd::FromPointer(initial_state, W),
                                         m->abs_tol, m->rel_tol);
        std::vector<SimulationResult> runs;
        const auto results = analyser.scenarioAnalysisFromSamples(smp, *m->pm, burn_in, thinning, &runs);
        for (size_t r = 0; r < results.size(); ++r) {
            store_metrics(results[r].second, n, nk, out_scalars ? out_scalars + r * SEPAIHRD_HOST_NUM_METRICS : nullptr,
                          out_age ? out_age + r * 4 * static_cast<size_t>(n) : nullptr, out_kappa ? out_kappa + r * static_cast<size_t>(nk) : nullptr);
            if (out_trajectories)
                for (size_t k = 0; k < runs[r].solution.size(); ++k)
                    std::copy(runs[r].solution[k].begin(), runs[r].solution[k].end(), out_trajectories + (r * runs[r].solution.size() + k) * static_cast<size_t>(W));
        }
        if (csv_path && *csv_path) PostCalibrationAnalyser::writeScenarioComparison(csv_path, results);
    });
}

extern "C" int32_t sepaihrd_host_model_analyze_runs(sepaihrd_host_model* m, const double* samples, int64_t S, int32_t burn_in, int32_t thinning,
                                                    const double* initial_state, double* out_scalars, double* out_rt_quantiles,
                                                    double* out_sero_quantiles, int64_t* out_runs) {
    return guarded([&] {
        if (!m || !samples || S <= 0 || !initial_state) throw InvalidParameterException("sepaihrd_host_model_analyze_runs", "bad argument");
        const auto P = static_cast<std::ptrdiff_t>(m->pm->getParameterCount());
        const int W = m->model->getStateSize();
        std::vector<VectorXd> smp;
        for (int64_t i = 0; i < S; ++i) smp.push_back(VectorXd::FromPointer(samples + i * P, P));
        PostCalibrationAnalyser analyser(m->model, std::make_shared<Dopri5SolverStrategy>(), m->times, VectorXd::FromPointer(initial_state, W),
                                         m->abs_tol, m->rel_tol);
        const auto res = analyser.analyzeMCMCRuns(smp, *m->pm, burn_in, thinning);
        if (out_runs) *out_runs = static_cast<int64_t>(res.metrics.size());
        for (size_t r = 0; r < res.metrics.size(); ++r)
            store_metrics(res.metrics[r], 0, 0, out_scalars ? out_scalars + r * SEPAIHRD_HOST_NUM_METRICS : nullptr, nullptr, nullptr);
        auto put = [&](const AggregatedTrajectory& a, double* out) {
            if (!out) return;
            const size_t K = a.median.size();
            const std::vector<double>* q[5] = {&a.q025, &a.q05, &a.median, &a.q95, &a.q975};
            for (int j = 0; j < 5; ++j) std::copy(q[j]->begin(), q[j]->end(), out + static_cast<size_t>(j) * K);
        };
        put(res.rt, out_rt_quantiles);
        put(res.seroprevalence, out_sero_quantiles);
    });
}
