// epidemic_host.cpp -- host mirror of the reference's model / parameter-manager / simulator / objective classes
// over the C ABI of include/sepaihrd_b200.h.  See epidemic_host.hpp for the reference files each class mirrors.
#include "epidemic_host.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <random>

namespace epidemic {

namespace {

const char* const kCompartments[SEPAIHRD_NUM_COMPARTMENTS] = {"S", "E", "P", "A", "I", "H", "ICU", "R", "D", "CumH", "CumICU"};

struct SlotLayout {   // same layout as sepaihrd_b200.h
    int n, nb, nk;
    int beta0() const { return 0; }
    int kappa0() const { return nb; }
    int scal0() const { return nb + nk; }
    int age0() const { return scal0() + 7; }
    int mult0() const { return age0() + 8 * n; }
    int seed() const { return mult0() + 8; }
    int runup() const { return mult0() + 9; }
    int beta_scalar() const { return mult0() + 10; }
    int count() const { return mult0() + 11; }
};

void check_size(const VectorXd& v, int n, const char* what) {
    if (v.size() != n) throw InvalidParameterException("AgeSEPAIHRDModel", std::string("size mismatch for ") + what);
}

[[noreturn]] void throw_capi(const char* where) {
    throw ModelConstructionException(where, std::string("B200 evaluator: ") + sepaihrd_last_error());
}

}  // namespace

// ---- SEPAIHRDParameters ------------------------------------------------------------------------------------
bool SEPAIHRDParameters::validate() const {
    const auto n = N.size();
    if (n == 0) return false;
    if (M_baseline.rows() != n || M_baseline.cols() != n) return false;
    for (const VectorXd* v : {&p, &h, &icu, &d_H, &d_ICU, &a, &h_infec})
        if (v->size() != n) return false;
    if (d_community.size() > 0 && d_community.size() != n) return false;
    if (kappa_end_times.size() != kappa_values.size() || beta_end_times.size() != beta_values.size()) return false;
    double prev = -std::numeric_limits<double>::infinity();
    for (double t : beta_end_times) {
        if (!(t > prev)) return false;
        prev = t;
    }
    return true;
}

// ---- PiecewiseConstantNpiStrategy --------------------------------------------------------------------------
PiecewiseConstantNpiStrategy::PiecewiseConstantNpiStrategy(const std::vector<double>& end_times, const std::vector<double>& values,
                                                           const std::map<std::string, std::pair<double, double>>& bounds,
                                                           double baseline_kappa, double baseline_end, bool fixed_baseline,
                                                           const std::vector<std::string>& names)
    : baseline_kappa_value_(baseline_kappa), is_baseline_fixed_(fixed_baseline), baseline_period_end_time_(baseline_end),
      npi_period_end_times_(end_times), npi_kappa_values_(values), npi_param_names_(names), param_bounds_map_(bounds) {
    const char* src = "PiecewiseConstantNpiStrategy";
    if (end_times.size() != values.size()) throw InvalidParameterException(src, "NPI end times and values must have the same size.");
    if (baseline_kappa < 0.0) throw InvalidParameterException(src, "Baseline kappa must be non-negative.");
    double prev = baseline_end;
    for (size_t i = 0; i < end_times.size(); ++i) {
        if (!(end_times[i] > prev)) throw InvalidParameterException(src, "NPI end times must be strictly increasing and after the baseline period.");
        if (values[i] < 0.0) throw InvalidParameterException(src, "NPI kappa values must be non-negative.");
        prev = end_times[i];
    }
    if (npi_param_names_.empty())
        for (size_t i = 0; i < values.size(); ++i) npi_param_names_.push_back("kappa_" + std::to_string(i + 2));
    if (npi_param_names_.size() != values.size()) throw InvalidParameterException(src, "One name per NPI value is required.");
}

double PiecewiseConstantNpiStrategy::getReductionFactor(double time) const {
    if (time < 0.0 || time <= baseline_period_end_time_) return baseline_kappa_value_;
    for (size_t k = 0; k < npi_period_end_times_.size(); ++k)
        if (time <= npi_period_end_times_[k]) return npi_kappa_values_[k];
    return npi_kappa_values_.empty() ? baseline_kappa_value_ : npi_kappa_values_.back();
}

std::vector<double> PiecewiseConstantNpiStrategy::getValues() const {
    std::vector<double> all{baseline_kappa_value_};
    all.insert(all.end(), npi_kappa_values_.begin(), npi_kappa_values_.end());
    return all;
}

void PiecewiseConstantNpiStrategy::setValues(const std::vector<double>& v) {
    if (v.size() != npi_kappa_values_.size())
        throw InvalidParameterException("PiecewiseConstantNpiStrategy::setValues", "New NPI values vector size must match existing number of changeable NPI periods.");
    for (double k : v)
        if (k < 0.0) throw InvalidParameterException("PiecewiseConstantNpiStrategy::setValues", "NPI kappa values must be non-negative.");
    npi_kappa_values_ = v;
}

size_t PiecewiseConstantNpiStrategy::getNumCalibratableNpiParams() const { return npi_kappa_values_.size() + (is_baseline_fixed_ ? 0 : 1); }

std::string PiecewiseConstantNpiStrategy::getNpiParamName(int idx) const {
    if (idx < 0 || static_cast<size_t>(idx) >= getNumCalibratableNpiParams())
        throw InvalidParameterException("PiecewiseConstantNpiStrategy::getNpiParamName", "calibratable_idx out of range.");
    if (!is_baseline_fixed_) return idx == 0 ? std::string("kappa_baseline") : npi_param_names_[idx - 1];
    return npi_param_names_[idx];
}

void PiecewiseConstantNpiStrategy::setCalibratableValues(const std::vector<double>& v) {
    const char* src = "PiecewiseConstantNpiStrategy::setCalibratableValues";
    if (v.size() != getNumCalibratableNpiParams()) throw InvalidParameterException(src, "Input vector size does not match the number of calibratable NPI parameters.");
    for (double k : v)
        if (k < 0.0) throw InvalidParameterException(src, "All NPI kappa values must be non-negative.");
    size_t i = 0;
    if (!is_baseline_fixed_) baseline_kappa_value_ = v[i++];
    for (size_t k = 0; k < npi_kappa_values_.size(); ++k) npi_kappa_values_[k] = v[i++];
}

std::vector<double> PiecewiseConstantNpiStrategy::getCalibratableValues() const {
    std::vector<double> v;
    if (!is_baseline_fixed_) v.push_back(baseline_kappa_value_);
    v.insert(v.end(), npi_kappa_values_.begin(), npi_kappa_values_.end());
    return v;
}

double PiecewiseConstantNpiStrategy::getLowerBoundForParamIndex(int idx) const {
    auto it = param_bounds_map_.find(getNpiParamName(idx));
    return it != param_bounds_map_.end() ? it->second.first : 0.0;     // constants::DEFAULT_NPI_LOWER_BOUND
}
double PiecewiseConstantNpiStrategy::getUpperBoundForParamIndex(int idx) const {
    auto it = param_bounds_map_.find(getNpiParamName(idx));
    return it != param_bounds_map_.end() ? it->second.second : 1.0;    // constants::DEFAULT_NPI_UPPER_BOUND
}

// ---- AgeSEPAIHRDModel ---------------------------------------------------------------------------------------
AgeSEPAIHRDModel::AgeSEPAIHRDModel(const SEPAIHRDParameters& params, std::shared_ptr<INpiStrategy> npi)
    : params_(params), npi_strategy_(std::move(npi)) {
    if (!npi_strategy_) throw InvalidParameterException("AgeSEPAIHRDModel", "NPI strategy pointer cannot be null.");
    if (!params_.validate()) throw InvalidParameterException("AgeSEPAIHRDModel", "Invalid SEPAIHRD parameters dimensions or NPI values.");
    for (std::ptrdiff_t i = 0; i < params_.N.size(); ++i)
        if (!(params_.N(i) > 0.0)) throw InvalidParameterException("AgeSEPAIHRDModel", "All population sizes N must be positive.");
    if (params_.d_community.size() == 0) params_.d_community = VectorXd::Zero(params_.N.size());
}

std::shared_ptr<AgeSEPAIHRDModel> AgeSEPAIHRDModel::clone() const {
    return std::make_shared<AgeSEPAIHRDModel>(getModelParameters(), npi_strategy_->clone());
}

std::vector<std::string> AgeSEPAIHRDModel::getStateNames() const {
    std::vector<std::string> names;
    for (const char* c : kCompartments)
        for (int i = 0; i < getNumAgeClasses(); ++i) names.push_back(std::string(c) + std::to_string(i));
    return names;
}

SEPAIHRDParameters AgeSEPAIHRDModel::getModelParameters() const {
    SEPAIHRDParameters p = params_;
    p.kappa_values = npi_strategy_->getValues();
    p.kappa_end_times = kappaEndTimes();
    return p;
}

void AgeSEPAIHRDModel::setModelParameters(const SEPAIHRDParameters& p) {
    const int n = getNumAgeClasses();
    check_size(p.N, n, "N");
    if (p.M_baseline.rows() != n || p.M_baseline.cols() != n) throw InvalidParameterException("AgeSEPAIHRDModel::setModelParameters", "contact matrix dimension mismatch");
    for (const VectorXd* v : {&p.a, &p.h_infec, &p.p, &p.h, &p.icu, &p.d_H, &p.d_ICU}) check_size(*v, n, "age vector");
    if (p.beta_end_times.size() != p.beta_values.size()) throw InvalidParameterException("AgeSEPAIHRDModel::setModelParameters", "beta schedule size mismatch");
    const VectorXd keep_dc = params_.d_community;
    params_ = p;
    if (params_.d_community.size() == 0) params_.d_community = keep_dc;
}

double AgeSEPAIHRDModel::computeBeta(double time) const {
    const auto& e = params_.beta_end_times;
    if (e.empty()) return params_.beta;
    for (size_t k = 0; k < e.size(); ++k)
        if (time <= e[k]) return params_.beta_values[k];
    return params_.beta_values.back();
}

std::vector<double> AgeSEPAIHRDModel::kappaEndTimes() const {
    std::vector<double> t;
    if (auto pw = std::dynamic_pointer_cast<PiecewiseConstantNpiStrategy>(npi_strategy_)) t.push_back(pw->getBaselinePeriodEndTime());
    else t.push_back(0.0);
    const auto& e = npi_strategy_->getEndTimes();
    t.insert(t.end(), e.begin(), e.end());
    return t;
}

int AgeSEPAIHRDModel::numKappa() const { return static_cast<int>(npi_strategy_->getValues().size()); }

std::vector<double> AgeSEPAIHRDModel::slotVector() const {
    const int n = getNumAgeClasses();
    const SlotLayout L{n, static_cast<int>(params_.beta_values.size()), numKappa()};
    std::vector<double> s(static_cast<size_t>(L.count()), 0.0);
    std::copy(params_.beta_values.begin(), params_.beta_values.end(), s.begin() + L.beta0());
    const auto kv = npi_strategy_->getValues();
    std::copy(kv.begin(), kv.end(), s.begin() + L.kappa0());
    const double scal[7] = {params_.theta, params_.sigma, params_.gamma_p, params_.gamma_A, params_.gamma_I, params_.gamma_H, params_.gamma_ICU};
    std::copy(scal, scal + 7, s.begin() + L.scal0());
    const VectorXd* blocks[8] = {&params_.a, &params_.h_infec, &params_.p, &params_.h, &params_.icu, &params_.d_H, &params_.d_ICU, &params_.d_community};
    for (int b = 0; b < 8; ++b)
        for (int i = 0; i < n; ++i) s[L.age0() + b * n + i] = (*blocks[b])(i);
    const double mult[8] = {params_.E0_multiplier, params_.P0_multiplier, params_.A0_multiplier, params_.I0_multiplier,
                            params_.H0_multiplier, params_.ICU0_multiplier, params_.R0_multiplier, params_.D0_multiplier};
    std::copy(mult, mult + 8, s.begin() + L.mult0());
    s[L.seed()] = params_.seed_exposed;
    s[L.runup()] = params_.runup_days;
    s[L.beta_scalar()] = params_.beta;
    return s;
}

// ---- CalibrationData ------------------------------------------------------------------------------------------
CalibrationData::CalibrationData(const MatrixXd& nh, const MatrixXd& ni, const MatrixXd& nd, const VectorXd& pop, const VectorXd& init)
    : new_hospitalizations_(nh), new_icu_(ni), new_deaths_(nd), population_(pop), initial_state_(init) {
    const auto n = pop.size();
    if (n == 0) throw InvalidParameterException("CalibrationData", "Population vector cannot be empty.");
    for (const MatrixXd* m : {&nh, &ni, &nd})
        if (m->cols() != n || m->rows() != nh.rows()) throw InvalidParameterException("CalibrationData", "Input data matrices have inconsistent dimensions.");
    if (init.size() != 0 && init.size() != SEPAIHRD_NUM_COMPARTMENTS * n)
        throw InvalidParameterException("CalibrationData", "Initial state must have 11 entries per age class.");
}

// ---- interfaces ---------------------------------------------------------------------------------------------
void IObjectiveFunction::calculateBatch(const double* params, int64_t B, int64_t ld, double* out) const {
    const auto P = static_cast<std::ptrdiff_t>(getParameterNames().size());
    for (int64_t b = 0; b < B; ++b) out[b] = calculate(VectorXd::FromPointer(params + b * ld, P));
}

void Dopri5SolverStrategy::integrate(const std::function<void(const state_type&, state_type&, double)>&, state_type&,
                                     const std::vector<double>&, double, std::function<void(const state_type&, double)>, double,
                                     double) const {
    throw SimulationException("Dopri5SolverStrategy::integrate",
                              "the B200 backend integrates AgeSEPAIHRDModel on the device (AgeSEPAIHRDSimulator::run / "
                              "SEPAIHRDObjectiveFunction); an arbitrary std::function system has no device form and there is no CPU fallback");
}

// ---- SEPAIHRDParameterManager -----------------------------------------------------------------------------------
SEPAIHRDParameterManager::SEPAIHRDParameterManager(std::shared_ptr<AgeSEPAIHRDModel> model, const std::vector<std::string>& names,
                                                   const std::map<std::string, double>& sigmas,
                                                   const std::map<std::string, std::pair<double, double>>& bounds)
    : model_(std::move(model)), param_names_(names), proposal_sigmas_(sigmas), param_bounds_(bounds) {
    const char* src = "SEPAIHRDParameterManager";
    if (!model_) throw InvalidParameterException(src, "Model pointer cannot be null.");
    if (param_names_.empty()) throw InvalidParameterException(src, "Parameter names list (params_to_calibrate) cannot be empty.");
    const int n = model_->getNumAgeClasses();
    const auto mp = model_->getModelParameters();
    const int nb = static_cast<int>(mp.beta_values.size()), nk = model_->numKappa();
    for (const auto& name : param_names_) {
        if (!proposal_sigmas_.count(name)) throw InvalidParameterException(src, "Missing proposal sigma for parameter: " + name);
        if (!param_bounds_.count(name)) throw InvalidParameterException(src, "Missing bounds for parameter: " + name);
        const int32_t slot = sepaihrd_slot_for_name(n, nb, nk, name.c_str());   // same dispatch order as .cpp:197-267
        if (slot == -2) throw InvalidParameterException(src, "Parameter '" + name + "' is not a calibratable parameter of this model (fixed baseline kappa or index out of range).");
        slots_.push_back(slot);     // -1: a name updateModelParameters would only warn about (.cpp:265-267)
        const auto& b = param_bounds_.at(name);                // resolved once: applyConstraints runs per proposal (.cpp:321-335 looks the name up every time)
        lower_.push_back(b.first); upper_.push_back(b.second);
    }
}

VectorXd SEPAIHRDParameterManager::getCurrentParameters() const {
    const auto s = model_->slotVector();
    VectorXd v(static_cast<std::ptrdiff_t>(param_names_.size()));
    for (size_t i = 0; i < slots_.size(); ++i) v(static_cast<std::ptrdiff_t>(i)) = slots_[i] >= 0 ? s[static_cast<size_t>(slots_[i])] : 0.0;
    return v;
}

void SEPAIHRDParameterManager::updateModelParameters(const VectorXd& parameters) { updateModelParameters(parameters, model_); }

void SEPAIHRDParameterManager::updateModelParameters(const VectorXd& parameters, std::shared_ptr<AgeSEPAIHRDModel> target) {
    const char* src = "SEPAIHRDParameterManager::updateModelParameters";
    if (!target) throw InvalidParameterException(src, "Target model pointer cannot be null.");
    if (static_cast<size_t>(parameters.size()) != param_names_.size()) throw InvalidParameterException(src, "Parameter vector size mismatch.");
    const VectorXd c = applyConstraints(parameters);
    std::vector<double> s = target->slotVector();
    for (size_t i = 0; i < slots_.size(); ++i)
        if (slots_[i] >= 0) s[static_cast<size_t>(slots_[i])] = c(static_cast<std::ptrdiff_t>(i));
    SEPAIHRDParameters p = target->getModelParameters();
    const int n = target->getNumAgeClasses();
    const SlotLayout L{n, static_cast<int>(p.beta_values.size()), target->numKappa()};
    std::copy(s.begin() + L.beta0(), s.begin() + L.beta0() + L.nb, p.beta_values.begin());
    p.theta = s[L.scal0()]; p.sigma = s[L.scal0() + 1]; p.gamma_p = s[L.scal0() + 2]; p.gamma_A = s[L.scal0() + 3];
    p.gamma_I = s[L.scal0() + 4]; p.gamma_H = s[L.scal0() + 5]; p.gamma_ICU = s[L.scal0() + 6];
    VectorXd* blocks[8] = {&p.a, &p.h_infec, &p.p, &p.h, &p.icu, &p.d_H, &p.d_ICU, &p.d_community};
    for (int b = 0; b < 8; ++b) {
        if (blocks[b]->size() != n) *blocks[b] = VectorXd::Zero(n);
        for (int i = 0; i < n; ++i) (*blocks[b])(i) = s[L.age0() + b * n + i];
    }
    p.E0_multiplier = s[L.mult0()]; p.P0_multiplier = s[L.mult0() + 1]; p.A0_multiplier = s[L.mult0() + 2];
    p.I0_multiplier = s[L.mult0() + 3]; p.H0_multiplier = s[L.mult0() + 4]; p.ICU0_multiplier = s[L.mult0() + 5];
    p.R0_multiplier = s[L.mult0() + 6]; p.D0_multiplier = s[L.mult0() + 7];
    p.seed_exposed = s[L.seed()]; p.runup_days = s[L.runup()]; p.beta = s[L.beta_scalar()];
    target->setModelParameters(p);
    // NPI values go through the strategy, which rejects negatives (PieceWiseConstantNPIStrategy.cpp:238-242)
    bool kappa_touched = false;
    for (int32_t sl : slots_) kappa_touched |= (sl >= L.kappa0() && sl < L.kappa0() + L.nk);
    if (kappa_touched) {
        auto pw = std::dynamic_pointer_cast<PiecewiseConstantNpiStrategy>(target->getNpiStrategy());
        if (!pw) throw ModelException(src, "NPI strategy is not a PiecewiseConstantNpiStrategy.");
        std::vector<double> kv(s.begin() + L.kappa0() + (pw->isBaselineFixed() ? 1 : 0), s.begin() + L.kappa0() + L.nk);
        pw->setCalibratableValues(kv);
    }
}

double SEPAIHRDParameterManager::getSigmaForParamIndex(int index) const {
    if (index < 0 || static_cast<size_t>(index) >= param_names_.size()) throw std::out_of_range("SEPAIHRDParameterManager::getSigmaForParamIndex: Index out of bounds.");
    return proposal_sigmas_.at(param_names_[static_cast<size_t>(index)]);
}

double SEPAIHRDParameterManager::reflectBound(double value, double min_b, double max_b) {
    if (min_b >= max_b) return min_b;
    const double width = max_b - min_b;
    double y = std::fmod(value - min_b, 2.0 * width);
    if (y < 0) y += 2.0 * width;
    return (y <= width) ? min_b + y : max_b - (y - width);
}

VectorXd SEPAIHRDParameterManager::applyConstraints(const VectorXd& parameters) const {
    if (static_cast<size_t>(parameters.size()) != param_names_.size())
        throw InvalidParameterException("SEPAIHRDParameterManager::applyConstraints", "Parameter vector size mismatch.");
    VectorXd out = parameters;
    for (size_t i = 0; i < param_names_.size(); ++i) {       // every calibrated name has bounds (checked by the constructor)
        const auto k = static_cast<std::ptrdiff_t>(i);
        double lo = lower_[i], hi = upper_[i];
        if (lo > hi) std::swap(lo, hi);
        out(k) = (mode_ == ConstraintMode::OPTIMIZATION_CLAMP) ? std::min(std::max(parameters(k), lo), hi) : reflectBound(parameters(k), lo, hi);
    }
    return out;
}

int SEPAIHRDParameterManager::getIndexForParam(const std::string& name) const {
    auto it = std::find(param_names_.begin(), param_names_.end(), name);
    return it == param_names_.end() ? -1 : static_cast<int>(it - param_names_.begin());
}

double SEPAIHRDParameterManager::getLowerBoundForParamIndex(int idx) const {
    if (idx < 0 || static_cast<size_t>(idx) >= param_names_.size()) throw OutOfRangeException("SEPAIHRDParameterManager::getLowerBoundForParamIndex", "Index out of bounds.");
    return lower_[static_cast<size_t>(idx)];
}
double SEPAIHRDParameterManager::getUpperBoundForParamIndex(int idx) const {
    if (idx < 0 || static_cast<size_t>(idx) >= param_names_.size()) throw OutOfRangeException("SEPAIHRDParameterManager::getUpperBoundForParamIndex", "Index out of bounds.");
    return upper_[static_cast<size_t>(idx)];
}

void SEPAIHRDParameterManager::setConstraintMode(ConstraintMode mode) {
    mode_ = mode;
    for (auto& cb : mode_listeners_) if (cb) cb(mode);
}

// ---- DeviceContext ----------------------------------------------------------------------------------------------
DeviceContext::DeviceContext(const AgeSEPAIHRDModel& model, const std::vector<double>& times, const MatrixXd* oh, const MatrixXd* oi,
                             const MatrixXd* od, const VectorXd& data_init, const std::vector<int32_t>& slots,
                             const std::vector<double>& lower, const std::vector<double>& upper, ConstraintMode mode, double abs_error,
                             double rel_error, double dt_hint, int device) {
    const int n = model.getNumAgeClasses();
    const SEPAIHRDParameters mp = model.getModelParameters();
    const std::vector<double> base = model.slotVector();
    const std::vector<double> kend = model.kappaEndTimes();
    // runup offset: first output time >= 0 (SEPAIHRDObjectiveFunction.cpp:39-46); the observation matrices cover the rest
    int off = 0;
    while (off < static_cast<int>(times.size()) && times[static_cast<size_t>(off)] < 0.0) ++off;
    const int n_obs = static_cast<int>(times.size()) - off;
    auto rowmajor = [&](const MatrixXd* m) {
        std::vector<double> v(static_cast<size_t>(std::max(n_obs, 0)) * n, -1.0);   // -1 = skipped observation
        if (m) {
            if (m->cols() != n) throw InvalidParameterException("DeviceContext", "observation matrix has the wrong number of age classes");
            v.assign(static_cast<size_t>(m->rows()) * n, 0.0);
            for (std::ptrdiff_t r = 0; r < m->rows(); ++r)
                for (int a = 0; a < n; ++a) v[static_cast<size_t>(r) * n + a] = (*m)(r, a);
        }
        return v;
    };
    const std::vector<double> vh = rowmajor(oh), vi = rowmajor(oi), vd = rowmajor(od);
    std::vector<double> init(static_cast<size_t>(SEPAIHRD_NUM_COMPARTMENTS) * n, 0.0);
    if (data_init.size() == static_cast<std::ptrdiff_t>(init.size())) std::copy(data_init.data(), data_init.data() + init.size(), init.begin());
    else for (int a = 0; a < n; ++a) init[static_cast<size_t>(a)] = mp.N(a);    // S = N, everything else empty
    std::vector<double> lo = lower, hi = upper;

    sepaihrd_problem pb;
    std::memset(&pb, 0, sizeof(pb));
    pb.abi_version = SEPAIHRD_ABI_VERSION;
    pb.n_ages = n;
    pb.n_times = static_cast<int32_t>(times.size());
    pb.n_obs = static_cast<int32_t>(oh ? oh->rows() : n_obs);
    pb.times = times.data();
    pb.obs_hosp = vh.data(); pb.obs_icu = vi.data(); pb.obs_deaths = vd.data();
    pb.population = mp.N.data();
    pb.contact_matrix = mp.M_baseline.data();     // column-major, like Eigen's default
    pb.n_beta = static_cast<int32_t>(mp.beta_end_times.size());
    pb.n_kappa = static_cast<int32_t>(kend.size());
    pb.beta_end_times = mp.beta_end_times.data();
    pb.kappa_end_times = kend.data();
    pb.base_slots = base.data();
    pb.data_initial_state = init.data();
    pb.n_params = static_cast<int32_t>(slots.size());
    pb.constraint_mode = (mode == ConstraintMode::MCMC_REFLECT) ? 1 : 0;
    pb.param_slot = slots.data();
    pb.lower_bound = lo.data(); pb.upper_bound = hi.data();
    pb.abs_tol = abs_error; pb.rel_tol = rel_error; pb.dt_hint = dt_hint;
    if (sepaihrd_create(&pb, device, &ctx_) != SEPAIHRD_OK) throw_capi("DeviceContext");
    n_params_ = pb.n_params; n_times_ = pb.n_times; state_size_ = SEPAIHRD_NUM_COMPARTMENTS * n;
}

DeviceContext::~DeviceContext() { if (ctx_) sepaihrd_destroy(ctx_); }

void DeviceContext::setConstraintMode(ConstraintMode mode) {
    if (sepaihrd_set_constraint_mode(ctx_, mode == ConstraintMode::MCMC_REFLECT ? 1 : 0) != SEPAIHRD_OK) throw_capi("DeviceContext::setConstraintMode");
}

// ---- AgeSEPAIHRDSimulator ---------------------------------------------------------------------------------------
AgeSEPAIHRDSimulator::AgeSEPAIHRDSimulator(std::shared_ptr<AgeSEPAIHRDModel> model, std::shared_ptr<IOdeSolverStrategy> solver,
                                           double start_time, double end_time, double time_step, double abs_error, double rel_error)
    : model_(std::move(model)), solver_(std::move(solver)), start_time_(start_time), end_time_(end_time), time_step_(time_step),
      abs_err_(abs_error), rel_err_(rel_error) {
    const char* src = "Simulator";
    if (!model_) throw InvalidParameterException(src, "Model pointer cannot be null.");
    if (!solver_) throw InvalidParameterException(src, "Solver strategy pointer cannot be null.");
    if (!dynamic_cast<Dopri5SolverStrategy*>(solver_.get()))
        throw InvalidParameterException(src, "the B200 backend implements the adaptive Dopri5 strategy only (Dopri5SolverStrategy)");
    if (end_time_ < start_time_) throw InvalidParameterException(src, "End time must be greater than or equal to start time.");
    if (!(time_step_ > 0.0)) throw InvalidParameterException(src, "Time step hint must be positive.");
    if (!(abs_err_ > 0.0) || !(rel_err_ > 0.0)) throw InvalidParameterException(src, "Error tolerances must be positive.");
}

AgeSEPAIHRDSimulator::~AgeSEPAIHRDSimulator() = default;

void AgeSEPAIHRDSimulator::ensureContext(const std::vector<double>& times) {
    if (dev_ && times == dev_times_) return;
    // every slot is a "parameter" with bounds (-inf, +inf): run() feeds the model's CURRENT values, unconstrained
    const int nslots = static_cast<int>(model_->slotVector().size());
    std::vector<int32_t> slots(static_cast<size_t>(nslots));
    for (int i = 0; i < nslots; ++i) slots[static_cast<size_t>(i)] = i;
    const double inf = std::numeric_limits<double>::infinity();
    std::vector<double> lo(static_cast<size_t>(nslots), -inf), hi(static_cast<size_t>(nslots), inf);
    dev_.reset();
    dev_ = std::make_unique<DeviceContext>(*model_, times, nullptr, nullptr, nullptr, VectorXd(), slots, lo, hi,
                                           ConstraintMode::OPTIMIZATION_CLAMP, abs_err_, rel_err_, time_step_);
    dev_times_ = times;
}

std::vector<uint32_t> AgeSEPAIHRDSimulator::runBatch(const VectorXd& initial_state, const std::vector<double>& times,
                                                      const double* slot_rows, int64_t B, double* out) {
    const char* src = "Simulator::run";
    if (initial_state.size() != model_->getStateSize()) throw InvalidParameterException(src, "Initial state size does not match model state size.");
    if (times.empty()) throw InvalidParameterException(src, "Output time points vector is empty.");
    for (size_t i = 0; i < times.size(); ++i) {
        if (times[i] < start_time_ - 1e-9 || times[i] > end_time_ + 1e-9) throw InvalidParameterException(src, "Output time points must be within the simulation time range.");
        if (i > 0 && !(times[i] > times[i - 1])) throw InvalidParameterException(src, "Output time points must be strictly increasing.");
    }
    ensureContext(times);
    std::vector<uint32_t> status(static_cast<size_t>(B), 0u);
    const int64_t ld = dev_->numParams();
    if (sepaihrd_simulate_from_state(dev_->get(), slot_rows, B, ld, initial_state.data(), 0, SEPAIHRD_TRAJ_FULL, 1, out, status.data()) != SEPAIHRD_OK)
        throw SimulationException(src, std::string("device integration failed: ") + sepaihrd_last_error());
    return status;
}

SimulationResult AgeSEPAIHRDSimulator::run(const VectorXd& initial_state, const std::vector<double>& times) {
    const std::vector<double> row = model_->slotVector();
    const size_t W = static_cast<size_t>(model_->getStateSize());
    std::vector<double> out(times.size() * W);
    const auto st = runBatch(initial_state, times, row.data(), 1, out.data());
    if (st[0] & SEPAIHRD_ST_STEP_FAILURE) throw SimulationException("Simulator::run", "ODE integration failed: too many consecutive rejected steps.");
    if (st[0] & SEPAIHRD_ST_INVALID_PARAM) throw InvalidParameterException("Simulator::run", "All NPI kappa values must be non-negative.");
    SimulationResult r;
    r.time_points = times;
    r.solution.resize(times.size());
    for (size_t k = 0; k < times.size(); ++k) r.solution[k].assign(out.begin() + static_cast<std::ptrdiff_t>(k * W), out.begin() + static_cast<std::ptrdiff_t>((k + 1) * W));
    r.compartment_names.assign(kCompartments, kCompartments + SEPAIHRD_NUM_COMPARTMENTS);
    r.num_age_classes = model_->getNumAgeClasses();
    return r;
}

// ---- SEPAIHRDObjectiveFunction ------------------------------------------------------------------------------------
SEPAIHRDObjectiveFunction::SEPAIHRDObjectiveFunction(std::shared_ptr<AgeSEPAIHRDModel> model, IParameterManager& pm, ISimulationCache& cache,
                                                     const CalibrationData& data, const std::vector<double>& time_points,
                                                     const VectorXd& initial_state, std::shared_ptr<IOdeSolverStrategy> solver,
                                                     double abs_error, double rel_error)
    : parameterManager_(pm), cache_(cache), model_(std::move(model)) {
    const char* src = "SEPAIHRDObjectiveFunction";
    if (!model_) throw InvalidParameterException(src, "Model pointer is null.");
    if (!solver) throw InvalidParameterException(src, "Solver strategy pointer is null.");
    if (!dynamic_cast<Dopri5SolverStrategy*>(solver.get())) throw InvalidParameterException(src, "the B200 backend implements Dopri5SolverStrategy only");
    if (time_points.empty()) throw InvalidParameterException(src, "Time points vector is empty.");
    if (initial_state.size() != model_->getStateSize()) throw InvalidParameterException(src, "Initial state size does not match model state size.");
    auto* spm = dynamic_cast<SEPAIHRDParameterManager*>(&pm);
    if (!spm) throw InvalidParameterException(src, "parameterManager must be a SEPAIHRDParameterManager");
    std::vector<double> lo, hi;
    for (size_t i = 0; i < spm->getParameterCount(); ++i) {
        lo.push_back(spm->getLowerBoundForParamIndex(static_cast<int>(i)));
        hi.push_back(spm->getUpperBoundForParamIndex(static_cast<int>(i)));
    }
    dev_ = std::make_unique<DeviceContext>(*model_, time_points, &data.getNewHospitalizations(), &data.getNewICU(), &data.getNewDeaths(),
                                           initial_state, spm->slots(), lo, hi, spm->getConstraintMode(), abs_error, rel_error,
                                           1.0 /* the objective builds its simulator with time_step 1.0, .cpp:113 */);
    DeviceContext* dev = dev_.get();
    mode_listener_id_ = spm->onConstraintModeChange([dev](ConstraintMode m) { dev->setConstraintMode(m); });
}

SEPAIHRDObjectiveFunction::~SEPAIHRDObjectiveFunction() {
    // the parameter manager may outlive this objective: stop forwarding its constraint mode to a dead device context
    if (auto* spm = dynamic_cast<SEPAIHRDParameterManager*>(&parameterManager_)) spm->removeConstraintModeListener(mode_listener_id_);
}

const std::vector<std::string>& SEPAIHRDObjectiveFunction::getParameterNames() const { return parameterManager_.getParameterNames(); }

double SEPAIHRDObjectiveFunction::calculate(const VectorXd& parameters) const {
    if (parameters.size() != dev_->numParams()) return std::numeric_limits<double>::lowest();   // updateModelParameters throws -> caught, .cpp:117-122
    double out = 0.0;
    calculateBatch(parameters.data(), 1, parameters.size(), &out);
    return out;
}

void SEPAIHRDObjectiveFunction::evaluateRows(const double* params, int64_t B, int64_t ld, double* out, uint32_t* status, int32_t* steps) const {
    if (sepaihrd_eval_batch(dev_->get(), params, B, ld, out, status, steps) != SEPAIHRD_OK)
        throw SimulationException("SEPAIHRDObjectiveFunction::calculateBatch", std::string("device evaluation failed: ") + sepaihrd_last_error());
}

void SEPAIHRDObjectiveFunction::calculateBatch(const double* params, int64_t B, int64_t ld, double* out, uint32_t* status, int32_t* steps) const {
    evaluateRows(params, B, ld, out, status, steps);
}

void SEPAIHRDObjectiveFunction::calculateBatch(const double* params, int64_t B, int64_t ld, double* out) const {
    evaluateThroughCache(cache_, dev_->numParams(), params, B, ld, out,
                         [this](const double* rows, int64_t M, int64_t rld, double* vals, uint32_t* st) { evaluateRows(rows, M, rld, vals, st, nullptr); });
}

// calculate() row by row as far as the cache is concerned (ObjectiveFunction.cpp:63-77): probe, remember the misses, evaluate
// them as ONE batch, store what calculate() would have stored (.cpp:227-234).
void evaluateThroughCache(ISimulationCache& cache_, std::ptrdiff_t P, const double* params, int64_t B, int64_t ld, double* out,
                          const std::function<void(const double*, int64_t, int64_t, double*, uint32_t*)>& evaluate) {
    if (B <= 0) return;
    auto* fast = dynamic_cast<SimulationCache*>(&cache_);
    const bool uncached = dynamic_cast<NullSimulationCache*>(&cache_) != nullptr || (fast != nullptr && static_cast<size_t>(B) > fast->capacity());
    if (uncached) {
        std::vector<uint32_t> st(static_cast<size_t>(B));
        return evaluate(params, B, ld, out, st.data());
    }
    std::vector<size_t> fkey(fast ? static_cast<size_t>(B) : 0);
    std::vector<std::string> skey(fast ? 0 : static_cast<size_t>(B));
    std::vector<int64_t> miss, first_of(static_cast<size_t>(B), -1);
    std::map<std::string, int64_t> seen_s;
    std::map<size_t, int64_t> seen_f;
    for (int64_t i = 0; i < B; ++i) {
        const double* row = params + i * ld;
        double v = 0.0;
        bool hit;
        if (fast) { fkey[static_cast<size_t>(i)] = fast->computeHash(row, P); hit = fast->getLikelihood(fkey[static_cast<size_t>(i)], v); }
        else { skey[static_cast<size_t>(i)] = cache_.createCacheKey(VectorXd::FromPointer(row, P)); hit = cache_.getLikelihood(skey[static_cast<size_t>(i)], v); }
        if (hit) { out[i] = v; first_of[static_cast<size_t>(i)] = i; continue; }
        const auto ins = fast ? seen_f.emplace(fkey[static_cast<size_t>(i)], i).first->second : seen_s.emplace(skey[static_cast<size_t>(i)], i).first->second;
        if (ins != i) first_of[static_cast<size_t>(i)] = ins;      // the same key earlier in this batch: evaluated once
        else miss.push_back(i);
    }
    if (!miss.empty()) {
        const int64_t M = static_cast<int64_t>(miss.size());
        std::vector<double> rows, vals(static_cast<size_t>(M));
        std::vector<uint32_t> st(static_cast<size_t>(M));
        const double* src = params;
        int64_t src_ld = ld;
        if (M != B) {                                              // gather the missing rows
            rows.resize(static_cast<size_t>(M) * static_cast<size_t>(P));
            for (int64_t j = 0; j < M; ++j) std::copy(params + miss[static_cast<size_t>(j)] * ld, params + miss[static_cast<size_t>(j)] * ld + P, rows.begin() + j * P);
            src = rows.data(); src_ld = P;
        }
        evaluate(src, M, src_ld, vals.data(), st.data());
        for (int64_t j = 0; j < M; ++j) {
            const int64_t i = miss[static_cast<size_t>(j)];
            out[i] = vals[static_cast<size_t>(j)];
            // calculate() reaches its store only on the complete path; the early `return lowest()` exits (bad parameters, S
            // overflow, invalid simulation result) leave the cache untouched
            if ((st[static_cast<size_t>(j)] & ~static_cast<uint32_t>(SEPAIHRD_ST_NONFINITE)) == 0) {
                if (fast) fast->storeLikelihood(fkey[static_cast<size_t>(i)], out[i]);
                else cache_.storeLikelihood(skey[static_cast<size_t>(i)], out[i]);
            }
        }
    }
    for (int64_t i = 0; i < B; ++i) {
        const int64_t f = first_of[static_cast<size_t>(i)];
        if (f >= 0 && f != i) {                                    // the repeat would have been a hit in calculate(): count it as one
            double v = 0.0;
            const bool hit = fast ? fast->getLikelihood(fkey[static_cast<size_t>(i)], v) : cache_.getLikelihood(skey[static_cast<size_t>(i)], v);
            out[i] = hit ? v : out[f];
        }
    }
}

// ---- finite-difference gradient ---------------------------------------------------------------------------------------------
void ForwardDifferences::perturb(const VectorXd& params, double epsilon, std::vector<double>& rows, std::vector<double>& steps) {
    const std::ptrdiff_t P = params.size();
    rows.resize(static_cast<size_t>(P * P));
    steps.resize(static_cast<size_t>(P));
    for (std::ptrdiff_t i = 0; i < P; ++i) {
        const double param_scale = std::max(std::abs(params(i)), epsilon);      // .cpp:35-36
        steps[static_cast<size_t>(i)] = epsilon * param_scale;
        double* row = rows.data() + i * P;
        std::copy(params.data(), params.data() + P, row);
        row[i] += steps[static_cast<size_t>(i)];
    }
}

void ForwardDifferences::gradient(double f_center, const double* f_plus, const uint8_t* skip, const std::vector<double>& steps, VectorXd& grad) {
    const std::ptrdiff_t P = static_cast<std::ptrdiff_t>(steps.size());
    grad.resize(P);
    for (std::ptrdiff_t i = 0; i < P; ++i) {
        if ((skip && skip[i]) || !std::isfinite(f_plus[i])) grad(i) = 0.0;      // .cpp:50-53, 101-104, 163-167
        else grad(i) = (f_plus[i] - f_center) / steps[static_cast<size_t>(i)];
    }
}

double SEPAIHRDGradientObjectiveFunction::evaluate_with_gradient(const VectorXd& params, VectorXd& grad) const {
    const std::ptrdiff_t P = params.size();
    grad.resize(P);
    const double f_center = SEPAIHRDObjectiveFunction::calculate(params);
    if (!std::isfinite(f_center)) {                                             // .cpp:24-29
        grad.setZero();
        return f_center;
    }
    std::vector<double> rows, steps, f_plus(static_cast<size_t>(P));
    std::vector<uint32_t> status(static_cast<size_t>(P));
    ForwardDifferences::perturb(params, epsilon_, rows, steps);
    auto* spm = dynamic_cast<SEPAIHRDParameterManager*>(&parameterManager_);
    const ConstraintMode mode = spm ? spm->getConstraintMode() : ConstraintMode::OPTIMIZATION_CLAMP;
    if (mode != ConstraintMode::OPTIMIZATION_CLAMP) device().setConstraintMode(ConstraintMode::OPTIMIZATION_CLAMP);
    try {
        evaluateRows(rows.data(), P, P, f_plus.data(), status.data(), nullptr);
    } catch (...) {
        if (mode != ConstraintMode::OPTIMIZATION_CLAMP) device().setConstraintMode(mode);
        throw;
    }
    if (mode != ConstraintMode::OPTIMIZATION_CLAMP) device().setConstraintMode(mode);
    ++gradient_batches_;
    // a perturbed vector the parameter manager rejects (negative kappa) or whose initial state does not fit into the
    // population contributes a zero component (.cpp:50-53, 93-104); every other failure is the finite sentinel lowest()
    std::vector<uint8_t> skip(static_cast<size_t>(P));
    for (std::ptrdiff_t i = 0; i < P; ++i) skip[static_cast<size_t>(i)] = (status[static_cast<size_t>(i)] & (SEPAIHRD_ST_INVALID_PARAM | SEPAIHRD_ST_S_OVERFLOW)) != 0;
    ForwardDifferences::gradient(f_center, f_plus.data(), skip.data(), steps, grad);
    return f_center;
}

// ---- SimulationCache (src/sir_age_structured/caching/SimulationCache.cpp) ---------------------------------------------
namespace {
inline size_t mix_hash(size_t k) {                                 // the 64-bit finaliser the reference mixes each coordinate with (.cpp:11-19)
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}
}  // namespace

SimulationCache::SimulationCache(size_t max_size) : capacity_(max_size) {
    if (max_size == 0) throw std::invalid_argument("SimulationCache: max_size must be > 0.");
    keys_.assign(capacity_, 0); values_.assign(capacity_, 0.0); frequencies_.assign(capacity_, 0); timestamps_.assign(capacity_, 0);
    occupied_.assign(capacity_, 0);
}

size_t SimulationCache::computeHash(const double* p, std::ptrdiff_t count) const {      // .cpp:35-52
    size_t seed = 0;
    for (std::ptrdiff_t i = 0; i < count; ++i) {
        const long long quantized = static_cast<long long>(p[i] * 1e8 + 0.5);
        seed ^= mix_hash(static_cast<size_t>(quantized)) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
    }
    return seed;
}
size_t SimulationCache::computeHash(const VectorXd& parameters) const { return computeHash(parameters.data(), parameters.size()); }
std::string SimulationCache::createCacheKey(const VectorXd& parameters) const { return std::to_string(computeHash(parameters)); }

size_t SimulationCache::findIndex(size_t key) const {              // linear probing from key % capacity (.cpp:58-72)
    size_t idx = key % capacity_;
    const size_t start = idx;
    while (occupied_[idx]) {
        if (keys_[idx] == key) return idx;
        if (++idx == capacity_) idx = 0;
        if (idx == start) break;
    }
    return NOT_FOUND;
}

size_t SimulationCache::evict() {                                  // least frequently used, oldest among equals (.cpp:74-104)
    size_t victim = 0;
    uint32_t min_freq = std::numeric_limits<uint32_t>::max(), min_time = std::numeric_limits<uint32_t>::max();
    for (size_t i = 0; i < capacity_; ++i) {
        if (!occupied_[i]) continue;
        if (frequencies_[i] < min_freq || (frequencies_[i] == min_freq && timestamps_[i] < min_time)) {
            min_freq = frequencies_[i]; min_time = timestamps_[i]; victim = i;
        }
    }
    occupied_[victim] = 0;
    --count_;
    return victim;
}

bool SimulationCache::lookupLocked(size_t key, double& value) {
    const size_t idx = findIndex(key);
    if (idx == NOT_FOUND) return false;
    frequencies_[idx]++;
    timestamps_[idx] = ++current_tick_;
    value = values_[idx];
    return true;
}

void SimulationCache::storeLocked(size_t key, double value) {
    const size_t idx = findIndex(key);
    if (idx != NOT_FOUND) {                                        // update in place
        values_[idx] = value;
        frequencies_[idx]++;
        timestamps_[idx] = ++current_tick_;
        return;
    }
    if (count_ >= capacity_) evict();
    size_t at = key % capacity_;
    while (occupied_[at]) if (++at == capacity_) at = 0;
    keys_[at] = key; values_[at] = value; frequencies_[at] = 1; timestamps_[at] = ++current_tick_; occupied_[at] = 1;
    ++count_;
}

std::optional<double> SimulationCache::get(const VectorXd& parameters) {
    std::lock_guard<std::mutex> lock(mutex_);
    double v = 0.0;
    if (lookupLocked(computeHash(parameters), v)) return v;
    return std::nullopt;
}
void SimulationCache::set(const VectorXd& parameters, double result) {
    std::lock_guard<std::mutex> lock(mutex_);
    storeLocked(computeHash(parameters), result);
}
void SimulationCache::clear() {
    std::lock_guard<std::mutex> lock(mutex_);
    std::fill(occupied_.begin(), occupied_.end(), 0);
    std::fill(frequencies_.begin(), frequencies_.end(), 0);
    count_ = 0;
    current_tick_ = 0;
}
size_t SimulationCache::size() const {
    std::lock_guard<std::mutex> lock(mutex_);
    return count_;
}
bool SimulationCache::getLikelihood(size_t key, double& value) {
    get_calls_.fetch_add(1, std::memory_order_relaxed);
    std::lock_guard<std::mutex> lock(mutex_);
    if (!lookupLocked(key, value)) return false;
    get_hits_.fetch_add(1, std::memory_order_relaxed);
    return true;
}
void SimulationCache::storeLikelihood(size_t key, double value) {
    store_calls_.fetch_add(1, std::memory_order_relaxed);
    std::lock_guard<std::mutex> lock(mutex_);
    storeLocked(key, value);
}
bool SimulationCache::getLikelihood(const std::string& key, double& value) {            // a key that is not a number is a miss (.cpp:165-180)
    size_t k = 0;
    try { k = std::stoull(key); } catch (...) { get_calls_.fetch_add(1, std::memory_order_relaxed); return false; }
    return getLikelihood(k, value);
}
void SimulationCache::storeLikelihood(const std::string& key, double value) {
    size_t k = 0;
    try { k = std::stoull(key); } catch (...) { store_calls_.fetch_add(1, std::memory_order_relaxed); return; }
    storeLikelihood(k, value);
}

// ---- ResultAggregator ---------------------------------------------------------------------------------------------
PosteriorPredictiveData ResultAggregator::aggregatePosteriorPredictives(const std::vector<VectorXd>& samples, SEPAIHRDParameterManager& pm,
                                                                        int num_samples_for_ppc, const std::vector<double>& time_points,
                                                                        const VectorXd& initial_state, const CalibrationData& observed,
                                                                        std::shared_ptr<AgeSEPAIHRDModel> model, unsigned int random_seed) const {
    PosteriorPredictiveData ppd;
    for (double t : time_points) if (t >= 0.0) ppd.time_points.push_back(t);
    ppd.daily_hospitalizations.observed = observed.getNewHospitalizations();
    ppd.daily_icu_admissions.observed = observed.getNewICU();
    ppd.daily_deaths.observed = observed.getNewDeaths();
    if (ppd.time_points.empty() || samples.empty()) return ppd;                    // .cpp:197-200, 218-221
    if (!model) throw InvalidParameterException("ResultAggregator", "model template is null");
    std::vector<size_t> pick;
    if (num_samples_for_ppc > 0 && static_cast<size_t>(num_samples_for_ppc) < samples.size()) {
        std::mt19937 gen(random_seed != 0 ? random_seed : std::random_device{}());
        std::uniform_int_distribution<> distrib(0, static_cast<int>(samples.size()) - 1);
        for (int i = 0; i < num_samples_for_ppc; ++i) pick.push_back(static_cast<size_t>(distrib(gen)));
    } else {
        for (size_t i = 0; i < samples.size(); ++i) pick.push_back(i);
    }
    const int64_t P = static_cast<int64_t>(pm.getParameterCount());
    std::vector<double> rows(pick.size() * static_cast<size_t>(P));
    for (size_t i = 0; i < pick.size(); ++i) {
        if (samples[pick[i]].size() != P) throw InvalidParameterException("ResultAggregator", "Parameter vector size mismatch.");
        std::copy(samples[pick[i]].data(), samples[pick[i]].data() + P, rows.begin() + static_cast<std::ptrdiff_t>(i * static_cast<size_t>(P)));
    }
    std::vector<double> lo, hi;
    for (int i = 0; i < static_cast<int>(P); ++i) { lo.push_back(pm.getLowerBoundForParamIndex(i)); hi.push_back(pm.getUpperBoundForParamIndex(i)); }
    DeviceContext dev(*model, time_points, nullptr, nullptr, nullptr, initial_state, pm.slots(), lo, hi, pm.getConstraintMode(), 1e-6, 1e-6, 1.0);
    const double probs[5] = {0.025, 0.05, 0.5, 0.95, 0.975};                        // .cpp:233
    const int n = model->getNumAgeClasses();
    const std::ptrdiff_t T = static_cast<std::ptrdiff_t>(ppd.time_points.size());
    std::vector<double> q(static_cast<size_t>(6 * T * n * 5));
    if (sepaihrd_posterior_predictive(dev.get(), rows.data(), static_cast<int64_t>(pick.size()), P, initial_state.data(), 5, probs, q.data(),
                                      &ppd.samples_used) != SEPAIHRD_OK)
        throw SimulationException("ResultAggregator", std::string("device aggregation failed: ") + sepaihrd_last_error());
    PosteriorPredictiveData::IncidenceData* series[6] = {&ppd.daily_hospitalizations, &ppd.daily_icu_admissions, &ppd.daily_deaths,
                                                         &ppd.cumulative_hospitalizations, &ppd.cumulative_icu_admissions, &ppd.cumulative_deaths};
    for (int s = 0; s < 6; ++s) {
        MatrixXd* out[5] = {&series[s]->lower_95, &series[s]->lower_90, &series[s]->median, &series[s]->upper_90, &series[s]->upper_95};
        for (MatrixXd* m : out) m->resize(T, n);
        for (std::ptrdiff_t t = 0; t < T; ++t)
            for (int a = 0; a < n; ++a)
                for (int k = 0; k < 5; ++k) (*out[k])(t, a) = q[static_cast<size_t>(((s * T + t) * n + a) * 5 + k)];
    }
    return ppd;
}

}  // namespace epidemic
