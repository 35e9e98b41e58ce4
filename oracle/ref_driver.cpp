// ref_driver.cpp -- drives the UNMODIFIED reference (adjo0043/Mathematical-Modeling-Of-Infectious-Diseases-V1) through its own
// public API so that the oracle can be pinned against it:  oracle/build_ref.sh compiles this file together with the reference's
// sources, where they lie, against a Boost + Eigen given by BOOST_ROOT / EIGEN_ROOT.  TEST INFRASTRUCTURE ONLY (never shipped,
// never on the product path).  Neither Boost nor Eigen exists in the build image of this repository, so this file has not been
// compiled there; it only uses the reference's declared interfaces (cited below).
//
//   ref_driver <reference_root> <params.bin> <out.bin> [clamp|reflect]
//     params.bin : int64 B, int64 P, then B*P doubles (rows = parameter vectors in params_to_calibrate.txt order)
//     out.bin    : B doubles logL | B int64 accepted | B int64 rejected | B doubles logL from the counting pass
//
// Pass 1 evaluates SEPAIHRDObjectiveFunction::calculate (src/model/objectives/SEPAIHRDObjectiveFunction.cpp:62-235) with the
// reference's own Dopri5SolverStrategy (src/sir_age_structured/solvers/Dopri5SolverStrategy.cpp:9-43) and a null cache
// (pattern: src/model/sepaihrd_objective_benchmark_main.cpp:229-238).  Pass 2 repeats it with a solver strategy that makes the
// same three Boost calls but counts right-hand-side evaluations and takes integrate_times' return value (accepted steps):
// attempts = (rhs_calls - 1) / 6, rejected = attempts - accepted.  Both passes must return identical log-likelihoods.
// The object graph is assembled exactly like sepaihrd_objective_benchmark_main.cpp:287-407.
#include <Eigen/Dense>

#include <cstdint>
#include <cstdio>
#include <filesystem>
#include <iostream>
#include <map>
#include <memory>
#include <optional>
#include <string>
#include <vector>

#include <boost/numeric/odeint/integrate/integrate_times.hpp>
#include <boost/numeric/odeint/stepper/generation.hpp>
#include <boost/numeric/odeint/stepper/runge_kutta_dopri5.hpp>

#include "exceptions/Exceptions.hpp"
#include "model/AgeSEPAIHRDModel.hpp"
#include "model/ModelConstants.hpp"
#include "model/PieceWiseConstantNPIStrategy.hpp"
#include "model/objectives/SEPAIHRDObjectiveFunction.hpp"
#include "model/parameters/SEPAIHRDParameterManager.hpp"
#include "sir_age_structured/interfaces/IOdeSolverStrategy.hpp"
#include "sir_age_structured/interfaces/ISimulationCache.hpp"
#include "sir_age_structured/solvers/Dopri5SolverStrategy.hpp"
#include "utils/FileUtils.hpp"
#include "utils/GetCalibrationData.hpp"
#include "utils/Logger.hpp"
#include "utils/ReadCalibrationConfiguration.hpp"
#include "utils/ReadContactMatrix.hpp"

using namespace epidemic;
using Eigen::MatrixXd;
using Eigen::VectorXd;

namespace {

class NullCache final : public ISimulationCache {
public:
    std::optional<double> get(const VectorXd&) override { return std::nullopt; }
    void set(const VectorXd&, double) override {}
    void clear() override {}
    size_t size() const override { return 0; }
    std::string createCacheKey(const VectorXd&) const override { return std::string(); }
    bool getLikelihood(const std::string&, double&) override { return false; }
    void storeLikelihood(const std::string&, double) override {}
};

// The same Boost calls as Dopri5SolverStrategy::integrate, plus counters.
class CountingDopri5 final : public IOdeSolverStrategy {
public:
    mutable long long rhs_calls = 0, accepted = 0;
    void integrate(const std::function<void(const state_type&, state_type&, double)>& system, state_type& initial_state,
                   const std::vector<double>& times, double dt_hint, std::function<void(const state_type&, double)> observer,
                   double abs_error, double rel_error) const override {
        using namespace boost::numeric::odeint;
        auto counted = [&](const state_type& x, state_type& dxdt, double t) { ++rhs_calls; system(x, dxdt, t); };
        auto stepper = make_controlled<runge_kutta_dopri5<state_type>>(abs_error, rel_error);
        accepted += static_cast<long long>(integrate_times(stepper, counted, initial_state, times.begin(), times.end(), dt_hint, observer));
    }
};

std::shared_ptr<PiecewiseConstantNpiStrategy> make_npi(const SEPAIHRDParameters& params, const std::vector<std::string>& kappa_names,
                                                       const std::map<std::string, std::pair<double, double>>& bounds) {
    std::map<std::string, std::pair<double, double>> npi_bounds;
    for (const auto& name : kappa_names) {
        if (name == "kappa_1") continue;
        auto it = bounds.find(name);
        if (it != bounds.end()) npi_bounds[name] = it->second;
    }
    std::vector<double> ends(params.kappa_end_times.begin() + 1, params.kappa_end_times.end());
    std::vector<double> vals(params.kappa_values.begin() + 1, params.kappa_values.end());
    std::vector<std::string> names(kappa_names.begin() + 1, kappa_names.end());
    return std::make_shared<PiecewiseConstantNpiStrategy>(ends, vals, npi_bounds, params.kappa_values.at(0), params.kappa_end_times.at(0), true, names);
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 4) { std::fprintf(stderr, "usage: ref_driver <reference_root> <params.bin> <out.bin> [clamp|reflect]\n"); return 2; }
    const std::string root = argv[1], in_path = argv[2], out_path = argv[3];
    const bool reflect = argc > 4 && std::string(argv[4]) == "reflect";
    Logger::getInstance().setLogLevel(LogLevel::ERROR);
    try {
        std::filesystem::current_path(root);                              // FileUtils::getProjectRoot() starts from the cwd
        const int n = constants::DEFAULT_NUM_AGE_CLASSES;
        const std::string pr = FileUtils::getProjectRoot();
        CalibrationData data(FileUtils::joinPaths(pr, "data/processed/processed_data.csv"), "2020-03-01", "2020-12-31");
        MatrixXd C = readMatrixFromCSV(FileUtils::joinPaths(pr, "data/contacts.csv"), n, n);
        SEPAIHRDParameters params = readSEPAIHRDParameters(FileUtils::joinPaths(pr, "data/configuration/initial_guess.txt"), n);
        params.N = data.getPopulationByAgeGroup();
        params.M_baseline = C;
        if (!params.validate()) throw std::runtime_error("SEPAIHRDParameters validation failed");
        std::vector<std::string> kappa_names;
        for (size_t i = 0; i < params.kappa_values.size(); ++i) kappa_names.push_back("kappa_" + std::to_string(i + 1));
        auto bounds = readParamBounds(FileUtils::joinPaths(pr, "data/configuration/param_bounds.txt"));
        auto sigmas = readProposalSigmas(FileUtils::joinPaths(pr, "data/configuration/proposal_sigmas.txt"));
        auto names = readParamsToCalibrate(FileUtils::joinPaths(pr, "data/configuration/params_to_calibrate.txt"));

        const double runup_days = params.runup_days;
        const int num_days = data.getNumDataPoints();
        std::vector<double> time_points;
        for (int t = -static_cast<int>(runup_days); t < num_days; ++t) time_points.push_back(static_cast<double>(t));

        // the objective rebuilds the initial state per evaluation (ObjectiveFunction.cpp:124-163); this one only seeds the ctor
        VectorXd initial_state = data.getInitialSEPAIHRDState(params.sigma, params.gamma_p, params.gamma_A, params.gamma_I, params.p, params.h);

        auto npi = make_npi(params, kappa_names, bounds);
        auto model = std::make_shared<AgeSEPAIHRDModel>(params, npi);
        SEPAIHRDParameterManager pm(model, names, sigmas, bounds);
        pm.setConstraintMode(reflect ? ConstraintMode::MCMC_REFLECT : ConstraintMode::OPTIMIZATION_CLAMP);
        NullCache cache;

        std::FILE* fi = std::fopen(in_path.c_str(), "rb");
        if (!fi) throw std::runtime_error("cannot open " + in_path);
        int64_t B = 0, P = 0;
        if (std::fread(&B, 8, 1, fi) != 1 || std::fread(&P, 8, 1, fi) != 1) throw std::runtime_error("short header");
        if (P != static_cast<int64_t>(names.size())) throw std::runtime_error("P does not match params_to_calibrate.txt");
        std::vector<double> rows(static_cast<size_t>(B * P));
        if (std::fread(rows.data(), 8, rows.size(), fi) != rows.size()) throw std::runtime_error("short parameter block");
        std::fclose(fi);

        std::vector<double> ll(B), ll2(B);
        std::vector<int64_t> acc(B), rej(B);
        {
            auto solver = std::make_shared<Dopri5SolverStrategy>();
            SEPAIHRDObjectiveFunction objective(model, pm, cache, data, time_points, initial_state, solver, 1.0e-6, 1.0e-6);
            for (int64_t b = 0; b < B; ++b) ll[b] = objective.calculate(Eigen::Map<const VectorXd>(rows.data() + b * P, P));
        }
        {
            auto solver = std::make_shared<CountingDopri5>();
            SEPAIHRDObjectiveFunction objective(model, pm, cache, data, time_points, initial_state, solver, 1.0e-6, 1.0e-6);
            for (int64_t b = 0; b < B; ++b) {
                solver->rhs_calls = 0; solver->accepted = 0;
                ll2[b] = objective.calculate(Eigen::Map<const VectorXd>(rows.data() + b * P, P));
                const long long attempts = solver->rhs_calls > 0 ? (solver->rhs_calls - 1) / 6 : 0;
                acc[b] = solver->accepted; rej[b] = attempts - solver->accepted;
            }
        }
        std::FILE* fo = std::fopen(out_path.c_str(), "wb");
        if (!fo) throw std::runtime_error("cannot open " + out_path);
        std::fwrite(ll.data(), 8, B, fo); std::fwrite(acc.data(), 8, B, fo); std::fwrite(rej.data(), 8, B, fo); std::fwrite(ll2.data(), 8, B, fo);
        std::fclose(fo);
        std::printf("ref_driver: %lld sets, logL[0] = %.12e, accepted/rejected[0] = %lld/%lld\n", (long long)B, ll[0], (long long)acc[0], (long long)rej[0]);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ref_driver: %s\n", e.what());
        return 1;
    }
    return 0;
}
