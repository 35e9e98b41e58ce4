// benchmark_main.cpp -- `sepaihrd_objective_benchmark` over the B200 evaluator: the reference's own timing harness
// (src/model/sepaihrd_objective_benchmark_main.cpp:52-560) with the same command line, the same object graph and the
// same jitter recipe, written against the C++ host mirror.  Everything is C++ over the C ABI: the reference-style
// project tree is read by config_io, the objects are the mirrored classes, and every objective call is a device batch.
//
// What changes against the reference harness, on purpose:
//   * `--repeats N` and `--jitters N` are each ONE batch call of N parameter sets (calculateBatch) instead of N calls of
//     calculate(); the jittered sets are the reference's: base_i + sigma_i N(0,1) from std::mt19937(seed) +
//     std::normal_distribution, set by set, then applyConstraints (:452-460);
//   * the likelihood cache is OFF unless `--cache-size N` is given (the reference defaults to a cache of 10000 entries and
//     takes `--no-cache`; quirk Q6: a hit may return another vector's value, so measurements and parity runs go without).
//     With a cache, calculate() and every batch of at most N sets consult it like the reference's calculate() does; the cache
//     statistics line of the reference is printed after each mode;
//   * `--project-root PATH` names the tree (the reference finds it by walking up from the working directory);
//   * `--mode pso` and `--chains N` exist in addition (batched callers), and `--json` prints one machine-readable line;
//   * `--mode threads` is the UNCHANGED reference calling pattern: an OpenMP loop (`--threads N` of them, more than cores is
//     the point) in which every thread calls calculate() for one jittered vector at a time, as ParticleSwarmOptimizer.cpp:368-424
//     does.  The C ABI merges the calls that arrive while a launch is in flight into the next launch.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <omp.h>
#include <random>
#include <string>
#include <vector>

#include "config_io.hpp"
#include "host_capi.h"
#include "optimizers.hpp"

using namespace epidemic;

namespace {

struct Args {
    int repeats = 200, jitters = 200, seed = 1, threads = 0;
    std::string startDate = "2020-03-01", endDate = "2020-12-31", constraintMode = "opt", mode = "micro", root = ".";
    std::string hillSettingsPath, mcmcSettingsPath, psoSettingsPath;
    bool useFileIters = false, json = false;
    int hillIters = 200, mcmcIters = 2000, psoIters = 30, chains = 1, swarm = 4096;
    bool psoAsConfigured = false;
    size_t cacheSize = 0;
};

void usage(const char* prog) {
    std::cout << "Usage: " << prog << " --project-root PATH [--repeats N] [--jitters N] [--seed N] [--threads N]\n"
              << "       [--start YYYY-MM-DD] [--end YYYY-MM-DD] [--constraints opt|mcmc] [--no-cache] [--cache-size N]\n"
              << "       [--mode micro|hill|mcmc|hillmcmc|pso|threads|all] [--hill-settings PATH] [--mcmc-settings PATH] [--pso-settings PATH]\n"
              << "       [--hill-iters N] [--mcmc-iters N] [--pso-iters N] [--chains N] [--swarm N] [--pso-as-configured] [--use-file-iters] [--json]\n";
}

Args parse(int argc, char** argv) {
    Args a;
    for (int i = 1; i < argc; ++i) {
        const std::string f = argv[i];
        auto value = [&]() -> std::string {
            if (i + 1 >= argc) throw std::runtime_error("Missing value after " + f);
            return argv[++i];
        };
        if (f == "--help" || f == "-h") { usage(argv[0]); std::exit(0); }
        else if (f == "--no-cache") a.cacheSize = 0;
        else if (f == "--cache-size") a.cacheSize = static_cast<size_t>(std::stoull(value()));
        else if (f == "--repeats") a.repeats = std::stoi(value());
        else if (f == "--jitters") a.jitters = std::stoi(value());
        else if (f == "--seed") a.seed = std::stoi(value());
        else if (f == "--threads") a.threads = std::stoi(value());
        else if (f == "--start") a.startDate = value();
        else if (f == "--end") a.endDate = value();
        else if (f == "--constraints") a.constraintMode = value();
        else if (f == "--mode") a.mode = value();
        else if (f == "--project-root") a.root = value();
        else if (f == "--hill-settings") a.hillSettingsPath = value();
        else if (f == "--mcmc-settings") a.mcmcSettingsPath = value();
        else if (f == "--pso-settings") a.psoSettingsPath = value();
        else if (f == "--use-file-iters") a.useFileIters = true;
        else if (f == "--hill-iters") a.hillIters = std::stoi(value());
        else if (f == "--mcmc-iters") a.mcmcIters = std::stoi(value());
        else if (f == "--pso-iters") a.psoIters = std::stoi(value());
        else if (f == "--pso-as-configured") a.psoAsConfigured = true;
        else if (f == "--chains") a.chains = std::stoi(value());
        else if (f == "--swarm") a.swarm = std::stoi(value());
        else if (f == "--json") a.json = true;
        else throw std::runtime_error("Unknown argument: " + f);
    }
    if (a.repeats < 0 || a.jitters < 0) throw std::runtime_error("repeats/jitters must be non-negative");
    if (a.constraintMode != "opt" && a.constraintMode != "mcmc") throw std::runtime_error("--constraints must be 'opt' or 'mcmc'");
    static const char* modes[] = {"micro", "hill", "mcmc", "hillmcmc", "pso", "threads", "all"};
    bool ok = false;
    for (const char* m : modes) ok = ok || a.mode == m;
    if (!ok) throw std::runtime_error("--mode must be one of: micro, hill, mcmc, hillmcmc, pso, threads, all");
    const std::string cfg = a.root + "/data/configuration/";
    if (a.hillSettingsPath.empty()) a.hillSettingsPath = cfg + "hill_climbing_settings.txt";
    if (a.mcmcSettingsPath.empty()) a.mcmcSettingsPath = cfg + "mcmc_settings.txt";
    if (a.psoSettingsPath.empty()) a.psoSettingsPath = cfg + "pso_settings.txt";
    return a;
}

using Clock = std::chrono::steady_clock;
double ms_since(Clock::time_point t0) { return std::chrono::duration<double, std::milli>(Clock::now() - t0).count(); }

std::map<std::string, double> settings_or_empty(const std::string& path, std::map<std::string, double> (*reader)(const std::string&)) {
    std::ifstream probe(path);
    if (!probe.is_open()) return {};
    return reader(path);
}

// page-locked buffer of doubles (sepaihrd_alloc_pinned)
class PinnedDoubles {
public:
    explicit PinnedDoubles(size_t count) {
        void* p = nullptr;
        if (sepaihrd_alloc_pinned(count * sizeof(double), &p) != SEPAIHRD_OK) throw std::runtime_error(std::string("sepaihrd_alloc_pinned: ") + sepaihrd_last_error());
        p_ = static_cast<double*>(p);
    }
    ~PinnedDoubles() { sepaihrd_free_pinned(p_); }
    PinnedDoubles(const PinnedDoubles&) = delete;
    PinnedDoubles& operator=(const PinnedDoubles&) = delete;
    double* data() { return p_; }
private:
    double* p_ = nullptr;
};

// counts what goes through it, like the reference's CountingObjective
class CountingObjective : public IObjectiveFunction {
public:
    explicit CountingObjective(SEPAIHRDObjectiveFunction& inner) : inner_(inner) {}
    double calculate(const VectorXd& p) const override { ++calls_; return inner_.calculate(p); }
    void calculateBatch(const double* params, int64_t B, int64_t ld, double* out) const override { calls_ += B; inner_.calculateBatch(params, B, ld, out); }
    const std::vector<std::string>& getParameterNames() const override { return inner_.getParameterNames(); }
    long long calls() const { return calls_; }
    void resetCalls() { calls_ = 0; }
private:
    SEPAIHRDObjectiveFunction& inner_;
    mutable long long calls_ = 0;
};

}  // namespace

int main(int argc, char** argv) {
    try {
        const Args args = parse(argc, argv);
        if (args.threads > 0) sepaihrd_host_set_threads(args.threads);
        const ReferenceProject prj = loadReferenceProject(args.root, args.startDate, args.endDate);
        SEPAIHRDParameterManager pm(prj.model, prj.params_to_calibrate, prj.proposal_sigmas, prj.param_bounds);
        pm.setConstraintMode(args.constraintMode == "mcmc" ? ConstraintMode::MCMC_REFLECT : ConstraintMode::OPTIMIZATION_CLAMP);
        std::unique_ptr<ISimulationCache> cache_owner;
        SimulationCache* sim_cache = nullptr;
        if (args.cacheSize > 0) { auto c = std::make_unique<SimulationCache>(args.cacheSize); sim_cache = c.get(); cache_owner = std::move(c); }
        else cache_owner = std::make_unique<NullSimulationCache>();
        ISimulationCache& cache = *cache_owner;
        auto print_cache_stats = [&]() {                                       // printCacheStats, reference harness :260-272
            if (!sim_cache) return;
            const size_t calls = sim_cache->getLikelihoodCalls(), hits = sim_cache->getLikelihoodHits();
            std::printf("Cache stats: size=%zu, get_calls=%zu, hits=%zu (%.2f%%), store_calls=%zu\n", sim_cache->size(), calls, hits,
                        calls ? 100.0 * static_cast<double>(hits) / static_cast<double>(calls) : 0.0, sim_cache->storeLikelihoodCalls());
            sim_cache->resetLikelihoodStats();
        };
        const CalibrationData data = prj.data->toCalibrationData(prj.data_initial_state);
        SEPAIHRDObjectiveFunction objective(prj.model, pm, cache, data, prj.time_points, prj.data_initial_state,
                                            std::make_shared<Dopri5SolverStrategy>(), prj.abs_error, prj.rel_error);
        CountingObjective counting(objective);
        const VectorXd base = pm.getCurrentParameters();
        const auto P = base.size();

        std::cout << "SEPAIHRD objective benchmark (B200 evaluator)\n"
                  << "Parameters: " << P << ", output points: " << prj.time_points.size() << ", data days: " << prj.data->getNumDataPoints() << "\n"
                  << "Cache enabled: " << (sim_cache ? "yes, capacity " + std::to_string(args.cacheSize) : std::string("no (default here; --cache-size N enables it, quirk Q6)")) << "\n"
                  << "Mode: " << args.mode << "\nConstraint mode: " << args.constraintMode << "\n"
                  << "Data window: " << args.startDate << " .. " << args.endDate << "\n";
        std::map<std::string, double> report;

        auto run_micro = [&]() {
            counting.resetCalls();
            auto t0 = Clock::now();
            const double warmup_val = counting.calculate(base);
            const double warmup_ms = ms_since(t0);

            // parameter rows and results in page-locked memory: the host-buffer call then overlaps its copies with the kernel
            const size_t most = static_cast<size_t>(std::max(std::max(args.repeats, args.jitters), 1));
            PinnedDoubles rows(most * static_cast<size_t>(P)), out(most);
            auto batch = [&](int n, double& sum) {
                const auto t = Clock::now();
                if (n > 0) counting.calculateBatch(rows.data(), n, P, out.data());
                const double took = ms_since(t);
                sum = 0.0;
                for (int i = 0; i < n; ++i) sum += out.data()[i];
                return took;
            };
            for (int i = 0; i < args.repeats; ++i) std::copy(base.data(), base.data() + P, rows.data() + static_cast<size_t>(i) * P);
            double repeat_sum = 0.0;
            const double repeats_ms = batch(args.repeats, repeat_sum);

            // the reference's recipe, set by set (draw order included): one generator, drawn sequentially; the constraints
            // are applied afterwards, row-parallel
            std::mt19937 rng(static_cast<unsigned>(args.seed));
            std::normal_distribution<double> normal(0.0, 1.0);
            t0 = Clock::now();
            std::vector<double> sigma(static_cast<size_t>(P));
            for (std::ptrdiff_t i = 0; i < P; ++i) sigma[static_cast<size_t>(i)] = pm.getSigmaForParamIndex(static_cast<int>(i));
            for (int k = 0; k < args.jitters; ++k) {
                double* row = rows.data() + static_cast<size_t>(k) * P;
                for (std::ptrdiff_t i = 0; i < P; ++i) row[i] = base[i] + sigma[static_cast<size_t>(i)] * normal(rng);
            }
#pragma omp parallel for schedule(static)
            for (int k = 0; k < args.jitters; ++k) {
                double* row = rows.data() + static_cast<size_t>(k) * P;
                const VectorXd c = pm.applyConstraints(VectorXd::FromPointer(row, P));
                std::copy(c.data(), c.data() + P, row);
            }
            const double gen_ms = ms_since(t0);
            double jitter_sum = 0.0;
            const double jitters_ms = batch(args.jitters, jitter_sum);

            std::printf("\n--- Micro ---\nWarmup: 1 eval => %.3f ms (value=%.12e)\n", warmup_ms, warmup_val);
            if (args.repeats > 0)
                std::printf("Repeat: %d evals => %.3f ms (avg %.4f us/eval, %.4e evals/s)\n", args.repeats, repeats_ms, repeats_ms * 1000.0 / args.repeats,
                            args.repeats / repeats_ms * 1e3);
            if (args.jitters > 0)
                std::printf("Jitter: %d evals => %.3f ms (avg %.4f us/eval, %.4e evals/s; host generation of the sets %.3f ms; sum logL %.12e)\n",
                            args.jitters, jitters_ms, jitters_ms * 1000.0 / args.jitters, args.jitters / jitters_ms * 1e3, gen_ms, jitter_sum);
            std::printf("Objective calls: %lld\n", counting.calls());
            print_cache_stats();
            report["warmup_value"] = warmup_val; report["repeat_ms"] = repeats_ms; report["jitter_ms"] = jitters_ms;
            report["jitter_sum"] = jitter_sum; report["repeat_sum"] = repeat_sum;
            report["jitter_evals_per_s"] = args.jitters > 0 ? args.jitters / jitters_ms * 1e3 : 0.0;
        };

        auto run_hill = [&]() {
            counting.resetCalls();
            std::map<std::string, double> st = settings_or_empty(args.hillSettingsPath, readHillClimbingSettings);
            if (!args.useFileIters) st["iterations"] = args.hillIters;
            st["seed"] = args.seed;
            pm.setConstraintMode(ConstraintMode::OPTIMIZATION_CLAMP);      // calibration behaviour: clamp in phase 1
            HillClimbingOptimizer hill;
            hill.configure(st);
            const auto t0 = Clock::now();
            const OptimizationResult res = hill.optimize(base, counting, pm);
            const double took = ms_since(t0);
            std::printf("\n--- Hill Climbing ---\nTime: %.3f ms\nObjective calls: %lld\nBest logL: %.12e\n", took, counting.calls(), res.bestObjectiveValue);
            print_cache_stats();
            report["hill_ms"] = took; report["hill_best"] = res.bestObjectiveValue; report["hill_calls"] = static_cast<double>(counting.calls());
            return res.bestParameters;
        };

        auto run_mcmc = [&](const VectorXd& start) {
            counting.resetCalls();
            std::map<std::string, double> st = settings_or_empty(args.mcmcSettingsPath, readMetropolisHastingsSettings);
            if (!args.useFileIters) st["mcmc_iterations"] = args.mcmcIters;
            st["store_samples"] = 0.0;
            st["n_chains"] = args.chains;
            st["seed"] = args.seed;
            st["write_checkpoints"] = 0.0; st["write_trace"] = 0.0;      // the timing harness runs with file output disabled, like the reference's (:484-542)
            MetropolisHastingsSampler mcmc;
            mcmc.configure(st);
            const auto t0 = Clock::now();
            const OptimizationResult res = mcmc.optimize(start, counting, pm);       // flips the manager to MCMC_REFLECT like the reference
            const double took = ms_since(t0);
            std::printf("\n--- MCMC (SA-MH, %d chains) ---\nTime: %.3f ms\nObjective calls: %lld\nBest logL: %.12e\n", args.chains, took, counting.calls(),
                        res.bestObjectiveValue);
            print_cache_stats();
            report["mcmc_ms"] = took; report["mcmc_best"] = res.bestObjectiveValue; report["mcmc_calls"] = static_cast<double>(counting.calls());
        };

        auto run_pso = [&]() {
            std::map<std::string, double> st = settings_or_empty(args.psoSettingsPath, readParticleSwarmSettings);
            if (!args.useFileIters) { st["iterations"] = args.psoIters; st["swarm_size"] = args.swarm; }
            st["seed"] = args.seed;
            // default: the STANDARD / GLOBAL_BEST swarm of BASELINE.json configs[3] (device-resident); --pso-as-configured keeps the
            // variant / topology / opposition / adaptation switches of the settings file (whole-swarm engine, one batch per evaluation)
            if (!args.psoAsConfigured)
                for (const char* k : {"variant", "topology", "use_opposition_learning", "use_adaptive_parameters"}) st[k] = 0.0;
            pm.setConstraintMode(ConstraintMode::OPTIMIZATION_CLAMP);
            ParticleSwarmOptimization pso;
            pso.configure(st);
            const auto t0 = Clock::now();
            const OptimizationResult res = pso.optimize(base, objective, pm);          // the device-resident swarm needs the device objective itself
            const double took = ms_since(t0);
            const double evals = static_cast<double>(pso.evaluations());
            std::printf("\n--- PSO (%s) ---\nTime: %.3f ms\nObjective calls: %.0f (%.4e evals/s)\nBest logL: %.12e\n", pso.onDevice() ? "device-resident swarm" : "host-resident swarm, batched evaluations", took, evals,
                        evals / took * 1e3, res.bestObjectiveValue);
            report["pso_ms"] = took; report["pso_best"] = res.bestObjectiveValue; report["pso_calls"] = evals;
        };

        auto run_threads = [&]() {
            const int n = std::max(args.jitters, 1);
            std::vector<double> rows(static_cast<size_t>(n) * static_cast<size_t>(P)), out(static_cast<size_t>(n));
            std::mt19937 rng(static_cast<unsigned>(args.seed));
            std::normal_distribution<double> normal(0.0, 1.0);
            for (int k = 0; k < n; ++k) {
                VectorXd v(P);
                for (std::ptrdiff_t i = 0; i < P; ++i) v(i) = base[i] + pm.getSigmaForParamIndex(static_cast<int>(i)) * normal(rng);
                const VectorXd c = pm.applyConstraints(v);
                std::copy(c.data(), c.data() + P, rows.begin() + static_cast<std::ptrdiff_t>(k) * P);
            }
            (void)objective.calculate(base);                                   // warm-up
            int64_t ml0 = 0, mr0 = 0, l0 = 0, s0 = 0;
            sepaihrd_get_merge_counters(objective.device().get(), &ml0, &mr0);
            sepaihrd_get_counters(objective.device().get(), &l0, &s0);
            const int nthreads = args.threads > 0 ? args.threads : 0;
            const auto t0 = Clock::now();
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : omp_get_max_threads())
            for (int k = 0; k < n; ++k) out[static_cast<size_t>(k)] = objective.calculate(VectorXd::FromPointer(rows.data() + static_cast<std::ptrdiff_t>(k) * P, P));
            const double took = ms_since(t0);
            int64_t ml1 = 0, mr1 = 0, l1 = 0, s1 = 0;
            sepaihrd_get_merge_counters(objective.device().get(), &ml1, &mr1);
            sepaihrd_get_counters(objective.device().get(), &l1, &s1);
            double sum = 0.0;
            for (double v : out) sum += v;
            std::printf("\n--- calculate() from an OpenMP loop (%d threads) ---\n%d evals => %.3f ms (%.4e evals/s; sum logL %.12e)\n"
                        "Kernel launches: %lld (%.1f calls per launch; %lld launches served more than one call)\n",
                        nthreads > 0 ? nthreads : omp_get_max_threads(), n, took, n / took * 1e3, sum, static_cast<long long>(l1 - l0),
                        static_cast<double>(n) / std::max<int64_t>(l1 - l0, 1), static_cast<long long>(ml1 - ml0));
            report["threads_ms"] = took; report["threads_evals_per_s"] = n / took * 1e3; report["threads_launches"] = static_cast<double>(l1 - l0);
            report["threads_sum"] = sum;
        };

        if (args.mode == "micro") run_micro();
        else if (args.mode == "threads") run_threads();
        else if (args.mode == "hill") (void)run_hill();
        else if (args.mode == "mcmc") run_mcmc(base);
        else if (args.mode == "pso") run_pso();
        else if (args.mode == "hillmcmc") { const VectorXd best = run_hill(); run_mcmc(best); }
        else { run_micro(); const VectorXd best = run_hill(); run_mcmc(best); run_pso(); }

        if (args.json) {
            std::printf("JSON {");
            bool first = true;
            for (const auto& [k, v] : report) { std::printf("%s\"%s\": %.17g", first ? "" : ", ", k.c_str(), v); first = false; }
            std::printf("}\n");
        }
        return 0;
    } catch (const std::exception& e) {
        std::cerr << "sepaihrd_objective_benchmark: " << e.what() << "\n";
        return 1;
    }
}
