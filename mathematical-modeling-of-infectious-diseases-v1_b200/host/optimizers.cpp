// optimizers.cpp -- batched calibrators over IObjectiveFunction::calculateBatch.  See optimizers.hpp.
#include "optimizers.hpp"
#include "../csrc/det_math.h"      // log / exp with the same bits on the host and on the device (see there)

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <iostream>

namespace epidemic {

namespace {

double setting(const std::map<std::string, double>& s, const std::string& key, double def) {
    auto it = s.find(key);
    return it != s.end() ? it->second : def;
}

std::vector<double> evaluate_rows(IObjectiveFunction& f, const std::vector<double>& rows, int64_t B, int64_t P) {
    std::vector<double> out(static_cast<size_t>(B));
    if (B > 0) f.calculateBatch(rows.data(), B, P, out.data());
    for (double& v : out) v = MetropolisHastingsSampler::safeValue(v);
    return out;
}

// std::normal_distribution<double>(0, 1) as libstdc++ draws it (Marsaglia's polar method over generate_canonical<double, 53>;
// the second value of a pair is kept for the next call), with det_math's logarithm in place of libm's and every product
// rounded on its own: csrc/sepaihrd_mh.cu draws the same numbers from the same generator state on the device.
struct PolarNormal {
    bool has_saved = false;
    double saved = 0.0;
    double operator()(std::mt19937& gen) {
        if (has_saved) { has_saved = false; return saved; }
        double x, y, r2;
        do {
            x = 2.0 * std::generate_canonical<double, 53>(gen) - 1.0;
            y = 2.0 * std::generate_canonical<double, 53>(gen) - 1.0;
            r2 = detm::opaque(x * x) + detm::opaque(y * y);
        } while (r2 > 1.0 || r2 == 0.0);
        const double mult = std::sqrt(-2.0 * detm::log(r2) / r2);
        saved = x * mult;
        has_saved = true;
        return y * mult;
    }
};

MatrixXd outer(const VectorXd& d) {
    const auto n = d.size();
    MatrixXd m(n, n);
    for (std::ptrdiff_t j = 0; j < n; ++j)
        for (std::ptrdiff_t i = 0; i < n; ++i) m(i, j) = d(i) * d(j);
    return m;
}

}  // namespace

// =====================================================================================================================
// Metropolis-Hastings (adaptive Metropolis, Haario et al.; Robbins-Monro global scale)
// =====================================================================================================================
MetropolisHastingsSampler::MetropolisHastingsSampler() = default;

void MetropolisHastingsSampler::configure(const std::map<std::string, double>& s) {
    iterations_ = static_cast<int>(setting(s, "mcmc_iterations", 10000.0));
    burn_in_ = static_cast<int>(setting(s, "burn_in", 1000.0));
    adaptation_period_ = std::max(1, static_cast<int>(setting(s, "adaptation_period", 100.0)));
    report_interval_ = std::max(1, static_cast<int>(setting(s, "report_interval", 100.0)));
    thinning_ = std::max(1, static_cast<int>(setting(s, "thinning", 1.0)));
    regularization_epsilon_ = setting(s, "regularization_epsilon", 1e-6);
    target_acceptance_rate_ = setting(s, "target_acceptance_rate", 0.234);
    adapt_scale_ = setting(s, "adapt_scale", 1.0) != 0.0;
    store_samples_ = setting(s, "store_samples", 1.0) != 0.0;
    write_checkpoints_ = setting(s, "write_checkpoints", 1.0) != 0.0;
    write_trace_ = setting(s, "write_trace", 1.0) != 0.0;
    // batched / repeatable extensions
    n_chains_ = std::max(1, static_cast<int>(setting(s, "n_chains", 1.0)));
    lookahead_ = std::max(0, static_cast<int>(setting(s, "lookahead", 0.0)));
    chain_offset_ = static_cast<long>(setting(s, "chain_offset", 0.0));
    has_seed_ = s.count("seed") != 0;
    seed_ = static_cast<unsigned>(setting(s, "seed", 0.0));
}

void MetropolisHastingsSampler::setInitialCovariance(const MatrixXd& cov) {
    hasInitialCovariance_ = cov.rows() > 0 && cov.rows() == cov.cols();
    if (hasInitialCovariance_) initialCovariance_ = cov;
}

void MetropolisHastingsSampler::begin(const VectorXd& initial, const double* initial_logpost, IParameterManager& pm) {
    n_params_ = static_cast<int>(initial.size());
    const auto P = static_cast<std::ptrdiff_t>(n_params_);
    // initial covariance (.cpp:216-240)
    if (hasInitialCovariance_ && initialCovariance_.rows() == P) {
        shared_cov_ = initialCovariance_;
    } else {
        shared_cov_ = MatrixXd::Identity(P, P);
        for (std::ptrdiff_t i = 0; i < P; ++i) {
            const double sg = pm.getSigmaForParamIndex(static_cast<int>(i));
            shared_cov_(i, i) = (sg > 0 ? sg * sg : 1e-6);
        }
        shared_cov_ *= (2.38 * 2.38) / static_cast<double>(n_params_);
    }
    shared_cov_ += regularization_epsilon_ * MatrixXd::Identity(P, P);
    if (!linalg::cholesky_lower(shared_cov_, shared_chol_)) shared_chol_ = MatrixXd::Identity(P, P) * 0.1;

    keep_history_ = iterations_ - 1 > burn_in_;      // the covariance adaptation reads the whole chain history
    shared_diagonal_ = true;                          // the start kernel built from the sigmas is diagonal: L z is then P products
    for (std::ptrdiff_t j = 0; j < P && shared_diagonal_; ++j)
        for (std::ptrdiff_t i = 0; i < P; ++i)
            if (i != j && shared_chol_(i, j) != 0.0) { shared_diagonal_ = false; break; }
    std::random_device rd;
    chains_.clear();
    chains_.resize(static_cast<size_t>(n_chains_));   // constructed in place (a chain carries a 2.5 KB generator)
    cur_x_.assign(static_cast<size_t>(n_chains_) * n_params_, 0.0);
    cur_lp_.assign(static_cast<size_t>(n_chains_), 0.0);
    prop_x_.assign(cur_x_.size(), 0.0);
    std::vector<unsigned> fallback_seeds;
    if (!has_seed_) { fallback_seeds.resize(static_cast<size_t>(n_chains_)); for (auto& v : fallback_seeds) v = rd(); }
#pragma omp parallel for schedule(static)
    for (int c = 0; c < n_chains_; ++c) {
        Chain& ch = chains_[static_cast<size_t>(c)];
        if (has_seed_) {
            std::seed_seq seq{seed_, static_cast<unsigned>(chain_offset_ + c)};
            ch.gen.seed(seq);
        } else {
            ch.gen.seed(fallback_seeds[static_cast<size_t>(c)]);
        }
        std::copy(initial.data(), initial.data() + n_params_, cur_x_.begin() + static_cast<std::ptrdiff_t>(c) * n_params_);
        cur_lp_[static_cast<size_t>(c)] = safeValue(initial_logpost[c]);
        ch.t = 1;
        ch.running_mean = initial;
        ch.best_x = initial;
        ch.best_lp = cur_lp_[static_cast<size_t>(c)];
        ch.hist.clear(); ch.hist_sum = VectorXd::Zero(static_cast<std::ptrdiff_t>(n_params_));
        if (keep_history_) pushHistory(ch, initial.data());
        if (store_samples_) { ch.samples.push_back(initial); ch.sample_lp.push_back(ch.best_lp); }
    }
    t_ = 1;
}

void MetropolisHastingsSampler::ownKernel(Chain& c) const {
    if (c.own_kernel) return;
    c.cov = shared_cov_;
    c.chol = shared_chol_;
    c.own_kernel = true;
}

// chain_history_.push_back (.cpp:369): the states in one contiguous buffer, and their running sum (the mean the full
// recomputation needs is this sum: the same additions in the same order as summing the history again)
void MetropolisHastingsSampler::pushHistory(Chain& c, const double* x) const {
    const auto P = static_cast<std::ptrdiff_t>(n_params_);
    c.hist.insert(c.hist.end(), x, x + P);
    double* s = c.hist_sum.data();
    for (std::ptrdiff_t i = 0; i < P; ++i) s[i] += x[i];
}

void MetropolisHastingsSampler::updateCovarianceRank1(Chain& c, int step) const {
    if (c.hist.empty()) return;
    const auto P = static_cast<std::ptrdiff_t>(n_params_);
    const double gamma = 10.0 / (step + 100.0);
    const VectorXd diff = VectorXd::FromPointer(c.hist.data() + c.hist.size() - static_cast<size_t>(P), P) - c.running_mean;
    c.running_mean += gamma * diff;
    // cov = (1 - gamma) cov + gamma diff diff^T, element by element in place (the same three products and one sum per element
    // as the matrix expression, without its four P x P temporaries)
    const double keep = 1.0 - gamma;
    double* cov = c.cov.data();
    const double* d = diff.data();
    for (std::ptrdiff_t j = 0; j < P; ++j) {
        const double dj = d[j];
        double* col = cov + j * P;
        for (std::ptrdiff_t i = 0; i < P; ++i) col[i] = keep * col[i] + gamma * (d[i] * dj);
    }
}

namespace {
// cv(i, j) += sum over the rows t of D of D(t, i) D(t, j) for the columns [j0, j1) of the lower triangle, in row order, ONE pass
// over D; four rows per pass over a column (added one after the other, so every element sees the additions of the plain loop).
// Compiled a second time for AVX2 (four elements per instruction; no fused multiply-add: the products are rounded on their own
// on every path) and picked at load time.
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
__attribute__((target_clones("avx2", "default")))
#endif
void accumulate_lower(const double* D, std::ptrdiff_t n, std::ptrdiff_t P, std::ptrdiff_t j0, std::ptrdiff_t j1, double* cv) {
    std::ptrdiff_t t = 0;
    for (; t + 4 <= n; t += 4) {
        const double* __restrict r0 = D + t * P;
        const double* __restrict r1 = r0 + P;
        const double* __restrict r2 = r1 + P;
        const double* __restrict r3 = r2 + P;
        for (std::ptrdiff_t j = j0; j < j1; ++j) {
            const double d0 = r0[j], d1 = r1[j], d2 = r2[j], d3 = r3[j];
            double* __restrict col = cv + j * P;
#pragma omp simd
            for (std::ptrdiff_t i = j; i < P; ++i) col[i] = (((col[i] + r0[i] * d0) + r1[i] * d1) + r2[i] * d2) + r3[i] * d3;
        }
    }
    for (; t < n; ++t) {
        const double* __restrict row = D + t * P;
        for (std::ptrdiff_t j = j0; j < j1; ++j) {
            const double dj = row[j];
            double* __restrict col = cv + j * P;
#pragma omp simd
            for (std::ptrdiff_t i = j; i < P; ++i) col[i] += row[i] * dj;
        }
    }
}
}  // namespace

// .cpp:170-199.  The reference recomputes mean and covariance over the WHOLE chain history every adaptation_period
// iterations (O(t P^2) each; Eigen's GEMM there).  Same here, every sum in history order: the centred history is laid out
// once, the lower triangle is accumulated column by column (d_i d_j == d_j d_i bit for bit, so the upper triangle is a copy),
// and for a long history the columns are shared among the host threads -- a column's sums never leave their thread, so the
// result does not depend on the thread count.
void MetropolisHastingsSampler::recomputeFullCovariance(Chain& c) const {
    const auto P = static_cast<std::ptrdiff_t>(n_params_);
    const std::ptrdiff_t n = static_cast<std::ptrdiff_t>(c.hist.size()) / P;
    if (static_cast<size_t>(n) < static_cast<size_t>(n_params_) + 10) return;
    // the mean: the history summed in order -- which is what the running sum kept by pushHistory holds, addition by addition
    VectorXd mean = c.hist_sum;
    mean /= static_cast<double>(n);
    c.running_mean = mean;
    std::vector<double>& centred = c.centred;                            // work buffer kept with the chain (50 MB at 100 000 states)
    centred.resize(static_cast<size_t>(n) * static_cast<size_t>(P));
    const double* m = mean.data();
    const bool threads = n_chains_ == 1 && n * P * P > (1 << 22);        // a long history: rows / column pairs shared among the host threads
#pragma omp parallel for schedule(static) if (threads)
    for (std::ptrdiff_t t = 0; t < n; ++t) {
        const double* v = c.hist.data() + t * P;
        double* row = centred.data() + t * P;
        for (std::ptrdiff_t i = 0; i < P; ++i) row[i] = v[i] - m[i];
    }
    MatrixXd cov = MatrixXd::Zero(P, P);
    double* cv = cov.data();
    const double* D = centred.data();
    auto accumulate = [&](std::ptrdiff_t j0, std::ptrdiff_t j1) { accumulate_lower(D, n, P, j0, j1, cv); };
    if (threads) {
        const std::ptrdiff_t blocks = (P + 1) / 2;
#pragma omp parallel for schedule(dynamic, 1)
        for (std::ptrdiff_t b = 0; b < blocks; ++b) accumulate(2 * b, std::min<std::ptrdiff_t>(2 * b + 2, P));
    } else {
        accumulate(0, P);
    }
    for (std::ptrdiff_t j = 0; j < P; ++j)
        for (std::ptrdiff_t i = j + 1; i < P; ++i) cv[i * P + j] = cv[j * P + i];
    cov *= 1.0 / static_cast<double>(n - 1);
    c.cov = ((2.38 * 2.38) / static_cast<double>(n_params_)) * cov + regularization_epsilon_ * MatrixXd::Identity(P, P);
    MatrixXd L;
    if (linalg::cholesky_lower(c.cov, L)) c.chol = L;
}

void MetropolisHastingsSampler::adaptGlobalScale(ScaleState& c, bool accepted, int step) const {
    if (!adapt_scale_) return;
    c.recent.push_back(accepted ? 1 : 0);
    c.recent_sum += accepted ? 1 : 0;
    if (c.recent.size() > 1000) { c.recent_sum -= c.recent.front(); c.recent.pop_front(); }
    const double rate = c.recent.empty() ? 0.0 : static_cast<double>(c.recent_sum) / static_cast<double>(c.recent.size());
    if (c.recent.size() >= 1000 && rate < 0.001) {
        c.log_scale -= 0.7;
        c.emergency_shrink_count++;
    } else if (rate < 0.02 && c.recent.size() >= 500) {
        const double g = std::min(5.0 / std::sqrt(static_cast<double>(step) + 1.0), 0.3);
        c.log_scale += g * (0.0 - target_acceptance_rate_);
    } else {
        const double g = std::min(1.0 / std::sqrt(static_cast<double>(step) + 1.0), 0.1);
        c.log_scale += g * ((accepted ? 1.0 : 0.0) - target_acceptance_rate_);
    }
    if (c.global_scale <= 0.011 && rate > 0.15 && rate < 0.30) c.log_scale += 0.01;
    c.log_scale = std::max(std::min(c.log_scale, 2.3), -6.9);
    c.global_scale = detm::exp(c.log_scale);
}

// 1. adaptation (only after burn-in), .cpp:286-303
void MetropolisHastingsSampler::adaptKernel(Chain& c, int t) const {
    if (t <= burn_in_) return;
    const auto P = static_cast<std::ptrdiff_t>(n_params_);
    ownKernel(c);
    updateCovarianceRank1(c, t);
    if (t % adaptation_period_ == 0) {
        recomputeFullCovariance(c);
        MatrixXd L;
        if (linalg::cholesky_lower(c.cov + regularization_epsilon_ * MatrixXd::Identity(P, P), L)) c.chol = L;
    }
}

// 2. proposal  Y = X + scale * L z,  z ~ N(0, I)   (generateProposal, .cpp:91-102), 2b. constraints (reflection in MCMC mode, .cpp:308)
void MetropolisHastingsSampler::drawProposal(std::mt19937& gen, const Chain& c, double scale, const double* x, IParameterManager& pm,
                                             double* out) const {
    const auto P = static_cast<std::ptrdiff_t>(n_params_);
    VectorXd z(P);
    PolarNormal dist;                                   // a fresh distribution per proposal, like the reference's local object
    for (std::ptrdiff_t i = 0; i < P; ++i) z(i) = dist(gen);
    // L is a lower Cholesky factor (diagonal for the start kernel): the structural zeros are skipped, which leaves every
    // sum unchanged (they would add +-0)
    const MatrixXd& L = c.own_kernel ? c.chol : shared_chol_;
    VectorXd step = VectorXd::Zero(P);
    if (!c.own_kernel && shared_diagonal_) {
        for (std::ptrdiff_t i = 0; i < P; ++i) step(i) = L(i, i) * z(i);
    } else {
        for (std::ptrdiff_t j = 0; j < P; ++j) {
            const double zj = z(j);
            for (std::ptrdiff_t i = j; i < P; ++i) step(i) += L(i, j) * zj;
        }
    }
    VectorXd y(P);
    for (std::ptrdiff_t i = 0; i < P; ++i) y(i) = x[i] + scale * step(i);
    const VectorXd yc = pm.applyConstraints(y);
    std::copy(yc.data(), yc.data() + P, out);
}

void MetropolisHastingsSampler::propose(IParameterManager& pm, double* out) {
    const auto P = static_cast<std::ptrdiff_t>(n_params_);
#pragma omp parallel for schedule(static) if (n_chains_ >= 8)      // one chain (the reference's run): no thread team to wake per iteration
    for (int ci = 0; ci < n_chains_; ++ci) {
        Chain& c = chains_[static_cast<size_t>(ci)];
        adaptKernel(c, c.t);
        drawProposal(c.gen, c, c.global_scale, cur_x_.data() + static_cast<std::ptrdiff_t>(ci) * P, pm,
                     prop_x_.data() + static_cast<std::ptrdiff_t>(ci) * P);
    }
    std::copy(prop_x_.begin(), prop_x_.end(), out);
}

// Steps 4-7 of the loop for ONE chain at ITS iteration c.t (.cpp:310-367): log-space accept test, state, MAP estimate, scale
// adaptation, history, thinned samples.  The proposal is row ci of prop_x_.
bool MetropolisHastingsSampler::acceptOne(Chain& c, int ci, double proposed_logpost) {
    const auto P = static_cast<std::ptrdiff_t>(n_params_);
    const int t = c.t;
    const double plp = safeValue(proposed_logpost);
    double& clp = cur_lp_[static_cast<size_t>(ci)];
    const double log_ratio = plp - clp;
    bool acc = false;
    if (log_ratio >= 0.0) {
        acc = true;
    } else {
        std::uniform_real_distribution<double> u(0.0, 1.0);      // drawn ONLY for downhill proposals (.cpp:323-329)
        if (detm::log(u(c.gen)) < log_ratio) acc = true;
    }
    double* x = cur_x_.data() + static_cast<std::ptrdiff_t>(ci) * P;
    if (acc) {
        std::copy(prop_x_.begin() + static_cast<std::ptrdiff_t>(ci) * P, prop_x_.begin() + static_cast<std::ptrdiff_t>(ci + 1) * P, x);
        clp = plp;
        c.accepted++;
        if (clp > c.best_lp) { c.best_lp = clp; c.best_x = VectorXd::FromPointer(x, P); }
    }
    if (adapt_scale_) adaptGlobalScale(c, acc, t);
    if (keep_history_) pushHistory(c, x);
    if (store_samples_ && (t % thinning_ == 0)) { c.samples.push_back(VectorXd::FromPointer(x, P)); c.sample_lp.push_back(clp); }
    ++c.t;
    return acc;
}

void MetropolisHastingsSampler::accept(const double* proposed_logpost, uint8_t* accepted_out) {
#pragma omp parallel for schedule(static) if (n_chains_ >= 8)
    for (int ci = 0; ci < n_chains_; ++ci) {
        const bool acc = acceptOne(chains_[static_cast<size_t>(ci)], ci, proposed_logpost[ci]);
        if (accepted_out) accepted_out[ci] = acc ? 1 : 0;
    }
    ++t_;
}

// upto >= 0: only the samples of iterations <= upto (a look-ahead run writes its checkpoints with every chain cut at the
// reporting iteration, although some chains are already past it)
OptimizationResult MetropolisHastingsSampler::result(int upto) const {
    OptimizationResult r;
    long accepted = 0;
    for (const Chain& c : chains_) {
        if (c.best_lp > r.bestObjectiveValue || r.bestParameters.size() == 0) { r.bestObjectiveValue = c.best_lp; r.bestParameters = c.best_x; }
        accepted += c.accepted;
        const size_t keep = upto < 0 ? c.samples.size() : std::min(c.samples.size(), static_cast<size_t>(1 + upto / thinning_));
        r.samples.insert(r.samples.end(), c.samples.begin(), c.samples.begin() + static_cast<std::ptrdiff_t>(keep));     // chain-major
        r.sampleObjectiveValues.insert(r.sampleObjectiveValues.end(), c.sample_lp.begin(), c.sample_lp.begin() + static_cast<std::ptrdiff_t>(keep));
    }
    r.finalCovariance = chains_.empty() || !chains_.front().own_kernel ? shared_cov_ : chains_.front().cov;
    r.additionalStats["acceptance_rate"] = static_cast<double>(accepted) / (static_cast<double>(iterations_) * std::max(1, n_chains_));
    r.additionalStats["final_scale"] = chains_.empty() ? 1.0 : chains_.front().global_scale;
    r.additionalStats["burn_in"] = burn_in_;
    r.additionalStats["total_iterations"] = iterations_;
    r.additionalStats["n_chains"] = n_chains_;
    return r;
}

// How many iterations ahead to evaluate.  Proposal k of a window is reached with probability (1 - rate)^(k-1) and then commits
// exactly one iteration, so K proposals commit g(K) = (1 - (1 - rate)^K) / rate iterations on average; they cost the host
// n_chains * K * c of arithmetic (normals, L z, reflection; c measured per window, ~6 us on one thread, less when the chains are
// shared among the threads) on top of the launch time L (measured, ~0.6 ms whatever the launch holds up to ~4 000 sets).
// K = argmax g(K) / (L + n K c), with the acceptance rate of the last <= 1000 iterations (the window the scale adaptation keeps
// anyway).  Measured on the device objective at 21 % acceptance, one chain: K = 8 5.7 k, 16 6.2 k, 32 5.9 k, 128 5.0 k iterations/s.
int MetropolisHastingsSampler::windowLength(const Chain& c, int running, int share) const {
    if (lookahead_ > 1) return std::min(lookahead_, share);
    if (calls_seen_ < 4) return std::min(calls_seen_ % 2 == 0 ? 4 : 16, share);      // the probe windows (see runLookahead)
    const double rate = c.recent.size() < 50 ? 0.234
                                             : std::min(std::max(static_cast<double>(c.recent_sum) / static_cast<double>(c.recent.size()), 0.02), 0.9);
    // a call of the objective over B rows is taken to cost launch + B * row: the device objective is all `launch` (~0.6 ms, ~2 us per
    // row), an objective that scores its rows one after the other on the host is all `row` -- and then looking ahead only wastes
    // evaluations, K comes out as 1 and the run is the sequential one
    const double launch = launch_seconds_ > 0 ? launch_seconds_ : 6e-4, per = (proposal_seconds_ > 0 ? proposal_seconds_ : 6e-6) + row_seconds_;
    const double commit = commit_seconds_ > 0 ? commit_seconds_ : 3e-6;       // host time per COMMITTED iteration (draws replayed, accept, adaptation)
    int best = 1;
    double best_rate = 0.0, miss = 1.0;
    for (int k = 1; k <= std::min(share, 128); ++k) {
        miss *= 1.0 - rate;
        const double g = (1.0 - miss) / rate;
        const double r = g / (launch + static_cast<double>(running) * (k * per + g * commit));
        if (r > best_rate) { best_rate = r; best = k; }
    }
    return best;
}

// K iterations of every chain per device launch (see optimizers.hpp).  A chain object only ever changes through the calls the
// sequential loop makes -- adaptKernel, the generator draws, acceptOne -- in the sequential order; the speculation works on
// copies of the generator and of the scale state.  Chains advance by different amounts per launch (each up to its first
// accepted proposal), so every chain carries its own iteration counter; a launch holds at most LOOKAHEAD_SETS proposals
// (the range in which its cost does not depend on their number), shared evenly among the chains still running.
void MetropolisHastingsSampler::runLookahead(IObjectiveFunction& f, IParameterManager& pm, const std::string& dir) {
    const auto P = static_cast<std::ptrdiff_t>(n_params_);
    const int n = n_chains_;
    std::vector<int> K(static_cast<size_t>(n));
    std::vector<std::ptrdiff_t> first(static_cast<size_t>(n) + 1);
    std::vector<double> props;
    static const bool debug_timing = std::getenv("SEPAIHRD_HOST_DEBUG_TIMING") != nullptr;
    double dbg[5] = {0, 0, 0, 0, 0};
    long dbg_windows = 0;
    struct DebugReport {
        const double* d; const long* w; bool on;
        ~DebugReport() { if (on) std::fprintf(stderr, "look-ahead windows %ld: plan %.3f s, propose %.3f s, objective %.3f s, commit %.3f s, tail %.3f s\n", *w, d[0], d[1], d[2], d[3], d[4]); }
    } debug_report{dbg, &dbg_windows, debug_timing};
    while (!done()) {
        const auto clock_plan = std::chrono::steady_clock::now();
        int running = 0;
        for (const Chain& c : chains_) running += c.t < iterations_;
        const int share = std::max(1, LOOKAHEAD_SETS / std::max(running, 1));
        first[0] = 0;
        for (int ci = 0; ci < n; ++ci) {
            const Chain& c = chains_[static_cast<size_t>(ci)];
            int k = std::min(windowLength(c, running, share), iterations_ - c.t);
            // the proposal kernel L must not change inside the window: an iteration that refactors it (.cpp:292-302) may open
            // a window (its adaptation runs before anything is drawn), it may not sit inside one
            for (int j = 1; j < k; ++j)
                if (c.t + j > burn_in_ && (c.t + j) % adaptation_period_ == 0) { k = j; break; }
            K[static_cast<size_t>(ci)] = std::max(k, 0);
            first[static_cast<size_t>(ci) + 1] = first[static_cast<size_t>(ci)] + K[static_cast<size_t>(ci)];
        }
        const int64_t total = first[static_cast<size_t>(n)];
        props.resize(static_cast<size_t>(total) * static_cast<size_t>(P));
        const auto clock0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(static) if (n >= 8)
        for (int ci = 0; ci < n; ++ci) {
            Chain& c = chains_[static_cast<size_t>(ci)];
            const int k = K[static_cast<size_t>(ci)];
            if (k == 0) continue;
            adaptKernel(c, c.t);
            if (k == 1) {                                            // nothing to look ahead to: the sequential iteration, drawn from the chain's own generator
                drawProposal(c.gen, c, c.global_scale, cur_x_.data() + static_cast<std::ptrdiff_t>(ci) * P, pm, props.data() + first[static_cast<size_t>(ci)] * P);
                continue;
            }
            std::mt19937 gen = c.gen;
            ScaleState sc = c;
            std::uniform_real_distribution<double> u01(0.0, 1.0);
            for (int j = 0; j < k; ++j) {
                drawProposal(gen, c, sc.global_scale, cur_x_.data() + static_cast<std::ptrdiff_t>(ci) * P, pm,
                             props.data() + (first[static_cast<size_t>(ci)] + j) * P);
                (void)u01(gen);                                      // a rejected proposal was a downhill one: its uniform is drawn
                if (adapt_scale_) adaptGlobalScale(sc, false, c.t + j);
            }
        }
        const auto clock1 = std::chrono::steady_clock::now();
        const std::vector<double> plp = evaluate_rows(f, props, total, P);
        const auto clock2 = std::chrono::steady_clock::now();
        speculated_ += total;
        long committed = 0;
#pragma omp parallel for schedule(static) reduction(+ : committed) if (n >= 8)
        for (int ci = 0; ci < n; ++ci) {
            Chain& c = chains_[static_cast<size_t>(ci)];
            for (int j = 0; j < K[static_cast<size_t>(ci)]; ++j) {
                if (j > 0) adaptKernel(c, c.t);                      // rank-1 update of the covariance only (no refactoring, see above)
                if (K[static_cast<size_t>(ci)] > 1) {                // the generator makes the draws of this iteration's proposal (a window of one drew them itself)
                    PolarNormal dist;
                    for (std::ptrdiff_t i = 0; i < P; ++i) (void)dist(c.gen);
                }
                const double* row = props.data() + (first[static_cast<size_t>(ci)] + j) * P;
                std::copy(row, row + P, prop_x_.begin() + static_cast<std::ptrdiff_t>(ci) * P);
                ++committed;
                if (acceptOne(c, ci, plp[static_cast<size_t>(first[static_cast<size_t>(ci)] + j)])) break;   // the uniform (if downhill), state, scale, history, samples
            }
        }
        committed_ += committed;
        const auto clock3 = std::chrono::steady_clock::now();
        {   // what a proposal costs the host (drawing + its share of the commit) and what a launch costs: running means for windowLength
            const double per = std::chrono::duration<double>(clock1 - clock0).count() / static_cast<double>(std::max<int64_t>(total, 1));
            const double com = std::chrono::duration<double>(std::chrono::steady_clock::now() - clock2).count() / static_cast<double>(std::max<long>(committed, 1));
            const double call = std::chrono::duration<double>(clock2 - clock1).count();
            proposal_seconds_ = proposal_seconds_ > 0 ? 0.8 * proposal_seconds_ + 0.2 * per : per;
            commit_seconds_ = commit_seconds_ > 0 ? 0.8 * commit_seconds_ + 0.2 * com : com;
            // Split the call into its fixed part and its per-row part from the first four windows, which are given the lengths 4, 16,
            // 4, 16 on purpose (fresh proposals, so a likelihood cache cannot fake a cheap call); the cheaper call of each pair counts
            // (the first call of a size may pay for buffers that grow), and a per-row part below 30 % of the call is taken as zero:
            // the split only has to tell an objective that scores its rows one after the other from one launch for all of them.
            if (calls_seen_ < 4) {
                double& rows = (calls_seen_ % 2 == 0) ? probe_rows_[0] : probe_rows_[1];
                double& secs = (calls_seen_ % 2 == 0) ? probe_seconds_[0] : probe_seconds_[1];
                if (calls_seen_ < 2 || call / static_cast<double>(total) < secs / rows) { rows = static_cast<double>(total); secs = call; }
                if (calls_seen_ == 3 && probe_rows_[1] > probe_rows_[0]) {
                    const double row = std::max(0.0, (probe_seconds_[1] - probe_seconds_[0]) / (probe_rows_[1] - probe_rows_[0]));
                    row_seconds_ = row * probe_rows_[1] < 0.3 * probe_seconds_[1] ? 0.0 : row;
                }
            }
            const double launch = std::max(call - row_seconds_ * static_cast<double>(total), 0.05 * call);
            launch_seconds_ = launch_seconds_ > 0 && calls_seen_ > 3 ? 0.8 * launch_seconds_ + 0.2 * launch : launch;
            ++calls_seen_;
        }
        // the lockstep loop writes a checkpoint after iteration t when (t + 1) % report_interval == 0 (.cpp:380-382): here once ALL
        // chains have completed such an iteration, with every chain cut at it
        const int t_before = t_;
        t_ = iterations_;
        for (const Chain& c : chains_) t_ = std::min(t_, c.t);
        if (!dir.empty() && write_checkpoints_) {
            int last = -1;
            for (int t = t_before; t < t_; ++t) if ((t + 1) % report_interval_ == 0) last = t;
            if (last >= 0) saveCheckpoint(result(last), pm, false, dir);
        }
        if (debug_timing) {
            auto sec = [](auto a, auto b) { return std::chrono::duration<double>(b - a).count(); };
            dbg[0] += sec(clock_plan, clock0); dbg[1] += sec(clock0, clock1); dbg[2] += sec(clock1, clock2); dbg[3] += sec(clock2, clock3);
            dbg[4] += sec(clock3, std::chrono::steady_clock::now());
            ++dbg_windows;
        }
    }
}

OptimizationResult MetropolisHastingsSampler::optimize(const VectorXd& initial, IObjectiveFunction& f, IParameterManager& pm) {
    if (auto* spm = dynamic_cast<SEPAIHRDParameterManager*>(&pm)) spm->setConstraintMode(ConstraintMode::MCMC_REFLECT);   // .cpp:207-210
    const int64_t P = initial.size();
    // every chain starts from the same point: one evaluation, replicated
    const double lp0 = safeValue(f.calculate(initial));
    std::vector<double> lp(static_cast<size_t>(n_chains_), lp0);
    begin(initial, lp.data(), pm);
    std::vector<double> prop(static_cast<size_t>(n_chains_) * static_cast<size_t>(P));
    const std::string dir = (store_samples_ && (write_checkpoints_ || write_trace_)) ? traceDirectory() : std::string();
    speculated_ = committed_ = 0;
    launch_seconds_ = proposal_seconds_ = row_seconds_ = commit_seconds_ = 0.0;
    calls_seen_ = 0;
    if (lookahead_ != 1 && 2 * n_chains_ <= LOOKAHEAD_SETS) runLookahead(f, pm, dir);
    while (!done()) {
        const int t = t_;
        propose(pm, prop.data());
        const std::vector<double> plp = evaluate_rows(f, prop, n_chains_, P);
        accept(plp.data());
        if (!dir.empty() && write_checkpoints_ && (t + 1) % report_interval_ == 0) saveCheckpoint(result(), pm, false, dir);   // .cpp:380-382
    }
    OptimizationResult r = result();
    if (!dir.empty() && write_checkpoints_) saveCheckpoint(r, pm, true, dir);                                                  // .cpp:399-401
    if (!dir.empty() && write_trace_ && !r.samples.empty())                                                                    // .cpp:403-409
        saveSamplesToCSV(r.samples, r.sampleObjectiveValues, pm.getParameterNames(), dir + "/posterior_trace.csv");
    return r;
}

namespace {
std::string g_default_trace_dir;
}
void MetropolisHastingsSampler::setDefaultOutputDirectory(const std::string& dir) { g_default_trace_dir = dir; }

std::string MetropolisHastingsSampler::traceDirectory() const {
    namespace fs = std::filesystem;
    std::string dir = !output_dir_.empty() ? output_dir_ : g_default_trace_dir;
    if (dir.empty()) {
        // FileUtils::getProjectRoot (src/utils/FileUtils.cpp:25-46): the working directory or one of its five ancestors that
        // holds data/, include/ and src/
        std::error_code ec;
        fs::path cur = fs::current_path(ec);
        for (int i = 0; i <= 5 && !ec && !cur.empty(); ++i) {
            if (fs::exists(cur / "data") && fs::exists(cur / "include") && fs::exists(cur / "src")) { dir = (cur / "data" / "mcmc_samples").string(); break; }
            if (!cur.has_parent_path() || cur.parent_path() == cur) break;
            cur = cur.parent_path();
        }
    }
    if (dir.empty()) return dir;
    std::error_code ec;
    fs::create_directories(dir, ec);
    return ec ? std::string() : dir;
}

// iter, log_posterior, one column per parameter; everything after the index in %.6e (the reference sets std::scientific and
// setprecision(6) on the stream before the objective value and never resets them, .cpp:429-433)
void MetropolisHastingsSampler::saveSamplesToCSV(const std::vector<VectorXd>& samples, const std::vector<double>& values,
                                                 const std::vector<std::string>& names, const std::string& filepath, size_t first) {
    std::FILE* file = std::fopen(filepath.c_str(), "w");
    if (!file) return;
    std::fputs("iter,log_posterior", file);
    for (const auto& n : names) std::fprintf(file, ",%s", n.c_str());
    std::fputc('\n', file);
    for (size_t i = first; i < samples.size(); ++i) {
        std::fprintf(file, "%zu,%.6e", i, values[i]);
        for (std::ptrdiff_t j = 0; j < samples[i].size(); ++j) std::fprintf(file, ",%.6e", samples[i](j));
        std::fputc('\n', file);
    }
    std::fclose(file);
}

void MetropolisHastingsSampler::saveCheckpoint(const OptimizationResult& res, IParameterManager& pm, bool final, const std::string& dir) const {   // .cpp:440-469
    if (res.samples.empty()) return;
    const size_t first = final ? 0 : (res.samples.size() > 5000 ? res.samples.size() - 5000 : 0);
    saveSamplesToCSV(res.samples, res.sampleObjectiveValues, pm.getParameterNames(),
                     dir + (final ? "/posterior_trace_final.csv" : "/posterior_trace_checkpoint.csv"), first);
}

// =====================================================================================================================
// Particle swarm (ParticleSwarmOptimizer.cpp, whole class).  optimize() is the reference's algorithm with every swarm
// evaluation handed to calculateBatch as ONE batch; the STANDARD / GLOBAL_BEST swarm with linear schedules also exists
// step-wise (begin / tell / step: shardable over processes) and device-resident (beginDevice / evaluateDevice / stepDevice).
//
// Where batching forces a choice the reference leaves open:
//   * the reference's OpenMP loop lets particle i read personal bests its neighbours may or may not have updated yet in
//     the same iteration (sequentially: always for j < i); here every neighbourhood best is taken from the personal bests
//     at the START of the iteration -- the synchronous swarm;
//   * RANDOM_DYNAMIC shuffles with the shared master generator inside that OpenMP loop (a data race in the reference);
//     here the shuffles run in particle order before the updates, which is the reference's single-thread order;
//   * HYBRID passes an empty mean-best vector into quantumPSOUpdate (it is only computed for QUANTUM, .cpp:359-362:
//     out-of-bounds reads); here the mean of the personal bests is computed for HYBRID as well;
//   * std::sort on equal fitness values (opposition selection, restart) is replaced by std::stable_sort, so ties keep
//     the lower particle index;
//   * the <= 3 sequential trial evaluations of the elitist learning strategy are generated ahead and evaluated as one
//     batch of 3; the master generator is then rewound to where the reference would have stopped drawing.
// =====================================================================================================================
void ParticleSwarmOptimization::configure(const std::map<std::string, double>& s) {
    // .cpp:10-104: only the keys present are touched, each validated with the reference's message
    for (const auto& [key, value] : s) {
        if (key == "iterations") { if (value <= 0) throw std::invalid_argument("iterations must be positive"); iterations_ = static_cast<int>(value); }
        else if (key == "swarm_size") { if (value <= 0) throw std::invalid_argument("swarm_size must be positive"); swarm_size_ = static_cast<int>(value); }
        else if (key == "omega_start") { if (value < 0) throw std::invalid_argument("omega_start must be non-negative"); omega_start_ = value; }
        else if (key == "omega_end") { if (value < 0) throw std::invalid_argument("omega_end must be non-negative"); omega_end_ = value; }
        else if (key == "c1_initial") { if (value < 0) throw std::invalid_argument("c1_initial must be non-negative"); c1_initial_ = value; }
        else if (key == "c1_final") { if (value < 0) throw std::invalid_argument("c1_final must be non-negative"); c1_final_ = value; }
        else if (key == "c2_initial") { if (value < 0) throw std::invalid_argument("c2_initial must be non-negative"); c2_initial_ = value; }
        else if (key == "c2_final") { if (value < 0) throw std::invalid_argument("c2_final must be non-negative"); c2_final_ = value; }
        else if (key == "report_interval") { if (value <= 0) throw std::invalid_argument("report_interval must be positive"); report_interval_ = static_cast<int>(value); }
        else if (key == "variant") {
            const int v = static_cast<int>(value);
            if (v < 0 || v > 4) throw std::invalid_argument("variant must be between 0 and 4");
            variant_ = static_cast<PSOVariant>(v);
        } else if (key == "topology") {
            const int t = static_cast<int>(value);
            if (t < 0 || t > 3) throw std::invalid_argument("topology must be between 0 and 3");
            topology_ = static_cast<TopologyType>(t);
        }
        else if (key == "use_opposition_learning") use_opposition_learning_ = (value != 0.0);
        else if (key == "use_parallel") use_parallel_ = (value != 0.0);           // kept for the settings file; the batch IS the parallel loop
        else if (key == "use_adaptive_parameters") use_adaptive_parameters_ = (value != 0.0);
        else if (key == "diversity_threshold") diversity_threshold_ = value;
        else if (key == "restart_threshold") restart_threshold_ = value;
        else if (key == "quantum_beta") quantum_beta_ = value;
        else if (key == "levy_alpha") levy_alpha_ = value;
        else if (key == "max_stagnation") { if (value <= 0) throw std::invalid_argument("max_stagnation must be positive"); max_stagnation_ = static_cast<int>(value); }
        else if (key == "log_evolutionary_state") log_evolutionary_state_ = (value != 0.0);
        // extensions of this build: sharding, repeatable runs, device-resident swarm
        else if (key == "particle_offset") particle_offset_ = static_cast<long>(value);
        else if (key == "local_count") local_count_setting_ = static_cast<int>(value);
        else if (key == "seed") { has_seed_ = true; seed_ = static_cast<unsigned>(value); }
        else if (key == "device_resident") device_resident_ = (value != 0.0);
    }
}

bool ParticleSwarmOptimization::isBasicSwarm() const {
    return variant_ == PSOVariant::STANDARD && topology_ == TopologyType::GLOBAL_BEST && !use_opposition_learning_ && !use_adaptive_parameters_;
}

ParticleSwarmOptimization::~ParticleSwarmOptimization() {
    if (dev_swarm_) sepaihrd_swarm_destroy(dev_swarm_);
}

void ParticleSwarmOptimization::setupRun(IParameterManager& pm) {
    n_ = static_cast<int>(pm.getParameterCount());
    local_ = local_count_setting_ >= 0 ? local_count_setting_ : swarm_size_ - static_cast<int>(particle_offset_);
    if (particle_offset_ < 0 || particle_offset_ + local_ > swarm_size_) throw std::invalid_argument("particle_offset/local_count outside the swarm");
    lb_.resize(static_cast<size_t>(n_)); ub_.resize(static_cast<size_t>(n_));
    for (int k = 0; k < n_; ++k) { lb_[static_cast<size_t>(k)] = pm.getLowerBoundForParamIndex(k); ub_[static_cast<size_t>(k)] = pm.getUpperBoundForParamIndex(k); }
    if (has_seed_) rng_.seed(seed_); else rng_.seed(std::random_device{}());
    uniform_dist_.reset(); normal_dist_.reset();
    const size_t tot = static_cast<size_t>(local_) * static_cast<size_t>(n_);
    pos_.assign(tot, 0.0); vel_.assign(tot, 0.0); pbest_.assign(tot, 0.0);
    pbest_val_.assign(static_cast<size_t>(local_), -std::numeric_limits<double>::infinity());
    cur_fit_.assign(static_cast<size_t>(local_), -std::numeric_limits<double>::infinity());
    success_rate_.assign(static_cast<size_t>(local_), 0.0);
    success_count_.assign(static_cast<size_t>(local_), 0); total_updates_.assign(static_cast<size_t>(local_), 0);
    gbest_.assign(static_cast<size_t>(n_), 0.0);
    gbest_value_ = -std::numeric_limits<double>::infinity();
    first_tell_ = true;
    restarts_ = 0; els_trials_ = 0; evaluations_ = 0;
    if (dev_swarm_) { sepaihrd_swarm_destroy(dev_swarm_); dev_swarm_ = nullptr; }
}

// one seed per particle of the WHOLE swarm from the master generator (.cpp:268-270, :365-371): every shard draws them all
std::vector<uint32_t> ParticleSwarmOptimization::drawSeeds() {
    std::vector<uint32_t> seeds(static_cast<size_t>(swarm_size_));
    for (auto& sd : seeds) sd = static_cast<uint32_t>(rng_());
    return seeds;
}

void ParticleSwarmOptimization::coefficients(int iter, double& omega, double& c1, double& c2) const {
    const double ratio = (iterations_ > 1) ? static_cast<double>(iter) / (iterations_ - 1) : 0.0;    // .cpp:352-357
    omega = omega_start_ + (omega_end_ - omega_start_) * ratio;
    c1 = c1_initial_ + (c1_final_ - c1_initial_) * ratio;
    c2 = c2_initial_ + (c2_final_ - c2_initial_) * ratio;
}

namespace {
// the reference's per-iteration progress line (.cpp:184-211) goes to its Logger; here it is printed to stderr when the
// environment variable SEPAIHRD_PSO_REPORT is set (report_interval applies)
bool report_enabled() {
    static const bool on = std::getenv("SEPAIHRD_PSO_REPORT") != nullptr;
    return on;
}
void require_basic(const ParticleSwarmOptimization& s, const char* who) {
    if (!s.isBasicSwarm())
        throw std::invalid_argument(std::string(who) + ": the step-wise / device-resident swarm is the STANDARD variant on the GLOBAL_BEST topology "
                                    "(settings variant 0, topology 0, use_opposition_learning 0, use_adaptive_parameters 0); optimize() runs every other configuration");
}
}  // namespace

void ParticleSwarmOptimization::beginDevice(const VectorXd* init, IParameterManager& pm, sepaihrd_ctx* ctx) {
    if (!ctx) throw std::invalid_argument("beginDevice: null device context");
    require_basic(*this, "beginDevice");
    setupRun(pm);
    if (sepaihrd_swarm_create(ctx, swarm_size_, particle_offset_, local_, &dev_swarm_) != SEPAIHRD_OK)
        throw std::runtime_error(std::string("sepaihrd_swarm_create: ") + sepaihrd_last_error());
    const std::vector<uint32_t> seeds = drawSeeds();
    const bool use_init = init != nullptr && init->size() == n_;
    if (sepaihrd_swarm_init(dev_swarm_, seeds.data(), use_init ? init->data() : nullptr) != SEPAIHRD_OK)
        throw std::runtime_error(std::string("sepaihrd_swarm_init: ") + sepaihrd_last_error());
}

std::pair<double, int> ParticleSwarmOptimization::evaluateDevice(double* best_position) {
    if (!dev_swarm_) throw std::logic_error("evaluateDevice before beginDevice");
    double value = 0.0;
    int64_t index = -1;
    if (sepaihrd_swarm_evaluate(dev_swarm_, &value, &index, best_position) != SEPAIHRD_OK)
        throw std::runtime_error(std::string("sepaihrd_swarm_evaluate: ") + sepaihrd_last_error());
    first_tell_ = false;
    evaluations_ += local_;
    return {value, static_cast<int>(index)};
}

void ParticleSwarmOptimization::stepDevice(int iter) {
    if (!dev_swarm_) throw std::logic_error("stepDevice before beginDevice");
    double omega, c1, c2;
    coefficients(iter, omega, c1, c2);
    const std::vector<uint32_t> seeds = drawSeeds();
    if (sepaihrd_swarm_step(dev_swarm_, seeds.data(), omega, c1, c2, gbest_.data()) != SEPAIHRD_OK)
        throw std::runtime_error(std::string("sepaihrd_swarm_step: ") + sepaihrd_last_error());
}

void ParticleSwarmOptimization::fetchPersonalBests(bool with_positions) {
    if (!dev_swarm_ || local_ == 0) return;
    bool ok = sepaihrd_swarm_read(dev_swarm_, SEPAIHRD_SWARM_PERSONAL_BEST, pbest_.data()) == SEPAIHRD_OK &&
              sepaihrd_swarm_read(dev_swarm_, SEPAIHRD_SWARM_PERSONAL_BEST_VALUES, pbest_val_.data()) == SEPAIHRD_OK;
    if (ok && with_positions)
        ok = sepaihrd_swarm_read(dev_swarm_, SEPAIHRD_SWARM_POSITIONS, pos_.data()) == SEPAIHRD_OK &&
             sepaihrd_swarm_read(dev_swarm_, SEPAIHRD_SWARM_VELOCITIES, vel_.data()) == SEPAIHRD_OK;
    if (!ok) throw std::runtime_error(std::string("sepaihrd_swarm_read: ") + sepaihrd_last_error());
}

// the draws of initializeSwarm (.cpp:273-298) for the local particles: position, then velocity, from the particle's own generator
void ParticleSwarmOptimization::drawInitialSwarm(const VectorXd* init) {
    const std::vector<uint32_t> seeds = drawSeeds();
#pragma omp parallel for schedule(static)
    for (int li = 0; li < local_; ++li) {
        const long gi = particle_offset_ + li;
        std::mt19937 local_rng(seeds[static_cast<size_t>(gi)]);
        std::uniform_real_distribution<> U(0.0, 1.0);
        double* p = pos_.data() + static_cast<size_t>(li) * n_;
        double* v = vel_.data() + static_cast<size_t>(li) * n_;
        if (gi == 0 && init != nullptr && init->size() == n_) {
            for (int k = 0; k < n_; ++k) p[k] = std::clamp((*init)(k), lb_[static_cast<size_t>(k)], ub_[static_cast<size_t>(k)]);
        } else {
            for (int k = 0; k < n_; ++k) p[k] = lb_[static_cast<size_t>(k)] + U(local_rng) * (ub_[static_cast<size_t>(k)] - lb_[static_cast<size_t>(k)]);
        }
        for (int k = 0; k < n_; ++k) {
            const double vmax = 0.2 * (ub_[static_cast<size_t>(k)] - lb_[static_cast<size_t>(k)]);
            v[k] = -vmax + 2 * vmax * U(local_rng);
        }
    }
}

void ParticleSwarmOptimization::begin(const VectorXd* init, IParameterManager& pm) {
    require_basic(*this, "begin");
    setupRun(pm);
    drawInitialSwarm(init);
}

std::pair<double, int> ParticleSwarmOptimization::tell(const double* fitness) {
    double best = -std::numeric_limits<double>::infinity();
    int best_i = -1;
    for (int li = 0; li < local_; ++li) {
        const double f = fitness[li];
        if (first_tell_ || f > pbest_val_[static_cast<size_t>(li)]) {       // .cpp:301-303 / :417-421
            pbest_val_[static_cast<size_t>(li)] = f;
            std::copy(pos_.begin() + static_cast<std::ptrdiff_t>(li) * n_, pos_.begin() + static_cast<std::ptrdiff_t>(li + 1) * n_,
                      pbest_.begin() + static_cast<std::ptrdiff_t>(li) * n_);
        }
        if (pbest_val_[static_cast<size_t>(li)] > best) { best = pbest_val_[static_cast<size_t>(li)]; best_i = li; }   // first maximum wins (.cpp:149-156)
    }
    first_tell_ = false;
    evaluations_ += local_;
    return {best, best_i};
}

void ParticleSwarmOptimization::setGlobalBest(double value, const double* position) {
    if (value > gbest_value_) {            // strict: an equal value keeps the earlier best (.cpp:150)
        gbest_value_ = value;
        gbest_.assign(position, position + n_);
    }
}

// standardPSOUpdate (.cpp:576-618) for local particle i towards `lbest`: r1_k, r2_k interleaved per dimension
void ParticleSwarmOptimization::standardPSOUpdate(int i, const double* lbest, double omega, double c1, double c2, std::mt19937& rng) {
    std::uniform_real_distribution<> U(0.0, 1.0);
    double* p = pos_.data() + static_cast<size_t>(i) * n_;
    double* v = vel_.data() + static_cast<size_t>(i) * n_;
    const double* pb = pbest_.data() + static_cast<size_t>(i) * n_;
    std::vector<double> r1(static_cast<size_t>(n_)), r2(static_cast<size_t>(n_));
    for (int k = 0; k < n_; ++k) { r1[static_cast<size_t>(k)] = U(rng); r2[static_cast<size_t>(k)] = U(rng); }
    for (int k = 0; k < n_; ++k) {
        const size_t kk = static_cast<size_t>(k);
        const double cognitive = c1 * (r1[kk] * (pb[k] - p[k]));
        const double social = c2 * (r2[kk] * (lbest[k] - p[k]));
        double vk = omega * v[k] + cognitive + social;
        const double vmax = 0.2 * (ub_[kk] - lb_[kk]);
        vk = std::clamp(vk, -vmax, vmax);
        double pk = p[k] + vk;
        if (pk < lb_[kk]) { pk = lb_[kk] + std::abs(pk - lb_[kk]); vk *= -0.5; }      // reflection with velocity dampening
        else if (pk > ub_[kk]) { pk = ub_[kk] - std::abs(pk - ub_[kk]); vk *= -0.5; }
        p[k] = std::clamp(pk, lb_[kk], ub_[kk]);
        v[k] = vk;
    }
}

void ParticleSwarmOptimization::step(int iter) {
    double omega, c1, c2;
    coefficients(iter, omega, c1, c2);
    const std::vector<uint32_t> seeds = drawSeeds();
#pragma omp parallel for schedule(static)
    for (int li = 0; li < local_; ++li) {
        std::mt19937 local_rng(seeds[static_cast<size_t>(particle_offset_ + li)]);
        standardPSOUpdate(li, gbest_.data(), omega, c1, c2, local_rng);
    }
}

MatrixXd ParticleSwarmOptimization::personalBestScatter(VectorXd& mean) const {
    const auto P = static_cast<std::ptrdiff_t>(n_);
    mean = VectorXd::Zero(P);
    for (int li = 0; li < local_; ++li)
        for (std::ptrdiff_t k = 0; k < P; ++k) mean(k) += pbest_[static_cast<size_t>(li) * n_ + static_cast<size_t>(k)];
    mean /= static_cast<double>(std::max(local_, 1));
    MatrixXd sc = MatrixXd::Zero(P, P);
    for (int li = 0; li < local_; ++li) {
        VectorXd d(P);
        for (std::ptrdiff_t k = 0; k < P; ++k) d(k) = pbest_[static_cast<size_t>(li) * n_ + static_cast<size_t>(k)] - mean(k);
        for (std::ptrdiff_t j = 0; j < P; ++j)
            for (std::ptrdiff_t i = 0; i < P; ++i) sc(i, j) += d(i) * d(j);
    }
    return sc;
}

// ---- the whole-swarm engine ------------------------------------------------------------------------------------------
void ParticleSwarmOptimization::evaluateSwarm(IObjectiveFunction& f, int first, std::vector<double>& fitness) {
    if (first >= local_) return;
    f.calculateBatch(pos_.data() + static_cast<size_t>(first) * n_, local_ - first, n_, fitness.data() + first);   // not sanitised (.cpp:298, :412)
    evaluations_ += local_ - first;
}

void ParticleSwarmOptimization::rescanGlobalBest() {                       // .cpp:149-156, :316-321: strict >, index order
    for (int i = 0; i < local_; ++i)
        if (pbest_val_[static_cast<size_t>(i)] > gbest_value_) {
            gbest_value_ = pbest_val_[static_cast<size_t>(i)];
            gbest_.assign(pbest_.begin() + static_cast<std::ptrdiff_t>(i) * n_, pbest_.begin() + static_cast<std::ptrdiff_t>(i + 1) * n_);
        }
}

void ParticleSwarmOptimization::permuteSwarm(const std::vector<int>& order) {
    auto rows = [&](std::vector<double>& a) {
        std::vector<double> b(a.size());
        for (int i = 0; i < local_; ++i)
            std::copy(a.begin() + static_cast<std::ptrdiff_t>(order[static_cast<size_t>(i)]) * n_, a.begin() + static_cast<std::ptrdiff_t>(order[static_cast<size_t>(i)] + 1) * n_,
                      b.begin() + static_cast<std::ptrdiff_t>(i) * n_);
        a.swap(b);
    };
    auto scal = [&](auto& a) {
        auto b = a;
        for (int i = 0; i < local_; ++i) b[static_cast<size_t>(i)] = a[static_cast<size_t>(order[static_cast<size_t>(i)])];
        a.swap(b);
    };
    rows(pos_); rows(vel_); rows(pbest_);
    scal(pbest_val_); scal(cur_fit_); scal(success_rate_); scal(success_count_); scal(total_updates_);
}

// oppositionBasedInitialization (.cpp:507-560).  The opposite particles are built but never evaluated before the selection
// (their pbest_value is the struct default -inf, ParticleSwarmOptimizer.hpp:258), so with finite fitness values the
// selection keeps the original swarm, sorted by fitness; an opposite particle enters only where an original scored -inf.
void ParticleSwarmOptimization::oppositionBasedInitialization() {
    const int N = local_;
    const double NEG_INF = -std::numeric_limits<double>::infinity();
    std::vector<std::pair<double, int>> cand;
    cand.reserve(static_cast<size_t>(2 * N));
    for (int i = 0; i < N; ++i) { cand.push_back({pbest_val_[static_cast<size_t>(i)], i}); cand.push_back({NEG_INF, i + N}); }
    std::stable_sort(cand.begin(), cand.end(), [](const auto& a, const auto& b) { return a.first > b.first; });
    std::vector<double> pos(pos_.size()), vel(vel_.size()), pb(pbest_.size()), pbv(static_cast<size_t>(N)), cf(static_cast<size_t>(N)), sr(static_cast<size_t>(N));
    std::vector<int> sc(static_cast<size_t>(N)), tu(static_cast<size_t>(N));
    for (int i = 0; i < N; ++i) {
        const int idx = cand[static_cast<size_t>(i)].second;
        const size_t d = static_cast<size_t>(i) * n_;
        if (idx < N) {
            const size_t s = static_cast<size_t>(idx) * n_;
            std::copy(pos_.begin() + s, pos_.begin() + s + n_, pos.begin() + d);
            std::copy(vel_.begin() + s, vel_.begin() + s + n_, vel.begin() + d);
            std::copy(pbest_.begin() + s, pbest_.begin() + s + n_, pb.begin() + d);
            pbv[static_cast<size_t>(i)] = pbest_val_[static_cast<size_t>(idx)]; cf[static_cast<size_t>(i)] = cur_fit_[static_cast<size_t>(idx)];
            sr[static_cast<size_t>(i)] = success_rate_[static_cast<size_t>(idx)]; sc[static_cast<size_t>(i)] = success_count_[static_cast<size_t>(idx)];
            tu[static_cast<size_t>(i)] = total_updates_[static_cast<size_t>(idx)];
        } else {
            const size_t s = static_cast<size_t>(idx - N) * n_;
            for (int k = 0; k < n_; ++k) {
                pos[d + k] = lb_[static_cast<size_t>(k)] + ub_[static_cast<size_t>(k)] - pos_[s + k];
                vel[d + k] = -vel_[s + k];
                pb[d + k] = pos[d + k];
            }
            pbv[static_cast<size_t>(i)] = NEG_INF; cf[static_cast<size_t>(i)] = NEG_INF; sr[static_cast<size_t>(i)] = 0.0; sc[static_cast<size_t>(i)] = 0; tu[static_cast<size_t>(i)] = 0;
        }
    }
    pos_.swap(pos); vel_.swap(vel); pbest_.swap(pb); pbest_val_.swap(pbv); cur_fit_.swap(cf); success_rate_.swap(sr); success_count_.swap(sc); total_updates_.swap(tu);
}

void ParticleSwarmOptimization::initializeSwarmFull(const VectorXd* init, IObjectiveFunction& f, IParameterManager& pm) {
    setupRun(pm);
    drawInitialSwarm(init);
    evaluateSwarm(f, 0, cur_fit_);
    pbest_ = pos_; pbest_val_ = cur_fit_;
    if (use_opposition_learning_) {                                      // .cpp:306-315: select, then evaluate the whole swarm again
        oppositionBasedInitialization();
        evaluateSwarm(f, 0, cur_fit_);
        pbest_ = pos_; pbest_val_ = cur_fit_;
    }
    gbest_value_ = -std::numeric_limits<double>::infinity();
    rescanGlobalBest();
    first_tell_ = false;
}

std::vector<int> ParticleSwarmOptimization::getNeighbors(int particle_idx) {
    std::vector<int> nb;
    const int N = swarm_size_;
    switch (topology_) {
        case TopologyType::GLOBAL_BEST:
            nb.resize(static_cast<size_t>(N));
            for (int i = 0; i < N; ++i) nb[static_cast<size_t>(i)] = i;
            break;
        case TopologyType::LOCAL_BEST:                                   // ring, two neighbours on each side
            nb.push_back(particle_idx);
            for (int j = 1; j <= 2; ++j) { nb.push_back((particle_idx - j + N) % N); nb.push_back((particle_idx + j) % N); }
            break;
        case TopologyType::VON_NEUMANN: {                                // ceil(sqrt(N))-wide grid, no wrap-around
            const int g = static_cast<int>(std::ceil(std::sqrt(N)));
            const int row = particle_idx / g, col = particle_idx % g;
            nb.push_back(particle_idx);
            if (row > 0) { const int idx = (row - 1) * g + col; if (idx < N) nb.push_back(idx); }
            if (row < g - 1) { const int idx = (row + 1) * g + col; if (idx < N) nb.push_back(idx); }
            if (col > 0) { const int idx = row * g + (col - 1); if (idx < N) nb.push_back(idx); }
            if (col < g - 1) { const int idx = row * g + (col + 1); if (idx < N) nb.push_back(idx); }
            break;
        }
        case TopologyType::RANDOM_DYNAMIC: {                             // four random others, redrawn on every call (master generator)
            nb.push_back(particle_idx);
            std::vector<int> cand;
            cand.reserve(static_cast<size_t>(std::max(N - 1, 0)));
            for (int i = 0; i < N; ++i) if (i != particle_idx) cand.push_back(i);
            std::shuffle(cand.begin(), cand.end(), rng_);
            const int k = std::min(4, static_cast<int>(cand.size()));
            nb.insert(nb.end(), cand.begin(), cand.begin() + k);
            break;
        }
    }
    return nb;
}

double ParticleSwarmOptimization::calculateEvolutionaryFactor() const {   // .cpp:446-482
    const double INF = std::numeric_limits<double>::infinity();
    double mean_distance = 0.0, max_distance = 0.0, mean_fitness = 0.0, max_fitness = -INF, min_fitness = INF;
    for (int i = 0; i < local_; ++i) {
        double d2 = 0.0;
        for (int k = 0; k < n_; ++k) { const double d = pos_[static_cast<size_t>(i) * n_ + static_cast<size_t>(k)] - gbest_[static_cast<size_t>(k)]; d2 += d * d; }
        const double dist = std::sqrt(d2);
        mean_distance += dist;
        max_distance = std::max(max_distance, dist);
        const double cf = cur_fit_[static_cast<size_t>(i)];
        mean_fitness += cf;
        max_fitness = std::max(max_fitness, cf);
        min_fitness = std::min(min_fitness, cf);
    }
    mean_distance /= swarm_size_;
    mean_fitness /= swarm_size_;
    const double fitness_range = (max_fitness - min_fitness) > 1e-10 ? (max_fitness - min_fitness) : 1e-10;
    const double distance_factor = (max_distance > 0) ? mean_distance / max_distance : 0.0;
    const double fitness_factor = (max_fitness - mean_fitness) / fitness_range;
    return 0.5 * distance_factor + 0.5 * (1.0 - fitness_factor);
}

ParticleSwarmOptimization::EvolutionaryState ParticleSwarmOptimization::estimateEvolutionaryState() const {   // .cpp:427-444
    const double ef = calculateEvolutionaryFactor();
    if (ef > 0.7) return EvolutionaryState::EXPLORATION;
    if (ef > 0.4) return EvolutionaryState::EXPLOITATION;
    if (ef > 0.2) return EvolutionaryState::CONVERGENCE;
    return EvolutionaryState::JUMPING_OUT;
}

void ParticleSwarmOptimization::adaptParameters(EvolutionaryState state, int iter, double& omega, double& c1, double& c2) {   // .cpp:484-505
    const double PI = 3.14159265358979323846;
    const double ratio = (iterations_ > 1) ? static_cast<double>(iter) / (iterations_ - 1) : 0.0;
    switch (state) {
        case EvolutionaryState::EXPLORATION:
            omega = 0.9 - 0.2 * ratio; c1 = 1.5 + 0.5 * std::sin(ratio * PI); c2 = 1.5 - 0.5 * std::sin(ratio * PI); break;
        case EvolutionaryState::EXPLOITATION:
            omega = 0.7 - 0.3 * ratio; c1 = 2.0 - ratio; c2 = 1.0 + ratio; break;
        case EvolutionaryState::CONVERGENCE:
            omega = 0.4 - 0.3 * ratio; c1 = 1.0 - 0.5 * ratio; c2 = 2.0 + 0.5 * ratio; break;
        case EvolutionaryState::JUMPING_OUT:                              // three draws from the master generator, in this order
            omega = 0.9 + 0.1 * uniform_dist_(rng_); c1 = 2.5 + uniform_dist_(rng_); c2 = 0.5 + uniform_dist_(rng_); break;
    }
    omega = std::clamp(omega, 0.1, 1.0);
    c1 = std::clamp(c1, 0.0, 4.0);
    c2 = std::clamp(c2, 0.0, 4.0);
}

// quantumPSOUpdate (.cpp:620-652): attractor between personal and global best, jump length from the mean-best distance
void ParticleSwarmOptimization::quantumPSOUpdate(int i, const std::vector<double>& mean_best, int iter, std::mt19937& rng) {
    std::uniform_real_distribution<> U(0.0, 1.0);
    double* p = pos_.data() + static_cast<size_t>(i) * n_;
    const double* pb = pbest_.data() + static_cast<size_t>(i) * n_;
    const double phi = U(rng);
    const double beta = quantum_beta_ * (1.0 - 0.5 * static_cast<double>(iter) / iterations_);
    for (int k = 0; k < n_; ++k) {
        const size_t kk = static_cast<size_t>(k);
        const double attractor = phi * pb[k] + (1 - phi) * gbest_[kk];
        const double u = U(rng);
        const double L = 2.0 * beta * std::abs(mean_best[kk] - p[k]);
        if (U(rng) < 0.5) p[k] = attractor + L * std::log(1.0 / u);
        else p[k] = attractor - L * std::log(1.0 / u);
        p[k] = std::clamp(p[k], lb_[kk], ub_[kk]);
    }
}

// generateLevyNumber (.cpp:917-934): Mantegna's algorithm; a fresh normal_distribution per number, as in the reference
double ParticleSwarmOptimization::generateLevyNumber(std::mt19937& rng) const {
    const double PI = 3.14159265358979323846;
    const double sigma_u = std::pow(std::tgamma(1 + levy_alpha_) * std::sin(PI * levy_alpha_ / 2) /
                                        (std::tgamma((1 + levy_alpha_) / 2) * levy_alpha_ * std::pow(2, (levy_alpha_ - 1) / 2)),
                                    1.0 / levy_alpha_);
    std::normal_distribution<> local_normal(0.0, 1.0);
    const double u = local_normal(rng) * sigma_u;
    const double v = std::max(std::abs(local_normal(rng)), 1e-10);
    const double levy_step = u / std::pow(v, 1.0 / levy_alpha_);
    return std::clamp(levy_step, -100.0, 100.0);
}

// levyFlightUpdate (.cpp:654-677): standard update towards the GLOBAL best, then an occasional heavy-tailed jump
void ParticleSwarmOptimization::levyFlightUpdate(int i, double omega, double c1, double c2, std::mt19937& rng) {
    standardPSOUpdate(i, gbest_.data(), omega, c1, c2, rng);
    std::uniform_real_distribution<> U(0.0, 1.0);
    const double levy_prob = 0.1 * (1.0 + success_rate_[static_cast<size_t>(i)]);
    if (U(rng) < levy_prob) {
        std::vector<double> levy(static_cast<size_t>(n_));
        for (auto& l : levy) l = generateLevyNumber(rng);
        const double step_scale = 0.01 * (1.0 - stagnation_counter_ / static_cast<double>(max_stagnation_));
        double* p = pos_.data() + static_cast<size_t>(i) * n_;
        for (int k = 0; k < n_; ++k) {
            const size_t kk = static_cast<size_t>(k);
            const double scale = step_scale * (ub_[kk] - lb_[kk]);
            p[k] += scale * levy[kk];
            p[k] = std::clamp(p[k], lb_[kk], ub_[kk]);
        }
    }
}

void ParticleSwarmOptimization::updateParticles(int iter, IObjectiveFunction& f) {   // .cpp:330-425
    double omega = omega_start_, c1 = c1_initial_, c2 = c2_initial_;
    if (use_adaptive_parameters_) adaptParameters(estimateEvolutionaryState(), iter, omega, c1, c2);
    else coefficients(iter, omega, c1, c2);
    std::vector<double> mean_best;
    if (variant_ == PSOVariant::QUANTUM || variant_ == PSOVariant::HYBRID) {          // calculateMeanBestPosition (.cpp:936-947)
        mean_best.assign(static_cast<size_t>(n_), 0.0);
        for (int i = 0; i < local_; ++i)
            for (int k = 0; k < n_; ++k) mean_best[static_cast<size_t>(k)] += pbest_[static_cast<size_t>(i) * n_ + static_cast<size_t>(k)];
        for (auto& m : mean_best) m /= swarm_size_;
    }
    const std::vector<uint32_t> seeds = drawSeeds();
    // getNeighborhoodBest (.cpp:816-834) for every particle from the personal bests at the start of the iteration
    std::vector<int> lbest_of;
    std::vector<double> lbest_rows;
    if (topology_ != TopologyType::GLOBAL_BEST) {
        lbest_of.resize(static_cast<size_t>(local_));
        for (int i = 0; i < local_; ++i) {
            int best = i;
            double best_value = pbest_val_[static_cast<size_t>(i)];
            for (int nb : getNeighbors(i))
                if (nb >= 0 && nb < swarm_size_ && pbest_val_[static_cast<size_t>(nb)] > best_value) { best_value = pbest_val_[static_cast<size_t>(nb)]; best = nb; }
            lbest_of[static_cast<size_t>(i)] = best;
        }
        lbest_rows = pbest_;           // snapshot: an update below never changes pbest_, but keep the read side explicit
    }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < local_; ++i) {
        std::mt19937 local_rng(seeds[static_cast<size_t>(i)]);
        std::uniform_real_distribution<> local_uniform(0.0, 1.0);
        const double* lbest = (topology_ == TopologyType::GLOBAL_BEST) ? gbest_.data()
                                                                       : lbest_rows.data() + static_cast<size_t>(lbest_of[static_cast<size_t>(i)]) * n_;
        switch (variant_) {
            case PSOVariant::STANDARD:
            case PSOVariant::ADAPTIVE:                                   // standard update with the adapted coefficients
                standardPSOUpdate(i, lbest, omega, c1, c2, local_rng);
                break;
            case PSOVariant::QUANTUM:
                quantumPSOUpdate(i, mean_best, iter, local_rng);
                break;
            case PSOVariant::LEVY_FLIGHT:
                levyFlightUpdate(i, omega, c1, c2, local_rng);
                break;
            case PSOVariant::HYBRID: {                                   // by success rate; the uniform is drawn only when its test is reached
                const double sr = success_rate_[static_cast<size_t>(i)];
                if (sr < 0.3 && local_uniform(local_rng) < 0.5) levyFlightUpdate(i, omega, c1, c2, local_rng);
                else if (sr > 0.7 && local_uniform(local_rng) < 0.3) quantumPSOUpdate(i, mean_best, iter, local_rng);
                else standardPSOUpdate(i, lbest, omega, c1, c2, local_rng);
                break;
            }
        }
    }
    evaluateSwarm(f, 0, cur_fit_);
    for (int i = 0; i < local_; ++i) {
        const size_t ii = static_cast<size_t>(i);
        total_updates_[ii]++;
        if (cur_fit_[ii] > pbest_val_[ii]) {
            pbest_val_[ii] = cur_fit_[ii];
            std::copy(pos_.begin() + static_cast<std::ptrdiff_t>(i) * n_, pos_.begin() + static_cast<std::ptrdiff_t>(i + 1) * n_, pbest_.begin() + static_cast<std::ptrdiff_t>(i) * n_);
            success_count_[ii]++;
        }
        success_rate_[ii] = (total_updates_[ii] > 0) ? static_cast<double>(success_count_[ii]) / total_updates_[ii] : 0.0;
    }
}

// applyElitistLearningStrategy (.cpp:705-742): up to three Gaussian trials around the best particle's POSITION with a
// halving radius; the first improvement of its personal best is taken.
void ParticleSwarmOptimization::applyElitistLearningStrategy(int best, IObjectiveFunction& f) {
    const size_t b = static_cast<size_t>(best);
    const double* p = pos_.data() + b * n_;
    double sigma_scale = 0.1 * std::exp(-2.0 * success_rate_[b]);
    std::vector<double> trials(static_cast<size_t>(3) * n_), fit(3);
    std::mt19937 rng_after[3];
    std::normal_distribution<> dist_after[3];
    for (int a = 0; a < 3; ++a) {
        for (int k = 0; k < n_; ++k) {
            const size_t kk = static_cast<size_t>(k);
            const double sigma = sigma_scale * (ub_[kk] - lb_[kk]);
            trials[static_cast<size_t>(a) * n_ + kk] = std::clamp(p[k] + sigma * normal_dist_(rng_), lb_[kk], ub_[kk]);
        }
        rng_after[a] = rng_; dist_after[a] = normal_dist_;
        sigma_scale *= 0.5;
    }
    f.calculateBatch(trials.data(), 3, n_, fit.data());
    evaluations_ += 3;
    for (int a = 0; a < 3; ++a) {
        ++els_trials_;
        if (fit[static_cast<size_t>(a)] > pbest_val_[b]) {
            std::copy(trials.begin() + static_cast<std::ptrdiff_t>(a) * n_, trials.begin() + static_cast<std::ptrdiff_t>(a + 1) * n_, pos_.begin() + static_cast<std::ptrdiff_t>(best) * n_);
            std::copy(trials.begin() + static_cast<std::ptrdiff_t>(a) * n_, trials.begin() + static_cast<std::ptrdiff_t>(a + 1) * n_, pbest_.begin() + static_cast<std::ptrdiff_t>(best) * n_);
            pbest_val_[b] = fit[static_cast<size_t>(a)];
            cur_fit_[b] = fit[static_cast<size_t>(a)];
            rng_ = rng_after[a]; normal_dist_ = dist_after[a];          // the reference stops drawing here
            break;
        }
    }
}

// restartSwarm (.cpp:744-814): keep the best `keep_best_count` particles, re-seed the others around the elites (70 % of the
// coordinates) or uniformly in bounds (30 %)
void ParticleSwarmOptimization::restartSwarm(IObjectiveFunction& f, int keep_best_count) {
    ++restarts_;
    std::vector<int> order(static_cast<size_t>(local_));
    for (int i = 0; i < local_; ++i) order[static_cast<size_t>(i)] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return pbest_val_[static_cast<size_t>(a)] > pbest_val_[static_cast<size_t>(b)]; });
    permuteSwarm(order);
    const int n_elite = std::min(keep_best_count, swarm_size_);
    const std::vector<double> elite(pos_.begin(), pos_.begin() + static_cast<std::ptrdiff_t>(n_elite) * n_);
    const std::vector<uint32_t> seeds = drawSeeds();
#pragma omp parallel for schedule(static)
    for (int i = keep_best_count; i < local_; ++i) {
        std::mt19937 local_rng(seeds[static_cast<size_t>(i)]);
        std::uniform_real_distribution<> U(0.0, 1.0);
        std::normal_distribution<> Nrm(0.0, 1.0);                        // one object per particle: its cached second value carries over
        const double* el = elite.data() + static_cast<size_t>(i % n_elite) * n_;
        double* p = pos_.data() + static_cast<size_t>(i) * n_;
        double* v = vel_.data() + static_cast<size_t>(i) * n_;
        for (int k = 0; k < n_; ++k) {
            const size_t kk = static_cast<size_t>(k);
            if (U(local_rng) < 0.7) {
                const double range = ub_[kk] - lb_[kk];
                const double sigma = 0.3 * range * (1.0 + 0.5 * U(local_rng));
                p[k] = el[k] + sigma * Nrm(local_rng);
            } else {
                p[k] = lb_[kk] + U(local_rng) * (ub_[kk] - lb_[kk]);
            }
            p[k] = std::clamp(p[k], lb_[kk], ub_[kk]);
            const double vmax = 0.2 * (ub_[kk] - lb_[kk]);
            v[k] = -vmax + 2 * vmax * U(local_rng);
        }
    }
    evaluateSwarm(f, keep_best_count, cur_fit_);
    for (int i = keep_best_count; i < local_; ++i) {
        const size_t ii = static_cast<size_t>(i);
        std::copy(pos_.begin() + static_cast<std::ptrdiff_t>(i) * n_, pos_.begin() + static_cast<std::ptrdiff_t>(i + 1) * n_, pbest_.begin() + static_cast<std::ptrdiff_t>(i) * n_);
        pbest_val_[ii] = cur_fit_[ii];
        success_count_[ii] = 0; total_updates_[ii] = 0; success_rate_[ii] = 0.0;
    }
    gbest_value_ = pbest_val_[0];                                        // the sorted swarm's first particle, unconditionally (.cpp:806-807)
    gbest_.assign(pbest_.begin(), pbest_.begin() + n_);
}

void ParticleSwarmOptimization::runHostLoop(int start_iter, double previous_gbest, IObjectiveFunction& f) {   // .cpp:131-182
    for (int iter = start_iter; iter < iterations_; ++iter) {
        if (std::abs(gbest_value_ - previous_gbest) < restart_threshold_) {
            stagnation_counter_++;
            if (stagnation_counter_ > max_stagnation_) {
                restartSwarm(f);
                stagnation_counter_ = 0;
            }
        } else {
            stagnation_counter_ = 0;
        }
        previous_gbest = gbest_value_;
        updateParticles(iter, f);
        rescanGlobalBest();
        if ((variant_ == PSOVariant::ADAPTIVE || variant_ == PSOVariant::HYBRID) && (iter % 5 == 0)) {
            const int best = static_cast<int>(std::max_element(pbest_val_.begin(), pbest_val_.end()) - pbest_val_.begin());   // first maximum
            applyElitistLearningStrategy(best, f);
            if (pbest_val_[static_cast<size_t>(best)] > gbest_value_) {
                gbest_value_ = pbest_val_[static_cast<size_t>(best)];
                gbest_.assign(pbest_.begin() + static_cast<std::ptrdiff_t>(best) * n_, pbest_.begin() + static_cast<std::ptrdiff_t>(best + 1) * n_);
            }
        }
        reportProgress(iter, "host");
    }
}

void ParticleSwarmOptimization::reportProgress(int iter, const char* where) const {
    if (!report_enabled() || !((iter + 1) % report_interval_ == 0 || iter == iterations_ - 1)) return;
    std::fprintf(stderr, "[PSO %s] Iteration %d/%d | Best: %.17g | Stagnation: %d | Restarts: %d\n", where, iter + 1, iterations_, gbest_value_,
                 stagnation_counter_, restarts_);
}

double ParticleSwarmOptimization::swarmDiversity() const {
    if (local_ == 0) return 0.0;
    std::vector<double> centroid(static_cast<size_t>(n_), 0.0);
    for (int i = 0; i < local_; ++i)
        for (int k = 0; k < n_; ++k) centroid[static_cast<size_t>(k)] += pos_[static_cast<size_t>(i) * n_ + static_cast<size_t>(k)];
    for (auto& c : centroid) c /= swarm_size_;
    double avg = 0.0, mx = 0.0;
    for (int i = 0; i < local_; ++i) {
        double d2 = 0.0;
        for (int k = 0; k < n_; ++k) { const double d = pos_[static_cast<size_t>(i) * n_ + static_cast<size_t>(k)] - centroid[static_cast<size_t>(k)]; d2 += d * d; }
        const double dist = std::sqrt(d2);
        avg += dist;
        mx = std::max(mx, dist);
    }
    avg /= swarm_size_;
    return (mx > 0) ? avg / mx : 0.0;
}

OptimizationResult ParticleSwarmOptimization::finish() const {
    OptimizationResult r;
    r.bestParameters = VectorXd::FromPointer(gbest_.data(), n_);
    r.bestObjectiveValue = gbest_value_;
    VectorXd mean;
    r.finalCovariance = personalBestScatter(mean);                    // covariance of the personal bests for phase 2 (.cpp:226-242)
    r.finalCovariance *= 1.0 / static_cast<double>(std::max(local_ - 1, 1));
    r.finalCovariance += 1e-6 * MatrixXd::Identity(n_, n_);
    return r;
}

OptimizationResult ParticleSwarmOptimization::optimize(const VectorXd& initial, IObjectiveFunction& f, IParameterManager& pm) {
    const int n = static_cast<int>(pm.getParameterCount());
    const VectorXd* init = initial.size() == n ? &initial : nullptr;
    const int local = local_count_setting_ >= 0 ? local_count_setting_ : swarm_size_ - static_cast<int>(particle_offset_);
    const bool whole = particle_offset_ == 0 && local == swarm_size_;
    auto* device_objective = dynamic_cast<SEPAIHRDObjectiveFunction*>(&f);
    if (!whole) {
        // one shard of a swarm run on its own (the sharded drivers exchange the global best between the calls of the
        // step-wise form instead): STANDARD / GLOBAL_BEST only, no restart
        begin(init, pm);
        auto evaluate = [&]() {
            std::vector<double> fit(static_cast<size_t>(local_));
            f.calculateBatch(pos_.data(), local_, n_, fit.data());
            const auto best = tell(fit.data());
            if (best.second >= 0) setGlobalBest(best.first, personalBest(best.second));
        };
        evaluate();
        for (int iter = 0; iter < iterations_; ++iter) { step(iter); evaluate(); }
        return finish();
    }
    if (device_objective != nullptr && device_resident_ && isBasicSwarm()) {
        // the swarm stays in HBM: per iteration one seed per particle goes down, one (value, index, position) triple comes
        // back.  The stagnation test of the main loop (.cpp:133-146) runs here on the host; when it asks for a restart the
        // swarm is read back once and the run continues in the host-resident engine.
        beginDevice(init, pm, device_objective->device().get());
        std::vector<double> best_pos(static_cast<size_t>(n));
        auto evaluate = [&]() {
            const auto best = evaluateDevice(best_pos.data());
            if (best.second >= 0) setGlobalBest(best.first, best_pos.data());
        };
        evaluate();
        double previous_gbest = -std::numeric_limits<double>::infinity();
        for (int iter = 0; iter < iterations_; ++iter) {
            const bool stagnant = std::abs(gbest_value_ - previous_gbest) < restart_threshold_;
            if (stagnant && stagnation_counter_ + 1 > max_stagnation_) {
                fetchPersonalBests(true);
                cur_fit_ = pbest_val_;             // current fitness is only read by the adaptive variants, which never start on the device
                sepaihrd_swarm_destroy(dev_swarm_); dev_swarm_ = nullptr;
                runHostLoop(iter, previous_gbest, f);
                return finish();
            }
            stagnation_counter_ = stagnant ? stagnation_counter_ + 1 : 0;
            previous_gbest = gbest_value_;
            stepDevice(iter);
            evaluate();
            reportProgress(iter, "device");
        }
        fetchPersonalBests(false);                 // the personal bests feed the covariance hand-off; positions stay on the device
        return finish();
    }
    initializeSwarmFull(init, f, pm);
    runHostLoop(0, -std::numeric_limits<double>::infinity(), f);
    return finish();
}

// =====================================================================================================================
// NUTS (src/model/optimizers/NUTSSampler.cpp)
// =====================================================================================================================
namespace {
double dot(const VectorXd& a, const VectorXd& b) {
    double s = 0.0;
    for (std::ptrdiff_t i = 0; i < a.size(); ++i) s += a(i) * b(i);
    return s;
}
void clip_gradient(VectorXd& grad) {                                  // .cpp:80-88, 285-289, 296-299
    const double norm = std::sqrt(dot(grad, grad));
    if (norm > 1000.0) grad *= 1000.0 / norm;
}
}  // namespace

NUTSSampler::NUTSSampler() : rng_(std::random_device{}()) {}

void NUTSSampler::configure(const std::map<std::string, double>& s) {
    num_iterations_ = static_cast<int>(setting(s, "nuts_iterations", 2000.0));
    adaptation_window_ = static_cast<int>(setting(s, "nuts_adaptation_window", 500.0));
    delta_target_ = setting(s, "nuts_delta_target", 0.8);
    max_tree_depth_ = static_cast<int>(setting(s, "nuts_max_tree_depth", 10.0));
    has_seed_ = s.count("seed") != 0;
    seed_ = static_cast<unsigned>(setting(s, "seed", 0.0));
}

double NUTSSampler::gradientAt(IGradientObjectiveFunction& objective, const VectorXd& theta, VectorXd& grad) const {
    if (memo_valid_ && memo_theta_ == theta) { grad = memo_grad_; return memo_value_; }
    const double v = objective.evaluate_with_gradient(theta, grad);
    ++gradient_evaluations_;
    memo_theta_ = theta; memo_grad_ = grad; memo_value_ = v; memo_valid_ = true;
    return v;
}

OptimizationResult NUTSSampler::optimize(const VectorXd& initialParameters, IObjectiveFunction& objectiveFunction, IParameterManager& pm) {
    auto* grad_obj = dynamic_cast<IGradientObjectiveFunction*>(&objectiveFunction);
    if (!grad_obj) throw InvalidParameterException("NUTSSampler", "Objective function must implement IGradientObjectiveFunction for NUTS.");
    if (has_seed_) rng_.seed(seed_);
    memo_valid_ = false; gradient_evaluations_ = 0; tree_depths_.clear();
    OptimizationResult result;
    result.samples.reserve(static_cast<size_t>(std::max(num_iterations_, 0)));
    result.sampleObjectiveValues.reserve(static_cast<size_t>(std::max(num_iterations_, 0)));
    VectorXd theta_m = initialParameters;
    const std::ptrdiff_t P = theta_m.size();

    double epsilon = findReasonableEpsilon(*grad_obj, theta_m, pm);
    // dual averaging (.cpp:63-70)
    const double mu = std::log(10.0 * epsilon);
    double epsilon_bar = epsilon, H_bar = 0.0;
    const double gamma = 0.05, t0 = 10.0, kappa = 0.75;

    for (int m = 1; m <= num_iterations_; ++m) {
        std::normal_distribution<> normal(0.0, 1.0);                     // a fresh distribution per iteration (.cpp:74)
        VectorXd r0(P);
        for (std::ptrdiff_t i = 0; i < P; ++i) r0(i) = normal(rng_);
        VectorXd grad(P);
        const double log_p = gradientAt(*grad_obj, theta_m, grad);
        clip_gradient(grad);
        if (!std::isfinite(log_p)) {                                     // .cpp:98-105: repeat the last sample, if there is one
            if (!result.samples.empty()) {
                result.samples.push_back(result.samples.back());
                result.sampleObjectiveValues.push_back(result.sampleObjectiveValues.back());
            }
            continue;
        }
        const double H0 = log_p - 0.5 * dot(r0, r0);
        const double log_u_slice = H0 - std::exponential_distribution<>(1.0)(rng_);
        VectorXd theta_minus = theta_m, theta_plus = theta_m, r_minus = r0, r_plus = r0, theta_next = theta_m;
        int j = 0, n = 1, n_alpha = 0;
        bool s = true;
        double alpha = 0.0;
        while (s && j < max_tree_depth_) {
            const int v = (std::uniform_int_distribution<>(0, 1)(rng_) * 2) - 1;
            Tree subtree;
            if (v == -1) {
                buildTree(*grad_obj, theta_minus, r_minus, log_u_slice, v, j, epsilon, H0, pm, subtree);
                theta_minus = subtree.theta_minus; r_minus = subtree.r_minus;
            } else {
                buildTree(*grad_obj, theta_plus, r_plus, log_u_slice, v, j, epsilon, H0, pm, subtree);
                theta_plus = subtree.theta_plus; r_plus = subtree.r_plus;
            }
            if (subtree.s && checkNoUTurn(theta_minus, theta_plus, r_minus, r_plus)) {
                const double acceptance_prob = static_cast<double>(subtree.n_valid) / static_cast<double>(n + subtree.n_valid);
                if (std::uniform_real_distribution<>(0.0, 1.0)(rng_) < acceptance_prob) theta_next = subtree.theta_prime;
                n += subtree.n_valid;
                alpha += subtree.alpha;
                n_alpha += subtree.n_alpha;
                j++;
            } else {
                s = false;
            }
        }
        theta_m = theta_next;
        tree_depths_.push_back(j);
        if (m <= adaptation_window_) {                                   // .cpp:165-179
            const double avg_alpha = (n_alpha > 0) ? alpha / n_alpha : 0.0;
            const double eta = 1.0 / (m + t0);
            H_bar = (1.0 - eta) * H_bar + eta * (delta_target_ - avg_alpha);
            const double log_epsilon = mu - (std::sqrt(m) / gamma) * H_bar;
            epsilon = std::exp(log_epsilon);
            const double m_kappa = std::pow(m, -kappa);
            const double log_epsilon_bar = m_kappa * log_epsilon + (1.0 - m_kappa) * std::log(epsilon_bar);
            epsilon_bar = std::exp(log_epsilon_bar);
        } else {
            epsilon = epsilon_bar;
        }
        const VectorXd constrained_theta = pm.applyConstraints(theta_m);
        result.samples.push_back(constrained_theta);
        const double final_obj = objectiveFunction.calculate(constrained_theta);
        result.sampleObjectiveValues.push_back(final_obj);
        if (final_obj > result.bestObjectiveValue) {
            result.bestObjectiveValue = final_obj;
            result.bestParameters = constrained_theta;
        }
    }
    final_epsilon_ = epsilon;
    return result;
}

// heuristic first step size (.cpp:213-279): a tenth of the mean proposal sigma, clipped to [1e-6, 0.1], then at most five
// halvings / 1.5-fold increases steered by the acceptance probability of one leapfrog step
double NUTSSampler::findReasonableEpsilon(IGradientObjectiveFunction& objective, const VectorXd& theta, IParameterManager& pm) const {
    const std::ptrdiff_t P = theta.size();
    double avg_scale = 0.0;
    for (std::ptrdiff_t i = 0; i < P; ++i) avg_scale += pm.getSigmaForParamIndex(static_cast<int>(i));
    avg_scale /= static_cast<double>(P);
    double epsilon = std::max(1e-6, std::min(avg_scale * 0.1, 0.1));
    std::normal_distribution<> normal(0.0, 1.0);
    VectorXd r(P);
    for (std::ptrdiff_t i = 0; i < P; ++i) r(i) = normal(rng_);
    VectorXd grad(P);
    const double log_p = gradientAt(objective, theta, grad);
    if (!std::isfinite(log_p)) return epsilon;
    const double H0 = log_p - 0.5 * dot(r, r);
    VectorXd theta_prime = theta, r_prime = r;
    leapfrog(objective, theta_prime, r_prime, epsilon, pm);
    double log_p_prime = gradientAt(objective, theta_prime, grad);
    double H_prime = log_p_prime - 0.5 * dot(r_prime, r_prime);
    double accept_prob = std::exp(std::min(0.0, H_prime - H0));
    for (int iter = 0; iter < 5; ++iter) {
        if (accept_prob < 0.1 && epsilon > 1e-8) epsilon *= 0.5;
        else if (accept_prob > 0.9 && epsilon < 1.0) epsilon *= 1.5;
        else break;
        theta_prime = theta; r_prime = r;
        leapfrog(objective, theta_prime, r_prime, epsilon, pm);
        log_p_prime = gradientAt(objective, theta_prime, grad);
        if (!std::isfinite(log_p_prime)) { epsilon *= 0.5; continue; }
        H_prime = log_p_prime - 0.5 * dot(r_prime, r_prime);
        accept_prob = std::exp(std::min(0.0, H_prime - H0));
    }
    return epsilon;
}

void NUTSSampler::leapfrog(IGradientObjectiveFunction& objective, VectorXd& theta, VectorXd& r, double epsilon, IParameterManager& pm) const {   // .cpp:282-306
    VectorXd grad(theta.size());
    gradientAt(objective, theta, grad);
    clip_gradient(grad);
    r += (0.5 * epsilon) * grad;
    theta += epsilon * r;
    theta = pm.applyConstraints(theta);
    gradientAt(objective, theta, grad);
    clip_gradient(grad);
    r += (0.5 * epsilon) * grad;
}

void NUTSSampler::buildTree(IGradientObjectiveFunction& objective, const VectorXd& theta, const VectorXd& r, double log_u_slice, int v, int j,
                            double epsilon, double H0, IParameterManager& pm, Tree& tree) const {                                          // .cpp:309-407
    if (j == 0) {
        VectorXd theta_prime = theta, r_prime = r;
        leapfrog(objective, theta_prime, r_prime, v * epsilon, pm);
        VectorXd grad(theta_prime.size());
        const double log_p = gradientAt(objective, theta_prime, grad);
        const double H_prime = log_p - 0.5 * dot(r_prime, r_prime);
        tree.n_valid = (log_u_slice <= H_prime) ? 1 : 0;
        tree.s = (log_u_slice < H_prime + DELTA_MAX);
        tree.theta_minus = theta_prime; tree.theta_plus = theta_prime;
        tree.r_minus = r_prime; tree.r_plus = r_prime;
        tree.theta_prime = theta_prime;
        tree.alpha = std::min(1.0, std::exp(H_prime - H0));
        tree.n_alpha = 1;
        return;
    }
    Tree left;
    buildTree(objective, theta, r, log_u_slice, v, j - 1, epsilon, H0, pm, left);
    if (!left.s) { tree = left; return; }
    Tree right;
    if (v == -1) {
        buildTree(objective, left.theta_minus, left.r_minus, log_u_slice, v, j - 1, epsilon, H0, pm, right);
        tree.theta_minus = right.theta_minus; tree.r_minus = right.r_minus;
        tree.theta_plus = left.theta_plus; tree.r_plus = left.r_plus;
    } else {
        buildTree(objective, left.theta_plus, left.r_plus, log_u_slice, v, j - 1, epsilon, H0, pm, right);
        tree.theta_minus = left.theta_minus; tree.r_minus = left.r_minus;
        tree.theta_plus = right.theta_plus; tree.r_plus = right.r_plus;
    }
    if (right.s) {
        tree.n_valid = left.n_valid + right.n_valid;
        const double prob = (tree.n_valid > 0) ? static_cast<double>(right.n_valid) / static_cast<double>(tree.n_valid) : 0.0;
        tree.theta_prime = (std::uniform_real_distribution<>(0.0, 1.0)(rng_) < prob) ? right.theta_prime : left.theta_prime;
        tree.alpha = left.alpha + right.alpha;
        tree.n_alpha = left.n_alpha + right.n_alpha;
        tree.s = left.s && right.s && checkNoUTurn(tree.theta_minus, tree.theta_plus, tree.r_minus, tree.r_plus);
    } else {
        tree.theta_prime = left.theta_prime;
        tree.n_valid = left.n_valid;
        tree.s = false;
        tree.alpha = left.alpha;
        tree.n_alpha = left.n_alpha;
    }
}

bool NUTSSampler::checkNoUTurn(const VectorXd& theta_minus, const VectorXd& theta_plus, const VectorXd& r_minus, const VectorXd& r_plus) {   // .cpp:410-424
    const VectorXd delta = theta_plus - theta_minus;
    return dot(delta, r_minus) >= 0 && dot(delta, r_plus) >= 0;
}

// =====================================================================================================================
// Hill climbing: candidate cloud + robust line search
// =====================================================================================================================
void HillClimbingOptimizer::configure(const std::map<std::string, double>& s) {
    iterations_ = static_cast<int>(setting(s, "iterations", 2000.0));
    report_interval_ = std::max(1, static_cast<int>(setting(s, "report_interval", 100.0)));
    cloud_size_multiplier_ = std::max(1, static_cast<int>(setting(s, "cloud_size_multiplier", 8.0)));
    // the reference sizes the cloud as max(4, omp threads * multiplier) (.cpp:170-175); on the GPU the natural unit is
    // a wave of lane groups, so an explicit "cloud_size" can be given (default: 8 "threads" worth)
    cloud_size_ = static_cast<int>(setting(s, "cloud_size", 0.0));
    has_seed_ = s.count("seed") != 0;
    seed_ = static_cast<unsigned>(setting(s, "seed", 0.0));
}

bool HillClimbingOptimizer::lineSearch(VectorXd& current, double& current_logL, const VectorXd& direction, IObjectiveFunction& func,
                                       IParameterManager& pm) const {
    const int max_backtrack = 10, max_expansion = 12;
    const int64_t P = current.size();
    auto sqdist = [](const VectorXd& a, const VectorXd& b) { double s = 0; for (std::ptrdiff_t i = 0; i < a.size(); ++i) s += (a(i) - b(i)) * (a(i) - b(i)); return s; };
    // backtracking: candidates at step 1, 1/2, 1/4, ... -- evaluated as ONE batch, consumed in order
    std::vector<VectorXd> cand;
    double step = 1.0;
    for (int i = 0; i < max_backtrack; ++i) {
        VectorXd c = pm.applyConstraints(current + direction * step);
        if (sqdist(c, current) < 1e-16) break;
        cand.push_back(c);
        step *= 0.5;
    }
    if (cand.empty()) return false;
    std::vector<double> rows(cand.size() * static_cast<size_t>(P));
    for (size_t i = 0; i < cand.size(); ++i) std::copy(cand[i].data(), cand[i].data() + P, rows.begin() + static_cast<std::ptrdiff_t>(i * static_cast<size_t>(P)));
    std::vector<double> val = evaluate_rows(func, rows, static_cast<int64_t>(cand.size()), P);
    int foothold = -1;
    for (size_t i = 0; i < cand.size(); ++i)
        if (val[i] > current_logL) { foothold = static_cast<int>(i); break; }
    if (foothold < 0) return false;
    VectorXd best = cand[static_cast<size_t>(foothold)];
    double best_logL = val[static_cast<size_t>(foothold)];
    // expansion: c_k = constrain(c_{k-1} + 2^k s) as long as every previous candidate improved
    VectorXd cur_step = best - current;
    std::vector<VectorXd> exp_c;
    VectorXd base = best;
    for (int i = 0; i < max_expansion; ++i) {
        cur_step *= 2.0;
        VectorXd c = pm.applyConstraints(base + cur_step);
        exp_c.push_back(c);
        base = c;
    }
    rows.assign(exp_c.size() * static_cast<size_t>(P), 0.0);
    for (size_t i = 0; i < exp_c.size(); ++i) std::copy(exp_c[i].data(), exp_c[i].data() + P, rows.begin() + static_cast<std::ptrdiff_t>(i * static_cast<size_t>(P)));
    val = evaluate_rows(func, rows, static_cast<int64_t>(exp_c.size()), P);
    for (size_t i = 0; i < exp_c.size(); ++i) {
        if (val[i] > best_logL) { best = exp_c[i]; best_logL = val[i]; }
        else break;
    }
    current = best;
    current_logL = best_logL;
    return true;
}

OptimizationResult HillClimbingOptimizer::optimize(const VectorXd& initial, IObjectiveFunction& f, IParameterManager& pm) {
    OptimizationResult result;
    const auto P = initial.size();
    const int n_params = static_cast<int>(P);
    result.bestParameters = initial;
    result.bestObjectiveValue = MetropolisHastingsSampler::safeValue(f.calculate(initial));
    VectorXd current = initial, prev = initial;
    double current_logL = result.bestObjectiveValue;
    MatrixXd cov = MatrixXd::Identity(P, P);
    for (std::ptrdiff_t i = 0; i < P; ++i) { const double s = pm.getSigmaForParamIndex(static_cast<int>(i)); cov(i, i) = (s > 0 ? s * s : 1e-4); }
    MatrixXd L;
    linalg::cholesky_lower(cov, L);
    const int num_candidates = cloud_size_ > 0 ? std::max(4, cloud_size_) : std::max(4, 8 * cloud_size_multiplier_);
    std::mt19937 gen(has_seed_ ? seed_ : std::random_device{}());
    std::normal_distribution<double> norm(0.0, 1.0);
    std::vector<double> rows(static_cast<size_t>(num_candidates) * static_cast<size_t>(P));
    for (int iter = 0; iter < iterations_; ++iter) {
        // half the cloud follows the learned covariance, half moves along one coordinate (.cpp:198-223)
        for (int i = 0; i < num_candidates; ++i) {
            VectorXd d = VectorXd::Zero(P);
            if (i < num_candidates / 2) {
                VectorXd z(P);
                for (std::ptrdiff_t k = 0; k < P; ++k) z(k) = norm(gen);
                d = L * z;
            } else {
                std::uniform_int_distribution<int> pick(0, n_params - 1);
                const int idx = pick(gen);
                d(idx) = std::sqrt(cov(idx, idx)) * norm(gen);
            }
            const VectorXd c = pm.applyConstraints(current + d);
            std::copy(c.data(), c.data() + P, rows.begin() + static_cast<std::ptrdiff_t>(i) * P);
        }
        const std::vector<double> scores = evaluate_rows(f, rows, num_candidates, P);
        int best_idx = -1;
        double best_val = -1e18;
        for (int i = 0; i < num_candidates; ++i)
            if (scores[static_cast<size_t>(i)] > best_val) { best_val = scores[static_cast<size_t>(i)]; best_idx = i; }
        bool moved = false;
        if (best_idx != -1 && best_val > -1e18) {
            const VectorXd best_point = VectorXd::FromPointer(rows.data() + static_cast<std::ptrdiff_t>(best_idx) * P, P);
            const VectorXd direction = best_point - current;
            if (best_val > current_logL) { current = best_point; current_logL = best_val; moved = true; }
            moved = lineSearch(current, current_logL, direction, f, pm) || moved;
        }
        if (moved) {
            if (current_logL > result.bestObjectiveValue) { result.bestObjectiveValue = current_logL; result.bestParameters = current; }
            const VectorXd stepv = current - prev;
            double sq = 0;
            for (std::ptrdiff_t i = 0; i < P; ++i) sq += stepv(i) * stepv(i);
            if (sq > 1e-14) {
                const double alpha = 2.0 / (n_params + 2.0);
                cov = (1.0 - alpha) * cov + alpha * outer(stepv);
                cov = 0.5 * (cov + cov.transpose());
                cov += (1e-8 * cov.trace() / n_params) * MatrixXd::Identity(P, P);
                for (std::ptrdiff_t i = 0; i < P; ++i) {
                    double mv = pm.getSigmaForParamIndex(static_cast<int>(i));
                    mv = (mv > 0 ? mv * mv * 0.01 : 1e-8);
                    if (cov(i, i) < mv) cov(i, i) = mv;
                }
            }
            prev = current;
        }
        if (iter > 0 && iter % 10 == 0) {
            MatrixXd Ln;
            if (linalg::cholesky_lower(cov, Ln)) {
                L = Ln;
            } else {
                double lambda = 1e-6 * cov.trace() / n_params;
                bool ok = false;
                for (int attempt = 0; attempt < 5 && !ok; ++attempt) {
                    cov += lambda * MatrixXd::Identity(P, P);
                    ok = linalg::cholesky_lower(cov, Ln);
                    lambda *= 10.0;
                }
                if (ok) L = Ln;
                else {
                    MatrixXd dg = MatrixXd::Zero(P, P);
                    L = MatrixXd::Zero(P, P);
                    for (std::ptrdiff_t i = 0; i < P; ++i) { dg(i, i) = cov(i, i); L(i, i) = std::sqrt(cov(i, i)); }
                    cov = dg;
                }
            }
        }
    }
    result.finalCovariance = cov;
    return result;
}

// =====================================================================================================================
// ModelCalibrator / SEPAIHRDModelCalibration
// =====================================================================================================================
ModelCalibrator::ModelCalibrator(std::unique_ptr<IParameterManager> pm, std::unique_ptr<IObjectiveFunction> f,
                                 std::map<std::string, std::unique_ptr<IOptimizationAlgorithm>> algorithms)
    : parameterManager_(std::move(pm)), objectiveFunction_(std::move(f)), optimization_algorithms_(std::move(algorithms)) {
    const char* src = "ModelCalibrator";
    if (!parameterManager_) throw InvalidParameterException(src, "Parameter manager cannot be null.");
    if (!objectiveFunction_) throw InvalidParameterException(src, "Objective function cannot be null.");
    if (optimization_algorithms_.empty()) throw InvalidParameterException(src, "At least one optimization algorithm must be provided.");
    if (parameterManager_->getParameterNames() != objectiveFunction_->getParameterNames())
        throw InvalidParameterException(src, "Parameter names mismatch between ParameterManager and ObjectiveFunction.");
    best_params_vector_ = parameterManager_->getCurrentParameters();
    best_objective_value_ = objectiveFunction_->calculate(best_params_vector_);
    if (std::isnan(best_objective_value_) || std::isinf(best_objective_value_)) best_objective_value_ = -std::numeric_limits<double>::infinity();
}

void ModelCalibrator::calibrate(const std::map<std::string, double>& phase1_settings, const std::map<std::string, double>& phase2_settings) {
    VectorXd current_best = best_params_vector_;
    auto* spm = dynamic_cast<SEPAIHRDParameterManager*>(parameterManager_.get());
    auto it1 = optimization_algorithms_.find(PHASE1_NAME);
    if (it1 != optimization_algorithms_.end()) {
        if (spm) spm->setConstraintMode(ConstraintMode::OPTIMIZATION_CLAMP);
        it1->second->configure(phase1_settings);
        phase1_result_ = it1->second->optimize(current_best, *objectiveFunction_, *parameterManager_);
        if (phase1_result_.bestObjectiveValue > best_objective_value_) {
            best_objective_value_ = phase1_result_.bestObjectiveValue;
            best_params_vector_ = phase1_result_.bestParameters;
        }
        current_best = best_params_vector_;
    }
    auto it2 = optimization_algorithms_.find(PHASE2_NAME);
    if (it2 != optimization_algorithms_.end()) {
        if (spm) spm->setConstraintMode(ConstraintMode::MCMC_REFLECT);
        it2->second->configure(phase2_settings);
        if (phase1_result_.finalCovariance.size() > 0) {
            if (auto* mh = dynamic_cast<MetropolisHastingsSampler*>(it2->second.get())) {
                // condition the phase-1 covariance: symmetrise, floor the eigenvalues at (0.1 sigma)^2, inflate x4, ridge
                MatrixXd cov = phase1_result_.finalCovariance;
                const auto n = cov.rows();
                cov = 0.5 * (cov + cov.transpose());
                VectorXd evals; MatrixXd evecs;
                linalg::symmetric_eigen(cov, evals, evecs);
                // ascending order like Eigen::SelfAdjointEigenSolver, so that floor i pairs with parameter i's sigma as in the reference
                std::vector<std::ptrdiff_t> order(static_cast<size_t>(n));
                for (std::ptrdiff_t i = 0; i < n; ++i) order[static_cast<size_t>(i)] = i;
                std::sort(order.begin(), order.end(), [&](std::ptrdiff_t a, std::ptrdiff_t b) { return evals(a) < evals(b); });
                MatrixXd floored = MatrixXd::Zero(n, n);
                for (std::ptrdiff_t r = 0; r < n; ++r) {
                    const std::ptrdiff_t e = order[static_cast<size_t>(r)];
                    const double prior = parameterManager_->getSigmaForParamIndex(static_cast<int>(r));
                    const double lam = std::max(evals(e), std::pow(prior * 0.1, 2));
                    for (std::ptrdiff_t j = 0; j < n; ++j)
                        for (std::ptrdiff_t i = 0; i < n; ++i) floored(i, j) += lam * evecs(i, e) * evecs(j, e);
                }
                MatrixXd phase2 = floored * 4.0;
                phase2 += (1e-8 * phase2.trace() / static_cast<double>(n)) * MatrixXd::Identity(n, n);
                mh->setInitialCovariance(phase2);
            }
        }
        phase2_result_ = it2->second->optimize(current_best, *objectiveFunction_, *parameterManager_);
        if (phase2_result_.bestObjectiveValue > best_objective_value_) {
            best_objective_value_ = phase2_result_.bestObjectiveValue;
            best_params_vector_ = phase2_result_.bestParameters;
        }
        // re-score every stored sample (.cpp:144-147): one batch instead of one calculate() per sample
        const auto& S = phase2_result_.samples;
        if (!S.empty()) {
            const int64_t P = S.front().size();
            std::vector<double> rows(S.size() * static_cast<size_t>(P));
            for (size_t i = 0; i < S.size(); ++i) std::copy(S[i].data(), S[i].data() + P, rows.begin() + static_cast<std::ptrdiff_t>(i * static_cast<size_t>(P)));
            mcmcObjectiveValues_.assign(S.size(), 0.0);
            objectiveFunction_->calculateBatch(rows.data(), static_cast<int64_t>(S.size()), P, mcmcObjectiveValues_.data());
        }
    }
    parameterManager_->updateModelParameters(best_params_vector_);
}

SEPAIHRDModelCalibration::SEPAIHRDModelCalibration(std::shared_ptr<AgeSEPAIHRDModel> model, const CalibrationData& data,
                                                   const std::vector<double>& time_points, const std::vector<std::string>& names,
                                                   const std::map<std::string, double>& sigmas,
                                                   const std::map<std::string, std::pair<double, double>>& bounds,
                                                   std::shared_ptr<IOdeSolverStrategy> solver, std::shared_ptr<ISimulationCache> cache)
    : model_(std::move(model)), observed_data_(data), time_points_(time_points), params_to_calibrate_(names), proposal_sigmas_(sigmas),
      param_bounds_(bounds), solver_strategy_(std::move(solver)), cache_(std::move(cache)) {
    const char* src = "SEPAIHRDModelCalibration";
    if (!model_) throw InvalidParameterException(src, "Model pointer is null.");
    if (!solver_strategy_) throw InvalidParameterException(src, "Solver strategy pointer is null.");
    if (!cache_) throw InvalidParameterException(src, "Cache pointer is null.");
    if (time_points_.empty()) throw InvalidParameterException(src, "Time points vector is empty.");
    if (params_to_calibrate_.empty()) throw InvalidParameterException(src, "Parameters to calibrate list is empty.");
    initial_state_cached_ = observed_data_.getInitialSEPAIHRDState();
    if (initial_state_cached_.size() != model_->getStateSize())
        throw ModelConstructionException(src, "Failed to get initial state from calibration data: size mismatch.");
    try {
        parameter_manager_ = std::make_unique<SEPAIHRDParameterManager>(model_, params_to_calibrate_, proposal_sigmas_, param_bounds_);
    } catch (const std::exception& e) {
        throw ModelConstructionException(src, std::string("Failed to create parameter manager: ") + e.what());
    }
}

SEPAIHRDParameterManager& SEPAIHRDModelCalibration::getParameterManager() { return *parameter_manager_; }
VectorXd SEPAIHRDModelCalibration::getCurrentParameterValues() { return parameter_manager_->getCurrentParameters(); }

ModelCalibrator SEPAIHRDModelCalibration::setupCalibrator(std::map<std::string, std::unique_ptr<IOptimizationAlgorithm>> algorithms) {
    const char* src = "SEPAIHRDModelCalibration::setupCalibrator";
    std::unique_ptr<SEPAIHRDParameterManager> pm;
    std::unique_ptr<IObjectiveFunction> f;
    try {
        pm = std::make_unique<SEPAIHRDParameterManager>(model_, params_to_calibrate_, proposal_sigmas_, param_bounds_);
        bool needs_gradient = false;                              // .cpp:77-83: a NUTS sampler among the algorithms asks for the gradient objective
        for (const auto& kv : algorithms) needs_gradient = needs_gradient || dynamic_cast<NUTSSampler*>(kv.second.get()) != nullptr;
        if (needs_gradient)
            f = std::make_unique<SEPAIHRDGradientObjectiveFunction>(model_, *pm, *cache_, observed_data_, time_points_, initial_state_cached_, solver_strategy_);
        else
            f = std::make_unique<SEPAIHRDObjectiveFunction>(model_, *pm, *cache_, observed_data_, time_points_, initial_state_cached_, solver_strategy_);
    } catch (const std::exception& e) {
        throw ModelConstructionException(src, std::string("Failed to create ObjectiveFunction: ") + e.what());
    }
    return ModelCalibrator(std::move(pm), std::move(f), std::move(algorithms));
}

ModelCalibrator SEPAIHRDModelCalibration::runPSOMCMC(const std::map<std::string, double>& s1, const std::map<std::string, double>& s2) {
    std::map<std::string, std::unique_ptr<IOptimizationAlgorithm>> algos;
    algos[ModelCalibrator::PHASE1_NAME] = std::make_unique<ParticleSwarmOptimization>();
    algos[ModelCalibrator::PHASE2_NAME] = std::make_unique<MetropolisHastingsSampler>();
    ModelCalibrator c = setupCalibrator(std::move(algos));
    c.calibrate(s1, s2);
    return c;
}

ModelCalibrator SEPAIHRDModelCalibration::runHillClimbingMCMC(const std::map<std::string, double>& s1, const std::map<std::string, double>& s2) {
    std::map<std::string, std::unique_ptr<IOptimizationAlgorithm>> algos;
    algos[ModelCalibrator::PHASE1_NAME] = std::make_unique<HillClimbingOptimizer>();
    algos[ModelCalibrator::PHASE2_NAME] = std::make_unique<MetropolisHastingsSampler>();
    ModelCalibrator c = setupCalibrator(std::move(algos));
    c.calibrate(s1, s2);
    return c;
}

ModelCalibrator SEPAIHRDModelCalibration::runNUTS(const std::map<std::string, double>& nuts_settings) {
    std::map<std::string, std::unique_ptr<IOptimizationAlgorithm>> algos;
    algos[ModelCalibrator::PHASE2_NAME] = std::make_unique<NUTSSampler>();
    ModelCalibrator c = setupCalibrator(std::move(algos));
    c.calibrate({}, nuts_settings);                                   // no phase 1
    return c;
}

}  // namespace epidemic
