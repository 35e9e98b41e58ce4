// sepaihrd_oracle.cpp -- CPU oracle (TEST INFRASTRUCTURE ONLY; see sepaihrd_oracle.h for the
// scope statement and the parity status: the Dopri5 controller is "parity unpinned").
//
// Build: g++ -std=c++17 -O2 -ffp-contract=off -fopenmp -shared -fPIC  (oracle/Makefile)
// Every function cites the reference file:line it restates.  Paths are relative to the
// reference repository root.

#include "sepaihrd_oracle.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <random>
#include <string>
#include <vector>

#if defined(_OPENMP)
#include <omp.h>
#endif

namespace {

constexpr int NC = SEPAIHRD_NUM_COMPARTMENTS;   // ModelConstants.hpp:18
constexpr int NPOP = 9;                          // ModelConstants.hpp:22 (S..D, excludes CumH/CumICU)
constexpr int MAXN = SEPAIHRD_MAX_AGES;
constexpr int MAXS = NC * MAXN;

// ---- slot layout (sepaihrd_b200.h) --------------------------------------------------------------
struct Slots {
    int n, nb, nk;
    int beta0() const { return 0; }
    int kappa0() const { return nb; }
    int scal0() const { return nb + nk; }              // theta sigma gamma_p gamma_A gamma_I gamma_H gamma_ICU
    int theta() const { return scal0() + 0; }
    int sigma() const { return scal0() + 1; }
    int gamma_p() const { return scal0() + 2; }
    int gamma_A() const { return scal0() + 3; }
    int gamma_I() const { return scal0() + 4; }
    int gamma_H() const { return scal0() + 5; }
    int gamma_ICU() const { return scal0() + 6; }
    int age0() const { return scal0() + 7; }
    int a(int i) const { return age0() + 0 * n + i; }
    int h_infec(int i) const { return age0() + 1 * n + i; }
    int p(int i) const { return age0() + 2 * n + i; }
    int h(int i) const { return age0() + 3 * n + i; }
    int icu(int i) const { return age0() + 4 * n + i; }
    int d_H(int i) const { return age0() + 5 * n + i; }
    int d_ICU(int i) const { return age0() + 6 * n + i; }
    int d_comm(int i) const { return age0() + 7 * n + i; }
    int mult0() const { return age0() + 8 * n; }       // E0 P0 A0 I0 H0 ICU0 R0 D0
    int seed_exposed() const { return mult0() + 8; }
    int runup_days() const { return mult0() + 9; }
    int beta_scalar() const { return mult0() + 10; }
    int count() const { return mult0() + 11; }
};

inline Slots slots_of(const sepaihrd_problem* pb) { return Slots{pb->n_ages, pb->n_beta, pb->n_kappa}; }

bool starts_with(const std::string& s, const char* prefix) { return s.rfind(prefix, 0) == 0; }

// parse "<prefix><unsigned>" like std::stoul(name.substr(len)); returns false on junk.
bool parse_index(const std::string& s, size_t off, unsigned long& out) {
    if (off >= s.size()) return false;
    char* end = nullptr;
    const char* b = s.c_str() + off;
    if (*b < '0' || *b > '9') return false;
    out = std::strtoul(b, &end, 10);
    return end != b;
}

// ---- schedules ------------------------------------------------------------------------------------
// PiecewiseConstantParameterStrategy::getValue (src/model/PiecewiseConstantParameterStrategy.cpp:37-74)
// and PiecewiseConstantNpiStrategy::getReductionFactor (src/model/PieceWiseConstantNPIStrategy.cpp:86-127).
// Both: t <= end[0] -> v[0]; else the first k >= 1 with t <= end[k] -> v[k]; past the last -> v[m-1].
// (The mutable cached index in the reference is an optimisation: its forward scan and its
// lower_bound branch select the same element.)  The NPI variant also maps t < 0 to the baseline,
// which is implied because its baseline end time is required to be >= 0 (NPI.cpp:24-26).
inline double piecewise_value(const double* end_times, const double* values, int m, double t) {
    if (t <= end_times[0]) return values[0];
    if (m == 1) return values[0];
    for (int k = 1; k < m; ++k)
        if (!(t > end_times[k])) return values[k];   // "while (time > end[idx]) ++idx"
    return values[m - 1];
}

// ---- RHS ------------------------------------------------------------------------------------------
struct Model {
    const sepaihrd_problem* pb;
    const double* s;        // slot vector
    Slots L;
    double inv_N[MAXN];
    long rhs_calls = 0;

    void init(const sepaihrd_problem* p, const double* slot_values) {
        pb = p; s = slot_values; L = slots_of(p);
        // AgeSEPAIHRDModel.cpp:46-49 / 332-334: inv_N = N > 1e-9 ? 1/N : 0
        for (int i = 0; i < L.n; ++i) inv_N[i] = (pb->population[i] > 1e-9) ? (1.0 / pb->population[i]) : 0.0;
    }

    double beta_at(double t) const {   // AgeSEPAIHRDModel.cpp:366-368 (quirk Q1: scalar beta only without a schedule)
        if (L.nb > 0) return piecewise_value(pb->beta_end_times, s + L.beta0(), L.nb, t);
        return s[L.beta_scalar()];
    }
    double kappa_at(double t) const {
        if (t < 0) return s[L.kappa0()];
        return piecewise_value(pb->kappa_end_times, s + L.kappa0(), L.nk, t);
    }

    // AgeSEPAIHRDModel::computeDerivatives (src/model/AgeSEPAIHRDModel.cpp:101-228)
    void operator()(const double* x, double* d, double t) {
        ++rhs_calls;
        const int n = L.n;
        const double* S = x + 0 * n; const double* E = x + 1 * n; const double* P = x + 2 * n;
        const double* A = x + 3 * n; const double* I = x + 4 * n; const double* H = x + 5 * n;
        const double* ICU = x + 6 * n;
        const double theta = s[L.theta()];
        double pressure[MAXN], lambda[MAXN];
        for (int i = 0; i < n; ++i) {                         // :154-157
            double total_inf = P[i] + A[i] + theta * I[i];
            pressure[i] = total_inf * s[L.h_infec(i)] * inv_N[i];
        }
        for (int i = 0; i < n; ++i) lambda[i] = 0.0;          // :162-164
        const double* M = pb->contact_matrix;
        for (int j = 0; j < n; ++j) {                         // :166-174, column-major, j outer
            const double inf_j = pressure[j];
            for (int i = 0; i < n; ++i) lambda[i] += M[j * n + i] * inf_j;
        }
        const double beta_eff = beta_at(t) * kappa_at(t);     // :176-178
        for (int i = 0; i < n; ++i) lambda[i] *= beta_eff * s[L.a(i)];   // :181-183
        const double sigma = s[L.sigma()], gamma_p = s[L.gamma_p()], gamma_A = s[L.gamma_A()];
        const double gamma_I = s[L.gamma_I()], gamma_H = s[L.gamma_H()], gamma_ICU = s[L.gamma_ICU()];
        for (int i = 0; i < n; ++i) {                         // :195-227
            double lambda_val = std::max(0.0, lambda[i]);
            double flow_SE = lambda_val * S[i];
            double flow_EP = sigma * E[i];
            double flow_P_out = gamma_p * P[i];
            double flow_PA = s[L.p(i)] * flow_P_out;
            double flow_PI = flow_P_out - flow_PA;
            double flow_IH = s[L.h(i)] * I[i];
            double flow_IR = gamma_I * I[i];
            double flow_ID_community = s[L.d_comm(i)] * I[i];
            double I_out = flow_IR + flow_IH + flow_ID_community;
            double flow_H_ICU = s[L.icu(i)] * H[i];
            double H_out = gamma_H * H[i] + s[L.d_H(i)] * H[i] + flow_H_ICU;
            double ICU_out = (gamma_ICU + s[L.d_ICU(i)]) * ICU[i];
            d[0 * n + i] = -flow_SE;
            d[1 * n + i] = flow_SE - flow_EP;
            d[2 * n + i] = flow_EP - flow_P_out;
            d[3 * n + i] = flow_PA - gamma_A * A[i];
            d[4 * n + i] = flow_PI - I_out;
            d[5 * n + i] = flow_IH - H_out;
            d[6 * n + i] = flow_H_ICU - ICU_out;
            d[7 * n + i] = gamma_A * A[i] + flow_IR + gamma_H * H[i] + gamma_ICU * ICU[i];
            d[8 * n + i] = s[L.d_H(i)] * H[i] + s[L.d_ICU(i)] * ICU[i] + flow_ID_community;
            d[9 * n + i] = flow_IH;
            d[10 * n + i] = flow_H_ICU;
        }
    }
};

// ---- Boost.Odeint controlled Dopri5 (EXTERNAL to the reference; restated, see header) --------------
// boost/numeric/odeint/stepper/runge_kutta_dopri5.hpp : do_step_impl (FSAL, with error estimate)
// boost/numeric/odeint/stepper/controlled_runge_kutta.hpp : try_step (explicit_error_stepper_fsal_tag),
//     default_error_checker::error, default_step_adjuster::{decrease_step,increase_step}
// boost/numeric/odeint/integrate/detail/integrate_times.hpp : controlled_stepper_tag overload
// boost/numeric/odeint/integrate/max_step_checker.hpp : failed_step_checker (500)
struct Dopri5 {
    int dim;
    double eps_abs, eps_rel;
    bool first_call = true;
    double* trace = nullptr; long trace_cap = 0, trace_n = 0;   // diagnostics: (t, dt, err) of every attempt (sepaihrd_oracle_trace_one)
    double dxdt[MAXS], xnew[MAXS], dxdt_new[MAXS], xerr[MAXS];
    double xt[MAXS], k2[MAXS], k3[MAXS], k4[MAXS], k5[MAXS], k6[MAXS];

    template <class Sys>
    void do_step(Sys& sys, const double* in, const double* dxdt_in, double t, double* out, double* dxdt_out,
                 double dt, double* err) {
        const double a2 = 1.0 / 5.0, a3 = 3.0 / 10.0, a4 = 4.0 / 5.0, a5 = 8.0 / 9.0;
        const double b21 = 1.0 / 5.0;
        const double b31 = 3.0 / 40.0, b32 = 9.0 / 40.0;
        const double b41 = 44.0 / 45.0, b42 = -56.0 / 15.0, b43 = 32.0 / 9.0;
        const double b51 = 19372.0 / 6561.0, b52 = -25360.0 / 2187.0, b53 = 64448.0 / 6561.0, b54 = -212.0 / 729.0;
        const double b61 = 9017.0 / 3168.0, b62 = -355.0 / 33.0, b63 = 46732.0 / 5247.0, b64 = 49.0 / 176.0,
                     b65 = -5103.0 / 18656.0;
        const double c1 = 35.0 / 384.0, c3 = 500.0 / 1113.0, c4 = 125.0 / 192.0, c5 = -2187.0 / 6784.0,
                     c6 = 11.0 / 84.0;
        const double dc1 = c1 - 5179.0 / 57600.0, dc3 = c3 - 7571.0 / 16695.0, dc4 = c4 - 393.0 / 640.0;
        const double dc5 = c5 - -92097.0 / 339200.0, dc6 = c6 - 187.0 / 2100.0, dc7 = -1.0 / 40.0;
        const int N = dim;
        // scale_sumK: t1 = a1*t2 + a2*t3 + ... evaluated left to right; coefficients dt*b formed first.
        { const double f1 = dt * b21;
          for (int i = 0; i < N; ++i) xt[i] = 1.0 * in[i] + f1 * dxdt_in[i]; }
        sys(xt, k2, t + dt * a2);
        { const double f1 = dt * b31, f2 = dt * b32;
          for (int i = 0; i < N; ++i) xt[i] = 1.0 * in[i] + f1 * dxdt_in[i] + f2 * k2[i]; }
        sys(xt, k3, t + dt * a3);
        { const double f1 = dt * b41, f2 = dt * b42, f3 = dt * b43;
          for (int i = 0; i < N; ++i) xt[i] = 1.0 * in[i] + f1 * dxdt_in[i] + f2 * k2[i] + f3 * k3[i]; }
        sys(xt, k4, t + dt * a4);
        { const double f1 = dt * b51, f2 = dt * b52, f3 = dt * b53, f4 = dt * b54;
          for (int i = 0; i < N; ++i) xt[i] = 1.0 * in[i] + f1 * dxdt_in[i] + f2 * k2[i] + f3 * k3[i] + f4 * k4[i]; }
        sys(xt, k5, t + dt * a5);
        { const double f1 = dt * b61, f2 = dt * b62, f3 = dt * b63, f4 = dt * b64, f5 = dt * b65;
          for (int i = 0; i < N; ++i)
              xt[i] = 1.0 * in[i] + f1 * dxdt_in[i] + f2 * k2[i] + f3 * k3[i] + f4 * k4[i] + f5 * k5[i]; }
        sys(xt, k6, t + dt);
        { const double f1 = dt * c1, f3 = dt * c3, f4 = dt * c4, f5 = dt * c5, f6 = dt * c6;
          for (int i = 0; i < N; ++i)
              out[i] = 1.0 * in[i] + f1 * dxdt_in[i] + f3 * k3[i] + f4 * k4[i] + f5 * k5[i] + f6 * k6[i]; }
        sys(out, dxdt_out, t + dt);   // FSAL derivative
        { const double e1 = dt * dc1, e3 = dt * dc3, e4 = dt * dc4, e5 = dt * dc5, e6 = dt * dc6, e7 = dt * dc7;
          for (int i = 0; i < N; ++i)
              err[i] = e1 * dxdt_in[i] + e3 * k3[i] + e4 * k4[i] + e5 * k5[i] + e6 * k6[i] + e7 * dxdt_out[i]; }
    }

    // returns true on success; x, t, dt updated as controlled_runge_kutta::try_step does.
    template <class Sys>
    bool try_step(Sys& sys, double* x, double& t, double& dt) {
        if (first_call) { sys(x, dxdt, t); first_call = false; }       // try_step_v1 -> initialize
        do_step(sys, x, dxdt, t, xnew, dxdt_new, dt, xerr);
        // default_error_checker::error: |err_i| / (eps_abs + eps_rel*(a_x*|x_i| + a_dxdt*dt*|dxdt_i|)), a_x = a_dxdt = 1
        const double a_x = 1.0, a_dxdt_dt = 1.0 * dt;
        double max_rel_err = 0.0;
        for (int i = 0; i < dim; ++i) {
            double v = std::fabs(xerr[i]) / (eps_abs + eps_rel * (a_x * std::fabs(x[i]) + a_dxdt_dt * std::fabs(dxdt[i])));
            max_rel_err = std::max(max_rel_err, std::fabs(v));           // norm_inf: std::max(init, |v|) ignores NaN
        }
        if (trace && trace_n < trace_cap) { trace[3 * trace_n] = t; trace[3 * trace_n + 1] = dt; trace[3 * trace_n + 2] = max_rel_err; }
        ++trace_n;
        if (max_rel_err > 1.0) {
            // decrease_step(dt, err, error_order = 4): dt *= max(0.9 * err^(-1/(4-1)), 1/5)
            dt *= std::max(9.0 / 10.0 * std::pow(max_rel_err, -1.0 / (4 - 1)), 1.0 / 5.0);
            return false;
        }
        t += dt;
        // increase_step(dt, err, stepper_order = 5)
        if (max_rel_err < 0.5) {
            double error = std::max(std::pow(5.0, -5.0), max_rel_err);
            dt *= 9.0 / 10.0 * std::pow(error, -1.0 / 5);
        }
        std::memcpy(x, xnew, sizeof(double) * dim);
        std::memcpy(dxdt, dxdt_new, sizeof(double) * dim);
        return true;
    }
};

struct IntegrateStats { long accepted = 0, rejected = 0; };

// integrate_times(controlled stepper): returns false when failed_step_checker would throw.
template <class Sys, class Obs>
bool integrate_times(Sys& sys, Dopri5& st, double* x, const double* times, int K, double dt, Obs&& obs,
                     IntegrateStats& stats, int32_t* interval_steps) {
    int i = 0;
    while (true) {
        double current_time = times[i++];
        obs(x, current_time, i - 1);
        if (i == K) break;
        int fail_steps = 0;                                   // failed_step_checker, reset on success
        long acc0 = stats.accepted, rej0 = stats.rejected;
        // less_with_sign(t1, t2, dt>0): (t2 - t1) > epsilon
        while ((times[i] - current_time) > std::numeric_limits<double>::epsilon()) {
            double current_dt = std::min(dt, times[i] - current_time);   // min_abs
            if (st.try_step(sys, x, current_time, current_dt)) {
                ++stats.accepted;
                fail_steps = 0;
                dt = std::max(dt, current_dt);                // max_abs: continue with the larger step
            } else {
                ++stats.rejected;
                if (fail_steps++ >= 500) return false;        // "if (m_steps++ >= m_max_steps) throw"
                dt = current_dt;
            }
        }
        if (interval_steps) {
            interval_steps[2 * (i - 1) + 0] = (int32_t)(stats.accepted - acc0);
            interval_steps[2 * (i - 1) + 1] = (int32_t)(stats.rejected - rej0);
        }
    }
    return true;
}

// ---- constraints ----------------------------------------------------------------------------------
// reflectBound (src/model/parameters/SEPAIHRDParameterManager.cpp:302-313)
double reflect_bound(double value, double minb, double maxb) {
    if (minb >= maxb) return minb;
    double width = maxb - minb;
    double y = std::fmod(value - minb, 2.0 * width);
    if (y < 0) y += 2.0 * width;
    if (y <= width) return minb + y;
    return maxb - (y - width);
}

// applyConstraints (.cpp:315-347)
void apply_constraints(const sepaihrd_problem* pb, int mode, const double* in, double* out) {
    for (int i = 0; i < pb->n_params; ++i) {
        double lo = pb->lower_bound[i], hi = pb->upper_bound[i];
        if (!std::isnan(lo)) {
            if (lo > hi) std::swap(lo, hi);
            out[i] = (mode == 0) ? std::min(std::max(in[i], lo), hi) : reflect_bound(in[i], lo, hi);
        } else {
            out[i] = (mode == 0) ? std::max(0.0, in[i]) : std::abs(in[i]);
        }
    }
}

// updateModelParameters (.cpp:164-287): constrained vector -> slot vector; false = the reference throws
// (negative kappa at setCalibratableValues, NPI.cpp:238-242) and calculate() returns lowest().
bool build_slots(const sepaihrd_problem* pb, const double* params, double* slots) {
    const Slots L = slots_of(pb);
    std::memcpy(slots, pb->base_slots, sizeof(double) * L.count());
    std::vector<double> c(pb->n_params);
    apply_constraints(pb, pb->constraint_mode, params, c.data());
    bool kappa_touched = false;
    for (int i = 0; i < pb->n_params; ++i) {
        int sl = pb->param_slot[i];
        if (sl < 0) continue;                                 // unknown name: warning only (.cpp:264-266)
        slots[sl] = c[i];
        if (sl >= L.kappa0() && sl < L.kappa0() + L.nk) kappa_touched = true;
    }
    if (kappa_touched)
        for (int k = 1; k < L.nk; ++k)
            if (slots[L.kappa0() + k] < 0.0) return false;
    return true;
}

// initial-state rule of calculate() (src/model/objectives/SEPAIHRDObjectiveFunction.cpp:124-163)
bool initial_state(const sepaihrd_problem* pb, const double* slots, double* x0) {
    const Slots L = slots_of(pb);
    const int n = L.n;
    std::memcpy(x0, pb->data_initial_state, sizeof(double) * NC * n);
    const double runup_days = slots[L.runup_days()], seed_exposed = slots[L.seed_exposed()];
    if (runup_days > 0 && seed_exposed > 0) {
        double total_pop = 0.0;                               // N.sum() (:102); exact for integer populations
        for (int i = 0; i < n; ++i) total_pop += pb->population[i];
        for (int i = 0; i < n; ++i) {
            double age_fraction = (total_pop > 0.0) ? pb->population[i] / total_pop : 0.0;   // :103-107
            x0[1 * n + i] = seed_exposed * age_fraction;
            for (int c = 2; c < NC; ++c) x0[c * n + i] = 0.0;
        }
    } else {
        for (int c = 1; c <= 8; ++c)
            for (int i = 0; i < n; ++i) x0[c * n + i] *= slots[L.mult0() + (c - 1)];
    }
    for (int i = 0; i < n; ++i) {
        double sum = 0;
        for (int j = 1; j < NPOP; ++j) sum += x0[j * n + i];
        if (sum > pb->population[i]) return false;
        x0[i] = pb->population[i] - sum;
    }
    return true;
}

// calculateSingleLogLikelihood (.cpp:241-279), rows in sequence (quirk Q7: OMP_NUM_THREADS=1 order)
double poisson_ll(const double* simulated, const double* observed, int rows, int cols) {
    const double epsilon = 1e-10;
    double log_likelihood = 0.0;
    for (int i = 0; i < rows; ++i) {
        double row_sum = 0.0;
        for (int j = 0; j < cols; ++j) {
            const double obs = observed[i * cols + j];
            if (obs >= 0.0 && std::isfinite(obs)) {
                double sim = simulated[i * cols + j];
                if (sim < 0.0) sim = 0.0;
                sim += epsilon;
                row_sum += (obs * std::log(sim) - sim);
            }
        }
        log_likelihood += row_sum;
    }
    return log_likelihood;
}

struct EvalScratch {
    std::vector<double> slots, traj, inc_h, inc_icu, inc_d;
    double* trace = nullptr; long trace_cap = 0;
};

// calculate() (.cpp:62-235) with a null cache. traj_out: [K][11n] (optional).
double eval_one(const sepaihrd_problem* pb, const double* params, uint32_t* status_out, double* traj_out,
                int32_t* interval_steps, int64_t* counts, EvalScratch& sc, const double* given_state = nullptr) {
    const double LOWEST = std::numeric_limits<double>::lowest();
    const Slots L = slots_of(pb);
    const int n = L.n, K = pb->n_times, dim = NC * n;
    uint32_t st = SEPAIHRD_ST_OK;
    if (counts) counts[0] = counts[1] = counts[2] = 0;
    sc.slots.resize(L.count());
    if (!build_slots(pb, params, sc.slots.data())) { if (status_out) *status_out = SEPAIHRD_ST_INVALID_PARAM; return LOWEST; }
    double x[MAXS];
    if (given_state) std::memcpy(x, given_state, sizeof(double) * dim);   // Simulator::run(initial_state, ...) (Simulator.cpp:60-150)
    else if (!initial_state(pb, sc.slots.data(), x)) { if (status_out) *status_out = SEPAIHRD_ST_S_OVERFLOW; return LOWEST; }
    double x0[MAXS];
    std::memcpy(x0, x, sizeof(double) * dim);

    sc.traj.resize((size_t)K * dim);
    double* traj = sc.traj.data();
    Model model; model.init(pb, sc.slots.data());
    Dopri5 stepper; stepper.dim = dim; stepper.eps_abs = pb->abs_tol; stepper.eps_rel = pb->rel_tol;
    stepper.trace = sc.trace; stepper.trace_cap = sc.trace_cap;
    IntegrateStats stats;
    bool ok = integrate_times(model, stepper, x, pb->times, K, pb->dt_hint,
        [&](const double* xs, double, int idx) { std::memcpy(traj + (size_t)idx * dim, xs, sizeof(double) * dim); },
        stats, interval_steps);
    if (counts) { counts[0] = stats.accepted; counts[1] = stats.rejected; counts[2] = model.rhs_calls; }
    if (traj_out) std::memcpy(traj_out, traj, sizeof(double) * (size_t)K * dim);
    if (!ok) { if (status_out) *status_out = SEPAIHRD_ST_STEP_FAILURE; return LOWEST; }

    // runup_offset_ = first index with t >= 0 (:39-46); num_obs_points_ must match the data rows (:176-178)
    int runup_offset = 0;
    for (int i = 0; i < K; ++i) if (pb->times[i] >= 0.0) { runup_offset = i; break; }
    const int num_obs = K - runup_offset;
    if (num_obs != pb->n_obs) { if (status_out) *status_out = SEPAIHRD_ST_INVALID_PARAM; return LOWEST; }

    // daily incidence = first difference, row 0 against the initial state, clamped at 0 (:191-215)
    sc.inc_h.resize((size_t)K * n); sc.inc_icu.resize((size_t)K * n); sc.inc_d.resize((size_t)K * n);
    for (int r = 0; r < K; ++r) {
        const double* cur = traj + (size_t)r * dim;
        const double* prev = (r == 0) ? x0 : traj + (size_t)(r - 1) * dim;
        for (int a = 0; a < n; ++a) {
            sc.inc_h[(size_t)r * n + a] = std::max(cur[9 * n + a] - prev[9 * n + a], 0.0);
            sc.inc_icu[(size_t)r * n + a] = std::max(cur[10 * n + a] - prev[10 * n + a], 0.0);
            sc.inc_d[(size_t)r * n + a] = std::max(cur[8 * n + a] - prev[8 * n + a], 0.0);
        }
    }
    double ll_hosp = poisson_ll(sc.inc_h.data() + (size_t)runup_offset * n, pb->obs_hosp, num_obs, n);
    double ll_icu = poisson_ll(sc.inc_icu.data() + (size_t)runup_offset * n, pb->obs_icu, num_obs, n);
    double ll_deaths = poisson_ll(sc.inc_d.data() + (size_t)runup_offset * n, pb->obs_deaths, num_obs, n);
    double total = ll_hosp + ll_icu + ll_deaths;              // :222
    if (std::isnan(total) || std::isinf(total)) { total = LOWEST; st |= SEPAIHRD_ST_NONFINITE; }
    if (status_out) *status_out = st;
    return total;
}

}  // namespace

// =====================================================================================================
extern "C" {

int32_t sepaihrd_oracle_slot_for_name(int32_t n, int32_t nb, int32_t nk, const char* cname) {
    // dispatch order of SEPAIHRDParameterManager::updateModelParameters (.cpp:197-267); construction-time
    // validation (.cpp:45-88) yields -2.
    const Slots L{n, nb, nk};
    const std::string name(cname ? cname : "");
    unsigned long idx = 0;
    auto age = [&](size_t off, int base) -> int32_t {
        if (!parse_index(name, off, idx)) return -2;
        if (idx >= (unsigned long)n) return -2;
        return base + (int)idx;
    };
    if (name == "beta") return L.beta_scalar();
    if (starts_with(name, "beta_")) {
        if (!parse_index(name, 5, idx)) return -2;
        if (idx < 1 || idx > (unsigned long)nb) return -2;   // beta_idx < beta_values.size() else throw (.cpp:203-207)
        return L.beta0() + (int)idx - 1;
    }
    if (name == "theta") return L.theta();
    if (name == "sigma") return L.sigma();
    if (name == "gamma_p") return L.gamma_p();
    if (name == "gamma_A") return L.gamma_A();
    if (name == "gamma_I") return L.gamma_I();
    if (name == "gamma_H") return L.gamma_H();
    if (name == "gamma_ICU") return L.gamma_ICU();
    if (starts_with(name, "a_")) return age(2, L.a(0));
    if (starts_with(name, "h_infec_")) return age(8, L.h_infec(0));
    if (starts_with(name, "p_")) return age(2, L.p(0));
    if (starts_with(name, "h_")) return age(2, L.h(0));
    if (starts_with(name, "icu_")) return age(4, L.icu(0));
    if (starts_with(name, "d_H_")) return age(4, L.d_H(0));
    if (starts_with(name, "d_ICU_")) return age(6, L.d_ICU(0));
    if (starts_with(name, "d_community_")) return age(12, L.d_comm(0));
    if (name == "seed_exposed") return L.seed_exposed();
    if (name == "runup_days") return L.runup_days();
    static const char* mult[8] = {"E0_multiplier", "P0_multiplier", "A0_multiplier", "I0_multiplier",
                                  "H0_multiplier", "ICU0_multiplier", "R0_multiplier", "D0_multiplier"};
    for (int m = 0; m < 8; ++m) if (name == mult[m]) return L.mult0() + m;
    if (starts_with(name, "kappa_")) {
        // calibratable NPI names are kappa_2..kappa_nk (main.cpp:104-123, fixed baseline kappa_1)
        if (!parse_index(name, 6, idx)) return -2;
        if (idx < 2 || idx > (unsigned long)nk) return -2;
        return L.kappa0() + (int)idx - 1;
    }
    return -1;
}

void sepaihrd_oracle_apply_constraints(const sepaihrd_problem* pb, int32_t mode, const double* in, double* out) {
    apply_constraints(pb, mode, in, out);
}

void sepaihrd_oracle_rhs(const sepaihrd_problem* pb, const double* slots, const double* state, double t, double* dxdt) {
    Model m; m.init(pb, slots);
    m(state, dxdt, t);
}

double sepaihrd_oracle_poisson_ll(const double* simulated, const double* observed, int32_t rows, int32_t cols) {
    return poisson_ll(simulated, observed, rows, cols);
}

// CalibrationData::getInitialSEPAIHRDState (src/utils/GetCalibrationData.cpp:107-234)
void sepaihrd_oracle_initial_state_from_data(int32_t n, const double* N, const double* cum_confirmed0,
                                             const double* cum_deaths0, const double* cum_hosp0,
                                             const double* cum_icu0, double sigma, double gamma_p,
                                             double gamma_a, double gamma_i, const double* p_asym, double* out) {
    std::vector<double> D0(n), H0(n), ICU0(n), CumH0(n), CumICU0(n), I0(n), E0(n), P0(n), A0(n), R0(n, 0.0);
    for (int i = 0; i < n; ++i) {
        D0[i] = std::max(cum_deaths0[i], 0.0);                           // :139-145
        H0[i] = std::max(cum_hosp0[i], 0.0);
        ICU0[i] = std::max(cum_icu0[i], 0.0);
        CumH0[i] = std::max(cum_hosp0[i], 0.0);
        CumICU0[i] = std::max(cum_icu0[i], 0.0);
        I0[i] = std::max(cum_confirmed0[i] - D0[i], 0.0);                // :148
    }
    for (int i = 0; i < n; ++i) {                                        // :154-167
        double p_i = std::clamp(p_asym[i], 0.0, 1.0);
        double one_minus_p_i = 1.0 - p_i;
        if (gamma_p > 1e-9 && one_minus_p_i > 1e-9) P0[i] = I0[i] * gamma_i / (one_minus_p_i * gamma_p);
        else P0[i] = I0[i];
        if (gamma_a > 1e-9) A0[i] = P0[i] * p_i * gamma_p / gamma_a; else A0[i] = P0[i] * p_i;
        if (sigma > 1e-9) E0[i] = P0[i] * gamma_p / sigma; else E0[i] = P0[i];
    }
    for (int i = 0; i < n; ++i) { E0[i] = std::max(E0[i], 0.0); P0[i] = std::max(P0[i], 0.0); A0[i] = std::max(A0[i], 0.0); }
    for (int i = 0; i < n; ++i) {                                        // :174-180
        D0[i] = std::min(D0[i], N[i]);
        ICU0[i] = std::min(ICU0[i], std::max(0.0, N[i] - D0[i]));
        H0[i] = std::min(H0[i], std::max(0.0, N[i] - D0[i] - ICU0[i]));
        I0[i] = std::min(I0[i], std::max(0.0, N[i] - D0[i] - ICU0[i] - H0[i]));
        R0[i] = std::min(R0[i], std::max(0.0, N[i] - D0[i] - ICU0[i] - H0[i] - I0[i]));
    }
    for (int i = 0; i < n; ++i) {                                        // :188-202
        double sum_set = I0[i] + H0[i] + ICU0[i] + R0[i] + D0[i];
        double sum_inferred = E0[i] + P0[i] + A0[i];
        double available = N[i] - sum_set;
        if (available < 0) available = 0;
        if (sum_inferred > available) {
            double scale = (sum_inferred > 1e-9) ? available / sum_inferred : 0.0;
            E0[i] *= scale; P0[i] *= scale; A0[i] *= scale;
        }
    }
    for (int i = 0; i < NC * n; ++i) out[i] = 0.0;
    for (int i = 0; i < n; ++i) {
        out[1 * n + i] = E0[i]; out[2 * n + i] = P0[i]; out[3 * n + i] = A0[i]; out[4 * n + i] = I0[i];
        out[5 * n + i] = H0[i]; out[6 * n + i] = ICU0[i]; out[7 * n + i] = R0[i]; out[8 * n + i] = D0[i];
        out[9 * n + i] = CumH0[i]; out[10 * n + i] = CumICU0[i];
    }
    for (int i = 0; i < n; ++i) {                                        // :216-223
        double sum_non_S = 0;
        for (int j = 1; j < 9; ++j) sum_non_S += out[j * n + i];
        out[i] = std::max(0.0, N[i] - sum_non_S);
    }
}

double sepaihrd_oracle_eval_one(const sepaihrd_problem* pb, const double* params, uint32_t* out_status,
                                double* out_traj, int32_t* out_interval_steps, int64_t* out_counts) {
    EvalScratch sc;
    return eval_one(pb, params, out_status, out_traj, out_interval_steps, out_counts, sc);
}

int64_t sepaihrd_oracle_trace_one(const sepaihrd_problem* pb, const double* params, int64_t cap, double* out_trace, double* out_ll) {
    EvalScratch sc;
    sc.trace = out_trace; sc.trace_cap = (long)cap;
    int64_t counts[3] = {0, 0, 0};
    uint32_t st = 0;
    const double ll = eval_one(pb, params, &st, nullptr, nullptr, counts, sc);
    if (out_ll) *out_ll = ll;
    return counts[0] + counts[1];
}

int32_t sepaihrd_oracle_eval_batch(const sepaihrd_problem* pb, const double* params, int64_t B, int64_t ld,
                                   double* out_ll, uint32_t* out_status, int32_t* out_steps, int32_t nthreads) {
    int used = 1;
#if defined(_OPENMP)
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    used = nthreads;
#else
    (void)nthreads;
#endif
#pragma omp parallel num_threads(used)
    {
        EvalScratch sc;
#pragma omp for schedule(dynamic, 16)
        for (int64_t b = 0; b < B; ++b) {
            uint32_t st = 0; int64_t counts[3];
            out_ll[b] = eval_one(pb, params + b * ld, &st, nullptr, nullptr, counts, sc);
            if (out_status) out_status[b] = st;
            if (out_steps) { out_steps[2 * b] = (int32_t)counts[0]; out_steps[2 * b + 1] = (int32_t)counts[1]; }
        }
    }
    return used;
}

static int32_t simulate_impl(const sepaihrd_problem* pb, const double* params, int64_t B, int64_t ld,
                             const double* init_states, int64_t init_stride,
                             int32_t what, int32_t stride, double* out, uint32_t* out_status,
                             int32_t nthreads) {
    const int n = pb->n_ages, K = pb->n_times, dim = NC * n;
    const int W = (what == SEPAIHRD_TRAJ_FULL) ? dim : 3 * n;
    if (stride < 1) stride = 1;
    const int Kout = (K + stride - 1) / stride;
    int used = 1;
#if defined(_OPENMP)
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    used = nthreads;
#else
    (void)nthreads;
#endif
#pragma omp parallel num_threads(used)
    {
        EvalScratch sc;
        std::vector<double> traj((size_t)K * dim);
#pragma omp for schedule(dynamic, 4)
        for (int64_t b = 0; b < B; ++b) {
            uint32_t st = 0;
            eval_one(pb, params + b * ld, &st, traj.data(), nullptr, nullptr, sc, init_states ? init_states + b * init_stride : nullptr);
            double* o = out + (size_t)b * Kout * W;
            const bool bad = (st & (SEPAIHRD_ST_S_OVERFLOW | SEPAIHRD_ST_INVALID_PARAM | SEPAIHRD_ST_STEP_FAILURE)) != 0;
            for (int r = 0; r < Kout; ++r) {
                const double* x = traj.data() + (size_t)(r * stride) * dim;
                for (int w = 0; w < W; ++w) {
                    double v;
                    if (bad) v = std::numeric_limits<double>::quiet_NaN();
                    else if (what == SEPAIHRD_TRAJ_FULL) v = x[w];
                    else { int blk = w / n, a = w % n; int comp = (blk == 0) ? 8 : (blk == 1 ? 9 : 10); v = x[comp * n + a]; }
                    o[(size_t)r * W + w] = v;
                }
            }
            if (out_status) out_status[b] = st & ~SEPAIHRD_ST_NONFINITE;
        }
    }
    return used;
}

int32_t sepaihrd_oracle_simulate_batch(const sepaihrd_problem* pb, const double* params, int64_t B, int64_t ld,
                                       int32_t what, int32_t stride, double* out, uint32_t* out_status,
                                       int32_t nthreads) {
    return simulate_impl(pb, params, B, ld, nullptr, 0, what, stride, out, out_status, nthreads);
}

int32_t sepaihrd_oracle_simulate_from_state(const sepaihrd_problem* pb, const double* params, int64_t B, int64_t ld,
                                            const double* initial_states, int64_t state_stride, int32_t what,
                                            int32_t stride, double* out, uint32_t* out_status, int32_t nthreads) {
    return simulate_impl(pb, params, B, ld, initial_states, state_stride, what, stride, out, out_status, nthreads);
}

void sepaihrd_oracle_jitter_params(const sepaihrd_problem* pb, const double* base, const double* sigmas,
                                   uint32_t seed, int64_t B, double* out) {
    // sepaihrd_objective_benchmark_main.cpp:418-419, 452-460
    std::mt19937 rng(seed);
    std::normal_distribution<double> normal(0.0, 1.0);
    const int P = pb->n_params;
    std::vector<double> cand(P);
    for (int64_t k = 0; k < B; ++k) {
        for (int i = 0; i < P; ++i) cand[i] = base[i] + sigmas[i] * normal(rng);
        apply_constraints(pb, pb->constraint_mode, cand.data(), out + k * P);
    }
}

void sepaihrd_oracle_uniform_params(const sepaihrd_problem* pb, uint32_t seed, int64_t B, double* out) {
    std::mt19937 rng(seed);
    std::uniform_real_distribution<> uni(0.0, 1.0);
    const int P = pb->n_params;
    for (int64_t k = 0; k < B; ++k)
        for (int i = 0; i < P; ++i) {
            double lo = pb->lower_bound[i], hi = pb->upper_bound[i];
            out[k * P + i] = lo + uni(rng) * (hi - lo);      // ParticleSwarmOptimizer.cpp:291
        }
}

}  // extern "C"
