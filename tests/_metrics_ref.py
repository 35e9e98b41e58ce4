"""numpy restatement of the reference's post-calibration metrics, written independently of host/analysis.cpp:
MetricsCalculator::calculateEssentialMetrics (src/model/MetricsCalculator.cpp:8-166) and the next-generation-matrix
reproduction numbers (src/model/ReproductionNumberCalculator.cpp:18-171, full 4n x 4n F V^-1 + numpy eigenvalues)."""
import numpy as np


def unpack(problem, slots=None):
    """Model parameters by name from a slot vector (base slots by default)."""
    lay = problem.layout
    s = np.array(problem.base_slots if slots is None else slots, dtype=float)
    n = problem.n_ages
    d = dict(beta_values=s[lay.beta0:lay.beta0 + lay.nb], kappa_values=s[lay.kappa0:lay.kappa0 + lay.nk],
             beta=s[lay.beta_scalar], N=np.array(problem.population), M=np.array(problem.contact_matrix),
             beta_end=np.array(problem.beta_end_times), kappa_end=np.array(problem.kappa_end_times))
    for name in ("theta", "sigma", "gamma_p", "gamma_A", "gamma_I", "gamma_H", "gamma_ICU"):
        d[name] = s[lay.scalar(name)]
    for blk in ("a", "h_infec", "p", "h", "icu", "d_H", "d_ICU", "d_community"):
        d[blk] = s[lay.age(blk, 0):lay.age(blk, 0) + n]
    return d


def piecewise(t, ends, values):
    for k, e in enumerate(ends):
        if t <= e:
            return values[k]
    return values[-1]


def kappa_at(t, q, kappa_values=None):
    kv = q["kappa_values"] if kappa_values is None else kappa_values
    if t < 0:
        return kv[0]
    return piecewise(t, q["kappa_end"], kv)


def ngm_radius(q, X, t, clamp, kappa_values=None):
    n = len(q["N"])
    beta = piecewise(t, q["beta_end"], q["beta_values"]) if len(q["beta_end"]) else q["beta"]
    kappa = kappa_at(t, q, kappa_values)
    F = np.zeros((4 * n, 4 * n)); V = np.zeros((4 * n, 4 * n))
    for i in range(n):
        for j in range(n):
            T = beta * kappa * q["M"][i, j] * q["a"][i] * q["h_infec"][j] * (X[i] / q["N"][j])
            if clamp:
                T = max(0.0, T)
            F[i, n + j] = T; F[i, 2 * n + j] = T; F[i, 3 * n + j] = q["theta"] * T
    for a in range(n):
        e, p, aa, ii = a, n + a, 2 * n + a, 3 * n + a
        V[e, e] = q["sigma"]; V[p, e] = -q["sigma"]; V[p, p] = q["gamma_p"]
        V[aa, p] = -q["p"][a] * q["gamma_p"]; V[ii, p] = -(1.0 - q["p"][a]) * q["gamma_p"]
        V[aa, aa] = q["gamma_A"]; V[ii, ii] = q["gamma_I"] + q["h"][a]
    return float(np.abs(np.linalg.eigvals(F @ np.linalg.inv(V))).max())


def essential_metrics(q, times, traj, x0, npi_kappa_values=None):
    """q: parameters of the run; npi_kappa_values: the kappa values of the metrics model's NPI schedule (the template's
    for a scenario run, the run's own otherwise).  Returns (scalars[12], age[4, n], Rt[K], sero[K])."""
    n = len(q["N"]); K = len(times)
    traj = np.asarray(traj).reshape(K, 11, n); x0 = np.asarray(x0).reshape(11, n)
    N = q["N"]; total = N.sum()
    cum = x0[1:8].sum(axis=0).copy()
    R0 = ngm_radius(q, N, 0.0, False, npi_kappa_values)
    target = int(np.argmin(np.abs(np.asarray(times) - 64.0)))
    rt = np.zeros(K); sero = np.zeros(K)
    peakH = peakU = tH = tU = 0.0
    sero_target = 0.0
    for t in range(K):
        S, P, A, I, H, U = traj[t, 0], traj[t, 2], traj[t, 3], traj[t, 4], traj[t, 5], traj[t, 6]
        dt = times[t] - times[t - 1] if t > 0 else 1.0
        rt[t] = ngm_radius(q, S, times[t], True, npi_kappa_values)
        sero[t] = (total - S.sum()) / total
        if H.sum() > peakH:
            peakH, tH = H.sum(), times[t]
        if U.sum() > peakU:
            peakU, tU = U.sum(), times[t]
        beta = q["beta"] if np.isfinite(q["beta"]) else piecewise(times[t], q["beta_end"], q["beta_values"])
        lam = beta * kappa_at(times[t], q, npi_kappa_values) * (q["M"] @ ((P + A + q["theta"] * I) / N))
        cum += lam * S * dt
        if t == target:
            sero_target = sero[t]
    deaths = traj[-1, 8] - x0[8]; ch = traj[-1, 9] - x0[9]; cu = traj[-1, 10] - x0[10]
    age = np.zeros((4, n))
    age[3] = cum / N
    ok = cum > 1.0
    age[0, ok] = np.clip(deaths[ok] / cum[ok], 0, 1); age[1, ok] = np.clip(ch[ok] / cum[ok], 0, 1); age[2, ok] = np.clip(cu[ok] / cum[ok], 0, 1)
    scal = np.array([R0, deaths.sum() / cum.sum() if cum.sum() > 1e-9 else 0.0, cum.sum() / total, peakH, peakU, tH, tU, deaths.sum(),
                     max(0.0, rt.max()), min(1e6, rt.min()), rt[-1], sero_target])
    return scal, age, rt, sero
