import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import __graft_entry__ as g
pkg = g.load_package(); orc = g.load_oracle()
p = pkg.load_default_problem(); o = orc.Oracle(p)
N = 2048
for name, P in (("uniform", o.uniform_params(N, seed=2)), ("jitter", o.jitter_params(N, seed=1))):
    A = np.array([o.eval_one(x, want_interval_steps=True)["interval_steps"].sum(axis=1) for x in P])   # [N][325] attempts per day
    tot = A.sum(1)
    def cost(order):
        B = A[order].reshape(N // 8, 8, -1)
        return B.max(axis=1).sum()          # warp-attempts
    ideal = A.sum() / 8.0
    rnd = cost(np.arange(N))
    by_total = cost(np.argsort(tot, kind="stable"))
    free = A.reshape(N // 8, 8, -1).sum(2).max(1).sum()      # free-running bound: max over sets of the total
    # similarity ordering: sort by the day-profile projected on its first principal components
    X = A - A.mean(0); U, S, Vt = np.linalg.svd(X, full_matrices=False)
    pc1 = cost(np.argsort(U[:, 0], kind="stable"))
    print(f"{name}: mean attempts {tot.mean():.1f} (std {tot.std():.1f}); idle lane-attempts: as given {1 - ideal / rnd:.3f}, sorted by total {1 - ideal / by_total:.3f}, "
          f"sorted by PC1 of the day profile {1 - ideal / pc1:.3f}, free-running bound {1 - ideal / free:.3f}")
    # which parameters explain the total?
    c = [abs(np.corrcoef(P[:, j], tot)[0, 1]) for j in range(P.shape[1])]
    top = np.argsort(c)[::-1][:6]
    print("   |corr(param, total attempts)| top:", [(p.param_names[j], round(c[j], 2)) for j in top])

# orderings available BEFORE the evaluation: by single parameters, and by a linear predictor of the day-profile component
P = o.uniform_params(N, seed=2)
A = np.array([o.eval_one(x, want_interval_steps=True)["interval_steps"].sum(axis=1) for x in P])
ideal = A.sum() / 8.0
cost = lambda order: A[order].reshape(N // 8, 8, -1).max(axis=1).sum()
names = p.param_names
for nm in ("sigma", "beta_2", "kappa_2", "gamma_A", "gamma_p"):
    j = names.index(nm)
    print(f"sorted by {nm}: idle {1 - ideal / cost(np.argsort(P[:, j], kind='stable')):.3f}")
X = A - A.mean(0); U, S, Vt = np.linalg.svd(X, full_matrices=False)
Z = (P - P.mean(0)) / (P.std(0) + 1e-300)
for k in (1, 2):
    w, *_ = np.linalg.lstsq(Z[:N // 2], U[:N // 2, :k] * S[:k], rcond=None)      # fit on one half
    pred = Z @ w
    order = np.argsort(pred[:, 0], kind="stable")
    print(f"sorted by a linear predictor of PC1 (fit on half): idle {1 - ideal / cost(order):.3f}; top weights",
          [(names[j], round(float(w[j, 0]), 2)) for j in np.argsort(-np.abs(w[:, 0]))[:5]])
    break
# two-key bucketing: sigma deciles, then beta_2 inside
js, jb = names.index("sigma"), names.index("beta_2")
dec = np.floor((np.argsort(np.argsort(P[:, js])) / N) * 16).astype(int)
order = np.lexsort((P[:, jb], dec))
print(f"sigma in 16 bins, beta_2 inside: idle {1 - ideal / cost(order):.3f}")
