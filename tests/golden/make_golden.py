#!/usr/bin/env python
"""Regenerate tests/golden/spain2020_golden.json.

The reference itself cannot be built in the container (Boost/Eigen absent), so these
vectors are produced by the CPU oracle (oracle/sepaihrd_oracle.cpp) on the committed Spain-2020
problem fixture.  They pin the oracle against regressions and give the GPU tests fixed inputs and
outputs.  The only numbers with an origin outside this repository are in "survey_anchor": the
independent transcription made during the survey.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
orc = entry.load_oracle()
p = pkg.load_default_problem()
o = orc.Oracle(p)

base = p.base_params()
r = o.eval_one(base, want_traj=True, want_interval_steps=True)
times = p.times
n = p.n_ages


def at(t, comp, age):
    return float(r["traj"][int(np.where(times == t)[0][0]), comp * n + age])


jit = o.jitter_params(24, seed=1)
uni = o.uniform_params(24, seed=2)
ll_j, st_j, steps_j, _ = o.eval_batch(jit)
ll_u, st_u, steps_u, _ = o.eval_batch(uni)
o_reflect = orc.Oracle(p, constraint_mode=pkg.REFLECT)
wild = base[None, :] + 6.0 * p.sigmas[None, :] * np.random.default_rng(11).standard_normal((16, p.n_params))
ll_r, st_r, steps_r, _ = o_reflect.eval_batch(wild)
traj_rows = [0, 20, 33, 83, 131, 203, 257, 325]

gold = dict(
    note="oracle-generated; see make_golden.py",
    survey_anchor=dict(logL=1.206869676728e+06, ll_H=9.879936740455e+05, ll_ICU=4.328482013563e+04,
                       ll_D=1.755911825467e+05, accepted=441, rejected=45, rhs_calls=2917,
                       D_age3_t305=1.771146891e+04, CumH_age0_t13=1.728118076e+02, S_age0_t0=1.407358505e+07),
    default=dict(params=[float(x) for x in base], logL=r["ll"], accepted=r["accepted"], rejected=r["rejected"],
                 rhs_calls=r["rhs_calls"], D_age3_t305=at(305, 8, 3), CumH_age0_t13=at(13, 9, 0),
                 S_age0_t0=at(0, 0, 0), interval_steps=r["interval_steps"].tolist(),
                 traj_rows=traj_rows, traj=[[float(v) for v in r["traj"][i]] for i in traj_rows]),
    jitter=dict(seed=1, params=jit.tolist(), logL=ll_j.tolist(), steps=steps_j.tolist(), status=st_j.tolist()),
    uniform=dict(seed=2, params=uni.tolist(), logL=ll_u.tolist(), steps=steps_u.tolist(), status=st_u.tolist()),
    reflect=dict(params=wild.tolist(), logL=ll_r.tolist(), steps=steps_r.tolist(), status=st_r.tolist()),
)
out = os.path.join(ROOT, "tests", "golden", "spain2020_golden.json")
with open(out, "w") as f:
    json.dump(gold, f, separators=(",", ":"))
    f.write("\n")
print("wrote", out, os.path.getsize(out), "bytes; default logL", repr(r["ll"]))
