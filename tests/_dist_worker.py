"""Worker of the world-size-2 tests: one rank of a sharded Metropolis-Hastings or PSO run with the CPU oracle as the
evaluator (test infrastructure) and gloo as the backend.  Writes its result to <out>.rank<r>.npz."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

what, out = sys.argv[1], sys.argv[2]
pkg = entry.load_package(); orc = entry.load_oracle()
from sepaihrd_b200 import drivers  # noqa: E402
from sepaihrd_b200.distributed import Comm  # noqa: E402

p = pkg.load_default_problem()
reflect = p.__class__.from_json(dict(p.to_json(), constraint_mode=1))
o_clamp, o_reflect = orc.Oracle(p), orc.Oracle(reflect)
comm = Comm(backend="gloo")
if what == "mh":
    r = drivers.run_multichain_mh(lambda x: o_reflect.eval_batch(x, nthreads=1)[0], p.sigmas, p.lower_bound, p.upper_bound, p.base_params(),
                                  n_chains=6, iterations=9, seed=1234, comm=comm)
    np.savez(f"{out}.rank{comm.rank}.npz", lo=r["chains"][0], hi=r["chains"][1], x=r["x"], logpost=r["logpost"], accepts=r["accepts"],
             all_logpost=r["all_logpost"], scale=r["scale"])
else:
    r = drivers.run_pso(lambda x: o_clamp.eval_batch(x, nthreads=1)[0], p.sigmas, p.lower_bound, p.upper_bound, swarm_size=10, iterations=4,
                        seed=77, initial=p.base_params(), comm=comm)
    np.savez(f"{out}.rank{comm.rank}.npz", lo=r["particles"][0], hi=r["particles"][1], best_value=r["best_value"],
             best_position=r["best_position"], trace=r["trace"])
comm.close()
