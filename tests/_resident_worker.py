"""Worker of the two-process GPU test of the resident samplers: one rank of a sharded device-resident run.  Both ranks use
cuda:0 (CUDA IPC works between processes on one device), rendezvous over gloo (NCCL refuses two ranks on one GPU), the data
path is the peer-memory exchange (transport "p2p").  Writes <out>.rank<r>.npz."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

what, out = sys.argv[1], sys.argv[2]
import torch
import torch.distributed as dist

pkg = entry.load_package()
from sepaihrd_b200 import resident  # noqa: E402
from sepaihrd_b200.evaluator import BatchEvaluator  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("gloo", rank=rank, world_size=world)
torch.cuda.set_device(0)
p = pkg.load_default_problem()
if what in ("mh", "mhw"):
    rp = p.__class__.from_json(dict(p.to_json(), constraint_mode=1))
    with BatchEvaluator(rp, device=0) as ev:
        r = resident.run_mh_resident(ev, p.sigmas, p.base_params(), n_chains=203, iterations=12 if what == "mh" else 40, seed=1234, rank=rank,
                                     world=world, transport="p2p", lookahead=1 if what == "mh" else 6)
    np.savez(f"{out}.rank{rank}.npz", lo=r["chains"][0], hi=r["chains"][1], x=r["x"], logpost=r["logpost"], accepts=r["accepts"],
             all_logpost=r["all_logpost"], scale=r["scale"], trace=r["best_trace"], status=r["exchange_status"])
else:
    with BatchEvaluator(p, device=0) as ev:
        r = resident.run_pso_resident(ev, swarm_size=301, iterations=5, seed=7, initial=p.base_params(), rank=rank, world=world, transport="p2p",
                                      return_positions=True)
    np.savez(f"{out}.rank{rank}.npz", lo=r["particles"][0], hi=r["particles"][1], best_value=r["best_value"], best_position=r["best_position"],
             trace=r["trace"], positions=r["final_positions"], status=r["exchange_status"])
dist.barrier()
dist.destroy_process_group()
