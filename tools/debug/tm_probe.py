"""development aid: one Wilson D application with the TM_DEBUG build (bounded mbarrier waits) against the oracle"""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path[:0] = [str(ROOT), str(ROOT / "latticeqcd.jl_b200")]
import numpy as np
import lqcd_b200 as q
from oracle import oracle as orc
dims = tuple(int(v) for v in sys.argv[1].split("x"))
Uh = orc.random_su3(dims, seed=5)
U = q.gaugefields_from_array(Uh)
x = q.Initialize_pseudofermion_fields(U[0], "Wilson")
D = q.Dirac_operator(U, x, {"Dirac_operator": "Wilson", "κ": 0.12, "r": 1.0, "boundarycondition": [1, 1, 1, -1]})
src = orc.gaussian_field(dims, orc.WILSON, seed=13)
x.from_host(src)
y = q.similar(x)
q.mul_(y, D, x)
want = orc.apply(orc.make_op(dims, kappa=0.12), orc.WILSON, orc.D, Uh, src)
print(f"probe {dims}: rel err {np.abs(y.to_host() - want).max() / np.abs(want).max():.2e}", flush=True)
