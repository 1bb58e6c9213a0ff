"""
Writes tests/golden/oracle_regression.json: numbers produced by THE ORACLE ITSELF on the in-tree fixtures with seeded
sources (NOT reference outputs -- the reference stores none).  They freeze the oracle's behaviour so that an
accidental change of oracle/lqcd_oracle.c (the checker every GPU parity test trusts) is caught by the CPU suite.
Run in the build container:  python tests/golden/make_oracle_regression.py
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as orc  # noqa: E402

G = Path(__file__).parent
DIMS = (4, 4, 4, 4)
out = {}
Uw, Us = np.load(G / "wilson_4444.npy"), np.load(G / "staggered_4444.npy")
opw, ops = orc.make_op(DIMS, kappa=0.141139), orc.make_op(DIMS, mass=0.5)
pw, ps = orc.gaussian_field(DIMS, orc.WILSON, seed=112), orc.gaussian_field(DIMS, orc.STAGGERED, seed=112)


def sig(a):
    a = np.asarray(a).ravel()
    w = np.cos(np.arange(a.size) * 0.37) + 1j * np.sin(np.arange(a.size) * 0.11)
    return {"norm2": float(np.vdot(a, a).real), "probe_re": float(np.vdot(w, a).real), "probe_im": float(np.vdot(w, a).imag)}


out["wilson_D"] = sig(orc.apply(opw, orc.WILSON, orc.D, Uw, pw))
out["wilson_Ddag"] = sig(orc.apply(opw, orc.WILSON, orc.DDAG, Uw, pw))
out["stag_D"] = sig(orc.apply(ops, orc.STAGGERED, orc.D, Us, ps))
r = orc.cg(opw, orc.WILSON, Uw, pw)
out["wilson_cg"] = {"iters": r["iters"], **sig(r["x"])}
r = orc.cgnr(opw, orc.WILSON, Uw, orc.point_source(DIMS, orc.WILSON, 0, 0))
out["wilson_cgnr_point"] = {"iters": r["iters"], **sig(r["x"])}
r = orc.cg(ops, orc.STAGGERED, Us, ps)
out["stag_cg"] = {"iters": r["iters"], **sig(r["x"])}
X = orc.cg(opw, orc.WILSON, Uw, pw, eps=1e-24)["x"]
out["wilson_force"] = sig(orc.force(opw, orc.WILSON, Uw, X, orc.apply(opw, orc.WILSON, orc.D, Uw, X)))
(G / "oracle_regression.json").write_text(json.dumps(out, indent=1))
print(json.dumps(out, indent=1))
