"""
Regenerates tests/golden/*.npy from the reference's in-tree gauge fixtures (run in the build container
only; /root/reference does not exist on the GPU box).

Inputs  : /root/reference/test/confs_*/conf_00000100.ildg{,.txt}   (thermalised 4^4 configurations, the
          start points of test/test_wilson.toml:17, test_staggered.toml, test_Nf2.toml ...)
Outputs : <name>.npy  complex128 links in the numpy host layout [mu,t,z,y,x,b,a] (oracle/oracle.py)
          fixtures.json: plaquette of each fixture computed by the oracle, compared with SURVEY.md section 4.
These are golden INPUTS.  The reference stores no vector-level outputs for this path (parity unpinned).
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as orc  # noqa: E402

REF = Path("/root/reference/test")
FIX = {
    "wilson_4444": ("confs_HMC_L04040404_beta5.7_Wilson_kappa0.141139", (4, 4, 4, 4), 0.565800226845),
    "staggered_4444": ("confs_HMC_L04040404_beta5.7_Staggered_mass0.5", (4, 4, 4, 4), 0.575584039475),
    "staggered_nf2_4444": ("confs_HMC_L04040404_beta5.7_Staggered_mass0.5_Nf2", (4, 4, 4, 4), 0.566501729368),
    "staggered_nf3_4444": ("confs_HMC_L04040404_beta5.7_Staggered_mass0.5_Nf3", (4, 4, 4, 4), 0.570837085972),
    "quenched_su3_4444": ("confs_HMC_L04040404_beta5.7_quenched_su3", (4, 4, 4, 4), 0.568215750149),
}
out = {}
for name, (d, dims, plaq_survey) in FIX.items():
    U_txt = orc.load_bridgetext(REF / d / "conf_00000100.ildg.txt", dims)
    U_bin = orc.load_ildg(REF / d / "conf_00000100.ildg", dims)
    assert np.array_equal(U_txt, U_bin), name          # text and LIME payload agree bit for bit
    p = orc.plaquette(dims, U_bin)
    assert abs(p - plaq_survey) < 1e-11, (name, p)
    np.save(Path(__file__).parent / f"{name}.npy", U_bin)
    out[name] = {"dims": dims, "plaquette": p, "source": f"test/{d}/conf_00000100.ildg"}
    print(name, p)
(Path(__file__).parent / "fixtures.json").write_text(json.dumps(out, indent=1))
