"""quick device timing of the Dslash kernels (development aid; bench.py is the contract)."""
import ctypes as C, sys, os
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "latticeqcd.jl_b200")]
import lqcd_b200 as q
from lqcd_b200 import _lib as L
sizes = [(16,)*4, (32,)*4] if len(sys.argv) < 2 else [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for dims in sizes:
    ctx = q.get_context(dims)
    ctx.call("lqcd_gauge_random", 111, -1.0)
    for kind, name, bps, fps in [(L.WILSON, "wilson", 960, 1368), (L.STAGGERED, "stag", 672, 582)]:
        x, y = q.FermionField(ctx, kind), q.FermionField(ctx, kind)
        q.gauss_distribution_fermion_(x, 1)
        op = L.LqcdOp(); op.kind = kind; op.kappa = 0.12; op.r = 1.0; op.mass = 0.1
        for i, b in enumerate([1, 1, 1, -1]): op.bc[i] = b
        for flush in (0, 1):
            mean, mn = C.c_double(), C.c_double()
            ctx.call("lqcd_time_dslash", C.byref(op), y.h, x.h, 0, 5, flush, C.byref(mean), C.byref(mn))
            ctx.call("lqcd_time_dslash", C.byref(op), y.h, x.h, 0, 20, flush, C.byref(mean), C.byref(mn))
            V = dims[0]*dims[1]*dims[2]*dims[3]
            print(f"{name} {dims} flush={flush} mean {mean.value*1e3:8.1f} us min {mn.value*1e3:8.1f} us  "
                  f"{bps*V/mean.value/1e6:8.1f} GB/s  {fps*V/mean.value/1e6:8.1f} GFLOP/s  (WPC={os.environ.get('LQCD_WPC','4')} TILE={os.environ.get('LQCD_TILE','auto')})", flush=True)
