"""Developer check: finite-difference directional derivative of the C3 aggregate against the two adjoint answers."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from eigd_b200 import device as D, topo as T, arpack
D.init()
arpack.BLOCK_SIZE = int(sys.argv[1]) if len(sys.argv) > 1 else 1
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "fullsize_c3.npz"))
node = int(g["node"])
bk = T.make_buckling_model(nx=352, ny=704, N=20, m=60, sigma=3.0, solver_type="IRAM", adjoint_method="sibk",
                           adjoint_options={"lanczos_guess": True}, rtol=1e-12)
pert = np.random.default_rng(777).uniform(size=bk.x.shape)
x0 = bk.x.copy()
def h_of(x):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        bk.initialize(x=x)
    return bk.get_eigenvector_aggregate(100.0, node, mode="tanh"), np.asarray(bk.BLF).copy()
h0, blf0 = h_of(x0)
print("h0 %.12f  golden mean q^2 %.12f" % (h0, float(np.mean(g["qnode"] ** 2))))
for eps in (1e-5, 1e-6, 1e-7):
    hp, bp = h_of(x0 + eps * pert)
    hm, bm = h_of(x0 - eps * pert)
    print("eps %.0e: FD dh %.8f   dBLF1 FD %.8f  max |dBLF| %.3e" % (eps, (hp - hm) / (2 * eps), (bp[0] - bm[0]) / (2 * eps), np.abs(bp - bm).max()))
print("reference pert.xb %.8f ; guess-free adjoint gave -3.65414033" % float(g["pert_dot_xb"]))
