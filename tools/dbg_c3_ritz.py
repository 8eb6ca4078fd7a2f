import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from eigd_b200 import device as D, topo as T, arpack
D.init()
for P in (1, 2):
    arpack.BLOCK_SIZE = P
    bk = T.make_buckling_model(nx=352, ny=704, N=20, m=60, sigma=3.0, solver_type="IRAM", adjoint_method="sibk",
                               adjoint_options={"lanczos_guess": True}, rtol=1e-12)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        bk.initialize()
    es = bk.eig_solver; st = es.lanczos_state
    th = es.theta; lam = 3.0 * th / (th - 1.0)
    order = np.argsort(-np.abs(th))
    print("P=%d: Ritz values by |theta| (rank: theta, lambda, in indices[:N]?, in returned set?)" % P)
    first = set(es.indices[:20].tolist()); ret = set(st.sel.tolist())
    for r, i in enumerate(order[:30]):
        print("  %2d: theta %12.5f  lam %.8f  %s %s" % (r, th[i], lam[i], "idxN" if i in first else "    ", "ret" if i in ret else ""))
    neg = [(th[i], lam[i]) for i in range(len(th)) if lam[i] < 2.57 and lam[i] > 0]
    print("  Ritz values with 0 < lam < 2.57:", neg)
