import sys, os, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from conftest import load_golden, align_signs
import eigd_b200 as E
from eigd_b200 import device
device.init()
g = load_golden("buckling_basiclanczos")
rel = lambda a, b: np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300)
A, B, sigma = g["A"], g["B"], float(g["sigma"])
f = E.SpLuOperator((B + sigma * A).tocsc())
print("factor info", f.info, "refine", f.refine)
for name, s in (("basic", E.BasicLanczos(N=5, m=24, tol=1e-14, mode="buckling")), ("iram", E.IRAM(N=5, m=24, mode="buckling"))):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        lam, Phi = s.solve(A, B, f, sigma)
    Pa, sgn = align_signs(Phi, g["Phi"])
    res = [np.linalg.norm(B @ Phi[:, i] + lam[i] * (A @ Phi[:, i])) / np.linalg.norm(B @ Phi[:, i]) for i in range(5)]
    print(name, "lam rel", rel(lam, g["lam"]), "Phi rel", rel(Pa, g["Phi"]), "orth", np.abs(Phi.T @ (B @ Phi) - np.eye(5)).max(), "eig res", np.array(res))
    psi, data = s.solve_adjoint(g["Phib"] * sgn, method="sibk", rtol=1e-12)
    print("   psi rel", rel(psi * sgn, g["psi_sibk"]), "corr", data)
    r = s.eval_adjoint_residual_norm(g["Phib"] * sgn, psi)
    print("   adjoint residual", r)
