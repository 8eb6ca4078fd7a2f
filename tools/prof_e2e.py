"""Developer profile of the numpy-API (end-to-end) gradient step of bench.py: cProfile of the host side."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import eigd_b200 as E
from eigd_b200 import device as D, topo as T
D.init()
N, SIGMA = 10, -0.1
model = T.make_thermal_model(nx=500, ny=500, N=N, m=60, sigma=SIGMA, solver_type="IRAM", adjoint_method="sibk",
                             adjoint_options={"lanczos_guess": True}, rtol=1e-10, deriv_type="tensor", seed=0)
x_d = D.to_device(np.random.default_rng(0).uniform(0.3, 1.0, model.nnodes))
vec_h = np.random.default_rng(12345).uniform(size=model.nnodes)
model.initialize(x=x_d)
K_h, M_h = model.K.to_scipy(), model.M.to_scipy()
mat_h = (K_h - SIGMA * M_h).tocsc()
if "--pinned" in sys.argv:
    for A_h in (K_h, M_h, mat_h):
        vals = D.pinned_empty(A_h.data.shape); vals[...] = A_h.data; A_h.data = vals
prob = model.prob

Phib_pin = D.pinned_empty((model.nnodes, N)) if "--pinned" in sys.argv else None
STAMPS = {}


def step():
    t = [time.perf_counter()]
    f = E.SpLuOperator(mat_h, coords=model.X, dof_per_node=1); t.append(time.perf_counter())
    s = E.IRAM(N=N, m=60); s.seed = 0
    lam, Phi = s.solve(K_h, M_h, f, SIGMA); t.append(time.perf_counter())
    c = Phi.T @ vec_h
    Phib = np.multiply.outer(vec_h, 2.0 * c / lam, out=Phib_pin); lamb = -(c * c) / lam**2
    Phib[:, 0], lamb[0] = 0.0, 0.0; t.append(time.perf_counter())
    psi, data = s.solve_adjoint(Phib, method="sibk", rtol=1e-10, lanczos_guess=True); t.append(time.perf_counter())
    dfdx = np.zeros(prob.nelems)
    s.add_total_derivative(lamb, Phib, psi, prob.dAdx, prob.dBdx, dfdx, adj_corr_data=data, deriv_type="tensor")
    t.append(time.perf_counter())
    for k, a, b in zip(("SpLuOperator", "solve", "host Phib", "solve_adjoint", "add_total_derivative"), t, t[1:]):
        STAMPS[k] = STAMPS.get(k, 0.0) + (b - a)
    return dfdx

for _ in range(3): step()
D.COPY_STATS.clear(); STAMPS.clear()
D.Timeline.reset(); D.Timeline.enabled = True; D.solve_timing_begin()
t0 = time.perf_counter()
for _ in range(3): step()
print("e2e step %.1f ms" % ((time.perf_counter() - t0) / 3 * 1e3))
D.Timeline.enabled = False
print("  GPU time by entry point (CUDA events, ms/step):", {k: round(v["ms"] / 3, 2) for k, v in D.Timeline.summary().items()})
print("  solves by k (calls/step, ms/call):", {k: (v[0] / 3, round(v[1] / max(v[0], 1), 4)) for k, v in D.solve_timing_end().items()})
for k, (c, sec, nb) in D.COPY_STATS.items():
    print("  %-52s %5.1f calls/step %7.2f ms/step %7.1f MB/step" % (k, c / 3, sec / 3 * 1e3, nb / 3 / 1e6))
print("  host wall clock per call: " + ", ".join("%s %.2f ms" % (k, v / 3 * 1e3) for k, v in STAMPS.items()))
if "--no-cprofile" in sys.argv: sys.exit(0)
pr = cProfile.Profile(); pr.enable()
for _ in range(3): step()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
