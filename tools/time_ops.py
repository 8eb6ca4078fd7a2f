"""Developer micro-timing of individual device ops at benchmark size."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from eigd_b200 import device as D
D.init()
n = 251001
m = 61
def bench(name, fn, reps=20):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("%-40s host %.1f us  total %.1f us" % (name, (t1 - t0) / reps * 1e6, (t2 - t0) / reps * 1e6))
Vt = torch.randn(m, n, dtype=torch.float64, device="cuda")
w = torch.randn(n, dtype=torch.float64, device="cuda")
h = D.zeros(m)
X10 = torch.randn(n, 10, dtype=torch.float64, device="cuda")
Y10 = torch.randn(n, 10, dtype=torch.float64, device="cuda")
s10 = torch.rand(10, dtype=torch.float64, device="cuda") + 0.5
for j in (1, 10, 30, 60):
    Vj = Vt[:j].T
    bench("gemm_tn vecmajor j=%d" % j, lambda: D.gemm_tn(Vj, w, out=h[:j].unsqueeze(1)))
    bench("gemm_nn vecmajor j=%d" % j, lambda: D.gemm_nn(Vj, h[:j].unsqueeze(1), w, alpha=-1e-9, beta=1.0))
bench("col_dot k=1", lambda: D.col_dot(w, Vt[0], out=h[:1]))
bench("col_scale k=1", lambda: D.col_scale(Vt[1], h[1:2], mode=0))
bench("copy_ row", lambda: Vt[2].copy_(w))
bench("axpby 1", lambda: D.axpby(1.0, h[0:1], 1.0, h[1:2], out=h[2:3]))
bench("col_dot k=10", lambda: D.col_dot(X10, Y10))
bench("col_axpy k=10", lambda: D.col_axpy(X10, s10, Y10, sign=-1e-9))
bench("gemm_tn rowmajor 10x10", lambda: D.gemm_tn(X10, Y10))
S = torch.randn(10, 10, dtype=torch.float64, device="cuda") * 1e-9
bench("gemm_nn rowmajor 10x10", lambda: D.gemm_nn(X10, S, Y10, alpha=1.0, beta=1.0))
bench("project 10", lambda: D.project(X10, X10, Y10))
bench("torch.empty", lambda: D.empty(n, 10))
bench("clone", lambda: X10.clone())
