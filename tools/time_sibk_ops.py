"""Developer A/B timing (same box, L2 flushed) of the fused sibk step kernels against the launches they replace."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eigd_b200 import device as D
D.init()
n, k = 251001, 10
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
g = torch.Generator(device="cuda"); g.manual_seed(0)
U = torch.randn(n, k, dtype=torch.float64, device="cuda", generator=g) / n**0.5
V = torch.randn(n, k, dtype=torch.float64, device="cuda", generator=g)
X0 = torch.randn(n, k, dtype=torch.float64, device="cuda", generator=g)
Ws = [torch.randn(n, k, dtype=torch.float64, device="cuda", generator=g) / n**0.5 for _ in range(8)]
H = D.zeros(10, k)


def t(fn, name, reps=12):
    ts = []
    for _ in range(reps):
        X = X0.clone()
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(X); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print("%-58s min %6.1f us  median %6.1f us" % (name, ts[0], ts[len(ts) // 2]), flush=True)


def loop_sweep(X, j):
    for i in range(j):
        D.col_dot(X, Ws[i], out=H[i])
        D.col_axpy(X, H[i], Ws[i], sign=-1.0)


for j in (1, 4, 8):
    t(lambda X: loop_sweep(X, j), "MGS against %d blocks, dot / axpy launches" % j)
    t(lambda X: D.mgs_sweep(X, Ws[:j], [H[i] for i in range(j)]), "MGS against %d blocks, one cooperative launch" % j)
