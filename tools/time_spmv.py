"""Developer timing of the CSR SpMV for the two stencils of the examples (9 and 18 non-zeros per row)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from eigd_b200 import device as D, fe
D.init()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for kind, nx in (("thermal", 500), ("plane_stress", 352)):
    conn, X = fe.grid_mesh(nx, nx, 1.0, 1.0)
    prob = fe.Q4Problem(conn, X, kind)
    prob.set_density(rhoE=np.full(prob.nelems, 0.7))
    K, M = prob.assemble()
    x = torch.randn(K.shape[0], dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    ts = []
    for it in range(10):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); M.spmm(x, out=y); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    by = K.nnz * 12 + K.shape[0] * 20
    print("%s n=%d nnz=%d  G=%s  %.1f us  %.0f GB/s" % (kind, K.shape[0], K.nnz, os.environ.get("EIGD_SPMV_G", "4"), min(ts) * 1e3, by / min(ts) / 1e6))
    for k in (2, 10, 20):
        xk = torch.randn(K.shape[0], k, dtype=torch.float64, device="cuda")
        yk = torch.empty_like(xk)
        ts = []
        for it in range(10):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); M.spmm(xk, out=yk); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        by = K.nnz * 12 + K.shape[0] * (4 + 16 * k)
        print("   spmm k=%d  %.1f us  %.0f GB/s" % (k, min(ts) * 1e3, by / min(ts) / 1e6))
