"""Developer microbenchmark of the host <-> device copy paths behind the numpy API (20 MB operands)."""
import time
import numpy as np
import torch

n = 251001 * 10
a = np.random.rand(n)
src = torch.from_numpy(a)
pin = torch.empty(n, dtype=torch.float64, pin_memory=True)
pin_np = pin.numpy()
dev = torch.empty(n, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()


def t(f, name, reps=10):
    f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print("%-46s %7.2f ms  %6.1f GB/s" % (name, dt * 1e3, n * 8 / dt / 1e9), flush=True)


print("torch threads", torch.get_num_threads())
t(lambda: torch.as_tensor(a, device="cuda"), "as_tensor pageable -> device")
t(lambda: np.copyto(pin_np, a), "np.copyto pageable -> pinned")
for nt in (16, 8, 4, 2, 1):
    torch.set_num_threads(nt)
    t(lambda: pin.copy_(src), "torch copy_ pageable -> pinned (%d threads)" % nt)
torch.set_num_threads(16)
t(lambda: dev.copy_(pin, non_blocking=True), "pinned -> device (async + sync at end)")
t(lambda: pin.copy_(dev, non_blocking=True), "device -> pinned (async + sync at end)")
t(lambda: dev.cpu(), "device -> pageable (.cpu())")
b = np.empty(n)
t(lambda: np.copyto(b, a), "np.copyto pageable -> pageable")


def fresh():
    p = torch.empty(n, dtype=torch.float64, pin_memory=True)
    p.copy_(src)
    return p.to("cuda", non_blocking=True)


t(fresh, "cached pinned block + copy_ + async upload")


# ---- the pattern of device.h2d / d2h, piece by piece ------------------------------------------------
def pieces(sync_each, reps=8):
    acc = {"empty": 0.0, "copy_": 0.0, "to": 0.0}
    for _ in range(reps):
        t0 = time.perf_counter()
        p = torch.empty(n, dtype=torch.float64, pin_memory=True)
        t1 = time.perf_counter()
        p.copy_(src)
        t2 = time.perf_counter()
        d = p.to("cuda", non_blocking=True)
        t3 = time.perf_counter()
        acc["empty"] += t1 - t0
        acc["copy_"] += t2 - t1
        acc["to"] += t3 - t2
        del p
        if sync_each:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    print("h2d pieces (sync each iteration: %s): " % sync_each + ", ".join("%s %.2f ms" % (k, v / reps * 1e3) for k, v in acc.items()), flush=True)


pieces(True)
pieces(False)
pieces(True)

# d2h while the GPU is busy: is copy_(non_blocking) into a pinned block really asynchronous?
big = torch.empty(1 << 28, dtype=torch.float64, device="cuda")     # 2 GiB fill ~ 0.6 ms each
for sync_each in (True, False):
    acc = {"empty": 0.0, "copy_": 0.0, "wait": 0.0}
    for _ in range(8):
        for _ in range(10):
            big.fill_(1.0)
        t0 = time.perf_counter()
        st = torch.empty_like(dev, device="cpu", pin_memory=True)
        t1 = time.perf_counter()
        st.copy_(dev, non_blocking=True)
        t2 = time.perf_counter()
        ev = torch.cuda.Event()
        ev.record()
        ev.synchronize()
        t3 = time.perf_counter()
        acc["empty"] += t1 - t0
        acc["copy_"] += t2 - t1
        acc["wait"] += t3 - t2
        out = st.numpy()
        del st
    print("d2h pieces behind ~6 ms of GPU work: " + ", ".join("%s %.2f ms" % (k, v / 8 * 1e3) for k, v in acc.items()), "pinned:", torch.from_numpy(out).is_pinned(), flush=True)
