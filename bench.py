#!/usr/bin/env python
"""Time-to-gradient benchmark (BASELINE.json metric) of the eigd gradient path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--nx NX] [--modes NM]

Workload (config.workload): BASELINE.json configs[1] -- the thermal eigenproblem of
examples/thermal.py at nx = ny = 500 (251 001 DOF), 10 modes, m = 60, sigma = -0.1, IRAM + sibk
(lanczos_guess, rtol 1e-10), tensor derivative, modal thermal-compliance objective; synthetic
design x ~ U(0.3, 1) (breaks the symmetric-pair degeneracy of the uniform square).
A "step" is one pass design -> gradient: filter, material, K/M assembly, numeric LDL^T of
K - sigma M, eigensolve, adjoint right-hand sides, adjoint solve, df/dx, node gather, filter^T.

value   : seconds per gradient with the design already in HBM (CUDA events, sum over K steps / K).
e2e     : seconds per gradient through the reference-facing numpy API: host scipy CSR K, M in,
          host df/dx out, every host<->device copy inside the timed region.
N > 1   : strong scaling -- eigensolve replicated, per-mode adjoint solves sharded round robin,
          element ranges of df/dx sharded (eigd_b200/dist.py); max over ranks.
--impl reference : the CPU path of the reference (oracle port: scipy SuperLU + ARPACK + numpy) on
          the host cores, bounded sample per step (see cpu_reference()).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SIGMA, MLANCZOS = -0.1, 60
# operation counts of the unmodified reference on this exact configuration, measured in the build
# container through oracle/ref_loader.py (thermal nx=ny=500, N=10, m=60)
REF_EIG_SOLVES, REF_ADJ_SOLVES, REF_B_PER_EIG_SOLVE = 61, 214, 3


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--nx", type=int, default=500)
    ap.add_argument("--modes", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-API leg")
    ap.add_argument("--shard", default="designs", choices=["designs", "modes"],
                    help="N > 1: 'designs' = one independent design per GPU (weak scaling, no data-path collective); "
                         "'modes' = one design, per-mode adjoint solves and df/dx element ranges sharded (strong scaling)")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the extra strong-scaling measurement")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
# CPU reference arm (oracle port of the reference path; the only place the product tree runs oracle/)
# ------------------------------------------------------------------------------------------
class CpuReference:
    """Bounded sample of the reference's CPU path at FULL problem size.

    The unmodified reference needs ~80 s per gradient on this workload, so one
    step times a slice of it and scales by the reference's own operation counts:
      setup (timed once, added to every step): SuperLU factorisation of K - sigma M (splu);
      per step: S SuperLU solves, S B-products (the two operations 90 % of the reference's time
      goes to), one full evaluation of the dK/dM einsum callbacks for N modes;
      value = t_factor + (61 + 214) * t_solve + (3*61 + 214 + 2N + m) * t_spmv + t_dfdx.
    Krylov orthogonalisation and Python overhead of the reference are NOT included, so the number
    is a lower bound on the reference's time-to-gradient."""

    def __init__(self, nx, N):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import eigd_oracle as eo
        import fe_oracle as fo
        self.N = N
        conn, X = fo.grid_mesh(nx, nx, 1.0, 1.0)
        self.mdl = fo.Q4Model(conn, X, "thermal")
        rng = np.random.default_rng(0)
        self.rhoE = rng.uniform(0.3, 1.0, self.mdl.nelems)
        self.K, self.M = self.mdl.assemble(self.rhoE)
        t0 = time.perf_counter()
        self.factor = eo.SpLu(self.K - SIGMA * self.M)
        self.t_factor = time.perf_counter() - t0
        self.rng = rng
        self.n = self.K.shape[0]

    def step(self, S=3):
        n, N = self.n, self.N
        b = self.rng.normal(size=n)
        t0 = time.perf_counter()
        for _ in range(S):
            b = self.factor(b)
        t_solve = (time.perf_counter() - t0) / S
        t0 = time.perf_counter()
        for _ in range(S):
            b = self.M @ b
        t_spmv = (time.perf_counter() - t0) / S
        W = self.rng.normal(size=(n, N))
        V = self.rng.normal(size=(n, N))
        t0 = time.perf_counter()
        self.mdl.dK(self.rhoE, W, V)
        self.mdl.dM(self.rhoE, W, V)
        t_dfdx = time.perf_counter() - t0
        nspmv = REF_B_PER_EIG_SOLVE * REF_EIG_SOLVES + REF_ADJ_SOLVES + 2 * N + MLANCZOS
        total = self.t_factor + (REF_EIG_SOLVES + REF_ADJ_SOLVES) * t_solve + nspmv * t_spmv + t_dfdx
        self.last = {"t_factor": self.t_factor, "t_solve": t_solve, "t_spmv": t_spmv, "t_dfdx": t_dfdx}
        return total

    def describe(self, S=3):
        return ("full-size %d-DOF K, M; splu factor timed once (%.2f s); per step %d SuperLU solves, %d B-products, "
                "one dK/dM einsum pass for N=%d, scaled by the reference's measured counts (%d+%d solves); "
                "orthogonalisation and Python overhead excluded (lower bound)"
                % (self.n, self.t_factor, S, S, self.N, REF_EIG_SOLVES, REF_ADJ_SOLVES))


def run_reference(args, rank):
    if rank != 0:
        return
    ref = CpuReference(args.nx, args.modes)
    for _ in range(args.warmup):
        ref.step()
    vals = [ref.step() for _ in range(args.steps)]
    v = float(np.mean(vals))
    line = base_line(args, v, v * 1e3)
    line.update({"impl": "reference", "n_gpus": args.gpus, "gpu_launches": 0,
                 "cpu_baseline": {"value": v, "unit": "s", "cores": 1, "kind": "port", "sample": ref.describe(),
                                  "detail": ref.last},
                 "e2e": {"value": v, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line))


def base_line(args, value, ms):
    n = (args.nx + 1) ** 2
    weak = args.gpus > 1 and args.shard == "designs"
    return {"metric": "time_to_gradient", "value": value, "unit": "s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "weak" if weak else "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "thermal_q4_nx%d_ny%d_%ddof_N%d_m%d_iram_sibk" % (args.nx, args.nx, n, args.modes, MLANCZOS),
                       "baseline_config": "configs[1]: examples/thermal.py scaled to ~250k DOF, 10 modes, single B200",
                       "sigma": SIGMA, "rtol": 1e-10, "deriv_type": "tensor",
                       "l2": "explicit 256 MiB L2 flush between timed steps; per-step working set ~1 GB > 126 MB L2",
                       "parallelism": "1 GPU" if args.gpus == 1 else
                       ("design-batch sweep (BASELINE configs[4] pattern): %d independent designs of the same mesh, one per "
                        "GPU, no data-path collective; value = step time (max over ranks) / %d designs" % (args.gpus, args.gpus)
                        if weak else
                        "one design: eigensolve replicated, per-mode adjoint + element-range dfdx sharded over %d GPUs"
                        % args.gpus)}}


# ------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------
class Clocks:
    """One long-running ``nvidia-smi -lms 200`` process sampling this rank's GPU during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        if os.environ.get("EIGD_BENCH_NO_CLOCKS"):
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        samples = []
        if self.proc is not None:
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                self.proc.kill()
                out = ""
            for ln in out.strip().splitlines():
                f = [x.strip() for x in ln.split(",")]
                if len(f) >= 6:
                    samples.append(f)
        sm = [float(s[0]) for s in samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in samples for i in range(4) if s[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(samples)}


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def run_ours(args, rank, world):
    import torch
    import torch.distributed as dist
    import eigd_b200 as E
    from eigd_b200 import device as D, topo as T, _hostdev as H

    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    D.init("cuda:%d" % local)
    shard = None
    designs = world > 1 and args.shard == "designs"
    if world > 1 and not designs:
        from eigd_b200.dist import ModeSharding
        shard = ModeSharding()
    N = args.modes
    model = T.make_thermal_model(nx=args.nx, ny=args.nx, N=N, m=MLANCZOS, sigma=SIGMA, solver_type="IRAM",
                                 adjoint_method="sibk", adjoint_options={"lanczos_guess": True}, rtol=1e-10,
                                 deriv_type="tensor", seed=0)
    model.sharding = shard
    rng = np.random.default_rng(rank if designs else 0)     # designs mode: rank r evaluates design r
    x_h = rng.uniform(0.3, 1.0, model.nnodes)
    vec_h = np.random.default_rng(12345).uniform(size=model.nnodes)
    x_d, vec_d = D.to_device(x_h), D.to_device(vec_h)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def step_device():
        model.initialize(x=x_d)
        model.initialize_adjoint()
        model.add_thermal_compliance_derivative(1.0, vec_d)
        model.finalize_adjoint()
        return model.xb

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 1)):
        step_device()
    barrier()
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    D.Timeline.reset()
    D.Timeline.enabled = True
    D.solve_timing_begin()                  # native-side CUDA events around every solve launch of the timed region
    l0 = D.launch_count()
    evs = []
    stage = {"eigenvalue solve time": 0.0, "adjoint solution time": 0.0, "total derivative time": 0.0,
             "matrix assembly time": 0.0}
    prof = None
    if os.environ.get("EIGD_BENCH_CPROFILE"):
        import cProfile
        prof = cProfile.Profile()
    for _ in range(args.steps):
        flush.fill_(1)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if prof:
            prof.enable()
        step_device()
        if prof:
            torch.cuda.synchronize()
            prof.disable()
        e1.record()
        evs.append((e0, e1))
        torch.cuda.synchronize()
        for k in stage:
            stage[k] += model.profile[k] / args.steps
    barrier()
    if prof:
        import pstats
        pstats.Stats(prof, stream=sys.stderr).sort_stats("tottime").print_stats(18)
    launches = D.launch_count() - l0
    solve_by_k = D.solve_timing_end()
    D.Timeline.enabled = False
    tl = D.Timeline.summary()
    per_step_ms = [a.elapsed_time(b) for a, b in evs]
    ms = sum(per_step_ms) / args.steps
    # ---- N > 1, designs mode: also time ONE design with the sharded stages (strong scaling) -----------
    strong = None
    if designs and not args.no_strong:
        from eigd_b200.dist import ModeSharding
        sh = ModeSharding()
        x0_d = D.to_device(np.random.default_rng(0).uniform(0.3, 1.0, model.nnodes))   # same design on every rank
        model.sharding = sh

        def step_strong():
            model.initialize(x=x0_d)
            model.initialize_adjoint()
            model.add_thermal_compliance_derivative(1.0, vec_d)
            model.finalize_adjoint()

        for _ in range(2):
            step_strong()
        tms = []
        stage_s = {"eigenvalue solve time": 0.0, "adjoint solution time": 0.0, "total derivative time": 0.0}
        for _ in range(args.steps):
            flush.fill_(1)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step_strong()
            e1.record()
            torch.cuda.synchronize()
            tms.append(e0.elapsed_time(e1))
            for k in stage_s:
                stage_s[k] += model.profile[k] / args.steps
        barrier()
        t = torch.tensor([sum(tms) / len(tms)] + [stage_s[k] for k in sorted(stage_s)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        strong = {"value": float(t[0]) / 1e3, "unit": "s", "scaling": "strong",
                  "parallelism": "one design: eigensolve + factorisation replicated, per-mode adjoint solves (mode i -> "
                                 "rank i mod N, one all-gather) and df/dx element ranges (one all-gather) sharded",
                  "stages_s_max_over_ranks": {k: float(v) for k, v in zip(sorted(stage_s), t[1:].tolist())}}
        model.sharding = None
    # ---- end to end through the numpy API (host CSR in, host df/dx out) ---------------------------
    K_h, M_h = model.K.to_scipy(), model.M.to_scipy()
    mat_h = (K_h - SIGMA * M_h).tocsc()     # the caller's host inputs: K, M and the shifted matrix (thermal.py:288-290)
    for A_h in (K_h, M_h, mat_h):           # their values live in page-locked host memory (bench contract); the
        vals = D.pinned_empty(A_h.data.shape)   # so does the right-hand side array the host code below fills
        vals[...] = A_h.data
        A_h.data = vals
    prob = model.prob

    Phib_h = D.pinned_empty((model.nnodes, N))

    def step_e2e():
        f = E.SpLuOperator(mat_h, coords=model.X, dof_per_node=1)
        s = E.IRAM(N=N, m=MLANCZOS)
        s.seed = 0
        s.sharding = shard
        lam, Phi = s.solve(K_h, M_h, f, SIGMA)
        c = Phi.T @ vec_h                                   # objective seeds on the host, as the example does
        Phib = np.multiply.outer(vec_h, 2.0 * c / lam, out=Phib_h)     # written into page-locked memory
        lamb = -(c * c) / lam**2
        Phib[:, 0], lamb[0] = 0.0, 0.0
        psi, data = s.solve_adjoint(Phib, method="sibk", rtol=1e-10, lanczos_guess=True)
        dfdx = np.zeros(prob.nelems)
        s.add_total_derivative(lamb, Phib, psi, prob.dAdx, prob.dBdx, dfdx, adj_corr_data=data, deriv_type="tensor")
        return dfdx

    e2e_s, h2d, d2h = None, 0, 0
    if not args.no_e2e:
        for _ in range(2):
            step_e2e()
        barrier()
        H.XFER["h2d"] = H.XFER["d2h"] = 0
        ke = max(1, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(ke):
            step_e2e()
        barrier()
        e2e_s = (time.perf_counter() - t0) / ke
        h2d, d2h = H.XFER["h2d"] // ke, H.XFER["d2h"] // ke
    clk = clocks.stop() if rank == 0 else None
    # ---- max over ranks ------------------------------------------------------------------------------
    if world > 1:
        t = torch.tensor([ms, e2e_s or 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1])
    units = world if designs else 1          # gradients produced per step by the whole job
    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # dominant kernel: the persistent triangular-solve kernel; headline = its single-RHS launches (the Lanczos
    # recurrence), the multi-RHS launches of the adjoint solvers are listed beside it
    fac = model.factor.lu
    by_rhs = {}
    tot_ms = 0.0
    for k, (calls, kms) in sorted(solve_by_k.items()):
        byts = fac.solve_bytes(k)
        by_rhs[str(k)] = {"launches_per_step": calls / max(args.steps, 1), "ms_per_launch": kms / calls,
                          "algorithmic_bytes_per_launch": byts, "achieved_gbs": byts / 1e9 / (kms / calls / 1e3)}
        tot_ms += kms
    kdom = max(solve_by_k, key=lambda k: solve_by_k[k][1]) if solve_by_k else 1
    dom = by_rhs.get(str(kdom), {"launches_per_step": 0, "ms_per_launch": 0.0, "algorithmic_bytes_per_launch": 0, "achieved_gbs": 0.0})
    achieved = dom["achieved_gbs"]
    traffic = None
    try:
        ncu = json.load(open(os.path.join(ROOT, "profiles", "r1_solve_kernel_ncu.json")))
        traffic = ncu.get("dram_bytes_per_launch", {}).get(str(kdom))
    except Exception:
        pass
    line = base_line(args, ms / 1e3 / units, ms)
    if strong is not None:
        line["strong_single_gradient"] = strong
    line.update({
        "impl": "ours", "gpu_launches": int(launches // max(args.steps, 1)), "clocks": clk,
        "e2e": {"value": (e2e_s / units) if e2e_s else e2e_s, "unit": "s", "h2d_bytes_per_step": int(h2d) * units,
                "d2h_bytes_per_step": int(d2h) * units,
                "note": "reference-facing numpy API: host scipy K, M, K - sigma*M (values; the shared int32 pattern is uploaded once "
                        "per mesh) and host Phib (written by the host code each step) in page-locked host memory in; host lam, "
                        "Phi, psi, dfdx out (numpy views of page-locked blocks); copies >= 1 MB go through a copy stream"},
        "stages_s": stage, "per_step_ms": per_step_ms,
        "roofline": {"bound": "hbm", "kernel": "multifrontal LDL^T triangular solve, forward + backward sweep, %d right-hand side(s): "
                                               "subtree_kernel<%d> (forward, TMA-staged fronts) + solve_kernel<%d> (persistent cooperative "
                                               "level phases) + subtree_kernel<%d> (backward); 'launch' below = one such solve call"
                                               % (kdom, kdom, kdom, kdom),
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                     "traffic": traffic, "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"],
                     "launches_per_step": dom["launches_per_step"], "ms_per_launch": dom["ms_per_launch"],
                     "share_of_step": tot_ms / (ms * args.steps) if ms else None,
                     "note": "issue/latency-bound, not DRAM-bound: 2 subtree phases (front mode) + 22 level phases separated by grid barriers "
                             "(DESIGN.md section 5 has the ablation)",
                     "by_rhs": by_rhs},
        "timeline_ms_per_step": {k: v["ms"] / args.steps for k, v in tl.items()},
        "counts": {"eig_solves": model.profile["solve preconditioner count"],
                   "adjoint_solves": model.profile["adjoint preconditioner count"],
                   "nnzL": model.symbolic[0].query("nnzL"), "n": model.nvars},
    })
    if world == 1 and not args.no_cpu_baseline:
        ref = CpuReference(args.nx, N)
        v = ref.step()
        line["cpu_baseline"] = {"value": v, "unit": "s", "cores": 1, "kind": "port", "sample": ref.describe(),
                                "detail": ref.last}
    print(json.dumps(line))


def main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        dist.init_process_group("nccl")
    try:
        run_ours(args, rank, world)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
