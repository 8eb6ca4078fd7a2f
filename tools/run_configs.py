"""Full-size runs of the BASELINE.json configurations that are not the bench line:

    python tools/run_configs.py c3 [--nx 352 --ny 704 --modes 20]     buckling, ~497k DOF, 20 modes
    python tools/run_configs.py c5 [--designs 8]                      design sweep, 202k-DOF natural-frequency mesh
    python tools/run_configs.py c4 [--nx 408 --ny 408 --modes 20 --fd]   1M-DOF synthetic shell (6 dof/node), crm.py driver

Each prints one JSON line: stage times (CUDA-synchronised wall clock of the driver's own timers, as the
reference examples report them), solve counts, factor statistics, and size-independent acceptance checks
(eigen-residuals, B-orthonormality, adjoint residual of the converged psi) -- the reference cannot be run at
these sizes inside the GPU job, so parity at full size is by those properties (tests/test_fullsize_gpu.py does
the same for the bench configuration)."""
import argparse
import json
import os
import sys
import time
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def now():
    import torch
    torch.cuda.synchronize()
    return time.perf_counter()


def residuals(model, A, B, lam, Phi_d, mode):
    """max_i ||L_i phi_i|| / (||A phi_i|| + |lam_i| ||B phi_i||) and max |Phi^T B Phi - I| on the device."""
    from eigd_b200 import device as D
    import torch
    AP, BP = A.spmm(Phi_d), B.spmm(Phi_d)
    lam_d = torch.as_tensor(np.asarray(lam), device=Phi_d.device)
    R = (AP - BP * lam_d) if mode == "normal" else (BP + AP * lam_d)
    res = (R.norm(dim=0) / (AP.norm(dim=0) + (BP * lam_d).norm(dim=0))).max().item()      # relative to |A phi| + |lam B phi|
    G = D.gemm_tn(Phi_d, BP).cpu().numpy()
    return res, float(np.abs(G - np.eye(G.shape[0])).max())


def setup_dist():
    """torchrun launch: one process per GPU, per-mode adjoint shards (eigd_b200/dist.py)."""
    import torch
    import torch.distributed as dist
    from eigd_b200 import device as D
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    D.init("cuda:%d" % local)
    if world == 1:
        return None, 0, 1
    dist.init_process_group("nccl")
    from eigd_b200.dist import ModeSharding
    return ModeSharding(), dist.get_rank(), world


def run_c3(args):
    from eigd_b200 import device as D, topo as T
    shard, rank, world = setup_dist()
    t0 = now()
    model = T.make_buckling_model(nx=args.nx, ny=args.ny, N=args.modes, m=60, sigma=3.0, solver_type="IRAM",
                                  adjoint_method="sibk", adjoint_options={"lanczos_guess": True}, rtol=1e-10,
                                  deriv_type="tensor")
    t_setup = now() - t0
    model.sharding = shard
    out = {"config": "C3 buckling nx=%d ny=%d" % (args.nx, args.ny), "n": int(model.prob.nred), "nnz": int(model.prob.nnz),
           "N": args.modes, "host_setup_s": t_setup, "n_gpus": world,
           "parallelism": "1 GPU" if world == 1 else "eigensolve replicated, adjoint modes i -> rank i mod %d" % world}
    rng = np.random.default_rng(0)
    reps = []
    for rep in range(args.reps):
        l0 = D.launch_count()
        ta = now()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            model.initialize()
        model.initialize_adjoint()
        if rep == 0:                                                # the dof of largest |phi_1| (non-degenerate objective)
            node = int(np.argmax(np.abs(model.prob.full_vector(model.Qr[:, 0].contiguous()).cpu().numpy())))
        h = model.add_eigenvector_aggregate_derivative(1.0, 100.0, node, mode="tanh")
        model.finalize_adjoint()
        tb = now()
        p = model.profile
        reps.append({"wall_s": tb - ta, "assembly_s": p["matrix assembly time"], "eig_s": p["eigenvalue solve time"],
                     "adjoint_s": p["adjoint solution time"], "dfdx_s": p["total derivative time"],
                     "time_to_gradient_s": model.time_to_gradient(), "eig_solves": p["solve preconditioner count"],
                     "adjoint_solves": p["adjoint preconditioner count"], "launches": D.launch_count() - l0,
                     "symbolic_s": p.get("symbolic analysis time")})
    out["reps"] = reps
    out["BLF"] = [float(v) for v in model.BLF]
    out["factor_info"] = model.factor.info
    out["refine_steps"] = model.factor.refine
    out["probe_residual_unrefined"] = model.factor._probe_residual()
    out["symbolic"] = model.symbolic[0].stats()
    res, orth = residuals(model, model.Gr, model.Kr, model.BLF, model.Qr, "buckling")
    out["eigen_residual_max"] = res
    out["orthonormality_defect"] = orth
    ar, ao = model.eig_solver.eval_adjoint_residual_norm(model.Qrb, model.psir, b_ortho=True)
    rhs = float(model.Qrb.norm(dim=0).max().item())
    out["adjoint_residual_over_rtol_rhs"] = float(np.max(ar)) / (1e-10 * rhs)      # <= O(1): the solver's own acceptance level
    assert out["adjoint_residual_over_rtol_rhs"] < 50.0, out["adjoint_residual_over_rtol_rhs"]
    assert res < 1e-8 and orth < 1e-10, (res, orth)
    out["aggregate_h"] = h
    out["xb_norm"] = float(model.xb.norm().item())
    return out


def run_c4(args):
    """BASELINE configs[3]: CRM-scale synthetic shell, 409 x 409 nodes x 6 DOF = 1 003 686 DOF (1 001 232 free), 20 modes, IRAM +
    sibk, modal compliance, per-mode "vector" total derivative over 400 component thicknesses (driver of examples/crm.py).
    The reference cannot run at this size inside the job (and needs TACS for its own model): acceptance = normalised
    eigen / adjoint residuals and a finite-difference check of the gradient, as SURVEY.md 8d prescribes for C4."""
    from eigd_b200 import device as D, shell as S
    shard, rank, world = setup_dist()
    t0 = now()
    model = S.make_shell_model(nx=args.nx, ny=args.ny, ncx=20, ncy=20, N=args.modes, m=60, omega0=10.0, solver_type="IRAM",
                               adjoint_method="sibk", adjoint_options={"lanczos_guess": True}, rtol=1e-10, deriv_type="vector")
    t_setup = now() - t0
    model.sharding = shard
    x0 = np.random.default_rng(0).uniform(0.6, 1.4, model.prob.ncomp)
    out = {"config": "C4 synthetic shell nx=%d ny=%d (6 dof/node, cylindrical panel, clamped edge)" % (args.nx, args.ny),
           "n": int(model.prob.ndof), "nnz": int(model.prob.nnz), "N": args.modes, "ncomp": int(model.prob.ncomp),
           "host_setup_s": t_setup, "n_gpus": world,
           "parallelism": "1 GPU" if world == 1 else "factorisation + eigensolve replicated, adjoint modes and per-mode derivative "
                          "calls i -> rank i mod %d (one packed all-gather, one all-reduce)" % world}
    reps = []
    for rep in range(args.reps):
        model.set_design_vars(x0)
        l0 = D.launch_count()
        ta = now()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            model.initialize()
        model.initialize_adjoint()
        model.add_compliance_derivative()
        model.finalize_adjoint()
        tb = now()
        p = model.profile
        reps.append({"wall_s": tb - ta, "assembly_s": p["matrix assembly time"], "eig_s": p["eigenvalue solve time"],
                     "adjoint_s": p["adjoint solution time"], "dfdx_s": p["total derivative time"],
                     "time_to_gradient_s": model.time_to_gradient(), "eig_solves": p["solve preconditioner count"],
                     "adjoint_solves": p["adjoint preconditioner count"], "launches": D.launch_count() - l0,
                     "symbolic_s": p.get("symbolic analysis time")})
    out["reps"] = reps
    out["lam_first"] = [float(v) for v in model.lam[:4]]
    out["factor_info"] = model.factor.info
    out["symbolic"] = model.symbolic[0].stats()
    res, orth = residuals(model, model.Kr, model.Mr, model.lam, model.Q, "normal")
    out["eigen_residual_max_rel"] = res
    out["orthonormality_defect"] = orth
    ar, ao = model.eig_solver.eval_adjoint_residual_norm(model.Qb, model.psi, b_ortho=True)
    rhs = float(model.Qb.norm(dim=0).max().item())
    out["adjoint_residual_over_rtol_rhs"] = float(np.max(ar)) / (1e-10 * rhs)
    grad = model.grad.cpu().numpy().copy()
    out["grad_norm"] = float(np.linalg.norm(grad))
    if args.fd:
        pert = np.random.default_rng(3).uniform(size=x0.shape)
        h = 1e-4
        vals = []
        for sgn in (1.0, -1.0):
            model.set_design_vars(x0 + sgn * h * pert)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                model.initialize()
            vals.append(model.get_compliance())
        fd = (vals[0] - vals[1]) / (2 * h)
        out["fd_directional"] = fd
        out["adjoint_directional"] = float(grad @ pert)
        out["fd_rel_err"] = abs(fd - float(grad @ pert)) / abs(fd)
    return out


def run_nf(args, tag):
    from eigd_b200 import device as D, topo as T
    D.init()
    t0 = now()
    model = T.make_natural_frequency_model(nx=args.nx, ny=args.ny, Lx=2.0, Ly=1.0, N=args.modes, m=60, sigma=-10.0,
                                           solver_type="IRAM", adjoint_method="sibk",
                                           adjoint_options={"lanczos_guess": True}, rtol=1e-10, deriv_type="tensor")
    t_setup = now() - t0
    out = {"config": "%s natural frequency nx=%d ny=%d" % (tag, args.nx, args.ny), "n": int(model.nvars), "N": args.modes,
           "host_setup_s": t_setup}
    w = np.random.default_rng(99).normal(size=(model.nvars, args.modes))
    w_d = D.to_device(w)
    reps = []
    for b in range(args.designs):
        x = np.random.default_rng(b).uniform(0.3, 1.0, model.fltr.num_design_vars)
        l0 = D.launch_count()
        ta = now()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            model.initialize(x=x)
        model.initialize_adjoint()
        model.add_modal_function_derivative(w_d)
        model.finalize_adjoint()
        tb = now()
        p = model.profile
        reps.append({"design": b, "wall_s": tb - ta, "assembly_s": p["matrix assembly time"], "eig_s": p["eigenvalue solve time"],
                     "adjoint_s": p["adjoint solution time"], "dfdx_s": p["total derivative time"],
                     "time_to_gradient_s": model.time_to_gradient(), "eig_solves": p["solve preconditioner count"],
                     "adjoint_solves": p["adjoint preconditioner count"], "launches": D.launch_count() - l0,
                     "lam_first": [float(v) for v in model.lam[:3]]})
    out["reps"] = reps
    out["factor_info"] = model.factor.info
    out["symbolic"] = model.symbolic[0].stats()
    res, orth = residuals(model, model.K, model.M, model.lam0, model.Q0, "normal")
    out["eigen_residual_max"] = res
    out["orthonormality_defect"] = orth
    ar, ao = model.eig_solver.eval_adjoint_residual_norm(model.Q0b, model.psi0, b_ortho=True)
    rhs = float(model.Q0b.norm(dim=0).max().item())
    out["adjoint_residual_over_rtol_rhs"] = float(np.max(ar)) / (1e-10 * rhs)
    assert out["adjoint_residual_over_rtol_rhs"] < 50.0 and orth < 1e-10, (out["adjoint_residual_over_rtol_rhs"], orth)
    out["xb_norm"] = float(model.xb.norm().item())
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["c3", "c4", "c4plate", "c5"])
    ap.add_argument("--fd", action="store_true", help="c4: add the finite-difference check of the gradient")
    ap.add_argument("--nx", type=int, default=None)
    ap.add_argument("--ny", type=int, default=None)
    ap.add_argument("--modes", type=int, default=None)
    ap.add_argument("--designs", type=int, default=None)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    if a.config == "c3":
        a.nx, a.ny, a.modes = a.nx or 352, a.ny or 704, a.modes or 20
        res = run_c3(a)
    elif a.config == "c5":
        a.nx, a.ny, a.modes, a.designs = a.nx or 448, a.ny or 224, a.modes or 6, a.designs or 8
        res = run_nf(a, "C5 design sweep")
    elif a.config == "c4":
        a.nx, a.ny, a.modes = a.nx or 408, a.ny or 408, a.modes or 20
        res = run_c4(a)
    else:
        a.nx, a.ny, a.modes, a.designs = a.nx or 1000, a.ny or 500, a.modes or 20, a.designs or 2
        res = run_nf(a, "round-1 C4 stand-in (1M-DOF 2-dof plate)")
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps(res))
