#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json): 3D ADMM outer iters/s at fixed DoFs,
plus GMG V-cycle ms and SpMV HBM GB/s on a refinement level larger than L2.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one ADMM outer iteration of 3d_admm.lua:875-1304 (q-step, prox, Newton/Schur loop with six
GMG-preconditioned BiCGStab solves per Newton iteration, dual update, norms) on box_3D_elongated at the
script's default refinement (numRefs = 2, 44 730 deformation DoFs) with a synthetic J' (no Navier-Stokes).

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM; `e2e`: J' uploaded from pinned host memory
and u downloaded every step through the public API.  `roofline`: the dominant kernel (k_bsr_spmv family)
timed with CUDA events on a level whose matrix exceeds L2.  `cpu_baseline` / `--impl reference`: the CPU
oracle port of the same algorithm timed on this box's host cores (the only places oracle/ is executed).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GRID3D = os.path.join(ROOT, "grids", "box_3D_elongated.npz")
METRIC = "3D ADMM outer iters/s at fixed DoFs"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def spmv_bytes(dim, nb, nnzb):
    """B_spmv = nnzb*(8 d^2 + 4) + nb*(4 + 16 d)."""
    return nnzb * (8 * dim * dim + 4) + nb * (4 + 16 * dim)


def vcycle_bytes(dim, levels):
    """B_V = sum_{l>=1} [7 B_spmv(A_l) + 2 B_P(l)], B_P = nnzP*12 + nb_l*4 + 8 d (nb_l + nb_{l-1}), nnzP = nnzb_{l-1}."""
    tot = 0
    for l in range(1, len(levels)):
        nb, nnzb = levels[l]
        nbc, nnzbc = levels[l - 1]
        tot += 7 * spmv_bytes(dim, nb, nnzb) + 2 * (nnzbc * 12 + nb * 4 + 8 * dim * (nb + nbc))
    return tot


# ------------------------------------------------------------------------------------------------
# CPU oracle legs (cpu_baseline / --impl reference)
# ------------------------------------------------------------------------------------------------
def oracle_problem(refs):
    from admm_optim_b200.driver import ObstacleOptim
    from oracle import ug4_np
    ug = ug4_np.Backend(smoother="gs")       # lexicographic Gauss-Seidel: what the reference's descriptor asks for (u3:16)
    p = ObstacleOptim(ug, 3, numRefs=refs, grid=GRID3D).setup()
    p.set_sensitivity(p.synthetic_sensitivity(0.5))
    p.begin_step()
    return p


def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np  # noqa: F401
    p = oracle_problem(args.refs)
    for _ in range(args.warmup):
        p.admm_iteration()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rec = p.admm_iteration()
        assert rec is not None, "oracle ADMM iteration failed"
    dt = time.perf_counter() - t0
    v = args.steps / dt
    cores = cpu_cores()
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "iters/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "3d_admm.lua ADMM loop on box_3D_elongated.ugx, numRefs=%d, synthetic J'" % args.refs, "numRefs": args.refs,
                       "note": "UG4 is not installable here; this is the CPU oracle port (NumPy/SciPy, lexicographic GS V(3,3), SuperLU base solve)"},
            "cpu_baseline": {"value": v, "unit": "iters/s", "cores": cores, "kind": "port",
                             "sample": "%d full ADMM iterations (NumPy/SciPy oracle; BLAS/SuperLU threads as available)" % args.steps},
            "e2e": {"value": v, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def cpu_baseline_sample(refs):
    p = oracle_problem(refs)
    t0 = time.perf_counter()
    rec = p.admm_iteration()
    dt = time.perf_counter() - t0
    assert rec is not None
    return {"value": 1.0 / dt, "unit": "iters/s", "cores": cpu_cores(), "kind": "port",
            "sample": "1 ADMM iteration (first of the loop, %d Newton its) of the same workload, NumPy/SciPy oracle with lexicographic GS" % len(rec["newton"]),
            "newton_iterations": len(rec["newton"]),
            "bicgstab_iterations_first_newton": rec["newton"][0]["its"]}


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from admm_optim_b200 import ug4
    from admm_optim_b200.driver import ObstacleOptim

    stream = torch.cuda.Stream()
    ug = ug4.Backend(device=local, stream=stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # multi-GPU (this revision): the path shards by domain decomposition; until the halo layer lands each
    # rank runs the whole problem as an independent replica and the aggregate is reported as weak scaling.
    prob = ObstacleOptim(ug, 3, numRefs=args.refs, grid=GRID3D).setup()
    ndofs = prob.DeformationSpace_ApproxSpace.num_dofs()
    J_host = torch.from_numpy(prob.synthetic_sensitivity(0.5)).pin_memory()
    u_host = torch.empty(ndofs, dtype=torch.float64).pin_memory()

    def timed_leg(e2e):
        prob.set_sensitivity(J_host.numpy())
        prob.begin_step()
        for _ in range(args.warmup):
            assert prob.admm_iteration() is not None
        newton, its = 0, 0
        total_ms = 0.0
        launches0 = ug.launch_count()
        barrier()
        for _ in range(args.steps):
            with torch.cuda.stream(stream):
                flush.zero_()                                     # L2 flush between timed iterations (untimed)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            if e2e:
                prob.set_sensitivity(J_host.numpy())             # H2D from pinned memory
            rec = prob.admm_iteration()
            if e2e:
                prob.u.to_numpy(u_host.numpy())                  # D2H of the step's result
            e1.record(stream)
            e1.synchronize()
            assert rec is not None, "ADMM iteration failed"
            total_ms += e0.elapsed_time(e1)
            newton += len(rec["newton"])
            its += sum(n["its"]["rhs"] + n["its"]["large"] + sum(n["its"]["B"]) for n in rec["newton"])
        barrier()
        launches = ug.launch_count() - launches0
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), launches, newton, its, rec

    clocks = ClockSampler(local)
    clocks.start()
    ms_dev, launches, newton, its, rec = timed_leg(False)
    ms_e2e, _, _, _, _ = timed_leg(True)
    clk = clocks.stop()

    # ---- roofline leg: SpMV / V-cycle on a level larger than L2 ---------------------------------
    roof, extra = None, {}
    if rank == 0 and args.roofline_refs > 0:
        big = ObstacleOptim(ug, 3, numRefs=args.roofline_refs, grid=GRID3D).setup()
        DD = big.DeformationEquation_DomainDisc
        DD.assemble_jacobian(big.A_u_Hessian, big.u)
        _, nb, nnzb = big.A_u_Hessian.info()
        n = nb * 3
        x = np.random.default_rng(1).standard_normal(n)
        big.sigma.from_numpy(x)
        DD.adjust_solution(big.sigma)
        for _ in range(3):
            big.A_u_Hessian.apply(big.Lu, big.sigma)
        reps = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(reps):
            big.A_u_Hessian.apply(big.Lu, big.sigma)
        e1.record(stream)
        e1.synchronize()
        t_spmv = e0.elapsed_time(e1) / reps * 1e-3
        bytes_spmv = spmv_bytes(3, nb, nnzb)
        peak, peak_src = measured_peak()
        ach = bytes_spmv / t_spmv / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                "kernel": "k_bsr_spmv_tma<3,0,0,3> (y = A x, BSR 3x3 fp64, TMA-staged tiles)", "peak_source": peak_src,
                "bytes_per_launch": bytes_spmv, "us_per_launch": t_spmv * 1e6,
                "workload": "box_3D_elongated numRefs=%d: %d block rows, %d blocks (matrix %.2f GB > L2)" % (args.roofline_refs, nb, nnzb, nnzb * 76 / 1e9)}
        s = big.SmallProblemRHS_Solver
        s.init(big.A_u_Hessian, big.sigma)
        levels = [s.level_info(l) for l in range(args.roofline_refs + 1)]
        for _ in range(2):
            s.vcycle(big.delta_u, big.sigma)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(10):
            s.vcycle(big.delta_u, big.sigma)
        e1.record(stream)
        e1.synchronize()
        t_v = e0.elapsed_time(e1) / 10 * 1e-3
        bv = vcycle_bytes(3, levels)
        extra = {"spmv_gbs": ach, "vcycle_ms": t_v * 1e3, "vcycle_gbs": bv / t_v / 1e9, "vcycle_frac": bv / t_v / 1e9 / peak,
                 "vcycle_bytes": bv, "roofline_levels": levels}
        del big

    if rank != 0:
        return
    value = world * args.steps / (ms_dev * 1e-3)
    e2e_v = world * args.steps / (ms_e2e * 1e-3)
    line = {"metric": METRIC, "value": value, "unit": "iters/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "3d_admm.lua ADMM loop (3d_admm.lua:875-1304) on box_3D_elongated.ugx, numRefs=%d, %d deformation DoFs, synthetic J'" % (args.refs, ndofs),
                       "numRefs": args.refs, "dofs": ndofs, "parallelism": "replicas x%d" % world if world > 1 else "1 GPU",
                       "l2": "L2 flushed (256 MB write) between timed iterations; working set itself is L2-sized",
                       "smoother": "Chebyshev(3)-Jacobi (stated equivalent of the reference's sequential GS, DESIGN.md)",
                       "newton_its_per_step": newton / args.steps, "bicgstab_its_per_step": its / args.steps},
            "e2e": {"value": e2e_v, "unit": "iters/s", "h2d_bytes_per_step": ndofs * 8, "d2h_bytes_per_step": ndofs * 8 + 8 * 16,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clk}
    if roof:
        line["roofline"] = roof
        line.update(extra)
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_sample(args.refs)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--refs", type=int, default=2, help="numRefs of the ADMM workload (3d_admm.lua:46 default 2)")
    ap.add_argument("--roofline-refs", type=int, default=4, help="refinement level of the SpMV / V-cycle roofline leg (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
