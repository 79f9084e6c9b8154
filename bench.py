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
GRID2D = os.path.join(ROOT, "grids", "refined.npz")
METRIC = "3D ADMM outer iters/s at fixed DoFs"
# dram__bytes_read.sum + dram__bytes_write.sum per launch of k_bsr_spmv_tma<3,0,0,3> from the committed `ncu --set full`
# captures (one GPU): numRefs=5 profiles/r01b_spmv_tma_L5_hint_raw.csv (current kernel, L2 evict-first hint on the matrix stream;
# without the hint 8.70 GB + 0.17 GB, r01b_spmv_tma_L5_raw.csv), numRefs=4 profiles/r01_spmv_tma_raw.csv; other sizes -> null
NCU_TRAFFIC_BYTES = {5: 8517273000 + 150910464, 4: 1059398000 + 24325888}


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


def global_counts(refs, dim=3):
    """(block rows, blocks) = (V, V + 2E) of every level of the GLOBAL hierarchy from the regular-refinement recurrences
    V' = V+E, E' = 2E+3F(+T), F' = 4F(+8T), T' = 8T seeded with the level-0 entity counts of the grid."""
    import numpy as np
    z = np.load(GRID3D if dim == 3 else GRID2D)
    el = z["elems"]
    V = len(z["xyz"])
    if dim == 3:
        T = len(el)
        edges = np.unique(np.sort(np.concatenate([el[:, [i, j]] for i, j in ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3))]), axis=1), axis=0)
        faces = np.unique(np.sort(np.concatenate([el[:, list(f)] for f in ((0, 1, 2), (0, 1, 3), (0, 2, 3), (1, 2, 3))]), axis=1), axis=0)
        E, F = len(edges), len(faces)
    else:
        T, F = 0, len(el)
        E = len(np.unique(np.sort(np.concatenate([el[:, [i, j]] for i, j in ((0, 1), (1, 2), (0, 2))]), axis=1), axis=0))
    out = [(V, V + 2 * E)]
    for _ in range(refs):
        V, E, F, T = V + E, 2 * E + 3 * F + T, 4 * F + 8 * T, 8 * T
        out.append((V, V + 2 * E))
    return out


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
def oracle_threads(refs):
    """Host threads the CPU port uses: the C kernels (GS sweep, SpMV) fork per call, which only pays off on big levels."""
    return 1 if refs <= 2 else max(1, min(cpu_cores(), 16))


def oracle_problem(refs):
    from admm_optim_b200.driver import ObstacleOptim
    from oracle import ug4_np
    # Gauss-Seidel is what the reference's descriptor asks for (u3:16): sequential lexicographic on one thread,
    # block-Jacobi across threads otherwise (UG4's behaviour under mpirun)
    # fast_assembly: the element loops of the P1 assembly run in C (oracle/oracle_kernels.c) -- with NumPy assembly two thirds of
    # the CPU iteration were temporaries, which no compiled reference would pay
    ug = ug4_np.Backend(smoother="gs", threads=oracle_threads(refs), fast_assembly=True)
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
    cores = oracle_threads(args.refs)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "iters/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "3d_admm.lua ADMM loop on box_3D_elongated.ugx, numRefs=%d, synthetic J'" % args.refs, "numRefs": args.refs,
                       "note": "UG4 is not installable here; this is the CPU oracle port (NumPy/SciPy + C kernels for the P1 assembly, the Gauss-Seidel sweep and SpMV, V(3,3), SuperLU base solve)",
                       "host_cores_available": cpu_cores()},
            "cpu_baseline": {"value": v, "unit": "iters/s", "cores": cores, "kind": "port",
                             "sample": "%d full ADMM iterations (NumPy/SciPy oracle with C kernels for assembly / GS / SpMV, %d thread(s))" % (args.steps, cores)},
            "e2e": {"value": v, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def cpu_baseline_sample(refs):
    p = oracle_problem(refs)
    t0 = time.perf_counter()
    rec = p.admm_iteration()
    dt = time.perf_counter() - t0
    assert rec is not None
    return {"value": 1.0 / dt, "unit": "iters/s", "cores": oracle_threads(refs), "kind": "port", "host_cores_available": cpu_cores(),
            "sample": "1 ADMM iteration (first of the loop, %d Newton its) of the same workload, NumPy/SciPy oracle with C kernels for the P1 assembly, "
                      "the lexicographic GS sweep and the SpMV" % len(rec["newton"]),
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
    ug = ug4.Backend(device=local, stream=stream.cuda_stream, distributed=world > 1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # multi-GPU: the SAME global problem is domain-decomposed over the ranks (strong scaling): level-0 elements are
    # partitioned (RCB), every rank refines its sub-grid, interface sums + all-reduces run on NCCL
    prob = ObstacleOptim(ug, 3, numRefs=args.refs, grid=GRID3D).setup()
    ndofs_local = prob.DeformationSpace_ApproxSpace.num_dofs()
    J_host = torch.from_numpy(prob.synthetic_sensitivity(0.5)).pin_memory()
    u_host = torch.empty(ndofs_local, dtype=torch.float64).pin_memory()
    ndofs = global_counts(args.refs)[-1][0] * 3 if world > 1 else ndofs_local

    def timed_leg(e2e):
        prob.set_sensitivity(J_host.numpy())
        prob.begin_step()
        for _ in range(args.warmup):
            assert prob.admm_iteration() is not None
        newton, its = 0, 0
        total_ms = 0.0
        h2d = d2h = 0
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
                h2d = ndofs_local * 8
            rec = prob.admm_iteration()
            if e2e:
                prob.u.to_numpy(u_host.numpy())                  # D2H of the step's result
                d2h = ndofs_local * 8 + 8 * 16
            e1.record(stream)
            e1.synchronize()
            assert rec is not None, "ADMM iteration failed"
            total_ms += e0.elapsed_time(e1)
            newton += len(rec["newton"])
            its += sum(n["its"]["rhs"] + n["its"]["large"] + sum(n["its"]["B"]) for n in rec["newton"])
        barrier()
        launches = ug.launch_count() - launches0
        t = torch.tensor([total_ms, float(h2d), float(d2h)], dtype=torch.float64, device="cuda")
        if world > 1:
            tm = t.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t[0] = tm[0]
        return float(t[0].item()), launches, newton, its, rec, int(t[1].item()), int(t[2].item())

    clocks = ClockSampler(local)
    clocks.start()
    ms_dev, launches, newton, its, rec, _, _ = timed_leg(False)
    ms_e2e, _, _, _, _, h2d_bytes, d2h_bytes = timed_leg(True)
    clk = clocks.stop()

    # ---- roofline leg: SpMV / V-cycle on a level larger than L2 ---------------------------------
    roof, extra = None, {}
    if args.roofline_refs > 0:
        big = ObstacleOptim(ug, 3, numRefs=args.roofline_refs, grid=GRID3D).setup()
        DD = big.DeformationEquation_DomainDisc
        DD.assemble_jacobian(big.A_u_Hessian, big.u)
        levels = global_counts(args.roofline_refs) if world > 1 else None
        _, nb_loc, nnzb_loc = big.A_u_Hessian.info()
        nb, nnzb = (levels[-1] if world > 1 else (nb_loc, nnzb_loc))
        x = np.random.default_rng(1 + rank).standard_normal(nb_loc * 3)
        big.sigma.from_numpy(x)
        DD.adjust_solution(big.sigma)

        def timeit(fn, reps):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record(stream)
            for _ in range(reps):
                fn()
            e1.record(stream)
            e1.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / reps * 1e-3], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        t_spmv = timeit(lambda: big.A_u_Hessian.apply(big.Lu, big.sigma), 20)
        bytes_spmv = spmv_bytes(3, nb, nnzb)
        peak, peak_src = measured_peak()
        ach = bytes_spmv / t_spmv / 1e9
        roof = {"bound": "hbm", "achieved": ach / world, "peak": peak, "unit": "GB/s", "frac": ach / world / peak,
                "traffic": NCU_TRAFFIC_BYTES.get(args.roofline_refs) if world == 1 else None,
                "kernel": "k_bsr_spmv_tma<3,0,0,3> (y = A x, BSR 3x3 fp64, TMA-staged tiles)", "peak_source": peak_src,
                "bytes_per_launch": bytes_spmv // world, "us_per_launch": t_spmv * 1e6,
                "workload": "box_3D_elongated numRefs=%d: %d block rows, %d blocks (matrix %.2f GB > L2)%s" %
                            (args.roofline_refs, nb, nnzb, nnzb * 76 / 1e9, " over %d GPUs; per-GPU figures" % world if world > 1 else "")}
        s = big.SmallProblemRHS_Solver
        s.init(big.A_u_Hessian, big.sigma)
        if levels is None:
            levels = [s.level_info(l) for l in range(args.roofline_refs + 1)]
        t_v = timeit(lambda: s.vcycle(big.delta_u, big.sigma), 10)
        bv = vcycle_bytes(3, levels)
        # one full solve on the big level (GMG-preconditioned BiCGStab to the script tolerance); the first call allocates the
        # Krylov workspace and captures the iteration graph (one-off), the second one is timed
        big.Lu.from_numpy(x, 2)
        DD.adjust_solution(big.Lu)
        for _ in range(2):
            big.sigma.set(0.0)
            barrier()
            t0 = time.perf_counter()
            ok = s.apply(big.sigma, big.Lu)
            ug.synchronize()
            t_solve = time.perf_counter() - t0
        extra = {"spmv_gbs": ach, "vcycle_ms": t_v * 1e3, "vcycle_gbs": bv / t_v / 1e9, "vcycle_frac": bv / t_v / 1e9 / (peak * world),
                 "vcycle_bytes": bv, "roofline_levels": levels, "solve_ms": t_solve * 1e3, "solve_its": s.step(), "solve_converged": bool(ok)}
        if world > 1:
            st = big.dom.p2p_status()
            assert st["error"] == 0, "peer-to-peer interface exchange timed out (error %d)" % st["error"]
            extra["decomposed"] = bool(big.dom.decomposed)
            extra["gather_level"] = big.dom._dist["gather_level"] if big.dom.decomposed else None
        del big

    if rank != 0:
        return
    value = args.steps / (ms_dev * 1e-3)
    e2e_v = args.steps / (ms_e2e * 1e-3)
    line = {"metric": METRIC, "value": value, "unit": "iters/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "3d_admm.lua ADMM loop (3d_admm.lua:875-1304) on box_3D_elongated.ugx, numRefs=%d, %d deformation DoFs, synthetic J'" % (args.refs, ndofs),
                       "numRefs": args.refs, "dofs": ndofs,
                       "parallelism": ("domain decomposition x%d (RCB of the level-0 grid, NCCL interface sums + all-reduces)" % world) if world > 1 else "1 GPU",
                       "l2": "L2 flushed (256 MB write) between timed iterations; working set itself is L2-sized",
                       "smoother": "Chebyshev(3)-Jacobi (stated equivalent of the reference's sequential GS, DESIGN.md)",
                       "newton_its_per_step": newton / args.steps, "bicgstab_its_per_step": its / args.steps,
                       "scaling_note": "value at N>1 = the SAME 44 730-DoF problem domain-decomposed (latency-bound, SURVEY 8e); the strong-scaling "
                                       "numbers that matter are spmv_gbs / vcycle_ms / solve_ms at roofline numRefs and profiles/r01_scaling.md "
                                       "(numRefs 6: 7.4x from 1 to 8 GPUs)"},
            "e2e": {"value": e2e_v, "unit": "iters/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clk}
    if roof:
        line["roofline"] = roof
        line.update(extra)
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline_sample(args.refs)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--refs", type=int, default=2, help="numRefs of the ADMM workload (3d_admm.lua:46 default 2)")
    ap.add_argument("--roofline-refs", type=int, default=5, help="refinement level of the SpMV / V-cycle / solve roofline leg (0 = skip); 5 = 20.3 M DoFs, 7.6 GB matrix")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
