#!/usr/bin/env python
"""Generate tests/golden/admm_trace_*.json with the CPU oracle (oracle/ug4_np.py driven by the script replay).

The reference ships no golden vectors and cannot run here, so these pins come from the
oracle restatement itself: they freeze the per-ADMM-iteration scalars the drivers write to __ADMMStats_step_*.txt
(3d_admm.lua:1265-1276) plus the Newton / Lagrange-multiplier trace, for regression and for the GPU parity tests.
Run:  python tools/make_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from admm_optim_b200.driver import ObstacleOptim
from oracle import ug4_np

# the last two are the scripts' DEFAULT refinements (BASELINE.json configs[1] / configs[0]: 44 730 / 18 016 deformation DoFs)
CASES = {"3d_refs1": (3, "box_3D_elongated.npz", 1), "2d_refs2": (2, "refined.npz", 2),
         "3d_refs2": (3, "box_3D_elongated.npz", 2), "2d_refs3": (2, "refined.npz", 3)}
only = sys.argv[1:]
for name, (dim, grid, refs) in CASES.items():
    if only and name not in only:
        continue
    p = ObstacleOptim(ug4_np.Backend(smoother="cheb"), dim, numRefs=refs, grid=os.path.join(ROOT, "grids", grid), admmSteps=2).setup()
    for s in [p.SmallProblemRHS_Solver, p.LargeProblem_Solver] + p.B_Solver:
        s.desc["convCheck"]["absolute"] = 1e-13
    p.set_sensitivity(p.synthetic_sensitivity(0.5))
    tr = p.run_admm()
    u = p.u.to_numpy()
    out = dict(dim=dim, grid=grid, numRefs=refs, amplitude=0.5, abs_tol=1e-13, reference_volume=p.ReferenceVolume,
               admm=[dict(u_diff=r["u_diff"], lambda_inc=r["lambda_inc"], max_norm=r["max_norm"], Lambda=[float(x) for x in r["Lambda"]],
                          L_lambda=[float(x) for x in r["L_lambda"]], newton_its=len(r["newton"]),
                          delta_lambda=[n["delta_lambda"] for n in r["newton"]]) for r in tr],
               u_l2=float(np.linalg.norm(u)), u_sum=float(u.sum()), u_absmax=float(np.abs(u).max()),
               u_probe=[float(x) for x in u[:: max(1, len(u) // 16)][:16]])
    path = os.path.join(ROOT, "tests", "golden", "admm_trace_%s.json" % name)
    json.dump(out, open(path, "w"), indent=1)
    print(path, out["admm"][0]["u_diff"], out["admm"][1]["u_diff"])


# ---- trace files in the reference's format, written by the driver replay on the oracle ----------------
# tests/test_host.py::test_trace_files_match_committed_golden replays this on the CPU and diffs the files byte for byte.
if not only or "traces" in only:
    import shutil
    d = os.path.join(ROOT, "tests", "golden", "traces_3d_refs1")
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    p = ObstacleOptim(ug4_np.Backend(smoother="cheb"), 3, numRefs=1, grid=os.path.join(ROOT, "grids", "box_3D_elongated.npz"), admmSteps=3,
                      trace_dir=d, newton_output=True).setup()
    p.set_sensitivity(p.synthetic_sensitivity(0.5))
    p.run_admm()
    print(d, sorted(os.listdir(d)))
