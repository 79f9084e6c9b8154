"""A/B of the device-side BiCGStab loop (conditional WHILE graph node) against per-iteration graph replay: the same solve
repeated, wall clock per solve.   python tools/loop_ab.py [numRefs]"""
import os, sys, time
os.environ.setdefault("ADMM_B200_TRACE", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from admm_optim_b200 import ug4
from admm_optim_b200.driver import ObstacleOptim
refs = int(sys.argv[1]) if len(sys.argv) > 1 else 2
stream = torch.cuda.Stream()
ug = ug4.Backend(device=0, stream=stream.cuda_stream)
p = ObstacleOptim(ug, 3, numRefs=refs, grid="grids/box_3D_elongated.npz").setup()
DD = p.DeformationEquation_DomainDisc
DD.assemble_jacobian(p.A_u_Hessian, p.u)
b = np.random.default_rng(1).standard_normal(p.DeformationSpace_ApproxSpace.num_dofs())
p.Lu.from_numpy(b, 2)
DD.adjust_solution(p.Lu)
s = p.SmallProblemRHS_Solver
os.environ["ADMM_B200_TRACE"] = "1"
s.init(p.A_u_Hessian, p.sigma)
p.sigma.set(0.0); assert s.apply(p.sigma, p.Lu)
os.environ["ADMM_B200_TRACE"] = "0"
for loop in (1, 0, 1, 0):
    ug.set_tuning("loop", loop)
    p.sigma.set(0.0); s.apply(p.sigma, p.Lu)
    ug.synchronize()
    t = time.perf_counter()
    for _ in range(40):
        p.sigma.set(0.0)
        assert s.apply(p.sigma, p.Lu)
    ug.synchronize()
    dt = (time.perf_counter() - t) / 40
    print("loop=%d: %.1f us per solve, %d iterations -> %.1f us per iteration (incl. solve prologue)" % (loop, dt * 1e6, s.step(), dt * 1e6 / s.step()))
