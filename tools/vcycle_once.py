"""Minimal V-cycle driver for ncu: build the level-`refs` hierarchy and run a few V(3,3) cycles (smoother = k_bsr_spmv_tma<D,2,0,U>).
    ncu --set full -k regex:k_bsr_spmv_tma --launch-skip N -c 1 python tools/vcycle_once.py 5"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from admm_optim_b200 import ug4  # noqa: E402
from admm_optim_b200.driver import ObstacleOptim  # noqa: E402
import bench  # noqa: E402

refs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ug = ug4.Backend(device=0)
big = ObstacleOptim(ug, 3, numRefs=refs, grid=bench.GRID3D).setup()
DD = big.DeformationEquation_DomainDisc
DD.assemble_jacobian(big.A_u_Hessian, big.u)
_, nb, nnzb = big.A_u_Hessian.info()
big.sigma.from_numpy(np.random.default_rng(1).standard_normal(nb * 3))
DD.adjust_solution(big.sigma)
s = big.SmallProblemRHS_Solver
s.init(big.A_u_Hessian, big.sigma)
n0 = ug.launch_count()
for _ in range(reps):
    s.vcycle(big.delta_u, big.sigma)
ug.synchronize()
print("done", nb, nnzb, "launches per cycle", (ug.launch_count() - n0) // reps)
