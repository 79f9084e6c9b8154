"""Minimal SpMV driver for ncu: assemble the level-`refs` Hessian and launch the SpMV a few times."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from admm_optim_b200 import ug4
from admm_optim_b200.driver import ObstacleOptim
import bench
refs = int(sys.argv[1]); variant = int(sys.argv[2]); waves = int(sys.argv[3]); reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
ug = ug4.Backend(device=0)
ug.set_tuning("spmv_variant", variant); ug.set_tuning("spmv_waves", waves)
big = ObstacleOptim(ug, 3, numRefs=refs, grid=bench.GRID3D).setup()
DD = big.DeformationEquation_DomainDisc
DD.assemble_jacobian(big.A_u_Hessian, big.u)
_, nb, nnzb = big.A_u_Hessian.info()
big.sigma.from_numpy(np.random.default_rng(1).standard_normal(nb * 3))
for _ in range(reps):
    big.A_u_Hessian.apply(big.Lu, big.sigma)
ug.synchronize()
print("done", nb, nnzb)
