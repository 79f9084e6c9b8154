"""Phase timing of the cluster tail kernel (ADMM_B200_TAIL_PROF=1, graphs off so that every launch is a stream launch)."""
import os, sys
os.environ["ADMM_B200_TAIL_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from admm_optim_b200 import ug4
from admm_optim_b200.driver import ObstacleOptim
stream = torch.cuda.Stream()
ug = ug4.Backend(device=0, stream=stream.cuda_stream)
ug.set_tuning("graph", 0)
p = ObstacleOptim(ug, 3, numRefs=2, grid="grids/box_3D_elongated.npz").setup()
DD = p.DeformationEquation_DomainDisc
DD.assemble_jacobian(p.A_u_Hessian, p.u)
p.Lu.from_numpy(np.random.default_rng(1).standard_normal(p.DeformationSpace_ApproxSpace.num_dofs()), 2)
DD.adjust_solution(p.Lu)
s = p.SmallProblemRHS_Solver
s.init(p.A_u_Hessian, p.sigma)
for _ in range(8):
    s.vcycle(p.sigma, p.Lu)
ug.synchronize()
