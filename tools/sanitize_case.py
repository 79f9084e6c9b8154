"""Small end-to-end case for compute-sanitizer: one ADMM iteration (3D numRefs=1 and 2D numRefs=2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from admm_optim_b200 import ug4
from admm_optim_b200.driver import ObstacleOptim
ug = ug4.Backend(device=0)
for dim, grid, refs in ((3, "grids/box_3D_elongated.npz", 1), (2, "grids/refined.npz", 2)):
    p = ObstacleOptim(ug, dim, numRefs=refs, grid=grid, admmSteps=1).setup()
    p.set_sensitivity(p.synthetic_sensitivity(0.5))
    tr = p.run_admm()
    assert tr and not p.p_solver_failure
    print("dim", dim, "newton its", len(tr[0]["newton"]), "u_diff", tr[0]["u_diff"])
ug.synchronize()
print("SANITIZE CASE DONE")
