import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from admm_optim_b200 import ug4
from admm_optim_b200.driver import ObstacleOptim
refs = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ug = ug4.Backend(device=0)
p = ObstacleOptim(ug, 3, numRefs=refs, grid="grids/box_3D_elongated.npz").setup()
p.set_sensitivity(p.synthetic_sensitivity(0.5))
p.begin_step()
p.admm_iteration()
os.environ["ADMM_B200_TRACE"] = "1"
t = time.perf_counter(); p.admm_iteration(); ug.synchronize(); print("iteration wall %.1f ms" % (1e3 * (time.perf_counter() - t)))
