"""Host-side wall-clock profile of the ADMM loop on the GPU backend (which C-ABI calls dominate)."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from admm_optim_b200 import ug4, _lib
from admm_optim_b200.driver import ObstacleOptim

refs = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ug = ug4.Backend(device=0)
p = ObstacleOptim(ug, 3, numRefs=refs, grid="grids/box_3D_elongated.npz").setup()
p.set_sensitivity(p.synthetic_sensitivity(0.5))
p.begin_step()
p.admm_iteration()
# per C function timing
import collections
acc = collections.defaultdict(lambda: [0, 0.0])
orig = _lib.call
def timed(name, *a):
    t = time.perf_counter(); orig(name, *a); dt = time.perf_counter() - t
    acc[name][0] += 1; acc[name][1] += dt
ug4.call = timed
t0 = time.perf_counter()
n = 3
for _ in range(n):
    p.admm_iteration()
ug.synchronize()
tot = time.perf_counter() - t0
print("total per ADMM iteration: %.1f ms" % (1e3 * tot / n))
for k, (c, t) in sorted(acc.items(), key=lambda kv: -kv[1][1])[:14]:
    print("%-40s calls/it %7.1f   ms/it %8.2f   us/call %8.1f" % (k, c / n, 1e3 * t / n, 1e6 * t / c))
