"""Where does one ADMM iteration go?  Wraps every C-ABI entry point (admm_optim_b200._lib.call) with a synchronising
wall-clock timer and prints the per-entry totals of one iteration, next to the un-instrumented iteration time with the
BiCGStab iteration graph on and off.

    python tools/iter_profile.py [numRefs] [dim] [graph,pdl,coarse_variant,spmv_variant ...]
"""
import collections
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from admm_optim_b200 import ug4  # noqa: E402
from admm_optim_b200.driver import ObstacleOptim  # noqa: E402

refs = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 3
grid = "grids/box_3D_elongated.npz" if dim == 3 else "grids/refined.npz"

import torch  # noqa: E402

stream = torch.cuda.Stream()
ug = ug4.Backend(device=0, stream=stream.cuda_stream)
p = ObstacleOptim(ug, dim, numRefs=refs, grid=grid).setup()
p.set_sensitivity(p.synthetic_sensitivity(0.5))
p.begin_step()
for _ in range(3):
    assert p.admm_iteration() is not None


def timed_iterations(k=5):
    ug.synchronize()
    l0 = ug.launch_count()
    t = time.perf_counter()
    for _ in range(k):
        rec = p.admm_iteration()
        assert rec is not None
    ug.synchronize()
    dt = (time.perf_counter() - t) / k
    its = sum(n["its"]["rhs"] + n["its"]["large"] + sum(n["its"]["B"]) for n in rec["newton"])
    return dt, (ug.launch_count() - l0) / k, len(rec["newton"]), its


CONFIGS = [tuple(int(x) for x in a.split(",")) for a in sys.argv[3:]] or [(1, 1, 0, 0), (1, 0, 0, 0), (0, 1, 0, 0), (0, 0, 0, 0), (1, 1, 1, 0), (1, 1, 0, 1), (1, 1, 0, 0)]
for g, pdl, cv, sv in CONFIGS:
    ug.set_tuning("graph", g)
    ug.set_tuning("pdl", pdl)
    ug.set_tuning("coarse_variant", cv)
    ug.set_tuning("loop", 0 if sv >= 1000 else 1)            # spmv_variant + 1000: device-side BiCGStab loop (conditional graph node) off
    sv = sv % 1000
    ug.set_tuning("spmv_variant", sv % 10)
    ug.set_tuning("tma_small_ctas", 1 if sv >= 100 else 2)   # spmv_variant + 100: one persistent SpMV CTA per SM on small levels
    p.admm_iteration()
    dt, launches, nn, its = timed_iterations()
    print("graph=%d pdl=%d coarse_variant=%d spmv_variant=%d: %.2f ms / ADMM iteration (%.1f us per BiCGStab it incl. everything), %.0f launches, %d Newton its, %d BiCGStab its (last)" %
          (g, pdl, cv, sv, dt * 1e3, dt * 1e6 / max(its, 1), launches, nn, its))

# ---- per-entry attribution (synchronising: adds overhead, read the SHARES) -------------------------------------------
acc = collections.defaultdict(lambda: [0, 0.0])
orig_call = ug4.call


def timed_call(name, *args):
    ug.lib.ab_context_synchronize(ug.ctx)
    t = time.perf_counter()
    r = orig_call(name, *args)
    ug.lib.ab_context_synchronize(ug.ctx)
    a = acc[name]
    a[0] += 1
    a[1] += time.perf_counter() - t
    return r


ug4.call = timed_call
t = time.perf_counter()
rec = p.admm_iteration()
ug.synchronize()
total = time.perf_counter() - t
ug4.call = orig_call
print("instrumented iteration: %.2f ms wall" % (total * 1e3))
tsum = sum(v[1] for v in acc.values())
for name, (cnt, sec) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print("  %-36s %5d calls %9.3f ms  %5.1f %%" % (name, cnt, sec * 1e3, 100 * sec / tsum))
print("  %-36s %15s %9.3f ms (python + ctypes outside the calls: %.3f ms)" % ("sum", "", tsum * 1e3, (total - tsum) * 1e3))

os.environ["ADMM_B200_TRACE"] = "1"
print("---- setup trace of one more iteration (ADMM_B200_TRACE=1) ----", flush=True)
p.admm_iteration()
ug.synchronize()
