"""Dry run of bench.py's B200 arm WITHOUT a GPU: torch.cuda and the CUDA backend are replaced by stand-ins (wall-clock "events",
the NumPy oracle behind the ug4.Backend surface) so that the control flow of the arm -- every leg, every key of the JSON line --
executes on a CPU box.  It checks the Python of the arm, nothing about the kernels or the numbers.
    python tools/bench_dryrun.py [bench.py arguments, e.g. --refs 1 --roofline-refs 1 --admm-refs 1 --dim2-refs 2]"""
import contextlib
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from oracle import ug4_np


# ---- torch.cuda stand-ins ------------------------------------------------------------------------------------------------
class _Event:
    def __init__(self, enable_timing=False): self.t = None
    def record(self, stream=None): self.t = time.perf_counter()
    def synchronize(self): pass
    def elapsed_time(self, other): return (other.t - self.t) * 1e3


class _Stream:
    cuda_stream = 0


torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda d: None
torch.cuda.synchronize = lambda *a: None
torch.cuda.Stream = _Stream
torch.cuda.Event = _Event
torch.cuda.stream = lambda s: contextlib.nullcontext()
_empty, _tensor = torch.empty, torch.tensor
torch.empty = lambda *a, **k: _empty(*a, **{**k, "device": "cpu"})
torch.tensor = lambda *a, **k: _tensor(*a, **{**k, "device": "cpu"})
torch.Tensor.pin_memory = lambda self: self


# ---- the ug4.Backend surface on top of the oracle ------------------------------------------------------------------------
class _Desc:
    """attribute view of the oracle solver's descriptor dict (bench.py writes s.desc.verbose / s.desc.abs_tol)"""
    def __init__(self, d): object.__setattr__(self, "d", d)
    def __setattr__(self, k, v):
        self.d["convCheck"][{"abs_tol": "absolute", "verbose": "verbose"}[k]] = v
    def __getitem__(self, k): return self.d[k]
    def get(self, k, default=None): return self.d.get(k, default)


class _Solver(ug4_np.BiCGStabGMG):
    def __init__(self, ug, desc):
        super().__init__(ug, desc)
        self.desc = _Desc(desc)


DECOMPOSED = os.environ.get("DRYRUN_DECOMPOSED") == "1"      # pretend the larger legs are domain-decomposed (multi-rank control flow)


class _Domain(ug4_np.Domain):
    @property
    def decomposed(self):       # only the legs above the headline size pretend, and only on the "distributed" backend
        return DECOMPOSED and getattr(self, "_backend_distributed", False) and self.top.nv > 1000
    @property
    def _dist(self):
        return dict(gather_level=0) if self.decomposed else None
    def p2p_status(self): return dict(connected=False, error=0)
    def get_level(self, level, elems=True):
        l = self.levels[level]
        return dict(xyz=l.xyz, elems=l.elems)


class FakeBackend(ug4_np.Backend):
    name = "dryrun"
    def __init__(self, device=0, stream=None, distributed=False):
        super().__init__(smoother="cheb", threads=1, fast_assembly=True, c_solver=True)   # compiled oracle paths: the dry runs stay short
        self.rank, self.nranks = (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))) if distributed else (0, 1)
        self._distributed = bool(distributed)
        self.util.solver.CreateSolver = lambda desc: _Solver(self, desc)
        self._launches = 0
    def Domain(self):
        d = _Domain()
        d._backend_distributed = self._distributed
        return d
    def launch_count(self):
        self._launches += 1
        return self._launches
    def set_tuning(self, *a): pass


class _Op(ug4_np.AssembledLinearOperator):
    def info(self): return (self.dd.space.dom.dim, self.dd.space.dom.top.nv, 0)


FakeBackend.AssembledLinearOperator = lambda self, dd: _Op(dd)

import admm_optim_b200.ug4 as ug4   # noqa: E402  (host-only import: the shared library is not touched)
ug4.Backend = FakeBackend

import torch.distributed as _dist   # noqa: E402
_init = _dist.init_process_group
_dist.init_process_group = lambda backend=None, **kw: _init("gloo")      # the multi-rank flow on CPU: gloo instead of nccl

import bench   # noqa: E402

FAIL_LEG = os.environ.get("DRYRUN_FAIL_IN_2D_LEG")      # "raise": an exception inside the last leg; "hang": it never returns
if FAIL_LEG:
    _L2 = FakeBackend.L2Norm

    def _l2(self, gf, cmp, *a):
        if gf.space.dom.dim == 2 and gf.space.dom.top.nv > 1000:
            if FAIL_LEG == "hang":
                time.sleep(3600)
            raise RuntimeError("injected failure in the 2D leg")
        return _L2(self, gf, cmp, *a)
    FakeBackend.L2Norm = _l2

bench.golden_parity.__defaults__ = ("3d_refs1",)      # the smaller golden trace: the oracle stands in for the GPU here

if __name__ == "__main__":
    sys.argv = [sys.argv[0]] + (sys.argv[1:] or ["--refs", "1", "--roofline-refs", "1", "--admm-refs", "1", "--dim2-refs", "2", "--steps", "1", "--warmup", "1", "--no-cpu-baseline"])
    bench.main()
