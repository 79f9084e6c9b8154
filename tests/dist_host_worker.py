"""Worker of test_host.py::test_partition_and_interfaces_gloo_world_size_2 -- run as one of N gloo ranks (CPU only).
Exercises the host side of the multi-GPU path: RCB partition, sub-grid extraction, native refinement of the local
sub-grid, candidate masks, coordinate matching over torch.distributed, the C-ABI interface registration and the
vertical-interface maps of the agglomerated coarse levels (local -> global vertices / matrix blocks on the gather level)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch.distributed as dist

from admm_optim_b200 import partition as P
from admm_optim_b200 import ug4

grid, refs = sys.argv[1], int(sys.argv[2])
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
def gather(obj):
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out


ug = ug4.Backend.host_only(rank, world, gather)      # host-only: no GPU context
dom = ug4.Domain(ug)
ug.LoadDomain(dom, grid)
ug._create_regular_hierarchy(dom, refs)        # decides: undivided, or refine + match + ab_domain_set_interface + gather maps
if not dom.decomposed:
    out = dict(decomposed=False, levels=[dom.level_info(l)["nv"] for l in range(dom.num_levels())])
    allres = gather(out)
    if rank == 0:
        print(json.dumps(allres))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0)
res = []
for level in range(refs + 1):
    lv = dom.get_level(level, elems=False)
    I = dom._iface[level]
    shared = {int(q): lv["xyz"][I["idx"][I["offsets"][k]:I["offsets"][k + 1]]].tolist() for k, q in enumerate(I["neigh"])}
    res.append(dict(nv=int(lv["xyz"].shape[0]), owned=int(I["owned"].sum()), shared=shared))
# vertical interface: every rank's element-incidence counts per local block, summed through gpos, must give the global counts
G = dom._gather
lg = G["level"]
ll, gl = dom.get_level(lg), dom._cdom.get_level(lg)
nvl, nvg = len(ll["xyz"]), G["nv_global"]
def incidence(elems, nv):
    el = np.asarray(elems, np.int64)
    n = el.shape[1]
    keys = np.concatenate([el[:, a] * nv + el[:, b] for a in range(n) for b in range(n)])
    return np.unique(keys, return_counts=True)
lk, lc = incidence(ll["elems"], nvl)
assert np.array_equal(lk, P.pattern_keys(ll["elems"], nvl)) and len(lk) == len(G["gpos"])
# the C++ pattern (what the device matrices use) has exactly this order
nn = C.c_int64()
ug4.call("ab_domain_level_pattern", dom.h, lg, C.byref(nn), None, None)
rp, ci = np.empty(nvl + 1, np.int32), np.empty(nn.value, np.int32)
ug4.call("ab_domain_level_pattern", dom.h, lg, C.byref(nn), rp.ctypes.data_as(C.POINTER(C.c_int32)), ci.ctypes.data_as(C.POINTER(C.c_int32)))
rows = np.repeat(np.arange(nvl, dtype=np.int64), np.diff(rp))
assert np.array_equal(rows * nvl + ci, lk), "numpy pattern order differs from build_pattern"
# the product derives gpos from the two native patterns; the element-based NumPy derivation must give the same positions
assert np.array_equal(G["gpos"], P.block_positions(ll["elems"], nvl, G["l2g"], P.pattern_keys(gl["elems"], nvg), nvg))
parts = gather((G["gpos"], lc, G["l2g"], ll["xyz"]))
ok = True
if rank == 0:
    gk, gc = incidence(gl["elems"], nvg)
    tot = np.zeros(len(gk), np.int64)
    for gpos, cnt, l2g, xyz in parts:
        assert len(np.unique(gpos)) == len(gpos)                   # injective per rank
        np.add.at(tot, gpos, cnt)
        assert np.array_equal(gl["xyz"][l2g], xyz)                 # vertex map hits identical coordinates
    ok = bool(np.array_equal(tot, gc))
    covered = np.zeros(nvg, bool)
    for _, _, l2g, _ in parts:
        covered[l2g] = True
    ok = ok and bool(covered.all())
# shared matrix blocks of the top level (exact Gershgorin bound, partition.match_blocks): a NumPy twin of the device steps of
# lib.cu gmg_setup_kernels -- pack the shared blocks, add the neighbours' parts, replace |own part| by |sum| / mult in the local
# row sums, interface-sum the rows -- must reproduce the row sums of the GLOBAL operator (scalar P1 stiffness as test matrix)
import scipy.sparse as sp


def stiffness(X, el):
    n, dd = el.shape[1], X.shape[1]
    J = np.transpose(X[el[:, 1:]] - X[el[:, :1]], (0, 2, 1))
    vol = np.abs(np.linalg.det(J)) / (2 if dd == 2 else 6)
    Gi = np.linalg.inv(J)
    G = np.concatenate([-Gi.sum(axis=1, keepdims=True), Gi], axis=1)
    K = vol[:, None, None] * np.einsum("eac,ebc->eab", G, G)
    A = sp.csr_matrix((K.ravel(), (np.repeat(el, n, axis=1).ravel(), np.tile(el, (1, n)).ravel())), shape=(len(X), len(X)))
    A.sum_duplicates(); A.sort_indices()
    return A


top = refs
lt, It, Bt = dom.get_level(top), dom._iface[top], dom._biface[top]
A = stiffness(lt["xyz"], lt["elems"])
nvt = len(lt["xyz"])
nn = C.c_int64()
rp, ci = np.empty(nvt + 1, np.int32), np.empty(A.nnz, np.int32)
ug4.call("ab_domain_level_pattern", dom.h, top, C.byref(nn), rp.ctypes.data_as(C.POINTER(C.c_int32)), ci.ctypes.data_as(C.POINTER(C.c_int32)))
assert nn.value == A.nnz and np.array_equal(A.indptr, rp) and np.array_equal(A.indices, ci)
vals = A.data


def vertex_sum(v):
    send = {int(q): v[It["idx"][It["offsets"][n]:It["offsets"][n + 1]]] for n, q in enumerate(It["neigh"])}
    ev = gather(send)
    out = v.copy()
    for n, q in enumerate(It["neigh"]):
        np.add.at(out, It["idx"][It["offsets"][n]:It["offsets"][n + 1]], ev[int(q)][rank])
    return out


rowabs = np.add.reduceat(np.abs(vals), rp[:-1])
loose = vertex_sum(rowabs)
cv = vals[Bt["bpos"]].copy()
ev = gather({int(q): cv[Bt["slot_block"][Bt["offsets"][n]:Bt["offsets"][n + 1]]] for n, q in enumerate(It["neigh"])})
tot = cv.copy()
for n, q in enumerate(It["neigh"]):
    np.add.at(tot, Bt["slot_block"][Bt["offsets"][n]:Bt["offsets"][n + 1]], ev[int(q)][rank])
np.add.at(rowabs, Bt["brow"], np.abs(tot) / Bt["mult"] - np.abs(vals[Bt["bpos"]]))
exact = vertex_sum(rowabs)
# the same steps through the kernel SOURCE of the library compiled for the host (tests/cuda_host_shim/gershgorin_emulation.cpp,
# path in ADMM_B200_GERSH_EMU): d x d blocks a_ij * M with a sign-mixed M, every kernel of lib.cu gmg_setup_kernels' exact-bound
# branch in its order; the result must be the global scalar row sums times the absolute row sums of M
emu_path = os.environ.get("ADMM_B200_GERSH_EMU")
emu_worst = None
if emu_path:
    emu = C.CDLL(emu_path)
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    d = lt["xyz"].shape[1]
    DD = d * d
    M = np.array([[1.0, -0.3, 0.2], [0.25, 1.0, -0.5], [-0.125, 0.4, 1.0]])[:d, :d]
    bvals = np.ascontiguousarray((vals[:, None, None] * M[None]).reshape(-1))
    diagpos = np.ascontiguousarray(np.array([rp[i] + np.searchsorted(ci[rp[i]:rp[i + 1]], i) for i in range(nvt)], np.int32))
    diag, rab = np.zeros(nvt * d), np.zeros(nvt * d)
    emu.emu_diag_rowabs(d, nvt, ip(rp), ip(diagpos), dp(bvals), dp(diag), dp(rab))
    assert np.array_equal(diag.reshape(nvt, d), vals[diagpos][:, None] * np.diag(M)[None])
    nsb = len(Bt["bpos"])
    bpos, brow, mult, slotb = (np.ascontiguousarray(Bt[k], np.int32) for k in ("bpos", "brow", "mult", "slot_block"))
    cvb = np.zeros(max(nsb, 1) * DD)
    emu.emu_pack_blocks(C.c_int64(nsb * DD), DD, ip(bpos), dp(bvals), dp(cvb))
    totalb = int(Bt["offsets"][-1])
    send, recv = np.zeros(max(totalb, 1) * DD), np.zeros(max(totalb, 1) * DD)
    emu.emu_iface_pack(totalb, DD, ip(slotb), dp(cvb), dp(send))
    evb = gather({int(q): send[Bt["offsets"][n] * DD:Bt["offsets"][n + 1] * DD] for n, q in enumerate(It["neigh"])})     # ncclSend / ncclRecv
    for n, q in enumerate(It["neigh"]):
        recv[Bt["offsets"][n] * DD:Bt["offsets"][n + 1] * DD] = evb[int(q)][rank]
    emu.emu_iface_unpack_add(totalb, DD, ip(slotb), dp(recv), dp(cvb))
    emu.emu_rowabs_fix(d, nsb, ip(bpos), ip(brow), ip(mult), dp(bvals), dp(cvb), dp(rab))
    # interface sum of the rows (exchange_sum; its own kernel is covered by tests/cuda_host_shim/xchg_emulation.cpp): the NCCL form
    total_v = int(It["offsets"][-1])
    idx_v = np.ascontiguousarray(It["idx"], np.int32)
    sendv, recvv = np.zeros(max(total_v, 1) * d), np.zeros(max(total_v, 1) * d)
    emu.emu_iface_pack(total_v, d, ip(idx_v), dp(rab), dp(sendv))
    evv = gather({int(q): sendv[It["offsets"][n] * d:It["offsets"][n + 1] * d] for n, q in enumerate(It["neigh"])})
    for n, q in enumerate(It["neigh"]):
        recvv[It["offsets"][n] * d:It["offsets"][n + 1] * d] = evv[int(q)][rank]
    emu.emu_iface_unpack_add(total_v, d, ip(idx_v), dp(recvv), dp(rab))
    want = exact[:, None] * np.abs(M).sum(axis=1)[None]          # `exact` was checked against the global operator below
    emu_worst = float(np.abs(rab.reshape(nvt, d) - want).max() / want.max())
rows = gather((lt["xyz"], exact, loose, emu_worst))
gersh = dict(ok=True)
if rank == 0:
    gdom = ug4.Domain(ug)
    ug._create_from_dict(gdom, dom._global, host_only=True)
    ug4.call("ab_domain_refine", gdom.h, top)
    gl2 = gdom.get_level(top)
    Ag = stiffness(gl2["xyz"], gl2["elems"])
    lut = {tuple(x): v for x, v in zip(gl2["xyz"].tolist(), np.abs(Ag).sum(axis=1).A1)}
    worst, differing = 0.0, 0
    for Xr, ex, lo, _ in rows:
        ref = np.array([lut[tuple(x)] for x in Xr.tolist()])
        worst = max(worst, float(np.abs(ex - ref).max() / ref.max()))
        differing += int((np.abs(lo - ref) > 1e-9 * ref.max()).sum())
    emu_all = [r[3] for r in rows]
    gersh = dict(ok=bool(worst < 1e-12) and all(e is None or e < 1e-12 for e in emu_all), worst=worst, rows_where_loose_differs=differing,
                 shared_blocks=int(len(Bt["bpos"])), kernel_source_emulation_worst=emu_all)
allres = gather(dict(decomposed=True, gather_level=lg, levels=res, blocks_ok=ok, gershgorin=gersh))
if rank == 0:
    print(json.dumps(allres))
dist.barrier()
dist.destroy_process_group()
