"""ORACLE (test infrastructure, not product code) -- P1/P0 finite-element restatement in NumPy/SciPy.

PARITY UNPINNED: none of the element formulas exist in /root/reference (the plugins FluidOptim /
PLaplacian / ADMMOptim named at 3d_admm.lua:1-3 are un-vendored, SURVEY.md section 0 + App. B).
Every function states the script evidence it rests on.  The model:

    L(u,q,lam,Lam) = J'(u) + (lam, grad u - q) + tau/2 |grad u - q|^2 + sum_i Lam_i g_i(u),  |q_T|_F <= sigma
    g_vol(u)  = int det(I + grad u) dx - V_ref            (VolumeDefect,     3d_admm.lua:780,1167)
    g_k(u)    = int (x_k + u_k) det(I + grad u) dx         (BarycenterDefect, 3d_admm.lua:1168)

DoF layout: P1 index = vertex*dim + comp; P0 index = element*dim*dim + (row*dim + col)
(l1..l9 = lambda00,01,02,10,..,22 -- 3d_admm.lua:343-351).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .mesh_np import Mesh, p1_pattern

# ------------------------------------------------------------------------------------------
# element geometry
# ------------------------------------------------------------------------------------------


_GEOM_CACHE = {}


def geometry(mesh: Mesh):
    """P1 gradients G (ne,d+1,d), element measure vol (ne,) (|det J|/d!  -- refined.ugx has 134
    clockwise triangles), centroid xbar (ne,d).  Cached per coordinate array content."""
    key = (id(mesh), mesh.xyz.ctypes.data, hash(mesh.xyz[:: max(1, mesh.nv // 64)].tobytes()), float(mesh.xyz.sum()))
    hit = _GEOM_CACHE.get(id(mesh))
    if hit is not None and hit[0] == key:
        return hit[1]
    out = _geometry(mesh)
    _GEOM_CACHE[id(mesh)] = (key, out)
    return out


def _geometry(mesh: Mesh):
    d = mesh.dim
    X = mesh.xyz[mesh.elems]                         # (ne,d+1,d)
    J = np.transpose(X[:, 1:, :] - X[:, :1, :], (0, 2, 1))   # columns = edge vectors
    Jinv = np.linalg.inv(J)
    G = np.empty((mesh.ne, d + 1, d))
    G[:, 1:, :] = Jinv                               # rows of J^-1
    G[:, 0, :] = -Jinv.sum(axis=1)
    vol = np.abs(np.linalg.det(J)) / math.factorial(d)
    return G, vol, X.mean(axis=1)


def grad_u(mesh: Mesh, G, u):
    """(grad u)_{ij} = sum_a u_{a,i} G_{a,j}, constant per element. u is (nv*d,) or None (=0)."""
    d = mesh.dim
    if u is None:
        return np.zeros((mesh.ne, d, d))
    U = u.reshape(-1, d)[mesh.elems]                 # (ne,d+1,d)
    return np.einsum("eai,eaj->eij", U, G)


def cofactor(F):
    """cof(F) = det(F) F^-T, written polynomially so that singular F is fine."""
    d = F.shape[-1]
    C = np.empty_like(F)
    if d == 2:
        C[:, 0, 0], C[:, 0, 1] = F[:, 1, 1], -F[:, 1, 0]
        C[:, 1, 0], C[:, 1, 1] = -F[:, 0, 1], F[:, 0, 0]
    else:
        for i in range(3):
            C[:, i, :] = np.cross(F[:, (i + 1) % 3, :], F[:, (i + 2) % 3, :])
    return C


def det(F):
    d = F.shape[-1]
    if d == 2:
        return F[:, 0, 0] * F[:, 1, 1] - F[:, 0, 1] * F[:, 1, 0]
    return np.einsum("ei,ei->e", F[:, 0, :], np.cross(F[:, 1, :], F[:, 2, :]))


def _eps_det(mesh, G, F):
    """T[e,a,i,b,j] = sum_k eps_{ijk} det[G_a, G_b, F_k]   (3D)   /   eps_{ij} det[G_a, G_b]  (2D)
    = d cof(F)[e_j (x) G_b] : (e_i (x) G_a), the second derivative kernel of det(I+grad u)."""
    d = mesh.dim
    ne = mesh.ne
    T = np.zeros((ne, d + 1, d, d + 1, d))
    if d == 2:
        D = G[:, :, None, 0] * G[:, None, :, 1] - G[:, :, None, 1] * G[:, None, :, 0]   # (ne,a,b)
        T[:, :, 0, :, 1] = D
        T[:, :, 1, :, 0] = -D
    else:
        for i in range(3):
            j, k = (i + 1) % 3, (i + 2) % 3
            # det[G_a, G_b, F_k] = G_a . (G_b x F_k)
            GbxFk = np.cross(G[:, :, :], F[:, None, k, :])            # (ne,b,3)
            T[:, :, i, :, j] = np.einsum("eax,ebx->eab", G, GbxFk)
            GbxFj = np.cross(G[:, :, :], F[:, None, j, :])
            T[:, :, i, :, k] = -np.einsum("eax,ebx->eab", G, GbxFj)
    return T


# ------------------------------------------------------------------------------------------
# P1 assembly
# ------------------------------------------------------------------------------------------


def dirichlet_dofs(mesh: Mesh, subsets=("inlet", "wall", "outlet")):
    """u = 0 on inlet, wall, outlet for every component; obstacle_surface free (3d_admm.lua:445-457)."""
    m = mesh.vertex_mask(list(subsets))
    return np.repeat(m, mesh.dim)


def hessian_matrix(mesh: Mesh, u, c=1.0, lam_vol=0.0, lam_bary=None, dmask=None):
    """DeformationEquation jacobian (3d_admm.lua:393-405, assembled at :972,:1008,... ):
        K[(a,i),(b,j)] = vol*( c*delta_ij G_a.G_b
                               + (Lam_vol + sum_k Lam_k (xbar_k+ubar_k)) * T[a,i,b,j]
                               + sum_k Lam_k (delta_ik (C G_b)_j + delta_jk (C G_a)_i)/(d+1) )
    followed by symmetric Dirichlet elimination (rows AND columns -> identity)."""
    d = mesh.dim
    ne = mesh.ne
    G, vol, xbar = geometry(mesh)
    lam_bary = np.zeros(d) if lam_bary is None else np.asarray(lam_bary, float)
    K = np.zeros((ne, d + 1, d, d + 1, d))
    GG = np.einsum("eax,ebx->eab", G, G)
    for i in range(d):
        K[:, :, i, :, i] += c * GG
    if lam_vol != 0.0 or np.any(lam_bary != 0.0):
        F = np.eye(d)[None] + grad_u(mesh, G, u)
        C = cofactor(F)
        ubar = np.zeros((ne, d)) if u is None else u.reshape(-1, d)[mesh.elems].mean(axis=1)
        w = lam_vol + (xbar + ubar) @ lam_bary
        K += w[:, None, None, None, None] * _eps_det(mesh, G, F)
        CG = np.einsum("eij,eaj->eai", C, G)          # (C G_a)_i
        for k in range(d):
            if lam_bary[k] != 0.0:
                K[:, :, k, :, :] += lam_bary[k] / (d + 1) * CG[:, None, :, :]
                K[:, :, :, :, k] += lam_bary[k] / (d + 1) * CG[:, :, :, None]
    K *= vol[:, None, None, None, None]
    dof = (mesh.elems[:, :, None].astype(np.int64) * d + np.arange(d)[None, None, :])   # (ne,a,i)
    rows = np.broadcast_to(dof[:, :, :, None, None], K.shape).ravel()
    cols = np.broadcast_to(dof[:, None, None, :, :], K.shape).ravel()
    n = mesh.nv * d
    A = sp.csr_matrix((K.ravel(), (rows, cols)), shape=(n, n))
    A.sum_duplicates()
    if dmask is not None:
        A = apply_dirichlet_sym(A, dmask)
    return A


def apply_dirichlet_sym(A, dmask):
    keep = sp.diags((~dmask).astype(float))
    A = keep @ A @ keep + sp.diags(dmask.astype(float))
    A = A.tocsr()
    A.sort_indices()
    return A


def load_vector(mesh: Mesh, u, S=None, w=None, sign=1.0):
    """Generic P1 element load vector (one kernel serves all RHS-type ElemDiscs):
        f[(a,i)] = sign * vol * ( ((S + wc*C) G_a)_i + w_i det(F)/(d+1) ),   wc = w_vol + sum_k w_k (xbar_k+ubar_k)
    with w = (w_vol, w_1..w_d), C = cof(F), F = I + grad u.
      DeformationEquationRHS            S = lam + tau(grad u - q), w = Lam        sign +1  (3d_admm.lua:407-442,973)
      DeformationEquationLargeProblemRHS S as above,               w = Lam + mult  sign +1  (3d_admm.lua:472-507,1081-1091)
      VolumeConstraintSecondDerivative   S = 0,                     w = (1,0,0,0)   sign -1  (3d_admm.lua:559,954-959)
      SecondDerivativeBarycenter(k)      S = 0,                     w = e_k         sign -1  (3d_admm.lua:577-617)
    """
    d = mesh.dim
    G, vol, xbar = geometry(mesh)
    gu = grad_u(mesh, G, u)
    M = np.zeros((mesh.ne, d, d)) if S is None else S.copy()
    f = np.zeros((mesh.ne, d + 1, d))
    if w is not None and np.any(np.asarray(w) != 0.0):
        w = np.asarray(w, float)
        F = np.eye(d)[None] + gu
        C = cofactor(F)
        ubar = np.zeros((mesh.ne, d)) if u is None else u.reshape(-1, d)[mesh.elems].mean(axis=1)
        wc = w[0] + (xbar + ubar) @ w[1:]
        M = M + wc[:, None, None] * C
        f += (det(F)[:, None, None] / (d + 1)) * w[None, None, 1:]
    f += np.einsum("eij,eaj->eai", M, G)
    f *= sign * vol[:, None, None]
    out = np.zeros(mesh.nv * d)
    dof = (mesh.elems[:, :, None].astype(np.int64) * d + np.arange(d)[None, None, :])
    np.add.at(out, dof.ravel(), f.ravel())
    return out


# ------------------------------------------------------------------------------------------
# P0 (tensor) operations
# ------------------------------------------------------------------------------------------


def p0_grad(mesh, u):
    G, _, _ = geometry(mesh)
    return grad_u(mesh, G, u).reshape(-1)


def mass_model(mesh, u, lam):
    """MassModel (3d_admm.lua:650-674, assembled :899-900): diagonal P0 mass matrix diag = |T| and
    defect = -|T| (grad u + lam); the script solves DiagQ q = defect and negates (:903-905)."""
    d2 = mesh.dim ** 2
    _, vol, _ = geometry(mesh)
    diag = np.repeat(vol, d2)
    rhs = -diag * (p0_grad(mesh, u) + lam)
    return diag, rhs


def project_frobenius(q, sigma, d):
    """Testing(q_projected,q_piecewise,cmps,sigma) (3d_admm.lua:910): q * min(1, sigma/|q|_F) per element."""
    Q = q.reshape(-1, d * d)
    nrm = np.sqrt((Q * Q).sum(axis=1))
    s = np.where(nrm > sigma, sigma / np.where(nrm > 0, nrm, 1.0), 1.0)
    return (Q * s[:, None]).reshape(-1)


def _svd2(Q):
    """Closed-form SVD pieces of 2x2 matrices: returns s1>=s2>=0 and rotation angles."""
    a, b, c, dd = Q[:, 0], Q[:, 1], Q[:, 2], Q[:, 3]
    E, Fh, Gh, H = (a + dd) / 2, (a - dd) / 2, (c + b) / 2, (c - b) / 2
    q_, r_ = np.hypot(E, H), np.hypot(Fh, Gh)
    s1, s2 = q_ + r_, q_ - r_           # s2 may be negative (signed), |s2| is the singular value
    a1, a2 = np.arctan2(Gh, Fh), np.arctan2(H, E)
    theta, phi = (a2 - a1) / 2, (a2 + a1) / 2
    return s1, s2, theta, phi


def project_spectral(q, sigma):
    """ProjectWithSpectralNorm (2d_admm.lua:902): clip the singular values of the 2x2 at sigma."""
    Q = q.reshape(-1, 4)
    s1, s2, theta, phi = _svd2(Q)
    t1 = np.minimum(s1, sigma)
    t2 = np.clip(s2, -sigma, sigma)
    cp, sp_, ct, st = np.cos(phi), np.sin(phi), np.cos(theta), np.sin(theta)
    # Q = R(phi) diag(s1,s2) R(theta)   with R(x) = [[cos,-sin],[sin,cos]]
    out = np.empty_like(Q)
    out[:, 0] = cp * t1 * ct - sp_ * t2 * st
    out[:, 1] = -cp * t1 * st - sp_ * t2 * ct
    out[:, 2] = sp_ * t1 * ct + cp * t2 * st
    out[:, 3] = -sp_ * t1 * st + cp * t2 * ct
    return out.reshape(-1)


def max_frobenius_norm(mesh, u):
    """MaximumFrobeniusNorm(u_old,cmps,"outer",4) (3d_admm.lua:916)."""
    g = p0_grad(mesh, u).reshape(-1, mesh.dim ** 2)
    return float(np.sqrt((g * g).sum(axis=1)).max())


def max_spectral_norm(mesh, u):
    """MaxSpectralNorm (2d_admm.lua:901): max over elements of the largest singular value."""
    g = p0_grad(mesh, u).reshape(-1, 4)
    s1, _, _, _ = _svd2(g)
    return float(s1.max())


def lambda_update_defect(mesh, u, qproj, tau=1.0):
    """LambdaUpdate (3d_admm.lua:677-694, :1221): defect = -tau (grad u - q_proj) pointwise per P0 dof
    (the script negates it and adds it to lambda: "lambda += tau(Grad_u-q_proj)", :1220-1223)."""
    return -tau * (p0_grad(mesh, u) - qproj)


# ------------------------------------------------------------------------------------------
# integrals / norms
# ------------------------------------------------------------------------------------------


def l2norm_p1(mesh, v, comp):
    """L2Norm(gf,"u1",4,"outer") (3d_admm.lua:1137-1146): exact P1 mass-matrix norm of one component."""
    d = mesh.dim
    _, vol, _ = geometry(mesh)
    f = v.reshape(-1, d)[:, comp][mesh.elems]
    val = (vol / ((d + 1) * (d + 2)) * ((f * f).sum(axis=1) + f.sum(axis=1) ** 2)).sum()
    return math.sqrt(val)


def l2norm_p0(mesh, v, comp):
    """L2Norm(temp1_piecewise,"l1",4,"outer") (3d_admm.lua:1243-1251)."""
    _, vol, _ = geometry(mesh)
    f = v.reshape(-1, mesh.dim ** 2)[:, comp]
    return math.sqrt((vol * f * f).sum())


def volume_defect(mesh, u, vref):
    """VolumeDefect(u,Vref,"outer",cmps,4,false,1,false) (3d_admm.lua:780,1167)."""
    G, vol, _ = geometry(mesh)
    F = np.eye(mesh.dim)[None] + grad_u(mesh, G, u)
    return float((vol * det(F)).sum() - vref)


def barycenter_defect(mesh, u):
    """BarycenterDefect(u,cmps,"outer",4) (3d_admm.lua:1168): int (x_k+u_k) det(I+grad u) dx, k=1..d."""
    d = mesh.dim
    G, vol, xbar = geometry(mesh)
    F = np.eye(d)[None] + grad_u(mesh, G, u)
    ubar = u.reshape(-1, d)[mesh.elems].mean(axis=1)
    return ((vol * det(F))[:, None] * (xbar + ubar)).sum(axis=0)


# ------------------------------------------------------------------------------------------
# multigrid + Krylov (obstacle_optim_3d_util.lua:9-43)
# ------------------------------------------------------------------------------------------
import ctypes as _C
import os as _os

_CLIB = None


def c_kernels():
    """oracle/liboracle_c.so (oracle_kernels.c): forward Gauss-Seidel sweep + CSR SpMV; None when not built."""
    global _CLIB
    if _CLIB is None:
        path = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "liboracle_c.so")
        if not _os.path.exists(path):
            _CLIB = False
        else:
            lib = _C.CDLL(path)
            ip, dp = _C.POINTER(_C.c_int), _C.POINTER(_C.c_double)
            lib.oracle_gs_forward.argtypes = [_C.c_int, ip, ip, dp, dp, dp, _C.c_int, dp]
            lib.oracle_gs_forward.restype = None
            lib.oracle_spmv.argtypes = [_C.c_int, ip, ip, dp, dp, dp, _C.c_int]
            lib.oracle_spmv.restype = None
            lib.oracle_max_threads.restype = _C.c_int
            if hasattr(lib, "oracle_set_elem_threads"):
                lib.oracle_set_elem_threads.argtypes = [_C.c_int]
                lib.oracle_set_elem_threads.restype = None
            if hasattr(lib, "oracle_hessian_scatter"):
                llp = _C.POINTER(_C.c_longlong)
                lib.oracle_hessian_scatter.argtypes = [_C.c_int, _C.c_long, ip, dp, dp, _C.c_double, _C.c_double, dp, llp, dp]
                lib.oracle_hessian_scatter.restype = None
                lib.oracle_load_scatter.argtypes = [_C.c_int, _C.c_long, ip, dp, dp, dp, dp, _C.c_double, _C.c_int, dp, _C.c_int, _C.c_double, dp]
                lib.oracle_load_scatter.restype = None
            if hasattr(lib, "oracle_gmg_create"):
                vp, ucp = _C.c_void_p, _C.POINTER(_C.c_ubyte)
                lib.oracle_gmg_create.argtypes = [_C.c_int, _C.c_int, _C.c_int, _C.c_int, _C.c_int, _C.c_double, _C.c_double]
                lib.oracle_gmg_create.restype = vp
                lib.oracle_gmg_set_level.argtypes = [vp, _C.c_int, _C.c_int, _C.c_int, ip, ip, dp, ip, ip, dp, ucp]
                lib.oracle_gmg_set_level.restype = None
                lib.oracle_gmg_setup.argtypes = [vp, ip, ip, dp]
                lib.oracle_gmg_setup.restype = _C.c_int
                lib.oracle_gmg_vcycle.argtypes = [vp, dp, dp]
                lib.oracle_gmg_vcycle.restype = None
                lib.oracle_bicgstab_gmg.argtypes = [vp, dp, dp, _C.c_double, _C.c_double, _C.c_int, _C.POINTER(_C.c_int), dp]
                lib.oracle_bicgstab_gmg.restype = _C.c_int
                lib.oracle_gmg_destroy.argtypes = [vp]
                lib.oracle_gmg_destroy.restype = None
            _CLIB = lib
    return _CLIB or None


# ------------------------------------------------------------------------------------------
# compiled assembly (bench.py's CPU legs: ug4_np.Backend(fast_assembly=True)); the NumPy functions above stay the checker
# ------------------------------------------------------------------------------------------
_PLAN_CACHE = {}


def _scatter_plan(mesh: Mesh):
    """CSR structure of the P1 operator on `mesh` and, for every element-matrix entry (e,a,i,b,j) in K.ravel() order, its slot
    in the CSR data array.  Built once per mesh (the pattern never changes)."""
    hit = _PLAN_CACHE.get(id(mesh))
    if hit is not None and hit["nv"] == mesh.nv and hit["ne"] == mesh.ne:
        return hit
    d, n = mesh.dim, mesh.nv * mesh.dim
    dof = (mesh.elems[:, :, None].astype(np.int64) * d + np.arange(d)[None, None, :])   # (ne,a,i)
    shape = (mesh.ne, d + 1, d, d + 1, d)
    key = (np.broadcast_to(dof[:, :, :, None, None], shape) * n + np.broadcast_to(dof[:, None, None, :, :], shape)).ravel()
    uniq, slot = np.unique(key, return_inverse=True)
    rows, cols = uniq // n, uniq % n
    indptr = np.searchsorted(rows, np.arange(n + 1)).astype(np.int32)
    plan = dict(nv=mesh.nv, ne=mesh.ne, n=n, slot=np.ascontiguousarray(slot.ravel(), np.int64), rows=rows, indptr=indptr,
                indices=cols.astype(np.int32), diag=np.flatnonzero(rows == cols), elems=np.ascontiguousarray(mesh.elems, np.int32), dmask={})
    _PLAN_CACHE[id(mesh)] = plan
    return plan


def hessian_matrix_fast(mesh: Mesh, u, c=1.0, lam_vol=0.0, lam_bary=None, dmask=None):
    """Same operator as hessian_matrix() (entries agree to rounding, tests/test_oracle.py), element loop and scatter in C.
    Dirichlet rows / columns are zeroed in place (explicit zeros stay in the pattern), diagonal 1."""
    lib = c_kernels()
    if lib is None or not hasattr(lib, "oracle_hessian_scatter"):
        return hessian_matrix(mesh, u, c, lam_vol, lam_bary, dmask)
    d = mesh.dim
    plan = _scatter_plan(mesh)
    lb = np.zeros(3)
    if lam_bary is not None:
        lb[:d] = np.asarray(lam_bary, float)[:d]
    data = np.zeros(len(plan["indices"]))
    xyz = np.ascontiguousarray(mesh.xyz, np.float64)
    uu = None if u is None else np.ascontiguousarray(u, np.float64)
    dp, ip = _C.POINTER(_C.c_double), _C.POINTER(_C.c_int)
    lib.oracle_hessian_scatter(d, mesh.ne, plan["elems"].ctypes.data_as(ip), xyz.ctypes.data_as(dp),
                               uu.ctypes.data_as(dp) if uu is not None else None, float(c), float(lam_vol), lb.ctypes.data_as(dp),
                               plan["slot"].ctypes.data_as(_C.POINTER(_C.c_longlong)), data.ctypes.data_as(dp))
    if dmask is not None:
        k = dmask.tobytes()
        ent = plan["dmask"].get(k)
        if ent is None:
            ent = (np.flatnonzero(dmask[plan["rows"]] | dmask[plan["indices"]]), plan["diag"][dmask])
            plan["dmask"] = {k: ent}
        data[ent[0]] = 0.0
        data[ent[1]] = 1.0
    return sp.csr_matrix((data, plan["indices"], plan["indptr"]), shape=(plan["n"], plan["n"]))


def load_vector_fast(mesh: Mesh, u, lam=None, q=None, tau=0.0, w=None, sign=1.0):
    """load_vector() with S = lam + tau (grad u - q) formed inside the C element loop (S = 0 when lam is None)."""
    lib = c_kernels()
    d = mesh.dim
    if lib is None or not hasattr(lib, "oracle_load_scatter"):
        S = None
        if lam is not None:
            G, _, _ = geometry(mesh)
            S = lam.reshape(-1, d, d) + tau * (grad_u(mesh, G, u) - q.reshape(-1, d, d))
        return load_vector(mesh, u, S, w, sign)
    plan = _scatter_plan(mesh)
    out = np.zeros(mesh.nv * d)
    xyz = np.ascontiguousarray(mesh.xyz, np.float64)
    uu = None if u is None else np.ascontiguousarray(u, np.float64)
    ww = np.zeros(4)
    has_w = w is not None and bool(np.any(np.asarray(w) != 0.0))
    if has_w:
        ww[:d + 1] = np.asarray(w, float)[:d + 1]
    use_S = lam is not None
    la = np.ascontiguousarray(lam, np.float64).ravel() if use_S else None
    qq = np.ascontiguousarray(q, np.float64).ravel() if use_S else None
    dp, ip = _C.POINTER(_C.c_double), _C.POINTER(_C.c_int)
    lib.oracle_load_scatter(d, mesh.ne, plan["elems"].ctypes.data_as(ip), xyz.ctypes.data_as(dp), uu.ctypes.data_as(dp) if uu is not None else None,
                            la.ctypes.data_as(dp) if use_S else None, qq.ctypes.data_as(dp) if use_S else None, float(tau), int(use_S),
                            ww.ctypes.data_as(dp), int(has_w), float(sign), out.ctypes.data_as(dp))
    return out


class CsrC:
    """CSR matrix with C SpMV / GS (falls back to SciPy when the C library is not built)."""

    def __init__(self, A, threads=1):
        A = A.tocsr()
        A.sort_indices()
        self.A, self.n, self.threads = A, A.shape[0], threads
        self.ip = np.ascontiguousarray(A.indptr, np.int32)
        self.idx = np.ascontiguousarray(A.indices, np.int32)
        self.a = np.ascontiguousarray(A.data, np.float64)
        self.lib = c_kernels()
        self._xold = np.empty(self.n)
        self._LD = None

    def _p(self, arr, t):
        return arr.ctypes.data_as(_C.POINTER(t))

    def matvec(self, x):
        if self.lib is None:
            return self.A @ x
        x = np.ascontiguousarray(x, np.float64)
        y = np.empty(self.n)
        self.lib.oracle_spmv(self.n, self._p(self.ip, _C.c_int), self._p(self.idx, _C.c_int), self._p(self.a, _C.c_double),
                             self._p(x, _C.c_double), self._p(y, _C.c_double), self.threads)
        return y

    def gs_sweep(self, x, b, blocks=1):
        """one forward sweep, in place on a copy; blocks > 1 = block-Jacobi across `blocks` row blocks"""
        if self.lib is None:
            if self._LD is None:
                self._LD = sp.tril(self.A, 0, format="csr")
            return x + spla.spsolve_triangular(self._LD, b - self.A @ x, lower=True)
        x = np.array(x, dtype=np.float64, copy=True)
        b = np.ascontiguousarray(b, np.float64)
        self.lib.oracle_gs_forward(self.n, self._p(self.ip, _C.c_int), self._p(self.idx, _C.c_int), self._p(self.a, _C.c_double),
                                   self._p(b, _C.c_double), self._p(x, _C.c_double), blocks, self._p(self._xold, _C.c_double))
        return x



def prolongation(fine: Mesh, d: int):
    """StdTransfer for nested P1: copies weight 1, edge midpoints 1/2,1/2; same for all comps."""
    nvc, nvf = fine.nv_coarse, fine.nv
    k = nvf - nvc
    rows = np.concatenate([np.arange(nvc), nvc + np.arange(k), nvc + np.arange(k)])
    cols = np.concatenate([np.arange(nvc), fine.parent_a, fine.parent_b])
    vals = np.concatenate([np.ones(nvc), 0.5 * np.ones(2 * k)])
    Ps = sp.csr_matrix((vals, (rows, cols)), shape=(nvf, nvc))
    return sp.kron(Ps, sp.identity(d), format="csr")


def gershgorin_lmax(A):
    """Upper bound of lambda_max(D^-1 A): max_i sum_j |a_ij| / a_ii (used by the Chebyshev smoother)."""
    rs = np.asarray(abs(A).sum(axis=1)).ravel()
    return float((rs / A.diagonal()).max())


class GMG:
    """V(nu1,nu2) geometric multigrid, Galerkin RAP coarse operators, dense/LU base solve on level 0.
    smoother: 'gs'   forward lexicographic Gauss-Seidel (what the reference asks for, u3:16 -- [UPSTREAM-UNVERIFIED] C5)
              'cheb' Chebyshev(point-Jacobi) of degree nu on [lmax/4, lmax], lmax = 1.1*Gershgorin bound? no: = Gershgorin
              'jac'  damped point Jacobi, omega = 0.66
    The product (CUDA) path implements 'cheb' and 'jac'; 'gs' is kept for side-by-side iteration counts."""

    def __init__(self, levels, A_top, dmasks, smoother="cheb", nu1=3, nu2=3, cheb_ratio=4.0, omega=0.66, threads=1):
        d = levels[0].dim
        self.smoother, self.nu1, self.nu2, self.omega, self.threads = smoother, nu1, nu2, omega, threads
        self.A = [None] * len(levels)
        self.P = [None] * len(levels)
        self.keep = [(~m).astype(float) for m in dmasks]
        self.A[-1] = A_top.tocsr()
        for l in range(len(levels) - 1, 0, -1):
            P = prolongation(levels[l], d)
            self.P[l] = P
            Ac = (P.T @ self.A[l] @ P).tocsr()
            self.A[l - 1] = apply_dirichlet_sym(Ac, dmasks[l - 1])
        self.lu = spla.splu(self.A[0].tocsc())
        self.dinv = [1.0 / A.diagonal() for A in self.A]
        self.lmax = [gershgorin_lmax(A) for A in self.A]
        self.cheb_ratio = cheb_ratio
        self.C = [CsrC(A, threads) for A in self.A]

    # -- smoothers -------------------------------------------------------------------------
    def _smooth(self, l, x, b, nu, zero_guess):
        A = self.C[l]
        if self.smoother == "gs":
            for _ in range(nu):
                x = A.gs_sweep(x, b, self.threads)
            return x
        if self.smoother == "jac":
            for it in range(nu):
                r = b if (zero_guess and it == 0) else b - A.matvec(x)
                x = x + self.omega * self.dinv[l] * r
            return x
        # Chebyshev (Saad, Alg. 12.1) preconditioned by point Jacobi
        lmax = self.lmax[l]
        lmin = lmax / self.cheb_ratio
        theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
        sigma1 = theta / delta
        rho = 1.0 / sigma1
        r = b if zero_guess else b - A.matvec(x)
        dvec = self.dinv[l] * r / theta
        x = x + dvec
        for _ in range(nu - 1):
            rho_new = 1.0 / (2.0 * sigma1 - rho)
            r = b - A.matvec(x)
            dvec = rho_new * rho * dvec + (2.0 * rho_new / delta) * (self.dinv[l] * r)
            x = x + dvec
            rho = rho_new
        return x

    def vcycle(self, l, b):
        if l == 0:
            return self.lu.solve(b)
        x = self._smooth(l, np.zeros_like(b), b, self.nu1, True)
        r = b - self.C[l].matvec(x)
        bc = self.keep[l - 1] * (self.P[l].T @ r)
        x = x + self.P[l] @ self.vcycle(l - 1, bc)
        return self._smooth(l, x, b, self.nu2, False)

    def apply(self, b):
        return self.vcycle(len(self.A) - 1, b)


class GMGC:
    """The hierarchy of `GMG` (same Galerkin operators, smoothers, transfers, direct base solve) built and applied by
    oracle/solver_c.c on `threads` OpenMP threads, together with the BiCGStab loop: bench.py's multi-threaded CPU baseline.
    The static part (transfers, Dirichlet masks, work vectors) is set once per mesh hierarchy; `setup(A)` runs what the reference
    does at every solver:init (RAP chain, smoother data, base-solver factorisation).  Gauss-Seidel runs inside each thread's row
    block with Jacobi coupling between the blocks (threads = 1: the sequential lexicographic sweep of `GMG`)."""
    SMOOTHERS = {"gs": 0, "cheb": 1, "jac": 2}

    def __init__(self, levels, dmasks, smoother="gs", nu1=3, nu2=3, cheb_ratio=6.0, omega=0.66, threads=1):
        self.lib = c_kernels()
        if self.lib is None or not hasattr(self.lib, "oracle_gmg_create"):
            raise RuntimeError("oracle/liboracle_c.so is not built (make -C oracle)")
        d = levels[0].dim
        self.n = levels[-1].nv * d
        self.h = self.lib.oracle_gmg_create(len(levels), int(threads), self.SMOOTHERS[smoother], nu1, nu2, cheb_ratio, omega)
        self._keep = []
        ip_t, dp_t, uc_t = _C.POINTER(_C.c_int), _C.POINTER(_C.c_double), _C.POINTER(_C.c_ubyte)
        for l, lev in enumerate(levels):
            mask = np.ascontiguousarray(dmasks[l], np.uint8)
            if l == 0:
                self._keep.append(mask)
                self.lib.oracle_gmg_set_level(self.h, 0, lev.nv * d, 0, None, None, None, None, None, None, mask.ctypes.data_as(uc_t))
                continue
            P = prolongation(lev, d).tocsr(); P.sort_indices()
            R = P.T.tocsr(); R.sort_indices()
            arrs = [np.ascontiguousarray(P.indptr, np.int32), np.ascontiguousarray(P.indices, np.int32), np.ascontiguousarray(P.data, np.float64),
                    np.ascontiguousarray(R.indptr, np.int32), np.ascontiguousarray(R.indices, np.int32), np.ascontiguousarray(R.data, np.float64), mask]
            self._keep.append(arrs)
            self.lib.oracle_gmg_set_level(self.h, l, lev.nv * d, levels[l - 1].nv * d, arrs[0].ctypes.data_as(ip_t), arrs[1].ctypes.data_as(ip_t),
                                          arrs[2].ctypes.data_as(dp_t), arrs[3].ctypes.data_as(ip_t), arrs[4].ctypes.data_as(ip_t),
                                          arrs[5].ctypes.data_as(dp_t), mask.ctypes.data_as(uc_t))
        self._A = None

    def setup(self, A_top):
        A = A_top.tocsr(); A.sort_indices()
        self._A = (np.ascontiguousarray(A.indptr, np.int32), np.ascontiguousarray(A.indices, np.int32), np.ascontiguousarray(A.data, np.float64))
        ip_t, dp_t = _C.POINTER(_C.c_int), _C.POINTER(_C.c_double)
        rc = self.lib.oracle_gmg_setup(self.h, self._A[0].ctypes.data_as(ip_t), self._A[1].ctypes.data_as(ip_t), self._A[2].ctypes.data_as(dp_t))
        if rc != 0:
            raise RuntimeError("coarse-level matrix is singular")

    def apply(self, b):
        b = np.ascontiguousarray(b, np.float64)
        x = np.empty(self.n)
        dp_t = _C.POINTER(_C.c_double)
        self.lib.oracle_gmg_vcycle(self.h, b.ctypes.data_as(dp_t), x.ctypes.data_as(dp_t))
        return x

    def solve(self, b, x0, abs_tol=1e-10, max_it=3000, red_tol=0.0):
        b = np.ascontiguousarray(b, np.float64)
        x = np.array(x0, dtype=np.float64, copy=True)
        r = np.empty(self.n)
        ok = _C.c_int(0)
        dp_t = _C.POINTER(_C.c_double)
        its = self.lib.oracle_bicgstab_gmg(self.h, b.ctypes.data_as(dp_t), x.ctypes.data_as(dp_t), abs_tol, red_tol, max_it, _C.byref(ok), r.ctypes.data_as(dp_t))
        return x, bool(ok.value), int(its), r

    def __del__(self):
        try:
            if self.h:
                self.lib.oracle_gmg_destroy(self.h)
                self.h = None
        except Exception:
            pass


def bicgstab(A, b, x0, precond, abs_tol=1e-10, max_it=3000, red_tol=0.0, check_half=False):
    """Right-preconditioned BiCGStab with the reference's ConvCheck semantics (u3:32-38):
    stop when |r|_2 < abs_tol (or |r|/|r0| < red_tol), fail after max_it. Returns (x, ok, its, r)."""
    x = x0.copy()
    r = b - A @ x
    nr0 = nr = np.linalg.norm(r)
    if nr < abs_tol:
        return x, True, 0, r
    rh = r.copy()
    rho_old = alpha = omega = 1.0
    v = np.zeros_like(b)
    p = np.zeros_like(b)
    for it in range(1, max_it + 1):
        rho = rh @ r
        if rho == 0.0 or not np.isfinite(rho):
            return x, False, it, r
        beta = (rho / rho_old) * (alpha / omega)
        p = r + beta * (p - omega * v)
        ph = precond(p)
        v = A @ ph
        alpha = rho / (rh @ v)
        s = r - alpha * v
        if check_half and np.linalg.norm(s) < abs_tol:
            return x + alpha * ph, True, it, s
        sh = precond(s)
        t = A @ sh
        tt = t @ t
        omega = (t @ s) / tt if tt > 0 else 0.0
        x = x + alpha * ph + omega * sh
        r = s - omega * t
        rho_old = rho
        nr = np.linalg.norm(r)
        if nr < abs_tol or nr < red_tol * nr0:
            return x, True, it, r
        if omega == 0.0 or not np.isfinite(nr):
            return x, False, it, r
    return x, False, max_it, r


def cg_jacobi(diag, b, x0, damp=0.66, abs_tol=1e-9, max_it=2000):
    """CG + Jacobi(0.66) + ConvCheck(2000,1e-9,0,true) on the diagonal P0 mass matrix (3d_admm.lua:701-703)."""
    x = x0.copy()
    r = b - diag * x
    if np.linalg.norm(r) < abs_tol:
        return x, True, 0
    z = damp * r / diag
    p = z.copy()
    rz = r @ z
    for it in range(1, max_it + 1):
        q = diag * p
        a = rz / (p @ q)
        x += a * p
        r -= a * q
        if np.linalg.norm(r) < abs_tol:
            return x, True, it
        z = damp * r / diag
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, False, max_it


# ------------------------------------------------------------------------------------------
# lua-matrix restatement (host Schur algebra, 3d_admm.lua:1063-1078)
# ------------------------------------------------------------------------------------------


def lua_matrix_invert(S):
    """matrix.invert (lua-matrix/matrix.lua:513-534) = Gauss-Jordan on [S | I] via dogauss (:450-507)
    whose pivot rule picks the SMALLEST non-zero |entry| of the column (pivotOk, :422-442)."""
    n = len(S)
    m = [list(map(float, S[i])) + [1.0 if i == j else 0.0 for j in range(n)] for i in range(n)]
    cols = 2 * n
    for j in range(n):
        imin, nmin = None, math.inf
        for i in range(j, n):
            a = abs(m[i][j])
            if 0 < a < nmin:
                imin, nmin = i, a
        if imin is None:
            return None
        if imin != j:
            m[j], m[imin] = m[imin], m[j]
        for i in range(j + 1, n):
            if m[i][j] != 0:
                fac = m[i][j] / m[j][j]
                m[i][j] = 0.0
                for jj in range(j + 1, cols):
                    m[i][jj] = m[i][jj] - fac * m[j][jj]
    for j in range(n - 1, -1, -1):
        div = m[j][j]
        for jj in range(j + 1, cols):
            m[j][jj] = m[j][jj] / div
        for i in range(j - 1, -1, -1):
            if m[i][j] != 0:
                fac = m[i][j]
                for jj in range(j + 1, cols):
                    m[i][jj] = m[i][jj] - fac * m[j][jj]
                m[i][j] = 0.0
        m[j][j] = 1.0
    return [row[n:] for row in m]


def lua_matrix_mul(A, B):
    """matrix.mul (lua-matrix/matrix.lua:223-237): plain triple loop, left-to-right accumulation."""
    n, k, mcols = len(A), len(B), len(B[0])
    out = [[0.0] * mcols for _ in range(n)]
    for i in range(n):
        for j in range(mcols):
            num = A[i][0] * B[0][j]
            for t in range(1, k):
                num = num + A[i][t] * B[t][j]
            out[i][j] = num
    return out
