"""ORACLE (test infrastructure, not product code) -- the UG4-style object API on NumPy/SciPy.

Same names and call semantics as admm_optim_b200/ug4.py (which mirrors the Lua-registered objects of
3d_admm.lua / 2d_admm.lua), so that one driver replay runs on both and the traces can be diffed.
PARITY UNPINNED by the reference (no UG4 here): this restates the inferred model of
oracle/fem_np.py.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import fem_np as F
from . import mesh_np as M

PST_CONSISTENT, PST_ADDITIVE = 1, 2


class OracleError(RuntimeError):
    pass


class Domain:
    def __init__(self):
        self.levels = None
        self.dim = None
        self.coords_version = 0

    @property
    def top(self):
        return self.levels[-1]

    def num_levels(self):
        return len(self.levels)

    class _Info:
        def __init__(self, dom):
            self.dom = dom

        def num_surface_elements(self):
            return self.dom.top.ne

        def to_string(self):
            return "\n".join("lvl %d: nv=%d ne=%d" % (i, l.nv, l.ne) for i, l in enumerate(self.dom.levels))

    def domain_info(self):
        return Domain._Info(self)


class ApproximationSpace:
    def __init__(self, dom):
        self.dom, self.names, self.kind = dom, [], None

    def add_fct(self, names, fe_type, order=None):
        self.kind = {"Lagrange": 1, "Piecewise-Constant": 0}[fe_type]
        self.names += [n.strip() for n in names.split(",")]

    def init_levels(self): pass
    def init_top_surface(self): pass
    def print_statistic(self): print("  oracle space %s: %d dofs" % (",".join(self.names), self.num_dofs()))

    def num_dofs(self):
        t = self.dom.top
        return (t.nv if self.kind == 1 else t.ne) * len(self.names)

    def fct_index(self, name):
        return self.names.index(name)


class GridFunction:
    def __init__(self, space):
        self.space = space
        self.v = np.zeros(space.num_dofs())
        self.storage = PST_CONSISTENT

    def set(self, c):
        self.v[:] = c
        self.storage = PST_CONSISTENT

    def has_storage_type_additive(self): return bool(self.storage & PST_ADDITIVE)
    def change_storage_type_to_consistent(self): self.storage = PST_CONSISTENT
    def change_storage_type_to_additive(self): self.storage = PST_ADDITIVE

    def from_numpy(self, a, storage=PST_CONSISTENT):
        self.v[:] = np.asarray(a, float).ravel()
        self.storage = storage

    def to_numpy(self, out=None):
        if out is None:
            return self.v.copy()
        out[:] = self.v
        return out


class _Import:
    def __init__(self, gf, fct, what):
        self.gf, self.comp, self.what = gf, gf.space.fct_index(fct), what


class ElemDisc:
    def __init__(self, ug, class_name, fcts, subsets):
        self.ug, self.kind = ug, class_name
        self.p = dict(lambda_vol=0.0, lambda_bary=[0.0, 0.0, 0.0], step_length=1.0, tau=1.0, index=1,
                      mult=[0.0, 0.0, 0.0, 0.0], second_order=False)
        self.imp = {}

    def set_quad_order(self, o): pass
    def set_lambda_vol(self, v): self.p["lambda_vol"] = float(v)
    def set_lambda_barycenter(self, x, y, z=0.0): self.p["lambda_bary"] = [float(x), float(y), float(z)]
    def set_step_length(self, v): self.p["step_length"] = float(v)
    def set_tau(self, v): self.p["tau"] = float(v)
    def set_index(self, k): self.p["index"] = int(k)
    def set_multiplier_vol(self, v): self.p["mult"][0] = float(v)
    def set_multiplier_bx(self, v): self.p["mult"][1] = float(v)
    def set_multiplier_by(self, v): self.p["mult"][2] = float(v)
    def set_multiplier_bz(self, v): self.p["mult"][3] = float(v)
    def set_scaling(self, v): pass
    def set_high_order_scaling(self, v): pass
    def set_second_order(self, b): self.p["second_order"] = bool(b)

    def __getattr__(self, name):
        if name.startswith("set_deformation"):
            return lambda imp: self.imp.__setitem__("u", imp.gf)
        if name.startswith("set_lambda"):
            return lambda imp: self.imp.__setitem__("lam", imp.gf)
        if name.startswith("set_q"):
            return lambda imp: self.imp.__setitem__("q", imp.gf)
        raise AttributeError(name)


class DirichletBoundary:
    def __init__(self):
        self.entries = []

    def add(self, value, fct, subset):
        self.entries.append((float(value), fct, subset))


class DomainDiscretization:
    def __init__(self, ug, space):
        self.ug, self.space, self.discs, self.dir = ug, space, [], []

    def add(self, obj):
        if isinstance(obj, ElemDisc):
            self.discs.append(obj)
        else:
            self.dir += [(self.space.fct_index(f), s) for _, f, s in obj.entries]

    def dmask(self, mesh):
        d = len(self.space.names)
        m = np.zeros((mesh.nv, d), bool)
        for comp, subset in self.dir:
            m[mesh.vertex_mask(subset), comp] = True
        return m.ravel()

    def _u(self, disc, uarg):
        gf = disc.imp.get("u", uarg)
        return gf.v if gf is not None else None

    def assemble_jacobian(self, A, u):
        mesh = self.space.dom.top
        d = mesh.dim
        for disc in self.discs:
            if disc.kind == "DeformationEquation":
                if disc.p["second_order"]:
                    raise OracleError("second-order (J'') terms are outside the hot path")
                uu = self._u(disc, u)
                has_lam = disc.p["lambda_vol"] != 0.0 or any(v != 0.0 for v in disc.p["lambda_bary"][:d])
                # the six operators of one Newton iteration are identical (same Hessian disc + Dirichlet set):
                # assemble once, like the GPU path's signature cache (DESIGN.md "Operator sharing")
                key = (disc.p["step_length"], disc.p["lambda_vol"], tuple(disc.p["lambda_bary"][:d]),
                       hash(uu.tobytes()) if (has_lam and uu is not None) else 0, self.space.dom.coords_version, tuple(sorted(self.dir)))
                cache = self.ug._asm_cache
                if cache.get("key") != key:
                    cache["key"] = key
                    asm = F.hessian_matrix_fast if self.ug.fast_assembly else F.hessian_matrix
                    cache["mat"] = asm(mesh, uu, c=disc.p["step_length"], lam_vol=disc.p["lambda_vol"],
                                       lam_bary=disc.p["lambda_bary"][:d], dmask=self.dmask(mesh) if self.dir else None)
                A.mat = cache["mat"]
                A.diag = None
                return
            if disc.kind == "MassModel":
                A.diag, _ = F.mass_model(mesh, self._u(disc, u), np.zeros(mesh.ne * d * d))
                A.mat = None
                return
        raise OracleError("no jacobian-contributing ElemDisc")

    def assemble_defect(self, dvec, u):
        mesh = self.space.dom.top
        d = mesh.dim
        out = np.zeros_like(dvec.v)
        for disc in self.discs:
            uu = self._u(disc, u)
            k = disc.kind
            if k == "DeformationEquation":
                continue
            fast = self.ug.fast_assembly
            if k in ("DeformationEquationRHS", "DeformationEquationLargeProblemRHS"):
                w = np.array([disc.p["lambda_vol"]] + disc.p["lambda_bary"][:d])
                if k == "DeformationEquationLargeProblemRHS":
                    w = w + np.array(disc.p["mult"][:d + 1])
                sgn = 1.0 if d == 3 else -1.0                                    # sign conventions: DESIGN.md 'Signs'
                if fast:
                    out += F.load_vector_fast(mesh, uu, disc.imp["lam"].v, disc.imp["q"].v, disc.p["tau"], w, sgn)
                else:
                    lam = disc.imp["lam"].v.reshape(-1, d, d)
                    q = disc.imp["q"].v.reshape(-1, d, d)
                    G, _, _ = F.geometry(mesh)
                    S = lam + disc.p["tau"] * (F.grad_u(mesh, G, uu) - q)
                    out += F.load_vector(mesh, uu, S, w, sgn)
            elif k in ("VolumeConstraintSecondDerivative", "SecondDerivativeVolume"):
                w = np.zeros(d + 1); w[0] = 1.0
                out += (F.load_vector_fast(mesh, uu, None, None, 0.0, w, -1.0 if d == 3 else 1.0) if fast
                        else F.load_vector(mesh, uu, None, w, -1.0 if d == 3 else 1.0))
            elif k in ("SecondDerivativeBarycenter", "XBarycenterConstraintSecondDerivative"):
                w = np.zeros(d + 1); w[disc.p["index"]] = 1.0
                out += (F.load_vector_fast(mesh, uu, None, None, 0.0, w, -1.0 if d == 3 else 1.0) if fast
                        else F.load_vector(mesh, uu, None, w, -1.0 if d == 3 else 1.0))
            elif k == "MassModel":
                _, rhs = F.mass_model(mesh, uu, disc.imp["lam"].v)
                out += rhs
            elif k == "LambdaUpdate":
                out += F.lambda_update_defect(mesh, uu, disc.imp["q"].v, disc.p["tau"])
            else:
                raise OracleError("unknown disc " + k)
        if self.dir:
            out[self.dmask(mesh)] = 0.0
        dvec.v[:] = out
        dvec.storage = PST_ADDITIVE

    def adjust_solution(self, u):
        if self.dir:
            u.v[self.dmask(self.space.dom.top)] = 0.0


class AssembledLinearOperator:
    def __init__(self, dd):
        self.dd, self.mat, self.diag = dd, None, None

    def apply(self, y, x):
        y.v[:] = self.mat @ x.v if self.mat is not None else self.diag * x.v
        y.storage = PST_ADDITIVE

    def to_scipy(self):
        return self.mat if self.mat is not None else sp.diags(self.diag)


class ConvCheck:
    def __init__(self, max_its=100, abs_tol=1e-12, reduction=1e-12, verbose=False):
        self.max_its, self.abs_tol, self.reduction, self.verbose = int(max_its), float(abs_tol), float(reduction), verbose


class Jacobi:
    def __init__(self, damp=1.0):
        self.damp = damp


class SuperLU:
    pass


class CG:
    def __init__(self):
        self.precond, self.cc, self.A, self.steps = Jacobi(1.0), ConvCheck(), None, 0

    def set_preconditioner(self, p): self.precond = p
    def set_convergence_check(self, cc): self.cc = cc

    def init(self, A, x=None):
        self.A = A
        return True

    def apply(self, x, b):
        sol, ok, its = F.cg_jacobi(self.A.diag, b.v, x.v, self.precond.damp, self.cc.abs_tol, self.cc.max_its)
        x.v[:] = sol
        x.storage = PST_CONSISTENT
        self.steps = its
        return ok

    def step(self):
        return self.steps


class BiCGStabGMG:
    def __init__(self, ug, desc):
        self.ug, self.desc = ug, desc
        self.A, self.gmg, self.steps, self.last_defect = None, None, 0, 0.0

    def init(self, A, x=None):
        self.A = A
        pre = self.desc["precond"]
        dom = A.dd.space.dom
        cache = self.ug._gmg_cache
        if self.ug.c_solver:
            # compiled solve path (oracle/solver_c.c): the static part lives with the mesh hierarchy + Dirichlet set, every new
            # matrix triggers what the reference does at solver:init (RAP chain, smoother data, base factorisation)
            skey = ("c", id(dom), tuple(sorted(A.dd.dir)), self.ug.smoother, self.ug.threads)
            if skey not in cache:
                dmasks = [A.dd.dmask(l) if A.dd.dir else np.zeros(l.nv * l.dim, bool) for l in dom.levels]
                cache[skey] = [F.GMGC(dom.levels, dmasks, smoother=self.ug.smoother, nu1=pre.get("preSmooth", 3), nu2=pre.get("postSmooth", 3),
                                      cheb_ratio=self.ug.cheb_ratio, threads=self.ug.threads), None]
            ent = cache[skey]
            if ent[1] is not A.mat:
                ent[0].setup(A.mat)
                ent[1] = A.mat
            self.gmg = ent[0]
            return True
        dmasks = [A.dd.dmask(l) if A.dd.dir else np.zeros(l.nv * l.dim, bool) for l in dom.levels]
        key = (id(A.mat), self.ug.smoother)
        if key not in cache:
            cache.clear()
            cache[key] = F.GMG(dom.levels, A.mat, dmasks, smoother=self.ug.smoother, nu1=pre.get("preSmooth", 3),
                               nu2=pre.get("postSmooth", 3), cheb_ratio=self.ug.cheb_ratio, threads=self.ug.threads)
        self.gmg = cache[key]
        return True

    def _solve(self, x, b):
        cc = self.desc["convCheck"]
        if self.ug.c_solver:
            sol, ok, its, r = self.gmg.solve(b.v, x.v, abs_tol=cc["absolute"], max_it=cc["iterations"], red_tol=cc.get("reduction", 0.0))
        else:
            sol, ok, its, r = F.bicgstab(self.A.mat, b.v, x.v, self.gmg.apply, abs_tol=cc["absolute"], max_it=cc["iterations"],
                                         red_tol=cc.get("reduction", 0.0))
        x.v[:] = sol
        x.storage = PST_CONSISTENT
        self.steps, self.last_defect = its, float(np.linalg.norm(r))
        return ok, r

    def apply(self, x, b):
        return self._solve(x, b)[0]

    def apply_return_defect(self, x, b):
        ok, r = self._solve(x, b)
        b.v[:] = r
        return ok

    def step(self):
        return self.steps

    def defect(self):
        return self.last_defect

    def vcycle(self, z, r):
        """z = one V-cycle applied to r (the preconditioner alone; mirrors ab_solver_vcycle of the CUDA library)."""
        z.v[:] = self.gmg.apply(r.v)
        z.storage = PST_CONSISTENT


class _NS:
    pass


class Backend:
    """NumPy twin of admm_optim_b200.ug4.Backend. `smoother`: 'cheb' | 'jac' (what the CUDA path runs) or
    'gs' (lexicographic Gauss-Seidel, what the reference asks for -- iteration counts side by side)."""
    name = "oracle"

    def __init__(self, smoother="cheb", cheb_ratio=6.0, threads=1, fast_assembly=False, c_solver=False):
        """fast_assembly: element loops of the P1 assembly in C (oracle_kernels.c) instead of NumPy -- used by bench.py's CPU
        legs so that the CPU baseline is not dominated by NumPy temporaries; the tests keep the NumPy path as the checker.
        c_solver: solver:init + solver:apply (RAP chain, V-cycle, BiCGStab) in C / OpenMP on `threads` threads (oracle/solver_c.c),
        bench.py's multi-threaded CPU baseline; pinned against the NumPy statement by tests/test_oracle.py."""
        self.dim = None
        self.smoother, self.cheb_ratio, self.threads = smoother, cheb_ratio, threads
        self.fast_assembly = bool(fast_assembly)
        self.c_solver = bool(c_solver)
        if self.fast_assembly and self.c_solver and F.c_kernels() is not None and hasattr(F.c_kernels(), "oracle_set_elem_threads"):
            F.c_kernels().oracle_set_elem_threads(int(threads))     # CPU-baseline mode: the element loops run on the same threads
        self._gmg_cache = {}
        self._asm_cache = {}
        self.util = _NS()
        self.util.refinement = _NS()
        self.util.refinement.CreateRegularHierarchy = self._refine
        self.util.solver = _NS()
        self.util.solver.CreateSolver = lambda desc: BiCGStabGMG(self, desc)

    def InitUG(self, dim, algebra=None): self.dim = dim
    def AlgebraType(self, n, b): return (n, b)
    def synchronize(self): pass
    def launch_count(self): return 0

    def Domain(self): return Domain()

    def LoadDomain(self, dom, name):
        m = M.load_npz(name) if name.endswith(".npz") else M.load_ugx(name)
        dom.levels, dom.dim = [m], m.dim
        if self.dim is None:
            self.dim = m.dim

    def _refine(self, dom, num_refs, verbose=False, desc=None):
        dom.levels = M.build_hierarchy(dom.levels[0], num_refs)

    def ApproximationSpace(self, dom): return ApproximationSpace(dom)
    def GridFunction(self, space): return GridFunction(space)
    AdvancedGridFunction = GridFunction
    def GlobalGridFunctionNumberData(self, gf, fct): return _Import(gf, fct, "value")
    def GlobalGridFunctionGradientData(self, gf, fct): return _Import(gf, fct, "gradient")
    def DirichletBoundary(self): return DirichletBoundary()
    def DomainDiscretization(self, space): return DomainDiscretization(self, space)
    def AssembledLinearOperator(self, dd): return AssembledLinearOperator(dd)

    def __getattr__(self, name):
        if name in ("DeformationEquation", "DeformationEquationRHS", "DeformationEquationLargeProblemRHS",
                    "VolumeConstraintSecondDerivative", "SecondDerivativeVolume", "SecondDerivativeBarycenter",
                    "XBarycenterConstraintSecondDerivative", "MassModel", "LambdaUpdate"):
            return lambda fcts, subsets: ElemDisc(self, name, fcts, subsets)
        raise AttributeError(name)

    def CG(self): return CG()
    Jacobi = staticmethod(Jacobi)
    ConvCheck = staticmethod(ConvCheck)
    SuperLU = staticmethod(SuperLU)

    def VecScaleAssign(self, dst, a, src):
        dst.v[:] = a * src.v
        dst.storage = src.storage

    def VecScaleAdd2(self, dst, a, x, b, y):
        dst.v[:] = a * x.v + b * y.v
        dst.storage = x.storage if x.storage == y.storage else (x.storage & y.storage or x.storage)

    def VecProd(self, x, y): return float(x.v @ y.v)
    def VecProdMulti(self, xs, y): return [float(x.v @ y.v) for x in xs]
    def VecNorm(self, x): return float(np.linalg.norm(x.v))

    def L2Norm(self, gf, fct, quad_order=None, subsets=None):
        mesh = gf.space.dom.top
        c = gf.space.fct_index(fct)
        return F.l2norm_p1(mesh, gf.v, c) if gf.space.kind == 1 else F.l2norm_p0(mesh, gf.v, c)

    def L2NormAll(self, gf):
        return [self.L2Norm(gf, n) for n in gf.space.names]

    def Testing(self, qp, q, cmps, sigma):
        qp.v[:] = F.project_frobenius(q.v, sigma, q.space.dom.dim)
        qp.storage = q.storage

    def ProjectWithSpectralNorm(self, qp, q, cmps, sigma):
        qp.v[:] = F.project_spectral(q.v, sigma)
        qp.storage = q.storage

    def MaximumFrobeniusNorm(self, u, cmps, subsets, qo): return F.max_frobenius_norm(u.space.dom.top, u.v)
    def MaxSpectralNorm(self, u, cmps, subsets, qo): return F.max_spectral_norm(u.space.dom.top, u.v)
    def VolumeDefect(self, u, vref, subsets, cmps, qo, *unused): return F.volume_defect(u.space.dom.top, u.v, vref)
    def BarycenterDefect(self, u, cmps, subsets, qo): return list(F.barycenter_defect(u.space.dom.top, u.v))

    def SetZeroAwayFromSubset(self, gf, cmps, subset):
        mesh = gf.space.dom.top
        keep = np.repeat(mesh.vertex_mask(subset), len(gf.space.names))
        gf.v[~keep] = 0.0

    def TransformDomainByDisplacement(self, u, cmps):
        dom = u.space.dom
        d = dom.dim
        disp = u.v.reshape(-1, d)
        for l in dom.levels:
            l.xyz += disp[: l.nv]
        dom.coords_version += 1
