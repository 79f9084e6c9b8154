// UG4 plugin shim: registers the Lua-visible objects the reference drivers call on the deformation / extension hot path
// (3d_admm.lua / 2d_admm.lua / obstacle_optim_*_util.lua) and forwards every call to the C ABI of libadmm_b200
// (include/admm_b200.h).  Built inside a UG4 tree as plugins/ADMMOptimB200 (CMake: -DADMMOptimB200=ON) and loaded by
// ugshell from bin/plugins [UPSTREAM-UNVERIFIED, SURVEY.md App. C14]; in this repository it is compiled against the stand-in
// tests/ug4_stub/bridge/util.h and its registrations are checked by tests/test_host.py (no UG4 tree exists in the image).
//
// Names.  The classes and functions of the plugins this one REPLACES on the hot path (ADMMOptim, FluidOptim's deformation
// part, PLaplacian: 3d_admm.lua:1-3) keep their Lua names: DeformationEquation, DeformationEquationRHS, ..., Testing,
// MaximumFrobeniusNorm, VolumeDefect ... (3d_admm.lua:393-694, 910-916, 1167-1168).  Objects whose names belong to ugcore
// (Domain, ApproximationSpace, GridFunction, DomainDiscretization, GeometricMultiGrid, BiCGStab, VecProd ...) are registered
// with the prefix "B200"; admm_b200_prelude.lua binds the ugcore names to them for the deformation objects before it runs the
// UNCHANGED driver script, so that the Navier-Stokes / adjoint part keeps UG4's CPU objects (out of scope).
#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "bridge/util.h"

#include "admm_b200.h"

namespace ug {
namespace ADMMOptimB200 {

static void check(int rc) {
    if (rc != AB_OK) UG_THROW("ADMMOptimB200: " << ab_last_error());
}

static std::vector<std::string> tokenize(const char* csv) {
    std::vector<std::string> out;
    std::string cur;
    for (const char* p = csv; *p; ++p) {
        if (*p == ',') { out.push_back(cur); cur.clear(); }
        else if (*p != ' ' && *p != '\t') cur.push_back(*p);
    }
    out.push_back(cur);
    return out;
}

// one GPU context per ugshell process (one MPI rank = one GPU); device from ADMM_B200_DEVICE or the local rank
static ab_context* session() {
    static ab_context* ctx = nullptr;
    if (!ctx) {
        const char* d = std::getenv("ADMM_B200_DEVICE");
        if (!d) d = std::getenv("OMPI_COMM_WORLD_LOCAL_RANK");
        check(ab_context_create(d ? std::atoi(d) : 0, nullptr, &ctx));
    }
    return ctx;
}

// ---- Domain() + LoadDomain(dom, gridName) + util.refinement.CreateRegularHierarchy   3d_admm.lua:108-109,186 ----------------
class B200Domain {
 public:
    ab_domain* h = nullptr;
    ~B200Domain() { if (h) ab_domain_destroy(h); }
    void load(const char* file) { check(ab_domain_load_ugx(session(), file, &h)); }
    void refine(int num_refs) { check(ab_domain_refine(h, num_refs)); }
    int num_surface_elements() {                                                              // dom:domain_info()  3d:112
        int nl = 0, ne = 0;
        check(ab_domain_num_levels(h, &nl));
        check(ab_domain_level_info(h, nl - 1, nullptr, nullptr, &ne, nullptr, nullptr));
        return ne;
    }
    int dim() { int d = 0; check(ab_domain_level_info(h, 0, &d, nullptr, nullptr, nullptr, nullptr)); return d; }
};
static void LoadDomain(B200Domain& dom, const char* file) { dom.load(file); }
static void CreateRegularHierarchy(B200Domain& dom, int num_refs) { dom.refine(num_refs); }

// ---- ApproximationSpace(dom):add_fct / init_levels / init_top_surface   3d_admm.lua:329-333,367-370 ---------------------------
class B200ApproximationSpace {
 public:
    SmartPtr<B200Domain> dom;
    ab_space* h = nullptr;
    std::vector<std::string> names;
    int kind = -1;
    explicit B200ApproximationSpace(SmartPtr<B200Domain> d) : dom(d) {}
    ~B200ApproximationSpace() { if (h) ab_space_destroy(h); }
    void add_fct(const char* fcts, const char* type, int order) {
        const int k = std::strcmp(type, "Lagrange") == 0 ? AB_SPACE_P1 : (std::strcmp(type, "Piecewise-Constant") == 0 ? AB_SPACE_P0 : -1);
        if (h || k < 0 || (k == AB_SPACE_P1 && order != 1) || (kind >= 0 && kind != k))
            UG_THROW("ADMMOptimB200: the GPU backend serves ('Lagrange',1) and 'Piecewise-Constant' spaces (one type per space)");
        kind = k;
        for (const std::string& n : tokenize(fcts)) names.push_back(n);
    }
    void add_fct_p0(const char* fcts, const char* type) { add_fct(fcts, type, 0); }
    void ensure() { if (!h) check(ab_space_create(dom->h, kind, (int)names.size(), &h)); }
    void init_levels() { ensure(); }
    void init_top_surface() { ensure(); }
    void print_statistic() { ensure(); }
    int fct_index(const std::string& n) const {
        for (size_t i = 0; i < names.size(); ++i) if (names[i] == n) return (int)i;
        UG_THROW("ADMMOptimB200: unknown function '" << n << "'");
    }
};

// ---- GridFunction / AdvancedGridFunction   3d_admm.lua:337-341,375-383 --------------------------------------------------------
class B200GridFunction {
 public:
    SmartPtr<B200ApproximationSpace> space;
    ab_vector* h = nullptr;
    explicit B200GridFunction(SmartPtr<B200ApproximationSpace> s) : space(s) { s->ensure(); check(ab_vector_create(s->h, &h)); }
    ~B200GridFunction() { if (h) ab_vector_destroy(h); }
    void set(number c) { check(ab_vector_set(h, c)); }                                                          // 3d:951
    int storage() const { int s = 0; check(ab_vector_storage(h, &s)); return s; }
    bool has_storage_type_additive() { return (storage() & AB_PST_ADDITIVE) != 0; }                               // 3d:978
    bool has_storage_type_consistent() { return (storage() & AB_PST_CONSISTENT) != 0; }
    void change_storage_type_to_consistent() { check(ab_vector_change_storage(h, AB_PST_CONSISTENT)); }           // 3d:912,982,1096
    void change_storage_type_to_additive() { check(ab_vector_change_storage(h, AB_PST_ADDITIVE)); }
    // the boundary to the UG4/CPU side: J' enters (3d:816-817), u leaves (3d:1333); n = number of dofs on this rank
    void assign_from_host(const number* values, int storage_type) { check(ab_vector_upload(h, values, storage_type)); }
    void copy_to_host(number* values) { check(ab_vector_download(h, values)); }
    size_t num_dofs() { int64_t n = 0; check(ab_space_num_dofs(space->h, &n)); return (size_t)n; }
};

// ---- GlobalGridFunctionNumberData / GlobalGridFunctionGradientData   3d_admm.lua:343-363,384-389 ---------------------------------
class B200GridFunctionData {
 public:
    SmartPtr<B200GridFunction> gf;   // null: a field of the UG4/CPU side (Navier-Stokes / adjoint imports of the 2D Hessian, 2d:396-419)
    int comp = -1;
    bool gradient = false;
    B200GridFunctionData(SmartPtr<B200GridFunction> g, const char* fct, bool grad) : gf(g), gradient(grad) { comp = g->space->fct_index(fct); }
    B200GridFunctionData() {}
};
class B200GridFunctionNumberData : public B200GridFunctionData {
 public:
    B200GridFunctionNumberData(SmartPtr<B200GridFunction> g, const char* fct) : B200GridFunctionData(g, fct, false) {}
    B200GridFunctionNumberData() {}
};
class B200GridFunctionGradientData : public B200GridFunctionData {
 public:
    B200GridFunctionGradientData(SmartPtr<B200GridFunction> g, const char* fct) : B200GridFunctionData(g, fct, true) {}
    B200GridFunctionGradientData() {}
};

// ---- element discretisations   3d_admm.lua:393-694 ----------------------------------------------------------------------------
class B200ElemDisc {
 public:
    int kind;
    std::vector<std::string> fcts;
    ab_elemdisc* h = nullptr;
    B200ApproximationSpace* space = nullptr;
    std::map<int, double> params;
    std::map<int, SmartPtr<B200GridFunction>> imports;
    B200ElemDisc(int k, const char* functions, const char* subsets) : kind(k), fcts(tokenize(functions)) {
        if (tokenize(subsets) != std::vector<std::string>{"outer"}) UG_THROW("ADMMOptimB200: element discs are assembled on subset 'outer' (3d_admm.lua:393)");
    }
    virtual ~B200ElemDisc() { if (h) ab_elemdisc_destroy(h); }
    // the C object needs the space, known when the disc joins a DomainDiscretization
    void attach(B200ApproximationSpace& s) {
        if (h) { if (&s != space) UG_THROW("ADMMOptimB200: ElemDisc added to DomainDiscretizations of different spaces"); return; }
        if (fcts != s.names) UG_THROW("ADMMOptimB200: ElemDisc functions do not match the ApproximationSpace");
        space = &s;
        s.ensure();
        check(ab_elemdisc_create(s.h, kind, &h));
        for (auto& p : params) check(ab_elemdisc_set_param(h, p.first, p.second));
        for (auto& i : imports) check(ab_elemdisc_bind(h, i.first, i.second->h));
    }
    void param(int id, double v) { params[id] = v; if (h) check(ab_elemdisc_set_param(h, id, v)); }
    void bind(int which, SmartPtr<B200GridFunctionData> d, int comp, bool gradient) {
        if (!d->gf) return;                                    // a UG4/CPU field: not consumed on the hot path
        if (d->comp != comp || d->gradient != gradient) UG_THROW("ADMMOptimB200: only the canonical import wiring of the scripts is supported");
        auto it = imports.find(which);
        if (it != imports.end() && it->second.get() != d->gf.get()) UG_THROW("ADMMOptimB200: all components of one import must come from the same grid function");
        imports[which] = d->gf;
        if (h) check(ab_elemdisc_bind(h, which, d->gf->h));
    }
    int dim() const { return kind >= AB_DISC_MASS_MODEL ? (fcts.size() == 4 ? 2 : 3) : (int)fcts.size(); }
    // scalar setters (3d:393-396, 411, 473, 578, 1081-1084; 2d:389-394)
    void set_quad_order(int o) { param(AB_PARAM_QUAD_ORDER, o); }
    void set_lambda_vol(number v) { param(AB_PARAM_LAMBDA_VOL, v); }
    void set_lambda_barycenter(number x, number y, number z) { param(AB_PARAM_LAMBDA_BARY_X, x); param(AB_PARAM_LAMBDA_BARY_Y, y); param(AB_PARAM_LAMBDA_BARY_Z, z); }
    void set_step_length(number v) { param(AB_PARAM_STEP_LENGTH, v); }
    void set_tau(number v) { param(AB_PARAM_TAU, v); }
    void set_index(int k) { param(AB_PARAM_INDEX, k); }
    void set_multiplier_vol(number v) { param(AB_PARAM_MULT_VOL, v); }
    void set_multiplier_bx(number v) { param(AB_PARAM_MULT_BX, v); }
    void set_multiplier_by(number v) { param(AB_PARAM_MULT_BY, v); }
    void set_multiplier_bz(number v) { param(AB_PARAM_MULT_BZ, v); }
    void set_scaling(number v) { param(AB_PARAM_SCALING, v); }
    void set_high_order_scaling(number v) { param(AB_PARAM_HIGH_ORDER_SCALING, v); }
    void set_second_order(bool b) { param(AB_PARAM_SECOND_ORDER, b ? 1.0 : 0.0); }
    void set_kinematic_viscosity(number) {}                    // 2d:395: used by the J'' terms only (set_second_order(true) is unsupported)
    // imports: deformation value / gradient per component (3d:399-405)
#define AB_DEF_IMPORT(k)                                                                                                    \
    void set_deformation_d##k(SmartPtr<B200GridFunctionData> d) { bind(AB_IMPORT_DEFORMATION, d, k - 1, false); }           \
    void set_deformation_vector_d##k(SmartPtr<B200GridFunctionData> d) { bind(AB_IMPORT_DEFORMATION, d, k - 1, true); }     \
    void set_velocity_d##k(SmartPtr<B200GridFunctionData>) {}                                                               \
    void set_velocity_vector_d##k(SmartPtr<B200GridFunctionData>) {}                                                        \
    void set_adjoint_velocity_d##k(SmartPtr<B200GridFunctionData>) {}                                                       \
    void set_adjoint_velocity_vector_d##k(SmartPtr<B200GridFunctionData>) {}
    AB_DEF_IMPORT(1) AB_DEF_IMPORT(2) AB_DEF_IMPORT(3)
#undef AB_DEF_IMPORT
    void set_pressure(SmartPtr<B200GridFunctionData>) {}
    void set_adjoint_pressure(SmartPtr<B200GridFunctionData>) {}
    // tensor imports, components row-major l1..l(d*d) (3d:343-363, 423-442, 683-691)
#define AB_DEF_TENSOR(i, j)                                                                                                 \
    void set_lambda##i##j(SmartPtr<B200GridFunctionData> d) { bind(AB_IMPORT_LAMBDA, d, i * dim() + j, false); }            \
    void set_q##i##j(SmartPtr<B200GridFunctionData> d) { bind(AB_IMPORT_Q, d, i * dim() + j, false); }                      \
    void set_qproj##i##j(SmartPtr<B200GridFunctionData> d) { bind(AB_IMPORT_Q, d, i * dim() + j, false); }
    AB_DEF_TENSOR(0, 0) AB_DEF_TENSOR(0, 1) AB_DEF_TENSOR(0, 2) AB_DEF_TENSOR(1, 0) AB_DEF_TENSOR(1, 1) AB_DEF_TENSOR(1, 2)
    AB_DEF_TENSOR(2, 0) AB_DEF_TENSOR(2, 1) AB_DEF_TENSOR(2, 2)
#undef AB_DEF_TENSOR
};
template <int KIND, int TAG = 0>      // TAG: two Lua names may share one kind but need distinct C++ types (one registry entry per type)
class B200ElemDiscT : public B200ElemDisc {
 public:
    B200ElemDiscT(const char* functions, const char* subsets) : B200ElemDisc(KIND, functions, subsets) {}
};
typedef B200ElemDiscT<AB_DISC_DEFORMATION_EQUATION> DeformationEquation;                       // 3d:393
typedef B200ElemDiscT<AB_DISC_DEFORMATION_RHS> DeformationEquationRHS;                         // 3d:407
typedef B200ElemDiscT<AB_DISC_DEFORMATION_LARGE_RHS> DeformationEquationLargeProblemRHS;       // 3d:472
typedef B200ElemDiscT<AB_DISC_VOLUME_CONSTRAINT> VolumeConstraintSecondDerivative;             // 3d:559
typedef B200ElemDiscT<AB_DISC_VOLUME_CONSTRAINT, 1> SecondDerivativeVolume;                    // 2d:564
typedef B200ElemDiscT<AB_DISC_BARYCENTER_CONSTRAINT> SecondDerivativeBarycenter;               // 3d:577,596 / 2d:580,597
typedef B200ElemDiscT<AB_DISC_BARYCENTER_CONSTRAINT, 1> XBarycenterConstraintSecondDerivative; // 3d:616 (the z component uses this class name)
typedef B200ElemDiscT<AB_DISC_MASS_MODEL> MassModel;                                           // 3d:652
typedef B200ElemDiscT<AB_DISC_LAMBDA_UPDATE> LambdaUpdate;                                     // 3d:677

// ---- DirichletBoundary / DomainDiscretization / AssembledLinearOperator   3d_admm.lua:445-467 ---------------------------------
class B200DirichletBoundary {
 public:
    struct Entry { number value; std::string fct, subset; };
    std::vector<Entry> entries;
    void add(number value, const char* fct, const char* subset) { entries.push_back({value, fct, subset}); }
};
class B200AssembledLinearOperator;
class B200DomainDiscretization {
 public:
    SmartPtr<B200ApproximationSpace> space;
    ab_domaindisc* h = nullptr;
    std::vector<SmartPtr<B200ElemDisc>> keep;
    explicit B200DomainDiscretization(SmartPtr<B200ApproximationSpace> s) : space(s) { s->ensure(); check(ab_domaindisc_create(s->h, &h)); }
    ~B200DomainDiscretization() { if (h) ab_domaindisc_destroy(h); }
    void add(SmartPtr<B200ElemDisc> d) { d->attach(*space); check(ab_domaindisc_add_elemdisc(h, d->h)); keep.push_back(d); }
    void add_dirichlet(SmartPtr<B200DirichletBoundary> b) {
        for (auto& e : b->entries) check(ab_domaindisc_add_dirichlet(h, e.subset.c_str(), space->fct_index(e.fct), e.value));
    }
    void assemble_jacobian(B200AssembledLinearOperator& A, B200GridFunction& u);                                      // 3d:972
    void assemble_defect(B200GridFunction& d, B200GridFunction& u) { check(ab_domaindisc_assemble_defect(h, d.h, u.h)); }   // 3d:973
    void adjust_solution(B200GridFunction& u) { check(ab_domaindisc_adjust_solution(h, u.h)); }                       // 3d:971
};
class B200AssembledLinearOperator {
 public:
    SmartPtr<B200DomainDiscretization> dd;
    ab_operator* h = nullptr;
    explicit B200AssembledLinearOperator(SmartPtr<B200DomainDiscretization> d) : dd(d) { check(ab_operator_create(d->h, &h)); }
    ~B200AssembledLinearOperator() { if (h) ab_operator_destroy(h); }
    void apply(B200GridFunction& y, B200GridFunction& x) { check(ab_operator_apply(h, y.h, x.h)); }
};
void B200DomainDiscretization::assemble_jacobian(B200AssembledLinearOperator& A, B200GridFunction& u) {
    check(ab_domaindisc_assemble_jacobian(h, A.h, u.h));
}

// ---- solver components   obstacle_optim_3d_util.lua:9-43, 159-172; 3d_admm.lua:701-703 ----------------------------------------
class B200ConvCheck {
 public:
    int max_its; number abs_tol, reduction; bool verbose;
    B200ConvCheck(int m, number a, number r, bool v) : max_its(m), abs_tol(a), reduction(r), verbose(v) {}
    B200ConvCheck() : max_its(100), abs_tol(1e-12), reduction(1e-12), verbose(false) {}
};
class B200Jacobi { public: number damp; explicit B200Jacobi(number d) : damp(d) {} B200Jacobi() : damp(1.0) {} };
class B200GaussSeidel {       // smoother = "gs" (u3:16): served by the stated GPU equivalent (Chebyshev-Jacobi)
 public:
    number damp = 1.0;
    void set_damp(number d) { damp = d; }          // u3:160 (linear_solver_damping, unused by the drivers)
};
class B200SuperLU {};         // baseSolver = SuperLU() (u3:21): served by the dense coarse inverse
class B200StdTransfer {};     // transfer = "std" (u3:28)
class B200GeometricMultiGrid {        // setters of obstacle_optim_3d_util.lua:159-172
 public:
    SmartPtr<B200ApproximationSpace> space;
    ab_gmg_desc desc;
    explicit B200GeometricMultiGrid(SmartPtr<B200ApproximationSpace> s) : space(s) {
        std::memset(&desc, 0, sizeof desc);
        desc.smoother = AB_SMOOTHER_CHEBYSHEV; desc.pre_smooth = 2; desc.post_smooth = 2; desc.rap = 0;
    }
    void set_base_level(int l) { desc.base_level = l; }
    void set_base_solver(SmartPtr<B200SuperLU>) {}
    void set_gathered_base_solver_if_ambiguous(bool) {}
    void set_smoother(SmartPtr<B200GaussSeidel>) { desc.smoother = AB_SMOOTHER_CHEBYSHEV; }
    void set_smoother_jacobi(SmartPtr<B200Jacobi> j) { desc.smoother = AB_SMOOTHER_JACOBI; desc.jacobi_damp = j->damp; }
    void set_cycle_type(const char* c) { if (std::strcmp(c, "V") != 0) UG_THROW("ADMMOptimB200: only the V-cycle is served (u3:23)"); }
    void set_num_presmooth(int n) { desc.pre_smooth = n; }
    void set_num_postsmooth(int n) { desc.post_smooth = n; }
    void set_rap(bool b) { desc.rap = b ? 1 : 0; }
    void set_discretization(SmartPtr<B200DomainDiscretization>) {}
    void set_transfer(SmartPtr<B200StdTransfer>) {}
};
class B200LinearSolver {
 public:
    ab_solver* h = nullptr;
    SmartPtr<B200ConvCheck> cc;
    virtual ~B200LinearSolver() { if (h) ab_solver_destroy(h); }
    virtual void create(B200ApproximationSpace& s) = 0;
    void set_convergence_check(SmartPtr<B200ConvCheck> c) { cc = c; }
    bool init(B200AssembledLinearOperator& A, B200GridFunction& x) {                                                  // 3d:979
        if (!h) create(*A.dd->space);
        check(ab_solver_init(h, A.h, x.h));
        return true;
    }
    bool apply(B200GridFunction& x, B200GridFunction& b) { int ok = 0; check(ab_solver_apply(h, x.h, b.h, &ok)); return ok != 0; }                             // 3d:980
    bool apply_return_defect(B200GridFunction& x, B200GridFunction& b) { int ok = 0; check(ab_solver_apply_return_defect(h, x.h, b.h, &ok)); return ok != 0; }  // 3d:1095
    int step() { int n = 0; check(ab_solver_step(h, &n)); return n; }                                                 // 3d:1160
    number defect() { double d = 0; check(ab_solver_last_defect(h, &d)); return d; }
};
class B200BiCGStab : public B200LinearSolver {
 public:
    SmartPtr<B200GeometricMultiGrid> gmg;
    void set_preconditioner(SmartPtr<B200GeometricMultiGrid> g) { gmg = g; }
    void create(B200ApproximationSpace& s) override {
        if (!gmg) UG_THROW("ADMMOptimB200: BiCGStab is served with a GeometricMultiGrid preconditioner (u3:10-31)");
        ab_gmg_desc d = gmg->desc;
        if (cc) { d.max_iterations = cc->max_its; d.abs_tol = cc->abs_tol; d.red_tol = cc->reduction; d.verbose = cc->verbose ? 1 : 0; }
        else { d.max_iterations = 100; d.abs_tol = 1e-12; }
        check(ab_solver_create_bicgstab_gmg(s.h, &d, &h));
    }
};
class B200CG : public B200LinearSolver {
 public:
    SmartPtr<B200Jacobi> jac;
    void set_preconditioner(SmartPtr<B200Jacobi> j) { jac = j; }
    void create(B200ApproximationSpace& s) override {
        B200ConvCheck c = cc ? *cc : B200ConvCheck();
        check(ab_solver_create_cg_jacobi(s.h, jac ? jac->damp : 1.0, c.max_its, c.abs_tol, c.reduction, c.verbose ? 1 : 0, &h));
    }
};

// ---- algebra / plugin free functions   3d_admm.lua:760,976,994; 910,916; 1137,1167-1168; 817,1333 -----------------------------
static void VecScaleAssign(B200GridFunction& dst, number a, B200GridFunction& src) { check(ab_vec_scale_assign(dst.h, a, src.h)); }
static void VecScaleAdd2(B200GridFunction& dst, number a, B200GridFunction& x, number b, B200GridFunction& y) { check(ab_vec_scale_add2(dst.h, a, x.h, b, y.h)); }
static number VecProd(B200GridFunction& x, B200GridFunction& y) { double v = 0; check(ab_vec_prod(x.h, y.h, &v)); return v; }
static number VecNorm(B200GridFunction& x) { double v = 0; check(ab_vec_norm(x.h, &v)); return v; }
static number L2Norm(B200GridFunction& gf, const char* fct, int /*quadOrder*/, const char* /*subsets*/) {
    double v = 0;
    check(ab_l2norm(gf.h, gf.space->fct_index(fct), &v));
    return v;
}
static void Testing(B200GridFunction& q_projected, B200GridFunction& q, const char* /*cmps*/, number sigma) { check(ab_project_frobenius(q_projected.h, q.h, sigma)); }
static void ProjectWithSpectralNorm(B200GridFunction& q_projected, B200GridFunction& q, const char*, number sigma) { check(ab_project_spectral(q_projected.h, q.h, sigma)); }
static number MaximumFrobeniusNorm(B200GridFunction& u, const char*, const char*, int) { double v = 0; check(ab_max_frobenius_norm(u.h, &v)); return v; }
static number MaxSpectralNorm(B200GridFunction& u, const char*, const char*, int) { double v = 0; check(ab_max_spectral_norm(u.h, &v)); return v; }
static number VolumeDefect(B200GridFunction& u, number reference_volume, const char*, const char*, int, bool, int, bool) {
    double v = 0;
    check(ab_volume_defect(u.h, reference_volume, &v));
    return v;
}
static std::vector<number> BarycenterDefect(B200GridFunction& u, const char*, const char*, int) {
    double b[3] = {0, 0, 0};
    check(ab_barycenter_defect(u.h, b));
    return std::vector<number>(b, b + u.space->dom->dim());
}
static void SetZeroAwayFromSubset(B200GridFunction& gf, const char*, const char* subset) { check(ab_set_zero_away_from_subset(gf.h, subset)); }
static void TransformDomainByDisplacement(B200GridFunction& u, const char*) { check(ab_transform_domain_by_displacement(u.space->dom->h, u.h)); }


// ---- output of GPU-resident data ------------------------------------------------------------------------------------------
// SaveGridLevelToFile(dom:grid(), dom:subset_handler(), numRefs, "Mesh_lev..step...ugx")  3d_admm.lua:795 (bDebugOutput) for the
// GPU-side copy of the grid: one level with its CURRENT coordinates in the layout of the shipped grids -- all edges, all
// triangles (3D: the faces of the tetrahedra), the tetrahedra, and a subset handler in which every entity has exactly one subset
// (boundary edges / faces keep the subset handed down by the refinement, interior ones fall into the subset of the volumes).
// Same statement as admm_optim_b200/ugx.py (the Python mirror); LoadDomain reads the file back bit for bit.
void SaveGridLevelToFileB200(ab_domain* dom, int level, const char* filename) {
    int dim = 0, nv = 0, ne = 0, nse = 0, nsf = 0, nsub = 0;
    check(ab_domain_level_info(dom, level, &dim, &nv, &ne, nullptr, nullptr));
    check(ab_domain_special_info(dom, level, &nse, &nsf, &nsub));
    const int nen = dim + 1;
    std::vector<double> xyz((size_t)nv * dim);
    std::vector<int32_t> el((size_t)ne * nen), vsub((size_t)nv), esub((size_t)ne), se((size_t)nse * 2), ses((size_t)nse), sf((size_t)nsf * 3), sfs((size_t)nsf);
    check(ab_domain_get_level(dom, level, xyz.data(), el.data(), vsub.data(), nullptr, nullptr));
    check(ab_domain_get_special(dom, level, nse ? se.data() : nullptr, nse ? ses.data() : nullptr, nsf ? sf.data() : nullptr, nsf ? sfs.data() : nullptr, esub.data()));
    if (ne == 0) UG_THROW("ADMMOptimB200: SaveGridLevelToFile on an empty level");
    const int vol_sub = esub[0];
    typedef std::array<int32_t, 2> E2;
    typedef std::array<int32_t, 3> F3;
    static const int LE[6][2] = {{0, 1}, {1, 2}, {0, 2}, {0, 3}, {1, 3}, {2, 3}};
    static const int LF[4][3] = {{0, 1, 2}, {0, 1, 3}, {1, 2, 3}, {0, 2, 3}};
    std::vector<E2> edges;
    std::vector<F3> faces;
    edges.reserve((size_t)ne * (dim == 3 ? 6 : 3));
    for (int e = 0; e < ne; ++e) {
        const int32_t* v = &el[(size_t)e * nen];
        for (int k = 0; k < (dim == 3 ? 6 : 3); ++k) {
            const int32_t a = v[LE[k][0]], b = v[LE[k][1]];
            edges.push_back({std::min(a, b), std::max(a, b)});
        }
        if (dim == 3)
            for (int k = 0; k < 4; ++k) {
                F3 f = {v[LF[k][0]], v[LF[k][1]], v[LF[k][2]]};
                std::sort(f.begin(), f.end());
                faces.push_back(f);
            }
    }
    std::sort(edges.begin(), edges.end());
    edges.erase(std::unique(edges.begin(), edges.end()), edges.end());
    std::sort(faces.begin(), faces.end());
    faces.erase(std::unique(faces.begin(), faces.end()), faces.end());
    std::vector<int32_t> edge_sub(edges.size(), vol_sub), face_sub(faces.size(), vol_sub);
    for (int i = 0; i < nse; ++i) {
        const E2 k = {std::min(se[2 * i], se[2 * i + 1]), std::max(se[2 * i], se[2 * i + 1])};
        auto it = std::lower_bound(edges.begin(), edges.end(), k);
        if (it == edges.end() || *it != k) UG_THROW("ADMMOptimB200: a boundary edge is not a side of any element");
        edge_sub[it - edges.begin()] = ses[i];
    }
    for (int i = 0; i < nsf; ++i) {
        F3 k = {sf[3 * i], sf[3 * i + 1], sf[3 * i + 2]};
        std::sort(k.begin(), k.end());
        auto it = std::lower_bound(faces.begin(), faces.end(), k);
        if (it == faces.end() || *it != k) UG_THROW("ADMMOptimB200: a boundary face is not a side of any element");
        face_sub[it - faces.begin()] = sfs[i];
    }
    FILE* f = std::fopen(filename, "w");
    if (!f) UG_THROW("ADMMOptimB200: cannot write " << filename);
    std::fprintf(f, "<?xml version=\"1.0\" encoding=\"utf-8\"?>\n<grid name=\"defGrid\">\n\t<vertices coords=\"%d\">", dim);
    for (size_t i = 0; i < xyz.size(); ++i) std::fprintf(f, "%s%.17g", i ? " " : "", xyz[i]);
    std::fprintf(f, "</vertices>\n\t<edges>");
    for (size_t i = 0; i < edges.size(); ++i) std::fprintf(f, "%s%d %d", i ? " " : "", edges[i][0], edges[i][1]);
    std::fprintf(f, "</edges>\n\t<triangles>");
    if (dim == 3) for (size_t i = 0; i < faces.size(); ++i) std::fprintf(f, "%s%d %d %d", i ? " " : "", faces[i][0], faces[i][1], faces[i][2]);
    else for (int e = 0; e < ne; ++e) std::fprintf(f, "%s%d %d %d", e ? " " : "", el[3 * (size_t)e], el[3 * (size_t)e + 1], el[3 * (size_t)e + 2]);
    std::fprintf(f, "</triangles>\n");
    if (dim == 3) {
        std::fprintf(f, "\t<tetrahedrons>");
        for (size_t i = 0; i < el.size(); ++i) std::fprintf(f, "%s%d", i ? " " : "", el[i]);
        std::fprintf(f, "</tetrahedrons>\n");
    }
    std::fprintf(f, "\t<subset_handler name=\"defSH\">\n");
    for (int s = 0; s < nsub; ++s) {
        char name[256];
        check(ab_domain_subset_name(dom, s, name, (int)sizeof(name)));
        std::fprintf(f, "\t\t<subset name=\"%s\" color=\"%.4f %.4f %.4f 1\" state=\"0\">\n", name, 0.15 + 0.7 * ((s * 37) % 10) / 10.0, 0.15 + 0.7 * ((s * 53 + 3) % 10) / 10.0, 0.15 + 0.7 * ((s * 71 + 6) % 10) / 10.0);
        struct Item { const char* tag; const std::vector<int32_t>* sub; };
        const Item items3[4] = {{"vertices", &vsub}, {"edges", &edge_sub}, {"faces", &face_sub}, {"volumes", &esub}};
        const Item items2[3] = {{"vertices", &vsub}, {"edges", &edge_sub}, {"faces", &esub}};
        const Item* items = dim == 3 ? items3 : items2;
        for (int t = 0; t < (dim == 3 ? 4 : 3); ++t) {
            bool any = false;
            const std::vector<int32_t>& sub = *items[t].sub;
            for (size_t i = 0; i < sub.size(); ++i) {
                if (sub[i] != s) continue;
                if (!any) { std::fprintf(f, "\t\t\t<%s>%zu", items[t].tag, i); any = true; }
                else std::fprintf(f, " %zu", i);
            }
            if (any) std::fprintf(f, "</%s>\n", items[t].tag);
        }
        std::fprintf(f, "\t\t</subset>\n");
    }
    std::fprintf(f, "\t</subset_handler>\n</grid>\n");
    std::fclose(f);
}
static void SaveGridLevelToFile(B200Domain& dom, int level, const char* filename) { SaveGridLevelToFileB200(dom.h, level, filename); }

// XML UnstructuredGrid (.vtu, ASCII) of nodal data on a simplex grid: points padded to 3 coordinates, 2- and 3-vectors padded to 3
// components (same layout as admm_optim_b200/vtk.py)
void WriteVTUB200(const char* path, int dim, int nv, const double* xyz, int ne, const int32_t* elems,
                  const std::vector<std::string>& names, const std::vector<std::vector<int>>& comps, int nfct, const double* values) {
    FILE* f = std::fopen(path, "w");
    if (!f) UG_THROW("ADMMOptimB200: cannot write " << path);
    const int nen = dim + 1;
    std::fprintf(f, "<?xml version=\"1.0\"?>\n<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n <UnstructuredGrid>\n");
    std::fprintf(f, "  <Piece NumberOfPoints=\"%d\" NumberOfCells=\"%d\">\n   <Points>\n    <DataArray type=\"Float64\" NumberOfComponents=\"3\" format=\"ascii\">\n", nv, ne);
    for (int v = 0; v < nv; ++v) {
        for (int c = 0; c < 3; ++c) std::fprintf(f, "%s%.17g", c ? " " : "", c < dim ? xyz[(size_t)v * dim + c] : 0.0);
        std::fprintf(f, "\n");
    }
    std::fprintf(f, "    </DataArray>\n   </Points>\n   <Cells>\n    <DataArray type=\"Int32\" Name=\"connectivity\" format=\"ascii\">\n");
    for (int e = 0; e < ne; ++e) {
        for (int k = 0; k < nen; ++k) std::fprintf(f, "%s%d", k ? " " : "", elems[(size_t)e * nen + k]);
        std::fprintf(f, "\n");
    }
    std::fprintf(f, "    </DataArray>\n    <DataArray type=\"Int32\" Name=\"offsets\" format=\"ascii\">\n");
    for (int e = 0; e < ne; ++e) std::fprintf(f, "%s%d", e ? " " : "", nen * (e + 1));
    std::fprintf(f, "\n    </DataArray>\n    <DataArray type=\"UInt8\" Name=\"types\" format=\"ascii\">\n");
    for (int e = 0; e < ne; ++e) std::fprintf(f, "%s%d", e ? " " : "", dim == 3 ? 10 : 5);     // VTK_TETRA / VTK_TRIANGLE
    std::fprintf(f, "\n    </DataArray>\n   </Cells>\n   <PointData>\n");
    for (size_t s = 0; s < names.size(); ++s) {
        const int nc = (int)comps[s].size(), out = (nc == 2 || nc == 3) ? 3 : nc;
        std::fprintf(f, "    <DataArray type=\"Float64\" Name=\"%s\" NumberOfComponents=\"%d\" format=\"ascii\">\n", names[s].c_str(), out);
        for (int v = 0; v < nv; ++v) {
            for (int c = 0; c < out; ++c) std::fprintf(f, "%s%.17g", c ? " " : "", c < nc ? values[(size_t)v * nfct + comps[s][c]] : 0.0);
            std::fprintf(f, "\n");
        }
        std::fprintf(f, "    </DataArray>\n");
    }
    std::fprintf(f, "   </PointData>\n  </Piece>\n </UnstructuredGrid>\n</VTKFile>\n");
    std::fclose(f);
}
// VTKOutput on deformation-space functions: vtkWriter:clear_selection(); vtkWriter:select_nodal("u1,u2,u3","u");
// vtkWriter:print("u", u, step+1, step+1, false)   3d_admm.lua:716, 985-987, 1099-1101, 1400-1406
class B200VTKOutput {
 public:
    std::vector<std::string> names;
    std::vector<std::vector<std::string>> fcts;
    void clear_selection() { names.clear(); fcts.clear(); }
    void select_nodal(const char* functions, const char* name) { fcts.push_back(tokenize(functions)); names.push_back(name); }
    void select_all(bool flag) { if (!flag) clear_selection(); }
    void print(const char* filename, B200GridFunction& gf, int step, number /*time*/, bool make_consistent) {
        B200ApproximationSpace& sp = *gf.space;
        if (sp.kind != AB_SPACE_P1) UG_THROW("ADMMOptimB200: VTKOutput writes nodal (Lagrange-1) functions");
        ab_domain* dom = sp.dom->h;
        int nl = 0, dim = 0, nv = 0, ne = 0;
        check(ab_domain_num_levels(dom, &nl));
        check(ab_domain_level_info(dom, nl - 1, &dim, &nv, &ne, nullptr, nullptr));
        std::vector<double> xyz((size_t)nv * dim), vals(gf.num_dofs());
        std::vector<int32_t> el((size_t)ne * (dim + 1)), vsub((size_t)nv);
        check(ab_domain_get_level(dom, nl - 1, xyz.data(), el.data(), vsub.data(), nullptr, nullptr));
        if (make_consistent) gf.change_storage_type_to_consistent();
        gf.copy_to_host(vals.data());
        std::vector<std::string> nm = names;
        std::vector<std::vector<std::string>> fc = fcts;
        if (nm.empty()) { nm.push_back("u"); fc.push_back(sp.names); }
        std::vector<std::vector<int>> comps;
        for (auto& group : fc) {
            comps.emplace_back();
            for (auto& n : group) comps.back().push_back(sp.fct_index(n));
        }
        char path[1024];
        std::snprintf(path, sizeof(path), "%s_t%04d.vtu", filename, step);
        WriteVTUB200(path, dim, nv, xyz.data(), ne, el.data(), nm, comps, (int)sp.names.size(), vals.data());
    }
};

template <typename T>
static void register_elemdisc(bridge::Registry& reg, const std::string& name, const std::string& grp) {
    typedef B200ElemDisc B;
    reg.add_class_<T, B>(name, grp).template add_constructor<void (*)(const char*, const char*)>("Function(s)#Subset(s)").set_construct_as_smart_pointer(true);
}

static void register_all(bridge::Registry& reg, const std::string& grp) {
    // grid + spaces
    reg.add_class_<B200Domain>("B200Domain", grp).add_constructor()
        .add_method("num_surface_elements", &B200Domain::num_surface_elements).add_method("dim", &B200Domain::dim)
        .set_construct_as_smart_pointer(true);
    reg.add_function("B200LoadDomain", &LoadDomain, grp, "", "Domain#Filename");
    reg.add_function("B200CreateRegularHierarchy", &CreateRegularHierarchy, grp, "", "Domain#NumRefs");
    reg.add_class_<B200ApproximationSpace>("B200ApproximationSpace", grp)
        .template add_constructor<void (*)(SmartPtr<B200Domain>)>("Domain")
        .add_method("add_fct", &B200ApproximationSpace::add_fct, "", "Functions#Type#Order")
        .add_method("add_fct", &B200ApproximationSpace::add_fct_p0, "", "Functions#Type")
        .add_method("init_levels", &B200ApproximationSpace::init_levels)
        .add_method("init_top_surface", &B200ApproximationSpace::init_top_surface)
        .add_method("print_statistic", &B200ApproximationSpace::print_statistic)
        .set_construct_as_smart_pointer(true);
    reg.add_class_<B200GridFunction>("B200GridFunction", grp)
        .template add_constructor<void (*)(SmartPtr<B200ApproximationSpace>)>("ApproximationSpace")
        .add_method("set", &B200GridFunction::set)
        .add_method("has_storage_type_additive", &B200GridFunction::has_storage_type_additive)
        .add_method("has_storage_type_consistent", &B200GridFunction::has_storage_type_consistent)
        .add_method("change_storage_type_to_consistent", &B200GridFunction::change_storage_type_to_consistent)
        .add_method("change_storage_type_to_additive", &B200GridFunction::change_storage_type_to_additive)
        .add_method("assign_from_host", &B200GridFunction::assign_from_host)
        .add_method("copy_to_host", &B200GridFunction::copy_to_host)
        .add_method("num_dofs", &B200GridFunction::num_dofs)
        .set_construct_as_smart_pointer(true);
    reg.add_class_<B200GridFunctionData>("B200GridFunctionData", grp).add_constructor().set_construct_as_smart_pointer(true);
    reg.add_class_<B200GridFunctionNumberData, B200GridFunctionData>("B200GridFunctionNumberData", grp)
        .template add_constructor<void (*)(SmartPtr<B200GridFunction>, const char*)>("GridFunction#Component").add_constructor()
        .set_construct_as_smart_pointer(true);
    reg.add_class_<B200GridFunctionGradientData, B200GridFunctionData>("B200GridFunctionGradientData", grp)
        .template add_constructor<void (*)(SmartPtr<B200GridFunction>, const char*)>("GridFunction#Component").add_constructor()
        .set_construct_as_smart_pointer(true);

    // element discretisations: the base class carries every setter the scripts call (3d:393-694, 2d:388-669)
    {
        typedef B200ElemDisc T;
        auto c = reg.add_class_<T>("B200ElemDisc", grp);
        c.add_method("set_quad_order", &T::set_quad_order).add_method("set_lambda_vol", &T::set_lambda_vol)
            .add_method("set_lambda_barycenter", &T::set_lambda_barycenter).add_method("set_step_length", &T::set_step_length)
            .add_method("set_tau", &T::set_tau).add_method("set_index", &T::set_index)
            .add_method("set_multiplier_vol", &T::set_multiplier_vol).add_method("set_multiplier_bx", &T::set_multiplier_bx)
            .add_method("set_multiplier_by", &T::set_multiplier_by).add_method("set_multiplier_bz", &T::set_multiplier_bz)
            .add_method("set_scaling", &T::set_scaling).add_method("set_high_order_scaling", &T::set_high_order_scaling)
            .add_method("set_second_order", &T::set_second_order).add_method("set_kinematic_viscosity", &T::set_kinematic_viscosity)
            .add_method("set_pressure", &T::set_pressure).add_method("set_adjoint_pressure", &T::set_adjoint_pressure);
#define AB_REG_IMPORT(k)                                                                                                    \
        c.add_method("set_deformation_d" #k, &T::set_deformation_d##k).add_method("set_deformation_vector_d" #k, &T::set_deformation_vector_d##k) \
            .add_method("set_velocity_d" #k, &T::set_velocity_d##k).add_method("set_velocity_vector_d" #k, &T::set_velocity_vector_d##k)           \
            .add_method("set_adjoint_velocity_d" #k, &T::set_adjoint_velocity_d##k)                                                                \
            .add_method("set_adjoint_velocity_vector_d" #k, &T::set_adjoint_velocity_vector_d##k);
        AB_REG_IMPORT(1) AB_REG_IMPORT(2) AB_REG_IMPORT(3)
#undef AB_REG_IMPORT
#define AB_REG_TENSOR(i, j)                                                                                                 \
        c.add_method("set_lambda" #i #j, &T::set_lambda##i##j).add_method("set_q" #i #j, &T::set_q##i##j).add_method("set_qproj" #i #j, &T::set_qproj##i##j);
        AB_REG_TENSOR(0, 0) AB_REG_TENSOR(0, 1) AB_REG_TENSOR(0, 2) AB_REG_TENSOR(1, 0) AB_REG_TENSOR(1, 1) AB_REG_TENSOR(1, 2)
        AB_REG_TENSOR(2, 0) AB_REG_TENSOR(2, 1) AB_REG_TENSOR(2, 2)
#undef AB_REG_TENSOR
    }
    register_elemdisc<DeformationEquation>(reg, "DeformationEquation", grp);
    register_elemdisc<DeformationEquationRHS>(reg, "DeformationEquationRHS", grp);
    register_elemdisc<DeformationEquationLargeProblemRHS>(reg, "DeformationEquationLargeProblemRHS", grp);
    register_elemdisc<VolumeConstraintSecondDerivative>(reg, "VolumeConstraintSecondDerivative", grp);
    register_elemdisc<SecondDerivativeVolume>(reg, "SecondDerivativeVolume", grp);
    register_elemdisc<SecondDerivativeBarycenter>(reg, "SecondDerivativeBarycenter", grp);
    register_elemdisc<XBarycenterConstraintSecondDerivative>(reg, "XBarycenterConstraintSecondDerivative", grp);
    register_elemdisc<MassModel>(reg, "MassModel", grp);
    register_elemdisc<LambdaUpdate>(reg, "LambdaUpdate", grp);

    // discretisation
    reg.add_class_<B200DirichletBoundary>("B200DirichletBoundary", grp).add_constructor()
        .add_method("add", &B200DirichletBoundary::add, "", "Value#Function#Subsets").set_construct_as_smart_pointer(true);
    reg.add_class_<B200DomainDiscretization>("B200DomainDiscretization", grp)
        .template add_constructor<void (*)(SmartPtr<B200ApproximationSpace>)>("ApproximationSpace")
        .add_method("add", &B200DomainDiscretization::add).add_method("add", &B200DomainDiscretization::add_dirichlet)
        .add_method("assemble_jacobian", &B200DomainDiscretization::assemble_jacobian)
        .add_method("assemble_defect", &B200DomainDiscretization::assemble_defect)
        .add_method("adjust_solution", &B200DomainDiscretization::adjust_solution)
        .set_construct_as_smart_pointer(true);
    reg.add_class_<B200AssembledLinearOperator>("B200AssembledLinearOperator", grp)
        .template add_constructor<void (*)(SmartPtr<B200DomainDiscretization>)>("DomainDiscretization")
        .add_method("apply", &B200AssembledLinearOperator::apply).set_construct_as_smart_pointer(true);

    // solvers
    reg.add_class_<B200ConvCheck>("B200ConvCheck", grp).add_constructor()
        .template add_constructor<void (*)(int, number, number, bool)>("MaxIts#AbsTol#Reduction#Verbose").set_construct_as_smart_pointer(true);
    reg.add_class_<B200Jacobi>("B200Jacobi", grp).add_constructor().template add_constructor<void (*)(number)>("Damping").set_construct_as_smart_pointer(true);
    reg.add_class_<B200GaussSeidel>("B200GaussSeidel", grp).add_constructor().add_method("set_damp", &B200GaussSeidel::set_damp).set_construct_as_smart_pointer(true);
    reg.add_class_<B200SuperLU>("B200SuperLU", grp).add_constructor().set_construct_as_smart_pointer(true);
    reg.add_class_<B200StdTransfer>("B200StdTransfer", grp).add_constructor().set_construct_as_smart_pointer(true);
    reg.add_class_<B200GeometricMultiGrid>("B200GeometricMultiGrid", grp)
        .template add_constructor<void (*)(SmartPtr<B200ApproximationSpace>)>("ApproximationSpace")
        .add_method("set_base_level", &B200GeometricMultiGrid::set_base_level)
        .add_method("set_base_solver", &B200GeometricMultiGrid::set_base_solver)
        .add_method("set_gathered_base_solver_if_ambiguous", &B200GeometricMultiGrid::set_gathered_base_solver_if_ambiguous)
        .add_method("set_smoother", &B200GeometricMultiGrid::set_smoother).add_method("set_smoother", &B200GeometricMultiGrid::set_smoother_jacobi)
        .add_method("set_cycle_type", &B200GeometricMultiGrid::set_cycle_type)
        .add_method("set_num_presmooth", &B200GeometricMultiGrid::set_num_presmooth)
        .add_method("set_num_postsmooth", &B200GeometricMultiGrid::set_num_postsmooth)
        .add_method("set_rap", &B200GeometricMultiGrid::set_rap)
        .add_method("set_discretization", &B200GeometricMultiGrid::set_discretization)
        .add_method("set_transfer", &B200GeometricMultiGrid::set_transfer)
        .set_construct_as_smart_pointer(true);
    {
        typedef B200LinearSolver T;
        reg.add_class_<T>("B200LinearSolver", grp)
            .add_method("set_convergence_check", &T::set_convergence_check).add_method("init", &T::init).add_method("apply", &T::apply)
            .add_method("apply_return_defect", &T::apply_return_defect).add_method("step", &T::step).add_method("defect", &T::defect);
    }
    reg.add_class_<B200BiCGStab, B200LinearSolver>("B200BiCGStab", grp).add_constructor()
        .add_method("set_preconditioner", &B200BiCGStab::set_preconditioner).set_construct_as_smart_pointer(true);
    reg.add_class_<B200CG, B200LinearSolver>("B200CG", grp).add_constructor()
        .add_method("set_preconditioner", &B200CG::set_preconditioner).set_construct_as_smart_pointer(true);

    // algebra (ugcore names get the prefix; the prelude dispatches on the argument type)
    reg.add_function("B200VecScaleAssign", &VecScaleAssign, grp);
    reg.add_function("B200VecScaleAdd2", &VecScaleAdd2, grp);
    reg.add_function("B200VecProd", &VecProd, grp);
    reg.add_function("B200VecNorm", &VecNorm, grp);
    reg.add_function("B200L2Norm", &L2Norm, grp);
    // functions of the replaced plugins keep their names (3d:910,916,1167,1168,817,1333; 2d:901-902)
    reg.add_function("Testing", &Testing, grp);
    reg.add_function("ProjectWithSpectralNorm", &ProjectWithSpectralNorm, grp);
    reg.add_function("MaximumFrobeniusNorm", &MaximumFrobeniusNorm, grp);
    reg.add_function("MaxSpectralNorm", &MaxSpectralNorm, grp);
    reg.add_function("VolumeDefect", &VolumeDefect, grp);
    reg.add_function("BarycenterDefect", &BarycenterDefect, grp);
    reg.add_function("SetZeroAwayFromSubset", &SetZeroAwayFromSubset, grp);
    reg.add_function("TransformDomainByDisplacement", &TransformDomainByDisplacement, grp);
    // output of GPU-resident data
    reg.add_function("B200SaveGridLevelToFile", &SaveGridLevelToFile, grp, "", "Domain#Level#Filename");
    reg.add_class_<B200VTKOutput>("B200VTKOutput", grp).add_constructor()
        .add_method("clear_selection", &B200VTKOutput::clear_selection).add_method("select_nodal", &B200VTKOutput::select_nodal)
        .add_method("select_all", &B200VTKOutput::select_all).add_method("print", &B200VTKOutput::print)
        .set_construct_as_smart_pointer(true);
}

}  // namespace ADMMOptimB200
}  // namespace ug

extern "C" void InitUGPlugin_ADMMOptimB200(ug::bridge::Registry* reg, std::string grp) {
    grp.append("/ADMMOptimB200");
    try {
        ug::ADMMOptimB200::register_all(*reg, grp);
    }
    UG_REGISTRY_CATCH_THROW(grp);
}
