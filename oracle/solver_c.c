/* ORACLE (test infrastructure, not product code) -- the solve path of the CPU restatement in C / OpenMP.
 *
 * What the reference runs per `solver:init` + `solver:apply` (obstacle_optim_3d_util.lua:9-43, called from
 * 3d_admm.lua:979-980, 1009-1011, 1094-1095): Galerkin coarse operators (rap = true, u3:27), a direct base solve on level 0
 * (SuperLU, u3:21), a V(3,3) cycle with Gauss-Seidel smoothing (u3:16,25-26) and P1 standard transfers (u3:28) as the
 * preconditioner of BiCGStab with the ConvCheck of u3:32-38.  oracle/fem_np.py states the same algorithm in NumPy/SciPy (the
 * checker of the tests); this file is the multi-threaded CPU baseline bench.py times on all host cores: Gauss-Seidel inside
 * each thread's block of rows, Jacobi coupling between the blocks -- what UG4 does across MPI ranks under
 * `mpirun -np T` (3d_admm.lua:25).  tests/test_oracle.py pins it against the NumPy statement.
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this library.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void oracle_gs_forward(int n, const int* ip, const int* idx, const double* a, const double* b, double* x, int nblocks, double* xold);

typedef struct {
    int n;                 /* unknowns of the level */
    int *ip, *idx;         /* operator, CSR (owned for coarse levels, borrowed for the top level) */
    double* a;
    int owned;
    int nc;                /* unknowns of the next coarser level */
    const int *pip, *pidx; /* prolongation n x nc, CSR (borrowed) */
    const double* pa;
    const int *rip, *ridx; /* restriction nc x n = P^T, CSR (borrowed) */
    const double* ra;
    const unsigned char* dir; /* Dirichlet mask of THIS level (borrowed) */
    double *dinv, lmax;
    double *x, *b, *r, *d, *t; /* work vectors */
} Level;

typedef struct {
    int nl, threads, smoother, nu1, nu2; /* smoother: 0 gs, 1 chebyshev, 2 jacobi */
    double cheb_ratio, omega;
    Level* L;
    double* lu; /* dense LU of level 0, row-major, partial pivoting */
    int* piv;
    double *kr, *krh, *kp, *kv, *ks, *kt, *kph, *ksh; /* BiCGStab vectors */
} Gmg;

void* oracle_gmg_create(int nl, int threads, int smoother, int nu1, int nu2, double cheb_ratio, double omega) {
    Gmg* G = (Gmg*)calloc(1, sizeof(Gmg));
    G->nl = nl; G->threads = threads > 0 ? threads : 1; G->smoother = smoother; G->nu1 = nu1; G->nu2 = nu2;
    G->cheb_ratio = cheb_ratio; G->omega = omega;
    G->L = (Level*)calloc((size_t)nl, sizeof(Level));
    return G;
}

/* static part of a level: sizes, transfers to the next coarser level, Dirichlet mask */
void oracle_gmg_set_level(void* h, int l, int n, int nc, const int* pip, const int* pidx, const double* pa, const int* rip, const int* ridx,
                          const double* ra, const unsigned char* dir) {
    Gmg* G = (Gmg*)h;
    Level* L = &G->L[l];
    L->n = n; L->nc = nc; L->pip = pip; L->pidx = pidx; L->pa = pa; L->rip = rip; L->ridx = ridx; L->ra = ra; L->dir = dir;
    if (!L->x) {
        L->x = (double*)malloc(sizeof(double) * n); L->b = (double*)malloc(sizeof(double) * n); L->r = (double*)malloc(sizeof(double) * n);
        L->d = (double*)malloc(sizeof(double) * n); L->t = (double*)malloc(sizeof(double) * n); L->dinv = (double*)malloc(sizeof(double) * n);
    }
}

static void spmv(const Gmg* G, int n, const int* ip, const int* idx, const double* a, const double* x, double* y) {
#pragma omp parallel for schedule(static) num_threads(G->threads)
    for (int i = 0; i < n; ++i) {
        double s = 0.0;
        for (int k = ip[i]; k < ip[i + 1]; ++k) s += a[k] * x[idx[k]];
        y[i] = s;
    }
}
static void residual(const Gmg* G, const Level* L, const double* x, const double* b, double* r) {
#pragma omp parallel for schedule(static) num_threads(G->threads)
    for (int i = 0; i < L->n; ++i) {
        double s = b[i];
        for (int k = L->ip[i]; k < L->ip[i + 1]; ++k) s -= L->a[k] * x[L->idx[k]];
        r[i] = s;
    }
}

/* C = R A P (Gustavson, row-parallel, two passes), then symmetric Dirichlet elimination with unit diagonal */
static void rap(const Gmg* G, const Level* F, Level* C) {
    const int nc = F->nc, T = G->threads;
    int* cnt = (int*)calloc((size_t)nc + 1, sizeof(int));
    int** marks = (int**)malloc(sizeof(int*) * T);
    double** accs = (double**)malloc(sizeof(double*) * T);
    for (int t = 0; t < T; ++t) { marks[t] = (int*)malloc(sizeof(int) * nc); accs[t] = (double*)malloc(sizeof(double) * nc); }
    for (int pass = 0; pass < 2; ++pass) {
#pragma omp parallel num_threads(T)
        {
#ifdef _OPENMP
            const int t = omp_get_thread_num();
#else
            const int t = 0;
#endif
            int* mark = marks[t];
            double* acc = accs[t];
            for (int j = 0; j < nc; ++j) mark[j] = -1;
            int* cols = (int*)malloc(sizeof(int) * nc);
#pragma omp for schedule(dynamic, 64)
            for (int I = 0; I < nc; ++I) {
                int len = 0;
                for (int kr = F->rip[I]; kr < F->rip[I + 1]; ++kr) {
                    const int i = F->ridx[kr];
                    const double rv = F->ra[kr];
                    for (int ka = F->ip[i]; ka < F->ip[i + 1]; ++ka) {
                        const int k = F->idx[ka];
                        const double av = rv * F->a[ka];
                        for (int kp = F->pip[k]; kp < F->pip[k + 1]; ++kp) {
                            const int J = F->pidx[kp];
                            if (mark[J] != I) { mark[J] = I; acc[J] = 0.0; cols[len++] = J; }
                            acc[J] += av * F->pa[kp];
                        }
                    }
                }
                if (pass == 0) { cnt[I + 1] = len; continue; }
                /* sort the row's columns (insertion sort: rows are short) and write */
                for (int p = 1; p < len; ++p) { int c = cols[p], q = p - 1; while (q >= 0 && cols[q] > c) { cols[q + 1] = cols[q]; --q; } cols[q + 1] = c; }
                int* oi = C->idx + C->ip[I];
                double* oa = C->a + C->ip[I];
                const int dI = C->dir ? C->dir[I] : 0;
                for (int p = 0; p < len; ++p) {
                    const int J = cols[p];
                    double v = acc[J];
                    if (dI || (C->dir && C->dir[J])) v = (J == I) ? 1.0 : 0.0;
                    oi[p] = J; oa[p] = v;
                }
            }
            free(cols);
        }
        if (pass == 0) {
            for (int I = 0; I < nc; ++I) cnt[I + 1] += cnt[I];
            if (C->owned) { free(C->ip); free(C->idx); free(C->a); }
            C->ip = cnt; C->owned = 1;
            C->idx = (int*)malloc(sizeof(int) * (size_t)cnt[nc]);
            C->a = (double*)malloc(sizeof(double) * (size_t)cnt[nc]);
        }
    }
    for (int t = 0; t < T; ++t) { free(marks[t]); free(accs[t]); }
    free(marks); free(accs);
}

/* new top-level operator: Galerkin chain, smoother data, dense LU of level 0.  Returns 0, or 1 when level 0 is singular. */
int oracle_gmg_setup(void* h, const int* ip, const int* idx, const double* a) {
    Gmg* G = (Gmg*)h;
    Level* top = &G->L[G->nl - 1];
    if (top->owned) { free(top->ip); free(top->idx); free(top->a); top->owned = 0; }
    top->ip = (int*)ip; top->idx = (int*)idx; top->a = (double*)a;
    for (int l = G->nl - 1; l > 0; --l) rap(G, &G->L[l], &G->L[l - 1]);
    for (int l = 0; l < G->nl; ++l) {
        Level* L = &G->L[l];
        double lmax = 0.0;
#pragma omp parallel for schedule(static) num_threads(G->threads) reduction(max : lmax)
        for (int i = 0; i < L->n; ++i) {
            double s = 0.0, d = 1.0;
            for (int k = L->ip[i]; k < L->ip[i + 1]; ++k) { s += fabs(L->a[k]); if (L->idx[k] == i) d = L->a[k]; }
            L->dinv[i] = 1.0 / d;
            if (s / d > lmax) lmax = s / d;
        }
        L->lmax = lmax;
    }
    /* dense LU with partial pivoting of the level-0 operator (the SuperLU() base solver, u3:21) */
    Level* L0 = &G->L[0];
    const int n = L0->n;
    if (!G->lu) { G->lu = (double*)malloc(sizeof(double) * (size_t)n * n); G->piv = (int*)malloc(sizeof(int) * n); }
    memset(G->lu, 0, sizeof(double) * (size_t)n * n);
    for (int i = 0; i < n; ++i)
        for (int k = L0->ip[i]; k < L0->ip[i + 1]; ++k) G->lu[(size_t)i * n + L0->idx[k]] = L0->a[k];
    for (int k = 0; k < n; ++k) {
        int p = k;
        double best = fabs(G->lu[(size_t)k * n + k]);
        for (int i = k + 1; i < n; ++i) { const double v = fabs(G->lu[(size_t)i * n + k]); if (v > best) { best = v; p = i; } }
        if (!(best > 0.0)) return 1;
        G->piv[k] = p;
        if (p != k) for (int c = 0; c < n; ++c) { const double t = G->lu[(size_t)k * n + c]; G->lu[(size_t)k * n + c] = G->lu[(size_t)p * n + c]; G->lu[(size_t)p * n + c] = t; }
        const double inv = 1.0 / G->lu[(size_t)k * n + k];
#pragma omp parallel for schedule(static) num_threads(G->threads)
        for (int i = k + 1; i < n; ++i) {
            double* row = G->lu + (size_t)i * n;
            const double f = row[k] * inv;
            if (f == 0.0) continue;
            row[k] = f;
            const double* pr = G->lu + (size_t)k * n;
            for (int c = k + 1; c < n; ++c) row[c] -= f * pr[c];
        }
    }
    if (!G->kr) {
        const int nt = top->n;
        double** v[] = {&G->kr, &G->krh, &G->kp, &G->kv, &G->ks, &G->kt, &G->kph, &G->ksh};
        for (int q = 0; q < 8; ++q) *v[q] = (double*)malloc(sizeof(double) * nt);
    }
    return 0;
}

static void coarse_solve(const Gmg* G, const double* b, double* x) {
    const int n = G->L[0].n;
    memcpy(x, b, sizeof(double) * n);
    for (int k = 0; k < n; ++k) { const int p = G->piv[k]; if (p != k) { const double t = x[k]; x[k] = x[p]; x[p] = t; } }
    for (int i = 1; i < n; ++i) { const double* row = G->lu + (size_t)i * n; double s = x[i]; for (int c = 0; c < i; ++c) s -= row[c] * x[c]; x[i] = s; }
    for (int i = n - 1; i >= 0; --i) { const double* row = G->lu + (size_t)i * n; double s = x[i]; for (int c = i + 1; c < n; ++c) s -= row[c] * x[c]; x[i] = s / row[i]; }
}

static void smooth(const Gmg* G, Level* L, double* x, const double* b, int nu, int zero_guess) {
    const int n = L->n, T = G->threads;
    if (G->smoother == 0) {      /* Gauss-Seidel inside the thread blocks, Jacobi between them */
        if (zero_guess) memset(x, 0, sizeof(double) * n);
        for (int s = 0; s < nu; ++s) oracle_gs_forward(n, L->ip, L->idx, L->a, b, x, T, L->t);
        return;
    }
    if (G->smoother == 2) {      /* damped point Jacobi */
        for (int s = 0; s < nu; ++s) {
            if (zero_guess && s == 0) {
#pragma omp parallel for schedule(static) num_threads(T)
                for (int i = 0; i < n; ++i) x[i] = G->omega * L->dinv[i] * b[i];
            } else {
                residual(G, L, x, b, L->t);
#pragma omp parallel for schedule(static) num_threads(T)
                for (int i = 0; i < n; ++i) x[i] += G->omega * L->dinv[i] * L->t[i];
            }
        }
        return;
    }
    /* Chebyshev polynomial in D^-1 A on [lmax / ratio, lmax] (the smoother of the CUDA path) */
    const double lmax = L->lmax, lmin = lmax / G->cheb_ratio;
    const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma1 = theta / delta;
    double rho = 1.0 / sigma1;
    double* d = L->d;
    if (zero_guess) {
#pragma omp parallel for schedule(static) num_threads(T)
        for (int i = 0; i < n; ++i) { d[i] = L->dinv[i] * b[i] / theta; x[i] = d[i]; }
    } else {
        residual(G, L, x, b, L->t);
#pragma omp parallel for schedule(static) num_threads(T)
        for (int i = 0; i < n; ++i) { d[i] = L->dinv[i] * L->t[i] / theta; x[i] += d[i]; }
    }
    for (int s = 1; s < nu; ++s) {
        const double rho_new = 1.0 / (2.0 * sigma1 - rho);
        const double c1 = rho_new * rho, c2 = 2.0 * rho_new / delta;
        residual(G, L, x, b, L->t);
#pragma omp parallel for schedule(static) num_threads(T)
        for (int i = 0; i < n; ++i) { d[i] = c1 * d[i] + c2 * (L->dinv[i] * L->t[i]); x[i] += d[i]; }
        rho = rho_new;
    }
}

static void vcycle(const Gmg* G, int l, const double* b, double* x) {
    if (l == 0) { coarse_solve(G, b, x); return; }
    Level* L = &G->L[l];
    Level* C = &G->L[l - 1];
    smooth(G, L, x, b, G->nu1, 1);
    residual(G, L, x, b, L->r);
    spmv(G, L->nc, L->rip, L->ridx, L->ra, L->r, C->b);
    if (C->dir) for (int i = 0; i < C->n; ++i) if (C->dir[i]) C->b[i] = 0.0;
    vcycle(G, l - 1, C->b, C->x);
    spmv(G, L->n, L->pip, L->pidx, L->pa, C->x, L->r);
#pragma omp parallel for schedule(static) num_threads(G->threads)
    for (int i = 0; i < L->n; ++i) x[i] += L->r[i];
    smooth(G, L, x, b, G->nu2, 0);
}

void oracle_gmg_vcycle(void* h, const double* b, double* x) {
    Gmg* G = (Gmg*)h;
    vcycle(G, G->nl - 1, b, x);
}

static double dot(const Gmg* G, int n, const double* x, const double* y) {
    double s = 0.0;
#pragma omp parallel for schedule(static) num_threads(G->threads) reduction(+ : s)
    for (int i = 0; i < n; ++i) s += x[i] * y[i];
    return s;
}

/* right-preconditioned BiCGStab with the ConvCheck of u3:32-38 (fem_np.bicgstab states the same loop).  Returns the iteration
 * count; *ok = converged; r_out (may be NULL) receives the final residual. */
int oracle_bicgstab_gmg(void* h, const double* b, double* x, double abs_tol, double red_tol, int max_it, int* ok, double* r_out) {
    Gmg* G = (Gmg*)h;
    const Level* A = &G->L[G->nl - 1];
    const int n = A->n, T = G->threads;
    double *r = G->kr, *rh = G->krh, *p = G->kp, *v = G->kv, *s = G->ks, *t = G->kt, *ph = G->kph, *sh = G->ksh;
    residual(G, A, x, b, r);
    double nr = sqrt(dot(G, n, r, r));
    const double nr0 = nr;
    int it = 0;
    *ok = 0;
    if (nr < abs_tol) { *ok = 1; if (r_out) memcpy(r_out, r, sizeof(double) * n); return 0; }
    memcpy(rh, r, sizeof(double) * n);
    memset(p, 0, sizeof(double) * n);
    memset(v, 0, sizeof(double) * n);
    double rho_old = 1.0, alpha = 1.0, omega = 1.0;
    for (it = 1; it <= max_it; ++it) {
        const double rho = dot(G, n, rh, r);
        if (rho == 0.0 || !isfinite(rho)) break;
        const double beta = (rho / rho_old) * (alpha / omega);
#pragma omp parallel for schedule(static) num_threads(T)
        for (int i = 0; i < n; ++i) p[i] = r[i] + beta * (p[i] - omega * v[i]);
        vcycle(G, G->nl - 1, p, ph);
        spmv(G, n, A->ip, A->idx, A->a, ph, v);
        alpha = rho / dot(G, n, rh, v);
#pragma omp parallel for schedule(static) num_threads(T)
        for (int i = 0; i < n; ++i) s[i] = r[i] - alpha * v[i];
        vcycle(G, G->nl - 1, s, sh);
        spmv(G, n, A->ip, A->idx, A->a, sh, t);
        const double tt = dot(G, n, t, t);
        omega = tt > 0.0 ? dot(G, n, t, s) / tt : 0.0;
#pragma omp parallel for schedule(static) num_threads(T)
        for (int i = 0; i < n; ++i) { x[i] += alpha * ph[i] + omega * sh[i]; r[i] = s[i] - omega * t[i]; }
        rho_old = rho;
        nr = sqrt(dot(G, n, r, r));
        if (nr < abs_tol || nr < red_tol * nr0) { *ok = 1; break; }
        if (omega == 0.0 || !isfinite(nr)) break;
    }
    if (it > max_it) it = max_it;
    if (r_out) memcpy(r_out, r, sizeof(double) * n);
    return it;
}

void oracle_gmg_destroy(void* h) {
    Gmg* G = (Gmg*)h;
    if (!G) return;
    for (int l = 0; l < G->nl; ++l) {
        Level* L = &G->L[l];
        if (L->owned) { free(L->ip); free(L->idx); free(L->a); }
        free(L->x); free(L->b); free(L->r); free(L->d); free(L->t); free(L->dinv);
    }
    double* v[] = {G->kr, G->krh, G->kp, G->kv, G->ks, G->kt, G->kph, G->ksh};
    for (int q = 0; q < 8; ++q) free(v[q]);
    free(G->lu); free(G->piv); free(G->L); free(G);
}
