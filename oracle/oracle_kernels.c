/* ORACLE (test infrastructure, not product code) -- plain C kernels of the CPU restatement.
 *
 * The two loops NumPy cannot express efficiently: the forward Gauss-Seidel sweep of the reference's smoother
 * (smoother = "gs", obstacle_optim_3d_util.lua:16; V(3,3), :25-26) and the CSR matrix-vector product inside
 * BiCGStab / the V-cycle.  nblocks = 1 is the sequential lexicographic sweep of a serial `ugshell` run;
 * nblocks = T reproduces what UG4 does under `mpirun -np T` (3d_admm.lua:25): Gauss-Seidel inside each process'
 * block of rows, Jacobi coupling between blocks [UPSTREAM-UNVERIFIED, SURVEY.md App. C5] -- that variant is the
 * multi-threaded CPU baseline of bench.py.
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this library.
 */
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* x <- x + (D+L)^-1 (b - A x) restricted to row blocks: one forward sweep */
void oracle_gs_forward(int n, const int* ip, const int* idx, const double* a, const double* b, double* x, int nblocks, double* xold) {
    if (nblocks <= 1) {
        for (int i = 0; i < n; ++i) {
            double s = b[i], d = 1.0;
            for (int k = ip[i]; k < ip[i + 1]; ++k) {
                const int j = idx[k];
                if (j == i) d = a[k]; else s -= a[k] * x[j];
            }
            x[i] = s / d;
        }
        return;
    }
    memcpy(xold, x, (size_t)n * sizeof(double));
#pragma omp parallel for schedule(static, 1) num_threads(nblocks)
    for (int blk = 0; blk < nblocks; ++blk) {
        const int lo = (int)((long long)n * blk / nblocks), hi = (int)((long long)n * (blk + 1) / nblocks);
        for (int i = lo; i < hi; ++i) {
            double s = b[i], d = 1.0;
            for (int k = ip[i]; k < ip[i + 1]; ++k) {
                const int j = idx[k];
                /* both loads are issued, the select is branch-free (an unpredictable branch per entry costs more than the sweep) */
                const double xo = xold[j], xn = x[j];
                const double xv = ((unsigned)(j - lo) < (unsigned)(hi - lo)) ? xn : xo;
                const double av = a[k];
                d = (j == i) ? av : d;
                s -= (j == i) ? 0.0 : av * xv;
            }
            x[i] = s / d;
        }
    }
}

void oracle_spmv(int n, const int* ip, const int* idx, const double* a, const double* x, double* y, int nthreads) {
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (int i = 0; i < n; ++i) {
        double s = 0.0;
        for (int k = ip[i]; k < ip[i + 1]; ++k) s += a[k] * x[idx[k]];
        y[i] = s;
    }
}

/* threads of the element loops below (bench.py's CPU legs set it; default 1 = the deterministic serial scatter the tests pin) */
static int g_elem_threads = 1;
void oracle_set_elem_threads(int t) { g_elem_threads = t > 0 ? t : 1; }

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------------------------------
 * P1 element kernels in C (bench.py's CPU legs only: `ug4_np.Backend(fast_assembly=True)`).  Same formulas as
 * fem_np.hessian_matrix / fem_np.load_vector (the NumPy versions stay the checker of the tests; tests/test_oracle.py
 * compares the two entry by entry).  A compiled CPU reference would not spend two thirds of an ADMM iteration in NumPy
 * temporaries -- this keeps the CPU baseline honest.
 * ------------------------------------------------------------------------------------------------------------------ */
#include <math.h>

/* G (d+1 x d): P1 gradients, vol = |det J| / d!, for one simplex with corner coordinates X ((d+1) x d) */
static void oracle_elem_geom(int d, const double* X, double* G, double* vol) {
    if (d == 2) {
        const double a = X[2] - X[0], c = X[3] - X[1];      /* J = [[a, b], [c, e]]: columns = edge vectors */
        const double b = X[4] - X[0], e = X[5] - X[1];
        const double det = a * e - b * c, inv = 1.0 / det;
        /* rows of J^-1 */
        G[2] = e * inv;  G[3] = -b * inv;
        G[4] = -c * inv; G[5] = a * inv;
        G[0] = -(G[2] + G[4]); G[1] = -(G[3] + G[5]);
        *vol = fabs(det) / 2.0;
    } else {
        double J[9];                                        /* J[r][c] = X[c+1][r] - X[0][r] */
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) J[r * 3 + c] = X[(c + 1) * 3 + r] - X[r];
        const double c00 = J[4] * J[8] - J[5] * J[7], c01 = J[5] * J[6] - J[3] * J[8], c02 = J[3] * J[7] - J[4] * J[6];
        const double det = J[0] * c00 + J[1] * c01 + J[2] * c02, inv = 1.0 / det;
        double Ji[9];                                       /* inverse = adj / det */
        Ji[0] = c00 * inv; Ji[1] = (J[2] * J[7] - J[1] * J[8]) * inv; Ji[2] = (J[1] * J[5] - J[2] * J[4]) * inv;
        Ji[3] = c01 * inv; Ji[4] = (J[0] * J[8] - J[2] * J[6]) * inv; Ji[5] = (J[2] * J[3] - J[0] * J[5]) * inv;
        Ji[6] = c02 * inv; Ji[7] = (J[1] * J[6] - J[0] * J[7]) * inv; Ji[8] = (J[0] * J[4] - J[1] * J[3]) * inv;
        for (int a = 0; a < 3; ++a)
            for (int x = 0; x < 3; ++x) G[(a + 1) * 3 + x] = Ji[a * 3 + x];
        for (int x = 0; x < 3; ++x) G[x] = -(Ji[x] + Ji[3 + x] + Ji[6 + x]);
        *vol = fabs(det) / 6.0;
    }
}
static void oracle_cross(const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
/* F = I + sum_a u_a (x) G_a ; C = cof F ; returns det F */
static double oracle_F_cof(int d, const double* U /* (d+1) x d or NULL */, const double* G, double* F, double* C) {
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) {
            double s = (i == j) ? 1.0 : 0.0;
            if (U) for (int a = 0; a <= d; ++a) s += U[a * d + i] * G[a * d + j];
            F[i * d + j] = s;
        }
    if (d == 2) {
        C[0] = F[3]; C[1] = -F[2]; C[2] = -F[1]; C[3] = F[0];
        return F[0] * F[3] - F[1] * F[2];
    }
    for (int i = 0; i < 3; ++i) oracle_cross(F + ((i + 1) % 3) * 3, F + ((i + 2) % 3) * 3, C + i * 3);
    double t[3];
    oracle_cross(F + 3, F + 6, t);
    return F[0] * t[0] + F[1] * t[1] + F[2] * t[2];
}

/* DeformationEquation jacobian: data[slot[e][a][i][b][j]] += K_e[(a,i),(b,j)]   (fem_np.hessian_matrix) */
void oracle_hessian_scatter(int d, long ne, const int* elems, const double* xyz, const double* u, double c, double lam_vol,
                            const double* lam_b, const long long* slot, double* data) {
    const int nd = (d + 1) * d;
    const int has_lam = lam_vol != 0.0 || lam_b[0] != 0.0 || lam_b[1] != 0.0 || (d == 3 && lam_b[2] != 0.0);
#pragma omp parallel for schedule(static) num_threads(g_elem_threads)
    for (long e = 0; e < ne; ++e) {
        double X[12], U[12], G[12], F[9], C[9], K[144], vol;
        for (int a = 0; a <= d; ++a) {
            const long v = elems[e * (d + 1) + a];
            for (int x = 0; x < d; ++x) { X[a * d + x] = xyz[v * d + x]; U[a * d + x] = u ? u[v * d + x] : 0.0; }
        }
        oracle_elem_geom(d, X, G, &vol);
        for (int k = 0; k < nd * nd; ++k) K[k] = 0.0;
        for (int a = 0; a <= d; ++a)
            for (int b = 0; b <= d; ++b) {
                double gg = 0.0;
                for (int x = 0; x < d; ++x) gg += G[a * d + x] * G[b * d + x];
                for (int i = 0; i < d; ++i) K[(a * d + i) * nd + b * d + i] += c * gg;
            }
        if (has_lam) {
            oracle_F_cof(d, u ? U : (const double*)0, G, F, C);
            double w = lam_vol;
            for (int k = 0; k < d; ++k) {
                double xb = 0.0;
                for (int a = 0; a <= d; ++a) xb += X[a * d + k] + U[a * d + k];
                w += lam_b[k] * xb / (d + 1);
            }
            double CG[12];                                  /* (C G_a)_i */
            for (int a = 0; a <= d; ++a)
                for (int i = 0; i < d; ++i) {
                    double s = 0.0;
                    for (int j = 0; j < d; ++j) s += C[i * d + j] * G[a * d + j];
                    CG[a * d + i] = s;
                }
            for (int a = 0; a <= d; ++a)
                for (int b = 0; b <= d; ++b) {
                    if (d == 2) {
                        const double D = G[a * 2] * G[b * 2 + 1] - G[a * 2 + 1] * G[b * 2];
                        K[(a * 2 + 0) * nd + b * 2 + 1] += w * D;
                        K[(a * 2 + 1) * nd + b * 2 + 0] -= w * D;
                    } else {
                        for (int i = 0; i < 3; ++i) {
                            const int j = (i + 1) % 3, k = (i + 2) % 3;
                            double t[3];
                            oracle_cross(G + b * 3, F + k * 3, t);            /* G_b x F_k */
                            K[(a * 3 + i) * nd + b * 3 + j] += w * (G[a * 3] * t[0] + G[a * 3 + 1] * t[1] + G[a * 3 + 2] * t[2]);
                            oracle_cross(G + b * 3, F + j * 3, t);            /* G_b x F_j */
                            K[(a * 3 + i) * nd + b * 3 + k] -= w * (G[a * 3] * t[0] + G[a * 3 + 1] * t[1] + G[a * 3 + 2] * t[2]);
                        }
                    }
                    for (int k = 0; k < d; ++k) {
                        if (lam_b[k] == 0.0) continue;
                        const double f = lam_b[k] / (d + 1);
                        for (int j = 0; j < d; ++j) K[(a * d + k) * nd + b * d + j] += f * CG[b * d + j];
                        for (int i = 0; i < d; ++i) K[(a * d + i) * nd + b * d + k] += f * CG[a * d + i];
                    }
                }
        }
        const long long* sl = slot + (long long)e * nd * nd;
        if (g_elem_threads == 1) {
            for (int k = 0; k < nd * nd; ++k) data[sl[k]] += vol * K[k];
        } else {
            for (int k = 0; k < nd * nd; ++k) {
                const double v = vol * K[k];
                if (v != 0.0) {
#pragma omp atomic
                    data[sl[k]] += v;
                }
            }
        }
    }
}

/* generic P1 load vector (fem_np.load_vector with S = lam + tau (grad u - q) when use_S):
 *   out[(v_a, i)] += sign vol ( ((S + wc C) G_a)_i + w_i det F / (d+1) ),  wc = w_0 + sum_k w_k (xbar_k + ubar_k) */
void oracle_load_scatter(int d, long ne, const int* elems, const double* xyz, const double* u, const double* lam, const double* q, double tau,
                         int use_S, const double* w, int has_w, double sign, double* out) {
#pragma omp parallel for schedule(static) num_threads(g_elem_threads)
    for (long e = 0; e < ne; ++e) {
        double X[12], U[12], G[12], F[9], C[9], M[9], vol;
        for (int a = 0; a <= d; ++a) {
            const long v = elems[e * (d + 1) + a];
            for (int x = 0; x < d; ++x) { X[a * d + x] = xyz[v * d + x]; U[a * d + x] = u ? u[v * d + x] : 0.0; }
        }
        oracle_elem_geom(d, X, G, &vol);
        const double detF = oracle_F_cof(d, u ? U : (const double*)0, G, F, C);
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j) {
                double m = 0.0;
                if (use_S) {
                    const double gu = F[i * d + j] - (i == j ? 1.0 : 0.0);
                    m = lam[e * d * d + i * d + j] + tau * (gu - q[e * d * d + i * d + j]);
                }
                M[i * d + j] = m;
            }
        double add[3] = {0.0, 0.0, 0.0};
        if (has_w) {
            double wc = w[0];
            for (int k = 0; k < d; ++k) {
                double xb = 0.0;
                for (int a = 0; a <= d; ++a) xb += X[a * d + k] + U[a * d + k];
                wc += w[1 + k] * xb / (d + 1);
            }
            for (int k = 0; k < d * d; ++k) M[k] += wc * C[k];
            for (int i = 0; i < d; ++i) add[i] = w[1 + i] * detF / (d + 1);
        }
        for (int a = 0; a <= d; ++a) {
            const long v = elems[e * (d + 1) + a];
            for (int i = 0; i < d; ++i) {
                double s = add[i];
                for (int j = 0; j < d; ++j) s += M[i * d + j] * G[a * d + j];
                const double val = sign * vol * s;
                if (g_elem_threads == 1) out[v * d + i] += val;
                else {
#pragma omp atomic
                    out[v * d + i] += val;
                }
            }
        }
    }
}
