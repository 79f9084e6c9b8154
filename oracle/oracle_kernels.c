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
                if (j == i) d = a[k];
                else s -= a[k] * ((j >= lo && j < hi) ? x[j] : xold[j]);
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

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
