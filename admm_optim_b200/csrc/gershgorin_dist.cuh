// Kernels of the exact, partition-independent Gershgorin bound of a domain-decomposed operator (lib.cu gmg_setup_kernels) and the
// pack / unpack kernels of the NCCL interface path.  Kept free of every other dependency so that
// tests/cuda_host_shim/gershgorin_emulation.cpp can compile THIS source for the host and tests/dist_host_worker.py can drive the
// real kernel code on 2-4 gloo ranks against the row sums of the global operator (no GPU needed for the index arithmetic).
#pragma once
#include <cmath>
#include <cstdint>

#ifndef AB_HOST_EMULATION
#define AB_GD_KERNEL __global__ void
#endif

namespace ab {

// point-Jacobi data, distributed: additive diagonal and additive absolute row sums (made consistent by an interface sum)
template <int D>
AB_GD_KERNEL k_diag_rowabs(int nb, const int* __restrict__ rowptr, const int* __restrict__ diagpos, const double* __restrict__ vals,
                              double* __restrict__ diag, double* __restrict__ rowabs) {
    constexpr int DD = D * D;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < (int64_t)nb * D; t += (int64_t)gridDim.x * blockDim.x) {
        const int row = (int)(t / D), r = (int)(t - (int64_t)row * D);
        const int s = rowptr[row], e = rowptr[row + 1];
        double sum = 0.0;
        for (int k = s; k < e; ++k) {
#pragma unroll
            for (int c = 0; c < D; ++c) sum += fabs(vals[(int64_t)k * DD + r * D + c]);
        }
        diag[t] = vals[(int64_t)diagpos[row] * DD + r * D + r];
        rowabs[t] = sum;
    }
}
// exact row sums of an additive operator (multi-GPU): compact copy of the blocks shared with neighbour ranks ...
AB_GD_KERNEL k_pack_blocks(int64_t n, int DD, const int* __restrict__ bpos, const double* __restrict__ vals, double* __restrict__ cv) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = t / DD;
        cv[t] = vals[(int64_t)bpos[k] * DD + (t - k * DD)];
    }
}
// ... and, once the neighbours' parts were added to cv, the correction of the local row sums: every rank holding a shared block
// contributes |sum| / mult instead of |its own part|, so that the interface sum of the rows counts |sum| exactly once
template <int D>
AB_GD_KERNEL k_rowabs_fix(int nsb, const int* __restrict__ bpos, const int* __restrict__ brow, const int* __restrict__ mult,
                             const double* __restrict__ vals, const double* __restrict__ cv, double* rowabs) {
    constexpr int DD = D * D;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < (int64_t)nsb * D; t += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(t / D), r = (int)(t - (int64_t)k * D);
        const double inv = 1.0 / (double)mult[k];
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < D; ++c) s += fabs(cv[(int64_t)k * DD + r * D + c]) * inv - fabs(vals[(int64_t)bpos[k] * DD + r * D + c]);
        atomicAdd(rowabs + (int64_t)brow[k] * D + r, s);
    }
}
// pack / unpack of interface slots (P1 vectors: D = components; shared matrix blocks: D = d*d values per block): the NCCL form of
// the interface sum and the exchange of the shared blocks at solver:init
AB_GD_KERNEL k_iface_pack(int total, int D, const int* __restrict__ idx, const double* __restrict__ v, double* __restrict__ buf) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total * D; t += gridDim.x * blockDim.x) {
        const int k = t / D, c = t - k * D;
        buf[t] = v[(int64_t)idx[k] * D + c];
    }
}
AB_GD_KERNEL k_iface_unpack_add(int total, int D, const int* __restrict__ idx, const double* __restrict__ buf, double* __restrict__ v) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total * D; t += gridDim.x * blockDim.x) {
        const int k = t / D, c = t - k * D;
        atomicAdd(v + (int64_t)idx[k] * D + c, buf[t]);
    }
}

}  // namespace ab
