// HOST EMULATION of the exact-Gershgorin kernels of a decomposed operator and of the interface pack / unpack kernels
// (admm_optim_b200/csrc/gershgorin_dist.cuh) -- test infrastructure.  The kernel SOURCE the library ships is compiled for the
// CPU and every (block, thread) of a launch is run in turn (these kernels have no intra-grid synchronisation; atomicAdd becomes
// a plain add).  tests/dist_host_worker.py loads this as a shared library and drives the device steps of lib.cu
// gmg_setup_kernels with it on 2-4 gloo ranks: k_diag_rowabs -> k_pack_blocks -> k_iface_pack -> (exchange) ->
// k_iface_unpack_add -> k_rowabs_fix -> interface sum, against the row sums of the GLOBAL operator.
#include <cstdint>

struct EmuDim { unsigned x = 0; };
static EmuDim blockIdx, threadIdx, blockDim, gridDim;
#define AB_HOST_EMULATION 1
#define AB_GD_KERNEL void
#define __restrict__
static inline double atomicAdd(double* p, double v) { const double old = *p; *p = old + v; return old; }

// the library itself (loaded in the same process by the worker) exports host stubs of these kernels under the same mangled names:
// give the host-compiled copies their own namespace so that the dynamic linker cannot bind the calls below to the stubs
#define ab ab_host_emulation
#include "gershgorin_dist.cuh"

template <class F>
static void launch(int grid, int block, F f) {
    gridDim.x = grid; blockDim.x = block;
    for (int b = 0; b < grid; ++b)
        for (int t = 0; t < block; ++t) { blockIdx.x = b; threadIdx.x = t; f(); }
}

extern "C" {
// a deliberately small launch shape (fewer threads than entries): the grid-stride loops are part of what is checked
void emu_diag_rowabs(int D, int nb, const int* rowptr, const int* diagpos, const double* vals, double* diag, double* rowabs) {
    if (D == 2) launch(3, 8, [&] { ab::k_diag_rowabs<2>(nb, rowptr, diagpos, vals, diag, rowabs); });
    else launch(3, 8, [&] { ab::k_diag_rowabs<3>(nb, rowptr, diagpos, vals, diag, rowabs); });
}
void emu_pack_blocks(int64_t n, int DD, const int* bpos, const double* vals, double* cv) {
    launch(2, 16, [&] { ab::k_pack_blocks(n, DD, bpos, vals, cv); });
}
void emu_iface_pack(int total, int D, const int* idx, const double* v, double* buf) {
    launch(2, 16, [&] { ab::k_iface_pack(total, D, idx, v, buf); });
}
void emu_iface_unpack_add(int total, int D, const int* idx, const double* buf, double* v) {
    launch(2, 16, [&] { ab::k_iface_unpack_add(total, D, idx, buf, v); });
}
void emu_rowabs_fix(int D, int nsb, const int* bpos, const int* brow, const int* mult, const double* vals, const double* cv, double* rowabs) {
    if (D == 2) launch(3, 8, [&] { ab::k_rowabs_fix<2>(nsb, bpos, brow, mult, vals, cv, rowabs); });
    else launch(3, 8, [&] { ab::k_rowabs_fix<3>(nsb, bpos, brow, mult, vals, cv, rowabs); });
}
}
