// Multi-GPU communication layer: one process per GPU, NCCL over NVLink/NVSwitch.
// Replaces UG4's pcl/MPI layer (`mpirun -np 4 ugshell ...`, 3d_admm.lua:25; storage-type conversions
// 3d_admm.lua:912,982,1096,1222): interface sums (additive -> consistent) and scalar all-reduces.
// NCCL is resolved at run time from the library already loaded in the process (torch's bundled
// libnccl.so.2) so that libadmm_b200.so has no link-time dependency and never mixes two NCCL copies.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"
#include "iface_xchg.cuh"
#include "gershgorin_dist.cuh"

namespace ab {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    static NcclApi& get() {
        static NcclApi api;
        static bool loaded = false;
        if (!loaded) {
            void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // torch's copy when already loaded
            if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
            AB_REQUIRE(h, -4, std::string("cannot load libnccl.so.2: ") + dlerror());
#define AB_SYM(name) api.name = reinterpret_cast<decltype(api.name)>(dlsym(h, "nccl" #name)); AB_REQUIRE(api.name, -4, "libnccl.so.2 lacks nccl" #name)
            AB_SYM(GetUniqueId); AB_SYM(CommInitRank); AB_SYM(CommDestroy); AB_SYM(CommAbort); AB_SYM(AllReduce); AB_SYM(Send); AB_SYM(Recv);
            AB_SYM(GroupStart); AB_SYM(GroupEnd); AB_SYM(GetErrorString);
#undef AB_SYM
            loaded = true;
        }
        return api;
    }
};

#define AB_NCCL(call)                                                                                                       \
    do {                                                                                                                    \
        ncclResult_t r__ = (call);                                                                                          \
        if (r__ != ncclSuccess) throw ab::Error(-2, std::string(#call) + " failed: " + ab::NcclApi::get().GetErrorString(r__)); \
    } while (0)

struct Comm {
    int rank = 0, nranks = 1;
    ncclComm_t comm = nullptr;
    // ncclCommDestroy is an intra-node collective (it blocks until every rank of the node calls it); a context may be torn down
    // by a garbage collector at a different time on every rank, or not at all on a rank that is exiting.  All work of the
    // context has been synchronised by then (ab_context_destroy), so the communicator is released with ncclCommAbort, which
    // frees the local resources without waiting for the peers.
    ~Comm() { if (comm) NcclApi::get().CommAbort(comm); }
    void allreduce(double* d, int n, bool max_op, cudaStream_t s) {
        AB_NCCL(NcclApi::get().AllReduce(d, d, (size_t)n, ncclFloat64, max_op ? ncclMax : ncclSum, comm, s));
    }
};

// interface (shared-vertex) lists of one level: for every neighbour rank the local vertex ids of the shared
// vertices in a canonical order both sides agree on (sorted by coordinates at setup)
struct Interface {
    std::vector<int> neigh, offset;   // neighbour ranks (ascending), offset[n]..offset[n+1] into idx
    DevBuf<int> idx;                  // concatenated per-neighbour local vertex ids ("slots")
    DevBuf<int> iv;                   // unique interface vertices, ordered by (first neighbour, slot): the stores into the first
                                      // neighbour's window are then consecutive
    DevBuf<int> iv_ptr, iv_slot, iv_nb;   // CSR unique vertex -> its slots and their neighbour index, neighbour ranks ascending
    DevBuf<unsigned char> owned;      // per local vertex: 1 when this rank is the lowest rank sharing it
    DevBuf<double> send, recv, save;  // NCCL fallback: packed buffers (total * maxcomp), smoother scratch (niv * 2 * D)
    int total = 0, niv = 0, my_pos = 0;   // my_pos: number of neighbours with a rank below mine
    // peer-to-peer path (NVLink, CUDA IPC): neighbours write straight into this rank's window
    DevBuf<int> d_offset, d_neigh;                 // device copies of offset / neigh
    DevBuf<unsigned long long> d_peer_dst;          // per neighbour: address (in this process) of my slot in the neighbour's window, parity 0
    DevBuf<unsigned long long> d_peer_stride;       // per neighbour: byte distance between the neighbour's two parity buffers
    DevBuf<unsigned long long> d_peer_flag;         // per neighbour: address of the neighbour's flag word for this rank
    double* win_recv = nullptr;                     // my two parity buffers inside the window (total*D doubles each)
    unsigned long long* win_flags = nullptr;        // my flag words (one per rank) for this level
    DevBuf<unsigned long long> state;               // device side: [0] epoch of the last completed exchange, [1] arrive counter, [2] done counter
};


// zero the copies this rank does not own (consistent -> unique, a valid additive representation)
__global__ void k_zero_not_owned(int64_t n, int D, const unsigned char* __restrict__ owned, double* __restrict__ v) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        if (!owned[t / D]) v[t] = 0.0;
}
// dot products that count every shared vertex once: sum over owned vertices of x_k . y
template <int NX>
__global__ void __launch_bounds__(256) k_dot_owned(int64_t n, int D, const unsigned char* __restrict__ owned, const double* x0, const double* x1,
                                                   const double* __restrict__ y, double* partials, unsigned int* ticket, double* out) {
    const double* xs[2] = {x0, x1};
    double v[NX];
#pragma unroll
    for (int k = 0; k < NX; ++k) v[k] = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (owned[i / D]) {
            const double yi = y[i];
#pragma unroll
            for (int k = 0; k < NX; ++k) v[k] += xs[k][i] * yi;
        }
    }
    grid_reduce<NX, 0>(v, partials, ticket, out);
}
// smoother interface fix-up (see Gmg::smooth): save d and x at the interface vertices before the fused step ...
__global__ void k_iface_save(int niv, int D, const int* __restrict__ iv, const double* __restrict__ d, const double* __restrict__ x,
                             double* __restrict__ save) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < niv * D; t += gridDim.x * blockDim.x) {
        const int k = t / D, c = t - k * D;
        const int64_t i = (int64_t)iv[k] * D + c;
        save[t] = d ? d[i] : 0.0;
        save[niv * D + t] = x ? x[i] : 0.0;
    }
}
// ... turn the locally updated d into the additive increment inc = d_new_local - c1*d_old ...
__global__ void k_iface_inc(int niv, int D, const int* __restrict__ iv, double c1, const double* __restrict__ save, double* __restrict__ d) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < niv * D; t += gridDim.x * blockDim.x) {
        const int k = t / D, c = t - k * D;
        const int64_t i = (int64_t)iv[k] * D + c;
        d[i] = d[i] - c1 * save[t];
    }
}
// ... and after the interface sum rebuild d = c1*d_old + inc_total, x_new = x_old + d
__global__ void k_iface_fix(int niv, int D, const int* __restrict__ iv, double c1, const double* __restrict__ save, double* __restrict__ d,
                            double* __restrict__ xnew) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < niv * D; t += gridDim.x * blockDim.x) {
        const int k = t / D, c = t - k * D;
        const int64_t i = (int64_t)iv[k] * D + c;
        const double dn = c1 * save[t] + d[i];
        d[i] = dn;
        xnew[i] = save[niv * D + t] + dn;
    }
}

}  // namespace ab
