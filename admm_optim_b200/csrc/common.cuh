// Shared infrastructure of libadmm_b200: error handling, device buffers, context, reductions.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <map>
#include <memory>
#include <unordered_map>
#include <stdexcept>
#include <string>
#include <vector>

namespace ab {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define AB_CUDA(call)                                                                                  \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess)                                                                        \
            throw ab::Error(-2, std::string(#call) + " failed: " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
    } while (0)
#define AB_REQUIRE(cond, code, msg)                 \
    do {                                            \
        if (!(cond)) throw ab::Error((code), (msg)); \
    } while (0)

struct Comm;  // multi-GPU communicator (comm.cuh)

struct Context {
    int device = 0;
    cudaStream_t stream = nullptr;
    int num_sms = 148;
    int64_t launches = 0;
    int spmv_variant = 0;   // ADMM_B200_SPMV_VARIANT
    int spmv_waves = 8;     // ADMM_B200_SPMV_WAVES: persistent CTAs per SM for the SpMV kernels
    bool use_graph = true;  // ADMM_B200_GRAPH: replay the BiCGStab iteration as a CUDA graph (single GPU)
    bool own_stream = false;
    int tma_small_ctas = 2; // persistent TMA-SpMV CTAs per SM on levels that fit L2 (tuning key "tma_small_ctas")
    bool l2_hint = true;    // ADMM_B200_L2_HINT: evict-first hint on the matrix stream of levels larger than L2
    bool use_cache = true;  // result caches (VecProd batches, L2Norm components); ADMM_B200_NO_CACHE=1 disables
    bool use_loop = true;   // ADMM_B200_LOOP: the whole BiCGStab loop as one graph launch (conditional WHILE node, device-side ConvCheck)
    bool use_pdl = true;    // ADMM_B200_PDL: programmatic dependent launch for the V-cycle / BiCGStab kernel chain
    int assembly_variant = 0; // ADMM_B200_ASSEMBLY: 0 = row-owner gather (no atomics, reproducible), 1 / "atomic" = per-element atomic scatter
    int spmv2d_lanes = 4;   // ADMM_B200_SPMV2D_LANES: lanes per 2x2-block row in the TMA SpMV (4: 8 rows per warp in flight; 8: quarter-warp rows)
    int coarse_variant = 0; // ADMM_B200_COARSE_VARIANT: 0 = shared-memory-resident blocked Gauss-Jordan, 1 = rows in global memory
    // small device scratch for reductions: partial sums + ticket counters + result slots
    double* d_partials = nullptr;   // kMaxBlocks * kMaxVals
    unsigned int* d_tickets = nullptr;
    double* d_results = nullptr;    // kResultSlots doubles (device)
    double* h_results = nullptr;    // pinned mirror
    std::shared_ptr<Comm> comm;
    static constexpr int kMaxBlocks = 4736;   // 148 SMs x 32
    static constexpr int kMaxVals = 16;
    static constexpr int kResultSlots = 64;
};

// Caching device allocator: per-Newton-iteration matrices / multigrid hierarchies come and go, and raw
// cudaMalloc/cudaFree cost milliseconds each (and cudaFree synchronises the device).  Blocks are recycled by
// (device, size): a block is only ever handed back to a context on the device it was allocated on.  Reuse across
// streams of one device is safe because a block is released by host code that runs after the work using it was
// enqueued, and every context synchronises its stream before it is destroyed; two contexts on the same device that
// run concurrently on different streams must not share vectors (the C ABI never lets them).  trim() returns the
// cached blocks of a device to the driver (ab_context_destroy calls it for its device).
struct DevicePool {
    std::multimap<std::pair<int, size_t>, void*> free_;
    std::unordered_map<void*, std::pair<int, size_t>> size_;
    size_t bytes_allocated = 0;
    static DevicePool& get() { static DevicePool* p = new DevicePool(); return *p; }   // leaked on purpose (CUDA teardown order)
    void* alloc(size_t bytes) {
        bytes = (bytes + 511) & ~(size_t)511;
        int dev = 0;
        cudaGetDevice(&dev);
        auto it = free_.lower_bound({dev, bytes});
        if (it != free_.end() && it->first.first == dev && it->first.second <= bytes + bytes / 8 + 4096) {
            void* p = it->second;
            free_.erase(it);
            return p;
        }
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {   // out of memory: drop this device's cache and retry once
            cudaGetLastError();
            trim(dev);
            e = cudaMalloc(&p, bytes);
        }
        if (e != cudaSuccess) throw ab::Error(-2, std::string("cudaMalloc of ") + std::to_string(bytes) + " bytes failed: " + cudaGetErrorString(e));
        size_[p] = {dev, bytes};
        bytes_allocated += bytes;
        return p;
    }
    void release(void* p) {
        auto it = size_.find(p);
        if (it == size_.end()) { cudaFree(p); return; }
        free_.insert({it->second, p});
    }
    void trim(int dev) {
        for (auto it = free_.begin(); it != free_.end();) {
            if (it->first.first == dev) {
                cudaFree(it->second);
                size_.erase(it->second);
                bytes_allocated -= it->first.second;
                it = free_.erase(it);
            } else ++it;
        }
    }
};

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    explicit DevBuf(size_t n_) { alloc(n_); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t n_) {
        release();
        n = n_;
        if (n) p = (T*)DevicePool::get().alloc(n * sizeof(T) + 16);   // +16: bulk copies may read up to 15 bytes past the end
    }
    void release() {
        if (p) DevicePool::get().release(p);
        p = nullptr;
        n = 0;
    }
    void upload(const T* h, size_t cnt, cudaStream_t s) { AB_CUDA(cudaMemcpyAsync(p, h, cnt * sizeof(T), cudaMemcpyHostToDevice, s)); }
    void upload(const std::vector<T>& h, cudaStream_t s) {
        if (n < h.size()) alloc(h.size());
        if (!h.empty()) { AB_CUDA(cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s)); AB_CUDA(cudaStreamSynchronize(s)); }
    }
    void zero(cudaStream_t s) { if (n) AB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
};

#define AB_LAUNCH(ctx, kernel, grid, block, smem, ...)                  \
    do {                                                                \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__); \
        (ctx)->launches++;                                              \
    } while (0)

// Programmatic dependent launch (PDL): the kernels of the V-cycle / BiCGStab chain are a few microseconds each, so the
// launch latency between two dependent kernels is a large share of the chain.  Launched with the programmatic-stream-
// serialization attribute (a programmatic edge when captured into a graph), kernel N+1 is scheduled while kernel N still
// runs; every such kernel calls pdl_prologue() before its first global-memory access: it releases ITS dependents, then
// blocks until the preceding grid has completed and its writes are visible (griddepcontrol.wait).
template <typename... KArgs, typename... Args>
inline void launch_pdl(Context* ctx, void (*kernel)(KArgs...), int grid, int block, size_t smem, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = ctx->use_pdl ? 1 : 0;
    AB_CUDA(cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...));
    ctx->launches++;
}
#define AB_LAUNCH_PDL(ctx, kernel, grid, block, smem, ...) ab::launch_pdl((ctx), kernel, (grid), (block), (smem), __VA_ARGS__)

// Data that no kernel of the chain writes (matrix values, patterns, transfer tables, the coarse inverse: all final
// before the stream synchronisation that ends every setup) may be fetched BEFORE pdl_wait(): that part of a kernel's
// latency then overlaps the tail of its predecessor.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
    pdl_trigger();
    pdl_wait();
}
// Loads of static data that must be ISSUED before pdl_wait(): volatile asm keeps its program order relative to the
// griddepcontrol instructions (the compiler otherwise sinks or hoists LDG.CONSTANT loads freely across them).
__device__ __forceinline__ double ld_static(const double* p) {
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int ld_static(const int* p) {
    int v;
    asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

inline int grid_for(int64_t n, int block, int max_blocks) {
    int64_t g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > max_blocks) g = max_blocks;
    return (int)g;
}

// ---------------------------------------------------------------------------------------------
// device-side deterministic reduction: per-block partials + "last block" finalisation
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Reduce NV per-thread values over the whole grid (block size: multiple of 32, <= 1024). OP: 0 = sum, 1 = max.
// Results are written to out[0..NV) by the last block to finish, combining the per-block partials in block
// order (bitwise reproducible for a fixed launch configuration).
// Returns true in every thread of the block that finished last (block-uniform), e.g. to append a scalar epilogue
// (after a __syncthreads(): out[] is written by lane 0 of warps 0..NV-1).
template <int NV, int OP>
__device__ __forceinline__ bool grid_reduce(double (&v)[NV], double* __restrict__ partials, unsigned int* __restrict__ ticket,
                                            double* __restrict__ out) {
    __shared__ double sm[NV][32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double w = OP == 0 ? warp_sum(v[k]) : warp_max(v[k]);
        if (lane == 0) sm[k][warp] = w;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double acc = sm[threadIdx.x][0];
        for (int w = 1; w < nwarp; ++w) acc = OP == 0 ? acc + sm[threadIdx.x][w] : fmax(acc, sm[threadIdx.x][w]);
        partials[(size_t)blockIdx.x * NV + threadIdx.x] = acc;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        for (int k = warp; k < NV; k += nwarp) {      // warp k reduces value k over all blocks -- fixed order
            double acc = OP == 0 ? 0.0 : -1.0e300;
            for (unsigned int b = lane; b < gridDim.x; b += 32) {
                double p = partials[(size_t)b * NV + k];
                acc = OP == 0 ? acc + p : fmax(acc, p);
            }
            acc = OP == 0 ? warp_sum(acc) : warp_max(acc);
            if (lane == 0) out[k] = acc;
        }
        if (threadIdx.x == 0) *ticket = 0u;
    }
    return is_last;
}

// ---------------------------------------------------------------------------------------------
// mbarrier + 1-D bulk async copy (TMA engine, no tensor map needed) -- sm_90+ PTX, used by k_bsr_spmv_tma
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
// global -> shared bulk copy; src, dst 16-byte aligned, bytes a multiple of 16; completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// same with an L2 eviction-priority hint: the matrix stream is read exactly once per product, marking it evict-first keeps the
// gathered vector (re-read ~27 times per row neighbourhood) resident in L2 instead of being flushed by the stream
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}

}  // namespace ab
