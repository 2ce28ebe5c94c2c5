// Device layer behind include/descent_cuda.h: context, stream-ordered buffers, pinned staging,
// NVRTC JIT for sm_100a, launches, CUDA-graph replay, events, NCCL data parallel.
// Replaces the reference's Vulkan runtime (src/device/*.rs) and shader-module creation
// (src/kernel.rs:946-1034); nothing here is a translation of it -- the mechanisms are CUDA's own
// (memory pools instead of a buddy heap, graphs instead of command buffers, events instead of
// timestamp query pools).
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <nvrtc.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/descent_cuda.h"

namespace {

thread_local std::string g_last_error;

int set_error(int code, const char* fmt, ...) {
    char buf[4096];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                                      \
    do {                                                                                                    \
        cudaError_t e_ = (expr);                                                                            \
        if (e_ != cudaSuccess) return set_error(DSC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// Driver entry points come through the runtime so the library never links libcuda directly
// (it must load on a GPU-less build box for the symbol/compile checks).
struct DriverApi {
    CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*ModuleUnload)(CUmodule) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
    CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
    CUresult (*TensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill) = nullptr;
    bool loaded = false;
};
DriverApi g_drv;

int load_driver_api() {
    if (g_drv.loaded) return DSC_OK;
    auto get = [](const char* name, void** fn) -> int {
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || *fn == nullptr)
            return set_error(DSC_ERR_CUDA, "driver entry point %s unavailable: %s", name, cudaGetErrorString(e));
        return DSC_OK;
    };
    int rc;
    if ((rc = get("cuModuleLoadData", (void**)&g_drv.ModuleLoadData))) return rc;
    if ((rc = get("cuModuleUnload", (void**)&g_drv.ModuleUnload))) return rc;
    if ((rc = get("cuModuleGetFunction", (void**)&g_drv.ModuleGetFunction))) return rc;
    if ((rc = get("cuLaunchKernel", (void**)&g_drv.LaunchKernel))) return rc;
    if ((rc = get("cuFuncSetAttribute", (void**)&g_drv.FuncSetAttribute))) return rc;
    if ((rc = get("cuGetErrorString", (void**)&g_drv.GetErrorString))) return rc;
    if ((rc = get("cuTensorMapEncodeTiled", (void**)&g_drv.TensorMapEncodeTiled))) return rc;
    g_drv.loaded = true;
    return DSC_OK;
}

int cu_error(CUresult r, const char* what) {
    const char* msg = nullptr;
    if (g_drv.GetErrorString) g_drv.GetErrorString(r, &msg);
    return set_error(DSC_ERR_CUDA, "%s failed: %s", what, msg ? msg : "unknown driver error");
}
#define CU_TRY(expr)                                 \
    do {                                             \
        CUresult r_ = (expr);                        \
        if (r_ != CUDA_SUCCESS) return cu_error(r_, #expr); \
    } while (0)

// NCCL is optional (N=1 never touches it) and is resolved at run time so that importing torch first
// and this library second share one libnccl.
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
    if (g_nccl.handle) return DSC_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return set_error(DSC_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(h, "ncclAllGather");  // optional: only the peer-memory set-up uses it
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy || !g_nccl.GetErrorString)
        return set_error(DSC_ERR_NCCL, "libnccl.so.2 is missing expected symbols");
    g_nccl.handle = h;
    return DSC_OK;
}
#define NCCL_TRY(expr)                                                                                  \
    do {                                                                                                \
        ncclResult_t r_ = (expr);                                                                       \
        if (r_ != ncclSuccess) return set_error(DSC_ERR_NCCL, "%s failed: %s", #expr, g_nccl.GetErrorString(r_)); \
    } while (0)

__global__ void dsc_set_u32_kernel(unsigned* p, unsigned v) { *p = v; }
__global__ void dsc_fill_u32_kernel(unsigned* p, unsigned v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

// ---- one-shot all-reduce over peer-mapped memory (the small, late gradient bucket) -------------------------------
// NCCL's latency for a 20 KB message (15-35 us at 2-8 ranks) is all exposed at the end of the backward pass.  Every
// rank owns an exchange area (two data slots + flags + an epoch counter) that its peers map through CUDA IPC over
// NVLink.  One CTA per rank: publish my values into my slot (epoch parity), release-store the epoch into my flag on
// every peer, spin until every peer's flag in my area shows the epoch, then read all slots with P2P loads and add them
// in rank order -- the same order on every rank, so the replicas stay bitwise identical.  No second barrier: a peer
// cannot overwrite the slot of epoch e before it has seen my flag of epoch e + 1, which I send after reading.
constexpr int kXchgMaxRanks = 16;
constexpr size_t kXchgSlotFloats = 64 * 1024;  // 256 KB per slot
struct XchgArea {
    unsigned epoch;
    unsigned done;   // CTAs of the current launch that have finished (the last one advances `epoch`)
    unsigned pad[30];
    unsigned flags[2][kXchgMaxRanks * 32];  // one 128-byte line per (slot, peer)
    float data[2][kXchgSlotFloats];
};
struct XchgPeers { XchgArea* area[kXchgMaxRanks]; };

// Several CTAs per rank (round 2: the hash tables' 150 KB bucket took 20+ us of dependent P2P loads in one CTA): CTA c owns a
// contiguous slice of the bucket and has its own flag per (slot, peer) -- 32 flags fit the peer's 128-byte line -- so the
// CTAs never wait for each other; the last CTA to finish advances the epoch (it can only be last once every CTA of this
// launch has read the old value).
__global__ void __launch_bounds__(1024) dsc_oneshot_allreduce_kernel(float* bucket, unsigned count, XchgPeers peers, int rank, int world) {
    XchgArea* mine = peers.area[rank];
    const unsigned epoch = mine->epoch + 1u, slot = epoch & 1u;
    const unsigned tid = threadIdx.x, cta = blockIdx.x;
    const unsigned chunk = ((count + gridDim.x - 1u) / gridDim.x + 3u) & ~3u;
    const unsigned begin = min(count, cta * chunk), end = min(count, begin + chunk);
    for (unsigned i = begin + tid * 4u; i < end; i += 4096u) {
        if (i + 4u <= end) *reinterpret_cast<float4*>(&mine->data[slot][i]) = *reinterpret_cast<const float4*>(bucket + i);
        else for (unsigned j = i; j < end; ++j) mine->data[slot][j] = bucket[j];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < (unsigned)world) {
        unsigned* flag = &peers.area[tid]->flags[slot][rank * 32 + cta];
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
        const unsigned* wait_on = &mine->flags[slot][tid * 32 + cta];
        unsigned seen;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(wait_on) : "memory");
        } while (seen != epoch);
    }
    __syncthreads();
    for (unsigned i = begin + tid * 4u; i < end; i += 4096u) {
        if (i + 4u <= end) {
            float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int p = 0; p < world; ++p) {
                float4 v;
                asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(&peers.area[p]->data[slot][i]) : "memory");
                sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
            }
            *reinterpret_cast<float4*>(bucket + i) = sum;
        } else {
            for (unsigned j = i; j < end; ++j) {
                float sum = 0.f;
                for (int p = 0; p < world; ++p) {
                    float v;
                    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(&peers.area[p]->data[slot][j]) : "memory");
                    sum += v;
                }
                bucket[j] = sum;
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(&mine->done, 1u) == gridDim.x - 1u) {
            mine->done = 0u;
            __threadfence();
            mine->epoch = epoch;
        }
    }
}

constexpr size_t kStagingSlotBytes = 16u << 20;
constexpr int kStagingSlots = 2;

}  // namespace

struct dsc_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    unsigned* step_params = nullptr;  // device: [0] rand_seed
    char* staging[kStagingSlots] = {nullptr, nullptr};
    cudaEvent_t staging_done[kStagingSlots] = {nullptr, nullptr};
    int staging_next = 0;
    cudaStream_t copy_stream = nullptr;   // host -> device prefetch of the next batch, overlapping the running graph
    cudaEvent_t prefetch_done = nullptr;  // recorded on copy_stream after a prefetch
    cudaEvent_t staged_read = nullptr;    // recorded on stream after a commit read the staging buffer
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0;
    cudaStream_t comm_stream = nullptr;   // early gradient bucket: all-reduced here while the backward pass continues on `stream`
    cudaEvent_t comm_fork = nullptr, comm_join = nullptr;
    bool comm_pending = false;
    // independent kernels of one dependency level on parallel branches (dsc_branch_*): kernels, fills and copies go to
    // `launch_stream`, which is `stream` outside a fork
    static constexpr int kBranches = 16;
    cudaStream_t branch[kBranches] = {};
    cudaEvent_t branch_fork = nullptr, branch_done[kBranches] = {};
    bool branch_used[kBranches] = {};
    bool forked = false;
    cudaStream_t launch_stream = nullptr;
    XchgPeers xchg;             // peer-mapped exchange areas (own entry = local allocation); valid when xchg_ready
    bool xchg_ready = false;
    bool capturing = false;
};
struct dsc_module {
    dsc_ctx* ctx;
    CUmodule module;
};
struct dsc_graph {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
};

extern "C" {

const char* dsc_last_error(void) { return g_last_error.c_str(); }

int dsc_device_count(int* count) {
    *count = 0;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        return set_error(DSC_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    return DSC_OK;
}

int dsc_ctx_create(int device, dsc_ctx** out) {
    *out = nullptr;
    int count = 0;
    CUDA_TRY(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return set_error(DSC_ERR_INVALID, "device %d out of range (%d devices)", device, count);
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaFree(nullptr));  // create the primary context
    int rc = load_driver_api();
    if (rc) return rc;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return set_error(DSC_ERR_UNSUPPORTED, "device %d is sm_%d%d; this backend is written for sm_100a (B200)", device, prop.major, prop.minor);
    dsc_ctx* ctx = new dsc_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->launch_stream = ctx->stream;
    CUDA_TRY(cudaEventCreateWithFlags(&ctx->branch_fork, cudaEventDisableTiming));  // (created here: nothing is created while a graph is being captured)
    for (int b = 0; b < dsc_ctx::kBranches; ++b) {
        CUDA_TRY(cudaStreamCreateWithFlags(&ctx->branch[b], cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&ctx->branch_done[b], cudaEventDisableTiming));
    }
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&ctx->prefetch_done, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&ctx->staged_read, cudaEventDisableTiming));
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&ctx->comm_fork, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&ctx->comm_join, cudaEventDisableTiming));
    cudaMemPool_t pool;
    CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t threshold = UINT64_MAX;  // keep freed memory cached in the pool: alloc/free cost no driver calls
    CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
    CUDA_TRY(cudaMalloc(&ctx->step_params, 4 * sizeof(unsigned)));
    CUDA_TRY(cudaMemset(ctx->step_params, 0, 4 * sizeof(unsigned)));
    for (int i = 0; i < kStagingSlots; ++i) {
        CUDA_TRY(cudaHostAlloc(&ctx->staging[i], kStagingSlotBytes, cudaHostAllocDefault));
        CUDA_TRY(cudaEventCreateWithFlags(&ctx->staging_done[i], cudaEventDisableTiming));
    }
    *out = ctx;
    return DSC_OK;
}

int dsc_ctx_destroy(dsc_ctx* ctx) {
    if (!ctx) return DSC_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
    for (int i = 0; i < kStagingSlots; ++i) {
        if (ctx->staging[i]) cudaFreeHost(ctx->staging[i]);
        if (ctx->staging_done[i]) cudaEventDestroy(ctx->staging_done[i]);
    }
    cudaFree(ctx->step_params);
    cudaStreamSynchronize(ctx->copy_stream);
    cudaEventDestroy(ctx->prefetch_done);
    cudaEventDestroy(ctx->staged_read);
    if (ctx->xchg_ready) {
        for (int p = 0; p < ctx->world; ++p) {
            // the local area is deliberately not freed: a slower peer's last reduction may still be reading it (0.5 MB per
            // context, returned at process exit)
            if (p != ctx->rank) cudaIpcCloseMemHandle(ctx->xchg.area[p]);
        }
        ctx->xchg_ready = false;
    }
    cudaStreamSynchronize(ctx->comm_stream);
    cudaEventDestroy(ctx->comm_fork);
    cudaEventDestroy(ctx->comm_join);
    cudaStreamDestroy(ctx->comm_stream);
    cudaStreamDestroy(ctx->copy_stream);
    if (ctx->branch_fork) {
        cudaEventDestroy(ctx->branch_fork);
        for (int b = 0; b < dsc_ctx::kBranches; ++b) { cudaEventDestroy(ctx->branch_done[b]); cudaStreamDestroy(ctx->branch[b]); }
    }
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return DSC_OK;
}

int dsc_ctx_device(dsc_ctx* ctx, int* device) { *device = ctx->device; return DSC_OK; }
int dsc_ctx_stream(dsc_ctx* ctx, void** s) { *s = (void*)ctx->stream; return DSC_OK; }
int dsc_ctx_sm_count(dsc_ctx* ctx, int* count) { *count = ctx->sm_count; return DSC_OK; }

int dsc_alloc(dsc_ctx* ctx, size_t bytes, uint64_t* id) {
    CUDA_TRY(cudaSetDevice(ctx->device));
    void* p = nullptr;
    CUDA_TRY(cudaMallocAsync(&p, bytes ? bytes : 4, ctx->stream));
    *id = (uint64_t)p;
    return DSC_OK;
}
int dsc_free(dsc_ctx* ctx, uint64_t id) {
    if (!id) return DSC_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaFreeAsync((void*)id, ctx->stream));
    return DSC_OK;
}
int dsc_fill_u32(dsc_ctx* ctx, uint64_t id, size_t offset_bytes, uint32_t value, size_t count) {
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (count == 0) return DSC_OK;
    unsigned* p = (unsigned*)(id + offset_bytes);
    if (value == 0) {
        CUDA_TRY(cudaMemsetAsync(p, 0, count * 4, ctx->launch_stream));
    } else {
        unsigned blocks = (unsigned)std::min<size_t>((count + 255) / 256, (size_t)ctx->sm_count * 8);
        dsc_fill_u32_kernel<<<blocks, 256, 0, ctx->launch_stream>>>(p, value, count);
        CUDA_TRY(cudaGetLastError());
    }
    return DSC_OK;
}
int dsc_copy(dsc_ctx* ctx, uint64_t dst, uint64_t src, size_t bytes) {
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaMemcpyAsync((void*)dst, (const void*)src, bytes, cudaMemcpyDeviceToDevice, ctx->launch_stream));
    return DSC_OK;
}

int dsc_upload(dsc_ctx* ctx, uint64_t id, size_t offset, const void* src, size_t n, size_t zero_tail_to, int src_is_pinned) {
    CUDA_TRY(cudaSetDevice(ctx->device));
    char* dst = (char*)id + offset;
    if (src_is_pinned) {
        if (n) CUDA_TRY(cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, ctx->stream));
    } else {
        const char* s = (const char*)src;
        size_t done = 0;
        while (done < n) {
            int slot = ctx->staging_next;
            ctx->staging_next = (slot + 1) % kStagingSlots;
            CUDA_TRY(cudaEventSynchronize(ctx->staging_done[slot]));
            size_t chunk = std::min(kStagingSlotBytes, n - done);
            memcpy(ctx->staging[slot], s + done, chunk);
            CUDA_TRY(cudaMemcpyAsync(dst + done, ctx->staging[slot], chunk, cudaMemcpyHostToDevice, ctx->stream));
            CUDA_TRY(cudaEventRecord(ctx->staging_done[slot], ctx->stream));
            done += chunk;
        }
    }
    if (zero_tail_to > offset + n) CUDA_TRY(cudaMemsetAsync(dst + n, 0, zero_tail_to - offset - n, ctx->stream));
    return DSC_OK;
}
int dsc_prefetch(dsc_ctx* ctx, uint64_t staging, const void* pinned_src, size_t n) {
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, ctx->staged_read, 0));  // the last commit has finished reading the staging buffers
    if (n) CUDA_TRY(cudaMemcpyAsync((void*)staging, pinned_src, n, cudaMemcpyHostToDevice, ctx->copy_stream));
    CUDA_TRY(cudaEventRecord(ctx->prefetch_done, ctx->copy_stream));
    return DSC_OK;
}
int dsc_prefetch_commit(dsc_ctx* ctx, uint64_t dst, uint64_t staging, size_t n) {
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->prefetch_done, 0));
    if (n) CUDA_TRY(cudaMemcpyAsync((void*)dst, (const void*)staging, n, cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_TRY(cudaEventRecord(ctx->staged_read, ctx->stream));
    return DSC_OK;
}
int dsc_download(dsc_ctx* ctx, uint64_t id, size_t offset, void* dst, size_t n) {
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (n) CUDA_TRY(cudaMemcpyAsync(dst, (const char*)id + offset, n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return DSC_OK;
}
int dsc_host_alloc(size_t bytes, void** out) {
    CUDA_TRY(cudaHostAlloc(out, bytes ? bytes : 4, cudaHostAllocDefault));
    return DSC_OK;
}
int dsc_host_free(void* p) {
    if (p) CUDA_TRY(cudaFreeHost(p));
    return DSC_OK;
}

int dsc_nvrtc_compile(const char* cuda_source, const char* const* options, int num_options, void** cubin, size_t* bytes) {
    *cubin = nullptr;
    *bytes = 0;
    nvrtcProgram prog;
    nvrtcResult r = nvrtcCreateProgram(&prog, cuda_source, "descent_kernels.cu", 0, nullptr, nullptr);
    if (r != NVRTC_SUCCESS) return set_error(DSC_ERR_NVRTC, "nvrtcCreateProgram: %s", nvrtcGetErrorString(r));
    std::vector<const char*> opts = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "--extra-device-vectorization"};
    for (int i = 0; i < num_options; ++i) opts.push_back(options[i]);
    r = nvrtcCompileProgram(prog, (int)opts.size(), opts.data());
    if (r != NVRTC_SUCCESS) {
        size_t log_size = 0;
        nvrtcGetProgramLogSize(prog, &log_size);
        std::string log(log_size, '\0');
        nvrtcGetProgramLog(prog, log.data());
        nvrtcDestroyProgram(&prog);
        if (log.size() > 3500) log.resize(3500);
        return set_error(DSC_ERR_NVRTC, "NVRTC compilation failed (%s):\n%s", nvrtcGetErrorString(r), log.c_str());
    }
    size_t size = 0;
    r = nvrtcGetCUBINSize(prog, &size);
    if (r != NVRTC_SUCCESS || size == 0) {
        nvrtcDestroyProgram(&prog);
        return set_error(DSC_ERR_NVRTC, "nvrtcGetCUBINSize: %s", nvrtcGetErrorString(r));
    }
    void* data = malloc(size);
    nvrtcGetCUBIN(prog, (char*)data);
    nvrtcDestroyProgram(&prog);
    *cubin = data;
    *bytes = size;
    return DSC_OK;
}
int dsc_host_buffer_free(void* p) { free(p); return DSC_OK; }

int dsc_module_load_cubin(dsc_ctx* ctx, const void* cubin, size_t bytes, dsc_module** out) {
    (void)bytes;
    *out = nullptr;
    CUDA_TRY(cudaSetDevice(ctx->device));
    int rc = load_driver_api();
    if (rc) return rc;
    CUmodule m;
    CU_TRY(g_drv.ModuleLoadData(&m, cubin));
    *out = new dsc_module{ctx, m};
    return DSC_OK;
}
int dsc_module_jit(dsc_ctx* ctx, const char* cuda_source, const char* const* options, int num_options, dsc_module** out) {
    void* cubin = nullptr;
    size_t bytes = 0;
    int rc = dsc_nvrtc_compile(cuda_source, options, num_options, &cubin, &bytes);
    if (rc) return rc;
    rc = dsc_module_load_cubin(ctx, cubin, bytes, out);
    free(cubin);
    return rc;
}
int dsc_module_get_kernel(dsc_module* module, const char* entry, dsc_kernel* out) {
    CUfunction f;
    CUDA_TRY(cudaSetDevice(module->ctx->device));
    CUresult r = g_drv.ModuleGetFunction(&f, module->module, entry);
    if (r != CUDA_SUCCESS) return set_error(DSC_ERR_INVALID, "kernel '%s' not found in module", entry);
    *out = (dsc_kernel)f;
    return DSC_OK;
}
int dsc_module_destroy(dsc_module* module) {
    if (!module) return DSC_OK;
    cudaSetDevice(module->ctx->device);
    g_drv.ModuleUnload(module->module);
    delete module;
    return DSC_OK;
}
int dsc_kernel_set_max_dynamic_smem(dsc_kernel kernel, int bytes) {
    CU_TRY(g_drv.FuncSetAttribute((CUfunction)kernel, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, bytes));
    return DSC_OK;
}

int dsc_launch(dsc_ctx* ctx, dsc_kernel kernel, uint32_t gx, uint32_t gy, uint32_t gz, uint32_t block_x, uint32_t smem,
               const uint64_t* buffers, int num_buffers) {
    if (num_buffers < 0 || num_buffers > 254) return set_error(DSC_ERR_INVALID, "too many kernel buffers (%d)", num_buffers);
    if (gx == 0 || gy == 0 || gz == 0) return DSC_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    uint64_t values[256];  // 4 KB of kernel parameters hold 512 pointers; multi-tensor optimiser launches bind ~8 per tensor
    void* params[256];
    for (int i = 0; i < num_buffers; ++i) {
        values[i] = buffers[i];
        params[i] = &values[i];
    }
    values[num_buffers] = (uint64_t)ctx->step_params;
    params[num_buffers] = &values[num_buffers];
    CU_TRY(g_drv.LaunchKernel((CUfunction)kernel, gx, gy, gz, block_x, 1, 1, smem, (CUstream)ctx->launch_stream, params, nullptr));
    return DSC_OK;
}
int dsc_set_rand_seed(dsc_ctx* ctx, uint32_t rand_seed) {
    CUDA_TRY(cudaSetDevice(ctx->device));
    dsc_set_u32_kernel<<<1, 1, 0, ctx->stream>>>(ctx->step_params, rand_seed);
    CUDA_TRY(cudaGetLastError());
    return DSC_OK;
}

int dsc_graph_begin_capture(dsc_ctx* ctx) {
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    ctx->capturing = true;
    return DSC_OK;
}
int dsc_graph_end_capture(dsc_ctx* ctx, dsc_graph** out) {
    *out = nullptr;
    ctx->capturing = false;
    cudaGraph_t g = nullptr;
    CUDA_TRY(cudaStreamEndCapture(ctx->stream, &g));
    cudaGraphExec_t exec = nullptr;
    cudaError_t e = cudaGraphInstantiate(&exec, g, 0);
    if (e != cudaSuccess) {
        cudaGraphDestroy(g);
        return set_error(DSC_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
    }
    dsc_graph* gr = new dsc_graph();
    gr->graph = g;
    gr->exec = exec;
    *out = gr;
    return DSC_OK;
}
int dsc_graph_launch(dsc_ctx* ctx, dsc_graph* graph, uint32_t rand_seed) {
    CUDA_TRY(cudaSetDevice(ctx->device));
    dsc_set_u32_kernel<<<1, 1, 0, ctx->stream>>>(ctx->step_params, rand_seed);
    CUDA_TRY(cudaGraphLaunch(graph->exec, ctx->stream));
    return DSC_OK;
}
int dsc_graph_destroy(dsc_graph* graph) {
    if (!graph) return DSC_OK;
    if (graph->exec) cudaGraphExecDestroy(graph->exec);
    if (graph->graph) cudaGraphDestroy(graph->graph);
    delete graph;
    return DSC_OK;
}

int dsc_event_create(void** event) {
    cudaEvent_t e;
    CUDA_TRY(cudaEventCreate(&e));
    *event = (void*)e;
    return DSC_OK;
}
int dsc_event_record(dsc_ctx* ctx, void* event) {
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaEventRecord((cudaEvent_t)event, ctx->stream));
    return DSC_OK;
}
int dsc_event_elapsed_ms(void* start, void* end, float* ms) {
    CUDA_TRY(cudaEventSynchronize((cudaEvent_t)end));
    CUDA_TRY(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)end));
    return DSC_OK;
}
int dsc_event_destroy(void* event) {
    if (event) CUDA_TRY(cudaEventDestroy((cudaEvent_t)event));
    return DSC_OK;
}
int dsc_sync(dsc_ctx* ctx) {
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return DSC_OK;
}

int dsc_dp_unique_id(void* out128) {
    int rc = load_nccl();
    if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(out128, &id, 128);
    return DSC_OK;
}
int dsc_dp_init(dsc_ctx* ctx, const void* unique_id128, int world, int rank) {
    if (world < 1 || rank < 0 || rank >= world) return set_error(DSC_ERR_INVALID, "bad data-parallel world/rank %d/%d", world, rank);
    ctx->world = world;
    ctx->rank = rank;
    if (world == 1) return DSC_OK;
    int rc = load_nccl();
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, unique_id128, 128);
    NCCL_TRY(g_nccl.CommInitRank(&ctx->comm, world, id, rank));
    // peer-memory exchange areas for the one-shot all-reduce; any failure here just leaves the NCCL path in charge
    ctx->xchg_ready = false;
    const char* off = getenv("DSC_DP_ONESHOT");
    if (world <= kXchgMaxRanks && g_nccl.AllGather && !(off && atoi(off) == 0)) {
        XchgArea* local = nullptr;
        cudaIpcMemHandle_t* handles = nullptr;
        bool ok = cudaMalloc(&local, sizeof(XchgArea)) == cudaSuccess && cudaMemset(local, 0, sizeof(XchgArea)) == cudaSuccess &&
                  cudaMalloc(&handles, sizeof(cudaIpcMemHandle_t) * world) == cudaSuccess;
        cudaIpcMemHandle_t mine;
        ok = ok && cudaIpcGetMemHandle(&mine, local) == cudaSuccess &&
             cudaMemcpy(handles + rank, &mine, sizeof(mine), cudaMemcpyHostToDevice) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess;
        // every rank must take part in the collective even if its own set-up failed, so the outcome is gathered too
        std::vector<cudaIpcMemHandle_t> all(world);
        if (handles) {
            if (!ok) cudaMemset(handles + rank, 0, sizeof(mine));
            ncclResult_t r = g_nccl.AllGather(handles + rank, handles, sizeof(mine), ncclInt8, ctx->comm, ctx->stream);
            ok = ok && r == ncclSuccess && cudaStreamSynchronize(ctx->stream) == cudaSuccess &&
                 cudaMemcpy(all.data(), handles, sizeof(mine) * world, cudaMemcpyDeviceToHost) == cudaSuccess;
        }
        for (int p = 0; ok && p < world; ++p) {
            if (p == rank) { ctx->xchg.area[p] = local; continue; }
            void* mapped = nullptr;
            ok = cudaIpcOpenMemHandle(&mapped, all[p], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
            ctx->xchg.area[p] = (XchgArea*)mapped;
        }
        if (handles) cudaFree(handles);
        cudaGetLastError();  // a failed attempt must not poison later calls
        // all ranks agree on the outcome: a rank that failed falls back to NCCL, so everyone must
        int* flag = nullptr;
        int agreed = 0;
        if (cudaMalloc(&flag, sizeof(int)) == cudaSuccess) {
            const int mine_ok = ok ? 1 : 0;
            cudaMemcpy(flag, &mine_ok, sizeof(int), cudaMemcpyHostToDevice);
            if (g_nccl.AllReduce(flag, flag, 1, ncclInt32, ncclMin, ctx->comm, ctx->stream) == ncclSuccess && cudaStreamSynchronize(ctx->stream) == cudaSuccess)
                cudaMemcpy(&agreed, flag, sizeof(int), cudaMemcpyDeviceToHost);
            cudaFree(flag);
        }
        ctx->xchg_ready = agreed == 1;
        cudaGetLastError();
    }
    return DSC_OK;
}
int dsc_dp_allreduce_sum_f32(dsc_ctx* ctx, uint64_t id, size_t count) {
    if (ctx->world == 1 || count == 0) return DSC_OK;
    if (!ctx->comm) return set_error(DSC_ERR_NCCL, "data-parallel communicator not initialised");
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (ctx->xchg_ready && count <= kXchgSlotFloats && (id & 15) == 0) {  // small bucket: one CTA over peer-mapped memory (NVLink P2P)
        // one CTA per 16 KB of gradients, at most the 32 flags of a peer's line
        const unsigned ctas = (unsigned)std::min<size_t>(32, std::max<size_t>(1, (count * 4 + 16383) / 16384));
        dsc_oneshot_allreduce_kernel<<<ctas, 1024, 0, ctx->stream>>>((float*)id, (unsigned)count, ctx->xchg, ctx->rank, ctx->world);
        CUDA_TRY(cudaGetLastError());
        return DSC_OK;
    }
    NCCL_TRY(g_nccl.AllReduce((const void*)id, (void*)id, count, ncclFloat32, ncclSum, ctx->comm, ctx->stream));
    return DSC_OK;
}
// The same all-reduce on the context's side stream: it starts once everything issued so far on the context's stream has
// finished and runs beside whatever is issued next; dsc_dp_allreduce_join makes the context's stream wait for it.
// Both are capturable: inside dsc_graph_begin_capture / end_capture they become a fork and a join of the CUDA graph.
int dsc_dp_allreduce_sum_f32_async(dsc_ctx* ctx, uint64_t id, size_t count) {
    if (ctx->world == 1 || count == 0) return DSC_OK;
    if (!ctx->comm) return set_error(DSC_ERR_NCCL, "data-parallel communicator not initialised");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaEventRecord(ctx->comm_fork, ctx->stream));
    CUDA_TRY(cudaStreamWaitEvent(ctx->comm_stream, ctx->comm_fork, 0));
    NCCL_TRY(g_nccl.AllReduce((const void*)id, (void*)id, count, ncclFloat32, ncclSum, ctx->comm, ctx->comm_stream));
    CUDA_TRY(cudaEventRecord(ctx->comm_join, ctx->comm_stream));
    ctx->comm_pending = true;
    return DSC_OK;
}
int dsc_dp_allreduce_join(dsc_ctx* ctx) {
    if (!ctx->comm_pending) return DSC_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->comm_join, 0));
    ctx->comm_pending = false;
    return DSC_OK;
}
int dsc_dp_peer_memory_ready(dsc_ctx* ctx, int* ready) {
    *ready = ctx->xchg_ready ? 1 : 0;
    return DSC_OK;
}
int dsc_dp_world(dsc_ctx* ctx, int* world, int* rank) {
    *world = ctx->world;
    *rank = ctx->rank;
    return DSC_OK;
}

}  // extern "C"

// shared with gemm_tcgen05.cu
extern "C" int dsc_internal_encode_tiled_2d_f32(void* tensor_map, uint64_t base, uint64_t dim0, uint64_t dim1, uint64_t row_stride_bytes,
                                                uint32_t box0, uint32_t box1, int swizzle_atom_32b) {
    int rc = load_driver_api();
    if (rc) return rc;
    cuuint64_t dims[2] = {dim0, dim1};
    cuuint64_t strides[1] = {row_stride_bytes};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t elem[2] = {1, 1};
    CU_TRY(g_drv.TensorMapEncodeTiled((CUtensorMap*)tensor_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, elem,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_atom_32b ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
    return DSC_OK;
}
extern "C" void* dsc_internal_stream(dsc_ctx* ctx) { return (void*)ctx->launch_stream; }

// ---- parallel branches --------------------------------------------------------------------------------------------------
// Kernels of one dependency level are independent by construction (graph.cpp build_clusters).  dsc_branch_fork makes up to
// sixteen side streams wait for everything issued so far on the context's stream; dsc_branch_select routes the following
// dsc_launch / dsc_gemm_tf32* / dsc_fill_u32 / dsc_copy calls to one of them; dsc_branch_join makes the context's stream
// wait for every branch that was used.  All of it is event record / wait, so inside dsc_graph_begin_capture / end_capture
// the level becomes parallel nodes of the CUDA graph (a fork and a join), and the GPU overlaps the small latency-bound
// kernels of a level instead of running them back to back.
int dsc_branch_fork(dsc_ctx* ctx) {
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (ctx->forked) return set_error(DSC_ERR_INVALID, "dsc_branch_fork: already forked");
    CUDA_TRY(cudaEventRecord(ctx->branch_fork, ctx->stream));
    for (int b = 0; b < dsc_ctx::kBranches; ++b) ctx->branch_used[b] = false;
    ctx->forked = true;
    return DSC_OK;
}
int dsc_branch_select(dsc_ctx* ctx, int branch) {
    if (!ctx->forked) return set_error(DSC_ERR_INVALID, "dsc_branch_select outside dsc_branch_fork / dsc_branch_join");
    const int b = ((branch % dsc_ctx::kBranches) + dsc_ctx::kBranches) % dsc_ctx::kBranches;
    if (!ctx->branch_used[b]) {
        CUDA_TRY(cudaStreamWaitEvent(ctx->branch[b], ctx->branch_fork, 0));
        ctx->branch_used[b] = true;
    }
    ctx->launch_stream = ctx->branch[b];
    return DSC_OK;
}
int dsc_branch_join(dsc_ctx* ctx) {
    if (!ctx->forked) return DSC_OK;
    for (int b = 0; b < dsc_ctx::kBranches; ++b) {
        if (!ctx->branch_used[b]) continue;
        CUDA_TRY(cudaEventRecord(ctx->branch_done[b], ctx->branch[b]));
        CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->branch_done[b], 0));
    }
    ctx->launch_stream = ctx->stream;
    ctx->forked = false;
    return DSC_OK;
}
extern "C" int dsc_internal_set_error(int code, const char* msg) { return set_error(code, "%s", msg); }
