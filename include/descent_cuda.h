/* descent_cuda.h -- C ABI of the B200 device layer that replaces descent's `src/device/` and the
 * shader-module half of `src/kernel.rs` (reference: sjb3d/descent, cited as file:line).
 *
 * This is the seam a Rust `src/environment.rs` would bind through `extern "C"`;
 * in this repository the C++ restatement of the frontend (descent_b200/csrc) calls exactly these
 * entry points.  Plain pointers and sizes only.  Every function returns 0 on success and a non-zero
 * DSC_ERR_* code otherwise; `dsc_last_error()` gives the message for the calling thread.  There is no
 * CPU fallback: without a CUDA device `dsc_ctx_create` fails.
 *
 * Threading: a context is used by one thread at a time (the reference is `!Send`, parameter.rs:35).
 * All device work of a context is ordered on one CUDA stream.
 */
#ifndef DESCENT_CUDA_H
#define DESCENT_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSC_OK 0
#define DSC_ERR_CUDA 1        /* a CUDA runtime/driver call failed */
#define DSC_ERR_NVRTC 2       /* JIT compilation failed; the log is in dsc_last_error() */
#define DSC_ERR_INVALID 3     /* bad argument */
#define DSC_ERR_NCCL 4        /* NCCL missing or a collective failed */
#define DSC_ERR_UNSUPPORTED 5 /* shape/layout outside what the kernel implements */

typedef struct dsc_ctx dsc_ctx;       /* device + stream + staging + step parameters */
typedef struct dsc_module dsc_module; /* one NVRTC-compiled cubin */
typedef struct dsc_graph dsc_graph;   /* an instantiated CUDA graph */
typedef void* dsc_kernel;             /* a kernel entry point inside a module */

const char* dsc_last_error(void);
int dsc_device_count(int* count);

/* Context::new (context.rs:34-164) / Drop for Environment (environment.rs:523-528). */
int dsc_ctx_create(int device, dsc_ctx** out);
int dsc_ctx_destroy(dsc_ctx* ctx);
int dsc_ctx_device(dsc_ctx* ctx, int* device);
int dsc_ctx_stream(dsc_ctx* ctx, void** cuda_stream); /* cudaStream_t, for interop only */
int dsc_ctx_sm_count(dsc_ctx* ctx, int* count);

/* BufferHeap::alloc / free (buffer_heap.rs:79-110).  Buffer ids are device addresses; allocation is
 * stream-ordered from the device's memory pool, so alloc/free never synchronise. */
int dsc_alloc(dsc_ctx* ctx, size_t bytes, uint64_t* id);
int dsc_free(dsc_ctx* ctx, uint64_t id);
int dsc_fill_u32(dsc_ctx* ctx, uint64_t id, size_t offset_bytes, uint32_t value, size_t count); /* FillKernel kernel.rs:145-193 */
int dsc_copy(dsc_ctx* ctx, uint64_t dst, uint64_t src, size_t bytes);

/* StagingWriter (staging.rs:85-187): copies n bytes from host to the buffer at `offset`; when
 * `zero_tail_to` > offset+n the bytes up to `zero_tail_to` are zero-filled, which is what dropping a
 * partially written ParameterWriter does (staging.rs:181-187).  `src_is_pinned` != 0 promises the host
 * memory came from dsc_host_alloc (the copy is then asynchronous and the caller must not touch `src`
 * until dsc_sync or a later download); otherwise the bytes are staged through the context's pinned
 * ring and `src` is free on return. */
int dsc_upload(dsc_ctx* ctx, uint64_t id, size_t offset, const void* src, size_t n, size_t zero_tail_to, int src_is_pinned);
/* Double-buffered form of the StagingWriter for the next mini-batch (the reference's staging ring overlaps
 * host writes with GPU work the same way, staging.rs:100-140).  dsc_prefetch copies n bytes of pinned host
 * memory into the device buffer `staging` on the context's copy stream, concurrently with whatever runs on
 * the compute stream; dsc_prefetch_commit makes the compute stream wait for every prefetch issued so far and
 * copies staging -> dst device-side.  A later dsc_prefetch waits for the last commit before overwriting. */
int dsc_prefetch(dsc_ctx* ctx, uint64_t staging, const void* pinned_src, size_t n);
int dsc_prefetch_commit(dsc_ctx* ctx, uint64_t dst, uint64_t staging, size_t n);
/* StagingReader (staging.rs:198-302): blocks until the bytes are on the host. */
int dsc_download(dsc_ctx* ctx, uint64_t id, size_t offset, void* dst, size_t n);
int dsc_host_alloc(size_t bytes, void** out); /* pinned host memory */
int dsc_host_free(void* p);

/* KernelCacheWorker::create_module (kernel.rs:946-1034): NVRTC for sm_100a, cubin loaded into the
 * context's device.  `options` may be NULL; "-fmad=false" is how strict-FP32 per-element code is built. */
int dsc_module_jit(dsc_ctx* ctx, const char* cuda_source, const char* const* options, int num_options, dsc_module** out);
/* Compile only (works without a GPU: used by the build check and the CPU test-suite).  The cubin is
 * malloc'ed; release it with dsc_host_buffer_free. */
int dsc_nvrtc_compile(const char* cuda_source, const char* const* options, int num_options, void** cubin, size_t* bytes);
int dsc_host_buffer_free(void* p);
int dsc_module_load_cubin(dsc_ctx* ctx, const void* cubin, size_t bytes, dsc_module** out);
int dsc_module_get_kernel(dsc_module* module, const char* entry, dsc_kernel* out);
int dsc_module_destroy(dsc_module* module);
int dsc_kernel_set_max_dynamic_smem(dsc_kernel kernel, int bytes);

/* Environment::run_kernel (environment.rs:241-324).  The kernel is launched with the buffers as its
 * first `num_buffers` pointer parameters (inputs then outputs, environment.rs:465-471) followed by one
 * `const unsigned*` pointing at the context's step parameters: [0] = rand_seed of the current run
 * (the reference's push constant, kernel.rs:1000-1004).  No barrier call exists: launches on the
 * context's stream are ordered. */
int dsc_launch(dsc_ctx* ctx, dsc_kernel kernel, uint32_t grid_x, uint32_t grid_y, uint32_t grid_z, uint32_t block_x,
               uint32_t dynamic_smem_bytes, const uint64_t* buffers, int num_buffers);
int dsc_set_rand_seed(dsc_ctx* ctx, uint32_t rand_seed);

/* One queue submit per step (command_buffer.rs:105-123) becomes one CUDA graph launch per step:
 * capture the launches of a step once, then replay. */
int dsc_graph_begin_capture(dsc_ctx* ctx);
int dsc_graph_end_capture(dsc_ctx* ctx, dsc_graph** out);
int dsc_graph_launch(dsc_ctx* ctx, dsc_graph* graph, uint32_t rand_seed);
int dsc_graph_destroy(dsc_graph* graph);

/* TimestampSets (timestamp.rs): events on the context's stream. */
int dsc_event_create(void** event);
int dsc_event_record(dsc_ctx* ctx, void* event);
int dsc_event_elapsed_ms(void* start, void* end, float* ms); /* synchronises on `end` */
int dsc_event_destroy(void* event);
int dsc_sync(dsc_ctx* ctx);

/* Dense GEMM on the tcgen05 tensor cores (replaces MatMulKernel + kernel_matmul.glsl for plain
 * operands, kernel.rs:385-557): C[M,N] (row-major, ldc=N) = A * B with FP32 storage, TF32 operands and
 * FP32 accumulation in TMEM.  a_is_mk != 0: A is row-major [M,K]; else A is stored [K,M] (i.e. the
 * transpose view the reference's dW = a^T * dc uses, array.rs:875).  b_is_kn != 0: B is row-major
 * [K,N]; else B is stored [N,K].  Any M, N, K (edge tiles are zero-filled by TMA and masked in the
 * epilogue); the contiguous extent of A and of B must be a multiple of 4 floats and all three buffers
 * 16-byte aligned, otherwise DSC_ERR_UNSUPPORTED and the caller uses the JIT strict-FP32 path. */
int dsc_gemm_tf32(dsc_ctx* ctx, uint64_t a, uint64_t b, uint64_t c, int64_t m, int64_t n, int64_t k, int a_is_mk, int b_is_kn);
/* The same GEMM with the k range cut into `splits` slices of whole 32-wide k blocks (the reference splits long
 * reductions the same way, r = ceil(K / 1024), op.rs:74, array.rs:515): slice s writes its partial product to
 * c + s * M * N, i.e. `c` is a [splits, M, N] workspace that the caller sums in slice order.  `splits` must leave
 * no slice empty: ceil(kb / ceil(kb / splits)) == splits for kb = ceil(K / 32), else DSC_ERR_UNSUPPORTED. */
int dsc_gemm_tf32_split_k(dsc_ctx* ctx, uint64_t a, uint64_t b, uint64_t c, int64_t m, int64_t n, int64_t k, int a_is_mk, int b_is_kn, int splits);

/* Data parallel (new): one NCCL communicator per context/rank. */
int dsc_dp_unique_id(void* out128);                                  /* ncclGetUniqueId; 128 bytes */
int dsc_dp_init(dsc_ctx* ctx, const void* unique_id128, int world, int rank);
int dsc_dp_allreduce_sum_f32(dsc_ctx* ctx, uint64_t id, size_t count); /* in place, on the context's stream */
/* the same on a side stream behind the work issued so far; the context's stream continues and waits at the join (both capturable) */
int dsc_dp_allreduce_sum_f32_async(dsc_ctx* ctx, uint64_t id, size_t count);
int dsc_dp_allreduce_join(dsc_ctx* ctx);
/* 1 when small buckets (<= 256 KB) are reduced by the one-CTA kernel over CUDA-IPC peer-mapped memory instead of NCCL */
int dsc_dp_peer_memory_ready(dsc_ctx* ctx, int* ready);
int dsc_dp_world(dsc_ctx* ctx, int* world, int* rank);

/* Parallel branches for the independent kernels of one dependency level (the reference issues every cluster with a full
 * barrier in between, environment.rs:326-516): fork makes up to sixteen side streams wait for the context's stream, select
 * routes the following dsc_launch / dsc_gemm_tf32* / dsc_fill_u32 / dsc_copy calls to one of them, join makes the context's
 * stream wait for the branches used.  Capturable: inside dsc_graph_begin_capture they become parallel nodes of the graph. */
int dsc_branch_fork(dsc_ctx* ctx);
int dsc_branch_select(dsc_ctx* ctx, int branch);
int dsc_branch_join(dsc_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* DESCENT_CUDA_H */
