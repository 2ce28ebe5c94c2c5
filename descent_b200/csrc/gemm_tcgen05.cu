// Dense TF32 GEMM on the Blackwell tensor cores: TMA (cp.async.bulk.tensor) -> 128B-swizzled shared
// tiles -> tcgen05.mma.kind::tf32 with FP32 accumulators in TMEM -> tcgen05.ld epilogue.
// Replaces the reference's 16x16x16 shared-memory SIMT shader (src/kernel_matmul.glsl) for plain
// (untransformed or transposed) operands; FP32 storage, operands read as TF32, FP32 accumulation.
//
// One persistent CTA per SM, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (one elected
// lane), warp 2 = TMEM allocator, warps 4..7 = epilogue (each owns the 32 TMEM lanes of its quadrant).
// Pipelines: smem full/empty ring (TMA <-> MMA) and a 2-deep TMEM full/empty ring (MMA <-> epilogue),
// so the epilogue of tile i overlaps the main loop of tile i+1.
//
// Operand layouts (all four combinations): K-contiguous operands land as the canonical K-major
// SWIZZLE_128B layout (rows of 32 floats); MN-contiguous operands land as MN-major SWIZZLE_128B
// (rows of 32 floats along M/N, one row per k), and the instruction descriptor's major bits say which.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#include "../../include/descent_cuda.h"

extern "C" int dsc_internal_encode_tiled_2d_f32(void* tensor_map, uint64_t base, uint64_t dim0, uint64_t dim1, uint64_t row_stride_bytes,
                                                uint32_t box0, uint32_t box1, int swizzle_atom_32b);
extern "C" void* dsc_internal_stream(dsc_ctx* ctx);
extern "C" int dsc_internal_set_error(int code, const char* msg);

namespace {

constexpr int BM = 128;        // UMMA M (cta_group::1)
constexpr int BK = 32;         // floats per stage along K = one 128-byte swizzle span
constexpr int UMMA_K = 8;      // tf32: 32 bytes of K per instruction
constexpr int NUM_THREADS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y)
                 : "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): address, LBO, SBO in 16-byte units,
// version 1 (Blackwell), layout type in bits [61,64): SWIZZLE_128B = 2 (K-major operands) or
// SWIZZLE_128B_BASE32B = 1, the only swizzled layout tf32 accepts for MN-major operands
// (cutlass/gemm/collective/builders/sm100_common.inl:92): 32-byte chunks XOR-ed with the row index mod 4.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= 1ull << 46;
    d |= (uint64_t)layout_type << 61;
    return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, majors, N>>3, M>>4
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, bool a_mn_major, bool b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(pred));
    return pred != 0;
}

template <int BN, int STAGES>
struct SharedStorage {
    alignas(1024) float a[STAGES][BM * BK];
    alignas(1024) float b[STAGES][BN * BK];
    alignas(8) uint64_t full[STAGES];
    uint64_t empty[STAGES];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
    // epilogue: each warp turns its 32 rows x 32 columns (one row per lane out of TMEM) into row-contiguous stores;
    // 36-float rows keep the quarter-warp phases of the 128-bit accesses on distinct banks
    alignas(16) float stage[4][32][36];
};

template <int BN, int STAGES, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, float* __restrict__ c, int m, int n, int k, int splits) {
    extern __shared__ uint8_t smem_raw[];
    using Storage = SharedStorage<BN, STAGES>;
    Storage& s = *reinterpret_cast<Storage*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    // (shuffles: warp-uniform values the compiler can keep in uniform registers, which TMA and tcgen05.mma operands need)
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x / 32, 0), lane = threadIdx.x % 32;
    const int tiles_m = (m + BM - 1) / BM, tiles_n = (n + BN - 1) / BN, num_tiles = tiles_m * tiles_n, k_blocks = (k + BK - 1) / BK;  // TMA zero-fills out-of-range boxes
    // split K: work item = (split, tile); split s accumulates k-blocks [s * kbs, (s + 1) * kbs) into c + s * m * n
    const int kbs = (k_blocks + splits - 1) / splits, num_items = num_tiles * splits;
    constexpr uint32_t STAGE_BYTES = (BM + BN) * BK * 4;
    constexpr uint32_t TMEM_COLS = 2 * BN;  // two accumulator stages

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&s.full[i], 1);
            mbar_init(&s.empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s.tmem_full[i], 1);
            mbar_init(&s.tmem_empty[i], 4);  // one arrive per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, s.tmem_base, 0);

    // the single producer / issuer threads are chosen with elect.sync: under `lane == 0` the compiler must assume any subset
    // of lanes and wraps every TMA / MMA in an elect-and-broadcast loop (BRA.U.ANY per instruction in SASS)
    if (warp == 0) { if (elect_one()) {
        // ===== TMA producer =====
        uint32_t stage = 0, phase = 0;
        for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
            const int tile = item % num_tiles, kb_lo = (item / num_tiles) * kbs, kb_hi = min(k_blocks, kb_lo + kbs);
            const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
            for (int kb = kb_lo; kb < kb_hi; ++kb) {
                mbar_wait(&s.empty[stage], phase ^ 1);
                mbar_expect_tx(&s.full[stage], STAGE_BYTES);
                const int k0 = kb * BK;
                if (A_MN) {  // A stored [K, M]: boxes of 32 (m) x BK (k), one per 32 rows of the tile
                    for (int j = 0; j < BM / 32; ++j) tma_load_2d(&map_a, &s.full[stage], &s.a[stage][j * 32 * BK], m0 + j * 32, k0);
                } else {     // A stored [M, K]: one box of BK (k) x BM (m)
                    tma_load_2d(&map_a, &s.full[stage], &s.a[stage][0], k0, m0);
                }
                if (B_MN) {  // B stored [K, N]
                    for (int j = 0; j < BN / 32; ++j) tma_load_2d(&map_b, &s.full[stage], &s.b[stage][j * 32 * BK], n0 + j * 32, k0);
                } else {     // B stored [N, K]
                    tma_load_2d(&map_b, &s.full[stage], &s.b[stage][0], k0, n0);
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } } else if (warp == 1) { if (elect_one()) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = make_idesc(BM, BN, A_MN, B_MN);
        // K-major SW128: 8-row groups 1024 B apart (SBO), LBO unused; step 32 B per UMMA_K inside the 128 B row.
        // MN-major SW128_BASE32B: 32-wide M/N blocks BK*128 B apart (LBO), 4-k groups 512 B apart (SBO); one UMMA_K
        // (8 k) is two such groups, so the start address steps 1024 B.
        constexpr uint32_t A_LBO = A_MN ? BK * 128 : 16, B_LBO = B_MN ? BK * 128 : 16;
        constexpr uint32_t A_SBO = A_MN ? 512 : 1024, B_SBO = B_MN ? 512 : 1024;
        constexpr uint32_t A_TYPE = A_MN ? 1 : 2, B_TYPE = B_MN ? 1 : 2;
        constexpr uint32_t A_KSTEP = A_MN ? 1024 : 32, B_KSTEP = B_MN ? 1024 : 32;
        uint32_t stage = 0, phase = 0, acc_stage = 0, acc_phase = 0;
        for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
            const int kb_lo = (item / num_tiles) * kbs, kb_hi = min(k_blocks, kb_lo + kbs);
            mbar_wait(&s.tmem_empty[acc_stage], acc_phase ^ 1);
            tcgen05_fence_after();
            const uint32_t tmem_d = tmem_base + acc_stage * BN;
            for (int kb = kb_lo; kb < kb_hi; ++kb) {
                mbar_wait(&s.full[stage], phase);
                tcgen05_fence_after();
                const uint32_t a_addr = smem_u32(&s.a[stage][0]), b_addr = smem_u32(&s.b[stage][0]);
#pragma unroll
                for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                    const uint64_t adesc = make_smem_desc(a_addr + kk * A_KSTEP, A_LBO, A_SBO, A_TYPE);
                    const uint64_t bdesc = make_smem_desc(b_addr + kk * B_KSTEP, B_LBO, B_SBO, B_TYPE);
                    umma_tf32(tmem_d, adesc, bdesc, idesc, (kb != kb_lo || kk != 0) ? 1u : 0u);
                }
                tcgen05_commit(&s.empty[stage]);  // frees the smem slot when these MMAs retire
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            tcgen05_commit(&s.tmem_full[acc_stage]);  // accumulator complete
            if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
        }
    } } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> global =====
        const int quad = warp % 4;  // TMEM lanes [32*quad, 32*quad+32)
        uint32_t acc_stage = 0, acc_phase = 0;
        for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
            const int tile = item % num_tiles, split = item / num_tiles;
            const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
            mbar_wait(&s.tmem_full[acc_stage], acc_phase);
            tcgen05_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < BN && n0 + c0 < n; c0 += 32) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc_stage * BN + c0;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                      "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                      "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                      "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                // a lane holds 32 consecutive columns of ONE row: stored as is, every instruction of the warp would touch 32
                // rows with 16 bytes each (half-filled sectors; measured 43 us for the 51 MB product of a K = 128 GEMM).
                // Through the warp's staging tile each instruction writes whole 128-byte row segments instead.
                float (*tile)[36] = s.stage[quad];
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(&tile[lane][j]) =
                        make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                __syncwarp();
                const int row0 = m0 + quad * 32;
                float* base = c + ((size_t)split * m + row0) * n + n0 + c0;
                if ((n & 3) == 0) {
                    const int cpos = (lane & 7) * 4;
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int r = it * 4 + (lane >> 3);
                        if (row0 + r < m && n0 + c0 + cpos < n)
                            *reinterpret_cast<float4*>(base + (size_t)r * n + cpos) = *reinterpret_cast<const float4*>(&tile[r][cpos]);
                    }
                } else {
#pragma unroll 4
                    for (int r = 0; r < 32; ++r)
                        if (row0 + r < m && n0 + c0 + lane < n) base[(size_t)r * n + lane] = tile[r][lane];
                }
                __syncwarp();
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s.tmem_empty[acc_stage]);
            if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

template <int BN, int STAGES, bool A_MN, bool B_MN>
int launch(dsc_ctx* ctx, const CUtensorMap& ma, const CUtensorMap& mb, float* c, int m, int n, int k, int sm_count, int splits) {
    auto kernel = gemm_tf32_kernel<BN, STAGES, A_MN, B_MN>;
    const int smem = (int)sizeof(SharedStorage<BN, STAGES>) + 1024;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return dsc_internal_set_error(DSC_ERR_CUDA, cudaGetErrorString(e));
    const int tiles = ((m + BM - 1) / BM) * ((n + BN - 1) / BN) * splits;
    const int grid = tiles < sm_count ? tiles : sm_count;
    kernel<<<grid, NUM_THREADS, smem, (cudaStream_t)dsc_internal_stream(ctx)>>>(ma, mb, c, m, n, k, splits);
    e = cudaGetLastError();
    if (e != cudaSuccess) return dsc_internal_set_error(DSC_ERR_CUDA, cudaGetErrorString(e));
    return DSC_OK;
}

}  // namespace

extern "C" int dsc_gemm_tf32(dsc_ctx* ctx, uint64_t a, uint64_t b, uint64_t c, int64_t m, int64_t n, int64_t k, int a_is_mk, int b_is_kn) {
    return dsc_gemm_tf32_split_k(ctx, a, b, c, m, n, k, a_is_mk, b_is_kn, 1);
}

extern "C" int dsc_gemm_tf32_split_k(dsc_ctx* ctx, uint64_t a, uint64_t b, uint64_t c, int64_t m, int64_t n, int64_t k, int a_is_mk, int b_is_kn, int splits) {
    if (splits < 1 || (int64_t)(splits - 1) * 32 >= k) return dsc_internal_set_error(DSC_ERR_UNSUPPORTED, "dsc_gemm_tf32_split_k: every split needs at least one 32-wide k block");
    {   // every split must own at least one k block: shrink to the number of non-empty slices
        const int64_t k_blocks = (k + 31) / 32, kbs = (k_blocks + splits - 1) / splits;
        if ((k_blocks + kbs - 1) / kbs != splits) return dsc_internal_set_error(DSC_ERR_UNSUPPORTED, "dsc_gemm_tf32_split_k: splits must divide the k blocks into non-empty slices");
    }
    if (m <= 0 || n <= 0 || k <= 0 || m > INT32_MAX || n > INT32_MAX || k > INT32_MAX)
        return dsc_internal_set_error(DSC_ERR_UNSUPPORTED, "dsc_gemm_tf32: bad shape");
    // TMA needs 16-byte row pitches: the contiguous extent of each operand must be a multiple of 4 floats
    if ((a_is_mk ? k : m) % 4 != 0 || (b_is_kn ? n : k) % 4 != 0)
        return dsc_internal_set_error(DSC_ERR_UNSUPPORTED, "dsc_gemm_tf32 needs the contiguous extent of A and of B to be a multiple of 4");
    if ((a | b | c) & 15) return dsc_internal_set_error(DSC_ERR_UNSUPPORTED, "dsc_gemm_tf32 needs 16-byte aligned operands");
    int device = 0, sm_count = 148;
    dsc_ctx_device(ctx, &device);
    dsc_ctx_sm_count(ctx, &sm_count);
    cudaSetDevice(device);
    const bool a_mn = a_is_mk == 0, b_mn = b_is_kn != 0;
    alignas(64) CUtensorMap ma, mb;
    int rc;
    // tensor maps: dim0 is the contiguous dimension; boxes are 32 floats (128 B, the swizzle span) wide
    if (a_mn) rc = dsc_internal_encode_tiled_2d_f32(&ma, a, (uint64_t)m, (uint64_t)k, (uint64_t)m * 4, 32, BK, 1);
    else rc = dsc_internal_encode_tiled_2d_f32(&ma, a, (uint64_t)k, (uint64_t)m, (uint64_t)k * 4, BK, BM, 0);
    if (rc) return rc;
    // 128x256 tiles (4 stages, all 512 TMEM columns) halve the B-operand shared-memory traffic per MMA: with
    // 128x128 tiles a tf32 MMA reads 8 KB per 64 cycles = the whole 128 B/clk of shared-memory bandwidth.
    const bool wide = n >= 1024 && ((int64_t)((m + BM - 1) / BM) * ((n + 255) / 256)) >= sm_count;
    const uint32_t bn = wide ? 256 : 128;
    if (b_mn) rc = dsc_internal_encode_tiled_2d_f32(&mb, b, (uint64_t)n, (uint64_t)k, (uint64_t)n * 4, 32, BK, 1);
    else rc = dsc_internal_encode_tiled_2d_f32(&mb, b, (uint64_t)k, (uint64_t)n, (uint64_t)k * 4, BK, bn, 0);
    if (rc) return rc;
    float* cp = (float*)c;
    const int mi = (int)m, ni = (int)n, ki = (int)k;
    if (wide) {
        if (!a_mn && !b_mn) return launch<256, 4, false, false>(ctx, ma, mb, cp, mi, ni, ki, sm_count, splits);
        if (!a_mn && b_mn) return launch<256, 4, false, true>(ctx, ma, mb, cp, mi, ni, ki, sm_count, splits);
        if (a_mn && !b_mn) return launch<256, 4, true, false>(ctx, ma, mb, cp, mi, ni, ki, sm_count, splits);
        return launch<256, 4, true, true>(ctx, ma, mb, cp, mi, ni, ki, sm_count, splits);
    }
    if (!a_mn && !b_mn) return launch<128, 6, false, false>(ctx, ma, mb, cp, mi, ni, ki, sm_count, splits);
    if (!a_mn && b_mn) return launch<128, 6, false, true>(ctx, ma, mb, cp, mi, ni, ki, sm_count, splits);
    if (a_mn && !b_mn) return launch<128, 6, true, false>(ctx, ma, mb, cp, mi, ni, ki, sm_count, splits);
    return launch<128, 6, true, true>(ctx, ma, mb, cp, mi, ni, ki, sm_count, splits);
}
