// Voltage screening contraction on the 5th-generation tensor cores: tcgen05.mma with the
// accumulator in TMEM, operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle).
//
// Same product as screen_bf16.cu (V~ = R~ * G~^T, BF16 in, FP32 out, time-major) and the
// same role in the loop; this is the sm_100a-native implementation of it.
//
//   CTA = one 128-row tile of one feeder's sensitivity block x all (<= 96) hours.
//   warp 0 (one lane)  TMA producer: per 64-wide k block one box of A (128 x 64 bf16) from
//                      the feeder's tensor map and one box of B (96 x 64) from the map of
//                      the time-major schedule, into a 3-stage ring (28 KB per stage, 2 CTAs per SM)
//   warp 1 (one lane)  MMA issuer: 4 x tcgen05.mma.kind::f16 (M=128, N=96, K=16) per stage,
//                      tcgen05.commit frees the stage / signals the epilogue
//   warps 2..5         epilogue: tcgen05.ld 32 lanes x 96 columns each (thread = row),
//                      coalesced time-major FP32 stores
// Out-of-range rows / k / hours are zero-filled by TMA, so ragged feeders need no special
// cases.  Tiles whose feeder has no running column exit before touching anything.
#include <cuda.h>
#include <cuda_bf16.h>

#include "kernels.cuh"

namespace revs {

namespace {

constexpr int kBM = 128, kBN = 96, kBK = 64, kStagesT = 3;   // 87 KB -> two CTAs per SM
constexpr int kUmmaK = 16;
constexpr int kThreadsT = 192;                       // 6 warps
constexpr uint32_t kABytes = kBM * kBK * 2, kBBytes = kBN * kBK * 2;
constexpr uint32_t kStageBytes = kABytes + kBBytes;   // 28672, multiple of 1024
constexpr uint32_t kTmemCols = 128;                   // power of two >= 96
constexpr size_t kSmemT = 1024 + (size_t)kStagesT * kStageBytes + 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const void* tmap, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(smem_u32(dst)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    // K-major, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor, sm100 version)
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(1024 >> 4) << 32;      // stride byte offset
    d |= (uint64_t)1 << 46;                // descriptor version
    d |= (uint64_t)2 << 61;                // LayoutType::SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_c),
        "l"(da), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}

// idesc: c=F32 (1<<4), a=BF16 (1<<7), b=BF16 (1<<10), K-major both, N>>3 at bit 17, M>>4 at bit 24
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);

__global__ void __launch_bounds__(kThreadsT, 2)
screen_tc5_kernel(const ScreenProblem* __restrict__ problems, const ContractTile* __restrict__ tiles,
                  const CUtensorMap* __restrict__ maps_a, const CUtensorMap* __restrict__ map_b,
                  const int* __restrict__ b_col0, int T, double thr) {
    extern __shared__ unsigned char smem_raw[];
    const ContractTile tile = tiles[blockIdx.x];
    const ScreenProblem pb = problems[tile.problem];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (pb.col_status) {       // nothing to do for a feeder whose columns have all converged
        bool on = false;
        for (int t = threadIdx.x; t < T; t += kThreadsT) on |= pb.col_status[t] == 0;
        if (__syncthreads_or(on) == 0) return;
    }

    unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* stage_mem = base;                                   // [stages][A | B], 1024-aligned
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(base + (size_t)kStagesT * kStageBytes);
    uint64_t* empty_bar = full_bar + kStagesT;
    uint64_t* tmem_full = empty_bar + kStagesT;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const CUtensorMap* map_a = maps_a + tile.problem;
    const int nk = (pb.K + kBK - 1) / kBK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStagesT; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 2) {           // TMEM allocation by one full warp
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];\n" ::"l"(map_a) : "memory");
            asm volatile("prefetch.tensormap [%0];\n" ::"l"(map_b) : "memory");
            const int bcol = b_col0[tile.problem];
            for (int kb = 0; kb < nk; ++kb) {
                const int s = kb % kStagesT;
                mbar_wait(&empty_bar[s], ((kb / kStagesT) & 1) ^ 1);
                mbar_expect_tx(&full_bar[s], kStageBytes);
                unsigned char* a = stage_mem + (size_t)s * kStageBytes;
                tma_load_2d(a, map_a, &full_bar[s], kb * kBK, tile.row0);
                tma_load_2d(a + kABytes, map_b, &full_bar[s], bcol + kb * kBK, 0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int kb = 0; kb < nk; ++kb) {
                const int s = kb % kStagesT;
                mbar_wait(&full_bar[s], (kb / kStagesT) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint32_t a_addr = smem_u32(stage_mem + (size_t)s * kStageBytes);
                const uint64_t da = make_desc_sw128(a_addr), db = make_desc_sw128(a_addr + kABytes);
#pragma unroll
                for (int k = 0; k < kBK / kUmmaK; ++k)   // +32 bytes per K=16 slice inside the 128-byte swizzle row
                    umma_bf16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), kIdesc, (kb | k) ? 1u : 0u);
                umma_commit(&empty_bar[s]);              // stage reusable once these MMAs have read it
            }
            umma_commit(tmem_full);                      // accumulator complete
        }
    } else {
        // epilogue warps 2..5: warp w may touch TMEM lanes 32*(w%4) .. +31
        mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const int q = warp & 3;
        const int row = tile.row0 + q * 32 + lane;
#pragma unroll
        for (int c0 = 0; c0 < kBN; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            if (row < pb.M) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int t = c0 + j;
                    if (t < T) {
                        const float a = __uint_as_float(v[j]);
                        pb.out[(size_t)t * pb.ldo + row] = a;
                        if (pb.col_cand && (double)a > thr) pb.col_cand[t] = 1;   // the column needs the QP kernel
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

}  // namespace

int screen_tc5_tile_rows() { return kBM; }
size_t screen_tc5_map_bytes() { return sizeof(CUtensorMap); }

// Encode one 2-D BF16 tensor map (rows x cols, row pitch in elements) with a (box_rows x 64) box.
cudaError_t screen_tc5_encode(void* host_map, const void* gptr, uint64_t rows, uint64_t cols, uint64_t pitch_elems,
                              uint32_t box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return cudaErrorNotSupported;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {pitch_elems * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(reinterpret_cast<CUtensorMap*>(host_map), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(gptr), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

uint32_t screen_tc5_box_rows_a() { return kBM; }
uint32_t screen_tc5_box_rows_b() { return kBN; }

cudaError_t screen_tc5_prepare() {
    static std::atomic<unsigned long long> attr_devices{0};
    if (first_use_on_device(attr_devices))
        return cudaFuncSetAttribute(screen_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemT);
    return cudaSuccess;
}

cudaError_t launch_screen_tc5(const ScreenProblem* d_problems, const ContractTile* d_tiles, int n_tiles, const void* d_maps_a,
                              const void* d_map_b, const int* d_b_col0, int T, double thr, cudaStream_t stream) {
    if (n_tiles == 0) return cudaSuccess;
    if (T > kBN) return cudaErrorInvalidValue;
    cudaError_t pe = screen_tc5_prepare();
    if (pe != cudaSuccess) return pe;
    screen_tc5_kernel<<<n_tiles, kThreadsT, kSmemT, stream>>>(d_problems, d_tiles, reinterpret_cast<const CUtensorMap*>(d_maps_a),
                                                            reinterpret_cast<const CUtensorMap*>(d_map_b), d_b_col0, T, thr);
    return cudaGetLastError();
}

}  // namespace revs
