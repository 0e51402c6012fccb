// Sensitivity contraction  C = A * B^T  in FP64 on the tensor cores (DMMA m8n8k4).
//
// This is the LinDistFlow reliability check of the reference: R_res @ g in
// Utility.network (lpsolver.py:183-194) and R @ P / A_inv @ P in drawing.py:29-78.
// A is a feeder's dense sensitivity block (voltage: R, flow: subtree incidence), B^T is
// the schedule in time-major layout [T][homes], so the contraction runs over the homes
// and all T hours of a feeder share one pass over the sensitivity block.
//
// One CTA owns a BM x BN tile of C; K is streamed through a 3-stage cp.async ring in
// BK=16 slices.  Shared tiles have a row stride of BK+4 doubles so that the DMMA operand
// reads (lane l -> row l/4, k l%4) touch each 8-byte bank at most twice per 256-byte
// request, which is the floor for 64-bit shared loads.  FP64 has no tcgen05 kind, so the
// tensor path for double precision on sm_100a is mma.sync.m8n8k4.f64.
#include "kernels.cuh"

namespace revs {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

constexpr int kBK = 16;
constexpr int kLds = kBK + 4;
constexpr int kStages = 3;

template <int BM, int BN, int WARPS_M, int WARPS_N>
struct ContractCfg {
    static constexpr int kThreads = 32 * WARPS_M * WARPS_N;
    static constexpr int kWM = BM / WARPS_M;        // rows per warp
    static constexpr int kWN = BN / WARPS_N;        // cols per warp
    static constexpr int kMT = kWM / 8;
    static constexpr int kNT = kWN / 8;
    static constexpr size_t kSmem = (size_t)kStages * (BM + BN) * kLds * sizeof(double);
    static_assert(kWM % 8 == 0 && kWN % 8 == 0, "warp tile must be a multiple of 8x8");
};

template <int BM, int BN, int WARPS_M, int WARPS_N>
__global__ void __launch_bounds__(32 * WARPS_M * WARPS_N)
contract_f64_kernel(const ContractProblem* __restrict__ problems,
                    const ContractTile* __restrict__ tiles, int T, int mode, double v2) {
    using Cfg = ContractCfg<BM, BN, WARPS_M, WARPS_N>;
    extern __shared__ __align__(16) double smem[];
    double* sA = smem;                                   // [stages][BM][kLds]
    double* sB = smem + (size_t)kStages * BM * kLds;     // [stages][BN][kLds]

    const ContractTile tile = tiles[blockIdx.x];
    const ContractProblem pb = problems[tile.problem];
    const int m0 = tile.row0;
    const int n0 = blockIdx.y * BN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp % WARPS_M) * Cfg::kWM;
    const int wn = (warp / WARPS_M) * Cfg::kWN;
    const int nk = pb.K / kBK;

    // columns whose QP has converged keep their voltages: skip their 8-wide groups, and the
    // whole tile when no column of this feeder is still running
    unsigned act = 0xffffffffu;
    if (pb.col_status) {
        act = 0u;
#pragma unroll
        for (int j = 0; j < Cfg::kNT; ++j) {
            const int colj = n0 + wn + j * 8 + (lane & 7);
            const bool on = colj < T && pb.col_status[colj] == 0;
            if (__any_sync(0xffffffffu, on)) act |= 1u << j;
        }
        if (__syncthreads_or(act != 0u) == 0) return;
    }

    auto load_stage = [&](int stage, int kt) {
        const int k0 = kt * kBK;
        double* a = sA + (size_t)stage * BM * kLds;
        double* b = sB + (size_t)stage * BN * kLds;
        for (int c = tid; c < BM * (kBK / 2); c += Cfg::kThreads) {
            int r = c / (kBK / 2), q = c % (kBK / 2);
            bool ok = (m0 + r) < pb.M;
            const double* src = pb.A + (size_t)(ok ? m0 + r : 0) * pb.lda + k0 + 2 * q;
            cp_async16(a + r * kLds + 2 * q, src, ok);
        }
        for (int c = tid; c < BN * (kBK / 2); c += Cfg::kThreads) {
            int r = c / (kBK / 2), q = c % (kBK / 2);
            bool ok = (n0 + r) < T;
            const double* src = pb.Bt + (size_t)(ok ? n0 + r : 0) * pb.ldb + k0 + 2 * q;
            cp_async16(b + r * kLds + 2 * q, src, ok);
        }
    };

    double acc[Cfg::kMT][Cfg::kNT][2];
#pragma unroll
    for (int i = 0; i < Cfg::kMT; ++i)
#pragma unroll
        for (int j = 0; j < Cfg::kNT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < kStages - 1; ++s) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }

    const int fr = lane >> 2, fk = lane & 3;
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<kStages - 2>();
        __syncthreads();
        {   // refill the slot freed in the previous iteration
            int nxt = kt + kStages - 1;
            if (nxt < nk) load_stage(nxt % kStages, nxt);
            cp_async_commit();
        }
        const double* a = sA + (size_t)(kt % kStages) * BM * kLds + (wm + fr) * kLds + fk;
        const double* b = sB + (size_t)(kt % kStages) * BN * kLds + (wn + fr) * kLds + fk;
#pragma unroll
        for (int kk = 0; kk < kBK; kk += 4) {
            double af[Cfg::kMT], bf[Cfg::kNT];
#pragma unroll
            for (int i = 0; i < Cfg::kMT; ++i) af[i] = a[i * 8 * kLds + kk];
#pragma unroll
            for (int j = 0; j < Cfg::kNT; ++j) bf[j] = b[j * 8 * kLds + kk];
#pragma unroll
            for (int i = 0; i < Cfg::kMT; ++i)
#pragma unroll
                for (int j = 0; j < Cfg::kNT; ++j)
                    if (act & (1u << j)) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: lane holds C[row = fr][col = 2*fk, 2*fk+1] of every 8x8 tile
#pragma unroll
    for (int i = 0; i < Cfg::kMT; ++i) {
        const int row = m0 + wm + i * 8 + fr;
        if (row >= pb.M) continue;
        const double sc = (mode == kOutScaled && pb.scale) ? pb.scale[row] : 1.0;
#pragma unroll
        for (int j = 0; j < Cfg::kNT; ++j) {
            if (!(act & (1u << j))) continue;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int col = n0 + wn + j * 8 + 2 * fk + e;
                if (col >= T) continue;
                double v = acc[i][j][e];
                if (mode == kOutTimeMajor) {
                    pb.out[(size_t)col * pb.ldo + row] = v;
                } else {
                    if (mode == kOutVoltage) v = sqrt(v2 - v);
                    else if (mode == kOutScaled) v = sc * v;
                    if (pb.gather) {
                        // fused all-gather: the element goes to every rank's buffer over NVLink
                        const GatherDev& G = *pb.gather;
                        const size_t at = (size_t)(G.row_base + row) * G.ldo + col;
                        for (int r = 0; r < G.world; ++r) G.out[r][at] = v;
                    } else {
                        pb.out[(size_t)row * pb.ldo + col] = v;
                    }
                }
            }
        }
    }
    if (pb.gather) {
        // every thread makes its peer stores visible system-wide, the last CTA of the launch raises the arrival flags
        const GatherDev& G = *pb.gather;
        __shared__ unsigned s_last;
        __threadfence_system();
        __syncthreads();
        if (tid == 0) s_last = atomicAdd(G.ticket, 1u) == gridDim.x * gridDim.y - 1 ? 1u : 0u;
        __syncthreads();
        if (s_last && tid < G.world) {
            __threadfence_system();
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(G.flag[tid] + G.rank), "l"(G.seq) : "memory");
        }
    }
}

// waits until every rank's arrival flag in THIS rank's gather buffer carries `seq` (one warp; ~10 s timeout)
__global__ void gather_wait_kernel(const unsigned long long* flags, int world, unsigned long long seq, int* timeout) {
    const int r = threadIdx.x;
    if (r >= world) return;
    const long long t0 = clock64();
    for (;;) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + r) : "memory");
        if (v >= seq) break;
        if (clock64() - t0 > 20000000000ll) { atomicExch(timeout, 1); break; }
    }
}

cudaError_t launch_gather_wait(const unsigned long long* flags, int world, unsigned long long seq, int* timeout, cudaStream_t s) {
    gather_wait_kernel<<<1, 32, 0, s>>>(flags, world, seq, timeout);
    return cudaGetLastError();
}

// Tile shape by horizon: T<=24 (the reference's hourly day) is HBM-bound -> tall, narrow
// tiles; up to 96 quarter-hours keeps all hours of a feeder in one pass over R.
int contract_tile_rows(int T) { return T <= 24 ? 128 : (T <= 48 ? 128 : 64); }

cudaError_t launch_contract(const ContractProblem* d_problems, const ContractTile* d_tiles,
                            int n_tiles, int T, int mode, double v2, cudaStream_t stream) {
    if (n_tiles == 0) return cudaSuccess;
    if (T <= 24) {
        using Cfg = ContractCfg<128, 24, 4, 1>;
        auto k = contract_f64_kernel<128, 24, 4, 1>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem);
        k<<<dim3(n_tiles, 1), Cfg::kThreads, Cfg::kSmem, stream>>>(d_problems, d_tiles, T, mode, v2);
    } else if (T <= 48) {
        using Cfg = ContractCfg<128, 48, 4, 1>;
        auto k = contract_f64_kernel<128, 48, 4, 1>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem);
        k<<<dim3(n_tiles, 1), Cfg::kThreads, Cfg::kSmem, stream>>>(d_problems, d_tiles, T, mode, v2);
    } else {
        using Cfg = ContractCfg<64, 96, 2, 2>;
        auto k = contract_f64_kernel<64, 96, 2, 2>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem);
        k<<<dim3(n_tiles, (T + 95) / 96), Cfg::kThreads, Cfg::kSmem, stream>>>(d_problems, d_tiles, T,
                                                                              mode, v2);
    }
    return cudaGetLastError();
}

}  // namespace revs
