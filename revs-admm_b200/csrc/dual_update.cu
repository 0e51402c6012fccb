// Fused dual-price update, per-home convergence value, primal/dual residuals, the
// convergence test and the target of the next utility step -- one kernel per ADMM
// iteration, nothing leaves the device.
//
// Reference: the tail of the home loop of solve_ADMM (lpsolver.py:280-284)
//     check = P_est[k+1] - P_sch[k+1]
//     G[k+1] = G[k] + kappa/2 * check
//     diff[k+1][h] = ||check|| / T
// plus the objective data of the next Utility (lpsolver.py:202-204), which is the
// projection target  z = (P_est + P_sch)/2 - G/kappa.
//
// Layout: the utility side keeps its arrays time-major [T][Hp] (one contiguous column per
// (feeder,hour) QP), the home side home-major [Hp][T] (one contiguous row per home).  A
// CTA owns 32 homes x T hours and transposes through shared memory, so both sides are
// read and written in full 256-byte runs.  Per-home norms use warp shuffles; the two
// global sums are one atomicAdd per CTA, and the last CTA to finish turns them into the
// residuals and the converged flag.
#include <cuda_bf16.h>

#include "kernels.cuh"

namespace revs {



__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(256, 4) dual_update_kernel(DualParams P) {
    extern __shared__ double tile[];   // [32][T+1]
    __shared__ double s_part[2][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h0 = blockIdx.x * 32;
    const int ldt = P.T + 1;
    const int it = P.iter ? *P.iter : 0;          // read by every CTA before the last one increments it (ticket below)
    const double* __restrict__ p_sch_new = (it & 1) ? P.p_sch_old : P.p_sch_new;
    const double* __restrict__ p_sch_old = (it & 1) ? P.p_sch_new : P.p_sch_old;
    double* __restrict__ diff_k = P.diff_k + (size_t)it * P.Hp;

    for (int t = warp; t < P.T; t += 8) {
        int h = h0 + lane;
        tile[lane * ldt + t] = h < P.Hp ? P.g_t[(size_t)t * P.Hp + h] : 0.0;
    }
    __syncthreads();

    const double hk = 0.5 * P.kappa;
    double blk_p = 0.0, blk_d = 0.0;
    for (int hl = warp; hl < 32; hl += 8) {
        const int h = h0 + hl;
        if (h >= P.Hp) break;
        const size_t base = (size_t)h * P.T;
        double a1 = 0.0, a2 = 0.0;
        for (int t = lane; t < P.T; t += 32) {
            const double e = tile[hl * ldt + t];
            const double sn = p_sch_new[base + t];
            const double so = p_sch_old[base + t];
            const double gm = P.gamma[base + t];
            const double check = __dadd_rn(e, -sn);
            const double g2 = __dadd_rn(gm, __dmul_rn(hk, check));
            P.gamma[base + t] = g2;
            P.p_est[base + t] = e;
            tile[hl * ldt + t] = __dadd_rn(__dmul_rn(__dadd_rn(e, sn), 0.5), -__ddiv_rn(g2, P.kappa));
            a1 = fma(check, check, a1);
            const double ds = sn - so;
            a2 = fma(ds, ds, a2);
        }
        a1 = warp_sum(a1);
        a2 = warp_sum(a2);
        if (lane == 0) diff_k[h] = sqrt(a1) / (double)P.T;
        blk_p += a1;
        blk_d += a2;
    }
    if (lane == 0) { s_part[0][warp] = blk_p; s_part[1][warp] = blk_d; }
    __syncthreads();

    __nv_bfloat16* gbf = reinterpret_cast<__nv_bfloat16*>(P.gbf_next);
    for (int t = warp; t < P.T; t += 8) {
        int h = h0 + lane;
        if (h < P.Hp) {
            const double zv = tile[lane * ldt + t];
            const size_t o = (size_t)t * P.Hp + h;
            P.z_t[o] = zv;
            if (P.g_next) {                    // start of the next utility solve: g = [z]_+ wherever no multiplier is stored
                const double gv = fmax(zv, 0.0);
                P.g_next[o] = gv;
                if (gbf) gbf[o] = __float2bfloat16_rn((float)gv);
            }
        }
    }

    // residuals: per-CTA partial sums, added by the last CTA in a fixed order (no floating-point atomics: the
    // residuals and the convergence decision are reproducible run to run)
    __shared__ int s_last;
    __shared__ double s_fin[2][8];
    if (threadIdx.x == 0) {
        double sp = 0.0, sd = 0.0;
        for (int w = 0; w < 8; ++w) { sp += s_part[0][w]; sd += s_part[1][w]; }
        P.partials[blockIdx.x] = sp;
        P.partials[gridDim.x + blockIdx.x] = sd;
        __threadfence();
        const unsigned done = atomicAdd(&P.res->ticket, 1u);
        s_last = done == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double tp = 0.0, td = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
        tp += __ldcg(P.partials + b);
        td += __ldcg(P.partials + gridDim.x + b);
    }
    tp = warp_sum(tp);
    td = warp_sum(td);
    if (lane == 0) { s_fin[0][warp] = tp; s_fin[1][warp] = td; }
    __syncthreads();
    __shared__ double s_peer[3][kMaxPeers];
    if (P.peer.world > 1) {
        // ---- all-reduce over the GPUs of the box (see PeerReduce): thread r talks to rank r
        const int r = threadIdx.x;
        if (r < P.peer.world) {
            double lp = 0.0, ld = 0.0;
            for (int w = 0; w < 8; ++w) { lp += s_fin[0][w]; ld += s_fin[1][w]; }
            const unsigned long long seq = *P.peer.run_seq + (unsigned long long)it + 1ull;
            const int par = it & 1;
            PeerSlot* out = P.peer.box[r] + par * P.peer.world + P.peer.rank;       // my slot in rank r's mailbox
            out->v[0] = lp;
            out->v[1] = ld;
            out->v[2] = P.count;
            __threadfence_system();
            st_release_sys(&out->seq, seq);
            PeerSlot* in = P.peer.box[P.peer.rank] + par * P.peer.world + r;         // rank r's slot in my mailbox
            const long long t0 = clock64();
            bool ok = true;
            while (ld_acquire_sys(&in->seq) != seq) {
                if (clock64() - t0 > 20000000000ll) { ok = false; break; }           // ~10 s: a peer died
                __nanosleep(200);
            }
            if (!ok) atomicExch(P.peer.timeout, 1);
            s_peer[0][r] = ok ? in->v[0] : 0.0;
            s_peer[1][r] = ok ? in->v[1] : 0.0;
            s_peer[2][r] = ok ? in->v[2] : 0.0;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        tp = 0.0; td = 0.0;
        double count = P.count;
        if (P.peer.world > 1) {
            count = 0.0;
            for (int r = 0; r < P.peer.world; ++r) { tp += s_peer[0][r]; td += s_peer[1][r]; count += s_peer[2][r]; }
        } else {
            for (int w = 0; w < 8; ++w) { tp += s_fin[0][w]; td += s_fin[1][w]; }
        }
        const double r = sqrt(tp / count), sres = P.kappa * sqrt(td / count);
        const int conv = (P.tol > 0.0 && r < P.tol && sres < P.tol) ? 1 : 0;
        P.res->sum_primal = tp;                    // global sums when peers are attached
        P.res->sum_dual = td;
        P.res->count = count;
        P.res->primal = r;
        P.res->dual = sres;
        P.res->converged = conv;
        P.res->ticket = 0u;                        // ready for the next iteration of a captured loop
        if (P.iter) {
            const int k = it + 1;
            *P.iter = k;
            if (P.use_cond) {
                const bool err = (P.err_a && *P.err_a) || (P.err_b && *P.err_b) || (P.peer.world > 1 && *P.peer.timeout);
                cudaGraphSetConditional((cudaGraphConditionalHandle)P.cond_loop, (k < P.iter_max && !conv && !err) ? 1u : 0u);
            }
        }
    }
}

cudaError_t launch_dual_update(const DualParams& P, cudaStream_t stream) {
    size_t smem = (size_t)32 * (P.T + 1) * sizeof(double);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(dual_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dual_update_kernel<<<(P.Hp + 31) / 32, 256, smem, stream>>>(P);
    return cudaGetLastError();
}

}  // namespace revs
