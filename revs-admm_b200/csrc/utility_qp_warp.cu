// Utility QP, small columns: ONE WARP per (zone, hour) column, no block barriers.
//
// After a feeder is split into its voltage zones (feeder.py:split_zones) most columns have
// n ~ 10^2 residences and a handful of binding voltage rows.  For those the CTA-wide kernel of
// utility_qp.cu spends its time in barriers and in latency it cannot hide; here a column
// lives in the registers of a single warp:
//   lanes own the homes   j = lane + 32 k  (k < NJ, n <= 32 NJ): z_j, g_j, trial g_j
//   lanes own the rows    a < m <= 16 of the working set: idx_a, lam_a, grad_a, H[a][.]
// and every step of the algorithm of utility_qp.cu (same fixed point, same tolerances) is a
// few shuffles: Hessian = rank-1 updates over the homes of F (row a of H accumulates in lane
// a's registers; later pieces apply signed updates for the homes that crossed g = 0),
// primal-dual active-set loop with a 16x16 Cholesky in shared memory, segment line search.
// Anything unusual -- working set outgrowing 16 rows, active-set guesses cycling, line search
// failing -- hands the column, untouched, to the CTA kernel of the next class.
//
// qp_init_kernel (also one warp per column, any size) starts a utility solve: working set =
// support of the stored multipliers, class by its size, g = [z - R lam]_+ for the new target.
#include <cuda_bf16.h>

#include "kernels.cuh"

namespace revs {

namespace {

constexpr int kWarpsPerCta = 1;              // one column per CTA: a finished column frees its slot at once
constexpr int kHW = kWW + 1;                 // leading dim of the per-warp 16x16 matrices
constexpr double kArcMinW = 9.5367431640625e-07;
constexpr int kPdasMaxW = 40;
constexpr double kHessShiftW = 1e-12;
constexpr int kAddMaxW = 8;                  // violated rows admitted per round by the warp kernel

struct WarpSmem {
    double H[kWW * kHW];      // model Hessian, full symmetric
    double L[kWW * kHW];      // Cholesky factor of the active sub-matrix
};

__device__ __forceinline__ double warp_bcast(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

}  // namespace

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) qp_init_kernel(QpParams P, int max_warp_n) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= P.ncols) return;
    const int f = c / P.T, t = c % P.T;
    const FeederDev fd = P.feeders[f];
    const int n = fd.n, ld = fd.np;
    const double* R = P.Rpool + fd.roff;
    const size_t col = (size_t)t * P.Hp + fd.off;
    const double* z = P.z_t + col;
    double* lam_g = P.lam_t + col;
    double* g = P.g_t + col;
    __nv_bfloat16* gbf = P.gbf_t ? reinterpret_cast<__nv_bfloat16*>(P.gbf_t) + col : nullptr;
    int* widx = P.widx + (size_t)c * kWMax;

    // working set = rows with a positive multiplier, in row order (reproducible)
    int m = 0;
    for (int j0 = 0; j0 < n; j0 += 32) {
        const int j = j0 + lane;
        const bool on = j < n && lam_g[j] > 0.0;
        const unsigned bal = __ballot_sync(0xffffffffu, on);
        if (on) {
            const int pos = m + __popc(bal & ((1u << lane) - 1));
            if (pos < kWMax) widx[pos] = j;
            else lam_g[j] = 0.0;                 // cannot be carried; re-admitted if violated
        }
        m += __popc(bal);
    }
    m = min(m, kWMax);
    __syncwarp();
    int cl = (n <= max_warp_n) ? 0 : 1;
    while (cl < kQpClasses - 1 && m > qp_class_cap(cl)) ++cl;

    // g = [z - R_W lam]_+
    for (int j0 = 0; j0 < n; j0 += 32) {
        const int j = j0 + lane;
        if (j >= n) break;
        double pi = 0.0;
        for (int a = 0; a < m; ++a) {
            const int i = widx[a];
            pi = fma(R[(size_t)i * ld + j], lam_g[i], pi);
        }
        const double gj = fmax(z[j] - pi, 0.0);
        g[j] = gj;
        if (gbf) gbf[j] = __float2bfloat16_rn((float)gj);
    }
    if (lane == 0) {
        P.wcount[c] = m;
        P.cls[c] = cl;
        P.status[c] = 0;
        P.inner_ok[c] = 0;
    }
}

cudaError_t launch_qp_init(const QpParams& P, int max_warp_n, cudaStream_t stream) {
    const int wpc = 8;
    qp_init_kernel<<<(P.ncols + wpc - 1) / wpc, 32 * wpc, 0, stream>>>(P, max_warp_n);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
template <int NJ>
__global__ void __launch_bounds__(32 * kWarpsPerCta, 20) utility_qp_warp_kernel(QpParams P) {
    __shared__ WarpSmem smem_all[kWarpsPerCta];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    WarpSmem& sm = smem_all[wib];
    const int slot = blockIdx.x * kWarpsPerCta + wib;
    if (slot >= P.order_count[0]) return;
    const int c = P.order[slot];
    if (P.status[c] != 0 || P.cls[c] != 0) return;
    const int f = c / P.T, t = c % P.T;
    const FeederDev fd = P.feeders[f];
    const int n = fd.n, ld = fd.np;
    const double* R = P.Rpool + fd.roff;
    const size_t col = (size_t)t * P.Hp + fd.off;
    const double* z = P.z_t + col;
    double* lam_g = P.lam_t + col;
    double* g = P.g_t + col;
    double* v = P.v_t + col;
    const double u = P.u, tol = P.tol;
    int* widx = P.widx + (size_t)c * kWMax;
    const unsigned full = 0xffffffffu;

    long long tr_start = 0;
    if (P.trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_start));
    const long long tr_clk0 = clock64();
    const int m_old = P.wcount[c];
    const double thr = (1.0 - kScreenMargin) * u;
    // ---- fast exit: no multipliers and every screened voltage below the safe threshold ->
    // g = [z]_+ from qp_init_kernel is already the projection (most columns, most rounds)
    if (m_old == 0) {
        bool any_cand = false;
#pragma unroll
        for (int k = 0; k < NJ; ++k) {
            const int j = lane + 32 * k;
            if (j < n) any_cand |= P.v32_t ? ((double)(P.v32_t + col)[j] > thr) : (v[j] - u > tol);
        }
        if (!__any_sync(full, any_cand)) {
            if (lane == 0) { P.status[c] = 1; P.inner_ok[c] = 1; atomicAdd(P.n_cls + 0, 1); }
            return;
        }
    }

    // ---- this lane's homes
    double zj[NJ], gj[NJ], vj[NJ];
    bool haslam[NJ];
#pragma unroll
    for (int k = 0; k < NJ; ++k) {
        const int j = lane + 32 * k;
        zj[k] = j < n ? z[j] : 0.0;
        gj[k] = j < n ? g[j] : 0.0;
        haslam[k] = m_old > 0 && j < n && lam_g[j] > 0.0;
    }
    // ---- voltages: screened (BF16) values, exact FP64 recheck of the candidates
    if (P.v32_t) {
        const float* v32 = P.v32_t + col;
#pragma unroll
        for (int k = 0; k < NJ; ++k) {
            const int j = lane + 32 * k;
            const double a = j < n ? (double)v32[j] : 0.0;
            vj[k] = a;
            unsigned cand = __ballot_sync(full, j < n && a > thr && !haslam[k]);
            while (cand) {
                const int src = __ffs(cand) - 1;
                cand &= cand - 1;
                const double* row = R + (size_t)(src + 32 * k) * ld;
                double acc = 0.0;
#pragma unroll
                for (int kk = 0; kk < NJ; ++kk) {
                    const int jj = lane + 32 * kk;
                    if (jj < n) acc = fma(row[jj], gj[kk], acc);
                }
                acc = warp_sum(acc);
                if (lane == src) vj[k] = acc;
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < NJ; ++k) { const int j = lane + 32 * k; vj[k] = j < n ? v[j] : 0.0; }
    }

    // ---- working set: keep rows with a positive multiplier (order preserved), lanes = rows
    int idx = 0;
    double lam = 0.0;
    int m = 0;
    {
        const int i = lane < m_old ? widx[lane] : 0;
        const double l = lane < m_old ? lam_g[i] : 0.0;
        const unsigned keep = __ballot_sync(full, lane < m_old && l > 0.0);
        m = __popc(keep);
        const int src = __fns(keep, 0, lane + 1);          // lane a takes the a-th kept row
        const int si = __shfl_sync(full, i, src & 31);
        const double sl = __shfl_sync(full, l, src & 31);
        if (lane < m) { idx = si; lam = sl; }
    }
    // ---- violated rows outside W, most violated first (ties: lowest row)
    double viol[NJ];
#pragma unroll
    for (int k = 0; k < NJ; ++k) {
        const int j = lane + 32 * k;
        viol[k] = (j < n && !haslam[k] && vj[k] - u > tol) ? vj[k] - u : -1.0;
    }
    int added = 0;
    const int room = min(kAddMaxW, kWW - m);
    bool left = false;
    for (int r = 0; r <= room; ++r) {
        double best = -1.0;
        int bj = 0x7fffffff;
#pragma unroll
        for (int k = 0; k < NJ; ++k)
            if (viol[k] > best) { best = viol[k]; bj = lane + 32 * k; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(full, best, o);
            const int oj = __shfl_xor_sync(full, bj, o);
            if (ob > best || (ob == best && oj < bj)) { best = ob; bj = oj; }
        }
        if (best < 0.0) break;
        if (r == room) { left = true; break; }
        if (lane == m + added) { idx = bj; lam = 0.0; }
#pragma unroll
        for (int k = 0; k < NJ; ++k)
            if (bj == lane + 32 * k) viol[k] = -1.0;
        ++added;
    }
    if (added == 0 && !left && P.inner_ok[c]) {
        if (lane == 0) P.status[c] = 1;
        return;
    }
    auto hand_over = [&]() {
        if (lane == 0) { P.cls[c] = 1; atomicAdd(P.n_running, 1); atomicAdd(P.n_cls + 1, 1); }
    };
    if (left && m + added == kWW) { hand_over(); return; }
    const bool clean = (added == 0 && !left);
    m += added;
    const bool row = lane < m;

    // curvature scale of the Hessian shift
    double scale = warp_sum(row ? P.rn2[fd.off + idx] : 0.0) / (double)max(m, 1);
    const double shift = kHessShiftW * scale + 1e-300;

    // ---- evaluation of phi at multipliers (lane a holds lam_a): fills out[] = [z - R lam]_+
    auto eval = [&](double lam_a, double (&out)[NJ]) -> double {
        double pi[NJ];
#pragma unroll
        for (int k = 0; k < NJ; ++k) pi[k] = 0.0;
        for (int a = 0; a < m; ++a) {
            const double la = warp_bcast(lam_a, a);
            const int ia = __shfl_sync(full, idx, a);
            if (la != 0.0) {
                const double* rr = R + (size_t)ia * ld;
#pragma unroll
                for (int k = 0; k < NJ; ++k) {
                    const int j = lane + 32 * k;
                    if (j < n) pi[k] = fma(rr[j], la, pi[k]);
                }
            }
        }
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < NJ; ++k) {
            const int j = lane + 32 * k;
            out[k] = j < n ? fmax(zj[k] - pi[k], 0.0) : 0.0;
            acc = fma(out[k], out[k], acc);
        }
        return 0.5 * warp_sum(acc) + u * warp_sum(row ? lam_a : 0.0);
    };

    double phi = eval(lam, gj);
    double hrow[kWW];                 // row `lane` of H (columns <= lane are maintained)
#pragma unroll
    for (int q = 0; q < kWW; ++q) hrow[q] = 0.0;
    unsigned fbits = 0;               // bit k: home lane+32k was in F when H was last updated
    bool have_H = false;
    int ok = 0, its = 0;
    bool bail = false;
    double flops = 2.0 * m * n;

    for (; its < P.inner_max; ++its) {
        // gradient on W
        double grad = 0.0;
        for (int a = 0; a < m; ++a) {
            const int ia = __shfl_sync(full, idx, a);
            const double* rr = R + (size_t)ia * ld;
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < NJ; ++k) {
                const int j = lane + 32 * k;
                if (j < n) acc = fma(rr[j], gj[k], acc);
            }
            acc = warp_sum(acc);
            if (lane == a) grad = u - acc;
        }
        flops += 2.0 * m * n;
        const double kk = row ? fabs(lam > 0.0 ? grad : fmin(grad, 0.0)) : 0.0;
        const double kkt = warp_max(kk);
        if (kkt < tol) { ok = 1; break; }

        // Hessian.  First piece: H[p][q] = sum over F of row p times row q, every row read
        // coalesced (lanes = homes), two q at a time so the loads overlap.  Later pieces:
        // signed rank-1 updates for the homes whose membership of F changed.
        {
            int nupd = 0;
            if (!have_H) {
                for (int p = 0; p < m; ++p) {
                    const double* rp_ptr = R + (size_t)__shfl_sync(full, idx, p) * ld;
                    double rp[NJ];
#pragma unroll
                    for (int k = 0; k < NJ; ++k) {
                        const int j = lane + 32 * k;
                        rp[k] = (j < n && gj[k] > 0.0) ? rp_ptr[j] : 0.0;
                    }
                    for (int q = 0; q <= p; q += 2) {
                        const double* r0 = R + (size_t)__shfl_sync(full, idx, q) * ld;
                        const double* r1 = R + (size_t)__shfl_sync(full, idx, min(q + 1, p)) * ld;
                        double a0 = 0.0, a1 = 0.0;
#pragma unroll
                        for (int k = 0; k < NJ; ++k) {
                            const int j = lane + 32 * k;
                            if (j < n) { a0 = fma(rp[k], r0[j], a0); a1 = fma(rp[k], r1[j], a1); }
                        }
                        a0 = warp_sum(a0);
                        a1 = warp_sum(a1);
                        if (lane == p) {
#pragma unroll
                            for (int qq = 0; qq < kWW; ++qq) {
                                if (qq == q) hrow[qq] = a0;
                                if (qq == q + 1 && q + 1 <= p) hrow[qq] = a1;
                            }
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < NJ; ++k) { if (gj[k] > 0.0) { fbits |= 1u << k; ++nupd; } else fbits &= ~(1u << k); }
                nupd = __reduce_add_sync(full, nupd);
            } else {
#pragma unroll
                for (int k = 0; k < NJ; ++k) {
                    const bool now = gj[k] > 0.0;
                    const bool was = (fbits >> k) & 1u;
                    unsigned chg = __ballot_sync(full, now != was);
                    const unsigned nowb = __ballot_sync(full, now);
                    while (chg) {
                        const int src = __ffs(chg) - 1;
                        chg &= chg - 1;
                        const int j = src + 32 * k;
                        const double sgn = ((nowb >> src) & 1u) ? 1.0 : -1.0;
                        const double ra = row ? R[(size_t)idx * ld + j] : 0.0;
#pragma unroll
                        for (int q = 0; q < kWW; ++q) {
                            const double rq = warp_bcast(ra, q);
                            if (q <= lane) hrow[q] = fma(sgn * ra, rq, hrow[q]);
                        }
                        ++nupd;
                    }
                    fbits = now ? (fbits | (1u << k)) : (fbits & ~(1u << k));
                }
            }
            have_H = true;
            flops += (double)m * (m + 1) * nupd;
            __syncwarp();
#pragma unroll
            for (int q = 0; q < kWW; ++q)
                if (row && q <= lane) { sm.H[lane * kHW + q] = hrow[q]; sm.H[q * kHW + lane] = hrow[q]; }
            __syncwarp();
        }

        // ---- exact minimiser of the piece over lam_W >= 0: primal-dual active set, lanes = rows
        double b = 0.0;
        for (int q = 0; q < m; ++q) {
            const double lq = warp_bcast(lam, q);
            if (row && lq != 0.0) b = fma(sm.H[lane * kHW + q], lq, b);
        }
        b += shift * lam - grad;
        bool inA = row && (lam > 0.0 || grad < 0.0);
        double x = 0.0;
        bool pdas_ok = false;
        for (int guess = 0; guess < kPdasMaxW; ++guess) {
            const unsigned Am = __ballot_sync(full, inA);
            const int ma = __popc(Am);
            const int pos = __popc(Am & ((1u << lane) - 1));
            double xs = 0.0;
            if (ma > 0) {
                const int o = (lane < ma) ? (int)__fns(Am, 0, lane + 1) : 0;     // original row of compact row `lane`
                // gather H_AA (+ shift) into L, lane = compact row
                for (int cidx = 0; cidx < ma; ++cidx) {
                    const int oc = __shfl_sync(full, o, cidx);
                    if (lane < ma && cidx <= lane) sm.L[lane * kHW + cidx] = sm.H[o * kHW + oc] + (cidx == lane ? shift : 0.0);
                }
                __syncwarp();
                for (int k2 = 0; k2 < ma; ++k2) {            // Cholesky, lanes own rows
                    const double dkk = sqrt(fmax(sm.L[k2 * kHW + k2], 1e-300));
                    __syncwarp();
                    if (lane == k2) sm.L[k2 * kHW + k2] = dkk;
                    double lik = 0.0;
                    if (lane > k2 && lane < ma) { lik = sm.L[lane * kHW + k2] / dkk; sm.L[lane * kHW + k2] = lik; }
                    __syncwarp();
                    if (lane > k2 && lane < ma)
                        for (int j2 = k2 + 1; j2 <= lane; ++j2)
                            sm.L[lane * kHW + j2] = fma(-lik, sm.L[j2 * kHW + k2], sm.L[lane * kHW + j2]);
                    __syncwarp();
                }
                double y = warp_bcast(b, o);                   // rhs of compact row `lane`
                if (lane >= ma) y = 0.0;
                for (int k2 = 0; k2 < ma; ++k2) {
                    const double yk = warp_bcast(y, k2) / sm.L[k2 * kHW + k2];
                    if (lane == k2) y = yk;
                    if (lane > k2 && lane < ma) y = fma(-sm.L[lane * kHW + k2], yk, y);
                }
                for (int k2 = ma - 1; k2 >= 0; --k2) {
                    const double xk = warp_bcast(y, k2) / sm.L[k2 * kHW + k2];
                    if (lane == k2) y = xk;
                    if (lane < k2) y = fma(-sm.L[k2 * kHW + lane], xk, y);
                }
                xs = y;
                flops += (2.0 / 3.0) * ma * ma * ma + 4.0 * ma * ma + 2.0 * m * ma;
            }
            const double xg = warp_bcast(xs, pos & 31);
            x = inA ? xg : 0.0;
            double mu = 0.0;
            for (int q = 0; q < m; ++q) {
                const double xq = warp_bcast(x, q);
                if (row && xq != 0.0) mu = fma(sm.H[lane * kHW + q], xq, mu);
            }
            mu -= b;                                           // excludes the shift term: x_i = 0 off A
            const bool bad = row && (inA ? (x <= 0.0) : (mu < 0.0));
            if (!__any_sync(full, bad)) { pdas_ok = true; break; }
            if (bad) inA = !inA;
        }
        if (!pdas_ok) { bail = true; break; }

        // ---- line search of phi on the segment lam -> x
        const double dir = x - lam;
        const double slope0 = warp_sum(row ? grad * dir : 0.0);
        double gt[NJ];
        double alpha = 1.0, phin = phi, lt = lam;
        bool stepped = false;
        for (; alpha >= kArcMinW; alpha *= 0.5) {
            lt = row ? fmax(fma(alpha, dir, lam), 0.0) : 0.0;
            phin = eval(lt, gt);
            flops += 2.0 * m * n;
            const double slope = warp_sum(row ? grad * (lt - lam) : 0.0);
            if (phin <= phi + 1e-4 * slope + 1e-14 * fabs(phi)) { stepped = true; break; }
        }
        (void)slope0;
        if (!stepped) { bail = true; break; }
        lam = lt;
        phi = phin;
#pragma unroll
        for (int k = 0; k < NJ; ++k) gj[k] = gt[k];
    }
    if (bail) { hand_over(); return; }       // nothing was written: the CTA kernel redoes the column

    // ---- persist
    if (lane < m_old) lam_g[widx[lane]] = 0.0;
    __syncwarp();
    if (row) { lam_g[idx] = lam; widx[lane] = idx; }
    if (its > 0) {
        __nv_bfloat16* gbf = P.gbf_t ? reinterpret_cast<__nv_bfloat16*>(P.gbf_t) + col : nullptr;
#pragma unroll
        for (int k = 0; k < NJ; ++k) {
            const int j = lane + 32 * k;
            if (j < n) {
                g[j] = gj[k];
                if (gbf) gbf[j] = __float2bfloat16_rn((float)gj[k]);
            }
        }
    }
    const bool done = clean && ok && its == 0;
    if (lane == 0) {
        P.wcount[c] = m;
        P.inner_ok[c] = ok;
        P.status[c] = done ? 1 : 0;
        if (!done) atomicAdd(P.n_running, 1);
        atomicAdd(P.n_cls + 0, 1);
        atomicAdd(P.newton_its, (unsigned long long)its);
        atomicMax(P.max_ws, m);
        atomicAdd(P.flops, (unsigned long long)flops);
        if (P.trace) {
            long long tr_end; unsigned smid;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_end));
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            long long* rec = P.trace + 12 * (size_t)c;
            rec[0] = tr_start; rec[1] = tr_end; rec[2] = smid; rec[3] = ((long long)m << 20) | its;
            rec[11] = clock64() - tr_clk0;
        }
    }
}

int qp_warp_max_n() { return 32 * 16; }

cudaError_t launch_utility_qp_warp(const QpParams& P, int n_cols_bound, int max_n, cudaStream_t stream) {
    if (n_cols_bound <= 0) return cudaSuccess;
    const int grid = (n_cols_bound + kWarpsPerCta - 1) / kWarpsPerCta;
    if (max_n <= 128) utility_qp_warp_kernel<4><<<grid, 32 * kWarpsPerCta, 0, stream>>>(P);
    else if (max_n <= 256) utility_qp_warp_kernel<8><<<grid, 32 * kWarpsPerCta, 0, stream>>>(P);
    else utility_qp_warp_kernel<16><<<grid, 32 * kWarpsPerCta, 0, stream>>>(P);
    return cudaGetLastError();
}

}  // namespace revs
