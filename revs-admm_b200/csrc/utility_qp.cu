// Operator (utility) sub-problem of the ADMM loop: one CTA per (feeder, hour) column.
//
// Reference: class Utility (lpsolver.py:160-240), a Gurobi QP over all residences and
// hours at once.  It separates over hours; each hour is the Euclidean projection of
//     z = (P_est + P_sch)/2 - Gamma/kappa
// onto { g >= 0,  R g <= u },  u = vhigh^2 - vset^2  (Gurobi's default lb=0; the vlow row
// is vacuous for g>=0, R>=0, vlow<=vset -- checked on the host).
//
// Method (exact, terminates on KKT residuals, same fixed point as the oracle's
// project_voltage): work on the dual
//     min_{lam>=0}  phi(lam) = 1/2 || [z - R lam]_+ ||^2 + u sum(lam)
// with a WORKING SET W of voltage rows (only a few tens of the ~10^3 rows of a feeder ever
// carry a multiplier).  A launch of this kernel
//   1. reads the voltages  v = R g  of the current iterate, produced for ALL rows and all
//      hours at once by the tensor-core contraction (contract_f64.cu),
//   2. drops rows whose multiplier is zero, admits the most violated rows (v > u),
//   3. solves the dual restricted to W by projected Newton: Hessian R_WF R_FW over the
//      homes F with g>0, Cholesky in shared memory, Armijo search along the projection
//      arc.  This touches only |W| rows of R (coalesced row reads), never the full block.
// and the host alternates it with the contraction until no column has a violated row.
#include "kernels.cuh"

namespace revs {

constexpr int kQpThreads = 256;
constexpr int kJT = 32;                  // columns of R per Hessian tile
constexpr int kHld = kWMax + 1;          // leading dim of H in shared memory
constexpr int kTld = kJT + 1;
constexpr double kArcMin = 9.5367431640625e-07;   // 2^-20, shortest arc-search step


struct QpSmem {
    double H[kWMax * kHld];
    double tileR[kWMax * kTld];
    double lam[kWMax], trial[kWMax], grad[kWMax], dir[kWMax];
    double red[kQpThreads / 32];
    double bcast[2];
    int idx[kWMax], fl[kWMax];
    int ired[kQpThreads / 32];
    int ibcast[2];
};

__device__ __forceinline__ double block_sum(double v, QpSmem& S) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) S.red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
#pragma unroll
    for (int w = 0; w < kQpThreads / 32; ++w) r += S.red[w];
    return r;
}
__device__ __forceinline__ double block_max(double v, QpSmem& S) {
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) S.red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = S.red[0];
#pragma unroll
    for (int w = 1; w < kQpThreads / 32; ++w) r = fmax(r, S.red[w]);
    return r;
}

// phi(lam) for multipliers lam[0..m) on rows idx[0..m); optionally stores g.
__device__ double eval_phi(const double* __restrict__ R, int ld, int n, const double* __restrict__ z,
                           const int* idx, const double* lam, int m, double u, double* g_store,
                           QpSmem& S) {
    double part = 0.0;
    for (int j = threadIdx.x; j < n; j += kQpThreads) {
        double pi = 0.0;
        for (int a = 0; a < m; ++a) {
            const double l = lam[a];
            if (l != 0.0) pi = fma(R[(size_t)idx[a] * ld + j], l, pi);
        }
        const double gj = fmax(z[j] - pi, 0.0);
        if (g_store) g_store[j] = gj;
        part = fma(gj, gj, part);
    }
    double sl = 0.0;
    for (int a = threadIdx.x; a < m; a += kQpThreads) sl += lam[a];
    return 0.5 * block_sum(part, S) + u * block_sum(sl, S);
}

// H[p][q] = sum_{j: g_j>0} R[row_p][j] R[row_q][j] for the mf free rows, NB = ceil(mf/16).
template <int NB>
__device__ void hessian(const double* __restrict__ R, int ld, int n, const double* g, int mf, QpSmem& S) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = tid & 15, ty = tid >> 4;
    double acc[NB][NB];
#pragma unroll
    for (int a = 0; a < NB; ++a)
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[a][b] = 0.0;
    // rows beyond mf read a zero row of the tile
    for (int p = mf + warp; p < 16 * NB; p += kQpThreads / 32) S.tileR[p * kTld + lane] = 0.0;
    for (int j0 = 0; j0 < n; j0 += kJT) {
        __syncthreads();
        for (int p = warp; p < mf; p += kQpThreads / 32) {
            const int j = j0 + lane;
            double val = 0.0;
            if (j < n && g[j] > 0.0) val = R[(size_t)S.idx[S.fl[p]] * ld + j];
            S.tileR[p * kTld + lane] = val;
        }
        __syncthreads();
#pragma unroll 4
        for (int jj = 0; jj < kJT; ++jj) {
            double pa[NB], qb[NB];
#pragma unroll
            for (int a = 0; a < NB; ++a) pa[a] = S.tileR[(ty + 16 * a) * kTld + jj];
#pragma unroll
            for (int b = 0; b < NB; ++b) qb[b] = S.tileR[(tx + 16 * b) * kTld + jj];
#pragma unroll
            for (int a = 0; a < NB; ++a)
#pragma unroll
                for (int b = 0; b < NB; ++b) acc[a][b] = fma(pa[a], qb[b], acc[a][b]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < NB; ++a)
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int p = ty + 16 * a, q = tx + 16 * b;
            if (p < mf && q < mf) S.H[p * kHld + q] = acc[a][b];
        }
    __syncthreads();
}

__global__ void __launch_bounds__(kQpThreads, 1) utility_qp_kernel(QpParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    QpSmem& S = *reinterpret_cast<QpSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = blockIdx.x;
    const int f = c / P.T, t = c % P.T;
    if (!P.init && P.status[c] != 0) return;

    const FeederDev fd = P.feeders[f];
    const int n = fd.n, ld = fd.np;
    const double* R = P.Rpool + fd.roff;
    const size_t col = (size_t)t * P.Hp + fd.off;
    const double* z = P.z_t + col;
    double* lam_g = P.lam_t + col;
    double* g = P.g_t + col;
    const double* v = P.v_t + col;
    const double u = P.u, tol = P.tol;
    const double* rn2 = P.rn2 + fd.off;

    // ------------------------------------------------------------ working set
    int m = 0;
    if (P.init) {
        // warm start: rows that carried a multiplier in the previous ADMM iteration
        // (ordered compaction, so the working-set order -- and with it every rounding -- is
        // reproducible from run to run)
        int base = 0;
        for (int j0 = 0; j0 < n; j0 += kQpThreads) {
            const int j = j0 + tid;
            const bool on = j < n && lam_g[j] > 0.0;
            const unsigned bal = __ballot_sync(0xffffffffu, on);
            __syncthreads();
            if (lane == 0) S.ired[warp] = __popc(bal);
            __syncthreads();
            int before = 0, total = 0;
#pragma unroll
            for (int w = 0; w < kQpThreads / 32; ++w) {
                before += (w < warp) ? S.ired[w] : 0;
                total += S.ired[w];
            }
            if (on) {
                const int pos = base + before + __popc(bal & ((1u << lane) - 1));
                if (pos < kWMax) { S.idx[pos] = j; S.lam[pos] = lam_g[j]; }
                else lam_g[j] = 0.0;   // cannot be carried; re-admitted if violated
            }
            base += total;
        }
        __syncthreads();
        m = min(base, kWMax);
    } else {
        const int m_old = P.wcount[c];
        // keep rows with a positive multiplier (serial compaction keeps the order stable)
        if (tid == 0) {
            int k = 0;
            for (int a = 0; a < m_old; ++a) {
                int i = P.widx[(size_t)c * kWMax + a];
                double l = lam_g[i];
                if (l > 0.0) { S.idx[k] = i; S.lam[k] = l; ++k; }
            }
            S.ibcast[0] = k;
        }
        __syncthreads();
        m = S.ibcast[0];
        // violated rows outside W, most violated first; key order (viol desc, index asc)
        double prev_v = 1e300;
        int prev_i = -1;
        int added = 0, n_viol_left = 0;
        const int room = min(kAddMax, kWMax - m);
        for (int round = 0; round <= room; ++round) {
            double best = -1.0;
            int besti = 0x7fffffff;
            for (int j = tid; j < n; j += kQpThreads) {
                const double viol = v[j] - u;
                if (viol > tol && !(lam_g[j] > 0.0)) {
                    bool after_prev = (viol < prev_v) || (viol == prev_v && j > prev_i);
                    if (after_prev && (viol > best || (viol == best && j < besti))) { best = viol; besti = j; }
                }
            }
            // block arg-max
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                double ob = __shfl_xor_sync(0xffffffffu, best, o);
                int oi = __shfl_xor_sync(0xffffffffu, besti, o);
                if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
            }
            __syncthreads();
            if (lane == 0) { S.red[warp] = best; S.ired[warp] = besti; }
            __syncthreads();
            best = S.red[0]; besti = S.ired[0];
#pragma unroll
            for (int w = 1; w < kQpThreads / 32; ++w) {
                if (S.red[w] > best || (S.red[w] == best && S.ired[w] < besti)) { best = S.red[w]; besti = S.ired[w]; }
            }
            if (best < 0.0) break;              // no further violated row
            if (round == room) { n_viol_left = 1; break; }
            if (tid == 0) { S.idx[m + added] = besti; S.lam[m + added] = 0.0; }
            ++added;
            prev_v = best; prev_i = besti;
        }
        __syncthreads();
        if (added == 0 && n_viol_left == 0 && P.inner_ok[c]) {
            if (tid == 0) P.status[c] = 1;      // KKT point of the full problem
            return;
        }
        if (added == 0 && n_viol_left && m == kWMax) {
            if (tid == 0) { P.status[c] = 2; atomicAdd(P.n_failed, 1); }
            return;
        }
        // clear the stored multipliers of the old set; rewritten at the end
        for (int a = tid; a < m_old; a += kQpThreads) lam_g[P.widx[(size_t)c * kWMax + a]] = 0.0;
        m += added;
        __syncthreads();
    }

    // ------------------------------------------------------------ restricted projected Newton
    double phi = eval_phi(R, ld, n, z, S.idx, S.lam, m, u, g, S);
    double tau = 1.0;
    int ok = 0, its = 0;
    for (; its < P.inner_max; ++its) {
        __syncthreads();
        // gradient on W:  u - R[idx_a] . g
        for (int a = warp; a < m; a += kQpThreads / 32) {
            const double* row = R + (size_t)S.idx[a] * ld;
            double acc = 0.0;
            for (int j = lane; j < n; j += 32) acc = fma(row[j], g[j], acc);
            acc = warp_sum(acc);
            if (lane == 0) S.grad[a] = u - acc;
        }
        __syncthreads();
        double kk = 0.0;
        for (int a = tid; a < m; a += kQpThreads) {
            double gr = S.grad[a];
            kk = fmax(kk, fabs(S.lam[a] > 0.0 ? gr : fmin(gr, 0.0)));
        }
        const double kkt = block_max(kk, S);
        if (kkt < tol) { ok = 1; break; }

        // free rows; rows pinned at (almost) zero with a positive gradient go to exactly 0
        const double eps = fmin(1e-8, kkt);
        if (tid == 0) {
            int k = 0;
            double sc = 0.0;
            for (int a = 0; a < m; ++a) {
                bool bound = (S.lam[a] <= eps) && (S.grad[a] > 0.0);
                S.dir[a] = -S.lam[a];
                if (!bound) { S.fl[k++] = a; sc += rn2[S.idx[a]]; }
            }
            S.ibcast[1] = k;
            S.bcast[0] = k ? sc / (double)k : 0.0;
        }
        __syncthreads();
        const int mf = S.ibcast[1];
        const double scale = S.bcast[0];    // mean |R_a|^2 of the free rows: curvature scale

        double alpha = 1.0, phin = phi;
        for (;;) {   // Levenberg-Marquardt safeguard: raise the shift until the arc search succeeds
            if (mf > 0) {
                // Hessian H = R_{A,F} R_{F,A}, register-blocked over a 16x16 thread grid
                const int nb = (mf + 15) >> 4;
                if (nb <= 1) hessian<1>(R, ld, n, g, mf, S);
                else if (nb <= 2) hessian<2>(R, ld, n, g, mf, S);
                else if (nb <= 3) hessian<3>(R, ld, n, g, mf, S);
                else if (nb <= 4) hessian<4>(R, ld, n, g, mf, S);
                else if (nb <= 6) hessian<6>(R, ld, n, g, mf, S);
                else hessian<8>(R, ld, n, g, mf, S);
                const double reg = 1e-10 * tau * scale + 1e-300;
                for (int p = tid; p < mf; p += kQpThreads) S.H[p * kHld + p] += reg;
                __syncthreads();

                // Cholesky H = L L^T (lower, in place)
                for (int k = 0; k < mf; ++k) {
                    if (tid == 0) S.H[k * kHld + k] = sqrt(fmax(S.H[k * kHld + k], 1e-300));
                    __syncthreads();
                    const double dkk = S.H[k * kHld + k];
                    for (int i = k + 1 + tid; i < mf; i += kQpThreads) S.H[i * kHld + k] /= dkk;
                    __syncthreads();
                    const int cnt = mf - k - 1;
                    for (int e = tid; e < cnt * cnt; e += kQpThreads) {
                        int i = k + 1 + e / cnt, j = k + 1 + e % cnt;
                        if (j <= i) S.H[i * kHld + j] = fma(-S.H[i * kHld + k], S.H[j * kHld + k], S.H[i * kHld + j]);
                    }
                    __syncthreads();
                }
                // solve L y = -grad_A ; L^T d = y   (trial[] is scratch for y)
                for (int p = tid; p < mf; p += kQpThreads) S.trial[p] = -S.grad[S.fl[p]];
                __syncthreads();
                for (int k = 0; k < mf; ++k) {
                    if (tid == 0) S.trial[k] /= S.H[k * kHld + k];
                    __syncthreads();
                    const double yk = S.trial[k];
                    for (int i = k + 1 + tid; i < mf; i += kQpThreads) S.trial[i] = fma(-S.H[i * kHld + k], yk, S.trial[i]);
                    __syncthreads();
                }
                for (int k = mf - 1; k >= 0; --k) {
                    if (tid == 0) S.trial[k] /= S.H[k * kHld + k];
                    __syncthreads();
                    const double xk = S.trial[k];
                    for (int i = tid; i < k; i += kQpThreads) S.trial[i] = fma(-S.H[k * kHld + i], xk, S.trial[i]);
                    __syncthreads();
                }
                for (int p = tid; p < mf; p += kQpThreads) S.dir[S.fl[p]] = S.trial[p];
                __syncthreads();
            }

            // Armijo search along the projection arc
            bool found = false;
            for (alpha = 1.0; alpha >= kArcMin; alpha *= 0.5) {
                for (int a = tid; a < m; a += kQpThreads) S.trial[a] = fmax(fma(alpha, S.dir[a], S.lam[a]), 0.0);
                __syncthreads();
                double sl = 0.0;
                for (int a = tid; a < m; a += kQpThreads) sl = fma(S.grad[a], S.trial[a] - S.lam[a], sl);
                const double slope = block_sum(sl, S);
                phin = eval_phi(R, ld, n, z, S.idx, S.trial, m, u, nullptr, S);
                // + rounding noise of phi itself, see oracle/revs_oracle.py:project_voltage
                if (phin <= phi + 1e-4 * slope + 1e-14 * fabs(phi)) { found = true; break; }
            }
            if (found || tau > 1e40 || mf == 0) break;
            tau *= 1e3;
        }
        if (alpha == 1.0) tau = fmax(1.0, tau / 10.0);
        __syncthreads();
        for (int a = tid; a < m; a += kQpThreads) S.lam[a] = S.trial[a];
        __syncthreads();
        phi = eval_phi(R, ld, n, z, S.idx, S.lam, m, u, g, S);
    }

    // ------------------------------------------------------------ persist
    __syncthreads();
    for (int a = tid; a < m; a += kQpThreads) {
        lam_g[S.idx[a]] = S.lam[a];
        P.widx[(size_t)c * kWMax + a] = S.idx[a];
    }
    if (tid == 0) {
        P.wcount[c] = m;
        P.inner_ok[c] = ok;
        P.status[c] = 0;
        atomicAdd(P.n_running, 1);
        atomicAdd(P.newton_its, (unsigned long long)its);
        atomicMax(P.max_ws, m);
    }
}

cudaError_t launch_utility_qp(const QpParams& P, int ncols, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(utility_qp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(QpSmem));
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    utility_qp_kernel<<<ncols, kQpThreads, sizeof(QpSmem), stream>>>(P);
    return cudaGetLastError();
}

}  // namespace revs
