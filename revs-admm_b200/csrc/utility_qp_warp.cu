// Utility QP, small columns: ONE WARP per (zone, hour) column, no block barriers, no host
// round trips.
//
// After a feeder is split into its voltage zones (feeder.py:split_zones) most columns have
// n ~ 10^2 residences and one or two binding voltage rows.  A column lives in the registers
// of a single warp:
//   lanes own the homes   j = lane + 32 k  (k < NJ, n <= 32 NJ): z_j, g_j, voltage bound of row j
//   lanes own the rows    a < m <= 16 of the working set: idx_a, lam_a, grad_a, H[a][.]
// and the kernel is persistent: every warp pulls columns from a device-side queue (hardest
// first, lists built by order_columns_kernel), so a finished column frees its slot at once and
// columns that need no work never occupy one.
//
// Per column (same fixed point and tolerances as utility_qp.cu / oracle project_voltage):
//   1. screened voltages (BF16 tensor pass) -> exact FP64 recheck of the candidate rows,
//   2. admit the most violated rows, copy the working rows of R into shared memory (cp.async, all
//      rows in flight at once; 8 KB per warp for zones up to 128 residences, 20 KB above) so that
//      gradient / Hessian / line search never go back to L2,
//   3. minimise the dual on W: |W| = 1 is a monotone Newton iteration on a convex piecewise
//      linear function in registers; otherwise piecewise-quadratic descent with an exact
//      primal-dual active-set step (16x16 Cholesky in shared memory) and a segment line search,
//   4. VERIFY every row outside W without leaving the kernel: with D+ = sum_j (g_new - g_old)_+,
//      v_i(g_new) <= v_i(g_old) + max_j R_ij * D+  (all entries of R are >= 0), so rows whose bound
//      stays below u are proven feasible -- usually all of them, because raising multipliers
//      only lowers g -- and the few others are recomputed exactly; violated rows join W -> 2.
// Nothing is written until the column is finished; a column that outgrows 16 rows, cycles or
// fails a line search is handed, untouched, to the CTA kernel of the next class (same round).
// Every group of loads a phase starts with (the column's z / g / screened voltages, the row maxima
// of the verification) is issued before its first use, with clamped indices instead of branches:
// left to the compiler each slot's load sat next to its use, one round trip to memory per slot.
//
// Three kernels live in this file:
//   utility_qp_warp_kernel<NJ>   the general kernel described above (NJ = 4: zones <= 128 residences,
//                                NJ = 6 / 8: zones <= 192 / 256), drains a queue of work lists;
//   utility_qp_fast_kernel       columns with a stored working set of 0 or 1 rows (85 % of the
//                                columns that need a QP kernel): steps 1, 3 (Newton) and 4 entirely
//                                in ~60 registers, 32 resident warps per SM; what needs a second row
//                                is appended to the general kernel's queue;
//   qp_init_kernel               (one warp per column, any size) starts a utility solve: working set
//                                = support of the stored multipliers, class by its size,
//                                g = [z - R lam]_+ for the new target (inside the ADMM loop only for
//                                columns that carry multipliers: dual_update_kernel wrote g = [z]_+).
#include <cstdlib>

#include <cuda_bf16.h>

#include "kernels.cuh"

namespace revs {

namespace {

#ifndef REVS_WARPS_PER_CTA        // build-time experiment knob (profiles/build_variants.sh): warps (= columns in flight) per CTA
#define REVS_WARPS_PER_CTA 4
#endif
constexpr int kWarpsPerCta = REVS_WARPS_PER_CTA;
constexpr int kCtasPerSm = 16 / kWarpsPerCta;   // 16 resident columns per SM
constexpr int kHW = kWW + 1;                 // leading dim of the per-warp 16x16 matrices
constexpr int kCacheDoubles = 1024;          // per-warp cache of working rows of R (zones of up to 128 residences: 8 rows)
constexpr double kArcMinW = 9.5367431640625e-07;
constexpr int kPdasMaxW = 40;
constexpr double kHessShiftW = 1e-12;
constexpr int kAddMaxW = 8;                  // violated rows admitted per pass
constexpr int kPassMaxW = 10;                // admit / solve / verify passes before the column is handed on
constexpr int kRecheckMaxW = 256;             // rows re-evaluated exactly per verification; more -> next screening pass
constexpr int kHysteresisW = 8;              // warm working sets above kWW - this start in class 1 (long columns: a whole CTA each)

template <int CACHE>
struct WarpSmemT {
    double H[kWW * kHW];      // model Hessian, full symmetric
    double L[kWW * kHW];      // Cholesky factor of the active sub-matrix
    double rows[CACHE];
};
// zones of up to 128 residences (NJ = 4, 4 CTAs per SM): 1024 doubles of cached working rows per warp;
// larger zones (NJ >= 6, 2 CTAs per SM): REVS_WARP_CACHE_BIG doubles
// CTAs per SM of the NJ >= 6 instantiations.  Measured on the reference-shaped population (profiles/README_r02.md):
// at 4 (128 registers) and 3 (168) the kernels spill 130..570 bytes per thread into a 28 KB L1 and every phase of a column
// runs twice as long; at 2 (255 registers, no spills) the step is 13 % faster in spite of half the resident warps.
#ifndef REVS_WARP_CTAS_SMALL      // build-time experiment knobs (profiles/build_variants.sh)
#define REVS_WARP_CTAS_SMALL (8 / REVS_WARPS_PER_CTA)
#endif
#ifndef REVS_WARP_CTAS_BIG
#define REVS_WARP_CTAS_BIG (8 / REVS_WARPS_PER_CTA)
#endif
#ifndef REVS_WARP_CACHE_BIG         // doubles of cached working rows per warp at NJ >= 6 (2 CTAs of 4 warps per SM): 20 KB = 13 / 10 / 8
#define REVS_WARP_CACHE_BIG 2560    // rows of a zone of 192 / 256 / 320 residences, all rows of 97 % of the columns
#endif
template <int NJ> struct WarpCfg {
    static constexpr int kCache = NJ <= 4 ? kCacheDoubles : REVS_WARP_CACHE_BIG;
    static constexpr int kCtas = NJ <= 8 ? (NJ <= 4 ? kCtasPerSm : REVS_WARP_CTAS_SMALL) : REVS_WARP_CTAS_BIG;
};

struct WarpStats {
    unsigned long long its = 0;
    double flops = 0.0;
    int handed = 0, deferred = 0, max_ws = 0, cols = 0;
};

__device__ __forceinline__ double warp_bcast(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// Exact FP64 voltages of the rows flagged in candk (bit k of a lane: row lane+32k) for the
// schedule in gj[]: four rows at a time so that their loads and reductions overlap.  The loop
// over k is a run-time loop on purpose (code size: the kernel must stay in the instruction cache).
template <int NJ>
__device__ __forceinline__ void recheck_rows(unsigned candk, const double* __restrict__ R, int ld, int n,
                                             const double (&gj)[NJ], double (&vub)[NJ]) {
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    if (!__any_sync(full, candk != 0u)) return;
#pragma unroll 1
    for (int k = 0; k < NJ; ++k) {
        unsigned cand = __ballot_sync(full, (candk >> k) & 1u);
        while (cand) {
            int s[4], cnt = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (cand) { s[q] = __ffs(cand) - 1; cand &= cand - 1; ++cnt; }
                else s[q] = s[0];
            }
            double a[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int kk = 0; kk < NJ; ++kk) {
                const int jj = lane + 32 * kk;
                if (jj < n) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) a[q] = fma(R[(size_t)(s[q] + 32 * k) * ld + jj], gj[kk], a[q]);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int q = 0; q < 4; ++q) a[q] += __shfl_xor_sync(full, a[q], o);
            }
            double mine = 0.0;
            bool hit = false;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (q < cnt && lane == s[q]) { mine = a[q]; hit = true; }
            if (hit) {
#pragma unroll
                for (int kk = 0; kk < NJ; ++kk)
                    if (kk == k) vub[kk] = mine;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
template <int NJ>
__device__ __forceinline__ void solve_column(const QpParams& P, const int4 ent, WarpSmemT<WarpCfg<NJ>::kCache>& sm, WarpStats& st) {
    constexpr int kCacheD = WarpCfg<NJ>::kCache;
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    // the work-list entry carries the zone geometry, so every load of the column starts at once
    const int c = ent.x;
    const int n = ent.z & 0xffff, ld = ent.z >> 16;
    const int t = c % P.T;
    const size_t hoff = (size_t)ent.y;
    const double* __restrict__ R = P.Rpool + ((size_t)ent.w << 4);
    const size_t col = (size_t)t * P.Hp + hoff;
    const double* z = P.z_t + col;
    double* lam_g = P.lam_t + col;
    double* g = P.g_t + col;
    const double u = P.u, tol = P.tol;
    int* widx = P.widx + (size_t)c * kWMax;
    const int m_old = P.wcount[c];
    const bool solved_before = P.inner_ok[c] == 1;
    const int wi = lane < kWW ? widx[lane] : 0;        // speculative (valid for lane < m_old): the multipliers depend on it

    long long tr_start = 0;
    if (P.trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_start));
    const long long tr_clk0 = clock64();
    const double thr = (1.0 - kScreenMargin) * u;

    // ---- this lane's homes; vub[k] is an upper bound of the voltage of row lane+32k for the g
    // in g0[] and EXACT whenever it exceeds u (invariant kept by every step below)
    double zj[NJ], gj[NJ], g0[NJ], vub[NJ];
    unsigned inw = 0;                          // bit k: row lane+32k is in the working set
    unsigned candk = 0;                        // bit k: row lane+32k must be re-evaluated exactly
    // (every load of the column is issued before the first value is used, with clamped indices instead of
    // branches: one round trip to memory for the whole column, not one per slot)
    if (P.v32_t) {
        const float* v32 = P.v32_t + col;
        float vf[NJ];
#pragma unroll
        for (int k = 0; k < NJ; ++k) {
            const int jc = min(lane + 32 * k, n - 1);
            zj[k] = z[jc];
            gj[k] = g[jc];
            vf[k] = v32[jc];
        }
        // the row maxima are needed after the solve (verification): on their way to L2 meanwhile
        if (lane < (n + 15) / 16) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.rmax + hoff + 16 * lane));
#pragma unroll
        for (int k = 0; k < NJ; ++k) {
            const bool in = lane + 32 * k < n;
            if (!in) { zj[k] = 0.0; gj[k] = 0.0; }
            g0[k] = gj[k];
            const double a = in ? (double)vf[k] : 0.0;
            vub[k] = kScreenUp * a;
            if (in && a > thr) candk |= 1u << k;
        }
    } else {
        const double* v64 = P.v_t + col;
#pragma unroll
        for (int k = 0; k < NJ; ++k) {
            const int jc = min(lane + 32 * k, n - 1);
            zj[k] = z[jc];
            gj[k] = g[jc];
            vub[k] = v64[jc];
        }
#pragma unroll
        for (int k = 0; k < NJ; ++k) {
            if (lane + 32 * k >= n) { zj[k] = 0.0; gj[k] = 0.0; vub[k] = 0.0; }
            g0[k] = gj[k];
        }
    }

    // ---- working set: rows with a positive multiplier (order preserved), lanes = rows
    int idx = 0, m = 0;
    double lam = 0.0;
    if (m_old > 0) {
        const double l = lane < m_old ? lam_g[wi] : 0.0;
        const unsigned keep = __ballot_sync(full, lane < m_old && l > 0.0);
        m = __popc(keep);
        const int src = __fns(keep, 0, lane + 1);          // lane a takes the a-th kept row
        const int si = __shfl_sync(full, wi, src & 31);
        const double sl = __shfl_sync(full, l, src & 31);
        if (lane < m) { idx = si; lam = sl; }
#pragma unroll 1
        for (int a = 0; a < m; ++a) {                      // membership bits of the homes that own a working row
            const int h = __shfl_sync(full, idx, a);
            if ((h & 31) == lane) inw |= 1u << (h >> 5);
        }
        candk &= ~inw;
    }

    double flops = 0.0;
    int its_total = 0;
    bool changed = false;                      // g differs from the stored iterate
    long long tph[4] = {0, 0, 0, 0}, tq0 = clock64(), tq1;   // debug trace: cycles in load / admit+cache / solve / verify
    int npass = 0;
#define WPHASE(i) do { if (P.trace) { tq1 = clock64(); tph[i] += tq1 - tq0; tq0 = tq1; } } while (0)
    auto trace_out = [&](int kind) {
        if (P.trace && lane == 0) {
            long long tr_end; unsigned smid;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_end));
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            long long* rec = P.trace + 12 * (size_t)c;
            rec[0] = tr_start; rec[1] = tr_end; rec[2] = smid; rec[3] = ((long long)m << 20) | its_total;
            for (int i = 0; i < 4; ++i) rec[4 + i] = tph[i];
            rec[8] = npass; rec[9] = kind; rec[10] = m_old; rec[11] = clock64() - tr_clk0;
        }
    };
    tph[0] = tq0 - tr_clk0;
    // exit kinds: 0 finished (persist), 1 deferred to the next screening pass (persist, still running),
    // 2 handed to the CTA class untouched
    int exit_kind = 0, reason = 0;             // reason of a hand-over (debug trace)
    double grad = 0.0;                         // u - v of this lane's row at the last solution

    for (int pass = 0;; ++pass) {
        // ---- exact voltages of the candidate rows (screening candidates / rows whose bound failed)
        recheck_rows<NJ>(candk, R, ld, n, gj, vub);
        candk = 0;
        WPHASE(pass == 0 ? 0 : 3);

        // ---- rows whose multiplier went to zero leave W (their exact voltage is u - grad)
        if (pass > 0) {
            const bool row = lane < m;
            unsigned drop = __ballot_sync(full, row && !(lam > 0.0));
            if (drop) {
                const unsigned keep = __ballot_sync(full, row && lam > 0.0);
                while (drop) {
                    const int a = __ffs(drop) - 1;
                    drop &= drop - 1;
                    const int h = __shfl_sync(full, idx, a);
                    const double vr = u - warp_bcast(grad, a);
                    if ((h & 31) == lane) {
#pragma unroll
                        for (int k = 0; k < NJ; ++k)
                            if (k == (h >> 5)) { vub[k] = vr; inw &= ~(1u << k); }
                    }
                }
                m = __popc(keep);
                const int src = __fns(keep, 0, lane + 1);
                const int si = __shfl_sync(full, idx, src & 31);
                const double sl = __shfl_sync(full, lam, src & 31);
                idx = lane < m ? si : 0;
                lam = lane < m ? sl : 0.0;
            }
        }

        // ---- violated rows outside W, most violated first (ties: lowest row)
        int added = 0;
        bool left = false;
        const int room = min(kAddMaxW, kWW - m);
#pragma unroll 1
        for (int r = 0; r <= room; ++r) {
            double best = -1.0;
            int bj = 0x7fffffff;
#pragma unroll
            for (int k = 0; k < NJ; ++k) {
                const int j = lane + 32 * k;
                const double viol = vub[k] - u;
                if (j < n && !((inw >> k) & 1u) && viol > tol && viol > best) { best = viol; bj = j; }
            }
            if (!__any_sync(full, best >= 0.0)) break;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(full, best, o);
                const int oj = __shfl_xor_sync(full, bj, o);
                if (ob > best || (ob == best && oj < bj)) { best = ob; bj = oj; }
            }
            if (r == room) { left = true; break; }
            if (lane == m + added) { idx = bj; lam = 0.0; }
            if ((bj & 31) == lane) inw |= 1u << (bj >> 5);
            ++added;
        }
        if (added == 0 && !left) {
            // no violated row: after a solve + verification this is the KKT point; at the start it
            // is one if there is nothing to solve or the stored iterate was solved already
            if (pass > 0 || m == 0 || solved_before) break;
        }
        if ((left && m + added == kWW) || pass >= kPassMaxW) { exit_kind = 2; reason = pass >= kPassMaxW ? 2 : 1; break; }
        m += added;
        ++npass;
        const bool row = lane < m;

        // ---- working rows of R into shared memory (as many as fit)
        // (cp.async, 16 bytes per request: all rows are in flight at once -- one L2 round trip for the whole
        // working set instead of one per row; rows start on 128-byte boundaries, ld is a multiple of 16)
        const int ncache = min(m, kCacheD / ld);
        __syncwarp();
        {
            const int nchunk = (n + 1) >> 1;
#pragma unroll 1
            for (int a = 0; a < ncache; ++a) {
                const double* src = R + (size_t)__shfl_sync(full, idx, a) * ld;
                const unsigned dst = (unsigned)__cvta_generic_to_shared(sm.rows + a * ld);
                for (int ch = lane; ch < nchunk; ch += 32)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * ch), "l"(src + 2 * ch) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp();
        auto rowp = [&](int a) -> const double* {      // a is warp-uniform
            const int ia = __shfl_sync(full, idx, a);
            return a < ncache ? sm.rows + a * ld : R + (size_t)ia * ld;
        };

        int ok = 0, its = 0;
        bool bail = false;
        WPHASE(1);

        // ---- |W| = 1: the dual is one-dimensional, v(lam) = r . [z - r lam]_+ is convex, piecewise
        // linear and non-increasing; Newton on v(lam) = u converges monotonically from the left and
        // lands on the exact root of the final piece
        if (m == 1) {
            const double* rp = rowp(0);
            double r1[NJ];
#pragma unroll
            for (int k = 0; k < NJ; ++k) {
                const int j = lane + 32 * k;
                r1[k] = j < n ? rp[j] : 0.0;
            }
            const double l_in = warp_bcast(lam, 0);
            double l = l_in;
#pragma unroll 1
            for (int it = 0; it < 48; ++it) {
                double v = 0.0, S = 0.0;
#pragma unroll
                for (int k = 0; k < NJ; ++k) {
                    const double gk = fmax(zj[k] - r1[k] * l, 0.0);
                    gj[k] = gk;
                    v = fma(r1[k], gk, v);
                    if (gk > 0.0) S = fma(r1[k], r1[k], S);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    v += __shfl_xor_sync(full, v, o);
                    S += __shfl_xor_sync(full, S, o);
                }
                ++its;
                const double fr = v - u;
                if ((l > 0.0 ? fabs(fr) : fmax(fr, 0.0)) < tol) { ok = 1; grad = -fr; break; }
                double ln = S > 0.0 ? l + fr / S : 0.0;
                if (ln < 0.0) ln = 0.0;
                if (ln == l) break;                    // stagnation: let the general path decide
                l = ln;
            }
            flops += 4.0 * n * its;
            lam = lane == 0 ? l : 0.0;                 // gj[] is [z - r l]_+ either way
            if (l != l_in) changed = true;
            if (ok && l == l_in) its = 0;
            if (!ok) its = 0;
        }

        if (!ok) {
            // ---- general path: piecewise-quadratic descent on W
            const double scale = warp_sum(row ? P.rn2[hoff + idx] : 0.0) / (double)max(m, 1);
            const double shift = kHessShiftW * scale + 1e-300;
            double phi;
            {
                double acc = 0.0;
#pragma unroll
                for (int k = 0; k < NJ; ++k) acc = fma(gj[k], gj[k], acc);
                phi = 0.5 * warp_sum(acc) + u * warp_sum(row ? lam : 0.0);
            }
            double hrow[kWW];                 // row `lane` of H (columns <= lane are maintained)
#pragma unroll
            for (int q = 0; q < kWW; ++q) hrow[q] = 0.0;
            unsigned fbits = 0;               // bit k: home lane+32k was in F when H was last updated
            bool have_H = false;

#pragma unroll 1
            for (; its < P.inner_max; ++its) {
                // gradient on W
                grad = 0.0;
#pragma unroll 1
                for (int a = 0; a < m; ++a) {
                    const double* rr = rowp(a);
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

                // Hessian.  First piece: H[p][q] = sum over F of row p times row q (lanes = homes),
                // two q at a time so the loads overlap.  Later pieces: signed rank-1 updates for the
                // homes whose membership of F changed.
                {
                    int nupd = 0;
                    unsigned nowmask = 0;
#pragma unroll
                    for (int k = 0; k < NJ; ++k)
                        if (gj[k] > 0.0) nowmask |= 1u << k;
                    if (!have_H) {
#pragma unroll 1
                        for (int p = 0; p < m; ++p) {
                            const double* rp_ptr = rowp(p);
                            double rp[NJ];
#pragma unroll
                            for (int k = 0; k < NJ; ++k) {
                                const int j = lane + 32 * k;
                                rp[k] = (j < n && gj[k] > 0.0) ? rp_ptr[j] : 0.0;
                            }
#pragma unroll 1
                            for (int q = 0; q <= p; q += 2) {
                                const double* r0 = rowp(q);
                                const double* r1 = rowp(min(q + 1, p));
                                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                                for (int k = 0; k < NJ; ++k) {
                                    const int j = lane + 32 * k;
                                    if (j < n) { a0 = fma(rp[k], r0[j], a0); a1 = fma(rp[k], r1[j], a1); }
                                }
#pragma unroll
                                for (int o = 16; o > 0; o >>= 1) {
                                    a0 += __shfl_xor_sync(full, a0, o);
                                    a1 += __shfl_xor_sync(full, a1, o);
                                }
                                if (lane == p) {
#pragma unroll
                                    for (int qq = 0; qq < kWW; ++qq) {
                                        if (qq == q) hrow[qq] = a0;
                                        if (qq == q + 1 && q + 1 <= p) hrow[qq] = a1;
                                    }
                                }
                            }
                        }
                        nupd = __reduce_add_sync(full, __popc(nowmask));
                    } else {
#pragma unroll 1
                        for (int k = 0; k < NJ; ++k) {
                            const bool now = (nowmask >> k) & 1u;
                            const bool was = (fbits >> k) & 1u;
                            unsigned chg = __ballot_sync(full, now != was);
                            const unsigned nowb = __ballot_sync(full, now);
                            while (chg) {
                                const int src = __ffs(chg) - 1;
                                chg &= chg - 1;
                                const int j = src + 32 * k;
                                const double sgn = ((nowb >> src) & 1u) ? 1.0 : -1.0;
                                const double ra = row ? (lane < ncache ? sm.rows[lane * ld + j] : R[(size_t)idx * ld + j]) : 0.0;
#pragma unroll
                                for (int q = 0; q < kWW; ++q) {
                                    const double rq = warp_bcast(ra, q);
                                    if (q <= lane) hrow[q] = fma(sgn * ra, rq, hrow[q]);
                                }
                                ++nupd;
                            }
                        }
                    }
                    fbits = nowmask;
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
#pragma unroll 1
                for (int q = 0; q < m; ++q) {
                    const double lq = warp_bcast(lam, q);
                    if (row && lq != 0.0) b = fma(sm.H[lane * kHW + q], lq, b);
                }
                b += shift * lam - grad;
                bool inA = row && (lam > 0.0 || grad < 0.0);
                double x = 0.0;
                bool pdas_ok = false;
#pragma unroll 1
                for (int guess = 0; guess < kPdasMaxW; ++guess) {
                    const unsigned Am = __ballot_sync(full, inA);
                    const int ma = __popc(Am);
                    const int pos = __popc(Am & ((1u << lane) - 1));
                    double xs = 0.0;
                    if (ma > 0) {
                        const int o = (lane < ma) ? (int)__fns(Am, 0, lane + 1) : 0;     // original row of compact row `lane`
                        // gather H_AA (+ shift) into L, lane = compact row
#pragma unroll 1
                        for (int cidx = 0; cidx < ma; ++cidx) {
                            const int oc = __shfl_sync(full, o, cidx);
                            if (lane < ma && cidx <= lane) sm.L[lane * kHW + cidx] = sm.H[o * kHW + oc] + (cidx == lane ? shift : 0.0);
                        }
                        __syncwarp();
                        double rdiag = 1.0;
#pragma unroll 1
                        for (int k2 = 0; k2 < ma; ++k2) {            // Cholesky, left-looking: lanes own rows, one column
                            double sv = 0.0;                         // per step, no stores inside the dot product
                            if (lane >= k2 && lane < ma) {
                                sv = sm.L[lane * kHW + k2];
                                for (int p2 = 0; p2 < k2; ++p2) sv = fma(-sm.L[lane * kHW + p2], sm.L[k2 * kHW + p2], sv);
                            }
                            const double skk = fmax(warp_bcast(sv, k2), 1e-300);
                            const double rk = rsqrt(skk);            // one special-function call per column instead of sqrt + divide
                            const double dkk = skk * rk;
                            if (lane == k2) rdiag = rk;              // reciprocal pivots stay in registers: no division in the solves
                            if (lane >= k2 && lane < ma) sm.L[lane * kHW + k2] = lane == k2 ? dkk : sv * rk;
                            __syncwarp();
                        }
                        double y = warp_bcast(b, o);                   // rhs of compact row `lane`
                        if (lane >= ma) y = 0.0;
#pragma unroll 1
                        for (int k2 = 0; k2 < ma; ++k2) {
                            const double yk = warp_bcast(y, k2) * warp_bcast(rdiag, k2);
                            if (lane == k2) y = yk;
                            if (lane > k2 && lane < ma) y = fma(-sm.L[lane * kHW + k2], yk, y);
                        }
#pragma unroll 1
                        for (int k2 = ma - 1; k2 >= 0; --k2) {
                            const double xk = warp_bcast(y, k2) * warp_bcast(rdiag, k2);
                            if (lane == k2) y = xk;
                            if (lane < k2) y = fma(-sm.L[k2 * kHW + lane], xk, y);
                        }
                        xs = y;
                        flops += (2.0 / 3.0) * ma * ma * ma + 4.0 * ma * ma + 2.0 * m * ma;
                    }
                    const double xg = warp_bcast(xs, pos & 31);
                    x = inA ? xg : 0.0;
                    double mu = 0.0;
#pragma unroll 1
                    for (int q = 0; q < m; ++q) {
                        const double xq = warp_bcast(x, q);
                        if (row && xq != 0.0) mu = fma(sm.H[lane * kHW + q], xq, mu);
                    }
                    mu -= b;                                           // excludes the shift term: x_i = 0 off A
                    const bool bad = row && (inA ? (x <= 0.0) : (mu < 0.0));
                    if (!__any_sync(full, bad)) { pdas_ok = true; break; }
                    if (bad) inA = !inA;
                }
                if (!pdas_ok) { bail = true; reason = 3; break; }

                // ---- line search of phi on the segment lam -> x
                const double dir = x - lam;
                double gt[NJ];
                double alpha = 1.0, phin = phi, lt = lam;
                bool stepped = false;
#pragma unroll 1
                for (; alpha >= kArcMinW; alpha *= 0.5) {
                    lt = row ? fmax(fma(alpha, dir, lam), 0.0) : 0.0;
                    {   // phi(lt), gt = [z - R lt]_+
                        double pi[NJ];
#pragma unroll
                        for (int k = 0; k < NJ; ++k) pi[k] = 0.0;
#pragma unroll 1
                        for (int a = 0; a < m; ++a) {
                            const double la = warp_bcast(lt, a);
                            const double* rr = rowp(a);
                            if (la != 0.0) {
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
                            gt[k] = j < n ? fmax(zj[k] - pi[k], 0.0) : 0.0;
                            acc = fma(gt[k], gt[k], acc);
                        }
                        phin = 0.5 * warp_sum(acc) + u * warp_sum(row ? lt : 0.0);
                    }
                    flops += 2.0 * m * n;
                    const double slope = warp_sum(row ? grad * (lt - lam) : 0.0);
                    if (phin <= phi + 1e-4 * slope + 1e-14 * fabs(phi)) { stepped = true; break; }
                }
                if (!stepped) { bail = true; reason = 4; break; }
                lam = lt;
                phi = phin;
                changed = true;
#pragma unroll
                for (int k = 0; k < NJ; ++k) gj[k] = gt[k];
            }
        }
        its_total += its;
        WPHASE(2);
        if (bail || !ok) { exit_kind = 2; if (!reason) reason = 5; break; }

        // ---- verification of the rows outside W for the new g (see the header): rows whose
        // bound fails are re-evaluated at the top of the next pass
        double dp = 0.0;
#pragma unroll
        for (int k = 0; k < NJ; ++k) dp += fmax(gj[k] - g0[k], 0.0);
        dp = warp_sum(dp) * (1.0 + 1e-9);
        int ncand = 0;
        {
            const double* rmax = P.rmax + hoff;
            double rm[NJ];
#pragma unroll
            for (int k = 0; k < NJ; ++k) rm[k] = rmax[min(lane + 32 * k, n - 1)];     // all in flight at once
#pragma unroll
            for (int k = 0; k < NJ; ++k) {
                const int j = lane + 32 * k;
                if (j < n && !((inw >> k) & 1u)) {
                    const double bnd = dp > 0.0 ? fma(rm[k], dp, vub[k]) : vub[k];
                    if (bnd - u > tol) { candk |= 1u << k; ++ncand; }
                    else vub[k] = bnd;
                }
            }
            ncand = __reduce_add_sync(full, ncand);
        }
#pragma unroll
        for (int k = 0; k < NJ; ++k) g0[k] = gj[k];
        flops += 2.0 * n * min(ncand, kRecheckMaxW);
        if (ncand == 0) break;                         // every row proven feasible: KKT point
        if (ncand > kRecheckMaxW) { exit_kind = 1; break; }   // too many rows to settle here
    }
    WPHASE(3);

    if (exit_kind == 2) {                      // nothing was written: the next class redoes the column
        if (lane == 0) { P.cls[c] = 1; P.inner_ok[c] = 2; }
        ++st.handed;
        trace_out(2 + 16 * reason);
        return;
    }
    // ---- persist: finished column, or (exit_kind 1) an iterate for the next tensor-core
    // screening pass to check
    if (changed || m != m_old) {
        if (lane < m_old) lam_g[wi] = 0.0;
        __syncwarp();
        if (lane < m) { lam_g[idx] = lam; widx[lane] = idx; }
        if (changed) {
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
    }
    if (lane == 0) {
        P.wcount[c] = m;
        P.inner_ok[c] = 1;
        if (exit_kind == 0) P.status[c] = 1;
    }
    if (exit_kind == 1) ++st.deferred;
    trace_out(exit_kind);
    st.its += (unsigned long long)its_total;
    st.flops += flops;
    st.max_ws = max(st.max_ws, m);
#undef WPHASE
}

}  // namespace

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) qp_init_kernel(QpParams P, int max_warp_n) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (blockIdx.x == 0 && threadIdx.x == 0 && P.round_ctr) {     // start of a utility solve: first working-set round follows
        *P.round_ctr = 0;
        if (P.use_cond) cudaGraphSetConditional((cudaGraphConditionalHandle)P.cond_round, 1u);
    }
    if (c >= P.ncols) return;
    const int f = c / P.T, t = c % P.T;
    const FeederDev fd = P.feeders[f];
    const int n = fd.n, ld = fd.np;
    if (fd.roff < 0) {                 // zone of the tree-Newton path (tree_newton.cu): not a column of the dense kernels
        if (lane == 0) { P.cls[c] = 0; P.status[c] = 1; P.inner_ok[c] = 1; if (P.cand) P.cand[c] = 0; }
        return;
    }
    const double* R = P.Rpool + fd.roff;
    const size_t col = (size_t)t * P.Hp + fd.off;
    const double* z = P.z_t + col;
    double* lam_g = P.lam_t + col;
    double* g = P.g_t + col;
    __nv_bfloat16* gbf = P.gbf_t ? reinterpret_cast<__nv_bfloat16*>(P.gbf_t) + col : nullptr;
    int* widx = P.widx + (size_t)c * kWMax;

    int m = 0;
    if (P.init == 2) {
        // inside the ADMM loop: the multipliers live on the stored working rows only, and
        // dual_update_kernel has already written g = [z]_+ -- a column without a stored row is ready
        const int wc = P.wcount[c];
        if (wc == 0) {
            if (lane == 0) {
                P.cls[c] = n <= max_warp_n ? 0 : 1;
                P.status[c] = 0;
                P.inner_ok[c] = 0;
                if (P.cand) P.cand[c] = 0;
            }
            return;
        }
        for (int a0 = 0; a0 < wc; a0 += 32) {      // compact in place: rows with a positive multiplier, order kept
            const int a = a0 + lane;
            const int i = a < wc ? widx[a] : 0;
            const bool on = a < wc && lam_g[i] > 0.0;
            const unsigned bal = __ballot_sync(0xffffffffu, on);
            __syncwarp();
            if (on) widx[m + __popc(bal & ((1u << lane) - 1))] = i;
            m += __popc(bal);
        }
    } else {
        // working set = rows with a positive multiplier, in row order (reproducible)
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
    }
    __syncwarp();
    int cl = (n <= max_warp_n && m <= (n <= 128 ? P.warp_m_max : P.warp_m_max_big)) ? 0 : 1;
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
        if (P.cand) P.cand[c] = 0;
    }
}

cudaError_t launch_qp_init(const QpParams& P, int max_warp_n, cudaStream_t stream) {
    const int wpc = 8;
    qp_init_kernel<<<(P.ncols + wpc - 1) / wpc, 32 * wpc, 0, stream>>>(P, max_warp_n);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
template <int NJ>
__global__ void __launch_bounds__(32 * kWarpsPerCta, WarpCfg<NJ>::kCtas) utility_qp_warp_kernel(QpParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using Smem = WarpSmemT<WarpCfg<NJ>::kCache>;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    Smem& sm = reinterpret_cast<Smem*>(smem_raw)[wib];
    // the bucket lists [list0, list0 + nlists) (hardest first), then list_extra, form one queue
    const int l0 = P.list0, nl = P.nlists;
    int cnt[kQpBuckets], total = 0;
#pragma unroll
    for (int b = 0; b < kQpBuckets; ++b) {
        const int li = b < nl ? l0 + b : ((b == nl && P.list_extra >= 0) ? P.list_extra : -1);
        cnt[b] = li >= 0 ? P.order_count[li] : 0;
        total += cnt[b];
    }
    WarpStats st;
    for (;;) {
        int slot = 0;
        if (lane == 0) slot = atomicAdd(P.queue + l0, 1);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (slot >= total) break;
        int b = 0;
#pragma unroll
        for (int bb = 0; bb < kQpBuckets - 1; ++bb)
            if (b == bb && slot >= cnt[bb]) { slot -= cnt[bb]; ++b; }
        const int li = b < nl ? l0 + b : P.list_extra;
        const int4 ent = P.order4[(size_t)(li - kQpClasses) * P.ncols + slot];
        solve_column<NJ>(P, ent, sm, st);
        ++st.cols;
        __syncwarp();
    }
    if (lane == 0) {
        if (st.its) atomicAdd(P.newton_its, st.its);
        if (st.flops > 0.0) atomicAdd(P.flops, (unsigned long long)st.flops);
        if (st.max_ws) atomicMax(P.max_ws, st.max_ws);
        if (st.cols) atomicAdd(P.cols, (unsigned long long)st.cols);
        if (st.handed) { atomicAdd(P.n_running, st.handed); atomicAdd(P.n_cls + 1, st.handed); }
        if (st.deferred) { atomicAdd(P.n_running, st.deferred); atomicAdd(P.n_cls + 0, st.deferred); }
    }
}

// ------------------------------------------------------------------------------------------
// One-row columns of the small zones (stored working set of 0 or 1 rows: 85 % of the columns that
// need a QP kernel).  The whole column lives in ~60 registers per lane -- z, the single working
// row of R, the voltage bounds -- so twice as many columns are resident per SM as in the general
// kernel, and the code is a tenth of its size.  Handled here: no violated row (clean), or exactly
// one binding row before and after the solve (monotone Newton on v(lam) = u, then the same
// in-kernel verification as the general kernel).  Anything else -- a second violated row, a
// returning column -- is appended, untouched, to the leftover list for the general kernel.
constexpr int kFastWarps = 8;

__global__ void __launch_bounds__(32 * kFastWarps, 4) utility_qp_fast_kernel(QpParams P) {
    constexpr int NJ = 4;
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    const int l0 = kQpClasses + 2;
    const int c1 = P.order_count[l0], total = c1 + P.order_count[l0 + 1];
    const double u = P.u, tol = P.tol, thr = (1.0 - kScreenMargin) * u;
    unsigned long long its_sum = 0;
    double flops = 0.0;
    int cols = 0, passed = 0, maxws = 0;
    for (;;) {
        int slot = 0;
        if (lane == 0) slot = atomicAdd(P.queue + l0, 1);
        slot = __shfl_sync(full, slot, 0);
        if (slot >= total) break;
        const int4 ent = slot < c1 ? P.order4[(size_t)(l0 - kQpClasses) * P.ncols + slot]
                                   : P.order4[(size_t)(l0 + 1 - kQpClasses) * P.ncols + slot - c1];
        const int c = ent.x, n = ent.z & 0xffff, ld = ent.z >> 16;
        const size_t hoff = (size_t)ent.y;
        const double* __restrict__ R = P.Rpool + ((size_t)ent.w << 4);
        const size_t col = (size_t)(c % P.T) * P.Hp + hoff;
        const double* z = P.z_t + col;
        double* lam_g = P.lam_t + col;
        int* widx = P.widx + (size_t)c * kWMax;
        const int m_old = P.wcount[c];
        const int solved_before = P.inner_ok[c];
        const int w0 = widx[0];                          // valid if m_old == 1
        ++cols;

        double zj[NJ], rj[NJ], vub[NJ], gj[NJ];
        unsigned candk = 0;
        if (P.v32_t) {                                   // (loads first, uses afterwards: see solve_column)
            const float* v32 = P.v32_t + col;
            float vf[NJ];
#pragma unroll
            for (int k = 0; k < NJ; ++k) {
                const int jc = min(lane + 32 * k, n - 1);
                zj[k] = z[jc];
                vf[k] = v32[jc];
            }
#pragma unroll
            for (int k = 0; k < NJ; ++k) {
                const bool in = lane + 32 * k < n;
                if (!in) zj[k] = 0.0;
                const double a = in ? (double)vf[k] : 0.0;
                vub[k] = kScreenUp * a;
                if (in && a > thr) candk |= 1u << k;
            }
        } else {
            const double* v64 = P.v_t + col;
#pragma unroll
            for (int k = 0; k < NJ; ++k) {
                const int jc = min(lane + 32 * k, n - 1);
                zj[k] = z[jc];
                vub[k] = v64[jc];
            }
#pragma unroll
            for (int k = 0; k < NJ; ++k)
                if (lane + 32 * k >= n) { zj[k] = 0.0; vub[k] = 0.0; }
        }
        auto pass_on = [&]() {                           // untouched: the general kernel redoes the column
            if (lane == 0) {
                const int pos = atomicAdd(const_cast<int*>(P.order_count) + kListLeftover, 1);
                P.order4[(size_t)(kListLeftover - kQpClasses) * P.ncols + pos] = ent;
            }
            ++passed;
        };
        if (m_old > 1 || solved_before != 0) { pass_on(); continue; }
        int i0 = -1;                                     // the working row
        double lam_old = 0.0;
        if (m_old == 1) {
            lam_old = lam_g[w0];
            if (lam_old > 0.0) i0 = w0; else lam_old = 0.0;
        }
        // stored iterate g = [z - r lam_old]_+ (qp_init_kernel / dual_update_kernel wrote exactly this)
#pragma unroll
        for (int k = 0; k < NJ; ++k) {
            const int j = lane + 32 * k;
            rj[k] = (i0 >= 0 && j < n) ? R[(size_t)i0 * ld + j] : 0.0;
            gj[k] = fmax(zj[k] - rj[k] * lam_old, 0.0);
        }
        if (i0 >= 0 && (i0 & 31) == lane) candk &= ~(1u << (i0 >> 5));
        recheck_rows<NJ>(candk, R, ld, n, gj, vub);
        // violated rows outside W (vub is exact wherever it exceeds u)
        double best = -1.0;
        int bj = 0x7fffffff, nviol = 0;
#pragma unroll
        for (int k = 0; k < NJ; ++k) {
            const int j = lane + 32 * k;
            const double viol = vub[k] - u;
            if (j < n && j != i0 && viol > tol) { ++nviol; if (viol > best) { best = viol; bj = j; } }
        }
        nviol = __reduce_add_sync(full, nviol);
        if (nviol > 1 || (nviol == 1 && i0 >= 0)) { pass_on(); continue; }
        if (nviol == 1) {                                // admit the one violated row
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(full, best, o);
                const int oj = __shfl_xor_sync(full, bj, o);
                if (ob > best || (ob == best && oj < bj)) { best = ob; bj = oj; }
            }
            i0 = bj;
#pragma unroll
            for (int k = 0; k < NJ; ++k) {
                const int j = lane + 32 * k;
                rj[k] = j < n ? R[(size_t)i0 * ld + j] : 0.0;
            }
        }
        double l = lam_old;
        int its = 0;
        bool ok = i0 < 0;                                // nothing to solve: g = [z]_+ is the projection
        if (i0 >= 0) {
#pragma unroll 1
            for (int it = 0; it < 48; ++it) {
                double v = 0.0, S = 0.0;
#pragma unroll
                for (int k = 0; k < NJ; ++k) {
                    const double gk = fmax(zj[k] - rj[k] * l, 0.0);
                    gj[k] = gk;
                    v = fma(rj[k], gk, v);
                    if (gk > 0.0) S = fma(rj[k], rj[k], S);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    v += __shfl_xor_sync(full, v, o);
                    S += __shfl_xor_sync(full, S, o);
                }
                ++its;
                const double fr = v - u;
                if ((l > 0.0 ? fabs(fr) : fmax(fr, 0.0)) < tol) { ok = true; break; }
                double ln = S > 0.0 ? l + fr / S : 0.0;
                if (ln < 0.0) ln = 0.0;
                if (ln == l) break;
                l = ln;
            }
            flops += 4.0 * n * its;
        }
        if (!ok) { pass_on(); continue; }
        // verification of the other rows for the new g: monotone bound, exact recheck of the rest
        double dp = 0.0;
#pragma unroll
        for (int k = 0; k < NJ; ++k) dp += fmax(gj[k] - fmax(zj[k] - rj[k] * lam_old, 0.0), 0.0);   // lam_old = 0 for a newly admitted row
        dp = warp_sum(dp) * (1.0 + 1e-9);
        candk = 0;
        {
            const double* rmax = P.rmax + hoff;
            int ncand = 0;
            double rm[NJ];
#pragma unroll
            for (int k = 0; k < NJ; ++k) rm[k] = rmax[min(lane + 32 * k, n - 1)];
#pragma unroll
            for (int k = 0; k < NJ; ++k) {
                const int j = lane + 32 * k;
                if (j < n && j != i0) {
                    const double bnd = dp > 0.0 ? fma(rm[k], dp, vub[k]) : vub[k];
                    if (bnd - u > tol) { candk |= 1u << k; ++ncand; }
                    else vub[k] = bnd;
                }
            }
            flops += 2.0 * n * __reduce_add_sync(full, ncand);
        }
        recheck_rows<NJ>(candk, R, ld, n, gj, vub);
        bool viol = false;
#pragma unroll
        for (int k = 0; k < NJ; ++k) {
            const int j = lane + 32 * k;
            viol |= j < n && j != i0 && vub[k] - u > tol;
        }
        if (__any_sync(full, viol)) { pass_on(); continue; }       // a second row binds: general kernel

        // persist
        const bool changed = (i0 >= 0) && (l != lam_old || nviol == 1);
        const int m_new = i0 >= 0 ? 1 : 0;
        if (lane == 0) {
            if (m_old == 1) lam_g[w0] = 0.0;
            if (i0 >= 0) { lam_g[i0] = l; widx[0] = i0; }
            P.wcount[c] = m_new;
            P.inner_ok[c] = 1;
            P.status[c] = 1;
        }
        if (changed) {
            double* g = P.g_t + col;
            __nv_bfloat16* gbf = P.gbf_t ? reinterpret_cast<__nv_bfloat16*>(P.gbf_t) + col : nullptr;
#pragma unroll
            for (int k = 0; k < NJ; ++k) {
                const int j = lane + 32 * k;
                if (j < n) {
                    g[j] = gj[k];
                    if (gbf) gbf[j] = __float2bfloat16_rn((float)gj[k]);
                }
            }
            its_sum += (unsigned long long)its;
        }
        maxws = max(maxws, m_new);
        __syncwarp();
    }
    if (lane == 0) {
        if (its_sum) atomicAdd(P.newton_its, its_sum);
        if (flops > 0.0) atomicAdd(P.flops, (unsigned long long)flops);
        if (maxws) atomicMax(P.max_ws, maxws);
        if (cols - passed > 0) atomicAdd(P.cols, (unsigned long long)(cols - passed));
        if (P.dbg && passed) atomicAdd(P.dbg + 3, (unsigned long long)passed);     // debug: shown as `fallbacks`
    }
}

cudaError_t launch_utility_qp_fast(const QpParams& P, cudaStream_t stream) {
    int dev = 0, n_sm = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    utility_qp_fast_kernel<<<n_sm * 4, 32 * kFastWarps, 0, stream>>>(P);
    return cudaGetLastError();
}

int qp_warp_max_n() { return kWarpMaxN; }

// Zones up to 128 residences run the NJ = 4 instantiation, larger ones (own work lists) NJ = 6 / 8 / 10 by the
// largest zone present (<= 192 / 256 / 320 residences); each fits the instruction cache.  Function attributes
// are per device: they are set once for every device this process launches on.
static int g_warp_n_sm[4][64] = {};     // per instantiation and device: SM count once the attributes are set

template <int NJ>
static cudaError_t prepare_warp_nj(int* n_sm_out) {
    const int smem = (int)sizeof(WarpSmemT<WarpCfg<NJ>::kCache>) * kWarpsPerCta;
    int* n_sm = g_warp_n_sm[(NJ - 4) / 2];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    dev &= 63;
    if (!n_sm[dev]) {
        int n = 0;
        e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(utility_qp_warp_kernel<NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
#ifndef REVS_WARP_NO_CARVEOUT
        if (e == cudaSuccess) e = cudaFuncSetAttribute(utility_qp_warp_kernel<NJ>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
#endif
        if (e != cudaSuccess) return e;
        n_sm[dev] = n;
    }
    if (n_sm_out) *n_sm_out = n_sm[dev];
    return cudaSuccess;
}

cudaError_t qp_warp_prepare() {
    cudaError_t e = prepare_warp_nj<4>(nullptr);
    if (e == cudaSuccess) e = prepare_warp_nj<6>(nullptr);
    if (e == cudaSuccess) e = prepare_warp_nj<8>(nullptr);
    if (e == cudaSuccess) e = prepare_warp_nj<10>(nullptr);
    return e;
}

template <int NJ>
static cudaError_t launch_warp_nj(const QpParams& P, int ctas_per_sm, cudaStream_t stream) {
    const int smem = (int)sizeof(WarpSmemT<WarpCfg<NJ>::kCache>) * kWarpsPerCta;
    int nsm = 0;
    cudaError_t e = prepare_warp_nj<NJ>(&nsm);
    if (e != cudaSuccess) return e;
    int n_sm[1] = {nsm};
    const int dev = 0;
    const int cmax = WarpCfg<NJ>::kCtas;
    ctas_per_sm = ctas_per_sm < 1 ? 1 : (ctas_per_sm > cmax ? cmax : ctas_per_sm);
    utility_qp_warp_kernel<NJ><<<n_sm[dev] * ctas_per_sm, 32 * kWarpsPerCta, smem, stream>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_utility_qp_warp(const QpParams& P, int nj, int ctas_per_sm, cudaStream_t stream) {
    if (nj == 10) return launch_warp_nj<10>(P, ctas_per_sm, stream);
    if (nj == 8) return launch_warp_nj<8>(P, ctas_per_sm, stream);
    if (nj == 6) return launch_warp_nj<6>(P, ctas_per_sm, stream);
    return launch_warp_nj<4>(P, ctas_per_sm, stream);
}

int qp_warp_ctas_per_sm() { return kCtasPerSm; }
int qp_warp_m_max_default() { return kWW - kHysteresisW; }

}  // namespace revs
