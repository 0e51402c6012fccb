// Utility QP on the feeder TREE: one warp per (zone, hour) column, no sensitivity matrix at all.
//
// Reference: class Utility (lpsolver.py:163-238) builds R = 2 F D F^T densely (compute_Rmat,
// lpsolver.py:17-26) and hands the rows R g <= u to Gurobi.  For a radial feeder R has a closed form:
// with the residences of a zone in depth-first order,
//     R[i][j] = 2 cumr(lca(i, j)) = min(c[i .. j-1])   (i < j),   R[i][i] = d[i],
// c[p] = 2 cumr(lca(p, p+1)) being the sensitivity of two depth-first neighbours.  Everything the
// projection needs follows from 2 n numbers per zone instead of n^2:
//
//   * a ROW of R is a prefix-min / suffix-min scan of c away from its diagonal (gen_row: two warp scans),
//   * the PRODUCT R x for all rows at once is three prefix sums: the Cartesian tree of c has one node q
//     per c-position, covering the leaves [lo_q, hi_q] with weight w_q = c[q] - c[parent(q)], and
//         (R x)[p] = sum_{q : lo_q <= p <= hi_q} w_q X_q + e[p] x[p],     X_q = sum_{lo_q <= j <= hi_q} x[j];
//     X_q are differences of the prefix sum of x, and the sum over the nodes that cover p is a prefix sum
//     over the nodes sorted by lo minus one over the nodes sorted by hi (tree_product).  O(n) work, exact
//     to rounding (1e-16 relative), no memory traffic beyond ~40 bytes of static data per residence.
//
// So a column is solved from z alone: g = [z - R lam]_+ for the warm start, the voltages of ALL rows
// exactly (no BF16 screening pass, no candidate rechecks, no verification bounds), the working-set
// solve (same fixed point, tolerances and safeguards as utility_qp_warp.cu / oracle project_voltage:
// monotone Newton for one row, piecewise-quadratic descent with an exact primal-dual active-set step
// otherwise -- the Hessian column of row q is one product of that row restricted to {g > 0}), and the
// exact voltages again to admit what the new iterate violates.  Nothing is written until the column is
// finished; a column that outgrows 16 rows or fails a safeguard is left, untouched, to the dense
// kernels (counted in TreeParams::left).  Lanes own CONTIGUOUS runs of NJ depth-first positions, so
// every scan is NJ serial steps plus five shuffles.
#include <math_constants.h>

#include "kernels.cuh"

namespace revs {

namespace {

constexpr int kTWarps = 4;                   // warps (columns in flight) per CTA
constexpr int kTH = kWW + 1;                 // leading dimension of the 16 x 16 matrices
constexpr double kArcMinT = 9.5367431640625e-07;
constexpr int kPdasMaxT = 40;
constexpr double kHessShiftT = 1e-12;
constexpr int kAddMaxT = 8;                  // violated rows admitted per pass
constexpr int kPassMaxT = 12;                // admit / solve / verify passes before the column is left to the dense path

template <int NJ>
struct TreeSmem {
    double G[32 * NJ + 1];                   // prefix sums / scatter-gather scratch
    double S1[32 * NJ + 1];
    double S2[32 * NJ + 1];
    double H[kWW * kTH];
    double L[kWW * kTH];
};

template <int NJ> struct TreeCfg { static constexpr int kCtas = NJ <= 6 ? 4 : 3; };

struct ZonePtr {                             // static arrays of one zone (TreeParams pools + zone offset)
    const int* perm;
    const int* iperm;
    const double* c;
    const double* d;
    const double* e;
    const int* nodeA;
    const double* wA;
    const int* nodeB;
    const double* wB;
    const int* cnt;
    int n;
};

// in-place inclusive prefix sum over the warp's 32 * NJ positions (lane-major, contiguous)
template <int NJ>
__device__ __forceinline__ void scan_incl(double (&x)[NJ]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 1; k < NJ; ++k) x[k] += x[k - 1];
    double t = x[NJ - 1];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
    }
    double ex = __shfl_up_sync(0xffffffffu, t, 1);
    if (lane == 0) ex = 0.0;
#pragma unroll
    for (int k = 0; k < NJ; ++k) x[k] += ex;
}

// v = R x for every row of the zone (see the header).  x is kept; smem arrays are scratch.
template <int NJ>
__device__ __forceinline__ void tree_product(const ZonePtr& Z, const double (&x)[NJ], double (&v)[NJ], TreeSmem<NJ>& sm) {
    const int lane = threadIdx.x & 31, p0 = lane * NJ, n = Z.n;
    double t[NJ];
#pragma unroll
    for (int k = 0; k < NJ; ++k) t[k] = x[k];
    scan_incl<NJ>(t);
    __syncwarp();
    if (lane == 0) { sm.G[0] = 0.0; sm.S1[0] = 0.0; sm.S2[0] = 0.0; }
#pragma unroll
    for (int k = 0; k < NJ; ++k) sm.G[p0 + k + 1] = t[k];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < NJ; ++k) {
        const int q = p0 + k;
        double tq = 0.0;
        if (q < n - 1) {
            const int nd = Z.nodeA[q];
            tq = Z.wA[q] * (sm.G[(nd >> 16) + 1] - sm.G[nd & 0xffff]);
        }
        t[k] = tq;
    }
    scan_incl<NJ>(t);
#pragma unroll
    for (int k = 0; k < NJ; ++k) sm.S1[p0 + k + 1] = t[k];
#pragma unroll
    for (int k = 0; k < NJ; ++k) {
        const int q = p0 + k;
        double tq = 0.0;
        if (q < n - 1) {
            const int nd = Z.nodeB[q];
            tq = Z.wB[q] * (sm.G[(nd >> 16) + 1] - sm.G[nd & 0xffff]);
        }
        t[k] = tq;
    }
    scan_incl<NJ>(t);
#pragma unroll
    for (int k = 0; k < NJ; ++k) sm.S2[p0 + k + 1] = t[k];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < NJ; ++k) {
        const int p = p0 + k;
        double r = 0.0;
        if (p < n) {
            const int cn = Z.cnt[p];
            r = fma(Z.e[p], x[k], sm.S1[cn & 0xffff] - sm.S2[cn >> 16]);
        }
        v[k] = r;
    }
}

// row i of R (depth-first positions): min of c over the positions between i and j, d[i] on the diagonal
template <int NJ>
__device__ __forceinline__ void gen_row(const ZonePtr& Z, int i, double (&row)[NJ]) {
    const int lane = threadIdx.x & 31, p0 = lane * NJ, n = Z.n;
    double cq[NJ], right[NJ], left[NJ];
#pragma unroll
    for (int k = 0; k < NJ; ++k) cq[k] = (p0 + k < n - 1) ? Z.c[p0 + k] : 0.0;     // beyond the zone: 0 -> padded entries are 0
    double run = CUDART_INF;
#pragma unroll
    for (int k = 0; k < NJ; ++k) {                     // j > i: min over c-positions [i, j-1]
        right[k] = run;
        if (p0 + k >= i) run = fmin(run, cq[k]);
    }
    double t = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t = fmin(t, y);
    }
    double offr = __shfl_up_sync(0xffffffffu, t, 1);
    if (lane == 0) offr = CUDART_INF;
    run = CUDART_INF;
#pragma unroll
    for (int k = NJ - 1; k >= 0; --k) {                // j < i: min over c-positions [j, i-1]
        if (p0 + k < i) run = fmin(run, cq[k]);
        left[k] = run;
    }
    t = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double y = __shfl_down_sync(0xffffffffu, t, o);
        if (lane + o < 32) t = fmin(t, y);
    }
    double offl = __shfl_down_sync(0xffffffffu, t, 1);
    if (lane == 31) offl = CUDART_INF;
    const double di = Z.d[i];
#pragma unroll
    for (int k = 0; k < NJ; ++k) {
        const int p = p0 + k;
        row[k] = p > i ? fmin(offr, right[k]) : (p < i ? fmin(offl, left[k]) : di);
    }
}

// x[p] = val_a at the positions pos_a of the working rows (lanes a < m), 0 elsewhere
template <int NJ>
__device__ __forceinline__ void scatter_rows(int m, int pos, double val, double (&x)[NJ], TreeSmem<NJ>& sm) {
    const int lane = threadIdx.x & 31, p0 = lane * NJ;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < NJ; ++k) sm.G[p0 + k] = 0.0;
    __syncwarp();
    if (lane < m) sm.G[pos] = val;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < NJ; ++k) x[k] = sm.G[p0 + k];
}

// value of the position vector v at the position of this lane's working row
template <int NJ>
__device__ __forceinline__ double gather_rows(int m, int pos, const double (&v)[NJ], TreeSmem<NJ>& sm) {
    const int lane = threadIdx.x & 31, p0 = lane * NJ;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < NJ; ++k) sm.G[p0 + k] = v[k];
    __syncwarp();
    return lane < m ? sm.G[pos] : 0.0;
}

struct TreeStats {
    unsigned long long its = 0;
    double flops = 0.0;
    int left = 0, max_ws = 0, cols = 0;
};

template <int NJ>
__device__ __forceinline__ void tree_column(const QpParams& P, const TreeParams& TP, const int c, TreeSmem<NJ>& sm, TreeStats& st) {
    const int lane = threadIdx.x & 31, p0 = lane * NJ;
    const unsigned full = 0xffffffffu;
    const int f = c / P.T, t = c % P.T;
    const FeederDev fd = P.feeders[f];
    const int n = fd.n;
    const size_t zo = (size_t)fd.off;
    ZonePtr Z{TP.perm + zo, TP.iperm + zo, TP.c + zo, TP.d + zo, TP.e + zo, TP.nodeA + zo, TP.wA + zo, TP.nodeB + zo, TP.wB + zo, TP.cnt + zo, n};
    const size_t col = (size_t)t * P.Hp + zo;
    const double* __restrict__ z = P.z_t + col;
    double* lam_g = P.lam_t + col;
    const double u = P.u, tol = P.tol;
    int* widx = P.widx + (size_t)c * kWMax;
    const int m_old = P.wcount[c];
    if (m_old > kWW) { ++st.left; return; }                // stored set beyond this kernel: dense path
    const int wi = lane < m_old ? widx[lane] : 0;          // home index of the stored rows

    int hk[NJ];
    double zj[NJ], gj[NJ], v[NJ];
#pragma unroll
    for (int k = 0; k < NJ; ++k) {
        const int p = p0 + k;
        hk[k] = p < n ? Z.perm[p] : -1;
        zj[k] = p < n ? z[hk[k]] : 0.0;
    }

    // ---- working set: rows with a positive multiplier (order kept), lanes = rows; pos = depth-first position
    int pos = 0, m = 0;
    double lam = 0.0;
    if (m_old > 0) {
        const double l = lane < m_old ? lam_g[wi] : 0.0;
        const unsigned keep = __ballot_sync(full, lane < m_old && l > 0.0);
        m = __popc(keep);
        const int src = __fns(keep, 0, lane + 1);
        const int si = __shfl_sync(full, wi, src & 31);
        const double sl = __shfl_sync(full, l, src & 31);
        if (lane < m) { pos = Z.iperm[si]; lam = sl; }
    }
    // g = [z - R lam]_+ , v = R g
    {
        double pi[NJ];
        if (m > 0) {
            double x[NJ];
            scatter_rows<NJ>(m, pos, lam, x, sm);
            tree_product<NJ>(Z, x, pi, sm);
        } else {
#pragma unroll
            for (int k = 0; k < NJ; ++k) pi[k] = 0.0;
        }
#pragma unroll
        for (int k = 0; k < NJ; ++k) gj[k] = fmax(zj[k] - pi[k], 0.0);
    }
    tree_product<NJ>(Z, gj, v, sm);

    double flops = 0.0;
    int its_total = 0;
    bool changed = false, give_up = false;
    unsigned inw = 0;                                      // bit k: position p0 + k is a working row
    for (int a = 0; a < m; ++a) {
        const int pa = __shfl_sync(full, pos, a);
        if (pa / NJ == lane) inw |= 1u << (pa % NJ);
    }

    for (int pass = 0;; ++pass) {
        // ---- violated rows outside W, most violated first (ties: lowest position)
        int added = 0;
        bool more = false;
        const int room = min(kAddMaxT, kWW - m);
#pragma unroll 1
        for (int r = 0; r <= room; ++r) {
            double best = -1.0;
            int bp = 0x7fffffff;
#pragma unroll
            for (int k = 0; k < NJ; ++k) {
                const double viol = v[k] - u;
                if (p0 + k < n && !((inw >> k) & 1u) && viol > tol && viol > best) { best = viol; bp = p0 + k; }
            }
            if (!__any_sync(full, best >= 0.0)) break;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(full, best, o);
                const int op = __shfl_xor_sync(full, bp, o);
                if (ob > best || (ob == best && op < bp)) { best = ob; bp = op; }
            }
            if (r == room) { more = true; break; }
            if (lane == m + added) { pos = bp; lam = 0.0; }
            if (bp / NJ == lane) inw |= 1u << (bp % NJ);
            ++added;
        }
        if (added == 0 && !more && (pass > 0 || m == 0)) break;       // KKT point of the whole column
        if ((more && m + added == kWW) || pass >= kPassMaxT) { give_up = true; break; }
        m += added;
        const bool row = lane < m;
        int ok = 0, its = 0;
        double grad = 0.0;

        if (m == 1) {
            // one row: v(l) = r . [z - r l]_+ is convex, piecewise linear, non-increasing; Newton on v(l) = u
            double r1[NJ];
            gen_row<NJ>(Z, __shfl_sync(full, pos, 0), r1);
            const double l_in = __shfl_sync(full, lam, 0);
            double l = l_in;
#pragma unroll 1
            for (int it = 0; it < 48; ++it) {
                double vv = 0.0, S = 0.0;
#pragma unroll
                for (int k = 0; k < NJ; ++k) {
                    const double gk = fmax(zj[k] - r1[k] * l, 0.0);
                    gj[k] = gk;
                    vv = fma(r1[k], gk, vv);
                    if (gk > 0.0) S = fma(r1[k], r1[k], S);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    vv += __shfl_xor_sync(full, vv, o);
                    S += __shfl_xor_sync(full, S, o);
                }
                ++its;
                const double fr = vv - u;
                if ((l > 0.0 ? fabs(fr) : fmax(fr, 0.0)) < tol) { ok = 1; break; }
                double ln = S > 0.0 ? l + fr / S : 0.0;
                if (ln < 0.0) ln = 0.0;
                if (ln == l) break;                        // stagnation: the general path decides
                l = ln;
            }
            flops += 4.0 * n * its;
            lam = lane == 0 ? l : 0.0;
            if (l != l_in) changed = true;
            if (!ok) its = 0;
        }

        if (!ok) {
            // ---- piecewise-quadratic descent on W; every matrix-vector product is a tree product
            const double scale = warp_sum(row ? P.rn2[zo + Z.perm[pos]] : 0.0) / (double)max(m, 1);
            const double shift = kHessShiftT * scale + 1e-300;
            double phi;
            {
                double acc = 0.0;
#pragma unroll
                for (int k = 0; k < NJ; ++k) acc = fma(gj[k], gj[k], acc);
                phi = 0.5 * warp_sum(acc) + u * warp_sum(row ? lam : 0.0);
            }
            bool bail = false;
#pragma unroll 1
            for (; its < P.inner_max; ++its) {
                // gradient on W from the exact voltages of the current g
                {
                    double vv[NJ];
                    tree_product<NJ>(Z, gj, vv, sm);
                    grad = u - gather_rows<NJ>(m, pos, vv, sm);
                }
                flops += 6.0 * n;
                const double kk = row ? fabs(lam > 0.0 ? grad : fmin(grad, 0.0)) : 0.0;
                if (warp_max(kk) < tol) { ok = 1; break; }

                // Hessian of the current piece: column q = R (row_q restricted to {g > 0}) at the working rows
#pragma unroll 1
                for (int q = 0; q < m; ++q) {
                    double x[NJ], y[NJ];
                    gen_row<NJ>(Z, __shfl_sync(full, pos, q), x);
#pragma unroll
                    for (int k = 0; k < NJ; ++k)
                        if (!(gj[k] > 0.0)) x[k] = 0.0;
                    tree_product<NJ>(Z, x, y, sm);
                    const double h = gather_rows<NJ>(m, pos, y, sm);
                    if (row) sm.H[lane * kTH + q] = h;
                }
                flops += 8.0 * n * m;
                __syncwarp();
                if (row) {                                 // exact symmetry for the factorisation: mirror the lower triangle
                    for (int q = lane + 1; q < m; ++q) sm.H[lane * kTH + q] = sm.H[q * kTH + lane];
                }
                __syncwarp();

                // ---- exact minimiser of the piece over lam_W >= 0: primal-dual active set, lanes = rows
                double b = 0.0;
#pragma unroll 1
                for (int q = 0; q < m; ++q) {
                    const double lq = __shfl_sync(full, lam, q);
                    if (row && lq != 0.0) b = fma(sm.H[lane * kTH + q], lq, b);
                }
                b += shift * lam - grad;
                bool inA = row && (lam > 0.0 || grad < 0.0);
                double xs_all = 0.0;
                bool pdas_ok = false;
#pragma unroll 1
                for (int guess = 0; guess < kPdasMaxT; ++guess) {
                    const unsigned Am = __ballot_sync(full, inA);
                    const int ma = __popc(Am);
                    const int cpos = __popc(Am & ((1u << lane) - 1));
                    double xs = 0.0;
                    if (ma > 0) {
                        const int o = (lane < ma) ? (int)__fns(Am, 0, lane + 1) : 0;     // original row of compact row `lane`
#pragma unroll 1
                        for (int cidx = 0; cidx < ma; ++cidx) {
                            const int oc = __shfl_sync(full, o, cidx);
                            if (lane < ma && cidx <= lane) sm.L[lane * kTH + cidx] = sm.H[o * kTH + oc] + (cidx == lane ? shift : 0.0);
                        }
                        __syncwarp();
                        double rdiag = 1.0;
#pragma unroll 1
                        for (int k2 = 0; k2 < ma; ++k2) {            // Cholesky, left-looking: lanes own rows
                            double sv = 0.0;
                            if (lane >= k2 && lane < ma) {
                                sv = sm.L[lane * kTH + k2];
                                for (int p2 = 0; p2 < k2; ++p2) sv = fma(-sm.L[lane * kTH + p2], sm.L[k2 * kTH + p2], sv);
                            }
                            const double skk = fmax(__shfl_sync(full, sv, k2), 1e-300);
                            const double rk = rsqrt(skk);
                            const double dkk = skk * rk;
                            if (lane == k2) rdiag = rk;
                            if (lane >= k2 && lane < ma) sm.L[lane * kTH + k2] = lane == k2 ? dkk : sv * rk;
                            __syncwarp();
                        }
                        double y = __shfl_sync(full, b, o);
                        if (lane >= ma) y = 0.0;
#pragma unroll 1
                        for (int k2 = 0; k2 < ma; ++k2) {
                            const double yk = __shfl_sync(full, y, k2) * __shfl_sync(full, rdiag, k2);
                            if (lane == k2) y = yk;
                            if (lane > k2 && lane < ma) y = fma(-sm.L[lane * kTH + k2], yk, y);
                        }
#pragma unroll 1
                        for (int k2 = ma - 1; k2 >= 0; --k2) {
                            const double xk = __shfl_sync(full, y, k2) * __shfl_sync(full, rdiag, k2);
                            if (lane == k2) y = xk;
                            if (lane < k2) y = fma(-sm.L[k2 * kTH + lane], xk, y);
                        }
                        xs = y;
                        flops += (2.0 / 3.0) * ma * ma * ma + 4.0 * ma * ma + 2.0 * m * ma;
                    }
                    const double xg = __shfl_sync(full, xs, cpos & 31);
                    xs_all = inA ? xg : 0.0;
                    double mu = 0.0;
#pragma unroll 1
                    for (int q = 0; q < m; ++q) {
                        const double xq = __shfl_sync(full, xs_all, q);
                        if (row && xq != 0.0) mu = fma(sm.H[lane * kTH + q], xq, mu);
                    }
                    mu -= b;
                    const bool bad = row && (inA ? (xs_all <= 0.0) : (mu < 0.0));
                    if (!__any_sync(full, bad)) { pdas_ok = true; break; }
                    if (bad) inA = !inA;
                }
                if (!pdas_ok) { bail = true; break; }

                // ---- line search of phi on the segment lam -> minimiser
                const double dir = xs_all - lam;
                double gt[NJ];
                double alpha = 1.0, phin = phi, lt = lam;
                bool stepped = false;
#pragma unroll 1
                for (; alpha >= kArcMinT; alpha *= 0.5) {
                    lt = row ? fmax(fma(alpha, dir, lam), 0.0) : 0.0;
                    {
                        double x[NJ], pi[NJ];
                        scatter_rows<NJ>(m, pos, lt, x, sm);
                        tree_product<NJ>(Z, x, pi, sm);
                        double acc = 0.0;
#pragma unroll
                        for (int k = 0; k < NJ; ++k) {
                            gt[k] = fmax(zj[k] - pi[k], 0.0);
                            acc = fma(gt[k], gt[k], acc);
                        }
                        phin = 0.5 * warp_sum(acc) + u * warp_sum(row ? lt : 0.0);
                    }
                    flops += 6.0 * n;
                    const double slope = warp_sum(row ? grad * (lt - lam) : 0.0);
                    if (phin <= phi + 1e-4 * slope + 1e-14 * fabs(phi)) { stepped = true; break; }
                }
                if (!stepped) { bail = true; break; }
                lam = lt;
                phi = phin;
                changed = true;
#pragma unroll
                for (int k = 0; k < NJ; ++k) gj[k] = gt[k];
            }
            if (bail || !ok) { give_up = true; break; }
        }
        its_total += its;

        // ---- rows whose multiplier went to zero leave W; exact voltages of all rows for the new g
        {
            const unsigned keep = __ballot_sync(full, row && lam > 0.0);
            const int mk = __popc(keep);
            if (mk != m) {
                const int src = __fns(keep, 0, lane + 1);
                const int sp = __shfl_sync(full, pos, src & 31);
                const double sl = __shfl_sync(full, lam, src & 31);
                m = mk;
                pos = lane < m ? sp : 0;
                lam = lane < m ? sl : 0.0;
                inw = 0;
                for (int a = 0; a < m; ++a) {
                    const int pa = __shfl_sync(full, pos, a);
                    if (pa / NJ == lane) inw |= 1u << (pa % NJ);
                }
            }
        }
        tree_product<NJ>(Z, gj, v, sm);
        flops += 6.0 * n;
    }

    ++st.cols;
    if (give_up) { ++st.left; return; }                    // nothing was written: the dense path redoes the column
    // ---- persist
    if (changed || m != m_old || m > 0) {
        if (lane < m_old) lam_g[wi] = 0.0;
        __syncwarp();
        if (lane < m) {
            const int h = Z.perm[pos];
            lam_g[h] = lam;
            widx[lane] = h;
        }
        double* g = P.g_t + col;
#pragma unroll
        for (int k = 0; k < NJ; ++k)
            if (hk[k] >= 0) g[hk[k]] = gj[k];
    }
    if (lane == 0) {
        P.wcount[c] = m;
        P.inner_ok[c] = 1;
        P.status[c] = 1;
    }
    st.its += (unsigned long long)its_total;
    st.flops += flops;
    st.max_ws = max(st.max_ws, m);
}

template <int NJ>
__global__ void __launch_bounds__(32 * kTWarps, TreeCfg<NJ>::kCtas) tree_qp_kernel(QpParams P, TreeParams TP, const int* __restrict__ cols,
                                                                                  int ncols, int* __restrict__ queue) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using Smem = TreeSmem<NJ>;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    Smem& sm = reinterpret_cast<Smem*>(smem_raw)[wib];
    TreeStats st;
    for (;;) {
        int slot = 0;
        if (lane == 0) slot = atomicAdd(queue, 1);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (slot >= ncols) break;
        tree_column<NJ>(P, TP, cols[slot], sm, st);
        __syncwarp();
    }
    if (lane == 0) {
        if (st.its) atomicAdd(P.newton_its, st.its);
        if (st.flops > 0.0) atomicAdd(P.flops, (unsigned long long)st.flops);
        if (st.max_ws) atomicMax(P.max_ws, st.max_ws);
        if (st.cols) atomicAdd(P.cols, (unsigned long long)st.cols);
        if (st.left) atomicAdd(TP.left, st.left);
    }
}

template <int NJ>
cudaError_t prepare_tree_nj(int* n_sm_out) {
    static int n_sm[64] = {0};
    const int smem = (int)sizeof(TreeSmem<NJ>) * kTWarps;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    dev &= 63;
    if (!n_sm[dev]) {
        int n = 0;
        e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tree_qp_kernel<NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        n_sm[dev] = n;
    }
    if (n_sm_out) *n_sm_out = n_sm[dev];
    return cudaSuccess;
}

template <int NJ>
cudaError_t launch_tree_nj(const QpParams& P, const TreeParams& TP, const int* cols, int ncols, int* queue, cudaStream_t stream) {
    int n_sm = 0;
    cudaError_t e = prepare_tree_nj<NJ>(&n_sm);
    if (e != cudaSuccess) return e;
    const int smem = (int)sizeof(TreeSmem<NJ>) * kTWarps;
    tree_qp_kernel<NJ><<<n_sm * TreeCfg<NJ>::kCtas, 32 * kTWarps, smem, stream>>>(P, TP, cols, ncols, queue);
    return cudaGetLastError();
}

}  // namespace

// captured loop: the working-set rounds of the dense kernels run only if the tree kernels left columns behind
__global__ void tree_gate_kernel(const int* __restrict__ left, unsigned long long cond_round) {
    if (threadIdx.x == 0) cudaGraphSetConditional((cudaGraphConditionalHandle)cond_round, *left > 0 ? 1u : 0u);
}

cudaError_t launch_tree_gate(const int* left, unsigned long long cond_round, cudaStream_t stream) {
    tree_gate_kernel<<<1, 32, 0, stream>>>(left, cond_round);
    return cudaGetLastError();
}

cudaError_t tree_qp_prepare() {
    cudaError_t e = prepare_tree_nj<4>(nullptr);
    if (e == cudaSuccess) e = prepare_tree_nj<6>(nullptr);
    if (e == cudaSuccess) e = prepare_tree_nj<8>(nullptr);
    if (e == cudaSuccess) e = prepare_tree_nj<10>(nullptr);
    return e;
}

int tree_qp_group(int n) { return n <= 128 ? 0 : (n <= 192 ? 1 : (n <= 256 ? 2 : (n <= kWarpMaxN ? 3 : -1))); }

cudaError_t launch_tree_qp(const QpParams& P, const TreeParams& TP, int group, const int* cols, int ncols, int* queue, cudaStream_t stream) {
    if (ncols <= 0) return cudaSuccess;
    switch (group) {
        case 0: return launch_tree_nj<4>(P, TP, cols, ncols, queue, stream);
        case 1: return launch_tree_nj<6>(P, TP, cols, ncols, queue, stream);
        case 2: return launch_tree_nj<8>(P, TP, cols, ncols, queue, stream);
        case 3: return launch_tree_nj<10>(P, TP, cols, ncols, queue, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace revs
