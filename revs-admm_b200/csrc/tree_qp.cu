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
    double X[32 * NJ + 1];                   // argument of tree_product / result of gen_row   (slot layout, see CtaZone)
    double V[32 * NJ + 1];                   // result of tree_product; inside it: the prefix sums of X (slot 32 NJ holds 0)
    double S1[32 * NJ + 1];                  // scratch of tree_product (slot 32 NJ holds 0)
    double S2[32 * NJ + 1];
    double H[kWW * kTH];
    double L[kWW * kTH];
};

template <int NJ> struct TreeCfg { static constexpr int kCtas = NJ <= 4 ? 4 : 3; };

// Static arrays of one zone, staged in shared memory once per CTA and chunk of columns (all four warps of a CTA work
// on columns of the same zone).  Position p (lane p / NJ, slot p % NJ) is stored at slot(p) = (p % NJ) * 32 + p / NJ,
// so the k-th element of every lane is one conflict-free 32-wide access; the Cartesian-tree nodes and the counts carry
// ready-made slots into the per-warp arrays (the extra slot 32 NJ holds 0).
template <int NJ>
struct CtaZone {
    double c[32 * NJ];    // [slot(p)] c[p] = 2 cumr(lca(p, p+1)), 0 beyond the zone
    double d[32 * NJ];    // R[p][p]
    double e[32 * NJ];    // d[p] - max(c[p-1], c[p])
    double wA[32 * NJ];   // weight of the k-th Cartesian-tree node in lo-order, at slot(k)
    int perm[32 * NJ];    // home index of position p, -1 beyond the zone
    int nodeA[32 * NJ];   // k-th node in lo-order: slot of G[hi] | slot of G[lo - 1] << 16
    int permB[32 * NJ];   // k-th node in hi-order: slot of that node in lo-order
    int cnt[32 * NJ];     // slot of S1[#lo <= p] | slot of S2[#hi < p] << 16
};
template <int NJ> __device__ __forceinline__ int slot_of(int p) { return (p % NJ) * 32 + p / NJ; }

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

// sm.V = R sm.X for every row of the zone (see the header).  ONE copy of this code per kernel (not inlined: with the
// products inlined at their seven call sites the kernel outgrew the instruction cache and stalled on fetches);
// vectors travel through shared memory in the slot layout.
template <int NJ>
__device__ __noinline__ void tree_product(const CtaZone<NJ>* Zp, TreeSmem<NJ>* smp) {
    const CtaZone<NJ>& Z = *Zp;
    TreeSmem<NJ>& sm = *smp;
    double* G = sm.V;
    const int lane = threadIdx.x & 31;
    double t[NJ];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < NJ; ++k) t[k] = sm.X[k * 32 + lane];
    scan_incl<NJ>(t);
#pragma unroll
    for (int k = 0; k < NJ; ++k) G[k * 32 + lane] = t[k];             // G[slot(p)] = sum of x over positions <= p
    __syncwarp();
#pragma unroll
    for (int k = 0; k < NJ; ++k) {                                     // node weights times subtree sums, nodes in lo-order
        const int nd = Z.nodeA[k * 32 + lane];
        t[k] = Z.wA[k * 32 + lane] * (G[nd & 0xffff] - G[nd >> 16]);
        sm.S2[k * 32 + lane] = t[k];
    }
    scan_incl<NJ>(t);
#pragma unroll
    for (int k = 0; k < NJ; ++k) sm.S1[k * 32 + lane] = t[k];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < NJ; ++k) t[k] = sm.S2[Z.permB[k * 32 + lane]];  // the same products, nodes in hi-order
    __syncwarp();
    scan_incl<NJ>(t);
#pragma unroll
    for (int k = 0; k < NJ; ++k) sm.S2[k * 32 + lane] = t[k];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < NJ; ++k) {
        const int cn = Z.cnt[k * 32 + lane];
        sm.V[k * 32 + lane] = fma(Z.e[k * 32 + lane], sm.X[k * 32 + lane], sm.S1[cn & 0xffff] - sm.S2[cn >> 16]);
    }
    __syncwarp();
}

// sm.X = row i of R (depth-first positions): min of c over the positions between i and j, d[i] on the diagonal
template <int NJ>
__device__ __noinline__ void gen_row(const CtaZone<NJ>* Zp, int i, TreeSmem<NJ>* smp) {
    const CtaZone<NJ>& Z = *Zp;
    TreeSmem<NJ>& sm = *smp;
    const int lane = threadIdx.x & 31, p0 = lane * NJ;
    double cq[NJ], right[NJ], left[NJ];
#pragma unroll
    for (int k = 0; k < NJ; ++k) cq[k] = Z.c[k * 32 + lane];          // beyond the zone: 0 -> padded entries are 0
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
    const double di = Z.d[slot_of<NJ>(i)];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < NJ; ++k) {
        const int p = p0 + k;
        sm.X[k * 32 + lane] = p > i ? fmin(offr, right[k]) : (p < i ? fmin(offl, left[k]) : di);
    }
    __syncwarp();
}

// sm.X[p] = val_a at the positions pos_a of the working rows (lanes a < m), 0 elsewhere
template <int NJ>
__device__ __forceinline__ void scatter_rows(int m, int pos, double val, TreeSmem<NJ>& sm) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < NJ; ++k) sm.X[k * 32 + lane] = 0.0;
    __syncwarp();
    if (lane < m) sm.X[slot_of<NJ>(pos)] = val;
}

template <int NJ>
__device__ __forceinline__ void store_x(const double (&x)[NJ], TreeSmem<NJ>& sm) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < NJ; ++k) sm.X[k * 32 + lane] = x[k];
}

struct TreeStats {
    unsigned long long its = 0;
    double flops = 0.0;
    int left = 0, max_ws = 0, cols = 0;
};

template <int NJ>
__device__ __forceinline__ void tree_column(const QpParams& P, const TreeParams& TP, const int c, const CtaZone<NJ>* Zp, TreeSmem<NJ>& sm, TreeStats& st) {
    const int lane = threadIdx.x & 31, p0 = lane * NJ;
    const unsigned full = 0xffffffffu;
    const int f = c / P.T, t = c % P.T;
    const FeederDev fd = P.feeders[f];
    const int n = fd.n;
    const size_t zo = (size_t)fd.off;
    const CtaZone<NJ>& Z = *Zp;
    const CtaZone<NJ>* Zc = Zp;
    const int* __restrict__ iperm = TP.iperm + zo;
    if (lane == 0) { sm.V[32 * NJ] = 0.0; sm.S1[32 * NJ] = 0.0; sm.S2[32 * NJ] = 0.0; }     // the zero slot of the products
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
        hk[k] = Z.perm[k * 32 + lane];
        zj[k] = (p < n && hk[k] >= 0) ? z[hk[k]] : 0.0;
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
        if (lane < m) { pos = iperm[si]; lam = sl; }
    }
    // g = [z - R lam]_+ , v = R g
    if (m > 0) {
        scatter_rows<NJ>(m, pos, lam, sm);
        tree_product<NJ>(Zc, &sm);
#pragma unroll
        for (int k = 0; k < NJ; ++k) gj[k] = fmax(zj[k] - sm.V[k * 32 + lane], 0.0);
    } else {
#pragma unroll
        for (int k = 0; k < NJ; ++k) gj[k] = fmax(zj[k], 0.0);
    }
    store_x<NJ>(gj, sm);
    tree_product<NJ>(Zc, &sm);
#pragma unroll
    for (int k = 0; k < NJ; ++k) v[k] = sm.V[k * 32 + lane];

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
            gen_row<NJ>(Zc, __shfl_sync(full, pos, 0), &sm);
#pragma unroll
            for (int k = 0; k < NJ; ++k) r1[k] = sm.X[k * 32 + lane];
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
            const double scale = warp_sum(row ? P.rn2[zo + Z.perm[slot_of<NJ>(pos)]] : 0.0) / (double)max(m, 1);
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
                store_x<NJ>(gj, sm);
                tree_product<NJ>(Zc, &sm);
                grad = row ? u - sm.V[slot_of<NJ>(pos)] : 0.0;
                flops += 6.0 * n;
                const double kk = row ? fabs(lam > 0.0 ? grad : fmin(grad, 0.0)) : 0.0;
                if (warp_max(kk) < tol) { ok = 1; break; }

                // Hessian of the current piece: column q = R (row_q restricted to {g > 0}) at the working rows
#pragma unroll 1
                for (int q = 0; q < m; ++q) {
                    gen_row<NJ>(Zc, __shfl_sync(full, pos, q), &sm);
#pragma unroll
                    for (int k = 0; k < NJ; ++k)
                        if (!(gj[k] > 0.0)) sm.X[k * 32 + lane] = 0.0;
                    tree_product<NJ>(Zc, &sm);
                    if (row) sm.H[lane * kTH + q] = sm.V[slot_of<NJ>(pos)];
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
                        scatter_rows<NJ>(m, pos, lt, sm);
                        tree_product<NJ>(Zc, &sm);
                        double acc = 0.0;
#pragma unroll
                        for (int k = 0; k < NJ; ++k) {
                            gt[k] = fmax(zj[k] - sm.V[k * 32 + lane], 0.0);
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
        store_x<NJ>(gj, sm);
        tree_product<NJ>(Zc, &sm);
#pragma unroll
        for (int k = 0; k < NJ; ++k) v[k] = sm.V[k * 32 + lane];
        flops += 6.0 * n;
    }

    ++st.cols;
    if (give_up) { ++st.left; return; }                    // nothing was written: the dense path redoes the column
    // ---- persist
    if (changed || m != m_old || m > 0) {
        if (lane < m_old) lam_g[wi] = 0.0;
        __syncwarp();
        if (lane < m) {
            const int h = Z.perm[slot_of<NJ>(pos)];
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

// Work comes in CHUNKS of columns of one zone (int2 {zone, first hour | count << 16}, built on the host): the CTA stages
// the zone's static arrays in shared memory once per chunk (skipped when the zone is the one already staged) and its
// four warps take the chunk's columns round-robin.
template <int NJ>
__global__ void __launch_bounds__(32 * kTWarps, TreeCfg<NJ>::kCtas) tree_qp_kernel(QpParams P, TreeParams TP, const int2* __restrict__ chunks,
                                                                                  int nchunks, int* __restrict__ queue) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using Smem = TreeSmem<NJ>;
    CtaZone<NJ>& Z = *reinterpret_cast<CtaZone<NJ>*>(smem_raw);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    Smem& sm = reinterpret_cast<Smem*>(smem_raw + sizeof(CtaZone<NJ>))[wib];
    __shared__ int s_slot;
    TreeStats st;
    int staged = -1;
    for (;;) {
        __syncthreads();                                   // everybody is done with the staged zone and with s_slot
        if (threadIdx.x == 0) s_slot = atomicAdd(queue, 1);
        __syncthreads();
        const int slot = s_slot;
        if (slot >= nchunks) break;
        const int2 ch = chunks[slot];
        const int f = ch.x, t0 = ch.y & 0xffff, cnt = ch.y >> 16;
        if (f != staged) {
            const size_t so = (size_t)TP.zoff[f];
            for (int i = threadIdx.x; i < 32 * NJ; i += 32 * kTWarps) {
                Z.c[i] = TP.c[so + i]; Z.d[i] = TP.d[so + i]; Z.e[i] = TP.e[so + i]; Z.wA[i] = TP.wA[so + i];
                Z.perm[i] = TP.perm[so + i]; Z.nodeA[i] = TP.nodeA[so + i]; Z.permB[i] = TP.permB[so + i]; Z.cnt[i] = TP.cnt[so + i];
            }
            staged = f;
            __syncthreads();
        }
        for (int k = wib; k < cnt; k += kTWarps) {
            tree_column<NJ>(P, TP, f * P.T + t0 + k, &Z, sm, st);
            __syncwarp();
        }
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
    const int smem = (int)(sizeof(TreeSmem<NJ>) * kTWarps + sizeof(CtaZone<NJ>));
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
cudaError_t launch_tree_nj(const QpParams& P, const TreeParams& TP, const int2* cols, int ncols, int* queue, cudaStream_t stream) {
    int n_sm = 0;
    cudaError_t e = prepare_tree_nj<NJ>(&n_sm);
    if (e != cudaSuccess) return e;
    const int smem = (int)(sizeof(TreeSmem<NJ>) * kTWarps + sizeof(CtaZone<NJ>));
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

int tree_qp_chunk() { return 2 * kTWarps; }

int tree_qp_group(int n) { return n <= 128 ? 0 : (n <= 192 ? 1 : (n <= 256 ? 2 : (n <= kWarpMaxN ? 3 : -1))); }

cudaError_t launch_tree_qp(const QpParams& P, const TreeParams& TP, int group, const int2* cols, int ncols, int* queue, cudaStream_t stream) {
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
