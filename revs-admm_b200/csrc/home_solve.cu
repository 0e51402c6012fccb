// Batched per-consumer charging sub-problem: one warp per home.
//
// Reference: class Home (lpsolver.py:44-160) and solve_residence (lpsolver.py:430-460),
// one Gurobi MIQP per home per ADMM iteration.  With p[t] = e[t]*rating, e binary, the
// objective is separable and linear in e, and the SOC rows collapse to a window on the
// number of charging hours, so the exact optimum is a selection: the n_min cheapest
// hours of the plug-in window, then further hours while their cost is negative (up to
// n_max); hour costs are compared on a 2^-20 grid and ties go to the earliest hour.  Lanes hold the hours (t = lane + 32 j); hours are
// picked with warp-shuffle arg-min rounds (or a full shuffle ranking when many hours are
// needed), no sort and no shared memory.
//
// Besides the schedule the kernel leaves, per home, the sum over the hours of (new schedule -
// previous schedule)^2 -- the dual residual of the ADMM loop, which dual_update_kernel then only
// adds up (the previous schedule is the zero start in iteration 0, the load outside the plug-in
// window, and the value the cost evaluation reads anyway inside it).
//
// The hour cost is evaluated with individually rounded operations (__dmul_rn/__dadd_rn)
// in the order oracle/revs_oracle.py:home_delta uses, so that the selection is
// bit-identical with the CPU oracle whenever the inputs are.
#include <math_constants.h>

#include "kernels.cuh"

namespace revs {

constexpr int kMaxSlots = 8;   // hours per lane -> T <= 256
constexpr int kSelectRounds = 24;   // up to this many charging hours: iterative arg-min, else full ranking


template <int SLOTS>
__global__ void __launch_bounds__(256, SLOTS <= 3 ? 6 : 1) home_solve_kernel(HomeParams P) {
    const int lane = threadIdx.x & 31;
    const int h = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (h >= P.Hp) return;
    const size_t base = (size_t)h * P.T;
    const bool ev = P.has_ev[h] != 0;
    // (the per-home parameters travel together with the EV flag: one round trip instead of two for an EV home)
    const double rate = P.rating[h];
    const int st = P.start[h], en = P.end[h];
    const int nmin = P.n_min[h], nmax = P.n_max[h];
    // ping-pong of the schedules by the device iteration counter (see HomeParams::iter)
    const bool odd = P.iter != nullptr && (*P.iter & 1);
    const double* __restrict__ p_sch_in = odd ? P.p_sch_new : P.p_sch;
    double* __restrict__ p_sch_out = odd ? const_cast<double*>(P.p_sch) : P.p_sch_new;

    double ld[SLOTS];
#pragma unroll
    for (int j = 0; j < SLOTS; ++j) {
        int t = lane + 32 * j;
        ld[j] = t < P.T ? P.load[base + t] : 0.0;
    }
    // dual residual of the ADMM loop: sum over the hours of (new schedule - previous schedule)^2, for dual_update_kernel.
    // The previous schedule is the zero start in the first iteration, otherwise what this kernel wrote last time: the
    // load outside the plug-in window (difference exactly 0) and the value read below inside it.
    const bool first = P.iter ? (*P.iter == 0) : (P.first != 0);
    if (!ev) {   // warp-uniform: a home without EV only moves its load through
        double a2 = 0.0;
#pragma unroll
        for (int j = 0; j < SLOTS; ++j) {
            int t = lane + 32 * j;
            if (t < P.T) {
                p_sch_out[base + t] = ld[j];
                P.p_ev[base + t] = 0.0;
                const double ds = first ? ld[j] : 0.0;
                a2 = fma(ds, ds, a2);
            }
        }
        if (P.dsum) {
            a2 = warp_sum(a2);
            if (lane == 0) P.dsum[h] = a2;
        }
        return;
    }

    const double kap = P.kappa;
    const double c0 = __dmul_rn(__dmul_rn(0.5 * kap, rate), rate);

    double d[SLOTS], prev[SLOTS];
#pragma unroll
    for (int j = 0; j < SLOTS; ++j) {
        int t = lane + 32 * j;
        double v = CUDART_INF;
        prev[j] = first ? 0.0 : ld[j];
        if (t < P.T && t >= st && t < en) {
            if (P.individual) {
                // (0.01*c_t)*rating - 0.99*(rating/capacity)   (lpsolver.py:407-415)
                v = __dadd_rn(__dmul_rn(__dmul_rn(0.01, P.cost[t]), rate), P.ind_const[h]);
            } else {
                const double ps = p_sch_in[base + t];
                prev[j] = ps;
                double s = __dadd_rn(P.p_est[base + t], ps);
                double a = __dadd_rn(P.gamma[base + t], __dmul_rn(0.5 * kap, s));
                double x = __dmul_rn(rate, __dadd_rn(P.cost[t], -a));
                double y = __dmul_rn(__dmul_rn(kap, ld[j]), rate);
                v = __dadd_rn(__dadd_rn(x, y), c0);
            }
        }
        // hour costs are compared on a grid of 2^-20 (oracle/revs_oracle.py:TIE_GRID): the scale is a
        // power of two and rint rounds once, so the key is bit-identical with the oracle's, and the
        // ~1e-12 noise of the utility QP cannot reorder two hours that tie mathematically
        d[j] = v < CUDART_INF ? rint(v * 1048576.0) : v;
    }

    int in_window = 0;
#pragma unroll
    for (int j = 0; j < SLOTS; ++j) in_window += __popc(__ballot_sync(0xffffffffu, d[j] < CUDART_INF));
    if (nmin > nmax || nmin > in_window) {
        if (lane == 0) atomicExch(P.infeasible, 1);
    }

    bool on[SLOTS];
#pragma unroll
    for (int j = 0; j < SLOTS; ++j) on[j] = false;
    if (nmax <= kSelectRounds) {
        // few charging hours (the usual case): pull the cheapest remaining hour out of the
        // warp nmax times.  Costs become order-preserving 64-bit integer keys and every lane
        // keeps its own hours sorted (ties: earlier hour first), so a pick is two hardware warp
        // reductions (REDUX) on the lanes' current heads plus a ballot; only if two lanes hold
        // the same cost does a third reduction on the hour break the tie.
        unsigned long long key[SLOTS];
        unsigned slots = 0;                       // nibble i: slot index of the i-th smallest key of this lane
#pragma unroll
        for (int j = 0; j < SLOTS; ++j) {
            const unsigned long long b = (unsigned long long)__double_as_longlong(d[j] + 0.0);   // -0 -> +0
            key[j] = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
            slots |= (unsigned)j << (4 * j);
        }
#pragma unroll
        for (int i = 1; i < SLOTS; ++i)            // insertion sort, stable (strict <)
#pragma unroll
            for (int jj = i; jj >= 1; --jj) {
                const bool sw = key[jj] < key[jj - 1];
                const unsigned long long lo_k = sw ? key[jj] : key[jj - 1], hi_k = sw ? key[jj - 1] : key[jj];
                key[jj - 1] = lo_k;
                key[jj] = hi_k;
                if (sw) {
                    const unsigned a = (slots >> (4 * jj)) & 15u, c = (slots >> (4 * (jj - 1))) & 15u;
                    slots = (slots & ~((15u << (4 * jj)) | (15u << (4 * (jj - 1))))) | (c << (4 * jj)) | (a << (4 * (jj - 1)));
                }
            }
        unsigned onmask = 0;
        for (int cnt = 0; cnt < nmax; ++cnt) {
            // the high word (sign, exponent, 20 mantissa bits) of the cheapest head almost always
            // identifies the winner on its own; low word and hour are only consulted on ties
            const unsigned long long head = key[0];
            const unsigned hw = (unsigned)(head >> 32);
            const unsigned hi = __reduce_min_sync(0xffffffffu, hw);
            if (hi >= 0xFFF00000u) break;                             // window exhausted (+inf)
            if (cnt >= nmin && hi >= 0x80000000u) break;              // optional hours only while cost < 0
            unsigned tied = __ballot_sync(0xffffffffu, hw == hi);
            int wl = __ffs(tied) - 1;
            if (tied & (tied - 1)) {
                const unsigned lo = __reduce_min_sync(0xffffffffu, hw == hi ? (unsigned)head : 0xffffffffu);
                const bool mine = hw == hi && (unsigned)head == lo;
                tied = __ballot_sync(0xffffffffu, mine);
                wl = __ffs(tied) - 1;
                if (tied & (tied - 1)) {                              // same cost in several lanes: earliest hour wins
                    const unsigned myt = mine ? (unsigned)(lane + 32 * (int)(slots & 15u)) : 0x7fffffffu;
                    wl = (int)(__reduce_min_sync(0xffffffffu, myt) & 31u);
                }
            }
            if (lane == wl) {
                onmask |= 1u << (slots & 15u);
                slots >>= 4;
#pragma unroll
                for (int j = 0; j + 1 < SLOTS; ++j) key[j] = key[j + 1];
                key[SLOTS - 1] = ~0ull;
            }
        }
#pragma unroll
        for (int j = 0; j < SLOTS; ++j) on[j] = (onmask >> j) & 1u;
    } else {
        // rank[j] = #{hours s : d_s < d_t  or (d_s == d_t and s < t)}
        int rank[SLOTS];
#pragma unroll
        for (int j = 0; j < SLOTS; ++j) rank[j] = 0;
#pragma unroll
        for (int js = 0; js < SLOTS; ++js) {
            for (int src = 0; src < 32; ++src) {
                double o = __shfl_sync(0xffffffffu, d[js], src);
                int s = src + 32 * js;
#pragma unroll
                for (int j = 0; j < SLOTS; ++j) {
                    int t = lane + 32 * j;
                    rank[j] += (o < d[j]) || (o == d[j] && s < t);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < SLOTS; ++j)
            on[j] = d[j] < CUDART_INF && (rank[j] < nmin || (rank[j] < nmax && d[j] < 0.0));
    }
    double a2 = 0.0;
#pragma unroll
    for (int j = 0; j < SLOTS; ++j) {
        int t = lane + 32 * j;
        if (t >= P.T) continue;
        double p = on[j] ? rate : 0.0;
        P.p_ev[base + t] = p;
        const double sn = __dadd_rn(ld[j], p);
        p_sch_out[base + t] = sn;
        const double ds = sn - prev[j];
        a2 = fma(ds, ds, a2);
    }
    if (P.dsum) {
        a2 = warp_sum(a2);
        if (lane == 0) P.dsum[h] = a2;
    }
}

cudaError_t launch_home_solve(const HomeParams& P, cudaStream_t stream) {
    const int warps = 8;
    dim3 grid((P.Hp + warps - 1) / warps), block(32 * warps);
    int slots = (P.T + 31) / 32;
    if (slots <= 1) home_solve_kernel<1><<<grid, block, 0, stream>>>(P);
    else if (slots <= 3) home_solve_kernel<3><<<grid, block, 0, stream>>>(P);
    else if (slots <= kMaxSlots) home_solve_kernel<kMaxSlots><<<grid, block, 0, stream>>>(P);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

// SOC[h][t+1] = SOC[h][t] + p_ev[h][t]/capacity  (lpsolver.py:105-108), one thread per home
// hour would need a scan; T is tiny, so one lane walks a home and a warp covers 32 homes.
__global__ void soc_profile_kernel(const double* __restrict__ p_ev, const uint8_t* __restrict__ has_ev,
                                   const double* __restrict__ capacity,
                                   const double* __restrict__ initial, double* __restrict__ soc,
                                   int Hp, int T) {
    int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= Hp) return;
    double* s = soc + (size_t)h * (T + 1);
    if (!has_ev[h]) {
        for (int t = 0; t <= T; ++t) s[t] = 0.0;
        return;
    }
    const double cap = capacity[h];
    double acc = initial[h];
    s[0] = acc;
    for (int t = 0; t < T; ++t) {
        acc = __dadd_rn(acc, __ddiv_rn(p_ev[(size_t)h * T + t], cap));
        s[t + 1] = acc;
    }
}

cudaError_t launch_soc_profile(const double* p_ev, const uint8_t* has_ev, const double* capacity,
                               const double* initial, double* soc, int Hp, int T,
                               cudaStream_t stream) {
    soc_profile_kernel<<<(Hp + 127) / 128, 128, 0, stream>>>(p_ev, has_ev, capacity, initial, soc, Hp, T);
    return cudaGetLastError();
}

}  // namespace revs
