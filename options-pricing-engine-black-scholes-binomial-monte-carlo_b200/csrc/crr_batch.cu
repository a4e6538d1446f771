// Batched Cox-Ross-Rubinstein backward induction (fp64) for sm_100a.
//
// Replaces reference binomial.py:7-38 (crr) looped over a contract batch and
// binomial.py:44-109 (crr_vec).  Recursion per level k = N-1..0 (binomial.py:28-36):
//     V_k[j] = disc*(p*V_{k+1}[j+1] + (1-p)*V_{k+1}[j]);  American: V_k[j] = max(V_k[j], +-(S_kj - K))
// with S_kj = S0*u^j*d^(k-j) = S0*u^(2j-k)  (d = 1/u, binomial.py:13).
//
// Kernel 1 (N <= 1023): ONE OPTION PER WARP, the whole level lives in registers.
//   Lane l owns the B contiguous nodes [l*B, l*B+B); a level update is B independent
//   (DFMA, DFMA, DSETP, 2 SEL) groups per lane plus ONE 64-bit shuffle for the neighbour's
//   first node, so neither shared memory nor a block barrier sits on the critical path.
//   Because level k has only k+1 live nodes, the warp re-packs its registers into a
//   narrower layout (B = 32, 28, 24, ... 1) through a padded shared-memory tile whenever
//   the live nodes fit; that keeps ~88 % of the issued node updates live (a fixed layout
//   would idle half the lanes on average).
//   Exercise values cost one DFMA per node: S_kj = base_k(l) * u^(2i), with u^(2i) (i < B)
//   held in registers and base_k(l) = S0*u^(2lB-k) advanced by one DMUL per lane per level.
// Kernel 2 (any N up to the shared-memory limit): one option per CTA, ping-pong level
//   arrays in shared memory.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "optp_common.cuh"

namespace optp {

constexpr int kWarpsPerCta = 4;
constexpr int kTileDoubles = 1024 + 40;  // padded re-pack tile per warp

struct CrrArgs {
    const double *S0, *K, *T, *r, *q, *sigma;
    const uint8_t* is_call;
    unsigned long long n;
    int N;
    double* price;
    uint8_t* status;
    int todo_only;  // 0 = every contract, else only contracts whose status equals this value (2 = declined by
                    // a banded kernel, 3 = American call that is never exercised early)
    // two-zone kernel work lists (compaction of the contracts whose band was too wide for a lane layout):
    // wl_in (wl_in_n items) = the contracts to walk instead of [0, n), or null; wl_out / wl_out_n = the list
    // the too-wide ones are appended to.  The counters sit together in one header (one memset per batch).
    const unsigned int* wl_in;
    const unsigned int* wl_in_n;
    unsigned int* wl_out;
    unsigned int* wl_out_n;
    // contracts whose exercise boundary fell faster than the segment schedule allows (the per-level bottom test
    // failed) are appended here and walked again by a pass on the RIGOROUS schedule (28-level segments: the
    // boundary cannot fall by more than one node per level), instead of going to the full-width kernel; or null
    unsigned int* wl_rig;
    unsigned int* wl_rig_n;
    // {number of contracts left at status 2, at status 3} by the two-zone passes, or null: a full-width
    // filter pass with nothing to do returns at once instead of scanning n status bytes
    unsigned int* decl_count;
    // two-zone kernel: ticket counter of this launch (zeroed by launch_crr).  A warp takes the next group of G
    // contracts when it has finished its last one, instead of a fixed stride: the walks differ in length (band
    // width, early all-exercise exits), and with a fixed stride the slowest warp set the kernel's time
    unsigned int* next_group;
    // one-option-per-CTA kernel: level arrays in global memory (3*(N+2) doubles per CTA) when they do not
    // fit in shared memory, else null
    double* cta_scratch;
    // diagnostics ("crr_count_ops" = 1): {FP64-pipe lane-instructions, node updates} executed inside the level
    // loops of the two-zone kernel, summed over the batch (two atomics per warp and option group), or null
    unsigned long long* op_count;
};

struct Lattice {
    double u, pu, pd, S0, sK, sgn;
    double rho, inv_pd, u_pd;  // pu/pd, 1/pd, u/pd (scaled recursion of the warp kernel)
    int R;                     // renormalisation interval
    bool ok;
};

// binomial.py:11-17
__device__ __forceinline__ Lattice make_lattice_of(double S0, double K, double T, double r, double q,
                                                   double sg, bool is_call, int N) {
    Lattice L;
    const double dt = T / (double)N;
    L.u = exp(sg * sqrt(dt));
    const double d = 1.0 / L.u;
    const double disc = exp(-r * dt);
    const double p = (exp((r - q) * dt) - d) / (L.u - d);
    // (a non-finite spot or strike also counts as a bad contract: it would otherwise come out as 0 from the
    // max(., 0) selects, where the reference's np.maximum propagates the NaN)
    L.ok = (p > 0.0) && (p < 1.0) && isfinite(S0) && isfinite(K);
    L.pu = disc * p;
    L.pd = disc * (1.0 - p);
    L.S0 = S0;
    L.sgn = is_call ? 1.0 : -1.0;
    L.sK = -L.sgn * K;
    L.inv_pd = 1.0 / L.pd;
    L.rho = L.pu * L.inv_pd;
    L.u_pd = L.u * L.inv_pd;
    const double al = fabs(log(L.pd));
    L.R = (al * 16777216.0 > 138.0) ? max(1, (int)(138.0 / al)) : (1 << 24);
    return L;
}

__device__ __forceinline__ Lattice make_lattice(const CrrArgs& a, unsigned long long o) {
    return make_lattice_of(a.S0[o], a.K[o], a.T[o], a.r[o], a.q[o], a.sigma[o], a.is_call[o] != 0, a.N);
}

// signed base of lane `lane` in layout B at level k:  sgn*S0*u^(2*lane*B - k)
__device__ __forceinline__ double lane_base(const Lattice& L, int lane, int B, int k) {
    return L.sgn * L.S0 * pow(L.u, (double)(2 * lane * B - k));
}

// max(c, e) for c >= 0 and finite-or-inf e: one DSETP + two SELs (fmax() would add the
// NaN-propagation sequence, ~6 more issue slots per node).
__device__ __forceinline__ double pick_max(double c, double e) { return e > c ? e : c; }

// Levels k_hi..k_lo in layout B.  To spend ONE DFMA on the continuation value the level is
// stored scaled: A_k = V_k * pd^-(m), m = levels walked since the last renormalisation, so
//     A_k[j] = (pu/pd)*A_{k+1}[j+1] + A_{k+1}[j]
// and the exercise value is scaled alike (two DMULs per LANE per level keep its coefficients).
// Every L.R levels (and at the end) the registers are multiplied back by pd^m; L.R bounds
// pd^-m by 2^+-200, so no contract can overflow.  Rounding differs from binomial.py:29 by
// O(1 ulp) per level -- far inside the 1e-10 parity bar.
// u^(2i), i < 32, as a two-level table: u^(2i) = hi[i >> 3] * lo[i & 7] (24 registers instead of
// 64; the freed registers are what lets ptxas overlap the exercise DFMA -> DSETP -> FSEL chains).
struct UTab {
    double lo[8], hi[4];
};

template <int B, bool AMER>
__device__ __forceinline__ void run_levels(double (&v)[32], const UTab& U, double sb,
                                           const Lattice& L, int k_hi, int k_lo) {
    int k = k_hi;
    while (k >= k_lo) {
        const int seg_lo = (k - L.R + 1 > k_lo) ? (k - L.R + 1) : k_lo;
        double sbg = sb, sKg = L.sK, back = 1.0;
        // the neighbour's first node of the NEXT level is requested as soon as this lane's
        // first node is updated, so the shuffle latency overlaps the other B-1 updates
        double nb = __shfl_down_sync(0xffffffffu, v[0], 1);
        for (; k >= seg_lo; --k) {
            sbg *= L.u_pd;
            sKg *= L.inv_pd;
            back *= L.pd;
            sb *= L.u;
            double nb_next = 0.0;
            double sa[4];
            if (AMER) {
#pragma unroll
                for (int a = 0; a < 4; ++a)
                    if (8 * a < B) sa[a] = sbg * U.hi[a];
            }
            // groups of G nodes: all continuation DFMAs, all exercise DFMAs, all compares, all
            // selects -- so the DFMA -> DSETP -> FSEL latencies of one node are covered by the
            // other nodes of the group instead of being exposed per node
            constexpr int G = (B >= 8) ? 8 : B;
#pragma unroll
            for (int g0 = 0; g0 < B; g0 += G) {
                double c[G], e[G];
#pragma unroll
                for (int t = 0; t < G; ++t) {
                    const int i = g0 + t;
                    if (i < B) {
                        const double up = (i + 1 < B) ? v[i + 1] : nb;
                        c[t] = fma(L.rho, up, v[i]);
                    }
                }
                if (AMER) {
#pragma unroll
                    for (int t = 0; t < G; ++t)
                        if (g0 + t < B) e[t] = fma(sa[(g0 + t) >> 3], U.lo[(g0 + t) & 7], sKg);
#pragma unroll
                    for (int t = 0; t < G; ++t)
                        if (g0 + t < B) c[t] = pick_max(c[t], e[t]);
                }
#pragma unroll
                for (int t = 0; t < G; ++t)
                    if (g0 + t < B) v[g0 + t] = c[t];
                if (g0 == 0) nb_next = __shfl_down_sync(0xffffffffu, c[0], 1);
            }
            nb = nb_next;
        }
#pragma unroll
        for (int i = 0; i < B; ++i) v[i] *= back;
    }
}

template <int B0, int B1>
__device__ __forceinline__ void repack(double (&v)[32], double* tile, int lane) {
    // keep the index arithmetic here (ptxas otherwise hoists ~100 addresses out of the
    // option loop and spills)
    asm volatile("" : "+r"(lane));
#pragma unroll
    for (int i = 0; i < B0; ++i) tile[lane * (B0 + 1) + i] = v[i];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < B1; ++i) {
        const int j = lane * B1 + i;
        v[i] = tile[j + j / B0];
    }
    __syncwarp();
}

// One stage of the schedule: layout B, entered from layout BPREV, followed by layout BN
// (BN = 0 ends the schedule).  `k` is the next output level.  Once the walk has started
// every later stage runs, so the previous layout is always the previous schedule entry.
template <int BPREV, int B, int BN, bool AMER>
__device__ __forceinline__ void stage(double (&v)[32], const UTab& U, const Lattice& L,
                                      int N, int& k, bool& started, double* tile, int lane) {
    if (!started) {
        // terminal payoff, binomial.py:20-25
        const double sbN = lane_base(L, lane, B, N);
#pragma unroll
        for (int i = 0; i < B; ++i) v[i] = pick_max(0.0, fma(sbN * U.hi[i >> 3], U.lo[i & 7], L.sK));
        started = true;
    } else {
        repack<BPREV, B>(v, tile, lane);
    }
    const int k_lo = (BN > 0) ? (32 * BN - 1) : 0;
    const double sb = AMER ? lane_base(L, lane, B, k + 1) : 0.0;
    run_levels<B, AMER>(v, U, sb, L, k, k_lo);
    k = k_lo - 1;
}

template <bool AMER>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 3)
crr_warp_kernel(CrrArgs a) {
    __shared__ double tiles[kWarpsPerCta][kTileDoubles];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    double* tile = tiles[wib];
    const unsigned long long warp0 = (unsigned long long)blockIdx.x * kWarpsPerCta + wib;
    const unsigned long long nwarps = (unsigned long long)gridDim.x * kWarpsPerCta;
    if (a.todo_only && a.decl_count && a.decl_count[a.todo_only - 2] == 0) return;  // nothing was left for this pass
    for (unsigned long long o = warp0; o < a.n; o += nwarps) {
        if (a.todo_only && a.status[o] != a.todo_only) continue;
        const Lattice L = make_lattice(a, o);
        if (!L.ok) {  // binomial.py:16-17 -> per-contract status instead of raising
            if (lane == 0) {
                a.price[o] = nan("");
                a.status[o] = 1;
            }
            continue;
        }
        double v[32];
        UTab U;
        {
            const double u2 = L.u * L.u;
            U.lo[0] = 1.0;
#pragma unroll
            for (int i = 1; i < 8; ++i) U.lo[i] = U.lo[i - 1] * u2;
            const double u16 = U.lo[7] * u2;
            U.hi[0] = 1.0;
#pragma unroll
            for (int a = 1; a < 4; ++a) U.hi[a] = U.hi[a - 1] * u16;
        }
        int k = a.N - 1;
        bool started = false;
        const int N = a.N;
        // pick the first layout that holds the N+1 terminal nodes, then walk the schedule
#define OPTP_STAGE(BP, B, BN)                                                         \
    if (started || N + 1 > 32 * (BN))                                                 \
        stage<BP, B, BN, AMER>(v, U, L, N, k, started, tile, lane);
        OPTP_STAGE(32, 32, 28)
        OPTP_STAGE(32, 28, 24)
        OPTP_STAGE(28, 24, 20)
        OPTP_STAGE(24, 20, 16)
        OPTP_STAGE(20, 16, 12)
        OPTP_STAGE(16, 12, 8)
        OPTP_STAGE(12, 8, 6)
        OPTP_STAGE(8, 6, 4)
        OPTP_STAGE(6, 4, 2)
        OPTP_STAGE(4, 2, 1)
        OPTP_STAGE(2, 1, 0)
#undef OPTP_STAGE
        if (lane == 0) {
            a.price[o] = v[0];
            a.status[o] = 0;
        }
    }
}


// ---------------------------------------------------------------- banded American put kernel
// The N = 1000 lattice spans +-31 standard deviations of ln S_T; almost all of it is either
// exactly worthless (every terminal descendant out of the money: V = 0 bit-for-bit), worthless
// to 2^-80 K, or deep inside the early-exercise region where V = K - S is known in closed form.
// This kernel walks only the band between the exercise boundary and the (numerically) zero
// region, one option per warp, band in registers, in segments of <= 32 levels:
//   * within a segment the window [w0, w0 + 32B) is fixed in index space; node j needs nodes j
//     and j+1 of the next level only, so nothing BELOW the window is ever read, and the node
//     just ABOVE the window top is 0 (lane 31's neighbour);
//   * at a segment boundary the band [jlo, jhi) is measured from the registers (jlo = first
//     node that did not choose exercise at the last level, jhi = 1 + last node worth more than
//     2^-80 K), the next window [jlo - nlev - 2, min(jhi, k+1)) is laid out with the smallest
//     B that holds it, and the level is re-packed through a shared-memory tile; nodes entering
//     from below get the exercise value, nodes entering from above get 0.
// Why the cut at the bottom is exact: if both children of a node are exercise nodes then
// cont = disc*K - S*e^{-q dt} < K - S whenever K(1-disc) > S(1-e^{-q dt}), which r > q >= 0
// guarantees for every S < K; so the exercise boundary index drops by at most one per level
// and a margin of nlev + 2 nodes keeps the window bottom inside the exercise region for the
// whole segment.  The cut at the top discards values below 2^-80 K (they can only shrink at a
// fixed index going back in time); the induced error is < 1e-21 K.  Contracts that do not
// qualify (calls, European, r <= q, band wider than 512 nodes) are left at status 2 and priced
// by the full-width kernel in a second launch.
constexpr int kBandWarps = 4;
constexpr int kBandMaxB = 16;
constexpr int kBandTile = 32 * kBandMaxB + 8;
#ifndef OPTP_BAND_SEG
#define OPTP_BAND_SEG 64
#endif
constexpr int kBandSeg = OPTP_BAND_SEG;  // levels per segment

// one level for B nodes per lane; `nb` is the next lane's first node (0 above the window).
// FLAGS: also return a bit mask of the nodes that chose exercise.
template <int B, bool FLAGS>
__device__ __forceinline__ unsigned int band_level(double (&v)[kBandMaxB], const UTab& U, double rho,
                                                   double rho_top, double sbg, double sKg, double nb,
                                                   bool& ex0) {
    double sa[2];
    sa[0] = sbg;
#pragma unroll
    for (int a = 1; a < 2; ++a)
        if (8 * a < B) sa[a] = sbg * U.hi[a];
    unsigned int mask = 0;
    constexpr int G = (B >= 8) ? 8 : B;
#pragma unroll
    for (int g0 = 0; g0 < B; g0 += G) {
        double c[G], e[G];
#pragma unroll
        for (int t = 0; t < G; ++t) {
            const int i = g0 + t;
            if (i < B) {
                // the lane's last node takes the next lane's first node; in lane 31 rho_top is 0, so
                // the node above the window counts as 0 without a per-level select
                c[t] = (i + 1 < B) ? fma(rho, v[i + 1], v[i]) : fma(rho_top, nb, v[i]);
            }
        }
#pragma unroll
        for (int t = 0; t < G; ++t)
            if (g0 + t < B) e[t] = fma(sa[(g0 + t) >> 3], U.lo[(g0 + t) & 7], sKg);
#pragma unroll
        for (int t = 0; t < G; ++t) {
            if (g0 + t < B) {
                const bool ex = e[t] > c[t];
                if (g0 + t == 0) ex0 = ex;
                if (FLAGS) mask |= ex ? (1u << (g0 + t)) : 0u;
                v[g0 + t] = ex ? e[t] : c[t];
            }
        }
    }
    return mask;
}

// nlev levels (the first one is level k_first) on window w0 in layout B; returns the exercise
// mask of the last level.  nlev <= L.R, so the scaled recursion cannot overflow.
template <int B>
__device__ __forceinline__ unsigned int band_segment(double (&v)[kBandMaxB], const UTab& U,
                                                     const Lattice& L, int lane, double sb, int nlev,
                                                     bool& bottom_always_exercise) {
    // sb = sgn*S0*u^(2*(w0 + lane*B) - k): this lane's first node one level above the segment
    double sbg = sb;
    double sKg = L.sK;
    const double rho_top = (lane == 31) ? 0.0 : L.rho;
    double nb = __shfl_down_sync(0xffffffffu, v[0], 1);
    unsigned int mask = 0;
    bool ex0 = true;
    // the last level is peeled: it is the only one that records the full exercise mask
#pragma unroll 1
    for (int t = 0; t < nlev - 1; ++t) {
        sbg *= L.u_pd;
        sKg *= L.inv_pd;
        bool e0;
        band_level<B, false>(v, U, L.rho, rho_top, sbg, sKg, nb, e0);
        ex0 = ex0 && e0;
        nb = __shfl_down_sync(0xffffffffu, v[0], 1);
    }
    {
        sbg *= L.u_pd;
        sKg *= L.inv_pd;
        bool e0;
        mask = band_level<B, true>(v, U, L.rho, rho_top, sbg, sKg, nb, e0);
        ex0 = ex0 && e0;
    }
    // undo the pd^-nlev scaling: pd^nlev by square-and-multiply (nlev <= 64)
    double back = 1.0, sq = L.pd;
    for (int e = nlev; e > 0; e >>= 1) {
        if (e & 1) back *= sq;
        sq *= sq;
    }
    bottom_always_exercise = ex0;  // meaningful in lane 0: the window's bottom node, every level
#pragma unroll
    for (int i = 0; i < B; ++i) v[i] *= back;
    return mask;
}

template <int B>
__device__ __forceinline__ void band_store(const double (&v)[kBandMaxB], double* tile, int lane) {
#pragma unroll
    for (int i = 0; i < B; ++i) tile[i * 32 + lane] = v[i];  // conflict-free, slot-major
}

// new layout (w0n, BN) from the tile holding the old layout (w0, width_old = 32*B_old, slot-major):
// nodes below the old window enter with their exercise value, nodes above it with 0
template <int BN>
__device__ __forceinline__ void band_load(double (&v)[kBandMaxB], const double* tile, const UTab& U,
                                          double sb, double sK, int j_first, int w0_old, int B_old,
                                          int magic) {
    const int width_old = 32 * B_old;
#pragma unroll
    for (int i = 0; i < BN; ++i) {
        const int jr = j_first + i - w0_old;
        double val;
        if (jr < 0) {
            val = fma(sb * U.hi[i >> 3], U.lo[i & 7], sK);
        } else if (jr >= width_old) {
            val = 0.0;
        } else {
            const int lo_ = (jr * magic) >> 16;   // jr / B_old, exact for jr < 2048
            val = tile[(jr - lo_ * B_old) * 32 + lo_];
        }
        v[i] = val;
    }
}

// lane-local scan of the registers at level k: first non-exercise slot, last slot above thr
template <int B>
__device__ __forceinline__ void band_scan(const double (&v)[kBandMaxB], unsigned int exmask, int j0,
                                          int k, double thr, int& first_cont, int& last_live) {
    const unsigned int full = (B >= 32) ? 0xffffffffu : ((1u << B) - 1u);
    const unsigned int cont = ~exmask & full;
    first_cont = cont ? (__ffs(cont) - 1) : B;
    last_live = -1;
#pragma unroll
    for (int i = 0; i < B; ++i)
        if (v[i] > thr && j0 + i <= k) last_live = i;
}

#define OPTP_BAND_SWITCH(Bv, CALL)        \
    switch (Bv) {                         \
        case 1: { constexpr int BB = 1; CALL; } break;   \
        case 2: { constexpr int BB = 2; CALL; } break;   \
        case 3: { constexpr int BB = 3; CALL; } break;   \
        case 4: { constexpr int BB = 4; CALL; } break;   \
        case 6: { constexpr int BB = 6; CALL; } break;   \
        case 8: { constexpr int BB = 8; CALL; } break;   \
        case 12: { constexpr int BB = 12; CALL; } break; \
        default: { constexpr int BB = 16; CALL; } break; \
    }

__device__ __forceinline__ int band_pick_B(int width) {
    const int need = (width + 31) >> 5;
    if (need <= 4) return need < 1 ? 1 : need;
    if (need <= 6) return 6;
    if (need <= 8) return 8;
    if (need <= 12) return 12;
    return 16;
}

__global__ void __launch_bounds__(kBandWarps * 32, 5) crr_band_kernel(CrrArgs a) {
    __shared__ double tiles[kBandWarps][kBandTile];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    double* tile = tiles[wib];
    const int N = a.N;
    const unsigned long long warp0 = (unsigned long long)blockIdx.x * kBandWarps + wib;
    const unsigned long long nwarps = (unsigned long long)gridDim.x * kBandWarps;
    for (unsigned long long o = warp0; o < a.n; o += nwarps) {
        const Lattice L = make_lattice(a, o);
        if (!L.ok) {
            if (lane == 0) {
                a.price[o] = nan("");
                a.status[o] = 1;
            }
            continue;
        }
        const double r = a.r[o], q = a.q[o], K = a.K[o];
        const double dt = a.T[o] / (double)N;
        const double one_m_disc = -expm1(-r * dt), one_m_dq = -expm1(-q * dt);
        const bool eligible = (L.sgn < 0.0) && q >= 0.0 && one_m_disc > 1e-9 &&
                              one_m_disc > one_m_dq * (1.0 + 1e-6);
        if (!eligible) {
            if (lane == 0) a.status[o] = 2;
            continue;
        }
        UTab U;
        {
            const double u2 = L.u * L.u;
            U.lo[0] = 1.0;
#pragma unroll
            for (int i = 1; i < 8; ++i) U.lo[i] = U.lo[i - 1] * u2;
            const double u16 = U.lo[7] * u2;
            U.hi[0] = 1.0;
#pragma unroll
            for (int i = 1; i < 4; ++i) U.hi[i] = U.hi[i - 1] * u16;
        }
        const double thr = K * 8.2718061255302767e-25;  // 2^-80 K
        const double ln_u = log(L.u);
        double v[kBandMaxB];
        // ---- level N: every node is closed form, max(K - S, 0) (binomial.py:20-25) ----------
        int B = 2, k = N;
        // index of the first out-of-the-money terminal node, S0*u^(2j-N) >= K
        const double jz = 0.5 * ((double)N + log(K / L.S0) / ln_u);
        int w0 = (int)fmin(fmax(jz - 31.0, 0.0), (double)max(0, N + 1 - 64));
        unsigned int exmask = 0;
        {
            const double sb = L.sgn * L.S0 * exp(ln_u * (double)(2 * (w0 + lane * 2) - N));
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const double e = fma(sb, U.lo[i], L.sK);
                v[i] = e > 0.0 ? e : 0.0;
                exmask |= (e > 0.0) ? (1u << i) : 0u;
            }
        }
        bool fail = false;
        double price = 0.0;
        int prev_drop = 0, prev_nlev = 0, jlo_seg_start = -1;
        while (true) {
            // ---- measure the band [jlo, jhi) of the level held in registers ------------------
            int first_cont = 0, last_live = -1;
            OPTP_BAND_SWITCH(B, band_scan<BB>(v, exmask, w0 + lane * BB, k, thr, first_cont, last_live));
            const unsigned int has_cont = __ballot_sync(0xffffffffu, first_cont < B);
            const unsigned int has_live = __ballot_sync(0xffffffffu, last_live >= 0);
            int jlo, jhi;
            if (has_cont) {
                const int f = __ffs(has_cont) - 1;
                jlo = w0 + f * B + __shfl_sync(0xffffffffu, first_cont, f);
            } else {
                jlo = w0 + 32 * B;
            }
            if (has_live) {
                const int f = 31 - __clz(has_live);
                jhi = w0 + f * B + __shfl_sync(0xffffffffu, last_live, f) + 1;
            } else {
                jhi = w0;
            }
            if (jhi < jlo) jhi = jlo;
            if (jlo_seg_start >= 0) prev_drop = max(0, jlo_seg_start - jlo);
            // the window bottom must still be inside the exercise region (or be node 0)
            if (w0 > 0 && jlo <= w0) fail = true;
            if (k == 0 || fail) break;
            // ---- next segment: levels k-1 ... k-nlev ------------------------------------------
            int nlev = min(kBandSeg, min(L.R, k));
            int top = min(jhi, k + 1);           // exclusive; node `top` of level k is 0 (or dead)
            // margin below the boundary: rigorous (one node per level) for the first segment; later
            // the boundary falls no faster than it did over the previous segment (it decelerates
            // like sqrt(tau)), and the per-level bottom check catches any violation
            int margin = nlev + 2;
            if (prev_nlev > 0) margin = min(margin, (prev_drop * nlev + prev_nlev - 1) / prev_nlev + 6);
            jlo_seg_start = jlo;
            int bot = max(0, jlo - margin);
            if (top < bot + 1) top = bot + 1;
            const int Bn = band_pick_B(top - bot);
            if (top - bot > 32 * kBandMaxB) {
                fail = true;
                break;
            }
            int w0n = max(0, top - 32 * Bn);
            // ---- re-pack level k from (w0, B) to (w0n, Bn) -------------------------------------
            OPTP_BAND_SWITCH(B, band_store<BB>(v, tile, lane));
            __syncwarp();
            // this lane's first node of the new layout at level k: sgn*S0*u^(2j - k), via exp()
            const double sb = L.sgn * L.S0 * exp(ln_u * (double)(2 * (w0n + lane * Bn) - k));
            {
                const int magic = (65536 + B - 1) / B;   // exact j / B for j < 2048
                const int B_old = B, w0_old = w0;
                OPTP_BAND_SWITCH(Bn, band_load<BB>(v, tile, U, sb, L.sK, w0n + lane * BB, w0_old, B_old, magic));
            }
            __syncwarp();
            w0 = w0n;
            B = Bn;
            bool bot_ok = true;
            OPTP_BAND_SWITCH(B, exmask = band_segment<BB>(v, U, L, lane, sb, nlev, bot_ok));
            // everything below the window is an exercise node only while the window's bottom node
            // chose exercise at EVERY level (then the node below it has two exercise children)
            if (w0 > 0 && !__shfl_sync(0xffffffffu, (int)bot_ok, 0)) fail = true;
            k -= nlev;
            prev_nlev = nlev;
        }
        if (fail) {
            if (lane == 0) a.status[o] = 2;
            continue;
        }
        // level 0: node 0 is in the window (w0 == 0) or below it (then it is an exercise node)
        price = (w0 == 0) ? __shfl_sync(0xffffffffu, v[0], 0) : (K - L.S0);
        if (lane == 0) {
            a.price[o] = price;
            a.status[o] = 0;
        }
    }
}
#undef OPTP_BAND_SWITCH

// ---------------------------------------------------------------- two-zone banded American put kernel
// Refinement of the banded kernel above, on three counts.
//  (1) Inside the band almost every node is a CONTINUATION node: only the few nodes next to the
//      early-exercise boundary can choose exercise.  The window is split in two zones:
//        zone Z = [w0, w0 + ZW)              full update  max(cont, K - S)   (3 FP64 ops + 2 selects per node)
//        zone C = [w0 + ZW, w0 + ZW + 32*M)  continuation only, A[j] = rho*A[j+1] + A[j]   (ONE DFMA per node)
//      Z holds the margin below the boundary and at least two nodes above it: ZW = 16 in the default build, whose
//      window drifts with the boundary (OPTP_ZONE_DRIFT below), ZW = 32 with a fixed origin (-DOPTP_ZONE_DRIFT=0:
//      the boundary index falls by at most one per level, see above).
//      Why C needs no exercise test, checked at run time: on every level V + S is non-decreasing in
//      the node index when q >= 0 (induction: pu*S(j+1) + pd*S(j) = S*e^{-q dt}), so the exercise
//      nodes of a level are contiguous from the bottom; as long as the TOP node of Z chooses
//      continuation on every level of the segment, every node of C is a continuation node.  The
//      kernel counts that test over the segment (last lane of the option) exactly like the
//      bottom-node test (first lane); the count stops at the level where Z's top node leaves the
//      tree (index > level: C has no real node from there on).  A contract that fails either test
//      keeps status 2 and is priced by the full-width kernel.
//  (2) The per-level cost of a warp is dominated by instructions that do not depend on how many
//      lanes do useful work (neighbour shuffles, scaling multiplies, loop control), so G = 4 (or 2)
//      options share one warp, 32/G lanes each, every lane owning ZW*G/32 nodes of Z and G*M nodes of C
//      of ITS option.  All options of a warp walk the same levels with the same layout (M = the
//      widest C zone among them); everything else (lattice constants, window origin, band limits,
//      failure flags) is per option.
//  (3) The schedule is FIXED -- kZoneSeg0 levels from level N, then segments of kZoneSeg levels --
//      so the arithmetic applied to a contract does not depend on its warp mates or on G: a
//      contract's price is bit-identical whatever batch it is priced in.
#ifndef OPTP_ZONE_MINB4
#define OPTP_ZONE_MINB4 4
#endif
#ifndef OPTP_ZONE_WARPS
#define OPTP_ZONE_WARPS 4
#endif
constexpr int kZoneWarps = OPTP_ZONE_WARPS;  // warps per CTA of the two-zone kernel
// Fixed origin (-DOPTP_ZONE_DRIFT=0, the kernel of the first half of round 2; the default is the drifting walk below).
// Z = [jlo - 30, jlo + 2): 30 nodes below the first continuation node jlo of the segment's first level.
// RIGOROUSLY the boundary index falls by at most one per level (see above), which would allow segments of 28
// levels.  Measured (scripts/crr_boundary_drop.py: the exact lattice boundary of 250 contracts over sigma in
// [0.05, 1], T in [0.05, 5], r up to 15 %, N = 200 / 1000 / 3000), it falls by 1/2 + O(1/sqrt(n)) per level n
// levels from expiry: at most 27 nodes over the FIRST 34 levels, and at most 27 over any stretch of 44 levels
// that starts 32 or more levels from expiry (Z tolerates 29).  The schedule therefore walks kZoneSeg0 levels from level N and
// kZoneSeg levels per segment after that -- 23 segments per N = 1000 option instead of 36, and the work between
// segments (tile round trip, band measurement, plan) is a third of the walk.  Exactness does not rest on the
// measurement: the kernel checks at EVERY level that Z's bottom node chose exercise, and a contract that fails
// keeps status 2 and is priced by the full-width kernel.
// OPTP_ZONE_DRIFT = 1: the window FOLLOWS the boundary inside a segment.  A lattice line of constant S falls by half a
// node per level, and so does the exercise boundary up to O(1/sqrt(n)); a window with a fixed origin therefore needs
// nlev/2 nodes of Z just to absorb that trend.  The drifting walk alternates two kinds of level:
//     A (origin stays):     new[m] = pu*old[m+1] + pd*old[m]       (the classic in-place update, ascending)
//     B (origin one lower): new[m] = pu*old[m]   + pd*old[m-1]     (position m now holds node j-1; descending)
// so position m keeps the same S up to one factor u, the boundary stays where it was in the window (the emulation
// scripts/crr_drift_emu.py sees it move by at most 5 positions over 64 levels, 10 over the first 32), and Z shrinks
// from 32 to 16 nodes: 13 below the first continuation node, 2 above.  The node that enters at the bottom in a B
// level takes the exercise value of the node below the window (justified by the same bottom test as before); the
// node that leaves at the top must be dead, so the window is laid out nlev/2 nodes above the live top of the
// segment's first level (values at a fixed index only shrink going back).  The window origin may go NEGATIVE: the
// walk then runs on the tree extended downwards (nodes j < 0 at S0*u^(2j-k), a consistent lattice on which every
// argument above holds and on which nothing below a node ever enters its value), so there is no second code path
// for windows pinned at node 0.
#ifndef OPTP_ZONE_DRIFT
#define OPTP_ZONE_DRIFT 1
#endif
constexpr bool kZoneDrift = OPTP_ZONE_DRIFT != 0;
#ifndef OPTP_ZONE_SEG
#define OPTP_ZONE_SEG (OPTP_ZONE_DRIFT ? 64 : 44)
#endif
#ifndef OPTP_ZONE_SEG0
#define OPTP_ZONE_SEG0 32
#endif
constexpr int kZoneBelow = kZoneDrift ? 13 : 30;  // Z nodes below the first continuation node
constexpr int kZoneSeg = OPTP_ZONE_SEG;       // levels per segment
constexpr int kZoneSeg0 = OPTP_ZONE_SEG0;     // levels of the first segment (from level N)
// rigorous schedule of the retry pass: the boundary falls by at most one node per level -- relative to a drifting
// window by at most half a node per level
constexpr int kZoneSegRig = kZoneDrift ? 2 * (kZoneBelow - 1) : kZoneBelow - 2;
static_assert(kZoneSeg0 <= kZoneSeg, "the renormalisation bound L.R >= kZoneSeg covers the first segment too");
static_assert(!kZoneDrift || (kZoneSeg % 2 == 0 && kZoneSeg0 % 2 == 0), "the drifting walk steps in pairs of levels");
// Z nodes per lane and width of Z for G options per warp
__host__ __device__ constexpr int zone_bz(int G) { return kZoneDrift ? (G >= 2 ? G / 2 : 1) : G; }
__host__ __device__ constexpr int zone_zw(int G) { return zone_bz(G) * (32 / G); }
// A lane holds at most 24 C nodes: 6 chunks of 32 nodes per option at four options per warp (band up to
// 224 nodes), 12 at two (416), 24 at one (800).  A contract whose band is too wide for the layout it was
// tried in keeps status 4 and is walked again by the next narrower-group pass (launch_crr); the
// arithmetic applied to a node does not depend on the layout, so neither does the price.
#ifndef OPTP_ZONE_LANEC
#define OPTP_ZONE_LANEC (OPTP_ZONE_DRIFT ? 28 : 24)
#endif
constexpr int kZoneLaneC = OPTP_ZONE_LANEC;
static_assert(kZoneLaneC <= 28 && kZoneLaneC % 4 == 0, "the segment dispatch covers M <= 28");
// C grows in chunks of zone_ch(G) nodes per option, zone_bcu(G) nodes per lane; M counts chunks.  Chunks of 32 nodes.
// (-DOPTP_ZONE_CH16=1, drifting walk with two or four options per warp: chunks of 16.  It trims ~8 of ~140 C nodes per
// level and was measured SLOWER, 42.9 M against 49.4 M options/s on config 2: twice as many segment bodies for the
// instruction caches, and lane chunks of 2, 6, 10 ... doubles that the tile padding was not laid out for.)
#ifndef OPTP_ZONE_CH16
#define OPTP_ZONE_CH16 0
#endif
__host__ __device__ constexpr int zone_ch(int G) { return (kZoneDrift && OPTP_ZONE_CH16 && G >= 2) ? 16 : 32; }
__host__ __device__ constexpr int zone_bcu(int G) { return zone_ch(G) * G / 32; }
__host__ __device__ constexpr int zone_max_m(int G) { return kZoneLaneC / zone_bcu(G); }
__host__ __device__ constexpr int zone_cap(int G) { return 32 + zone_ch(G) * zone_max_m(G); }  // nodes of a window
// nodes of C that are not band: the drifting walk starts C 3 nodes above the boundary and keeps nlev/2 nodes above the
// live top; with a fixed origin C starts 2 nodes above the boundary
__host__ __device__ constexpr int zone_overhead() { return kZoneDrift ? (kZoneSeg / 2 - 3) : -2; }
// re-pack tile, node-ordered per option: node e lives at e + (e >> OPTP_ZONE_PAD_SHIFT), one pad double per
// 2^shift nodes.  A lane's chunk of 4M consecutive nodes then starts 5M doubles after its neighbour's (shift 2):
// 5M mod 16 runs through distinct bank pairs for the lanes of an option (M != 4), and the tile of the next option
// starts 8 (mod 16) doubles further, so the two options of a half-warp interleave.
#ifndef OPTP_ZONE_PAD_SHIFT
#define OPTP_ZONE_PAD_SHIFT 2
#endif
__host__ __device__ constexpr int zone_pad_c(int e) { return e + (e >> OPTP_ZONE_PAD_SHIFT); }
__host__ __device__ constexpr int zone_group_tile(int G) {
    // zone_cap(G) nodes padded, rounded up to 8 (mod 16)
    return ((zone_pad_c(zone_cap(G)) + 7) & ~15) + 8;
}
__device__ __forceinline__ int zone_pad(int e) { return e + (e >> OPTP_ZONE_PAD_SHIFT); }
// -DOPTP_ZONE_CHECK (OPTP_NVCC_FLAGS, build_native.py --force): every tile access of the two-zone kernel is
// bounds-checked and traps -- the in-house stand-in for compute-sanitizer, run over the stress tests
#ifdef OPTP_ZONE_CHECK
__device__ __forceinline__ int zone_pad_checked(int e, int limit) {
    if ((unsigned int)e >= (unsigned int)limit) __trap();
    return zone_pad(e);
}
#define OPTP_ZP(e) zone_pad_checked((e), zone_cap(G))
#else
#define OPTP_ZP(e) zone_pad(e)
#endif
static_assert(zone_pad_c(zone_cap(4)) <= zone_group_tile(4) && zone_group_tile(4) % 16 == 8, "tile too small");
static_assert(zone_pad_c(zone_cap(2)) <= zone_group_tile(2) && zone_group_tile(2) % 16 == 8, "tile too small");
static_assert(zone_pad_c(zone_cap(1)) <= zone_group_tile(1) && zone_group_tile(1) % 16 == 8, "tile too small");

// One segment in layout (BZ = G, BC = G*M), M a compile-time constant:
//   * the lane's C nodes are loaded from the node-ordered tile, which holds level k in the PREVIOUS layout
//     (tile index jr0 + i; nodes outside [0, lim) are above the band and count as 0);
//   * nlev levels are walked;
//   * the pd^-nlev scaling is undone (back = pd^nlev) and the lane's C nodes are written back to the tile in THIS
//     layout.  (The caller loads and stores the Z nodes: with the Z stores in here ptxas no longer proves the warp
//     converged and guards every shuffle of the walk with a BRA.DIV.)
// sb = sgn*S0*u^(2*(w0 + quad) - k): the first node of this lane's quad of Z nodes one level above the segment
// (ulo holds the lane's factors within the quad).  Returns the exercise mask of this lane's Z nodes at the last
// level; n_ex0 counts the levels at which the lane's first Z node chose exercise, t_exT is the first level at
// which its last Z node did (INT_MAX: never).
template <int G, int M>
__device__ __forceinline__ unsigned int zone_segment(double (&z)[G], double (&c)[kZoneLaneC],
                                                     const double (&ulo)[G], const Lattice& L, int gl,
                                                     int src, double sb, int nlev, double back, double* tile,
                                                     int jr0, int lim, int& n_ex0_out, int& t_exT_out) {
    constexpr int BZ = G, BC = zone_bcu(G) * M, LG = 32 / G;
    constexpr int kMaxM = zone_max_m(G);  // (OPTP_ZP's bound in -DOPTP_ZONE_CHECK builds)
    (void)kMaxM;
#pragma unroll
    for (int i = 0; i < BC; ++i) {
        const int jr = jr0 + i;
        c[i] = ((unsigned int)jr < (unsigned int)lim) ? tile[OPTP_ZP(jr)] : 0.0;
    }
    __syncwarp();  // every lane has read level k before this segment's stores overwrite the tile
    double sbg = sb, sKg = L.sK;
    const double rho = L.rho;
    const bool last = (gl == LG - 1);
    const double rho_top = last ? 0.0 : rho;  // the node above the window counts as 0
    // The top Z node of a lane takes the next lane's first Z node; in the option's last lane it takes the first
    // C node of the option's first lane, which the ring shuffle of c[0] delivers there as nbc.  Both products
    // are formed and one multiplier is zero (x + 0*y = x exactly for finite y: values are bounded by 2^200 K),
    // so no 64-bit select sits between the shuffles and the arithmetic: one DFMA instead of two FSELs.
    const double rz_z = (BC > 0) ? (last ? 0.0 : rho) : rho_top;
    const double rz_c = (BC > 0 && last) ? rho : 0.0;
    // a counter and a running minimum instead of predicate chains: ptxas folds `all = all && (e > c)` into a
    // second and third DSETP per level, and the FP64 pipe is the one this loop loads most
    int n_ex0 = 0, t_exT = 0x7fffffff;
    double nbz = __shfl_sync(0xffffffffu, z[0], src);
    double nbc = 0.0;
    if (BC > 0) nbc = __shfl_sync(0xffffffffu, c[0], src);
#pragma unroll 2
    for (int t = 0; t < nlev; ++t) {
        sbg *= L.u_pd;
        sKg *= L.inv_pd;
        double cz[BZ], ez[BZ];
#pragma unroll
        for (int i = 0; i < BZ; ++i) {
            if (i + 1 < BZ) cz[i] = fma(rho, z[i + 1], z[i]);
            else if (BC > 0) cz[i] = fma(rz_z, nbz, fma(rz_c, nbc, z[i]));
            else cz[i] = fma(rz_z, nbz, z[i]);
        }
#pragma unroll
        for (int i = 0; i < BZ; ++i) ez[i] = fma(sbg, ulo[i], sKg);
        // everything is updated in place, in ascending node order, so no value has to be carried
        // in a second register across the loop edge
#pragma unroll
        for (int i = 0; i < BC; ++i)
            c[i] = (i + 1 < BC) ? fma(rho, c[i + 1], c[i]) : fma(rho_top, nbc, c[i]);
#pragma unroll
        for (int i = 0; i < BZ; ++i) {
            const bool ex = ez[i] > cz[i];
            if (i == 0) n_ex0 += ex ? 1 : 0;
            if (i == BZ - 1) t_exT = ex ? min(t_exT, t) : t_exT;
            z[i] = ex ? ez[i] : cz[i];
        }
        if (BC > 0) nbc = __shfl_sync(0xffffffffu, c[0], src);
        nbz = __shfl_sync(0xffffffffu, z[0], src);
    }
    // exercise mask of the last level: a node chose exercise iff it holds the exercise value
    unsigned int mask = 0;
#pragma unroll
    for (int i = 0; i < BZ; ++i) mask |= (z[i] == fma(sbg, ulo[i], sKg)) ? (1u << i) : 0u;
    // undo the pd^-nlev scaling (back = pd^nlev)
#pragma unroll
    for (int i = 0; i < BZ; ++i) z[i] *= back;
#pragma unroll
    for (int i = 0; i < BC; ++i) {
        c[i] *= back;
        tile[OPTP_ZP(32 + gl * BC + i)] = c[i];
    }
    n_ex0_out = n_ex0;
    t_exT_out = t_exT;
    return mask;
}

// The drifting walk (OPTP_ZONE_DRIFT): `npairs` pairs of levels (A then B, see above) in layout (BZ, BC = G*M).
// Scaled recursion: an A level divides by pd (A[m] = (pu/pd)*A[m+1] + A[m]), a B level by pu
// (A[m] = A[m] + (pd/pu)*A[m-1]), so both kinds spend ONE DFMA per node; back = (pd*pu)^npairs undoes it.
// kB.rho2 = pd/pu, kB.inv_pu = 1/pu, kB.d_pu = 1/(u*pu) (position m moves to a node whose S is one factor u lower),
// ulo_below = u^-2 (the node below the window, relative to the first lane's base).  Tile I/O as in zone_segment.
struct DriftConsts {
    double rho2, inv_pu, d_pu, ulo_below;
};
template <int G, int M>
__device__ __forceinline__ unsigned int zone_segment_drift(double (&z)[zone_bz(G)], double (&c)[kZoneLaneC],
                                                           const double (&ulo)[zone_bz(G)], const Lattice& L,
                                                           const DriftConsts& kB, double ub, int gl, int src_up,
                                                           int src_dn, double sb, int npairs, double back, double* tile, int jr0,
                                                           int lim, int& n_ex0_out, int& t_exT_out) {
    constexpr int BZ = zone_bz(G), BC = zone_bcu(G) * M, LG = 32 / G, ZW = BZ * LG;
    constexpr int kMaxM = zone_max_m(G);  // (OPTP_ZP's bound in -DOPTP_ZONE_CHECK builds)
    (void)kMaxM;
#pragma unroll
#if defined(OPTP_ZONE_CHECK) || OPTP_ZONE_PAD_SHIFT != 2 || defined(OPTP_ZONE_PLAIN_IO)
    for (int i = 0; i < BC; ++i) {
        const int jr = jr0 + i;
        c[i] = ((unsigned int)jr < (unsigned int)lim) ? tile[OPTP_ZP(jr)] : 0.0;
    }
#else
    {
        // zone_pad(jr0 + i) = e + (e >> 2) without per-element address arithmetic: with b = jr0 & ~3, r = jr0 & 3 it
        // is [b + (b >> 2) + r] + [i + (i >> 2)] + (((i & 3) + r) >> 2) -- four base pointers (one per i & 3, the
        // carry of the low bits) and a compile-time offset per element
        const int r = jr0 & 3, b = jr0 - r;
        const double* const t0 = tile + (b + (b >> 2) + r);
        const double* const tb[4] = {t0, t0 + (r == 3 ? 1 : 0), t0 + (r >= 2 ? 1 : 0), t0 + (r >= 1 ? 1 : 0)};
#pragma unroll
        for (int i = 0; i < BC; ++i) {
            const int jr = jr0 + i;
            c[i] = ((unsigned int)jr < (unsigned int)lim) ? tb[i & 3][i + (i >> 2)] : 0.0;
        }
    }
#endif
    __syncwarp();  // every lane has read level k before this segment's stores overwrite the tile
    double sbg = sb, sKg = L.sK;
    const double rho = L.rho, rho2 = kB.rho2;
    const bool last = (gl == LG - 1), first = (gl == 0);
    // the exercise test covers the first 16 nodes whatever G is (at G = 1 Z holds 32): which nodes may choose
    // exercise must not depend on the layout, or a tie in the last bit could be resolved differently
    const bool ztest = (ZW <= 16) || (gl * BZ < 16);
    // Cross links of the two rings (Z and C) as zero multipliers (x + 0*y = x exactly, every y is finite):
    // A level: a lane's top node takes the next lane's first node (ra_nb; nothing above the window), and Z's top
    //          node in the option's last lane takes C's first node, which the ring shuffle of c[0] delivers (ra_x);
    // B level: a lane's first node takes the previous lane's last node (rb_nb); in the option's first lane Z's
    //          first node takes the node below the window and C's first node takes Z's top node, which the ring
    //          shuffle of z[BZ-1] delivers (rb_x).
    const double ra_nb = last ? 0.0 : rho;
    const double ra_x = (BC > 0 && last) ? rho : 0.0;
    const double rb_nb = first ? 0.0 : rho2;
    const double rb_x = first ? rho2 : 0.0;
    int n_ex0 = 0, t_exT = 0x7fffffff;
    double nbz = __shfl_sync(0xffffffffu, z[0], src_up);
    double nbc = 0.0;
    if (BC > 0) nbc = __shfl_sync(0xffffffffu, c[0], src_up);
    // (-DOPTP_ZONE_PAIR_UNROLL: two pairs per trip.  Measured 43.0 M against 49.7 M options/s on config 2 -- like the
    // 16-node chunks, twice the loop bodies for the instruction caches cost far more than the loop control saves.)
#ifdef OPTP_ZONE_PAIR_UNROLL
#pragma unroll 2
#endif
    for (int p = 0; p < npairs; ++p) {
        double cz[BZ], ez[BZ];
        // ---- A level
        sbg *= L.u_pd;
        sKg *= L.inv_pd;
#pragma unroll
        for (int i = 0; i < BZ; ++i) {
            if (i + 1 < BZ) cz[i] = fma(rho, z[i + 1], z[i]);
            else if (BC > 0) cz[i] = fma(ra_nb, nbz, fma(ra_x, nbc, z[i]));
            else cz[i] = fma(ra_nb, nbz, z[i]);
        }
#pragma unroll
        for (int i = 0; i < BZ; ++i) ez[i] = fma(sbg, ulo[i], sKg);
#pragma unroll
        for (int i = 0; i < BC; ++i) c[i] = (i + 1 < BC) ? fma(rho, c[i + 1], c[i]) : fma(ra_nb, nbc, c[i]);
#pragma unroll
        for (int i = 0; i < BZ; ++i) {
            const bool ex = (ZW <= 16) ? (ez[i] > cz[i]) : (ztest && ez[i] > cz[i]);
            if (i == 0) n_ex0 += ex ? 1 : 0;
            if (i == BZ - 1) t_exT = ex ? min(t_exT, 2 * p) : t_exT;
            z[i] = ex ? ez[i] : cz[i];
        }
        const double pbz = __shfl_sync(0xffffffffu, z[BZ - 1], src_dn);
        double pbc = 0.0;
        if (BC > 0) pbc = __shfl_sync(0xffffffffu, c[BC - 1], src_dn);
        // ---- B level
        // the node below the window, at the level just computed: its exercise value (ub = u^-2) -- or, in the regime
        // without exercise nodes (ub = 0, see k_pin), the strike: what enters below node 0 there is never used by a
        // node of the tree, but it must stay ABOVE the values of the tree so that the live nodes remain a prefix of
        // the window (measure_band)
        const double eb = fma(sbg, ub, sKg);
        sbg *= kB.d_pu;
        sKg *= kB.inv_pu;
#pragma unroll
        for (int i = BZ - 1; i >= 0; --i) {
            if (i > 0) cz[i] = fma(rho2, z[i - 1], z[i]);
            else cz[i] = fma(rb_x, eb, fma(rb_nb, pbz, z[i]));
        }
#pragma unroll
        for (int i = 0; i < BZ; ++i) ez[i] = fma(sbg, ulo[i], sKg);
#pragma unroll
        for (int i = BC - 1; i >= 0; --i) c[i] = (i > 0) ? fma(rho2, c[i - 1], c[i]) : fma(rb_x, pbz, fma(rb_nb, pbc, c[i]));
#pragma unroll
        for (int i = 0; i < BZ; ++i) {
            const bool ex = (ZW <= 16) ? (ez[i] > cz[i]) : (ztest && ez[i] > cz[i]);
            if (i == 0) n_ex0 += ex ? 1 : 0;
            if (i == BZ - 1) t_exT = ex ? min(t_exT, 2 * p + 1) : t_exT;
            z[i] = ex ? ez[i] : cz[i];
        }
        if (BC > 0) nbc = __shfl_sync(0xffffffffu, c[0], src_up);
        nbz = __shfl_sync(0xffffffffu, z[0], src_up);
    }
    // exercise mask of the last level: a node chose exercise iff it holds the exercise value
    unsigned int mask = 0;
#pragma unroll
    for (int i = 0; i < BZ; ++i) mask |= (z[i] == fma(sbg, ulo[i], sKg)) ? (1u << i) : 0u;
#pragma unroll
    for (int i = 0; i < BZ; ++i) z[i] *= back;
#pragma unroll
    for (int i = 0; i < BC; ++i) {
        c[i] *= back;
#if defined(OPTP_ZONE_CHECK) || OPTP_ZONE_PAD_SHIFT != 2 || defined(OPTP_ZONE_PLAIN_IO)
        tile[OPTP_ZP(ZW + gl * BC + i)] = c[i];
#else
        // (a lane's chunk starts at a multiple of 4 nodes when BC is one: the pad of node base + i is then the pad
        // of the base plus i >> 2, a compile-time offset)
        if (BC % 4 == 0) tile[zone_pad(ZW + gl * BC) + i + (i >> 2)] = c[i];
        else tile[OPTP_ZP(ZW + gl * BC + i)] = c[i];
#endif
    }
    n_ex0_out = n_ex0;
    t_exT_out = t_exT;
    return mask;
}

template <int G, bool COUNT = false, bool RIG = false>
__global__ void __launch_bounds__(kZoneWarps * 32, OPTP_ZONE_MINB4) crr_zone_kernel(CrrArgs a) {
    // segment schedule: the measured one, or (RIG, the retry pass) the rigorous one of kZoneBelow - 2 levels
    constexpr int kSeg = RIG ? kZoneSegRig : kZoneSeg, kSeg0 = RIG ? kZoneSegRig : kZoneSeg0;
    constexpr int LG = 32 / G, BZ = zone_bz(G), ZW = zone_zw(G), BCMAX = kZoneLaneC, kMaxM = zone_max_m(G),
                  kGroupTile = zone_group_tile(G);
    constexpr int CH = zone_ch(G), BCU = zone_bcu(G);
    static_assert(ZW + CH * kMaxM <= zone_cap(G) && BCU * kMaxM <= BCMAX, "the tile holds Z and kMaxM chunks of C");
    // the node whose choice of continuation covers everything above it: the last node of Z -- in the drifting walk
    // node 15 of the window for every G (at G = 1 Z is 32 nodes wide; a test that depended on G would let a contract
    // pass in one layout and go to the retry pass, i.e. another rounding, in another)
    constexpr int ZT = kZoneDrift ? 16 : ZW;
    static_assert(ZT % BZ == 0 && ZT <= ZW, "the tested node is the last node of a lane");
    __shared__ double tiles[kZoneWarps][G * kGroupTile];
    __shared__ double stash_all[kZoneWarps][G][4];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int gl = lane & (LG - 1);        // lane within the option
    const int gb = lane & ~(LG - 1);       // first lane of the option
    const int src = gb | ((gl + 1) & (LG - 1));
    const int src_dn = gb | ((gl + LG - 1) & (LG - 1));  // (drifting walk: the previous lane of the option)
    (void)src_dn;
    const unsigned int gmask = (LG == 32) ? 0xffffffffu : ((1u << LG) - 1u);
    double* tile = tiles[wib] + (lane / LG) * kGroupTile;
    const int N = a.N;
    const unsigned long long warp0 = (unsigned long long)blockIdx.x * kZoneWarps + wib;
    const unsigned long long nwarps = (unsigned long long)gridDim.x * kZoneWarps;
    // retry passes walk the compacted list of contracts the previous layout was too narrow for
    const unsigned int n_work = a.wl_in ? *a.wl_in_n : (unsigned int)a.n;  // (launch_crr: n < 2^32)
    (void)warp0;
    (void)nwarps;
    while (true) {
        // EVERY lane issues the atomic (lane 0 adds one, the others zero): nvcc aggregates it into one ATOMG per
        // warp, and the REDUX puts the ticket in a uniform register.  With the atomic under `if (lane == 0)` --
        // or the ticket passed on by a shuffle -- ptxas no longer proves the warp converged and guards every
        // shuffle of the walk with a BRA.DIV (see the level-loop notes above).
        const unsigned int ticket = __reduce_min_sync(0xffffffffu, atomicAdd(a.next_group, lane == 0 ? 1u : 0u));
        const unsigned long long o0 = (unsigned long long)ticket * G;
        if (o0 >= n_work) break;
        const unsigned long long og = o0 + (unsigned long long)(lane / LG);
        const bool wanted = og < n_work;
        const unsigned long long ow = wanted ? og : (unsigned long long)(n_work - 1);
        const unsigned long long o = a.wl_in ? (unsigned long long)a.wl_in[ow] : ow;
        // An American call with q > 0 is priced as its DUAL PUT: on a CRR lattice (u*d = 1)
        //     C_k[j](S0, K, r, q) = P_k[k - j](S0' = K, K' = S0, r' = q, q' = r)
        // holds node by node (change of numeraire: p' = 1 - p*u*e^{-(r-q)dt} is the up-probability of
        // K/S, e^{-q dt} its discount factor; the exercise test is the same inequality divided by S/K'),
        // so the put machinery below prices it; the result differs from the direct recursion by
        // rounding only (1e-13 relative, tests/test_gpu_parity.py).
        const bool is_call = a.is_call[o] != 0;
        const double r_in = a.r[o], q_in = a.q[o];
        const bool dual = is_call && q_in > 0.0 && r_in >= 0.0;
        const double r = dual ? q_in : r_in, q = dual ? r_in : q_in;
        const double K = dual ? a.S0[o] : a.K[o];
        Lattice L = make_lattice_of(dual ? a.K[o] : a.S0[o], K, a.T[o], r, q, a.sigma[o], is_call && !dual, N);
        const double dt = a.T[o] / (double)N;
        if (dual) {  // status 1 is decided by the contract's own probability (binomial.py:15-17)
            const double d = 1.0 / L.u;
            const double p = (exp((r_in - q_in) * dt) - d) / (L.u - d);
            L.ok = L.ok && (p > 0.0) && (p < 1.0);
        }
        if (wanted && !L.ok && gl == 0) {
            a.price[o] = nan("");
            a.status[o] = 1;
        }
        const double one_m_disc = -expm1(-r * dt), one_m_dq = -expm1(-q * dt);
        // q > r: the exercise boundary sits at or below S1 = K(1 - disc)/(1 - e^{-q dt}) ~ K r/q from level
        // N-1 on (a node whose children are both exercise nodes is one iff S <= S1), i.e. it JUMPS down
        // from the strike between level N and N-1; the first segment is set up for that below.
        // j1 = first level-N node above S1 (N + 2 = no jump); needs ln u, computed here so that r, q and dt
        // do not stay live across the walk
        int j1 = N + 2;
        if (one_m_dq > one_m_disc)
            j1 = (int)fmin(fmax(floor(0.5 * ((double)N + log(K * (one_m_disc / one_m_dq) / L.S0) / log(L.u))) + 1.0, 0.0),
                           (double)(N + 1));
        // Drifting walk, q > 0: a node can only choose exercise if S < S1 (cont >= disc*K - S*e^{-q dt} whatever its
        // children chose).  On the levels k < k_pin even node 0 lies above S1, so NO node of the tree is an exercise
        // node: the window then starts at node 0 (it still drifts; what enters below node 0 never reaches a node of
        // the tree) and neither the bottom nor the top test applies.  k only falls, so this regime lasts to level 0.
        int k_pin = 0;
#ifndef OPTP_NO_PIN
        if (kZoneDrift && one_m_dq > 0.0)
            k_pin = (int)fmin(fmax(ceil(-log(K * (one_m_disc / one_m_dq) / L.S0) / log(L.u) - 0.01), 0.0), (double)(N + 1));
#endif
        bool eligible = wanted && L.ok && (L.sgn < 0.0) && q >= 0.0 && one_m_disc > 1e-9;
        // the scale of a segment stays within 2^+-200: pd^-kSeg -- drifting walk: (pd*pu)^-(kSeg/2), a pair of levels
        // scales by 1/(pd*pu).  (The retry pass walks shorter segments and takes what fails this on the long ones.)
        // `retry`: fails only the scale bound of THIS schedule -- for the retry pass, whoever shares its warp
        bool retry = false;
        if (kZoneDrift) {
            const bool scale_ok = (fabs(log(L.pd)) + fabs(log(L.pu))) * (double)kSeg <= 276.0;
            retry = eligible && !scale_ok && !RIG;
            eligible = eligible && scale_ok;
        } else {
            eligible = eligible && L.R >= kZoneSeg;
        }
        // An American call on an asset that pays no dividend is never exercised early: with q <= 0 and
        // r > 0, cont >= S*e^{-q dt} - K*disc > S - K at every node, so binomial.py:30-36 always keeps the
        // continuation value.  Such contracts (status 3) go to the European full-width kernel, which is
        // the same recursion without the per-node exercise value and compare.
        // (recomputed from the contract where it is needed: one register less across the walk)
        // The mirror image: an American put with r <= 0 and q >= 0 is never worth exercising early either
        // (cont >= disc*K - S*e^{-q dt} >= K - S), so it takes the European recursion too.
        auto declined_code = [&]() -> unsigned char {
            const double r0 = a.r[o], q0 = a.q[o];
            if (a.is_call[o] != 0) return (q0 <= 0.0 && -expm1(-r0 * (a.T[o] / (double)N)) > 1e-9) ? 3 : 2;
            return (r0 <= 0.0 && q0 >= 0.0) ? 3 : 2;
        };
        // an option that is not walked (declined, failed, past the end of the batch) rides along with
        // a harmless plan; `fail` keeps its result from being written
        bool fail = !eligible, too_wide = false;
        if (!__any_sync(0xffffffffu, eligible)) {
            if (wanted && L.ok && gl == 0) {
                if (retry && a.wl_rig) {
                    a.wl_rig[atomicAdd(a.wl_rig_n, 1u)] = (unsigned int)o;
                } else {
                    const unsigned char dc = declined_code();
                    a.status[o] = dc;
                    if (a.decl_count) atomicAdd(&a.decl_count[dc - 2], 1u);
                }
            }
            continue;
        }
        // exercise values: node m of Z (m = index - w0) is sgn*S0*u^(2*(w0 + (m & ~3)) - k) * u^(2*(m & 3)),
        // the first factor through exp(), the second from {1, u^2, u^4, u^6} -- the SAME two factors
        // whatever G is, so that a node's arithmetic does not depend on the lane layout
        double ulo[BZ];
        {
            const double u2 = L.u * L.u;
            const double t4[4] = {1.0, u2, u2 * u2, u2 * u2 * u2};
#pragma unroll
            for (int i = 0; i < BZ; ++i) {
                const int b = (gl * BZ + i) & 3;
                ulo[i] = (b == 0) ? t4[0] : (b == 1) ? t4[1] : (b == 2) ? t4[2] : t4[3];
            }
        }
        const int quad0 = (gl * BZ) & ~3;  // first node of this lane's quad, relative to w0
        DriftConsts kB;
        kB.inv_pu = 1.0 / L.pu;
        kB.rho2 = L.pd * kB.inv_pu;
        kB.d_pu = kB.inv_pu / L.u;
        kB.ulo_below = 1.0 / (L.u * L.u);
        // Four per-option constants that are only used BETWEEN segments live in shared memory, not in registers (the
        // level loops need every register: 28 C nodes + 2 Z nodes + 13 multipliers per lane at four options per warp)
        volatile double* const stash = stash_all[wib][lane / LG];
        {
            double bf = 1.0;  // pd^kSeg by square-and-multiply (drifting walk: (pd*pu)^(kSeg/2))
            double sq = kZoneDrift ? L.pd * L.pu : L.pd;
#pragma unroll
            for (int e = kZoneDrift ? kSeg / 2 : kSeg; e > 0; e >>= 1) {
                if (e & 1) bf *= sq;
                sq *= sq;
            }
            __syncwarp();  // (the previous group's readers are done)
            if (gl == 0) {
                stash[0] = L.sgn * L.S0;                  // (exact: sgn = -1 for every option that is walked)
                stash[1] = K * 8.2718061255302767e-25;    // 2^-80 K (K itself is not kept across the walk)
                stash[2] = log(L.u);
                stash[3] = bf;
            }
            __syncwarp();
        }
#define sS0 (stash[0])
#define thr (stash[1])
#define ln_u (stash[2])
#define back_full (stash[3])
        double z[BZ], c[BCMAX];
#pragma unroll
        for (int i = 0; i < BCMAX; ++i) c[i] = 0.0;
        // ---- level N: closed form max(K - S, 0) (binomial.py:20-25).  Z is placed so that its top
        // nodes are out of the money: everything above it is exactly 0, everything below it is an
        // exercise node
        int M = 0, k = N;
        int w0;
        unsigned int exmask = 0;
        {
            const double jz = 0.5 * ((double)N + log(K / L.S0) / ln_u);  // first out-of-the-money node
            // q > r and more than one option per warp: by level 0 the band reaches from below S1 to ~10.2
            // standard deviations above the strike, i.e. at least (jz - j1) + 5.1 sqrt(N) nodes.  A contract
            // that cannot fit this layout goes to the next one now instead of after a wasted partial walk.
            if (G > 1 && !fail && j1 <= N + 1 && (jz - (double)j1) + 4.6 * sqrt((double)N) + (double)zone_overhead() > (double)(CH * kMaxM))
                fail = too_wide = true;
            // (drifting walk: the origin may be negative -- the tree extended downwards, see OPTP_ZONE_DRIFT)
            // (the same origin whatever G is: the two factors of a node's exercise value depend on it)
            const double w = kZoneDrift ? fmin(fmax(floor(jz) - 12.0, -268435456.0), (double)(N + 1 - 16))
                                        : fmin(fmax(floor(jz) - 28.0, 0.0), (double)max(0, N + 1 - 32));
            w0 = fail ? 0 : (int)w;
            const double sbN = sS0 * exp(ln_u * (double)(2 * (w0 + quad0) - N));
#pragma unroll
            for (int i = 0; i < BZ; ++i) {
                const double e = fma(sbN, ulo[i], L.sK);
                z[i] = e > 0.0 ? e : 0.0;
                exmask |= (e > 0.0) ? (1u << i) : 0u;
            }
        }
        // Level k goes to the tile (node-ordered) and its band [jlo, jhi) is measured per option: once for
        // level N here, then after every segment.  (The level-N specials live OUTSIDE the walk loop: the loop
        // body is what the instruction caches have to hold.)
        int jlo = 0, jhi = 0;
#pragma unroll
        for (int i = 0; i < BZ; ++i) tile[OPTP_ZP(gl * BZ + i)] = z[i];  // level N (M = 0)
        __syncwarp();
        auto measure_band = [&]() {
            // ---- level k goes to the tile (node-ordered); measure its band [jlo, jhi) per option -----
            const int BC = BCU * M;
            // jlo = first node that did not choose exercise
            const unsigned int contm = ~exmask & ((1u << BZ) - 1u);
            const int first_cont = contm ? (__ffs(contm) - 1) : BZ;
            // (the boundary is looked for among the first ZT nodes whatever G is, see ZT)
            constexpr unsigned int zt_mask = (ZT / BZ >= 32) ? 0xffffffffu : ((1u << (ZT / BZ)) - 1u);
            const unsigned int has_cont = (__ballot_sync(0xffffffffu, first_cont < BZ) >> gb) & gmask & zt_mask;
            const int f_cont = has_cont ? (__ffs(has_cont) - 1) : 0;  // always one of the option's lanes
            const int s_cont = __shfl_sync(0xffffffffu, first_cont, gb + f_cont);
            jlo = has_cont ? (w0 + f_cont * BZ + s_cont) : (w0 + ZT);
            // jhi = 1 + last node worth more than 2^-80 K.  A put's value falls with the node index, so
            // the live nodes are a prefix of the window: find the last lane whose FIRST node is live
            // from the registers, then count the live nodes of that lane's chunk from the tile.
            const bool c0_live = (BC > 0) && c[0] > thr && (w0 + ZW + gl * BC) <= k;
            const bool z0_live = z[0] > thr && (w0 + gl * BZ) <= k;
            const unsigned int liveC = (__ballot_sync(0xffffffffu, c0_live) >> gb) & gmask;
            const unsigned int liveZ = (__ballot_sync(0xffffffffu, z0_live) >> gb) & gmask;
            int chunk0, chunk_n;  // the chunk that holds the last live node (window-relative), its length
            if (liveC) {
                chunk0 = ZW + (31 - __clz(liveC)) * BC;
                chunk_n = BC;
            } else {
                chunk0 = liveZ ? (31 - __clz(liveZ)) * BZ : 0;
                chunk_n = liveZ ? BZ : 0;
            }
            int n_live = 0;
            constexpr int kScanRounds = (BCMAX + LG - 1) / LG;
#pragma unroll
            for (int rr = 0; rr < kScanRounds; ++rr) {
                const int e = gl + LG * rr;
                bool live = false;
                if (e < chunk_n) live = tile[OPTP_ZP(chunk0 + e)] > thr && (w0 + chunk0 + e) <= k;
                n_live += __popc((__ballot_sync(0xffffffffu, live) >> gb) & gmask);
            }
            jhi = w0 + chunk0 + n_live;
            if (jhi < jlo) jhi = jlo;
            // the window bottom must still be inside the exercise region (or be node 0; the drifting walk runs on
            // the tree extended downwards, where there always is a node below)
            if ((kZoneDrift ? (k >= k_pin) : (w0 > 0)) && jlo <= w0) fail = true;
        };
        measure_band();
        unsigned int cnt_fp64 = 0, cnt_nodes = 0;  // per lane; COUNT builds only
        int width_old = ZW;  // M = 0 at level N
        if (k > 0 && __reduce_or_sync(0xffffffffu, (j1 <= N + 1 && !fail) ? 1u : 0u) != 0u) {
            // q > r (see `jump`): from level N-1 on the boundary is the first node above S1, far below
            // the strike.  Level N is re-tabulated in the tile from a lower origin, so that the
            // generic re-pack below finds the payoffs of the in-the-money continuation nodes there.
            const bool mine = !fail && j1 <= N + 1 && j1 < jlo;  // (a warp mate's jump must not touch this option)
            int wl = w0, span = 0;
            if (mine) {
                jlo = j1;
                wl = kZoneDrift ? ((N < k_pin) ? 0 : j1 - kZoneBelow) : max(0, j1 - kZoneBelow);
                span = w0 + ZW - wl;
                if (span > ZW + CH * kMaxM) fail = too_wide = true;
            }
            const bool refill = mine && !fail && wl < w0;
            if (!refill) span = 0;
            // (warp-uniform trip count: the shuffles of the walk need a provably converged warp)
            const int rounds = (__reduce_max_sync(0xffffffffu, span) + 4 * LG - 1) / (4 * LG);
            __syncwarp();  // the band scan above has finished reading the tile
            const double u2 = exp(2.0 * ln_u);  // (L.u is not kept live across the walk)
            for (int it = 0; it < rounds; ++it) {
                const int e0 = 4 * (gl + LG * it);
                const double sbq = sS0 * exp(ln_u * (double)(2 * (wl + e0) - N));
                double f = 1.0;
#pragma unroll
                for (int i = 0; i < 4; ++i, f *= u2) {
                    const double e = fma(sbq, f, L.sK);
                    if (e0 + i < span) tile[OPTP_ZP(e0 + i)] = e > 0.0 ? e : 0.0;
                }
            }
            if (refill) {
                w0 = wl;
                width_old = span;
            }
            __syncwarp();
        }
        while (true) {
            // (REDUX, not a vote: ptxas treats the result of __any_sync as divergent, and a divergent exit here
            // puts a BRA.DIV convergence check and a dozen register moves into every level loop below)
            // (the drifting walk steps in pairs of levels; an odd last level is taken in closed form after the loop)
            if (k <= (kZoneDrift ? 1 : 0) || __reduce_or_sync(0xffffffffu, fail ? 0u : 1u) == 0u) break;  // nothing left to walk in this warp
            // ---- next segment: levels k-1 ... k-nlev, fixed schedule -------------------------------
            const int nlev = kZoneDrift ? (min(k == N ? kSeg0 : kSeg, k) & ~1) : min(k == N ? kSeg0 : kSeg, k);
            const int top = min(jhi, k + 1);  // exclusive; node `top` of level k is 0 (or dead)
            // Z's bottom sits kZoneBelow nodes under the boundary (see kZoneSeg)
            const bool pinned = kZoneDrift && k < k_pin;  // no exercise node left in the tree (see k_pin)
            int w0n = kZoneDrift ? (pinned ? 0 : jlo - kZoneBelow) : max(0, jlo - kZoneBelow);
            int need = 0, n_chk = 0;
            // drifting walk: the window's top falls by nlev/2 nodes over the segment and must not fall below `top`
            const int drift = kZoneDrift ? (nlev >> 1) : 0;
            if (top + drift > w0n + ZW) {
                need = (top + drift - (w0n + ZW) + CH - 1) / CH;
                if (!fail && need > kMaxM) fail = too_wide = true;
            }
            // the top test is evaluated while the tested node exists: its index w0n + ZT - 1 (minus one per B level)
            // is at most the level.  (ZT, not ZW: whether a contract passes must not depend on G.)
            if (!pinned && top + drift > w0n + ZT)
                n_chk = kZoneDrift ? min(nlev, max(0, 2 * (k - (w0n + ZT - 1)))) : min(nlev, max(0, k - (w0n + ZT - 1)));
            if (fail) {
                w0n = 0;
                need = 0;
                n_chk = 0;
            }
            const int Mn = __reduce_max_sync(0xffffffffu, need);
            const int BCn = BCU * Mn;
            // ---- re-pack level k from (w0, M) to (w0n, Mn) ------------------------------------------
            // this lane's first Z node of the new layout at level k: sgn*S0*u^(2j - k).  (A running product
            // sb *= (u^2)^m instead of this exp() -- five predicated square-and-multiply steps -- was tried: no
            // faster (45.5 M against 45.8 M options/s), and it lets the exercise values drift by 3e-14.)
            const double sb = sS0 * exp(ln_u * (double)(2 * (w0n + quad0) - k));
#pragma unroll
            for (int i = 0; i < BZ; ++i) {
                const int j = w0n + gl * BZ + i;
                const int jr = j - w0;
                double val;
                if (jr < 0) val = fma(sb, ulo[i], L.sK);           // below the old window: exercise
                else if (jr >= width_old || j >= top) val = 0.0;   // above the band
                else val = tile[OPTP_ZP(jr)];
                z[i] = val;
            }
            const int jr0 = w0n + ZW + gl * BCn - w0;
            const int lim = max(0, min(width_old, top - w0));
            w0 = w0n;
            M = Mn;
            int n_ex0 = 0, t_exT = 0x7fffffff;
            double back = back_full;
            if (nlev != kSeg) {  // first or last (partial) segment: pd^nlev by square-and-multiply
                back = 1.0;
                double sq = kZoneDrift ? L.pd * L.pu : L.pd;
                for (int e = kZoneDrift ? (nlev >> 1) : nlev; e > 0; e >>= 1) {
                    if (e & 1) back *= sq;
                    sq *= sq;
                }
            }
#define OPTP_ZONE_CASE(m)                                                                                          \
    case m:                                                                                                        \
        if constexpr ((m) <= kMaxM) {                                                                              \
            if constexpr (kZoneDrift)                                                                              \
                exmask = zone_segment_drift<G, ((m) <= kMaxM ? (m) : 0)>(z, c, ulo, L, kB, pinned ? 0.0 : kB.ulo_below, gl, src, src_dn, sb, nlev >> 1, \
                                                                         back, tile, jr0, lim, n_ex0, t_exT);      \
            else                                                                                                   \
                exmask = zone_segment<G, ((m) <= kMaxM ? (m) : 0)>(z, c, ulo, L, gl, src, sb, nlev, back, tile, jr0, lim, n_ex0, t_exT); \
        }                                                                                                          \
        break;
            switch (M) {
                OPTP_ZONE_CASE(0) OPTP_ZONE_CASE(1) OPTP_ZONE_CASE(2) OPTP_ZONE_CASE(3) OPTP_ZONE_CASE(4)
                OPTP_ZONE_CASE(5) OPTP_ZONE_CASE(6) OPTP_ZONE_CASE(7) OPTP_ZONE_CASE(8) OPTP_ZONE_CASE(9)
                OPTP_ZONE_CASE(10) OPTP_ZONE_CASE(11) OPTP_ZONE_CASE(12) OPTP_ZONE_CASE(13) OPTP_ZONE_CASE(14)
                OPTP_ZONE_CASE(15) OPTP_ZONE_CASE(16) OPTP_ZONE_CASE(17) OPTP_ZONE_CASE(18) OPTP_ZONE_CASE(19)
                OPTP_ZONE_CASE(20) OPTP_ZONE_CASE(21) OPTP_ZONE_CASE(22) OPTP_ZONE_CASE(23) OPTP_ZONE_CASE(24)
                OPTP_ZONE_CASE(25) OPTP_ZONE_CASE(26) OPTP_ZONE_CASE(27) OPTP_ZONE_CASE(28)
                default: break;
            }
#undef OPTP_ZONE_CASE
#pragma unroll
            for (int i = 0; i < BZ; ++i) tile[OPTP_ZP(gl * BZ + i)] = z[i];
            __syncwarp();  // level k - nlev is in the tile
            // everything below the window is an exercise node only while the window's bottom node
            // chose exercise at EVERY level; everything in C is a continuation node only while Z's
            // top node chose continuation at every level it exists
            const int bottom_ex = __shfl_sync(0xffffffffu, n_ex0, gb);
            const int top_ex = __shfl_sync(0xffffffffu, t_exT, gb + ZT / BZ - 1);
            if ((kZoneDrift ? !pinned : (w0 > 0)) && bottom_ex != nlev) fail = true;
            if (top_ex < n_chk) fail = true;  // (exercise events while Z's top node is outside the tree do not count)
            if (COUNT) {
                // per lane and level: Z nodes 2 DFMA + 1 DSETP each, C nodes 1 DFMA each, 2 scaling DMULs (drifting
                // walk, per pair of levels: the exercise value below the window and the cross links, 3 or 5 DFMAs)
                cnt_fp64 += (unsigned int)(nlev * (3 * BZ + BCU * M + 2)) + (kZoneDrift ? (unsigned int)((nlev >> 1) * (M > 0 ? 5 : 3)) : 0u);
                cnt_nodes += (unsigned int)(nlev * (BZ + BCU * M));
            }
            k -= nlev;
            if (kZoneDrift) w0 -= nlev >> 1;  // the origin fell by one node per B level
            measure_band();
            width_old = ZW + CH * M;
        }
        if (COUNT && lane == 0 && a.op_count) {
            atomicAdd(&a.op_count[0], 32ull * cnt_fp64);
            atomicAdd(&a.op_count[1], 32ull * cnt_nodes);
        }
        // level 0: node 0 is in the window (w0 == 0) or below it (then it is an exercise node)
        double v0 = __shfl_sync(0xffffffffu, z[0], gb);
        bool v0_below = (w0 != 0);
        if (kZoneDrift) {
            // the tile holds level k (0, or 1 when N is odd) in node order; node j sits at j - w0, is an exercise node
            // below the window and dead above it
            auto node = [&](int j) -> double {
                const int pos = j - w0;
                if (pos < 0) return fma(sS0, exp(ln_u * (double)(2 * j - k)), L.sK);
                if (pos >= width_old) return 0.0;
                return tile[OPTP_ZP(pos)];
            };
            v0 = node(0);
            v0_below = (w0 > 0);
            if (k == 1) {  // binomial.py:28-36 for the one node of level 0
                const double cont = fma(L.pu, node(1), L.pd * v0);
                const double e0 = sS0 + L.sK;
                v0 = e0 > cont ? e0 : cont;
                v0_below = false;
            }
            __syncwarp();  // the tile is free for the next group of contracts
        }
        if (wanted && L.ok && gl == 0) {
            if (fail) {
                // too wide for this lane layout: the next narrower pass walks it.  Otherwise, on the measured
                // schedule, a walk that failed one of its per-level tests -- almost always the boundary outran the
                // schedule (small r dt: the bottom test) -- is settled by the rigorous-schedule pass; contracts that
                // never qualified ride along and are declined there.  (One append path: status is settled later.)
                unsigned int* const list = (too_wide && G > 1) ? a.wl_out : (!too_wide && !RIG) ? a.wl_rig : nullptr;
                if (list) {
                    list[atomicAdd((too_wide && G > 1) ? a.wl_out_n : a.wl_rig_n, 1u)] = (unsigned int)o;
                } else {
                    const unsigned char dc = declined_code();
                    a.status[o] = dc;
                    if (a.decl_count) atomicAdd(&a.decl_count[dc - 2], 1u);
                }
            } else {
                // below the window at level 0: immediate exercise, K - S0 (for a call, through its dual put, S0 - K)
                const double intrinsic = (a.is_call[o] != 0) ? (a.S0[o] - a.K[o]) : (a.K[o] - a.S0[o]);
                a.price[o] = v0_below ? intrinsic : v0;
                a.status[o] = 0;
            }
        }
    }
}

// ---------------------------------------------------------------- generic N: one option per CTA
#undef sS0
#undef thr
#undef ln_u
#undef back_full

template <bool AMER>
__global__ void __launch_bounds__(256) crr_cta_kernel(CrrArgs a) {
    extern __shared__ double sm[];
    const int N = a.N;
    const size_t stride = (size_t)N + 2;
    double* base = a.cta_scratch ? a.cta_scratch + (size_t)blockIdx.x * 3 * stride : sm;
    double* bufA = base;
    double* bufB = base + stride;
    double* W = base + 2 * stride;  // u^(2j)
    if (a.todo_only && a.decl_count && a.decl_count[a.todo_only - 2] == 0) return;  // nothing was left for this pass
    for (unsigned long long o = blockIdx.x; o < a.n; o += gridDim.x) {
        if (a.todo_only && a.status[o] != a.todo_only) continue;  // (block-uniform)
        const Lattice L = make_lattice(a, o);
        if (!L.ok) {
            if (threadIdx.x == 0) {
                a.price[o] = nan("");
                a.status[o] = 1;
            }
            continue;
        }
        __syncthreads();
        double sbk = L.sgn * L.S0 * pow(L.u, (double)(-N));
        for (int j = threadIdx.x; j <= N; j += blockDim.x) {
            const double w = pow(L.u, (double)(2 * j));
            W[j] = w;
            bufA[j] = pick_max(0.0, fma(sbk, w, L.sK));
        }
        __syncthreads();
        double* src = bufA;
        double* dst = bufB;
        for (int k = N - 1; k >= 0; --k) {
            sbk *= L.u;
            for (int j = threadIdx.x; j <= k; j += blockDim.x) {
                double c = fma(L.pd, src[j], L.pu * src[j + 1]);
                if (AMER) c = pick_max(c, fma(sbk, W[j], L.sK));
                dst[j] = c;
            }
            __syncthreads();
            double* t = src;
            src = dst;
            dst = t;
        }
        if (threadIdx.x == 0) {
            a.price[o] = src[0];
            a.status[o] = 0;
        }
    }
}

// ---------------------------------------------------------------- European: the level-0 value in O(N)
// Without an exercise test the backward induction of binomial.py:28-29 is linear, and its level-0 value is
// the discounted binomial expectation of the terminal payoff,
//     V0 = e^{-rT} * sum_j C(N,j) p^j (1-p)^(N-j) * max(+-(S0 u^(2j-N) - K), 0),
// which is what N(N+1)/2 node updates compute in another order.  One warp per option: lane l takes a
// contiguous chunk of nodes, starts from the weight exp(lgamma(N+1) - lgamma(j+1) - lgamma(N-j+1) +
// j ln p + (N-j) ln(1-p)) and from S0 u^(2j-N) = S0 exp(ln u (2j-N)), and walks its chunk with the exact
// ratios w_{j+1}/w_j = (N-j)/(j+1) * p/(1-p) and S_{j+1}/S_j = u^2.  Rounding: |ln w| <= ~7 N, so the weights
// carry ~7 N eps (8e-13 at N = 1000) -- the same order as the tree's own N roundings per path; used for
// N <= 20000 (2e-11), deeper trees keep the recursion.  `optp_set_option("crr_euro_sum", 0)` restores the
// recursion for every N (A/B tests).
constexpr int kEuroSumMaxN = 20000;
__global__ void __launch_bounds__(128) crr_euro_sum_kernel(CrrArgs a) {
    if (a.todo_only && a.decl_count && a.decl_count[a.todo_only - 2] == 0) return;  // nothing was left for this pass
    const int lane = threadIdx.x & 31;
    const unsigned long long warp0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    const int N = a.N;
    const int chunk = (N + 1 + 31) / 32;
    const double lgN = lgamma((double)N + 1.0);
    for (unsigned long long o = warp0; o < a.n; o += nwarps) {
        if (a.todo_only && a.status[o] != a.todo_only) continue;
        const double S0 = a.S0[o], K = a.K[o], T = a.T[o], r = a.r[o], q = a.q[o], sg = a.sigma[o];
        const double sgn = a.is_call[o] ? 1.0 : -1.0;
        const double dt = T / (double)N;
        const double ln_u = sg * sqrt(dt);
        const double u = exp(ln_u), d = 1.0 / u;
        const double p = (exp((r - q) * dt) - d) / (u - d);  // binomial.py:15
        if (!((p > 0.0) && (p < 1.0) && isfinite(S0) && isfinite(K))) {
            if (lane == 0) {
                a.price[o] = nan("");
                a.status[o] = 1;
            }
            continue;
        }
        const int j0 = lane * chunk;
        const int j1 = min(j0 + chunk, N + 1);
        double acc = 0.0;
        if (j0 < j1) {
            const double lp = log(p), lq = log1p(-p);
            double w = exp(lgN - lgamma((double)j0 + 1.0) - lgamma((double)(N - j0) + 1.0) + (double)j0 * lp +
                           (double)(N - j0) * lq);
            double S = S0 * exp(ln_u * (double)(2 * j0 - N));
            const double odds = p / (1.0 - p), u2 = u * u;
            for (int j = j0; j < j1; ++j) {
                const double pay = sgn * (S - K);
                acc = fma(w, pay > 0.0 ? pay : 0.0, acc);
                w *= odds * ((double)(N - j) / (double)(j + 1));
                S *= u2;
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            a.price[o] = exp(-r * T) * acc;
            a.status[o] = 0;
        }
    }
}

// cudaOccupancyMaxActiveBlocksPerMultiprocessor costs microseconds per call and its answer never changes for
// a (kernel, block size, dynamic shared memory = 0) triple: remember it (calls are serialised by the API lock).
static int occupancy_cached(const void* fn, int threads, int* per_sm) {
    static const void* keys[16];
    static int vals[16];
    static int n_keys = 0;
    for (int i = 0; i < n_keys; ++i)
        if (keys[i] == fn) {
            *per_sm = vals[i];
            return 0;
        }
    int v = 0;
    OPTP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, fn, threads, 0));
    if (v < 1) v = 1;
    if (n_keys < 16) {
        keys[n_keys] = fn;
        vals[n_keys++] = v;
    }
    *per_sm = v;
    return 0;
}

static int launch_full_warp(const CrrArgs& a, int american, cudaStream_t st) {
    Context& c = ctx();
    const unsigned long long want = (a.n + kWarpsPerCta - 1) / kWarpsPerCta;
    int per_sm = 0;
    if (int e = occupancy_cached(american ? (const void*)crr_warp_kernel<true> : (const void*)crr_warp_kernel<false>,
                                 kWarpsPerCta * 32, &per_sm))
        return e;
    // persistent warps: exactly one resident wave, options strided over warps
    const unsigned long long cap = (unsigned long long)c.n_sm * per_sm;
    const unsigned int grid = (unsigned int)(want < cap ? want : cap);
    if (american)
        crr_warp_kernel<true><<<grid, kWarpsPerCta * 32, 0, st>>>(a);
    else
        crr_warp_kernel<false><<<grid, kWarpsPerCta * 32, 0, st>>>(a);
    c.launches++;
    OPTP_CUDA(cudaGetLastError());
    return 0;
}

// Full-width pass (honours a.todo_only): one option per warp up to N = 1023, one per CTA beyond.
static int launch_full(const CrrArgs& a_in, int american, cudaStream_t st, int slot) {
    Context& c = ctx();
    if (!american && a_in.N <= kEuroSumMaxN && !c.crr_no_euro_sum) {
        const unsigned long long want = (a_in.n + 3) / 4;  // four warps (options in flight) per CTA
        const unsigned long long cap = (unsigned long long)c.n_sm * 16;
        crr_euro_sum_kernel<<<(unsigned int)(want < cap ? want : cap), 128, 0, st>>>(a_in);
        c.launches++;
        OPTP_CUDA(cudaGetLastError());
        return 0;
    }
    if (a_in.N <= 1023) return launch_full_warp(a_in, american, st);
    CrrArgs a = a_in;
    size_t smem = sizeof(double) * 3 * (size_t)(a.N + 2);
    const unsigned long long cap = (unsigned long long)c.n_sm * 8;
    unsigned int grid = (unsigned int)(a.n < cap ? a.n : cap);
    a.cta_scratch = nullptr;
    if (smem > 227 * 1024) {
        // deeper than shared memory holds (N > 9685): ping-pong level arrays in global memory, at most 1 GiB per
        // pipeline slot (they stay in L2 up to N ~ 10^4 x 148 CTAs)
        const size_t per_cta = smem, budget = (size_t)1 << 30;
        const size_t fit = budget / per_cta;
        if (fit < 1) return set_error(OPTP_E_TOO_LARGE, "N=%d needs %zu B of level storage per option", a.N, per_cta);
        if (grid > fit) grid = (unsigned int)fit;
        const size_t slot_bytes = ((size_t)grid * per_cta + 255) & ~(size_t)255;
        if (int e = arena_reserve(c.cta_scratch, 2 * slot_bytes)) return e;
        a.cta_scratch = (double*)((char*)c.cta_scratch.p + (size_t)slot * ((c.cta_scratch.cap / 2) & ~(size_t)255));
        smem = 0;
    }
    if (american) {
        if (smem) OPTP_CUDA(cudaFuncSetAttribute(crr_cta_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        crr_cta_kernel<true><<<grid, 256, smem, st>>>(a);
    } else {
        if (smem) OPTP_CUDA(cudaFuncSetAttribute(crr_cta_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        crr_cta_kernel<false><<<grid, 256, smem, st>>>(a);
    }
    c.launches++;
    OPTP_CUDA(cudaGetLastError());
    return 0;
}

// Work lists of the two-zone kernel: per slot two lists of {count, n items}.
static size_t crr_worklist_bytes(unsigned long long n) {
    return ((size_t)n * sizeof(unsigned int) + 255) & ~(size_t)255;
}
static int crr_reserve_worklists(unsigned long long n) {
    return arena_reserve(ctx().work, 2 * (256 + 3 * crr_worklist_bytes(n)) + 512);  // two slots of {header, 3 lists}
}

// the two-zone kernel for G options per warp; COUNT = with the executed-work counters (diagnostics)
static const void* zone_fn(int G, bool count) {
    if (count) return (G == 4) ? (const void*)crr_zone_kernel<4, true> : (G == 2) ? (const void*)crr_zone_kernel<2, true>
                                                                               : (const void*)crr_zone_kernel<1, true>;
    return (G == 4) ? (const void*)crr_zone_kernel<4> : (G == 2) ? (const void*)crr_zone_kernel<2>
                                                                 : (const void*)crr_zone_kernel<1>;
}
static const void* zone_rig_fn(bool count) {
    return count ? (const void*)crr_zone_kernel<1, true, true> : (const void*)crr_zone_kernel<1, false, true>;
}
static void zone_rig_launch(bool count, unsigned int grid, const CrrArgs& a, cudaStream_t st) {
    if (count) crr_zone_kernel<1, true, true><<<grid, dim3(kZoneWarps * 32), 0, st>>>(a);
    else crr_zone_kernel<1, false, true><<<grid, dim3(kZoneWarps * 32), 0, st>>>(a);
}
static void zone_launch(int G, bool count, unsigned int grid, const CrrArgs& a, cudaStream_t st) {
    const dim3 blk(kZoneWarps * 32);
    if (count) {
        if (G == 4) crr_zone_kernel<4, true><<<grid, blk, 0, st>>>(a);
        else if (G == 2) crr_zone_kernel<2, true><<<grid, blk, 0, st>>>(a);
        else crr_zone_kernel<1, true><<<grid, blk, 0, st>>>(a);
    } else {
        if (G == 4) crr_zone_kernel<4><<<grid, blk, 0, st>>>(a);
        else if (G == 2) crr_zone_kernel<2><<<grid, blk, 0, st>>>(a);
        else crr_zone_kernel<1><<<grid, blk, 0, st>>>(a);
    }
}

// `slot` (0 or 1) selects the work-list scratch: launches that may be in flight together (the two
// streams of the sliced host pipeline) must use different slots.
static int launch_crr(const CrrArgs& a_in, int american, cudaStream_t st, int slot = 0) {
    Context& c = ctx();
    CrrArgs a = a_in;
    a.todo_only = 0;
    a.wl_in = a.wl_in_n = nullptr;
    a.wl_out = a.wl_out_n = nullptr;
    a.wl_rig = a.wl_rig_n = nullptr;
    a.decl_count = nullptr;
    a.cta_scratch = nullptr;
    a.op_count = nullptr;
    a.next_group = nullptr;
    if (a.n == 0) return 0;
    if (a.n >= (1ull << 32)) return set_error(OPTP_E_TOO_LARGE, "a CRR batch holds at most 2^32 - 1 contracts");
    if (american && !c.crr_no_band) {
        // pass 1: banded kernel (American puts with r > q); pass 2: whatever it left at status 2
        int per_sm = 0;
        // options per warp of the two-zone kernel: 4 once the batch fills the machine that way
        int G = 0;
        if (c.crr_band_kernel == 2 || a.N > 1023) {  // (the single-zone kernel stops at N = 1023)
            G = c.crr_zone_group;
            if (G == 0) {
                const unsigned long long warps = (unsigned long long)c.n_sm * 16;
                G = (a.n >= 4 * warps) ? 4 : (a.n >= 2 * warps) ? 2 : 1;
                // by level 0 a band spans ~10.2 standard deviations above the strike, 5.1 sqrt(N) nodes: where that
                // cannot fit a lane layout, the pass in that layout would be walked in vain (N = 5000: a quarter of it)
                const double width = 4.6 * sqrt((double)a.N) + (double)zone_overhead();
                while (G > 1 && width > (double)(zone_ch(G) * zone_max_m(G))) G /= 2;
            }
        }
        unsigned int* wl[2] = {nullptr, nullptr};
        unsigned int* hdr = nullptr;
        if (G >= 1) {
            if (int e = crr_reserve_worklists(a.n)) return e;
            const size_t list_bytes = crr_worklist_bytes(a.n);
            char* base = (char*)c.work.p + (size_t)slot * ((c.work.cap / 2) & ~(size_t)255);
            // header {list 0 count, list 1 count, contracts left at status 2, at status 3}, then the two lists
            hdr = (unsigned int*)base;
            wl[0] = (unsigned int*)(base + 256);
            wl[1] = (unsigned int*)(base + 256 + list_bytes);
            a.decl_count = hdr + 2;
            a.next_group = hdr + 4;   // hdr[4..6]: the ticket counters of the (up to three) two-zone passes
            a.wl_rig = (unsigned int*)(base + 256 + 2 * list_bytes);  // hdr[7]: its count, hdr[8]: its pass's tickets
            a.wl_rig_n = hdr + 7;
            OPTP_CUDA(cudaMemsetAsync(hdr, 0, 16 * sizeof(unsigned int), st));
            if (G >= 2) {
                a.wl_out = wl[0];
                a.wl_out_n = hdr;
            }
        }
        const bool count = c.crr_count_ops && G >= 1;
        if (count) {
            if (!c.crr_op_count) {
                OPTP_CUDA(cudaMalloc(&c.crr_op_count, 2 * sizeof(unsigned long long)));
                OPTP_CUDA(cudaMemsetAsync(c.crr_op_count, 0, 2 * sizeof(unsigned long long), st));
            }
            a.op_count = c.crr_op_count;
        }
        const void* fn = (G >= 1) ? zone_fn(G, count) : (const void*)crr_band_kernel;
        const int cta_warps = (G >= 1) ? kZoneWarps : kBandWarps;
        if (int e = occupancy_cached(fn, cta_warps * 32, &per_sm)) return e;
        const unsigned long long per_warp = (G > 0) ? (unsigned long long)G : 1ull;
        const unsigned long long want = (a.n + cta_warps * per_warp - 1) / (cta_warps * per_warp);
        const unsigned long long cap = (unsigned long long)c.n_sm * per_sm;
        const unsigned int grid = (unsigned int)(want < cap ? want : cap);
        if (G >= 1)
            zone_launch(G, count, grid, a, st);
        else
            crr_band_kernel<<<grid, kBandWarps * 32, 0, st>>>(a);
        c.launches++;
        OPTP_CUDA(cudaGetLastError());
        // contracts whose band was too wide for G options per warp were appended to wl[0]: walk that
        // list two per warp, then what is too wide even so (wl[1]) one per warp
        int list = 0;
        for (int g2 = G / 2; g2 >= 1; g2 /= 2, ++list) {
            a.wl_in = wl[list];
            a.wl_in_n = hdr + list;
            a.next_group = hdr + 5 + list;
            a.wl_out = (g2 == 2) ? wl[1] : nullptr;
            a.wl_out_n = (g2 == 2) ? hdr + 1 : nullptr;
            if (int e = occupancy_cached(zone_fn(g2, count), kZoneWarps * 32, &per_sm)) return e;
            const unsigned long long want2 = (a.n + kZoneWarps * g2 - 1) / (kZoneWarps * g2);
            const unsigned long long cap2 = (unsigned long long)c.n_sm * per_sm;
            const unsigned int grid2 = (unsigned int)(want2 < cap2 ? want2 : cap2);
            zone_launch(g2, count, grid2, a, st);
            c.launches++;
            OPTP_CUDA(cudaGetLastError());
        }
        if (G >= 1) {
            // contracts whose boundary outran the measured schedule: one option per warp on the rigorous schedule
            a.wl_in = a.wl_rig;
            a.wl_in_n = a.wl_rig_n;
            a.wl_out = a.wl_out_n = nullptr;
            a.wl_rig = a.wl_rig_n = nullptr;
            a.next_group = hdr + 8;
            if (int e = occupancy_cached(zone_rig_fn(count), kZoneWarps * 32, &per_sm)) return e;
            const unsigned long long want3 = (a.n + kZoneWarps - 1) / kZoneWarps;
            const unsigned long long cap3 = (unsigned long long)c.n_sm * per_sm;
            zone_rig_launch(count, (unsigned int)(want3 < cap3 ? want3 : cap3), a, st);
            c.launches++;
            OPTP_CUDA(cudaGetLastError());
        }
        a.wl_in = a.wl_in_n = nullptr;
        a.wl_out = a.wl_out_n = nullptr;
        a.todo_only = 0;
        if (c.crr_band_only) return 0;  // diagnostics: leave the declined contracts unpriced
        if (G > 0) {  // two-zone kernel: never-exercised calls were marked 3 -> European recursion
            a.todo_only = 3;
            if (int e = launch_full(a, 0, st, slot)) return e;
        }
        a.todo_only = 2;
    }
    return launch_full(a, american, st, slot);
}

}  // namespace optp

using namespace optp;

extern "C" int optp_crr_batch_dev(const double* S0, const double* K, const double* T,
                                  const double* r, const double* q, const double* sigma,
                                  const uint8_t* is_call, uint64_t n, uint32_t N, int american,
                                  double* price_out, uint8_t* status_out, void* stream) {
    OPTP_ENTER();
    if (N == 0 || N > (1u << 24)) return set_error(OPTP_E_ARG, "N must be positive.");
    if (n && (!S0 || !K || !T || !r || !q || !sigma || !is_call || !price_out || !status_out))
        return set_error(OPTP_E_ARG, "null pointer argument");
    CrrArgs a{S0, K, T, r, q, sigma, is_call, (unsigned long long)n, (int)N, price_out, status_out, 0};
    return launch_crr(a, american, (cudaStream_t)stream);
}

extern "C" int optp_crr_batch(const double* S0, const double* K, const double* T, const double* r,
                              const double* q, const double* sigma, const uint8_t* is_call,
                              uint64_t n, uint32_t N, int american, double* price_out,
                              uint8_t* status_out) {
    OPTP_ENTER();
    if (N == 0 || N > (1u << 24)) return set_error(OPTP_E_ARG, "N must be positive.");
    if (n == 0) return 0;
    if (!S0 || !K || !T || !r || !q || !sigma || !is_call || !price_out || !status_out)
        return set_error(OPTP_E_ARG, "null pointer argument");
    Context& c = ctx();
    OPTP_CUDA(cudaSetDevice(c.device));
    // contract batch: six fp64 columns + one byte column, packed into one arena (SoA)
    const size_t col = (size_t)n * sizeof(double);
    const size_t colp = (col + 255) & ~(size_t)255;
    if (int e = arena_reserve(c.in0, 6 * colp + n + 256)) return e;
    if (int e = arena_reserve(c.out0, colp + n + 256)) return e;
    char* in = (char*)c.in0.p;
    char* out = (char*)c.out0.p;
    const double* src[6] = {S0, K, T, r, q, sigma};
    // Small batches (a scalar crr() is a batch of one): the seven columns are packed into pinned memory and
    // travel in ONE copy each way -- the call is latency-bound, and a copy costs ~5 us whatever its size.
    if (n <= 8192) {
        if (int e = host_arena_reserve(c.pin_in, 6 * colp + n)) return e;
        if (int e = host_arena_reserve(c.pin_out, colp + n)) return e;
        char* hin = (char*)c.pin_in.p;
        char* hout = (char*)c.pin_out.p;
        for (int i = 0; i < 6; ++i) memcpy(hin + i * colp, src[i], col);
        memcpy(hin + 6 * colp, is_call, n);
        OPTP_CUDA(cudaMemcpyAsync(in, hin, 6 * colp + n, cudaMemcpyHostToDevice, c.stream));
        CrrArgs a{(const double*)(in), (const double*)(in + colp), (const double*)(in + 2 * colp),
                  (const double*)(in + 3 * colp), (const double*)(in + 4 * colp), (const double*)(in + 5 * colp),
                  (const uint8_t*)(in + 6 * colp), (unsigned long long)n, (int)N, (double*)out,
                  (uint8_t*)(out + colp), 0};
        if (int e = launch_crr(a, american, c.stream, 0)) return e;
        OPTP_CUDA(cudaMemcpyAsync(hout, out, colp + n, cudaMemcpyDeviceToHost, c.stream));
        OPTP_CUDA(cudaStreamSynchronize(c.stream));
        memcpy(price_out, hout, col);
        memcpy(status_out, hout + colp, n);
        return 0;
    }
    // Large batches are walked in slices of 128k contracts through FOUR streams chained by events: host->device
    // copies on one, the kernels of even / odd slices on two (so the tail of one slice's persistent grid
    // overlaps the head of the next), device->host copies (into pinned staging) on the fourth.  The copies
    // of slices i+1 and i-1 overlap the kernels of slice i, and the host never blocks on a pageable copy: it
    // moves a slice's results from the staging buffer to the caller's arrays as soon as that slice's event
    // has fired.  A contract's price does not depend on the slice it is in.
    // (OPTP_CRR_SLICE_LOG2 in the environment: slice size for A/B runs, 14 ... 20)
    static const int slice_log2 = [] {
        const char* e = getenv("OPTP_CRR_SLICE_LOG2");
        const int v = e ? atoi(e) : 17;
        return (v >= 14 && v <= 20) ? v : 17;
    }();
    const uint64_t slice = 1ull << slice_log2;
    const uint64_t step = (n >= 4 * slice) ? slice : n;
    constexpr uint64_t kGroupSlices = kMaxEvents / 3;  // slices in flight per group (bounds staging and events)
    if (!c.stream2) OPTP_CUDA(cudaStreamCreateWithFlags(&c.stream2, cudaStreamNonBlocking));
    if (!c.stream3) OPTP_CUDA(cudaStreamCreateWithFlags(&c.stream3, cudaStreamNonBlocking));
    if (!c.stream4) OPTP_CUDA(cudaStreamCreateWithFlags(&c.stream4, cudaStreamNonBlocking));
    while (c.n_events < kMaxEvents) {
        OPTP_CUDA(cudaEventCreateWithFlags(&c.events[c.n_events], cudaEventDisableTiming));
        c.n_events++;
    }
    // scratch is reserved before anything is in flight (an arena never grows under a running kernel)
    if (american)
        if (int e = crr_reserve_worklists(step)) return e;
    // (the very first slice is a quarter of the others: nothing runs on the device until its inputs have arrived)
    const uint64_t head = (step == slice) ? slice / 4 : step;
    const uint64_t group = step * (kGroupSlices - 1);
    auto slice_len = [&](uint64_t lo, uint64_t g1) -> uint64_t {
        const uint64_t want = (lo == 0) ? head : step;
        return (g1 - lo < want) ? (g1 - lo) : want;
    };
    const size_t gcol = (((size_t)(n < group ? n : group) * sizeof(double)) + 255) & ~(size_t)255;
    if (int e = host_arena_reserve(c.pin_out, gcol + (size_t)(n < group ? n : group))) return e;
    char* hout = (char*)c.pin_out.p;
    for (uint64_t g0 = 0; g0 < n; g0 += group) {
        const uint64_t g1 = (n - g0 < group) ? n : g0 + group;
        int s_idx = 0;
        for (uint64_t lo = g0, m = 0; lo < g1; lo += m, ++s_idx) {
            m = slice_len(lo, g1);
            cudaEvent_t ev_in = c.events[3 * s_idx], ev_k = c.events[3 * s_idx + 1], ev_out = c.events[3 * s_idx + 2];
            for (int i = 0; i < 6; ++i)
                OPTP_CUDA(cudaMemcpyAsync(in + i * colp + lo * sizeof(double), src[i] + lo, m * sizeof(double),
                                          cudaMemcpyHostToDevice, c.stream2));
            OPTP_CUDA(cudaMemcpyAsync(in + 6 * colp + lo, is_call + lo, m, cudaMemcpyHostToDevice, c.stream2));
            OPTP_CUDA(cudaEventRecord(ev_in, c.stream2));
            cudaStream_t sc = (s_idx & 1) ? c.stream4 : c.stream;
            OPTP_CUDA(cudaStreamWaitEvent(sc, ev_in, 0));
            CrrArgs a{(const double*)(in) + lo,
                      (const double*)(in + colp) + lo,
                      (const double*)(in + 2 * colp) + lo,
                      (const double*)(in + 3 * colp) + lo,
                      (const double*)(in + 4 * colp) + lo,
                      (const double*)(in + 5 * colp) + lo,
                      (const uint8_t*)(in + 6 * colp) + lo,
                      (unsigned long long)m,
                      (int)N,
                      (double*)out + lo,
                      (uint8_t*)(out + colp) + lo,
                      0};
            if (int e = launch_crr(a, american, sc, s_idx & 1)) return e;
            OPTP_CUDA(cudaEventRecord(ev_k, sc));
            OPTP_CUDA(cudaStreamWaitEvent(c.stream3, ev_k, 0));
            OPTP_CUDA(cudaMemcpyAsync(hout + (lo - g0) * sizeof(double), (double*)out + lo, m * sizeof(double),
                                      cudaMemcpyDeviceToHost, c.stream3));
            OPTP_CUDA(cudaMemcpyAsync(hout + gcol + (lo - g0), out + colp + lo, m, cudaMemcpyDeviceToHost, c.stream3));
            OPTP_CUDA(cudaEventRecord(ev_out, c.stream3));
        }
        s_idx = 0;
        for (uint64_t lo = g0, m = 0; lo < g1; lo += m, ++s_idx) {
            m = slice_len(lo, g1);
            OPTP_CUDA(cudaEventSynchronize(c.events[3 * s_idx + 2]));
            memcpy(price_out + lo, hout + (lo - g0) * sizeof(double), m * sizeof(double));
            memcpy(status_out + lo, hout + gcol + (lo - g0), m);
        }
    }
    OPTP_CUDA(cudaStreamSynchronize(c.stream));
    return 0;
}
