// Production Monte Carlo kernels (sm_100a): fp32 path state, Philox4x32-10 + Box-Muller
// normals built in registers, payoff statistics accumulated in fp64.  Paths are never
// written to HBM: each thread carries one base path AND its antithetic mirror (the mirror
// re-uses the normal, so the RNG cost is halved per effective path-step), keeps the running
// path state the payoffs need (max / min / sum S / sum ln S) in registers, and contributes
// (X, X^2, XY, Y, Y^2) once per path to block-level fp64 partials.
//
//   terminal_fp32_kernel   <- _mc_chunk_sumstats          monte_carlo.py:28-64
//   path_fp32_kernel<GBM>  <- gbm_paths + exotics          processes.py:44-55, exotics.py:18-273
//   path_fp32_kernel<HES>  <- heston_paths + exotics       processes.py:134-160
//   path_fp32_kernel<GBMM> <- gbm_milstein_paths           processes.py:284-303
//   lv_fp32_kernel<LVM>    <- milstein_local_vol_paths     processes.py:345-379
//   lv_fp32_kernel<LVE>    <- local_vol_paths              processes.py:240-256
//
// All log-space work is done in base 2 so exp/log are single MUFU.EX2 / MUFU.LG2 issues.
// The Philox counter layout (path, step block, stream) is the same one mc_fp64.cu uses in
// FP64_PHILOX mode, so both modes see the same normals and differ only in arithmetic.
#include <math.h>
#include <stdlib.h>
#include <type_traits>
#include "mc_kernels.cuh"

namespace optp {

constexpr uint32_t kTermStream = 0x7465726du;
constexpr float kLog2e = 1.4426950408889634f;

// per-step path state requested by the payoffs (finer than OBS_*: max and min are split so
// an up-barrier does not pay for the running minimum)
enum : int { M_MAX = 1, M_MIN = 2, M_SUM = 4, M_LOG = 8 };

static int fine_mask(const optp_payoff* p, int n) {
    int m = 0;
    for (int i = 0; i < n; ++i) {
        const bool call = (p[i].flags & OPTP_F_CALL) != 0;
        switch (p[i].kind) {
            case OPTP_PAY_BARRIER: m |= (p[i].flags & OPTP_F_UP) ? M_MAX : M_MIN; break;
            case OPTP_PAY_LOOKBACK:
                if (p[i].flags & OPTP_F_FLOATING) m |= call ? M_MIN : M_MAX;
                else m |= call ? M_MAX : M_MIN;
                break;
            case OPTP_PAY_ASIAN: m |= (p[i].flags & OPTP_F_GEOMETRIC) ? M_LOG : M_SUM; break;
            default: break;
        }
    }
    return m;
}

// the instantiated state sets; a request is served by the first superset in this list
static const int kMaskSets[] = {0, M_MAX, M_MIN, M_MAX | M_MIN, M_SUM, M_LOG,
                                M_MAX | M_MIN | M_SUM | M_LOG};
static int cover_mask(int m) {
    for (int s : kMaskSets)
        if ((s & m) == m) return s;
    return M_MAX | M_MIN | M_SUM | M_LOG;
}

// ------------------------------------------------------------------ terminal (1 step / path)
// Each thread owns Philox groups of four consecutive base paths.  Sums are kept in fp32 for
// kFlush groups (<= 64 effective paths) and then folded into fp64 registers; the discount
// factor is applied once, to the totals.
struct TermConst {
    float c0, sg2, sgn, sgnK;
};

// (sum XY is not accumulated: pay * S_T = pay^2 + K pay for a call and K pay - pay^2 for a put -- both sides
// vanish where the payoff does -- so sum XY = +-sum X^2 + K sum X, formed once per thread from the fp64 totals)
__device__ __forceinline__ void term_accum(float st, const TermConst& k, float& sx, float& sxx,
                                           float& sy, float& syy) {
    const float pay = fmaxf(fmaf(k.sgn, st, k.sgnK), 0.0f);
    sx += pay;
    sxx = fmaf(pay, pay, sxx);
    sy += st;
    syy = fmaf(st, st, syy);
}

template <bool ANTI>
__global__ void __launch_bounds__(256) terminal_fp32_kernel(TermArgs a) {
    __shared__ double red[32];
    __shared__ double blockvals[8];
    const double mu = (a.r - a.q - 0.5 * a.sigma * a.sigma) * a.T;
    TermConst k;
    k.c0 = (float)((mu + log(a.S0)) * 1.4426950408889634);
    k.sg2 = (float)(a.sigma * sqrt(a.T) * 1.4426950408889634);
    k.sgn = a.is_call ? 1.0f : -1.0f;
    k.sgnK = (float)(a.is_call ? -a.K : a.K);
    const double df = exp(-a.r * a.T);
    const float bm_c = (float)(-1.3862943611198906 * (a.sigma * a.sigma * a.T) * (1.4426950408889634 * 1.4426950408889634));

    const unsigned long long first = a.path_offset, last = a.path_offset + a.n_base;  // [first,last)
    const unsigned long long g_lo = first >> 2, g_hi = (last + 3) >> 2;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    double dsx = 0, dsxx = 0, dsy = 0, dsyy = 0;
    constexpr int kFlush = 8;
    unsigned long long g = g_lo + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    while (g < g_hi) {
        float sx = 0, sxx = 0, sy = 0, syy = 0;
        uint4 w = philox_words(a.seed, g, 0u, kTermStream);
#pragma unroll 1
        for (int it = 0; it < kFlush && g < g_hi; ++it, g += stride) {
            float z[4];
            const uint4 wn = philox_words(a.seed, g + stride, 0u, kTermStream);  // software pipeline
            normals_from_words_scaled(w, bm_c, z);  // z = sigma*sqrt(T)*log2e * Z: the scale rides in the radius
            w = wn;
            const bool full = (g << 2) >= first && (g << 2) + 4 <= last;
            if (full) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float x = z[j];
                    term_accum(mufu_ex2(k.c0 + x), k, sx, sxx, sy, syy);
                    if (ANTI) term_accum(mufu_ex2(k.c0 - x), k, sx, sxx, sy, syy);
                }
            } else {
                for (int j = 0; j < 4; ++j) {
                    const unsigned long long p = (g << 2) + j;
                    if (p < first || p >= last) continue;
                    const float x = z[j];
                    term_accum(mufu_ex2(k.c0 + x), k, sx, sxx, sy, syy);
                    if (ANTI) term_accum(mufu_ex2(k.c0 - x), k, sx, sxx, sy, syy);
                }
            }
        }
        dsx += (double)sx;
        dsxx += (double)sxx;
        dsy += (double)sy;
        dsyy += (double)syy;
    }
    const double Kf = fabs((double)k.sgnK);  // the strike as the payoff above used it (rounded to fp32)
    const double dsxy = a.is_call ? dsxx + Kf * dsx : Kf * dsx - dsxx;
    double vals[5] = {dsx * df, dsxx * df * df, dsy * df, dsyy * df * df, dsxy * df * df};
    for (int i = 0; i < 5; ++i) {
        const double t = block_sum(vals[i], red);
        if (threadIdx.x == 0) blockvals[i] = t;
    }
    __syncthreads();
    double* stats = a.stats;
    const double n = (double)(ANTI ? 2 * a.n_base : a.n_base);
    finalize_partials(a.partials, a.ticket, blockvals, 5, [=](int i, double s) {
        stats[1 + i] = s;
        if (i == 0) {
            stats[0] = n;
            stats[6] = 0.0;
            stats[7] = 0.0;
        }
    });
}

int launch_terminal_fp32(const TermArgs& a, cudaStream_t st) {
    const unsigned long long groups = (a.n_base + 3) / 4 + 1;
    unsigned long long blocks = (groups + 255) / 256;
    const unsigned long long cap = (unsigned long long)ctx().n_sm * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (a.anti)
        terminal_fp32_kernel<true><<<(unsigned int)blocks, 256, 0, st>>>(a);
    else
        terminal_fp32_kernel<false><<<(unsigned int)blocks, 256, 0, st>>>(a);
    ctx().launches++;
    OPTP_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------ per-path epilogue
// Turns the running state of ONE effective path into payoffs and adds them to the thread's
// fp64 accumulators (shared memory, acc[value][thread]).
struct PathEnd {
    float ST, smax, smin, avgA, avgG;
};

__device__ __forceinline__ void path_epilogue(const PathArgs& a, const PathEnd& e, float df,
                                              unsigned long long col, unsigned long long n_eff,
                                              double* acc) {
    if (a.terminal_out) a.terminal_out[col] = (double)e.ST;
    const float Y = df * e.ST;
    acc_add(acc, 0, (double)Y);
    acc_add(acc, 1, (double)Y * (double)Y);
    for (int p = 0; p < a.n_pay; ++p) {
        const float pay = eval_payoff<float>(a.pay[p], e.ST, e.smax, e.smin, e.avgA, e.avgG);
        if (a.payoff_out) a.payoff_out[(size_t)p * n_eff + col] = (double)pay;
        const float X = df * pay;
        acc_add(acc, 3 + 3 * p, (double)X);
        acc_add(acc, 4 + 3 * p, (double)X * (double)X);
        acc_add(acc, 5 + 3 * p, (double)X * (double)Y);
    }
}

// ------------------------------------------------------------------ GBM / Heston / GBM-Milstein
// One thread = one base path (+ mirror).  State lives in log2(S/S0) for GBM and Heston.
struct LogState {  // one effective path in log2 space
    float x = 0.f, xmax = 0.f, xmin = 0.f, sum = 0.f, slog = 0.f;
};

#ifdef OPTP_POLY_EX2
// A/B experiment (task 5): 2^x on the FMA pipe instead of MUFU.EX2 -- magic-number rounding, a
// degree-5 polynomial on [-1/2, 1/2] and an exponent add: 3 FADD + 5 FFMA + SHL + IADD against one MUFU issue.
// Measured on B200 (scripts/mc_ab.py, Heston FTE + Asian): see DESIGN.md section 4 -- the kernel is bound by the
// issue rate, not by the XU pipe, so trading one MUFU for ten FMA-pipe instructions makes it slower.
__device__ __forceinline__ float poly_ex2(float x) {
    const float t = x + 12582912.0f;
    const float n = t - 12582912.0f;
    const float f = x - n;
    float p = 1.3333558e-3f;
    p = fmaf(p, f, 9.6181291e-3f);
    p = fmaf(p, f, 5.5504109e-2f);
    p = fmaf(p, f, 2.4022651e-1f);
    p = fmaf(p, f, 6.9314718e-1f);
    p = fmaf(p, f, 1.0f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
#define OPTP_OBS_EX2(x) poly_ex2(x)
#else
#define OPTP_OBS_EX2(x) mufu_ex2(x)
#endif

template <int MASK>
__device__ __forceinline__ void observe_log(LogState& s) {
    if (MASK & M_MAX) s.xmax = fmaxf(s.xmax, s.x);
    if (MASK & M_MIN) s.xmin = fminf(s.xmin, s.x);
    if (MASK & M_SUM) s.sum += OPTP_OBS_EX2(s.x);
    if (MASK & M_LOG) s.slog += s.x;
}

template <int MASK>
__device__ __forceinline__ PathEnd finish_log(const LogState& s, float S0, float inv_n) {
    PathEnd e;
    e.ST = S0 * mufu_ex2(s.x);
    e.smax = (MASK & M_MAX) ? S0 * mufu_ex2(s.xmax) : e.ST;
    e.smin = (MASK & M_MIN) ? S0 * mufu_ex2(s.xmin) : e.ST;
    e.avgA = (MASK & M_SUM) ? S0 * s.sum * inv_n : 0.f;
    e.avgG = (MASK & M_LOG) ? S0 * mufu_ex2(s.slog * inv_n) : 0.f;
    return e;
}

struct LinState {  // one effective path in linear space (Milstein schemes)
    float S, smax, smin, sum = 0.f, slog = 0.f;
    // local-vol kernels: extrema of the table position of this path's lookups (two FMNMX per step instead
    // of a compare + predicated add; the path counts as clamped if it ever left [0, pos_max])
    float pmin = 0.f, pmax = 0.f;
};

template <int MASK>
__device__ __forceinline__ void observe_lin(LinState& s) {
    if (MASK & M_MAX) s.smax = fmaxf(s.smax, s.S);
    if (MASK & M_MIN) s.smin = fminf(s.smin, s.S);
    if (MASK & M_SUM) s.sum += s.S;
    if (MASK & M_LOG) s.slog += mufu_lg2(s.S);
}

template <int MASK>
__device__ __forceinline__ PathEnd finish_lin(const LinState& s, float inv_n) {
    PathEnd e;
    e.ST = s.S;
    e.smax = (MASK & M_MAX) ? s.smax : s.S;
    e.smin = (MASK & M_MIN) ? s.smin : s.S;
    e.avgA = (MASK & M_SUM) ? s.sum * inv_n : 0.f;
    e.avgG = (MASK & M_LOG) ? mufu_ex2(s.slog * inv_n) : 0.f;
    return e;
}

struct StepConst {
    // GBM (log2 units)
    float vol2, drift2;
    // Heston (x in log2 units, v in natural units)
    float rho, rhoc, h_sdt2, h_mhalf2, h_rq2, h_c1, h_kt, h_xs;
    // Heston FTE with the normals pre-scaled by s = |xi| sqrt(dt) inside Box-Muller (h_bm = -2 ln2 s^2):
    // zc' = h_A*z0s + h_B*z1s = sqrt(dt)*log2e*(rho Z0 + rhoc Z1)
    float h_bm, h_A, h_B;
    // GBM Milstein (linear)
    float m_c1, m_ssdt, m_hs2dt;
    // Heston QE (Andersen 2008); x in log2 units
    float q_theta, q_E, q_c1, q_c2, q_K0, q_K1, q_K2, q_K3;
    float q_A, q_hK3, q_rq2;  // martingale correction: A = K2 + K3/2 (natural units), K3/2 and (r-q) dt in log2 units
    // Merton jumps (log2 units): drift incl. compensator, mJ, sJ, mu = lam*dt, exp(-mu)
    float j_drift2, j_m2, j_s2, j_mu, j_p0;
    // SABR: log2 sigma += s_nu2*Z2 + s_vdr2; beta = 1: log2 S += sigma*zc*s_sdt2 + sigma^2*s_mhdt2 + s_rq2;
    // beta < 1: S += S*s_rqdt + sigma*S^beta*s_sdt*zc
    float s_nu2, s_vdr2, s_beta, s_sdt2, s_mhdt2, s_rq2, s_rqdt, s_sdt;
};

// one time step of a base path and its mirror, per process
template <int MASK, bool ANTI>
__device__ __forceinline__ void gbm_step(LogState& p, LogState& m, float z, const StepConst& k) {
    // log2 S accumulates drift + vol*Z exactly like the reference's cumsum (processes.py:52-53)
    p.x += fmaf(k.vol2, z, k.drift2);
    observe_log<MASK>(p);
    if (ANTI) {
        m.x += fmaf(-k.vol2, z, k.drift2);
        observe_log<MASK>(m);
    }
}

template <int MASK, bool ANTI>
__device__ __forceinline__ void heston_step(LogState& p, LogState& m, float& vp, float& vm, float z0,
                                            float z1, const StepConst& k) {
    // full-truncation Euler (processes.py:150-160); v >= 0 is re-established at the end of every
    // step, so max(v,0) at the start of the next one is the identity.
    // z0, z1 arrive scaled by s = |xi| sqrt(dt) (free: folded into the Box-Muller radius), so the variance
    // update is fma(sqrt(v), z0, ...) and the correlated log-price shock is fma(sqrt(v), zc, ...) with
    // zc = (rho Z0 + rhoc Z1) sqrt(dt) log2e: 7 FP32 instructions per path-step instead of 9.
    const float zc = fmaf(k.h_A, z0, k.h_B * z1);
    {
        const float sq = mufu_sqrt(vp);
        p.x = fmaf(sq, zc, fmaf(vp, k.h_mhalf2, p.x + k.h_rq2));
        vp = fmaxf(fmaf(sq, z0, fmaf(vp, k.h_c1, k.h_kt)), 0.f);
        observe_log<MASK>(p);
    }
    if (ANTI) {
        const float sq = mufu_sqrt(vm);
        m.x = fmaf(-sq, zc, fmaf(vm, k.h_mhalf2, m.x + k.h_rq2));
        vm = fmaxf(fmaf(-sq, z0, fmaf(vm, k.h_c1, k.h_kt)), 0.f);
        observe_log<MASK>(m);
    }
}

// Heston quadratic-exponential step of ONE path (no reference code: Andersen 2008, see
// oracle/path_schemes.py:heston_qe).  psi <= 1.5  <=>  tt = 2 m^2 / s^2 >= 4/3.
template <int MASK, bool MART>
__device__ __forceinline__ void heston_qe_one(LogState& s, float& v, float z0, float z1,
                                              const StepConst& k) {
    const float m = fmaf(v - k.q_theta, k.q_E, k.q_theta);
    const float s2 = fmaf(v, k.q_c1, k.q_c2);
    float inv_s2;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv_s2) : "f"(s2));
    const float m2 = m * m;
    const float tt = (m2 + m2) * inv_s2;
    float vn;
    // MART: log2 of M = E[exp(A v') | v] (Andersen 2008, eqs. 4.14-4.16); NaN where its regularity condition
    // fails, which keeps the uncorrected drift below
    float lg2M = 0.0f;
    if (tt >= 1.3333334f) {
        const float r1 = mufu_sqrt(fmaf(tt, tt, -tt));
        const float b2 = (tt - 1.0f) + r1;
        float inv_1b2;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv_1b2) : "f"(1.0f + b2));
        const float bz = mufu_sqrt(b2) + z0;
        const float aq = m * inv_1b2;
        vn = aq * (bz * bz);
        if (MART) {  // M = exp(A b^2 a / (1 - 2 A a)) / sqrt(1 - 2 A a)
            const float den = fmaf(-2.0f * k.q_A, aq, 1.0f);
            float inv_den;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv_den) : "f"(den));
            lg2M = fmaf(1.4426950408889634f * k.q_A * b2 * aq, inv_den, -0.5f * mufu_lg2(den));  // den <= 0 -> NaN
        }
    } else {  // exponential branch (small variance, psi = 2/tt > 1.5): v' = 0 with probability p, else exponential
        // p = (psi - 1)/(psi + 1) = (2 - tt)/(2 + tt), 1 - p = 2 tt/(2 + tt), v' = ln((1 - p)/(1 - u)) m/(1 - p)
        // with u = Phi(z0); u <= p  <=>  1 - u >= 1 - p.  Reciprocals and the log are single MUFU ops here
        // (a divergent warp pays for this branch in full, so it must be short); one normcdff, not two.
        float inv_2ptt, inv_tt;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv_2ptt) : "f"(2.0f + tt));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv_tt) : "f"(tt));
        const float omp = (tt + tt) * inv_2ptt;               // 1 - p
        const float omu = normcdff(-z0);                      // 1 - u
        float inv_omu;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv_omu) : "f"(omu));
        const float scale = 0.5f * m * (2.0f + tt) * inv_tt;  // m / (1 - p)
        vn = (omu >= omp) ? 0.0f : (0.6931471805599453f * mufu_lg2(omp * inv_omu)) * scale;
        if (MART) {  // M = p + beta (1 - p)/(beta - A), beta = (1 - p)/m  ->  p + (1-p)^2 / ((1-p) - A m)
            const float den = fmaf(-k.q_A, m, omp);
            float inv_den;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv_den) : "f"(den));
            lg2M = (den > 0.0f) ? mufu_lg2(fmaf(omp * omp, inv_den, 1.0f - omp)) : __int_as_float(0x7fc00000);
        }
    }
    const float vol = mufu_sqrt(fmaxf(k.q_K3 * (v + vn), 0.0f));
    float drift = fmaf(k.q_K1, v, s.x + k.q_K0);
    if (MART) {
        const float corrected = fmaf(-k.q_hK3, v, (s.x + k.q_rq2) - lg2M);
        drift = (lg2M == lg2M) ? corrected : drift;
    }
    s.x = fmaf(vol, z1, fmaf(k.q_K2, vn, drift));
    v = vn;
    observe_log<MASK>(s);
}

template <int MASK, bool ANTI>
__device__ __forceinline__ void gbmm_step(LinState& p, LinState& m, float z, const StepConst& k) {
    // S' = S + S*g with g kept as a SMALL number (processes.py:296-303): folding the 1 into a
    // rounded fp32 constant would bias every step by up to 6e-8
    const float zs = k.m_ssdt * z;
    const float zq = fmaf(k.m_hs2dt * z, z, k.m_c1);
    p.S = fmaxf(fmaf(p.S, zq + zs, p.S), 1e-10f);
    observe_lin<MASK>(p);
    if (ANTI) {
        m.S = fmaxf(fmaf(m.S, zq - zs, m.S), 1e-10f);
        observe_lin<MASK>(m);
    }
}

// SABR with lognormal volatility (processes.py:199-215): the volatility is evolved exactly in log2 space
// (one FFMA + one MUFU.EX2), then the price takes a log-Euler step (beta = 1) or an Euler step with the
// reference's 1e-12 floor (beta < 1, S^beta = ex2(beta*lg2 S)).  The mirror negates both normals.
template <int MASK, bool ANTI>
__device__ __forceinline__ void sabr_step_log(LogState& p, LogState& m, float& lp, float& lm, float z0, float z1,
                                              const StepConst& k) {
    const float zc = fmaf(k.rho, z0, k.rhoc * z1);
    {
        lp += fmaf(k.s_nu2, z0, k.s_vdr2);
        const float sg = mufu_ex2(lp);
        p.x += fmaf(sg * zc, k.s_sdt2, fmaf(sg * sg, k.s_mhdt2, k.s_rq2));
        observe_log<MASK>(p);
    }
    if (ANTI) {
        lm += fmaf(-k.s_nu2, z0, k.s_vdr2);
        const float sg = mufu_ex2(lm);
        m.x += fmaf(-(sg * zc), k.s_sdt2, fmaf(sg * sg, k.s_mhdt2, k.s_rq2));
        observe_log<MASK>(m);
    }
}

template <int MASK, bool ANTI>
__device__ __forceinline__ void sabr_step_lin(LinState& p, LinState& m, float& lp, float& lm, float z0, float z1,
                                              const StepConst& k) {
    const float zc = k.s_sdt * fmaf(k.rho, z0, k.rhoc * z1);
    {
        lp += fmaf(k.s_nu2, z0, k.s_vdr2);
        const float sb = mufu_ex2(fmaf(k.s_beta, mufu_lg2(p.S), lp));  // sigma * S^beta
        p.S = fmaxf(fmaf(sb, zc, fmaf(p.S, k.s_rqdt, p.S)), 1e-12f);
        observe_lin<MASK>(p);
    }
    if (ANTI) {
        lm += fmaf(-k.s_nu2, z0, k.s_vdr2);
        const float sb = mufu_ex2(fmaf(k.s_beta, mufu_lg2(m.S), lm));
        m.S = fmaxf(fmaf(-sb, zc, fmaf(m.S, k.s_rqdt, m.S)), 1e-12f);
        observe_lin<MASK>(m);
    }
}

template <int PROC, int MASK, bool ANTI, int RNG = RNG_PHILOX10>
__global__ void __launch_bounds__(256, (PROC == OPTP_PROC_HESTON_FTE || PROC == OPTP_PROC_HESTON_QE) ? 4 : 5) path_fp32_kernel(PathArgs a) {
    extern __shared__ double dyn[];
    __shared__ double blockvals[kMaxVals];
    double* acc = dyn;
    acc_zero(acc, 3 + 3 * a.n_pay);

    const unsigned int n_steps = a.n_steps;
    const double dt = a.T / (double)n_steps;
    const double sdt = sqrt(dt);
    constexpr double L2E = 1.4426950408889634;
    StepConst k;
    k.vol2 = (float)(a.sigma * sdt * L2E);
    k.drift2 = (float)((a.r - a.q - 0.5 * a.sigma * a.sigma) * dt * L2E);
    k.rho = (float)a.rho;
    k.rhoc = (float)sqrt(fmax(0.0, 1.0 - a.rho * a.rho));
    k.h_sdt2 = (float)(sdt * L2E);
    k.h_mhalf2 = (float)(-0.5 * dt * L2E);
    k.h_rq2 = (float)((a.r - a.q) * dt * L2E);
    k.h_c1 = (float)(1.0 - a.kappa * dt);
    k.h_kt = (float)(a.kappa * a.theta * dt);
    k.h_xs = (float)(a.xi * sdt);
    {
        // xi < 0 is the same process as xi > 0 with rho -> -rho once Z0 is relabelled -Z0 (statistically
        // identical; not path-identical to the fp64 kernel for xi < 0); xi = 0 keeps a representable scale
        // (the variance shock then is 1e-15 * Z0)
        const double s = fmax(fabs(a.xi) * sdt, 1e-15);
        const double rho_eff = (a.xi < 0.0) ? -a.rho : a.rho;
        k.h_bm = (float)(-1.3862943611198906 * s * s);
        k.h_A = (float)(rho_eff * sdt * L2E / s);
        k.h_B = (float)(sqrt(fmax(0.0, 1.0 - a.rho * a.rho)) * sdt * L2E / s);
    }
    if (PROC == OPTP_PROC_HESTON_QE) {
        const double E = exp(-a.kappa * dt);
        k.q_theta = (float)a.theta;
        k.q_E = (float)E;
        k.q_c1 = (float)(a.xi * a.xi * E * (1.0 - E) / a.kappa);
        k.q_c2 = (float)(a.theta * a.xi * a.xi * (1.0 - E) * (1.0 - E) / (2.0 * a.kappa));
        k.q_K0 = (float)(((a.r - a.q) * dt - a.rho * a.kappa * a.theta * dt / a.xi) * L2E);
        k.q_K1 = (float)((0.5 * dt * (a.kappa * a.rho / a.xi - 0.5) - a.rho / a.xi) * L2E);
        k.q_K2 = (float)((0.5 * dt * (a.kappa * a.rho / a.xi - 0.5) + a.rho / a.xi) * L2E);
        k.q_K3 = (float)(0.5 * dt * (1.0 - a.rho * a.rho) * L2E * L2E);
        k.q_A = (float)((0.5 * dt * (a.kappa * a.rho / a.xi - 0.5) + a.rho / a.xi) + 0.25 * dt * (1.0 - a.rho * a.rho));
        k.q_hK3 = (float)(0.25 * dt * (1.0 - a.rho * a.rho) * L2E);
        k.q_rq2 = (float)((a.r - a.q) * dt * L2E);
    }
    if (PROC == OPTP_PROC_MERTON) {
        const double kj = exp(a.kappa + 0.5 * a.theta * a.theta) - 1.0;
        k.j_drift2 = (float)((a.r - a.q - 0.5 * a.sigma * a.sigma - a.v0 * kj) * dt * L2E);
        k.j_m2 = (float)(a.kappa * L2E);
        k.j_s2 = (float)(a.theta * L2E);
        k.j_mu = (float)(a.v0 * dt);
        k.j_p0 = (float)exp(-a.v0 * dt);
    }
    if (PROC == OPTP_PROC_SABR) {  // v0 = alpha0, kappa = beta, xi = nu (mc_fp64.cu, processes.py:199-215)
        k.s_nu2 = (float)(a.xi * sdt * L2E);
        k.s_vdr2 = (float)(-0.5 * a.xi * a.xi * dt * L2E);
        k.s_beta = (float)a.kappa;
        k.s_sdt2 = (float)(sdt * L2E);
        k.s_mhdt2 = (float)(-0.5 * dt * L2E);
        k.s_rq2 = (float)((a.r - a.q) * dt * L2E);
        k.s_rqdt = (float)((a.r - a.q) * dt);
        k.s_sdt = (float)sdt;
    }
    k.m_c1 = (float)((a.r - a.q) * dt - 0.5 * a.sigma * a.sigma * dt);
    k.m_ssdt = (float)(a.sigma * sdt);
    k.m_hs2dt = (float)(0.5 * a.sigma * a.sigma * dt);
    const float S0 = (float)a.S0;
    const float inv_n = (float)(1.0 / (double)n_steps);
    const float df = (float)exp(-a.pay[0].r_disc * a.pay[0].T_disc);
    const unsigned long long n_eff = ANTI ? 2 * a.n_base : a.n_base;
    constexpr int kDims =
        (PROC == OPTP_PROC_HESTON_FTE || PROC == OPTP_PROC_HESTON_QE || PROC == OPTP_PROC_SABR) ? 2 : 1;
    constexpr int kPerCall = RngTraits<RNG>::normals / kDims;  // steps served by one Philox call
    const unsigned int n_full = n_steps / kPerCall, n_tail = n_steps % kPerCall;

    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
         b < a.n_base; b += stride) {
        const unsigned long long gpath = a.path_offset + b;
        PathEnd ep, em;
        float z[RngTraits<RNG>::normals];
        if (PROC == OPTP_PROC_GBM) {
            // (RNG = 0 is bit-identical to normals_from_words: same words, same radius constant)
            LogState p, m;
            uint4 w = philox_words_m<RNG>(a.seed, gpath, 0u, 0u);
#pragma unroll 1
            for (unsigned int c = 0; c < n_full; ++c) {
                const uint4 wn = philox_words_m<RNG>(a.seed, gpath, c + 1, 0u);  // software pipeline
                normals_from_words_m<RNG>(w, -1.3862943611198906f, z);
#pragma unroll
                for (int j = 0; j < kPerCall; ++j) gbm_step<MASK, ANTI>(p, m, z[j], k);
                w = wn;
            }
            if (n_tail) {
                normals_from_words_m<RNG>(w, -1.3862943611198906f, z);
                for (unsigned int j = 0; j < n_tail; ++j) gbm_step<MASK, ANTI>(p, m, z[j], k);
            }
            ep = finish_log<MASK>(p, S0, inv_n);
            if (ANTI) em = finish_log<MASK>(m, S0, inv_n);
        } else if (PROC == OPTP_PROC_HESTON_FTE) {
            LogState p, m;
            float vp = fmaxf((float)a.v0, 0.f), vm = vp;
            uint4 w = philox_words_m<RNG>(a.seed, gpath, 0u, 0u);
#pragma unroll 1
            for (unsigned int c = 0; c < n_full; ++c) {
                const uint4 wn = philox_words_m<RNG>(a.seed, gpath, c + 1, 0u);
                normals_from_words_m<RNG>(w, k.h_bm, z);  // (RNG = 0: the words and normals of normals_from_words_scaled)
#pragma unroll
                for (int j = 0; j < kPerCall; ++j) heston_step<MASK, ANTI>(p, m, vp, vm, z[2 * j], z[2 * j + 1], k);
                w = wn;
            }
            if (n_tail) {
                normals_from_words_m<RNG>(w, k.h_bm, z);
                for (unsigned int j = 0; j < n_tail; ++j) heston_step<MASK, ANTI>(p, m, vp, vm, z[2 * j], z[2 * j + 1], k);
            }
            ep = finish_log<MASK>(p, S0, inv_n);
            if (ANTI) em = finish_log<MASK>(m, S0, inv_n);
        } else if (PROC == OPTP_PROC_MERTON) {
            // one Philox call per step: (Z, ZJ) from one Box-Muller pair, the Poisson count from the
            // third word; the mirror negates Z and ZJ and shares the count (processes.py:95-99)
            LogState p, m;
            uint4 w = philox_words(a.seed, gpath, 0u, 0u);
#pragma unroll 1
            for (unsigned int t = 0; t < n_steps; ++t) {
                const uint4 wn = philox_words(a.seed, gpath, t + 1, 0u);
                float z0, zj;
                box_muller(w.x, w.y, z0, zj);
                const float kj = poisson_inv(w.z, k.j_mu, k.j_p0);
                const float dif = k.vol2 * z0;
                float jm = 0.f, js = 0.f;
                if (kj > 0.f) {
                    jm = k.j_m2 * kj;
                    js = k.j_s2 * mufu_sqrt(kj) * zj;
                }
                p.x += (k.j_drift2 + jm) + (dif + js);
                observe_log<MASK>(p);
                if (ANTI) {
                    m.x += (k.j_drift2 + jm) - (dif + js);
                    observe_log<MASK>(m);
                }
                w = wn;
            }
            ep = finish_log<MASK>(p, S0, inv_n);
            if (ANTI) em = finish_log<MASK>(m, S0, inv_n);
        } else if (PROC == OPTP_PROC_HESTON_QE) {
            LogState p, m;
            float vp = fmaxf((float)a.v0, 0.f), vm = vp;
            uint4 w = philox_words(a.seed, gpath, 0u, 0u);
            auto qe_steps = [&](auto mart) {
                constexpr bool MART = decltype(mart)::value;
#pragma unroll 1
                for (unsigned int c = 0; c < n_full; ++c) {
                    const uint4 wn = philox_words(a.seed, gpath, c + 1, 0u);
                    normals_from_words(w, z);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        heston_qe_one<MASK, MART>(p, vp, z[2 * j], z[2 * j + 1], k);
                        if (ANTI) heston_qe_one<MASK, MART>(m, vm, -z[2 * j], -z[2 * j + 1], k);
                    }
                    w = wn;
                }
                if (n_tail) {
                    normals_from_words(w, z);
                    heston_qe_one<MASK, MART>(p, vp, z[0], z[1], k);
                    if (ANTI) heston_qe_one<MASK, MART>(m, vm, -z[0], -z[1], k);
                }
            };
            if (a.qe_martingale) qe_steps(std::true_type{});
            else qe_steps(std::false_type{});
            ep = finish_log<MASK>(p, S0, inv_n);
            if (ANTI) em = finish_log<MASK>(m, S0, inv_n);
        } else if (PROC == OPTP_PROC_SABR) {
            float lp = mufu_lg2((float)a.v0), lm = lp;  // log2 alpha0
            uint4 w = philox_words(a.seed, gpath, 0u, 0u);
            if (a.kappa == 1.0) {
                LogState p, m;
#pragma unroll 1
                for (unsigned int c = 0; c < n_full; ++c) {
                    const uint4 wn = philox_words(a.seed, gpath, c + 1, 0u);
                    normals_from_words(w, z);
                    sabr_step_log<MASK, ANTI>(p, m, lp, lm, z[0], z[1], k);
                    sabr_step_log<MASK, ANTI>(p, m, lp, lm, z[2], z[3], k);
                    w = wn;
                }
                if (n_tail) {
                    normals_from_words(w, z);
                    sabr_step_log<MASK, ANTI>(p, m, lp, lm, z[0], z[1], k);
                }
                ep = finish_log<MASK>(p, S0, inv_n);
                if (ANTI) em = finish_log<MASK>(m, S0, inv_n);
            } else {
                LinState p, m;
                p.S = p.smax = p.smin = S0;
                m = p;
#pragma unroll 1
                for (unsigned int c = 0; c < n_full; ++c) {
                    const uint4 wn = philox_words(a.seed, gpath, c + 1, 0u);
                    normals_from_words(w, z);
                    sabr_step_lin<MASK, ANTI>(p, m, lp, lm, z[0], z[1], k);
                    sabr_step_lin<MASK, ANTI>(p, m, lp, lm, z[2], z[3], k);
                    w = wn;
                }
                if (n_tail) {
                    normals_from_words(w, z);
                    sabr_step_lin<MASK, ANTI>(p, m, lp, lm, z[0], z[1], k);
                }
                ep = finish_lin<MASK>(p, inv_n);
                if (ANTI) em = finish_lin<MASK>(m, inv_n);
            }
        } else {  // OPTP_PROC_GBM_MILSTEIN
            LinState p, m;
            p.S = p.smax = p.smin = S0;
            m = p;
            uint4 w = philox_words(a.seed, gpath, 0u, 0u);
#pragma unroll 1
            for (unsigned int c = 0; c < n_full; ++c) {
                const uint4 wn = philox_words(a.seed, gpath, c + 1, 0u);
                normals_from_words(w, z);
#pragma unroll
                for (int j = 0; j < 4; ++j) gbmm_step<MASK, ANTI>(p, m, z[j], k);
                w = wn;
            }
            if (n_tail) {
                normals_from_words(w, z);
                for (unsigned int j = 0; j < n_tail; ++j) gbmm_step<MASK, ANTI>(p, m, z[j], k);
            }
            ep = finish_lin<MASK>(p, inv_n);
            if (ANTI) em = finish_lin<MASK>(m, inv_n);
        }
        path_epilogue(a, ep, df, b, n_eff, acc);
        if (ANTI) path_epilogue(a, em, df, a.n_base + b, n_eff, acc);
    }
    reduce_and_emit(acc, blockvals, a.n_pay, (double)n_eff, a.partials, a.ticket, a.stats);
}

// ------------------------------------------------------------------ local volatility
// sigma(S,t) table: row i (step time i*dt) holds n_nodes (value, next-value minus value) pairs
// on a uniform ln S grid.  The CTA walks the steps in lock step: a chunk of rows is staged in
// shared memory, every thread advances its paths through the chunk, the next chunk is staged.
// Lookup: pos = (log2 S - x0_2)/h_2 clamped to [0, n-1]; the cell index comes from the
// round-to-nearest "magic number" add (no F2I on the XU pipe), then one LDS.64 and one FFMA.
struct LvConst {
    float inv_h2, c0, pos_max, c_up, c_dn, k_up, k_dn, c1, sdt, hdt, bm_dt;
    float e_sdt2, e_mhalf2, e_rq2;  // log-Euler variant
};

__device__ __forceinline__ float table_sigma32(const float2* row, float pos, float pos_max) {
    const float pc = fminf(fmaxf(pos, 0.f), pos_max);
    const float t = (pc - 0.5f) + 12582912.0f;
    const int j = __float_as_int(t) - 0x4B400000;
    const float w = pc - (t - 12582912.0f);
    const float2 e = row[j];
    return fmaf(w, e.y, e.x);
}

template <int PROC, int MASK>
__device__ __forceinline__ void lv_step(LinState& s, const float2* row, const LvConst& k, float zs,
                                        float zzh) {
    const float pos = fmaf(mufu_lg2(s.S), k.inv_h2, k.c0);
    s.pmin = fminf(s.pmin, pos);
    s.pmax = fmaxf(s.pmax, pos);
    float sig = table_sigma32(row, pos, k.pos_max);
    if (PROC == OPTP_PROC_LV_MILSTEIN) {
        // processes.py:361-379.  s_up = S(1+b), s_dn = S(1-b): the bumped lookups are constant
        // offsets in pos, and S cancels from d(sigma*S)/dS.
        sig = fminf(fmaxf(sig, 1e-8f), 10.0f);
        const float sig_up = table_sigma32(row, pos + k.c_up, k.pos_max);
        const float sig_dn = table_sigma32(row, pos + k.c_dn, k.pos_max);
        const float da = fmaf(sig_up, k.k_up, -(sig_dn * k.k_dn));
        const float av = sig * s.S;
        s.S = fmaxf(fmaf(av, fmaf(da, zzh, zs), fmaf(s.S, k.c1, s.S)), 1e-10f);
    } else {  // LV_EULER, processes.py:253-256
        sig = fmaxf(sig, 0.f);
        const float ex = fmaf(sig * zs, 1.4426950408889634f, fmaf(sig * sig, k.e_mhalf2, k.e_rq2));  // zs = sqrt(dt) Z
        s.S = s.S * mufu_ex2(ex);
    }
    observe_lin<MASK>(s);
}

// Milstein step with the "window" table: node j holds (v_j, s_j, s_{j+1}, s_{j-1}) with
// s_j = v_{j+1} - v_j, so ONE LDS.128 serves sigma(S), sigma(S(1+b)) and sigma(S(1-b)) as long as
// the bump is smaller than a cell (|c_up|, |c_dn| < 1): the bumped points are in the same cell or in
// the neighbouring one.  Rows carry two replicated nodes on each side (flat extrapolation), pos is
// clamped to [-1, n]; that reproduces the per-lookup clamping of the (value, slope) path exactly.
template <int MASK, bool LDG = false, bool CLIP = true>
__device__ __forceinline__ void lvm_step_win(LinState& s, const float4* row, const LvConst& k, float zs,
                                             float zzh) {
    // Issue-cost model of sm_100a (profiles/r1_pipe_model.md): FFMA/FADD/FMNMX cost one issue
    // cycle, FSETP/IADD/LOP3/LEA two.  The piecewise-linear evaluations are therefore written with
    // min/max instead of compare + select, and the clamp counter costs one compare.
    const float pos = fmaf(mufu_lg2(s.S), k.inv_h2, k.c0);
    s.pmin = fminf(s.pmin, pos);                          // lookups outside the tabulated range are counted
    s.pmax = fmaxf(s.pmax, pos);                          // per path, from these extrema
    const float pc = fminf(fmaxf(pos, -1.0f), k.pos_max + 1.0f);
    const float t = __fadd_rd(pc, 12582912.0f);           // floor(pc) + 1.5*2^23, one FADD.RM
    const int j = __float_as_int(t) - 0x4B400000;         // floor(pc) in [-1, n]
    const float f = pc - (t - 12582912.0f);               // fraction in [0, 1)
    const float4 e = LDG ? __ldg(row + j + 2) : row[j + 2];
    float sig = fmaf(f, e.y, e.x);
    // sigma at x + c_up (c_up in (0,1)):  v_j + min(tu,1) s_j + max(tu-1,0) s_{j+1}
    // (max(tu-1, 0) = tu - min(tu, 1) and min(td, 0) = td - max(td, 0), bit for bit: one FMNMX + one FADD per
    // bumped lookup instead of two FMNMX + one FADD)
    const float tu = f + k.c_up;
    const float tu1 = fminf(tu, 1.0f);
    const float sig_up = fmaf(tu - tu1, e.z, fmaf(tu1, e.y, e.x));
    // sigma at x + c_dn (c_dn in (-1,0)): v_j + max(td,0) s_j + min(td,0) s_{j-1}
    const float td = f + k.c_dn;
    const float td0 = fmaxf(td, 0.0f);
    const float sig_dn = fmaf(td - td0, e.w, fmaf(td0, e.y, e.x));
    // processes.py:361-362 clips sigma to [1e-8, 10]; a table whose nodes all lie inside that range (every
    // Dupire table: calibration.py:355 clips to [0.01, 5]) cannot leave it by linear interpolation
    if (CLIP) sig = fminf(fmaxf(sig, 1e-8f), 10.0f);
    const float da = fmaf(sig_up, k.k_up, -(sig_dn * k.k_dn));
    const float av = sig * s.S;
    s.S = fmaxf(fmaf(av, fmaf(da, zzh, zs), fmaf(s.S, k.c1, s.S)), 1e-10f);
    observe_lin<MASK>(s);
}

constexpr int kLvThreads = 256;
#ifndef OPTP_LV_TMA
#define OPTP_LV_TMA 1
#endif

// ---- TMA bulk copy (cp.async.bulk -> SASS UBLKCP) + mbarrier helpers --------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// global -> shared bulk copy executed by the TMA unit; completion is signalled on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (!done && ++spins > (1u << 24)) __trap();  // never hang the device on a protocol bug
    }
}

// WINMODE: 0 = (value, slope) table, three lookups; 1 = window table, one load; 2 = window table whose nodes
// all lie inside the reference's clip range [1e-8, 10] (no per-step clip)
template <int PROC, int MASK, bool ANTI, int WINMODE>
__global__ void __launch_bounds__(kLvThreads, 4) lv_fp32_kernel(PathArgs a, int rows_per_chunk) {
    constexpr bool WIN = WINMODE != 0;
    using Node = typename std::conditional<WIN, float4, float2>::type;
    extern __shared__ double dyn[];
    __shared__ double blockvals[kMaxVals];
    const int n_vals = 3 + 3 * a.n_pay;
    double* acc = dyn;
    Node* rows = reinterpret_cast<Node*>(dyn + (size_t)n_vals * kLvThreads);
    const Node* table = WIN ? reinterpret_cast<const Node*>(a.table4) : reinterpret_cast<const Node*>(a.table32);
    acc_zero(acc, n_vals);

    const unsigned int n_steps = a.n_steps;
    const double dt = a.T / (double)n_steps;
    const double sdt = sqrt(dt);
    constexpr double L2E = 1.4426950408889634, LN2 = 0.6931471805599453;
    LvConst k;
    k.inv_h2 = (float)(LN2 / a.h);              // d pos / d log2 S
    k.c0 = (float)(-a.x0 / a.h);
    k.pos_max = (float)(a.n_nodes - 1);
    k.c_up = (float)(log1p(a.bump) / a.h);
    k.c_dn = (float)(log1p(-fmin(a.bump, 0.999999)) / a.h);
    k.k_up = (float)((1.0 + a.bump) / (2.0 * a.bump));
    k.k_dn = (float)((1.0 - a.bump) / (2.0 * a.bump));
    k.c1 = (float)((a.r - a.q) * dt);  // S*(1+c1) is evaluated as fma(S, c1, S)
    k.sdt = (float)sdt;
    k.hdt = (float)(0.5 * dt);
    k.bm_dt = (float)(-1.3862943611198906 * dt);
    k.e_sdt2 = (float)(sdt * L2E);
    k.e_mhalf2 = (float)(-0.5 * dt * L2E);
    k.e_rq2 = (float)((a.r - a.q) * dt * L2E);
    const float S0 = (float)a.S0;
    const float inv_n = (float)(1.0 / (double)n_steps);
    const float df = (float)exp(-a.pay[0].r_disc * a.pay[0].T_disc);
    const unsigned long long n_eff = ANTI ? 2 * a.n_base : a.n_base;
    // row length: (value, slope) table padded to an even node count; window table n + 4 nodes
    const int n_ld = WIN ? a.n_nodes + 4 : ((a.n_nodes + 1) & ~1);

    // Staging of the table rows, selectable at compile time (OPTP_LV_TMA, default 1):
    //   OPTP_LV_TMA=1 (default): the TMA unit (cp.async.bulk + mbarrier) fills one of two ~36 KB buffers while
    //     the CTA walks the other; a buffer is re-armed by thread 0 after the __syncthreads that ends its chunk.
    //     Chunks are numbered globally (across slabs) so the mbarrier phase of buffer b is (use count of b) & 1.
    //     Measured on B200, config 3: 4.9e11 path-steps/s against 4.5e11 for the copy loop below (no LDG -> STS
    //     instructions in the issue stream, one barrier per chunk instead of two).
    //   OPTP_LV_TMA=0: the CTA copies one chunk itself between two __syncthreads (kept for A/B runs).
    __shared__ __align__(8) uint64_t full_bar[2];
    const unsigned int n_chunks = (n_steps + rows_per_chunk - 1) / rows_per_chunk;
    const unsigned int row_bytes = (unsigned int)n_ld * (unsigned int)sizeof(Node);
    const unsigned int buf_elems = (unsigned int)rows_per_chunk * (unsigned int)n_ld;
    const unsigned long long n_slabs = (a.n_base + kLvThreads - 1) / kLvThreads;
    const unsigned long long my_slabs =
        blockIdx.x < n_slabs ? (n_slabs - blockIdx.x + gridDim.x - 1) / gridDim.x : 0ull;
    const unsigned long long total_chunks = my_slabs * n_chunks;
    auto issue = [&](unsigned long long gc) {  // thread 0: arm buffer gc&1 with chunk (gc mod n_chunks)
        const unsigned int c = (unsigned int)(gc % n_chunks);
        const unsigned int t0 = c * rows_per_chunk;
        const unsigned int nr = min((unsigned int)rows_per_chunk, n_steps - t0);
        uint64_t* bar = &full_bar[gc & 1];
        mbar_expect_tx(bar, nr * row_bytes);
        bulk_g2s(rows + (size_t)(gc & 1) * buf_elems, table + (size_t)t0 * n_ld, nr * row_bytes, bar);
    };
    if (OPTP_LV_TMA) {
        if (threadIdx.x == 0) {
            mbar_init(&full_bar[0], 1);
            mbar_init(&full_bar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            if (total_chunks > 0) issue(0);
            if (total_chunks > 1) issue(1);
        }
    }
    unsigned long long g = 0;  // global chunk counter of this CTA
    for (unsigned long long slab = blockIdx.x; slab < n_slabs; slab += gridDim.x) {
        const unsigned long long b = slab * kLvThreads + threadIdx.x;
        const bool active = b < a.n_base;
        const unsigned long long gpath = a.path_offset + b;
        LinState p, m;
        p.S = p.smax = p.smin = S0;
        m = p;
        p.pmin = p.pmax = m.pmin = m.pmax = 0.5f * k.pos_max;
        // the Philox words of the next 4-step block are always requested one block ahead, across
        // chunk boundaries too (chunks are multiples of 4 rows)
        uint4 w = philox_words(a.seed, gpath, 0u, 0u);
        for (unsigned int t0 = 0; t0 < n_steps; t0 += rows_per_chunk, ++g) {
            const unsigned int nr = min((unsigned int)rows_per_chunk, n_steps - t0);
            const Node* buf = rows + (OPTP_LV_TMA ? (size_t)(g & 1) * buf_elems : 0);
            if (OPTP_LV_TMA) {
                mbar_wait(&full_bar[g & 1], (uint32_t)((g >> 1) & 1));
            } else {
                __syncthreads();  // everyone is done with the previous chunk
                const float4* src = reinterpret_cast<const float4*>(table + (size_t)t0 * n_ld);
                float4* dst = reinterpret_cast<float4*>(rows);
                const unsigned int n4 = nr * row_bytes / 16u;
                for (unsigned int i = threadIdx.x; i < n4; i += kLvThreads) dst[i] = __ldg(src + i);
                __syncthreads();
            }
            if (active) {
                // one step of the pair on table row `row`
                auto step = [&](const Node* row, float zs) {
                    const float zzh = fmaf(0.5f * zs, zs, -k.hdt);  // (Z^2 - 1) dt / 2
                    if constexpr (WIN) {
                        lvm_step_win<MASK, false, WINMODE != 2>(p, row, k, zs, zzh);
                        if (ANTI) lvm_step_win<MASK, false, WINMODE != 2>(m, row, k, -zs, zzh);
                    } else {
                        lv_step<PROC, MASK>(p, row, k, zs, zzh);
                        if (ANTI) lv_step<PROC, MASK>(m, row, k, -zs, zzh);
                    }
                };
                // full groups of four rows run unguarded with a running row pointer (the per-step bound check
                // and row-offset multiply were 16 of the 344 instructions of a group); only the table's last
                // chunk can end in a partial group
                const unsigned int n4 = nr & ~3u;
                const Node* row = buf;
                float z[4];
#pragma unroll 1
                for (unsigned int r = 0; r < n4; r += 4) {
                    const uint4 wn = philox_words(a.seed, gpath, ((t0 + r) >> 2) + 1, 0u);
                    normals_from_words_scaled(w, k.bm_dt, z);  // z = sqrt(dt) Z (the scale rides in the radius)
                    w = wn;
#pragma unroll
                    for (int j = 0; j < 4; ++j, row += n_ld) step(row, z[j]);
                }
                if (n4 < nr) {
                    normals_from_words_scaled(w, k.bm_dt, z);
                    for (unsigned int j = 0; j < nr - n4; ++j, row += n_ld) step(row, z[j]);
                }
            }
            if (OPTP_LV_TMA) {
                __syncthreads();  // every thread is done reading this buffer
                if (threadIdx.x == 0 && g + 2 < total_chunks) issue(g + 2);
            }
        }
        if (active) {
            const int left_p = (p.pmin < 0.f || p.pmax > k.pos_max) ? 1 : 0;
            const int left_m = (ANTI && (m.pmin < 0.f || m.pmax > k.pos_max)) ? 1 : 0;
            acc_add(acc, 2, (double)(left_p + left_m));
            path_epilogue(a, finish_lin<MASK>(p, inv_n), df, b, n_eff, acc);
            if (ANTI) path_epilogue(a, finish_lin<MASK>(m, inv_n), df, a.n_base + b, n_eff, acc);
        }
    }
    reduce_and_emit(acc, blockvals, a.n_pay, (double)n_eff, a.partials, a.ticket, a.stats);
}

// Same Milstein step, but the window nodes are gathered straight from global memory through L1
// (ld.global.nc): no shared-memory staging, hence no CTA-wide barrier every chunk and no chunk
// copy on the critical path.  The rows in flight (all warps of an SM are within a few steps of each
// other) are a few tens of KB and stay L1-resident; the 1 MB table stays L2-resident.
template <int MASK, bool ANTI>
__global__ void __launch_bounds__(256, 4) lvm_direct_kernel(PathArgs a) {
    extern __shared__ double dyn[];
    __shared__ double blockvals[kMaxVals];
    double* acc = dyn;
    acc_zero(acc, 3 + 3 * a.n_pay);
    const unsigned int n_steps = a.n_steps;
    const double dt = a.T / (double)n_steps;
    constexpr double LN2 = 0.6931471805599453;
    LvConst k;
    k.inv_h2 = (float)(LN2 / a.h);
    k.c0 = (float)(-a.x0 / a.h);
    k.pos_max = (float)(a.n_nodes - 1);
    k.c_up = (float)(log1p(a.bump) / a.h);
    k.c_dn = (float)(log1p(-fmin(a.bump, 0.999999)) / a.h);
    k.k_up = (float)((1.0 + a.bump) / (2.0 * a.bump));
    k.k_dn = (float)((1.0 - a.bump) / (2.0 * a.bump));
    k.c1 = (float)((a.r - a.q) * dt);
    k.sdt = (float)sqrt(dt);
    k.hdt = (float)(0.5 * dt);
    k.bm_dt = (float)(-1.3862943611198906 * dt);
    const float S0 = (float)a.S0;
    const float inv_n = (float)(1.0 / (double)n_steps);
    const float df = (float)exp(-a.pay[0].r_disc * a.pay[0].T_disc);
    const unsigned long long n_eff = ANTI ? 2 * a.n_base : a.n_base;
    const int n_ld = a.n_nodes + 4;
    const unsigned int n_full = n_steps >> 2, n_tail = n_steps & 3;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < a.n_base;
         b += stride) {
        const unsigned long long gpath = a.path_offset + b;
        LinState p, m;
        p.S = p.smax = p.smin = S0;
        m = p;
        p.pmin = p.pmax = m.pmin = m.pmax = 0.5f * k.pos_max;
        float z[4];
        uint4 w = philox_words(a.seed, gpath, 0u, 0u);
        const float4* row = a.table4;
#pragma unroll 1
        for (unsigned int c = 0; c < n_full; ++c) {
            const uint4 wn = philox_words(a.seed, gpath, c + 1, 0u);
            normals_from_words_scaled(w, k.bm_dt, z);
            w = wn;
#pragma unroll
            for (int j = 0; j < 4; ++j, row += n_ld) {
                const float zs = z[j];
                const float zzh = fmaf(0.5f * zs, zs, -k.hdt);
                lvm_step_win<MASK, true>(p, row, k, zs, zzh);
                if (ANTI) lvm_step_win<MASK, true>(m, row, k, -zs, zzh);
            }
        }
        if (n_tail) {
            normals_from_words_scaled(w, k.bm_dt, z);
            for (unsigned int j = 0; j < n_tail; ++j, row += n_ld) {
                const float zs = z[j];
                const float zzh = fmaf(0.5f * zs, zs, -k.hdt);
                lvm_step_win<MASK, true>(p, row, k, zs, zzh);
                if (ANTI) lvm_step_win<MASK, true>(m, row, k, -zs, zzh);
            }
        }
        {
            const int left_p = (p.pmin < 0.f || p.pmax > k.pos_max) ? 1 : 0;
            const int left_m = (ANTI && (m.pmin < 0.f || m.pmax > k.pos_max)) ? 1 : 0;
            acc_add(acc, 2, (double)(left_p + left_m));
        }
        path_epilogue(a, finish_lin<MASK>(p, inv_n), df, b, n_eff, acc);
        if (ANTI) path_epilogue(a, finish_lin<MASK>(m, inv_n), df, a.n_base + b, n_eff, acc);
    }
    reduce_and_emit(acc, blockvals, a.n_pay, (double)n_eff, a.partials, a.ticket, a.stats);
}

template <int MASK, bool ANTI>
static int launch_lvm_direct(const PathArgs& a, cudaStream_t st) {
    const int threads = 256;
    const size_t smem = sizeof(double) * (3 + 3 * a.n_pay) * threads;
    auto kern = lvm_direct_kernel<MASK, ANTI>;
    if (smem > 48 * 1024)
        OPTP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    OPTP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) per_sm = 1;
    unsigned long long blocks = (a.n_base + threads - 1) / threads;
    unsigned long long cap = (unsigned long long)ctx().n_sm * per_sm;
    if (cap > kMaxBlocks) cap = kMaxBlocks;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned int)blocks, threads, smem, st>>>(a);
    ctx().launches++;
    OPTP_CUDA(cudaGetLastError());
    return 0;
}

// fp64 table -> (value, slope) fp32 pairs, rows padded to an even node count (16-byte rows)
__global__ void table_convert_kernel(const double* t64, float2* t32, unsigned int n_steps,
                                     int n_nodes) {
    const int n_ld = (n_nodes + 1) & ~1;
    const size_t n = (size_t)n_steps * n_ld;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(i % n_ld);
        const size_t r = i / n_ld;
        const int jc = j < n_nodes ? j : n_nodes - 1;
        const double v = t64[r * n_nodes + jc];
        const double nx = (jc + 1 < n_nodes) ? t64[r * n_nodes + jc + 1] : v;
        t32[i] = make_float2((float)v, (float)(nx - v));
    }
}

// fp64 table -> window nodes (v_j, s_j, s_{j+1}, s_{j-1}), j = -2 .. n+1 (edges replicated)
__global__ void table_window_kernel(const double* t64, float4* t4, unsigned int n_steps, int n_nodes) {
    const int n_ld = n_nodes + 4;
    const size_t n = (size_t)n_steps * n_ld;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(i % n_ld) - 2;
        const double* row = t64 + (i / n_ld) * (size_t)n_nodes;
        auto V = [&](int q) { return row[min(max(q, 0), n_nodes - 1)]; };
        t4[i] = make_float4((float)V(j), (float)(V(j + 1) - V(j)), (float)(V(j + 2) - V(j + 1)),
                            (float)(V(j) - V(j - 1)));
    }
}

int launch_table_convert(const double* t64, float2* t32, float4* t4, unsigned int n_steps, int n_nodes,
                         cudaStream_t st) {
    table_convert_kernel<<<64, 256, 0, st>>>(t64, t32, n_steps, n_nodes);
    table_window_kernel<<<64, 256, 0, st>>>(t64, t4, n_steps, n_nodes);
    ctx().launches += 2;
    OPTP_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------ dispatch
bool path_fp32_supported(int proc) {
    return proc == OPTP_PROC_GBM || proc == OPTP_PROC_HESTON_FTE || proc == OPTP_PROC_GBM_MILSTEIN ||
           proc == OPTP_PROC_HESTON_QE || proc == OPTP_PROC_MERTON || proc == OPTP_PROC_SABR ||
           proc == OPTP_PROC_LV_MILSTEIN || proc == OPTP_PROC_LV_EULER;
}

template <int PROC, int MASK, bool ANTI, int RNG = RNG_PHILOX10>
static int launch_generic(const PathArgs& a, cudaStream_t st) {
    const int threads = 256;
    const size_t smem = sizeof(double) * (3 + 3 * a.n_pay) * threads;
    if constexpr ((PROC == OPTP_PROC_GBM || PROC == OPTP_PROC_HESTON_FTE) && RNG == RNG_PHILOX10) {  // opt-in generator modes
        if (a.rng_mode == RNG_PHILOX7) return launch_generic<PROC, MASK, ANTI, RNG_PHILOX7>(a, st);
        if (a.rng_mode == RNG_PHILOX10_X6) return launch_generic<PROC, MASK, ANTI, RNG_PHILOX10_X6>(a, st);
    }
    if (smem > 48 * 1024)
        OPTP_CUDA(cudaFuncSetAttribute(path_fp32_kernel<PROC, MASK, ANTI, RNG>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    OPTP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
        &per_sm, path_fp32_kernel<PROC, MASK, ANTI, RNG>, threads, smem));
    if (per_sm < 1) per_sm = 1;
    unsigned long long blocks = (a.n_base + threads - 1) / threads;
    unsigned long long cap = (unsigned long long)ctx().n_sm * per_sm;
    if (cap > kMaxBlocks) cap = kMaxBlocks;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    path_fp32_kernel<PROC, MASK, ANTI, RNG><<<(unsigned int)blocks, threads, smem, st>>>(a);
    ctx().launches++;
    OPTP_CUDA(cudaGetLastError());
    return 0;
}

template <int PROC, int MASK, bool ANTI, int WINMODE>
static int launch_lv_impl(const PathArgs& a, cudaStream_t st) {
    constexpr bool WIN = WINMODE != 0;
    const size_t acc_bytes = sizeof(double) * (3 + 3 * a.n_pay) * kLvThreads;
    const size_t row_bytes = WIN ? sizeof(float4) * (size_t)(a.n_nodes + 4)
                                 : sizeof(float2) * (size_t)((a.n_nodes + 1) & ~1);
    // chunk of rows: a multiple of 4 (one Philox call = 4 steps).  The CTA meets at a barrier once per chunk, so
    // the chunk is as large as TWO resident CTAs per SM allow (two buffers each with the TMA staging): measured
    // on B200, config 3 (252 x 260-node rows of 4160 B): 4 rows per chunk (4 CTAs/SM) 5.02e11 path-steps/s,
    // 8 rows 5.23e11, 12 rows 5.48e11, 16 rows (one CTA per SM) 4.07e11.  OPTP_LV_CHUNK_KB overrides (A/B runs).
    static const int chunk_kb = getenv("OPTP_LV_CHUNK_KB") ? atoi(getenv("OPTP_LV_CHUNK_KB")) : 0;  // per buffer
    const size_t per_cta = (227 * 1024) / 2 - 1024;  // dynamic shared memory of one of two resident CTAs
    const size_t per_buf = per_cta > acc_bytes ? (per_cta - acc_bytes) / (OPTP_LV_TMA ? 2 : 1) : 0;
    int rows = (int)((chunk_kb > 0 ? (size_t)chunk_kb * 1024 : per_buf) / row_bytes) & ~3;
    if (rows < 4) rows = 4;
    const int max_rows = (int)((a.n_steps + 3) & ~3u);
    if (rows > max_rows) rows = max_rows;
    const size_t smem = acc_bytes + (OPTP_LV_TMA ? 2 : 1) * rows * row_bytes;
    if (smem > 200 * 1024)
        return set_error(OPTP_E_TOO_LARGE, "sigma table rows of %d nodes do not fit shared memory",
                         a.n_nodes);
    auto kern = lv_fp32_kernel<PROC, MASK, ANTI, WINMODE>;
    if (smem > 48 * 1024)
        OPTP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    OPTP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kLvThreads, smem));
    if (per_sm < 1) per_sm = 1;
    unsigned long long blocks = (a.n_base + kLvThreads - 1) / kLvThreads;
    unsigned long long cap = (unsigned long long)ctx().n_sm * per_sm;
    if (cap > kMaxBlocks) cap = kMaxBlocks;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned int)blocks, kLvThreads, smem, st>>>(a, rows);
    ctx().launches++;
    OPTP_CUDA(cudaGetLastError());
    return 0;
}

template <int MASK, bool ANTI>
static int launch_lvm_direct(const PathArgs& a, cudaStream_t st);

template <int PROC, int MASK, bool ANTI>
static int launch_lv(const PathArgs& a, cudaStream_t st) {
    if constexpr (PROC == OPTP_PROC_LV_MILSTEIN) {
        // one-load window lookup when the bump S(1 +- b) stays within one cell of ln S
        const double cu = log1p(a.bump) / a.h, cd = -log1p(-fmin(a.bump, 0.999999)) / a.h;
        if (cu < 1.0 && cd < 1.0 && !ctx().lv_no_window) {
            if (ctx().lv_direct) return launch_lvm_direct<MASK, ANTI>(a, st);
            return a.table_in_clip ? launch_lv_impl<PROC, MASK, ANTI, 2>(a, st)
                                       : launch_lv_impl<PROC, MASK, ANTI, 1>(a, st);
        }
    }
    return launch_lv_impl<PROC, MASK, ANTI, 0>(a, st);
}

template <int PROC, int MASK>
static int launch_anti(const PathArgs& a, cudaStream_t st) {
    if constexpr (PROC == OPTP_PROC_LV_MILSTEIN || PROC == OPTP_PROC_LV_EULER)
        return a.anti ? launch_lv<PROC, MASK, true>(a, st) : launch_lv<PROC, MASK, false>(a, st);
    else
        return a.anti ? launch_generic<PROC, MASK, true>(a, st)
                      : launch_generic<PROC, MASK, false>(a, st);
}

template <int PROC>
static int launch_mask(const PathArgs& a, int mask, cudaStream_t st) {
    switch (mask) {
        case 0: return launch_anti<PROC, 0>(a, st);
        case M_MAX: return launch_anti<PROC, M_MAX>(a, st);
        case M_MIN: return launch_anti<PROC, M_MIN>(a, st);
        case M_MAX | M_MIN: return launch_anti<PROC, M_MAX | M_MIN>(a, st);
        case M_SUM: return launch_anti<PROC, M_SUM>(a, st);
        case M_LOG: return launch_anti<PROC, M_LOG>(a, st);
        default: return launch_anti<PROC, M_MAX | M_MIN | M_SUM | M_LOG>(a, st);
    }
}

int launch_path_fp32(const PathArgs& a, cudaStream_t st) {
    const int mask = cover_mask(fine_mask(a.pay, a.n_pay));
    switch (a.proc) {
        case OPTP_PROC_GBM: return launch_mask<OPTP_PROC_GBM>(a, mask, st);
        case OPTP_PROC_HESTON_FTE: return launch_mask<OPTP_PROC_HESTON_FTE>(a, mask, st);
        case OPTP_PROC_HESTON_QE: return launch_mask<OPTP_PROC_HESTON_QE>(a, mask, st);
        case OPTP_PROC_MERTON: return launch_mask<OPTP_PROC_MERTON>(a, mask, st);
        case OPTP_PROC_SABR: return launch_mask<OPTP_PROC_SABR>(a, mask, st);
        case OPTP_PROC_GBM_MILSTEIN: return launch_mask<OPTP_PROC_GBM_MILSTEIN>(a, mask, st);
        case OPTP_PROC_LV_MILSTEIN: return launch_mask<OPTP_PROC_LV_MILSTEIN>(a, mask, st);
        case OPTP_PROC_LV_EULER: return launch_mask<OPTP_PROC_LV_EULER>(a, mask, st);
        default: return set_error(OPTP_E_KIND, "process kind %d has no fp32 kernel", a.proc);
    }
}

}  // namespace optp
