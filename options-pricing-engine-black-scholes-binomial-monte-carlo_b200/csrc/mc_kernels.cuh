// Launch interface between the C-ABI front end (mc_host.cu) and the two Monte Carlo
// translation units: mc_fp32.cu (production arithmetic) and mc_fp64.cu (reference operation
// order, compiled with -fmad=false so nothing is contracted).
#pragma once
#include "optp_common.cuh"

namespace optp {

struct TermArgs {
    double S0, K, T, r, q, sigma;
    int is_call;
    unsigned long long n_base;
    int anti;
    int fp32_cast;
    int use_philox;
    unsigned long long seed, path_offset;
    const double* z;  // device, n_base normals (HOSTZ mode)
    double* partials;
    unsigned int* ticket;
    double* stats;  // device, OPTP_NSTATS
};

struct PathArgs {
    int proc;
    int n_nodes;
    double S0, r, q, T, sigma, v0, kappa, theta, xi, rho, bump, x0, h;
    const float2* table32;  // (value, next-value minus value) per node, n_steps rows
    const float4* table4;   // window nodes (v_j, s_j, s_{j+1}, s_{j-1}), n_nodes + 4 per row
    const double* table64;  // n_steps x n_nodes
    int table_in_clip;      // every node of the table lies in the reference's clip range [1e-8, 10]
    unsigned int n_steps;
    unsigned long long n_base;
    int anti;
    int use_philox;
    int rng_mode;  // RNG_* (optp_common.cuh); honoured by the fp32 GBM and Heston FTE path kernels
    int qe_martingale;  // OPTP_PROC_HESTON_QE_M: proc = HESTON_QE with the martingale-corrected K0*
    unsigned long long seed, path_offset;
    const double* z;  // device, n_streams x n_steps x n_base (HOSTZ mode)
    int n_pay;
    optp_payoff pay[kMaxPayoffs];
    double* payoff_out;    // n_pay x n_eff or null
    double* terminal_out;  // n_eff or null
    double* paths_out;     // (n_steps+1) x n_eff or null (fp64 kernels only)
    double* var_out;       // same shape or null (Heston)
    double* partials;
    unsigned int* ticket;
    double* stats;  // device, n_pay x OPTP_NSTATS
};

int launch_terminal_fp32(const TermArgs& a, cudaStream_t st);
int launch_terminal_fp64(const TermArgs& a, cudaStream_t st);
int launch_path_fp32(const PathArgs& a, cudaStream_t st);
int launch_path_fp64(const PathArgs& a, cudaStream_t st);
bool path_fp32_supported(int proc);
// builds the (value, slope) fp32 table from an fp64 device table
int launch_table_convert(const double* t64, float2* t32, float4* t4, unsigned int n_steps, int n_nodes,
                         cudaStream_t st);

// which per-step path state the payoffs need
enum : int { OBS_EXT = 1, OBS_SUMS = 2, OBS_SUMLOG = 4 };

__host__ __device__ inline int observer_mask(const optp_payoff* p, int n) {
    int m = 0;
    for (int i = 0; i < n; ++i) {
        if (p[i].kind == OPTP_PAY_BARRIER || p[i].kind == OPTP_PAY_LOOKBACK) m |= OBS_EXT;
        if (p[i].kind == OPTP_PAY_ASIAN)
            m |= (p[i].flags & OPTP_F_GEOMETRIC) ? OBS_SUMLOG : OBS_SUMS;
    }
    return m;
}

__host__ __device__ inline int n_streams_of(int proc) {
    if (proc == OPTP_PROC_MERTON) return 3;
    return (proc == OPTP_PROC_HESTON_FTE || proc == OPTP_PROC_SABR || proc == OPTP_PROC_HESTON_QE ||
            proc == OPTP_PROC_HESTON_QE_M) ? 2
                                                                                                  : 1;
}

#ifdef __CUDACC__
// Payoff at the end of a path from the running path state (exotics.py:77-94, 136-159,
// 201-205, 252-267).  smax/smin include the t=0 node, the averages exclude it.
template <typename R>
__device__ __forceinline__ R eval_payoff(const optp_payoff& p, R ST, R smax, R smin, R avgA,
                                         R avgG) {
    const bool call = (p.flags & OPTP_F_CALL) != 0;
    const R K = (R)p.K;
    const R zero = (R)0;
    switch (p.kind) {
        case OPTP_PAY_VANILLA:
            return call ? max(ST - K, zero) : max(K - ST, zero);
        case OPTP_PAY_DIGITAL:
            return call ? (ST > K ? (R)p.payout : zero) : (ST < K ? (R)p.payout : zero);
        case OPTP_PAY_BARRIER: {
            const R B = (R)p.barrier;
            const bool hit = (p.flags & OPTP_F_UP) ? (smax >= B) : (smin <= B);
            const R van = call ? max(ST - K, zero) : max(K - ST, zero);
            const bool pays_vanilla = (p.flags & OPTP_F_IN) ? hit : !hit;
            return pays_vanilla ? van : (R)p.rebate;
        }
        case OPTP_PAY_ASIAN: {
            const R avg = (p.flags & OPTP_F_GEOMETRIC) ? avgG : avgA;
            if (p.flags & OPTP_F_FLOATING) return call ? max(ST - avg, zero) : max(avg - ST, zero);
            return call ? max(avg - K, zero) : max(K - avg, zero);
        }
        case OPTP_PAY_LOOKBACK:
            if (p.flags & OPTP_F_FLOATING) return call ? ST - smin : smax - ST;
            return call ? max(smax - K, zero) : max(K - smin, zero);
        default:
            return zero;
    }
}

// Thread-private accumulators in shared memory: acc[value][thread].  Values: 0 sumY, 1 sumY2,
// 2 clamp count, then (sumX, sumX2, sumXY) per payoff.
__device__ __forceinline__ void acc_zero(double* acc, int n_vals) {
    for (int v = 0; v < n_vals; ++v) acc[v * blockDim.x + threadIdx.x] = 0.0;
}

__device__ __forceinline__ void acc_add(double* acc, int v, double x) {
    acc[v * blockDim.x + threadIdx.x] += x;
}

// Reduce acc over the block, publish block partials, last block writes the stats blocks.
__device__ __forceinline__ void reduce_and_emit(double* acc, double* blockvals, int n_pay,
                                                double n_eff, double* partials,
                                                unsigned int* ticket, double* stats) {
    const int n_vals = 3 + 3 * n_pay;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    for (int v = wid; v < n_vals; v += nw) {
        double s = 0.0;
        for (int t = lane; t < (int)blockDim.x; t += 32) s += acc[v * blockDim.x + t];
        s = warp_sum(s);
        if (lane == 0) blockvals[v] = s;
    }
    __syncthreads();
    finalize_partials(partials, ticket, blockvals, n_vals, [=](int i, double s) {
        if (i < 3) {
            const int slot = (i == 0) ? 3 : (i == 1) ? 4 : 6;
            for (int p = 0; p < n_pay; ++p) {
                stats[p * OPTP_NSTATS + slot] = s;
                if (i == 0) {
                    stats[p * OPTP_NSTATS + 0] = n_eff;
                    stats[p * OPTP_NSTATS + 7] = 0.0;
                }
            }
        } else {
            const int p = (i - 3) / 3, w = (i - 3) % 3;
            stats[p * OPTP_NSTATS + (w == 0 ? 1 : (w == 1 ? 2 : 5))] = s;
        }
    });
}
#endif

}  // namespace optp
