// fp64 Monte Carlo kernels in the REFERENCE'S OPERATION ORDER (compiled with -fmad=false).
// They serve the deterministic-check mode (normals supplied by the caller: the reference's
// own PCG64 draws), the fp64-with-Philox mode, and path materialisation for API parity.
// One thread per effective column; columns [0,n_base) are base paths, [n_base,2n_base) the
// antithetic mirrors (processes.py:49-50).
//   terminal:  _mc_chunk_sumstats            monte_carlo.py:28-64
//   GBM:       gbm_paths                     processes.py:44-55
//   Heston:    heston_paths                  processes.py:134-160
//   LV Milst.: milstein_local_vol_paths      processes.py:345-379
//   LV Euler:  local_vol_paths               processes.py:240-256
//   GBM Milst: gbm_milstein_paths            processes.py:284-303
//   SABR:      sabr_paths                    processes.py:185-215
#include <math.h>
#include "mc_kernels.cuh"

namespace optp {

constexpr uint32_t kTermStream = 0x7465726du;

// ------------------------------------------------------------------ terminal
__global__ void __launch_bounds__(256) terminal_fp64_kernel(TermArgs a) {
    __shared__ double red[32];
    __shared__ double blockvals[8];
    const double mu = (a.r - a.q - 0.5 * a.sigma * a.sigma) * a.T;  // monte_carlo.py:30
    const double sig = a.sigma * sqrt(a.T);
    const double df = exp(-a.r * a.T);
    const float muf = (float)mu, sigf = (float)sig, dff = (float)df, S0f = (float)a.S0,
                Kf = (float)a.K;
    const unsigned long long n_eff = a.anti ? 2 * a.n_base : a.n_base;
    double sx = 0, sxx = 0, sy = 0, syy = 0, sxy = 0;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; c < n_eff;
         c += stride) {
        const unsigned long long b = c < a.n_base ? c : c - a.n_base;
        double z;
        if (a.use_philox) {
            const unsigned long long g = a.path_offset + b;
            float zz[4];
            philox_normals4(a.seed, g >> 2, 0u, kTermStream, zz);
            z = (double)zz[g & 3];
        } else {
            z = a.z[b];
        }
        if (c >= a.n_base) z = -z;
        double X, Y;
        if (a.fp32_cast) {  // dtype=np.float32: monte_carlo.py:35,42,46,54-55
            const float zf = (float)z;
            const float st = S0f * expf(muf + sigf * zf);
            const float pay = a.is_call ? fmaxf(st - Kf, 0.0f) : fmaxf(Kf - st, 0.0f);
            const float xf = dff * pay, yf = dff * st;
            X = (double)xf;
            Y = (double)yf;
            sxx += (double)(xf * xf);
            syy += (double)(yf * yf);
            sxy += (double)(xf * yf);
        } else {
            const double st = a.S0 * exp(mu + sig * z);
            const double pay = a.is_call ? fmax(st - a.K, 0.0) : fmax(a.K - st, 0.0);
            X = df * pay;
            Y = df * st;
            sxx += X * X;
            syy += Y * Y;
            sxy += X * Y;
        }
        sx += X;
        sy += Y;
    }
    double vals[5] = {sx, sxx, sy, syy, sxy};
    for (int i = 0; i < 5; ++i) {
        const double t = block_sum(vals[i], red);
        if (threadIdx.x == 0) blockvals[i] = t;
    }
    __syncthreads();
    double* stats = a.stats;
    const double n = (double)n_eff;
    finalize_partials(a.partials, a.ticket, blockvals, 5, [=](int i, double s) {
        stats[1 + i] = s;
        if (i == 0) {
            stats[0] = n;
            stats[6] = 0.0;
            stats[7] = 0.0;
        }
    });
}

int launch_terminal_fp64(const TermArgs& a, cudaStream_t st) {
    const unsigned long long n_eff = a.anti ? 2 * a.n_base : a.n_base;
    unsigned long long blocks = (n_eff + 255) / 256;
    const unsigned long long cap = (unsigned long long)ctx().n_sm * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    terminal_fp64_kernel<<<(unsigned int)blocks, 256, 0, st>>>(a);
    ctx().launches++;
    OPTP_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------ sigma(S,t) table, fp64
// Identical arithmetic to oracle/sigma_grid.py (the table callable that drives the
// reference's Milstein recursion as the oracle).
__device__ __forceinline__ double table_sigma64(const double* row, int n, double x0, double inv_h,
                                                double S, double& clamps) {
    double pos = (log(S) - x0) * inv_h;
    if (pos < 0.0 || pos > (double)(n - 1)) clamps += 1.0;
    pos = fmin(fmax(pos, 0.0), (double)(n - 1));
    const double jf = fmin(floor(pos), (double)(n - 2));
    const double w = pos - jf;
    const int j = (int)jf;
    const double lo = row[j];
    return lo + w * (row[j + 1] - lo);
}

// ------------------------------------------------------------------ paths
template <int PROC>
__global__ void __launch_bounds__(128) path_fp64_kernel(PathArgs a) {
    extern __shared__ double dyn[];
    __shared__ double blockvals[kMaxVals];
    const int n_vals = 3 + 3 * a.n_pay;
    double* acc = dyn;
    acc_zero(acc, n_vals);

    constexpr int kDims =
        (PROC == OPTP_PROC_HESTON_FTE || PROC == OPTP_PROC_SABR || PROC == OPTP_PROC_HESTON_QE) ? 2 : 1;
    constexpr int kPerCall = (PROC == OPTP_PROC_MERTON) ? 1 : 4 / kDims;  // Merton: one call per step
    const unsigned long long n_eff = a.anti ? 2 * a.n_base : a.n_base;
    const unsigned int n_steps = a.n_steps;
    const double dt = a.T / (double)n_steps;
    const double sdt = sqrt(dt);
    const double df = exp(-a.pay[0].r_disc * a.pay[0].T_disc);
    const int mask = observer_mask(a.pay, a.n_pay);
    const size_t zstream = (size_t)n_steps * a.n_base;

    // per-process constants, same expression order as the reference
    const double gbm_drift = (a.r - a.q - 0.5 * a.sigma * a.sigma) * dt;
    const double gbm_vol = a.sigma * sdt;
    // Merton: v0 = lam, kappa = mJ, theta = sJ; compensator kappa_J = exp(mJ + sJ^2/2) - 1 (processes.py:84-85)
    const double mer_kj = exp(a.kappa + 0.5 * a.theta * a.theta) - 1.0;
    const double mer_drift = (a.r - a.q - 0.5 * a.sigma * a.sigma - a.v0 * mer_kj) * dt;
    const double rhoc = sqrt(fmax(0.0, 1.0 - a.rho * a.rho));
    const double inv_h = 1.0 / a.h;
    const double rq = a.r - a.q;
    // Heston QE constants (Andersen 2008, central discretisation gamma1 = gamma2 = 1/2)
    const double qe_E = exp(-a.kappa * dt);
    const double qe_c1 = a.xi * a.xi * qe_E * (1.0 - qe_E) / a.kappa;
    const double qe_c2 = a.theta * a.xi * a.xi * ((1.0 - qe_E) * (1.0 - qe_E)) / (2.0 * a.kappa);
    const double qe_K0 = -a.rho * a.kappa * a.theta * dt / a.xi;
    const double qe_K1 = 0.5 * dt * (a.kappa * a.rho / a.xi - 0.5) - a.rho / a.xi;
    const double qe_K2 = 0.5 * dt * (a.kappa * a.rho / a.xi - 0.5) + a.rho / a.xi;
    const double qe_K3 = 0.5 * dt * (1.0 - a.rho * a.rho);

    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; c < n_eff;
         c += stride) {
        const unsigned long long b = c < a.n_base ? c : c - a.n_base;
        const double sgn = c < a.n_base ? 1.0 : -1.0;
        const unsigned long long gpath = a.path_offset + b;
        double S = a.S0, logs = 0.0, clamps = 0.0;
        double v = 0.0;
        if (PROC == OPTP_PROC_HESTON_FTE || PROC == OPTP_PROC_HESTON_QE) v = fmax(a.v0, 0.0);  // processes.py:150
        if (PROC == OPTP_PROC_HESTON_QE) logs = log(a.S0);
        if (PROC == OPTP_PROC_SABR) v = a.v0;                   // alpha0, processes.py:199
        double smax = S, smin = S, sumS = 0.0, sumLog = 0.0;
        if (a.paths_out) a.paths_out[c] = S;
        if (a.var_out) a.var_out[c] = v;
        float zz[4] = {0.f, 0.f, 0.f, 0.f};
        for (unsigned int t = 0; t < n_steps; ++t) {
            double z0, z1 = 0.0, kj = 0.0;
            if (PROC == OPTP_PROC_MERTON) {
                // Z (mirrored), Poisson count (shared by the pair), ZJ (mirrored): processes.py:89-99
                if (a.use_philox) {
                    const uint4 w = philox_words(a.seed, gpath, t, 0u);
                    float n0, n1;
                    box_muller(w.x, w.y, n0, n1);
                    z0 = sgn * (double)n0;
                    z1 = sgn * (double)n1;
                    kj = (double)poisson_inv(w.z, (float)(a.v0 * dt), (float)exp(-a.v0 * dt));
                } else {
                    const size_t o = (size_t)t * a.n_base + b;
                    z0 = sgn * a.z[o];
                    kj = a.z[zstream + o];
                    z1 = sgn * a.z[2 * zstream + o];
                }
            } else if (a.use_philox) {
                const unsigned int j = t % kPerCall;
                if (j == 0) philox_normals4(a.seed, gpath, t / kPerCall, 0u, zz);
                z0 = sgn * (double)zz[j * kDims];
                if (kDims == 2) z1 = sgn * (double)zz[j * kDims + 1];
            } else {
                const size_t o = (size_t)t * a.n_base + b;
                z0 = sgn * a.z[o];
                if (kDims == 2) z1 = sgn * a.z[zstream + o];
            }
            if (PROC == OPTP_PROC_MERTON) {
                // log increment = drift + vol*Z + (mJ*K + sJ*sqrt(K)*ZJ), processes.py:102-106
                const double ysum = a.kappa * kj + a.theta * sqrt(kj) * z1;
                logs += mer_drift + gbm_vol * z0 + ysum;
                S = a.S0 * exp(logs);
            } else if (PROC == OPTP_PROC_GBM) {
                logs += gbm_drift + gbm_vol * z0;  // cumsum, processes.py:53
                S = a.S0 * exp(logs);
            } else if (PROC == OPTP_PROC_HESTON_FTE) {
                const double zc = a.rho * z0 + rhoc * z1;  // Z1 = rho*Z2 + ..*Zp, processes.py:143
                const double vp = fmax(v, 0.0);
                const double sq = sqrt(vp);
                S = S * exp((rq - 0.5 * vp) * dt + sq * sdt * zc);
                v = v + a.kappa * (a.theta - vp) * dt + a.xi * sq * sdt * z0;
                v = fmax(v, 0.0);
            } else if (PROC == OPTP_PROC_HESTON_QE) {
                // Andersen (2008) QE, same expression order as oracle/path_schemes.py:heston_qe
                const double m = a.theta + (v - a.theta) * qe_E;
                const double s2 = v * qe_c1 + qe_c2;
                const double psi = s2 / (m * m);
                double vn;
                if (psi <= 1.5) {
                    const double tt = 2.0 / psi;
                    const double b2 = fmax(tt - 1.0 + sqrt(tt) * sqrt(fmax(tt - 1.0, 0.0)), 0.0);
                    const double aa = m / (1.0 + b2);
                    const double bz = sqrt(b2) + z0;
                    vn = aa * (bz * bz);
                } else {
                    const double p = (psi - 1.0) / (psi + 1.0);
                    const double beta = (1.0 - p) / m;
                    const double u = 0.5 * erfc(-z0 * 0.7071067811865476);
                    const double omu = 0.5 * erfc(z0 * 0.7071067811865476);
                    vn = (u <= p) ? 0.0 : log((1.0 - p) / omu) / beta;
                }
                double k0 = qe_K0 + qe_K1 * v;
                if (a.qe_martingale) {  // Andersen's K0* (oracle/path_schemes.py:heston_qe, martingale_correction)
                    const double A = qe_K2 + 0.5 * qe_K3;
                    if (psi <= 1.5) {
                        const double tt = 2.0 / psi;
                        const double b2 = fmax(tt - 1.0 + sqrt(tt) * sqrt(fmax(tt - 1.0, 0.0)), 0.0);
                        const double aa = m / (1.0 + b2);
                        const double den = 1.0 - 2.0 * A * aa;
                        if (den > 0.0) k0 = -(A * b2 * aa / den - 0.5 * log(den)) - 0.5 * qe_K3 * v;
                    } else {
                        const double p = (psi - 1.0) / (psi + 1.0);
                        const double beta = (1.0 - p) / m;
                        const double den = beta - A;
                        if (den > 0.0) k0 = -log(p + beta * (1.0 - p) / den) - 0.5 * qe_K3 * v;
                    }
                }
                logs = logs + rq * dt + k0 + qe_K2 * vn + sqrt(fmax(qe_K3 * (v + vn), 0.0)) * z1;
                v = vn;
                S = exp(logs);
            } else if (PROC == OPTP_PROC_LV_MILSTEIN) {
                const double* row = a.table64 + (size_t)t * a.n_nodes;
                double sig = table_sigma64(row, a.n_nodes, a.x0, inv_h, S, clamps);
                sig = fmin(fmax(sig, 1e-8), 10.0);
                const double eps = a.bump * S;
                const double s_up = S + eps;
                const double s_dn = fmax(S - eps, 1e-10);
                double dummy = 0.0;
                const double sig_up = table_sigma64(row, a.n_nodes, a.x0, inv_h, s_up, dummy);
                const double sig_dn = table_sigma64(row, a.n_nodes, a.x0, inv_h, s_dn, dummy);
                const double da = (sig_up * s_up - sig_dn * s_dn) / (s_up - s_dn);
                const double aa = sig * S;
                S = S + rq * S * dt + aa * sdt * z0 + 0.5 * aa * da * (z0 * z0 - 1.0) * dt;
                S = fmax(S, 1e-10);
            } else if (PROC == OPTP_PROC_LV_EULER) {
                const double* row = a.table64 + (size_t)t * a.n_nodes;
                double sig = table_sigma64(row, a.n_nodes, a.x0, inv_h, S, clamps);
                sig = fmax(sig, 0.0);
                S = S * exp((rq - 0.5 * sig * sig) * dt + sig * sdt * z0);
            } else if (PROC == OPTP_PROC_GBM_MILSTEIN) {
                S = S + rq * S * dt + a.sigma * S * sdt * z0 +
                    0.5 * (a.sigma * a.sigma) * S * (z0 * z0 - 1.0) * dt;
                S = fmax(S, 1e-10);
            } else if (PROC == OPTP_PROC_SABR) {
                const double zc = a.rho * z0 + rhoc * z1;
                const double nu = a.xi, beta = a.kappa;
                v = v * exp(nu * sdt * z0 - 0.5 * (nu * nu) * dt);
                if (beta == 1.0) {
                    S = S * exp((rq - 0.5 * (v * v)) * dt + v * sdt * zc);
                } else {
                    S = S + rq * S * dt + v * pow(S, beta) * sdt * zc;
                    S = fmax(S, 1e-12);
                }
            }
            if (mask & OBS_EXT) {
                smax = fmax(smax, S);
                smin = fmin(smin, S);
            }
            if (mask & OBS_SUMS) sumS += S;
            if (mask & OBS_SUMLOG) sumLog += log(S);
            if (a.paths_out) a.paths_out[(size_t)(t + 1) * n_eff + c] = S;
            if (a.var_out) a.var_out[(size_t)(t + 1) * n_eff + c] = v;
        }
        const double avgA = sumS / (double)n_steps;
        const double avgG = (mask & OBS_SUMLOG) ? exp(sumLog / (double)n_steps) : 0.0;
        if (a.terminal_out) a.terminal_out[c] = S;
        const double Y = df * S;
        acc_add(acc, 0, Y);
        acc_add(acc, 1, Y * Y);
        acc_add(acc, 2, clamps);
        for (int p = 0; p < a.n_pay; ++p) {
            const double pay = eval_payoff<double>(a.pay[p], S, smax, smin, avgA, avgG);
            if (a.payoff_out) a.payoff_out[(size_t)p * n_eff + c] = pay;
            const double X = df * pay;
            acc_add(acc, 3 + 3 * p, X);
            acc_add(acc, 4 + 3 * p, X * X);
            acc_add(acc, 5 + 3 * p, X * Y);
        }
    }
    reduce_and_emit(acc, blockvals, a.n_pay, (double)n_eff, a.partials, a.ticket, a.stats);
}

template <int PROC>
static int launch_one64(const PathArgs& a, cudaStream_t st) {
    const unsigned long long n_eff = a.anti ? 2 * a.n_base : a.n_base;
    const int threads = 128;
    const size_t smem = sizeof(double) * (3 + 3 * a.n_pay) * threads;
    if (smem > 48 * 1024)
        OPTP_CUDA(cudaFuncSetAttribute(path_fp64_kernel<PROC>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned long long blocks = (n_eff + threads - 1) / threads;
    const unsigned long long cap = (unsigned long long)ctx().n_sm * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    path_fp64_kernel<PROC><<<(unsigned int)blocks, threads, smem, st>>>(a);
    ctx().launches++;
    OPTP_CUDA(cudaGetLastError());
    return 0;
}

int launch_path_fp64(const PathArgs& a, cudaStream_t st) {
    switch (a.proc) {
        case OPTP_PROC_GBM: return launch_one64<OPTP_PROC_GBM>(a, st);
        case OPTP_PROC_HESTON_FTE: return launch_one64<OPTP_PROC_HESTON_FTE>(a, st);
        case OPTP_PROC_HESTON_QE: return launch_one64<OPTP_PROC_HESTON_QE>(a, st);
        case OPTP_PROC_LV_MILSTEIN: return launch_one64<OPTP_PROC_LV_MILSTEIN>(a, st);
        case OPTP_PROC_LV_EULER: return launch_one64<OPTP_PROC_LV_EULER>(a, st);
        case OPTP_PROC_GBM_MILSTEIN: return launch_one64<OPTP_PROC_GBM_MILSTEIN>(a, st);
        case OPTP_PROC_SABR: return launch_one64<OPTP_PROC_SABR>(a, st);
        case OPTP_PROC_MERTON: return launch_one64<OPTP_PROC_MERTON>(a, st);
        default: return set_error(OPTP_E_KIND, "process kind %d has no fp64 kernel", a.proc);
    }
}

// ------------------------------------------------------------------ payoffs over given paths
// The literal exotics.py contract: reduce a caller-supplied (n_rows, n_cols) array.
__global__ void __launch_bounds__(128) payoff_from_paths_kernel(const double* paths,
                                                               unsigned long long n_rows,
                                                               unsigned long long n_cols,
                                                               optp_payoff pay, double* payoff_out,
                                                               double* partials,
                                                               unsigned int* ticket, double* stats) {
    extern __shared__ double dyn[];
    __shared__ double blockvals[kMaxVals];
    double* acc = dyn;
    acc_zero(acc, 6);
    const double df = exp(-pay.r_disc * pay.T_disc);
    const int mask = observer_mask(&pay, 1);
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; c < n_cols;
         c += stride) {
        double S = paths[c];
        double smax = S, smin = S, sumS = 0.0, sumLog = 0.0;
        for (unsigned long long t = 1; t < n_rows; ++t) {
            S = paths[t * n_cols + c];
            if (mask & OBS_EXT) {
                smax = fmax(smax, S);
                smin = fmin(smin, S);
            }
            if (mask & OBS_SUMS) sumS += S;
            if (mask & OBS_SUMLOG) sumLog += log(S);
        }
        const double nobs = (double)(n_rows - 1);
        const double avgA = sumS / nobs;
        const double avgG = (mask & OBS_SUMLOG) ? exp(sumLog / nobs) : 0.0;
        const double p = eval_payoff<double>(pay, S, smax, smin, avgA, avgG);
        if (payoff_out) payoff_out[c] = p;
        const double X = df * p, Y = df * S;
        acc_add(acc, 0, Y);
        acc_add(acc, 1, Y * Y);
        acc_add(acc, 3, X);
        acc_add(acc, 4, X * X);
        acc_add(acc, 5, X * Y);
    }
    reduce_and_emit(acc, blockvals, 1, (double)n_cols, partials, ticket, stats);
}

int launch_payoff_from_paths(const double* paths, unsigned long long n_rows,
                             unsigned long long n_cols, const optp_payoff& pay, double* payoff_out,
                             double* stats, cudaStream_t st) {
    Context& c = ctx();
    unsigned long long blocks = (n_cols + 127) / 128;
    const unsigned long long cap = (unsigned long long)c.n_sm * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    double* partials;
    unsigned int* ticket;
    reduction_scratch(&partials, &ticket);
    payoff_from_paths_kernel<<<(unsigned int)blocks, 128, sizeof(double) * 6 * 128, st>>>(
        paths, n_rows, n_cols, pay, payoff_out, partials, ticket, stats);
    c.launches++;
    OPTP_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace optp
