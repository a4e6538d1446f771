// Device-side tabulation of the Dupire local volatility of a raw-SVI surface (fp64, compiled with
// -fmad=false: it restates the reference's operation order).
//
// Replaces the host loop `for i: table[i] = sigma_loc(nodes, i*dt)` over the reference's callable
// dupire_local_vol_func(surface, r, q) (calibration.py:362-390 -> dupire_local_vol, :268-359) for the
// table the local-vol kernels read: one thread per (step time i, node j).
//   t = max(i*dt, 1e-8)                                      calibration.py:314
//   F = forward curve at t: exact key, else np.interp with flat ends        :317-320, :108-122
//   k = ln(S/F); analytic w, dw/dk, d2w/dk2 of the nearest-above slice      :325-334, :37-66
//   dw/dT by central difference (dT) on the T-interpolated surface          :337-343, :125-148
//   Gatheral's denominator floored at 1e-8, sigma clipped to [0.01, 5]      :346-355
// r and q do not enter the reference's formula (the forward comes from the surface's curve).
#include <math.h>
#include "mc_kernels.cuh"

namespace optp {

struct DupireArgs {
    const double* expiries;  // n_exp, ascending
    const double* svi;       // n_exp x 5: a, b, rho, m, sigma
    const double* fwd_T;     // n_fwd, ascending
    const double* fwd_F;
    int n_exp, n_fwd;
    double dt, x0, h, dT;
    unsigned int n_steps, n_nodes;
    double* table;           // n_steps x n_nodes
};

__device__ __forceinline__ double svi_w(const double* p, double k) {
    const double km = k - p[3];
    return p[0] + p[1] * (p[2] * km + sqrt(km * km + p[4] * p[4]));
}

// searchsorted(a, x, side="left"): first index with a[i] >= x
__device__ __forceinline__ int lower_bound(const double* a, int n, double x) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < x) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// VolSurface._get_forward, calibration.py:108-122
__device__ __forceinline__ double forward_at(const DupireArgs& a, double T) {
    const int i = lower_bound(a.fwd_T, a.n_fwd, T);
    if (i < a.n_fwd && a.fwd_T[i] == T) return a.fwd_F[i];
    if (a.n_fwd == 1) return a.fwd_F[0];
    if (i == 0) return a.fwd_F[0];                      // np.interp: flat below the first point
    if (i >= a.n_fwd) return a.fwd_F[a.n_fwd - 1];      // ... and above the last
    const double slope = (a.fwd_F[i] - a.fwd_F[i - 1]) / (a.fwd_T[i] - a.fwd_T[i - 1]);
    return slope * (T - a.fwd_T[i - 1]) + a.fwd_F[i - 1];
}

// VolSurface.iv_from_logm, calibration.py:125-148
__device__ __forceinline__ double iv_from_logm(const DupireArgs& a, double k, double T) {
    const int i = lower_bound(a.expiries, a.n_exp, T);
    int exact = -1;
    if (i < a.n_exp && a.expiries[i] == T) exact = i;
    else if (i == 0) exact = 0;
    else if (i >= a.n_exp) exact = a.n_exp - 1;
    if (exact >= 0) return sqrt(fmax(svi_w(a.svi + 5 * exact, k), 0.0) / a.expiries[exact]);
    const double lo = a.expiries[i - 1], hi = a.expiries[i];
    const double w_lo = svi_w(a.svi + 5 * (i - 1), k) * lo;
    const double w_hi = svi_w(a.svi + 5 * i, k) * hi;
    const double al = (T - lo) / (hi - lo);
    const double wT = (1.0 - al) * w_lo + al * w_hi;
    return sqrt(fmax(wT, 0.0) / T);
}

__global__ void __launch_bounds__(256) dupire_table_kernel(DupireArgs a) {
    const unsigned long long n = (unsigned long long)a.n_steps * a.n_nodes;
    for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < n;
         e += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned int i = (unsigned int)(e / a.n_nodes), j = (unsigned int)(e % a.n_nodes);
        const double S = exp(a.x0 + a.h * (double)j);
        const double t = fmax((double)i * a.dt, 1e-8);
        const double F = forward_at(a, t);
        const double k = log(S / F);
        int idx = lower_bound(a.expiries, a.n_exp, t);
        idx = max(0, min(idx, a.n_exp - 1));
        const double* p = a.svi + 5 * idx;
        const double w = fmax(svi_w(p, k), 1e-12);
        const double x = k - p[3];
        const double x2s2 = x * x + p[4] * p[4];
        const double dw = p[1] * (p[2] + x / sqrt(x2s2));
        const double d2w = p[1] * (p[4] * p[4]) / pow(x2s2, 1.5);
        const double t_up = t + a.dT;
        const double t_dn = fmax(t - a.dT, 1e-8);
        const double iv_up = iv_from_logm(a, k, t_up), iv_dn = iv_from_logm(a, k, t_dn);
        const double w_up = iv_up * iv_up * t_up, w_dn = iv_dn * iv_dn * t_dn;
        const double numer = fmax((w_up - w_dn) / (t_up - t_dn), 1e-12);
        const double kw = k / w;
        double denom = 1.0 - kw * dw + 0.25 * (-0.25 - 1.0 / w + kw * kw) * (dw * dw) + 0.5 * d2w;
        denom = fmax(denom, 1e-8);
        const double s = sqrt(fmax(numer / denom, 0.0));
        a.table[e] = fmin(fmax(s, 0.01), 5.0);
    }
}

}  // namespace optp

using namespace optp;

extern "C" int optp_dupire_table(const double* expiries, const double* svi, uint32_t n_exp,
                                 const double* fwd_T, const double* fwd_F, uint32_t n_fwd, double T,
                                 uint32_t n_steps, double x0, double h, uint32_t n_nodes, double dT,
                                 double* table_out) {
    OPTP_ENTER();
    if (!expiries || !svi || n_exp == 0)
        return set_error(OPTP_E_ARG, "At least one SVI slice is required.");
    if (!fwd_T || !fwd_F || n_fwd == 0)
        return set_error(OPTP_E_ARG, "a forward curve is required to tabulate the surface");
    if (n_steps == 0 || n_nodes < 2 || !(h > 0.0) || !(T > 0.0) || !(dT > 0.0))
        return set_error(OPTP_E_ARG, "dupire table needs n_steps > 0, n_nodes >= 2, h > 0, T > 0, dT > 0");
    for (uint32_t i = 1; i < n_exp; ++i)
        if (!(expiries[i] > expiries[i - 1])) return set_error(OPTP_E_ARG, "expiries must ascend");
    for (uint32_t i = 1; i < n_fwd; ++i)
        if (!(fwd_T[i] > fwd_T[i - 1])) return set_error(OPTP_E_ARG, "forward curve times must ascend");
    Context& c = ctx();
    OPTP_CUDA(cudaSetDevice(c.device));
    const size_t n = (size_t)n_steps * n_nodes;
    const size_t n_ld = (n_nodes + 1) & ~1u;
    // surface description: expiries | svi | fwd_T | fwd_F in one staging arena
    const size_t n_in = (size_t)n_exp * 6 + (size_t)n_fwd * 2;
    if (int e = arena_reserve(c.in1, n_in * sizeof(double))) return e;
    if (int e = arena_reserve(c.tables[0].f64, n * sizeof(double))) return e;
    if (int e = arena_reserve(c.tables[0].f32, (size_t)n_steps * n_ld * sizeof(float2))) return e;
    if (int e = arena_reserve(c.tables[0].f4, (size_t)n_steps * (n_nodes + 4) * sizeof(float4))) return e;
    double* in = (double*)c.in1.p;
    OPTP_CUDA(cudaMemcpyAsync(in, expiries, n_exp * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    OPTP_CUDA(cudaMemcpyAsync(in + n_exp, svi, 5 * (size_t)n_exp * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    OPTP_CUDA(cudaMemcpyAsync(in + 6 * (size_t)n_exp, fwd_T, n_fwd * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    OPTP_CUDA(cudaMemcpyAsync(in + 6 * (size_t)n_exp + n_fwd, fwd_F, n_fwd * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    DupireArgs a;
    a.expiries = in;
    a.svi = in + n_exp;
    a.fwd_T = in + 6 * (size_t)n_exp;
    a.fwd_F = a.fwd_T + n_fwd;
    a.n_exp = (int)n_exp;
    a.n_fwd = (int)n_fwd;
    a.dt = T / (double)n_steps;
    a.x0 = x0;
    a.h = h;
    a.dT = dT;
    a.n_steps = n_steps;
    a.n_nodes = n_nodes;
    a.table = (double*)c.tables[0].f64.p;
    const unsigned int grid = (unsigned int)((n + 255) / 256 < (size_t)c.n_sm * 8 ? (n + 255) / 256 : (size_t)c.n_sm * 8);
    dupire_table_kernel<<<grid, 256, 0, c.stream>>>(a);
    c.launches++;
    OPTP_CUDA(cudaGetLastError());
    // the fp32 (value, slope) and window tables the production kernels read
    if (int e = launch_table_convert((const double*)c.tables[0].f64.p, (float2*)c.tables[0].f32.p,
                                     (float4*)c.tables[0].f4.p, n_steps, (int)n_nodes, c.stream))
        return e;
    c.tables[0].steps = n_steps;
    c.tables[0].nodes = n_nodes;
    c.tables[0].in_clip = true;  // the tabulation clips to [0.01, 5] (calibration.py:355)
    if (table_out)
        OPTP_CUDA(cudaMemcpyAsync(table_out, c.tables[0].f64.p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    OPTP_CUDA(cudaStreamSynchronize(c.stream));
    return 0;
}
