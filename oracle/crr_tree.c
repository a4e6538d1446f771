/* Oracle: Cox-Ross-Rubinstein backward induction in plain C (TEST INFRASTRUCTURE, see
 * oracle/__init__.py).  Two restatements of reference src/optpricer/binomial.py:7-38 for a
 * heterogeneous contract batch (the oracle SURVEY 8c names: a loop over `crr`):
 *
 *   oracle_crr_loop_ref   the reference's operation order, node by node: d = 1/u (:13),
 *                         terminal S0*u^j*d^(N-j) with two pow() (:20-21), V = disc*(p*V[j+1] +
 *                         (1-p)*V[j]) (:29), American exercise value recomputed with two pow() per
 *                         node per level (:30-36).  Matches the numpy reference to the last few ulps
 *                         (numpy's power is libm pow()).
 *   oracle_crr_loop_fast  the same lattice with S_kj = S0*u^(2j-k) from a table of u^m
 *                         (m = -N..N) -- no pow() in the level loop.  Differs from the reference by
 *                         O(1e-13) relative; it exists so parity samples of thousands of N = 1000
 *                         contracts finish in seconds, and as the optimised-C CPU baseline of bench.py.
 * Both set status 1 / price NaN when the risk-neutral probability is outside (0,1) (:16-17) and
 * treat any kind other than call as a put (:22-25).  Contracts are independent; callers that want
 * several cores split the batch over processes (this image has no OpenMP runtime).
 */
#include <math.h>
#include <stdlib.h>

static int lattice(double T, double r, double q, double sigma, int N, double* u, double* d,
                   double* disc, double* p) {
    const double dt = T / (double)N;
    *u = exp(sigma * sqrt(dt));
    *d = 1.0 / *u;
    *disc = exp(-r * dt);
    *p = (exp((r - q) * dt) - *d) / (*u - *d);
    return (*p > 0.0) && (*p < 1.0);
}

int oracle_crr_loop_ref(const double* S0, const double* K, const double* T, const double* r,
                        const double* q, const double* sigma, const unsigned char* is_call, long n,
                        int N, int american, double* price, unsigned char* status) {
    if (N <= 0) return 1;
    double* V = (double*)malloc(sizeof(double) * (size_t)(N + 1));
    if (!V) return 2;
    for (long o = 0; o < n; ++o) {
        double u, d, disc, p;
        if (!lattice(T[o], r[o], q[o], sigma[o], N, &u, &d, &disc, &p)) {
            price[o] = NAN;
            status[o] = 1;
            continue;
        }
        const int call = is_call[o] != 0;
        for (int j = 0; j <= N; ++j) {
            const double st = S0[o] * pow(u, (double)j) * pow(d, (double)(N - j));
            const double pay = call ? st - K[o] : K[o] - st;
            V[j] = pay > 0.0 ? pay : 0.0;
        }
        for (int k = N - 1; k >= 0; --k) {
            for (int j = 0; j <= k; ++j) {
                double v = disc * (p * V[j + 1] + (1.0 - p) * V[j]);
                if (american) {
                    const double s = S0[o] * pow(u, (double)j) * pow(d, (double)(k - j));
                    const double ex = call ? s - K[o] : K[o] - s;
                    if (ex > v) v = ex;
                }
                V[j] = v;
            }
        }
        price[o] = V[0];
        status[o] = 0;
    }
    free(V);
    return 0;
}

int oracle_crr_loop_fast(const double* S0, const double* K, const double* T, const double* r,
                         const double* q, const double* sigma, const unsigned char* is_call, long n,
                         int N, int american, double* price, unsigned char* status) {
    if (N <= 0) return 1;
    int err = 0;
    {
        double* V = (double*)malloc(sizeof(double) * (size_t)(N + 1));
        double* W = (double*)malloc(sizeof(double) * (size_t)(2 * N + 1)); /* W[m + N] = S0*u^m */
        if (!V || !W) err = 2;
        for (long o = 0; o < n; ++o) {
            if (err) continue;
            double u, d, disc, p;
            if (!lattice(T[o], r[o], q[o], sigma[o], N, &u, &d, &disc, &p)) {
                price[o] = NAN;
                status[o] = 1;
                continue;
            }
            const int call = is_call[o] != 0;
            const double pu = disc * p, pd = disc * (1.0 - p);
            W[N] = S0[o];
            for (int m = 1; m <= N; ++m) {
                W[N + m] = W[N + m - 1] * u;
                W[N - m] = W[N - m + 1] * d;
            }
            for (int j = 0; j <= N; ++j) {
                const double st = W[N + 2 * j - N];
                const double pay = call ? st - K[o] : K[o] - st;
                V[j] = pay > 0.0 ? pay : 0.0;
            }
            const double sgn = call ? 1.0 : -1.0, sK = call ? -K[o] : K[o];
            for (int k = N - 1; k >= 0; --k) {
                const double* w = W + N - k; /* w[2j] = S0*u^(2j-k) */
                if (american) {
                    for (int j = 0; j <= k; ++j) {
                        const double v = pu * V[j + 1] + pd * V[j];
                        const double ex = sgn * w[2 * j] + sK;
                        V[j] = ex > v ? ex : v;
                    }
                } else {
                    for (int j = 0; j <= k; ++j) V[j] = pu * V[j + 1] + pd * V[j];
                }
            }
            price[o] = V[0];
            status[o] = 0;
        }
        free(V);
        free(W);
    }
    return err;
}
