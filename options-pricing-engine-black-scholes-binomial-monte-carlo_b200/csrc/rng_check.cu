// Generator self-checks behind the C ABI (test infrastructure of the PRODUCT's own generator, not a
// pricing path): a host-side Philox4x32-10 known-answer entry point, and a device kernel that draws
// normals through exactly the code the fused path kernels use (philox_words_m + normals_from_words_m,
// counter layout (path, block, stream 0)) and reduces their moments, tail counts and pair products.
#include <math.h>
#include "mc_kernels.cuh"

namespace optp {

constexpr int kNormVals = 15;  // sums below; the maximum travels through an atomicMax

// out[0] n  [1] sum z  [2] sum z^2  [3] sum z^3  [4] sum z^4
//    [5..8] counts of |z| > 3, 4, 5, 6
//    [9]  sum z_a z_b over the two normals of a Box-Muller pair (r cos, r sin)     [13] number of such pairs
//    [10] sum z_a z_b over first normals of different pairs of one call             [14] number of such products
//    [11] sum (last normal of call c) * (first normal of call c+1) along a path     (n_paths * (n_blocks-1) products)
//    [12] max |z|
template <int RNG>
__global__ void __launch_bounds__(256) normal_stats_kernel(unsigned long long seed, unsigned long long path0,
                                                           unsigned long long n_paths, unsigned int n_blocks,
                                                           double* partials, unsigned int* ticket,
                                                           unsigned long long* max_bits, double* out) {
    constexpr int NZ = RngTraits<RNG>::normals;
    __shared__ double red[32];
    __shared__ double blockvals[kNormVals];
    double s1 = 0, s2 = 0, s3 = 0, s4 = 0, c3 = 0, c4 = 0, c5 = 0, c6 = 0, pp = 0, px = 0, pl = 0;
    float zmax = 0.f;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < n_paths; b += stride) {
        float prev = 0.f;
        for (unsigned int c = 0; c < n_blocks; ++c) {
            float z[NZ];
            const uint4 w = philox_words_m<RNG>(seed, path0 + b, c, 0u);
            normals_from_words_m<RNG>(w, -1.3862943611198906f, z);
            // fp32 partial sums over one call, folded into fp64 per call
            float t1 = 0.f, t2 = 0.f, t3 = 0.f, t4 = 0.f;
            int n3 = 0, n4 = 0, n5 = 0, n6 = 0;
#pragma unroll
            for (int j = 0; j < NZ; ++j) {
                const float x = z[j], x2 = x * x, ax = fabsf(x);
                t1 += x;
                t2 += x2;
                t3 = fmaf(x2, x, t3);
                t4 = fmaf(x2, x2, t4);
                n3 += ax > 3.f;
                n4 += ax > 4.f;
                n5 += ax > 5.f;
                n6 += ax > 6.f;
                zmax = fmaxf(zmax, ax);
            }
            s1 += (double)t1;
            s2 += (double)t2;
            s3 += (double)t3;
            s4 += (double)t4;
            c3 += (double)n3;
            c4 += (double)n4;
            c5 += (double)n5;
            c6 += (double)n6;
            float tp = 0.f, tx = 0.f;
#pragma unroll
            for (int j = 0; j < NZ; j += 2) tp = fmaf(z[j], z[j + 1], tp);
#pragma unroll
            for (int j = 0; j + 2 < NZ; j += 2) tx = fmaf(z[j], z[j + 2], tx);
            pp += (double)tp;
            px += (double)tx;
            if (c > 0) pl += (double)(prev * z[0]);
            prev = z[NZ - 1];
        }
    }
    const double n_calls = (double)n_paths * (double)n_blocks;
    double vals[kNormVals] = {0.0, s1, s2, s3, s4, c3, c4, c5, c6, pp, px, pl, 0.0, 0.0, 0.0};
    for (int i = 1; i < 12; ++i) {
        const double t = block_sum(vals[i], red);
        if (threadIdx.x == 0) blockvals[i] = t;
    }
    if (threadIdx.x == 0) blockvals[0] = blockvals[12] = blockvals[13] = blockvals[14] = 0.0;
    // non-negative floats order like their bit patterns
    float m = zmax;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(max_bits, (unsigned long long)__float_as_uint(m));
    __syncthreads();
    finalize_partials(partials, ticket, blockvals, kNormVals, [=](int i, double s) {
        if (i == 0) out[0] = n_calls * NZ;
        else if (i == 12) out[12] = (double)__uint_as_float((unsigned int)__ldcg(max_bits));
        else if (i == 13) out[13] = n_calls * (NZ / 2);
        else if (i == 14) out[14] = n_calls * (NZ / 2 - 1);
        else out[i] = s;
    });
}

}  // namespace optp

using namespace optp;

extern "C" {

// Host-only: the library's Philox4x32-10 (the very function the kernels inline) on one counter/key pair,
// for known-answer tests against the Random123 vectors.  Needs no device and no optp_init.
int optp_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    if (!ctr || !key || !out) return set_error(OPTP_E_ARG, "null pointer argument");
    const uint4 r = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    out[0] = r.x;
    out[1] = r.y;
    out[2] = r.z;
    out[3] = r.w;
    return 0;
}

int optp_normal_stats(uint64_t seed, uint64_t path_offset, uint64_t n_paths, uint32_t n_blocks, int rng_mode,
                      double stats_out[16]) {
    OPTP_ENTER();
    if (!stats_out || n_paths == 0 || n_blocks == 0) return set_error(OPTP_E_ARG, "empty draw request");
    if (rng_mode < 0 || rng_mode >= RNG_MODES) return set_error(OPTP_E_ARG, "rng_mode must be 0, 1 or 2");
    Context& c = ctx();
    OPTP_CUDA(cudaSetDevice(c.device));
    if (int e = arena_reserve(c.out1, 32 * sizeof(double))) return e;
    double* out = (double*)c.out1.p;
    unsigned long long* max_bits = (unsigned long long*)(out + 16);
    OPTP_CUDA(cudaMemsetAsync(out, 0, 32 * sizeof(double), c.stream));
    unsigned long long blocks = (n_paths + 255) / 256;
    const unsigned long long cap = (unsigned long long)c.n_sm * 8;
    if (blocks > cap) blocks = cap;
    const unsigned int grid = (unsigned int)blocks;
    double* partials;
    unsigned int* ticket;
    reduction_scratch(&partials, &ticket);
    if (rng_mode == RNG_PHILOX7)
        normal_stats_kernel<RNG_PHILOX7><<<grid, 256, 0, c.stream>>>(seed, path_offset, n_paths, n_blocks, partials,
                                                                     ticket, max_bits, out);
    else if (rng_mode == RNG_PHILOX10_X6)
        normal_stats_kernel<RNG_PHILOX10_X6><<<grid, 256, 0, c.stream>>>(seed, path_offset, n_paths, n_blocks,
                                                                         partials, ticket, max_bits, out);
    else
        normal_stats_kernel<RNG_PHILOX10><<<grid, 256, 0, c.stream>>>(seed, path_offset, n_paths, n_blocks, partials,
                                                                      ticket, max_bits, out);
    c.launches++;
    OPTP_CUDA(cudaGetLastError());
    OPTP_CUDA(cudaMemcpyAsync(c.stats_host, out, 16 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    OPTP_CUDA(cudaStreamSynchronize(c.stream));
    for (int i = 0; i < 16; ++i) stats_out[i] = c.stats_host[i];
    stats_out[15] = (double)n_paths * (double)(n_blocks - 1);
    return 0;
}

}  // extern "C"
