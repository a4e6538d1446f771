// Shared device/host helpers of the B200 pricing library: context, Philox4x32-10,
// Box-Muller normals built in registers, fp64 block reduction with a deterministic
// last-block finalise.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/optp_b200.h"

namespace optp {

// ------------------------------------------------------------------ host context
struct Arena {          // grow-only device scratch
    void* p = nullptr;
    size_t cap = 0;
};

// A sigma(S,t) table staged in HBM: the fp64 values, the (value, slope) fp32 pairs and the window nodes.
struct SigmaTableDev {
    Arena f64, f32, f4;
    uint32_t steps = 0, nodes = 0;
    bool in_clip = false;  // every node lies in [1e-8, 10] (processes.py:361-362): no per-step clip needed
    bool used = false;     // slots 1.. are handed out by optp_sigma_table_create
};
constexpr int kMaxTables = 64;

struct Context {
    int device = -1;
    bool ready = false;
    int n_sm = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;  // host->device copies of the sliced CRR batch pipeline (created on demand)
    cudaStream_t stream3 = nullptr;  // its device->host copies
    cudaStream_t stream4 = nullptr;  // the kernels of its odd slices (even slices: `stream`)
    cudaStream_t book_stream[3] = {};  // optp_mc_book deals its launches over `stream` and these (created on demand)
    cudaEvent_t book_event[3] = {};
    cudaEvent_t events[384] = {};    // its per-slice events (copy-in done, kernels done, copy-out done)
    int n_events = 0;
    double* partials = nullptr;   // [kReduceSlots][kMaxBlocks][kMaxVals]
    unsigned int* ticket = nullptr;  // [kReduceSlots]
    unsigned int reduce_slot = 0;    // next slot of the reduction scratch (reduction_scratch())
    double* stats_dev = nullptr;  // kMaxPayoffs * OPTP_NSTATS
    double* stats_host = nullptr; // pinned mirror
    Arena in0, in1, out0, out1;
    // slot 0: the table of optp_sigma_table_upload / optp_dupire_table / optp_mc_path; slots 1..: handles
    SigmaTableDev tables[kMaxTables];
    Arena work;  // work lists of the two-zone CRR kernel (crr_batch.cu)
    Arena cta_scratch;  // level arrays of the one-option-per-CTA CRR kernel when N > 9685
    Arena pin_in, pin_out;  // PINNED host staging of small CRR batches (one copy in, one copy out)
    uint64_t launches = 0;
    bool lv_direct = false;      // "lv_direct" = 1: window nodes gathered from global/L1, no smem staging
    bool lv_no_window = false;   // "lv_window" = 0: three-lookup (value, slope) path for LV Milstein
    bool crr_band_only = false;  // OPTP_CRR_BAND_ONLY=1: skip the fallback pass (diagnostics)
    bool crr_no_band = false;  // OPTP_CRR_NO_BAND=1: full-width CRR kernels only (A/B testing)
    bool crr_no_euro_sum = false;  // "crr_euro_sum" = 0: European trees by backward induction for every N
    int crr_band_kernel = 2;   // 2 = two-zone banded kernel (default), 1 = single-zone banded kernel
    int crr_zone_group = 0;    // two-zone kernel: options per warp (1, 2, 4); 0 = by batch size
    bool crr_count_ops = false;  // "crr_count_ops": the two-zone kernel counts its executed level-loop work
    unsigned long long* crr_op_count = nullptr;  // device {FP64 lane-instructions, node updates}
    int rng_mode = 0;          // "rng_mode": generator mode of the fp32 GBM / Heston FTE path kernels (RNG_*), default Philox4x32-10 x 4 normals
};

constexpr int kMaxEvents = 384;
constexpr int kMaxPayoffs = 16;
constexpr int kMaxBlocks = 148 * 16;
constexpr int kMaxVals = 3 * kMaxPayoffs + 4;
// Monte Carlo launches take the slots of the reduction scratch (block partials + ticket) round-robin, so up to
// kReduceSlots launches may be in flight on DIFFERENT streams at once; launches on one stream are ordered anyway.
constexpr int kReduceSlots = 8;

Context& ctx();
int set_error(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
int arena_reserve(Arena& a, size_t bytes);
int host_arena_reserve(Arena& a, size_t bytes);  // pinned host memory (cudaHostAlloc), grow-only
int require_ready();
// the block-partial buffer and ticket of the next Monte Carlo launch (callers hold the API lock)
void reduction_scratch(double** partials, unsigned int** ticket);

// The library keeps ONE context (stream, reduction scratch, staging arenas): every C-ABI entry point
// holds this lock for its whole duration, so calls from several host threads are serialised instead
// of racing on the scratch (ctypes releases the GIL around a call).
struct ApiLock {
    ApiLock();
    ~ApiLock();
    ApiLock(const ApiLock&) = delete;
    ApiLock& operator=(const ApiLock&) = delete;
};
#define OPTP_ENTER()                                   \
    optp::ApiLock api_lock__;                          \
    do {                                               \
        if (int e__ = optp::require_ready()) return e__; \
    } while (0)

#define OPTP_CUDA(call)                                           \
    do {                                                          \
        cudaError_t e__ = (call);                                 \
        if (e__ != cudaSuccess) return optp::cuda_fail(e__, #call); \
    } while (0)

// ------------------------------------------------------------------ Philox4x32-10
// Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC11).  Counter-based:
// the 128-bit counter is (base path index lo, hi, step block, stream), the 64-bit key is
// the user's seed, so a draw depends only on (seed, path, step) -- never on the grid
// shape or on how paths are sharded over GPUs.
// 32x32 -> 64 multiply as ONE IMAD.WIDE.U32 (writing the product as a 64-bit C++ expression
// makes nvcc add a dead "hi + 0" VIADD per multiply inside unrolled loops).
__host__ __device__ __forceinline__ void mulhilo32(uint32_t a, uint32_t b, uint32_t& hi,
                                                  uint32_t& lo) {
#ifdef __CUDA_ARCH__
    asm("{\n\t.reg .b64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}"
        : "=r"(lo), "=r"(hi)
        : "r"(a), "r"(b));
#else
    const uint64_t p = (uint64_t)a * b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
#endif
}

__host__ __device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2,
                                                      uint32_t& c3, uint32_t k0, uint32_t k1) {
    uint32_t h0, l0, h1, l1;
    mulhilo32(0xD2511F53u, c0, h0, l0);
    mulhilo32(0xCD9E8D57u, c2, h1, l1);
    c0 = h1 ^ c1 ^ k0;
    c2 = h0 ^ c3 ^ k1;
    c1 = l1;
    c3 = l0;
}

template <int ROUNDS>
__host__ __device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                     uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < ROUNDS; ++i) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                        uint32_t c3, uint32_t k0, uint32_t k1) {
    return philox4x32<10>(c0, c1, c2, c3, k0, k1);
}

// Opt-in generator modes of the fused GBM and Heston FTE kernels (optp_set_option("rng_mode", m)); the default is 0 and
// every parity statement in DESIGN.md is made for it.
//   0  Philox4x32-10, two Box-Muller pairs of (32-bit radius, 32-bit angle) per call  -> 4 normals
//   1  Philox4x32-7 (the smallest round count Salmon et al. report as Crush-resistant), same mapping
//   2  Philox4x32-10, THREE Box-Muller pairs per call: radii from the top 26 bits of words 0..2, angles
//      from the two halves of word 3 (16 bits each) and from the low 6 bits of words 0..2 (18 bits)
//      -> 6 normals; tails reach 6.1 sigma instead of 6.7, 65 536 / 262 144 directions per pair
enum : int { RNG_PHILOX10 = 0, RNG_PHILOX7 = 1, RNG_PHILOX10_X6 = 2, RNG_MODES = 3 };
template <int RNG>
struct RngTraits {
    static constexpr int rounds = (RNG == RNG_PHILOX7) ? 7 : 10;
    static constexpr int normals = (RNG == RNG_PHILOX10_X6) ? 6 : 4;
};

#ifdef __CUDACC__
// ------------------------------------------------------------------ uniforms / normals
// One 32-bit word -> u in (0,1]:  u = (w + 0.5) * 2^-32 as I2FP + FFMA (two issue slots; all 32
// bits count at the low end, so the radius sqrt(-2 ln u) reaches 6.7 sigma; w = 2^32-1 rounds
// to u = 1 -> radius 0, never a NaN).
__device__ __forceinline__ float u01_open_low(uint32_t w) {
    return fmaf(__uint2float_rn(w), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
}

__device__ __forceinline__ float mufu_lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float mufu_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float mufu_sqrt(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float mufu_sin(float x) {
    float y;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float mufu_cos(float x) {
    float y;
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Box-Muller on two words: radius from w0, angle from w1.  4 MUFU per pair.
__device__ __forceinline__ void box_muller(uint32_t w0, uint32_t w1, float& z0, float& z1) {
    const float u = u01_open_low(w0);
    // r = sqrt(-2 ln u) = sqrt(-2 ln2 * lg2 u)
    const float r = mufu_sqrt(-1.3862943611198906f * mufu_lg2(u));
    // angle = 2*pi*w1/2^32 in [0, 2*pi]
    const float a = __uint2float_rn(w1) * 1.4629180792671596e-9f;
    z0 = r * mufu_cos(a);
    z1 = r * mufu_sin(a);
}

// Same pair scaled by s >= 0 at no cost: the caller passes c = -2 ln2 * s^2, so the radius comes out as
// s * sqrt(-2 ln u) (z0 = s*Z0, z1 = s*Z1 with the SAME Z0, Z1 as box_muller on these words).
__device__ __forceinline__ void box_muller_scaled(uint32_t w0, uint32_t w1, float c, float& z0, float& z1) {
    const float u = u01_open_low(w0);
    const float r = mufu_sqrt(c * mufu_lg2(u));
    const float a = __uint2float_rn(w1) * 1.4629180792671596e-9f;
    z0 = r * mufu_cos(a);
    z1 = r * mufu_sin(a);
}

// Four random words of (seed, base path, block, stream), and the four normals they map to.
// The two halves are separate so a kernel can request the words of block c+1 before it turns
// the words of block c into normals: the integer Philox chain (IMAD.WIDE / LOP3) then overlaps
// the MUFU-bound Box-Muller of the previous block inside one warp's instruction stream.
__device__ __forceinline__ uint4 philox_words(uint64_t seed, uint64_t path, uint32_t block,
                                              uint32_t stream) {
    return philox4x32_10((uint32_t)path, (uint32_t)(path >> 32), block, stream, (uint32_t)seed,
                         (uint32_t)(seed >> 32));
}

__device__ __forceinline__ void normals_from_words(const uint4& w, float z[4]) {
    box_muller(w.x, w.y, z[0], z[1]);
    box_muller(w.z, w.w, z[2], z[3]);
}

__device__ __forceinline__ void normals_from_words_scaled(const uint4& w, float c, float z[4]) {
    box_muller_scaled(w.x, w.y, c, z[0], z[1]);
    box_muller_scaled(w.z, w.w, c, z[2], z[3]);
}

__device__ __forceinline__ void philox_normals4(uint64_t seed, uint64_t path, uint32_t block,
                                                uint32_t stream, float z[4]) {
    normals_from_words(philox_words(seed, path, block, stream), z);
}

// The same two halves for the generator modes above: words of one call, and the 4 (or 6) normals they map
// to, scaled by s when the caller passes c = -2 ln2 * s^2 (c = -2 ln2 for plain N(0,1)).
template <int RNG>
__device__ __forceinline__ uint4 philox_words_m(uint64_t seed, uint64_t path, uint32_t block, uint32_t stream) {
    return philox4x32<RngTraits<RNG>::rounds>((uint32_t)path, (uint32_t)(path >> 32), block, stream,
                                              (uint32_t)seed, (uint32_t)(seed >> 32));
}

__device__ __forceinline__ void bm_pair(float u, float a, float c, float& z0, float& z1) {
    const float r = mufu_sqrt(c * mufu_lg2(u));
    z0 = r * mufu_cos(a);
    z1 = r * mufu_sin(a);
}

template <int RNG>
__device__ __forceinline__ void normals_from_words_m(const uint4& w, float c, float* z) {
    if (RNG == RNG_PHILOX10_X6) {
        // radius: u = ((w >> 6) + 1/2) 2^-26 in (0, 1); angle: (k + 1/2) 2 pi / 2^bits
        const float ux = fmaf(__uint2float_rn(w.x & 0xFFFFFFC0u), 2.3283064365386963e-10f, 7.450580596923828e-9f);
        const float uy = fmaf(__uint2float_rn(w.y & 0xFFFFFFC0u), 2.3283064365386963e-10f, 7.450580596923828e-9f);
        const float uz = fmaf(__uint2float_rn(w.z & 0xFFFFFFC0u), 2.3283064365386963e-10f, 7.450580596923828e-9f);
        const float a0 = fmaf(__uint2float_rn(w.w >> 16), 9.587379924285257e-5f, 4.7936899621426287e-5f);
        const float a1 = fmaf(__uint2float_rn(w.w & 0xFFFFu), 9.587379924285257e-5f, 4.7936899621426287e-5f);
        const uint32_t k2 = ((w.x & 63u) << 12) | ((w.y & 63u) << 6) | (w.z & 63u);
        const float a2 = fmaf(__uint2float_rn(k2), 2.3968449810713143e-5f, 1.1984224905356572e-5f);
        bm_pair(ux, a0, c, z[0], z[1]);
        bm_pair(uy, a1, c, z[2], z[3]);
        bm_pair(uz, a2, c, z[4], z[5]);
    } else {
        box_muller_scaled(w.x, w.y, c, z[0], z[1]);
        box_muller_scaled(w.z, w.w, c, z[2], z[3]);
    }
}

// Poisson(mu) count by CDF inversion from one 32-bit word (mu = lambda*dt is small: the loop
// almost always stops at k = 0).  p0 = exp(-mu).
__device__ __forceinline__ float poisson_inv(uint32_t w, float mu, float p0) {
    const float u = fmaf(__uint2float_rn(w), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    float k = 0.f, p = p0, c = p0;
    while (u > c && k < 64.f) {
        k += 1.f;
        p *= mu / k;
        c += p;
    }
    return k;
}

// ------------------------------------------------------------------ fp64 reductions
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum `v` over the block; result valid in thread 0.  `smem` needs 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* smem) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) smem[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    double t = 0.0;
    if (wid == 0) {
        t = lane < nw ? smem[lane] : 0.0;
        t = warp_sum(t);
    }
    return t;
}

// Each block deposits `n_vals` partial sums; the last block to arrive adds them in block
// order (deterministic for a given grid) and writes totals through `emit`.
template <class Emit>
__device__ __forceinline__ void finalize_partials(double* partials, unsigned int* ticket,
                                                  const double* block_vals_smem, int n_vals,
                                                  Emit emit) {
    __shared__ bool is_last;
    for (int i = threadIdx.x; i < n_vals; i += blockDim.x)
        partials[(size_t)blockIdx.x * kMaxVals + i] = block_vals_smem[i];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int i = threadIdx.x; i < n_vals; i += blockDim.x) {
        double s = 0.0;
        for (unsigned int b = 0; b < gridDim.x; ++b)
            s += __ldcg(&partials[(size_t)b * kMaxVals + i]);
        emit(i, s);
    }
    if (threadIdx.x == 0) *ticket = 0u;
}
#endif  // __CUDACC__

}  // namespace optp
