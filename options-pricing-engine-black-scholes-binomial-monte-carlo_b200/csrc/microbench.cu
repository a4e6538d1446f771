// Roofline denominators MEASURED_PEAKS.json does not carry: the FP32 FFMA issue rate, the MUFU
// rate, the FP64 DFMA rate, shared-memory LDS.64 bandwidth and the IMAD.WIDE rate (the Philox
// multiplier).  Each kernel keeps 8 independent dependency chains per thread, 1024 threads per
// CTA and two CTAs per SM, so the measured figure is the pipe's issue limit, not its latency.
// bench.py calls these in the same run as the timed kernels so the clock state is comparable.
#include "optp_common.cuh"

namespace optp {

constexpr int kIters = 4096;
constexpr int kChains = 8;

template <int WHICH>
__global__ void __launch_bounds__(1024, 2)
peak_kernel(float* sink, long long* clocks, float seed_f, double seed_d, unsigned int seed_u) {
    __shared__ double lds_buf[2048];
    const long long c0 = clock64();
    if (WHICH == 0) {  // FFMA: 256 per loop trip, so the 3 loop-control instructions cost ~1 % (32 per trip cost 9 %)
        float v[kChains];
#pragma unroll
        for (int i = 0; i < kChains; ++i) v[i] = seed_f + (float)(threadIdx.x + i);
        // ONE multiplicand and addend for all chains: they stay in the operand-reuse cache (distinct multiplicands
        // landed in the accumulators' register bank and halved the rate)
        const float m = seed_f * 0.5f + 0.499f, c = seed_f;
#pragma unroll 1
        for (int it = 0; it < kIters / 8; ++it) {
#pragma unroll
            for (int r = 0; r < 32; ++r)
#pragma unroll
                for (int i = 0; i < kChains; ++i) v[i] = fmaf(v[i], m, c);
        }
        float s = 0;
#pragma unroll
        for (int i = 0; i < kChains; ++i) s += v[i];
        if (s == 123.456f) sink[0] = s;
    } else if (WHICH == 1) {  // MUFU.EX2
        float v[kChains];
#pragma unroll
        for (int i = 0; i < kChains; ++i) v[i] = seed_f - (float)i * 0.01f;
#pragma unroll 1
        for (int it = 0; it < kIters; ++it) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < kChains; ++i) v[i] = mufu_ex2(v[i]);
        }
        float s = 0;
#pragma unroll
        for (int i = 0; i < kChains; ++i) s += v[i];
        if (s == 123.456f) sink[0] = s;
    } else if (WHICH == 2) {  // DFMA
        double v[kChains];
#pragma unroll
        for (int i = 0; i < kChains; ++i) v[i] = seed_d + (double)(threadIdx.x + i);
        const double m = seed_d * 0.5 + 0.999, c = seed_d;
#pragma unroll 1
        for (int it = 0; it < kIters; ++it) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < kChains; ++i) v[i] = fma(v[i], m, c);
        }
        double s = 0;
#pragma unroll
        for (int i = 0; i < kChains; ++i) s += v[i];
        if (s == 123.456) sink[0] = (float)s;
    } else if (WHICH == 3) {  // LDS.64, conflict free (consecutive lanes, consecutive doubles)
        for (int i = threadIdx.x; i < 2048; i += blockDim.x) lds_buf[i] = seed_d + i;
        __syncthreads();
        double s[kChains];
#pragma unroll
        for (int i = 0; i < kChains; ++i) s[i] = 0.0;
        unsigned int idx = threadIdx.x;
#pragma unroll 1
        for (int it = 0; it < kIters; ++it) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
                for (int i = 0; i < kChains; ++i) {
                    const volatile double* p = lds_buf + ((idx + 32u * i) & 2047u);
                    s[i] += *p;
                }
                idx += seed_u;
            }
        }
        double t = 0;
#pragma unroll
        for (int i = 0; i < kChains; ++i) t += s[i];
        if (t == 123.456) sink[0] = (float)t;
    } else if (WHICH == 18 || WHICH == 19) {
        // CRR node-update mix, 4 independent chains: DFMA, DFMA, DSETP, 2 FSEL (18, American) or one
        // DFMA (19, European); counted as one op per node update (the caller halves kChains)
        constexpr int NC = 4;
        double v[NC], U[NC];
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            v[i] = seed_d + (double)(threadIdx.x + i);
            U[i] = seed_d * (1.0 + 0.01 * i);
        }
        const double rho = seed_d * 0.25 + 0.5;
        double a = seed_d * 1e-3, kk = -seed_d;
#pragma unroll 1
        for (int it = 0; it < kIters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
#pragma unroll
                for (int i = 0; i < NC; ++i) {
                    double c = fma(rho, v[(i + 1) % NC], v[i]);
                    if (WHICH == 18) {
                        const double e = fma(a, U[i], kk);
                        c = e > c ? e : c;
                    }
                    v[i] = c;
                }
                if (WHICH == 18) a = -a;  // keeps the exercise DFMA inside the loop (one FADD-class op per 4 nodes)
            }
        }
        double t = 0;
#pragma unroll
        for (int i = 0; i < NC; ++i) t += v[i];
        if (t == 123.456) sink[0] = (float)t;
    } else if (WHICH >= 5) {  // single-instruction probes (inline PTX pins the opcode)
        unsigned int v[kChains];
        float f[kChains];
#pragma unroll
        for (int i = 0; i < kChains; ++i) {
            v[i] = seed_u + threadIdx.x + i;
            f[i] = seed_f * 0.5f + (float)i * 0.01f;
        }
#pragma unroll 1
        for (int it = 0; it < kIters; ++it) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < kChains; ++i) {
                    if (WHICH == 5) {  // IMAD.WIDE only: hi word feeds the next multiply
                        unsigned int lo, hi;
                        asm volatile("{\n\t.reg .b64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}"
                                     : "=r"(lo), "=r"(hi) : "r"(v[i]), "r"(0xD2511F53u));
                        v[i] = hi;
                    } else if (WHICH == 6) {  // IMAD.HI
                        asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(v[i]) : "r"(v[i]), "r"(0xD2511F53u));
                    } else if (WHICH == 7) {  // IMAD (low 32 bits)
                        asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(v[i]) : "r"(v[i]), "r"(0xD2511F53u), "r"(seed_u));
                    } else if (WHICH == 8) {  // LOP3
                        asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(v[i]) : "r"(v[i]), "r"(seed_u), "r"(it));
                    } else if (WHICH == 9) {  // I2FP
                        asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f[i]) : "r"(v[i]));
                        v[i] = __float_as_uint(f[i]);
                    } else if (WHICH == 10) {  // MUFU.SIN (+ the range-reduction FMUL)
                        asm volatile("sin.approx.ftz.f32 %0, %1;" : "=f"(f[i]) : "f"(f[i]));
                    } else if (WHICH == 11) {  // MUFU.LG2
                        asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(f[i]) : "f"(f[i]));
                    } else if (WHICH == 12) {  // MUFU.SQRT
                        asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(f[i]) : "f"(f[i]));
                    } else if (WHICH == 13) {  // FMNMX
                        asm volatile("max.f32 %0, %1, %2;" : "=f"(f[i]) : "f"(f[i]), "f"(seed_f));
                    } else if (WHICH == 14) {  // FADD
                        asm volatile("add.f32 %0, %1, %2;" : "=f"(f[i]) : "f"(f[i]), "f"(seed_f));
                    } else if (WHICH == 15 || WHICH == 16 || WHICH == 17) {
                        // does IMAD.WIDE hold the issue port?  1 IMAD.WIDE + 3 FFMA (15), + 3 FMNMX (16),
                        // + 1 MUFU.EX2 (17) per group; counted as one op per group
                        unsigned int lo, hi;
                        asm volatile("{\n\t.reg .b64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}"
                                     : "=r"(lo), "=r"(hi) : "r"(v[i]), "r"(0xD2511F53u));
                        v[i] = hi;
                        if (WHICH == 15) {
                            asm volatile("fma.rn.f32 %0, %0, %1, %2;\n\tfma.rn.f32 %0, %0, %1, %2;\n\tfma.rn.f32 %0, %0, %1, %2;"
                                         : "+f"(f[i]) : "f"(seed_f), "f"(0.25f));
                        } else if (WHICH == 16) {
                            asm volatile("max.f32 %0, %0, %1;\n\tmin.f32 %0, %0, %2;\n\tmax.f32 %0, %0, %1;"
                                         : "+f"(f[i]) : "f"(seed_f), "f"(0.25f));
                        } else {
                            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
                        }
                    }
                }
        }
        unsigned int sv = 0;
        float sf = 0;
#pragma unroll
        for (int i = 0; i < kChains; ++i) {
            sv ^= v[i];
            sf += f[i];
        }
        if (sv == 0x12345u && sf == 1.5f) sink[0] = sf;
    } else {  // IMAD.WIDE (32x32 -> 64), the Philox round multiplier
        unsigned int v[kChains];
#pragma unroll
        for (int i = 0; i < kChains; ++i) v[i] = seed_u + threadIdx.x + i;
#pragma unroll 1
        for (int it = 0; it < kIters; ++it) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < kChains; ++i) {
                    const unsigned long long p = (unsigned long long)v[i] * 0xD2511F53u;
                    v[i] = (unsigned int)(p >> 32) ^ (unsigned int)p;
                }
        }
        unsigned int s = 0;
#pragma unroll
        for (int i = 0; i < kChains; ++i) s ^= v[i];
        if (s == 0x12345u) sink[0] = (float)s;
    }
    const long long c1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) clocks[0] = c1 - c0;
}

template <int WHICH>
static int run_peak(double* per_second, double* mhz) {
    Context& c = ctx();
    OPTP_CUDA(cudaSetDevice(c.device));
    if (int e = arena_reserve(c.out1, 256)) return e;
    float* sink = (float*)c.out1.p;
    long long* clocks = (long long*)((char*)c.out1.p + 64);
    const int grid = c.n_sm * 2;
    cudaEvent_t e0, e1;
    OPTP_CUDA(cudaEventCreate(&e0));
    OPTP_CUDA(cudaEventCreate(&e1));
    float best_ms = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {  // rep 0 warms up
        OPTP_CUDA(cudaEventRecord(e0, c.stream));
        peak_kernel<WHICH><<<grid, 1024, 0, c.stream>>>(sink, clocks, 1.0f, 1.0, 1u);
        c.launches++;
        OPTP_CUDA(cudaEventRecord(e1, c.stream));
        OPTP_CUDA(cudaEventSynchronize(e1));
        OPTP_CUDA(cudaGetLastError());
        float ms = 0;
        OPTP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best_ms) best_ms = ms;
    }
    long long cyc = 0;
    OPTP_CUDA(cudaMemcpy(&cyc, clocks, sizeof(cyc), cudaMemcpyDeviceToHost));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double ops = (double)grid * 1024.0 * (double)kIters * 4.0 * kChains;
    const double scale = (WHICH == 3) ? 8.0 : 1.0;  // bytes per LDS.64 lane-op
    *per_second = ops * scale / (best_ms * 1e-3);
    // cycles one CTA spent inside the kernel (NOT a clock estimate: the grid runs in several waves)
    if (mhz) *mhz = (double)cyc;
    return 0;
}

}  // namespace optp

using namespace optp;

extern "C" int optp_microbench(int which, double* per_second_out, double* cta_cycles_out) {
    OPTP_ENTER();
    if (!per_second_out) return set_error(OPTP_E_ARG, "null output pointer");
    switch (which) {
        case 0: return run_peak<0>(per_second_out, cta_cycles_out);
        case 1: return run_peak<1>(per_second_out, cta_cycles_out);
        case 2: return run_peak<2>(per_second_out, cta_cycles_out);
        case 3: return run_peak<3>(per_second_out, cta_cycles_out);
        case 4: return run_peak<4>(per_second_out, cta_cycles_out);
        case 5: return run_peak<5>(per_second_out, cta_cycles_out);
        case 6: return run_peak<6>(per_second_out, cta_cycles_out);
        case 7: return run_peak<7>(per_second_out, cta_cycles_out);
        case 8: return run_peak<8>(per_second_out, cta_cycles_out);
        case 9: return run_peak<9>(per_second_out, cta_cycles_out);
        case 10: return run_peak<10>(per_second_out, cta_cycles_out);
        case 11: return run_peak<11>(per_second_out, cta_cycles_out);
        case 12: return run_peak<12>(per_second_out, cta_cycles_out);
        case 13: return run_peak<13>(per_second_out, cta_cycles_out);
        case 14: return run_peak<14>(per_second_out, cta_cycles_out);
        case 15: return run_peak<15>(per_second_out, cta_cycles_out);
        case 16: return run_peak<16>(per_second_out, cta_cycles_out);
        case 17: return run_peak<17>(per_second_out, cta_cycles_out);
        case 18: return run_peak<18>(per_second_out, cta_cycles_out);
        case 19: return run_peak<19>(per_second_out, cta_cycles_out);
        default: return set_error(OPTP_E_ARG, "unknown microbenchmark %d", which);
    }
}
