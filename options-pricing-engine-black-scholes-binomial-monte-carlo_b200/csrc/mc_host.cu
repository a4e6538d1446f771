// C-ABI front end of the Monte Carlo path (include/optp_b200.h): argument checks, staging of
// host buffers into the per-device arenas, kernel dispatch, result read-back.
#include <string.h>
#include "mc_kernels.cuh"

namespace optp {
int launch_payoff_from_paths(const double* paths, unsigned long long n_rows,
                             unsigned long long n_cols, const optp_payoff& pay, double* payoff_out,
                             double* stats, cudaStream_t st);

static int check_payoffs(const optp_payoff* p, uint32_t n) {
    if (!p || n == 0 || n > (uint32_t)kMaxPayoffs)
        return set_error(OPTP_E_ARG, "n_payoffs must be in [1, %d]", kMaxPayoffs);
    for (uint32_t i = 0; i < n; ++i) {
        if (p[i].kind < OPTP_PAY_VANILLA || p[i].kind > OPTP_PAY_LOOKBACK)
            return set_error(OPTP_E_KIND, "unknown payoff kind %d", p[i].kind);
        if (p[i].r_disc != p[0].r_disc || p[i].T_disc != p[0].T_disc)
            return set_error(OPTP_E_ARG, "payoffs of one call must share r_disc and T_disc");
    }
    return 0;
}

static int check_process(const optp_process* pr, uint32_t n_steps, uint64_t n_base) {
    if (!pr) return set_error(OPTP_E_ARG, "null process");
    if (n_steps == 0 || n_base == 0)
        return set_error(OPTP_E_ARG, "n_steps and n_paths must be positive.");
    if (pr->kind < OPTP_PROC_GBM || pr->kind > OPTP_PROC_HESTON_QE_M)
        return set_error(OPTP_E_KIND, "unknown process kind %d", pr->kind);
    if ((pr->kind == OPTP_PROC_HESTON_FTE || pr->kind == OPTP_PROC_SABR ||
         pr->kind == OPTP_PROC_HESTON_QE || pr->kind == OPTP_PROC_HESTON_QE_M) &&
        !(pr->rho >= -1.0 && pr->rho <= 1.0))
        return set_error(OPTP_E_ARG, "rho must be in [-1, 1].");
    if (pr->kind == OPTP_PROC_MERTON && (pr->v0 < 0.0 || pr->theta < 0.0))
        return set_error(OPTP_E_ARG, "lam and sJ must be non-negative.");
    if ((pr->kind == OPTP_PROC_HESTON_QE || pr->kind == OPTP_PROC_HESTON_QE_M) &&
        !(pr->xi > 0.0 && pr->kappa > 0.0 && pr->theta > 0.0))
        return set_error(OPTP_E_ARG, "the QE scheme needs xi > 0, kappa > 0 and theta > 0");
    if (pr->kind == OPTP_PROC_LV_MILSTEIN || pr->kind == OPTP_PROC_LV_EULER) {
        if (pr->n_nodes < 2 || !(pr->h > 0.0))
            return set_error(OPTP_E_ARG, "sigma table needs n_nodes >= 2 and h > 0");
        if (pr->kind == OPTP_PROC_LV_MILSTEIN && !(pr->dS_bump > 0.0 && pr->dS_bump < 1.0))
            return set_error(OPTP_E_ARG, "dS_bump must be in (0, 1)");
    }
    return 0;
}

static bool is_lv(int kind) { return kind == OPTP_PROC_LV_MILSTEIN || kind == OPTP_PROC_LV_EULER; }

static void fill_args(PathArgs& a, const optp_process* pr, const optp_payoff* pay, uint32_t n_pay,
                      uint32_t n_steps, uint64_t n_base, int anti, int mode, uint64_t seed,
                      uint64_t path_offset, int table_slot = 0) {
    Context& c = ctx();
    const SigmaTableDev& tb = c.tables[table_slot];
    memset(&a, 0, sizeof(a));
    a.proc = (pr->kind == OPTP_PROC_HESTON_QE_M) ? OPTP_PROC_HESTON_QE : pr->kind;
    a.qe_martingale = (pr->kind == OPTP_PROC_HESTON_QE_M) ? 1 : 0;
    a.n_nodes = pr->n_nodes;
    a.S0 = pr->S0; a.r = pr->r; a.q = pr->q; a.T = pr->T; a.sigma = pr->sigma;
    a.v0 = pr->v0; a.kappa = pr->kappa; a.theta = pr->theta; a.xi = pr->xi; a.rho = pr->rho;
    a.bump = pr->dS_bump; a.x0 = pr->x0; a.h = pr->h;
    a.table32 = (const float2*)tb.f32.p;
    a.table4 = (const float4*)tb.f4.p;
    a.table64 = (const double*)tb.f64.p;
    a.table_in_clip = tb.in_clip ? 1 : 0;
    a.n_steps = n_steps;
    a.n_base = n_base;
    a.anti = anti ? 1 : 0;
    a.use_philox = (mode != OPTP_MODE_FP64_HOSTZ);
    a.rng_mode = c.rng_mode;
    a.seed = seed;
    a.path_offset = path_offset;
    a.n_pay = (int)n_pay;
    for (uint32_t i = 0; i < n_pay; ++i) a.pay[i] = pay[i];
    reduction_scratch(&a.partials, &a.ticket);
}

static int upload_table(const double* table, uint32_t n_steps, uint32_t n_nodes, int slot = 0) {
    Context& c = ctx();
    SigmaTableDev& tb = c.tables[slot];
    if (!table || n_steps == 0 || n_nodes < 2)
        return set_error(OPTP_E_ARG, "sigma table must be n_steps x n_nodes with n_nodes >= 2");
    const size_t n = (size_t)n_steps * n_nodes;
    const size_t n_ld = (n_nodes + 1) & ~1u;
    if (int e = arena_reserve(tb.f64, n * sizeof(double))) return e;
    if (int e = arena_reserve(tb.f32, (size_t)n_steps * n_ld * sizeof(float2))) return e;
    if (int e = arena_reserve(tb.f4, (size_t)n_steps * (n_nodes + 4) * sizeof(float4))) return e;
    OPTP_CUDA(cudaMemcpyAsync(tb.f64.p, table, n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    if (int e = launch_table_convert((const double*)tb.f64.p, (float2*)tb.f32.p, (float4*)tb.f4.p, n_steps,
                                     (int)n_nodes, c.stream))
        return e;
    tb.steps = n_steps;
    tb.nodes = n_nodes;
    bool in_clip = true;  // NaNs fail the comparison and keep the per-step clip
    for (size_t i = 0; i < n && in_clip; ++i) in_clip = (table[i] >= 1e-8) && (table[i] <= 10.0);
    tb.in_clip = in_clip;
    return 0;
}

static int check_table_slot(int slot, uint32_t n_steps, int n_nodes) {
    Context& c = ctx();
    if (slot < 0 || slot >= kMaxTables || (slot > 0 && !c.tables[slot].used))
        return set_error(OPTP_E_ARG, "sigma table handle %d is not live", slot);
    const SigmaTableDev& tb = c.tables[slot];
    if (tb.steps < n_steps || tb.nodes != (uint32_t)n_nodes || !tb.f64.p)
        return set_error(OPTP_E_ARG, "no sigma table of %u x %d staged (optp_sigma_table_upload / _create)", n_steps,
                         n_nodes);
    return 0;
}

static int dispatch_path(const PathArgs& a, int mode, cudaStream_t st) {
    if (mode == OPTP_MODE_FP32_PHILOX) {
        if (!path_fp32_supported(a.proc))
            return set_error(OPTP_E_KIND, "process kind %d has no fp32 kernel; use an fp64 mode",
                             a.proc);
        if (a.paths_out || a.var_out)
            return set_error(OPTP_E_ARG, "path materialisation needs an fp64 mode");
        return launch_path_fp32(a, st);
    }
    if (mode == OPTP_MODE_FP64_PHILOX || mode == OPTP_MODE_FP64_HOSTZ) return launch_path_fp64(a, st);
    return set_error(OPTP_E_ARG, "unknown mode %d", mode);
}

}  // namespace optp

using namespace optp;

extern "C" {

int optp_sigma_table_upload(const double* table, uint32_t n_steps, uint32_t n_nodes) {
    OPTP_ENTER();
    OPTP_CUDA(cudaSetDevice(ctx().device));
    if (int e = upload_table(table, n_steps, n_nodes)) return e;
    OPTP_CUDA(cudaStreamSynchronize(ctx().stream));
    return 0;
}

int optp_sigma_table_create(const double* table, uint32_t n_steps, uint32_t n_nodes, int32_t* handle_out) {
    OPTP_ENTER();
    if (!handle_out) return set_error(OPTP_E_ARG, "null handle pointer");
    Context& c = ctx();
    OPTP_CUDA(cudaSetDevice(c.device));
    int slot = 0;
    for (int i = 1; i < kMaxTables && !slot; ++i)
        if (!c.tables[i].used) slot = i;
    if (!slot) return set_error(OPTP_E_TOO_LARGE, "all %d sigma table handles are in use", kMaxTables - 1);
    if (int e = upload_table(table, n_steps, n_nodes, slot)) return e;
    OPTP_CUDA(cudaStreamSynchronize(c.stream));
    c.tables[slot].used = true;
    *handle_out = slot;
    return 0;
}

int optp_sigma_table_destroy(int32_t handle) {
    OPTP_ENTER();
    Context& c = ctx();
    if (handle < 1 || handle >= kMaxTables || !c.tables[handle].used)
        return set_error(OPTP_E_ARG, "sigma table handle %d is not live", (int)handle);
    OPTP_CUDA(cudaSetDevice(c.device));
    OPTP_CUDA(cudaStreamSynchronize(c.stream));
    SigmaTableDev& tb = c.tables[handle];
    for (Arena* a : {&tb.f64, &tb.f32, &tb.f4}) {
        if (a->p) cudaFree(a->p);
        a->p = nullptr;
        a->cap = 0;
    }
    tb = SigmaTableDev();
    return 0;
}

// ------------------------------------------------------------------ terminal
int optp_mc_terminal_dev(double S0, double K, double T, double r, double q, double sigma,
                         int is_call, uint64_t n_base, int antithetic, int mode, int fp32_cast,
                         uint64_t seed, uint64_t path_offset, const double* z_dev,
                         double* stats_dev, void* stream) {
    OPTP_ENTER();
    if (!stats_dev) return set_error(OPTP_E_ARG, "null stats pointer");
    if (n_base == 0) return set_error(OPTP_E_ARG, "n_base must be positive");
    if (mode == OPTP_MODE_FP64_HOSTZ && !z_dev)
        return set_error(OPTP_E_ARG, "HOSTZ mode needs the normal draws");
    Context& c = ctx();
    TermArgs a;
    a.S0 = S0; a.K = K; a.T = T; a.r = r; a.q = q; a.sigma = sigma;
    a.is_call = is_call ? 1 : 0;
    a.n_base = n_base;
    a.anti = antithetic ? 1 : 0;
    a.fp32_cast = fp32_cast ? 1 : 0;
    a.use_philox = (mode != OPTP_MODE_FP64_HOSTZ);
    a.seed = seed;
    a.path_offset = path_offset;
    a.z = z_dev;
    reduction_scratch(&a.partials, &a.ticket);
    a.stats = stats_dev;
    if (mode == OPTP_MODE_FP32_PHILOX) return launch_terminal_fp32(a, (cudaStream_t)stream);
    if (mode == OPTP_MODE_FP64_PHILOX || mode == OPTP_MODE_FP64_HOSTZ)
        return launch_terminal_fp64(a, (cudaStream_t)stream);
    return set_error(OPTP_E_ARG, "unknown mode %d", mode);
}

int optp_mc_terminal(double S0, double K, double T, double r, double q, double sigma, int is_call,
                     uint64_t n_base, int antithetic, int mode, int fp32_cast, uint64_t seed,
                     uint64_t path_offset, const double* z, double stats_out[OPTP_NSTATS]) {
    OPTP_ENTER();
    if (!stats_out) return set_error(OPTP_E_ARG, "null stats pointer");
    Context& c = ctx();
    OPTP_CUDA(cudaSetDevice(c.device));
    if (n_base == 0) {  // monte_carlo.py:23-24
        memset(stats_out, 0, sizeof(double) * OPTP_NSTATS);
        return 0;
    }
    const double* z_dev = nullptr;
    if (mode == OPTP_MODE_FP64_HOSTZ) {
        if (!z) return set_error(OPTP_E_ARG, "HOSTZ mode needs the normal draws");
        if (int e = arena_reserve(c.in0, n_base * sizeof(double))) return e;
        OPTP_CUDA(cudaMemcpyAsync(c.in0.p, z, n_base * sizeof(double), cudaMemcpyHostToDevice,
                                  c.stream));
        z_dev = (const double*)c.in0.p;
    }
    if (int e = optp_mc_terminal_dev(S0, K, T, r, q, sigma, is_call, n_base, antithetic, mode,
                                     fp32_cast, seed, path_offset, z_dev, c.stats_dev, c.stream))
        return e;
    OPTP_CUDA(cudaMemcpyAsync(c.stats_host, c.stats_dev, sizeof(double) * OPTP_NSTATS,
                              cudaMemcpyDeviceToHost, c.stream));
    OPTP_CUDA(cudaStreamSynchronize(c.stream));
    memcpy(stats_out, c.stats_host, sizeof(double) * OPTP_NSTATS);
    return 0;
}

// ------------------------------------------------------------------ fused paths + payoffs
static int mc_path_dev_slot(const optp_process* proc, const optp_payoff* payoffs, uint32_t n_payoffs,
                            uint32_t n_steps, uint64_t n_base, int antithetic, int mode, uint64_t seed,
                            uint64_t path_offset, const double* z_dev, double* payoff_out_dev,
                            double* terminal_out_dev, double* stats_dev, void* stream, int table_slot) {
    if (int e = check_process(proc, n_steps, n_base)) return e;
    if (int e = check_payoffs(payoffs, n_payoffs)) return e;
    if (!stats_dev) return set_error(OPTP_E_ARG, "null stats pointer");
    if (mode == OPTP_MODE_FP64_HOSTZ && !z_dev)
        return set_error(OPTP_E_ARG, "HOSTZ mode needs the normal draws");
    if (is_lv(proc->kind))
        if (int e = check_table_slot(table_slot, n_steps, proc->n_nodes)) return e;
    PathArgs a;
    fill_args(a, proc, payoffs, n_payoffs, n_steps, n_base, antithetic, mode, seed, path_offset, table_slot);
    a.z = z_dev;
    a.payoff_out = payoff_out_dev;
    a.terminal_out = terminal_out_dev;
    a.stats = stats_dev;
    return dispatch_path(a, mode, (cudaStream_t)stream);
}

int optp_mc_path_dev(const optp_process* proc, const optp_payoff* payoffs, uint32_t n_payoffs,
                     uint32_t n_steps, uint64_t n_base, int antithetic, int mode, uint64_t seed,
                     uint64_t path_offset, const double* z_dev, double* payoff_out_dev,
                     double* terminal_out_dev, double* stats_dev, void* stream) {
    OPTP_ENTER();
    return mc_path_dev_slot(proc, payoffs, n_payoffs, n_steps, n_base, antithetic, mode, seed, path_offset, z_dev,
                            payoff_out_dev, terminal_out_dev, stats_dev, stream, 0);
}

int optp_mc_path(const optp_process* proc, const optp_payoff* payoffs, uint32_t n_payoffs,
                 uint32_t n_steps, uint64_t n_base, int antithetic, int mode, uint64_t seed,
                 uint64_t path_offset, const double* z, double* payoff_out, double* terminal_out,
                 double* stats_out) {
    OPTP_ENTER();
    if (int e = check_process(proc, n_steps, n_base)) return e;
    if (int e = check_payoffs(payoffs, n_payoffs)) return e;
    if (!stats_out) return set_error(OPTP_E_ARG, "null stats pointer");
    Context& c = ctx();
    OPTP_CUDA(cudaSetDevice(c.device));
    if (is_lv(proc->kind)) {
        if (!proc->sigma_table) return set_error(OPTP_E_ARG, "LV process without a sigma table");
        if (int e = upload_table(proc->sigma_table, n_steps, (uint32_t)proc->n_nodes)) return e;
    }
    const uint64_t n_eff = antithetic ? 2 * n_base : n_base;
    const double* z_dev = nullptr;
    if (mode == OPTP_MODE_FP64_HOSTZ) {
        if (!z) return set_error(OPTP_E_ARG, "HOSTZ mode needs the normal draws");
        const size_t bytes = sizeof(double) * (size_t)n_streams_of(proc->kind) * n_steps * n_base;
        if (int e = arena_reserve(c.in0, bytes)) return e;
        OPTP_CUDA(cudaMemcpyAsync(c.in0.p, z, bytes, cudaMemcpyHostToDevice, c.stream));
        z_dev = (const double*)c.in0.p;
    }
    double *pay_dev = nullptr, *term_dev = nullptr;
    if (payoff_out) {
        if (int e = arena_reserve(c.out0, sizeof(double) * n_payoffs * n_eff)) return e;
        pay_dev = (double*)c.out0.p;
    }
    if (terminal_out) {
        if (int e = arena_reserve(c.out1, sizeof(double) * n_eff)) return e;
        term_dev = (double*)c.out1.p;
    }
    if (int e = optp_mc_path_dev(proc, payoffs, n_payoffs, n_steps, n_base, antithetic, mode, seed,
                                 path_offset, z_dev, pay_dev, term_dev, c.stats_dev, c.stream))
        return e;
    const size_t sbytes = sizeof(double) * OPTP_NSTATS * n_payoffs;
    OPTP_CUDA(cudaMemcpyAsync(c.stats_host, c.stats_dev, sbytes, cudaMemcpyDeviceToHost, c.stream));
    if (payoff_out)
        OPTP_CUDA(cudaMemcpyAsync(payoff_out, pay_dev, sizeof(double) * n_payoffs * n_eff,
                                  cudaMemcpyDeviceToHost, c.stream));
    if (terminal_out)
        OPTP_CUDA(cudaMemcpyAsync(terminal_out, term_dev, sizeof(double) * n_eff,
                                  cudaMemcpyDeviceToHost, c.stream));
    OPTP_CUDA(cudaStreamSynchronize(c.stream));
    memcpy(stats_out, c.stats_host, sbytes);
    return 0;
}

// ------------------------------------------------------------------ book of items
int optp_mc_book(const optp_book_item* items, uint64_t n_items, const optp_payoff* payoffs,
                 uint64_t n_payoffs_total, double* stats_out) {
    OPTP_ENTER();
    if (n_items == 0) return 0;
    if (!items || !stats_out) return set_error(OPTP_E_ARG, "null pointer argument");
    Context& c = ctx();
    OPTP_CUDA(cudaSetDevice(c.device));
    // validate everything before the first launch, and size the statistics buffer
    uint64_t n_blocks = 0;
    for (uint64_t i = 0; i < n_items; ++i) {
        const optp_book_item& it = items[i];
        if (it.terminal_only) {
            if (it.n_base == 0) return set_error(OPTP_E_ARG, "item %llu: n_base must be positive",
                                                 (unsigned long long)i);
            if (it.mode == OPTP_MODE_FP64_HOSTZ)
                return set_error(OPTP_E_ARG, "item %llu: a book has no host normals", (unsigned long long)i);
            n_blocks += 1;
        } else {
            if (is_lv(it.proc.kind)) {  // local-vol items name a staged table (optp_sigma_table_create)
                if (it.table_handle < 1)
                    return set_error(OPTP_E_ARG, "item %llu: a local-vol item needs a table handle (optp_sigma_table_create)",
                                     (unsigned long long)i);
                if (int e = check_table_slot(it.table_handle, it.n_steps, it.proc.n_nodes)) return e;
            }
            if (it.mode == OPTP_MODE_FP64_HOSTZ)
                return set_error(OPTP_E_ARG, "item %llu: a book has no host normals", (unsigned long long)i);
            if (int e = check_process(&it.proc, it.n_steps, it.n_base)) return e;
            if (!payoffs || (uint64_t)it.first_payoff + (uint64_t)it.n_payoffs > n_payoffs_total)
                return set_error(OPTP_E_ARG, "item %llu: payoff range out of bounds", (unsigned long long)i);
            if (int e = check_payoffs(payoffs + it.first_payoff, (uint32_t)it.n_payoffs)) return e;
            n_blocks += (uint64_t)it.n_payoffs;
        }
    }
    const size_t bytes = sizeof(double) * OPTP_NSTATS * n_blocks;
    if (int e = arena_reserve(c.out1, bytes)) return e;
    double* stats_dev = (double*)c.out1.p;
    // The launches are dealt over four streams: an item of 1e5 paths does not fill the machine and its tail would
    // otherwise sit in front of the next item's head.  The reduction scratch has kReduceSlots = 8 slots taken
    // round-robin, so launches i and i + 8 share a slot -- and a stream (8 is a multiple of 4): ordered.
    constexpr int kBookStreams = 4;
    static_assert(kReduceSlots % kBookStreams == 0, "slot reuse must stay on one stream");
    const bool spread = n_items >= 2 * kBookStreams;
    cudaStream_t streams[kBookStreams] = {c.stream, c.stream, c.stream, c.stream};
    if (spread) {
        for (int i = 0; i < 3; ++i) {
            if (!c.book_stream[i]) OPTP_CUDA(cudaStreamCreateWithFlags(&c.book_stream[i], cudaStreamNonBlocking));
            if (!c.book_event[i]) OPTP_CUDA(cudaEventCreateWithFlags(&c.book_event[i], cudaEventDisableTiming));
            streams[i + 1] = c.book_stream[i];
        }
        // whatever the context's stream still has queued (table uploads) comes first
        OPTP_CUDA(cudaEventRecord(c.book_event[0], c.stream));
        for (int i = 0; i < 3; ++i) OPTP_CUDA(cudaStreamWaitEvent(c.book_stream[i], c.book_event[0], 0));
    }
    uint64_t blk = 0;
    for (uint64_t i = 0; i < n_items; ++i) {
        const optp_book_item& it = items[i];
        double* dst = stats_dev + blk * OPTP_NSTATS;
        cudaStream_t st = streams[i % kBookStreams];
        if (it.terminal_only) {
            if (int e = optp_mc_terminal_dev(it.proc.S0, it.K, it.proc.T, it.proc.r, it.proc.q,
                                             it.proc.sigma, it.is_call, it.n_base, it.antithetic, it.mode,
                                             0, it.seed, it.path_offset, nullptr, dst, st))
                return e;
            blk += 1;
        } else {
            if (int e = mc_path_dev_slot(&it.proc, payoffs + it.first_payoff, (uint32_t)it.n_payoffs, it.n_steps,
                                         it.n_base, it.antithetic, it.mode, it.seed, it.path_offset, nullptr, nullptr,
                                         nullptr, dst, st, is_lv(it.proc.kind) ? it.table_handle : 0))
                return e;
            blk += (uint64_t)it.n_payoffs;
        }
    }
    if (spread)
        for (int i = 0; i < 3; ++i) {
            OPTP_CUDA(cudaEventRecord(c.book_event[i], c.book_stream[i]));
            OPTP_CUDA(cudaStreamWaitEvent(c.stream, c.book_event[i], 0));
        }
    OPTP_CUDA(cudaMemcpyAsync(stats_out, stats_dev, bytes, cudaMemcpyDeviceToHost, c.stream));
    OPTP_CUDA(cudaStreamSynchronize(c.stream));
    return 0;
}

// ------------------------------------------------------------------ materialised paths
int optp_paths(const optp_process* proc, uint32_t n_steps, uint64_t n_base, int antithetic, int mode,
               uint64_t seed, uint64_t path_offset, const double* z, double* paths_out,
               double* var_out) {
    OPTP_ENTER();
    if (int e = check_process(proc, n_steps, n_base)) return e;
    if (!paths_out) return set_error(OPTP_E_ARG, "null paths pointer");
    if (mode != OPTP_MODE_FP64_PHILOX && mode != OPTP_MODE_FP64_HOSTZ)
        return set_error(OPTP_E_ARG, "path materialisation needs an fp64 mode");
    Context& c = ctx();
    OPTP_CUDA(cudaSetDevice(c.device));
    if (is_lv(proc->kind)) {
        if (!proc->sigma_table) return set_error(OPTP_E_ARG, "LV process without a sigma table");
        if (int e = upload_table(proc->sigma_table, n_steps, (uint32_t)proc->n_nodes)) return e;
    }
    const uint64_t n_eff = antithetic ? 2 * n_base : n_base;
    const size_t pbytes = sizeof(double) * (size_t)(n_steps + 1) * n_eff;
    const double* z_dev = nullptr;
    if (mode == OPTP_MODE_FP64_HOSTZ) {
        if (!z) return set_error(OPTP_E_ARG, "HOSTZ mode needs the normal draws");
        const size_t bytes = sizeof(double) * (size_t)n_streams_of(proc->kind) * n_steps * n_base;
        if (int e = arena_reserve(c.in0, bytes)) return e;
        OPTP_CUDA(cudaMemcpyAsync(c.in0.p, z, bytes, cudaMemcpyHostToDevice, c.stream));
        z_dev = (const double*)c.in0.p;
    }
    if (int e = arena_reserve(c.out0, pbytes)) return e;
    if (var_out)
        if (int e = arena_reserve(c.out1, pbytes)) return e;
    optp_payoff dummy;
    memset(&dummy, 0, sizeof(dummy));
    dummy.kind = OPTP_PAY_VANILLA;
    dummy.flags = OPTP_F_CALL;
    PathArgs a;
    fill_args(a, proc, &dummy, 1, n_steps, n_base, antithetic, mode, seed, path_offset);
    a.z = z_dev;
    a.paths_out = (double*)c.out0.p;
    a.var_out = var_out ? (double*)c.out1.p : nullptr;
    a.stats = c.stats_dev;
    if (int e = dispatch_path(a, mode, c.stream)) return e;
    OPTP_CUDA(cudaMemcpyAsync(paths_out, c.out0.p, pbytes, cudaMemcpyDeviceToHost, c.stream));
    if (var_out)
        OPTP_CUDA(cudaMemcpyAsync(var_out, c.out1.p, pbytes, cudaMemcpyDeviceToHost, c.stream));
    OPTP_CUDA(cudaStreamSynchronize(c.stream));
    return 0;
}

// ------------------------------------------------------------------ payoffs over given paths
int optp_payoff_from_paths(const double* paths, uint64_t n_rows, uint64_t n_cols,
                           const optp_payoff* payoff, double* payoff_out,
                           double stats_out[OPTP_NSTATS]) {
    OPTP_ENTER();
    if (!paths || n_rows < 1 || n_cols < 1) return set_error(OPTP_E_ARG, "empty paths array");
    if (int e = check_payoffs(payoff, 1)) return e;
    if (!stats_out) return set_error(OPTP_E_ARG, "null stats pointer");
    Context& c = ctx();
    OPTP_CUDA(cudaSetDevice(c.device));
    const size_t bytes = sizeof(double) * (size_t)n_rows * n_cols;
    if (int e = arena_reserve(c.in0, bytes)) return e;
    OPTP_CUDA(cudaMemcpyAsync(c.in0.p, paths, bytes, cudaMemcpyHostToDevice, c.stream));
    double* pay_dev = nullptr;
    if (payoff_out) {
        if (int e = arena_reserve(c.out0, sizeof(double) * n_cols)) return e;
        pay_dev = (double*)c.out0.p;
    }
    if (int e = launch_payoff_from_paths((const double*)c.in0.p, n_rows, n_cols, *payoff, pay_dev,
                                         c.stats_dev, c.stream))
        return e;
    OPTP_CUDA(cudaMemcpyAsync(c.stats_host, c.stats_dev, sizeof(double) * OPTP_NSTATS,
                              cudaMemcpyDeviceToHost, c.stream));
    if (payoff_out)
        OPTP_CUDA(cudaMemcpyAsync(payoff_out, pay_dev, sizeof(double) * n_cols,
                                  cudaMemcpyDeviceToHost, c.stream));
    OPTP_CUDA(cudaStreamSynchronize(c.stream));
    memcpy(stats_out, c.stats_host, sizeof(double) * OPTP_NSTATS);
    return 0;
}

}  // extern "C"
