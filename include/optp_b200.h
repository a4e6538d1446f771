/* optp_b200.h -- C ABI of the B200-native pricing hot path.
 *
 * The reference (optpricer 0.3.0) is pure Python: it has no FFI layer, so the
 * "interface each entry point replaces" is a Python function of the reference.
 * Every declaration below cites that function (path:line under the reference's
 * src/optpricer/).  The Python host package (optpricer_b200) keeps the reference's
 * names/signatures and forwards to these symbols through ctypes.
 *
 * Conventions
 *   - plain C types only; all pointers are HOST pointers unless the function name
 *     ends in _dev (then they are DEVICE pointers on the current device and the
 *     call is asynchronous on the given cudaStream_t, passed as void*).
 *   - every function returns an int status: 0 ok, >0 argument error (OPTP_E_*),
 *     <0 CUDA runtime error (-(cudaError_t)); optp_last_error() gives the text.
 *   - the library keeps no pointer after a call returns.
 *   - threads: the library owns ONE context per process (stream, reduction scratch, staging
 *     arenas); every entry point takes a process-wide lock, so concurrent calls from several
 *     host threads are serialised, never interleaved (the lock covers the host side of a call: the
 *     _dev variants return once their kernels are queued).  Monte Carlo launches take the 8 slots of the
 *     reduction scratch round-robin, so _dev launches queued on up to 8 different streams overlap on the
 *     device; the CRR _dev entry point uses one work-list scratch: one stream at a time.
 *   - no CPU fallback exists: without a usable CUDA device every compute entry
 *     point returns a negative status.
 */
#ifndef OPTP_B200_H
#define OPTP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OPTP_ABI_VERSION 7

/* argument errors */
#define OPTP_E_ARG 1        /* bad scalar argument / null pointer            */
#define OPTP_E_KIND 2       /* unknown process / payoff kind                 */
#define OPTP_E_TOO_LARGE 3  /* N or table too large for the device kernels   */
#define OPTP_E_NOINIT 4     /* optp_init was not called on this device       */

/* ---- process kinds (reference processes.py) ------------------------------ */
#define OPTP_PROC_GBM 0          /* gbm_paths                 processes.py:30-56   */
#define OPTP_PROC_HESTON_FTE 1   /* heston_paths              processes.py:117-162 */
#define OPTP_PROC_LV_MILSTEIN 2  /* milstein_local_vol_paths  processes.py:311-381 */
#define OPTP_PROC_LV_EULER 3     /* local_vol_paths           processes.py:223-258 */
#define OPTP_PROC_GBM_MILSTEIN 4 /* gbm_milstein_paths        processes.py:264-305 */
#define OPTP_PROC_SABR 5         /* sabr_paths                processes.py:168-217 */
#define OPTP_PROC_HESTON_QE 6    /* Andersen (2008) QE -- no reference code, parity unpinned */
#define OPTP_PROC_MERTON 7       /* merton_jump_paths         processes.py:62-111  */
#define OPTP_PROC_HESTON_QE_M 8  /* QE with Andersen's martingale correction K0* (2008, sec. 4.3.2): E[S_T] is exact,
                                    which the control variate needs -- no reference code, parity unpinned */

/* ---- payoff kinds (reference exotics.py) ---------------------------------- */
#define OPTP_PAY_VANILLA 0   /* vanilla leg of barrier_price  exotics.py:83-88   */
#define OPTP_PAY_DIGITAL 1   /* digital_price                 exotics.py:169-209 */
#define OPTP_PAY_BARRIER 2   /* barrier_price                 exotics.py:33-96   */
#define OPTP_PAY_ASIAN 3     /* asian_price                   exotics.py:102-163 */
#define OPTP_PAY_LOOKBACK 4  /* lookback_price                exotics.py:215-273 */

/* payoff flag bits */
#define OPTP_F_CALL 1u        /* else put                                        */
#define OPTP_F_UP 2u          /* barrier: up  (else down)                        */
#define OPTP_F_IN 4u          /* barrier: knock-in (else knock-out)              */
#define OPTP_F_GEOMETRIC 8u   /* asian: geometric (else arithmetic) average      */
#define OPTP_F_FLOATING 16u   /* asian / lookback: floating strike (else fixed)  */

/* ---- arithmetic / rng modes ------------------------------------------------ */
#define OPTP_MODE_FP32_PHILOX 0 /* production: fp32 path state, Philox4x32-10 + Box-Muller in
                                   registers, fp64 accumulators                              */
#define OPTP_MODE_FP64_PHILOX 1 /* reference operation order in fp64, on-device Philox normals */
#define OPTP_MODE_FP64_HOSTZ 2  /* deterministic check: fp64, consumes the caller's normals    */

typedef struct optp_process {
    int32_t kind;      /* OPTP_PROC_*                                             */
    int32_t n_nodes;   /* sigma table: nodes per row (LV_* only)                  */
    double S0, r, q, T;
    double sigma;                      /* GBM, GBM_MILSTEIN                          */
    double v0, kappa, theta, xi, rho;  /* HESTON_*; SABR: v0=alpha0, xi=nu, kappa=beta;
                                          MERTON: v0=lam, kappa=mJ, theta=sJ           */
    double dS_bump;                    /* LV_MILSTEIN (processes.py:320)             */
    /* sigma(S,t) table, row i = step time i*T/n_steps (i < n_steps), n_nodes log-uniform
       nodes ln S = x0 + h*j; linear interpolation in ln S, flat outside.  The table is
       this library's statement of the sigma_loc callable contract, processes.py:361-369. */
    const double* sigma_table;
    double x0, h;
} optp_process;

typedef struct optp_payoff {
    int32_t kind;    /* OPTP_PAY_*  */
    uint32_t flags;  /* OPTP_F_*    */
    double K, barrier, rebate, payout;
    double r_disc, T_disc; /* the r,T the payoff function discounts with, exotics.py:22 */
} optp_payoff;

/* statistics block written by the Monte Carlo entry points:
 *   [0] n (effective paths)  [1] sum X  [2] sum X^2  [3] sum Y  [4] sum Y^2  [5] sum XY
 *   [6] number of effective paths with a table lookup outside the tabulated range, i.e. clamped to
 *       the grid edge (LV_* only; fp64 kernels count the lookups themselves)  [7] reserved
 * X = exp(-r_disc*T_disc)*payoff, Y = exp(-r_disc*T_disc)*S_T  (monte_carlo.py:54-64). */
#define OPTP_NSTATS 8

/* ---- lifetime ------------------------------------------------------------- */
int optp_abi_version(void);
/* Create (idempotent) the per-device context: stream, scratch arena, partial-sum buffers. */
int optp_init(int device);
int optp_shutdown(void);
const char* optp_last_error(void);
int optp_device_count(void);
/* Number of kernels this library has launched since optp_init (bench.py "gpu_launches"). */
uint64_t optp_launch_count(void);
/* Tuning / diagnostics switches (all default to the production setting):
 *   "crr_band"      1 = American puts (r > 0, q >= 0) and, as dual puts, American calls (q > 0) go
 *                   through the banded kernel (default), 0 = every
 *                   contract through the full-width kernels (A/B checks of the band cuts)
 *   "lv_window"     1 = LV Milstein uses the one-load window table when the bump is smaller than a
 *                   grid cell (default), 0 = always the three-lookup (value, slope) path
 *   "crr_band_only" 1 = skip the full-width fallback pass; contracts the banded kernel declined
 *                   keep status 2 (diagnostics only)
 *   "crr_band_kernel" 2 = two-zone banded kernel (default), 1 = the earlier single-zone banded kernel
 *   "crr_zone_group"  options per warp of the two-zone kernel: 1, 2, 4, or 0 = by batch size (default);
 *                   prices are bit-identical for every setting
 *   "rng_mode"      generator mode of the fp32 GBM and Heston (FTE) path kernels: 0 = Philox4x32-10, two Box-Muller pairs of
 *                   (32-bit radius, 32-bit angle) per call (default; every parity statement is made for it),
 *                   1 = Philox4x32-7, same mapping, 2 = Philox4x32-10 with three Box-Muller pairs of
 *                   (26-bit radius, 16/16/18-bit angle) per call (6 normals; tails end at 6.1 sigma).
 *                   Modes 1 and 2 trade generator margin for speed and are opt-in; see optp_normal_stats
 *   "crr_euro_sum"  1 = European trees (and American contracts that are never exercised early) with
 *                   N <= 20000 are priced as the discounted binomial expectation of the terminal payoff,
 *                   O(N) per option (default); 0 = backward induction for every N
 *   "crr_count_ops" 1 = the two-zone CRR kernel counts the work its level loops execute (diagnostic build of
 *                   the same kernel; setting it also resets the counters), 0 = production kernels (default)
 * Returns OPTP_E_ARG for an unknown key. */
int optp_set_option(const char* key, long value);
/* Diagnostics counters (bench.py's physical roofline figure): "crr_fp64_lane_ops" = FP64-pipe lane-instructions
 * (DFMA, DMUL, DSETP) executed inside the level loops of the two-zone CRR kernel since "crr_count_ops" was
 * set, "crr_node_updates" = node updates executed there.  Synchronises the device. */
int optp_get_counter(const char* key, uint64_t* value_out);
/* Pinned host staging memory for callers that want DMA-able buffers. */
int optp_host_alloc(void** p, uint64_t bytes);
int optp_host_free(void* p);

/* ---- terminal-only European Monte Carlo ------------------------------------
 * Replaces _mc_chunk_sumstats (monte_carlo.py:11-64): draws (or consumes) n_base normals,
 * optional antithetic mirror, S_T = S0*exp((r-q-sigma^2/2)T + sigma*sqrt(T)*Z), and returns
 * the six sufficient statistics.  The estimator (monte_carlo.py:142-167) stays on the host.
 *   mode FP64_HOSTZ: z points to n_base fp64 normals (the reference's PCG64 draws).
 *   fp32_cast != 0 reproduces dtype=np.float32 (monte_carlo.py:35,42) in HOSTZ mode.
 *   path_offset: global index of the first base path (Philox counter; multi-GPU sharding). */
int optp_mc_terminal(double S0, double K, double T, double r, double q, double sigma,
                     int is_call, uint64_t n_base, int antithetic, int mode, int fp32_cast,
                     uint64_t seed, uint64_t path_offset, const double* z,
                     double stats_out[OPTP_NSTATS]);
int optp_mc_terminal_dev(double S0, double K, double T, double r, double q, double sigma,
                         int is_call, uint64_t n_base, int antithetic, int mode, int fp32_cast,
                         uint64_t seed, uint64_t path_offset, const double* z_dev,
                         double* stats_dev, void* stream);

/* ---- fused path simulation + payoff reduction ------------------------------
 * Replaces the pair  <generator>(...) -> ndarray -> <payoff>_price(paths, ...)
 * (processes.py:30-56,117-162,311-381 + exotics.py:18-273) without materialising paths.
 *   n_payoffs payoffs are evaluated on the SAME simulated path set (price_book.py:80,92-93
 *   seeds every row alike, i.e. common random numbers); stats_out holds n_payoffs blocks.
 *   z (HOSTZ mode): n_streams consecutive (n_steps, n_base) C-order blocks, in the
 *   reference's draw order (GBM/LV: Z; Heston/SABR: Z2 then Zp -- processes.py:138-139;
 *   Merton: Z, then the Poisson counts K as doubles, then ZJ -- processes.py:89-93).
 *   payoff_out (optional, may be NULL): n_payoffs x n_eff undiscounted per-path payoffs,
 *   columns [0,n_base) base paths, [n_base,2n_base) antithetic mirrors (processes.py:49-50).
 *   terminal_out (optional): n_eff terminal prices S_T. */
int optp_mc_path(const optp_process* proc, const optp_payoff* payoffs, uint32_t n_payoffs,
                 uint32_t n_steps, uint64_t n_base, int antithetic, int mode,
                 uint64_t seed, uint64_t path_offset, const double* z,
                 double* payoff_out, double* terminal_out, double* stats_out);
/* Device-resident variant: the sigma table must already be uploaded with
 * optp_sigma_table_upload (proc->sigma_table is ignored); stats_dev receives
 * n_payoffs*OPTP_NSTATS doubles; asynchronous on stream. */
int optp_mc_path_dev(const optp_process* proc, const optp_payoff* payoffs, uint32_t n_payoffs,
                     uint32_t n_steps, uint64_t n_base, int antithetic, int mode,
                     uint64_t seed, uint64_t path_offset, const double* z_dev,
                     double* payoff_out_dev, double* terminal_out_dev, double* stats_dev,
                     void* stream);
/* Stage a sigma(S,t) table (n_steps x n_nodes fp64, host) in HBM for later _dev calls. */
int optp_sigma_table_upload(const double* table, uint32_t n_steps, uint32_t n_nodes);
/* Handle-based staging, so that ONE optp_mc_book call can carry local-vol items with different tables
 * (bumped surfaces of a risk run; the desk workflow's barrier + vanilla legs share one).  Handles are small
 * positive integers owned by the context; at most 63 are live at a time. */
int optp_sigma_table_create(const double* table, uint32_t n_steps, uint32_t n_nodes, int32_t* handle_out);
int optp_sigma_table_destroy(int32_t handle);
/* Tabulate the Dupire local volatility of a raw-SVI surface ON THE DEVICE and stage it like
 * optp_sigma_table_upload does: row i = step time i*T/n_steps, node j at ln S = x0 + h*j.
 * Replaces the host loop over the reference's callable dupire_local_vol_func(surface, r, q)
 * (calibration.py:362-390 -> dupire_local_vol, :268-359; slices: SVIParams :14-66, surface lookup
 * VolSurface._get_forward :108-122 and .iv_from_logm :125-148).
 *   expiries[n_exp] ascending; svi[n_exp][5] = a, b, rho, m, sigma per slice;
 *   (fwd_T, fwd_F)[n_fwd] the surface's forward curve (required; np.interp with flat ends);
 *   dT = the finite-difference bump of dw/dT (reference default 1e-4);
 *   table_out (optional, host): receives the n_steps x n_nodes fp64 table. */
int optp_dupire_table(const double* expiries, const double* svi, uint32_t n_exp,
                      const double* fwd_T, const double* fwd_F, uint32_t n_fwd, double T,
                      uint32_t n_steps, double x0, double h, uint32_t n_nodes, double dT,
                      double* table_out);

/* Materialise paths for API compatibility with the generators' ndarray return value
 * (processes.py:56,162,381): paths_out is (n_steps+1, n_eff) fp64 C-order with the t=0 row;
 * var_out (optional, Heston return_variance=True, processes.py:147,161) has the same shape.
 * mode must be FP64_PHILOX or FP64_HOSTZ. */
int optp_paths(const optp_process* proc, uint32_t n_steps, uint64_t n_base, int antithetic,
               int mode, uint64_t seed, uint64_t path_offset, const double* z,
               double* paths_out, double* var_out);

/* Payoff reduction over caller-supplied paths (the literal exotics.py:18-273 contract:
 * a (n_rows, n_cols) fp64 C-order array whose row 0 is t=0).  One payoff per call. */
int optp_payoff_from_paths(const double* paths, uint64_t n_rows, uint64_t n_cols,
                           const optp_payoff* payoff, double* payoff_out,
                           double stats_out[OPTP_NSTATS]);

/* ---- book of Monte Carlo items ----------------------------------------------
 * Replaces the serial row loop of scripts/price_book.py:139-145 for its "mc" and "exotic" rows
 * (price_book.py:78-114): every item is one euro_price_mc call (terminal_only) or one
 * <generator> + <payoff(s)> pair; all items are launched back to back on the context's stream
 * and the statistics blocks come back in ONE device-to-host copy.
 *   stats_out: for every item, in order, n_payoffs blocks of OPTP_NSTATS doubles
 *   (terminal_only items: one block).  Local-vol items (milstein_local_vol_paths + payoffs, the desk
 *   workflow scripts/desk_workflow_localvol_barrier.py:136-148) name their table by handle
 *   (optp_sigma_table_create); proc.sigma_table is ignored.  path_offset shards the base paths of an item
 *   across GPUs. */
typedef struct optp_book_item {
    int32_t terminal_only;  /* 1: optp_mc_terminal semantics, uses proc.{S0,r,q,T,sigma}, K, is_call */
    int32_t n_payoffs;      /* payoffs[first_payoff .. first_payoff + n_payoffs), ignored if terminal_only */
    uint32_t first_payoff;
    uint32_t n_steps;
    uint64_t n_base, seed, path_offset;
    int32_t antithetic, mode;
    int32_t is_call;
    int32_t table_handle;   /* LV_* items: handle of the staged sigma table (>= 1); else ignored */
    double K;
    optp_process proc;
} optp_book_item;
int optp_mc_book(const optp_book_item* items, uint64_t n_items, const optp_payoff* payoffs,
                 uint64_t n_payoffs_total, double* stats_out);

/* ---- batched CRR binomial ---------------------------------------------------
 * Replaces crr (binomial.py:7-38) looped over a heterogeneous contract batch, and
 * crr_vec (binomial.py:44-109) when S0,T,r,q,sigma are constant across the batch.
 *   status_out[i] = 1 and price_out[i] = NaN when the risk-neutral p is outside (0,1)
 *   (binomial.py:16-17) or the spot / strike is not finite (the reference returns NaN there);
 *   N <= 0 is an argument error (binomial.py:9-10).  A bad contract never affects the others. */
int optp_crr_batch(const double* S0, const double* K, const double* T, const double* r,
                   const double* q, const double* sigma, const uint8_t* is_call,
                   uint64_t n, uint32_t N, int american, double* price_out,
                   uint8_t* status_out);
int optp_crr_batch_dev(const double* S0, const double* K, const double* T, const double* r,
                       const double* q, const double* sigma, const uint8_t* is_call,
                       uint64_t n, uint32_t N, int american, double* price_out,
                       uint8_t* status_out, void* stream);

/* ---- multi-GPU reduction (one process per GPU) -------------------------------------
 * Replaces _aggregate_stats over the reference's worker processes (monte_carlo.py:67-74, filled by the
 * ProcessPoolExecutor of monte_carlo.py:126-139): the statistics blocks are additive, so ONE sum over the
 * ranks gives the whole run.  The library owns one NCCL communicator per context (libnccl.so.2 is resolved
 * at run time with dlopen -- the copy already loaded in the process, e.g. torch's, is preferred).
 *   optp_comm_unique_id: rank 0 creates the 128-byte id; the caller ships it to the other ranks by any
 *                        means it has (torch.distributed broadcast, MPI, a file).
 *   optp_comm_init:      collective over the `world` ranks; binds to the device of optp_init.
 *   optp_allreduce_stats:     HOST buffer, in place (staged through pinned memory; synchronous).
 *   optp_allreduce_stats_dev: DEVICE buffer, in place, asynchronous on `stream` -- queue it behind the
 *                        *_dev kernel launches that produced the blocks and no host round trip is paid.
 * Without a communicator (world of one) both reductions are the identity and return 0. */
#define OPTP_COMM_ID_BYTES 128
int optp_comm_unique_id(char id_out[OPTP_COMM_ID_BYTES]);
int optp_comm_init(int rank, int world, const char id[OPTP_COMM_ID_BYTES]);
int optp_comm_destroy(void);
int optp_comm_info(int* rank_out, int* world_out);
int optp_allreduce_stats(double* stats, uint64_t n_doubles);
int optp_allreduce_stats_dev(double* stats_dev, uint64_t n_doubles, void* stream);

/* ---- generator self-checks ----------------------------------------------------------
 * optp_philox4x32_10: HOST-only evaluation of the library's Philox4x32-10 (the function the kernels inline)
 *   for known-answer tests against the Random123 vectors; needs neither a device nor optp_init.
 * optp_normal_stats: draws normals on the device exactly as the fused path kernels do (counter = (path,
 *   block, stream 0), key = seed; n_paths x n_blocks Philox calls, 4 or 6 normals each) and returns
 *   [0] n  [1..4] sum z, z^2, z^3, z^4  [5..8] counts of |z| > 3, 4, 5, 6  [9] sum of products of the two
 *   normals of a Box-Muller pair  [10] sum of products of the first normals of different pairs of a call
 *   [11] sum of (last normal of block c) x (first normal of block c+1) along a path  [12] max |z|
 *   [13], [14], [15] the number of products behind [9], [10], [11].
 * The reference draws with numpy's PCG64 ziggurat (processes.py:24,48); these checks are what stands behind
 * replacing it with a counter-based generator. */
int optp_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
int optp_normal_stats(uint64_t seed, uint64_t path_offset, uint64_t n_paths, uint32_t n_blocks, int rng_mode,
                      double stats_out[16]);

/* ---- roofline microbenchmarks (bench.py only) ---------------------------------
 * Measure the peaks MEASURED_PEAKS.json lacks, on the device, with CUDA events:
 *   which = 0 FP32 FFMA lane-ops/s, 1 MUFU.EX2 ops/s, 2 FP64 DFMA lane-ops/s,
 *           3 shared-memory bytes/s (LDS.64 conflict-free), 4 IMAD.WIDE+LOP3 pairs/s,
 *           5..14 single-opcode probes used to build the pipe model in profiles/ (IMAD.WIDE,
 *           IMAD.HI, IMAD, LOP3, I2FP, MUFU.SIN, MUFU.LG2, MUFU.SQRT, FMNMX, FADD lane-ops/s). */
int optp_microbench(int which, double* per_second_out, double* cta_cycles_out /* may be NULL */);

#ifdef __cplusplus
}
#endif
#endif /* OPTP_B200_H */
