/* leniax_b200 — C ABI of the B200-native Lenia simulation hot path.
 *
 * This is the drop-in boundary for the path BASELINE.json names.  The reference (morgangiraud/leniax) is pure
 * Python on JAX and has no FFI of its own; every entry point below cites the Python interface it replaces
 * (file:line in the reference checkout).  INTEGRATION.md shows the ctypes binding a maintainer adds on the
 * reference side.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer marked "device" is a CUDA device pointer owned by the caller
 *     (PyTorch on the Python side).  The library allocates nothing persistent except plan-owned memory.
 *   - every call returns 0 (LNX_OK) or a negative lnx_status; the message is available through lnx_last_error()
 *     (thread local).  No C++ exception crosses the ABI.
 *   - launches go to the CUDA stream passed as `stream` (a cudaStream_t cast to void*; NULL = default stream).
 *     Calls are asynchronous like JAX dispatch; errors of asynchronous work surface at the caller's next sync.
 *   - a plan is immutable after creation: concurrent calls on different streams are safe.
 */
#ifndef LENIAX_B200_H
#define LENIAX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LNX_VERSION 100 /* 0.1.0 */
#define LNX_MAX_CHANNELS 8
#define LNX_MAX_KERNELS 32
#define LNX_NB_STATS 11 /* scalar statistics per world-step, order = lnx_stat_key */
#define LNX_COUT_ANY (-1)
#define LNX_COUT_NONE (-2)

typedef enum {
    LNX_OK = 0,
    LNX_ERR_INVALID = -1,     /* bad argument / unsupported configuration (Python side raises ValueError) */
    LNX_ERR_UNSUPPORTED = -2, /* valid leniax configuration that this build cannot run (NotImplementedError) */
    LNX_ERR_CUDA = -3,        /* CUDA runtime error (RuntimeError) */
    LNX_ERR_NO_DEVICE = -4    /* no usable sm_100 device: there is NO CPU fallback */
} lnx_status;

/* growth functions: leniax/growth_functions.py:256-264 (register) */
typedef enum {
    LNX_GF_POLY_QUAD4 = 0, LNX_GF_GAUSSIAN = 1, LNX_GF_GAUSSIAN_TARGET = 2, LNX_GF_STEP = 3,
    LNX_GF_STAIRCASE = 4, LNX_GF_TRIANGLE = 5, LNX_GF_IDENTITY = 6
} lnx_growth_fn;

/* state update functions: leniax/core.py:322-326 (register) */
typedef enum { LNX_STATE_V1 = 0, LNX_STATE_V2 = 1, LNX_STATE_SIMPLE = 2 } lnx_state_fn;

/* scalar statistics, leniax/statistics.py:102-115 (channel_mass is returned separately, shape [..., C]) */
typedef enum {
    LNX_ST_MASS = 0, LNX_ST_MASS_VOLUME, LNX_ST_MASS_DENSITY, LNX_ST_GROWTH, LNX_ST_GROWTH_VOLUME,
    LNX_ST_GROWTH_DENSITY, LNX_ST_MASS_SPEED, LNX_ST_MASS_ANGLE_SPEED, LNX_ST_MASS_GROWTH_DIST, LNX_ST_INERTIA,
    LNX_ST_POTENTIAL_VOLUME
} lnx_stat_key;

/* plan flags (lnx_desc.flags) */
#define LNX_PLAN_FORCE_TILED 1u /* run 128x128 worlds through the tiled multi-pass engine too (cross-checks, tests) */

/* run flags */
#define LNX_RUN_EARLY_STOP 1u /* stop a world once its stop criteria fired and >= 128 stat rows exist (rows after
                                 the stop are left untouched); default off = bit-for-bit [max_run_iter] rows like
                                 runner.py:207-213 */
#define LNX_RUN_NO_STATS 2u   /* reserved */
#define LNX_RUN_GENERIC_OLD 0x400u   /* several channels / kernels: use the older lnx_world128_generic kernel (everything in an L2
                                        scratch, statistics warp) instead of lnx_world128_gen_tm; cross-check in the tests */
#define LNX_RUN_GENERIC_1CTA 0x1000u  /* several channels / kernels: one world per SM (lnx_world128_gen_tm) even when the two-worlds-per-SM
                                        kernel applies; A/B runs and cross-check in the tests */
#define LNX_RUN_WEIGHTS_MATCH_COUT 0x2000u /* caller checked that `weights` is zero outside the (c_out[k], k) entries the plan declares:
                                        required by lnx_world128_gen2, which reads one weight per kernel */
#define LNX_RUN_TILED_GENERIC 0x800u /* 64^3 and 2048^2 one-channel one-kernel worlds: use the generic tiled passes instead of
                                        the thread-per-line / four-step kernels (A/B runs, cross-check in the tests) */
#define LNX_RUN_T64_LINE 0x4000u /* 64^3 one-channel one-kernel worlds: the round-1 thread-per-line step kernels (lnx_tiled64.cuh) instead of
                                        the half-line kernels (lnx_tiled64h.cuh); A/B runs, cross-check in the tests */
#define LNX_RUN_T2K_REAL_ROWS 0x8000u /* 2048^2 one-channel one-kernel worlds: rows kernels with one real row per warp (twice the warps, half the
                                        chain; measured no faster) instead of a packed row pair per warp; A/B runs, cross-check in the tests */
#define LNX_RUN_T64_STEPWISE 0x10000u /* 64^3 one-channel one-kernel worlds: one launch per pass and step (lead_h, plane_step, pass D) whatever the
                                        number of worlds (default above 128 worlds); A/B runs, cross-check in the tests */
#define LNX_RUN_T64_WHOLE_SCAN 0x20000u /* ... the persistent whole-scan kernel whatever the number of worlds (default up to 128 worlds) */
#define LNX_RUN_ASSUME_FINITE 0x100u /* caller checked that no growth s == 0 and no weight row sums to 0: NaN cannot
                                        appear, the fused kernel may use min/max clamps that do not propagate NaN */

/* Static description of the update + statistics functions.  Replaces the callables built by
 * helpers.build_update_fn (leniax/helpers.py:401-427: slugs + tc_indices + mean/sum) and
 * statistics.build_compute_stats_fn (leniax/statistics.py:11-33: R, dt = 1/T of the build-time config). */
typedef struct {
    int32_t nb_dims;                     /* world dimensions: 2 or 3 */
    int32_t dims[3];                     /* world size per dimension, powers of two in [8, 4096]; 128x128 worlds run in the
                                            shared-memory resident kernels, everything else in the tiled multi-pass engine.
                                            Any other size in [1, 4096] gives a plan that serves lnx_compute_stats only: such
                                            worlds are stepped by lnx_update_conv (2-D), the FFT entry points refuse the plan */
    int32_t nb_channels;                 /* C */
    int32_t nb_kernels;                  /* K = number of true kernels (after tc_indices) */
    int32_t nb_slots;                    /* C * max_k_per_channel = leading size of the reference's K tensor */
    int32_t slot[LNX_MAX_KERNELS];       /* tc_indices: slot of kernel k inside [C * max_k] (helpers.py:449-456) */
    int32_t c_in[LNX_MAX_KERNELS];       /* input channel of kernel k ( = slot / max_k, kernels.py:122-143) */
    int32_t gf_id[LNX_MAX_KERNELS];      /* lnx_growth_fn of kernel k (mapping.cin_gfs flattened) */
    int32_t c_out[LNX_MAX_KERNELS];      /* the ONE channel whose weight row is non-zero in column k (kernels.py:110-111: W[c_out][k] = h),
                                            LNX_COUT_NONE when the column is all zero; LNX_COUT_ANY in every entry = not declared: the
                                            weights tensor is used as given (slower kernel for several channels) */
    int32_t state_fn;                    /* lnx_state_fn (world_params.get_state_fn_slug) */
    int32_t weighted_average;            /* 1: core.weighted_mean, 0: core.weighted_sum (core.py:202-242) */
    float R;                             /* world_params.R used by the statistics (statistics.py:23) */
    float stats_dt;                      /* 1 / world_params.T of the build-time config (statistics.py:24) */
    uint32_t flags;                      /* LNX_PLAN_* */
} lnx_desc;

typedef struct lnx_plan lnx_plan;

int lnx_version(void);
const char* lnx_last_error(void);

/* Number of CUDA devices usable by this library (compute capability 10.x); 0 when none.  */
int lnx_device_count(void);

int lnx_plan_create(const lnx_desc* desc, lnx_plan** out);
int lnx_plan_destroy(lnx_plan* plan);

/* Size in bytes of the prepared kernel-spectrum table for ONE solution (all K kernels). */
size_t lnx_kernel_table_bytes(const lnx_plan* plan);

/* Re-pack the reference's FFT kernels (kernels.get_kernels_and_mapping, leniax/kernels.py:145-149:
 * K = fftn(fftshift(padded kernels)), complex64 [n_sols][nb_slots][H][W], device) into the engine's per-thread
 * half-spectrum layout (scaled by 1/(2 H W)).  table: device, n_sols * lnx_kernel_table_bytes(). */
int lnx_kernels_prepare(const lnx_plan* plan, int32_t n_sols, const void* K_fft, void* table, void* stream);

/* Forward 2-D FFT of real images: images float32 [n][H][W] (device) -> spectra complex64 [n][H][W] (device), same
 * convention as jnp.fft.fftn.  Used by the Python layer to build K = fftn(fftshift(kernels)) (leniax/kernels.py:145-149)
 * with the engine's own butterflies (no cuFFT). */
int lnx_rfft2(const lnx_plan* plan, int32_t n_images, const float* images, void* spectra, void* stream);

/* Same for any supported world shape (2-D / 3-D, power-of-two dims): images float32 [n][dims...] -> complex64 [n][dims...]. */
int lnx_rfftn(int32_t nb_dims, const int32_t* dims, int32_t n_images, const float* images, void* spectra, void* stream);

/* Measure the FP32 FMA throughput of the current device with a register-resident FMA loop (iters x 128 FMA per thread,
 * 4 CTAs of 512 threads per SM).  Synchronous.  This is the denominator of the FP32 roofline bench.py reports. */
int lnx_measure_fp32_peak(int32_t iters, double* tflops, double* ms, void* stream);

/* Bytes of device scratch needed by lnx_run_scan for this plan (independent of the number of worlds; the fused
 * single-channel kernel only uses the first 256 bytes).  Concurrent calls must use distinct workspaces. */
size_t lnx_workspace_bytes(const lnx_plan* plan);

/* The scan: replaces runner.run_scan / runner.run_scan_mem_optimized (leniax/runner.py:119-215) including
 * _scan_fn (295-334), core.update (core.py:13-49), compute_stats (statistics.py:36-126) and
 * check_heuristics (statistics.py:134-205).
 *
 * World w = sol * n_init + init.  All arrays are float32 device pointers, C-contiguous:
 *   cells0        [n_sols][n_init][C][H][W]                      initial states
 *   table         lnx_kernels_prepare output for the n_sols solutions
 *   gf_params     [n_sols][K][2]                                 (m, s) per kernel
 *   weights       [n_sols][C][K]                                 kernels_weight_per_channel
 *   dt            [n_sols]                                       1 / T per solution (runner.py:307)
 *   stats         [LNX_NB_STATS][n_sols][max_run_iter][n_init]   out
 *   channel_mass  [n_sols][max_run_iter][n_init][C]              out
 *   n_alive       [n_sols][n_init]                               out: stats['N'] (runner.py:161-162, 212-213)
 *   final_cells   [n_sols][n_init][C][H][W]                      out, may be NULL
 *   cells_out / field_out   [n_sols][max_run_iter][n_init][C][H][W], potential_out [..][K][H][W]
 *                                                                out, each may be NULL (run_scan keeps them,
 *                                                                run_scan_mem_optimized does not)
 *   workspace     lnx_workspace_bytes() bytes of device scratch (contents undefined afterwards)
 *
 * Asynchronous: all work is enqueued on `stream`.  Worlds that are not 128 x 128 run as a loop of passes per step; for
 * max_run_iter >= 8 that loop is one CUDA graph (captured on a private stream, thread-local capture mode) launched
 * max_run_iter times into `stream`; the graph is released by a later call once its last launch has completed. */
/* Scratch needed by lnx_run_scan for n_sols x n_init worlds (the tiled engine keeps per-world spectra in the workspace). */
size_t lnx_workspace_bytes_for(const lnx_plan* plan, int32_t n_sols, int32_t n_init);

int lnx_run_scan(const lnx_plan* plan, int32_t n_sols, int32_t n_init, int32_t max_run_iter, uint32_t run_flags,
                 const float* cells0, const void* table, const float* gf_params, const float* weights, const float* dt,
                 float* stats, float* channel_mass, float* n_alive, float* final_cells, float* cells_out, float* field_out,
                 float* potential_out, void* workspace, size_t workspace_bytes, void* stream);

/* Stand-alone statistics of given arrays: replaces a direct call of the closure returned by
 * statistics.build_compute_stats_fn (leniax/statistics.py:36-126).  cells/field [n_worlds][C][dims...], potential
 * [n_worlds][K][dims...]; carry arrays are updated in place: total_shift_idx int32 [n_worlds][nb_dims], mass_centroid
 * float [nb_dims][n_worlds], mass_angle float [n_worlds].  Out: stats [LNX_NB_STATS][n_worlds], channel_mass [n_worlds][C]. */
int lnx_compute_stats(const lnx_plan* plan, int32_t n_worlds, const float* cells, const float* field, const float* potential,
                      int32_t* total_shift_idx, float* mass_centroid, float* mass_angle, float* stats, float* channel_mass, void* stream);

/* What qd.update_individuals consumes of a scan (leniax/qd.py:168-186), reduced on the device: for every world N and the mean of
 * rows [ns - window, ns) of each scalar statistic, ns = max(int(N), window) clamped to T.  planes: HOST array of LNX_NB_STATS
 * device pointers (order lnx_stat_key), each [n_sols][T][n_init]; n_alive [n_sols][n_init]; out [n_sols][n_init][1 + LNX_NB_STATS]
 * = (N, means...).  This block is what a sharded run all-gathers per generation. */
int lnx_summarize_stats(const float* const* planes, const float* n_alive, int32_t n_sols, int32_t T, int32_t n_init, int32_t window,
                        float* out, void* stream);

/* One step: replaces a call of the update_fn built by helpers.build_update_fn, i.e. core.update (leniax/core.py:13-49):
 * state [n_worlds][C][dims...] -> state_out (same shape), field_out [n_worlds][C][dims...], potential_out [n_worlds][K][dims...].
 * table: lnx_kernels_prepare() output for ONE solution; gf_params [K][2], weights [C][K], dt [1] (device).  Thin wrapper over
 * lnx_run_scan with max_run_iter = 1 (its scratch is stream-ordered, cudaMallocAsync on `stream`). */
int lnx_update(const lnx_plan* plan, int32_t n_worlds, const float* state, const void* table, const float* gf_params,
               const float* weights, const float* dt, float* state_out, float* field_out, float* potential_out, void* stream);

/* One step of core.update with the DIRECT-CONVOLUTION potential (leniax/core.py:105-146 get_potential, selected by
 * fft=False in helpers.build_get_potential_fn, helpers.py:464-488) followed by get_field / weighted mean-sum / get_state.
 * 2-D worlds of any size (desc->dims, no power-of-two restriction); desc also gives C, K, slot[], c_in[], gf_id[],
 * state_fn, weighted_average.  state [n][C][H][W], kernels = the reference's cropped K [nb_slots][1][kh][kw]
 * (kernels.py:116-117, 153-156), gf_params [K][2], weights [C][K]; out: state_out, field_out [n][C][H][W],
 * potential_out [n][K][H][W].  This is the reference's cross-check path, not the throughput path. */
int lnx_update_conv(const lnx_desc* desc, int32_t n_worlds, int32_t kh, int32_t kw, const float* state, const float* kernels,
                    const float* gf_params, const float* weights, float dt, float* state_out, float* field_out, float* potential_out,
                    void* stream);

/* ---- set-up either side of the scan, batched over all individuals of a QD generation ---- */

/* kernel shapes: leniax/kernels.py:312-317 (register; `raw` needs no rasterisation), kernel functions: leniax/kernel_functions.py:253-261 */
typedef enum { LNX_KSHAPE_EMPTY = 0, LNX_KSHAPE_CIRCLE_2D = 1, LNX_KSHAPE_ELLIPSE_2D = 2, LNX_KSHAPE_ORIENTED_ELLIPSE_2D = 3 } lnx_kernel_shape;
typedef enum {
    LNX_KF_POLY_QUAD = 0, LNX_KF_GAUSS_BUMP = 1, LNX_KF_STEP = 2, LNX_KF_GAUSS = 3, LNX_KF_THRESHOLD = 4, LNX_KF_STAIRCASE = 5,
    LNX_KF_TRIANGLE = 6
} lnx_kernel_fn;
#define LNX_MAX_RINGS 8
typedef struct lnx_kernel_spec {
    int32_t shape;               /* lnx_kernel_shape; EMPTY = a padded slot of the reference's K tensor (all zeros, kernels.py:122-143) */
    int32_t kf;                  /* lnx_kernel_fn (kf_slug) */
    int32_t nb_b;                /* number of rings = len(k_params[1]) */
    float r;                     /* k_params[0]: relative radius, the kernel covers ceil(r R) pixels either side */
    float bs[LNX_MAX_RINGS];     /* k_params[1]: ring heights */
    float kf_params[2];          /* q, or (m, s) for staircase / triangle */
    float a, b;                  /* ellipses: k_params[2], k_params[3] */
    float cos_theta, sin_theta;  /* ellipses: cos / sin of k_params[4] * pi */
} lnx_kernel_spec;

/* Rasterise n kernels in one launch: replaces kernels.circle_2d / ellipse_2d / oriented_ellipse_2d (leniax/kernels.py:176-309) and the
 * kernel functions (leniax/kernel_functions.py:7-250), fp32 in the reference's operation order.  specs: HOST array; out: device float32
 * [n][side][side], side even and >= 2 ceil(r R) for every kernel; a kernel of radius k px sits at offset side / 2 - k, i.e. centre-padding
 * `out` into the world places it exactly where the reference's own centre padding does (leniax/utils.py:231-263). */
int lnx_rasterize_kernels(int32_t n, const lnx_kernel_spec* specs, float R, int32_t side, float* out, void* stream);

/* K = fftn(fftshift(centre-padded kernel)) (leniax/kernels.py:145-149) for n kernels in 2 (2-D) or 3 (3-D) launches: an exact separable
 * DFT over the small support (fp64 accumulation and twiddles, one rounding to complex64).  spatial: device float32 [n][support...],
 * K_out: device complex64 [n][dims...].  Any world size (no power-of-two restriction); no cuFFT. */
int lnx_kernel_spectrum(int32_t nb_dims, const int32_t* dims, int32_t n, const int32_t* support, const float* spatial, void* K_out, void* stream);

/* n uniform numbers in [0, 1) from a counter-based generator (SplitMix64 of seed + index): stands in for jax.random.uniform
 * (initializations.py:24-28, 63-65; bit-parity with threefry is not provided, no reference test pins a random draw). */
int lnx_random_uniform(uint64_t seed, int64_t n, float* out, void* stream);

/* initializations.random_uniform (leniax/initializations.py:24-30): out[w][i] = make_array_compressible(u * maxvals[w]). */
int lnx_init_uniform(uint64_t seed, int32_t n_worlds, int64_t cells_per_world, const float* maxvals, float* out, void* stream);

/* initializations.perlin after the random draw (leniax/initializations.py:66-75) with perlin.generate_perlin_noise_2d
 * (leniax/perlin.py:16-71) and loader.make_array_compressible (leniax/loader.py:16-30), one CTA per world:
 * angles [n][res0][res1] (device) -> out [n][H][W] = quantise((noise - min) / max * scaling[w]).  noise_out != NULL: write the plain
 * noise there instead (scaling / out may then be NULL). */
int lnx_init_perlin(int32_t n_worlds, int32_t H, int32_t W, int32_t res0, int32_t res1, const float* angles, const float* scaling, float* out,
                    float* noise_out, void* stream);

/* The perlin initial states of a whole QD generation in ONE launch (replaces the per-individual calls of leniax/qd.py:121-125): world
 * (s, i) of individual s draws its angles 2 pi u(seeds[s], i res0 res1 + j) itself - the numbers lnx_random_uniform(seeds[s]) gives - and
 * proceeds as lnx_init_perlin.  seeds: HOST array [n_seeds]; scaling: device [n_seeds * nb_init]; out: device [n_seeds * nb_init][H][W]. */
int lnx_init_perlin_seeded(int32_t n_seeds, const uint64_t* seeds, int32_t nb_init, int32_t H, int32_t W, int32_t res0, int32_t res1,
                           const float* scaling, float* out, void* stream);

/* Host-only (no GPU needed): the channel-update order and tensor-memory slots lnx_world128_gen2 would use for this description
 * (c_in sorted, c_out declared).  Returns 1 and fills acc_slot[K] (slot kernel k adds into), acc_first[K] (1: first touch of the slot in
 * a step), upd_mask[K] (bit c: channel c is updated after kernel k), chan_slot[C] (slot holding channel c's field at its update, -1: no
 * kernel feeds it); 0 when the kernel graph needs more than two live accumulators or is too large (the scan then runs one world per
 * SM); negative lnx_status on bad arguments.  For tests and for callers that want to know which kernel a configuration will take. */
int lnx_gen2_schedule(const lnx_desc* desc, int32_t* acc_slot, int32_t* acc_first, int32_t* upd_mask, int32_t* chan_slot);

/* Name of the CUDA kernel family lnx_run_scan would launch for this plan/arguments ("fused", "generic2", "generic", "tiled"), for tests. */
const char* lnx_run_scan_variant(const lnx_plan* plan, int32_t with_trajectory);

#ifdef __cplusplus
}
#endif
#endif /* LENIAX_B200_H */
