/* libsdfrt -- C ABI of the B200-native differentiable SDF rendering hot path.
 *
 * Drop-in boundary for rgl-epfl/differentiable-sdf-rendering (reference paths are relative to
 * /root/reference).  The reference has no FFI of its own -- its device code is JIT-traced
 * from Python by Dr.Jit -- so every entry point cites the Python interface it replaces.
 *
 * Conventions: plain pointers and sizes only (no torch / Dr.Jit types).  Unless a function
 * name ends in `_host`, every data pointer is a DEVICE pointer owned by the caller and
 * every call is asynchronous on `stream` (a cudaStream_t passed as void*).  Functions
 * return 0 on success, <0 on error; sdfrt_last_error() describes the last failure of the
 * context.  One context per (process, device); a context is not thread-safe.
 * There is no CPU fallback: creation fails if no sm_100 device is usable.
 */
#ifndef SDFRT_H
#define SDFRT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sdfrt_ctx sdfrt_ctx;

enum { SDFRT_OK = 0, SDFRT_ERR_INVALID = -1, SDFRT_ERR_CUDA = -2, SDFRT_ERR_STATE = -3 };

/* integrator plugins, python/integrators/{sdf_silhouette,sdf_simple_shading,sdf_direct,sdf_prb}_reparam.py */
enum { SDFRT_INTEGRATOR_SILHOUETTE = 0, SDFRT_INTEGRATOR_SIMPLE_SHADING = 1, SDFRT_INTEGRATOR_DIRECT = 2, SDFRT_INTEGRATOR_PRB = 3 };
/* warp field kinds, python/warp.py: None / DummyWarpField(:179) / WarpField2D(:7) */
enum { SDFRT_WARP_NONE = 0, SDFRT_WARP_DUMMY = 1, SDFRT_WARP_2D = 2, SDFRT_WARP_CONV = 3 /* WarpFieldConvolution, python/warp.py:126-176 */ };

/* SDFBase tracer knobs, python/shapes.py:28-41 (same defaults via sdfrt_tracer_defaults) */
typedef struct sdfrt_tracer_params {
    float step_scale;            /* :30 */
    float trace_eps;             /* :31 */
    int32_t refine_intersection; /* :33 */
    int32_t use_extra_weight;    /* :34 */
    float extra_thresh;          /* :35 */
    float sil_weight_offset;     /* :36 */
    float sil_weight_epsilon;    /* :37 */
    int32_t weight_power;        /* :38 */
    int32_t use_weight_sum_weight; /* :39 */
    int32_t max_steps;           /* safety cap on the march (reference: unbounded, :170) */
} sdfrt_tracer_params;

/* WarpField2D attributes, python/warp.py:10-23 and python/configs.py:36-40 */
typedef struct sdfrt_warp_params {
    int32_t kind;               /* SDFRT_WARP_* */
    float edge_eps;             /* :13 (already divided by 4 for strategy 2, :22-23) */
    int32_t weight_strategy;    /* :14 */
    float clamping_thresh;      /* :16 */
    int32_t normalize_warp_field; /* :20 */
    int32_t max_reparam_depth;  /* :11 (WarpFieldConvolution: python/warp.py:132) */
    /* SDFRT_WARP_CONV only -- WarpFieldConvolution (Bangaru et al. 2020, the comparison baseline), python/warp.py:129-136 */
    int32_t n_aux_rays;         /* :131, 1..64 */
    float kappa;                /* :134 von Mises-Fisher concentration, 1e5 */
    float power;                /* :135 exponent of the harmonic weights, 3 */
} sdfrt_warp_params;

/* perspective sensor + hdrfilm + gaussian rfilter as built by python/util.py:115-138 */
typedef struct sdfrt_sensor {
    float to_world[16];   /* row-major 4x4, mitsuba look_at convention (columns left, up, dir, origin) */
    float fov_x_deg;      /* 39 */
    float near_clip, far_clip;
    int32_t width, height;
    float filter_radius;  /* gaussian: 4*stddev = 2 */
    float filter_stddev;  /* 0.5 */
    int32_t sample_border;
} sdfrt_sensor;

/* `constant` environment emitter + diffuse BSDF of the dummy SDF shape */
typedef struct sdfrt_shading {
    float env_radiance[3];
    float env_dist;        /* 2 * scene bounding-sphere radius (shadow-ray target distance) */
    float albedo[3];       /* used when no albedo grid is bound */
    int32_t hide_emitters; /* python/integrators/sdf_direct_reparam.py:12 */
} sdfrt_shading;

/* props of the path-replay integrator, python/integrators/sdf_prb_reparam.py:35-48 + reparam.py:14 (max_depth);
 * diffuse BSDF + constant environment emitter (the shading block above) */
typedef struct sdfrt_prb_params {
    int32_t max_depth;  /* loop iterations = path vertices incl. the escaping one; default 4 */
    int32_t rr_depth;   /* Russian roulette from this depth on; default 5 */
} sdfrt_prb_params;

/* device-counted work, for the roofline byte count */
typedef struct sdfrt_stats {
    uint64_t n_rays;          /* camera lanes processed */
    uint64_t n_evals_value;   /* 64-tap value-only evaluations */
    uint64_t n_evals_grad;    /* value+gradient evaluations */
    uint64_t n_evals_hess;    /* value+gradient+Hessian evaluations */
    uint64_t n_scatter_sites; /* 64-tap adjoint scatter sites */
    uint64_t n_albedo_lookups;
    uint64_t n_steps_max;     /* longest march */
    uint64_t n_hits;
    uint64_t n_evals_adjoint; /* value+gradient+Hessian evaluations of the adjoint sweep (k_backward) */
    uint64_t n_tap_loads;     /* marcher evaluations that actually gathered their 4x4x4 block (tap-cache misses) */
    uint64_t n_skipped;       /* camera rays proven clear of the surface by the brick lower bound (never marched) */
    uint64_t n_launches;      /* kernels launched by the library (host-side count) */
    uint64_t n_box_loads;     /* tap loads served from a TMA-staged shared-memory box instead of L1/L2 (shadow marchers, SDFRT_BOX) */
} sdfrt_stats;

/* per-ray tracer outputs of SDFBase.ray_intersect, python/shapes.py:288 (test hook) */
typedef struct sdfrt_ray_out {
    float* its_t;          /* [n] */
    float* warp_t;         /* [n] */
    float* warp_t_d;       /* [n*3] */
    float* warp_weight;    /* [n] */
    float* warp_weight_d;  /* [n*3] */
    int32_t* n_steps;      /* [n] or NULL */
} sdfrt_ray_out;

/* ---- lifetime ---------------------------------------------------------------------- */
sdfrt_ctx* sdfrt_create(int device, int* status);
void sdfrt_destroy(sdfrt_ctx* ctx);
const char* sdfrt_last_error(const sdfrt_ctx* ctx);
int sdfrt_version(void);

void sdfrt_tracer_defaults(sdfrt_tracer_params* p);  /* python/shapes.py:28-41 */
void sdfrt_warp_defaults(sdfrt_warp_params* p);      /* python/configs.py:12-40 ("Warp" config) */
void sdfrt_sensor_defaults(sdfrt_sensor* s);         /* python/util.py:115-138 */
void sdfrt_shading_defaults(sdfrt_shading* s);

/* ---- scene state (borrowed pointers; valid until the next set_* of the same kind) ------ */
/* Grid3d(data) + traverse()'s `sdf.data` / `sdf.p`: python/shapes.py:378-392,469-479.
 * sdf is (rz, ry, rx) C-order, x fastest; to_world: sdfrt_set_sdf_transform (identity unless set). */
int sdfrt_set_sdf(sdfrt_ctx* ctx, const float* sdf, int rx, int ry, int rz, const float p[3]);
/* NOTE: like Grid3d.parameters_changed -> texture.set_tensor (python/shapes.py:477-479), sdfrt_set_sdf must be called again after
 * the grid's VALUES change (e.g. an optimiser step in place): the marchers read derived copies of the grid (row-expanded copies,
 * per-brick minima) that are rebuilt by the first sweep after a sdfrt_set_sdf. */
/* Grid3d(data, transform) / integrator prop `sdf_to_world` (python/shapes.py:378-403, python/integrators/reparam.py:21-24):
 * affine to_world as a row-major 4x4 (last row 0 0 0 1), or NULL for the identity.  Lookups use to_local @ (x - p), gradients
 * and Hessians are pulled back with to_local^T, the bounding box becomes the AABB of the transformed unit cube +- 0.05. */
int sdfrt_set_sdf_transform(sdfrt_ctx* ctx, const float* to_world);
/* `main-bsdf.reflectance.volume.data`: (rz, ry, rx, 3) trilinear grid on [0,1]^3, or NULL */
int sdfrt_set_albedo(sdfrt_ctx* ctx, const float* rgb_grid, int rx, int ry, int rz);
int sdfrt_set_tracer(sdfrt_ctx* ctx, const sdfrt_tracer_params* p);
int sdfrt_set_warp(sdfrt_ctx* ctx, const sdfrt_warp_params* p);   /* integrator.warp_field = ... (python/shape_opt.py:47) */
int sdfrt_set_sensor(sdfrt_ctx* ctx, const sdfrt_sensor* s);
int sdfrt_set_shading(sdfrt_ctx* ctx, const sdfrt_shading* s);
int sdfrt_set_integrator(sdfrt_ctx* ctx, int integrator);         /* SDFRT_INTEGRATOR_* */
/* props `antithetic_sampling` of ReparamIntegrator (python/integrators/reparam.py:19,167-178; default off, "used in evaluation"):
 * every wavefront lane is evaluated twice, the second time at the mirrored film position with a clone of the sampler.
 * Applies to the three ReparamIntegrator.render-based integrators; sdf_prb_reparam has its own render without it. */
int sdfrt_set_antithetic(sdfrt_ctx* ctx, int enable);
/* props `detach_indirect_si` / `decouple_reparam` of sdf_direct_reparam (python/integrators/sdf_direct_reparam.py:13-14,44-49): the
 * interaction the shadow ray is spawned from -- 0 = si (default), 1 = detach(si), 2 = si_d0 (interaction of the detached primary ray) */
int sdfrt_set_shadow_origin(sdfrt_ctx* ctx, int mode);
/* prop `use_mis` of ReparamIntegrator (python/integrators/reparam.py:17) as sdf_direct_reparam uses it
 * (python/integrators/sdf_direct_reparam.py:76-79,87-107): the emitter sample is weighted by the power heuristic against the BSDF's pdf
 * and a second, BSDF-sampled (cosine-hemisphere) ray is traced from every hit; rays that escape see the environment emitter.  Its warp
 * record (reparameterised like the shadow ray, depth 1) goes to nine more rows of the lane records.  Default off (the reference's). */
int sdfrt_set_use_mis(sdfrt_ctx* ctx, int enable);

/* `envmap` emitter instead of the constant one (the scenes of the reference's figures light the object with a lat-long bitmap,
 * <emitter type="envmap"> with the identity to_world; here the bitmap comes as an array).  rgb_host: (height, width, 3) floats in
 * HOST memory, row 0 = +y pole, copied (and scaled by `scale`) into context-owned device memory; NULL: back to the constant emitter
 * of sdfrt_set_shading.  Lookup: Mitsuba's envmap eval (uv = (atan2(d.x, -d.z) / 2pi, acos(d.y) / pi), bilinear, first column
 * repeated), followed with the reparameterised directions, so dLe/dd enters the adjoint (python/integrators/sdf_direct_reparam.py:31,
 * 62-70,105).  Emitter sampling: piecewise-constant over the bilinear cells by luminance x sin(theta) (oracle/ref_semantics.md S18 --
 * Mitsuba's hierarchical sample warp is not restated; the estimator is unbiased either way).  sdf_direct_reparam only: a render with
 * sdf_prb_reparam fails with SDFRT_ERR_INVALID while an envmap is bound.  Synchronises the device. */
int sdfrt_set_envmap_host(sdfrt_ctx* ctx, const float* rgb_host, int width, int height, float scale);
/* `to_world` of the envmap emitter (<transform name="to_world"> of the scene file): row-major 4x4 in HOST memory, or NULL for the
 * identity.  Must be a rotation (the emitter is at infinity: a translation is ignored, anything else is refused).  Lookup direction
 * = to_world^-1 d, sampled directions are mapped with to_world (Mitsuba's envmap.cpp).  Kept across sdfrt_set_envmap_host. */
int sdfrt_set_envmap_transform(sdfrt_ctx* ctx, const float* to_world);
void sdfrt_prb_defaults(sdfrt_prb_params* p);
int sdfrt_set_prb(sdfrt_ctx* ctx, const sdfrt_prb_params* p);      /* used when the integrator is SDFRT_INTEGRATOR_PRB */
/* multi-GPU ray partition: the film rows (border included) are cut into bands of 8 rows; this context renders the
 * samples of the bands rank, rank + world, ... (interleaved for load balance).  RNG streams stay keyed by the global
 * wavefront lane index (python/integrators/reparam.py:51), so results do not depend on the world size. */
int sdfrt_set_partition(sdfrt_ctx* ctx, int rank, int world);
/* The same with the granularity of the assignment as a parameter: sub-bands of `rows` film rows (1, 2, 4 or 8; sdfrt_set_partition = 8).
 * Finer units level the row counts of the ranks at the price of coherence (measured: no gain at 8 ranks).
 * Whatever the granularity, a rank walks its own rows in tiles of 8 rows, so the single-rank traversal is unchanged. */
int sdfrt_set_partition_rows(sdfrt_ctx* ctx, int rank, int world, int rows);

/* Validation hook: restrict adjoint sweep 2 (sdfrt_adjoint_backprop) to the wavefront lanes [lane_lo, lane_hi) -- lane =
 * (pixel_y * (W+2b) + pixel_x) * spp + sample, python/integrators/reparam.py:141-155 -- so that the gradient contribution of a
 * lane subsample can be compared with a reference that evaluates only those lanes.  (0, 0) = all lanes (default). */
int sdfrt_set_lane_window(sdfrt_ctx* ctx, uint64_t lane_lo, uint64_t lane_hi);

/* The conservative empty-space test (one brick-minimum DDA per film pixel; pixels whose sample rays provably never evaluate an
 * SDF value below the hit / warp threshold are not marched) changes no result, only the work done.  enable = 0 marches every
 * camera ray like the reference does (python/shapes.py:166-238), which makes the device counters of sdfrt_get_stats comparable
 * with a count of the reference's evaluations.  Default: enabled. */
int sdfrt_set_empty_space_skip(sdfrt_ctx* ctx, int enable);

/* ---- rendering ---------------------------------------------------------------------- */
/* film accumulator geometry: (H+2b, W+2b, 4) floats, b = border of the reconstruction filter */
int sdfrt_film_dims(const sdfrt_ctx* ctx, int* height_b, int* width_b, int* border);
uint64_t sdfrt_wavefront_size(const sdfrt_ctx* ctx, int spp);

/* mi.render(...) primal: ReparamIntegrator.render(mode=Primal), python/integrators/reparam.py:120-185.
 * Zeroes and fills film_accum (this rank's lanes only); no develop. */
int sdfrt_render_primal(sdfrt_ctx* ctx, uint32_t seed, int spp, float* film_accum, void* stream);
/* film.develop(): image (H, W, 3) = rgb / weight, border cropped */
int sdfrt_develop(sdfrt_ctx* ctx, const float* film_accum, float* image, void* stream);

/* render_backward, python/integrators/reparam.py:187-190, split in two phases so that a
 * multi-GPU caller can all-reduce the film between them:
 *  accumulate: re-render at (seed_grad, spp_grad) with the differentiable tracer; fills
 *              film_accum and keeps per-lane tracer records inside the context.
 *  backprop:   given the (all-reduced) film_accum of that re-render and grad_image =
 *              dLoss/dImage (H, W, 3), ADDS dLoss/d(sdf.data) into grad_sdf (rz, ry, rx),
 *              dLoss/d(albedo) into grad_albedo (or NULL) and dLoss/d(sdf.p) into grad_p[3]
 *              (device pointer, or NULL).
 * The SDF tensor must stay unchanged until the view's backprop has been enqueued (with SDFRT_DEFER_SIDE=1 the warp records of the few
 * shadow segments that need the differentiable march are still being filled on a stream of the context when accumulate returns;
 * backprop and every later sweep of the view slot wait for them). */
int sdfrt_adjoint_accumulate(sdfrt_ctx* ctx, uint32_t seed_grad, int spp_grad, float* film_accum, void* stream);
/* Primal render at (seed, spp) and adjoint sweep 1 at (seed_grad, spp_grad) in ONE pipelined call: the two sweeps do not
 * depend on each other (the re-render of render_backward does not need the loss, python/integrators/reparam.py:187-189),
 * so their batches share the pipeline and fill each other's latency tails.  Equivalent to sdfrt_render_primal followed by
 * sdfrt_adjoint_accumulate; sdfrt_adjoint_backprop follows once dLoss/dImage is known. */
int sdfrt_render_view(sdfrt_ctx* ctx, uint32_t seed, int spp, float* film_primal, uint32_t seed_grad, int spp_grad,
                      float* film_adjoint, void* stream);
int sdfrt_adjoint_backprop(sdfrt_ctx* ctx, const float* film_accum, const float* grad_image,
                           float* grad_sdf, float* grad_albedo, float* grad_p, void* stream);
/* The context keeps the lane records of ONE adjoint sweep 1 (the most recent sdfrt_adjoint_accumulate / sdfrt_render_view).
 * Returns a token that changes with every such sweep and is 0 when no records are valid (no sweep yet, or a set_* call since):
 * a caller that renders several views before back-propagating them (mi.render's CustomOp holds its inputs per view,
 * python/shape_opt.py:78-83) compares tokens and re-runs sweep 1 of a view whose records were overwritten. */
uint64_t sdfrt_adjoint_token(const sdfrt_ctx* ctx);
/* Views in flight.  The adjoint state of a view (lane records of sweep 1, the scene parameters it was rendered with, the film
 * adjoint, the per-view scratch of the empty-space test) lives in one of sdfrt_num_view_slots() slots; sdfrt_set_view_slot selects
 * the slot that the following sdfrt_adjoint_accumulate / sdfrt_render_view / sdfrt_adjoint_backprop / sdfrt_adjoint_token calls use
 * (default 0).  The views of one optimisation iteration are independent (python/shape_opt.py:77-83): a caller may enqueue the sweeps
 * of view v+1 (slot B, stream B) before the backprop of view v (slot A, stream A) so that the persistent kernels of one view fill
 * the latency tails of the other.  set_* calls only invalidate the selected slot; a slot back-propagates with the scene parameters
 * its sweep 1 saw.  The grid's derived copies are shared: do not rebind the SDF (sdfrt_set_sdf) while a view is in flight. */
int sdfrt_set_view_slot(sdfrt_ctx* ctx, int slot);
int sdfrt_num_view_slots(void);

/* device work counters accumulated since the last reset (synchronises `stream`) */
int sdfrt_get_stats(sdfrt_ctx* ctx, sdfrt_stats* out, int reset, void* stream);

/* Optional per-kernel timing with CUDA events recorded on the launching stream (for roofline reporting).
 * classes: 0 gen_primary, 1 march (primary segments), 2 shade, 3 march (shadow segments), 4 shadow_splat,
 *          5 film_adjoint, 6 backward, 7 develop.  sdfrt_get_timing synchronises the events, adds the elapsed
 * milliseconds of every launch since the last call into ms[8] and the launch counts into launches[8]. */
enum { SDFRT_NUM_KERNEL_CLASSES = 8 };
int sdfrt_set_profiling(sdfrt_ctx* ctx, int enable);
/* Batches of a forward sweep normally alternate between two internal streams (the tail of one batch's persistent
 * kernels overlaps the next batch).  enable = 0 serialises them on the caller's stream, which makes the per-kernel
 * event timings additive (used for roofline reporting). */
int sdfrt_set_overlap(sdfrt_ctx* ctx, int enable);
int sdfrt_get_timing(sdfrt_ctx* ctx, float* ms, int* launches);
/* Same, plus ms_union[9] (optional; entry 8 = all k_march launches, classes 1 and 3 together): per class, the length of the UNION of its launches' [start, stop] intervals -- with overlapping
 * batches on the internal streams the per-launch durations are not additive, the union is the time the class was on the device. */
int sdfrt_get_timing_ex(sdfrt_ctx* ctx, float* ms, int* launches, float* ms_union);

/* ---- test hooks / helpers used by the reference's callers ------------------------------ */
/* SDFBase.ray_intersect / ray_intersect_non_diff, python/shapes.py:115-339; o, d: [n*3], maxt: [n] or NULL */
int sdfrt_trace_rays(sdfrt_ctx* ctx, const float* o, const float* d, const float* maxt, int n,
                     int differentiable, const sdfrt_ray_out* out, void* stream);
/* Grid3d.eval / eval_and_grad / eval_all, python/shapes.py:420-450; x: [n*3];
 * out: order 0 -> [n], 1 -> [n*4] (v, g), 2 -> [n*10] (v, g, Hxx Hyy Hzz Hxy Hxz Hyz).
 * Also serves variables.upsample_sdf (python/variables.py:18-23). */
int sdfrt_eval_sdf(sdfrt_ctx* ctx, const float* x, int n, int order, float* out, void* stream);

/* ---- per-iteration grid update (the step on the other side of the gradient) ------------------
 * These calls work on caller-owned device grids of (rz, ry, rx) floats and do not touch the bound scene. */

/* redistancing.redistance(phi) = fastsweep.redistance, python/redistancing.py:4-13, as applied by
 * SdfVariable.validate (python/variables.py:185-189) to max(sdf, bbox_sdf): `lower_bound` (optional) is that
 * bbox_sdf grid.  Unit-cube domain, node spacing 1/resolution per axis, sign of the input kept.  out != phi.
 * Asynchronous on `stream` unless iters_out (host int, optional: number of work lists the solver ran) is given. */
int sdfrt_redistance(sdfrt_ctx* ctx, const float* phi, const float* lower_bound, float* out, int rx, int ry, int rz,
                     int* iters_out, void* stream);
/* Outcome of the last asynchronous sdfrt_redistance (waits for it): SDFRT_OK, or SDFRT_ERR_CUDA if the solver aborted
 * (grid barrier time-out / iteration cap) and `out` holds unreached nodes.  Also checked by the next sdfrt_redistance. */
int sdfrt_redistance_status(sdfrt_ctx* ctx);

/* mi.ad.Adam.step (python/shape_opt.py:53,102) for one parameter tensor of n floats with the reference's validation fused
 * in: gradient NaN scrub + clamp to [-grad_clamp, grad_clamp] (SdfVariable.validate_gradient, python/variables.py:193-199;
 * grad_clamp <= 0 disables it; the scrubbed gradient is written back) and value clamp to [lo, hi]
 * (VolumeVariable.validate, python/variables.py:114-117; lo > hi disables it).  t = 1-based step count of this tensor.
 * The hyper-parameters are doubles like the reference's Python scalars (beta and 1 - beta are rounded to fp32 separately). */
int sdfrt_adam_step(sdfrt_ctx* ctx, float* param, float* grad, float* m, float* v, size_t n, double lr, double beta1, double beta2,
                    double eps, int t, float grad_clamp, float lo, float hi, void* stream);

/* regularizations.eval_discrete_laplacian_reg (python/regularizations.py:5-25): *loss_out (device double, optional) +=
 * weight * sum_i (c_i - mean of the 6 clamped neighbours)^2, grad (optional) += weight * d(sum)/dc. */
int sdfrt_laplacian_reg(sdfrt_ctx* ctx, const float* data, int rx, int ry, int rz, float weight, double* loss_out, float* grad,
                        void* stream);

/* ---- end-to-end entry point with HOST buffers (what a plugin without device memory calls) --
 * One optimisation-loop view: upload sdf (host, rz*ry*rx floats), render primal at (seed, spp),
 * re-render + backprop at (seed_grad, spp_grad) against grad_image_host (H*W*3; if NULL the
 * reference benchmark's loss dr.sum(img) is used, figures/benchmark/benchmark.py:44), download
 * image_host (H*W*3) and grad_sdf_host (rz*ry*rx).  Synchronous for the caller; the work runs on a non-blocking stream owned by the
 * context, so host threads driving DIFFERENT contexts (one per thread; a context is not thread-safe) overlap each other's copies
 * and kernels -- e.g. the views of one optimisation iteration, which are independent (python/shape_opt.py:77-83).  Use pinned host
 * memory for asynchronous copies. */
int sdfrt_render_view_host(sdfrt_ctx* ctx, const float* sdf_host, int rx, int ry, int rz,
                           uint32_t seed, int spp, uint32_t seed_grad, int spp_grad,
                           const float* grad_image_host, float* image_host, float* grad_sdf_host);

/* Same with the reflectance grid as a second parameter (`main-bsdf.reflectance.volume.data`, python/opt_configs.py:286):
 * albedo_host (az*ay*ax*3 floats, or NULL to keep the bound grid) is uploaded too and, if grad_albedo_host is given,
 * dLoss/d(reflectance) comes back with the SDF gradient -- one textured optimisation-loop view, host buffers in and out. */
int sdfrt_render_view_host_ex(sdfrt_ctx* ctx, const float* sdf_host, int rx, int ry, int rz, const float* albedo_host, int ax, int ay,
                              int az, uint32_t seed, int spp, uint32_t seed_grad, int spp_grad, const float* grad_image_host,
                              float* image_host, float* grad_sdf_host, float* grad_albedo_host);

#ifdef __cplusplus
}
#endif
#endif /* SDFRT_H */
