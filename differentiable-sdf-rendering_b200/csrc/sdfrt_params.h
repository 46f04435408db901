// sdfrt_params.h -- host-side translation of the public C structs (include/sdfrt.h) into the
// plain parameter blocks consumed by the per-lane code (sdfrt_core.cuh / sdfrt_lane.cuh).
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>
#include <cstring>

#include "../../include/sdfrt.h"
#include "sdfrt_prb.cuh"

namespace sdfrt {

inline void tracer_defaults(sdfrt_tracer_params* p) {  // python/shapes.py:28-41
    p->step_scale = 1.0f;
    p->trace_eps = 1e-6f;
    p->refine_intersection = 1;
    p->use_extra_weight = 1;
    p->extra_thresh = 0.05f;
    p->sil_weight_offset = 0.05f;
    p->sil_weight_epsilon = 1e-6f;
    p->weight_power = 3;
    p->use_weight_sum_weight = 1;
    p->max_steps = 1 << 16;
}
inline void warp_defaults(sdfrt_warp_params* p) {  // python/configs.py:12-40 ("Warp")
    p->kind = SDFRT_WARP_2D;
    p->edge_eps = 0.01f;
    p->weight_strategy = 6;
    p->clamping_thresh = 0.05f;
    p->normalize_warp_field = 1;
    p->max_reparam_depth = -1;
    p->n_aux_rays = 16;   // python/configs.py:123 ("ConvolutionWarp"); only read for SDFRT_WARP_CONV
    p->kappa = 1e5f;
    p->power = 3.f;
}
inline void sensor_defaults(sdfrt_sensor* s) {  // python/util.py:115-138, sensor 0 of a 6-camera rig is set by the caller
    std::memset(s, 0, sizeof(*s));
    for (int i = 0; i < 4; ++i) s->to_world[i * 4 + i] = 1.f;
    s->fov_x_deg = 39.f;
    s->near_clip = 1e-2f;
    s->far_clip = 1e4f;
    s->width = 128;
    s->height = 128;
    s->filter_radius = 2.f;
    s->filter_stddev = 0.5f;
    s->sample_border = 1;
}
inline void shading_defaults(sdfrt_shading* s) {
    s->env_radiance[0] = s->env_radiance[1] = s->env_radiance[2] = 1.f;
    s->env_dist = 2.f * std::sqrt(3.f);
    s->albedo[0] = s->albedo[1] = s->albedo[2] = 0.5f;
    s->hide_emitters = 0;
}

inline void prb_defaults(sdfrt_prb_params* p) {  // python/integrators/reparam.py:14, sdf_prb_reparam.py:41
    p->max_depth = 4;
    p->rr_depth = 5;
}
inline PrbParams make_prb(const sdfrt_prb_params& p) {
    PrbParams o;
    o.max_depth = p.max_depth;
    o.rr_depth = p.rr_depth;
    return o;
}

inline TracerParams make_tracer(const sdfrt_tracer_params& p) {
    TracerParams t;
    t.step_scale = p.step_scale; t.trace_eps = p.trace_eps;
    t.refine_intersection = p.refine_intersection; t.use_extra_weight = p.use_extra_weight;
    t.extra_thresh = p.extra_thresh; t.sil_weight_offset = p.sil_weight_offset; t.sil_weight_epsilon = p.sil_weight_epsilon;
    t.weight_power = p.weight_power; t.use_weight_sum_weight = p.use_weight_sum_weight;
    t.max_steps = p.max_steps > 0 ? p.max_steps : (1 << 16);
    for (int i = 0; i < 3; ++i) { t.bmin[i] = -kBBoxDelta; t.bmax[i] = 1.f + kBBoxDelta; }  // identity to_world
    return t;
}
// Grid3d(data, transform) (python/shapes.py:378-403): to_local = inverse(to_world) (affine part), bbox = AABB of the eight
// transformed unit-cube corners +- 0.05.  to_world: row-major 4x4, or nullptr for the identity.  Returns false if singular.
inline bool set_grid_transform(GridView& G, TracerParams& T, const float* to_world) {
    G.has_transform = 0;
    for (int i = 0; i < 9; ++i) G.A[i] = (i % 4 == 0) ? 1.f : 0.f;
    G.tl[0] = G.tl[1] = G.tl[2] = 0.f;
    for (int i = 0; i < 3; ++i) { T.bmin[i] = -kBBoxDelta; T.bmax[i] = 1.f + kBBoxDelta; }
    if (to_world == nullptr) return true;
    double M[3][3], t[3];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) M[i][j] = to_world[i * 4 + j];
        t[i] = to_world[i * 4 + 3];
    }
    const double det = M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
                       M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
    if (!(std::fabs(det) > 1e-30)) return false;
    double inv[3][3];
    inv[0][0] = (M[1][1] * M[2][2] - M[1][2] * M[2][1]) / det; inv[0][1] = (M[0][2] * M[2][1] - M[0][1] * M[2][2]) / det;
    inv[0][2] = (M[0][1] * M[1][2] - M[0][2] * M[1][1]) / det; inv[1][0] = (M[1][2] * M[2][0] - M[1][0] * M[2][2]) / det;
    inv[1][1] = (M[0][0] * M[2][2] - M[0][2] * M[2][0]) / det; inv[1][2] = (M[0][2] * M[1][0] - M[0][0] * M[1][2]) / det;
    inv[2][0] = (M[1][0] * M[2][1] - M[1][1] * M[2][0]) / det; inv[2][1] = (M[0][1] * M[2][0] - M[0][0] * M[2][1]) / det;
    inv[2][2] = (M[0][0] * M[1][1] - M[0][1] * M[1][0]) / det;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) G.A[i * 3 + j] = (float)inv[i][j];
        G.tl[i] = (float)(-(inv[i][0] * t[0] + inv[i][1] * t[1] + inv[i][2] * t[2]));
    }
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int c = 0; c < 8; ++c) {
        const double x = c & 1, y = (c >> 1) & 1, z = (c >> 2) & 1;
        for (int i = 0; i < 3; ++i) {
            const double w = (double)(float)(M[i][0] * x + M[i][1] * y + M[i][2] * z + t[i]);
            lo[i] = std::min(lo[i], w); hi[i] = std::max(hi[i], w);
        }
    }
    for (int i = 0; i < 3; ++i) { T.bmin[i] = (float)lo[i] - kBBoxDelta; T.bmax[i] = (float)hi[i] + kBBoxDelta; }
    G.has_transform = 1;
    return true;
}
inline WarpParams make_warp(const sdfrt_warp_params& p) {
    WarpParams w;
    w.kind = p.kind; w.edge_eps = p.edge_eps; w.weight_strategy = p.weight_strategy;
    w.clamping_thresh = p.clamping_thresh; w.normalize_warp_field = p.normalize_warp_field;
    w.max_reparam_depth = p.max_reparam_depth;
    w.n_aux_rays = p.n_aux_rays; w.kappa = p.kappa; w.power = p.power;
    return w;
}
inline SensorParams make_sensor(const sdfrt_sensor& s) {
    SensorParams o;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) o.R[i][j] = s.to_world[i * 4 + j];
        o.c[i] = s.to_world[i * 4 + 3];
    }
    o.tan_half_fov = (float)std::tan(0.5 * (double)s.fov_x_deg * 3.14159265358979323846 / 180.0);
    o.aspect = (float)s.width / (float)s.height;
    o.near_clip = s.near_clip; o.far_clip = s.far_clip;
    o.W = s.width; o.H = s.height;
    // border_size = ceil(radius - 0.5) (mitsuba ReconstructionFilter), only sampled if sample_border
    o.border = s.sample_border ? (int)std::ceil(s.filter_radius - 0.5f) : 0;
    o.filter_radius = s.filter_radius;
    o.filter_alpha = -1.f / (2.f * s.filter_stddev * s.filter_stddev);
    o.filter_bias = (float)std::exp((double)o.filter_alpha * s.filter_radius * s.filter_radius);
    return o;
}
inline ShadingParams make_shading(const sdfrt_shading& s) {
    ShadingParams o;
    for (int i = 0; i < 3; ++i) o.Le[i] = s.env_radiance[i];
    o.env_dist = s.env_dist;
    o.hide_emitters = s.hide_emitters;
    o.shadow_origin = 0;
    o.use_mis = 0;
    return o;
}
// host tables of the `envmap` emitter (oracle/ref_semantics.md S18; oracle/sdf_oracle.py EnvMap): the bitmap with its first column
// repeated, and the piecewise-constant sampling distribution over the bilinear cells (computed in double, stored as float)
struct EnvTables {
    std::vector<float> data, marg_cdf, cond_cdf, cell_prob;
    int W1 = 0, H = 0;
};
inline bool build_envmap(const float* rgb, int w, int h, float scale, EnvTables& t) {
    if (rgb == nullptr || w < 1 || h < 2) return false;
    t.W1 = w + 1; t.H = h;
    t.data.resize((size_t)h * t.W1 * 3);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < t.W1; ++x)
            for (int c = 0; c < 3; ++c) t.data[((size_t)y * t.W1 + x) * 3 + c] = rgb[((size_t)y * w + (x % w)) * 3 + c] * scale;
    const int nx = w, ny = h - 1;
    std::vector<double> wgt((size_t)nx * ny), rows(ny, 0.0);
    auto lum = [&](int y, int x) {
        const float* q = &t.data[((size_t)y * t.W1 + x) * 3];
        return 0.212671 * (double)q[0] + 0.715160 * (double)q[1] + 0.072169 * (double)q[2];
    };
    double total = 0.0;
    for (int y = 0; y < ny; ++y) {
        const double st = std::sin(((double)y + 0.5) / (double)ny * 3.14159265358979323846);
        for (int x = 0; x < nx; ++x) {
            const double v = 0.25 * (lum(y, x) + lum(y, x + 1) + lum(y + 1, x) + lum(y + 1, x + 1)) * st;
            if (!(v >= 0.0)) return false;   // negative or NaN radiance
            wgt[(size_t)y * nx + x] = v;
            rows[y] += v;
        }
        total += rows[y];
    }
    if (!(total > 0.0)) return false;
    t.cell_prob.resize((size_t)nx * ny); t.cond_cdf.resize((size_t)nx * ny); t.marg_cdf.resize(ny);
    double acc = 0.0;
    for (int y = 0; y < ny; ++y) {
        acc += rows[y] / total;
        t.marg_cdf[y] = (float)acc;
        double a = 0.0;
        for (int x = 0; x < nx; ++x) {
            t.cell_prob[(size_t)y * nx + x] = (float)(wgt[(size_t)y * nx + x] / total);
            a += wgt[(size_t)y * nx + x] / std::max(rows[y], 1e-300);
            t.cond_cdf[(size_t)y * nx + x] = (float)a;
        }
    }
    return true;
}
inline PassParams make_pass(const SensorParams& S, uint32_t seed, int spp, int rank, int world, int antithetic = 0, int assign_rows = kBandRows) {
    PassParams p;
    p.seed = seed;
    p.spp = spp;
    p.log2_spp = -1;
    for (int k = 0; k < 31; ++k)
        if ((1 << k) == spp) p.log2_spp = k;
    p.rank = rank;
    p.world = world;
    p.Wb = S.W + 2 * S.border;
    p.Hb = S.H + 2 * S.border;
    // sub-bands rank, rank + world, ... of assign_rows film rows each (the last sub-band of the film may be shorter)
    p.assign_rows = (assign_rows == 1 || assign_rows == 2 || assign_rows == 4) ? assign_rows : kBandRows;
    const int nsub = (p.Hb + p.assign_rows - 1) / p.assign_rows;
    unsigned long long rows = 0;
    for (int g = rank; g < nsub; g += world) rows += (unsigned long long)std::min(p.assign_rows, p.Hb - g * p.assign_rows);
    p.rows_local = (int)rows;
    p.antithetic = antithetic ? 1 : 0;
    p.n_items = rows * (unsigned long long)p.Wb * (unsigned long long)spp * (unsigned long long)(1 + p.antithetic);
    p.win_lo = 0ull;
    p.win_hi = ~0ull;
    return p;
}

}  // namespace sdfrt
