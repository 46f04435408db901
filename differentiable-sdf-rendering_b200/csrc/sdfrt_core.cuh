// sdfrt_core.cuh -- per-ray math of the differentiable SDF rendering hot path (fp32).
//
// Everything here is straight-line per-lane arithmetic shared by the CUDA kernels in
// sdfrt_kernels.cu.  The functions are `__host__ __device__` so that a developer harness
// (tests/host_harness.cpp, CPU-only unit tests of the *math*) can call them; the shipped
// library only ever calls them from device code.
//
// Reference semantics (paths relative to the upstream repository):
//   python/shapes.py      SDFBase.ray_intersect & friends, Grid3d cubic B-spline lookups
//   python/warp.py        WarpField2D
//   python/math_util.py   normalize_sqr, bbox_distance_inside_d
//   python/integrators/   reparam.py, sdf_direct_reparam.py, sdf_silhouette_reparam.py,
//                         sdf_simple_shading_reparam.py
// Third-party behaviour (cubic texture, PCG32/TEA, perspective sensor, gaussian film,
// constant emitter, diffuse BSDF) is restated in oracle/ref_semantics.md.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define SDFRT_HD __host__ __device__ __forceinline__
#else
#define SDFRT_HD inline
#endif
// Division in the silhouette-weight arithmetic of the differentiable tracer (weights, not positions): on the device the
// 2-ulp MUFU-based form; the hit/miss decisions and step lengths never go through it.
#if defined(__CUDA_ARCH__)
#define SDFRT_FDIV(a, b) __fdividef((a), (b))
#else
#define SDFRT_FDIV(a, b) ((a) / (b))
#endif

#if !defined(__CUDACC__)
struct alignas(16) float4 { float x, y, z, w; };  // host builds (test harness, CPU port)
#endif

namespace sdfrt {

// ------------------------------------------------------------------------------------
// vectors
// ------------------------------------------------------------------------------------
struct V3 {
    float x, y, z;
};
SDFRT_HD V3 v3(float x, float y, float z) { return V3{x, y, z}; }
SDFRT_HD V3 v3(float s) { return V3{s, s, s}; }
SDFRT_HD V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
SDFRT_HD V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
SDFRT_HD V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
SDFRT_HD V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
SDFRT_HD V3 operator*(float s, V3 a) { return V3{a.x * s, a.y * s, a.z * s}; }
SDFRT_HD V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
SDFRT_HD V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }
SDFRT_HD V3& operator+=(V3& a, V3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
SDFRT_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
SDFRT_HD V3 normalize(V3 a) { return a / sqrtf(dot(a, a)); }
SDFRT_HD float sign1(float x) { return x >= 0.f ? 1.f : -1.f; }  // dr.sign
SDFRT_HD float fminf3(V3 a) { return fminf(a.x, fminf(a.y, a.z)); }
SDFRT_HD float comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

// symmetric 3x3 (Hessian): xx yy zz xy xz yz
struct S3 {
    float xx, yy, zz, xy, xz, yz;
};
SDFRT_HD V3 mul(const S3& H, V3 v) {
    return V3{H.xx * v.x + H.xy * v.y + H.xz * v.z, H.xy * v.x + H.yy * v.y + H.yz * v.z,
              H.xz * v.x + H.yz * v.y + H.zz * v.z};
}
// general 3x3, row major
struct M3 {
    float m[3][3];
};
SDFRT_HD V3 mul(const M3& A, V3 v) {
    return V3{A.m[0][0] * v.x + A.m[0][1] * v.y + A.m[0][2] * v.z, A.m[1][0] * v.x + A.m[1][1] * v.y + A.m[1][2] * v.z,
              A.m[2][0] * v.x + A.m[2][1] * v.y + A.m[2][2] * v.z};
}
SDFRT_HD V3 mulT(const M3& A, V3 v) {
    return V3{A.m[0][0] * v.x + A.m[1][0] * v.y + A.m[2][0] * v.z, A.m[0][1] * v.x + A.m[1][1] * v.y + A.m[2][1] * v.z,
              A.m[0][2] * v.x + A.m[1][2] * v.y + A.m[2][2] * v.z};
}

// ------------------------------------------------------------------------------------
// parameters (plain data; mirrors include/sdfrt.h)
// ------------------------------------------------------------------------------------
struct GridView {
    const float* data;  // (rz, ry, rx), x fastest
    int rx, ry, rz;
    float px, py, pz;  // sdf.p: lookups use x - p (python/shapes.py:408-414)
    // optional row-expanded copy for the marchers: rows4[(z * ry + y) * rx + x] = data[z][y][x .. x+3] (x <= rx - 4),
    // so that one 16-byte load fetches a whole x-row of a 4x4x4 block.  nullptr = read `data`.
    const float4* rows4;
    // ... or 8x expanded: rows8[((z * ry + y) * rx + x) * 8 + 0..7] = data[z][y][x .. x+3], data[z][y+1][x .. x+3], 32-byte aligned, so
    // that one 256-bit load (LDG.E.256, sm_100) fetches TWO x-rows of a block: 8 load instructions per lookup instead of 16.  For
    // incoherent (shadow) rays the L1 tag stage does one lookup per lane and load instruction, so this halves its work.
    const float* rows8;
    // Grid3d(data, transform) (python/shapes.py:378-392): lookups use x_local = A (x - p) + tl (`to_local`), gradients and
    // Hessians are pulled back with A^T (python/shapes.py:423-448).  has_transform == 0: identity (the reference's fast path).
    int has_transform;
    float A[9], tl[3];
};
struct AlbedoView {
    const float* data;  // (rz, ry, rx, 3) or nullptr
    int rx, ry, rz;
    float cr, cg, cb;  // constant reflectance when data == nullptr
};
struct TracerParams {
    float step_scale, trace_eps;
    int refine_intersection, use_extra_weight;
    float extra_thresh, sil_weight_offset, sil_weight_epsilon;
    int weight_power, use_weight_sum_weight, max_steps;
    // Grid3d.bbox(): world AABB of the transformed unit cube, inflated by 0.05 (python/shapes.py:394-403,416-418).  Kept with the
    // tracer parameters because every tracer routine that needs it already receives them.
    float bmin[3], bmax[3];
};
struct WarpParams {
    int kind;
    float edge_eps;
    int weight_strategy;
    float clamping_thresh;
    int normalize_warp_field, max_reparam_depth;
    int n_aux_rays;      // WARP_CONV (python/warp.py:129-136)
    float kappa, power;
};
struct SensorParams {
    float R[3][3];  // camera -> world rotation (columns left, up, dir)
    float c[3];     // camera origin
    float tan_half_fov, aspect, near_clip, far_clip;
    int W, H, border;
    float filter_radius, filter_alpha, filter_bias;  // gaussian: exp(alpha x^2) - bias
};
struct ShadingParams {
    float Le[3];
    float env_dist;
    int hide_emitters;
    // which interaction the shadow ray of sdf_direct_reparam is spawned from (python/integrators/sdf_direct_reparam.py:44-49):
    // 0 = si (attached, default), 1 = detach(si) (`detach_indirect_si`), 2 = si_d0 (`decouple_reparam`: the interaction of the
    // detached primary ray -- it still moves with the SDF, but not with the primary ray's warp)
    int shadow_origin;
    // ReparamIntegrator prop `use_mis` (python/integrators/reparam.py:17): sdf_direct_reparam combines the emitter sample with a
    // BSDF sample by the power heuristic (python/integrators/sdf_direct_reparam.py:76-79,87-107)
    int use_mis;
};
// `envmap` emitter (lat-long bitmap, identity to_world; oracle/ref_semantics.md S18).  data == nullptr: the `constant` emitter of
// ShadingParams.  The bitmap is stored with its first column repeated at the end (W1 = width + 1); the sampling tables describe a
// piecewise-constant distribution over the (W1 - 1) x (H - 1) bilinear cells, weight = mean corner luminance * sin(theta_centre).
struct EnvView {
    const float* data;       // (H, W1, 3)
    const float* marg_cdf;   // [H - 1] inclusive CDF over cell rows
    const float* cond_cdf;   // [(H - 1) * (W1 - 1)] inclusive CDF inside each row
    const float* cell_prob;  // [(H - 1) * (W1 - 1)]
    int W1, H;
    int has_rot;    // to_world is not the identity
    float R[9];     // to_world rotation, row-major: d_world = R d_local, d_local = R^T d_world
};
struct Counters {  // device work counters (unsigned long long for atomicAdd)
    unsigned long long n_rays, n_evals_value, n_evals_grad, n_evals_hess, n_scatter_sites, n_albedo_lookups,
        n_steps_max, n_hits, n_evals_adjoint, n_tap_loads, n_skipped, n_box_loads;
};

constexpr float kBBoxDelta = 0.05f;     // python/shapes.py:416-418
constexpr float kRayEpsilon = 1500.f * 5.9604644775390625e-08f;  // mitsuba math::RayEpsilon<float>
constexpr float kShadowEpsilon = 10.f * kRayEpsilon;
constexpr float kPi = 3.14159265358979323846f;

// ------------------------------------------------------------------------------------
// Sampler: PCG32 seeded through TEA (mitsuba `independent`), oracle/ref_semantics.md S2
// ------------------------------------------------------------------------------------
struct Pcg32 {
    uint64_t state, inc;
};
SDFRT_HD uint32_t pcg_next_u32(Pcg32& r) {
    uint64_t old = r.state;
    r.state = old * 0x5851F42D4C957F2DULL + r.inc;
    uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t)(old >> 59u);
    return (xs >> rot) | (xs << ((32u - rot) & 31u));
}
SDFRT_HD float pcg_next_f32(Pcg32& r) {
    uint32_t bits = (pcg_next_u32(r) >> 9) | 0x3f800000u;
    float f;
#if defined(__CUDA_ARCH__)
    f = __uint_as_float(bits);
#else
    union { uint32_t u; float f; } cv; cv.u = bits; f = cv.f;
#endif
    return f - 1.f;
}
SDFRT_HD Pcg32 sampler_seed(uint32_t seed, uint32_t lane) {
    uint32_t v0 = seed, v1 = lane, sum = 0;
    for (int i = 0; i < 4; ++i) {
        sum += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + sum) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + sum) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    Pcg32 r;
    r.state = 0;
    r.inc = ((uint64_t)v1 << 1) | 1ull;
    pcg_next_u32(r);
    r.state += (uint64_t)v0;
    pcg_next_u32(r);
    return r;
}

// ------------------------------------------------------------------------------------
// Tricubic B-spline lookups (drjit Texture3f.eval_cubic*, python/shapes.py:420-450)
// ------------------------------------------------------------------------------------
SDFRT_HD void bspline_w(float a, float w[4]) {
    float a2 = a * a, a3 = a2 * a;
    w[0] = (-a3 + 3.f * a2 - 3.f * a + 1.f) * (1.f / 6.f);
    w[1] = (3.f * a3 - 6.f * a2 + 4.f) * (1.f / 6.f);
    w[2] = (-3.f * a3 + 3.f * a2 + 3.f * a + 1.f) * (1.f / 6.f);
    w[3] = a3 * (1.f / 6.f);
}
SDFRT_HD void bspline_dw(float a, float w[4]) {
    float a2 = a * a;
    w[0] = -0.5f * a2 + a - 0.5f;
    w[1] = 1.5f * a2 - 2.f * a;
    w[2] = -1.5f * a2 + a + 0.5f;
    w[3] = 0.5f * a2;
}
SDFRT_HD void bspline_ddw(float a, float w[4]) {
    w[0] = 1.f - a;
    w[1] = 3.f * a - 2.f;
    w[2] = -3.f * a + 1.f;
    w[3] = a;
}

struct CellCoord {  // base tap index (may be out of range -> clamped per tap) and fractional offsets
    int ix, iy, iz;
    float ax, ay, az;
    bool interior;
};
SDFRT_HD CellCoord cell_coord(const GridView& G, V3 x) {
    CellCoord c;
    float lx = x.x - G.px, ly = x.y - G.py, lz = x.z - G.pz;
    if (G.has_transform) {  // call_wrap: to_local @ (x - p), python/shapes.py:408-414
        const float tx = G.A[0] * lx + G.A[1] * ly + G.A[2] * lz + G.tl[0];
        const float ty = G.A[3] * lx + G.A[4] * ly + G.A[5] * lz + G.tl[1];
        const float tz = G.A[6] * lx + G.A[7] * ly + G.A[8] * lz + G.tl[2];
        lx = tx; ly = ty; lz = tz;
    }
    float ux = lx * (float)G.rx - 0.5f;
    float uy = ly * (float)G.ry - 0.5f;
    float uz = lz * (float)G.rz - 0.5f;
    float fx = floorf(ux), fy = floorf(uy), fz = floorf(uz);
    c.ax = ux - fx; c.ay = uy - fy; c.az = uz - fz;
    // keep the int conversion defined for far-away points (texels clamp anyway)
    fx = fminf(fmaxf(fx, -4.f), (float)G.rx + 4.f);
    fy = fminf(fmaxf(fy, -4.f), (float)G.ry + 4.f);
    fz = fminf(fmaxf(fz, -4.f), (float)G.rz + 4.f);
    c.ix = (int)fx - 1; c.iy = (int)fy - 1; c.iz = (int)fz - 1;
    c.interior = (c.ix >= 0) & (c.iy >= 0) & (c.iz >= 0) & (c.ix + 3 < G.rx) & (c.iy + 3 < G.ry) & (c.iz + 3 < G.rz);
    return c;
}
SDFRT_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
SDFRT_HD int min_i(int a, int b) { return a < b ? a : b; }

// Clamp-to-edge addressing (WrapMode::Clamp) folded into the WEIGHTS: the interpolant is linear in the taps, so
// taps i-1..i+2 clamped to [0, r-1] equal the 4 in-range taps b..b+3 (b = clamp(i, 0, r-4)) with the weights of
// coinciding taps added up.  After folding every lookup is an unclamped 4x4x4 block read (r >= 4 per axis).
SDFRT_HD int fold_base(int i, int r) { return clampi(i, 0, r - 4); }
SDFRT_HD void fold_weights(int i, int b, int r, float w[4]) {
    if (i == b) return;  // interior along this axis
    float f[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int m = clampi(i + k, 0, r - 1) - b;  // 0..3
        f[0] += (m == 0) ? w[k] : 0.f;
        f[1] += (m == 1) ? w[k] : 0.f;
        f[2] += (m == 2) ? w[k] : 0.f;
        f[3] += (m == 3) ? w[k] : 0.f;
    }
    w[0] = f[0]; w[1] = f[1]; w[2] = f[2]; w[3] = f[3];
}

// A lookup is split in three stages so that the marcher can keep the 64 taps of the current cell across steps:
//   cell_setup   -> base cell (after clamp folding) + the (folded) B-spline weights of the three axes
//   load_taps    -> the 4x4x4 block f[(k*4+j)*4+i] = data[bz+k][by+j][bx+i]
//   reduce_taps  -> separable reduction (x per row, then y, then z: ~1/2 the FMAs of the naive 64 x 10 form)
// ORDER 0: value; 1: + gradient; 2: + Hessian.
template <int ORDER>
struct CellSetup {
    int bx, by, bz;
    float wx[4], wy[4], wz[4];
    float dx[ORDER >= 1 ? 4 : 1], dy[ORDER >= 1 ? 4 : 1], dz[ORDER >= 1 ? 4 : 1];
    float ex[ORDER >= 2 ? 4 : 1], ey[ORDER >= 2 ? 4 : 1], ez[ORDER >= 2 ? 4 : 1];
};
template <int ORDER>
SDFRT_HD void cell_setup(const GridView& G, V3 x, CellSetup<ORDER>& s) {
    const CellCoord c = cell_coord(G, x);
    bspline_w(c.ax, s.wx); bspline_w(c.ay, s.wy); bspline_w(c.az, s.wz);
    if (ORDER >= 1) { bspline_dw(c.ax, s.dx); bspline_dw(c.ay, s.dy); bspline_dw(c.az, s.dz); }
    if (ORDER >= 2) { bspline_ddw(c.ax, s.ex); bspline_ddw(c.ay, s.ey); bspline_ddw(c.az, s.ez); }
    s.bx = fold_base(c.ix, G.rx); s.by = fold_base(c.iy, G.ry); s.bz = fold_base(c.iz, G.rz);
    if (!c.interior) {
        fold_weights(c.ix, s.bx, G.rx, s.wx); fold_weights(c.iy, s.by, G.ry, s.wy); fold_weights(c.iz, s.bz, G.rz, s.wz);
        if (ORDER >= 1) { fold_weights(c.ix, s.bx, G.rx, s.dx); fold_weights(c.iy, s.by, G.ry, s.dy); fold_weights(c.iz, s.bz, G.rz, s.dz); }
        if (ORDER >= 2) { fold_weights(c.ix, s.bx, G.rx, s.ex); fold_weights(c.iy, s.by, G.ry, s.ey); fold_weights(c.iz, s.bz, G.rz, s.ez); }
    }
}
template <int ROWS = 0>  // 0: plain grid, 4: row-expanded copy (float4 per row), 8: two rows per 256-bit load
SDFRT_HD void load_taps(const GridView& G, int bx, int by, int bz, float f[64]) {
    // 32-bit element offsets (grids up to 2^32 texels): one address computation per row, taps at +0..+3
    const unsigned sy = (unsigned)G.rx, sz = (unsigned)G.rx * (unsigned)G.ry;
    const unsigned o = (unsigned)bz * sz + (unsigned)by * sy + (unsigned)bx;
#if defined(__CUDA_ARCH__)
    if (ROWS == 8) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int j = 0; j < 4; j += 2) {
                const float* q = G.rows8 + (size_t)(o + (unsigned)k * sz + (unsigned)j * sy) * 8u;
                float* d = f + (k * 4 + j) * 4;
                asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]), "=f"(d[4]), "=f"(d[5]), "=f"(d[6]), "=f"(d[7])
                             : "l"(q));
            }
        return;
    }
#endif
    if (ROWS == 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 r = G.rows4[o + (unsigned)k * sz + (unsigned)j * sy];
                f[(k * 4 + j) * 4 + 0] = r.x; f[(k * 4 + j) * 4 + 1] = r.y; f[(k * 4 + j) * 4 + 2] = r.z; f[(k * 4 + j) * 4 + 3] = r.w;
            }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float* row = G.data + (o + (unsigned)k * sz + (unsigned)j * sy);
#pragma unroll
                for (int i = 0; i < 4; ++i) f[(k * 4 + j) * 4 + i] = row[i];
            }
    }
}
template <int ORDER>
SDFRT_HD void reduce_taps(const float f[64], const CellSetup<ORDER>& s, const GridView& G, float& v, V3& g, S3& H) {
    float r_v = 0, r_gx = 0, r_gy = 0, r_gz = 0, r_xx = 0, r_yy = 0, r_zz = 0, r_xy = 0, r_xz = 0, r_yz = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float A0 = 0, A1 = 0, A2 = 0, B0 = 0, B1 = 0, C0 = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float f0 = f[(k * 4 + j) * 4 + 0], f1 = f[(k * 4 + j) * 4 + 1], f2 = f[(k * 4 + j) * 4 + 2], f3 = f[(k * 4 + j) * 4 + 3];
            const float a = s.wx[0] * f0 + s.wx[1] * f1 + s.wx[2] * f2 + s.wx[3] * f3;
            A0 += s.wy[j] * a;
            if (ORDER >= 1) {
                const float b = s.dx[0] * f0 + s.dx[1] * f1 + s.dx[2] * f2 + s.dx[3] * f3;
                A1 += s.dy[j] * a;
                B0 += s.wy[j] * b;
                if (ORDER >= 2) {
                    const float cc = s.ex[0] * f0 + s.ex[1] * f1 + s.ex[2] * f2 + s.ex[3] * f3;
                    A2 += s.ey[j] * a;
                    B1 += s.dy[j] * b;
                    C0 += s.wy[j] * cc;
                }
            }
        }
        r_v += s.wz[k] * A0;
        if (ORDER >= 1) {
            r_gz += s.dz[k] * A0;
            r_gy += s.wz[k] * A1;
            r_gx += s.wz[k] * B0;
            if (ORDER >= 2) {
                r_zz += s.ez[k] * A0;
                r_yz += s.dz[k] * A1;
                r_yy += s.wz[k] * A2;
                r_xz += s.dz[k] * B0;
                r_xy += s.wz[k] * B1;
                r_xx += s.wz[k] * C0;
            }
        }
    }
    v = r_v;
    if (ORDER >= 1) {
        const float sx = (float)G.rx, sy_ = (float)G.ry, sz_ = (float)G.rz;
        g = V3{r_gx * sx, r_gy * sy_, r_gz * sz_};
        if (ORDER >= 2) {
            H.xx = r_xx * sx * sx; H.yy = r_yy * sy_ * sy_; H.zz = r_zz * sz_ * sz_;
            H.xy = r_xy * sx * sy_; H.xz = r_xz * sx * sz_; H.yz = r_yz * sy_ * sz_;
        }
        if (G.has_transform) {  // g <- A^T g, H <- A^T H A (python/shapes.py:426-428,445-448)
            const float* A = G.A;
            const V3 gl = g;
            g = V3{A[0] * gl.x + A[3] * gl.y + A[6] * gl.z, A[1] * gl.x + A[4] * gl.y + A[7] * gl.z, A[2] * gl.x + A[5] * gl.y + A[8] * gl.z};
            if (ORDER >= 2) {
                const float Hl[3][3] = {{H.xx, H.xy, H.xz}, {H.xy, H.yy, H.yz}, {H.xz, H.yz, H.zz}};
                float HA[3][3], R[3][3];
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) HA[i][j] = Hl[i][0] * A[0 + j] + Hl[i][1] * A[3 + j] + Hl[i][2] * A[6 + j];
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) R[i][j] = A[0 + i] * HA[0][j] + A[3 + i] * HA[1][j] + A[6 + i] * HA[2][j];
                H.xx = R[0][0]; H.yy = R[1][1]; H.zz = R[2][2];
                H.xy = 0.5f * (R[0][1] + R[1][0]); H.xz = 0.5f * (R[0][2] + R[2][0]); H.yz = 0.5f * (R[1][2] + R[2][1]);
            }
        }
    }
}
template <int ORDER>
SDFRT_HD void eval_sdf(const GridView& G, V3 x, float& v, V3& g, S3& H) {
    CellSetup<ORDER> s;
    cell_setup<ORDER>(G, x, s);
    float f[64];
    load_taps(G, s.bx, s.by, s.bz, f);
    reduce_taps<ORDER>(f, s, G, v, g, H);
}
SDFRT_HD float eval_value(const GridView& G, V3 x) {
    float v; V3 g; S3 H;
    eval_sdf<0>(G, x, v, g, H);
    return v;
}

// ------------------------------------------------------------------------------------
// math_util.bbox_distance_inside_d (python/math_util.py:31-41); bbox = Grid3d.bbox() (unit cube +- 0.05 without a transform)
// ------------------------------------------------------------------------------------
SDFRT_HD float bbox_distance_inside_d(const TracerParams& P, V3 x, V3& dist_d) {
    const V3 dmin = V3{x.x - P.bmin[0], x.y - P.bmin[1], x.z - P.bmin[2]};
    const V3 dmax = V3{P.bmax[0] - x.x, P.bmax[1] - x.y, P.bmax[2] - x.z};
    const float dist = fmaxf(0.f, fminf(fminf3(dmin), fminf3(dmax)));
    const V3 amax = V3{fabsf(dmax.x), fabsf(dmax.y), fabsf(dmax.z)};
    const V3 amin = V3{fabsf(dmin.x), fabsf(dmin.y), fabsf(dmin.z)};
    const V3 md = V3{fminf(amin.x, amax.x), fminf(amin.y, amax.y), fminf(amin.z, amax.z)};
    V3 n = v3(0.f);
    if ((md.x < md.y) & (md.x < md.z)) n = v3(1, 0, 0);
    if ((md.y < md.z) & (md.y < md.x)) n = v3(0, 1, 0);
    if ((md.z < md.x) & (md.z < md.y)) n = v3(0, 0, 1);
    dist_d = v3(0.f);
    if (dist > 0.f) dist_d = V3{n.x * sign1(amax.x - amin.x), n.y * sign1(amax.y - amin.y), n.z * sign1(amax.z - amin.z)};
    return dist;
}

// ------------------------------------------------------------------------------------
// Ray / AABB (mitsuba BoundingBox3f::ray_intersect + contains), python/shapes.py:129-141
// ------------------------------------------------------------------------------------
struct BoxHit {
    bool intersects, inside;
    float mint, maxt;
};
SDFRT_HD BoxHit ray_box(const TracerParams& P, V3 o, V3 d) {
    BoxHit b;
    const V3 r = V3{1.f / d.x, 1.f / d.y, 1.f / d.z};
    const V3 t1 = V3{(P.bmin[0] - o.x) * r.x, (P.bmin[1] - o.y) * r.y, (P.bmin[2] - o.z) * r.z};
    const V3 t2 = V3{(P.bmax[0] - o.x) * r.x, (P.bmax[1] - o.y) * r.y, (P.bmax[2] - o.z) * r.z};
    b.mint = fmaxf(fminf(t1.x, t2.x), fmaxf(fminf(t1.y, t2.y), fminf(t1.z, t2.z)));
    b.maxt = fminf(fmaxf(t1.x, t2.x), fminf(fmaxf(t1.y, t2.y), fmaxf(t1.z, t2.z)));
    const bool inx = (o.x >= P.bmin[0]) & (o.x <= P.bmax[0]), iny = (o.y >= P.bmin[1]) & (o.y <= P.bmax[1]),
               inz = (o.z >= P.bmin[2]) & (o.z <= P.bmax[2]);
    b.inside = inx & iny & inz;
    const bool act = ((d.x != 0.f) | inx) & ((d.y != 0.f) | iny) & ((d.z != 0.f) | inz);
    b.intersects = act & (b.maxt >= b.mint);
    b.intersects = b.intersects & ((b.mint > 0.f) | b.inside);  // python/shapes.py:132
    return b;
}

// ------------------------------------------------------------------------------------
// SDFBase.eval_trace_weight (python/shapes.py:68-113, use_weight_ad == False)
// ------------------------------------------------------------------------------------
SDFRT_HD float ipow(float b, int e) {
    float r = 1.f;
    for (int i = 0; i < e; ++i) r *= b;
    return r;
}
SDFRT_HD float eval_trace_weight(const TracerParams& P, V3 d, int i, V3 x, float sdf_value, V3 g, const S3& H, V3& weight_d) {
    const float n_dot_d = dot(g, d);
    const float n_dot_n = dot(g, g);
    const float dot_ratio = SDFRT_FDIV(n_dot_d, n_dot_n);
    const float denom = P.sil_weight_epsilon + fabsf(sdf_value) + P.sil_weight_offset * n_dot_d * dot_ratio;
    const float dist_weight = SDFRT_FDIV(1.f, ipow(denom, P.weight_power));
    V3 bbox_dist_d;
    const float bbox_dist = bbox_distance_inside_d(P, x, bbox_dist_d);
    const float bbox_eps = 0.01f;
    const float bbox_weight = (i > 0) ? SDFRT_FDIV(fminf(bbox_dist, bbox_eps), bbox_eps) : 1.f;
    const float weight = dist_weight * bbox_weight;
    const V3 bbox_weight_d = ((i > 0) & (bbox_dist < bbox_eps)) ? bbox_dist_d * SDFRT_FDIV(1.f, bbox_eps) : v3(0.f);
    const V3 gradient = (2.f * dot_ratio) * (d - dot_ratio * g);
    const V3 denom_d = sign1(sdf_value) * g + P.sil_weight_offset * mul(H, gradient);  // gradient @ H (H symmetric)
    const V3 dist_weight_d = SDFRT_FDIV(-(float)P.weight_power * dist_weight, denom) * denom_d;
    weight_d = dist_weight * bbox_weight_d + bbox_weight * dist_weight_d;
    return weight;
}

// ------------------------------------------------------------------------------------
// Tracer outputs (python/shapes.py:288)
// ------------------------------------------------------------------------------------
struct TraceOut {
    float its_t, warp_t, warp_weight;
    V3 warp_t_d, warp_weight_d;
    int n_steps;
};

// ------------------------------------------------------------------------------------
// The tracers as resumable state machines: one ray *segment* (primary or shadow ray) is
//   segment_begin -> {value_step | diff_step}* -> refine_step* -> diff_finalize
// with exactly one SDF evaluation between two steps.  trace_non_diff / trace_diff below run a segment
// to completion (per-lane form used by the test hooks and the host harness); the persistent wavefront
// kernel (sdfrt_march.cuh) interleaves the steps of 32 segments per warp and refills finished slots.
// ------------------------------------------------------------------------------------
struct Segment {
    V3 o, d;       // d normalised (python/shapes.py:124)
    float t, maxt, trace_eps, its_t;
    int i;         // march step counter, then refine iteration counter
    bool intersects, inside;
};

// python/shapes.py:124-141.  Returns whether the march is active.
SDFRT_HD bool segment_begin(Segment& s, const TracerParams& P, V3 o, V3 d, float ray_maxt) {
    s.o = o;
    s.d = normalize(d);
    const BoxHit b = ray_box(P, o, s.d);
    s.maxt = fminf(b.maxt, ray_maxt);
    s.trace_eps = P.trace_eps * fmaxf(s.maxt, 1.f);
    s.its_t = INFINITY;
    s.t = b.inside ? 0.f : b.mint + 1e-5f;
    s.i = 0;
    s.intersects = b.intersects;
    s.inside = b.inside;
    return b.intersects;
}

// one iteration of SDFBase.ray_intersect_non_diff's loop (python/shapes.py:309-320); v = sdf(o + t d)
SDFRT_HD bool value_step(Segment& s, const TracerParams& P, float v) {
    if (P.step_scale != 1.f) v *= P.step_scale;
    const bool hit = v < s.trace_eps;
    if (hit) s.its_t = s.t;
    bool active = (s.t <= s.maxt) & !hit;
    s.t += hit ? 0.f : fabsf(v);
    active = active & (s.t <= s.maxt);
    ++s.i;
    return active & (s.i < P.max_steps);
}

// refine loop body, python/shapes.py:245-257 / :323-334; md = sdf(o + its_t d); s.i counts iterations
SDFRT_HD bool refine_step(Segment& s, float md) {
    s.its_t += md * (10.f / (10.f + (float)s.i));
    ++s.i;
    return ((md <= 0.f) | (md > s.trace_eps)) & (s.i < 10);
}

struct DiffAcc {
    float warp_t, prev_surf_dist, weight_sum, extra_weight_sum;
    V3 prev_sdf_grad_c, mixed_sum_d, weight_d_sum, extra_weight_sum_d, t_d;
};

// entry face of the bounding box and initial dt/dx (python/shapes.py:142-164)
SDFRT_HD void diff_begin(DiffAcc& a, const Segment& s, const TracerParams& P) {
    a.warp_t = 0.f; a.prev_surf_dist = 0.f; a.weight_sum = 0.f; a.extra_weight_sum = 0.f;
    a.prev_sdf_grad_c = v3(0.f); a.mixed_sum_d = v3(0.f); a.weight_d_sum = v3(0.f); a.extra_weight_sum_d = v3(0.f);
    a.t_d = v3(0.f);
    const V3 q = s.o + s.t * s.d;
    const V3 md = V3{fminf(fabsf(P.bmin[0] - q.x), fabsf(P.bmax[0] - q.x)), fminf(fabsf(P.bmin[1] - q.y), fabsf(P.bmax[1] - q.y)),
                     fminf(fabsf(P.bmin[2] - q.z), fabsf(P.bmax[2] - q.z))};
    V3 n = v3(0.f);
    if ((md.x < md.y) & (md.x < md.z)) n = v3(1, 0, 0);
    if ((md.y < md.z) & (md.y < md.x)) n = v3(0, 1, 0);
    if ((md.z < md.x) & (md.z < md.y)) n = v3(0, 0, 1);
    const float d_dot_n = dot(s.d, n);
    if (!s.inside && fabsf(d_dot_n) > 0.f) a.t_d = (-1.f / d_dot_n * s.t) * n;
}

// one iteration of SDFBase.ray_intersect's loop (python/shapes.py:170-238); (sdf_value, g, H) at o + t d
SDFRT_HD bool diff_step(Segment& s, DiffAcc& a, const TracerParams& P, float sdf_value, V3 g, S3 H) {
    const V3 d = s.d;
    const float t = s.t;
    const V3 x = s.o + t * d;
    if (P.step_scale != 1.f) {
        const float sc = P.step_scale;
        sdf_value *= sc; g = g * sc;
        H.xx *= sc; H.yy *= sc; H.zz *= sc; H.xy *= sc; H.xz *= sc; H.yz *= sc;
    }
    const bool hit = sdf_value < s.trace_eps;
    if (hit) s.its_t = t;
    const float surf_dist = fabsf(sdf_value);
    V3 weight_d;
    float weight = eval_trace_weight(P, d, s.i, x, sdf_value, g, H, weight_d);
    float inv_extra_w_den = 0.f, dist_difference = 0.f;
    if (P.use_extra_weight) {
        inv_extra_w_den = SDFRT_FDIV(1.f, fminf(P.extra_thresh, surf_dist));
        dist_difference = a.prev_surf_dist - surf_dist;
        a.extra_weight_sum += (dist_difference >= 0.f) ? dist_difference * inv_extra_w_den : 0.f;
        a.extra_weight_sum = fminf(a.extra_weight_sum, 1.f);
    }
    const float curr_segment_value = hit ? 0.f : surf_dist;
    const float segment_length = 0.5f * (curr_segment_value + a.prev_surf_dist);
    float weight_increment = segment_length * weight;
    if (P.use_extra_weight) weight_increment *= a.extra_weight_sum;
    a.weight_sum += weight_increment;
    a.warp_t += weight_increment * t;
    // convert_deriv(in_d, t, t_d) = t * in_d + dot(d, in_d) * t_d   (python/shapes.py:126-127)
    weight_d = t * weight_d + dot(d, weight_d) * a.t_d;
    const V3 sdf_grad_c = t * g + dot(d, g) * a.t_d;
    const V3 segment_d = 0.5f * (sdf_grad_c + a.prev_sdf_grad_c);
    if (P.use_extra_weight) {
        const V3 surf_dist_d = sign1(sdf_value) * sdf_grad_c;
        V3 extra_w_d = (a.prev_sdf_grad_c - surf_dist_d) * inv_extra_w_den;
        if (sdf_value < P.extra_thresh) extra_w_d = extra_w_d - (dist_difference * inv_extra_w_den * inv_extra_w_den) * surf_dist_d;
        if (dist_difference > 0.f) a.extra_weight_sum_d += extra_w_d;
        if ((a.extra_weight_sum >= 1.f) | (a.extra_weight_sum <= 0.f)) a.extra_weight_sum_d = v3(0.f);
        weight_d = weight * a.extra_weight_sum_d + weight_d * a.extra_weight_sum;
        weight *= a.extra_weight_sum;
    }
    const V3 weight_increment_d = weight * segment_d + weight_d * segment_length;
    a.mixed_sum_d += weight_increment_d * t + (weight * segment_length) * a.t_d;
    a.t_d += sdf_grad_c;
    a.weight_d_sum += weight_increment_d;
    ++s.i;
    s.t = t + curr_segment_value;
    a.prev_surf_dist = surf_dist;
    a.prev_sdf_grad_c = sdf_grad_c;
    return (s.t <= s.maxt) & !hit & (s.i < P.max_steps);
}

// python/shapes.py:259-283 (its_t is taken from the segment after the refine loop)
SDFRT_HD void diff_finalize(const Segment& s, const DiffAcc& a, const TracerParams& P, TraceOut& out) {
    out.its_t = s.its_t;
    const float inv_weight_sum = 1.f / a.weight_sum;
    float warp_t = a.warp_t * inv_weight_sum;
    V3 warp_t_d = (-warp_t * a.weight_d_sum + a.mixed_sum_d) * inv_weight_sum;
    float warp_weight = 1.f;
    V3 warp_weight_d = v3(0.f);
    if (P.use_weight_sum_weight) {
        warp_weight = fminf(fmaxf(a.weight_sum, 0.f), 1.f);
        warp_weight_d = ((a.weight_sum > 0.f) & (a.weight_sum < 1.f)) ? a.weight_d_sum : v3(0.f);
    }
    const bool invalid = (a.weight_sum < 1e-7f) | !s.intersects | !isfinite(warp_t);
    if (invalid) {
        warp_t = INFINITY;
        warp_t_d = v3(0.f);
        warp_weight = 0.f;
        warp_weight_d = v3(0.f);
    }
    out.warp_t = warp_t; out.warp_t_d = warp_t_d; out.warp_weight = warp_weight; out.warp_weight_d = warp_weight_d;
}

// SDFBase.ray_intersect_non_diff, python/shapes.py:290-339
SDFRT_HD float trace_non_diff(const GridView& G, const TracerParams& P, V3 o, V3 d, float ray_maxt, bool do_refine,
                              int& n_steps, int& n_evals) {
    Segment s;
    bool active = segment_begin(s, P, o, d, ray_maxt);
    while (active) active = value_step(s, P, eval_value(G, s.o + s.t * s.d));
    n_steps = s.i;
    n_evals += s.i;
    if (do_refine && P.refine_intersection && isfinite(s.its_t)) {
        s.i = 0;
        bool refining = true;
        while (refining) refining = refine_step(s, eval_value(G, s.o + s.its_t * s.d));
        n_evals += s.i;
    }
    return s.its_t;
}

// SDFBase.ray_intersect (differentiable tracer), python/shapes.py:115-288
SDFRT_HD void trace_diff(const GridView& G, const TracerParams& P, V3 o, V3 d, float ray_maxt, bool do_refine, TraceOut& out,
                         int& n_evals_hess, int& n_evals_value) {
    Segment s;
    DiffAcc a;
    bool active = segment_begin(s, P, o, d, ray_maxt);
    diff_begin(a, s, P);
    while (active) {
        float v; V3 g; S3 H;
        eval_sdf<2>(G, s.o + s.t * s.d, v, g, H);
        active = diff_step(s, a, P, v, g, H);
    }
    n_evals_hess += s.i;
    out.n_steps = s.i;
    if (do_refine && P.refine_intersection && isfinite(s.its_t)) {
        s.i = 0;
        bool refining = true;
        while (refining) refining = refine_step(s, eval_value(G, s.o + s.its_t * s.d));
        n_evals_value += s.i;
    }
    diff_finalize(s, a, P, out);
}

// ------------------------------------------------------------------------------------
// WarpField2D.weight / weight_epsilon (python/warp.py:25-45)
// ------------------------------------------------------------------------------------
SDFRT_HD float warp_weight_fn(const TracerParams& P, V3 x, float sdf_value, V3 g, float edge_eps, V3& w_d, float& edge_eps_d_out) {
    V3 bbox_dist_d;
    const float bbox_dist = bbox_distance_inside_d(P, x, bbox_dist_d);
    const bool use_edge_eps = edge_eps <= bbox_dist;
    const V3 edge_eps_d = use_edge_eps ? v3(0.f) : bbox_dist_d;
    edge_eps = fminf(edge_eps, bbox_dist);
    const float inv_edge_eps = 1.f / edge_eps;
    const float surf_dist = fabsf(sdf_value);
    const float fac = 1.f - surf_dist * inv_edge_eps;
    const float w = fmaxf(fac, 0.f);
    w_d = (-sign1(sdf_value) * inv_edge_eps) * g + (surf_dist * inv_edge_eps * inv_edge_eps) * edge_eps_d;
    if (!(fac >= 0.f)) w_d = v3(0.f);
    edge_eps_d_out = (use_edge_eps & (fac >= 0.f)) ? surf_dist * inv_edge_eps * inv_edge_eps : 0.f;
    return w;
}

// The warp site: everything WarpField2D.eval (python/warp.py:47-96) contributes to the adjoint.
// With outputs d = d0 <- dir and div = 1 <- tr(P J M), the derivative is linear in
// (dphi, dg) at x = o + warp_t d0:
//    d(div) = cphi_div * dphi + cg_div . dg          d(dir) = cphi_dir * dphi   (vector)
// (derivation in DESIGN.md, "closed-form adjoint").
struct WarpSite {
    bool active;
    float cphi_div;  // d div / d phi
    V3 cg_div;       // d div / d g
    V3 cphi_dir;     // d dir / d phi
    float phi; V3 g; S3 H;  // evaluation at the site (for the chain rule through an attached origin)
};
SDFRT_HD WarpSite warp_site(const GridView& G, const WarpParams& W, const TracerParams& T, V3 o, V3 d0, float tau, V3 tau_d,
                            float omega, V3 omega_d) {
    WarpSite s;
    s.active = false;
    s.cphi_div = 0.f; s.cg_div = v3(0.f); s.cphi_dir = v3(0.f);
    s.phi = 0.f; s.g = v3(0.f); s.H = S3{0, 0, 0, 0, 0, 0};
    if (!isfinite(tau)) return s;
    const V3 x = o + tau * d0;
    float phi; V3 g; S3 H;
    eval_sdf<2>(G, x, phi, g, H);
    s.phi = phi; s.g = g; s.H = H;
    // n' and J_n (python/warp.py:56-63)
    V3 np;
    M3 JnH;  // J_n @ H
    if (W.normalize_warp_field) {
        const float g2 = dot(g, g);
        np = g / g2;
        const float c2 = 2.f / (g2 * g2);
        const V3 Hg = mul(H, g);  // (g g^T H)_ij = g_i (H g)_j   (H symmetric)
        const float Hm[3][3] = {{H.xx, H.xy, H.xz}, {H.xy, H.yy, H.yz}, {H.xz, H.yz, H.zz}};
        const float gv[3] = {g.x, g.y, g.z}, Hgv[3] = {Hg.x, Hg.y, Hg.z};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) JnH.m[i][j] = Hm[i][j] / g2 - c2 * gv[i] * Hgv[j];
    } else {
        np = g;
        JnH.m[0][0] = H.xx; JnH.m[0][1] = H.xy; JnH.m[0][2] = H.xz;
        JnH.m[1][0] = H.xy; JnH.m[1][1] = H.yy; JnH.m[1][2] = H.yz;
        JnH.m[2][0] = H.xz; JnH.m[2][1] = H.yz; JnH.m[2][2] = H.zz;
    }
    // weight of the warp field (python/warp.py:66-76)
    const float eps = (W.weight_strategy == 6) ? W.edge_eps * tau : W.edge_eps;
    V3 w_g; float eps_d;
    float w = warp_weight_fn(T, x, phi, g, eps, w_g, eps_d);
    w_g = w_g + (eps_d * W.edge_eps) * d0;
    if (T.use_weight_sum_weight) {
        w_g = w_g * omega + w * omega_d;
        w *= omega;
    }
    if (!(w > 0.f)) return s;
    // A = M P,  M = I + d0 (x) (tau_d / tau),  P = I - d0 d0^T   (python/warp.py:84-88)
    const V3 q = tau_d / tau;
    const float dv[3] = {d0.x, d0.y, d0.z}, qv[3] = {q.x, q.y, q.z};
    float Pm[3][3], Mm[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            Pm[i][j] = (i == j ? 1.f : 0.f) - dv[i] * dv[j];
            Mm[i][j] = (i == j ? 1.f : 0.f) + dv[i] * qv[j];
        }
    M3 A;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) A.m[i][j] = Mm[i][0] * Pm[0][j] + Mm[i][1] * Pm[1][j] + Mm[i][2] * Pm[2][j];
    const V3 Anp = mul(A, np);
    float trJHA = 0.f;  // tr(J_n H A)
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) trJHA += JnH.m[i][j] * A.m[j][i];
    s.cphi_div = -dot(w_g, Anp) - w * trJHA;
    s.cg_div = (-w) * Anp;
    const float tc = fmaxf(W.clamping_thresh, tau);
    const V3 Pnp = V3{Pm[0][0] * np.x + Pm[0][1] * np.y + Pm[0][2] * np.z, Pm[1][0] * np.x + Pm[1][1] * np.y + Pm[1][2] * np.z,
                      Pm[2][0] * np.x + Pm[2][1] * np.y + Pm[2][2] * np.z};
    s.cphi_dir = (-w / tc) * Pnp;
    s.active = isfinite(s.cphi_div) & isfinite(s.cg_div.x + s.cg_div.y + s.cg_div.z) &
               isfinite(s.cphi_dir.x + s.cphi_dir.y + s.cphi_dir.z);
    return s;
}

// ------------------------------------------------------------------------------------
// Sensor (mitsuba PerspectiveCamera) and film filter, oracle/ref_semantics.md S3-S4
// ------------------------------------------------------------------------------------
SDFRT_HD void sensor_sample_ray(const SensorParams& S, float px, float py, V3& o, V3& d, float& maxt) {
    const float sx = px / (float)S.W, sy = py / (float)S.H;
    V3 np = V3{(1.f - 2.f * sx) * S.tan_half_fov * S.near_clip, (1.f - 2.f * sy) * S.tan_half_fov / S.aspect * S.near_clip, S.near_clip};
    const V3 dl = normalize(np);
    d = V3{S.R[0][0] * dl.x + S.R[0][1] * dl.y + S.R[0][2] * dl.z, S.R[1][0] * dl.x + S.R[1][1] * dl.y + S.R[1][2] * dl.z,
           S.R[2][0] * dl.x + S.R[2][1] * dl.y + S.R[2][2] * dl.z};
    const float inv_z = 1.f / dl.z;
    const float nt = S.near_clip * inv_z;
    o = V3{S.c[0] + d.x * nt, S.c[1] + d.y * nt, S.c[2] + d.z * nt};
    maxt = (S.far_clip - S.near_clip) * inv_z;
}
// sample_direction(it.p = o + d): uv (pixels); `inside` = importance > 0; ref = camera-space point
SDFRT_HD void sensor_project(const SensorParams& S, V3 p, float& u, float& v, bool& inside, V3& ref) {
    const V3 q = V3{p.x - S.c[0], p.y - S.c[1], p.z - S.c[2]};
    ref = V3{S.R[0][0] * q.x + S.R[1][0] * q.y + S.R[2][0] * q.z, S.R[0][1] * q.x + S.R[1][1] * q.y + S.R[2][1] * q.z,
             S.R[0][2] * q.x + S.R[1][2] * q.y + S.R[2][2] * q.z};
    const float sx = 0.5f - 0.5f * ref.x / (ref.z * S.tan_half_fov);
    const float sy = 0.5f - 0.5f * S.aspect * ref.y / (ref.z * S.tan_half_fov);
    u = sx * (float)S.W;
    v = sy * (float)S.H;
    inside = (sx >= 0.f) & (sx <= 1.f) & (sy >= 0.f) & (sy <= 1.f) & (ref.z >= S.near_clip) & (ref.z <= S.far_clip);
}
SDFRT_HD float filter_eval(const SensorParams& S, float x) { return fmaxf(0.f, expf(S.filter_alpha * x * x) - S.filter_bias); }
SDFRT_HD float filter_deriv(const SensorParams& S, float x) {
    const float e = expf(S.filter_alpha * x * x);
    return (e - S.filter_bias > 0.f) ? 2.f * S.filter_alpha * x * e : 0.f;
}

SDFRT_HD V3 square_to_uniform_sphere(float u0, float u1) {
    const float z = 1.f - 2.f * u1;
    const float r = sqrtf(fmaxf(0.f, 1.f - z * z));
    const float phi = 2.f * kPi * u0;
    return V3{r * cosf(phi), r * sinf(phi), z};
}

// ------------------------------------------------------------------------------------
// envmap emitter: lookup (+ derivative w.r.t. the direction), importance sampling, pdf
// ------------------------------------------------------------------------------------
SDFRT_HD V3 env_to_local(const EnvView& E, V3 d) {
    if (!E.has_rot) return d;
    return V3{E.R[0] * d.x + E.R[3] * d.y + E.R[6] * d.z, E.R[1] * d.x + E.R[4] * d.y + E.R[7] * d.z, E.R[2] * d.x + E.R[5] * d.y + E.R[8] * d.z};
}
SDFRT_HD V3 env_to_world(const EnvView& E, V3 d) {
    if (!E.has_rot) return d;
    return V3{E.R[0] * d.x + E.R[1] * d.y + E.R[2] * d.z, E.R[3] * d.x + E.R[4] * d.y + E.R[5] * d.z, E.R[6] * d.x + E.R[7] * d.y + E.R[8] * d.z};
}
// (d: direction in the emitter's local frame)
SDFRT_HD void env_coord(const EnvView& E, V3 d, float& cx, float& cy) {
    float u = atan2f(d.x, -d.z) * (0.5f / kPi) - 0.5f / (float)(E.W1 - 1);
    float v = acosf(fminf(fmaxf(d.y, -1.f), 1.f)) * (1.f / kPi);
    u -= floorf(u);
    v -= floorf(v);
    cx = u * (float)(E.W1 - 1);
    cy = v * (float)(E.H - 1);
}
// radiance from the unit direction d; dLe (optional): dLe[c] = gradient of channel c w.r.t. d
SDFRT_HD V3 env_eval(const EnvView& E, V3 d, V3* dLe) {
    d = env_to_local(E, d);   // v = to_world^-1 (-si.wi)
    float cx, cy;
    env_coord(E, d, cx, cy);
    const int ix = min_i((int)floorf(cx), E.W1 - 2), iy = min_i((int)floorf(cy), E.H - 2);
    const float wx1 = cx - (float)ix, wy1 = cy - (float)iy, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
    const float* r0 = E.data + ((size_t)iy * E.W1 + ix) * 3;
    const float* r1 = r0 + (size_t)E.W1 * 3;
    const V3 v00 = V3{r0[0], r0[1], r0[2]}, v10 = V3{r0[3], r0[4], r0[5]}, v01 = V3{r1[0], r1[1], r1[2]}, v11 = V3{r1[3], r1[4], r1[5]};
    const V3 Le = (wx0 * wy0) * v00 + (wx1 * wy0) * v10 + (wx0 * wy1) * v01 + (wx1 * wy1) * v11;
    if (dLe) {
        const V3 dcx = wy0 * (v10 - v00) + wy1 * (v11 - v01);   // d Le / d cx
        const V3 dcy = wx0 * (v01 - v00) + wx1 * (v11 - v10);   // d Le / d cy
        const float r2 = d.x * d.x + d.z * d.z;
        const float kx = r2 > 0.f ? (float)(E.W1 - 1) * (0.5f / kPi) / r2 : 0.f;
        const V3 gx = V3{-d.z * kx, 0.f, d.x * kx};             // grad cx
        const float s2 = 1.f - d.y * d.y;
        const float ky = s2 > 0.f ? -(float)(E.H - 1) * (1.f / kPi) / sqrtf(s2) : 0.f;
        const V3 gy = V3{0.f, ky, 0.f};                         // grad cy
        dLe[0] = env_to_world(E, dcx.x * gx + dcy.x * gy);   // d/d(d_world) = R d/d(d_local)
        dLe[1] = env_to_world(E, dcx.y * gx + dcy.y * gy);
        dLe[2] = env_to_world(E, dcx.z * gx + dcy.z * gy);
    }
    return Le;
}
SDFRT_HD float env_pdf(const EnvView& E, V3 d) {
    d = env_to_local(E, d);
    float cx, cy;
    env_coord(E, d, cx, cy);
    const int ix = min_i((int)floorf(cx), E.W1 - 2), iy = min_i((int)floorf(cy), E.H - 2);
    const float sin_t = sqrtf(fmaxf(0.f, 1.f - d.y * d.y));
    if (!(sin_t > 0.f)) return 0.f;
    return E.cell_prob[(size_t)iy * (E.W1 - 1) + ix] * (float)((E.W1 - 1) * (E.H - 1)) / (2.f * kPi * kPi * sin_t);
}
// first index with cdf[i] > u  (inclusive CDF, n entries), clamped to n - 1
SDFRT_HD int cdf_search(const float* cdf, int n, float u) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cdf[mid] > u) hi = mid; else lo = mid + 1;
    }
    return lo;
}
SDFRT_HD V3 env_sample(const EnvView& E, float u0, float u1, float& pdf) {
    const int nx = E.W1 - 1, ny = E.H - 1;
    const int iy = cdf_search(E.marg_cdf, ny, u1);
    const float ylo = iy > 0 ? E.marg_cdf[iy - 1] : 0.f;
    const float fy = fminf(fmaxf((u1 - ylo) / fmaxf(E.marg_cdf[iy] - ylo, 1e-30f), 0.f), 1.f);
    const float* row = E.cond_cdf + (size_t)iy * nx;
    const int ix = cdf_search(row, nx, u0);
    const float xlo = ix > 0 ? row[ix - 1] : 0.f;
    const float fx = fminf(fmaxf((u0 - xlo) / fmaxf(row[ix] - xlo, 1e-30f), 0.f), 1.f);
    const float uu = ((float)ix + fx) / (float)nx + 0.5f / (float)nx;
    const float vv = ((float)iy + fy) / (float)ny;
    const float theta = vv * kPi, phi = uu * 2.f * kPi;
    const float st = sinf(theta);
    // pdf from the sampled cell and sin(theta) themselves (re-deriving them from the rotated direction loses digits near the poles)
    pdf = st > 0.f ? E.cell_prob[(size_t)iy * nx + ix] * (float)(nx * ny) / (2.f * kPi * kPi * st) : 0.f;
    return env_to_world(E, V3{sinf(phi) * st, cosf(theta), -cosf(phi) * st});
}

// diffuse reflectance: constant or trilinear gridvolume on [0,1]^3 (clamp); optional gradient wrt p
struct AlbedoTaps {
    int idx[8];
    float w[8];
};
SDFRT_HD V3 eval_albedo(const AlbedoView& A, V3 p, AlbedoTaps* taps, V3 dwd[3] /* d rho_c / d p */) {
    if (A.data == nullptr) {
        if (dwd) { dwd[0] = dwd[1] = dwd[2] = v3(0.f); }
        return V3{A.cr, A.cg, A.cb};
    }
    const float ux = p.x * (float)A.rx - 0.5f, uy = p.y * (float)A.ry - 0.5f, uz = p.z * (float)A.rz - 0.5f;
    float fx = floorf(ux), fy = floorf(uy), fz = floorf(uz);
    const float ax = ux - fx, ay = uy - fy, az = uz - fz;
    fx = fminf(fmaxf(fx, -2.f), (float)A.rx + 2.f);
    fy = fminf(fmaxf(fy, -2.f), (float)A.ry + 2.f);
    fz = fminf(fmaxf(fz, -2.f), (float)A.rz + 2.f);
    const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
    V3 rho = v3(0.f);
    V3 gx = v3(0.f), gy = v3(0.f), gz = v3(0.f);  // per-channel derivative wrt (x, y, z)
    for (int k = 0; k < 8; ++k) {
        const int bx = k & 1, by = (k >> 1) & 1, bz = (k >> 2) & 1;
        const int xx = clampi(ix + bx, 0, A.rx - 1), yy = clampi(iy + by, 0, A.ry - 1), zz = clampi(iz + bz, 0, A.rz - 1);
        const float wx = bx ? ax : 1.f - ax, wy = by ? ay : 1.f - ay, wz = bz ? az : 1.f - az;
        const int lin = ((zz * A.ry + yy) * A.rx + xx);
        const float* t = A.data + (size_t)lin * 3;
        const V3 c = V3{t[0], t[1], t[2]};
        const float w = wx * wy * wz;
        rho += w * c;
        if (taps) { taps->idx[k] = lin; taps->w[k] = w; }
        if (dwd) {
            gx += ((bx ? 1.f : -1.f) * wy * wz * (float)A.rx) * c;
            gy += ((by ? 1.f : -1.f) * wx * wz * (float)A.ry) * c;
            gz += ((bz ? 1.f : -1.f) * wx * wy * (float)A.rz) * c;
        }
    }
    if (dwd) {  // dwd[c] = gradient of channel c wrt p
        dwd[0] = V3{gx.x, gy.x, gz.x};
        dwd[1] = V3{gx.y, gy.y, gz.y};
        dwd[2] = V3{gx.z, gy.z, gz.z};
    }
    return rho;
}

// ------------------------------------------------------------------------------------
// Conservative empty-space test: per-brick SDF minima + 3-D DDA (rationale in sdfrt_march.cuh)
// ------------------------------------------------------------------------------------
constexpr int kBrick = 4;  // brick edge in texels
struct BrickView {
    const float* data;  // (nz, ny, nx) brick minima dilated by `dilation` bricks, or nullptr
    int nx, ny, nz;
};

// 3-D DDA over the brick lattice in texel coordinates c = x * R - 0.5 (cells of kBrick texels; indices clamped on read):
// true iff every brick pierced by o + t d, t in [t0, t1], has a (dilated) minimum > thresh.
SDFRT_HD bool ray_provably_clear(const BrickView& B, const GridView& G, V3 o, V3 d, float t0, float t1, float thresh) {
    if (B.data == nullptr || !(t1 >= t0)) return false;
    const float inv = 1.f / (float)kBrick;
    const float cx = ((o.x - G.px) * (float)G.rx - 0.5f) * inv, cy = ((o.y - G.py) * (float)G.ry - 0.5f) * inv,
                cz = ((o.z - G.pz) * (float)G.rz - 0.5f) * inv;
    const float dx = d.x * (float)G.rx * inv, dy = d.y * (float)G.ry * inv, dz = d.z * (float)G.rz * inv;
    const float px = cx + t0 * dx, py = cy + t0 * dy, pz = cz + t0 * dz;
    int ix = (int)floorf(px), iy = (int)floorf(py), iz = (int)floorf(pz);
    const int sx = dx > 0.f ? 1 : -1, sy = dy > 0.f ? 1 : -1, sz = dz > 0.f ? 1 : -1;
    const float idx = dx != 0.f ? 1.f / fabsf(dx) : INFINITY, idy = dy != 0.f ? 1.f / fabsf(dy) : INFINITY,
                idz = dz != 0.f ? 1.f / fabsf(dz) : INFINITY;
    float tx = dx != 0.f ? t0 + ((dx > 0.f ? (float)(ix + 1) - px : px - (float)ix)) * idx : INFINITY;
    float ty = dy != 0.f ? t0 + ((dy > 0.f ? (float)(iy + 1) - py : py - (float)iy)) * idy : INFINITY;
    float tz = dz != 0.f ? t0 + ((dz > 0.f ? (float)(iz + 1) - pz : pz - (float)iz)) * idz : INFINITY;
    for (int it = 0; it < 8192; ++it) {
        const int qx = clampi(ix, 0, B.nx - 1), qy = clampi(iy, 0, B.ny - 1), qz = clampi(iz, 0, B.nz - 1);
        if (!(B.data[(qz * B.ny + qy) * B.nx + qx] > thresh)) return false;
        const float tn = fminf(tx, fminf(ty, tz));
        if (tn > t1) return true;
        if (tx <= ty && tx <= tz) { ix += sx; tx += idx; }
        else if (ty <= tz) { iy += sy; ty += idy; }
        else { iz += sz; tz += idz; }
    }
    return false;
}

// The test is made once per PIXEL for the ray through the pixel centre.  Every sample ray of the pixel passes through the
// camera apex within the angle `half_diag` of the centre ray, so each of its points y has a centre-line point y' with
// |y - y'| <= delta = apex_dist_max * half_diag.  The base texel of y is then within ceil(delta R) + 1 texels of that of
// y', and its 4 taps per axis within the brick of y' dilated by `dilation` = ceil((2 + ceil(delta R) + 1) / kBrick) bricks.
struct EmptySpaceParams {
    float delta;          // max distance between a sample-ray point and the centre line
    float maxt_bound;     // upper bound of any segment's maxt (scales trace_eps / the warp epsilon)
    float skip_eps;       // 0 (value pass) or the warp's edge_eps (DIFF pass, weight strategy 6); < 0 disables the test
    int dilation;         // bricks
};
SDFRT_HD EmptySpaceParams empty_space_params(const SensorParams& S, const GridView& G, float skip_eps) {
    EmptySpaceParams e;
    const float ax = S.tan_half_fov / (float)S.W, ay = S.tan_half_fov / (S.aspect * (float)S.H);
    const float half_diag = sqrtf(ax * ax + ay * ay);
    const float cx = S.c[0] - 0.5f, cy = S.c[1] - 0.5f, cz = S.c[2] - 0.5f;
    const float apex_dist_max = sqrtf(cx * cx + cy * cy + cz * cz) + 0.5f * 1.7320508f * (1.f + 2.f * kBBoxDelta) + 0.05f;
    e.delta = apex_dist_max * half_diag * 1.05f;
    e.maxt_bound = apex_dist_max;
    e.skip_eps = skip_eps;
    const float rmax = (float)(G.rx > G.ry ? (G.rx > G.rz ? G.rx : G.rz) : (G.ry > G.rz ? G.ry : G.rz));
    const int k = (int)ceilf(e.delta * rmax) + 1;
    e.dilation = (2 + k + kBrick - 1) / kBrick;
    if (e.dilation > 4) e.skip_eps = -1.f;  // pixels too coarse relative to the grid: the bound would be useless
    return e;
}
SDFRT_HD float empty_space_threshold(const TracerParams& T, const EmptySpaceParams& e) {
    return fmaxf(T.trace_eps, e.skip_eps) * fmaxf(e.maxt_bound, 1.f) + 1e-4f;
}
SDFRT_HD bool pixel_provably_clear(const SensorParams& S, const GridView& G, const TracerParams& T, const BrickView& B,
                                   const EmptySpaceParams& e, int px, int py) {
    if (e.skip_eps < 0.f) return false;
    V3 o, d; float maxt;
    sensor_sample_ray(S, (float)(px - S.border) + 0.5f, (float)(py - S.border) + 0.5f, o, d, maxt);
    d = normalize(d);
    // centre line clipped against the bounding box inflated by delta (slab test as ray_box, python/shapes.py:129-131)
    const float lo = -kBBoxDelta - e.delta, hi = 1.f + kBBoxDelta + e.delta;
    const V3 r = V3{1.f / d.x, 1.f / d.y, 1.f / d.z};
    const V3 t1 = V3{(lo - o.x) * r.x, (lo - o.y) * r.y, (lo - o.z) * r.z};
    const V3 t2 = V3{(hi - o.x) * r.x, (hi - o.y) * r.y, (hi - o.z) * r.z};
    float tn = fmaxf(fminf(t1.x, t2.x), fmaxf(fminf(t1.y, t2.y), fminf(t1.z, t2.z)));
    float tf = fminf(fmaxf(t1.x, t2.x), fminf(fmaxf(t1.y, t2.y), fmaxf(t1.z, t2.z)));
    if (!(tf >= tn)) return true;  // the whole pixel misses the bounding box: nothing to march anyway
    tn = fmaxf(tn - e.delta, -e.delta);
    tf = tf + e.delta;
    return ray_provably_clear(B, G, o, d, tn, tf, empty_space_threshold(T, e));
}

}  // namespace sdfrt
