// sdfrt_lane.cuh -- one wavefront lane (= one camera sample) of the reparameterised integrators.
//
//   lane_forward<DIFF>   ReparamIntegrator.render / eval_sample (python/integrators/reparam.py:82-185)
//                        + <integrator>.sample; DIFF selects the differentiable tracer and writes the
//                        per-lane tracer record that the adjoint consumes.
//   lane_backward        closed-form adjoint of the same lane (what dr.backward_from(image * grad_in)
//                        produces through Dr.Jit's AD in python/integrators/reparam.py:187-190),
//                        scattering dL/d(sdf.data) with 64-tap B-spline weights.
// See DESIGN.md for the derivation.
#pragma once
#include "sdfrt_core.cuh"

namespace sdfrt {

enum { INTEGRATOR_SILHOUETTE = 0, INTEGRATOR_SIMPLE_SHADING = 1, INTEGRATOR_DIRECT = 2, INTEGRATOR_PRB = 3 };
enum { WARP_NONE = 0, WARP_DUMMY = 1, WARP_2D = 2, WARP_CONV = 3 };

struct PrbParams {  // props of SdfPrbReparamIntegrator (python/integrators/sdf_prb_reparam.py:35-48, reparam.py:14)
    int max_depth, rr_depth;
};

struct SceneParams {
    GridView G;
    AlbedoView A;
    TracerParams T;
    WarpParams W;
    SensorParams S;
    ShadingParams Sh;
    int integrator;
    PrbParams prb;
    EnvView E;   // `envmap` emitter (data == nullptr: the constant emitter of Sh)
};

struct PassParams {
    uint32_t seed;
    int spp;
    int log2_spp;               // >= 0 if spp is a power of two (python/integrators/reparam.py:144-148)
    // Multi-GPU partition: the film rows are cut into sub-bands of `assign_rows` rows (1, 2, 4 or 8); rank r owns the sub-bands
    // r, r + world, ... (interleaved, so that every rank sees a similar mix of object and background), `rows_local` rows in all.
    // A rank walks ITS rows in tiles of kBandRows rows, whatever the assignment granularity.  `n_items` = camera samples of this rank.
    int rank, world;
    int assign_rows, rows_local;
    int Wb, Hb;                 // film size incl. border
    unsigned long long n_items;
    // antithetic_sampling (python/integrators/reparam.py:19,167-178): every wavefront lane is evaluated twice, the second time
    // at the mirrored film position pos - r + 1 with a clone of the sampler; work items 2l and 2l + 1 are the two samples
    int antithetic;
    // validation hook (sdfrt_set_lane_window): adjoint sweep 2 only back-propagates the wavefront lanes in [win_lo, win_hi)
    unsigned long long win_lo, win_hi;
};

constexpr unsigned long long kAntiBit = 1ull << 63;  // WorkItem::lane carries the "mirrored sample" flag in bit 63

constexpr int kBandRows = 8;

// Work item k of this rank -> pixel and wavefront lane.  Items walk the rank's rows in tiles of kBandRows rows, each tile column by column
// (8 rows x spp samples per column), so neighbouring items are neighbouring samples; lane indices -- and with them the
// sampler streams -- stay the reference's scanline wavefront indices (python/integrators/reparam.py:141-155).
struct WorkItem {
    int px, py;
    unsigned long long lane;  // wavefront lane index | kAntiBit for the mirrored sample of an antithetic pair
};
SDFRT_HD WorkItem decode_item(const PassParams& ps, unsigned long long k) {
    const unsigned long long anti = ps.antithetic ? (k & 1ull) : 0ull;
    if (ps.antithetic) k >>= 1;
    const unsigned Wb = (unsigned)ps.Wb;
    const unsigned long long per_tile = (unsigned long long)kBandRows * Wb * (unsigned)ps.spp;  // < 2^32 (checked on the host)
    const unsigned bl = (unsigned)(k / per_tile);
    const unsigned r = (unsigned)(k - (unsigned long long)bl * per_tile);
    const unsigned row_l0 = bl * (unsigned)kBandRows;                  // first local row of the tile
    const unsigned left = (unsigned)ps.rows_local - row_l0;
    const unsigned rows_here = left < (unsigned)kBandRows ? left : (unsigned)kBandRows;
    const unsigned per_col = rows_here * (unsigned)ps.spp;
    const unsigned x = r / per_col;
    const unsigned rr = r - x * per_col;
    const unsigned y_in = ps.log2_spp >= 0 ? (rr >> ps.log2_spp) : (rr / (unsigned)ps.spp);
    const unsigned sidx = rr - y_in * (unsigned)ps.spp;
    // local row -> film row: sub-band (j / A) of this rank is the global sub-band (j / A) * world + rank
    const unsigned j = row_l0 + y_in, A = (unsigned)ps.assign_rows;
    const unsigned sb = j / A;
    const unsigned row = (sb * (unsigned)ps.world + (unsigned)ps.rank) * A + (j - sb * A);
    WorkItem w;
    w.px = (int)x;
    w.py = (int)row;
    w.lane = (((unsigned long long)w.py * Wb + x) * (unsigned)ps.spp + sidx) | (anti ? kAntiBit : 0ull);
    return w;
}

// per-lane record written by lane_forward<true>, SoA with stride = number of local lanes
enum {
    REC_ITS_T = 0, REC_TAU, REC_TAU_DX, REC_TAU_DY, REC_TAU_DZ, REC_OMEGA, REC_OMEGA_DX, REC_OMEGA_DY, REC_OMEGA_DZ,
    REC_S_STATE,  // shadow ray: 0 = no NEE contribution, 1 = unoccluded and gates passed
    REC_S_TAU, REC_S_TAU_DX, REC_S_TAU_DY, REC_S_TAU_DZ, REC_S_OMEGA, REC_S_OMEGA_DX, REC_S_OMEGA_DY, REC_S_OMEGA_DZ,
    REC_COUNT,         // rows of a record without the BSDF-sampled ray
    // use_mis: the BSDF-sampled ray (python/integrators/sdf_direct_reparam.py:87-107): 0 = no contribution, 1 = escaped to the emitter
    REC_B_STATE = REC_COUNT,
    REC_B_TAU, REC_B_TAU_DX, REC_B_TAU_DY, REC_B_TAU_DZ, REC_B_OMEGA, REC_B_OMEGA_DX, REC_B_OMEGA_DY, REC_B_OMEGA_DZ,
    REC_COUNT_MIS
};

struct LaneStats {
    unsigned n_value, n_grad, n_hess, n_sites, n_albedo, n_steps_max, n_hit;
};

#if defined(__CUDA_ARCH__)
#define SDFRT_ATOMIC_ADD(ptr, val) atomicAdd((ptr), (val))
#elif !defined(SDFRT_ATOMIC_ADD)
#define SDFRT_ATOMIC_ADD(ptr, val) (*(ptr) += (val))  // host builds (test harness): single-threaded unless overridden
#endif

// WarpFieldConvolution.ray_intersect: `reparam and (max_reparam_depth < 0 or depth <= max_reparam_depth)` (python/warp.py:142)
SDFRT_HD bool conv_at_depth(const WarpParams& W, int depth) {
    return W.kind == WARP_CONV && (W.max_reparam_depth < 0 || depth <= W.max_reparam_depth);
}
SDFRT_HD bool reparam_at_depth(const WarpParams& W, int depth) {
    return W.kind == WARP_2D && (W.max_reparam_depth < 0 || depth <= W.max_reparam_depth);
}

struct CameraSample {
    V3 o, d;
    float maxt;
    Pcg32 rng;
};
// film position of a sample from its pixel and jitter; the mirrored sample of an antithetic pair sits at pos - r + 1
// (python/integrators/reparam.py:171-174), evaluated in that order
SDFRT_HD float sample_position(int pix, float r, bool anti) { return anti ? ((float)pix - r) + 1.f : (float)pix + r; }

SDFRT_HD CameraSample camera_sample(const SceneParams& sc, const PassParams& ps, unsigned long long lane) {
    const bool anti = (lane & kAntiBit) != 0ull;
    lane &= ~kAntiBit;
    const int b = sc.S.border;
    const int Wb = sc.S.W + 2 * b;
    const unsigned long long pix = ps.log2_spp >= 0 ? (lane >> ps.log2_spp) : (lane / (unsigned long long)ps.spp);
    const int py = (int)(pix / (unsigned long long)Wb);
    const int px = (int)(pix - (unsigned long long)py * Wb);
    CameraSample cs;
    cs.rng = sampler_seed(ps.seed, (uint32_t)lane);
    const float r0 = pcg_next_f32(cs.rng), r1 = pcg_next_f32(cs.rng);
    (void)pcg_next_f32(cs.rng);  // wavelength sample, consumed even in RGB (python/integrators/reparam.py:90)
    sensor_sample_ray(sc.S, sample_position(px - b, r0, anti), sample_position(py - b, r1, anti), cs.o, cs.d, cs.maxt);
    return cs;
}

// shadow ray of the emitter sample (python/integrators/sdf_direct_reparam.py:38-51 + spawn_ray_to)
struct ShadowRay {
    V3 o, d;
    float maxt;
};
SDFRT_HD ShadowRay make_shadow_ray_dir(const ShadingParams& Sh, V3 p, V3 n, V3 d_e);
SDFRT_HD ShadowRay make_shadow_ray(const ShadingParams& Sh, V3 p, V3 n, float u0, float u1) {
    return make_shadow_ray_dir(Sh, p, n, square_to_uniform_sphere(u0, u1));
}
// emitter sample of the direct integrator (scene.sample_emitter_direction, python/integrators/sdf_direct_reparam.py:40): uniform sphere
// for the constant emitter (pdf 1 / 4 pi), importance sampling of the envmap otherwise
SDFRT_HD ShadowRay sample_emitter_ray(const SceneParams& sc, V3 p, V3 n, float u0, float u1, float& pdf) {
    if (sc.E.data == nullptr) {
        pdf = 0.07957747154594767f;
        return make_shadow_ray(sc.Sh, p, n, u0, u1);
    }
    const V3 d_e = env_sample(sc.E, u0, u1, pdf);
    return make_shadow_ray_dir(sc.Sh, p, n, d_e);
}
SDFRT_HD ShadowRay make_shadow_ray_dir(const ShadingParams& Sh, V3 p, V3 n, V3 d_e) {
    const V3 ds_p = p + Sh.env_dist * d_e;
    const float mag = (1.f + fmaxf(fabsf(p.x), fmaxf(fabsf(p.y), fabsf(p.z)))) * kRayEpsilon;
    const float sgn = sign1(dot(n, ds_p - p));
    ShadowRay s;
    s.o = p + (mag * sgn) * n;
    const V3 dv = ds_p - s.o;
    const float dist = sqrtf(dot(dv, dv));
    s.d = dv / dist;
    s.maxt = dist * (1.f - kShadowEpsilon);
    return s;
}

constexpr float kInv4Pi = 0.07957747154594767f;

// mitsuba `coordinate_system` (Duff et al. 2017) -- the frame initialize_sh_frame builds when dp_du == 0
// (python/shapes.py:360); s = first vector, t = cross(n, s).  oracle/ref_semantics.md S14
SDFRT_HD void sh_frame(V3 n, V3& s, V3& t) {
    const float sign = n.z >= 0.f ? 1.f : -1.f;
    const float a = -1.f / (sign + n.z);
    const float b = n.x * n.y * a;
    s = V3{sign * n.x * n.x * a + 1.f, sign * b, -sign * n.x};
    t = V3{n.y * s.z - n.z * s.y, n.z * s.x - n.x * s.z, n.x * s.y - n.y * s.x};
}
// warp::square_to_cosine_hemisphere through the concentric disk mapping.  S15
SDFRT_HD void square_to_uniform_disk_concentric(float u0, float u1, float& px, float& py) {
    const float x = 2.f * u0 - 1.f, y = 2.f * u1 - 1.f;
    const bool is_zero = (x == 0.f) & (y == 0.f);
    const bool q13 = fabsf(x) < fabsf(y);
    const float r = q13 ? y : x, rp = q13 ? x : y;
    float phi = 0.25f * kPi * rp / (is_zero ? 1.f : r);
    if (q13) phi = 0.5f * kPi - phi;
    if (is_zero) phi = 0.f;
    px = r * cosf(phi); py = r * sinf(phi);
}
SDFRT_HD V3 square_to_cosine_hemisphere(float u0, float u1) {
    float px, py;
    square_to_uniform_disk_concentric(u0, u1, px, py);
    return V3{px, py, sqrtf(fmaxf(0.f, 1.f - (px * px + py * py)))};
}
// mitsuba.ad.integrators.common.mis_weight (power heuristic, 0 where undefined)
SDFRT_HD float mis_weight(float pdf_a, float pdf_b) {
    const float a2 = pdf_a * pdf_a, den = a2 + pdf_b * pdf_b;
    return den > 0.f ? a2 / den : 0.f;
}
// Interaction.offset_p (spawn_ray): p + n * mulsign((1 + max|p|) * RayEpsilon, n . d).  S6
SDFRT_HD V3 spawn_offset(V3 p, V3 n, V3 d) {
    const float mag = (1.f + fmaxf(fabsf(p.x), fmaxf(fabsf(p.y), fabsf(p.z)))) * kRayEpsilon;
    return p + (mag * sign1(dot(n, d))) * n;
}

// BSDF sample of the MIS branch of sdf_direct_reparam (python/integrators/sdf_direct_reparam.py:87-94): cosine-hemisphere direction
// in the shading frame of the detached interaction, ray spawned from the attached interaction with a detached direction.
struct BsdfRay {
    bool active;   // cos_i > 0 and pdf > 0
    V3 o, d;
    float pdf;     // cos / pi
};
SDFRT_HD BsdfRay make_bsdf_ray(V3 p, V3 n, float cos_i, float u0, float u1) {
    BsdfRay r;
    const V3 wo = square_to_cosine_hemisphere(u0, u1);
    r.pdf = wo.z * (1.f / kPi);
    r.active = (cos_i > 0.f) & (r.pdf > 0.f);
    V3 fs, ft;
    sh_frame(n, fs, ft);
    r.d = fs * wo.x + ft * wo.y + n * wo.z;
    r.o = spawn_offset(p, n, r.d);
    return r;
}
// weights of the two strategies for the constant emitter (pdf 1 / 4 pi) and the diffuse BSDF (pdf cos / pi); both detached
SDFRT_HD float mis_weight_nee(float cos_o) { return mis_weight(kInv4Pi, cos_o * (1.f / kPi)); }
SDFRT_HD float mis_weight_bsdf(float pdf) { return mis_weight(pdf, kInv4Pi); }

// The three places radiance enters sdf_direct_reparam, for the constant emitter (Sh.Le) or the envmap (sc.E):
// an escaped camera ray (python/integrators/sdf_direct_reparam.py:31), the emitter sample (:68-80; emitter_val / ds.pdf, optional
// power-heuristic weight against the detached BSDF pdf) and an escaped BSDF-sampled ray (:105-107; bsdf_val / bs.pdf = rho).
SDFRT_HD V3 miss_rgb(const SceneParams& sc, V3 d) {
    if (sc.E.data == nullptr) return V3{sc.Sh.Le[0], sc.Sh.Le[1], sc.Sh.Le[2]};
    return env_eval(sc.E, d, nullptr);
}
SDFRT_HD V3 nee_rgb(const SceneParams& sc, V3 rho, float cos_o, V3 d_s, float pdf, bool mis) {
    if (sc.E.data == nullptr) {
        float k = cos_o * 4.f;  // (rho / pi) * cos_o * Le / (1 / (4 pi))
        if (mis) k *= mis_weight_nee(cos_o);
        return V3{rho.x * k * sc.Sh.Le[0], rho.y * k * sc.Sh.Le[1], rho.z * k * sc.Sh.Le[2]};
    }
    float k = cos_o * (1.f / kPi) / pdf;
    if (mis) k *= mis_weight(pdf, cos_o * (1.f / kPi));
    const V3 Le = env_eval(sc.E, d_s, nullptr);
    return V3{rho.x * k * Le.x, rho.y * k * Le.y, rho.z * k * Le.z};
}
SDFRT_HD V3 bsdf_rgb(const SceneParams& sc, V3 rho, V3 d_b, float bs_pdf) {
    if (sc.E.data == nullptr) {
        const float wb = mis_weight_bsdf(bs_pdf);
        return V3{rho.x * wb * sc.Sh.Le[0], rho.y * wb * sc.Sh.Le[1], rho.z * wb * sc.Sh.Le[2]};
    }
    const float wb = mis_weight(bs_pdf, env_pdf(sc.E, d_b));
    const V3 Le = env_eval(sc.E, d_b, nullptr);
    return V3{rho.x * wb * Le.x, rho.y * wb * Le.y, rho.z * wb * Le.z};
}

// ------------------------------------------------------------------------------------------------
// forward lane
// ------------------------------------------------------------------------------------------------
template <bool DIFF>
SDFRT_HD void lane_forward(const SceneParams& sc, const PassParams& ps, unsigned long long lane, float* film,
                           float* rec /* already offset to this sample's record */, unsigned long long rec_stride, LaneStats& st) {
    CameraSample cs = camera_sample(sc, ps, lane);
    const V3 o = cs.o, d0 = cs.d;
    int nv = 0, nh = 0, nsteps = 0;
    TraceOut tr;
    tr.warp_t = INFINITY; tr.warp_t_d = v3(0.f); tr.warp_weight = 0.f; tr.warp_weight_d = v3(0.f);
    if (DIFF && reparam_at_depth(sc.W, 0)) {
        trace_diff(sc.G, sc.T, o, d0, cs.maxt, true, tr, nh, nv);
        nsteps = tr.n_steps;
    } else {
        tr.its_t = trace_non_diff(sc.G, sc.T, o, d0, cs.maxt, true, nsteps, nv);
    }
    const bool hit = isfinite(tr.its_t);
    V3 rgb = v3(0.f);
    float s_state = 0.f, b_state = 0.f;
    TraceOut ts, tb;
    ts.warp_t = INFINITY; ts.warp_t_d = v3(0.f); ts.warp_weight = 0.f; ts.warp_weight_d = v3(0.f);
    tb = ts;
    if (sc.integrator == INTEGRATOR_SILHOUETTE) {
        rgb = v3(hit ? 1.f : 0.f);
    } else if (sc.integrator == INTEGRATOR_SIMPLE_SHADING) {
        if (hit) {
            float phi; V3 g; S3 H;
            eval_sdf<1>(sc.G, o + tr.its_t * d0, phi, g, H);
            st.n_grad += 1;
            const V3 n = normalize(g);
            rgb = v3(fmaxf(0.f, (n.x + n.y + n.z) * 0.57735026918962576f));
        }
    } else {
        if (!hit) {
            if (!sc.Sh.hide_emitters) rgb = miss_rgb(sc, d0);
        } else {
            const V3 p = o + tr.its_t * d0;
            float phi; V3 g; S3 H;
            eval_sdf<1>(sc.G, p, phi, g, H);
            st.n_grad += 1;
            const V3 n = normalize(g);
            const float cos_i = -dot(n, d0);
            const float u0 = pcg_next_f32(cs.rng), u1 = pcg_next_f32(cs.rng);
            float pdf_e;
            const ShadowRay sr = sample_emitter_ray(sc, p, n, u0, u1, pdf_e);
            const float cos_o = dot(n, sr.d);
            // the reference marches every shadow ray; a lane whose BSDF gate is closed contributes
            // exactly zero whatever the march returns, so it is skipped here.
            if ((cos_i > 0.f) & (cos_o > 0.f) & (pdf_e > 0.f)) {
                float its_s;
                int ns = 0;
                if (DIFF && reparam_at_depth(sc.W, 1)) {
                    trace_diff(sc.G, sc.T, sr.o, sr.d, sr.maxt, false, ts, nh, nv);
                    its_s = ts.its_t;
                    ns = ts.n_steps;
                } else {
                    its_s = trace_non_diff(sc.G, sc.T, sr.o, sr.d, sr.maxt, false, ns, nv);
                }
                nsteps = nsteps > ns ? nsteps : ns;
                if (!isfinite(its_s)) {
                    s_state = 1.f;
                    const V3 rho = eval_albedo(sc.A, p, nullptr, nullptr);
                    st.n_albedo += (sc.A.data != nullptr);
                    rgb = nee_rgb(sc, rho, cos_o, sr.d, pdf_e, sc.Sh.use_mis != 0);
                }
            }
            if (sc.Sh.use_mis) {   // BSDF sampling (python/integrators/sdf_direct_reparam.py:87-107)
                (void)pcg_next_f32(cs.rng);  // component sample
                const float b0 = pcg_next_f32(cs.rng), b1 = pcg_next_f32(cs.rng);
                const BsdfRay br = make_bsdf_ray(p, n, cos_i, b0, b1);
                if (br.active) {
                    float its_b;
                    int nb = 0;
                    if (DIFF && reparam_at_depth(sc.W, 1)) {
                        trace_diff(sc.G, sc.T, br.o, br.d, INFINITY, false, tb, nh, nv);
                        its_b = tb.its_t;
                        nb = tb.n_steps;
                    } else {
                        its_b = trace_non_diff(sc.G, sc.T, br.o, br.d, INFINITY, false, nb, nv);
                    }
                    nsteps = nsteps > nb ? nsteps : nb;
                    if (!isfinite(its_b)) {   // escaped: the environment emitter; bsdf_val / pdf = rho
                        b_state = 1.f;
                        const V3 rho = eval_albedo(sc.A, p, nullptr, nullptr);
                        st.n_albedo += (sc.A.data != nullptr);
                        rgb += bsdf_rgb(sc, rho, br.d, br.pdf);
                    }
                }
            }
        }
    }
    st.n_value += nv;
    st.n_hess += nh;
    st.n_hit += hit;
    st.n_steps_max = st.n_steps_max > (unsigned)nsteps ? st.n_steps_max : (unsigned)nsteps;

    if (DIFF) {
        float* r = rec;
        r[REC_ITS_T * rec_stride] = tr.its_t;
        r[REC_TAU * rec_stride] = tr.warp_t;
        r[REC_TAU_DX * rec_stride] = tr.warp_t_d.x; r[REC_TAU_DY * rec_stride] = tr.warp_t_d.y; r[REC_TAU_DZ * rec_stride] = tr.warp_t_d.z;
        r[REC_OMEGA * rec_stride] = tr.warp_weight;
        r[REC_OMEGA_DX * rec_stride] = tr.warp_weight_d.x; r[REC_OMEGA_DY * rec_stride] = tr.warp_weight_d.y; r[REC_OMEGA_DZ * rec_stride] = tr.warp_weight_d.z;
        r[REC_S_STATE * rec_stride] = s_state;
        r[REC_S_TAU * rec_stride] = ts.warp_t;
        r[REC_S_TAU_DX * rec_stride] = ts.warp_t_d.x; r[REC_S_TAU_DY * rec_stride] = ts.warp_t_d.y; r[REC_S_TAU_DZ * rec_stride] = ts.warp_t_d.z;
        r[REC_S_OMEGA * rec_stride] = ts.warp_weight;
        r[REC_S_OMEGA_DX * rec_stride] = ts.warp_weight_d.x; r[REC_S_OMEGA_DY * rec_stride] = ts.warp_weight_d.y; r[REC_S_OMEGA_DZ * rec_stride] = ts.warp_weight_d.z;
        if (sc.Sh.use_mis && sc.integrator == INTEGRATOR_DIRECT) {
            r[REC_B_STATE * rec_stride] = b_state;
            r[REC_B_TAU * rec_stride] = tb.warp_t;
            r[REC_B_TAU_DX * rec_stride] = tb.warp_t_d.x; r[REC_B_TAU_DY * rec_stride] = tb.warp_t_d.y; r[REC_B_TAU_DZ * rec_stride] = tb.warp_t_d.z;
            r[REC_B_OMEGA * rec_stride] = tb.warp_weight;
            r[REC_B_OMEGA_DX * rec_stride] = tb.warp_weight_d.x; r[REC_B_OMEGA_DY * rec_stride] = tb.warp_weight_d.y; r[REC_B_OMEGA_DZ * rec_stride] = tb.warp_weight_d.z;
        }
    }

    // re-projection + splat (python/integrators/reparam.py:100-118; ImageBlock::put)
    float u, v; bool inside; V3 ref;
    sensor_project(sc.S, o + d0, u, v, inside, ref);
    const int b = sc.S.border;
    const int Wb = sc.S.W + 2 * b, Hb = sc.S.H + 2 * b;
    const float pfx = u + ((float)b - 0.5f), pfy = v + ((float)b - 0.5f);
    const float rad = sc.S.filter_radius;
    const int x0 = (int)ceilf(pfx - rad), y0 = (int)ceilf(pfy - rad);
    const int x1 = (int)floorf(pfx + rad), y1 = (int)floorf(pfy + rad);
    for (int by = y0; by <= y1; ++by) {
        if (by < 0 || by >= Hb) continue;
        const float wy = filter_eval(sc.S, (float)by - pfy);
        if (wy == 0.f) continue;
        for (int bx = x0; bx <= x1; ++bx) {
            if (bx < 0 || bx >= Wb) continue;
            const float w = wy * filter_eval(sc.S, (float)bx - pfx);
            if (w == 0.f) continue;
            float* px = film + ((size_t)by * Wb + bx) * 4;
            if (rgb.x != 0.f) SDFRT_ATOMIC_ADD(px + 0, rgb.x * w);
            if (rgb.y != 0.f) SDFRT_ATOMIC_ADD(px + 1, rgb.y * w);
            if (rgb.z != 0.f) SDFRT_ATOMIC_ADD(px + 2, rgb.z * w);
            SDFRT_ATOMIC_ADD(px + 3, w);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// adjoint scatter of one B-spline site: grad[k] += phi_hat * B_k(x) + g_hat . grad B_k(x) over the 64 taps.
// Clamp-to-edge is folded into the weights (cell_setup), so the taps are the unclamped block [b, b+3]^3.
// `add(ix, iy, iz, v)` receives grid coordinates.
// ------------------------------------------------------------------------------------------------
template <class Add>
SDFRT_HD void scatter_site_t(const GridView& G, V3 x, float phi_hat, V3 g_hat, Add add) {
    CellSetup<1> s;
    cell_setup<1>(G, x, s);
    if (G.has_transform)  // g_world = A^T g_local  =>  adjoint: g_hat_local = A g_hat_world
        g_hat = V3{G.A[0] * g_hat.x + G.A[1] * g_hat.y + G.A[2] * g_hat.z, G.A[3] * g_hat.x + G.A[4] * g_hat.y + G.A[5] * g_hat.z,
                   G.A[6] * g_hat.x + G.A[7] * g_hat.y + G.A[8] * g_hat.z};
    const float gx = g_hat.x * (float)G.rx, gy = g_hat.y * (float)G.ry, gz = g_hat.z * (float)G.rz;
    float ex[4];  // phi_hat * wx + gx * dx  (x-direction factor shared by all rows)
#pragma unroll
    for (int i = 0; i < 4; ++i) ex[i] = phi_hat * s.wx[i] + gx * s.dx[i];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float a = s.wy[j] * s.wz[k];                                   // multiplies ex[i]
            const float bq = gy * s.dy[j] * s.wz[k] + gz * s.wy[j] * s.dz[k];    // multiplies wx[i]
#pragma unroll
            for (int i = 0; i < 4; ++i) add(s.bx + i, s.by + j, s.bz + k, a * ex[i] + bq * s.wx[i]);
        }
    }
}
// the 64 tap values of a site in block order t = (k*4 + j)*4 + i, and the block's base texel
template <class Store>
SDFRT_HD void site_tap_values(const GridView& G, V3 x, float phi_hat, V3 g_hat, int& bx, int& by, int& bz, Store store) {
    CellSetup<1> s;
    cell_setup<1>(G, x, s);
    bx = s.bx; by = s.by; bz = s.bz;
    if (G.has_transform)  // g_world = A^T g_local  =>  adjoint: g_hat_local = A g_hat_world
        g_hat = V3{G.A[0] * g_hat.x + G.A[1] * g_hat.y + G.A[2] * g_hat.z, G.A[3] * g_hat.x + G.A[4] * g_hat.y + G.A[5] * g_hat.z,
                   G.A[6] * g_hat.x + G.A[7] * g_hat.y + G.A[8] * g_hat.z};
    const float gx = g_hat.x * (float)G.rx, gy = g_hat.y * (float)G.ry, gz = g_hat.z * (float)G.rz;
    float ex[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) ex[i] = phi_hat * s.wx[i] + gx * s.dx[i];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float a = s.wy[j] * s.wz[k];
            const float bq = gy * s.dy[j] * s.wz[k] + gz * s.wy[j] * s.dz[k];
#pragma unroll
            for (int i = 0; i < 4; ++i) store((k * 4 + j) * 4 + i, a * ex[i] + bq * s.wx[i]);
        }
    }
}
struct GlobalAdd {
    float* grad;
    int rx, ry;
    SDFRT_HD void operator()(int ix, int iy, int iz, float v) const {
        SDFRT_ATOMIC_ADD(grad + ((size_t)iz * ry + iy) * (size_t)rx + ix, v);
    }
};
SDFRT_HD void scatter_site(float* grad, const GridView& G, V3 x, float phi_hat, V3 g_hat) {
    scatter_site_t(G, x, phi_hat, g_hat, GlobalAdd{grad, G.rx, G.ry});
}

// The <= 3 scatter sites of one lane (shadow warp site, surface interaction, primary warp site).  lane_backward only
// collects them; the caller scatters (directly, or warp-aggregated through shared memory in k_backward).
struct Site {
    V3 x;
    float phi_hat;
    V3 g_hat;
};
struct SiteList {
    int n;
    Site s[5];  // (use_mis: BSDF ray's warp site), shadow warp site, (decouple_reparam: detached-ray interaction), surface interaction, primary warp site
};
SDFRT_HD void push_site(SiteList& L, V3 x, float phi_hat, V3 g_hat) {
    L.s[L.n].x = x; L.s[L.n].phi_hat = phi_hat; L.s[L.n].g_hat = g_hat;
    ++L.n;
}

// Queue of scatter sites (device): producers append (position, value adjoint, gradient adjoint), k_scatter_sites drains it with full
// warps through the grouped transposed scatter.  a == nullptr: no queue, sites are scattered where they arise.
struct SiteQueue {
    float4* a;          // (x, phi_hat)
    float4* b;          // (g_hat, unused)
    unsigned* count;    // appended so far (may exceed capacity: the excess was scattered directly)
    unsigned capacity;
};
struct BackwardTargets {
    float* grad_sdf;     // (rz, ry, rx), accumulated (by the caller, from the SiteList)
    float* grad_albedo;  // (rz, ry, rx, 3) or nullptr
    SiteQueue queue;     // sites produced inside the lane code (convolution warp: one per auxiliary hit)
};
// value-only site of the convolution warp's adjoint: into the queue if there is one (opportunistic warp aggregation), else in place
SDFRT_HD void emit_value_site(const BackwardTargets& tg, const GridView& G, V3 x, float phi_hat) {
    if (tg.grad_sdf == nullptr) return;
#if defined(__CUDA_ARCH__)
    if (tg.queue.a != nullptr) {
        const unsigned m = __activemask();
        const int lane_id = threadIdx.x & 31, leader = __ffs(m) - 1;
        unsigned base = 0;
        if (lane_id == leader) base = atomicAdd(tg.queue.count, (unsigned)__popc(m));
        base = __shfl_sync(m, base, leader);
        const unsigned slot = base + (unsigned)__popc(m & ((1u << lane_id) - 1u));
        if (slot < tg.queue.capacity) {
            tg.queue.a[slot] = float4{x.x, x.y, x.z, phi_hat};
            tg.queue.b[slot] = float4{0.f, 0.f, 0.f, 0.f};
            return;
        }
    }
#endif
    scatter_site(tg.grad_sdf, G, x, phi_hat, v3(0.f));
}

// ------------------------------------------------------------------------------------------------
// WarpFieldConvolution (Bangaru et al. 2020; python/warp.py:126-176, python/warp_conv.py) -- the comparison baseline.
// The reference's reparameterisation is a CustomOp whose primal returns (ray.d, 0): the auxiliary rays are only traced when
// the AD graph is traversed, so ALL of it lives in the adjoint sweep.  conv_reparam_adjoint is that `backward`
// (python/warp_conv.py:185-282) for one ray in closed form: pass 1 marches the auxiliary rays and accumulates the normalisation
// Z and its direction derivative dZ, pass 2 pushes  dL/dV_direct = w * grad_V + d_w * grad_div_V_1  through
// V_direct = normalize(detach(p) + (W - detach(W)) - o), W = -g phi / |g|^2 at the auxiliary hit point p: one B-spline site per
// auxiliary hit (value adjoint only) and the adjoint of the ray origin o.  oracle/ref_semantics.md S19.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxAuxRays = 64;

// mitsuba `coordinate_system` (Duff et al. 2017) as mi.Frame3f(n) builds it: both vectors from the closed form
SDFRT_HD void coordinate_system(V3 n, V3& s, V3& t) {
    const float sign = n.z >= 0.f ? 1.f : -1.f;
    const float a = -1.f / (sign + n.z);
    const float b = n.x * n.y * a;
    s = V3{sign * n.x * n.x * a + 1.f, sign * b, -sign * n.x};
    t = V3{b, sign + n.y * n.y * a, -n.y};
}
// One auxiliary direction: warp::square_to_von_mises_fisher (concentric-disk variant) in the frame of d, with the quantities
// the harmonic weight needs, all from m = 1 - cos(theta) = -log(xi) / kappa so that nothing is lost to 1 - (1 - m) in fp32:
// pdf = xi kappa / (2 pi (1 - e^{-2 kappa})),  D = exp(kappa - kappa d.omega) - 1 = 1 / xi - 1.
struct AuxDir {
    V3 omega, tang;   // direction; its component orthogonal to d (omega - d (d . omega))
    float pdf, D;
};
SDFRT_HD AuxDir aux_direction(V3 d, V3 fs, V3 ft, float u0, float u1, float kappa) {
    float px, py;
    square_to_uniform_disk_concentric(u0, u1, px, py);
    const float r2 = px * px + py * py;
    const float sy = fmaxf(1.f - r2, 1e-6f);
    const float e2k = expf(-2.f * kappa);
    const float xi = sy + (1.f - sy) * e2k;
    const float m = -logf(xi) / kappa;
    const float sc = r2 > 0.f ? sqrtf(fmaxf(m * (2.f - m), 0.f) / r2) : 0.f;
    AuxDir a;
    a.tang = (px * sc) * fs + (py * sc) * ft;
    a.omega = a.tang + (1.f - m) * d;
    a.pdf = xi * kappa * (0.5f / kPi) / (1.f - e2k);
    a.D = 1.f / xi - 1.f;
    return a;
}
// harmonic weight w and the factor of its direction derivative (d_w = dwf * tang), both already divided by the pdf (python/warp_conv.py:50-66)
SDFRT_HD void aux_weight(const AuxDir& a, float B, float kappa, float power, float& w, float& dwf) {
    const float den = a.D + B;
    const bool ok = den > 1e-7f;
    const float ip = ok ? powf(den, -power) : 0.f;
    w = ip / a.pdf;
    const float tmp1 = (a.D + 1.f) * (ok ? ip / den : 0.f) * kappa * power;
    dwf = fminf(fabsf(tmp1), 1e10f) / a.pdf;   // tmp1 >= 0
}
SDFRT_HD void rng_skip(Pcg32& rng, int n) {
    for (int i = 0; i < n; ++i) (void)pcg_next_u32(rng);
}
// Adjoint of one reparameterised ray.  g_dir / g_div: adjoints of the returned direction and divergence.  Draws 2 n_aux_rays samples
// from `rng` whatever happens.  Emits one value-only site per auxiliary hit (tg.grad_sdf may be null), accumulates dL/d(sdf.p) and
// returns dL/d(ray origin).
SDFRT_HD V3 conv_reparam_adjoint(const SceneParams& sc, V3 o, V3 d, float maxt, Pcg32& rng, V3 g_dir, float g_div, const BackwardTargets& tg,
                                 V3& grad_p, LaneStats& st) {
    const int N = sc.W.n_aux_rays < kMaxAuxRays ? sc.W.n_aux_rays : kMaxAuxRays;
    const float kappa = sc.W.kappa, power = sc.W.power;
    V3 o_hat = v3(0.f);
    if (!((g_div != 0.f) | (g_dir.x != 0.f) | (g_dir.y != 0.f) | (g_dir.z != 0.f))) {
        rng_skip(rng, 2 * N);
        return o_hat;
    }
    V3 fs, ft;
    coordinate_system(d, fs, ft);
    float its[kMaxAuxRays];
    Pcg32 rng2 = rng;
    float Z = 0.f;
    V3 dZ = v3(0.f);
    for (int i = 0; i < N; ++i) {
        const float u0 = pcg_next_f32(rng), u1 = pcg_next_f32(rng);
        const AuxDir a = aux_direction(d, fs, ft, u0, u1, kappa);
        int ns = 0, nv = 0;
        its[i] = trace_non_diff(sc.G, sc.T, o, a.omega, maxt, true, ns, nv);
        st.n_value += (unsigned)nv;
        float B = 1.f;
        if (isfinite(its[i])) {
            float phi; V3 g; S3 H;
            eval_sdf<1>(sc.G, o + its[i] * a.omega, phi, g, H);
            st.n_grad += 1;
            const float c = dot(g, a.omega);
            B = c * c / dot(g, g);   // boundary test (n . -omega)^2
        }
        float w, dwf;
        aux_weight(a, B, kappa, power, w, dwf);
        Z += w;
        dZ += dwf * a.tang;
    }
    Z = fmaxf(Z, 1e-8f);
    const float iZ = 1.f / Z;
    // direction = normalize(d + V_theta), divergence = (div_V_1 - V_theta . dZ) / Z at V_theta = 0 (python/warp_conv.py:226-236)
    const V3 grad_V = ((g_dir - dot(d, g_dir) * d) - (g_div * iZ) * dZ) * iZ;
    const float grad_div1 = g_div * iZ;
    for (int i = 0; i < N; ++i) {
        const float u0 = pcg_next_f32(rng2), u1 = pcg_next_f32(rng2);
        if (!isfinite(its[i])) continue;   // the background does not move (:47)
        const AuxDir a = aux_direction(d, fs, ft, u0, u1, kappa);
        const V3 p = o + its[i] * a.omega;
        float phi; V3 g; S3 H;
        eval_sdf<1>(sc.G, p, phi, g, H);
        st.n_grad += 1;
        const float gg = dot(g, g), c = dot(g, a.omega);
        float w, dwf;
        aux_weight(a, c * c / gg, kappa, power, w, dwf);
        const V3 V_hat = w * grad_V + (dwf * grad_div1) * a.tang;
        const V3 q_hat = (V_hat - dot(a.omega, V_hat) * a.omega) / its[i];   // through normalize(p - o), |p - o| = its_t
        const float phi_hat = -dot(g, q_hat) / gg;                           // W = -(g / |g|^2) phi(p)
        if (!isfinite(phi_hat) | !isfinite(q_hat.x + q_hat.y + q_hat.z)) continue;
        emit_value_site(tg, sc.G, p, phi_hat);
        st.n_sites += 1;
        grad_p = grad_p - phi_hat * g;
        o_hat = o_hat - q_hat;
    }
    return o_hat;
}


// ------------------------------------------------------------------------------------------------
// backward lane.  film_adj: (Hb, Wb, 4) = (dL/dV_rgb, dL/dWt) per film pixel (see k_film_adjoint).
// Returns this lane's contribution to dL/d(sdf.p) in grad_p.
// ------------------------------------------------------------------------------------------------
// CONV: the warp field is a WarpFieldConvolution (kind WARP_CONV): no tracer records are read, the warp sites are replaced by
// conv_reparam_adjoint, whose auxiliary samples continue the lane's sampler stream after its forward draws in the order the
// reference's AD traversal visits the CustomOps: BSDF-sampled ray (every lane draws), shadow ray (lanes with an emitter sample),
// camera ray (every lane).  oracle/ref_semantics.md S19
template <bool CONV = false>
SDFRT_HD void lane_backward(const SceneParams& sc, const PassParams& ps, unsigned long long lane, const float* film_adj,
                            const float* rec /* already offset to this lane's record */, unsigned long long rec_stride, const BackwardTargets& tg, V3& grad_p,
                            SiteList& sites, LaneStats& st) {
    sites.n = 0;
    const float* r = rec;
    const float its_t = r[REC_ITS_T * rec_stride];
    const bool hit = isfinite(its_t);
    // A sample that misses and has no valid warp site (warp_t = inf: weight below threshold, bbox miss, or a ray the
    // empty-space test skipped) carries no derivative at all: its only differentiable quantities are the warp outputs.
    if (!CONV && !hit && !(reparam_at_depth(sc.W, 0) && isfinite(r[REC_TAU * rec_stride]))) return;
    CameraSample cs = camera_sample(sc, ps, lane);
    const V3 o = cs.o, d0 = cs.d;

    // ---- recompute the forward values that the adjoint needs -----------------------------------
    V3 rgb = v3(0.f);
    V3 p = v3(0.f), n = v3(0.f), g_p = v3(0.f);
    S3 H_p = S3{0, 0, 0, 0, 0, 0};
    float cos_o = 0.f;
    ShadowRay sr; sr.o = v3(0.f); sr.d = v3(0.f); sr.maxt = 0.f;
    V3 rho = v3(0.f);
    V3 drho[3];
    AlbedoTaps taps;
    bool nee = false, bsdf_hit_env = false;
    float shade = 0.f, w_nee = 1.f, w_b = 0.f;
    V3 rgb_b = v3(0.f);
    BsdfRay br; br.active = false; br.o = v3(0.f); br.d = v3(0.f, 0.f, 1.f); br.pdf = 0.f;
    const bool mis = sc.Sh.use_mis && sc.integrator == INTEGRATOR_DIRECT;
    const bool env = sc.E.data != nullptr && sc.integrator == INTEGRATOR_DIRECT;
    V3 K = v3(0.f), Kb = v3(0.f);          // rgb_c = rho_c * (cos_o * K_c + Kb_c)
    V3 dLe_0[3], dLe_s[3], dLe_b[3];       // envmap: d Le_c / d direction for the camera / emitter-sampled / BSDF-sampled ray
    float pdf_e = 0.f, k_env = 0.f;
    if (hit && sc.integrator != INTEGRATOR_SILHOUETTE) {
        p = o + its_t * d0;
        float phi;
        eval_sdf<2>(sc.G, p, phi, g_p, H_p);
        st.n_hess += 1;
        n = normalize(g_p);
    }
    if (sc.integrator == INTEGRATOR_SILHOUETTE) {
        rgb = v3(hit ? 1.f : 0.f);
    } else if (sc.integrator == INTEGRATOR_SIMPLE_SHADING) {
        if (hit) {
            shade = (n.x + n.y + n.z) * 0.57735026918962576f;
            rgb = v3(fmaxf(0.f, shade));
        }
    } else {
        if (!hit) {
            if (!sc.Sh.hide_emitters) {
                rgb = miss_rgb(sc, d0);
                if (env) (void)env_eval(sc.E, d0, dLe_0);
            }
        } else {
            nee = r[REC_S_STATE * rec_stride] != 0.f;
            bsdf_hit_env = mis && r[REC_B_STATE * rec_stride] != 0.f;
            if (nee | bsdf_hit_env | CONV) {   // (CONV: the sampler position after the forward draws matters for every lane)
                const float u0 = pcg_next_f32(cs.rng), u1 = pcg_next_f32(cs.rng);
                sr = sample_emitter_ray(sc, p, n, u0, u1, pdf_e);
                cos_o = dot(n, sr.d);
                rho = eval_albedo(sc.A, p, &taps, drho);
            }
            if (nee) {
                // rgb_c = rho_c * cos_o * K_c;  K = 4 w Le (constant emitter) or w Le(d_s) / (pi pdf) (envmap), w detached
                if (!env) {
                    if (mis) w_nee = mis_weight_nee(cos_o);
                    const float k = cos_o * 4.f * w_nee;
                    rgb = V3{rho.x * k * sc.Sh.Le[0], rho.y * k * sc.Sh.Le[1], rho.z * k * sc.Sh.Le[2]};
                    K = V3{4.f * w_nee * sc.Sh.Le[0], 4.f * w_nee * sc.Sh.Le[1], 4.f * w_nee * sc.Sh.Le[2]};
                } else {
                    if (mis) w_nee = mis_weight(pdf_e, cos_o * (1.f / kPi));
                    k_env = (1.f / kPi) / pdf_e * w_nee;
                    K = k_env * env_eval(sc.E, sr.d, dLe_s);
                    rgb = V3{rho.x * cos_o * K.x, rho.y * cos_o * K.y, rho.z * cos_o * K.z};
                }
            }
            if (CONV && mis && !bsdf_hit_env) rng_skip(cs.rng, 3);
            if (bsdf_hit_env) {
                (void)pcg_next_f32(cs.rng);
                const float b0 = pcg_next_f32(cs.rng), b1 = pcg_next_f32(cs.rng);
                br = make_bsdf_ray(p, n, -dot(n, d0), b0, b1);
                if (!env) {
                    w_b = mis_weight_bsdf(br.pdf);
                    Kb = V3{w_b * sc.Sh.Le[0], w_b * sc.Sh.Le[1], w_b * sc.Sh.Le[2]};
                } else {
                    w_b = mis_weight(br.pdf, env_pdf(sc.E, br.d));
                    Kb = w_b * env_eval(sc.E, br.d, dLe_b);
                }
                rgb_b = V3{rho.x * Kb.x, rho.y * Kb.y, rho.z * Kb.z};
                rgb += rgb_b;
            }
        }
    }

    // ---- WarpFieldConvolution: where each reparameterised ray's auxiliary draws start in the lane's stream ------
    Pcg32 rng_b = cs.rng, rng_s = cs.rng, rng_p = cs.rng;
    if (CONV) {
        const int n2 = 2 * (sc.W.n_aux_rays < kMaxAuxRays ? sc.W.n_aux_rays : kMaxAuxRays);
        const bool direct = sc.integrator == INTEGRATOR_DIRECT;
        if (direct && mis && conv_at_depth(sc.W, 1)) rng_skip(cs.rng, n2);                         // BSDF-sampled ray: active = True (:95)
        rng_s = cs.rng;
        if (direct && hit && pdf_e > 0.f && conv_at_depth(sc.W, 1)) rng_skip(cs.rng, n2);          // shadow ray: active_e (:42,52)
        rng_p = cs.rng;
    }

    // ---- film adjoint: gather over the filter footprint ------------------------------------------
    float u, v; bool inside; V3 ref;
    sensor_project(sc.S, o + d0, u, v, inside, ref);
    const int b = sc.S.border;
    const int Wb = sc.S.W + 2 * b, Hb = sc.S.H + 2 * b;
    const float pfx = u + ((float)b - 0.5f), pfy = v + ((float)b - 0.5f);
    const float rad = sc.S.filter_radius;
    const int x0 = (int)ceilf(pfx - rad), y0 = (int)ceilf(pfy - rad);
    const int x1 = (int)floorf(pfx + rad), y1 = (int)floorf(pfy + rad);
    V3 c_hat = v3(0.f);
    float wch_hat = 0.f, u_hat = 0.f, v_hat = 0.f;
    {
        // separable filter: <= 5 taps per axis (radius <= 2.5), weights and derivatives evaluated once per axis
        float fxs[6], dfxs[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const float t = (float)(x0 + i) - pfx;
            const bool in = (x0 + i <= x1) & (x0 + i >= 0) & (x0 + i < Wb);
            fxs[i] = in ? filter_eval(sc.S, t) : 0.f;
            dfxs[i] = in ? filter_deriv(sc.S, t) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const int by = y0 + j;
            if (by > y1 || by < 0 || by >= Hb) continue;
            const float fy = filter_eval(sc.S, (float)by - pfy);
            const float dfy = filter_deriv(sc.S, (float)by - pfy);
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                if (fxs[i] == 0.f && dfxs[i] == 0.f) continue;
                const float* a = film_adj + ((size_t)by * Wb + (x0 + i)) * 4;
                const float a0 = a[0], a1 = a[1], a2 = a[2], a3 = a[3];
                const float w = fxs[i] * fy;
                c_hat += w * V3{a0, a1, a2};
                wch_hat += w * a3;
                const float sq = a0 * rgb.x + a1 * rgb.y + a2 * rgb.z + a3;  // d L / d f_q  (c = rgb, wch = 1)
                u_hat += -dfxs[i] * fy * sq;  // d f_q / d u = -fx'(bx - pfx) fy
                v_hat += -fxs[i] * dfy * sq;
            }
        }
    }
    // c = rw * rgb, wch = div * rw  (values rw = div = 1)
    const float crgb = dot(c_hat, rgb);
    const float rw_hat = crgb + wch_hat;
    float div_hat = wch_hat + crgb;  // rgb is proportional to the primary det in every integrator

    // ---- re-projection: it.p = o + d  ->  (uv, rw) ------------------------------------------------
    V3 ref_hat = v3(0.f);
    {
        const float iz = 1.f / ref.z, it = 1.f / sc.S.tan_half_fov;
        const float Wf = (float)sc.S.W, Hf = (float)sc.S.H;
        // u = W (0.5 - 0.5 x / (z tan)),  v = H (0.5 - 0.5 aspect y / (z tan))
        ref_hat.x += u_hat * (-0.5f * Wf * iz * it);
        ref_hat.z += u_hat * (0.5f * Wf * ref.x * iz * iz * it);
        ref_hat.y += v_hat * (-0.5f * Hf * sc.S.aspect * iz * it);
        ref_hat.z += v_hat * (0.5f * Hf * sc.S.aspect * ref.y * iz * iz * it);
        if (inside) {  // rw = W_e / detach(W_e), W_e ~ dist / z^3
            const float inv_d2 = 1.f / dot(ref, ref);
            ref_hat += rw_hat * V3{ref.x * inv_d2, ref.y * inv_d2, ref.z * inv_d2 - 3.f * iz};
        }
    }
    V3 d_hat = V3{sc.S.R[0][0] * ref_hat.x + sc.S.R[0][1] * ref_hat.y + sc.S.R[0][2] * ref_hat.z,
                  sc.S.R[1][0] * ref_hat.x + sc.S.R[1][1] * ref_hat.y + sc.S.R[1][2] * ref_hat.z,
                  sc.S.R[2][0] * ref_hat.x + sc.S.R[2][1] * ref_hat.y + sc.S.R[2][2] * ref_hat.z};
    if (env && !hit && !sc.Sh.hide_emitters)   // escaped camera ray: the envmap lookup follows the reparameterised direction
        d_hat += c_hat.x * dLe_0[0] + c_hat.y * dLe_0[1] + c_hat.z * dLe_0[2];

    // ---- shading -> surface interaction -------------------------------------------------------------
    if (hit && sc.integrator != INTEGRATOR_SILHOUETTE) {
        V3 n_hat = v3(0.f), p_hat = v3(0.f);
        if (sc.integrator == INTEGRATOR_SIMPLE_SHADING) {
            if (shade > 0.f) n_hat = v3((c_hat.x + c_hat.y + c_hat.z) * 0.57735026918962576f);
        } else if (nee | bsdf_hit_env) {
            // rgb_c = div * (div_s * rho_c * cos_o * 4 Le_c * w_nee  +  div_b * rho_c * Le_c * w_b); the MIS weights are detached
            const float cos_hat = c_hat.x * rho.x * K.x + c_hat.y * rho.y * K.y + c_hat.z * rho.z * K.z;
            n_hat = cos_hat * sr.d;
            V3 ds_hat = cos_hat * n;
            V3 db_hat = v3(0.f);
            if (env) {   // the envmap is looked up with the attached (reparameterised) directions
                if (nee) ds_hat += (cos_o * k_env) * ((c_hat.x * rho.x) * dLe_s[0] + (c_hat.y * rho.y) * dLe_s[1] + (c_hat.z * rho.z) * dLe_s[2]);
                if (bsdf_hit_env) db_hat = w_b * ((c_hat.x * rho.x) * dLe_b[0] + (c_hat.y * rho.y) * dLe_b[1] + (c_hat.z * rho.z) * dLe_b[2]);
            }
            const float div_s_hat = crgb - dot(c_hat, rgb_b);   // the emitter-sampled part only
            // albedo (both strategies)
            const V3 rho_hat = V3{c_hat.x * (K.x * cos_o + Kb.x), c_hat.y * (K.y * cos_o + Kb.y), c_hat.z * (K.z * cos_o + Kb.z)};
            if (sc.A.data != nullptr) {
                st.n_albedo += 1;
                p_hat += rho_hat.x * drho[0] + rho_hat.y * drho[1] + rho_hat.z * drho[2];
                if (tg.grad_albedo != nullptr) {
                    for (int k = 0; k < 8; ++k) {
                        float* t = tg.grad_albedo + (size_t)taps.idx[k] * 3;
                        SDFRT_ATOMIC_ADD(t + 0, taps.w[k] * rho_hat.x);
                        SDFRT_ATOMIC_ADD(t + 1, taps.w[k] * rho_hat.y);
                        SDFRT_ATOMIC_ADD(t + 2, taps.w[k] * rho_hat.z);
                    }
                }
            }
            // BSDF-sampled ray's warp site: spawned from the attached interaction (si.spawn_ray, :93), direction detached -- only its
            // divergence carries a derivative
            if (CONV && bsdf_hit_env && conv_at_depth(sc.W, 1)) {
                const V3 ob_hat = conv_reparam_adjoint(sc, br.o, br.d, INFINITY, rng_b, db_hat, dot(c_hat, rgb_b), tg, grad_p, st);
                p_hat += ob_hat;
            }
            if (bsdf_hit_env && reparam_at_depth(sc.W, 1)) {
                const V3 tau_d = V3{r[REC_B_TAU_DX * rec_stride], r[REC_B_TAU_DY * rec_stride], r[REC_B_TAU_DZ * rec_stride]};
                const V3 om_d = V3{r[REC_B_OMEGA_DX * rec_stride], r[REC_B_OMEGA_DY * rec_stride], r[REC_B_OMEGA_DZ * rec_stride]};
                const float tau = r[REC_B_TAU * rec_stride];
                const WarpSite ws = warp_site(sc.G, sc.W, sc.T, br.o, br.d, tau, tau_d, r[REC_B_OMEGA * rec_stride], om_d);
                st.n_hess += isfinite(tau);
                if (ws.active) {
                    const float div_b_hat = dot(c_hat, rgb_b);
                    const float phi_hat = div_b_hat * ws.cphi_div + dot(db_hat, ws.cphi_dir);
                    const V3 g_hat = div_b_hat * ws.cg_div;
                    if (isfinite(phi_hat) & isfinite(g_hat.x + g_hat.y + g_hat.z)) {
                        push_site(sites, br.o + tau * br.d, phi_hat, g_hat);
                        st.n_sites += 1;
                        const V3 ob_hat = phi_hat * ws.g + mul(ws.H, g_hat);
                        grad_p = grad_p - ob_hat;
                        p_hat += ob_hat;
                    }
                }
            }
            if (CONV && nee && conv_at_depth(sc.W, 1)) {
                const V3 os_hat = conv_reparam_adjoint(sc, sr.o, sr.d, sr.maxt, rng_s, ds_hat, div_s_hat, tg, grad_p, st);
                if (sc.Sh.shadow_origin == 0) {
                    p_hat += os_hat;
                } else if (sc.Sh.shadow_origin == 2) {
                    const float phi0_hat = dot(os_hat, d0) / (-dot(g_p, d0));
                    if (isfinite(phi0_hat)) {
                        push_site(sites, p, phi0_hat, v3(0.f));
                        st.n_sites += 1;
                        grad_p = grad_p - phi0_hat * g_p;
                    }
                }
            }
            // shadow-ray warp site (attached through its origin o_s = p + const)
            if (nee && reparam_at_depth(sc.W, 1)) {
                const V3 tau_d = V3{r[REC_S_TAU_DX * rec_stride], r[REC_S_TAU_DY * rec_stride], r[REC_S_TAU_DZ * rec_stride]};
                const V3 om_d = V3{r[REC_S_OMEGA_DX * rec_stride], r[REC_S_OMEGA_DY * rec_stride], r[REC_S_OMEGA_DZ * rec_stride]};
                const float tau = r[REC_S_TAU * rec_stride];
                const WarpSite ws = warp_site(sc.G, sc.W, sc.T, sr.o, sr.d, tau, tau_d, r[REC_S_OMEGA * rec_stride], om_d);
                st.n_hess += isfinite(tau);
                if (ws.active) {
                    const float phi_hat = div_s_hat * ws.cphi_div + dot(ds_hat, ws.cphi_dir);
                    const V3 g_hat = div_s_hat * ws.cg_div;
                    push_site(sites, sr.o + tau * sr.d, phi_hat, g_hat);
                    st.n_sites += 1;
                    const V3 os_hat = phi_hat * ws.g + mul(ws.H, g_hat);
                    grad_p = grad_p - os_hat;
                    if (sc.Sh.shadow_origin == 0) {
                        p_hat += os_hat;                       // spawned from the attached si
                    } else if (sc.Sh.shadow_origin == 2) {     // si_d0: p0 = ray0(t~), t~ = t <- phi(p) / detach(g . -d0); no path to d
                        const float phi0_hat = dot(os_hat, d0) / (-dot(g_p, d0));
                        if (isfinite(phi0_hat)) {
                            push_site(sites, p, phi0_hat, v3(0.f));
                            st.n_sites += 1;
                            grad_p = grad_p - phi0_hat * g_p;
                        }
                    }                                          // 1: detach(si) -- the origin carries no derivative
                }
            }
        }
        // compute_surface_interaction (python/shapes.py:347-366): n = normalize(g(p)), p = o + t d,
        // t = t* <- phi(p0) / detach(g . -d)
        const float inv_gn = 1.f / sqrtf(dot(g_p, g_p));
        const V3 gp_hat = (n_hat - dot(n, n_hat) * n) * inv_gn;
        p_hat += mul(H_p, gp_hat);
        const float t_hat = dot(p_hat, d0);
        const float phi2_hat = t_hat / (-dot(g_p, d0));
        const V3 dd = its_t * p_hat + (its_t * phi2_hat) * g_p;
        if (isfinite(phi2_hat) & isfinite(gp_hat.x + gp_hat.y + gp_hat.z) & isfinite(dd.x + dd.y + dd.z)) {
            d_hat += dd;
            push_site(sites, p, phi2_hat, gp_hat);
            st.n_sites += 1;
            grad_p = grad_p - (phi2_hat * g_p + mul(H_p, gp_hat));
        }
    }

    // ---- primary warp site ---------------------------------------------------------------------------
    if (CONV && conv_at_depth(sc.W, 0)) (void)conv_reparam_adjoint(sc, o, d0, cs.maxt, rng_p, d_hat, div_hat, tg, grad_p, st);
    if (reparam_at_depth(sc.W, 0)) {
        const V3 tau_d = V3{r[REC_TAU_DX * rec_stride], r[REC_TAU_DY * rec_stride], r[REC_TAU_DZ * rec_stride]};
        const V3 om_d = V3{r[REC_OMEGA_DX * rec_stride], r[REC_OMEGA_DY * rec_stride], r[REC_OMEGA_DZ * rec_stride]};
        const float tau = r[REC_TAU * rec_stride];
        const WarpSite ws = warp_site(sc.G, sc.W, sc.T, o, d0, tau, tau_d, r[REC_OMEGA * rec_stride], om_d);
        st.n_hess += isfinite(tau);
        if (ws.active) {
            const float phi_hat = div_hat * ws.cphi_div + dot(d_hat, ws.cphi_dir);
            const V3 g_hat = div_hat * ws.cg_div;
            if (isfinite(phi_hat) & isfinite(g_hat.x + g_hat.y + g_hat.z)) {
                push_site(sites, o + tau * d0, phi_hat, g_hat);
                st.n_sites += 1;
                grad_p = grad_p - (phi_hat * ws.g + mul(ws.H, g_hat));
            }
        }
    }
}

}  // namespace sdfrt
