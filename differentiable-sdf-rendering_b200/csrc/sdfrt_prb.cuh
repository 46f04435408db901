// sdfrt_prb.cuh -- one wavefront lane of the path-replay integrator `sdf_prb_reparam`
// (python/integrators/sdf_prb_reparam.py), diffuse BSDF + constant environment emitter.
//
//   lane_prb<false>   SdfPrbReparamIntegrator.sample in ADMode.Primal (:242-439): emitter sampling with MIS,
//                     cosine-hemisphere BSDF sampling, Russian roulette from rr_depth on.  Returns L.
//   lane_prb<true>    the same path replayed in ADMode.Backward with delta_L = dL/dL of the film (:222-234): at every
//                     vertex the closed-form adjoint of  Lo = (Le + Lr_dir + Lr_ind) * det + extra  (:414-428),
//                     emitted as B-spline scatter sites.  Starts from L = state_in and subtracts, as path replay does.
//
// What is attached at a vertex (everything else is detached by the reference's suspend_grad / detach calls):
//   segment warp      ray_reparam.d = d <- warp dir,  det = 1 <- div  (WarpField2D.reparam on si_prev.spawn_ray(d), :294);
//                     the warp site x = o_s + warp_t d is attached through its origin o_s = si_prev.p + const, and
//                     si_prev.p = ray_prev(t~), t~ = t <- phi(p) / detach(g . -d)  (python/shapes.py:354-356)
//   si_cur            p = o + t~ d~,  n = normalize(g(p))                       (:300, python/shapes.py:347-366)
//   Lr_dir            beta * mis * (rho(p)/pi) * (n . d_s) * 4 pi Le * det_s, shadow direction and det_s from the shadow
//                     ray's warp site (attached through its origin p + const)   (:313-333)
//   Lr_ind            L * (bsdf_val / detach(bsdf_val)),  bsdf_val = rho(p)/pi * (n . wo)               (:407-412)
//   extra             L_prev * cos(n_prev, normalize(p - p_prev)) / detach(.)   (:389-406; the `next` term of :399,406
//                     carries no derivative for a diffuse BSDF, and Le_next none for a constant emitter)
// The walker marches inline (per-lane form, host harness) or reads the marches from per-vertex records (wavefront form: PrbRecords,
// PrbStop below).  Derivation and parity: DESIGN.md section 8, tests/test_prb.py (against the autograd oracle), tests/test_gpu_prb.py.
#pragma once
#include "sdfrt_lane.cuh"

namespace sdfrt {


// The path walker marches up to three kinds of segment per vertex and evaluates up to three warp sites.  Inlining every
// call site made k_prb_backward instruction-fetch bound (ncu: `no_instruction` 5.6 stall cycles per issued instruction, 255
// registers; profiles/r1_prb_per_lane.md), so the three heavy routines are real calls here, one copy each.
#if defined(__CUDACC__)
#define SDFRT_HD_NOINLINE __host__ __device__ __noinline__
#else
#define SDFRT_HD_NOINLINE __attribute__((noinline))
#endif
struct ValueTrace {
    float its_t;
    int n_steps, n_evals;
};
SDFRT_HD_NOINLINE ValueTrace prb_trace_value(const GridView& G, const TracerParams& T, V3 o, V3 d, float ray_maxt, bool do_refine) {
    ValueTrace r;
    r.n_steps = 0; r.n_evals = 0;
    r.its_t = trace_non_diff(G, T, o, d, ray_maxt, do_refine, r.n_steps, r.n_evals);
    return r;
}
// primal walker: inlined (ncu: the call costs it 18 % -- long_scoreboard 1.9 -> 4.7 stall cycles per issue, fewer loads in
// flight across the call boundary); replay walker: the real call (its problem was instruction fetch)
template <bool CALL>
SDFRT_HD ValueTrace trace_value_sel(const GridView& G, const TracerParams& T, V3 o, V3 d, float ray_maxt, bool do_refine) {
    if (CALL) return prb_trace_value(G, T, o, d, ray_maxt, do_refine);
    ValueTrace r;
    r.n_steps = 0; r.n_evals = 0;
    r.its_t = trace_non_diff(G, T, o, d, ray_maxt, do_refine, r.n_steps, r.n_evals);
    return r;
}
struct DiffTrace {
    TraceOut out;
    int n_hess, n_value;
};
SDFRT_HD_NOINLINE DiffTrace prb_trace_diff(const GridView& G, const TracerParams& T, V3 o, V3 d, float ray_maxt) {
    DiffTrace r;
    r.n_hess = 0; r.n_value = 0;
    r.out.warp_t = INFINITY; r.out.warp_t_d = v3(0.f); r.out.warp_weight = 0.f; r.out.warp_weight_d = v3(0.f);
    trace_diff(G, T, o, d, ray_maxt, false, r.out, r.n_hess, r.n_value);
    return r;
}
struct Lookup2 {
    float phi;
    V3 g;
    S3 H;
};
SDFRT_HD_NOINLINE Lookup2 prb_eval2(const GridView& G, V3 x) {
    Lookup2 r;
    eval_sdf<2>(G, x, r.phi, r.g, r.H);
    return r;
}
SDFRT_HD_NOINLINE void prb_scatter_site(float* grad, const GridView& G, V3 x, float phi_hat, V3 g_hat) {
    scatter_site(grad, G, x, phi_hat, g_hat);
}
SDFRT_HD_NOINLINE WarpSite prb_warp_site(const GridView& G, const WarpParams& W, const TracerParams& T, V3 o, V3 d0, const TraceOut& tr) {
    return warp_site(G, W, T, o, d0, tr.warp_t, tr.warp_t_d, tr.warp_weight, tr.warp_weight_d);
}

// adjoint of the film splat of one sample whose splatted value is `rgb` with weight channel 1:
// c_hat = dL/d rgb, wch_hat = dL/d weight, (u_hat, v_hat) = dL/d uv
SDFRT_HD void film_gather(const SensorParams& S, const float* film_adj, float u, float v, V3 rgb, V3& c_hat, float& wch_hat,
                          float& u_hat, float& v_hat) {
    const int b = S.border;
    const int Wb = S.W + 2 * b, Hb = S.H + 2 * b;
    const float pfx = u + ((float)b - 0.5f), pfy = v + ((float)b - 0.5f);
    const float rad = S.filter_radius;
    const int x0 = (int)ceilf(pfx - rad), y0 = (int)ceilf(pfy - rad);
    const int x1 = (int)floorf(pfx + rad), y1 = (int)floorf(pfy + rad);
    c_hat = v3(0.f);
    wch_hat = 0.f; u_hat = 0.f; v_hat = 0.f;
    for (int by = y0; by <= y1; ++by) {
        if (by < 0 || by >= Hb) continue;
        const float fy = filter_eval(S, (float)by - pfy), dfy = filter_deriv(S, (float)by - pfy);
        for (int bx = x0; bx <= x1; ++bx) {
            if (bx < 0 || bx >= Wb) continue;
            const float fx = filter_eval(S, (float)bx - pfx), dfx = filter_deriv(S, (float)bx - pfx);
            if (fx == 0.f && dfx == 0.f) continue;
            const float* a = film_adj + ((size_t)by * Wb + bx) * 4;
            const float a0 = a[0], a1 = a[1], a2 = a[2], a3 = a[3];
            const float w = fx * fy;
            c_hat += w * V3{a0, a1, a2};
            wch_hat += w * a3;
            const float sq = a0 * rgb.x + a1 * rgb.y + a2 * rgb.z + a3;
            u_hat += -dfx * fy * sq;
            v_hat += -fx * dfy * sq;
        }
    }
}

// Result of the camera ray's march when it was done by the wavefront marcher (k_march over the primary segments, with the
// empty-space test): hit distance and, for the replay, the differentiable tracer's record.  nullptr = march here.
struct PrbPrimary {
    float its_t;
    TraceOut tr;   // warp_t = inf if there is none
};

// Wavefront form: every segment of the path is marched by k_march, bounce by bounce, and its result is
// kept as a 9-float record (its_t + the differentiable tracer's outputs; a value march fills row 0 only).  Record (v, s) belongs to
// vertex v: s = 0 the segment that ENDS at the vertex (camera ray for v = 0), s = 1 its emitter-sample shadow ray.
// Row r of record (v, s) of sample i is base[((v * 2 + s) * 9 + r) * stride + i]; `base` is already offset to the sample.
struct PrbRecords {
    const float* base;  // nullptr: march inline (per-lane form, host harness)
    unsigned long long stride;
    int rows;           // rows per record: 9, or 1 in value-only sweeps (its_t alone)
    SDFRT_HD float its(int v, int s) const { return base[(size_t)((v * 2 + s) * rows) * stride]; }
    SDFRT_HD TraceOut rec(int v, int s) const {  // rows == 9 only
        const float* r = base + (size_t)((v * 2 + s) * 9) * stride;
        TraceOut t;
        t.its_t = r[0];
        t.warp_t = r[1 * stride];
        t.warp_t_d = V3{r[2 * stride], r[3 * stride], r[4 * stride]};
        t.warp_weight = r[5 * stride];
        t.warp_weight_d = V3{r[6 * stride], r[7 * stride], r[8 * stride]};
        t.n_steps = 0;
        return t;
    }
};
constexpr int kPrbRecFloats = 18;  // per vertex: segment record + shadow record
// The walker as segment generator: walk (from records) up to vertex `stop_at`, report the two rays that leave it, return.
struct PrbStop {
    int stop_at;
    bool shadow_active, next_active;
    V3 shadow_o, shadow_d, next_o, next_d;
    float shadow_maxt;
};

template <bool BWD, class Emit>
SDFRT_HD V3 lane_prb(const SceneParams& sc, const PassParams& ps, unsigned long long lane, const float* film_adj, V3 L_total, Emit& emit,
                     V3& grad_p, LaneStats& st, const PrbPrimary* pre = nullptr, PrbRecords recs = PrbRecords{nullptr, 0, 9},
                     PrbStop* stop = nullptr) {
    const GridView& G = sc.G;
    const TracerParams& T = sc.T;
    const WarpParams& W = sc.W;
    const int max_depth = sc.prb.max_depth, rr_depth = sc.prb.rr_depth;
    const bool warp2d = W.kind == WARP_2D;
    // :267-272: reparam_max_depth (and with it shade_gradient_max_depth) follows the warp field
    const int rmd = warp2d ? (W.max_reparam_depth >= 0 ? W.max_reparam_depth : 100000) : max_depth;
    const V3 Le_env = V3{sc.Sh.Le[0], sc.Sh.Le[1], sc.Sh.Le[2]};
    CameraSample cs = camera_sample(sc, ps, lane);  // sample_rays (:441-497): jitter (2), sample1 (1)
    int nv = 0, nh = 0, nsteps = 0;

    V3 o_cur = cs.o, d_cur = cs.d;
    float t_cur;
    if (recs.base != nullptr) {
        t_cur = recs.its(0, 0);
    } else if (pre != nullptr) {
        t_cur = pre->its_t;
    } else {  // ray_intersect_preliminary (:264)
        const ValueTrace vt = trace_value_sel<BWD>(G, T, o_cur, d_cur, cs.maxt, true);
        t_cur = vt.its_t; nsteps = vt.n_steps; nv += vt.n_evals;
    }
    st.n_hit += isfinite(t_cur);
    V3 L = BWD ? L_total : v3(0.f);
    V3 beta = v3(1.f);
    float mis_em = 1.f;
    int depth = 0;
    V3 o_prev = v3(0.f), d_prev = v3(0.f), g_prev = v3(0.f);
    float t_prev = INFINITY, cos_i_prev = 0.f;

    // ---- differential phase: film adjoint of this sample and the primary ray's warp record ----------------
    V3 dL = v3(0.f), d_hat0 = v3(0.f);
    float det_hat0 = 0.f;
    TraceOut tr0;
    tr0.warp_t = INFINITY; tr0.warp_t_d = v3(0.f); tr0.warp_weight = 0.f; tr0.warp_weight_d = v3(0.f);
    if (BWD) {
        if (warp2d && recs.base != nullptr) {
            tr0 = recs.rec(0, 0);
        } else if (warp2d && pre != nullptr) {
            tr0 = pre->tr;
        } else if (warp2d) {  // sample_rays with reparam (:499-511)
            const DiffTrace dt = prb_trace_diff(G, T, o_cur, d_cur, cs.maxt);
            tr0 = dt.out; nh += dt.n_hess; nv += dt.n_value;
        }
        // a sample that escapes and has no valid primary warp carries no derivative at all
        if (!isfinite(t_cur) && !isfinite(tr0.warp_t)) {
            st.n_value += nv; st.n_hess += nh;
            return L;
        }
        float u, v; bool inside; V3 ref;
        sensor_project(sc.S, cs.o + cs.d, u, v, inside, ref);
        float wch_hat, u_hat, v_hat;
        film_gather(sc.S, film_adj, u, v, L_total, dL, wch_hat, u_hat, v_hat);  // block.put(pos, L * weight * det, weight = det) (:211-213)
        det_hat0 = dot(dL, L_total) + wch_hat;
        const float iz = 1.f / ref.z, itn = 1.f / sc.S.tan_half_fov;
        const float Wf = (float)sc.S.W, Hf = (float)sc.S.H;
        V3 ref_hat;
        ref_hat.x = u_hat * (-0.5f * Wf * iz * itn);
        ref_hat.y = v_hat * (-0.5f * Hf * sc.S.aspect * iz * itn);
        ref_hat.z = u_hat * (0.5f * Wf * ref.x * iz * iz * itn) + v_hat * (0.5f * Hf * sc.S.aspect * ref.y * iz * iz * itn);
        d_hat0 = V3{sc.S.R[0][0] * ref_hat.x + sc.S.R[0][1] * ref_hat.y + sc.S.R[0][2] * ref_hat.z,
                    sc.S.R[1][0] * ref_hat.x + sc.S.R[1][1] * ref_hat.y + sc.S.R[1][2] * ref_hat.z,
                    sc.S.R[2][0] * ref_hat.x + sc.S.R[2][1] * ref_hat.y + sc.S.R[2][2] * ref_hat.z};
    }

    for (int it = 0; it < max_depth; ++it) {  // loop.set_max_iterations(self.max_depth) (:283)
        const float u_em0 = pcg_next_f32(cs.rng), u_em1 = pcg_next_f32(cs.rng);   // emitter sample (:310)
        (void)pcg_next_f32(cs.rng);                                               // BSDF component sample (:336), unused by `diffuse`
        const float u_b0 = pcg_next_f32(cs.rng), u_b1 = pcg_next_f32(cs.rng);     // BSDF direction sample
        const float u_rr = pcg_next_f32(cs.rng);                                  // Russian roulette (:348)
        const bool first = depth == 0;
        const bool hit = isfinite(t_cur);
        // Lo[depth > shade_gradient_max_depth] = 0 (:421) and the reparam gate depth <= reparam_max_depth share one threshold
        const bool grad_here = BWD && depth <= rmd;

        // ---- reparameterise the segment that ends at this vertex (:290-296) ----------------------------------
        TraceOut trs;
        trs.warp_t = INFINITY; trs.warp_t_d = v3(0.f); trs.warp_weight = 0.f; trs.warp_weight_d = v3(0.f);
        WarpSite wseg;
        wseg.active = false;
        if (grad_here && !first && warp2d) {
            if (recs.base != nullptr) {
                trs = recs.rec(it, 0);
            } else {
                const DiffTrace dt = prb_trace_diff(G, T, o_cur, d_cur, INFINITY);  // o_cur == si_prev.spawn_ray(ray_cur.d).o
                trs = dt.out; nh += dt.n_hess; nv += dt.n_value;
            }
            wseg = prb_warp_site(G, W, T, o_cur, d_cur, trs);
            nh += isfinite(trs.warp_t);
        }

        // ---- current interaction, emission, next event estimation (:298-333) ---------------------------------
        V3 p = v3(0.f), n = v3(0.f), g_p = v3(0.f);
        S3 H_p = S3{0, 0, 0, 0, 0, 0};
        float cos_i = 0.f, cos_o = 0.f, mis_direct = 0.f;
        V3 rho = v3(0.f);
        V3 drho[3];
        AlbedoTaps taps;
        V3 Le = v3(0.f), Lr_dir = v3(0.f);
        bool active_next = (depth + 1 < max_depth) && hit;
        bool nee = false;
        ShadowRay sr; sr.o = v3(0.f); sr.d = v3(0.f); sr.maxt = 0.f;
        TraceOut ts;
        ts.warp_t = INFINITY; ts.warp_t_d = v3(0.f); ts.warp_weight = 0.f; ts.warp_weight_d = v3(0.f);
        V3 bsdf_weight = v3(0.f), wo_w = v3(0.f), wo_l = v3(0.f), o_next = v3(0.f);
        float bs_pdf = 0.f;
        if (!hit) {
            Le = (mis_em * beta) * Le_env;  // emitter.eval(si_cur): the constant environment for an escaped ray
        } else {
            p = o_cur + t_cur * d_cur;
            float phi;
            eval_sdf<1>(G, p, phi, g_p, H_p);  // the replay takes the normal from the same lookup as the primal pass
            st.n_grad += 1;
            if (grad_here) { H_p = prb_eval2(G, p).H; nh += 1; }
            n = normalize(g_p);
            cos_i = -dot(n, d_cur);
            if (active_next) {
                rho = eval_albedo(sc.A, p, BWD ? &taps : nullptr, BWD ? drho : nullptr);
                st.n_albedo += (sc.A.data != nullptr);
                sr = make_shadow_ray(sc.Sh, p, n, u_em0, u_em1);  // si_cur.spawn_ray_to(ds.p), d detached (:316-317)
                cos_o = dot(n, sr.d);
                if (stop != nullptr && it == stop->stop_at) {
                    if ((cos_i > 0.f) & (cos_o > 0.f)) {
                        stop->shadow_active = true;
                        stop->shadow_o = sr.o; stop->shadow_d = sr.d; stop->shadow_maxt = sr.maxt;
                    }
                } else if ((cos_i > 0.f) & (cos_o > 0.f)) {  // a closed BSDF gate contributes exactly zero: not marched
                    float its_s;
                    if (recs.base != nullptr) {
                        its_s = recs.its(it, 1);
                    } else {
                        const ValueTrace vs = trace_value_sel<BWD>(G, T, sr.o, sr.d, sr.maxt, false);
                        nv += vs.n_evals;
                        its_s = vs.its_t;
                    }
                    if (!isfinite(its_s)) {
                        nee = true;
                        mis_direct = mis_weight(kInv4Pi, cos_o * (1.f / kPi));
                        const float k = 4.f * cos_o * mis_direct;  // (rho / pi) cos_o * Le / (1 / 4 pi) * mis
                        Lr_dir = k * (beta * rho * Le_env);
                        // the shadow ray's warp record is only needed where it is reparameterised (static depth 1, :321)
                        if (grad_here && reparam_at_depth(W, 1)) {
                            if (recs.base != nullptr) {
                                ts = recs.rec(it, 1);
                            } else {
                                const DiffTrace dt = prb_trace_diff(G, T, sr.o, sr.d, sr.maxt);
                                ts = dt.out; nh += dt.n_hess; nv += dt.n_value;
                            }
                        }
                    }
                }
                // BSDF sampling (:335-338): cosine hemisphere
                wo_l = square_to_cosine_hemisphere(u_b0, u_b1);
                const float pdf = wo_l.z * (1.f / kPi);
                if ((cos_i > 0.f) & (pdf > 0.f)) { bsdf_weight = rho; bs_pdf = pdf; }
                V3 fs, ft;
                sh_frame(n, fs, ft);
                wo_w = wo_l.x * fs + wo_l.y * ft + wo_l.z * n;
                o_next = spawn_offset(p, n, wo_w);  // si_cur.spawn_ray (:352)
            }
        }
        V3 beta_new = beta * bsdf_weight;
        const V3 L_prev = L;
        L = BWD ? (L - Le - Lr_dir) : (L + Le + Lr_dir);  // PRB update rule (:340)

        // ---- stopping criterion (:342-349) -----------------------------------------------------------------------
        const float beta_max = fmaxf(beta_new.x, fmaxf(beta_new.y, beta_new.z));
        active_next = active_next && (beta_max != 0.f);
        const float rr_prob = fminf(beta_max, 0.95f);  // eta == 1
        if (depth >= rr_depth) {
            if (rr_prob > 0.f) beta_new = beta_new * (1.f / rr_prob);
            active_next = active_next && (u_rr < rr_prob);
        }

        // ---- next surface and the MIS weight of hitting the emitter by BSDF sampling (:351-363) ---------------
        if (stop != nullptr && it == stop->stop_at) {  // wavefront: the continuation ray of this vertex goes to the marcher
            stop->next_active = active_next;
            stop->next_o = o_next; stop->next_d = wo_w;
            return L;
        }
        float t_next = INFINITY;
        if (active_next && recs.base != nullptr) {
            t_next = recs.its(it + 1, 0);
        } else if (active_next) {
            const ValueTrace vn = trace_value_sel<BWD>(G, T, o_next, wo_w, INFINITY, true);
            t_next = vn.its_t; nv += vn.n_evals;
            nsteps = nsteps > vn.n_steps ? nsteps : vn.n_steps;
        }
        const float mis_em_new = mis_weight(bs_pdf, isfinite(t_next) ? 0.f : kInv4Pi);

        // ---- adjoint of Lo at this vertex (:414-428) --------------------------------------------------------------
        if (grad_here) {
            const float det_hat = first ? 0.f : dot(dL, Le + Lr_dir + L);  // ray_reparam_det[first_vertex] = 1 (:295)
            V3 d_hat = v3(0.f);
            if (hit) {
                V3 n_hat = v3(0.f), p_hat = v3(0.f), rho_hat = v3(0.f);
                if (nee) {
                    const V3 K = (4.f * mis_direct) * (beta * Le_env);  // Lr_dir_c = K_c rho_c cos_o det_s
                    const V3 a = dL * K;
                    const float cos_hat = dot(a, rho);
                    n_hat += cos_hat * sr.d;
                    const V3 ds_hat = cos_hat * n;
                    const float det_s_hat = dot(dL, Lr_dir);
                    rho_hat += cos_o * a;
                    if (reparam_at_depth(W, 1)) {
                        const WarpSite ws = prb_warp_site(G, W, T, sr.o, sr.d, ts);
                        nh += isfinite(ts.warp_t);
                        if (ws.active) {
                            const float phi_hat = det_s_hat * ws.cphi_div + dot(ds_hat, ws.cphi_dir);
                            const V3 g_hat = det_s_hat * ws.cg_div;
                            emit.site(sr.o + ts.warp_t * sr.d, phi_hat, g_hat);
                            st.n_sites += 1;
                            const V3 os_hat = phi_hat * ws.g + mul(ws.H, g_hat);  // attached through its origin p + const
                            p_hat += os_hat;
                            grad_p = grad_p - os_hat;
                        }
                    }
                }
                if (active_next && (cos_i > 0.f) && (bs_pdf > 0.f)) {  // Lr_ind = L * replace_grad(1, bsdf_val / detach) (:407-412)
                    const float cos_n = dot(n, wo_w);
                    if (cos_n > 0.f) {
                        const V3 aL = dL * L;
                        const float ratio = cos_n / wo_l.z;
                        float cn_hat = 0.f;
                        if (rho.x != 0.f) { rho_hat.x += aL.x * ratio / rho.x; cn_hat += aL.x; }
                        if (rho.y != 0.f) { rho_hat.y += aL.y * ratio / rho.y; cn_hat += aL.y; }
                        if (rho.z != 0.f) { rho_hat.z += aL.z * ratio / rho.z; cn_hat += aL.z; }
                        n_hat += (cn_hat / wo_l.z) * wo_w;
                    }
                }
                if (!first && cos_i_prev > 0.f) {  // extra: previous vertex' BSDF seen from the moving si_cur.p (:389-406)
                    const V3 p_prev = o_prev + t_prev * d_prev;
                    const V3 n_prev = normalize(g_prev);
                    const V3 dv = p - p_prev;
                    const float len = sqrtf(dot(dv, dv));
                    const V3 wo_prev = dv / len;
                    const float cos_p = dot(n_prev, wo_prev);
                    if (cos_p > 0.f) {
                        const V3 wo_hat = (dot(dL, L_prev) / cos_p) * n_prev;
                        p_hat += (wo_hat - dot(wo_prev, wo_hat) * wo_prev) / len;
                    }
                }
                if (sc.A.data != nullptr && (rho_hat.x != 0.f || rho_hat.y != 0.f || rho_hat.z != 0.f)) {
                    p_hat += rho_hat.x * drho[0] + rho_hat.y * drho[1] + rho_hat.z * drho[2];
                    emit.albedo(taps, rho_hat);
                }
                // compute_surface_interaction (python/shapes.py:347-366)
                const float inv_gn = 1.f / sqrtf(dot(g_p, g_p));
                const V3 gp_hat = (n_hat - dot(n, n_hat) * n) * inv_gn;
                p_hat += mul(H_p, gp_hat);
                const float t_hat = dot(p_hat, d_cur);
                const float phi2_hat = t_hat / (-dot(g_p, d_cur));
                const V3 dd = t_cur * p_hat + (t_cur * phi2_hat) * g_p;
                if (isfinite(phi2_hat) & isfinite(gp_hat.x + gp_hat.y + gp_hat.z) & isfinite(dd.x + dd.y + dd.z)) {
                    d_hat = dd;
                    emit.site(p, phi2_hat, gp_hat);
                    st.n_sites += 1;
                    grad_p = grad_p - (phi2_hat * g_p + mul(H_p, gp_hat));
                }
            }
            if (first) {
                if (warp2d) {  // the primary warp site collects the film's and the first vertex' direction adjoints
                    const WarpSite ws = prb_warp_site(G, W, T, cs.o, cs.d, tr0);
                    nh += isfinite(tr0.warp_t);
                    if (ws.active) {
                        const float phi_hat = det_hat0 * ws.cphi_div + dot(d_hat0 + d_hat, ws.cphi_dir);
                        const V3 g_hat = det_hat0 * ws.cg_div;
                        if (isfinite(phi_hat) & isfinite(g_hat.x + g_hat.y + g_hat.z)) {
                            emit.site(cs.o + tr0.warp_t * cs.d, phi_hat, g_hat);
                            st.n_sites += 1;
                            grad_p = grad_p - (phi_hat * ws.g + mul(ws.H, g_hat));
                        }
                    }
                }
            } else if (wseg.active) {
                const float phi_hat = det_hat * wseg.cphi_div + dot(d_hat, wseg.cphi_dir);
                const V3 g_hat = det_hat * wseg.cg_div;
                if (isfinite(phi_hat) & isfinite(g_hat.x + g_hat.y + g_hat.z)) {
                    emit.site(o_cur + trs.warp_t * d_cur, phi_hat, g_hat);
                    st.n_sites += 1;
                    const V3 os_hat = phi_hat * wseg.g + mul(wseg.H, g_hat);
                    grad_p = grad_p - os_hat;
                    // ... and moves with si_prev.p = ray_prev(t~), t~ = t <- phi(p_prev) / detach(g_prev . -d_prev)
                    const float phi_prev_hat = dot(os_hat, d_prev) / (-dot(g_prev, d_prev));
                    if (isfinite(phi_prev_hat)) {
                        emit.site(o_prev + t_prev * d_prev, phi_prev_hat, v3(0.f));
                        st.n_sites += 1;
                        grad_p = grad_p - phi_prev_hat * g_prev;
                    }
                }
            }
        }

        // ---- advance (:430-438) --------------------------------------------------------------------------------------
        o_prev = o_cur; d_prev = d_cur; t_prev = t_cur; g_prev = g_p; cos_i_prev = cos_i;
        o_cur = o_next; d_cur = wo_w; t_cur = t_next;
        beta = beta_new;
        mis_em = mis_em_new;
        depth += hit ? 1 : 0;
        if (!active_next) break;
    }
    st.n_value += nv;
    st.n_hess += nh;
    st.n_steps_max = st.n_steps_max > (unsigned)nsteps ? st.n_steps_max : (unsigned)nsteps;
    return L;
}

struct NoEmit {
    SDFRT_HD void site(V3, float, V3) {}
    SDFRT_HD void albedo(const AlbedoTaps&, V3) {}
};
// direct accumulation (device: red.global.add; host harness: plain adds)
struct DirectEmit {
    float* grad_sdf;
    float* grad_albedo;
    GridView G;
    SDFRT_HD void site(V3 x, float phi_hat, V3 g_hat) { if (grad_sdf != nullptr) prb_scatter_site(grad_sdf, G, x, phi_hat, g_hat); }
    SDFRT_HD void albedo(const AlbedoTaps& taps, V3 rho_hat) {
        if (grad_albedo == nullptr) return;
        for (int k = 0; k < 8; ++k) {
            float* t = grad_albedo + (size_t)taps.idx[k] * 3;
            SDFRT_ATOMIC_ADD(t + 0, taps.w[k] * rho_hat.x);
            SDFRT_ATOMIC_ADD(t + 1, taps.w[k] * rho_hat.y);
            SDFRT_ATOMIC_ADD(t + 2, taps.w[k] * rho_hat.z);
        }
    }
};

}  // namespace sdfrt
