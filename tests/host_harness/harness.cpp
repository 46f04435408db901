// Developer harness: compiles the per-lane math of csrc/sdfrt_core.cuh / sdfrt_lane.cuh for the HOST so
// that CPU-only unit tests (-m "not gpu") can compare the kernel arithmetic against the oracle without
// a GPU.  It is test infrastructure: libsdfrt.so never links or calls it, and it is single-threaded.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../differentiable-sdf-rendering_b200/csrc/sdfrt_params.h"
#include "../../differentiable-sdf-rendering_b200/csrc/sdfrt_grid.cuh"

using namespace sdfrt;

struct hh_scene {
    const float* sdf; int rx, ry, rz; float p[3];
    const float* albedo; int ax, ay, az;
    sdfrt_tracer_params tracer;
    sdfrt_warp_params warp;
    sdfrt_sensor sensor;
    sdfrt_shading shading;
    int integrator;
    sdfrt_prb_params prb;
    int antithetic;   // ReparamIntegrator.antithetic_sampling (hh_forward / hh_backward / hh_local_lanes)
    int has_transform;   // Grid3d(data, transform): row-major 4x4 to_world
    float to_world[16];
    int shadow_origin;   // sdf_direct_reparam: 0 si, 1 detach_indirect_si, 2 decouple_reparam
    int use_mis;         // ReparamIntegrator prop use_mis (sdf_direct_reparam's BSDF-sampling branch)
    const void* envmap;  // hh_envmap_create handle or NULL (constant emitter)
    int env_has_rot;     // to_world rotation of the envmap (row-major 3x3)
    float env_R[9];
};

static SceneParams build(const hh_scene* h) {
    SceneParams sc;
    std::memset(&sc, 0, sizeof(sc));
    sc.G.data = h->sdf; sc.G.rx = h->rx; sc.G.ry = h->ry; sc.G.rz = h->rz;
    sc.G.px = h->p[0]; sc.G.py = h->p[1]; sc.G.pz = h->p[2];
    sc.A.data = h->albedo; sc.A.rx = h->ax; sc.A.ry = h->ay; sc.A.rz = h->az;
    sc.A.cr = h->shading.albedo[0]; sc.A.cg = h->shading.albedo[1]; sc.A.cb = h->shading.albedo[2];
    sc.T = make_tracer(h->tracer);
    sc.W = make_warp(h->warp);
    sc.S = make_sensor(h->sensor);
    sc.Sh = make_shading(h->shading);
    sc.Sh.shadow_origin = h->shadow_origin;
    sc.Sh.use_mis = h->use_mis;
    sc.integrator = h->integrator;
    sc.prb = make_prb(h->prb);
    set_grid_transform(sc.G, sc.T, h->has_transform ? h->to_world : nullptr);
    if (h->envmap != nullptr) {
        const EnvTables* t = (const EnvTables*)h->envmap;
        sc.E.data = t->data.data(); sc.E.marg_cdf = t->marg_cdf.data(); sc.E.cond_cdf = t->cond_cdf.data();
        sc.E.cell_prob = t->cell_prob.data(); sc.E.W1 = t->W1; sc.E.H = t->H;
        sc.E.has_rot = h->env_has_rot;
        for (int i = 0; i < 9; ++i) sc.E.R[i] = h->env_R[i];
    }
    return sc;
}

static int g_assign_rows = kBandRows;   // granularity of the row partition (sdfrt_set_partition_rows)

extern "C" {

void hh_set_assign_rows(int rows) { g_assign_rows = rows; }

// `envmap` emitter tables (the same builder sdfrt_set_envmap_host uses)
void* hh_envmap_create(const float* rgb, int w, int h, float scale) {
    EnvTables* t = new EnvTables();
    if (!build_envmap(rgb, w, h, scale, *t)) { delete t; return nullptr; }
    return t;
}
void hh_envmap_destroy(void* t) { delete (EnvTables*)t; }
// lookup / sampling of the emitter alone: out (n, 17) = Le rgb, pdf(d), sampled direction, its pdf, dLe_c / dd (3 x 3, row c)
int hh_envmap_query(const void* e, const float* R /* 9 floats or NULL */, const float* d, const float* u, int n, float* out) {
    const EnvTables* t = (const EnvTables*)e;
    EnvView E;
    E.has_rot = R != nullptr;
    for (int i = 0; i < 9; ++i) E.R[i] = R ? R[i] : (i % 4 == 0 ? 1.f : 0.f);
    E.data = t->data.data(); E.marg_cdf = t->marg_cdf.data(); E.cond_cdf = t->cond_cdf.data(); E.cell_prob = t->cell_prob.data();
    E.W1 = t->W1; E.H = t->H;
    for (int i = 0; i < n; ++i) {
        const V3 di = V3{d[3 * i], d[3 * i + 1], d[3 * i + 2]};
        V3 dLe[3];
        const V3 Le = env_eval(E, di, dLe);
        float pdf;
        const V3 s = env_sample(E, u[2 * i], u[2 * i + 1], pdf);
        float* o = out + (size_t)i * 17;
        o[0] = Le.x; o[1] = Le.y; o[2] = Le.z; o[3] = env_pdf(E, di); o[4] = s.x; o[5] = s.y; o[6] = s.z; o[7] = pdf;
        for (int c = 0; c < 3; ++c) { o[8 + 3 * c] = dLe[c].x; o[9 + 3 * c] = dLe[c].y; o[10 + 3 * c] = dLe[c].z; }
    }
    return 0;
}

void hh_defaults(hh_scene* h) {
    std::memset(h, 0, sizeof(*h));
    tracer_defaults(&h->tracer);
    warp_defaults(&h->warp);
    sensor_defaults(&h->sensor);
    shading_defaults(&h->shading);
    h->integrator = INTEGRATOR_DIRECT;
    prb_defaults(&h->prb);
}

int hh_eval_sdf(const hh_scene* h, const float* x, int n, int order, float* out) {
    SceneParams sc = build(h);
    for (int i = 0; i < n; ++i) {
        V3 p = V3{x[3 * i], x[3 * i + 1], x[3 * i + 2]};
        float v; V3 g; S3 H;
        if (order == 0) { eval_sdf<0>(sc.G, p, v, g, H); out[i] = v; }
        else if (order == 1) { eval_sdf<1>(sc.G, p, v, g, H); float* q = out + 4 * (size_t)i; q[0] = v; q[1] = g.x; q[2] = g.y; q[3] = g.z; }
        else { eval_sdf<2>(sc.G, p, v, g, H); float* q = out + 10 * (size_t)i; q[0] = v; q[1] = g.x; q[2] = g.y; q[3] = g.z;
               q[4] = H.xx; q[5] = H.yy; q[6] = H.zz; q[7] = H.xy; q[8] = H.xz; q[9] = H.yz; }
    }
    return 0;
}

int hh_trace(const hh_scene* h, const float* o, const float* d, const float* maxt, int n, int differentiable, float* its_t,
             float* warp_t, float* warp_t_d, float* warp_weight, float* warp_weight_d, int* n_steps) {
    SceneParams sc = build(h);
    for (int i = 0; i < n; ++i) {
        V3 ro = V3{o[3 * i], o[3 * i + 1], o[3 * i + 2]}, rd = V3{d[3 * i], d[3 * i + 1], d[3 * i + 2]};
        TraceOut tr; tr.warp_t = 0; tr.warp_t_d = v3(0.f); tr.warp_weight = 0; tr.warp_weight_d = v3(0.f); tr.n_steps = 0;
        int nh = 0, nv = 0;
        if (differentiable) trace_diff(sc.G, sc.T, ro, rd, maxt ? maxt[i] : INFINITY, true, tr, nh, nv);
        else tr.its_t = trace_non_diff(sc.G, sc.T, ro, rd, maxt ? maxt[i] : INFINITY, true, tr.n_steps, nv);
        its_t[i] = tr.its_t; warp_t[i] = tr.warp_t; warp_weight[i] = tr.warp_weight;
        warp_t_d[3 * i] = tr.warp_t_d.x; warp_t_d[3 * i + 1] = tr.warp_t_d.y; warp_t_d[3 * i + 2] = tr.warp_t_d.z;
        warp_weight_d[3 * i] = tr.warp_weight_d.x; warp_weight_d[3 * i + 1] = tr.warp_weight_d.y; warp_weight_d[3 * i + 2] = tr.warp_weight_d.z;
        if (n_steps) n_steps[i] = tr.n_steps;
    }
    return 0;
}

// forward pass over the whole wavefront; film: (Hb, Wb, 4) zeroed here; rec: REC_COUNT * n floats (diff only)
int hh_forward(const hh_scene* h, unsigned seed, int spp, int diff, float* film, float* rec, unsigned long long* stats, int rank, int world) {
    SceneParams sc = build(h);
    PassParams ps = make_pass(sc.S, seed, spp, rank, world, h->antithetic, g_assign_rows);
    const size_t nf = (size_t)(sc.S.W + 2 * sc.S.border) * (sc.S.H + 2 * sc.S.border) * 4;
    std::memset(film, 0, nf * sizeof(float));
    LaneStats st = {0, 0, 0, 0, 0, 0, 0};
    const unsigned long long n = ps.n_items;
    for (unsigned long long k = 0; k < n; ++k) {
        const unsigned long long lane = decode_item(ps, k).lane;
        if (diff) lane_forward<true>(sc, ps, lane, film, rec + k, n, st);
        else lane_forward<false>(sc, ps, lane, film, rec, n, st);
    }
    if (stats) { stats[0] = n; stats[1] = st.n_value; stats[2] = st.n_grad; stats[3] = st.n_hess; stats[4] = st.n_sites; stats[5] = st.n_steps_max; stats[6] = st.n_hit; }
    return 0;
}

int hh_develop(const hh_scene* h, const float* film, float* image) {
    SensorParams S = make_sensor(h->sensor);
    const int Wb = S.W + 2 * S.border;
    for (int y = 0; y < S.H; ++y)
        for (int x = 0; x < S.W; ++x) {
            const float* a = film + ((size_t)(y + S.border) * Wb + x + S.border) * 4;
            const float iw = a[3] == 0.f ? 1.f : 1.f / a[3];
            for (int c = 0; c < 3; ++c) image[((size_t)y * S.W + x) * 3 + c] = a[c] * iw;
        }
    return 0;
}

// lane window of hh_backward (the host twin of sdfrt_set_lane_window); (0, 0) = all lanes
static unsigned long long g_win_lo = 0ull, g_win_hi = ~0ull;
void hh_set_lane_window(unsigned long long lo, unsigned long long hi) {
    if (lo == 0ull && hi == 0ull) { g_win_lo = 0ull; g_win_hi = ~0ull; } else { g_win_lo = lo; g_win_hi = hi; }
}

int hh_backward(const hh_scene* h, unsigned seed, int spp, const float* film, const float* rec, const float* grad_image,
                float* grad_sdf, float* grad_albedo, float* grad_p, unsigned long long* stats, int rank, int world) {
    SceneParams sc = build(h);
    PassParams ps = make_pass(sc.S, seed, spp, rank, world, h->antithetic, g_assign_rows);
    const SensorParams& S = sc.S;
    const int Wb = S.W + 2 * S.border, Hb = S.H + 2 * S.border;
    std::vector<float> adj((size_t)Wb * Hb * 4, 0.f);
    for (int by = 0; by < Hb; ++by)
        for (int bx = 0; bx < Wb; ++bx) {
            const int qx = bx - S.border, qy = by - S.border;
            if (qx < 0 || qx >= S.W || qy < 0 || qy >= S.H) continue;
            const float* a = film + ((size_t)by * Wb + bx) * 4;
            const float* g = grad_image + ((size_t)qy * S.W + qx) * 3;
            if (a[3] == 0.f) continue;
            const float iw = 1.f / a[3];
            float* o = adj.data() + ((size_t)by * Wb + bx) * 4;
            o[0] = g[0] * iw; o[1] = g[1] * iw; o[2] = g[2] * iw;
            o[3] = -(g[0] * a[0] + g[1] * a[1] + g[2] * a[2]) * iw * iw;
        }
    BackwardTargets tg; tg.grad_sdf = grad_sdf; tg.grad_albedo = grad_albedo; tg.queue = SiteQueue{nullptr, nullptr, nullptr, 0u};
    LaneStats st = {0, 0, 0, 0, 0, 0, 0};
    V3 gp = v3(0.f);
    const unsigned long long n = ps.n_items;
    for (unsigned long long k = 0; k < n; ++k) {
        const unsigned long long lane = decode_item(ps, k).lane;
        if ((lane & ~kAntiBit) < g_win_lo || (lane & ~kAntiBit) >= g_win_hi) continue;
        SiteList sites;
        if (sc.W.kind == WARP_CONV) lane_backward<true>(sc, ps, lane, adj.data(), rec + k, n, tg, gp, sites, st);
        else lane_backward<false>(sc, ps, lane, adj.data(), rec + k, n, tg, gp, sites, st);
        for (int q = 0; q < sites.n; ++q) scatter_site(grad_sdf, sc.G, sites.s[q].x, sites.s[q].phi_hat, sites.s[q].g_hat);
    }
    if (grad_p) { grad_p[0] += gp.x; grad_p[1] += gp.y; grad_p[2] += gp.z; }
    if (stats) { stats[0] = n; stats[3] = st.n_hess; stats[4] = st.n_sites; }
    return 0;
}

// sdf_prb_reparam: primal sweep (film zeroed here; L_rec: 3 * n floats, may be null)
int hh_prb_forward(const hh_scene* h, unsigned seed, int spp, float* film, float* L_rec, int rank, int world) {
    SceneParams sc = build(h);
    PassParams ps = make_pass(sc.S, seed, spp, rank, world, 0, g_assign_rows);
    const int Wb = sc.S.W + 2 * sc.S.border, Hb = sc.S.H + 2 * sc.S.border;
    std::memset(film, 0, (size_t)Wb * Hb * 4 * sizeof(float));
    LaneStats st = {0, 0, 0, 0, 0, 0, 0};
    const unsigned long long n = ps.n_items;
    for (unsigned long long k = 0; k < n; ++k) {
        const WorkItem w = decode_item(ps, k);
        NoEmit ne;
        V3 gp = v3(0.f);
        const V3 L = lane_prb<false>(sc, ps, w.lane, nullptr, v3(0.f), ne, gp, st);
        if (L_rec) { L_rec[k] = L.x; L_rec[k + n] = L.y; L_rec[k + 2 * n] = L.z; }
        Pcg32 rng = sampler_seed(ps.seed, (uint32_t)w.lane);
        const float r0 = pcg_next_f32(rng), r1 = pcg_next_f32(rng);
        const float pfx = (float)(w.px - sc.S.border) + r0 + ((float)sc.S.border - 0.5f);
        const float pfy = (float)(w.py - sc.S.border) + r1 + ((float)sc.S.border - 0.5f);
        const float rad = sc.S.filter_radius;
        for (int by = (int)ceilf(pfy - rad); by <= (int)floorf(pfy + rad); ++by) {
            if (by < 0 || by >= Hb) continue;
            const float wy = filter_eval(sc.S, (float)by - pfy);
            for (int bx = (int)ceilf(pfx - rad); bx <= (int)floorf(pfx + rad); ++bx) {
                if (bx < 0 || bx >= Wb) continue;
                const float wgt = wy * filter_eval(sc.S, (float)bx - pfx);
                float* px = film + ((size_t)by * Wb + bx) * 4;
                px[0] += L.x * wgt; px[1] += L.y * wgt; px[2] += L.z * wgt; px[3] += wgt;
            }
        }
    }
    return 0;
}

int hh_prb_backward(const hh_scene* h, unsigned seed, int spp, const float* film, const float* L_rec, const float* grad_image,
                    float* grad_sdf, float* grad_albedo, float* grad_p, unsigned long long* stats, int rank, int world) {
    SceneParams sc = build(h);
    PassParams ps = make_pass(sc.S, seed, spp, rank, world, 0, g_assign_rows);
    const SensorParams& S = sc.S;
    const int Wb = S.W + 2 * S.border, Hb = S.H + 2 * S.border;
    std::vector<float> adj((size_t)Wb * Hb * 4, 0.f);
    for (int by = 0; by < Hb; ++by)
        for (int bx = 0; bx < Wb; ++bx) {
            const int qx = bx - S.border, qy = by - S.border;
            if (qx < 0 || qx >= S.W || qy < 0 || qy >= S.H) continue;
            const float* a = film + ((size_t)by * Wb + bx) * 4;
            const float* g = grad_image + ((size_t)qy * S.W + qx) * 3;
            if (a[3] == 0.f) continue;
            const float iw = 1.f / a[3];
            float* o = adj.data() + ((size_t)by * Wb + bx) * 4;
            o[0] = g[0] * iw; o[1] = g[1] * iw; o[2] = g[2] * iw;
            o[3] = -(g[0] * a[0] + g[1] * a[1] + g[2] * a[2]) * iw * iw;
        }
    LaneStats st = {0, 0, 0, 0, 0, 0, 0};
    V3 gp = v3(0.f);
    const unsigned long long n = ps.n_items;
    for (unsigned long long k = 0; k < n; ++k) {
        const WorkItem w = decode_item(ps, k);
        DirectEmit em{grad_sdf, grad_albedo, sc.G};
        (void)lane_prb<true>(sc, ps, w.lane, adj.data(), V3{L_rec[k], L_rec[k + n], L_rec[k + 2 * n]}, em, gp, st);
    }
    if (grad_p) { grad_p[0] += gp.x; grad_p[1] += gp.y; grad_p[2] += gp.z; }
    if (stats) { stats[0] = n; stats[3] = st.n_hess; stats[4] = st.n_sites; }
    return 0;
}

// ---- wavefront form of sdf_prb_reparam, emulated on the host: the same walker (lane_prb with records / stop hooks) that the
// device kernels call, with k_march replaced by the per-ray tracers.  recs: max_depth * 18 rows of n floats (PrbRecords layout).
static void hh_march_into(const SceneParams& sc, V3 o, V3 d, float maxt, bool refine, bool diff, float* rec, unsigned long long stride) {
    if (diff) {
        TraceOut tr; tr.warp_t = INFINITY; tr.warp_t_d = v3(0.f); tr.warp_weight = 0.f; tr.warp_weight_d = v3(0.f);
        int nh = 0, nv = 0;
        trace_diff(sc.G, sc.T, o, d, maxt, refine, tr, nh, nv);
        rec[0] = tr.its_t; rec[1 * stride] = tr.warp_t;
        rec[2 * stride] = tr.warp_t_d.x; rec[3 * stride] = tr.warp_t_d.y; rec[4 * stride] = tr.warp_t_d.z;
        rec[5 * stride] = tr.warp_weight;
        rec[6 * stride] = tr.warp_weight_d.x; rec[7 * stride] = tr.warp_weight_d.y; rec[8 * stride] = tr.warp_weight_d.z;
    } else {
        int ns = 0, nv = 0;
        rec[0] = trace_non_diff(sc.G, sc.T, o, d, maxt, refine, ns, nv);
    }
}

int hh_prb_wf_forward(const hh_scene* h, unsigned seed, int spp, int diff, float* film, float* L_rec, float* recs, int rank, int world) {
    SceneParams sc = build(h);
    PassParams ps = make_pass(sc.S, seed, spp, rank, world, 0, g_assign_rows);
    const int Wb = sc.S.W + 2 * sc.S.border, Hb = sc.S.H + 2 * sc.S.border;
    std::memset(film, 0, (size_t)Wb * Hb * 4 * sizeof(float));
    const unsigned long long n = ps.n_items;
    const bool warp2d = sc.W.kind == WARP_2D;
    const int rmd = warp2d ? (sc.W.max_reparam_depth >= 0 ? sc.W.max_reparam_depth : 100000) : sc.prb.max_depth;
    for (size_t q = 0; q < (size_t)n * sc.prb.max_depth * kPrbRecFloats; ++q) recs[q] = NAN;  // reading an unwritten row must show
    LaneStats st = {0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < sc.prb.max_depth; ++k) {
        for (unsigned long long i = 0; i < n; ++i) {
            const WorkItem w = decode_item(ps, i);
            float* base = recs + i;
            if (k == 0) {
                const CameraSample cs = camera_sample(sc, ps, w.lane);
                hh_march_into(sc, cs.o, cs.d, cs.maxt, true, diff && warp2d, base, n);
            }
            PrbStop stop;
            stop.stop_at = k; stop.shadow_active = false; stop.next_active = false;
            NoEmit ne; V3 gp = v3(0.f);
            (void)lane_prb<false>(sc, ps, w.lane, nullptr, v3(0.f), ne, gp, st, nullptr, PrbRecords{base, n, 9}, &stop);
            float* srec = base + (size_t)((k * 2 + 1) * 9) * n;
            if (stop.shadow_active) hh_march_into(sc, stop.shadow_o, stop.shadow_d, stop.shadow_maxt, false,
                                                  diff && warp2d && reparam_at_depth(sc.W, 1) && k <= rmd, srec, n);
            else srec[0] = INFINITY;
            if (k + 1 < sc.prb.max_depth) {
                float* nrec = base + (size_t)(((k + 1) * 2) * 9) * n;
                if (stop.next_active) hh_march_into(sc, stop.next_o, stop.next_d, INFINITY, true, diff && warp2d && (k + 1) <= rmd, nrec, n);
                else nrec[0] = INFINITY;
            }
        }
    }
    for (unsigned long long i = 0; i < n; ++i) {
        const WorkItem w = decode_item(ps, i);
        NoEmit ne; V3 gp = v3(0.f);
        const V3 L = lane_prb<false>(sc, ps, w.lane, nullptr, v3(0.f), ne, gp, st, nullptr, PrbRecords{recs + i, n, 9});
        if (L_rec) { L_rec[i] = L.x; L_rec[i + n] = L.y; L_rec[i + 2 * n] = L.z; }
        Pcg32 rng = sampler_seed(ps.seed, (uint32_t)w.lane);
        const float r0 = pcg_next_f32(rng), r1 = pcg_next_f32(rng);
        const float pfx = (float)(w.px - sc.S.border) + r0 + ((float)sc.S.border - 0.5f);
        const float pfy = (float)(w.py - sc.S.border) + r1 + ((float)sc.S.border - 0.5f);
        const float rad = sc.S.filter_radius;
        for (int by = (int)ceilf(pfy - rad); by <= (int)floorf(pfy + rad); ++by) {
            if (by < 0 || by >= Hb) continue;
            const float wy = filter_eval(sc.S, (float)by - pfy);
            for (int bx = (int)ceilf(pfx - rad); bx <= (int)floorf(pfx + rad); ++bx) {
                if (bx < 0 || bx >= Wb) continue;
                const float wgt = wy * filter_eval(sc.S, (float)bx - pfx);
                float* px = film + ((size_t)by * Wb + bx) * 4;
                px[0] += L.x * wgt; px[1] += L.y * wgt; px[2] += L.z * wgt; px[3] += wgt;
            }
        }
    }
    return 0;
}

int hh_prb_wf_backward(const hh_scene* h, unsigned seed, int spp, const float* film, const float* L_rec, const float* recs,
                       const float* grad_image, float* grad_sdf, float* grad_albedo, float* grad_p, int rank, int world) {
    SceneParams sc = build(h);
    PassParams ps = make_pass(sc.S, seed, spp, rank, world, 0, g_assign_rows);
    const SensorParams& S = sc.S;
    const int Wb = S.W + 2 * S.border, Hb = S.H + 2 * S.border;
    std::vector<float> adj((size_t)Wb * Hb * 4, 0.f);
    for (int by = 0; by < Hb; ++by)
        for (int bx = 0; bx < Wb; ++bx) {
            const int qx = bx - S.border, qy = by - S.border;
            if (qx < 0 || qx >= S.W || qy < 0 || qy >= S.H) continue;
            const float* a = film + ((size_t)by * Wb + bx) * 4;
            const float* g = grad_image + ((size_t)qy * S.W + qx) * 3;
            if (a[3] == 0.f) continue;
            const float iw = 1.f / a[3];
            float* o = adj.data() + ((size_t)by * Wb + bx) * 4;
            o[0] = g[0] * iw; o[1] = g[1] * iw; o[2] = g[2] * iw;
            o[3] = -(g[0] * a[0] + g[1] * a[1] + g[2] * a[2]) * iw * iw;
        }
    LaneStats st = {0, 0, 0, 0, 0, 0, 0};
    V3 gp = v3(0.f);
    const unsigned long long n = ps.n_items;
    for (unsigned long long k = 0; k < n; ++k) {
        const WorkItem w = decode_item(ps, k);
        DirectEmit em{grad_sdf, grad_albedo, sc.G};
        (void)lane_prb<true>(sc, ps, w.lane, adj.data(), V3{L_rec[k], L_rec[k + n], L_rec[k + 2 * n]}, em, gp, st, nullptr,
                             PrbRecords{recs + k, n, 9});
    }
    if (grad_p) { grad_p[0] += gp.x; grad_p[1] += gp.y; grad_p[2] += gp.z; }
    return 0;
}

int hh_rec_count() { return REC_COUNT_MIS; }
unsigned long long hh_local_lanes(const hh_scene* h, int spp, int rank, int world) {
    SensorParams S = make_sensor(h->sensor);
    PassParams ps = make_pass(S, 0, spp, rank, world, h->antithetic, g_assign_rows);
    return ps.n_items;
}
}

// ---- developer statistics: how the march spends its evaluations (guides kernel design, not a test) ----
extern "C" int hh_march_profile(const hh_scene* h, unsigned seed, int spp, int stride, double* out) {
    SceneParams sc = build(h);
    PassParams ps = make_pass(sc.S, seed, spp, 0, 1);
    double n_rays = 0, n_evals = 0, n_interior = 0, n_same_cell = 0, n_refine = 0, n_hit = 0, n_same_cell_refine = 0, n_x_aligned = 0;
    for (unsigned long long lane = 0; lane < ps.n_items; lane += stride) {
        CameraSample cs = camera_sample(sc, ps, lane);
        Segment s;
        bool active = segment_begin(s, sc.T, cs.o, cs.d, cs.maxt);
        CellCoord prev; prev.ix = -1000; prev.iy = prev.iz = 0;
        n_rays += 1;
        while (active) {
            V3 x = s.o + s.t * s.d;
            CellCoord c = cell_coord(sc.G, x);
            n_evals += 1; n_interior += c.interior; n_x_aligned += ((c.ix & 3) == 0);
            n_same_cell += (c.ix == prev.ix && c.iy == prev.iy && c.iz == prev.iz);
            prev = c;
            active = value_step(s, sc.T, eval_value(sc.G, x));
        }
        if (isfinite(s.its_t)) {
            n_hit += 1;
            s.i = 0; bool refining = true;
            while (refining) {
                V3 x = s.o + s.its_t * s.d;
                CellCoord c = cell_coord(sc.G, x);
                n_refine += 1; n_same_cell_refine += (c.ix == prev.ix && c.iy == prev.iy && c.iz == prev.iz);
                prev = c;
                refining = refine_step(s, eval_value(sc.G, x));
            }
        }
    }
    out[0] = n_rays; out[1] = n_evals; out[2] = n_interior; out[3] = n_same_cell; out[4] = n_refine; out[5] = n_hit; out[6] = n_same_cell_refine; out[7] = n_x_aligned;
    return 0;
}

// developer statistic: share of primary-march evaluations spent on rays that stay farther than `margin` from the surface
extern "C" int hh_miss_profile(const hh_scene* h, unsigned seed, int spp, int stride, float margin, double* out) {
    SceneParams sc = build(h);
    PassParams ps = make_pass(sc.S, seed, spp, 0, 1);
    double n_rays = 0, n_evals = 0, n_far_rays = 0, n_far_evals = 0, n_hit = 0, n_hit_evals = 0;
    for (unsigned long long lane = 0; lane < ps.n_items; lane += stride) {
        CameraSample cs = camera_sample(sc, ps, lane);
        Segment s;
        bool active = segment_begin(s, sc.T, cs.o, cs.d, cs.maxt);
        float minphi = 1e30f; int n = 0;
        while (active) { float v = eval_value(sc.G, s.o + s.t * s.d); minphi = fminf(minphi, v); ++n; active = value_step(s, sc.T, v); }
        n_rays += 1; n_evals += n;
        if (isfinite(s.its_t)) { n_hit += 1; n_hit_evals += n; }
        else if (minphi > margin) { n_far_rays += 1; n_far_evals += n; }
    }
    out[0] = n_rays; out[1] = n_evals; out[2] = n_far_rays; out[3] = n_far_evals; out[4] = n_hit; out[5] = n_hit_evals;
    return 0;
}

// ---- empty-space test (csrc pixel_provably_clear): host construction of the dilated brick minima + property check ----
// For every film pixel the test declares clear, EVERY sample ray of that pixel (all spp jitters of this seed) must never
// see an SDF value <= its own threshold (hence cannot hit, and has zero warp weight).  Returns the number of violations;
// out[0] = active rays in clear pixels, out[1] = active rays, out[2] = dilation used.
extern "C" int hh_check_empty_space(const hh_scene* h, unsigned seed, int spp, float skip_eps, double* out) {
    SceneParams sc = build(h);
    PassParams ps = make_pass(sc.S, seed, spp, 0, 1);
    const GridView& G = sc.G;
    const EmptySpaceParams e = empty_space_params(sc.S, G, skip_eps);
    out[2] = e.dilation;
    const int nx = (G.rx + kBrick - 1) / kBrick, ny = (G.ry + kBrick - 1) / kBrick, nz = (G.rz + kBrick - 1) / kBrick;
    std::vector<float> raw((size_t)nx * ny * nz, INFINITY), dil((size_t)nx * ny * nz, INFINITY);
    for (int z = 0; z < G.rz; ++z) for (int y = 0; y < G.ry; ++y) for (int x = 0; x < G.rx; ++x) {
        float& m = raw[((size_t)(z / kBrick) * ny + y / kBrick) * nx + x / kBrick];
        m = fminf(m, G.data[((size_t)z * G.ry + y) * G.rx + x]);
    }
    const int D = e.dilation;
    for (int bz = 0; bz < nz; ++bz) for (int by = 0; by < ny; ++by) for (int bx = 0; bx < nx; ++bx) {
        float m = INFINITY;
        for (int dz = -D; dz <= D; ++dz) for (int dy = -D; dy <= D; ++dy) for (int dx = -D; dx <= D; ++dx)
            m = fminf(m, raw[((size_t)clampi(bz + dz, 0, nz - 1) * ny + clampi(by + dy, 0, ny - 1)) * nx + clampi(bx + dx, 0, nx - 1)]);
        dil[((size_t)bz * ny + by) * nx + bx] = m;
    }
    BrickView B; B.data = dil.data(); B.nx = nx; B.ny = ny; B.nz = nz;
    const int Wb = sc.S.W + 2 * sc.S.border;
    int violations = 0; double n_clear = 0, n_rays = 0;
    for (unsigned long long lane = 0; lane < ps.n_items; ++lane) {
        const unsigned long long pix = lane / (unsigned long long)spp;
        const int py = (int)(pix / Wb), px = (int)(pix - (unsigned long long)py * Wb);
        CameraSample cs = camera_sample(sc, ps, lane);
        Segment s;
        bool active = segment_begin(s, sc.T, cs.o, cs.d, cs.maxt);
        if (!active) continue;
        n_rays += 1;
        if (!pixel_provably_clear(sc.S, G, sc.T, B, e, px, py)) continue;
        n_clear += 1;
        if (s.maxt > e.maxt_bound) ++violations;  // the bound used for the threshold must dominate every segment's maxt
        const float own = fmaxf(sc.T.trace_eps, skip_eps) * fmaxf(s.maxt, 1.f);
        while (active) {
            const float v = eval_value(G, s.o + s.t * s.d);
            if (!(v > own)) { ++violations; break; }
            active = value_step(s, sc.T, v);
        }
        if (isfinite(s.its_t)) ++violations;
    }
    out[0] = n_clear; out[1] = n_rays;
    return violations;
}

// developer statistic: shadow rays (emitter samples of hit lanes): occluded / unoccluded-monotone / unoccluded-non-monotone
extern "C" int hh_shadow_profile(const hh_scene* h, unsigned seed, int spp, int stride, double* out) {
    SceneParams sc = build(h);
    PassParams ps = make_pass(sc.S, seed, spp, 0, 1);
    double n_shadow = 0, n_occ = 0, n_mono = 0, n_nonmono = 0, ev_occ = 0, ev_mono = 0, ev_nonmono = 0;
    for (unsigned long long lane = 0; lane < ps.n_items; lane += stride) {
        CameraSample cs = camera_sample(sc, ps, lane);
        int ns = 0, nv = 0;
        const float its_t = trace_non_diff(sc.G, sc.T, cs.o, cs.d, cs.maxt, true, ns, nv);
        if (!isfinite(its_t)) continue;
        const V3 p = cs.o + its_t * normalize(cs.d);
        float phi; V3 g; S3 H;
        eval_sdf<1>(sc.G, p, phi, g, H);
        const V3 n = normalize(g);
        const float u0 = pcg_next_f32(cs.rng), u1 = pcg_next_f32(cs.rng);
        const ShadowRay sr = make_shadow_ray(sc.Sh, p, n, u0, u1);
        if (!((-dot(n, normalize(cs.d)) > 0.f) & (dot(n, sr.d) > 0.f))) continue;
        Segment s;
        bool active = segment_begin(s, sc.T, sr.o, sr.d, sr.maxt);
        float prev = 0.f; bool mono = true; int steps = 0;
        while (active) {
            const float v = eval_value(sc.G, s.o + s.t * s.d);
            if (!(fabsf(v) > prev)) mono = false;
            prev = fabsf(v); ++steps;
            active = value_step(s, sc.T, v);
        }
        n_shadow += 1;
        if (isfinite(s.its_t)) { n_occ += 1; ev_occ += steps; }
        else if (mono) { n_mono += 1; ev_mono += steps; }
        else { n_nonmono += 1; ev_nonmono += steps; }
    }
    out[0] = n_shadow; out[1] = n_occ; out[2] = n_mono; out[3] = n_nonmono; out[4] = ev_occ; out[5] = ev_mono; out[6] = ev_nonmono;
    return 0;
}

// Sequential emulation of the block-asynchronous re-distancing order of k_rd_solve (csrc/sdfrt_grid.cuh): same node
// arithmetic (rd_init_node / rd_solve are shared source), same tile size, round cap, work lists and wake rules; the tiles
// of a list are visited one after the other instead of by concurrent blocks.  Lets CPU-only tests check that this update
// order reaches the fixed point the oracle's sweeps reach.  Returns the number of work lists; visits_out = tile visits.
extern "C" int hh_redistance_tiles(const float* phi, const float* lower_bound, float* out, int rx, int ry, int rz, long long* visits_out) {
    RdGrid g;
    g.phi = phi; g.lower_bound = lower_bound; g.rx = rx; g.ry = ry; g.rz = rz;
    g.hx = 1.0f / (float)rx; g.hy = 1.0f / (float)ry; g.hz = 1.0f / (float)rz;
    const size_t sy = (size_t)rx, sz = (size_t)rx * ry, n = sz * rz;
    std::vector<float> u(n);
    std::vector<unsigned char> frozen(n);
    for (int z = 0; z < rz; ++z) for (int y = 0; y < ry; ++y) for (int x = 0; x < rx; ++x) {
        const size_t i = (size_t)z * sz + (size_t)y * sy + (size_t)x;
        const float d = rd_init_node(g, x, y, z);
        frozen[i] = d < kRdInf;
        u[i] = d;
    }
    const int T = kRdTile, P = T + 2;
    const int ntx = (rx + T - 1) / T, nty = (ry + T - 1) / T, ntz = (rz + T - 1) / T;
    const int n_tiles = ntx * nty * ntz;
    std::vector<int> list(n_tiles), next;
    for (int t = 0; t < n_tiles; ++t) list[t] = t;
    std::vector<unsigned char> flag(n_tiles, 0);
    std::vector<float> s(P * P * P), s2(P * P * P);
    long long visits = 0;
    int iters = 0;
    while (!list.empty()) {
        next.clear();
        std::fill(flag.begin(), flag.end(), 0);
        for (int tile : list) {
            const int tix = tile % ntx, tiy = (tile / ntx) % nty, tiz = tile / (ntx * nty);
            const int x0 = tix * T - 1, y0 = tiy * T - 1, z0 = tiz * T - 1;
            bool any_known = false;
            for (int q = 0; q < P * P * P; ++q) {
                const int gx = x0 + q % P, gy = y0 + (q / P) % P, gz = z0 + q / (P * P);
                float v = kRdInf;
                if (gx >= 0 && gx < rx && gy >= 0 && gy < ry && gz >= 0 && gz < rz) v = u[(size_t)gz * sz + (size_t)gy * sy + (size_t)gx];
                s[q] = v;
                any_known |= v < kRdInf;
            }
            if (!any_known) continue;
            ++visits;
            const std::vector<float> orig = s;
            bool capped = true;
            for (int round = 0; round < kRdMaxRounds; ++round) {
                bool any = false;
                s2 = s;
                for (int tz = 0; tz < T; ++tz) for (int ty = 0; ty < T; ++ty) for (int tx = 0; tx < T; ++tx) {
                    const int gx = x0 + 1 + tx, gy = y0 + 1 + ty, gz = z0 + 1 + tz;
                    if (gx >= rx || gy >= ry || gz >= rz) continue;
                    if (frozen[(size_t)gz * sz + (size_t)gy * sy + (size_t)gx]) continue;
                    const int me = ((tz + 1) * P + (ty + 1)) * P + (tx + 1);
                    const float cand = rd_solve(fminf(s[me - 1], s[me + 1]), rd_axis(g.hx), fminf(s[me - P], s[me + P]), rd_axis(g.hy),
                                                fminf(s[me - P * P], s[me + P * P]), rd_axis(g.hz));
                    if (cand < s[me]) { s2[me] = cand; any = true; }
                }
                s.swap(s2);
                if (!any) { capped = false; break; }
            }
            bool face[6] = {false, false, false, false, false, false};
            for (int tz = 0; tz < T; ++tz) for (int ty = 0; ty < T; ++ty) for (int tx = 0; tx < T; ++tx) {
                const int me = ((tz + 1) * P + (ty + 1)) * P + (tx + 1);
                if (!(s[me] < orig[me])) continue;
                const int gx = x0 + 1 + tx, gy = y0 + 1 + ty, gz = z0 + 1 + tz;
                u[(size_t)gz * sz + (size_t)gy * sy + (size_t)gx] = s[me];
                if (tx == 0) face[0] = true; if (tx == T - 1) face[1] = true;
                if (ty == 0) face[2] = true; if (ty == T - 1) face[3] = true;
                if (tz == 0) face[4] = true; if (tz == T - 1) face[5] = true;
            }
            const int dn[7][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}, {0, 0, 0}};
            for (int f = 0; f < 7; ++f) {
                if (!(f < 6 ? face[f] : capped)) continue;
                const int nx = tix + dn[f][0], ny = tiy + dn[f][1], nz = tiz + dn[f][2];
                if (nx < 0 || nx >= ntx || ny < 0 || ny >= nty || nz < 0 || nz >= ntz) continue;
                const int nt = (nz * nty + ny) * ntx + nx;
                if (!flag[nt]) { flag[nt] = 1; next.push_back(nt); }
            }
        }
        list.swap(next);
        ++iters;
    }
    for (size_t i = 0; i < n; ++i) out[i] = rd_input(g, i) < 0.0f ? -u[i] : u[i];
    if (visits_out) *visits_out = visits;
    return iters;
}
