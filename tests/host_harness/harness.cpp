// Developer harness: compiles the per-lane math of csrc/sdfrt_core.cuh / sdfrt_lane.cuh for the HOST so
// that CPU-only unit tests (-m "not gpu") can compare the kernel arithmetic against the oracle without
// a GPU.  It is test infrastructure: libsdfrt.so never links or calls it, and it is single-threaded.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../differentiable-sdf-rendering_b200/csrc/sdfrt_params.h"

using namespace sdfrt;

struct hh_scene {
    const float* sdf; int rx, ry, rz; float p[3];
    const float* albedo; int ax, ay, az;
    sdfrt_tracer_params tracer;
    sdfrt_warp_params warp;
    sdfrt_sensor sensor;
    sdfrt_shading shading;
    int integrator;
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
    sc.integrator = h->integrator;
    return sc;
}

extern "C" {

void hh_defaults(hh_scene* h) {
    std::memset(h, 0, sizeof(*h));
    tracer_defaults(&h->tracer);
    warp_defaults(&h->warp);
    sensor_defaults(&h->sensor);
    shading_defaults(&h->shading);
    h->integrator = INTEGRATOR_DIRECT;
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
    PassParams ps = make_pass(sc.S, seed, spp, rank, world);
    const size_t nf = (size_t)(sc.S.W + 2 * sc.S.border) * (sc.S.H + 2 * sc.S.border) * 4;
    std::memset(film, 0, nf * sizeof(float));
    LaneStats st = {0, 0, 0, 0, 0, 0, 0};
    const unsigned long long n = ps.lane_end - ps.lane_begin;
    for (unsigned long long lane = ps.lane_begin; lane < ps.lane_end; ++lane) {
        if (diff) lane_forward<true>(sc, ps, lane, film, rec, n, st);
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

int hh_backward(const hh_scene* h, unsigned seed, int spp, const float* film, const float* rec, const float* grad_image,
                float* grad_sdf, float* grad_albedo, float* grad_p, unsigned long long* stats, int rank, int world) {
    SceneParams sc = build(h);
    PassParams ps = make_pass(sc.S, seed, spp, rank, world);
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
    BackwardTargets tg; tg.grad_sdf = grad_sdf; tg.grad_albedo = grad_albedo;
    LaneStats st = {0, 0, 0, 0, 0, 0, 0};
    V3 gp = v3(0.f);
    const unsigned long long n = ps.lane_end - ps.lane_begin;
    for (unsigned long long lane = ps.lane_begin; lane < ps.lane_end; ++lane) {
        SiteList sites;
        lane_backward(sc, ps, lane, adj.data(), rec + (lane - ps.lane_begin), n, tg, gp, sites, st);
        for (int q = 0; q < sites.n; ++q) scatter_site(grad_sdf, sc.G, sites.s[q].x, sites.s[q].phi_hat, sites.s[q].g_hat);
    }
    if (grad_p) { grad_p[0] += gp.x; grad_p[1] += gp.y; grad_p[2] += gp.z; }
    if (stats) { stats[0] = n; stats[3] = st.n_hess; stats[4] = st.n_sites; }
    return 0;
}

int hh_rec_count() { return REC_COUNT; }
unsigned long long hh_local_lanes(const hh_scene* h, int spp, int rank, int world) {
    SensorParams S = make_sensor(h->sensor);
    PassParams ps = make_pass(S, 0, spp, rank, world);
    return ps.lane_end - ps.lane_begin;
}
}

// ---- developer statistics: how the march spends its evaluations (guides kernel design, not a test) ----
extern "C" int hh_march_profile(const hh_scene* h, unsigned seed, int spp, int stride, double* out) {
    SceneParams sc = build(h);
    PassParams ps = make_pass(sc.S, seed, spp, 0, 1);
    double n_rays = 0, n_evals = 0, n_interior = 0, n_same_cell = 0, n_refine = 0, n_hit = 0, n_same_cell_refine = 0, n_x_aligned = 0;
    for (unsigned long long lane = ps.lane_begin; lane < ps.lane_end; lane += stride) {
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
