// oracle/cpu_port.cpp -- multi-threaded CPU build of the per-lane restatement, used ONLY as the timed CPU baseline
// (bench.py `cpu_baseline` / `--impl reference`).  TEST/BENCH INFRASTRUCTURE: the product never links it.
//
// The reference's own CPU path (Mitsuba `llvm_ad_rgb`, python/optimize.py:70-78) cannot be installed offline, and the
// parity oracle (oracle/sdf_oracle.py, torch autograd) is a vectorised-Python program that is ~100x slower than a
// compiled implementation would be -- timing it would flatter the GPU.  This file therefore compiles the per-lane
// forward/adjoint code (csrc/sdfrt_lane.cuh, `__host__ __device__`) with g++ -O3 -march=native -fopenmp so that the CPU
// baseline is a compiled, all-cores implementation of the same algorithm (same evaluation counts per ray).
// Parity is NOT judged against this file (it shares source with the kernels); it is judged against sdf_oracle.py.
#include <omp.h>

#include <chrono>
#include <cstring>
#include <vector>

#define SDFRT_ATOMIC_ADD(ptr, val)                \
    do {                                          \
        _Pragma("omp atomic") *(ptr) += (val);    \
    } while (0)
#include "../differentiable-sdf-rendering_b200/csrc/sdfrt_params.h"

using namespace sdfrt;

struct cp_scene {
    const float* sdf; int rx, ry, rz;
    const float* albedo; int ax, ay, az;
    sdfrt_tracer_params tracer;
    sdfrt_warp_params warp;
    sdfrt_sensor sensor;
    sdfrt_shading shading;
    int integrator;
    sdfrt_prb_params prb;   // used when integrator == INTEGRATOR_PRB
};

// per-sample film splat at the jittered pixel position (ImageBlock::put), atomics via SDFRT_ATOMIC_ADD
static void splat_at(const SceneParams& sc, const PassParams& ps, const WorkItem& w, V3 rgb, float* film) {
    Pcg32 rng = sampler_seed(ps.seed, (uint32_t)w.lane);
    const float r0 = pcg_next_f32(rng), r1 = pcg_next_f32(rng);
    const int b = sc.S.border, Wb = sc.S.W + 2 * b, Hb = sc.S.H + 2 * b;
    const float pfx = (float)(w.px - b) + r0 + ((float)b - 0.5f), pfy = (float)(w.py - b) + r1 + ((float)b - 0.5f);
    const float rad = sc.S.filter_radius;
    for (int by = (int)ceilf(pfy - rad); by <= (int)floorf(pfy + rad); ++by) {
        if (by < 0 || by >= Hb) continue;
        const float wy = filter_eval(sc.S, (float)by - pfy);
        for (int bx = (int)ceilf(pfx - rad); bx <= (int)floorf(pfx + rad); ++bx) {
            if (bx < 0 || bx >= Wb) continue;
            const float wgt = wy * filter_eval(sc.S, (float)bx - pfx);
            if (wgt == 0.f) continue;
            float* px = film + ((size_t)by * Wb + bx) * 4;
            SDFRT_ATOMIC_ADD(px + 0, rgb.x * wgt); SDFRT_ATOMIC_ADD(px + 1, rgb.y * wgt); SDFRT_ATOMIC_ADD(px + 2, rgb.z * wgt);
            SDFRT_ATOMIC_ADD(px + 3, wgt);
        }
    }
}

extern "C" {

void cp_defaults(cp_scene* h) {
    std::memset(h, 0, sizeof(*h));
    tracer_defaults(&h->tracer); warp_defaults(&h->warp); sensor_defaults(&h->sensor); shading_defaults(&h->shading);
    h->integrator = INTEGRATOR_DIRECT;
    prb_defaults(&h->prb);
}

int cp_threads() { return omp_get_max_threads(); }
void cp_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

// Sample k of `n_sample` -> work item: chunks of 64 consecutive items (the samples of a pixel stay together, as they would in a
// full render) spread evenly over the WHOLE wavefront of `n_items`, so that the sample sees the workload's mix of rays that miss
// the bounding box, cross empty space, graze and hit.
static inline unsigned long long strided_item(long long k, unsigned long long n_items, long long n_sample) {
    const long long chunk = 64;
    const unsigned long long total_chunks = (n_items + chunk - 1) / chunk;
    const unsigned long long sample_chunks = (unsigned long long)((n_sample + chunk - 1) / chunk);
    if (sample_chunks >= total_chunks) return (unsigned long long)k < n_items ? (unsigned long long)k : n_items - 1;
    const unsigned long long c = (unsigned long long)(k / chunk);
    const unsigned long long it = (c * total_chunks / sample_chunks) * chunk + (unsigned long long)(k % chunk);
    return it < n_items ? it : n_items - 1;
}

// Times `n_primal` primal lanes (seed, spp*mult) and `n_adjoint` adjoint lanes (forward with the differentiable tracer +
// closed-form backward with grad_image = 1) strided over the whole of the respective wavefronts (strided_item).
// out[0] = seconds primal, out[1] = seconds adjoint, out[2] = value evals, out[3] = hess evals, out[4] = camera rays that hit
int cp_time_sample(const cp_scene* h, unsigned seed, int spp_primal, int spp_adjoint, long long n_primal, long long n_adjoint,
                   float* grad_sdf /* rz*ry*rx, accumulated */, double* out) {
    SceneParams sc;
    std::memset(&sc, 0, sizeof(sc));
    sc.G.data = h->sdf; sc.G.rx = h->rx; sc.G.ry = h->ry; sc.G.rz = h->rz;
    sc.A.data = h->albedo; sc.A.rx = h->ax; sc.A.ry = h->ay; sc.A.rz = h->az;
    sc.A.cr = h->shading.albedo[0]; sc.A.cg = h->shading.albedo[1]; sc.A.cb = h->shading.albedo[2];
    sc.T = make_tracer(h->tracer); sc.W = make_warp(h->warp); sc.S = make_sensor(h->sensor); sc.Sh = make_shading(h->shading);
    sc.integrator = h->integrator;
    sc.prb = make_prb(h->prb);
    const bool prb = sc.integrator == INTEGRATOR_PRB;
    const size_t nf = (size_t)(sc.S.W + 2 * sc.S.border) * (sc.S.H + 2 * sc.S.border) * 4;
    std::vector<float> film(nf, 0.f), film_adj(nf, 0.f);
    double nv = 0, nh = 0, nhit = 0;
    // ---- primal sample
    PassParams pp = make_pass(sc.S, seed, spp_primal, 0, 1);
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : nv, nh, nhit)
    for (long long k = 0; k < n_primal; ++k) {
        const long long l = (long long)strided_item(k, pp.n_items, n_primal);
        LaneStats st = {0, 0, 0, 0, 0, 0, 0};
        if (prb) {
            const WorkItem w = decode_item(pp, (unsigned long long)l);
            NoEmit ne; V3 gp = v3(0.f);
            splat_at(sc, pp, w, lane_prb<false>(sc, pp, w.lane, nullptr, v3(0.f), ne, gp, st), film.data());
        } else {
            lane_forward<false>(sc, pp, decode_item(pp, (unsigned long long)l).lane, film.data(), nullptr, 0, st);
        }
        nv += st.n_value + st.n_grad; nh += st.n_hess; nhit += st.n_hit;
    }
    auto t1 = std::chrono::steady_clock::now();
    // ---- adjoint sample: forward (records) + backward
    PassParams pa = make_pass(sc.S, seed + 13, spp_adjoint, 0, 1);
    const long long la = 0;   // records are indexed by the sample index k
    const PassParams& sub = pa;
    std::vector<float> rec((size_t)REC_COUNT_MIS * (size_t)n_adjoint, 0.f);
    std::fill(film.begin(), film.end(), 0.f);
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : nv, nh, nhit)
    for (long long k = 0; k < n_adjoint; ++k) {
        const long long l = k;
        const unsigned long long item = strided_item(k, sub.n_items, n_adjoint);
        LaneStats st = {0, 0, 0, 0, 0, 0, 0};
        if (prb) {  // re-render of render_backward: keeps L per sample (rows 0..2 of the records)
            const WorkItem w = decode_item(sub, item);
            NoEmit ne; V3 gp = v3(0.f);
            const V3 L = lane_prb<false>(sc, sub, w.lane, nullptr, v3(0.f), ne, gp, st);
            float* r = rec.data() + (l - la);
            r[0] = L.x; r[n_adjoint] = L.y; r[2 * n_adjoint] = L.z;
            splat_at(sc, sub, w, L, film.data());
        } else {
            lane_forward<true>(sc, sub, decode_item(sub, item).lane, film.data(), rec.data() + (l - la), (unsigned long long)n_adjoint, st);
        }
        nv += st.n_value + st.n_grad; nh += st.n_hess; nhit += st.n_hit;
    }
    const int Wb = sc.S.W + 2 * sc.S.border, Hb = sc.S.H + 2 * sc.S.border;
    for (int i = 0; i < Wb * Hb; ++i) {  // grad_image = 1 (figures/benchmark/benchmark.py:44)
        const float* a = film.data() + (size_t)i * 4;
        if (a[3] == 0.f) continue;
        const float iw = 1.f / a[3];
        float* o = film_adj.data() + (size_t)i * 4;
        o[0] = o[1] = o[2] = iw;
        o[3] = -(a[0] + a[1] + a[2]) * iw * iw;
    }
    BackwardTargets tg; tg.grad_sdf = grad_sdf; tg.grad_albedo = nullptr; tg.queue = SiteQueue{nullptr, nullptr, nullptr, 0u};
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : nh)
    for (long long k = 0; k < n_adjoint; ++k) {
        const long long l = k;
        const unsigned long long item = strided_item(k, sub.n_items, n_adjoint);
        LaneStats st = {0, 0, 0, 0, 0, 0, 0};
        V3 gp = v3(0.f);
        if (prb) {
            const float* r = rec.data() + (l - la);
            DirectEmit em{grad_sdf, nullptr, sc.G};
            (void)lane_prb<true>(sc, sub, decode_item(sub, item).lane, film_adj.data(), V3{r[0], r[n_adjoint], r[2 * n_adjoint]}, em, gp, st);
            nh += st.n_hess;
            continue;
        }
        SiteList sites;
        lane_backward(sc, sub, decode_item(sub, item).lane, film_adj.data(), rec.data() + (l - la), (unsigned long long)n_adjoint, tg, gp, sites, st);
        for (int q = 0; q < sites.n; ++q) scatter_site(grad_sdf, sc.G, sites.s[q].x, sites.s[q].phi_hat, sites.s[q].g_hat);
        nh += st.n_hess;
    }
    auto t2 = std::chrono::steady_clock::now();
    out[0] = std::chrono::duration<double>(t1 - t0).count();
    out[1] = std::chrono::duration<double>(t2 - t1).count();
    out[2] = nv; out[3] = nh; out[4] = nhit;
    return 0;
}
}
