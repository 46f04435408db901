// scripts/brick_analysis.cpp -- offline study for the SMEM brick staging experiment (not part of the library).
// For regions (= one 8-pixel film column x spp samples) of the c3 workload, replay the value marcher per ray on the host and record,
// per B-spline lookup, the base texel of its 4x4x4 block.  Reports (a) tap-cache hit rate (same block as the previous lookup of the
// ray), (b) for SHADOW rays: fraction of tap LOADS (cache misses) whose block lies inside a box of E^3 texels centred on the
// region's shadow-origin patch, (c) for PRIMARY rays: loads per distinct 8^3 brick of a region (reuse available to a brick cache).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <vector>
#include "../differentiable-sdf-rendering_b200/csrc/sdfrt_params.h"
using namespace sdfrt;

struct Rec { int bx, by, bz; };

static void march_record(const GridView& G, const TracerParams& T, V3 o, V3 d, float maxt, bool refine, std::vector<Rec>& out, float& its) {
    Segment s;
    bool active = segment_begin(s, T, o, d, maxt);
    while (active) {
        const V3 x = s.o + s.t * s.d;
        CellSetup<0> cs; cell_setup<0>(G, x, cs);
        out.push_back({cs.bx, cs.by, cs.bz});
        active = value_step(s, T, eval_value(G, x));
    }
    if (refine && isfinite(s.its_t)) {
        s.i = 0; bool r = true;
        while (r) { const V3 x = s.o + s.its_t * s.d; CellSetup<0> cs; cell_setup<0>(G, x, cs); out.push_back({cs.bx, cs.by, cs.bz}); r = refine_step(s, eval_value(G, x)); }
    }
    its = s.its_t;
}

int main(int argc, char** argv) {
    const int R = argc > 1 ? atoi(argv[1]) : 256, W = 512, spp = argc > 2 ? atoi(argv[2]) : 256;
    // blobs grid (synthetic.make_sdf 'blobs') is read from a raw float file written by the python driver
    std::vector<float> grid((size_t)R * R * R);
    FILE* f = fopen(argv[3], "rb"); if (!f || fread(grid.data(), 4, grid.size(), f) != grid.size()) { fprintf(stderr, "grid?\n"); return 1; } fclose(f);
    float tw[16]; f = fopen(argv[4], "rb"); if (!f || fread(tw, 4, 16, f) != 16) return 1; fclose(f);
    SceneParams sc; memset(&sc, 0, sizeof(sc));
    sdfrt_tracer_params tp; tracer_defaults(&tp); sc.T = make_tracer(tp);
    set_grid_transform(sc.G, sc.T, nullptr);
    sc.G.data = grid.data(); sc.G.rx = sc.G.ry = sc.G.rz = R;
    sdfrt_sensor sp; sensor_defaults(&sp); memcpy(sp.to_world, tw, sizeof(tw)); sp.width = W; sp.height = W; sc.S = make_sensor(sp);
    sdfrt_shading sh; shading_defaults(&sh); sc.Sh = make_shading(sh);
    PassParams ps = make_pass(sc.S, 1000, spp, 0, 1);
    const unsigned per_col = kBandRows * spp;
    const unsigned long long per_band = (unsigned long long)per_col * ps.Wb;
    double prim_evals = 0, prim_loads = 0, sh_evals = 0, sh_loads = 0, n_sh = 0, n_prim = 0;
    const int E[4] = {12, 16, 24, 32};
    double sh_in[4] = {0, 0, 0, 0}, sh_in_ev[4] = {0, 0, 0, 0};
    double patch_ext_sum = 0, nreg = 0;
    std::map<long long, int> brick_hist;   // loads per (region, brick) -> histogram buckets
    double prim_bricks = 0, prim_loads_b = 0;
    double prim_box[4] = {0, 0, 0, 0};
    // sample regions: band rows through the object, every 16th column
    for (int band = 20; band <= 44; band += 6)
        for (int col = 100; col < 420; col += 24) {
            const unsigned long long item0 = band * per_band + (unsigned long long)col * per_col;
            std::vector<std::vector<Rec>> prim(per_col), shad;
            std::vector<V3> so;
            for (unsigned k = 0; k < per_col; ++k) {
                const WorkItem w = decode_item(ps, item0 + k);
                CameraSample cs = camera_sample(sc, ps, w.lane);
                float its;
                march_record(sc.G, sc.T, cs.o, cs.d, cs.maxt, true, prim[k], its);
                if (!isfinite(its)) continue;
                const V3 dn = normalize(cs.d);
                const V3 p = cs.o + its * dn;
                float phi; V3 g; S3 H; eval_sdf<1>(sc.G, p, phi, g, H);
                const V3 n = normalize(g);
                const float u0 = pcg_next_f32(cs.rng), u1 = pcg_next_f32(cs.rng);
                const ShadowRay sr = make_shadow_ray(sc.Sh, p, n, u0, u1);
                if (!((-dot(n, dn) > 0.f) & (dot(n, sr.d) > 0.f))) continue;
                shad.emplace_back(); float its_s;
                march_record(sc.G, sc.T, sr.o, sr.d, sr.maxt, false, shad.back(), its_s);
                so.push_back(sr.o);
            }
            // primary stats
            std::map<long long, int> bricks;
            int pmin[3] = {1 << 30, 1 << 30, 1 << 30}, pmax[3] = {-1, -1, -1};
            for (auto& r : prim) {
                if (r.empty()) continue;
                ++n_prim;
                Rec prev{-99, -99, -99};
                for (auto& e : r) {
                    ++prim_evals;
                    if (e.bx != prev.bx || e.by != prev.by || e.bz != prev.bz) {
                        ++prim_loads;
                        bricks[((long long)(e.bz >> 3) << 40) | ((long long)(e.by >> 3) << 20) | (e.bx >> 3)]++;
                    }
                    prev = e;
                }
            }
            for (auto& b : bricks) { prim_bricks += 1; prim_loads_b += b.second; int h = b.second >= 256 ? 256 : (b.second >= 64 ? 64 : (b.second >= 16 ? 16 : (b.second >= 4 ? 4 : 1))); brick_hist[h] += b.second; }
            // shadow stats: box centred on the origin patch
            if (so.empty()) continue;
            float lo[3] = {1e9, 1e9, 1e9}, hi[3] = {-1e9, -1e9, -1e9};
            for (auto& o : so) { const float c[3] = {o.x * R - 0.5f, o.y * R - 0.5f, o.z * R - 0.5f}; for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], c[a]); hi[a] = std::max(hi[a], c[a]); } }
            patch_ext_sum += std::max(hi[0] - lo[0], std::max(hi[1] - lo[1], hi[2] - lo[2])); nreg += 1;
            for (auto& r : shad) {
                ++n_sh;
                Rec prev{-99, -99, -99};
                for (auto& e : r) {
                    ++sh_evals;
                    const bool load = (e.bx != prev.bx || e.by != prev.by || e.bz != prev.bz);
                    if (load) ++sh_loads;
                    for (int q = 0; q < 4; ++q) {
                        bool in = true;
                        const int b[3] = {e.bx, e.by, e.bz};
                        for (int a = 0; a < 3; ++a) { const float c = 0.5f * (lo[a] + hi[a]); const int b0 = (int)floorf(c - 0.5f * E[q]); in = in && (b[a] >= b0) && (b[a] + 3 < b0 + E[q]); }
                        if (in) { sh_in_ev[q] += 1; if (load) sh_in[q] += 1; }
                    }
                    prev = e;
                }
            }
        }
    printf("primary: rays %.0f evals %.0f loads %.0f (cache hit %.1f%%), distinct 8^3 bricks per region %.1f, loads per brick %.1f\n", n_prim, prim_evals, prim_loads,
           100 * (1 - prim_loads / prim_evals), prim_bricks / std::max(1.0, nreg), prim_loads_b / std::max(1.0, prim_bricks));
    for (auto& h : brick_hist) printf("  primary loads in bricks with >= %d loads per region: %.1f%%\n", h.first, 100.0 * h.second / prim_loads);
    printf("shadow: rays %.0f evals %.0f (%.1f per ray) loads %.0f (cache hit %.1f%%); origin patch extent %.1f texels (mean of max axis)\n", n_sh, sh_evals, sh_evals / n_sh, sh_loads,
           100 * (1 - sh_loads / sh_evals), patch_ext_sum / nreg);
    for (int q = 0; q < 4; ++q) printf("  box %d^3 texels around the patch centre: %.1f%% of the shadow tap loads, %.1f%% of the evals\n", E[q], 100 * sh_in[q] / sh_loads, 100 * sh_in_ev[q] / sh_evals);
    return 0;
}
