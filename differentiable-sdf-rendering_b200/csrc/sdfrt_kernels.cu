// sdfrt_kernels.cu -- CUDA kernels (sm_100a) and the C ABI of libsdfrt.so (include/sdfrt.h).
//
// Kernels
//   k_march<DIFF>     (sdfrt_march.cuh) persistent-threads wavefront marcher: camera sample -> sphere trace
//                     through the tricubic SDF grid -> shading -> gaussian film splat.  DIFF=true runs the
//                     differentiable tracer (python/shapes.py:115-288) and stores its outputs per lane.
//   k_film_adjoint    per film pixel: (dL/dV, dL/dWt) of I = V / Wt (HDRFilm::develop).
//   k_backward        one lane per thread: closed-form adjoint + 64-tap atomic scatter into dL/d(sdf).
//   k_develop, k_trace_rays, k_eval_sdf  small helpers / test hooks.
// No CPU fallback exists: every entry point launches device work or fails.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "sdfrt_params.h"
#include "sdfrt_march.cuh"
#include "sdfrt_grid.cuh"

using namespace sdfrt;

namespace {

constexpr int kBlock = 128;

__device__ __forceinline__ unsigned warp_sum(unsigned v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Adjoint sweep 2: closed-form adjoint per lane, then the 64-tap scatter of its <= 3 sites.
// ncu on the one-atomic-per-tap version: issue active 15 %, lg_throttle 6.2 + long_scoreboard 5.8 stall cycles per issued
// instruction -- bound by the RED queue, not by arithmetic (profiles/r1_backward_direct_atomics.md).  Lanes of a warp are
// samples of one pixel, so many of their sites share the SAME 4x4x4 block.  Each lane therefore publishes its 64 tap
// values in shared memory ([lane][65], conflict free); lanes with equal block are grouped (ballot on the block key), the
// group's values are summed "transposed" (tap t by lane t mod 32) and one `red.global.add.f32` per tap and group is issued.
// (A per-warp 8^3 accumulation window with shared atomics was tried first and was slower: 10.1 vs 6.3 ms.)
struct RowStore {
    float* row;
    __device__ __forceinline__ void operator()(int t, float v) const { row[t] = v; }
};

// One round of the grouped transposed scatter: every lane of the warp brings at most one site.  Each publishes its 64 tap values in
// shared memory ([lane][65], conflict free); lanes with equal 4x4x4 block are grouped (ballot on the block key), the group's values
// are summed "transposed" (tap t by lane t mod 32) and ONE red.global.add.f32 per tap and group is issued.  Convergent: all 32 lanes.
__device__ __forceinline__ void grouped_scatter_round(const GridView& G, bool has, V3 x, float phi_hat, V3 g_hat, float (*vals)[65], float* grad_sdf) {
    const unsigned FULL = 0xffffffffu;
    const int lane_id = threadIdx.x & 31;
    unsigned rem = __ballot_sync(FULL, has);
    if (rem == 0u) return;
    int bx = 0, by = 0, bz = 0, key = -1;
    if (has) {
        site_tap_values(G, x, phi_hat, g_hat, bx, by, bz, RowStore{vals[lane_id]});
        key = (bz * G.ry + by) * G.rx + bx;
    }
    __syncwarp();
    while (rem) {
        const int leader = __ffs(rem) - 1;
        const int lkey = __shfl_sync(FULL, key, leader);
        const unsigned members = __ballot_sync(FULL, has && key == lkey);
        const int lbx = __shfl_sync(FULL, bx, leader), lby = __shfl_sync(FULL, by, leader), lbz = __shfl_sync(FULL, bz, leader);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int t = lane_id + 32 * h;
            float acc = 0.f;
            for (unsigned m = members; m; m &= m - 1) acc += vals[__ffs(m) - 1][t];
            if (acc != 0.f && grad_sdf != nullptr) {
                const int i = t & 3, j = (t >> 2) & 3, k = t >> 4;
                atomicAdd(grad_sdf + ((size_t)(lbz + k) * G.ry + (lby + j)) * (size_t)G.rx + (lbx + i), acc);
            }
        }
        rem &= ~members;
    }
    __syncwarp();
}

template <bool CONV>
__global__ void __launch_bounds__(kBlock) k_backward(const __grid_constant__ SceneParams sc, const __grid_constant__ PassParams ps,
                                                      const float* __restrict__ film_adj, const float* __restrict__ rec,
                                                      unsigned long long rec_stride, BackwardTargets tg, float* grad_p,
                                                      Counters* counters) {
    __shared__ float s_vals[kBlock / 32][32][65];
    const unsigned FULL = 0xffffffffu;
    const int lane_id = threadIdx.x & 31;
    float(*vals)[65] = s_vals[threadIdx.x >> 5];
    // records are stored in work-item (band) order; the lane index is needed for the sampler and the pixel
    const unsigned long long item = (unsigned long long)blockIdx.x * kBlock + threadIdx.x;
    LaneStats st = {0, 0, 0, 0, 0, 0, 0};
    V3 gp = v3(0.f);
    SiteList sites;
    sites.n = 0;
    if (item < ps.n_items) {
        const WorkItem w = decode_item(ps, item);
        const unsigned long long wl = w.lane & ~kAntiBit;
        if (wl >= ps.win_lo && wl < ps.win_hi) lane_backward<CONV>(sc, ps, w.lane, film_adj, rec + item, rec_stride, tg, gp, sites, st);
    }
    for (int q = 0; q < 5; ++q) {
        const bool has = q < sites.n;
        if (__ballot_sync(FULL, has) == 0u) break;
        grouped_scatter_round(sc.G, has, sites.s[q].x, sites.s[q].phi_hat, sites.s[q].g_hat, vals, tg.grad_sdf);
    }
    {   // the adjoint sweep's evaluations are counted separately from the marchers'
        const unsigned h = warp_sum(st.n_hess + (CONV ? st.n_value + st.n_grad : 0u)), nsites = warp_sum(st.n_sites), al = warp_sum(st.n_albedo);
        if (lane_id == 0) {
            if (h) atomicAdd(&counters->n_evals_adjoint, (unsigned long long)h);
            if (nsites) atomicAdd(&counters->n_scatter_sites, (unsigned long long)nsites);
            if (al) atomicAdd(&counters->n_albedo_lookups, (unsigned long long)al);
        }
    }
    if (grad_p != nullptr) {
        const float gx = warp_sum_f(gp.x), gy = warp_sum_f(gp.y), gz = warp_sum_f(gp.z);
        if (lane_id == 0) {
            if (gx != 0.f) atomicAdd(grad_p + 0, gx);
            if (gy != 0.f) atomicAdd(grad_p + 1, gy);
            if (gz != 0.f) atomicAdd(grad_p + 2, gz);
        }
    }
}

// I = V / Wt on the crop window  =>  dL/dV = G / Wt,  dL/dWt = -(G . I) / Wt   (zero on the border)
__global__ void k_film_adjoint(const float* __restrict__ film, const float* __restrict__ grad_image, float* __restrict__ film_adj,
                               int W, int H, int border) {
    const int Wb = W + 2 * border, Hb = H + 2 * border;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Wb * Hb) return;
    const int by = i / Wb, bx = i - by * Wb;
    const int qx = bx - border, qy = by - border;
    float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
    if (qx >= 0 && qx < W && qy >= 0 && qy < H) {
        const float4 a = reinterpret_cast<const float4*>(film)[i];
        const float* g = grad_image + ((size_t)qy * W + qx) * 3;
        if (a.w != 0.f) {
            const float iw = 1.f / a.w;
            out.x = g[0] * iw; out.y = g[1] * iw; out.z = g[2] * iw;
            out.w = -(g[0] * a.x + g[1] * a.y + g[2] * a.z) * iw * iw;
        }
    }
    reinterpret_cast<float4*>(film_adj)[i] = out;
}

__global__ void k_develop(const float* __restrict__ film, float* __restrict__ image, int W, int H, int border) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W * H) return;
    const int qy = i / W, qx = i - qy * W;
    const float4 a = reinterpret_cast<const float4*>(film)[(size_t)(qy + border) * (W + 2 * border) + (qx + border)];
    const float iw = a.w == 0.f ? 1.f : 1.f / a.w;
    image[(size_t)i * 3 + 0] = a.x * iw;
    image[(size_t)i * 3 + 1] = a.y * iw;
    image[(size_t)i * 3 + 2] = a.z * iw;
}

__global__ void k_ones(float* p, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 1.f;
}

__global__ void k_trace_rays(GridView G, TracerParams T, const float* __restrict__ o, const float* __restrict__ d,
                             const float* __restrict__ maxt, int n, int differentiable, sdfrt_ray_out out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const V3 ro = V3{o[3 * i], o[3 * i + 1], o[3 * i + 2]}, rd = V3{d[3 * i], d[3 * i + 1], d[3 * i + 2]};
    const float mt = maxt ? maxt[i] : INFINITY;
    TraceOut tr;
    tr.warp_t = 0.f; tr.warp_t_d = v3(0.f); tr.warp_weight = 0.f; tr.warp_weight_d = v3(0.f); tr.n_steps = 0;
    int nh = 0, nv = 0;
    if (differentiable) {
        trace_diff(G, T, ro, rd, mt, true, tr, nh, nv);
    } else {
        tr.its_t = trace_non_diff(G, T, ro, rd, mt, true, tr.n_steps, nv);
    }
    out.its_t[i] = tr.its_t;
    if (out.warp_t) out.warp_t[i] = tr.warp_t;
    if (out.warp_weight) out.warp_weight[i] = tr.warp_weight;
    if (out.warp_t_d) { out.warp_t_d[3 * i] = tr.warp_t_d.x; out.warp_t_d[3 * i + 1] = tr.warp_t_d.y; out.warp_t_d[3 * i + 2] = tr.warp_t_d.z; }
    if (out.warp_weight_d) { out.warp_weight_d[3 * i] = tr.warp_weight_d.x; out.warp_weight_d[3 * i + 1] = tr.warp_weight_d.y; out.warp_weight_d[3 * i + 2] = tr.warp_weight_d.z; }
    if (out.n_steps) out.n_steps[i] = tr.n_steps;
}

__global__ void k_eval_sdf(GridView G, const float* __restrict__ x, int n, int order, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const V3 p = V3{x[3 * i], x[3 * i + 1], x[3 * i + 2]};
    float v; V3 g; S3 H;
    if (order == 0) {
        eval_sdf<0>(G, p, v, g, H);
        out[i] = v;
    } else if (order == 1) {
        eval_sdf<1>(G, p, v, g, H);
        float* q = out + (size_t)i * 4;
        q[0] = v; q[1] = g.x; q[2] = g.y; q[3] = g.z;
    } else {
        eval_sdf<2>(G, p, v, g, H);
        float* q = out + (size_t)i * 10;
        q[0] = v; q[1] = g.x; q[2] = g.y; q[3] = g.z;
        q[4] = H.xx; q[5] = H.yy; q[6] = H.zz; q[7] = H.xy; q[8] = H.xz; q[9] = H.yz;
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Path replay integrator (sdfrt_prb.cuh).  Default = wavefront form: every segment of every path is marched by k_march, bounce by
// bounce (enqueue_batch), and kept as a per-vertex record; the kernels here only WALK records:
//   k_prb_wf_emit   walker as segment generator: the shadow and continuation rays leaving vertex k -> SegIn arrays for k_march
//   k_prb_primal    final walk: L per sample -> warp-cooperative film splat; the re-render of render_backward also keeps L
//   k_prb_backward  the replay: vertex adjoints from the records (no marching), scattered with red.global.add
// Per-lane form (SDFRT_PRB_WAVEFRONT=0; also what the host harness runs): the same two kernels march inline with the per-ray
// tracers, optionally taking the camera ray's march from k_march (SDFRT_PRB_PRIMARY).
// ---------------------------------------------------------------------------------------------------------------
// items [item0, item0 + cnt); `its` (or nullptr): hit distance of the camera ray per item of this range, from the wavefront marcher;
// L_rec (or nullptr): rows of `rec_stride` floats, already offset to item0
__global__ void __launch_bounds__(kBlock) k_prb_primal(const __grid_constant__ SceneParams sc, const __grid_constant__ PassParams ps,
                                                        unsigned long long item0, unsigned long long cnt, const float* __restrict__ its,
                                                        PrbRecords recs /* wavefront form: records of all segments, offset to item0 */,
                                                        float* __restrict__ film, float* __restrict__ L_rec, unsigned long long rec_stride,
                                                        Counters* counters) {
    __shared__ SplatSmem s_splat[kBlock / 32];
    const int lane_id = threadIdx.x & 31;
    const unsigned long long idx = (unsigned long long)blockIdx.x * kBlock + threadIdx.x;
    const unsigned long long item = item0 + idx;
    const bool valid = idx < cnt;
    LaneStats st = {0, 0, 0, 0, 0, 0, 0};
    V3 L = v3(0.f);
    float pfx = 0.f, pfy = 0.f;
    if (valid) {
        const WorkItem w = decode_item(ps, item);
        NoEmit ne;
        V3 gp = v3(0.f);
        PrbPrimary pre;
        if (its) { pre.its_t = its[idx]; pre.tr.warp_t = INFINITY; }
        PrbRecords my = recs;
        if (my.base) my.base += idx;
        L = lane_prb<false>(sc, ps, w.lane, nullptr, v3(0.f), ne, gp, st, its ? &pre : nullptr, my);
        if (L_rec) {
            L_rec[idx] = L.x; L_rec[idx + rec_stride] = L.y; L_rec[idx + 2 * rec_stride] = L.z;
        }
        // sample position of the splat: the (un-reparameterised) jittered pixel position re-projected (render(), :54-76)
        Pcg32 rng = sampler_seed(ps.seed, (uint32_t)w.lane);
        const float r0 = pcg_next_f32(rng), r1 = pcg_next_f32(rng);
        const bool anti = (w.lane & kAntiBit) != 0ull;
        pfx = sample_position(w.px - sc.S.border, r0, anti) + ((float)sc.S.border - 0.5f);
        pfy = sample_position(w.py - sc.S.border, r1, anti) + ((float)sc.S.border - 0.5f);
    }
    warp_splat(sc.S, film, s_splat[threadIdx.x >> 5], valid, pfx, pfy, L);
    const unsigned nvv = warp_sum(st.n_value), ng = warp_sum(st.n_grad), al = warp_sum(st.n_albedo), hits = warp_sum(st.n_hit);
    const unsigned nr = warp_sum(valid ? 1u : 0u);
    if (lane_id == 0) {
        atomicAdd(&counters->n_rays, (unsigned long long)nr);
        if (nvv) atomicAdd(&counters->n_evals_value, (unsigned long long)nvv);
        if (ng) atomicAdd(&counters->n_evals_grad, (unsigned long long)ng);
        if (al) atomicAdd(&counters->n_albedo_lookups, (unsigned long long)al);
        if (hits) atomicAdd(&counters->n_hits, (unsigned long long)hits);
    }
}

// Wavefront form, segment generator of bounce `k`: walks the path of every sample from the records of the marches done so far and
// writes the two rays that leave vertex k -- emitter-sample shadow ray and BSDF-sampled continuation -- as segments for k_march
// (inactive segments for samples whose path has ended or whose BSDF gate is closed).
__global__ void __launch_bounds__(kBlock) k_prb_wf_emit(const __grid_constant__ SceneParams sc, const __grid_constant__ PassParams ps,
                                                         unsigned long long item0, unsigned cnt, int k, PrbRecords recs,
                                                         SegIn* __restrict__ next_segs, SegIn* __restrict__ shadow_segs) {
    const unsigned idx = blockIdx.x * kBlock + threadIdx.x;
    if (idx >= cnt) return;
    SegIn dead;
    dead.a = make_float4(0.f, 0.f, 0.f, 0.f);
    dead.b = make_float4(0.f, 0.f, 1.f, -1.f);
    // a path emits at vertex k only if the segment that ends there hit the surface (three camera samples in four do not even reach
    // vertex 0): its record decides before anything of the path is re-walked
    if (!isfinite(recs.base[(size_t)((k * 2) * recs.rows) * recs.stride + idx])) {
        shadow_segs[idx] = dead;
        next_segs[idx] = dead;
        return;
    }
    const WorkItem w = decode_item(ps, item0 + idx);
    PrbStop stop;
    stop.stop_at = k; stop.shadow_active = false; stop.next_active = false;
    stop.shadow_o = v3(0.f); stop.shadow_d = v3(0.f, 0.f, 1.f); stop.next_o = v3(0.f); stop.next_d = v3(0.f, 0.f, 1.f); stop.shadow_maxt = 0.f;
    NoEmit ne;
    V3 gp = v3(0.f);
    LaneStats st = {0, 0, 0, 0, 0, 0, 0};
    PrbRecords my = recs;
    my.base += idx;
    (void)lane_prb<false>(sc, ps, w.lane, nullptr, v3(0.f), ne, gp, st, nullptr, my, &stop);
    shadow_segs[idx] = stop.shadow_active ? make_seg(sc.T, stop.shadow_o, stop.shadow_d, stop.shadow_maxt) : dead;
    next_segs[idx] = stop.next_active ? make_seg(sc.T, stop.next_o, stop.next_d, INFINITY) : dead;
}

// Site queue of the replay.  k_prb_backward used to scatter every B-spline site where the walker produced it: 64 red.global.add by ONE
// thread of a divergent warp, 22 of the kernel's 35 ms at c5 (scripts/prb_timing.py: 15 ms with the scatter switched off).  The walker
// now only appends the site (position, value adjoint, gradient adjoint: 32 B) to a queue -- opportunistically warp-aggregated, one
// atomic per converged group -- and k_scatter_sites drains the queue with full warps through the grouped transposed scatter of
// k_backward.  A site that does not fit into the queue is scattered on the spot, so the capacity is a tuning knob, not a limit.
// (struct SiteQueue: sdfrt_lane.cuh -- the convolution warp's adjoint appends to the same kind of queue.)
struct QueueEmit {
    float* grad_sdf;
    float* grad_albedo;
    GridView G;
    SiteQueue q;
    __device__ void site(V3 x, float phi_hat, V3 g_hat) {
        if (grad_sdf == nullptr) return;
        const unsigned m = __activemask();
        const int lane_id = threadIdx.x & 31, leader = __ffs(m) - 1;
        unsigned base = 0;
        if (lane_id == leader) base = atomicAdd(q.count, (unsigned)__popc(m));
        base = __shfl_sync(m, base, leader);
        const unsigned slot = base + (unsigned)__popc(m & ((1u << lane_id) - 1u));
        if (slot < q.capacity) {
            q.a[slot] = make_float4(x.x, x.y, x.z, phi_hat);
            q.b[slot] = make_float4(g_hat.x, g_hat.y, g_hat.z, 0.f);
        } else {
            prb_scatter_site(grad_sdf, G, x, phi_hat, g_hat);
        }
    }
    __device__ void albedo(const AlbedoTaps& taps, V3 rho_hat) {
        if (grad_albedo == nullptr) return;
        for (int k = 0; k < 8; ++k) {
            float* t = grad_albedo + (size_t)taps.idx[k] * 3;
            atomicAdd(t + 0, taps.w[k] * rho_hat.x);
            atomicAdd(t + 1, taps.w[k] * rho_hat.y);
            atomicAdd(t + 2, taps.w[k] * rho_hat.z);
        }
    }
};
__global__ void __launch_bounds__(kBlock) k_scatter_sites(const GridView G, SiteQueue q, float* __restrict__ grad_sdf) {
    __shared__ float s_vals[kBlock / 32][32][65];
    float(*vals)[65] = s_vals[threadIdx.x >> 5];
    const unsigned n = min(*q.count, q.capacity);
    const unsigned warps = gridDim.x * (kBlock / 32);
    const unsigned lane_id = threadIdx.x & 31;
    for (unsigned w = blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5); (unsigned long long)w * 32ull < n; w += warps) {
        const unsigned i = w * 32u + lane_id;
        const bool has = i < n;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (has) { a = q.a[i]; b = q.b[i]; }
        grouped_scatter_round(G, has, V3{a.x, a.y, a.z}, a.w, V3{b.x, b.y, b.z}, vals, grad_sdf);
    }
}

// Wavefront form, adjoint sweep 1 with the deferred differentiable march: the segments of one bounce were marched by the VALUE
// tracer with APPROACH (hit distance in row 0 of their records; a miss whose |phi| stopped growing somewhere comes back as -inf).
// Only segments that can end with a valid warp record go to the differentiable marcher: misses flagged -inf, and -- for
// continuation segments, whose record the replay reads whatever they hit -- hits.  They are appended (warp-aggregated) to the
// batch's list together with the offset of their record; every other segment gets warp_t = inf (rows 2..8 are then never used).
template <bool CONT>
__global__ void __launch_bounds__(256) k_prb_flag(const SegIn* __restrict__ segs, float* __restrict__ rec, unsigned long long stride, unsigned cnt,
                                                   unsigned dest_base, SegIn* __restrict__ fsegs, unsigned* __restrict__ fdest, unsigned* __restrict__ fcount) {
    const unsigned i = blockIdx.x * 256u + threadIdx.x;
    const unsigned FULL = 0xffffffffu;
    const int lane_id = threadIdx.x & 31;
    bool take = false;
    if (i < cnt) {
        const float its = rec[i];
        if (its == -INFINITY) { take = true; rec[i] = INFINITY; }
        else if (CONT && isfinite(its)) take = true;
        if (!take) rec[stride + i] = INFINITY;
    }
    const unsigned m = __ballot_sync(FULL, take);
    if (m == 0u) return;
    unsigned base = 0;
    if (lane_id == (__ffs(m) - 1)) base = atomicAdd(fcount, (unsigned)__popc(m));
    base = __shfl_sync(FULL, base, __ffs(m) - 1);
    if (take) {
        const unsigned slot = base + (unsigned)__popc(m & ((1u << lane_id) - 1u));
        fsegs[slot] = segs[i];
        fdest[slot] = dest_base + i;
    }
}

template <int MIN_BLOCKS, bool QUEUE>  // MIN_BLOCKS: resident blocks per SM the register allocation is tuned for (2: 255 registers, 3: 168)
__global__ void __launch_bounds__(kBlock, MIN_BLOCKS) k_prb_backward(const __grid_constant__ SceneParams sc, const __grid_constant__ PassParams ps,
                                                          const float* __restrict__ film_adj, const float* __restrict__ L_rec,
                                                          const float* __restrict__ prim_rec /* rows REC_ITS_T.. of the lane records, or nullptr */,
                                                          PrbRecords recs /* wavefront form: no marching in the replay */,
                                                          unsigned long long rec_stride, BackwardTargets tg, float* grad_p, Counters* counters,
                                                          SiteQueue queue) {
    const int lane_id = threadIdx.x & 31;
    const unsigned long long item = (unsigned long long)blockIdx.x * kBlock + threadIdx.x;
    LaneStats st = {0, 0, 0, 0, 0, 0, 0};
    V3 gp = v3(0.f);
    if (item < ps.n_items && (decode_item(ps, item).lane & ~kAntiBit) >= ps.win_lo && (decode_item(ps, item).lane & ~kAntiBit) < ps.win_hi) {
        const WorkItem w = decode_item(ps, item);
        const V3 L = V3{L_rec[item], L_rec[item + rec_stride], L_rec[item + 2 * rec_stride]};
        typename std::conditional<QUEUE, QueueEmit, DirectEmit>::type em{tg.grad_sdf, tg.grad_albedo, sc.G};
        if constexpr (QUEUE) em.q = queue;
        PrbPrimary pre;
        if (prim_rec) {
            const float* r = prim_rec + item;
            pre.its_t = r[REC_ITS_T * rec_stride];
            pre.tr.its_t = pre.its_t;
            pre.tr.warp_t = r[REC_TAU * rec_stride];
            pre.tr.warp_t_d = V3{r[REC_TAU_DX * rec_stride], r[REC_TAU_DY * rec_stride], r[REC_TAU_DZ * rec_stride]};
            pre.tr.warp_weight = r[REC_OMEGA * rec_stride];
            pre.tr.warp_weight_d = V3{r[REC_OMEGA_DX * rec_stride], r[REC_OMEGA_DY * rec_stride], r[REC_OMEGA_DZ * rec_stride]};
            pre.tr.n_steps = 0;
        }
        PrbRecords my = recs;
        if (my.base) my.base += item;
        (void)lane_prb<true>(sc, ps, w.lane, film_adj, L, em, gp, st, prim_rec ? &pre : nullptr, my);
    }
    const unsigned nvv = warp_sum(st.n_value), h = warp_sum(st.n_hess), nsites = warp_sum(st.n_sites), al = warp_sum(st.n_albedo),
                   ng = warp_sum(st.n_grad);
    if (lane_id == 0) {
        if (nvv) atomicAdd(&counters->n_evals_value, (unsigned long long)nvv);
        if (ng) atomicAdd(&counters->n_evals_grad, (unsigned long long)ng);
        if (h) atomicAdd(&counters->n_evals_adjoint, (unsigned long long)h);
        if (nsites) atomicAdd(&counters->n_scatter_sites, (unsigned long long)nsites);
        if (al) atomicAdd(&counters->n_albedo_lookups, (unsigned long long)al);
    }
    if (grad_p != nullptr) {
        const float gx = warp_sum_f(gp.x), gy = warp_sum_f(gp.y), gz = warp_sum_f(gp.z);
        if (lane_id == 0) {
            if (gx != 0.f) atomicAdd(grad_p + 0, gx);
            if (gy != 0.f) atomicAdd(grad_p + 1, gy);
            if (gz != 0.f) atomicAdd(grad_p + 2, gz);
        }
    }
}

}  // namespace

// ----------------------------------------------------------------------------------------------------
// context
// ----------------------------------------------------------------------------------------------------
constexpr int kViewSlots = 4;  // views in flight (adjoint-state slots)
#define VS(c) ((c)->view[(c)->cur_view])
constexpr int kSlots = 8;  // pipeline slots (internal streams + scratch sets) available to a forward sweep

struct sdfrt_ctx {
    int device = 0;
    std::string err;
    SceneParams sc;
    sdfrt_tracer_params tracer_pub;
    sdfrt_warp_params warp_pub;
    sdfrt_sensor sensor_pub;
    sdfrt_shading shading_pub;
    sdfrt_prb_params prb_pub;
    int rank = 0, world = 1;
    float to_world[16];      // Grid3d's transform (row-major) when has_xf
    bool has_xf = false;
    int shadow_origin = 0;   // sdf_direct_reparam props detach_indirect_si (1) / decouple_reparam (2)
    int assign_rows = kBandRows;   // granularity of the row partition (sdfrt_set_partition_rows)
    int use_mis = 0;         // ReparamIntegrator prop use_mis (python/integrators/reparam.py:17)
    float* env_buf = nullptr;   // `envmap` emitter: bitmap + sampling tables in one allocation (sc.E points into it)
    int env_has_rot = 0;        // to_world of the emitter (sdfrt_set_envmap_transform), kept across sdfrt_set_envmap_host
    float env_R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    int antithetic = 0;  // ReparamIntegrator.antithetic_sampling (python/integrators/reparam.py:19); not used by sdf_prb_reparam
    bool have_sdf = false;
    // Adjoint state of a VIEW: the lane records of adjoint sweep 1, the scene it was rendered with and the per-view scratch of the
    // empty-space test.  kViewSlots views can be in flight (sdfrt_set_view_slot): the views of one optimisation iteration are
    // independent (python/shape_opt.py:77-83), so a caller may enqueue the sweeps of the next view on another stream before the
    // previous view's backprop -- the persistent kernels of one view then fill the latency tails of the other.
    struct ViewSlot {
        float* rec = nullptr;
        size_t rec_capacity = 0;  // floats
        unsigned long long rec_stride = 0;
        PassParams adj_pass;
        bool adj_ready = false;
        unsigned long long generation = 0;
        SceneParams sc;               // snapshot taken by adjoint sweep 1, used by sweep 2
        float* film_adj = nullptr;
        size_t film_adj_capacity = 0;
        float* prb_rec = nullptr;     // sdf_prb_reparam, DIFF sweep: per-vertex march records [max_depth * 18][n_items]
        size_t prb_rec_n = 0;
        float *siteq_a = nullptr, *siteq_b = nullptr;        // site queue of k_prb_backward (one per view: backprops of views in flight overlap)
        size_t siteq_a_n = 0, siteq_b_n = 0;
        unsigned* siteq_count = nullptr;
        cudaStream_t defer_stream = nullptr;                 // the deferred differentiable march of this view runs beside the caller's stream:
        cudaEvent_t defer_ready = nullptr, defer_done = nullptr;   // only the adjoint (and whoever reuses the records) waits for it
        bool defer_pending = false;
        float *flag_segs = nullptr, *flag_dest = nullptr;   // deferred differentiable march: segments + record offsets, and
        size_t flag_segs_n = 0, flag_dest_n = 0;           // [0] their count, [1] the marcher's region counter
        unsigned* flag_count = nullptr;
        float *brick_a = nullptr, *brick_b = nullptr, *pix_mask = nullptr;  // dilated brick minima (ping-pong) + per-pixel masks
        size_t brick_a_n = 0, brick_b_n = 0, pix_mask_n = 0;
    } view[kViewSlots];
    int cur_view = 0;
    unsigned long long adj_generation = 0;   // bumped by every adjoint sweep 1; sdfrt_adjoint_token
    bool grid_dirty = true, rows4_valid = false, rows8_valid = false, brick_raw_valid = false;                  // sdfrt_set_sdf since the derived copies (rows4 / rows8 / brick minima) were built
    unsigned long long win_lo = 0ull, win_hi = ~0ull;  // sdfrt_set_lane_window
    unsigned* rd_err_host = nullptr;         // pinned copy of the re-distancing solver's error word (async callers)
    cudaEvent_t rd_err_event = nullptr;
    bool rd_err_pending = false;
    Counters* counters = nullptr;
    unsigned* work = nullptr;  // WQ_COUNT counters
    unsigned long long n_launches = 0;
    bool profiling = false;
    bool overlap = true;
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> timed;  // (class, start/stop) of launches not yet collected
    std::vector<cudaEvent_t> event_pool;
    // per-batch scratch of the wavefront pipeline
    struct Scratch {
        float *segs = nullptr, *its = nullptr, *ssegs = nullptr, *sinfo = nullptr, *s_its = nullptr, *srec = nullptr, *scount = nullptr, *prb_rec = nullptr, *sbox = nullptr,
              *bsegs = nullptr, *binfo = nullptr, *b_its = nullptr, *brec = nullptr, *bcount = nullptr;   // use_mis: BSDF-sampled rays
        size_t segs_n = 0, its_n = 0, ssegs_n = 0, sinfo_n = 0, s_its_n = 0, srec_n = 0, scount_n = 0, prb_rec_n = 0, sbox_n = 0, bsegs_n = 0, binfo_n = 0, b_its_n = 0, brec_n = 0, bcount_n = 0;
    } scratch[kSlots];            // one per pipeline slot
    cudaStream_t host_stream = nullptr;   // stream of the host-buffer entry points
    cudaStream_t aux[kSlots] = {};   // batches rotate over internal streams so that the tail of one batch's persistent
    cudaEvent_t ev_fork = nullptr, ev_join[kSlots] = {};  // kernels overlaps the following batches
    float* brick_raw = nullptr;   // per-brick SDF minima (undilated; rebuilt when the grid changes)
    size_t brick_raw_n = 0;
    bool skip_empty = true;
    bool defer_side_stream = false;  // ... on a stream of its own beside the caller's (SDFRT_DEFER_SIDE=1).  Measured: 4.88 vs 4.92 ms for one rank of
                                     // eight, 31.25 vs 30.92 ms at N = 1 -- nothing gained, so the march stays on the caller's stream
    bool defer_diff = true;     // deferred differentiable march of the shadow segments (SDFRT_DEFER=0: march them differentiably per batch)
    int min_live = kMinActive;  // slots that must be stepping for a warp to stay in the stepping loop
    int max_slots = 6;
    unsigned long long batch_target = 0;  // camera samples per batch; 0 = n / 6 clamped to [2^23, 2^24]
    int num_sms = 0;
    int march_blocks[6] = {0, 0, 0, 0, 0, 0};  // resident blocks per SM of k_march<DIFF, CACHE> at [3*DIFF + CACHE]
    // Row-expanded copy of the SDF grid for the marchers (GridView::rows4): 4x the memory, one 16-byte load per x-row of a
    // block instead of four 4-byte loads.  Pays while the copy stays mostly L2-resident (256^3: +18 %), loses at 512^3
    // (2 GiB copy, DRAM-bound: -12 %), hence the size gate.  rows4_roles = bitmask over the four marcher roles, -1 = by size.
    // grid-update scratch (sdfrt_grid.cuh)
    unsigned* rd_work = nullptr;             // [2][n_tiles] work lists + [2][n_tiles] flags
    size_t rd_work_n = 0;
    RdState* rd_state = nullptr;
    int rd_blocks = 0;                       // co-resident blocks of k_rd_solve
    float* lap_r = nullptr;
    size_t lap_r_n = 0;
    float4* rows4 = nullptr;
    size_t rows4_n = 0;
    TmaMap tmap;                             // 3-D tensor map of the SDF grid, box (kBoxX, kBoxY, kBoxZ): SMEM box staging of the shadow marchers
    const float* tmap_grid = nullptr;        // grid the map was encoded for
    int tmap_dims[3] = {0, 0, 0};
    int box_roles = 0;                       // role mask of the SMEM box staging (SDFRT_BOX; roles 1 and 3 = shadow segments), 0 = off
    bool box_on = false;
    float* rows8 = nullptr;                  // two-rows-per-256-bit-load copy (GridView::rows8), 8x the grid
    size_t rows8_n = 0;
    int rows8_roles = 15;                    // role mask (SDFRT_ROWS8; 0 = off): measured best at c3 with every role on the 8x copy
    bool rows8_on = false;
    bool shade_rows8 = false;  // k_shade's lookup from the 8x row copy (SDFRT_SHADE_ROWS8=1): measured 2755 vs 2757 Mrays/s at c3 -- no gain, off
    int rows4_roles = -1, march_role = 0;
    bool rows4_on = false;
    bool prb_wavefront = true;               // sdf_prb_reparam bounce by bounce through k_march (SDFRT_PRB_WAVEFRONT=0: per-lane walker)
    bool prb_primary_wavefront = true;       // camera rays of sdf_prb_reparam through k_march (SDFRT_PRB_PRIMARY=0: per lane)
    int prb_min_blocks = 2;                  // register-allocation variant of k_prb_backward (SDFRT_PRB_OCC)
    unsigned long long site_queue_cap = 1ull << 27;   // entries (SDFRT_SITEQ_CAP); sites beyond it are scattered in place
    bool site_queue = true;                  // k_prb_backward appends its scatter sites to a queue drained by k_scatter_sites (SDFRT_SITEQ=0: scatter in place)

    int cache_mode[4] = {1, 1, 0, 1};        // tap-cache mode of [value primary, value shadow, diff primary, diff shadow]
    // buffers of the host entry point
    float *h_sdf = nullptr, *h_film = nullptr, *h_film2 = nullptr, *h_image = nullptr, *h_gimg = nullptr, *h_grad = nullptr;
    float *h_alb = nullptr, *h_galb = nullptr;
    size_t h_sdf_n = 0, h_film_n = 0, h_film2_n = 0, h_image_n = 0, h_grad_n = 0, h_alb_n = 0, h_galb_n = 0;
};

namespace {

int fail(sdfrt_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}
#define SDFRT_CUDA(ctx, expr)                                                                                \
    do {                                                                                                     \
        cudaError_t e__ = (expr);                                                                            \
        if (e__ != cudaSuccess)                                                                              \
            return fail(ctx, SDFRT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));           \
    } while (0)

int ensure(sdfrt_ctx* ctx, float** buf, size_t* cap, size_t need) {
    if (*cap >= need) return SDFRT_OK;
    if (*buf) cudaFree(*buf);
    *buf = nullptr;
    *cap = 0;
    SDFRT_CUDA(ctx, cudaMalloc((void**)buf, need * sizeof(float)));
    *cap = need;
    return SDFRT_OK;
}

// RAII scope around one kernel launch: counts it and, when profiling, brackets it with events on `st`
struct LaunchScope {
    sdfrt_ctx* ctx;
    cudaStream_t st;
    cudaEvent_t e1 = nullptr;
    LaunchScope(sdfrt_ctx* c, int cls, cudaStream_t s) : ctx(c), st(s) {
        ++ctx->n_launches;
        if (!ctx->profiling) return;
        cudaEvent_t ev[2];
        for (int i = 0; i < 2; ++i) {
            if (!ctx->event_pool.empty()) { ev[i] = ctx->event_pool.back(); ctx->event_pool.pop_back(); }
            else if (cudaEventCreate(&ev[i]) != cudaSuccess) return;
        }
        cudaEventRecord(ev[0], st);
        e1 = ev[1];
        ctx->timed.push_back({cls, {ev[0], ev[1]}});
    }
    ~LaunchScope() { if (e1) cudaEventRecord(e1, st); }
};

size_t film_floats(const sdfrt_ctx* ctx) {
    const SensorParams& S = ctx->sc.S;
    return (size_t)(S.W + 2 * S.border) * (size_t)(S.H + 2 * S.border) * 4;
}

int rec_rows(const sdfrt_ctx* ctx) {   // rows of a lane record: the BSDF-sampled ray of the MIS branch adds nine
    return (ctx->sc.Sh.use_mis && ctx->sc.integrator == INTEGRATOR_DIRECT) ? (int)REC_COUNT_MIS : (int)REC_COUNT;
}

int check_ready(sdfrt_ctx* ctx, int spp) {
    if (!ctx) return SDFRT_ERR_INVALID;
    if (!ctx->have_sdf) return fail(ctx, SDFRT_ERR_STATE, "no SDF bound (sdfrt_set_sdf)");
    if (spp <= 0) return fail(ctx, SDFRT_ERR_INVALID, "spp must be positive");
    const SensorParams& S = ctx->sc.S;
    const unsigned long long wf = (unsigned long long)(S.W + 2 * S.border) * (S.H + 2 * S.border) * (unsigned long long)spp;
    // the sampler is seeded per 32-bit lane index (python/integrators/reparam.py:47-50: 2^30 on CUDA, 2^32 on LLVM)
    if (wf > 0xffffffffull) return fail(ctx, SDFRT_ERR_INVALID, "wavefront exceeds 2^32 lanes");
    return SDFRT_OK;
}


// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
int encode_grid_tmap(sdfrt_ctx* ctx) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
            return fail(ctx, SDFRT_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
        fn = (EncodeFn)p;
    }
    const GridView& G = ctx->sc.G;
    static_assert(sizeof(CUtensorMap) == sizeof(TmaMap), "tensor map size");
    const cuuint64_t dims[3] = {(cuuint64_t)G.rx, (cuuint64_t)G.ry, (cuuint64_t)G.rz};
    const cuuint64_t strides[2] = {(cuuint64_t)G.rx * 4ull, (cuuint64_t)G.rx * (cuuint64_t)G.ry * 4ull};
    const cuuint32_t box[3] = {(cuuint32_t)kBoxX, (cuuint32_t)kBoxY, (cuuint32_t)kBoxZ};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(reinterpret_cast<CUtensorMap*>(&ctx->tmap), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(G.data), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, SDFRT_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    ctx->tmap_grid = G.data;
    ctx->tmap_dims[0] = G.rx; ctx->tmap_dims[1] = G.ry; ctx->tmap_dims[2] = G.rz;
    return SDFRT_OK;
}

struct MarchExtra {   // optional arguments of a k_march launch
    const int4* region_box = nullptr;     // SMEM box staging (shadow segments)
    bool approach = false;                // value marcher: report misses whose |phi| stopped growing as -inf
    const unsigned* dest = nullptr;       // DIFF: record offset per segment
    const unsigned* n_slots_dev = nullptr;
    bool keep_row0 = false;               // DIFF: leave row 0 (its_t) of the record alone
};

template <bool DIFF, int CACHE>
void launch_march(sdfrt_ctx* ctx, unsigned grid, cudaStream_t st, const SegIn* segs, const unsigned* region_fill, unsigned n_slots, unsigned region,
                  int do_refine, float* out_its, float* rec_out, unsigned long long rec_stride, unsigned* region_counter,
                  const MarchExtra& ex = MarchExtra()) {
    GridView G = ctx->sc.G;
    const int role = ctx->march_role;
    const bool use8 = ctx->rows8_on && (ctx->rows8_roles >> role & 1);
    const bool use4 = !use8 && ctx->rows4_on && ((ctx->rows4_roles < 0 ? 15 : ctx->rows4_roles) >> role & 1);
    G.rows4 = use4 ? ctx->rows4 : nullptr;
    G.rows8 = use8 ? ctx->rows8 : nullptr;
#define SDFRT_MARCH_ARGS G, ctx->sc.T, segs, region_fill, n_slots, region, do_refine, ctx->min_live, out_its, rec_out, rec_stride, region_counter, ctx->counters, ctx->tmap
    if (ex.region_box != nullptr && CACHE != 2) {   // shadow segments with a staged SMEM box per region (TMA)
        const size_t smem = sizeof(BoxShared) + 128;
        if (use8) k_march<DIFF, CACHE, 8, true><<<grid, kMarchBlock, smem, st>>>(SDFRT_MARCH_ARGS, ex.region_box, ex.dest, ex.n_slots_dev, ex.keep_row0);
        else k_march<DIFF, CACHE, 4, true><<<grid, kMarchBlock, smem, st>>>(SDFRT_MARCH_ARGS, ex.region_box, ex.dest, ex.n_slots_dev, ex.keep_row0);
        return;
    }
    if constexpr (!DIFF && CACHE == 1) {
        if (ex.approach) {   // (only the register-cached value marcher is instantiated with APPROACH: the shadow role's mode)
            if (use8) k_march<false, 1, 8, false, true><<<grid, kMarchBlock, 0, st>>>(SDFRT_MARCH_ARGS, nullptr, ex.dest, ex.n_slots_dev, ex.keep_row0);
            else if (use4) k_march<false, 1, 4, false, true><<<grid, kMarchBlock, 0, st>>>(SDFRT_MARCH_ARGS, nullptr, ex.dest, ex.n_slots_dev, ex.keep_row0);
            else k_march<false, 1, 0, false, true><<<grid, kMarchBlock, 0, st>>>(SDFRT_MARCH_ARGS, nullptr, ex.dest, ex.n_slots_dev, ex.keep_row0);
            return;
        }
    }
    if (use8) k_march<DIFF, CACHE, 8><<<grid, kMarchBlock, 0, st>>>(SDFRT_MARCH_ARGS, nullptr, ex.dest, ex.n_slots_dev, ex.keep_row0);
    else if (use4) k_march<DIFF, CACHE, 4><<<grid, kMarchBlock, 0, st>>>(SDFRT_MARCH_ARGS, nullptr, ex.dest, ex.n_slots_dev, ex.keep_row0);
    else k_march<DIFF, CACHE, 0><<<grid, kMarchBlock, 0, st>>>(SDFRT_MARCH_ARGS, nullptr, ex.dest, ex.n_slots_dev, ex.keep_row0);
#undef SDFRT_MARCH_ARGS
}
template <bool DIFF>
void launch_march_mode(sdfrt_ctx* ctx, int role /* index into cache_mode */, unsigned need, cudaStream_t st, const SegIn* segs, const unsigned* region_fill,
                       unsigned n_slots, unsigned region, int do_refine, float* out_its, float* rec_out, unsigned long long rec_stride,
                       unsigned* region_counter, const MarchExtra& ex = MarchExtra()) {
    const int mode = ctx->cache_mode[role];
    ctx->march_role = role;
    int bps = ctx->march_blocks[(DIFF ? 3 : 0) + mode];
    if (ex.region_box != nullptr && mode != 2) bps = std::min(bps, 3);   // 58 KB of shared memory per block
    unsigned grid = (unsigned)ctx->num_sms * (unsigned)bps;
    if (grid > need) grid = need;
    if (mode == 0) launch_march<DIFF, 0>(ctx, grid, st, segs, region_fill, n_slots, region, do_refine, out_its, rec_out, rec_stride, region_counter, ex);
    else if (mode == 1) launch_march<DIFF, 1>(ctx, grid, st, segs, region_fill, n_slots, region, do_refine, out_its, rec_out, rec_stride, region_counter, ex);
    else launch_march<DIFF, 2>(ctx, grid, st, segs, region_fill, n_slots, region, do_refine, out_its, rec_out, rec_stride, region_counter, ex);
}  // camera samples per batch of the wavefront pipeline

// One forward sweep over this rank's camera samples: the primal render (value tracers) or adjoint sweep 1 (differentiable
// tracers + lane records).  Several sweeps can be enqueued together (run_sweeps): their batches then share the pipeline
// slots, so the latency tail of one sweep's persistent kernels is filled with the other's work.
constexpr size_t kRows4MaxTexels = (size_t)1 << 24;  // 256^3: 256 MiB expanded copy

struct Sweep {
    bool diff;
    PassParams ps;
    float* film;
    const unsigned char* mask;  // per-pixel result of the empty-space test (or nullptr)
    unsigned long long B;       // camera samples per batch
    FlagList flag;              // DIFF sweep: list for the deferred differentiable march of shadow / BSDF segments (segs == nullptr: off)
};

template <bool DIFF>
int enqueue_batch(sdfrt_ctx* ctx, const Sweep& sw, unsigned long long item0, int sl, cudaStream_t bs) {
    const PassParams& ps = sw.ps;
    const unsigned long long n = ps.n_items, B = sw.B;
    float* film = sw.film;
    sdfrt_ctx::Scratch& sc = ctx->scratch[sl];
    unsigned* work = ctx->work + sl * WQ_COUNT;
    ShadowQueue q;
    q.segs = reinterpret_cast<SegIn*>(sc.ssegs);
    q.info = reinterpret_cast<float4*>(sc.sinfo);
    q.count = reinterpret_cast<unsigned*>(sc.scount);
    const int shadow_role = (DIFF && reparam_at_depth(ctx->sc.W, 1)) ? 3 : 1;   // which marcher walks this batch's shadow segments
    const bool box_shadow = ctx->box_on && ctx->sc.integrator == INTEGRATOR_DIRECT && ((ctx->box_roles >> shadow_role) & 1);
    q.box = box_shadow ? reinterpret_cast<int4*>(sc.sbox) : nullptr;
    const bool mis = ctx->sc.Sh.use_mis && ctx->sc.integrator == INTEGRATOR_DIRECT;
    ShadowQueue qb;   // use_mis: the BSDF-sampled rays of the batch, region-compacted like the shadow segments
    qb.segs = reinterpret_cast<SegIn*>(sc.bsegs);
    qb.info = reinterpret_cast<float4*>(sc.binfo);
    qb.count = reinterpret_cast<unsigned*>(sc.bcount);
    qb.box = nullptr;
    const unsigned cnt = (unsigned)((n - item0 < B) ? (n - item0) : B);
    const unsigned blocks256 = (cnt + 255u) / 256u;
    // region size: large enough to keep a block's rays on one film column, small enough to balance the persistent grid
    unsigned region = kMaxRegion;
    while (region > 256u && (unsigned long long)cnt / region < 8ull * (unsigned long long)ctx->num_sms * 4ull) region >>= 1;
    const unsigned regions = (cnt + region - 1) / region;
    SDFRT_CUDA(ctx, cudaMemsetAsync(work, 0, WQ_COUNT * sizeof(unsigned), bs));
    { LaunchScope ls(ctx, 0, bs); k_gen_primary<<<blocks256, 256, 0, bs>>>(ctx->sc, ps, item0, cnt, reinterpret_cast<SegIn*>(sc.segs), sw.mask, ctx->counters); }
    // primary march (+ refine); DIFF writes its_t into row 0 of the lane records
    float* rec_b = DIFF ? VS(ctx).rec + item0 : nullptr;
    unsigned long long rec_stride = VS(ctx).rec_stride;
    float* its_out = sc.its;
    // sdf_prb_reparam, wavefront form: every segment's march lands in the per-vertex records (PrbRecords layout)
    const bool prb_wf = ctx->sc.integrator == INTEGRATOR_PRB && ctx->prb_wavefront;
    PrbRecords prec{nullptr, 0, 9};
    if (prb_wf) {
        if (DIFF) { prec.base = VS(ctx).prb_rec + item0; prec.stride = ps.n_items; prec.rows = 9; }
        else { prec.base = sc.prb_rec; prec.stride = B; prec.rows = 1; }
        rec_b = const_cast<float*>(prec.base);
        rec_stride = prec.stride;
        its_out = const_cast<float*>(prec.base);
    }
    const unsigned need = (cnt + kMarchBlock - 1) / kMarchBlock;
    { LaunchScope ls(ctx, 1, bs);
      if (DIFF && reparam_at_depth(ctx->sc.W, 0))
          launch_march_mode<true>(ctx, 2, need, bs, reinterpret_cast<const SegIn*>(sc.segs), nullptr, cnt, region, 1,
                                  nullptr, rec_b, rec_stride, work + WQ_PRIMARY_CHUNK);
      else  // value tracer; a DIFF pass without a primary warp (DummyWarpField / None) only needs its_t (record row 0)
          launch_march_mode<false>(ctx, 0, need, bs, reinterpret_cast<const SegIn*>(sc.segs), nullptr, cnt, region, 1,
                                   DIFF ? rec_b : its_out, nullptr, 0, work + WQ_PRIMARY_CHUNK); }
    const float* its = DIFF ? rec_b : its_out;
    if (prb_wf) {
        // bounce by bounce: k_prb_wf_emit regenerates the path from the records and emits the rays leaving vertex k, k_march
        // marches them (differentiable tracer where the replay will read a warp record), the final kernel walks the finished
        // records once more for L and splats.  No per-sample path state is stored: the walk is arithmetic + one lookup per vertex.
        const int max_depth = ctx->sc.prb.max_depth;
        const bool warp2d = ctx->sc.W.kind == WARP_2D;
        const int rmd = warp2d ? (ctx->sc.W.max_reparam_depth >= 0 ? ctx->sc.W.max_reparam_depth : 100000) : max_depth;
        SegIn* nsegs = reinterpret_cast<SegIn*>(sc.segs);
        SegIn* ssegs = reinterpret_cast<SegIn*>(sc.ssegs);
        float* rb = const_cast<float*>(prec.base);
        // Deferred differentiable march (as for the direct integrator's shadow segments): value march with APPROACH first, then only
        // the segments k_prb_flag selects go through the differentiable marcher, which fills rows 1..8 of their records in place.
        const bool defer = DIFF && warp2d && ctx->defer_diff && !ctx->box_on && ctx->cache_mode[1] == 1 &&
                           (unsigned long long)(max_depth * 2 * 9) * prec.stride + (unsigned long long)cnt <= 0xffffffffull;
        SegIn* fsegs = reinterpret_cast<SegIn*>(sc.bsegs);
        unsigned* fdest = reinterpret_cast<unsigned*>(sc.b_its);
        auto remarch = [&](int role, int v, int s_idx, const SegIn* segs_in, bool cont) -> int {
            float* rec_vs = rb + (size_t)((v * 2 + s_idx) * prec.rows) * prec.stride;
            const unsigned dest_base = (unsigned)((unsigned long long)((v * 2 + s_idx) * 9) * prec.stride);
            SDFRT_CUDA(ctx, cudaMemsetAsync(work + WQ_FLAG_COUNT, 0, 2 * sizeof(unsigned), bs));
            { LaunchScope ls(ctx, 2, bs);
              if (cont) k_prb_flag<true><<<(cnt + 255u) / 256u, 256, 0, bs>>>(segs_in, rec_vs, prec.stride, cnt, dest_base, fsegs, fdest, work + WQ_FLAG_COUNT);
              else k_prb_flag<false><<<(cnt + 255u) / 256u, 256, 0, bs>>>(segs_in, rec_vs, prec.stride, cnt, dest_base, fsegs, fdest, work + WQ_FLAG_COUNT); }
            MarchExtra ex;
            ex.dest = fdest; ex.n_slots_dev = work + WQ_FLAG_COUNT; ex.keep_row0 = true;
            LaunchScope ls(ctx, role == 3 ? 3 : 1, bs);
            launch_march_mode<true>(ctx, role, need, bs, fsegs, nullptr, cnt, 256u /* short lists: small work units */, 0, nullptr, rb, prec.stride, work + WQ_FLAG_CHUNK, ex);
            return SDFRT_OK;
        };
        for (int k = 0; k < max_depth; ++k) {
            { LaunchScope ls(ctx, 2, bs);
              k_prb_wf_emit<<<(cnt + kBlock - 1) / kBlock, kBlock, 0, bs>>>(ctx->sc, ps, item0, cnt, k, prec, nsegs, ssegs); }
            float* srec = rb + (size_t)((k * 2 + 1) * prec.rows) * prec.stride;
            SDFRT_CUDA(ctx, cudaMemsetAsync(work, 0, WQ_COUNT * sizeof(unsigned), bs));
            const bool shadow_diff = DIFF && warp2d && reparam_at_depth(ctx->sc.W, 1) && k <= rmd;
            MarchExtra exa;
            exa.approach = defer;
            { LaunchScope ls(ctx, 3, bs);
              if (shadow_diff && !defer)
                  launch_march_mode<true>(ctx, 3, need, bs, ssegs, nullptr, cnt, region, 0, nullptr, srec, prec.stride, work + WQ_SHADOW_CHUNK);
              else
                  launch_march_mode<false>(ctx, 1, need, bs, ssegs, nullptr, cnt, region, 0, srec, nullptr, 0, work + WQ_SHADOW_CHUNK, shadow_diff ? exa : MarchExtra()); }
            if (shadow_diff && defer) { const int rr = remarch(3, k, 1, ssegs, false); if (rr) return rr; }
            if (k + 1 < max_depth) {
                float* nrec = rb + (size_t)(((k + 1) * 2) * prec.rows) * prec.stride;
                const bool next_diff = DIFF && warp2d && (k + 1) <= rmd;
                { LaunchScope ls(ctx, 1, bs);
                  if (next_diff && !defer)
                      launch_march_mode<true>(ctx, 2, need, bs, nsegs, nullptr, cnt, region, 1, nullptr, nrec, prec.stride, work + WQ_PRIMARY_CHUNK);
                  else
                      launch_march_mode<false>(ctx, next_diff ? 1 : 0, need, bs, nsegs, nullptr, cnt, region, 1, nrec, nullptr, 0, work + WQ_PRIMARY_CHUNK, next_diff ? exa : MarchExtra()); }
                if (next_diff && defer) { const int rr = remarch(2, k + 1, 0, nsegs, true); if (rr) return rr; }
            }
        }
        { LaunchScope ls(ctx, 4, bs);
          k_prb_primal<<<(cnt + kBlock - 1) / kBlock, kBlock, 0, bs>>>(ctx->sc, ps, item0, (unsigned long long)cnt, nullptr, prec, film,
                                                                        DIFF ? VS(ctx).rec + item0 + (size_t)REC_S_STATE * VS(ctx).rec_stride : nullptr,
                                                                        VS(ctx).rec_stride, ctx->counters); }
        SDFRT_CUDA(ctx, cudaGetLastError());
        return SDFRT_OK;
    }
    if (ctx->sc.integrator == INTEGRATOR_PRB) {
        // path replay integrator, per-lane form: the camera ray was marched above by the wavefront marcher (empty-space test
        // included); the rest of the path is walked per sample.  The re-render of render_backward (DIFF) keeps L in rows 9..11.
        LaunchScope ls(ctx, 2, bs);
        k_prb_primal<<<(cnt + kBlock - 1) / kBlock, kBlock, 0, bs>>>(ctx->sc, ps, item0, (unsigned long long)cnt, its, PrbRecords{nullptr, 0, 9}, film,
                                                                      DIFF ? rec_b + (size_t)REC_S_STATE * VS(ctx).rec_stride : nullptr, VS(ctx).rec_stride, ctx->counters);
        SDFRT_CUDA(ctx, cudaGetLastError());
        return SDFRT_OK;
    }
    { LaunchScope ls(ctx, 2, bs);
      float* rs = DIFF ? rec_b + REC_S_STATE * VS(ctx).rec_stride : nullptr;
      float* rbs = (DIFF && mis) ? rec_b + REC_B_STATE * VS(ctx).rec_stride : nullptr;
      if (ctx->rows8_on && ctx->shade_rows8) {
          SceneParams scs = ctx->sc;
          scs.G.rows8 = ctx->rows8;
          k_shade<DIFF, true><<<regions, 256, 0, bs>>>(scs, ps, item0, cnt, region, reinterpret_cast<const SegIn*>(sc.segs), its, q, qb, film, rs, rbs, ctx->counters);
      } else {
          k_shade<DIFF, false><<<regions, 256, 0, bs>>>(ctx->sc, ps, item0, cnt, region, reinterpret_cast<const SegIn*>(sc.segs), its, q, qb, film, rs, rbs, ctx->counters);
      } }
    if (ctx->sc.integrator == INTEGRATOR_DIRECT) {
        // shadow march over the region-compacted queue.  The differentiable tracer is only needed if shadow rays are
        // reparameterised (python/warp.py:103: max_reparam_depth); otherwise the value tracer's its_t is all the pass uses.
        const bool shadow_diff = DIFF && reparam_at_depth(ctx->sc.W, 1);
        // Deferred differentiable march (sw.flag): the shadow segments of a DIFF sweep are first marched by the VALUE tracer, which
        // also reports whether |phi| ever stopped growing along the segment; only those (about 1 in 10, and only if unoccluded) can
        // end with a valid warp record and are collected for one differentiable march at the end of the sweep (run_sweeps).
        const bool deferred = shadow_diff && sw.flag.segs != nullptr && ctx->cache_mode[1] == 1 && !box_shadow;
        const bool shadow_diff_now = shadow_diff && !deferred;
        MarchExtra ex;
        ex.region_box = q.box;
        ex.approach = deferred;
        { LaunchScope ls(ctx, 3, bs);
          if (shadow_diff_now)
              launch_march_mode<true>(ctx, 3, need, bs, reinterpret_cast<const SegIn*>(sc.ssegs), q.count,
                                      regions * region, region, 0, nullptr, sc.srec, B, work + WQ_SHADOW_CHUNK, ex);
          else
              launch_march_mode<false>(ctx, 1, need, bs, reinterpret_cast<const SegIn*>(sc.ssegs), q.count,
                                       regions * region, region, 0, sc.s_its, nullptr, 0, work + WQ_SHADOW_CHUNK, ex); }
        const float* shadow_its = shadow_diff_now ? sc.srec : sc.s_its;
        { LaunchScope ls(ctx, 4, bs);
          k_shadow_splat<DIFF><<<regions, 256, 0, bs>>>(ctx->sc, ps, item0, region, reinterpret_cast<const SegIn*>(sc.segs), q.count,
                                                        reinterpret_cast<const float4*>(sc.sinfo), shadow_its, sc.srec, B, film, rec_b,
                                                        VS(ctx).rec_stride, reinterpret_cast<const SegIn*>(sc.ssegs),
                                                        deferred ? sw.flag : FlagList{nullptr, nullptr, nullptr, 0u}); }
        if (mis) {   // BSDF-sampled rays: same marchers (no refinement: only hit / escape and the warp record matter), second splat
            SDFRT_CUDA(ctx, cudaMemsetAsync(work + WQ_SHADOW_CHUNK, 0, sizeof(unsigned), bs));
            MarchExtra exb;
            exb.approach = deferred;
            { LaunchScope ls(ctx, 3, bs);
              if (shadow_diff_now)
                  launch_march_mode<true>(ctx, 3, need, bs, reinterpret_cast<const SegIn*>(sc.bsegs), qb.count,
                                          regions * region, region, 0, nullptr, sc.brec, B, work + WQ_SHADOW_CHUNK, exb);
              else
                  launch_march_mode<false>(ctx, 1, need, bs, reinterpret_cast<const SegIn*>(sc.bsegs), qb.count,
                                           regions * region, region, 0, sc.b_its, nullptr, 0, work + WQ_SHADOW_CHUNK, exb); }
            const float* b_its = shadow_diff_now ? sc.brec : sc.b_its;
            { LaunchScope ls(ctx, 4, bs);
              k_shadow_splat<DIFF, true><<<regions, 256, 0, bs>>>(ctx->sc, ps, item0, region, reinterpret_cast<const SegIn*>(sc.segs), qb.count,
                                                                  reinterpret_cast<const float4*>(sc.binfo), b_its, sc.brec, B, film, rec_b,
                                                                  VS(ctx).rec_stride, reinterpret_cast<const SegIn*>(sc.bsegs),
                                                                  deferred ? sw.flag : FlagList{nullptr, nullptr, nullptr, 0u}); }
        }
    }
    SDFRT_CUDA(ctx, cudaGetLastError());
    return SDFRT_OK;
}

int wait_deferred(sdfrt_ctx* ctx, cudaStream_t st, int slot);

int run_sweeps(sdfrt_ctx* ctx, Sweep* sweeps, int ns, cudaStream_t st) {
    SDFRT_CUDA(ctx, cudaSetDevice(ctx->device));
    {   // a deferred march still in flight reads this view slot's records and flag list, and -- any slot's -- the derived grid copies
        const int rcw = wait_deferred(ctx, st, ctx->grid_dirty ? -1 : ctx->cur_view);
        if (rcw) return rcw;
    }
    if (ctx->sc.W.kind == WARP_CONV && ctx->sc.integrator == INTEGRATOR_PRB)
        return fail(ctx, SDFRT_ERR_INVALID, "WarpFieldConvolution is implemented for the direct-illumination integrators only (not sdf_prb_reparam)");
    if (ctx->sc.E.data != nullptr && ctx->sc.integrator == INTEGRATOR_PRB)
        return fail(ctx, SDFRT_ERR_INVALID, "the envmap emitter is implemented for sdf_direct_reparam only (sdf_prb_reparam: constant emitter)");
    const SensorParams& S = ctx->sc.S;
    int rc;
    unsigned long long Bmax = 0, total_batches = 0;
    bool any_diff = false, any_value = false;
    for (int k = 0; k < ns; ++k) {
        Sweep& sw = sweeps[k];
        SDFRT_CUDA(ctx, cudaMemsetAsync(sw.film, 0, film_floats(ctx) * sizeof(float), st));
        sw.mask = nullptr;
        sw.B = 0;
        sw.flag = FlagList{nullptr, nullptr, nullptr, 0u};
        const unsigned long long n = sw.ps.n_items;
        if (n == 0) continue;
        if ((unsigned long long)kBandRows * (unsigned long long)(S.W + 2 * S.border) * (unsigned long long)sw.ps.spp * (1ull + sw.ps.antithetic) > 0xffffffffull)
            return fail(ctx, SDFRT_ERR_INVALID, "film row band exceeds 2^32 samples");
        unsigned long long target = ctx->batch_target;
        if (target == 0) target = std::max<unsigned long long>(1ull << 23, std::min<unsigned long long>(1ull << 24, n / 6));   // measured: fewer, larger batches
                                                                        // win once a rank holds < 50 M samples (profiles/r2_call4)
        const unsigned long long nbatch = (n + target - 1) / target;
        unsigned long long B = (n + nbatch - 1) / nbatch;
        sw.B = (B + kMaxRegion - 1) / kMaxRegion * kMaxRegion;  // whole regions per batch
        Bmax = std::max(Bmax, sw.B);
        total_batches += (n + sw.B - 1) / sw.B;
        (sw.diff ? any_diff : any_value) = true;
    }
    if (total_batches == 0) return SDFRT_OK;
    // list of the deferred differentiable march (per view slot; one DIFF sweep per call)
    for (int k = 0; k < ns; ++k) {
        Sweep& sw = sweeps[k];
        const unsigned long long n = sw.ps.n_items;
        if (!sw.diff || n == 0 || !ctx->defer_diff || ctx->sc.integrator != INTEGRATOR_DIRECT || !reparam_at_depth(ctx->sc.W, 1)) continue;
        const unsigned long long rows = (unsigned long long)rec_rows(ctx);
        if (n * rows > 0xffffffffull) continue;                       // record offsets are 32-bit
        const bool mis = ctx->sc.Sh.use_mis;
        const unsigned long long cap = n * (mis ? 2ull : 1ull);       // every sample could contribute a shadow (and a BSDF) segment
        if (cap > 0xffffffffull) continue;
        sdfrt_ctx::ViewSlot& V = VS(ctx);
        if ((rc = ensure(ctx, &V.flag_segs, &V.flag_segs_n, (size_t)cap * 8))) return rc;
        if ((rc = ensure(ctx, &V.flag_dest, &V.flag_dest_n, (size_t)cap))) return rc;
        if (!V.flag_count) SDFRT_CUDA(ctx, cudaMalloc((void**)&V.flag_count, 4 * sizeof(unsigned)));
        SDFRT_CUDA(ctx, cudaMemsetAsync(V.flag_count, 0, 4 * sizeof(unsigned), st));
        sw.flag = FlagList{reinterpret_cast<SegIn*>(V.flag_segs), reinterpret_cast<unsigned*>(V.flag_dest), V.flag_count, (unsigned)cap};
    }
    const int nslots = (total_batches > 1 && ctx->overlap) ? (int)std::min<unsigned long long>(total_batches, (unsigned long long)ctx->max_slots) : 1;
    for (int sl = 0; sl < nslots; ++sl) {
        sdfrt_ctx::Scratch& sc = ctx->scratch[sl];
        if ((rc = ensure(ctx, &sc.segs, &sc.segs_n, (size_t)Bmax * 8))) return rc;
        if ((rc = ensure(ctx, &sc.ssegs, &sc.ssegs_n, (size_t)Bmax * 8))) return rc;
        if ((rc = ensure(ctx, &sc.sinfo, &sc.sinfo_n, (size_t)Bmax * 4))) return rc;
        if ((rc = ensure(ctx, &sc.scount, &sc.scount_n, (size_t)(Bmax / 256) + 1))) return rc;
        if ((rc = ensure(ctx, &sc.sbox, &sc.sbox_n, 4 * ((size_t)(Bmax / 256) + 1)))) return rc;
        if ((rc = ensure(ctx, &sc.s_its, &sc.s_its_n, (size_t)Bmax))) return rc;
        if (any_diff && (rc = ensure(ctx, &sc.srec, &sc.srec_n, (size_t)Bmax * 9))) return rc;
        if (ctx->sc.Sh.use_mis && ctx->sc.integrator == INTEGRATOR_DIRECT) {
            if ((rc = ensure(ctx, &sc.bsegs, &sc.bsegs_n, (size_t)Bmax * 8))) return rc;
            if ((rc = ensure(ctx, &sc.binfo, &sc.binfo_n, (size_t)Bmax * 4))) return rc;
            if ((rc = ensure(ctx, &sc.bcount, &sc.bcount_n, (size_t)(Bmax / 256) + 1))) return rc;
            if ((rc = ensure(ctx, &sc.b_its, &sc.b_its_n, (size_t)Bmax))) return rc;
            if (any_diff && (rc = ensure(ctx, &sc.brec, &sc.brec_n, (size_t)Bmax * 9))) return rc;
        }
        if (any_value && (rc = ensure(ctx, &sc.its, &sc.its_n, (size_t)Bmax))) return rc;
        if (any_value && ctx->sc.integrator == INTEGRATOR_PRB && ctx->prb_wavefront &&
            (rc = ensure(ctx, &sc.prb_rec, &sc.prb_rec_n, (size_t)Bmax * 2 * (size_t)ctx->sc.prb.max_depth))) return rc;
        if (any_diff && ctx->sc.integrator == INTEGRATOR_PRB && ctx->prb_wavefront && ctx->defer_diff) {   // list of the deferred differentiable march
            if ((rc = ensure(ctx, &sc.bsegs, &sc.bsegs_n, (size_t)Bmax * 8))) return rc;
            if ((rc = ensure(ctx, &sc.b_its, &sc.b_its_n, (size_t)Bmax))) return rc;
        }
    }
    // Derived copies of the grid (row-expanded copies for the marchers, per-brick minima) are rebuilt when the grid was (re)bound
    // since they were built -- sdfrt_set_sdf is the `parameters_changed` of this boundary (python/shapes.py:477-479) and must follow
    // every change of the grid's values.  They are shared by all views in flight; everything per view lives in its ViewSlot.
    const bool dirty = ctx->grid_dirty;
    sdfrt_ctx::ViewSlot& V = VS(ctx);
    // empty-space test: brick minima -> dilation -> one DDA per film pixel and sweep -> masks consumed by k_gen_primary
    if (ctx->skip_empty && !ctx->sc.G.has_transform) {  // the brick DDA walks grid space = world space: identity to_world only
        const GridView& G = ctx->sc.G;
        const int npix = (S.W + 2 * S.border) * (S.H + 2 * S.border);
        const size_t mask_words = (size_t)npix / 4 + 1;
        bool bricks_ready = false;
        BrickView bricks;
        bricks.data = nullptr;
        bricks.nx = (G.rx + kBrick - 1) / kBrick; bricks.ny = (G.ry + kBrick - 1) / kBrick; bricks.nz = (G.rz + kBrick - 1) / kBrick;
        for (int k = 0; k < ns; ++k) {
            Sweep& sw = sweeps[k];
            if (sw.ps.n_items == 0) continue;
            // value pass: only the hit threshold matters.  DIFF pass: also the largest warp epsilon a lane can see
            // (strategy 6: edge_eps * warp_t <= edge_eps * maxt, python/warp.py:41-45), so that its warp weight is exactly 0.
            float skip_eps = 0.f;
            if (sw.diff && ctx->sc.W.kind == WARP_2D) skip_eps = ctx->sc.W.weight_strategy == 6 ? ctx->sc.W.edge_eps : -1.f;
            if (ctx->sc.T.step_scale != 1.f) skip_eps = -1.f;  // scaled SDF values: the bound does not apply as is
            const EmptySpaceParams e = empty_space_params(S, G, skip_eps);
            if (e.skip_eps < 0.f) continue;
            if (!bricks_ready) {
                const size_t nb = (size_t)bricks.nx * bricks.ny * bricks.nz;
                const bool fresh = ctx->brick_raw_n < nb;
                if ((rc = ensure(ctx, &ctx->brick_raw, &ctx->brick_raw_n, nb))) return rc;
                if ((rc = ensure(ctx, &V.brick_a, &V.brick_a_n, nb))) return rc;
                if ((rc = ensure(ctx, &V.brick_b, &V.brick_b_n, nb))) return rc;
                if ((rc = ensure(ctx, &V.pix_mask, &V.pix_mask_n, 2 * mask_words))) return rc;
                const unsigned gb = (unsigned)((nb + 255) / 256);
                LaunchScope ls(ctx, 0, st);
                if (dirty || fresh || !ctx->brick_raw_valid) {
                    k_brick_min<<<gb, 256, 0, st>>>(G, ctx->brick_raw, bricks.nx, bricks.ny, bricks.nz);
                    ctx->brick_raw_valid = true;
                    ctx->n_launches += 1;
                }
                k_brick_dilate<<<gb, 256, 0, st>>>(ctx->brick_raw, V.brick_a, bricks.nx, bricks.ny, bricks.nz, e.dilation, 0);
                k_brick_dilate<<<gb, 256, 0, st>>>(V.brick_a, V.brick_b, bricks.nx, bricks.ny, bricks.nz, e.dilation, 1);
                k_brick_dilate<<<gb, 256, 0, st>>>(V.brick_b, V.brick_a, bricks.nx, bricks.ny, bricks.nz, e.dilation, 2);
                ctx->n_launches += 2;
                bricks.data = V.brick_a;
                bricks_ready = true;
            }
            unsigned char* m = reinterpret_cast<unsigned char*>(V.pix_mask + (sw.diff ? mask_words : 0));
            { LaunchScope ls(ctx, 0, st); k_pixel_clear<<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(ctx->sc, bricks, e, m); }
            sw.mask = m;
        }
    }
    {
        const GridView& G = ctx->sc.G;
        const size_t nt = (size_t)G.rx * G.ry * G.rz;
        ctx->rows4_on = ctx->rows4_roles != 0 && (ctx->rows4_roles > 0 || nt <= kRows4MaxTexels);
        ctx->rows8_on = ctx->rows8_roles != 0 && nt <= kRows4MaxTexels;
        if (ctx->rows8_on && (ctx->rows8_roles & 15) == 15) ctx->rows4_on = false;   // every role on the 8x copy
        if (ctx->rows8_on) {
            const bool fresh = ctx->rows8_n < nt;
            if (fresh) {
                cudaFree(ctx->rows8); ctx->rows8 = nullptr; ctx->rows8_n = 0;
                SDFRT_CUDA(ctx, cudaMalloc(&ctx->rows8, nt * 8 * sizeof(float)));
                ctx->rows8_n = nt;
            }
            if (dirty || fresh || !ctx->rows8_valid) {
                LaunchScope ls(ctx, 0, st);
                k_expand_rows8<<<(unsigned)ctx->num_sms * 8, 256, 0, st>>>(G, ctx->rows8);
                ctx->rows8_valid = true;
            }
        }
        if (ctx->rows4_on) {
            const bool fresh = ctx->rows4_n < nt;
            if (fresh) {
                cudaFree(ctx->rows4); ctx->rows4 = nullptr; ctx->rows4_n = 0;
                SDFRT_CUDA(ctx, cudaMalloc(&ctx->rows4, nt * sizeof(float4)));
                ctx->rows4_n = nt;
            }
            if (dirty || fresh || !ctx->rows4_valid) {
                LaunchScope ls(ctx, 0, st);
                k_expand_rows<<<(unsigned)ctx->num_sms * 8, 256, 0, st>>>(G, ctx->rows4);
                ctx->rows4_valid = true;
            }
        }
    }
    // tensor map of the grid for the SMEM box staging (TMA): needs 16-byte row pitch and an identity to_world
    ctx->box_on = false;
    if (ctx->box_roles != 0 && !ctx->sc.G.has_transform && (ctx->sc.G.rx % 4) == 0 && ctx->sc.G.rx >= kBoxBase + 4 &&
        ctx->sc.G.ry >= kBoxBase + 3 && ctx->sc.G.rz >= kBoxBase + 3) {
        const GridView& G = ctx->sc.G;
        if (ctx->tmap_grid != G.data || ctx->tmap_dims[0] != G.rx || ctx->tmap_dims[1] != G.ry || ctx->tmap_dims[2] != G.rz) {
            if ((rc = encode_grid_tmap(ctx))) return rc;
        }
        ctx->box_on = true;
    }
    ctx->grid_dirty = false;
    // fork: the internal streams wait for everything queued on the caller's stream (film memsets, masks)
    if (nslots > 1) {
        SDFRT_CUDA(ctx, cudaEventRecord(ctx->ev_fork, st));
        for (int sl = 0; sl < nslots; ++sl) SDFRT_CUDA(ctx, cudaStreamWaitEvent(ctx->aux[sl], ctx->ev_fork, 0));
    }
    // batches of the sweeps, interleaved round-robin over the slots
    unsigned long long next[4] = {0, 0, 0, 0};
    unsigned long long bi = 0;
    for (bool more = true; more;) {
        more = false;
        for (int k = 0; k < ns; ++k) {
            Sweep& sw = sweeps[k];
            if (sw.B == 0 || next[k] >= sw.ps.n_items) continue;
            const int sl = (int)(bi % (unsigned long long)nslots);
            cudaStream_t bs = nslots > 1 ? ctx->aux[sl] : st;
            rc = sw.diff ? enqueue_batch<true>(ctx, sw, next[k], sl, bs) : enqueue_batch<false>(ctx, sw, next[k], sl, bs);
            if (rc) return rc;
            next[k] += sw.B;
            ++bi;
            more = true;
        }
    }
    // join: the caller's stream continues after all internal streams are done
    if (nslots > 1) {
        for (int sl = 0; sl < nslots; ++sl) {
            SDFRT_CUDA(ctx, cudaEventRecord(ctx->ev_join[sl], ctx->aux[sl]));
            SDFRT_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_join[sl], 0));
        }
    }
    // the deferred differentiable march: every flagged shadow / BSDF segment of the sweep in ONE persistent launch; the records go
    // straight to the lanes' rows (row 0 = its_t lands in the state row: non-zero for an unoccluded segment, as the adjoint expects)
    for (int k = 0; k < ns; ++k) {
        Sweep& sw = sweeps[k];
        if (sw.flag.segs == nullptr) continue;
        MarchExtra ex;
        ex.dest = sw.flag.dest;
        ex.n_slots_dev = sw.flag.count;
        const unsigned region = 256u;
        // The films of the sweep are complete at this point; what is still missing are warp records of the flagged segments, which only
        // adjoint sweep 2 reads.  The march is a few thousand long grazing rays -- latency, not throughput (2.3 ms for one rank of eight
        // against 3.3 ms for the whole film).  Optionally (SDFRT_DEFER_SIDE=1) it goes to a stream of its own so that the caller's stream
        // (film all-reduce, develop, the next view) does not queue behind it; sdfrt_adjoint_backprop and every reuse of the records then
        // wait for `defer_done`.  With two views in flight that bought nothing, so it is off by default.
        sdfrt_ctx::ViewSlot& V = VS(ctx);
        cudaStream_t ds = st;
        if (ctx->defer_side_stream) {
            if (!V.defer_stream) {
                SDFRT_CUDA(ctx, cudaStreamCreateWithFlags(&V.defer_stream, cudaStreamNonBlocking));
                SDFRT_CUDA(ctx, cudaEventCreateWithFlags(&V.defer_ready, cudaEventDisableTiming));
                SDFRT_CUDA(ctx, cudaEventCreateWithFlags(&V.defer_done, cudaEventDisableTiming));
            }
            ds = V.defer_stream;
            SDFRT_CUDA(ctx, cudaEventRecord(V.defer_ready, st));
            SDFRT_CUDA(ctx, cudaStreamWaitEvent(ds, V.defer_ready, 0));
        }
        { LaunchScope ls(ctx, 3, ds);
          launch_march_mode<true>(ctx, 3, 0xffffffffu, ds, sw.flag.segs, nullptr, sw.flag.capacity, region, 0, nullptr, V.rec, V.rec_stride,
                                  sw.flag.count + 1, ex); }
        SDFRT_CUDA(ctx, cudaGetLastError());
        if (ds != st) {
            SDFRT_CUDA(ctx, cudaEventRecord(V.defer_done, ds));
            V.defer_pending = true;
        }
    }
    return SDFRT_OK;
}
// make `st` wait for the deferred march of one view slot (or of all of them)
int wait_deferred(sdfrt_ctx* ctx, cudaStream_t st, int slot /* -1: every slot */) {
    for (int i = 0; i < kViewSlots; ++i) {
        if (slot >= 0 && i != slot) continue;
        sdfrt_ctx::ViewSlot& V = ctx->view[i];
        if (V.defer_pending) SDFRT_CUDA(ctx, cudaStreamWaitEvent(st, V.defer_done, 0));
    }
    return SDFRT_OK;
}


// sdf_prb_reparam forward sweep: primal render, or the re-render of render_backward that also keeps L per sample
int run_prb_sweep(sdfrt_ctx* ctx, const PassParams& ps, float* film, float* L_rec, unsigned long long rec_stride, cudaStream_t st) {
    SDFRT_CUDA(ctx, cudaSetDevice(ctx->device));
    SDFRT_CUDA(ctx, cudaMemsetAsync(film, 0, film_floats(ctx) * sizeof(float), st));
    if (ps.n_items == 0) return SDFRT_OK;
    if ((unsigned long long)kBandRows * (unsigned long long)ps.Wb * (unsigned long long)ps.spp > 0xffffffffull)
        return fail(ctx, SDFRT_ERR_INVALID, "film row band exceeds 2^32 samples");
    const unsigned long long blocks = (ps.n_items + kBlock - 1) / kBlock;
    if (blocks > 0x7fffffffull) return fail(ctx, SDFRT_ERR_INVALID, "too many samples for one launch");
    { LaunchScope ls(ctx, 1, st); k_prb_primal<<<(unsigned)blocks, kBlock, 0, st>>>(ctx->sc, ps, 0ull, ps.n_items, nullptr, PrbRecords{nullptr, 0, 9}, film, L_rec, rec_stride, ctx->counters); }
    SDFRT_CUDA(ctx, cudaGetLastError());
    return SDFRT_OK;
}

}  // namespace

// ----------------------------------------------------------------------------------------------------
// C ABI
// ----------------------------------------------------------------------------------------------------
extern "C" {

int sdfrt_version(void) { return 100; }

void sdfrt_tracer_defaults(sdfrt_tracer_params* p) { tracer_defaults(p); }
void sdfrt_warp_defaults(sdfrt_warp_params* p) { warp_defaults(p); }
void sdfrt_sensor_defaults(sdfrt_sensor* s) { sensor_defaults(s); }
void sdfrt_shading_defaults(sdfrt_shading* s) { shading_defaults(s); }
void sdfrt_prb_defaults(sdfrt_prb_params* p) { prb_defaults(p); }

sdfrt_ctx* sdfrt_create(int device, int* status) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || device < 0 || device >= n) {
        if (status) *status = SDFRT_ERR_CUDA;
        return nullptr;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) {
        if (status) *status = SDFRT_ERR_CUDA;  // built for sm_100a only; there is no fallback path
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) {
        if (status) *status = SDFRT_ERR_CUDA;
        return nullptr;
    }
    sdfrt_ctx* ctx = new sdfrt_ctx();
    ctx->device = device;
    std::memset(&ctx->sc, 0, sizeof(ctx->sc));
    tracer_defaults(&ctx->tracer_pub);
    warp_defaults(&ctx->warp_pub);
    sensor_defaults(&ctx->sensor_pub);
    shading_defaults(&ctx->shading_pub);
    ctx->sc.T = make_tracer(ctx->tracer_pub);
    set_grid_transform(ctx->sc.G, ctx->sc.T, nullptr);
    ctx->sc.W = make_warp(ctx->warp_pub);
    ctx->sc.S = make_sensor(ctx->sensor_pub);
    ctx->sc.Sh = make_shading(ctx->shading_pub);
    prb_defaults(&ctx->prb_pub);
    ctx->sc.prb = make_prb(ctx->prb_pub);
    ctx->sc.A.data = nullptr;
    ctx->sc.A.cr = ctx->shading_pub.albedo[0]; ctx->sc.A.cg = ctx->shading_pub.albedo[1]; ctx->sc.A.cb = ctx->shading_pub.albedo[2];
    ctx->sc.integrator = INTEGRATOR_DIRECT;
    if (cudaMalloc((void**)&ctx->counters, sizeof(Counters)) != cudaSuccess || cudaMemset(ctx->counters, 0, sizeof(Counters)) != cudaSuccess) {
        delete ctx;
        if (status) *status = SDFRT_ERR_CUDA;
        return nullptr;
    }
    ctx->num_sms = prop.multiProcessorCount;
    bool ok = cudaMalloc((void**)&ctx->work, kSlots * WQ_COUNT * sizeof(unsigned)) == cudaSuccess;
    if (const char* e = std::getenv("SDFRT_ROWS4")) ctx->rows4_roles = std::atoi(e) & 15;   // developer override: 0 = off, else role mask
    std::memset(&ctx->tmap, 0, sizeof(ctx->tmap));
    if (const char* e = std::getenv("SDFRT_BOX")) ctx->box_roles = std::atoi(e) & 10;      // developer override: SMEM box staging, roles 1 / 3
    {
        const int smem = (int)sizeof(BoxShared) + 128;
        ok = ok && cudaFuncSetAttribute(k_march<false, 0, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(k_march<false, 1, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(k_march<true, 0, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(k_march<true, 1, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(k_march<false, 0, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(k_march<false, 1, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(k_march<true, 0, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(k_march<true, 1, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess;
    }
    if (const char* e = std::getenv("SDFRT_ROWS8")) ctx->rows8_roles = std::atoi(e) & 15;  // developer override: role mask of the 8x copy
    if (const char* e = std::getenv("SDFRT_SLOTS")) ctx->max_slots = std::max(1, std::min(kSlots, std::atoi(e)));          // developer overrides
    if (const char* e = std::getenv("SDFRT_MIN_LIVE")) ctx->min_live = std::max(1, std::min(32, std::atoi(e)));
    if (const char* e = std::getenv("SDFRT_PRB_OCC")) ctx->prb_min_blocks = std::atoi(e);
    if (const char* e = std::getenv("SDFRT_SITEQ")) ctx->site_queue = std::atoi(e) != 0;
    if (const char* e = std::getenv("SDFRT_DEFER_SIDE")) ctx->defer_side_stream = std::atoi(e) != 0;
    if (const char* e = std::getenv("SDFRT_SHADE_ROWS8")) ctx->shade_rows8 = std::atoi(e) != 0;
    if (const char* e = std::getenv("SDFRT_SITEQ_CAP")) ctx->site_queue_cap = std::max(1ll, std::atoll(e));
    if (const char* e = std::getenv("SDFRT_PRB_WAVEFRONT")) ctx->prb_wavefront = std::atoi(e) != 0;
    if (const char* e = std::getenv("SDFRT_PRB_PRIMARY")) ctx->prb_primary_wavefront = std::atoi(e) != 0;
    if (const char* e = std::getenv("SDFRT_DEFER")) ctx->defer_diff = std::atoi(e) != 0;
    if (const char* e = std::getenv("SDFRT_SKIP_EMPTY")) ctx->skip_empty = std::atoi(e) != 0;
    if (const char* e = std::getenv("SDFRT_BATCH_LOG2")) ctx->batch_target = 1ull << std::max(16, std::min(28, std::atoi(e)));
    for (int i = 0; i < kSlots; ++i) {
        ok = ok && cudaStreamCreateWithFlags(&ctx->aux[i], cudaStreamNonBlocking) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming) == cudaSuccess;
    }
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&ctx->host_stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->march_blocks[0], k_march<false, 0, 0>, kMarchBlock, 0) == cudaSuccess;
    ok = ok && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->march_blocks[1], k_march<false, 1, 0>, kMarchBlock, 0) == cudaSuccess;
    ok = ok && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->march_blocks[2], k_march<false, 2, 0>, kMarchBlock, 0) == cudaSuccess;
    ok = ok && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->march_blocks[3], k_march<true, 0, 0>, kMarchBlock, 0) == cudaSuccess;
    ok = ok && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->march_blocks[4], k_march<true, 1, 0>, kMarchBlock, 0) == cudaSuccess;
    ok = ok && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->march_blocks[5], k_march<true, 2, 0>, kMarchBlock, 0) == cudaSuccess;
    for (int i = 0; i < 6; ++i) ok = ok && ctx->march_blocks[i] >= 1;
    if (const char* e = std::getenv("SDFRT_CACHE_MODES")) {  // developer override, e.g. "1102"
        for (int i = 0; i < 4 && e[i]; ++i) ctx->cache_mode[i] = (e[i] - '0') % 3;
    }
    if (!ok) {
        sdfrt_destroy(ctx);
        if (status) *status = SDFRT_ERR_CUDA;
        return nullptr;
    }
    if (status) *status = SDFRT_OK;
    return ctx;
}

void sdfrt_destroy(sdfrt_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (auto& t : ctx->timed) { cudaEventDestroy(t.second.first); cudaEventDestroy(t.second.second); }
    for (auto e : ctx->event_pool) cudaEventDestroy(e);
    for (auto& v : ctx->view) { if (v.defer_stream) cudaStreamDestroy(v.defer_stream); if (v.defer_ready) cudaEventDestroy(v.defer_ready); if (v.defer_done) cudaEventDestroy(v.defer_done); }
    for (auto& v : ctx->view) { cudaFree(v.rec); cudaFree(v.film_adj); cudaFree(v.prb_rec); cudaFree(v.brick_a); cudaFree(v.brick_b); cudaFree(v.pix_mask); cudaFree(v.flag_segs); cudaFree(v.flag_dest); cudaFree(v.flag_count); cudaFree(v.siteq_a); cudaFree(v.siteq_b); cudaFree(v.siteq_count); }
    cudaFree(ctx->counters);
    cudaFree(ctx->work);
    cudaFree(ctx->rows4); cudaFree(ctx->rows8); cudaFree(ctx->brick_raw); cudaFree(ctx->env_buf);
    cudaFree(ctx->rd_work); cudaFree(ctx->rd_state); cudaFree(ctx->lap_r);
    for (auto& sc : ctx->scratch) {
        cudaFree(sc.prb_rec); cudaFree(sc.segs); cudaFree(sc.its); cudaFree(sc.ssegs); cudaFree(sc.sinfo); cudaFree(sc.s_its); cudaFree(sc.srec); cudaFree(sc.scount); cudaFree(sc.sbox);
        cudaFree(sc.bsegs); cudaFree(sc.binfo); cudaFree(sc.b_its); cudaFree(sc.brec); cudaFree(sc.bcount);
    }
    for (int i = 0; i < kSlots; ++i) {
        if (ctx->aux[i]) cudaStreamDestroy(ctx->aux[i]);
        if (ctx->ev_join[i]) cudaEventDestroy(ctx->ev_join[i]);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->host_stream) cudaStreamDestroy(ctx->host_stream);
    if (ctx->rd_err_host) cudaFreeHost(ctx->rd_err_host);
    if (ctx->rd_err_event) cudaEventDestroy(ctx->rd_err_event);
    cudaFree(ctx->h_alb); cudaFree(ctx->h_galb);
    cudaFree(ctx->h_sdf); cudaFree(ctx->h_film); cudaFree(ctx->h_film2); cudaFree(ctx->h_image); cudaFree(ctx->h_gimg); cudaFree(ctx->h_grad);
    delete ctx;
}

const char* sdfrt_last_error(const sdfrt_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int sdfrt_set_sdf(sdfrt_ctx* ctx, const float* sdf, int rx, int ry, int rz, const float p[3]) {
    if (!ctx) return SDFRT_ERR_INVALID;
    if (!sdf || rx < 4 || ry < 4 || rz < 4) return fail(ctx, SDFRT_ERR_INVALID, "bad SDF grid (need >= 4 texels per axis)");
    ctx->sc.G.data = sdf;
    ctx->sc.G.rows4 = nullptr;
    ctx->sc.G.rows8 = nullptr;
    ctx->sc.G.rx = rx; ctx->sc.G.ry = ry; ctx->sc.G.rz = rz;
    ctx->sc.G.px = p ? p[0] : 0.f; ctx->sc.G.py = p ? p[1] : 0.f; ctx->sc.G.pz = p ? p[2] : 0.f;
    ctx->have_sdf = true;
    ctx->grid_dirty = true;
    ctx->rows4_valid = ctx->rows8_valid = ctx->brick_raw_valid = false;
    VS(ctx).adj_ready = false;
    return SDFRT_OK;
}

int sdfrt_set_albedo(sdfrt_ctx* ctx, const float* rgb_grid, int rx, int ry, int rz) {
    if (!ctx) return SDFRT_ERR_INVALID;
    if (rgb_grid && (rx < 1 || ry < 1 || rz < 1)) return fail(ctx, SDFRT_ERR_INVALID, "bad albedo grid");
    ctx->sc.A.data = rgb_grid;
    ctx->sc.A.rx = rx; ctx->sc.A.ry = ry; ctx->sc.A.rz = rz;
    VS(ctx).adj_ready = false;
    return SDFRT_OK;
}

int sdfrt_set_tracer(sdfrt_ctx* ctx, const sdfrt_tracer_params* p) {
    if (!ctx || !p) return SDFRT_ERR_INVALID;
    if (p->weight_power < 1 || p->weight_power > 8) return fail(ctx, SDFRT_ERR_INVALID, "weight_power out of range");
    ctx->tracer_pub = *p;
    ctx->sc.T = make_tracer(*p);
    set_grid_transform(ctx->sc.G, ctx->sc.T, ctx->has_xf ? ctx->to_world : nullptr);  // the bbox travels with the tracer block
    VS(ctx).adj_ready = false;
    return SDFRT_OK;
}

int sdfrt_set_warp(sdfrt_ctx* ctx, const sdfrt_warp_params* p) {
    if (!ctx || !p) return SDFRT_ERR_INVALID;
    if (p->kind < SDFRT_WARP_NONE || p->kind > SDFRT_WARP_CONV) return fail(ctx, SDFRT_ERR_INVALID, "unknown warp kind");
    if (p->kind == SDFRT_WARP_CONV && (p->n_aux_rays < 1 || p->n_aux_rays > kMaxAuxRays || !(p->kappa > 0.f) || !(p->power > 0.f)))
        return fail(ctx, SDFRT_ERR_INVALID, "bad convolution warp parameters (need 1 <= n_aux_rays <= 64, kappa > 0, power > 0)");
    ctx->warp_pub = *p;
    ctx->sc.W = make_warp(*p);
    VS(ctx).adj_ready = false;
    return SDFRT_OK;
}

int sdfrt_set_sensor(sdfrt_ctx* ctx, const sdfrt_sensor* s) {
    if (!ctx || !s) return SDFRT_ERR_INVALID;
    if (s->width < 1 || s->height < 1 || !(s->filter_radius > 0.f) || !(s->filter_stddev > 0.f) || s->filter_radius > 2.5f)
        return fail(ctx, SDFRT_ERR_INVALID, "bad sensor");
    ctx->sensor_pub = *s;
    ctx->sc.S = make_sensor(*s);
    VS(ctx).adj_ready = false;
    return SDFRT_OK;
}

int sdfrt_set_shading(sdfrt_ctx* ctx, const sdfrt_shading* s) {
    if (!ctx || !s) return SDFRT_ERR_INVALID;
    ctx->shading_pub = *s;
    ctx->sc.Sh = make_shading(*s);
    ctx->sc.Sh.shadow_origin = ctx->shadow_origin;
    ctx->sc.Sh.use_mis = ctx->use_mis;
    ctx->sc.A.cr = s->albedo[0]; ctx->sc.A.cg = s->albedo[1]; ctx->sc.A.cb = s->albedo[2];
    VS(ctx).adj_ready = false;
    return SDFRT_OK;
}

int sdfrt_set_integrator(sdfrt_ctx* ctx, int integrator) {
    if (!ctx) return SDFRT_ERR_INVALID;
    if (integrator < SDFRT_INTEGRATOR_SILHOUETTE || integrator > SDFRT_INTEGRATOR_PRB)
        return fail(ctx, SDFRT_ERR_INVALID, "unknown integrator");
    ctx->sc.integrator = integrator;
    VS(ctx).adj_ready = false;
    return SDFRT_OK;
}

int sdfrt_set_sdf_transform(sdfrt_ctx* ctx, const float* to_world) {
    if (!ctx) return SDFRT_ERR_INVALID;
    if (to_world) {
        for (int i = 0; i < 16; ++i)
            if (!std::isfinite(to_world[i])) return fail(ctx, SDFRT_ERR_INVALID, "sdf_to_world has non-finite entries");
        if (to_world[12] != 0.f || to_world[13] != 0.f || to_world[14] != 0.f || to_world[15] != 1.f)
            return fail(ctx, SDFRT_ERR_INVALID, "sdf_to_world must be affine (last row 0 0 0 1)");
    }
    GridView G = ctx->sc.G;
    TracerParams T = ctx->sc.T;
    if (!set_grid_transform(G, T, to_world)) return fail(ctx, SDFRT_ERR_INVALID, "sdf_to_world is singular");
    ctx->sc.G = G;
    ctx->sc.T = T;
    ctx->has_xf = to_world != nullptr;
    if (to_world) std::memcpy(ctx->to_world, to_world, sizeof(ctx->to_world));
    VS(ctx).adj_ready = false;
    return SDFRT_OK;
}

int sdfrt_set_shadow_origin(sdfrt_ctx* ctx, int mode) {
    if (!ctx) return SDFRT_ERR_INVALID;
    if (mode < 0 || mode > 2) return fail(ctx, SDFRT_ERR_INVALID, "shadow origin mode must be 0 (si), 1 (detach_indirect_si) or 2 (decouple_reparam)");
    ctx->shadow_origin = mode;
    ctx->sc.Sh.shadow_origin = mode;
    VS(ctx).adj_ready = false;
    return SDFRT_OK;
}

int sdfrt_set_use_mis(sdfrt_ctx* ctx, int enable) {
    if (!ctx) return SDFRT_ERR_INVALID;
    ctx->use_mis = enable ? 1 : 0;
    ctx->sc.Sh.use_mis = ctx->use_mis;
    VS(ctx).adj_ready = false;
    return SDFRT_OK;
}

int sdfrt_set_envmap_host(sdfrt_ctx* ctx, const float* rgb_host, int width, int height, float scale) {
    if (!ctx) return SDFRT_ERR_INVALID;
    SDFRT_CUDA(ctx, cudaSetDevice(ctx->device));
    SDFRT_CUDA(ctx, cudaDeviceSynchronize());   // views in flight may still read the previous tables
    cudaFree(ctx->env_buf);
    ctx->env_buf = nullptr;
    std::memset(&ctx->sc.E, 0, sizeof(ctx->sc.E));
    ctx->sc.E.has_rot = ctx->env_has_rot;
    for (int i = 0; i < 9; ++i) ctx->sc.E.R[i] = ctx->env_R[i];
    for (auto& v : ctx->view) v.adj_ready = false;
    if (rgb_host == nullptr) return SDFRT_OK;
    EnvTables t;
    if (!build_envmap(rgb_host, width, height, scale, t))
        return fail(ctx, SDFRT_ERR_INVALID, "bad envmap (need width >= 1, height >= 2, non-negative finite radiance, not all black)");
    const size_t n_data = t.data.size(), n_m = t.marg_cdf.size(), n_c = t.cond_cdf.size();
    const size_t total = n_data + n_m + 2 * n_c;
    SDFRT_CUDA(ctx, cudaMalloc((void**)&ctx->env_buf, total * sizeof(float)));
    float* d = ctx->env_buf;
    SDFRT_CUDA(ctx, cudaMemcpy(d, t.data.data(), n_data * sizeof(float), cudaMemcpyHostToDevice));
    SDFRT_CUDA(ctx, cudaMemcpy(d + n_data, t.marg_cdf.data(), n_m * sizeof(float), cudaMemcpyHostToDevice));
    SDFRT_CUDA(ctx, cudaMemcpy(d + n_data + n_m, t.cond_cdf.data(), n_c * sizeof(float), cudaMemcpyHostToDevice));
    SDFRT_CUDA(ctx, cudaMemcpy(d + n_data + n_m + n_c, t.cell_prob.data(), n_c * sizeof(float), cudaMemcpyHostToDevice));
    ctx->sc.E.data = d; ctx->sc.E.marg_cdf = d + n_data; ctx->sc.E.cond_cdf = d + n_data + n_m; ctx->sc.E.cell_prob = d + n_data + n_m + n_c;
    ctx->sc.E.W1 = t.W1; ctx->sc.E.H = t.H;
    return SDFRT_OK;
}

int sdfrt_set_envmap_transform(sdfrt_ctx* ctx, const float* to_world) {
    if (!ctx) return SDFRT_ERR_INVALID;
    float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    int has = 0;
    if (to_world != nullptr) {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) R[i * 3 + j] = to_world[i * 4 + j];
        double err = 0.0;   // a rotation: R R^T = I, det = +1 (the emitter is at infinity: translation is ignored, scale is not allowed)
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double v = 0.0;
                for (int k = 0; k < 3; ++k) v += (double)R[i * 3 + k] * (double)R[j * 3 + k];
                err = std::max(err, std::fabs(v - (i == j ? 1.0 : 0.0)));
            }
        const double det = (double)R[0] * ((double)R[4] * R[8] - (double)R[5] * R[7]) - (double)R[1] * ((double)R[3] * R[8] - (double)R[5] * R[6]) +
                           (double)R[2] * ((double)R[3] * R[7] - (double)R[4] * R[6]);
        if (err > 1e-4 || det < 0.0) return fail(ctx, SDFRT_ERR_INVALID, "the envmap's to_world must be a rotation");
        has = !(R[0] == 1.f && R[4] == 1.f && R[8] == 1.f);
    }
    ctx->env_has_rot = has;
    for (int i = 0; i < 9; ++i) { ctx->env_R[i] = R[i]; ctx->sc.E.R[i] = R[i]; }
    ctx->sc.E.has_rot = has;
    for (auto& v : ctx->view) v.adj_ready = false;
    return SDFRT_OK;
}

int sdfrt_set_antithetic(sdfrt_ctx* ctx, int enable) {
    if (!ctx) return SDFRT_ERR_INVALID;
    ctx->antithetic = enable ? 1 : 0;
    VS(ctx).adj_ready = false;
    return SDFRT_OK;
}

int sdfrt_set_prb(sdfrt_ctx* ctx, const sdfrt_prb_params* p) {
    if (!ctx || !p) return SDFRT_ERR_INVALID;
    if (p->max_depth < 1 || p->max_depth > 1024 || p->rr_depth < 0) return fail(ctx, SDFRT_ERR_INVALID, "bad PRB parameters (need 1 <= max_depth <= 1024, rr_depth >= 0)");
    ctx->prb_pub = *p;
    ctx->sc.prb = make_prb(*p);
    VS(ctx).adj_ready = false;
    return SDFRT_OK;
}

int sdfrt_set_partition_rows(sdfrt_ctx* ctx, int rank, int world, int rows) {
    if (!ctx) return SDFRT_ERR_INVALID;
    if (world < 1 || rank < 0 || rank >= world) return fail(ctx, SDFRT_ERR_INVALID, "bad partition");
    if (rows != 1 && rows != 2 && rows != 4 && rows != 8) return fail(ctx, SDFRT_ERR_INVALID, "partition granularity must be 1, 2, 4 or 8 film rows");
    ctx->rank = rank;
    ctx->world = world;
    ctx->assign_rows = rows;
    VS(ctx).adj_ready = false;
    return SDFRT_OK;
}
int sdfrt_set_partition(sdfrt_ctx* ctx, int rank, int world) { return sdfrt_set_partition_rows(ctx, rank, world, kBandRows); }

int sdfrt_set_lane_window(sdfrt_ctx* ctx, uint64_t lane_lo, uint64_t lane_hi) {
    if (!ctx) return SDFRT_ERR_INVALID;
    if (lane_lo == 0 && lane_hi == 0) { ctx->win_lo = 0ull; ctx->win_hi = ~0ull; return SDFRT_OK; }
    if (lane_hi <= lane_lo) return fail(ctx, SDFRT_ERR_INVALID, "empty lane window");
    ctx->win_lo = lane_lo;
    ctx->win_hi = lane_hi;
    return SDFRT_OK;
}

int sdfrt_set_empty_space_skip(sdfrt_ctx* ctx, int enable) {
    if (!ctx) return SDFRT_ERR_INVALID;
    ctx->skip_empty = enable != 0;
    return SDFRT_OK;
}

uint64_t sdfrt_adjoint_token(const sdfrt_ctx* ctx) { return (ctx && VS(ctx).adj_ready) ? VS(ctx).generation : 0ull; }

int sdfrt_set_view_slot(sdfrt_ctx* ctx, int slot) {
    if (!ctx) return SDFRT_ERR_INVALID;
    if (slot < 0 || slot >= kViewSlots) return fail(ctx, SDFRT_ERR_INVALID, "view slot out of range");
    ctx->cur_view = slot;
    return SDFRT_OK;
}
int sdfrt_num_view_slots(void) { return kViewSlots; }

int sdfrt_film_dims(const sdfrt_ctx* ctx, int* height_b, int* width_b, int* border) {
    if (!ctx) return SDFRT_ERR_INVALID;
    const SensorParams& S = ctx->sc.S;
    if (height_b) *height_b = S.H + 2 * S.border;
    if (width_b) *width_b = S.W + 2 * S.border;
    if (border) *border = S.border;
    return SDFRT_OK;
}

uint64_t sdfrt_wavefront_size(const sdfrt_ctx* ctx, int spp) {
    if (!ctx || spp <= 0) return 0;
    const SensorParams& S = ctx->sc.S;
    return (uint64_t)(S.W + 2 * S.border) * (uint64_t)(S.H + 2 * S.border) * (uint64_t)spp;
}

int sdfrt_render_primal(sdfrt_ctx* ctx, uint32_t seed, int spp, float* film_accum, void* stream) {
    int rc = check_ready(ctx, spp);
    if (rc) return rc;
    if (!film_accum) return fail(ctx, SDFRT_ERR_INVALID, "film_accum is null");
    const bool prb = ctx->sc.integrator == INTEGRATOR_PRB;
    if (prb && !ctx->prb_primary_wavefront && !ctx->prb_wavefront)
        return run_prb_sweep(ctx, make_pass(ctx->sc.S, seed, spp, ctx->rank, ctx->world, 0, ctx->assign_rows), film_accum, nullptr, 0, (cudaStream_t)stream);
    Sweep sw;
    sw.diff = false;
    sw.ps = make_pass(ctx->sc.S, seed, spp, ctx->rank, ctx->world, prb ? 0 : ctx->antithetic, ctx->assign_rows);
    sw.film = film_accum;
    return run_sweeps(ctx, &sw, 1, (cudaStream_t)stream);
}

int sdfrt_develop(sdfrt_ctx* ctx, const float* film_accum, float* image, void* stream) {
    if (!ctx || !film_accum || !image) return SDFRT_ERR_INVALID;
    SDFRT_CUDA(ctx, cudaSetDevice(ctx->device));
    const SensorParams& S = ctx->sc.S;
    const int n = S.W * S.H;
    { LaunchScope ls(ctx, 7, (cudaStream_t)stream); k_develop<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(film_accum, image, S.W, S.H, S.border); }
    SDFRT_CUDA(ctx, cudaGetLastError());
    return SDFRT_OK;
}

int sdfrt_adjoint_accumulate(sdfrt_ctx* ctx, uint32_t seed_grad, int spp_grad, float* film_accum, void* stream) {
    int rc = check_ready(ctx, spp_grad);
    if (rc) return rc;
    if (!film_accum) return fail(ctx, SDFRT_ERR_INVALID, "film_accum is null");
    SDFRT_CUDA(ctx, cudaSetDevice(ctx->device));
    Sweep sw;
    sw.diff = true;
    sw.ps = make_pass(ctx->sc.S, seed_grad, spp_grad, ctx->rank, ctx->world, ctx->sc.integrator == INTEGRATOR_PRB ? 0 : ctx->antithetic, ctx->assign_rows);
    sw.film = film_accum;
    rc = ensure(ctx, &VS(ctx).rec, &VS(ctx).rec_capacity, (size_t)sw.ps.n_items * (size_t)rec_rows(ctx) + 1);
    if (rc) return rc;
    VS(ctx).rec_stride = sw.ps.n_items;
    if (ctx->sc.integrator == INTEGRATOR_PRB && ctx->prb_wavefront &&
        (rc = ensure(ctx, &VS(ctx).prb_rec, &VS(ctx).prb_rec_n, (size_t)sw.ps.n_items * kPrbRecFloats * (size_t)ctx->sc.prb.max_depth + 1))) return rc;
    if (ctx->sc.integrator == INTEGRATOR_PRB && !ctx->prb_primary_wavefront && !ctx->prb_wavefront)  // re-render of render_backward (:204-206), keeps L per sample
        rc = run_prb_sweep(ctx, sw.ps, film_accum, VS(ctx).rec + (size_t)REC_S_STATE * VS(ctx).rec_stride, VS(ctx).rec_stride, (cudaStream_t)stream);
    else
        rc = run_sweeps(ctx, &sw, 1, (cudaStream_t)stream);
    if (rc) return rc;
    VS(ctx).adj_pass = sw.ps;
    VS(ctx).sc = ctx->sc;
    VS(ctx).adj_ready = true;
    VS(ctx).generation = ++ctx->adj_generation;
    return SDFRT_OK;
}

int sdfrt_render_view(sdfrt_ctx* ctx, uint32_t seed, int spp, float* film_primal, uint32_t seed_grad, int spp_grad,
                      float* film_adjoint, void* stream) {
    int rc = check_ready(ctx, spp);
    if (rc) return rc;
    if ((rc = check_ready(ctx, spp_grad))) return rc;
    if (!film_primal || !film_adjoint) return fail(ctx, SDFRT_ERR_INVALID, "film buffer is null");
    SDFRT_CUDA(ctx, cudaSetDevice(ctx->device));
    Sweep sw[2];
    sw[0].diff = false;
    const int anti = ctx->sc.integrator == INTEGRATOR_PRB ? 0 : ctx->antithetic;  // sdf_prb_reparam.sample_rays has no antithetic mode
    sw[0].ps = make_pass(ctx->sc.S, seed, spp, ctx->rank, ctx->world, anti, ctx->assign_rows);
    sw[0].film = film_primal;
    sw[1].diff = true;
    sw[1].ps = make_pass(ctx->sc.S, seed_grad, spp_grad, ctx->rank, ctx->world, anti, ctx->assign_rows);
    sw[1].film = film_adjoint;
    rc = ensure(ctx, &VS(ctx).rec, &VS(ctx).rec_capacity, (size_t)sw[1].ps.n_items * (size_t)rec_rows(ctx) + 1);
    if (rc) return rc;
    VS(ctx).rec_stride = sw[1].ps.n_items;
    if (ctx->sc.integrator == INTEGRATOR_PRB && ctx->prb_wavefront &&
        (rc = ensure(ctx, &VS(ctx).prb_rec, &VS(ctx).prb_rec_n, (size_t)sw[1].ps.n_items * kPrbRecFloats * (size_t)ctx->sc.prb.max_depth + 1))) return rc;
    if (ctx->sc.integrator == INTEGRATOR_PRB && !ctx->prb_primary_wavefront && !ctx->prb_wavefront) {
        rc = run_prb_sweep(ctx, sw[0].ps, film_primal, nullptr, 0, (cudaStream_t)stream);
        if (!rc) rc = run_prb_sweep(ctx, sw[1].ps, film_adjoint, VS(ctx).rec + (size_t)REC_S_STATE * VS(ctx).rec_stride, VS(ctx).rec_stride, (cudaStream_t)stream);
    } else {
        rc = run_sweeps(ctx, sw, 2, (cudaStream_t)stream);
    }
    if (rc) return rc;
    VS(ctx).adj_pass = sw[1].ps;
    VS(ctx).sc = ctx->sc;
    VS(ctx).adj_ready = true;
    VS(ctx).generation = ++ctx->adj_generation;
    return SDFRT_OK;
}

int sdfrt_adjoint_backprop(sdfrt_ctx* ctx, const float* film_accum, const float* grad_image, float* grad_sdf,
                           float* grad_albedo, float* grad_p, void* stream) {
    if (!ctx) return SDFRT_ERR_INVALID;
    if (!VS(ctx).adj_ready) return fail(ctx, SDFRT_ERR_STATE, "sdfrt_adjoint_accumulate must run first (and no set_* in between)");
    if (!film_accum || !grad_image) return fail(ctx, SDFRT_ERR_INVALID, "null buffer");
    if (!grad_sdf && !grad_albedo && !grad_p) return fail(ctx, SDFRT_ERR_INVALID, "no gradient target");
    SDFRT_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    { const int rcw = wait_deferred(ctx, st, ctx->cur_view); if (rcw) return rcw; }   // the warp records of the flagged segments
    const SceneParams& vsc = VS(ctx).sc;   // the scene adjoint sweep 1 of this view was rendered with
    const SensorParams& S = vsc.S;
    int rc = ensure(ctx, &VS(ctx).film_adj, &VS(ctx).film_adj_capacity, (size_t)(S.W + 2 * S.border) * (size_t)(S.H + 2 * S.border) * 4);
    if (rc) return rc;
    const int npix = (S.W + 2 * S.border) * (S.H + 2 * S.border);
    { LaunchScope ls(ctx, 5, st); k_film_adjoint<<<(npix + 255) / 256, 256, 0, st>>>(film_accum, grad_image, VS(ctx).film_adj, S.W, S.H, S.border); }
    SDFRT_CUDA(ctx, cudaGetLastError());
    PassParams ps = VS(ctx).adj_pass;
    ps.win_lo = ctx->win_lo;
    ps.win_hi = ctx->win_hi;
    const unsigned long long n = ps.n_items;
    if (n == 0) return SDFRT_OK;
    BackwardTargets tg;
    tg.grad_sdf = grad_sdf;
    tg.grad_albedo = grad_albedo;
    tg.queue = SiteQueue{nullptr, nullptr, nullptr, 0u};
    const unsigned long long blocks = (n + kBlock - 1) / kBlock;
    auto make_queue = [&](SiteQueue& q) -> int {   // per-view site queue (views in flight run their adjoints concurrently)
        const size_t cap = (size_t)std::min<unsigned long long>(n, ctx->site_queue_cap);
        int r;
        if ((r = ensure(ctx, &VS(ctx).siteq_a, &VS(ctx).siteq_a_n, cap * 4))) return r;
        if ((r = ensure(ctx, &VS(ctx).siteq_b, &VS(ctx).siteq_b_n, cap * 4))) return r;
        if (!VS(ctx).siteq_count) SDFRT_CUDA(ctx, cudaMalloc((void**)&VS(ctx).siteq_count, sizeof(unsigned)));
        SDFRT_CUDA(ctx, cudaMemsetAsync(VS(ctx).siteq_count, 0, sizeof(unsigned), st));
        q = SiteQueue{reinterpret_cast<float4*>(VS(ctx).siteq_a), reinterpret_cast<float4*>(VS(ctx).siteq_b), VS(ctx).siteq_count, (unsigned)cap};
        return SDFRT_OK;
    };
    if (vsc.integrator == INTEGRATOR_PRB) {
        LaunchScope ls(ctx, 6, st);
        const float* L_rec = VS(ctx).rec + (size_t)REC_S_STATE * VS(ctx).rec_stride;   // rows 9..11: L of the re-render
        const float* prim = (ctx->prb_primary_wavefront && !ctx->prb_wavefront) ? VS(ctx).rec : nullptr;  // rows 0..8: camera ray's march by k_march
        const PrbRecords recs = ctx->prb_wavefront ? PrbRecords{VS(ctx).prb_rec, ps.n_items, 9} : PrbRecords{nullptr, 0, 9};
        SiteQueue queue{nullptr, nullptr, nullptr, 0u};
        if (ctx->site_queue && tg.grad_sdf != nullptr && (rc = make_queue(queue))) return rc;   // sites go through a queue and the grouped scatter (k_scatter_sites)
#define SDFRT_PRB_BWD_ARGS vsc, ps, VS(ctx).film_adj, L_rec, prim, recs, VS(ctx).rec_stride, tg, grad_p, ctx->counters, queue
        if (queue.a != nullptr) {
            if (ctx->prb_min_blocks >= 3) k_prb_backward<3, true><<<(unsigned)blocks, kBlock, 0, st>>>(SDFRT_PRB_BWD_ARGS);
            else k_prb_backward<2, true><<<(unsigned)blocks, kBlock, 0, st>>>(SDFRT_PRB_BWD_ARGS);
            k_scatter_sites<<<(unsigned)ctx->num_sms * 8u, kBlock, 0, st>>>(vsc.G, queue, tg.grad_sdf);
        } else {
            if (ctx->prb_min_blocks >= 3) k_prb_backward<3, false><<<(unsigned)blocks, kBlock, 0, st>>>(SDFRT_PRB_BWD_ARGS);
            else k_prb_backward<2, false><<<(unsigned)blocks, kBlock, 0, st>>>(SDFRT_PRB_BWD_ARGS);
        }
#undef SDFRT_PRB_BWD_ARGS
    } else {
        LaunchScope ls(ctx, 6, st);
        if (vsc.W.kind == WARP_CONV) {   // WarpFieldConvolution: the auxiliary rays are marched inside the adjoint (python/warp_conv.py:185-282)
            if (ctx->site_queue && tg.grad_sdf != nullptr && (rc = make_queue(tg.queue))) return rc;   // one site per auxiliary hit: queued like the replay's
            k_backward<true><<<(unsigned)blocks, kBlock, 0, st>>>(vsc, ps, VS(ctx).film_adj, VS(ctx).rec, VS(ctx).rec_stride, tg, grad_p, ctx->counters);
            if (tg.queue.a != nullptr) k_scatter_sites<<<(unsigned)ctx->num_sms * 8u, kBlock, 0, st>>>(vsc.G, tg.queue, tg.grad_sdf);
        } else
            k_backward<false><<<(unsigned)blocks, kBlock, 0, st>>>(vsc, ps, VS(ctx).film_adj, VS(ctx).rec, VS(ctx).rec_stride, tg, grad_p, ctx->counters);
    }
    SDFRT_CUDA(ctx, cudaGetLastError());
    return SDFRT_OK;
}

int sdfrt_get_stats(sdfrt_ctx* ctx, sdfrt_stats* out, int reset, void* stream) {
    if (!ctx || !out) return SDFRT_ERR_INVALID;
    SDFRT_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    { const int rcw = wait_deferred(ctx, st, -1); if (rcw) return rcw; }
    Counters c;
    SDFRT_CUDA(ctx, cudaMemcpyAsync(&c, ctx->counters, sizeof(c), cudaMemcpyDeviceToHost, st));
    SDFRT_CUDA(ctx, cudaStreamSynchronize(st));
    out->n_rays = c.n_rays; out->n_evals_value = c.n_evals_value; out->n_evals_grad = c.n_evals_grad;
    out->n_evals_hess = c.n_evals_hess; out->n_scatter_sites = c.n_scatter_sites; out->n_albedo_lookups = c.n_albedo_lookups;
    out->n_steps_max = c.n_steps_max; out->n_hits = c.n_hits;
    out->n_evals_adjoint = c.n_evals_adjoint;
    out->n_tap_loads = c.n_tap_loads;
    out->n_skipped = c.n_skipped;
    out->n_box_loads = c.n_box_loads;
    out->n_launches = ctx->n_launches;
    if (reset) {
        SDFRT_CUDA(ctx, cudaMemsetAsync(ctx->counters, 0, sizeof(Counters), st));
        ctx->n_launches = 0;
    }
    return SDFRT_OK;
}

int sdfrt_set_profiling(sdfrt_ctx* ctx, int enable) {
    if (!ctx) return SDFRT_ERR_INVALID;
    ctx->profiling = enable != 0;
    return SDFRT_OK;
}

int sdfrt_set_overlap(sdfrt_ctx* ctx, int enable) {
    if (!ctx) return SDFRT_ERR_INVALID;
    ctx->overlap = enable != 0;
    return SDFRT_OK;
}

int sdfrt_get_timing_ex(sdfrt_ctx* ctx, float* ms, int* launches, float* ms_union) {
    if (!ctx || !ms || !launches) return SDFRT_ERR_INVALID;
    SDFRT_CUDA(ctx, cudaSetDevice(ctx->device));
    for (int i = 0; i < SDFRT_NUM_KERNEL_CLASSES; ++i) { ms[i] = 0.f; launches[i] = 0; if (ms_union) ms_union[i] = 0.f; }
    if (ms_union) ms_union[SDFRT_NUM_KERNEL_CLASSES] = 0.f;
    // intervals relative to the first recorded event (launches of different batches overlap on the internal streams);
    // slot [SDFRT_NUM_KERNEL_CLASSES] collects every k_march launch (classes 1 and 3)
    std::vector<std::pair<float, float>> iv[SDFRT_NUM_KERNEL_CLASSES + 1];
    const cudaEvent_t ref = ctx->timed.empty() ? nullptr : ctx->timed.front().second.first;
    for (auto& t : ctx->timed) {
        float e = 0.f;
        SDFRT_CUDA(ctx, cudaEventSynchronize(t.second.second));
        SDFRT_CUDA(ctx, cudaEventElapsedTime(&e, t.second.first, t.second.second));
        ms[t.first] += e;
        launches[t.first] += 1;
        if (ms_union) {
            float a = 0.f;
            if (t.second.first != ref) {
                SDFRT_CUDA(ctx, cudaEventSynchronize(t.second.first));
                SDFRT_CUDA(ctx, cudaEventElapsedTime(&a, ref, t.second.first));
            }
            iv[t.first].push_back({a, a + e});
            if (t.first == 1 || t.first == 3) iv[SDFRT_NUM_KERNEL_CLASSES].push_back({a, a + e});
        }
    }
    if (ms_union)
        for (int c = 0; c <= SDFRT_NUM_KERNEL_CLASSES; ++c) {
            std::sort(iv[c].begin(), iv[c].end());
            float end = -1e30f;
            for (auto& q : iv[c]) {
                if (q.first > end) { ms_union[c] += q.second - q.first; end = q.second; }
                else if (q.second > end) { ms_union[c] += q.second - end; end = q.second; }
            }
        }
    for (auto& t : ctx->timed) {
        ctx->event_pool.push_back(t.second.first);
        ctx->event_pool.push_back(t.second.second);
    }
    ctx->timed.clear();
    return SDFRT_OK;
}

int sdfrt_get_timing(sdfrt_ctx* ctx, float* ms, int* launches) { return sdfrt_get_timing_ex(ctx, ms, launches, nullptr); }

int sdfrt_trace_rays(sdfrt_ctx* ctx, const float* o, const float* d, const float* maxt, int n, int differentiable,
                     const sdfrt_ray_out* out, void* stream) {
    if (!ctx || !o || !d || !out || !out->its_t || n < 0) return SDFRT_ERR_INVALID;
    if (!ctx->have_sdf) return fail(ctx, SDFRT_ERR_STATE, "no SDF bound (sdfrt_set_sdf)");
    if (n == 0) return SDFRT_OK;
    SDFRT_CUDA(ctx, cudaSetDevice(ctx->device));
    k_trace_rays<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ctx->sc.G, ctx->sc.T, o, d, maxt, n, differentiable, *out);
    SDFRT_CUDA(ctx, cudaGetLastError());
    return SDFRT_OK;
}

int sdfrt_eval_sdf(sdfrt_ctx* ctx, const float* x, int n, int order, float* out, void* stream) {
    if (!ctx || !x || !out || n < 0 || order < 0 || order > 2) return SDFRT_ERR_INVALID;
    if (!ctx->have_sdf) return fail(ctx, SDFRT_ERR_STATE, "no SDF bound (sdfrt_set_sdf)");
    if (n == 0) return SDFRT_OK;
    SDFRT_CUDA(ctx, cudaSetDevice(ctx->device));
    k_eval_sdf<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ctx->sc.G, x, n, order, out);
    SDFRT_CUDA(ctx, cudaGetLastError());
    return SDFRT_OK;
}

int sdfrt_redistance(sdfrt_ctx* ctx, const float* phi, const float* lower_bound, float* out, int rx, int ry, int rz, int* iters_out,
                     void* stream) {
    if (!ctx || !phi || !out) return SDFRT_ERR_INVALID;
    if (rx < 1 || ry < 1 || rz < 1) return fail(ctx, SDFRT_ERR_INVALID, "bad grid size");
    if (phi == out) return fail(ctx, SDFRT_ERR_INVALID, "redistance cannot run in place");
    SDFRT_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)rx * ry * rz;
    const size_t n_tiles = (size_t)((rx + kRdTile - 1) / kRdTile) * ((ry + kRdTile - 1) / kRdTile) * ((rz + kRdTile - 1) / kRdTile);
    if (n_tiles > 0x7fffffffull) return fail(ctx, SDFRT_ERR_INVALID, "grid too large");
    if (ctx->rd_work_n < 4 * n_tiles) {
        cudaFree(ctx->rd_work); ctx->rd_work = nullptr; ctx->rd_work_n = 0;
        SDFRT_CUDA(ctx, cudaMalloc(&ctx->rd_work, 4 * n_tiles * sizeof(unsigned)));
        ctx->rd_work_n = 4 * n_tiles;
    }
    if (!ctx->rd_state) SDFRT_CUDA(ctx, cudaMalloc(&ctx->rd_state, sizeof(RdState)));
    if (!ctx->rd_err_host) {
        SDFRT_CUDA(ctx, cudaMallocHost((void**)&ctx->rd_err_host, 4 * sizeof(unsigned)));
        SDFRT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->rd_err_event, cudaEventDisableTiming));
    }
    {   // an asynchronous solve that aborted earlier (barrier time-out / iteration cap) must not go unnoticed
        const int prev = sdfrt_redistance_status(ctx);
        if (prev != SDFRT_OK) return prev;
    }
    if (ctx->rd_blocks == 0) {
        int coop = 0, per_sm = 0;
        SDFRT_CUDA(ctx, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
        if (!coop) return fail(ctx, SDFRT_ERR_STATE, "device lacks cooperative launch");
        SDFRT_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_rd_solve, kRdThreads, 0));
        if (per_sm < 1) return fail(ctx, SDFRT_ERR_STATE, "k_rd_solve does not fit on an SM");
        ctx->rd_blocks = per_sm * ctx->num_sms;
    }
    RdGrid g;
    g.phi = phi; g.lower_bound = lower_bound; g.rx = rx; g.ry = ry; g.rz = rz;
    g.hx = 1.0f / (float)rx; g.hy = 1.0f / (float)ry; g.hz = 1.0f / (float)rz;
    const unsigned gb = (unsigned)((n + 255) / 256);
    unsigned* lists = ctx->rd_work;
    unsigned* flags = ctx->rd_work + 2 * n_tiles;
    SDFRT_CUDA(ctx, cudaMemsetAsync(ctx->rd_state, 0, sizeof(RdState), st));
    SDFRT_CUDA(ctx, cudaMemsetAsync(flags, 0, 2 * n_tiles * sizeof(unsigned), st));
    k_rd_init<<<gb, 256, 0, st>>>(g, out);
    {
        RdState* state = ctx->rd_state;
        float hx = g.hx, hy = g.hy, hz = g.hz;
        void* args[] = {&out, &rx, &ry, &rz, &hx, &hy, &hz, &lists, &flags, &state};
        unsigned blocks = (unsigned)std::min<size_t>((size_t)ctx->rd_blocks, n_tiles);
        SDFRT_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)k_rd_solve, dim3(blocks), dim3(kRdThreads), args, 0, st));
    }
    k_rd_finish<<<gb, 256, 0, st>>>(g, out);
    SDFRT_CUDA(ctx, cudaGetLastError());
    ctx->n_launches += 3;
    // the solver's error word always travels to a pinned host word; synchronous callers check it now, asynchronous callers
    // at their next sdfrt_redistance / sdfrt_redistance_status call (an aborted solve leaves unreached nodes at +-1e30)
    SDFRT_CUDA(ctx, cudaMemcpyAsync(ctx->rd_err_host, ctx->rd_state, 4 * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    SDFRT_CUDA(ctx, cudaEventRecord(ctx->rd_err_event, st));
    ctx->rd_err_pending = true;
    if (iters_out) {
        const int rc = sdfrt_redistance_status(ctx);
        if (rc != SDFRT_OK) return rc;
        *iters_out = (int)ctx->rd_err_host[3];
    }
    return SDFRT_OK;
}

int sdfrt_redistance_status(sdfrt_ctx* ctx) {
    if (!ctx) return SDFRT_ERR_INVALID;
    if (!ctx->rd_err_pending) return SDFRT_OK;
    SDFRT_CUDA(ctx, cudaEventSynchronize(ctx->rd_err_event));
    ctx->rd_err_pending = false;
    const unsigned e = ctx->rd_err_host[2];
    if (e != 0u) return fail(ctx, SDFRT_ERR_CUDA, e == 1u ? "redistance: grid barrier timed out" : "redistance: iteration cap reached");
    return SDFRT_OK;
}

int sdfrt_adam_step(sdfrt_ctx* ctx, float* param, float* grad, float* m, float* v, size_t n, double lr, double beta1, double beta2, double eps,
                    int t, float grad_clamp, float lo, float hi, void* stream) {
    if (!ctx || !param || !grad || !m || !v) return SDFRT_ERR_INVALID;
    if (t < 1) return fail(ctx, SDFRT_ERR_INVALID, "Adam step count starts at 1");
    if (n == 0) return SDFRT_OK;
    SDFRT_CUDA(ctx, cudaSetDevice(ctx->device));
    AdamParams a;
    // lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t), evaluated in double like the reference's Python scalars
    a.lr_t = (float)(lr * std::sqrt(1.0 - std::pow(beta2, (double)t)) / (1.0 - std::pow(beta1, (double)t)));
    // the reference's hyper-parameters are Python (double) scalars that meet fp32 arrays one by one: beta and (1 - beta) are
    // rounded to fp32 separately (1 - 0.999 -> 0.001f, not 1.f - 0.999f which is 1.3e-5 off)
    a.beta1 = (float)beta1; a.beta2 = (float)beta2; a.omb1 = (float)(1.0 - beta1); a.omb2 = (float)(1.0 - beta2);
    a.eps = (float)eps; a.grad_clamp = grad_clamp; a.lo = lo; a.hi = hi;
    const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)ctx->num_sms * 16);
    k_adam_step<<<grid, 256, 0, (cudaStream_t)stream>>>(param, grad, m, v, n, a);
    SDFRT_CUDA(ctx, cudaGetLastError());
    ctx->n_launches += 1;
    return SDFRT_OK;
}

int sdfrt_laplacian_reg(sdfrt_ctx* ctx, const float* data, int rx, int ry, int rz, float weight, double* loss_out, float* grad,
                        void* stream) {
    if (!ctx || !data) return SDFRT_ERR_INVALID;
    if (rx < 1 || ry < 1 || rz < 1) return fail(ctx, SDFRT_ERR_INVALID, "bad grid size");
    SDFRT_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)rx * ry * rz;
    int rc;
    if ((rc = ensure(ctx, &ctx->lap_r, &ctx->lap_r_n, n))) return rc;
    const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)ctx->num_sms * 8);
    k_laplacian_residual<<<grid, 256, 0, st>>>(data, ctx->lap_r, rx, ry, rz, (double)weight, loss_out);
    if (grad) k_laplacian_grad<<<grid, 256, 0, st>>>(ctx->lap_r, grad, rx, ry, rz, weight);
    SDFRT_CUDA(ctx, cudaGetLastError());
    ctx->n_launches += grad ? 2 : 1;
    return SDFRT_OK;
}

int sdfrt_render_view_host_ex(sdfrt_ctx* ctx, const float* sdf_host, int rx, int ry, int rz, const float* albedo_host, int ax, int ay,
                              int az, uint32_t seed, int spp, uint32_t seed_grad, int spp_grad, const float* grad_image_host,
                              float* image_host, float* grad_sdf_host, float* grad_albedo_host) {
    if (!ctx || !sdf_host || !image_host || !grad_sdf_host) return SDFRT_ERR_INVALID;
    if (grad_albedo_host && !albedo_host) return fail(ctx, SDFRT_ERR_INVALID, "grad_albedo_host needs albedo_host");
    if (albedo_host && (ax < 1 || ay < 1 || az < 1)) return fail(ctx, SDFRT_ERR_INVALID, "bad albedo grid");
    if (ctx->world != 1) return fail(ctx, SDFRT_ERR_STATE, "host entry point is single-GPU; use the device API with a partition");
    SDFRT_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t nvox = (size_t)rx * ry * rz;
    const size_t nalb = albedo_host ? (size_t)ax * ay * az * 3 : 0;
    const SensorParams& S = ctx->sc.S;
    const size_t nimg = (size_t)S.W * S.H * 3;
    int rc;
    if ((rc = ensure(ctx, &ctx->h_sdf, &ctx->h_sdf_n, nvox))) return rc;
    if ((rc = ensure(ctx, &ctx->h_grad, &ctx->h_grad_n, nvox))) return rc;
    if (albedo_host && (rc = ensure(ctx, &ctx->h_alb, &ctx->h_alb_n, nalb))) return rc;
    if (grad_albedo_host && (rc = ensure(ctx, &ctx->h_galb, &ctx->h_galb_n, nalb))) return rc;
    if ((rc = ensure(ctx, &ctx->h_film, &ctx->h_film_n, film_floats(ctx)))) return rc;
    if ((rc = ensure(ctx, &ctx->h_film2, &ctx->h_film2_n, film_floats(ctx)))) return rc;
    if (ctx->h_image_n < nimg) {
        size_t cap = ctx->h_image_n;
        if ((rc = ensure(ctx, &ctx->h_image, &cap, nimg))) return rc;
        cap = ctx->h_image_n;
        if ((rc = ensure(ctx, &ctx->h_gimg, &cap, nimg))) return rc;
        ctx->h_image_n = nimg;
    }
    // the context's own non-blocking stream: calls on DIFFERENT contexts (one per host thread) overlap their copies and kernels
    cudaStream_t st = ctx->host_stream;
    SDFRT_CUDA(ctx, cudaMemcpyAsync(ctx->h_sdf, sdf_host, nvox * sizeof(float), cudaMemcpyHostToDevice, st));
    const float p0[3] = {ctx->sc.G.px, ctx->sc.G.py, ctx->sc.G.pz};
    if ((rc = sdfrt_set_sdf(ctx, ctx->h_sdf, rx, ry, rz, ctx->have_sdf ? p0 : nullptr))) return rc;
    if (albedo_host) {
        SDFRT_CUDA(ctx, cudaMemcpyAsync(ctx->h_alb, albedo_host, nalb * sizeof(float), cudaMemcpyHostToDevice, st));
        if ((rc = sdfrt_set_albedo(ctx, ctx->h_alb, ax, ay, az))) return rc;
    }
    // primal render and adjoint sweep 1 do not depend on each other: one pipelined call
    if ((rc = sdfrt_render_view(ctx, seed, spp, ctx->h_film, seed_grad, spp_grad, ctx->h_film2, st))) return rc;
    if ((rc = sdfrt_develop(ctx, ctx->h_film, ctx->h_image, st))) return rc;
    SDFRT_CUDA(ctx, cudaMemcpyAsync(image_host, ctx->h_image, nimg * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (grad_image_host) {
        SDFRT_CUDA(ctx, cudaMemcpyAsync(ctx->h_gimg, grad_image_host, nimg * sizeof(float), cudaMemcpyHostToDevice, st));
    } else {
        k_ones<<<(unsigned)((nimg + 255) / 256), 256, 0, st>>>(ctx->h_gimg, nimg);
        SDFRT_CUDA(ctx, cudaGetLastError());
    }
    SDFRT_CUDA(ctx, cudaMemsetAsync(ctx->h_grad, 0, nvox * sizeof(float), st));
    if (grad_albedo_host) SDFRT_CUDA(ctx, cudaMemsetAsync(ctx->h_galb, 0, nalb * sizeof(float), st));
    if ((rc = sdfrt_adjoint_backprop(ctx, ctx->h_film2, ctx->h_gimg, ctx->h_grad, grad_albedo_host ? ctx->h_galb : nullptr, nullptr, st))) return rc;
    SDFRT_CUDA(ctx, cudaMemcpyAsync(grad_sdf_host, ctx->h_grad, nvox * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (grad_albedo_host) SDFRT_CUDA(ctx, cudaMemcpyAsync(grad_albedo_host, ctx->h_galb, nalb * sizeof(float), cudaMemcpyDeviceToHost, st));
    SDFRT_CUDA(ctx, cudaStreamSynchronize(st));
    return SDFRT_OK;
}

int sdfrt_render_view_host(sdfrt_ctx* ctx, const float* sdf_host, int rx, int ry, int rz, uint32_t seed, int spp,
                           uint32_t seed_grad, int spp_grad, const float* grad_image_host, float* image_host,
                           float* grad_sdf_host) {
    return sdfrt_render_view_host_ex(ctx, sdf_host, rx, ry, rz, nullptr, 0, 0, 0, seed, spp, seed_grad, spp_grad, grad_image_host, image_host,
                                     grad_sdf_host, nullptr);
}

}  // extern "C"
