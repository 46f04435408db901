// sdfrt_march.cuh -- persistent-threads wavefront marcher (device only).
//
// One warp holds 32 ray slots.  Every trip of the inner loop performs exactly ONE tricubic SDF
// evaluation for every live slot (primary march, refine iteration or shadow march alike), so the
// 64-tap gather -- the cost of this path -- always runs converged.  Slots whose segment finished wait
// until fewer than kMinActive slots are stepping; then the warp (i) finishes them in one batch
// (surface interaction, emitter sample, film splat, start of the shadow segment) and (ii) refills idle
// slots with new camera samples drawn from a block-level chunk of a global work counter.  The
// reference runs the same work as two Dr.Jit megakernels in which every lane waits for the slowest lane
// of its warp (python/shapes.py:166-238 `mi.Loop`); ncu on the one-thread-per-lane version of this file
// measured 13 of 32 threads active per instruction (profiles/r1_baseline_summary.md).
//
// Work items are camera samples in band order: the rank's film rows are cut into bands of 8 rows and a
// band is walked column by column, so the samples in flight at any time cover a compact 8 x k pixel tile
// (SDF bricks shared through L1/L2), while lane indices -- and therefore the sampler streams -- stay the
// reference's scanline wavefront indices.
#pragma once
#include "sdfrt_lane.cuh"

namespace sdfrt {

constexpr int kMarchBlock = 128;        // threads per block
constexpr int kBlockChunk = 2048;       // work items a block takes from the global counter at a time
constexpr int kMinActive = 20;          // leave the stepping loop when fewer slots are live
constexpr int kBandRows = 8;

struct WorkQueue {
    unsigned long long next;  // global item counter (zeroed before every launch)
};

struct BlockQueue {
    unsigned long long cur, end;
    int lock;
};

// item -> (px, py, sample, lane)
struct WorkItem {
    int px, py;
    unsigned long long lane;
};
__device__ __forceinline__ WorkItem decode_item(const SensorParams& S, const PassParams& ps, unsigned long long k) {
    const unsigned Wb = (unsigned)(S.W + 2 * S.border);
    const unsigned long long row0 = ps.lane_begin / ((unsigned long long)Wb * (unsigned)ps.spp);
    const unsigned nrows = (unsigned)((ps.lane_end - ps.lane_begin) / ((unsigned long long)Wb * (unsigned)ps.spp));
    const unsigned long long per_band = (unsigned long long)kBandRows * Wb * (unsigned)ps.spp;
    const unsigned band = (unsigned)(k / per_band);
    const unsigned r = (unsigned)(k - (unsigned long long)band * per_band);  // < 8 * Wb * spp (fits 32 bits: asserted on the host)
    const unsigned rows_here = min((unsigned)kBandRows, nrows - band * kBandRows);
    const unsigned per_col = rows_here * (unsigned)ps.spp;
    const unsigned x = r / per_col;
    const unsigned rr = r - x * per_col;
    const unsigned y_in = ps.log2_spp >= 0 ? (rr >> ps.log2_spp) : (rr / (unsigned)ps.spp);
    const unsigned s = rr - y_in * (unsigned)ps.spp;
    WorkItem w;
    w.px = (int)x;
    w.py = (int)(row0 + band * kBandRows + y_in);
    w.lane = ((unsigned long long)w.py * Wb + x) * (unsigned)ps.spp + s;
    return w;
}

// leader-only: take up to n consecutive items from the block's chunk (refilling it from the global counter)
__device__ __forceinline__ int fetch_items(BlockQueue* bq, WorkQueue* gq, unsigned long long total, int n, unsigned long long& pos) {
    while (atomicCAS(&bq->lock, 0, 1) != 0) __nanosleep(32);
    __threadfence_block();
    unsigned long long cur = *(volatile unsigned long long*)&bq->cur, end = *(volatile unsigned long long*)&bq->end;
    if (cur >= end && cur < total) {
        const unsigned long long base = atomicAdd(&gq->next, (unsigned long long)kBlockChunk);
        if (base >= total) {
            cur = end = total;
        } else {
            cur = base;
            end = base + kBlockChunk < total ? base + kBlockChunk : total;
        }
    }
    const unsigned long long avail = end - cur;
    const int got = avail < (unsigned long long)n ? (int)avail : n;
    pos = cur;
    *(volatile unsigned long long*)&bq->cur = cur + got;
    *(volatile unsigned long long*)&bq->end = end;
    __threadfence_block();
    atomicExch(&bq->lock, 0);
    return got;
}

__device__ __forceinline__ void splat(const SensorParams& S, float* film, float pfx, float pfy, V3 rgb) {
    const int Wb = S.W + 2 * S.border, Hb = S.H + 2 * S.border;
    const float rad = S.filter_radius;
    const int x0 = (int)ceilf(pfx - rad), y0 = (int)ceilf(pfy - rad);
    const int x1 = (int)floorf(pfx + rad), y1 = (int)floorf(pfy + rad);
    for (int by = y0; by <= y1; ++by) {
        if (by < 0 || by >= Hb) continue;
        const float wy = filter_eval(S, (float)by - pfy);
        if (wy == 0.f) continue;
        for (int bx = x0; bx <= x1; ++bx) {
            if (bx < 0 || bx >= Wb) continue;
            const float w = wy * filter_eval(S, (float)bx - pfx);
            if (w == 0.f) continue;
            float* px = film + ((size_t)by * Wb + bx) * 4;
            if (rgb.x != 0.f) atomicAdd(px + 0, rgb.x * w);
            if (rgb.y != 0.f) atomicAdd(px + 1, rgb.y * w);
            if (rgb.z != 0.f) atomicAdd(px + 2, rgb.z * w);
            atomicAdd(px + 3, w);
        }
    }
}

enum { SLOT_IDLE = 0, SLOT_MARCH = 1, SLOT_REFINE = 2, SLOT_DONE = 3 };
enum { KIND_PRIMARY = 0, KIND_SHADOW = 1 };

template <bool DIFF>
__global__ void __launch_bounds__(kMarchBlock) k_march(const __grid_constant__ SceneParams sc, const __grid_constant__ PassParams ps,
                                                        WorkQueue* __restrict__ gq, float* __restrict__ film,
                                                        float* __restrict__ rec, unsigned long long rec_stride,
                                                        Counters* __restrict__ counters) {
    __shared__ BlockQueue bq;
    if (threadIdx.x == 0) { bq.cur = 0; bq.end = 0; bq.lock = 0; }
    __syncthreads();
    const unsigned FULL = 0xffffffffu;
    const int lane_id = threadIdx.x & 31;
    const unsigned long long total = ps.lane_end - ps.lane_begin;
    const bool warp_primary = DIFF && reparam_at_depth(sc.W, 0);
    const bool warp_shadow = DIFF && reparam_at_depth(sc.W, 1);

    // slot state
    int state = SLOT_IDLE, kind = KIND_PRIMARY;
    Segment seg;
    seg.o = v3(0.f); seg.d = v3(0.f, 0.f, 1.f); seg.t = 0.f; seg.maxt = 0.f; seg.trace_eps = 0.f; seg.its_t = INFINITY; seg.i = 0;
    seg.intersects = false; seg.inside = false;
    DiffAcc acc;
    if (DIFF) { acc.warp_t = 0; acc.prev_surf_dist = 0; acc.weight_sum = 0; acc.extra_weight_sum = 0; acc.prev_sdf_grad_c = v3(0.f);
                acc.mixed_sum_d = v3(0.f); acc.weight_d_sum = v3(0.f); acc.extra_weight_sum_d = v3(0.f); acc.t_d = v3(0.f); }
    bool seg_diff = false;             // this segment runs the differentiable tracer
    unsigned long long lane = 0;       // wavefront lane of the camera sample in this slot
    float pfx = 0.f, pfy = 0.f;        // film position of the sample (block coordinates)
    V3 pend_rgb = v3(0.f);             // radiance added if the shadow segment is unoccluded
    LaneStats st = {0, 0, 0, 0, 0, 0, 0};
    unsigned n_rays = 0;
    bool exhausted = false;

    while (true) {
        // ---- (1) finish completed segments -------------------------------------------------------------
        if (state == SLOT_DONE) {
            float* r = rec + (lane - ps.lane_begin);
            st.n_steps_max = max(st.n_steps_max, (unsigned)seg.i);
            if (kind == KIND_PRIMARY) {
                if (DIFF) {
                    TraceOut tr;
                    if (seg_diff) {
                        diff_finalize(seg, acc, sc.T, tr);
                    } else {
                        tr.its_t = seg.its_t; tr.warp_t = INFINITY; tr.warp_t_d = v3(0.f); tr.warp_weight = 0.f; tr.warp_weight_d = v3(0.f);
                    }
                    r[REC_ITS_T * rec_stride] = tr.its_t;
                    r[REC_TAU * rec_stride] = tr.warp_t;
                    r[REC_TAU_DX * rec_stride] = tr.warp_t_d.x; r[REC_TAU_DY * rec_stride] = tr.warp_t_d.y; r[REC_TAU_DZ * rec_stride] = tr.warp_t_d.z;
                    r[REC_OMEGA * rec_stride] = tr.warp_weight;
                    r[REC_OMEGA_DX * rec_stride] = tr.warp_weight_d.x; r[REC_OMEGA_DY * rec_stride] = tr.warp_weight_d.y; r[REC_OMEGA_DZ * rec_stride] = tr.warp_weight_d.z;
                    r[REC_S_STATE * rec_stride] = 0.f;
                }
                const bool hit = isfinite(seg.its_t);
                st.n_hit += hit;
                V3 rgb = v3(0.f);
                bool shadow = false;
                if (sc.integrator == INTEGRATOR_SILHOUETTE) {
                    rgb = v3(hit ? 1.f : 0.f);
                } else if (!hit) {
                    if (sc.integrator == INTEGRATOR_DIRECT && !sc.Sh.hide_emitters) rgb = V3{sc.Sh.Le[0], sc.Sh.Le[1], sc.Sh.Le[2]};
                } else {
                    const V3 p = seg.o + seg.its_t * seg.d;
                    float phi; V3 g; S3 H;
                    eval_sdf<1>(sc.G, p, phi, g, H);
                    st.n_grad += 1;
                    const V3 n = normalize(g);
                    if (sc.integrator == INTEGRATOR_SIMPLE_SHADING) {
                        rgb = v3(fmaxf(0.f, (n.x + n.y + n.z) * 0.57735026918962576f));
                    } else {
                        const float cos_i = -dot(n, seg.d);
                        Pcg32 rng = sampler_seed(ps.seed, (uint32_t)lane);
                        (void)pcg_next_u32(rng); (void)pcg_next_u32(rng); (void)pcg_next_u32(rng);
                        const float u0 = pcg_next_f32(rng), u1 = pcg_next_f32(rng);
                        const ShadowRay sr = make_shadow_ray(sc.Sh, p, n, u0, u1);
                        const float cos_o = dot(n, sr.d);
                        if ((cos_i > 0.f) & (cos_o > 0.f)) {
                            const V3 rho = eval_albedo(sc.A, p, nullptr, nullptr);
                            st.n_albedo += (sc.A.data != nullptr);
                            const float k = cos_o * 4.f;
                            pend_rgb = V3{rho.x * k * sc.Sh.Le[0], rho.y * k * sc.Sh.Le[1], rho.z * k * sc.Sh.Le[2]};
                            shadow = true;
                            kind = KIND_SHADOW;
                            seg_diff = warp_shadow;
                            const bool active = segment_begin(seg, sc.T, sr.o, sr.d, sr.maxt);
                            if (DIFF && seg_diff) diff_begin(acc, seg);
                            state = active ? SLOT_MARCH : SLOT_DONE;
                        }
                    }
                }
                if (!shadow) {
                    splat(sc.S, film, pfx, pfy, rgb);
                    state = SLOT_IDLE;
                }
            }
            if (state == SLOT_DONE && kind == KIND_SHADOW) {
                const bool occluded = isfinite(seg.its_t);
                if (DIFF && !occluded) {
                    r[REC_S_STATE * rec_stride] = 1.f;
                    TraceOut ts;
                    if (seg_diff) {
                        diff_finalize(seg, acc, sc.T, ts);
                    } else {
                        ts.warp_t = INFINITY; ts.warp_t_d = v3(0.f); ts.warp_weight = 0.f; ts.warp_weight_d = v3(0.f);
                    }
                    r[REC_S_TAU * rec_stride] = ts.warp_t;
                    r[REC_S_TAU_DX * rec_stride] = ts.warp_t_d.x; r[REC_S_TAU_DY * rec_stride] = ts.warp_t_d.y; r[REC_S_TAU_DZ * rec_stride] = ts.warp_t_d.z;
                    r[REC_S_OMEGA * rec_stride] = ts.warp_weight;
                    r[REC_S_OMEGA_DX * rec_stride] = ts.warp_weight_d.x; r[REC_S_OMEGA_DY * rec_stride] = ts.warp_weight_d.y; r[REC_S_OMEGA_DZ * rec_stride] = ts.warp_weight_d.z;
                }
                splat(sc.S, film, pfx, pfy, occluded ? v3(0.f) : pend_rgb);
                state = SLOT_IDLE;
            }
        }

        // ---- (2) refill idle slots ----------------------------------------------------------------------
        const unsigned idle = __ballot_sync(FULL, state == SLOT_IDLE);
        if (idle != 0u && !exhausted) {
            const int n_idle = __popc(idle);
            unsigned long long pos = 0;
            int got = 0;
            if (lane_id == 0) got = fetch_items(&bq, gq, total, n_idle, pos);
            got = __shfl_sync(FULL, got, 0);
            pos = __shfl_sync(FULL, pos, 0);
            if (got == 0) exhausted = true;
            const int my = __popc(idle & ((1u << lane_id) - 1u));
            if (state == SLOT_IDLE && my < got) {
                const WorkItem w = decode_item(sc.S, ps, pos + (unsigned long long)my);
                lane = w.lane;
                Pcg32 rng = sampler_seed(ps.seed, (uint32_t)lane);
                const float r0 = pcg_next_f32(rng), r1 = pcg_next_f32(rng);
                V3 o, d; float maxt;
                sensor_sample_ray(sc.S, (float)(w.px - sc.S.border) + r0, (float)(w.py - sc.S.border) + r1, o, d, maxt);
                float u, v; bool inside; V3 ref;
                sensor_project(sc.S, o + d, u, v, inside, ref);  // re-projection, python/integrators/reparam.py:100-105
                pfx = u + ((float)sc.S.border - 0.5f);
                pfy = v + ((float)sc.S.border - 0.5f);
                kind = KIND_PRIMARY;
                seg_diff = warp_primary;
                const bool active = segment_begin(seg, sc.T, o, d, maxt);
                if (DIFF && seg_diff) diff_begin(acc, seg);
                state = active ? SLOT_MARCH : SLOT_DONE;
                ++n_rays;
            }
        }
        if (__ballot_sync(FULL, state != SLOT_IDLE) == 0u) {
            if (exhausted) break;
            continue;
        }

        // ---- (3) stepping loop: one SDF evaluation per live slot per trip ----------------------------------
        const int min_active = exhausted ? 1 : kMinActive;
        while (true) {
            const bool live = (state == SLOT_MARCH) | (state == SLOT_REFINE);
            if (__popc(__ballot_sync(FULL, live)) < min_active) break;
            if (live) {
                const float tt = (state == SLOT_MARCH) ? seg.t : seg.its_t;
                const V3 x = seg.o + tt * seg.d;
                if (DIFF) {
                    float v; V3 g; S3 H;
                    eval_sdf<2>(sc.G, x, v, g, H);
                    if (state == SLOT_MARCH) {
                        bool active;
                        if (seg_diff) {
                            active = diff_step(seg, acc, sc.T, v, g, H);
                            st.n_hess += 1;
                        } else {
                            active = value_step(seg, sc.T, v);
                            st.n_value += 1;
                        }
                        if (!active) {
                            st.n_steps_max = max(st.n_steps_max, (unsigned)seg.i);
                            if (kind == KIND_PRIMARY && sc.T.refine_intersection && isfinite(seg.its_t)) { state = SLOT_REFINE; seg.i = 0; }
                            else state = SLOT_DONE;
                        }
                    } else {
                        st.n_value += 1;
                        if (!refine_step(seg, v)) state = SLOT_DONE;
                    }
                } else {
                    const float v = eval_value(sc.G, x);
                    st.n_value += 1;
                    if (state == SLOT_MARCH) {
                        if (!value_step(seg, sc.T, v)) {
                            st.n_steps_max = max(st.n_steps_max, (unsigned)seg.i);
                            if (kind == KIND_PRIMARY && sc.T.refine_intersection && isfinite(seg.its_t)) { state = SLOT_REFINE; seg.i = 0; }
                            else state = SLOT_DONE;
                        }
                    } else {
                        if (!refine_step(seg, v)) state = SLOT_DONE;
                    }
                }
            }
        }
    }

    // ---- counters ------------------------------------------------------------------------------------------
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        st.n_value += __shfl_xor_sync(FULL, st.n_value, o);
        st.n_grad += __shfl_xor_sync(FULL, st.n_grad, o);
        st.n_hess += __shfl_xor_sync(FULL, st.n_hess, o);
        st.n_albedo += __shfl_xor_sync(FULL, st.n_albedo, o);
        st.n_hit += __shfl_xor_sync(FULL, st.n_hit, o);
        n_rays += __shfl_xor_sync(FULL, n_rays, o);
        st.n_steps_max = max(st.n_steps_max, __shfl_xor_sync(FULL, st.n_steps_max, o));
    }
    if (lane_id == 0) {
        if (n_rays) atomicAdd(&counters->n_rays, (unsigned long long)n_rays);
        if (st.n_value) atomicAdd(&counters->n_evals_value, (unsigned long long)st.n_value);
        if (st.n_grad) atomicAdd(&counters->n_evals_grad, (unsigned long long)st.n_grad);
        if (st.n_hess) atomicAdd(&counters->n_evals_hess, (unsigned long long)st.n_hess);
        if (st.n_albedo) atomicAdd(&counters->n_albedo_lookups, (unsigned long long)st.n_albedo);
        if (st.n_hit) atomicAdd(&counters->n_hits, (unsigned long long)st.n_hit);
        atomicMax(&counters->n_steps_max, (unsigned long long)st.n_steps_max);
    }
}

}  // namespace sdfrt
