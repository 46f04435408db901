// sdfrt_march.cuh -- the forward sweep as a WAVEFRONT pipeline of compact kernels (device only).
//
//   k_gen_primary   one thread per camera sample: sampler + sensor ray + AABB clip  -> SegIn[item]
//   k_march<DIFF>   persistent-threads sphere tracer over an array of ray segments (primary or shadow):
//                   one warp = 32 ray slots; every trip of the stepping loop performs exactly ONE tricubic
//                   SDF evaluation for every live slot, finished slots are refilled (ballot + popc ranks) from a
//                   block-level chunk of a global work counter by two coalesced 16-byte loads.
//   k_shade         one thread per camera sample: hit -> normal, emitter sample, BSDF gate -> pushes a shadow
//                   segment into a compacted queue (warp-aggregated atomic); everything else is splatted.
//   k_shadow_splat  one thread per shadow segment: visibility -> film splat (+ per-lane shadow record).
//
// The reference runs the same work as one Dr.Jit megakernel per pass in which every lane of a warp waits for the
// slowest lane of its two recorded loops (python/shapes.py:166-238, sdf_direct_reparam.py:21-52).  Splitting at the
// loop boundaries costs ~40 B of queue traffic per ray (HBM: < 1 ms per pass at C3) and keeps the marcher -- where
// > 85 % of the time goes -- small (I-cache resident) and converged.  The one-thread-per-lane form of this code ran
// with 13/32 threads active per instruction, the fused persistent form with 20/32 (profiles/).
//
// Work items are camera samples in band order: the rank's film rows are cut into bands of 8 rows and a band is
// walked column by column, so the samples in flight cover a compact pixel tile (SDF cells shared through L1/L2),
// while lane indices -- and therefore the sampler streams -- stay the reference's scanline wavefront indices.
#pragma once
#include "sdfrt_lane.cuh"

namespace sdfrt {

constexpr int kMarchBlock = 128;       // threads per block of k_march
constexpr unsigned kMaxRegion = 2048;  // segments per region (a power of two in [256, 2048] chosen per launch): the unit a
                                       // marcher block takes from the global counter (2048 = one 8-pixel film column at
                                       // 256 spp); shadow segments are compacted per region
constexpr int kMinActive = 24;         // leave the stepping loop when fewer slots are live

// device work counters of one batch (zeroed before every batch)
enum { WQ_PRIMARY_CHUNK = 0, WQ_SHADOW_CHUNK = 1, WQ_FLAG_COUNT = 2, WQ_FLAG_CHUNK = 3, WQ_COUNT = 4 };

// one ray segment ready to march: d normalised; maxt < 0 marks a segment that never enters the bounding box;
// t0 == 0 <=> origin inside the box (python/shapes.py:141)
struct SegIn {
    float4 a;  // o.xyz, t0
    float4 b;  // d.xyz, maxt
};

// camera sample of a work item: ray + film position of its (re-projected) splat
struct ItemRay {
    V3 o, d;
    float maxt, pfx, pfy;
    unsigned long long lane;
};
__device__ __forceinline__ ItemRay item_ray(const SceneParams& sc, const PassParams& ps, unsigned long long k) {
    const WorkItem w = decode_item(ps, k);
    ItemRay r;
    r.lane = w.lane;
    Pcg32 rng = sampler_seed(ps.seed, (uint32_t)w.lane);
    const float r0 = pcg_next_f32(rng), r1 = pcg_next_f32(rng);
    const bool anti = (w.lane & kAntiBit) != 0ull;
    sensor_sample_ray(sc.S, sample_position(w.px - sc.S.border, r0, anti), sample_position(w.py - sc.S.border, r1, anti), r.o, r.d, r.maxt);
    float u, v; bool inside; V3 ref;
    sensor_project(sc.S, r.o + r.d, u, v, inside, ref);  // re-projection, python/integrators/reparam.py:100-105
    r.pfx = u + ((float)sc.S.border - 0.5f);
    r.pfy = v + ((float)sc.S.border - 0.5f);
    return r;
}

// film position of the re-projected sample from the (unit) camera-ray direction: sensor_project(o + d) only depends on d
__device__ __forceinline__ void film_pos_from_dir(const SensorParams& S, V3 d, float& pfx, float& pfy) {
    const V3 ref = V3{S.R[0][0] * d.x + S.R[1][0] * d.y + S.R[2][0] * d.z, S.R[0][1] * d.x + S.R[1][1] * d.y + S.R[2][1] * d.z,
                      S.R[0][2] * d.x + S.R[1][2] * d.y + S.R[2][2] * d.z};
    const float sx = 0.5f - 0.5f * ref.x / (ref.z * S.tan_half_fov);
    const float sy = 0.5f - 0.5f * S.aspect * ref.y / (ref.z * S.tan_half_fov);
    pfx = sx * (float)S.W + ((float)S.border - 0.5f);
    pfy = sy * (float)S.H + ((float)S.border - 0.5f);
}

__device__ __forceinline__ SegIn make_seg(const TracerParams& T, V3 o, V3 d, float ray_maxt) {
    Segment s;
    const bool active = segment_begin(s, T, o, d, ray_maxt);
    SegIn g;
    g.a = make_float4(s.o.x, s.o.y, s.o.z, s.t);
    g.b = make_float4(s.d.x, s.d.y, s.d.z, active ? s.maxt : -1.f);
    return g;
}

// ---------------------------------------------------------------------------------------------------------------
// Conservative empty-space test.  The tricubic B-spline weights are non-negative and sum to one, so the interpolant at
// x is bounded below by the smallest of its 64 taps.  `brick` holds, per kBrick^3-texel brick, the minimum SDF value over
// the brick dilated by a few bricks (enough to cover the taps of every lookup any sample ray of a pixel can make while the
// pixel's centre ray is in that brick -- sdfrt_core.cuh, empty_space_params).  If all bricks pierced by the centre ray of
// a pixel exceed a threshold T, no sample ray of that pixel can ever evaluate an SDF value <= T:
//   value pass, T = trace_eps:        the march cannot hit, its_t = inf exactly as if it had been marched;
//   DIFF pass,  T = max warp epsilon: additionally the warp weight max(1 - |phi|/eps, 0) at x = o + warp_t d is exactly 0,
//                                     so the lane's tracer record is never used by the adjoint.
// The segments of such pixels are marked inactive by k_gen_primary and never reach the marcher (about 2/3 of the camera
// rays of the benchmark scene).  Results are unchanged; only work is skipped.  One DDA per PIXEL (not per sample).
// ---------------------------------------------------------------------------------------------------------------
// rows4[(z * ry + y) * rx + x] = data[z][y][x .. x+3]; entries with x > rx - 4 are never read (fold_base) and stay unwritten
__global__ void k_expand_rows(GridView G, float4* __restrict__ out) {
    const size_t n = (size_t)G.rx * G.ry * G.rz;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % (size_t)G.rx);
        if (x > G.rx - 4) continue;
        out[i] = make_float4(G.data[i], G.data[i + 1], G.data[i + 2], G.data[i + 3]);
    }
}

// rows8[i * 8 + 0..3] = data[i .. i+3], [4..7] = data[i + rx .. i + rx + 3]; only bases with x <= rx - 4 and y <= ry - 2 are ever read
// with an even row offset j in {0, 2} from by <= ry - 4, i.e. y + 1 <= ry - 1 (fold_base); the rest stays unwritten
__global__ void k_expand_rows8(GridView G, float* __restrict__ out) {
    const size_t n = (size_t)G.rx * G.ry * G.rz;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % (size_t)G.rx), y = (int)((i / (size_t)G.rx) % (size_t)G.ry);
        if (x > G.rx - 4 || y > G.ry - 2) continue;
        float4* o = reinterpret_cast<float4*>(out + i * 8);
        o[0] = make_float4(G.data[i], G.data[i + 1], G.data[i + 2], G.data[i + 3]);
        const size_t u = i + (size_t)G.rx;
        o[1] = make_float4(G.data[u], G.data[u + 1], G.data[u + 2], G.data[u + 3]);
    }
}

__global__ void k_brick_min(GridView G, float* __restrict__ out, int nx, int ny, int nz) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per brick (kBrick^3 = 64 texels)
    if (i >= nx * ny * nz) return;
    const int bx = i % nx, by = (i / nx) % ny, bz = i / (nx * ny);
    float m = INFINITY;
    for (int z = bz * kBrick; z < min((bz + 1) * kBrick, G.rz); ++z)
        for (int y = by * kBrick; y < min((by + 1) * kBrick, G.ry); ++y)
            for (int x = bx * kBrick; x < min((bx + 1) * kBrick, G.rx); ++x) m = fminf(m, G.data[((size_t)z * G.ry + y) * G.rx + x]);
    out[i] = m;
}
// separable dilation (min over +-dil bricks) along one axis: axis 0 = x, 1 = y, 2 = z
__global__ void k_brick_dilate(const float* __restrict__ in, float* __restrict__ out, int nx, int ny, int nz, int dil, int axis) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nx * ny * nz) return;
    const int bx = i % nx, by = (i / nx) % ny, bz = i / (nx * ny);
    float m = INFINITY;
    for (int k = -dil; k <= dil; ++k) {
        const int x = axis == 0 ? min(max(bx + k, 0), nx - 1) : bx, y = axis == 1 ? min(max(by + k, 0), ny - 1) : by,
                  z = axis == 2 ? min(max(bz + k, 0), nz - 1) : bz;
        m = fminf(m, in[(z * ny + y) * nx + x]);
    }
    out[i] = m;
}
__global__ void k_pixel_clear(const __grid_constant__ SceneParams sc, BrickView B, EmptySpaceParams e, unsigned char* __restrict__ mask) {
    const int Wb = sc.S.W + 2 * sc.S.border, Hb = sc.S.H + 2 * sc.S.border;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Wb * Hb) return;
    mask[i] = pixel_provably_clear(sc.S, sc.G, sc.T, B, e, i % Wb, i / Wb) ? 1 : 0;
}

// `mask` (per film pixel, or nullptr): 1 = no sample ray of the pixel needs marching
__global__ void __launch_bounds__(256) k_gen_primary(const __grid_constant__ SceneParams sc, const __grid_constant__ PassParams ps,
                                                      unsigned long long item0, unsigned n, SegIn* __restrict__ segs,
                                                      const unsigned char* __restrict__ mask, Counters* __restrict__ counters) {
    const unsigned i = blockIdx.x * 256u + threadIdx.x;
    bool skipped = false;
    if (i < n) {
        const WorkItem w = decode_item(ps, item0 + i);
        const ItemRay r = item_ray(sc, ps, item0 + i);
        SegIn g = make_seg(sc.T, r.o, r.d, r.maxt);
        if (mask != nullptr && g.b.w >= 0.f && mask[w.py * (sc.S.W + 2 * sc.S.border) + w.px]) {
            g.b.w = -1.f;
            skipped = true;
        }
        segs[i] = g;
    }
    const unsigned m = __ballot_sync(0xffffffffu, skipped);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&counters->n_skipped, (unsigned long long)__popc(m));
}

// ---------------------------------------------------------------------------------------------------------------
// film splat (ImageBlock::put, oracle/ref_semantics.md S4)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void splat(const SensorParams& S, float* film, float pfx, float pfy, V3 rgb, float wgt = 1.f) {
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
            if (wgt != 0.f) atomicAdd(px + 3, w * wgt);
        }
    }
}

// Warp-cooperative splat.  The 32 lanes of a warp hold consecutive work items, i.e. samples of one pixel (or of a few
// vertically adjacent pixels), so their filter footprints fall into one small window of the film.  Instead of
// 32 x ~16 x 4 same-address atomics, every lane publishes its separable filter weights in shared memory and the window's
// cells are summed "transposed" (one cell per lane, loop over the 32 samples), leaving <= 64 x 4 atomics per warp.
// Must be called by all 32 lanes; falls back to per-lane atomics when the window is larger than 8 x 16 pixels.
struct SplatSmem {
    float fx[32][8];
    float fy[32][16];
    float4 c[32];
};
// `wgt`: what the sample adds to the weight channel (1; 0 for the second partial splat of a sample that is splatted in two parts)
__device__ __forceinline__ void warp_splat(const SensorParams& S, float* film, SplatSmem& sm, bool active, float pfx, float pfy, V3 rgb,
                                           float wgt = 1.f) {
    const unsigned FULL = 0xffffffffu;
    const int lane_id = threadIdx.x & 31;
    const float rad = S.filter_radius;
    int x0 = 0x7fffffff, y0 = 0x7fffffff, x1 = -0x7fffffff, y1 = -0x7fffffff;
    if (active) {
        x0 = (int)ceilf(pfx - rad); y0 = (int)ceilf(pfy - rad);
        x1 = (int)floorf(pfx + rad); y1 = (int)floorf(pfy + rad);
    }
    int wx0 = x0, wy0 = y0, wx1 = x1, wy1 = y1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        wx0 = min(wx0, __shfl_xor_sync(FULL, wx0, o)); wy0 = min(wy0, __shfl_xor_sync(FULL, wy0, o));
        wx1 = max(wx1, __shfl_xor_sync(FULL, wx1, o)); wy1 = max(wy1, __shfl_xor_sync(FULL, wy1, o));
    }
    if (wx1 < wx0) return;  // no active lane
    const int Wx = wx1 - wx0 + 1, Wy = wy1 - wy0 + 1;
    if (Wx > 8 || Wy > 16) {
        if (active) splat(S, film, pfx, pfy, rgb, wgt);
        return;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int bx = wx0 + i;
        sm.fx[lane_id][i] = (active && bx >= x0 && bx <= x1) ? filter_eval(S, (float)bx - pfx) : 0.f;
    }
    for (int i = 0; i < Wy; ++i) {
        const int by = wy0 + i;
        sm.fy[lane_id][i] = (active && by >= y0 && by <= y1) ? filter_eval(S, (float)by - pfy) : 0.f;
    }
    sm.c[lane_id] = make_float4(rgb.x, rgb.y, rgb.z, active ? wgt : 0.f);
    __syncwarp();
    const int Wb = S.W + 2 * S.border, Hb = S.H + 2 * S.border;
    for (int cell = lane_id; cell < Wx * Wy; cell += 32) {
        const int cy = cell / Wx, cx = cell - cy * Wx;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
        for (int s = 0; s < 32; ++s) {
            const float w = sm.fx[s][cx] * sm.fy[s][cy];
            const float4 c = sm.c[s];
            acc.x += w * c.x; acc.y += w * c.y; acc.z += w * c.z; acc.w += w * c.w;
        }
        const int bx = wx0 + cx, by = wy0 + cy;
        if (bx >= 0 && bx < Wb && by >= 0 && by < Hb) {
            float* px = film + ((size_t)by * Wb + bx) * 4;
            if (acc.x != 0.f) atomicAdd(px + 0, acc.x);
            if (acc.y != 0.f) atomicAdd(px + 1, acc.y);
            if (acc.z != 0.f) atomicAdd(px + 2, acc.z);
            if (acc.w != 0.f) atomicAdd(px + 3, acc.w);
        }
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------
// k_march: persistent sphere tracer over segments
// ---------------------------------------------------------------------------------------------------------------

// ---------------------------------------------------------------------------------------------------------------
// SMEM box staging for the shadow marchers (north star: "TMA-stages SDF bricks into shared memory").
// The shadow rays of one region (= the samples of one 8-pixel film column) start on a small surface patch.  k_shade records the
// patch's position, and the marcher block that takes the region asks the TMA unit for the box of kBoxBase^3 block bases
// ((kBoxBase + 3)^3 texels) around it: one `cp.async.bulk.tensor.3d` from the SDF grid into one of two shared-memory buffers,
// completion signalled on an mbarrier.  A lookup whose 4x4x4 block lies inside the box of its region reads its 64 taps from shared
// memory (no L1 tag work: the shadow marcher's limit, profiles/); everything else -- rays that left the box, regions whose box is
// still in flight -- takes the global (L1/L2) path as before.  Nobody ever waits for a box.
//   buffer b holds region s_box_region[b] once s_ready[b] == s_fill[b]; a reader re-checks s_fill[b] after its loads (seqlock): the
//   installer bumps s_fill[b] BEFORE it hands the buffer to the TMA unit, so an unchanged value proves the data was intact.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kBoxBase = 16;                 // block bases per axis covered by a box
constexpr int kBoxY = kBoxBase + 3, kBoxZ = kBoxBase + 3;
constexpr int kBoxX = kBoxBase + 4;          // inner extent rounded up to a multiple of 16 bytes (TMA box rule)
constexpr int kBoxFloats = kBoxX * kBoxY * kBoxZ;
constexpr unsigned kBoxBytes = (unsigned)kBoxFloats * 4u;          // bytes one TMA fill delivers
constexpr int kBoxStride = (kBoxFloats + 31) / 32 * 32;            // buffer stride: every TMA destination must be 128-byte aligned

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_test(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0u;
}
// 3-D tiled TMA load: box of the tensor map at element coordinates (x, y, z) -> shared memory, bytes counted on `bar`
__device__ __forceinline__ void tma_load_3d(void* dst, const void* tmap, int x, int y, int z, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(tmap)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
                 : "memory");
}

struct TmaMap { alignas(64) unsigned long long opaque[16]; };  // CUtensorMap (128 bytes), filled by the host with cuTensorMapEncodeTiled

struct BoxShared {
    alignas(128) float data[2][kBoxStride];
    unsigned long long mbar[2];
    int region[2];        // region id held by (or being loaded into) the buffer, -1 = none
    int ox[2], oy[2], oz[2];
    unsigned fill[2];     // fills handed to the TMA unit
    unsigned ready[2];    // fills known to have landed
    unsigned seq;         // regions installed by this block
};

// Hand buffer (seq & 1) to the TMA unit for region `id` (box origin from k_shade, w = 0: the region has no shadow rays / no box).
// One thread.  If the buffer's previous fill is still in flight it is waited for (microseconds; a block installs a region only
// after handing out all slots of the previous one).
__device__ __forceinline__ void box_install(BoxShared& B, const TmaMap* tmap, const int4* __restrict__ region_box, unsigned id) {
    const unsigned b = atomicAdd(&B.seq, 1u) & 1u;
    const int4 org = region_box[id];
    if (B.fill[b] != B.ready[b]) {
        const unsigned par = (B.fill[b] - 1u) & 1u;
        while (!mbar_test(&B.mbar[b], par)) __nanosleep(20);
        B.ready[b] = B.fill[b];
    }
    *(volatile int*)&B.region[b] = -1;
    if (org.w == 0) return;
    *(volatile unsigned*)&B.fill[b] = B.fill[b] + 1u;        // readers of the old contents fail their re-check from here on
    __threadfence_block();
    B.ox[b] = org.x; B.oy[b] = org.y; B.oz[b] = org.z;
    __threadfence_block();
    *(volatile int*)&B.region[b] = (int)id;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&B.mbar[b], kBoxBytes);
    tma_load_3d(&B.data[b][0], tmap, org.x, org.y, org.z, &B.mbar[b]);
}

// Try to serve the 4x4x4 block at (bx, by, bz) of a ray of region `my_region` from shared memory.  Returns false (f untouched or
// garbage) if the region has no ready box, the block is outside it, or the buffer was re-used while reading.
__device__ __forceinline__ bool box_lookup(BoxShared& B, int my_region, int bx, int by, int bz, float f[64]) {
    const int b = (*(volatile int*)&B.region[0] == my_region) ? 0 : ((*(volatile int*)&B.region[1] == my_region) ? 1 : -1);
    if (b < 0) return false;
    const unsigned fill = *(volatile unsigned*)&B.fill[b];
    if (*(volatile unsigned*)&B.ready[b] != fill) {
        if (!mbar_test(&B.mbar[b], (fill - 1u) & 1u)) return false;   // still in flight: global path, no waiting
        *(volatile unsigned*)&B.ready[b] = fill;
    }
    const int lx = bx - B.ox[b], ly = by - B.oy[b], lz = bz - B.oz[b];
    if ((unsigned)lx >= (unsigned)kBoxBase || (unsigned)ly >= (unsigned)kBoxBase || (unsigned)lz >= (unsigned)kBoxBase) return false;
    const float* src = &B.data[b][(lz * kBoxY + ly) * kBoxX + lx];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) f[(k * 4 + j) * 4 + i] = src[(k * kBoxY + j) * kBoxX + i];
    // seqlock re-check (same thread, shared loads in program order): unchanged => the TMA unit did not get the buffer meanwhile
    return *(volatile unsigned*)&B.fill[b] == fill && *(volatile int*)&B.region[b] == my_region;
}

constexpr unsigned long long kQueueExhausted = 0xFFFFFFFFull << 32;

// Segments live in REGIONS of `region` slots; region r holds fill(r) valid segments at its start (primary segments: all
// slots; shadow segments: the count k_shade wrote).  Leader only.  The block's current region is one 64-bit shared word
// (region id << 32 | slots handed out); a fetch is a single shared atomicAdd.  Returns the number of segments taken
// (>= 0) or -1 when all regions are exhausted.
__device__ __forceinline__ unsigned region_fill(const unsigned* region_count, unsigned id, unsigned n_slots, unsigned region) {
    if (region_count) return region_count[id];
    const unsigned long long base = (unsigned long long)id * region;
    return (base + region <= n_slots) ? region : (unsigned)(n_slots - base);
}
template <bool BOX>
__device__ __forceinline__ int fetch_items(unsigned long long* word, unsigned* region_counter, const unsigned* region_count,
                                           unsigned n_slots, unsigned region, int n, unsigned& pos, BoxShared* box, const TmaMap* tmap,
                                           const int4* region_box) {
    for (;;) {
        const unsigned long long old = atomicAdd(word, (unsigned long long)n);
        const unsigned id = (unsigned)(old >> 32), off = (unsigned)old;
        if (id == 0xFFFFFFFFu) return -1;
        const unsigned fill = region_fill(region_count, id, n_slots, region);
        if (off < fill || (off == 0 && fill == 0)) {
            int got = (off + (unsigned)n <= fill) ? n : (int)(fill - off);
            if (off + (unsigned)n >= fill) {  // this fetch used the region up: install the next one
                const unsigned nid = atomicAdd(region_counter, 1u);
                const unsigned long long nw = ((unsigned long long)nid * region >= n_slots) ? kQueueExhausted
                                                                                          : ((unsigned long long)nid << 32);
                if (BOX && nw != kQueueExhausted) box_install(*box, tmap, region_box, nid);   // before the region becomes visible
                atomicExch(word, nw);
            }
            pos = id * region + off;
            return got;
        }
        __nanosleep(40);  // another warp is installing the next region
    }
}

enum { SLOT_IDLE = 0, SLOT_MARCH = 1, SLOT_REFINE = 2, SLOT_DONE = 3 };

// rec_out (DIFF): 9 rows [its_t, warp_t, warp_t_d xyz, warp_weight, warp_weight_d xyz] of stride rec_stride.
// CACHE: 0 = gather every evaluation, 1 = tap cache in registers, 2 = tap cache in shared memory
// APPROACH (value marcher only): a miss is reported as -inf instead of +inf if |phi| ever failed to grow from one step to the next.
//   The differentiable tracer only accumulates silhouette weight once that has happened (`extra_weight_sum`, python/shapes.py:195-201),
//   so a +inf segment provably ends with an invalid warp record and never needs the differentiable march (shadow segments of
//   adjoint sweep 1: 9 in 10).
// dest (DIFF, optional): record of segment `item` goes to rec_out + dest[item] instead of rec_out + item.
// n_slots_dev (optional): the number of valid slots is read from device memory (a list filled by earlier kernels of the stream).
// keep_row0 (DIFF): rows 1..8 of the record are written, row 0 (its_t) is left as the earlier value march of the same segment wrote it.
template <bool DIFF, int CACHE, int ROWS, bool BOX = false, bool APPROACH = false>
__global__ void __launch_bounds__(kMarchBlock, BOX ? 3 : (DIFF ? (CACHE == 1 ? 3 : 4) : (CACHE == 1 ? 4 : 7))) k_march(const GridView G, const TracerParams T, const SegIn* __restrict__ segs,
                                                        const unsigned* __restrict__ region_count, unsigned n_slots, unsigned region, int do_refine, int min_live,
                                                        float* __restrict__ out_its, float* __restrict__ rec_out,
                                                        unsigned long long rec_stride, unsigned* __restrict__ region_counter,
                                                        Counters* __restrict__ counters, const __grid_constant__ TmaMap tmap,
                                                        const int4* __restrict__ region_box, const unsigned* __restrict__ dest = nullptr,
                                                        const unsigned* __restrict__ n_slots_dev = nullptr, bool keep_row0 = false) {
    if (n_slots_dev != nullptr) n_slots = *n_slots_dev;
    __shared__ unsigned long long s_word;
    extern __shared__ unsigned char s_dyn[];   // BOX: one BoxShared (128-byte aligned by the launch)
    BoxShared* box = BOX ? reinterpret_cast<BoxShared*>((reinterpret_cast<size_t>(s_dyn) + 127) & ~(size_t)127) : nullptr;
    unsigned n_box = 0;
    // Tap cache: the 4x4x4 block of the cell a slot evaluated last.  Near a surface the sphere-tracing steps are far
    // smaller than a voxel, so consecutive evaluations mostly fall into the same cell and skip the 64-tap gather.
    // The value-only marcher keeps the block in registers; the differentiable marcher (register-bound) in shared memory,
    // laid out [tap][thread] so that every access is conflict free.
    __shared__ float s_taps[CACHE == 2 ? 64 * kMarchBlock : 1];
    float taps[CACHE == 1 ? 64 : 1];
    int tap_key = -1;
    if (threadIdx.x == 0) {
        const unsigned id = atomicAdd(region_counter, 1u);
        s_word = ((unsigned long long)id * region >= n_slots) ? kQueueExhausted : ((unsigned long long)id << 32);
        if (BOX) {
            box->seq = 0u;
            for (int b = 0; b < 2; ++b) { mbar_init(&box->mbar[b], 1u); box->region[b] = -1; box->fill[b] = 0u; box->ready[b] = 0u; }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            if (s_word != kQueueExhausted) box_install(*box, &tmap, region_box, id);
        }
    }
    __syncthreads();
    const unsigned FULL = 0xffffffffu;
    int my_region = -1;
    const int lane_id = threadIdx.x & 31;
    int state = SLOT_IDLE;
    unsigned item = 0;
    Segment seg;
    seg.o = v3(0.f); seg.d = v3(0.f, 0.f, 1.f); seg.t = 0.f; seg.maxt = -1.f; seg.trace_eps = 0.f; seg.its_t = INFINITY; seg.i = 0;
    seg.intersects = false; seg.inside = false;
    DiffAcc acc;
    if (DIFF) { acc.warp_t = 0; acc.prev_surf_dist = 0; acc.weight_sum = 0; acc.extra_weight_sum = 0; acc.prev_sdf_grad_c = v3(0.f);
                acc.mixed_sum_d = v3(0.f); acc.weight_d_sum = v3(0.f); acc.extra_weight_sum_d = v3(0.f); acc.t_d = v3(0.f); }
    unsigned n_evals = 0, n_steps_max = 0, n_loads = 0;
    bool exhausted = false;
    float prev_surf = 0.f;      // APPROACH: |phi| of the previous march step of this slot's segment
    bool approached = false;

    while (true) {
        // ---- finish (DIFF): tracer outputs of completed segments ----------------------------------------------
        if (DIFF && state == SLOT_DONE) {
            TraceOut tr;
            diff_finalize(seg, acc, T, tr);
            float* r = rec_out + (dest != nullptr ? dest[item] : item);
            if (!keep_row0) r[0 * rec_stride] = tr.its_t;   // (keep_row0: the hit distance of an earlier value march stays)
            r[1 * rec_stride] = tr.warp_t;
            r[2 * rec_stride] = tr.warp_t_d.x; r[3 * rec_stride] = tr.warp_t_d.y; r[4 * rec_stride] = tr.warp_t_d.z;
            r[5 * rec_stride] = tr.warp_weight;
            r[6 * rec_stride] = tr.warp_weight_d.x; r[7 * rec_stride] = tr.warp_weight_d.y; r[8 * rec_stride] = tr.warp_weight_d.z;
            state = SLOT_IDLE;
        }
        // ---- refill idle slots ----------------------------------------------------------------------------------
        const unsigned idle = __ballot_sync(FULL, state == SLOT_IDLE);
        if (idle != 0u && !exhausted) {
            unsigned pos = 0;
            int got = 0;
            if (lane_id == 0) got = fetch_items<BOX>(&s_word, region_counter, region_count, n_slots, region, __popc(idle), pos, box, &tmap, region_box);
            got = __shfl_sync(FULL, got, 0);
            pos = __shfl_sync(FULL, pos, 0);
            if (got < 0) exhausted = true;
            const int my = __popc(idle & ((1u << lane_id) - 1u));
            if (state == SLOT_IDLE && my < got) {
                item = pos + (unsigned)my;
                if (BOX) my_region = (int)(item / region);
                const float4 a = __ldg(&segs[item].a), b = __ldg(&segs[item].b);
                seg.o = V3{a.x, a.y, a.z}; seg.d = V3{b.x, b.y, b.z};
                seg.t = a.w; seg.maxt = b.w;
                seg.inside = (a.w == 0.f);
                seg.intersects = b.w >= 0.f;
                seg.trace_eps = T.trace_eps * fmaxf(b.w, 1.f);
                seg.its_t = INFINITY;
                seg.i = 0;
                if (DIFF) diff_begin(acc, seg, T);
                if (APPROACH) { prev_surf = 0.f; approached = false; }
                state = seg.intersects ? SLOT_MARCH : SLOT_DONE;
                if (!DIFF && state == SLOT_DONE) { out_its[item] = INFINITY; state = SLOT_IDLE; }
            }
        }
        const bool stepping = (state == SLOT_MARCH) | (state == SLOT_REFINE);
        if (__ballot_sync(FULL, stepping) == 0u) {
            if (__ballot_sync(FULL, state == SLOT_DONE) != 0u) continue;
            if (exhausted) break;
            continue;
        }
        // ---- stepping loop: one SDF evaluation per live slot per trip ------------------------------------------------
        const int min_active = exhausted ? 1 : min_live;
        for (;;) {
            const bool live = (state == SLOT_MARCH) | (state == SLOT_REFINE);
            if (__popc(__ballot_sync(FULL, live)) < min_active) break;
            if (live) {
                const float tt = (state == SLOT_MARCH) ? seg.t : seg.its_t;
                const V3 x = seg.o + tt * seg.d;
                ++n_evals;
                bool cont;
                float v; V3 g; S3 H;
                constexpr int ORD = DIFF ? 2 : 0;
                CellSetup<ORD> cs;
                cell_setup<ORD>(G, x, cs);
                const int key = (cs.bz * G.ry + cs.by) * G.rx + cs.bx;
                if constexpr (CACHE == 1) {          // block of the last cell kept in registers
                    if (key != tap_key) {
                        bool from_box = false;
                        if (BOX) { from_box = box_lookup(*box, my_region, cs.bx, cs.by, cs.bz, taps); n_box += from_box; }
                        if (!from_box) load_taps<ROWS>(G, cs.bx, cs.by, cs.bz, taps);
                        tap_key = key;
                        ++n_loads;
                    }
                    reduce_taps<ORD>(taps, cs, G, v, g, H);
                } else if constexpr (CACHE == 2) {   // ... in shared memory, [tap][thread]
                    float f[64];
                    float* st = s_taps + threadIdx.x;
                    if (key != tap_key) {
                        load_taps<ROWS>(G, cs.bx, cs.by, cs.bz, f);
#pragma unroll
                        for (int q = 0; q < 64; ++q) st[q * kMarchBlock] = f[q];
                        tap_key = key;
                        ++n_loads;
                    } else {
#pragma unroll
                        for (int q = 0; q < 64; ++q) f[q] = st[q * kMarchBlock];
                    }
                    reduce_taps<ORD>(f, cs, G, v, g, H);
                } else {
                    float f[64];
                    bool from_box = false;
                    if (BOX) { from_box = box_lookup(*box, my_region, cs.bx, cs.by, cs.bz, f); n_box += from_box; }
                    if (!from_box) load_taps<ROWS>(G, cs.bx, cs.by, cs.bz, f);
                    ++n_loads;
                    reduce_taps<ORD>(f, cs, G, v, g, H);
                }
                if (DIFF) cont = (state == SLOT_MARCH) ? diff_step(seg, acc, T, v, g, H) : refine_step(seg, v);
                else {
                    if (APPROACH && state == SLOT_MARCH) {   // dist_difference >= 0 of python/shapes.py:199-200 (conservatively: incl. 0)
                        const float surf = fabsf(T.step_scale != 1.f ? v * T.step_scale : v);
                        approached = approached | (prev_surf - surf >= 0.f);
                        prev_surf = surf;
                    }
                    cont = (state == SLOT_MARCH) ? value_step(seg, T, v) : refine_step(seg, v);
                }
                if (!cont) {
                    if (state == SLOT_MARCH) {
                        n_steps_max = max(n_steps_max, (unsigned)seg.i);
                        if (do_refine && T.refine_intersection && isfinite(seg.its_t)) { state = SLOT_REFINE; seg.i = 0; }
                        else state = SLOT_DONE;
                    } else {
                        state = SLOT_DONE;
                    }
                    if (!DIFF && state == SLOT_DONE) {
                        out_its[item] = (APPROACH && approached && !isfinite(seg.its_t)) ? -INFINITY : seg.its_t;
                        state = SLOT_IDLE;
                    }
                }
            }
        }
    }
    // ---- counters --------------------------------------------------------------------------------------------------
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_evals += __shfl_xor_sync(FULL, n_evals, o);
        n_loads += __shfl_xor_sync(FULL, n_loads, o);
        if (BOX) n_box += __shfl_xor_sync(FULL, n_box, o);
        n_steps_max = max(n_steps_max, __shfl_xor_sync(FULL, n_steps_max, o));
    }
    if (lane_id == 0) {
        if (n_evals) atomicAdd(DIFF ? &counters->n_evals_hess : &counters->n_evals_value, (unsigned long long)n_evals);
        if (n_loads) atomicAdd(&counters->n_tap_loads, (unsigned long long)n_loads);
        if (BOX && n_box) atomicAdd(&counters->n_box_loads, (unsigned long long)n_box);
        atomicMax(&counters->n_steps_max, (unsigned long long)n_steps_max);
    }
    if (BOX) {   // a block must not exit while the TMA unit may still write into its shared memory
        __syncthreads();
        if (threadIdx.x == 0)
            for (int b = 0; b < 2; ++b)
                if (box->fill[b] != box->ready[b])
                    while (!mbar_test(&box->mbar[b], (box->fill[b] - 1u) & 1u)) __nanosleep(20);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k_shade: <integrator>.sample after the primary intersection (python/integrators/sdf_direct_reparam.py:24-84).
// One block per region of `region` camera samples (region / 256 iterations).  Shadow segments are written compacted at the
// start of the block's own region of the shadow arrays, so their order stays the band order of the camera samples (the
// shadow marcher and the shadow splat then see rays of one 8-pixel film column at a time) and no global atomics or
// prefix sums are needed.
// ---------------------------------------------------------------------------------------------------------------
struct ShadowQueue {
    SegIn* segs;      // [n_slots] shadow segments, region-compacted
    float4* info;     // [n_slots] rgb added if unoccluded, item index (bit pattern) in .w
    unsigned* count;  // [n_regions] valid segments per region
    int4* box;        // [n_regions] or nullptr: origin (block-base texel coordinates) of the SMEM box of the region's shadow rays, w = valid
};

// R8: the order-1 lookup at the hit point reads the marchers' 8x row copy (8 x 256-bit loads instead of 64 scalar ones; same taps,
// same reduction, bit-identical result).  Measured at c3: 30.92 vs 30.90 ms per step -- the lookup is not what binds the kernel's L1
// data stage (the reflectance gather and the splat are); kept as an option (SDFRT_SHADE_ROWS8=1), off by default.
template <bool DIFF, bool R8 = false>
__global__ void __launch_bounds__(256, 3) k_shade(const __grid_constant__ SceneParams sc, const __grid_constant__ PassParams ps,
                                                unsigned long long item0, unsigned n, unsigned region, const SegIn* __restrict__ psegs,
                                                const float* __restrict__ its, ShadowQueue q, ShadowQueue qb /* use_mis: BSDF-sampled rays */,
                                                float* __restrict__ film, float* __restrict__ rec_state, float* __restrict__ rec_bstate,
                                                Counters* __restrict__ counters) {
    __shared__ SplatSmem s_splat[8];
    __shared__ unsigned s_count, s_count_b;
    __shared__ int s_lo[3], s_hi[3];   // block bases of the shadow origins (for the marcher's SMEM box)
    const bool mis = sc.Sh.use_mis && sc.integrator == INTEGRATOR_DIRECT;
    if (threadIdx.x == 0) {
        s_count = 0;
        s_count_b = 0;
        for (int a = 0; a < 3; ++a) { s_lo[a] = 0x7fffffff; s_hi[a] = -0x7fffffff; }
    }
    __syncthreads();
    const unsigned FULL = 0xffffffffu;
    const int lane_id = threadIdx.x & 31;
    const unsigned region_base = blockIdx.x * region;
    unsigned n_grad = 0, n_alb = 0, n_hit = 0, n_rays = 0;
    int my_lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, my_hi[3] = {-0x7fffffff, -0x7fffffff, -0x7fffffff};
    for (unsigned it = 0; it < region / 256u; ++it) {
        const unsigned i = region_base + it * 256u + threadIdx.x;
        if (region_base + it * 256u >= n) break;  // block-uniform
        const bool live = i < n;
        bool push = false, push_b = false;
        SegIn sseg, bseg;
        float4 info = make_float4(0.f, 0.f, 0.f, 0.f), info_b = make_float4(0.f, 0.f, 0.f, 0.f);
        float pfx = 0.f, pfy = 0.f;
        V3 rgb = v3(0.f);
        if (live) {
            ++n_rays;
            const float its_t = its[i];
            const bool hit = isfinite(its_t);
            n_hit += hit;
            // the camera ray comes back from the primary segment (o, unit d): no sampler / sensor math for the 3/4 of the
            // samples that miss; the wavefront lane index is only needed for the emitter sample of a hit
            const float4 sa = __ldg(&psegs[i].a), sb = __ldg(&psegs[i].b);
            struct { V3 o, d; } r;
            r.o = V3{sa.x, sa.y, sa.z}; r.d = V3{sb.x, sb.y, sb.z};
            film_pos_from_dir(sc.S, r.d, pfx, pfy);
            if (sc.integrator == INTEGRATOR_SILHOUETTE) {
                rgb = v3(hit ? 1.f : 0.f);
            } else if (!hit) {
                if (sc.integrator == INTEGRATOR_DIRECT && !sc.Sh.hide_emitters) rgb = miss_rgb(sc, r.d);
            } else {
                const V3 p = r.o + its_t * r.d;
                float phi; V3 g; S3 H;
                if constexpr (R8) {
                    CellSetup<1> cs;
                    cell_setup<1>(sc.G, p, cs);
                    float f[64];
                    load_taps<8>(sc.G, cs.bx, cs.by, cs.bz, f);
                    reduce_taps<1>(f, cs, sc.G, phi, g, H);
                } else {
                    eval_sdf<1>(sc.G, p, phi, g, H);
                }
                ++n_grad;
                const V3 nrm = normalize(g);
                if (sc.integrator == INTEGRATOR_SIMPLE_SHADING) {
                    rgb = v3(fmaxf(0.f, (nrm.x + nrm.y + nrm.z) * 0.57735026918962576f));
                } else {
                    const float cos_i = -dot(nrm, r.d);
                    Pcg32 rng = sampler_seed(ps.seed, (uint32_t)decode_item(ps, item0 + i).lane);
                    (void)pcg_next_u32(rng); (void)pcg_next_u32(rng); (void)pcg_next_u32(rng);
                    const float u0 = pcg_next_f32(rng), u1 = pcg_next_f32(rng);
                    float pdf_e;
                    const ShadowRay sr = sample_emitter_ray(sc, p, nrm, u0, u1, pdf_e);
                    const float cos_o = dot(nrm, sr.d);
                    // a shadow ray whose BSDF gate is closed contributes exactly zero whatever it hits: not traced
                    if ((cos_i > 0.f) & (cos_o > 0.f) & (pdf_e > 0.f)) {
                        const V3 rho = eval_albedo(sc.A, p, nullptr, nullptr);
                        n_alb += (sc.A.data != nullptr);
                        const V3 c = nee_rgb(sc, rho, cos_o, sr.d, pdf_e, mis);
                        info = make_float4(c.x, c.y, c.z, __uint_as_float(i));
                        sseg = make_seg(sc.T, sr.o, sr.d, sr.maxt);
                        push = true;
                        if (q.box != nullptr) {   // (boxes are staged for the emitter-sampled rays only)
                            const CellCoord cc = cell_coord(sc.G, sr.o);
                            my_lo[0] = min(my_lo[0], cc.ix); my_lo[1] = min(my_lo[1], cc.iy); my_lo[2] = min(my_lo[2], cc.iz);
                            my_hi[0] = max(my_hi[0], cc.ix); my_hi[1] = max(my_hi[1], cc.iy); my_hi[2] = max(my_hi[2], cc.iz);
                        }
                    }
                    if (mis) {   // BSDF sampling (python/integrators/sdf_direct_reparam.py:87-107): escaped rays see the emitter
                        (void)pcg_next_u32(rng);
                        const float b0 = pcg_next_f32(rng), b1 = pcg_next_f32(rng);
                        const BsdfRay br = make_bsdf_ray(p, nrm, cos_i, b0, b1);
                        if (br.active) {
                            const V3 rho = eval_albedo(sc.A, p, nullptr, nullptr);
                            n_alb += (sc.A.data != nullptr);
                            const V3 c = bsdf_rgb(sc, rho, br.d, br.pdf);
                            // bit 31 of the item word: the sample's weight channel is splatted with its emitter-sampled part
                            info_b = make_float4(c.x, c.y, c.z, __uint_as_float(i | (push ? 0x80000000u : 0u)));
                            bseg = make_seg(sc.T, br.o, br.d, INFINITY);
                            push_b = true;
                        }
                    }
                }
            }
            if (DIFF) { rec_state[i] = 0.f; if (mis) rec_bstate[i] = 0.f; }
        }
        warp_splat(sc.S, film, s_splat[threadIdx.x >> 5], live && !push && !push_b, pfx, pfy, rgb);
        if (mis) {
            const unsigned mb = __ballot_sync(FULL, push_b);
            if (mb) {
                unsigned base = 0;
                if (lane_id == (__ffs(mb) - 1)) base = atomicAdd(&s_count_b, (unsigned)__popc(mb));
                base = __shfl_sync(FULL, base, __ffs(mb) - 1);
                if (push_b) {
                    const unsigned slot = region_base + base + (unsigned)__popc(mb & ((1u << lane_id) - 1u));
                    qb.segs[slot] = bseg;
                    qb.info[slot] = info_b;
                }
            }
        }
        const unsigned m = __ballot_sync(FULL, push);
        if (m) {
            unsigned base = 0;
            if (lane_id == (__ffs(m) - 1)) base = atomicAdd(&s_count, (unsigned)__popc(m));
            base = __shfl_sync(FULL, base, __ffs(m) - 1);
            if (push) {
                const unsigned slot = region_base + base + (unsigned)__popc(m & ((1u << lane_id) - 1u));
                q.segs[slot] = sseg;
                q.info[slot] = info;
            }
        }
    }
    if (q.box != nullptr) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            int lo = my_lo[a], hi = my_hi[a];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { lo = min(lo, __shfl_xor_sync(FULL, lo, o)); hi = max(hi, __shfl_xor_sync(FULL, hi, o)); }
            if (lane_id == 0 && hi >= lo) { atomicMin(&s_lo[a], lo); atomicMax(&s_hi[a], hi); }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        q.count[blockIdx.x] = s_count;
        if (mis) qb.count[blockIdx.x] = s_count_b;
        if (q.box != nullptr) {   // box of kBoxBase^3 block bases centred on the origin patch, kept inside the grid
            int4 b = make_int4(0, 0, 0, 0);
            if (s_count > 0u) {
                const int r[3] = {sc.G.rx, sc.G.ry, sc.G.rz};
                int o3[3];
                for (int a = 0; a < 3; ++a) {
                    int c = (s_lo[a] + s_hi[a]) / 2 - kBoxBase / 2;
                    c = max(0, min(c, r[a] - (kBoxBase + 3)));
                    if (a == 0) c &= ~3;   // the TMA unit faults on a box whose first row does not start on a 16-byte boundary
                    o3[a] = c;
                }
                b = make_int4(o3[0], o3[1], o3[2], 1);
            }
            q.box[blockIdx.x] = b;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_grad += __shfl_xor_sync(FULL, n_grad, o);
        n_alb += __shfl_xor_sync(FULL, n_alb, o);
        n_hit += __shfl_xor_sync(FULL, n_hit, o);
        n_rays += __shfl_xor_sync(FULL, n_rays, o);
    }
    if (lane_id == 0) {
        if (n_rays) atomicAdd(&counters->n_rays, (unsigned long long)n_rays);
        if (n_grad) atomicAdd(&counters->n_evals_grad, (unsigned long long)n_grad);
        if (n_alb) atomicAdd(&counters->n_albedo_lookups, (unsigned long long)n_alb);
        if (n_hit) atomicAdd(&counters->n_hits, (unsigned long long)n_hit);
    }
}

// one block per region: visibility -> splat; DIFF: scatter the shadow tracer record to the lane records
// BQ: the queue of BSDF-sampled rays (use_mis): escaped = contribution, records go to the REC_B_* rows, and the weight channel is
// only added for samples that have no emitter-sampled part (bit 31 of the item word)
// Deferred differentiable march (DIFF, `flag` given): the segments were marched by the value tracer with APPROACH; unoccluded ones
// get state 1 and an invalid warp record (tau = inf), and those reported as -inf are appended to the sweep's list `flag` for the one
// differentiable march at the end of the sweep, which overwrites their record rows in place.
struct FlagList {
    SegIn* segs;        // [capacity]
    unsigned* dest;     // [capacity] record offset of the segment's lane: row_base * stride + lane item (the marcher adds row * stride)
    unsigned* count;    // device counter
    unsigned capacity;
};
template <bool DIFF, bool BQ = false>
__global__ void __launch_bounds__(256) k_shadow_splat(const __grid_constant__ SceneParams sc, const __grid_constant__ PassParams ps,
                                                       unsigned long long item0, unsigned region, const SegIn* __restrict__ psegs,
                                                       const unsigned* __restrict__ region_count, const float4* __restrict__ info, const float* __restrict__ s_its,
                                                       const float* __restrict__ srec, unsigned long long srec_stride,
                                                       float* __restrict__ film, float* __restrict__ rec,
                                                       unsigned long long rec_stride, const SegIn* __restrict__ ssegs = nullptr, FlagList flag = FlagList{nullptr, nullptr, nullptr, 0u}) {
    __shared__ SplatSmem s_splat[8];
    const unsigned cnt = region_count[blockIdx.x];
    const unsigned region_base = blockIdx.x * region;
    for (unsigned j0 = 0; j0 < cnt; j0 += 256u) {
        const unsigned j = j0 + threadIdx.x;
        const bool live = j < cnt;
        float pfx = 0.f, pfy = 0.f, wgt = 1.f;
        V3 rgb = v3(0.f);
        bool want_flag = false;
        unsigned flag_item = 0u;
        if (live) {
            const unsigned slot = region_base + j;
            const float4 inf = info[slot];
            const unsigned iw = __float_as_uint(inf.w);
            const unsigned i = BQ ? (iw & 0x7fffffffu) : iw;
            if (BQ && (iw & 0x80000000u)) wgt = 0.f;
            flag_item = (unsigned)item0 + i;   // lane item within the sweep (records are indexed by it)
            const float its_s = s_its[slot];  // its_t of the shadow segment (value tracer's output, or row 0 of the DIFF records)
            const bool occluded = isfinite(its_s);
            const float4 sb = __ldg(&psegs[i].b);
            film_pos_from_dir(sc.S, V3{sb.x, sb.y, sb.z}, pfx, pfy);
            if (!occluded) rgb = V3{inf.x, inf.y, inf.z};
            if (DIFF && !occluded) {
                float* d = rec + i;
                d[(BQ ? REC_B_STATE : REC_S_STATE) * rec_stride] = 1.f;
                if (flag.segs != nullptr) {
                    d[(BQ ? REC_B_TAU : REC_S_TAU) * rec_stride] = INFINITY;   // invalid warp record unless the deferred march says otherwise
                    want_flag = its_s < 0.f;                                  // -inf: |phi| stopped growing somewhere along the segment
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) d[((BQ ? REC_B_TAU : REC_S_TAU) + k) * rec_stride] = srec[(1 + k) * srec_stride + slot];
                }
            }
        }
        if (DIFF && flag.segs != nullptr) {   // warp-aggregated append to the sweep's list
            const unsigned m = __ballot_sync(0xffffffffu, want_flag);
            if (m) {
                const int lane_id = threadIdx.x & 31;
                unsigned base = 0;
                if (lane_id == (__ffs(m) - 1)) base = atomicAdd(flag.count, (unsigned)__popc(m));
                base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
                if (want_flag) {
                    const unsigned k = base + (unsigned)__popc(m & ((1u << lane_id) - 1u));
                    if (k < flag.capacity) {
                        flag.segs[k] = ssegs[region_base + j];
                        flag.dest[k] = (unsigned)((BQ ? REC_B_STATE : REC_S_STATE) * rec_stride) + flag_item;
                    }
                }
            }
        }
        warp_splat(sc.S, film, s_splat[threadIdx.x >> 5], live, pfx, pfy, rgb, wgt);
    }
}

}  // namespace sdfrt
