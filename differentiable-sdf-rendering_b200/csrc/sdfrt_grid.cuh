// sdfrt_grid.cuh -- the per-iteration grid update on the other side of the gradient:
//
//   redistance            redistancing.redistance(sdf) = fastsweep.redistance (python/redistancing.py:4-13), applied to
//                         max(sdf, bbox_sdf) after every optimiser step (python/variables.py:185-189)
//   adam_step             SdfVariable.validate_gradient (python/variables.py:193-199: NaN scrub + clamp to +-0.1) fused with
//                         mi.ad.Adam.step (python/shape_opt.py:53,102) and the value clamps of VolumeVariable.validate
//                         (python/variables.py:114-117)
//   laplacian_reg         regularizations.eval_discrete_laplacian_reg (python/regularizations.py:5-25) and its gradient
//
// Redistancing.  The discrete problem is the one fastsweep solves (first-order Godunov upwind Eikonal equation, nodes next
// to the zero level set frozen at their sub-voxel distance; restated in oracle/redistance_oracle.c).  Its solution is the
// unique fixed point of the monotone update u <- min(u, solve(neighbours)), so any update order reaches it.  fastsweep's
// order (Detrixhe et al.: 8 sweeps x (rx+ry+rz) diagonal planes, one launch or grid-wide barrier per plane) is latency
// bound on a B200 (6 k dependent steps at 256^3), so the order used here is a block-asynchronous one:
//   - the grid is cut into 8^3 tiles; a tile is solved to ITS fixed point in shared memory (Jacobi rounds on the 8^3 nodes
//     against a frozen 1-node halo), then written back;
//   - a tile whose face values changed puts its face neighbours on the next work list (flag + list append, no duplicates);
//   - one persistent cooperative kernel runs list after list with a grid barrier in between until a list comes back empty.
// The front moves one tile (8 nodes) per barrier instead of one plane, and only tiles on the front are touched:
// ~(rx+ry+rz)/8 barriers of a few microseconds and a few sweeps' worth of memory traffic in total.
#pragma once
#include "sdfrt_core.cuh"

namespace sdfrt {

constexpr float kRdInf = 1.0e30f;
constexpr int kRdTile = 8;
constexpr int kRdThreads = kRdTile * kRdTile * kRdTile;
constexpr int kRdMaxRounds = 48;      // Jacobi rounds per tile visit (a visit that hits the cap re-queues the tile)
constexpr int kRdMaxIters = 4096;     // work lists per solve

// fp32 operations spelled out so that host (oracle order, -ffp-contract=off) and device agree step by step
#if defined(__CUDA_ARCH__)
#define RD_MUL(a, b) __fmul_rn((a), (b))
#define RD_ADD(a, b) __fadd_rn((a), (b))
#define RD_SUB(a, b) __fsub_rn((a), (b))
#define RD_DIV(a, b) __fdiv_rn((a), (b))
#define RD_SQRT(a) __fsqrt_rn((a))
#else
#define RD_MUL(a, b) ((a) * (b))
#define RD_ADD(a, b) ((a) + (b))
#define RD_SUB(a, b) ((a) - (b))
#define RD_DIV(a, b) ((a) / (b))
#define RD_SQRT(a) sqrtf((a))
#endif

// Godunov update: smallest u with sum_k max((u - a_k) / h_k, 0)^2 = 1 (a_k = smaller neighbour along axis k).
// w_k = 1 / h_k^2 is passed in (same value the oracle computes inline, hoisted out of the solver's inner loop).
struct RdAxis {
    float h, w;
};
SDFRT_HD RdAxis rd_axis(float h) { return RdAxis{h, RD_DIV(1.0f, RD_MUL(h, h))}; }
SDFRT_HD float rd_solve(float a0, RdAxis x0, float a1, RdAxis x1, float a2, RdAxis x2) {
    float t;
    RdAxis tx;
    if (a1 < a0) { t = a0; a0 = a1; a1 = t; tx = x0; x0 = x1; x1 = tx; }
    if (a2 < a1) { t = a1; a1 = a2; a2 = t; tx = x1; x1 = x2; x2 = tx; }
    if (a1 < a0) { t = a0; a0 = a1; a1 = t; tx = x0; x0 = x1; x1 = tx; }
    if (a0 >= kRdInf) return kRdInf;
    float u = RD_ADD(a0, x0.h);
    if (u <= a1) return u;
    const float w0 = x0.w, w1 = x1.w;
    const float b1 = RD_SUB(a1, a0);
    {
        const float A = RD_ADD(w0, w1);
        const float B = RD_MUL(w1, b1);
        const float C = RD_SUB(RD_MUL(RD_MUL(w1, b1), b1), 1.0f);
        const float D = RD_SUB(RD_MUL(B, B), RD_MUL(A, C));
        u = RD_ADD(a0, RD_DIV(RD_ADD(B, RD_SQRT(D > 0.0f ? D : 0.0f)), A));
        if (u <= a2) return u;
    }
    {
        const float w2 = x2.w;
        const float b2 = RD_SUB(a2, a0);
        const float A = RD_ADD(RD_ADD(w0, w1), w2);
        const float B = RD_ADD(RD_MUL(w1, b1), RD_MUL(w2, b2));
        const float C = RD_SUB(RD_ADD(RD_MUL(RD_MUL(w1, b1), b1), RD_MUL(RD_MUL(w2, b2), b2)), 1.0f);
        const float D = RD_SUB(RD_MUL(B, B), RD_MUL(A, C));
        u = RD_ADD(a0, RD_DIV(RD_ADD(B, RD_SQRT(D > 0.0f ? D : 0.0f)), A));
    }
    return u;
}

struct RdGrid {
    const float* phi;          // input
    const float* lower_bound;  // optional: the input is max(phi, lower_bound) (python/variables.py:185-187), or nullptr
    int rx, ry, rz;
    float hx, hy, hz;
};
SDFRT_HD float rd_input(const RdGrid& g, size_t i) {
    const float p = g.phi[i];
    return g.lower_bound ? fmaxf(p, g.lower_bound[i]) : p;
}

// frozen-node initialisation (scikit-fmm style, see the oracle): unsigned distance or kRdInf
SDFRT_HD float rd_init_node(const RdGrid& g, int x, int y, int z) {
    const size_t sy = (size_t)g.rx, sz = (size_t)g.rx * g.ry;
    const size_t i = (size_t)z * sz + (size_t)y * sy + (size_t)x;
    const float p = rd_input(g, i);
    if (p == 0.0f) return 0.0f;
    float dsum = 0.0f;
    bool borders = false;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float best = 0.0f;
        const float h = a == 0 ? g.hx : (a == 1 ? g.hy : g.hz);
#pragma unroll
        for (int s = -1; s <= 1; s += 2) {
            const int xx = x + (a == 0 ? s : 0), yy = y + (a == 1 ? s : 0), zz = z + (a == 2 ? s : 0);
            if (xx < 0 || xx >= g.rx || yy < 0 || yy >= g.ry || zz < 0 || zz >= g.rz) continue;
            const float q = rd_input(g, (size_t)zz * sz + (size_t)yy * sy + (size_t)xx);
            if (RD_MUL(p, q) < 0.0f) {
                const float d = RD_DIV(RD_MUL(h, p), RD_SUB(p, q));
                if (best == 0.0f || best > d) best = d;
                borders = true;
            }
        }
        if (best > 0.0f) dsum = RD_ADD(dsum, RD_DIV(1.0f, RD_MUL(best, best)));
    }
    if (!borders) return kRdInf;
    return RD_SQRT(RD_DIV(1.0f, dsum));
}

#if defined(__CUDACC__)

// Working encoding of u: frozen nodes carry the sign bit (-u, -0.0f for a zero), free nodes are >= 0 (kRdInf = unknown).
__global__ void k_rd_init(RdGrid g, float* __restrict__ u) {
    const size_t n = (size_t)g.rx * g.ry * g.rz;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int x = (int)(i % (size_t)g.rx), y = (int)((i / (size_t)g.rx) % (size_t)g.ry), z = (int)(i / ((size_t)g.rx * g.ry));
    const float d = rd_init_node(g, x, y, z);
    u[i] = d < kRdInf ? -d : kRdInf;
}

__global__ void k_rd_finish(RdGrid g, float* __restrict__ u) {
    const size_t n = (size_t)g.rx * g.ry * g.rz;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a = fabsf(u[i]);
    u[i] = rd_input(g, i) < 0.0f ? -a : a;
}

struct RdState {
    unsigned barrier_count, barrier_gen;
    unsigned error;                  // 1 = barrier timed out, 2 = iteration cap
    unsigned iters;                  // work lists processed
    unsigned long long tile_visits;
    unsigned count[kRdMaxIters + 2]; // entries of work list i
};

// all blocks of the (cooperatively launched, hence co-resident) grid
__device__ __forceinline__ bool rd_grid_barrier(RdState* s, unsigned nblocks) {
    __shared__ unsigned s_ok;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned ok = 1u;
        __threadfence();
        const unsigned gen = atomicAdd(&s->barrier_gen, 0u);
        if (atomicAdd(&s->barrier_count, 1u) == nblocks - 1u) {
            atomicExch(&s->barrier_count, 0u);
            __threadfence();
            atomicAdd(&s->barrier_gen, 1u);
        } else {
            unsigned spins = 0;
            while (atomicAdd(&s->barrier_gen, 0u) == gen) {
                __nanosleep(64);
                if (++spins > (1u << 24)) { atomicExch(&s->error, 1u); ok = 0u; break; }  // ~1 s: never on a healthy launch
            }
        }
        __threadfence();
        s_ok = ok;
    }
    __syncthreads();
    return s_ok != 0u;
}

// lists[2][n_tiles], flags[2][n_tiles] (parity of the work list); u in the working encoding.
__global__ void __launch_bounds__(kRdThreads) k_rd_solve(float* __restrict__ u, int rx, int ry, int rz, float hx, float hy, float hz,
                                                         unsigned* __restrict__ lists, unsigned* __restrict__ flags, RdState* __restrict__ st) {
    constexpr int P = kRdTile + 2;
    __shared__ float s_u[P * P * P];
    __shared__ unsigned s_face[2];
    const int tx = threadIdx.x & 7, ty = (threadIdx.x >> 3) & 7, tz = threadIdx.x >> 6;
    const int ntx = (rx + kRdTile - 1) / kRdTile, nty = (ry + kRdTile - 1) / kRdTile, ntz = (rz + kRdTile - 1) / kRdTile;
    const unsigned n_tiles = (unsigned)(ntx * nty * ntz);
    const size_t sy = (size_t)rx, sz = (size_t)rx * ry;
    const RdAxis axx = rd_axis(hx), axy = rd_axis(hy), axz = rd_axis(hz);
    unsigned visits = 0;
    unsigned it = 0;
    for (;; ++it) {
        const unsigned count = it == 0 ? n_tiles : __ldcg(&st->count[it]);  // list 0 = every tile
        if (count == 0u) break;
        if (it >= (unsigned)kRdMaxIters) { if (threadIdx.x == 0) atomicExch(&st->error, 2u); break; }
        const unsigned* list = lists + (size_t)(it & 1u) * n_tiles;
        unsigned* next_list = lists + (size_t)((it + 1u) & 1u) * n_tiles;
        unsigned* my_flags = flags + (size_t)(it & 1u) * n_tiles;
        unsigned* next_flags = flags + (size_t)((it + 1u) & 1u) * n_tiles;
        for (unsigned k = blockIdx.x; k < count; k += gridDim.x) {
            const unsigned tile = it == 0 ? k : __ldcg(&list[k]);
            const int tix = (int)(tile % (unsigned)ntx), tiy = (int)((tile / (unsigned)ntx) % (unsigned)nty), tiz = (int)(tile / (unsigned)(ntx * nty));
            const int x0 = tix * kRdTile - 1, y0 = tiy * kRdTile - 1, z0 = tiz * kRdTile - 1;
            if (threadIdx.x == 0 && it != 0) my_flags[tile] = 0u;
            // tile + halo -> shared memory (outside the grid: unknown)
            bool any_known = false;
            for (int q = threadIdx.x; q < P * P * P; q += kRdThreads) {
                const int lx = q % P, ly = (q / P) % P, lz = q / (P * P);
                const int gx = x0 + lx, gy = y0 + ly, gz = z0 + lz;
                float v = kRdInf;
                if (gx >= 0 && gx < rx && gy >= 0 && gy < ry && gz >= 0 && gz < rz) v = __ldcg(&u[(size_t)gz * sz + (size_t)gy * sy + (size_t)gx]);
                s_u[q] = v;
                any_known |= fabsf(v) < kRdInf;
            }
            if (threadIdx.x == 0) s_face[(visits + 1u) & 1u] = 0u;  // buffer of the visit that starts here
            if (!__syncthreads_or(any_known)) continue;  // nothing to propagate yet (block-uniform)
            ++visits;
            const int me = ((tz + 1) * P + (ty + 1)) * P + (tx + 1);
            const int gx = x0 + 1 + tx, gy = y0 + 1 + ty, gz = z0 + 1 + tz;
            const bool inside = gx < rx && gy < ry && gz < rz;
            const float orig = s_u[me];
            const bool is_free = inside && !(__float_as_uint(orig) >> 31);
            float cur = orig;
            bool capped = true;
            float pax = -1.f, pay = -1.f, paz = -1.f;  // neighbour minima the last solve saw: unchanged inputs, unchanged result
            for (int round = 0; round < kRdMaxRounds; ++round) {
                bool ch = false;
                float cand = cur;
                if (is_free) {
                    const float ax = fminf(fabsf(s_u[me - 1]), fabsf(s_u[me + 1]));
                    const float ay = fminf(fabsf(s_u[me - P]), fabsf(s_u[me + P]));
                    const float az = fminf(fabsf(s_u[me - P * P]), fabsf(s_u[me + P * P]));
                    if (ax != pax || ay != pay || az != paz) {
                        cand = rd_solve(ax, axx, ay, axy, az, axz);
                        ch = cand < cur;
                        pax = ax; pay = ay; paz = az;
                    }
                }
                __syncthreads();
                if (ch) { s_u[me] = cand; cur = cand; }
                if (!__syncthreads_or(ch)) { capped = false; break; }
            }
            const bool changed = cur < orig;
            if (changed) __stcg(&u[(size_t)gz * sz + (size_t)gy * sy + (size_t)gx], cur);
            // a changed face wakes the neighbour behind it; a visit that ran into the round cap re-queues the tile itself
            unsigned fm = 0u;
            if (changed) fm = (tx == 0 ? 1u : 0u) | (tx == kRdTile - 1 ? 2u : 0u) | (ty == 0 ? 4u : 0u) | (ty == kRdTile - 1 ? 8u : 0u) |
                              (tz == 0 ? 16u : 0u) | (tz == kRdTile - 1 ? 32u : 0u);
            fm = __reduce_or_sync(0xffffffffu, fm);
            unsigned* face = &s_face[visits & 1u];          // double-buffered: the next visit's reset cannot race with this read
            if ((threadIdx.x & 31) == 0 && fm) atomicOr(face, fm);
            __syncthreads();
            const unsigned faces = *face;
            const int f0 = faces & 1u, f1 = faces & 2u, f2 = faces & 4u, f3 = faces & 8u, f4 = faces & 16u, f5 = faces & 32u;
            if (threadIdx.x < 7) {
                int nx = tix, ny = tiy, nz = tiz;
                int wake;
                switch (threadIdx.x) {
                    case 0: wake = f0; nx -= 1; break;
                    case 1: wake = f1; nx += 1; break;
                    case 2: wake = f2; ny -= 1; break;
                    case 3: wake = f3; ny += 1; break;
                    case 4: wake = f4; nz -= 1; break;
                    case 5: wake = f5; nz += 1; break;
                    default: wake = capped; break;
                }
                if (wake && nx >= 0 && nx < ntx && ny >= 0 && ny < nty && nz >= 0 && nz < ntz) {
                    const unsigned nt = (unsigned)((nz * nty + ny) * ntx + nx);
                    if (atomicExch(&next_flags[nt], 1u) == 0u) next_list[atomicAdd(&st->count[it + 1u], 1u)] = nt;
                }
            }
        }
        if (!rd_grid_barrier(st, gridDim.x)) break;
    }
    if (threadIdx.x == 0 && visits) atomicAdd(&st->tile_visits, (unsigned long long)visits);
    if (blockIdx.x == 0 && threadIdx.x == 0) st->iters = it;
}

// ---------------------------------------------------------------------------------------------------------------
// Adam step with the reference's gradient / value validation fused in (one pass over param, grad, m, v).
//   g   <- isnan(g) ? 0 : clamp(g, -grad_clamp, grad_clamp)        python/variables.py:193-199 (grad_clamp <= 0: off)
//   m   <- beta1 m + (1 - beta1) g ;  v <- beta2 v + (1 - beta2) g^2
//   p   <- p - lr_t m / (sqrt(v) + eps),  lr_t = lr sqrt(1 - beta2^t) / (1 - beta1^t)     (mi.ad.Adam.step)
//   p   <- clamp(p, lo, hi)                                          python/variables.py:114-117 (lo > hi: off)
// ---------------------------------------------------------------------------------------------------------------
struct AdamParams {
    float lr_t, beta1, beta2, omb1, omb2, eps, grad_clamp, lo, hi;
};
__global__ void k_adam_step(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n, AdamParams a) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float gi = g[i];
        if (a.grad_clamp > 0.f) {
            gi = isnan(gi) ? 0.f : fminf(fmaxf(gi, -a.grad_clamp), a.grad_clamp);
            g[i] = gi;
        }
        const float mi = a.beta1 * m[i] + a.omb1 * gi;
        const float vi = a.beta2 * v[i] + a.omb2 * gi * gi;
        m[i] = mi;
        v[i] = vi;
        float pi = fmaf(-a.lr_t, mi / (sqrtf(vi) + a.eps), p[i]);
        if (a.lo <= a.hi) pi = fminf(fmaxf(pi, a.lo), a.hi);
        p[i] = pi;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Discrete Laplacian regulariser (python/regularizations.py:5-25): R = sum_i r_i^2, r_i = c_i - mean of the 6 clamped
// axis neighbours.  k_laplacian_residual writes r and accumulates R; k_laplacian_grad adds weight * dR/dc, which by the
// symmetry of clamped neighbour access is 2 r_i - (1/3) * (sum of the 6 clamped neighbours of r)_i.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float lap_sum6(const float* __restrict__ a, int x, int y, int z, int rx, int ry, int rz) {
    const size_t sy = (size_t)rx, sz = (size_t)rx * ry;
    const size_t i = (size_t)z * sz + (size_t)y * sy + (size_t)x;
    const float vx0 = a[x > 0 ? i - 1 : i], vx1 = a[x < rx - 1 ? i + 1 : i];
    const float vy0 = a[y > 0 ? i - sy : i], vy1 = a[y < ry - 1 ? i + sy : i];
    const float vz0 = a[z > 0 ? i - sz : i], vz1 = a[z < rz - 1 ? i + sz : i];
    return vx0 + vx1 + vy0 + vy1 + vz0 + vz1;  // the reference's summation order
}
__global__ void __launch_bounds__(256) k_laplacian_residual(const float* __restrict__ c, float* __restrict__ r, int rx, int ry, int rz,
                                                           double weight, double* __restrict__ loss /* optional */) {
    __shared__ double s_part[8];
    const size_t n = (size_t)rx * ry * rz;
    double acc = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % (size_t)rx), y = (int)((i / (size_t)rx) % (size_t)ry), z = (int)(i / ((size_t)rx * ry));
        const float ri = c[i] - lap_sum6(c, x, y, z, rx, ry, rz) / 6.f;
        r[i] = ri;
        acc += (double)ri * (double)ri;
    }
    if (!loss) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_part[w];
        atomicAdd(loss, weight * t);
    }
}
__global__ void k_laplacian_grad(const float* __restrict__ r, float* __restrict__ grad, int rx, int ry, int rz, float weight) {
    const size_t n = (size_t)rx * ry * rz;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % (size_t)rx), y = (int)((i / (size_t)rx) % (size_t)ry), z = (int)(i / ((size_t)rx * ry));
        grad[i] += weight * (2.f * r[i] - lap_sum6(r, x, y, z, rx, ry, rz) / 3.f);
    }
}

#endif  // __CUDACC__

}  // namespace sdfrt
