/* oracle/redistance_oracle.c -- CPU restatement of `redistancing.redistance(phi)` (python/redistancing.py:4-13), i.e. of
 * `fastsweep.redistance`, the re-distancing step python/variables.py:189 runs on the SDF grid after every optimiser step.
 *
 * TEST INFRASTRUCTURE ONLY: tests/ and bench.py's cpu_baseline leg load it as the checker / the timed CPU side; the
 * product (libsdfrt.so) never links or calls it.
 *
 * PARITY UNPINNED.  `fastsweep` (rgl-epfl/fastsweep, `pip install fastsweep`, version unpinned in README.md:48-50) is a
 * third-party native module that is neither vendored in /root/reference nor installable offline, and the reference holds no
 * test or golden vector for it.  What is restated here is its published algorithm:
 *   - the fast sweeping method for the Eikonal equation |grad u| = 1 (Zhao 2005) with the first-order Godunov upwind
 *     discretisation; fastsweep runs its parallel variant (Detrixhe, Gibou, Min 2013), which converges to the same
 *     discrete solution -- the fixed point of the upwind update is unique for given frozen values, so the sweep order
 *     only changes the iteration count, not the result;
 *   - the sub-voxel initialisation of the nodes next to the zero level set in the style of scikit-fmm (the reference's
 *     alternative back end, python/redistancing.py:8-11): along every axis with a sign change towards a neighbour the
 *     distance to the linearly interpolated crossing d = h * phi_i / (phi_i - phi_n), the smaller of the two sides per
 *     axis, combined as 1 / sqrt(sum_axes 1 / d^2); nodes with phi == 0 are frozen at 0;
 *   - domain = unit cube, node spacing h_axis = 1 / resolution_axis (python/redistancing.py:11 passes the same spacing to
 *     skfmm), sign of the input kept.
 * Storage order (z, y, x), x fastest, as the tensors of python/shapes.py:378-392.
 *
 * The solver below is the textbook sequential Gauss-Seidel form: 8 sweep orderings, repeated until a full cycle changes
 * no node.  All arithmetic is fp32 and written operation by operation (compile with -ffp-contract=off) so that the CUDA
 * kernel can follow the same sequence.
 */
#include <math.h>
#include <stdlib.h>

#define RD_INF 1.0e30f

/* first-order Godunov update: smallest u with sum_k max((u - a_k) / h_k, 0)^2 = 1; a sorted ascending with their h */
static float rd_solve(float a0, float h0, float a1, float h1, float a2, float h2) {
    float t;
    /* sort (a0,h0) <= (a1,h1) <= (a2,h2) */
    if (a1 < a0) { t = a0; a0 = a1; a1 = t; t = h0; h0 = h1; h1 = t; }
    if (a2 < a1) { t = a1; a1 = a2; a2 = t; t = h1; h1 = h2; h2 = t; }
    if (a1 < a0) { t = a0; a0 = a1; a1 = t; t = h0; h0 = h1; h1 = t; }
    if (a0 >= RD_INF) return RD_INF;
    float u = a0 + h0;
    if (u <= a1) return u;
    /* the quadratics are solved for s = u - a0 (differences of neighbouring values are O(h), so nothing cancels) */
    {   /* two terms: s^2/h0^2 + (s-b1)^2/h1^2 = 1 */
        const float b1 = a1 - a0;
        const float w0 = 1.0f / (h0 * h0), w1 = 1.0f / (h1 * h1);
        const float A = w0 + w1;
        const float B = w1 * b1;
        const float C = w1 * b1 * b1 - 1.0f;
        const float D = B * B - A * C;
        u = a0 + (B + sqrtf(D > 0.0f ? D : 0.0f)) / A;
        if (u <= a2) return u;
    }
    {   /* three terms */
        const float b1 = a1 - a0, b2 = a2 - a0;
        const float w0 = 1.0f / (h0 * h0), w1 = 1.0f / (h1 * h1), w2 = 1.0f / (h2 * h2);
        const float A = w0 + w1 + w2;
        const float B = w1 * b1 + w2 * b2;
        const float C = w1 * b1 * b1 + w2 * b2 * b2 - 1.0f;
        const float D = B * B - A * C;
        u = a0 + (B + sqrtf(D > 0.0f ? D : 0.0f)) / A;
    }
    return u;
}

/* frozen-node initialisation; returns the unsigned distance, or RD_INF for a node away from the interface */
static float rd_init_node(const float* phi, int rx, int ry, int rz, int x, int y, int z, const float h[3]) {
    const size_t sy = (size_t)rx, sz = (size_t)rx * ry;
    const size_t i = (size_t)z * sz + (size_t)y * sy + (size_t)x;
    const float p = phi[i];
    if (p == 0.0f) return 0.0f;
    float dsum = 0.0f;
    int borders = 0;
    for (int a = 0; a < 3; ++a) {
        float best = 0.0f;
        for (int s = -1; s <= 1; s += 2) {
            int xx = x, yy = y, zz = z;
            if (a == 0) xx += s; else if (a == 1) yy += s; else zz += s;
            if (xx < 0 || xx >= rx || yy < 0 || yy >= ry || zz < 0 || zz >= rz) continue;
            const float q = phi[(size_t)zz * sz + (size_t)yy * sy + (size_t)xx];
            if (p * q < 0.0f) {
                const float d = h[a] * p / (p - q);
                if (best == 0.0f || best > d) best = d;
                borders = 1;
            }
        }
        if (best > 0.0f) dsum += 1.0f / (best * best);
    }
    if (!borders) return RD_INF;
    return sqrtf(1.0f / dsum);
}

/* out = redistance(phi).  Returns the number of 8-sweep cycles run (the last one changes nothing). */
int rd_redistance(const float* phi, float* out, int rx, int ry, int rz) {
    const float h[3] = {1.0f / (float)rx, 1.0f / (float)ry, 1.0f / (float)rz};
    const size_t sy = (size_t)rx, sz = (size_t)rx * ry, n = sz * rz;
    unsigned char* frozen = (unsigned char*)malloc(n);
    float* u = out;
    for (int z = 0; z < rz; ++z)
        for (int y = 0; y < ry; ++y)
            for (int x = 0; x < rx; ++x) {
                const size_t i = (size_t)z * sz + (size_t)y * sy + (size_t)x;
                u[i] = rd_init_node(phi, rx, ry, rz, x, y, z, h);
                frozen[i] = u[i] < RD_INF;
            }
    int cycles = 0;
    for (;;) {
        int changed = 0;
        for (int dir = 0; dir < 8; ++dir) {
            const int x0 = (dir & 1) ? rx - 1 : 0, xs = (dir & 1) ? -1 : 1;
            const int y0 = (dir & 2) ? ry - 1 : 0, ys = (dir & 2) ? -1 : 1;
            const int z0 = (dir & 4) ? rz - 1 : 0, zs = (dir & 4) ? -1 : 1;
            for (int zi = 0, z = z0; zi < rz; ++zi, z += zs)
                for (int yi = 0, y = y0; yi < ry; ++yi, y += ys)
                    for (int xi = 0, x = x0; xi < rx; ++xi, x += xs) {
                        const size_t i = (size_t)z * sz + (size_t)y * sy + (size_t)x;
                        if (frozen[i]) continue;
                        const float xm = x > 0 ? u[i - 1] : RD_INF, xp = x < rx - 1 ? u[i + 1] : RD_INF;
                        const float ym = y > 0 ? u[i - sy] : RD_INF, yp = y < ry - 1 ? u[i + sy] : RD_INF;
                        const float zm = z > 0 ? u[i - sz] : RD_INF, zp = z < rz - 1 ? u[i + sz] : RD_INF;
                        const float v = rd_solve(xm < xp ? xm : xp, h[0], ym < yp ? ym : yp, h[1], zm < zp ? zm : zp, h[2]);
                        if (v < u[i]) { u[i] = v; changed = 1; }
                    }
        }
        ++cycles;
        if (!changed) break;
    }
    for (size_t i = 0; i < n; ++i) out[i] = phi[i] < 0.0f ? -u[i] : u[i];
    free(frozen);
    return cycles;
}

/* largest violation of the discrete upwind Eikonal equation over the non-frozen nodes, and of the frozen values against
 * the initialisation: the size-independent property the full-size GPU tests check.  res[0] = max |u - solve(neighbours)|,
 * res[1] = max |u_frozen - init|, res[2] = number of sign mismatches against phi. */
void rd_check(const float* phi, const float* dist, int rx, int ry, int rz, double res[3]) {
    const float h[3] = {1.0f / (float)rx, 1.0f / (float)ry, 1.0f / (float)rz};
    const size_t sy = (size_t)rx, sz = (size_t)rx * ry;
    double e_eik = 0.0, e_init = 0.0, n_sign = 0.0;
    for (int z = 0; z < rz; ++z)
        for (int y = 0; y < ry; ++y)
            for (int x = 0; x < rx; ++x) {
                const size_t i = (size_t)z * sz + (size_t)y * sy + (size_t)x;
                const float ui = fabsf(dist[i]);
                if ((phi[i] < 0.0f) != (dist[i] < 0.0f) && dist[i] != 0.0f) n_sign += 1.0;
                const float f = rd_init_node(phi, rx, ry, rz, x, y, z, h);
                if (f < RD_INF) {
                    const double e = fabs((double)ui - (double)f);
                    if (e > e_init) e_init = e;
                    continue;
                }
#define RD_ABS(j) fabsf(dist[j])
                const float xm = x > 0 ? RD_ABS(i - 1) : RD_INF, xp = x < rx - 1 ? RD_ABS(i + 1) : RD_INF;
                const float ym = y > 0 ? RD_ABS(i - sy) : RD_INF, yp = y < ry - 1 ? RD_ABS(i + sy) : RD_INF;
                const float zm = z > 0 ? RD_ABS(i - sz) : RD_INF, zp = z < rz - 1 ? RD_ABS(i + sz) : RD_INF;
#undef RD_ABS
                const float v = rd_solve(xm < xp ? xm : xp, h[0], ym < yp ? ym : yp, h[1], zm < zp ? zm : zp, h[2]);
                const double e = fabs((double)ui - (double)v);
                if (e > e_eik) e_eik = e;
            }
    res[0] = e_eik; res[1] = e_init; res[2] = n_sign;
}
