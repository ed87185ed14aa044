/* oracle/section_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the ground-truth obstacle map of the reference's data collection,
 * get_binary_obstacle_array (next_best_path/utility/utils.py:226-262, called at nbp_utils.py:638):
 *   trimesh.intersections.mesh_plane(mesh, normal (0,1,0), origin (0, y_cam, 0))  -> 3-D line segments
 *   matplotlib: every segment drawn as a black line (x against z) in an 80 x 80 window centred on the camera,
 *   saved as a PNG, resized to 256 x 256, flipped left-right, thresholded: dark pixel -> 1.
 *
 * PARITY UNPINNED: trimesh 4.1.2 (environment.yml:344) and matplotlib are neither vendored in the reference nor
 * installed in this image, and the reference holds no fixture for this function.  What is restated:
 *   - the section: trimesh's per-face case analysis on the signed vertex distances d_i = y_i - y_cam with |d| < 1e-8
 *     counted as on-plane: edge on the plane -> that edge; one vertex on the plane and the other two on opposite sides ->
 *     vertex to the crossing of the opposite edge; one vertex on one side and two on the other -> the two edge crossings;
 *     faces that only touch the plane or lie in it produce nothing.  Crossing of edge (a, b): a + d_a / (d_a - d_b) * (b - a).
 *   - the picture: image row r, column c (after the left-right flip) has its centre at
 *         z = z_cam + V/2 - (r + 0.5) * V/S,    x = x_cam + V/2 - (c + 0.5) * V/S          (V = 80, S = 256)
 *     i.e. rows run towards -z and columns towards -x, like the egocentric count grids (utils.py:166-223).  A pixel is 1 iff
 *     its centre lies within `half_width` pixels of a segment (round caps).  matplotlib's default 1.5 pt line at 100 dpi is
 *     2.08 px in the ~198 px wide saved axes, 2.7 px after the resize to 256: half_width = 1.35.  (Deviations from the real
 *     pipeline that this restatement does not model: Agg anti-aliasing + Lanczos resampling at the line border, and the
 *     0.6 % widening of the x range that `set_aspect('equal', adjustable='datalim')` applies to the 198 x 197 px axes.)
 * All arithmetic is float32 in one fixed order (compile with -ffp-contract=off): the CUDA kernel is bit-identical.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static void mark_segment(float u0, float v0, float u1, float v1, float hw, int S, float* out) {
    /* (u, v) = continuous (column, row) image coordinates; pixel (c, r) has its centre at (c + 0.5, r + 0.5) */
    const float du = u1 - u0, dv = v1 - v0;
    const float len2 = du * du + dv * dv;
    const float hw2 = hw * hw;
    float lo_u = fminf(u0, u1) - hw, hi_u = fmaxf(u0, u1) + hw, lo_v = fminf(v0, v1) - hw, hi_v = fmaxf(v0, v1) + hw;
    int c0 = (int)floorf(lo_u - 0.5f), c1 = (int)ceilf(hi_u - 0.5f), r0 = (int)floorf(lo_v - 0.5f), r1 = (int)ceilf(hi_v - 0.5f);
    if (c0 < 0) c0 = 0;
    if (r0 < 0) r0 = 0;
    if (c1 > S - 1) c1 = S - 1;
    if (r1 > S - 1) r1 = S - 1;
    if (c0 > c1 || r0 > r1) return;
    const int major_u = fabsf(du) >= fabsf(dv);
    /* walk the major axis; per major pixel only the minor pixels the capsule can reach are tested (exact test below) */
    const int m0 = major_u ? c0 : r0, m1 = major_u ? c1 : r1;
    const float a0 = major_u ? u0 : v0, b0 = major_u ? v0 : u0, da = major_u ? du : dv, db = major_u ? dv : du;
    const float slope = (da != 0.0f) ? db / da : 0.0f;
    const float reach = hw * 1.5f + 1.0f;                       /* >= hw * sqrt(1 + slope^2) + half a pixel, |slope| <= 1 */
    const float a_lo = fminf(a0, a0 + da), a_hi = fmaxf(a0, a0 + da);
    for (int m = m0; m <= m1; ++m) {
        float am = (float)m + 0.5f;
        if (am < a_lo) am = a_lo;
        if (am > a_hi) am = a_hi;
        const float bm = b0 + (am - a0) * slope;                 /* minor coordinate of the line at this major position */
        int n0 = (int)floorf(bm - reach - 0.5f), n1 = (int)ceilf(bm + reach - 0.5f);
        const int nlo = major_u ? r0 : c0, nhi = major_u ? r1 : c1;
        if (n0 < nlo) n0 = nlo;
        if (n1 > nhi) n1 = nhi;
        for (int n = n0; n <= n1; ++n) {
            const int c = major_u ? m : n, r = major_u ? n : m;
            const float pu = ((float)c + 0.5f) - u0, pv = ((float)r + 0.5f) - v0;
            float t = 0.0f;
            if (len2 > 0.0f) {
                t = (pu * du + pv * dv) / len2;
                if (t < 0.0f) t = 0.0f;
                if (t > 1.0f) t = 1.0f;
            }
            const float eu = pu - t * du, ev = pv - t * dv;
            if (eu * eu + ev * ev <= hw2) out[(size_t)r * S + c] = 1.0f;
        }
    }
}

/* returns the number of section segments; out (S*S) is zeroed here */
int nbp_oracle_plane_section_map(const float* verts, const int64_t* faces, int64_t n_faces, const float* pose /* x,y,z,.. */,
                                 int S, float view, float half_width, float* out, float* segments /* optional [n_faces][4] x0,z0,x1,z1 */) {
    const float tol = 1e-8f;
    const float cx = pose[0], y0 = pose[1], cz = pose[2];
    const float half = view * 0.5f, scale = (float)S / view;
    int n_seg = 0;
    memset(out, 0, sizeof(float) * (size_t)S * S);
    for (int64_t f = 0; f < n_faces; ++f) {
        float p[3][3], d[3];
        int s[3], nz = 0, npos = 0, nneg = 0;
        for (int k = 0; k < 3; ++k) {
            const float* v = verts + 3 * faces[3 * f + k];
            p[k][0] = v[0]; p[k][1] = v[1]; p[k][2] = v[2];
            d[k] = v[1] - y0;
            s[k] = fabsf(d[k]) < tol ? 0 : (d[k] > 0.0f ? 1 : -1);
            nz += s[k] == 0; npos += s[k] > 0; nneg += s[k] < 0;
        }
        float q[2][2];
        int nq = 0;
        if (nz == 2 && (npos + nneg) == 1) {                          /* an edge lies in the plane */
            for (int k = 0; k < 3; ++k) if (s[k] == 0) { q[nq][0] = p[k][0]; q[nq][1] = p[k][2]; ++nq; }
        } else if (nz <= 1 && npos >= 1 && nneg >= 1) {               /* the plane cuts through the face */
            for (int k = 0; k < 3 && nq < 2; ++k) if (s[k] == 0) { q[nq][0] = p[k][0]; q[nq][1] = p[k][2]; ++nq; }
            for (int k = 0; k < 3 && nq < 2; ++k) {
                const int a = k, b = (k + 1) % 3;
                if (s[a] * s[b] < 0) {
                    const float t = d[a] / (d[a] - d[b]);
                    q[nq][0] = p[a][0] + t * (p[b][0] - p[a][0]);
                    q[nq][1] = p[a][2] + t * (p[b][2] - p[a][2]);
                    ++nq;
                }
            }
        }
        if (nq != 2) continue;
        if (segments) { segments[4 * n_seg] = q[0][0]; segments[4 * n_seg + 1] = q[0][1]; segments[4 * n_seg + 2] = q[1][0]; segments[4 * n_seg + 3] = q[1][1]; }
        ++n_seg;
        const float u0 = ((cx + half) - q[0][0]) * scale, v0 = ((cz + half) - q[0][1]) * scale;
        const float u1 = ((cx + half) - q[1][0]) * scale, v1 = ((cz + half) - q[1][1]) * scale;
        mark_segment(u0, v0, u1, v1, half_width, S, out);
    }
    return n_seg;
}
