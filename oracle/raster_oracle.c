/*
 * oracle/raster_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, strict fp32, no FMA contraction) of the depth
 * rasterisation the reference obtains from PyTorch3D 0.7.4 at
 *   /root/reference/macarons/utility/macarons_utils.py:2759  (renderer(mesh, cameras=fov_camera))
 *   /root/reference/macarons/utility/macarons_utils.py:905-937 (RasterizationSettings: blur 0, K=1)
 * PyTorch3D (pinned `pytorch3d==0.7.4`, environment.yml:204) is NOT vendored in
 * /root/reference and not installable here, so this follows its published
 * algorithm (renderer/mesh/rasterizer.py MeshRasterizer.transform,
 * renderer/mesh/clip.py clip_faces, csrc/rasterize_meshes/rasterize_meshes_cpu.cpp
 * RasterizeMeshesNaiveCpu, renderer/cameras.py FoVPerspectiveCameras):
 *
 *   1. verts_view = verts_world * R + T          (row vectors)
 *   2. x_ndc = (x_v * f) / z_v ; y_ndc = (y_v * f) / z_v ; z kept as view z
 *      (f = 1/tan(fov/2), znear 1, aspect 1 -- FoVPerspectiveCameras defaults)
 *   3. faces are clipped against z = z_clip (= znear/2 = 0.5): fully-behind
 *      faces are dropped, one-vertex-behind faces become two triangles,
 *      two-vertex-behind faces become one; intersection points are found in
 *      (x_ndc*z, y_ndc*z, z) space (perspective_correct=True)
 *   4. naive per-pixel loop: pixel (yi,xi) <-> NDC via NonSquarePixToNdc with
 *      the y/x flip, bbox reject, |area|<=1e-8 reject, barycentrics with
 *      area+1e-8, perspective correction with max(denom,1e-8), pz<0 reject,
 *      strict >0 inside test, keep min (pz, face).
 *
 * PARITY UNPINNED by the reference itself: the reference has no test, fixture
 * or golden vector touching rasterisation, and
 * PyTorch3D cannot be run here.  This file fixes ONE fp32 evaluation order; the
 * CUDA rasteriser is compared bit-for-bit against it.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may call this.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define K_EPS 1e-8f

typedef struct { float x, y, z; } v3;

static inline float edge_fn(float px, float py, float ax, float ay, float bx, float by) {
    /* EdgeFunctionForward(p, v0=a, v1=b) */
    float t0 = (px - ax) * (by - ay);
    float t1 = (py - ay) * (bx - ax);
    return t0 - t1;
}

/* NonSquarePixToNdc(i, S1, S2) */
static inline float pix_to_ndc(int i, int S1, int S2) {
    float range = (S1 > S2) ? (2.0f * (float)S1) / (float)S2 : 2.0f;
    float offset = range / 2.0f;
    return -offset + (range * (float)i + offset) / (float)S1;
}

/* world -> (x_ndc, y_ndc, z_view) for one vertex; R row-major 3x3, row-vector convention */
static inline v3 project_vertex(const float *p, const float *R, const float *T, float focal) {
    float xv = ((p[0] * R[0] + p[1] * R[3]) + p[2] * R[6]) + T[0];
    float yv = ((p[0] * R[1] + p[1] * R[4]) + p[2] * R[7]) + T[1];
    float zv = ((p[0] * R[2] + p[1] * R[5]) + p[2] * R[8]) + T[2];
    v3 o;
    o.x = (xv * focal) / zv;
    o.y = (yv * focal) / zv;
    o.z = zv;
    return o;
}

/* intersection of edge p1->p2 with the plane z = clip, perspective-correct */
static inline v3 clip_edge(v3 p1, v3 p2, float clip) {
    float w = (p1.z - clip) / (p1.z - p2.z);
    float omw = 1.0f - w;
    float p1wx = p1.x * p1.z, p1wy = p1.y * p1.z;
    float p2wx = p2.x * p2.z, p2wy = p2.y * p2.z;
    v3 o;
    o.x = (p1wx * omw + p2wx * w) / clip;
    o.y = (p1wy * omw + p2wy * w) / clip;
    o.z = clip;
    return o;
}

/*
 * Clip one projected face.  Writes 0, 1 or 2 triangles (9 floats each:
 * x0 y0 z0 x1 y1 z1 x2 y2 z2) to out; returns the count.
 */
static int clip_face(const v3 v[3], float clip, float *out) {
    int behind[3], nb = 0;
    for (int i = 0; i < 3; ++i) { behind[i] = v[i].z < clip; nb += behind[i]; }
    if (nb == 3) return 0;
    if (nb == 0) {
        for (int i = 0; i < 3; ++i) { out[3*i] = v[i].x; out[3*i+1] = v[i].y; out[3*i+2] = v[i].z; }
        return 1;
    }
    if (nb == 1) {
        /* p1 = the vertex behind; p2, p3 follow cyclically. quad p4 p2 p3 p5 -> (p4,p2,p5),(p5,p2,p3) */
        int i1 = behind[0] ? 0 : (behind[1] ? 1 : 2);
        v3 p1 = v[i1], p2 = v[(i1 + 1) % 3], p3 = v[(i1 + 2) % 3];
        v3 p4 = clip_edge(p1, p2, clip), p5 = clip_edge(p1, p3, clip);
        v3 t[6] = { p4, p2, p5, p5, p2, p3 };
        for (int i = 0; i < 6; ++i) { out[3*i] = t[i].x; out[3*i+1] = t[i].y; out[3*i+2] = t[i].z; }
        return 2;
    }
    /* nb == 2: p1 = the vertex in front; triangle (p1,p4,p5) */
    int i1 = !behind[0] ? 0 : (!behind[1] ? 1 : 2);
    v3 p1 = v[i1], p2 = v[(i1 + 1) % 3], p3 = v[(i1 + 2) % 3];
    v3 p4 = clip_edge(p1, p2, clip), p5 = clip_edge(p1, p3, clip);
    v3 t[3] = { p1, p4, p5 };
    for (int i = 0; i < 3; ++i) { out[3*i] = t[i].x; out[3*i+1] = t[i].y; out[3*i+2] = t[i].z; }
    return 1;
}

/*
 * Stage 1+2+3: project and clip all faces of one mesh for one camera.
 * tris: caller buffer of 2*F*9 floats; tri_face: 2*F int32 (source face of each
 * clipped triangle).  Returns the number of triangles written.
 */
int nbp_oracle_project_clip(const float *verts, const int64_t *faces, int64_t F,
                            const float *R, const float *T, float focal, float z_clip,
                            float *tris, int32_t *tri_face) {
    int64_t n = 0;
    for (int64_t f = 0; f < F; ++f) {
        v3 v[3];
        for (int k = 0; k < 3; ++k) v[k] = project_vertex(verts + 3 * faces[3 * f + k], R, T, focal);
        int c = clip_face(v, z_clip, tris + 9 * n);
        for (int k = 0; k < c; ++k) tri_face[n + k] = (int32_t)f;
        n += c;
    }
    return (int)n;
}

/*
 * Stage 4: naive rasterisation of n clipped triangles into an H x W zbuf
 * (view-space z, -1 = no face) and pix_to_face (source face index, -1 = none).
 * Ties on z keep the lowest source face index.  Rows are independent, so they
 * may be split over `nthreads` POSIX threads; that only changes wall time.
 */
typedef struct {
    const float *tris; const int32_t *tri_face; const float *bb; const float *area;
    int n, H, W, y0, y1; float *zbuf; int32_t *pix_to_face;
} row_job;

static void *raster_rows(void *arg) {
    const row_job *j = (const row_job *)arg;
    const float *tris = j->tris, *bb = j->bb, *area = j->area;
    const int n = j->n, H = j->H, W = j->W;
    for (int yi = j->y0; yi < j->y1; ++yi) {
        const float yf = pix_to_ndc(H - 1 - yi, H, W);
        for (int xi = 0; xi < W; ++xi) {
            const float xf = pix_to_ndc(W - 1 - xi, W, H);
            float best_z = -1.0f; int32_t best_f = -1;
            for (int t = 0; t < n; ++t) {
                if (xf > bb[4*t+1] || xf < bb[4*t+0] || yf > bb[4*t+3] || yf < bb[4*t+2]) continue;
                const float a = area[t];
                if (a <= K_EPS && a >= -K_EPS) continue;
                const float *q = tris + 9 * t;
                const float den = a + K_EPS;
                const float w0 = edge_fn(xf, yf, q[3], q[4], q[6], q[7]) / den;
                const float w1 = edge_fn(xf, yf, q[6], q[7], q[0], q[1]) / den;
                const float w2 = edge_fn(xf, yf, q[0], q[1], q[3], q[4]) / den;
                const float z0 = q[2], z1 = q[5], z2 = q[8];
                const float t0 = (w0 * z1) * z2;
                const float t1 = (z0 * w1) * z2;
                const float t2 = (z0 * z1) * w2;
                const float d = fmaxf((t0 + t1) + t2, K_EPS);
                const float l0 = t0 / d, l1 = t1 / d, l2 = t2 / d;
                const float pz = (l0 * z0 + l1 * z1) + l2 * z2;
                if (pz < 0.0f) continue;
                if (!(l0 > 0.0f && l1 > 0.0f && l2 > 0.0f)) continue;
                const int32_t f = j->tri_face[t];
                if (best_f < 0 || pz < best_z || (pz == best_z && f < best_f)) { best_z = pz; best_f = f; }
            }
            j->zbuf[(size_t)yi * W + xi] = best_z;
            j->pix_to_face[(size_t)yi * W + xi] = best_f;
        }
    }
    return NULL;
}

void nbp_oracle_raster_naive(const float *tris, const int32_t *tri_face, int n,
                             int H, int W, float *zbuf, int32_t *pix_to_face, int nthreads) {
    float *bb = (float *)malloc(sizeof(float) * 4 * (size_t)(n > 0 ? n : 1));
    float *area = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    for (int t = 0; t < n; ++t) {
        const float *q = tris + 9 * t;
        bb[4*t+0] = fminf(fminf(q[0], q[3]), q[6]);
        bb[4*t+1] = fmaxf(fmaxf(q[0], q[3]), q[6]);
        bb[4*t+2] = fminf(fminf(q[1], q[4]), q[7]);
        bb[4*t+3] = fmaxf(fmaxf(q[1], q[4]), q[7]);
        area[t] = edge_fn(q[6], q[7], q[0], q[1], q[3], q[4]);   /* EdgeFunction(v2, v0, v1) */
    }
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if (nthreads > H) nthreads = H;
    row_job jobs[256]; pthread_t th[256];
    for (int k = 0; k < nthreads; ++k) {
        row_job jb = { tris, tri_face, bb, area, n, H, W,
                       (int)((long)H * k / nthreads), (int)((long)H * (k + 1) / nthreads), zbuf, pix_to_face };
        jobs[k] = jb;
    }
    if (nthreads == 1) raster_rows(&jobs[0]);
    else {
        for (int k = 0; k < nthreads; ++k) pthread_create(&th[k], NULL, raster_rows, &jobs[k]);
        for (int k = 0; k < nthreads; ++k) pthread_join(th[k], NULL);
    }
    free(bb); free(area);
}

/* Convenience: full depth render of one mesh from one camera. */
int nbp_oracle_render_depth(const float *verts, int64_t V, const int64_t *faces, int64_t F,
                            const float *R, const float *T, float focal, float z_clip,
                            int H, int W, float *zbuf, int32_t *pix_to_face, int nthreads) {
    (void)V;
    float *tris = (float *)malloc(sizeof(float) * 18 * (size_t)(F > 0 ? F : 1));
    int32_t *tf = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(F > 0 ? F : 1));
    if (!tris || !tf) { free(tris); free(tf); return -1; }
    int n = nbp_oracle_project_clip(verts, faces, F, R, T, focal, z_clip, tris, tf);
    nbp_oracle_raster_naive(tris, tf, n, H, W, zbuf, pix_to_face, nthreads);
    free(tris); free(tf);
    return n;
}
