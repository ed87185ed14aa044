// Batched segment-vs-mesh and ray-vs-mesh tests: the B200-native replacement of the Trimesh ray
// casts the reference's planner performs one at a time on the host:
//   line_segment_mesh_intersection   /root/reference/macarons/utility/macarons_utils.py:120-151
//       -> "does the straight move start -> end cross any triangle?"   used by the Dijkstra neighbour expansion
//          (next_best_path/utility/long_term_utils.py:347) and the path check (next_best_path/testers/nbp_planning.py:142,245)
//   check_camera_in_mesh             next_best_path/utility/long_term_utils.py:158-170
//       -> parity of the hit counts of three axis-aligned rays
// One WARP per segment / ray; lanes stride over the triangles of the segment's scene (packed meshes as in raster.cu) and
// ballot their hits.  fp64, pinned op order (file compiled with -fmad=false), the formulation Trimesh's ray_triangle module
// uses: intersect the ray with the triangle's plane, take the barycentric coordinates of the intersection point (Cramer),
// accept when all lie in [-tol, 1+tol] and the point is forward of the origin; a segment additionally requires
// |location - start| < |end - start| (macarons_utils.py:143-144).  PARITY UNPINNED against Trimesh itself (trimesh 4.1.2 is not
// in the reference tree nor installable here): the pin is oracle/oracle.py::segment_mesh_hits, which this kernel matches
// bit for bit.
#include "nbp_common.cuh"

namespace nbp {

static constexpr double COL_TOL_ZERO = 1e-13;      // trimesh.constants.tol.zero (float64 resolution * 100)
static constexpr double COL_FORWARD = -1e-6;       // ray_triangle_id keeps hits with distance > -1e-6

struct D3 { double x, y, z; };
__device__ __forceinline__ D3 dsub(D3 a, D3 b) { return {__dsub_rn(a.x, b.x), __dsub_rn(a.y, b.y), __dsub_rn(a.z, b.z)}; }
__device__ __forceinline__ double ddot(D3 a, D3 b) { return __dadd_rn(__dadd_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)), __dmul_rn(a.z, b.z)); }
__device__ __forceinline__ D3 dcross(D3 a, D3 b) {
    return {__dsub_rn(__dmul_rn(a.y, b.z), __dmul_rn(a.z, b.y)), __dsub_rn(__dmul_rn(a.z, b.x), __dmul_rn(a.x, b.z)),
            __dsub_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x))};
}

// returns true and the distance along the (unit) direction if the ray origin + t*dir hits triangle (v0, v1, v2)
__device__ __forceinline__ bool ray_triangle(D3 o, D3 dir, D3 v0, D3 v1, D3 v2, double& t_out) {
    const D3 e1 = dsub(v1, v0), e2 = dsub(v2, v0);
    const D3 n = dcross(e1, e2);                                   // plane normal (not normalised: the test is scale-free)
    const double nn = ddot(n, n);
    if (!(nn > 0.0)) return false;                                 // degenerate triangle
    const double denom = ddot(n, dir);
    if (!(fabs(denom) > COL_TOL_ZERO * sqrt(nn))) return false;    // ray parallel to the plane
    const double t = __ddiv_rn(ddot(n, dsub(v0, o)), denom);
    if (!(t > COL_FORWARD)) return false;                          // behind the origin
    const D3 p = {__dadd_rn(o.x, __dmul_rn(t, dir.x)), __dadd_rn(o.y, __dmul_rn(t, dir.y)), __dadd_rn(o.z, __dmul_rn(t, dir.z))};
    // barycentric coordinates of p (Cramer, as trimesh.triangles.points_to_barycentric): w = p - v0
    const D3 w = dsub(p, v0);
    const double d00 = ddot(e1, e1), d01 = ddot(e1, e2), d11 = ddot(e2, e2), d20 = ddot(w, e1), d21 = ddot(w, e2);
    const double inv = __ddiv_rn(1.0, __dsub_rn(__dmul_rn(d00, d11), __dmul_rn(d01, d01)));
    const double b1 = __dmul_rn(__dsub_rn(__dmul_rn(d11, d20), __dmul_rn(d01, d21)), inv);
    const double b2 = __dmul_rn(__dsub_rn(__dmul_rn(d00, d21), __dmul_rn(d01, d20)), inv);
    const double b0 = __dsub_rn(__dsub_rn(1.0, b1), b2);
    const double lo = -COL_TOL_ZERO, hi = 1.0 + COL_TOL_ZERO;
    if (!(b0 > lo && b1 > lo && b2 > lo && b0 < hi && b1 < hi && b2 < hi)) return false;
    t_out = t;
    return true;
}

struct ColParams {
    const float* verts; const int32_t* faces; const int64_t* vert_off; const int64_t* face_off;
    const float* seg;            // [n][6]: start xyz, end xyz (segments)  |  origin xyz, direction xyz (rays)
    const int32_t* seg_scene; int n; int rays;
    uint8_t* hit; int32_t* count;
};

__global__ void __launch_bounds__(256) segments_mesh_kernel(ColParams p) {
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp_global >= p.n) return;
    const float* s = p.seg + 6 * (size_t)warp_global;
    const D3 o = {(double)s[0], (double)s[1], (double)s[2]};
    D3 d = {(double)s[3], (double)s[4], (double)s[5]};
    double length = 0.0;
    if (!p.rays) {
        d = dsub(d, o);
        length = sqrt(ddot(d, d));                               // np.linalg.norm(direction)
        d = {__ddiv_rn(d.x, length), __ddiv_rn(d.y, length), __ddiv_rn(d.z, length)};
    }
    const int sc = p.seg_scene[warp_global];
    const float* vb = p.verts + 3 * p.vert_off[sc];
    const int32_t* fb = p.faces + 3 * p.face_off[sc];
    const int nf = (int)(p.face_off[sc + 1] - p.face_off[sc]);
    int cnt = 0;
    bool any = false;
    for (int f0 = 0; f0 < nf; f0 += 32) {
        const int f = f0 + lane;
        bool h = false;
        if (f < nf) {
            const int i0 = fb[3 * f], i1 = fb[3 * f + 1], i2 = fb[3 * f + 2];
            const D3 v0 = {(double)vb[3 * i0], (double)vb[3 * i0 + 1], (double)vb[3 * i0 + 2]};
            const D3 v1 = {(double)vb[3 * i1], (double)vb[3 * i1 + 1], (double)vb[3 * i1 + 2]};
            const D3 v2 = {(double)vb[3 * i2], (double)vb[3 * i2 + 1], (double)vb[3 * i2 + 2]};
            double t;
            if (ray_triangle(o, d, v0, v1, v2, t)) {
                if (p.rays) h = true;
                else {
                    // distances = norm(location - start) < line_length, location = origin + t * direction (macarons_utils.py:143-144)
                    const D3 loc = {__dadd_rn(o.x, __dmul_rn(t, d.x)), __dadd_rn(o.y, __dmul_rn(t, d.y)), __dadd_rn(o.z, __dmul_rn(t, d.z))};
                    const D3 dl = dsub(loc, o);
                    h = sqrt(ddot(dl, dl)) < length;
                }
            }
        }
        const unsigned ball = __ballot_sync(0xffffffffu, h);
        cnt += __popc(ball);
        any |= ball != 0;
        if (any && !p.count) break;                              // any-hit is enough unless the caller wants the count
    }
    if (lane == 0) {
        if (p.hit) p.hit[warp_global] = any ? 1 : 0;
        if (p.count) p.count[warp_global] = cnt;
    }
}

}  // namespace nbp

using namespace nbp;

extern "C" int nbp_segments_hit_mesh(const float* verts, const int32_t* faces, const int64_t* vert_offsets, const int64_t* face_offsets,
                                     int n_scenes, const float* segments, const int32_t* seg_scene, int n_segments, int rays,
                                     uint8_t* hit, int32_t* count, void* stream) {
    if (n_segments == 0) return NBP_OK;
    if (!verts || !faces || !vert_offsets || !face_offsets || !segments || !seg_scene || (!hit && !count))
        return invalid("nbp_segments_hit_mesh: null pointer argument");
    if (n_segments < 0 || n_scenes <= 0) return invalid("nbp_segments_hit_mesh: bad sizes n_segments=%d n_scenes=%d", n_segments, n_scenes);
    ColParams p{verts, faces, vert_offsets, face_offsets, segments, seg_scene, n_segments, rays ? 1 : 0, hit, count};
    const int warps_per_block = 8;
    segments_mesh_kernel<<<(n_segments + warps_per_block - 1) / warps_per_block, 256, 0, (cudaStream_t)stream>>>(p);
    count_launch();
    return check_cuda(cudaGetLastError(), "nbp_segments_hit_mesh launch");
}
