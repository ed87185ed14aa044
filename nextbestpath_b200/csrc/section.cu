// SURVEY.md section 8(f) row 3, second half -- the ground-truth obstacle map of the reference's data collection:
// get_binary_obstacle_array (/root/reference/next_best_path/utility/utils.py:226-262, called at nbp_utils.py:638) cuts the mesh with
// the horizontal plane through the camera (trimesh.intersections.mesh_plane), draws the section with matplotlib into an 80 x 80
// window, resizes the PNG to 256 x 256, flips it and thresholds it -- per pose, on the host.  Here: one launch for all scenes,
// one thread per (map, face): trimesh's per-face case analysis on the signed vertex distances, then the segment is stamped into
// the map as a capsule of the line's half width (idempotent 1.0f stores, no atomics).  Arithmetic pinned op-for-op (this file is
// compiled with -fmad=false) to oracle/section_oracle.c, which documents the restated picture geometry: bit-identical maps.
#include "nbp_common.cuh"

namespace nbp {

__device__ void mark_segment(float u0, float v0, float u1, float v1, float hw, int S, float* __restrict__ out) {
    const float du = fsub(u1, u0), dv = fsub(v1, v0);
    const float len2 = fadd(fmul(du, du), fmul(dv, dv));
    const float hw2 = fmul(hw, hw);
    const float lo_u = fsub(fminf(u0, u1), hw), hi_u = fadd(fmaxf(u0, u1), hw), lo_v = fsub(fminf(v0, v1), hw), hi_v = fadd(fmaxf(v0, v1), hw);
    int c0 = (int)floorf(fsub(lo_u, 0.5f)), c1 = (int)ceilf(fsub(hi_u, 0.5f)), r0 = (int)floorf(fsub(lo_v, 0.5f)), r1 = (int)ceilf(fsub(hi_v, 0.5f));
    c0 = max(c0, 0); r0 = max(r0, 0); c1 = min(c1, S - 1); r1 = min(r1, S - 1);
    if (c0 > c1 || r0 > r1) return;
    const bool major_u = fabsf(du) >= fabsf(dv);
    const int m0 = major_u ? c0 : r0, m1 = major_u ? c1 : r1;
    const float a0 = major_u ? u0 : v0, b0 = major_u ? v0 : u0, da = major_u ? du : dv, db = major_u ? dv : du;
    const float slope = (da != 0.0f) ? __fdiv_rn(db, da) : 0.0f;
    const float reach = fadd(fmul(hw, 1.5f), 1.0f);
    const float a_end = fadd(a0, da);
    const float a_lo = fminf(a0, a_end), a_hi = fmaxf(a0, a_end);
    const int nlo = major_u ? r0 : c0, nhi = major_u ? r1 : c1;
    for (int m = m0; m <= m1; ++m) {
        float am = fadd((float)m, 0.5f);
        am = fminf(fmaxf(am, a_lo), a_hi);
        const float bm = fadd(b0, fmul(fsub(am, a0), slope));
        const int n0 = max((int)floorf(fsub(fsub(bm, reach), 0.5f)), nlo), n1 = min((int)ceilf(fsub(fadd(bm, reach), 0.5f)), nhi);
        for (int n = n0; n <= n1; ++n) {
            const int c = major_u ? m : n, r = major_u ? n : m;
            const float pu = fsub(fadd((float)c, 0.5f), u0), pv = fsub(fadd((float)r, 0.5f), v0);
            float t = 0.0f;
            if (len2 > 0.0f) {
                t = __fdiv_rn(fadd(fmul(pu, du), fmul(pv, dv)), len2);
                t = fminf(fmaxf(t, 0.0f), 1.0f);
            }
            const float eu = fsub(pu, fmul(t, du)), ev = fsub(pv, fmul(t, dv));
            if (fadd(fmul(eu, eu), fmul(ev, ev)) <= hw2) out[(size_t)r * S + c] = 1.0f;
        }
    }
}

__global__ void __launch_bounds__(128) plane_section_kernel(const float* __restrict__ verts, const int32_t* __restrict__ faces,
                                                            const int64_t* __restrict__ vert_off, const int64_t* __restrict__ face_off,
                                                            const int32_t* __restrict__ map_scene, const float* __restrict__ pose, int S,
                                                            float view, float hw, float* __restrict__ out) {
    const int b = blockIdx.y;
    const int sc = map_scene ? map_scene[b] : b;
    const int64_t f0 = face_off[sc], nf = face_off[sc + 1] - f0;
    const float* V = verts + 3 * vert_off[sc];
    const float cx = pose[5 * b], y0 = pose[5 * b + 1], cz = pose[5 * b + 2];
    const float half = fmul(view, 0.5f), scale = __fdiv_rn((float)S, view);
    float* o = out + (size_t)b * S * S;
    for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += (int64_t)gridDim.x * blockDim.x) {
        float p[3][3], d[3];
        int s[3], nz = 0, npos = 0, nneg = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float* v = V + 3 * (int64_t)faces[3 * (f0 + f) + k];
            p[k][0] = v[0]; p[k][1] = v[1]; p[k][2] = v[2];
            d[k] = fsub(v[1], y0);
            s[k] = fabsf(d[k]) < 1e-8f ? 0 : (d[k] > 0.0f ? 1 : -1);
            nz += s[k] == 0; npos += s[k] > 0; nneg += s[k] < 0;
        }
        float q[2][2];
        int nq = 0;
        if (nz == 2 && (npos + nneg) == 1) {
#pragma unroll
            for (int k = 0; k < 3; ++k) if (s[k] == 0) { q[nq][0] = p[k][0]; q[nq][1] = p[k][2]; ++nq; }
        } else if (nz <= 1 && npos >= 1 && nneg >= 1) {
#pragma unroll
            for (int k = 0; k < 3; ++k) if (nq < 2 && s[k] == 0) { q[nq][0] = p[k][0]; q[nq][1] = p[k][2]; ++nq; }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int a = k, bb = (k + 1) % 3;
                if (nq < 2 && s[a] * s[bb] < 0) {
                    const float t = __fdiv_rn(d[a], fsub(d[a], d[bb]));
                    q[nq][0] = fadd(p[a][0], fmul(t, fsub(p[bb][0], p[a][0])));
                    q[nq][1] = fadd(p[a][2], fmul(t, fsub(p[bb][2], p[a][2])));
                    ++nq;
                }
            }
        }
        if (nq != 2) continue;
        const float u0 = fmul(fsub(fadd(cx, half), q[0][0]), scale), v0 = fmul(fsub(fadd(cz, half), q[0][1]), scale);
        const float u1 = fmul(fsub(fadd(cx, half), q[1][0]), scale), v1 = fmul(fsub(fadd(cz, half), q[1][1]), scale);
        mark_segment(u0, v0, u1, v1, hw, S, o);
    }
}

}  // namespace nbp

using namespace nbp;

extern "C" int nbp_gt_obstacle_map(const float* verts, const int32_t* faces, const int64_t* vert_offsets, const int64_t* face_offsets,
                                   const int32_t* map_scene, const float* pose, int n_maps, int64_t max_faces_per_scene, int S,
                                   float view_size, float half_width_px, float* out, void* stream) {
    if (!verts || !faces || !vert_offsets || !face_offsets || !pose || !out) return invalid("nbp_gt_obstacle_map: null pointer argument");
    if (n_maps <= 0 || S <= 0 || S > 4096 || !(view_size > 0.0f) || !(half_width_px > 0.0f) || max_faces_per_scene < 0)
        return invalid("nbp_gt_obstacle_map: bad sizes n_maps=%d S=%d view=%f half_width=%f", n_maps, S, view_size, half_width_px);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_cuda(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)n_maps * S * S, st), "nbp_gt_obstacle_map memset");
    if (rc) return rc;
    if (max_faces_per_scene == 0) return NBP_OK;
    int gx = (int)((max_faces_per_scene + 127) / 128);
    if (gx > 148 * 8) gx = 148 * 8;
    plane_section_kernel<<<dim3(gx, n_maps), 128, 0, st>>>(verts, faces, vert_offsets, face_offsets, map_scene, pose, S, view_size, half_width_px, out);
    count_launch();
    return check_cuda(cudaGetLastError(), "nbp_gt_obstacle_map launch");
}
