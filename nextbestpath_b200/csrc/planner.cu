// SURVEY.md section 8(f) row 1 -- the step immediately downstream of NBP.forward in the reference's re-plan branch
// (/root/reference/next_best_path/testers/nbp_planning.py:166-233; check_pixel_values macarons/utility/macarons_utils.py:86-100):
//   obstacle-map fusion   predicted map thresholded at 0.13, overwritten by the observed height-slice occupancy wherever the
//                         cloud has any point, cleared along the trajectory                               (:168-190)
//   candidate scoring     for every lattice position: heading-max of the value map at its S/4 cell, point-density penalty and
//                         the +-10 px "has the camera seen anything near here" test on the S grid         (:193-231)
// The reference runs the second part as a Python loop with several .item() syncs per candidate; here it is one launch over
// (scene, candidate).  Cell indices use the same pinned fp32 rounding as scatter.cu (bit-exact with torch).
#include "nbp_common.cuh"

namespace nbp {

__device__ __forceinline__ float cellf(float v, float lo, float scale) { return rintf(fmul(fsub(v, lo), scale)); }

__global__ void __launch_bounds__(256) obstacle_fuse_kernel(const float* __restrict__ pred, const float* __restrict__ all_cnt, int all_stride,
                                                            const float* __restrict__ slice_cnt, int slice_stride,
                                                            const float* __restrict__ traj, int traj_stride, int n, int cells,
                                                            float threshold, float* __restrict__ fused, float* __restrict__ full_proj) {
    const size_t total = (size_t)n * cells;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / cells, c = i - b * cells;
        float o = pred[i] >= threshold ? 1.0f : 0.0f;                       // (predicted_obstacle_map >= 0.13).float()
        const bool seen = all_cnt[b * all_stride + c] > 0.0f;               // mask_layout = full_pc_projection > 0
        if (seen) o = slice_cnt[b * slice_stride + c] > 0.0f ? 1.0f : 0.0f;  // <- filt_pc_selection_img
        if (traj[b * traj_stride + c] > 0.0f) o = 0.0f;                     // previous trajectory is passable
        fused[i] = o;
        full_proj[i] = seen ? 1.0f : 0.0f;                                   // full_pc_projection clipped to {0,1}
    }
}

// python-style index wrap of a (possibly negative) integer index into [0, S) -- torch indexing semantics
__device__ __forceinline__ int wrap_index(int i, int S) { return i < 0 ? i + S : i; }

__global__ void __launch_bounds__(128) candidate_score_kernel(const float* __restrict__ cand, const int32_t* __restrict__ n_cand, int max_cand,
                                                              const uint8_t* __restrict__ skip, const float* __restrict__ pose,
                                                              const float* __restrict__ value_map, int n_ch, int Sv,
                                                              const float* __restrict__ full_proj, int S, float lo, float scale_v, float scale_s,
                                                              int window, float* __restrict__ value, float* __restrict__ density,
                                                              int32_t* __restrict__ cell, uint8_t* __restrict__ valid) {
    const int b = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= max_cand) return;
    const size_t o = (size_t)b * max_cand + j;
    valid[o] = 0; value[o] = 0.0f; density[o] = 0.0f; cell[2 * o] = -1; cell[2 * o + 1] = -1;
    if (j >= n_cand[b] || (skip && skip[o])) return;
    const float cx = pose[b * 5 + 0], cz = pose[b * 5 + 2];
    const float p0 = -fsub(cand[3 * o + 2], cz), p1 = -fsub(cand[3 * o + 0], cx);          // transform_points_to_n_pieces
    const float g0 = cellf(p0, lo, scale_v), g1 = cellf(p1, lo, scale_v);                  // value-map cell (S/4 grid)
    if (!(g0 >= 0.0f && g0 < (float)Sv && g1 >= 0.0f && g1 < (float)Sv)) return;
    const int r = (int)g0, c = (int)g1;
    cell[2 * o] = r; cell[2 * o + 1] = c;
    const float* vm = value_map + (size_t)b * n_ch * Sv * Sv + (size_t)r * Sv + c;
    float best = vm[0];
    for (int ch = 1; ch < n_ch; ++ch) best = fmaxf(best, vm[(size_t)ch * Sv * Sv]);         // torch.max over the 8 headings
    const int x = (int)cellf(p0, lo, scale_s), y = (int)cellf(p1, lo, scale_s);            // S-grid cell, may be slightly negative
    const float* fp = full_proj + (size_t)b * S * S;
    const int xw = wrap_index(x, S), yw = wrap_index(y, S);
    if (xw < 0 || xw >= S || yw < 0 || yw >= S) return;                                     // torch would raise IndexError
    const float dens = fp[(size_t)xw * S + yw];
    // check_pixel_values: any cell == 1 in rows [max(x-w,0), min(x+w+1,S)) x cols [max(y-w,0), min(y+w+1,S))
    bool any = false;
    for (int rr = max(x - window, 0); rr < min(x + window + 1, S) && !any; ++rr)
        for (int cc = max(y - window, 0); cc < min(y + window + 1, S); ++cc)
            if (fp[(size_t)rr * S + cc] == 1.0f) { any = true; break; }
    if (!any) return;
    valid[o] = 1; value[o] = best; density[o] = dens;
}

}  // namespace nbp

using namespace nbp;

extern "C" int nbp_obstacle_fuse(const float* pred, const float* all_cnt, int64_t all_stride, const float* slice_cnt, int64_t slice_stride,
                                 const float* traj, int64_t traj_stride, int n, int S, float threshold, float* fused, float* full_proj, void* stream) {
    if (!pred || !all_cnt || !slice_cnt || !traj || !fused || !full_proj) return invalid("nbp_obstacle_fuse: null pointer argument");
    if (n <= 0 || S <= 0) return invalid("nbp_obstacle_fuse: bad sizes");
    const size_t total = (size_t)n * S * S;
    size_t g = (total + 255) / 256; if (g > 148 * 16) g = 148 * 16;
    obstacle_fuse_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(pred, all_cnt, (int)all_stride, slice_cnt, (int)slice_stride, traj, (int)traj_stride,
                                                                  n, S * S, threshold, fused, full_proj);
    count_launch();
    return check_cuda(cudaGetLastError(), "nbp_obstacle_fuse launch");
}

extern "C" int nbp_candidate_scores(const float* cand, const int32_t* n_cand, int max_cand, const uint8_t* skip, const float* pose,
                                    const float* value_map, int n_ch, int Sv, const float* full_proj, int S, int n,
                                    float range_lo, float range_hi, int window, float* value, float* density, int32_t* cell, uint8_t* valid, void* stream) {
    if (!cand || !n_cand || !pose || !value_map || !full_proj || !value || !density || !cell || !valid) return invalid("nbp_candidate_scores: null pointer argument");
    if (n <= 0 || n > 65535 || max_cand <= 0 || n_ch <= 0 || Sv <= 0 || S <= 0 || !(range_hi > range_lo) || window < 0) return invalid("nbp_candidate_scores: bad sizes");
    const float sv = (float)((double)Sv / ((double)range_hi - (double)range_lo)), ss = (float)((double)S / ((double)range_hi - (double)range_lo));
    candidate_score_kernel<<<dim3((max_cand + 127) / 128, n), 128, 0, (cudaStream_t)stream>>>(cand, n_cand, max_cand, skip, pose, value_map, n_ch, Sv, full_proj, S,
                                                                                            range_lo, sv, ss, window, value, density, cell, valid);
    count_launch();
    return check_cuda(cudaGetLastError(), "nbp_candidate_scores launch");
}
