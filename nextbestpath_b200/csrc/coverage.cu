// Batched surface-coverage metric: the B200-native replacement of
// calculate_coverage_percentage (/root/reference/next_best_path/utility/long_term_utils.py:437-468), which both drivers call
// at the top of every pose iteration (next_best_path/testers/nbp_planning.py:71, next_best_path/utility/nbp_utils.py:572):
//
//     sampled = pc2 if len(pc2) <= weight*len(pc1) else pc2[randperm(len(pc2))[:weight*len(pc1)]]
//     coverage = mean_i [ min_j || pc1[i] - sampled[j] ||_2 < threshold ]            (pc1 = ground-truth surface points)
//
// The reference materialises the full |pc1| x |sampled| distance matrix with torch.cdist (20 000 x 40 000 floats per call).
// Here the ground-truth cloud of every scene is bucketed ONCE into a uniform grid with cell edge >= threshold (host set-up,
// coverage.py); per step one thread per sampled reconstruction point visits the 3x3 rows of 3 x-adjacent cells around it
// (9 contiguous index ranges of the cell-sorted ground truth) and flags the ground-truth points closer than the threshold;
// a second kernel counts the flags.  O(|sampled| * local density) instead of O(|pc1| * |sampled|), all scenes in one launch.
//
// Arithmetic is pinned (one rounding per operation, file compiled with -fmad=false): d2 = (dx*dx + dy*dy) + dz*dz < thr*thr,
// identical to oracle/oracle.py::coverage_percentage, so flag counts are bit-exact against the oracle.  (torch.cdist switches
// to a matmul expansion for large inputs, so the reference itself is only reproducible to ~1e-3 in distance; the fixture
// generated from the reference's own function pins the metric to within the points that close to the threshold.)
#include "nbp_common.cuh"

namespace nbp {

__device__ __forceinline__ uint32_t cov_mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// keyed bijection of [0, n): 4-round balanced Feistel network on 2*half bits with cycle walking (expected < 4 walks)
__device__ __forceinline__ uint32_t cov_perm(uint32_t i, uint32_t n, int half, uint32_t k0, uint32_t k1) {
    const uint32_t mask = (1u << half) - 1u;
    uint32_t v = i;
    do {
        uint32_t l = v >> half, r = v & mask;
#pragma unroll
        for (int round = 0; round < 4; ++round) {
            const uint32_t f = cov_mix(r ^ (round & 1 ? k1 : k0) ^ (0x9e3779b9u * (uint32_t)(round + 1))) & mask;
            const uint32_t t = l ^ f; l = r; r = t;
        }
        v = (l << half) | r;
    } while (v >= n);
    return v;
}

struct CovParams {
    const float* cloud; int64_t cloud_stride; const int32_t* cloud_len;    // reconstruction: [B][cloud_stride][3]
    const int64_t* sample_idx; int64_t sample_stride;                      // optional explicit sample (parity path): [B][sample_stride]
    const float* gt; const int64_t* gt_off;                                // ground truth sorted by cell: [sum G][3], [B+1]
    const int32_t* cell_start; const int64_t* cell_off;                    // per scene ncells+1 starts (scene-local), [B+1]
    const float* origin; const int32_t* dims;                              // [B][3] grid origin, [B][3] cells along x, y, z
    float inv_cell, thr2; int weight; uint64_t seed;
    uint8_t* covered;                                                      // [sum G]
};

__global__ void __launch_bounds__(256) coverage_mark(CovParams p) {
    const int b = blockIdx.y;
    const int64_t g0 = p.gt_off[b];
    const int64_t G = p.gt_off[b + 1] - g0;
    const int64_t N = p.cloud_len[b];
    if (G <= 0 || N <= 0) return;
    const int64_t want = (int64_t)p.weight * G;
    const int64_t K = N <= want ? N : want;                                 // random_sample_pc: all points, or `want` of them
    int half = 1;
    while ((1ll << (2 * half)) < N) ++half;
    const uint32_t k0 = cov_mix((uint32_t)p.seed ^ (uint32_t)(b * 0x85ebca6bu)), k1 = cov_mix((uint32_t)(p.seed >> 32) + 0x27d4eb2fu * (uint32_t)(b + 1));
    const float ox = p.origin[3 * b], oy = p.origin[3 * b + 1], oz = p.origin[3 * b + 2];
    const int nx = p.dims[3 * b], ny = p.dims[3 * b + 1], nz = p.dims[3 * b + 2];
    const int32_t* cs = p.cell_start + p.cell_off[b];
    const float* gt = p.gt + 3 * g0;
    uint8_t* cov = p.covered + g0;
    const float* cl = p.cloud + (size_t)b * p.cloud_stride * 3;

    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < K; j += (int64_t)gridDim.x * blockDim.x) {
        int64_t idx = j;
        if (K < N) idx = p.sample_idx ? p.sample_idx[(size_t)b * p.sample_stride + j] : (int64_t)cov_perm((uint32_t)j, (uint32_t)N, half, k0, k1);
        if (idx < 0 || idx >= N) continue;                                   // malformed explicit sample: ignore the entry
        const float qx = cl[3 * idx], qy = cl[3 * idx + 1], qz = cl[3 * idx + 2];
        const float fx = floorf(fmul(fsub(qx, ox), p.inv_cell)), fy = floorf(fmul(fsub(qy, oy), p.inv_cell)), fz = floorf(fmul(fsub(qz, oz), p.inv_cell));
        // a point more than one cell outside the grid cannot be within the threshold of any ground-truth point
        if (!(fx >= -1.0f && fx <= (float)nx && fy >= -1.0f && fy <= (float)ny && fz >= -1.0f && fz <= (float)nz)) continue;
        const int cx = (int)fx, cy = (int)fy, cz = (int)fz;
        const int x0 = max(cx - 1, 0), x1 = min(cx + 1, nx - 1);
        if (x0 > x1) continue;
        for (int z = max(cz - 1, 0); z <= min(cz + 1, nz - 1); ++z)
            for (int y = max(cy - 1, 0); y <= min(cy + 1, ny - 1); ++y) {
                const int64_t row = ((int64_t)z * ny + y) * nx;
                const int s = cs[row + x0], e = cs[row + x1 + 1];           // x-adjacent cells are contiguous in the cell order
                for (int i = s; i < e; ++i) {
                    const float dx = fsub(gt[3 * i], qx), dy = fsub(gt[3 * i + 1], qy), dz = fsub(gt[3 * i + 2], qz);
                    const float d2 = fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
                    if (d2 < p.thr2) cov[i] = 1;                             // idempotent store: no atomics needed
                }
            }
    }
}

__global__ void __launch_bounds__(256) coverage_count(const uint8_t* __restrict__ covered, const int64_t* __restrict__ gt_off,
                                                      const int32_t* __restrict__ cloud_len, float* __restrict__ out, int32_t* __restrict__ counts) {
    __shared__ int s_w[8];
    const int b = blockIdx.x;
    const int64_t g0 = gt_off[b], G = gt_off[b + 1] - g0;
    int c = 0;
    for (int64_t i = threadIdx.x; i < G; i += blockDim.x) c += covered[g0 + i];
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if (lane_id() == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += s_w[w];
        if (counts) counts[b] = t;
        // (nearest < threshold).float().mean(): an exact integer sum divided in fp32; empty reconstruction -> 0 (long_term_utils.py:461-462)
        out[b] = (G > 0 && cloud_len[b] > 0) ? fdiv((float)t, (float)G) : 0.0f;
    }
}

}  // namespace nbp

using namespace nbp;

extern "C" int nbp_coverage_percentage(const float* cloud, int64_t cloud_stride, const int32_t* cloud_len, const int64_t* sample_idx,
                                       int64_t sample_stride, const float* gt_sorted, const int64_t* gt_off, const int32_t* cell_start,
                                       const int64_t* cell_off, const float* origin, const int32_t* dims, int n_scenes, int64_t max_samples,
                                       int64_t total_gt, float cell, float threshold, int weight, uint64_t seed, uint8_t* covered,
                                       float* coverage, int32_t* counts, void* stream) {
    if (n_scenes == 0) return NBP_OK;
    if (!cloud || !cloud_len || !gt_sorted || !gt_off || !cell_start || !cell_off || !origin || !dims || !covered || !coverage)
        return invalid("nbp_coverage_percentage: null pointer argument");
    if (n_scenes < 0 || n_scenes > 65535 || max_samples < 0 || total_gt < 0 || weight <= 0)
        return invalid("nbp_coverage_percentage: bad sizes n_scenes=%d max_samples=%lld weight=%d", n_scenes, (long long)max_samples, weight);
    if (!(threshold > 0.0f) || !(cell >= threshold)) return invalid("nbp_coverage_percentage: need cell >= threshold > 0 (cell=%g threshold=%g)", cell, threshold);
    if (sample_idx && sample_stride <= 0) return invalid("nbp_coverage_percentage: sample_stride must be positive (>= weight*G_b of every sub-sampled scene)");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_cuda(cudaMemsetAsync(covered, 0, (size_t)total_gt, st), "memset(covered)");
    if (rc) return rc;
    if (max_samples > 0 && total_gt > 0) {
        CovParams p{cloud, cloud_stride, cloud_len, sample_idx, sample_stride, gt_sorted, gt_off, cell_start, cell_off, origin, dims,
                    1.0f / cell, threshold * threshold, weight, seed, covered};
        int gx = (int)((max_samples + 255) / 256);
        if (gx > 148 * 4) gx = 148 * 4;
        coverage_mark<<<dim3(gx, n_scenes), 256, 0, st>>>(p);
        count_launch();
    }
    coverage_count<<<n_scenes, 256, 0, st>>>(covered, gt_off, cloud_len, coverage, counts);
    count_launch();
    return check_cuda(cudaGetLastError(), "nbp_coverage_percentage launch");
}
