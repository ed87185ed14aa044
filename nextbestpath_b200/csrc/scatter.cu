// Height-slab split + egocentric transform + histogram into the (n_pieces+1, S, S) model-input grid
// (SURVEY.md section 8 rows a6-a10).  Replaces, per scene and per step,
//   torch.bucketize(full_pc[:,1], y_bins[:-1]) - 1            next_best_path/testers/nbp_planning.py:114-115
//   transform_points_to_n_pieces(no_rotation=True)            next_best_path/utility/utils.py:166-196
//   map_points_to_n_imgs -> index_put_(accumulate=True)       next_best_path/utility/utils.py:198-223
//   the trajectory image + torch.cat                          nbp_planning.py:126-132
//
// HBM-bound streaming kernel: each thread reads 4 points as three coalesced float4 loads (48 B), maps
// them to cells with op-for-op pinned fp32 arithmetic (so cell indices are bit-exact with torch), and
// warp-aggregates equal cells (__match_any_sync) into one fp32 RED.ADD per distinct cell per warp.
// Counts are integers < 2^24, so fp32 accumulation is exact and order-independent.
#include <cstdlib>

#include "nbp_common.cuh"

namespace nbp {

static constexpr int SC_THREADS = 256;
static constexpr int SC_PTS_PER_THREAD = 4;
static constexpr int SC_MAX_BOUNDS = 8;

struct ScatterParams {
    const float* cloud; const int32_t* cloud_len; int64_t cap;
    const float* traj; const int32_t* traj_len; int64_t tcap;
    const float* pose; const float* bounds; const int32_t* n_bounds; int max_bounds;
    int n_pieces, S; float lo, scale; float* grid;
};

// rint((v - lo) * scale) with torch's rounding: each op rounded to fp32, half-to-even
__device__ __forceinline__ float cell_coord(float v, float lo, float scale) { return rintf(fmul(fsub(v, lo), scale)); }

__device__ __forceinline__ int point_cell(float x, float y, float z, float cx, float cz, const float* b, int nb,
                                          int n_pieces, int S, float lo, float scale) {
    int slab = -1;
#pragma unroll
    for (int q = 0; q < SC_MAX_BOUNDS; ++q) slab += (y > b[q]) ? 1 : 0;     // bucketize(right=False) - 1; b[q >= nb] = +inf (set by the callers)
    if (slab < 0 || slab >= n_pieces) return -1;
    const float r = cell_coord(-fsub(z, cz), lo, scale);      // rows <- -(z - c_z)
    const float c = cell_coord(-fsub(x, cx), lo, scale);      // cols <- -(x - c_x)
    if (!(r >= 0.0f && r < (float)S && c >= 0.0f && c < (float)S)) return -1;
    return (slab * S + (int)r) * S + (int)c;
}

__device__ __forceinline__ void warp_aggregated_add(float* grid, int cell) {
    // all 32 lanes call; cell < 0 = nothing to add
    const unsigned peers = __match_any_sync(0xffffffffu, cell);
    if (cell >= 0 && lane_id() == (__ffs(peers) - 1)) atomicAdd(grid + cell, (float)__popc(peers));
}

__global__ void __launch_bounds__(SC_THREADS) grid_scatter(ScatterParams p) {
    const int scene = blockIdx.y;
    const int n = p.cloud_len[scene];
    __shared__ float s_b[SC_MAX_BOUNDS];
    __shared__ float s_c[2];
    __shared__ int s_nb;
    if (threadIdx.x < SC_MAX_BOUNDS)
        s_b[threadIdx.x] = (int)threadIdx.x < min(p.n_bounds[scene], p.max_bounds) ? p.bounds[(size_t)scene * p.max_bounds + threadIdx.x] : INFINITY;
    if (threadIdx.x == 0) {
        s_c[0] = p.pose[scene * 5 + 0]; s_c[1] = p.pose[scene * 5 + 2];
        s_nb = min(p.n_bounds[scene], p.max_bounds);
    }
    __syncthreads();
    float b[SC_MAX_BOUNDS];
#pragma unroll
    for (int q = 0; q < SC_MAX_BOUNDS; ++q) b[q] = s_b[q];
    const int nb = s_nb;
    const float cx = s_c[0], cz = s_c[1];
    float* g = p.grid + (size_t)scene * (size_t)(p.n_pieces + 1) * p.S * p.S;

    // ---- the cloud: groups of 4 points = 3 float4
    const float4* src = reinterpret_cast<const float4*>(p.cloud + (size_t)scene * (size_t)p.cap * 3);
    const int n_groups = (n + 3) / 4;
    const int groups_padded = (n_groups + 31) & ~31;               // keep warps converged for match_any
    for (int gi = blockIdx.x * SC_THREADS + threadIdx.x; gi < groups_padded; gi += gridDim.x * SC_THREADS) {
        float v[12];
        const bool live = gi < n_groups;
        if (live) {
            const float4 a = __ldcs(src + 3 * (size_t)gi), bq = __ldcs(src + 3 * (size_t)gi + 1), c = __ldcs(src + 3 * (size_t)gi + 2);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = bq.x; v[5] = bq.y; v[6] = bq.z; v[7] = bq.w;
            v[8] = c.x; v[9] = c.y; v[10] = c.z; v[11] = c.w;
        }
#pragma unroll
        for (int k = 0; k < SC_PTS_PER_THREAD; ++k) {
            int cell = -1;
            if (live && 4 * gi + k < n) cell = point_cell(v[3 * k], v[3 * k + 1], v[3 * k + 2], cx, cz, b, nb, p.n_pieces, p.S, p.lo, p.scale);
            warp_aggregated_add(g, cell);
        }
    }

    // ---- the trajectory image (channel n_pieces): a few hundred points, first CTA of the scene
    if (blockIdx.x == 0 && p.traj) {
        const int nt = p.traj_len[scene];
        const float* t = p.traj + (size_t)scene * (size_t)p.tcap * 3;
        float* gt = g + (size_t)p.n_pieces * p.S * p.S;
        const int padded = (nt + 31) & ~31;
        for (int i = threadIdx.x; i < padded; i += SC_THREADS) {
            int cell = -1;
            if (i < nt) {
                const float r = cell_coord(-fsub(t[3 * i + 2], cz), p.lo, p.scale);
                const float c = cell_coord(-fsub(t[3 * i + 0], cx), p.lo, p.scale);
                if (r >= 0.0f && r < (float)p.S && c >= 0.0f && c < (float)p.S) cell = (int)r * p.S + (int)c;
            }
            warp_aggregated_add(gt, cell);
        }
    }
}

// Rollout clouds hit the same wall cells again and again (every frame that sees a wall adds ~10 points per cell and height slab), and
// the kernel above is then limited by same-address fp32 REDs in L2 (ncu: 19 M RED sectors per 48 M points, 0.26 of the HBM peak), not
// by the 12 bytes per point it streams.  This variant gives every CTA a contiguous chunk of SH_CHUNK_GROUPS * 4 points (~3 frames of a
// rollout) and counts it in a shared-memory hash table (open addressing, integer counts) first; each distinct cell of the chunk then
// costs ONE RED.  Counts are integers, so the grid is bit-identical to the direct kernel's; a crowded table (16 probes) falls back to
// a direct RED for that point.
static constexpr int SH_THREADS = 512;
static constexpr int SH_SLOTS = 8192;                      // 32 KB keys + 32 KB counts
static constexpr int SH_CHUNK_GROUPS = 4096;               // groups of 4 points per chunk

__global__ void __launch_bounds__(SH_THREADS) grid_scatter_hash(ScatterParams p) {
    extern __shared__ int sh_tab[];
    volatile int* keys = sh_tab;
    int* cnts = sh_tab + SH_SLOTS;
    const int scene = blockIdx.y;
    const int n = p.cloud_len[scene];
    __shared__ float s_b[SC_MAX_BOUNDS];
    __shared__ float s_c[2];
    __shared__ int s_nb;
    if (threadIdx.x < SC_MAX_BOUNDS)
        s_b[threadIdx.x] = (int)threadIdx.x < min(p.n_bounds[scene], p.max_bounds) ? p.bounds[(size_t)scene * p.max_bounds + threadIdx.x] : INFINITY;
    if (threadIdx.x == 0) {
        s_c[0] = p.pose[scene * 5 + 0]; s_c[1] = p.pose[scene * 5 + 2];
        s_nb = min(p.n_bounds[scene], p.max_bounds);
    }
    __syncthreads();
    float b[SC_MAX_BOUNDS];
#pragma unroll
    for (int q = 0; q < SC_MAX_BOUNDS; ++q) b[q] = s_b[q];
    const int nb = s_nb;
    const float cx = s_c[0], cz = s_c[1];
    float* g = p.grid + (size_t)scene * (size_t)(p.n_pieces + 1) * p.S * p.S;
    const float4* src = reinterpret_cast<const float4*>(p.cloud + (size_t)scene * (size_t)p.cap * 3);
    const int n_groups = (n + 3) / 4;
    const int n_chunks = (n_groups + SH_CHUNK_GROUPS - 1) / SH_CHUNK_GROUPS;
    if ((int)blockIdx.x < n_chunks) for (int i = threadIdx.x; i < SH_SLOTS; i += SH_THREADS) { keys[i] = -1; cnts[i] = 0; }
    __syncthreads();
    for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const int g1 = min(n_groups, (chunk + 1) * SH_CHUNK_GROUPS);
        for (int gi = chunk * SH_CHUNK_GROUPS + threadIdx.x; gi < g1; gi += SH_THREADS) {
            const float4 a = __ldcs(src + 3 * (size_t)gi), bq = __ldcs(src + 3 * (size_t)gi + 1), c = __ldcs(src + 3 * (size_t)gi + 2);
            const float v[12] = {a.x, a.y, a.z, a.w, bq.x, bq.y, bq.z, bq.w, c.x, c.y, c.z, c.w};
#pragma unroll
            for (int k = 0; k < SC_PTS_PER_THREAD; ++k) {
                if (4 * gi + k >= n) break;
                const int cell = point_cell(v[3 * k], v[3 * k + 1], v[3 * k + 2], cx, cz, b, nb, p.n_pieces, p.S, p.lo, p.scale);
                if (cell < 0) continue;
                unsigned slot = ((unsigned)cell * 2654435761u) >> 19;            // 13 bits
                bool done = false;
#pragma unroll 1
                for (int probe = 0; probe < 16 && !done; ++probe) {
                    int cur = keys[slot];
                    if (cur == -1) cur = atomicCAS(const_cast<int*>(keys) + slot, -1, cell);
                    if (cur == cell || cur == -1) { atomicAdd(cnts + slot, 1); done = true; }
                    slot = (slot + 1) & (SH_SLOTS - 1);
                }
                if (!done) atomicAdd(g + cell, 1.0f);
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < SH_SLOTS; i += SH_THREADS) {              // flush, and leave the slot empty for the next chunk
            const int key = keys[i];
            if (key >= 0) { atomicAdd(g + key, (float)cnts[i]); keys[i] = -1; cnts[i] = 0; }
        }
        __syncthreads();
    }

    // ---- the trajectory image (channel n_pieces): a few hundred points, first CTA of the scene
    if (blockIdx.x == 0 && p.traj) {
        const int nt = p.traj_len[scene];
        const float* t = p.traj + (size_t)scene * (size_t)p.tcap * 3;
        float* gt = g + (size_t)p.n_pieces * p.S * p.S;
        const int padded = (nt + 31) & ~31;
        for (int i = threadIdx.x; i < padded; i += SH_THREADS) {
            int cell = -1;
            if (i < nt) {
                const float r = cell_coord(-fsub(t[3 * i + 2], cz), p.lo, p.scale);
                const float c = cell_coord(-fsub(t[3 * i + 0], cx), p.lo, p.scale);
                if (r >= 0.0f && r < (float)p.S && c >= 0.0f && c < (float)p.S) cell = (int)r * p.S + (int)c;
            }
            warp_aggregated_add(gt, cell);
        }
    }
}

// plain map_points_to_n_imgs: points_2d [n, m, 2]
__global__ void __launch_bounds__(SC_THREADS) map_points_kernel(const float* pts, const int32_t* lens, int64_t m, int S0, int S1,
                                                                float lo, float sx, float sy, float* out) {
    const int img = blockIdx.y;
    const int64_t cnt = lens ? (int64_t)lens[img] : m;
    const float2* src = reinterpret_cast<const float2*>(pts) + (size_t)img * m;
    float* g = out + (size_t)img * S0 * S1;
    const int64_t padded = (cnt + 31) & ~(int64_t)31;
    for (int64_t i = (int64_t)blockIdx.x * SC_THREADS + threadIdx.x; i < padded; i += (int64_t)gridDim.x * SC_THREADS) {
        int cell = -1;
        if (i < cnt) {
            const float2 q = src[i];
            const float r = cell_coord(q.x, lo, sx), c = cell_coord(q.y, lo, sy);
            if (r >= 0.0f && r < (float)S0 && c >= 0.0f && c < (float)S1) cell = (int)r * S1 + (int)c;
        }
        warp_aggregated_add(g, cell);
    }
}

__global__ void point_cells_kernel(const float* pts, int64_t n, float lo, float sx, float sy, int64_t* cells) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        cells[i] = (int64_t)cell_coord(pts[2 * i], lo, sx);
        cells[n + i] = (int64_t)cell_coord(pts[2 * i + 1], lo, sy);
    }
}

}  // namespace nbp

using namespace nbp;

extern "C" int nbp_grid_scatter(const float* cloud, const int32_t* cloud_len, int64_t cloud_capacity,
                                const float* traj, const int32_t* traj_len, int64_t traj_capacity,
                                const float* pose, const float* slab_bounds, const int32_t* n_bounds, int max_bounds,
                                int n_scenes, int n_pieces, int S, float range_lo, float range_hi,
                                int64_t max_points, float* grid, void* stream) {
    if (n_scenes == 0) return NBP_OK;
    if (!cloud || !cloud_len || !pose || !slab_bounds || !n_bounds || !grid)
        return invalid("nbp_grid_scatter: null pointer argument");
    if (traj && !traj_len) return invalid("nbp_grid_scatter: traj given without traj_len");
    if (n_scenes < 0 || n_scenes > 65535 || n_pieces <= 0 || S <= 0 || max_bounds <= 0 || max_bounds > SC_MAX_BOUNDS)
        return invalid("nbp_grid_scatter: bad sizes n_scenes=%d n_pieces=%d S=%d max_bounds=%d (<=%d)", n_scenes, n_pieces, S,
                       max_bounds, SC_MAX_BOUNDS);
    if (cloud_capacity <= 0 || (cloud_capacity & 3)) return invalid("nbp_grid_scatter: cloud_capacity must be a positive multiple of 4");
    if (!(range_hi > range_lo)) return invalid("nbp_grid_scatter: empty grid range");
    if ((int64_t)(n_pieces + 1) * S * S >= (1LL << 31)) return invalid("nbp_grid_scatter: grid too large");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t bytes = sizeof(float) * (size_t)n_scenes * (size_t)(n_pieces + 1) * S * S;
    int rc = check_cuda(cudaMemsetAsync(grid, 0, bytes, st), "nbp_grid_scatter memset");
    if (rc) return rc;
    count_launch();
    if (max_points < 0 || max_points > cloud_capacity) max_points = cloud_capacity;
    const int64_t groups = (max_points + 3) / 4;
    int gx = (int)((groups + SC_THREADS * 4 - 1) / (SC_THREADS * 4));       // ~4 groups (16 points) per thread
    if (gx < 1) gx = 1;
    if (gx > 148 * 8) gx = 148 * 8;
    // torch: scale = S / (hi - lo) as a python float, cast to fp32 when multiplied with the fp32 tensor
    const float scale = (float)((double)S / ((double)range_hi - (double)range_lo));
    ScatterParams p{cloud, cloud_len, cloud_capacity, traj, traj_len, traj_capacity, pose, slab_bounds, n_bounds, max_bounds,
                    n_pieces, S, range_lo, scale, grid};
    // clouds of a few frames or more go through the shared-memory hash (NBP_SCATTER_HASH=0: always the direct kernel, A/B switch)
    static int hash_env = -1;
    if (hash_env < 0) { const char* e = getenv("NBP_SCATTER_HASH"); hash_env = e ? atoi(e) : 1; }
    if (hash_env && groups >= 2 * SH_CHUNK_GROUPS) {
        static bool attr_d[NBP_MAX_DEVICES] = {};
        bool& attr = attr_d[device_slot()];
        const int smem = SH_SLOTS * 2 * (int)sizeof(int);
        if (!attr) {
            rc = check_cuda(cudaFuncSetAttribute(grid_scatter_hash, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "cudaFuncSetAttribute(grid_scatter_hash)");
            if (rc) return rc;
            attr = true;
        }
        int hx = (int)((groups + SH_CHUNK_GROUPS - 1) / SH_CHUNK_GROUPS);
        if (hx > 148 * 4) hx = 148 * 4;
        grid_scatter_hash<<<dim3(hx, n_scenes), SH_THREADS, smem, st>>>(p);
        count_launch();
        return check_cuda(cudaGetLastError(), "nbp_grid_scatter launch");
    }
    grid_scatter<<<dim3(gx, n_scenes), SC_THREADS, 0, st>>>(p);
    count_launch();
    return check_cuda(cudaGetLastError(), "nbp_grid_scatter launch");
}

extern "C" int nbp_map_points(const float* points_2d, const int32_t* lens, int n, int64_t m, int S0, int S1,
                              float range_lo, float range_hi, float* out, void* stream) {
    if (n == 0) return NBP_OK;
    if (!out || (m > 0 && !points_2d)) return invalid("nbp_map_points: null pointer argument");
    if (n < 0 || n > 65535 || m < 0 || S0 <= 0 || S1 <= 0 || (int64_t)S0 * S1 >= (1LL << 31))
        return invalid("nbp_map_points: bad sizes n=%d m=%lld S=(%d,%d)", n, (long long)m, S0, S1);
    if (!(range_hi > range_lo)) return invalid("nbp_map_points: empty grid range");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_cuda(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)n * S0 * S1, st), "nbp_map_points memset");
    if (rc) return rc;
    count_launch();
    if (m == 0) return NBP_OK;
    const float sx = (float)((double)S0 / ((double)range_hi - (double)range_lo));
    const float sy = (float)((double)S1 / ((double)range_hi - (double)range_lo));
    int gx = (int)((m + SC_THREADS * 8 - 1) / (SC_THREADS * 8));
    if (gx < 1) gx = 1;
    if (gx > 148 * 8) gx = 148 * 8;
    map_points_kernel<<<dim3(gx, n), SC_THREADS, 0, st>>>(points_2d, lens, m, S0, S1, range_lo, sx, sy, out);
    count_launch();
    return check_cuda(cudaGetLastError(), "nbp_map_points launch");
}

extern "C" int nbp_point_cells(const float* points_2d, int64_t n, int S0, int S1, float range_lo, float range_hi,
                               int64_t* cells, void* stream) {
    if (n == 0) return NBP_OK;
    if (!points_2d || !cells) return invalid("nbp_point_cells: null pointer argument");
    if (n < 0 || S0 <= 0 || S1 <= 0 || !(range_hi > range_lo)) return invalid("nbp_point_cells: bad arguments");
    const float sx = (float)((double)S0 / ((double)range_hi - (double)range_lo));
    const float sy = (float)((double)S1 / ((double)range_hi - (double)range_lo));
    int g = (int)((n + 255) / 256);
    if (g > 148 * 8) g = 148 * 8;
    point_cells_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(points_2d, n, range_lo, sx, sy, cells);
    count_launch();
    return check_cuda(cudaGetLastError(), "nbp_point_cells launch");
}
