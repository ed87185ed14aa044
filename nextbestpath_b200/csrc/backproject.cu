// Depth back-projection + validity filter + k-subset selection + append to per-scene clouds
// (SURVEY.md section 8 rows a4, a5).  Replaces Camera.compute_partial_point_cloud /
// Camera.project_depth_in_3D (/root/reference/macarons/utility/macarons_utils.py:2788-2847) and the
// growing torch.vstack at next_best_path/testers/nbp_planning.py:105,352.
//
// Three launches per call, one CTA (1024 threads) per frame in the two heavy ones:
//   bp_select : n = valid pixels, k = int(n*gf); if k < n, radix-select (9-bit digits, shared histogram; the first
//               pass doubles as the count) the k-th smallest key of a keyed Feistel bijection of the pixel index.
//   bp_offsets: per-frame append offsets (frames of one scene append in frame order) + new cloud_len.
//   bp_write  : ordered compaction of the selected pixels (keep / drop decided once, cached as 4-bit masks in
//               registers), closed-form un-projection (oracle/oracle.py::unproject is the pin: one fp32 rounding
//               per op), float stores.
// Pixels are read four at a time (16-byte loads, four independent Feistel chains per thread): round 1's scalar loops
// were latency-bound at IPC 0.5 (profiles/r02_geometry_ncu_full.txt).  zbuf is read 3x (2 select passes + write), only
// the first read comes from HBM (467 KB per frame stays in the 126 MB L2).
#include "nbp_common.cuh"

namespace nbp {

#ifndef NBP_BP_THREADS
#define NBP_BP_THREADS 1024
#endif
static constexpr int BP_THREADS = NBP_BP_THREADS;          // measured: two 512-thread frames per SM are slower (stage A 0.81 vs 0.66 ms at 256 frames)

struct FrameSel { int32_t n, k; uint32_t tau; int32_t base; int32_t final_len; int32_t pad[3]; };

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

// keyed bijection of [0, 2^(2*half)) : 6-round balanced Feistel network
struct Feistel {
    uint32_t rk[6]; int half; uint32_t hmask;
    __device__ __forceinline__ uint32_t operator()(uint32_t i) const {
        uint32_t L = i >> half, R = i & hmask;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            const uint32_t F = mix32(R * 0x9E3779B1U + rk[r]) & hmask;
            const uint32_t nL = R;
            R = L ^ F;
            L = nL;
        }
        return (L << half) | R;
    }
};

__device__ __forceinline__ Feistel make_feistel(uint64_t seed, int32_t uid, int key_bits) {
    Feistel f;
    f.half = key_bits / 2;
    f.hmask = (1u << f.half) - 1u;
    const uint32_t s0 = (uint32_t)seed, s1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 6; ++r) f.rk[r] = mix32(mix32(mix32(s0 + 0x632be5abU * (uint32_t)(r + 1)) ^ s1) ^ (uint32_t)uid);
    return f;
}

struct BpParams {
    const float* zbuf; const uint8_t* mask; const float* R; const float* T;
    const int32_t* frame_scene; const int32_t* frame_uid; int n_frames;
    int H, W, HW; float wm, hm, mm1; float tan_half, fov_range; double gf; uint64_t seed; int key_bits;
    float* cloud; int32_t* cloud_len; int64_t cap; int n_scenes;
    int32_t* frame_valid; int32_t* frame_kept; int32_t* overflow;
    FrameSel* sel;
    uint32_t* keys;        // optional [n_frames][HW] scratch: the key of every pixel (BP_NO_KEY = invalid pixel), written by the first select
                           // pass and read by the later passes and by bp_write instead of re-evaluating validity + the 6-round Feistel key
};
static constexpr uint32_t BP_NO_KEY = 0xffffffffu;        // keys have at most 26 bits

__device__ __forceinline__ bool depth_valid(const BpParams& p, float d, bool masked_in) {
    bool v = masked_in;                            // an explicit mask replaces the default zbuf > -1 (macarons_utils.py:2771,2825)
    if (p.fov_range > 0.0f) v = v && (d < p.fov_range);
    return v;
}

__device__ __forceinline__ bool pixel_valid(const BpParams& p, const float* z, const uint8_t* m, int i) {
    const float d = __ldg(z + i);
    return depth_valid(p, d, m ? (m[i] != 0) : (d > -1.0f));
}

// validity of the 4 consecutive pixels 4q .. 4q+3 as a bit mask (one 16-byte load; frames are 16-byte aligned when HW % 4 == 0)
__device__ __forceinline__ unsigned quad_valid(const BpParams& p, const float* z, const uint8_t* m, int q, float* d4) {
    const float4 d = __ldg(reinterpret_cast<const float4*>(z) + q);
    d4[0] = d.x; d4[1] = d.y; d4[2] = d.z; d4[3] = d.w;
    unsigned in = 0;
    if (m) {
        const uchar4 mm = *reinterpret_cast<const uchar4*>(m + 4 * (size_t)q);
        in = (mm.x != 0) | ((mm.y != 0) << 1) | ((mm.z != 0) << 2) | ((mm.w != 0) << 3);
    } else {
        in = (d.x > -1.0f) | ((d.y > -1.0f) << 1) | ((d.z > -1.0f) << 2) | ((d.w > -1.0f) << 3);
    }
    unsigned v = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) v |= (depth_valid(p, d4[j], (in >> j) & 1u) ? 1u : 0u) << j;
    return v;
}

// Per frame: n = number of valid pixels, k = int(n * gf), and -- when k < n -- the k-th smallest key tau of a keyed bijection of the
// pixel index (radix select, 9-bit digits, shared histogram): exactly k valid pixels have key <= tau.  The first radix pass doubles as
// the count (n = sum of its histogram); pixels are read four at a time so that four independent Feistel chains are in flight per thread.
__global__ void __launch_bounds__(BP_THREADS, BP_THREADS <= 512 ? 2 : 1) bp_select(BpParams p) {
    __shared__ int s_hist[512];
    __shared__ int s_red[BP_THREADS / 32];
    __shared__ uint32_t s_prefix; __shared__ int s_krem; __shared__ int s_n; __shared__ int s_k;

    const int f = blockIdx.x;
    const float* z = p.zbuf + (size_t)f * p.HW;
    const uint8_t* m = p.mask ? p.mask + (size_t)f * p.HW : nullptr;
    const bool vec = (p.HW & 3) == 0;
    const int nq = vec ? p.HW >> 2 : 0;                       // quads; the scalar loop below covers everything when !vec
    const bool subsample = p.gf < 1.0;
    const Feistel perm = make_feistel(p.seed, p.frame_uid ? p.frame_uid[f] : f, p.key_bits);

    uint32_t tau = 0xffffffffu;
    int bits_left = p.key_bits;
    bool first = true;
    if (threadIdx.x == 0) { s_prefix = 0; s_krem = 0; s_k = 0; }
    do {
        const int db = bits_left >= 9 ? 9 : bits_left;       // digit width of this pass
        const int shift = bits_left - db;
        for (int b = threadIdx.x; b < 512; b += BP_THREADS) s_hist[b] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix;                    // already-fixed high bits (aligned at `bits_left`)
        int cnt = 0;
        auto tally = [&](uint32_t key) {
            if ((key >> bits_left) != prefix) return;
            atomicAdd(&s_hist[(key >> shift) & ((1u << db) - 1u)], 1);
        };
        uint32_t* kf = (p.keys && subsample) ? p.keys + (size_t)f * p.HW : nullptr;
        if (kf && !first) {
            // later passes: the keys are in the scratch (L2), no depth read and no key evaluation
            for (int q = threadIdx.x; q < nq; q += BP_THREADS) {
                const uint4 k4 = *(reinterpret_cast<const uint4*>(kf) + q);
                if (k4.x != BP_NO_KEY) tally(k4.x);
                if (k4.y != BP_NO_KEY) tally(k4.y);
                if (k4.z != BP_NO_KEY) tally(k4.z);
                if (k4.w != BP_NO_KEY) tally(k4.w);
            }
            if (!vec) for (int i = threadIdx.x; i < p.HW; i += BP_THREADS) { const uint32_t k1 = kf[i]; if (k1 != BP_NO_KEY) tally(k1); }
        } else {
            for (int q = threadIdx.x; q < nq; q += BP_THREADS) {
                float d4[4];
                const unsigned v = quad_valid(p, z, m, q, d4);
                if (!subsample) { cnt += __popc(v); continue; }
                uint32_t k4[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    k4[j] = BP_NO_KEY;
                    if ((v >> j) & 1u) { k4[j] = perm((uint32_t)(4 * q + j)); tally(k4[j]); }
                }
                if (kf) *(reinterpret_cast<uint4*>(kf) + q) = make_uint4(k4[0], k4[1], k4[2], k4[3]);
            }
            if (!vec) for (int i = threadIdx.x; i < p.HW; i += BP_THREADS) {
                const bool v = pixel_valid(p, z, m, i);
                if (!subsample) { cnt += v ? 1 : 0; continue; }
                const uint32_t k1 = v ? perm((uint32_t)i) : BP_NO_KEY;
                if (v) tally(k1);
                if (kf) kf[i] = k1;
            }
        }
        if (first && !subsample) {                           // keep everything: only the count is needed
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
            if (lane_id() == 0) s_red[threadIdx.x >> 5] = cnt;
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            // warp 0 scans the histogram: lane l owns bins [l * per, (l + 1) * per) (a 512-iteration loop of one thread was ~10 % of the
            // kernel: every other warp of the CTA waits for it)
            const int nb = 1 << db, per = (nb + 31) >> 5;
            const int b0 = lane_id() * per, b1 = min(nb, b0 + per);
            int mine = 0;
            for (int d = b0; d < b1; ++d) mine += s_hist[d];
            int incl = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane_id() >= d) incl += t; }
            if (first) {
                int n = 0;
                if (subsample) n = __shfl_sync(0xffffffffu, incl, 31);
                else for (int w = 0; w < BP_THREADS / 32; ++w) n += s_red[w];
                int k = subsample ? (int)((double)n * p.gf) : n;
                if (k > n) k = n;
                if (k < 0) k = 0;
                if (lane_id() == 0) { s_n = n; s_k = k; s_krem = k; }
                __syncwarp();
            }
            const int n_all = s_n, k_all = s_k, krem = s_krem;
            __syncwarp();
            if (subsample && k_all > 0 && k_all < n_all) {
                // the digit d with acc(d) < krem <= acc(d) + hist[d] lies in the first lane whose inclusive sum reaches krem
                const unsigned reach = __ballot_sync(0xffffffffu, incl >= krem);
                const int owner = reach ? __ffs(reach) - 1 : 31;
                if (lane_id() == owner) {
                    int acc = incl - mine, d = b0;
                    for (; d < b1; ++d) {
                        if (acc + s_hist[d] >= krem) break;
                        acc += s_hist[d];
                    }
                    if (d >= nb) d = nb - 1;                  // cannot happen (k <= n)
                    s_krem = krem - acc;
                    s_prefix = (prefix << db) | (uint32_t)d;
                }
            }
        }
        __syncthreads();
        first = false;
        bits_left = shift;
    } while (subsample && bits_left > 0 && s_k > 0 && s_k < s_n);
    const int n = s_n, k = s_k;
    if (subsample && k > 0 && k < n) tau = s_prefix;         // keys are unique: exactly k valid pixels have key <= tau
    if (threadIdx.x == 0) {
        FrameSel s; s.n = n; s.k = k; s.tau = tau; s.base = 0; s.final_len = -1; s.pad[0] = s.pad[1] = s.pad[2] = 0;
        p.sel[f] = s;
        if (p.frame_valid) p.frame_valid[f] = n;
        if (p.frame_kept) p.frame_kept[f] = k;
    }
}

__global__ void bp_offsets(BpParams p) {
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < p.n_frames; f += gridDim.x * blockDim.x) {
        const int s = p.frame_scene[f];
        int base = p.cloud_len[s];
        bool last = true;
        for (int g = 0; g < p.n_frames; ++g) {
            if (p.frame_scene[g] != s) continue;
            if (g < f) base += p.sel[g].k;
            if (g > f) last = false;
        }
        p.sel[f].base = base;
        if (last) {
            int64_t e = (int64_t)base + p.sel[f].k;
            p.sel[f].final_len = (int32_t)(e > p.cap ? p.cap : e);
        }
    }
}

static constexpr int BP_MASK_WORDS = 8;        // 4-bit keep masks of 8 quads per word: a warp range of up to 8 * 8 * 128 = 8192 pixels stays in registers

__device__ __forceinline__ void unproject_store(const BpParams& p, const float* sR, const float* sT, int i, float d, float* o) {
    const int row = i / p.W, col = i - row * p.W;
    // NDC tables of macarons_utils.py:2270-2279
    const float nx = fsub(p.wm, fmul(fdiv((float)col, p.mm1), 2.0f));
    const float ny = fsub(p.hm, fmul(fdiv((float)row, p.mm1), 2.0f));
    const float sc = fmul(d, p.tan_half);
    const float dx = fsub(fmul(nx, sc), sT[0]);
    const float dy = fsub(fmul(ny, sc), sT[1]);
    const float dz = fsub(d, sT[2]);
    o[0] = fadd(fadd(fmul(dx, sR[0]), fmul(dy, sR[1])), fmul(dz, sR[2]));
    o[1] = fadd(fadd(fmul(dx, sR[3]), fmul(dy, sR[4])), fmul(dz, sR[5]));
    o[2] = fadd(fadd(fmul(dx, sR[6]), fmul(dy, sR[7])), fmul(dz, sR[8]));
}

__global__ void __launch_bounds__(BP_THREADS, BP_THREADS <= 512 ? 2 : 1) bp_write(BpParams p) {
    __shared__ int s_wcnt[BP_THREADS / 32];
    __shared__ float sR[9], sT[3];
    const int f = blockIdx.x;
    const FrameSel sel = p.sel[f];
    const int scene = p.frame_scene[f];
    if (threadIdx.x < 9) sR[threadIdx.x] = p.R[f * 9 + threadIdx.x];
    if (threadIdx.x < 3) sT[threadIdx.x] = p.T[f * 3 + threadIdx.x];
    if (threadIdx.x == 0 && sel.final_len >= 0) p.cloud_len[scene] = sel.final_len;   // bp_offsets already read the old value
    if (sel.k == 0) return;                                                            // block-uniform
    // (sR / sT become visible at the __syncthreads() below, before their first use)

    const float* z = p.zbuf + (size_t)f * p.HW;
    const uint8_t* m = p.mask ? p.mask + (size_t)f * p.HW : nullptr;
    const bool all = sel.k == sel.n;
    const Feistel perm = make_feistel(p.seed, p.frame_uid ? p.frame_uid[f] : f, p.key_bits);
    float* out = p.cloud + (size_t)scene * (size_t)p.cap * 3;
    const int warp = threadIdx.x >> 5, lane = lane_id(), nwarp = BP_THREADS / 32;
    int dropped = 0;
    const uint32_t* kf = (p.keys && p.gf < 1.0) ? p.keys + (size_t)f * p.HW : nullptr;    // keys left by bp_select (BP_NO_KEY = invalid pixel)
    const uint32_t tau_all = all ? BP_NO_KEY - 1u : sel.tau;                              // keep = key <= tau_all

    // every warp owns one contiguous range of pixels (row-major order is preserved).  Fast path: a lane reads 4 consecutive pixels per
    // iteration (one 16-byte load), decides keep / drop ONCE (validity + key <= tau), remembers the 4-bit decisions in registers, and after
    // the block-wide scan of the per-warp counts writes the kept pixels at (frame base + ranges before + rank inside the range).
    const int per_q = ((((p.HW >> 2) + nwarp - 1) / nwarp) + 31) & ~31;            // quads per warp range
    if ((p.HW & 3) == 0 && per_q <= BP_MASK_WORDS * 8 * 32) {
        const int q0 = warp * per_q, q1 = min(p.HW >> 2, q0 + per_q);
        uint32_t keep[BP_MASK_WORDS];
#pragma unroll
        for (int w = 0; w < BP_MASK_WORDS; ++w) keep[w] = 0;
        const int n_it = per_q >> 5;                                                 // block-uniform; empty tail iterations are skipped
        int cnt = 0;
#pragma unroll
        for (int it = 0; it < BP_MASK_WORDS * 8; ++it) {
            if (it >= n_it) break;
            const int q = q0 + it * 32 + lane;
            unsigned mk = 0;
            if (q < q1 && kf) {
                const uint4 k4 = *(reinterpret_cast<const uint4*>(kf) + q);
                mk = (k4.x <= tau_all ? 1u : 0u) | (k4.y <= tau_all ? 2u : 0u) | (k4.z <= tau_all ? 4u : 0u) | (k4.w <= tau_all ? 8u : 0u);
            } else if (q < q1) {
                float d4[4];
                mk = quad_valid(p, z, m, q, d4);
                if (!all) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (((mk >> j) & 1u) && perm((uint32_t)(4 * q + j)) > sel.tau) mk &= ~(1u << j);
                }
            }
            keep[it >> 3] |= mk << (4 * (it & 7));
            cnt += __popc(mk);
        }
        int wtot = cnt;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) wtot += __shfl_xor_sync(0xffffffffu, wtot, d);
        if (lane == 0) s_wcnt[warp] = wtot;
        __syncthreads();
        int running = sel.base;
        for (int w = 0; w < warp; ++w) running += s_wcnt[w];
#pragma unroll
        for (int it = 0; it < BP_MASK_WORDS * 8; ++it) {
            if (it >= n_it) break;
            const unsigned mk = (keep[it >> 3] >> (4 * (it & 7))) & 15u;
            const int c = __popc(mk);
            int incl = c;                                                            // inclusive scan of the lane counts of this iteration
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
            const int tot = __shfl_sync(0xffffffffu, incl, 31);
            if (mk) {
                const int q = q0 + it * 32 + lane;
                const float4 d = __ldg(reinterpret_cast<const float4*>(z) + q);
                const float dd[4] = {d.x, d.y, d.z, d.w};
                int slot = running + incl - c;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if ((mk >> j) & 1u) {
                        if (slot < p.cap) unproject_store(p, sR, sT, 4 * q + j, dd[j], out + (size_t)slot * 3);
                        else ++dropped;
                        ++slot;
                    }
                }
            }
            running += tot;
        }
    } else {
        // general path (odd frame sizes / very large frames): two passes over the range, the first counts, the second writes
        const int per = (((p.HW + nwarp - 1) / nwarp) + 31) & ~31;
        const int i0 = warp * per, i1 = min(p.HW, i0 + per);
        int cnt = 0;
        for (int b = i0; b < i1; b += 32) {
            const int i = b + lane;
            const bool kp = i < i1 && (kf ? kf[i] <= tau_all : (pixel_valid(p, z, m, i) && (all || perm((uint32_t)i) <= sel.tau)));
            cnt += __popc(__ballot_sync(0xffffffffu, kp));
        }
        if (lane == 0) s_wcnt[warp] = cnt;
        __syncthreads();
        int running = sel.base;
        for (int w = 0; w < warp; ++w) running += s_wcnt[w];
        for (int b = i0; b < i1; b += 32) {
            const int i = b + lane;
            const bool kp = i < i1 && (kf ? kf[i] <= tau_all : (pixel_valid(p, z, m, i) && (all || perm((uint32_t)i) <= sel.tau)));
            const unsigned ball = __ballot_sync(0xffffffffu, kp);
            const int slot = running + __popc(ball & ((1u << lane) - 1u));
            running += __popc(ball);
            if (kp) {
                if (slot < p.cap) unproject_store(p, sR, sT, i, z[i], out + (size_t)slot * 3);
                else ++dropped;
            }
        }
    }
    if (dropped && p.overflow) atomicAdd(p.overflow, dropped);
}

}  // namespace nbp

using namespace nbp;

extern "C" size_t nbp_backproject_workspace_bytes(int n_frames) {
    if (n_frames < 0) return 0;
    return sizeof(FrameSel) * (size_t)(n_frames + 1);
}

static size_t bp_sel_bytes(int n_frames) { return (sizeof(FrameSel) * (size_t)(n_frames + 1) + 255) / 256 * 256; }

extern "C" size_t nbp_backproject_key_cache_bytes(int n_frames, int H, int W) {
    if (n_frames < 0 || H <= 0 || W <= 0) return 0;
    return bp_sel_bytes(n_frames) - nbp_backproject_workspace_bytes(n_frames) + sizeof(uint32_t) * (size_t)n_frames * (size_t)H * (size_t)W;
}

extern "C" int nbp_backproject_append(const float* zbuf, const uint8_t* mask, const float* R, const float* T,
                                      const int32_t* frame_scene, const int32_t* frame_uid, int n_frames,
                                      int H, int W, float tan_half_fov, float fov_range,
                                      double gathering_factor, uint64_t seed,
                                      float* cloud, int32_t* cloud_len, int64_t cloud_capacity, int n_scenes,
                                      int32_t* frame_valid, int32_t* frame_kept, int32_t* overflow,
                                      void* workspace, size_t workspace_bytes, void* stream) {
    if (n_frames == 0) return NBP_OK;
    if (!zbuf || !R || !T || !frame_scene || !cloud || !cloud_len)
        return invalid("nbp_backproject_append: null pointer argument");
    if (n_frames < 0 || H <= 0 || W <= 0 || n_scenes <= 0 || cloud_capacity <= 0 || (int64_t)H * W > (1 << 26))
        return invalid("nbp_backproject_append: bad sizes n_frames=%d H=%d W=%d n_scenes=%d cap=%lld", n_frames, H, W,
                       n_scenes, (long long)cloud_capacity);
    if (cloud_capacity > 0x7fffffffLL) return invalid("nbp_backproject_append: cloud_capacity must fit int32");
    if (!(tan_half_fov > 0.0f)) return invalid("nbp_backproject_append: tan_half_fov must be positive");
    if (!(gathering_factor >= 0.0)) return invalid("nbp_backproject_append: gathering_factor must be >= 0");
    const size_t need = nbp_backproject_workspace_bytes(n_frames);
    if (!workspace || workspace_bytes < need) {
        set_error("nbp_backproject_append: workspace too small (%zu < %zu)", workspace_bytes, need);
        return NBP_ERR_WORKSPACE;
    }
    int key_bits = 2;
    while ((1LL << key_bits) < (int64_t)H * W) key_bits += 2;     // even, so the Feistel halves balance
    // python: W / min(W, H) is a double, cast to fp32 when combined with the fp32 table (macarons_utils.py:2272-2277)
    const int mn = H < W ? H : W;
    const float wm = (float)((double)W / (double)mn), hm = (float)((double)H / (double)mn), mm1 = (float)(mn - 1);
    if (mn < 2) return invalid("nbp_backproject_append: H and W must be >= 2");
    BpParams p{zbuf, mask, R, T, frame_scene, frame_uid, n_frames, H, W, H * W, wm, hm, mm1, tan_half_fov, fov_range,
               gathering_factor, seed, key_bits, cloud, cloud_len, cloud_capacity, n_scenes,
               frame_valid, frame_kept, overflow, (FrameSel*)workspace, nullptr};
    // a workspace that also holds nbp_backproject_key_cache_bytes() lets the kernels evaluate every pixel's key once instead of three times
    if (gathering_factor < 1.0 && workspace_bytes >= need + nbp_backproject_key_cache_bytes(n_frames, H, W))
        p.keys = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(workspace) + bp_sel_bytes(n_frames));
    cudaStream_t st = (cudaStream_t)stream;
    bp_select<<<n_frames, BP_THREADS, 0, st>>>(p);
    count_launch();
    bp_offsets<<<(n_frames + 255) / 256, 256, 0, st>>>(p);
    count_launch();
    bp_write<<<n_frames, BP_THREADS, 0, st>>>(p);
    count_launch();
    return check_cuda(cudaGetLastError(), "nbp_backproject_append launch");
}
