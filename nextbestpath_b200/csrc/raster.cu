// Batched triangle depth rasteriser for sm_100a.
//
// Replaces the PyTorch3D MeshRasterizer call behind Camera.capture_image
// (/root/reference/macarons/utility/macarons_utils.py:2759; settings :905-937).
//
// Two kernels per launch, both over ALL views of the batch:
//   1. raster_setup : one thread per (view, face): world->view->NDC, clip against z = z_clip, cull
//      (non-finite, zero-area, off-screen), append 96-byte triangle records + a pixel-space bounding
//      box to the view's list (warp-aggregated atomic slot claim), and append the record index to the
//      list of every COARSE BIN (<= 32 bins of CB x CB pixels per view, CB a multiple of 16) its box touches.
//   2. raster_tiles : one CTA per (view, 16x16 pixel tile): scans the index list of ITS coarse bin 256
//      entries at a time (not the whole view: 10-30x fewer box tests on indoor meshes), warp-ballot
//      compacts the triangles overlapping the tile into shared memory, then every thread (one pixel)
//      walks the compacted list.
//
// Arithmetic is pinned op-for-op to oracle/raster_oracle.c (one fp32 rounding per operation, no FMA),
// so zbuf is bit-identical to the CPU oracle; the min over (z, face) is order-independent, so the
// non-deterministic append order of kernel 1 does not affect results.
#include <cstdlib>

#include "nbp_common.cuh"

namespace nbp {

static constexpr int TILE = 16;
static constexpr int MAX_BINS = 32;            // coarse bins per view (workspace is sized for this many lists)
static constexpr int RT_THREADS = 256;
static constexpr float K_EPS = 1e-8f;

struct __align__(16) TriRec {          // 24 words
    float v0x, v0y, v1x, v1y;          // NDC xy
    float v2x, v2y, z0, z1;
    float z2, d12x, d12y, d20x;        // dIJ = vJ - vI
    float d20y, d01x, d01y, den;       // den = area + 1e-8
    float xmin, xmax, ymin, ymax;      // NDC bbox
    int face; float zprune; int pad1, pad2;     // zprune: a lower bound of every depth this triangle can produce (see raster_tiles)
};
static_assert(sizeof(TriRec) == 96, "TriRec must be 96 bytes");

struct V3 { float x, y, z; };

__device__ __forceinline__ V3 project_vertex(const float* __restrict__ p, const float* R, const float* T, float focal) {
    const float px = p[0], py = p[1], pz = p[2];
    const float xv = fadd(fadd(fadd(fmul(px, R[0]), fmul(py, R[3])), fmul(pz, R[6])), T[0]);
    const float yv = fadd(fadd(fadd(fmul(px, R[1]), fmul(py, R[4])), fmul(pz, R[7])), T[1]);
    const float zv = fadd(fadd(fadd(fmul(px, R[2]), fmul(py, R[5])), fmul(pz, R[8])), T[2]);
    V3 o;
    o.x = fdiv(fmul(xv, focal), zv);
    o.y = fdiv(fmul(yv, focal), zv);
    o.z = zv;
    return o;
}

__device__ __forceinline__ V3 clip_edge(V3 p1, V3 p2, float clip) {
    const float w = fdiv(fsub(p1.z, clip), fsub(p1.z, p2.z));
    const float omw = fsub(1.0f, w);
    const float p1wx = fmul(p1.x, p1.z), p1wy = fmul(p1.y, p1.z);
    const float p2wx = fmul(p2.x, p2.z), p2wy = fmul(p2.y, p2.z);
    V3 o;
    o.x = fdiv(fadd(fmul(p1wx, omw), fmul(p2wx, w)), clip);
    o.y = fdiv(fadd(fmul(p1wy, omw), fmul(p2wy, w)), clip);
    o.z = clip;
    return o;
}

// NonSquarePixToNdc(i, S1, S2)
__device__ __forceinline__ float pix_to_ndc(int i, int S1, int S2) {
    const float range = (S1 > S2) ? fdiv(fmul(2.0f, (float)S1), (float)S2) : 2.0f;
    const float offset = fdiv(range, 2.0f);
    return fadd(-offset, fdiv(fadd(fmul(range, (float)i), offset), (float)S1));
}

// Conservative pixel index interval [lo, hi] whose centres can satisfy ndc_min <= ndc(pixel) <= ndc_max.
// Output pixel index o maps to NDC through i = S1-1-o (the y/x flip), ndc is increasing in i.
__device__ __forceinline__ void ndc_to_pixel_range(float ndc_min, float ndc_max, int S1, int S2, int& lo, int& hi) {
    const float range = (S1 > S2) ? (2.0f * (float)S1) / (float)S2 : 2.0f;
    const float offset = 0.5f * range;
    // i(ndc) = ((ndc + offset) * S1 - offset) / range
    float i_min = ((ndc_min + offset) * (float)S1 - offset) / range;
    float i_max = ((ndc_max + offset) * (float)S1 - offset) / range;
    i_min = fminf(fmaxf(i_min, -4.0f), (float)S1 + 4.0f);
    i_max = fminf(fmaxf(i_max, -4.0f), (float)S1 + 4.0f);
    const int ii_min = (int)floorf(i_min) - 1, ii_max = (int)ceilf(i_max) + 1;
    lo = S1 - 1 - ii_max;
    hi = S1 - 1 - ii_min;
}

struct SetupParams {
    const float* verts; const int32_t* faces;
    const int64_t* vert_offsets; const int64_t* face_offsets;
    const int32_t* view_scene; const float* R; const float* T;
    int H, W; float focal, z_clip;
    int32_t* tri_count; const int64_t* tri_off; uint2* bbox; TriRec* recs;
    int cb, bins_x, nbins; int32_t* bin_count; uint2* bin_list;       // bin_list[(tri_off[view]*MAX_BINS) + bin*cap(view) + i] = {record, box clamped to the bin}
};

__device__ __forceinline__ bool make_record(const V3& a, const V3& b, const V3& c, int face, int H, int W,
                                            TriRec& r, uint2& bb) {
    // non-finite coordinates can never produce a hit in the oracle (NaN propagates to the inside test)
    const float s = a.x + a.y + a.z + b.x + b.y + b.z + c.x + c.y + c.z;
    if (!isfinite(s)) return false;
    // EdgeFunction(v2, v0, v1)
    const float area = fsub(fmul(fsub(c.x, a.x), fsub(b.y, a.y)), fmul(fsub(c.y, a.y), fsub(b.x, a.x)));
    if (area <= K_EPS && area >= -K_EPS) return false;
    r.v0x = a.x; r.v0y = a.y; r.v1x = b.x; r.v1y = b.y; r.v2x = c.x; r.v2y = c.y;
    r.z0 = a.z; r.z1 = b.z; r.z2 = c.z;
    r.d12x = fsub(c.x, b.x); r.d12y = fsub(c.y, b.y);
    r.d20x = fsub(a.x, c.x); r.d20y = fsub(a.y, c.y);
    r.d01x = fsub(b.x, a.x); r.d01y = fsub(b.y, a.y);
    r.den = fadd(area, K_EPS);
    r.xmin = fminf(fminf(a.x, b.x), c.x); r.xmax = fmaxf(fmaxf(a.x, b.x), c.x);
    r.ymin = fminf(fminf(a.y, b.y), c.y); r.ymax = fmaxf(fmaxf(a.y, b.y), c.y);
    r.face = face; r.pad1 = r.pad2 = 0;
    // the interpolated depth is a convex combination (weights > 0, sum 1 +- a few ulp) of z0, z1, z2 > 0: never below 0.99999 * min z
    r.zprune = fminf(fminf(a.z, b.z), c.z) * 0.99999f;
    int x0, x1, y0, y1;
    ndc_to_pixel_range(r.xmin, r.xmax, W, H, x0, x1);
    ndc_to_pixel_range(r.ymin, r.ymax, H, W, y0, y1);
    if (x1 < 0 || x0 > W - 1 || y1 < 0 || y0 > H - 1) return false;
    x0 = max(x0, 0); y0 = max(y0, 0); x1 = min(x1, W - 1); y1 = min(y1, H - 1);
    bb.x = (uint32_t)x0 | ((uint32_t)x1 << 16);
    bb.y = (uint32_t)y0 | ((uint32_t)y1 << 16);
    return true;
}

__global__ void __launch_bounds__(256) raster_setup(SetupParams p) {
    const int view = blockIdx.y;
    const int scene = p.view_scene[view];
    const int64_t f0 = p.face_offsets[scene];
    const int nf = (int)(p.face_offsets[scene + 1] - f0);
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.x * blockDim.x >= nf) return;          // whole block idle

    __shared__ float sR[9], sT[3];
    if (threadIdx.x < 9) sR[threadIdx.x] = p.R[view * 9 + threadIdx.x];
    if (threadIdx.x < 3) sT[threadIdx.x] = p.T[view * 3 + threadIdx.x];
    __syncthreads();

    TriRec rec[2]; uint2 bb[2]; int n = 0;
    if (f < nf) {
        const float* vbase = p.verts + 3 * p.vert_offsets[scene];
        const int32_t* fi = p.faces + 3 * (f0 + f);
        V3 v[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) v[k] = project_vertex(vbase + 3 * (int64_t)fi[k], sR, sT, p.focal);
        const float clip = p.z_clip;
        const bool b0 = v[0].z < clip, b1 = v[1].z < clip, b2 = v[2].z < clip;
        const int nb = (int)b0 + (int)b1 + (int)b2;
        if (nb == 0) {
            if (make_record(v[0], v[1], v[2], f, p.H, p.W, rec[n], bb[n])) ++n;
        } else if (nb == 1) {
            const int i1 = b0 ? 0 : (b1 ? 1 : 2);
            const V3 p1 = v[i1], p2 = v[(i1 + 1) % 3], p3 = v[(i1 + 2) % 3];
            const V3 p4 = clip_edge(p1, p2, clip), p5 = clip_edge(p1, p3, clip);
            if (make_record(p4, p2, p5, f, p.H, p.W, rec[n], bb[n])) ++n;
            if (make_record(p5, p2, p3, f, p.H, p.W, rec[n], bb[n])) ++n;
        } else if (nb == 2) {
            const int i1 = !b0 ? 0 : (!b1 ? 1 : 2);
            const V3 p1 = v[i1], p2 = v[(i1 + 1) % 3], p3 = v[(i1 + 2) % 3];
            const V3 p4 = clip_edge(p1, p2, clip), p5 = clip_edge(p1, p3, clip);
            if (make_record(p1, p4, p5, f, p.H, p.W, rec[n], bb[n])) ++n;
        }
    }
    // warp-aggregated slot claim: one atomic per warp
    const int lane = lane_id();
    int incl = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    const int warp_total = __shfl_sync(0xffffffffu, incl, 31);
    int base = 0;
    if (lane == 31 && warp_total > 0) base = atomicAdd(&p.tri_count[view], warp_total);
    base = __shfl_sync(0xffffffffu, base, 31);
    const int64_t voff = p.tri_off[view];
    const int64_t slot = voff + base + (incl - n);
    const int64_t cap = p.tri_off[view + 1] - voff;
    for (int k = 0; k < n; ++k) {
        p.bbox[slot + k] = bb[k];
        float4* dst = reinterpret_cast<float4*>(&p.recs[slot + k]);
        const float4* src = reinterpret_cast<const float4*>(&rec[k]);
#pragma unroll
        for (int q = 0; q < 6; ++q) dst[q] = src[q];
        // coarse binning: the record's index (inside the view) goes to every bin its pixel box touches
        const int x0 = (int)(bb[k].x & 0xffff), x1 = (int)(bb[k].x >> 16), y0 = (int)(bb[k].y & 0xffff), y1 = (int)(bb[k].y >> 16);
        const int bx0 = x0 / p.cb, bx1 = x1 / p.cb, by0 = y0 / p.cb, by1 = y1 / p.cb;
        const int idx = base + (incl - n) + k;
        for (int by = by0; by <= by1; ++by)
            for (int bx = bx0; bx <= bx1; ++bx) {
                const int b = by * p.bins_x + bx;
                const int ox = bx * p.cb, oy = by * p.cb;                  // box relative to the bin origin, clamped to the bin (cb <= 256)
                const uint32_t rel = (uint32_t)max(x0 - ox, 0) | ((uint32_t)min(x1 - ox, p.cb - 1) << 8) |
                                     ((uint32_t)max(y0 - oy, 0) << 16) | ((uint32_t)min(y1 - oy, p.cb - 1) << 24);
                const int pos = atomicAdd(&p.bin_count[view * MAX_BINS + b], 1);
                p.bin_list[voff * MAX_BINS + (int64_t)b * cap + pos] = make_uint2((uint32_t)idx, rel);
            }
    }
}

// per-view triangle-list offsets (capacity 2 * faces(scene(view))): one block, chunked sequential prefix + block scan;
// also zeroes the per-view and per-bin counters
__global__ void __launch_bounds__(1024) raster_offsets(const int64_t* face_offsets, const int32_t* view_scene, int n_views,
                                                       int64_t* tri_off, int32_t* tri_count, int32_t* bin_count) {
    __shared__ int64_t s_part[1024];
    const int per = (n_views + 1023) / 1024;
    const int v0 = threadIdx.x * per, v1 = min(n_views, v0 + per);
    int64_t sum = 0;
    for (int v = v0; v < v1; ++v) { const int s = view_scene[v]; sum += 2 * (face_offsets[s + 1] - face_offsets[s]); }
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {                     // inclusive Hillis-Steele scan of the per-thread sums
        const int64_t t = threadIdx.x >= d ? s_part[threadIdx.x - d] : 0;
        __syncthreads();
        s_part[threadIdx.x] += t;
        __syncthreads();
    }
    int64_t acc = s_part[threadIdx.x] - sum;                 // exclusive prefix of this thread's chunk
    for (int v = v0; v < v1; ++v) {
        tri_off[v] = acc;
        const int s = view_scene[v];
        acc += 2 * (face_offsets[s + 1] - face_offsets[s]);
        tri_count[v] = 0;
    }
    if (threadIdx.x == 1023) tri_off[n_views] = s_part[1023];
    for (int i = threadIdx.x; i < n_views * MAX_BINS; i += 1024) bin_count[i] = 0;
}

struct TileParams {
    const int32_t* tri_count; const int64_t* tri_off; const uint2* bbox; const TriRec* recs;
    int H, W, tiles_x; float* zbuf; int32_t* pix_to_face;
    int cb, bins_x; const int32_t* bin_count; const uint2* bin_list;
    int prune;
    int warp8x4;
};

__global__ void __launch_bounds__(RT_THREADS) raster_tiles(TileParams p) {
    __shared__ TriRec s_rec[RT_THREADS];
    __shared__ int s_wcnt[RT_THREADS / 32];

    const int view = blockIdx.y;
    const int tile_x = blockIdx.x % p.tiles_x, tile_y = blockIdx.x / p.tiles_x;
    const int px0 = tile_x * TILE, py0 = tile_y * TILE;
    // a warp covers an 8 x 4 pixel block of the tile (p.warp8x4; 16 x 2 otherwise): the per-triangle box and edge rejections below only
    // save time when ALL lanes of a warp reject, and a compact footprint is rejected by more triangles than a 16-pixel strip
    const int wl = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int tx = p.warp8x4 ? ((wi & 1) * 8 + (wl & 7)) : (threadIdx.x & (TILE - 1));
    const int ty = p.warp8x4 ? ((wi >> 1) * 4 + (wl >> 3)) : (threadIdx.x >> 4);
    const int xi = px0 + tx, yi = py0 + ty;
    const float xf = pix_to_ndc(p.W - 1 - xi, p.W, p.H);
    const float yf = pix_to_ndc(p.H - 1 - yi, p.H, p.W);

    const int64_t base = p.tri_off[view];
    const int64_t cap = p.tri_off[view + 1] - base;
    const int bin_x = px0 / p.cb, bin_y = py0 / p.cb;               // TILE divides cb: a tile lies in exactly one coarse bin
    const int bin = bin_y * p.bins_x + bin_x;
    const int n = p.bin_count[view * MAX_BINS + bin];
    const uint2* list = p.bin_list + base * MAX_BINS + (int64_t)bin * cap;
    const int rx0 = px0 - bin_x * p.cb, ry0 = py0 - bin_y * p.cb;   // tile origin relative to the bin
    float best_z = -1.0f; int best_f = -1;

    // the list entry carries the (bin-clamped) box, so the overlap test needs no dependent load; the next chunk's entries are
    // fetched before the current chunk is processed
    uint2 e_next = make_uint2(0u, 0u);
    if ((int)threadIdx.x < n) e_next = __ldg(&list[threadIdx.x]);
    for (int c0 = 0; c0 < n; c0 += RT_THREADS) {
        const int li = c0 + threadIdx.x;
        const uint2 e = e_next;
        if (li + RT_THREADS < n) e_next = __ldg(&list[li + RT_THREADS]);
        bool ov = false;
        const int t = (int)e.x;
        if (li < n) {
            const int x0 = e.y & 0xff, x1 = (e.y >> 8) & 0xff, y0 = (e.y >> 16) & 0xff, y1 = e.y >> 24;
            ov = !(x1 < rx0 || x0 > rx0 + TILE - 1 || y1 < ry0 || y0 > ry0 + TILE - 1);
        }
        int m;
        const int slot = block_compact(ov, s_wcnt, m);
        if (ov) {
            const float4* src = reinterpret_cast<const float4*>(&p.recs[base + t]);
            float4* dst = reinterpret_cast<float4*>(&s_rec[slot]);
#pragma unroll
            for (int q = 0; q < 6; ++q) dst[q] = __ldg(src + q);
        }
        __syncthreads();
        for (int j = 0; j < m; ++j) {
            const TriRec& r = s_rec[j];
            if (xf > r.xmax || xf < r.xmin || yf > r.ymax || yf < r.ymin) continue;
            const float e0 = fsub(fmul(fsub(xf, r.v1x), r.d12y), fmul(fsub(yf, r.v1y), r.d12x));
            const float e1 = fsub(fmul(fsub(xf, r.v2x), r.d20y), fmul(fsub(yf, r.v2y), r.d20x));
            const float e2 = fsub(fmul(fsub(xf, r.v0x), r.d01y), fmul(fsub(yf, r.v0y), r.d01x));
            const float den = r.den;
            const bool cand = (den > 0.0f) ? (e0 > 0.0f && e1 > 0.0f && e2 > 0.0f)
                                           : (e0 < 0.0f && e1 < 0.0f && e2 < 0.0f);
            if (!cand) continue;
            // a triangle that lies wholly behind the depth already found cannot win (nor tie): skip its six IEEE divisions.  Only valid
            // for vertex depths > 0 (z_clip > 0): p.prune
            if (p.prune && best_f >= 0 && r.zprune > best_z) continue;
            const float w0 = fdiv(e0, den), w1 = fdiv(e1, den), w2 = fdiv(e2, den);
            const float z0 = r.z0, z1 = r.z1, z2 = r.z2;
            const float t0 = fmul(fmul(w0, z1), z2);
            const float t1 = fmul(fmul(z0, w1), z2);
            const float t2 = fmul(fmul(z0, z1), w2);
            const float d = fmaxf(fadd(fadd(t0, t1), t2), K_EPS);
            const float l0 = fdiv(t0, d), l1 = fdiv(t1, d), l2 = fdiv(t2, d);
            const float pz = fadd(fadd(fmul(l0, z0), fmul(l1, z1)), fmul(l2, z2));
            if (pz < 0.0f) continue;
            if (!(l0 > 0.0f && l1 > 0.0f && l2 > 0.0f)) continue;
            const int f = r.face;
            if (best_f < 0 || pz < best_z || (pz == best_z && f < best_f)) { best_z = pz; best_f = f; }
        }
        __syncthreads();
    }
    if (xi < p.W && yi < p.H) {
        const size_t o = ((size_t)view * p.H + yi) * p.W + xi;
        p.zbuf[o] = best_z;
        if (p.pix_to_face) p.pix_to_face[o] = best_f;
    }
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace nbp

using namespace nbp;

extern "C" size_t nbp_raster_workspace_bytes(int n_views, int64_t total_view_faces) {
    if (n_views < 0 || total_view_faces < 0) return 0;
    size_t b = 0;
    b += align_up(sizeof(int32_t) * (size_t)(n_views + 1), 256);          // tri_count
    b += align_up(sizeof(int64_t) * (size_t)(n_views + 1), 256);          // tri_off
    b += align_up(sizeof(uint2) * (size_t)(2 * total_view_faces + 1), 256);   // bbox
    b += align_up(sizeof(TriRec) * (size_t)(2 * total_view_faces + 1), 256);  // records
    b += align_up(sizeof(int32_t) * (size_t)MAX_BINS * (size_t)(n_views + 1), 256);                 // coarse-bin counters
    b += align_up(sizeof(uint2) * (size_t)MAX_BINS * (size_t)(2 * total_view_faces + 1), 256);      // coarse-bin lists {record, box}
    return b;
}

extern "C" int nbp_raster_depth_batched(const float* verts, const int32_t* faces,
                                        const int64_t* vert_offsets, const int64_t* face_offsets, int n_scenes,
                                        const int32_t* view_scene, const float* R, const float* T, int n_views,
                                        int64_t total_view_faces, int max_faces_per_scene,
                                        int H, int W, float tan_half_fov, float z_clip,
                                        float* zbuf, int32_t* pix_to_face,
                                        void* workspace, size_t workspace_bytes, void* stream) {
    if (n_views == 0) return NBP_OK;
    if (!vert_offsets || !face_offsets || !view_scene || !R || !T || !zbuf)
        return invalid("nbp_raster_depth_batched: null pointer argument");
    if (max_faces_per_scene > 0 && (!verts || !faces))
        return invalid("nbp_raster_depth_batched: null verts/faces with a non-empty mesh");
    if (n_views < 0 || n_scenes <= 0 || H <= 0 || W <= 0 || H > 32767 || W > 32767)
        return invalid("nbp_raster_depth_batched: bad sizes n_views=%d n_scenes=%d H=%d W=%d", n_views, n_scenes, H, W);
    if (n_views > 65535) return invalid("nbp_raster_depth_batched: n_views=%d exceeds 65535 per call", n_views);
    if (!(tan_half_fov > 0.0f) || !(z_clip > 0.0f))
        return invalid("nbp_raster_depth_batched: tan_half_fov and z_clip must be positive");
    if (max_faces_per_scene < 0 || total_view_faces < 0) return invalid("nbp_raster_depth_batched: negative face count");
    const size_t need = nbp_raster_workspace_bytes(n_views, total_view_faces);
    if (!workspace || workspace_bytes < need) {
        set_error("nbp_raster_depth_batched: workspace too small (%zu < %zu)", workspace_bytes, need);
        return NBP_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    char* w = (char*)workspace;
    int32_t* tri_count = (int32_t*)w; w += align_up(sizeof(int32_t) * (size_t)(n_views + 1), 256);
    int64_t* tri_off = (int64_t*)w;   w += align_up(sizeof(int64_t) * (size_t)(n_views + 1), 256);
    uint2* bbox = (uint2*)w;          w += align_up(sizeof(uint2) * (size_t)(2 * total_view_faces + 1), 256);
    TriRec* recs = (TriRec*)w;        w += align_up(sizeof(TriRec) * (size_t)(2 * total_view_faces + 1), 256);
    int32_t* bin_count = (int32_t*)w; w += align_up(sizeof(int32_t) * (size_t)MAX_BINS * (size_t)(n_views + 1), 256);
    uint2* bin_list = (uint2*)w;
    // coarse bin edge: the smallest multiple of the tile size that keeps the bin count within MAX_BINS
    int cb = TILE;
    while (((W + cb - 1) / cb) * ((H + cb - 1) / cb) > MAX_BINS) cb += TILE;
    const int bins_x = (W + cb - 1) / cb, nbins = bins_x * ((H + cb - 1) / cb);
    if (cb > 256) return invalid("nbp_raster_depth_batched: image %dx%d needs coarse bins wider than 256 pixels (unsupported)", H, W);

    raster_offsets<<<1, 1024, 0, st>>>(face_offsets, view_scene, n_views, tri_off, tri_count, bin_count);
    count_launch();
    if (max_faces_per_scene > 0) {
        SetupParams sp{verts, faces, vert_offsets, face_offsets, view_scene, R, T, H, W,
                       1.0f / tan_half_fov, z_clip, tri_count, tri_off, bbox, recs, cb, bins_x, nbins, bin_count, bin_list};
        dim3 g((max_faces_per_scene + 255) / 256, n_views);
        raster_setup<<<g, 256, 0, st>>>(sp);
        count_launch();
    }
    const int tiles_x = (W + TILE - 1) / TILE, tiles_y = (H + TILE - 1) / TILE;
    TileParams tp{tri_count, tri_off, bbox, recs, H, W, tiles_x, zbuf, pix_to_face, cb, bins_x, bin_count, bin_list, z_clip > 1e-6f ? 1 : 0, 1};
    {
        static int w84_env = -1;                             // NBP_RASTER_WARP8X4=0: A/B switch back to 16 x 2 warps
        if (w84_env < 0) { const char* e = getenv("NBP_RASTER_WARP8X4"); w84_env = e ? atoi(e) : 1; }
        tp.warp8x4 = w84_env;
    }
    raster_tiles<<<dim3(tiles_x * tiles_y, n_views), RT_THREADS, 0, st>>>(tp);
    count_launch();
    return check_cuda(cudaGetLastError(), "nbp_raster_depth_batched launch");
}
