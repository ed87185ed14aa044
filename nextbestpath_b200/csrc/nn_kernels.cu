// CUDA-core kernels around the tcgen05 convolution (the degenerate
// GEMM shapes K=45, N=1, N=8 and the pointwise glue of /root/reference/next_best_path/networks/nbp_model.py):
//   conv_first     Conv1.conv.0 (5->64, 3x3) + folded BN + ReLU, fp32 NCHW counts in, NHWC fp16 out   (:11-13)
//   maxpool2x2     nn.MaxPool2d(2,2)                                                              (:68)
//   upsample2x     nn.Upsample(scale_factor=2) (nearest)                                          (:27)
//   att_gate       psi = sigmoid(BN(conv1x1(a))) ; out = x * psi                                   (:49-62)
//   conv1x1_head   Final1 (256->8) / Final2 (64->1 + sigmoid), NHWC fp16 in, NCHW fp32 out         (:89,106-108)
// All are HBM-bound streaming kernels: 16-byte vector accesses, one pass over their inputs.
#include <cuda_fp16.h>

#include <cstdlib>

#include "nbp_common.cuh"

namespace nbp {

// Activation format helpers.  A tensor is NHWC fp16; `pix` points at the first element of a pixel, `ch` is a channel index
// (multiple of 8) and `lo` the offset of the pixel's second plane (0: single fp16 plane).  Second-plane formats (`fmt`):
//   1  fp16x2: fp16 (value - hi) * 2048 at the same channel index
//   2  e4m3 pair: per 64-channel group 64 bytes e4m3(value / 8) followed by 64 bytes e4m3((value - hi) * 2048 / 8) -- the operand layout
//      of the conv kernel's fp16 + e4m3 mode (NBP_E4M3_ACT_SCALE); the value read back is hi + 8 * e4m3_lo / 2048 (~15 bits)
static constexpr float LO_SCALE = 2048.0f;

__device__ __forceinline__ uint32_t e4m3x2(float a, float b) {           // low byte = a; round to nearest even, saturating
    uint16_t r;
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %2, %1;" : "=h"(r) : "f"(a), "f"(b));
    return (uint32_t)r;
}
__device__ __forceinline__ float2 e4m3x2_to_float2(uint32_t v) {
    uint32_t h2;
    asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(h2) : "h"((uint16_t)v));
    return __half22float2(*reinterpret_cast<const __half2*>(&h2));
}

__device__ __forceinline__ void load8(const __half* pix, int ch, int lo, int fmt, float* f) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(pix + ch));
    const __half2* hq = reinterpret_cast<const __half2*>(&q);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 t = __half22float2(hq[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
    if (lo && fmt == 2) {
        const uint2 r = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(pix + lo) + (ch >> 6) * 128 + 64 + (ch & 63)));
        const uint32_t w[2] = {r.x, r.y};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 t = e4m3x2_to_float2((w[j >> 1] >> (16 * (j & 1))) & 0xffffu);
            f[2 * j] = fmaf(t.x, 1.0f / (LO_SCALE * NBP_E4M3_ACT_SCALE), f[2 * j]);
            f[2 * j + 1] = fmaf(t.y, 1.0f / (LO_SCALE * NBP_E4M3_ACT_SCALE), f[2 * j + 1]);
        }
    } else if (lo) {
        const uint4 r = __ldg(reinterpret_cast<const uint4*>(pix + ch + lo));
        const __half2* hr = reinterpret_cast<const __half2*>(&r);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 t = __half22float2(hr[j]);
            f[2 * j] = fmaf(t.x, 1.0f / LO_SCALE, f[2 * j]);
            f[2 * j + 1] = fmaf(t.y, 1.0f / LO_SCALE, f[2 * j + 1]);
        }
    }
}

__device__ __forceinline__ void store8(__half* pix, int ch, int lo, int fmt, const float* f) {
    uint32_t hi[4], lw[4], qh[4], ql[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float a0 = fminf(fmaxf(f[2 * j], -65504.0f), 65504.0f), a1 = fminf(fmaxf(f[2 * j + 1], -65504.0f), 65504.0f);
        const __half2 h = __floats2half2_rn(a0, a1);
        hi[j] = *reinterpret_cast<const uint32_t*>(&h);
        const float2 hf = __half22float2(h);
        const float r0 = (a0 - hf.x) * LO_SCALE, r1 = (a1 - hf.y) * LO_SCALE;
        const __half2 l = __floats2half2_rn(r0, r1);
        lw[j] = *reinterpret_cast<const uint32_t*>(&l);
        qh[j] = e4m3x2(a0 * NBP_E4M3_ACT_SCALE, a1 * NBP_E4M3_ACT_SCALE); ql[j] = e4m3x2(r0 * NBP_E4M3_ACT_SCALE, r1 * NBP_E4M3_ACT_SCALE);
    }
    *reinterpret_cast<uint4*>(pix + ch) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (lo && fmt == 2) {
        uint8_t* g = reinterpret_cast<uint8_t*>(pix + lo) + (ch >> 6) * 128 + (ch & 63);
        *reinterpret_cast<uint2*>(g) = make_uint2(qh[0] | (qh[1] << 16), qh[2] | (qh[3] << 16));
        *reinterpret_cast<uint2*>(g + 64) = make_uint2(ql[0] | (ql[1] << 16), ql[2] | (ql[3] << 16));
    } else if (lo) {
        *reinterpret_cast<uint4*>(pix + ch + lo) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    }
}

// ------------------------------------------------------------------------------------------------ conv_first
// One thread per pixel computes the 64 output channels (input loads coalesced across the warp, weights broadcast from shared memory,
// zero inputs skipped); the warp then transposes its 32 x 64 results through shared memory so that 8 lanes write the 128 contiguous
// bytes of one pixel and plane (round 1 wrote 16-byte pieces at a 256-byte stride: 32 partial-sector transactions per store
// instruction; an 8-lanes-per-pixel variant fixed the stores but multiplied the input loads by 8 and was 2.5x slower).
static constexpr int CF_ROW = 68;                 // floats per staged pixel row (64 + 4: 16-byte aligned, conflict-free float4 columns)

// CIN > 0: compile-time input channels (5 = the NBP model input); CIN = 0: run-time `cin` (one tap per load batch)
template <int COUT, int CIN>
__global__ void __launch_bounds__(128, 4) conv_first_kernel(const float* __restrict__ x, int n, int cin_rt, int h, int w,
                                                         const float* __restrict__ wt,      // [9*cin][COUT]
                                                         const float* __restrict__ scale, const float* __restrict__ shift,
                                                         __half* __restrict__ dst, int dst_ld, int dst_lo, int relu, int fmt) {
    static_assert(COUT == 64, "the staging tile is 64 channels wide");
    const int cin = CIN > 0 ? CIN : cin_rt;
    constexpr int CMAX = CIN > 0 ? CIN : 16;          // nbp_conv_first admits c_in <= 16
    constexpr int TXB = CIN > 0 ? 3 : 1;              // taps per load batch
    extern __shared__ float s_w[];                    // 9*cin*COUT weights, scale, shift, then 4 warps x 32 x CF_ROW staging
    const int nw = 9 * cin * COUT;
    for (int i = threadIdx.x; i < nw; i += blockDim.x) s_w[i] = wt[i];
    float* s_sc = s_w + nw; float* s_sh = s_sc + COUT;
    for (int i = threadIdx.x; i < COUT; i += blockDim.x) { s_sc[i] = scale[i]; s_sh[i] = shift[i]; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* tile = s_sh + COUT + warp * 32 * CF_ROW;
    __syncthreads();
    const size_t hw = (size_t)h * w;
    const size_t total = (size_t)n * hw;
    const size_t warp_stride = (size_t)gridDim.x * (blockDim.x >> 5) * 32;
    for (size_t base = ((size_t)blockIdx.x * (blockDim.x >> 5) + warp) * 32; base < total; base += warp_stride) {
        const size_t pix = base + lane;
        if (pix < total) {
            const int img = (int)(pix / hw);
            const int rem = (int)(pix - (size_t)img * hw);
            const int y = rem / w, xx = rem - y * w;
            float acc[COUT];
#pragma unroll
            for (int c = 0; c < COUT; ++c) acc[c] = 0.0f;
            const float* xin = x + (size_t)img * cin * hw;
            // TXB taps of one row at a time: their input values are loaded back to back (one exposed load latency per batch, not per
            // value: the kernel was bound by 45 dependent load -> FMA steps per pixel)
#pragma unroll 1
            for (int tb = 0; tb < 9; tb += TXB) {
                const int ty = tb / 3, tx0 = tb - 3 * ty;
                const int yy = y + ty - 1;
                const bool row_in = yy >= 0 && yy < h;
                const float* row = xin + (size_t)(row_in ? yy : 0) * w;
                float v[TXB][CMAX];
#pragma unroll
                for (int t = 0; t < TXB; ++t) {
                    const int xc = xx + tx0 + t - 1;
                    const bool in = row_in && xc >= 0 && xc < w;
#pragma unroll
                    for (int ci = 0; ci < CMAX; ++ci) v[t][ci] = (in && ci < cin) ? __ldg(row + (size_t)ci * hw + xc) : 0.0f;
                }
#pragma unroll
                for (int t = 0; t < TXB; ++t) {
#pragma unroll
                    for (int ci = 0; ci < CMAX; ++ci) {
                        if (ci >= cin) break;
                        const float vv = v[t][ci];
                        if (vv == 0.0f) continue;                         // count images are sparse
                        const float4* wr = reinterpret_cast<const float4*>(s_w + ((tb + t) * cin + ci) * COUT);
#pragma unroll
                        for (int c4 = 0; c4 < COUT / 4; ++c4) {
                            const float4 q = wr[c4];
                            acc[4 * c4 + 0] = fmaf(vv, q.x, acc[4 * c4 + 0]);
                            acc[4 * c4 + 1] = fmaf(vv, q.y, acc[4 * c4 + 1]);
                            acc[4 * c4 + 2] = fmaf(vv, q.z, acc[4 * c4 + 2]);
                            acc[4 * c4 + 3] = fmaf(vv, q.w, acc[4 * c4 + 3]);
                        }
                    }
                }
            }
#pragma unroll
            for (int c4 = 0; c4 < COUT / 4; ++c4) {
                float4 o;
                o.x = fmaf(acc[4 * c4 + 0], s_sc[4 * c4 + 0], s_sh[4 * c4 + 0]); o.y = fmaf(acc[4 * c4 + 1], s_sc[4 * c4 + 1], s_sh[4 * c4 + 1]);
                o.z = fmaf(acc[4 * c4 + 2], s_sc[4 * c4 + 2], s_sh[4 * c4 + 2]); o.w = fmaf(acc[4 * c4 + 3], s_sc[4 * c4 + 3], s_sh[4 * c4 + 3]);
                if (relu) { o.x = fmaxf(o.x, 0.0f); o.y = fmaxf(o.y, 0.0f); o.z = fmaxf(o.z, 0.0f); o.w = fmaxf(o.w, 0.0f); }
                *reinterpret_cast<float4*>(tile + lane * CF_ROW + 4 * c4) = o;
            }
        }
        __syncwarp();
        // 8 lanes per pixel, 4 pixels per pass: lane -> (pixel 4*it + lane/8, channel octet lane%8)
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int pl = 4 * it + (lane >> 3), c8 = lane & 7;
            const size_t pp = base + pl;
            if (pp < total) {
                const float4 a = *reinterpret_cast<const float4*>(tile + pl * CF_ROW + 8 * c8), b = *reinterpret_cast<const float4*>(tile + pl * CF_ROW + 8 * c8 + 4);
                const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                if (dst_lo < 0) {                     // plain fp32 destination (train mode: the raw pre-BatchNorm tensor), row stride dst_ld floats
                    float4* of = reinterpret_cast<float4*>(reinterpret_cast<float*>(dst) + pp * dst_ld + 8 * c8);
                    of[0] = a; of[1] = b;
                } else {
                    store8(dst + pp * dst_ld, 8 * c8, dst_lo, fmt, f);
                }
            }
        }
        __syncwarp();
    }
}

// The NBP stem itself (5 input channels, even image width): the generic kernel above is bound by FFMA issue (45 x 64 FFMA per pixel at
// one 3-register FFMA per two clocks and scheduler) and, behind that, by its 16 LDS.128 of weights per 64 FFMA.  Here a thread owns TWO
// horizontally adjacent pixels x 32 output channels (lanes 0-15: channels 0-31, lanes 16-31: channels 32-63 of the same 32 pixels):
// every weight float4 feeds four packed fp32x2 FMAs (`fma.rn.f32x2`, two channels per instruction), i.e. 8 LDS.128 + 32 FFMA2 per
// (tap, input channel) and pixel pair -- a quarter of the issue slots per pixel.  Every output channel still sums its 45 products in
// the same order with one fused multiply-add each, so the results are bit-identical to the generic kernel's.
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f32x2(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void ffma2(unsigned long long& acc, unsigned long long a, unsigned long long b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}

// shared-memory position of output channel c inside a 64-float weight row: the float4 k of channels 0-31 sits next to the float4 k of
// channels 32-63, so the two half-warps' broadcast reads of one LDS.128 fall into one 32-byte piece (no bank conflict)
__device__ __forceinline__ int stem_wpos(int c) { return ((c & 31) >> 2) * 8 + (c >> 5) * 4 + (c & 3); }

__global__ void __launch_bounds__(128, 4) conv_first5_kernel(const float* __restrict__ x, int n, int h, int w,
                                                            const float* __restrict__ wt,      // [45][64]
                                                            const float* __restrict__ scale, const float* __restrict__ shift,
                                                            __half* __restrict__ dst, int dst_ld, int dst_lo, int relu, int fmt) {
    constexpr int COUT = 64, CIN = 5;
    extern __shared__ float s_w[];                    // 45 x 64 weights (channel-permuted rows), scale, shift, then 4 warps x 32 x CF_ROW staging
    constexpr int nw = 9 * CIN * COUT;
    for (int i = threadIdx.x; i < nw; i += blockDim.x) s_w[(i & ~63) + stem_wpos(i & 63)] = wt[i];
    float* s_sc = s_w + nw; float* s_sh = s_sc + COUT;
    for (int i = threadIdx.x; i < COUT; i += blockDim.x) { s_sc[i] = scale[i]; s_sh[i] = shift[i]; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pr = lane & 15, half = lane >> 4;       // pixel pair of the warp's 32 pixels, channel half
    float* tile = s_sh + COUT + warp * 32 * CF_ROW;
    __syncthreads();
    const size_t hw = (size_t)h * w;
    const size_t total = (size_t)n * hw;              // even (w is)
    const size_t warp_stride = (size_t)gridDim.x * (blockDim.x >> 5) * 32;
    for (size_t base = ((size_t)blockIdx.x * (blockDim.x >> 5) + warp) * 32; base < total; base += warp_stride) {
        const size_t pix = base + 2 * pr;             // even pixel index -> even column, its right neighbour is in the same row
        if (pix < total) {
            const int img = (int)(pix / hw);
            const int rem = (int)(pix - (size_t)img * hw);
            const int y = rem / w, xx = rem - y * w;
            unsigned long long acc0[COUT / 4], acc1[COUT / 4];      // [pixel][16 channel pairs of this half]
#pragma unroll
            for (int c = 0; c < COUT / 4; ++c) { acc0[c] = 0ull; acc1[c] = 0ull; }
            const float* xin = x + (size_t)img * CIN * hw;
            const bool left = xx > 0, right = xx + 2 < w;
#pragma unroll 1
            for (int ty = 0; ty < 3; ++ty) {
                const int yy = y + ty - 1;
                if (yy < 0 || yy >= h) continue;                   // a padding row contributes exact zeros
                const float* row = xin + (size_t)yy * w + xx;
                float v[CIN][4];                                     // columns xx-1 .. xx+2 of the 5 input channels, loaded back to back
#pragma unroll
                for (int ci = 0; ci < CIN; ++ci) {
                    const float* q = row + (size_t)ci * hw;
                    v[ci][0] = left ? __ldg(q - 1) : 0.0f;
                    v[ci][1] = __ldg(q); v[ci][2] = __ldg(q + 1);
                    v[ci][3] = right ? __ldg(q + 2) : 0.0f;
                }
#pragma unroll
                for (int t = 0; t < 3; ++t) {
#pragma unroll
                    for (int ci = 0; ci < CIN; ++ci) {
                        const float v0 = v[ci][t], v1 = v[ci][t + 1];
                        if (v0 == 0.0f && v1 == 0.0f) continue;       // count images are sparse
                        const unsigned long long a0 = pack_f32x2(v0, v0), a1 = pack_f32x2(v1, v1);
                        const float4* wr = reinterpret_cast<const float4*>(s_w + ((ty * 3 + t) * CIN + ci) * COUT) + half;
#pragma unroll
                        for (int c4 = 0; c4 < COUT / 8; ++c4) {
                            const float4 q = wr[2 * c4];
                            const unsigned long long w01 = pack_f32x2(q.x, q.y), w23 = pack_f32x2(q.z, q.w);
                            ffma2(acc0[2 * c4], a0, w01); ffma2(acc0[2 * c4 + 1], a0, w23);
                            ffma2(acc1[2 * c4], a1, w01); ffma2(acc1[2 * c4 + 1], a1, w23);
                        }
                    }
                }
            }
            const float* sc = s_sc + half * 32; const float* sh = s_sh + half * 32;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
#pragma unroll
                for (int c4 = 0; c4 < COUT / 8; ++c4) {
                    float a[4];
                    unpack_f32x2(k ? acc1[2 * c4] : acc0[2 * c4], a[0], a[1]);
                    unpack_f32x2(k ? acc1[2 * c4 + 1] : acc0[2 * c4 + 1], a[2], a[3]);
                    float4 o;
                    o.x = fmaf(a[0], sc[4 * c4 + 0], sh[4 * c4 + 0]); o.y = fmaf(a[1], sc[4 * c4 + 1], sh[4 * c4 + 1]);
                    o.z = fmaf(a[2], sc[4 * c4 + 2], sh[4 * c4 + 2]); o.w = fmaf(a[3], sc[4 * c4 + 3], sh[4 * c4 + 3]);
                    if (relu) { o.x = fmaxf(o.x, 0.0f); o.y = fmaxf(o.y, 0.0f); o.z = fmaxf(o.z, 0.0f); o.w = fmaxf(o.w, 0.0f); }
                    *reinterpret_cast<float4*>(tile + (2 * pr + k) * CF_ROW + half * 32 + 4 * c4) = o;
                }
            }
        }
        __syncwarp();
        // 8 lanes per pixel, 4 pixels per pass: lane -> (pixel 4*it + lane/8, channel octet lane%8)
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int pl = 4 * it + (lane >> 3), c8 = lane & 7;
            const size_t pp = base + pl;
            if (pp < total) {
                const float4 a = *reinterpret_cast<const float4*>(tile + pl * CF_ROW + 8 * c8), b = *reinterpret_cast<const float4*>(tile + pl * CF_ROW + 8 * c8 + 4);
                const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                if (dst_lo < 0) {                     // plain fp32 destination (train mode: the raw pre-BatchNorm tensor), row stride dst_ld floats
                    float4* of = reinterpret_cast<float4*>(reinterpret_cast<float*>(dst) + pp * dst_ld + 8 * c8);
                    of[0] = a; of[1] = b;
                } else {
                    store8(dst + pp * dst_ld, 8 * c8, dst_lo, fmt, f);
                }
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------ pool / upsample
__global__ void __launch_bounds__(256) maxpool2x2_kernel(const __half* __restrict__ src, int n, int h, int w, int c, int ld_src, int lo_src,
                                                         __half* __restrict__ dst, int ld_dst, int lo_dst) {
    const int ho = h / 2, wo = w / 2, c8 = c / 8;
    const size_t total = (size_t)n * ho * wo * c8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int cc = (int)(i % c8); size_t t = i / c8;
        const int xo = (int)(t % wo); t /= wo;
        const int yo = (int)(t % ho); const int img = (int)(t / ho);
        const __half* p = src + (((size_t)img * h + 2 * yo) * w + 2 * xo) * ld_src;
        float a[8], b[8], d[8], e[8];
        load8(p, 8 * cc, lo_src, 1, a); load8(p + ld_src, 8 * cc, lo_src, 1, b);
        load8(p + (size_t)w * ld_src, 8 * cc, lo_src, 1, d); load8(p + (size_t)w * ld_src + ld_src, 8 * cc, lo_src, 1, e);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = fmaxf(fmaxf(a[j], b[j]), fmaxf(d[j], e[j]));
        store8(dst + (((size_t)img * ho + yo) * wo + xo) * ld_dst, 8 * cc, lo_dst, 1, a);    // re-split of the fp32 sum hi + lo/2048
    }
}

__global__ void __launch_bounds__(256) upsample2x_kernel(const __half* __restrict__ src, int n, int h, int w, int c, int ld_src, int lo_src,
                                                         __half* __restrict__ dst, int ld_dst, int lo_dst) {
    const int ho = 2 * h, wo = 2 * w, c8 = c / 8;
    const size_t total = (size_t)n * ho * wo * c8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int cc = (int)(i % c8); size_t t = i / c8;
        const int xo = (int)(t % wo); t /= wo;
        const int yo = (int)(t % ho); const int img = (int)(t / ho);
        const __half* sp = src + (((size_t)img * h + yo / 2) * w + xo / 2) * ld_src + 8 * cc;
        __half* dp = dst + (((size_t)img * ho + yo) * wo + xo) * ld_dst + 8 * cc;
        *reinterpret_cast<uint4*>(dp) = __ldg(reinterpret_cast<const uint4*>(sp));
        if (lo_src) *reinterpret_cast<uint4*>(dp + lo_dst) = __ldg(reinterpret_cast<const uint4*>(sp + lo_src));
    }
}

// ------------------------------------------------------------------------------------------------ attention gate
// a [P][f_int] (already ReLU'd), x [P][ld_x] (first f_l channels), out [P][ld_dst] channels [c_off, c_off+f_l)
// GS lanes cooperate on one pixel (GS = min(32, f_l/8)), 32/GS pixels per warp iteration.
__global__ void __launch_bounds__(256) att_gate_kernel(const __half* __restrict__ a, int f_int, int ld_a, int lo_a,
                                                       const __half* __restrict__ x, int f_l, int ld_x, int lo_x,
                                                       const float* __restrict__ w_psi, float psi_scale, float psi_shift,
                                                       __half* __restrict__ dst, int ld_dst, int c_off, int lo_dst, size_t npix, int gs, int fmt,
                                                       const float* __restrict__ psi_pre) {      // non-NULL: psi per pixel is given (f_int = 0)
    extern __shared__ float s_wp[];
    for (int i = threadIdx.x; i < f_int; i += blockDim.x) s_wp[i] = w_psi[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int ppw = 32 / gs;                         // pixels per warp iteration
    const int sub = lane / gs, gl = lane % gs;
    const size_t warp_global = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t warp_stride = ((size_t)gridDim.x * blockDim.x) >> 5;
    const size_t n_iter = (npix + ppw - 1) / ppw;
    for (size_t it = warp_global; it < n_iter; it += warp_stride) {
        const size_t pix = it * ppw + sub;
        const bool live = pix < npix;
        float dot = 0.0f;
        if (live) {
            const __half* ap = a + pix * ld_a;
            for (int ch = gl; ch < f_int / 8; ch += gs) {
                float f[8];
                load8(ap, 8 * ch, lo_a, 1, f);                    // `a` is never a GEMM operand: always the fp16 lo plane
#pragma unroll
                for (int j = 0; j < 8; ++j) dot = fmaf(f[j], s_wp[8 * ch + j], dot);
            }
        }
        for (int d = gs >> 1; d > 0; d >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, d);
        const float z = fmaf(dot, psi_scale, psi_shift);
        const float psi = psi_pre ? (live ? __ldg(psi_pre + pix) : 0.0f) : 1.0f / (1.0f + expf(-z));
        if (live) {
            const __half* xp = x + pix * ld_x;
            __half* op = dst + pix * ld_dst;
            for (int ch = gl; ch < f_l / 8; ch += gs) {
                float f[8];
                load8(xp, 8 * ch, lo_x, fmt, f);
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] *= psi;
                store8(op, c_off + 8 * ch, lo_dst, fmt, f);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ small-N 1x1 head
template <int COUT>
__global__ void __launch_bounds__(256) conv1x1_head_kernel(const __half* __restrict__ src, int c_in, int ld_src, int lo_src,
                                                           const float* __restrict__ wt,     // [COUT][c_in]
                                                           const float* __restrict__ bias, int sigmoid,
                                                           float* __restrict__ dst, float* __restrict__ dst_max, int n, size_t hw, int fmt) {
    extern __shared__ float s_w[];
    for (int i = threadIdx.x; i < COUT * c_in; i += blockDim.x) s_w[i] = wt[i];
    __syncthreads();
    const size_t total = (size_t)n * hw;
    for (size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += (size_t)gridDim.x * blockDim.x) {
        float acc[COUT];
#pragma unroll
        for (int c = 0; c < COUT; ++c) acc[c] = bias[c];
        const __half* sp = src + pix * ld_src;
        for (int ch = 0; ch < c_in / 8; ++ch) {
            float f[8];
            load8(sp, 8 * ch, lo_src, fmt, f);
#pragma unroll
            for (int c = 0; c < COUT; ++c) {
                const float* wr = s_w + c * c_in + 8 * ch;
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[c] = fmaf(f[j], wr[j], acc[c]);
            }
        }
        const size_t img = pix / hw, rem = pix - img * hw;
        float vmax = -INFINITY;
#pragma unroll
        for (int c = 0; c < COUT; ++c) {
            float v = acc[c];
            if (sigmoid) v = 1.0f / (1.0f + expf(-v));
            dst[(img * COUT + c) * hw + rem] = v;
            vmax = fmaxf(vmax, v);
        }
        if (dst_max) dst_max[pix] = vmax;                  // torch.max(predicted_value_map, dim=1) (nbp_planning.py:193)
    }
}

static int grid_for(size_t work_items, int threads, int per_thread = 1) {
    size_t g = (work_items + (size_t)threads * per_thread - 1) / ((size_t)threads * per_thread);
    if (g < 1) g = 1;
    if (g > 148 * 16) g = 148 * 16;
    return (int)g;
}

}  // namespace nbp

using namespace nbp;

static int check_plane(const char* who, int c, int ld, int lo) {
    if (lo < 0 || lo % 8 || (lo > 0 && lo < c) || ld < lo + c || ld % 8) return invalid("%s: bad plane layout c=%d ld=%d lo=%d", who, c, ld, lo);
    return NBP_OK;
}

static int check_fmt(const char* who, int fmt, int lo) {
    if (fmt != 1 && fmt != 2) return invalid("%s: fmt must be 1 (fp16 lo plane) or 2 (e4m3 pair plane), got %d", who, fmt);
    if (fmt == 2 && (lo <= 0 || lo % 64)) return invalid("%s: the e4m3 pair plane needs a 64-channel aligned second-plane offset (lo=%d)", who, lo);
    return NBP_OK;
}

extern "C" int nbp_conv_first(const float* x, int n, int c_in, int h, int w, const float* weight, const float* scale,
                              const float* shift, int c_out, int relu, void* dst, int dst_ld, int dst_lo, int fmt, void* stream) {
    if (!x || !weight || !scale || !shift || !dst) return invalid("nbp_conv_first: null pointer argument");
    if (n <= 0 || h <= 0 || w <= 0 || c_in <= 0 || c_in > 16) return invalid("nbp_conv_first: bad sizes n=%d c_in=%d h=%d w=%d", n, c_in, h, w);
    if (c_out != 64) return invalid("nbp_conv_first: c_out must be 64 (got %d)", c_out);
    int rc = dst_lo < 0 ? ((dst_ld < c_out || dst_ld % 4) ? invalid("nbp_conv_first: bad fp32 destination stride %d", dst_ld) : NBP_OK)
                        : check_plane("nbp_conv_first", c_out, dst_ld, dst_lo);
    if (rc) return rc;
    if ((uintptr_t)dst & 15) return invalid("nbp_conv_first: dst must be 16-byte aligned");
    if (dst_lo >= 0 && (rc = check_fmt("nbp_conv_first", fmt, fmt == 2 ? dst_lo : 64))) return rc;
    const size_t smem = sizeof(float) * (size_t)(9 * c_in * 64 + 128 + 4 * 32 * CF_ROW);
    if (smem > 48 * 1024) return invalid("nbp_conv_first: c_in=%d needs more than 48 KB of shared memory", c_in);
    const int g = grid_for((size_t)n * h * w, 128);
    static int pair_env = -1;                             // NBP_STEM_PAIR=0: A/B switch back to the one-pixel-per-thread kernel
    if (pair_env < 0) { const char* e = getenv("NBP_STEM_PAIR"); pair_env = e ? atoi(e) : 1; }
    if (c_in == 5 && !(w & 1) && pair_env) conv_first5_kernel<<<g, 128, smem, (cudaStream_t)stream>>>(x, n, h, w, weight, scale, shift, (__half*)dst, dst_ld, dst_lo, relu, fmt);
    else if (c_in == 5) conv_first_kernel<64, 5><<<g, 128, smem, (cudaStream_t)stream>>>(x, n, c_in, h, w, weight, scale, shift, (__half*)dst, dst_ld, dst_lo, relu, fmt);
    else conv_first_kernel<64, 0><<<g, 128, smem, (cudaStream_t)stream>>>(x, n, c_in, h, w, weight, scale, shift, (__half*)dst, dst_ld, dst_lo, relu, fmt);
    count_launch();
    return check_cuda(cudaGetLastError(), "nbp_conv_first launch");
}

static int check_nhwc(const char* who, const void* src, const void* dst, int n, int h, int w, int c, int ld_src, int lo_src, int ld_dst, int lo_dst) {
    if (!src || !dst) return invalid("%s: null pointer argument", who);
    if (n <= 0 || h <= 0 || w <= 0 || c <= 0 || c % 8) return invalid("%s: bad sizes n=%d h=%d w=%d c=%d (c must be a multiple of 8)", who, n, h, w, c);
    int rc = check_plane(who, c, ld_src, lo_src);
    if (rc) return rc;
    rc = check_plane(who, c, ld_dst, lo_dst);
    if (rc) return rc;
    if ((lo_src == 0) != (lo_dst == 0)) return invalid("%s: source and destination must use the same format", who);
    if (((uintptr_t)src | (uintptr_t)dst) & 15) return invalid("%s: pointers must be 16-byte aligned", who);
    return NBP_OK;
}

extern "C" int nbp_maxpool2x2(const void* src, int n, int h, int w, int c, int ld_src, int lo_src, void* dst, int ld_dst, int lo_dst, void* stream) {
    int rc = check_nhwc("nbp_maxpool2x2", src, dst, n, h, w, c, ld_src, lo_src, ld_dst, lo_dst);
    if (rc) return rc;
    if ((h | w) & 1) return invalid("nbp_maxpool2x2: h and w must be even");
    maxpool2x2_kernel<<<grid_for((size_t)n * (h / 2) * (w / 2) * (c / 8), 256), 256, 0, (cudaStream_t)stream>>>(
        (const __half*)src, n, h, w, c, ld_src, lo_src, (__half*)dst, ld_dst, lo_dst);
    count_launch();
    return check_cuda(cudaGetLastError(), "nbp_maxpool2x2 launch");
}

extern "C" int nbp_upsample2x(const void* src, int n, int h, int w, int c, int ld_src, int lo_src, void* dst, int ld_dst, int lo_dst, void* stream) {
    int rc = check_nhwc("nbp_upsample2x", src, dst, n, h, w, c, ld_src, lo_src, ld_dst, lo_dst);
    if (rc) return rc;
    upsample2x_kernel<<<grid_for((size_t)n * h * w * 4 * (c / 8), 256), 256, 0, (cudaStream_t)stream>>>(
        (const __half*)src, n, h, w, c, ld_src, lo_src, (__half*)dst, ld_dst, lo_dst);
    count_launch();
    return check_cuda(cudaGetLastError(), "nbp_upsample2x launch");
}

extern "C" int nbp_att_gate(const void* a, int f_int, int ld_a, int lo_a, const void* x, int f_l, int ld_x, int lo_x,
                            const float* w_psi, float psi_scale, float psi_shift,
                            void* dst, int dst_ld, int dst_c_off, int dst_lo, int64_t npix, int fmt, void* stream) {
    if (!a || !x || !w_psi || !dst) return invalid("nbp_att_gate: null pointer argument");
    if (int rf = check_fmt("nbp_att_gate", fmt, fmt == 2 ? (lo_x | dst_lo | (dst_c_off % 64 ? 1 : 0)) : 64)) return rf;
    if (f_int <= 0 || f_int % 8 || f_l <= 0 || f_l % 8 || npix <= 0) return invalid("nbp_att_gate: bad sizes f_int=%d f_l=%d npix=%lld", f_int, f_l, (long long)npix);
    int rc = check_plane("nbp_att_gate(a)", f_int, ld_a, lo_a);
    if (rc) return rc;
    rc = check_plane("nbp_att_gate(x)", f_l, ld_x, lo_x);
    if (rc) return rc;
    if (dst_ld % 8 || dst_c_off % 8 || dst_lo % 8 || dst_c_off + dst_lo + f_l > dst_ld) return invalid("nbp_att_gate: bad destination layout");
    if (((uintptr_t)a | (uintptr_t)x | (uintptr_t)dst) & 15) return invalid("nbp_att_gate: pointers must be 16-byte aligned");
    int gs = 1;
    while (gs * 2 <= 32 && gs * 2 <= f_l / 8) gs *= 2;
    const size_t warps = ((size_t)npix + (32 / gs) - 1) / (32 / gs);
    att_gate_kernel<<<grid_for(warps * 32, 256, 2), 256, sizeof(float) * f_int, (cudaStream_t)stream>>>(
        (const __half*)a, f_int, ld_a, lo_a, (const __half*)x, f_l, ld_x, lo_x, w_psi, psi_scale, psi_shift,
        (__half*)dst, dst_ld, dst_c_off, dst_lo, (size_t)npix, gs, fmt, nullptr);
    count_launch();
    return check_cuda(cudaGetLastError(), "nbp_att_gate launch");
}

extern "C" int nbp_att_scale(const float* psi, const void* x, int f_l, int ld_x, int lo_x,
                             void* dst, int dst_ld, int dst_c_off, int dst_lo, int64_t npix, int fmt, void* stream) {
    if (!psi || !x || !dst) return invalid("nbp_att_scale: null pointer argument");
    if (int rf = check_fmt("nbp_att_scale", fmt, fmt == 2 ? (lo_x | dst_lo | (dst_c_off % 64 ? 1 : 0)) : 64)) return rf;
    if (f_l <= 0 || f_l % 8 || npix <= 0) return invalid("nbp_att_scale: bad sizes f_l=%d npix=%lld", f_l, (long long)npix);
    int rc = check_plane("nbp_att_scale(x)", f_l, ld_x, lo_x);
    if (rc) return rc;
    if (dst_ld % 8 || dst_c_off % 8 || dst_lo % 8 || dst_c_off + dst_lo + f_l > dst_ld) return invalid("nbp_att_scale: bad destination layout");
    if (((uintptr_t)x | (uintptr_t)dst) & 15) return invalid("nbp_att_scale: pointers must be 16-byte aligned");
    int gs = 1;
    while (gs * 2 <= 32 && gs * 2 <= f_l / 8) gs *= 2;
    const size_t warps = ((size_t)npix + (32 / gs) - 1) / (32 / gs);
    att_gate_kernel<<<grid_for(warps * 32, 256, 2), 256, 0, (cudaStream_t)stream>>>(
        nullptr, 0, 0, 0, (const __half*)x, f_l, ld_x, lo_x, nullptr, 1.0f, 0.0f,
        (__half*)dst, dst_ld, dst_c_off, dst_lo, (size_t)npix, gs, fmt, psi);
    count_launch();
    return check_cuda(cudaGetLastError(), "nbp_att_scale launch");
}

extern "C" int nbp_conv1x1_head(const void* src, int c_in, int ld_src, int lo_src, const float* weight, const float* bias, int c_out,
                                int sigmoid, float* dst, float* dst_max, int n, int64_t hw, int fmt, void* stream) {
    if (!src || !weight || !bias || !dst) return invalid("nbp_conv1x1_head: null pointer argument");
    if (int rf = check_fmt("nbp_conv1x1_head", fmt, fmt == 2 ? lo_src : 64)) return rf;
    if (c_in <= 0 || c_in % 8 || n <= 0 || hw <= 0) return invalid("nbp_conv1x1_head: bad sizes");
    int rc = check_plane("nbp_conv1x1_head", c_in, ld_src, lo_src);
    if (rc) return rc;
    if ((uintptr_t)src & 15) return invalid("nbp_conv1x1_head: src must be 16-byte aligned");
    const size_t smem = sizeof(float) * (size_t)c_out * c_in;
    const int g = grid_for((size_t)n * hw, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (c_out == 8) conv1x1_head_kernel<8><<<g, 256, smem, st>>>((const __half*)src, c_in, ld_src, lo_src, weight, bias, sigmoid, dst, dst_max, n, (size_t)hw, fmt);
    else if (c_out == 1) conv1x1_head_kernel<1><<<g, 256, smem, st>>>((const __half*)src, c_in, ld_src, lo_src, weight, bias, sigmoid, dst, dst_max, n, (size_t)hw, fmt);
    else return invalid("nbp_conv1x1_head: c_out must be 1 or 8 (got %d)", c_out);
    count_launch();
    return check_cuda(cudaGetLastError(), "nbp_conv1x1_head launch");
}
