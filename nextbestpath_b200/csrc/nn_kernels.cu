// CUDA-core kernels around the tcgen05 convolution (the degenerate
// GEMM shapes K=45, N=1, N=8 and the pointwise glue of /root/reference/next_best_path/networks/nbp_model.py):
//   conv_first     Conv1.conv.0 (5->64, 3x3) + folded BN + ReLU, fp32 NCHW counts in, NHWC fp16 out   (:11-13)
//   maxpool2x2     nn.MaxPool2d(2,2)                                                              (:68)
//   upsample2x     nn.Upsample(scale_factor=2) (nearest)                                          (:27)
//   att_gate       psi = sigmoid(BN(conv1x1(a))) ; out = x * psi                                   (:49-62)
//   conv1x1_head   Final1 (256->8) / Final2 (64->1 + sigmoid), NHWC fp16 in, NCHW fp32 out         (:89,106-108)
// All are HBM-bound streaming kernels: 16-byte vector accesses, one pass over their inputs.
#include <cuda_fp16.h>

#include "nbp_common.cuh"

namespace nbp {

// ------------------------------------------------------------------------------------------------ conv_first
template <int COUT>
__global__ void __launch_bounds__(128) conv_first_kernel(const float* __restrict__ x, int n, int cin, int h, int w,
                                                         const float* __restrict__ wt,      // [9*cin][COUT]
                                                         const float* __restrict__ scale, const float* __restrict__ shift,
                                                         __half* __restrict__ dst, int dst_ld) {
    extern __shared__ float s_w[];                    // 9*cin*COUT weights, then scale, shift
    const int nw = 9 * cin * COUT;
    for (int i = threadIdx.x; i < nw; i += blockDim.x) s_w[i] = wt[i];
    float* s_sc = s_w + nw; float* s_sh = s_sc + COUT;
    for (int i = threadIdx.x; i < COUT; i += blockDim.x) { s_sc[i] = scale[i]; s_sh[i] = shift[i]; }
    __syncthreads();
    const size_t hw = (size_t)h * w;
    const size_t total = (size_t)n * hw;
    for (size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += (size_t)gridDim.x * blockDim.x) {
        const int img = (int)(pix / hw);
        const int rem = (int)(pix - (size_t)img * hw);
        const int y = rem / w, xx = rem - y * w;
        float acc[COUT];
#pragma unroll
        for (int c = 0; c < COUT; ++c) acc[c] = 0.0f;
        const float* xin = x + (size_t)img * cin * hw;
        for (int tap = 0; tap < 9; ++tap) {
            const int yy = y + tap / 3 - 1, xc = xx + tap % 3 - 1;
            if (yy < 0 || yy >= h || xc < 0 || xc >= w) continue;
            for (int ci = 0; ci < cin; ++ci) {
                const float v = __ldg(xin + (size_t)ci * hw + (size_t)yy * w + xc);
                if (v == 0.0f) continue;                         // count images are sparse
                const float4* wr = reinterpret_cast<const float4*>(s_w + (tap * cin + ci) * COUT);
#pragma unroll
                for (int c4 = 0; c4 < COUT / 4; ++c4) {
                    const float4 q = wr[c4];
                    acc[4 * c4 + 0] = fmaf(v, q.x, acc[4 * c4 + 0]);
                    acc[4 * c4 + 1] = fmaf(v, q.y, acc[4 * c4 + 1]);
                    acc[4 * c4 + 2] = fmaf(v, q.z, acc[4 * c4 + 2]);
                    acc[4 * c4 + 3] = fmaf(v, q.w, acc[4 * c4 + 3]);
                }
            }
        }
        uint4* o = reinterpret_cast<uint4*>(dst + pix * dst_ld);
#pragma unroll
        for (int c8 = 0; c8 < COUT / 8; ++c8) {
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = 8 * c8 + 2 * j;
                float a0 = fmaxf(fmaf(acc[c], s_sc[c], s_sh[c]), 0.0f);
                float a1 = fmaxf(fmaf(acc[c + 1], s_sc[c + 1], s_sh[c + 1]), 0.0f);
                const __half2 hh = __floats2half2_rn(fminf(a0, 65504.0f), fminf(a1, 65504.0f));
                pk[j] = *reinterpret_cast<const uint32_t*>(&hh);
            }
            o[c8] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------ pool / upsample
__device__ __forceinline__ uint4 hmax8(uint4 a, uint4 b) {
    uint4 r;
    const __half2* pa = reinterpret_cast<const __half2*>(&a); const __half2* pb = reinterpret_cast<const __half2*>(&b);
    __half2* pr = reinterpret_cast<__half2*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
    return r;
}

__global__ void __launch_bounds__(256) maxpool2x2_kernel(const __half* __restrict__ src, int n, int h, int w, int c, int ld_src,
                                                         __half* __restrict__ dst, int ld_dst) {
    const int ho = h / 2, wo = w / 2, c8 = c / 8;
    const size_t total = (size_t)n * ho * wo * c8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int cc = (int)(i % c8); size_t t = i / c8;
        const int xo = (int)(t % wo); t /= wo;
        const int yo = (int)(t % ho); const int img = (int)(t / ho);
        const __half* p = src + (((size_t)img * h + 2 * yo) * w + 2 * xo) * ld_src + 8 * cc;
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(p + ld_src));
        const uint4 d = __ldg(reinterpret_cast<const uint4*>(p + (size_t)w * ld_src));
        const uint4 e = __ldg(reinterpret_cast<const uint4*>(p + (size_t)w * ld_src + ld_src));
        *reinterpret_cast<uint4*>(dst + (((size_t)img * ho + yo) * wo + xo) * ld_dst + 8 * cc) = hmax8(hmax8(a, b), hmax8(d, e));
    }
}

__global__ void __launch_bounds__(256) upsample2x_kernel(const __half* __restrict__ src, int n, int h, int w, int c, int ld_src,
                                                         __half* __restrict__ dst, int ld_dst) {
    const int ho = 2 * h, wo = 2 * w, c8 = c / 8;
    const size_t total = (size_t)n * ho * wo * c8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int cc = (int)(i % c8); size_t t = i / c8;
        const int xo = (int)(t % wo); t /= wo;
        const int yo = (int)(t % ho); const int img = (int)(t / ho);
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + (((size_t)img * h + yo / 2) * w + xo / 2) * ld_src + 8 * cc));
        *reinterpret_cast<uint4*>(dst + (((size_t)img * ho + yo) * wo + xo) * ld_dst + 8 * cc) = v;
    }
}

// ------------------------------------------------------------------------------------------------ attention gate
// a [P][f_int] (already ReLU'd), x [P][ld_x] (first f_l channels), out [P][ld_dst] channels [c_off, c_off+f_l)
// GS lanes cooperate on one pixel (GS = min(32, f_l/8)), 32/GS pixels per warp iteration.
__global__ void __launch_bounds__(256) att_gate_kernel(const __half* __restrict__ a, int f_int, const __half* __restrict__ x, int f_l,
                                                       int ld_x, const float* __restrict__ w_psi, float psi_scale, float psi_shift,
                                                       __half* __restrict__ dst, int ld_dst, int c_off, size_t npix, int gs) {
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
            const uint4* ap = reinterpret_cast<const uint4*>(a + pix * f_int);
            for (int ch = gl; ch < f_int / 8; ch += gs) {
                const uint4 q = __ldg(ap + ch);
                const __half2* hq = reinterpret_cast<const __half2*>(&q);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __half22float2(hq[j]);
                    dot = fmaf(f.x, s_wp[8 * ch + 2 * j], dot);
                    dot = fmaf(f.y, s_wp[8 * ch + 2 * j + 1], dot);
                }
            }
        }
        for (int d = gs >> 1; d > 0; d >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, d);
        const float z = fmaf(dot, psi_scale, psi_shift);
        const float psi = 1.0f / (1.0f + __expf(-z));
        if (live) {
            const uint4* xp = reinterpret_cast<const uint4*>(x + pix * ld_x);
            uint4* op = reinterpret_cast<uint4*>(dst + pix * ld_dst + c_off);
            for (int ch = gl; ch < f_l / 8; ch += gs) {
                uint4 q = __ldg(xp + ch);
                __half2* hq = reinterpret_cast<__half2*>(&q);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __half22float2(hq[j]);
                    hq[j] = __floats2half2_rn(f.x * psi, f.y * psi);
                }
                op[ch] = q;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ small-N 1x1 head
template <int COUT>
__global__ void __launch_bounds__(256) conv1x1_head_kernel(const __half* __restrict__ src, int c_in, int ld_src,
                                                           const float* __restrict__ wt,     // [COUT][c_in]
                                                           const float* __restrict__ bias, int sigmoid,
                                                           float* __restrict__ dst, int n, size_t hw) {
    extern __shared__ float s_w[];
    for (int i = threadIdx.x; i < COUT * c_in; i += blockDim.x) s_w[i] = wt[i];
    __syncthreads();
    const size_t total = (size_t)n * hw;
    for (size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += (size_t)gridDim.x * blockDim.x) {
        float acc[COUT];
#pragma unroll
        for (int c = 0; c < COUT; ++c) acc[c] = bias[c];
        const uint4* sp = reinterpret_cast<const uint4*>(src + pix * ld_src);
        for (int ch = 0; ch < c_in / 8; ++ch) {
            const uint4 q = __ldg(sp + ch);
            const __half2* hq = reinterpret_cast<const __half2*>(&q);
            float f[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float2 t = __half22float2(hq[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
#pragma unroll
            for (int c = 0; c < COUT; ++c) {
                const float* wr = s_w + c * c_in + 8 * ch;
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[c] = fmaf(f[j], wr[j], acc[c]);
            }
        }
        const size_t img = pix / hw, rem = pix - img * hw;
#pragma unroll
        for (int c = 0; c < COUT; ++c) {
            float v = acc[c];
            if (sigmoid) v = 1.0f / (1.0f + __expf(-v));
            dst[(img * COUT + c) * hw + rem] = v;
        }
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

extern "C" int nbp_conv_first(const float* x, int n, int c_in, int h, int w, const float* weight, const float* scale,
                              const float* shift, int c_out, void* dst, int dst_ld, void* stream) {
    if (!x || !weight || !scale || !shift || !dst) return invalid("nbp_conv_first: null pointer argument");
    if (n <= 0 || h <= 0 || w <= 0 || c_in <= 0 || c_in > 16) return invalid("nbp_conv_first: bad sizes n=%d c_in=%d h=%d w=%d", n, c_in, h, w);
    if (c_out != 64) return invalid("nbp_conv_first: c_out must be 64 (got %d)", c_out);
    if (dst_ld < c_out || dst_ld % 8 || ((uintptr_t)dst & 15)) return invalid("nbp_conv_first: bad destination layout");
    const size_t smem = sizeof(float) * (size_t)(9 * c_in * 64 + 128);
    static bool attr = false;
    if (!attr) {
        int rc = check_cuda(cudaFuncSetAttribute(conv_first_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 9 * 16 * 64 * 4 + 512),
                            "cudaFuncSetAttribute(conv_first)");
        if (rc) return rc;
        attr = true;
    }
    conv_first_kernel<64><<<grid_for((size_t)n * h * w, 128), 128, smem, (cudaStream_t)stream>>>(x, n, c_in, h, w, weight, scale, shift,
                                                                                                  (__half*)dst, dst_ld);
    count_launch();
    return check_cuda(cudaGetLastError(), "nbp_conv_first launch");
}

static int check_nhwc(const char* who, const void* src, const void* dst, int n, int h, int w, int c, int ld_src, int ld_dst) {
    if (!src || !dst) return invalid("%s: null pointer argument", who);
    if (n <= 0 || h <= 0 || w <= 0 || c <= 0 || c % 8) return invalid("%s: bad sizes n=%d h=%d w=%d c=%d (c must be a multiple of 8)", who, n, h, w, c);
    if (ld_src < c || ld_dst < c || ld_src % 8 || ld_dst % 8) return invalid("%s: channel strides must be >= c and multiples of 8", who);
    if (((uintptr_t)src | (uintptr_t)dst) & 15) return invalid("%s: pointers must be 16-byte aligned", who);
    return NBP_OK;
}

extern "C" int nbp_maxpool2x2(const void* src, int n, int h, int w, int c, int ld_src, void* dst, int ld_dst, void* stream) {
    int rc = check_nhwc("nbp_maxpool2x2", src, dst, n, h, w, c, ld_src, ld_dst);
    if (rc) return rc;
    if ((h | w) & 1) return invalid("nbp_maxpool2x2: h and w must be even");
    maxpool2x2_kernel<<<grid_for((size_t)n * (h / 2) * (w / 2) * (c / 8), 256), 256, 0, (cudaStream_t)stream>>>(
        (const __half*)src, n, h, w, c, ld_src, (__half*)dst, ld_dst);
    count_launch();
    return check_cuda(cudaGetLastError(), "nbp_maxpool2x2 launch");
}

extern "C" int nbp_upsample2x(const void* src, int n, int h, int w, int c, int ld_src, void* dst, int ld_dst, void* stream) {
    int rc = check_nhwc("nbp_upsample2x", src, dst, n, h, w, c, ld_src, ld_dst);
    if (rc) return rc;
    upsample2x_kernel<<<grid_for((size_t)n * h * w * 4 * (c / 8), 256), 256, 0, (cudaStream_t)stream>>>(
        (const __half*)src, n, h, w, c, ld_src, (__half*)dst, ld_dst);
    count_launch();
    return check_cuda(cudaGetLastError(), "nbp_upsample2x launch");
}

extern "C" int nbp_att_gate(const void* a, int f_int, const void* x, int f_l, int ld_x, const float* w_psi, float psi_scale,
                            float psi_shift, void* dst, int dst_ld, int dst_c_off, int64_t npix, void* stream) {
    if (!a || !x || !w_psi || !dst) return invalid("nbp_att_gate: null pointer argument");
    if (f_int <= 0 || f_int % 8 || f_l <= 0 || f_l % 8 || npix <= 0) return invalid("nbp_att_gate: bad sizes f_int=%d f_l=%d npix=%lld", f_int, f_l, (long long)npix);
    if (ld_x < f_l || ld_x % 8 || dst_ld % 8 || dst_c_off % 8 || dst_c_off + f_l > dst_ld) return invalid("nbp_att_gate: bad channel layout");
    if (((uintptr_t)a | (uintptr_t)x | (uintptr_t)dst) & 15) return invalid("nbp_att_gate: pointers must be 16-byte aligned");
    int gs = 1;
    while (gs * 2 <= 32 && gs * 2 <= f_l / 8) gs *= 2;
    const size_t warps = ((size_t)npix + (32 / gs) - 1) / (32 / gs);
    att_gate_kernel<<<grid_for(warps * 32, 256, 2), 256, sizeof(float) * f_int, (cudaStream_t)stream>>>(
        (const __half*)a, f_int, (const __half*)x, f_l, ld_x, w_psi, psi_scale, psi_shift, (__half*)dst, dst_ld, dst_c_off, (size_t)npix, gs);
    count_launch();
    return check_cuda(cudaGetLastError(), "nbp_att_gate launch");
}

extern "C" int nbp_conv1x1_head(const void* src, int c_in, int ld_src, const float* weight, const float* bias, int c_out,
                                int sigmoid, float* dst, int n, int64_t hw, void* stream) {
    if (!src || !weight || !bias || !dst) return invalid("nbp_conv1x1_head: null pointer argument");
    if (c_in <= 0 || c_in % 8 || ld_src < c_in || ld_src % 8 || n <= 0 || hw <= 0) return invalid("nbp_conv1x1_head: bad sizes");
    if ((uintptr_t)src & 15) return invalid("nbp_conv1x1_head: src must be 16-byte aligned");
    const size_t smem = sizeof(float) * (size_t)c_out * c_in;
    const int g = grid_for((size_t)n * hw, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (c_out == 8) conv1x1_head_kernel<8><<<g, 256, smem, st>>>((const __half*)src, c_in, ld_src, weight, bias, sigmoid, dst, n, (size_t)hw);
    else if (c_out == 1) conv1x1_head_kernel<1><<<g, 256, smem, st>>>((const __half*)src, c_in, ld_src, weight, bias, sigmoid, dst, n, (size_t)hw);
    else return invalid("nbp_conv1x1_head: c_out must be 1 or 8 (got %d)", c_out);
    count_launch();
    return check_cuda(cudaGetLastError(), "nbp_conv1x1_head launch");
}
