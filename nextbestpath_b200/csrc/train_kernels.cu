// Train-mode kernels of the NBP network (SURVEY.md section 8 rows a11 train mode, a13):
// BatchNorm with batch statistics (forward + backward), ReLU / max-pool / nearest-upsample / attention-gate /
// head backward passes, operand preparation for the tensor-core dgrad / wgrad GEMMs.
// Reference semantics: torch.nn.BatchNorm2d in train mode (biased variance for normalisation, unbiased for the
// running estimate, momentum 0.1, eps 1e-5) as used by /root/reference/next_best_path/networks/nbp_model.py:8-62 and
// autograd through NBP.forward (:110-160) as driven by next_best_path/utility/nbp_utils.py:378-390.
//
// Formats: forward activations are NHWC fp16x2 "split" tensors (hi plane + lo*2048 plane, see conv_tc.cu);
// gradients flowing between layers are plain NHWC fp32; GEMM operands are produced from them by to_split_* with a
// per-tensor power-of-two scale (fp16 range), undone in the GEMM epilogue.
#include <cuda_fp16.h>

#include "nbp_common.cuh"

namespace nbp {

static constexpr float LO_SCALE = 2048.0f;

__device__ __forceinline__ void ld8(const __half* p, int lo, float* f) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
    const __half2* hq = reinterpret_cast<const __half2*>(&q);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 t = __half22float2(hq[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
    if (lo) {
        const uint4 r = __ldg(reinterpret_cast<const uint4*>(p + lo));
        const __half2* hr = reinterpret_cast<const __half2*>(&r);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 t = __half22float2(hr[j]);
            f[2 * j] = fmaf(t.x, 1.0f / LO_SCALE, f[2 * j]);
            f[2 * j + 1] = fmaf(t.y, 1.0f / LO_SCALE, f[2 * j + 1]);
        }
    }
}

__device__ __forceinline__ void st8(__half* p, int lo, const float* f) {
    uint32_t hi[4], lw[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float a0 = fminf(fmaxf(f[2 * j], -65504.0f), 65504.0f), a1 = fminf(fmaxf(f[2 * j + 1], -65504.0f), 65504.0f);
        const __half2 h = __floats2half2_rn(a0, a1);
        hi[j] = *reinterpret_cast<const uint32_t*>(&h);
        const float2 hf = __half22float2(h);
        const __half2 l = __floats2half2_rn((a0 - hf.x) * LO_SCALE, (a1 - hf.y) * LO_SCALE);
        lw[j] = *reinterpret_cast<const uint32_t*>(&l);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (lo) *reinterpret_cast<uint4*>(p + lo) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

__device__ __forceinline__ void ld8f(const float* p, float* f);
// The raw conv output z feeds BatchNorm (statistics, normalisation, backward).  It is stored either as a split fp16x2 tensor
// (lo >= 0) or -- train mode, same 4 bytes per element -- as plain fp32 (lo < 0, row stride ld floats): with 22 significant bits
// the cancellation in (z - mean) flips ReLU / max-pool decisions that the fp32 reference takes the other way, which is what
// dominated the whole-network gradient error (tests/studies/gradient_study.py).
__device__ __forceinline__ void ldz(const __half* z, size_t pix, int ld, int lo, int g, float* f) {
    if (lo < 0) ld8f(reinterpret_cast<const float*>(z) + pix * (size_t)ld + 8 * g, f);
    else ld8(z + pix * (size_t)ld + 8 * g, lo, f);
}

__device__ __forceinline__ void ld8f(const float* p, float* f) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void st8f(float* p, const float* f) {
    reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
}

// ---------------------------------------------------------------------------------------------------------------
// Per-channel reductions over pixels.  Block = 256 threads = (256 / G) pixel lanes x G channel groups of 8, G = C/8
// (C <= 2048).  Every thread accumulates NQ quantities for its 8 channels in fp32 over <= ~64 pixels, block-reduces
// through shared memory and adds to the global fp64 accumulators (so the cross-block sum is fp64).
template <int NQ, class F>
__device__ __forceinline__ void channel_reduce(size_t npix, int C, double* out /* [NQ][C] */, F&& body) {
    extern __shared__ float s_red[];                        // [lanes][NQ][C]
    const int G = C >> 3;
    const int lanes = blockDim.x / G;
    const int g = threadIdx.x % G, lane = threadIdx.x / G;
    float acc[NQ][8];
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[q][j] = 0.0f;
    if (lane < lanes)
        for (size_t p = (size_t)blockIdx.x * lanes + lane; p < npix; p += (size_t)gridDim.x * lanes) body(p, g, acc);
    if (lane < lanes) {
#pragma unroll
        for (int q = 0; q < NQ; ++q)
#pragma unroll
            for (int j = 0; j < 8; ++j) s_red[((size_t)lane * NQ + q) * C + 8 * g + j] = acc[q][j];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NQ * C; i += blockDim.x) {
        double t = 0.0;
        for (int l = 0; l < lanes; ++l) t += (double)s_red[(size_t)l * NQ * C + i];
        atomicAdd(out + i, t);
    }
}

// One pass over z: sum(z - s) -> acc[0][C] and sum((z - s)^2) -> acc[1][C] with the per-channel shift s = z[pixel 0] (the first pixel of
// the tensor: deterministic, and a value of the distribution, so |mean - s| is a few standard deviations at most and the cancellation
// in var = E[(z-s)^2] - (mean - s)^2 costs a factor (1 + (mean-s)^2/var) on a 1e-7 relative error; partial sums: fp32 over <= ~64
// pixels per thread, fp64 across threads and blocks).  Round 2 replaced the two sweeps (mean, then squared deviations) by this one.
__global__ void __launch_bounds__(256) bn_moments_kernel(const __half* __restrict__ z, int ld, int lo, size_t npix, int C, double* acc) {
    float s[8];
    ldz(z, 0, ld, lo, threadIdx.x % (C >> 3), s);           // this thread's channel group (as in channel_reduce)
    channel_reduce<2>(npix, C, acc, [&](size_t p, int g, float (*a)[8]) {
        float f[8];
        ldz(z, p, ld, lo, g, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = f[j] - s[j]; a[0][j] += d; a[1][j] = fmaf(d, d, a[1][j]); }
    });
}

__device__ __forceinline__ float z_elem(const __half* z, int ld, int lo, int c) {          // element c of pixel 0
    if (lo < 0) return reinterpret_cast<const float*>(z)[c];
    return __half2float(z[c]) + __half2float(z[c + lo]) * (1.0f / LO_SCALE);
}

// batch mean / biased var -> (scale, shift) of the normalisation, invstd, and the running-statistics update
__global__ void bn_finalize_kernel(const double* acc, size_t npix, int C, const float* gamma, const float* beta,
                                   float* running_mean, float* running_var, float momentum, float eps,
                                   float* mean, float* invstd, float* scale, float* shift,
                                   const __half* z0, int ld, int lo) {       // z0 != nullptr: acc holds the moments about z[pixel 0]
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double m = acc[c] / (double)npix, ss = acc[C + c];
    if (z0) {
        ss = ss - m * m * (double)npix;                      // sum of squared deviations about the mean
        if (ss < 0.0) ss = 0.0;
        m += (double)z_elem(z0, ld, lo, c);
    }
    const double var = ss / (double)npix;
    const float is = (float)(1.0 / sqrt(var + (double)eps));
    mean[c] = (float)m; invstd[c] = is;
    const float sc = gamma[c] * is;
    scale[c] = sc; shift[c] = beta[c] - (float)m * sc;
    if (running_mean) {
        const double unbiased = npix > 1 ? ss / (double)(npix - 1) : var;
        running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * (float)m;
        running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}

// y = act(z * scale + shift)
__global__ void __launch_bounds__(256) affine_act_kernel(const __half* __restrict__ z, int ld_z, int lo_z, size_t npix, int C,
                                                         const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                                                         __half* __restrict__ y, int ld_y, int lo_y) {
    const int G = C >> 3;
    const size_t total = npix * G;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t p = i / G; const int g = (int)(i - p * G);
        float f[8];
        ldz(z, p, ld_z, lo_z, g, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) { f[j] = fmaf(f[j], scale[8 * g + j], shift[8 * g + j]); if (relu) f[j] = fmaxf(f[j], 0.0f); }
        st8(y + p * ld_y + 8 * g, lo_y, f);
    }
}

// a = relu(zg*sg+tg + zx*sx+tx)   (Attention_block nbp_model.py:57-59)
__global__ void __launch_bounds__(256) att_pre_kernel(const __half* __restrict__ zg, const __half* __restrict__ zx, int ld, int lo, size_t npix, int C,
                                                      const float* sg, const float* tg, const float* sx, const float* tx,
                                                      __half* __restrict__ a, int ld_a, int lo_a) {
    const int G = C >> 3;
    const size_t total = npix * G;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t p = i / G; const int g = (int)(i - p * G);
        float u[8], v[8];
        ldz(zg, p, ld, lo, g, u); ldz(zx, p, ld, lo, g, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) u[j] = fmaxf(fmaf(u[j], sg[8 * g + j], tg[8 * g + j]) + fmaf(v[j], sx[8 * g + j], tx[8 * g + j]), 0.0f);
        st8(a + p * ld_a + 8 * g, lo_a, u);
    }
}

// zpsi[p] = dot(a[p], w) + b : one warp-group of GS lanes per pixel
__global__ void __launch_bounds__(256) psi_dot_kernel(const __half* __restrict__ a, int ld, int lo, size_t npix, int C, const float* __restrict__ w,
                                                      const float* __restrict__ bias, float* __restrict__ zpsi, int gs) {
    const int lane = threadIdx.x & 31, sub = lane / gs, gl = lane % gs, ppw = 32 / gs;
    const size_t warp_global = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, warp_stride = ((size_t)gridDim.x * blockDim.x) >> 5;
    const size_t n_iter = (npix + ppw - 1) / ppw;
    for (size_t it = warp_global; it < n_iter; it += warp_stride) {
        const size_t p = it * ppw + sub;
        float dot = 0.0f;
        if (p < npix)
            for (int ch = gl; ch < C / 8; ch += gs) {
                float f[8];
                ld8(a + p * ld + 8 * ch, lo, f);
#pragma unroll
                for (int j = 0; j < 8; ++j) dot = fmaf(f[j], __ldg(w + 8 * ch + j), dot);
            }
        for (int d = gs >> 1; d > 0; d >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, d);
        if (p < npix && gl == 0) zpsi[p] = dot + bias[0];
    }
}

// scalar-channel statistics (the psi BatchNorm2d(1)): acc[0] = sum, acc[1] = sum sq dev
__global__ void __launch_bounds__(256) stats1_kernel(const float* __restrict__ v, size_t n, double* acc, int phase) {
    __shared__ double s[8];
    const double mean = phase ? acc[0] / (double)n : 0.0;
    double t = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double x = (double)v[i];
        t += phase ? (x - mean) * (x - mean) : x;
    }
    for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
    if (lane_id() == 0) s[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) { double r = 0; for (int w = 0; w < 8; ++w) r += s[w]; atomicAdd(acc + phase, r); }
}

// psi = sigmoid(zpsi*scale+shift) (device scalars); out = x * psi
__global__ void __launch_bounds__(256) att_apply_kernel(const float* __restrict__ zpsi, const float* __restrict__ ps, const float* __restrict__ pt,
                                                        const __half* __restrict__ x, int ld_x, int lo_x, size_t npix, int C,
                                                        __half* __restrict__ dst, int ld_d, int c_off, int lo_d, float* __restrict__ psi_out) {
    const int G = C >> 3;
    const size_t total = npix * G;
    const float s = ps[0], t = pt[0];
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t p = i / G; const int g = (int)(i - p * G);
        const float psi = 1.0f / (1.0f + expf(-fmaf(zpsi[p], s, t)));
        if (g == 0 && psi_out) psi_out[p] = psi;
        float f[8];
        ld8(x + p * ld_x + 8 * g, lo_x, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] *= psi;
        st8(dst + p * ld_d + c_off + 8 * g, lo_d, f);
    }
}

__device__ __forceinline__ void atomic_max_abs(float* addr, float v) {       // non-negative floats order like uints
    atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(fabsf(v)));
}

// ---------------------------------------------------------------------------------------------------------------
// BatchNorm(+ReLU) backward.  y = act(z*scale+shift), xhat = (z-mean)*invstd, dy_m = dy * (y > 0 if relu)
//   s1 = sum dy_m, s2 = sum dy_m*xhat ;  dz = gamma*invstd*(dy_m - s1/N - xhat*s2/N) ; dgamma = s2 ; dbeta = s1
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float* __restrict__ dy, int ld_dy, const __half* __restrict__ z, int ld_z, int lo_z,
                                                            size_t npix, int C, const float* __restrict__ scale, const float* __restrict__ shift,
                                                            const float* __restrict__ mean, const float* __restrict__ invstd, int relu, double* acc,
                                                            float* amax2 /* optional: [0] = max|dy_m|, [1] = max|xhat| */) {
    float mx_d = 0.0f, mx_x = 0.0f;
    // per-channel (scale, shift, mean, invstd) staged in shared memory behind channel_reduce's scratch ([lanes][2][C])
    extern __shared__ float s_red[];
    float* s_co = s_red + (size_t)(blockDim.x / (C >> 3)) * 2 * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) { s_co[c] = scale[c]; s_co[C + c] = shift[c]; s_co[2 * C + c] = mean[c]; s_co[3 * C + c] = invstd[c]; }
    __syncthreads();
    channel_reduce<2>(npix, C, acc, [&](size_t p, int g, float (*a)[8]) {
        float f[8], d[8], co[4][8];
        ldz(z, p, ld_z, lo_z, g, f);
        ld8f(dy + p * ld_dy + 8 * g, d);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 u = *reinterpret_cast<const float4*>(s_co + q * C + 8 * g), v = *reinterpret_cast<const float4*>(s_co + q * C + 8 * g + 4);
            co[q][0] = u.x; co[q][1] = u.y; co[q][2] = u.z; co[q][3] = u.w; co[q][4] = v.x; co[q][5] = v.y; co[q][6] = v.z; co[q][7] = v.w;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float y = fmaf(f[j], co[0][j], co[1][j]);
            const float dm = (relu && !(y > 0.0f)) ? 0.0f : d[j];
            const float xh = (f[j] - co[2][j]) * co[3][j];
            a[0][j] += dm;
            a[1][j] = fmaf(dm, xh, a[1][j]);
            mx_d = fmaxf(mx_d, fabsf(dm)); mx_x = fmaxf(mx_x, fabsf(xh));
        }
    });
    if (amax2) {
        for (int d = 16; d > 0; d >>= 1) { mx_d = fmaxf(mx_d, __shfl_xor_sync(0xffffffffu, mx_d, d)); mx_x = fmaxf(mx_x, __shfl_xor_sync(0xffffffffu, mx_x, d)); }
        if (lane_id() == 0) { atomic_max_abs(amax2, mx_d); atomic_max_abs(amax2 + 1, mx_x); }
    }
}

__device__ __forceinline__ float pow2_scale_for(float amax) {
    if (!(amax > 0.0f) || !isfinite(amax)) return 1.0f;
    int e; frexpf(amax, &e);                    // amax = m * 2^e, m in [0.5, 1)
    return ldexpf(1.0f, 8 - e);                 // amax * scale in [128, 256)
}

// Split mode (dzs != nullptr): dz goes straight out as the fp16x2 GEMM operand of dgrad / wgrad, scaled by a power of two
// derived from an upper BOUND of max|dz| (from the reduce pass: gamma*invstd*(max|dy_m| + |s1|/N + max|xhat|*|s2|/N)),
// so no fp32 dz tensor and no separate re-split pass exist.
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ dy, int ld_dy, const __half* __restrict__ z, int ld_z, int lo_z,
                                                           size_t npix, int C, const float* __restrict__ scale, const float* __restrict__ shift,
                                                           const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                           int relu, const double* __restrict__ acc, float* __restrict__ dz, int ld_dz, float* amax,
                                                           float* dgamma, float* dbeta,
                                                           __half* __restrict__ dzs, int ld_s, int lo_s, const float* __restrict__ amax2,
                                                           float* inv_scale_vec, int n_vec) {
    const int G = C >> 3;
    const size_t total = npix * G;
    const float invn = 1.0f / (float)npix;
    float local_max = 0.0f;
    float sc = 1.0f;
    if (dzs) {
        __shared__ float s_b[8];
        float b = 0.0f;
        const float md = amax2[0], mxh = amax2[1];
        for (int c = threadIdx.x; c < C; c += blockDim.x)
            b = fmaxf(b, fabsf(gamma[c]) * invstd[c] * (md + fabsf((float)acc[c]) * invn + mxh * fabsf((float)acc[C + c]) * invn));
        for (int d = 16; d > 0; d >>= 1) b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, d));
        if (lane_id() == 0) s_b[threadIdx.x >> 5] = b;
        __syncthreads();
        b = 0.0f;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) b = fmaxf(b, s_b[i]);
        sc = pow2_scale_for(b);
        if (blockIdx.x == 0 && inv_scale_vec) for (int i = threadIdx.x; i < n_vec; i += blockDim.x) inv_scale_vec[i] = 1.0f / sc;
    }
    // per-channel coefficients staged once per block in shared memory ([7][C]: scale, shift, mean, invstd, k = gamma*invstd,
    // a1 = s1/N, a2 = s2/N): the element loop then issues 16-byte LDS instead of 7 scalar global loads per channel
    extern __shared__ float s_co[];
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        s_co[c] = scale[c]; s_co[C + c] = shift[c]; s_co[2 * C + c] = mean[c]; s_co[3 * C + c] = invstd[c];
        s_co[4 * C + c] = gamma[c] * invstd[c]; s_co[5 * C + c] = (float)acc[c] * invn; s_co[6 * C + c] = (float)acc[C + c] * invn;
    }
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t p = i / G; const int g = (int)(i - p * G);
        float f[8], d[8], o[8], co[7][8];
        ldz(z, p, ld_z, lo_z, g, f);
        ld8f(dy + p * ld_dy + 8 * g, d);
#pragma unroll
        for (int q = 0; q < 7; ++q) {
            const float4 u = *reinterpret_cast<const float4*>(s_co + q * C + 8 * g), v = *reinterpret_cast<const float4*>(s_co + q * C + 8 * g + 4);
            co[q][0] = u.x; co[q][1] = u.y; co[q][2] = u.z; co[q][3] = u.w; co[q][4] = v.x; co[q][5] = v.y; co[q][6] = v.z; co[q][7] = v.w;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float y = fmaf(f[j], co[0][j], co[1][j]);
            const float dm = (relu && !(y > 0.0f)) ? 0.0f : d[j];
            const float xh = (f[j] - co[2][j]) * co[3][j];
            o[j] = co[4][j] * (dm - co[5][j] - xh * co[6][j]);
            local_max = fmaxf(local_max, fabsf(o[j]));
        }
        if (dzs) {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] *= sc;
            st8(dzs + p * ld_s + 8 * g, lo_s, o);
        } else {
            st8f(dz + p * ld_dz + 8 * g, o);
        }
    }
    for (int d = 16; d > 0; d >>= 1) local_max = fmaxf(local_max, __shfl_xor_sync(0xffffffffu, local_max, d));
    if (lane_id() == 0 && amax) atomic_max_abs(amax, local_max);
    if (blockIdx.x == 0 && dgamma)
        for (int c = threadIdx.x; c < C; c += blockDim.x) { dgamma[c] += (float)acc[C + c]; dbeta[c] += (float)acc[c]; }
}

// ---------------------------------------------------------------------------------------------------------------
// GEMM operand preparation: fp32 NHWC gradient -> fp16x2 split, scaled by 2^k so that amax lands near 2^8.
// *inv_scale_out (device) receives 2^-k for the GEMM epilogue.  Two layouts: NHWC (dgrad A operand) and channel-major
// CNHW [C][n][h][w] (wgrad operands, K = pixels contiguous).
__global__ void __launch_bounds__(256) to_split_nhwc_kernel(const float* __restrict__ src, int ld_s, size_t npix, int C, const float* __restrict__ amax,
                                                            __half* __restrict__ dst, int ld_d, int lo_d, float* inv_scale_vec, int n_vec) {
    const float sc = amax ? pow2_scale_for(amax[0]) : 1.0f;
    if (blockIdx.x == 0 && inv_scale_vec) for (int i = threadIdx.x; i < n_vec; i += blockDim.x) inv_scale_vec[i] = 1.0f / sc;
    const int G = C >> 3;
    const size_t total = npix * G;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t p = i / G; const int g = (int)(i - p * G);
        float f[8];
        ld8f(src + p * ld_s + 8 * g, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] *= sc;
        st8(dst + p * ld_d + 8 * g, lo_d, f);
    }
}

// transpose to channel-major through a 32x32 shared tile.  src is either fp32 NHWC (src_f) or split NHWC (src_h).
// dst planes: hi [C][npix_pad], lo [C][npix_pad] (lo plane `plane_stride` elements after hi).
__global__ void __launch_bounds__(256) to_split_cnhw_kernel(const float* __restrict__ src_f, const __half* __restrict__ src_h, int ld_s, int lo_s,
                                                            size_t npix, int w, int wp, int dx, int C, const float* __restrict__ amax, __half* __restrict__ dst,
                                                            size_t row_stride, size_t plane_stride, float* inv_scale_out) {
    __shared__ float tile[32][33];
    const float sc = amax ? pow2_scale_for(amax[0]) : 1.0f;
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && inv_scale_out) inv_scale_out[0] = 1.0f / sc;
    const size_t p0 = (size_t)blockIdx.x * 32; const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const size_t p = p0 + r; const int c = c0 + tx;
        float v = 0.0f;
        if (p < npix && c < C) {
            if (src_f) v = src_f[p * ld_s + c] * sc;
            else { v = __half2float(src_h[p * ld_s + c]); if (lo_s) v += __half2float(src_h[p * ld_s + lo_s + c]) * (1.0f / LO_SCALE); }
        }
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r; const size_t p = p0 + tx;
        if (c < C && p < npix) {
            const float v = fminf(fmaxf(tile[tx][r], -65504.0f), 65504.0f);
            const __half h = __float2half_rn(v);
            // rows are padded from w to wp pixels; dst[.., x'] = src[.., x' + dx] (TMA cannot shift the innermost
            // coordinate by less than 16 bytes, so x-shifted copies are materialised); unwritten entries stay zero
            const int xs = (int)(p % w) - dx;
            if (xs >= 0 && xs < w) {
                const size_t q = (p / w) * wp + xs;
                dst[(size_t)c * row_stride + q] = h;
                dst[plane_stride + (size_t)c * row_stride + q] = __float2half_rn((v - __half2float(h)) * LO_SCALE);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// pooling / upsampling backward, fp32 NHWC
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const float* __restrict__ dy, const __half* __restrict__ x, int ld_x, int lo_x,
                                                          int n, int h, int w, int C, float* __restrict__ dx, int accumulate) {
    const int ho = h / 2, wo = w / 2, G = C >> 3;
    const size_t total = (size_t)n * ho * wo * G;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int g = (int)(i % G); size_t t = i / G;
        const int xo = (int)(t % wo); t /= wo;
        const int yo = (int)(t % ho); const int img = (int)(t / ho);
        float d[8], v[4][8];
        ld8f(dy + (((size_t)img * ho + yo) * wo + xo) * C + 8 * g, d);
        const size_t base = ((size_t)img * h + 2 * yo) * w + 2 * xo;
        const size_t off[4] = {base, base + 1, base + w, base + w + 1};
#pragma unroll
        for (int k = 0; k < 4; ++k) ld8(x + off[k] * ld_x + 8 * g, lo_x, v[k]);
        float o[4][8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int best = 0;                                        // first maximum in window order, as torch
#pragma unroll
            for (int k = 1; k < 4; ++k) if (v[k][j] > v[best][j]) best = k;
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k][j] = (k == best) ? d[j] : 0.0f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float* q = dx + off[k] * C + 8 * g;
            if (accumulate) { float c[8]; ld8f(q, c);
#pragma unroll
                for (int j = 0; j < 8; ++j) o[k][j] += c[j]; }
            st8f(q, o[k]);
        }
    }
}

__global__ void __launch_bounds__(256) upsample_bwd_kernel(const float* __restrict__ dy, int n, int h, int w, int C, float* __restrict__ dx, int accumulate) {
    const int G = C >> 3;
    const size_t total = (size_t)n * h * w * G;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int g = (int)(i % G); size_t t = i / G;
        const int x = (int)(t % w); t /= w;
        const int y = (int)(t % h); const int img = (int)(t / h);
        const size_t base = ((size_t)img * 2 * h + 2 * y) * (2 * w) + 2 * x;
        float a[8], b[8], c[8], d[8];
        ld8f(dy + base * C + 8 * g, a); ld8f(dy + (base + 1) * C + 8 * g, b);
        ld8f(dy + (base + 2 * w) * C + 8 * g, c); ld8f(dy + (base + 2 * w + 1) * C + 8 * g, d);
        float* q = dx + (((size_t)img * h + y) * w + x) * C + 8 * g;
        float o[8];
        if (accumulate) ld8f(q, o);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (accumulate ? o[j] : 0.0f) + ((a[j] + b[j]) + (c[j] + d[j]));
        st8f(q, o);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// attention gate backward, part 1: out = x*psi (psi per pixel)
//   dx_skip += dout*psi ; dt[p] = (sum_c dout*x) * psi*(1-psi)       (t = BN1(zpsi), psi = sigmoid(t))
__global__ void __launch_bounds__(256) att_apply_bwd_kernel(const float* __restrict__ dout, int ld_do, const __half* __restrict__ x, int ld_x, int lo_x,
                                                            const float* __restrict__ psi, size_t npix, int C, float* __restrict__ dx, int accumulate,
                                                            float* __restrict__ dt, int gs) {
    const int lane = threadIdx.x & 31, sub = lane / gs, gl = lane % gs, ppw = 32 / gs;
    const size_t warp_global = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, warp_stride = ((size_t)gridDim.x * blockDim.x) >> 5;
    const size_t n_iter = (npix + ppw - 1) / ppw;
    for (size_t it = warp_global; it < n_iter; it += warp_stride) {
        const size_t p = it * ppw + sub;
        const bool live = p < npix;
        const float ps = live ? psi[p] : 0.0f;
        float dot = 0.0f;
        if (live)
            for (int ch = gl; ch < C / 8; ch += gs) {
                float f[8], d[8], o[8];
                ld8(x + p * ld_x + 8 * ch, lo_x, f);
                ld8f(dout + p * ld_do + 8 * ch, d);
                float* q = dx + p * C + 8 * ch;
                if (accumulate) ld8f(q, o);
#pragma unroll
                for (int j = 0; j < 8; ++j) { dot = fmaf(d[j], f[j], dot); o[j] = (accumulate ? o[j] : 0.0f) + d[j] * ps; }
                st8f(q, o);
            }
        for (int d = gs >> 1; d > 0; d >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, d);
        if (live && gl == 0) dt[p] = dot * ps * (1.0f - ps);
    }
}

// scalar BN backward reductions for psi: acc[0] = sum dt, acc[1] = sum dt*xhat
__global__ void __launch_bounds__(256) bn1_bwd_reduce_kernel(const float* __restrict__ dt, const float* __restrict__ zpsi, size_t n,
                                                             const float* __restrict__ mean, const float* __restrict__ invstd, double* acc) {
    __shared__ double s[2][8];
    double a = 0.0, b = 0.0;
    const float m = mean[0], is = invstd[0];
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double d = dt[i];
        a += d; b += d * (double)((zpsi[i] - m) * is);
    }
    for (int d = 16; d > 0; d >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, d); b += __shfl_xor_sync(0xffffffffu, b, d); }
    if (lane_id() == 0) { s[0][threadIdx.x >> 5] = a; s[1][threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) { double ra = 0, rb = 0; for (int w = 0; w < 8; ++w) { ra += s[0][w]; rb += s[1][w]; } atomicAdd(acc, ra); atomicAdd(acc + 1, rb); }
}

// attention gate backward, part 2: dzpsi = gamma1*invstd1*(dt - s1/N - xhat*s2/N); d_pre = dzpsi * w_psi * (a > 0);
// dw_psi += sum_p dzpsi * a ; dgamma1/dbeta1 from acc
__global__ void __launch_bounds__(256) psi_bwd_kernel(const float* __restrict__ dt, const float* __restrict__ zpsi, size_t npix,
                                                      const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma1,
                                                      const double* __restrict__ acc, const __half* __restrict__ a, int ld_a, int lo_a, int C,
                                                      const float* __restrict__ w_psi, float* __restrict__ dpre, double* dw_acc /* [C] */,
                                                      float* dgamma1, float* dbeta1) {
    channel_reduce<1>(npix, C, dw_acc, [&](size_t p, int g, float (*r)[8]) {
        const float xh = (zpsi[p] - mean[0]) * invstd[0];
        const float invn = 1.0f / (float)npix;
        const float dz = gamma1[0] * invstd[0] * (dt[p] - (float)acc[0] * invn - xh * (float)acc[1] * invn);
        float f[8], o[8];
        ld8(a + p * ld_a + 8 * g, lo_a, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            r[0][j] = fmaf(dz, f[j], r[0][j]);
            o[j] = f[j] > 0.0f ? dz * w_psi[8 * g + j] : 0.0f;
        }
        st8f(dpre + p * C + 8 * g, o);
    });
    if (blockIdx.x == 0 && threadIdx.x == 0) { dgamma1[0] += (float)acc[1]; dbeta1[0] += (float)acc[0]; }
}

// ---------------------------------------------------------------------------------------------------------------
// heads backward: out[o] = act(sum_c W[o][c] d[c] + b[o]).  dD (fp32 NHWC), dW, db.
template <int COUT>
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out_sig /* nullptr: linear */,
                                                       const __half* __restrict__ src, int c_in, int ld_s, int lo_s, const float* __restrict__ wt,
                                                       int n, size_t hw, float* __restrict__ dsrc, double* dw_acc /* [COUT][c_in] then [COUT] */) {
    extern __shared__ float s_w[];                 // [COUT][c_in]
    for (int i = threadIdx.x; i < COUT * c_in; i += blockDim.x) s_w[i] = wt[i];
    __syncthreads();
    const size_t npix = (size_t)n * hw;
    // channel groups of 8 over c_in; lanes over pixels.  dW[o][c] = sum_p dpre[o][p]*d[p][c]
    const int G = c_in >> 3, lanes = blockDim.x / G, g = threadIdx.x % G, lane = threadIdx.x / G;
    float acc[COUT][8];
    float accb[COUT];
#pragma unroll
    for (int o = 0; o < COUT; ++o) { accb[o] = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[o][j] = 0.0f; }
    if (lane < lanes)
        for (size_t p = (size_t)blockIdx.x * lanes + lane; p < npix; p += (size_t)gridDim.x * lanes) {
            const size_t img = p / hw, rem = p - img * hw;
            float dp[COUT];
#pragma unroll
            for (int o = 0; o < COUT; ++o) {
                float d = dout[(img * COUT + o) * hw + rem];
                if (out_sig) { const float s = out_sig[(img * COUT + o) * hw + rem]; d *= s * (1.0f - s); }
                dp[o] = d;
                if (g == 0) accb[o] += d;
            }
            float f[8], ds[8];
            ld8(src + p * ld_s + 8 * g, lo_s, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) ds[j] = 0.0f;
#pragma unroll
            for (int o = 0; o < COUT; ++o)
#pragma unroll
                for (int j = 0; j < 8; ++j) { acc[o][j] = fmaf(dp[o], f[j], acc[o][j]); ds[j] = fmaf(dp[o], s_w[o * c_in + 8 * g + j], ds[j]); }
            st8f(dsrc + p * c_in + 8 * g, ds);
        }
    __syncthreads();
    float* s_red = s_w;                             // reuse: [lanes][c_in] per output channel, one o at a time
    for (int o = 0; o < COUT; ++o) {
        if (lane < lanes)
#pragma unroll
            for (int j = 0; j < 8; ++j) s_red[(size_t)lane * c_in + 8 * g + j] = acc[o][j];
        __syncthreads();
        for (int c = threadIdx.x; c < c_in; c += blockDim.x) {
            double t = 0.0;
            for (int l = 0; l < lanes; ++l) t += (double)s_red[(size_t)l * c_in + c];
            atomicAdd(dw_acc + (size_t)o * c_in + c, t);
        }
        __syncthreads();
        if (lane < lanes && g == 0) s_red[lane] = accb[o];
        __syncthreads();
        if (threadIdx.x == 0) { double t = 0.0; for (int l = 0; l < lanes; ++l) t += (double)s_red[l]; atomicAdd(dw_acc + (size_t)COUT * c_in + o, t); }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// stem weight gradient: dW[tap*cin+ci][co] += sum_p x[p+tap][ci] * dz[p][co]   (x fp32 NCHW counts, sparse)
__global__ void __launch_bounds__(256) stem_wgrad_kernel(const float* __restrict__ x, int n, int cin, int h, int w, const float* __restrict__ dz /* NHWC 64 */,
                                                         double* dw_acc /* [9*cin][64] */) {
    // one warp per (pixel stripe); lanes = 64 output channels as 2 per lane
    extern __shared__ float s_acc[];               // [9*cin][64]
    const int nk = 9 * cin;
    for (int i = threadIdx.x; i < nk * 64; i += blockDim.x) s_acc[i] = 0.0f;
    __syncthreads();
    const size_t hw = (size_t)h * w, total = (size_t)n * hw;
    const int lane = threadIdx.x & 31;
    const size_t warp_global = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, warp_stride = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t p = warp_global; p < total; p += warp_stride) {
        const int img = (int)(p / hw); const int rem = (int)(p - (size_t)img * hw);
        const int y = rem / w, xx = rem - y * w;
        const float2 d = *reinterpret_cast<const float2*>(dz + p * 64 + 2 * lane);
        if (__all_sync(0xffffffffu, d.x == 0.0f && d.y == 0.0f)) continue;
        // the 9*cin input values of this pixel's 3x3 neighbourhood, one per lane; the count images are sparse, so only
        // the non-zero ones (ballot) are multiplied into the shared accumulator
        for (int base = 0; base < nk; base += 32) {
            const int k = base + lane;
            float v = 0.0f;
            if (k < nk) {
                const int tap = k / cin, ci = k - tap * cin;
                const int yy = y + tap / 3 - 1, xc = xx + tap % 3 - 1;
                if (yy >= 0 && yy < h && xc >= 0 && xc < w) v = __ldg(x + ((size_t)img * cin + ci) * hw + (size_t)yy * w + xc);
            }
            unsigned mask = __ballot_sync(0xffffffffu, v != 0.0f);
            while (mask) {
                const int b = __ffs(mask) - 1;
                mask &= mask - 1;
                const float vv = __shfl_sync(0xffffffffu, v, b);
                atomicAdd(&s_acc[(base + b) * 64 + 2 * lane], vv * d.x);
                atomicAdd(&s_acc[(base + b) * 64 + 2 * lane + 1], vv * d.y);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nk * 64; i += blockDim.x) if (s_acc[i] != 0.0f) atomicAdd(dw_acc + i, (double)s_acc[i]);
}

__global__ void axpy_kernel(float* __restrict__ y, const float* __restrict__ x, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] += x[i];
}
__global__ void add_strided_kernel(float* __restrict__ y, const float* __restrict__ x, int ld_x, size_t npix, int C) {   // y[p][c] += x[p*ld_x + c]
    const size_t total = npix * (size_t)C;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t p = i / C; const int c = (int)(i - p * C);
        y[i] += x[p * ld_x + c];
    }
}
__global__ void d2f_kernel(const double* __restrict__ src, float* __restrict__ dst, size_t n, float mul, int accumulate) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = (accumulate ? dst[i] : 0.0f) + (float)(src[i] * (double)mul);
}

static int tk_grid(size_t items, int threads) {
    size_t g = (items + threads - 1) / threads;
    if (g < 1) g = 1;
    if (g > 148 * 8) g = 148 * 8;
    return (int)g;
}
static int red_grid(size_t npix, int C) {
    const int lanes = 256 / (C / 8);
    size_t g = (npix + (size_t)lanes * 32 - 1) / ((size_t)lanes * 32);
    if (g < 1) g = 1;
    if (g > 148 * 4) g = 148 * 4;
    return (int)g;
}
static int group_size(int c) { int gs = 1; while (gs * 2 <= 32 && gs * 2 <= c / 8) gs *= 2; return gs; }

}  // namespace nbp

using namespace nbp;
#define ST ((cudaStream_t)stream)
#define H16(p) ((const __half*)(p))

static int chk_c(const char* who, int C) {
    if (C <= 0 || C % 8 || C > 2048) return invalid("%s: channel count must be a multiple of 8 in [8, 2048] (got %d)", who, C);
    return NBP_OK;
}

extern "C" int nbp_bn_train_stats(const void* z, int ld, int lo, int64_t npix, int C, const float* gamma, const float* beta,
                                  float* running_mean, float* running_var, float momentum, float eps,
                                  float* mean, float* invstd, float* scale, float* shift, double* workspace /* [2C], zeroed here */, void* stream) {
    if (!z || !gamma || !beta || !mean || !invstd || !scale || !shift || !workspace) return invalid("nbp_bn_train_stats: null pointer argument");
    int rc = chk_c("nbp_bn_train_stats", C);
    if (rc) return rc;
    if (npix <= 0) return invalid("nbp_bn_train_stats: npix must be positive");
    rc = check_cuda(cudaMemsetAsync(workspace, 0, sizeof(double) * 2 * C, ST), "memset");
    if (rc) return rc;
    const size_t smem = sizeof(float) * 2 * (size_t)(256 / (C / 8)) * C;
    bn_moments_kernel<<<red_grid(npix, C), 256, smem, ST>>>(H16(z), ld, lo, (size_t)npix, C, workspace);
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, ST>>>(workspace, (size_t)npix, C, gamma, beta, running_mean, running_var, momentum, eps,
                                                        mean, invstd, scale, shift, H16(z), ld, lo);
    count_launch(3);
    return check_cuda(cudaGetLastError(), "nbp_bn_train_stats launch");
}

extern "C" int nbp_affine_act(const void* z, int ld_z, int lo_z, int64_t npix, int C, const float* scale, const float* shift, int relu,
                              void* y, int ld_y, int lo_y, void* stream) {
    if (!z || !scale || !shift || !y) return invalid("nbp_affine_act: null pointer argument");
    int rc = chk_c("nbp_affine_act", C);
    if (rc) return rc;
    affine_act_kernel<<<tk_grid((size_t)npix * (C / 8), 256), 256, 0, ST>>>(H16(z), ld_z, lo_z, (size_t)npix, C, scale, shift, relu, (__half*)y, ld_y, lo_y);
    count_launch();
    return check_cuda(cudaGetLastError(), "nbp_affine_act launch");
}

extern "C" int nbp_att_pre(const void* zg, const void* zx, int ld, int lo, int64_t npix, int C, const float* sg, const float* tg,
                           const float* sx, const float* tx, void* a, int ld_a, int lo_a, void* stream) {
    if (!zg || !zx || !sg || !tg || !sx || !tx || !a) return invalid("nbp_att_pre: null pointer argument");
    int rc = chk_c("nbp_att_pre", C);
    if (rc) return rc;
    att_pre_kernel<<<tk_grid((size_t)npix * (C / 8), 256), 256, 0, ST>>>(H16(zg), H16(zx), ld, lo, (size_t)npix, C, sg, tg, sx, tx, (__half*)a, ld_a, lo_a);
    count_launch();
    return check_cuda(cudaGetLastError(), "nbp_att_pre launch");
}

extern "C" int nbp_psi_train(const void* a, int ld_a, int lo_a, int64_t npix, int C, const float* w_psi, const float* b_psi,
                             const float* gamma1, const float* beta1, float* running_mean1, float* running_var1, float momentum, float eps,
                             float* zpsi, float* stat4 /* mean, invstd, scale, shift */, double* workspace /* [2] */, void* stream) {
    if (!a || !w_psi || !b_psi || !gamma1 || !beta1 || !zpsi || !stat4 || !workspace) return invalid("nbp_psi_train: null pointer argument");
    int rc = chk_c("nbp_psi_train", C);
    if (rc) return rc;
    const int gs = group_size(C);
    const size_t warps = ((size_t)npix + (32 / gs) - 1) / (32 / gs);
    psi_dot_kernel<<<tk_grid(warps * 32, 256), 256, 0, ST>>>(H16(a), ld_a, lo_a, (size_t)npix, C, w_psi, b_psi, zpsi, gs);
    rc = check_cuda(cudaMemsetAsync(workspace, 0, sizeof(double) * 2, ST), "memset");
    if (rc) return rc;
    const int g = tk_grid((size_t)npix, 256 * 8);
    stats1_kernel<<<g, 256, 0, ST>>>(zpsi, (size_t)npix, workspace, 0);
    stats1_kernel<<<g, 256, 0, ST>>>(zpsi, (size_t)npix, workspace, 1);
    bn_finalize_kernel<<<1, 32, 0, ST>>>(workspace, (size_t)npix, 1, gamma1, beta1, running_mean1, running_var1, momentum, eps,
                                         stat4, stat4 + 1, stat4 + 2, stat4 + 3, nullptr, 0, 0);
    count_launch(5);
    return check_cuda(cudaGetLastError(), "nbp_psi_train launch");
}

extern "C" int nbp_att_apply(const float* zpsi, const float* psi_scale, const float* psi_shift, const void* x, int ld_x, int lo_x, int64_t npix, int C,
                             void* dst, int ld_d, int c_off, int lo_d, float* psi_out, void* stream) {
    if (!zpsi || !psi_scale || !psi_shift || !x || !dst) return invalid("nbp_att_apply: null pointer argument");
    int rc = chk_c("nbp_att_apply", C);
    if (rc) return rc;
    att_apply_kernel<<<tk_grid((size_t)npix * (C / 8), 256), 256, 0, ST>>>(zpsi, psi_scale, psi_shift, H16(x), ld_x, lo_x, (size_t)npix, C,
                                                                            (__half*)dst, ld_d, c_off, lo_d, psi_out);
    count_launch();
    return check_cuda(cudaGetLastError(), "nbp_att_apply launch");
}

extern "C" int nbp_bn_bwd(const float* dy, int ld_dy, const void* z, int ld_z, int lo_z, int64_t npix, int C, const float* scale, const float* shift,
                          const float* mean, const float* invstd, const float* gamma, int relu, float* dz, int ld_dz, float* amax /* zeroed here */,
                          float* dgamma, float* dbeta, double* workspace /* [2C] */, void* stream) {
    return nbp_bn_bwd_split(dy, ld_dy, z, ld_z, lo_z, npix, C, scale, shift, mean, invstd, gamma, relu, dz, ld_dz, amax, dgamma, dbeta, workspace,
                            nullptr, 0, 0, nullptr, nullptr, 0, stream);
}

extern "C" int nbp_bn_bwd_split(const float* dy, int ld_dy, const void* z, int ld_z, int lo_z, int64_t npix, int C, const float* scale, const float* shift,
                                const float* mean, const float* invstd, const float* gamma, int relu, float* dz, int ld_dz, float* amax,
                                float* dgamma, float* dbeta, double* workspace, void* dz_split, int ld_s, int lo_s, float* amax2 /* [2] scratch */,
                                float* inv_scale_vec, int n_vec, void* stream) {
    if (!dy || !z || !scale || !shift || !mean || !invstd || !gamma || (!dz && !dz_split) || !workspace) return invalid("nbp_bn_bwd: null pointer argument");
    if (dz_split && (!amax2 || lo_s < C || ld_s < lo_s + C || ld_s % 8 || lo_s % 8)) return invalid("nbp_bn_bwd_split: bad split destination (ld=%d lo=%d C=%d)", ld_s, lo_s, C);
    int rc = chk_c("nbp_bn_bwd", C);
    if (rc) return rc;
    if (dz_split) { rc = check_cuda(cudaMemsetAsync(amax2, 0, 2 * sizeof(float), ST), "memset"); if (rc) return rc; }
    rc = check_cuda(cudaMemsetAsync(workspace, 0, sizeof(double) * 2 * C, ST), "memset");
    if (rc) return rc;
    if (amax) { rc = check_cuda(cudaMemsetAsync(amax, 0, sizeof(float), ST), "memset"); if (rc) return rc; }
    const size_t smem = sizeof(float) * ((size_t)(256 / (C / 8)) * 2 * C + 4 * (size_t)C);
    static bool attr_d[NBP_MAX_DEVICES] = {};
    bool& attr = attr_d[device_slot()];
    if (!attr) {
        cudaFuncSetAttribute(bn_bwd_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        cudaFuncSetAttribute(bn_bwd_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 7 * 2048 * 4);
        attr = true;
    }
    bn_bwd_reduce_kernel<<<red_grid(npix, C), 256, smem, ST>>>(dy, ld_dy, H16(z), ld_z, lo_z, (size_t)npix, C, scale, shift, mean, invstd, relu, workspace,
                                                               dz_split ? amax2 : nullptr);
    bn_bwd_apply_kernel<<<tk_grid((size_t)npix * (C / 8), 256), 256, sizeof(float) * 7 * (size_t)C, ST>>>(dy, ld_dy, H16(z), ld_z, lo_z, (size_t)npix, C, scale, shift, mean, invstd,
                                                                               gamma, relu, workspace, dz, ld_dz, amax, dgamma, dbeta,
                                                                               (__half*)dz_split, ld_s, lo_s, amax2, inv_scale_vec, n_vec);
    count_launch(4);
    return check_cuda(cudaGetLastError(), "nbp_bn_bwd launch");
}

extern "C" int nbp_to_split_nhwc(const float* src, int ld_s, int64_t npix, int C, const float* amax, void* dst, int ld_d, int lo_d,
                                 float* inv_scale_vec, int n_vec, void* stream) {
    if (!src || !dst) return invalid("nbp_to_split_nhwc: null pointer argument");
    int rc = chk_c("nbp_to_split_nhwc", C);
    if (rc) return rc;
    to_split_nhwc_kernel<<<tk_grid((size_t)npix * (C / 8), 256), 256, 0, ST>>>(src, ld_s, (size_t)npix, C, amax, (__half*)dst, ld_d, lo_d, inv_scale_vec, n_vec);
    count_launch();
    return check_cuda(cudaGetLastError(), "nbp_to_split_nhwc launch");
}

extern "C" int nbp_to_split_cnhw(const float* src_f32, const void* src_split, int ld_s, int lo_s, int64_t npix, int w, int w_pad, int dx, int C, const float* amax,
                                 void* dst, int64_t row_stride, int64_t plane_stride, float* inv_scale_out, void* stream) {
    if ((!src_f32 && !src_split) || !dst) return invalid("nbp_to_split_cnhw: null pointer argument");
    if (npix <= 0 || C <= 0 || w <= 0 || w_pad < w || npix % w || row_stride < npix / w * w_pad) return invalid("nbp_to_split_cnhw: bad sizes");
    dim3 g((unsigned)((npix + 31) / 32), (unsigned)((C + 31) / 32));
    to_split_cnhw_kernel<<<g, 256, 0, ST>>>(src_f32, H16(src_split), ld_s, lo_s, (size_t)npix, w, w_pad, dx, C, amax, (__half*)dst, (size_t)row_stride, (size_t)plane_stride, inv_scale_out);
    count_launch();
    return check_cuda(cudaGetLastError(), "nbp_to_split_cnhw launch");
}

extern "C" int nbp_maxpool2x2_bwd(const float* dy, const void* x, int ld_x, int lo_x, int n, int h, int w, int C, float* dx, int accumulate, void* stream) {
    if (!dy || !x || !dx) return invalid("nbp_maxpool2x2_bwd: null pointer argument");
    int rc = chk_c("nbp_maxpool2x2_bwd", C);
    if (rc) return rc;
    maxpool_bwd_kernel<<<tk_grid((size_t)n * (h / 2) * (w / 2) * (C / 8), 256), 256, 0, ST>>>(dy, H16(x), ld_x, lo_x, n, h, w, C, dx, accumulate);
    count_launch();
    return check_cuda(cudaGetLastError(), "nbp_maxpool2x2_bwd launch");
}

extern "C" int nbp_upsample2x_bwd(const float* dy, int n, int h, int w, int C, float* dx, int accumulate, void* stream) {
    if (!dy || !dx) return invalid("nbp_upsample2x_bwd: null pointer argument");
    int rc = chk_c("nbp_upsample2x_bwd", C);
    if (rc) return rc;
    upsample_bwd_kernel<<<tk_grid((size_t)n * h * w * (C / 8), 256), 256, 0, ST>>>(dy, n, h, w, C, dx, accumulate);
    count_launch();
    return check_cuda(cudaGetLastError(), "nbp_upsample2x_bwd launch");
}

extern "C" int nbp_att_bwd(const float* dout, int ld_do, const void* x, int ld_x, int lo_x, const float* psi, const float* zpsi, int64_t npix, int C_l,
                           const float* stat4, const float* gamma1, const void* a, int ld_a, int lo_a, int C_int, const float* w_psi,
                           float* dx_skip, int accumulate, float* dpre /* [npix][C_int] */, float* dt /* [npix] scratch */,
                           float* dw_psi /* [C_int] += */, float* dgamma1, float* dbeta1, double* workspace /* [2 + C_int] */, void* stream) {
    if (!dout || !x || !psi || !zpsi || !stat4 || !gamma1 || !a || !w_psi || !dx_skip || !dpre || !dt || !dw_psi || !dgamma1 || !dbeta1 || !workspace)
        return invalid("nbp_att_bwd: null pointer argument");
    int rc = chk_c("nbp_att_bwd", C_l);
    if (rc) return rc;
    rc = chk_c("nbp_att_bwd", C_int);
    if (rc) return rc;
    rc = check_cuda(cudaMemsetAsync(workspace, 0, sizeof(double) * (2 + C_int), ST), "memset");
    if (rc) return rc;
    const int gs = group_size(C_l);
    const size_t warps = ((size_t)npix + (32 / gs) - 1) / (32 / gs);
    att_apply_bwd_kernel<<<tk_grid(warps * 32, 256), 256, 0, ST>>>(dout, ld_do, H16(x), ld_x, lo_x, psi, (size_t)npix, C_l, dx_skip, accumulate, dt, gs);
    bn1_bwd_reduce_kernel<<<tk_grid((size_t)npix, 256 * 8), 256, 0, ST>>>(dt, zpsi, (size_t)npix, stat4, stat4 + 1, workspace);
    const size_t smem = sizeof(float) * (size_t)(256 / (C_int / 8)) * C_int;
    psi_bwd_kernel<<<red_grid(npix, C_int), 256, smem, ST>>>(dt, zpsi, (size_t)npix, stat4, stat4 + 1, gamma1, workspace, H16(a), ld_a, lo_a, C_int,
                                                             w_psi, dpre, workspace + 2, dgamma1, dbeta1);
    d2f_kernel<<<1, 256, 0, ST>>>(workspace + 2, dw_psi, (size_t)C_int, 1.0f, 1);
    count_launch(5);
    return check_cuda(cudaGetLastError(), "nbp_att_bwd launch");
}

extern "C" int nbp_head_bwd(const float* dout, const float* out_sigmoid, const void* src, int c_in, int ld_s, int lo_s, const float* weight, int c_out,
                            int n, int64_t hw, float* dsrc, float* dweight /* += [c_out][c_in] */, float* dbias /* += */, double* workspace, void* stream) {
    if (!dout || !src || !weight || !dsrc || !dweight || !dbias || !workspace) return invalid("nbp_head_bwd: null pointer argument");
    int rc = chk_c("nbp_head_bwd", c_in);
    if (rc) return rc;
    const size_t nacc = (size_t)c_out * c_in + c_out;
    rc = check_cuda(cudaMemsetAsync(workspace, 0, sizeof(double) * nacc, ST), "memset");
    if (rc) return rc;
    const int lanes = 256 / (c_in / 8);
    size_t smem = sizeof(float) * (size_t)c_in * (c_out > lanes ? c_out : lanes);
    const int g = red_grid((size_t)n * hw, c_in);
    if (c_out == 8) head_bwd_kernel<8><<<g, 256, smem, ST>>>(dout, out_sigmoid, H16(src), c_in, ld_s, lo_s, weight, n, (size_t)hw, dsrc, workspace);
    else if (c_out == 1) head_bwd_kernel<1><<<g, 256, smem, ST>>>(dout, out_sigmoid, H16(src), c_in, ld_s, lo_s, weight, n, (size_t)hw, dsrc, workspace);
    else return invalid("nbp_head_bwd: c_out must be 1 or 8");
    d2f_kernel<<<tk_grid((size_t)c_out * c_in, 256), 256, 0, ST>>>(workspace, dweight, (size_t)c_out * c_in, 1.0f, 1);
    d2f_kernel<<<1, 32, 0, ST>>>(workspace + (size_t)c_out * c_in, dbias, (size_t)c_out, 1.0f, 1);
    count_launch(4);
    return check_cuda(cudaGetLastError(), "nbp_head_bwd launch");
}

extern "C" int nbp_stem_wgrad(const float* x, int n, int c_in, int h, int w, const float* dz, float* dweight /* += [9*c_in][64] */, double* workspace, void* stream) {
    if (!x || !dz || !dweight || !workspace) return invalid("nbp_stem_wgrad: null pointer argument");
    if (c_in <= 0 || c_in > 16) return invalid("nbp_stem_wgrad: c_in must be in [1, 16]");
    const size_t nacc = (size_t)9 * c_in * 64;
    int rc = check_cuda(cudaMemsetAsync(workspace, 0, sizeof(double) * nacc, ST), "memset");
    if (rc) return rc;
    static bool attr_d[NBP_MAX_DEVICES] = {};
    bool& attr = attr_d[device_slot()];
    if (!attr) { cudaFuncSetAttribute(stem_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 9 * 16 * 64 * 4); attr = true; }
    stem_wgrad_kernel<<<148 * 2, 256, sizeof(float) * nacc, ST>>>(x, n, c_in, h, w, dz, workspace);
    d2f_kernel<<<tk_grid(nacc, 256), 256, 0, ST>>>(workspace, dweight, nacc, 1.0f, 1);
    count_launch(3);
    return check_cuda(cudaGetLastError(), "nbp_stem_wgrad launch");
}

extern "C" int nbp_add_f32(float* y, const float* x, int ld_x, int64_t npix, int C, void* stream) {
    if (!y || !x) return invalid("nbp_add_f32: null pointer argument");
    if (ld_x == C) axpy_kernel<<<tk_grid((size_t)npix * C, 256 * 4), 256, 0, ST>>>(y, x, (size_t)npix * C);
    else add_strided_kernel<<<tk_grid((size_t)npix * C, 256 * 4), 256, 0, ST>>>(y, x, ld_x, (size_t)npix, C);
    count_launch();
    return check_cuda(cudaGetLastError(), "nbp_add_f32 launch");
}
