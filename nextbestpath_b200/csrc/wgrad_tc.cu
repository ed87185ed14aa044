// Weight-gradient GEMM of the NBP convolutions on tcgen05/TMEM/TMA:
//
//   dW[co][tap][ci] = sum over pixels p of  dz[p][co] * x[p + tap][ci]           (tap = (dy,dx), zero outside the image)
//
// The reduction dimension is the PIXEL index.  Both operands stay in their NHWC fp16x2 layout (no transposed copies):
// a K slice of 64 pixels x 64 channels is the same 4-D TMA box (64 ch, tw, th, tn) the forward kernel loads, with the
// tap shift applied to the pixel coordinates (TMA zero fill = conv padding).  In shared memory such a box is 64 rows
// (pixels = K) of 128 bytes (64 channels = M or N), SWIZZLE_128B: the canonical **MN-major** UMMA operand
// (instruction-descriptor a_major = b_major = 1; smem descriptor LBO = 8 KB between 64-channel atoms, SBO = 1 KB
// between 8-pixel groups).
//
// fp16x2 operands (hi + lo/2048) as in conv_tc.cu: UMMA#1  R_hi x [Cc_hi ; Cc_lo]  (N = 2*BLOCK_N),
// UMMA#2  R_lo x Cc_hi into the lo columns; epilogue acc_hi + acc_lo/2048, times the inverse gradient scale,
// fp32 atomicAdd into dW (split-K across CTAs).
// R ("row operand", M = 128 channels) and Cc ("column operand", N = BLOCK_N channels) are dz and x or x and dz: the host
// picks the orientation that fills M; `shift_rows` says which one carries the tap shift (always x).
#include <cuda.h>

#include "nbp_common.cuh"
#include "tc_ptx.cuh"

namespace nbp {

using namespace tc;

static constexpr int WG_M = 128;
static constexpr int WG_K = 64;
static constexpr int WG_THREADS = 256;
static constexpr int WG_ATOM_BYTES = 64 * WG_K * 2;         // one TMA box: 64 pixels x 64 channels
static constexpr int WG_A_BYTES = WG_M * WG_K * 2;          // one plane of the row operand = 2 atoms
static constexpr int WG_SMEM_LIMIT = 232448;
static constexpr int WG_AUX = 1024;

struct WgradParams {
    int n, h, w;                       // image batch
    int tw, th, tn, tiles_x, tiles_y, tiles_n, k_tiles;
    int m_tiles, n_tiles, taps, splits;
    int lo_r, lo_c;                    // channel offset of the lo plane in each NHWC tensor
    int shift_rows;                    // 1: the row operand is x (gets the tap shift); 0: the column operand is x
    int m_valid, n_valid;              // channels actually present (rows beyond are junk / zero and are not stored)
    long long stride_m, stride_n, stride_tap;
    const float* inv_scale;            // device scalar (1 / gradient scale)
    float* out;
};

template <int BLOCK_N>
struct WgCfg {
    static constexpr int B_BYTES = 2 * BLOCK_N * WG_K * 2;                 // [Cc_hi atoms ; Cc_lo atoms], 8 KB each
    static constexpr int STAGE_BYTES = 2 * WG_A_BYTES + B_BYTES;
    static constexpr int STAGES_RAW = (WG_SMEM_LIMIT - WG_AUX - 1024) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
    static constexpr int ACC_COLS = 2 * BLOCK_N;
    static constexpr int TMEM_COLS = (2 * ACC_COLS <= 64) ? 64 : (2 * ACC_COLS <= 128) ? 128 : (2 * ACC_COLS <= 256) ? 256 : 512;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + WG_AUX + 1024;
    static_assert(STAGES >= 2, "pipeline needs two stages");
};

template <int BLOCK_N>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_gemm_f16x2(const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmC, const WgradParams p) {
    using Cfg = WgCfg<BLOCK_N>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * 2 * WG_A_BYTES;
    uint8_t* aux = smem + STAGES * Cfg::STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) { prefetch_tmap(&tmR); prefetch_tmap(&tmC); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 4); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // work item = (output tile (m_tile, n_tile, tap), K split)
    const int out_tiles = p.m_tiles * p.n_tiles * p.taps;
    const int num_items = out_tiles * p.splits;
    const int k_per = (p.k_tiles + p.splits - 1) / p.splits;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                const int split = item / out_tiles; int ot = item - split * out_tiles;
                const int tap = ot % p.taps; ot /= p.taps;
                const int n_tile = ot % p.n_tiles, m_tile = ot / p.n_tiles;
                const int dy = p.taps == 9 ? tap / 3 - 1 : 0, dx = p.taps == 9 ? tap % 3 - 1 : 0;
                const int k0 = split * k_per, k1 = min(p.k_tiles, k0 + k_per);
                for (int kt = k0; kt < k1; ++kt) {
                    int t = kt;
                    const int tx = t % p.tiles_x; t /= p.tiles_x;
                    const int ty = t % p.tiles_y; const int tb = t / p.tiles_y;
                    const int x0 = tx * p.tw, y0 = ty * p.th, n0 = tb * p.tn;
                    const int rx = p.shift_rows ? dx : 0, ry = p.shift_rows ? dy : 0;
                    const int cx = p.shift_rows ? 0 : dx, cy = p.shift_rows ? 0 : dy;
                    mbar_wait(&empty_bar[stage], phase ^ 1, 101);
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                    uint8_t* sa = smem_a + stage * 2 * WG_A_BYTES;
                    uint8_t* sb = smem_b + stage * Cfg::B_BYTES;
#pragma unroll
                    for (int a = 0; a < 2; ++a) {                  // row operand: hi atoms 0,1 then lo atoms 0,1
                        tma_load_4d(sa + a * WG_ATOM_BYTES, &tmR, &full_bar[stage], m_tile * WG_M + a * 64, x0 + rx, y0 + ry, n0);
                        tma_load_4d(sa + WG_A_BYTES + a * WG_ATOM_BYTES, &tmR, &full_bar[stage], p.lo_r + m_tile * WG_M + a * 64, x0 + rx, y0 + ry, n0);
                    }
#pragma unroll
                    for (int a = 0; a < BLOCK_N / 64; ++a) {       // column operand: hi atoms then lo atoms
                        tma_load_4d(sb + a * WG_ATOM_BYTES, &tmC, &full_bar[stage], n_tile * BLOCK_N + a * 64, x0 + cx, y0 + cy, n0);
                        tma_load_4d(sb + (BLOCK_N / 64 + a) * WG_ATOM_BYTES, &tmC, &full_bar[stage], p.lo_c + n_tile * BLOCK_N + a * 64, x0 + cx, y0 + cy, n0);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc_main = umma_idesc_f16_mn(WG_M, 2 * BLOCK_N);
            constexpr uint32_t idesc_lo = umma_idesc_f16_mn(WG_M, BLOCK_N);
            int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                const int split = item / out_tiles;
                const int k0 = split * k_per, k1 = min(p.k_tiles, k0 + k_per);
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1, 102);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * Cfg::ACC_COLS);
                for (int kt = k0; kt < k1; ++kt) {
                    mbar_wait(&full_bar[stage], phase, 103);
                    tc_fence_after();
                    const uint32_t a_base = smem_u32(smem_a + stage * 2 * WG_A_BYTES);
                    const uint32_t b_base = smem_u32(smem_b + stage * Cfg::B_BYTES);
#pragma unroll
                    for (int k = 0; k < WG_K / 16; ++k) {
                        // one UMMA consumes 16 pixels = two 8-row groups = 2048 bytes of every 64-channel atom
                        const uint64_t ahi = umma_desc_mnmajor_sw128(a_base + k * 2048, WG_ATOM_BYTES, 1024);
                        const uint64_t alo = umma_desc_mnmajor_sw128(a_base + WG_A_BYTES + k * 2048, WG_ATOM_BYTES, 1024);
                        const uint64_t bd = umma_desc_mnmajor_sw128(b_base + k * 2048, WG_ATOM_BYTES, 1024);
                        umma_f16(d_tmem, ahi, bd, idesc_main, (kt > k0 || k > 0) ? 1u : 0u);
                        umma_f16(d_tmem + (uint32_t)BLOCK_N, alo, bd, idesc_lo, 1u);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull_bar[acc]);          // (an empty K range still commits: nothing was issued, arrives at once)
                acc ^= 1; if (acc == 0) acc_phase ^= 1;
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int m = q * 32 + lane;
        int acc = 0; uint32_t acc_phase = 0;
        const float inv = p.inv_scale ? p.inv_scale[0] : 1.0f;
        for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
            const int split = item / out_tiles; int ot = item - split * out_tiles;
            const int tap = ot % p.taps; ot /= p.taps;
            const int n_tile = ot % p.n_tiles, m_tile = ot / p.n_tiles;
            const int k0 = split * k_per, k1 = min(p.k_tiles, k0 + k_per);
            mbar_wait(&tfull_bar[acc], acc_phase, 104);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * Cfg::ACC_COLS);
            const int mg = m_tile * WG_M + m;
            if (k1 > k0) {
#pragma unroll 1
                for (int c = 0; c < BLOCK_N / 32; ++c) {
                    uint32_t v[32], vl[32];
                    tmem_ld_32x32(t_row + (uint32_t)(c * 32), v);
                    tmem_ld_32x32(t_row + (uint32_t)(BLOCK_N + c * 32), vl);
                    tmem_ld_wait();
                    if (mg < p.m_valid) {
                        float* o = p.out + (long long)mg * p.stride_m + (long long)tap * p.stride_tap;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int ng = n_tile * BLOCK_N + c * 32 + j;
                            if (ng < p.n_valid) atomicAdd(o + (long long)ng * p.stride_n, fmaf(__uint_as_float(vl[j]), 1.0f / 2048.0f, __uint_as_float(v[j])) * inv);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            acc ^= 1; if (acc == 0) acc_phase ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

typedef CUresult (*PFN_encodeTiledW)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiledW wg_encode() {
    static PFN_encodeTiledW fn = nullptr;
    if (!fn) {
        void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
        fn = (PFN_encodeTiledW)ptr;
    }
    return fn;
}

// NHWC fp16x2 tensor: `span` channels per pixel are addressable (hi plane + lo plane), pixel stride ld
static int make_nhwc_map(CUtensorMap* m, const void* ptr, int span, int ld, int n, int h, int w, int tw, int th, int tn) {
    PFN_encodeTiledW enc = wg_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return NBP_ERR_UNSUPPORTED; }
    cuuint64_t dims[4] = {(cuuint64_t)span, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)w * ld * 2, (cuuint64_t)h * w * ld * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tn};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(nhwc span=%d ld=%d n=%d h=%d w=%d box=%d,%d,%d) failed: %d", span, ld, n, h, w, tw, th, tn, (int)r); return NBP_ERR_INVALID; }
    return NBP_OK;
}

template <int BLOCK_N>
static int launch_wgrad(const CUtensorMap& r, const CUtensorMap& c, const WgradParams& p, int sms, cudaStream_t st) {
    using Cfg = WgCfg<BLOCK_N>;
    static bool attr_d[NBP_MAX_DEVICES] = {};
    bool& attr = attr_d[device_slot()];
    if (!attr) {
        int rc = check_cuda(cudaFuncSetAttribute(wgrad_gemm_f16x2<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES), "cudaFuncSetAttribute(wgrad)");
        if (rc) return rc;
        attr = true;
    }
    const int items = p.m_tiles * p.n_tiles * p.taps * p.splits;
    wgrad_gemm_f16x2<BLOCK_N><<<items < sms ? items : sms, WG_THREADS, Cfg::SMEM_BYTES, st>>>(r, c, p);
    count_launch();
    return check_cuda(cudaGetLastError(), "wgrad_gemm_f16x2 launch");
}

}  // namespace nbp

using namespace nbp;

static int wg_pow2_floor(int v) { int p = 1; while (p * 2 <= v) p *= 2; return p; }
static int wg_pow2_ceil(int v) { int p = 1; while (p < v) p *= 2; return p; }

extern "C" int nbp_conv_wgrad(const void* dz, int c_out, int ld_dz, int lo_dz, const void* x, int c_in, int ld_x, int lo_x,
                              int n, int h, int w, int taps, const float* inv_scale, float* dweight, int max_k_tiles, void* stream) {
    if (!dz || !x || !dweight) return invalid("nbp_conv_wgrad: null pointer argument");
    if (taps != 1 && taps != 9) return invalid("nbp_conv_wgrad: taps must be 1 or 9");
    if (c_out <= 0 || c_in <= 0 || c_out % 64 || c_in % 64) return invalid("nbp_conv_wgrad: channel counts must be positive multiples of 64 (pad with zeros): c_out=%d c_in=%d", c_out, c_in);
    if (lo_dz < c_out || lo_x < c_in || lo_dz % 8 || lo_x % 8 || ld_dz < lo_dz + c_out || ld_x < lo_x + c_in || ld_dz % 8 || ld_x % 8)
        return invalid("nbp_conv_wgrad: operands must be fp16x2 split tensors (bad ld/lo)");
    if (n <= 0 || h <= 0 || w <= 0) return invalid("nbp_conv_wgrad: bad image dims");
    if (((uintptr_t)dz | (uintptr_t)x) & 15) return invalid("nbp_conv_wgrad: pointers must be 16-byte aligned");

    WgradParams p{};
    p.n = n; p.h = h; p.w = w;
    p.tw = w >= 16 ? 16 : wg_pow2_floor(w);                 // 64-pixel K slice = (tw, th, tn) box of the image batch
    int th = 64 / p.tw; const int hc = wg_pow2_ceil(h);
    p.th = th < hc ? th : hc;
    p.tn = 64 / (p.tw * p.th);
    p.tiles_x = (w + p.tw - 1) / p.tw; p.tiles_y = (h + p.th - 1) / p.th; p.tiles_n = (n + p.tn - 1) / p.tn;
    p.k_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
    p.taps = taps;
    // orientation: the operand with more channels fills M = 128
    const bool x_rows = c_in > c_out;
    const int c_r = x_rows ? c_in : c_out, c_c = x_rows ? c_out : c_in;
    const int block_n = c_c >= 128 ? 128 : 64;
    p.shift_rows = x_rows ? 1 : 0;
    p.m_tiles = (c_r + WG_M - 1) / WG_M; p.n_tiles = (c_c + block_n - 1) / block_n;
    p.lo_r = x_rows ? lo_x : lo_dz; p.lo_c = x_rows ? lo_dz : lo_x;
    p.m_valid = c_r; p.n_valid = c_c;
    const long long ktc = (long long)taps * c_in;
    if (x_rows) { p.stride_m = 1; p.stride_n = ktc; } else { p.stride_m = ktc; p.stride_n = 1; }
    p.stride_tap = c_in;
    p.inv_scale = inv_scale; p.out = dweight;
    static int sms_d[NBP_MAX_DEVICES] = {};
    int& sms = sms_d[device_slot()];
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    const int out_tiles = p.m_tiles * p.n_tiles * taps;
    int splits = (2 * sms + out_tiles - 1) / out_tiles;
    // The tensor core's fp32 accumulator truncates (error grows linearly with the chain length); weight gradients are
    // sums with heavy cancellation, so the in-TMEM chain is bounded to max_k_tiles 64-pixel slices and the partial
    // results are combined by fp32 round-to-nearest atomics.
    if (max_k_tiles > 0 && (p.k_tiles + splits - 1) / splits > max_k_tiles) splits = (p.k_tiles + max_k_tiles - 1) / max_k_tiles;
    if (splits > p.k_tiles) splits = p.k_tiles;
    if (splits < 1) splits = 1;
    p.splits = splits;

    CUtensorMap tr, tc_;
    const void* r_ptr = x_rows ? x : dz; const void* c_ptr = x_rows ? dz : x;
    const int ld_r = x_rows ? ld_x : ld_dz, ld_c = x_rows ? ld_dz : ld_x;
    int rc = make_nhwc_map(&tr, r_ptr, p.lo_r + c_r, ld_r, n, h, w, p.tw, p.th, p.tn);
    if (rc) return rc;
    rc = make_nhwc_map(&tc_, c_ptr, p.lo_c + c_c, ld_c, n, h, w, p.tw, p.th, p.tn);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (block_n == 128) return launch_wgrad<128>(tr, tc_, p, sms, st);
    return launch_wgrad<64>(tr, tc_, p, sms, st);
}

// debugging aid: attach a zero-copy host word block (>= 4 ints) that a timed-out mbarrier wait of the wgrad kernel fills
extern "C" int nbp_debug_attach_wgrad(int* device_visible_host_ptr) {
    return check_cuda(cudaMemcpyToSymbol(nbp::tc::nbp_dbg_ptr, &device_visible_host_ptr, sizeof(int*)), "cudaMemcpyToSymbol(dbg)");
}
