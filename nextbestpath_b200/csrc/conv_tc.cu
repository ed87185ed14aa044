// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05 + TMEM + TMA) for the NBP
// attention-U-Net (/root/reference/next_best_path/networks/nbp_model.py:8-62,110-160).
//
//   out[n,y,x,co] = act( scale[co] * sum_{tap,ci} in[n, y+dy, x+dx, ci] * w[co, tap, ci] + shift[co] )
//
// Two numeric modes (template PRECISE), both fp32-accumulating in TMEM with an fp32 epilogue affine
// (folded conv bias + BatchNorm):
//   PRECISE = true  "fp16x2": every activation and weight is an unevaluated sum hi + lo/2048 of two fp16
//       numbers (22 significand bits).  Per K slice the tensor core computes
//           acc_hi  = A_hi * W_hi                       } one UMMA with N = 2*BLOCK_N: B rows [W_hi ; W_lo]
//           acc_lo  = A_hi * W_lo  (+)  A_lo * W_hi       } second UMMA (N = BLOCK_N) into the acc_lo columns
//       and the epilogue returns acc_hi + acc_lo/2048.  Measured on the BN-calibrated parity weights this
//       is 7e-6 relative to the fp32 reference (plain fp16 or tf32 storage: 7e-3, bf16: 5e-2), which is what
//       the "value maps within 1e-3 of fp32" bar needs ("numeric format").  3 MMA passes.
//   MODE 2 "fp16+e4m3": the same hi planes, but both correction products run on the 8-bit floating-point pipe (kind::f8f6f4,
//       twice the MAC rate).  The second plane of an activation holds, per 64-channel group, 64 bytes e4m3(x) followed by
//       64 bytes e4m3((x - hi) * 2048); the matching weight rows hold e4m3(W_lo * 2048 * s) followed by e4m3(W_hi * s) (s = a
//       per-layer power of two).  ONE K = 128 fp8 reduction per 64-channel slice then yields
//           acc_lo = A_hi8 * W_lo8  +  A_lo8 * W_hi8
//       and the epilogue returns acc_hi + acc_lo / (2048 s).  Operands keep ~15 bits instead of 22 and a slice costs
//       4 + 4 half-rate-equivalent UMMAs instead of 12: 2/3 of the tensor time.  Measured network error with the five most
//       sensitive encoder layers kept in fp16x2: 1.0-1.3e-4 (value map) / 2.4-4.8e-4 (obstacle map), bar 1e-3.
//   PRECISE = false "fp16": single fp16 plane, 1 MMA pass; 7e-3 relative on the same weights.
// Layout: activations NHWC fp16, channel counts multiples of 64; PRECISE tensors carry a second plane
// (lo * 2048) `lo_off` elements after the first.  Weights fp16 [c_out][taps][c_in] (K-major).
//
// GEMM view per CTA tile: M = 128 output pixels (a tw x th x tn box of the image batch), N = BLOCK_N output
// channels, K = taps * c_in walked in 64-channel slices.  No im2col buffer exists anywhere: the A operand of
// tap (dy,dx) is the SAME 4-D TMA box shifted by (dx,dy); TMA zero-fills the out-of-image halo, which is the
// conv's zero padding.  Channel concatenation (torch.cat((skip*psi, up), 1), nbp_model.py:128) is fused by
// walking two source tensors inside the K loop.
//
// Warp roles (320 threads, 1 CTA / SM, persistent over tiles; CTA pairs: see ConvCfg::PAIR):
//   warp 0 : TMA producer (one elected lane)      smem ring of STAGES x {A block(s), weight tile(s)} -- a plain stage is one tap x one
//            64-channel slice, a halo stage the (th+2)-row A block of one column offset + the weight tiles of its 2-3 taps
//   warp 1 : tcgen05.mma issuer (one elected lane; in a CTA pair the even CTA issues cta_group::2 MMAs for both)
//   warps 2-9 : epilogue, TMEM -> registers -> affine/ReLU -> fp16 planes -> global (two warps per TMEM lane quarter, alternate
//            32-column groups); double-buffered accumulators so the epilogue of tile i overlaps the main loop of tile i+1.
//            Optional epilogue work: fused 2x2 max-pool (pool_dst), the "dot" contraction with a vector instead of the store
//            (dot_w / dot_out: attention psi, Final2) and the attention gate on top of it (gate_src)
#include <cuda.h>

#include <cstdlib>
#include <vector>

#include "nbp_common.cuh"
#include "tc_ptx.cuh"

namespace nbp {

using namespace tc;

// two floats -> packed e4m3x2 (round to nearest even, saturating at +-448); low byte = first value
__device__ __forceinline__ uint32_t e4m3x2(float a, float b) {
    uint16_t r;
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %2, %1;" : "=h"(r) : "f"(a), "f"(b));
    return (uint32_t)r;
}

static constexpr int BLOCK_M = 128;
static constexpr int BLOCK_K = 64;                         // fp16 elements = one 128-byte swizzle row
static constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;     // one plane of a plain 128-pixel A block
static constexpr int HALO_ROWS = 160;                          // (th + 2) * tw pixels of a halo block (tw = 16, th = 8)
static constexpr int A_HALO_BYTES = HALO_ROWS * BLOCK_K * 2;   // one plane of a halo A block
#ifndef NBP_EPI_WARPS
#define NBP_EPI_WARPS 8
#endif
// warp 0 = TMA producer, warp 1 = MMA issuer, warps 2.. = epilogue.  An epilogue warp may only read the TMEM lane quarter (warp & 3), so
// warps 2,3,4,5 cover the 128 accumulator rows once; with 8 warps two warps share each quarter (and each scheduler) and take alternate
// 32-column groups -- one warp per scheduler cannot hide the latencies of the epilogue (TMEM loads, shuffles, stores that wait for the LSU)
static constexpr int EPI_WARPS = NBP_EPI_WARPS;
static constexpr int EPI_SPLIT = EPI_WARPS / 4;                 // warps sharing one lane quarter
static_assert(EPI_WARPS == 4 || EPI_WARPS == 8, "4 or 8 epilogue warps");
static constexpr int CONV_THREADS = 64 + 32 * EPI_WARPS;
static constexpr int SMEM_LIMIT = 232448;                  // 227 KB opt-in maximum per CTA
static constexpr int SMEM_AUX = 4096;                      // barriers + tmem ptr + scale/shift staging

// n / d and n % d for 0 <= n < 2^31 by one multiply-high (Granlund-Montgomery round-up multiplier)
struct FastDiv { uint32_t d, mul, shr; };
static FastDiv make_fastdiv(uint32_t d) {
    FastDiv f{d, 0u, 0u};
    if (d <= 1) return f;
    uint32_t lg = 0;
    while ((1ull << lg) < d) ++lg;
    f.mul = (uint32_t)(((1ull << (31 + lg)) + d - 1) / d);
    f.shr = lg - 1;
    return f;
}
__device__ __forceinline__ void fast_divmod(const FastDiv& f, int n, int& q, int& r) {
    q = f.d == 1 ? n : (int)(__umulhi((uint32_t)n, f.mul) >> f.shr);
    r = n - q * (int)f.d;
}

struct ConvKParams {
    FastDiv fd_n_tiles, fd_tiles_x, fd_tiles_y;   // tile index -> (n-tile, x, y, image group)
    int n, h, w;
    int tw, th, tn, tiles_x, tiles_y, tiles_n;
    int m_tiles, n_tiles;
    int taps, kc0, kc1;
    int lo0, lo1;                       // element offset of the lo plane inside the tensor maps (PRECISE)
    float lo_scale;                     // acc = acc_hi + acc_lo * lo_scale (1/2048 for fp16x2, 1/(2048 s) for fp16+e4m3)
    int dst_fmt, pool_fmt;              // second-plane format written by the epilogue: 1 = fp16 lo*2048, 2 = e4m3 pair (MODE 2 consumers)
    int cluster;                        // 1, or 2: CTA pairs (thread-block cluster) work on two m-tiles of the SAME n-tile in lock step and each
                                        // loads half of every weight tile, multicast into both CTAs' shared memory (half the L2->SM weight bytes)
    int interleave;                     // MODE 2: 1 (default) = alternate the fp16 and e4m3 UMMAs per 16-element K step, 0 = two runs per stage
    unsigned long long* sat_count;      // optional: += number of (pixel, 32-channel group) stores in which an e4m3 value saturated
    int out_f32;                        // 1: the destination is plain fp32 NHWC (gradients), no fp16 planes
    int kchunk;                         // K stages accumulated inside TMEM before the partial sum is folded into fp32
                                        // registers (round-to-nearest); the tensor core's own accumulator truncates
    int up2x;                           // 1: fused nearest-2x upsample + 3x3 conv as 4 parity-specific 2x2 convs (taps == 4)
    int b_rows_per_parity;              // rows of the weight matrix per parity block (up2x)
    // K loop geometry.  plain: one stage = one tap x one 64-channel slice (ngroups = taps, gtaps = 1).  halo (HALO kernels): one
    // stage = the (th+2)-row A block of one column offset dx + the B tiles of its gtaps taps (3x3: 3 groups x 3 taps, up2x: 2 x 2)
    int ngroups, gtaps;
    int a_tx;                           // bytes the A loads of one stage deliver
    int aoff_step;                      // halo: byte offset of one tile row inside the A block (tw * 128)
    const float* scale; const float* shift;
    int relu;
    __half* dst; int dst_ld, dst_c_off, dst_lo_off;
    __half* pool; int pool_ld, pool_lo_off;   // optional second output: the 2x2 max-pooled activation [n][h/2][w/2] (fused nn.MaxPool2d)
    // optional "dot" epilogue (n_tiles == 1, one accumulation chain): the activation tile is NOT stored; every pixel's c_out activations
    // are contracted with dot_w in fp32 and dot_out[pixel] = f(dot * dot_scale + dot_shift), f = sigmoid or identity.  Fuses the 1-channel
    // 1x1 convolution that follows (Attention_block.psi, Final2) so that its input never goes to memory.
    const float* dot_w; float* dot_out; float dot_scale, dot_shift; int dot_sigmoid;
    int dot_n;                          // 1 .. 8 vectors (rows of dot_w, c_out apart): dot_out is [n][dot_n][oh][ow] (a fused 1x1 conv to dot_n channels)
    const float* dot_bias;              // optional [dot_n], added after the scale
    float* dot_max;                     // optional [n][oh][ow]: max over the dot_n outputs (the heading read-out of Final1)
    // optional gate on top of the dot epilogue (Attention_block, nbp_model.py:62): dst[pixel][dst_c_off + c] = gate_src[pixel][c] * f(dot),
    // c < gate_c, written in dst_fmt like a normal output; gate_src is an NHWC tensor in the sources' format (gate_fmt)
    const __half* gate_src; int gate_c, gate_ld, gate_lo, gate_fmt;
};

// MODE 0 fp16 | 1 fp16x2 | 2 fp16+e4m3 | 3 fp16+e4m3 with SPLIT stages; HALO = 0: plain stages; 2 | 3: halo stages carrying that many taps.
// SPLIT (MODE 3, plain stages only): a 64-channel K slice travels as TWO pipeline stages -- {A_hi, W_hi} for the 4 fp16 UMMAs and
// {A_p8, W_p8} for the 4 e4m3 UMMAs -- of half the size, so the ring holds 7 stages = 224 KB of operands in flight instead of
// 3 x 64 = 192 KB.  These launches are bound by the TMA round trip (three 512-clock stages cannot cover it: measured 765 clocks per
// slice, tensor pipe 65 % busy), i.e. by the bytes in flight.
// PAIR: two CTAs of a cluster form one tcgen05 cta_group::2 unit: an MMA of M = 256 (two adjacent m-tiles) x BLOCK_N whose B operand is
// split between them -- each CTA stages its own 128 pixel rows of A and only HALF of the weight rows, so the bytes an SM takes in per flop
// fall by a quarter (these launches run at the ~85-90 B/clk an SM ingests from L2: DESIGN.md 4.1).  The even CTA issues every MMA.
template <int BLOCK_N, int MODE, int HALO, bool PAIR = false>
struct ConvCfg {
    static constexpr bool PRECISE = MODE != 0;
    static constexpr bool SPLIT = MODE == 3;
    static_assert(!SPLIT || HALO == 0, "split stages exist for the plain path only");
    static_assert(!PAIR || (HALO == 0 && (MODE == 1 || MODE == 3)) || (HALO != 0 && (MODE == 1 || MODE == 2)),
                  "CTA pairs: plain fp16x2 / split fp16+e4m3 launches and the halo launches of both parity modes");
    static constexpr int PLANES = PRECISE ? 2 : 1;                        // planes per tensor / accumulators per tile
    static constexpr int SPLANES = SPLIT ? 1 : PLANES;                    // planes carried by ONE pipeline stage
    static constexpr int A_PLANE = HALO ? A_HALO_BYTES : A_STAGE_BYTES;
    static constexpr int A_BYTES = SPLANES * A_PLANE;                     // A_hi [, A_lo]
    static constexpr int NTILE_ROWS = PLANES * BLOCK_N;                   // rows of the packed weight matrix per n-tile: W_hi rows [, W_lo rows]
    static constexpr int B_ROWS = SPLANES * BLOCK_N / (PAIR ? 2 : 1);     // weight rows carried by one stage of one CTA
    static constexpr int B_TILE_BYTES = B_ROWS * BLOCK_K * 2;             // one tap x one 64-channel slice of the weights
    static constexpr int B_STAGE_BYTES = (HALO ? HALO : 1) * B_TILE_BYTES;
    static constexpr int STAGE_BYTES = A_BYTES + B_STAGE_BYTES;
    // SPLIT lays the auxiliary block (barriers, TMEM pointer, affine staging: < 3 KB) in FRONT of the ring and relies on the 1024-byte
    // alignment of the dynamic shared-memory window (checked at run time) instead of reserving an alignment slack: 3 KB + 7 x 32 KB is
    // exactly the 227 KB limit for 128-column tiles
    static constexpr int AUX_FRONT = 3072;
    static constexpr int STAGES_RAW = SPLIT ? (SMEM_LIMIT - AUX_FRONT) / STAGE_BYTES : (SMEM_LIMIT - SMEM_AUX - 1024) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > (PAIR ? 9 : 8) ? (PAIR ? 9 : 8) : STAGES_RAW;
    static constexpr int ACC_COLS = PLANES * BLOCK_N;                     // acc_hi [, acc_lo] columns per buffer
    static constexpr int TMEM_COLS = (2 * ACC_COLS <= 32) ? 32 : (2 * ACC_COLS <= 64) ? 64 : (2 * ACC_COLS <= 128) ? 128 : (2 * ACC_COLS <= 256) ? 256 : 512;
    static constexpr int SMEM_BYTES = SPLIT ? AUX_FRONT + STAGES * STAGE_BYTES : STAGES * STAGE_BYTES + SMEM_AUX + 1024;   // +1024: manual 1024-B alignment
    static_assert(!SPLIT || (2 * STAGES + 4) * 8 + 8 <= 512, "barriers overflow their slot");
    static_assert(!SPLIT || 512 + 2 * 2 * BLOCK_N * 4 <= AUX_FRONT, "affine staging overflows the front block");
    static_assert(2 * ACC_COLS <= 512, "accumulators exceed TMEM");
    static_assert(STAGES >= 2, "not enough shared memory for a pipeline");
};

template <int BLOCK_N, int MODE, int HALO, bool PAIR = false>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_gemm_f16(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
              const __grid_constant__ CUtensorMap tmB, const ConvKParams p) {
    using Cfg = ConvCfg<BLOCK_N, MODE, HALO, PAIR>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr bool PRECISE = MODE != 0;
    constexpr bool FP8 = MODE >= 2;
    constexpr bool SPLIT = Cfg::SPLIT;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = SPLIT ? smem_raw + Cfg::AUX_FRONT
                          : reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    if (SPLIT && (smem_u32(smem_raw) & 1023u)) __trap();                // the swizzled operand tiles need 1024-byte aligned bases
    uint8_t* smem_a = smem;                                            // [STAGES][A_hi | A_lo]
    uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;                     // [STAGES][W_hi rows | W_lo rows]
    uint8_t* aux = SPLIT ? smem_raw : smem + STAGES * Cfg::STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);              // [STAGES]
    uint64_t* empty_bar = full_bar + STAGES;                            // [STAGES]
    uint64_t* tfull_bar = empty_bar + STAGES;                           // [2]
    uint64_t* tempty_bar = tfull_bar + 2;                               // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* s_affine = reinterpret_cast<float*>(aux + 512);              // [2 acc][2][BLOCK_N]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA0); prefetch_tmap(&tmA1); prefetch_tmap(&tmB);
    }
    const int CL = p.cluster;
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
    const uint16_t cmask = (uint16_t)((1u << CL) - 1u);
    if (warp == 1 && lane == 0) {
        // a slot is free again when the MMAs of EVERY CTA of the cluster have read it (peers multicast weight tiles into it)
        // PAIR: the even CTA's full barrier collects one arrival per CTA (+ both CTAs' bytes), its MMAs release the slot in both CTAs
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], PAIR ? 2u : 1u); mbar_init(&empty_bar[s], PAIR ? 1u : (uint32_t)CL); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], (PAIR ? 2 : 1) * EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 2) { if (PAIR) tmem_alloc_pair(tmem_ptr, Cfg::TMEM_COLS); else tmem_alloc(tmem_ptr, Cfg::TMEM_COLS); }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                           // the peers' barriers exist before anything is multicast at them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // work items: (parity, n-tile, group of CL consecutive m-tiles); the CTA of rank r in its cluster takes m-tile group * CL + r.  A rank
    // beyond the last m-tile recomputes that tile as a ghost (it must keep the cluster's barrier handshakes going) and stores nothing.
    const int m_groups = (p.m_tiles + CL - 1) / CL;
    const int tiles_per_parity = m_groups * p.n_tiles;
    const int num_tiles = (p.up2x ? 4 : 1) * tiles_per_parity;
    const int tile0 = blockIdx.x / CL, tile_step = gridDim.x / CL;
    const int kc_total = p.kc0 + p.kc1;
    const int num_k = p.ngroups * kc_total * (SPLIT ? 2 : 1);            // stages per output tile (p.kchunk counts stages too)

    if (warp == 0) {
        // ================================================================= TMA producer
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            // one weight tile of `rows` rows: the whole box, or this CTA's 1/CL of the rows multicast to every CTA of the cluster
            auto load_b = [&](uint8_t* dst, uint64_t* bar, int k_elem, int row, int rows) {
                if (CL > 1 && !PAIR) {
                    const int part = rows / CL;
                    tma_load_2d_mc(dst + (size_t)crank * part * (BLOCK_K * 2), &tmB, bar, k_elem, row + (int)crank * part, cmask);
                } else {
                    tma_load_2d(dst, &tmB, bar, k_elem, row);
                }
            };
            for (int tile = tile0; tile < num_tiles; tile += tile_step) {
                // parities of one m-tile are adjacent work items (concurrent CTAs): the A block comes from DRAM once, not once per parity sweep
                const int parity = p.up2x ? (tile & 3) : 0;              // (py, px) = (parity >> 1, parity & 1)
                const int tpl = p.up2x ? (tile >> 2) : tile;
                int n_tile, mt, tx, ty, tb;
                fast_divmod(p.fd_n_tiles, tpl, mt, n_tile);
                mt = min(mt * CL + (int)crank, p.m_tiles - 1);
                fast_divmod(p.fd_tiles_x, mt, mt, tx);
                fast_divmod(p.fd_tiles_y, mt, tb, ty);
                const int x0 = tx * p.tw, y0 = ty * p.th, n0 = tb * p.tn;
                const int b_row0 = parity * p.b_rows_per_parity + n_tile * Cfg::NTILE_ROWS;
                for (int g = 0; g < p.ngroups; ++g) {
                    int dy = 0, dx = 0;
                    if (HALO) { dy = -1; dx = g - 1 + (p.up2x ? (parity & 1) : 0); }          // rows y0-1 .. y0+th of column offset dx
                    else if (p.up2x) { dy = (g >> 1) - 1 + (parity >> 1); dx = (g & 1) - 1 + (parity & 1); }
                    else if (p.taps == 9) { dy = g / 3 - 1; dx = g % 3 - 1; }
                    for (int kc = 0; kc < kc_total; ++kc) {
                        const bool first = kc < p.kc0;
                        const CUtensorMap* tm = first ? &tmA0 : &tmA1;
                        const int c = (first ? kc : kc - p.kc0) * BLOCK_K;
                        if (SPLIT) {
#pragma unroll
                            for (int hf = 0; hf < 2; ++hf) {            // {A_hi, W_hi} then {A_p8, W_p8}: one stage each
                                uint8_t* sa2 = smem_a + stage * Cfg::A_BYTES;
                                uint8_t* sb2 = smem_b + stage * Cfg::B_STAGE_BYTES;
                                mbar_wait(&empty_bar[stage], phase ^ 1);
                                if (PAIR) {
                                    // both CTAs' boxes are credited to the even CTA's barrier, which expects the bytes of the pair
                                    if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * ((uint32_t)p.a_tx + (uint32_t)Cfg::B_STAGE_BYTES));
                                    else mbar_arrive_remote(&full_bar[stage], 0u);
                                    tma_load_4d_pair(sa2, tm, &full_bar[stage], c + (hf ? (first ? p.lo0 : p.lo1) : 0), x0 + dx, y0 + dy, n0);
                                    tma_load_2d_pair(sb2, &tmB, &full_bar[stage], (g * kc_total + kc) * BLOCK_K,
                                                     b_row0 + hf * BLOCK_N + (int)crank * Cfg::B_ROWS);
                                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                                    continue;
                                }
                                mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)p.a_tx + (uint32_t)Cfg::B_STAGE_BYTES);
                                tma_load_4d(sa2, tm, &full_bar[stage], c + (hf ? (first ? p.lo0 : p.lo1) : 0), x0 + dx, y0 + dy, n0);
                                load_b(sb2, &full_bar[stage], (g * kc_total + kc) * BLOCK_K, b_row0 + hf * BLOCK_N, Cfg::B_ROWS);
                                if (++stage == STAGES) { stage = 0; phase ^= 1; }
                            }
                            continue;
                        }
                        uint8_t* sa = smem_a + stage * Cfg::A_BYTES;
                        uint8_t* sb = smem_b + stage * Cfg::B_STAGE_BYTES;
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        if (PAIR) {        // fp16x2 pair: A_hi, A_lo of this CTA's pixels + its half of the W_hi rows and of the W_lo rows
                            if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * ((uint32_t)p.a_tx + (uint32_t)Cfg::B_STAGE_BYTES));
                            else mbar_arrive_remote(&full_bar[stage], 0u);
                            const int lo = first ? p.lo0 : p.lo1;
                            tma_load_4d_pair(sa, tm, &full_bar[stage], c, x0 + dx, y0 + dy, n0);
                            tma_load_4d_pair(sa + Cfg::A_PLANE, tm, &full_bar[stage], c + lo, x0 + dx, y0 + dy, n0);
#pragma unroll
                            for (int j = 0; j < (HALO ? HALO : 1); ++j) {            // per tap tile: [half of the hi rows | half of the second-plane rows]
                                const int tap = !HALO ? g : p.up2x ? j * 2 + g : j * 3 + g;
                                const int kel = (tap * kc_total + kc) * BLOCK_K;
                                uint8_t* dstb = sb + j * Cfg::B_TILE_BYTES;
                                tma_load_2d_pair(dstb, &tmB, &full_bar[stage], kel, b_row0 + (int)crank * (BLOCK_N / 2));
                                tma_load_2d_pair(dstb + (BLOCK_N / 2) * BLOCK_K * 2, &tmB, &full_bar[stage], kel, b_row0 + BLOCK_N + (int)crank * (BLOCK_N / 2));
                            }
                            if (++stage == STAGES) { stage = 0; phase ^= 1; }
                            continue;
                        }
                        mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)p.a_tx + (uint32_t)Cfg::B_STAGE_BYTES);
                        tma_load_4d(sa, tm, &full_bar[stage], c, x0 + dx, y0 + dy, n0);
                        if (PRECISE) tma_load_4d(sa + Cfg::A_PLANE, tm, &full_bar[stage], c + (first ? p.lo0 : p.lo1), x0 + dx, y0 + dy, n0);
                        if (HALO) {
#pragma unroll
                            for (int j = 0; j < HALO; ++j) {                                   // taps (dy = j-1 | j-1+py) of this column
                                const int tap = p.up2x ? j * 2 + g : j * 3 + g;
                                load_b(sb + j * Cfg::B_TILE_BYTES, &full_bar[stage], (tap * kc_total + kc) * BLOCK_K, b_row0, Cfg::B_ROWS);
                            }
                        } else {
                            load_b(sb, &full_bar[stage], (g * kc_total + kc) * BLOCK_K, b_row0, Cfg::B_ROWS);
                        }
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1 && !(PAIR && crank != 0)) {
        // ================================================================= MMA issuer (PAIR: the even CTA issues for both)
        if (elect_one()) {
            constexpr uint32_t idesc_main = umma_idesc_f16(BLOCK_M, Cfg::B_ROWS);   // N = BLOCK_N, or 2*BLOCK_N over [W_hi ; W_lo]
            constexpr uint32_t idesc_lo = umma_idesc_f16(PAIR ? 2 * BLOCK_M : BLOCK_M, BLOCK_N);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = tile0; tile < num_tiles; tile += tile_step) {
              for (int ks0 = 0; ks0 < num_k; ks0 += p.kchunk) {
                const int ks1 = min(num_k, ks0 + p.kchunk);
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * Cfg::ACC_COLS);
                for (int ks = ks0; ks < ks1; ++ks) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem_a + stage * Cfg::A_BYTES);
                    const uint32_t b_addr = smem_u32(smem_b + stage * Cfg::B_STAGE_BYTES);
                    if (HALO) {
                        // tap j of the column reads the SAME A block from tile row (row0 + j): a multiple of 8 pixel rows = 1024 B,
                        // so the swizzle phase is unchanged and only the descriptor address moves
                        const int row0 = p.up2x ? ((tile & 3) >> 1) : 0;
#pragma unroll
                        for (int j = 0; j < HALO; ++j) {
                            const uint64_t adesc = umma_desc_kmajor_sw128(a_addr + (uint32_t)((row0 + j) * p.aoff_step));
                            const uint64_t alo = umma_desc_kmajor_sw128(a_addr + (uint32_t)((row0 + j) * p.aoff_step) + Cfg::A_PLANE);
                            const uint64_t bdesc = umma_desc_kmajor_sw128(b_addr + (uint32_t)(j * Cfg::B_TILE_BYTES));
                            const uint64_t blo = umma_desc_kmajor_sw128(b_addr + (uint32_t)(j * Cfg::B_TILE_BYTES + BLOCK_N * BLOCK_K * 2));
                            if (PAIR) {
                                // this CTA's tap tile = [half of the W_hi rows | half of the second-plane rows]; three (fp16x2) or two (fp16 + e4m3) pair MMAs
                                const uint64_t b2 = umma_desc_kmajor_sw128(b_addr + (uint32_t)(j * Cfg::B_TILE_BYTES + (BLOCK_N / 2) * BLOCK_K * 2));
#pragma unroll
                                for (int k = 0; k < BLOCK_K / 16; ++k) {
                                    const uint32_t accum = (ks > ks0 || j > 0 || k > 0) ? 1u : 0u;
                                    umma_f16_pair(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_lo, accum);
                                    if (FP8) {
                                        umma_f8_pair(d_tmem + (uint32_t)BLOCK_N, alo + (uint64_t)(2 * k), b2 + (uint64_t)(2 * k), idesc_lo, accum);
                                    } else {
                                        umma_f16_pair(d_tmem + (uint32_t)BLOCK_N, adesc + (uint64_t)(2 * k), b2 + (uint64_t)(2 * k), idesc_lo, accum);
                                        umma_f16_pair(d_tmem + (uint32_t)BLOCK_N, alo + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_lo, 1u);
                                    }
                                }
                            } else if (FP8 && !p.interleave) {
#pragma unroll
                                for (int k = 0; k < BLOCK_K / 16; ++k)
                                    umma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_lo, (ks > ks0 || j > 0 || k > 0) ? 1u : 0u);
#pragma unroll
                                for (int k = 0; k < BLOCK_K / 16; ++k)
                                    umma_f8(d_tmem + (uint32_t)BLOCK_N, alo + (uint64_t)(2 * k), blo + (uint64_t)(2 * k), idesc_lo, (ks > ks0 || j > 0 || k > 0) ? 1u : 0u);
                            } else {
#pragma unroll
                            for (int k = 0; k < BLOCK_K / 16; ++k) {
                                const uint32_t accum = (ks > ks0 || j > 0 || k > 0) ? 1u : 0u;
                                if (FP8) {
                                    umma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_lo, accum);
                                    umma_f8(d_tmem + (uint32_t)BLOCK_N, alo + (uint64_t)(2 * k), blo + (uint64_t)(2 * k), idesc_lo, accum);
                                } else {
                                    umma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_main, accum);
                                    if (PRECISE) umma_f16(d_tmem + (uint32_t)BLOCK_N, alo + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_lo, 1u);
                                }
                            }
                            }
                        }
                    } else {
                    const uint64_t adesc = umma_desc_kmajor_sw128(a_addr);
                    const uint64_t bdesc = umma_desc_kmajor_sw128(b_addr);
                    const uint64_t alo = umma_desc_kmajor_sw128(a_addr + A_STAGE_BYTES);
                    const uint64_t blo = umma_desc_kmajor_sw128(b_addr + BLOCK_N * BLOCK_K * 2);     // the e4m3 weight rows (MODE 2)
                    if (PAIR && !SPLIT) {
                        // fp16x2 over a CTA pair: three M = 256 x N = BLOCK_N products per 16-element K step; this CTA's B tile holds its half
                        // of the W_hi rows followed by its half of the W_lo rows
                        const uint64_t bhi = bdesc, blo2 = umma_desc_kmajor_sw128(b_addr + (BLOCK_N / 2) * BLOCK_K * 2);
#pragma unroll
                        for (int k = 0; k < BLOCK_K / 16; ++k) {
                            const uint32_t accum = (ks > ks0 || k > 0) ? 1u : 0u;
                            umma_f16_pair(d_tmem, adesc + (uint64_t)(2 * k), bhi + (uint64_t)(2 * k), idesc_lo, accum);
                            umma_f16_pair(d_tmem + (uint32_t)BLOCK_N, adesc + (uint64_t)(2 * k), blo2 + (uint64_t)(2 * k), idesc_lo, accum);
                            umma_f16_pair(d_tmem + (uint32_t)BLOCK_N, alo + (uint64_t)(2 * k), bhi + (uint64_t)(2 * k), idesc_lo, 1u);
                        }
                    } else if (SPLIT) {
                        // even stages carry the fp16 operands (-> acc_hi), odd stages the e4m3 operands (-> acc_lo); chains start on an even stage
                        if (((ks - ks0) & 1) == 0) {
#pragma unroll
                            for (int k = 0; k < BLOCK_K / 16; ++k) {
                                if (PAIR) umma_f16_pair(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_lo, (ks > ks0 || k > 0) ? 1u : 0u);
                                else umma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_lo, (ks > ks0 || k > 0) ? 1u : 0u);
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < BLOCK_K / 16; ++k) {
                                if (PAIR) umma_f8_pair(d_tmem + (uint32_t)BLOCK_N, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_lo, (ks > ks0 + 1 || k > 0) ? 1u : 0u);
                                else umma_f8(d_tmem + (uint32_t)BLOCK_N, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_lo, (ks > ks0 + 1 || k > 0) ? 1u : 0u);
                            }
                        }
                    } else if (FP8) {
                        // hi product on the fp16 pipe, then both corrections as one K = 128 e4m3 reduction ([A_hi8 | A_lo8] . [W_lo8 ; W_hi8]);
                        // NBP_CONV_INTERLEAVE=0 issues the two kinds as two runs instead of alternating them (A/B switch; measured 1.3 % slower)
                        if (!p.interleave) {
#pragma unroll
                            for (int k = 0; k < BLOCK_K / 16; ++k)
                                umma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_lo, (ks > ks0 || k > 0) ? 1u : 0u);
#pragma unroll
                            for (int k = 0; k < BLOCK_K / 16; ++k)
                                umma_f8(d_tmem + (uint32_t)BLOCK_N, alo + (uint64_t)(2 * k), blo + (uint64_t)(2 * k), idesc_lo, (ks > ks0 || k > 0) ? 1u : 0u);
                        } else {
#pragma unroll
                            for (int k = 0; k < BLOCK_K / 16; ++k) {
                                const uint32_t accum = (ks > ks0 || k > 0) ? 1u : 0u;
                                umma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_lo, accum);
                                umma_f8(d_tmem + (uint32_t)BLOCK_N, alo + (uint64_t)(2 * k), blo + (uint64_t)(2 * k), idesc_lo, accum);
                            }
                        }
                    } else {
#pragma unroll
                    for (int k = 0; k < BLOCK_K / 16; ++k) {
                        // +32 bytes along K inside the 128-byte swizzle row (16 fp16): +2 in the encoded address
                        const uint32_t accum = (ks > ks0 || k > 0) ? 1u : 0u;
                        umma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_main, accum);
                        // A_lo * W_hi accumulates into the acc_lo columns that the UMMA above just wrote (in-order pipe)
                        if (PRECISE) umma_f16(d_tmem + (uint32_t)BLOCK_N, alo + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_lo, 1u);
                    }
                    }
                    }
                    // frees the smem slot once these MMAs have read it -- in every CTA of the cluster, whose producers write into it
                    if (PAIR) umma_commit_pair(&empty_bar[stage]);
                    else if (CL > 1) umma_commit_mc(&empty_bar[stage], cmask); else umma_commit(&empty_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (PAIR) umma_commit_pair(&tfull_bar[acc]); else umma_commit(&tfull_bar[acc]);   // (partial) accumulator complete -> epilogue(s)
                acc ^= 1; if (acc == 0) acc_phase ^= 1;
              }
            }
        }
    } else if (warp >= 2) {
        // ================================================================= epilogue (128 TMEM lanes x EPI_SPLIT column-group sets)
        const int q = warp & 3;                                 // TMEM lane quarter this warp may access
        const int cg0 = (warp - 2) >> 2;                        // first 32-column group of this warp (then every EPI_SPLIT-th)
        const int m = q * 32 + lane;                            // accumulator row = pixel of the tile
        const int et = threadIdx.x - 64;                        // 0 .. 32 * EPI_WARPS - 1
        const uint32_t s_affine_addr = smem_u32(s_affine);
        int staged_n_tile = -1;
        int acc = 0; uint32_t acc_phase = 0;
        int tsel = 0;
        const int n_chunks = (num_k + p.kchunk - 1) / p.kchunk;
        // dot epilogue: one thread contracts ALL column groups of its pixel (warps of the first column-group set; the others only keep
        // the accumulator hand-shake going): the tile's values stay in registers, no stores, so four warps are plenty
        // With a gate every warp contracts the whole tile (the dot is cheap) so that both warps of a lane quarter know the pixel's factor
        // and share the gate's channel groups.
        const bool dotm = p.dot_w != nullptr;
        const bool gatem = p.gate_src != nullptr;
        const int c_first = (dotm && gatem) ? 0 : cg0;
        const int c_step = dotm ? 1 : EPI_SPLIT;
        const int c_end = (dotm && !gatem && cg0 != 0) ? 0 : BLOCK_N / 32;
        for (int tile = tile0; tile < num_tiles; tile += tile_step) {
            float dot[8];
#pragma unroll
            for (int o = 0; o < 8; ++o) dot[o] = 0.0f;
            const int parity = p.up2x ? (tile & 3) : 0;
            const int tpl = p.up2x ? (tile >> 2) : tile;
            int n_tile, mt, tx, ty, tb;
            fast_divmod(p.fd_n_tiles, tpl, mt, n_tile);
            mt = mt * CL + (int)crank;
            const bool ghost = mt >= p.m_tiles;
            mt = min(mt, p.m_tiles - 1);
            fast_divmod(p.fd_tiles_x, mt, mt, tx);
            fast_divmod(p.fd_tiles_y, mt, tb, ty);
            const int wi = m % p.tw, hi = (m / p.tw) % p.th, ni = m / (p.tw * p.th);
            const int x = tx * p.tw + wi, y = ty * p.th + hi, nn = tb * p.tn + ni;
            const bool valid = x < p.w && y < p.h && nn < p.n && !ghost;
            // destination pixel: identity, or (2y+py, 2x+px) of the 2h x 2w image for the fused upsample
            const int oh = p.up2x ? 2 * p.h : p.h, ow = p.up2x ? 2 * p.w : p.w;
            const int oy = p.up2x ? 2 * y + (parity >> 1) : y, ox = p.up2x ? 2 * x + (parity & 1) : x;

            // stage the n-tile's affine into smem when it changes (two buffers alternate per change: every warp that still reads the
            // buffer being rewritten has yet to pass the previous change's barrier, which the writer is already behind)
            if (n_tile != staged_n_tile) {
                staged_n_tile = n_tile;
                tsel ^= 1;
                for (int i = et; i < BLOCK_N; i += 32 * EPI_WARPS) {
                    st_shared_f32(s_affine_addr + (uint32_t)(tsel * 2 * BLOCK_N + i) * 4u, p.scale[n_tile * BLOCK_N + i]);
                    st_shared_f32(s_affine_addr + (uint32_t)(tsel * 2 * BLOCK_N + BLOCK_N + i) * 4u, p.shift[n_tile * BLOCK_N + i]);
                }
                asm volatile("bar.sync 1, %0;" :: "n"(32 * EPI_WARPS) : "memory");
            }
            const uint32_t sa_addr = s_affine_addr + (uint32_t)(tsel * 2 * BLOCK_N) * 4u;

            const size_t opix = (size_t)((size_t)nn * oh + oy) * ow + ox;
            float* orow_f = reinterpret_cast<float*>(p.dst) + opix * p.dst_ld + p.dst_c_off + n_tile * BLOCK_N;

            // fp32 -> hi plane (fp16) + second plane -> 16-byte stores.  `pix` = first element of the pixel, `ch` = channel of x[0]
            // inside the pixel, `lo_off` = offset of the second plane; fmt 1: fp16 (x - hi) * 2048 at the same channel index,
            // fmt 2: per 64-channel group 64 bytes e4m3(x) then 64 bytes e4m3((x - hi) * 2048)
            auto split_store = [&](__half* pix, int ch, int lo_off, int fmt, const float* x) {
                uint32_t packed[16];
                if (PRECISE && fmt == 2) {          // hi plane + e4m3 pair plane (the fp16 lo values are never formed)
                    uint32_t p8h[8], p8l[8];
                    bool sat = false;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const __half2 h = __floats2half2_rn(x[2 * j], x[2 * j + 1]);
                        packed[j] = *reinterpret_cast<const uint32_t*>(&h);
                        const float2 hf = __half22float2(h);
                        const float r0 = (x[2 * j] - hf.x) * (2048.0f * NBP_E4M3_ACT_SCALE), r1 = (x[2 * j + 1] - hf.y) * (2048.0f * NBP_E4M3_ACT_SCALE);
                        const uint32_t qh = e4m3x2(x[2 * j] * NBP_E4M3_ACT_SCALE, x[2 * j + 1] * NBP_E4M3_ACT_SCALE);
                        const uint32_t ql = e4m3x2(r0, r1);
                        sat |= fmaxf(fabsf(x[2 * j]), fabsf(x[2 * j + 1])) > NBP_E4M3_MAX / NBP_E4M3_ACT_SCALE;
                        if (j & 1) { p8h[j >> 1] |= qh << 16; p8l[j >> 1] |= ql << 16; } else { p8h[j >> 1] = qh; p8l[j >> 1] = ql; }
                    }
                    st_global_v8(pix + ch, packed); st_global_v8(pix + ch + 16, packed + 8);
                    uint8_t* g = reinterpret_cast<uint8_t*>(pix + lo_off) + (ch >> 6) * 128 + (ch & 63);
                    st_global_v8(g, p8h); st_global_v8(g + 64, p8l);
                    if (sat && p.sat_count) atomicAdd(p.sat_count, 1ull);          // rare by construction: out-of-range inputs only
                    return;
                }
                uint32_t packed_lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const __half2 h = __floats2half2_rn(x[2 * j], x[2 * j + 1]);
                    packed[j] = *reinterpret_cast<const uint32_t*>(&h);
                    if (PRECISE) {
                        const float2 hf = __half22float2(h);
                        const __half2 l = __floats2half2_rn((x[2 * j] - hf.x) * 2048.0f, (x[2 * j + 1] - hf.y) * 2048.0f);
                        packed_lo[j] = *reinterpret_cast<const uint32_t*>(&l);
                    }
                }
                st_global_v8(pix + ch, packed); st_global_v8(pix + ch + 16, packed + 8);
                if (PRECISE) { st_global_v8(pix + ch + lo_off, packed_lo); st_global_v8(pix + ch + lo_off + 16, packed_lo + 8); }
            };
            // affine + activation + store of 32 consecutive output channels held as fp32 in v[]
            const float lower = p.relu ? 0.0f : -65504.0f;
            auto finish = [&](int c, const float* v) {
                float sc[32], sh[32];                // this group's scale / shift
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    ld_shared_f32x4(sa_addr + (uint32_t)(c * 32 + 4 * j) * 4u, &sc[4 * j]);
                    ld_shared_f32x4(sa_addr + (uint32_t)(BLOCK_N + c * 32 + 4 * j) * 4u, &sh[4 * j]);
                }
                if (p.out_f32) {                     // dgrad: fp32 NHWC destination
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint32_t o8[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) o8[e] = __float_as_uint(fmaf(v[8 * j + e], sc[8 * j + e], sh[8 * j + e]));
                            st_global_v8(orow_f + c * 32 + 8 * j, o8);
                        }
                    }
                    return;
                }
                float a[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    a[j] = fminf(fmaxf(fmaf(v[j], sc[j], sh[j]), lower), 65504.0f);      // ReLU (lower = 0) and the fp16 range clamp in one
                }
                if (dotm) {                          // channels in ascending order, one fused multiply-add each: deterministic
#pragma unroll
                    for (int o = 0; o < 8; ++o) {
                        if (o < p.dot_n) {
                            const float4* dw = reinterpret_cast<const float4*>(p.dot_w + o * (p.n_tiles * BLOCK_N) + n_tile * BLOCK_N + c * 32);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 q = __ldg(dw + j);
                                dot[o] = fmaf(a[4 * j], q.x, dot[o]); dot[o] = fmaf(a[4 * j + 1], q.y, dot[o]);
                                dot[o] = fmaf(a[4 * j + 2], q.z, dot[o]); dot[o] = fmaf(a[4 * j + 3], q.w, dot[o]);
                            }
                        }
                    }
                    return;
                }
                if (valid) split_store(p.dst + opix * p.dst_ld, p.dst_c_off + n_tile * BLOCK_N + c * 32, p.dst_lo_off, p.dst_fmt, a);
                if (p.pool) {
                    // fused nn.MaxPool2d(2,2): the 2x2 window of a pixel lives in lanes {l, l^1, l^tw, l^(tw+1)} of this warp (tile rows
                    // are tw <= 16 pixels wide and a warp holds 32 consecutive tile pixels); max commutes with the monotonic hi/lo split
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float t = fmaxf(a[j], __shfl_xor_sync(0xffffffffu, a[j], 1));
                        a[j] = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, p.tw));
                    }
                    if (valid && !((wi | hi) & 1)) {
                        const size_t ppix = ((size_t)nn * (p.h >> 1) + (y >> 1)) * (size_t)(p.w >> 1) + (x >> 1);
                        split_store(p.pool + ppix * p.pool_ld, n_tile * BLOCK_N + c * 32, p.pool_lo_off, p.pool_fmt, a);
                    }
                }
            };
            // one 32-column group of the current TMEM buffer as fp32 (acc_hi + acc_lo/2048 in the fp16x2 mode)
            auto load_group = [&](uint32_t t_row, int c, float* out) {
                uint32_t v[32];
                tmem_ld_32x32(t_row + (uint32_t)(c * 32), v);
                if (PRECISE) {
                    uint32_t vl[32];
                    tmem_ld_32x32(t_row + (uint32_t)(BLOCK_N + c * 32), vl);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) out[j] = fmaf(__uint_as_float(vl[j]), p.lo_scale, __uint_as_float(v[j]));
                } else {
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) out[j] = __uint_as_float(v[j]);
                }
            };
            auto release = [&]() {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if (PAIR) mbar_arrive_remote(&tempty_bar[acc], 0u); else mbar_arrive(&tempty_bar[acc]); }
                acc ^= 1; if (acc == 0) acc_phase ^= 1;
            };

            if (n_chunks == 1) {
                mbar_wait(&tfull_bar[acc], acc_phase);
                tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * Cfg::ACC_COLS);
#pragma unroll 1
                for (int c = c_first; c < c_end; c += c_step) {
                    float v[32];
                    load_group(t_row, c, v);
                    finish(c, v);
                }
                release();
                if (dotm) {
                    float zs[8];
#pragma unroll
                    for (int o = 0; o < 8; ++o) {
                        zs[o] = fmaf(dot[o], p.dot_scale, p.dot_shift + ((p.dot_bias && o < p.dot_n) ? __ldg(p.dot_bias + o) : 0.0f));
                        if (p.dot_sigmoid) zs[o] = 1.0f / (1.0f + expf(-zs[o]));
                    }
                    const float z = zs[0];
                    if (p.dot_out && cg0 == 0 && valid) {
                        // NCHW fp32 [n][dot_n][oh][ow] (dot_n = 1: one value per pixel)
                        const size_t ohw = (size_t)oh * ow, pin = (size_t)oy * ow + ox;
                        float vmax = -INFINITY;
#pragma unroll
                        for (int o = 0; o < 8; ++o) {
                            if (o < p.dot_n) { p.dot_out[((size_t)nn * p.dot_n + o) * ohw + pin] = zs[o]; vmax = fmaxf(vmax, zs[o]); }
                        }
                        if (p.dot_max) p.dot_max[(size_t)nn * ohw + pin] = vmax;
                    }
                    if (gatem && valid) {
                        // this pixel's gate_c channels times z, 32 channels at a time, alternate groups per warp of the lane quarter
                        const __half* xp = p.gate_src + opix * p.gate_ld;
                        for (int gg = cg0; gg < p.gate_c / 32; gg += EPI_SPLIT) {
                            uint32_t h[16];
                            ld_global_nc_v8(xp + gg * 32, h); ld_global_nc_v8(xp + gg * 32 + 16, h + 8);
                            float x[32];
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&h[j]));
                                x[2 * j] = t.x; x[2 * j + 1] = t.y;
                            }
                            if (PRECISE && p.gate_fmt == 2) {       // second plane: 64 e4m3(x / 8) bytes, then 64 e4m3((x - hi) * 2048 / 8) bytes per 64 channels
                                uint32_t q[8];
                                ld_global_nc_v8(reinterpret_cast<const uint8_t*>(xp + p.gate_lo) + (gg >> 1) * 128 + 64 + (gg & 1) * 32, q);
#pragma unroll
                                for (int j = 0; j < 16; ++j) {
                                    uint32_t h2;
                                    asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(h2) : "h"((uint16_t)(q[j >> 1] >> (16 * (j & 1)))));
                                    const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&h2));
                                    x[2 * j] = fmaf(t.x, 1.0f / (2048.0f * NBP_E4M3_ACT_SCALE), x[2 * j]);
                                    x[2 * j + 1] = fmaf(t.y, 1.0f / (2048.0f * NBP_E4M3_ACT_SCALE), x[2 * j + 1]);
                                }
                            } else if (PRECISE) {
                                uint32_t l[16];
                                ld_global_nc_v8(xp + p.gate_lo + gg * 32, l); ld_global_nc_v8(xp + p.gate_lo + gg * 32 + 16, l + 8);
#pragma unroll
                                for (int j = 0; j < 16; ++j) {
                                    const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&l[j]));
                                    x[2 * j] = fmaf(t.x, 1.0f / 2048.0f, x[2 * j]); x[2 * j + 1] = fmaf(t.y, 1.0f / 2048.0f, x[2 * j + 1]);
                                }
                            }
#pragma unroll
                            for (int j = 0; j < 32; ++j) x[j] = fminf(fmaxf(x[j] * z, -65504.0f), 65504.0f);
                            split_store(p.dst + opix * p.dst_ld, p.dst_c_off + gg * 32, p.dst_lo_off, p.dst_fmt, x);
                        }
                    }
                }
            } else {
                // long reductions: every K chunk is summed inside TMEM (truncating accumulator), the chunks are summed here
                // in fp32 round-to-nearest -- error grows with sqrt(chunks) instead of linearly with K
                constexpr int MY_GROUPS = (BLOCK_N / 32 + EPI_SPLIT - 1) / EPI_SPLIT;       // column groups cg0, cg0 + EPI_SPLIT, ...
                float racc[MY_GROUPS * 32];
#pragma unroll
                for (int j = 0; j < MY_GROUPS * 32; ++j) racc[j] = 0.0f;
                for (int ch = 0; ch < n_chunks; ++ch) {
                    mbar_wait(&tfull_bar[acc], acc_phase);
                    tc_fence_after();
                    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * Cfg::ACC_COLS);
#pragma unroll
                    for (int g = 0; g < MY_GROUPS; ++g) {
                        const int c = cg0 + g * EPI_SPLIT;
                        if (c < BLOCK_N / 32) {
                            float v[32];
                            load_group(t_row, c, v);
#pragma unroll
                            for (int j = 0; j < 32; ++j) racc[g * 32 + j] += v[j];
                        }
                    }
                    release();
                }
#pragma unroll
                for (int g = 0; g < MY_GROUPS; ++g)
                    if (cg0 + g * EPI_SPLIT < BLOCK_N / 32) finish(cg0 + g * EPI_SPLIT, &racc[g * 32]);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                           // no CTA leaves while a peer may still multicast into it or arrive on its barriers
    if (warp == 2) { if (PAIR) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS); else tmem_dealloc(tmem_base, Cfg::TMEM_COLS); }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// NHWC fp16 activation [n][h][w][ld] (first `c` channels used) -> 4-D map (c, w, h, n), box (64, tw, th, tn), 128B swizzle
static int make_act_map(CUtensorMap* m, const void* ptr, int c, int ld, int n, int h, int w, int tw, int th, int tn) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return NBP_ERR_UNSUPPORTED; }
    cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)w * ld * 2, (cuuint64_t)h * w * ld * 2};
    cuuint32_t box[4] = {BLOCK_K, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tn};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(activation c=%d ld=%d n=%d h=%d w=%d box=%d,%d,%d) failed: %d", c, ld, n, h, w, tw, th, tn, (int)r); return NBP_ERR_INVALID; }
    return NBP_OK;
}

static int make_weight_map(CUtensorMap* m, const void* ptr, int k_total, int c_out, int block_n) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return NBP_ERR_UNSUPPORTED; }
    cuuint64_t dims[2] = {(cuuint64_t)k_total, (cuuint64_t)c_out};
    cuuint64_t strides[1] = {(cuuint64_t)k_total * 2};
    cuuint32_t box[2] = {BLOCK_K, (cuuint32_t)block_n};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weight k=%d c_out=%d) failed: %d", k_total, c_out, (int)r); return NBP_ERR_INVALID; }
    return NBP_OK;
}

// ---- optional per-launch timing of conv_gemm_f16 (bench.py roofline): CUDA events recorded on the launch stream
struct ConvProfile {
    bool enabled = false;
    std::vector<cudaEvent_t> ev;       // pairs (start, stop)
    size_t used = 0;                   // events consumed
    double flops = 0.0;                // algorithmic 2*M*N*K of the recorded launches
    uint64_t dropped = 0;
};
static ConvProfile g_prof;

static int pow2_floor(int v) { int p = 1; while (p * 2 <= v) p *= 2; return p; }
static int pow2_ceil(int v) { int p = 1; while (p < v) p *= 2; return p; }

template <int BLOCK_N, int MODE, int HALO = 0, bool PAIR = false>
static int launch_conv(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const ConvKParams& kp, int sms, cudaStream_t st) {
    using Cfg = ConvCfg<BLOCK_N, MODE, HALO, PAIR>;
    static bool attr_set_d[NBP_MAX_DEVICES] = {};
    static int max_clusters2_d[NBP_MAX_DEVICES] = {};   // co-resident 2-CTA clusters of this instantiation (0 = not queried yet)
    const int dslot = device_slot();
    bool& attr_set = attr_set_d[dslot];
    int& max_clusters2 = max_clusters2_d[dslot];
    auto kern = conv_gemm_f16<BLOCK_N, MODE, HALO, PAIR>;
    if (!attr_set) {
        int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES),
                            "cudaFuncSetAttribute(conv_gemm_f16)");
        if (rc) return rc;
        attr_set = true;
    }
    const int CL = kp.cluster;
    const int work = ((kp.m_tiles + CL - 1) / CL) * kp.n_tiles * (kp.up2x ? 4 : 1);       // work items of one cluster (or CTA)
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(CONV_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    int slots = sms;
    if (CL > 1) {
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        if (!max_clusters2) {
            cfg.gridDim = dim3((unsigned)(sms / CL * CL));
            int n = 0;
            int rc = check_cuda(cudaOccupancyMaxActiveClusters(&n, kern, &cfg), "cudaOccupancyMaxActiveClusters(conv_gemm_f16)");
            if (rc) return rc;
            if (n < 1) { set_error("conv_gemm_f16: no 2-CTA cluster fits on this device"); return NBP_ERR_UNSUPPORTED; }
            max_clusters2 = n;
        }
        slots = max_clusters2;
    }
    cfg.gridDim = dim3((unsigned)((work < slots ? work : slots) * CL));
    const bool prof = g_prof.enabled && g_prof.used + 2 <= g_prof.ev.size();
    if (g_prof.enabled && !prof) ++g_prof.dropped;
    if (prof) cudaEventRecord(g_prof.ev[g_prof.used], st);
    int rc = check_cuda(cudaLaunchKernelEx(&cfg, kern, a0, a1, b, kp), "conv_gemm_f16 launch");
    if (prof) {
        cudaEventRecord(g_prof.ev[g_prof.used + 1], st);
        g_prof.used += 2;
        // algorithmic flops of the reference op: a fused upsample+3x3 counts as the 3x3 conv on the 2h x 2w image
        const double pix = (double)kp.n * kp.h * kp.w * (kp.up2x ? 4.0 : 1.0), ktaps = kp.up2x ? 9.0 : (double)kp.taps;
        g_prof.flops += 2.0 * pix * (double)(kp.n_tiles * BLOCK_N) * ktaps * (double)((kp.kc0 + kp.kc1) * BLOCK_K);
    }
    count_launch();
    return rc;
}

}  // namespace nbp

using namespace nbp;

extern "C" int nbp_conv_fwd(const nbp_conv_desc* d, void* stream) {
    if (!d) return invalid("nbp_conv_fwd: null descriptor");
    const bool dot_any = d->dot_out != nullptr || d->gate_src != nullptr;   // dot epilogue requested
    const bool gate = d->gate_src != nullptr;
    const bool dotm = dot_any && !gate;                                      // dot epilogue that writes no activation: dst is unused
    if (!d->src0 || !d->weight || !d->scale || !d->shift || (!d->dst && !dotm)) return invalid("nbp_conv_fwd: null pointer in descriptor");
    if (dot_any && (!d->dot_w || d->out_f32 || d->pool_dst || (d->c_out != 32 && d->c_out != 64 && d->c_out != 128) || ((uintptr_t)d->dot_w & 15)))
        return invalid("nbp_conv_fwd: the dot epilogue needs dot_w (16-byte aligned), c_out = 32, 64 or 128 (one n-tile; got %d), no fp32 / pooled output", d->c_out);
    const int dot_n = d->dot_n > 0 ? d->dot_n : 1;
    if (dot_any && (dot_n > 8 || (gate && dot_n != 1) || (d->dot_max && !d->dot_out)))
        return invalid("nbp_conv_fwd: dot_n must be 1..8 (1 with a gate); dot_max needs dot_out (dot_n=%d)", d->dot_n);
    if (gate) {
        const int gfmt = d->precise;                                          // the gate source is in the sources' format
        if (d->up2x || d->gate_c <= 0 || d->gate_c % 32 || d->gate_ld % 16 || ((uintptr_t)d->gate_src & 31) ||
            (gfmt && (d->gate_lo < d->gate_c || d->gate_lo % 16 || d->gate_lo + d->gate_c > d->gate_ld)) || (!gfmt && d->gate_c > d->gate_ld) ||
            (gfmt == 2 && (d->gate_lo % 64 || d->gate_c % 64)))
            return invalid("nbp_conv_fwd: bad gate source layout c=%d ld=%d lo=%d (32-byte pieces; e4m3 pair planes in 64-channel groups; not with up2x)",
                           d->gate_c, d->gate_ld, d->gate_lo);
    }
    const int c_written = gate ? d->gate_c : d->c_out;                       // channels the epilogue stores per pixel
    if (d->taps != 1 && d->taps != 9 && !(d->taps == 4 && d->up2x)) return invalid("nbp_conv_fwd: taps must be 1 or 9, or 4 with up2x (got %d)", d->taps);
    if (d->up2x && d->taps != 4) return invalid("nbp_conv_fwd: up2x needs the 4-tap parity weights");
    if (d->n <= 0 || d->h <= 0 || d->w <= 0) return invalid("nbp_conv_fwd: bad image dims n=%d h=%d w=%d", d->n, d->h, d->w);
    if (d->c0 <= 0 || d->c0 % BLOCK_K || d->c1 < 0 || d->c1 % BLOCK_K || (d->c1 > 0 && !d->src1))
        return invalid("nbp_conv_fwd: source channels must be positive multiples of 64 (c0=%d c1=%d)", d->c0, d->c1);
    if (d->precise < 0 || d->precise > 2) return invalid("nbp_conv_fwd: precise must be 0 (fp16), 1 (fp16x2) or 2 (fp16 + e4m3 corrections), got %d", d->precise);
    const bool precise = d->precise != 0;
    const bool fp8 = d->precise == 2;
    if (fp8 && d->out_f32) return invalid("nbp_conv_fwd: the fp16+e4m3 mode has no fp32 output (it is an eval-path format)");
    if (fp8 && !(d->w_lo_scale > 0.0f)) return invalid("nbp_conv_fwd: precise = 2 needs w_lo_scale = 1 / (2048 s) > 0");
    const int dst_fmt = d->dst_fmt ? d->dst_fmt : d->precise, pool_fmt = d->pool_fmt ? d->pool_fmt : d->precise;   // 0 = the mode's own format
    if (!dotm && precise && !d->out_f32 && dst_fmt != 1 && dst_fmt != 2) return invalid("nbp_conv_fwd: dst_fmt must be 0 (as the sources), 1 (fp16 lo plane) or 2 (e4m3 pair plane)");
    if (precise && d->pool_dst && pool_fmt != 1 && pool_fmt != 2) return invalid("nbp_conv_fwd: pool_fmt must be 0, 1 or 2");
    if (precise && ((!dotm && dst_fmt == 2 && (d->dst_c_off % 32 || d->dst_lo_off % 64)) || (d->pool_dst && pool_fmt == 2 && d->pool_lo_off % 64)))
        return invalid("nbp_conv_fwd: e4m3 pair planes need 64-channel aligned plane offsets");
    const int span0 = precise ? d->lo0 + d->c0 : d->c0, span1 = precise ? d->lo1 + d->c1 : d->c1;
    if (precise && (d->lo0 < d->c0 || d->lo0 % 8 || (d->c1 > 0 && (d->lo1 < d->c1 || d->lo1 % 8)) ||
                    (!d->out_f32 && !dotm && (d->dst_lo_off < c_written || d->dst_lo_off % 8))))
        return invalid("nbp_conv_fwd: lo-plane offsets must be >= the channel count and multiples of 8");
    if (d->ld0 < span0 || d->ld0 % 8 || (d->c1 > 0 && (d->ld1 < span1 || d->ld1 % 8)))
        return invalid("nbp_conv_fwd: source pixel strides must cover the planes and be multiples of 8");
    if (d->c_out <= 0 || d->c_out % 32) return invalid("nbp_conv_fwd: c_out must be a positive multiple of 32 (got %d)", d->c_out);
    if (!dotm && (d->dst_ld % 8 || d->dst_c_off % 8 || d->dst_c_off + ((precise && !d->out_f32) ? d->dst_lo_off : 0) + c_written > d->dst_ld))
        return invalid("nbp_conv_fwd: bad destination channel layout ld=%d off=%d lo_off=%d c_out=%d", d->dst_ld, d->dst_c_off, d->dst_lo_off, d->c_out);
    if (((uintptr_t)d->src0 | (uintptr_t)d->src1 | (uintptr_t)d->weight) & 15)
        return invalid("nbp_conv_fwd: source and weight pointers must be 16-byte aligned");
    // the epilogue writes 32-byte pieces (16 fp16 channels / 8 fp32 values / 32 e4m3 bytes per store)
    if (((uintptr_t)d->dst | (uintptr_t)d->pool_dst) & 31) return invalid("nbp_conv_fwd: destination pointers must be 32-byte aligned");
    if (!dotm && !d->out_f32 && (d->dst_ld % 16 || d->dst_c_off % 16 || (precise && d->dst_lo_off % 16)))
        return invalid("nbp_conv_fwd: fp16 destinations need ld, channel offset and plane offset in multiples of 16 (ld=%d off=%d lo_off=%d)",
                       d->dst_ld, d->dst_c_off, d->dst_lo_off);
    if (d->pool_dst && (d->pool_ld % 16 || (precise && d->pool_lo_off % 16)))
        return invalid("nbp_conv_fwd: the pooled destination needs ld and plane offset in multiples of 16 (ld=%d lo_off=%d)", d->pool_ld, d->pool_lo_off);

    const int block_n = (d->c_out % 128 == 0) ? 128 : (d->c_out % 64 == 0) ? 64 : 32;
    ConvKParams kp{};
    kp.n = d->n; kp.h = d->h; kp.w = d->w;
    kp.tw = d->w >= 16 ? 16 : pow2_floor(d->w);
    int th = BLOCK_M / kp.tw; const int hc = pow2_ceil(d->h);
    kp.th = th < hc ? th : hc;
    kp.tn = BLOCK_M / (kp.tw * kp.th);
    kp.tiles_x = (d->w + kp.tw - 1) / kp.tw; kp.tiles_y = (d->h + kp.th - 1) / kp.th; kp.tiles_n = (d->n + kp.tn - 1) / kp.tn;
    kp.m_tiles = kp.tiles_x * kp.tiles_y * kp.tiles_n; kp.n_tiles = d->c_out / block_n;
    kp.fd_n_tiles = make_fastdiv((uint32_t)kp.n_tiles); kp.fd_tiles_x = make_fastdiv((uint32_t)kp.tiles_x); kp.fd_tiles_y = make_fastdiv((uint32_t)kp.tiles_y);
    kp.taps = d->taps; kp.kc0 = d->c0 / BLOCK_K; kp.kc1 = d->c1 / BLOCK_K;
    kp.lo0 = d->lo0; kp.lo1 = d->lo1;
    kp.lo_scale = fp8 ? d->w_lo_scale / NBP_E4M3_ACT_SCALE : 1.0f / 2048.0f;
    kp.sat_count = (unsigned long long*)d->sat_count;
    kp.dst_fmt = dst_fmt; kp.pool_fmt = pool_fmt;
    kp.up2x = d->up2x ? 1 : 0;
    kp.out_f32 = d->out_f32 ? 1 : 0;
    // ---- vertical halo reuse (64- and 32-channel tiles, whose K steps are too short to hide the operand latency): tiles inside one
    // image whose rows are whole 8-pixel swizzle groups; training's short accumulation chains (k_chunk 1..2) keep the plain path
    static int halo_env = -1, kchunk_env = -1, inter_env = -1, k8_env = -1;
    if (inter_env < 0) { const char* e = getenv("NBP_CONV_INTERLEAVE"); inter_env = e ? atoi(e) : 1; }
    if (k8_env < 0) { const char* e = getenv("NBP_CONV_KCHUNK_E4M3"); k8_env = e ? atoi(e) : 0; }
    kp.interleave = inter_env;
    if (halo_env < 0) { const char* e = getenv("NBP_CONV_HALO"); halo_env = e ? atoi(e) : 1; }
    if (kchunk_env < 0) { const char* e = getenv("NBP_CONV_KCHUNK"); kchunk_env = e ? atoi(e) : 8; }
    // K slices (64 elements each) per in-TMEM accumulation chain; 0 = unbounded.  The fp16+e4m3 mode keeps whole reductions in TMEM by
    // default: the truncating accumulator costs 1.2e-9 * K relative (1.1e-5 at K = 9216), an order below that mode's operand error
    const int want = d->k_chunk > 0 ? d->k_chunk : (fp8 ? k8_env : kchunk_env);
    // 128-column tiles too, as CTA pairs: a pair stages only half of the weight rows per SM, so a halo stage (A block of one column
    // offset + the weight tiles of its 3 taps) is 88 KB and two fit; per 64-channel slice and 3 taps an SM then takes in 88 KB instead
    // of 3 x 48 KB.  Measured on the 32-scene forward: -7.8 % conv time, tensor pipe 82-92 % busy on those layers (was 70-83 %);
    // NBP_CONV_HALO128=0 is the A/B switch back to plain split stages
    static int halo128_env = -1;
    if (halo128_env < 0) { const char* e = getenv("NBP_CONV_HALO128"); halo128_env = e ? atoi(e) : 1; }
    const bool halo = halo_env && precise && (block_n <= 64 || (block_n == 128 && halo128_env && kp.m_tiles >= 2)) && (d->taps == 9 || d->up2x) &&
                      kp.tn == 1 && kp.tw >= 8 && (kp.th + 2) * kp.tw <= HALO_ROWS && !(want > 0 && want < 3);
    static int split_env = -1, cluster_env = -1;
    if (cluster_env < 0) { const char* e = getenv("NBP_CONV_CLUSTER"); cluster_env = e ? atoi(e) : 2; }
    kp.cluster = (cluster_env >= 2 && kp.m_tiles >= 2) ? 2 : 1;
    if (split_env < 0) { const char* e = getenv("NBP_CONV_SPLIT"); split_env = e ? atoi(e) : 1; }
    const bool split = fp8 && !halo && split_env;                     // two half-size stages per K slice (ConvCfg, MODE 3)
    static int pair_env = -1;
    if (pair_env < 0) { const char* e = getenv("NBP_CONV_PAIR"); pair_env = e ? atoi(e) : 1; }
    // cta_group::2 tiles (ConvCfg, PAIR): the 128-column plain launches of the fp16+e4m3 (split) and fp16x2 modes
    const bool pair = pair_env && kp.m_tiles >= 2 &&
                      ((block_n == 128 && !halo && (split || d->precise == 1)) || (block_n >= 64 && halo && precise));
    if (pair) kp.cluster = 2;
    const int planes_ = split ? 1 : precise ? 2 : 1;                  // planes carried by one pipeline stage
    kp.gtaps = halo ? (d->up2x ? 2 : 3) : 1;
    kp.ngroups = halo ? (d->up2x ? 2 : 3) : d->taps;
    kp.a_tx = planes_ * (halo ? kp.th + 2 : kp.th) * kp.tw * kp.tn * BLOCK_K * 2;
    kp.aoff_step = kp.tw * BLOCK_K * 2;
    {   // kp.kchunk = STAGES per accumulation chain (a halo stage carries gtaps K slices)
        const int smult = split ? 2 : 1;                              // stages per K slice
        const int total = kp.ngroups * (kp.kc0 + kp.kc1) * smult;
        int want_st = want <= 0 ? total : (want + kp.gtaps - 1) / kp.gtaps * smult;
        kp.kchunk = !precise ? total : want_st;
        if (kp.kchunk > total) kp.kchunk = total;
        if (kp.kchunk < 1) kp.kchunk = 1;
        // A chain only slightly longer than the target (K = 9 slices of a 64-channel 3x3 layer vs 8) stays single: cutting it as
        // 8 + 1 makes the epilogue warps fold two chunks per tile for nothing (measured -3..-13 % on those layers); spreading
        // longer reductions evenly (18 as 6+6+6 instead of 8+8+2) was measured slower and is not done.
        if (precise && want > 0 && total > kp.kchunk && total / smult * kp.gtaps <= want + want / 4) kp.kchunk = total;
        if (dot_any) kp.kchunk = total;           // the dot epilogue contracts one whole accumulator per tile
    }
    kp.b_rows_per_parity = (precise ? 2 : 1) * d->c_out;
    kp.scale = d->scale; kp.shift = d->shift; kp.relu = d->relu;
    kp.dst = (__half*)d->dst; kp.dst_ld = d->dst_ld; kp.dst_c_off = d->dst_c_off; kp.dst_lo_off = d->dst_lo_off;
    kp.pool = (__half*)d->pool_dst; kp.pool_ld = d->pool_ld; kp.pool_lo_off = d->pool_lo_off;
    kp.dot_w = dot_any ? d->dot_w : nullptr; kp.dot_out = d->dot_out; kp.dot_scale = d->dot_scale; kp.dot_shift = d->dot_shift; kp.dot_sigmoid = d->dot_sigmoid ? 1 : 0;
    kp.dot_n = dot_n; kp.dot_bias = dot_any ? d->dot_bias : nullptr; kp.dot_max = dot_any ? d->dot_max : nullptr;
    kp.gate_src = (const __half*)d->gate_src; kp.gate_c = d->gate_c; kp.gate_ld = d->gate_ld; kp.gate_lo = d->gate_lo; kp.gate_fmt = d->precise;
    if (d->pool_dst) {
        if (d->up2x || d->out_f32) return invalid("nbp_conv_fwd: pool_dst cannot be combined with up2x / out_f32");
        if ((d->h | d->w) & 1) return invalid("nbp_conv_fwd: pool_dst needs even h and w (got %d x %d)", d->h, d->w);
        if (kp.tw < 2 || kp.tw > 16 || kp.th < 2) return invalid("nbp_conv_fwd: pool_dst needs a tile of >= 2 rows of 2..16 pixels (w=%d h=%d)", d->w, d->h);
        if (d->pool_ld % 8 || d->pool_lo_off % 8 || (precise ? d->pool_lo_off : 0) + d->c_out > d->pool_ld || (precise && d->pool_lo_off < d->c_out) ||
            ((uintptr_t)d->pool_dst & 15))
            return invalid("nbp_conv_fwd: bad pooled destination layout ld=%d lo_off=%d c_out=%d", d->pool_ld, d->pool_lo_off, d->c_out);
    }
    if (kp.tn > 256 || kp.th > 256) return invalid("nbp_conv_fwd: image too small for a 128-pixel tile (w=%d h=%d)", d->w, d->h);

    CUtensorMap a0, a1, b;
    const int box_rows = halo ? kp.th + 2 : kp.th;
    int rc = make_act_map(&a0, d->src0, span0, d->ld0, d->n, d->h, d->w, kp.tw, box_rows, kp.tn);
    if (rc) return rc;
    if (d->c1 > 0) rc = make_act_map(&a1, d->src1, span1, d->ld1, d->n, d->h, d->w, kp.tw, box_rows, kp.tn);
    else a1 = a0;
    if (rc) return rc;
    // weights: fast [c_out][K]; precise [(c_out/block_n) tiles][W_hi rows ; W_lo rows][K]
    const int planes = precise ? 2 : 1;
    rc = make_weight_map(&b, d->weight, d->taps * (d->c0 + d->c1), (d->up2x ? 4 : 1) * planes * d->c_out, pair ? block_n / 2 : (split ? 1 : planes) * block_n / kp.cluster);
    if (rc) return rc;

    static int sms_d[NBP_MAX_DEVICES] = {};
    int& sms = sms_d[device_slot()];
    if (!sms) {
        int dev = 0;
        rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
        if (rc) return rc;
        rc = check_cuda(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute");
        if (rc) return rc;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (pair && halo && block_n == 128) {
        if (fp8) return d->up2x ? launch_conv<128, 2, 2, true>(a0, a1, b, kp, sms, st) : launch_conv<128, 2, 3, true>(a0, a1, b, kp, sms, st);
        return d->up2x ? launch_conv<128, 1, 2, true>(a0, a1, b, kp, sms, st) : launch_conv<128, 1, 3, true>(a0, a1, b, kp, sms, st);
    }
    if (pair && halo) {
        if (fp8) return d->up2x ? launch_conv<64, 2, 2, true>(a0, a1, b, kp, sms, st) : launch_conv<64, 2, 3, true>(a0, a1, b, kp, sms, st);
        return d->up2x ? launch_conv<64, 1, 2, true>(a0, a1, b, kp, sms, st) : launch_conv<64, 1, 3, true>(a0, a1, b, kp, sms, st);
    }
    if (pair) return split ? launch_conv<128, 3, 0, true>(a0, a1, b, kp, sms, st) : launch_conv<128, 1, 0, true>(a0, a1, b, kp, sms, st);
    if (split) {
        switch (block_n) {
            case 128: return launch_conv<128, 3>(a0, a1, b, kp, sms, st);
            case 64:  return launch_conv<64, 3>(a0, a1, b, kp, sms, st);
            default:  return launch_conv<32, 3>(a0, a1, b, kp, sms, st);
        }
    }
    if (fp8) {
        switch (block_n) {
            case 128: return launch_conv<128, 2>(a0, a1, b, kp, sms, st);
            case 64:  return !halo ? launch_conv<64, 2>(a0, a1, b, kp, sms, st)
                             : d->up2x ? launch_conv<64, 2, 2>(a0, a1, b, kp, sms, st) : launch_conv<64, 2, 3>(a0, a1, b, kp, sms, st);
            default:  return !halo ? launch_conv<32, 2>(a0, a1, b, kp, sms, st)
                             : d->up2x ? launch_conv<32, 2, 2>(a0, a1, b, kp, sms, st) : launch_conv<32, 2, 3>(a0, a1, b, kp, sms, st);
        }
    }
    if (precise) {
        switch (block_n) {
            case 128: return launch_conv<128, 1>(a0, a1, b, kp, sms, st);
            case 64:  return !halo ? launch_conv<64, 1>(a0, a1, b, kp, sms, st)
                             : d->up2x ? launch_conv<64, 1, 2>(a0, a1, b, kp, sms, st) : launch_conv<64, 1, 3>(a0, a1, b, kp, sms, st);
            default:  return !halo ? launch_conv<32, 1>(a0, a1, b, kp, sms, st)
                             : d->up2x ? launch_conv<32, 1, 2>(a0, a1, b, kp, sms, st) : launch_conv<32, 1, 3>(a0, a1, b, kp, sms, st);
        }
    }
    switch (block_n) {
        case 128: return launch_conv<128, 0>(a0, a1, b, kp, sms, st);
        case 64:  return launch_conv<64, 0>(a0, a1, b, kp, sms, st);
        default:  return launch_conv<32, 0>(a0, a1, b, kp, sms, st);
    }
}

extern "C" int nbp_conv_profile_begin(int max_launches) {
    if (max_launches <= 0) return invalid("nbp_conv_profile_begin: max_launches must be positive");
    while (g_prof.ev.size() < (size_t)2 * max_launches) {
        cudaEvent_t e;
        int rc = check_cuda(cudaEventCreate(&e), "cudaEventCreate");
        if (rc) return rc;
        g_prof.ev.push_back(e);
    }
    g_prof.used = 0; g_prof.flops = 0.0; g_prof.dropped = 0; g_prof.enabled = true;
    return NBP_OK;
}

extern "C" int nbp_conv_profile_end(double* total_ms, double* total_flops, uint64_t* launches, uint64_t* dropped) {
    g_prof.enabled = false;
    double ms = 0.0;
    for (size_t i = 0; i + 1 < g_prof.used; i += 2) {
        int rc = check_cuda(cudaEventSynchronize(g_prof.ev[i + 1]), "cudaEventSynchronize");
        if (rc) return rc;
        float t = 0.f;
        rc = check_cuda(cudaEventElapsedTime(&t, g_prof.ev[i], g_prof.ev[i + 1]), "cudaEventElapsedTime");
        if (rc) return rc;
        ms += t;
    }
    if (total_ms) *total_ms = ms;
    if (total_flops) *total_flops = g_prof.flops;
    if (launches) *launches = g_prof.used / 2;
    if (dropped) *dropped = g_prof.dropped;
    return NBP_OK;
}
