// Shared helpers for the nextbestpath_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nbp_b200.h"

namespace nbp {

// thread-local last-error message (nbp_last_error)
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);      // 0 or positive CUDA error code
int invalid(const char* fmt, ...);                    // always returns NBP_ERR_INVALID
void count_launch(uint64_t n = 1);                    // nbp_launch_count accounting

// One-time per-DEVICE state (function attributes, SM counts): a process may drive several GPUs, and cudaFuncSetAttribute /
// occupancy results belong to the device that was current when they were made.
static constexpr int NBP_MAX_DEVICES = 64;
inline int device_slot() {
    int d = 0;
    cudaGetDevice(&d);
    return (d >= 0 && d < NBP_MAX_DEVICES) ? d : 0;
}

// Pinned fp32 arithmetic: one IEEE rounding per operation, never contracted into FMA, so that
// results are bit-identical to the strict-fp32 CPU oracle (oracle/raster_oracle.c, oracle/oracle.py).
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// e4m3 pair planes (nbp_conv_desc mode 2) hold e4m3(x * E4M3_ACT_SCALE) and e4m3((x - hi) * 2048 * E4M3_ACT_SCALE): the power-of-two
// pre-scale moves the format's window [2^-9, 448] to [2^-6, 3584] -- BatchNorm-ed activations are O(1), so nothing is lost at the bottom
// (measured: same network error for 1, 1/8, 1/32) and inputs 8x beyond the calibrated range still do not saturate.
#define NBP_E4M3_ACT_SCALE 0.125f
#define NBP_E4M3_MAX 448.0f

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// ordered in-block compaction helper: returns this thread's slot (valid only if flag) and the
// block total.  All threads of the block must call it.  warp_cnt: smem array of blockDim/32 ints.
__device__ __forceinline__ int block_compact(bool flag, int* warp_cnt, int& total) {
    const unsigned ball = __ballot_sync(0xffffffffu, flag);
    const int lane = lane_id(), warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    if (lane == 0) warp_cnt[warp] = __popc(ball);
    __syncthreads();
    int base = 0, tot = 0;
    for (int w = 0; w < nwarp; ++w) {
        const int c = warp_cnt[w];
        if (w < warp) base += c;
        tot += c;
    }
    total = tot;
    __syncthreads();   // warp_cnt may be reused by the caller's next round
    return base + __popc(ball & ((1u << lane) - 1u));
}

}  // namespace nbp
