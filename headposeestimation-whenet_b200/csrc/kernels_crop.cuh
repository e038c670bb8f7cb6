// kernels_crop.cuh - crop front-end of the stream path.
//
// For every head box of a frame: slice [y0:y1, x0:x1] out of the BGR frame, swap to RGB and resize to
// 224x224 exactly as cv2.resize's 8-bit INTER_LINEAR kernel does (reference demo_video.py:21-23,
// demo.py:10-11): half-pixel centres, 11-bit weights rounded to nearest-even, int32 horizontal pass,
// ((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2 >> 2 vertical pass, 2x2 box for exact 2x down-scaling.
// All heads of a frame come out as ONE uint8 NHWC batch that feeds the stem kernel directly - the
// reference resizes and runs the network one head at a time (demo_video.py:57-58).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace whenet {

struct AxisTap { int s0, s1, w0, w1; };

// OpenCV resize(): left source index + the two 11-bit weights for destination index d.
// clamp_weights: x axis resets the fraction at the borders, y axis only clips the row indices.
__device__ __forceinline__ AxisTap axis_tap(int d, int src, bool clamp_weights) {
    const double scale = 1.0 / (224.0 / (double)src);
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    AxisTap t;
    if (clamp_weights) {
        if (s < 0) { f = 0.f; s = 0; }
        if (s >= src - 1) { f = 0.f; s = src - 1; }
        t.s0 = s;
        t.s1 = min(s + 1, src - 1);
    } else {
        t.s0 = min(max(s, 0), src - 1);
        t.s1 = min(max(s + 1, 0), src - 1);
    }
    t.w1 = __float2int_rn(f * 2048.f);              // saturate_cast<short>(float): round half to even
    t.w0 = __float2int_rn((1.f - f) * 2048.f);
    return t;
}

// grid = (ceil(224*224/256), M).  rects[m] = (y0, y1, x0, x1) slice bounds inside the frame.
__global__ void __launch_bounds__(256) crop_resize_kernel(const uint8_t* __restrict__ frame, int H, int W,
                                                          const int4* __restrict__ rects, uint8_t* __restrict__ out,
                                                          int swap_rb) {
    const int m = blockIdx.y;
    const int pix = blockIdx.x * 256 + threadIdx.x;
    if (pix >= 224 * 224) return;
    const int dy = pix / 224, dx = pix - dy * 224;
    const int4 r = rects[m];
    const int y0 = r.x, h = r.y - r.x, x0 = r.z, w = r.w - r.z;
    const uint8_t* src = frame + ((long long)y0 * W + x0) * 3;
    const long long pitch = (long long)W * 3;
    uint8_t* dst = out + ((long long)m * 224 * 224 + pix) * 3;
    int v[3];
    if (h == 448 && w == 448) {                     // resize(): INTER_LINEAR with exact 2x down-scale -> 2x2 box
        const uint8_t* p = src + (long long)(2 * dy) * pitch + (2 * dx) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = (p[c] + p[3 + c] + p[pitch + c] + p[pitch + 3 + c] + 2) >> 2;
    } else {
        const AxisTap tx = axis_tap(dx, w, true), ty = axis_tap(dy, h, false);
        const uint8_t* r0 = src + (long long)ty.s0 * pitch;
        const uint8_t* r1 = src + (long long)ty.s1 * pitch;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int h0 = r0[tx.s0 * 3 + c] * tx.w0 + r0[tx.s1 * 3 + c] * tx.w1;
            const int h1 = r1[tx.s0 * 3 + c] * tx.w0 + r1[tx.s1 * 3 + c] * tx.w1;
            v[c] = (((ty.w0 * (h0 >> 4)) >> 16) + ((ty.w1 * (h1 >> 4)) >> 16) + 2) >> 2;
        }
    }
    if (swap_rb) { dst[0] = (uint8_t)v[2]; dst[1] = (uint8_t)v[1]; dst[2] = (uint8_t)v[0]; }
    else { dst[0] = (uint8_t)v[0]; dst[1] = (uint8_t)v[1]; dst[2] = (uint8_t)v[2]; }
}

}  // namespace whenet
