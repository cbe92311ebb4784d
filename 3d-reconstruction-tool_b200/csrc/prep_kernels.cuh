// prep_kernels.cuh -- working-resolution colour + gray on the device.
//
// Reference: `_prepare_images` (src/core/mvs_patchmatch.py:167-191, src/core/dense_stereo.py:156-176):
//   cv.resize(img, (new_w, new_h))  ->  cv.cvtColor(BGR2GRAY)  ->  astype(float32) / 255.
// Both OpenCV steps are fixed-point for uint8 and are reproduced bit for bit (checked against cv2 for the tables in
// tests/test_prepare.py and for the kernel in tests/test_gpu_prepare.py):
//   resize INTER_LINEAR: 11-bit coefficients a = round((1-f)*2048), round(f*2048) per output column / row (tables
//     built on the host exactly like OpenCV builds them), horizontal pass in int32, vertical pass
//     (((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2;
//   BGR2GRAY: (3735 B + 19235 G + 9798 R + 2^14) >> 15.
// One thread per output pixel, HBM-bound (reads 4 source pixels x 3 channels, writes 3 B + 4 B).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200d {

struct PrepArgs {
    const uint8_t *src;          // (H,W,3) BGR
    int H, W, nh, nw;
    const int32_t *xi, *yi;      // left / top source index per output column / row
    const int16_t *xa, *ya;      // (n,2) coefficient pairs
    uint8_t *color;              // out (nh,nw,3) BGR
    float *gray;                 // out (nh,nw) in [0,1]
};

__global__ void __launch_bounds__(256) prepare_image_kernel(const __grid_constant__ PrepArgs a) {
    const int dx = blockIdx.x * 32 + threadIdx.x, dy = blockIdx.y * 8 + threadIdx.y;
    if (dx >= a.nw || dy >= a.nh) return;
    const int x0 = a.xi[dx], x1 = min(x0 + 1, a.W - 1);
    const int sy = a.yi[dy];                                    // unclamped: OpenCV clips the two rows, not the index
    const int y0 = min(max(sy, 0), a.H - 1), y1 = min(max(sy + 1, 0), a.H - 1);
    const int a0 = a.xa[2 * dx], a1 = a.xa[2 * dx + 1], b0 = a.ya[2 * dy], b1 = a.ya[2 * dy + 1];
    const uint8_t *r0 = a.src + (size_t)y0 * a.W * 3, *r1 = a.src + (size_t)y1 * a.W * 3;
    int px[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int h0 = r0[3 * x0 + c] * a0 + r0[3 * x1 + c] * a1;
        const int h1 = r1[3 * x0 + c] * a0 + r1[3 * x1 + c] * a1;
        const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
        px[c] = min(max(v, 0), 255);
    }
    const size_t o = (size_t)dy * a.nw + dx;
    a.color[3 * o] = (uint8_t)px[0];
    a.color[3 * o + 1] = (uint8_t)px[1];
    a.color[3 * o + 2] = (uint8_t)px[2];
    const int g = (px[0] * 3735 + px[1] * 19235 + px[2] * 9798 + (1 << 14)) >> 15;
    a.gray[o] = __fdiv_rn((float)g, 255.0f);
}

}  // namespace b200d
