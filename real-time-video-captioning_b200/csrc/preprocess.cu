// Frame preprocessing fused into one pass: OpenCV frame (uint8, H x W x 3, BGR) ->
// ToTensor -> bicubic Resize(224) (no antialias, align_corners = False, A = -0.75) -> CenterCrop(224) -> BGR->RGB ->
// CLIP Normalize -> fp32 NCHW, i.e. the reference's image_transform() (src/utils/dataloader.py:18-32).  Only the
// 224 x 224 crop is ever computed; the 16 taps of an output pixel come straight from the uint8 frame (4x fewer
// host->device bytes than shipping normalised fp32 frames).  HBM / L2 bound; one thread per output pixel, 3 channels.
#include "common.cuh"
#include "kernels.h"

namespace {

__device__ __forceinline__ float cubic1(float x, float A) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic2(float x, float A) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }
__device__ __forceinline__ void cubic_coeffs(float t, float (&c)[4]) {
  const float A = -0.75f;
  c[0] = cubic2(t + 1.f, A);
  c[1] = cubic1(t, A);
  c[2] = cubic1(1.f - t, A);
  c[3] = cubic2(2.f - t, A);
}

// TO_PATCHES: write the pixel as bf16 straight into the patch matrix the patch-embedding GEMM reads (the im2col layout of
// elementwise.cu: row = (frame, patch row, patch column), column = c*P*P + ky*P + kx, columns >= 3*P*P zero) instead of
// fp32 NCHW -- same value, same round-to-nearest cast as im2col_kernel, so the two routes agree bit for bit.
template <bool TO_PATCHES>
__global__ void __launch_bounds__(256)
preprocess_kernel(const uint8_t* __restrict__ frames, int n, int H, int W, int nh, int nw, int top, int left, int size,
                  float* __restrict__ out, bf16* __restrict__ patches, int P, int kpad) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = n * size * size;
  if (idx >= total) return;
  const int ox = idx % size, oy = (idx / size) % size, f = idx / (size * size);
  // position in the (virtual) resized image, then source coordinates (align_corners = False)
  const float sy = ((float)(oy + top) + 0.5f) * ((float)H / (float)nh) - 0.5f;
  const float sx = ((float)(ox + left) + 0.5f) * ((float)W / (float)nw) - 0.5f;
  const float fy = floorf(sy), fx = floorf(sx);
  const int iy = (int)fy, ix = (int)fx;
  float cy[4], cx[4];
  cubic_coeffs(sy - fy, cy);
  cubic_coeffs(sx - fx, cx);
  float acc[3] = {0.f, 0.f, 0.f};
  const uint8_t* base = frames + (size_t)f * H * W * 3;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int y = min(max(iy - 1 + j, 0), H - 1);
    float row[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int x = min(max(ix - 1 + i, 0), W - 1);
      const uint8_t* p = base + ((size_t)y * W + x) * 3;
      row[0] = fmaf((float)p[0] * (1.f / 255.f), cx[i], row[0]);
      row[1] = fmaf((float)p[1] * (1.f / 255.f), cx[i], row[1]);
      row[2] = fmaf((float)p[2] * (1.f / 255.f), cx[i], row[2]);
    }
    acc[0] = fmaf(row[0], cy[j], acc[0]);
    acc[1] = fmaf(row[1], cy[j], acc[1]);
    acc[2] = fmaf(row[2], cy[j], acc[2]);
  }
  // BGR -> RGB and CLIP normalisation: output channel c reads input channel 2 - c
  const float mean[3] = {0.48145466f, 0.4578275f, 0.40821073f};
  const float stdv[3] = {0.26862954f, 0.26130258f, 0.27577711f};
  if (TO_PATCHES) {
    const int G = size / P;
    const int gy = oy / P, ky = oy - gy * P, gx = ox / P, kx = ox - gx * P;
    bf16* o = patches + ((size_t)f * G * G + (size_t)gy * G + gx) * kpad;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c * P * P + ky * P + kx] = __float2bfloat16_rn((acc[2 - c] - mean[c]) / stdv[c]);
    if (ky == 0 && kx == 0)
      for (int k = 3 * P * P; k < kpad; ++k) o[k] = __float2bfloat16_rn(0.f);
  } else {
    float* o = out + (size_t)f * 3 * size * size + (size_t)oy * size + ox;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[(size_t)c * size * size] = (acc[2 - c] - mean[c]) / stdv[c];
  }
}

}  // namespace

cudaError_t preprocess_frames_u8(const uint8_t* frames, int n, int H, int W, int size, float* out, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  if (H < 1 || W < 1 || size < 1) return cudaErrorInvalidValue;
  int nh, nw;
  if (H <= W) { nh = size; nw = (int)((double)size * W / H); } else { nh = (int)((double)size * H / W); nw = size; }
  const int top = (int)lrint((nh - size) / 2.0), left = (int)lrint((nw - size) / 2.0);
  const int total = n * size * size;
  preprocess_kernel<false><<<(total + 255) / 256, 256, 0, stream>>>(frames, n, H, W, nh, nw, top, left, size, out, nullptr, 0, 0);
  note_launch();
  return cudaGetLastError();
}

cudaError_t preprocess_frames_u8_to_patches(const uint8_t* frames, int n, int H, int W, int size, int patch, int kpad,
                                            bf16* patches, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  if (H < 1 || W < 1 || size < 1 || patch < 1 || size % patch != 0 || kpad < 3 * patch * patch) return cudaErrorInvalidValue;
  int nh, nw;
  if (H <= W) { nh = size; nw = (int)((double)size * W / H); } else { nh = (int)((double)size * H / W); nw = size; }
  const int top = (int)lrint((nh - size) / 2.0), left = (int)lrint((nw - size) / 2.0);
  const int total = n * size * size;
  preprocess_kernel<true><<<(total + 255) / 256, 256, 0, stream>>>(frames, n, H, W, nh, nw, top, left, size, nullptr, patches, patch, kpad);
  note_launch();
  return cudaGetLastError();
}
