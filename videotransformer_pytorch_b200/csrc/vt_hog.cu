// MaskFeat HOG target on the GPU (reference: dataset.py:39-45 -> skimage.feature.hog, 9 orientations,
// 8x8 cells, 1x1 blocks, L2 norm, per colour channel, then the 2x2 cell regroup to (14,14,108)).
// One CTA per output patch (16x16 pixels = 2x2 cells): 18x18x3 u8 halo tile in shared memory,
// integer central differences, orientation bin from a 511x511 host-built LUT (bit-exact vs numpy's
// arctan2/rad2deg/%180), fp32 magnitudes, then one thread per (cell, channel, bin) sums its 64 pixels in
// the reference's row-major order (deterministic), /64, L2-normalises and writes the 108-vector.
#include "vt_common.cuh"

namespace vt {

__global__ void __launch_bounds__(256)
hog_kernel(const uint8_t* __restrict__ frames, const uint8_t* __restrict__ lut, float* __restrict__ feat,
           uint8_t* __restrict__ bins, int H, int W) {
  __shared__ uint8_t tile[18][18][3];
  __shared__ float mag[3][16][16];
  __shared__ uint8_t bin[3][16][16];
  __shared__ float hist[4][3][9];
  const int PW = W / 16, PH = H / 16;
  const int pw = blockIdx.x % PW, ph = (blockIdx.x / PW) % PH, f = blockIdx.x / (PW * PH);
  const uint8_t* img = frames + (long long)f * H * W * 3;
  const int y0 = ph * 16 - 1, x0 = pw * 16 - 1;
  for (int idx = threadIdx.x; idx < 18 * 18; idx += 256) {
    const int ty = idx / 18, tx = idx % 18;
    const int y = y0 + ty, x = x0 + tx;
    uint8_t r = 0, g = 0, b = 0;
    if (y >= 0 && y < H && x >= 0 && x < W) {
      const uint8_t* px = img + ((long long)y * W + x) * 3;
      r = px[0]; g = px[1]; b = px[2];
    }
    tile[ty][tx][0] = r; tile[ty][tx][1] = g; tile[ty][tx][2] = b;
  }
  __syncthreads();
  {
    const int ly = threadIdx.x >> 4, lx = threadIdx.x & 15;
    const int y = ph * 16 + ly, x = pw * 16 + lx;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      int gy = 0, gx = 0;
      if (y > 0 && y < H - 1) gy = (int)tile[ly + 2][lx + 1][c] - (int)tile[ly][lx + 1][c];
      if (x > 0 && x < W - 1) gx = (int)tile[ly + 1][lx + 2][c] - (int)tile[ly + 1][lx][c];
      const uint8_t bi = __ldg(lut + (gy + 255) * 511 + (gx + 255));
      mag[c][ly][lx] = sqrtf((float)(gx * gx + gy * gy));
      bin[c][ly][lx] = bi;
      if (bins) bins[(((long long)f * 3 + c) * H + y) * W + x] = bi;
    }
  }
  __syncthreads();
  if (threadIdx.x < 108) {
    const int k = threadIdx.x % 9, c = (threadIdx.x / 9) % 3, cell = threadIdx.x / 27;  // cell = dh*2+dw
    const int cy = (cell >> 1) * 8, cx = (cell & 1) * 8;
    float s = 0.f;
    for (int i = 0; i < 8; ++i)
      for (int j = 0; j < 8; ++j)
        if (bin[c][cy + i][cx + j] == k) s += mag[c][cy + i][cx + j];
    hist[cell][c][k] = s * (1.0f / 64.0f);
  }
  __syncthreads();
  if (threadIdx.x < 108) {
    const int k = threadIdx.x % 9, c = (threadIdx.x / 9) % 3, cell = threadIdx.x / 27;
    float ss = 0.f;
#pragma unroll
    for (int q = 0; q < 9; ++q) ss += hist[cell][c][q] * hist[cell][c][q];
    const float v = hist[cell][c][k] / sqrtf(ss + 1e-10f);
    feat[(((long long)f * PH + ph) * PW + pw) * 108 + cell * 27 + c * 9 + k] = v;
  }
}

}  // namespace vt

extern "C" int vt_hog(const vt_hog_params* p, void* stream) {
  using namespace vt;
  VT_REQUIRE(p && p->frames && p->lut && p->feat, "vt_hog: null pointer");
  VT_REQUIRE(p->F > 0 && p->H >= 16 && p->W >= 16 && p->H % 16 == 0 && p->W % 16 == 0, "vt_hog: bad geometry F=%d H=%d W=%d",
             p->F, p->H, p->W);
  const int blocks = p->F * (p->H / 16) * (p->W / 16);
  hog_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(p->frames, p->lut, p->feat, p->bins, p->H, p->W);
  return check_launch("hog_kernel");
}
