// layers.cu -- the small memory-bound layers around the tensor-core convolution:
//   * lowering of strided / 7x7 convolutions to GEMM rows (gather of the receptive field into
//     a [pixels, taps*C] fp16 matrix; used for the ResNet stem, the three stride-2 3x3 and 1x1
//     convolutions of the backbone, and FPN pyramid6 / pyramid7 -- together < 5 % of the FLOPs);
//   * 3x3 stride-2 max-pool (torchvision resnet.py maxpool), NHWC fp16.
// Reference call sites: odtk/backbones/resnet.py:25-28, odtk/backbones/fpn.py:54-55.
#include <cuda_fp16.h>

#include "common.cuh"
#include "prof.cuh"

namespace {

// C % 8 == 0: one thread moves 8 channels (16 B) of one tap of one output pixel.
__global__ void lower_vec8_kernel(const __half *__restrict__ x, __half *__restrict__ out, int N, int H, int W, int C,
                                  int OH, int OW, int ks, int stride, int pad, int kpad, int relu) {
  const int c8n = C >> 3;
  const long long total = (long long)N * OH * OW * ks * ks * c8n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c8 = (int)(i % c8n);
    long long r = i / c8n;
    int tap = (int)(r % (ks * ks));
    long long m = r / (ks * ks);
    int ow = (int)(m % OW);
    long long t = m / OW;
    int oh = (int)(t % OH);
    int n = (int)(t / OH);
    int ih = oh * stride + tap / ks - pad, iw = ow * stride + tap % ks - pad;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
      v = __ldg(reinterpret_cast<const uint4 *>(x + (((long long)n * H + ih) * W + iw) * C + c8 * 8));
      if (relu) {
        __half2 *h = reinterpret_cast<__half2 *>(&v);
        const __half2 z = __float2half2_rn(0.0f);
#pragma unroll
        for (int j = 0; j < 4; j++) h[j] = __hmax2(h[j], z);
      }
    }
    *reinterpret_cast<uint4 *>(out + m * kpad + (long long)tap * C + c8 * 8) = v;
  }
}

// generic small-C path (the RGB stem): one thread produces 8 consecutive K entries.
__global__ void lower_generic_kernel(const __half *__restrict__ x, __half *__restrict__ out, int N, int H, int W, int C,
                                     int OH, int OW, int ks, int stride, int pad, int kpad) {
  const int k8n = kpad >> 3;
  const int kreal = ks * ks * C;
  const long long total = (long long)N * OH * OW * k8n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int k8 = (int)(i % k8n);
    long long m = i / k8n;
    int ow = (int)(m % OW);
    long long t = m / OW;
    int oh = (int)(t % OH);
    int n = (int)(t / OH);
    __align__(16) __half v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      int k = k8 * 8 + j;
      __half val = __float2half_rn(0.0f);
      if (k < kreal) {
        int tap = k / C, c = k - tap * C;
        int ih = oh * stride + tap / ks - pad, iw = ow * stride + tap % ks - pad;
        if (ih >= 0 && ih < H && iw >= 0 && iw < W) val = x[(((long long)n * H + ih) * W + iw) * C + c];
      }
      v[j] = val;
    }
    *reinterpret_cast<uint4 *>(out + m * kpad + k8 * 8) = *reinterpret_cast<uint4 *>(v);
  }
}

__global__ void maxpool3x3s2_kernel(const __half *__restrict__ x, __half *__restrict__ y, int N, int H, int W, int C,
                                    int OH, int OW) {
  const int c8n = C >> 3;
  const long long total = (long long)N * OH * OW * c8n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c8 = (int)(i % c8n);
    long long m = i / c8n;
    int ow = (int)(m % OW);
    long long t = m / OW;
    int oh = (int)(t % OH);
    int n = (int)(t / OH);
    __half2 best[4];
    const __half2 ninf = __float2half2_rn(-65504.0f);
#pragma unroll
    for (int j = 0; j < 4; j++) best[j] = ninf;
#pragma unroll
    for (int dy = 0; dy < 3; dy++) {
      int ih = oh * 2 + dy - 1;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; dx++) {
        int iw = ow * 2 + dx - 1;
        if (iw < 0 || iw >= W) continue;
        uint4 v = __ldg(reinterpret_cast<const uint4 *>(x + (((long long)n * H + ih) * W + iw) * C + c8 * 8));
        const __half2 *h = reinterpret_cast<const __half2 *>(&v);
#pragma unroll
        for (int j = 0; j < 4; j++) best[j] = __hmax2(best[j], h[j]);
      }
    }
    *reinterpret_cast<uint4 *>(y + m * C + c8 * 8) = *reinterpret_cast<uint4 *>(best);
  }
}

// RGB NHWC3 fp16 -> zero-padded NHWC4 [n, h+6, w+8, 4]: 3 rows/columns of zeros before, 3/5 after
// (the stem's overlapping-window tensor map never leaves the buffer).  One 8-byte pixel per thread.
__global__ void pad_input_kernel(const __half *__restrict__ x, __half *__restrict__ y, int N, int H, int W) {
  const int HP = H + 6, WP = W + 8;
  const long long total = (long long)N * HP * WP;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int xp = (int)(i % WP);
    long long t = i / WP;
    int yp = (int)(t % HP);
    int n = (int)(t / HP);
    int ih = yp - 3, iw = xp - 3;
    uint2 v = make_uint2(0u, 0u);
    if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
      const unsigned short *s = reinterpret_cast<const unsigned short *>(x) + (((long long)n * H + ih) * W + iw) * 3;
      v.x = (unsigned)s[0] | ((unsigned)s[1] << 16);
      v.y = (unsigned)s[2];
    }
    reinterpret_cast<uint2 *>(y)[i] = v;
  }
}

// Same, for W % 8 == 0 (input rows start 16-byte aligned): one block per padded row; the 6-byte pixels of the input row
// are staged through shared memory with coalesced 16-byte loads, the padded row leaves as coalesced 16-byte stores
// (two 8-byte pixels each) -- the per-pixel version above issues three 2-byte loads per pixel.
__global__ void pad_input_rows_kernel(const __half *__restrict__ x, __half *__restrict__ y, int N, int H, int W) {
  extern __shared__ uint4 srow[];                                 // W * 6 bytes of the input row
  const int HP = H + 6, WP = W + 8;
  const int n = blockIdx.x / HP, yp = blockIdx.x - n * HP, ih = yp - 3;
  uint4 *dst = reinterpret_cast<uint4 *>(y + ((long long)n * HP + yp) * WP * 4);
  const int pairs = WP / 2;
  if (ih < 0 || ih >= H) {
    for (int i = threadIdx.x; i < pairs; i += blockDim.x) dst[i] = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  const uint4 *src = reinterpret_cast<const uint4 *>(x + ((long long)n * H + ih) * W * 3);
  const int nv = W * 6 / 16;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) srow[i] = __ldg(src + i);
  __syncthreads();
  const unsigned short *s = reinterpret_cast<const unsigned short *>(srow);
  for (int i = threadIdx.x; i < pairs; i += blockDim.x) {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    const int iw0 = 2 * i - 3, iw1 = iw0 + 1;
    if (iw0 >= 0 && iw0 < W) { v.x = (unsigned)s[iw0 * 3] | ((unsigned)s[iw0 * 3 + 1] << 16); v.y = (unsigned)s[iw0 * 3 + 2]; }
    if (iw1 >= 0 && iw1 < W) { v.z = (unsigned)s[iw1 * 3] | ((unsigned)s[iw1 * 3 + 1] << 16); v.w = (unsigned)s[iw1 * 3 + 2]; }
    dst[i] = v;
  }
}

// Input side of odtk infer (reference odtk/data.py:113-123): uint8 HWC image -> /255 -> (x - mean) / std
// -> zero padding to a multiple of the model stride, fused with the stem's own zero padding and the
// RGB -> NHWC4 widening: one 8-byte store per padded pixel, the fp32 full-image pass disappears.
__global__ void preprocess_u8_kernel(const unsigned char *__restrict__ x, __half *__restrict__ y, int N, int H, int W,
                                     int HS, int WS, float m0, float m1, float m2, float s0, float s1, float s2) {
  const int HP = HS + 6, WP = WS + 8;
  const long long total = (long long)N * HP * WP;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int xp = (int)(i % WP);
    long long t = i / WP;
    int yp = (int)(t % HP);
    int n = (int)(t / HP);
    int ih = yp - 3, iw = xp - 3;
    uint2 v = make_uint2(0u, 0u);
    if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
      const unsigned char *s = x + (((long long)n * H + ih) * W + iw) * 3;
      // same operation order as the reference: float(u8) / 255, then - mean, then / std
      float r = ((float)s[0] / 255.0f - m0) / s0, g = ((float)s[1] / 255.0f - m1) / s1, b = ((float)s[2] / 255.0f - m2) / s2;
      __half2 rg = __floats2half2_rn(r, g);
      __half2 bz = __floats2half2_rn(b, 0.0f);
      v.x = *reinterpret_cast<unsigned *>(&rg);
      v.y = *reinterpret_cast<unsigned *>(&bz);
    }
    reinterpret_cast<uint2 *>(y)[i] = v;
  }
}

// y = max(x, 0) on fp16, 8 elements per thread (FPN pyramid7 reads ReLU(P6), fpn.py:55, while P6 itself is an output)
__global__ void relu_f16_kernel(const uint4 *__restrict__ x, uint4 *__restrict__ y, long long n8) {
  const __half2 z = __float2half2_rn(0.0f);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    uint4 v = __ldg(x + i);
    __half2 *h = reinterpret_cast<__half2 *>(&v);
#pragma unroll
    for (int j = 0; j < 4; j++) h[j] = __hmax2(h[j], z);
    y[i] = v;
  }
}

// 3-D strided copy (+ optional ReLU) of fp16 rows, 16 bytes per thread: moves a pyramid level between a dense
// [n, rows, row_elems] tensor and its rectangle inside the atlas (pitches in elements).
__global__ void copy_rows_f16_kernel(const __half *__restrict__ x, __half *__restrict__ y, int n, int rows, int row8,
                                     long long x_img, long long x_row, long long y_img, long long y_row, int relu) {
  const long long total = (long long)n * rows * row8;
  const __half2 z = __float2half2_rn(0.0f);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % row8);
    const long long t = i / row8;
    const int r = (int)(t % rows), im = (int)(t / rows);
    uint4 v = __ldg(reinterpret_cast<const uint4 *>(x + im * x_img + r * x_row) + c);
    if (relu) {
      __half2 *h = reinterpret_cast<__half2 *>(&v);
#pragma unroll
      for (int j = 0; j < 4; j++) h[j] = __hmax2(h[j], z);
    }
    reinterpret_cast<uint4 *>(y + im * y_img + r * y_row)[c] = v;
  }
}

inline int grid_for(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  long long cap = (long long)odtk_sm_count() * 16;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace

extern "C" int odtk_lower_conv(const void *x, void *out, int n, int h, int w, int c, int ksize, int stride, int pad,
                               int kpad, int relu, odtk_stream_t stream_) {
  if (!x || !out || n <= 0 || h <= 0 || w <= 0 || c <= 0 || ksize <= 0 || stride <= 0 || pad < 0) return ODTK_E_INVALID;
  if (kpad < ksize * ksize * c || (kpad % 8)) return ODTK_E_INVALID;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int oh = (h + 2 * pad - ksize) / stride + 1, ow = (w + 2 * pad - ksize) / stride + 1;
  if (c % 8 == 0 && kpad != ksize * ksize * c) return ODTK_E_INVALID;
  if (c % 8 != 0 && relu) return ODTK_E_UNSUPPORTED;
  OdtkProfScope prof(ODTK_PROF_LAYER, stream);
  if (c % 8 == 0) {
    long long total = (long long)n * oh * ow * ksize * ksize * (c / 8);
    lower_vec8_kernel<<<grid_for(total, 256), 256, 0, stream>>>((const __half *)x, (__half *)out, n, h, w, c, oh, ow,
                                                                ksize, stride, pad, kpad, relu);
  } else {
    long long total = (long long)n * oh * ow * (kpad / 8);
    lower_generic_kernel<<<grid_for(total, 256), 256, 0, stream>>>((const __half *)x, (__half *)out, n, h, w, c, oh,
                                                                   ow, ksize, stride, pad, kpad);
  }
  return cudaGetLastError() == cudaSuccess ? ODTK_OK : ODTK_E_CUDA;
}

extern "C" int odtk_maxpool3x3s2(const void *x, void *y, int n, int h, int w, int c, odtk_stream_t stream_) {
  if (!x || !y || n <= 0 || h <= 0 || w <= 0 || c <= 0 || (c % 8)) return ODTK_E_INVALID;
  const int oh = (h + 2 - 3) / 2 + 1, ow = (w + 2 - 3) / 2 + 1;
  long long total = (long long)n * oh * ow * (c / 8);
  OdtkProfScope prof(ODTK_PROF_LAYER, (cudaStream_t)stream_);
  maxpool3x3s2_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream_>>>((const __half *)x, (__half *)y, n, h, w,
                                                                               c, oh, ow);
  return cudaGetLastError() == cudaSuccess ? ODTK_OK : ODTK_E_CUDA;
}

extern "C" int odtk_pad_input(const void *x, void *y, int n, int h, int w, odtk_stream_t stream_) {
  if (!x || !y || n <= 0 || h <= 0 || w <= 0) return ODTK_E_INVALID;
  long long total = (long long)n * (h + 6) * (w + 8);
  OdtkProfScope prof(ODTK_PROF_LAYER, (cudaStream_t)stream_);
  if (w % 8 == 0 && (((uintptr_t)x | (uintptr_t)y) & 15) == 0 && (size_t)w * 6 <= 48 * 1024 && (long long)n * (h + 6) < (1ll << 31))
    pad_input_rows_kernel<<<n * (h + 6), 256, (size_t)w * 6, (cudaStream_t)stream_>>>((const __half *)x, (__half *)y, n, h, w);
  else
    pad_input_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream_>>>((const __half *)x, (__half *)y, n, h, w);
  return cudaGetLastError() == cudaSuccess ? ODTK_OK : ODTK_E_CUDA;
}

extern "C" int odtk_preprocess_u8(const void *x, void *y, int n, int h, int w, int hs, int ws, const float *mean,
                                  const float *std, odtk_stream_t stream_) {
  if (!x || !y || !mean || !std || n <= 0 || h <= 0 || w <= 0 || hs < h || ws < w) return ODTK_E_INVALID;
  long long total = (long long)n * (hs + 6) * (ws + 8);
  OdtkProfScope prof(ODTK_PROF_LAYER, (cudaStream_t)stream_);
  preprocess_u8_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream_>>>(
      (const unsigned char *)x, (__half *)y, n, h, w, hs, ws, mean[0], mean[1], mean[2], std[0], std[1], std[2]);
  return cudaGetLastError() == cudaSuccess ? ODTK_OK : ODTK_E_CUDA;
}

// Depthwise 3x3 convolution (MobileNetV2's inverted residual blocks, torchvision mobilenetv2.py InvertedResidual behind
// odtk/backbones/mobilenet.py:5-25), pad 1, stride 1 / 2, + bias (folded BatchNorm) + ReLU6.  NHWC fp16, 8 channels per
// thread (16-byte loads), fp32 accumulation in tap order; w: [9][c] fp16.  No contraction over channels: HBM / L2-bound.
__global__ void depthwise3x3_kernel(const __half *__restrict__ x, const __half *__restrict__ w, const float *__restrict__ bias,
                                    __half *__restrict__ y, int n, int h, int wd, int c, int oh, int ow, int stride, int act) {
  const int c8 = c >> 3;
  const long long total = (long long)n * oh * ow * c8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cc = (int)(i % c8);
    long long pix = i / c8;
    const int ox = (int)(pix % ow);
    pix /= ow;
    const int oy = (int)(pix % oh), img = (int)(pix / oh);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = 0.0f;
#pragma unroll
    for (int r = 0; r < 3; r++) {
      const int iy = oy * stride + r - 1;
      if (iy < 0 || iy >= h) continue;
#pragma unroll
      for (int s = 0; s < 3; s++) {
        const int ix = ox * stride + s - 1;
        if (ix < 0 || ix >= wd) continue;
        const uint4 xv = __ldg(reinterpret_cast<const uint4 *>(x + (((long long)img * h + iy) * wd + ix) * c + cc * 8));
        const uint4 wv = __ldg(reinterpret_cast<const uint4 *>(w + (long long)(r * 3 + s) * c + cc * 8));
        const __half2 *xh = reinterpret_cast<const __half2 *>(&xv), *wh = reinterpret_cast<const __half2 *>(&wv);
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const float2 a = __half22float2(xh[j]), b = __half22float2(wh[j]);
          acc[2 * j] = fmaf(a.x, b.x, acc[2 * j]);
          acc[2 * j + 1] = fmaf(a.y, b.y, acc[2 * j + 1]);
        }
      }
    }
    uint4 o;
    __half2 *oh2 = reinterpret_cast<__half2 *>(&o);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float a = acc[2 * j] + (bias ? __ldg(bias + cc * 8 + 2 * j) : 0.0f), b = acc[2 * j + 1] + (bias ? __ldg(bias + cc * 8 + 2 * j + 1) : 0.0f);
      if (act) { a = fmaxf(a, 0.0f); b = fmaxf(b, 0.0f); }
      if (act == 2) { a = fminf(a, 6.0f); b = fminf(b, 6.0f); }
      oh2[j] = __floats2half2_rn(a, b);
    }
    *reinterpret_cast<uint4 *>(y + (((long long)img * oh + oy) * ow + ox) * c + cc * 8) = o;
  }
}

extern "C" int odtk_depthwise3x3(const void *x, const void *w, const float *bias, void *y, int n, int h, int width, int c,
                                 int stride, int act, odtk_stream_t stream_) {
  if (!x || !w || !y || n <= 0 || h <= 0 || width <= 0 || c <= 0 || (c % 8) || (stride != 1 && stride != 2) || act < 0 || act > 2)
    return ODTK_E_INVALID;
  if (((uintptr_t)x | (uintptr_t)w | (uintptr_t)y) & 15) return ODTK_E_INVALID;
  const int oh = (h - 1) / stride + 1, ow = (width - 1) / stride + 1;
  const long long total = (long long)n * oh * ow * (c / 8);
  OdtkProfScope prof(ODTK_PROF_LAYER, (cudaStream_t)stream_);
  depthwise3x3_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream_>>>((const __half *)x, (const __half *)w, bias, (__half *)y, n, h,
                                                                              width, c, oh, ow, stride, act);
  return cudaGetLastError() == cudaSuccess ? ODTK_OK : ODTK_E_CUDA;
}

extern "C" int odtk_relu_f16(const void *x, void *y, long long n, odtk_stream_t stream_) {
  if (!x || !y || n <= 0 || (n % 8) || (((uintptr_t)x | (uintptr_t)y) & 15)) return ODTK_E_INVALID;
  OdtkProfScope prof(ODTK_PROF_LAYER, (cudaStream_t)stream_);
  relu_f16_kernel<<<grid_for(n / 8, 256), 256, 0, (cudaStream_t)stream_>>>((const uint4 *)x, (uint4 *)y, n / 8);
  return cudaGetLastError() == cudaSuccess ? ODTK_OK : ODTK_E_CUDA;
}

extern "C" int odtk_copy_rows_f16(const void *x, void *y, int n, int rows, int row_elems, long long x_img_pitch,
                                  long long x_row_pitch, long long y_img_pitch, long long y_row_pitch, int relu,
                                  odtk_stream_t stream_) {
  if (!x || !y || n <= 0 || rows <= 0 || row_elems <= 0 || (row_elems % 8)) return ODTK_E_INVALID;
  if ((x_img_pitch | x_row_pitch | y_img_pitch | y_row_pitch) % 8 || (((uintptr_t)x | (uintptr_t)y) & 15)) return ODTK_E_INVALID;
  const long long total = (long long)n * rows * (row_elems / 8);
  OdtkProfScope prof(ODTK_PROF_LAYER, (cudaStream_t)stream_);
  copy_rows_f16_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream_>>>((const __half *)x, (__half *)y, n, rows, row_elems / 8,
                                                                               x_img_pitch, x_row_pitch, y_img_pitch, y_row_pitch, relu);
  return cudaGetLastError() == cudaSuccess ? ODTK_OK : ODTK_E_CUDA;
}
