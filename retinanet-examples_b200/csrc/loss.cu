// loss.cu -- fused focal loss forward + backward for sm_100a (HBM-bound, one pass).
//
// Replaces FocalLoss.forward (reference odtk/loss.py:13-18) as used by Model._compute_loss
// (odtk/model.py:195-199): sigmoid, BCE-with-logits, alpha_t, p_t, the modulating factor, the
// (depth >= 0) mask, the sum, and -- in the same pass -- the gradient of that sum w.r.t. the
// logits, which the reference obtains by autograd through ~8 elementwise kernels, each a full
// pass over [B, A, C, H, W] fp32 (15.36 M elements per image at 3x800x1280).
//
// Targets come either dense (fp32 one-hot, the reference's layout) or as class indices
// [B*A, H*W] (class id, -1 = background, -2 = ignored), which removes the one-hot read
// altogether.  fp32 maths; 128-bit loads/stores; block partial sums in double, summed in a
// fixed order by a second tiny kernel, so the result is deterministic.
#include <cuda_fp16.h>

#include "common.cuh"
#include "prof.cuh"

namespace {

constexpr int kThreads = 256;

struct LossParams {
  const float *logits;
  const float *target;   // dense one-hot or NULL
  const float *mask;     // dense mask or NULL
  const int *cls_index;  // [groups, hw] or NULL
  long long n;
  int num_classes, hw;   // indexed mode: element e -> group e / (C*hw), class (e / hw) % C, pos e % hw
  float alpha, gamma, grad_scale;
  float *loss_elem;      // or NULL
  float *grad;           // or NULL
  double *partials;      // [gridDim.x]
};

__device__ __forceinline__ void focal_one(float x, float t, float m, const LossParams &p, float &loss, float &grad) {
  const float e = expf(-fabsf(x));
  const float inv = 1.0f / (1.0f + e);
  const float pr = x >= 0.0f ? inv : e * inv;          // sigmoid(x)
  const float ce = fmaxf(x, 0.0f) - x * t + log1pf(e);   // BCE with logits
  const float a = t * p.alpha + (1.0f - t) * (1.0f - p.alpha);
  const bool pos = (t == 1.0f);
  const float q = pos ? 1.0f - pr : pr;                  // 1 - p_t
  float w, dwq;                                          // q^gamma, gamma * q^(gamma-1)
  if (p.gamma == 2.0f) { w = q * q; dwq = 2.0f * q; }
  else { w = powf(q, p.gamma); dwq = (p.gamma == 0.0f) ? 0.0f : p.gamma * powf(q, p.gamma - 1.0f); }
  loss = m * a * w * ce;
  const float dq = (pos ? -1.0f : 1.0f) * pr * (1.0f - pr);
  grad = m * a * (dwq * dq * ce + w * (pr - t)) * p.grad_scale;
}

__global__ void __launch_bounds__(kThreads) focal_loss_kernel(LossParams p) {
  __shared__ double s_part[kThreads / 32];
  double acc = 0.0;
  const long long nvec = p.n >> 2;
  const bool vec_ok = ((p.n & 3) == 0) && (p.cls_index == nullptr || (p.hw & 3) == 0);
  if (vec_ok) {
    for (long long v = (long long)blockIdx.x * kThreads + threadIdx.x; v < nvec; v += (long long)gridDim.x * kThreads) {
      const long long e0 = v << 2;
      float4 x4 = odtk_ld_stream_f4(reinterpret_cast<const float4 *>(p.logits) + v);
      float x[4] = {x4.x, x4.y, x4.z, x4.w}, t[4], m[4] = {1.f, 1.f, 1.f, 1.f};
      if (p.cls_index) {
        const long long chw = (long long)p.num_classes * p.hw;
        const long long g = e0 / chw;
        const long long r = e0 - g * chw;
        const int c = (int)(r / p.hw), pos = (int)(r - (long long)c * p.hw);
        int4 ci = *reinterpret_cast<const int4 *>(p.cls_index + g * p.hw + pos);
        int idx[4] = {ci.x, ci.y, ci.z, ci.w};
#pragma unroll
        for (int j = 0; j < 4; j++) { t[j] = (idx[j] == c) ? 1.0f : 0.0f; m[j] = (idx[j] == -2) ? 0.0f : 1.0f; }
      } else {
        float4 t4 = odtk_ld_stream_f4(reinterpret_cast<const float4 *>(p.target) + v);
        t[0] = t4.x; t[1] = t4.y; t[2] = t4.z; t[3] = t4.w;
        if (p.mask) {
          float4 m4 = odtk_ld_stream_f4(reinterpret_cast<const float4 *>(p.mask) + v);
          m[0] = m4.x; m[1] = m4.y; m[2] = m4.z; m[3] = m4.w;
        }
      }
      float l[4], g4[4];
#pragma unroll
      for (int j = 0; j < 4; j++) { focal_one(x[j], t[j], m[j], p, l[j], g4[j]); acc += (double)l[j]; }
      if (p.loss_elem) reinterpret_cast<float4 *>(p.loss_elem)[v] = make_float4(l[0], l[1], l[2], l[3]);
      if (p.grad) reinterpret_cast<float4 *>(p.grad)[v] = make_float4(g4[0], g4[1], g4[2], g4[3]);
    }
  } else {
    for (long long e = (long long)blockIdx.x * kThreads + threadIdx.x; e < p.n; e += (long long)gridDim.x * kThreads) {
      float x = p.logits[e], t, m = 1.0f;
      if (p.cls_index) {
        const long long chw = (long long)p.num_classes * p.hw;
        const long long g = e / chw;
        const long long r = e - g * chw;
        const int c = (int)(r / p.hw), pos = (int)(r - (long long)c * p.hw);
        const int idx = p.cls_index[g * p.hw + pos];
        t = (idx == c) ? 1.0f : 0.0f;
        m = (idx == -2) ? 0.0f : 1.0f;
      } else {
        t = p.target[e];
        if (p.mask) m = p.mask[e];
      }
      float l, g;
      focal_one(x, t, m, p, l, g);
      acc += (double)l;
      if (p.loss_elem) p.loss_elem[e] = l;
      if (p.grad) p.grad[e] = g;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < kThreads / 32; w++) s += s_part[w];
    p.partials[blockIdx.x] = s;
  }
}

// Smooth L1 (reference odtk/loss.py:27-31): x = |pred - target|; x >= beta ? x - beta/2 : x^2 / (2 beta);
// masked sum + gradient in the same pass (box regression loss of Model._compute_loss, model.py:201-205).
__global__ void __launch_bounds__(kThreads) smooth_l1_kernel(const float *pred, const float *target, const float *mask,
                                                             long long n, float beta, float grad_scale,
                                                             float *loss_elem, float *grad, double *partials) {
  __shared__ double s_part[kThreads / 32];
  double acc = 0.0;
  for (long long e = (long long)blockIdx.x * kThreads + threadIdx.x; e < n; e += (long long)gridDim.x * kThreads) {
    const float d = pred[e] - target[e], x = fabsf(d), m = mask ? mask[e] : 1.0f;
    const bool lin = x >= beta;
    const float l = m * (lin ? x - 0.5f * beta : 0.5f * x * x / beta);
    acc += (double)l;
    if (loss_elem) loss_elem[e] = l;
    if (grad) grad[e] = m * grad_scale * (lin ? (d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f)) : d / beta);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < kThreads / 32; w++) s += s_part[w];
    partials[blockIdx.x] = s;
  }
}

__global__ void focal_loss_finish_kernel(const double *partials, int n, float *out) {
  double s = 0.0;
  for (int i = 0; i < n; i++) s += partials[i];  // fixed order: deterministic
  *out = (float)s;
}

int loss_grid(long long n) {
  long long b = (n / 4 + kThreads - 1) / kThreads / 4;
  if (b < 1) b = 1;
  if (b > odtk_sm_count() * 8) b = odtk_sm_count() * 8;
  return (int)b;
}

}  // namespace

extern "C" long long odtk_focal_loss(const float *logits, const float *target, const float *mask,
                                     const int *cls_index, long long n, int num_classes, int hw, float alpha,
                                     float gamma, float grad_scale, float *loss_elem, float *loss_sum, float *grad,
                                     void *workspace, size_t workspace_size, odtk_stream_t stream_) {
  if (n <= 0) return ODTK_E_INVALID;
  const int grid = loss_grid(n);
  const size_t need = odtk_align_up((size_t)grid * sizeof(double));
  if (!workspace || !workspace_size) return (long long)need;
  if (workspace_size < need) return ODTK_E_WORKSPACE;
  if (!logits || !loss_sum || (!target && !cls_index)) return ODTK_E_INVALID;
  if (cls_index && (num_classes <= 0 || hw <= 0 || n % ((long long)num_classes * hw))) return ODTK_E_INVALID;
  if (((uintptr_t)logits | (uintptr_t)target | (uintptr_t)mask | (uintptr_t)cls_index | (uintptr_t)loss_elem |
       (uintptr_t)grad) & 15)
    return ODTK_E_INVALID;
  cudaStream_t stream = (cudaStream_t)stream_;
  LossParams p;
  p.logits = logits; p.target = target; p.mask = mask; p.cls_index = cls_index;
  p.n = n; p.num_classes = num_classes; p.hw = hw;
  p.alpha = alpha; p.gamma = gamma; p.grad_scale = grad_scale;
  p.loss_elem = loss_elem; p.grad = grad; p.partials = (double *)workspace;
  {
    OdtkProfScope prof(ODTK_PROF_LOSS, stream);
    focal_loss_kernel<<<grid, kThreads, 0, stream>>>(p);
  }
  focal_loss_finish_kernel<<<1, 1, 0, stream>>>(p.partials, grid, loss_sum);
  return cudaGetLastError() == cudaSuccess ? ODTK_OK : ODTK_E_CUDA;
}

extern "C" long long odtk_smooth_l1_loss(const float *pred, const float *target, const float *mask, long long n,
                                         float beta, float grad_scale, float *loss_elem, float *loss_sum, float *grad,
                                         void *workspace, size_t workspace_size, odtk_stream_t stream_) {
  if (n <= 0 || !(beta > 0.0f)) return ODTK_E_INVALID;
  const int grid = loss_grid(n);
  const size_t need = odtk_align_up((size_t)grid * sizeof(double));
  if (!workspace || !workspace_size) return (long long)need;
  if (workspace_size < need) return ODTK_E_WORKSPACE;
  if (!pred || !target || !loss_sum) return ODTK_E_INVALID;
  cudaStream_t stream = (cudaStream_t)stream_;
  {
    OdtkProfScope prof(ODTK_PROF_LOSS, stream);
    smooth_l1_kernel<<<grid, kThreads, 0, stream>>>(pred, target, mask, n, beta, grad_scale, loss_elem, grad,
                                                    (double *)workspace);
  }
  focal_loss_finish_kernel<<<1, 1, 0, stream>>>((const double *)workspace, grid, loss_sum);
  return cudaGetLastError() == cudaSuccess ? ODTK_OK : ODTK_E_CUDA;
}
