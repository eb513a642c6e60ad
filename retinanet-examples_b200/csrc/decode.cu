// decode.cu -- score filter + exact top-n + anchor box decode for sm_100a.
//
// Replaces odtk::cuda::decode / decode_rotate (reference csrc/cuda/decode.cu:44-171,
// decode_rotate.cu:42-179), which run, PER IMAGE and with a host sync in between,
// thrust::transform -> cub::DeviceSelect -> D2H count -> gather -> cub radix sort ->
// thrust::transform.  Here one level of the whole batch is two launches and no host sync:
//
//   K1 score_filter_kernel   HBM-bound streaming pass over [B, A*C*H*W] fp32 (128-bit
//                            L1-bypassing loads, 4 in flight per lane).  Survivors
//                            (score > thresh, ~0.5 %) are staged per warp in shared
//                            memory and flushed with ONE global atomic per >= 32 of them
//                            into a per-image candidate list (key, flat index), while a
//                            2048-bin histogram of the score keys is accumulated.
//   K2 select_decode_kernel  one CTA per image: the histogram gives the bin that holds
//                            the top_n-th score; only candidates at or above it (about
//                            top_n of them) are pulled into shared memory and ordered
//                            with a bitonic network on the unique composite key
//                            (score key, ~flat index) == the reference's stable order;
//                            then every kept index is decoded (fp32, IEEE, no FMA
//                            contraction, same operation order as decode.cu:133-156).
//
// Exactness never depends on the data: if the candidate list overflows, or a histogram
// bin holds more ties than the sort capacity, K2 falls back to an 8-pass radix select on
// the composite key (slow, exact).  Compile with -fmad=false (see Makefile).
#include "common.cuh"
#include "prof.cuh"

namespace {

constexpr int kHistBins = 2048;
constexpr int kSortCap = ODTK_MAX_TOP_N;  // 4096 keys of 8 B = 32 KB shared
constexpr int kMaxAnchors = 32;           // anchor table travels as a kernel parameter
constexpr int kFilterThreads = 256;
constexpr int kWarpsPerBlock = kFilterThreads / 32;
constexpr int kTile = 512;                 // elements per warp per iteration (4 x float4 x 32)
constexpr int kStage = 32 + kTile;         // per-warp staging entries

struct AnchorTable {
  float v[4 * kMaxAnchors];
};

struct FilterParams {
  const float *scores;   // [B, n]
  long long n;           // elements per image
  float thresh;
  uint32_t key_thresh;
  int shift;             // histogram bin = min((key - key_thresh) >> shift, kHistBins-1)
  int *counts;           // [B]
  uint32_t *hist;        // [B, kHistBins]
  uint2 *cand;           // [B, cap]  (x = score key, y = flat index)
  long long cap;
};

__device__ __forceinline__ int hist_bin(uint32_t key, uint32_t key_thresh, int shift) {
  uint32_t d = (key - key_thresh) >> shift;
  return d < (uint32_t)(kHistBins - 1) ? (int)d : (kHistBins - 1);
}

// ------------------------------------------------------------------------------------
// K1: streaming filter.  grid = (blocks, B), block = 256.
template <bool VEC>
__global__ void __launch_bounds__(kFilterThreads) score_filter_kernel(FilterParams p) {
  __shared__ uint2 stage[kWarpsPerBlock][kStage];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int img = blockIdx.y;
  const float *s = p.scores + (long long)img * p.n;
  uint2 *cand = p.cand + (long long)img * p.cap;
  uint32_t *hist = p.hist + (long long)img * kHistBins;
  uint2 *st = stage[warp];
  const unsigned lt_mask = (1u << lane) - 1u;
  int nstaged = 0;  // warp-uniform

  auto flush = [&]() {
    if (nstaged > 0) {
      __syncwarp();
      int base = 0;
      if (lane == 0) base = atomicAdd(p.counts + img, nstaged);
      base = __shfl_sync(0xffffffffu, base, 0);
      for (int j = lane; j < nstaged; j += 32) {
        uint2 c = st[j];
        long long dst = (long long)base + j;
        if (dst < p.cap) cand[dst] = c;
        atomicAdd(hist + hist_bin(c.x, p.key_thresh, p.shift), 1u);
      }
      __syncwarp();
      nstaged = 0;
    }
  };
  auto push = [&](bool pass, float v, long long idx) {
    unsigned m = __ballot_sync(0xffffffffu, pass);
    if (m) {
      if (pass) st[nstaged + __popc(m & lt_mask)] = make_uint2(odtk_float_key(v), (uint32_t)idx);
      nstaged += __popc(m);
    }
  };

  const long long ntiles = (p.n + kTile - 1) / kTile;
  const long long wstride = (long long)gridDim.x * kWarpsPerBlock;
  for (long long t = (long long)blockIdx.x * kWarpsPerBlock + warp; t < ntiles; t += wstride) {
    const long long e0 = t * kTile;
    if (VEC) {
      float4 v[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        long long e = e0 + (long long)(j * 32 + lane) * 4;
        v[j] = (e < p.n) ? odtk_ld_stream_f4(reinterpret_cast<const float4 *>(s + e))
                         : make_float4(p.thresh, p.thresh, p.thresh, p.thresh);
      }
#pragma unroll
      for (int j = 0; j < 4; j++) {
        long long e = e0 + (long long)(j * 32 + lane) * 4;
        // (e < n) is implied by the thresh fill: thresh > thresh is false
        push(v[j].x > p.thresh, v[j].x, e + 0);
        push(v[j].y > p.thresh, v[j].y, e + 1);
        push(v[j].z > p.thresh, v[j].z, e + 2);
        push(v[j].w > p.thresh, v[j].w, e + 3);
      }
    } else {
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; j++) {
        long long e = e0 + j * 32 + lane;
        v[j] = (e < p.n) ? odtk_ld_stream_f1(s + e) : p.thresh;
      }
#pragma unroll
      for (int j = 0; j < 16; j++) push(v[j] > p.thresh, v[j], e0 + j * 32 + lane);
    }
    if (nstaged >= 32) flush();
  }
  flush();
}

// ------------------------------------------------------------------------------------
struct SelectParams {
  const float *scores;  // [B, n] dense scores (slow path + nothing else); may be NULL
  const float *deltas;  // [B, A*NBOX, H, W]
  long long n;
  int height, width, scale, num_anchors, num_classes;
  int has_anchors;
  float thresh;
  uint32_t key_thresh;
  int shift;
  int top_n;
  const int *counts;
  const uint32_t *hist;
  const uint2 *cand;
  long long cap;
  float *out_scores, *out_boxes, *out_classes;
  long long out_stride, out_offset;
};

// Block-wide radix select (8 passes x 8 bits, MSB first) of the `need`-th largest unique
// 64-bit composite among the items produced by src(j) for j in [0, total).  Returns the
// smallest composite that belongs to the top `need`.  Exact for any input; only used when
// the fast path cannot be (candidate overflow / a tie bin larger than kSortCap).
template <class Src>
__device__ unsigned long long radix_select_kth(Src src, long long total, int need, uint32_t *sh256,
                                               int *sh_misc) {
  unsigned long long prefix = 0;
  for (int pass = 7; pass >= 0; pass--) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sh256[i] = 0;
    __syncthreads();
    const int sh = pass * 8;
    for (long long j = threadIdx.x; j < total; j += blockDim.x) {
      unsigned long long c;
      if (src(j, c)) {
        bool match = (pass == 7) ? true : ((c >> (sh + 8)) == prefix);
        if (match) atomicAdd(&sh256[(c >> sh) & 255], 1u);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int acc = 0, d = 255;
      for (; d > 0; d--) {
        if (acc + (int)sh256[d] >= need) break;
        acc += (int)sh256[d];
      }
      sh_misc[0] = d;
      sh_misc[1] = need - acc;             // how many are still needed inside digit d
      sh_misc[2] = (int)sh256[d];
    }
    __syncthreads();
    prefix = (prefix << 8) | (unsigned long long)sh_misc[0];
    need = sh_misc[1];
    const int in_bucket = sh_misc[2];
    __syncthreads();
    if (in_bucket == need) return prefix << sh;  // the whole bucket is selected
  }
  return prefix;
}

// K2: grid = B, block = 1024.
template <int NBOX>
__global__ void __launch_bounds__(1024) select_decode_kernel(SelectParams p, AnchorTable anchors) {
  __shared__ unsigned long long skey[kSortCap];
  __shared__ uint32_t shist[kHistBins];
  __shared__ int s_wsum[32];
  __shared__ int s_misc[8];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int img = blockIdx.x;
  const long long total = p.counts[img];
  const uint2 *cand = p.cand + (long long)img * p.cap;
  const float *dense = p.scores ? p.scores + (long long)img * p.n : nullptr;

  // composite in SCORE mode: (score key << 32) | ~index      -> score desc, index asc
  // composite in INDEX mode: (~index << 32) | score key       -> index asc
  bool index_mode = total <= (long long)p.top_n;
  int nsel = 0;

  if (t == 0) s_misc[3] = 0;
  __syncthreads();

  if (index_mode) {
    nsel = (int)total;  // total <= top_n <= cap
    for (int j = t; j < nsel; j += blockDim.x) {
      uint2 c = cand[j];
      skey[j] = ((unsigned long long)(~c.y) << 32) | c.x;
    }
  } else {
    bool slow = total > p.cap;
    int bstar = 0;
    if (!slow) {
      // histogram suffix scan: find the highest bin b* with sum_{bin >= b*} >= top_n
      const uint32_t *hist = p.hist + (long long)img * kHistBins;
      for (int i = t; i < kHistBins; i += blockDim.x) shist[i] = hist[i];
      __syncthreads();
      // thread t owns bins (kHistBins-1-2t, kHistBins-2-2t): descending order
      int b0 = kHistBins - 1 - 2 * t, b1 = b0 - 1;
      int c0 = (int)shist[b0], c1 = (int)shist[b1];
      int own = c0 + c1, incl = own;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      if (lane == 31) s_wsum[warp] = incl;
      __syncthreads();
      if (warp == 0) {
        int w = s_wsum[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          int v = __shfl_up_sync(0xffffffffu, wi, o);
          if (lane >= o) wi += v;
        }
        s_wsum[lane] = wi - w;  // exclusive
      }
      __syncthreads();
      incl += s_wsum[warp];
      int excl = incl - own;
      if (excl < p.top_n && incl >= p.top_n) {
        if (excl + c0 >= p.top_n) { s_misc[0] = b0; s_misc[1] = excl + c0; }
        else                      { s_misc[0] = b1; s_misc[1] = incl; }
      }
      __syncthreads();
      bstar = s_misc[0];
      nsel = s_misc[1];
      if (nsel > kSortCap) slow = true;
    }
    if (!slow) {
      const int cnt = (int)total;
      for (int j = t; j < cnt; j += blockDim.x) {
        uint2 c = cand[j];
        if (hist_bin(c.x, p.key_thresh, p.shift) >= bstar) {
          int pos = atomicAdd(&s_misc[3], 1);
          skey[pos] = ((unsigned long long)c.x << 32) | (uint32_t)(~c.y);
        }
      }
      __syncthreads();
      nsel = s_misc[3];
    } else {
      // exact slow path: radix select on the composite key
      unsigned long long kth;
      const float thresh = p.thresh;
      if (total > p.cap) {
        if (dense == nullptr) { __trap(); }
        auto src = [=](long long j, unsigned long long &c) {
          float v = dense[j];
          if (!(v > thresh)) return false;
          c = ((unsigned long long)odtk_float_key(v) << 32) | (uint32_t)(~(uint32_t)j);
          return true;
        };
        kth = radix_select_kth(src, p.n, p.top_n, shist, s_misc);
        for (long long j = t; j < p.n; j += blockDim.x) {
          unsigned long long c;
          if (src(j, c) && c >= kth) skey[atomicAdd(&s_misc[3], 1)] = c;
        }
      } else {
        auto src = [=](long long j, unsigned long long &c) {
          uint2 v = cand[j];
          c = ((unsigned long long)v.x << 32) | (uint32_t)(~v.y);
          return true;
        };
        kth = radix_select_kth(src, total, p.top_n, shist, s_misc);
        for (long long j = t; j < total; j += blockDim.x) {
          unsigned long long c;
          if (src(j, c) && c >= kth) skey[atomicAdd(&s_misc[3], 1)] = c;
        }
      }
      __syncthreads();
      nsel = s_misc[3];  // == top_n
    }
  }

  // order the selected keys
  const int P = odtk_next_pow2(nsel);
  for (int j = nsel + t; j < P; j += blockDim.x) skey[j] = 0ull;
  __syncthreads();
  odtk_bitonic_desc_u64(skey, P);

  // decode: reference decode.cu:119-159 / decode_rotate.cu:115-166
  const int n_out = nsel < p.top_n ? nsel : p.top_n;
  const int H = p.height, W = p.width, A = p.num_anchors, C = p.num_classes;
  const float *d = p.deltas + (long long)img * ((long long)A * NBOX * H * W);
  float *os = p.out_scores + (long long)img * p.out_stride + p.out_offset;
  float *ob = p.out_boxes + ((long long)img * p.out_stride + p.out_offset) * NBOX;
  float *oc = p.out_classes + (long long)img * p.out_stride + p.out_offset;
  for (int k = t; k < p.top_n; k += blockDim.x) {
    if (k < n_out) {
      unsigned long long c = skey[k];
      uint32_t key = index_mode ? (uint32_t)c : (uint32_t)(c >> 32);
      int i = (int)(index_mode ? ~(uint32_t)(c >> 32) : ~(uint32_t)c);
      int x = i % W;
      int y = (i / W) % H;
      int a = (i / C / H / W) % A;
      int cls = (i / H / W) % C;
      float box[NBOX];
#pragma unroll
      for (int q = 0; q < NBOX; q++) box[q] = d[((long long)(a * NBOX + q) * H + y) * W + x];
      if (p.has_anchors) {
        float fx = (float)((long long)x * p.scale);
        float fy = (float)((long long)y * p.scale);
        const float *an = anchors.v + 4 * a;
        float x1 = fx + an[0];
        float y1 = fy + an[1];
        float x2 = fx + an[2];
        float y2 = fy + an[3];
        float w = x2 - x1 + 1.0f;
        float h = y2 - y1 + 1.0f;
        float pred_ctr_x = box[0] * w + x1 + 0.5f * w;
        float pred_ctr_y = box[1] * h + y1 + 0.5f * h;
        float pred_w = expf(box[2]) * w;
        float pred_h = expf(box[3]) * h;
        box[0] = fmaxf(0.0f, pred_ctr_x - 0.5f * pred_w);
        box[1] = fmaxf(0.0f, pred_ctr_y - 0.5f * pred_h);
        box[2] = fminf(pred_ctr_x + 0.5f * pred_w - 1.0f, (float)((long long)W * p.scale) - 1.0f);
        box[3] = fminf(pred_ctr_y + 0.5f * pred_h - 1.0f, (float)((long long)H * p.scale) - 1.0f);
      }
      os[k] = odtk_key_float(key);
#pragma unroll
      for (int q = 0; q < NBOX; q++) ob[(long long)k * NBOX + q] = box[q];
      oc[k] = (float)cls;
    } else {
      os[k] = 0.0f;
#pragma unroll
      for (int q = 0; q < NBOX; q++) ob[(long long)k * NBOX + q] = 0.0f;
      oc[k] = 0.0f;
    }
  }
}

// workspace layout: counts | hist | cand
struct DecodeWs {
  size_t counts_off, hist_off, cand_off, total;
  long long cap;
};
DecodeWs decode_ws_layout(int batch, long long n, int top_n) {
  DecodeWs w;
  long long cap = n < (1ll << 20) ? n : (1ll << 20);
  if (cap < top_n) cap = top_n;
  w.cap = cap;
  w.counts_off = 0;
  w.hist_off = odtk_align_up((size_t)batch * sizeof(int));
  w.cand_off = w.hist_off + odtk_align_up((size_t)batch * kHistBins * sizeof(uint32_t));
  w.total = w.cand_off + odtk_align_up((size_t)batch * (size_t)cap * sizeof(uint2));
  return w;
}

int choose_shift(uint32_t key_thresh) {
  uint32_t hi = odtk_float_key(1.0f);
  if (hi <= key_thresh) hi = key_thresh + 1;
  unsigned long long span = (unsigned long long)hi - key_thresh + 1;
  int s = 0;
  while ((span >> s) >= (unsigned long long)kHistBins) s++;
  return s;
}

}  // namespace

extern "C" long long odtk_decode_ex(int batch, const void *const *inputs, void *const *outputs,
                                    size_t height, size_t width, size_t scale, size_t num_anchors,
                                    size_t num_classes, const float *anchors, size_t num_anchor_floats,
                                    float score_thresh, int top_n, int nbox, size_t out_stride,
                                    size_t out_offset, void *workspace, size_t workspace_size,
                                    odtk_stream_t stream_) {
  if (batch <= 0 || height == 0 || width == 0 || num_anchors == 0 || num_classes == 0 || top_n <= 0)
    return ODTK_E_INVALID;
  if (nbox != 4 && nbox != 6) return ODTK_E_INVALID;
  if (top_n > ODTK_MAX_TOP_N || num_anchors > (size_t)kMaxAnchors) return ODTK_E_UNSUPPORTED;
  if (num_anchor_floats != 0 && num_anchor_floats != 4 * num_anchors) return ODTK_E_INVALID;
  const long long n = (long long)num_anchors * num_classes * height * width;
  if (n >= (1ll << 31)) return ODTK_E_UNSUPPORTED;  // flat index is int32 (decode.cu:122)
  DecodeWs ws = decode_ws_layout(batch, n, top_n);
  if (!workspace || !workspace_size) return (long long)ws.total;
  if (workspace_size < ws.total) return ODTK_E_WORKSPACE;
  if (!inputs || !outputs || !inputs[0] || !inputs[1] || !outputs[0] || !outputs[1] || !outputs[2])
    return ODTK_E_INVALID;
  if (num_anchor_floats && !anchors) return ODTK_E_INVALID;
  if (out_stride < out_offset + (size_t)top_n) return ODTK_E_INVALID;
  cudaStream_t stream = (cudaStream_t)stream_;
  char *base = (char *)workspace;

  FilterParams fp;
  fp.scores = (const float *)inputs[0];
  fp.n = n;
  fp.thresh = score_thresh;
  fp.key_thresh = odtk_float_key(score_thresh);
  fp.shift = choose_shift(fp.key_thresh);
  fp.counts = (int *)(base + ws.counts_off);
  fp.hist = (uint32_t *)(base + ws.hist_off);
  fp.cand = (uint2 *)(base + ws.cand_off);
  fp.cap = ws.cap;
  if (cudaMemsetAsync(base, 0, ws.cand_off, stream) != cudaSuccess) return ODTK_E_CUDA;

  const long long ntiles = (n + kTile - 1) / kTile;
  long long blocks = (ntiles + kWarpsPerBlock - 1) / kWarpsPerBlock;
  // 6 resident CTAs per SM (34.8 KB of staging each) x 148 SMs, shared by the batch
  long long max_blocks = (148ll * 6 + batch - 1) / batch;
  if (max_blocks < 1) max_blocks = 1;
  if (blocks > max_blocks) blocks = max_blocks;
  dim3 grid((unsigned)blocks, (unsigned)batch);
  const bool vec = (n % 4 == 0) && (((uintptr_t)inputs[0]) % 16 == 0);
  {
    OdtkProfScope prof(ODTK_PROF_FILTER, stream);
    if (vec) score_filter_kernel<true><<<grid, kFilterThreads, 0, stream>>>(fp);
    else     score_filter_kernel<false><<<grid, kFilterThreads, 0, stream>>>(fp);
  }

  SelectParams sp;
  sp.scores = fp.scores;
  sp.deltas = (const float *)inputs[1];
  sp.n = n;
  sp.height = (int)height; sp.width = (int)width; sp.scale = (int)scale;
  sp.num_anchors = (int)num_anchors; sp.num_classes = (int)num_classes;
  sp.has_anchors = num_anchor_floats != 0;
  sp.thresh = score_thresh;
  sp.key_thresh = fp.key_thresh;
  sp.shift = fp.shift;
  sp.top_n = top_n;
  sp.counts = fp.counts; sp.hist = fp.hist; sp.cand = fp.cand; sp.cap = ws.cap;
  sp.out_scores = (float *)outputs[0];
  sp.out_boxes = (float *)outputs[1];
  sp.out_classes = (float *)outputs[2];
  sp.out_stride = (long long)out_stride;
  sp.out_offset = (long long)out_offset;
  AnchorTable at;
  for (size_t i = 0; i < 4 * (size_t)kMaxAnchors; i++) at.v[i] = i < num_anchor_floats ? anchors[i] : 0.0f;
  {
    OdtkProfScope prof(ODTK_PROF_SELECT, stream);
    if (nbox == 4) select_decode_kernel<4><<<batch, 1024, 0, stream>>>(sp, at);
    else           select_decode_kernel<6><<<batch, 1024, 0, stream>>>(sp, at);
  }
  return cudaGetLastError() == cudaSuccess ? ODTK_OK : ODTK_E_CUDA;
}

extern "C" long long odtk_decode(int batch, const void *const *inputs, void *const *outputs, size_t height,
                                 size_t width, size_t scale, size_t num_anchors, size_t num_classes,
                                 const float *anchors, size_t num_anchor_floats, float score_thresh,
                                 int top_n, void *workspace, size_t workspace_size, odtk_stream_t stream) {
  return odtk_decode_ex(batch, inputs, outputs, height, width, scale, num_anchors, num_classes, anchors,
                        num_anchor_floats, score_thresh, top_n, 4, (size_t)top_n, 0, workspace,
                        workspace_size, stream);
}

extern "C" long long odtk_decode_rotate(int batch, const void *const *inputs, void *const *outputs,
                                        size_t height, size_t width, size_t scale, size_t num_anchors,
                                        size_t num_classes, const float *anchors, size_t num_anchor_floats,
                                        float score_thresh, int top_n, void *workspace,
                                        size_t workspace_size, odtk_stream_t stream) {
  return odtk_decode_ex(batch, inputs, outputs, height, width, scale, num_anchors, num_classes, anchors,
                        num_anchor_floats, score_thresh, top_n, 6, (size_t)top_n, 0, workspace,
                        workspace_size, stream);
}
