// conv.cu -- im2col-free implicit-GEMM convolution on the 5th-gen tensor cores (sm_100a).
//
// Replaces the cuDNN convolutions behind nn.Conv2d in the reference's Model.forward
// (odtk/model.py:57-68,130-135, odtk/backbones/fpn.py:45-61, torchvision resnet blocks): the
// reference ships no convolution kernel of its own.  One persistent, warp-specialised kernel:
//
//   D[pixels, Cout] = sum over taps (r,s) and 64-channel chunks of  A_tap[pixels, 64] * W[Cout, 64]^T
//
//   warp 0   TMA producer: for every (tap, chunk) one cp.async.bulk.tensor load of the SHIFTED
//            activation patch straight from the NHWC tensor (4-D tensor map, box = 64 ch x TW x TH;
//            the hardware zero-fills out-of-bounds rows/columns == conv padding, so no im2col and
//            no border code) and one of the weight tile (2-D map over [Cout, taps*Cin]); both land
//            in 128-byte-swizzled shared memory stages guarded by full/empty mbarriers.
//   warp 1   MMA issuer: one thread issues tcgen05.mma (cta_group::1, kind::f16, M=128, N<=256,
//            K=16) x4 per stage, fp32 accumulation in TMEM; tcgen05.commit releases the stage and,
//            after the last K block, hands the accumulator to the epilogue.
//   warp 2   TMEM allocator (512 columns = two accumulator buffers, so the epilogue of tile i
//            overlaps the MMAs of tile i+1).
//   warps 4-7 epilogue: tcgen05.ld the accumulator (one output pixel per thread), fused
//            bias (folded BatchNorm) + residual add + FPN nearest-upsample add + ReLU, fp16 NHWC
//            store -- or, for the last convolution of a head, (sigmoid +) fp32 NCHW store in the
//            layout the reference's decode entry point expects.
//
// 1x1 convolutions use the same kernel with a 2-D [pixels, Cin] map (plain GEMM rows).
// Strided and 7x7 convolutions are lowered by layers.cu to one of the two forms.
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "conv.cuh"
#include "prof.cuh"

namespace {

constexpr int kStages = 4;
constexpr int kABytes = 128 * 128;          // 128 rows x 64 fp16
constexpr int kBBytesMax = 256 * 128;       // up to 256 rows x 64 fp16
constexpr int kStageBytes = kABytes + kBBytesMax;
constexpr int kSlabRowBytes = 128 + 16;       // 64 fp16 + 16 B pad: conflict-free 16-byte accesses
constexpr int kSlabBytes = 32 * kSlabRowBytes; // per epilogue warp
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + 4 * kSlabBytes;
constexpr int kThreads = 256;
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;             // columns between the two accumulator buffers

// ---------------------------------------------------------------------------------- PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major, 128-byte swizzle: rows of 64 fp16 = 128 B, 8-row
// groups 1024 B apart (SBO), LBO unused for swizzled K-major layouts, descriptor version 1.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);   // start address  [0,14)
  d |= (uint64_t)1 << 16;                      // LBO (ignored)  [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;            // SBO = 1024 B   [32,46)
  d |= (uint64_t)1 << 46;                      // version = 1    [46,48)
  d |= (uint64_t)2 << 61;                      // SWIZZLE_128B   [61,64)
  return d;
}

struct Barriers {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ float sigmoidf_accurate(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---------------------------------------------------------------------------------- kernel
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ ConvParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  Barriers *bars = reinterpret_cast<Barriers *>(smem + kStages * kStageBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; s++) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    for (int b = 0; b < 2; b++) { mbar_init(&bars->tmem_full[b], 1); mbar_init(&bars->tmem_empty[b], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                 "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  const int total_tiles = p.num_m_tiles * p.num_n_tiles;
  const int kblocks = p.taps * p.kblocks_per_tap;
  const uint32_t a_bytes = (p.mode == 1) ? (uint32_t)(p.TH * p.TW * 128) : (uint32_t)kABytes;
  const uint32_t b_bytes = (uint32_t)p.BN * 128u;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_tile = tile % p.num_m_tiles, n_tile = tile / p.num_m_tiles;
        const int n0 = n_tile * p.BN;
        int img = 0, h0 = 0, w0 = 0;
        if (p.mode == 1) {
          const int per_img = p.tiles_h * p.tiles_w;
          img = m_tile / per_img;
          const int r = m_tile - img * per_img;
          h0 = (r / p.tiles_w) * p.TH;
          w0 = (r % p.tiles_w) * p.TW;
        }
        for (int tap = 0; tap < p.taps; tap++) {
          const int dy = tap / p.kw - p.pad, dx = tap % p.kw - p.pad;
          for (int kb = 0; kb < p.kblocks_per_tap; kb++) {
            mbar_wait(&bars->empty[stage], phase ^ 1u);
            unsigned char *sa = smem + stage * kStageBytes;
            unsigned char *sb = sa + kABytes;
            mbar_arrive_expect_tx(&bars->full[stage], a_bytes + b_bytes);
            if (p.mode == 1) tma_load_4d(sa, &tmA, &bars->full[stage], kb * 64, w0 + dx, h0 + dy, img);
            else             tma_load_2d(sa, &tmA, &bars->full[stage], kb * 64, m_tile * 128);
            tma_load_2d(sb, &tmB, &bars->full[stage], (tap * p.kblocks_per_tap + kb) * 64, n0);
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ========================================
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=f16, both K-major, N = BN, M = 128
      const uint32_t idesc = (1u << 4) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
        const int buf = it & 1;
        mbar_wait(&bars->tmem_empty[buf], ((uint32_t)(it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * kAccStride);
        for (int kb = 0; kb < kblocks; kb++) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint64_t da = make_desc_sw128(sa), db = make_desc_sw128(sa + kABytes);
#pragma unroll
          for (int k = 0; k < 4; k++)  // 4 x UMMA_K(16) = 64 channels; +32 B per step inside the swizzle atom
            tc_mma_f16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
          tc_commit(&bars->empty[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        tc_commit(&bars->tmem_full[buf]);
      }
    }
  } else if (warp >= 4) {
    // ===================================== epilogue ==========================================
    // One accumulator row (= output pixel) per thread.  NHWC fp16 output goes through a padded
    // shared-memory staging slab per warp (32 rows x 64 channels) so that global traffic is
    // row-contiguous: residual / upsample rows are READ coalesced into the slab, combined in fp32
    // with the accumulator in place, and the slab is WRITTEN back as full 128-byte lines.
    const int q = warp - 4;             // TMEM lane quarter == warp_id % 4
    const int row = q * 32 + lane;      // accumulator row == pixel inside the tile
    unsigned char *slab = smem + kStages * kStageBytes + 256 + q * kSlabBytes;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
      const int buf = it & 1;
      const int m_tile = tile % p.num_m_tiles, n_tile = tile / p.num_m_tiles;
      const int n0 = n_tile * p.BN;
      // geometry of an arbitrary row of this tile (used for the own row and for slab rows)
      auto locate = [&](int rr, int &img, int &h, int &w) -> bool {
        if (p.mode == 1) {
          const int per_img = p.tiles_h * p.tiles_w;
          img = m_tile / per_img;
          const int r = m_tile - img * per_img;
          h = (r / p.tiles_w) * p.TH + rr / p.TW;
          w = (r % p.tiles_w) * p.TW + rr % p.TW;
          return rr < p.TH * p.TW && h < p.H && w < p.W;
        }
        const long long m = (long long)m_tile * 128 + rr;
        const int hw = p.H * p.W;
        img = (int)(m / hw);
        const int rem = (int)(m - (long long)img * hw);
        h = rem / p.W;
        w = rem - h * p.W;
        return m < p.M;
      };
      int img, h, w;
      const bool valid = locate(row, img, h, w);
      mbar_wait(&bars->tmem_full[buf], (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * kAccStride);
      const int nchunks = p.BN >> 4;

      if (p.out_mode == ODTK_OUT_NHWC_F16) {
        // slab geometry: 8 lanes x 16 B cover the 64 channels of one row; 4 rows per warp access
        const int sub = lane >> 3, l8 = lane & 7;
        int pix8[8], up8[8];   // pixel index (-1: outside) of slab rows sub, 4+sub, ..., 28+sub
#pragma unroll
        for (int k = 0; k < 8; k++) {
          int i2, h2, w2;
          const bool ok = locate(q * 32 + k * 4 + sub, i2, h2, w2);
          pix8[k] = ok ? (int)(((long long)i2 * p.H + h2) * p.W + w2) : -1;
          up8[k] = (ok && p.upsample) ? (int)(((long long)i2 * p.up_h + (h2 >> 1)) * p.up_w + (w2 >> 1)) : -1;
        }
        for (int seg = 0; seg * 4 < nchunks; seg++) {
          const int segc = min(4, nchunks - seg * 4);        // 16-column chunks in this segment
          const int colbase = n0 + seg * 64;
          const bool lane_on = l8 * 8 < segc * 16 && colbase + l8 * 8 < p.Cout;
          // ---- stage residual / upsample rows (coalesced 128-byte reads) ----
          if (p.residual || p.upsample) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
              const int r0 = k * 4;
              const bool ok = pix8[k] >= 0 && lane_on;
              uint4 acc = make_uint4(0u, 0u, 0u, 0u);
              if (ok && p.residual) {
                const long long px = pix8[k];
                acc = __ldg(reinterpret_cast<const uint4 *>(p.residual + px * p.ldr + colbase + l8 * 8));
              }
              if (ok && p.upsample) {
                const long long up = up8[k];
                uint4 u = __ldg(reinterpret_cast<const uint4 *>(p.upsample + up * p.Cout + colbase + l8 * 8));
                if (p.residual) {
                  __half2 *a2 = reinterpret_cast<__half2 *>(&acc);
                  const __half2 *u2 = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
                  for (int j = 0; j < 4; j++) a2[j] = __hadd2(a2[j], u2[j]);
                } else acc = u;
              }
              *reinterpret_cast<uint4 *>(slab + (r0 + sub) * kSlabRowBytes + l8 * 16) = acc;
            }
            __syncwarp();
          }
          // ---- accumulator (+bias, +staged addend, ReLU) -> fp16 into the own slab row ----
          for (int cc = 0; cc < segc; cc++) {
            const int c = seg * 4 + cc;
            uint32_t v[16];
            tc_ld16(taddr + (uint32_t)(c * 16), v);
            tc_ld_wait();
            const int col0 = n0 + c * 16;
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; j++) f[j] = __uint_as_float(v[j]);
            if (p.bias && col0 < p.Cout) {
              const float4 *bp = reinterpret_cast<const float4 *>(p.bias + col0);
#pragma unroll
              for (int j = 0; j < 4; j++) {
                float4 b4 = __ldg(bp + j);
                f[4 * j] += b4.x; f[4 * j + 1] += b4.y; f[4 * j + 2] += b4.z; f[4 * j + 3] += b4.w;
              }
            }
            uint4 *srow = reinterpret_cast<uint4 *>(slab + lane * kSlabRowBytes + cc * 32);
            if (p.residual || p.upsample) {
              uint4 r0 = srow[0], r1 = srow[1];
              const __half2 *h0 = reinterpret_cast<const __half2 *>(&r0), *h1 = reinterpret_cast<const __half2 *>(&r1);
#pragma unroll
              for (int j = 0; j < 4; j++) {
                float2 a = __half22float2(h0[j]), b = __half22float2(h1[j]);
                f[2 * j] += a.x; f[2 * j + 1] += a.y; f[8 + 2 * j] += b.x; f[8 + 2 * j + 1] += b.y;
              }
            }
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 16; j++) f[j] = fmaxf(f[j], 0.0f);
            }
            uint4 o0, o1;
            __half2 *q0 = reinterpret_cast<__half2 *>(&o0), *q1 = reinterpret_cast<__half2 *>(&o1);
#pragma unroll
            for (int j = 0; j < 4; j++) {
              q0[j] = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
              q1[j] = __floats2half2_rn(f[8 + 2 * j], f[8 + 2 * j + 1]);
            }
            srow[0] = o0;
            srow[1] = o1;
          }
          __syncwarp();
          // ---- slab -> global, 4 rows x 128 contiguous bytes per warp store ----
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const int r0 = k * 4;
            if (pix8[k] >= 0 && lane_on) {
              const long long px = pix8[k];
              uint4 val = *reinterpret_cast<const uint4 *>(slab + (r0 + sub) * kSlabRowBytes + l8 * 16);
              *reinterpret_cast<uint4 *>(reinterpret_cast<__half *>(p.out) + px * p.ldy + colbase + l8 * 8) = val;
            }
          }
          __syncwarp();
        }
      } else {
        // fp32 NCHW (+ sigmoid): lanes of a warp are consecutive pixels of a row -> coalesced
        const long long cs = (long long)p.H * p.W;
        for (int c = 0; c < nchunks; c++) {
          uint32_t v[16];
          tc_ld16(taddr + (uint32_t)(c * 16), v);
          tc_ld_wait();
          const int col0 = n0 + c * 16;
          const int ncol = min(16, p.Cout - col0);
          if (!valid || ncol <= 0) continue;
          float *o = reinterpret_cast<float *>(p.out) + ((long long)img * p.Cout + col0) * cs + (long long)h * p.W + w;
#pragma unroll
          for (int j = 0; j < 16; j++) {
            if (j < ncol) {
              float x = __uint_as_float(v[j]) + (p.bias ? __ldg(p.bias + col0 + j) : 0.0f);
              if (p.relu) x = fmaxf(x, 0.0f);
              if (p.out_mode == ODTK_OUT_NCHW_F32_SIGMOID) x = sigmoidf_accurate(x);
              o[j * cs] = x;
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&bars->tmem_empty[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
}

// ---------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

bool encode_map(CUtensorMap *m, const void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes,
                const uint32_t *box) {
  EncodeTiledFn fn = get_encode();
  if (!fn) return false;
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; i++) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i < rank - 1; i++) gstr[i] = strides_bytes[i];
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void *>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// pick the spatial patch (TW x TH <= 128 pixels) that wastes the fewest accumulator rows
void choose_patch(int H, int W, int &TH, int &TW) {
  double best = -1;
  TH = 1; TW = 1;
  for (int tw = 1; tw <= 128 && tw <= W; tw++) {
    int thmax = 128 / tw;
    if (thmax > H) thmax = H;
    for (int th = 1; th <= thmax; th++) {
      long long tiles = (long long)((W + tw - 1) / tw) * ((H + th - 1) / th);
      double eff = (double)H * W / ((double)tiles * 128.0);
      // prefer wide patches on ties (coalesced NCHW stores, fewer TMA rows)
      if (eff > best + 1e-9 || (eff > best - 1e-9 && tw > TW)) { best = eff; TH = th; TW = tw; }
    }
  }
}

int g_num_sms = 0;

}  // namespace

extern "C" int odtk_conv2d(const odtk_conv_t *d, odtk_stream_t stream_) {
  if (!d || !d->x || !d->w || !d->y) return ODTK_E_INVALID;
  if (d->n <= 0 || d->h <= 0 || d->width <= 0 || d->cin <= 0 || d->cout <= 0) return ODTK_E_INVALID;
  if (d->ksize != 1 && d->ksize != 3) return ODTK_E_UNSUPPORTED;
  if (d->cin % 64 != 0) return ODTK_E_UNSUPPORTED;  // 64-channel K blocks (128-byte swizzle rows)
  if (d->out_mode < 0 || d->out_mode > 2) return ODTK_E_INVALID;
  if (d->out_mode == ODTK_OUT_NHWC_F16 && (d->cout % 16)) return ODTK_E_UNSUPPORTED;
  if (d->out_mode != ODTK_OUT_NHWC_F16 && (d->residual || d->upsample)) return ODTK_E_UNSUPPORTED;
  if ((long long)d->n * d->h * d->width >= (1ll << 31)) return ODTK_E_UNSUPPORTED;
  if (((uintptr_t)d->x | (uintptr_t)d->w | (uintptr_t)d->y) & 15) return ODTK_E_INVALID;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaFuncSetAttribute(conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess)
      return ODTK_E_CUDA;
  }
  ConvParams p;
  memset(&p, 0, sizeof p);
  p.N = d->n; p.H = d->h; p.W = d->width; p.Cin = d->cin; p.Cout = d->cout;
  p.kw = d->ksize; p.taps = d->ksize * d->ksize; p.pad = d->ksize / 2;
  p.kblocks_per_tap = d->cin / 64;
  p.M = (long long)d->n * d->h * d->width;
  // N tile: whole Cout when it fits 256 columns, else the largest multiple-of-16 divisor-ish tile
  int BN;
  if (d->cout <= 256) BN = (d->cout + 15) / 16 * 16;
  else {
    int nt = (d->cout + 255) / 256;
    BN = ((d->cout + nt - 1) / nt + 15) / 16 * 16;
  }
  p.BN = BN;
  p.num_n_tiles = (d->cout + BN - 1) / BN;
  p.bias = d->bias;
  p.residual = (const __half *)d->residual;
  p.upsample = (const __half *)d->upsample;
  p.out = d->y;
  p.relu = d->relu;
  p.out_mode = d->out_mode;
  p.ldy = d->ldy > 0 ? d->ldy : d->cout;
  p.ldr = d->ldr > 0 ? d->ldr : d->cout;
  p.up_h = d->h / 2; p.up_w = d->width / 2;
  if (p.out_mode == ODTK_OUT_NHWC_F16 && (p.ldy % 8)) return ODTK_E_INVALID;
  if (d->upsample && ((d->h & 1) || (d->width & 1))) return ODTK_E_INVALID;

  CUtensorMap tmA, tmB;
  const uint64_t K = (uint64_t)p.taps * d->cin;
  {
    uint64_t dims[2] = {K, (uint64_t)d->cout};
    uint64_t str[1] = {K * 2};
    uint32_t box[2] = {64, (uint32_t)BN};
    if (!encode_map(&tmB, d->w, 2, dims, str, box)) return ODTK_E_CUDA;
  }
  if (d->ksize == 1) {
    p.mode = 0;
    p.num_m_tiles = (int)((p.M + 127) / 128);
    uint64_t dims[2] = {(uint64_t)d->cin, (uint64_t)p.M};
    uint64_t str[1] = {(uint64_t)d->cin * 2};
    uint32_t box[2] = {64, 128};
    if (!encode_map(&tmA, d->x, 2, dims, str, box)) return ODTK_E_CUDA;
  } else {
    p.mode = 1;
    choose_patch(d->h, d->width, p.TH, p.TW);
    p.tiles_h = (d->h + p.TH - 1) / p.TH;
    p.tiles_w = (d->width + p.TW - 1) / p.TW;
    p.num_m_tiles = d->n * p.tiles_h * p.tiles_w;
    uint64_t dims[4] = {(uint64_t)d->cin, (uint64_t)d->width, (uint64_t)d->h, (uint64_t)d->n};
    uint64_t str[3] = {(uint64_t)d->cin * 2, (uint64_t)d->width * d->cin * 2, (uint64_t)d->h * d->width * d->cin * 2};
    uint32_t box[4] = {64, (uint32_t)p.TW, (uint32_t)p.TH, 1};
    if (!encode_map(&tmA, d->x, 4, dims, str, box)) return ODTK_E_CUDA;
  }
  const int total = p.num_m_tiles * p.num_n_tiles;
  const int grid = total < g_num_sms ? total : g_num_sms;
  {
    OdtkProfScope prof(ODTK_PROF_CONV, stream);
    conv_gemm_kernel<<<grid, kThreads, kSmemBytes, stream>>>(tmA, tmB, p);
  }
  return cudaGetLastError() == cudaSuccess ? ODTK_OK : ODTK_E_CUDA;
}
