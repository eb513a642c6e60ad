// polygon.cuh -- box-overlap arithmetic shared by nms.cu (greedy NMS) and iou.cu (rotated target assignment):
// the Sutherland-Hodgman clip + shoelace of the reference's rotated IoU restated operation for operation
// (csrc/cuda/nms_iou.cu:87-169,182-248: bit-exact kept indices need the same arithmetic order), and the axis-aligned
// overlap of csrc/cuda/nms.cu:57-69.  Compile with -fmad=false -prec-div=true (csrc/Makefile EXACT flags).
#pragma once
#include <cuda_runtime.h>

namespace {

struct f2 { float x, y; };
struct line_t { float a, b, c; };

__device__ __forceinline__ line_t make_line(f2 v1, f2 v2) {  // nms_iou.cu:87
  line_t l;
  l.a = v2.y - v1.y;
  l.b = v1.x - v2.x;
  l.c = v2.x * v1.y - v2.y * v1.x;
  return l;
}
__device__ __forceinline__ float line_call(line_t l, f2 v) { return l.a * v.x + l.b * v.y + l.c; }
__device__ __forceinline__ f2 line_isect(line_t l, line_t o) {  // nms_iou.cu:93-96
  float w = l.a * o.b - l.b * o.a;
  f2 r;
  r.x = (l.b * o.c - l.c * o.b) / w;
  r.y = (l.c * o.a - l.a * o.c) / w;
  return r;
}

// Sutherland-Hodgman clip of `inter` (<= 8 points) against the 4 edges of mrect, then
// shoelace: nms_iou.cu:114-169, same operation order.  Writes beyond 8 points (undefined
// in the reference) are dropped.
__device__ float intersection_area(const f2 *mrect, f2 *inter) {
  int count = 4;
  for (int i = 0; i < 4; i++) {
    float lv[8];
    line_t l1 = make_line(mrect[i], mrect[(i + 1) & 3]);
#pragma unroll
    for (int j = 0; j < 8; j++) lv[j] = (j < count) ? line_call(l1, inter[j]) : 0.0f;
    f2 nw[8];
    int temp = count;
    count = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      if (j < temp) {
        int jn = (j + 1 == temp) ? 0 : j + 1;  // rotateLeft(count)
        float lvs = lv[jn];
        if (lv[j] <= 0) {
          if (count < 8) nw[count] = inter[j];
          count++;
        }
        if ((lv[j] * lvs) <= 0) {
          line_t l2 = make_line(inter[j], inter[jn]);
          if (count < 8) nw[count] = line_isect(l1, l2);
          count++;
        }
      }
    }
    if (count > 8) count = 8;
#pragma unroll
    for (int k = 0; k < 8; k++)
      if (k < count) inter[k] = nw[k];
  }
  float area = 0.0f;
  if (count > 2) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (k < count) {
        int kn = (k + 1 == count) ? 0 : k + 1;
        area += inter[k].x * inter[kn].y - inter[k].y * inter[kn].x;
      }
    }
  }
  return fabsf(area / 2.0f);
}

// nms_iou.cu:182-248.  ib / mb = (x1,y1,x2,y2,sin,cos).
__device__ float rotated_overlap(const float *ib, const float *mb, int fixed_angle) {
  const float is = ib[4], ic = ib[5];
  const float ms = fixed_angle ? mb[4] : ib[4], mc = fixed_angle ? mb[5] : ib[5];
  f2 inter[8], irect[4], mrect[4];
  const float icx = (ib[0] + ib[2]) / 2.0f, icy = (ib[1] + ib[3]) / 2.0f;
  const float mcx = (mb[0] + mb[2]) / 2.0f, mcy = (mb[1] + mb[3]) / 2.0f;
  const float ibx[4] = {ib[0] - icx, ib[2] - icx, ib[2] - icx, ib[0] - icx};
  const float iby[4] = {ib[1] - icy, ib[1] - icy, ib[3] - icy, ib[3] - icy};
  const float mbx[4] = {mb[0] - mcx, mb[2] - mcx, mb[2] - mcx, mb[0] - mcx};
  const float mby[4] = {mb[1] - mcy, mb[1] - mcy, mb[3] - mcy, mb[3] - mcy};
#pragma unroll
  for (int b = 0; b < 4; b++) {
    float ix = (ibx[b] * ic - iby[b] * is) + icx;
    float iy = (iby[b] * ic + ibx[b] * is) + icy;
    float mx = (mbx[b] * mc - mby[b] * ms) + mcx;
    float my = (mby[b] * mc + mbx[b] * ms) + mcy;
    float px = (ix == mx) ? 0.001f : 0.0f;
    float py = (iy == my) ? 0.001f : 0.0f;
    inter[b].x = ix + px; inter[b].y = iy + py;
    irect[b].x = ix; irect[b].y = iy;
    mrect[b].x = mx; mrect[b].y = my;
  }
#pragma unroll
  for (int b = 4; b < 8; b++) { inter[b].x = -1.0f; inter[b].y = -1.0f; }
  float ia = intersection_area(mrect, inter);
  float irect_area = 0.0f, mrect_area = 0.0f;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    int kn = (k + 1) & 3;
    irect_area += irect[k].x * irect[kn].y - irect[k].y * irect[kn].x;
    mrect_area += mrect[k].x * mrect[kn].y - mrect[k].y * mrect[kn].x;
  }
  float ua = (fabsf(irect_area) + fabsf(mrect_area)) / 2.0f;
  float overlap;
  if (isnan(ia) && isnan(ua)) overlap = 1.0f;
  else if (isnan(ia)) overlap = 0.0f;
  else overlap = ia / (ua - ia);
  return overlap;
}

// nms.cu:57-69
__device__ __forceinline__ float aligned_overlap(const float *ib, const float *mb) {
  float x1 = fmaxf(ib[0], mb[0]);
  float y1 = fmaxf(ib[1], mb[1]);
  float x2 = fminf(ib[2], mb[2]);
  float y2 = fminf(ib[3], mb[3]);
  float w = fmaxf(0.0f, x2 - x1 + 1);
  float h = fmaxf(0.0f, y2 - y1 + 1);
  float iarea = (ib[2] - ib[0] + 1) * (ib[3] - ib[1] + 1);
  float marea = (mb[2] - mb[0] + 1) * (mb[3] - mb[1] + 1);
  float inter = w * h;
  return inter / (iarea + marea - inter);
}

}  // namespace
