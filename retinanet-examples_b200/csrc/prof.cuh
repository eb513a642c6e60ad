// prof.cuh -- optional per-kernel CUDA-event timing (off by default; bench.py turns it on to
// obtain the live duration of the dominant kernel inside the timed region, on the launching stream).
#pragma once
#include <cuda_runtime.h>

enum OdtkProfTag { ODTK_PROF_FILTER = 0, ODTK_PROF_SELECT = 1, ODTK_PROF_NMS = 2, ODTK_PROF_CONV = 3,
                   ODTK_PROF_LOSS = 4, ODTK_PROF_LAYER = 5 /* pad, max-pool, lowering, preprocess */, ODTK_PROF_NTAGS = 8 };

void odtk_prof_begin(int tag, cudaStream_t s);
void odtk_prof_end(int tag, cudaStream_t s);

struct OdtkProfScope {
  int tag; cudaStream_t s;
  OdtkProfScope(int t, cudaStream_t st) : tag(t), s(st) { odtk_prof_begin(tag, s); }
  ~OdtkProfScope() { odtk_prof_end(tag, s); }
};
