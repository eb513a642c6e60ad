// peer.cu -- buffers one GPU process shares with the other ranks of its node (CUDA IPC over NVLink / NVSwitch): the memory
// the NMS kernel's in-kernel detection gather stores into (odtk_nms_gather, nms.cu).  Plain cudaMalloc + cudaIpc*: the
// importer opens the handle with ITS OWN device current, so the mapping lands in its address space with peer access
// enabled (cudaIpcMemLazyEnablePeerAccess) -- torch's tensor-sharing path opens handles under the OWNER's ordinal instead,
// which is right for same-device sharing and wrong for this.
#include "common.cuh"

extern "C" int odtk_peer_alloc(size_t bytes, void **ptr, void *handle64) {
  if (!ptr || !handle64 || bytes == 0) return ODTK_E_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handles travel as 64 bytes");
  void *p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) return ODTK_E_CUDA;
  if (cudaMemset(p, 0, bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { cudaFree(p); return ODTK_E_CUDA; }
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) { cudaFree(p); return ODTK_E_CUDA; }
  memcpy(handle64, &h, 64);
  *ptr = p;
  return ODTK_OK;
}

extern "C" int odtk_peer_open(const void *handle64, void **ptr) {
  if (!ptr || !handle64) return ODTK_E_INVALID;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  return cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess ? ODTK_OK : ODTK_E_CUDA;
}

extern "C" int odtk_peer_close(void *ptr) { return cudaIpcCloseMemHandle(ptr) == cudaSuccess ? ODTK_OK : ODTK_E_CUDA; }
extern "C" int odtk_peer_free(void *ptr) { return cudaFree(ptr) == cudaSuccess ? ODTK_OK : ODTK_E_CUDA; }
