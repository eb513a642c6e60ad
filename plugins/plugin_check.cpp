// plugin_check.cpp -- compile + link + size-query check of plugins/odtk_b200_plugin.h against libodtk_b200.so, no TensorRT
// and no GPU needed (the workspace queries return before any CUDA call).  Built and run by tests/test_host.py.
#include <cstdio>

#include "odtk_b200_plugin.h"

int main() {
  using namespace nvinfer1;
  std::vector<float> anchors(36, 1.0f);
  odtk_b200::DecodeBackend dec(0.05f, 1000, anchors, 8), decr(0.05f, 1000, std::vector<float>(108, 1.0f), 8, true);
  DynamicPluginTensorDesc in[2] = {}, out[3] = {};
  in[0].desc.dims = Dims{4, {8, 720, 100, 160}};
  in[1].desc.dims = Dims{4, {8, 36, 100, 160}};
  in[0].desc.type = in[1].desc.type = DataType::kFLOAT;
  dec.configurePlugin(in, 2, out, 3);
  PluginTensorDesc io[5] = {in[0].desc, in[1].desc, {}, {}, {}};
  const size_t ws = dec.getWorkspaceSize(io, 2, io + 2, 3);
  in[0].desc.dims = Dims{4, {8, 2160, 100, 160}};
  in[1].desc.dims = Dims{4, {8, 162, 100, 160}};
  decr.configurePlugin(in, 2, out, 3);
  PluginTensorDesc ior[5] = {in[0].desc, in[1].desc, {}, {}, {}};
  const size_t wsr = decr.getWorkspaceSize(ior, 2, ior + 2, 3);
  odtk_b200::NmsBackend nms(0.5f, 100), nmsr(0.5f, 100, true);
  DynamicPluginTensorDesc nin[3] = {};
  nin[0].desc.dims = Dims{2, {8, 5000}};
  nms.configurePlugin(nin, 3, out, 3);
  nmsr.configurePlugin(nin, 3, out, 3);
  PluginTensorDesc nio[1] = {nin[0].desc};
  const size_t wn = nms.getWorkspaceSize(nio, 3, nio, 3), wnr = nmsr.getWorkspaceSize(nio, 3, nio, 3);
  std::printf("decode_ws=%zu decode_rotate_ws=%zu nms_ws=%zu nms_rotate_ws=%zu format_ok=%d\n", ws, wsr, wn, wnr,
              (int)dec.supportsFormatCombination(0, io, 2, 3));
  return (ws > 0 && wsr >= ws && wn > 0 && wnr > 0) ? 0 : 1;
}
