"""Debug helper: run the 7x7 stem (odtk_pad_input + odtk_stem_conv) of a given libodtk_b200 build and compare with torch."""
import ctypes, sys, os
import torch, torch.nn.functional as F
lib = ctypes.CDLL(sys.argv[1])
DEV = "cuda:0"
g = torch.Generator().manual_seed(96)
n, h, w = 1, 32, 64
x = (torch.randn((n, h, w, 3), generator=g)).half()
wt = (torch.randn((64, 3, 7, 7), generator=g) * 0.1).half()
b = torch.randn(64, generator=g)
ws = wt.float().new_zeros((64, 7, 8, 4)); ws[:, :, :7, :3] = wt.float().permute(0, 2, 3, 1)
ws = ws.reshape(64, 224).half().contiguous().to(DEV)
xd, bd = x.to(DEV), b.to(DEV)
xp = torch.empty((n, h + 6, w + 8, 4), dtype=torch.float16, device=DEV)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
V = ctypes.c_void_p
for relu in (1, 0):
    lib.odtk_pad_input(V(xd.data_ptr()), V(xp.data_ptr()), n, h, w, st)
    y = torch.full((n, h // 2, w // 2, 64), 7.0, dtype=torch.float16, device=DEV)
    rc = lib.odtk_stem_conv(V(xp.data_ptr()), V(ws.data_ptr()), V(bd.data_ptr()), V(y.data_ptr()), n, h, w, 64, relu, st)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), b, stride=2, padding=3)
    if relu: ref = F.relu(ref)
    err = (y.float().cpu() - ref.permute(0, 2, 3, 1)).abs().max().item()
    print(os.path.basename(sys.argv[1]), "env", {k: v for k, v in os.environ.items() if k.startswith("ODTK")}, "relu", relu, "rc", rc, "max err %.4f" % err, "min %.3f" % float(y.min()), flush=True)
