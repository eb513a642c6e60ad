"""Wait profile of bottleneck_tail_kernel: cycles each role of CTA 0 spends waiting on each barrier, per launch.
    make -C retinanet-examples_b200/csrc btprof
    ODTK_B200_LIB=tools/_ab/lib_btprof.so python tools/bt_prof.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from retinanet_examples_b200 import engine, _lib
NAMES = {0: "mma acc1_empty", 1: "mma pfull", 2: "mma wfull", 3: "mma y1_full", 4: "mma acc2_empty", 6: "mma TOTAL",
         7: "prodA pempty", 8: "prodA wempty", 9: "prodA TOTAL", 10: "prodR rempty", 11: "prodR TOTAL",
         12: "epi acc1_full (sum of 8 warps)", 13: "epi y1_empty (8 warps)", 14: "epi acc2_full (8 warps)", 15: "epi rfull (8 warps)", 16: "epi TOTAL (warp 4)", 17: "epi2 wait_group.read (warp 4)", 18: "epi2 tmem ld + math + sts (warp 4)", 19: "epi2 fences + syncwarp (warp 4)", 20: "epi2 arrive + store issue (warp 4)"}
lib = _lib.lib()
g = torch.Generator().manual_seed(0)
for (n, h, w, c1) in ((32, 200, 320, 64), (32, 100, 160, 128)):
    c2 = 4 * c1
    x = (torch.randn((n, h, w, c1), generator=g)).half().cuda()
    res = (torch.randn((n, h, w, c2), generator=g)).half().cuda()
    w2 = engine.pack_weight(torch.randn((c1, c1, 3, 3), generator=g) * 0.04).cuda()
    w3 = engine.pack_weight(torch.randn((c2, c1, 1, 1), generator=g) * 0.08).cuda()
    b2, b3 = torch.randn(c1).cuda(), torch.randn(c2).cuda()
    for _ in range(2):
        engine.bottleneck_tail(x, w2, b2, w3, b3, res)
    buf = (ctypes.c_ulonglong * 32)()
    lib.odtk_bt_prof_read(buf, 1)
    reps = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        engine.bottleneck_tail(x, w2, b2, w3, b3, res)
    e1.record()
    torch.cuda.synchronize()
    lib.odtk_bt_prof_read(buf, 1)
    tiles = n * ((h + 7) // 8) * ((w + 15) // 16)
    per_cta = tiles / 148.0
    print("shape", (n, h, w, c1), "us/launch %.1f" % (e0.elapsed_time(e1) * 1e3 / reps), "tiles/CTA %.1f" % per_cta, flush=True)
    for k in sorted(NAMES):
        print("  %-34s %10.0f cyc/launch  %8.0f cyc/tile" % (NAMES[k], buf[k] / reps, buf[k] / reps / per_cta))
