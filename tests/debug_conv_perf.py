"""Manual diagnostic (not collected by pytest): time individual conv shapes of ResNet50FPN at batch 8."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from retinanet_examples_b200 import engine

DEV = "cuda:0"
SHAPES = [  # name, n, h, w, cin, cout, ks, residual, mode
    ("l1.conv3 1x1 64->256 +res", 8, 200, 320, 64, 256, 1, True, 0),
    ("l1.conv1 1x1 256->64", 8, 200, 320, 256, 64, 1, False, 0),
    ("l1.conv2 3x3 64->64", 8, 200, 320, 64, 64, 3, False, 0),
    ("l2.conv3 1x1 128->512 +res", 8, 100, 160, 128, 512, 1, True, 0),
    ("l3.conv3 1x1 256->1024 +res", 8, 50, 80, 256, 1024, 1, True, 0),
    ("head 3x3 256->256 P3", 8, 100, 160, 256, 256, 3, False, 0),
    ("cls final 3x3 256->720 P3", 8, 100, 160, 256, 720, 3, False, 2),
]
only = sys.argv[1] if len(sys.argv) > 1 else None
for name, n, h, w, cin, cout, ks, res, mode in SHAPES:
    if only and only not in name:
        continue
    x = torch.randn((n, h, w, cin), device=DEV).half()
    wt = (torch.randn((cout, ks * ks * cin), device=DEV) * 0.05).half()
    b = torch.randn(cout, device=DEV)
    r = torch.randn((n, h, w, cout), device=DEV).half() if res else None
    for _ in range(2):
        y = engine.conv2d(x, wt, b, cout, ks, relu=True if mode == 0 else False, residual=r, out_mode=mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        y = engine.conv2d(x, wt, b, cout, ks, relu=True if mode == 0 else False, residual=r, out_mode=mode)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 5 * 1e3
    flops = 2.0 * n * h * w * cout * ks * ks * cin
    byt = x.numel() * 2 + (r.numel() * 2 if res else 0) + y.numel() * y.element_size()
    print("%-30s %8.1f us  %7.1f TFLOP/s  %6.2f TB/s" % (name, us, flops / us / 1e6, byt / us / 1e6), flush=True)
