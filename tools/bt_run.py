"""Run bottleneck_tail a few times on the two benchmark shapes (ncu target)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from retinanet_examples_b200 import engine
g = torch.Generator().manual_seed(0)
for (n, h, w, c1) in ((32, 200, 320, 64), (32, 100, 160, 128)):
    c2 = 4 * c1
    x = (torch.randn((n, h, w, c1), generator=g)).half().cuda()
    res = (torch.randn((n, h, w, c2), generator=g)).half().cuda()
    w2 = engine.pack_weight(torch.randn((c1, c1, 3, 3), generator=g) * 0.04).cuda()
    w3 = engine.pack_weight(torch.randn((c2, c1, 1, 1), generator=g) * 0.08).cuda()
    b2, b3 = torch.randn(c1).cuda(), torch.randn(c2).cuda()
    for _ in range(3):
        engine.bottleneck_tail(x, w2, b2, w3, b3, res)
    torch.cuda.synchronize()
