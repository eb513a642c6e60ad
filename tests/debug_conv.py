"""Manual diagnostic (not collected by pytest): one small conv per mode, error summary."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from retinanet_examples_b200 import engine

DEV = "cuda:0"


def run(name, n, h, w, cin, cout, ks, mode=0):
    g = torch.Generator().manual_seed(1)
    x = (torch.randn((n, h, w, cin), generator=g)).half()
    wt = (torch.randn((cout, cin, ks, ks), generator=g) * 0.05).half()
    b = torch.randn(cout, generator=g)
    y = engine.conv2d(x.to(DEV), engine.pack_weight(wt.float()).to(DEV), b.to(DEV), cout, ks, out_mode=mode)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), b, padding=ks // 2)
    got = y.float().cpu() if mode else y.float().cpu().permute(0, 3, 1, 2)
    err = (got - ref).abs()
    print(name, "max err %.4g  ref max %.4g  mean err %.4g" % (err.max().item(), ref.abs().max().item(), err.mean().item()), flush=True)
    if err.max().item() > 0.05:
        idx = (err > 0.05).nonzero()
        print("  bad count", idx.shape[0], "of", err.numel(), "first", idx[:5].tolist(), flush=True)
        print("  got", got.flatten()[:8].tolist(), "\n  ref", ref.flatten()[:8].tolist(), flush=True)
        # which channels / rows are wrong?
        badc = (err > 0.05).any(dim=0).any(dim=1).any(dim=1).nonzero().flatten().tolist()
        print("  bad channels", badc[:40], flush=True)
        badh = (err > 0.05).any(dim=0).any(dim=0).any(dim=1).nonzero().flatten().tolist()
        badw = (err > 0.05).any(dim=0).any(dim=0).any(dim=0).nonzero().flatten().tolist()
        print("  bad h", badh[:40], "bad w", badw[:40], flush=True)


run("1x1 64->64   8x16", 1, 8, 16, 64, 64, 1)
run("1x1 128->256 8x16", 1, 8, 16, 128, 256, 1)
run("1x1 256->512 13x20 n2", 2, 13, 20, 256, 512, 1)
run("3x3 64->64   8x16", 1, 8, 16, 64, 64, 3)
run("3x3 256->256 13x20 n2", 2, 13, 20, 256, 256, 3)
run("3x3 128->128 33x47", 1, 33, 47, 128, 128, 3)
run("3x3 256->720 nchw", 1, 25, 40, 256, 720, 3, 1)
run("3x3 256->36 nchw", 1, 25, 40, 256, 36, 3, 1)
print("done")
