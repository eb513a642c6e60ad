"""Feasibility experiment: two half batches on two streams with complementary SM budgets, phase-shifted so that the
tensor-bound heads of one half run next to the (largely HBM-bound) backbone of the other.
    python tools/pipe_bench.py X Y [steps]      X = SMs for heads, Y = SMs for the backbone (0 0 = no budgets)"""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from retinanet_examples_b200 import engine, synth, _lib, _C
from retinanet_examples_b200.model import Model, make_state_dict

X, Y = int(sys.argv[1]), int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
half = int(os.environ.get("HALF", "16"))
dev = torch.device("cuda:0")
lib = _lib.lib()
sd = make_state_dict("ResNet50FPN", 80, 9, False, seed=0)
probe = Model("ResNet50FPN", classes=80)
xs = [torch.randn((half, 3, 800, 1280), generator=torch.Generator().manual_seed(7 + r)).half().contiguous(memory_format=torch.channels_last).to(dev) for r in range(2)]


def gpu_logits(s):
    probe.load_state_dict(s).cuda(0)
    return probe.forward_heads(xs[0][:2], sigmoid=False)[0]


sd = synth.calibrate_cls_head(sd, gpu_logits)
reps = []
for r in range(2):
    m = Model("ResNet50FPN", classes=80).load_state_dict(sd).cuda(0)
    m.parallel_heads = False
    st = torch.cuda.Stream()
    rep = {"m": m, "s": st, "x": xs[r]}
    with torch.no_grad(), torch.cuda.stream(st):
        # warm-up (lazy init, allocations), then capture the two halves of the forward as two graphs
        for _ in range(2):
            m.forward(rep["x"])
        torch.cuda.synchronize()
        lib.odtk_set_sm_budget(Y)
        gb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gb, stream=st):
            feats = m._features(m._stem(m._to_nhwc_half(rep["x"])))
        lib.odtk_set_sm_budget(X)
        sizes = tuple((f.shape[1], f.shape[2]) for f in feats)
        key = (half, sizes, 1280, m.threshold, m.top_n, m.rotated_bbox)
        fd = m._fused[key]
        gh = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gh, stream=st):
            sinks = fd.begin()
            _, box_heads = m._heads(feats, True, sinks)
            out = m._nms(fd.finish(box_heads))
        lib.odtk_set_sm_budget(0)
    rep.update(gb=gb, gh=gh, out=out, feats=feats)
    reps.append(rep)
torch.cuda.synchronize()
A, B = reps
main = torch.cuda.current_stream()


def run(n):
    evA_bb = [torch.cuda.Event() for _ in range(n)]
    evB_bb = [torch.cuda.Event() for _ in range(n)]
    for k in range(n):
        with torch.cuda.stream(A["s"]):
            if k > 0:
                A["s"].wait_event(evB_bb[k - 1])
            A["gb"].replay()
            evA_bb[k].record(A["s"])
            A["gh"].replay()
        with torch.cuda.stream(B["s"]):
            B["s"].wait_event(evA_bb[k])
            B["gb"].replay()
            evB_bb[k].record(B["s"])
            B["gh"].replay()


run(3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(main)
A["s"].wait_event(e0); B["s"].wait_event(e0)
run(steps)
ea, eb = torch.cuda.Event(), torch.cuda.Event()
ea.record(A["s"]); eb.record(B["s"])
main.wait_event(ea); main.wait_event(eb)
e1.record(main)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print("X=%d Y=%d half=%d: %.3f ms per %d images -> %.1f img/s; detections %d" % (X, Y, half, ms, 2 * half, 2 * half / ms * 1e3, int((A["out"][0] > 0).sum() + (B["out"][0] > 0).sum())), flush=True)
