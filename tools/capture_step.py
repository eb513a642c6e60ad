"""Run ONE eager, single-stream step of bench.py's default workload between cudaProfilerStart/Stop, so that
    ncu --profile-from-start off --metrics ... python tools/capture_step.py [--batch B] [--trace out.json]
captures exactly the launches of one step, in the order of the engine trace written to --trace
(one record per launch: kind, shape, algorithmic FLOPs and bytes).  tools/layer_table.py joins the two."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--backbone", default="ResNet50FPN")
    ap.add_argument("--trace", default="gpurun_out/step_trace.json")
    args = ap.parse_args()
    os.environ["ODTK_BENCH_CUDA_GRAPH"] = "0"
    import torch
    import bench
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    with torch.no_grad():
        wl = bench.FullWorkload(args.backbone, args.batch, 0, dev)
        json.dump(wl.trace, open(args.trace, "w"))
        for _ in range(3):
            wl.step_profile()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        wl.step_profile()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
