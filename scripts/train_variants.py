"""Training-step variants of the (PyTorch) network around the CM loss: eager vs one CUDA graph over
forward + loss + backward, fp32/TF32 vs bf16 autocast.  Prints ms/step, windows/s and the loss after the timed steps."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="train_128x128_b8")
ap.add_argument("--steps", type=int, default=8)
a = ap.parse_args()
wl = dict(bench.TRAIN_WORKLOADS[a.workload], name=a.workload)
for mode in ("eager", "graph"):
    for dtype in ("f32", "bf16"):
        try:
            res = bench.run_train(argparse.Namespace(steps=a.steps, warmup=3, train_mode=mode, train_dtype=dtype), wl, quiet=True)
            print("%-6s %-5s %8.2f ms/step %9.0f windows/s  loss %.6f" % (mode, dtype, res["ms_per_step"], res["value"], res["loss"]), flush=True)
        except Exception as exc:
            import traceback

            traceback.print_exc()
            print("%-6s %-5s FAILED %r" % (mode, dtype, exc), flush=True)
