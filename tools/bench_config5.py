"""BASELINE config 5 shapes on ONE GPU: the graphed training step with the loss fused into the vocabulary projection vs the two-step route.
    python tools/bench_config5.py"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one():
    import torch

    import bench
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    barrier = torch.cuda.synchronize
    out = bench.measure_config5(torch, None, dev, 1, 0, barrier, lambda x: x, bench.load_peaks(), 20)
    print(json.dumps({"AA_FUSED_CE": os.environ.get("AA_FUSED_CE", "auto"), "ms_per_step": out["ms_per_step"], "value": out["value"]}))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        one()
    else:
        for mode in ("0", "1", "auto"):
            subprocess.run([sys.executable, __file__, "run"], env=dict(os.environ, AA_FUSED_CE=mode), check=False)
