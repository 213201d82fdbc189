"""Small-batch greedy decoding only (bench.py's `decode_small` section): eager pipeline / CUDA-graph replay / persistent kernel.
    python tools/bench_decode_small.py  -> one JSON line"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import adaptive_b200  # noqa: E402
import bench  # noqa: E402
from adaptive_b200.synth import CFG_A, make_weights  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    dims = CFG_A

    class Cf:
        adaptive_word_embed_size, adaptive_lstm_hidden_size, vocab_length = dims.E, dims.H, dims.Vc

    model = adaptive_b200.Encoder2Decoder(Cf()).to(dev)
    w = make_weights(dims, seed=123)
    model.load_state_dict({"decoder." + k: torch.from_numpy(v) for k, v in w.items()}, strict=False)
    print(json.dumps(bench.measure_decode_small(torch, dev, model, dims, bench.load_peaks())))


if __name__ == "__main__":
    main()
