"""Generate the golden vectors under ``tests/golden/`` by running the UNMODIFIED reference
modules (``/root/reference/code_src/models/adaptive_attention.py``) in this container.

TEST INFRASTRUCTURE ONLY.  Run here (the reference does not exist on the GPU box):

    python oracle/gen_golden.py            # writes tests/golden/*.npz

Weights/inputs come from ``adaptive_b200.synth`` (numpy PCG64, reproducible anywhere) and are
pushed into the reference ``Decoder`` with ``load_state_dict``; only the reference's OUTPUTS
are stored, so the fixtures stay small.  Reference quirks handled as in SURVEY.md §8c:
batched greedy = loop body of ``sampler`` (adaptive_attention.py:186-216) with ``[1,B,H]``
states (Q9); gradients via ``Decoder`` directly with leaf inputs (Q10); fp64 runs via
``torch.set_default_dtype`` (Q11).
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

from adaptive_b200.synth import Dims, make_inputs, make_lengths, make_weights  # noqa: E402
from code_src.models import adaptive_attention as ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# name -> (Dims, B, T, greedy max_len)
CASES = {
    "tiny": (Dims(H=32, E=16, Vc=40, k=49), 3, 5, 6),
    "k196": (Dims(H=64, E=32, Vc=100, k=196), 2, 4, 5),       # Q1: regions != 49
    "odd": (Dims(H=48, E=20, Vc=77, k=10), 5, 7, 7),          # ragged sizes for tail paths
    "cfgA": (Dims(H=512, E=256, Vc=10000, k=49), 4, 18, 20),  # BASELINE config 1 shapes
}


def build_ref(dims: Dims, w, dtype):
    torch.set_default_dtype(dtype)
    dec = ref.Decoder(dims.E, dims.Vc, dims.H, None)
    sd = {k: torch.from_numpy(v).to(dtype) for k, v in w.items()}
    dec.load_state_dict(sd, strict=True)
    return dec.to(dtype)


def run_case(name, dims, B, T, L):
    out = {}
    for tag, tdt, ndt in (("f32", torch.float32, np.float32), ("f64", torch.float64, np.float64)):
        w = make_weights(dims, seed=123, dtype=np.float32, bias_scale=0.1)
        inp = make_inputs(dims, B, T, seed=1234, dtype=np.float32)
        w = {k: v.astype(ndt) for k, v in w.items()}
        dec = build_ref(dims, w, tdt)
        V = torch.from_numpy(inp["V"].astype(ndt)).requires_grad_(True)
        v_g = torch.from_numpy(inp["v_g"].astype(ndt)).requires_grad_(True)
        h0 = torch.from_numpy(inp["h0"].astype(ndt))[None].requires_grad_(True)
        c0 = torch.from_numpy(inp["c0"].astype(ndt))[None].requires_grad_(True)
        cap = torch.from_numpy(inp["captions"])
        # ---- teacher-forced forward (Decoder.forward) ----
        scores, alpha, beta, (hT, cT) = dec(V, v_g, cap, (h0, c0))
        # ---- backward with a fixed random upstream gradient ----
        rng = np.random.Generator(np.random.PCG64(99))
        dS = rng.standard_normal(scores.shape).astype(ndt) / scores.shape[-1]
        dA = rng.standard_normal(alpha.shape).astype(ndt) * 0.1
        dB = rng.standard_normal(beta.shape).astype(ndt) * 0.1
        dH = rng.standard_normal(hT.shape).astype(ndt) * 0.1
        dC = rng.standard_normal(cT.shape).astype(ndt) * 0.1
        loss = ((scores * torch.from_numpy(dS)).sum() + (alpha * torch.from_numpy(dA)).sum()
                + (beta * torch.from_numpy(dB)).sum() + (hT * torch.from_numpy(dH)).sum()
                + (cT * torch.from_numpy(dC)).sum())
        loss.backward()
        big = scores.numel() > 200000
        sc = scores.detach().numpy()
        if big:  # keep the fixture small: strided sample + per-row summaries
            out[tag + "_scores_sub"] = sc[:, :, ::97].copy()
            out[tag + "_scores_max"] = sc.max(-1)
            out[tag + "_scores_argmax"] = sc.argmax(-1)
            out[tag + "_scores_sum"] = sc.sum(-1)
        else:
            out[tag + "_scores"] = sc
        out[tag + "_alpha"] = alpha.detach().numpy()
        out[tag + "_beta"] = beta.detach().numpy()
        out[tag + "_hT"] = hT.detach().numpy()[0]
        out[tag + "_cT"] = cT.detach().numpy()[0]
        for k, p in dec.named_parameters():
            g = p.grad.numpy()
            if big and g.size > 60000:
                out[tag + "_grad_sub_" + k] = g.reshape(-1)[::251].copy()
                out[tag + "_grad_norm_" + k] = np.asarray(np.sqrt((g.astype(np.float64) ** 2).sum()))
            else:
                out[tag + "_grad_" + k] = g
        out[tag + "_grad_V"] = V.grad.numpy()
        out[tag + "_grad_v_g"] = v_g.grad.numpy()
        out[tag + "_grad_h0"] = h0.grad.numpy()[0]
        out[tag + "_grad_c0"] = c0.grad.numpy()[0]
        # ---- packed forward (what Encoder2Decoder.forward returns, baseline_attention.py:228) ----
        lengths = make_lengths(B, T, seed=1234)
        packed = torch.nn.utils.rnn.pack_padded_sequence(scores.detach(), lengths, batch_first=True)
        out["lengths"] = np.asarray(lengths)
        out[tag + "_packed_batch_sizes"] = packed.batch_sizes.numpy()
        pk = packed.data.numpy()
        out[tag + "_packed_data_sub"] = pk[:, ::97].copy() if big else pk
        # ---- CE loss on the packed rows (train.py:102,208) ----
        tgt = torch.nn.utils.rnn.pack_padded_sequence(cap[:, 1:], lengths, batch_first=True).data
        out["packed_targets"] = tgt.numpy()
        out[tag + "_ce_loss"] = np.asarray(torch.nn.functional.cross_entropy(packed.data, tgt).item())
        # ---- batched greedy (sampler loop body with [1,B,H] states; Q3, Q9, Q12) ----
        with torch.no_grad():
            states = (h0.detach(), c0.detach())
            tok = torch.ones(B, 1, dtype=torch.long)
            ids, att, bet, top2 = [], [], [], []
            for _ in range(L):
                s1, a1, b1, states = dec(V.detach(), v_g.detach(), tok, states)
                tok = s1.max(2)[1]
                ids.append(tok)
                att.append(a1)
                bet.append(b1)
                t2 = torch.topk(s1[:, 0], 2, dim=1).values
                top2.append((t2[:, 0] - t2[:, 1]).numpy())
            out[tag + "_greedy_ids"] = torch.cat(ids, 1).numpy()
            out[tag + "_greedy_alpha"] = torch.cat(att, 1).numpy()
            out[tag + "_greedy_beta"] = torch.cat(bet, 1).numpy()
            out[tag + "_greedy_gap"] = np.stack(top2, 1)
            # Q2/Q3: stepwise teacher-forced scores differ from batched teacher-forced for t>=1
            states = (h0.detach(), c0.detach())
            sw = []
            for t in range(T):
                s1, _, _, states = dec(V.detach(), v_g.detach(), cap[:, t:t + 1], states)
                sw.append(s1)
            sw = torch.cat(sw, 1).numpy()
            out[tag + "_stepwise_scores_sub"] = sw[:, :, ::97].copy() if big else sw
    torch.set_default_dtype(torch.float32)
    out["meta"] = np.asarray([dims.H, dims.E, dims.Vc, dims.k, B, T, L])
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


# ---- encoder heads: the reference's AttentiveCNN with the ResNet trunk replaced by nn.Identity ----
ENC_CASES = {
    "enc_tiny": (Dims(H=32, E=16, Vc=40, k=49), 3, 64),
    "enc_odd": (Dims(H=48, E=20, Vc=77, k=49), 5, 100),
    "enc_cfgA": (Dims(H=512, E=256, Vc=10000, k=49), 4, 2048),     # BASELINE config 1 shapes
}


def run_encoder_case(name, dims, B, C):
    import torchvision
    from code_src.models import baseline_attention as refb
    from adaptive_b200.synth import make_encoder_weights, make_features

    orig = torchvision.models.resnet152
    torchvision.models.resnet152 = lambda pretrained=True: orig(weights=None)   # no download (offline)
    out = {}
    try:
        for tag, tdt, ndt in (("f32", torch.float32, np.float32), ("f64", torch.float64, np.float64)):
            torch.set_default_dtype(tdt)
            enc = refb.AttentiveCNN(dims.E, dims.H, None)
            enc.resnet_conv = torch.nn.Identity()
            for nm in ("affine_a", "affine_b", "affine_h0", "affine_c0"):     # heads sized for C input channels
                old = getattr(enc, nm)
                setattr(enc, nm, torch.nn.Linear(C, old.out_features))
            w = make_encoder_weights(dims, C, seed=321, bias_scale=0.1)
            enc.load_state_dict({k: torch.from_numpy(v.astype(ndt)) for k, v in w.items()}, strict=True)
            enc = enc.to(tdt)
            A = torch.from_numpy(make_features(B, C, (7, 7), seed=4321).astype(ndt)).requires_grad_(True)
            V, v_g, (h0, c0) = enc(A)
            rng = np.random.Generator(np.random.PCG64(77))
            ups = [rng.standard_normal(t.shape).astype(ndt) for t in (V, v_g, h0, c0)]
            sum((t * torch.from_numpy(u)).sum() for t, u in zip((V, v_g, h0, c0), ups)).backward()
            big = C * dims.H > 60000
            out[tag + "_V"] = V.detach().numpy()
            out[tag + "_v_g"] = v_g.detach().numpy()
            out[tag + "_h0"] = h0.detach().numpy()[:, 0]
            out[tag + "_c0"] = c0.detach().numpy()[:, 0]
            for k, p in enc.named_parameters():
                g = p.grad.numpy()
                if big and g.size > 60000:
                    out[tag + "_grad_sub_" + k] = g.reshape(-1)[::251].copy()
                    out[tag + "_grad_norm_" + k] = np.asarray(np.sqrt((g.astype(np.float64) ** 2).sum()))
                else:
                    out[tag + "_grad_" + k] = g
            gA = A.grad.numpy()
            out[tag + "_grad_A" + ("_sub" if big else "")] = gA.reshape(-1)[::251].copy() if big else gA
    finally:
        torchvision.models.resnet152 = orig
        torch.set_default_dtype(torch.float32)
    out["meta"] = np.asarray([dims.H, dims.E, C, B, 7, 7])
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


# ---- sentinel-less baseline decoder (baseline_attention.py:66-194) ----
BASE_CASES = {
    "base_tiny": (Dims(H=32, E=16, Vc=40, k=49), 3, 5, 6),
    "base_odd": (Dims(H=48, E=20, Vc=77, k=10), 5, 7, 7),
}


def run_baseline_case(name, dims, B, T, L):
    from code_src.models import baseline_attention as refb
    from adaptive_b200.synth import baseline_weights

    out = {}
    for tag, tdt, ndt in (("f32", torch.float32, np.float32), ("f64", torch.float64, np.float64)):
        torch.set_default_dtype(tdt)
        w = baseline_weights(make_weights(dims, seed=123, dtype=np.float32, bias_scale=0.1))
        inp = make_inputs(dims, B, T, seed=1234, dtype=np.float32)
        dec = refb.Decoder(dims.E, dims.Vc, dims.H)
        dec.load_state_dict({k: torch.from_numpy(v.astype(ndt)) for k, v in w.items()}, strict=True)
        dec = dec.to(tdt)
        V = torch.from_numpy(inp["V"].astype(ndt)).requires_grad_(True)
        v_g = torch.from_numpy(inp["v_g"].astype(ndt)).requires_grad_(True)
        h0 = torch.from_numpy(inp["h0"].astype(ndt))[None].requires_grad_(True)
        c0 = torch.from_numpy(inp["c0"].astype(ndt))[None].requires_grad_(True)
        cap = torch.from_numpy(inp["captions"])
        scores, alpha, (hT, cT) = dec(V, v_g, cap, (h0, c0))
        rng = np.random.Generator(np.random.PCG64(99))
        dS = rng.standard_normal(scores.shape).astype(ndt) / scores.shape[-1]
        dA = rng.standard_normal(alpha.shape).astype(ndt) * 0.1
        ((scores * torch.from_numpy(dS)).sum() + (alpha * torch.from_numpy(dA)).sum()).backward()
        out[tag + "_scores"] = scores.detach().numpy()
        out[tag + "_alpha"] = alpha.detach().numpy()
        out[tag + "_hT"] = hT.detach().numpy()[0]
        out[tag + "_cT"] = cT.detach().numpy()[0]
        for k, p in dec.named_parameters():
            out[tag + "_grad_" + k] = p.grad.numpy()
        out[tag + "_grad_V"] = V.grad.numpy()
        out[tag + "_grad_v_g"] = v_g.grad.numpy()
        out[tag + "_grad_h0"] = h0.grad.numpy()[0]
        out[tag + "_grad_c0"] = c0.grad.numpy()[0]
        with torch.no_grad():   # batched greedy = loop body of the baseline sampler (baseline_attention.py:262-281)
            states = (h0.detach(), c0.detach())
            tok = torch.ones(B, 1, dtype=torch.long)
            ids, att = [], []
            for _ in range(L):
                s1, a1, states = dec(V.detach(), v_g.detach(), tok, states)
                tok = s1.max(2)[1]
                ids.append(tok)
                att.append(a1)
            out[tag + "_greedy_ids"] = torch.cat(ids, 1).numpy()
            out[tag + "_greedy_alpha"] = torch.cat(att, 1).numpy()
    torch.set_default_dtype(torch.float32)
    out["meta"] = np.asarray([dims.H, dims.E, dims.Vc, dims.k, B, T, L])
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    only = sys.argv[1:]   # e.g. `python oracle/gen_golden.py enc base` regenerates only those families
    want = lambda fam: not only or fam in only
    if want("dec"):
        for name, (dims, B, T, L) in CASES.items():
            run_case(name, dims, B, T, L)
    if want("enc"):
        for name, (dims, B, C) in ENC_CASES.items():
            run_encoder_case(name, dims, B, C)
    if want("base"):
        for name, (dims, B, T, L) in BASE_CASES.items():
            run_baseline_case(name, dims, B, T, L)
