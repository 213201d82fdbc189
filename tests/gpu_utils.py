"""Helpers for the -m gpu tests: move synth/oracle numpy data to the device."""
import numpy as np
import torch

from adaptive_b200._lib import KEY_TO_FIELD, WEIGHT_FIELDS


def dev_weights(w, requires_grad=False):
    """numpy weight dict (reference state_dict keys) -> tuple of CUDA tensors in aa_weights order."""
    inv = {v: k for k, v in KEY_TO_FIELD.items()}
    out = []
    for f in WEIGHT_FIELDS:
        t = torch.from_numpy(np.ascontiguousarray(w[inv[f]], dtype=np.float32)).cuda()
        if requires_grad:
            t.requires_grad_(True)
        out.append(t)
    return tuple(out)


def dev_inputs(inp, requires_grad=False):
    V = torch.from_numpy(inp["V"].astype(np.float32)).cuda()
    v_g = torch.from_numpy(inp["v_g"].astype(np.float32)).cuda()
    h0 = torch.from_numpy(inp["h0"].astype(np.float32)).cuda()
    c0 = torch.from_numpy(inp["c0"].astype(np.float32)).cuda()
    cap = torch.from_numpy(inp["captions"]).cuda()
    if requires_grad:
        for t in (V, v_g, h0, c0):
            t.requires_grad_(True)
    return V, v_g, h0, c0, cap


def grad_key_order():
    inv = {v: k for k, v in KEY_TO_FIELD.items()}
    return [inv[f] for f in WEIGHT_FIELDS]


def near_tie_report(ids_gpu, ids_ref, gap, thresh):
    """Positions where greedy ids differ although the reference top1-top2 gap is above `thresh`
    and no earlier position of that row already diverged.  Returns (hard_mismatches, near_ties)."""
    bad = ids_gpu != ids_ref
    hard, near = [], []
    for b in range(bad.shape[0]):
        pos = np.flatnonzero(bad[b])
        if pos.size == 0:
            continue
        t = int(pos[0])  # after the first flip the sequences legitimately diverge
        (near if gap[b, t] <= thresh else hard).append((b, t, float(gap[b, t])))
    return hard, near
