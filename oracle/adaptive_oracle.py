"""CPU oracle for the adaptive-attention caption decoder — TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference's decoder hot path.  It is the checker
the CUDA path is compared with; it is never shipped, never imported by ``adaptive_b200``
and never measured as the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.

Parity status
-------------
* ``decoder_forward`` / ``decoder_backward`` / ``greedy_decode`` / ``pack_padded`` are
  **pinned**: ``oracle/gen_golden.py`` runs the unmodified reference modules
  (``/root/reference/code_src/models/adaptive_attention.py``) in this container on the same
  weights and inputs and commits their outputs under ``tests/golden/``;
  ``tests/test_oracle_golden.py`` checks this file against those vectors (fp32 and fp64).
  The reference itself ships no tests or golden vectors.
* ``beam_decode`` — **parity unpinned**: the reference has no beam search at all
  (SURVEY.md Q14); the definition here is the specification.

The arithmetic that the reference delegates to PyTorch (``nn.LSTM``, ``nn.Linear``,
``F.softmax``, ``torch.bmm``; torch 2.11 in this image, the reference pins no version)
is restated from the published definitions: LSTM gate order i,f,g,o with
``c' = σ(f)c + σ(i)tanh(g)``, ``h' = σ(o)tanh(c')``.

All functions compute in the dtype of the weights they are given (float32 or float64).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

Weights = Dict[str, np.ndarray]


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _softmax(z):
    m = z.max(axis=-1, keepdims=True)
    e = np.exp(z - m)
    return e / e.sum(axis=-1, keepdims=True)


def _log_softmax(z):
    m = z.max(axis=-1, keepdims=True)
    s = z - m
    return s - np.log(np.exp(s).sum(axis=-1, keepdims=True))


# --------------------------------------------------------------------------------------
# forward
# --------------------------------------------------------------------------------------
def lstm_step(W: Weights, x_t, h, c):
    """One ``nn.LSTM`` step (``baseline_attention.py:140,172``). Returns h', c', (i,f,g,o)."""
    H = h.shape[1]
    pre = x_t @ W["LSTM.weight_ih_l0"].T + W["LSTM.bias_ih_l0"] + h @ W["LSTM.weight_hh_l0"].T + W["LSTM.bias_hh_l0"]
    i = _sigmoid(pre[:, 0 * H:1 * H])
    f = _sigmoid(pre[:, 1 * H:2 * H])
    g = np.tanh(pre[:, 2 * H:3 * H])
    o = _sigmoid(pre[:, 3 * H:4 * H])
    c2 = f * c + i * g
    h2 = o * np.tanh(c2)
    return h2, c2, (i, f, g, o)


def is_baseline(W: Weights) -> bool:
    """True for the weights of the sentinel-less baseline decoder (``baseline_attention.py:66-194``): no sentinel,
    no ``affine_s``."""
    return "adaptive.sentinel.affine_x.weight" not in W


def sentinel_forward(W: Weights, x, h_prev, cells):
    """``Sentinel.forward`` (``adaptive_attention.py:75-85``): s = σ(W_x x + W_h h̃) ⊙ tanh(c_t)."""
    ga = x @ W["adaptive.sentinel.affine_x.weight"].T + h_prev @ W["adaptive.sentinel.affine_h.weight"].T
    g = _sigmoid(ga)
    return g * np.tanh(cells), g


def atten_forward(W: Weights, V, h, s):
    """``Atten.forward`` (``adaptive_attention.py:26-58``).

    V [B,k,H]; h,s [B,T,H] -> c_hat [B,T,H], alpha [B,T,k], beta [B,T,1] and a cache.

    Without ``affine_s`` in ``W`` this is the sentinel-less ``Atten.forward`` of the baseline model
    (``baseline_attention.py:79-100``): the same scores, softmax and context, c_hat = ctx (beta = 0)."""
    Wv = W["adaptive.atten.affine_v.weight"]
    Wg = W["adaptive.atten.affine_g.weight"]
    if "adaptive.atten.affine_s.weight" not in W:
        wh = W["adaptive.atten.affine_h.weight"][0]
        P = V @ Wv.T                                                          # baseline_attention.py:88
        q = h @ Wg.T                                                          # :89
        tp = np.tanh(P[:, None, :, :] + q[:, :, None, :])
        alpha = _softmax(tp @ wh)                                             # :92-93
        ctx = alpha @ V                                                       # :96
        beta = np.zeros(alpha.shape[:2] + (1,), dtype=alpha.dtype)
        return ctx, alpha, beta, dict(P=P, q=q, tp=tp, alpha=alpha, ctx=ctx, tr=np.zeros_like(q), beta=beta)
    Ws = W["adaptive.atten.affine_s.weight"]
    wh = W["adaptive.atten.affine_h.weight"][0]
    P = V @ Wv.T                                  # [B,k,a]      :34
    q = h @ Wg.T                                  # [B,T,a]      :35
    tp = np.tanh(P[:, None, :, :] + q[:, :, None, :])   # [B,T,k,a]
    z = tp @ wh                                   # [B,T,k]      :38
    alpha = _softmax(z)                           # k-way        :39
    ctx = alpha @ V                               # bmm          :42
    r = s @ Ws.T + q                              #              :45
    tr = np.tanh(r)
    zs = tr @ wh                                  # [B,T]        :47
    ext = np.concatenate([z, zs[..., None]], axis=-1)
    beta = _softmax(ext)[..., -1:]                # (k+1)-way    :50-52
    c_hat = beta * s + (1 - beta) * ctx           #              :56
    cache = dict(P=P, q=q, tp=tp, alpha=alpha, ctx=ctx, tr=tr, beta=beta)
    return c_hat, alpha, beta, cache


def adaptive_forward(W: Weights, x, hiddens, cells, V):
    """``AdaptiveBlock.forward`` (``adaptive_attention.py:110-134``) incl. the zero-h0 shift
    for the sentinel (Q2/Q3: h̃_0 = 0; for T == 1 h̃ = 0)."""
    B, T, H = hiddens.shape
    hs_prev = np.zeros_like(hiddens)
    if T > 1:
        hs_prev[:, 1:] = hiddens[:, :-1]
    if is_baseline(W):   # sentinel-less baseline block (baseline_attention.py:121-128): scores = mlp(c + h)
        s, g = np.zeros_like(hiddens), np.zeros_like(hiddens)
    else:
        s, g = sentinel_forward(W, x, hs_prev, cells)
    c_hat, alpha, beta, cache = atten_forward(W, V, hiddens, s)
    u = c_hat + hiddens
    scores = u @ W["adaptive.mlp.weight"].T + W["adaptive.mlp.bias"]   # :132
    cache.update(hs_prev=hs_prev, s=s, g=g, u=u)
    return scores, alpha, beta, cache


def decoder_forward(W: Weights, V, v_g, captions, h0, c0, want_cache: bool = False):
    """``Decoder.forward`` (``baseline_attention.py:148-194`` with the adaptive block of
    ``adaptive_attention.py:155``).  ``h0, c0`` are [B,H].  Returns
    ``scores [B,T,Vc], alpha [B,T,k], beta [B,T,1], (hT, cT)`` (+ cache)."""
    dt = W["adaptive.mlp.weight"].dtype
    B, T = captions.shape
    H = h0.shape[1]
    emb = W["embed.weight"][captions]                                        # :151
    x = np.concatenate([emb, np.broadcast_to(v_g[:, None, :], emb.shape)], axis=2)   # :154
    hiddens = np.zeros((B, T, H), dtype=dt)
    cells = np.zeros((B, T, H), dtype=dt)
    acts = np.zeros((B, T, 4, H), dtype=dt)
    h, c = h0.astype(dt), c0.astype(dt)
    for t in range(T):                                                       # :167-178
        h, c, (i, f, g, o) = lstm_step(W, x[:, t], h, c)
        hiddens[:, t], cells[:, t] = h, c
        acts[:, t, 0], acts[:, t, 1], acts[:, t, 2], acts[:, t, 3] = i, f, g, o
    scores, alpha, beta, cache = adaptive_forward(W, x, hiddens, cells, V)
    if want_cache:
        cache.update(x=x, hiddens=hiddens, cells=cells, acts=acts, h0=h0, c0=c0, V=V, captions=captions)
        return scores, alpha, beta, (h, c), cache
    return scores, alpha, beta, (h, c)


def pack_padded(scores, lengths: Sequence[int]):
    """``pack_padded_sequence(scores, lengths, batch_first=True)``
    (``baseline_attention.py:228``): time-major concatenation of the valid rows.
    ``lengths`` must be sorted descending. Returns (data [sum(lengths), ...], batch_sizes)."""
    lengths = list(lengths)
    assert all(lengths[i] >= lengths[i + 1] for i in range(len(lengths) - 1)), "lengths must be sorted descending"
    rows, bs = [], []
    for t in range(max(lengths)):
        n = sum(1 for L in lengths if L > t)
        bs.append(n)
        rows.append(scores[:n, t])
    return np.concatenate(rows, axis=0), np.asarray(bs, dtype=np.int64)


def e2d_forward(W: Weights, V, v_g, captions, lengths, h0, c0):
    """``Encoder2Decoder.forward`` after the (out-of-scope) encoder
    (``baseline_attention.py:222-230``): decoder + pack."""
    scores = decoder_forward(W, V, v_g, captions, h0, c0)[0]
    return pack_padded(scores, lengths)


def cross_entropy(logits, targets):
    """``nn.CrossEntropyLoss()`` (mean) as used at ``train.py:63,208``; returns loss, dlogits."""
    ls = _log_softmax(logits)
    n = logits.shape[0]
    loss = -ls[np.arange(n), targets].mean()
    d = np.exp(ls)
    d[np.arange(n), targets] -= 1.0
    return loss, d / n


def packed_targets(captions, lengths):
    """``pack_padded_sequence(captions[:, 1:], lengths)`` (``train.py:102``)."""
    return pack_padded(captions[:, 1:], lengths)[0]


# --------------------------------------------------------------------------------------
# backward (what autograd computes for the reference; formulas in SURVEY.md §7)
# --------------------------------------------------------------------------------------
def decoder_backward(W: Weights, cache, d_scores, d_alpha=None, d_beta=None, d_hT=None, d_cT=None):
    """Gradients of ``sum(d_scores*scores) [+ d_alpha.alpha + d_beta.beta + d_hT.hT + d_cT.cT]``
    w.r.t. the 13 decoder parameters and ``V, v_g, h0, c0``."""
    x, hiddens, cells, acts = cache["x"], cache["hiddens"], cache["cells"], cache["acts"]
    V, captions = cache["V"], cache["captions"]
    B, T, H = hiddens.shape
    E = W["embed.weight"].shape[1]
    Wv = W["adaptive.atten.affine_v.weight"]
    Wg = W["adaptive.atten.affine_g.weight"]
    base = is_baseline(W)   # beta = 0: every sentinel term below vanishes
    Ws = None if base else W["adaptive.atten.affine_s.weight"]
    wh = W["adaptive.atten.affine_h.weight"][0]
    Wp = W["adaptive.mlp.weight"]
    Wx = None if base else W["adaptive.sentinel.affine_x.weight"]
    Wh = None if base else W["adaptive.sentinel.affine_h.weight"]
    Wih, Whh = W["LSTM.weight_ih_l0"], W["LSTM.weight_hh_l0"]
    alpha, beta, ctx, tp, tr = cache["alpha"], cache["beta"], cache["ctx"], cache["tp"], cache["tr"]
    s, g, u, hs_prev = cache["s"], cache["g"], cache["u"], cache["hs_prev"]
    G: Dict[str, np.ndarray] = {}

    # vocabulary projection  (adaptive_attention.py:132)
    dS2 = d_scores.reshape(B * T, -1)
    du = d_scores @ Wp
    G["adaptive.mlp.weight"] = dS2.T @ u.reshape(B * T, H)
    G["adaptive.mlp.bias"] = dS2.sum(0)
    dh = du.copy()
    dchat = du
    # beta gate (:56)
    dbeta = (dchat * (s - ctx)).sum(-1, keepdims=True)
    if d_beta is not None:
        dbeta = dbeta + d_beta
    ds = beta * dchat
    dctx = (1 - beta) * dchat
    # context bmm (:42)
    dalpha = dctx @ V.transpose(0, 2, 1)
    if d_alpha is not None:
        dalpha = dalpha + d_alpha
    dV = alpha.transpose(0, 2, 1) @ dctx
    # two softmaxes (:39, :51)
    b1 = beta * (1 - beta) * dbeta                                  # [B,T,1]
    dz = alpha * (dalpha - (alpha * dalpha).sum(-1, keepdims=True)) - b1 * alpha
    dzs = b1[..., 0]
    # scores (:34-38, :45-47)
    dp = dz[..., None] * wh * (1 - tp * tp)                         # [B,T,k,a]
    dr = dzs[..., None] * wh * (1 - tr * tr)                        # [B,T,a]
    G["adaptive.atten.affine_h.weight"] = ((dz[..., None] * tp).sum((0, 1, 2)) + (dzs[..., None] * tr).sum((0, 1)))[None, :]
    dP = dp.sum(1)                                                  # [B,k,a]
    dq = dp.sum(2) + dr                                             # [B,T,a]
    if not base:
        ds = ds + dr @ Ws
        G["adaptive.atten.affine_s.weight"] = dr.reshape(B * T, -1).T @ s.reshape(B * T, H)
    dh += dq @ Wg
    G["adaptive.atten.affine_g.weight"] = dq.reshape(B * T, -1).T @ hiddens.reshape(B * T, H)
    dV = dV + dP @ Wv
    G["adaptive.atten.affine_v.weight"] = np.einsum("bka,bkh->ah", dP, V)
    # sentinel (:79-83)
    tc = np.tanh(cells)
    if base:
        dcell = np.zeros_like(cells)
        dx = np.zeros_like(x)
    else:
        dg = ds * tc
        dcell = ds * g * (1 - tc * tc)
        da = dg * g * (1 - g)
        dx = da @ Wx
        G["adaptive.sentinel.affine_x.weight"] = da.reshape(B * T, H).T @ x.reshape(B * T, -1)
        G["adaptive.sentinel.affine_h.weight"] = da.reshape(B * T, H).T @ hs_prev.reshape(B * T, H)
        if T > 1:
            dh[:, :-1] += (da @ Wh)[:, 1:]                          # h̃_t = h_{t-1}, t >= 1 (Q2)
    # LSTM BPTT (baseline_attention.py:167-178)
    dWih = np.zeros_like(Wih)
    dWhh = np.zeros_like(Whh)
    db = np.zeros(4 * H, dtype=Wih.dtype)
    dh_next = np.zeros((B, H), dtype=Wih.dtype) if d_hT is None else d_hT.copy()
    dc_next = np.zeros((B, H), dtype=Wih.dtype) if d_cT is None else d_cT.copy()
    for t in range(T - 1, -1, -1):
        i, f, gg, o = acts[:, t, 0], acts[:, t, 1], acts[:, t, 2], acts[:, t, 3]
        c_prev = cells[:, t - 1] if t > 0 else cache["c0"]
        h_prev = hiddens[:, t - 1] if t > 0 else cache["h0"]
        dh_t = dh[:, t] + dh_next
        dc_t = dcell[:, t] + dc_next + dh_t * o * (1 - tc[:, t] ** 2)
        dgates = np.concatenate([
            dc_t * gg * i * (1 - i),
            dc_t * c_prev * f * (1 - f),
            dc_t * i * (1 - gg * gg),
            dh_t * tc[:, t] * o * (1 - o),
        ], axis=1)
        dc_next = dc_t * f
        dh_next = dgates @ Whh
        dx[:, t] += dgates @ Wih
        dWih += dgates.T @ x[:, t]
        dWhh += dgates.T @ h_prev
        db += dgates.sum(0)
    G["LSTM.weight_ih_l0"], G["LSTM.weight_hh_l0"] = dWih, dWhh
    G["LSTM.bias_ih_l0"], G["LSTM.bias_hh_l0"] = db, db.copy()
    # x = [embed(w); v_g]  (:151-154)
    dE = np.zeros_like(W["embed.weight"])
    np.add.at(dE, captions.reshape(-1), dx[:, :, :E].reshape(B * T, E))
    G["embed.weight"] = dE
    G["V"], G["v_g"], G["h0"], G["c0"] = dV, dx[:, :, E:].sum(1), dh_next, dc_next
    return G


# --------------------------------------------------------------------------------------
# decoding
# --------------------------------------------------------------------------------------
def decode_step(W: Weights, V, v_g, tokens, h, c):
    """One sampler step = ``Decoder.forward`` with seq-len 1 (``adaptive_attention.py:198``):
    the sentinel sees h̃ = 0 (Q3).  tokens [B] -> scores [B,Vc], alpha [B,k], beta [B], h', c'."""
    scores, alpha, beta, (h2, c2) = decoder_forward(W, V, v_g, tokens[:, None], h, c)
    return scores[:, 0], alpha[:, 0], beta[:, 0, 0], h2, c2


def greedy_decode(W: Weights, V, v_g, h0, c0, max_len: int = 20, want_scores: bool = False):
    """Body of ``Encoder2Decoder.sampler`` (``adaptive_attention.py:186-216``) with correctly
    shaped states (Q9): start id 1, arg-max of raw logits (lowest index wins ties, Q12),
    never stops early.  Returns ids [B,L] int64, attention [B,L,k], Beta [B,L,1]."""
    B = V.shape[0]
    tok = np.ones(B, dtype=np.int64)
    h, c = h0, c0
    ids, att, bet, sc = [], [], [], []
    for _ in range(max_len):
        scores, alpha, beta, h, c = decode_step(W, V, v_g, tok, h, c)
        tok = scores.argmax(axis=1).astype(np.int64)
        ids.append(tok)
        att.append(alpha)
        bet.append(beta)
        if want_scores:
            sc.append(scores)
    out = (np.stack(ids, 1), np.stack(att, 1), np.stack(bet, 1)[..., None])
    if want_scores:
        out = out + (np.stack(sc, 1),)
    return out


END_ID = 2  # build_vocab.py:48-51


def beam_decode(W: Weights, V, v_g, h0, c0, beam: int = 3, max_len: int = 20):
    """Beam search over the reference's single-step decoder (definition: SURVEY.md §8c;
    PARITY UNPINNED — no beam search exists in the reference).

    Per image: hypotheses ranked by cumulative log-prob (log_softmax of the step scores),
    candidates = live beams x vocab, ties -> lowest flat index (beam-major); a hypothesis
    that emitted ``<end>`` is frozen (it competes with its score, emits ``<end>`` padding and
    keeps its state); no length normalisation; first step expands the single ``<start>``
    hypothesis.  Returns the top hypothesis: ids [B,L], alpha [B,L,k], beta [B,L,1],
    score [B]."""
    B, k = V.shape[0], V.shape[1]
    dt = V.dtype
    Vc = W["adaptive.mlp.weight"].shape[0]
    NEG = -np.inf
    # flat layout [B*beam]
    rep = lambda a: np.repeat(a, beam, axis=0)
    Vr, vgr = rep(V), rep(v_g)
    h, c = rep(h0), rep(c0)
    tok = np.ones(B * beam, dtype=np.int64)
    cum = np.zeros((B, beam), dtype=dt)
    cum[:, 1:] = NEG                      # only hypothesis 0 is live at step 0
    done = np.zeros((B, beam), dtype=bool)
    ids = np.zeros((B, beam, max_len), dtype=np.int64)
    att = np.zeros((B, beam, max_len, k), dtype=dt)
    bet = np.zeros((B, beam, max_len), dtype=dt)
    ar = np.arange(B)[:, None]
    for t in range(max_len):
        scores, alpha, beta, h2, c2 = decode_step(W, Vr, vgr, tok, h, c)
        lp = _log_softmax(scores).reshape(B, beam, Vc)
        cand = cum[:, :, None] + lp
        # frozen hypotheses: single candidate (<end>) carrying the unchanged score
        frozen = np.full((B, beam, Vc), NEG, dtype=dt)
        frozen[:, :, END_ID] = cum
        cand = np.where(done[:, :, None], frozen, cand)
        flat = cand.reshape(B, beam * Vc)
        # top-`beam`, ties -> lowest flat index: stable sort on (-value)
        order = np.argsort(-flat, axis=1, kind="stable")[:, :beam]
        src = order // Vc
        word = order % Vc
        cum = np.take_along_axis(flat, order, axis=1)
        was_done = np.take_along_axis(done, src, axis=1)
        # reorder histories and states by back-pointer
        ids = ids[ar, src]
        att = att[ar, src]
        bet = bet[ar, src]
        ids[:, :, t] = word
        a3 = alpha.reshape(B, beam, k)[ar, src]
        b3 = beta.reshape(B, beam)[ar, src]
        att[:, :, t] = np.where(was_done[..., None], 0, a3)
        bet[:, :, t] = np.where(was_done, 0, b3)
        H = h.shape[1]
        hn = h2.reshape(B, beam, H)[ar, src]
        cn = c2.reshape(B, beam, H)[ar, src]
        ho = h.reshape(B, beam, H)[ar, src]
        co = c.reshape(B, beam, H)[ar, src]
        h = np.where(was_done[..., None], ho, hn).reshape(B * beam, H)
        c = np.where(was_done[..., None], co, cn).reshape(B * beam, H)
        done = was_done | (word == END_ID)
        tok = word.reshape(-1)
    return ids[:, 0], att[:, 0], bet[:, 0][..., None], cum[:, 0]


# --------------------------------------------------------------------------------------
# encoder heads -- pinned by tests/golden/enc_*.npz (gen_golden.py runs the
# reference's AttentiveCNN with its ResNet trunk replaced by nn.Identity)
# --------------------------------------------------------------------------------------
def encoder_forward(We: Weights, A, want_cache: bool = False):
    """``AttentiveCNN.forward`` after the trunk (``baseline_attention.py:46-62``): A [B,C,h,w] last-conv
    feature map -> V [B,h*w,H], v_g [B,E], h0 [B,H], c0 [B,H].  The average pool covers the whole map
    (``AvgPool2d(7)`` on the reference's 7x7 maps, ``:46-47``)."""
    B, C = A.shape[0], A.shape[1]
    At = A.reshape(B, C, -1).transpose(0, 2, 1)                                   # :50
    a_g = At.mean(axis=1)                                                         # :46-47
    V = np.maximum(At @ We["affine_a.weight"].T + We["affine_a.bias"], 0)         # :51
    v_g = np.maximum(a_g @ We["affine_b.weight"].T + We["affine_b.bias"], 0)      # :53
    h0 = np.tanh(a_g @ We["affine_h0.weight"].T + We["affine_h0.bias"])           # :56
    c0 = np.tanh(a_g @ We["affine_c0.weight"].T + We["affine_c0.bias"])           # :58
    if want_cache:
        return V, v_g, h0, c0, dict(At=At, a_g=a_g, V=V, v_g=v_g, h0=h0, c0=c0, shape=A.shape)
    return V, v_g, h0, c0


def encoder_backward(We: Weights, cache, dV, dv_g, dh0, dc0):
    """Gradients of ``sum(dV*V) + sum(dv_g*v_g) + sum(dh0*h0) + sum(dc0*c0)`` w.r.t. the 8 head parameters and A."""
    At, a_g = cache["At"], cache["a_g"]
    B, hw, C = At.shape
    G: Dict[str, np.ndarray] = {}
    dVp = dV * (cache["V"] > 0)
    G["affine_a.weight"] = dVp.reshape(B * hw, -1).T @ At.reshape(B * hw, C)
    G["affine_a.bias"] = dVp.sum((0, 1))
    dAt = dVp @ We["affine_a.weight"]
    dag = np.zeros_like(a_g)
    for name, up, act in (("affine_b", dv_g, "relu"), ("affine_h0", dh0, "tanh"), ("affine_c0", dc0, "tanh")):
        out = cache[{"affine_b": "v_g", "affine_h0": "h0", "affine_c0": "c0"}[name]]
        dp = up * (out > 0) if act == "relu" else up * (1 - out * out)
        G[name + ".weight"] = dp.T @ a_g
        G[name + ".bias"] = dp.sum(0)
        dag = dag + dp @ We[name + ".weight"]
    dAt = dAt + dag[:, None, :] / hw
    G["A"] = dAt.transpose(0, 2, 1).reshape(cache["shape"])
    return G
