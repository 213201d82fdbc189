"""Drop-in ``nn.Module`` surface of the reference's sentinel-less baseline captioner
(``code_src/models/baseline_attention.py:66-283``: spatial attention without the visual sentinel).  SURVEY.md §8f rank 4.

Same class names, constructor arguments, return tuples (no beta anywhere), attribute paths and ``state_dict`` keys as the
reference.  The arithmetic is the adaptive operator of ``libadaptive_sm100.so`` with the three sentinel weights absent
(``sen_wx = sen_wh = att_ws = NULL``): the kernels force beta = 0, so c_hat = ctx and ``scores = mlp(ctx + h)``; the
sentinel contractions, their gradients and the r = s W_s^T + q contraction are skipped on the host side.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import functional as F_aa
from .modules import ATT_DIM, AttentiveCNN, _kaiming_normal, _lstm_init


def _xavier_normal(nonlinearity, *modules):
    """model_utils.xavier_normal (model_utils.py:19-31)."""
    gain = nn.init.calculate_gain(nonlinearity)
    for m in modules:
        nn.init.xavier_normal_(m.weight, gain)
        if m.bias is not None:
            m.bias.data.fill_(0)


class Atten(nn.Module):
    """baseline_attention.py:66-100."""

    def __init__(self, hidden_size):
        super().__init__()
        self.affine_v = nn.Linear(hidden_size, ATT_DIM, bias=False)
        self.affine_g = nn.Linear(hidden_size, ATT_DIM, bias=False)
        self.affine_h = nn.Linear(ATT_DIM, 1, bias=False)
        self.dropout = nn.Dropout(0)
        _xavier_normal("tanh", self.affine_v, self.affine_g)
        _kaiming_normal("relu", 0, self.affine_h)

    def forward(self, V, h_t):
        """-> c_t [B,T,H], alpha_t [B,T,k] (forward only; training runs through ``Decoder.forward``)."""
        c_t, alpha, _ = F_aa.atten_forward(self.affine_v.weight, self.affine_g.weight, None, self.affine_h.weight, V, h_t, None)
        return c_t, alpha


class AdaptiveBlock(nn.Module):
    """baseline_attention.py:104-128 (the reference keeps the name although the block has no sentinel)."""

    def __init__(self, hidden_size, vocab_size):
        super().__init__()
        self.atten = Atten(hidden_size)
        self.mlp = nn.Linear(hidden_size, vocab_size)
        self.dropout = nn.Dropout(0)
        _kaiming_normal("relu", 0, self.mlp)

    def _weights13(self, embed=None, lstm=None):
        z = self.mlp.weight.new_empty(0)
        e = embed.weight if embed is not None else z
        l = [lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0] if lstm is not None else [z] * 4
        return (e, *l, None, None, self.atten.affine_v.weight, self.atten.affine_g.weight, None, self.atten.affine_h.weight,
                self.mlp.weight, self.mlp.bias)

    def forward(self, x, hiddens, cells, V):
        """-> scores [B,T,Vc], atten_weights [B,T,k] (forward only)."""
        scores, alpha, _ = F_aa.adaptive_forward(self._weights13(), x, hiddens, cells, V)
        return scores, alpha


class Decoder(nn.Module):
    """baseline_attention.py:132-194."""

    def __init__(self, embed_size, vocab_size, hidden_size, cf=None):
        super().__init__()
        self.embed = nn.Embedding(vocab_size, embed_size)
        self.LSTM = nn.LSTM(embed_size * 2, hidden_size, 1, batch_first=True)   # parameter container only
        self.adaptive = AdaptiveBlock(hidden_size, vocab_size)
        _lstm_init(self.LSTM)
        self.precision = getattr(cf, "precision", "fp32") if cf is not None else "fp32"
        self.decode_precision = getattr(cf, "decode_precision", "tf32x3") if cf is not None else "tf32x3"

    def weights(self):
        return self.adaptive._weights13(self.embed, self.LSTM)

    def forward(self, V, v_g, captions, states=None):
        """-> scores [B,T,Vc], atten_weights [B,T,k], (h_n, c_n) each [1,B,H]."""
        h0, c0 = states if states is not None else (None, None)
        scores, alpha, _, hT, cT = F_aa.decoder_forward(self.weights(), V, v_g, captions, h0, c0, self.precision)
        return scores, alpha, (hT.unsqueeze(0), cT.unsqueeze(0))


class _Cfg:
    base_word_embed_size = 256
    base_lstm_hidden_size = 512
    vocab_length = 10000


class Encoder2Decoder(nn.Module):
    """baseline_attention.py:198-283.  ``images``: feature maps ``[B,2048,h,w]`` or an encoded tuple ``(V, v_g, (h0, c0))``."""

    def __init__(self, cf=None):
        super().__init__()
        cf = cf if cf is not None else _Cfg()
        self.encoder = AttentiveCNN(cf.base_word_embed_size, cf.base_lstm_hidden_size, cf)
        self.decoder = Decoder(cf.base_word_embed_size, cf.vocab_length, cf.base_lstm_hidden_size, cf)

    def _encode(self, images):
        if isinstance(images, (tuple, list)):
            return images
        return self.encoder(images)

    def forward(self, images, captions, lengths):
        """-> PackedSequence of scores (baseline_attention.py:206-230); the projection runs over the kept rows only."""
        V, v_g, states = self._encode(images)
        h0, c0 = states if states is not None else (None, None)
        if V.shape[2] % 4 == 0:
            return F_aa.decoder_forward_packed(self.decoder.weights(), V, v_g, captions, lengths, h0, c0, self.decoder.precision)[0]
        return F_aa.pack_scores(self.decoder(V, v_g, captions, states)[0], lengths)

    def forward_loss(self, images, captions, lengths, targets=None):
        """``criterion(self(images, captions, lengths).data, targets)`` (``train.py:205-208``) as one operator; same routes as the
        adaptive model's ``forward_loss`` (the fused operator takes the sentinel-less weights like every other entry point)."""
        V, v_g, states = self._encode(images)
        h0, c0 = states if states is not None else (None, None)
        if (self.decoder.precision == "bf16" and V.shape[2] % 8 == 0
                and F_aa.fused_loss_pays(sum(int(x) for x in lengths), self.decoder.embed.num_embeddings)):
            return F_aa.decoder_forward_loss(self.decoder.weights(), V, v_g, captions, lengths, targets, h0, c0)[0]
        packed = self.forward((V, v_g, states), captions, lengths)
        if targets is None:
            targets = F_aa.packed_targets(captions, lengths)
        return F_aa.cross_entropy(packed.data, targets)

    def sampler(self, images, max_len=30):
        """Greedy search (baseline_attention.py:233-283) -> sampled_ids [B,max_len], attention [B,max_len,k]."""
        V, v_g, states = self._encode(images)
        h0, c0 = states if states is not None else (None, None)
        ids, att, _ = F_aa.greedy_decode(self.decoder.weights(), V, v_g, h0, c0, max_len, precision=self.decoder.decode_precision)
        return ids, att
