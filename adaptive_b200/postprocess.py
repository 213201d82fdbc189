"""Eval post-processing next to the sampler: sampled ids -> caption strings cut at ``<end>`` and the
result records the COCO caption metrics consume (``tools/utils.py:180-195,225``), plus the reference's vocabulary wrapper
format (``data/build_vocab.py:9-56``: ``<pad>``=0, ``<start>``=1, ``<end>``=2, ``<unk>``=3, then the words)."""
from __future__ import annotations

import json
from typing import Dict, Iterable, List, Sequence

SPECIALS = ("<pad>", "<start>", "<end>", "<unk>")


class Vocabulary:
    """``build_vocab.Vocabulary`` (same attributes, so a pickled reference vocabulary can be swapped in)."""

    def __init__(self, words: Iterable[str] = ()):
        self.word2idx: Dict[str, int] = {}
        self.idx2word: Dict[int, str] = {}
        self.idx = 0
        for w in SPECIALS:
            self.add_word(w)
        for w in words:
            self.add_word(w)

    def add_word(self, word: str):
        if word not in self.word2idx:
            self.word2idx[word] = self.idx
            self.idx2word[self.idx] = word
            self.idx += 1

    def __call__(self, word: str) -> int:
        return self.word2idx.get(word, self.word2idx["<unk>"])

    def __len__(self) -> int:
        return len(self.word2idx)


def ids_to_captions(sampled_ids, vocab) -> List[str]:
    """``tools/utils.py:180-191``: words up to (not including) the first ``<end>``, joined by spaces.
    ``sampled_ids``: [B, L] tensor / array / nested list of int ids (as returned by ``Encoder2Decoder.sampler``)."""
    if hasattr(sampled_ids, "detach"):
        sampled_ids = sampled_ids.detach().cpu().numpy()
    out = []
    for row in sampled_ids:
        words = []
        for word_id in row:
            word = vocab.idx2word[int(word_id)]
            if word == "<end>":
                break
            words.append(word)
        out.append(" ".join(words))
    return out


def caption_results(img_ids: Sequence[int], sampled_ids, vocab) -> List[dict]:
    """``[{'image_id': int, 'caption': str}, ...]`` -- the records of ``tools/utils.py:193-195``."""
    caps = ids_to_captions(sampled_ids, vocab)
    if len(caps) != len(img_ids):
        raise ValueError("%d captions for %d image ids" % (len(caps), len(img_ids)))
    return [{"image_id": int(i), "caption": c} for i, c in zip(img_ids, caps)]


def dump_results(results: List[dict], path: str):
    """``json.dump(results, open(resFile, 'w'))`` (``tools/utils.py:225``)."""
    with open(path, "w") as f:
        json.dump(results, f)
