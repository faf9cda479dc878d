"""Seeded synthetic inputs with the shapes and mask semantics of the reference's data pipeline
(dataset/multi_speaker_dataset.py:13-45 mixes two utterances and builds masks {0,1,2}; dataset/collate_fn.py
pads lips/text/audio with 0 and masks with 3).  Used by bench.py, smoke() and the tests; no datasets or
checkpoints are reachable offline."""
from __future__ import annotations

import numpy as np
import torch


class CharTokenizer:
    """Stand-in for utils/tokenizer.py over the 800-entry vocab: ids 0..3 = <unk>,<s>,</s>,<blank>, 4 = U+2581."""

    def __init__(self, vocab_size=800):
        self.id_to_token = ["<unk>", "<s>", "</s>", "<blank>", "▁"] + [chr(0xAC00 + i) for i in range(vocab_size - 5)]
        self.token_to_id = {t: i for i, t in enumerate(self.id_to_token)}

    vocab_size = property(lambda self: len(self.id_to_token))
    blank_id = property(lambda self: 3)
    unk_id = property(lambda self: 0)
    pad_id = property(lambda self: 0)

    def encode(self, text):
        return [self.token_to_id.get("▁" if ch == " " else ch, 0) for ch in text]

    def decode(self, ids):
        return "".join(self.id_to_token[i] for i in ids if 0 <= i < len(self.id_to_token)).replace("▁", " ").strip()


def pair_masks(n1, n2, n_max):
    """mask1/mask2 of one mixed pair, padded with 3 to n_max samples."""
    n = max(n1, n2)
    m1 = np.full(n_max, 3, dtype=np.int64); m2 = np.full(n_max, 3, dtype=np.int64)
    m1[:n] = 0; m2[:n] = 0
    both = min(n1, n2)
    m1[:both] = 1; m2[:both] = 1
    if n1 > n2:
        m1[n2:n1] = 2
    elif n2 > n1:
        m2[n1:n2] = 2
    return m1, m2


def make_batch(pairs=8, seconds=5.0, t_v=150, vocab=800, blank=3, seed=1234, l_range=(20, 58), lips=True,
               pin=False):
    """One collated batch (BASELINE config 4 defaults: 8 pairs, 5 s audio -> T_enc 249, 150 lip frames,
    utterance lengths U[2.7 s, 5 s], labels U{20..58} over [4, vocab))."""
    rng = np.random.default_rng(seed)
    g = torch.Generator().manual_seed(seed)
    n_max = int(seconds * 16000)
    lo = int(min(2.7, seconds * 0.54) * 16000)
    m1s, m2s = [], []
    for _ in range(pairs):
        n1, n2 = int(rng.integers(lo, n_max + 1)), int(rng.integers(lo, n_max + 1))
        if rng.random() < 0.5:
            n1 = n_max
        else:
            n2 = n_max
        a, b = pair_masks(n1, n2, n_max)
        m1s.append(a); m2s.append(b)
    lens = [rng.integers(l_range[0], l_range[1] + 1, size=pairs) for _ in range(2)]
    texts = []
    for s in range(2):
        t = np.zeros((pairs, int(lens[s].max())), dtype=np.int64)
        for i in range(pairs):
            t[i, :lens[s][i]] = rng.integers(4, vocab, size=lens[s][i])
        texts.append(t)
    batch = {
        "audio": 0.1 * torch.randn(pairs, n_max, generator=g),
        "audio_lengths": torch.full((pairs,), n_max),
        "mask1": torch.from_numpy(np.stack(m1s)), "mask2": torch.from_numpy(np.stack(m2s)),
        "text1": torch.from_numpy(texts[0]), "text2": torch.from_numpy(texts[1]),
        "text1_lengths": torch.from_numpy(lens[0].astype(np.int64)), "text2_lengths": torch.from_numpy(lens[1].astype(np.int64)),
    }
    if lips:
        batch["lip1"] = torch.rand(pairs, t_v, 1, 96, 96, generator=g)
        batch["lip2"] = torch.rand(pairs, t_v, 1, 96, 96, generator=g)
        batch["lip1_lengths"] = torch.full((pairs,), t_v); batch["lip2_lengths"] = torch.full((pairs,), t_v)
    if pin:
        batch = {k: v.pin_memory() for k, v in batch.items()}
    return batch


def make_features(pairs=8, t_v=150, t_enc=249, seed=1234, n_samples=80000, dtype=torch.float32):
    """Encoder-feature level inputs for the hot path (visual [B,T_v,512], audio/middle [B,T_enc,1024]) plus the
    sample-rate masks and labels of make_batch with the same seed."""
    b = make_batch(pairs=pairs, seconds=n_samples / 16000, t_v=t_v, seed=seed, lips=False)
    g = torch.Generator().manual_seed(seed + 1)
    feats = {"visual": [torch.randn(pairs, t_v, 512, generator=g).to(dtype) for _ in range(2)],
             "audio": [torch.randn(pairs, t_enc, 1024, generator=g).to(dtype) for _ in range(2)],
             "middle": [torch.randn(pairs, t_enc, 1024, generator=g).to(dtype) for _ in range(2)],
             "masks": [b["mask1"], b["mask2"]], "texts": [b["text1"], b["text2"]],
             "lens": [b["text1_lengths"], b["text2_lengths"]]}
    return feats
