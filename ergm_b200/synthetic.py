"""Synthetic MELD-shaped batches — shared by tests and bench.py.

Layout follows the reference collate (custom_dataset.py:47-68,107-132): flattened alternating
speaker utterances, token_type_ids alternating <sp1>/<sp2>, labels = -100 over the history
and the response tokens + eos as targets, right-padding with eos / -100; emotion label in
[0,7); caption ids [B,Tc]; audio [B,Ta,768] and visual [B,197*Kf,768] feature sequences whose
time-means are the [B,768] / [B,1,768] tensors model.py:497-498 adds.
"""
import torch

VOCAB = 50260
EOS, BOS, SP1, SP2 = 50256, 50257, 50258, 50259


def make_batch(B, T, seed=1234, n_embd=768, ta=113, kf=1, ragged=True, feat_dim=768, tc=None,
               vocab=VOCAB):
    """vocab < 50260 builds the same layout for tiny test models: the four special ids sit
    at the top of the vocabulary (eos = vocab-4, bos, sp1, sp2 = vocab-1)."""
    g = torch.Generator().manual_seed(seed)
    EOS, BOS, SP1, SP2 = vocab - 4, vocab - 3, vocab - 2, vocab - 1
    tc = T if tc is None else tc
    ids = torch.full((B, T), EOS, dtype=torch.long)
    tt = torch.full((B, T), EOS, dtype=torch.long)
    lab = torch.full((B, T), -100, dtype=torch.long)
    for b in range(B):
        n = int(torch.randint(T // 2, T + 1, (1,), generator=g)) if ragged else T
        resp = min(n - 1, int(torch.randint(16, 33, (1,), generator=g)))
        toks = torch.randint(0, vocab - 3, (n,), generator=g)
        toks[0] = BOS
        toks[n - 1] = EOS
        ids[b, :n] = toks
        pos, spk = 0, 0
        while pos < n - resp:
            seg = int(torch.randint(8, 25, (1,), generator=g))
            tt[b, pos:min(pos + seg, n - resp)] = SP1 if spk == 0 else SP2
            pos += seg
            spk ^= 1
        tt[b, n - resp:n] = SP2
        lab[b, n - resp:n] = ids[b, n - resp:n]
    emo = torch.randint(0, 7, (B,), generator=g)
    cap = torch.randint(0, vocab - 3, (B, tc), generator=g)
    aud_seq = torch.randn(B, ta, feat_dim, generator=g)
    vis_seq = torch.randn(B, 197 * kf, feat_dim, generator=g)
    return dict(input_ids=ids, token_type_ids=tt, labels=lab, emotion_labels=emo, caption_ids=cap,
                aud_seq=aud_seq, vis_seq=vis_seq,
                auds=aud_seq.mean(1), imgs=vis_seq.mean(1, keepdim=True))


def gv1_inputs():
    """The exact GV-1 generator of SURVEY.md §8(c)."""
    B, T = 4, 128
    g = torch.Generator().manual_seed(1234)
    ids = torch.randint(0, 50257, (B, T), generator=g)
    cap = torch.randint(0, 50257, (B, T), generator=g)
    tt = torch.where(torch.arange(T)[None].expand(B, T) % 32 < 16, 50258, 50259)
    lab = ids.clone()
    lab[:, :T // 2] = -100
    emo = torch.randint(0, 7, (B,), generator=g)
    imgs = torch.randn(B, 1, 768, generator=g)
    auds = torch.randn(B, 768, generator=g)
    return dict(input_ids=ids, caption_ids=cap, token_type_ids=tt, labels=lab, emotion_labels=emo,
                imgs=imgs, auds=auds)
