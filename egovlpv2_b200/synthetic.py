"""Seeded synthetic inputs of SURVEY.md section 8(d): N(0,1) video (post-normalisation statistics), RoBERTa-style token
ids with <s>/</s>/pad structure, 15 % MLM masking (80/10/10), multi-hot noun / verb vectors."""
import torch


def synthetic_batch(B, T, img, S, seed=1234, vocab=50265, n_noun=582, n_verb=118, pin=False):
    g = torch.Generator().manual_seed(seed)
    video = torch.randn(B, T, 3, img, img, generator=g)
    lens = torch.randint(min(8, S), S + 1, (B,), generator=g)
    ids = torch.randint(3, vocab - 1, (B, S), generator=g)
    pos = torch.arange(S)[None]
    ids[:, 0] = 0
    ids = torch.where(pos == (lens[:, None] - 1), torch.full_like(ids, 2), ids)
    ids = torch.where(pos >= lens[:, None], torch.full_like(ids, 1), ids)
    am = (pos < lens[:, None]).to(torch.int64)
    pick = (torch.rand(B, S, generator=g) < 0.15) & ~(ids <= 2)
    for b in range(B):
        if not pick[b].any():
            pick[b, 1] = True
    labels = torch.where(pick, ids, torch.full_like(ids, -100))
    r = torch.rand(B, S, generator=g)
    rnd = torch.randint(3, vocab - 1, (B, S), generator=g)
    mlm_ids = ids.clone()
    mlm_ids[pick & (r < 0.8)] = vocab - 1
    sel = pick & (r >= 0.8) & (r < 0.9)
    mlm_ids[sel] = rnd[sel]
    noun = (torch.rand(B, n_noun, generator=g) < 0.01).float()
    verb = (torch.rand(B, n_verb, generator=g) < 0.02).float()
    noun[torch.arange(B), torch.randint(0, n_noun, (B,), generator=g)] = 1.0
    verb[torch.arange(B), torch.randint(0, n_verb, (B,), generator=g)] = 1.0
    d = dict(video=video, input_ids=ids, attention_mask=am, text_mlm_ids=mlm_ids, text_mlm_labels=labels, noun_vec=noun,
             verb_vec=verb)
    if pin and torch.cuda.is_available():
        d = {k: v.pin_memory() for k, v in d.items()}
    return d
