"""Dropout randomness of the text tower in train mode (roberta.py:162,244,337,418: hidden_dropout_prob =
attention_probs_dropout_prob = 0.1).

The kernels draw keep masks from Philox4x32-10 keyed by (step seed, site id) -- csrc/philox.cuh.  The step seed is ONE int64
in device memory per device: `advance()` bumps it with a kernel, so a captured CUDA graph that contains the call draws new
masks on every replay; the site id names the call site: (text-tower invocation of this step) << 16 | layer slot << 4 |
kind (functional.drop_site).  Nothing is stored for the backward: it regenerates the masks from the same pair."""
import types

import torch

from . import lib as _lib

_seeds = {}
_pass = 0


def _norm(device):
    device = torch.device(device)
    if device.type == "cuda" and device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


def seed_tensor(device):
    """the device-resident step seed (created from torch's global seed on first use)"""
    device = _norm(device)
    key = (device.type, device.index)
    t = _seeds.get(key)
    if t is None:
        t = _seeds[key] = torch.tensor([torch.initial_seed() & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device=device)
    return t


def manual_seed(seed, device=None):
    """set the step seed of `device` (default: of every device seen so far)"""
    if device is not None:
        seed_tensor(device).fill_(int(seed) & 0x7FFFFFFFFFFFFFFF)
        return
    for t in _seeds.values():
        t.fill_(int(seed) & 0x7FFFFFFFFFFFFFFF)


def advance(device):
    """start of a training step: new step seed (a kernel launch: captured with the step), site numbering restarts"""
    global _pass
    _pass = 0
    _lib.kernels().rng_advance(seed_tensor(device))


def begin_pass():
    """a new text-tower invocation (RobertaEmbeddings.forward): its dropout sites get a fresh id range"""
    global _pass
    _pass += 1
    return _pass << 16


def current_base():
    return _pass << 16


def drop_cfg(p_hidden, p_attn, device, base=None):
    """the `drop` argument of functional.text_layer_fwd / text_embeddings_fwd"""
    return types.SimpleNamespace(p=float(p_hidden), p_attn=float(p_attn), seed=seed_tensor(device),
                                 base=current_base() if base is None else base)
