"""TEST INFRASTRUCTURE ONLY -- loads the UNMODIFIED reference from /root/reference.

This module exists so that `oracle/make_golden.py` can execute the reference's
own PyTorch code (EgoVLPv2/model/{model,video_transformer,roberta,heads,loss}.py)
in this container and dump golden input/output vectors into tests/golden/.
It is never imported by the product package, by `bench.py`, by `smoke()` or by
any `-m gpu` test: /root/reference does not exist on the GPU box.

The reference pins torch 1.13 / transformers 4.30 / timm 0.4.12; this image has
torch 2.11 / transformers 5.5 / no timm, so a handful of names are stubbed
before the reference modules are imported.  None of the
stubs touches arithmetic on the path: DropPath is the identity at rate 0 (the
only rate pre-training uses), trunc_normal_ only affects init (we overwrite all
weights with a seeded state dict afterwards).
"""
import contextlib
import functools
import importlib
import os
import sys
import types

import torch

REF_ROOT = "/root/reference/EgoVLPv2"

_loaded = {}


def _stub_module(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def available():
    return os.path.isdir(REF_ROOT)


def load(use_checkpoint=False, root=None):
    """Import the reference model modules; returns a namespace with mm/vt/rb/heads/loss.
    `root`: another directory holding the same unmodified files (tools/install_ref.py stages baseline/_ref/EgoVLPv2 for
    the GPU-eager baseline of tools/bench_ref_gpu.py, the only caller that passes it)."""
    global REF_ROOT
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if root is not None:
        REF_ROOT = root
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError("reference tree not present (expected on the build container only)")

    import transformers  # noqa: F401  (must precede the timm stub)
    import transformers.modeling_utils as mu
    from transformers import RobertaConfig
    from transformers.modeling_utils import PreTrainedModel

    if not hasattr(mu, "find_pruneable_heads_and_indices"):
        mu.find_pruneable_heads_and_indices = None
    if not hasattr(mu, "prune_linear_layer"):
        mu.prune_linear_layer = None

    class _DropPath(torch.nn.Module):
        def __init__(self, p=0.0):
            super().__init__()
            assert p == 0.0, "shim only models the pre-training setting (rate 0)"

        def forward(self, x):
            return x

    def _to_2tuple(x):
        return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

    def _trunc_normal_(t, std=1.0, **_):
        return torch.nn.init.trunc_normal_(t, std=std)

    layers = _stub_module("timm.models.layers", DropPath=_DropPath, to_2tuple=_to_2tuple,
                          trunc_normal_=_trunc_normal_)
    models = _stub_module("timm.models", layers=layers)
    _stub_module("timm", models=models)
    _stub_module("av")
    _stub_module("ffmpeg")
    _stub_module("humanize")
    bridge = types.SimpleNamespace(set_bridge=lambda *_a, **_k: None)
    _stub_module("decord", bridge=bridge, VideoReader=None, cpu=None)
    try:
        import cv2  # noqa: F401
    except Exception:
        _stub_module("cv2")

    cwd = os.getcwd()
    os.chdir(REF_ROOT)
    sys.path.insert(0, REF_ROOT)
    # the repo under test also ships a package named `model`; make sure the
    # reference's wins inside this process
    for k in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
        del sys.modules[k]
    real_load = torch.load

    def _fake_load(path, *a, **k):
        if isinstance(path, str) and "jx_vit" in path:
            return {"shim_unused_key": torch.zeros(1)}
        return real_load(path, *a, **k)

    try:
        torch.load = _fake_load
        rb = importlib.import_module("model.roberta")

        def _init_weights(self):
            if not getattr(self, "_shim_inited", False):
                self._shim_inited = True
                self.post_init()

        rb.RobertaModel.init_weights = _init_weights
        rb.RobertaModel.get_head_mask = lambda self, hm, n, *a, **k: [None] * n
        rb.RobertaModel.get_extended_attention_mask = (
            lambda self, am, shp, device=None, dtype=None:
            PreTrainedModel.get_extended_attention_mask(self, am, shp, torch.float32))
        rb._shim_text_config = dict(vocab_size=50265, hidden_size=768, num_hidden_layers=12,
                                    num_attention_heads=12, intermediate_size=3072,
                                    max_position_embeddings=514, type_vocab_size=1,
                                    pad_token_id=1, layer_norm_eps=1e-5)
        rb.RobertaModel.from_pretrained = classmethod(
            lambda cls, name, **k: cls(RobertaConfig(**rb._shim_text_config)))
        vt = importlib.import_module("model.video_transformer")
        mm = importlib.import_module("model.model")
        heads = importlib.import_module("model.heads")
        loss = importlib.import_module("model.loss")
    finally:
        os.chdir(cwd)
    mm._shim_fake_load = _fake_load
    mm._shim_real_load = real_load
    for mod in (mm, vt, rb):
        cfg = getattr(mod, "config", None) or getattr(mod, "config_yaml")
        cfg["use_checkpoint"] = use_checkpoint
    mm.config["use_checkpoint"] = use_checkpoint
    vt.config_yaml["use_checkpoint"] = use_checkpoint
    rb.config_yaml["use_checkpoint"] = use_checkpoint

    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29577")
        dist.init_process_group("gloo", rank=0, world_size=1)
    _loaded.update(mm=mm, vt=vt, rb=rb, heads=heads, loss=loss, RobertaConfig=RobertaConfig)
    return types.SimpleNamespace(**_loaded)


@contextlib.contextmanager
def tiny_wrapper(ref, *, img_size, embed_dim, depth, num_heads, text_cfg, dim_cross):
    """Make FrozenInTime build a small model: the wrapper hard-codes 12 layers /
    224^2 (model.py:77-83) and width 768 for the cross-attention K/V inputs
    (video_transformer.py:33, roberta.py:24)."""
    mm, vt, rb = ref.mm, ref.vt, ref.rb
    saved = (mm.SpaceTimeTransformer, vt.DIM_TEXT, rb.DIM_IMG, dict(rb._shim_text_config), rb.NUM_FUSE_BLOCK)
    mm.SpaceTimeTransformer = functools.partial(vt.SpaceTimeTransformer, img_size=img_size,
                                                embed_dim=embed_dim, depth=depth, num_heads=num_heads)
    vt.DIM_TEXT = dim_cross
    rb.DIM_IMG = dim_cross
    rb._shim_text_config.update(text_cfg)
    torch.load = mm._shim_fake_load
    try:
        yield
    finally:
        torch.load = mm._shim_real_load
        mm.SpaceTimeTransformer, vt.DIM_TEXT, rb.DIM_IMG, cfg, rb.NUM_FUSE_BLOCK = saved
        rb._shim_text_config.clear()
        rb._shim_text_config.update(cfg)
