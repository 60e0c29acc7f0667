"""B200 drop-in for EgoVLPv2/model/roberta.py (HF RoBERTa fork with gated text->video cross-attention).

Same module tree / parameter names as the reference.  It does not depend on the
`transformers` package: the configuration is a plain namespace with the RobertaConfig field names the callers read.
`RobertaLayer.forward` is one autograd node running the sm_100a kernels.  In train mode the reference's dropouts
(p=0.1: roberta.py:162,203 embeddings; :244,313 attention probabilities; :337,342 and :418,422 dense outputs; the
text->video cross-attention included) are applied inside those kernels from a Philox stream (egovlpv2_b200/rng.py).
"""
import types

import torch
from torch import nn

from .. import autograd as A
from .. import functional as Fn
from .. import rng
from ..weights import cache

NUM_FUSE_BLOCK = 6   # assigned by FrozenInTime (model.py:141)
DIM_IMG = 768        # roberta.py:24


class RobertaConfig(types.SimpleNamespace):
    """The RobertaConfig fields used on this path (defaults = roberta-base)."""

    def __init__(self, vocab_size=50265, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                 intermediate_size=3072, hidden_act="gelu", hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1,
                 max_position_embeddings=514, type_vocab_size=1, initializer_range=0.02, layer_norm_eps=1e-5,
                 pad_token_id=1, bos_token_id=0, eos_token_id=2, **kw):
        super().__init__(vocab_size=vocab_size, hidden_size=hidden_size, num_hidden_layers=num_hidden_layers,
                         num_attention_heads=num_attention_heads, intermediate_size=intermediate_size,
                         hidden_act=hidden_act, hidden_dropout_prob=hidden_dropout_prob,
                         attention_probs_dropout_prob=attention_probs_dropout_prob,
                         max_position_embeddings=max_position_embeddings, type_vocab_size=type_vocab_size,
                         initializer_range=initializer_range, layer_norm_eps=layer_norm_eps, pad_token_id=pad_token_id,
                         bos_token_id=bos_token_id, eos_token_id=eos_token_id, is_decoder=False,
                         add_cross_attention=False, chunk_size_feed_forward=0, **kw)
        if hidden_act != "gelu":
            raise NotImplementedError("only erf-GELU is implemented (roberta-base)")


class RobertaEmbeddings(nn.Module):
    """roberta.py:150-204."""

    def __init__(self, config):
        super().__init__()
        self.word_embeddings = nn.Embedding(config.vocab_size, config.hidden_size, padding_idx=config.pad_token_id)
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, config.hidden_size,
                                                padding_idx=config.pad_token_id)
        self.token_type_embeddings = nn.Embedding(config.type_vocab_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.register_buffer("position_ids", torch.arange(config.max_position_embeddings).expand((1, -1)))
        self.padding_idx = config.pad_token_id
        self._eps = config.layer_norm_eps
        self._p_attn = config.attention_probs_dropout_prob

    def forward(self, input_ids=None, token_type_ids=None, position_ids=None, inputs_embeds=None, past_key_values_length=0):
        if input_ids is None or token_type_ids is not None or position_ids is not None or inputs_embeds is not None:
            raise NotImplementedError("only input_ids with default token types / positions is on the EgoVLPv2 path")
        params = [self.word_embeddings.weight, self.position_embeddings.weight, self.token_type_embeddings.weight,
                  self.LayerNorm.weight, self.LayerNorm.bias]
        drop = None
        if self.training and torch.is_grad_enabled() and self.dropout.p > 0:
            drop = rng.drop_cfg(self.dropout.p, self._p_attn, input_ids.device, base=rng.begin_pass())
        return A.TextEmbedFn.apply(A.cfg(eps=self._eps, pad_id=self.padding_idx, drop=drop), input_ids, *params)


class RobertaSelfAttention(nn.Module):
    def __init__(self, config, layer_index=None):
        super().__init__()
        self.num_attention_heads = config.num_attention_heads
        self.attention_head_size = config.hidden_size // config.num_attention_heads
        if self.attention_head_size != 64:
            raise NotImplementedError("the attention kernels are specialised for head_dim 64")
        self.all_head_size = config.hidden_size
        self.query = nn.Linear(config.hidden_size, self.all_head_size)
        kv_in = DIM_IMG if layer_index is not None else config.hidden_size   # roberta.py:241-242
        self.key = nn.Linear(kv_in, self.all_head_size)
        self.value = nn.Linear(kv_in, self.all_head_size)
        self.dropout = nn.Dropout(config.attention_probs_dropout_prob)


class RobertaSelfOutput(nn.Module):
    def __init__(self, config, layer_index=None):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        if layer_index is None:   # the cross-attention output has no LayerNorm (roberta.py:335-336)
            self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)


class RobertaAttention(nn.Module):
    def __init__(self, config, layer_index=None):
        super().__init__()
        self.self = RobertaSelfAttention(config, layer_index=layer_index)
        self.output = RobertaSelfOutput(config, layer_index=layer_index)


class RobertaIntermediate(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.intermediate_size)


class RobertaOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.intermediate_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)


class RobertaLayer(nn.Module):
    """roberta.py:430-505."""

    def __init__(self, config, layer_index=None):
        super().__init__()
        self.attention = RobertaAttention(config)
        n_layers = config.num_hidden_layers
        if layer_index is not None and layer_index >= n_layers - NUM_FUSE_BLOCK:   # reference: 12 - NUM_FUSE_BLOCK (:438)
            self.crossattention_t2i = RobertaAttention(config, layer_index=layer_index)
            self.alpha_t2i = nn.Parameter(torch.Tensor([0]))
        self.intermediate = RobertaIntermediate(config)
        self.output = RobertaOutput(config)
        self.num_heads = config.num_attention_heads
        self._eps = config.layer_norm_eps
        self.layer_index = 0 if layer_index is None else layer_index

    def forward(self, hidden_states, attention_mask=None, head_mask=None, encoder_hidden_states=None,
                encoder_attention_mask=None, past_key_value=None, output_attentions=False, last_norm=True):
        if head_mask is not None or past_key_value is not None or output_attentions or encoder_attention_mask is not None:
            raise NotImplementedError("head masks / kv caches / attention outputs / encoder masks are not on the EgoVLPv2 path")
        if not last_norm:
            raise NotImplementedError("last_norm=False is never used by this code base")
        fused = encoder_hidden_states is not None
        if fused:
            assert hasattr(self, "crossattention_t2i"), \
                "If `encoder_hidden_states` are passed, the layer has to be instantiated with cross-attention"
        names = Fn.TEXT_LAYER_PARAMS + (Fn.TEXT_FUSE_PARAMS if fused else [])
        params = [self.get_parameter(n) for n in names]
        byname = dict(zip(names, params))
        c = cache()
        wnames = ["attention.output.dense.weight", "intermediate.dense.weight", "output.dense.weight"]
        if fused:
            wnames += ["crossattention_t2i.self.query.weight", "crossattention_t2i.output.dense.weight"]
        w = {n: c.bf16(byname[n]) for n in wnames}
        sa = "attention.self."
        w["qkv"] = c.cat_bf16([byname[sa + "query.weight"], byname[sa + "key.weight"], byname[sa + "value.weight"]])
        pcat = {"qkv.bias": c.cat_f32([byname[sa + "query.bias"], byname[sa + "key.bias"], byname[sa + "value.bias"]])}
        if fused:
            ca = "crossattention_t2i.self."
            w["cross.kv"] = c.cat_bf16([byname[ca + "key.weight"], byname[ca + "value.weight"]])
            pcat["cross.kv.bias"] = c.cat_f32([byname[ca + "key.bias"], byname[ca + "value.bias"]])
        B, S, _ = hidden_states.shape
        if attention_mask is None:
            key_bias = None
        else:   # extended additive mask [B,1,1,S] (0 / finfo.min)
            key_bias = attention_mask.reshape(B, S).float().contiguous()
        drop = None
        p_hidden, p_attn = self.output.dropout.p, self.attention.self.dropout.p
        if self.training and torch.is_grad_enabled() and (p_hidden > 0 or p_attn > 0):
            if S > 64:
                raise NotImplementedError("train-mode dropout of the text tower is implemented for <= 64 tokens (got %d)" % S)
            drop = rng.drop_cfg(p_hidden, p_attn, hidden_states.device)
        cfg = A.cfg(names=names, H=self.num_heads, eps=self._eps, drop=drop, layer=self.layer_index)
        out = A.TextLayerFn.apply(cfg, w, pcat, hidden_states, key_bias, encoder_hidden_states, *params)
        return (out,)


class RobertaEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.layer = nn.ModuleList([RobertaLayer(config, layer_index=i) for i in range(config.num_hidden_layers)])

    def forward(self, hidden_states, attention_mask=None, **unused):
        from .. import reduce as _reduce
        red = _reduce.active()
        for layer in self.layer:
            if red is not None:
                red.watch(hidden_states, layer)   # EgoNCE pass: see reduce.OverlappedGradReducer
            hidden_states = layer(hidden_states, attention_mask)[0]
        return types.SimpleNamespace(last_hidden_state=hidden_states)


class RobertaPooler(nn.Module):
    """roberta.py:610-622: the dense layer was stripped; tanh of the first token (unused by FrozenInTime)."""

    def __init__(self, config):
        super().__init__()
        self.activation = nn.Tanh()

    def forward(self, hidden_states):
        return self.activation(hidden_states[:, 0])


class ModelOutput(dict):
    """Minimal BaseModelOutputWithPoolingAndCrossAttentions: attribute, key and index access."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __getitem__(self, k):
        if isinstance(k, int):
            return list(self.values())[k]
        return dict.__getitem__(self, k)


class RobertaModel(nn.Module):
    """roberta.py:712-878."""

    _keys_to_ignore_on_load_missing = [r"position_ids"]

    def __init__(self, config, add_pooling_layer=True):
        super().__init__()
        self.config = config
        self.embeddings = RobertaEmbeddings(config)
        self.encoder = RobertaEncoder(config)
        self.pooler = RobertaPooler(config) if add_pooling_layer else None
        self.apply(self._init_weights)

    def _init_weights(self, module):
        if isinstance(module, nn.Linear):
            module.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
            if module.bias is not None:
                module.bias.data.zero_()
        elif isinstance(module, nn.Embedding):
            module.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
            if module.padding_idx is not None:
                module.weight.data[module.padding_idx].zero_()
        elif isinstance(module, nn.LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)

    @classmethod
    def from_pretrained(cls, name_or_path="roberta-base", state_dict=None, allow_random_init=False, **kw):
        """roberta-base architecture.  Weights, in this order: `state_dict`; a local file `name_or_path` (torch.save'd
        state_dict); the Hugging Face cache / hub through `transformers` when it is importable (the reference always
        starts from pretrained roberta-base: model.py:69).  If none of them yields weights the model keeps its random
        init and says so LOUDLY (a warning; silence it with allow_random_init=True -- synthetic benchmarks and tests)."""
        import os
        import warnings
        allow_random_init = allow_random_init or os.environ.get("EGV_ALLOW_RANDOM_INIT", "0") == "1"
        model = cls(RobertaConfig(**kw))
        if state_dict is None and isinstance(name_or_path, str) and os.path.isfile(name_or_path):
            state_dict = torch.load(name_or_path, map_location="cpu")
        if state_dict is None and not allow_random_init and model.config.hidden_size == 768 and model.config.num_hidden_layers == 12:
            try:
                os.environ.setdefault("HF_HUB_OFFLINE", "1")     # never block on the network
                from transformers import RobertaModel as _HF
                state_dict = _HF.from_pretrained(name_or_path, add_pooling_layer=False).state_dict()
            except Exception:
                state_dict = None
        if state_dict is not None:
            state_dict = {k[len("roberta."):] if k.startswith("roberta.") else k: v for k, v in state_dict.items()}
            missing, unexpected = model.load_state_dict(state_dict, strict=False)
            missing = [k for k in missing if "crossattention_t2i" not in k and "alpha_t2i" not in k and "position_ids" not in k]
            if missing or unexpected:
                warnings.warn("RobertaModel.from_pretrained(%r): %d missing keys (e.g. %s), %d unexpected keys (e.g. %s)"
                              % (name_or_path, len(missing), missing[:3], len(unexpected), list(unexpected)[:3]))
        elif not allow_random_init:
            warnings.warn("RobertaModel.from_pretrained(%r): no pretrained weights found (no local file, no Hugging Face cache): "
                          "the text tower keeps its RANDOM initialisation, unlike the reference, which starts from "
                          "roberta-base.  Pass state_dict= / a local path, or allow_random_init=True to silence this."
                          % (name_or_path,))
        return model

    def get_input_embeddings(self):
        return self.embeddings.word_embeddings

    def get_extended_attention_mask(self, attention_mask, input_shape=None, device=None, dtype=None):
        """(1 - mask)[:, None, None, :] * finfo(fp32).min  (transformers.PreTrainedModel; model.py:251)."""
        if attention_mask.dim() != 2:
            raise NotImplementedError("only [B, S] padding masks are on the EgoVLPv2 path")
        m = attention_mask[:, None, None, :].to(torch.float32)
        return (1.0 - m) * torch.finfo(torch.float32).min

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, position_ids=None, head_mask=None,
                inputs_embeds=None, encoder_hidden_states=None, encoder_attention_mask=None, past_key_values=None,
                use_cache=None, output_attentions=None, output_hidden_states=None, return_dict=None):
        if input_ids is None:
            raise ValueError("You have to specify input_ids")
        if attention_mask is None:
            attention_mask = torch.ones_like(input_ids)
        ext = self.get_extended_attention_mask(attention_mask, input_ids.shape, input_ids.device)
        h = self.embeddings(input_ids=input_ids, token_type_ids=token_type_ids, position_ids=position_ids,
                            inputs_embeds=inputs_embeds)
        h = self.encoder(h, ext).last_hidden_state
        pooled = self.pooler(h) if self.pooler is not None else None
        return ModelOutput(last_hidden_state=h, pooler_output=pooled)
