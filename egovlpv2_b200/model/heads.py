"""B200 drop-in for EgoVLPv2/model/heads.py: Pooler, ITMHead, MLMHead (parameter containers + kernel-backed forward)."""
import torch
from torch import nn

from .. import autograd as A
from ..lib import ACT_GELU, ACT_NONE, ACT_TANH
from ..weights import cache


def _linear(x, lin, act=ACT_NONE):
    """x [..., K] -> act(x W^T + b) through the tcgen05 GEMM with fused epilogue."""
    shp = x.shape
    has_b = lin.bias is not None
    params = (lin.weight, lin.bias) if has_b else (lin.weight,)
    out = A.MlpChainFn.apply(A.cfg(acts=[act], has_bias=[has_b]), [cache().bf16(lin.weight)], x.reshape(-1, shp[-1]), *params)
    return out.view(*shp[:-1], lin.weight.shape[0])


class Pooler(nn.Module):
    """heads.py:15-27"""

    def __init__(self, hidden_size):
        super().__init__()
        self.dense = nn.Linear(hidden_size, hidden_size)
        self.activation = nn.Tanh()

    def forward(self, hidden_states):
        return _linear(hidden_states, self.dense, ACT_TANH)


class ITMHead(nn.Module):
    """heads.py:30-35"""

    def __init__(self, hidden_size):
        super().__init__()
        self.fc = nn.Linear(hidden_size, 2)

    def forward(self, x):
        return _linear(x, self.fc)


class BertPredictionHeadTransform(nn.Module):
    """transformers BertPredictionHeadTransform (heads.py:12,41): dense -> erf-GELU -> LayerNorm(config.layer_norm_eps)."""

    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=getattr(config, "layer_norm_eps", 1e-12))

    def forward(self, x):
        shp = x.shape
        t = _linear(x, self.dense, ACT_GELU).reshape(-1, shp[-1])
        return A.LayerNormRowsFn.apply(t, self.LayerNorm.weight, self.LayerNorm.bias, self.LayerNorm.eps).view(shp)


class MLMHead(nn.Module):
    """heads.py:38-50"""

    def __init__(self, config, weight=None):
        super().__init__()
        self.transform = BertPredictionHeadTransform(config)
        self.decoder = nn.Linear(config.hidden_size, config.vocab_size, bias=False)
        self.bias = nn.Parameter(torch.zeros(config.vocab_size))
        if weight is not None:
            self.decoder.weight = weight

    def forward(self, x):
        shp = x.shape
        t = self.transform(x).reshape(-1, shp[-1])
        out = A.MlpChainFn.apply(A.cfg(acts=[ACT_NONE], has_bias=[True]), [cache().bf16(self.decoder.weight)], t,
                                 self.decoder.weight, self.bias)
        return out.view(*shp[:-1], self.decoder.weight.shape[0])
