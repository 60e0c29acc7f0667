"""B200 drop-in for EgoVLPv2/model/video_transformer.py (TimeSformer / Frozen-in-Time space-time transformer).

Same module tree and parameter names as the reference; `SpaceTimeBlock.forward` and
`SpaceTimeTransformer.forward_features` dispatch to the sm_100a kernels through egovlpv2_b200.autograd.
The nn.Linear / nn.LayerNorm / nn.Conv2d children are parameter containers only: their own forward is never
called on the hot path.
"""
import os
from functools import partial

import torch
from torch import nn

from .. import autograd as A
from .. import functional as Fn
from ..weights import cache

NUM_FUSE_BLOCK = 6
# the last block of a pass that consumes only `x[:, 0]` computes its post-attention part on the CLS rows only
# (exactly the same CLS row; set EGV_CLS_ONLY_LAST=0 to run the full block)
CLS_ONLY_LAST_BLOCK = os.environ.get("EGV_CLS_ONLY_LAST", "1") != "0"
DIM_TEXT = 768   # video_transformer.py:33


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


class Mlp(nn.Module):
    """video_transformer.py:42-58 (parameter container; fused into SpaceTimeBlock)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)
        if drop != 0.:
            raise NotImplementedError("dropout > 0 in the video tower is not used by pre-training (SURVEY.md Q7)")


class VideoPatchEmbed(nn.Module):
    """video_transformer.py:62-83: Conv2d(k=s=patch) as an im2col kernel + tcgen05 GEMM."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, num_frames=8):
        super().__init__()
        img_size, patch_size = to_2tuple(img_size), to_2tuple(patch_size)
        self.img_size, self.patch_size = img_size, patch_size
        self.num_patches = (img_size[1] // patch_size[1]) * (img_size[0] // patch_size[0]) * num_frames
        self.num_frames = num_frames
        self.embed_dim = embed_dim
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        B, F, C, H, W = x.shape
        assert F == self.num_frames, (F, self.num_frames)
        w = self.proj.weight
        layers_w = [cache().bf16(w, (w.shape[0], w[0].numel()))]
        gh, gw = H // self.patch_size[0], W // self.patch_size[1]
        # [B*F*gh*gw, 3*p*p] patches -> GEMM -> [B*F, gh, gw, C] -> the reference's [B*F, C, gh, gw] view
        from .. import lib as L
        K = L.kernels()
        cols = torch.empty(B * F * gh * gw, C * self.patch_size[0] * self.patch_size[1], dtype=torch.bfloat16, device=x.device)
        K.patchify(x.reshape(B * F, C, H, W).float().contiguous(), self.patch_size[0], cols)
        out = A.MlpChainFn.apply(A.cfg(acts=[0], has_bias=[True]), layers_w, cols, w, self.proj.bias)
        return out.view(B * F, gh, gw, self.embed_dim).permute(0, 3, 1, 2)


class VarAttention(nn.Module):
    """video_transformer.py:86-185 (parameter container; the math lives in functional.video_block_fwd)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0., initialize='random',
                 dim_text=None, norm_layer=nn.LayerNorm, space_attn=True):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        if head_dim != 64:
            raise NotImplementedError("the attention kernels are specialised for head_dim 64 (got %d)" % head_dim)
        if qk_scale is not None and abs(qk_scale - head_dim ** -0.5) > 1e-12:
            raise NotImplementedError("custom qk_scale")
        if not qkv_bias or attn_drop != 0. or proj_drop != 0.:
            raise NotImplementedError("qkv_bias=False / dropout > 0 are not used on the pre-training path")
        self.scale = head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        if initialize == 'zeros':   # video_transformer.py:96-102
            self.qkv.weight.data.fill_(0)
            self.qkv.bias.data.fill_(0)
            self.proj.weight.data.fill_(1)
            self.proj.bias.data.fill_(0)
        if dim_text is not None and space_attn:
            self.qkv_text_i2t = nn.Linear(dim_text, dim * 2, bias=qkv_bias)
            self.qkv_i2t = nn.Linear(dim, dim, bias=qkv_bias)
            self.proj_i2t = nn.Linear(dim, dim)
            self.alpha_i2t = nn.Parameter(torch.Tensor([0]))
            self.norm_i2t_i = norm_layer(dim)


class SpaceTimeBlock(nn.Module):
    """video_transformer.py:188-228.  One autograd node per block."""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0., drop_path=0.,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm, time_init='zeros', attention_style='frozen-in-time',
                 dim_text=None):
        super().__init__()
        if attention_style != 'frozen-in-time':
            raise NotImplementedError
        if drop_path > 0.:
            raise NotImplementedError("stochastic depth is not used by pre-training (SURVEY.md Q7)")
        self.norm1 = norm_layer(dim)
        self.attn = VarAttention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                                 proj_drop=drop, dim_text=dim_text, norm_layer=norm_layer, space_attn=True)
        self.timeattn = VarAttention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                                     proj_drop=drop, initialize=time_init, dim_text=dim_text, norm_layer=norm_layer,
                                     space_attn=False)
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        self.norm3 = norm_layer(dim)
        self.attention_style = attention_style
        self.num_heads = num_heads
        self.has_fusion = dim_text is not None
        self._eps = self.norm1.eps

    def _params(self, fused):
        names = Fn.VIDEO_BLOCK_PARAMS + (Fn.VIDEO_FUSE_PARAMS if fused else [])
        return names, [self.get_parameter(n) for n in names]

    def i2t_prep(self, y, y_mask=None):
        """(extension) The text side of this block's gated video->text cross-attention, re-associated around the text
        tokens (xattn_reassoc.py): (Mt, U, bias1), functions of the text states and weights only.  FrozenInTime calls it
        on the text tower's stream and hands the result to forward(..., i2t_prep=...); None when the re-associated
        kernels do not cover the shape (forward then takes y and does everything itself)."""
        from .. import xattn_reassoc as XR
        dim = self.norm1.normalized_shape[0]
        if not (self.has_fusion and Fn.reassoc_enabled() and XR.i2t_supported(y.shape[1], dim, self.num_heads)):
            return None
        params = [self.get_parameter(n) for n in Fn.I2T_PREP_PARAMS]
        w = {n: cache().bf16(p) for n, p in zip(Fn.I2T_PREP_PARAMS, params) if p.dim() == 2}
        y_bias = None if y_mask is None else y_mask.reshape(y.shape[0], -1).float().contiguous()
        return A.I2TPrepFn.apply(A.cfg(H=self.num_heads), w, y, y_bias, *params)

    def forward(self, x, einops_from_space, einops_to_space, einops_from_time, einops_to_time, time_n, space_f, y=None,
                y_mask=None, cls_only=False, i2t_prep=None):
        """`cls_only` (extension, default off): return only the CLS row, [B, 1, C] -- for a block whose output is
        consumed as `x[:, 0]` (the last block of the EgoNCE and ITM passes); the row equals the full forward's.
        `i2t_prep` (extension): the result of self.i2t_prep(y, y_mask)."""
        fused = y is not None
        if fused and not self.has_fusion:
            raise AttributeError("this SpaceTimeBlock was built without cross-attention parameters (dim_text=None)")
        names, params = self._params(fused)
        w = {n: cache().bf16(p) for n, p in zip(names, params) if p.dim() == 2}
        y_bias = None
        if fused and y_mask is not None:
            y_bias = y_mask.reshape(y.shape[0], -1).float().contiguous()   # [B,1,1,S] additive mask -> [B,S]
        cfg = A.cfg(names=names, H=self.num_heads, T=space_f, Nf=time_n, eps=self._eps)
        if cls_only:
            return A.VideoBlockClsFn.apply(cfg, w, x, y, y_bias, *params).unsqueeze(1)
        Mt, U, bias1 = i2t_prep if i2t_prep is not None else (None, None, None)
        return A.VideoBlockFn.apply(cfg, w, x, y, y_bias, Mt, U, bias1, *params)


class SpaceTimeTransformer(nn.Module):
    """video_transformer.py:231-399."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=4., qkv_bias=True, qk_scale=None, representation_size=None, drop_rate=0., attn_drop_rate=0.,
                 drop_path_rate=0., hybrid_backbone=None, num_frames=8, time_init='rand', attention_style='frozen-in-time',
                 norm_layer=nn.LayerNorm, dim_text=None, fuse_from=6):
        super().__init__()
        if hybrid_backbone is not None:
            raise NotImplementedError('hybrid backbone not implemented')
        if drop_rate != 0. or drop_path_rate != 0. or representation_size:
            raise NotImplementedError("dropout / stochastic depth / representation layer are unused by pre-training")
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.num_frames = num_frames
        # statistics applied to uint8 frames on the device (the reference's loader applies them on the host)
        self.norm_mean, self.norm_std = Fn.VIDEO_NORM_MEAN, Fn.VIDEO_NORM_STD
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        self.patch_embed = VideoPatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim,
                                           num_frames=num_frames)
        self.patches_per_frame = self.patch_embed.num_patches // num_frames
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patches_per_frame + 1, embed_dim))
        self.temporal_embed = nn.Parameter(torch.zeros(1, num_frames, embed_dim))
        self.pos_drop = nn.Dropout(p=drop_rate)
        # cross-attention parameters exist for blocks >= 6 with text width DIM_TEXT (video_transformer.py:302)
        self.blocks = nn.ModuleList([
            SpaceTimeBlock(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                           drop=drop_rate, attn_drop=attn_drop_rate, drop_path=0., norm_layer=norm_layer,
                           time_init=time_init, attention_style=attention_style,
                           dim_text=None if i < fuse_from else (dim_text or DIM_TEXT))
            for i in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.pre_logits = nn.Identity()
        self.head = nn.Linear(self.num_features, num_classes) if num_classes > 0 else nn.Identity()
        nn.init.trunc_normal_(self.pos_embed, std=.02)
        nn.init.trunc_normal_(self.cls_token, std=.02)
        if num_frames == 1:
            self.apply(self._init_weights)
        self.einops_from_space = 'b (f n) d'
        self.einops_to_space = '(b f) n d'
        self.einops_from_time = 'b (f n) d'
        self.einops_to_time = '(b n) f d'

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_embed', 'cls_token'}

    def get_classifier(self):
        return self.head

    def reset_classifier(self, num_classes, global_pool=''):
        self.num_classes = num_classes
        self.head = nn.Linear(self.embed_dim, num_classes) if num_classes > 0 else nn.Identity()

    def tokens(self, x, cls_token=None):
        """patch embedding + CLS + tiled spatial / repeated temporal position terms (video_transformer.py:354-372;
        model.py:211-232 passes FrozenInTime.cls_token).  x [B,T,3,H,W] -> [B, 1+T*Nf, C] f32."""
        B, T = x.shape[:2]
        assert T == self.patch_embed.num_frames, (T, self.patch_embed.num_frames)   # video_transformer.py:80
        pw = self.patch_embed.proj.weight
        w = {"patch_embed.proj.weight": cache().bf16(pw, (pw.shape[0], pw[0].numel()))}
        cls = self.cls_token if cls_token is None else cls_token
        # uint8 frames are normalised inside the im2col kernel with the loader's statistics (transforms.py:39-49)
        cfg = A.cfg(patch=self.patch_embed.patch_size[0], norm=(self.norm_mean, self.norm_std))
        return A.VideoTokensFn.apply(cfg, w, x, pw, self.patch_embed.proj.bias, self.pos_embed, self.temporal_embed, cls)

    def forward_features(self, x):
        b, curr_frames = x.shape[:2]
        x = self.tokens(x)
        n, f = self.patches_per_frame, curr_frames
        from .. import reduce as _reduce
        red = _reduce.active()
        for i, blk in enumerate(self.blocks):
            if red is not None:
                red.watch(x, blk)     # EgoNCE pass: block i's parameter gradients are final once d(loss)/dx_i exists
            # only the CLS row of the last block is consumed below: its per-token tail is skipped
            x = blk(x, self.einops_from_space, self.einops_to_space, self.einops_from_time, self.einops_to_time,
                    time_n=n, space_f=f, cls_only=CLS_ONLY_LAST_BLOCK and i == len(self.blocks) - 1)
        # only the CLS row of the final norm is consumed (video_transformer.py:391)
        x = A.LayerNormRowsFn.apply(x[:, 0], self.norm.weight, self.norm.bias, self.norm.eps)
        return self.pre_logits(x)

    def forward(self, x):
        x = self.forward_features(x)
        if isinstance(self.head, nn.Linear):
            raise NotImplementedError("classification head is not on the EgoVLPv2 path (FrozenInTime sets head = Identity)")
        return self.head(x)
