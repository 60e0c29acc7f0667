"""B200 drop-in for EgoVLPv2/model/model.py: `FrozenInTime`, `sim_matrix`, `sim_matrix_batch_val`.

Same constructor, `forward` / `infer` / `compute_*` signatures, attribute tree and state_dict keys as the reference
(model.py:46-595).  The three passes of one pre-training step (EgoNCE dual-encoder pass, fused MLM pass, fused ITM
pass with hard negatives; model.py:370-487) run on the sm_100a kernels of libegovlp_b200.so.

Differences from the reference that do not change results:
  * activation checkpointing (`use_checkpoint`) is accepted and ignored: B200 has 180 GB of HBM;
  * the MLM pass skips the last video block, whose output the reference computes and discards;
  * MLM / ITM cross-entropies are reduced as (local sum, local count) + a scalar all-reduce instead of all-gathering
    the [B*S, 50265] logits; the value and the gradients are identical;
  * ITM hard negatives are drawn on the device (same distribution, no per-row host sync);
  * the ITM pass reuses the MLM pass's activations of the unfused video blocks for the rows whose clip is unchanged
    (the label-1 half of the batch) and recomputes them only for the label-0 half: the video side has no dropout or
    stochastic depth, both passes use the same CLS token and weights, so the values are identical.
Known gap: text-side dropout (p = 0.1 in the reference's train mode) is not applied (parity is defined in eval mode).
"""
import os
from functools import partial

import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn

from .. import autograd as A
from .. import rng
from .. import streams
from ..lib import ACT_NONE, ACT_RELU, ACT_TANH
from ..weights import cache
from . import heads, roberta, video_transformer
from .roberta import RobertaConfig, RobertaModel
from .video_transformer import SpaceTimeTransformer

DEFAULT_CONFIG = dict(input_image_embed_size=768, vocab_size=50265, mlm_prob=0.15, input_text_embed_size=768,
                      hidden_size=768, num_heads=12, num_layers=12, mlp_ratio=4, drop_rate=0.1, num_fuse_block=6,
                      use_checkpoint=True, decay_power="cosine", end_lr=1e-7, warmup_steps=0.1)


def _load_yaml_config():
    """The reference opens ./EgoNCE_MLM_ITM_Config.yml relative to the CWD at import (model.py:32)."""
    cfg = dict(DEFAULT_CONFIG)
    path = './EgoNCE_MLM_ITM_Config.yml'
    if os.path.exists(path):
        try:
            import yaml
            with open(path) as f:
                cfg.update(yaml.load(f, Loader=yaml.FullLoader) or {})
        except Exception as e:   # the reference would fail at import; falling back silently would hide a broken config
            import warnings
            warnings.warn("egovlpv2_b200: could not read %s (%r); using the built-in defaults of EgoNCE_MLM_ITM_Config.yml" % (path, e))
    return cfg


config = _load_yaml_config()


def init_weights(module):
    """model.py:34-43"""
    if isinstance(module, (nn.Linear, nn.Embedding)):
        module.weight.data.normal_(mean=0.0, std=0.02)
    elif isinstance(module, nn.LayerNorm):
        module.bias.data.zero_()
        module.weight.data.fill_(1.0)
    if isinstance(module, nn.Linear) and module.bias is not None:
        module.bias.data.zero_()


def state_dict_data_parallel_fix(load_state_dict, curr_state_dict):
    """utils/util.py:31-57: add / strip the DataParallel 'module.' prefix so the keys match the current model."""
    load_keys, curr_keys = list(load_state_dict.keys()), list(curr_state_dict.keys())
    if not load_keys or not curr_keys:
        return load_state_dict
    redo_dp = undo_dp = False
    if not curr_keys[0].startswith('module.') and load_keys[0].startswith('module.'):
        undo_dp = True
    elif curr_keys[0].startswith('module.') and not load_keys[0].startswith('module.'):
        redo_dp = True
    if undo_dp:
        return {k[7:]: v for k, v in load_state_dict.items()}
    if redo_dp:
        return {'module.' + k: v for k, v in load_state_dict.items()}
    return load_state_dict


class FrozenInTime(nn.Module):
    DUAL_TASK = 'EgoNCE'     # task name of the dual-encoder branch of infer() (model.py:198; 'Dual' in model_epic_charades.py:218)

    def __init__(self, video_params, text_params, projection_dim=4096, load_checkpoint=None, projection='minimal',
                 load_temporal_fix='bilinear', config=config, task_names='EgoNCE_ITM_MLM', norm_layer=None, embed_dim=768):
        super().__init__()
        self.video_params = video_params
        self.text_params = text_params
        self.load_temporal_fix = load_temporal_fix
        self.config = config
        self.task_names = task_names
        if not text_params['pretrained']:
            raise NotImplementedError("Huggingface text models require pretrained init.")
        fused_heads = ('MLM' in self.task_names or 'ITM' in self.task_names)
        n_layers, n_fuse = self.config["num_layers"], self.config["num_fuse_block"]
        if fused_heads:   # model.py:141-143: module globals consumed when the towers are built
            roberta.NUM_FUSE_BLOCK = n_fuse
            roberta.DIM_IMG = self.config["input_image_embed_size"]

        if self.text_params['model'].startswith('roberta'):
            tcfg = dict(text_params.get('config') or {})
            self.text_model = RobertaModel.from_pretrained(text_params.get('checkpoint', "roberta-base"),
                                                           allow_random_init=text_params.get('allow_random_init', False), **tcfg)
        else:
            raise NotImplementedError(f"{text_params['model']} not implemented")
        self.text_model.train()

        if video_params['model'] == "SpaceTimeTransformer":
            self.num_frames = video_params['num_frames']
            if video_params.get('arch_config', 'base_patch16_224') != 'base_patch16_224':
                raise NotImplementedError
            vkw = {k: video_params[k] for k in ('img_size', 'patch_size', 'embed_dim', 'depth', 'num_heads') if k in video_params}
            depth = vkw.get('depth', 12)
            model = SpaceTimeTransformer(num_frames=self.num_frames, time_init='zeros', attention_style='frozen-in-time',
                                         dim_text=self.config["input_text_embed_size"],
                                         fuse_from=(depth - n_fuse) if fused_heads else 6, **vkw)
            model.head = nn.Identity()
            model.pre_logits = nn.Identity()
            ftr_dim = model.embed_dim
            vit_path = video_params.get('vit_checkpoint')
            if load_checkpoint in ["", None] and vit_path and os.path.isfile(vit_path):
                vit_checkpoint = torch.load(vit_path, map_location="cpu")
                model.load_state_dict(state_dict_data_parallel_fix(vit_checkpoint, model.state_dict()), strict=False)
            self.video_model = model
        else:
            raise NotImplementedError(f"{video_params['model']} not implemented")
        self.video_model.fc = nn.Identity()

        txt_proj, vid_proj = self._build_projections(projection, self.text_model.config.hidden_size, ftr_dim, projection_dim)
        self.txt_proj, self.vid_proj = txt_proj, vid_proj

        if fused_heads:
            bert_config = RobertaConfig(vocab_size=self.config["vocab_size"], hidden_size=self.config["hidden_size"],
                                        num_hidden_layers=n_layers, num_attention_heads=self.config["num_heads"],
                                        intermediate_size=self.config["hidden_size"] * self.config["mlp_ratio"],
                                        hidden_dropout_prob=self.config["drop_rate"],
                                        attention_probs_dropout_prob=self.config["drop_rate"], layer_norm_eps=1e-12)
            self.num_fuse_block = n_fuse
            self.num_text_layer = n_layers
            self.video_model.NUM_FUSE_BLOCK = n_fuse
            self.video_model.DIM_TXT = self.config["input_text_embed_size"]
            self.cross_modal_text_transform = nn.Linear(self.config["input_text_embed_size"], self.config["hidden_size"])
            self.cross_modal_text_transform.apply(init_weights)
            self.cross_modal_video_transform = nn.Linear(self.config["input_image_embed_size"], self.config["hidden_size"])
            self.cross_modal_video_transform.apply(init_weights)
            self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
            self.num_patches = self.video_model.patch_embed.num_patches
            self.patches_per_frame = self.num_patches // self.num_frames
            norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
            self.norm = norm_layer(embed_dim)
            self.pre_logits = nn.Identity()
            self.avgpool = nn.AdaptiveAvgPool1d(1)
            self.cross_modal_video_pooler = heads.Pooler(self.config["hidden_size"])
            self.cross_modal_video_pooler.apply(init_weights)
            self.cross_modal_text_pooler = heads.Pooler(self.config["hidden_size"])
            self.cross_modal_text_pooler.apply(init_weights)
            self.einops_from_space = 'b (f n) d'
            self.einops_to_space = '(b f) n d'
            self.einops_from_time = 'b (f n) d'
            self.einops_to_time = '(b n) f d'
        if 'MLM' in self.task_names:
            self.mlm_score = heads.MLMHead(bert_config)
            self.mlm_score.apply(init_weights)
        if 'ITM' in self.task_names:
            self.itm_score = heads.ITMHead(self.config["hidden_size"] * 2)
            self.itm_score.apply(init_weights)

        self.itm_plan = None   # test hook: dict(labels, swap_video, neg_idx) replacing the random ITM plan
        self.share_itm_prefix = True   # reuse MLM-pass activations of the unfused video blocks in the ITM pass (Q7-iii)

        if load_checkpoint not in ["", None]:
            checkpoint = torch.load(load_checkpoint, map_location='cpu')
            state_dict = checkpoint['state_dict']
            new_state_dict = state_dict_data_parallel_fix(state_dict, self.state_dict())
            new_state_dict = self._inflate_positional_embeds(new_state_dict)
            self.load_state_dict(new_state_dict, strict=False)

    def _build_projections(self, projection, txt_dim, ftr_dim, projection_dim):
        """model.py:104-121"""
        if projection == 'minimal':
            txt_proj = nn.Sequential(nn.Linear(txt_dim, projection_dim, bias=False),
                                     nn.ReLU(inplace=True), nn.Linear(projection_dim, projection_dim, bias=True),
                                     nn.ReLU(inplace=True), nn.Linear(projection_dim, projection_dim, bias=True))
            vid_proj = nn.Sequential(nn.Linear(ftr_dim, projection_dim, bias=False),
                                     nn.ReLU(inplace=True), nn.Linear(projection_dim, projection_dim, bias=True),
                                     nn.ReLU(inplace=True), nn.Linear(projection_dim, projection_dim, bias=True))
        elif projection == '':
            txt_proj, vid_proj = nn.Identity(), nn.Identity()
        else:
            raise NotImplementedError
        return txt_proj, vid_proj

    def set_device(self, device):
        self.device = device

    # ------------------------------------------------------------------------------------------ towers
    def _proj(self, seq, x):
        if isinstance(seq, nn.Identity):
            return x
        lins = [m for m in seq if isinstance(m, nn.Linear)]
        acts = [ACT_RELU] * (len(lins) - 1) + [ACT_NONE]
        if isinstance(seq[0], nn.ReLU):    # model_epic_charades.py:118: Sequential(ReLU(), Linear)
            x = A.ReluRowsFn.apply(x.reshape(-1, x.shape[-1])).view(x.shape)
        params = []
        for lin in lins:
            params.append(lin.weight)
            if lin.bias is not None:
                params.append(lin.bias)
        shp = x.shape
        out = A.MlpChainFn.apply(A.cfg(acts=acts, has_bias=[lin.bias is not None for lin in lins]),
                                 [cache().bf16(lin.weight) for lin in lins], x.reshape(-1, shp[-1]), *params)
        return out.view(*shp[:-1], out.shape[-1])

    def compute_text(self, text_data):
        if not self.text_params['model'].startswith('roberta'):
            raise NotImplementedError
        h = self.text_model(**text_data).last_hidden_state[:, 0, :]
        return self._proj(self.txt_proj, h)

    def compute_text_tokens(self, text_data):
        if not self.text_params['model'].startswith('roberta'):
            raise NotImplementedError
        h = self.text_model(**text_data).last_hidden_state
        return self._proj(self.txt_proj, h)

    def compute_video(self, video_data):
        return self._proj(self.vid_proj, self.video_model(video_data))

    def _video_prefix(self, video_data):
        """tokens + the unfused video blocks (model.py:211-244): [B,T,3,H,W] -> [B,N,C] entering the first fused block"""
        vm = self.video_model
        n, f = self.patches_per_frame, video_data.shape[1]
        x = vm.tokens(video_data, cls_token=self.cls_token)
        x = vm.pos_drop(x)
        unfused = self.num_text_layer - self.num_fuse_block
        es = (self.einops_from_space, self.einops_to_space, self.einops_from_time, self.einops_to_time)
        for blk in vm.blocks[:unfused]:
            x = blk(x, *es, time_n=n, space_f=f)
        return x

    def _fused_stack(self, video_data, input_ids, attention_mask, need_video_out=True, video_prefix=None):
        """model.py:211-271 / 295-357: (num_layers - num_fuse_block) plain blocks per tower, then the fused pairs: video
        block i reads the text entering layer i, text layer i reads the video ENTERING block i.
        `video_prefix`: precomputed output of the unfused video blocks (see _video_prefix), or a callable producing it
        (evaluated after the text prefix has been enqueued, so the two can overlap on two streams).
        Returns (video out, text out, video prefix)."""
        vm, tm = self.video_model, self.text_model
        n, f = self.patches_per_frame, video_data.shape[1]
        unfused = self.num_text_layer - self.num_fuse_block
        es = (self.einops_from_space, self.einops_to_space, self.einops_from_time, self.einops_to_time)
        ext = tm.get_extended_attention_mask(attention_mask, attention_mask.size(), input_ids.device)
        streams.to_side(ext, input_ids)
        with streams.side():   # text prefix next to the video prefix
            h = tm.embeddings(input_ids=input_ids)
            for layer in tm.encoder.layer[:unfused]:
                h = layer(h, ext)[0]
        if video_prefix is None:
            video_prefix = self._video_prefix(video_data)
        elif callable(video_prefix):
            video_prefix = video_prefix()
        x = video_prefix
        last = self.num_text_layer - 1
        for i in range(unfused, self.num_text_layer):
            run_video = need_video_out or i < last   # the last video block's output is unused by the MLM head
            cls_only = video_transformer.CLS_ONLY_LAST_BLOCK and i == last   # the ITM head reads only its CLS row (model.py:275-278)
            prep = None
            if run_video and not cls_only:
                with streams.side():   # text side of video block i's cross-attention: on the text stream, behind text layer i - 1
                    prep = vm.blocks[i].i2t_prep(h, ext)
            streams.exchange()        # level i reads both outputs of level i - 1
            streams.to_main(h)
            if prep is not None:
                streams.to_main(*prep)
            streams.to_side(x)
            fuse_x = None
            if run_video:
                fuse_x = vm.blocks[i](x, *es, y=h, y_mask=ext, time_n=n, space_f=f, cls_only=cls_only, i2t_prep=prep)
            with streams.side():
                h = tm.encoder.layer[i](h, ext, encoder_hidden_states=x, last_norm=True)[0]
            x = fuse_x
        streams.join(h)
        return x, h, video_prefix

    def infer(self, data, video_only=False, return_embeds=True, task_names=None, ret=None):
        ret = {} if ret is None else ret    # (the reference's mutable default leaks state across calls: SURVEY.md Q10)
        text_data, video_data = data['text'], data['video']
        if task_names is not None:
            self.task_names = task_names
        if self.DUAL_TASK in self.task_names:
            streams.to_side(*[v for v in text_data.values() if torch.is_tensor(v)])
            with streams.side():     # the text tower runs next to the video tower
                text_embeddings = self.compute_text(text_data)
            video_embeddings = self.compute_video(video_data)
            streams.join(text_embeddings)
            if return_embeds:
                ret.update({'text_embeds': text_embeddings, 'video_embeds': video_embeddings})
        if 'ITM' in self.task_names:
            x, h, _ = self._fused_stack(video_data, text_data['input_ids'], text_data['attention_mask'],
                                        video_prefix=data.get('_video_prefix'))
            v = A.LayerNormRowsFn.apply(x[:, 0], self.norm.weight, self.norm.bias, self.norm.eps)
            v = self.pre_logits(v)
            # transform -> pooler (tanh) per modality, then fc on the concatenation (model.py:279-290)
            t_feats = self._chain([self.cross_modal_text_transform, self.cross_modal_text_pooler.dense], [ACT_NONE, ACT_TANH], h[:, 0])
            v_feats = self._chain([self.cross_modal_video_transform, self.cross_modal_video_pooler.dense], [ACT_NONE, ACT_TANH], v)
            cls_feats = torch.cat([t_feats, v_feats], dim=-1)
            ret.update({"cross_attn_itm_logits": self.itm_score(cls_feats)})
        if 'MLM' in self.task_names:
            _, h, prefix = self._fused_stack(video_data, data['text_mlm_ids'], text_data['attention_mask'],
                                             need_video_out=False)
            ret['_video_prefix_mlm'] = prefix   # consumed (and removed) by forward() for the ITM pass
            if 'text_mlm_labels' in data and torch.is_grad_enabled():
                names = A.MlmLossFn.NAMES
                params = [self.get_parameter(nm) for nm in names]
                w = {nm: cache().bf16(p) for nm, p in zip(names, params) if p.dim() == 2}
                loss_sum, count, logits = A.MlmLossFn.apply(w, h, data['text_mlm_labels'], *params)
                ret.update({"cross_attn_mlm_logits": logits, "_mlm_loss_sum": loss_sum, "_mlm_count": count})
            else:
                t = heads._linear(h, self.cross_modal_text_transform)
                ret.update({"cross_attn_mlm_logits": self.mlm_score(t)})
        return ret

    def _chain(self, lins, acts, x):
        params = []
        for lin in lins:
            params += [lin.weight, lin.bias]
        return A.MlpChainFn.apply(A.cfg(acts=acts, has_bias=[True] * len(lins)), [cache().bf16(lin.weight) for lin in lins],
                                  x, *params)

    # ------------------------------------------------------------------------------------------ training step
    @staticmethod
    def _global_mean(loss_sum, count):
        """mean over the GLOBAL valid count with gradients through the local sum only (model.py:411-418 + Q9)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            tot = torch.cat([loss_sum.detach().reshape(1), count.reshape(1)])
            dist.all_reduce(tot)
            return (loss_sum.reshape(()) - loss_sum.detach().reshape(()) + tot[0]) / tot[1].clamp_min(1.0)
        return loss_sum.reshape(()) / count.reshape(()).clamp_min(1.0)

    def forward(self, data, n_embeds, v_embeds, allgather, n_gpu, args, config, loss_egonce, gpu, return_embeds=True,
                task_names='EgoNCE_ITM_MLM'):
        ret, loss_dict = {}, {}
        if 'Feature_Extraction' in task_names:
            return self.compute_video(data['video'])
        if self.training and torch.is_grad_enabled():
            rng.advance(data['video'].device)    # new dropout masks for this step (a kernel: captured with the step)
        if 'EgoNCE' not in task_names:
            raise NotImplementedError("MLM / ITM need the EgoNCE branch's similarities (model.py:420,443; SURVEY.md Q11)")

        ret = self.infer(data, task_names='EgoNCE', ret=ret)
        video_embeds, text_embeds = ret['video_embeds'], ret['text_embeds']
        if config['loss']['type'] != 'EgoNCE' or not getattr(loss_egonce, 'fused_kernel', False):
            raise NotImplementedError("only loss type 'EgoNCE' (egovlpv2_b200.model.loss.EgoNCE) is on the pre-training path")
        bsz = video_embeds.shape[0]
        rank = getattr(args, 'rank', 0)
        v_all = allgather(video_embeds.detach().float(), n_gpu, args)
        t_all = allgather(text_embeds.detach().float(), n_gpu, args)
        n_all = allgather(n_embeds.float(), n_gpu, args)
        vb_all = allgather(v_embeds.float(), n_gpu, args)
        temp = loss_egonce.temperature
        loss, sim, mask = A.EgoNceFn.apply(text_embeds, video_embeds, t_all, v_all, n_all, vb_all, temp, rank * bsz)
        mask_bool = mask.bool()
        ret.update({"sim_v2t": sim, "sim_t2v": sim.t()})
        loss_dict.update({'EgoNCE': loss})

        if 'MLM' in task_names:
            ret = self.infer(data, task_names='MLM', ret=ret)
            if "_mlm_loss_sum" not in ret:   # forward() under torch.no_grad(): infer() produced logits only
                lg = ret["cross_attn_mlm_logits"]
                ls, cnt = A.XentFn.apply(lg.reshape(-1, lg.shape[-1]).float(), data['text_mlm_labels'].reshape(-1))
                ret.update({"_mlm_loss_sum": ls, "_mlm_count": cnt})
            loss_mlm = self._global_mean(ret.pop("_mlm_loss_sum"), ret.pop("_mlm_count"))
            loss = loss + loss_mlm
            loss_dict.update({"loss_mlm": loss_mlm})

        prefix_mlm = ret.pop('_video_prefix_mlm', None)
        if 'ITM' in task_names:
            data_itm, itm_labels = self._build_itm_batch(data, sim, mask_bool, temp, rank, allgather, n_gpu, args)
            if self.share_itm_prefix and prefix_mlm is not None:
                # label-1 rows keep their own clip: their unfused-block activations equal the MLM pass's.  Recompute
                # only the label-0 rows (a fixed bsz - bsz//2 of them: static shapes, CUDA-graph safe).
                n_neg = bsz - bsz // 2
                order = torch.argsort(itm_labels, stable=True)            # label-0 rows first, then label-1 rows
                redo_idx, keep_idx = order[:n_neg], order[n_neg:]

                def itm_prefix(video=data_itm['video']):
                    x_redo = self._video_prefix(video.index_select(0, redo_idx))
                    x_cat = torch.cat([x_redo, prefix_mlm.index_select(0, keep_idx)], 0)
                    return x_cat.index_select(0, torch.argsort(order))

                data_itm['_video_prefix'] = itm_prefix   # evaluated inside _fused_stack, after the text prefix is enqueued
            ret = self.infer(data_itm, task_names='ITM', ret=ret)
            loss_sum, count = A.XentFn.apply(ret["cross_attn_itm_logits"], itm_labels.long())
            loss_itm = self._global_mean(loss_sum, count)
            loss = loss + 2 * loss_itm
            loss_dict.update({"loss_itm": loss_itm})

        loss_dict.update({"loss_total": loss})
        return loss, loss_dict, ret

    def _build_itm_batch(self, data, sim, mask_bool, temp, rank, allgather, n_gpu, args):
        """model.py:426-468.  Half the rows keep their pair (label 1); a label-0 row swaps either its clip (hard negative
        drawn from softmax(sim^T/temp) with EgoNCE positives zeroed) or its caption (from softmax(sim/temp))."""
        video = data['video']
        ids, am = data['text']['input_ids'], data['text']['attention_mask']
        bsz, dev = video.shape[0], video.device
        world = getattr(args, 'world_size', 1) if (dist.is_available() and dist.is_initialized()) else 1
        if world > 1:
            # The reference gathers the raw fp32 clips of every rank (model.py:430) to use at most B/2 of them.  The patch
            # embedding rounds every pixel to bf16 before its GEMM (egv_patchify), and bf16 -> fp32 -> bf16 is the
            # identity, so gathering the clips in bf16 gives bit-identical ITM inputs at half the bytes (uint8 frames
            # travel as they are).
            vg = video.to(torch.bfloat16) if video.dtype == torch.float32 else video
            all_video, all_ids, all_am = allgather(vg, n_gpu, args), allgather(ids, n_gpu, args), allgather(am, n_gpu, args)
        else:
            all_video, all_ids, all_am = video, ids, am
        rows = slice(bsz * rank, bsz * (rank + 1))
        own = torch.arange(bsz * rank, bsz * (rank + 1), device=dev)
        if self.itm_plan is not None:
            labels = self.itm_plan["labels"].to(dev).float()
            swap_video = self.itm_plan["swap_video"].to(dev).bool()
            neg = self.itm_plan["neg_idx"].to(dev).long()
            neg_v, neg_t = neg, neg
        else:
            pos_len = bsz // 2
            labels = torch.cat([torch.ones(pos_len, device=dev), torch.zeros(bsz - pos_len, device=dev)])
            labels = labels[torch.randperm(bsz, device=dev)]
            with torch.no_grad():
                w_v2t = F.softmax(sim[rows, :] / temp, dim=1).masked_fill(mask_bool[rows, :], 0)
                w_t2v = F.softmax(sim.t()[rows, :] / temp, dim=1).masked_fill(mask_bool[rows, :], 0)
                neg_v = torch.multinomial(w_t2v + 1e-9, 1).squeeze(1)
                neg_t = torch.multinomial(w_v2t + 1e-9, 1).squeeze(1)
            swap_video = torch.rand(bsz, device=dev) > 0.5
        is_neg = labels == 0
        vid_idx = torch.where(is_neg & swap_video, neg_v, own)
        txt_idx = torch.where(is_neg & ~swap_video, neg_t, own)
        vid_sel = all_video.index_select(0, vid_idx)
        if vid_sel.dtype != video.dtype:
            vid_sel = vid_sel.to(video.dtype)
        data_itm = {'video': vid_sel,
                    'text': {'input_ids': all_ids.index_select(0, txt_idx), 'attention_mask': all_am.index_select(0, txt_idx)}}
        return data_itm, labels

    # ------------------------------------------------------------------------------------------ checkpoints
    def _inflate_positional_embeds(self, new_state_dict):
        """model.py:532-574: load a checkpoint trained with a different number of frames."""
        curr_keys = list(self.state_dict().keys())
        key = 'video_model.temporal_embed'
        if key in new_state_dict and key in curr_keys:
            load_temporal_embed = new_state_dict[key]
            load_num_frames = load_temporal_embed.shape[1]
            curr_num_frames = self.video_params['num_frames']
            embed_dim = load_temporal_embed.shape[2]
            if load_num_frames != curr_num_frames:
                if load_num_frames > curr_num_frames:
                    new_temporal_embed = load_temporal_embed[:, :curr_num_frames, :]
                elif self.load_temporal_fix == 'zeros':
                    new_temporal_embed = torch.zeros([load_temporal_embed.shape[0], curr_num_frames, embed_dim])
                    new_temporal_embed[:, :load_num_frames] = load_temporal_embed
                elif self.load_temporal_fix in ['interp', 'bilinear']:
                    mode = 'bilinear' if self.load_temporal_fix == 'bilinear' else 'nearest'
                    kw = dict(align_corners=True) if mode == 'bilinear' else {}
                    new_temporal_embed = F.interpolate(load_temporal_embed.unsqueeze(0), (curr_num_frames, embed_dim),
                                                       mode=mode, **kw).squeeze(0)
                else:
                    raise NotImplementedError
                new_state_dict[key] = new_temporal_embed
        key = 'video_model.pos_embed'
        if key in new_state_dict and key in curr_keys:
            if new_state_dict[key].shape[1] != self.state_dict()[key].shape[1]:
                raise NotImplementedError(
                    'Loading models with different spatial resolution / patch number not yet implemented, sorry.')
        return new_state_dict


def sim_matrix(a, b, eps=1e-8):
    """model.py:576-584 (stand-alone utility for evaluation code; the training step uses the fused EgoNCE kernel)."""
    a_n, b_n = a.norm(dim=1)[:, None], b.norm(dim=1)[:, None]
    a_norm = a / torch.max(a_n, eps * torch.ones_like(a_n))
    b_norm = b / torch.max(b_n, eps * torch.ones_like(b_n))
    return torch.mm(a_norm, b_norm.transpose(0, 1))


def sim_matrix_batch_val(a, b, eps=1e-8):
    """model.py:587-595"""
    a_n, b_n = a.norm(dim=-1).unsqueeze(-1), b.norm(dim=-1).unsqueeze(-1)
    a_norm = a / torch.max(a_n, eps * torch.ones_like(a_n))
    b_norm = b / torch.max(b_n, eps * torch.ones_like(b_n))
    return torch.bmm(a_norm, b_norm.transpose(1, 2))
