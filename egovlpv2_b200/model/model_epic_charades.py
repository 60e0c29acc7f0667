"""B200 drop-in for EgoVLPv2/model/model_epic_charades.py: the fine-tuning `FrozenInTime` of the EPIC-Kitchens-100 MIR and
Charades-Ego mains (multinode_train_epic.py, multinode_train_charades.py).

Same towers, kernels and state_dict prefix as model.FrozenInTime; what differs (reference lines of model_epic_charades.py):
  * `video_params["drop_path_rate"]` is a required key (:83); every shipped config sets it to 0.0 (configs/ft/*.json:13),
    where timm's DropPath is the identity -- a non-zero rate raises NotImplementedError here;
  * projection 'minimal' = txt_proj Sequential(ReLU(), Linear(C, 256)), vid_proj Sequential(Linear(C, 256)) (:118-119),
    i.e. state_dict keys txt_proj.1.{weight,bias}, vid_proj.0.{weight,bias};
  * the dual-encoder branch of infer() is named 'Dual' (:218);
  * forward(data, allgather, n_gpu, args, config, loss_dual, gpu, ..., task_names='Dual', dataset_name) (:408-445):
    all-gather of the embeddings, sim_matrix(text, video), then `loss_dual(sim, relation)` (dataset 'epic',
    AdaptiveMaxMarginRankingLoss) or `loss_dual(sim)` (dataset 'charades', NormSoftmaxLoss) -- here ONE fused kernel
    sequence (egv_dual_loss) that also emits this rank's slice of the embedding gradients, which is exactly the backward
    of the reference's AllGather_multi (trainer_epic.py:21-41)."""
import torch
from torch import nn

from .. import autograd as A
from .model import FrozenInTime as _PretrainFrozenInTime
from .model import config, sim_matrix, sim_matrix_batch_val, state_dict_data_parallel_fix  # noqa: F401  (module surface)


class FrozenInTime(_PretrainFrozenInTime):
    DUAL_TASK = 'Dual'

    def __init__(self, video_params, text_params, projection_dim=4096, load_checkpoint=None, projection='minimal',
                 load_temporal_fix='bilinear', config=config, task_names='EgoNCE_ITM_MLM', norm_layer=None, embed_dim=768):
        self.drop_path_rate = video_params["drop_path_rate"]      # KeyError when absent, like the reference (:83)
        if self.drop_path_rate:
            raise NotImplementedError("stochastic depth (drop_path_rate > 0) is not used by any shipped config "
                                      "(configs/ft/epic.json:13, configs/ft/charades.json:13)")
        super().__init__(video_params, text_params, projection_dim=projection_dim, load_checkpoint=load_checkpoint,
                         projection=projection, load_temporal_fix=load_temporal_fix, config=config, task_names=task_names,
                         norm_layer=norm_layer, embed_dim=embed_dim)

    def _build_projections(self, projection, txt_dim, ftr_dim, projection_dim):
        """model_epic_charades.py:116-138 (projection_dim is ignored by the reference: both heads are 256 wide)"""
        if projection == 'minimal':
            return nn.Sequential(nn.ReLU(), nn.Linear(txt_dim, 256)), nn.Sequential(nn.Linear(ftr_dim, 256))
        if projection == '':
            return nn.Identity(), nn.Identity()
        raise NotImplementedError

    def create_mask_relevancy(self, relevancy_mat, item_v, item_t):
        """model_epic_charades.py:398-406"""
        mask_rel = torch.zeros(len(item_v), len(item_t))
        item_t_list = item_t.tolist()
        for _i in range(len(item_v)):
            current_item_v = int(item_v[_i].item())
            mask_rel[current_item_v] = torch.as_tensor(relevancy_mat[current_item_v][item_t_list] > 0.1)
        return mask_rel

    def forward(self, data, allgather, n_gpu, args, config, loss_dual, gpu, return_embeds=True, task_names='Dual',
                dataset_name='charades'):
        ret, loss_dict = {}, {}
        loss = None
        if 'Dual' in task_names:
            if dataset_name not in ('epic', 'charades'):
                raise NameError()
            if not getattr(loss_dual, 'fused_kernel', False) or not hasattr(loss_dual, 'fused_kind'):
                raise NotImplementedError("loss_dual must be one of egovlpv2_b200.model.loss.{NormSoftmaxLoss, "
                                          "MaxMarginRankingLoss, AdaptiveMaxMarginRankingLoss}")
            ret = self.infer(data, task_names='Dual', ret=ret)
            video_embeds, text_embeds = ret['video_embeds'], ret['text_embeds']
            bsz, rank = video_embeds.shape[0], getattr(args, 'rank', 0)
            v_all = allgather(video_embeds.detach().float(), n_gpu, args)
            t_all = allgather(text_embeds.detach().float(), n_gpu, args)
            w_all = None
            if dataset_name == 'epic':
                w_all = allgather(data['relation'].float().reshape(-1), n_gpu, args)
            elif loss_dual.fused_kind == 2:
                raise TypeError("AdaptiveMaxMarginRankingLoss needs data['relation'] (dataset_name='epic')")
            loss, output = A.DualLossFn.apply(text_embeds, video_embeds, t_all, v_all, w_all, loss_dual.fused_kind,
                                              float(loss_dual.fused_param), bool(loss_dual.fix_norm), rank * bsz)
            if dataset_name == 'epic':
                ret.update({"sim_v2t": output, "sim_t2v": output.t(), 'epic_relation': w_all})
            else:
                ret.update({"sim_v2t": output, "sim_t2v": output.t()})
            loss_dict.update({'Dual': loss})
        return loss, loss_dict, ret
