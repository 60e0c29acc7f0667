"""B200 drop-in for EgoVLPv2/model/loss.py: EgoNCE (loss.py:33-61) for pre-training; NormSoftmaxLoss, MaxMarginRankingLoss,
AdaptiveMaxMarginRankingLoss (loss.py:13-31, 65-143) for the fine-tuning dual-encoder path; CrossEntropy (loss.py:145-151)."""
import torch
from torch import nn


class EgoNCE(nn.Module):
    """Same constructor and call contract as the reference: forward(x, mask_v, mask_n) -> (loss, mask_bool, temperature)
    on a precomputed similarity matrix.  FrozenInTime.forward recognises this class and runs the fused
    sim_matrix + EgoNCE kernel on the gathered embeddings instead (egovlpv2_b200.autograd.EgoNceFn)."""

    fused_kernel = True

    def __init__(self, temperature=0.05, noun=True, verb=True):
        super().__init__()
        if not (noun and verb):
            raise NotImplementedError("the pre-training configs use noun=True, verb=True (configs/pt/egoclip.json)")
        self.noun, self.verb, self.temperature = noun, verb, temperature

    def forward(self, x, mask_v, mask_n):
        """Stand-alone evaluation on an existing similarity matrix (not the training hot path): plain tensor algebra,
        kept for callers that use the loss module directly."""
        mask = mask_v * mask_n + torch.eye(x.shape[0], device=x.device, dtype=x.dtype)
        mask_bool = mask > 0
        i_sm = torch.softmax(x / self.temperature, dim=1)
        j_sm = torch.softmax(x.t() / self.temperature, dim=1)
        loss_i = torch.log(torch.sum(i_sm * mask_bool, dim=1)).mean()
        loss_j = torch.log(torch.sum(j_sm * mask_bool, dim=1)).mean()
        return -loss_i - loss_j, mask_bool, self.temperature


class _DualLoss(nn.Module):
    """Losses of the fine-tuning ('Dual') path.  `fused_kind` / `fused_param` / `fix_norm` tell
    model_epic_charades.FrozenInTime.forward which mode of the fused sim_matrix + loss kernel (egv_dual_loss) to run on
    the gathered embeddings; calling the module on an existing similarity matrix evaluates the same formula with plain
    tensor algebra (not the training path)."""

    fused_kernel = True


class NormSoftmaxLoss(_DualLoss):
    """loss.py:13-31 -> (loss, temperature)"""
    fused_kind = 0

    def __init__(self, temperature=0.05):
        super().__init__()
        self.temperature = temperature
        self.fix_norm = True

    @property
    def fused_param(self):
        return self.temperature

    def forward(self, x):
        i_logsm = torch.log_softmax(x / self.temperature, dim=1)
        j_logsm = torch.log_softmax(x.t() / self.temperature, dim=1)
        return -torch.diag(i_logsm).mean() - torch.diag(j_logsm).mean(), self.temperature


def _ranking_terms(x, margin_rows, fix_norm):
    """mean over (i, j) of relu(m_i - (x_ii - x_ij)) and relu(m_i - (x_ii - x_ji)); fix_norm drops i == j (loss.py:72-100)"""
    n = x.shape[0]
    d = torch.diag(x).reshape(n, 1)
    terms = torch.stack([torch.relu(margin_rows - (d - x)), torch.relu(margin_rows - (d - x.t()))])
    if fix_norm:
        return terms[:, ~torch.eye(n, dtype=torch.bool, device=x.device)].mean()
    return terms.mean()


class MaxMarginRankingLoss(_DualLoss):
    """loss.py:65-100"""
    fused_kind = 1

    def __init__(self, margin=0.2, fix_norm=True):
        super().__init__()
        self.margin, self.fix_norm = margin, fix_norm

    @property
    def fused_param(self):
        return self.margin

    def forward(self, x, weight=None):
        return _ranking_terms(x, torch.full((x.shape[0], 1), self.margin, device=x.device, dtype=x.dtype), self.fix_norm)


class AdaptiveMaxMarginRankingLoss(_DualLoss):
    """loss.py:102-143: the margin of row i is weight[i] * margin (weight = data['relation'], EpicKitchens_MIR_dataset.py:176)"""
    fused_kind = 2

    def __init__(self, margin=0.4, fix_norm=True):
        super().__init__()
        self.margin, self.fix_norm = margin, fix_norm

    @property
    def fused_param(self):
        return self.margin

    def forward(self, x, weight=None):
        return _ranking_terms(x, weight.reshape(-1, 1).to(x.dtype) * self.margin, self.fix_norm)


class CrossEntropy(nn.Module):
    """loss.py:145-151"""

    def __init__(self):
        super().__init__()
        self.loss = nn.CrossEntropyLoss()

    def forward(self, output, target):
        return self.loss(output, target)
