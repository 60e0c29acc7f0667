"""B200 drop-in for the pre-training loss of EgoVLPv2/model/loss.py: EgoNCE (loss.py:33-61)."""
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
