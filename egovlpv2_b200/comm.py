"""Embedding all-gather for the EgoNCE negatives (reference: trainer_egoclip.py:25-41 `AllGather_multi`, a NCCL
all_gather with local-slice backward; called 11x per step from model.py:385-476).

`P2PAllGather` keeps one symmetric buffer per rank (CUDA IPC over NVLink / NVSwitch) and gathers with ONE kernel launch
per call (csrc/comm.cu): every rank stores its rows straight into all peers' buffers and spins on sequence flags.
Both classes are callables with the reference's `allgather(tensor, n_gpu, args)` signature, so `FrozenInTime.forward`
takes either.  Gradients are not routed through the gather: the fused EgoNCE kernel produces the local-slice
gradients directly, which is exactly the reference's backward."""
import torch
import torch.distributed as dist

from . import lib as _lib


class NcclAllGather:
    def __call__(self, tensor, n_gpu=None, args=None):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return tensor
        out = [torch.empty_like(tensor) for _ in range(dist.get_world_size())]
        dist.all_gather(out, tensor.contiguous())
        return torch.cat(out, 0)


class P2PAllGather:
    SLOT_BYTES = 1 << 20   # per rank and call: fits [64, 4096] fp32 embeddings; larger tensors go through NCCL

    def __init__(self, device):
        assert dist.is_initialized()
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = device
        K = _lib.kernels()
        self.K = K
        self.nccl = NcclAllGather()
        nslots = 2 * self.world * self.SLOT_BYTES          # two parities (consecutive gathers alternate)
        self.slots_ptr, h_slots = K.p2p_alloc(nslots)
        self.flags_ptr, h_flags = K.p2p_alloc(32 * 4)
        handles = [None] * self.world
        dist.all_gather_object(handles, (h_slots, h_flags))
        self.peer_slots, self.peer_flags = [], []
        for r, (hs, hf) in enumerate(handles):
            if r == self.rank:
                self.peer_slots.append(self.slots_ptr)
                self.peer_flags.append(self.flags_ptr)
            else:
                self.peer_slots.append(K.p2p_open(hs))
                self.peer_flags.append(K.p2p_open(hf))
        dist.barrier()

    def check(self):
        """Raise if a gather of this rank ever gave up waiting for a peer (the kernel leaves a sticky error word instead of
        trapping: the CUDA context survives).  Synchronous 4-byte read: call it where the trainer synchronises anyway."""
        err = self.K.p2p_error(self.flags_ptr)
        if err:
            raise RuntimeError("P2P all-gather: rank %d timed out waiting for rank %d" % (self.rank, err & 0xFF))

    def __call__(self, tensor, n_gpu=None, args=None):
        t = tensor.contiguous()
        nbytes = t.numel() * t.element_size()
        if nbytes > self.SLOT_BYTES or nbytes % 16 or t.data_ptr() % 16:
            return self.nccl(t)
        out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        self.K.p2p_allgather(t, nbytes, self.SLOT_BYTES, self.peer_slots, self.peer_flags, self.rank, self.world, out)
        return out
