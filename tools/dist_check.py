"""Multi-GPU checks (run under torchrun, one rank per GPU):
  1. the NVSwitch P2P all-gather kernel returns exactly what NCCL all_gather returns (odd number of calls: both slot sets);
  2. multi-rank parity of the pre-training step: every rank's losses equal the single-process losses on the
     concatenated batch, and the rank-summed gradients equal the full-batch gradients."""
import os
import sys
import types

import faulthandler
faulthandler.dump_traceback_later(int(os.environ.get("DIST_CHECK_WATCHDOG_S", "240")), exit=True)   # a hang ends with a traceback, not with the caller's timeout
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from egovlpv2_b200 import lib as L  # noqa: E402
from egovlpv2_b200.comm import NcclAllGather, P2PAllGather  # noqa: E402
from egovlpv2_b200.model.loss import EgoNCE  # noqa: E402
from egovlpv2_b200.synthetic import synthetic_batch  # noqa: E402
from egovlpv2_b200.trainer import build_model, randomize_gates  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
K = L.kernels()
p2p, nccl = P2PAllGather(dev), NcclAllGather()
g = torch.Generator(device="cpu").manual_seed(100 + rank)
for i, shape in enumerate([(8, 4096), (4, 582), (4, 118), (8, 4096), (4, 32), (16, 4096), (8, 4096)]):
    dt = torch.float32 if i != 4 else torch.int64
    t = (torch.randn(shape, generator=g) * 3).to(dt).to(dev)
    a, b = p2p(t), nccl(t)
    torch.cuda.synchronize()
    assert a.shape == b.shape and torch.equal(a, b), ("p2p gather mismatch", i, rank)
if rank == 0:
    print("p2p all-gather == nccl all_gather: ok (%d ranks)" % world)

# ---- step parity
c = dict(C=128, heads=2, depth=8, n_fuse=2, T=2, img=64, S=8, proj=256, vocab=50265)
Bl = 4
torch.manual_seed(0)
model = build_model(T=c["T"], img=c["img"], C=c["C"], heads=c["heads"], depth=c["depth"], n_fuse=c["n_fuse"], vocab=c["vocab"],
                    proj=c["proj"])
randomize_gates(model)
model.eval().to(dev)
for p in model.parameters():
    dist.broadcast(p.data, 0)
full = synthetic_batch(Bl * world, c["T"], c["img"], c["S"], seed=7)
full = {k: v.to(dev) for k, v in full.items()}
gp = torch.Generator().manual_seed(5)
G = Bl * world
labels = torch.cat([torch.ones(Bl // 2), torch.zeros(Bl - Bl // 2)]).repeat(world)
swap = torch.rand(G, generator=gp) > 0.5
neg = (torch.arange(G) + torch.randint(1, G, (G,), generator=gp)) % G


def run(batch, plan, allgather, args):
    model.zero_grad(set_to_none=True)
    model.itm_plan = plan
    data = {"video": batch["video"], "text": {"input_ids": batch["input_ids"], "attention_mask": batch["attention_mask"]},
            "text_mlm_ids": batch["text_mlm_ids"], "text_mlm_labels": batch["text_mlm_labels"]}
    loss, ld, _ = model(data, batch["noun_vec"], batch["verb_vec"], allgather, world, args, {"loss": {"type": "EgoNCE"}},
                        EgoNCE(), local, task_names="EgoNCE_MLM_ITM")
    loss.backward()
    return {k: float(v) for k, v in ld.items()}, {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}


sl = slice(rank * Bl, (rank + 1) * Bl)
loc = {k: v[sl].contiguous() for k, v in full.items()}
ld_r, g_r = run(loc, dict(labels=labels[sl], swap_video=swap[sl], neg_idx=neg[sl]), p2p,
                types.SimpleNamespace(world_size=world, rank=rank))
# single-process reference on the concatenated batch: no cross-rank reduction of the loss sums
model._global_mean = lambda ls, cnt: ls.reshape(()) / cnt.reshape(()).clamp_min(1.0)
ld_f, g_f = run(full, dict(labels=labels, swap_video=swap, neg_idx=neg), lambda t, n=None, a=None: t,
                types.SimpleNamespace(world_size=1, rank=0))
for k in ld_f:
    assert abs(ld_r[k] - ld_f[k]) <= 2e-2 * max(1.0, abs(ld_f[k])), (k, ld_r[k], ld_f[k], rank)
worst = 0.0
names = sorted(g_f)[::7]
gmax = max(g_f[n].norm().item() for n in names)
for n in names:
    s = g_r[n].clone()
    dist.all_reduce(s)
    ref = g_f[n]
    # attention key biases have an analytically ZERO gradient (softmax is invariant to a per-query constant): what both
    # runs hold there is rounding noise ~1e-4 of the other gradients, so errors are measured against a floor
    den = max(ref.norm().item(), 1e-3 * gmax)
    err = (s - ref).norm().item() / den
    worst = max(worst, err)
    assert err <= 0.2, (n, err, ref.norm().item(), gmax)
if rank == 0:
    print("multi-rank step parity ok: losses %s ; worst sampled gradient rel-L2 %.3f" % (ld_r, worst))
# ---- trainer.PretrainStep: bucketed all-reduce overlapped with the backward (reduce.py) == one all-reduce after the backward
from egovlpv2_b200 import weights  # noqa: E402
from egovlpv2_b200.trainer import PretrainStep  # noqa: E402
ref_grad = None
for mode in ("0", "1"):
    os.environ["EGV_OVERLAP_ALLREDUCE"] = mode
    weights.cache().arena = None
    torch.manual_seed(0)
    m2 = build_model(T=c["T"], img=c["img"], C=c["C"], heads=c["heads"], depth=c["depth"], n_fuse=c["n_fuse"], vocab=c["vocab"],
                     proj=c["proj"])
    randomize_gates(m2)
    m2.eval().to(dev)
    for p in m2.parameters():
        dist.broadcast(p.data, 0)
    st = PretrainStep(m2, dev, lr=0.0, weight_decay=0.0)
    m2.itm_plan = dict(labels=labels[sl].to(dev), swap_video=swap[sl].to(dev), neg_idx=neg[sl].to(dev))   # on the device: the step is captured below
    st.step(loc)
    torch.cuda.synchronize()
    gsum = st.opt.arena.grad.clone()
    if mode == "0":
        ref_grad = gsum
    else:
        assert st.reducer.calls > 2 * c["depth"], st.reducer.calls
        # (the backward itself is not bit-reproducible run to run on the GPU -- split-K / bias-gradient atomics -- so the two
        # runs are compared to rounding noise; the 2-rank gloo test on CPU pins bit-equality of the reduction logic)
        assert (gsum - ref_grad).norm().item() <= 1e-5 * ref_grad.norm().item(), (gsum - ref_grad).abs().max().item()
        # and through the captured CUDA graph (the bench path)
        print("[rank %d] capturing the step" % rank, flush=True)
        st.capture(loc, warmup=1)
        print("[rank %d] captured; replaying" % rank, flush=True)
        st.step_graph(loc)
        torch.cuda.synchronize()
        print("[rank %d] replay done" % rank, flush=True)
        g2 = st.opt.arena.grad
        assert (g2 - ref_grad).norm().item() <= 1e-5 * ref_grad.norm().item(), "graph replay of the overlapped reduction differs"
if rank == 0:
    print("overlapped gradient all-reduce == single all-reduce: ok (%d buckets)" % st.reducer.calls)
dist.barrier()
torch.cuda.synchronize()
sys.stdout.flush()
# (no destroy_process_group: tearing the NCCL communicator down while a captured graph that holds its collectives is alive
#  blocks in ncclCommDestroy; the process ends here)
os._exit(0)
