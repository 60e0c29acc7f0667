"""Fine-tuning dual-encoder step (configs/ft/epic.json: 16 frames, batch 16 per GPU,
AdaptiveMaxMarginRankingLoss; --dataset charades: NormSoftmaxLoss) on one B200: forward + backward + fused AdamW, replayed
as a CUDA graph, timed with CUDA events.  Prints one JSON line (not the driver's bench contract: see bench.py)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dataset", default="epic", choices=["epic", "charades"])
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--seq", type=int, default=32)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    from egovlpv2_b200 import lib as L
    from egovlpv2_b200.model import loss as Lm
    from egovlpv2_b200.synthetic import synthetic_batch
    from egovlpv2_b200.trainer import FinetuneStep, build_dual_model, randomize_gates
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    K = L.kernels()
    torch.manual_seed(0)
    model = build_dual_model(T=a.frames)
    randomize_gates(model)
    model.eval()
    loss_fn = Lm.AdaptiveMaxMarginRankingLoss(margin=0.2) if a.dataset == "epic" else Lm.NormSoftmaxLoss()
    step = FinetuneStep(model, dev, loss_fn, dataset_name=a.dataset, max_steps=10000, warmup_steps=100)
    host = synthetic_batch(a.batch, a.frames, 224, a.seq, seed=1234, pin=True)
    host = {k: host[k] for k in ("video", "input_ids", "attention_mask")}
    host["relation"] = torch.rand(a.batch, generator=torch.Generator().manual_seed(3)).pin_memory()
    batch = step.to_device(host)
    for _ in range(max(a.warmup, 3)):
        step.step(batch)
    step.capture(batch, warmup=1)
    for _ in range(2):
        step.step_graph(batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss, _ = step.step_graph(batch)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    # algorithmic FLOPs: one dual-encoder pass (SURVEY.md 8(d) `pass_nce` with 256-wide projections), fwd + bwd = 3x
    C, N, Nf, S, T = 768, a.frames * 196 + 1, 196, a.seq, a.frames
    vblk = 32 * N * C * C + 4 * (N - 1) * C * ((T + 1) + (Nf + 1)) + 8 * N * C
    tblk = 24 * S * C * C + 4 * S * S * C
    fwd = 2 * T * Nf * 768 * C + 12 * vblk + 12 * tblk + 2 * 2 * C * 256
    flops = 3.0 * a.batch * fwd
    print(json.dumps({"metric": "finetune_dual_clips_per_sec", "value": a.batch / (ms * 1e-3), "unit": "clips/s",
                      "ms_per_step": ms, "config": {"workload": "model_epic_charades.FrozenInTime, task 'Dual', dataset %s, "
                                                   "%d frames 224^2, seq=%d, batch %d, fwd+bwd+AdamW, CUDA graph"
                                                   % (a.dataset, a.frames, a.seq, a.batch)},
                      "step_tflop_algorithmic": round(flops / 1e12, 2),
                      "achieved_tflops": round(flops / (ms * 1e-3) / 1e12, 1),
                      "gpu_launches_per_step": step.launches_per_step, "last_loss": float(loss)}))


if __name__ == "__main__":
    main()
