"""Run warm-up steps, then ONE pre-training step between cudaProfilerStart/Stop (for ncu --profile-from-start off)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from egovlpv2_b200.synthetic import synthetic_batch  # noqa: E402
from egovlpv2_b200.trainer import PretrainStep, build_model, randomize_gates  # noqa: E402

B = int(os.environ.get("PROF_BATCH", "8"))
T = int(os.environ.get("PROF_FRAMES", "16"))
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = build_model(T=T)
randomize_gates(model)
model.train()   # the reference's step runs in train mode (text-tower dropout)
step = PretrainStep(model, dev)
batch = step.to_device(synthetic_batch(B, T, 224, 32, seed=1234))
for _ in range(int(os.environ.get("PROF_WARMUP", "2"))):
    step.step(batch)
torch.cuda.synchronize()
torch.cuda.profiler.start()
step.step(batch)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
