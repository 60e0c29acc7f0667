#!/usr/bin/env python
"""bench.py -- pre-train clips/s for one EgoNCE+MLM+ITM step (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--frames T]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE configs[2]: TimeSformer-B/16 + RoBERTa-base, 16 frames 224^2, 32 tokens, per-GPU
batch 8, fusion ON (top-6 cross-attention layers), EgoNCE+MLM+ITM, forward + backward + AdamW; synthetic inputs and
random-init weights of the reference architecture (fusion gates / time attention given non-zero values).

One JSON line on rank 0.  `value` = whole-job clips/s with the batch resident in HBM; `e2e` = the same through the
public API with the host->device copy of the batch from pinned memory and a device->host read of the loss inside the
timed region.  `roofline` describes the dominant kernel family (the tcgen05 GEMM): algorithmic FLOPs of all its launches
in one step / their summed CUDA-event durations (instrumented extra step, not the timed one).  `cpu_baseline` = the
oracle port of the reference's PyTorch path on the host cores over a bounded sample.
`--impl reference` times that CPU port alone (the reference itself is pure Python under /root/reference, which does
not exist on the GPU box)."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pretrain_clips_per_sec"
UNIT = "clips/s"


# model / input dimensions per BASELINE.json config; cfg3 is the headline workload (the default), cfg5 the stress case
CONFIGS = {
    "cfg3": dict(name="TimeSformer-B/16 + RoBERTa-base", img=224, patch=16, C=768, heads=12, depth=12, n_fuse=6),
    "cfg5": dict(name="TimeSformer-L/14 + RoBERTa-large", img=336, patch=14, C=1024, heads=16, depth=24, n_fuse=6),
}


def workload_name(a):
    m = CONFIGS[a.config]
    return ("%s, %d frames %d^2, seq=%d, per-GPU batch %d, fusion ON (top-6), EgoNCE+MLM+ITM, fwd+bwd+AdamW"
            % (m["name"], a.frames, m["img"], a.seq, a.batch))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1400.0), d.get("bf16_tflops", 1590.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.QUERY,
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            sm.sort()
            out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


# ----------------------------------------------------------------------------------------------- CPU port (oracle)
class CpuPort:
    """The oracle (oracle/egovlp_oracle.py: fp32 functional restatement of the reference's PyTorch path) timed on the
    host cores: forward + backward of one EgoNCE+MLM+ITM step on `sample_batch` clips of the workload."""

    def __init__(self, a, sample_batch):
        from oracle import egovlp_oracle as O
        self.O = O
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        shapes = O.key_shapes(T=a.frames)
        self.sd = {k: v.requires_grad_(True) for k, v in O.seeded_state(shapes, seed=0).items()}
        self.data = O.synthetic_batch(sample_batch, a.frames, 224, a.seq, seed=1234)
        self.plan = O.synthetic_itm_plan(sample_batch)
        self.n = sample_batch

    def step(self):
        t0 = time.perf_counter()
        out = self.O.pretrain_step(self.data, self.sd, 12, 12, 6, self.plan)
        out["loss_total"].backward()
        dt = time.perf_counter() - t0
        for v in self.sd.values():
            v.grad = None
        return dt


def cpu_port_clips_per_sec(a, sample_batch):
    port = CpuPort(a, sample_batch)
    dt = port.step()
    return sample_batch / dt, port.cores, dt


def run_reference(a, rank):
    if rank != 0:
        return
    # every one of the K + W steps is a bounded sample: 3 clips (~11 s on 16 host threads) keep the driver's
    # `--steps 20 --warmup 5` run at ~5 minutes; --cpu-sample 8 gives the workload's own batch (~38 s per step)
    sample = a.cpu_sample if a.cpu_sample > 0 else 3
    port = CpuPort(a, sample)
    times = []
    for i in range(a.warmup + a.steps):
        dt = port.step()
        if i >= a.warmup:
            times.append(dt)
    mean_dt = sum(times) / len(times)
    val, cores = sample / mean_dt, port.cores
    desc = "B=%d clips of the workload per step, fwd+bwd (no optimizer), fp32 oracle port, %d torch threads" % (sample, cores)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * mean_dt, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "sample": desc},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(a, rank, world, local_rank):
    import torch.distributed as dist
    from egovlpv2_b200 import lib as L
    from egovlpv2_b200.synthetic import synthetic_batch
    from egovlpv2_b200.trainer import PretrainStep, build_model, randomize_gates, step_flops

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    K = L.kernels()   # raises if the CUDA library is missing: there is no fallback
    torch.manual_seed(0)
    m = CONFIGS[a.config]
    model = build_model(T=a.frames, img=m["img"], C=m["C"], heads=m["heads"], depth=m["depth"], n_fuse=m["n_fuse"], patch=m["patch"])
    randomize_gates(model)
    model.train()     # the reference's step runs in train mode: text-tower dropout (p = 0.1) is part of the timed work
    use_graph = not a.no_graph
    step = PretrainStep(model, dev, max_steps=10000, warmup_steps=100, gather=a.gather)
    host = synthetic_batch(a.batch, a.frames, m["img"], a.seq, seed=1234 + rank, pin=True)
    if a.input_dtype == "u8":
        # decoder-format frames: the loader's / 255 + NormalizeVideo run inside the im2col kernel,
        # the H2D copy carries 1 byte per sample instead of 4
        gen = torch.Generator().manual_seed(99 + rank)
        host["video"] = torch.randint(0, 256, tuple(host["video"].shape), generator=gen, dtype=torch.uint8).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    dev_batch = step.to_device(host)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, e2e):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        last = None
        if e2e and use_graph:
            step.prefetch(host)               # H2D of step 0's inputs (exposed: nothing to overlap with yet)
        for i in range(n):
            if use_graph and e2e:
                # every step's inputs travel pinned host -> device inside the timed region; the copy of step i+1 is
                # issued on a copy stream before step i is replayed, so it overlaps step i's kernels (input prefetch)
                loss, _ = step.step_graph_prefetched()
                if i + 1 < n:
                    step.prefetch(host)
            elif use_graph:
                loss, _ = step.step_graph(dev_batch)
            else:
                b = step.to_device(host) if e2e else dev_batch
                loss, _ = step.step(b)
            if e2e:
                last = float(loss.item())     # D2H read of the step's result
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if e2e:
            ms = max(ms, (time.perf_counter() - t0) * 1e3)   # host-visible time bounds the end-to-end number
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, last

    def note(msg):
        if os.environ.get("BENCH_VERBOSE"):
            torch.cuda.synchronize()
            print("[bench rank %d] %s" % (rank, msg), file=sys.stderr, flush=True)

    note("model built, gather=%s" % step.gather_kind)
    for i in range(max(a.warmup, 3)):
        step.step(dev_batch)
        note("eager warm-up step %d done" % i)
    # instrumented extra step: per-launch CUDA events -> GEMM / attention time and work (not part of the timed region;
    # run before the graph capture: its eager allocations and the graph's private pool do not fit side by side at cfg 5)
    # (every rank runs it -- a step contains collectives -- but only rank 0 keeps the timings)
    prof = None
    torch.cuda.synchronize()
    from egovlpv2_b200 import streams
    two_streams = streams.enabled()
    streams.enable(False)    # per-launch events must not time kernels that share the GPU with the other stream
    if rank == 0:
        K.start_profile()
    step.step(dev_batch)
    torch.cuda.synchronize()
    if rank == 0:
        prof = K.stop_profile()
    streams.enable(two_streams)
    if world > 1:
        dist.barrier()
    if use_graph:
        import gc
        gc.collect()
        torch.cuda.empty_cache()   # the eager warm-up's cached blocks would sit next to the graph's private pool (cfg 5: 57 + 119 GB)
        step.capture(dev_batch, warmup=1)
        note("graph captured")
        for _ in range(2):
            step.step_graph(dev_batch)
        # warm the end-to-end path as well (untimed): its staging buffers and copy stream are created on first use, and
        # the allocator's cache was just emptied -- that one-time cudaMalloc does not belong in the timed region
        step.prefetch(host)
        step.step_graph_prefetched()
        torch.cuda.synchronize()
        note("graph replays ok")
    launches0 = K.launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms, _ = timed(a.steps, e2e=False)
    launches = (step.launches_per_step * a.steps) if use_graph else (K.launch_count() - launches0)
    ms_e2e, last_loss = timed(a.steps, e2e=True)
    clocks = sampler.stop() if sampler else None

    if rank != 0:
        return

    clips = a.batch * world * a.steps
    value = clips / (ms * 1e-3)
    e2e_value = clips / (ms_e2e * 1e-3)
    flops, parts = step_flops(a.batch, a.frames, img=m["img"], patch=m["patch"], S=a.seq, C=m["C"], depth=m["depth"], n_fuse=m["n_fuse"])
    executed = sum(w for (t, w, n) in prof.values())      # FLOPs of the products actually issued (exact shortcuts Q6 / Q7 and
    #                                                        the re-associated cross-attention make it smaller than `flops`)
    ref_gpu = None
    rp = os.path.join(ROOT, "profiles", "r02_ref_gpu_eager.jsonl")
    if os.path.exists(rp):    # the reference's own 1-GPU PyTorch eager step on this pool's B200 (tools/bench_ref_gpu.py)
        for ln in open(rp):
            try:
                d = json.loads(ln)
            except ValueError:
                continue
            if "cfg 3" in d.get("config", {}).get("workload", "") and d["config"].get("use_checkpoint"):
                ref_gpu = d
    sustained, burst, hbm, peak_src = peaks()
    traffic = None
    import glob
    tps = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_gemm_traffic.json")))
    tp = tps[-1] if tps else None
    if tp:   # dram__bytes_read+write per GEMM launch from the latest committed ncu capture of the same step
        traffic = json.load(open(tp)).get("gemm_dram_bytes_per_launch")
    gemm_t = sum(t for (r, k), (t, w, n) in prof.items() if k == "gemm")
    gemm_w = sum(w for (r, k), (t, w, n) in prof.items() if k == "gemm")
    gemm_n = sum(n for (r, k), (t, w, n) in prof.items() if k == "gemm")
    prof_total = sum(t for (t, w, n) in prof.values())
    achieved = gemm_w / gemm_t / 1e12 if gemm_t > 0 else 0.0

    alg = dict(K.alg)

    def region(prefix):
        """all launches of one gated cross-attention direction in the instrumented step (text-side preparation included).
        `tflops_algorithmic` divides the FLOPs of the REFERENCE formulation (B (4NC^2 + 4SC_tC + 4NSC) per
        block forward, twice that backward) by the time -- BASELINE.json's second metric; `tflops_executed` counts the
        re-associated products actually issued (about half)."""
        t = sum(t for (r, k), (t, w, n) in prof.items() if r.startswith(prefix))
        w = sum(w for (r, k), (t, w, n) in prof.items() if r.startswith(prefix))
        a_ = sum(v for r, v in alg.items() if r.startswith(prefix))
        return {"ms": round(t * 1e3, 3), "tflops_executed": round(w / t / 1e12, 1) if t > 0 else 0.0,
                "tflops_algorithmic": round(a_ / t / 1e12, 1) if t > 0 else 0.0,
                "frac_of_burst_peak": round(a_ / t / 1e12 / burst, 4) if t > 0 else 0.0}

    cpu_val, cores, cpu_dt = (None, None, None)
    cpu_desc = None
    if world == 1 and not a.no_cpu_baseline and a.config == "cfg3":   # (the fp32 port of a cfg-5 clip would take many minutes)
        n_cpu = a.cpu_sample if a.cpu_sample > 0 else a.batch     # default: the workload's own per-GPU batch, one step
        cpu_val, cores, cpu_dt = cpu_port_clips_per_sec(a, n_cpu)
        cpu_desc = ("B=%d clips of the workload, 1 step fwd+bwd (no optimizer), fp32 oracle port, %d torch threads, %.1f s"
                    % (n_cpu, cores, cpu_dt))
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": workload_name(a), "global_batch": a.batch * world, "parallelism": "dp%d" % world,
                   "l2_policy": "per-step working set (>30 GB of activations) exceeds the 126 MB L2; no flush needed",
                   "embedding_gather": step.gather_kind, "cuda_graph": use_graph, "input_dtype": a.input_dtype, "text_tower_side_stream": two_streams, "step_tflop_algorithmic": round(flops / 1e12, 2),
                   "step_frac_of_sustained_peak": round(flops / (ms / a.steps * 1e-3) / 1e12 / sustained, 4),
                   "step_tflop_executed": round(executed / 1e12, 2),
                   "step_frac_of_sustained_peak_executed": round(executed / (ms / a.steps * 1e-3) / 1e12 / sustained, 4),
                   "train_mode": True,
                   "reference_gpu_eager": None if ref_gpu is None or world != 1 or a.batch != 8 or a.frames != 16 or a.config != "cfg3" else {
                       "clips_per_sec": round(ref_gpu["value"], 3), "ms_per_step": round(ref_gpu["ms_per_step"], 1),
                       "what": "UNMODIFIED reference, 1 B200 of this pool, fp16 autocast + GradScaler, yaml defaults "
                               "(use_checkpoint: True), DDP static_graph, torch AdamW -- tools/bench_ref_gpu.py, "
                               "profiles/r02_ref_gpu_eager.jsonl",
                       "value_over_reference_gpu_eager": round(value / ref_gpu["value"], 3),
                       "e2e_over_reference_gpu_eager": round(e2e_value / ref_gpu["value"], 3)},
                   "xattn_i2t_fwd": region("xattn_i2t_fwd"), "xattn_t2i_fwd": region("xattn_t2i_fwd"),
                   "xattn_i2t_bwd": region("xattn_i2t_bwd"), "xattn_t2i_bwd": region("xattn_t2i_bwd"),
                   "gemm_share_of_kernel_time": round(gemm_t / prof_total, 4) if prof_total else None,
                   "last_loss": last_loss},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / a.steps,
                "input_prefetch": "H2D of step i+1 is issued before step i replays (copy stream); step 0's copy is exposed"},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05, all GEMMs of one step)", "achieved": achieved,
                     "peak": sustained, "unit": "TFLOP/s", "frac": achieved / sustained, "peak_source": peak_src,
                     "launches": gemm_n, "avg_launch_us": gemm_t / max(gemm_n, 1) * 1e6, "traffic": traffic,
                     "traffic_note": "avg DRAM bytes per GEMM launch, ncu capture %s" % (os.path.relpath(tp, ROOT) if tp else None)},
    }
    if cpu_val is not None:
        line["cpu_baseline"] = {"value": cpu_val, "unit": UNIT, "cores": cores, "kind": "port", "sample": cpu_desc}
    emit(line)


def _protect_stdout():
    """The contract is ONE JSON line on stdout; libraries (NCCL's version banner, ...) also write to fd 1.
    Point fd 1 at stderr for everybody else and keep a private handle on the real stdout for the result line."""
    global _RESULT_OUT
    real = os.dup(1)
    os.dup2(2, 1)
    _RESULT_OUT = os.fdopen(real, "w")


_RESULT_OUT = None


def emit(line):
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS),
                    help="BASELINE.json config: cfg3 = the headline workload; cfg5 = TimeSformer-L/14 + RoBERTa-large, 32 frames "
                         "336^2, seq 64, per-GPU batch 2 (bs 16 over 8 GPUs)")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: 8 for cfg3, 2 for cfg5)")
    ap.add_argument("--frames", type=int, default=0)
    ap.add_argument("--seq", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=0, dest="cpu_sample",
                    help="clips per step of the CPU port; 0 = the workload's batch (8) for the cpu_baseline of the default arm "
                         "(one step, ~38 s on 16 host threads) and 3 for every step of --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="auto", choices=["auto", "p2p", "nccl"], help="embedding all-gather implementation")
    ap.add_argument("--no-graph", action="store_true", help="eager kernel launches instead of replaying a captured CUDA graph")
    ap.add_argument("--input-dtype", default="f32", choices=["f32", "u8"], dest="input_dtype",
                    help="video frames as the reference's loader ships them (f32, normalised) or as decoded (uint8)")
    a = ap.parse_args()
    dflt = dict(cfg3=(8, 16, 32), cfg5=(2, 32, 64))[a.config]
    a.batch, a.frames, a.seq = a.batch or dflt[0], a.frames or dflt[1], a.seq or dflt[2]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference(a, rank)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(a, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
