"""CPU-only checks: the C-ABI library loads and exports every symbol include/egovlp_b200.h declares (no compute
calls), and the multi-rank host logic (global-mean losses, gather callables, rank-major ITM indexing) on a 2-rank gloo
group."""
import ctypes
import os
import re
import socket
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from egovlpv2_b200 import build, lib
    build.build(verbose=False)
    hdr = open(os.path.join(ROOT, "include", "egovlp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(egv_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 30, names
    so = ctypes.CDLL(lib.LIB_PATH)
    missing = [n for n in names if not hasattr(so, n)]
    assert not missing, missing
    assert so.egv_version() >= 100
    so.egv_last_error.restype = ctypes.c_char_p
    assert isinstance(so.egv_last_error(), bytes)
    # argument validation happens before any CUDA call: a NULL struct is rejected with EGV_ERR_ARG
    assert so.egv_gemm_bf16(None, None) == -1
    assert b"null" in so.egv_last_error()


def test_product_fails_loudly_without_library(monkeypatch):
    from egovlpv2_b200 import lib
    monkeypatch.setattr(lib, "LIB_PATH", os.path.join(ROOT, "does_not_exist.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        lib.Kernels()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "egovlpv2_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in re.sub(r'"""(.|\n)*?"""', "", src), f


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from egovlpv2_b200 import lib as L
        from egovlpv2_b200.comm import NcclAllGather
        from egovlpv2_b200.model.model import FrozenInTime
        from tests.fake_kernels import FakeKernels
        L.set_kernels(FakeKernels())
        # 1. gather callable: rank-major concatenation (trainer_egoclip.py:34)
        g = NcclAllGather()
        t = torch.full((2, 3), float(rank))
        out = g(t, world, None)
        assert out.shape == (2 * world, 3) and torch.equal(out[2 * rank:2 * rank + 2], t)
        # 2. global mean with gradient through the local sum only (model.py:411-418 + Q9)
        ls = torch.tensor([3.0 + rank], requires_grad=True)
        cnt = torch.tensor([2.0 + rank])
        m = FrozenInTime._global_mean(ls, cnt)
        assert abs(m.item() - (3.0 + 4.0) / (2.0 + 3.0)) < 1e-6
        m.backward()
        assert abs(ls.grad.item() - 1.0 / 5.0) < 1e-6
        # 3. ITM batch: label-0 rows pull their negative from the GLOBAL pool with rank-major indexing (model.py:443-468)
        B = 4
        video = torch.arange(B).float().reshape(B, 1, 1, 1, 1) + 100 * rank
        ids = (torch.arange(B).reshape(B, 1) + 100 * rank).long()
        data = {"video": video, "text": {"input_ids": ids, "attention_mask": torch.ones_like(ids)}}
        dummy = types.SimpleNamespace(itm_plan=dict(labels=torch.tensor([1., 0., 0., 1.]),
                                                    swap_video=torch.tensor([False, True, False, False]),
                                                    neg_idx=torch.tensor([0, 5, 6, 0])))
        args = types.SimpleNamespace(world_size=world, rank=rank)
        d_itm, labels = FrozenInTime._build_itm_batch(dummy, data, None, None, 0.05, rank, g, world, args)
        own = torch.arange(B).float() + 100 * rank
        v = d_itm["video"].reshape(B)
        assert v[0] == own[0] and v[3] == own[3] and v[2] == own[2]
        assert v[1] == 100 * 1 + 1                       # global row 5 = rank 1, local row 1
        t_ids = d_itm["text"]["input_ids"].reshape(B)
        assert t_ids[2] == 100 * 1 + 2 and t_ids[1] == ids[1, 0]
        # 4. fine-tuning Dual loss: every rank evaluates the loss on the gathered embeddings and keeps the gradient
        #    slice of its own rows (model_epic_charades.py:416-431; trainer_epic.py:21-41 AllGather_multi backward)
        from egovlpv2_b200 import autograd as A
        gen = torch.Generator().manual_seed(5)
        full_t, full_v = torch.randn(2 * world, 16, generator=gen), torch.randn(2 * world, 16, generator=gen)
        rel_w = torch.rand(2 * world, generator=gen)
        for kind, param in ((0, 0.05), (2, 0.3)):
            lt = full_t[2 * rank:2 * rank + 2].clone().requires_grad_(True)
            lv = full_v[2 * rank:2 * rank + 2].clone().requires_grad_(True)
            w_all = g(rel_w[2 * rank:2 * rank + 2], world, None) if kind == 2 else None
            loss, sim = A.DualLossFn.apply(lt, lv, g(lt.detach(), world, None), g(lv.detach(), world, None), w_all,
                                           kind, param, True, 2 * rank)
            loss.backward()
            ft, fv = full_t.clone().requires_grad_(True), full_v.clone().requires_grad_(True)
            tn, vn = ft / ft.norm(dim=1, keepdim=True), fv / fv.norm(dim=1, keepdim=True)
            x = tn @ vn.t()
            if kind == 0:
                ref = -torch.log_softmax(x / param, 1).diag().mean() - torch.log_softmax(x.t() / param, 1).diag().mean()
            else:
                n = x.shape[0]
                d, m = x.diag().reshape(n, 1), rel_w.reshape(n, 1) * param
                off = ~torch.eye(n, dtype=torch.bool)
                ref = torch.stack([torch.relu(m - (d - x)), torch.relu(m - (d - x.t()))])[:, off].mean()
            ref.backward()
            assert abs(loss.item() - ref.item()) < 1e-5
            assert torch.allclose(sim, x.detach(), atol=1e-6)
            assert torch.allclose(lt.grad, ft.grad[2 * rank:2 * rank + 2], atol=1e-6)
            assert torch.allclose(lv.grad, fv.grad[2 * rank:2 * rank + 2], atol=1e-6)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_two_rank_host_logic_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def _ddp_worker(rank, world, port, q, golden_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from egovlpv2_b200 import functional as Fn
        from egovlpv2_b200 import lib as L
        from egovlpv2_b200.comm import NcclAllGather
        from egovlpv2_b200.model.loss import EgoNCE
        from oracle import egovlp_oracle as O
        from tests.fake_kernels import FakeKernels
        from tests.test_model_cpu import _golden, build_tiny
        L.set_kernels(FakeKernels())
        Fn.BF16 = torch.float32
        torch.set_num_threads(2)
        fx, c, shapes, sd, _, _ = _golden(golden_dir)
        Bl = 2
        G = Bl * world
        full = O.synthetic_batch(G, c["T"], c["img"], c["S"], seed=31)
        plan = dict(labels=torch.tensor([1., 0.] * world), swap_video=torch.tensor([False, True, False, False][:G]),
                    neg_idx=torch.tensor([0, 2, 0, 1][:G]))

        def run(model, batch, pl, gather, args, n):
            inner = model.module if hasattr(model, "module") else model
            inner.itm_plan = pl
            data = {"video": batch["video"], "text": {"input_ids": batch["input_ids"], "attention_mask": batch["attention_mask"]},
                    "text_mlm_ids": batch["text_mlm_ids"], "text_mlm_labels": batch["text_mlm_labels"]}
            # the call of trainer_egoclip.py:145, through the DDP wrapper
            loss, ld, _ = model(data, batch["noun_vec"], batch["verb_vec"], gather, n, args, {"loss": {"type": "EgoNCE"}},
                                EgoNCE(), 0, task_names="EgoNCE_MLM_ITM")
            loss.backward()
            return float(loss.detach()), {k: p.grad.clone() for k, p in inner.named_parameters() if p.grad is not None}

        model = build_tiny(c)
        model.load_state_dict(sd, strict=False)
        model.eval()
        ddp = torch.nn.parallel.DistributedDataParallel(model, static_graph=True)      # base_trainer.py:268-269
        sl = slice(rank * Bl, (rank + 1) * Bl)
        loc = {k: v[sl].contiguous() for k, v in full.items()}
        loss_r, g_r = run(ddp, loc, {k: v[sl] for k, v in plan.items()}, NcclAllGather(),
                          types.SimpleNamespace(world_size=world, rank=rank), world)
        # single-process run of the concatenated batch on a fresh copy (no cross-rank reduction of the loss sums)
        ref = build_tiny(c)
        ref.load_state_dict(sd, strict=False)
        ref.eval()
        ref._global_mean = lambda ls, cnt: ls.reshape(()) / cnt.reshape(()).clamp_min(1.0)
        loss_f, g_f = run(ref, full, plan, lambda t, n=None, a=None: t, types.SimpleNamespace(world_size=1, rank=0), 1)
        assert abs(loss_r - loss_f) <= 1e-4 * max(1.0, abs(loss_f)), (loss_r, loss_f)
        gmax = max(g.norm().item() for g in g_f.values())
        worst = 0.0
        for k, gf in g_f.items():
            # DDP averages over ranks: mean_r(local gradient) == full-batch gradient / world
            err = (g_r[k] * world - gf).norm().item() / max(gf.norm().item(), 1e-3 * gmax)
            worst = max(worst, err)
            assert err <= 2e-3, (k, err)
        q.put((rank, "ok %.1e" % worst))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_two_rank_ddp_static_graph_training_call(golden_dir):
    """The reference's training configuration on CPU: DistributedDataParallel(static_graph=True) (base_trainer.py:268-269)
    around the drop-in FrozenInTime, the trainer's call signature (trainer_egoclip.py:145), an all_gather callable, two
    gloo ranks: DDP-averaged gradients x world == gradients of the concatenated batch in one process."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q, golden_dir)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1].startswith("ok") for r in res), res


def _reducer_worker(rank, world, port, q, golden_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from egovlpv2_b200 import functional as Fn
        from egovlpv2_b200 import lib as L
        from egovlpv2_b200 import weights
        from egovlpv2_b200.trainer import PretrainStep
        from oracle import egovlp_oracle as O
        from tests.fake_kernels import FakeKernels
        from tests.test_model_cpu import _golden, build_tiny
        L.set_kernels(FakeKernels())
        Fn.BF16 = torch.float32
        torch.set_num_threads(2)
        fx, c, shapes, sd, _, _ = _golden(golden_dir)
        batch = O.synthetic_batch(2, c["T"], c["img"], c["S"], seed=50 + rank)
        grads = {}
        for mode in ("1", "0"):
            os.environ["EGV_OVERLAP_ALLREDUCE"] = mode
            weights.cache().arena = None
            model = build_tiny(c)
            model.load_state_dict(sd, strict=False)
            model.eval()
            step = PretrainStep(model, torch.device("cpu"), lr=0.0, weight_decay=0.0, gather="nccl")
            model.itm_plan = dict(labels=torch.tensor([1., 0.]), swap_video=torch.tensor([False, True]), neg_idx=torch.tensor([1 - rank, 2 * (1 - rank)]))
            step.step(batch)
            grads[mode] = step.opt.arena.grad.clone()
            if mode == "1":
                calls, done = step.reducer.calls, step.reducer.done
                assert calls > 2 * c["depth"], calls       # per block / layer buckets, not one call
                covered = sorted(done)
                assert covered[0][0] == 0 and max(hi for _, hi in covered) == step.opt.arena.numel
        # the bucketed, overlapped reduction == one all-reduce of the whole buffer, bit for bit (2 ranks: a + b commutes)
        assert torch.equal(grads["1"], grads["0"]), (grads["1"] - grads["0"]).abs().max().item()
        assert grads["1"].abs().sum().item() > 0
        # opt-in bf16 buckets (EGV_ALLREDUCE_BF16=1): same buckets, payload rounded to bf16 before the sum -- the result
        # differs from the fp32 reduction by bf16 rounding only (rel-L2 <= 2^-8), and is written back into the fp32 buffer
        os.environ["EGV_OVERLAP_ALLREDUCE"] = "1"
        os.environ["EGV_ALLREDUCE_BF16"] = "1"
        try:
            weights.cache().arena = None
            model = build_tiny(c)
            model.load_state_dict(sd, strict=False)
            model.eval()
            step = PretrainStep(model, torch.device("cpu"), lr=0.0, weight_decay=0.0, gather="nccl")
            model.itm_plan = dict(labels=torch.tensor([1., 0.]), swap_video=torch.tensor([False, True]), neg_idx=torch.tensor([1 - rank, 2 * (1 - rank)]))
            step.step(batch)
            g16 = step.opt.arena.grad
            assert g16.dtype == torch.float32 and step.reducer.bf16
            err = ((g16 - grads["0"]).norm() / grads["0"].norm()).item()
            assert 0.0 < err <= 2.0 ** -8, err
        finally:
            os.environ["EGV_ALLREDUCE_BF16"] = "0"
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_two_rank_overlapped_gradient_reduction_equals_single_allreduce(golden_dir):
    """reduce.OverlappedGradReducer (buckets all-reduced from backward hooks of the EgoNCE pass) on two gloo
    ranks through trainer.PretrainStep: identical to the single all-reduce of the flat gradient buffer after the backward,
    i.e. no bucket is sent before its gradients are final (the ITM / MLM passes back-propagate first)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_reducer_worker, args=(r, 2, port, q, golden_dir)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1].startswith("ok") for r in res), res


def test_bench_reference_arm_line():
    """`bench.py --impl reference` (the CPU port of the path on the host cores) prints ONE JSON line with the contract's
    keys: impl, metric / unit of our arm, cpu_baseline{kind, cores, sample, value == the line's}, e2e with zero copy bytes."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample", "1",
                        "--frames", "4"], cwd=root, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "pretrain_clips_per_sec" and d["unit"] == "clips/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
