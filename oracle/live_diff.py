"""TEST INFRASTRUCTURE ONLY -- live differential check, in the build container, of the drop-in module API against the
UNMODIFIED reference (/root/reference through oracle/ref_shim.py) on FRESH seeded inputs, beyond the committed goldens:
odd / single-row batches, ragged caption lengths, other frame counts and image sizes.

    python -m oracle.live_diff            # exit 0 = every case within tolerance; prints one line per case

The drop-in runs its real host logic (egovlpv2_b200.model / autograd / functional) over the torch restatement of the
kernel interface in exact fp32 mode (tests/fake_kernels.py), so any difference beyond op-order noise is a logic error.
Run by tests/test_oracle_golden.py::test_live_differential_vs_reference in a subprocess (the shim patches process-wide
state); skipped where /root/reference does not exist (the GPU box)."""
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import egovlp_oracle as O  # noqa: E402
from oracle import make_golden as MG  # noqa: E402
from oracle import ref_shim  # noqa: E402

CASES = [  # (B, T, img, S, data seed, weight seed)
    dict(B=1, T=2, img=64, S=8, seed=11, wseed=1),
    dict(B=5, T=2, img=32, S=6, seed=12, wseed=2),
    dict(B=3, T=4, img=48, S=12, seed=13, wseed=3),
]
CHECK_GRADS = ["video_model.blocks.0.timeattn.qkv.weight", "video_model.blocks.7.attn.alpha_i2t",
               "video_model.blocks.6.attn.qkv_i2t.weight", "video_model.patch_embed.proj.weight", "video_model.temporal_embed",
               "cls_token", "text_model.encoder.layer.7.alpha_t2i", "text_model.encoder.layer.6.crossattention_t2i.self.value.weight",
               "text_model.embeddings.word_embeddings.weight", "txt_proj.2.weight", "vid_proj.0.weight", "mlm_score.bias",
               "itm_score.fc.weight", "cross_modal_text_transform.weight"]


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


def main():
    ref = ref_shim.load(use_checkpoint=False)
    from egovlpv2_b200 import functional as Fn
    from egovlpv2_b200 import lib as L
    from egovlpv2_b200.model.loss import EgoNCE
    from tests.fake_kernels import FakeKernels
    from tests.test_model_cpu import build_tiny
    L.set_kernels(FakeKernels())
    Fn.BF16 = torch.float32
    worst = 0.0
    for case in CASES:
        c = dict(MG.TINY, B=case["B"], T=case["T"], img=case["img"], S=case["S"])
        model_ref, shapes, sd = MG.build_tiny_reference(ref, c, seed=case["wseed"])
        data = O.synthetic_batch(c["B"], c["T"], c["img"], c["S"], seed=case["seed"])
        plan = O.synthetic_itm_plan(c["B"], seed=case["seed"] + 100)
        model_ref.zero_grad()
        loss_r, ld_r, ret_r = MG.reference_step(ref, model_ref, data, plan, grad=True)
        loss_r.backward()
        g_r = dict((n, p.grad) for n, p in model_ref.named_parameters())

        mine = build_tiny(c)
        mine.load_state_dict(sd, strict=False)
        mine.eval()
        mine.itm_plan = plan
        batch = {"video": data["video"], "text": {"input_ids": data["input_ids"], "attention_mask": data["attention_mask"]},
                 "text_mlm_ids": data["text_mlm_ids"], "text_mlm_labels": data["text_mlm_labels"]}
        args = types.SimpleNamespace(world_size=1, rank=0)
        loss_m, ld_m, ret_m = mine(batch, data["noun_vec"], data["verb_vec"], lambda t, n, a: t, 1, args,
                                   {"loss": {"type": "EgoNCE"}}, EgoNCE(), 0, task_names="EgoNCE_MLM_ITM")
        loss_m.backward()
        g_m = dict((n, p.grad) for n, p in mine.named_parameters())
        errs = {k: abs(float(ld_m[k]) - float(ld_r[k])) / max(1.0, abs(float(ld_r[k]))) for k in ld_r}
        errs["sim_v2t"] = (ret_m["sim_v2t"] - ret_r["sim_v2t"]).abs().max().item()
        errs["itm_logits"] = rel(ret_m["cross_attn_itm_logits"], ret_r["cross_attn_itm_logits"])
        errs["mlm_logits"] = rel(ret_m["cross_attn_mlm_logits"], ret_r["cross_attn_mlm_logits"])
        # B = 1: EgoNCE has no negatives (its loss and gradients vanish identically) -> gradients compared absolutely
        gmax = max(g_r[k].norm().item() for k in CHECK_GRADS)
        for k in CHECK_GRADS:
            errs["d " + k[-40:]] = (g_m[k] - g_r[k]).norm().item() / max(g_r[k].norm().item(), 1e-3 * gmax)
        w = max(errs.values())
        worst = max(worst, w)
        print("case %s: worst error %.2e (%s)" % (case, w, max(errs, key=errs.get)))
        assert w <= 2e-3, errs
    print("live differential vs reference ok: worst %.2e" % worst)


if __name__ == "__main__":
    main()
