#!/usr/bin/env python
"""MEASUREMENT INFRASTRUCTURE ONLY -- stage the UNMODIFIED reference model code for the GPU-eager baseline.

    python tools/install_ref.py

Copies the reference's own files for the pre-training path (EgoVLPv2/{model,base,utils}, parse_config.py and the yaml the
model modules open at import time) from /root/reference into baseline/_ref/EgoVLPv2/.  `baseline/_ref/` is git-ignored
(reference sources never enter this repository's history) but NOT gpurun-ignored, so it travels to the GPU box, where
/root/reference does not exist.  Only `tools/bench_ref_gpu.py` (the ">= 6x" denominator of BASELINE.md B1/B2) reads it;
nothing in egovlpv2_b200/, tests/ or bench.py's own arm does."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/EgoVLPv2"
DST = os.path.join(ROOT, "baseline", "_ref", "EgoVLPv2")
ITEMS = ["model", "base", "utils", "parse_config.py", "EgoNCE_MLM_ITM_Config.yml"]


def main():
    if not os.path.isdir(SRC):
        print("reference tree %s not present: nothing installed" % SRC)
        return 1
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    for it in ITEMS:
        s, d = os.path.join(SRC, it), os.path.join(DST, it)
        if os.path.isdir(s):
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        else:
            shutil.copy2(s, d)
    print("installed the unmodified reference model code into", DST)
    return 0


if __name__ == "__main__":
    sys.exit(main())
