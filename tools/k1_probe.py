#!/usr/bin/env python3
"""Per-kernel CUDA-event times of strictly serial 1024-proof passes (the numbers bench.py reports as kernel_ms), for experiment builds.
  SSYM_NVCC_EXTRA=-DK1_R_ADDMODE=0 python stark-symphony_b200/build.py && python tools/k1_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import stark_symphony_b200 as S

ver = S.Verifier(0)
cfg = S.stwo_config("prod", S.MODE_REF_LITERAL)
one, bad = S.witness.pack_stwo_wits([open(os.path.join(ROOT, "tests", "golden", "stwo_proof_prod.wit")).read()], cfg)
n = 1024
dev = torch.from_numpy(np.tile(one, n).view(np.int32)).cuda()
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ver.set_stream(stream.cuda_stream)
acc = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
st = torch.zeros(n, dtype=torch.int32, device="cuda")
for _ in range(20):
    ver.stwo_verify_batch(dev, cfg, n, accept_out=acc, status_out=st)
torch.cuda.synchronize()
ver.profile_read()
ver.profile_enable(True)
for _ in range(200):
    ver.stwo_verify_batch(dev, cfg, n, accept_out=acc, status_out=st)
torch.cuda.synchronize()
ver.profile_enable(False)
print(os.environ.get("SSYM_NVCC_EXTRA", "release"), {k: round(v[0] / max(v[1], 1) * 1e3, 2) for k, v in ver.profile_read().items() if v[1]}, "us; status[0] = 0x%x" % (int(st[0]) & 0xffffffff))
