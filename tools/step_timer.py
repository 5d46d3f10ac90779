#!/usr/bin/env python
"""Steady-state step time of the resident hot path (A/B runs on one box): python tools/step_timer.py [--batch 1024]."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cuahn_vio_b200 import api, build, synthetic as S, weights  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--rounds", type=int, default=5)
ap.add_argument("--tag", default="")
a = ap.parse_args()
if not os.environ.get("UAHN_LIB_PATH"):
    build.build()
dev = torch.device("cuda", 0)
n = a.batch
hp, hc, _, hpr = S.tiled_batch(n, unique=16)
p, c, pr = (torch.from_numpy(x).to(dev) for x in (hp, hc, hpr.reshape(n, 8)))
mean, cov = torch.empty(n, 8, device=dev), torch.empty(n, 64, device=dev)
net = api.Uahn(weights.synthetic_weights_file(0), "prior3", precision="bf16", max_batch=n)


def step():
    net.infer_batch_ptrs(n, p.data_ptr(), c.data_ptr(), pr.data_ptr(), mean.data_ptr(), cov.data_ptr(), None, seed=1)


for _ in range(60):       # long enough for the power cap to settle
    step()
net.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for r in range(a.rounds):
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        step()
    net.synchronize()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 20)
# stage split of the same steady state (per-stage CUDA events: warp / conv / MC GEMM / small kernels)
net.profile_enable(True)
for _ in range(20):
    step()
net.synchronize()
ms, cnt = net.profile_read()
net.profile_enable(False)
print(a.tag, "ms/step:", " ".join("%.4f" % t for t in ts), "| stages ms/step (warp, conv, mc gemm, small):",
      " ".join("%.4f" % (m / 20) for m in ms))
