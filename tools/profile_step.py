#!/usr/bin/env python
"""Run a few hot-path steps with nothing else around them — the command ncu wraps (B200_PROFILING.md).

    ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 10 -c 2 -o gpurun_out/prof \
        python tools/profile_step.py --batch 256 --steps 1
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from cuahn_vio_b200 import api, build, synthetic as S, weights  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--variant", default="prior3")
ap.add_argument("--show-error", action="store_true")
a = ap.parse_args()
build.build()
dev = torch.device("cuda", 0)
hp, hc, _, hprior = S.tiled_batch(a.batch, unique=16)
p, c, pr = (torch.from_numpy(x).to(dev) for x in (hp, hc, hprior.reshape(a.batch, 8)))
mean, cov = torch.empty(a.batch, 8, device=dev), torch.empty(a.batch, 64, device=dev)
err = torch.empty(a.batch, 224 * 320, device=dev) if a.show_error else None
net = api.Uahn(weights.synthetic_weights_file(0), a.variant, show_error=a.show_error, precision=a.precision,
               max_batch=a.batch)
for i in range(a.steps):
    net.infer_batch_ptrs(a.batch, p.data_ptr(), c.data_ptr(), pr.data_ptr() if a.variant != "full" else None,
                         mean.data_ptr(), cov.data_ptr(), err.data_ptr() if err is not None else None, seed=1)
net.synchronize()
print("ok", float(mean.abs().mean()))
