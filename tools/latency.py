#!/usr/bin/env python
"""p50 / p90 batch-1 latency of uahn_infer (CUDA-graph replay) for one variant and precision — a quick A/B tool.

    python tools/latency.py [--variant prior3|full] [--precision bf16|fp32] [--calls 2000] [--show-error]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from cuahn_vio_b200 import api, build, synthetic as S, weights  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--variant", default="prior3")
ap.add_argument("--precision", default="bf16")
ap.add_argument("--calls", type=int, default=2000)
ap.add_argument("--show-error", action="store_true")
ap.add_argument("--tag", default="")
a = ap.parse_args()
if not os.environ.get("UAHN_LIB_PATH"):
    build.build()
frames, _, prior = S.synthetic_sequence(3, seed=1)
with api.Uahn(weights.synthetic_weights_file(0), a.variant, show_error=a.show_error, precision=a.precision, max_batch=1) as net:
    net.load_image(frames[0], 0.0)
    net.load_image(frames[1], 1.0)
    pr = prior[0].reshape(8).astype(np.float64) if a.variant != "full" else None
    for i in range(100):
        net.infer(pr, seed=1, pair_index=i, want_error=a.show_error)
    ts = []
    for i in range(a.calls):
        t0 = time.perf_counter()
        net.infer(pr, seed=1, pair_index=i, want_error=a.show_error)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    l0 = net.launch_count
    net.infer(pr, seed=1, pair_index=0, want_error=a.show_error)
    print(f"{a.tag} {a.variant} {a.precision}: p50 {1e3 * ts[len(ts) // 2]:.4f} ms  p90 {1e3 * ts[int(0.9 * len(ts))]:.4f} ms  "
          f"p99 {1e3 * ts[int(0.99 * len(ts))]:.4f} ms  kernels/call {net.launch_count - l0}")
