#!/usr/bin/env python
"""Experiment: two handles on two streams per GPU (sub-batches overlap: one's HBM-bound convs with the other's ALU-bound warps)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cuahn_vio_b200 import api, build, synthetic as S, weights
build.build()
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
NH = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sub = B // NH
hp, hc, _, hprior = S.tiled_batch(B, unique=32)
p, c, pr = (torch.from_numpy(x).to(dev) for x in (hp, hc, hprior.reshape(B, 8)))
mean, cov = torch.empty(B, 8, device=dev), torch.empty(B, 64, device=dev)
w = weights.synthetic_weights_file(0)
streams = [torch.cuda.Stream(dev) for _ in range(NH)]
nets = [api.Uahn(w, "prior3", precision="bf16", max_batch=sub, stream=s.cuda_stream) for s in streams]
def step():
    for i, net in enumerate(nets):
        o = i * sub
        net.infer_batch_ptrs(sub, p[o:].data_ptr(), c[o:].data_ptr(), pr[o:].data_ptr(), mean[o:].data_ptr(), cov[o:].data_ptr(), seed=1, first_pair=o)
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
K = 30
for _ in range(K): step()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"handles={NH} batch={B} reserve={os.environ.get('UAHN_TMA_SMEM_RESERVE','0')}: {B*K/dt:.0f} pairs/s  ({dt/K*1e3:.3f} ms/step)")
