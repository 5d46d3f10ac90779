"""Seeded synthetic inputs for the UAHN hot path: weights and textured image pairs.

The reference checkpoint (`trace_pytorch_model/UAHN_fcdrop05_16.pth.tar`) is not
shipped (reference `.MISSING_LARGE_BLOBS:2`), so every parity / bench run uses a
seeded synthetic ``state_dict`` with the exact schema the reference factory loads
(`model_to_trace.py:333-350`, 54 tensors, 6 541 312 parameters).  The generators
are deterministic functions of their seeds on CPU, so the GPU box regenerates the
same tensors the golden fixtures were produced with.

Nothing here is on the product compute path; it only manufactures inputs.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch

IMG_H = 224
IMG_W = 320
# corner order UL, BL, BR, UR as (u, v) — model_to_trace.py:78-83
ORIGIN_4PT = np.array([[0.0, 0.0], [0.0, IMG_H - 1.0], [IMG_W - 1.0, IMG_H - 1.0], [IMG_W - 1.0, 0.0]],
                      dtype=np.float32)

P1 = "model_part1."
P4 = "model_last_block_list.0."

# (key prefix, Cout, Cin, k, stride) for every conv, in forward order  (model_to_trace.py:93-113, 210-216)
CONV_LAYERS = OrderedDict([
    ("block_1_1", (P1, 128, 2, 7, 2)), ("block_1_2", (P1, 128, 128, 5, 2)), ("block_1_3", (P1, 256, 128, 3, 2)),
    ("block_2_1", (P1, 64, 2, 7, 2)), ("block_2_2", (P1, 128, 64, 5, 2)), ("block_2_3", (P1, 256, 128, 3, 2)),
    ("block_2_4", (P1, 256, 256, 3, 2)),
    ("block_3_0", (P1, 16, 2, 7, 1)), ("block_3_1", (P1, 32, 16, 5, 2)), ("block_3_2", (P1, 64, 32, 3, 2)),
    ("block_3_3", (P1, 128, 64, 3, 2)), ("block_3_4", (P1, 256, 128, 3, 2)), ("block_3_5", (P1, 256, 256, 3, 2)),
    ("block_4_0", (P4, 8, 2, 7, 1)), ("block_4_1", (P4, 16, 8, 5, 2)), ("block_4_2", (P4, 32, 16, 3, 2)),
    ("block_4_3", (P4, 64, 32, 3, 2)), ("block_4_4", (P4, 128, 64, 3, 2)), ("block_4_5", (P4, 256, 128, 3, 2)),
    ("block_4_6", (P4, 256, 256, 3, 2)),
])
# (key, out, in)  (model_to_trace.py:97,105,115,222-235)
FC_LAYERS = OrderedDict([
    ("fc_block_1", (P1 + "fc_block_1", 8, 5120)),
    ("fc_block_2", (P1 + "fc_block_2", 8, 5120)),
    ("fc_block_3", (P1 + "fc_block_3", 8, 5120)),
    ("fc4_mean_1", (P4 + "fc_block_4_mean.1", 256, 5120)),
    ("fc4_mean_4", (P4 + "fc_block_4_mean.4", 8, 256)),
    ("fc4_unc_1", (P4 + "fc_block_4_uncertainty.1", 256, 5120)),
    ("fc4_unc_4", (P4 + "fc_block_4_uncertainty.4", 8, 256)),
])


def state_dict_schema():
    """Ordered {key: shape} exactly as `combined_stu_model.state_dict()` yields it."""
    out = OrderedDict()

    def conv(name):
        pre, co, ci, k, _ = CONV_LAYERS[name]
        out[f"{pre}{name}.0.weight"] = (co, ci, k, k)
        out[f"{pre}{name}.0.bias"] = (co,)

    def fc(name):
        key, o, i = FC_LAYERS[name]
        out[key + ".weight"] = (o, i)
        out[key + ".bias"] = (o,)

    for n in ("block_1_1", "block_1_2", "block_1_3"):
        conv(n)
    fc("fc_block_1")
    for n in ("block_2_1", "block_2_2", "block_2_3", "block_2_4"):
        conv(n)
    fc("fc_block_2")
    for n in ("block_3_0", "block_3_1", "block_3_2", "block_3_3", "block_3_4", "block_3_5"):
        conv(n)
    fc("fc_block_3")
    for n in ("block_4_0", "block_4_1", "block_4_2", "block_4_3", "block_4_4", "block_4_5", "block_4_6"):
        conv(n)
    for n in ("fc4_mean_1", "fc4_mean_4", "fc4_unc_1", "fc4_unc_4"):
        fc(n)
    return out


def synthetic_state_dict(seed: int = 0, fc_gain: float = 1.0) -> "OrderedDict[str, torch.Tensor]":
    """Seeded synthetic weights with the reference schema.

    Conv weights are He-uniform for LeakyReLU(0.1) so activations keep O(1) magnitude through
    the 3–7 layer stacks (the default nn init would shrink them and make outputs prior-dominated,
    hiding conv errors from the parity tests).  The 5120→8 regression heads are scaled so each
    cascade block moves the corners by a few pixels; the log-variance head gives σ² of O(1) px².
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(1000003 * seed + 17)
    sd = OrderedDict()
    for key, shape in state_dict_schema().items():
        if key.endswith(".bias"):
            fan_in = None
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
        else:
            fan_in = int(np.prod(shape[1:]))
            bound = math.sqrt(6.0 / (1.01 * fan_in))
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
            if len(shape) == 2 and shape[0] == 8:
                # final regression layers: a few px per block
                t = t * (2.0 * fc_gain)
        sd[key] = t.float().contiguous()
    return sd


def _texture_canvas(gen: torch.Generator, ch: int = 288, cw: int = 384) -> torch.Tensor:
    """Band-limited noise: 3 octaves of uniform noise, bicubic-upsampled."""
    acc = torch.zeros(1, 1, ch, cw)
    for octave, amp in ((8, 1.0), (16, 0.6), (32, 0.35)):
        hk, wk = ch // octave + 2, cw // octave + 2
        n = torch.rand(1, 1, hk, wk, generator=gen)
        acc = acc + amp * torch.nn.functional.interpolate(n, size=(ch, cw), mode="bicubic", align_corners=True)
    acc = acc - acc.amin()
    acc = acc / acc.amax().clamp_min(1e-6)
    return acc


def dlt_numpy(src: np.ndarray, dst: np.ndarray) -> np.ndarray:
    """4-point homography in float64 (input manufacture only, not the kernel under test)."""
    A = np.zeros((8, 8))
    b = np.zeros(8)
    for i in range(4):
        x, y = src[i]
        u, v = dst[i]
        A[2 * i] = [x, y, 1, 0, 0, 0, -u * x, -u * y]
        A[2 * i + 1] = [0, 0, 0, x, y, 1, -v * x, -v * y]
        b[2 * i], b[2 * i + 1] = u, v
    h = np.linalg.solve(A, b)
    return np.append(h, 1.0).reshape(3, 3)


def synthetic_pair(index: int, base_seed: int = 20240, max_disp: float = 16.0, prior_sigma: float = 2.0,
                   disp: np.ndarray | None = None):
    """One synthetic textured pair.

    Returns (img_prev u8[224,320], img_curr u8[224,320], gt_offset f32[4,2], prior f32[4,2]).
    img_curr is the canvas seen through a ground-truth homography given by corner displacements
    U(-max_disp, max_disp); the prior is GT + N(0, prior_sigma²).
    """
    gen = torch.Generator(device="cpu")
    gen.manual_seed(base_seed + index)
    ch, cw = 288, 384
    canvas = _texture_canvas(gen, ch, cw)
    if disp is None:
        disp = ((torch.rand(4, 2, generator=gen) * 2 - 1) * max_disp).numpy().astype(np.float64)
    noise = (torch.randn(4, 2, generator=gen) * prior_sigma).numpy()
    oy, ox = (ch - IMG_H) // 2, (cw - IMG_W) // 2
    src = ORIGIN_4PT.astype(np.float64)
    Hm = dlt_numpy(src, src + disp)  # img1 pixel -> img2 pixel
    # img2(p2) = canvas(crop + H^-1 p2): sample canvas at inverse-mapped positions
    Hinv = np.linalg.inv(Hm)
    v, u = np.meshgrid(np.arange(IMG_H, dtype=np.float64), np.arange(IMG_W, dtype=np.float64), indexing="ij")
    p = Hinv @ np.stack([u.ravel(), v.ravel(), np.ones(u.size)])
    x = p[0] / p[2] + ox
    y = p[1] / p[2] + oy
    grid = torch.from_numpy(np.stack([x / (cw - 1) * 2 - 1, y / (ch - 1) * 2 - 1], -1).reshape(1, IMG_H, IMG_W, 2)).float()
    img2 = torch.nn.functional.grid_sample(canvas, grid, mode="bilinear", padding_mode="border", align_corners=True)
    img1 = canvas[:, :, oy:oy + IMG_H, ox:ox + IMG_W]
    to_u8 = lambda t: (t[0, 0] * 255.0).round().clamp(0, 255).to(torch.uint8).numpy()
    prior = (disp + noise).astype(np.float32)
    return to_u8(img1), to_u8(img2), disp.astype(np.float32), prior


def synthetic_batch(n: int, start: int = 0, base_seed: int = 20240, **kw):
    """Stack `n` pairs: prev u8[n,224,320], curr u8[n,224,320], gt f32[n,4,2], prior f32[n,4,2]."""
    prev = np.empty((n, IMG_H, IMG_W), np.uint8)
    curr = np.empty((n, IMG_H, IMG_W), np.uint8)
    gt = np.empty((n, 4, 2), np.float32)
    prior = np.empty((n, 4, 2), np.float32)
    for i in range(n):
        prev[i], curr[i], gt[i], prior[i] = synthetic_pair(start + i, base_seed, **kw)
    return prev, curr, gt, prior


def tiled_batch(n: int, unique: int = 64, base_seed: int = 20240):
    """`n` pairs built by cycling `unique` generated pairs (bench input manufacture; cheap on CPU)."""
    prev, curr, gt, prior = synthetic_batch(min(unique, n), 0, base_seed)
    reps = (n + prev.shape[0] - 1) // prev.shape[0]
    tile = lambda a: np.concatenate([a] * reps, 0)[:n].copy()
    return tile(prev), tile(curr), tile(gt), tile(prior)


def synthetic_sequence(n_frames: int, seed: int = 20240, sigma: float = 4.0, rho: float = 0.95,
                       prior_sigma: float = 2.0):
    """One synthetic sequence: a textured canvas seen through a homography whose four
    corner displacements follow a smooth AR(1) random walk (rho, stationary sigma px per coordinate).

    Returns (frames u8[n_frames,224,320], gt f32[n_frames-1,4,2], prior f32[n_frames-1,4,2]); pair i is
    (frames[i], frames[i+1]), gt[i] the displacement of frame i's corners into frame i+1, prior = gt + N(0, prior_sigma²).
    """
    gen = torch.Generator(device="cpu")
    gen.manual_seed(seed)
    ch, cw = 288, 384
    canvas = _texture_canvas(gen, ch, cw)
    oy, ox = (ch - IMG_H) // 2, (cw - IMG_W) // 2
    src = ORIGIN_4PT.astype(np.float64)
    v, u = np.meshgrid(np.arange(IMG_H, dtype=np.float64), np.arange(IMG_W, dtype=np.float64), indexing="ij")
    pix = np.stack([u.ravel(), v.ravel(), np.ones(u.size)])
    frames = np.empty((n_frames, IMG_H, IMG_W), np.uint8)
    Hs = []
    d = (torch.randn(4, 2, generator=gen) * sigma).numpy().astype(np.float64)
    for t in range(n_frames):
        Hm = dlt_numpy(src, src + d)              # canvas-crop pixel -> frame-t pixel
        Hs.append(Hm)
        p = np.linalg.inv(Hm) @ pix
        x, y = p[0] / p[2] + ox, p[1] / p[2] + oy
        grid = torch.from_numpy(np.stack([x / (cw - 1) * 2 - 1, y / (ch - 1) * 2 - 1], -1).reshape(1, IMG_H, IMG_W, 2)).float()
        img = torch.nn.functional.grid_sample(canvas, grid, mode="bilinear", padding_mode="border", align_corners=True)
        frames[t] = (img[0, 0] * 255.0).round().clamp(0, 255).to(torch.uint8).numpy()
        d = rho * d + np.sqrt(1.0 - rho * rho) * (torch.randn(4, 2, generator=gen) * sigma).numpy()
    gt = np.empty((n_frames - 1, 4, 2), np.float32)
    c = np.concatenate([src.T, np.ones((1, 4))])
    for t in range(n_frames - 1):
        q = Hs[t + 1] @ np.linalg.inv(Hs[t]) @ c   # frame-t pixel -> frame-(t+1) pixel
        gt[t] = ((q[:2] / q[2]).T - src).astype(np.float32)
    prior = gt + (torch.randn(n_frames - 1, 4, 2, generator=gen) * prior_sigma).numpy().astype(np.float32)
    return frames, gt, prior.astype(np.float32)


def synthetic_raw_frame(seed: int, rows: int = 480, cols: int = 640) -> np.ndarray:
    """Seeded raw camera frame (u8 [rows, cols]) with structure at several scales; numpy only."""
    rng = np.random.default_rng(seed)
    v, u = np.meshgrid(np.arange(rows, dtype=np.float64), np.arange(cols, dtype=np.float64), indexing="ij")
    acc = np.zeros((rows, cols))
    for cell, amp in ((8, 1.0), (32, 0.7), (96, 0.5)):
        n = rng.random((rows // cell + 2, cols // cell + 2))
        # bilinear upsampling of the coarse noise grid
        gy, gx = v / cell, u / cell
        y0, x0 = gy.astype(np.int64), gx.astype(np.int64)
        fy, fx = gy - y0, gx - x0
        acc += amp * ((1 - fy) * ((1 - fx) * n[y0, x0] + fx * n[y0, x0 + 1]) + fy * ((1 - fx) * n[y0 + 1, x0] + fx * n[y0 + 1, x0 + 1]))
    acc += 0.3 * np.sin(u * 0.11) * np.cos(v * 0.07)
    acc = (acc - acc.min()) / (acc.max() - acc.min())
    return np.round(acc * 255).astype(np.uint8)


def torch_dropout_masks(seed: int, mc: int = 16):
    """The four MC-dropout masks the reference consumes after ``torch.manual_seed(seed)``.

    Order: mean head input [mc,5120], mean head hidden [mc,256], then the same two for the
    uncertainty head.  Values ∈ {0, 1/0.95}.  Indexing of the 5120 axis is the reference's NCHW
    flatten order c*20 + h*5 + w (model_to_trace.py:253).
    """
    torch.manual_seed(seed)
    F = torch.nn.functional
    m1 = F.dropout(torch.ones(mc, 5120), 0.05, True)
    m2 = F.dropout(torch.ones(mc, 256), 0.05, True)
    u1 = F.dropout(torch.ones(mc, 5120), 0.05, True)
    u2 = F.dropout(torch.ones(mc, 256), 0.05, True)
    return m1, m2, u1, u2
