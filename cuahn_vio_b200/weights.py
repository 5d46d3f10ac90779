"""One-time weight export: reference checkpoint / state_dict -> flat file read by libuahn.so.

The reference loads `torch.load("UAHN_fcdrop05_16.pth.tar")['state_dict']` (trace_model.py:14,
model_to_trace.py:344).  This writes the same 54 tensors, fp32, little-endian:

    "UAHNWTS1" | u32 count | count x { u32 name_len | name | u32 ndim | u32 dims[ndim] | f32 data }

All kernel-specific re-layout (im2col K order, NHWC permutation of the 5120-wide FC inputs,
bf16 conversion, UMMA swizzle, Toeplitz expansion of the 2-channel first layers) happens inside
the library at load time, so the file stays a faithful copy of the checkpoint.
"""
from __future__ import annotations

import struct

import numpy as np

from .synthetic import state_dict_schema


def export_state_dict(state_dict, path: str) -> None:
    schema = state_dict_schema()
    missing = [k for k in schema if k not in state_dict]
    if missing:
        raise KeyError(f"state_dict lacks {len(missing)} tensors, e.g. {missing[0]}")
    with open(path, "wb") as f:
        f.write(b"UAHNWTS1")
        f.write(struct.pack("<I", len(schema)))
        for key, shape in schema.items():
            t = state_dict[key]
            a = np.ascontiguousarray(t.detach().cpu().numpy() if hasattr(t, "detach") else t, dtype="<f4")
            if tuple(a.shape) != tuple(shape):
                raise ValueError(f"{key}: shape {a.shape} != {shape}")
            name = key.encode()
            f.write(struct.pack("<I", len(name)))
            f.write(name)
            f.write(struct.pack("<I", a.ndim))
            f.write(struct.pack(f"<{a.ndim}I", *a.shape))
            f.write(a.tobytes())


def export_checkpoint(pth_tar: str, path: str) -> None:
    """Convert the reference's `.pth.tar` (dict with key 'state_dict')."""
    import torch
    ck = torch.load(pth_tar, map_location="cpu")
    export_state_dict(ck["state_dict"] if "state_dict" in ck else ck, path)


def synthetic_weights_file(seed: int = 0, directory: str | None = None) -> str:
    """Export the seeded synthetic state_dict (cached by seed) and return the file path."""
    import os
    import tempfile
    from .synthetic import synthetic_state_dict
    directory = directory or os.path.join(tempfile.gettempdir(), "uahn_weights")
    os.makedirs(directory, exist_ok=True)
    path = os.path.join(directory, f"uahn_synth_seed{seed}.bin")
    if not os.path.exists(path):
        tmp = path + f".{os.getpid()}.tmp"
        export_state_dict(synthetic_state_dict(seed), tmp)
        os.replace(tmp, path)
    return path
