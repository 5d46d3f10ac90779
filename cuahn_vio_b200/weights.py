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


VARIANT_IDS = {"full": 0, "prior3": 1, "prior2": 2, "prior1": 3}


def export_state_dict(state_dict, path: str, meta: dict | None = None) -> None:
    """Write the 54 tensors (+ optional ``__meta__.*`` one-float records, e.g. which traced graph the file came from)."""
    schema = state_dict_schema()
    missing = [k for k in schema if k not in state_dict]
    if missing:
        raise KeyError(f"state_dict lacks {len(missing)} tensors, e.g. {missing[0]}")
    meta = meta or {}
    with open(path, "wb") as f:
        f.write(b"UAHNWTS1")
        f.write(struct.pack("<I", len(schema) + len(meta)))
        for key, value in meta.items():
            name = ("__meta__." + key).encode()
            f.write(struct.pack("<I", len(name)))
            f.write(name)
            f.write(struct.pack("<II", 1, 1))
            f.write(struct.pack("<f", float(value)))
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


def inspect_torchscript(module) -> dict:
    """Which reference graph a traced module holds, read off its inlined graph.

    The reference bakes the Python control flow (`blocks_to_run`, `prior is None`, `show_photometric_error`,
    model_to_trace.py:72,129-132,319) into the trace (trace_model.py:36-46), so the op counts identify it:
    convolutions 20 / 17 / 13 / 7 = full / 3-block prior / 2-block prior / 1-block prior (Appendix A of SURVEY.md);
    one `grid_sampler` per executed warp, plus one for the photometric-error map.
    """
    import collections
    cnt = collections.Counter(n.kind() for n in module.inlined_graph.nodes())
    convs = cnt.get("aten::_convolution", 0) + cnt.get("aten::conv2d", 0)
    variant = {20: "full", 17: "prior3", 13: "prior2", 7: "prior1"}.get(convs)
    if variant is None:
        raise ValueError(f"not a UAHN trace: {convs} convolutions in the graph")
    warps = {"full": 3, "prior3": 3, "prior2": 2, "prior1": 1}[variant]
    extra = cnt.get("aten::grid_sampler", 0) - warps
    if extra not in (0, 1):
        raise ValueError(f"not a UAHN trace: {cnt.get('aten::grid_sampler', 0)} grid_sampler nodes for variant {variant}")
    n_inputs = len(list(module.graph.inputs())) - 1     # minus `self`
    if n_inputs != (2 if variant == "full" else 3):
        raise ValueError(f"variant {variant} with {n_inputs} inputs")
    return {"variant": variant, "show_error": bool(extra)}


def export_torchscript(pt_path: str, path: str) -> dict:
    """Convert a TorchScript `.pt` as produced by the reference's `trace_model.py:41-46` — what
    `HomographyNet.cpp:81-124` loads and `uzhfpv.launch:58` points at — into the flat file.

    The traced module keeps the 54 parameters under the checkpoint's own names, so the export is byte-identical to the
    export of the state_dict it was traced from, plus two `__meta__` records (variant, show_error) that let
    `UAHN_VARIANT_AUTO` / `UAHN_SHOW_ERROR_AUTO` handles reproduce "whatever graph the file holds".
    Returns {"variant": ..., "show_error": ...}.
    """
    import torch
    module = torch.jit.load(pt_path, map_location="cpu")
    info = inspect_torchscript(module)
    export_state_dict(module.state_dict(), path, meta={"variant": VARIANT_IDS[info["variant"]],
                                                       "show_error": int(info["show_error"])})
    return info


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
