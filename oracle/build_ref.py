"""ORACLE tooling — build the CPU baseline artefact from the UNMODIFIED reference (container only).

Traces the reference model exactly as `trace_model.py:36-46` does (torch.jit.trace, check_trace=False) on the
seeded synthetic state_dict and saves the TorchScript files under `oracle/_ref/` (git-ignored; travels to the
GPU box with the snapshot).  `bench.py --impl reference` / the cpu_baseline leg load these with torch.jit.load,
i.e. they time the reference's own graph on libtorch CPU.  No reference source is copied into the repo.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "oracle", "_ref")


def main(quiet: bool = False, force: bool = False, all_variants: bool = False) -> None:
    import torch
    from cuahn_vio_b200 import synthetic as S
    from oracle import ref_import as R
    if not R.available():
        raise FileNotFoundError("/root/reference not present (GPU box): using prebuilt oracle/_ref")
    os.makedirs(OUT, exist_ok=True)
    names = {True: ("traced_full_model_showError.pt", "traced_model_3_blocks_using_prior_showError.pt"),
             False: ("traced_full_model.pt", "traced_model_3_blocks_using_prior.pt")}
    shows = (True, False) if all_variants else (False,)   # the _showError twins only on request (26 MB each)
    if not force and all(os.path.exists(os.path.join(OUT, n)) for s_ in shows for n in (names[s_] if all_variants else names[s_][1:])):
        return
    sd = S.synthetic_state_dict(0)
    img1 = torch.ones(1, 1, 224, 320) * 0.2        # trace_model.py:20-22
    img2 = torch.ones(1, 1, 224, 320) * 0.5
    homo8 = torch.ones(1, 1, 4, 2) * 0.9
    for show in shows:
        net, _ = R.build_reference_model(sd, show_error=show)
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            full = torch.jit.trace(net, (img1, img2), check_trace=False)
            prior = torch.jit.trace(net, (img1, img2, homo8), check_trace=False)
        if all_variants:
            full.save(os.path.join(OUT, names[show][0]))
        prior.save(os.path.join(OUT, names[show][1]))
    with open(os.path.join(OUT, "README.txt"), "w") as f:
        f.write(f"TorchScript traces of the unmodified reference model, synthetic weights seed 0, torch {torch.__version__}\n")
    if not quiet:
        for n in sorted(os.listdir(OUT)):
            print(n, os.path.getsize(os.path.join(OUT, n)) // 1024, "KiB")


if __name__ == "__main__":
    main(force="--force" in sys.argv, all_variants="--all" in sys.argv)
