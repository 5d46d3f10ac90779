"""ORACLE tooling — imports the UNMODIFIED reference model by path (container only).

`/root/reference` does not exist on the GPU box; nothing that runs there may import this.
Used by `oracle/make_golden.py` (fixtures) and `oracle/build_ref.py` (TorchScript baseline).
"""
import contextlib
import io
import os
import sys

REF_DIR = os.environ.get("UAHN_REFERENCE_DIR", "/root/reference/trace_pytorch_model")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "model_to_trace.py"))


def build_reference_model(state_dict, show_error: bool, blocks_to_run: int = 3):
    """Reference factory `HomoNet_ICSTN_Down_stu` (model_to_trace.py:333-350) on a given state_dict."""
    import torch
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import model_to_trace as m  # noqa: the reference's own module
    with contextlib.redirect_stdout(io.StringIO()):
        net = m.HomoNet_ICSTN_Down_stu(224, 320, "cpu", pretrained_stu_model={"state_dict": state_dict},
                                       dropout_rate=0.05, show_photometric_error=show_error)
    net.model_part1.blocks_to_run = blocks_to_run      # the reference's own knob (model_to_trace.py:72)
    net.eval()
    for p in net.parameters():
        p.requires_grad = False                        # trace_model.py:26-27
    return net, m
