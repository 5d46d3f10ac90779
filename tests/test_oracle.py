"""CPU: the oracle restatement must reproduce the vectors recorded from the unmodified reference."""
import numpy as np
import torch

from conftest import unpack_masks
from cuahn_vio_b200 import synthetic as S
from oracle import uahn_oracle as O


def _tol(g):
    # bit-exact in the container that made the fixtures (same torch build, same CPU kernels: /root/reference is only
    # mounted there); on other hosts oneDNN / MKL pick other code paths for the same ops, so: fp32 noise
    import os
    same_host = str(g["torch_version"]) == torch.__version__ and os.path.isdir("/root/reference")
    return 0.0 if same_host else 2e-4


def test_masks_replay_matches_torch_seed(golden_e2e):
    """the masks the reference consumed == F.dropout replay after the same manual_seed."""
    if str(golden_e2e["torch_version"]) != torch.__version__:
        import pytest
        pytest.skip("mask replay is only defined for the recording torch version")
    for i in range(3):
        stored = unpack_masks(golden_e2e, i)
        replay = S.torch_dropout_masks(20240 + 10 ** 6 + i)
        for a, b in zip(stored, replay):
            assert torch.equal(a, b)


def test_dropout_scale_value():
    m = S.torch_dropout_masks(3)[0]
    vals = torch.unique(m)
    assert vals.numel() == 2 and vals[0] == 0 and vals[1] == torch.tensor(1.0 / 0.95, dtype=torch.float32)


def test_e2e_matches_reference(golden_e2e, synth_sd):
    g = golden_e2e
    tol = _tol(g)
    for i in range(3):
        masks = unpack_masks(g, i)
        i1, i2 = O.u8_to_unit(g["prev"][i]), O.u8_to_unit(g["curr"][i])
        for variant in ("prior3", "full"):
            pr = torch.from_numpy(g["prior"][i]).view(1, 1, 4, 2) if variant == "prior3" else None
            t = O.Taps()
            flow, cov, err = O.forward(i1, i2, synth_sd, masks, pr, show_error=True, taps=t)
            k = f"{variant}_{i}"
            assert np.abs(flow.numpy().reshape(8) - g[f"flow_{k}"]).max() <= tol
            assert np.abs(cov.numpy() - g[f"cov_{k}"]).max() <= tol
            assert abs(err.double().sum().item() - float(g[f"errsum_{k}"])) <= tol * 1e5
            for b in (2, 3) + ((1,) if variant == "full" else ()):
                assert np.abs(t.d[b].numpy() - g[f"d{b}_{k}"]).max() <= tol
            if i == 0:
                assert np.abs(err[0, 0].numpy() - g[f"err_{k}"]).max() <= tol * 255
                assert np.abs(t.feat[4].numpy().reshape(256, 4, 5) - g[f"feat4_{k}"]).max() <= tol


def test_blocks_to_run_variants(golden_e2e, synth_sd):
    g = golden_e2e
    masks = unpack_masks(g, 0)
    i1, i2 = O.u8_to_unit(g["prev"][0]), O.u8_to_unit(g["curr"][0])
    pr = torch.from_numpy(g["prior"][0]).view(1, 1, 4, 2)
    for btr in (2, 1):
        flow, cov, _ = O.forward(i1, i2, synth_sd, masks, pr, show_error=True, blocks_to_run=btr)
        assert np.abs(flow.numpy().reshape(8) - g[f"flow_prior{btr}_0"]).max() <= _tol(g)
        assert np.abs(cov.numpy() - g[f"cov_prior{btr}_0"]).max() <= _tol(g)


def test_stage_known_answers(golden_stages):
    g = golden_stages
    tol = _tol(g)
    pts0 = O.origin_4pt().unsqueeze(0)
    offs = torch.from_numpy(g["dlt_offsets"])
    for i in range(8):
        Hm = O.dlt_solve(pts0, pts0 + offs[i:i + 1])
        assert np.abs(Hm.numpy()[0] - g["dlt_H"][i]).max() <= tol
    # analytic KATs: DLT(p,p)=I ; DLT(p,p+t)=translation
    assert np.abs(g["dlt_H"][0] - np.eye(3)).max() < 1e-4
    T = np.eye(3); T[0, 2], T[1, 2] = 3.5, -2.25
    assert np.abs(g["dlt_H"][1] - T).max() < 2e-3
    img = O.u8_to_unit(g["warp_src_u8"])
    for i in range(4):
        Hm = torch.from_numpy(g["warp_H"][i])
        assert np.abs(O.warp_image(img, Hm)[0, 0].numpy() - g["warp_out"][i]).max() <= tol
        ix, iy, _, _ = O.sample_indices(Hm)
        assert np.array_equal(ix.numpy().astype(np.int16), g["warp_ix"][i])
        assert np.array_equal(iy.numpy().astype(np.int16), g["warp_iy"][i])
    p2, cov = O.transfer_mean_var_single(torch.from_numpy(g["tr_var"]), torch.from_numpy(g["tr_H"]),
                                         torch.from_numpy(g["tr_pts"]))
    assert np.abs(p2.numpy() - g["tr_p2"]).max() <= tol and np.abs(cov.numpy() - g["tr_cov"]).max() <= tol


def test_warp_identity_and_constant():
    img = torch.rand(1, 1, 224, 320)
    # the normalise/un-normalise round trip (warp.py:70) leaves ~1e-5 px of coordinate noise
    assert (O.warp_image(img, torch.eye(3)) - img).abs().max() < 1e-4
    Hm = torch.eye(3); Hm[0, 2] = 1000.0
    assert O.warp_image(torch.ones(1, 1, 224, 320), Hm).abs().max() == 0


def test_transfer_identity():
    var = torch.rand(1, 4, 2) + 0.1
    pts = O.origin_4pt().unsqueeze(0)
    p2, cov = O.transfer_mean_var_single(var, torch.eye(3).unsqueeze(0), pts)
    for i in range(4):
        assert torch.allclose(cov[0, i], torch.diag(var[0, i]))


def test_dlt_closed_form_of_the_fixed_rectangle_equals_the_linear_solve(golden_stages):
    """The product path's DLT (head_kernels.cu dlt_rect) is the square-to-quadrilateral mapping composed with the frame's
    scaling, because the four source points are always the frame corners (model_to_trace.py:78-83).  Restated here in numpy
    and held against the oracle's 8x8 solve (model_to_trace.py:42-61) in fp64, and against the reference-made fixture."""
    def dlt_rect(dst):                       # dst [4, 2] in the model's corner order UL, BL, BR, UR
        (x0, y0), (x3, y3), (x2, y2), (x1, y1) = dst
        dx1, dx2, dy1, dy2 = x1 - x2, x3 - x2, y1 - y2, y3 - y2
        sx, sy = (x0 - x1) + (x2 - x3), (y0 - y1) + (y2 - y3)
        inv = 1.0 / (dx1 * dy2 - dy1 * dx2)
        g, k = (sx * dy2 - sy * dx2) * inv, (dx1 * sy - dy1 * sx) * inv
        iw, ih = 1.0 / 319.0, 1.0 / 223.0
        return np.array([[((x1 - x0) + g * x1) * iw, ((x3 - x0) + k * x3) * ih, x0],
                         [((y1 - y0) + g * y1) * iw, ((y3 - y0) + k * y3) * ih, y0],
                         [g * iw, k * ih, 1.0]])

    pts0 = O.origin_4pt().unsqueeze(0).double()
    rng = np.random.default_rng(5)
    for _ in range(200):
        off = (rng.random((1, 4, 2)) * 2 - 1) * 40
        ref = O.dlt_solve(pts0, pts0 + torch.from_numpy(off))[0].numpy()
        got = dlt_rect((pts0 + torch.from_numpy(off))[0].numpy())
        assert np.abs(got - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max())
    assert np.array_equal(dlt_rect(pts0[0].numpy()), np.eye(3))          # DLT(p, p) = I exactly
    g = golden_stages                                                     # the reference's own fp32 torch.inverse results
    for off, href in zip(g["dlt_offsets"].reshape(-1, 4, 2), g["dlt_H"]):
        got = dlt_rect(pts0[0].numpy() + off.astype(np.float64))
        assert np.abs(got - href).max() < 2e-4
