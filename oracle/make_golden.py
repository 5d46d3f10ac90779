"""ORACLE tooling — generate tests/golden/*.npz from the UNMODIFIED reference (run in the container).

    python -m oracle.make_golden

Imports `/root/reference/trace_pytorch_model/{model_to_trace,warp}.py` as they are, loads the seeded
synthetic state_dict through the reference's own factory (`model_to_trace.py:333-350`), and records
inputs, MC-dropout masks and outputs of `combined_stu_model.forward` plus stage values captured with
forward hooks (no reference code is changed or copied).  The reference holds no golden vectors of its
own for this path, so these files are what pins the oracle restatement.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cuahn_vio_b200 import synthetic as S  # noqa: E402
from oracle import ref_import as R  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
WEIGHT_SEED = 0
N_PAIRS = 3
MASK_SEED0 = 20240 + 10 ** 6   # dropout seed = base + 1e6 + i


def pack_masks(masks):
    return [np.packbits((m.numpy() != 0).astype(np.uint8)) for m in masks]


def run_case(net, img1, img2, prior, seed, hooks_out):
    hooks_out.clear()
    torch.manual_seed(seed)
    with torch.no_grad():
        out = net(img1, img2, prior) if prior is not None else net(img1, img2)
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    sd = S.synthetic_state_dict(WEIGHT_SEED)
    prev, curr, gt, prior = S.synthetic_batch(N_PAIRS)
    net, m = R.build_reference_model(sd, show_error=True)
    import warp as refwarp  # the reference's module (on sys.path after build_reference_model)

    cap = {}
    p1, lb = net.model_part1, net.model_last_block_list[0]
    for name in ("fc_block_1", "fc_block_2", "fc_block_3"):
        getattr(p1, name).register_forward_hook(
            lambda mod, inp, out, name=name: cap.__setitem__(name, (inp[0].detach().clone(), out.detach().clone())))
    lb.block_4_6.register_forward_hook(lambda mod, inp, out: cap.__setitem__("feat4", out.detach().clone()))

    rec = {"torch_version": np.array(torch.__version__), "weight_seed": np.array(WEIGHT_SEED),
           "prev": prev, "curr": curr, "prior": prior, "gt": gt}
    for i in range(N_PAIRS):
        i1 = torch.from_numpy(prev[i]).float().div(255.0).view(1, 1, 224, 320)   # HomographyNet.cpp:146
        i2 = torch.from_numpy(curr[i]).float().div(255.0).view(1, 1, 224, 320)
        pr = torch.from_numpy(prior[i]).view(1, 1, 4, 2)
        seed = MASK_SEED0 + i
        masks = S.torch_dropout_masks(seed)
        for j, pm in enumerate(pack_masks(masks)):
            rec[f"mask{j}_{i}"] = pm
        for variant, p in (("prior3", pr), ("full", None)):
            flow, cov, err = run_case(net, i1, i2, p, seed, cap)
            k = f"{variant}_{i}"
            rec[f"flow_{k}"] = flow.numpy().reshape(8)
            rec[f"cov_{k}"] = cov.numpy()
            if i == 0:
                rec[f"err_{k}"] = err[0, 0].numpy().astype(np.float32)
            rec[f"errsum_{k}"] = np.array(err.double().sum().item())
            for b in (1, 2, 3):
                if f"fc_block_{b}" in cap:
                    fin, fout = cap[f"fc_block_{b}"]
                    rec[f"d{b}_{k}"] = fout.numpy().reshape(8)
                    if i == 0:
                        rec[f"feat{b}_{k}"] = fin.numpy().reshape(256, 4, 5)
            if i == 0:
                rec[f"feat4_{k}"] = cap["feat4"].numpy().reshape(256, 4, 5)
    # blocks_to_run = 2 / 1 with prior ("iterative" model slot, HomographyNet.cpp:104-124)
    for btr in (2, 1):
        net.model_part1.blocks_to_run = btr
        i = 0
        i1 = torch.from_numpy(prev[i]).float().div(255.0).view(1, 1, 224, 320)
        i2 = torch.from_numpy(curr[i]).float().div(255.0).view(1, 1, 224, 320)
        flow, cov, err = run_case(net, i1, i2, torch.from_numpy(prior[i]).view(1, 1, 4, 2), MASK_SEED0 + i, cap)
        rec[f"flow_prior{btr}_0"] = flow.numpy().reshape(8)
        rec[f"cov_prior{btr}_0"] = cov.numpy()
    net.model_part1.blocks_to_run = 3
    np.savez_compressed(os.path.join(OUT, "e2e.npz"), **rec)

    # ---- stage-level known answers straight from the reference functions -------------------------
    g = torch.Generator().manual_seed(5)
    st = {"torch_version": np.array(torch.__version__)}
    pts0 = torch.from_numpy(S.ORIGIN_4PT).unsqueeze(0)
    offs = (torch.rand(8, 4, 2, generator=g) * 2 - 1) * 24.0
    offs[0] = 0.0                      # DLT(p, p) = I
    offs[1] = torch.tensor([3.5, -2.25]).expand(4, 2)   # pure translation
    Hs = torch.cat([m.DLT_solve(pts0, pts0 + offs[i:i + 1]) for i in range(8)], 0)   # model_to_trace.py:42-61
    st["dlt_offsets"] = offs.numpy()
    st["dlt_H"] = Hs.numpy()
    warper = refwarp.WarpImg(224, 320, "cpu") if False else p1.img_warper_full_size
    i2 = torch.from_numpy(curr[0]).float().div(255.0).view(1, 1, 224, 320)
    warped, ixs, iys = [], [], []
    for i in range(4):
        Hm = Hs[i + 2:i + 3]
        w = warper.warpSingleImage_H_Mtrx(i2, Hm)                                   # warp.py:60-79
        warped.append(w[0, 0].numpy())
        # sampling indices exactly as grid_sample derives them from the reference's grid (warp.py:65-70)
        uvz = torch.mm(Hm[0], warper.grid_uv1)
        uv = (uvz / uvz[2, :])[0:2].view(2, 224, 320)
        gn = uv.permute(1, 2, 0) * warper.sample_grid_factor - 1
        ix = ((gn[..., 0] + 1) / 2) * 319
        iy = ((gn[..., 1] + 1) / 2) * 223
        ixs.append(torch.floor(ix).to(torch.int16).numpy())
        iys.append(torch.floor(iy).to(torch.int16).numpy())
    st["warp_H"] = Hs[2:6].numpy()
    st["warp_src_u8"] = curr[0]
    st["warp_out"] = np.stack(warped).astype(np.float32)
    st["warp_ix"] = np.stack(ixs)
    st["warp_iy"] = np.stack(iys)
    var = torch.rand(1, 4, 2, generator=g) + 0.5
    ptsw = pts0 + (torch.rand(1, 4, 2, generator=g) * 2 - 1) * 3
    p2, cov4 = m.transfer_mean_var_single(var, Hs[3:4], ptsw)                        # model_to_trace.py:18-38
    st["tr_var"], st["tr_H"], st["tr_pts"] = var.numpy(), Hs[3:4].numpy(), ptsw.numpy()
    st["tr_p2"], st["tr_cov"] = p2.numpy(), cov4.numpy()
    np.savez_compressed(os.path.join(OUT, "stages.npz"), **st)
    for f in ("e2e.npz", "stages.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
