#!/usr/bin/env python
"""Turn the raw ncu outputs of one profiled step into the tracked summaries under profiles/.

    python tools/summarise_profiles.py gpurun_out/launches.csv /tmp/step_full.csv [bench.json] [--round r02] [--pairs 1024]

The first argument is the launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file …`), the second the
`ncu -i prof.ncu-rep --page raw --csv` dump of the `--set full` capture of the same command.  Both hold the 26 launches
of one 3-block step at 1024 pairs (tools/profile_step.py --batch 1024 --steps 2, -s 27 -c 26: one kernel zeroes the warp's cell array at create, 26 launches of
the warm-up step are skipped).
"""
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGES = ["dlt(prior)", "warp cells fill (current frames -> texture-gather array)", "warp+pool4 (block 2 input)", "block_2_1", "block_2_2", "block_2_3", "block_2_4", "fc8+dlt (block 2)",
          "warp+pool2 (block 3 input)", "block_3 front: conv7x7+conv5x5s2 fused (CTA pairs)", "block_3_2", "block_3_3",
          "block_3_4", "block_3_5", "fc8+dlt (block 3)", "warp (block 4 input)",
          "block_4 front: conv7x7+conv5x5s2 fused (CTA pairs)", "block_4_2", "block_4_3", "block_4_4", "block_4_5",
          "block_4_6", "mc mask bits", "mc fc1 GEMM (mean head)", "mc fc1 GEMM (uncertainty head)", "mc final"]
FULL_KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
             "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
             "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed",
             "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
             "l1tex__m_xbar2l1tex_read_bytes.sum",
             "smsp__issue_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
             "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
             "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
             "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]
UNIT = {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6, "ms": 1e3, "us": 1.0, "ns": 1e-3, "usecond": 1.0,
        "msecond": 1e3, "nsecond": 1e-3}


def short(name):
    name = name.replace("void ", "").replace("unnamed>::", "").replace("(anonymous namespace)::", "")
    return name.split("(")[0] if "<" not in name else name[:name.index(">(") + 1] if ">(" in name else name


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    out = []
    for r in rows:
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        t = float(d["Metric Value"].replace(",", "")) * UNIT.get(d["Metric Unit"], 1.0)
        out.append((short(d["Kernel Name"]), d["Grid Size"].replace(",", ""), d["Block Size"].replace(",", ""), t))
    return out


def main():
    argv = list(sys.argv)
    rnd, pairs = "r02", 1024
    for flag in ("--round", "--pairs"):
        if flag in argv:
            i = argv.index(flag)
            val = argv[i + 1]
            del argv[i:i + 2]
            if flag == "--round":
                rnd = val
            else:
                pairs = int(val)
    sys.argv = argv
    DESC = ("CTA-pair fused block fronts (2-row Toeplitz conv 1, N = 128; 64-byte-row input planes), TMA shifted-window + "
            "im2col-TMA igemm convs, fused MC GEMM, texture-gather warp")
    lpath, fpath = sys.argv[1], sys.argv[2]
    ls = launches(lpath)
    assert len(ls) == len(STAGES), (len(ls), len(STAGES))
    total = sum(x[3] for x in ls)
    with open(os.path.join(ROOT, "profiles", f"{rnd}_launches_step_b{pairs}.csv"), "w") as f:
        f.write(f"# ncu launch list — one 3-block UAHN step, {pairs} pairs, bf16 ({rnd}, final kernels: {DESC})\n")
        f.write("# command: ncu --metrics gpu__time_duration.sum --clock-control none -s 27 -c 26 --csv python "
                f"tools/profile_step.py --batch {pairs} --steps 2\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write("stage,kernel,grid,block,time_us,share\n")
        for s, (k, g, b, t) in zip(STAGES, ls):
            f.write(f'{s},"{k}",{g},{b},{t:.1f},{t / total:.3f}\n')
        f.write(f"total,,,,{total:.1f},1.000\n")
        groups = {"warp": 0.0, "conv": 0.0, "heads": 0.0}
        for s, x in zip(STAGES, ls):
            groups["warp" if s.startswith("warp") else "conv" if s.startswith("block") else "heads"] += x[3]
        f.write("# shares: " + ", ".join(f"{k} {v / total:.3f}" for k, v in groups.items()) + "\n")

    rows = list(csv.reader(open(fpath)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    assert len(data) == len(STAGES), len(data)
    idx = {k: hdr.index(k) for k in FULL_KEYS}
    conv_bytes = 0.0
    with open(os.path.join(ROOT, "profiles", f"{rnd}_ncu_step_full_b{pairs}.csv"), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on -s 27 -c 26 python tools/profile_step.py --batch "
                f"{pairs} --steps 2  ({rnd}, final kernels; one 3-block step, {pairs} pairs, bf16)\n")
        f.write("# per-launch values are cold-cache and serialised: compare SHARES, not absolutes.  Units: us, MB, MB, % of "
                "peak x3, MB (L2->SM), then % of peak\n")
        f.write("stage,kernel," + ",".join(FULL_KEYS) + "\n")
        for s, r in zip(STAGES, data):
            vals = []
            for k in FULL_KEYS:
                v, u = r[idx[k]].replace(",", ""), units[idx[k]]
                try:
                    x = float(v) * (UNIT.get(u, 1.0) if u in UNIT else 1.0)
                    vals.append(f"{x:.6f}" if "." in v or u in UNIT else v)
                except ValueError:
                    vals.append(v)
                    continue
                if s.startswith("block") and k.startswith("dram__bytes"):
                    conv_bytes += x * 1e6
            f.write(f'{s},"{short(r[hdr.index("Kernel Name")])}",' + ",".join(vals) + "\n")
    n_conv = sum(1 for s in STAGES if s.startswith("block"))
    tname = f"{rnd}_conv_traffic.json" if pairs == 1024 else f"{rnd}_conv_traffic_b{pairs}.json"
    json.dump({"pairs": pairs, "conv_group_dram_bytes_per_step": conv_bytes, "conv_launches": n_conv,
               "source": f"profiles/{rnd}_ncu_step_full_b{pairs}.csv (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum "
                         f"over the {n_conv} conv launches of one step)"},
              open(os.path.join(ROOT, "profiles", tname), "w"))
    if len(sys.argv) > 3:
        shutil.copy(sys.argv[3], os.path.join(ROOT, "profiles", f"{rnd}_bench_bf16.json"))
    print("launch total %.1f us, conv dram %.1f MB" % (total, conv_bytes / 1e6))


if __name__ == "__main__":
    main()
