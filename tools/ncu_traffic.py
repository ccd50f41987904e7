"""Per-kernel-family DRAM traffic of one embedding pass from an ncu launch list.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \\
        --log-file gpurun_out/traffic.csv python tools/profile_embed.py moco_aug_uber_34 192 3 1
    python tools/ncu_traffic.py gpurun_out/traffic.csv 576 profiles/r02_dram_traffic_uber34x3.json

Sums dram__bytes_read + dram__bytes_write over the launches of every kernel family (conv = every tcgen05 convolution
kernel, preprocess, heads, pooling) and divides by the number of frames: the figures bench.py reports as
`roofline.traffic` (scaled to the frames of a pass) next to the algorithmic bytes. The warm-up launches of the driver
(EmbeddingNet construction runs no kernels; `embed` is called once) are all part of the one pass.
"""
import csv
import json
import sys

FAMILIES = (("conv", ("conv_gemm_kernel", "conv3x3_patch_kernel", "conv_b2b")), ("preprocess", ("preprocess",)),
            ("heads", ("head_tail", "flatten")), ("pool", ("maxpool", "avgpool")))


def main():
    path, frames, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[h]
    per = {}
    for r in rows[h + 1:]:
        if len(r) < len(hdr):
            continue
        d = dict(zip(hdr, r))
        k = per.setdefault(d["ID"], {"name": d["Kernel Name"], "read": 0.0, "write": 0.0, "ns": 0.0})
        v = float(d["Metric Value"].replace(",", ""))
        unit = d["Metric Unit"].lower()
        scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6,
                 "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6}.get(unit, 1.0)
        if d["Metric Name"] == "dram__bytes_read.sum":
            k["read"] += v * scale
        elif d["Metric Name"] == "dram__bytes_write.sum":
            k["write"] += v * scale
        elif d["Metric Name"] == "gpu__time_duration.sum":
            k["ns"] += v * scale
    fam = {f: {"launches": 0, "read": 0.0, "write": 0.0, "ns": 0.0} for f, _ in FAMILIES}
    fam["other"] = {"launches": 0, "read": 0.0, "write": 0.0, "ns": 0.0}
    for k in per.values():
        f = next((f for f, pats in FAMILIES if any(p in k["name"] for p in pats)), "other")
        fam[f]["launches"] += 1
        for x in ("read", "write", "ns"):
            fam[f][x] += k[x]
    res = {"frames": frames, "source": path, "families": fam}
    for f in fam:
        res[f + "_bytes_per_frame"] = (fam[f]["read"] + fam[f]["write"]) / frames
    json.dump(res, open(out, "w"), indent=1)
    for f, v in fam.items():
        print(f"{f:10s} {v['launches']:4d} launches  {v['read'] / 1e6:9.1f} MB read  {v['write'] / 1e6:9.1f} MB written  "
              f"{(v['read'] + v['write']) / frames / 1e6:7.3f} MB/frame  {v['ns'] / 1e3:9.1f} us (serialised, cold)")


if __name__ == "__main__":
    main()
