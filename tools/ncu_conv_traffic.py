#!/usr/bin/env python
"""Per-layer DRAM traffic of the backbone from an ncu CSV (long format, --csv --log-file) captured with
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:'conv3x3_tc|conv1_1_tc' ... python bench.py --steps 1 --warmup 1 ...
Takes the LAST 13 matching launches (one full batch-32 step) and writes profiles/r02_conv_traffic.{json,md}.
usage: python tools/ncu_conv_traffic.py gpurun_out/conv_traffic.csv"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LAYERS = [("conv1_1", 3, 64, 480, 640, False), ("conv1_2", 64, 64, 480, 640, True), ("conv2_1", 64, 128, 240, 320, False),
          ("conv2_2", 128, 128, 240, 320, True), ("conv3_1", 128, 256, 120, 160, False), ("conv3_2", 256, 256, 120, 160, False),
          ("conv3_3", 256, 256, 120, 160, True), ("conv4_1", 256, 512, 60, 80, False), ("conv4_2", 512, 512, 60, 80, False),
          ("conv4_3", 512, 512, 60, 80, True), ("conv5_1", 512, 512, 30, 40, False), ("conv5_2", 512, 512, 30, 40, False),
          ("conv5_3", 512, 512, 30, 40, False)]
B = 32


def main():
    src = sys.argv[1]
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = {}
    order = []
    for r in rows[1:]:
        if "conv3x3_tc" not in r[ki] and "conv1_1_tc" not in r[ki]:
            continue
        if r[ii] not in per:
            per[r[ii]] = {"kernel": r[ki].split("(")[0]}
            order.append(r[ii])
        per[r[ii]][r[mi]] = float(r[vi].replace(",", ""))
    step = [per[i] for i in order[-13:]]
    assert len(step) == 13, f"need 13 conv launches, found {len(step)}"
    table, tot, tot_alg = [], 0.0, 0.0
    for (name, cin, cout, h, w, pool), k in zip(LAYERS, step):
        rd, wr = k.get("dram__bytes_read.sum", 0.0), k.get("dram__bytes_write.sum", 0.0)
        oh, ow = (h // 2, w // 2) if pool else (h, w)
        in_b = B * h * w * cin * 4                      # fp32 NCHW images for conv1_1; hi+lo bf16 planes (4 B/value) after
        out_b = B * oh * ow * cout * 4
        w_b = 9 * cin * cout * 4
        alg = in_b + out_b + w_b
        table.append({"layer": name, "shape": f"{cin}->{cout} @{h}x{w}{' +pool' if pool else ''}", "kernel": k["kernel"],
                      "us": k.get("gpu__time_duration.sum", 0.0) / 1e3, "dram_read": rd, "dram_write": wr,
                      "algorithmic_bytes": alg, "ratio": (rd + wr) / alg})
        if name != "conv1_1":
            tot += rd + wr
            tot_alg += alg
    out = {"captured": os.path.basename(src), "batch": B, "layers": table, "dram_bytes_total": tot,
           "algorithmic_bytes_total": tot_alg, "vs_algorithmic": tot / tot_alg,
           "note": "totals cover the 12 conv3x3_tc_kernel launches (conv1_2..conv5_3); conv1_1 is listed separately"}
    json.dump(out, open(os.path.join(ROOT, "profiles", "r02_conv_traffic.json"), "w"), indent=1)
    with open(os.path.join(ROOT, "profiles", "r02_conv_traffic.md"), "w") as f:
        f.write(f"# Backbone DRAM traffic per layer, batch {B}, 480x640 (ncu: dram__bytes_read.sum + dram__bytes_write.sum; {out['captured']})\n\n")
        f.write("| layer | shape | ncu time us (cold, serialised) | DRAM read MB | DRAM write MB | algorithmic MB (in+out+weights) | ratio |\n|---|---|---|---|---|---|---|\n")
        for t in table:
            f.write(f"| {t['layer']} | {t['shape']} | {t['us']:.0f} | {t['dram_read']/1e6:.1f} | {t['dram_write']/1e6:.1f} | "
                    f"{t['algorithmic_bytes']/1e6:.1f} | {t['ratio']:.2f} |\n")
        f.write(f"\n12 conv3x3_tc launches: {tot/1e6:.0f} MB DRAM vs {tot_alg/1e6:.0f} MB algorithmic = {tot/tot_alg:.2f}x\n")
    print(open(os.path.join(ROOT, "profiles", "r02_conv_traffic.md")).read())


if __name__ == "__main__":
    main()
