"""Summarise an ncu launch list (`--csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum`)
of `bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline` (or of the e2e forward): per-kernel launches / time / share /
DRAM bytes over the LAST `steps` steps; for the bench list also the mean DRAM traffic per conv_umma launch and per op
-> conv_umma_traffic.json (read by bench.py for `roofline.traffic`).
Usage: python profiles/summarize.py profiles/r02_launches_bench_steps2.csv [launches_per_step [steps [ops_per_step]]]"""
import collections
import csv
import json
import os
import sys


def rows_of(path):
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    per = collections.OrderedDict()
    for r in rd:
        d = per.setdefault(r["ID"], {"name": r["Kernel Name"], "t": 0.0, "rd": 0.0, "wr": 0.0})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            d["t"] = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)          # -> us
        else:
            b = v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            d["rd" if "read" in r["Metric Name"] else "wr"] = b
    return list(per.values())


import re

LABELS = {
    "conv_umma_kernel<1,0,0,1,0,0,0>": "quantizer fused (1x1 / stride 1; A-stationary when K > 256)",
    "conv_umma_kernel<1,0,0,1,0,0,1>": "fused stem (7x7 RGB: im2col rows built in shared memory)",
    "conv_umma_kernel<0,0,0,1,0,0,0>": "plain / halo (u8 NHWC workspace in, fp32 out)",
    "conv_umma_kernel<0,0,0,1,0,1,0>": "CTA pair (cta_group::2, deep reductions)",
    "conv_umma_kernel<0,1,1,1,0,0,0>": "conv3: residual + ReLU, fp32 and int8 out",
    "conv_umma_kernel<0,0,1,2,0,0,0>": "conv1 / conv2: int8 in, int8 out (two epilogue groups)",
    "conv_umma_kernel<0,0,1,1,0,0,0>": "int8 out, several channel tiles",
    "conv_umma_kernel<0,0,1,1,0,1,0>": "CTA pair, int8 out",
    "conv_umma_kernel<0,1,0,1,0,0,0>": "residual + ReLU, fp32 out (last block)",
}


def short(name):
    m = re.search(r"conv_umma_kernel<([^>]*)>", name)
    if m:
        return "conv_umma_kernel<" + m.group(1).replace(" ", "").replace("false", "0").replace("true", "1") + ">"
    m = re.search(r"(\w+_kernel)", name)
    return m.group(1) if m else name[:60]


def main():
    path = sys.argv[1]
    rows = rows_of(path)
    per_step = int(sys.argv[2]) if len(sys.argv) > 2 else None
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    ops_per_step = int(sys.argv[4]) if len(sys.argv) > 4 else 53
    if per_step:
        rows = rows[-steps * per_step:]
    agg = collections.OrderedDict()
    for r in rows:
        a = agg.setdefault(short(r["name"]), [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += r["t"]
        a[2] += r["rd"]
        a[3] += r["wr"]
    total = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | share | DRAM read GB | DRAM write GB | GB/s |\n|---|---|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        label = f"{k} ({LABELS[k]})" if k in LABELS else k
        print(f"| {label} | {a[0]} | {a[1]:.0f} | {a[1] / total:.3f} | {a[2] / 1e9:.2f} | {a[3] / 1e9:.2f} | {(a[2] + a[3]) / a[1] / 1e3:.0f} |")
    dram = sum(r["rd"] + r["wr"] for r in rows)
    print(f"total: {len(rows)} launches, {total:.0f} us, {dram / 1e9:.2f} GB of DRAM traffic ({steps} step(s))")
    conv = [r for r in rows if "conv_umma" in r["name"]]
    if "bench" in os.path.basename(path):
        out = {"dram_bytes_per_launch": int(sum(r["rd"] + r["wr"] for r in conv) / max(len(conv), 1)), "launches": len(conv),
               "op_dram_bytes_per_op": int(dram / (steps * ops_per_step)),
               "source": f"{path} (ncu dram__bytes_read.sum + dram__bytes_write.sum: mean over the conv_umma launches, and all "
                         f"quantizer + conv launches per op, of {steps} bench steps)"}
        with open(os.path.join(os.path.dirname(os.path.abspath(path)), "conv_umma_traffic.json"), "w") as f:
            json.dump(out, f, indent=1)
        print(json.dumps(out))
    print("conv share of kernel time: %.3f" % (sum(r["t"] for r in conv) / total))


if __name__ == "__main__":
    main()
