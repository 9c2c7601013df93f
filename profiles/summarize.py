"""Summarise an ncu launch list (`--csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum`)
of `bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline`: per-kernel launches / time / share / DRAM bytes over the
LAST `--steps` steps, and the mean DRAM traffic per conv_umma launch -> conv_umma_traffic.json (read by bench.py for
`roofline.traffic`).  Usage: python profiles/summarize.py profiles/r01_launches_bench_steps2.csv [launches_per_step]"""
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


def short(name):
    for key in ("conv_umma_kernel<1>", "conv_umma_kernel<0>", "act_quantize_nhwc_vec4_kernel",
                "act_quantize_nhwc_kernel", "act_quantize_im2col8_kernel", "act_quantize_im2col_kernel", "zero_pad_borders_kernel", "maxpool2d_kernel"):
        if key in name:
            return key
    return name[:60]


def main():
    path = sys.argv[1]
    rows = rows_of(path)
    per_step = int(sys.argv[2]) if len(sys.argv) > 2 else None
    steps = 2
    if per_step:
        rows = rows[-steps * per_step:]
    agg = collections.OrderedDict()
    for r in rows:
        a = agg.setdefault(short(r["name"]), [0, 0.0, 0.0])
        a[0] += 1
        a[1] += r["t"]
        a[2] += r["rd"] + r["wr"]
    total = sum(a[1] for a in agg.values())
    print("| kernel | launches | total ms | share | DRAM GB |\n|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {a[0]} | {a[1] / 1e3:.3f} | {a[1] / total:.3f} | {a[2] / 1e9:.2f} |")
    conv = [r for r in rows if "conv_umma" in r["name"]]
    out = {"dram_bytes_per_launch": int(sum(r["rd"] + r["wr"] for r in conv) / max(len(conv), 1)), "launches": len(conv),
           "source": f"{path} (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over the conv_umma launches of {steps} bench steps)"}
    with open(os.path.join(os.path.dirname(os.path.abspath(path)), "conv_umma_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))
    print("conv share of kernel time: %.3f" % (sum(r["t"] for r in conv) / total))


if __name__ == "__main__":
    main()
