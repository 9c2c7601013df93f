"""Packing microbench (BASELINE.json configs[4], north-star item (b)): tpack / tunpack HBM GB/s against the measured peak.

  python bench_pack.py [--log2n 27] [--iters 20]       one JSON line per (op, n_bits, dtype) + a summary line

x = randint(-2^(n-1), 2^(n-1), (2^27,), seed 0) as fp32 (what QuantConv2d.pack hands to tpack, reference
modelzoo/modules/quantconv2d.py:186-191) and as int8; n_bits in {2..8}.  Algorithmic bytes per element:
pack sizeof(dtype) + n/8, unpack n/8 + 1.  Device-timed with CUDA events through the C-ABI; inputs
(0.5 GB fp32) exceed L2, so every iteration streams from HBM.  Optionally times the unmodified reference kernels
(oracle/_ref) on the same tensors for context.
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=27)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--no-reference", action="store_true")
    a = ap.parse_args()
    import torch
    from quantize_b200 import capi
    if not torch.cuda.is_available():
        raise SystemExit("bench_pack.py needs a CUDA device (no CPU path)")
    L = capi.lib()
    peak = 6545.6
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    n = 1 << a.log2n
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device="cuda").manual_seed(0)
    ref = None
    if not a.no_reference:
        try:
            from oracle import build_ref
            ref = build_ref.load() if build_ref.available() else None
        except Exception:
            ref = None

    def timed(fn, iters):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e-3

    results = []
    for dt_name, dt, code, size in (("float32", torch.float32, capi.F32, 4), ("int8", torch.int8, capi.I8, 1)):
        for nb in (2, 3, 4, 5, 6, 7, 8):
            x = torch.randint(-(1 << (nb - 1)), 1 << (nb - 1), (n,), generator=g, device="cuda", dtype=torch.int32).to(dt)
            nbytes = int(L.qb200_packed_bytes(n, nb))
            packed = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
            flag = torch.zeros(1, dtype=torch.int32, device="cuda")
            out = torch.empty(n, dtype=torch.int8, device="cuda")

            def pack():
                capi.check(L.qb200_tpack(x.data_ptr(), code, n, nb, 1, packed.data_ptr(), flag.data_ptr(), st), "tpack")

            def unpack():
                capi.check(L.qb200_tunpack(packed.data_ptr(), n, nb, 1, out.data_ptr(), st), "tunpack")

            tp = timed(pack, a.iters)
            assert int(flag.item()) == 0
            tu = timed(unpack, a.iters)
            assert torch.equal(out.to(dt), x)                      # round trip at full size
            rec = {"dtype": dt_name, "n_bits": nb, "n": n,
                   "pack_gbs": round(n * (size + nb / 8) / tp / 1e9, 1), "unpack_gbs": round(n * (nb / 8 + 1) / tu / 1e9, 1)}
            rec["pack_frac"] = round(rec["pack_gbs"] / peak, 3)
            rec["unpack_frac"] = round(rec["unpack_gbs"] / peak, 3)
            if ref is not None and nb in (4, 8) and dt_name == "float32":
                tr = timed(lambda: ref.tpack(x, nb, True), 2)
                rp, rd = ref.tpack(x, nb, True)
                assert torch.equal(rp, packed)                     # same bytes as the compiled reference at 2^27
                tru = timed(lambda: ref.tunpack(rp, rd), 2)
                rec["reference_pack_gbs"] = round(n * (size + nb / 8) / tr / 1e9, 1)
                rec["reference_unpack_gbs"] = round(n * (nb / 8 + 1) / tru / 1e9, 1)
            results.append(rec)
            print(json.dumps(rec), flush=True)
            del x, packed, out
    f32 = [r for r in results if r["dtype"] == "float32"]
    print(json.dumps({"metric": "tensor_packing GB/s vs HBM peak", "peak_gbs": peak, "n": n,
                      "pack_f32_gbs_mean": round(sum(r["pack_gbs"] for r in f32) / len(f32), 1),
                      "unpack_gbs_mean": round(sum(r["unpack_gbs"] for r in results) / len(results), 1),
                      "pack_f32_frac_mean": round(sum(r["pack_frac"] for r in f32) / len(f32), 3),
                      "unpack_frac_mean": round(sum(r["unpack_frac"] for r in results) / len(results), 3)}), flush=True)


if __name__ == "__main__":
    main()
