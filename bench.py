"""bench.py — BASELINE.json's metric: ResNet-50 W8A8 images/s (conv stack through the fused hot path).

  python bench.py [--gpus N] [--steps K] [--warmup W]            one JSON line (this engine)
  python bench.py --impl reference [...]                          one JSON line (the reference's CPU path, host cores)
  torchrun --nproc-per-node N bench.py --gpus N ...               one process per GPU, weak scaling (256 images / GPU)

A step = one pass of the hot path over one batch: the 53 convolutions of ResNet-50 at batch 256, every layer through
qb200_conv_quantize_input + qb200_conv_from_workspace (the two kernels of quant_engine.quantconv2d_float_input's fused
path) on its own synthetic fp32 NCHW input that is resident in HBM when the timed region starts.
  value      images/s over all ranks, device-timed (two CUDA events around exactly K steps, max over ranks)
  e2e        the same metric through the public API a user calls — the packed ResNet-50 built from host.QuantConv2d
             layers, whose convs call quant_engine.quantconv2d_float_input — from PINNED HOST images to HOST logits,
             host<->device copies inside the timed region
  roofline   the op (activation-quantize + conv kernels): SURVEY 8(d) contract bytes per op / measured op duration vs the
             measured HBM peak; per-kernel durations come from K more steps instrumented with per-layer CUDA events
             (events between launches cost a few percent, so they stay out of the `value` loop)
  strong     BASELINE config 3 as written: a GLOBAL batch of 256 split over the ranks (256 / N images per GPU), with a
             cross-rank checksum of int32 accumulators (quantize_b200/dist.py) that verifies what the ranks computed
  packing    tpack / tunpack GB/s (fp32 and int8 inputs, 4 and 8 bits, 2^27 elements) vs the measured HBM peak
  cpu_baseline  the reference's CPU fake-quant conv path (oracle/fakequant.py port) on a bounded sample, rank 0, N=1
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL = "resnet50"
PER_GPU_BATCH = 256
W_BITS, A_BITS = 8, 8
INT8_PEAK_TOPS = 4500.0  # B200 dense int8 spec


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="images per GPU (BASELINE: 256)")
    ap.add_argument("--model", default=MODEL)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-chain", action="store_true", help="e2e: fp32 tensors between the convs of a block (no int8 hand-off)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layers", default="", help="debug: comma-separated layer indices to run")
    ap.add_argument("--per-layer", action="store_true", help="also print a per-layer table to stderr")
    ap.add_argument("--w-bits", type=int, default=8, help="weight bits (BASELINE headline: 8)")
    ap.add_argument("--a-bits", type=int, default=8, help="activation bits (BASELINE headline: 8)")
    ap.add_argument("--sweep-out", default="", help="write the per-unique-shape layer sweep (SURVEY 8d config 5) as markdown")
    ap.add_argument("--graph", action="store_true", help="replay one CUDA-graph capture of the step instead of launching it")
    ap.add_argument("--eager-e2e", action="store_true", help="e2e: launch the forward op by op instead of replaying CUDA graphs")
    ap.add_argument("--stem-chunks", type=int, default=1, help="e2e: H2D copy and stem (conv1 + max pool) in this many chunks of images per step")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling record (global batch 256 / N per GPU)")
    ap.add_argument("--no-packing", action="store_true", help="skip the tensor_packing GB/s record")
    a = ap.parse_args()
    global W_BITS, A_BITS
    W_BITS, A_BITS = a.w_bits, a.a_bits
    return a


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured", d
    return 6650.0, "fallback", {}


# ---------------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi) during the timed region
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def count_since(self, t_start):
        return sum(1 for t, _ in self.samples if t >= t_start)

    def stop(self, t_start=0.0):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, s in self.samples:
            if ts < t_start:
                continue
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# the reference arm / cpu baseline: reference's CPU fake-quant conv path (port), all host threads
# ---------------------------------------------------------------------------------------------------
def cpu_fakequant_stack(model, sample_batch, steps, warmup):
    import torch
    from quantize_b200 import models
    from oracle.fakequant import FakeQuantConvStack          # the one place bench.py executes oracle/
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    specs = models.conv_layer_specs(model, sample_batch)
    stack = FakeQuantConvStack(specs, sample_batch, W_BITS, A_BITS)
    sec = stack.time_steps(steps, warmup)
    return sample_batch / sec, sec, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 8
    ips, sec, cores = cpu_fakequant_stack(args.model, sample, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": ("ResNet-50 W8A8 images/sec" if (args.model, W_BITS, A_BITS) == ("resnet50", 8, 8)
                       else f"{args.model} W{W_BITS}A{A_BITS} images/sec"), "value": round(ips, 3), "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(sec * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.model} conv stack (53 convs) W{W_BITS}A{A_BITS} fake-quant on host CPU, "
                               f"bounded sample of {sample} images per step"},
        "cpu_baseline": {"value": round(ips, 3), "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} images/step x {args.steps} steps: Quantizer.simulate (act+weight) + fp32 "
                                   f"F.conv2d per layer (reference quantconv2d.py:154-168), torch {cores} threads"},
        "e2e": {"value": round(ips, 3), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# this engine
# ---------------------------------------------------------------------------------------------------
def im2col_kcol(C, R):
    """bytes per materialised im2col row of a few-channel layer (csrc/conv_common.cuh: im2col_kcol)"""
    kq = R * R * 4
    words = (kq + 127) // 128 * 128 if kq > 64 else (64 if kq > 32 else 32)
    k8 = C * R * 8
    grouped = (k8 + 63) // 64 * 64 if k8 > 64 else (64 if k8 > 32 else 32)
    return grouped if (R <= 8 and grouped < words) else words


class ConvStack:
    """per-layer device operands for the C-ABI hot path"""

    def __init__(self, specs, device, seed=0):
        import torch
        from quantize_b200 import capi, engine
        self.torch, self.capi = torch, capi
        self.L = capi.lib()
        qe = engine.load()
        g = torch.Generator(device=device).manual_seed(seed)
        self.layers = []
        max_ws = max_out = 0
        for i, s in enumerate(specs):
            x = torch.randn(s["N"], s["C"], s["H"], s["W"], generator=g, device=device)
            if s["relu_input"]:
                x.relu_()
            cg = s["C"] // s["groups"]
            wmax = (1 << (W_BITS - 1)) - 1                       # symmetric signed weights (range/minmax.py:123-135)
            qw = torch.randint(-wmax, wmax + 1, (s["K"], cg, s["R"], s["R"]), generator=g, device=device).float()
            packed, des = qe.tpack(qw, W_BITS, True)             # reference format (tpack.cu:203-255)
            shape = capi.conv_shape(s["N"], s["C"], s["H"], s["W"], s["K"], cg, s["R"], s["R"], s["stride"], s["pad"],
                                    W_BITS, 1)
            prepared = torch.empty(self.L.qb200_conv_prepared_bytes(ctypes.byref(shape)), dtype=torch.uint8, device=device)
            capi.check(self.L.qb200_conv_prepare_weights(ctypes.byref(shape), packed.data_ptr(), prepared.data_ptr(), None),
                       "prepare")
            qmax = float((1 << A_BITS) - 1)
            xmin, xmax = x.min(), x.max()
            a_scale = ((xmax - xmin) / qmax).reshape(1)
            a_zero = (xmin / a_scale).reshape(1)                  # range/minmax.py:136-143
            aq_t = [a_scale, a_zero, torch.zeros(1, device=device), torch.full((1,), qmax, device=device)]
            aq = capi.ActQuant(*[t.data_ptr() for t in aq_t])
            w_scale = torch.rand(s["K"], generator=g, device=device) * 1e-3 + 1e-4
            bias = torch.randn(s["K"], generator=g, device=device)
            P, Q = capi.conv_out_hw(shape)
            Cp = self.L.qb200_padded_channels(s["C"])
            max_ws = max(max_ws, self.L.qb200_conv_workspace_bytes(ctypes.byref(shape)))
            max_out = max(max_out, s["N"] * s["K"] * P * Q)
            ops = 2 * s["N"] * s["K"] * P * Q * cg * s["R"] * s["R"]
            single = bool(self.L.qb200_conv_is_single_kernel(ctypes.byref(shape), x.data_ptr()))
            chunked = int(self.L.qb200_conv_rows_chunk(ctypes.byref(shape))) > 0   # im2col rows in L2-sized chunks of images
            # strided 1x1 layers only touch the sampled pixels; few-channel layers go through im2col rows
            sub = s["R"] == 1 and s["stride"] > 1 and s["pad"] == 0
            in_pix = s["N"] * (P * Q if sub else s["H"] * s["W"])
            w_bytes = s["K"] * s["R"] * s["R"] * Cp + 12 * s["K"]
            if single:   # one kernel: fp32 in, fp32 out
                conv_bytes = 4 * s["N"] * s["C"] * s["H"] * s["W"] + w_bytes + 4 * s["N"] * s["K"] * P * Q
                quant_bytes = 0
            elif s["C"] <= 4 and s["R"] > 1:   # quantizer writes im2col rows [N*P*Q][Kcol]; the conv is a GEMM over them
                kcol = im2col_kcol(s["C"], s["R"])
                conv_bytes = s["N"] * P * Q * kcol + s["K"] * kcol + 12 * s["K"] + 4 * s["N"] * s["K"] * P * Q
                quant_bytes = 4 * s["N"] * s["C"] * s["H"] * s["W"] + s["N"] * P * Q * kcol
            else:        # quantizer kernel (fp32 in, u8 out) + conv kernel (u8 in, fp32 out)
                conv_bytes = in_pix * Cp + w_bytes + 4 * s["N"] * s["K"] * P * Q
                quant_bytes = 4 * s["C"] * in_pix + in_pix * Cp
            if chunked:      # one fused call (quantizer and conv kernels alternate per chunk): its bytes are both kernels'
                conv_bytes, quant_bytes = conv_bytes + quant_bytes, 0
            self.layers.append(dict(spec=s, x=x, shape=shape, prepared=prepared, aq=aq, aq_t=aq_t, w_scale=w_scale,
                                    bias=bias, ops=ops, conv_bytes=conv_bytes, quant_bytes=quant_bytes, packed=packed,
                                    single=single or chunked, chunked=chunked))
        self.ws = torch.empty(max_ws, dtype=torch.uint8, device=device)
        self.out = torch.empty(max_out, dtype=torch.float32, device=device)
        torch.cuda.synchronize()

    def step(self, stream, events=None):
        """one pass of the hot path.  Without `events` every layer is ONE call of the fused op's C-ABI entry point
        (qb200_quantconv2d_fused: what quant_engine.quantconv2d_float_input runs).  With `events` (per-layer: conv start,
        conv end, layer start) the two kernels of a two-kernel layer are launched through qb200_conv_quantize_input +
        qb200_conv_from_workspace — the same kernels — so that each can be timed; 1x1/stride-1 layers whose quantizer runs
        inside the conv kernel are one launch either way."""
        L, capi = self.L, self.capi
        for i, l in enumerate(self.layers):
            s = l["spec"]
            if events is None:
                capi.check(L.qb200_quantconv2d_fused(ctypes.byref(l["shape"]), l["x"].data_ptr(), l["prepared"].data_ptr(),
                                                     l["w_scale"].data_ptr(), s["K"], l["bias"].data_ptr(),
                                                     ctypes.byref(l["aq"]), self.ws.data_ptr(), self.out.data_ptr(),
                                                     capi.OUT_F32, stream), "quantconv2d_fused")
            else:
                events[i][2].record()
                if l["single"]:
                    events[i][0].record()
                    capi.check(L.qb200_quantconv2d_fused(ctypes.byref(l["shape"]), l["x"].data_ptr(), l["prepared"].data_ptr(),
                                                         l["w_scale"].data_ptr(), s["K"], l["bias"].data_ptr(),
                                                         ctypes.byref(l["aq"]), self.ws.data_ptr(), self.out.data_ptr(),
                                                         capi.OUT_F32, stream), "quantconv2d_fused")
                else:
                    capi.check(L.qb200_conv_quantize_input(ctypes.byref(l["shape"]), l["x"].data_ptr(), ctypes.byref(l["aq"]),
                                                           self.ws.data_ptr(), stream), "quantize_input")
                    events[i][0].record()
                    capi.check(L.qb200_conv_from_workspace(ctypes.byref(l["shape"]), self.ws.data_ptr(), l["prepared"].data_ptr(),
                                                           l["w_scale"].data_ptr(), s["K"], l["bias"].data_ptr(),
                                                           ctypes.byref(l["aq"]), self.out.data_ptr(), capi.OUT_F32, stream), "conv")
                events[i][1].record()
            if os.environ.get("QB200_BENCH_SYNC"):      # debugging aid: localise a failing launch
                try:
                    self.torch.cuda.synchronize()
                except Exception:
                    print(f"launch failure in layer {i}: {s} single={l['single']} watchdog={L.qb200_watchdog_code()}",
                          file=sys.stderr, flush=True)
                    raise

    def acc_checksum(self, layer_ids, lo, hi, stream):
        """int64 sum of the int32 accumulators (QB200_OUT_ACC) of images [lo, hi) of the given layers: a batch-shard
        invariant (integer sums are exact and order-free), used for the cross-rank verification."""
        torch, L, capi = self.torch, self.L, self.capi
        sums = []
        for i in layer_ids:
            l = self.layers[i]
            s = dict(l["spec"])
            n = hi - lo
            cg = s["C"] // s["groups"]
            shape = capi.conv_shape(n, s["C"], s["H"], s["W"], s["K"], cg, s["R"], s["R"], s["stride"], s["pad"], W_BITS, 1)
            P, Q = capi.conv_out_hw(shape)
            acc = torch.empty(n * s["K"] * P * Q, dtype=torch.int32, device=l["x"].device)
            x = l["x"][lo:hi].contiguous()
            capi.check(L.qb200_quantconv2d_fused(ctypes.byref(shape), x.data_ptr(), l["prepared"].data_ptr(), l["w_scale"].data_ptr(),
                                                 s["K"], l["bias"].data_ptr(), ctypes.byref(l["aq"]), self.ws.data_ptr(),
                                                 acc.data_ptr(), capi.OUT_ACC, stream), "quantconv2d_fused(acc)")
            sums.append(acc.sum(dtype=torch.int64))
        return torch.stack(sums)


def time_stack(torch, stack, stream, K, warmup, barrier, use_graph=False):
    """W untimed steps, then exactly K steps between two CUDA events (barrier + synchronize on both sides).
    Returns (ms for the K steps, engine launches inside the timed region)."""
    L = stack.L
    for _ in range(max(warmup, 1)):
        stack.step(stream)
    graph = None
    if use_graph:
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream()
        cap.wait_stream(torch.cuda.current_stream())
        with torch.cuda.graph(graph, stream=cap):
            stack.step(ctypes.c_void_p(cap.cuda_stream))
        torch.cuda.current_stream().wait_stream(cap)
        for _ in range(2):
            graph.replay()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    L.qb200_launch_count_reset()
    t0.record()
    if graph is not None:
        for _ in range(K):
            graph.replay()
    else:
        for _ in range(K):
            stack.step(stream)
    t1.record()
    barrier()
    launches = int(L.qb200_launch_count())
    if graph is not None:       # replays launch the captured kernels without passing through the library's counter
        L.qb200_launch_count_reset()
        stack.step(stream)
        torch.cuda.synchronize()
        launches = int(L.qb200_launch_count()) * K
    return t0.elapsed_time(t1), launches


def packing_record(torch, capi, hbm_peak):
    """tensor_packing GB/s (north-star item b): 2^27 signed values as fp32 (what QuantConv2d.pack hands to tpack) and as
    int8, 4 and 8 bits; algorithmic bytes sizeof(in) + n/8 (pack), n/8 + 1 (unpack); inputs >> L2."""
    L = capi.lib()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    n = 1 << 27
    g = torch.Generator(device="cuda").manual_seed(0)
    rec = {"n": n, "peak_gbs": hbm_peak}

    def timed(fn, iters=10):
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

    for dt_name, dt, code, size in (("f32", torch.float32, capi.F32, 4), ("i8", torch.int8, capi.I8, 1)):
        for nb in (4, 8):
            x = torch.randint(-(1 << (nb - 1)), 1 << (nb - 1), (n,), generator=g, device="cuda", dtype=torch.int32).to(dt)
            packed = torch.empty(int(L.qb200_packed_bytes(n, nb)), dtype=torch.uint8, device="cuda")
            flag = torch.zeros(1, dtype=torch.int32, device="cuda")
            out = torch.empty(n, dtype=torch.int8, device="cuda")
            tp = timed(lambda: capi.check(L.qb200_tpack(x.data_ptr(), code, n, nb, 1, packed.data_ptr(), flag.data_ptr(), st), "tpack"))
            tu = timed(lambda: capi.check(L.qb200_tunpack(packed.data_ptr(), n, nb, 1, out.data_ptr(), st), "tunpack"))
            ok = bool(torch.equal(out.to(dt), x)) and int(flag.item()) == 0     # round trip at full size
            rec[f"pack_{dt_name}_w{nb}"] = {"gbs": round(n * (size + nb / 8) / tp / 1e9, 1),
                                            "frac": round(n * (size + nb / 8) / tp / 1e9 / hbm_peak, 3)}
            rec[f"unpack_{dt_name}_w{nb}"] = {"gbs": round(n * (nb / 8 + 1) / tu / 1e9, 1),
                                              "frac": round(n * (nb / 8 + 1) / tu / 1e9 / hbm_peak, 3), "round_trip_ok": ok}
            del x, packed, out
    return rec


def run_b200(args):
    import torch
    from quantize_b200 import capi, models
    from quantize_b200 import dist as qdist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this engine has no CPU path (use --impl reference for the CPU arm)")
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    rank, world = qdist.init_from_env("nccl")
    if args.gpus != world and rank == 0 and world > 1:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE {world}", file=sys.stderr)
    n_gpus = world
    barrier = qdist.barrier

    specs = models.conv_layer_specs(args.model, args.batch)
    if args.layers:
        keep = [int(v) for v in args.layers.split(",")]
        specs = [specs[i] for i in keep]
    L = capi.lib()
    if os.environ.get("QB200_BENCH_ALGO"):               # A/B aid: 3 = plain two-kernel tensor-core path everywhere
        L.qb200_set_conv_algo(int(os.environ["QB200_BENCH_ALGO"]))
    stack = ConvStack(specs, device, seed=rank)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    K = args.steps
    nl = len(stack.layers)

    # ---- device-resident hot path: `value` ------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()          # nvidia-smi needs ~0.1-0.3 s to deliver its first line: start it before the warm-up
    for _ in range(max(args.warmup, 1)):
        stack.step(stream)
    torch.cuda.synchronize()
    t_load = time.time()     # only samples taken from here on (GPU under the benchmark's load) are reported
    ms, launches = time_stack(torch, stack, stream, K, 0 if not args.graph else 1, barrier, use_graph=args.graph)
    # ---- K more steps with per-layer events: the split into quantizer / conv kernel time -----------------
    events = [[tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(nl)] for _ in range(K)]
    i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    i0.record()
    for k in range(K):
        stack.step(stream, events[k])
    i1.record()
    torch.cuda.synchronize()
    # the timed region lasts tens of milliseconds; if the sampler caught fewer than 2 lines inside it, keep the same
    # load running (untimed) until it has, so that the reported clocks are clocks under this load
    extra = 0
    while sampler.count_since(t_load) < 2 and extra < 200 and sampler.proc is not None:
        stack.step(stream)
        torch.cuda.synchronize()
        extra += 1
    clocks = sampler.stop(t_load)
    clocks["extra_untimed_steps_for_sampling"] = extra
    instr_ms_per_step = i0.elapsed_time(i1) / K
    conv_ms = [sum(events[k][i][0].elapsed_time(events[k][i][1]) for k in range(K)) / K for i in range(nl)]
    quant_ms = [sum(events[k][i][2].elapsed_time(events[k][i][0]) for k in range(K)) / K for i in range(nl)]
    ms = qdist.max_over_ranks(ms, device)
    if world > 1:
        lc = torch.tensor([launches], device=device, dtype=torch.int64)
        torch.distributed.all_reduce(lc)
        launches = int(lc.item())
    ms_per_step = ms / K
    value = args.batch * n_gpus / (ms_per_step / 1e3)

    # ---- roofline (rank 0's kernels) ----------------------------------------------------------------------
    # The unit is the OP = activation-quantize + conv kernels of one layer; its algorithmic bytes are SURVEY 8(d)'s op
    # contract (fp32 NCHW in + fp32 NCHW out + packed weights + 12 B per output channel): 22.32 GB per ResNet-50 step.
    hbm_peak, peak_kind, peaks = measured_peaks()
    conv_total_ms = sum(conv_ms)
    op_total_ms = conv_total_ms + sum(q for q, l in zip(quant_ms, stack.layers) if not l["single"])
    conv_bytes = sum(l["conv_bytes"] for l in stack.layers)
    total_ops = sum(l["ops"] for l in stack.layers)
    contract_bytes = models.conv_stack_work(specs, W_BITS)[1]
    achieved = contract_bytes / (op_total_ms * 1e-3) / 1e9
    kernel_achieved = conv_bytes / (conv_total_ms * 1e-3) / 1e9
    traffic = kernel_dram_frac = None
    tp = os.path.join(ROOT, "profiles", "conv_umma_traffic.json")
    if os.path.exists(tp) and not args.layers and args.batch == PER_GPU_BATCH and args.model == MODEL:
        with open(tp) as f:
            tj = json.load(f)
        traffic = tj.get("op_dram_bytes_per_op", tj.get("dram_bytes_per_launch"))
        if tj.get("dram_bytes_per_launch"):
            kernel_dram_frac = round(tj["dram_bytes_per_launch"] / (conv_total_ms / nl * 1e-3) / 1e9 / hbm_peak, 4)
    int8_peak = None
    ip = os.path.join(ROOT, "profiles", "int8_peak.json")
    if os.path.exists(ip):
        with open(ip) as f:
            int8_peak = json.load(f).get("tops")
    step_tops = total_ops / (ms_per_step * 1e-3) / 1e12
    roofline = {"bound": "hbm", "kernel": "act_quantize_* + conv_umma_kernel (the op)", "achieved": round(achieved, 1),
                "peak": hbm_peak, "unit": "GB/s", "frac": round(achieved / hbm_peak, 4), "traffic": traffic,
                "peak_source": peak_kind,
                "bytes_per_op": round(contract_bytes / nl), "us_per_op": round(op_total_ms / nl * 1e3, 2), "ops_per_step": nl,
                "how": "contract bytes (SURVEY 8d) / (quantizer + conv kernel time), CUDA events around every kernel of K instrumented steps",
                "kernel_frac": round(kernel_achieved / hbm_peak, 4), "kernel_achieved_gbs": round(kernel_achieved, 1),
                "kernel_dram_frac": kernel_dram_frac,
                "kernel_note": "conv_umma_kernel alone on its own bytes (u8 NHWC or fp32 in + weights + fp32 out); kernel_dram_frac = ncu DRAM bytes of the same launches / their time",
                "conv_share_of_step": round(conv_total_ms / instr_ms_per_step, 4),
                "act_quantize_share_of_step": round((op_total_ms - conv_total_ms) / instr_ms_per_step, 4),
                "act_quantize_gbs": round(sum(l["quant_bytes"] for l in stack.layers) / (max(op_total_ms - conv_total_ms, 1e-9) * 1e-3) / 1e9, 1),
                "single_kernel_layers": sum(1 for l in stack.layers if l["single"] and not l["chunked"]),
                "instrumented_ms_per_step": round(instr_ms_per_step, 4),
                "tensor_tops": round(step_tops, 1),
                "tensor_frac_of_int8_spec": round(step_tops / INT8_PEAK_TOPS, 4),
                "tensor_frac_of_int8_measured": round(step_tops / int8_peak, 4) if int8_peak else None,
                "int8_peak_measured_tops": int8_peak,
                "step_contract_gbs": round(contract_bytes / (ms_per_step * 1e-3) / 1e9, 1),
                "step_contract_frac": round(contract_bytes / (ms_per_step * 1e-3) / 1e9 / hbm_peak, 4)}
    if args.per_layer and rank == 0:
        for i, l in enumerate(stack.layers):
            s = l["spec"]
            print(f"layer {i:2d} C{s['C']:4d} H{s['H']:3d} K{s['K']:4d} R{s['R']} s{s['stride']} {'1k' if l['single'] else '2k'} "
                  f"quant {quant_ms[i]*1e3:7.1f} us {l['quant_bytes']/max(quant_ms[i],1e-9)/1e6:6.0f} GB/s | conv {conv_ms[i]*1e3:7.1f} us "
                  f"{l['conv_bytes']/conv_ms[i]/1e6:6.0f} GB/s {l['ops']/conv_ms[i]/1e9:6.0f} TOPS", file=sys.stderr)

    if args.sweep_out and rank == 0:
        # per unique layer shape: time of the fused op (quantizer + conv), its contract bytes / ops, and how close it is
        # to its own roofline bound max(t_hbm, t_tensor)  (config 5)
        groups = {}
        for i, l in enumerate(stack.layers):
            sp = l["spec"]
            key = (sp["C"], sp["K"], sp["R"], sp["stride"], sp["H"], sp.get("groups", 1))
            g_ = groups.setdefault(key, {"n": 0, "t": 0.0, "ops": l["ops"], "single": l["single"], "spec": sp})
            g_["n"] += 1
            g_["t"] += (0.0 if l["single"] else quant_ms[i]) + conv_ms[i]
        with open(args.sweep_out, "w") as f:
            f.write(f"# layer sweep: {args.model} W{W_BITS}A{A_BITS}, {args.batch} images, 1 x B200 — fused op per unique shape\n\n")
            f.write("bytes = the op contract (fp32 in + fp32 out + packed weights + 12 B per output channel); bound = max(bytes / "
                    f"{hbm_peak:.0f} GB/s, ops / {INT8_PEAK_TOPS} TOPS); `frac` = bound / measured time.\n\n")
            f.write("| count | C->K | k | s | H | groups | kernels | us | GB/s (contract) | TOPS | % int8 spec | bound | frac of bound |\n")
            f.write("|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
            tot_t = tot_b = 0.0
            for key, g_ in groups.items():
                sp = g_["spec"]
                t_us = g_["t"] / g_["n"] * 1e3
                one = models.conv_stack_work([sp], W_BITS)
                ops, byts = one[0], one[1]
                t_hbm, t_tc = byts / (hbm_peak * 1e9) * 1e6, ops / (INT8_PEAK_TOPS * 1e12) * 1e6
                bound = max(t_hbm, t_tc)
                tot_t += g_["t"] * 1e3
                tot_b += bound * g_["n"]
                f.write(f"| {g_['n']} | {sp['C']}->{sp['K']} | {sp['R']} | {sp['stride']} | {sp['H']} | {sp.get('groups', 1)} | "
                        f"{1 if g_['single'] else 2} | {t_us:.1f} | {byts / t_us / 1e3:.0f} | {ops / t_us / 1e6:.0f} | "
                        f"{100 * ops / t_us / 1e6 / INT8_PEAK_TOPS:.1f} | {'hbm' if t_hbm >= t_tc else 'tensor'} | {bound / t_us:.2f} |\n")
            f.write(f"\nwhole stack: {tot_t:.0f} us measured, {tot_b:.0f} us at the per-layer bounds = {tot_b / tot_t:.2f}\n")

    # ---- strong scaling: BASELINE config 3 as written — a global batch of 256, 256 / N images per GPU ------------
    strong = None
    headline = (args.model, args.batch) == (MODEL, PER_GPU_BATCH) and not args.layers
    del stack
    torch.cuda.empty_cache()
    sstack = None
    if headline and not args.no_strong:
        lo, hi = qdist.shard_range(PER_GPU_BATCH, rank, world)
        if world == 1:
            s_ms_per_step, verify = ms_per_step, None
        else:
            # every rank builds the SAME global batch (seed 0) and keeps its slice: rows [r*256/N, (r+1)*256/N)
            gspecs = models.conv_layer_specs(args.model, PER_GPU_BATCH)
            gstack = ConvStack(gspecs, device, seed=0)
            vlayers = [0, 2, 13, 29, 48]                     # stem, 3x3 @56, 1x1 @28, 3x3 @14, 3x3 @7
            mine = gstack.acc_checksum(vlayers, lo, hi, stream)
            nlo, nhi = qdist.shard_range(PER_GPU_BATCH, (rank + 1) % world, world)
            neighbour = gstack.acc_checksum(vlayers, nlo, nhi, stream)   # this rank recomputes its neighbour's shard
            whole = gstack.acc_checksum(vlayers, 0, PER_GPU_BATCH, stream) if rank == 0 else None
            gathered = qdist.gather_batch(mine.reshape(1, -1), world)    # [world, layers] int64, via NCCL
            ok = bool(torch.equal(gathered[(rank + 1) % world], neighbour))
            if rank == 0:
                ok = ok and bool(torch.equal(gathered.sum(0), whole))    # shard sums add up to the unsharded batch
            okt = torch.tensor([1 if ok else 0], device=device)
            torch.distributed.all_reduce(okt, op=torch.distributed.ReduceOp.MIN)
            verify = {"cross_rank_checksum_ok": bool(okt.item()), "layers": vlayers,
                      "what": "int64 sums of int32 accumulators per shard: all-gathered (NCCL), each rank recomputes its "
                              "neighbour's shard, rank 0 checks that the shard sums add up to the unsharded batch",
                      "checksums_rank0_view": [int(v) for v in gathered.sum(0).tolist()]}
            del gstack, mine, neighbour, whole
            torch.cuda.empty_cache()
            sspecs = models.conv_layer_specs(args.model, hi - lo)
            sstack = ConvStack(sspecs, device, seed=rank)
            # 32 images per GPU: ~1 ms of kernels behind ~100 launches — replayed from one CUDA graph
            s_ms, _ = time_stack(torch, sstack, stream, K, max(args.warmup, 1), barrier, use_graph=True)
            s_ms_per_step = qdist.max_over_ranks(s_ms, device) / K
            del sstack
            torch.cuda.empty_cache()
        s_value = PER_GPU_BATCH / (s_ms_per_step / 1e3)
        strong = {"global_batch": PER_GPU_BATCH, "per_gpu_batch": hi - lo, "img_s": round(s_value, 1),
                  "ms_per_step": round(s_ms_per_step, 4),
                  # one GPU's rate on the full batch is what each rank of the weak run achieves: 256 / ms_per_step
                  "efficiency_vs_n1": round((s_value / n_gpus) / (PER_GPU_BATCH / (ms_per_step / 1e3)), 4),
                  "launch": "stream launches" if world == 1 and not args.graph else "cuda graph replay",
                  "verify": verify}

    # ---- end to end through the public op API: host images -> logits on host ----------------------------
    e2e = None
    if not args.no_e2e and not args.layers:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        net = models.build_packed(args.model, W_BITS, A_BITS, calib_batch=8, device=device, seed=0, fuse_blocks=True,
                                  chain_blocks=not args.no_chain, cross_block=not args.no_chain)
        hw = models.INPUT_HW[args.model]
        with torch.no_grad():
            n_classes = net(torch.zeros(1, 3, hw, hw, device=device)).shape[1]
        compute = torch.cuda.current_stream()
        copy_stream = torch.cuda.Stream(device=device)

        from quantize_b200 import host as qhost

        def run_e2e(batch):
            """every step: H2D of its inputs (pinned host memory), forward, D2H of its logits; the H2D of step k+1 runs on
            a copy stream into the second device buffer while step k computes.  The forward is replayed from CUDA graphs
            (host.GraphedForward: one graph per input buffer) unless --eager-e2e.  Returns ms per step (device events)."""
            host_in = torch.randn(batch, 3, hw, hw, generator=torch.Generator().manual_seed(100 + rank)).pin_memory()
            host_out = torch.empty(batch, n_classes, dtype=torch.float32).pin_memory()
            L.qb200_launch_count_reset()
            with torch.no_grad():
                net(torch.zeros(batch, 3, hw, hw, device=device))
            torch.cuda.synchronize()
            launches_per_forward = int(L.qb200_launch_count())
            graphed, why = None, None
            if not args.eager_e2e:
                try:
                    # --stem-chunks C: the images arrive in C chunks and the stem runs per chunk (it starts as soon as the
                    # first chunk has arrived).  Measured at C = 4: 48.4 k vs 48.8 k images/s over 20 steps, 47.4 k vs 47.3 k
                    # over 10 — the chunked stem costs per step what the shorter pipeline fill saves once — so the default is 1.
                    graphed = qhost.GraphedForward(net, torch.zeros(batch, 3, hw, hw, device=device), n_buffers=2,
                                                   stem_chunks=args.stem_chunks if batch % max(args.stem_chunks, 1) == 0 else 1)
                except Exception as ex:   # noqa: BLE001  (the eager path is the same computation)
                    why = f"{type(ex).__name__}: {ex}"[:200]
                    torch.cuda.synchronize()
            dev_in = graphed.inputs if graphed is not None else [torch.empty_like(host_in, device=device) for _ in range(2)]
            C = graphed.stem_chunks if graphed is not None else 1     # H2D chunks per step (the stem's chunks)
            cs = batch // C
            ev_copied = [[torch.cuda.Event() for _ in range(C)] for _ in range(2)]
            ev_used = [torch.cuda.Event() for _ in range(2)]

            def issue_copy(k):          # H2D of step k's images on the copy stream, into the buffer step k-2 has released
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(ev_used[k % 2])
                    for c in range(C):
                        dev_in[k % 2][c * cs:(c + 1) * cs].copy_(host_in[c * cs:(c + 1) * cs], non_blocking=True)
                        ev_copied[k % 2][c].record(copy_stream)

            def steps(n):
                issue_copy(0)
                for k in range(n):
                    if k + 1 < n:
                        issue_copy(k + 1)
                    if graphed is not None and C > 1:
                        for c in range(C):
                            compute.wait_event(ev_copied[k % 2][c])
                            graphed.replay_chunk(k % 2, c)
                        logits = graphed.replay_body(k % 2)
                    elif graphed is not None:
                        compute.wait_event(ev_copied[k % 2][0])
                        logits = graphed(k % 2)
                    else:
                        compute.wait_event(ev_copied[k % 2][0])
                        with torch.no_grad():
                            logits = net(dev_in[k % 2])
                    ev_used[k % 2].record(compute)
                    host_out.copy_(logits, non_blocking=True)

            for ev in ev_used:
                ev.record(compute)
            steps(max(args.warmup, 1))
            # the H2D copy alone (same buffers, nothing else running): names the bound of the e2e number at N GPUs
            barrier()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(3):
                dev_in[0].copy_(host_in, non_blocking=True)
            c1.record()
            barrier()
            h2d_gbs = host_in.numel() * 4 * 3 / (c0.elapsed_time(c1) * 1e-3) / 1e9
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            wall0 = time.perf_counter()
            e0.record()
            steps(K)
            e1.record()
            barrier()
            wall = time.perf_counter() - wall0
            e_ms = qdist.max_over_ranks(e0.elapsed_time(e1), device)
            mode = "cuda graph replay (host.GraphedForward)" if graphed is not None else ("eager" + (f" (graph capture failed: {why})" if why else ""))
            del graphed
            return e_ms / K, wall / K * 1e3, host_in.numel() * 4, host_out.numel() * 4, launches_per_forward, h2d_gbs, mode

        e_ms, wall_ms, h2d, d2h, e_launches, h2d_gbs, e_mode = run_e2e(args.batch)
        e2e = {"value": round(args.batch * n_gpus / (e_ms / 1e3), 1), "unit": "images/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": round(e_ms, 3), "wall_ms_per_step": round(wall_ms, 3),
               "h2d_gbs_per_gpu": round(h2d_gbs, 1),
               "h2d_ms_per_step_alone": round(h2d / (h2d_gbs * 1e9) * 1e3, 3),
               "api": "models.build_packed(resnet50, fuse_blocks=True, chain_blocks=%s) forward: host.QuantConv2d -> "
                      "quant_engine.quantconv2d_float_input / quantconv2d_chain (ReLU / residual add of each block run in the "
                      "conv epilogues; with chain_blocks the activations between the convs of a block stay int8)" % (not args.no_chain),
               "pipelining": "H2D of step k+1 (copy stream, double buffer) overlaps the forward of step k; all copies inside the timed region",
               "launch": e_mode,
               "engine_launches_per_step": e_launches}
        if strong is not None:
            if world == 1:
                strong["e2e_img_s"] = e2e["value"]
            else:
                lo, hi = qdist.shard_range(PER_GPU_BATCH, rank, world)
                se_ms = run_e2e(hi - lo)[0]
                strong["e2e_img_s"] = round(PER_GPU_BATCH / (se_ms / 1e3), 1)
                strong["e2e_ms_per_step"] = round(se_ms, 3)
        del net
        torch.cuda.empty_cache()

    packing = None
    if rank == 0 and n_gpus == 1 and not args.no_packing and not args.layers:
        packing = packing_record(torch, capi, hbm_peak)

    cpu_baseline = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline and not args.layers:
        ips, sec, cores = cpu_fakequant_stack(args.model, 8, 3, 1)
        cpu_baseline = {"value": round(ips, 3), "unit": "images/s", "cores": cores, "kind": "port",
                        "sample": "8 images/step x 3 steps (1 warm-up): Quantizer.simulate (act+weight) + fp32 F.conv2d "
                                  "per layer, the reference's CPU fake-quant path (quantconv2d.py:154-168)"}

    if rank == 0:
        line = {
            "metric": ("ResNet-50 W8A8 images/sec" if (args.model, W_BITS, A_BITS) == ("resnet50", 8, 8)
                       else f"{args.model} W{W_BITS}A{A_BITS} images/sec"), "value": round(value, 1), "unit": "images/s", "n_gpus": n_gpus,
            "steps": K, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8 x s8 -> s32 (fp32 dequant)", "data": "synthetic",
            "config": {"workload": f"{args.model} conv stack: {nl} convs via quantconv2d_float_input fused path "
                                   f"(act-quantize + tcgen05 int8 implicit GEMM + dequant), fp32 NCHW in/out per layer",
                       "per_gpu_batch": args.batch, "global_batch": args.batch * n_gpus, "w_bits": W_BITS, "a_bits": A_BITS,
                       "parallelism": f"batch-sharded x{n_gpus}, weights replicated, no collective on the data path",
                       "l2": "every layer streams its own input/output: 22.3 GB per step >> 126 MB L2, no flush needed",
                       "launch": "cuda graph replay" if args.graph else "stream launches (programmatic dependent launch)",
                       "ops_per_step": total_ops, "contract_bytes_per_step": contract_bytes},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "e2e": e2e, "strong": strong,
            "packing": packing, "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def _only_json_on_stdout():
    """NCCL / torch may print to fd 1 (e.g. 'NCCL version ...'); the contract is ONE JSON line on stdout.  Everything
    else is sent to stderr, and the JSON line is written to the real stdout at the end."""
    real = os.dup(1)
    os.dup2(2, 1)
    out = os.fdopen(real, "w")
    _print = print

    def json_print(*a, **k):
        if k.get("file") is None and a and isinstance(a[0], str) and a[0].startswith("{"):
            out.write(a[0] + "\n")
            out.flush()
        else:
            _print(*a, **k)
    return json_print


if __name__ == "__main__":
    print = _only_json_on_stdout()   # noqa: A001
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
