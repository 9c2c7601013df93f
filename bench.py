"""bench.py — BASELINE.json's metric: ResNet-50 W8A8 images/s (conv stack through the fused hot path).

  python bench.py [--gpus N] [--steps K] [--warmup W]            one JSON line (this engine)
  python bench.py --impl reference [...]                          one JSON line (the reference's CPU path, host cores)
  torchrun --nproc-per-node N bench.py --gpus N ...               one process per GPU, weak scaling (256 images / GPU)

A step = one pass of the hot path over one batch: the 53 convolutions of ResNet-50 at batch 256, every layer through
qb200_conv_quantize_input + qb200_conv_from_workspace (the two kernels of quant_engine.quantconv2d_float_input's fused
path) on its own synthetic fp32 NCHW input that is resident in HBM when the timed region starts.
  value      images/s over all ranks, device-timed (CUDA events, max over ranks)
  e2e        the same metric through the public API a user calls — the packed ResNet-50 built from host.QuantConv2d
             layers, whose convs call quant_engine.quantconv2d_float_input — from PINNED HOST images to HOST logits,
             host<->device copies inside the timed region
  roofline   the dominant kernel (conv_umma_kernel): algorithmic bytes per launch / measured duration vs measured HBM peak
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
            self.layers.append(dict(spec=s, x=x, shape=shape, prepared=prepared, aq=aq, aq_t=aq_t, w_scale=w_scale,
                                    bias=bias, ops=ops, conv_bytes=conv_bytes, quant_bytes=quant_bytes, packed=packed, single=single))
        self.ws = torch.empty(max_ws, dtype=torch.uint8, device=device)
        self.out = torch.empty(max_out, dtype=torch.float32, device=device)
        torch.cuda.synchronize()

    def step(self, stream, events=None):
        """one pass of the hot path; events: optional per-layer (conv start, conv end, layer start) CUDA events.
        1x1/stride-1 layers run as ONE kernel (quantizer fused into the conv's producer warps): for them the whole call
        is the conv kernel; the other layers run the quantizer kernel and the conv kernel."""
        L, capi = self.L, self.capi
        for i, l in enumerate(self.layers):
            s = l["spec"]
            if events is not None:
                events[i][2].record()
            if l["single"]:
                if events is not None:
                    events[i][0].record()
                capi.check(L.qb200_quantconv2d_fused(ctypes.byref(l["shape"]), l["x"].data_ptr(), l["prepared"].data_ptr(),
                                                     l["w_scale"].data_ptr(), s["K"], l["bias"].data_ptr(),
                                                     ctypes.byref(l["aq"]), self.ws.data_ptr(), self.out.data_ptr(),
                                                     capi.OUT_F32, stream), "quantconv2d_fused")
            else:
                capi.check(L.qb200_conv_quantize_input(ctypes.byref(l["shape"]), l["x"].data_ptr(), ctypes.byref(l["aq"]),
                                                       self.ws.data_ptr(), stream), "quantize_input")
                if events is not None:
                    events[i][0].record()
                capi.check(L.qb200_conv_from_workspace(ctypes.byref(l["shape"]), self.ws.data_ptr(), l["prepared"].data_ptr(),
                                                       l["w_scale"].data_ptr(), s["K"], l["bias"].data_ptr(),
                                                       ctypes.byref(l["aq"]), self.out.data_ptr(), capi.OUT_F32, stream), "conv")
            if events is not None:
                events[i][1].record()
            if os.environ.get("QB200_BENCH_SYNC"):      # debugging aid: localise a failing launch
                try:
                    self.torch.cuda.synchronize()
                except Exception:
                    print(f"launch failure in layer {i}: {s} single={l['single']} watchdog={L.qb200_watchdog_code()}",
                          file=sys.stderr, flush=True)
                    raise


def run_b200(args):
    import torch
    import torch.distributed as dist
    from quantize_b200 import capi, models

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this engine has no CPU path (use --impl reference for the CPU arm)")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    if args.gpus != world and rank == 0 and world > 1:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE {world}", file=sys.stderr)
    n_gpus = world

    specs = models.conv_layer_specs(args.model, args.batch)
    if args.layers:
        keep = [int(v) for v in args.layers.split(",")]
        specs = [specs[i] for i in keep]
    L = capi.lib()
    if os.environ.get("QB200_BENCH_ALGO"):               # A/B aid: 3 = plain two-kernel tensor-core path everywhere
        L.qb200_set_conv_algo(int(os.environ["QB200_BENCH_ALGO"]))
    stack = ConvStack(specs, device, seed=rank)
    stream_obj = torch.cuda.current_stream()
    stream = ctypes.c_void_p(stream_obj.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident hot path -------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()          # nvidia-smi needs ~0.1-0.3 s to deliver its first line: start it before the warm-up
    for _ in range(max(args.warmup, 1)):
        stack.step(stream)
    K = args.steps
    nl = len(stack.layers)
    events = [[tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(nl)] for _ in range(K)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_load = time.time()     # only samples taken from here on (GPU under the benchmark's load) are reported
    L.qb200_launch_count_reset()
    t0.record()
    for k in range(K):
        stack.step(stream, events[k])
    t1.record()
    barrier()
    launches = int(L.qb200_launch_count())
    # the timed region lasts tens of milliseconds; if the sampler caught fewer than 2 lines inside it, keep the same
    # load running (untimed) until it has, so that the reported clocks are clocks under this load
    extra = 0
    while sampler.count_since(t_load) < 2 and extra < 200 and sampler.proc is not None:
        stack.step(stream)
        torch.cuda.synchronize()
        extra += 1
    clocks = sampler.stop(t_load)
    clocks["extra_untimed_steps_for_sampling"] = extra
    ms = t0.elapsed_time(t1)
    conv_ms = [sum(events[k][i][0].elapsed_time(events[k][i][1]) for k in range(K)) / K for i in range(nl)]
    quant_ms = [sum(events[k][i][2].elapsed_time(events[k][i][0]) for k in range(K)) / K for i in range(nl)]
    if world > 1:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lc = torch.tensor([launches], device=device, dtype=torch.int64)
        dist.all_reduce(lc)
        launches = int(lc.item())
    ms_per_step = ms / K
    value = args.batch * n_gpus / (ms_per_step / 1e3)

    # ---- roofline of the dominant kernel (conv_umma_kernel), rank 0 ------------------------------------
    hbm_peak, peak_kind, peaks = measured_peaks()
    conv_total_ms = sum(conv_ms)
    conv_bytes = sum(l["conv_bytes"] for l in stack.layers)
    total_ops = sum(l["ops"] for l in stack.layers)
    contract_bytes = models.conv_stack_work(specs, W_BITS)[1]
    achieved = conv_bytes / nl / (conv_total_ms / nl * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "conv_umma_traffic.json")
    if os.path.exists(tp) and not args.layers and args.batch == PER_GPU_BATCH:
        with open(tp) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "conv_umma_kernel", "achieved": round(achieved, 1), "peak": hbm_peak,
                "unit": "GB/s", "frac": round(achieved / hbm_peak, 4), "traffic": traffic, "peak_source": peak_kind,
                "bytes_per_launch": round(conv_bytes / nl), "us_per_launch": round(conv_total_ms / nl * 1e3, 2),
                "launches_per_step": nl,
                "conv_share_of_step": round(conv_total_ms / ms_per_step, 4),
                "act_quantize_share_of_step": round(sum(quant_ms) / ms_per_step, 4),
                "act_quantize_gbs": round(sum(l["quant_bytes"] for l in stack.layers) / (max(sum(q for q, l in zip(quant_ms, stack.layers) if not l["single"]), 1e-9) * 1e-3) / 1e9, 1),
                "single_kernel_layers": sum(1 for l in stack.layers if l["single"]),
                "tensor_tops": round(total_ops / (conv_total_ms * 1e-3) / 1e12, 1),
                "tensor_frac_of_int8_spec": round(total_ops / (conv_total_ms * 1e-3) / 1e12 / INT8_PEAK_TOPS, 4),
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

    # ---- end to end through the public op API: host images -> logits on host ----------------------------
    e2e = None
    if not args.no_e2e and not args.layers:
        del stack
        torch.cuda.empty_cache()
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        net = models.build_packed(args.model, W_BITS, A_BITS, calib_batch=8, device=device, seed=0, fuse_blocks=True,
                                  chain_blocks=not args.no_chain, cross_block=not args.no_chain)
        hw = models.INPUT_HW[args.model]
        host_in = torch.randn(args.batch, 3, hw, hw, generator=torch.Generator().manual_seed(100 + rank)).pin_memory()
        with torch.no_grad():
            n_classes = net(torch.zeros(1, 3, hw, hw, device=device)).shape[1]
        host_out = torch.empty(args.batch, n_classes, dtype=torch.float32).pin_memory()
        dev_in = [torch.empty_like(host_in, device=device) for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=device)
        compute = torch.cuda.current_stream()
        ev_copied = [torch.cuda.Event() for _ in range(2)]
        ev_used = [torch.cuda.Event() for _ in range(2)]

        def issue_copy(k):          # H2D of step k's images on the copy stream, into the buffer step k-2 has released
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ev_used[k % 2])
                dev_in[k % 2].copy_(host_in, non_blocking=True)
                ev_copied[k % 2].record(copy_stream)

        def e2e_steps(n):           # every step: H2D of its inputs, forward, D2H of its logits; copies overlap compute
            issue_copy(0)
            for k in range(n):
                if k + 1 < n:
                    issue_copy(k + 1)
                compute.wait_event(ev_copied[k % 2])
                with torch.no_grad():
                    logits = net(dev_in[k % 2])
                ev_used[k % 2].record(compute)
                host_out.copy_(logits, non_blocking=True)

        for ev in ev_used:
            ev.record(compute)
        e2e_steps(max(args.warmup, 1))
        barrier()
        L.qb200_launch_count_reset()
        wall0 = time.perf_counter()
        t0.record()
        e2e_steps(K)
        t1.record()
        barrier()
        wall = time.perf_counter() - wall0
        e_ms = t0.elapsed_time(t1)
        if world > 1:
            t = torch.tensor([e_ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = float(t.item())
        e2e = {"value": round(args.batch * n_gpus / (e_ms / K / 1e3), 1), "unit": "images/s",
               "h2d_bytes_per_step": host_in.numel() * 4, "d2h_bytes_per_step": host_out.numel() * 4,
               "ms_per_step": round(e_ms / K, 3), "wall_ms_per_step": round(wall / K * 1e3, 3),
               "api": "models.build_packed(resnet50, fuse_blocks=True, chain_blocks=%s) forward: host.QuantConv2d -> "
                      "quant_engine.quantconv2d_float_input / quantconv2d_chain (ReLU / residual add of each block run in the "
                      "conv epilogues; with chain_blocks the activations between the convs of a block stay int8)" % (not args.no_chain),
               "pipelining": "H2D of step k+1 (copy stream, double buffer) overlaps the forward of step k; all copies inside the timed region",
               "engine_launches_per_step": int(L.qb200_launch_count()) // K}

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
                       "ops_per_step": total_ops, "contract_bytes_per_step": contract_bytes},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


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
