#!/usr/bin/env python
"""Benchmark of the DiT backbone forward (BASELINE.json metric: images/sec, bf16).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload base224|base512|large224]

* A "step" is one backbone forward over one batch of synthetic PubLayNet-shaped pages.
  N=1 workload (``base224``) = BASELINE.json configs[1]: DiT-base, batch 64, 224x224.
* ``value``  : whole-job images/s with the page batch already resident in HBM (the launch plan
  enqueued with programmatic dependent launch, or ``--graph``: CUDA-graph replay; per-step CUDA
  events, L2 flushed between steps).
* ``e2e``    : the same metric through the public module call ``DiTBackbone(...)(x)`` with
  HOST inputs: every step copies its fp16 pages from pinned host memory and reads the p5 tap
  back to the host, inside the timed region.
* ``roofline``: the dominant kernel (MLP up-projection tcgen05 GEMM, M x 3072 x 768 + erf-GELU)
  timed live with CUDA events around each of its launches inside eager forwards.
* ``cpu_baseline`` / ``--impl reference``: the reference's PyTorch-eager CPU forward
  (transformers BeitModel wrapped as R:dit_backbone.py:38-62, oracle/hf_reference.py) on the
  box's host cores, on a bounded sample of the same workload.
For N > 1 launch with torchrun (one rank per GPU); images shard by batch, every rank runs
the full backbone and the p5 taps are all-gathered over NCCL each step.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (config factory name, per-GPU batch, H, W)
    "base224": ("dit_base", 64, 224, 224),     # BASELINE.json configs[1]
    "base512": ("dit_base", 32, 512, 512),     # configs[2] geometry (per-GPU batch kept at 32)
    "large224": ("dit_large", 64, 224, 224),   # configs[3] geometry
}
CPU_SAMPLE_IMAGES = 8


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"],
                    bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


def profiled_traffic(kernel_regex: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the kernel, from the newest committed
    `ncu --set full` summary under profiles/ (tools/summarize_profiles.py); None if there is none."""
    import glob
    import re
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_full_summary.txt")), key=os.path.getmtime)
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for path in reversed(files):
        blocks = open(path).read().split("kernel: ")
        for blk in blocks[1:]:
            if not re.search(kernel_regex, blk.splitlines()[0]):
                continue
            tot = 0.0
            for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                m = re.search(re.escape(key) + r"\s+([0-9.,]+)\s+(\w+)", blk)
                if not m:
                    break
                tot += float(m.group(1).replace(",", "")) * unit.get(m.group(2), 1.0)
            else:
                return {"bytes": tot, "source": os.path.relpath(path, ROOT)}
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def cpu_reference_run(workload: str, steps: int, warmup: int, budget_s: float = 25.0):
    """Reference PyTorch-eager CPU forward on a bounded sample.  Returns (img/s, ms/step, steps, cores, sample)."""
    from layoutdit_b200 import config as cfgmod
    from layoutdit_b200.synth import make_state_dict, synthetic_pages
    from oracle import hf_reference   # the reference's CPU arithmetic; only ever the baseline, never the product

    fac, _, H, W = WORKLOADS[workload]
    cfg = getattr(cfgmod, fac)()
    model = hf_reference.build(cfg.to_dict(), make_state_dict(cfg, 0, False))
    x = synthetic_pages(CPU_SAMPLE_IMAGES, H, W, 1234)
    # all the host threads this process may use (torchrun pins OMP_NUM_THREADS=1 for its children)
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    torch.set_num_threads(max(torch.get_num_threads(), avail))
    cores = torch.get_num_threads()
    times = []
    with torch.no_grad():
        for _ in range(max(1, warmup)):
            t0 = time.perf_counter(); model(x); w = time.perf_counter() - t0
        # keep the whole run bounded: fewer timed steps if one step is slow on this host
        steps = max(1, min(steps, int(budget_s / max(w, 1e-3))))
        for _ in range(steps):
            t0 = time.perf_counter(); model(x); times.append(time.perf_counter() - t0)
    ms = 1e3 * statistics.median(times)
    sample = (f"{CPU_SAMPLE_IMAGES} of the workload's images per step ({workload}: {fac} {H}x{W}), fp32 eager, "
              f"{steps} timed steps, median; os.cpu_count()={os.cpu_count()}")
    return CPU_SAMPLE_IMAGES / (ms / 1e3), ms, steps, cores, sample


def run_reference(args, rank):
    if rank != 0:
        return
    fac, B, H, W = WORKLOADS[args.workload]
    v, ms, steps, cores, sample = cpu_reference_run(args.workload, args.steps, args.warmup, budget_s=120.0)
    print(json.dumps({
        "impl": "reference", "metric": "DiT backbone forward throughput", "value": round(v, 3), "unit": "images/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": round(ms, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {fac} forward, {H}x{W}, CPU sample of {CPU_SAMPLE_IMAGES} images/step"},
        "cpu_baseline": {"value": round(v, 3), "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(v, 3), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), file=_RESULT, flush=True)


def run_fpn_line(args, cfg, fac, B, H, W, dev, rank, world, peaks):
    """Secondary line for SURVEY 8 row f1: DiTWithFPN forward (graph replay), device-resident pages, same timing rules."""
    import torch.distributed as dist
    from layoutdit_b200 import DiTWithFPN
    from layoutdit_b200.config import flops_per_image
    from layoutdit_b200.synth import make_fpn_state_dict, make_state_dict, synthetic_pages
    model = DiTWithFPN(pretrained=False, config=cfg, state_dict=make_state_dict(cfg, 0, False),
                       fpn_state_dict=make_fpn_state_dict(cfg.hidden_size, 256, 0, False), use_cuda_graph=args.graph).to(dev).eval()
    x = synthetic_pages(B, H, W, 1234 + rank).to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    feats = model(x)
    eng = model.backbone._get_engine()
    launches = eng._geometry(B, H, W, 0, "fpn").launches
    for _ in range(args.warmup):
        model(x)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for s_, e_ in ev:
        flush.zero_()
        s_.record(); model(x); e_.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    total_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    Gh, Gw, C = H // 16, W // 16, 256
    pix = sum(int(Gh * s) * int(Gw * s) for s in (4.0, 2.0, 1.0, 0.5))
    fpn_flops = 2.0 * 4 * Gh * Gw * cfg.hidden_size * C + 2.0 * pix * 9 * C * C      # laterals on the token grid + 3x3 convolutions
    value = world * B * args.steps / (total_ms / 1e3)
    if rank == 0:
        print(json.dumps({
            "metric": "DiT backbone + FPN forward throughput", "value": round(value, 1), "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(total_ms / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.workload}+fpn: {fac} backbone + FeaturePyramidNetwork(256) forward, batch {B} per GPU, {H}x{W}, "
                                   f"random-init weights, p2..p5 + pool written", "global_batch": world * B,
                       "l2": "256 MiB buffer written between timed steps (outside the events)", "parallelism": f"dp{world}",
                       "timing": "CUDA-graph replay; per-step CUDA events summed; max over ranks"},
            "flops_per_image": flops_per_image(cfg, H, W) + fpn_flops,
            "model_tflops": round((flops_per_image(cfg, H, W) + fpn_flops) * value / 1e12, 1),
            "gpu_launches": launches * args.steps, "launches_per_step": launches,
            "outputs": {k: list(v.shape) for k, v in feats.items()}}), file=_RESULT, flush=True)
    if world > 1:
        dist.destroy_process_group()


_RESULT = sys.stdout   # where the ONE JSON line goes; main() re-points it at the real stdout


def isolate_stdout():
    """Keep stdout for the result line only: NCCL prints its version banner on fd 1 (and other libraries may chat
    there too), which would precede the JSON line.  fd 1 is re-pointed at stderr; the JSON goes to a dup of the
    original."""
    global _RESULT
    sys.stdout.flush()
    _RESULT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def main():
    isolate_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="base224", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--head", default="taps", choices=["taps", "fpn"],
                    help="taps (default, BASELINE.json's metric): DiTBackbone, four D-channel taps; fpn: DiTWithFPN "
                         "(SURVEY 8 row f1: laterals, top-down merges, 3x3 convolutions, pool) -- device-resident value only")
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="enqueue the launch plan on the stream every step instead of replaying its captured CUDA graph "
                         "(same device time within noise once the launches carry the PDL attribute, but exposed to host jitter)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    from layoutdit_b200 import DiTBackbone, _lib, config as cfgmod
    from layoutdit_b200.config import flops_per_image
    from layoutdit_b200.synth import make_state_dict, synthetic_pages

    assert torch.cuda.is_available(), "bench.py --impl ours needs a CUDA device"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    fac, B, H, W = WORKLOADS[args.workload]
    cfg = getattr(cfgmod, fac)()
    lib = _lib.load()
    peaks = measured_peaks()

    if args.head == "fpn":
        run_fpn_line(args, cfg, fac, B, H, W, dev, rank, world, peaks)
        return

    # random-init weights of the named architecture (HF init, seed 0), synthetic pages
    model = DiTBackbone(pretrained=False, config=cfg, state_dict=make_state_dict(cfg, 0, False),
                        use_cuda_graph=args.graph).to(dev).eval()
    eng = model._get_engine()
    pages = synthetic_pages(B, H, W, 1234 + rank)
    if args.graph:
        x_dev = eng.graph_input_buffer(B, H, W, torch.float32)  # the captured graph's own input tensor
        x_dev.copy_(pages)
    else:
        x_dev = pages.to(dev)                              # fp32, resident in HBM before the timed region
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    gathered = None
    def step_device():
        feats = model(x_dev)
        if world > 1:
            dist.all_gather_into_tensor(gathered, p5_flat(feats))
        return feats

    def p5_flat(feats):
        return feats["p5"].permute(0, 2, 3, 1).reshape(-1)   # channels-last memory of the static output buffer

    feats = model(x_dev)
    if world > 1:
        gathered = torch.empty(world * p5_flat(feats).numel(), dtype=torch.bfloat16, device=dev)
    launches_per_step = eng.last_launches(B, H, W)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ------------------------------------------------------------ device-resident timing
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for s, e in ev:
        flush.zero_()                                     # evict L2 between timed iterations (outside the events)
        s.record()
        step_device()
        e.record()
    barrier()
    per_step = [s.elapsed_time(e) for s, e in ev]
    total_ms = torch.tensor([sum(per_step)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)   # slowest rank defines the job time
    total_ms = float(total_ms.item())
    ms_per_step = total_ms / args.steps
    value = world * B * args.steps / (total_ms / 1e3)

    # ------------------------------------------------------------ end-to-end (host buffers)
    host_in = [pages.to(torch.float16).pin_memory() for _ in range(2)]
    feats0 = model(x_dev)
    host_out = [torch.empty(p5_flat(feats0).shape, dtype=torch.bfloat16).pin_memory() for _ in range(2)]
    e2e_model = DiTBackbone(pretrained=False, config=cfg, state_dict=None, use_cuda_graph=args.graph).to(dev).eval()
    e2e_model.dit.load_state_dict(model.dit.state_dict())


    def step_e2e(i):
        # the call a user makes for host-resident pages: H2D of this step's pages, forward, D2H of its result
        # (forward_host keeps the PCIe copies on their own streams, under the neighbouring steps' kernels)
        f, done = e2e_model.forward_host(host_in[i & 1], host_out[i & 1], "p5")
        if world > 1:
            dist.all_gather_into_tensor(gathered, p5_flat(f))
        return done

    for i in range(args.warmup):
        step_e2e(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        done = step_e2e(i)
    torch.cuda.current_stream(dev).wait_event(done)      # the last step's result has reached the host buffer
    e1.record()
    barrier()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_ms.item())
    e2e_value = world * B * args.steps / (e2e_ms / 1e3)
    clocks = sampler.stop() if rank == 0 else None
    h2d = host_in[0].numel() * host_in[0].element_size()
    d2h = host_out[0].numel() * host_out[0].element_size()

    # ------------------------------------------------------------ roofline of the dominant kernel
    # MLP up-projection: [M, D] x [I, D]^T + bias -> erf-GELU, one launch per layer; timed with
    # CUDA events around each launch inside eager forwards (so caches/clocks are those of a real step).
    D, I = cfg.hidden_size, cfg.intermediate_size
    geo = eng._geometry(B, H, W)
    outs = eng._alloc_outputs(geo)
    stream = torch.cuda.current_stream(dev)
    plan = eng._plan(geo, x_dev, outs, stream.cuda_stream)
    k_ev = []
    for _ in range(3):
        for name, fn, fargs in plan:
            if name in ("ldit_gemm_bias_gelu", "ldit_mlp_fused"):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream); _lib.check(fn(*fargs), name); b.record(stream)
                k_ev.append((a, b))
            else:
                _lib.check(fn(*fargs), name)
    torch.cuda.synchronize(dev)
    k_ms = statistics.mean(sorted(a.elapsed_time(b) for a, b in k_ev)[: max(1, len(k_ev) * 3 // 4)])
    fused_mlp = any(name == "ldit_mlp_fused" for name, _, _ in plan)
    k_flops = (4.0 if fused_mlp else 2.0) * geo.M * I * D
    achieved = k_flops / (k_ms / 1e3) / 1e12
    peak = peaks["bf16_tflops_sustained"]                 # kernel timed inside a long step -> sustained figure
    fl_img = flops_per_image(cfg, H, W)
    model_tflops = fl_img * value / 1e12
    # DRAM traffic of that kernel from the committed ncu --set full capture (same shape only: base224's fc1)
    traffic = profiled_traffic(r"gemm_tcgen05_kernel<\d+, 1, 2>") if args.workload == "base224" else None

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, ms, steps, cores, sample = cpu_reference_run(args.workload, 3, 1)
            cpu = {"value": round(v, 3), "unit": "images/s", "cores": cores, "kind": "port", "sample": sample}
        out = {
            "metric": "DiT backbone forward throughput", "value": round(value, 1), "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {fac} backbone forward, batch {B} per GPU, {H}x{W}, "
                                   f"random-init (HF init, seed 0) weights, 4 taps written",
                       "global_batch": world * B, "l2": "256 MiB buffer written between timed steps (outside the events)",
                       "parallelism": f"dp{world}", "gather": "p5 taps all-gathered over NCCL each step" if world > 1 else "none",
                       "timing": ("CUDA-graph replay" if args.graph else "stream launches with programmatic dependent launch (PDL)")
                                 + "; per-step CUDA events summed; max over ranks"},
            "model_tflops": round(model_tflops, 1),
            "model_frac_of_peak": round(model_tflops / (world * peaks["bf16_tflops"]), 4),
            "flops_per_image": fl_img,
            "roofline": {"bound": "tensor", "kernel": (f"mlp_tcgen05_kernel (fc1 + fc2 fused) M={geo.M} D={D} I={I}" if fused_mlp
                                    else f"gemm_tcgen05_kernel<EPI_BIAS_GELU> M={geo.M} N={I} K={D}"),
                         "achieved": round(achieved, 1), "peak": peak, "unit": "TFLOP/s", "frac": round(achieved / peak, 4),
                         "frac_of_burst_peak": round(achieved / peaks["bf16_tflops"], 4), "peak_source": peaks["source"],
                         "kernel_ms": round(k_ms, 4), "traffic": None if traffic is None else round(traffic["bytes"]),
                         "traffic_unit": "bytes per launch (dram read + write, ncu --set full)",
                         "traffic_source": None if traffic is None else traffic["source"],
                         "algorithmic_bytes": 2 * (geo.M * D + I * D + geo.M * I)},
            "e2e": {"value": round(e2e_value, 1), "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": round(e2e_ms / args.steps, 4), "input_dtype": "float16 pinned host",
                    "result": "p5 tap (bf16) copied to pinned host"},
            "gpu_launches": launches_per_step * args.steps,
            "launches_per_step": launches_per_step,
            "clocks": clocks,
        }
        if cpu is not None:
            out["cpu_baseline"] = cpu
        print(json.dumps(out), file=_RESULT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    import faulthandler
    faulthandler.enable()          # a fatal signal in any rank leaves a Python stack in stderr instead of nothing
    try:
        main()
    except BaseException:
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        raise
