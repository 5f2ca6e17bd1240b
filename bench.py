#!/usr/bin/env python
"""bench.py -- sampled motion frames/sec, 50-step DDIM (eta=0, respace '15,15,8,6,6'), (B x T x 322).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload t2m|s2g|m2d] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's CPU arithmetic (oracle port) on the host cores

One "step" = ONE complete 50-step DDIM sampling run producing B x T frames per GPU (BASELINE.json
configs[1]: t2m, B=256, T=196, D=322; weak scaling: every rank samples its own 256 rows of the global
batch, one all-gather of the results at the end, configs[4]).  Prints ONE JSON line (rank 0).

  value : whole-job frames/s with x_T already resident in HBM, device-timed (CUDA events, max over ranks)
  e2e   : the same through the host-buffer C-ABI call (mcm_sample_host): pinned x_T H2D + loop + x_0 D2H
  roofline     : tcgen05 GEMM kernel class -- algorithmic FLOPs / summed kernel time (CUDA events around every
                 launch, separate instrumented pass) against the measured bf16 peak
  cpu_baseline : the oracle (bit-identical restatement of the reference's CPU path) on a bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (B per GPU, T, control blocks, control feats, control length)   [BASELINE.json configs 1..3]
    "t2m": dict(B=256, T=196, n_ctrl=0, c_feats=0, c_len=0),
    "s2g": dict(B=128, T=300, n_ctrl=2, c_feats=2048, c_len=297),
    "m2d": dict(B=64, T=1024, n_ctrl=4, c_feats=35, c_len=1024),
}
RESPACE = "15,15,8,6,6"
# ncu (--set full, single stream, B=256 x T=196): dram__bytes_read.sum + dram__bytes_write.sum of one fused_block_kernel launch
FUSED_DRAM_BYTES_PER_LAUNCH = 441.4e6    # profiles/r02_fused_kernels_metrics.csv: 218.33 MB read + 223.05 MB written
N_STEPS = 50


def algorithmic_flops(T, n_ctrl, n_tokens=77, D=512, E=2048, F=1024, H=4, IN=322, Lt=256, L=8, c_feats=0, c_len=0):
    """live graph only, per sample; returns (per denoise step, one-off per run)."""
    layers = L + n_ctrl
    sa = 8 * D * T * T + 4 * D * T * T / H + 4 * E * T
    ffn = 4 * T * D * F + 4 * E * D + 2 * T * D * D
    ca = 4 * T * D * D + 2 * T * D * D / H + 4 * E * D
    per_step = layers * (sa + ffn + ca) + 4 * T * IN * D + 2 * D * E + 2 * E * E
    if n_ctrl:
        per_step += (n_ctrl + 1) * 2 * T * D * D
    once = layers * (4 * n_tokens * Lt * D + 2 * n_tokens * D * D / H) + 2 * c_len * c_feats * D
    return per_step, once


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(tflops=float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1421.6))), hbm=float(d["hbm_gbs"]),
                    source="measured (MEASURED_PEAKS.json, bf16 sustained)")
    return dict(tflops=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


_CPU_SETUP = {}


def oracle_cpu_setup(T, n_ctrl, c_feats, c_len, B_cpu):
    """Synthetic weights / inputs for the CPU arm (built once, outside any timed region)."""
    key = (T, n_ctrl, c_feats, c_len, B_cpu)
    if key in _CPU_SETUP:
        return _CPU_SETUP[key]
    import torch
    from motioncraft_b200 import modules, synth
    from oracle import mcm_oracle as O
    # the reference executes the dead ffn_channel branch of every layer (mcm.py:33-34, result discarded; 12.14 vs 7.6 GFLOP per
    # sample-step at T = 196): the timed CPU arm does the same work, the parity oracle (tests/) skips it
    O.EXECUTE_DEAD_FFN_CHANNEL = True
    shapes = modules.ctrl_state_shapes(T, n_ctrl, c_feats) if n_ctrl else modules.state_shapes(seq_len=T)
    sd = synth.synth_state_dict(shapes)
    x = synth.synth_rows("x_T", (T, 322), synth.SEED_XT, 0, B_cpu)
    xf_out = synth.synth_rows("xf_out", (77, 256), synth.SEED_XF_OUT, 0, B_cpu)
    xf_proj = synth.synth_rows("xf_proj", (2048,), synth.SEED_XF_PROJ, 0, B_cpu)
    c = synth.synth_rows("c", (c_len, c_feats), synth.SEED_C_EMB, 0, B_cpu) if n_ctrl else None
    if n_ctrl:
        fn = lambda xx, tt: O.control_forward(sd, xx, tt, xf_proj, xf_out, c)  # noqa: E731
    else:
        fn = lambda xx, tt: O.mcm_forward(sd, xx, tt, xf_proj, xf_out)  # noqa: E731
    with torch.no_grad():
        fn(x, torch.full((B_cpu,), 999, dtype=torch.long))          # warm-up (thread pool, allocator)
    _CPU_SETUP[key] = (fn, x)
    return _CPU_SETUP[key]


def oracle_cpu_rate(T, n_ctrl, c_feats, c_len, B_cpu, n_denoise):
    """frames/s of the reference's CPU arithmetic (oracle port) on a bounded sample: B_cpu samples,
    n_denoise of the 50 denoise steps + sampler updates, extrapolated to the 50-step run.
    This is the ONLY place outside tests/ and smoke() where oracle/ is executed, and only as the thing
    measured in the baseline / reference arm -- never on the product path."""
    import torch
    from oracle import mcm_oracle as O
    fn, x = oracle_cpu_setup(T, n_ctrl, c_feats, c_len, B_cpu)
    tables, tmap = O.spaced_tables(1000, RESPACE)
    sub_tables = {k: v[-n_denoise:] for k, v in tables.items()}
    with torch.no_grad():
        t0 = time.perf_counter()
        O.ddim_sample_loop(fn, x, sub_tables, tmap[-n_denoise:])
        dt = time.perf_counter() - t0
    return B_cpu * T / (N_STEPS * dt / n_denoise), dt, torch.get_num_threads()


def run_reference_arm(args, wl, rank, world):
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    B_cpu = args.cpu_batch or 64
    oracle_cpu_setup(wl["T"], wl["n_ctrl"], wl["c_feats"], wl["c_len"], B_cpu)
    for _ in range(args.warmup):
        oracle_cpu_rate(wl["T"], wl["n_ctrl"], wl["c_feats"], wl["c_len"], B_cpu, 1)
    vals, wall, cores = [], 0.0, None
    for _ in range(args.steps):
        v, dt, cores = oracle_cpu_rate(wl["T"], wl["n_ctrl"], wl["c_feats"], wl["c_len"], B_cpu, 10)
        vals.append(v)
        wall += dt
    value = sum(vals) / len(vals)
    sample = (f"reference CPU arithmetic (oracle port, bit-identical to mogen's PyTorch path on CPU, INCLUDING the dead "
              f"ffn_channel branch the reference executes and discards): B={B_cpu} samples x "
              f"10 of 50 DDIM steps per bench step, fp32, {cores} threads, frames/s = B*T/(50*t_denoise_step)")
    line = {"impl": "reference", "metric": "sampled motion frames/sec (50-step DDIM)", "value": value, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, wl, world),
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, wl, world, name=None):
    return {"workload": f"{name or args.workload}: synthetic (B={wl['B']}/GPU x T={wl['T']} x 322), 50-step DDIM eta=0 respace "
                        f"'{RESPACE}', MCMTransformer 8 layers" + (f" + {wl['n_ctrl']} control blocks" if wl["n_ctrl"] else ""),
            "global_batch": wl["B"] * world, "seq_len": wl["T"], "parallelism": f"dp{world} (batch shards, one all-gather)",
            "l2": "per-step working set (>1 GB activations + 124 MB weights) exceeds the 126 MB L2; no explicit flush"}


def build_engine_and_inputs(wl, rank, world, dev, precise=False):
    import torch
    from motioncraft_b200 import dist as mdist, modules, synth
    from motioncraft_b200.engine import DenoiserEngine
    B, T = wl["B"], wl["T"]
    n_total = B * world
    lo, hi = mdist.shard_range(n_total, rank, world)
    # synthetic weights (replicated) and this rank's rows of the globally seeded inputs
    if wl["n_ctrl"]:
        sd = modules.engine_state_from_ctrl(synth.synth_state_dict(modules.ctrl_state_shapes(T, wl["n_ctrl"], wl["c_feats"])))
    else:
        sd = {k: v for k, v in synth.synth_state_dict(modules.state_shapes(seq_len=T)).items() if ".ffn_channel." not in k}
    x_T = synth.synth_rows("x_T", (T, 322), synth.SEED_XT, lo, hi)
    xf_out = synth.synth_rows("xf_out", (77, 256), synth.SEED_XF_OUT, lo, hi)
    xf_proj = synth.synth_rows("xf_proj", (2048,), synth.SEED_XF_PROJ, lo, hi)
    c = synth.synth_rows("c", (wl["c_len"], wl["c_feats"]), synth.SEED_C_EMB, lo, hi) if wl["n_ctrl"] else None
    eng = DenoiserEngine(sd, seq_len=T, max_batch=B, num_ctrl_blocks=wl["n_ctrl"], ctrl_cond_feats=wl["c_feats"],
                         precise_all=precise, device=dev)
    del sd
    return eng, x_T, xf_out, xf_proj, c, n_total


def measure_workload(args, name, wl, rank, local_rank, world, dev, steps, warmup, with_clocks):
    """Times `steps` complete sampling runs of one workload; returns the numbers of its JSON block (rank 0) or None."""
    import torch
    import torch.distributed as dist
    from motioncraft_b200 import _lib
    from motioncraft_b200 import dist as mdist
    from motioncraft_b200.diffusion import build_diffusion
    from motioncraft_b200.engine import SamplerTables

    B, T = wl["B"], wl["T"]
    eng, x_T, xf_out, xf_proj, c, n_total = build_engine_and_inputs(wl, rank, world, dev, args.precise)
    d = build_diffusion(dict(beta_scheduler="linear", diffusion_steps=1000, model_mean_type="epsilon",
                             model_var_type="fixed_small", respace=RESPACE))
    tables = SamplerTables(d._tables(), d.timestep_map, "ddim", 0.0)
    x_dev = x_T.to(dev)
    # host side of the end-to-end leg: EVERYTHING a caller hands over lives in pinned host memory
    x_pin = x_T.pin_memory()
    out_pin = torch.empty_like(x_pin).pin_memory()
    cond_pin = [t.pin_memory() if t is not None else None for t in (xf_out, xf_proj, c)]
    cond = tuple(t.to(dev) if t is not None else None for t in (xf_out, xf_proj, c))
    full_pin = torch.empty((n_total, T, 322), dtype=torch.float32).pin_memory() if (world > 1 and rank == 0) else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def one_run_device():
        # the step-invariant condition work is part of every sampling run (counted in the algorithmic FLOPs)
        eng.prepare_conditions(*cond)
        x0 = eng.sample(tables, x_dev)
        return mdist.gather_rows(x0, n_total) if world > 1 else x0

    def one_run_host():
        # H2D of the conditions, then the host-buffer C-ABI call (H2D x_T + loop + D2H x_0)
        cd = [t.to(dev, non_blocking=True) if t is not None else None for t in cond_pin]
        eng.prepare_conditions(*cd)
        if world == 1:
            return eng.sample_host(tables, x_pin, out_pin)
        # N > 1: the result every caller receives is the GATHERED batch: H2D, loop, one NCCL all-gather, D2H on rank 0
        x0 = eng.sample(tables, x_pin.to(dev, non_blocking=True))
        full = mdist.gather_rows(x0, n_total)
        if rank == 0:
            full_pin.copy_(full, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return full_pin

    for _ in range(warmup):
        one_run_device()
    barrier()
    clocks = ClockSampler(local_rank) if with_clocks else None
    if clocks is not None and rank == 0:
        clocks.start()
    launches0 = _lib.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(steps):
        one_run_device()
    ev1.record()
    barrier()
    ms_dev = ev0.elapsed_time(ev1)
    launches = _lib.kernel_launches() - launches0
    clk = clocks.stop() if (clocks is not None and rank == 0) else None

    # ---- end-to-end through the host-buffer C-ABI call (H2D + loop + D2H inside the timed region)
    one_run_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_run_host()
    torch.cuda.synchronize(dev)
    ms_e2e = (time.perf_counter() - t0) * 1e3

    # ---- roofline leg: separate instrumented run (events around every launch), not part of the numbers above
    # (single stream, eager launches: with the two-stream schedule of the timed region kernels of different streams
    # share the SMs, so per-kernel event times would no longer be times of a kernel running alone)
    dual_default = os.environ.get("MCM_DUAL", "1") != "0"
    eng.set_option("dual", 0)
    _lib.timing_enable(True)
    one_run_device()
    tm = _lib.timing_collect()
    _lib.timing_enable(False)
    eng.set_option("dual", 1 if dual_default else 0)
    eng.close()

    if world > 1:
        t = torch.tensor([ms_dev, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        return None

    frames = n_total * T * steps
    value = frames / (ms_dev * 1e-3)
    e2e = frames / (ms_e2e * 1e-3)
    per_step, once = algorithmic_flops(T, wl["n_ctrl"], c_feats=wl["c_feats"], c_len=wl["c_len"])
    flops_run = B * (N_STEPS * per_step + once)            # per GPU per sampling run
    peaks = measured_peaks()
    gemm_ms = tm["gemm"]["ms"] + tm["fused"]["ms"]          # every tcgen05 kernel of the run
    all_tc = flops_run / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
    whole = flops_run / (ms_dev / steps * 1e-3) / 1e12
    fz = tm["fused"]
    if fz["launches"] > 0:
        # dominant kernel: the fused cross-attention + FFN token kernel (one launch per decoder layer and step).
        # achieved = its ALGORITHMIC flops (SURVEY.md 8a rows a9 + a10 without the AdaLN emb GEMM:
        # 2 * rows * (2*512*512 + 512*128 + 2*512*1024 + 512*512) per launch) / its CUDA-event device time
        kernel = ("fused_block_kernel (cross-attention + FFN of one decoder layer, tcgen05 cta_group::2; all launches of "
                  "one sampling run, timed single-stream with CUDA events around every launch)")
        achieved = fz["flops"] / (fz["ms"] * 1e-3) / 1e12
        dom_ms, dom_n = fz["ms"], fz["launches"]
        # dram__bytes_read + dram__bytes_write per launch, ncu --set full (profiles/); algorithmic minimum is one read +
        # one write of h = 2 * B*T*512*4 bytes
        traffic = FUSED_DRAM_BYTES_PER_LAUNCH if (name == "t2m" and B == 256) else None
        algo_bytes = 2.0 * B * T * 512 * 4
    else:
        kernel = "gemm_tc_kernel (tcgen05, all launches of one sampling run; timed single-stream, eager)"
        achieved, dom_ms, dom_n, traffic, algo_bytes = all_tc, tm["gemm"]["ms"], tm["gemm"]["launches"], None, None
    h2d = sum(int(t.numel() * 4) for t in [x_pin] + [t for t in cond_pin if t is not None])
    d2h = int((full_pin if full_pin is not None else out_pin).numel() * 4)
    return {
        "value": value, "ms_per_step": ms_dev / steps, "steps": steps, "warmup": warmup,
        "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": ("mcm_prepare_conditions + mcm_sample_host: pinned host x_T / xf_out / xf_proj / c -> x_0 in pinned host memory"
                        if world == 1 else
                        "pinned host inputs -> H2D -> mcm_sample -> ONE NCCL all-gather -> D2H of the gathered batch on rank 0")},
        "gpu_launches": int(launches), "clocks": clk,
        "roofline": {"bound": "tensor", "kernel": kernel,
                     "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                     "frac": (achieved / peaks["tflops"]) if achieved else None,
                     "traffic": traffic, "traffic_unit": "DRAM bytes per launch (ncu)", "algorithmic_bytes_per_launch": algo_bytes,
                     "peak_source": peaks["source"], "algorithmic_gflop_per_frame": flops_run / (B * T) / 1e9,
                     "kernel_ms_per_run": dom_ms, "kernel_launches_per_run": dom_n,
                     "kernel_share_of_device_time": dom_ms / (gemm_ms + tm["row"]["ms"]) if gemm_ms > 0 else None,
                     "all_tcgen05_kernels_ms_per_run": gemm_ms, "all_tcgen05_kernels_achieved": all_tc,
                     "other_tcgen05_launches_per_run": tm["gemm"]["launches"],
                     "row_kernel_ms_per_run": tm["row"]["ms"], "row_kernel_launches_per_run": tm["row"]["launches"],
                     "whole_step_achieved": whole, "whole_step_frac": whole / peaks["tflops"]},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="t2m", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override per-GPU batch")
    ap.add_argument("--cpu-batch", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the s2g / m2d extra_workloads block (N = 1 only)")
    ap.add_argument("--precise", action="store_true", help="debug: every GEMM in bf16x2-split mode")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.batch:
        wl["B"] = args.batch
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, wl, rank, world)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: motioncraft_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    main_res = measure_workload(args, args.workload, wl, rank, local_rank, world, dev, args.steps, args.warmup, True)
    extra = {}
    if world == 1 and not args.no_extra and args.workload == "t2m" and not args.batch:
        # BASELINE.json configs 2 and 3 on the same box, 2 timed runs each after 3 warm-up runs, so that the driver's
        # record carries their throughput and roofline fractions too (they are parity-test cases, not the headline)
        for name in ("s2g", "m2d"):
            r = measure_workload(args, name, dict(WORKLOADS[name]), rank, local_rank, world, dev, 2, 3, False)
            extra[name] = {"value": r["value"], "unit": "frames/s", "ms_per_step": r["ms_per_step"], "steps": 2, "warmup": 3,
                           "config": workload_config(args, dict(WORKLOADS[name]), world, name),
                           "e2e": r["e2e"], "gpu_launches": r["gpu_launches"],
                           "roofline": {k: r["roofline"][k] for k in ("kernel", "achieved", "peak", "unit", "frac",
                                                                      "kernel_share_of_device_time",
                                                                      "all_tcgen05_kernels_achieved", "whole_step_achieved",
                                                                      "whole_step_frac", "algorithmic_gflop_per_frame")}}

    if rank == 0:
        B, T = wl["B"], wl["T"]
        line = {
            "metric": "sampled motion frames/sec (50-step DDIM)", "value": main_res["value"], "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16 operands (bf16x2-split for embed/out/AdaLN), f32 accumulate/residual",
            "data": "synthetic", "config": workload_config(args, wl, world),
            "e2e": main_res["e2e"], "gpu_launches": main_res["gpu_launches"], "clocks": main_res["clocks"],
            "roofline": main_res["roofline"],
        }
        if extra:
            line["extra_workloads"] = extra
        if not args.no_cpu_baseline:
            B_cpu = args.cpu_batch or 64
            torch.set_num_threads(os.cpu_count() or 1)
            v, dt, cores = oracle_cpu_rate(T, wl["n_ctrl"], wl["c_feats"], wl["c_len"], B_cpu, 25)
            line["cpu_baseline"] = {"value": v, "unit": "frames/s", "cores": cores, "kind": "port",
                                    "sample": f"oracle (bit-identical restatement of mogen's CPU path, dead ffn_channel branch "
                                              f"executed as the reference does): B={B_cpu} x 25 of 50 DDIM steps, fp32, "
                                              f"{dt:.1f} s of CPU work, extrapolated B*T/(50*t_step)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
