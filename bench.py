#!/usr/bin/env python
"""bench.py -- the headline benchmark of the estimator hot path (BASELINE.json config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (per GPU, fixed as N grows -> weak scaling): 273 PRB / 30 kHz full-band comb-4 pilots,
4 ports (2 CDM groups x OCC-2), 4 rx antennas x 64 UEs = 256 units per slot, `--slots` slots per
step (default 64 -> 16 384 units per step), one pilot symbol, 14-symbol output grid, raised-cosine
smoothing, CFO compensation switched on.  One step = one launch of the fused kernel over the
step's units (+ the NCCL gather of the scalar table when N > 1).  Synthetic data, seeded.

`value`     channel estimates/s (units x 4 ports), device-timed, inputs resident in HBM.
`e2e`       the same metric through the host-buffer entry point (pinned host tensors in, dense
            grids + scalars back in host memory, copies inside the timed region).  The entry point
            moves the per-hop columns over PCIe and replicates them over the symbols on host threads
            (srsce_replicate_host: a copy, bit-identical grid); `e2e.dense_copy` is the same call
            with the device-written dense grid crossing the link instead.
`roofline`  algorithmic HBM bytes of the fused kernel / its CUDA-event duration vs the measured
            copy bandwidth in MEASURED_PEAKS.json.
`cpu_baseline` the numpy oracle port on all host cores (rank 0, N = 1), bounded sample.
`--impl reference` times that CPU path alone, same metric / config.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "SRS channel estimates/sec (UE x port x rx-ant) at 273 PRB"
UNIT = "estimates/s"
N_PRB, N_SYM, LAYERS, RX_ANT, UES, PILOT_SYMBOLS = 273, 14, 4, 4, 64, (13,)
UNITS_PER_SLOT = RX_ANT * UES
WORKLOAD = "config2"
FILL = "linear"          # "cnn": the ce_dl_cnn fill + blend (BASELINE configs[3])


def select_workload(name, slots):
    """config2 (headline): 4 rx x 64 UEs per slot.  config3: massive-MIMO, 64 rx x 128 UEs per slot
    (BASELINE.json configs[2]; its 1024 slots are streamed `slots` at a time through one reused output buffer)."""
    global RX_ANT, UES, UNITS_PER_SLOT, WORKLOAD, FILL
    WORKLOAD = name
    if name == "config3":
        RX_ANT, UES = 64, 128
    if name == "config4":                      # ce_dl_cnn fill (srsce_fill) on the config-2 geometry, 4096 UE-antenna grids per step
        FILL = "cnn"
    UNITS_PER_SLOT = RX_ANT * UES
    return slots if slots else {"config2": 64, "config3": 2, "config4": 16}[name]


def workload_config(slots, e2e_slots=None):
    variant = "ce_dl_cnn (inpainting fill + blend)" if FILL == "cnn" else "ce_rule_tensorized"
    cfg = {"workload": f"{WORKLOAD}: {variant} batched, 273 PRB 30 kHz, 4 ports, {RX_ANT} rx antennas, comb-4, "
                       f"{UES} UEs per slot, 1 pilot symbol, 14-symbol grid, filter smoothing, CFO compensation on",
           "units_per_step_per_gpu": slots * UNITS_PER_SLOT, "slots_per_step_per_gpu": slots, "estimates_per_unit": LAYERS,
           "rx_layout": "reference ([n_sc, n_sym] row major per unit, one grid per unit)",
           "l2_policy": "inputs+outputs per step (%.1f GB) exceed the 126 MB L2; no flush needed"
                        % (slots * UNITS_PER_SLOT * (algorithmic_bytes_per_unit() + 12 * N_PRB * N_SYM * 8) / 1e9)}
    if e2e_slots is not None:
        cfg["e2e_units_per_step"] = min(e2e_slots * UNITS_PER_SLOT, slots * UNITS_PER_SLOT, 2048)
    return cfg


def geometry():
    from srsran_ce_pytorch_b200 import synth
    hop1 = synth.make_hop(N_PRB, 0, N_PRB, list(PILOT_SYMBOLS), "comb4", 2)
    return hop1, synth.no_hop(), synth.make_config(30e3, "filter", True)


def algorithmic_bytes_per_unit():
    """8 * [G*N*S (rx pilot REs) + N*S*L/A (pilots shared by the A antennas) + n_sc*n_sym*L] + 48."""
    n, s, g = N_PRB * 3, len(PILOT_SYMBOLS), 2
    return 8 * (g * n * s + n * s * LAYERS // RX_ANT + 12 * N_PRB * N_SYM * LAYERS) + 48


# --------------------------------------------------------------------------------------------
# CPU arm (oracle port on all host cores)
# --------------------------------------------------------------------------------------------
def cpu_pool(n_distinct=None):
    from oracle import ce_oracle, cpu_bench
    from srsran_ce_pytorch_b200 import synth
    hop1, hop2, cfg = geometry()
    n_distinct = n_distinct or max(16, RX_ANT)
    d = synth.make_units(hop1, hop2, cfg, LAYERS, n_distinct, pilot_share=RX_ANT, seed=4321)
    return cpu_bench.OraclePool(d["rx"].numpy(), d["pilots"].numpy(), RX_ANT, 1.0, ce_oracle.hop_from_any(hop1),
                                ce_oracle.HopSpec.absent(), ce_oracle.est_from_any(cfg),
                                fill="cnn" if FILL == "cnn" else "tensorized")


def cpu_baseline(units_per_worker=480):
    if FILL == "cnn":
        units_per_worker = 48                      # the port of ce_dl_cnn's inpainting loop takes ~0.2 s per unit
    pool = cpu_pool()
    try:
        pool.run(4)
        wall = pool.run(units_per_worker)
    finally:
        pool.close()
    units = pool.cores * units_per_worker
    return {"value": LAYERS * units / wall, "unit": UNIT, "cores": pool.cores, "kind": "port",
            "sample": f"{units} units of the bench workload ({units_per_worker} per worker, 1 thread each, "
                      f"{wall:.2f} s wall); numpy port of " +
                      ("ce_dl_cnn" if FILL == "cnn" else "ce_rule_tensorized (2.9x faster than the torch reference in the "
                                                         "build container)")}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pool = cpu_pool()
    upw = 8 if FILL == "cnn" else 96              # units per worker and step: 1536 units per step on 16 cores, ~0.2 s
    try:
        for _ in range(max(1, args.warmup)):
            pool.run(4)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.run(upw)
        wall = time.perf_counter() - t0
    finally:
        pool.close()
    units = pool.cores * upw * args.steps
    value = LAYERS * units / wall
    cfg = workload_config(args.slots)
    cfg["reference_step"] = f"{pool.cores * upw} units per step (bounded sample of the step's units)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "complex64 (f32) data, f64 scalars", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": pool.cores, "kind": "port",
                         "sample": f"{units} units, {upw} per worker per step, 1 thread per worker"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock + throttle reasons of one GPU every ~5 ms through NVML while the timed region runs."""
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.max_mhz, self.stop_flag, self.err, self.ready = index, [], None, False, None, False

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                nv.nvmlDeviceGetCurrentClocksThrottleReasons
            self.ready = True
            while not self.stop_flag:
                self.samples.append((time.perf_counter(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), int(reasons(h))))
                time.sleep(0.002)
        except Exception as e:  # noqa: BLE001 -- the bench must not die on a monitoring problem
            self.err = repr(e)
            self.ready = True

    def summary(self, t0, t1):
        """Median SM clock and the union of throttle reasons over the samples taken in [t0, t1]
        (falls back to the nearest samples when the window is shorter than the sampling period)."""
        win = [x for x in self.samples if t0 <= x[0] <= t1]
        if not win and self.samples:
            win = sorted(self.samples, key=lambda x: abs(x[0] - 0.5 * (t0 + t1)))[:3]
        clocks = sorted(x[1] for x in win)
        mask = 0
        for x in win:
            mask |= x[2]
        return {"sm_mhz": clocks[len(clocks) // 2] if clocks else None, "sm_max_mhz": self.max_mhz,
                "reasons": [n for b, n in self.REASONS.items() if mask & b and n != "gpu_idle"],
                "samples": len(win), **({"error": self.err} if self.err else {})}


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except (ValueError, IndexError):
            return local
    return local


def make_step_inputs(slots, seed0, device):
    import torch
    from srsran_ce_pytorch_b200 import synth
    hop1, hop2, cfg = geometry()
    n = slots * UNITS_PER_SLOT
    rx = torch.empty((n, 12 * N_PRB, N_SYM), dtype=torch.complex64, device=device)
    pil = torch.empty((n // RX_ANT, N_PRB * 3, len(PILOT_SYMBOLS), LAYERS), dtype=torch.complex64, device=device)
    for s in range(slots):
        d = synth.make_units(hop1, hop2, cfg, LAYERS, UNITS_PER_SLOT, pilot_share=RX_ANT, seed=seed0 + s, device=device)
        rx[s * UNITS_PER_SLOT:(s + 1) * UNITS_PER_SLOT] = d["rx"]
        pil[s * UES:(s + 1) * UES] = d["pilots"]
    return rx, pil


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")

    # Libraries (NCCL's version banner, for one) write to fd 1: keep the process's stdout for the ONE JSON line
    # and send everything else to stderr.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline()                     # before CUDA is touched: workers are spawned, not forked

    from srsran_ce_pytorch_b200 import _lib, estimator, sharding
    from srsran_ce_pytorch_b200.hostio import HostPipeline

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    hop1, hop2, cfg = geometry()
    plan = estimator.make_plan(N_PRB, N_SYM, LAYERS, hop1, hop2, cfg, FILL)
    n_local = args.slots * UNITS_PER_SLOT
    n_global = n_local * world
    rx, pil = make_step_inputs(args.slots, 1234 + 1000 * rank, dev)
    out = torch.empty((n_local, 12 * N_PRB, N_SYM, LAYERS), dtype=torch.complex64, device=dev)
    sc = torch.empty((n_local, 6), dtype=torch.float64, device=dev)

    launches = 0

    def step():
        nonlocal launches
        plan.estimate(rx, pil, 1.0, RX_ANT, out=out, scalars=sc)
        launches += _lib.lib.srsce_last_launch_count()
        if world > 1:
            return sharding.gather_scalars(sc, n_global, RX_ANT)
        return sc

    sampler = ClockSampler(physical_gpu_index(local))
    sampler.start()
    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    for _ in range(200):                       # NVML is up before the timed region starts
        if sampler.ready:
            break
        time.sleep(0.01)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches = 0
    wall0 = time.perf_counter()
    k0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    k1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for k in range(args.steps):
        k0[k].record()
        plan.estimate(rx, pil, 1.0, RX_ANT, out=out, scalars=sc)
        k1[k].record()
        launches += _lib.lib.srsce_last_launch_count()
        if world > 1:
            sharding.gather_scalars(sc, n_global, RX_ANT)
    t1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    wall1 = time.perf_counter()
    sampler.stop_flag = True
    sampler.join(timeout=2.0)

    total_ms = t0.elapsed_time(t1)
    kern_ms = sum(a.elapsed_time(b) for a, b in zip(k0, k1)) / args.steps
    if world > 1:
        t = torch.tensor([total_ms, kern_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, kern_ms = float(t[0]), float(t[1])
    assert bool(torch.isfinite(sc[:, :4]).all()), "non-finite scalars"

    # ---- the same step with symbol-major grids ([B, n_sym, n_sc] memory viewed as [B, n_sc, n_sym]), the
    #      layout OFDM front ends produce; the batched entry point takes any strides
    alt = None
    if not args.no_alt_layout:
        rx_sm = rx.transpose(1, 2).contiguous().transpose(1, 2)
        for _ in range(3):
            plan.estimate(rx_sm, pil, 1.0, RX_ANT, out=out, scalars=sc)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a0.record()
        for _ in range(args.steps):
            plan.estimate(rx_sm, pil, 1.0, RX_ANT, out=out, scalars=sc)
        a1.record()
        torch.cuda.synchronize()
        alt_ms = a0.elapsed_time(a1) / args.steps
        alt = {"rx_layout": "symbol-major", "value": LAYERS * n_local / (alt_ms * 1e-3), "unit": UNIT,
               "kernel_ms": alt_ms, "n_gpus": 1}
        del rx_sm

    # ---- one received grid per (slot, rx antenna), shared by the UEs multiplexed on it (what a caller of the reference
    #      does by passing the same tensor to 64 calls): srsce_estimate_indexed, no copies materialised
    shared = None
    if not args.no_alt_layout:
        grids = rx.view(args.slots, UES, RX_ANT, 12 * N_PRB, N_SYM)[:, 0].reshape(args.slots * RX_ANT, 12 * N_PRB, N_SYM).contiguous()
        unit = torch.arange(n_local, device=dev)
        rx_index = ((unit // UNITS_PER_SLOT) * RX_ANT + unit % RX_ANT).to(torch.int64)
        for _ in range(3):
            plan.estimate_indexed(grids, rx_index, pil, 1.0, RX_ANT, out=out, scalars=sc)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s0.record()
        for _ in range(args.steps):
            plan.estimate_indexed(grids, rx_index, pil, 1.0, RX_ANT, out=out, scalars=sc)
        s1.record()
        torch.cuda.synchronize()
        sms_ = s0.elapsed_time(s1) / args.steps
        shared = {"rx": "one grid per (slot, rx antenna) shared by the slot's UEs (srsce_estimate_indexed)",
                  "value": LAYERS * n_local / (sms_ * 1e-3), "unit": UNIT, "kernel_ms": sms_, "n_gpus": 1}
        del grids

    # ---- compact output (srsce_estimate_compact, SURVEY 8f-2: outside the reference contract, reported separately)
    compact = None
    if not args.no_alt_layout:
        for _ in range(3):
            plan.estimate_compact(rx, pil, 1.0, RX_ANT)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        c0.record()
        for _ in range(args.steps):
            plan.estimate_compact(rx, pil, 1.0, RX_ANT)
        c1.record()
        torch.cuda.synchronize()
        cms = c0.elapsed_time(c1) / args.steps
        compact = {"output": "[B, hops, n_sc, L] + [B, n_sym] factors instead of the dense grid",
                   "value": LAYERS * n_local / (cms * 1e-3), "unit": UNIT, "kernel_ms": cms, "n_gpus": 1}

    # ---- end to end through the host-buffer entry point
    e2e_slots = min(args.e2e_slots, args.slots)
    n_e2e = min(e2e_slots * UNITS_PER_SLOT, n_local, 2048)     # <= 3 GB of pinned output
    h_rx = rx[:n_e2e].cpu().pin_memory()
    h_pil = pil[:n_e2e // RX_ANT].cpu().pin_memory()
    h_out = torch.empty((n_e2e, 12 * N_PRB, N_SYM, LAYERS), dtype=torch.complex64).pin_memory()
    h_sc = torch.empty((n_e2e, 6), dtype=torch.float64).pin_memory()
    # The entry point's default ("auto"): this plan's grid has no final rotation, so the per-hop columns cross PCIe and
    # the library's host threads replicate them over the symbols into h_out (srsce_replicate_host); the dense
    # device->host copy of the same grid is timed beside it.  Host threads are split between the ranks of a box.
    host_threads = 0 if world == 1 else max(2, (os.cpu_count() or 16) // world)
    pipe = HostPipeline(plan, 256, RX_ANT, dev, n_streams=3, replicate="auto", host_threads=host_threads)
    pipe_dense = HostPipeline(plan, 64, RX_ANT, dev)      # 32 chunks per pass: the fill / drain bubble of a pass is ~3 % (link-bound otherwise)
    e2e_steps = max(2, min(args.steps, args.e2e_steps))
    chk_out, chk_sc = plan.estimate(rx[:64], pil[:64 // RX_ANT], 1.0, RX_ANT)   # (out / sc were reused by the passes above)
    chk_out, chk_sc = chk_out.cpu(), chk_sc.cpu()

    def timed_e2e(p):
        h_out.fill_(float("nan"))
        for _ in range(2):
            p.run(h_rx, h_pil, 1.0, h_out, h_sc)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        w0 = time.perf_counter()
        for _ in range(e2e_steps):
            p.run(h_rx, h_pil, 1.0, h_out, h_sc)          # returns after the whole grid has landed in host memory
        torch.cuda.synchronize()
        dt = time.perf_counter() - w0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
        same_grid = bool(torch.equal(torch.view_as_real(h_out[:64]), torch.view_as_real(chk_out))) and \
            not bool(torch.isnan(torch.view_as_real(h_out[-1])).any())
        same_sc = bool(torch.allclose(h_sc[:64], chk_sc, rtol=2e-6, atol=1e-12))
        return dt, same_grid, same_sc

    dense_s, dense_grid_ok, dense_sc_ok = timed_e2e(pipe_dense)
    e2e_s, e2e_grid_ok, e2e_sc_ok = timed_e2e(pipe)
    del chk_out, chk_sc
    if compact is not None:
        del h_out
        pipe_c = HostPipeline(plan, 256, RX_ANT, dev, compact=True)
        hc = (torch.empty((n_e2e, 1, 12 * N_PRB, LAYERS), dtype=torch.complex64).pin_memory(),
              torch.empty((n_e2e, N_SYM), dtype=torch.complex64).pin_memory())
        for _ in range(2):
            pipe_c.run(h_rx, h_pil, 1.0, hc, h_sc)
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        for _ in range(e2e_steps):
            pipe_c.run(h_rx, h_pil, 1.0, hc, h_sc)
        torch.cuda.synchronize()
        compact["e2e_value"] = LAYERS * n_e2e * e2e_steps / (time.perf_counter() - w0)
        compact["e2e_h2d_bytes_per_step"], compact["e2e_d2h_bytes_per_step"] = pipe_c.h2d_bytes, pipe_c.d2h_bytes

    if rank == 0:
        peaks_path = ROOT / "MEASURED_PEAKS.json"
        if peaks_path.exists():
            peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (copy, burst)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        bpu = algorithmic_bytes_per_unit()
        achieved = bpu * n_local / (kern_ms * 1e-3) / 1e9
        traffic = None
        prof = ROOT / "profiles" / "traffic.json"
        if prof.exists():
            traffic = json.loads(prof.read_text()).get("dram_bytes_per_launch_at_bench_size") if WORKLOAD == "config2" and \
                args.slots == 64 else None
        line = {
            "metric": METRIC, "value": LAYERS * n_global * args.steps / (total_ms * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "complex64 (f32) data, f64 scalars", "data": "synthetic",
            "config": workload_config(args.slots, e2e_slots),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "estimate_kernel<%d,2,false,false,1>" % (512 if plan.info.threads > 512 else 256), "kernel_ms": kern_ms,
                         "algorithmic_bytes_per_unit": bpu, "units_per_launch": n_local, "peak_source": peak_src},
            "e2e": {"value": LAYERS * n_e2e * world * e2e_steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes,
                    "steps": e2e_steps, "units_per_step": n_e2e, "replication": pipe.replicate,
                    "host_threads": host_threads or os.cpu_count(), "grid_bit_identical_to_device_path": e2e_grid_ok,
                    "scalars_match_device_path_2e-6": e2e_sc_ok,
                    "dense_copy": {"value": LAYERS * n_e2e * world * e2e_steps / dense_s, "unit": UNIT,
                                   "h2d_bytes_per_step": pipe_dense.h2d_bytes, "d2h_bytes_per_step": pipe_dense.d2h_bytes,
                                   "grid_bit_identical_to_device_path": dense_grid_ok and dense_sc_ok,
                                   "note": "the fused kernel writes the dense grid on the device and the whole grid crosses PCIe"}},
            "gpu_launches": launches, "clocks": sampler.summary(wall0, wall1),
        }
        if shared is not None:
            shared["roofline_frac"] = (bpu - 8 * 2 * N_PRB * 3 * len(PILOT_SYMBOLS) * (UES - 1) // UES) * n_local / \
                (shared["kernel_ms"] * 1e-3) / 1e9 / peak
            line["shared_rx_grids"] = shared
        if compact is not None:
            line["compact_output"] = compact
        if alt is not None:
            alt["roofline_frac"] = bpu * n_local / (alt["kernel_ms"] * 1e-3) / 1e9 / peak
            line["alt_layout"] = alt
        if cpu is not None:
            line["cpu_baseline"] = cpu
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def run_config4(args):
    """BASELINE configs[3]: the ce_dl_cnn "CNN refinement" (fill_ch_est_cdm of ce_dl_cnn.py:233-322) on a batch of
    4096 (UE, antenna) smoothed pilot vectors at 273 PRB / comb-4 / 4 ports -> dense [3276, 14, 4] grids, through
    srsce_fill.  The line's value is the tcgen05 implementation the north star names; the direct implementation
    (float32 FMAs in the fused kernel's writer warps -- the "stencil comparator" of SURVEY 7.4) and the whole
    ce_dl_cnn estimator are reported next to it.  Single GPU (the scaling line is config 2's)."""
    import torch
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    from srsran_ce_pytorch_b200 import _lib, estimator, synth

    n_units = args.slots * UNITS_PER_SLOT                      # 16 slots x 256 = 4096 UE-antenna grids
    cpu = None
    if not args.no_cpu_baseline:
        cpu = cpu_fill_baseline()
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    hop1, hop2, cfg = geometry()
    plan = estimator.make_plan(N_PRB, N_SYM, LAYERS, hop1, hop2, cfg, "cnn")
    i = plan.info
    g = torch.Generator(device=dev).manual_seed(4)
    h = torch.view_as_complex(torch.randn(n_units, 1, i.n_re, LAYERS, 2, generator=g, device=dev))
    out = torch.empty((n_units, i.n_sc, N_SYM, LAYERS), dtype=torch.complex64, device=dev)
    bpu = 8 * (i.n_re * LAYERS + i.n_sc * N_SYM * LAYERS)       # read N*L*8, write n_sc*n_sym*L*8

    def timed(fn, steps):
        for _ in range(max(3, args.warmup)):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    sampler = ClockSampler(physical_gpu_index(0))
    sampler.start()
    for _ in range(200):
        if sampler.ready:
            break
        time.sleep(0.01)
    wall0 = time.perf_counter()
    launches = 0

    def tensor_step():
        nonlocal launches
        plan.fill(h, impl="tensor", out=out)
        launches += _lib.lib.srsce_last_launch_count()

    ms_t = timed(tensor_step, args.steps)
    launches_timed = args.steps
    wall1 = time.perf_counter()
    sampler.stop_flag = True
    sampler.join(timeout=2.0)
    ms_d = timed(lambda: plan.fill(h, impl="direct", out=out), args.steps)
    # the whole ce_dl_cnn estimator (gather ... smoothing ... CNN fill) on the same geometry, for reference
    rx, pil = make_step_inputs(min(args.slots, 8), 1234, dev)
    sc = torch.empty((rx.shape[0], 6), dtype=torch.float64, device=dev)
    ms_e = timed(lambda: plan.estimate(rx, pil, 1.0, RX_ANT, out=out[:rx.shape[0]], scalars=sc), args.steps)
    # end to end: pinned pilot vectors in, dense grids back in host memory (one chunked pass)
    n_e2e = min(n_units, 2048)
    h_h = h[:n_e2e].cpu().pin_memory()
    h_out = torch.empty((n_e2e, i.n_sc, N_SYM, LAYERS), dtype=torch.complex64).pin_memory()
    from srsran_ce_pytorch_b200.hostio import FillHostPipeline
    want64 = plan.fill(h[:64], impl="tensor").cpu()

    def e2e_of(pipe):
        h_out.fill_(float("nan"))
        pipe.run(h_h, h_out)
        w0 = time.perf_counter()
        for _ in range(3):
            pipe.run(h_h, h_out)
        dt = (time.perf_counter() - w0) / 3
        # (the direct and the tensor implementation agree to float32 rounding, not bit for bit: compare accordingly)
        same = bool(torch.equal(h_out[:64], want64)) if pipe.replicate == "device" else \
            bool(torch.equal(torch.view_as_real(h_out[:64]), torch.view_as_real(plan.fill(h[:64], impl="direct").cpu())))
        return dt, same and not bool(torch.isnan(torch.view_as_real(h_out[-1])).any()), pipe.h2d_bytes, pipe.d2h_bytes

    # the tensor-core fill writes the dense grid on the device and the whole grid crosses PCIe ...
    dense_s, dense_ok, dense_h2d, dense_d2h = e2e_of(FillHostPipeline(plan, 256, dev, n_streams=2, replicate="device", impl="tensor"))
    # ... or (the entry point's default) the per-hop columns cross PCIe and host threads replicate them over the symbols
    e2e_s, e2e_ok, e2e_h2d, e2e_d2h = e2e_of(FillHostPipeline(plan, 256, dev, n_streams=3, replicate="host"))

    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        pk = json.loads(peaks_path.read_text())
        peak, tpeak, peak_src = float(pk["hbm_gbs"]), float(pk["bf16_tflops"]), "MEASURED_PEAKS.json hbm_gbs (copy, burst)"
    else:
        peak, tpeak, peak_src = 6650.0, 1590.0, "fallback (B200_PROFILING.md)"
    prof = ROOT / "profiles" / "traffic.json"
    tr = json.loads(prof.read_text()) if prof.exists() else {}
    # useful operator FLOPs (2 * 2 * taps * n_sc * L, taps <= 8) and the padded bf16 MMA FLOPs actually issued
    mma_flops = 2.0 * 128 * (i.fill_tensor_k * i.fill_tensor_terms) * (2 * LAYERS) * i.fill_tensor_tiles      # per unit
    line = {
        "metric": METRIC, "value": LAYERS * n_units / (ms_t * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 x3 operand terms, f32 accumulation (tcgen05); complex64 data", "data": "synthetic",
        "config": {"workload": "config4: ce_dl_cnn CNN refinement (fill_ch_est_cdm) as a tcgen05 banded GEMM, 273 PRB, comb-4, "
                               "4 ports, batch of %d UE-antenna pilot vectors -> dense [3276, 14, 4] grids" % n_units,
                   "units_per_step_per_gpu": n_units, "estimates_per_unit": LAYERS,
                   "l2_policy": "output per step (%.1f GB) exceeds the 126 MB L2; no flush needed" % (n_units * bpu / 1e9)},
        "roofline": {"bound": "hbm", "achieved": bpu * n_units / (ms_t * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": bpu * n_units / (ms_t * 1e-3) / 1e9 / peak, "traffic": tr.get("fill_tc_dram_bytes_per_launch_4096_units"),
                     "kernel": "fill_tc_kernel<4>", "kernel_ms": ms_t, "algorithmic_bytes_per_unit": bpu,
                     "units_per_launch": n_units, "peak_source": peak_src},
        "tensor": {"mma_tflops_issued": mma_flops * n_units / (ms_t * 1e-3) / 1e12, "peak_bf16_tflops": tpeak,
                   "frac_of_tensor_peak": mma_flops * n_units / (ms_t * 1e-3) / 1e12 / tpeak,
                   "sm__pipe_tensor_cycles_active_pct": tr.get("fill_tc_tensor_pipe_active_pct"),
                   "note": "0.4 FLOP/B useful work: the kernel is HBM-write bound by construction; tensor-pipe utilisation "
                           "is reported because the north star asks for it, not as a target"},
        "direct_fill": {"value": LAYERS * n_units / (ms_d * 1e-3), "unit": UNIT, "kernel_ms": ms_d,
                        "roofline_frac": bpu * n_units / (ms_d * 1e-3) / 1e9 / peak,
                        "what": "same operator with float32 FMAs in the fused kernel's writer warps (srsce_fill direct)"},
        "full_ce_dl_cnn_estimator": {"value": LAYERS * rx.shape[0] / (ms_e * 1e-3), "unit": UNIT, "kernel_ms": ms_e, "units": rx.shape[0],
                                     "roofline_frac": algorithmic_bytes_per_unit() * rx.shape[0] / (ms_e * 1e-3) / 1e9 / peak},
        "e2e": {"value": LAYERS * n_e2e / e2e_s, "unit": UNIT, "h2d_bytes_per_step": e2e_h2d, "d2h_bytes_per_step": e2e_d2h,
                "units_per_step": n_e2e, "replication": "host (srsce_fill_compact + srsce_replicate_host)",
                "grid_bit_identical_to_device_path": e2e_ok,
                "dense_copy": {"value": LAYERS * n_e2e / dense_s, "unit": UNIT, "h2d_bytes_per_step": dense_h2d,
                               "d2h_bytes_per_step": dense_d2h, "grid_bit_identical_to_device_path": dense_ok,
                               "note": "fill_tc_kernel writes the dense grid on the device and the whole grid crosses PCIe"}},
        "gpu_launches": launches_timed, "clocks": sampler.summary(wall0, wall1),
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())


def cpu_fill_baseline(units_per_worker=24):
    """The oracle's ce_dl_cnn band fill (409 inpainting iterations at 273 PRB) on all host cores, bounded sample."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_fill_worker, [2] * cores)                 # warm-up (imports)
        t0 = time.perf_counter()
        pool.map(_cpu_fill_worker, [units_per_worker] * cores)
        wall = time.perf_counter() - t0
    units = cores * units_per_worker
    return {"value": LAYERS * units / wall, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{units} units ({units_per_worker} per worker, 1 thread each, {wall:.2f} s wall); numpy port of "
                      "ce_dl_cnn.fill_ch_est_cdm (both CDM groups of a 273-PRB comb-4 hop)"}


def _cpu_fill_worker(n):
    import numpy as np
    from oracle import ce_oracle
    from srsran_ce_pytorch_b200 import synth
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    spec = ce_oracle.hop_from_any(synth.make_hop(N_PRB, 0, N_PRB, list(PILOT_SYMBOLS), "comb4", 2))
    rng = np.random.default_rng(os.getpid())
    h = (rng.standard_normal((N_PRB * 3, 2)) + 1j * rng.standard_normal((N_PRB * 3, 2))).astype(np.complex64)
    for _ in range(n):
        for g in range(2):
            ce_oracle._fill_band(h, spec, g, "cnn")
    return n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--slots", type=int, default=0, help="slots per step and per GPU (default: 64 for config2, 2 for config3, 16 for config4)")
    ap.add_argument("--workload", choices=["config2", "config3", "config4"], default="config2")
    ap.add_argument("--e2e-slots", type=int, default=8)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-alt-layout", action="store_true")
    args = ap.parse_args()
    args.slots = select_workload(args.workload, args.slots)
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "config4":
        run_config4(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
