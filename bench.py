#!/usr/bin/env python3
"""bench.py -- ASTC encode throughput on B200 (the BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path (texel fetch -> PCA endpoint fit -> weight
quantisation -> BISE packing -> 128-bit block store) over one synthetic
16384x16384 RGBA8 texture per GPU, ASTC 4x4, RGB, linear -- the configuration
north_star quotes its target on.  Multi-GPU shards by texture (one texture of
the batch per rank, no collective), so scaling is "weak".

value      device-resident throughput, CUDA events on the launching stream,
           max over ranks (Mtexels/s, all ranks' texels / slowest rank's time).
e2e        the same metric through the C-ABI host call astc_b200_encode_host():
           pinned host input -> H2D -> kernel -> D2H -> pinned host output,
           every step, inside the timed region.
roofline   HBM roofline of the encode kernel from ALGORITHMIC bytes (5 B/texel
           at 4x4) over its average launch duration.
cpu_baseline  the CPU oracle (oracle/, a scalar port of the reference shader)
           with OpenMP on this host, on a bounded sample of the same texture.

--impl reference: the reference's own implementation cannot run on Linux (HLSL
cs_5_0 through D3D11), so this arm times the oracle port on the host cores.
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

METRIC = "Mtexels/s ASTC 4x4 encode, 16384x16384 RGBA8 (RGB linear)"
UNIT = "Mtexels/s"
W16K = 16384
BYTES_PER_TEXEL_4x4 = 5.0                  # 4 B read + 16/16 B written
HBM_FALLBACK_GBS = 6650.0                  # B200_PROFILING.md fallback


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            if "hbm_gbs" in d:
                return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _traffic(name: str):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture."""
    p = ROOT / "profiles" / "roofline_traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get(name)
        except Exception:
            return None
    return None


# --------------------------------------------------------------------- clocks --
class ClockSampler:
    """Polls NVML for SM clock and throttle reasons while a timed region runs."""
    _REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
                0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting",
                0x10: "sync_boost"}

    def __init__(self, torch_device_index: int, period_s: float = 0.004):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self._h = None
        self.period = period_s
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(torch_device_index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            try:
                self._h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self._h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
            self._nv = pynvml
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception as e:                       # noqa: BLE001
            self._err = repr(e)
            self._h = None

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                for bit, name in self._REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self._h is not None:
            self._stop.clear()
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        if self._thread is not None:
            self._stop.set()
            self._thread.join()
            self._thread = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unsampled"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------ reference --
def run_reference(args) -> int:
    """The reference's CPU-side stand-in: the oracle port on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    from astc_encoder_b200 import synth
    from oracle import oracle as O
    O.lib()
    rows_texels = 1024                                           # 16384 x 1024 texels per step
    img = synth.synth_rgba(W16K, rows_texels, synth.SEED_CFG5).numpy()
    # all host threads this process may run on -- explicitly, because torchrun exports OMP_NUM_THREADS=1
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    for _ in range(max(1, args.warmup)):
        _, used = O.encode_rows(img, 0, rows_texels // 4, block_dim=4, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, used = O.encode_rows(img, 0, rows_texels // 4, block_dim=4, threads=threads)
    dt = (time.perf_counter() - t0) / args.steps
    value = W16K * rows_texels / dt / 1e6
    sample = f"block rows 0..{rows_texels // 4 - 1} ({W16K}x{rows_texels} texels) of the 16384x16384 texture per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "16384x16384 RGBA8, ASTC 4x4, RGB linear", "sample": sample,
                   "note": "reference hot path is HLSL/D3D11 (not runnable on Linux): timed the scalar C port "
                           "oracle/astc_oracle.c, OpenMP over block rows"},
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": int(used), "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "host_cpus": threads,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------ b200 arm --
def _time_launches(torch, fn, steps, warmup, flush=None):
    """Average device time of fn() per call, each call bracketed by events (optionally an L2 flush before)."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(steps):
        if flush is not None:
            flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        total += a.elapsed_time(b)
    return total / steps


def run_b200(args) -> int:
    import numpy as np
    import torch
    import torch.distributed as dist
    import astc_encoder_b200 as A
    from astc_encoder_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            # convenience: re-launch under torchrun on this node
            port = os.environ.get("MASTER_PORT", "29531")
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", port, str(Path(__file__).resolve()), *sys.argv[1:]]
            os.execv(sys.executable, cmd)
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the encoder has no CPU path (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = None
    if world > 1 and not os.environ.get("ASTC_BENCH_NO_NUMA"):
        from astc_encoder_b200 import sharding
        numa = sharding.bind_host_to_gpu(local)                  # before the pinned buffers are allocated
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its "NCCL version ..." banner to STDOUT when the first communicator comes up;
        # stdout is reserved for the one JSON line, so fd 1 points at stderr until that has happened.
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    A.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    opt = A.encode_option()                                      # 4x4, RGB, linear
    tex = synth.synth_rgba(W16K, W16K, synth.SEED_CFG5 + rank, device=dev)   # texture `rank` of the batch
    out = torch.empty((A.output_size(W16K, W16K, opt) // 16, 16), dtype=torch.uint8, device=dev)
    texels = W16K * W16K
    stream = torch.cuda.current_stream()
    sampler = ClockSampler(local)

    # ---- device-resident: W warm-up, then exactly K timed steps ----
    for _ in range(args.warmup):
        A.encode_astc(tex, opt, out=out, stream=stream)
    launches0 = A.launch_count()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with sampler:
        ev[0].record(stream)
        for i in range(args.steps):
            A.encode_astc(tex, opt, out=out, stream=stream)
            ev[i + 1].record(stream)
        torch.cuda.synchronize()
    barrier()
    launches = A.launch_count() - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    per = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps))
    ms_per_step = max_over_ranks(total_ms / args.steps)
    kernel_ms = sum(per) / len(per)                              # == total/steps: launches are back to back
    value = world * texels / (ms_per_step * 1e-3) / 1e6

    # ---- end to end through the C-ABI host entry point, pinned host buffers ----
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    h_in = torch.empty((W16K, W16K, 4), dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty((out.shape[0], 16), dtype=torch.uint8, pin_memory=True)
    h_in.copy_(tex)
    torch.cuda.synchronize()
    h_in_np, h_out_np = h_in.numpy(), h_out.numpy()
    A.encode_astc_host(h_in_np, opt, out=h_out_np)              # warm-up (allocations land in the pool)
    l0 = A.launch_count()
    barrier()
    with sampler:
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            A.encode_astc_host(h_in_np, opt, out=h_out_np)      # synchronous: returns after the D2H copy
        e2e_s = (time.perf_counter() - t0) / e2e_steps
    barrier()
    e2e_launches = A.launch_count() - l0
    e2e_s = max_over_ranks(e2e_s)
    e2e_identical = bool(torch.equal(h_out, out.cpu()))        # before the probe below overwrites h_out
    # what the link alone costs: the same two pinned buffers copied H2D and D2H concurrently, no kernel
    d_probe_in, d_probe_out = torch.empty_like(tex), torch.empty_like(out)
    side = torch.cuda.Stream()

    def copies():
        d_probe_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(side):
            h_out.copy_(d_probe_out, non_blocking=True)
        torch.cuda.current_stream().wait_stream(side)

    copies()
    torch.cuda.synchronize()
    ca, cb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ca.record()
    for _ in range(3):
        copies()
    cb.record()
    torch.cuda.synchronize()
    link_ms = ca.elapsed_time(cb) / 3
    del d_probe_in, d_probe_out
    e2e_value = world * texels / e2e_s / 1e6

    # ---- BASELINE config 5 as written: ONE 16384^2 texture in block-row bands + 512 mip chains dealt by texture ----
    cfg5 = None
    if not args.no_others:
        del h_in, h_out, h_in_np, h_out_np
        cfg5 = config5(torch, dist, A, synth, dev, rank, world, barrier, max_over_ranks, tex if world == 1 else None, out, args)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = _peaks()
    achieved = texels * BYTES_PER_TEXEL_4x4 / (kernel_ms * 1e-3) / 1e9
    # The bound that actually binds: issue slots and the FP32 pipe.  Static per-block instruction counts
    # come from the committed ncu capture, time and SM clock are this run's.
    clocks = sampler.summary()
    sm_hz = float(clocks.get("sm_mhz") or 0) * 1e6
    smsps = torch.cuda.get_device_properties(0).multi_processor_count * 4
    instr32, fma32 = _traffic("encode4x4_rgb_warp_instr_per_32_blocks"), _traffic("encode4x4_rgb_fma_pipe_cycles_per_32_blocks")
    compute = None
    if instr32 and fma32 and sm_hz > 0:
        slots = kernel_ms * 1e-3 * sm_hz * smsps                     # issue slots (= FP32-pipe cycles) the launch had
        warp_blocks = texels / 16 / 32
        compute = {"bound": "issue / FP32 pipe", "issue_frac": round(instr32 * warp_blocks / slots, 4),
                   "fma_pipe_frac": round(fma32 * warp_blocks / slots, 4),
                   "warp_instr_per_32_blocks": instr32, "fma_pipe_cycles_per_32_blocks": fma32,
                   "note": "1 warp-instruction and 32 FP32 lanes per clock per SMSP; counts from profiles/roofline_traffic.json"}
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "16384x16384 RGBA8, ASTC 4x4, RGB linear (one texture per GPU, sharded by texture)",
                   "blocks_per_gpu": texels // 16, "l2": "input 1 GiB + output 256 MiB per step exceed the 126 MB L2; no flush needed",
                   "timing": "CUDA events on the launching stream, max over ranks"},
        "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": texels * 4,
                "d2h_bytes_per_step": texels, "steps": e2e_steps, "ms_per_step": round(e2e_s * 1e3, 3),
                "api": "astc_b200_encode_host (C ABI), pinned host buffers", "launches": int(e2e_launches),
                "matches_device_path": e2e_identical,
                "link_only_ms": round(link_ms, 3), "host_numa_binding": numa,
                "note": "PCIe-bound: link_only_ms is the same H2D + D2H traffic with no kernel at all"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": _traffic("encode4x4_rgb"),
                     "kernel": "encode4x4_kernel<rgb,linear>", "kernel_ms": round(kernel_ms, 4),
                     "kernel_ms_best": round(per[0], 4), "algorithmic_bytes_per_launch": int(texels * BYTES_PER_TEXEL_4x4),
                     "peak_source": peak_src, "compute": compute,
                     "note": "FP32-pipe / register-operand-bandwidth bound (~1050 warp-instructions per 32 blocks of 80 B): DESIGN.md 4.1"},
    }

    # ---- CPU baseline + parity on a bounded sample of the same texture (rank 0, N=1) ----
    if world == 1 and not args.no_cpu:
        from oracle import oracle as O
        O.lib()
        import numpy as np
        ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        # timing: the reference arm's crop (16384 x 1024 texels), one warm-up, median of 5
        crop_rows = 1024
        crop = tex[:crop_rows].cpu().numpy()

        def oracle_rate(rows_texels, threads, runs):
            O.encode_rows(crop, 0, rows_texels // 4, block_dim=4, threads=threads)          # warm-up
            ts, used = [], threads
            for _ in range(runs):
                t0 = time.perf_counter()
                _, used = O.encode_rows(crop, 0, rows_texels // 4, block_dim=4, threads=threads)
                ts.append(time.perf_counter() - t0)
            ts.sort()
            return W16K * rows_texels / ts[len(ts) // 2] / 1e6, int(used), ts

        mt, used, ts = oracle_rate(crop_rows, ncpu, 5)
        st_rows = 128                                               # single thread: 16384 x 128 texels per run
        st, _, _ = oracle_rate(st_rows, 1, 3)
        line["cpu_baseline"] = {"value": round(mt, 2), "unit": UNIT, "cores": used, "kind": "port",
                                "sample": f"block rows 0..{crop_rows // 4 - 1} ({W16K}x{crop_rows} texels) of the same texture, "
                                          "oracle/astc_oracle.c with OpenMP; 1 warm-up, median of 5",
                                "runs_ms": [round(t * 1e3, 1) for t in ts],
                                "single_thread": {"value": round(st, 2), "unit": UNIT, "cores": 1,
                                                  "sample": f"{W16K}x{st_rows} texels, 1 warm-up, median of 3"}}
        # parity: 4 194 304 blocks of the GPU output against the oracle
        rows_texels = 4096
        sample = tex[:rows_texels].cpu().numpy()
        want, _ = O.encode_rows(sample, 0, rows_texels // 4, block_dim=4, threads=ncpu)
        got = out[: want.shape[0]].cpu().numpy()
        same = int((got == want).all(axis=1).sum())
        line["parity"] = {"blocks_checked": int(want.shape[0]), "bit_identical": same}
        # decoded PSNR (per channel, dB) of the GPU's blocks (device decoder) and of the oracle's blocks
        # (oracle decoder) against the source, on the first 1024 texel rows of the sample: the metric's
        # "decoded PSNR delta vs ref" (0 when every block is bit-identical)
        prow = 1024
        nblk = (prow // 4) * (W16K // 4)
        dec_gpu = A.decode_astc(out[:nblk], W16K, prow, 4).cpu().numpy()
        dec_ref, nbad = O.decode_image(want[:nblk], W16K, prow, 4)
        src = sample[:prow]
        p_gpu, p_ref = O.psnr_per_channel(dec_gpu[..., :3], src[..., :3]), O.psnr_per_channel(dec_ref[..., :3], src[..., :3])
        line["parity"].update({"psnr_db_rgb": [round(float(v), 3) for v in p_gpu],
                               "psnr_delta_db_vs_oracle": round(float(np.max(np.abs(p_gpu - p_ref))), 4),
                               "undecodable_blocks": int(nbad)})

    # ---- the other BASELINE.json configs, kernel-only, L2 flushed between launches ----
    if world == 1 and not args.no_others:
        line["others"] = other_configs(torch, A, synth, dev, peak)
        line["e2e_small"] = host_small_textures(torch, A, synth)
    if cfg5 is not None:
        line["config5"] = cfg5

    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def config5(torch, dist, A, synth, dev, rank, world, barrier, max_over_ranks, full_tex, full_out, args):
    """BASELINE.json configs[4], sharded the two ways north_star names, device-resident, no collective on the
    data path: (a) ONE 16384x16384 texture cut into `world` contiguous block-row bands (astc_b200_band), every
    rank encoding its band into its slice of the output; (b) 512 2048x2048 12-level mip chains dealt out whole
    by sharding.assign_textures (longest first), one batch launch per rank.  Strong scaling: the total work is
    fixed.  Times are CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.
    Both outputs are checked byte for byte ON THE GPUS against the single-GPU encode of the same input."""
    from astc_encoder_b200 import sharding
    opt = A.encode_option()
    steps, warm = max(5, min(args.steps, 20)), 3
    stream = torch.cuda.current_stream()

    def all_true(flag: bool) -> bool:
        if world == 1:
            return bool(flag)
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def timed(fn):
        for _ in range(warm):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(steps):
            fn()
        b.record(stream)
        torch.cuda.synchronize()
        barrier()
        return max_over_ranks(a.elapsed_time(b) / steps)

    # ---- (a) one texture, `world` bands ----
    band = sharding.band_plan(W16K, W16K, opt, world)[rank]
    slab = full_tex if world == 1 else synth.synth_rgba(W16K, W16K, synth.SEED_CFG5, device=dev, row0=band.y0, rows=band.rows)
    band_out = torch.empty((band.nbytes // 16, 16), dtype=torch.uint8, device=dev)
    band_ms = timed(lambda: A.encode_astc(slab, opt, out=band_out, stream=stream))
    # the single-GPU encode of the whole texture, on this rank's GPU (deterministic, so the same on every rank)
    if world > 1:
        full_tex = synth.synth_rgba(W16K, W16K, synth.SEED_CFG5, device=dev)
        full_out = torch.empty((W16K // 4 * (W16K // 4), 16), dtype=torch.uint8, device=dev)
    single_ms = None
    if world > 1:
        for _ in range(warm):
            A.encode_astc(full_tex, opt, out=full_out, stream=stream)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(steps):
            A.encode_astc(full_tex, opt, out=full_out, stream=stream)
        b.record(stream)
        torch.cuda.synchronize()
        single_ms = max_over_ranks(a.elapsed_time(b) / steps)
    else:
        A.encode_astc(full_tex, opt, out=full_out, stream=stream)
        torch.cuda.synchronize()
    lo = band.byte_offset // 16
    band_same = all_true(torch.equal(full_out[lo:lo + band.nbytes // 16], band_out))
    if world > 1:
        del full_tex, full_out
    del slab
    texels_a = W16K * W16K

    # ---- (b) 512 mip chains dealt by texture ----
    chains = args.chains
    chain_texels = sum(max(1, 2048 >> l) ** 2 for l in range(12))
    mine = sharding.assign_textures([chain_texels] * chains, world)[rank]
    srcs = []
    for i in mine:
        srcs.extend(A.mip_chain(synth.synth_rgba(2048, 2048, synth.SEED_BATCH + i, device=dev)))     # device 2x2 box filter
    batch = A.Batch(srcs, opt)
    batch_ms = timed(lambda: batch.encode(stream=stream))
    outs = batch.encode(stream=stream)
    torch.cuda.synchronize()
    # byte check against the single-texture path (one encode_astc launch per level) on this GPU, and a checksum
    # of checksums that does not depend on how the chains were dealt
    same = True
    csum = torch.zeros((), dtype=torch.int64, device=dev)
    for src, o in zip(srcs, outs):
        same = same and bool(torch.equal(A.encode_astc(src, opt, stream=stream), o))
        csum += o.view(torch.int64).sum()
    if world > 1:
        dist.all_reduce(csum, op=dist.ReduceOp.SUM)
    batch_same = all_true(same)
    texels_b = chains * chain_texels

    # ---- the whole of config 5 in one step: band launch + batch launch per rank ----
    slab2 = synth.synth_rgba(W16K, W16K, synth.SEED_CFG5, device=dev, row0=band.y0, rows=band.rows)

    def both():
        A.encode_astc(slab2, opt, out=band_out, stream=stream)
        batch.encode(stream=stream)

    both_ms = timed(both)
    launches_per_step = 2
    host_batch = None
    if world == 1 and not args.no_host_batch:
        host_batch = host_batch_e2e(torch, A, srcs, outs, opt, chains, chain_texels)
    total_blocks = int(batch.total_blocks)
    batch.close()
    del srcs, outs, slab2
    res = {
        "workload": "BASELINE configs[4]: 16384x16384 RGBA8 4x4 RGB + 512 x 2048x2048 12-level mip chains, device-resident",
        "scaling": "strong", "n_gpus": world, "steps": steps, "unit": UNIT,
        "band": {"what": f"ONE 16384x16384 texture in {world} block-row band(s) (astc_b200_band), one launch per rank",
                 "ms": round(band_ms, 4), "value": round(texels_a / band_ms / 1e3, 1),
                 "bytes_identical_to_single_gpu_encode": band_same, "blocks_per_rank": band.nbytes // 16},
        "batch": {"what": f"{chains} mip chains dealt whole by sharding.assign_textures (LPT), ONE batch launch per rank",
                  "ms": round(batch_ms, 4), "value": round(texels_b / batch_ms / 1e3, 1), "chains_per_rank": len(mine),
                  "blocks_this_rank": total_blocks, "bytes_identical_to_per_texture_encode": batch_same,
                  "checksum_of_all_blocks": f"{int(csum.item()) & 0xFFFFFFFFFFFFFFFF:016x}"},
        "both": {"what": "band launch + batch launch per step (the whole of configs[4])", "ms": round(both_ms, 4),
                 "value": round((texels_a + texels_b) / both_ms / 1e3, 1), "launches_per_step_per_rank": launches_per_step},
    }
    if host_batch is not None:
        res["batch"]["e2e_host"] = host_batch
    if single_ms is not None:
        res["band"]["single_gpu_ms_same_run"] = round(single_ms, 4)
        res["band"]["strong_scaling_efficiency"] = round(single_ms / (world * band_ms), 4)
    return res


def _mem_available_gib() -> float:
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                return int(ln.split()[1]) / (1 << 20)
    except OSError:
        pass
    return 0.0


def host_batch_e2e(torch, A, srcs, outs, opt, chains, chain_texels):
    """The 512-chain batch END TO END from host memory: astc_b200_context_batch_encode_host with every level in
    one pinned host buffer (H2D of all levels + kernels + D2H of all blocks inside the timed region), checked
    against the device-resident batch output."""
    import numpy as np
    levels = len(srcs) // chains
    # pinned memory needed: all sources + all outputs; use fewer chains if the host is short of memory
    per_chain_in = sum(int(t.numel()) for t in srcs[:levels])
    per_chain_out = sum(int(t.numel()) for t in outs[:levels])
    avail = _mem_available_gib() * (1 << 30)
    use = chains
    while use > 8 and use * (per_chain_in + per_chain_out) > 0.4 * avail:
        use //= 2
    n = use * levels
    h_in = torch.empty(use * per_chain_in, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(use * per_chain_out, dtype=torch.uint8, pin_memory=True)
    imgs, dsts, oi, oo = [], [], 0, 0
    for t, o in zip(srcs[:n], outs[:n]):
        v = h_in[oi:oi + t.numel()].view(t.shape)
        v.copy_(t)
        imgs.append(v.numpy())
        dsts.append(h_out[oo:oo + o.numel()].view(o.shape).numpy())
        oi += t.numel()
        oo += o.numel()
    torch.cuda.synchronize()
    ctx = A.Context()
    ctx.batch_encode_host(imgs, opt, outs=dsts)                   # warm-up: workspace + staging grow here
    l0 = A.launch_count()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        ctx.batch_encode_host(imgs, opt, outs=dsts)
        ts.append(time.perf_counter() - t0)
    launches = (A.launch_count() - l0) // 3
    ts.sort()
    same = all(bool(np.array_equal(d, o.cpu().numpy())) for d, o in list(zip(dsts, outs[:n]))[:: max(1, n // 200)])
    # the same chains from their BASE levels only (astc_b200_context_batch_encode_mip_chains_host): the bases are
    # uploaded (three quarters of the bytes), the levels below are generated AND encoded on the device; the output
    # layout per chain -- base first, then level 1, 2, ... -- is the layout h_out already has
    import ctypes as C
    from astc_encoder_b200 import _HostImage
    expect = h_out.clone()
    h_out.zero_()
    recs = (_HostImage * use)()
    for c in range(use):
        b = imgs[c * levels]
        recs[c] = _HostImage(b.ctypes.data, dsts[c * levels].ctypes.data, b.strides[0], b.shape[1], b.shape[0])
    o = opt._abi()
    L = A.lib()
    assert L.astc_b200_context_batch_encode_mip_chains_host(ctx._h, recs, use, C.byref(o)) == 0
    bts = []
    for _ in range(3):
        t0 = time.perf_counter()
        rc = L.astc_b200_context_batch_encode_mip_chains_host(ctx._h, recs, use, C.byref(o))
        bts.append(time.perf_counter() - t0)
        assert rc == 0
    bts.sort()
    bases_same = bool(torch.equal(h_out, expect))
    from_bases = {"api": "astc_b200_context_batch_encode_mip_chains_host (C ABI): only the base levels are uploaded, pinned memory",
                  "chains": use, "ms": round(bts[1] * 1e3, 2), "value": round(use * chain_texels / bts[1] / 1e6, 1), "unit": UNIT,
                  "h2d_bytes": int(sum(int(imgs[c * levels].nbytes) for c in range(use))), "d2h_bytes": use * per_chain_out,
                  "matches_all_levels_uploaded": bases_same}
    # the same call from PAGEABLE memory (plain numpy arrays -- what a caller without CUDA allocations holds): the
    # levels travel through the context's pinned slots, copied by its worker threads; 64 chains bound the host memory
    pg_use = min(use, 64)
    pg_n = pg_use * levels
    pg_imgs = [np.array(v, copy=True) for v in imgs[:pg_n]]
    pg_dsts = [np.empty_like(d) for d in dsts[:pg_n]]
    ctx.batch_encode_host(pg_imgs, opt, outs=pg_dsts)
    pts = []
    for _ in range(3):
        t0 = time.perf_counter()
        ctx.batch_encode_host(pg_imgs, opt, outs=pg_dsts)
        pts.append(time.perf_counter() - t0)
    pts.sort()
    pg_same = all(bool(np.array_equal(a, b)) for a, b in zip(pg_dsts, dsts[:pg_n]))
    ctx.close()
    del h_in, h_out
    return {"api": "astc_b200_context_batch_encode_host (C ABI), every level in pinned host memory", "chains": use, "textures": n,
            "ms": round(ts[1] * 1e3, 2), "value": round(use * chain_texels / ts[1] / 1e6, 1), "unit": UNIT,
            "h2d_bytes": use * per_chain_in, "d2h_bytes": use * per_chain_out, "launches_per_call": int(launches),
            "matches_device_batch": same, "from_bases": from_bases,
            "pageable": {"chains": pg_use, "ms": round(pts[1] * 1e3, 2), "value": round(pg_use * chain_texels / pts[1] / 1e6, 1),
                         "unit": UNIT, "matches_pinned": pg_same,
                         "note": "sources and outputs in pageable numpy memory, staged through pinned slots by the context's copy workers"}}


def host_small_textures(torch, A, synth):
    """Per-call cost of astc_b200_encode_host (persistent thread-local context) on small textures, pinned and
    pageable host memory: what one texture costs end to end when the job is too small to hide anything."""
    import numpy as np
    res = []
    opt = A.encode_option()
    for size in (4, 256, 1024, 4096):
        src = synth.synth_rgba(size, size, 11)
        pinned = torch.empty((size, size, 4), dtype=torch.uint8, pin_memory=True)
        pinned.copy_(src)
        pout = torch.empty((A.output_size(size, size, opt) // 16, 16), dtype=torch.uint8, pin_memory=True)
        row = {"workload": f"{size}x{size} RGBA8, 4x4 RGB, one astc_b200_encode_host call"}
        for kind, arr, out in (("pinned", pinned.numpy(), pout.numpy()), ("pageable", src.numpy().copy(), np.empty(pout.shape, np.uint8))):
            for _ in range(5):
                A.encode_astc_host(arr, opt, out=out)
            iters = 200 if size <= 1024 else 30
            ts = []
            for _ in range(iters):
                t0 = time.perf_counter()
                A.encode_astc_host(arr, opt, out=out)
                ts.append(time.perf_counter() - t0)
            ts.sort()
            med = ts[len(ts) // 2]
            row[kind] = {"us_per_call": round(med * 1e6, 1), "value": round(size * size / med / 1e6, 1), "unit": UNIT}
        res.append(row)
    return res


def other_configs(torch, A, synth, dev, peak):
    res = []
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def flush():
        flush_buf.zero_()

    def one(name, img, opt, dim):
        out = A.encode_astc(img, opt)
        ms = _time_launches(torch, lambda: A.encode_astc(img, opt, out=out), 10, 3, flush)
        h, w = int(img.shape[0]), int(img.shape[1])
        nbytes = w * h * 4 + out.numel()
        # the same launch issued BACK TO BACK over rotating copies of the input that together exceed the L2 (what a stream of
        # textures looks like: programmatic dependent launch hides the next launch's start-up under the tail of this one);
        # two events around the whole run, so the timer's ~2 us tick is amortised
        copies = max(2, min(8, (400 << 20) // (w * h * 4)))
        imgs = [img] + [img.clone() for _ in range(copies - 1)]
        outs = [out] + [torch.empty_like(out) for _ in range(copies - 1)]
        for i, o in zip(imgs, outs):
            A.encode_astc(i, opt, out=o)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            for i, o in zip(imgs, outs):
                A.encode_astc(i, opt, out=o)
        b.record()
        b.synchronize()
        b2b = a.elapsed_time(b) / (3 * copies)
        res.append({"workload": name, "value": round(w * h / ms / 1e3, 1), "unit": UNIT, "kernel_ms": round(ms, 4),
                    "hbm_frac": round(nbytes / (ms * 1e-3) / 1e9 / peak, 4),
                    "back_to_back": {"kernel_ms": round(b2b, 4), "value": round(w * h / b2b / 1e3, 1),
                                     "hbm_frac": round(nbytes / (b2b * 1e-3) / 1e9 / peak, 4), "rotating_inputs": copies}})
        del imgs, outs

    one("4096x4096 RGBA8, 4x4, RGB linear", synth.synth_rgba(4096, 4096, synth.SEED_CFG2, device=dev),
        A.encode_option(), 4)
    one("8192x8192 RGBA8, 6x6, -alpha -srgb", synth.synth_rgba(8192, 8192, synth.SEED_CFG3, device=dev),
        A.encode_option(is6x6=True, has_alpha=True, srgb=True), 6)
    one("4096x4096 normal map, -norm -4x4", synth.synth_normal(4096, 4096, synth.SEED_CFG4, device=dev),
        A.encode_option(is_normal_map=True), 4)
    return res


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU oracle baseline / parity sample")
    ap.add_argument("--no-others", action="store_true", help="skip the secondary configs (incl. config 5)")
    ap.add_argument("--no-host-batch", action="store_true", help="skip the end-to-end host batch of config 5")
    ap.add_argument("--chains", type=int, default=512, help="mip chains in the config-5 batch (default: BASELINE's 512)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3                                          # timing rule: W >= 3
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
