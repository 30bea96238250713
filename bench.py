#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native rt.resynth hot path.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port) on the host cores

Headline metric (BASELINE.json): STFT frames/s at 4096 points. A step = one pass of the fused analysis kernel
(window -> real FFT -> |X|^2 -> peak extraction) over a batch of `streams_per_gpu` independent 60 s / 48 kHz streams
(BASELINE config 2 per stream, the 32-streams-per-GPU sharding of config 4), inputs resident in HBM. `e2e` is the same
metric through the host-buffer C-ABI entry point (pinned host input, peak lists copied back). `extra.synthesis` carries
the second half of the metric (oscillator-bank samples/s, BASELINE config 3). One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

SR, WINDOW, HOP, NFFT = 48000, 4096, 256, 4096
SECONDS = 60.0
MAX_PEAKS = 32
MIN_DB = -80.0
FLOP_PER_FRAME = 5 * NFFT // 2 * 12 + 3 * NFFT // 2  # 2.5 N log2 N for the transform + 3 N/2 for |X|^2: 129,024
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only when MEASURED_PEAKS.json is absent


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel: str):
    """Per-launch DRAM bytes of the dominant kernel from the committed ncu --set full capture (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get(kernel)
    return None


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe): NVML from a background
    thread every 10 ms (nvidia-smi -lms answers too slowly for a 40 ms region); nvidia-smi as the fallback."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = False
        self._thread = None
        self._smi = None
        self._smi_file = None

    def _nvml_loop(self, nv, h):
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = int(get_reasons(h))
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        try:
            import threading

            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it lists indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = self.index
            if vis and all(t.strip().isdigit() for t in vis.split(",")):
                ids = [int(t) for t in vis.split(",")]
                if self.index < len(ids):
                    phys = ids[self.index]
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self._thread.start()
            return
        except Exception:
            self._thread = None
        try:
            self._smi_file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self._smi = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=self._smi_file, stderr=subprocess.DEVNULL)
        except Exception:
            self._smi = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self._thread is not None:
            self._stop = True
            self._thread.join(timeout=2)
            if self.samples:
                out.update(sm_mhz=float(np.median(self.samples)), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons),
                           samples=len(self.samples), source="nvml, 10 ms period, warm-up + timed region")
            return out
        if self._smi is None:
            return out
        self._smi.terminate()
        try:
            self._smi.wait(timeout=5)
        except Exception:
            self._smi.kill()
        self._smi_file.flush()
        self._smi_file.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self._smi_file.read().splitlines():
            parts = [t.strip() for t in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm),
                       source="nvidia-smi -lms 100")
        try:
            os.unlink(self._smi_file.name)
        except OSError:
            pass
        return out


def workload_config(n_gpus: int, streams_per_gpu: int) -> dict:
    """The `config` object of the JSON line: the WORKLOAD, identical for both arms (the reference arm times a bounded sample of
    it and says which in cpu_baseline.sample)."""
    frames = streams_per_gpu * (1 + (int(round(SECONDS * SR)) - WINDOW) // HOP)
    return {"workload": "C2: rt.resynth periodic-FFT analysis, 60 s mono 48 kHz harmonic tone per stream, window 4096, hop "
                        "256, Gaussian window (sigma factor 4), |X|^2 spectrum + peak extraction",
            "streams_per_gpu": streams_per_gpu, "frames_per_step_per_gpu": frames, "window": WINDOW, "hop": HOP, "n_fft": NFFT,
            "max_peaks": MAX_PEAKS, "min_db": MIN_DB, "parallelism": f"{n_gpus} x independent streams, no data-path collective",
            "l2": "inputs+outputs per step (0.37 GB + 2.9 GB) exceed the 126 MB L2; no flush needed"}


def cpu_analysis_rate(orc, wl, threads: int, streams: int, seconds: float):
    """frames/s of the CPU oracle (the reference's structure: per-sample feed, double precision, one FFT per hop, peak
    extraction per frame) over `streams` independent streams of `seconds` each on `threads` host threads."""
    x = wl.harmonic_tone(int(round(seconds * SR)))
    t0 = time.perf_counter()
    r = orc.stft_analyze(x, WINDOW, HOP, sample_rate=SR, min_db=MIN_DB, max_peaks=MAX_PEAKS, want_sqmag=False, threads=threads,
                         repeat=streams, count_only=True)
    dt = time.perf_counter() - t0
    assert int(r["peak_count"].min()) == 16 and int(r["peak_count"].max()) == 16
    return streams * r["frames"] / dt, dt


def measure_copy_ceiling(torch, dist, dev, world, mib: int = 256, reps: int = 4):
    """Pinned host <-> device copy rates of this box with every rank copying at once (one cudaMemcpyAsync per direction in flight,
    `mib` MiB each): H2D alone, D2H alone, both directions together. Per rank and summed over the ranks, GB/s. This is the
    ceiling of every host-buffer entry point; it is a property of the host (PCIe root complexes, memory, virtualisation)."""
    n = mib << 20
    h_up = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_dn = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_up = torch.empty(n, dtype=torch.uint8, device=dev)
    d_dn = torch.empty(n, dtype=torch.uint8, device=dev)
    s_up, s_dn = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def run(up: bool, dn: bool):
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s_up.wait_event(e0)
        s_dn.wait_event(e0)
        for _ in range(reps):
            if up:
                with torch.cuda.stream(s_up):
                    d_up.copy_(h_up, non_blocking=True)
            if dn:
                with torch.cuda.stream(s_dn):
                    h_dn.copy_(d_dn, non_blocking=True)
        torch.cuda.current_stream(dev).wait_stream(s_up)
        torch.cuda.current_stream(dev).wait_stream(s_dn)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return reps * n / (ms * 1e-3) / 1e9  # GB/s per direction per rank (slowest rank)

    run(True, True)  # warm-up
    h2d, d2h, duplex = run(True, False), run(False, True), run(True, True)
    return {"mib_per_copy": mib, "ranks": world, "h2d_gbs_per_rank": h2d, "d2h_gbs_per_rank": d2h,
            "duplex_gbs_per_rank_per_direction": duplex, "h2d_gbs_all_ranks": h2d * world, "d2h_gbs_all_ranks": d2h * world,
            "duplex_gbs_all_ranks_both_directions": 2 * duplex * world,
            "note": "every rank copies at once; rates of the slowest rank (max time over ranks)"}


_RESULT_LINE: list = []  # the one line of stdout, printed by main() after the run


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    orc = graft.load_oracle()
    graft.load_package()
    import cpp_audio_b200.workloads as wl
    cores = orc.hardware_threads()
    streams = max(1, cores) * 4      # ~1.5 s of work per thread per step
    seconds = 60.0
    for _ in range(min(args.warmup, 2)):
        cpu_analysis_rate(orc, wl, cores, cores, 10.0)
    t_total, frames_total = 0.0, 0
    for _ in range(args.steps):
        rate, dt = cpu_analysis_rate(orc, wl, cores, streams, seconds)
        t_total += dt
        frames_total += rate * dt
    value = frames_total / t_total
    sample = f"{streams} streams x {seconds:.0f} s per step ({int(frames_total / args.steps)} frames), {cores} threads"
    line = {
        "impl": "reference", "metric": "stft_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus, args.streams_per_gpu),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": sample + "; the CPU oracle port of the reference (double precision, per-sample feed, one FFT "
                                            "per hop, peak extraction per frame; the spectrum is computed, not stored)"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _RESULT_LINE.append(json.dumps(line))
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    pkg = graft.load_package()
    import cpp_audio_b200.workloads as wl
    ctx = pkg.Context(local_rank)
    S = args.streams_per_gpu
    n_samples = int(round(SECONDS * SR))

    # ---- inputs: S streams per rank (weak scaling), resident in HBM before the timed region ----
    x_host_np = wl.tiled_streams(S, SECONDS)
    x_pinned = torch.from_numpy(x_host_np).pin_memory()
    x = x_pinned.to(dev)
    an = pkg.SqMagAnalyzer(ctx, WINDOW, HOP, sample_rate=SR, min_db=MIN_DB, max_peaks=MAX_PEAKS)
    frames = an.num_frames(n_samples)
    rows = S * frames
    out = an.allocate(S, n_samples, dev)

    def step():
        an(x, out)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)  # samples every 25 ms from the warm-up to the end of the timed region
    sampler.start()
    for _ in range(max(3, args.warmup)):
        step()
    sync_all()
    launches_warm = ctx.kernel_launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for a, b in ev:
        a.record()
        step()
        b.record()
    t_end.record()
    sync_all()
    clocks = sampler.stop()
    launches_timed = ctx.kernel_launches - launches_warm
    total_ms = t_start.elapsed_time(t_end)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    if world > 1:
        t = torch.tensor([total_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = world * rows * args.steps / (total_ms * 1e-3)

    # ---- sanity on the timed outputs (cheap, outside the timed region): 16 partials per frame ----
    cnt = out["peak_count"]
    assert int(cnt.min()) == 16 and int(cnt.max()) == 16, "unexpected peak counts: the timed kernel did not do the work"

    # ---- end to end through the host-buffer C ABI: pinned x -> H2D -> kernel -> D2H peak lists ----
    P = MAX_PEAKS
    h_bin = torch.empty((rows, P), dtype=torch.int32).pin_memory()
    h_freq = torch.empty((rows, P), dtype=torch.float32).pin_memory()
    h_db = torch.empty((rows, P), dtype=torch.float32).pin_memory()
    h_cnt = torch.empty((rows,), dtype=torch.int32).pin_memory()
    L = pkg.lib()

    def e2e_step():
        ctx.check(L.ca_stft_analyze_host(ctx.handle, x_pinned.data_ptr(), n_samples, S, n_samples, None, 0, h_bin.data_ptr(),
                                         h_freq.data_ptr(), h_db.data_ptr(), h_cnt.data_ptr()), "ca_stft_analyze_host")

    e2e_steps = max(2, min(args.steps, 5))
    e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * rows * e2e_steps / e2e_s
    assert int(h_cnt.min()) == 16 and int(h_cnt.max()) == 16
    h2d = x_pinned.numel() * 4
    d2h = (h_bin.numel() + h_freq.numel() + h_db.numel() + h_cnt.numel()) * 4

    # ---- the same call with the spectrum copied back as well (2.9 GB per step: what the device-timed step writes to HBM) ----
    e2e_sq = None
    if not args.skip_e2e_spectrum:
        h_sq = torch.empty((rows, NFFT // 2 + 1), dtype=torch.float32).pin_memory()

        def e2e_sq_step():
            ctx.check(L.ca_stft_analyze_host(ctx.handle, x_pinned.data_ptr(), n_samples, S, n_samples, h_sq.data_ptr(), NFFT // 2 + 1,
                                             h_bin.data_ptr(), h_freq.data_ptr(), h_db.data_ptr(), h_cnt.data_ptr()), "ca_stft_analyze_host")

        e2e_sq_step()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(2):
            e2e_sq_step()
        sq_s = (time.perf_counter() - t0) / 2
        if world > 1:
            t = torch.tensor([sq_s], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sq_s = float(t.item())
        e2e_sq = {"value": world * rows / sq_s, "unit": "frames/s", "ms_per_step": 1e3 * sq_s, "h2d_bytes_per_step": h2d,
                  "d2h_bytes_per_step": d2h + h_sq.numel() * 4,
                  "note": "peak lists AND the |X|^2 spectrum copied to the host: bound by the device-to-host copy of 8.2 KB per frame"}
        del h_sq

    # ---- final gather of the peak lists over NVLink (reported separately) ----
    gather_ms = None
    if world > 1:
        payload = torch.cat([out["peak_freq"][:, :16].reshape(-1), out["peak_db"][:, :16].reshape(-1)])
        bufs = [torch.empty_like(payload) for _ in range(world)]
        dist.all_gather(bufs, payload)
        torch.cuda.synchronize(dev)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        dist.all_gather(bufs, payload)
        g1.record()
        torch.cuda.synchronize(dev)
        gather_ms = g0.elapsed_time(g1)

    # ---- synthesis half of the metric: BASELINE config 3 (1024 voices x 64 partials x 10 s) per rank ----
    extra = {}
    if not args.skip_synth:
        b = wl.c3_bank(args.voices, 64)
        tt = {k: torch.from_numpy(v).to(dev) for k, v in b.items()}
        ns = int(round(args.synth_seconds * SR))
        bank = pkg.SineBank(ctx, args.voices, 64, SR)
        env = pkg.Ahdsr(0, 0, 0, 0, 0, 0, 0, 1.0)
        sout = torch.empty((args.voices, ns), dtype=torch.float32, device=dev)

        def synth_step():
            bank.render(ns, tt["f0_hz"], tt["volume"], tt["partial_volume"], tt["partial_phase"], env, out=sout)

        for _ in range(3):
            synth_step()
        sync_all()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(3, min(args.steps, 10))
        s0.record()
        for _ in range(reps):
            synth_step()
        s1.record()
        sync_all()
        sms = s0.elapsed_time(s1) / reps
        if world > 1:
            t = torch.tensor([sms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sms = float(t.item())
        voice_samples = world * args.voices * ns / (sms * 1e-3)

        def timed_variant(path):
            """The same render with the sustain-regime product on another pipe (CA_OSC_PATH)."""
            os.environ["CA_OSC_PATH"] = path
            for _ in range(2):
                synth_step()
            sync_all()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(reps):
                synth_step()
            f1.record()
            sync_all()
            del os.environ["CA_OSC_PATH"]
            ms = f0.elapsed_time(f1) / reps
            if world > 1:
                t = torch.tensor([ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            return ms

        mms, fms = timed_variant("mma"), timed_variant("fp32")
        fp32_peak = ctx.measure_fp32_peak()
        peak_gbs_s, _ = measured_peaks()
        part = voice_samples * 64
        per_rank = args.voices * ns
        extra["synthesis"] = {
            "workload": f"C3: {args.voices} voices x 64 partials x {args.synth_seconds:g} s @48 kHz per GPU, per-voice float32 output",
            "path": "sustain regime as a matrix product per chunk (samples x blocks x 2 partials), FP16 operands split hi/lo (3 products, "
                    "FP32-equivalent), FP32 accumulation: tcgen05.mma with the accumulator in tensor memory for the steady 16,384-sample "
                    "chunks, mma.sync m16n8k16 for the 2048-sample chunks with an envelope transition",
            "voice_samples_per_s": voice_samples, "partial_samples_per_s": part, "ms_per_render": sms,
            "bound": "hbm (the output: 4 B per voice-sample)",
            "hbm_write_gbs": voice_samples * 4 / world / 1e9,
            "hbm_frac": voice_samples * 4 / world / 1e9 / peak_gbs_s,
            "tensor_tflops_executed": part * 2 * 2 * 3 / world / 1e12,
            "fp32_peak_tflops_measured": fp32_peak,
            "fp32_frac_official_8flop": part * 8 / world / 1e12 / fp32_peak,
            "note": "official count = 8 FLOP per partial-sample (rotate 6 + MAC 2, SURVEY §8d): the product form needs 2 FMA per "
                    "partial-sample and runs them on the tensor pipe, so this fraction of the FP32 peak exceeds 1; what bounds the render "
                    "now is the 4 bytes written per voice-sample (hbm_frac)",
            # pipe utilisation as ncu counts it (committed captures; constants, not measured here)
            "ncu": ncu_traffic("oscbank_umma_counters"),
            "mma_sync_variant": {
                "path": "CA_OSC_PATH=mma: every chunk on mma.sync m16n8k16 (3 x FP16), no tcgen05 kernel",
                "voice_samples_per_s": world * per_rank / (mms * 1e-3), "ms_per_render": mms,
                "ncu": ncu_traffic("oscbank_mma_counters"),
            },
            "fp32_pipe_variant": {
                "path": "CA_OSC_PATH=fp32: the same product on the FP32 pipe (FFMA2, 2 FMA per partial-sample)",
                "voice_samples_per_s": world * per_rank / (fms * 1e-3), "ms_per_render": fms,
                "fp32_frac_executed_4flop": world * per_rank * 64 / (fms * 1e-3) * 4 / world / 1e12 / fp32_peak,
                "fp32_frac_official_8flop": world * per_rank * 64 / (fms * 1e-3) * 8 / world / 1e12 / fp32_peak,
                "ncu": ncu_traffic("oscbank_fp32_counters"),
            },
            "round1_kernel": ncu_traffic("oscbank_kernel_counters"),
        }
        del sout

    # ---- BASELINE configs 1 and 5 (batched real FFT), points of the sweep (2^15 ... 2^17: the cluster kernels); the full sweep is
    # profiles/sweep_rfft.py ----
    if not args.skip_rfft:
        peak_gbs, _ = measured_peaks()
        rf = {}
        for n_fft, batch in ((1024, 4096), (1024, 262144), (4096, 65536), (16384, 16384), (1 << 15, 8192), (1 << 16, 4096), (1 << 17, 2048),
                             (1 << 18, 1024)):
            xr = torch.empty((batch, n_fft), dtype=torch.float32, device=dev).uniform_(-1, 1)
            yr = pkg.rfft_batched(ctx, xr)
            for _ in range(3):
                pkg.rfft_batched(ctx, xr, out=yr)
            sync_all()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 50 if batch * n_fft <= (1 << 22) else 5
            r0.record()
            for _ in range(reps):
                pkg.rfft_batched(ctx, xr, out=yr)
            r1.record()
            sync_all()
            rms = r0.elapsed_time(r1) / reps
            gbs = batch * (4 * n_fft + 8 * (n_fft // 2 + 1)) / (rms * 1e-3) / 1e9
            rf[f"n{n_fft}_b{batch}"] = {"ms": rms, "frames_per_s": batch / (rms * 1e-3), "algorithmic_gbs": gbs, "frac_hbm": gbs / peak_gbs}
            if batch * n_fft <= (1 << 22):
                # launch-bound case: the same call captured in a CUDA graph (20 launches per replay), as SURVEY §8d prescribes
                try:
                    graph = torch.cuda.CUDAGraph()
                    cap = torch.cuda.Stream(device=dev)
                    cap.wait_stream(torch.cuda.current_stream(dev))
                    with torch.cuda.stream(cap):
                        pkg.rfft_batched(ctx, xr, out=yr)
                        with torch.cuda.graph(graph, stream=cap):
                            for _ in range(20):
                                pkg.rfft_batched(ctx, xr, out=yr)
                    torch.cuda.current_stream(dev).wait_stream(cap)
                    graph.replay()
                    sync_all()
                    r0.record()
                    for _ in range(10):
                        graph.replay()
                    r1.record()
                    sync_all()
                    gms = r0.elapsed_time(r1) / 200
                    ggbs = batch * (4 * n_fft + 8 * (n_fft // 2 + 1)) / (gms * 1e-3) / 1e9
                    rf[f"n{n_fft}_b{batch}"].update(graph_ms=gms, graph_algorithmic_gbs=ggbs, graph_frac_hbm=ggbs / peak_gbs)
                    del graph
                except Exception as exc:  # capture is an extra: the event-timed number above stands on its own
                    rf[f"n{n_fft}_b{batch}"]["graph_error"] = str(exc)[:200]
            del xr, yr
        rf["note"] = ("per rank; n1024_b4096 is BASELINE config 1 (33.6 MB, L2 resident, launch latency included), the others "
                      "are points of config 5 (1 GiB of input each)")
        extra["rfft"] = rf

        # ---- the fused analysis at the reference's own default window: 0.1814 s / stride 0.09 s (source/rt.resynth.lib.cpp:945-946),
        # at 48 kHz 8708 samples zero padded to 16384 points, hop 4320; same streams, device resident ----
        dw_window, dw_hop = 8708, 4320
        an_dw = pkg.SqMagAnalyzer(ctx, dw_window, dw_hop, sample_rate=SR, min_db=MIN_DB, max_peaks=MAX_PEAKS)
        out_dw = an_dw.allocate(S, n_samples, dev)
        for _ in range(3):
            an_dw(x, out_dw)
        sync_all()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record()
        for _ in range(10):
            an_dw(x, out_dw)
        d1.record()
        sync_all()
        dms = d0.elapsed_time(d1) / 10
        dw_frames = S * an_dw.num_frames(n_samples)
        dw_bytes = dw_frames * (4 * dw_hop + 4 * (an_dw.n_fft // 2 + 1))
        extra["stft_default_window"] = {
            "workload": "reference defaults: window 8708 samples (0.1814 s), hop 4320 (0.09 s), 16384-point frames, Gaussian window, |X|^2 + peaks",
            "kernel": "stft_cta_kernel<8> (one CTA per frame, csrc/rfftw.cu)", "ms": dms, "frames_per_s": dw_frames / (dms * 1e-3),
            "algorithmic_gbs": dw_bytes / (dms * 1e-3) / 1e9, "frac_hbm": dw_bytes / (dms * 1e-3) / 1e9 / peak_gbs,
            "note": "per rank; bound by the 16384-point transform (shared-memory wavefronts and latency, DESIGN.md K2 'other FFT lengths'), not by HBM",
        }
        del out_dw, an_dw
        an._plan(an.window)  # the context holds one user plan: back to the headline analysis

    # ---- what the host can copy with all ranks at it: the ceiling of the host-buffer paths below ----
    ceiling = None
    if not args.skip_roundtrip:
        ceiling = measure_copy_ceiling(torch, dist, dev, world)
        extra["host_copy_ceiling"] = ceiling

    def copy_bound_ms(h2d_bytes, d2h_bytes):
        """time the copies alone need on this box with every rank copying: the slower direction at the rate measured for that
        direction alone (the two directions overlap; when they slow each other down the bound is optimistic)"""
        return 1e3 * max(h2d_bytes / (ceiling["h2d_gbs_per_rank"] * 1e9), d2h_bytes / (ceiling["d2h_gbs_per_rank"] * 1e9))

    # ---- BASELINE config 4 per rank: host audio -> analysis -> pitch pipeline -> resynthesis -> host stereo audio ----
    if not args.skip_roundtrip:
        rt_in = pkg.pinned_empty(x_host_np.shape)
        rt_in[:] = x_host_np
        rt_out = pkg.pinned_empty(x_host_np.shape + (2,))
        pkg.resynth_roundtrip(ctx, rt_in, out=rt_out)  # first call sizes the context's staging buffers
        sync_all()
        t0 = time.perf_counter()
        rt_reps = 2
        for _ in range(rt_reps):
            _, rt_stats = pkg.resynth_roundtrip(ctx, rt_in, out=rt_out)
        rt_s = (time.perf_counter() - t0) / rt_reps
        if world > 1:
            t = torch.tensor([rt_s], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            rt_s = float(t.item())
        extra["roundtrip"] = {
            "workload": f"C4 share: {S} streams x {SECONDS:g} s per GPU, pinned host audio in, pinned host stereo audio out",
            "output_samples_per_s": world * S * n_samples / rt_s, "ms_per_call": 1e3 * rt_s,
            "notes_started": int(rt_stats[:, 0].sum()), "h2d_bytes": int(rt_in.nbytes), "d2h_bytes": int(rt_out.nbytes),
            "truncated_frames": pkg.truncated_frames(ctx),
            "copy_bound_ms": copy_bound_ms(int(rt_in.nbytes), int(rt_out.nbytes)),
            "frac_of_copy_ceiling": copy_bound_ms(int(rt_in.nbytes), int(rt_out.nbytes)) / (1e3 * rt_s),
            "note": "analysis (in FP64 on this path: the resynthesised phase integrates the peak frequencies), pitch pipeline (three "
                    "warps per stream pipelined over frames) and rendering all run on the device, as a pipeline over 16 time chunks: the "
                    "audio of the early chunks goes back while the tracker walks the later ones; only the audio crosses PCIe",
        }
        # ---- the same call fed with 16-bit PCM (the job's WAV files): half the input bytes ----
        rt_pcm = pkg.pinned_empty(x_host_np.shape, dtype=np.int16)
        rt_pcm[:] = np.round(x_host_np * 32767.0).astype(np.int16)
        pkg.resynth_roundtrip(ctx, rt_pcm, out=rt_out)
        sync_all()
        t0 = time.perf_counter()
        for _ in range(rt_reps):
            pkg.resynth_roundtrip(ctx, rt_pcm, out=rt_out)
        pcm_s = (time.perf_counter() - t0) / rt_reps
        if world > 1:
            t = torch.tensor([pcm_s], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            pcm_s = float(t.item())
        extra["roundtrip_pcm16"] = {
            "workload": f"C4 share with 16-bit PCM input: {S} streams x {SECONDS:g} s per GPU",
            "output_samples_per_s": world * S * n_samples / pcm_s, "ms_per_call": 1e3 * pcm_s, "h2d_bytes": int(rt_pcm.nbytes),
            "d2h_bytes": int(rt_out.nbytes), "copy_bound_ms": copy_bound_ms(int(rt_pcm.nbytes), int(rt_out.nbytes)),
            "frac_of_copy_ceiling": copy_bound_ms(int(rt_pcm.nbytes), int(rt_out.nbytes)) / (1e3 * pcm_s),
        }
        del rt_pcm
        # ---- FFT vocoder (source/rt.resynth.lib.vocoder.cpp) at the reference defaults, same streams as the modulator ----
        vc_car = pkg.pinned_empty(x_host_np.shape)
        vc_car[:] = np.random.default_rng(77 + rank).uniform(-0.5, 0.5, x_host_np.shape).astype(np.float32)
        vc_out = pkg.pinned_empty(x_host_np.shape)
        pkg.vocoder(ctx, rt_in, vc_car, out=vc_out)
        sync_all()
        t0 = time.perf_counter()
        for _ in range(rt_reps):
            pkg.vocoder(ctx, rt_in, vc_car, out=vc_out)
        vc_s = (time.perf_counter() - t0) / rt_reps
        if world > 1:
            t = torch.tensor([vc_s], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            vc_s = float(t.item())
        extra["vocoder"] = {
            "workload": f"{S} streams x {SECONDS:g} s per GPU, window 0.10 s, stride 0.005 s (241 samples, 512-point carrier FFT), 5 bands; pinned host audio in and out",
            "output_samples_per_s": world * S * n_samples / vc_s, "ms_per_call": 1e3 * vc_s,
            "h2d_bytes": int(2 * rt_in.nbytes), "d2h_bytes": int(vc_out.nbytes),
            "copy_bound_ms": copy_bound_ms(int(2 * rt_in.nbytes), int(vc_out.nbytes)),
            "frac_of_copy_ceiling": copy_bound_ms(int(2 * rt_in.nbytes), int(vc_out.nbytes)) / (1e3 * vc_s),
        }
        del rt_in, rt_out, vc_car, vc_out

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- CPU baseline: the oracle port on this box's host cores, bounded sample (rank 0, N = 1 only) ----
    cpu = None
    if world == 1 and not args.skip_cpu:
        orc = graft.load_oracle()
        cores = orc.hardware_threads()
        rate_1t, dt1 = cpu_analysis_rate(orc, wl, 1, 8, 60.0)
        n_mt = cores * 24
        rate_mt, dtm = cpu_analysis_rate(orc, wl, cores, n_mt, 60.0)
        cpu = {"value": rate_mt, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": f"{n_mt} streams x 60 s of config 2 on {cores} threads ({dtm:.1f} s wall); single thread: "
                         f"{rate_1t:.0f} frames/s on 8 streams x 60 s ({dt1:.1f} s)",
               "value_1_thread": rate_1t}

    peak, peak_src = measured_peaks()
    fp32_peak_stft = ctx.measure_fp32_peak()
    per_frame = 4 * HOP + 4 * (NFFT // 2 + 1) + 4 + 12 * 16
    achieved = rows * per_frame / (kernel_ms * 1e-3) / 1e9
    noreuse = rows * (4 * WINDOW + 4 * (NFFT // 2 + 1)) / (kernel_ms * 1e-3) / 1e9
    e2e_bound = None
    if ceiling:
        bound_s = max(h2d / (ceiling["h2d_gbs_per_rank"] * 1e9), d2h / (ceiling["d2h_gbs_per_rank"] * 1e9))
        e2e_bound = {"copy_bound_ms_per_step": 1e3 * bound_s, "frac_of_copy_ceiling": bound_s / (e2e_s / e2e_steps)}
    line = {
        "metric": "stft_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(world, S),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic("stft4096p_kernel"), "peak_source": peak_src,
                     "algorithmic_bytes_per_frame": per_frame, "kernel": "stft4096p_kernel", "kernel_ms": kernel_ms,
                     "achieved_no_reuse_convention_gbs": noreuse, "frac_no_reuse_convention": noreuse / peak,
                     # the kernel sits on the ridge (13 FLOP/B against a machine balance of 11.5): FP32 view next to the HBM view
                     "fp32_pipe_frac": rows * FLOP_PER_FRAME / (kernel_ms * 1e-3) / 1e12 / fp32_peak_stft,
                     "fp32_peak_tflops_measured": fp32_peak_stft, "algorithmic_flop_per_frame": FLOP_PER_FRAME,
                     "ncu": ncu_traffic("stft4096p_kernel_counters")},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "outputs": "peak lists (bin, freq, dB, count) copied to the host; the spectrum, which the device-timed step also "
                           "writes, stays in HBM here (extra.e2e_with_spectrum copies it too)",
                "copy_ceiling": e2e_bound},
        "gpu_launches": launches_timed, "clocks": clocks, "extra": dict(extra, final_gather_ms=gather_ms, e2e_with_spectrum=e2e_sq),
    }
    _RESULT_LINE.append(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams-per-gpu", type=int, default=32)
    ap.add_argument("--voices", type=int, default=1024)
    ap.add_argument("--synth-seconds", type=float, default=10.0)
    ap.add_argument("--skip-synth", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-roundtrip", action="store_true")
    ap.add_argument("--skip-rfft", action="store_true")
    ap.add_argument("--skip-e2e-spectrum", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON result: libraries that write to file descriptor 1 on their own (NCCL prints
    # its version banner there) are sent to stderr for the duration of the run
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        rc = run_reference(args) if args.impl == "reference" else run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if _RESULT_LINE:
        print(_RESULT_LINE[-1], flush=True)
    return rc


if __name__ == "__main__":
    sys.exit(main())
