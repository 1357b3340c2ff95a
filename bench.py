#!/usr/bin/env python
"""bench.py -- LERC encode+decode throughput of lerc_b200 on B200 (driver contract, task section 4).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c4|c5] [--gather]

One "step" = lerc_encode + lerc_decode of one synthetic raster of BASELINE.json configs[1]
(4096x4096 float32, 1 band, maxZError = 0.01; generator tests/cases.py:c2_raster).
  value   Gpixels/s with the raster / blob / output resident in HBM (device pointers through the C ABI,
          timed with CUDA events on the stream the kernels run on)
  e2e     the same calls with pinned HOST buffers: H2D of the raster, D2H of the blob, H2D of the blob,
          D2H of the pixels are inside the timed region
  roofline  dominant kernel: algorithmic bytes / its CUDA-event duration vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the unmodified reference (oracle/_ref/libLerc_ref.so) on the host cores, bounded sample
With N > 1 (torchrun) every rank codes its own raster (independent objects, weak scaling, no data-path
collective); rank 0 reports units of all ranks / max-over-ranks time.
The default line also carries the other BASELINE configs as sub-records, measured the same way under a time budget:
  "c5"  configs[4]: every rank codes its strip of the tiled 65536^2 raster (lerc_b200_encodeTiles / decodeTiles), then the
        per-tile streams of all ranks are gathered over NCCL into one container (lerc_b200/tiles.py): kernel-only and
        kernel+gather aggregate Gpixels/s, gather ms, bytes gathered, fraction of the NVLink peer bandwidth
  "c4"  configs[3]: 8192^2 x 3 uint8 lossless (Huffman path), one raster per rank
  "c3"  configs[2]: 16384^2 float32 x 4 bands at maxZError 0.001 (N = 1 only: 13 GB of buffers)
(--no-sub skips them; --workload c3|c4|c5 runs one of them as the main line.)
--impl reference times the reference's CPU implementation on the host cores and prints the same JSON line.
--workload c5 (BASELINE configs[4]): every rank holds an 8192 x 65536 strip of the 65536^2 float32 raster (8192 tiles of
256 x 256; 8 ranks = the whole raster) and codes it with ONE lerc_b200_encodeTiles + ONE lerc_b200_decodeTiles call per
step; --gather adds the all-gather of the per-tile streams (lerc_b200/tiles.py) as a separately reported time.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (rows, cols, nDepth, dtype code, maxZErr, description)
    "c2": (4096, 4096, 1, 6, 0.01, "4096x4096 float32 1-band encode+decode maxZError=0.01"),
    "c3": (16384, 16384, 1, 6, 0.001, "16384x16384 float32 4-band (nBands=4) encode+decode maxZError=0.001"),
    "c4": (8192, 8192, 3, 1, 0.0, "8192x8192 nDepth=3 uint8 lossless (Huffman path)"),
    # not a BASELINE config: configs[1]'s raster at maxZError 0 = the lossless float (FPL) codec, for profiling that path
    "c2l": (4096, 4096, 1, 6, 0.0, "4096x4096 float32 1-band encode+decode maxZError=0 (lossless float codec)"),
    # not a BASELINE config either: configs[1]'s raster under a california-shaped validity mask (ragged coast, one third invalid): the masked paths
    "c2m": (4096, 4096, 1, 6, 0.01, "4096x4096 float32 1-band, coast-shaped validity mask (one third invalid), encode+decode maxZError=0.01"),
    # per-rank strip of the 65536^2 raster; rows can be lowered with --strip-rows for a quick run
    "c5": (8192, 65536, 1, 6, 0.01, "65536x65536 float32 as 256x256 tiles, 8192-row strip (8192 tiles) per GPU, encodeTiles+decodeTiles maxZError=0.01"),
}
TILE = 256
BANDS = {"c3": 4}                      # bands per call (default 1)
NVLINK_PEER_GBS = 770.0                # measured peer copy per direction (B200_PROFILING.md)


METRIC = {"c2": "Gpixels/s encode+decode float32 @ maxZError=0.01; achieved HBM GB/s vs peak", "c3": "Gpixels/s (pixels x bands) encode+decode float32 4 bands @ maxZError=0.001", "c2l": "Gpixels/s encode+decode float32 lossless", "c2m": "Gpixels/s encode+decode float32 @ maxZError=0.01 with a validity mask", "c4": "Gpixels/s encode+decode uint8 nDepth=3 lossless",
          "c5": "Gpixels/s encode+decode float32 @ maxZError=0.01, 256x256 tiles (one blob per tile); achieved HBM GB/s vs peak"}


def make_raster(workload, seed):
    from cases import c2_raster, c4_raster
    rows, cols, depth, dt, mz, _ = WORKLOADS[workload]
    if workload in ("c2", "c2l"):
        return c2_raster(rows, cols, seed=seed, phase=0.1 * (seed % 7))
    return c4_raster(rows, cols, seed=seed)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region"""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = []
        for i, n in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows):
                reasons.append(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
def cpu_reference_run(workload, seconds_budget=20.0, threads=None):
    """The reference's own CPU implementation on the host cores (bounded sample)."""
    from lercapi import oracle_lib, ref_lib
    lib, kind = ref_lib(), "reference"
    if lib is None:
        lib, kind = oracle_lib(), "port"
    assert lib is not None, "neither oracle/_ref/libLerc_ref.so nor oracle/_build/liblerc_oracle.so present"
    rows, cols, depth, dt, mz, desc = WORKLOADS[workload]
    # bounded sample: a horizontal strip of the workload raster per thread (the library is single-threaded and
    # re-entrant: one independent call per core)
    try:                                               # (the GPU arm may have bound this process to its GPU's NUMA node: the CPU arm uses every core)
        os.sched_setaffinity(0, range(os.cpu_count() or 1))
    except OSError:
        pass
    threads = threads or os.cpu_count() or 1
    strip_rows = min(rows, 1024 if workload in ("c2", "c2l") else 512)
    if workload == "c5":                               # one 256 x 256 tile per call, as the reference's tile callers do
        from cases import c2_raster
        strip_rows, cols = TILE, TILE
        img = c2_raster(TILE, TILE, seed=1234)
    else:
        img = make_raster(workload, 1234)[:strip_rows]
    n_px = strip_rows * cols

    def one(out):
        t0 = time.perf_counter()
        st, blob, _ = lib.encode(img, mz, n_depth=depth)
        assert st == 0
        st, dec, _ = lib.decode(blob)
        assert st == 0
        out.append(time.perf_counter() - t0)

    one([])   # warm-up (page in the library, touch the buffers)
    single = []
    one(single)
    reps = max(1, int(seconds_budget / max(single[0], 1e-3) / 2))
    reps = min(reps, 8 if workload != "c5" else 2000)
    t0 = time.perf_counter()
    res = []
    th = [threading.Thread(target=lambda: [one(res) for _ in range(reps)]) for _ in range(threads)]
    [t.start() for t in th]
    [t.join() for t in th]
    wall = time.perf_counter() - t0
    gpx_all = threads * reps * n_px / wall / 1e9
    gpx_one = n_px / single[0] / 1e9
    return {"value": gpx_all, "unit": "Gpixels/s", "cores": threads, "kind": kind, "single_thread_value": gpx_one,
            "sample": f"{strip_rows}x{cols} strip of the workload raster, {reps} encode+decode calls on each of {threads} threads ({wall:.1f} s wall)"}, wall / (threads * reps) * 1e3



# ---------------------------------------------------------------------------------------------------
def device_c2_strip(torch, rows, cols, row0, seed):
    """tests/cases.py:c2_raster evaluated on the device for rows [row0, row0 + rows) of a cols-wide raster (fp64 math, cast to
    float32; the noise comes from torch's generator, so only the distribution -- not the values -- matches the numpy one)."""
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    out = torch.empty((rows, cols), dtype=torch.float32, device="cuda")
    xx = torch.arange(cols, dtype=torch.float64, device="cuda")[None, :]
    for r0 in range(0, rows, 512):
        r1 = min(rows, r0 + 512)
        yy = torch.arange(row0 + r0, row0 + r1, dtype=torch.float64, device="cuda")[:, None]
        z = 1000 + 300 * torch.sin(xx / 97) * torch.cos(yy / 131) + 50 * torch.sin(xx / 13 + yy / 17)
        z += torch.randn((r1 - r0, cols), dtype=torch.float64, device="cuda", generator=g) * 0.5
        out[r0:r1] = z.to(torch.float32)
    return out


def device_c4_raster(torch, rows, cols, seed):
    """tests/cases.py:c4_raster's formula on the device: three smooth sinusoid fields + N(0, 2), clipped to uint8, depth-interleaved"""
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    out = torch.empty((rows, cols, 3), dtype=torch.uint8, device="cuda")
    xx = torch.arange(cols, dtype=torch.float32, device="cuda")[None, :]
    for r0 in range(0, rows, 1024):
        r1 = min(rows, r0 + 1024)
        yy = torch.arange(r0, r1, dtype=torch.float32, device="cuda")[:, None]
        ch = [128 + 100 * torch.sin(xx / 53) * torch.cos(yy / 71), 128 + 90 * torch.sin(xx / 31 + 1) * torch.cos(yy / 47), 100 + 80 * torch.cos(xx / 91) * torch.sin(yy / 23)]
        for c in range(3):
            z = ch[c] + torch.randn((r1 - r0, cols), dtype=torch.float32, device="cuda", generator=g) * 2.0
            out[r0:r1, :, c] = z.round().clamp(0, 255).to(torch.uint8)
    return out


def bench_raster_sub(lib, workload, steps, rank, world, peak_gbs):
    """One of the whole-raster BASELINE configs as a sub-record of the default line: device-generated raster, device pointers through
    the C ABI, CUDA events around `steps` encode+decode steps (max over ranks), dominant kernel from an instrumented pass."""
    import torch
    import torch.distributed as dist
    import lerc_b200
    rows, cols, depth, dt, mz, desc = WORKLOADS[workload]
    bands = BANDS.get(workload, 1)
    enc, dec = lib.f["encode"], lib.f["decode"]
    if workload == "c4":
        d_img = device_c4_raster(torch, rows, cols, 7 + rank)
    else:
        d_img = torch.stack([device_c2_strip(torch, rows, cols, 0, 1234 + b + 16 * rank) for b in range(bands)])
    d_mask = d_dmask = None
    if workload == "c2m":                                                      # valid to the right of a ragged coast line
        yy = torch.arange(rows, device="cuda", dtype=torch.float32)[:, None]
        xx = torch.arange(cols, device="cuda", dtype=torch.float32)[None, :]
        coast = 0.34 * cols + 0.12 * cols * torch.sin(yy / 300.0) + 120 * torch.sin(yy / 37.0) + 25 * torch.sin(yy / 5.0)
        d_mask = (xx > coast).to(torch.uint8).contiguous()
        d_dmask = torch.empty_like(d_mask)
    n_masks = 1 if d_mask is not None else 0
    p_mask, p_dmask = (d_mask.data_ptr(), d_dmask.data_ptr()) if n_masks else (None, None)
    raw_bytes = d_img.numel() * d_img.element_size()
    cap = min(raw_bytes + raw_bytes // 8 + 4096, 0xF0000000)                  # (outBufferSize is a 32-bit count)
    d_blob = torch.empty(cap, dtype=torch.uint8, device="cuda")
    d_dec = torch.empty_like(d_img)
    n_written = C.c_uint(0)
    stream = torch.cuda.current_stream()
    lerc_b200.set_stream(stream.cuda_stream, True)
    tight = [cap]

    def step():
        st = enc(d_img.data_ptr(), dt, depth, cols, rows, bands, n_masks, p_mask, mz, d_blob.data_ptr(), tight[0], C.addressof(n_written))
        assert st == 0, f"{workload}: lerc_encode status {st}"
        st = dec(d_blob.data_ptr(), n_written.value, n_masks, p_dmask, depth, cols, rows, bands, dt, d_dec.data_ptr())
        assert st == 0, f"{workload}: lerc_decode status {st}"
        return n_written.value

    blob_bytes = step()
    tight[0] = min(cap, int(blob_bytes * 1.02) + 65536)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = lerc_b200.stats()[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = lerc_b200.stats()[0] - launches0
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    if n_masks:
        assert bool((d_dmask == d_mask).all()), f"{workload}: mask round trip"
        err = float(((d_dec.double() - d_img.double()).abs() * d_mask.double()).max().item())
    else:
        err = float((d_dec.double() - d_img.double()).abs().max().item())
    assert err <= max(mz, 0.5 if dt < 6 else 0) * 1.1 + 1e-12, f"{workload}: round trip error {err} exceeds maxZError {mz}"
    lerc_b200.kernel_times(reset=True)
    lerc_b200.profile(True)
    step()
    torch.cuda.synchronize()
    lerc_b200.profile(False)
    kt = lerc_b200.kernel_times(reset=True)
    top_name, (top_cnt, top_ms) = max(kt.items(), key=lambda kv: kv[1][1])
    n_px = rows * cols * bands
    algo = 2 * (raw_bytes + blob_bytes)                                        # both directions: raster + blob moved once each
    rec = {"workload": desc, "value": world * n_px / (ms * 1e-3) / 1e9, "unit": "Gpixels/s", "ms_per_step": ms, "steps": steps, "n_gpus": world,
           "blob_bytes": blob_bytes, "compression_ratio": raw_bytes / blob_bytes, "gpu_launches_per_step": launches / steps,
           "step_frac": algo / (ms * 1e-3) / 1e9 / peak_gbs, "dominant_kernel": top_name, "dominant_kernel_ms_per_step": top_ms,
           "kernels": {n: round(m, 4) for n, (c, m) in sorted(kt.items(), key=lambda kv: -kv[1][1])[:6]},
           "data": "synthetic, generated on the device (bench.py)", "sharding": "one raster per rank" if world > 1 else "single GPU"}
    del d_img, d_blob, d_dec
    torch.cuda.empty_cache()
    return rec


def main_tiles(args, lib, rank, world, local_rank, warm, peak_gbs, peak_src, sub=False):
    import torch
    import torch.distributed as dist
    import lerc_b200
    from lercapi import tiles_api
    from lerc_b200.tiles import gather_container

    rows, cols, depth, dt, mz, desc = WORKLOADS["c5"]
    if args.strip_rows:
        rows = args.strip_rows
    assert rows % TILE == 0 and cols % TILE == 0
    api = tiles_api(lib)
    n_px = rows * cols
    n_tiles = (rows // TILE) * (cols // TILE)
    raw_bytes = n_px * 4
    d_img = device_c2_strip(torch, rows, cols, rank * rows, 1234 + rank)
    cap = int(api.lerc_b200_tilesMaxBytes(dt, cols, rows, TILE, TILE))
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    d_off = torch.zeros(n_tiles + 1, dtype=torch.int64, device="cuda")
    d_dec = torch.empty_like(d_img)
    n_written = C.c_ulonglong(0)
    stream = torch.cuda.current_stream()
    lerc_b200.set_stream(stream.cuda_stream, True)

    def step_device():
        st = api.lerc_b200_encodeTiles(d_img.data_ptr(), dt, cols, rows, TILE, TILE, mz, d_out.data_ptr(), cap, d_off.data_ptr(), C.addressof(n_written))
        assert st == 0, f"lerc_b200_encodeTiles status {st}"
        st = api.lerc_b200_decodeTiles(d_out.data_ptr(), n_written.value, d_off.data_ptr(), dt, cols, rows, TILE, TILE, d_dec.data_ptr())
        assert st == 0, f"lerc_b200_decodeTiles status {st}"
        return n_written.value

    s0 = lerc_b200.stats()
    for _ in range(warm):
        blob_bytes = step_device()
    s1 = lerc_b200.stats()
    fast_enc, fast_dec = (s1[3] - s0[3]) // warm, (s1[4] - s0[4]) // warm
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_s = time.perf_counter()
    while len(sampler.rows) < 2 and time.perf_counter() - t_s < 3.0:
        step_device()
    torch.cuda.synchronize()
    launches0 = lerc_b200.stats()[0]
    ev_all = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    ev_all[0].record(stream)
    for _ in range(args.steps):
        step_device()
    ev_all[1].record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = lerc_b200.stats()[0] - launches0
    dev_ms = ev_all[0].elapsed_time(ev_all[1])
    clocks = sampler.stop()
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = world * n_px / (ms_per_step * 1e-3) / 1e9
    err = float((d_dec.double() - d_img.double()).abs().max().item())
    assert err <= mz * 1.1, f"round trip error {err} exceeds maxZError {mz}"

    # ---- per-kernel roofline (instrumented pass)
    lerc_b200.kernel_times(reset=True)
    lerc_b200.profile(True)
    PROF = 3
    for _ in range(PROF):
        step_device()
    torch.cuda.synchronize()
    lerc_b200.profile(False)
    kt = lerc_b200.kernel_times(reset=True)
    top_name, (top_cnt, top_ms) = max(kt.items(), key=lambda kv: kv[1][1])
    per_launch_ms = top_ms / top_cnt
    algo_bytes = raw_bytes + blob_bytes          # both dominant kernels (fused tile encoder, tile block decoder) move the raster and the blobs once
    achieved = algo_bytes / (per_launch_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": top_name, "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs, "traffic": None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes, "algorithmic_bytes_are": "strip read/written once + its blobs written/read once",
                "kernel_ms_per_launch": per_launch_ms, "kernel_share_of_step": top_ms / PROF / ms_per_step,
                "step_frac": (2 * algo_bytes) / (ms_per_step * 1e-3) / 1e9 / peak_gbs,
                "kernels": {n: {"launches_per_step": c / PROF, "ms_per_step": m / PROF} for n, (c, m) in sorted(kt.items(), key=lambda kv: -kv[1][1])[:10]}}

    # ---- the gather of the per-tile streams of all ranks into one container (the only exchange step of the path; reported beside
    # the codec time): sizes all-gather + prefix sum + every rank's streams straight into their place (lerc_b200/tiles.py)
    gather = None
    if args.gather or sub:
        torch.cuda.synchronize()
        total_guess = int(blob_bytes * world * 1.1) + (1 << 20)
        d_cont = torch.empty(total_guess, dtype=torch.uint8, device="cuda")
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        container, offsets = gather_container(d_out[:blob_bytes], d_off, n_tiles * world, out=d_cont)      # warm-up
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        g0.record(stream)
        for _ in range(3):
            container, offsets = gather_container(d_out[:blob_bytes], d_off, n_tiles * world, out=d_cont)
        g1.record(stream)
        torch.cuda.synchronize()
        tg = torch.tensor([g0.elapsed_time(g1) / 3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        g_ms = float(tg.item())
        cont_bytes = int(container.numel())
        ingress = cont_bytes - blob_bytes                                       # bytes that enter this GPU over NVLink
        gather = {"ms": g_ms, "container_bytes": cont_bytes, "tiles": int(offsets.numel() - 1), "comm_nranks": world,
                  "ingress_bytes_per_gpu": ingress, "ingress_gbs_per_gpu": ingress / (g_ms * 1e-3) / 1e9 if g_ms > 0 else None,
                  "frac_of_nvlink_peer": (ingress / (g_ms * 1e-3) / 1e9 / NVLINK_PEER_GBS) if (g_ms > 0 and world > 1) else None,
                  "nvlink_peer_gbs": NVLINK_PEER_GBS,
                  "kernel_plus_gather_gpixels": world * n_px / ((ms_per_step + g_ms) * 1e-3) / 1e9,
                  "limiter": "gather (NVLink ingress)" if (world > 1 and g_ms > ms_per_step) else "codec kernels"}
        del container, d_cont
    if sub:
        rec = {"workload": desc if rows == 8192 else desc.replace("8192-row strip (8192 tiles)", f"{rows}-row strip ({n_tiles} tiles)"),
               "value": value, "unit": "Gpixels/s", "per_gpu_gpixels": value / world, "ms_per_step": ms_per_step, "steps": args.steps, "n_gpus": world, "tiles_per_gpu": n_tiles,
               "blob_bytes_per_gpu": blob_bytes, "fused_encodes_per_step": fast_enc, "batch_decodes_per_step": fast_dec, "gpu_launches_per_step": launches / args.steps,
               "step_frac": roofline["step_frac"], "dominant_kernel": roofline["kernel"], "dominant_kernel_frac": roofline["frac"],
               "kernels": {n: round(v["ms_per_step"], 4) for n, v in roofline["kernels"].items()}, "gather": gather,
               "sharding": "contiguous tile rows per rank; codec without collective; streams gathered over NCCL", "data": "synthetic, generated on the device"}
        lerc_b200.set_stream(0, False)
        del d_img, d_out, d_dec
        torch.cuda.empty_cache()
        return rec

    # ---- end to end: pinned host buffers through the same calls
    lerc_b200.set_stream(0, False)
    e2e_rows = min(rows, 2048)                                      # bounded: a 2048-row strip (2048 tiles, 537 MB) per call
    e2e_px = e2e_rows * cols
    e2e_tiles = (e2e_rows // TILE) * (cols // TILE)
    h_in = d_img[:e2e_rows].cpu().pin_memory()
    e2e_cap = int(api.lerc_b200_tilesMaxBytes(dt, cols, e2e_rows, TILE, TILE))
    h_blob = torch.empty(e2e_cap, dtype=torch.uint8).pin_memory()
    h_off = torch.zeros(e2e_tiles + 1, dtype=torch.int64)
    h_out = torch.empty_like(h_in).pin_memory()

    def step_host():
        st = api.lerc_b200_encodeTiles(h_in.data_ptr(), dt, cols, e2e_rows, TILE, TILE, mz, h_blob.data_ptr(), e2e_cap, h_off.data_ptr(), C.addressof(n_written))
        assert st == 0
        st = api.lerc_b200_decodeTiles(h_blob.data_ptr(), n_written.value, h_off.data_ptr(), dt, cols, e2e_rows, TILE, TILE, h_out.data_ptr())
        assert st == 0
        return n_written.value

    step_host()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    E2E = 3
    t0 = time.perf_counter()
    for _ in range(E2E):
        nb = step_host()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / E2E
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e = {"value": world * e2e_px / float(t.item()) / 1e9, "unit": "Gpixels/s", "ms_per_step": float(t.item()) * 1e3,
           "h2d_bytes_per_step": e2e_px * 4 + nb, "d2h_bytes_per_step": nb + e2e_px * 4, "sample": f"{e2e_rows}x{cols} strip ({e2e_tiles} tiles) per call",
           "timer": "host wall clock around the synchronous C-API calls"}
    assert float((h_out.double() - h_in.double()).abs().max().item()) <= mz * 1.1

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and args.gpus == 1:
            cpu, _ = cpu_reference_run("c5", seconds_budget=15.0)
        line = {"metric": METRIC["c5"], "value": value, "unit": "Gpixels/s", "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": desc if not args.strip_rows else desc.replace("8192-row strip (8192 tiles)", f"{rows}-row strip ({n_tiles} tiles)"),
                           "generator": "bench.py:device_c2_strip (tests/cases.py:c2_raster formula on the device)", "tiles_per_gpu": n_tiles,
                           "blob_bytes": blob_bytes, "compression_ratio": raw_bytes / blob_bytes, "fused_encodes_per_step": fast_enc, "batch_decodes_per_step": fast_dec,
                           "l2": f"strip {raw_bytes // 2**20} MiB + blobs {blob_bytes // 2**20} MiB per direction >> 126 MB L2",
                           "sharding": "contiguous tile rows per rank, no data-path collective in the codec" + ("; stream all-gather timed separately" if args.gather else "")},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "gather": gather}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------
def bind_to_gpu_numa(local_rank):
    """pins this process to the CPUs nearest to its GPU (NVML's affinity mask), so that the pinned host buffers of the end-to-end
    leg are allocated on the GPU's NUMA node; returns the number of CPUs in the mask or None"""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local_rank]) if vis and vis.split(",")[local_rank].strip().isdigit() else local_rank
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [w * 64 + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", action="store_true", help="c5: also time the all-gather of the per-tile streams")
    ap.add_argument("--strip-rows", type=int, default=0, help="c5: rows of the per-rank strip (multiple of 256; default 8192)")
    ap.add_argument("--no-sub", action="store_true", help="default workload: skip the c3 / c4 / c5 sub-records")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    rows, cols, depth, dt, mz, desc = WORKLOADS[args.workload]
    n_px = rows * cols
    peak_gbs, peak_src = peaks()
    warm = max(args.warmup, 3)

    if args.impl == "reference":
        if rank != 0:
            return
        base, ms = cpu_reference_run(args.workload, seconds_budget=max(10.0, 2.0 * args.steps))
        line = {"impl": "reference", "metric": METRIC[args.workload], "value": base["value"], "unit": "Gpixels/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64" if dt >= 6 else "u8", "data": "synthetic", "config": {"workload": desc, "sample": base["sample"]},
                "cpu_baseline": base, "e2e": {"value": base["value"], "unit": "Gpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    import lerc_b200
    from lercapi import DT_NP, product_lib

    assert torch.cuda.is_available(), "bench.py needs a CUDA device: lerc_b200 has no CPU fallback"
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa(local_rank)                               # (before any pinned allocation: first touch puts staging next to the GPU)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = product_lib()
    assert lib is not None, "lerc_b200/libLerc.so.4 missing: run python __graft_entry__.py"
    if args.workload == "c5":
        return main_tiles(args, lib, rank, world, local_rank, warm, peak_gbs, peak_src)
    if args.workload in ("c3", "c4", "c2m"):                                   # the big rasters: device-generated, one buffer set (far larger than L2)
        rec = bench_raster_sub(lib, args.workload, max(3, min(args.steps, 10)), rank, world, peak_gbs)
        if rank == 0:
            print(json.dumps({"metric": METRIC[args.workload], "value": rec["value"], "unit": "Gpixels/s", "n_gpus": world, "steps": rec["steps"], "warmup": 3,
                              "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64" if dt >= 6 else "u8",
                              "data": "synthetic", "config": {"workload": desc}, "e2e": None, "gpu_launches": int(rec["gpu_launches_per_step"] * rec["steps"]), "record": rec}))
        if world > 1:
            dist.destroy_process_group()
        return
    enc, dec = lib.f["encode"], lib.f["decode"]
    ts = np.dtype(DT_NP[dt]).itemsize
    raw_bytes = n_px * depth * ts

    # rotating buffer sets: successive steps touch different inputs/outputs, > 126 MB L2 between reuses
    NBUF = 4
    host_imgs = [make_raster(args.workload, 1234 + rank * 16 + i) for i in range(NBUF)]
    d_imgs = [torch.from_numpy(h).cuda() for h in host_imgs]
    cap = raw_bytes + raw_bytes // 8 + 4096
    d_blobs = [torch.empty(cap, dtype=torch.uint8, device="cuda") for _ in range(NBUF)]
    d_decs = [torch.empty_like(d) for d in d_imgs]
    n_written = C.c_uint(0)
    stream = torch.cuda.current_stream()
    lerc_b200.set_stream(stream.cuda_stream, True)

    # lerc_encode zero-fills the WHOLE output buffer like the reference (Lerc.cpp:374), so the timed calls pass a buffer
    # sized like a caller that asked lerc_computeCompressedSize first (blob size + 2 % slack), not the raw-size bound
    tight = [cap]

    def step_device(i):
        k = i % NBUF
        st = enc(d_imgs[k].data_ptr(), dt, depth, cols, rows, 1, 0, None, mz, d_blobs[k].data_ptr(), tight[0], C.addressof(n_written))
        assert st == 0, f"lerc_encode status {st}"
        nb = n_written.value
        st = dec(d_blobs[k].data_ptr(), nb, 0, None, depth, cols, rows, 1, dt, d_decs[k].data_ptr())
        assert st == 0, f"lerc_decode status {st}"
        return nb

    # ---- device-resident timing ------------------------------------------------------------------
    for i in range(warm):
        blob_bytes = step_device(i)
    tight[0] = min(cap, int(blob_bytes * 1.02) + 65536)
    for i in range(warm):
        blob_bytes = step_device(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_s = time.perf_counter()
    while len(sampler.rows) < 2 and time.perf_counter() - t_s < 3.0:      # nvidia-smi needs a moment to start; keep the GPU under the same load meanwhile
        step_device(0)
    torch.cuda.synchronize()
    launches0 = lerc_b200.stats()[0]
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev_all = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    ev_all[0].record(stream)
    for i in range(args.steps):
        ev[i][0].record(stream)
        step_device(warm + i)
        ev[i][1].record(stream)
    ev_all[1].record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = lerc_b200.stats()[0] - launches0
    t_s = time.perf_counter()
    n_before = len(sampler.rows)
    while len(sampler.rows) < n_before + 3 and time.perf_counter() - t_s < 1.0:   # the timed region is only milliseconds long: keep sampling under the same load
        step_device(0)
    torch.cuda.synchronize()
    dev_ms = ev_all[0].elapsed_time(ev_all[1])          # one bracket around exactly K steps (host gaps between the calls included)
    if os.environ.get("BENCH_DEBUG"):
        print("per-step ms:", [round(a.elapsed_time(b), 3) for a, b in ev], "stats", lerc_b200.stats(), file=sys.stderr)
    clocks = sampler.stop()
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t.item())
    ms_per_step = dev_ms_max / args.steps
    value = world * n_px / (ms_per_step * 1e-3) / 1e9

    # correctness guard inside the bench: the decoded raster respects maxZError (reference slack 1.1x, Lerc.cpp:1137)
    k = (warm + args.steps - 1) % NBUF
    err = float((d_decs[k].double() - d_imgs[k].double()).abs().max().item())
    assert err <= max(mz, 0.5 if dt < 6 else 0) * 1.1 + 1e-12, f"round trip error {err} exceeds maxZError {mz}"

    # ---- per-kernel roofline (separate instrumented pass, CUDA events around every launch) ------------
    lerc_b200.kernel_times(reset=True)
    lerc_b200.profile(True)
    PROF = max(3, min(args.steps, 10))
    for i in range(PROF):
        step_device(i)
    torch.cuda.synchronize()
    lerc_b200.profile(False)
    kt = lerc_b200.kernel_times(reset=True)
    total_k = sum(v[1] for v in kt.values()) or 1.0
    top = max(kt.items(), key=lambda kv: kv[1][1])
    top_name, (top_cnt, top_ms) = top
    per_launch_ms = top_ms / top_cnt
    # algorithmic bytes of one launch of the dominant kernel: what the kernel must move once
    algo_table = {
        "k_encode_tile": (raw_bytes + blob_bytes, "raster read once + block stream written once"),
        "k_encode_fused": (raw_bytes + blob_bytes, "raster read once + block stream written once"),
        "k_decode_stream": (raw_bytes + blob_bytes, "block stream read once + raster written once"),
        "k_dec_blocks": (raw_bytes + blob_bytes, "block stream read once + raster written once"),
        "k_dec_walk": (blob_bytes, "block stream headers (bounded by the stream size)"),
        "k_dec_candidates": (blob_bytes, "block stream (bounded by the stream size)"),
        "k_fletcher_partial": (blob_bytes, "blob read once"),
        "k_tiles<T, true>": (raw_bytes + blob_bytes, "raster read + blob write (general path)"),
        "k_tiles<T, false>": (raw_bytes, "raster read (general path, count pass)"),
        "k_tiles_decode": (raw_bytes + blob_bytes, "blob read + raster write (general path)"),
        "k_walk_units": (blob_bytes, "block stream (general path, serial walk)"),
        "k_huffman": (raw_bytes + blob_bytes, "raster read + bit stream written"),
        "k_huff_chunks": (blob_bytes, "bit stream read once per synchronisation pass"),
        "k_huff_emit": (raw_bytes + blob_bytes, "bit stream read + symbols written"),
        "k_huff_rows": (2 * raw_bytes, "delta image read + raster written"),
        "k_histograms": (raw_bytes, "raster read"),
    }
    algo_bytes, which = raw_bytes + blob_bytes, "raster + blob"
    for key, (ab, wh) in algo_table.items():
        if key in top_name:
            algo_bytes, which = ab, wh
            break
    achieved = algo_bytes / (per_launch_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")   # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu --set full captures
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        for key, val in tj.get(args.workload, {}).items():
            if key in top_name:
                traffic = val
    roofline = {"bound": "hbm", "kernel": top_name, "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes, "algorithmic_bytes_are": which,
                "kernel_ms_per_launch": per_launch_ms, "kernel_share_of_step": top_ms / PROF / ms_per_step,
                "timer": "CUDA event pair around every launch on the library's stream (lerc_b200_profile), separate pass after the timed region",
                "step_achieved": (2 * (raw_bytes + blob_bytes)) / (ms_per_step * 1e-3) / 1e9,
                "step_frac": (2 * (raw_bytes + blob_bytes)) / (ms_per_step * 1e-3) / 1e9 / peak_gbs,
                "kernels": {n: {"launches_per_step": c / PROF, "ms_per_step": m / PROF} for n, (c, m) in sorted(kt.items(), key=lambda kv: -kv[1][1])[:10]}}

    # ---- end to end: pinned host buffers through the same C ABI ----------------------------------------
    lerc_b200.set_stream(0, False)
    h_in = [torch.from_numpy(h).pin_memory() for h in host_imgs[:2]]
    h_blob = torch.empty(cap, dtype=torch.uint8).pin_memory()
    h_out = torch.empty_like(h_in[0]).pin_memory()

    def step_host(i):
        src = h_in[i % 2]
        st = enc(src.data_ptr(), dt, depth, cols, rows, 1, 0, None, mz, h_blob.data_ptr(), tight[0], C.addressof(n_written))
        assert st == 0
        st = dec(h_blob.data_ptr(), n_written.value, 0, None, depth, cols, rows, 1, dt, h_out.data_ptr())
        assert st == 0
        return n_written.value

    for i in range(warm):
        step_host(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    E2E = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for i in range(E2E):
        nb = step_host(i)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / E2E
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item()) * 1e3
    e2e = {"value": world * n_px / float(t.item()) / 1e9, "unit": "Gpixels/s", "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": raw_bytes + nb, "d2h_bytes_per_step": nb + raw_bytes, "timer": "host wall clock around the synchronous C-API calls",
           "pcie_gbs_per_direction": (raw_bytes + nb) / (e2e_ms * 1e-3) / 1e9,
           "pipelining": "inside each call: strips of block rows / stream chunks, H2D | kernels | D2H on three streams (graded schedule: ~1 MB strips at both ends, up to 8x larger in between; LERC_B200_STRIP_LOG2)"}
    assert np.abs(h_out.numpy().astype(np.float64) - h_in[(E2E - 1) % 2].numpy().astype(np.float64)).max() <= max(mz, 0.5 if dt < 6 else 0) * 1.1 + 1e-12

    # ---- two callers: one thread encodes raster i + 1 while another decodes blob i (two library contexts; the calls are synchronous and
    # release the GIL), so that both PCIe directions are busy all the time.  Reported beside the single-caller figure, not instead of it.
    if True:
        import threading
        h_blob2 = [h_blob, torch.empty(cap, dtype=torch.uint8).pin_memory()]
        nw2 = [C.c_uint(0), C.c_uint(0)]
        errs = []

        def enc_job(i):
            if enc(h_in[i % 2].data_ptr(), dt, depth, cols, rows, 1, 0, None, mz, h_blob2[i % 2].data_ptr(), tight[0], C.addressof(nw2[i % 2])) != 0:
                errs.append(("enc", i))

        def dec_job(i):
            if dec(h_blob2[i % 2].data_ptr(), nw2[i % 2].value, 0, None, depth, cols, rows, 1, dt, h_out.data_ptr()) != 0:
                errs.append(("dec", i))

        def run_two(n):
            enc_job(0)
            for i in range(n):
                ta = threading.Thread(target=enc_job, args=(i + 1,))
                tb = threading.Thread(target=dec_job, args=(i,))
                ta.start(); tb.start(); ta.join(); tb.join()

        run_two(2)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        run_two(E2E)
        two_s = (time.perf_counter() - t0) / (E2E + 0.5)              # (E2E decodes + E2E + 1 encodes)
        t2 = torch.tensor([two_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        assert not errs, errs
        assert np.abs(h_out.numpy().astype(np.float64) - h_in[(E2E - 1) % 2].numpy().astype(np.float64)).max() <= max(mz, 0.5 if dt < 6 else 0) * 1.1 + 1e-12
        e2e["two_callers"] = {"value": world * n_px / float(t2.item()) / 1e9, "unit": "Gpixels/s", "ms_per_step": float(t2.item()) * 1e3,
                              "pcie_gbs_per_direction": (raw_bytes + nb) / float(t2.item()) / 1e9,
                              "what": "encode of raster i + 1 and decode of blob i issued by two host threads at the same time (same C-API calls, same host buffers)"}
    e2e["numa_bound_cpus"] = numa
    # what the link itself does on this box: one 64 MB pinned copy per direction alone, then both directions at once
    probe_d, probe_d2 = torch.empty_like(h_in[0], device="cuda"), torch.empty_like(h_in[0], device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def probe(fn, reps=3):
        best = 1e9
        for _ in range(reps):
            torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        return best

    def both():
        with torch.cuda.stream(s1):
            probe_d.copy_(h_in[0], non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(probe_d2, non_blocking=True)

    t_h2d = probe(lambda: probe_d.copy_(h_in[0], non_blocking=True))
    t_d2h = probe(lambda: h_out.copy_(probe_d2, non_blocking=True))
    t_both = probe(both)
    e2e["pcie_measured_gbs"] = {"h2d_alone": raw_bytes / t_h2d / 1e9, "d2h_alone": raw_bytes / t_d2h / 1e9, "each_direction_when_both_run": raw_bytes / t_both / 1e9,
                                "what": "one pinned copy of the raster per direction (torch copy_, wall clock), best of 3"}
    del probe_d, probe_d2

    # ---- the other BASELINE configs as sub-records (every rank takes part: c5 gathers over NCCL, c4 is one raster per rank)
    subs = {}
    if args.workload == "c2" and not args.no_sub:
        import copy
        sub_args = copy.copy(args)
        sub_args.steps, sub_args.gather = 5, True
        for name in ("c5", "c4", "c3", "c2m"):
            try:
                if name == "c5":
                    subs[name] = main_tiles(sub_args, lib, rank, world, local_rank, warm, peak_gbs, peak_src, sub=True)
                elif name == "c4" or world == 1:
                    subs[name] = bench_raster_sub(lib, name, 5 if name in ("c4", "c2m") else 3, rank, world, peak_gbs)
            except Exception as ex:                                      # a sub-record must not take the headline down with it
                subs[name] = {"error": f"{type(ex).__name__}: {ex}"[:300]}
                torch.cuda.empty_cache()

    if rank == 0:
        cpu = None
        if world >= 1 and not args.no_cpu_baseline and args.gpus == 1:
            cpu, _ = cpu_reference_run(args.workload, seconds_budget=15.0)
        line = {"metric": METRIC[args.workload],
                "value": value, "unit": "Gpixels/s", "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64" if dt >= 6 else "u8", "data": "synthetic",
                "config": {"workload": desc, "generator": "tests/cases.py", "blob_bytes": blob_bytes, "compression_ratio": raw_bytes / blob_bytes,
                           "l2": f"{NBUF} rotating raster/blob/output sets ({NBUF * (2 * raw_bytes + blob_bytes) // 2**20} MiB touched between reuses) > 126 MB L2",
                           "sharding": "one independent raster per rank, no data-path collective" if world > 1 else "single GPU"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu}
        line.update(subs)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
