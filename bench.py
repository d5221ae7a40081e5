#!/usr/bin/env python3
"""Throughput benchmark of the generator's hot path (BASELINE.json metric: image-pair + flow
samples/s at 1/2/4/8 B200, achieved HBM GB/s against the measured peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one batch of `--batch` (default 64) samples per GPU rendered by the hot path
(background preparation + tile render kernels) from scenes already flattened and resident in HBM;
`value` is whole-job samples/s over all ranks (weak scaling: per-GPU batch fixed, per-GPU seed
offset 45*rank). `e2e` is the same metric through the C-ABI call with HOST blobs: host parameter
draw + geometry flattening + scene upload + kernels + device-to-host copy of the three blobs, all
inside the timed region. `--impl reference` times the reference's own CPU generator (oracle/_ref: its
untouched sources compiled here; the oracle port when that library is absent) on all host threads, on
bounded samples of the same workload.

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_512x384 = 7471104  # 6,291,456 B blobs written + 1,179,648 B texels read per sample


def algo_bytes(W, H):
    return 2 * 3 * H * W * 4 + 2 * H * W * 4 + 2 * H * W * 3


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks/throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for l in self.proc.stdout:
            self.lines.append((time.time(), l.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    sm.append(float(f[1]))
                    mx = float(f[2])
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
            except ValueError:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_generator_rate(mode, W, H, n_samples, threads, tex, seed_offset=0, faithful=True):
    """samples/s of the CPU restatement (oracle port, reference structure: one worker thread per task). Used only when the
    reference build is not available (or for sizes other than the reference's compile-time 512 x 384)."""
    import ofdg_b200 as o
    from oracle import binding as ob
    tasks = o.ParamStream(mode, W, H, seed_offset).generate(n_samples)
    t0 = time.time()
    ob.render(tasks.struct(), tex, W=W, H=H, mode=mode, n_threads=threads, faithful=faithful)
    dt = time.time() - t0
    return n_samples / dt, dt


def reference_available(W, H):
    """The reference's own code (oracle/_ref, built from /root/reference by oracle/ref_build.sh) serves its compile-time size only."""
    try:
        from oracle import ref_binding as rb
        return (W, H) == (rb.W, rb.H) and rb.available()
    except Exception:
        return False


class ReferenceLayerRun:
    """The reference's DataGenerationLayer on this box's host cores: texture list of PPM files -> TextureCollection ->
    first_level_threads worker threads (DataGenerator::WorkerThreadLoop) -> load_batch -> Forward_cpu. One step = one Forward."""

    def __init__(self, mode, batch, threads, seed):
        import tempfile
        import ofdg_b200 as o
        from oracle import ref_binding as rb
        self.rb = rb
        self.dir = tempfile.mkdtemp(prefix="ofdg_ref_pool_")
        tex = o.synth_textures(16, 1024, 768, seed=seed)  # a small pool: the pool size does not change the CPU cost per sample
        lst = rb.write_ppm_pool(self.dir, tex)
        self.batch, self.threads = batch, threads
        # second_level_threads = 1 is the proto default (caffe.proto:10); prefetch 2 lets generation overlap the consumer
        self.layer = rb.Layer(mode, lst, batch=batch, prefetch=2, first_level_threads=threads, second_level_threads=1)

    def forward(self):
        self.layer.forward(copy=False)

    def close(self):
        import shutil
        self.layer.close()
        shutil.rmtree(self.dir, ignore_errors=True)


def reference_layer_rate(mode, cores, seconds, seed=0):
    """Bounded sample: forwards of `cores` samples until `seconds` of wall clock have passed (after one warm-up forward)."""
    run = ReferenceLayerRun(mode, batch=max(cores, 4), threads=cores, seed=seed)
    run.forward()
    n, t0 = 0, time.time()
    while True:
        run.forward()
        n += run.batch
        dt = time.time() - t0
        if dt >= seconds:
            break
    run.close()
    return n / dt, dt, n


def run_reference(args):
    """The reference arm: the reference's own CPU generator on this box's host cores, same workload/metric. Its own code
    (oracle/_ref) when that library is present, else the oracle port of it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    W, H, mode = args.width, args.height, args.mode
    if reference_available(W, H):
        from oracle import ref_binding as rb
        # calibrate, then size the per-step batch so that the whole run stays within ~3 minutes
        rate, _, _ = reference_layer_rate(mode, cores, 4.0, args.seed)
        per_step = int(max(4, min(4 * cores, rate * 150.0 / (args.steps + args.warmup))))
        run = ReferenceLayerRun(mode, per_step, cores, args.seed)
        for _ in range(max(args.warmup, 1)):
            run.forward()
        t0 = time.time()
        for _ in range(args.steps):
            run.forward()
        dt = time.time() - t0
        run.close()
        kind, threads = "reference", cores
        sample = (f"{per_step} samples per step x {args.steps} steps through the reference's own DataGenerationLayer::Forward_cpu "
                  f"(oracle/_ref: {rb.describe()}), mode {mode}, {W}x{H}, first_level_threads={cores}, second_level_threads=1, prefetch=2, "
                  f"16-texture PPM pool, {cores} host cores")
    else:
        import ofdg_b200 as o
        from oracle import binding as ob
        ob.build()
        tex = o.synth_textures(16, 2 * W, 2 * H, seed=args.seed)
        rate, _ = cpu_generator_rate(mode, W, H, min(cores, 16), min(cores, 16), tex)
        per_step = int(max(1, min(2 * cores, rate * 150.0 / (args.steps + args.warmup))))
        threads = min(cores, per_step)
        ps = o.ParamStream(mode, W, H, 0)
        tasks = o.Tasks()

        def step():
            tasks.clear()
            ps.generate(per_step, tasks)  # parameter draws are part of the reference's per-batch work too
            ob.render(tasks.struct(), tex, W=W, H=H, mode=mode, n_threads=threads, faithful=True)

        for _ in range(args.warmup):
            step()
        t0 = time.time()
        for _ in range(args.steps):
            step()
        dt = time.time() - t0
        kind = "port"
        sample = (f"{per_step} samples per step x {args.steps} steps, mode {mode}, {W}x{H}, oracle/liboracle.so with the reference's "
                  f"whole-image copies, {threads} worker threads of {cores} host cores")
    v = per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": "img-pair+flow samples/sec", "value": v, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_config(args):
    return {"workload": f"mode {args.mode} (example prototxt) {args.width}x{args.height}, batch {args.batch} per GPU per step, "
                        f"{args.textures}-texture procedural pool resident in HBM, host-RNG parameter stream",
            "global_batch": args.batch * args.gpus, "mode": args.mode, "textures": args.textures,
            "l2": "inputs+outputs larger than L2 (403 MB of blobs written per step per GPU; distinct scene batches cycled)",
            "parallelism": f"sample-sharded x{args.gpus}, seed offset 45*rank, no collective"}


def traffic_table():
    """DRAM bytes per launch of each kernel from the committed ncu --set full capture (batch 64, 512x384, mode 7)."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name))), name
        except Exception:
            pass
    return {}, None


def attribution_pass(o, args, device, task_sets, img0, img1, flow, stream, steps):
    """Per-kernel times of the step: a second generator created with the overlaps switched off (the switches are read at
    ofdg_create), so that bg_prep, bin_pairs, raster_pairs and shade run one after the other on one stream and the CUDA-event
    spans around each launch do not overlap. Returns {kernel: {ms, bytes (the kernel's own mandatory traffic), achieved GB/s,
    frac of the measured HBM peak, traffic (dram bytes per launch from the committed ncu capture)}}."""
    import torch
    W, H, B = args.width, args.height, args.batch
    keys = ("OFDG_PIPELINE", "OFDG_RASTER_OVERLAP", "OFDG_BIN_OVERLAP")
    saved = {k: os.environ.get(k) for k in keys}
    for k in keys:
        os.environ[k] = "0"
    try:
        g2 = o.Generator(device=device, width=W, height=H, mode=args.mode, max_batch=B)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    g2.synth_textures(args.textures, 2 * W, 2 * H, seed=args.seed)
    prepared = [g2.prepare(t) for t in task_sets]
    for i in range(8):
        g2.render_prepared(prepared[i % len(prepared)], img0, img1, flow, stream)
    g2.kernel_times()
    pairs = prep_px = src_px = 0
    for i in range(steps):
        g2.render_prepared(prepared[i % len(prepared)], img0, img1, flow, stream)
        if i < len(prepared):  # sizes of each distinct scene batch (synchronises; the spans are per launch, so this does not disturb them)
            a, b, c = g2.last_render_stats()
            pairs += a; prep_px += b; src_px += c
    prep_ms, render_ms, calls = g2.kernel_times()
    shade_ms = g2.last_shade_ms()
    bin_ms, raster_ms = g2.last_bin_raster_ms()
    n = min(steps, len(prepared))
    pairs, prep_px, src_px = pairs / n, prep_px / n, src_px / n
    peak, _ = measured_peak()
    traffic, traffic_src = traffic_table()
    own = {
        # prepared pixels written (RGBX) + the source texels under them read once (RGBX)
        "bg_prep_kernel": (prep_ms, 4 * prep_px + 4 * src_px),
        # reads the objects' boxes, writes the pair list and the tiles' ranges
        "bin_pairs_kernel": (bin_ms, 16 * pairs + 8 * B * ((W + 127) // 128) * ((H + 7) // 8)),
        # four 128x8 coverage masks per (object, tile) pair
        "raster_pairs_kernel": (raster_ms, 4096 * pairs),
        # blobs written + texels read
        "shade_kernel": (shade_ms, algo_bytes(W, H) * B),
    }
    out = {}
    for name, (tot_ms, nbytes) in own.items():
        ms = tot_ms / max(calls, 1)
        ach = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else None
        t = traffic.get(name, {}).get("traffic_bytes") if (B == 64 and W == 512 and H == 384 and args.mode == 7) else None
        out[name] = {"ms": ms, "bytes": int(nbytes), "achieved": ach, "frac": (ach / peak) if ach else None, "traffic": t}
    out["shade_kernel"]["note"] = "moves every algorithmic byte of the step (SURVEY 8d)"
    out["bg_prep_kernel"]["note"] = "own bytes: prepared RGBX pixels written + source RGBX texels read once; instruction-bound (CImg float/double chain)"
    out["raster_pairs_kernel"]["note"] = "own bytes: 4 KB of masks per (object, tile) pair, which stay in L2 for the shade kernel"
    del g2
    return out


def timed_render_leg(o, torch, g, task_sets, B, W, H, stream, steps):
    """samples/s and ms per step of render_prepared over the given scene batches (CUDA events on the launching stream)."""
    prepared = [g.prepare(t) for t in task_sets]
    img0 = torch.empty((B, 3, H, W), device="cuda", dtype=torch.float32)
    img1 = torch.empty_like(img0)
    flow = torch.empty((B, 2, H, W), device="cuda", dtype=torch.float32)
    for i in range(6):
        g.render_prepared(prepared[i % len(prepared)], img0, img1, flow, stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        g.render_prepared(prepared[i % len(prepared)], img0, img1, flow, stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    peak, _ = measured_peak()
    ab = algo_bytes(W, H) * B
    del prepared, img0, img1, flow
    return {"value": B / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms, "batch": B, "steps": steps,
            "algorithmic_bytes_per_step": ab, "whole_step_achieved_gbs": ab / (ms * 1e-3) / 1e9, "whole_step_frac": ab / (ms * 1e-3) / 1e9 / peak}


def other_config_legs(o, torch, args, device, stream):
    """BASELINE.json configs 3 and 5, device-resident scenes like the headline `value`:
    config 3 = mode 9 (all shapes, composites, thin objects, non-rigid motion) with a pool of 40 GPU-generated warp fields and the
    colour/noise augmentation, batch 64; config 5 = 1024 x 768 output, 64 and 128 forced foreground objects per sample,
    10,000-texture pool (31 GB) resident in HBM, batch 16 (the same 403 MB of blobs per step)."""
    out = {}
    steps = max(10, min(args.steps, 100))
    # ---- config 3
    g = o.Generator(device=device, width=512, height=384, mode=9, max_batch=64)
    g.synth_textures(args.textures, 1024, 768, seed=args.seed)
    g.generate_fields(5, 40)
    ps = o.ParamStream(9, 512, 384, n_fields=40)
    ps.enable_augmentation(True)
    leg = timed_render_leg(o, torch, g, [ps.generate(64) for _ in range(3)], 64, 512, 384, stream, steps)
    leg["workload"] = f"mode 9 + 40 generated warp fields + colour/noise augmentation, 512x384, batch 64, {args.textures}-texture pool"
    out["config3"] = leg
    g.close()
    # ---- config 5
    free, _ = torch.cuda.mem_get_info()
    n_tex = 10000 if free > 60e9 else 1000
    W5, H5, B5 = 1024, 768, 16
    g = o.Generator(device=device, width=W5, height=H5, mode=7, max_batch=B5)
    g.synth_textures(n_tex, 1024, 768, seed=args.seed)
    for n_obj in (64, 128):
        ps = o.ParamStream(7, W5, H5, fg_override=n_obj)
        leg = timed_render_leg(o, torch, g, [ps.generate(B5) for _ in range(3)], B5, W5, H5, stream, max(10, steps // 2))
        leg["workload"] = (f"mode 7, 1024x768 output, {n_obj} forced foreground objects per sample, batch {B5}, {n_tex} textures of 1024x768 "
                           f"({n_tex * 1024 * 768 * 4 / 1e9:.1f} GB as RGBX in HBM; smaller than 2W x 2H: the background takes getRandomizedCrop's resize-whole branch)")
        out[f"config5_{n_obj}_objects"] = leg
    g.close()
    return out


def host_memory_probe(o, threads=None, mb=96, reps=3):
    """What the host's memory system sustains for the host-blob path's last stage: `threads` threads widening uint8 -> float32 with
    non-temporal stores (csrc/host/expand.cpp), GB/s of DRAM traffic (1 byte read + 4 written per element)."""
    import numpy as np
    if threads is None:
        threads = max(4, min(32, (os.cpu_count() or 8) // 2))
    n = mb << 20
    src = [np.full(n, 7, np.uint8) for _ in range(threads)]
    dst = [np.empty(n, np.float32) for _ in range(threads)]
    for i in range(threads):
        o.expand_host(src[i], dst[i])  # touch the pages
    best = 0.0
    for _ in range(reps):
        ts = [threading.Thread(target=o.expand_host, args=(src[i], dst[i])) for i in range(threads)]
        t0 = time.time()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        dt = time.time() - t0
        best = max(best, threads * n * 5 / dt / 1e9)
    return {"threads": threads, "gbs": best, "what": "uint8 -> float32 widening with non-temporal stores, read + written bytes per second"}


def run_ours(args):
    import numpy as np
    import torch
    import ofdg_b200 as o
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert torch.cuda.is_available(), "bench.py needs a B200 (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H, B, mode = args.width, args.height, args.batch, args.mode
    g = o.Generator(device=local, width=W, height=H, mode=mode, max_batch=B)
    g.synth_textures(args.textures, 2 * W, 2 * H, seed=args.seed)
    ps = o.ParamStream(mode, W, H, seed_offset=45 * rank)
    n_sets = 4
    prepared_tasks = [ps.generate(B) for _ in range(n_sets)]
    prepared = [g.prepare(t) for t in prepared_tasks]
    img0 = torch.empty((B, 3, H, W), device="cuda", dtype=torch.float32)
    img1 = torch.empty_like(img0)
    flow = torch.empty((B, 2, H, W), device="cuda", dtype=torch.float32)
    tstream = torch.cuda.Stream(priority=int(os.environ.get("OFDG_BENCH_STREAM_PRIORITY", "0")))  # a non-default stream: torch events and our launches share it (experiments: -1 = high priority)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    t_load = time.time()
    # keep the GPU under the same load for at least ~0.5 s before timing, so that the clock samples
    # (nvidia-smi polls every 100 ms) describe the timed region even when --steps is small
    i = 0
    while i < args.warmup or time.time() - t_load < 0.5:
        g.render_prepared(prepared[i % n_sets], img0, img1, flow, stream)
        i += 1
        if i % 50 == 0:
            torch.cuda.synchronize()
    barrier()
    g.kernel_times()
    launches0 = g.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_start = time.time()
    e0.record()
    for i in range(args.steps):
        g.render_prepared(prepared[i % n_sets], img0, img1, flow, stream)
    e1.record()
    barrier()
    t_end = time.time()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t_load, t_end) if sampler else None
    launches = g.launch_count() - launches0
    prep_ms, render_ms, calls = g.kernel_times()
    shade_ms = g.last_shade_ms()  # 0 when the single-kernel render path is selected (OFDG_RENDER=fused)
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B * args.steps / (ms * 1e-3)

    # ---- attribution pass (rank 0): the same scenes with every kernel in line on one stream, per-kernel CUDA-event spans
    attribution = None
    if rank == 0 and not args.no_attribution and os.environ.get("OFDG_RENDER") != "fused":
        attribution = attribution_pass(o, args, local, prepared_tasks, img0, img1, flow, stream, min(args.steps, 200))

    # ---- production mode: parameters drawn + flattened on the device (Philox), nothing crosses PCIe
    production = None
    if mode != 9:
        for i in range(3):
            g.generate_philox(args.seed, i * B, B, img0, img1, flow, stream=stream)
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for i in range(args.steps):
            g.generate_philox(args.seed + rank, (3 + i) * B, B, img0, img1, flow, stream=stream)
        p1.record()
        barrier()
        pms = p0.elapsed_time(p1)
        if dist is not None:
            t = torch.tensor([pms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            pms = float(t.item())
        production = {"value": world * B * args.steps / (pms * 1e-3), "unit": "samples/s", "ms_per_step": pms / args.steps,
                      "what": "ofdg_generate_philox: device-side Philox parameter stream + device flattening + render, fresh scenes every step, no host work"}
        g.kernel_times()

    # ---- end to end through the C ABI with host blobs (pinned), copies inside the timed region
    h0 = torch.empty((B, 3, H, W), dtype=torch.float32).pin_memory()
    h1 = torch.empty_like(h0).pin_memory()
    hf = torch.empty((B, 2, H, W), dtype=torch.float32).pin_memory()
    e2e_steps = max(3, min(args.steps, args.e2e_steps))
    for _ in range(2):
        g.generate_host(ps, B, h0, h1, hf)
    barrier()
    t0 = time.time()
    h2d = 0
    for _ in range(e2e_steps):
        g.generate_host(ps, B, h0, h1, hf)  # parameter draw + flatten + H2D + kernels + D2H, all inside
        h2d += g.last_upload_bytes()
    d2h = g.last_download_bytes()
    torch.cuda.synchronize()
    dt = time.time() - t0
    if dist is not None:
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e = world * B * e2e_steps / dt
    checksum = float(h0.mean()) + float(h1.mean()) + float(hf.mean())  # every blob element was delivered to the host
    g.kernel_times()

    other = None
    if rank == 0 and world == 1 and not args.no_other_configs:
        other = other_config_legs(o, torch, args, local, stream)

    # ---- optional epilogue at N > 1: the ranks' finished blobs gathered to a training rank over NVLink (NCCL gather = send / recv)
    gather_leg = None
    if dist is not None and not args.no_gather:
        outs = None
        if rank == 0:
            outs = [torch.empty((world * B, 3, H, W), device="cuda"), torch.empty((world * B, 3, H, W), device="cuda"),
                    torch.empty((world * B, 2, H, W), device="cuda")]
        torch.cuda.set_stream(torch.cuda.default_stream())
        for _ in range(3):
            o.gather_blobs([img0, img1, flow], dst=0, out=outs)
        barrier()
        n_g = 20
        ga, gb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ga.record()
        for _ in range(n_g):
            o.gather_blobs([img0, img1, flow], dst=0, out=outs)
        gb.record()
        barrier()
        gms = ga.elapsed_time(gb) / n_g
        t = torch.tensor([gms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gms = float(t.item())
        recv = (world - 1) * B * 8 * H * W * 4
        gather_leg = {"ms_per_gather": gms, "bytes_into_rank0": recv, "gbs_into_rank0": recv / (gms * 1e-3) / 1e9,
                      "samples_per_s": world * B / (gms * 1e-3),
                      "what": f"ofdg_b200.gather_blobs: {world} x {B} samples of float blobs gathered to rank 0 with torch.distributed (NCCL) over NVLink; "
                              "not part of `value` (no collective on the generation path)"}
        torch.cuda.set_stream(tstream)
        del outs

    # ---- the Caffe-style layer, the actual drop-in: prototxt -> LayerRegistry -> LayerSetUp -> Forward_gpu x steps
    layer_leg = None
    if world == 1 and not args.no_layer and (W, H) == (512, 384):
        layer_leg = {}
        for name, extra in (("host_rng", ""), ("device_params", " device_params: true seed: %d" % args.seed)):
            if name == "device_params" and mode == 9:
                continue
            proto = ('layer { name: "gen" type: "DataGeneration" top: "img0" top: "img1" top: "flow" data_param { batch_size: %d prefetch: 4 } '
                     'data_generation_param { mode: %d texture_dbases: "synthetic:%d:%d"%s } }' % (B, mode, args.textures, args.seed, extra))
            layer = o.DataGenerationLayer(proto)
            layer.LayerSetUp()
            for _ in range(6):
                layer.Forward_gpu()
            torch.cuda.synchronize()
            if name == "host_rng":
                layer.producer_stats()
            n_fw = max(20, min(args.steps, 300))
            t0 = time.time()
            for _ in range(n_fw):
                layer.Forward_gpu()  # non-blocking: queues the batch on the layer's stream, the default stream waits by event
            torch.cuda.synchronize()
            dt = time.time() - t0
            layer_leg[name] = {"value": B * n_fw / dt, "unit": "samples/s", "ms_per_forward": 1e3 * dt / n_fw, "forwards": n_fw}
            if name == "host_rng":
                draw_ms, prep_ms_l, nb = layer.producer_stats()
                layer_leg[name]["producer"] = {"draw_ms_per_batch": draw_ms, "prepare_ms_per_batch": prep_ms_l, "batches": nb,
                                               "threads": "2 prefetch threads (draw of batch k+1 beside prepare of batch k) + look-ahead helpers + flatten pool"}
            layer.close()
        layer_leg["what"] = ("DataGenerationLayer::Forward_gpu into the top blobs' device memory, wall clock over back-to-back forwards incl. the "
                             "prefetch thread's parameter draw + flatten (host pool) + scene upload; host_rng = the reference's RNG stream, "
                             "device_params = the Philox production stream")

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    host_probe = host_memory_probe(o) if not args.no_cpu else None

    peak, peak_src = measured_peak()
    split = shade_ms > 0
    dominant = "shade_kernel" if split else "render_kernel"
    ab = algo_bytes(W, H) * B
    step_ms = ms / args.steps
    # The step is bg_prep_kernel + bin_pairs_kernel + raster_pairs_kernel (coverage masks of every (object, tile) pair) +
    # shade_kernel. shade_kernel reads the prepared background and the textures and writes the three blobs, i.e. it moves every
    # byte of SURVEY 8d's per-sample figure: the roofline entry is about it. In the timed region the kernels of consecutive
    # batches overlap (front end of batch k+1 beside the shade kernel of batch k), so per-kernel durations come from the
    # attribution pass above: the same scenes, same process, every kernel in line on one stream, CUDA events around each launch.
    kern_ms = attribution["shade_kernel"]["ms"] if attribution else (shade_ms if split else render_ms) / max(calls, 1)
    achieved = ab / (kern_ms * 1e-3) / 1e9
    traffic = args.traffic
    if traffic is None and attribution and B == 64 and W == 512 and H == 384:
        traffic = attribution["shade_kernel"].get("traffic")
    line = {
        "metric": "img-pair+flow samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload_config(args),
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d // e2e_steps,
                "d2h_bytes_per_step": d2h, "host_blob_bytes_per_step": B * (2 * 3 + 2) * H * W * 4,
                "transport": ("uint8 frames + float32 flow over PCIe, widened to the float blobs by host threads inside the timed region"
                              if d2h < B * 8 * H * W * 4 else "float32 blobs over PCIe"),
                "steps": e2e_steps, "checksum": checksum,
                # host DRAM traffic the path implies per step: the DMA writes (d2h), the byte frames read back by the widening
                # threads, the float frames they write (the flow lands in the caller's blob directly)
                "host_dram_bytes_per_step": (d2h + B * 6 * H * W + B * 6 * H * W * 4) if d2h < B * 8 * H * W * 4 else d2h,
                "host_dram_gbs_implied": ((d2h + B * 6 * H * W + B * 6 * H * W * 4) if d2h < B * 8 * H * W * 4 else d2h) * (e2e / (world * B)) * world / 1e9,
                "host_probe": host_probe},
        "gpu_launches": launches,
        "production_mode": production,
        "layer_forward_gpu": layer_leg,
        "other_configs": other,
        "gather_to_training_rank": gather_leg,
        "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": ab, "kernel_ms": kern_ms,
                     "timed": ("attribution pass: OFDG_PIPELINE=0 OFDG_RASTER_OVERLAP=0 OFDG_BIN_OVERLAP=0, every kernel in line, CUDA events "
                               "around each launch" if attribution else "CUDA events around the launches of the timed region"),
                     "kernels": attribution,
                     "serial_step_ms": sum(k["ms"] for k in attribution.values()) if attribution else None,
                     "shade_ms_in_step": (shade_ms / max(calls, 1)) if split else None,
                     "whole_step": {"ms": step_ms, "achieved": ab / (step_ms * 1e-3) / 1e9, "frac": ab / (step_ms * 1e-3) / 1e9 / peak,
                                    "what": "algorithmic bytes of the batch / ms_per_step of the timed region (kernels of consecutive batches overlap)"}},
    }
    # CPU generator next to it (N=1 only): bounded sample on this box's host cores
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        if reference_available(W, H):
            from oracle import ref_binding as rb
            rate, dt_cpu, n = reference_layer_rate(mode, cores, 12.0, args.seed)
            line["cpu_baseline"] = {"value": rate, "unit": "samples/s", "cores": cores, "kind": "reference",
                                    "sample": f"{n} samples of the same workload in {dt_cpu:.1f} s through the reference's own DataGenerationLayer::Forward_cpu "
                                              f"(oracle/_ref: {rb.describe()}), first_level_threads={cores}, second_level_threads=1, {cores} host cores"}
        else:
            from oracle import binding as ob
            ob.build()
            tex = o.synth_textures(16, 2 * W, 2 * H, seed=args.seed)
            n = int(max(8, min(128, 6 * cores)))  # ~5 s of wall clock on every core of the bench box (16 cores: 96 samples)
            rate, dt_cpu = cpu_generator_rate(mode, W, H, n, min(cores, n), tex)
            line["cpu_baseline"] = {"value": rate, "unit": "samples/s", "cores": min(cores, n), "kind": "port",
                                    "sample": f"{n} samples of the same workload in {dt_cpu:.1f} s (oracle/liboracle.so, reference structure "
                                              f"with its whole-image copies), {min(cores, n)} worker threads of {cores} host cores"}
    emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


_STDOUT_FD = None


def emit(line):
    sys.stdout.flush()
    if _STDOUT_FD is not None:
        os.dup2(_STDOUT_FD, 1)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--mode", type=int, default=7)
    ap.add_argument("--width", type=int, default=512)
    ap.add_argument("--height", type=int, default=384)
    ap.add_argument("--textures", type=int, default=1000)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=60)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-attribution", action="store_true", help="skip the per-kernel attribution pass")
    ap.add_argument("--no-layer", action="store_true", help="skip the DataGenerationLayer::Forward_gpu leg")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the config 3 / config 5 legs")
    ap.add_argument("--no-gather", action="store_true", help="N > 1: skip the gather-to-rank-0 leg")
    ap.add_argument("--traffic", type=float, default=None, help="dram bytes per launch from an ncu --set full capture (profiles/)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and args.impl == "ours":
        if world == 1 and args.gpus > 1:
            # relaunch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
            sys.exit(subprocess.call(cmd))
    # stdout carries the JSON line and nothing else: libraries that write banners to fd 1 (NCCL prints its version
    # there) are pointed at stderr until the line is printed
    sys.stdout.flush()
    global _STDOUT_FD
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
