#!/usr/bin/env python
"""bench.py — 500-bp POA windows/s at 32x coverage (BASELINE.json metric), one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (sm_100a kernels via the C ABI)
  python bench.py --impl reference ...                           the reference's own CPU path (oracle/_ref)
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N   one rank per GPU, windows sharded (weak scaling)

A "step" = one pass of the hot path over one batch: BASELINE config 2, 10 000 synthetic windows, w=500,
32 layers, 12 % ONT-like error, scores 3/-5/-4 (SURVEY.md §8d generator, seed 42).
  value      : windows/s with the batch resident in HBM (kernel launches only), CUDA-event timed
  e2e        : windows/s through the whole plugin call: add windows (host buffers -> pinned staging), rp_poa_run
               (H2D + kernel + D2H), rp_poa_fetch_all
  roofline   : algorithmic bytes (2 B x sum (L+1)[(N+1)+E], counted by the kernel's own
               device counters in a separate untimed launch) / kernel time, against the measured HBM peak
  cpu_baseline: the unmodified reference (oracle/_ref) on the host cores, bounded sample (rank 0, N=1)
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

WORKLOAD = "synthetic 10k windows, w=500, 32x ONT-error reads (12%), m/x/g=3/-5/-4, seed 42"
METRIC = "500-bp POA windows/sec at 32x coverage"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """Per-launch DRAM bytes of the POA kernel from the committed ncu --set full capture, if any."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args, rank, world):
    """The reference's own CPU implementation (Window::generate_consensus over spoa, compiled unmodified
    into oracle/_ref) on all host threads.  Each step = a bounded sample of the same workload."""
    if rank != 0:
        return
    from oracle import bindings as ob
    from racon_b200 import windows
    cores = host_cores()
    if not ob.have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libracon_ref.so not built"}))
        return
    n_sample = min(args.windows, 200 * cores)
    ws, _ = windows.synth_windows(n_sample, err=0.12)
    # the reference scales poorly past the physical cores (allocator contention in spoa::Graph): give it the
    # better of "all hardware threads" and "half of them", decided on an untimed probe
    threads = cores
    if cores >= 4:
        probe = ws.subset(range(min(n_sample, 16 * cores)))
        t_all = ob.ref_consensus(probe, threads=cores)[2]
        t_half = ob.ref_consensus(probe, threads=cores // 2)[2]
        threads = cores if t_all <= t_half else cores // 2
    for _ in range(args.warmup):
        ob.ref_consensus(ws, threads=threads)
    t0 = time.perf_counter()
    secs = 0.0
    for _ in range(args.steps):
        _, _, s = ob.ref_consensus(ws, threads=threads)
        secs += s
    wall = time.perf_counter() - t0
    value = n_sample * args.steps / secs
    sample = "%d of the workload's windows per step, %d host threads (of %d hardware threads; best of all/half), " \
             "consensus loop only" % (n_sample, threads, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "windows/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample, "wall_s": wall},
        "cpu_baseline": {"value": value, "unit": "windows/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--windows", type=int, default=10000, help="windows per GPU per step (BASELINE config 2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-batches", type=int, default=2,
                    help="batch objects the end-to-end arm cycles through (racon's -c/--cudapoa-batches)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = env_int("RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    local = env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from racon_b200 import api, shard, windows

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the POA hot path has no CPU fallback")
    torch.cuda.set_device(local)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # each rank generates ITS OWN shard of the stream of synthetic windows (weak scaling: per-GPU work fixed)
    state = 42
    for _ in range(rank):
        state = _skip_state(windows, args.windows, state)
    ws, _ = windows.synth_windows(args.windows, err=0.12, state=state)
    n = ws.n_windows

    batch = api.PoaBatch(device=local, window_length=500)
    stream = torch.cuda.current_stream()
    batch.set_stream(stream.cuda_stream)
    t_pack0 = time.perf_counter()
    took = batch.add_window_set(ws)
    pack_ms = 1e3 * (time.perf_counter() - t_pack0)
    assert took == n, "batch object could not hold the workload (%d of %d windows)" % (took, n)
    lens_in = np.diff(ws.seq_off.astype(np.int64))
    stride = int(2 * lens_in.max() + 64)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- counters launch (untimed): algorithmic bytes of this batch + a correctness probe ----------
    batch.enable_counters(True)
    batch.run()
    batch.sync()
    info = batch.info()
    out, lens, pol, st = batch.fetch_all(stride)
    if (st != 0).any() or not pol.all():
        raise SystemExit("bench.py: %d windows hit a device limit" % int((st != 0).sum()))
    checksum = "%016x" % windows.fnv1a64([out[i, :lens[i]].tobytes() for i in range(min(n, 200))])
    alg_bytes = 2.0 * (info["dp_cells"] + info["pred_cells"])
    batch.enable_counters(False)

    # ---- kernel-only arm: inputs resident in HBM --------------------------------------------------
    # (1) the kernel alone: isolated launches on one stream, CUDA events around each -> roofline.kernel_ms
    batch.upload()
    for _ in range(args.warmup):
        batch.launch()
    barrier()
    iso = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    iso[0].record(stream)
    for k in range(3):
        batch.launch()
        iso[k + 1].record(stream)
    barrier()
    kern_ms = [iso[k].elapsed_time(iso[k + 1]) for k in range(3)]
    # (2) the K timed steps: one launch per step over the resident batch, steps alternating between two batch objects
    # on two streams (both hold the whole batch in HBM), so that the tail of one step — the last windows of a launch
    # leave most SMs idle — is filled by the head of the next, as it is in a real multi-batch run
    batch2 = api.PoaBatch(device=local, window_length=500)
    stream2 = torch.cuda.Stream()
    batch2.set_stream(stream2.cuda_stream)
    assert batch2.add_window_set(ws) == n
    batch2.upload()
    batch2.launch()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1, evb = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    launches0 = batch.info()["launches"] + batch2.info()["launches"]
    ev0.record(stream)
    stream2.wait_event(ev0)
    for k in range(args.steps):
        (batch if k % 2 == 0 else batch2).launch()
    evb.record(stream2)
    stream.wait_event(evb)
    ev1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ev0.elapsed_time(ev1)
    gpu_launches = batch.info()["launches"] + batch2.info()["launches"] - launches0
    batch2.close()

    # ---- end-to-end arm: C-ABI call with host buffers (H2D + kernel + D2H + fetch) ------------------
    # what racon's CUDAPolisher does with `-c K` batch objects (cudapolisher.cpp:254-276): each object, on its own
    # stream, is reset, filled from host buffers (copied into pinned staging), run (H2D + kernel + D2H, asynchronous)
    # and read back; while one object's kernel runs, the host fills the next one.
    nb = max(1, args.e2e_batches)
    # (each on its own stream; torch's current stream stays free for the NCCL gather)
    objs = [api.PoaBatch(device=local, window_length=500) for _ in range(nb)]
    bounds = [n * k // nb for k in range(nb + 1)]

    last = {}

    def finalize(parts):
        out, lens, pol, st = (np.concatenate([p[i] for p in parts]) for i in range(4))
        if distributed:  # the one exchange step of the path: final consensus gather over NCCL
            flat = np.concatenate([out[i, :lens[i]] for i in range(len(lens))])
            last["gathered"] = shard.gather_packed(flat, lens, device="cuda")
        last["out"], last["lens"] = out, lens

    def plugin_steps(n_steps):
        # the batch objects stay in flight across steps, as CUDAPolisher keeps its batches busy until the window list
        # is exhausted: object k takes chunk k of every step; before it is refilled its previous results are read back
        pending = [None] * nb
        parts = {}

        def collect(k):
            objs[k].sync()
            s_done = pending[k]
            parts[s_done][k] = objs[k].fetch_all(stride)
            pending[k] = None
            if all(x is not None for x in parts[s_done]):
                finalize(parts.pop(s_done))

        for s_i in range(n_steps):
            parts[s_i] = [None] * nb
            for k, b in enumerate(objs):
                if pending[k] is not None:
                    collect(k)
                b.reset()
                cnt = bounds[k + 1] - bounds[k]
                assert b.add_window_set(ws, first=bounds[k], count=cnt) == cnt
                b.run()
                pending[k] = s_i
        for k in range(nb):
            if pending[k] is not None:
                collect(k)

    plugin_steps(2)
    barrier()
    t0 = time.perf_counter()
    plugin_steps(args.steps)
    barrier()
    out, lens = last["out"], last["lens"]
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    if windows.fnv1a64([out[i, :lens[i]].tobytes() for i in range(min(n, 200))]) != int(checksum, 16):
        raise SystemExit("bench.py: end-to-end arm produced a different consensus than the kernel-only arm")
    io = objs[0].info()
    io["h2d_bytes"] = sum(b.info()["h2d_bytes"] for b in objs)
    io["d2h_bytes"] = sum(b.info()["d2h_bytes"] for b in objs)

    # ---- max over ranks ---------------------------------------------------------------------------
    tt = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(tt[0]), float(tt[1])

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        ms_per_step = total_ms / args.steps
        value = world * n / (ms_per_step * 1e-3)
        kern_avg_ms = sum(kern_ms) / len(kern_ms)
        achieved = alg_bytes / (kern_avg_ms * 1e-3) / 1e9
        traffic = ncu_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": "windows/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "windows_per_gpu": n, "window_len": 500, "depth": 32,
                       "parallelism": "windows sharded across %d GPU(s), no data-path collective" % world,
                       "l2": "inputs (%.0f MB) + per-step DP scratch (>> 126 MB) exceed L2; no explicit flush"
                             % (io["h2d_bytes"] / 1e6),
                       "steps_overlap": "timed steps alternate between two resident batch objects on two streams; "
                                        "roofline.kernel_ms is an isolated launch",
                       "worker_warps": io["workers"], "first_pack_ms": pack_ms,
                       "consensus_fnv_first200": checksum},
            "e2e": {"value": world * n * args.steps / (e2e_ms * 1e-3), "unit": "windows/s",
                    "h2d_bytes_per_step": io["h2d_bytes"], "d2h_bytes_per_step": io["d2h_bytes"],
                    "batch_objects": nb,
                    "includes": "every step, per batch object: rp_poa_reset + rp_poa_add_window_set (host buffers -> pinned "
                                "staging) + rp_poa_run (H2D, kernel, D2H) + rp_poa_sync + rp_poa_fetch_all; the objects stay "
                                "in flight across steps (as CUDAPolisher keeps its batches busy), timed from the first add "
                                "to the last fetch"
                                + ("; + NCCL all_gather of consensus bytes" if distributed else "")},
            "gpu_launches": int(gpu_launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": (traffic or {}).get("dram_bytes_per_launch"),
                         "kernel": "rp_poa_kernel", "kernel_ms": kern_avg_ms,
                         "achieved_overlapped": alg_bytes / (ms_per_step * 1e-3) / 1e9,
                         "frac_overlapped": alg_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "algorithmic_bytes_per_window": alg_bytes / n, "peak_source": peak_src},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(line))
    for b in objs + [batch]:
        b.close()
    if distributed:
        dist.destroy_process_group()


def _skip_state(windows, n, state):
    """Advances the generator state past one rank's shard (so every rank draws a different shard)."""
    _, st = windows.synth_windows(n, err=0.12, state=state)
    return st


def cpu_baseline(args):
    """oracle/_ref (the unmodified reference) on the box's host cores, bounded sample (~10-30 s of CPU work)."""
    from oracle import bindings as ob
    from racon_b200 import windows
    cores = host_cores()
    if ob.have_ref():
        kind, fn = "reference", ob.ref_consensus
    else:
        kind, fn = "port", ob.oracle_consensus
    n_sample = min(args.windows, max(64, 300 * cores))
    ws, _ = windows.synth_windows(n_sample, err=0.12)
    threads = cores
    if cores >= 4:  # see run_reference(): best of all hardware threads / half of them
        probe = ws.subset(range(min(n_sample, 16 * cores)))
        threads = cores if fn(probe, threads=cores)[2] <= fn(probe, threads=cores // 2)[2] else cores // 2
    out = fn(ws, threads=threads)
    secs = out[2]
    res = {"value": n_sample / secs, "unit": "windows/s", "cores": threads, "kind": kind,
           "sample": "first %d windows of the workload, %d host threads (of %d hardware threads), consensus loop "
                     "only (%.1f s)" % (n_sample, threads, cores, secs)}
    res["gpu_reference"] = gpu_reference(args)
    return res


def gpu_reference(args):
    """The reference's OWN GPU path (GenomeWorks cudapoa, unmodified, oracle/_ref/libref_cudapoa.so) on the same
    workload and the same GPU — SURVEY §8(d)'s GPU baseline.  Separate process with a time limit: a failure inside the
    reference library must not take the bench line down.  Timing only (cudapoa's consensus is not spoa's)."""
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    try:
        p = subprocess.run([sys.executable, "-m", "oracle.cudapoa_time", "--windows", str(min(args.windows, 10000))],
                           cwd=here, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=240, text=True)
        lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
        return json.loads(lines[-1]) if lines else {"unavailable": "no output (exit %d)" % p.returncode}
    except Exception as e:  # noqa: BLE001 - reported, never fatal
        return {"unavailable": "%s" % e}


if __name__ == "__main__":
    main()
