#!/usr/bin/env python
"""bench.py — frames/sec of the per-frame Tessendorf hot path (spectrum + 3x IFFT + inversion + normals).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload default|c2|c3|c4|c5]

The JSON line's `value` is BASELINE.json configs[1] ("C2"): N=512 single patch, L=1000 m, wind 40 m/s, A=2, the
reference's PNG noise, a 600-frame time sweep t_f = f/60. One STEP = one full 600-frame sweep; every frame writes
its complete dy/dx/dz/normal set to HBM. `value` = frames/s with inputs resident in HBM (CUDA events on the launching
stream); `e2e` = the same sweep through the C ABI with HOST buffers (noise upload + h0 init + every frame's outputs
copied back to pinned host memory inside the timed region).

The default run (`--workload default`) ALSO measures the configurations the north star's targets are quoted on and puts
them in the same line under `configs`, each with its own CUDA-event timing, roofline and clock sample:
    configs.c3   N=2048 + Jacobian (the >= 0.70-of-HBM target); one replica per GPU
    configs.c4   64 cascades N=1024, SHARDED across the ranks under torchrun (the >= 7x-at-8-GPUs target)
    configs.c5   ONE N=32768 grid, slab-decomposed over the ranks (transpose over NVLink), checked against the single-GPU
                 path on a down-scaled grid of the same code before it is timed
`--workload c2|c3|c4|c5` runs one of them alone as the headline (development, profiling).

--impl reference times the reference's own compute shaders compiled for the CPU (oracle/_ref; the oracle port where that
library is absent) on ALL host cores of the box — also under torchrun, where rank 0 alone runs it — on a fixed sample of
the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "ocean frames/sec (spectrum+IFFT+normals)"
UNIT = "frames/s"
ALG_BYTES_PER_TEXEL = 44            # read h0k+h0minusk 16 B, write dy,dx,dz 12 B + normal 16 B (+4 Jacobian)
OWN_BYTES_PER_TEXEL = 36            # what THIS implementation must move: the folded spectrum is 8 B/texel, not 16
KERNEL_BYTES_PER_TEXEL = {          # compulsory bytes of each kernel of the 3-kernel frame
    "ow_row_kernel": 8 + 12,        # read the folded initial spectrum (16 B per texel PAIR) -> write Hermitian half-spectra (12)
    "ow_col_kernel": 12 + 12,       # read intermediate (12) -> write dy,dx,dz (12)
    "ow_normal_kernel": 4 + 16,     # read dy (4) -> write normal (16)
}
KERNELS = ["ow_row_kernel", "ow_col_kernel", "ow_normal_kernel"]

WORKLOADS = {
    # name: (N, frames per step, description, jacobian)
    "c2": (512, 600, "C2: N=512 single patch, 600-frame sweep t=f/60, L=1000 wind 40 A=2 lambda=1, PNG noise, dy/dx/dz+normal", False),
    "c3": (2048, 64, "C3: N=2048 single patch + Jacobian, 64-frame sweep t=f/60, L=1000 wind 40 A=2, rng(2048) noise", True),
    "c4": (1024, 64, "C4: 64 cascades N=1024 (L=100*1.08^c, wind 10+0.5c, dir 2*pi*c/64), one frame each at t=1", False),
    # C5: ONE N=32768 grid, slab-decomposed over the ranks (strong scaling). --c5-n 4096 runs the down-scaled grid
    # SURVEY.md §8 d2 names for the parity check (direct in-CTA lines instead of the N = A*B decomposition).
    "c5": (32768, 4, "C5: ONE grid N=32768, slab-decomposed 2-D IFFT over the ranks, Philox(32768) noise, "
                     "4-frame sweep t=f/60, dy/dx/dz+normal+Jacobian left column-slabbed", True),
}
# Fixed CPU-arm samples (the same at every --gpus N, so the driver's ratios compare like with like)
REFERENCE_SAMPLE = {"c2": 128, "c3": 4, "c4": 4, "c5": 8}


def host_cores() -> int:
    """Cores this process may run on. NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def workload_setup(name, only=None):
    """only = (lo, hi): build just that block of C4's cascades (a rank's shard)."""
    import fft_ocean_waves_b200 as fow
    N, frames, desc, jac = WORKLOADS[name]
    if name == "c4":
        lo, hi = only if only else (0, 64)
        casc, noise = [], []
        for c in range(lo, hi):
            ang = 2 * np.pi * c / 64
            casc.append(fow.OceanParams(L=float(100.0 * 1.08 ** c), wind_speed=float(10 + 0.5 * c),
                                        wind_dir=(float(np.cos(ang)), float(np.sin(ang))), amplitude=2.0, suppression=0.1, choppiness=1.0))
            noise.append(np.random.default_rng(1024 + c).integers(0, 256, (4, N, N), dtype=np.uint8))
        cascade_of = list(range(hi - lo))
        times = [1.0] * (hi - lo)
        frames = hi - lo
    else:
        casc = [fow.OceanParams(L=1000.0, wind_speed=40.0, wind_dir=(1.0, 1.0), amplitude=2.0, suppression=0.1, choppiness=1.0)]
        if name == "c5":
            noise = [_philox_noise(32768, N)]
        else:
            noise = [fow.default_noise() if name == "c2" else np.random.default_rng(2048).integers(0, 256, (4, N, N), dtype=np.uint8)]
        cascade_of = [0] * frames
        times = [float(np.float32(f / 60.0)) for f in range(frames)]
    return dict(name=name, N=N, frames=frames, desc=desc, jacobian=jac, cascades=casc, noise=noise, cascade_of=cascade_of, times=times)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": float(max(power))}


# ======================================================================================================================
# CPU arm
# ======================================================================================================================
def run_reference(args):
    """CPU arm: the reference's own shaders compiled for the CPU (oracle/_ref), else the oracle port, on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.oracle import OracleSim
    from oracle import ref as refmod
    name = "c2" if args.workload == "default" else args.workload
    if name == "c5":
        WORKLOADS["c5"] = (min(args.c5_n, 4096), 32, WORKLOADS["c5"][2].replace("N=32768", f"N={min(args.c5_n, 4096)} (CPU arm: down-scaled, the "
                           "oracle's reference textures need 116 GB at N=32768)"), True)
    w = workload_setup(name, only=(0, REFERENCE_SAMPLE["c4"]) if name == "c4" else None)
    N, cores = w["N"], host_cores()
    p = w["cascades"][0]
    # oracle/_ref = the reference's own compute shaders compiled for the CPU (oracle/make_ref.py): the reference arm proper. It needs
    # the reference's integer patch size and has no Jacobian (neither has the reference); otherwise the oracle port stands in.
    use_ref = refmod.available() and all(float(q.L) == int(q.L) for q in w["cascades"])
    sims = []
    for i, q in enumerate(w["cascades"]):
        if use_ref:
            sims.append(refmod.RefSim(N, int(q.L), q.wind_speed, q.wind_dir, q.amplitude, q.suppression, w["noise"][i], threads=cores))
        else:
            sims.append(OracleSim(N, q.L, q.wind_speed, q.wind_dir, q.amplitude, q.suppression, w["noise"][i], threads=cores))
    if use_ref:
        kind = "reference"
        what = ("oracle/_ref: the reference's own GLSL compute shaders (tilde_h0_t, butterfly x 2 log2 N x 3, inversion x 3, normal_map) compiled "
                "as C++ and dispatched in the reference's order on the host cores (OpenMP over rows); no Jacobian, as in the reference")
    else:
        kind = "port"
        what = "CPU oracle = scalar C++ restatement of the reference's GLSL dispatch chain (oracle/_ref not available or not applicable here)"
    lam = 1.0 if (w["jacobian"] and not use_ref) else None

    def one(f):
        s = sims[w["cascade_of"][f % len(w["cascade_of"])] % len(sims)]
        t = w["times"][f % len(w["times"])]
        return s.frame(t) if use_ref else s.frame(t, choppiness=lam)

    sample = min(REFERENCE_SAMPLE[name], w["frames"])
    one(0)                                                     # first call pays thread start-up and page faults

    def step():
        for f in range(sample):
            one(f)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    fps = sample * args.steps / dt
    desc = f"first {sample} frames of the workload's {WORKLOADS[name][1]} per step (fixed sample, the same at every --gpus)"
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "N": N, "frames_per_step": sample, "sample": desc, "what": what,
                       "threads": cores, "OMP_NUM_THREADS_env": os.environ.get("OMP_NUM_THREADS")},
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if use_ref:
        # for transparency: the hand-written restatement (oracle/ow_oracle.cpp) is faster than the reference's shader text compiled through
        # the GLSL emulation layer; its throughput on the same frames, same cores, is reported beside the reference arm's
        q = w["cascades"][0]
        port = OracleSim(N, q.L, q.wind_speed, q.wind_dir, q.amplitude, q.suppression, w["noise"][0], threads=cores)
        port.frame(w["times"][0])
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < 2.0 or n < 3:
            port.frame(w["times"][n % len(w["times"])])
            n += 1
        line["cpu_port"] = {"value": n / (time.perf_counter() - t0), "unit": UNIT, "cores": cores, "kind": "port",
                            "what": "oracle/ow_oracle.cpp on the same workload (not the number the driver compares against)"}
    print(json.dumps(line), flush=True)


def cpu_baseline(w, seconds=10.0):
    from oracle.oracle import OracleSim
    cores = host_cores()
    p = w["cascades"][0]
    sim = OracleSim(w["N"], p.L, p.wind_speed, p.wind_dir, p.amplitude, p.suppression, w["noise"][0], threads=cores)
    lam = 1.0 if w["jacobian"] else None
    sim.frame(0.0, choppiness=lam)   # warm-up
    n, t0 = 0, time.perf_counter()
    while True:
        sim.frame(w["times"][n % len(w["times"])], choppiness=lam)
        n += 1
        dt = time.perf_counter() - t0
        if (dt > seconds and n >= 3) or n >= w["frames"]:
            break
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {n} frames of the sweep ({dt:.1f} s), oracle/ow_oracle.cpp with OpenMP over rows on {cores} threads"}


def _philox_noise(seed, N):
    from oracle.numpy_ref import philox_noise
    return philox_noise(seed, N)


# ======================================================================================================================
# GPU arm plumbing
# ======================================================================================================================
class Env:
    """One process per GPU: torch for streams/events/pinned memory, torch.distributed (NCCL) for the barrier and the slab exchange."""

    def __init__(self):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        # A dedicated non-default stream: its handle is what the C ABI launches on, and the torch events are recorded
        # on the same stream (handle 0 would mean "the context's own stream" to ow_step*).
        self.stream = torch.cuda.Stream()
        torch.cuda.set_stream(self.stream)
        self.sp = self.stream.cuda_stream
        assert self.sp != 0
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
        self.flush_src = torch.zeros(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        if self.dist is None:
            return [float(v) for v in values]
        tt = self.torch.tensor(list(values), device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(tt, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in tt.cpu().numpy()]

    def event(self):
        return self.torch.cuda.Event(enable_timing=True)

    def flush_l2(self):
        """L2 flush between timed iterations, outside the event pair: write a 256 MiB buffer (> the 126 MB L2), then READ another one,
        so that the L2 is left full of clean lines - the write-back of the flush's own dirty lines is not billed to the timed step."""
        self.flush.zero_()
        self.flush_src.view(self.torch.int64).sum()

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()


def traffic_per_launch(name, kernel, frames_per_launch):
    """dram bytes per launch of `kernel` = this round's ncu capture (bytes per FRAME, profiles/traffic.json) x frames per launch."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        per_frame = t["per_frame"].get(f"{name}:{kernel}")
        return None if per_frame is None else float(per_frame) * frames_per_launch
    except Exception:
        return None


# ======================================================================================================================
# Independent patches / cascades (C2, C3, C4)
# ======================================================================================================================
def measure_patch(env, args, name, steps, warmup, full):
    """One JSON-able record for workload `name`. full=True adds e2e, the single-frame (drop-in) numbers, the cuFFT comparison
    and the CPU baseline (rank 0, world 1)."""
    import fft_ocean_waves_b200 as fow
    torch = env.torch
    world, rank = env.world, env.rank
    # C4 (BASELINE.json configs[3]): the 64 cascades are SHARDED, a contiguous block of 64/world per GPU,
    # no data-path collective -> strong scaling (total work fixed). Every other workload replicates per GPU (weak).
    sharded = name == "c4" and (world > 1 or args.c4_shard_of > 1)
    only = None
    if sharded:
        parts = world if world > 1 else args.c4_shard_of       # --c4-shard-of P: time ONE GPU's share of a P-GPU run (tuning aid)
        if 64 % parts:
            raise SystemExit("bench.py: c4 shards 64 cascades; --gpus must divide 64")
        per = 64 // parts
        only = (rank * per, rank * per + per)
    w = workload_setup(name, only=only)
    N, frames = w["N"], w["frames"]
    job_frames = (64 if world > 1 else frames) if sharded else world * frames            # frames the WHOLE job produces per step
    slots = min(args.slots or (300 if N <= 512 else 32), frames) if name != "c4" else frames
    sim = fow.FFTOceanWaves(N=N, cascades=w["cascades"], n_slots=max(slots, len(w["cascades"])), device=env.local, jacobian=w["jacobian"],
                            fused_normals=args.fused_normals)
    for i, nz in enumerate(w["noise"]):
        sim.set_noise(nz, cascade=i)
    sim.tilde_h0_k()
    if args.group:
        sim.set_group_size(args.group)
    if args.streams:
        sim.set_streams(args.streams)
    if args.row_kernel:
        sim.set_row_kernel(args.row_kernel)
    if args.col_kernel or args.fused >= 0:
        sim.set_column_kernel(args.col_kernel, args.fused)
    if args.discard:
        sim.set_discard_intermediate(True)
    sp, stream = env.sp, env.stream
    launches, groups = [0], [0]

    # C4 evaluates every cascade once at ONE time: that is ow_step (slot i <- cascade i at t), the drop-in call, which submits the whole
    # step as one CUDA graph launch; the time sweeps go through ow_step_multi
    whole_step = name == "c4" and len(set(w["times"])) == 1 and list(w["cascade_of"]) == list(range(frames)) and not args.no_graph

    def sweep():
        if whole_step:
            sim.update(w["times"][0], stream=sp)
            launches[0] += sim.last_launch_count()
            groups[0] += sim.last_group_count()
            return
        for base in range(0, frames, slots):
            n = min(slots, frames - base)
            sim.update_multi(w["cascade_of"][base:base + n], w["times"][base:base + n], stream=sp)
            launches[0] += sim.last_launch_count()
            groups[0] += sim.last_group_count()

    for _ in range(warmup):
        env.flush_l2()
        sweep()
    env.barrier()
    sampler = ClockSampler(env.local).start() if rank == 0 else None
    launches[0] = groups[0] = 0
    evs = []
    t_wall = time.perf_counter()
    for _ in range(steps):
        env.flush_l2()                      # L2 flush between timed iterations (outside the event pair)
        a, b = env.event(), env.event()
        a.record(stream)
        sweep()
        b.record(stream)
        evs.append((a, b))
    env.barrier()
    t_wall = time.perf_counter() - t_wall
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    timed_launches = launches[0]
    total_ms = env.max_over_ranks([total_ms])[0]
    value = job_frames * steps / (total_ms * 1e-3)

    if args.profile:
        sim.close()
        return None
    # ---- per-kernel durations (CUDA events around each kernel, same stream, same workload) -> roofline ----
    kms = np.zeros(3)
    prof_sweeps = 2
    for _ in range(prof_sweeps):
        env.flush_l2()
        for base in range(0, frames, slots):
            n = min(slots, frames - base)
            kms += np.array(sim.update_multi_timed(w["cascade_of"][base:base + n], w["times"][base:base + n], stream=sp))
    clocks = sampler.stop() if sampler else None
    kms /= prof_sweeps                                      # ms per sweep per kernel
    groups_per_sweep = groups[0] / steps
    modes = sim.kernel_modes()
    fused = modes["fused"]                                  # normal map produced by the column kernel's epilogue; the third slot is then the seam (+ Jacobian) pass
    peak, peak_src = measured_peak()
    texels = float(N) * N
    kb = dict(KERNEL_BYTES_PER_TEXEL)
    if w["jacobian"]:
        kb["ow_normal_kernel"] += 8 + 4                     # + read dx,dz, write J
    if fused:
        kb["ow_col_kernel"] += 12                           # the fused column kernel also writes 3/4 of the normal map and never re-reads dy
        kb["ow_normal_kernel"] = 4 + 4 + (8 + 4 if w["jacobian"] else 0)   # seam quads (a quarter of the normals) + the Jacobian pass
    per_kernel = []
    for i, k in enumerate(KERNELS):
        if kms[i] == 0.0:
            continue
        bpt = kb[k]
        if fused and k == "ow_col_kernel":
            k = "ow_col_fused_kernel"
        gbs = bpt * texels * frames / (kms[i] * 1e-3) / 1e9
        per_kernel.append({"kernel": k, "ms_per_launch": kms[i] / groups_per_sweep, "share": kms[i] / kms.sum(),
                           "bytes_per_texel": bpt, "achieved_gbs": gbs, "frac": gbs / peak})
    dom = int(np.argmax([pk["share"] for pk in per_kernel]))
    frames_per_launch = frames / groups_per_sweep
    alg = ALG_BYTES_PER_TEXEL + (4 if w["jacobian"] else 0)
    own = OWN_BYTES_PER_TEXEL + (4 if w["jacobian"] else 0)
    frame_gbs = alg * texels * value / world / 1e9
    own_gbs = own * texels * value / world / 1e9
    roofline = {"bound": "hbm", "kernel": per_kernel[dom]["kernel"], "achieved": per_kernel[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": per_kernel[dom]["frac"], "traffic": traffic_per_launch(name, per_kernel[dom]["kernel"], frames_per_launch),
                "peak_source": peak_src,
                "bytes_per_launch": per_kernel[dom]["bytes_per_texel"] * texels * frames_per_launch,
                "kernels": per_kernel,
                "frame": {"algorithmic_bytes_per_texel": alg, "achieved": frame_gbs, "frac": frame_gbs / peak,
                          "own_compulsory_bytes_per_texel": own, "own_achieved": own_gbs, "own_frac": own_gbs / peak,
                          "note": "whole frame per GPU. frac: on SURVEY.md §8 d4's 44 B/texel (48 with Jacobian), the figure the target is quoted on; "
                                  "own_frac: on the bytes THIS implementation must move (36 / 40: the folded spectrum is 8 B/texel). The separate-kernel "
                                  "frame moves 64 (76) B/texel through L2"}}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "N": N, "frames_per_step": job_frames if sharded else frames, "slots_per_launch": slots,
                       "launch_groups_per_step": groups_per_sweep,
                       "l2": "flushed between timed steps (256 MiB memset, then a 256 MiB read so the L2 is left clean; both outside the event pair); within a step the outputs "
                             f"({frames * sim.frame_bytes() / 1e9:.2f} GB) stream through L2, the folded spectrum ({8 * texels * len(w['cascades']) / 1e6:.1f} MB) is re-read every frame",
                       "parallelism": (f"64 cascades sharded {frames} per GPU over {world} GPUs, no communication" if sharded
                                       else f"{world} x independent patch per GPU, no communication"),
                       "kernels": modes, "wall_s_timed_region": t_wall,
                       "submission": "ow_step: one CUDA graph launch per step" if whole_step else "ow_step_multi: stream launches, launch groups on the context's auxiliary streams"},
            "clocks": clocks, "gpu_launches": int(timed_launches), "roofline": roofline}

    if name == "c4":
        try:
            line["config"]["compose"] = compose_numbers(env, sim, frames)
        except Exception as e:  # noqa: BLE001 - a side record must never take the bench line down
            line["config"]["compose"] = {"unavailable": f"{type(e).__name__}: {e}"}
    # ---- the drop-in call itself: ONE frame per ow_step (what FFTOceanWaves::update() does, src/main.cpp:240-244) ---------------
    if name != "c4":
        line["config"]["single_frame"] = single_frame_numbers(env, sim, w)
        line["config"]["single_slot_sequential_fps"] = line["config"]["single_frame"]["graph"]["back_to_back_fps"]
    if not full:
        sim.close()
        return line

    # ---- end to end through the C ABI with HOST buffers --------------------------------------------------
    fbytes = sim.frame_bytes()
    host = torch.empty(slots * fbytes, dtype=torch.uint8, pin_memory=True)
    pinned_noise = [torch.from_numpy(np.ascontiguousarray(nz)).pin_memory() for nz in w["noise"]]
    hp = host.data_ptr()

    def e2e_step():
        for i, nz in enumerate(pinned_noise):                  # H2D: the step's inputs (noise planes) from pinned memory
            sim.set_noise(nz.numpy(), cascade=i)
        sim.tilde_h0_k()
        for base in range(0, frames, slots):
            n = min(slots, frames - base)
            sim.update_multi(w["cascade_of"][base:base + n], w["times"][base:base + n], stream=sp)
            for s in range(n):                                 # D2H: every frame's dy,dx,dz,normal to pinned host memory
                sim.download_frame_async(s, hp + s * fbytes, fbytes, stream=sp)
        sim.sync(stream=sp)
        return float(host[:4].view(torch.float32)[0])          # host-side read of the result

    e2e_step()
    env.barrier()
    e2e_steps = max(1, min(steps, 3))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    env.barrier()
    e2e_t = env.max_over_ranks([time.perf_counter() - t0])[0]
    line["e2e"] = {"value": job_frames * e2e_steps / e2e_t, "unit": UNIT,
                   "h2d_bytes_per_step": int(sum(nz.numel() for nz in pinned_noise)), "d2h_bytes_per_step": int(frames * fbytes),
                   "steps": e2e_steps,
                   "what": "ow_set_noise + ow_init_spectrum + ow_step_multi + ow_download_frame_async of every frame into pinned host memory "
                           "(the reference's texture formats: 28 B/texel, +4 with the Jacobian)"}
    del host
    cmp_rec = None
    if rank == 0 and not args.no_compare and name != "c4":
        try:
            cmp_rec = compare_cufft(env, sim, w)
        except Exception as e:  # noqa: BLE001 - a comparison point must never take the bench line down
            cmp_rec = {"cufft": {"unavailable": f"{type(e).__name__}: {e}"}}
    sim.close()
    # ---- the same end-to-end sweep with the packed output set: 12 B/texel over PCIe instead of 28 ----
    if name != "c4":
        line["e2e_packed_f16"] = e2e_packed(env, fow, w, args, slots, pinned_noise, job_frames, steps)
    if cmp_rec is not None:
        line["comparison"] = cmp_rec
    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(w)
    return line


def compose_numbers(env, sim, n_slots, M=2048, extent=200.0, max_terms=8):
    """ow_compose_grid - the consumer's displacement (grid_tes.glsl:60-64) summed over the first cascades of this rank with
    unit blending weights, on an M x M clip-map image. CUDA events on the launching stream, L2 flushed before every repetition."""
    import ctypes as C
    torch = env.torch
    n = min(max_terms, n_slots)
    terms = sim._terms(list(range(n)), [1.0] * n)
    off = torch.empty((M, M, 4), dtype=torch.float32, device="cuda")
    nrm = torch.empty((M, M, 4), dtype=torch.float32, device="cuda")
    ms = []
    for i in range(6):
        env.flush_l2()
        a, b = env.event(), env.event()
        a.record(env.stream)
        rc = sim._lib.ow_compose_grid(sim._h, n, terms, 0.5, M, -100.0, -100.0, float(extent), off.data_ptr(), nrm.data_ptr(), C.c_void_p(env.sp))
        b.record(env.stream)
        if rc != 0:
            raise RuntimeError("ow_compose_grid failed")
        torch.cuda.synchronize()
        if i:
            ms.append(a.elapsed_time(b))
    t = float(np.median(ms)) * 1e-3
    src_bytes = n * 28.0 * sim.N * sim.N          # every source texel of every term at most once from DRAM (dy,dx,dz 12 B + normal 16 B)
    return {"what": f"ow_compose_grid: {n} cascades blended into one {M}x{M} offset + normal image over {extent:.0f} m", "terms": n, "M": M,
            "ms": t * 1e3, "output_mtexels_per_s": M * M / t / 1e6,
            "algorithmic_gbs": (32.0 * M * M + min(src_bytes, n * 4 * 28.0 * M * M)) / t / 1e9,
            "bytes_note": "32 B written per output texel + each term's source texels once (28 B each)"}


def single_frame_numbers(env, sim, w):
    """ow_step(t), one frame per call into slot 0 — the call that replaces the reference's update() chain. With the CUDA
    graph (default) and with plain launches: frames/s back to back (no host sync between frames: throughput of the call path) and
    the latency of one synchronous frame (ow_step + ow_sync, host clock). The loop calls the C ABI through pre-built ctypes
    arguments, so what is timed is the library and the GPU, not Python argument marshalling."""
    import ctypes as C
    torch = env.torch
    out = {}
    nseq = min(len(w["times"]), 200)
    lib, h, stp = sim._lib, sim._h, C.c_void_p(env.sp)
    ts = [C.c_float(t) for t in w["times"][:max(nseq, 50)]]
    step, sync = lib.ow_step, lib.ow_sync
    for label, on in (("graph", True), ("launches", False)):
        sim.set_graph(on)
        for f in range(10):
            assert step(h, ts[f % len(ts)], stp) == 0
        torch.cuda.synchronize()
        a, b = env.event(), env.event()
        a.record(env.stream)
        t0 = time.perf_counter()
        for f in range(nseq):
            step(h, ts[f], stp)
        host_us = (time.perf_counter() - t0) * 1e6 / nseq
        b.record(env.stream)
        torch.cuda.synchronize()
        fps = nseq / (a.elapsed_time(b) * 1e-3)
        lat = []
        for f in range(50):
            t0 = time.perf_counter()
            step(h, ts[f % len(ts)], stp)
            sync(h, stp)
            lat.append(time.perf_counter() - t0)
        out[label] = {"back_to_back_fps": fps, "sync_latency_us_median": float(np.median(lat) * 1e6), "host_us_per_call": host_us, "frames": nseq}
    sim.set_graph(True)
    out["what"] = ("ow_step(ctx, t, stream): slot 0 <- cascade 0 at time t, one call per frame (the reference's update(), src/main.cpp:240-244). "
                   "Consecutive frames write the same output slot, so they run one after the other on the GPU: the rate is set by the "
                   "latency of the frame's dependent kernels, not by throughput")
    return out


def e2e_packed(env, fow, w, args, slots, pinned_noise, job_frames, steps):
    """End to end with OW_FLAG_PACKED_F16: RGBA16F (dx,dy,dz,J) + RG16_SNORM normal.xz copied back instead of the reference formats."""
    torch = env.torch
    N, frames = w["N"], w["frames"]
    sim = fow.FFTOceanWaves(N=N, cascades=w["cascades"], n_slots=max(slots, len(w["cascades"])), device=env.local, jacobian=w["jacobian"], packed="f16")
    pbytes = sim.packed_bytes()
    host = torch.empty(slots * pbytes, dtype=torch.uint8, pin_memory=True)
    hp = host.data_ptr()

    def step():
        for i, nz in enumerate(pinned_noise):
            sim.set_noise(nz.numpy(), cascade=i)
        sim.tilde_h0_k()
        for base in range(0, frames, slots):
            n = min(slots, frames - base)
            sim.update_multi(w["cascade_of"][base:base + n], w["times"][base:base + n], stream=env.sp)
            for s in range(n):
                sim.download_packed_async(s, hp + s * pbytes, pbytes, stream=env.sp)
        sim.sync(stream=env.sp)
        return float(host[:2].view(torch.float16)[0])

    step()
    env.barrier()
    n = max(1, min(steps, 3))
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    env.barrier()
    dt = env.max_over_ranks([time.perf_counter() - t0])[0]
    sim.close()
    return {"value": job_frames * n / dt, "unit": UNIT, "h2d_bytes_per_step": int(sum(nz.numel() for nz in pinned_noise)),
            "d2h_bytes_per_step": int(frames * pbytes), "steps": n,
            "what": "same sweep, context created with OW_FLAG_PACKED_F16: ow_download_packed_async of every frame (12 B/texel: RGBA16F dx,dy,dz,J + "
                    "RG16_SNORM normal.xz; decode in INTEGRATION.md; tolerance 1e-3 of peak, tests/test_gpu_packed.py)"}


def compare_cufft(env, sim, w):
    """Comparison point only (never the product): the same frame through cuFFT (torch.fft.ifft2 = batched 2-D C2C inverse) on the
    identical spectra, plus torch element-wise ops for the inversion and the normal map. Also reports how far the cuFFT result is
    from this library's output (an independent check of the hand-written FFT)."""
    torch = env.torch
    N = w["N"]
    p = w["cascades"][0]
    dev = "cuda"
    h0k = torch.from_numpy(sim.download("h0k", 0)).to(dev)
    h0m = torch.from_numpy(sim.download("h0minusk", 0)).to(dev)
    h0k = torch.complex(h0k[..., 0], h0k[..., 1])
    h0m = torch.complex(h0m[..., 0], h0m[..., 1])
    idx = torch.arange(N, device=dev, dtype=torch.float32) - N / 2.0
    k1 = (2.0 * np.float32(np.pi) * idx) / np.float32(p.L)             # tilde_h0_t_cs.glsl:72-73
    kx, ky = k1[None, :].expand(N, N), k1[:, None].expand(N, N)
    km = torch.sqrt(kx * kx + ky * ky).clamp_min(1e-5)                  # :74-79
    wdisp = torch.sqrt(9.81 * km)

    def spectra(t):
        ph = wdisp * np.float32(t)
        e = torch.complex(torch.cos(ph), torch.sin(ph))
        h = h0k * e + h0m * torch.conj(e)                                  # :96-110 (conjugate() is a no-op in the shader: h0minusk is NOT conjugated)
        mi = torch.complex(torch.zeros_like(km), -torch.ones_like(km))
        return torch.stack([h, mi * (kx / km) * h, mi * (ky / km) * h])   # dy, dx, dz  (:113-126)

    def ifft_frame(H):
        D = torch.fft.ifft2(torch.fft.ifftshift(H, dim=(-2, -1)))         # = butterfly passes + inversion_cs.glsl
        return D.real.contiguous()

    def normals(h):                                                       # normal_map_cs.glsl:24-54; LINEAR/REPEAT corner taps = 2x2 box means
        box = 0.25 * (h + torch.roll(h, 1, 0) + torch.roll(h, 1, 1) + torch.roll(h, (1, 1), (0, 1)))
        z = lambda dx, dy: torch.roll(box, (-dy, -dx), (0, 1))
        nz = z(-1, -1) + 2 * z(0, -1) + z(1, -1) - z(-1, 1) - 2 * z(0, 1) - z(1, 1)
        nx = z(-1, -1) + 2 * z(-1, 0) + z(-1, 1) - z(1, -1) - 2 * z(1, 0) - z(1, 1)
        r = torch.rsqrt(nx * nx + 1.0 + nz * nz)
        return torch.stack([nx * r, r, nz * r], dim=-1)

    t = w["times"][min(1, len(w["times"]) - 1)]
    H = spectra(t)
    D = ifft_frame(H)
    ours = sim.frame(t)
    err = {k: float(np.abs(D[i].cpu().numpy() - ours[k]).max() / np.abs(ours[k]).max()) for i, k in enumerate(("dy", "dx", "dz"))}
    nrm = normals(D[0])
    err["normal_abs"] = float(np.abs(nrm.cpu().numpy() - ours["normal"][..., :3]).max())
    Hs = torch.fft.ifftshift(H, dim=(-2, -1)).contiguous()
    reps = 50 if N <= 1024 else 20

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = env.event(), env.event()
        a.record(env.stream)
        for _ in range(reps):
            fn()
        b.record(env.stream)
        torch.cuda.synchronize()
        return reps / (a.elapsed_time(b) * 1e-3)

    fft_only = timed(lambda: torch.fft.ifft2(Hs))
    pipeline = timed(lambda: normals(torch.fft.ifft2(Hs).real[0]))
    full = timed(lambda: normals(ifft_frame(spectra(t))[0]))
    return {"cufft": {"ifft2_only_fps": fft_only, "ifft2_plus_normals_fps": pipeline, "spectrum_ifft2_normals_fps": full, "unit": UNIT,
                      "max_rel_diff_vs_ours": err,
                      "what": "torch.fft.ifft2 (cuFFT batched 2-D C2C inverse, 3 x NxN complex64) on the identical spectra; +normals = .real + the "
                              "normal map in torch element-wise ops; spectrum_... also evaluates h(k,t) in torch ops. Comparison point only; back to "
                              "back on one stream, spectra resident in HBM (and in L2 for small N)"}}


# ======================================================================================================================
# One grid over all ranks (C5)
# ======================================================================================================================
def slab_parity_check(env, fow, N, transport):
    """Before timing: the slab path on all ranks against the single-GPU path, one frame, on a grid the single-GPU context fits
    beside the slab (same code: line decomposition when N > 4096, the transpose, the halo columns)."""
    p = fow.OceanParams(L=1000.0, wind_speed=40.0, wind_dir=(1.0, 1.0), amplitude=2.0, suppression=0.1, choppiness=1.0)
    t = 1.0
    slab = fow.SlabOcean(N=N, params=p, device=env.local, jacobian=True, transport=transport)
    slab.init(32768)
    for tt in (0.25, 0.5, 0.75, t):          # four frames: the pipelined path (rows of f+1 during the columns of f) reuses both receive buffers
        slab.update(tt)
    slab.sync()
    got = {k: slab.gather(k) for k in ("dy", "dx", "dz", "jacobian")}
    slab.close()
    res = None
    if env.rank == 0:
        with fow.FFTOceanWaves(N=N, cascades=[p], jacobian=True, device=env.local) as one:
            one.set_noise_seed(32768)
            one.tilde_h0_k()
            ref = one.frame(t)
        res = {"N": N, "world": env.world}
        for k in ("dy", "dx", "dz"):
            res[k + "_max_rel"] = float(np.abs(got[k] - ref[k]).max() / np.abs(ref[k]).max())
        res["jacobian_max_abs"] = float(np.abs(got["jacobian"] - ref["jacobian"]).max())
        res["ok"] = bool(max(res["dy_max_rel"], res["dx_max_rel"], res["dz_max_rel"]) <= 1e-4 and res["jacobian_max_abs"] <= 1e-3)
    env.barrier()
    return res


def measure_slab(env, args, n_grid, steps, warmup, full):
    """One grid over all ranks (strong scaling). A step = one sweep of SlabOcean.update."""
    import fft_ocean_waves_b200 as fow
    torch = env.torch
    world, rank, dist = env.world, env.rank, env.dist
    N, frames, desc, jac = WORKLOADS["c5"]
    if n_grid != N:
        N, frames = n_grid, (32 if n_grid <= 4096 else 4)
        desc = desc.replace("N=32768", f"N={N} (down-scaled)").replace("4-frame", f"{frames}-frame")
    parity = None
    if not args.profile and not args.no_slab_check:
        parity = slab_parity_check(env, fow, min(N, 8192), args.transport)
        if rank == 0 and not parity["ok"]:
            raise SystemExit(f"bench.py: slab path disagrees with the single-GPU path: {parity}")
    p = fow.OceanParams(L=1000.0, wind_speed=40.0, wind_dir=(1.0, 1.0), amplitude=2.0, suppression=0.1, choppiness=1.0)
    times = [float(np.float32(f / 60.0)) for f in range(frames)]
    sim = fow.SlabOcean(N=N, params=p, device=env.local, jacobian=jac, transport=args.transport, pipeline=(False if args.no_pipeline else None))
    if args.col_lines and hasattr(sim.backend, "set_column_lines"):
        sim.backend.set_column_lines(args.col_lines)
    if args.post_ctas >= 0 and hasattr(sim.backend, "set_post_ctas"):
        sim.backend.set_post_ctas(args.post_ctas)
    if args.line_clusters != 0 and hasattr(sim.backend, "set_line_clusters"):
        sim.backend.set_line_clusters(args.line_clusters)
    sim.init(32768)
    stream = env.stream

    def sweep():
        for t in times:
            sim.update(t)
        sim.flush()                 # pipelined frames: the caller's stream (and the event recorded on it next) waits for all of them

    for _ in range(warmup):
        env.flush_l2()
        sweep()
    env.barrier()
    sampler = ClockSampler(env.local).start() if rank == 0 else None
    evs = []
    for _ in range(steps):
        env.flush_l2()
        a, b = env.event(), env.event()
        a.record(stream)
        sweep()
        b.record(stream)
        evs.append((a, b))
    env.barrier()
    total_ms = env.max_over_ranks([sum(a.elapsed_time(b) for a, b in evs)])[0]
    value = frames * steps / (total_ms * 1e-3)
    if args.profile:
        sim.close()
        return None
    # per-phase durations: events around ow_slab_rows (1 kernel) and ow_slab_cols (column + normal kernels)
    b_ = sim.backend
    st = b_.current_stream()
    kms = np.zeros(3)
    for t in times:
        if sim.transport == "peer":
            sim._barrier()
        e = [env.event() for _ in range(4)]
        e[0].record(stream)
        b_.rows(t, 1 if sim.transport == "peer" else 0, st)
        e[1].record(stream)
        if sim.transport == "peer":
            sim._barrier()
        elif world == 1:
            b_.local_exchange(st)
        else:
            send, recv = b_.exchange_tensors()
            dist.all_to_all_single(recv, send)
        e[2].record(stream)
        b_.cols(st)
        e[3].record(stream)
        torch.cuda.synchronize()
        kms += np.array([e[0].elapsed_time(e[1]), e[2].elapsed_time(e[3]), e[1].elapsed_time(e[2])])
    kms = np.array(env.max_over_ranks(kms))
    clocks = sampler.stop() if sampler else None
    peak, peak_src = measured_peak()
    texels_rank = float(N) * N / world
    # compulsory bytes per texel of the two phases (folded spectrum 8 -> intermediate 12; intermediate 12 -> dy,dx,dz 12, then
    # dy,dx,dz 12 -> normal 16 + J 4); above N=4096 the line decomposition's scratch adds 24 B/texel per direction, not counted
    kb = {"ow_row_slab_kernel": 8 + 12, "ow_col_slab_kernel+ow_normal_slab_kernel": 12 + 12 + 4 + 8 + 16 + 4}
    per_kernel = []
    for i, k in enumerate(kb):
        gbs = kb[k] * texels_rank * frames / (kms[i] * 1e-3) / 1e9
        per_kernel.append({"kernel": k, "ms_per_launch": kms[i] / frames, "share": kms[i] / kms[:2].sum(), "bytes_per_texel": kb[k],
                           "achieved_gbs": gbs, "frac": gbs / peak})
    dom = int(np.argmax(kms[:2]))
    frame_gbs = 48 * texels_rank * value / 1e9
    xbytes = sim.exchange_bytes_per_frame()
    # NVLink: with peer stores the exchange happens INSIDE the row phase, so its bandwidth is bytes / row-phase time
    row_s = kms[0] * 1e-3 / frames
    nv_during = xbytes / row_s / 1e9 if sim.transport == "peer" and world > 1 else (xbytes / (kms[2] * 1e-3 / frames) / 1e9 if world > 1 and kms[2] > 0 else 0.0)
    roofline = {"bound": "hbm", "kernel": per_kernel[dom]["kernel"], "achieved": per_kernel[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": per_kernel[dom]["frac"], "traffic": None, "peak_source": peak_src,
                "bytes_per_launch": per_kernel[dom]["bytes_per_texel"] * texels_rank, "kernels": per_kernel,
                "frame": {"algorithmic_bytes_per_texel": 48, "achieved": frame_gbs, "frac": frame_gbs / peak,
                          "note": "per GPU: 48 B/texel x N^2/world texels x frames/s"},
                "nvlink": {"bytes_per_frame_per_gpu_per_direction": xbytes, "achieved_gbs_over_frame": xbytes * value / 1e9,
                           "achieved_gbs_during_exchange": nv_during, "peak_gbs": 770.0, "frac": nv_during / 770.0,
                           "exchange_ms_per_frame": (kms[0] if sim.transport == "peer" else kms[2]) / frames,
                           "note": "12 B/texel Hermitian-packed intermediate x (world-1)/world of this rank's texels (+ halo columns); "
                                   "frac = bytes / (time of the phase that moves them: the row kernel for peer stores, the all-to-all otherwise) against the "
                                   "measured peer copy 770 GB/s per direction (B200_PROFILING.md)"}}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "N": N, "frames_per_step": frames, "transport": sim.transport,
                       "l2": "flushed between timed steps (256 MiB memset, then a 256 MiB read so the L2 is left clean; both outside the event pair); a frame's working set "
                             f"({(16 + 12 + 12 + 20) * N * N / world / 1e6:.0f} MB per GPU) exceeds L2 at world <= 4",
                       "parallelism": f"slab{world}: row pairs -> transpose (peer stores / all-to-all over NVLink) -> column slabs",
                       "pipelined": bool(getattr(sim, "pipelined", False)),
                       "pipelined_note": "frame f+1's row pass (its peer stores are the exchange) overlaps frame f's column pass: two receive buffers, two streams",
                       "line_clusters": (sim.backend.line_clusters() if hasattr(sim.backend, "line_clusters") else 0),
                       "line_clusters_note": "N = A*B line decomposition: bit 0 rows, bit 1 columns run as thread-block clusters of A CTAs that combine "
                                             "their sub-lines through distributed shared memory (no global scratch), bit 2 = 8-column tiles; 0 = two kernels per direction",
                       "slab_vs_single_gpu_check": parity},
            "clocks": clocks, "gpu_launches": int(sim.launches_per_frame() * frames * steps), "roofline": roofline}
    if full:
        # ---- end to end: seed in, every frame's column slab copied to pinned host memory ---------------------------------
        outs = b_.output_tensors()
        host = {k: torch.empty(v.shape, dtype=torch.float32, pin_memory=True) for k, v in outs.items()}
        fbytes = sum(h.numel() * 4 for h in host.values())

        def e2e_step():
            sim.init(32768)
            for t in times:
                sim.update(t)
                sim.flush()                      # this frame's outputs are complete on the current stream before they are copied
                for k, v in outs.items():
                    host[k].copy_(v, non_blocking=True)
            torch.cuda.synchronize()
            return float(host["dy"][0, 0])

        e2e_step()
        env.barrier()
        e2e_steps = max(1, min(steps, 3))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        env.barrier()
        e2e_t = env.max_over_ranks([time.perf_counter() - t0])[0]
        line["e2e"] = {"value": frames * e2e_steps / e2e_t, "unit": UNIT, "h2d_bytes_per_step": 8, "d2h_bytes_per_step": int(frames * fbytes * world),
                       "steps": e2e_steps, "what": "ow_slab_init_spectrum_seeded (8-byte seed) + SlabOcean.update + every rank's column slab "
                                                   "(dy,dx,dz,normal,J) copied to pinned host memory every frame"}
        if rank == 0 and world == 1 and not args.no_cpu and N <= 4096:      # the CPU oracle's reference textures need 108 B/texel: 116 GB at N=32768
            w = dict(N=N, frames=frames, jacobian=True, cascades=[p], noise=[_philox_noise(32768, N)], times=times)
            line["cpu_baseline"] = cpu_baseline(w)
    sim.close()
    return line


def sub_record(line):
    """What a configs.* entry keeps of a full record."""
    if line is None:
        return None
    keep = ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "config", "clocks", "gpu_launches", "roofline", "e2e")
    return {k: line[k] for k in keep if k in line}


def run_ours(args):
    env = Env()
    name = args.workload
    line = None
    if name == "default":
        line = measure_patch(env, args, "c2", args.steps, args.warmup, full=True)
        if not args.profile:
            sub_steps = max(1, min(args.steps, 3))
            configs = {}
            for cname in ("c3", "c4"):
                try:
                    configs[cname] = sub_record(measure_patch(env, args, cname, sub_steps, 3, full=False))
                except Exception as e:  # noqa: BLE001 - a sub-record must not take the headline down; the failure is recorded instead
                    configs[cname] = {"error": f"{type(e).__name__}: {e}"}
                    if env.world > 1:
                        raise
            try:
                configs["c5"] = sub_record(measure_slab(env, args, args.c5_n, sub_steps, 3, full=False))
            except Exception as e:  # noqa: BLE001
                configs["c5"] = {"error": f"{type(e).__name__}: {e}"}
                if env.world > 1:
                    raise
            if line is not None:
                line["configs"] = configs
                line["config"]["workload_note"] = ("value/e2e/roofline = C2 (BASELINE.json configs[1]); configs.c3/c4/c5 = the target configurations "
                                                   "measured in the same run (c4 sharded and c5 slab-decomposed over the ranks under torchrun)")
    elif name == "c5":
        line = measure_slab(env, args, args.c5_n, args.steps, args.warmup, full=True)
    else:
        line = measure_patch(env, args, name, args.steps, args.warmup, full=True)
    env.close()
    if line is not None and env.rank == 0:
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=["default"] + sorted(WORKLOADS), default="default")
    ap.add_argument("--slots", type=int, default=0, help="frames evaluated per ow_step_multi call (0 = 300 for c2 (measured: 369 k frames/s vs 361 k at 128), 32 for c3, 64 for c4)")
    ap.add_argument("--group", type=int, default=0, help="slots per launch group (0 = library default)")
    ap.add_argument("--streams", type=int, default=0, help="internal streams the launch groups are spread over (0 = library default)")
    ap.add_argument("--row-kernel", type=int, default=0, help="ow_set_row_kernel mode (0 = per-N default, 1 = classic, 2 = persistent register-pipelined, 3 = persistent bulk-async staged)")
    ap.add_argument("--col-kernel", type=int, default=0, help="ow_set_column_kernel mode (0 = per-N default, 1 = ow_col_kernel, 2 = ow_col2_kernel, 3 = ow_col2_kernel TMA-staged)")
    ap.add_argument("--fused", type=int, default=-1, help="normal map as the column kernel's epilogue: -1 = per-N default, 0 = off, 1 = on")
    ap.add_argument("--discard", action="store_true", help="ow_set_discard_intermediate(1)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-compare", action="store_true", help="skip the cuFFT comparison leg")
    ap.add_argument("--no-slab-check", action="store_true", help="c5: skip the slab-vs-single-GPU agreement check before timing")
    ap.add_argument("--fused-normals", action="store_true", help="experimental OW_FLAG_FUSED_NORMALS (normal map as the column kernel's epilogue)")
    ap.add_argument("--col-lines", type=int, default=0, help="c5: ow_slab_set_column_lines (0 = default, 2 = 8-column tiles, 4 = persistent pipelined)")
    ap.add_argument("--post-ctas", type=int, default=-1, help="c5: ow_slab_set_post_ctas (CTAs per SM of the row pass's store kernel; -1 = the library's choice)")
    ap.add_argument("--no-pipeline", action="store_true", help="c5: one frame at a time (no overlap of the next frame's rows with this frame's columns)")
    ap.add_argument("--no-graph", action="store_true", help="c4: submit the step through ow_step_multi (stream launches) instead of ow_step (one graph launch)")
    ap.add_argument("--line-clusters", type=int, default=0, help="c5: ow_slab_set_line_clusters mode (0 = scratch path (default), -1 = clusters wherever possible, 1 / 3 / 7 = bit mask)")
    ap.add_argument("--c4-shard-of", type=int, default=1, help="c4 on one GPU only: run the 64/P cascades one rank of a P-GPU job would get")
    ap.add_argument("--c5-n", type=int, default=32768, help="c5 only: grid size (32768 = BASELINE config C5; 4096 = its down-scaled parity grid)")
    ap.add_argument("--transport", choices=["auto", "peer", "alltoall"], default="auto", help="c5 only: how the transpose crosses GPUs")
    ap.add_argument("--profile", action="store_true",
                    help="profiler mode (ncu): run warm-up + timed sweeps only and exit without printing a bench line")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours" and not args.profile:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
