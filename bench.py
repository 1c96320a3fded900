#!/usr/bin/env python
"""bench.py — frames/sec of the per-frame Tessendorf hot path (spectrum + 3x IFFT + inversion + normals).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4]

Default workload = BASELINE.json configs[1] ("C2"): N=512 single patch, L=1000 m, wind 40 m/s, A=2, the
reference's PNG noise, a 600-frame time sweep t_f = f/60. One STEP = one full 600-frame sweep; every frame writes
its complete dy/dx/dz/normal set to HBM. `value` = frames/s with inputs resident in HBM (CUDA events on the launching
stream); `e2e` = the same sweep through the C ABI with HOST buffers (noise upload + h0 init + every frame's outputs
copied back to pinned host memory inside the timed region).

--impl reference times the CPU oracle (oracle/ow_oracle.cpp, the scalar C++ restatement of the reference's GLSL
dispatch chain; the reference itself needs an OpenGL driver and cannot run here) on all host threads, on a
bounded sample of the same workload.

Under torchrun (--gpus N > 1) every rank runs the same sweep on its own GPU (independent patches, no data-path
collective; weak scaling); timing = max over ranks.
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
ALG_BYTES_PER_TEXEL = 44            # read h0k+h0minusk 16 B, write dy,dx,dz 12 B + normal 16 B
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


def workload_setup(name):
    import fft_ocean_waves_b200 as fow
    N, frames, desc, jac = WORKLOADS[name]
    if name == "c4":
        casc = []
        for c in range(64):
            ang = 2 * np.pi * c / 64
            casc.append(fow.OceanParams(L=float(100.0 * 1.08 ** c), wind_speed=float(10 + 0.5 * c),
                                        wind_dir=(float(np.cos(ang)), float(np.sin(ang))), amplitude=2.0, suppression=0.1, choppiness=1.0))
        noise = [np.random.default_rng(1024 + c).integers(0, 256, (4, N, N), dtype=np.uint8) for c in range(64)]
        cascade_of = list(range(64))
        times = [1.0] * 64
    else:
        casc = [fow.OceanParams(L=1000.0, wind_speed=40.0, wind_dir=(1.0, 1.0), amplitude=2.0, suppression=0.1, choppiness=1.0)]
        if name == "c5":
            noise = [_philox_noise(32768, N)]
        else:
            noise = [fow.default_noise() if name == "c2" else np.random.default_rng(2048).integers(0, 256, (4, N, N), dtype=np.uint8)]
        cascade_of = [0] * frames
        times = [float(np.float32(f / 60.0)) for f in range(frames)]
    return dict(N=N, frames=frames, desc=desc, jacobian=jac, cascades=casc, noise=noise, cascade_of=cascade_of, times=times)


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


def run_reference(args):
    """CPU arm: the oracle (scalar C++ restatement of the reference's GLSL chain) on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.oracle import OracleSim, max_threads
    from oracle import ref as refmod
    if args.workload == "c5":
        WORKLOADS["c5"] = (min(args.c5_n, 4096), 32, WORKLOADS["c5"][2].replace("N=32768", f"N={min(args.c5_n, 4096)} (CPU arm: down-scaled, the "
                           "oracle's reference textures need 116 GB at N=32768)"), True)
    w = workload_setup(args.workload)
    N, cores = w["N"], max_threads()
    p = w["cascades"][0]
    # oracle/_ref = the reference's own compute shaders compiled for the CPU (oracle/make_ref.py): the reference arm proper. It needs
    # the reference's integer patch size and has no Jacobian (neither has the reference); otherwise the oracle port stands in.
    use_ref = refmod.available() and float(p.L) == int(p.L)
    if use_ref:
        rs = refmod.RefSim(N, int(p.L), p.wind_speed, p.wind_dir, p.amplitude, p.suppression, w["noise"][0], threads=cores)

        class _Sim:
            def frame(self, t, choppiness=None):
                return rs.frame(t)
        sim, kind = _Sim(), "reference"
        what = ("oracle/_ref: the reference's own GLSL compute shaders (tilde_h0_t, butterfly x 2 log2 N x 3, inversion x 3, normal_map) compiled "
                "as C++ and dispatched in the reference's order on the host cores (OpenMP over rows); no Jacobian, as in the reference")
    else:
        sim, kind = OracleSim(N, p.L, p.wind_speed, p.wind_dir, p.amplitude, p.suppression, w["noise"][0], threads=cores), "port"
        what = "CPU oracle = scalar C++ restatement of the reference's GLSL dispatch chain (oracle/_ref not available here)"
    lam = 1.0 if (w["jacobian"] and not use_ref) else None
    sim.frame(w["times"][0], choppiness=lam)                  # first call pays thread start-up and page faults
    t0 = time.perf_counter()
    sim.frame(w["times"][1 % len(w["times"])], choppiness=lam)
    one = time.perf_counter() - t0
    budget = 120.0 / max(1, args.steps + args.warmup)          # whole run within a few minutes
    sample = int(max(1, min(w["frames"], min(budget, 8.0) / max(one, 1e-6))))
    def step():
        for f in range(sample):
            sim.frame(w["times"][f % len(w["times"])], choppiness=lam)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    fps = sample * args.steps / dt
    desc = f"first {sample} frames of the {w['frames']}-frame sweep per step"
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "N": N, "frames_per_step": sample, "sample": desc,
                       "what": what},
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if use_ref:
        # for transparency: the hand-written restatement (oracle/ow_oracle.cpp) is faster than the reference's shader text compiled through
        # the GLSL emulation layer; its throughput on the same frames, same cores, is reported beside the reference arm's
        port = OracleSim(N, p.L, p.wind_speed, p.wind_dir, p.amplitude, p.suppression, w["noise"][0], threads=cores)
        port.frame(w["times"][0])
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < 2.0 or n < 3:
            port.frame(w["times"][n % len(w["times"])])
            n += 1
        line["cpu_port"] = {"value": n / (time.perf_counter() - t0), "unit": UNIT, "cores": cores, "kind": "port",
                            "what": "oracle/ow_oracle.cpp on the same workload (not the number the driver compares against)"}
    print(json.dumps(line), flush=True)


def cpu_baseline(w, seconds=10.0):
    from oracle.oracle import OracleSim, max_threads
    cores = max_threads()
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
            "sample": f"first {n} frames of the sweep ({dt:.1f} s), oracle/ow_oracle.cpp with OpenMP over rows"}


def run_slab(args):
    """--workload c5: one grid over all ranks (strong scaling). A step = one 32-frame sweep of SlabOcean.update."""
    import torch
    import fft_ocean_waves_b200 as fow

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, frames, desc, jac = WORKLOADS["c5"]
    if args.c5_n != N:
        N, frames = args.c5_n, (32 if args.c5_n <= 4096 else 4)
        desc = desc.replace("N=32768", f"N={N} (down-scaled)").replace("4-frame", f"{frames}-frame")
    p = fow.OceanParams(L=1000.0, wind_speed=40.0, wind_dir=(1.0, 1.0), amplitude=2.0, suppression=0.1, choppiness=1.0)
    times = [float(np.float32(f / 60.0)) for f in range(frames)]
    sim = fow.SlabOcean(N=N, params=p, device=local, jacobian=jac, transport=args.transport)
    sim.init(32768)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def sweep():
        for t in times:
            sim.update(t)

    for _ in range(args.warmup):
        flush.zero_()
        sweep()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    evs = []
    for _ in range(args.steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        sweep()
        b.record(stream)
        evs.append((a, b))
    barrier()
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    if dist is not None:
        tt = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    value = frames * args.steps / (total_ms * 1e-3)
    if args.profile:
        sim.close()
        return
    # per-phase durations: events around ow_slab_rows (1 kernel) and ow_slab_cols (column + normal kernels)
    b_ = sim.backend
    st = b_.current_stream()
    kms = np.zeros(2)
    for t in times:
        if sim.transport == "peer":
            sim._barrier()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
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
        kms += np.array([e[0].elapsed_time(e[1]), e[2].elapsed_time(e[3])])
    if dist is not None:
        tt = torch.tensor(kms, device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        kms = tt.cpu().numpy()
    clocks = sampler.stop() if rank == 0 else None
    peak, peak_src = measured_peak()
    texels_rank = float(N) * N / world
    # compulsory bytes per texel of the two phases (folded spectrum 8 -> intermediate 12; intermediate 12 -> dy,dx,dz 12, then
    # dy,dx,dz 12 -> normal 16 + J 4); above N=4096 the line decomposition's scratch adds 24 B/texel per direction, not counted
    kb = {"ow_row_slab_kernel": 8 + 12, "ow_col_slab_kernel+ow_normal_slab_kernel": 12 + 12 + 4 + 8 + 16 + 4}
    per_kernel = []
    for i, k in enumerate(kb):
        gbs = kb[k] * texels_rank * frames / (kms[i] * 1e-3) / 1e9
        per_kernel.append({"kernel": k, "ms_per_launch": kms[i] / frames, "share": kms[i] / kms.sum(), "bytes_per_texel": kb[k],
                           "achieved_gbs": gbs, "frac": gbs / peak})
    dom = int(np.argmax(kms))
    frame_gbs = 48 * texels_rank * value / 1e9
    xbytes = sim.exchange_bytes_per_frame()
    roofline = {"bound": "hbm", "kernel": per_kernel[dom]["kernel"], "achieved": per_kernel[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": per_kernel[dom]["frac"], "traffic": None, "peak_source": peak_src,
                "bytes_per_launch": per_kernel[dom]["bytes_per_texel"] * texels_rank, "kernels": per_kernel,
                "frame": {"algorithmic_bytes_per_texel": 48, "achieved": frame_gbs, "frac": frame_gbs / peak,
                          "note": "per GPU: 48 B/texel x N^2/world texels x frames/s"},
                "nvlink": {"bytes_per_frame_per_gpu_per_direction": xbytes, "achieved_gbs": xbytes * value / 1e9, "peak_gbs": 770.0,
                           "frac": xbytes * value / 1e9 / 770.0,
                           "note": "12 B/texel Hermitian-packed intermediate x (world-1)/world of this rank's texels (+ halo columns); "
                                   "peak = measured peer copy 770 GB/s per direction (B200_PROFILING.md)"}}
    # ---- end to end: seed in, every frame's column slab copied to pinned host memory ---------------------------------
    outs = b_.output_tensors()
    host = {k: torch.empty(v.shape, dtype=torch.float32, pin_memory=True) for k, v in outs.items()}
    fbytes = sum(h.numel() * 4 for h in host.values())

    def e2e_step():
        sim.init(32768)
        for t in times:
            sim.update(t)
            for k, v in outs.items():
                host[k].copy_(v, non_blocking=True)
        torch.cuda.synchronize()
        return float(host["dy"][0, 0])

    e2e_step()
    barrier()
    e2e_steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_t = time.perf_counter() - t0
    if dist is not None:
        tt = torch.tensor([e2e_t], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_t = float(tt.item())
    e2e = {"value": frames * e2e_steps / e2e_t, "unit": UNIT, "h2d_bytes_per_step": 8, "d2h_bytes_per_step": int(frames * fbytes * world),
           "steps": e2e_steps, "what": "ow_slab_init_spectrum_seeded (8-byte seed) + SlabOcean.update + every rank's column slab "
                                       "(dy,dx,dz,normal,J) copied to pinned host memory every frame"}
    line = None
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": desc, "N": N, "frames_per_step": frames, "transport": sim.transport,
                           "l2": "flushed between timed steps (256 MiB memset outside the event pair); a frame's working set "
                                 f"({(16 + 12 + 12 + 20) * N * N / world / 1e6:.0f} MB per GPU) exceeds L2 at world <= 4",
                           "parallelism": f"slab{world}: row pairs -> transpose (peer stores / all-to-all over NVLink) -> column slabs"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(3 * frames * args.steps), "roofline": roofline}
        if world == 1 and not args.no_cpu and N <= 4096:      # the CPU oracle's reference textures need 108 B/texel: 116 GB at N=32768
            w = dict(N=N, frames=frames, jacobian=True, cascades=[p], noise=[_philox_noise(32768, N)], times=times)
            line["cpu_baseline"] = cpu_baseline(w)
    sim.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def _philox_noise(seed, N):
    from oracle.numpy_ref import philox_noise
    return philox_noise(seed, N)


def run_ours(args):
    import torch
    import fft_ocean_waves_b200 as fow

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    w = workload_setup(args.workload)
    N, frames = w["N"], w["frames"]
    # C4 (BASELINE.json configs[3]): the 64 cascades are SHARDED, a contiguous block of 64/world per GPU,
    # no data-path collective -> strong scaling (total work fixed). Every other workload replicates per GPU (weak).
    sharded = args.workload == "c4" and (world > 1 or args.c4_shard_of > 1)
    if sharded:
        parts = world if world > 1 else args.c4_shard_of       # --c4-shard-of P: time ONE GPU's share of a P-GPU run (tuning aid)
        if 64 % parts:
            raise SystemExit("bench.py: c4 shards 64 cascades; --gpus must divide 64")
        per = 64 // parts
        lo = rank * per
        w["cascades"], w["noise"] = w["cascades"][lo:lo + per], w["noise"][lo:lo + per]
        w["cascade_of"], w["times"] = list(range(per)), w["times"][lo:lo + per]
        frames = w["frames"] = per
    job_frames = (64 if world > 1 else frames) if sharded else world * frames            # frames the WHOLE job produces per step
    slots = min(args.slots or (128 if N <= 512 else 32), frames) if args.workload != "c4" else frames
    sim = fow.FFTOceanWaves(N=N, cascades=w["cascades"], n_slots=max(slots, len(w["cascades"])), device=local, jacobian=w["jacobian"],
                            fused_normals=args.fused_normals)
    for i, nz in enumerate(w["noise"]):
        sim.set_noise(nz, cascade=i)
    sim.tilde_h0_k()
    if args.group:
        sim.set_group_size(args.group)
    if args.streams:
        sim.set_streams(args.streams)
    # A dedicated non-default stream: its handle is what the C ABI launches on, and the torch events below are
    # recorded on the same stream (handle 0 would mean "the context's own stream" to ow_step*).
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    assert sp != 0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    launches, groups = [0], [0]

    def sweep():
        for base in range(0, frames, slots):
            n = min(slots, frames - base)
            sim.update_multi(w["cascade_of"][base:base + n], w["times"][base:base + n], stream=sp)
            launches[0] += sim.last_launch_count()
            groups[0] += sim.last_group_count()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        flush.zero_()
        sweep()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches[0] = groups[0] = 0
    evs = []
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()                      # L2 flush between timed iterations (outside the event pair)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        sweep()
        b.record(stream)
        evs.append((a, b))
    barrier()
    t_wall = time.perf_counter() - t_wall
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    timed_launches = launches[0]
    if dist is not None:
        tt = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    value = job_frames * args.steps / (total_ms * 1e-3)

    if args.profile:
        sim.close()
        return
    # ---- per-kernel durations (CUDA events around each kernel, same stream, same workload) -> roofline ----
    kms = np.zeros(3)
    prof_sweeps = 2
    for _ in range(prof_sweeps):
        flush.zero_()
        for base in range(0, frames, slots):
            n = min(slots, frames - base)
            kms += np.array(sim.update_multi_timed(w["cascade_of"][base:base + n], w["times"][base:base + n], stream=sp))
    clocks = sampler.stop() if rank == 0 else None
    kms /= prof_sweeps                                      # ms per sweep per kernel
    groups_per_sweep = groups[0] / args.steps
    fused = not w["jacobian"] and kms[2] == 0.0          # normal map produced by the column kernel's epilogue (no separate kernel)
    peak, peak_src = measured_peak()
    texels = float(N) * N
    kb = dict(KERNEL_BYTES_PER_TEXEL)
    if w["jacobian"]:
        kb["ow_normal_kernel"] += 8 + 4                     # + read dx,dz, write J
    if fused:
        kb["ow_col_kernel"] += 16                           # ow_col_fused_kernel also writes the normal map (and never re-reads dy)
    per_kernel = []
    for i, k in enumerate(KERNELS):
        if kms[i] == 0.0:
            continue
        if fused and k == "ow_col_kernel":
            k = "ow_col_fused_kernel"
        gbs = kb.get(k, kb["ow_col_kernel"]) * texels * frames / (kms[i] * 1e-3) / 1e9
        per_kernel.append({"kernel": k, "ms_per_launch": kms[i] / groups_per_sweep, "share": kms[i] / kms.sum(),
                           "bytes_per_texel": kb.get(k, kb["ow_col_kernel"]), "achieved_gbs": gbs, "frac": gbs / peak})
    dom = int(np.argmax([pk["share"] for pk in per_kernel]))
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"{args.workload}:{per_kernel[dom]['kernel']}")
    except Exception:
        pass
    alg = ALG_BYTES_PER_TEXEL + (4 if w["jacobian"] else 0)
    frame_gbs = alg * texels * value / world / 1e9
    roofline = {"bound": "hbm", "kernel": per_kernel[dom]["kernel"], "achieved": per_kernel[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": per_kernel[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
                "bytes_per_launch": per_kernel[dom]["bytes_per_texel"] * texels * frames / groups_per_sweep,
                "kernels": per_kernel,
                "frame": {"algorithmic_bytes_per_texel": alg, "achieved": frame_gbs, "frac": frame_gbs / peak,
                          "note": "whole frame on SURVEY.md's 44 B/texel (48 with Jacobian), per GPU; the 3-kernel design moves 64 B/texel (8+12, 12+12, 4+16) of which 36 are compulsory since the h0 fold"}}

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
    barrier()
    e2e_steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_t = time.perf_counter() - t0
    if dist is not None:
        tt = torch.tensor([e2e_t], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_t = float(tt.item())
    e2e = {"value": job_frames * e2e_steps / e2e_t, "unit": UNIT,
           "h2d_bytes_per_step": int(sum(nz.numel() for nz in pinned_noise)), "d2h_bytes_per_step": int(frames * fbytes),
           "steps": e2e_steps,
           "what": "ow_set_noise + ow_init_spectrum + ow_step_multi + ow_download_frame_async of every frame into pinned host memory"}

    # ---- latency-style number: one frame per call, single output slot (what an interactive renderer does) ----
    seq_fps = None
    if args.workload != "c4":
        nseq = min(frames, 200)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for f in range(nseq):
            sim.update_multi([0], [w["times"][f]], stream=sp)
        b.record(stream)
        torch.cuda.synchronize()
        seq_fps = nseq / (a.elapsed_time(b) * 1e-3)

    line = None
    if rank == 0:
        cpu = cpu_baseline(w) if world == 1 and not args.no_cpu else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": w["desc"], "N": N, "frames_per_step": job_frames if sharded else frames, "slots_per_launch": slots,
                           "launch_groups_per_step": groups_per_sweep,
                           "l2": "flushed between timed steps (256 MiB memset outside the event pair); within a step the outputs "
                                 f"({frames * fbytes / 1e9:.2f} GB) stream through L2, h0 ({16 * texels * len(w['cascades']) / 1e6:.1f} MB) is re-read every frame as in the reference",
                           "parallelism": (f"64 cascades sharded {frames} per GPU over {world} GPUs, no communication" if sharded
                                           else f"{world} x independent patch per GPU, no communication"),
                           "single_slot_sequential_fps": seq_fps, "wall_s_timed_region": t_wall},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(timed_launches), "roofline": roofline}
        if cpu is not None:
            line["cpu_baseline"] = cpu
    sim.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="c2")
    ap.add_argument("--slots", type=int, default=0, help="frames evaluated per ow_step_multi call (0 = 128 for c2, 32 for c3, 64 for c4)")
    ap.add_argument("--group", type=int, default=0, help="slots per launch group (0 = library default)")
    ap.add_argument("--streams", type=int, default=0, help="internal streams the launch groups are spread over (0 = library default)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--fused-normals", action="store_true", help="experimental OW_FLAG_FUSED_NORMALS (normal map as the column kernel's epilogue)")
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
    elif args.workload == "c5":
        run_slab(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
