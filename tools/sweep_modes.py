"""Whole-frame throughput of every (row kernel, column kernel, fused normals) combination, in the multi-frame sweep mode bench.py
times (development tool; one context per workload, modes switched through ow_set_row_kernel / ow_set_column_kernel).
usage: python tools/sweep_modes.py [c2 c3 c4 ...] [--streams 1,3] [--lib path]"""
import argparse, itertools, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workloads", nargs="*", default=["c2", "c3", "c4"])
    ap.add_argument("--streams", default="3")
    ap.add_argument("--groups", default="0")
    ap.add_argument("--rows", default="1,2,3")
    ap.add_argument("--cols", default="1:0,2:0,2:1,3:0,3:1")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--caps", default="0:0", help="resident CTAs per SM of the persistent row:column kernels, comma list (0 = what fits)")
    ap.add_argument("--c4-shard-of", type=int, default=1)
    ap.add_argument("--l2", default="0", help="ow_set_l2_persist modes, comma list")
    ap.add_argument("--discard", action="store_true", help="ow_set_discard_intermediate(1): the column kernel drops the intermediate's lines from L2 after reading them")
    ap.add_argument("--slots", type=int, default=0)
    ap.add_argument("--mega", type=int, default=0, help="ow_set_frame_kernel")
    ap.add_argument("--lat", type=int, default=1, help="ow_set_latency_shapes: 1 = default, 2 = wide row shape for every launch")
    ap.add_argument("--graph", action="store_true", help="c4: time ow_step (one CUDA graph launch per step) instead of ow_step_multi")
    ap.add_argument("--check", action="store_true", help="compare every combination's frame with the first combination's")
    args = ap.parse_args()
    import torch
    import bench
    import fft_ocean_waves_b200 as fow
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); sp = stream.cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    rows = [int(x) for x in args.rows.split(",")]
    cols = [tuple(int(y) for y in x.split(":")) for x in args.cols.split(",")]
    for name in args.workloads:
        w = bench.workload_setup(name, only=(0, 64 // args.c4_shard_of) if name == "c4" and args.c4_shard_of > 1 else None)
        N, frames = w["N"], w["frames"]
        slots = min(args.slots or (300 if N <= 512 else 32), frames) if name != "c4" else frames
        sim = fow.FFTOceanWaves(N=N, cascades=w["cascades"], n_slots=max(slots, len(w["cascades"])), device=0, jacobian=w["jacobian"])
        for i, nz in enumerate(w["noise"]):
            sim.set_noise(nz, cascade=i)
        sim.tilde_h0_k()
        if args.discard:
            sim.set_discard_intermediate(True)
        sim._check(sim._lib.ow_set_latency_shapes(sim._h, args.lat), 'ow_set_latency_shapes')
        if args.mega:
            sim.set_frame_kernel(args.mega)

        def sweep():
            if args.graph:
                sim.update(w["times"][0], stream=sp)
                return
            for base in range(0, frames, slots):
                n = min(slots, frames - base)
                sim.update_multi(w["cascade_of"][base:base + n], w["times"][base:base + n], stream=sp)
        ref = None
        for streams, group in itertools.product([int(x) for x in args.streams.split(",")], [int(x) for x in args.groups.split(",")]):
          for l2 in [int(x) for x in args.l2.split(",")]:
              sim.set_streams(streams); sim.set_group_size(group); sim.set_l2_persist(l2)
              for rm, (cm, fu), cap in itertools.product(rows, cols, [tuple(int(y) for y in x.split(":")) for x in args.caps.split(",")]):
                  sim.set_row_kernel(rm); sim.set_column_kernel(cm, fu); sim.set_resident_ctas(*cap)
                  try:
                      for _ in range(2):
                          sweep()
                      torch.cuda.synchronize()
                      ts = []
                      for _ in range(args.reps):
                          flush.zero_()
                          a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                          a.record(stream); sweep(); b.record(stream); torch.cuda.synchronize()
                          ts.append(a.elapsed_time(b))
                      ms = float(np.median(ts))
                      kms = np.zeros(3)
                      for base in range(0, frames, slots):
                          n = min(slots, frames - base)
                          kms += np.array(sim.update_multi_timed(w["cascade_of"][base:base + n], w["times"][base:base + n], stream=sp))
                      extra = ""
                      if args.check:
                          sim.update_multi([0], [w["times"][min(1, frames - 1)]], stream=sp); sim.sync(stream=sp)
                          got = {k: sim.download(k, 0) for k in ["dy", "dx", "dz", "normal"] + (["jacobian"] if w["jacobian"] else [])}
                          if ref is None:
                              ref = got
                          else:
                              errs = {k: float(np.abs(got[k] - ref[k]).max() / max(np.abs(ref[k]).max(), 1e-30)) for k in got}
                              extra = " maxrel " + " ".join("%s=%.1e" % kv for kv in errs.items())
                              if max(errs.values()) > 1e-5: extra += "  <-- MISMATCH"
                      print("%s N=%d streams=%d group=%d l2=%d row=%d col=%d fused=%d caps=%s : %8.0f fps  %7.2f us/frame   [row %.1f col %.1f nrm %.1f us/frame serial]%s" % (
                          name, N, streams, group, l2, rm, cm, fu, "%d:%d" % cap, frames / (ms * 1e-3), ms * 1e3 / frames, kms[0] * 1e3 / frames, kms[1] * 1e3 / frames, kms[2] * 1e3 / frames, extra), flush=True)
                  except Exception as e:
                      print("%s row=%d col=%d fused=%d FAILED: %s" % (name, rm, cm, fu, e), flush=True)
        sim.close()


if __name__ == "__main__":
    main()
