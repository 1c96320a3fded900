"""torchrun worker for the multi-GPU slab tests: every rank runs SlabOcean over NCCL; rank 0 also runs the single-GPU
path on the same Philox seed and compares the gathered column slabs with it (same phase functions; agreement to fp32
round-off, 1e-6 of peak — the single-GPU row kernel is the pipelined variant, so FMA contraction may differ).
    torchrun --nproc-per-node P tests/slab_worker.py N [shared]
"shared": every rank uses GPU 0 and the process group is gloo — real ow_slab contexts in separate processes, peer stores through CUDA
IPC mappings of each other's receive buffers, host-side barriers. That is the multi-rank slab path on a box with ONE GPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import fft_ocean_waves_b200 as fow

    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    shared = len(sys.argv) > 2 and sys.argv[2] == "shared"
    local = 0 if shared else int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if shared:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    p = fow.OceanParams(L=1000.0, wind_speed=40.0, wind_dir=(1.0, 1.0), amplitude=2.0, suppression=0.1, choppiness=1.0)
    seed, times = 32768, (0.25, 0.5, 0.75, 1.0)       # four frames: both receive buffers of the pipelined path are reused once
    ref = None
    if rank == 0:
        with fow.FFTOceanWaves(N=N, cascades=[p], jacobian=True, device=local) as one:
            one.set_noise_seed(seed)
            one.tilde_h0_k()
            ref = one.frame(times[-1])
    ok = True
    # (transport, pipeline): over NCCL the peer transport runs frame-pipelined by default (rows of frame f+1 during the columns of frame f)
    variants = (("peer", None),) if shared else (("peer", None), ("peer", False), ("alltoall", None), ("alltoall", False))
    for transport, pipeline in variants:
        with fow.SlabOcean(N=N, params=p, device=local, jacobian=True, transport=transport, pipeline=pipeline) as sim:
            sim.init(seed)
            assert sim.transport == transport
            assert sim.pipelined == (pipeline is None and not shared and world > 1), (transport, pipeline, sim.pipelined)
            transport = transport + ("+pipelined" if sim.pipelined else "")
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                for t in times:                      # frames back to back: exercises the buffer-reuse ordering
                    sim.update(t)
                sim.sync()
                full = {k: sim.gather(k) for k in ("dy", "dx", "dz", "normal", "jacobian")}
            if rank == 0:
                for k, v in full.items():
                    err = float(np.abs(v - ref[k]).max())
                    tol = 1e-6 * float(np.abs(ref[k]).max()) + (2e-6 if k in ("normal", "jacobian") else 0.0)
                    same = err <= tol
                    print(f"[slab N={N} world={world} {transport}] {k}: max err {err:.3e} (tol {tol:.3e}) {'ok' if same else 'MISMATCH'}", flush=True)
                    ok &= same
    flag = torch.tensor([1.0 if ok else 0.0], device="cpu" if shared else "cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("SLAB OK" if ok else "SLAB FAILED", flush=True)
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
