"""Single-frame (drop-in) path: ow_step one frame at a time, N from argv. With --ncu-friendly it just runs 6 frames without the graph
(for an ncu launch list); otherwise it times graph and plain launches back to back.  usage: python tools/single_frame_probe.py N [--plain]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import fft_ocean_waves_b200 as fow

N = int(sys.argv[1])
plain = "--plain" in sys.argv
torch.cuda.set_device(0)
st = torch.cuda.Stream(); sp = st.cuda_stream
p = fow.OceanParams(L=1000.0, wind_speed=40.0, wind_dir=(1.0, 1.0), amplitude=2.0, suppression=0.1, choppiness=1.0)
noise = np.random.default_rng(7).integers(0, 256, (4, N, N), dtype=np.uint8)
with fow.FFTOceanWaves(N=N, cascades=[p], device=0) as sim:
    sim.init(noise)
    if plain:
        sim.set_graph(False)
        for f in range(6):
            sim.update(f / 60.0, stream=sp)
        sim.sync(stream=sp)
        sys.exit(0)
    for graph, lat in ((True, True), (False, True), (True, False)):
        sim.set_graph(graph)
        sim.set_latency_shapes(lat)
        for f in range(20):
            sim.update(f / 60.0, stream=sp)
        sim.sync(stream=sp)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 300
        a.record(st)
        for f in range(n):
            sim.update(f / 60.0, stream=sp)
        b.record(st)
        torch.cuda.synchronize()
        print("N=%d %s %s: %.2f us/frame back to back (%.0f fps), %d launches" % (N, "graph" if graph else "launches", "latency shapes" if lat else "throughput shapes", a.elapsed_time(b) * 1e3 / n, n / (a.elapsed_time(b) * 1e-3), sim.last_launch_count()))
