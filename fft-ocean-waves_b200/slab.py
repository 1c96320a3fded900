"""One ocean grid over several GPUs: host side of the slab-decomposed frame (BASELINE config C5).

One process per GPU (torchrun); every process owns one `ow_slab` (include/oceanwaves.h) for its rank. The reference
functions replaced are the same as for the single-GPU path — tilde_h0_t(), butterfly_fft() x3, generate_normal_map()
(reference src/main.cpp:240-244) — split at the row/column boundary of the 2-D IFFT (src/main.cpp:626-661):

    rows    ow_slab_rows   spectrum + row IFFT of this rank's row pairs; results are stored in TRANSPOSED blocks
    exchange               transport "peer":     the row kernel stored straight into the column owners' buffers over
                                                 NVLink (CUDA IPC peer mappings); the host only orders rows before columns
                                                 (a stream-ordered all-reduce of one element as the barrier)
                           transport "alltoall": one equal-split all_to_all_single send -> recv (NCCL over NVLink)
    cols    ow_slab_cols   column IFFT + inversion + normals (+ Jacobian) on this rank's column slab

PyTorch is plumbing only: process group, stream handle, tensor views of the library's device buffers. The compute is the
C ABI; `backend=` exists so the host logic (partition, transports, ordering) can be driven by a stand-in in the CPU tests
(tests/test_slab_host.py, gloo, world_size 2) — there is no CPU fallback in the product: without the CUDA library
`SlabOcean()` raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from .sim import IMAGES, OW_FLAG_EXACT_SINCOS, OW_FLAG_JACOBIAN, OceanParams, OceanWavesError, _SlabInfo, load_library

HALO = 8                      # == kSlabHalo in csrc/ow_internal.h
SEND_BUFFER, PEER_STORES = 0, 1
IPC_HANDLE_BYTES = 64


def slab_plan(N: int, world: int, rank: int) -> dict:
    """Who owns what (the same arithmetic as ow_slab_create): row pairs, the h0 rows they need, columns, block sizes."""
    if world < 1 or (N // 2) % world or (N // world) % 128:
        raise ValueError(f"unsupported slab decomposition N={N} world={world}")
    PL, XL = N // 2 // world, N // world
    XH = XL + 2 * HALO
    pairs = range(rank * PL, (rank + 1) * PL)
    rows = [0 if p == 0 else p for p in pairs] + [N // 2 if p == 0 else N - p for p in pairs]   # local h0 row order
    return dict(N=N, world=world, rank=rank, pairs_per_rank=PL, cols_per_rank=XL, padded_cols=XH, halo=HALO,
                first_pair=rank * PL, first_col=rank * XL, h0_rows=rows, block_elems=PL * 3 * XH,
                block_bytes=PL * 3 * XH * 8)


class _DeviceBytes:
    """Just enough of the CUDA array interface for torch.as_tensor to alias library-owned device memory."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class LibSlabBackend:
    """The product backend: the ow_slab_* C ABI of liboceanwaves.so."""

    def __init__(self, N: int, world: int, rank: int, params: OceanParams, device: int, jacobian: bool, exact_sincos: bool):
        self._lib = load_library()
        self._h = C.c_void_p()
        self.device = int(device)
        self.jacobian = bool(jacobian)
        p = params.to_c()
        flags = (OW_FLAG_JACOBIAN if jacobian else 0) | (OW_FLAG_EXACT_SINCOS if exact_sincos else 0)
        rc = self._lib.ow_slab_create(int(N), int(world), int(rank), C.byref(p), self.device, flags, C.byref(self._h))
        if rc != 0:
            msg = self._lib.ow_slab_last_error(None)
            self._h = C.c_void_p()
            raise OceanWavesError(f"ow_slab_create failed ({rc}): {msg.decode() if msg else ''}")
        info = _SlabInfo()
        self._check(self._lib.ow_slab_get_info(self._h, C.byref(info)), "ow_slab_get_info")
        self.info = info
        self._views = None

    def _check(self, rc: int, what: str):
        if rc != 0:
            msg = self._lib.ow_slab_last_error(self._h)
            raise OceanWavesError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._views = None
            self._lib.ow_slab_destroy(self._h)
            self._h = C.c_void_p()

    # geometry
    @property
    def plan(self) -> dict:
        i = self.info
        return dict(N=i.N, world=i.world, rank=i.rank, pairs_per_rank=i.pairs_per_rank, cols_per_rank=i.cols_per_rank,
                    padded_cols=i.padded_cols, halo=i.halo, block_bytes=int(i.block_bytes), block_elems=int(i.block_bytes) // 8)

    # compute
    def init_spectrum(self, seed: int):
        self._check(self._lib.ow_slab_init_spectrum_seeded(self._h, C.c_uint64(int(seed))), "ow_slab_init_spectrum_seeded")

    def rows(self, t: float, transport: int, stream: int = 0):
        self._check(self._lib.ow_slab_rows(self._h, float(t), int(transport), C.c_void_p(stream or None)), "ow_slab_rows")

    def cols(self, stream: int = 0):
        self._check(self._lib.ow_slab_cols(self._h, C.c_void_p(stream or None)), "ow_slab_cols")

    def local_exchange(self, stream: int = 0):
        self._check(self._lib.ow_slab_local_exchange(self._h, C.c_void_p(stream or None)), "ow_slab_local_exchange")

    def sync(self, stream: int = 0):
        self._check(self._lib.ow_slab_sync(self._h, C.c_void_p(stream or None)), "ow_slab_sync")

    def set_line_clusters(self, mode: int = -1):
        self._check(self._lib.ow_slab_set_line_clusters(self._h, int(mode)), "ow_slab_set_line_clusters")

    def line_clusters(self) -> int:
        return int(self._lib.ow_slab_get_line_clusters(self._h))

    # transports
    def ipc_handle(self) -> bytes:
        buf = (C.c_ubyte * IPC_HANDLE_BYTES)()
        self._check(self._lib.ow_slab_ipc_handle(self._h, buf, IPC_HANDLE_BYTES), "ow_slab_ipc_handle")
        return bytes(buf)

    def open_peers(self, handles):
        blob = b"".join(handles)
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        self._check(self._lib.ow_slab_open_peers(self._h, buf, len(blob)), "ow_slab_open_peers")

    # two receive buffers: frame f+1's row pass (the exchange) overlaps frame f's column pass
    def enable_double_buffer(self):
        self._check(self._lib.ow_slab_enable_double_buffer(self._h), "ow_slab_enable_double_buffer")

    def ipc_handle_buf(self, buf: int) -> bytes:
        out = (C.c_ubyte * IPC_HANDLE_BYTES)()
        self._check(self._lib.ow_slab_ipc_handle_buf(self._h, int(buf), out, IPC_HANDLE_BYTES), "ow_slab_ipc_handle_buf")
        return bytes(out)

    def open_peers_buf(self, buf: int, handles):
        blob = b"".join(handles)
        arr = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        self._check(self._lib.ow_slab_open_peers_buf(self._h, int(buf), arr, len(blob)), "ow_slab_open_peers_buf")

    def rows_buf(self, t: float, transport: int, buf: int, stream: int = 0):
        self._check(self._lib.ow_slab_rows_buf(self._h, float(t), int(transport), int(buf), C.c_void_p(stream or None)), "ow_slab_rows_buf")

    def cols_buf(self, buf: int, stream: int = 0):
        self._check(self._lib.ow_slab_cols_buf(self._h, int(buf), C.c_void_p(stream or None)), "ow_slab_cols_buf")

    def set_column_lines(self, mode: int):
        self._check(self._lib.ow_slab_set_column_lines(self._h, int(mode)), "ow_slab_set_column_lines")

    def set_post_ctas(self, per_sm: int):
        self._check(self._lib.ow_slab_set_post_ctas(self._h, int(per_sm)), "ow_slab_set_post_ctas")

    def exchange_tensors(self, buf: int = 0):
        """(send, recv) as flat float32 torch tensors aliasing the library's device buffers (recv: receive buffer `buf`)."""
        if self._views is None:
            self._views = {}
        if buf not in self._views:
            import torch
            n = int(self.info.block_bytes) * self.info.world
            dev = f"cuda:{self.device}"
            send = torch.as_tensor(_DeviceBytes(self.info.send, n), device=dev).view(torch.float32)
            ptr = C.c_void_p()
            self._check(self._lib.ow_slab_recv_buffer(self._h, int(buf), C.byref(ptr)), "ow_slab_recv_buffer")
            recv = torch.as_tensor(_DeviceBytes(ptr.value, n), device=dev).view(torch.float32)
            self._views[buf] = (send, recv)
        return self._views[buf]

    def output_tensors(self) -> dict:
        """torch views of this rank's output slabs (halo columns cut off): dy/dx/dz [N][XL] (row stride XH), normal
        [N][XL][4], jacobian [N][XL]. Valid until close(); for stream-ordered D2H copies without extra staging."""
        import torch
        i, dev = self.info, f"cuda:{self.device}"
        out = {}
        for k, ptr in (("dy", i.dy), ("dx", i.dx), ("dz", i.dz)):
            full = torch.as_tensor(_DeviceBytes(ptr, i.N * i.padded_cols * 4), device=dev).view(torch.float32).view(i.N, i.padded_cols)
            out[k] = full[:, i.halo:i.halo + i.cols_per_rank]
        out["normal"] = torch.as_tensor(_DeviceBytes(i.normal, i.N * i.cols_per_rank * 16), device=dev).view(torch.float32).view(i.N, i.cols_per_rank, 4)
        if i.jacobian:
            out["jacobian"] = torch.as_tensor(_DeviceBytes(i.jacobian, i.N * i.cols_per_rank * 4), device=dev).view(torch.float32).view(i.N, i.cols_per_rank)
        return out

    def barrier_token(self):
        import torch
        return torch.zeros(1, device=f"cuda:{self.device}")

    def current_stream(self) -> int:
        """Handle of torch's current stream: the collectives are ordered on it, so the kernels must be too. torch's
        default stream is the legacy default stream (handle 0), which the C ABI would read as "the context's own
        stream": pass cudaStreamLegacy (0x1) instead."""
        import torch
        h = int(torch.cuda.current_stream(self.device).cuda_stream)
        return h if h != 0 else 1

    # outputs
    def download(self, name: str, stream: int = 0) -> np.ndarray:
        i = self.info
        shape = (i.N, i.cols_per_rank, 4) if name == "normal" else (i.N, i.cols_per_rank)
        out = np.empty(shape, np.float32)
        self._check(self._lib.ow_slab_download(self._h, IMAGES[name], out.ctypes.data, out.nbytes, C.c_void_p(stream or None)),
                    "ow_slab_download")
        return out


class SlabOcean:
    """FFTOceanWaves for ONE grid spread over the ranks of a torch.distributed process group (or a single process).

    sim = SlabOcean(N=4096, params=OceanParams(...)); sim.init(seed=32768); sim.update(t); dy_cols = sim.download("dy")
    """

    def __init__(self, N: int, params: Optional[OceanParams] = None, device: Optional[int] = None, jacobian: bool = False,
                 exact_sincos: bool = False, transport: str = "auto", group=None, backend=None, pipeline: Optional[bool] = None):
        """pipeline: overlap frame f+1's row pass and exchange (peer stores, or the all-to-all) with frame f's column pass, through a second
        receive buffer and two internal streams. None = whenever it applies (several ranks over NCCL); outputs are then complete after
        flush() / sync() / download()."""
        self._dist = None
        self.world, self.rank = 1, 0
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self._dist, self.group = dist, group
                self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        except ImportError:
            pass
        self.N = int(N)
        self.params = params or OceanParams()
        if backend is None:
            if device is None:
                import torch
                device = torch.cuda.current_device() if torch.cuda.is_available() else 0
            backend = LibSlabBackend(self.N, self.world, self.rank, self.params, device, jacobian, exact_sincos)
        self.backend = backend
        self.jacobian = bool(jacobian)
        self.plan = backend.plan
        if transport not in ("auto", "peer", "alltoall"):
            raise ValueError("transport must be 'auto', 'peer' or 'alltoall'")
        self.transport = transport
        self._peers_ready = False
        self._token = None
        self.frames = 0
        self._want_pipeline = pipeline
        self.pipelined = False
        self._pipe = None           # (rows stream, columns stream, rows-done events, columns-done events)
        self._fence_in = True       # the internal streams must first wait for the caller's stream

    def close(self):
        if self.pipelined and self._pipe is not None:
            self._pipe[0].synchronize()          # nothing of ours may still be running on the internal streams when the buffers go
            self._pipe[1].synchronize()
            self._pipe = None
            self.pipelined = False
        self.backend.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- init (reference init(): tilde_h0_k, src/main.cpp:218) ----------------------------------------------------
    def init(self, seed: int = 32768):
        self.backend.init_spectrum(seed)
        if self.transport in ("auto", "peer") and self.world > 1 and not self._peers_ready:
            # Every rank must end up on the SAME transport: a rank whose peer mapping failed cannot fall back on its own
            # while the others keep storing into peer buffers (mixed barriers and all-to-alls would hang or corrupt).
            err = None
            try:
                self._open_peers()
            except (OceanWavesError, NotImplementedError) as e:
                err = f"rank {self.rank}: {e}"
            errs = [None] * self.world
            self._dist.all_gather_object(errs, err, group=self.group)
            failed = [e for e in errs if e]
            if failed:
                if self.transport == "peer":
                    raise OceanWavesError("transport='peer' requested but peer mappings failed: " + "; ".join(failed))
                self.transport = "alltoall"          # decided identically on every rank
            else:
                self.transport = "peer"
        elif self.transport == "auto":
            self.transport = "peer"          # world == 1: the "peer" is this rank's own receive buffer
        self._setup_pipeline()
        return True

    def _setup_pipeline(self):
        """Second receive buffer + its peer mappings + two streams; decided identically on every rank."""
        can = (self.world > 1 and self._want_pipeline is not False and hasattr(self.backend, "rows_buf")
               and self._dist is not None and self._dist.get_backend(self.group) == "nccl")
        if not can:
            if self._want_pipeline:
                raise OceanWavesError("pipeline=True needs NCCL and more than one rank")
            return
        err = None
        try:
            self.backend.enable_double_buffer()
            if self.transport == "peer":
                handles = [None] * self.world
                self._dist.all_gather_object(handles, self.backend.ipc_handle_buf(1), group=self.group)
                self.backend.open_peers_buf(1, handles)
        except OceanWavesError as e:
            err = f"rank {self.rank}: {e}"
        errs = [None] * self.world
        self._dist.all_gather_object(errs, err, group=self.group)
        if any(errs):
            if self._want_pipeline:
                raise OceanWavesError("pipeline=True: " + "; ".join(e for e in errs if e))
            return
        import torch
        dev = self.backend.device
        # the ROWS stream has the higher priority: once both passes are eligible, frame f+1's rows must go first, so that their exchange (NVLink-
        # bound, few SMs) is under way while frame f's columns have the SMs; the other way round the exchange would queue up behind the columns
        self._pipe = (torch.cuda.Stream(dev, priority=-1), torch.cuda.Stream(dev), [torch.cuda.Event(), torch.cuda.Event()], [torch.cuda.Event(), torch.cuda.Event()])
        self.pipelined = True

    def _open_peers(self):
        handles = [None] * self.world
        self._dist.all_gather_object(handles, self.backend.ipc_handle(), group=self.group)
        self.backend.open_peers(handles)
        self._peers_ready = True

    def _barrier(self):
        """Stream-ordered barrier: an all-reduce of one element on the stream the kernels run on (NCCL). Under a host-side
        backend (gloo: the CPU tests, and several ranks sharing one GPU) the stream is drained first and the ranks meet on the host."""
        if self.world == 1:
            return
        if self._dist.get_backend(self.group) != "nccl":
            self.backend.sync(self.backend.current_stream())
            self._dist.barrier(group=self.group)
            return
        if self._token is None:
            self._token = self.backend.barrier_token()
        self._dist.all_reduce(self._token, group=self.group)

    # ---- per frame (reference update(): src/main.cpp:240-244) -------------------------------------------------------
    def update(self, t: float):
        if self.pipelined:
            return self._update_pipelined(t)
        b = self.backend
        st = b.current_stream()
        if self.transport == "peer":
            if self.frames:
                self._barrier()                  # every rank finished reading its receive buffer (previous columns)
            b.rows(t, PEER_STORES, st)
            self._barrier()                      # every rank's row results have landed
        else:
            b.rows(t, SEND_BUFFER, st)
            if self.world == 1:
                b.local_exchange(st)
            else:
                send, recv = b.exchange_tensors()
                self._dist.all_to_all_single(recv, send, group=self.group)
        b.cols(st)
        self.frames += 1

    def _update_pipelined(self, t: float):
        """rows(f) -> receive buffer f % 2 on the rows stream; cols(f) on the columns stream once every rank's rows(f) have landed;
        rows(f) only after every rank's cols(f-2) has finished reading that buffer. All barriers live on the rows stream, in the same order
        on every rank, so the columns of frame f run while the rows (and the NVLink stores) of frame f+1 are under way."""
        import torch
        b = self.backend
        R, Cs, ev_rows, ev_cols = self._pipe
        f, buf = self.frames, self.frames % 2
        if self._fence_in:                       # first frame after init / flush: order after whatever the caller queued so far
            e = torch.cuda.Event()
            e.record(torch.cuda.current_stream(b.device))
            R.wait_event(e)
            Cs.wait_event(e)
            self._fence_in = False
        with torch.cuda.stream(R):
            if self.transport == "peer":
                if f >= 2:
                    R.wait_event(ev_cols[buf])   # this rank's cols(f-2) ...
                    self._barrier()              # ... and every other rank's
                b.rows_buf(t, PEER_STORES, buf, int(R.cuda_stream))
                self._barrier()                  # every rank's rows(f) have landed
            else:
                b.rows_buf(t, SEND_BUFFER, buf, int(R.cuda_stream))      # into the (single) send buffer: the previous all-to-all precedes it on this stream
                if f >= 2:
                    R.wait_event(ev_cols[buf])   # the all-to-all overwrites MY receive buffer f % 2: my cols(f-2) must be done with it
                send, recv = b.exchange_tensors(buf)
                self._dist.all_to_all_single(recv, send, group=self.group)
            ev_rows[buf].record(R)
        with torch.cuda.stream(Cs):
            Cs.wait_event(ev_rows[buf])
            b.cols_buf(buf, int(Cs.cuda_stream))
            ev_cols[buf].record(Cs)
        self.frames += 1

    def flush(self):
        """Make the caller's current stream wait for every frame submitted so far (a no-op unless frames are pipelined)."""
        if self.pipelined and self.frames:
            import torch
            cur = torch.cuda.current_stream(self.backend.device)
            cur.wait_stream(self._pipe[0])
            cur.wait_stream(self._pipe[1])
            self._fence_in = True

    def sync(self):
        self.flush()
        self.backend.sync(self.backend.current_stream())

    # ---- outputs ---------------------------------------------------------------------------------------------------
    def download(self, name: str) -> np.ndarray:
        """This rank's column slab: [N][XL] (normal: [N][XL][4])."""
        self.flush()
        return self.backend.download(name, self.backend.current_stream())

    def gather(self, name: str) -> np.ndarray:
        """The full [N][N] image on every rank (test/debug helper: goes through host memory)."""
        mine = self.download(name)
        if self.world == 1:
            return mine
        parts = [None] * self.world
        self._dist.all_gather_object(parts, mine, group=self.group)
        return np.concatenate(parts, axis=1)

    def launches_per_frame(self) -> int:
        """Kernels this rank launches per frame: row, column, normal; above N = 4096 the line decomposition doubles the first two
        unless they run as thread-block clusters (ow_slab_get_line_clusters: bit 0 rows, bit 1 columns)."""
        if self.N <= 4096:
            return 3
        bits = self.backend.line_clusters() if hasattr(self.backend, "line_clusters") else 0
        return 5 - (bits & 1) - ((bits >> 1) & 1)

    def exchange_bytes_per_frame(self) -> int:
        """Bytes this rank sends to OTHER ranks per frame (NVLink traffic per direction)."""
        return self.plan["block_bytes"] * (self.world - 1)
