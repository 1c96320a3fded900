"""Host-side mirror of the reference's sim interface over the C ABI (ctypes).

The reference exposes no API; its "interface" is the FFTOceanWaves class' sim members and private methods
(reference src/main.cpp:553-744, 1630-1646). `FFTOceanWaves` below keeps those names and meanings:

    reference member / method                this class
    m_N, m_L, m_wind_speed, ...              OceanParams fields + FFTOceanWaves.N
    create_textures()        (:1083-1145)    __init__   -> ow_create
    tilde_h0_k()             (:553-583)      tilde_h0_k -> ow_set_noise + ow_init_spectrum
    init()                   (:199-225)      init()
    update(): tilde_h0_t(); butterfly_fft() x3; generate_normal_map()   (:240-244)
                                             update(t)  -> ow_step   (t replaces glfwGetTime(), :599)
    m_dy, m_dx, m_dz, m_normal_map           outputs(slot) device pointers / download(name)

Errors: the reference logs and returns false from init(); here every failing C call raises
OceanWavesError carrying ow_last_error().
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

EXPORTED_SYMBOLS = [
    "ow_create", "ow_destroy", "ow_last_error", "ow_set_params", "ow_set_noise", "ow_init_spectrum", "ow_set_h0",
    "ow_step", "ow_step_multi", "ow_step_multi_timed", "ow_sync", "ow_get_outputs", "ow_download", "ow_download_frame_async",
    "ow_frame_bytes", "ow_set_group_size", "ow_set_streams", "ow_last_launch_count", "ow_gl_register", "ow_gl_step", "ow_gl_unregister",
    "ow_set_noise_seed", "ow_last_group_count", "ow_init_spectrum_cascade", "ow_set_graph", "ow_get_packed", "ow_packed_bytes",
    "ow_download_packed_async", "ow_set_row_kernel", "ow_set_discard_intermediate", "ow_gl_register_packed", "ow_set_column_kernel", "ow_get_kernel_modes", "ow_set_resident_ctas", "ow_set_l2_persist", "ow_set_latency_shapes", "ow_set_frame_kernel", "ow_set_line_clusters", "ow_get_line_clusters",
    "ow_slab_set_line_clusters", "ow_slab_get_line_clusters", "ow_slab_enable_double_buffer", "ow_slab_ipc_handle_buf", "ow_slab_open_peers_buf",
    "ow_slab_rows_buf", "ow_slab_cols_buf", "ow_slab_set_post_ctas", "ow_slab_recv_buffer",
    "ow_slab_create", "ow_slab_destroy", "ow_slab_last_error", "ow_slab_get_info", "ow_slab_init_spectrum_seeded", "ow_slab_ipc_handle",
    "ow_slab_open_peers", "ow_slab_rows", "ow_slab_cols", "ow_slab_local_exchange", "ow_slab_sync", "ow_slab_download",
    "ow_slab_set_column_lines", "ow_sample_points", "ow_sample_points_host", "ow_compose_grid", "ow_set_time_scale", "ow_step_wall_clock",
]

OW_FLAG_JACOBIAN = 0x1
OW_FLAG_EXACT_SINCOS = 0x2
OW_FLAG_FOUR_STEP = 0x4
OW_FLAG_FUSED_NORMALS = 0x8
OW_FLAG_PACKED_F32 = 0x10
OW_FLAG_PACKED_F16 = 0x20
IMAGES = {"dy": 0, "dx": 1, "dz": 2, "normal": 3, "jacobian": 4, "h0k": 5, "h0minusk": 6}


class OceanWavesError(RuntimeError):
    pass


class _Params(C.Structure):
    _fields_ = [("L", C.c_float), ("wind_speed", C.c_float), ("wind_dir", C.c_float * 2), ("amplitude", C.c_float),
                ("suppression", C.c_float), ("choppiness", C.c_float)]


class _Outputs(C.Structure):
    _fields_ = [("N", C.c_int32), ("dy", C.c_void_p), ("dx", C.c_void_p), ("dz", C.c_void_p), ("normal", C.c_void_p),
                ("jacobian", C.c_void_p)]


class _Packed(C.Structure):
    _fields_ = [("N", C.c_int32), ("displacement_texel_bytes", C.c_int32), ("displacement", C.c_void_p), ("normal_xz", C.c_void_p)]


class _SlabInfo(C.Structure):
    _fields_ = [("N", C.c_int32), ("world", C.c_int32), ("rank", C.c_int32), ("pairs_per_rank", C.c_int32),
                ("cols_per_rank", C.c_int32), ("padded_cols", C.c_int32), ("halo", C.c_int32), ("block_bytes", C.c_size_t),
                ("send", C.c_void_p), ("recv", C.c_void_p), ("dy", C.c_void_p), ("dx", C.c_void_p), ("dz", C.c_void_p),
                ("normal", C.c_void_p), ("jacobian", C.c_void_p)]


@dataclass
class OceanParams:
    """Reference defaults: src/main.cpp:1641-1646 (wind 80, A 2, suppression 0.1, dir (1,1), L 1000), :1633."""
    L: float = 1000.0
    wind_speed: float = 80.0
    wind_dir: Sequence[float] = (1.0, 1.0)
    amplitude: float = 2.0
    suppression: float = 0.1
    choppiness: float = 0.75

    def to_c(self) -> _Params:
        p = _Params()
        p.L, p.wind_speed = float(self.L), float(self.wind_speed)
        p.wind_dir[0], p.wind_dir[1] = float(self.wind_dir[0]), float(self.wind_dir[1])
        p.amplitude, p.suppression, p.choppiness = float(self.amplitude), float(self.suppression), float(self.choppiness)
        return p


class _BlendTerm(C.Structure):
    _fields_ = [("slot", C.c_int32), ("weight", C.c_float)]


def lib_path() -> str:
    """In-tree library; OCEANWAVES_LIB points at another build of it (A/B runs of kernel variants)."""
    return os.environ.get("OCEANWAVES_LIB") or os.path.join(_HERE, "lib", "liboceanwaves.so")


_lib = None


def load_library():
    """dlopen liboceanwaves.so and declare the ABI. Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise OceanWavesError(f"{path} is missing: build it with `python fft-ocean-waves_b200/build.py` "
                              "(needs nvcc). There is no CPU fallback.")
    L = C.CDLL(path)
    vp, i32, u32, f32 = C.c_void_p, C.c_int32, C.c_uint32, C.c_float
    L.ow_create.argtypes = [i32, i32, i32, C.POINTER(_Params), i32, u32, C.POINTER(vp)]
    L.ow_destroy.argtypes = [vp]
    L.ow_destroy.restype = None
    L.ow_last_error.argtypes = [vp]
    L.ow_last_error.restype = C.c_char_p
    L.ow_set_params.argtypes = [vp, i32, C.POINTER(_Params)]
    L.ow_set_noise.argtypes = [vp, i32, C.POINTER(C.c_void_p), i32, i32]
    L.ow_init_spectrum.argtypes = [vp]
    L.ow_set_h0.argtypes = [vp, i32, vp, vp]
    L.ow_step.argtypes = [vp, f32, vp]
    L.ow_step_multi.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(f32), vp]
    L.ow_step_multi_timed.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(f32), vp, C.POINTER(f32)]
    L.ow_sync.argtypes = [vp, vp]
    L.ow_get_outputs.argtypes = [vp, i32, C.POINTER(_Outputs)]
    L.ow_download.argtypes = [vp, i32, i32, vp, C.c_size_t, vp]
    L.ow_download_frame_async.argtypes = [vp, i32, vp, C.c_size_t, vp]
    L.ow_frame_bytes.argtypes = [vp]
    L.ow_frame_bytes.restype = C.c_size_t
    L.ow_set_group_size.argtypes = [vp, i32]
    L.ow_set_streams.argtypes = [vp, i32]
    L.ow_last_launch_count.argtypes = [vp]
    L.ow_last_group_count.argtypes = [vp]
    L.ow_gl_register.argtypes = [vp, u32, u32, u32, u32]
    L.ow_gl_step.argtypes = [vp, f32]
    L.ow_gl_unregister.argtypes = [vp]
    L.ow_set_noise_seed.argtypes = [vp, i32, C.c_uint64]
    L.ow_init_spectrum_cascade.argtypes = [vp, i32]
    L.ow_set_graph.argtypes = [vp, i32]
    L.ow_get_packed.argtypes = [vp, i32, C.POINTER(_Packed)]
    L.ow_packed_bytes.argtypes = [vp]
    L.ow_packed_bytes.restype = C.c_size_t
    L.ow_download_packed_async.argtypes = [vp, i32, vp, C.c_size_t, vp]
    L.ow_set_row_kernel.argtypes = [vp, i32]
    L.ow_set_discard_intermediate.argtypes = [vp, i32]
    L.ow_set_column_kernel.argtypes = [vp, i32, i32]
    L.ow_set_resident_ctas.argtypes = [vp, i32, i32]
    L.ow_set_l2_persist.argtypes = [vp, i32]
    L.ow_set_latency_shapes.argtypes = [vp, i32]
    L.ow_set_frame_kernel.argtypes = [vp, i32]
    L.ow_set_line_clusters.argtypes = [vp, i32]
    L.ow_get_line_clusters.argtypes = [vp]
    L.ow_slab_set_line_clusters.argtypes = [vp, i32]
    L.ow_slab_get_line_clusters.argtypes = [vp]
    L.ow_get_kernel_modes.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.ow_gl_register_packed.argtypes = [vp, u32, u32]
    L.ow_sample_points.argtypes = [vp, i32, C.POINTER(_BlendTerm), f32, i32, vp, vp, vp]
    L.ow_sample_points_host.argtypes = [vp, i32, C.POINTER(_BlendTerm), f32, i32, vp, vp, vp]
    L.ow_compose_grid.argtypes = [vp, i32, C.POINTER(_BlendTerm), f32, i32, f32, f32, f32, vp, vp, vp]
    L.ow_set_time_scale.argtypes = [vp, f32, f32]
    L.ow_step_wall_clock.argtypes = [vp, C.c_double, vp]
    L.ow_slab_create.argtypes = [i32, i32, i32, C.POINTER(_Params), i32, u32, C.POINTER(vp)]
    L.ow_slab_destroy.argtypes = [vp]
    L.ow_slab_destroy.restype = None
    L.ow_slab_last_error.argtypes = [vp]
    L.ow_slab_last_error.restype = C.c_char_p
    L.ow_slab_get_info.argtypes = [vp, C.POINTER(_SlabInfo)]
    L.ow_slab_init_spectrum_seeded.argtypes = [vp, C.c_uint64]
    L.ow_slab_ipc_handle.argtypes = [vp, vp, C.c_size_t]
    L.ow_slab_open_peers.argtypes = [vp, vp, C.c_size_t]
    L.ow_slab_rows.argtypes = [vp, f32, i32, vp]
    L.ow_slab_enable_double_buffer.argtypes = [vp]
    L.ow_slab_ipc_handle_buf.argtypes = [vp, i32, vp, C.c_size_t]
    L.ow_slab_open_peers_buf.argtypes = [vp, i32, vp, C.c_size_t]
    L.ow_slab_rows_buf.argtypes = [vp, f32, i32, i32, vp]
    L.ow_slab_cols_buf.argtypes = [vp, i32, vp]
    L.ow_slab_set_post_ctas.argtypes = [vp, i32]
    L.ow_slab_set_column_lines.argtypes = [vp, i32]
    L.ow_slab_recv_buffer.argtypes = [vp, i32, C.POINTER(vp)]
    L.ow_slab_cols.argtypes = [vp, vp]
    L.ow_slab_local_exchange.argtypes = [vp, vp]
    L.ow_slab_sync.argtypes = [vp, vp]
    L.ow_slab_download.argtypes = [vp, i32, vp, C.c_size_t, vp]
    _lib = L
    return L


def default_noise() -> np.ndarray:
    """R channel of the reference's data/noise/LDR_LLL1_{0..3}.png (MIT, diharaw/fft-ocean-waves), (4,256,256) u8."""
    return np.fromfile(os.path.join(_HERE, "data", "noise_LDR_LLL1_R.u8"), dtype=np.uint8).reshape(4, 256, 256)


class FFTOceanWaves:
    """One context = `len(cascades)` independent N x N patches on one GPU (see module docstring)."""

    def __init__(self, N: int = 256, cascades: Optional[Sequence[OceanParams]] = None, n_slots: Optional[int] = None,
                 device: int = 0, jacobian: bool = False, exact_sincos: bool = False, four_step: bool = False,
                 fused_normals: bool = False, packed: Optional[str] = None):
        self._lib = load_library()
        self._h = C.c_void_p()
        self.N = int(N)
        self.cascades = list(cascades) if cascades is not None else [OceanParams()]
        self.n_slots = int(n_slots) if n_slots is not None else len(self.cascades)
        self.jacobian = bool(jacobian)
        self.device = int(device)
        if packed not in (None, "f32", "f16"):
            raise ValueError("packed must be None, 'f32' or 'f16'")
        self.packed = packed
        arr = (_Params * len(self.cascades))(*[c.to_c() for c in self.cascades])
        rc = self._lib.ow_create(self.N, len(self.cascades), self.n_slots, arr, int(device),
                                 (OW_FLAG_JACOBIAN if jacobian else 0) | (OW_FLAG_EXACT_SINCOS if exact_sincos else 0)
                                 | (OW_FLAG_FOUR_STEP if four_step else 0) | (OW_FLAG_FUSED_NORMALS if fused_normals else 0)
                                 | {None: 0, "f32": OW_FLAG_PACKED_F32, "f16": OW_FLAG_PACKED_F16}[packed],
                                 C.byref(self._h))
        if rc != 0:
            msg = self._lib.ow_last_error(None)
            self._h = C.c_void_p()
            raise OceanWavesError(f"ow_create failed ({rc}): {msg.decode() if msg else ''}")

    # ---- plumbing -------------------------------------------------------------------------------
    def _check(self, rc: int, what: str):
        if rc != 0:
            msg = self._lib.ow_last_error(self._h)
            raise OceanWavesError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.ow_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- init path (reference init(): src/main.cpp:218-220) ----------------------------------------
    def set_noise(self, noise: np.ndarray, cascade: int = -1):
        noise = np.ascontiguousarray(noise, dtype=np.uint8)
        if noise.ndim != 3 or noise.shape[0] != 4:
            raise ValueError("noise must have shape (4, h, w)")
        ptrs = (C.c_void_p * 4)(*[noise[j].ctypes.data for j in range(4)])
        self._check(self._lib.ow_set_noise(self._h, int(cascade), ptrs, noise.shape[2], noise.shape[1]), "ow_set_noise")

    def set_noise_seed(self, seed: int, cascade: int = -1):
        """Device-generated Philox noise (BASELINE config C5), identical to what the slab path draws."""
        self._check(self._lib.ow_set_noise_seed(self._h, int(cascade), C.c_uint64(int(seed))), "ow_set_noise_seed")

    def tilde_h0_k(self):
        self._check(self._lib.ow_init_spectrum(self._h), "ow_init_spectrum")

    def init(self, noise: Optional[np.ndarray] = None):
        self.set_noise(default_noise() if noise is None else noise)
        self.tilde_h0_k()
        return True

    def set_params(self, cascade: int, p: OceanParams):
        c = p.to_c()
        self._check(self._lib.ow_set_params(self._h, int(cascade), C.byref(c)), "ow_set_params")
        self.cascades[cascade] = p

    def tilde_h0_k_cascade(self, cascade: int):
        """Re-generate ONE cascade's initial spectrum (after set_params on it)."""
        self._check(self._lib.ow_init_spectrum_cascade(self._h, int(cascade)), "ow_init_spectrum_cascade")

    def set_h0(self, h0k: np.ndarray, h0minusk: np.ndarray, cascade: int = 0):
        a = np.ascontiguousarray(h0k, np.float32).reshape(self.N, self.N, 2)
        b = np.ascontiguousarray(h0minusk, np.float32).reshape(self.N, self.N, 2)
        self._check(self._lib.ow_set_h0(self._h, int(cascade), a.ctypes.data, b.ctypes.data), "ow_set_h0")

    # ---- per frame (reference update(): src/main.cpp:240-244) ---------------------------------------
    def update(self, t: float, stream: int = 0):
        self._check(self._lib.ow_step(self._h, float(t), C.c_void_p(stream or None)), "ow_step")

    def set_time_scale(self, scale: float = 1.0, offset: float = 0.0):
        """The demo's clock (src/main.cpp:599, t = glfwGetTime()): update_wall_clock(w) evaluates t = offset + scale * w."""
        self._check(self._lib.ow_set_time_scale(self._h, float(scale), float(offset)), "ow_set_time_scale")

    def update_wall_clock(self, wall_seconds: float, stream: int = 0):
        self._check(self._lib.ow_step_wall_clock(self._h, float(wall_seconds), C.c_void_p(stream or None)), "ow_step_wall_clock")

    # ---- multi-cascade composition (the consumer's sum grid_tes.glsl:60-64 over cascades) ---
    @staticmethod
    def _terms(slots: Sequence[int], weights: Sequence[float]):
        if len(slots) != len(weights):
            raise ValueError("one weight per slot")
        arr = (_BlendTerm * max(len(slots), 1))()
        for i, (s, w) in enumerate(zip(slots, weights)):
            arr[i].slot, arr[i].weight = int(s), float(w)
        return arr

    def sample_points(self, xz: np.ndarray, slots: Sequence[int], weights: Sequence[float], displacement_scale: float = 1.0,
                      stream: int = 0) -> dict:
        """Blended vertex offset and normal at world positions xz [n][2] (metres): dict(offset [n][4], normal [n][4])."""
        pts = np.ascontiguousarray(xz, np.float32).reshape(-1, 2)
        out = np.empty((pts.shape[0], 8), np.float32)
        self._check(self._lib.ow_sample_points_host(self._h, len(slots), self._terms(slots, weights), float(displacement_scale), pts.shape[0],
                                                    pts.ctypes.data, out.ctypes.data, C.c_void_p(stream or None)), "ow_sample_points_host")
        return dict(offset=out[:, :4].copy(), normal=out[:, 4:].copy())

    def compose_grid(self, M: int, origin: Sequence[float], extent: float, slots: Sequence[int], weights: Sequence[float],
                     displacement_scale: float = 1.0, stream: int = 0) -> dict:
        """The same sum on M x M world positions origin + (i + 0.5) * extent / M, evaluated into two device images and downloaded."""
        import torch
        dev = torch.device("cuda", self.device)
        off = torch.empty((M, M, 4), dtype=torch.float32, device=dev)
        nrm = torch.empty((M, M, 4), dtype=torch.float32, device=dev)
        self._check(self._lib.ow_compose_grid(self._h, len(slots), self._terms(slots, weights), float(displacement_scale), int(M), float(origin[0]),
                                              float(origin[1]), float(extent), off.data_ptr(), nrm.data_ptr(), C.c_void_p(stream or None)), "ow_compose_grid")
        self.sync(stream)
        return dict(offset=off.cpu().numpy(), normal=nrm.cpu().numpy())

    def update_multi(self, cascade_of_slot: Sequence[int], time_of_slot: Sequence[float], stream: int = 0):
        n = len(cascade_of_slot)
        ci = (C.c_int32 * n)(*[int(v) for v in cascade_of_slot])
        ti = (C.c_float * n)(*[float(v) for v in time_of_slot])
        self._check(self._lib.ow_step_multi(self._h, n, ci, ti, C.c_void_p(stream or None)), "ow_step_multi")

    def update_multi_timed(self, cascade_of_slot: Sequence[int], time_of_slot: Sequence[float], stream: int = 0):
        """Synchronous; returns (row_ms, col_ms, normal_ms) measured with CUDA events around each kernel."""
        n = len(cascade_of_slot)
        ci = (C.c_int32 * n)(*[int(v) for v in cascade_of_slot])
        ti = (C.c_float * n)(*[float(v) for v in time_of_slot])
        ms = (C.c_float * 3)()
        self._check(self._lib.ow_step_multi_timed(self._h, n, ci, ti, C.c_void_p(stream or None), ms), "ow_step_multi_timed")
        return float(ms[0]), float(ms[1]), float(ms[2])

    def sync(self, stream: int = 0):
        self._check(self._lib.ow_sync(self._h, C.c_void_p(stream or None)), "ow_sync")

    def set_group_size(self, g: int):
        self._check(self._lib.ow_set_group_size(self._h, int(g)), "ow_set_group_size")

    def set_streams(self, n: int):
        self._check(self._lib.ow_set_streams(self._h, int(n)), "ow_set_streams")

    def set_graph(self, enabled: bool):
        self._check(self._lib.ow_set_graph(self._h, int(bool(enabled))), "ow_set_graph")

    def set_row_kernel(self, mode: int):
        self._check(self._lib.ow_set_row_kernel(self._h, int(mode)), "ow_set_row_kernel")

    def set_column_kernel(self, mode: int, fused: int = -1):
        self._check(self._lib.ow_set_column_kernel(self._h, int(mode), int(fused)), "ow_set_column_kernel")

    def set_resident_ctas(self, row_per_sm: int = 0, col_per_sm: int = 0):
        self._check(self._lib.ow_set_resident_ctas(self._h, int(row_per_sm), int(col_per_sm)), "ow_set_resident_ctas")

    def set_line_clusters(self, mode: int = -1):
        self._check(self._lib.ow_set_line_clusters(self._h, int(mode)), "ow_set_line_clusters")

    def line_clusters(self) -> int:
        return int(self._lib.ow_get_line_clusters(self._h))

    def set_frame_kernel(self, mode: int):
        self._check(self._lib.ow_set_frame_kernel(self._h, int(mode)), "ow_set_frame_kernel")

    def set_latency_shapes(self, on: bool = True):
        self._check(self._lib.ow_set_latency_shapes(self._h, int(bool(on))), "ow_set_latency_shapes")

    def set_l2_persist(self, mode: int = -1):
        self._check(self._lib.ow_set_l2_persist(self._h, int(mode)), "ow_set_l2_persist")

    def kernel_modes(self) -> dict:
        r, k, f = C.c_int32(), C.c_int32(), C.c_int32()
        self._check(self._lib.ow_get_kernel_modes(self._h, C.byref(r), C.byref(k), C.byref(f)), "ow_get_kernel_modes")
        return {"row": r.value, "column": k.value, "fused": bool(f.value)}

    def set_discard_intermediate(self, on: bool):
        self._check(self._lib.ow_set_discard_intermediate(self._h, int(bool(on))), "ow_set_discard_intermediate")

    def last_launch_count(self) -> int:
        return int(self._lib.ow_last_launch_count(self._h))

    def last_group_count(self) -> int:
        return int(self._lib.ow_last_group_count(self._h))

    # ---- outputs -----------------------------------------------------------------------------------
    def outputs(self, slot: int = 0) -> _Outputs:
        o = _Outputs()
        self._check(self._lib.ow_get_outputs(self._h, int(slot), C.byref(o)), "ow_get_outputs")
        return o

    def frame_bytes(self) -> int:
        return int(self._lib.ow_frame_bytes(self._h))

    def download(self, name: str, index: int = 0, stream: int = 0) -> np.ndarray:
        n = self.N
        shape = {"dy": (n, n), "dx": (n, n), "dz": (n, n), "jacobian": (n, n), "normal": (n, n, 4), "h0k": (n, n, 2),
                 "h0minusk": (n, n, 2)}[name]
        out = np.empty(shape, np.float32)
        self._check(self._lib.ow_download(self._h, int(index), IMAGES[name], out.ctypes.data, out.nbytes,
                                          C.c_void_p(stream or None)), "ow_download")
        return out

    def download_frame_async(self, slot: int, host_ptr: int, nbytes: int, stream: int = 0):
        self._check(self._lib.ow_download_frame_async(self._h, int(slot), C.c_void_p(host_ptr), nbytes,
                                                      C.c_void_p(stream or None)), "ow_download_frame_async")

    # ---- packed output set (OW_FLAG_PACKED_*) ---------------------------------------
    def packed_bytes(self) -> int:
        return int(self._lib.ow_packed_bytes(self._h))

    def download_packed_async(self, slot: int, host_ptr: int, nbytes: int, stream: int = 0):
        self._check(self._lib.ow_download_packed_async(self._h, int(slot), C.c_void_p(host_ptr), nbytes,
                                                       C.c_void_p(stream or None)), "ow_download_packed_async")

    def download_packed(self, slot: int = 0, stream: int = 0) -> dict:
        """Raw packed images of `slot`: 'displacement' (N,N,4) float32 or float16 = (dx,dy,dz,J); 'normal_xz' (N,N,2) int16 (SNORM)."""
        nb = self.packed_bytes()
        buf = np.empty(nb, np.uint8)
        self.download_packed_async(slot, buf.ctypes.data, nb, stream)
        self.sync(stream)
        n = self.N
        tb = 8 if self.packed == "f16" else 16
        disp = buf[:n * n * tb].view(np.float16 if self.packed == "f16" else np.float32).reshape(n, n, 4)
        nxz = buf[n * n * tb:].view(np.int16).reshape(n, n, 2)
        return {"displacement": disp, "normal_xz": nxz}

    @staticmethod
    def decode_packed(pk: dict) -> dict:
        """What the consumer's shader does with the packed set: dx,dy,dz,J and the unit normal."""
        d = pk["displacement"].astype(np.float32)
        xz = np.maximum(pk["normal_xz"].astype(np.float32) / 32767.0, -1.0)
        ny = np.sqrt(np.maximum(0.0, 1.0 - xz[..., 0] ** 2 - xz[..., 1] ** 2))
        normal = np.stack([xz[..., 0], ny, xz[..., 1], np.ones_like(ny)], axis=-1)
        return {"dx": d[..., 0], "dy": d[..., 1], "dz": d[..., 2], "jacobian": d[..., 3], "normal": normal}

    def frame(self, t: float, slot: int = 0) -> dict:
        """Convenience for tests: update(t) then download every image of `slot`."""
        self.update(t)
        self.sync()
        names = ["dy", "dx", "dz", "normal"] + (["jacobian"] if self.jacobian else [])
        return {k: self.download(k, slot) for k in names}
