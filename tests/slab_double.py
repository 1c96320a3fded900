"""NumPy stand-in for the ow_slab_* C ABI (TEST INFRASTRUCTURE): lets the CPU tests drive the host side of the slab
path — partition, block layout with halo columns, transports, ordering — through a real torch.distributed group
(gloo, world_size 2) on a box without a GPU. It restates what the CUDA kernels store (csrc/ow_kernels.cuh: SlabRows,
SlabSink, the column kernel's Hermitian unpacking) in fp64 NumPy; the product never imports it."""
import numpy as np
import torch

import fft_ocean_waves_b200 as fow
from oracle import numpy_ref as R


class NumpySlabBackend:
    def __init__(self, N, world, rank, h0k, h0minusk, L, choppiness=1.0):
        self.plan = dict(fow.slab_plan(N, world, rank))
        self.N, self.world, self.rank, self.L, self.lam = N, world, rank, float(L), float(choppiness)
        rows = self.plan["h0_rows"]                       # this rank only ever touches the h0 rows it owns
        self._rows = rows
        self.h0k = {v: h0k[v].astype(np.complex128) for v in rows}
        self.h0m = {v: h0minusk[v].astype(np.complex128) for v in rows}
        n = self.plan["block_elems"] * world
        self.send = torch.zeros(2 * n, dtype=torch.float32)
        self.recv = torch.zeros(2 * n, dtype=torch.float32)
        self.out = {}

    def close(self):
        pass

    def init_spectrum(self, seed):
        pass

    # ---- rows: S = H + conj(H(-k)) on the owned row pairs, row IFFT, transposed blocks with halo columns ----------
    def _spectrum_rows(self, v, t):
        N = self.N
        full_a = np.zeros((N, N), np.complex128)
        full_b = np.zeros((N, N), np.complex128)
        full_a[v], full_b[v] = self.h0k[v], self.h0m[v]
        H = R.spectra(full_a, full_b, N, self.L, t)
        return [h[v] for h in H]

    def rows(self, t, transport, stream=0):
        assert transport == 0, "the CPU stand-in only has the send-buffer transport"
        N, P = self.N, self.plan
        PL, XL, XH, halo = P["pairs_per_rank"], P["cols_per_rank"], P["padded_cols"], P["halo"]
        idx = (-np.arange(N)) % N
        send = self.send.numpy().view(np.complex64).reshape(self.world, PL, 3, XH)
        for pl in range(PL):
            p = P["first_pair"] + pl
            va, vb = (0, N // 2) if p == 0 else (p, N - p)
            Ha, Hb = self._spectrum_rows(va, t), self._spectrum_rows(vb, t)
            for c in range(3):
                if p == 0:      # rows 0 and N/2 mirror onto themselves; both transforms are real: packed as re/im
                    line = (np.fft.ifft(Ha[c] + np.conj(Ha[c][idx])) * N).real + 1j * (np.fft.ifft(Hb[c] + np.conj(Hb[c][idx])) * N).real
                else:
                    line = np.fft.ifft(Ha[c] + np.conj(Hb[c][idx])) * N
                for h in range(self.world):
                    cols = (h * XL - halo + np.arange(XH)) % N
                    send[h, pl, c] = line[cols]

    def exchange_tensors(self):
        return self.send, self.recv

    def local_exchange(self, stream=0):
        self.recv.copy_(self.send)

    def barrier_token(self):
        return torch.zeros(1)

    def current_stream(self):
        return 0

    def sync(self, stream=0):
        pass

    # ---- cols: rebuild the conjugate rows, column IFFT, inversion sign/scale, stencils on the padded slab ------------
    def cols(self, stream=0):
        N, P = self.N, self.plan
        XL, XH, halo = P["cols_per_rank"], P["padded_cols"], P["halo"]
        recv = self.recv.numpy().view(np.complex64).reshape(N // 2, 3, XH).astype(np.complex128)
        xg = (P["first_col"] - halo + np.arange(XH)) % N
        sign = np.where((xg[None, :] + np.arange(N)[:, None]) % 2 == 0, 1.0, -1.0)
        d = []
        for c in range(3):
            I = recv[:, c, :]
            full = np.empty((N, XH), np.complex128)
            full[1:N // 2] = I[1:]
            full[N // 2 + 1:] = np.conj(I[1:][::-1])
            full[0], full[N // 2] = I[0].real, I[0].imag
            d.append(sign * 0.5 * np.fft.ifft(full, axis=0).real / N)      # (-1)^(x+y) * Re(.)/N^2, 1/2 from the Hermitian sum
        inner = slice(halo, halo + XL)
        self.out = dict(dy=d[0][:, inner], dx=d[1][:, inner], dz=d[2][:, inner],
                        normal=R.normal_map(d[0])[:, inner], jacobian=R.jacobian(d[1], d[2], self.L, self.lam)[:, inner])

    def download(self, name, stream=0):
        return np.ascontiguousarray(self.out[name], np.float32)
