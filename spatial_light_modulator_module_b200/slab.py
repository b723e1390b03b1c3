"""Gerchberg-Saxton on ONE very large square plane split by rows over the ranks of a process group
(BASELINE config 5: a 16384^2 grid on 2/4/8 GPUs): slab-decomposed 2-D transforms with all-to-all
transposes.

Rank p owns rows [p*h, (p+1)*h) of the N x N field, h = N / world.  One GS iteration
(reference: algorithms.py:30-38) is

    row pass      finish ifft2 along the rows, B = exp(1j*angle(A)), start fft2 along the rows   (local)
    exchange      pack blocks -> all-to-all -> every rank holds h whole COLUMNS as contiguous lines
    Fourier pass  finish fft2 along those lines, amplitude replacement + error sums, start ifft2   (local)
    all-reduce    max |C|^2 and the three error sums (4 doubles)
    exchange      all-to-all back -> unpack into the row slab

i.e. two all-to-alls and one tiny all-reduce per iteration; the target is exchanged once.  All array
work is done by kernels of libslmholo (``slm_rows_*``, ``slm_transpose_blocks``); torch.distributed
(NCCL over NVLink on GPUs) moves the blocks.  The loop is closed on the host (one scalar read per
iteration), which is negligible next to a multi-millisecond iteration at these sizes.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import numpy as np

from . import _ffi, host_logic as hl
from .engine import Engine, _PREC, _NP_REAL, _NP_CPLX


class SlabEngine(Engine):
    """Row-slab context: ``rows`` = N / world lines of N points on this rank's device."""

    def __init__(self, n: int, world: int, rank: int, precision: str = "fp32", device=None, stream=None, group=None):
        if n % world:
            raise ValueError("the grid size must be divisible by the number of ranks")
        self.n, self.world, self.rank, self.rows = int(n), int(world), int(rank), int(n) // int(world)
        self.precision = precision
        self.shape = (self.rows, self.n)
        self.max_batch = 1
        self.real_dtype, self.complex_dtype = _NP_REAL[precision], _NP_CPLX[precision]
        self.group = group
        self._lib = self._load_library()
        self._device_index = self._mem_init(device)
        self._stream_handle = self._mem_stream(stream)
        ctx = C.c_void_p()
        rc = self._lib.slm_rows_create(C.byref(ctx), self._device_index, self.rows, self.n, _PREC[precision], self._stream_handle)
        if rc == -2:
            raise ValueError(self._lib.slm_last_error().decode())
        _ffi.check(self._lib, rc)
        self._ctx = ctx
        self._amp_lut = hl.amplitude_lut()

    # ---- hooks the test-suite overrides together with the _mem_* ones ------------------------------------
    def _as_torch(self, buf):
        """torch view of a device buffer, as real numbers (NCCL has no complex types)."""
        torch = self._torch
        return torch.view_as_real(buf) if buf.is_complex() else buf

    def _dist(self):
        import torch.distributed as dist
        return dist

    # ---- collectives ---------------------------------------------------------------------------------------
    def _all_to_all(self, send, recv):
        if self.world == 1:
            self._copy(send, recv)
            return
        dist = self._dist()
        self._sync()
        dist.all_to_all_single(self._as_torch(recv), self._as_torch(send), group=self.group)

    def _all_reduce(self, values: np.ndarray, op: str) -> np.ndarray:
        if self.world == 1:
            return values
        import torch
        dist = self._dist()
        t = torch.from_numpy(np.ascontiguousarray(values, dtype=np.float64))
        if dist.get_backend(self.group) == "nccl":
            t = t.to(self._dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    def _sync(self):
        self._stream.synchronize()

    def _copy(self, src, dst):
        dst.copy_(src)

    # ---- kernels ----------------------------------------------------------------------------------------------
    def _transpose(self, src, dst, elem_bytes: int, from_exchange: bool):
        self._check(self._lib.slm_transpose_blocks(self._ctx, self._mem_ptr(src), self._mem_ptr(dst), self.rows, self.n,
                                                   elem_bytes, int(from_exchange)))

    def _exchange(self, slab, send, recv, elem_bytes):
        """row slab -> exchange layout of the columns this rank owns (in ``recv``)."""
        if self.world == 1:
            self._transpose(slab, recv, elem_bytes, False)
            return
        self._transpose(slab, send, elem_bytes, False)
        self._all_to_all(send, recv)

    def _exchange_back(self, lines, recv, slab, elem_bytes):
        """exchange layout (after a pass over the owned columns) -> row slab."""
        if self.world == 1:
            self._transpose(lines, slab, elem_bytes, True)
            return
        self._all_to_all(lines, recv)
        self._transpose(recv, slab, elem_bytes, True)

    def _rows_fft(self, src, dst, inverse, block_in=0, block_out=0, u8=None):
        lut = self._dp(self._amp_lut) if u8 is not None else None
        self._check(self._lib.slm_rows_fft(self._ctx, self._mem_ptr(src), self._mem_ptr(u8), lut, self._mem_ptr(dst),
                                           int(inverse), int(block_in), int(block_out)))

    def _partial_totals(self, partial) -> Tuple[float, np.ndarray]:
        p = self.to_host(partial).reshape(self.rows, 4)
        mx = self._all_reduce(np.array([p[:, 0].max()]), "max")[0]
        sums = self._all_reduce(p[:, 1:].sum(axis=0), "sum")
        return float(mx), sums

    # ---- Gerchberg-Saxton on the distributed plane --------------------------------------------------------------
    def gs(self, target_slab, max_loops: int, tolerance: float = 0.0, want_expected: bool = True, on_device: bool = False):
        """``target_slab``: this rank's uint8 rows [rows, N] of the target (host array or device buffer).
        Returns ``(hologram_slab float64 [rows, N], expected_slab or None, error_evolution list)``; the error
        curve is identical on every rank.  ``on_device`` leaves hologram / expected in device memory."""
        if max_loops < 1:
            raise UnboundLocalError("cannot access local variable 'expected_outcome' where it is not associated with a value")
        if self._mem_is_device(target_slab):
            T, t = target_slab, None
            if tuple(T.shape) != self.shape or self._mem_np_dtype(T) != np.uint8:
                raise ValueError(f"target slab must be uint8 {self.shape}")
            local_max = float(self.to_host(T.max() if hasattr(T, "is_cuda") else np.asarray(T).max()))
        else:
            t = np.ascontiguousarray(target_slab)
            if t.dtype != np.uint8 or t.shape != self.shape:
                raise ValueError(f"target slab must be uint8 {self.shape}")
            T, local_max = self._mem_upload(t), float(t.max())
        h, n, cs = self.rows, self.n, np.dtype(self.complex_dtype).itemsize
        norm = float(self._all_reduce(np.array([local_max]), "max")[0])
        Tx_send, Tx = self._mem_empty((self.world, h, h), np.uint8), self._mem_empty((self.world, h, h), np.uint8)
        self._exchange(T, Tx_send, Tx, 1)                                    # target columns, once
        X = self._mem_empty(self.shape, self.complex_dtype)                  # row slab
        S = self._mem_empty((self.world, h, h), self.complex_dtype)          # exchange layout (send / lines)
        Rv = self._mem_empty((self.world, h, h), self.complex_dtype)         # exchange layout (receive)
        partial = self._mem_empty((h, 4), np.float64)
        # A = ifft2(sqrt(T))  (algorithms.py:27), unnormalised: only its phase is used.  For 8-bit targets the
        # reference computes it (and the first phasor) in complex64, so an fp64 plane borrows an fp32 engine.
        if self.precision == "fp32":
            self._setup_field(T, X, S, Rv)
            A0, field_kind = X, 1
        else:
            helper = type(self)(self.n, self.world, self.rank, "fp32", self._device_index, None, self.group)
            A0 = helper._mem_empty(self.shape, np.complex64)
            helper._setup_field(T, A0, helper._mem_empty((self.world, h, h), np.complex64),
                                helper._mem_empty((self.world, h, h), np.complex64))
            helper._sync()
            helper.close()
            field_kind = 2
        Y = self._mem_empty(self.shape, self.complex_dtype)
        hw = float(n) * float(n)
        errors: List[float] = []
        s_prev = None
        inten = self._mem_empty((self.world, h, h), np.float64) if want_expected else None
        src, field = A0, field_kind
        for k in range(max_loops):
            cur = X if src is Y else Y                                        # receives the row-transformed B
            self._check(self._lib.slm_rows_gs_row_pass(self._ctx, self._mem_ptr(src), self._mem_ptr(cur), None, int(field), 0, None))
            self._exchange(cur, S, Rv, cs)
            if s_prev is None:                                                # exact scale of iteration 0: max pre-pass
                self._fourier(Rv, S, Tx, 1.0, partial, None)
                mx, _ = self._partial_totals(partial)
                s_prev = norm / mx
            last = k == max_loops - 1
            self._fourier(Rv, S, Tx, s_prev, partial, inten if (want_expected and (last or tolerance > 0)) else None)
            mx, (a, b, c) = self._partial_totals(partial)
            s = norm / mx                                                     # algorithms.py:37
            s0u = float(self.real_dtype(s_prev))                              # the scale the kernel used
            dl = s / s0u - 1.0 if s0u != 0.0 else 0.0
            err = (a + 2.0 * dl * b + dl * dl * c) / hw                       # algorithms.py:38,162
            errors.append(np.float64(err))
            s_prev = s
            self._exchange_back(S, Rv, cur, cs)                               # D with the columns inverse-transformed
            src, field = cur, 0
            if not (err > tolerance):
                break
        holo = self._mem_empty(self.shape, np.float64)
        self._check(self._lib.slm_rows_gs_row_pass(self._ctx, self._mem_ptr(src), None, None, 0, 1, self._mem_ptr(holo)))
        expected = None
        if want_expected:
            recv_i = self._mem_empty((self.world, h, h), np.float64)
            exp_slab = self._mem_empty(self.shape, np.float64)
            self._all_to_all(inten, recv_i)
            self._transpose(recv_i, exp_slab, 8, True)
            if on_device:
                exp_slab *= s_prev                                            # expected_outcome *= norm / max, :37
                expected = exp_slab
            else:
                expected = self.to_host(exp_slab) * s_prev
        return (holo if on_device else self.to_host(holo)), expected, errors

    def _setup_field(self, T, X, S, Rv):
        """A = ifft2(amplitude) of the distributed target into the row slab X (this engine's precision)."""
        h, cs = self.rows, np.dtype(self.complex_dtype).itemsize
        self._rows_fft(None, X, True, u8=T)
        self._exchange(X, S, Rv, cs)
        self._rows_fft(Rv, S, True, block_in=h, block_out=h)
        self._exchange_back(S, Rv, X, cs)

    def _fourier(self, lines_in, lines_out, Tx, s_prev, partial, inten):
        self._check(self._lib.slm_rows_gs_fourier_pass(self._ctx, self._mem_ptr(lines_in), self._mem_ptr(lines_out), self.rows,
                                                       self._mem_ptr(Tx), self._dp(self._amp_lut), float(s_prev),
                                                       self._mem_ptr(partial), self._mem_ptr(inten)))


def gerchberg_saxton_slab(target, max_loops: int, tolerance: float = 0.0, precision: str = "fp32", want_expected: bool = True,
                          engine_factory=None):
    """GS hologram of one large square uint8 ``target`` (every rank passes the same array, or only its
    own rows via ``target[lo:hi]`` semantics handled here) on all ranks of the default process group.
    Returns this rank's row slab of (hologram, expected) and the error curve."""
    try:
        import torch.distributed as dist
        world, rank = (dist.get_world_size(), dist.get_rank()) if dist.is_available() and dist.is_initialized() else (1, 0)
    except Exception:
        world, rank = 1, 0
    target = np.asarray(target)
    n = target.shape[1]
    lo, hi = rank * (n // world), (rank + 1) * (n // world)
    slab = target[lo:hi] if target.shape[0] == n else target
    eng = (engine_factory or SlabEngine)(n, world, rank, precision)
    try:
        return eng.gs(slab, max_loops, tolerance, want_expected)
    finally:
        eng.close()
