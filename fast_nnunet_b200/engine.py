"""Python handle of a libfnnu network engine + thin wrappers of the sliding-window operators.
torch is used for device memory and streams only; every computation below is a libfnnu call."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from .program import Program


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _np_ptr(a: Optional[np.ndarray]):
    return C.c_void_p(0 if a is None else a.ctypes.data)


def require_cuda_device(device: torch.device):
    lib = _lib.load()
    if device.type != 'cuda' or not torch.cuda.is_available():
        raise RuntimeError('fast_nnunet_b200 runs on a CUDA (sm_100a) device only; there is no CPU path')
    with torch.cuda.device(device):
        if not lib.fnnu_device_ok():
            raise RuntimeError('libfnnu: ' + lib.fnnu_last_error().decode())


class NetworkEngine:
    """A compiled network program for up to `max_batch` patches per launch sequence."""

    def __init__(self, program: Program, max_batch: int, device: torch.device, share_workspace_with=None):
        """`share_workspace_with`: another engine of the SAME program shape (another fold): the activation workspace
        is shared, only the packed parameters are separate."""
        self.lib = _lib.load()
        require_cuda_device(device)
        self.program = program
        self.max_batch = int(max_batch)
        self.device = device
        nb, no = len(program.buffers), len(program.ops)
        self._bufs = (_lib.BufferDesc * nb)()
        for i, (dims, ch) in enumerate(program.buffers):
            self._bufs[i].dims = _lib.i3(dims)
            self._bufs[i].channels = ch
        self._ops = (_lib.OpDesc * no)()
        self._keep = []
        for i, o in enumerate(program.ops):
            d = self._ops[i]
            d.op, d.src, d.src_coff, d.src2, d.src2_coff = o.op, o.src, o.src_coff, o.src2, o.src2_coff
            d.dst, d.dst_coff, d.cin, d.cout = o.dst, o.dst_coff, o.cin, o.cout
            d.kernel, d.stride = _lib.i3(o.kernel), _lib.i3(o.stride)
            d.has_bias, d.has_norm = int(o.bias is not None), int(o.has_norm)
            d.norm_eps, d.act_slope = float(o.eps), float(o.act_slope)
            for name in ('weight', 'bias', 'gamma', 'beta'):
                arr = getattr(o, name)
                if arr is not None:
                    arr = np.ascontiguousarray(arr, dtype=np.float32)
                    self._keep.append(arr)
                    setattr(d, name, arr.ctypes.data)
                else:
                    setattr(d, name, None)
        pb, wb = C.c_size_t(0), C.c_size_t(0)
        _lib.check(self.lib.fnnu_engine_sizes(self._bufs, nb, self._ops, no, self.max_batch, C.byref(pb), C.byref(wb)))
        self.param_bytes, self.workspace_bytes = pb.value, wb.value
        with torch.cuda.device(device):
            self.param_arena = torch.empty(self.param_bytes + 256, dtype=torch.uint8, device=device)
            if share_workspace_with is not None:
                assert share_workspace_with.workspace_bytes == self.workspace_bytes
                self.workspace = share_workspace_with.workspace
            else:
                self.workspace = torch.zeros(self.workspace_bytes + 256, dtype=torch.uint8, device=device)
            self._pa = (self.param_arena.data_ptr() + 255) // 256 * 256
            self._ws = (self.workspace.data_ptr() + 255) // 256 * 256
            h = C.c_void_p(0)
            _lib.check(self.lib.fnnu_engine_create(self._bufs, nb, self._ops, no, self.max_batch,
                                                   C.c_void_p(self._pa), self.param_bytes, C.c_void_p(self._ws),
                                                   self.workspace_bytes, _lib.stream_ptr(), C.byref(h)))
            torch.cuda.current_stream().synchronize()
        self.handle = h
        self._keep = []   # host parameter copies are no longer needed

    def __del__(self):
        h = getattr(self, 'handle', None)
        if h:
            self.lib.fnnu_engine_destroy(h)
            self.handle = None

    def buffer_ptr(self, index: int) -> int:
        return self.lib.fnnu_engine_buffer(self.handle, index)

    def buffer_tensor(self, index: int, batch: int) -> torch.Tensor:
        """fp16 view [batch, d0, d1, d2, C] of an activation buffer (tests / debugging)."""
        dims, ch = self.program.buffers[index]
        n = batch * int(np.prod(dims)) * ch
        off = self.buffer_ptr(index) - self.workspace.data_ptr()
        return self.workspace[off:off + 2 * n].view(torch.float16).view(batch, *dims, ch)

    def stats_tensor(self, index: int, batch: int) -> torch.Tensor:
        dims, ch = self.program.buffers[index]
        off = self.lib.fnnu_engine_stats(self.handle, index) - self.workspace.data_ptr()
        return self.workspace[off:off + 16 * batch * ch].view(torch.float64).view(batch, ch, 2)

    def set_backend(self, backend: int):
        _lib.check(self.lib.fnnu_engine_set_backend(self.handle, int(backend)))

    def forward(self, batch: int):
        _lib.check(self.lib.fnnu_engine_forward(self.handle, int(batch), _lib.stream_ptr()))

    def profile_op(self, index: int):
        _lib.check(self.lib.fnnu_engine_profile_op(self.handle, int(index)))

    def profile_ms(self) -> float:
        ms = C.c_float(0)
        _lib.check(self.lib.fnnu_engine_profile_ms(self.handle, C.byref(ms)))
        return ms.value

    def launch_counts(self):
        t, u = C.c_int(0), C.c_int(0)
        _lib.check(self.lib.fnnu_engine_launch_counts(self.handle, C.byref(t), C.byref(u)))
        return t.value, u.value


# ---- sliding-window operators ------------------------------------------------------------------

def gather_tiles(volume: torch.Tensor, starts_dev: torch.Tensor, n_tiles: int, patch: Sequence[int],
                 flip_masks: bytes, out_ptr: int, c_stride: int):
    lib = _lib.load()
    assert volume.dtype == torch.float32 and volume.is_contiguous() and volume.ndim == 4
    assert starts_dev.dtype == torch.int32 and starts_dev.is_contiguous()
    _lib.check(lib.fnnu_gather_tiles(_ptr(volume), volume.shape[0], _lib.i3(volume.shape[1:]), _ptr(starts_dev),
                                     n_tiles, _lib.i3(patch), flip_masks, len(flip_masks), C.c_void_p(out_ptr),
                                     c_stride, _lib.stream_ptr()))


def accumulate_tiles(preds_ptr: int, in_dtype: int, p_stride: int, heads: int, starts_host: np.ndarray,
                     patch: Sequence[int], flip_masks: bytes, gaussian: Optional[torch.Tensor], acc: torch.Tensor):
    lib = _lib.load()
    starts_host = np.ascontiguousarray(starts_host, dtype=np.int32)
    acc_dtype = _lib.ACC_F32 if acc.dtype == torch.float32 else _lib.ACC_F16
    assert acc.is_contiguous() and acc.shape[0] == heads
    if gaussian is not None:
        assert gaussian.dtype == torch.float16 and gaussian.is_contiguous()
    _lib.check(lib.fnnu_accumulate_tiles(C.c_void_p(preds_ptr), in_dtype, p_stride, heads,
                                         starts_host.ctypes.data_as(C.POINTER(C.c_int32)), len(starts_host),
                                         _lib.i3(patch), flip_masks, len(flip_masks), _ptr(gaussian), _ptr(acc),
                                         acc_dtype, _lib.i3(acc.shape[1:]), _lib.stream_ptr()))


def weight_sum(steps, patch: Sequence[int], gaussian: Optional[torch.Tensor], wsum: torch.Tensor):
    lib = _lib.load()
    arrs = [np.ascontiguousarray(s, dtype=np.int32) for s in steps]
    acc_dtype = _lib.ACC_F32 if wsum.dtype == torch.float32 else _lib.ACC_F16
    p = [a.ctypes.data_as(C.POINTER(C.c_int32)) for a in arrs]
    _lib.check(lib.fnnu_weight_sum(p[0], len(arrs[0]), p[1], len(arrs[1]), p[2], len(arrs[2]), _lib.i3(patch),
                                   _ptr(gaussian), _ptr(wsum), acc_dtype, _lib.i3(wsum.shape), _lib.stream_ptr()))


def finalize(acc: torch.Tensor, wsum: torch.Tensor, logits_out: Optional[torch.Tensor],
             labels_out: Optional[torch.Tensor], inf_flag: torch.Tensor):
    """acc may be a slab view [H, x0:x1] of a larger accumulator (planes contiguous, heads strided)."""
    lib = _lib.load()
    acc_dtype = _lib.ACC_F32 if acc.dtype == torch.float32 else _lib.ACC_F16
    assert wsum.dtype == acc.dtype and inf_flag.dtype == torch.int32 and wsum.is_contiguous()
    assert acc[0].is_contiguous() and tuple(wsum.shape) == tuple(acc.shape[1:])
    head_stride = acc.stride(0) if acc.shape[0] > 1 else 0
    if logits_out is not None:
        assert logits_out.dtype == torch.float16 and logits_out.is_contiguous() and logits_out.shape == acc.shape
    if labels_out is not None:
        assert labels_out.dtype == torch.uint8 and labels_out.is_contiguous()
    _lib.check(lib.fnnu_finalize(_ptr(acc), _ptr(wsum), acc_dtype, acc.shape[0], _lib.i3(acc.shape[1:]),
                                 head_stride, _ptr(logits_out), _ptr(labels_out), _ptr(inf_flag),
                                 _lib.stream_ptr()))


def scale_inplace(x: torch.Tensor, factor: float):
    lib = _lib.load()
    assert x.dtype == torch.float32 and x.is_contiguous()
    _lib.check(lib.fnnu_scale_inplace_f32(_ptr(x), C.c_float(factor), x.numel(), _lib.stream_ptr()))


def mem_launches() -> int:
    return int(_lib.load().fnnu_mem_launches())


def add_inplace(acc: torch.Tensor, other: torch.Tensor):
    lib = _lib.load()
    assert acc.dtype == torch.float32 and other.dtype == torch.float32 and acc.numel() == other.numel()
    assert acc.is_contiguous() and other.is_contiguous()
    if acc.numel() == 0:
        return
    _lib.check(lib.fnnu_add_inplace_f32(_ptr(acc), _ptr(other), acc.numel(), _lib.stream_ptr()))
