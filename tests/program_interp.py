"""Test infrastructure: a plain-PyTorch (CPU, fp32) interpreter of the flat buffer / operator program that
fast_nnunet_b200/program.py lowers a network into and libfnnu executes (include/fnnu.h).

It restates the program's SEMANTICS, not the kernels: every conv writes its RAW output; a buffer channel written by an
operator with a norm carries a pending transform lrelu(gamma * (x - mean) * rstd + beta, slope) with the InstanceNorm
statistics of that (sample, channel); whoever reads the channel applies it first.  Running a lowered program here and
comparing with the oracle network checks the lowering itself (key names, wiring, concat offsets, strides, which seg
layer is evaluated, dropped biases) without a GPU."""
import torch
import torch.nn.functional as F

from fast_nnunet_b200 import _lib


def run_program(prog, x: torch.Tensor) -> torch.Tensor:
    """x: (n, C, d0, d1, d2) fp32 -> logits (n, heads, d0, d1, d2) fp32."""
    n = x.shape[0]
    bufs = [torch.zeros((n, c, *dims), dtype=torch.float32) for dims, c in prog.buffers]
    # pending transform per buffer channel: None (ready to use) or (gamma, beta, eps, slope)
    pending = [[None] * c for _, c in prog.buffers]
    bufs[prog.input_buffer][:] = x

    def load(b, coff, c):
        raw = bufs[b][:, coff:coff + c]
        out = raw.clone()
        for j in range(c):
            t = pending[b][coff + j]
            if t is None:
                continue
            gamma, beta, eps, slope = t
            v = raw[:, j]
            mean = v.mean(dim=(1, 2, 3), keepdim=True)
            var = v.var(dim=(1, 2, 3), unbiased=False, keepdim=True)
            y = (v - mean) / torch.sqrt(var + eps) * gamma + beta
            out[:, j] = torch.where(y > 0, y, y * slope)
        return out

    for op in prog.ops:
        src = load(op.src, op.src_coff, op.cin)
        if op.op == _lib.OP_CONV:
            w = torch.from_numpy(op.weight)
            b = torch.from_numpy(op.bias) if op.bias is not None else None
            out = F.conv3d(src, w, b, stride=tuple(op.stride), padding=tuple((k - 1) // 2 for k in op.kernel))
        elif op.op == _lib.OP_TCONV:
            w = torch.from_numpy(op.weight)
            b = torch.from_numpy(op.bias) if op.bias is not None else None
            out = F.conv_transpose3d(src, w, b, stride=tuple(op.stride))
        elif op.op == _lib.OP_ADD_ACT:
            s = src + load(op.src2, op.src2_coff, op.cin)
            out = torch.where(s > 0, s, s * op.act_slope)
        elif op.op == _lib.OP_AVGPOOL:
            out = F.avg_pool3d(src, kernel_size=tuple(op.stride), stride=tuple(op.stride))
        else:
            raise ValueError(op.op)
        assert tuple(out.shape[2:]) == tuple(prog.buffers[op.dst][0]), (op.name, out.shape, prog.buffers[op.dst])
        bufs[op.dst][:, op.dst_coff:op.dst_coff + op.cout] = out
        for j in range(op.cout):
            pending[op.dst][op.dst_coff + j] = None if not op.has_norm else \
                (float(op.gamma[j]), float(op.beta[j]), float(op.eps), float(op.act_slope))
    assert all(t is None for t in pending[prog.output_buffer][:prog.num_heads])
    return bufs[prog.output_buffer][:, :prog.num_heads]
