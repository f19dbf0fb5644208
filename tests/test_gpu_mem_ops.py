"""GPU parity of the memory-bound sliding-window operators against the oracle (bit-exact where the
arithmetic is specified: gather, fp16-accumulator emulation, weight sum, argmax)."""
import numpy as np
import pytest
import torch

from fast_nnunet_b200 import _lib, engine as E
from fast_nnunet_b200 import sliding_window as sw
from oracle import predictor as OP
from oracle import sliding_window as osw

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda', 0)
FLIPS8 = bytes([0, 1, 2, 4, 3, 5, 6, 7])
AXES8 = [(), (0,), (1,), (2,), (0, 1), (0, 2), (1, 2), (0, 1, 2)]


@pytest.mark.parametrize('C,vol,patch,cs', [(1, (40, 48, 36), (32, 32, 32), 1), (4, (33, 35, 40), (16, 24, 20), 4),
                                            (2, (20, 20, 24), (20, 20, 24), 8)])
def test_gather_tiles_bit_exact(C, vol, patch, cs):
    g = torch.Generator().manual_seed(0)
    v = (torch.randn((C, *vol), generator=g) * 300).to(DEV)
    starts = sw.tile_starts(vol, patch, 0.5)
    sd = torch.from_numpy(starts).to(DEV)
    n = len(starts)
    out = torch.full((n * 8, *patch, cs), 7.0, dtype=torch.float16, device=DEV)
    E.gather_tiles(v, sd, n, patch, FLIPS8, out.data_ptr(), cs)
    torch.cuda.synchronize()
    for t, s in enumerate(starts):
        tile = v[:, s[0]:s[0] + patch[0], s[1]:s[1] + patch[1], s[2]:s[2] + patch[2]]
        for f, axes in enumerate(AXES8):
            want = torch.flip(tile, [a + 1 for a in axes]) if axes else tile
            want = want.permute(1, 2, 3, 0).half()
            got = out[t * 8 + f]
            assert torch.equal(got[..., :C], want), (t, f)
            if cs > C:
                assert torch.count_nonzero(got[..., C:]) == 0


def _tile_preds(n, heads, patch, seed, dtype):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn((heads, *patch), generator=g) * 3).to(dtype) for _ in range(n)]


@pytest.mark.parametrize('vol,patch,heads', [((40, 48, 36), (32, 32, 32), 2), ((33, 35, 41), (16, 24, 20), 3),
                                             ((32, 32, 32), (32, 32, 32), 5)])
def test_accumulate_fp16_emulation_bit_exact(vol, patch, heads):
    """fp32 tile predictions + fp16 accumulators == the reference's CPU arithmetic (PRED:611-620)."""
    sl = osw.slicers_for(vol, patch, 0.5)
    starts = sw.tile_starts(vol, patch, 0.5)
    preds = _tile_preds(len(sl), heads, patch, 1, torch.float32)
    want_logits, want_n = OP.accumulate_tiles(preds, sl, vol, heads, patch, True)
    g16 = torch.from_numpy(sw.compute_gaussian(patch, 1. / 8, 10, np.float16)).to(DEV)
    dev_preds = torch.stack([p.permute(1, 2, 3, 0) for p in preds]).contiguous().to(DEV)     # [n][x][y][z][h]
    acc = torch.zeros((heads, *vol), dtype=torch.float16, device=DEV)
    E.accumulate_tiles(dev_preds.data_ptr(), _lib.IN_F32, heads, heads, starts, patch, bytes([0]), g16, acc)
    wsum = torch.empty(vol, dtype=torch.float16, device=DEV)
    E.weight_sum(sw.compute_steps_for_sliding_window(vol, patch, 0.5), patch, g16, wsum)
    assert torch.equal(wsum.cpu().view(torch.int16), want_n.view(torch.int16))
    logits = torch.empty_like(acc)
    labels = torch.empty(vol, dtype=torch.uint8, device=DEV)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    E.finalize(acc, wsum, logits, labels, flag)
    torch.cuda.synchronize()
    assert int(flag.item()) == 0
    assert torch.equal(logits.cpu().view(torch.int16), want_logits.view(torch.int16))
    assert np.array_equal(labels.cpu().numpy(), OP.logits_to_segmentation(want_logits).astype(np.uint8))


@pytest.mark.parametrize('vol,patch,heads,pstride', [((64, 48, 64), (32, 32, 32), 2, 2), ((40, 48, 36), (32, 32, 32), 2, 2),
                                                     ((33, 35, 41), (16, 24, 20), 3, 8), ((48, 32, 32), (32, 32, 32), 61, 64),
                                                     # cluster kernel variants: 4 / 8 / 16 heads per load, 2 or 4 z voxels
                                                     ((48, 40, 48), (32, 32, 32), 4, 4), ((40, 40, 40), (32, 24, 32), 7, 8),
                                                     ((40, 48, 38), (32, 32, 32), 25, 32), ((36, 36, 44), (24, 24, 24), 3, 4),
                                                     # odd tile starts along z (cfg 5's tiles start at 46, 139, 185 ...): voxel-by-voxel loads
                                                     ((36, 40, 50), (24, 24, 32), 25, 32), ((30, 36, 50), (24, 24, 32), 61, 64),
                                                     ((30, 36, 50), (24, 24, 32), 7, 8)])
@pytest.mark.parametrize('chunk', [0, 4, 7])
def test_accumulate_tta_fp32(vol, patch, heads, pstride, chunk):
    """fp16 predictions of 8 mirrored passes -> un-flip, mean, Gaussian weight, fp32 accumulate, normalise."""
    starts = sw.tile_starts(vol, patch, 0.5)
    n = len(starts)
    g = torch.Generator().manual_seed(3)
    raw = (torch.randn((n * 8, *patch, pstride), generator=g) * 2).half().to(DEV)
    g16 = torch.from_numpy(sw.compute_gaussian(patch, 1. / 8, 10, np.float16)).to(DEV)
    acc = torch.zeros((heads, *vol), dtype=torch.float32, device=DEV)
    # chunk = 0: all tiles in one call; otherwise batches of `chunk` tiles as the predictor issues them (batches of
    # <= 8 tiles with 2 heads take the whole-batch kernel, everything else rounds of non-overlapping tiles)
    step = chunk if chunk else n
    per_tile = raw.shape[1] * raw.shape[2] * raw.shape[3] * pstride * 8 * 2
    for i in range(0, n, step):
        E.accumulate_tiles(raw.data_ptr() + i * per_tile, _lib.IN_F16, pstride, heads, starts[i:i + step], patch, FLIPS8,
                           g16, acc)
    wsum = torch.empty(vol, dtype=torch.float32, device=DEV)
    E.weight_sum(sw.compute_steps_for_sliding_window(vol, patch, 0.5), patch, g16, wsum)
    # reference arithmetic in fp32 on the GPU with torch ops
    ref = torch.zeros_like(acc)
    refw = torch.zeros_like(wsum)
    gf = g16.float()
    for t, s in enumerate(starts):
        p = None
        for f, axes in enumerate(AXES8):
            x = raw[t * 8 + f][..., :heads].permute(3, 0, 1, 2).float()
            x = torch.flip(x, [a + 1 for a in axes]) if axes else x
            p = x.clone() if p is None else p + x
        p = p / 8 * gf
        ref[:, s[0]:s[0] + patch[0], s[1]:s[1] + patch[1], s[2]:s[2] + patch[2]] += p
        refw[s[0]:s[0] + patch[0], s[1]:s[1] + patch[1], s[2]:s[2] + patch[2]] += gf
    torch.cuda.synchronize()
    assert torch.allclose(wsum, refw, rtol=1e-6, atol=0)
    assert torch.equal(acc, ref)          # same fp32 operation order -> identical
    logits = torch.empty((heads, *vol), dtype=torch.float16, device=DEV)
    labels = torch.empty(vol, dtype=torch.uint8, device=DEV)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    E.finalize(acc, wsum, logits, labels, flag)
    want = (ref / wsum).half()
    assert torch.equal(logits, want)
    assert torch.equal(labels.long(), torch.from_numpy(np.argmax(want.float().cpu().numpy(), 0)).to(DEV))


def test_no_gaussian_and_single_flip():
    vol, patch, heads = (24, 24, 24), (16, 16, 16), 2
    starts = sw.tile_starts(vol, patch, 0.5)
    g = torch.Generator().manual_seed(5)
    raw = torch.randn((len(starts), *patch, heads), generator=g).half().to(DEV)
    acc = torch.zeros((heads, *vol), dtype=torch.float32, device=DEV)
    E.accumulate_tiles(raw.data_ptr(), _lib.IN_F16, heads, heads, starts, patch, bytes([0]), None, acc)
    wsum = torch.empty(vol, dtype=torch.float32, device=DEV)
    E.weight_sum(sw.compute_steps_for_sliding_window(vol, patch, 0.5), patch, None, wsum)
    ref = torch.zeros_like(acc)
    refw = torch.zeros_like(wsum)
    for t, s in enumerate(starts):
        ref[:, s[0]:s[0] + 16, s[1]:s[1] + 16, s[2]:s[2] + 16] += raw[t].permute(3, 0, 1, 2).float()
        refw[s[0]:s[0] + 16, s[1]:s[1] + 16, s[2]:s[2] + 16] += 1
    assert torch.equal(acc, ref) and torch.equal(wsum, refw)


def test_inf_is_flagged():
    acc = torch.full((2, 8, 8, 8), 6.0e4, dtype=torch.float32, device=DEV)
    wsum = torch.full((8, 8, 8), 0.5, dtype=torch.float32, device=DEV)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    E.finalize(acc, wsum, torch.empty((2, 8, 8, 8), dtype=torch.float16, device=DEV), None, flag)
    assert int(flag.item()) == 1


def test_scale_inplace_and_launch_counter():
    a = torch.randn(100003, device=DEV)
    want = a * 3.0
    n0 = E.mem_launches()
    E.scale_inplace(a, 3.0)
    assert torch.equal(a, want)
    assert E.mem_launches() == n0 + 1


def test_add_inplace():
    a = torch.randn(1000003, device=DEV)
    b = torch.randn(1000003, device=DEV)
    want = a + b
    E.add_inplace(a, b)
    assert torch.equal(a, want)


def test_out_of_volume_tile_is_rejected():
    acc = torch.zeros((2, 16, 16, 16), dtype=torch.float32, device=DEV)
    raw = torch.zeros((1, 16, 16, 16, 2), dtype=torch.float16, device=DEV)
    with pytest.raises(_lib.FnnuError):
        E.accumulate_tiles(raw.data_ptr(), _lib.IN_F16, 2, 2, np.array([[4, 0, 0]], dtype=np.int32), (16, 16, 16),
                           bytes([0]), None, acc)
