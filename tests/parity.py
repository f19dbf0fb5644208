"""Parity helpers shared by the GPU tests: the oracle side of a sliding-window run and north_star's acceptance
criteria (tolerances are written here, once)."""
import numpy as np
import torch

from oracle import predictor as OP

DEV = torch.device('cuda', 0)

def oracle_volume(net, x, patch, use_gaussian=True, mirror_axes=(0, 1, 2), on_gpu=False):
    """Reference arithmetic (fp16 accumulators) and the exact-accumulation variant from ONE set of oracle
    tile predictions.  The fp32 oracle network runs on the GPU (TF32 off, tests/conftest.py) for speed; with
    on_gpu=True the accumulation does too (large volumes)."""
    net = net.to(DEV)
    _, tile_preds, slicers = OP.predict_sliding_window_return_logits(net, x.to(DEV), patch, 0.5, use_gaussian,
                                                                     mirror_axes, return_tile_predictions=True)
    net.cpu()
    if not on_gpu:
        tile_preds = [t.cpu() for t in tile_preds]
    xp, revert = OP.pad_to_patch(x, patch)
    heads = tile_preds[0].shape[0]
    ref16, n16 = OP.accumulate_tiles(tile_preds, slicers, tuple(xp.shape[1:]), heads, patch, use_gaussian, torch.half)
    ref32, _ = OP.accumulate_tiles(tile_preds, slicers, tuple(xp.shape[1:]), heads, patch, use_gaussian, torch.float32)
    crop = (slice(None), *revert[1:])
    return ref16[crop].cpu(), n16[tuple(revert[1:])].cpu(), ref32[crop].cpu()


LABEL_BAR = 0.999      # north_star: label-map voxel agreement >= 99.9 % ...
DICE_BAR = 0.999       # ... and per-class Dice >= 0.999 against the reference


def compare(got_logits, oracle, heads, tol_max=0.06, tol_mean=0.006):
    """Parity statement, fixtures = TRAINED oracle weights on a phantom (tests/nets.py):
      (1) against the oracle with exact (fp32) accumulation: max / mean |d| of the normalised logits on ALL voxels,
          tolerances scaled with the logit range (trained logits span +-10..20, the He-init ones +-1);
      (2) against the reference arithmetic (fp16 accumulators): the same on the voxels whose weight sum is a normal
          fp16 number (n_predictions >= 6.1e-5; below that the reference's own accumulators hold 1-2 significant bits);
      (3) north_star's label bar on ALL voxels: agreement >= 99.9 % and Dice >= 0.999 for every class that occurs."""
    ref16, n16, ref32 = oracle
    got = got_logits.float().cpu()
    scale = max(1.0, float(ref32.abs().max()) / 8.0)
    d32 = (got - ref32.float()).abs()
    ok16 = (n16.float() >= 6.1e-5)
    d16 = (got - ref16.float()).abs()[:, ok16]
    seg_g, seg_32, seg_16 = (OP.logits_to_segmentation(t) for t in (got, ref32, ref16))
    agree32 = float((seg_g == seg_32).mean())
    agree16 = float((seg_g == seg_16)[ok16.numpy()].mean())
    dice = [d for d in OP.dice_per_class(seg_g, seg_32, heads) if d == d]      # classes absent from both maps: nan
    print(f'vs exact-acc oracle: max|d|={d32.max():.4f} mean|d|={d32.mean():.5f} (logit range {float(ref32.abs().max()):.1f}) '
          f'labels agree={agree32:.6f} min dice={min(dice):.5f} over {len(dice)} classes | vs fp16-acc reference '
          f'arithmetic on {float(ok16.float().mean()):.4f} of voxels: max|d|={d16.max():.4f} mean|d|={d16.mean():.5f} '
          f'labels agree={agree16:.6f}')
    assert d32.max().item() <= tol_max * scale and d32.mean().item() <= tol_mean * scale
    assert d16.max().item() <= (tol_max + 0.05) * scale and d16.mean().item() <= (tol_mean + 0.002) * scale
    assert agree32 >= LABEL_BAR, f'label agreement {agree32:.6f} < {LABEL_BAR}'
    assert min(dice) >= DICE_BAR, f'per-class Dice {dice} < {DICE_BAR}'


