"""Run under torch.distributed.run with N ranks (one GPU each): the sharded prediction of one volume must equal the
single-GPU prediction up to fp32 summation order.  Used by tests/test_gpu_sharded.py and by hand:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/sharded_check.py
"""
import os
import sys
import tempfile

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    import nets
    from fast_nnunet_b200 import model_folder as M
    from fast_nnunet_b200 import nnUNetPredictor
    rank = int(os.environ['RANK'])
    world = int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    if os.environ.get('SHARDED_CHECK', 'student') == 'toy':
        spec, vol = nets.SMALL_PLAIN16, (112, 40, 48)
    else:
        # cfg-2-shaped slabs: the distilled student, 128^3 patches at pitch 64 (3 x 2 x 2 tiles x 8 mirror passes)
        spec, vol = nets.STUDENT, (256, 192, 192)
    sd, _ = nets.make(spec)
    x = nets.ct_like_volume(vol, 1)         # a HOST tensor: every rank uploads only the planes its tiles read
    with tempfile.TemporaryDirectory() as tmp:
        folder = M.write_model_folder(os.path.join(tmp, f'r{rank}', 'nnUNetTrainer__nnUNetPlans__3d_fullres'), spec['cls'],
                                      spec['kw'], spec['patch'], sd, 1, 2)
        p = nnUNetPredictor(device=dev, allow_tqdm=False)
        p.initialize_from_trained_model_folder(folder, use_folds=(0,))
    single = p.predict_sliding_window_return_logits(x)
    single_lab = p.predict_sliding_window_return_segmentation(x)
    sharded = p.predict_sliding_window_sharded(x, return_labels=False, gather_to=0)
    labels = p.predict_sliding_window_sharded(x, return_labels=True, gather_to=0)
    slab, (o0, o1) = p.predict_sliding_window_sharded(x, return_labels=False, gather_to=None)
    ok = torch.tensor([1], device=dev)
    d_slab = (slab.float() - single[:, o0:o1].float()).abs().max().item() if o1 > o0 else 0.0
    if d_slab > 2e-2:
        ok[0] = 0
    if rank == 0:
        d = (sharded.float() - single.float()).abs().max().item()
        agree = (labels == single_lab).float().mean().item()
        print(f'world={world}: max|sharded - single| = {d:.5f}, label agreement = {agree:.6f}', flush=True)
        if d > 2e-2 or agree < 0.999 or tuple(sharded.shape) != tuple(single.shape):
            ok[0] = 0
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if int(ok.item()) != 1:
        sys.exit(1)
    if rank == 0:
        print('SHARDED OK', flush=True)


if __name__ == '__main__':
    main()
