"""Golden logits produced by RUNNING the reference's own nnUNetPredictor (predict_from_raw_data.py) on the CPU.

Run in the build container only (needs /root/reference):
    python tests/golden/make_predictor_golden.py

`predict_from_raw_data.py` and `sliding_window_prediction.py` are imported by path.  The modules they import at the
top but that are absent from this image (acvl_utils, batchgenerators, dynamic_network_architectures users ...) are
replaced by name-only stubs; none of them is reached by the methods exercised here, with ONE exception:
`acvl_utils.cropping_and_padding.padding.pad_nd_image`, which is stubbed by a restatement of the published function.
For every case whose volume is at least as large as the patch the padding is the identity, so those goldens pin
    _internal_get_sliding_window_slicers, _internal_maybe_mirror_and_predict,
    _internal_predict_sliding_window_return_logits, predict_sliding_window_return_logits,
    predict_logits_from_preprocessed_data (fold loop)
to the reference's executed code; the `padded` case additionally depends on the pad_nd_image stub.
The network is the oracle's torch module (oracle/networks.py) with seeded weights — the reference treats the network
as an opaque nn.Module, so this pins the LOOP (rows a2, a5, a6, a7, a10 of SURVEY.md section 8), not the network.

The `model_folder` case goes through the reference's initialize_from_trained_model_folder on a folder written by
fast_nnunet_b200.model_folder.write_model_folder (the fixture every GPU predictor test loads) with the reference's REAL
plans_handler.py and label_handling.py: it shows that the reference reads that folder, records what it read, and pins
the two-fold ensemble through the file-based entry.  The trainer class it looks up is a stand-in whose
build_network_architecture returns the oracle network.

Writes tests/golden/predictor_golden.npz + .json (the inputs are regenerated from seeds by the test).
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

REF = '/root/reference/distillation/nnunetv2'
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

CASES = [
    # name, volume (c, x, y, z), patch, step, use_gaussian, use_mirroring, mirror axes, folds
    dict(name='gauss_mirror', volume=(1, 26, 24, 20), patch=(16, 16, 16), step=0.5, gaussian=True, mirroring=True, axes=(0, 1, 2), folds=1),
    dict(name='nogauss_mirror', volume=(1, 26, 24, 20), patch=(16, 16, 16), step=0.5, gaussian=False, mirroring=True, axes=(0, 1, 2), folds=1),
    dict(name='gauss_nomirror', volume=(1, 26, 24, 20), patch=(16, 16, 16), step=0.5, gaussian=True, mirroring=False, axes=(0, 1, 2), folds=1),
    dict(name='axes02_step075', volume=(1, 30, 16, 23), patch=(16, 16, 16), step=0.75, gaussian=True, mirroring=True, axes=(0, 2), folds=1),
    dict(name='two_folds', volume=(1, 20, 24, 16), patch=(16, 16, 16), step=0.5, gaussian=True, mirroring=True, axes=(0, 1, 2), folds=2),
    dict(name='padded', volume=(1, 12, 21, 16), patch=(16, 16, 16), step=0.5, gaussian=True, mirroring=True, axes=(0, 1, 2), folds=1),
]
NET = 'TINY_ONNX'          # tests/nets.py: 3 heads, anisotropic first stage, patch 16^3


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _module(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def pad_nd_image_stub(image, new_shape=None, mode='constant', kwargs=None, return_slicer=False, shape_must_be_divisible_by=None):
    """acvl_utils.cropping_and_padding.padding.pad_nd_image restated from the published package for the call
    predict_from_raw_data.py:657-659 makes (torch tensor, mode 'constant', value 0, trailing axes, centred)."""
    assert mode == 'constant' and shape_must_be_divisible_by is None
    kwargs = kwargs or {}
    old_shape = np.array(image.shape)
    n = len(new_shape)
    new_shape = np.array(list(old_shape[:len(old_shape) - n]) + list(new_shape))
    new_shape = np.maximum(new_shape, old_shape)
    difference = new_shape - old_shape
    pad_below = difference // 2
    pad_above = difference // 2 + difference % 2
    pad_list = [[int(a), int(b)] for a, b in zip(pad_below, pad_above)]
    if any(a or b for a, b in pad_list):
        torch_pad = [i for j in pad_list[::-1] for i in j]
        res = torch.nn.functional.pad(image, torch_pad, mode, value=kwargs.get('value', 0))
    else:
        res = image
    if not return_slicer:
        return res
    slicer = tuple(slice(a, int(s) - b) for (a, b), s in zip(pad_list, res.shape))
    return res, slicer


def install_stubs():
    nothing = lambda *a, **k: None      # noqa: E731
    _module('acvl_utils')
    _module('acvl_utils.cropping_and_padding')
    _module('acvl_utils.cropping_and_padding.padding', pad_nd_image=pad_nd_image_stub)
    _module('batchgenerators')
    _module('batchgenerators.dataloading')
    _module('batchgenerators.dataloading.multi_threaded_augmenter', MultiThreadedAugmenter=object)
    _module('batchgenerators.utilities')
    _module('batchgenerators.utilities.file_and_folder_operations', load_json=nothing, join=os.path.join,
            isfile=os.path.isfile, maybe_mkdir_p=nothing, isdir=os.path.isdir, subdirs=nothing, save_json=nothing)
    pkg = _module('nnunetv2')
    pkg.__path__ = ['/nonexistent']
    _module('nnunetv2.configuration', default_num_processes=8)
    _module('nnunetv2.inference')
    _module('nnunetv2.inference.data_iterators', PreprocessAdapterFromNpy=object, preprocessing_iterator_fromfiles=nothing,
            preprocessing_iterator_fromnpy=nothing)
    _module('nnunetv2.inference.export_prediction', export_prediction_from_logits=nothing,
            convert_predicted_logits_to_segmentation_with_correct_shape=nothing)
    _load(os.path.join(REF, 'inference/sliding_window_prediction.py'), 'nnunetv2.inference.sliding_window_prediction')
    _module('nnunetv2.utilities')
    _module('nnunetv2.utilities.file_path_utilities', get_output_folder=nothing, check_workers_alive_and_busy=nothing)
    _module('nnunetv2.utilities.find_class_by_name', recursive_find_python_class=nothing)
    _load(os.path.join(REF, 'utilities/helpers.py'), 'nnunetv2.utilities.helpers')
    _module('nnunetv2.utilities.json_export', recursive_fix_for_json_export=nothing)
    # the REAL label handling and plans handling (their own absent imports stubbed)
    import json as _json

    def load_json(path):
        with open(path) as f:
            return _json.load(f)

    sys.modules['batchgenerators.utilities.file_and_folder_operations'].load_json = load_json
    _module('acvl_utils.cropping_and_padding.bounding_boxes', bounding_box_to_slice=nothing, insert_crop_into_image=nothing)
    _module('nnunetv2.preprocessing')
    _module('nnunetv2.preprocessing.resampling')
    _module('nnunetv2.preprocessing.resampling.utils', recursive_find_resampling_fn_by_name=lambda name: (lambda *a, **k: None))
    _module('nnunetv2.imageio')
    _module('nnunetv2.imageio.reader_writer_registry', recursive_find_reader_writer_by_name=nothing)
    _module('dynamic_network_architectures')
    _module('dynamic_network_architectures.building_blocks')
    _module('dynamic_network_architectures.building_blocks.helper', convert_dim_to_conv_op=lambda dim: torch.nn.Conv3d,
            get_matching_instancenorm=lambda conv_op=None, dimension=None: torch.nn.InstanceNorm3d)
    _module('nnunetv2.utilities.label_handling')
    lab = _load(os.path.join(REF, 'utilities/label_handling/label_handling.py'), 'nnunetv2.utilities.label_handling.label_handling')
    _module('nnunetv2.utilities.plans_handling')
    _load(os.path.join(REF, 'utilities/plans_handling/plans_handler.py'), 'nnunetv2.utilities.plans_handling.plans_handler')

    class StandInTrainer:
        @staticmethod
        def build_network_architecture(architecture_class_name, arch_init_kwargs, arch_init_kwargs_req_import,
                                       num_input_channels, num_output_channels, enable_deep_supervision=True):
            from oracle import networks as N
            return N.build_from_arch(architecture_class_name, arch_init_kwargs, num_input_channels, num_output_channels,
                                     allow_init=False)

    def find_class(folder, class_name, current_module):
        return {'LabelManager': lab.LabelManager, 'nnUNetTrainer': StandInTrainer}.get(class_name)

    sys.modules['nnunetv2.utilities.find_class_by_name'].recursive_find_python_class = find_class
    lab.recursive_find_python_class = find_class
    sys.modules['nnunetv2.utilities.plans_handling.plans_handler'].recursive_find_python_class = find_class
    _module('nnunetv2.utilities.utils', create_lists_from_splitted_dataset_folder=nothing)


class _Labels:
    def __init__(self, heads):
        self.num_segmentation_heads = heads


class _Plans:
    def __init__(self, heads):
        self._heads = heads

    def get_label_manager(self, dataset_json):
        return _Labels(self._heads)


class _Config:
    def __init__(self, patch):
        self.patch_size = list(patch)


def main():
    install_stubs()
    ref = _load(os.path.join(REF, 'inference/predict_from_raw_data.py'), 'ref_predict_from_raw_data')
    import nets
    from fast_nnunet_b200 import model_folder as M
    from oracle import networks as N

    spec = getattr(nets, NET)
    torch.set_num_threads(8)
    arrays, meta = {}, {'net': NET, 'cases': []}
    for case in CASES:
        assert tuple(case['patch']) == tuple(spec['patch'])
        sds = [M.synthesize_state_dict(spec['cls'], spec['kw'], spec['in_ch'], spec['heads'], seed=4321 + f, randomize_affine=True)
               for f in range(case['folds'])]
        net = N.build_from_arch(spec['cls'], spec['kw'], spec['in_ch'], spec['heads'], allow_init=False)
        net.load_state_dict(sds[0], strict=True)
        p = ref.nnUNetPredictor(tile_step_size=case['step'], use_gaussian=case['gaussian'], use_mirroring=case['mirroring'],
                                perform_everything_on_device=False, device=torch.device('cpu'), verbose=False,
                                verbose_preprocessing=False, allow_tqdm=False)
        p.manual_initialization(net, _Plans(spec['heads']), _Config(case['patch']), sds, {}, 'nnUNetTrainer', tuple(case['axes']))
        x = nets.ct_like_volume(tuple(case['volume'][1:]), case['volume'][0], seed=11)
        slicers = p._internal_get_sliding_window_slicers(tuple(max(v, q) for v, q in zip(case['volume'][1:], case['patch'])))
        if case['folds'] == 1:
            out = p.predict_sliding_window_return_logits(x)
        else:
            out = p.predict_logits_from_preprocessed_data(x)
        assert out.dtype == torch.float16 and tuple(out.shape) == (spec['heads'], *case['volume'][1:])
        arrays[case['name']] = out.numpy()
        m = dict(case)
        m['n_tiles'] = len(slicers)
        m['first_slicers'] = [[[s.start, s.stop] for s in sl[1:]] for sl in slicers[:4]]
        meta['cases'].append(m)
        print(case['name'], tuple(out.shape), 'tiles', len(slicers), 'range', float(out.float().min()), float(out.float().max()))
    # ---- file-based entry: initialize_from_trained_model_folder on a folder written by write_model_folder ----------
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        folder = os.path.join(tmp, 'nnUNetTrainer__nnUNetPlans__3d_fullres')
        for f in range(2):
            sd = M.synthesize_state_dict(spec['cls'], spec['kw'], spec['in_ch'], spec['heads'], seed=777 + f, randomize_affine=True)
            M.write_model_folder(folder, spec['cls'], spec['kw'], spec['patch'], sd, spec['in_ch'], spec['heads'], fold=f,
                                 mirror_axes=(0, 1, 2))
        p = ref.nnUNetPredictor(tile_step_size=0.5, use_gaussian=True, use_mirroring=True, perform_everything_on_device=False,
                                device=torch.device('cpu'), verbose=False, verbose_preprocessing=False, allow_tqdm=False)
        p.initialize_from_trained_model_folder(folder, use_folds=(0, 1), checkpoint_name='checkpoint_final.pth')
        x = nets.ct_like_volume((20, 24, 16), spec['in_ch'], seed=12)
        out = p.predict_logits_from_preprocessed_data(x)
        arrays['model_folder'] = out.numpy()
        meta['model_folder'] = {
            'volume': [spec['in_ch'], 20, 24, 16], 'folds': [0, 1], 'seeds': [777, 778],
            'trainer_name': p.trainer_name, 'allowed_mirroring_axes': list(p.allowed_mirroring_axes),
            'patch_size': list(p.configuration_manager.patch_size), 'n_parameter_sets': len(p.list_of_parameters),
            'num_segmentation_heads': int(p.label_manager.num_segmentation_heads),
            'all_labels': [int(i) for i in p.label_manager.all_labels],
            'network_arch_class_name': p.configuration_manager.network_arch_class_name}
        print('model_folder', tuple(out.shape), meta['model_folder']['trainer_name'], meta['model_folder']['patch_size'])

    np.savez_compressed(os.path.join(HERE, 'predictor_golden.npz'), **arrays)
    with open(os.path.join(HERE, 'predictor_golden.json'), 'w') as f:
        json.dump(meta, f, indent=1)
    print('wrote predictor_golden.npz', os.path.getsize(os.path.join(HERE, 'predictor_golden.npz')) // 1024, 'KB')


if __name__ == '__main__':
    main()
