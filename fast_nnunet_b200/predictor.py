"""nnUNetPredictor — drop-in for the reference's sliding-window inference path
(distillation/nnunetv2/inference/predict_from_raw_data.py:39-680) running on libfnnu (sm_100a).

Same constructor keywords, method names, argument meaning and error behaviour as the reference class
for: __init__ (:40-65), initialize_from_trained_model_folder (:67-129), manual_initialization
(:131-154), predict_sliding_window_return_logits (:634-680), predict_logits_from_preprocessed_data
(:471-504), predict_single_npy_array (:423-468).  What differs is underneath:

  * the volume stays resident on the device; tiles (and their mirrored copies) are cut by a gather
    kernel instead of a Python producer thread + queue (:568-582);
  * `self.network(x)` and the 8-pass mirror loop (:541-557) become ONE batched engine forward over
    tiles x flips;
  * `prediction *= gaussian; predicted_logits[sl] += prediction` (:611-613) is one fused kernel with fp32
    accumulators (fp16 accumulators, the reference's arithmetic, on request);
  * `n_predictions` (:614) is produced by one pass (it does not depend on the image);
  * there is no CPU fallback (:663-672): errors surface as exceptions.
"""
from __future__ import annotations

import itertools
import os
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib, engine as E
from . import sliding_window as sw
from .model_folder import effective_arch
from .plans import ConfigurationManager, PlansManager, determine_num_input_channels, load_json
from .program import Program, build_program, clean_state_dict


def _infer_arch_from_weights(cls_name: str, kw: dict, sd: dict) -> dict:
    """Makes features / conv / block counts agree with the checkpoint itself (so a distilled student is
    loaded correctly even when the plans describe the teacher and init_args are incomplete)."""
    kw = dict(kw)
    sd = clean_state_dict(sd)
    n = int(kw['n_stages'])
    resenc = cls_name.endswith('ResidualEncoderUNet')
    feats = []
    for s in range(n):
        key = f'encoder.stages.{s}.blocks.0.conv1.conv.weight' if resenc else f'encoder.stages.{s}.0.convs.0.conv.weight'
        if key not in sd:
            raise KeyError(f'checkpoint is missing parameter {key!r}')
        feats.append(int(sd[key].shape[0]))
    kw['features_per_stage'] = feats
    if resenc:
        kw['n_blocks_per_stage'] = [
            len({k.split('.')[4] for k in sd if k.startswith(f'encoder.stages.{s}.blocks.')}) for s in range(n)]
    else:
        kw['n_conv_per_stage'] = [
            len({k.split('.')[5] for k in sd if k.startswith(f'encoder.stages.{s}.0.convs.')}) for s in range(n)]
    kw['n_conv_per_stage_decoder'] = [
        len({k.split('.')[4] for k in sd if k.startswith(f'decoder.stages.{l}.convs.')}) for l in range(n - 1)]
    kw['conv_bias'] = (('encoder.stem.convs.0.conv.bias' if resenc else 'encoder.stages.0.0.convs.0.conv.bias') in sd)
    return kw


class CompiledNetwork:
    """Stands where the reference keeps `self.network` (an nn.Module).  `load_state_dict` selects a set of
    weights; the first use lowers it into a libfnnu engine, which then stays resident (one per fold, sharing one
    activation workspace), so that a multi-fold ensemble never re-packs or re-uploads parameters
    (predict_from_raw_data.py:483-500 re-loads the state_dict per fold per case)."""

    MAX_RESIDENT = 8        # parameter sets kept lowered at the same time (5 folds + head-room)

    def __init__(self, network_class_name: str, arch_kwargs: dict, in_channels: int, num_heads: int,
                 patch_size: Sequence[int]):
        self.network_class_name = network_class_name
        self.arch_kwargs = arch_kwargs
        self.in_channels = in_channels
        self.num_heads = num_heads
        self.patch_size = tuple(int(p) for p in patch_size)
        self.device = None
        # every entry holds a strong reference to its state_dict, so the identity test below can never be
        # fooled by a recycled id()
        self._entries = []          # [{'params': dict, 'engines': {device str: NetworkEngine}}], most recent last
        self._current = None
        self.program: Optional[Program] = None

    # nn.Module-looking no-ops so that reference-style calling code keeps working
    def to(self, device):
        self.device = torch.device(device)
        return self

    def eval(self):
        return self

    def _entry_for(self, params: dict) -> dict:
        for i, e in enumerate(self._entries):
            if e['params'] is params:
                self._entries.append(self._entries.pop(i))
                return e
        e = {'params': params, 'engines': {}}
        self._entries.append(e)
        while len(self._entries) > self.MAX_RESIDENT:
            self._entries.pop(0)
        return e

    def load_state_dict(self, params: dict, strict: bool = True):
        self._current = self._entry_for(params)
        return self

    def engine(self, device: torch.device, max_batch: int, params: Optional[dict] = None) -> E.NetworkEngine:
        """Engine of `params` (default: the set selected by load_state_dict) able to run `max_batch` patches."""
        entry = self._current if params is None else self._entry_for(params)
        if entry is None:
            raise RuntimeError('no parameters loaded (call load_state_dict / initialize_* first)')
        key = str(device)
        eng = entry['engines'].get(key)
        if eng is None or eng.max_batch < int(max_batch):
            kw = _infer_arch_from_weights(self.network_class_name, self.arch_kwargs, entry['params'])
            prog = build_program(self.network_class_name, entry['params'], kw, self.in_channels, self.num_heads,
                                 self.patch_size)
            entry['engines'].pop(key, None)
            share = None
            for other in self._entries:
                o = other['engines'].get(key)
                if o is not None and o.max_batch == int(max_batch):
                    share = o
                    break
            eng = E.NetworkEngine(prog, int(max_batch), device, share_workspace_with=share)
            entry['engines'][key] = eng
        self.program = eng.program
        return eng

    @torch.inference_mode()
    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        """x: (n, c, *patch) on the CUDA device -> logits (n, heads, *patch) fp16.  What reference-side callers such
        as nnUNetDistillationTrainer.load_teacher_model's `teacher_model(x.float())` (:781-789) get."""
        assert x.ndim == 5 and tuple(x.shape[2:]) == self.patch_size and x.shape[1] == self.in_channels
        eng = self.engine(x.device, max(int(x.shape[0]), 1))
        n = x.shape[0]
        inp = eng.buffer_tensor(eng.program.input_buffer, n)
        inp.copy_(x.permute(0, 2, 3, 4, 1))
        eng.forward(n)
        out = eng.buffer_tensor(eng.program.output_buffer, n)[..., :self.num_heads]
        return out.permute(0, 4, 1, 2, 3).contiguous()


class nnUNetPredictor(object):
    def __init__(self,
                 tile_step_size: float = 0.5,
                 use_gaussian: bool = True,
                 use_mirroring: bool = True,
                 perform_everything_on_device: bool = True,
                 device: torch.device = torch.device('cuda'),
                 verbose: bool = False,
                 verbose_preprocessing: bool = False,
                 allow_tqdm: bool = True,
                 *,
                 tiles_per_batch: Optional[int] = None,
                 accumulator_dtype: torch.dtype = torch.float32):
        self.verbose = verbose
        self.verbose_preprocessing = verbose_preprocessing
        self.allow_tqdm = allow_tqdm

        self.plans_manager, self.configuration_manager, self.list_of_parameters, self.network, self.dataset_json, \
            self.trainer_name, self.allowed_mirroring_axes, self.label_manager = (None,) * 8

        self.tile_step_size = tile_step_size
        self.use_gaussian = use_gaussian
        self.use_mirroring = use_mirroring
        device = torch.device(device)
        if device.type != 'cuda':
            print('perform_everything_on_device=True is only supported for cuda devices! Setting this to False')
            perform_everything_on_device = False
        self.device = device
        self.perform_everything_on_device = perform_everything_on_device
        assert accumulator_dtype in (torch.float32, torch.float16)
        self.accumulator_dtype = accumulator_dtype
        self.tiles_per_batch = tiles_per_batch
        self.last_launches = 0          # kernels launched by the last predict_* call (bench evidence)
        self.last_tiles_per_batch = None
        self.collect_timing = False     # bench: CUDA events around each phase of the tile loop
        self.timing = {}
        self._gauss_cache = {}
        self._wsum_cache = {}           # n_predictions maps: input independent, 4 bytes per voxel -> keep the last two
        self._pinned_cache = {}

    # ------------------------------------------------------------------ model loading
    def initialize_from_trained_model_folder(self, model_training_output_dir: str,
                                             use_folds: Union[Tuple[Union[int, str]], None],
                                             checkpoint_name: str = 'checkpoint_final.pth'):
        """This is used when making predictions with a trained model."""
        if use_folds is None:
            use_folds = nnUNetPredictor.auto_detect_available_folds(model_training_output_dir, checkpoint_name)

        dataset_json = load_json(os.path.join(model_training_output_dir, 'dataset.json'))
        plans = load_json(os.path.join(model_training_output_dir, 'plans.json'))
        plans_manager = PlansManager(plans)

        if isinstance(use_folds, (str, int)):
            use_folds = [use_folds]

        parameters = []
        trainer_name = configuration_name = inference_allowed_mirroring_axes = init_args = None
        for i, f in enumerate(use_folds):
            f = int(f) if f != 'all' else f
            checkpoint = torch.load(os.path.join(model_training_output_dir, f'fold_{f}', checkpoint_name),
                                    map_location=torch.device('cpu'), weights_only=False)
            if i == 0:
                trainer_name = checkpoint['trainer_name']
                init_args = checkpoint['init_args']
                configuration_name = init_args['configuration']
                inference_allowed_mirroring_axes = checkpoint['inference_allowed_mirroring_axes'] if \
                    'inference_allowed_mirroring_axes' in checkpoint.keys() else None
            parameters.append(checkpoint['network_weights'])

        configuration_manager = plans_manager.get_configuration(configuration_name)
        num_input_channels = determine_num_input_channels(plans_manager, configuration_manager, dataset_json)
        label_manager = plans_manager.get_label_manager(dataset_json)
        cls, kw = effective_arch(configuration_manager.network_arch_class_name,
                                 configuration_manager.network_arch_init_kwargs, trainer_name, init_args,
                                 plans_manager.plans.get('plans_name', ''), parameters[0])
        network = CompiledNetwork(cls, kw, num_input_channels, label_manager.num_segmentation_heads,
                                  configuration_manager.patch_size)

        self.plans_manager = plans_manager
        self.configuration_manager = configuration_manager
        self.list_of_parameters = parameters
        network.load_state_dict(parameters[0])
        self.network = network
        self.dataset_json = dataset_json
        self.trainer_name = trainer_name
        self.allowed_mirroring_axes = inference_allowed_mirroring_axes
        self.label_manager = label_manager

    def manual_initialization(self, network, plans_manager: PlansManager,
                              configuration_manager: ConfigurationManager, parameters: Optional[List[dict]],
                              dataset_json: dict, trainer_name: str,
                              inference_allowed_mirroring_axes: Optional[Tuple[int, ...]]):
        """This is used by the nnUNetTrainer to initialize nnUNetPredictor for the final validation.
        `network` may be a live torch.nn.Module (its state_dict is lowered) or a CompiledNetwork."""
        self.plans_manager = plans_manager
        self.configuration_manager = configuration_manager
        self.dataset_json = dataset_json
        self.trainer_name = trainer_name
        self.allowed_mirroring_axes = inference_allowed_mirroring_axes
        self.label_manager = plans_manager.get_label_manager(dataset_json)
        if not isinstance(network, CompiledNetwork):
            sd = network.state_dict()
            num_input_channels = determine_num_input_channels(plans_manager, configuration_manager, dataset_json)
            mod = getattr(network, 'module', network)
            mod = getattr(mod, '_orig_mod', mod)
            cname = type(getattr(mod, 'network', mod)).__name__
            cls = configuration_manager.network_arch_class_name
            if cname in ('LiteResEncStudent', 'ResidualEncoderUNet'):
                cls = 'dynamic_network_architectures.architectures.unet.ResidualEncoderUNet'
            elif cname in ('LiteNNUNetStudent', 'PlainConvUNet'):
                cls = 'dynamic_network_architectures.architectures.unet.PlainConvUNet'
            compiled = CompiledNetwork(cls, configuration_manager.network_arch_init_kwargs, num_input_channels,
                                       self.label_manager.num_segmentation_heads, configuration_manager.patch_size)
            if parameters is None:
                parameters = [sd]
            compiled.load_state_dict(parameters[0])
            network = compiled
        self.list_of_parameters = parameters
        self.network = network

    @staticmethod
    def auto_detect_available_folds(model_training_output_dir, checkpoint_name):
        print('use_folds is None, attempting to auto detect available folds')
        fold_folders = [i for i in sorted(os.listdir(model_training_output_dir))
                        if i.startswith('fold_') and os.path.isdir(os.path.join(model_training_output_dir, i))]
        fold_folders = [i for i in fold_folders if i != 'fold_all']
        fold_folders = [i for i in fold_folders
                        if os.path.isfile(os.path.join(model_training_output_dir, i, checkpoint_name))]
        use_folds = [int(i.split('_')[-1]) for i in fold_folders]
        print(f'found the following folds: {use_folds}')
        return use_folds

    # ------------------------------------------------------------------ geometry helpers
    def _flip_masks(self) -> bytes:
        """Mirror combinations in the reference's order (:550-553): identity first, then every non-empty
        subset of the allowed axes ordered by size then lexicographically.  bit a = flip spatial axis a."""
        mirror_axes = self.allowed_mirroring_axes if self.use_mirroring else None
        masks = [0]
        if mirror_axes is not None:
            assert max(mirror_axes) <= 2, 'mirror_axes does not match the dimension of the input!'
            axes = list(mirror_axes)
            for i in range(len(axes)):
                for c in itertools.combinations(axes, i + 1):
                    masks.append(sum(1 << int(a) for a in c))
        return bytes(masks)

    def _internal_get_sliding_window_slicers(self, image_size: Tuple[int, ...]):
        """Same return value as the reference method (:506-538) for 3-D patches."""
        patch = self.configuration_manager.patch_size
        assert len(patch) == len(image_size) == 3, 'only 3-D patch sizes are supported by the B200 engine'
        starts = sw.tile_starts(image_size, patch, self.tile_step_size)
        return [tuple([slice(None), *[slice(int(si), int(si) + int(ti)) for si, ti in zip(s, patch)]]) for s in starts]

    def _gaussian(self, patch) -> Optional[torch.Tensor]:
        if not self.use_gaussian:
            return None
        key = (tuple(patch), str(self.device))
        if key not in self._gauss_cache:
            g = sw.compute_gaussian(tuple(patch), sigma_scale=1. / 8, value_scaling_factor=10, dtype=np.float16)
            self._gauss_cache = {key: torch.from_numpy(g).to(self.device)}
        return self._gauss_cache[key]

    def _choose_tiles_per_batch(self, n_flips: int, n_tiles: int) -> int:
        if self.tiles_per_batch is not None:
            return max(1, min(int(self.tiles_per_batch), n_tiles))
        # 32 patches per launch sequence keeps the low-resolution layers (8^3, 4^3 voxels per patch) wide enough
        # for 148 SMs; activations for 32 x 128^3 student patches take 15 GB of the 180 GB
        return max(1, min(n_tiles, max(1, 32 // n_flips)))

    def _engine_capacity(self, n_flips: int) -> int:
        """Patches per launch sequence the engines are built for — independent of the volume at hand, so that a
        small volume (or a short shard) never builds a second engine + workspace."""
        tpb = int(self.tiles_per_batch) if self.tiles_per_batch is not None else max(1, 32 // n_flips)
        return max(1, tpb) * n_flips

    @staticmethod
    def _batch_sizes(n_tiles: int, tpb: int):
        """Even split into ceil(n / tpb) batches: 37 tiles -> 4,4,4,4,4,4,4,3,3,3 instead of 9 x 4 + 1 (a 1-tile
        batch runs the 8^3 / 4^3 layers on a handful of SMs)."""
        nb = -(-n_tiles // tpb)
        base, rem = divmod(n_tiles, nb)
        return [base + 1] * rem + [base] * (nb - rem)

    def _weight_sum(self, vol, patch, x_offset: int = 0, n_planes: Optional[int] = None, scale: float = 1.0,
                    dtype=torch.float32) -> torch.Tensor:
        """n_predictions (:590, :614) for planes [x_offset, x_offset + n_planes) of a volume, times `scale` (fold
        count).  It does not depend on the image: computed once per geometry and kept."""
        n_planes = vol[0] if n_planes is None else n_planes
        key = (tuple(vol), tuple(patch), float(self.tile_step_size), bool(self.use_gaussian), str(dtype),
               str(self.device), int(x_offset), int(n_planes), float(scale))
        w = self._wsum_cache.get(key)
        if w is None:
            w = torch.empty((n_planes, vol[1], vol[2]), dtype=dtype, device=self.device)
            steps = sw.compute_steps_for_sliding_window(vol, patch, self.tile_step_size)
            steps = [[s_ - x_offset for s_ in steps[0]], steps[1], steps[2]]
            E.weight_sum(steps, patch, self._gaussian(patch), w)
            if scale != 1.0:
                assert dtype == torch.float32
                E.scale_inplace(w, scale)
            while len(self._wsum_cache) >= 2:
                self._wsum_cache.pop(next(iter(self._wsum_cache)))
            self._wsum_cache[key] = w
        return w

    def to_host(self, t: torch.Tensor) -> torch.Tensor:
        """Device -> pinned host copy through a buffer that is reused across volumes (a pageable `.cpu()` of a
        105 MB label map costs as much as the whole 8-GPU prediction)."""
        key = (tuple(t.shape), t.dtype)
        buf = self._pinned_cache.get(key)
        if buf is None:
            buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            self._pinned_cache = {key: buf}
        buf.copy_(t, non_blocking=True)
        torch.cuda.current_stream(t.device).synchronize()
        return buf

    # ------------------------------------------------------------------ the hot path
    @torch.inference_mode()
    def _sliding_window_accumulate(self, data: torch.Tensor, starts: np.ndarray, acc: torch.Tensor,
                                   acc_origin=(0, 0, 0), data_origin=(0, 0, 0), all_folds: bool = False):
        """Runs every tile in `starts` (volume coordinates) and adds its weighted prediction into `acc`, whose
        voxel (0,0,0) sits at volume coordinate `acc_origin`; `data` voxel (0,0,0) sits at `data_origin`.
        With `all_folds` every parameter set of list_of_parameters runs on the SAME gathered batch and is summed
        into the same accumulator (the caller divides by the fold count through the weight sum)."""
        patch = tuple(self.configuration_manager.patch_size)
        flips = self._flip_masks()
        nf = len(flips)
        tpb = self._choose_tiles_per_batch(nf, len(starts))
        cap = max(self._engine_capacity(nf), tpb * nf)
        if all_folds and self.list_of_parameters:
            engines = [self.network.engine(self.device, cap, params) for params in self.list_of_parameters]
        else:
            engines = [self.network.engine(self.device, cap)]
        prog = engines[0].program
        heads = self.label_manager.num_segmentation_heads
        assert prog.num_heads == heads
        gauss = self._gaussian(patch)
        starts = np.ascontiguousarray(starts, dtype=np.int32)
        local = starts - np.asarray(acc_origin, dtype=np.int32)[None]
        in_vol = np.ascontiguousarray(starts - np.asarray(data_origin, dtype=np.int32)[None])
        starts_dev = torch.from_numpy(in_vol).to(self.device)
        in_cs = prog.buffers[prog.input_buffer][1]
        out_cs = prog.buffers[prog.output_buffer][1]
        launches = 0
        m0 = E.mem_launches()
        self.last_tiles_per_batch = tpb
        i = 0
        for n in self._batch_sizes(len(starts), tpb):
            self._mark('gather')
            # folds share the workspace: ONE gather feeds every fold's forward
            E.gather_tiles(data, starts_dev[i:i + n], n, patch, flips, engines[0].buffer_ptr(prog.input_buffer), in_cs)
            for eng in engines:
                assert eng.buffer_ptr(prog.input_buffer) == engines[0].buffer_ptr(prog.input_buffer)
                self._mark('forward')
                eng.forward(n * nf)
                self._mark('accumulate')
                E.accumulate_tiles(eng.buffer_ptr(prog.output_buffer), _lib.IN_F16, out_cs, heads, local[i:i + n], patch,
                                   flips, gauss, acc)
                launches += eng.launch_counts()[0]
            self._mark(None)
            i += n
        self.last_launches += launches + (E.mem_launches() - m0)

    @torch.inference_mode()
    def profile_dominant_op(self, data: torch.Tensor, n_batches: int = 6, op_name: str = None):
        """bench.py: device time (CUDA events on the launching stream, inside libfnnu) of the network operator
        with the most FLOPs (or of the operator called `op_name`), over `n_batches` real tile batches of `data`."""
        patch = tuple(self.configuration_manager.patch_size)
        flips = self._flip_masks()
        nf = len(flips)
        starts = sw.tile_starts(tuple(data.shape[1:]), patch, self.tile_step_size)
        tpb = self._choose_tiles_per_batch(nf, len(starts))
        eng = self.network.engine(self.device, max(self._engine_capacity(nf), tpb * nf))
        prog = eng.program
        flops = [o.flops(prog.buffers[o.src][0], prog.buffers[o.dst][0]) for o in prog.ops]
        idx = int(np.argmax(flops))
        if op_name is not None:
            names = [o.name for o in prog.ops]
            if op_name not in names:
                raise KeyError(f'no operator named {op_name!r} in the program')
            idx = names.index(op_name)
        eng.profile_op(idx)
        starts_dev = torch.from_numpy(np.ascontiguousarray(starts, dtype=np.int32)).to(self.device)
        times = []
        for k in range(min(n_batches, len(starts) // tpb)):
            E.gather_tiles(data, starts_dev[k * tpb:(k + 1) * tpb], tpb, patch, flips, eng.buffer_ptr(prog.input_buffer),
                           prog.buffers[prog.input_buffer][1])
            eng.forward(tpb * nf)
            times.append(eng.profile_ms())
        eng.profile_op(-1)
        op = prog.ops[idx]
        in_d, out_d = prog.buffers[op.src][0], prog.buffers[op.dst][0]
        n = tpb * nf
        return {'op': op.name, 'cin': op.cin, 'cout': op.cout, 'out_dims': list(out_d), 'kernel': list(op.kernel),
                'patches_per_launch': n,
                'flop_per_launch': flops[idx] * n, 'ms': float(np.mean(times)), 'ms_all': [float(t) for t in times],
                'algorithmic_bytes_per_launch': float(n * 2 * (np.prod(in_d) * op.cin + np.prod(out_d) * op.cout))}

    # ------------------------------------------------------------------ phase timing (bench only)
    def _mark(self, phase):
        """Closes the running phase and opens `phase` with CUDA events on the launching stream."""
        if not self.collect_timing:
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        cur = self.timing.get('_open')
        if cur is not None:
            self.timing.setdefault(cur[0], []).append((cur[1], ev))
        self.timing['_open'] = (phase, ev) if phase is not None else None

    def timing_summary(self):
        """ms per phase summed over everything recorded since `timing` was reset, divided by the number of
        volumes (calls of the tile loop's owner)."""
        torch.cuda.synchronize()
        out = {}
        n = max(1, self.timing.get('_volumes', 1))
        for k, pairs in self.timing.items():
            if k.startswith('_'):
                continue
            out[k + '_ms'] = sum(a.elapsed_time(b) for a, b in pairs) / n
        return out

    @torch.inference_mode()
    def _internal_predict_sliding_window_return_logits(self, data: torch.Tensor, slicers,
                                                       do_on_device: bool = True, return_labels: bool = False,
                                                       all_folds: bool = False):
        if not do_on_device:
            raise RuntimeError('the B200 engine keeps the result arrays on the device; there is no CPU results path')
        patch = tuple(self.configuration_manager.patch_size)
        heads = self.label_manager.num_segmentation_heads
        data = data.to(self.device, dtype=torch.float32).contiguous()
        vol = tuple(data.shape[1:])
        starts = np.asarray([[s.start for s in sl[1:]] for sl in slicers], dtype=np.int32)
        if self.collect_timing:
            self.timing['_volumes'] = self.timing.get('_volumes', 0) + 1
        acc = torch.zeros((heads, *vol), dtype=self.accumulator_dtype, device=self.device)
        self._sliding_window_accumulate(data, starts, acc, all_folds=all_folds)
        self._mark('weight_sum')
        n_folds = len(self.list_of_parameters) if (all_folds and self.list_of_parameters) else 1
        m0 = E.mem_launches()
        wsum = self._weight_sum(vol, patch, scale=float(n_folds), dtype=self.accumulator_dtype)
        self._mark('finalize')
        inf_flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        logits = None if return_labels == 'only' else torch.empty((heads, *vol), dtype=torch.float16, device=self.device)
        labels = torch.empty(vol, dtype=torch.uint8, device=self.device) if return_labels else None
        E.finalize(acc, wsum, logits, labels, inf_flag)
        self._mark(None)
        self.last_launches += 1 + (E.mem_launches() - m0)
        del acc
        if int(inf_flag.item()) != 0:
            raise RuntimeError('Encountered inf in predicted array. Aborting... If this problem persists, '
                               'reduce value_scaling_factor in compute_gaussian or increase the dtype of '
                               'predicted_logits to fp32')
        return (logits, labels) if return_labels else logits

    def _check_ready(self, input_image):
        assert isinstance(input_image, torch.Tensor)
        if self.network is None:
            raise RuntimeError('predictor is not initialised')
        E.require_cuda_device(self.device)
        assert input_image.ndim == 4, 'input_image must be a 4D np.ndarray or torch.Tensor (c, x, y, z)'

    def _pad(self, input_image: torch.Tensor):
        patch = self.configuration_manager.patch_size
        below, above = sw.pad_amounts(input_image.shape[1:], patch)
        data = input_image.to(self.device, dtype=torch.float32)
        if any(b or a for b, a in zip(below, above)):
            pad = []
            for b, a in zip(below[::-1], above[::-1]):
                pad += [b, a]
            data = torch.nn.functional.pad(data, pad, mode='constant', value=0)
        slicer = tuple(slice(b, b + s) for b, s in zip(below, input_image.shape[1:]))
        return data.contiguous(), slicer

    @torch.inference_mode()
    def predict_sliding_window_return_logits(self, input_image: torch.Tensor, *, _all_folds: bool = False) -> torch.Tensor:
        self._check_ready(input_image)
        self.network = self.network.to(self.device)
        self.network.eval()
        self.last_launches = 0
        if self.verbose:
            print(f'Input shape: {input_image.shape}')
            print('step_size:', self.tile_step_size)
            print('mirror_axes:', self.allowed_mirroring_axes if self.use_mirroring else None)
        with torch.cuda.device(self.device):
            data, slicer_revert_padding = self._pad(input_image)
            slicers = self._internal_get_sliding_window_slicers(data.shape[1:])
            predicted_logits = self._internal_predict_sliding_window_return_logits(data, slicers, True,
                                                                                   all_folds=_all_folds)
            predicted_logits = predicted_logits[(slice(None), *slicer_revert_padding)]
        return predicted_logits

    @torch.inference_mode()
    def predict_sliding_window_return_segmentation(self, input_image: torch.Tensor) -> torch.Tensor:
        """Extension: label map (uint8, on device) with the argmax fused into the normalisation pass, so
        the (heads, x, y, z) logits never cross PCIe.  Equals
        label_manager.convert_logits_to_segmentation(predict_sliding_window_return_logits(x))."""
        self._check_ready(input_image)
        assert not self.label_manager.has_regions, 'region-based label maps go through the logits path'
        self.network = self.network.to(self.device)
        self.last_launches = 0
        with torch.cuda.device(self.device):
            data, slicer_revert_padding = self._pad(input_image)
            slicers = self._internal_get_sliding_window_slicers(data.shape[1:])
            _, labels = self._internal_predict_sliding_window_return_logits(data, slicers, True, return_labels='only')
            return labels[slicer_revert_padding]

    @torch.inference_mode()
    def predict_sliding_window_sharded(self, input_image: torch.Tensor, group=None, return_labels: bool = False,
                                       gather_to: Optional[int] = 0):
        """ONE volume across the ranks of `group` (one process per GPU): the tile list is cut into contiguous
        runs (x-slabs in the reference's tile order), each rank accumulates the planes its tiles touch, the
        overlapping partial sums are exchanged once (NCCL point-to-point over NVLink) and every rank normalises
        the planes it owns.  Every rank passes the same `input_image`.  Returns, on rank `gather_to`, the full
        (heads, x, y, z) fp16 logits (or the uint8 label map if `return_labels`); on other ranks their own slab.
        With gather_to=None every rank returns (slab, (x_lo, x_hi))."""
        import torch.distributed as dist
        from . import sharding
        self._check_ready(input_image)
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.network = self.network.to(self.device)
        self.last_launches = 0
        patch = tuple(self.configuration_manager.patch_size)
        heads = self.label_manager.num_segmentation_heads
        with torch.cuda.device(self.device):
            below, above = sw.pad_amounts(input_image.shape[1:], patch)
            padded = any(b or a for b, a in zip(below, above))
            if padded or input_image.device.type == 'cuda':
                data, revert = self._pad(input_image)            # tiny volume (or already resident): whole volume
                vol = tuple(data.shape[1:])
            else:
                data, vol = None, tuple(int(v) for v in input_image.shape[1:])
                revert = tuple(slice(0, v) for v in vol)
            starts = sw.tile_starts(vol, patch, self.tile_step_size)
            plan = sharding.plan_shards(starts, patch, vol, world)
            lo, hi = plan.tile_ranges[rank]
            l0, l1 = plan.local[rank]
            o0, o1 = plan.owned[rank]
            if self.collect_timing:
                self.timing['_volumes'] = self.timing.get('_volumes', 0) + 1
            acc = torch.zeros((heads, l1 - l0, vol[1], vol[2]), dtype=torch.float32, device=self.device)
            if hi > lo:
                if data is None:
                    # host input: upload only the planes this rank's tiles read (slab), not the whole volume
                    s0, s1 = plan.slabs[rank]
                    part = input_image[:, s0:s1]
                    if not part.is_pinned():
                        part = part.contiguous()
                    slab = part.to(self.device, dtype=torch.float32, non_blocking=True).contiguous()
                    self._sliding_window_accumulate(slab, starts[lo:hi], acc, acc_origin=(l0, 0, 0),
                                                    data_origin=(s0, 0, 0))
                else:
                    self._sliding_window_accumulate(data, starts[lo:hi], acc, acc_origin=(l0, 0, 0))
            self._mark('exchange')
            sharding.exchange_halos(acc, plan, rank, E.add_inplace, group)
            self._mark('weight_sum')
            n_own = o1 - o0
            out = None
            if n_own > 0:
                m0 = E.mem_launches()
                wsum = self._weight_sum(vol, patch, x_offset=o0, n_planes=n_own)
                self._mark('finalize')
                inf_flag = torch.zeros(1, dtype=torch.int32, device=self.device)
                view = acc[:, o0 - l0:o1 - l0]
                if return_labels:
                    out = torch.empty((n_own, vol[1], vol[2]), dtype=torch.uint8, device=self.device)
                    E.finalize(view, wsum, None, out, inf_flag)
                else:
                    out = torch.empty((heads, n_own, vol[1], vol[2]), dtype=torch.float16, device=self.device)
                    E.finalize(view, wsum, out, None, inf_flag)
                self.last_launches += 1 + (E.mem_launches() - m0)
                if int(inf_flag.item()) != 0:
                    raise RuntimeError('Encountered inf in predicted array.')
            self._mark('gather_result')
            if gather_to is None or world == 1:
                self._mark(None)
                if world == 1:
                    return out[revert] if return_labels else out[(slice(None), *revert)]
                return out, (o0, o1)
            full = None
            ops = []
            if rank == gather_to:
                full = torch.empty((vol if return_labels else (heads, *vol)),
                                   dtype=torch.uint8 if return_labels else torch.float16, device=self.device)
                for r in range(world):
                    a, b = plan.owned[r]
                    if b <= a:
                        continue
                    if r == rank:
                        if return_labels:
                            full[a:b] = out
                        else:
                            full[:, a:b] = out
                    elif return_labels:
                        ops.append(dist.P2POp(dist.irecv, full[a:b], sharding.global_rank(group, r), group=group))
                    else:
                        for h in range(heads):
                            ops.append(dist.P2POp(dist.irecv, full[h, a:b], sharding.global_rank(group, r), group=group))
            elif n_own > 0:
                if return_labels:
                    ops.append(dist.P2POp(dist.isend, out, sharding.global_rank(group, gather_to), group=group))
                else:
                    for h in range(heads):
                        ops.append(dist.P2POp(dist.isend, out[h], sharding.global_rank(group, gather_to), group=group))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
            self._mark(None)
            if rank == gather_to:
                return full[revert] if return_labels else full[(slice(None), *revert)]
            return out

    @torch.inference_mode()
    def predict_logits_from_preprocessed_data(self, data: torch.Tensor, *, on_device: bool = False) -> torch.Tensor:
        """Fold ensemble of the reference (:471-504).  The reference runs the whole sliding window once per fold,
        copies every fold's logits to the CPU and averages there; here every fold's engine is resident, all folds
        run on each gathered tile batch and are summed into ONE accumulator on the device (mean over folds = the
        weight sum times the fold count), and the result is returned on the CPU like the reference's
        (`on_device=True` leaves it on the GPU for the export step)."""
        n_folds = len(self.list_of_parameters) if self.list_of_parameters else 1
        if n_folds > 1 and self.accumulator_dtype == torch.float32:
            prediction = self.predict_sliding_window_return_logits(data, _all_folds=True)
        else:
            prediction = None
            for params in (self.list_of_parameters or [None]):
                if params is not None:
                    self.network.load_state_dict(params)
                if prediction is None:
                    prediction = self.predict_sliding_window_return_logits(data)
                else:
                    prediction += self.predict_sliding_window_return_logits(data)
            if n_folds > 1:
                prediction /= n_folds
        if self.verbose:
            print('Prediction done')
        return prediction if on_device else prediction.to('cpu')

    def predict_single_npy_array(self, input_image: np.ndarray, image_properties: dict,
                                 segmentation_previous_stage: np.ndarray = None,
                                 output_file_truncated: str = None,
                                 save_or_return_probabilities: bool = False):
        """Array-in / label-map-out entry (:423-468).  Pre-processing, the fold ensemble and the
        export to the image's own geometry (§8 f1) run on the device: the raw image goes up once, one uint8 per voxel
        comes back.  Writing files (`output_file_truncated`) is outside the inference path."""
        from . import export, preprocess
        if output_file_truncated is not None:
            raise NotImplementedError('file export is outside the B200 inference path (SURVEY.md §8 f1)')
        E.require_cuda_device(self.device)
        with torch.cuda.device(self.device):
            data, props = preprocess.run_case_npy(input_image, image_properties, segmentation_previous_stage,
                                                  self.plans_manager, self.configuration_manager, self.dataset_json,
                                                  self.label_manager, self.device)
            logits = self.predict_logits_from_preprocessed_data(data, on_device=True)
            return export.convert_predicted_logits_to_segmentation_with_correct_shape(
                logits, self.plans_manager, self.configuration_manager, self.label_manager, props,
                save_or_return_probabilities, to_host=self.to_host)
