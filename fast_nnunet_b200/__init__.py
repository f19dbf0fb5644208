"""fast_nnunet_b200 — B200-native (sm_100a) sliding-window inference for nnU-Net v2 / FastnnUNet.

Public surface mirrors the reference's inference entry point
(distillation/nnunetv2/inference/predict_from_raw_data.py): `nnUNetPredictor`.
Importing the package does not need a GPU; running a prediction does (no CPU fallback).
"""
from .predictor import nnUNetPredictor, CompiledNetwork  # noqa: F401
from .sliding_window import compute_gaussian, compute_steps_for_sliding_window  # noqa: F401

__all__ = ['nnUNetPredictor', 'CompiledNetwork', 'compute_gaussian', 'compute_steps_for_sliding_window']
