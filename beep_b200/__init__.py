"""beep_b200 -- B200-native engine for BEEP's batched latent-state hot path.

``beep_b200.engine`` binds the C-ABI shared library (hand-written sm_100a CUDA,
``beep_b200/csrc``); ``beep_b200.model`` holds the flat model description;
``beep_b200.synth`` generates the synthetic scale-up populations.
"""
from .model import ModelDesc  # noqa: F401
