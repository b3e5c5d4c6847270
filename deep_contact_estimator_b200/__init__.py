"""deep_contact_estimator_b200 — B200-native contact-classification inference path.

Drop-in for the hot path of UMich-CURLY/deep-contact-estimator:
the ``contact_cnn`` module surface, ``contact_dataset``, and the
``inference`` / ``inference_and_compute_acc`` / ``compute_accuracy`` loops, over
hand-written sm_100a kernels behind the C ABI in ``include/dce.h``.
"""
from .synth import WINDOW, CHANNELS, CLASSES, PARAM_NAMES, PARAM_SHAPES
from .engine import ContactEngine, LatencyRunner, RowRunner, default_precision
from .contact_cnn import contact_cnn
from .data_handler import contact_dataset
from .inference import (inference, inference_and_compute_acc, compute_accuracy, decimal2binary, evaluate,
                        metrics_from_counts, counts_from_arrays)
from .realtime import RealtimeContactEstimator

__all__ = [
    "contact_cnn", "contact_dataset", "ContactEngine", "LatencyRunner", "RowRunner", "RealtimeContactEstimator", "inference", "inference_and_compute_acc",
    "compute_accuracy", "decimal2binary", "default_precision", "evaluate", "metrics_from_counts", "counts_from_arrays",
    "WINDOW", "CHANNELS", "CLASSES", "PARAM_NAMES", "PARAM_SHAPES",
]
