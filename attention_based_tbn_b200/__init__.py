"""attention_based_tbn_b200 -- B200 (sm_100a) implementation of the Temporal Binding Network forward.

Drop-in surface (same names as the reference's core/models): ``TBNModel``, ``build_model``.
"""
from .config import Cfg, baseline_config, default_cfg, get_modality  # noqa: F401
from .evaluation import Metric, save_scores  # noqa: F401
from .model import TBNModel  # noqa: F401
from .model_builder import build_model  # noqa: F401

__all__ = ["TBNModel", "build_model", "Metric", "save_scores", "Cfg", "default_cfg", "baseline_config", "get_modality"]
