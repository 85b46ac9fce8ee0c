"""BatchNorm running-stat calibration for the synthetic test weights -- TEST INFRASTRUCTURE.

A deep random ReLU network with arbitrary BN statistics maps every input to (almost) the same
feature vector (hard part 1), which would make parity tests vacuous.  This
does what the survey prescribes: one pass over seeded synthetic frames in which every
BatchNorm takes its running mean/var from the batch (train-mode BN with momentum=1, i.e.
``running_var`` = unbiased batch variance), layer by layer.  The resulting statistics are
committed as ``tests/golden/bn_calib.npz`` (about 10k channels per backbone) so that every box
uses bit-identical weights; ``attention_based_tbn_b200.synthetic.make_state_dict(calib=...)``
reads that file (plain data, no oracle import).
"""
from __future__ import annotations

from typing import Dict, List

import torch

from . import tbn_oracle


def calibrate_backbones(sd, cfg, modality: List[str], inputs: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Overwrites ``Base_*.*_bn.running_{mean,var}`` in ``sd``; returns just those entries."""
    tbn_oracle._CALIBRATE = True
    try:
        with torch.no_grad():
            for m in modality:
                b, n, c, h, w = inputs[m].shape
                tbn_oracle.bninception_features(sd, f"Base_{m}.", inputs[m].reshape(b * n, c, h, w).float())
    finally:
        tbn_oracle._CALIBRATE = False
    return {k: v.clone() for k, v in sd.items() if k.endswith("running_mean") or k.endswith("running_var")}
