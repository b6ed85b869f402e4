"""Loss aggregation (reference ``simlingo_training/models/utils.py:7-41``)."""
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from simlingo_training.utils.custom_types import TrainingOutput


def summarise_losses(loss_dict: Dict[str, Tuple[Tensor, Tensor]], weights: Optional[Dict[str, float]] = None) -> TrainingOutput:
    """``loss = sum_k w_k * (sum(values_k) / sum(counts_k))`` with empty terms contributing 0.

    ``loss_dict[k] = (values [B], counts [B])``: per-sample summed loss and the number of items it sums over."""
    values, counts, averages = {}, {}, {}
    for key, (v, n) in loss_dict.items():
        values[key], counts[key] = v, n
        total = n.sum()
        averages[key] = torch.where(total > 0, v.sum() / total, 0.0)
    if weights is None:
        terms = list(averages.values())
    else:
        terms = [weights.get(k, 1.0) * a for k, a in averages.items()]
    return TrainingOutput(loss=torch.stack(terms).sum(), loss_averages=averages, loss_values=values, loss_counts=counts)
