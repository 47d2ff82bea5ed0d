"""Host-side helpers of the training scripts that sit on the optimiser boundary.

``get_parameter_groups_with_lrs`` follows flow2gan/utils.py:69-138: a module anywhere in the tree
may carry a float attribute ``lr_scale``; a parameter's learning rate is the base rate times the
``lr_scale`` of every module on its dotted path (root included).  Parameters are grouped by the
resulting rate, in first-seen order, as ScaledAdam's ``named_params`` / ``params`` groups
(pretrain.py / finetune.py build their optimisers from this)."""
from __future__ import annotations

import logging
from typing import Dict, List, Sequence

from torch import nn


def get_parameter_groups_with_lrs(model: nn.Module, lr: float, include_names: bool = False,
                                  freeze_modules: Sequence[str] = ()) -> List[dict]:
    scales: Dict[str, float] = {name: m.lr_scale for name, m in model.named_modules() if hasattr(m, "lr_scale")}

    def scale_of(prefix: str) -> float:
        return scales.get(prefix, 1.0)

    groups: Dict[float, list] = {}
    for name, param in model.named_parameters():
        parts = name.split(".")
        top = parts[1] if parts[0] == "module" and len(parts) > 1 else parts[0]     # DDP wrapper
        if top in freeze_modules:
            logging.info(f"Remove {name} from parameters")
            continue
        # same multiplication order as the reference, so equal rates compare equal as dict keys
        prefix = parts[0]
        rate = lr * scale_of(prefix)
        if prefix != "":
            rate *= scale_of("")
        for part in parts[1:]:
            prefix = prefix + "." + part
            rate *= scale_of(prefix)
        groups.setdefault(rate, []).append((name, param) if include_names else param)
    key = "named_params" if include_names else "params"
    return [{key: members, "lr": rate} for rate, members in groups.items()]
