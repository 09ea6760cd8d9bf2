"""Parameter discovery that also works on nn.DataParallel replicas."""
from __future__ import annotations

import torch.nn as nn


def is_replica(module: nn.Module) -> bool:
    """torch.nn.parallel.replicate() marks its shallow copies with a `_former_parameters` dict."""
    return "_former_parameters" in module.__dict__


def state_tensors(module: nn.Module):
    """The tensors of module.state_dict() in registration order.  A replica's `_parameters` dicts are empty (its
    broadcast copies are plain attributes listed in `_former_parameters`), so state_dict() would return nothing."""
    if not is_replica(module):
        return [p for _, p in module.state_dict(keep_vars=True).items()]
    out = []
    for mod in module.modules():
        src = mod.__dict__.get("_former_parameters", {})
        out += [src[n] for n in src]
        out += [v for n, v in mod._buffers.items() if v is not None and n not in mod._non_persistent_buffers_set]
    return out
