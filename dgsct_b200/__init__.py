"""dgsct_b200 -- B200-native (sm_100a) engine for the DG-SCT cross-modal adapter.

Public surface mirrors the reference's call boundary (DG-SCT/AVE/nets/net_trans.py:433-674):

    from dgsct_b200 import VisualAdapter          # same constructor, state_dict and forward contract
    dgsct_b200.patch(nets.net_trans, "ave")       # swap the class inside a reference task module
    dgsct_b200.AdapterDDP(adapters)               # one process per GPU: bucketed NCCL all-reduce of the adapter gradients
    dgsct_b200.FlatAdam(ddp.bucketer, lr=...)     # Adam over the flat gradient buckets, one fused kernel per bucket
"""
from .adapter import VisualAdapter, patch, adapter_forward_raw, VARIANT_DEFAULTS  # noqa: F401
from .adapter import adapter_backward_raw, adapter_pair  # noqa: F401
from .ddp import AdapterDDP, FlatAdam, GradBucketer  # noqa: F401
from .heads import TemporalGates, pool_tokens  # noqa: F401
from ._lib import build, lib  # noqa: F401

__all__ = ["VisualAdapter", "patch", "adapter_forward_raw", "adapter_backward_raw", "adapter_pair", "AdapterDDP", "FlatAdam", "GradBucketer", "build", "lib",
           "VARIANT_DEFAULTS", "TemporalGates", "pool_tokens"]
