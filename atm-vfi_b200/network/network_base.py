"""Base ATM-VFI network (51.56 M parameters) - drop-in for the reference's network/network_base.py."""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.dirname(_HERE), _HERE):       # importable both as ``network_base`` and ``network.network_base``
    if _p not in sys.path:
        sys.path.insert(0, _p)

from network._network import NetworkBase, _REFINE_PARTS  # noqa: E402
from atmvfi.arch import BASE


class Network(NetworkBase):
    ARCH = BASE

    def __finetune_refinenet_only__(self):          # network_base.py:316-334 (Base only)
        self.requires_grad_(False)
        self._grad(_REFINE_PARTS, True)
