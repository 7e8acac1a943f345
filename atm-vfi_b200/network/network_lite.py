"""Lite ATM-VFI network (11.98 M parameters) - drop-in for the reference's network/network_lite.py."""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.dirname(_HERE), _HERE):       # importable both as ``network_base`` and ``network.network_base``
    if _p not in sys.path:
        sys.path.insert(0, _p)

from network._network import NetworkBase  # noqa: E402
from atmvfi.arch import LITE


class Network(NetworkBase):
    ARCH = LITE
