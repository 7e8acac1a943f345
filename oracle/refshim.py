"""Import the UNMODIFIED reference from /root/reference (build container only).  TEST INFRASTRUCTURE ONLY.

The reference needs ``timm`` (three symbols) plus, for demo_2x, ``imageio`` / ``flow_vis`` at import
time; none is installed and there is no network.  The shim below provides exactly those names
(SURVEY.md section 8c).  /root/reference does not exist on the GPU box, so nothing that runs there may
import this module: it is used by ``oracle/gen_golden.py`` and by tests that skip when the tree is absent.
"""
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("ATMVFI_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "network"))


def install() -> None:
    if "timm" not in sys.modules:
        timm, models, layers = (types.ModuleType(n) for n in ("timm", "timm.models", "timm.models.layers"))
        layers.trunc_normal_ = torch.nn.init.trunc_normal_
        layers.to_2tuple = lambda x: x if isinstance(x, (tuple, list)) else (x, x)
        layers.DropPath = torch.nn.Identity
        timm.models, models.layers = models, layers
        sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})
    for stub in ("imageio", "flow_vis"):
        if stub not in sys.modules:
            m = types.ModuleType(stub)
            m.imread = m.imwrite = None          # names benchmark/utils.py:8 imports; never called on the forward path
            sys.modules[stub] = m


def load_reference_network(kind: str):
    """Returns the reference ``Network`` class for ``kind`` in {'base','lite'}, imported under a private
    module name so that it never collides with this repo's own ``network_base`` / ``network_lite``."""
    import importlib.util

    install()
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k in ("flow_warp", "network", "network.attention")}
    sys.path[:0] = [REF_ROOT, os.path.join(REF_ROOT, "network")]
    try:
        spec = importlib.util.spec_from_file_location(f"_ref_network_{kind}", os.path.join(REF_ROOT, "network", f"network_{kind}.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        ref_mods = {k: sys.modules[k] for k in ("flow_warp", "network", "network.attention") if k in sys.modules}
    finally:
        del sys.path[:2]
        for k in ("flow_warp", "network", "network.attention"):
            sys.modules.pop(k, None)
        sys.modules.update(saved)
    mod._ref_modules = ref_mods
    return mod.Network
