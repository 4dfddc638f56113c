"""Make the reference's own `models/`, `trainRGB.py`, `trainmask.py` run on the B200 modules, unchanged.

The reference imports its hot-path layers by module name:
    layers/Masked_Attention.py:6   from .masked_win_attention import *
    layers/Attention.py:6          from .win_attention import *
    layers/TransformRGB.py:4       from .GDN import *          (also layers/SupplyMask.py:4)
    layers/TransformRGB.py:10      from .Masked_Attention import *   (optional: the wrapper with the fused gate)
    models/*.py:3                  from layers.SupplyMask import *   (optional: the alpha pyramid in two launches)
`install(reference_root)` puts the reference tree on sys.path and pre-seeds `sys.modules` with this
package's drop-ins under those three names, so every later `import layers.X` / `from .X import *`
inside the reference resolves to the CUDA-backed modules.  `patch_model_rounding(model_module)`
swaps the `ste_round` helper the model files define for themselves (models/AutoEncoderRGB_Journal.py:31).

    import mwa_b200; mwa_b200.install("/path/to/reference")
    from models.AutoEncoderRGB_Journal import AutoEncoder      # the reference's file, our kernels
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

_DROPINS = ("GDN", "masked_win_attention", "win_attention")


def install(reference_root: str, extra_paths=(), fuse_gate: bool = True, fuse_pyramid: bool = True) -> None:
    """fuse_gate: also replace `layers.Masked_Attention` (the Win_noShift_Attention wrapper, layers/TransformRGB.py:10)
    with the drop-in whose gate `a * sigmoid(b) + x` is one fused kernel (same state-dict keys).
    fuse_pyramid: also replace `layers.SupplyMask` (the alpha pyramid, layers/SupplyMask.py:7-18) with the two-launch
    CUDA pyramid (bit-exact with the six AvgPool2d calls)."""
    reference_root = os.path.abspath(reference_root)
    if not os.path.isdir(os.path.join(reference_root, "layers")):
        raise FileNotFoundError(f"{reference_root} does not look like the reference tree (no layers/)")
    for p in (*extra_paths, reference_root):
        if p not in sys.path:
            sys.path.insert(0, p)
    if "layers" in sys.modules and not _is_reference_layers(sys.modules["layers"], reference_root):
        raise RuntimeError("a different top-level package named 'layers' is already imported")
    # A namespace for the reference's `layers` package whose __path__ still points at the reference tree
    # (so Masked_Attention.py, TransformRGB.py, SupplyMask.py ... load from there) ...
    if "layers" not in sys.modules:
        pkg = types.ModuleType("layers")
        pkg.__path__ = [os.path.join(reference_root, "layers")]
        pkg.__package__ = "layers"
        sys.modules["layers"] = pkg
    # ... but whose three hot-path submodules are ours.
    ours = importlib.import_module(__package__ + ".layers")
    for name in _DROPINS + (("Masked_Attention",) if fuse_gate else ()) + (("SupplyMask",) if fuse_pyramid else ()):
        mod = importlib.import_module(f"{__package__}.layers.{name}")
        sys.modules[f"layers.{name}"] = mod
        setattr(sys.modules["layers"], name, mod)
    sys.modules["layers"].__mwa_b200__ = ours


def _is_reference_layers(mod, root) -> bool:
    paths = [os.path.abspath(p) for p in getattr(mod, "__path__", [])]
    return os.path.join(root, "layers") in paths


def patch_model_rounding(model_module) -> None:
    """Replace the model file's module-level `ste_round` with the CUDA one (same forward value, same gradient)."""
    from .quant import ste_round
    if not hasattr(model_module, "ste_round"):
        raise AttributeError(f"{model_module.__name__} defines no ste_round")
    model_module.ste_round = ste_round


def accelerate_convs(model) -> int:
    """Switch the convolutions of an ALREADY BUILT model (the reference's AutoEncoder, built from its own files) to this
    package's drop-in classes, in place: every plain `torch.nn.Conv2d` / `ConvTranspose2d` becomes `layers.conv.Conv2d` /
    `ConvTranspose2d` and every plain `nn.Sequential` that holds convolutions becomes a `ConvStack` (activations folded
    into the convolutions' epilogues).  Parameters, buffers, hooks and state-dict keys are untouched -- only the class
    of the module objects changes.  In inference on CUDA the covered geometries then run csrc/conv_tc.cu, everything else
    (and every call that records autograd history) keeps going through torch.nn.functional.  Returns the number of
    convolutions switched."""
    import torch.nn as nn
    from .layers import conv as C
    n = 0
    for m in model.modules():
        if type(m) is nn.Conv2d:
            m.__class__ = C.Conv2d
            m._img = C._WeightImage()
            n += 1
        elif type(m) is nn.ConvTranspose2d:
            m.__class__ = C.ConvTranspose2d
            m._img = C._WeightImage()
            n += 1
    for m in model.modules():
        if type(m) is nn.Sequential and any(isinstance(c, (C.Conv2d, C.ConvTranspose2d)) or
                                            (type(c) is nn.Sequential and len(c) == 2 and isinstance(c[0], C.Conv2d))
                                            for c in m):
            if not (len(m) == 2 and isinstance(m[1], nn.PixelShuffle)):      # sub-pixel pairs stay plain: their parent fuses them
                m.__class__ = C.ConvStack
    return n


def uninstall() -> None:
    for name in ("layers",) + tuple(f"layers.{n}" for n in _DROPINS + ("Masked_Attention", "SupplyMask")):
        sys.modules.pop(name, None)
