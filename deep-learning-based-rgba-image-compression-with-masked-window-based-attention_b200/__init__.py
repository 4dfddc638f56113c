"""B200-native (sm_100a) hot path of the RGBA codec with masked window-based attention.

Drop-in replacements for the reference's `layers/masked_win_attention.py`,
`layers/win_attention.py`, `layers/GDN.py` and the `ste_round` helpers of its model files,
implemented as hand-written CUDA kernels behind the C ABI of include/mwa_b200.h.

    from <package>.layers.GDN import GDN
    from <package>.layers.masked_win_attention import WinBasedAttention
    from <package>.quant import ste_round
    <package>.install.install()      # make the reference's own models/ pick these modules up
"""
from . import _abi  # noqa: F401

__all__ = ["_abi"]
__version__ = "0.1.0"
