"""hydra_pspec_b200 -- B200-native (sm_100a) implementation of hydra-pspec's per-baseline Gibbs
sampling hot path, behind the reference's Python API (``hydra_pspec.pspec``).

Only the hot path lives here: ``pspec`` (the drop-in functions and the batched
:class:`~hydra_pspec_b200.pspec.GibbsEngine`) and the few ``utils`` helpers that path uses.
There is no CPU fallback: importing ``pspec`` works anywhere, calling it needs the compiled
``csrc/libhydra_pspec_b200.so`` and a CUDA device.
"""
__version__ = "0.1.0"

from . import utils, pspec  # noqa: E402,F401
