"""Drop-in for ``VLAAttacker/white_patch/TMA.py``: same class name and call signatures, CUDA engine inside."""
from ..attacker import TMAAttacker as OpenVLAAttacker  # noqa: F401
