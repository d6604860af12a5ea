"""Drop-in for ``VLAAttacker/white_patch/UADA.py``: same class name and call signatures, CUDA engine inside."""
from ..attacker import UADAAttacker as OpenVLAAttacker  # noqa: F401
