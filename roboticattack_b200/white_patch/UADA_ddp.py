"""Drop-in for ``VLAAttacker/white_patch/UADA_ddp.py``: same class name and call signatures, CUDA engine inside."""
from ..attacker import UADADDPAttacker as OpenVLAAttacker  # noqa: F401
