"""Drop-in for ``VLAAttacker/white_patch/UPA.py``: same class name and call signatures, CUDA engine inside."""
from ..attacker import UPAAttacker as OpenVLAAttacker  # noqa: F401
