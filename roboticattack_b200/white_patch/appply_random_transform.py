"""Drop-in for ``VLAAttacker/white_patch/appply_random_transform.py`` (the reference's spelling is kept)."""
from ..frontend import RandomPatchTransform  # noqa: F401
