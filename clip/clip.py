"""Import-compatible with the reference's `clip/clip.py`: `from clip.clip import CLIP`."""
from conzic_b200.clip.clip import CLIP  # noqa: F401
