"""Re-export: the generators live in the package (dagsfm_b200/tv_scene.py)."""
from dagsfm_b200.tv_scene import make_pairs, scene  # noqa: F401
