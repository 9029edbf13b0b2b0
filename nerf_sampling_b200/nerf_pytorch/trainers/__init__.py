from . import Blender, Trainer  # noqa: F401
