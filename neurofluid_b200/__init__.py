"""neurofluid_b200 -- B200-native (sm_100a) implementation of NeuroFluid's two hot paths behind the
reference's own module API: `RenderNet` (models/renderer.py) and `ParticleNet` (models/transmodel.py)."""
from .renderer import RenderNet, Renderer  # noqa: F401
from .nerf import NeRF, Embedding  # noqa: F401

__all__ = ["RenderNet", "Renderer", "NeRF", "Embedding"]
try:  # the transition model lands after the renderer
    from .transmodel import ParticleNet, TransModel  # noqa: F401
    __all__ += ["ParticleNet", "TransModel"]
except ImportError:  # pragma: no cover
    pass
