"""Import-path mirror of stoch_gpmp/envs/obst_map.py (implementation: occupancy.py)."""
from .occupancy import ObstacleMap, ObstacleRectangle, ObstacleCircle  # noqa: F401
