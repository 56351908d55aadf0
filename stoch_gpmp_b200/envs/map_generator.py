"""Import-path mirror of stoch_gpmp/envs/map_generator.py (implementation: occupancy.py)."""
from .occupancy import generate_obstacle_map  # noqa: F401
