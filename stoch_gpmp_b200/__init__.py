"""stoch_gpmp_b200 — B200-native StochGPMP optimisation loop behind the reference's Python API.

    from stoch_gpmp_b200.planner import StochGPMP, StochGPMPBatch, GPMP, GPMPBatch, print_info
    from stoch_gpmp_b200.costs.cost_functions import CostGP, CostGPTrajectory, CostGoalPrior, CostCollision, CostGoal, CostComposite
    from stoch_gpmp_b200.costs.fields import LinkDistanceField, LinkSelfDistanceField, EESE3DistanceField
    from stoch_gpmp_b200.parallel import shard_range, StreamShards
    from stoch_gpmp_b200.envs.map_generator import generate_obstacle_map
    from stoch_gpmp_b200.robots import PandaFK

The arithmetic lives in hand-written sm_100a CUDA kernels (csrc/, C-ABI in include/stoch_gpmp_b200.h);
build them with `python -m stoch_gpmp_b200.build`.  There is no CPU fallback.
"""
__version__ = "0.1.0"
