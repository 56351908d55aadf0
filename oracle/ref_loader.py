"""Import the REAL reference (anindex/stoch_gpmp) for golden-vector generation.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Build-container only: `/root/reference`
does not exist on the GPU box, and nothing in `-m gpu` tests, smoke() or bench.py calls this.

Two import stubs are needed:
  * matplotlib                — stoch_gpmp/envs/obst_map.py:4 imports pyplot only for plot()
  * torch_robotics...SE3_distance — stoch_gpmp/costs/fields.py:4, only used by EESE3DistanceField; bound to the
    restatement in oracle/se3.py (parity unpinned at that function)
"""
import os
import sys
import types

SEARCH = [os.environ.get("STOCH_GPMP_REF"), "/root/reference"]


def reference_root():
    for p in SEARCH:
        if p and os.path.isdir(os.path.join(p, "stoch_gpmp")):
            return p
    return None


def available():
    return reference_root() is not None


def load():
    """Returns a namespace with the reference's classes; raises RuntimeError if absent."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not found (searched $STOCH_GPMP_REF, /root/reference)")
    sys.dont_write_bytecode = True            # the reference mount is read-only
    if root not in sys.path:
        sys.path.insert(0, root)
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt
    tr = "torch_robotics.torch_kinematics_tree.geometrics.utils"
    if tr not in sys.modules:
        parts = tr.split(".")
        for i in range(1, len(parts) + 1):
            nm = ".".join(parts[:i])
            sys.modules.setdefault(nm, types.ModuleType(nm))
        # the absent dependency's function, restated (PARITY UNPINNED, see oracle/se3.py)
        from .se3 import se3_distance_torch
        sys.modules[tr].SE3_distance = se3_distance_torch
    ns = types.SimpleNamespace()
    from stoch_gpmp.planner import StochGPMP
    from stoch_gpmp.costs.cost_functions import CostCollision, CostComposite, CostGP, CostGoalPrior, CostGoal, CostGPTrajectory
    from stoch_gpmp.costs.fields import LinkDistanceField, LinkSelfDistanceField, EESE3DistanceField
    from stoch_gpmp.costs.factors.mp_priors_multi import MultiMPPrior
    from stoch_gpmp.envs.map_generator import generate_obstacle_map
    from stoch_gpmp.envs.obst_map import ObstacleMap, ObstacleRectangle, ObstacleCircle
    ns.StochGPMP = StochGPMP
    ns.CostCollision, ns.CostComposite, ns.CostGP, ns.CostGoalPrior = CostCollision, CostComposite, CostGP, CostGoalPrior
    ns.LinkDistanceField = LinkDistanceField
    ns.LinkSelfDistanceField, ns.EESE3DistanceField, ns.CostGoal = LinkSelfDistanceField, EESE3DistanceField, CostGoal
    ns.CostGPTrajectory = CostGPTrajectory
    ns.MultiMPPrior = MultiMPPrior
    ns.generate_obstacle_map = generate_obstacle_map
    ns.ObstacleMap, ns.ObstacleRectangle, ns.ObstacleCircle = ObstacleMap, ObstacleRectangle, ObstacleCircle
    ns.root = root
    return ns
