"""Oracle: SE(3) distance of the end-effector goal field.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED: the reference imports
`SE3_distance` from `torch_robotics.torch_kinematics_tree.geometrics.utils`
(stoch_gpmp/costs/fields.py:4, used at :142-144 by EESE3DistanceField.compute_distance), a
dependency that is neither vendored nor pinned nor installed, and no reference test holds a value of
it.  Restated here as the usual weighted sum of the translation distance and the geodesic rotation
angle,

    SE3_distance(H1, H2) = w_pos |p1 - p2| + w_rot acos( clamp( (tr(R1^T R2) - 1) / 2, -1, 1 ) ),

which is what the call site's keyword arguments (w_pos, w_rot) and EESE3DistanceField's `square`
option (fields.py:147-150) imply.  The golden Panda cases that include the EE-goal term inject THIS
definition into the unmodified reference through the import stub (oracle/ref_loader.py), so
everything around it (CostGoal, FieldFactor slicing of the last step, squaring, 1/sigma_goal^2
weighting, summation order) is pinned by executing the reference.
"""
import numpy as np


def se3_distance(H1, H2, w_pos=1., w_rot=1.):
    """H1 [..., 4, 4], H2 broadcastable -> [...]."""
    H1 = np.asarray(H1)
    H2 = np.asarray(H2)
    dp = np.linalg.norm(H1[..., :3, 3] - H2[..., :3, 3], axis=-1)
    tr = (H1[..., :3, :3] * H2[..., :3, :3]).sum((-1, -2))        # tr(R1^T R2) = sum_ij R1_ij R2_ij
    ang = np.arccos(np.clip((tr - 1.0) * 0.5, -1.0, 1.0))
    return w_pos * dp + w_rot * ang


def se3_distance_torch(H1, H2, w_pos=1., w_rot=1.):
    """Same definition on torch tensors (the stub bound to the reference's import)."""
    import torch
    dp = torch.linalg.norm(H1[..., :3, 3] - H2[..., :3, 3], dim=-1)
    tr = (H1[..., :3, :3] * H2[..., :3, :3]).sum((-1, -2))
    ang = torch.acos(torch.clamp((tr - 1.0) * 0.5, -1.0, 1.0))
    return w_pos * dp + w_rot * ang
