"""Distance fields (parameter carriers).  Mirrors stoch_gpmp/costs/fields.py: LinkDistanceField
(costs/fields.py:30-86: 'rbf' / 'sdf' / 'occupancy', link interpolation), LinkSelfDistanceField
(costs/fields.py:89-127) and EESE3DistanceField (costs/fields.py:130-153).  The arithmetic runs in
csrc/sgpmp_cost.cuh (link_fields, ee_se3_cost)."""
import torch


def _interp_alpha(num_interpolate):
    """alpha of costs/fields.py:69, formed exactly like the reference does: torch.linspace in the default dtype
    (float32), then cast — the kernels receive these values as doubles (sgpmp_cost_desc_t.*_interp_alpha)."""
    n = int(num_interpolate)
    return [float(a) for a in torch.linspace(0, 1, n + 2)[1:n + 1].to(torch.float32)]


def _check_interp(name, num_interpolate, rng):
    from .. import _lib
    n = int(num_interpolate)
    if n < 0 or n > _lib.MAX_INTERP:
        raise NotImplementedError("%s(num_interpolate=%d): the CUDA path supports 0..%d" % (name, n, _lib.MAX_INTERP))
    if n and (len(rng) != 2 or rng[0] < 0 or rng[1] < rng[0]):
        raise ValueError("%s: link_interpolate_range must be [lo, hi] with 0 <= lo <= hi, got %r" % (name, rng))


class DistanceField:
    """The arithmetic of a field normally runs inside the CUDA kernels.  `compute_cost` is the torch form with the
    reference's signature (link frames [..., L, 4, 4] -> one value per configuration): the planner uses it only for a
    term the kernels CANNOT lower — a `CostComposite(FK=<any callable>)` that is not a `SerialChainFK`, as the reference
    allows (cost_functions.py:39-52) — evaluated on the materialised samples like a user-defined term."""

    def __init__(self, tensor_args=None):
        self.tensor_args = tensor_args

    def zero_grad(self):
        pass


def _link_points(link_tensor, num_interpolate, rng):
    """Link-frame origins [..., L, 3] plus the interpolated points of costs/fields.py:68-74 appended link by link."""
    pts = link_tensor[..., :3, 3]
    n = int(num_interpolate)
    if n > 0:
        alpha = torch.linspace(0, 1, n + 2)[1:n + 1].to(torch.float32).to(pts.dtype).to(pts.device).reshape(n, 1)
        for i in range(int(rng[0]), int(rng[1])):
            a, b = pts[..., i:i + 1, :], pts[..., i + 1:i + 2, :]
            pts = torch.cat([pts, a + (b - a) * alpha], dim=-2)
    return pts


class LinkDistanceField(DistanceField):
    """'rbf': sum over link frames and obstacle spheres of exp(-0.5 |p - c|^2 / r^2); 'sdf': max of r - |p - c|
    (clamped to <= 0 with clamp_sdf); 'occupancy': number of (link, sphere) pairs in contact."""

    def __init__(self, field_type='rbf', clamp_sdf=False, num_interpolate=0, link_interpolate_range=(5, 7), **kwargs):
        super().__init__(**kwargs)
        self.field_type = field_type
        self.clamp_sdf = clamp_sdf
        self.num_interpolate = num_interpolate
        self.link_interpolate_range = list(link_interpolate_range)

    def field_code(self):
        """sgpmp_cost_desc_t.sphere_field_type for this field (costs/fields.py:78-86)."""
        from .. import _lib
        if self.field_type == 'rbf':
            return _lib.FIELD_RBF
        if self.field_type == 'sdf':
            return _lib.FIELD_SDF_CLAMPED if self.clamp_sdf else _lib.FIELD_SDF
        if self.field_type == 'occupancy':
            return _lib.FIELD_OCCUPANCY
        raise NotImplementedError("LinkDistanceField(field_type=%r) is unknown" % (self.field_type,))

    def check_lowerable(self):
        self.field_code()
        _check_interp("LinkDistanceField", self.num_interpolate, self.link_interpolate_range)

    def interp_alpha(self):
        return _interp_alpha(self.num_interpolate)

    def compute_cost(self, link_tensor, obstacle_spheres=None, **kwargs):
        """costs/fields.py:63-86 in torch (see DistanceField): obstacle_spheres [O, 4] or [1, O, 4] = (centre, radius)."""
        if obstacle_spheres is None:
            return 0
        sp = torch.as_tensor(obstacle_spheres).to(link_tensor).reshape(-1, 4)
        pts = _link_points(link_tensor, self.num_interpolate, self.link_interpolate_range).unsqueeze(-2)     # [..., P, 1, 3]
        d2 = ((pts - sp[:, :3]) ** 2).sum(-1)                                                                # [..., P, O]
        if self.field_type == 'rbf':
            return torch.exp(-0.5 * d2 / sp[:, 3] ** 2).sum((-1, -2))
        if self.field_type == 'sdf':
            sdf = sp[:, 3] - d2.sqrt()
            if self.clamp_sdf:
                sdf = sdf.clamp(max=0.)
            return sdf.amax((-1, -2))
        if self.field_type == 'occupancy':
            return (d2.sqrt() < sp[:, 3]).sum((-1, -2)).to(link_tensor.dtype)
        raise NotImplementedError("LinkDistanceField(field_type=%r) is unknown" % (self.field_type,))


class LinkSelfDistanceField(DistanceField):
    """sum over ALL ordered pairs of link frames (i == j included) of exp(-|p_i - p_j|^2 / (2 margin^2))
    (costs/fields.py:89-127).  Arithmetic: csrc/sgpmp_cost.cuh::link_fields."""

    def __init__(self, margin=0.03, num_interpolate=0, link_interpolate_range=(5, 7), **kwargs):
        super().__init__(**kwargs)
        self.margin = margin
        self.num_interpolate = num_interpolate
        self.link_interpolate_range = list(link_interpolate_range)

    def check_lowerable(self):
        _check_interp("LinkSelfDistanceField", self.num_interpolate, self.link_interpolate_range)

    def interp_alpha(self):
        return _interp_alpha(self.num_interpolate)

    def compute_cost(self, link_tensor, **kwargs):
        """costs/fields.py:114-124 in torch (see DistanceField): all ordered pairs of link points, i == j included."""
        pts = _link_points(link_tensor, self.num_interpolate, self.link_interpolate_range)
        d2 = ((pts.unsqueeze(-2) - pts.unsqueeze(-3)) ** 2).sum(-1)
        return torch.exp(-d2 / (2.0 * self.margin ** 2)).sum((-1, -2))


class EESE3DistanceField(DistanceField):
    """SE(3) distance of the end-effector (LAST link frame, costs/fields.py:142-144) to `target_H` [4,4] (or [1,4,4]),
    squared when `square` (costs/fields.py:146-150).  The reference takes `SE3_distance` from the absent torch_robotics;
    the CUDA path evaluates  w_pos |p - p*| + w_rot acos(clamp((tr(R^T R*) - 1)/2, -1, 1))  (csrc ee_se3_cost,
    oracle/se3.py; parity unpinned at that one function)."""

    def __init__(self, target_H, w_pos=1., w_rot=1., square=True, **kwargs):
        super().__init__(**kwargs)
        self.target_H = target_H
        self.square = square
        self.w_pos = w_pos
        self.w_rot = w_rot

    def update_target(self, target_H):
        self.target_H = target_H

    def target_matrix(self):
        """The target as a [4,4] float64 CPU tensor (validated)."""
        H = torch.as_tensor(self.target_H).detach().to(device='cpu', dtype=torch.float64).reshape(-1, 4, 4)
        if H.shape[0] != 1:
            raise NotImplementedError("EESE3DistanceField: one target pose per cost (got %d)" % H.shape[0])
        return H[0]

    def compute_cost(self, link_tensor, **kwargs):
        """costs/fields.py:142-150 in torch (see DistanceField), with the SE(3) metric the CUDA path uses (class docstring)."""
        H = link_tensor[..., -1, :, :]
        Ht = torch.as_tensor(self.target_H).to(H).reshape(-1, 4, 4)[0]
        dp = torch.linalg.norm(H[..., :3, 3] - Ht[:3, 3], dim=-1)
        tr = (H[..., :3, :3] * Ht[:3, :3]).sum((-1, -2))
        dist = self.w_pos * dp + self.w_rot * torch.acos(torch.clamp((tr - 1.0) * 0.5, -1.0, 1.0))
        return dist * dist if self.square else dist

    def check_lowerable(self):
        self.target_matrix()
        if not (self.w_pos >= 0 and self.w_rot >= 0):
            raise ValueError("EESE3DistanceField: w_pos and w_rot must be >= 0")
