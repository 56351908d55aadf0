"""Distance fields (parameter carriers).  Mirrors stoch_gpmp/costs/fields.py for the variants on the
StochGPMP hot path: LinkDistanceField with field_type='rbf' (costs/fields.py:30-79) and
LinkSelfDistanceField (costs/fields.py:89-127), both without link interpolation.  The arithmetic runs in
csrc/sgpmp_cost.cuh::link_fields."""


class DistanceField:
    def __init__(self, tensor_args=None):
        self.tensor_args = tensor_args

    def zero_grad(self):
        pass


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
        if self.num_interpolate:
            raise NotImplementedError("LinkDistanceField(num_interpolate>0) is not lowered yet (SURVEY §8f rank 1)")


class LinkSelfDistanceField(DistanceField):
    """sum over ALL ordered pairs of link frames (i == j included) of exp(-|p_i - p_j|^2 / (2 margin^2))
    (costs/fields.py:89-127).  Arithmetic: csrc/sgpmp_cost.cuh::link_fields."""

    def __init__(self, margin=0.03, num_interpolate=0, link_interpolate_range=(5, 7), **kwargs):
        super().__init__(**kwargs)
        self.margin = margin
        self.num_interpolate = num_interpolate
        self.link_interpolate_range = list(link_interpolate_range)

    def check_lowerable(self):
        if self.num_interpolate:
            raise NotImplementedError("LinkSelfDistanceField(num_interpolate>0) is not lowered yet (SURVEY §8f rank 1)")


class EESE3DistanceField(DistanceField):
    def __init__(self, *a, **k):
        raise NotImplementedError("EESE3DistanceField needs torch_robotics' SE3_distance (absent); SURVEY §8f rank 2")
