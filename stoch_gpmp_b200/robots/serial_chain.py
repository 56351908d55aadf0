"""Forward-kinematics descriptors consumed by the fused cost kernel.

The reference hands `CostComposite` an FK *callable* from torch_robotics
(`DifferentiableFrankaPanda(gripper=False).compute_forward_kinematics_all_links`,
examples/panda_environment.py:47,98; hook at costs/cost_functions.py:39-52).  torch_robotics is an
un-vendored dependency, so the chain is restated here from the reference's own
assets/franka_description/robots/panda_arm_no_gripper.urdf (joint origins at lines 43, 68, 93, 118, 143,
168, 193; fixed joints at 201, 209, 235).  A `SerialChainFK` does not compute anything on the host: it is
the parameter block (fixed transforms + joint order) that `StochGPMP` lowers into the CUDA kernel.
"""
import math
import xml.etree.ElementTree as ET


def _rpy_matrix(rpy):
    """URDF fixed-axis roll-pitch-yaw: R = Rz(yaw) Ry(pitch) Rx(roll), row-major 9-list."""
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr,
            sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr,
            -sp, cp * sr, cp * cr]


# (child link, xyz, rpy, joint index or -1)
PANDA_CHAIN = (
    ("panda_link1", (0.0, 0.0, 0.333), (0.0, 0.0, 0.0), 0),
    ("panda_link2", (0.0, 0.0, 0.0), (-1.57079632679, 0.0, 0.0), 1),
    ("panda_link3", (0.0, -0.316, 0.0), (1.57079632679, 0.0, 0.0), 2),
    ("panda_link4", (0.0825, 0.0, 0.0), (1.57079632679, 0.0, 0.0), 3),
    ("panda_link5", (-0.0825, 0.384, 0.0), (-1.57079632679, 0.0, 0.0), 4),
    ("panda_link6", (0.0, 0.0, 0.0), (1.57079632679, 0.0, 0.0), 5),
    ("panda_link7", (0.088, 0.0, 0.0), (1.57079632679, 0.0, 0.0), 6),
    ("panda_link8", (0.0, 0.0, 0.107), (0.0, 0.0, 0.0), -1),
    ("panda_hand", (0.0, 0.0, 0.0), (0.0, 0.0, -0.785398163397), -1),
    ("ee_link", (0.0, 0.0, 0.1), (0.0, 0.0, -1.57), -1),
)


class SerialChainFK:
    """Serial arm: frames 0..n_dof-1 are revolute about their own z axis (joint f), the rest are fixed.
    `include_base` adds the base frame (identity) to the list of link frames, as
    compute_forward_kinematics_all_links does for panda_link0."""

    def __init__(self, chain, include_base=True):
        self.names = [c[0] for c in chain]
        self.xyz = [tuple(float(v) for v in c[1]) for c in chain]
        self.R = [_rpy_matrix(c[2]) for c in chain]
        self.joint = [int(c[3]) for c in chain]
        self.include_base = bool(include_base)
        self.n_dofs = sum(1 for j in self.joint if j >= 0)
        for f, j in enumerate(self.joint):
            if j != (f if f < self.n_dofs else -1):
                raise NotImplementedError("SerialChainFK: joints must come first and in order (frame %d has joint %d)" % (f, j))

    def __call__(self, q):
        """Link frames as 4x4 transforms, [..., n_dofs] -> [..., num_links, 4, 4], in torch on q's device: the calling
        convention of the reference's FK hook (cost_functions.py:51-52).  Only user-defined cost terms go through this (they
        receive `x_trajs` like in the reference); the CUDA kernels evaluate the chain themselves from the descriptor."""
        import torch
        lead = q.shape[:-1]
        q = q.reshape(-1, q.shape[-1])
        N = q.shape[0]
        kw = dict(dtype=q.dtype, device=q.device)
        H = torch.eye(4, **kw).expand(N, 4, 4)
        frames = [H] if self.include_base else []
        for f, j in enumerate(self.joint):
            F = torch.eye(4, **kw)
            F[:3, :3] = torch.tensor(self.R[f], **kw).reshape(3, 3)
            F[:3, 3] = torch.tensor(self.xyz[f], **kw)
            H = H @ F
            if j >= 0:
                c, s_ = torch.cos(q[:, j]), torch.sin(q[:, j])
                Rz = torch.zeros(N, 4, 4, **kw)
                Rz[:, 0, 0], Rz[:, 0, 1], Rz[:, 1, 0], Rz[:, 1, 1] = c, -s_, s_, c
                Rz[:, 2, 2] = 1.0
                Rz[:, 3, 3] = 1.0
                H = H @ Rz
            frames.append(H)
        return torch.stack(frames, dim=1).reshape(*lead, len(frames), 4, 4)

    @property
    def num_links(self):
        return len(self.joint) + (1 if self.include_base else 0)

    @classmethod
    def from_urdf(cls, path, base_link, tip_link, include_base=True):
        """Build the chain base_link -> tip_link from a URDF (revolute z-axis and fixed joints only)."""
        root = ET.parse(path).getroot()
        by_child = {}
        for j in root.findall("joint"):
            by_child[j.find("child").get("link")] = j
        rows, link = [], tip_link
        while link != base_link:
            j = by_child[link]
            o = j.find("origin")
            xyz = tuple(float(v) for v in (o.get("xyz", "0 0 0") if o is not None else "0 0 0").split())
            rpy = tuple(float(v) for v in (o.get("rpy", "0 0 0") if o is not None else "0 0 0").split())
            jt = j.get("type")
            if jt == "revolute":
                axis = tuple(float(v) for v in j.find("axis").get("xyz").split())
                if axis != (0.0, 0.0, 1.0):
                    raise NotImplementedError("only z-axis revolute joints are supported (joint %s)" % j.get("name"))
            elif jt != "fixed":
                raise NotImplementedError("joint type %s is not supported" % jt)
            rows.append((link, xyz, rpy, jt == "revolute"))
            link = j.find("parent").get("link")
        rows.reverse()
        k, chain = 0, []
        for name, xyz, rpy, rev in rows:
            chain.append((name, xyz, rpy, k if rev else -1))
            k += int(rev)
        return cls(chain, include_base=include_base)


class PandaFK(SerialChainFK):
    """Franka Panda without gripper: 7 revolute joints + link8, hand, ee_link; 11 link frames."""

    def __init__(self, include_base=True):
        super().__init__(PANDA_CHAIN, include_base=include_base)
        self._n_dofs = self.n_dofs
