"""Oracle: serial-chain forward kinematics (Franka Panda, no gripper).

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED: the reference calls
`torch_robotics ... DifferentiableFrankaPanda(gripper=False).compute_forward_kinematics_all_links`
(examples/panda_environment.py:47,98; hook at stoch_gpmp/costs/cost_functions.py:51-52), a
dependency that is neither vendored nor pinned nor installed.  This file restates standard URDF
semantics ( child = parent * T(xyz) * R(rpy) * Rz(q) for a revolute z-axis joint ) over
assets/franka_description/robots/panda_arm_no_gripper.urdf, and is pinned by known answers only
(tests/test_oracle_golden.py::test_fk_known_answers).

Chain rows: (name, xyz, rpy, joint index or -1), in URDF order; the frame list returned is
[panda_link0, link1..link7, link8, panda_hand, ee_link] (L = 11), EE last
(stoch_gpmp/costs/fields.py:143-144 takes the last frame as the end-effector).
"""
import numpy as np

# panda_arm_no_gripper.urdf lines 43, 68, 93, 118, 143, 168, 193 (revolute, axis 0 0 1),
# 201 (joint8, fixed), 209 (hand joint, fixed), 235 (ee_fixed_joint).
PANDA_CHAIN = [
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
]


def rpy_matrix(rpy):
    """URDF fixed-axis roll-pitch-yaw: R = Rz(yaw) Ry(pitch) Rx(roll)."""
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def chain_tables(chain=PANDA_CHAIN):
    """(R_fix [F,3,3], p_fix [F,3], joint [F]) float64 — the descriptor the CUDA kernel consumes."""
    R = np.stack([rpy_matrix(c[2]) for c in chain])
    p = np.array([c[1] for c in chain], dtype=np.float64)
    j = np.array([c[3] for c in chain], dtype=np.int32)
    return R, p, j


def fk_all_links(q, chain=PANDA_CHAIN, include_base=True, dtype=None):
    """q [N, n] -> H [N, L, 4, 4] world transforms of every link frame."""
    q = np.asarray(q)
    dtype = dtype or q.dtype
    N = q.shape[0]
    Rf, pf, jn = chain_tables(chain)
    Rf = Rf.astype(dtype)
    pf = pf.astype(dtype)
    R = np.broadcast_to(np.eye(3, dtype=dtype), (N, 3, 3)).copy()
    p = np.zeros((N, 3), dtype=dtype)
    frames = []
    if include_base:
        H = np.zeros((N, 4, 4), dtype=dtype)
        H[:, :3, :3] = R
        H[:, 3, 3] = 1
        frames.append(H)
    for f in range(len(chain)):
        p = p + np.einsum('nij,j->ni', R, pf[f])
        R = R @ Rf[f]
        if jn[f] >= 0:
            c = np.cos(q[:, jn[f]]).astype(dtype)
            s = np.sin(q[:, jn[f]]).astype(dtype)
            Rz = np.zeros((N, 3, 3), dtype=dtype)
            Rz[:, 0, 0] = c
            Rz[:, 0, 1] = -s
            Rz[:, 1, 0] = s
            Rz[:, 1, 1] = c
            Rz[:, 2, 2] = 1
            R = R @ Rz
        H = np.zeros((N, 4, 4), dtype=dtype)
        H[:, :3, :3] = R
        H[:, :3, 3] = p
        H[:, 3, 3] = 1
        frames.append(H)
    return np.stack(frames, axis=1)


def fk_all_links_torch(chain=PANDA_CHAIN, include_base=True):
    """Returns a torch callable q [N,n] -> [N,L,4,4] with the same semantics, for injection
    through the reference's own hook CostComposite(FK=...) (cost_functions.py:39-44,51-52)."""
    import torch
    Rf_np, pf_np, jn = chain_tables(chain)

    def fk(q):
        N = q.shape[0]
        kw = dict(dtype=q.dtype, device=q.device)
        Rf = torch.as_tensor(Rf_np, **kw)
        pf = torch.as_tensor(pf_np, **kw)
        R = torch.eye(3, **kw).expand(N, 3, 3)
        p = torch.zeros(N, 3, **kw)
        frames = []

        def pack(R, p):
            H = torch.zeros(N, 4, 4, **kw)
            H[:, :3, :3] = R
            H[:, :3, 3] = p
            H[:, 3, 3] = 1
            return H
        if include_base:
            frames.append(pack(R, p))
        for f in range(len(chain)):
            p = p + (R @ pf[f])
            R = R @ Rf[f]
            if jn[f] >= 0:
                c = torch.cos(q[:, jn[f]])
                s = torch.sin(q[:, jn[f]])
                Rz = torch.zeros(N, 3, 3, **kw)
                Rz[:, 0, 0] = c
                Rz[:, 0, 1] = -s
                Rz[:, 1, 0] = s
                Rz[:, 1, 1] = c
                Rz[:, 2, 2] = 1
                R = R @ Rz
            frames.append(pack(R, p))
        return torch.stack(frames, dim=1)
    return fk
