"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU restatement (torch fp32, CPU semantics) of the step that follows the sampling loop in the reference's
sample/generate.py:161-171 for data_rep == 'hml_vec':

    sample = inv_transform(sample.cpu().permute(0, 2, 3, 1)).float()      # data * std + mean   dataset.py:309-310
    sample = recover_from_ric(sample, n_joints)                           # motion_process.py:437-452
    sample = sample.view(-1, *sample.shape[2:]).permute(0, 2, 3, 1)      # [B, n_joints, 3, T]

Pinned by tests/golden/ric.npz (outputs of the reference's own functions, oracle/gen_golden.py:gen_ric).

HumanML3D feature row (motion_process.py:360-365): [0] root yaw velocity, [1:3] root XZ velocity (in the root frame),
[3] root height, [4 : 4 + 3(J-1)] root-relative joint positions, then rotations / velocities / foot contacts.
"""
import torch


def recover_root_rot_pos(data):
    """motion_process.py:366-385.  data [..., T, D] -> (cos, sin) of the accumulated yaw [..., T], root position [..., T, 3].
    torch.cumsum on CPU accumulates fp32 inputs in fp64 and rounds every prefix to fp32; kept as is."""
    rot_vel = data[..., 0]
    ang = torch.zeros_like(rot_vel)
    ang[..., 1:] = rot_vel[..., :-1]                       # :370  yaw at frame t = sum of the velocities before t
    ang = torch.cumsum(ang, dim=-1)                        # :371
    c, s = torch.cos(ang), torch.sin(ang)                  # quaternion (c, 0, s, 0), :373-375
    vel = torch.zeros(data.shape[:-1] + (3,), dtype=data.dtype)
    vel[..., 1:, [0, 2]] = data[..., :-1, 1:3]             # :377-378
    vel = _rot_y_inv(c, s, vel)                            # qrot(qinv(q), v), :380
    pos = torch.cumsum(vel, dim=-2)                        # :382
    pos[..., 1] = data[..., 3]                             # :384
    return c, s, pos


def _rot_y_inv(c, s, v):
    """qrot(qinv((c,0,s,0)), v) written out (quaternion.py:16-20,56-75): qvec = (0,-s,0), uv = qvec x v,
    uuv = qvec x uv, result = v + 2 (c uv + uuv), with the reference's operation order."""
    qy = -s
    vx, vy, vz = v[..., 0], v[..., 1], v[..., 2]
    uvx, uvz = qy * vz, -(qy * vx)                          # cross((0,qy,0), v) = (qy vz, 0, -qy vx)
    uuvx, uuvz = qy * uvz, -(qy * uvx)                      # cross((0,qy,0), uv)
    return torch.stack([vx + 2 * (c * uvx + uuvx), vy + torch.zeros_like(c), vz + 2 * (c * uvz + uuvz)], dim=-1)


def recover_from_ric(data, joints_num):
    """motion_process.py:437-452.  data [..., T, D] (already de-normalised) -> joint positions [..., T, joints_num, 3]."""
    c, s, r_pos = recover_root_rot_pos(data)
    pos = data[..., 4:(joints_num - 1) * 3 + 4]
    pos = pos.reshape(pos.shape[:-1] + (-1, 3))
    pos = _rot_y_inv(c[..., None], s[..., None], pos)      # :443
    pos = pos.clone()
    pos[..., 0] += r_pos[..., 0:1]                         # :446-447
    pos[..., 2] += r_pos[..., 2:3]
    return torch.cat([r_pos.unsqueeze(-2), pos], dim=-2)   # :450


def sample_to_xyz(sample, mean, std):
    """generate.py:161-166: sample [B, D, 1, T] (model output, normalised) -> [B, n_joints, 3, T]."""
    n_joints = 22 if sample.shape[1] == 263 else 21
    data = sample.cpu().permute(0, 2, 3, 1) * std + mean   # inv_transform, dataset.py:309-310
    xyz = recover_from_ric(data.float(), n_joints)
    return xyz.view(-1, *xyz.shape[2:]).permute(0, 2, 3, 1)
