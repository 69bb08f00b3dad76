"""Planar robot with heading: dof 3, state (x, y, h, vx, vy, w), one collision sphere at (x, y)
(API mirror of reference ``robot_models/point_robot_xyh.py:5-61``).  The reference has no batch
method for this robot (its obstacle factor therefore cannot run batched); this mirror adds
``get_sphere_centers_batch`` with the Jacobian rows the reference's ``forward_kinematics_full``
uses (:28-35)."""
import torch

from .robot_model import RobotModel


class PointRobotXYH(RobotModel):
    def __init__(self, sphere_radii, use_cuda=False, batch_size=1, num_traj_states=1):
        super(PointRobotXYH, self).__init__(3, 1, 2, 6, sphere_radii, batch_size, num_traj_states, use_cuda)

    def _jac(self, device, dtype):
        J = torch.zeros(2, self.state_dim, device=device, dtype=dtype)
        J[0, 0] = 1.0
        J[1, 1] = 1.0
        return J

    def forward_kinematics(self, pose_config, vel_config=None):
        return pose_config, vel_config, torch.eye(self.state_dim, device=pose_config.device, dtype=pose_config.dtype)

    def get_sphere_centers(self, state):
        return state[0:2].reshape(self.nlinks, self.wksp_dim), self._jac(state.device, state.dtype)

    def get_sphere_centers_full(self, traj):
        c, J = self.get_sphere_centers_batch(traj.unsqueeze(0))
        return c[0], J[0]

    def get_sphere_centers_batch(self, trajb):
        B, T = trajb.shape[0], trajb.shape[1]
        return trajb[:, :, 0:2].reshape(B, T, self.nlinks, self.wksp_dim), self._jac(trajb.device, trajb.dtype).expand(B, T, -1, -1)
