"""2-D point robot: dof 2, state (x, y, vx, vy), one collision sphere at (x, y)
(API mirror of reference ``robot_models/point_robot_2d.py:5-71``).  Kinematics are the identity,
so the "sphere centre + FK Jacobian" step is a slice and a constant [I2 0] block; the CUDA
kernels read the position straight from the state (csrc/factors.cuh, assemble_node)."""
import torch

from .robot_model import RobotModel


class PointRobot2D(RobotModel):
    def __init__(self, sphere_radii, batch_size=1, num_traj_states=1, use_cuda=False):
        super(PointRobot2D, self).__init__(2, 1, 2, 4, sphere_radii, batch_size, num_traj_states, use_cuda)

    def forward_kinematics(self, pose_config, vel_config=None):
        return pose_config, vel_config, torch.eye(self.state_dim, device=pose_config.device, dtype=pose_config.dtype)

    def forward_kinematics_batch(self, pose_configb, vel_configb=None):
        B, T = pose_configb.shape[0], pose_configb.shape[1]
        J = torch.eye(self.state_dim, device=pose_configb.device, dtype=pose_configb.dtype).expand(B, T, -1, -1)
        vel = None if vel_configb is None else vel_configb.reshape(B, T, self.nlinks, self.wksp_dim)
        return pose_configb.reshape(B, T, self.nlinks, self.wksp_dim), vel, J

    def get_sphere_centers(self, state):
        J = torch.eye(self.state_dim, device=state.device, dtype=state.dtype)[0:self.nlinks * self.wksp_dim]
        return state[0:self.dofs].reshape(self.nlinks, self.wksp_dim), J

    def get_sphere_centers_full(self, traj):
        c, J = self.get_sphere_centers_batch(traj.unsqueeze(0))
        return c[0], J[0]

    def get_sphere_centers_batch(self, trajb):
        """trajb (B,T,4) -> centres (B,T,1,2), Jacobian (B,T,2,4) = [I2 0]."""
        B, T = trajb.shape[0], trajb.shape[1]
        J = torch.eye(self.state_dim, device=trajb.device, dtype=trajb.dtype)[0:self.nlinks * self.wksp_dim]
        return trajb[:, :, 0:self.dofs].reshape(B, T, self.nlinks, self.wksp_dim), J.expand(B, T, -1, -1)
