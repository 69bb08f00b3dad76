"""Minimal 2-D environment holder (import-compatible with reference ``diff_gpmp2/env/env_2d.py``).

Only what callers on either side of the GN path need: limits, image / SDF storage, resolution.
The reference's matplotlib plotting and its legacy single-point SDF query are out of scope;
plotting methods import matplotlib lazily and raise a clear error if it is absent."""
import numpy as np
import torch

from ..utils import sdf_utils


class Env2D():
    def __init__(self, params, use_cuda=False):
        self.plot_initialized = False
        self.image = None
        self.sedt = None
        self.sedt_available = False
        self.ndims = 2
        self.use_cuda = torch.cuda.is_available() if use_cuda else False
        self.device = torch.device('cuda') if self.use_cuda else torch.device('cpu')
        self.x_lims = params['x_lims']
        self.y_lims = params['y_lims']

    def _set_geometry(self):
        self.res = (self.x_lims[1] - self.x_lims[0]) / ((self.image.shape[1]) * 1.)
        self.orig_pix = torch.tensor([0 - self.x_lims[0] / self.res, 0 - self.y_lims[0] / self.res], device=self.device)
        self.MAX_D = (self.x_lims[1] - self.x_lims[0])

    def initialize_from_file(self, envfile):
        from PIL import Image
        self.image = np.asarray(Image.open(envfile).convert('L'), dtype=np.float64) / 255.0
        self._set_geometry()
        self.calculate_signed_distance_transform()

    def initialize_from_image(self, img, sedt=None):
        self.image = sdf_utils.rgb2gray(img) if len(img.shape) > 2 else img
        self._set_geometry()
        self.sedt = torch.as_tensor(sedt, device=self.device) if sedt is not None else None
        self.sedt_available = sedt is not None

    def calculate_signed_distance_transform(self, pad_len=1):
        if not self.sedt_available:
            self.sedt = torch.tensor(sdf_utils.sdf_2d(np.asarray(self.image), pad_len, self.res), device=self.device)
            self.sedt_available = True

    def in_limits(self, state):
        return bool(self.x_lims[0] <= state[0] < self.x_lims[1] and self.y_lims[0] <= state[1] < self.y_lims[1])

    def get_signed_obstacle_distance(self, stateb):
        """Batched SDF lookup (dist, J) through the CUDA library; stateb (N,1,2) or (B,N,2)."""
        pts = stateb.reshape(1, -1, 2) if stateb.dim() == 3 and stateb.shape[1] == 1 else stateb
        sdf = self.sedt.reshape(1, *self.sedt.shape[-2:])
        res = (self.x_lims[1] - self.x_lims[0]) / sdf.shape[-1]
        d, J = sdf_utils.bilinear_interpolate(sdf, pts, res, self.x_lims, self.y_lims)
        return d.reshape(-1, 1), J.reshape(-1, 1, 2)

    def _plt(self):
        try:
            import matplotlib.pyplot as plt
            return plt
        except ImportError as e:
            raise ImportError('Env2D plotting needs matplotlib, which is not installed') from e

    def initialize_plot(self, start, goal, grid_res=None, plot_grid=False):
        plt = self._plt()
        self.figure, self.axes = plt.subplots()
        self.axes.imshow(np.asarray(self.image), extent=(self.x_lims[0], self.x_lims[1], self.y_lims[0], self.y_lims[1]), cmap='gray')
        self.axes.plot(float(start[0]), float(start[1]), 'go')
        self.axes.plot(float(goal[0]), float(goal[1]), 'ro')
        self.plot_initialized = True

    def plot_edge(self, edge, linestyle='-', linewidth=1.0, color='blue', alpha=1.0, markerstyle='o', markersize=4.0, label=None):
        xs, ys = [float(p[0]) for p in edge], [float(p[1]) for p in edge]
        self.axes.plot(xs, ys, linestyle=linestyle, linewidth=linewidth, color=color, alpha=alpha, label=label)

    def plot_signed_distance_transform(self):
        self._plt()
        self.axes.imshow(np.asarray(self.sedt.cpu()), extent=(self.x_lims[0], self.x_lims[1], self.y_lims[0], self.y_lims[1]), alpha=0.5)

    def close_plot(self):
        if self.plot_initialized:
            self._plt().close(self.figure)
            self.plot_initialized = False
