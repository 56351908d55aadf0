"""Planar occupancy grids: the obstacle field of the planar example.

Host-side fixture code (numpy) + the parameter block the CUDA kernel consumes.  Same constructor
arguments, attributes and rasterisation semantics as the reference so that the same seeds give the same
grids (checked against the golden maps in tests/golden):
  ObstacleMap                         envs/obst_map.py:108-188
  ObstacleRectangle / ObstacleCircle  envs/obst_map.py:41-105
  generate_obstacle_map               envs/map_generator.py:9-92 (+ envs/obst_utils.py:12-26)
The lookup itself (ObstacleMap.get_collisions, envs/obst_map.py:164-182) runs inside the cost kernel
(csrc/sgpmp_cost.cuh::map_value); `compute_cost` here calls that kernel.
"""
import math
import random

import numpy as np
import torch


class ObstacleMap:
    """Occupancy-count grid centred on the origin; cell (iy, ix) covers [ix - ox, ix - ox + 1) * cell_size in x."""

    def __init__(self, map_dim, cell_size, tensor_args=None):
        if map_dim[0] % 2 or map_dim[1] % 2:
            raise AssertionError("map dimensions must be even")
        self.tensor_args = tensor_args or {'device': torch.device('cpu'), 'dtype': torch.float32}
        nx, ny = math.ceil(map_dim[0] / cell_size), math.ceil(map_dim[1] / cell_size)
        self.map = np.zeros([nx, ny])
        self.cell_size = cell_size
        self.origin_xi, self.origin_yi = int(nx / 2), int(ny / 2)
        self.x_dim, self.y_dim = self.map.shape
        self.xlim = [-cell_size * self.x_dim / 2, cell_size * self.x_dim / 2]
        self.ylim = [-cell_size * self.y_dim / 2, cell_size * self.y_dim / 2]
        self.map_torch = None

    def convert_map(self):
        """Upload the grid (in tensor_args' dtype) — what the kernel gathers from."""
        self.map_torch = torch.as_tensor(self.map, dtype=self.tensor_args['dtype']).to(self.tensor_args['device']).contiguous()
        return self.map_torch

    # -- field protocol (costs/fields.py:7-27) --------------------------------------------------------
    def compute_cost(self, X, **kwargs):
        """Occupancy value at positions X [..., 2] through the CUDA lookup."""
        from ..costs.cost_functions import _map_lookup_cuda
        return _map_lookup_cuda(self, X)

    get_collisions = compute_cost
    __call__ = compute_cost

    def zero_grad(self):
        pass


class _Obstacle:
    def __init__(self, center_x, center_y):
        self.center_x, self.center_y = center_x, center_y
        self.origin = np.array([center_x, center_y])

    def cells(self, grid):
        """(row index array, col index array) of the cells this obstacle covers."""
        raise NotImplementedError

    def _add_to_map(self, grid):
        iy, ix = self.cells(grid)
        ny, nx = grid.map.shape
        if iy.size and (iy.min() < 0 or iy.max() >= ny or ix.min() < 0 or ix.max() >= nx):
            raise IndexError("obstacle at (%g, %g) leaves the map" % (self.center_x, self.center_y))
        np.add.at(grid.map, (iy, ix), 1)
        return grid

    def overlaps(self, grid):
        iy, ix = self.cells(grid)
        ny, nx = grid.map.shape
        if iy.size and (iy.min() < 0 or iy.max() >= ny or ix.min() < 0 or ix.max() >= nx):
            raise IndexError("obstacle at (%g, %g) leaves the map" % (self.center_x, self.center_y))
        # the reference rejects a candidate when ANY cell of (grid + candidate) exceeds 1 (obst_map.py:21-26)
        return bool(np.any(grid.map > 1) or np.any(grid.map[iy, ix] >= 1))


class ObstacleRectangle(_Obstacle):
    def __init__(self, center_x=0, center_y=0, width=None, height=None):
        super().__init__(center_x, center_y)
        self.width, self.height = width, height

    def cells(self, grid):
        cs = grid.cell_size
        w, h = math.ceil(self.width / cs), math.ceil(self.height / cs)
        cx, cy = math.ceil(self.center_x / cs), math.ceil(self.center_y / cs)
        rows = np.arange(cy - math.ceil(h / 2.) + grid.origin_yi, cy + math.ceil(h / 2.) + grid.origin_yi)
        cols = np.arange(cx - math.ceil(w / 2.) + grid.origin_xi, cx + math.ceil(w / 2.) + grid.origin_xi)
        iy, ix = np.meshgrid(rows, cols, indexing='ij')
        return iy.reshape(-1), ix.reshape(-1)


class ObstacleCircle(_Obstacle):
    def __init__(self, center_x=0, center_y=0, radius=1.):
        super().__init__(center_x, center_y)
        self.radius = radius

    def cells(self, grid):
        cs = grid.cell_size
        cr = math.ceil(self.radius / cs)
        cx, cy = math.ceil(self.center_x / cs), math.ceil(self.center_y / cs)
        rows = np.arange(cy - 2 * cr + grid.origin_yi, cy + 2 * cr + grid.origin_yi)
        cols = np.arange(cx - 2 * cr + grid.origin_xi, cx + 2 * cr + grid.origin_xi)
        iy, ix = np.meshgrid(rows, cols, indexing='ij')
        px = (ix - grid.origin_xi) * cs
        py = (iy - grid.origin_yi) * cs
        inside = np.sqrt((px - self.center_x) ** 2 + (py - self.center_y) ** 2) <= self.radius
        return iy[inside], ix[inside]


def random_rect(xlim=(0, 0), ylim=(0, 0), width=2, height=2):
    cx = random.uniform(xlim[0], xlim[1])
    cy = random.uniform(ylim[0], ylim[1])
    return ObstacleRectangle(cx, cy, width, height)


def random_circle(xlim=(0, 0), ylim=(0, 0), radius=2):
    cx = random.uniform(xlim[0], xlim[1])
    cy = random.uniform(ylim[0], ylim[1])
    return ObstacleCircle(cx, cy, radius)


def generate_obstacle_map(map_dim=(10, 10), obst_list=(), cell_size=1., random_gen=False, num_obst=0,
                          rand_limits=None, rand_rect_shape=(2, 2), rand_circle_radius=1, tensor_args=None):
    """Grid with the listed obstacles plus random non-overlapping rectangles/circles up to `num_obst`.
    Draw order per attempt (np.random.choice(2), then random.uniform x2) matches the reference so that
    seeding `random` and `np.random` identically reproduces its maps.  Returns (ObstacleMap, obstacles)."""
    grid = ObstacleMap(map_dim, cell_size, tensor_args=tensor_args)
    placed = list(obst_list)
    for ob in placed:
        ob._add_to_map(grid)
    n_fixed = len(placed)
    if random_gen:
        if n_fixed > num_obst:
            raise AssertionError("num_obst must be >= len(obst_list)")
        for _ in range(num_obst - n_fixed):
            for attempt in range(26):
                if np.random.choice(2):
                    ob = random_rect(rand_limits[0], rand_limits[1], rand_rect_shape[0], rand_rect_shape[1])
                else:
                    ob = random_circle(rand_limits[0], rand_limits[1], rand_circle_radius)
                if not ob.overlaps(grid):
                    ob._add_to_map(grid)
                    placed.append(ob)
                    break
                if attempt == 25:
                    print("Obstacle generation: Max. number of attempts reached. ")
                    print("Total num. obstacles: {}.  Num. random obstacles: {}.\n".format(len(placed), len(placed) - n_fixed))
    grid.convert_map()
    return grid, placed
