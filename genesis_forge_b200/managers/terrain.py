"""
TerrainManager: terrain / sub-terrain bounds, cached height field, height lookup, spawn positions.

API of genesis_forge/managers/terrain_manager.py.  `get_bounds` feeds the out_of_bounds
termination and the height field feeds `rewards.base_height(terrain_manager=...)`, both evaluated
in the fused kernel from the values cached here (:285-359).  Spawn-position sampling
(:168-279) is reset-side work on the compacted reset indices and runs on the host.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .._gs import gs
from .base import BaseManager


class TerrainManager(BaseManager):
    def __init__(self, env, terrain_attr: str = "terrain"):
        super().__init__(env, type="terrain")
        self._origin = (0, 0, 0)
        self._bounds = (0, 0, 0, 0)  # x_min, x_max, y_min, y_max
        self._size = (0, 0)
        self._terrain = None
        self._terrain_attr = terrain_attr
        self._subterrain_bounds = {}
        self._subterrain_size = None
        self._height_field: torch.Tensor | None = None
        self._env_pos_buffer = torch.zeros((env.num_envs, 3), device=gs.device, dtype=gs.tc_float)

    def build(self):
        self._terrain = getattr(self.env, self._terrain_attr)
        self._map_terrain()

    def get_bounds(self, subterrain: str | None = None):
        if subterrain is not None and subterrain in self._subterrain_bounds:
            return self._subterrain_bounds[subterrain]
        return self._bounds

    @property
    def height_field(self) -> torch.Tensor | None:
        """(Hf, Wf) heights in metres, laid out for a (x -> column, y -> row) bilinear lookup."""
        return self._height_field

    def get_terrain_height(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """Bilinear terrain height at world (x, y) (terrain_manager.py:100-166)."""
        n = x.shape[0]
        if self._height_field is None:
            return torch.full((n,), float(self._origin[2]), device=gs.device, dtype=gs.tc_float)
        x_min, x_max, y_min, y_max = self._bounds
        gx = (x - x_min) / (x_max - x_min) * 2 - 1
        gy = (y - y_min) / (y_max - y_min) * 2 - 1
        grid = torch.stack([gx, gy], dim=-1).reshape(n, 1, 1, 2)
        field = self._height_field.unsqueeze(0).unsqueeze(0).expand(n, -1, -1, -1)
        out = F.grid_sample(field, grid, mode="bilinear", padding_mode="border", align_corners=True)
        return out[:, 0, 0, 0]

    def generate_random_positions(
        self, num: int | None = None, usable_ratio: float = 0.5, subterrain: str | None = None,
        height_offset: float = 0.1e-3, output: torch.Tensor | None = None, out_idx: torch.Tensor | None = None,
    ) -> torch.Tensor:
        """Random (x, y) inside the usable centre of the (sub)terrain, z = terrain height + offset."""
        assert output is not None or num is not None, "Either output or num must be provided"
        if output is None:
            output = torch.zeros(num, 3, device=gs.device)
        if out_idx is None:
            out_idx = torch.arange(output.shape[0], device=gs.device)
        bounds, size = self._bounds, self._size
        if subterrain is not None and subterrain in self._subterrain_bounds:
            size, bounds = self._subterrain_size, self._subterrain_bounds[subterrain]
        x_origin, _, y_origin, _ = bounds
        x_size, y_size = size
        margin_x = (x_size - x_size * usable_ratio) / 2
        margin_y = (y_size - y_size * usable_ratio) / 2
        x_lo, x_hi = x_origin + margin_x, x_origin + x_size - margin_x
        y_lo, y_hi = y_origin + margin_y, y_origin + y_size - margin_y
        like = output[out_idx, 0]
        output[out_idx, 0] = self.env.rng.uniform("spawn_x", like, 0.0, 1.0) * (x_hi - x_lo) + x_lo
        output[out_idx, 1] = self.env.rng.uniform("spawn_y", like, 0.0, 1.0) * (y_hi - y_lo) + y_lo
        heights = self.get_terrain_height(output[out_idx, 0], output[out_idx, 1])
        output[out_idx, 2] = heights + height_offset
        return output

    def generate_random_env_pos(
        self, envs_idx=None, usable_ratio: float = 0.5, subterrain: str | None = None, height_offset: float = 0.1e-3,
    ) -> torch.Tensor:
        if envs_idx is None:
            envs_idx = torch.arange(self.env.num_envs, device=gs.device)
        self.generate_random_positions(
            output=self._env_pos_buffer, out_idx=envs_idx, usable_ratio=usable_ratio,
            subterrain=subterrain, height_offset=height_offset,
        )
        return self._env_pos_buffer[envs_idx]

    def _map_terrain(self):
        (geom,) = self._terrain.geoms
        morph = self._terrain.morph
        if hasattr(morph, "pos") and getattr(morph, "n_subterrains", None) is not None:
            self._origin = morph.pos
            sx, sy = morph.subterrain_size
            nx, ny = morph.n_subterrains
            self._size = (sx * nx, sy * ny)
            x0, y0 = self._origin[0], self._origin[1]
            self._bounds = (x0, x0 + self._size[0], y0, y0 + self._size[1])
            self._subterrain_size = morph.subterrain_size
            self._subterrain_bounds = {}
            for ix in range(nx):
                for iy in range(ny):
                    name = morph.subterrain_types[ix][iy]
                    bx, by = x0 + ix * sx, y0 + iy * sy
                    self._subterrain_bounds[name] = (bx, bx + sx, by, by + sy)
        else:
            aabb, pos = geom.get_AABB(), geom.get_pos()
            if aabb.ndim == 3:
                aabb = aabb[0]
            if pos.ndim == 2:
                pos = pos[0]
            (x_min, y_min, _), (x_max, y_max, _) = aabb[0], aabb[1]
            self._origin = pos
            self._size = (x_max - x_min, y_max - y_min)
            self._bounds = (x_min, x_max, y_min, y_max)
        if "height_field" in geom.metadata:
            field = torch.as_tensor(geom.metadata["height_field"], device=gs.device, dtype=gs.tc_float)
            self._height_field = (field * morph.vertical_scale).T.contiguous()
