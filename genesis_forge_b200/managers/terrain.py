"""
TerrainManager: terrain / sub-terrain bounds, cached height field, height lookup, spawn positions.

API of genesis_forge/managers/terrain_manager.py.  `get_bounds` feeds the out_of_bounds
termination and the height field feeds `rewards.base_height(terrain_manager=...)`, both evaluated
in the fused kernel from the values cached here (:285-359).  Spawn-position sampling
(:168-279) is reset-side work on the compacted reset indices: one launch of the library's spawn
kernel (gfb_spawn_pose) instead of the reference's chain of indexed torch ops.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn.functional as F

from .. import _native as nat
from .._gs import gs
from ..rng import HostRng
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
        self._own_handle = None
        self._spawn_calls = 0
        self._spawn_cfgs: dict = {}

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
        self._spawn(output, out_idx, usable_ratio, subterrain, height_offset)
        return output

    def generate_random_env_pos(
        self, envs_idx=None, usable_ratio: float = 0.5, subterrain: str | None = None, height_offset: float = 0.1e-3,
    ) -> torch.Tensor:
        pos, _ = self._spawn(self._env_pos_buffer, envs_idx, usable_ratio, subterrain, height_offset, compact=True)
        return pos

    # -- the spawn kernel -------------------------------------------------------------------------
    def _handle(self):
        fused = getattr(self.env, "_fused", None)
        if fused is not None and not fused.dry_run:
            return fused.handle, fused.lib
        if self._own_handle is None:  # used before env.build(): a minimal handle of its own
            device = torch.device(gs.device)
            if device.type != "cuda":
                raise nat.NativeLibraryError(
                    f"spawn positions are sampled by the CUDA library (gs.device is {device}); "
                    "there is no CPU implementation in this package"
                )
            index = device.index if device.index is not None else torch.cuda.current_device()
            self._own_handle = nat.Handle(1, index)
        return self._own_handle, nat.lib()

    def _spawn(self, output: torch.Tensor, out_idx, usable_ratio, subterrain, height_offset, compact: bool = False,
               rotation: dict | None = None, rot_buffer: torch.Tensor | None = None,
               quat_buffer: torch.Tensor | None = None):
        """
        One launch of gfb_spawn_pose: positions (terrain_manager.py:168-279) for rows `out_idx` of
        `output`, and -- when `rotation` is given -- the Euler draws and quaternions of
        randomize_terrain_position.define_quat (mdp/reset.py:172-195).  Returns the compact
        (n,3) / (n,4) copies when `compact`.  Draws come from the kernel's Philox stream unless the
        environment carries a custom `rng` (parity harness), whose values are passed through.
        """
        handle, lib = self._handle()
        dev = output.device
        if output.dtype != torch.float32 or not output.is_contiguous() or output.dim() != 2 or output.shape[1] != 3:
            raise ValueError("spawn positions need a contiguous float32 (rows, 3) output tensor")
        if out_idx is not None:
            out_idx = torch.as_tensor(out_idx, device=dev)
            if out_idx.dtype == torch.bool:
                out_idx = out_idx.nonzero().reshape(-1)
            if out_idx.dtype != torch.int64 or not out_idx.is_contiguous():
                out_idx = out_idx.to(torch.int64).contiguous()
            n = out_idx.numel()
        else:
            n = output.shape[0]
        rng = getattr(self.env, "rng", None)
        in_kernel = rng is None or type(rng) is HostRng
        cfg = self._spawn_config(usable_ratio, subterrain, height_offset, rotation)
        field = self._height_field
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
        u_x = u_y = None
        u_rot = [None, None, None]
        if not in_kernel:
            like = torch.empty(n, device=dev)
            u_x = rng.uniform("spawn_x", like, 0.0, 1.0).to(dev, torch.float32).contiguous()
            u_y = rng.uniform("spawn_y", like, 0.0, 1.0).to(dev, torch.float32).contiguous()
            if rotation is not None:
                for col, axis in enumerate(("x", "y", "z")):
                    value = rotation.get(axis, 0)
                    if isinstance(value, tuple):
                        u_rot[col] = rng.uniform(f"spawn_rot_{axis}", like, *value).to(dev, torch.float32).contiguous()
        self._spawn_calls += 1
        cfg.rng_counter = (int(getattr(self.env, "step_count", 0)) << 20) ^ self._spawn_calls
        pos_out = torch.empty((n, 3), device=dev) if compact else None
        quat_out = torch.empty((n, 4), device=dev) if (compact and rotation is not None) else None
        stream = C.c_void_p(torch._C._cuda_getCurrentRawStream(dev.index if dev.index is not None else torch.cuda.current_device()))
        handle.check(
            lib.gfb_spawn_pose(
                handle.ptr, C.byref(cfg), ptr(out_idx), n, output.shape[0], ptr(field), ptr(u_x), ptr(u_y),
                ptr(u_rot[0]), ptr(u_rot[1]), ptr(u_rot[2]), ptr(output),
                ptr(rot_buffer) if rotation is not None else None,
                ptr(quat_buffer) if rotation is not None else None, ptr(pos_out), ptr(quat_out), stream,
            ),
            "gfb_spawn_pose",
        )
        # (temporaries handed to the launch may be released now: same-stream reuse is ordered after it)
        return pos_out, quat_out

    def _spawn_config(self, usable_ratio, subterrain, height_offset, rotation) -> "nat.Spawn":
        """The gfb_spawn block for one (area, offset, rotation) request; built once per distinct request."""
        rot_key = None if rotation is None else tuple(
            (axis, value) for axis in ("x", "y", "z") if isinstance(value := rotation.get(axis, 0), tuple)
        )
        fused = getattr(self.env, "_fused", None)
        key = (usable_ratio, subterrain, height_offset, rot_key, self._bounds, self._size,
               None if self._height_field is None else self._height_field.data_ptr(), getattr(fused, "rng_seed", None))
        cfg = self._spawn_cfgs.get(key)
        if cfg is not None:
            return cfg
        bounds, size = self._bounds, self._size
        if subterrain is not None and subterrain in self._subterrain_bounds:
            size, bounds = self._subterrain_size, self._subterrain_bounds[subterrain]
        x_origin, _, y_origin, _ = bounds
        x_size, y_size = size
        margin_x = (x_size - x_size * usable_ratio) / 2
        margin_y = (y_size - y_size * usable_ratio) / 2
        x_lo, x_hi = x_origin + margin_x, x_origin + x_size - margin_x
        y_lo, y_hi = y_origin + margin_y, y_origin + y_size - margin_y

        cfg = nat.Spawn()
        cfg.x_lo, cfg.x_span, cfg.y_lo, cfg.y_span = float(x_lo), float(x_hi - x_lo), float(y_lo), float(y_hi - y_lo)
        cfg.height_offset = float(height_offset)
        cfg.flat_height = float(self._origin[2])
        field = self._height_field
        if field is not None:
            cfg.height_field_rows, cfg.height_field_cols = field.shape
            for k in range(4):
                cfg.terrain_bounds[k] = float(self._bounds[k])
        if rotation is not None:
            cfg.with_rotation = 1
            for col, axis in enumerate(("x", "y", "z")):
                value = rotation.get(axis, 0)
                if isinstance(value, tuple):  # fixed values are ignored by the reference as well (reset.py:176-190)
                    cfg.rot_mode[col] = nat.K["GFB_SPAWN_ROT_DRAW"]
                    cfg.rot_lo[col], cfg.rot_hi[col] = float(value[0]), float(value[1])
        cfg.rng_seed = (getattr(fused, "rng_seed", 0x5EED) << 8) ^ 0x7E44A1  # per-rank seed, own stream
        if len(self._spawn_cfgs) > 64:
            self._spawn_cfgs.clear()
        self._spawn_cfgs[key] = cfg
        return cfg

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
