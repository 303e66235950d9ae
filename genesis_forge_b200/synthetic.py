"""
Synthetic physics engine: serves the slice of the Genesis entity / scene API that the manager step
touches from seeded synthetic state tensors instead of a simulation.

BASELINE.json measures the manager step "on synthetic physics-state tensors of the named shape";
this module is that state source.  It is used by bench.py, by the parity tests (one instance feeds
the torch-CPU oracle, an identically seeded one feeds the CUDA path) and by smoke().

State distributions follow SURVEY.md section 8(d): base xy ~ U(-10,10), z ~ N(0.32,0.03); base quat
= normalize([1, N(0,s)*3]) with s small for most envs and large for a few percent (so orientation
terminations and resets are exercised at a realistic rate); vel/ang ~ N(0,0.5); dof_pos ~ nominal +
N(0,0.2); dof_vel ~ N(0,2); C padded contact slots with n_valid ~ U{0..C}, forces ~ N(0,30) on valid
slots and zeros on padding.

Engine semantics encoded here (all [unverified] against real Genesis):
quaternions are w-first; get_vel/get_ang are world frame; padded contact slots are all-zero;
`force` is the force on link_b; get_links_quat() is indexed by global link idx; setters take effect
immediately so getters called after a reset return the reset state.
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field
from typing import Callable

import torch

from ._gs import gs


# --------------------------------------------------------------------------------------------
# Robot tables
# --------------------------------------------------------------------------------------------
@dataclass
class RobotModel:
    name: str
    joint_names: list[str]          # revolute joints, in DOF order
    link_names: list[str]           # robot links, local index order (0 = base)
    dof_lower: list[float]
    dof_upper: list[float]
    nominal_pos: list[float]        # centre of the synthetic dof_pos distribution
    dof_start: int = 6              # DOFs 0..5 belong to the free base joint


def _go2() -> RobotModel:
    legs = ["FL", "FR", "RL", "RR"]
    joints, links, lo, hi, nom = [], ["base"], [], [], []
    for leg in legs:
        rear = leg[0] == "R"
        joints += [f"{leg}_hip_joint", f"{leg}_thigh_joint", f"{leg}_calf_joint"]
        links += [f"{leg}_hip", f"{leg}_thigh", f"{leg}_calf", f"{leg}_foot"]
        lo += [-1.0472, -0.5236 if rear else -1.5708, -2.7227]
        hi += [1.0472, 4.5379 if rear else 3.4907, -0.83776]
        nom += [0.0, 1.0 if rear else 0.8, -1.5]
    return RobotModel("go2", joints, links, lo, hi, nom)


def _berkeley_humanoid() -> RobotModel:
    names = ["HR", "HAA", "HFE", "KFE", "FFE", "FAA"]
    lim = {
        "HR": (-0.610865, 0.610865), "HAA": (-0.610865, 0.610865), "HFE": (-1.74533, 0.523599),
        "KFE": (0.0, 2.0944), "FFE": (-0.523599, 0.698132), "FAA": (-0.523599, 0.523599),
    }
    nom_l = {"HR": -0.071, "HAA": 0.103, "HFE": -0.463, "KFE": 0.983, "FFE": -0.350, "FAA": 0.126}
    joints, links, lo, hi, nom = [], ["torso"], [], [], []
    for side, sgn in (("LL", 1.0), ("LR", -1.0)):
        for n in names:
            joints.append(f"{side}_{n}")
            links.append(f"{side.lower()}_{n.lower()}")
            lo.append(lim[n][0])
            hi.append(lim[n][1])
            nom.append(nom_l[n] * (sgn if n in ("HR", "HAA", "FAA") else 1.0))
    return RobotModel("berkeley_humanoid", joints, links, lo, hi, nom)


GO2 = _go2()
BERKELEY_HUMANOID = _berkeley_humanoid()
ROBOT_MODELS = {"go2": GO2, "berkeley_humanoid": BERKELEY_HUMANOID}


# --------------------------------------------------------------------------------------------
# State generation
# --------------------------------------------------------------------------------------------
STATE_KEYS = (
    "pos", "quat", "vel", "ang", "dof_pos", "dof_vel", "dof_force",
    "c_force", "c_pos", "c_link_a", "c_link_b", "links_quat", "links_vel", "links_pos",
)


def generate_state(
    model: RobotModel,
    n_envs: int,
    n_contacts: int,
    gen: torch.Generator,
    tilt_fraction: float = 0.015,
    xy_range: float = 10.0,
) -> dict[str, torch.Tensor]:
    """One synthetic physics state (CPU tensors) for `n_envs` copies of `model`."""
    N, D, C = n_envs, len(model.joint_names), n_contacts
    # + ground plane, global link idx 0.  Without contact slots nothing reads per-link state: keep
    # those arrays minimal so million-env benchmarks of contact-free configs do not carry them.
    L = 1 + len(model.link_names) if C > 0 else 1
    f32 = torch.float32

    def randn(*shape):
        return torch.randn(*shape, generator=gen, dtype=f32)

    def rand(*shape):
        return torch.rand(*shape, generator=gen, dtype=f32)

    s = {}
    pos = torch.empty(N, 3, dtype=f32)
    pos[:, :2] = (rand(N, 2) * 2 - 1) * xy_range
    pos[:, 2] = 0.32 + 0.03 * randn(N)
    s["pos"] = pos

    sigma = torch.where(rand(N) < tilt_fraction, 0.5, 0.02).unsqueeze(1)
    q = torch.cat([torch.ones(N, 1, dtype=f32), randn(N, 3) * sigma], dim=1)
    s["quat"] = q / q.norm(dim=1, keepdim=True)

    s["vel"] = 0.5 * randn(N, 3)
    s["ang"] = 0.5 * randn(N, 3)
    s["dof_pos"] = torch.tensor(model.nominal_pos, dtype=f32) + 0.2 * randn(N, D)
    s["dof_vel"] = 2.0 * randn(N, D)
    s["dof_force"] = 5.0 * randn(N, D)

    n_valid = torch.randint(0, C + 1, (N, 1), generator=gen)
    valid = torch.arange(C).unsqueeze(0) < n_valid  # (N, C)
    Lr = max(L, 3)
    link_b = torch.randint(1, Lr, (N, C), generator=gen)
    # the base/torso link (global idx 1) rarely touches anything: keep one in ten of its draws
    demote = (link_b == 1) & (rand(N, C) > 0.1)
    link_b = torch.where(demote, torch.randint(2, Lr, (N, C), generator=gen), link_b)
    # mostly ground contacts (link_a = plane = 0); one in five is a self contact
    self_contact = rand(N, C) < 0.2
    link_a = torch.where(self_contact, torch.randint(2, Lr, (N, C), generator=gen), 0)
    s["c_link_a"] = torch.where(valid, link_a, 0).to(torch.int32)
    s["c_link_b"] = torch.where(valid, link_b, 0).to(torch.int32)
    s["c_force"] = (30.0 * randn(N, C, 3)) * valid.unsqueeze(-1)
    s["c_pos"] = (rand(N, C, 3) * 2 - 1) * valid.unsqueeze(-1)

    lq = randn(N, L, 4)
    s["links_quat"] = lq / lq.norm(dim=-1, keepdim=True)
    s["links_vel"] = 0.5 * randn(N, L, 3)
    s["links_pos"] = randn(N, L, 3)
    return s


class StateSource:
    """Deterministic stream of states: `source(k)` is the state after the k-th physics step."""

    def __init__(self, model: RobotModel, n_envs: int, n_contacts: int = 8, seed: int = 1234, **kw):
        self.model, self.n_envs, self.n_contacts, self.seed, self.kw = model, n_envs, n_contacts, seed, kw

    def __call__(self, step_index: int) -> dict[str, torch.Tensor]:
        gen = torch.Generator().manual_seed(self.seed * 1_000_003 + step_index)
        return generate_state(self.model, self.n_envs, self.n_contacts, gen, **self.kw)


class CachedSource:
    """Wraps a source so that two engines (oracle / CUDA) receive the identical state objects."""

    def __init__(self, source: Callable[[int], dict], post: Callable[[dict, int], dict] | None = None):
        self._source, self._post, self._cache = source, post, {}

    def __call__(self, step_index: int) -> dict[str, torch.Tensor]:
        if step_index not in self._cache:
            state = self._source(step_index)
            if self._post is not None:
                state = self._post(state, step_index)
            self._cache = {k: v for k, v in self._cache.items() if k >= step_index - 2}
            self._cache[step_index] = state
        return {k: v.clone() for k, v in self._cache[step_index].items()}


# --------------------------------------------------------------------------------------------
# Engine objects
# --------------------------------------------------------------------------------------------
@dataclass
class SyntheticJoint:
    name: str
    type: object
    dof_start: int


@dataclass
class SyntheticLink:
    name: str
    idx: int
    idx_local: int
    _entity: object = field(default=None, repr=False)

    def get_vel(self) -> torch.Tensor:
        return self._entity._scene._get("links_vel")[:, self.idx, :]

    def get_pos(self) -> torch.Tensor:
        return self._entity._scene._get("links_pos")[:, self.idx, :]


class SyntheticPlane:
    """Ground plane entity: one link, global idx 0."""

    def __init__(self, scene):
        self._scene = scene
        self.links = [SyntheticLink("plane", 0, 0, self)]
        self.joints = []
        self.geoms = []
        self.morph = None

    def get_link(self, name):
        return self.links[0]


class _TerrainGeom:
    def __init__(self, terrain):
        self._t = terrain
        self.metadata = {"height_field": terrain.height_field}

    def get_AABB(self):
        t = self._t
        lo = torch.tensor([t.morph.pos[0], t.morph.pos[1], 0.0])
        hi = torch.tensor([t.morph.pos[0] + t.size[0], t.morph.pos[1] + t.size[1], 1.0])
        return torch.stack([lo, hi])

    def get_pos(self):
        return torch.tensor(self._t.morph.pos, dtype=torch.float32)


@dataclass
class _TerrainMorph:
    pos: tuple
    n_subterrains: tuple
    subterrain_size: tuple
    subterrain_types: list
    vertical_scale: float


class SyntheticTerrain:
    """
    Height-field terrain entity (what `gs.morphs.Terrain` produces, as far as TerrainManager reads it:
    genesis_forge/managers/terrain_manager.py:285-359).  One link, global idx 0.
    """

    def __init__(self, scene, pos=(-12.0, -12.0, 0.0), n_subterrains=(1, 1), subterrain_size=(24.0, 24.0),
                 subterrain_types=None, vertical_scale=0.001, cells=(96, 96), seed=7):
        self._scene = scene
        self.links = [SyntheticLink("terrain", 0, 0, self)]
        self.joints = []
        types = subterrain_types or [["random_uniform_terrain"] * n_subterrains[1] for _ in range(n_subterrains[0])]
        self.morph = _TerrainMorph(tuple(pos), tuple(n_subterrains), tuple(subterrain_size), types, vertical_scale)
        self.size = (subterrain_size[0] * n_subterrains[0], subterrain_size[1] * n_subterrains[1])
        gen = torch.Generator().manual_seed(seed)
        # integer heights in units of vertical_scale, like Genesis' int16 height fields
        self.height_field = torch.randint(0, 100, cells, generator=gen).to(torch.float32).numpy()
        self.geoms = [_TerrainGeom(self)]

    def get_link(self, name):
        return self.links[0]


class SyntheticRobot:
    """The articulated entity.  Getter/setter names and argument meaning follow RigidEntity."""

    def __init__(self, scene, model: RobotModel, link_offset: int = 1):
        self._scene = scene
        self.model = model
        JT = gs.JOINT_TYPE
        self.joints = [SyntheticJoint("root_joint", JT.FREE, 0)] + [
            SyntheticJoint(n, JT.REVOLUTE, model.dof_start + i) for i, n in enumerate(model.joint_names)
        ]
        self.links = [
            SyntheticLink(n, link_offset + i, i, self) for i, n in enumerate(model.link_names)
        ]
        self.n_dofs = len(model.joint_names)
        self.calls: list[tuple] = []          # record of setter calls (tests inspect it)
        self.record_calls = False
        self.last_control_target: torch.Tensor | None = None

    # -- topology -----------------------------------------------------------------------------
    def get_link(self, name: str) -> SyntheticLink:
        for link in self.links:
            if link.name == name:
                return link
        raise KeyError(name)

    def get_dofs_limit(self, dofs_idx=None):
        dev = self._scene.device
        return (
            torch.tensor(self.model.dof_lower, device=dev, dtype=torch.float32),
            torch.tensor(self.model.dof_upper, device=dev, dtype=torch.float32),
        )

    # -- getters ------------------------------------------------------------------------------
    def get_pos(self, envs_idx=None):
        return self._scene._get("pos")

    def get_quat(self, envs_idx=None):
        return self._scene._get("quat")

    def get_vel(self, envs_idx=None):
        return self._scene._get("vel")

    def get_ang(self, envs_idx=None):
        return self._scene._get("ang")

    def get_dofs_position(self, dofs_idx=None, envs_idx=None):
        return self._scene._get("dof_pos")

    def get_dofs_velocity(self, dofs_idx=None, envs_idx=None):
        return self._scene._get("dof_vel")

    def get_dofs_force(self, dofs_idx=None, envs_idx=None):
        return self._scene._get("dof_force")

    def _global_idx(self, links_idx_local):
        off = self.links[0].idx
        if links_idx_local is None:
            return torch.arange(off, off + len(self.links), device=self._scene.device)
        idx = torch.as_tensor(links_idx_local, device=self._scene.device).long()
        return idx + off

    def get_links_vel(self, links_idx_local=None, envs_idx=None):
        return self._scene._get("links_vel")[:, self._global_idx(links_idx_local), :]

    def get_links_pos(self, links_idx_local=None, envs_idx=None):
        return self._scene._get("links_pos")[:, self._global_idx(links_idx_local), :]

    def get_AABB(self):
        pos = self._scene._get("pos")
        return torch.stack([pos - 0.3, pos + 0.3], dim=1)

    # -- setters ------------------------------------------------------------------------------
    def _rec(self, name, *args):
        if self.record_calls:
            self.calls.append((name,) + tuple(a.detach().clone() if torch.is_tensor(a) else a for a in args))

    def control_dofs_position(self, position, dofs_idx_local=None, envs_idx=None):
        self.last_control_target = position
        self._rec("control_dofs_position", position)

    def _noop_setter(name):
        def fn(self, *args, **kwargs):
            self._rec(name, *args)
        fn.__name__ = name
        return fn

    set_dofs_kp = _noop_setter("set_dofs_kp")
    set_dofs_kv = _noop_setter("set_dofs_kv")
    set_dofs_damping = _noop_setter("set_dofs_damping")
    set_dofs_stiffness = _noop_setter("set_dofs_stiffness")
    set_dofs_frictionloss = _noop_setter("set_dofs_frictionloss")
    set_dofs_force_range = _noop_setter("set_dofs_force_range")
    set_mass_shift = _noop_setter("set_mass_shift")
    del _noop_setter

    def set_dofs_position(self, position, dofs_idx_local=None, envs_idx=None, zero_velocity=True):
        self._rec("set_dofs_position", position, envs_idx)
        if not self._scene.apply_setters:
            return
        st = self._scene.state
        st["dof_pos"][envs_idx] = position
        if zero_velocity:
            st["dof_vel"][envs_idx] = 0.0

    def _zero_base_velocity(self, envs_idx):
        if not self._scene.apply_setters:
            return
        st = self._scene.state
        st["vel"][envs_idx] = 0.0
        st["ang"][envs_idx] = 0.0
        st["dof_vel"][envs_idx] = 0.0

    def set_pos(self, pos, envs_idx=None, zero_velocity=True, relative=False):
        if self._scene.apply_setters:
            self._scene.state["pos"][envs_idx] = pos
        if zero_velocity:
            self._zero_base_velocity(envs_idx)
        self._rec("set_pos", pos, envs_idx)

    def set_quat(self, quat, envs_idx=None, zero_velocity=True, relative=False):
        if self._scene.apply_setters:
            self._scene.state["quat"][envs_idx] = quat
        if zero_velocity:
            self._zero_base_velocity(envs_idx)
        self._rec("set_quat", quat, envs_idx)

    def zero_all_dofs_velocity(self, envs_idx=None):
        self._zero_base_velocity(envs_idx)
        self._rec("zero_all_dofs_velocity", envs_idx)


class _Collider:
    def __init__(self, scene):
        self._scene = scene

    def get_contacts(self, as_tensor: bool = True, to_torch: bool = True):
        g = self._scene._get
        return {
            "force": g("c_force"), "position": g("c_pos"),
            "link_a": g("c_link_a"), "link_b": g("c_link_b"),
        }


class _RigidSolver:
    def __init__(self, scene):
        self._scene = scene
        self.collider = _Collider(scene)

    def get_links_quat(self):
        return self._scene._get("links_quat")


class SyntheticScene:
    """
    Stand-in for `gs.Scene`.

    mode "regen": every `step()` pulls a fresh state from `source(step_index)` (parity tests).
    mode "pool":  `pool` states are generated once on the device and `step()` rotates through
                  them (benchmarks: >= 4 sets, each larger than L2, so kernels read from HBM).
    copy_on_get:  getters return clones (what real Genesis does).  The reference's observation
                  manager scales getter results in place (observation_manager.py:243-250), so the
                  oracle side needs this; the CUDA path never writes to getter results and reads
                  the state tensors zero-copy.
    """

    def __init__(
        self,
        dt: float = 0.02,
        n_contacts: int = 8,
        seed: int = 1234,
        device=None,
        source: Callable[[int], dict] | None = None,
        copy_on_get: bool = False,
        pool: int = 0,
        source_kw: dict | None = None,
        apply_setters: bool = True,
    ):
        self.dt = dt
        self._source_kw = source_kw or {}
        # False: state setters are accepted but not applied.  Used by throughput benchmarks in pool
        # mode, where the state sets are reused: applying resets would leave every env of every set in
        # its reset pose after one cycle and no termination would ever fire again.  The manager path
        # (index compaction, reset fan-out, setter calls, re-observation) still runs in full.
        self.apply_setters = apply_setters
        self.n_contacts = n_contacts
        self.seed = seed
        self.device = torch.device(device) if device is not None else gs.device
        self._source = source
        self.copy_on_get = copy_on_get
        self.pool_size = pool
        self._pool: list[dict] = []
        self.state: dict[str, torch.Tensor] = {}
        self.step_index = 0
        self.n_envs = 0
        self.envs_offset = None
        self.robot: SyntheticRobot | None = None
        self.plane: SyntheticPlane | None = None
        self.rigid_solver = _RigidSolver(self)
        self.is_built = False

    # -- construction -------------------------------------------------------------------------
    def add_plane(self) -> SyntheticPlane:
        self.plane = SyntheticPlane(self)
        return self.plane

    def add_terrain(self, **kw) -> SyntheticTerrain:
        self.plane = SyntheticTerrain(self, **kw)
        return self.plane

    def add_robot(self, model: RobotModel | str) -> SyntheticRobot:
        if isinstance(model, str):
            model = ROBOT_MODELS[model]
        self.robot = SyntheticRobot(self, model, link_offset=1)
        return self.robot

    def build(self, n_envs: int = 1, **_):
        assert self.robot is not None, "add_robot() first"
        self.n_envs = n_envs
        if self._source is None:
            self._source = StateSource(self.robot.model, n_envs, self.n_contacts, self.seed, **self._source_kw)
        self.envs_offset = torch.zeros(n_envs, 3, device=self.device)
        if self.pool_size > 0:
            self._pool = [self._to_device(self._source(k)) for k in range(self.pool_size)]
            self.state = self._pool[0]
        else:
            self.state = self._to_device(self._source(0))
        self.is_built = True

    def _to_device(self, state: dict) -> dict:
        return {k: v.to(self.device).contiguous() for k, v in state.items()}

    # -- stepping -----------------------------------------------------------------------------
    def step(self):
        self.step_index += 1
        if self.pool_size > 0:
            self.state = self._pool[self.step_index % self.pool_size]
        else:
            self.state = self._to_device(self._source(self.step_index))

    def _get(self, key: str) -> torch.Tensor:
        t = self.state[key]
        return t.clone() if self.copy_on_get else t

    # -- viz no-ops ---------------------------------------------------------------------------
    def draw_debug_arrow(self, *a, **k):
        return None

    def draw_debug_spheres(self, *a, **k):
        return None

    def clear_debug_object(self, *a, **k):
        return None


def links_matching(entity, patterns: list[str]) -> list[SyntheticLink]:
    out = []
    for pattern in patterns:
        for link in entity.links:
            if pattern == link.name or re.match(f"^{pattern}$", link.name):
                out.append(link)
    return out
