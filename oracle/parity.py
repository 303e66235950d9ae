"""
TEST INFRASTRUCTURE (oracle).  Step-by-step parity harness: CUDA drop-in vs the oracle port.

One seeded synthetic state stream (guard-banded, see oracle/guard.py) feeds two environments:
  * `PortEnv` on torch-CPU (the oracle; draws its random numbers from torch's global CPU generator
    exactly where the reference does and logs every draw), and
  * the genesis_forge_b200 drop-in on the GPU, built by the same `env_builder.build_env` code that
    builds the unmodified reference, with the oracle's draws INJECTED: dense buffers for draws made
    inside the kernels, a replay queue for host-side (reset-time) draws.
After reset and after every step the harness compares every persistent buffer, every step output
and every logged extras entry:
    bool masks, int counters, reset indices        bit-exact
    fp32 values that are pure copies / single ops  bit-exact is expected, 1e-5 is asserted
    fp32 values in general                         |a-b| <= 1e-5 |b| + 1e-6
"""
from __future__ import annotations

import torch

from genesis_forge_b200.rng import ReplayRng
from genesis_forge_b200.synthetic import CachedSource, ROBOT_MODELS, StateSource

from configs import specs

from . import compare
from configs.env_builder import build_env, dropin_namespace, make_scene
from .guard import make_sanitizer
from .manager_port import PortEnv

RTOL, ATOL = 1e-5, 1e-6

EXACT_KEYS = ("episode_length", "max_episode_length", "terminated", "truncated")


class ParityFailure(AssertionError):
    pass


def _close(a: torch.Tensor, b: torch.Tensor):
    """(ok, max_abs_err, max_rel_err) of a against reference b; NaNs must coincide."""
    a, b = a.detach().cpu(), b.detach().cpu()
    if a.shape != b.shape:
        return False, float("inf"), float("inf")
    if not a.is_floating_point():
        return bool(torch.equal(a, b)), float((a != b).sum()), 0.0
    nan_a, nan_b = a.isnan(), b.isnan()
    if not torch.equal(nan_a, nan_b):
        return False, float("nan"), float("nan")
    a, b = torch.where(nan_a, 0.0, a.double()), torch.where(nan_b, 0.0, b.double())
    err = (a - b).abs()
    ok = bool((err <= RTOL * b.abs() + ATOL).all())
    rel = (err / b.abs().clamp(min=1e-3)).max().item() if err.numel() else 0.0
    return ok, err.max().item() if err.numel() else 0.0, rel


class _TransplantPort(PortEnv):
    """
    Oracle for the PRODUCTION kernels (in-kernel Philox draws, nothing injected): the draws the
    kernels made are read back from their results and used where the oracle would draw -- the new
    command of a resampled / reset env, the new maximum episode length of a reset env.  Everything
    else is computed by the oracle as usual, so every value that is not itself a draw is compared.
    (The draws' own distribution is covered by test_production_draws_philox.)
    """

    kernel_results: dict | None = None  # {"command/<name>": (N,K), "max_episode_length": (N,)} on the CPU

    def _resample(self, name, c, env_ids, tag):
        src = self.kernel_results[f"command/{name}"]
        for i in range(c["command"].shape[1]):
            c["command"][env_ids, i] = src[env_ids, i]

    def _draw_max_len(self, idx):
        # the u for which round(base + u * span) is the kernel's value (the sum lands within 1e-4 of an integer)
        span = self.base_max_episode_length * self.max_episode_random_scaling
        got = self.kernel_results["max_episode_length"][idx].float()
        return (got - self.base_max_episode_length) / span


class ParityRun:
    def __init__(self, spec_name: str, num_envs: int, device, seed: int = 1234, n_contacts: int = 8,
                 spec_override: dict | None = None, sanitize: bool = True, philox: bool = False):
        """
        `philox=True`: the drop-in runs in production mode (in-kernel Philox draws, the kernel binary
        bench.py times) and the oracle receives the kernels' draws instead of the other way round.
        Only for specs without observation noise and without host-side (reset-time) draws that feed
        compared values.
        """
        import genesis_forge_b200 as gfb

        self.philox = philox

        self.spec = specs.get(spec_name)
        if spec_override:
            self.spec.update(spec_override)
        self.N = num_envs
        self.device = torch.device(device)
        self.stats = {"max_abs": 0.0, "max_rel": 0.0, "steps": 0, "resets": 0, "compared": 0, "inexact": {}}

        model = ROBOT_MODELS[self.spec["robot"]]
        kw = {"xy_range": self.spec["xy_range"]} if "xy_range" in self.spec else {}
        base = StateSource(model, num_envs, n_contacts, seed, **kw)
        # a throw-away scene gives the sanitizer the link tables
        scene0, terrain0, robot0 = make_scene(self.spec, torch.device("cpu"))
        self.sanitizer = make_sanitizer(self.spec, robot0, terrain0)
        # sanitize=False reproduces the exact input stream of the golden fixtures (no guard-band nudging)
        self.source = CachedSource(base, post=self.sanitizer if sanitize else None)

        torch.manual_seed(seed)
        scene, terrain, robot = make_scene(self.spec, torch.device("cpu"), source=self.source, copy_on_get=True,
                                           n_contacts=n_contacts)
        self.port = (_TransplantPort if philox else PortEnv)(self.spec, num_envs, scene, terrain, robot,
                                                             record_margins=True)
        self.port.build()

        gfb.set_device(self.device)
        self.env = build_env(self.spec, dropin_namespace(), num_envs, self.device, source=self.source,
                             n_contacts=n_contacts)
        if not philox:
            self.env.rng = ReplayRng()
        self.env.build()
        self.action_gen = torch.Generator().manual_seed(seed + 77)
        self.group_names = list(self.spec["observations"].keys())
        self.command_names = list(self.spec["commands"].keys())

    # -- injection -------------------------------------------------------------------------------
    def _inject_from_log(self, reset_idx: torch.Tensor | None, resample_idx: dict | None):
        port, N, dev = self.port, self.N, self.device
        draws: dict[str, torch.Tensor] = {}
        for k, name in enumerate(self.command_names):
            kdim = port.command[name]["command"].shape[1]
            draws[f"cmd_step{k}"] = torch.zeros(N, kdim)
            draws[f"cmd_reset{k}"] = torch.zeros(N, kdim)
        draws["max_len"] = torch.zeros(N)
        for g, group in enumerate(self.group_names):
            draws[f"obs_noise{g}"] = torch.zeros(N, port.obs_groups[group]["single"])
        offsets = {}
        for group in self.group_names:
            off, table = 0, {}
            for tname, term in port.obs_groups[group]["terms"].items():
                table[tname] = off
                off += self._term_width(term)
            offsets[group] = table
        replay: ReplayRng = self.env.rng
        replay.clear()
        for tag, value in port.rng_log:
            parts = tag.split(":")
            if parts[0] == "cmd_step":
                k = self.command_names.index(parts[1])
                draws[f"cmd_step{k}"][resample_idx[parts[1]], int(parts[2])] = value
            elif parts[0] == "cmd_reset":
                k = self.command_names.index(parts[1])
                idx = reset_idx if reset_idx is not None else torch.arange(N)
                draws[f"cmd_reset{k}"][idx, int(parts[2])] = value
            elif parts[0] == "max_len":
                idx = reset_idx if reset_idx is not None else torch.arange(N)
                draws["max_len"][idx] = value
            elif parts[0] == "obs_noise":
                g = self.group_names.index(parts[1])
                off = offsets[parts[1]][parts[2]]
                draws[f"obs_noise{g}"][:, off:off + value.shape[1]] = value
            else:  # host-side draws: action_dr:*, spawn_x, spawn_y, spawn_rot_*, gait_* (user-level manager)
                replay.push(tag, value)
        self.env._fused.inject({k: v.to(dev).contiguous() for k, v in draws.items()})

    def _term_width(self, term: dict) -> int:
        kind = term["fn"]
        if callable(kind):
            return int(kind(env=self.port).reshape(self.N, -1).shape[1])
        if kind == "ang_vel_uncached":
            return 3
        if kind == "command":
            c = self.port.command[term["mgr"]]
            return (c["gait"].observation() if c.get("gait") is not None else c["command"]).shape[1]
        if kind in ("ang_vel", "lin_vel", "gravity"):
            return 3
        if kind == "contact_force":
            return self.port.contact[term["mgr"]]["contacts"].shape[1]
        return self.port.num_actions

    # -- comparison ------------------------------------------------------------------------------
    def _check(self, where: str, name: str, got: torch.Tensor, want: torch.Tensor, exact: bool):
        self.stats["compared"] += 1
        got, want = got.detach().cpu(), want.detach().cpu()
        if got.dtype == torch.uint8 and want.dtype == torch.bool:
            got = got.bool()
        if exact:
            if got.shape != want.shape or not torch.equal(got, want):
                n = int((got != want).sum()) if got.shape == want.shape else -1
                raise ParityFailure(f"{where}: {name} differs in {n} of {want.numel()} entries (must be bit-exact)")
            return
        ok, abs_err, rel_err = _close(got, want)
        self.stats["max_abs"] = max(self.stats["max_abs"], abs_err)
        self.stats["max_rel"] = max(self.stats["max_rel"], rel_err)
        if abs_err > 0:
            self.stats["inexact"][name] = max(self.stats["inexact"].get(name, 0.0), abs_err)
        if not ok:
            raise ParityFailure(f"{where}: {name} max abs err {abs_err:.3e}, max rel err {rel_err:.3e}")

    def _compare_all(self, where: str, extras_port: dict, extras_env: dict):
        snap_env = compare.reference_snapshot(self.env)
        snap_port = self.port.snapshot()
        for key, want in snap_port.items():
            if key not in snap_env:
                raise ParityFailure(f"{where}: drop-in has no buffer '{key}'")
            self._check(where, key, snap_env[key], want, exact=key in EXACT_KEYS)
        log_p = compare.extras_to_cpu(extras_port)
        log_e = compare.extras_to_cpu(extras_env)
        if set(log_p) != set(log_e):
            raise ParityFailure(f"{where}: extras keys differ: {sorted(set(log_p) ^ set(log_e))}")
        for key, want in log_p.items():
            self._check(where, f"extras[{key}]", log_e[key], want, exact=False)
        for group in self.group_names:
            self._check(where, f"obs[{group}]", extras_env["observations"][group],
                        extras_port["observations"][group], exact=False)

    def _assert_margins(self, where: str):
        """Every thresholded value the oracle saw this step must be clear of its guard band."""
        from .guard import BAND

        for what, value, thr in self.port.margins:
            if self.philox and what in ("feet_air_cmd", "stand_still_cmd"):
                continue  # thresholds on the kernels' own command draws: nothing to keep clear of in advance
            near = (value - thr).abs() <= 0.25 * BAND * max(abs(thr), 1e-3)
            if bool(near.any()):
                raise ParityFailure(
                    f"{where}: {int(near.sum())} value(s) of '{what}' within the guard band of {thr}; "
                    "choose another seed for this test (documented limitation, SURVEY.md 7-1)"
                )

    # -- drive -----------------------------------------------------------------------------------
    def _kernel_results(self) -> dict:
        out = {"max_episode_length": self.env.max_episode_length.cpu()}
        for name in self.command_names:
            out[f"command/{name}"] = getattr(self.env, name)._command.cpu()
        return out

    def reset(self):
        if self.philox:  # kernels first, their draws go to the oracle
            obs_e, extras_e = self.env.reset()
            self.port.kernel_results = self._kernel_results()
            obs_p, extras_p = self.port.reset()
        else:
            obs_p, extras_p = self.port.reset()
            self._inject_from_log(None, None)
            obs_e, extras_e = self.env.reset()
        self._compare_all("reset", extras_p, extras_e)
        self._check("reset", "obs", obs_e, obs_p, exact=False)

    def step(self, nan_action: bool = False):
        i = self.stats["steps"]
        actions = torch.randn(self.N, self.port.num_actions, generator=self.action_gen)
        if nan_action:
            actions[min(3, self.N - 1), 1] = float("nan")
        # (both sides get their own tensor: the within-limits action manager clamps its argument in place)
        self.last_action_args = (actions.clone(), actions.to(self.device))
        if self.philox:
            out_e = self.env.step(self.last_action_args[1])
            self.port.kernel_results = self._kernel_results()
            out_p = self.port.step(self.last_action_args[0])
            self._assert_margins(f"step {i}")
        else:
            out_p = self.port.step(self.last_action_args[0])
            self._assert_margins(f"step {i}")
            self._inject_from_log(self.port.reset_idx, self.port.resample_idx)
            out_e = self.env.step(self.last_action_args[1])
        where = f"step {i}"
        n_reset = int(self.port.reset_idx.numel())
        self._check(where, "reset_idx", self.env._fused.reset_idx[:n_reset], self.port.reset_idx, exact=True)
        if self.env._fused.report.n_reset != n_reset:
            raise ParityFailure(f"{where}: n_reset {self.env._fused.report.n_reset} != {n_reset}")
        self._check(where, "obs", out_e[0], out_p[0], exact=False)
        self._check(where, "rewards", out_e[1], out_p[1], exact=False)
        self._check(where, "terminated", out_e[2], out_p[2], exact=True)
        self._check(where, "truncated", out_e[3], out_p[3], exact=True)
        self._compare_all(where, out_p[4], out_e[4])
        self.stats["steps"] += 1
        self.stats["resets"] += n_reset
        return out_e, out_p

    def run(self, steps: int, nan_step: int | None = None):
        self.reset()
        for i in range(steps):
            self.step(nan_action=(nan_step is not None and i == nan_step))
        self.stats["guard"] = dict(self.sanitizer.stats)
        return self.stats
