"""
The stand-alone entry points of the C ABI against the oracle: gfb_contact_forces (the reference's
Taichi kernel with its own argument list, contact/kernel.py:5-90) and gfb_rotate
(transform_by_quat(v, inv_quat(q)) of entity_manager.py:130-146 / utils.py:13-55).
"""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


@pytest.mark.parametrize("with_filter", [False, True])
@pytest.mark.parametrize("n_envs,n_slots", [(1, 1), (257, 8), (1000, 20)])
def test_contact_forces_entry_point(n_envs, n_slots, with_filter, cuda_device):
    from genesis_forge_b200 import _native as nat
    from oracle.contact_kernel import kernel_get_contact_forces

    lib, handle = nat.lib(), nat.Handle(n_envs, cuda_device.index or 0)
    g = torch.Generator().manual_seed(1000 * n_envs + n_slots + int(with_filter))
    L, targets, withs = 14, torch.tensor([3, 6, 9, 12], dtype=torch.int32), torch.tensor([0, 5], dtype=torch.int32)
    force = 30.0 * torch.randn(n_envs, n_slots, 3, generator=g)
    pos = torch.rand(n_envs, n_slots, 3, generator=g) * 2 - 1
    link_a = torch.randint(0, L, (n_envs, n_slots), generator=g, dtype=torch.int32)
    link_b = torch.randint(0, L, (n_envs, n_slots), generator=g, dtype=torch.int32)
    # padding slots the way get_contacts(as_tensor=True) returns them: zero force between link 0 and link 0
    pad = torch.rand(n_envs, n_slots, generator=g) < 0.3
    link_a[pad], link_b[pad], force[pad], pos[pad] = 0, 0, 0.0, 0.0
    quat = torch.randn(n_envs, L, 4, generator=g)
    quat = quat / quat.norm(dim=-1, keepdim=True)

    want_f, want_p = torch.zeros(n_envs, 4, 3), torch.zeros(n_envs, 4, 3)
    want_c = torch.zeros(n_envs, 4)
    kernel_get_contact_forces(force, pos, link_a, link_b, quat, targets, withs, want_f, want_p, want_c, int(with_filter))

    dev = cuda_device
    d = [t.to(dev).contiguous() for t in (force, pos, link_a, link_b, quat, targets, withs)]
    # the kernel writes every output element: garbage in, results out (the caller zero-fills nothing)
    out_f = torch.full((n_envs, 4, 3), float("nan"), device=dev)
    out_p = torch.full((n_envs, 4, 3), float("nan"), device=dev)
    out_c = torch.full((n_envs, 4), float("nan"), device=dev)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    handle.check(
        lib.gfb_contact_forces(handle.ptr, *(_ptr(t) for t in d), _ptr(out_f), _ptr(out_p), _ptr(out_c),
                               n_envs, n_slots, L, 4, 2, int(with_filter), stream),
        "gfb_contact_forces",
    )
    # ordered accumulation on both sides: bit-exact
    assert torch.equal(out_c.cpu(), want_c)
    assert torch.equal(out_f.cpu(), want_f)
    assert torch.equal(out_p.cpu(), want_p)
    assert n_envs < 100 or bool((want_c > 0).any())  # the larger cases do contain hits


@pytest.mark.parametrize("n", [1, 33, 4097])
def test_rotate_entry_point(n, cuda_device):
    from genesis_forge_b200 import _native as nat
    from oracle.geom import inv_quat, transform_by_quat

    lib, handle = nat.lib(), nat.Handle(max(n, 1), cuda_device.index or 0)
    g = torch.Generator().manual_seed(n)
    vec = torch.randn(n, 3, generator=g)
    quat = torch.randn(n, 4, generator=g)
    quat = quat / quat.norm(dim=-1, keepdim=True)
    stream = C.c_void_p(torch.cuda.current_stream(cuda_device).cuda_stream)
    dv, dq = vec.to(cuda_device), quat.to(cuda_device)
    out = torch.empty(n, 3, device=cuda_device)
    # conjugate = 1: rotate by the inverse of q (what the uncached getters do, utils.py:23-24)
    handle.check(lib.gfb_rotate(handle.ptr, _ptr(dv), _ptr(dq), _ptr(out), n, 1, stream), "gfb_rotate")
    assert torch.equal(out.cpu(), transform_by_quat(vec, inv_quat(quat)))
    # conjugate = 0 with an already inverted quaternion (the cached path, entity_manager.py:134)
    diq = inv_quat(quat).to(cuda_device).contiguous()
    handle.check(lib.gfb_rotate(handle.ptr, _ptr(dv), _ptr(diq), _ptr(out), n, 0, stream), "gfb_rotate")
    assert torch.equal(out.cpu(), transform_by_quat(vec, inv_quat(quat)))
    # vec == NULL: projected gravity, R(q)^T (0, 0, -1)
    handle.check(lib.gfb_rotate(handle.ptr, None, _ptr(dq), _ptr(out), n, 1, stream), "gfb_rotate")
    gravity = torch.tensor([0.0, 0.0, -1.0]).expand(n, 3)
    assert torch.equal(out.cpu(), transform_by_quat(gravity, inv_quat(quat)))


def test_reset_rows_entry_point(cuda_device):
    """
    gfb_reset_rows: base + U(-1,1) * scale rows with injected draws are bit-equal
    to the reference's two torch ops (position_action_manager.py:516-525); U(a, b) rows are scattered
    into the per-env buffer at the given env ids; in-kernel Philox draws stay inside their range.
    """
    import genesis_forge_b200 as gfb
    from configs import specs
    from configs.env_builder import build_env, dropin_namespace
    from genesis_forge_b200.rng import ReplayRng

    gfb.set_device(cuda_device)
    env = build_env(specs.get("command_direction"), dropin_namespace(), 256, cuda_device, seed=3, n_contacts=0)
    env.build()
    fused = env._fused
    gen = torch.Generator().manual_seed(0)
    base = torch.randn(12, generator=gen).to(cuda_device)
    idx = torch.tensor([3, 7, 100, 255], device=cuda_device)
    u = (torch.rand(4, 12, generator=gen) * 2 - 1).to(cuda_device)
    env.rng = ReplayRng()
    env.rng.push("t", u)
    got = fused.reset_rows("noise", "t", idx, 4, 12, 0.05, base=base)
    assert torch.equal(got, base + u * 0.05)
    buf = torch.zeros((256, 5), device=cuda_device)
    vals = torch.rand(4, 5, generator=gen).to(cuda_device)
    env.rng.push("m", vals)
    out = fused.reset_rows("uniform", "m", idx, 4, 5, -0.2, 0.2, scatter=buf)
    assert torch.equal(out, vals) and torch.equal(buf[idx], vals) and int((buf != 0).any(dim=1).sum()) == 4
    from genesis_forge_b200.rng import HostRng

    env.rng = HostRng()
    big = torch.arange(256, device=cuda_device)
    draws = fused.reset_rows("uniform", "m", big, 256, 8, -0.2, 0.2)
    assert float(draws.min()) >= -0.2 and float(draws.max()) <= 0.2 and float(draws.std()) > 0.08
    noisy = fused.reset_rows("noise", "t", big, 256, 12, 0.05, base=base)
    assert float((noisy - base).abs().max()) <= 0.05 * (1 + 1e-6) and float((noisy - base).std()) > 0.02


def test_rsl_rl_wrapper_glue(cuda_device):
    """
    RslRlWrapper (wrappers/rsl_rl.py:41-72): step returns (obs, rewards, dones, extras) with
    dones == terminated | truncated -- here the kernel's own GFB_B_DONES output, no extra launch --
    and the critic observations default to the policy's.
    """
    import genesis_forge_b200 as gfb
    from configs import specs
    from configs.env_builder import build_env, dropin_namespace
    from genesis_forge_b200.wrappers import RslRlWrapper

    gfb.set_device(cuda_device)
    env = build_env(specs.get("command_direction"), dropin_namespace(), 2048, cuda_device, seed=5, n_contacts=0)
    wrapped = RslRlWrapper(env)
    wrapped.build()
    wrapped.reset()
    fired = 0
    for _ in range(5):
        launches = env._fused.launch_count()
        obs, rew, dones, extras = wrapped.step(torch.randn(2048, 12, device=cuda_device))
        term, trunc = env.termination_manager.terminated, env.termination_manager.truncated
        assert dones.dtype == torch.bool and torch.equal(dones, term | trunc)
        assert torch.equal(extras["time_outs"], trunc)
        assert extras["observations"]["critic"] is not None and obs is not None
        assert env._fused.launch_count() - launches <= 4  # action, post-physics, index compaction, re-observation
        fired += int(dones.sum())
    assert fired > 0
    with pytest.raises(AssertionError):
        RslRlWrapper(wrapped)  # can_be_wrapped = False


@pytest.mark.gpu
@pytest.mark.parametrize("tile,num_envs", [(32, 200_000), (64, 400_000), (128, 400_000)])
def test_persistent_loop_with_short_slabs(tile, num_envs):
    """The slab loop with the shortest possible iteration (entity phase alone), forced for every slab
    size (GFB_DEBUG=8), generic kernel: the launch shape that failed with "unspecified launch failure"
    before the uniform-datapath rule (csrc/device_utils.cuh).  Own process: a faulting kernel would
    poison this one's CUDA context."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, GFB_DEBUG="8", GFB_NO_SPEC="1", GFB_TILE=str(tile))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "loop_stress.py"), str(num_envs), "20"],
                         capture_output=True, text=True, env=env, timeout=600, cwd=root)
    assert out.returncode == 0 and "ok, copies equal: True" in out.stdout, out.stdout[-400:] + out.stderr[-400:]


@pytest.mark.parametrize("num_envs", [300, 4096 + 4, 333])
def test_action_step_ring_equals_the_copying_entry_point(num_envs, cuda_device):
    """gfb_action_step_ring (the caller exchanges its two action buffers, nothing is copied into
    last_actions) leaves the same values in env.actions / env.last_actions / targets / action-rate /
    episode_length as gfb_action_step (genesis_env.py:195-203 with the copy), TMA and fallback paths."""
    import genesis_forge_b200 as gfb
    from configs import specs
    from configs.env_builder import build_env, dropin_namespace

    dev = cuda_device
    gfb.set_device(dev)
    env = build_env(specs.get("command_direction"), dropin_namespace(), num_envs, dev, pool=2, seed=7)
    env.build()
    env.reset()
    fused, action = env._fused, env.managers["action"]
    g = torch.Generator().manual_seed(3)
    for _ in range(2):  # fill both buffers of the ring with history
        env.step(torch.randn(num_envs, fused.D, generator=g).to(dev))
    a0, l0, ep0 = env.actions.clone(), env.last_actions.clone(), env.episode_length.clone()
    raw = torch.randn(num_envs, fused.D, generator=g).to(dev)

    fused.action_step(raw)  # ring: exchanges env._actions / env._last_actions, then launches
    torch.cuda.synchronize(dev)
    ring = [t.clone() for t in (env.actions, env.last_actions, action._actions, fused.action_rate, env.episode_length)]
    assert torch.equal(ring[0], raw) and torch.equal(ring[1], a0)
    assert torch.equal(ring[4], ep0 + 1)

    # the same inputs through the copying entry point
    env._actions.copy_(a0)
    env._last_actions.copy_(l0)
    env.episode_length.copy_(ep0)
    action._actions.fill_(float("nan"))
    fused.action_rate.fill_(float("nan"))
    fused.bind_action_buffers()
    rc = fused.lib.gfb_action_step(fused._h, fused._buffers_ref, raw.data_ptr(), raw.data_ptr(), fused._stream())
    assert rc == 0
    torch.cuda.synchronize(dev)
    copy = (env.actions, env.last_actions, action._actions, fused.action_rate, env.episode_length)
    for r, c, what in zip(ring, copy, ("actions", "last_actions", "targets", "action_rate", "episode_length")):
        assert torch.equal(r, c), what
