"""
Sharded-env logging on real GPUs (run under torchrun, world >= 2): the in-kernel peer-memory exchange
must publish the same keys and values as the NCCL all-reduce path, and both must equal what the
ranks' local masks say globally.  Prints one line per rank; exit code 1 on mismatch.
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from bench import make_dropin_env
from configs import specs

rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
name = sys.argv[3] if len(sys.argv) > 3 else "command_direction"


def run(peer: bool):
    env = make_dropin_env(specs.get(name), n, dev, 4, 1234 + rank)
    env._fused.shard(dist.group.WORLD, n * world, peer=peer)
    assert env._fused.peer_mode == peer
    gen = torch.Generator().manual_seed(5 + rank)
    log = []
    for i in range(steps):
        actions = torch.randn(n, env._fused.D, generator=gen).to(dev)
        obs, rew, term, trunc, extras = env.step(actions)
        counts = torch.stack([term.sum(), (term | trunc).sum()]).double()
        dist.all_reduce(counts)
        entry = {k: float(v) for k, v in extras["episode"].items()}
        entry["_global_fall_fraction"] = float(counts[0]) / (n * world)
        entry["_global_resets"] = float(counts[1])
        entry["_report_resets"] = float(env._fused.report.global_n_reset if peer else env._fused.global_acc[-1])
        log.append(entry)
    return log


peer_log, nccl_log = run(True), run(False)
ok = True
for i, (a, b) in enumerate(zip(peer_log, nccl_log)):
    if set(a) != set(b):
        print(f"rank {rank} step {i}: keys differ {sorted(set(a) ^ set(b))}"); ok = False; continue
    for k in a:
        if abs(a[k] - b[k]) > 1e-6 * max(1.0, abs(b[k])):
            print(f"rank {rank} step {i}: {k} peer {a[k]!r} nccl {b[k]!r}"); ok = False
    if a["_report_resets"] != a["_global_resets"]:
        print(f"rank {rank} step {i}: global reset count {a['_report_resets']} != {a['_global_resets']}"); ok = False
    key = "Terminations / fall_over"
    if key in a and abs(a[key] - a["_global_fall_fraction"]) > 1e-7:
        print(f"rank {rank} step {i}: {key} {a[key]} != {a['_global_fall_fraction']}"); ok = False
gathered = [None] * world
dist.all_gather_object(gathered, peer_log)
if any(g != gathered[0] for g in gathered):  # every rank publishes the same global numbers, bit for bit
    print(f"rank {rank}: ranks disagree on the published values"); ok = False
n_keys = sum(len(e) - 3 for e in peer_log)
print(f"rank {rank}/{world}: {'PEER LOGGING OK' if ok else 'MISMATCH'} ({steps} steps, {n_keys} logged values compared)")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
