"""world_size-2 (gloo, CPU) test of the N>1 path's host logic: env sharding by contiguous global id ranges,
RNG keyed by the global env id, MEAN scaling by the global env count and the per-step exchange of dW
(allreduce SUM) — the same decomposition bench.py uses with NCCL.  The compute on each rank is the CPU oracle."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_global, steps, kw, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import pyoracle as O
    from rsrl_b200 import abi
    from rsrl_b200.sharding import shard_range

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = shard_range(n_global, rank, world)
    cfg = abi.default_config(n_envs=hi - lo, env_offset=lo, n_envs_global=n_global, **kw)
    e = O.Engine(cfg)
    for _ in range(steps):
        g = torch.from_numpy(e.step_local())
        if cfg.weight_mode == abi.SHARED:
            dist.all_reduce(g, op=dist.ReduceOp.SUM)   # the one data-path collective
        e.step_apply(g.numpy())
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), states=e.states(), actions=e.actions(), weights=e.weights(),
             hash=e.env_stats()[2], lo=lo, hi=hi)
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["shared", "per_env", "shared_pal_softmax"])
def test_two_rank_sharding_matches_single_rank(oracle, tmp_path, mode):
    import torch.multiprocessing as mp
    from rsrl_b200 import abi
    kw = dict(dtype=abi.F64, init_mode=abi.INIT_UNIFORM, init_lo=[-0.6, 0.0], init_hi=[-0.4, 0.0], max_episode_steps=40,
              seed=5, policy=abi.EPSILON_GREEDY, epsilon=0.2, update_scale=abi.SCALE_MEAN,
              weight_mode=abi.PER_ENV if mode == "per_env" else abi.SHARED)
    if mode == "shared_pal_softmax":   # PAL agent (control/td/pal.rs) behind a Softmax policy (tau = 0.5 in the epsilon field)
        kw.update(algo=abi.PAL, policy=abi.SOFTMAX, epsilon=0.5, alpha=0.5, lr=0.01)
    n_global, steps, world = 37, 60, 2   # ragged split: 19 + 18 envs
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_global, steps, kw, str(tmp_path)), nprocs=world, join=True)

    single = oracle.Engine(abi.default_config(n_envs=n_global, **kw))
    single.step(steps)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    assert [(int(p["lo"]), int(p["hi"])) for p in parts] == [(0, 19), (19, 37)]
    states = np.concatenate([p["states"] for p in parts])
    actions = np.concatenate([p["actions"] for p in parts])
    hashes = np.concatenate([p["hash"] for p in parts])
    assert (actions == single.actions()).all() and (hashes == single.env_stats()[2]).all()
    assert np.abs(states - single.states()).max() < 1e-12
    if mode != "per_env":
        # both ranks hold the same replicated W; it equals the single-rank W up to summation order
        assert (parts[0]["weights"] == parts[1]["weights"]).all()
        assert np.abs(parts[0]["weights"] - single.weights()).max() < 1e-12
    else:
        assert np.abs(np.concatenate([p["weights"] for p in parts]) - single.weights()).max() == 0.0


def test_shard_range_covers_everything():
    from rsrl_b200.sharding import shard_range
    for n, w in [(1, 1), (37, 2), (65536, 8), (1000003, 8), (5, 8)]:
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
