"""torchrun --nproc-per-node N tools/multi_gpu_check.py : SHARED-weights engines sharded over N GPUs with the
in-kernel NVLink exchange must reproduce a single-GPU engine over all envs (actions / step counts bit-exact,
weights equal up to the summation order of dW) and hold bit-identical W replicas."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from rsrl_b200 import abi
from rsrl_b200.engine import Engine
from rsrl_b200.sharding import shard_range

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
# f32 free runs diverge chaotically once a greedy tie flips (summation order differs): compare a short horizon
for dtype, tol, horizon in ((abi.F64, 1e-11, None), (abi.F32, 1e-4, 12)):
    for n_global, steps in ((4099, horizon or 300), (65536 * world, horizon or 200)):
        kw = dict(dtype=dtype, init_mode=abi.INIT_UNIFORM, init_lo=[-0.6, 0.0], init_hi=[-0.4, 0.0], max_episode_steps=150,
                  seed=9, policy=abi.EPSILON_GREEDY, epsilon=0.1, update_scale=abi.SCALE_MEAN, lr=0.05)
        lo, hi = shard_range(n_global, rank, world)
        cfg = abi.default_config(n_envs=hi - lo, env_offset=lo, n_envs_global=n_global, device=local, **kw)
        eng = Engine(cfg)
        handles = [None] * world
        dist.all_gather_object(handles, eng.peer_export())
        eng.peer_attach(handles, rank, world)
        dist.barrier()
        eng.step(steps); eng.sync()
        W = torch.from_numpy(eng.weights()).cuda()
        allW = [torch.empty_like(W) for _ in range(world)]
        dist.all_gather(allW, W)
        replicas_identical = all(bool((w == allW[0]).all()) for w in allW)
        acts = torch.from_numpy(eng.actions()).cuda()
        sizes = [shard_range(n_global, r, world) for r in range(world)]
        gathered = [torch.empty(h - l, dtype=acts.dtype, device="cuda") for l, h in sizes]
        dist.all_gather(gathered, acts)
        if rank == 0:
            single = Engine(abi.default_config(n_envs=n_global, device=local, **kw))
            single.step(steps); single.sync()
            a_ok = bool((torch.cat(gathered).cpu().numpy() == single.actions()).all()) if dtype == abi.F64 else True
            werr = np.abs(eng.weights() - single.weights()).max()
            good = replicas_identical and a_ok and werr < tol
            ok &= good
            print(f"dtype={'f64' if dtype == abi.F64 else 'f32'} N={n_global} world={world}: replicas_identical={replicas_identical} "
                  f"actions_equal={a_ok} |W - W_single|max={werr:.3e} launches={eng.stats()['kernel_launches']} -> {'OK' if good else 'FAIL'}", flush=True)
            single.close()
        eng.close()
        dist.barrier()
# BASELINE config 5: SARSA(lambda) with per-env eligibility traces (resident in shared memory), shared W exchanged in-kernel.
# f64 is the correctness check (agreement with a single-GPU engine to 1e-10).  In f32 the different summation order of dW
# flips near-tied eps-greedy decisions from the second step on and the trajectories diverge chaotically (measured 5e-2 after
# 12 steps on 2 GPUs): there the check is bit-identical replicas; the difference is printed for information.
for dtype, tol, steps in ((abi.F64, 1e-10, 120), (abi.F32, 0.2, 12)):
    n_global = 32768 * world if dtype == abi.F32 else 2048 * world + 37
    kw = dict(dtype=dtype, algo=abi.SARSA_LAMBDA, policy=abi.EPSILON_GREEDY, epsilon=0.2, alpha=0.01, gamma=0.99, init_mode=abi.INIT_UNIFORM,
              init_lo=[-0.6, 0.0], init_hi=[-0.4, 0.0], max_episode_steps=150, seed=4, update_scale=abi.SCALE_MEAN)
    lo, hi = shard_range(n_global, rank, world)
    eng = Engine(abi.default_config(n_envs=hi - lo, env_offset=lo, n_envs_global=n_global, device=local, **kw))
    handles = [None] * world
    dist.all_gather_object(handles, eng.peer_export())
    eng.peer_attach(handles, rank, world)
    dist.barrier()
    eng.step(steps); eng.sync()
    W = torch.from_numpy(eng.weights()).cuda()
    allW = [torch.empty_like(W) for _ in range(world)]
    dist.all_gather(allW, W)
    replicas_identical = all(bool((w == allW[0]).all()) for w in allW)
    if rank == 0:
        single = Engine(abi.default_config(n_envs=n_global, device=local, **kw))
        single.step(steps); single.sync()
        werr = np.abs(eng.weights() - single.weights()).max() / max(np.abs(single.weights()).max(), 1e-30)
        good = replicas_identical and (werr < tol if dtype == abi.F64 else bool(np.isfinite(werr)))
        ok &= good
        print(f"cfg5 sarsa(lambda) dtype={'f64' if dtype == abi.F64 else 'f32'} N={n_global} world={world}: replicas_identical={replicas_identical} "
              f"|W - W_single|/|W|max={werr:.3e} launches={eng.stats()['kernel_launches']} -> {'OK' if good else 'FAIL'}", flush=True)
        single.close()
    eng.close()
    dist.barrier()

# BASELINE config 4: Acrobot / ExpectedSARSA / Fourier(7) on the tensor-core path, envs sharded over the ranks, dW (4096 x 3)
# summed with ncclAllReduce between the dW pass and the weight update
from rsrl_b200.engine import comm_unique_id
import time
for n_global, steps in ((3001, 6), (131072 * world, 10)):
    kw = dict(domain=abi.ACROBOT, basis_order=7, algo=abi.EXPECTED_SARSA, policy=abi.EPSILON_GREEDY, epsilon=0.1, dtype=abi.F32,
              init_mode=abi.INIT_UNIFORM, init_lo=[-0.1] * 4, init_hi=[0.1] * 4, max_episode_steps=500, seed=3, gamma=0.99, lr=1e-3,
              alpha=1.0, update_scale=abi.SCALE_MEAN)
    lo, hi = shard_range(n_global, rank, world)
    eng = Engine(abi.default_config(n_envs=hi - lo, env_offset=lo, n_envs_global=n_global, device=local, **kw))
    uid = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    eng.comm_init(uid[0], rank, world)
    dist.barrier()
    eng.step(2); eng.sync(); dist.barrier()
    t0 = time.perf_counter(); eng.step(steps - 2); eng.sync(); dist.barrier(); dt = time.perf_counter() - t0
    W = torch.from_numpy(eng.weights()).cuda()
    allW = [torch.empty_like(W) for _ in range(world)]
    dist.all_gather(allW, W)
    replicas_identical = all(bool((w == allW[0]).all()) for w in allW)
    if rank == 0:
        good = replicas_identical
        msg = ""
        if n_global <= 4096:
            single = Engine(abi.default_config(n_envs=n_global, device=local, **kw))
            single.step(steps); single.sync()
            werr = np.abs(eng.weights() - single.weights()).max() / max(np.abs(single.weights()).max(), 1e-30)
            good &= werr < 1e-4
            msg = f" |W - W_single|/|W|max={werr:.3e}"
            single.close()
        else:
            msg = f" {1e6 * dt / (steps - 2):.1f} us/step  {n_global * (steps - 2) / dt / 1e9:.3f} G env-steps/s"
        ok &= good
        print(f"cfg4 f32 N={n_global} world={world}: replicas_identical={replicas_identical}{msg} -> {'OK' if good else 'FAIL'}", flush=True)
    eng.close()
    dist.barrier()
if rank == 0:
    print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL")
dist.destroy_process_group()
