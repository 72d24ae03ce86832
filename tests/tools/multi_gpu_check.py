"""torchrun --nproc-per-node N tests/tools/multi_gpu_check.py : SHARED-weights engines sharded over N GPUs.

  fp32 (bench dtype): every rank's states / actions / episode hashes and the weights must equal oracle32's replay of the
      same world (oracle/oracle32.cpp, world = N) BIT FOR BIT — the in-kernel NVLink exchange included;
  f64: agreement with ONE engine that runs all envs (actions / step counts exact, weights 1e-10: the sums associate differently);
  every case: bit-identical W replicas.
Covers BASELINE configs[1] (Q-learning), configs[4] (per-env traces, in-kernel exchange) and configs[3] (tensor-core path +
ncclAllReduce).  Prints MULTI_GPU_CHECK PASS / FAIL on rank 0."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import torch.distributed as dist
from rsrl_b200 import abi
from rsrl_b200.engine import Engine, comm_unique_id

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True


def gather(arr):
    t = torch.from_numpy(np.ascontiguousarray(arr)).cuda()
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [o.cpu().numpy() for o in out]


def attach(eng):
    handles = [None] * world
    dist.all_gather_object(handles, eng.peer_export())
    eng.peer_attach(handles, rank, world)
    dist.barrier()


MC = dict(init_mode=abi.INIT_UNIFORM, init_lo=[-0.6, 0.0], init_hi=[-0.4, 0.0], max_episode_steps=60, seed=9, update_scale=abi.SCALE_MEAN)
CASES = [
    ("cfg2 qlearning eps", dict(MC, policy=abi.EPSILON_GREEDY, epsilon=0.1, lr=0.05), [(1500, 90), (65536, 60)]),
    ("cfg5 sarsa(lambda)", dict(MC, algo=abi.SARSA_LAMBDA, policy=abi.EPSILON_GREEDY, epsilon=0.2, alpha=0.01, gamma=0.99), [(777, 60), (32768, 40)]),
]
for label, kw, sizes in CASES:
    for n_rank, steps in sizes:
        n_global = n_rank * world
        # ---- fp32: bit-exact against oracle32's replay of the whole world ----
        cfg = abi.default_config(n_envs=n_rank, env_offset=rank * n_rank, n_envs_global=n_global, device=local, dtype=abi.F32, **kw)
        eng = Engine(cfg)
        attach(eng)
        for k in (1, steps - 1):     # two launches: the exchange state has to carry over
            eng.step(k)
        eng.sync()
        Ws, Ss, As, Hs = gather(eng.weights()), gather(eng.states()), gather(eng.actions()), gather(eng.env_stats()[2].astype(np.int64))
        replicas = all((w == Ws[0]).all() for w in Ws)
        if rank == 0:
            from oracle import pyoracle32 as O32
            c0 = abi.default_config(n_envs=n_rank, env_offset=0, n_envs_global=n_global, dtype=abi.F32, **kw)
            o = O32.Engine(c0, eng.launch_shape(), world=world)
            o.step(steps)
            exact = (o.weights() == Ws[0]).all() and all((o.states(r) == Ss[r]).all() and (o.actions(r) == As[r]).all()
                                                         and (o.env_stats(r)[2].astype(np.int64) == Hs[r]).all() for r in range(world))
            good = bool(replicas and exact)
            ok &= good
            print(f"{label} f32 N={n_global} world={world} steps={steps}: replicas_identical={replicas} bit_exact_vs_oracle32={bool(exact)} "
                  f"launches={eng.stats()['kernel_launches']} -> {'OK' if good else 'FAIL'}", flush=True)
        # reset on peer-attached engines needs no barrier (the exchange epoch is not reset): run again and compare with the first run
        eng.reset()
        eng.step(steps)
        eng.sync()
        again = bool((eng.weights() == Ws[rank]).all())
        flags = gather(np.array([int(again)]))
        if rank == 0:
            good = all(int(f[0]) for f in flags)
            ok &= good
            print(f"{label} f32 reset + rerun reproduces the first run on every rank: {good}", flush=True)
        eng.close()
        dist.barrier()
        # ---- f64: against one engine that runs every env ----
        if n_rank > 4096:
            continue
        cfg = abi.default_config(n_envs=n_rank, env_offset=rank * n_rank, n_envs_global=n_global, device=local, dtype=abi.F64, **kw)
        eng = Engine(cfg)
        attach(eng)
        eng.step(steps)
        eng.sync()
        Ws, As = gather(eng.weights()), gather(eng.actions())
        if rank == 0:
            single = Engine(abi.default_config(n_envs=n_global, device=local, dtype=abi.F64, **kw))
            single.step(steps)
            single.sync()
            werr = np.abs(Ws[0] - single.weights()).max() / max(1.0, np.abs(single.weights()).max())
            good = bool(all((w == Ws[0]).all() for w in Ws) and (np.concatenate(As) == single.actions()).all() and werr < 1e-10)
            ok &= good
            print(f"{label} f64 N={n_global} world={world}: actions equal single GPU, |W - W_single| = {werr:.2e} -> {'OK' if good else 'FAIL'}", flush=True)
            single.close()
        eng.close()
        dist.barrier()

# BASELINE configs[3]: Acrobot / ExpectedSARSA / Fourier(7) on the tensor-core path, dW (4096 x 3) summed with ncclAllReduce
for n_rank, steps in ((1500, 6), (131072, 10)):
    n_global = n_rank * world
    kw = dict(domain=abi.ACROBOT, basis_order=7, algo=abi.EXPECTED_SARSA, policy=abi.EPSILON_GREEDY, epsilon=0.1, dtype=abi.F32,
              init_mode=abi.INIT_UNIFORM, init_lo=[-0.1] * 4, init_hi=[0.1] * 4, max_episode_steps=500, seed=3, gamma=0.99, lr=1e-3,
              alpha=1.0, update_scale=abi.SCALE_MEAN)
    eng = Engine(abi.default_config(n_envs=n_rank, env_offset=rank * n_rank, n_envs_global=n_global, device=local, **kw))
    uid = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    eng.comm_init(uid[0], rank, world)
    dist.barrier()
    eng.step(2); eng.sync(); dist.barrier()
    t0 = time.perf_counter(); eng.step(steps - 2); eng.sync(); dist.barrier(); dt = time.perf_counter() - t0
    Ws = gather(eng.weights())
    if rank == 0:
        good = all((w == Ws[0]).all() for w in Ws)
        msg = ""
        if n_global <= 16384:
            single = Engine(abi.default_config(n_envs=n_global, device=local, **kw))
            single.step(steps); single.sync()
            werr = np.abs(Ws[0] - single.weights()).max() / max(np.abs(single.weights()).max(), 1e-30)
            good &= werr < 1e-4
            msg = f" |W - W_single|/|W|max={werr:.3e}"
            single.close()
        else:
            msg = f" {1e6 * dt / (steps - 2):.1f} us/step  {n_global * (steps - 2) / dt / 1e9:.3f} G env-steps/s"
        ok &= bool(good)
        print(f"cfg4 f32 N={n_global} world={world}: replicas_identical={bool(good)}{msg} -> {'OK' if good else 'FAIL'}", flush=True)
    eng.close()
    dist.barrier()
if rank == 0:
    print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL")
dist.destroy_process_group()
