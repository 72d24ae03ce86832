"""torchrun --nproc-per-node N tools/multi_probe.py [ngroups:delay_ns[:world_poll_backoff_ns] ...] : us per batched step of the cfg2 kernel sharded over N GPUs
(in-kernel NVLink exchange) for exchange settings; RSRL_B200_PHASE_PROFILE=1 prints rank 0's phase table per setting."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from rsrl_b200 import abi
from rsrl_b200.engine import Engine

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = int(os.environ.get("N", 65536))
for setting in (sys.argv[1:] or ["8:400"]):
    ng, delay, wb = (setting.split(":") + ["0"])[:3]
    os.environ["RSRL_B200_NGROUPS"], os.environ["RSRL_B200_POLL_DELAY"], os.environ["RSRL_B200_WORLD_BACKOFF"] = ng, delay, wb
    cfg = abi.default_config(n_envs=N, env_offset=rank * N, n_envs_global=N * world, dtype=abi.F32, init_mode=abi.INIT_UNIFORM,
                             init_lo=[-0.6, 0.0], init_hi=[-0.4, 0.0], max_episode_steps=1000, seed=0, update_scale=abi.SCALE_MEAN)
    cfg.device = local
    if rank != 0:
        os.environ.pop("RSRL_B200_PHASE_PROFILE", None)
    e = Engine(cfg)
    if world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, e.peer_export())
        e.peer_attach(handles, rank, world)
    dist.barrier()
    e.step(500); e.sync(); dist.barrier()
    t0 = time.perf_counter(); e.step(4000); e.sync(); dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"world={world} ngroups={ng} delay={delay} world_backoff={wb}: {1e6 * float(t[0]) / 4000:.2f} us/step  {N * world * 4000 / float(t[0]) / 1e9:.2f} G env-steps/s", flush=True)
    dist.barrier()
    e.close()
    dist.barrier()
dist.destroy_process_group()
