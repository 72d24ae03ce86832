#!/usr/bin/env python
"""bench.py — env-steps/sec of the fused rsrl hot path on B200 (BASELINE.json metric).

Workload (config.workload): BASELINE configs[1] = MountainCar / Fourier(5)+bias / Q-learning / Greedy,
65 536 parallel envs per GPU, SHARED weights (one agent, dW summed over envs [and GPUs]), synthetic start
states x ~ U[-0.6,-0.4), episode cap 1000 with auto-reset.  One bench "step" = one call
rsrl_engine_step(K_INNER) = K_INNER batched steps of all envs.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
  torchrun ... bench.py --gpus N ...        (one rank per GPU, NCCL)

Prints ONE JSON line on rank 0 (see the driver contract in the task statement).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ENVS_PER_GPU = 65536
K_INNER = 2000                # batched steps per bench step
ALG_BYTES_PER_ENV_STEP = 48   # 16*D state r/w + 16 action/episode-counter r/w, D = 2, SHARED weights
METRIC = "env-steps/sec (N-env MountainCar QLearning) at 1/2/4/8 B200 vs CPU ref"
WORKLOAD = "cfg2: MountainCar QLearning Fourier(5)+bias Greedy, 65536 envs per GPU, SHARED weights (MEAN), cap 1000"


def make_cfg(abi, dtype, n_envs, env_offset=0, n_global=0, mode=None):
    return abi.default_config(
        n_envs=n_envs, env_offset=env_offset, n_envs_global=n_global or n_envs, dtype=dtype,
        weight_mode=abi.SHARED if mode is None else mode, update_scale=abi.SCALE_MEAN,
        init_mode=abi.INIT_UNIFORM, init_lo=[-0.6, 0.0], init_hi=[-0.4, 0.0], max_episode_steps=1000, seed=0)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for name, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def attach_peers(eng, dist, rank, world):
    """SHARED weights across GPUs: dW is summed inside the persistent kernel through NVLink peer mailboxes
    (cudaIpc handles gathered with torch.distributed); falls back to ncclAllReduce between per-step kernels."""
    from rsrl_b200.abi import RsrlError
    from rsrl_b200.engine import comm_unique_id
    ok, handles = 1, [None] * world
    try:
        dist.all_gather_object(handles, eng.peer_export())
        eng.peer_attach(handles, rank, world)
    except RsrlError as err:
        ok = 0
        print(f"[bench] rank {rank}: peer attach failed ({err}); falling back to NCCL", file=sys.stderr)
    import torch
    flag = torch.tensor([ok], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 1:
        return "in-kernel LL exchange over NVLink peer memory (rank-ordered sum)"
    uid = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    eng.comm_init(uid[0], rank, world)
    return "ncclAllReduce(dW) per step (per-step kernels)"


def run_reference(args, rank, world):
    """The reference's CPU implementation of the path (oracle port; rustc is unavailable),
    all host threads, bounded sample of the same workload.  Rank 0 only."""
    if rank != 0:
        return
    from rsrl_b200 import abi
    from oracle import pyoracle as O
    O.build()
    threads = os.cpu_count() or 1
    envs_per_thread, k = 64, 2000   # sample: threads*64 envs x 2000 steps per bench step (~0.4 s of CPU work each)
    cfg = make_cfg(abi, abi.F64, envs_per_thread, mode=abi.PER_ENV)
    for _ in range(args.warmup):
        O.baseline_run(cfg, threads, envs_per_thread, 50)
    secs, steps = 0.0, 0
    for _ in range(args.steps):
        s, n = O.baseline_run(cfg, threads, envs_per_thread, k)
        secs += s
        steps += n
    v = steps / secs
    sample = f"{threads} threads x {envs_per_thread} independent single-env reference-shaped agents x {k} steps per bench step"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import numpy as np
    import torch
    from rsrl_b200 import abi
    from rsrl_b200.engine import Engine, comm_unique_id

    assert torch.cuda.is_available(), "bench.py needs a CUDA device: rsrl_b200 has no CPU fallback"
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    dtype = abi.F32 if args.dtype == "f32" else abi.F64
    n_global = N_ENVS_PER_GPU * world
    cfg = make_cfg(abi, dtype, N_ENVS_PER_GPU, env_offset=rank * N_ENVS_PER_GPU, n_global=n_global)
    cfg.device = local_rank
    eng = Engine(cfg)
    exchange = "none"
    if world > 1:
        exchange = attach_peers(eng, dist, rank, world)
    stream = torch.cuda.ExternalStream(eng.stream(), device=torch.device("cuda", local_rank))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def timed_steps(n_steps, body):
        """device time of n_steps bench steps, L2 flushed between steps (outside the event pairs)"""
        total_ms = 0.0
        for _ in range(n_steps):
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            body()
            e1.record(stream)
            e1.synchronize()
            total_ms += e0.elapsed_time(e1)
        return total_ms

    # ---- device-resident throughput (`value`) ----
    for _ in range(args.warmup):
        eng.step(K_INNER)
    eng.sync()
    launches0 = eng.stats()["kernel_launches"]
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    ms = timed_steps(args.steps, lambda: eng.step(K_INNER))
    barrier()
    clocks = sampler.stop()
    launches = eng.stats()["kernel_launches"] - launches0
    eng.sync()

    # ---- end to end through the C ABI with HOST buffers (`e2e`) ----
    host_states = np.ascontiguousarray(eng.states())
    out_s, out_a, out_w = np.empty_like(host_states), np.empty(cfg.n_envs, dtype=np.int32), np.empty((36, 3))
    h2d = host_states.nbytes
    d2h = out_s.nbytes + out_a.nbytes + out_w.nbytes

    def e2e_body():
        eng.set_states(host_states)   # H2D: this step's inputs
        eng.step(K_INNER)
        eng.states(out_s)             # D2H: results
        eng.actions(out_a)
        eng.weights(out_w)

    e2e_body()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_body()
    barrier()
    e2e_s = time.perf_counter() - t0

    t = torch.tensor([ms, e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_s = float(t[0]), float(t[1])
    env_steps = float(n_global) * K_INNER * args.steps
    value = env_steps / (ms * 1e-3)

    peaks, peak_src = {}, "fallback"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak_src = "measured"
    except (OSError, ValueError):
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    ach_gbs = ALG_BYTES_PER_ENV_STEP * N_ENVS_PER_GPU * K_INNER * args.steps / (ms * 1e-3) / 1e9  # per GPU

    out = {
        "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 features/Q/weights + f64 physics" if dtype == abi.F32 else "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_envs_per_gpu": N_ENVS_PER_GPU, "batched_steps_per_bench_step": K_INNER,
                   "l2": "flushed (256 MiB write) between timed steps", "exchange": exchange},
        "clocks": clocks,
        "e2e": {"value": env_steps / e2e_s, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": ach_gbs, "peak": peak_gbs, "unit": "GB/s", "frac": ach_gbs / peak_gbs,
                     "traffic": 1.40e6, "peak_source": peak_src,
                     "note": "algorithmic 48 B/env-step x 65536 envs x 2000 steps per launch; traffic = dram bytes of one launch from "
                             "profiles/r01_final.md (env state is register-resident for the whole launch, so HBM is idle: the binding "
                             "roofs are instruction issue at 14 warps/SM and the per-step grid exchange, see DESIGN.md section 8)"},
    }

    # ---- the same workload in the other weight mode / dtype (extra information, not the headline) ----
    if world == 1:
        def side_run(label, **kw):
            c2 = make_cfg(abi, kw.get("dtype", dtype), N_ENVS_PER_GPU, mode=kw.get("mode"))
            with Engine(c2) as e2:
                st2 = torch.cuda.ExternalStream(e2.stream(), device=torch.device("cuda", local_rank))
                for _ in range(3):
                    e2.step(K_INNER)
                e2.sync()
                tot = 0.0
                for _ in range(5):
                    flush.zero_()
                    torch.cuda.synchronize()
                    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a0.record(st2)
                    e2.step(K_INNER)
                    a1.record(st2)
                    a1.synchronize()
                    tot += a0.elapsed_time(a1)
                return {"value": N_ENVS_PER_GPU * K_INNER * 5 / (tot * 1e-3), "unit": "env-steps/s", "ms_per_step": tot / 5, "what": label}
        pe = side_run("PER_ENV weights: 65536 independent reference agents, own W (36x3 fp32) resident in shared memory", mode=abi.PER_ENV)
        pe["roofline"] = {"bound": "hbm", "achieved": 624 * pe["value"] / 1e9, "peak": peak_gbs, "unit": "GB/s",
                          "frac": 624 * pe["value"] / 1e9 / peak_gbs,
                          "note": "algorithmic 624 B/env-step (SURVEY 8d: state + read W + write one column); > 1 because W never leaves the SM"}
        out["also"] = {"per_env_weights": pe,
                       "f64": side_run("SHARED weights, all arithmetic f64 (parity anchor dtype)", dtype=abi.F64)}

        # ---- the other BASELINE configs at their per-GPU shard size (parity-test cases; reported for context) ----
        def cfg_run(label, k, n_envs, **kw):
            base = dict(n_envs=n_envs, dtype=dtype, init_mode=abi.INIT_UNIFORM, seed=0, update_scale=abi.SCALE_MEAN)
            base.update(kw)
            with Engine(abi.default_config(**base)) as e3:
                st3 = torch.cuda.ExternalStream(e3.stream(), device=torch.device("cuda", local_rank))
                e3.step(max(3, k // 10))
                e3.sync()
                flush.zero_()
                torch.cuda.synchronize()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record(st3)
                e3.step(k)
                a1.record(st3)
                a1.synchronize()
                ms3 = a0.elapsed_time(a1)
                return {"value": n_envs * k / (ms3 * 1e-3), "unit": "env-steps/s", "us_per_batched_step": 1e3 * ms3 / k, "what": label}
        c3 = cfg_run("cfg3: CartPole SARSA TileCoding(8 tilings, 8 tiles/dim, 4096 rows) eps-greedy, 262144 envs, SHARED", 300, 262144,
                     domain=abi.CART_POLE, basis=abi.TILE_CODING, algo=abi.SARSA, policy=abi.EPSILON_GREEDY, epsilon=0.1, gamma=0.99,
                     lr=0.1 / 8, init_lo=[-0.05] * 4, init_hi=[0.05] * 4, max_episode_steps=500)
        c3["roofline"] = {"bound": "hbm", "achieved": 80 * c3["value"] / 1e9, "peak": peak_gbs, "unit": "GB/s", "frac": 80 * c3["value"] / 1e9 / peak_gbs,
                          "note": "algorithmic 80 B/env-step (SURVEY 8d); binding roof: instruction issue of the f64 RK4 + tile hashing, 4.7 k instructions per env-step (profiles/r01_final.md)"}
        c4 = cfg_run("cfg4 shard: Acrobot ExpectedSARSA Fourier(7)+bias (F=4096) eps-greedy, 131072 envs, SHARED, tcgen05 3xTF32 path", 40, 131072,
                     domain=abi.ACROBOT, basis_order=7, algo=abi.EXPECTED_SARSA, policy=abi.EPSILON_GREEDY, epsilon=0.1, gamma=0.99, lr=1e-4,
                     alpha=1.0, init_lo=[-0.1] * 4, init_hi=[0.1] * 4, max_episode_steps=500)
        tf32_peak = float(peaks.get("bf16_tflops", 1590.0)) / 2.0
        # per env-step: algorithmic 3 contractions x 2*4096*3; executed 3 TF32 passes x (2 evaluations x {re,im} x 192x64 MACs + {re,im} x 128x96 MACs) x 2
        alg_flop, exe_flop = 3 * 2 * 4096 * 3, 3 * (2 * 2 * 192 * 64 * 2 + 2 * 128 * 96 * 2)
        c4["roofline"] = {"bound": "tensor", "achieved": alg_flop * c4["value"] / 1e12, "peak": tf32_peak, "unit": "TFLOP/s",
                          "frac": alg_flop * c4["value"] / 1e12 / tf32_peak, "executed_tflops": exe_flop * c4["value"] / 1e12,
                          "note": "algorithmic 73.7 kFLOP/env-step (two Q = Phi W evaluations + dW = Phi^T D, SURVEY 8d); peak = measured bf16 "
                                  "cuBLAS TFLOP/s / 2 (kind::tf32 runs at half the bf16 rate); executed = 3xTF32 passes over the complex "
                                  "(real, imaginary) split, 6x the algorithmic count; see profiles/r01_f4tc.md"}
        c5 = cfg_run("cfg5 shard: MountainCar SARSA(lambda) replacing traces Fourier(5) eps-greedy, 32768 envs, per-env traces in shared memory", 1000, 32768,
                     algo=abi.SARSA_LAMBDA, policy=abi.EPSILON_GREEDY, epsilon=0.2, alpha=0.01, gamma=0.99, init_lo=[-0.6, 0.0],
                     init_hi=[-0.4, 0.0], max_episode_steps=1000)
        c5["roofline"] = {"bound": "hbm", "achieved": 912 * c5["value"] / 1e9, "peak": peak_gbs, "unit": "GB/s", "frac": 912 * c5["value"] / 1e9 / peak_gbs,
                          "note": "algorithmic 912 B/env-step (state + read/write 108 trace values); traces never leave shared memory"}
        out["also"].update({"cfg3_tile_coding": c3, "cfg4_fourier7_tensor_core": c4, "cfg5_sarsa_lambda": c5})

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import pyoracle as O
        O.build()
        threads = os.cpu_count() or 1
        ccfg = make_cfg(abi, abi.F64, 64, mode=abi.PER_ENV)
        O.baseline_run(ccfg, threads, 64, 100)
        secs, n = O.baseline_run(ccfg, threads, 64, 60000)   # ~10-15 s of CPU work on all host cores
        out["cpu_baseline"] = {"value": n / secs, "unit": "env-steps/s", "cores": threads, "kind": "port",
                               "sample": f"{threads} threads x 64 independent single-env reference-shaped agents x 60000 steps ({secs:.1f} s)"}
    if rank == 0:
        print(json.dumps(out))
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
