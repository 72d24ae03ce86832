#!/usr/bin/env python
"""bench.py — env-steps/sec of the fused rsrl hot path on B200 (BASELINE.json metric).

Workload (config.workload), default = BASELINE configs[1]: MountainCar / Fourier(5)+bias / Q-learning / Greedy,
65 536 parallel envs per GPU, SHARED weights (one agent, dW summed over envs [and GPUs]), synthetic start
states x ~ U[-0.6,-0.4), episode cap 1000 with auto-reset.  One bench "step" = one call
rsrl_engine_step(K_INNER) = K_INNER batched steps of all envs.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--config cfg2|cfg4|cfg5]
  torchrun ... bench.py --gpus N ...        (one rank per GPU)

--config cfg4 / cfg5 print the same kind of line for BASELINE configs[3] / configs[4] (their per-GPU shard x N GPUs).
Prints ONE JSON line on rank 0 (see the driver contract in the task statement).  At N > 1 the line carries
`replicas_identical` and `matches_single_gpu`; the process exits non-zero when either is false.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env-steps/sec (N-env MountainCar QLearning) at 1/2/4/8 B200 vs CPU ref"
# instruction count of the dominant kernel, measured with ncu (profiles/r02_persistent.md): warp-level instructions executed
# per env-step by persistent_kernel<float, MountainCar, Fourier, 5, 3, SHARED>; and its DRAM traffic per 2000-step launch
WARP_INST_PER_ENV_STEP = None
DRAM_BYTES_PER_LAUNCH = None
try:
    _m = json.load(open(os.path.join(ROOT, "profiles", "r02_measured_constants.json")))
    WARP_INST_PER_ENV_STEP = _m.get("warp_inst_per_env_step")
    DRAM_BYTES_PER_LAUNCH = _m.get("dram_bytes_per_2000_step_launch")
except (OSError, ValueError):
    pass


def workloads(abi):
    mc = dict(init_mode=abi.INIT_UNIFORM, init_lo=[-0.6, 0.0], init_hi=[-0.4, 0.0], max_episode_steps=1000, seed=0,
              update_scale=abi.SCALE_MEAN)
    return {
        "cfg2": dict(label="cfg2: MountainCar QLearning Fourier(5)+bias Greedy, 65536 envs per GPU, SHARED weights (MEAN), cap 1000",
                     n=65536, k_inner=2000, bytes=48, F=36, A=3, kw=mc),
        "cfg5": dict(label="cfg5: MountainCar SARSA(lambda) replacing traces Fourier(5)+bias eps-greedy 0.2, 32768 envs per GPU, SHARED weights "
                           "(MEAN), per-env traces resident in shared memory, cap 1000",
                     n=32768, k_inner=1000, bytes=912, F=36, A=3,
                     kw=dict(mc, algo=abi.SARSA_LAMBDA, policy=abi.EPSILON_GREEDY, epsilon=0.2, alpha=0.01, gamma=0.99)),
        "cfg4": dict(label="cfg4: Acrobot ExpectedSARSA Fourier(7)+bias (F=4096) eps-greedy 0.1, 131072 envs per GPU, SHARED weights (MEAN), "
                           "tcgen05 3xTF32 path, cap 500",
                     n=131072, k_inner=20, bytes=80, F=4096, A=3,
                     kw=dict(domain=abi.ACROBOT, basis_order=7, algo=abi.EXPECTED_SARSA, policy=abi.EPSILON_GREEDY, epsilon=0.1, gamma=0.99,
                             lr=1e-4, alpha=1.0, init_mode=abi.INIT_UNIFORM, init_lo=[-0.1] * 4, init_hi=[0.1] * 4, max_episode_steps=500,
                             seed=0, update_scale=abi.SCALE_MEAN)),
    }


def make_cfg(abi, wl, dtype, n_envs, env_offset=0, n_global=0, **over):
    kw = dict(wl["kw"])
    kw.update(over)
    return abi.default_config(n_envs=n_envs, env_offset=env_offset, n_envs_global=n_global or n_envs, dtype=dtype, **kw)


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


def attach_exchange(eng, dist, rank, world, want_peers):
    """SHARED weights across GPUs.  Persistent-kernel engines: dW is summed inside the kernel through NVLink peer mailboxes
    (cudaIpc handles gathered with torch.distributed).  Per-step-kernel engines (cfg4): ncclAllReduce between the kernels."""
    from rsrl_b200.abi import RsrlError
    from rsrl_b200.engine import comm_unique_id
    import torch
    if want_peers:
        ok, handles = 1, [None] * world
        try:
            dist.all_gather_object(handles, eng.peer_export())
            eng.peer_attach(handles, rank, world)
        except RsrlError as err:
            ok = 0
            print(f"[bench] rank {rank}: peer attach failed ({err})", file=sys.stderr)
        flag = torch.tensor([ok], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 1:
            if eng.launch_shape()["fx"]:
                return ("in-kernel exchange (fp32): 2^-40 fixed-point counting accumulators in L2, CTA-group tables whose leaders add the group "
                        "totals to every GPU's world table with system-scope reductions over NVLink peer pointers (order independent)")
            return "in-kernel exchange (f64): cluster DSMEM + NVLink peer-memory LL words + L2 LL lines (fixed order)"
        raise SystemExit("peer attach failed on some rank")
    uid = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    eng.comm_init(uid[0], rank, world)
    return "ncclAllReduce(dW) per step (per-step kernels)"


def cpu_baseline_leg(abi, wl, dtype_f64):
    """The reference's CPU implementation of the path (oracle port built -O3 -march=native on this machine; rustc is
    unavailable), BASELINE.md section 3: all host threads of independent single-env agents, one core, the fused
    scalar variant, and BASELINE configs[0] (examples/q_learning.rs itself, one env, one core)."""
    from oracle import pyoracle as O
    O.build()
    threads = os.cpu_count() or 1
    ccfg = make_cfg(abi, wl, dtype_f64, 64, weight_mode=abi.PER_ENV)
    O.baseline_run(ccfg, threads, 64, 100)
    secs, n = O.baseline_run(ccfg, threads, 64, 40000)        # ~8 s of CPU work on all host cores
    s1, n1 = O.baseline_run(ccfg, 1, 64, 12000)               # ~3 s on one core
    sf, nf = O.baseline_run(ccfg, threads, 64, 40000, fused=True)
    out = {"value": n / secs, "unit": "env-steps/s", "cores": threads, "kind": "port", "per_core": n / secs / threads,
           "one_core": n1 / s1, "build": O.baseline_build_flags() + ", engine create + thread start outside the timed region",
           "sample": f"{threads} threads x 64 independent single-env reference-shaped agents (4 projections + heap allocations per step) "
                     f"x 40000 steps ({secs:.1f} s); one_core = 1 thread x 64 agents x 12000 steps",
           "fused_scalar": {"value": nf / sf, "unit": "env-steps/s", "cores": threads,
                            "what": "same agents, 1 projection per step, no allocation (best simple scalar CPU implementation)"}}
    if wl["F"] == 36 and wl["kw"].get("algo", abi.QLEARNING) == abi.QLEARNING:
        c1 = abi.default_config(n_envs=1, dtype=dtype_f64, max_episode_steps=10000)   # examples/q_learning.rs + the 10 000-step cap of BASELINE.md
        sc, nc, lens = O.baseline_run_single(c1, 2000000, 40)
        out["cfg1_single_env"] = {"value": nc / sc, "unit": "env-steps/s", "cores": 1, "steps": nc,
                                  "what": "BASELINE configs[0]: examples/q_learning.rs (1 env, start (-0.5, 0), Greedy, seed 0), 2 000 000 steps, cap 10 000",
                                  "first_episode_lengths": lens}
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the host cores (oracle port, see cpu_baseline_leg).
    Rank 0 only; each bench step is a bounded sample of the workload."""
    if rank != 0:
        return
    from rsrl_b200 import abi
    from oracle import pyoracle as O
    O.build()
    wl = workloads(abi)[args.config]
    threads = os.cpu_count() or 1
    envs_per_thread, k = 64, 2000   # sample: threads*64 envs x 2000 steps per bench step
    cfg = make_cfg(abi, wl, abi.F64, envs_per_thread, weight_mode=abi.PER_ENV)
    for _ in range(args.warmup):
        O.baseline_run(cfg, threads, envs_per_thread, 50)
    secs, steps = 0.0, 0
    for _ in range(args.steps):
        s, n = O.baseline_run(cfg, threads, envs_per_thread, k)
        secs += s
        steps += n
    v = steps / secs
    sample = (f"{threads} threads x {envs_per_thread} independent single-env reference-shaped agents x {k} steps per bench step; "
              f"built {O.baseline_build_flags()}; engine create + thread start outside the timed region")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": wl["label"], "sample": sample},
        "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": threads, "kind": "port", "per_core": v / threads, "sample": sample},
        "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--config", default="cfg2", choices=["cfg2", "cfg4", "cfg5"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side-runs", action="store_true")
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
    from rsrl_b200.engine import Engine

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

    wl = workloads(abi)[args.config]
    N_PER_GPU, K_INNER = wl["n"], wl["k_inner"]
    dtype = abi.F32 if args.dtype == "f32" else abi.F64
    n_global = N_PER_GPU * world
    cfg = make_cfg(abi, wl, dtype, N_PER_GPU, env_offset=rank * N_PER_GPU, n_global=n_global)
    cfg.device = local_rank
    eng = Engine(cfg)
    shape = eng.launch_shape()
    exchange = "none"
    if world > 1:
        exchange = attach_exchange(eng, dist, rank, world, want_peers=bool(shape["persistent"]))
    stream = torch.cuda.ExternalStream(eng.stream(), device=torch.device("cuda", local_rank))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def timed_steps(n_steps, body):
        """device time of n_steps bench steps, L2 flushed between steps (outside the event pairs)"""
        total_ms = 0.0
        for _ in range(n_steps):
            flush.zero_()
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
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
    out_s, out_a, out_w = np.empty_like(host_states), np.empty(cfg.n_envs, dtype=np.int32), np.empty((wl["F"], wl["A"]))
    h2d = host_states.nbytes
    d2h = out_s.nbytes + out_a.nbytes + out_w.nbytes

    def e2e_body(k):
        eng.set_states(host_states)   # H2D: this step's inputs
        eng.step(k)
        eng.states(out_s)             # D2H: results
        eng.actions(out_a)
        eng.weights(out_w)

    e2e_body(K_INNER)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_body(K_INNER)
    barrier()
    e2e_s = time.perf_counter() - t0
    # the same with ONE batched step per call: the launch + copy floor of the drop-in loop (transition -> handle -> sample per call)
    n1 = 50
    e2e_body(1)
    barrier()
    t0 = time.perf_counter()
    for _ in range(n1):
        e2e_body(1)
    barrier()
    e2e1_s = time.perf_counter() - t0

    t = torch.tensor([ms, e2e_s, e2e1_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_s, e2e1_s = float(t[0]), float(t[1]), float(t[2])
    env_steps = float(n_global) * K_INNER * args.steps
    value = env_steps / (ms * 1e-3)

    # ---- N > 1: the W replicas must be bit-identical and agree with one GPU running all the envs ----
    checks = {}
    if world > 1:
        # Horizon: 2 batched steps for the persistent engines (step 1 checks the cross-GPU sum, step 2 that every replica applied it).
        # Not longer: 1 and N GPUs round the per-CTA fp32 partials differently (1e-7), and this workload (greedy, zero initial
        # weights, reward -1) sits on a knife edge — from step 3 on a 1e-7 difference in W flips the argmax of a large fraction of the
        # envs at once; the CPU replay of the device arithmetic shows the same 9e-5 jump at step 3 and 3.7e-4 after 40 steps for
        # 8 x 65 536 envs (profiles/r02_persistent.md).  The strict check is tests/tools/multi_gpu_check.py: bit-exact vs oracle32.
        H = 2 if shape["persistent"] else 6
        eng.reset()
        barrier()
        eng.step(H)
        eng.sync()
        Wm = eng.weights()
        digest = np.frombuffer(hashlib.sha256(Wm.tobytes()).digest()[:8], dtype=np.int64).copy()
        dg = torch.from_numpy(digest).cuda()
        allg = [torch.empty_like(dg) for _ in range(world)]
        dist.all_gather(allg, dg)
        checks["replicas_identical"] = bool(all(int(g.item()) == int(allg[0].item()) for g in allg))
        match = 1
        if rank == 0:
            single = make_cfg(abi, wl, dtype, n_global)
            single.device = local_rank
            with Engine(single) as se:
                se.step(H)
                se.sync()
                Ws = se.weights()
            # same envs, same RNG streams; the fp32 sums associate differently on 1 and N GPUs: tolerance, not identity
            err = float(np.abs(Wm - Ws).max() / max(np.abs(Ws).max(), 1e-30))
            tol = 1e-9 if dtype == abi.F64 else (2e-6 if shape["persistent"] else 2e-4)
            checks["single_gpu_rel_err"] = err
            checks["single_gpu_horizon_steps"] = H
            match = int(err < tol and np.isfinite(Wm).all())
        mt = torch.tensor([match], device="cuda")
        dist.all_reduce(mt, op=dist.ReduceOp.MIN)
        checks["matches_single_gpu"] = bool(int(mt.item()))

    peaks, peak_src = {}, "fallback"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak_src = "measured"
    except (OSError, ValueError):
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    ach_gbs = wl["bytes"] * N_PER_GPU * K_INNER * args.steps / (ms * 1e-3) / 1e9  # per GPU
    sm_hz = float(clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)) * 1e6

    out = {
        "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 features/Q/weights + f64 physics" if dtype == abi.F32 else "f64",
        "data": "synthetic",
        "config": {"workload": wl["label"], "n_envs_per_gpu": N_PER_GPU, "batched_steps_per_bench_step": K_INNER,
                   "us_per_batched_step": 1e3 * ms / args.steps / K_INNER,
                   "l2": "flushed (256 MiB write) between timed steps", "exchange": exchange,
                   "launch_shape": {k: shape[k] for k in ("persistent", "grid", "cluster_size", "n_clusters", "block", "smem")}},
        "clocks": clocks,
        "e2e": {"value": env_steps / e2e_s, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "e2e_single_step": {"value": float(n_global) * n1 / e2e1_s, "unit": "env-steps/s", "us_per_call": 1e6 * e2e1_s / n1,
                            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                            "what": "one batched step per ABI call (set_states, step(1), get states/actions/weights): launch + PCIe floor"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": ach_gbs, "peak": peak_gbs, "unit": "GB/s", "frac": ach_gbs / peak_gbs,
                     "traffic": DRAM_BYTES_PER_LAUNCH if args.config == "cfg2" else None, "peak_source": peak_src,
                     "note": f"algorithmic {wl['bytes']} B/env-step x {N_PER_GPU} envs x {K_INNER} steps per launch (SURVEY 8d); traffic = "
                             "dram__bytes_read+write of one such launch (ncu, profiles/r02_persistent.md).  Env state is register-resident "
                             "for the whole launch, so HBM is idle; the binding roof is instruction issue + the per-step grid exchange: "
                             "see roofline_issue"},
    }
    if args.config == "cfg2" and WARP_INST_PER_ENV_STEP:
        issue_peak = 148 * 4 * sm_hz
        ach = WARP_INST_PER_ENV_STEP * N_PER_GPU * K_INNER * args.steps / (ms * 1e-3)
        out["roofline_issue"] = {"bound": "issue", "achieved": ach, "peak": issue_peak, "unit": "warp-inst/s", "frac": ach / issue_peak,
                                 "warp_inst_per_env_step": WARP_INST_PER_ENV_STEP,
                                 "note": "peak = 148 SMs x 4 schedulers x SM clock (median under load); achieved = measured instruction count "
                                         "(ncu smsp__inst_executed.sum / env-steps, profiles/r02_persistent.md) x env-steps/s"}
    out.update(checks)

    # ---- the same workload in the other weight mode / dtype, and the other BASELINE configs (context, not the headline) ----
    if world == 1 and args.config == "cfg2" and not args.no_side_runs:
        def side_run(label, wlk="cfg2", k=None, **kw):
            w2 = workloads(abi)[wlk]
            k = k or w2["k_inner"]
            c2 = make_cfg(abi, w2, kw.pop("dtype", dtype), kw.pop("n_envs", w2["n"]), **kw)
            with Engine(c2) as e2:
                st2 = torch.cuda.ExternalStream(e2.stream(), device=torch.device("cuda", local_rank))
                for _ in range(3):
                    e2.step(max(2, k // 4))
                e2.sync()
                tot, reps = 0.0, 4
                for _ in range(reps):
                    flush.zero_()
                    torch.cuda.synchronize()
                    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a0.record(st2)
                    e2.step(k)
                    a1.record(st2)
                    a1.synchronize()
                    tot += a0.elapsed_time(a1)
                v = c2.n_envs * k * reps / (tot * 1e-3)
                return {"value": v, "unit": "env-steps/s", "us_per_batched_step": 1e3 * tot / reps / k, "what": label,
                        "roofline": {"bound": "hbm", "achieved": w2["bytes"] * v / 1e9, "peak": peak_gbs, "unit": "GB/s",
                                     "frac": w2["bytes"] * v / 1e9 / peak_gbs}}
        pe = side_run("PER_ENV weights: 65536 independent reference agents, own W (36x3 fp32) resident in shared memory",
                      weight_mode=abi.PER_ENV)
        pe["roofline"].update(achieved=624 * pe["value"] / 1e9, frac=624 * pe["value"] / 1e9 / peak_gbs,
                              note="algorithmic 624 B/env-step (SURVEY 8d: state + read W + write one column); > 1 because W never leaves the SM")
        out["also"] = {"per_env_weights": pe,
                       "f64": side_run("SHARED weights, all arithmetic f64 (the reference's operations; parity anchor dtype)", dtype=abi.F64)}
        c3 = abi.default_config(n_envs=262144, dtype=dtype, domain=abi.CART_POLE, basis=abi.TILE_CODING, algo=abi.SARSA, policy=abi.EPSILON_GREEDY,
                                epsilon=0.1, gamma=0.99, lr=0.1 / 8, init_mode=abi.INIT_UNIFORM, init_lo=[-0.05] * 4, init_hi=[0.05] * 4,
                                max_episode_steps=500, seed=0, update_scale=abi.SCALE_MEAN)
        with Engine(c3) as e3:
            st3 = torch.cuda.ExternalStream(e3.stream(), device=torch.device("cuda", local_rank))
            e3.step(30)
            e3.sync()
            flush.zero_()
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(st3)
            e3.step(300)
            a1.record(st3)
            a1.synchronize()
            ms3 = a0.elapsed_time(a1)
        v3 = 262144 * 300 / (ms3 * 1e-3)
        out["also"]["cfg3_tile_coding"] = {
            "value": v3, "unit": "env-steps/s", "us_per_batched_step": 1e3 * ms3 / 300,
            "what": "cfg3: CartPole SARSA TileCoding(8 tilings, 8 tiles/dim, 4096 rows) eps-greedy, 262144 envs, SHARED",
            "roofline": {"bound": "hbm", "achieved": 80 * v3 / 1e9, "peak": peak_gbs, "unit": "GB/s", "frac": 80 * v3 / 1e9 / peak_gbs,
                         "note": "algorithmic 80 B/env-step (SURVEY 8d); binding roof: instruction issue of the f64 RK4 + tile hashing"}}
        c4 = side_run(workloads(abi)["cfg4"]["label"], wlk="cfg4", k=20)
        tf32_peak = float(peaks.get("bf16_tflops", 1590.0)) / 2.0
        # per env-step: algorithmic 3 contractions x 2*4096*3; executed 3 TF32 passes x (2 evaluations x {re,im} x 192x64 MACs + {re,im} x 128x96 MACs) x 2
        alg_flop, exe_flop = 3 * 2 * 4096 * 3, 3 * (2 * 2 * 192 * 64 * 2 + 2 * 128 * 96 * 2)
        c4["roofline"] = {"bound": "tensor", "achieved": alg_flop * c4["value"] / 1e12, "peak": tf32_peak, "unit": "TFLOP/s",
                          "frac": alg_flop * c4["value"] / 1e12 / tf32_peak, "executed_tflops": exe_flop * c4["value"] / 1e12,
                          "note": "algorithmic 73.7 kFLOP/env-step (two Q = Phi W evaluations + dW = Phi^T D, SURVEY 8d); peak = measured bf16 "
                                  "cuBLAS TFLOP/s / 2 (kind::tf32 runs at half the bf16 rate); executed = 3xTF32 passes over the complex "
                                  "(real, imaginary) split, 6x the algorithmic count"}
        out["also"]["cfg4_fourier7_tensor_core"] = c4
        out["also"]["cfg5_sarsa_lambda"] = side_run(workloads(abi)["cfg5"]["label"], wlk="cfg5", k=1000)

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_leg(abi, wl, abi.F64)
    if rank == 0:
        print(json.dumps(out))
    eng.close()
    if dist is not None:
        dist.destroy_process_group()
    if checks and not (checks.get("replicas_identical") and checks.get("matches_single_gpu")):
        sys.exit(3)


if __name__ == "__main__":
    main()
