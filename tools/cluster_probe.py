"""us per batched step of the cfg2 kernel for exchange settings (run on the GPU box).
python tools/cluster_probe.py <poll delay ns>:<backoff ns>:<multi-GPU CTA groups> ...   (PHASES=1: phase table)"""
import os, subprocess, sys
code = r'''
import sys, os, time
sys.path.insert(0, os.getcwd())
from rsrl_b200 import abi
from rsrl_b200.engine import Engine
cfg = abi.default_config(n_envs=int(os.environ.get("N", 65536)), dtype=abi.F32, init_mode=abi.INIT_UNIFORM, init_lo=[-0.6, 0.0], init_hi=[-0.4, 0.0],
                         max_episode_steps=1000, seed=0, update_scale=abi.SCALE_MEAN)
with Engine(cfg) as e:
    sh = e.launch_shape()
    e.step(500); e.sync()
    t0 = time.perf_counter(); e.step(4000); e.sync(); dt = time.perf_counter() - t0
    print(f"ngroups={os.environ.get('RSRL_B200_NGROUPS')} delay={os.environ.get('RSRL_B200_POLL_DELAY')} backoff={os.environ.get('RSRL_B200_POLL_BACKOFF')} fx={sh['fx']} grid={sh['grid']} block={sh['block']}: {1e6*dt/4000:.2f} us/step {cfg.n_envs*4000/dt/1e9:.2f} G/s", flush=True)
'''
for cs in (sys.argv[1:] or ["400:0:1:16"]):
    f = (cs.split(":") + ["8"])[:3]
    base = dict(os.environ, RSRL_B200_POLL_DELAY=f[0], RSRL_B200_POLL_BACKOFF=f[1], RSRL_B200_NGROUPS=f[2])
    r = subprocess.run([sys.executable, "-c", code], env=base, capture_output=True, text=True, timeout=120)
    print(r.stdout.strip() or r.stderr[-300:], flush=True)
    if os.environ.get("PHASES"):
        r = subprocess.run([sys.executable, "-c", code], env=dict(base, RSRL_B200_PHASE_PROFILE="1"), capture_output=True, text=True, timeout=120)
        print("\n".join(l for l in r.stderr.splitlines() if "[phase] -" not in l and "leader" not in l)[-900:], flush=True)
