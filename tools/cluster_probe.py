"""us per batched step of the cfg2 kernel: counting exchange (fx) vs cluster sizes of the LL exchange (run on the GPU box).
python tools/cluster_probe.py [fx|<cluster size>] ..."""
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
    print(f"nsub={os.environ.get('RSRL_B200_NSUB')} delay={os.environ.get('RSRL_B200_POLL_DELAY')} backoff={os.environ.get('RSRL_B200_POLL_BACKOFF')} fx={sh['fx']} cluster={sh['cluster_size']} skip={os.environ.get('RSRL_B200_DEBUG_SKIP')} grid={sh['grid']} ncl={sh['n_clusters']} block={sh['block']}: {1e6*dt/4000:.2f} us/step {cfg.n_envs*4000/dt/1e9:.2f} G/s", flush=True)
'''
for cs in (sys.argv[1:] or ["400:0:1", "300:0:2", "500:0:2", "700:0:2", "300:0:4", "500:0:4", "700:0:4", "500:100:4"]):   # poll delay : backoff (ns) : sub-tables
    base = dict(os.environ, RSRL_B200_POLL_DELAY=cs.split(":")[0], RSRL_B200_POLL_BACKOFF=cs.split(":")[1], RSRL_B200_NSUB=cs.split(":")[2])
    for skip in ("0",):
        r = subprocess.run([sys.executable, "-c", code], env=dict(base, RSRL_B200_DEBUG_SKIP=skip), capture_output=True, text=True, timeout=120)
        print(r.stdout.strip() or r.stderr[-300:], flush=True)
    if os.environ.get("PHASES"):
        r = subprocess.run([sys.executable, "-c", code], env=dict(base, RSRL_B200_PHASE_PROFILE="1"), capture_output=True, text=True, timeout=120)
        print(r.stdout.strip(), "\n", "\n".join(l for l in r.stderr.splitlines() if "[phase] -" not in l and "leader" not in l)[-900:], flush=True)
