"""python tools/time_cfg.py <cfg5|cfg3|cfg2|per_env> [k]: us per batched step (honours RSRL_B200_DEBUG_SKIP / PHASE_PROFILE)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsrl_b200 import abi
from rsrl_b200.engine import Engine
which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
base = dict(n_envs=65536, dtype=abi.F32, init_mode=abi.INIT_UNIFORM, init_lo=[-0.6, 0.0], init_hi=[-0.4, 0.0],
            max_episode_steps=1000, seed=0, update_scale=abi.SCALE_MEAN)
if which == "cfg5":
    base.update(n_envs=32768, policy=abi.EPSILON_GREEDY, epsilon=0.2, algo=abi.SARSA_LAMBDA, alpha=0.01, gamma=0.99)
elif which == "cfg5_greedy":
    base.update(n_envs=32768, policy=abi.GREEDY, algo=abi.SARSA_LAMBDA, alpha=0.01, gamma=0.99)
elif which == "sarsa_eps_32k":
    base.update(n_envs=32768, policy=abi.EPSILON_GREEDY, epsilon=0.2, algo=abi.SARSA, gamma=0.99)
elif which == "ql_32k":
    base.update(n_envs=32768)
elif which == "per_env":
    base.update(weight_mode=abi.PER_ENV, update_scale=abi.SCALE_SUM)
cfg = abi.default_config(**base)
with Engine(cfg) as e:
    e.step(200); e.sync()
    t0 = time.perf_counter(); e.step(k); e.sync(); dt = time.perf_counter() - t0
    print(f"{which:14s} skip={os.environ.get('RSRL_B200_DEBUG_SKIP','0')} {1e6*dt/k:8.2f} us/step {cfg.n_envs*k/dt/1e9:7.2f} G env-steps/s", flush=True)
