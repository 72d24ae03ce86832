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
elif which == "cfg4":
    base.update(n_envs=131072, domain=abi.ACROBOT, basis_order=7, algo=abi.EXPECTED_SARSA, policy=abi.EPSILON_GREEDY, epsilon=0.1,
                gamma=0.99, lr=1e-4, alpha=1.0, init_lo=[-0.1] * 4, init_hi=[0.1] * 4, max_episode_steps=500)
elif which == "cfg3":
    base.update(n_envs=262144, domain=abi.CART_POLE, basis=abi.TILE_CODING, algo=abi.SARSA, policy=abi.EPSILON_GREEDY, epsilon=0.1,
                gamma=0.99, lr=0.1 / 8, init_lo=[-0.05] * 4, init_hi=[0.05] * 4, max_episode_steps=500)
elif which == "per_env":
    base.update(weight_mode=abi.PER_ENV, update_scale=abi.SCALE_SUM)
cfg = abi.default_config(**base)
with Engine(cfg) as e:
    e.step(int(os.environ.get('WARM', '200'))); e.sync()
    t0 = time.perf_counter(); e.step(k); e.sync(); dt = time.perf_counter() - t0
    print(f"{which:14s} skip={os.environ.get('RSRL_B200_DEBUG_SKIP','0')} {1e6*dt/k:8.2f} us/step {cfg.n_envs*k/dt/1e9:7.2f} G env-steps/s", flush=True)
