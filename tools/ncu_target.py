"""One short persistent launch for ncu: python tools/ncu_target.py [n_envs] [k_steps] [dtype]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsrl_b200 import abi
from rsrl_b200.engine import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
k = int(sys.argv[2]) if len(sys.argv) > 2 else 50
dt = abi.F64 if len(sys.argv) > 3 and sys.argv[3] == "f64" else abi.F32
extra = {}
if os.environ.get('RSRL_ALGO') == 'sarsa_lambda':
    extra = dict(policy=abi.EPSILON_GREEDY, epsilon=0.2, algo=abi.SARSA_LAMBDA, alpha=0.01, gamma=0.99)
cfg = abi.default_config(n_envs=n, **extra, dtype=dt, init_mode=abi.INIT_UNIFORM, init_lo=[-0.6, 0.0], init_hi=[-0.4, 0.0],
                         max_episode_steps=1000, seed=0, update_scale=abi.SCALE_MEAN)
with Engine(cfg) as e:
    e.step(k); e.sync()      # warm-up launch
    e.step(k); e.sync()      # profiled launch (ncu -s 1 skips init + warm-up kernels as needed)
    print(e.stats())
