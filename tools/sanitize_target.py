"""compute-sanitizer target: short runs of the main kernels (python tools/sanitize_target.py [cfg2|cfg5|tile|dyn|f64])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsrl_b200 import abi
from rsrl_b200.engine import Engine
which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
mc = dict(init_mode=abi.INIT_UNIFORM, init_lo=[-0.6, 0.0], init_hi=[-0.4, 0.0], max_episode_steps=20, seed=1, update_scale=abi.SCALE_MEAN)
cfgs = {
    "cfg2": abi.default_config(n_envs=20000, dtype=abi.F32, **mc),
    "f64": abi.default_config(n_envs=20000, dtype=abi.F64, **mc),
    "cfg5": abi.default_config(n_envs=9000, dtype=abi.F32, algo=abi.SARSA_LAMBDA, policy=abi.EPSILON_GREEDY, epsilon=0.2, alpha=0.01, gamma=0.99, **mc),
    "tile": abi.default_config(n_envs=40000, dtype=abi.F32, domain=abi.CART_POLE, basis=abi.TILE_CODING, n_tilings=8, tiles_per_dim=8, memory_size=4096,
                               algo=abi.SARSA, policy=abi.EPSILON_GREEDY, epsilon=0.1, gamma=0.99, lr=0.1 / 8, init_mode=abi.INIT_UNIFORM,
                               init_lo=[-0.05] * 4, init_hi=[0.05] * 4, max_episode_steps=30, seed=1, update_scale=abi.SCALE_MEAN),
    "dyn": abi.default_config(n_envs=3000, dtype=abi.F32, basis_order=4, **mc),
}
with Engine(cfgs[which]) as e:
    e.step(3); e.step(25); e.sync()
    print(which, e.launch_shape()["persistent"], e.stats()["total_steps"], flush=True)
