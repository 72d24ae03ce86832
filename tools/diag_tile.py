import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np
from rsrl_b200 import abi
from rsrl_b200.engine import Engine
from oracle import pyoracle as O
O.build()
base = dict(domain=abi.CART_POLE, basis=abi.TILE_CODING, n_tilings=8, tiles_per_dim=8, memory_size=4096, algo=abi.SARSA,
            policy=abi.EPSILON_GREEDY, epsilon=0.1, gamma=0.99, lr=0.1 / 8, dtype=abi.F64,
            init_mode=abi.INIT_UNIFORM, init_lo=[-0.05] * 4, init_hi=[0.05] * 4, max_episode_steps=500, seed=13,
            update_scale=abi.SCALE_MEAN, record_td_error=1)
for n in (2100, 1024, 130):
    for dense in ("1", "0"):
        os.environ["RSRL_B200_TILE_DENSE"] = dense
        cfg = abi.default_config(n_envs=n, **base)
        with Engine(cfg) as e:
            o = O.Engine(cfg)
            first = None
            for t in range(41):
                e.step(1); o.step(1); e.sync()
                bad = (e.actions() != o.actions())
                werr = np.abs(e.weights() - o.weights()).max()
                serr = np.abs(e.states() - o.states()).max()
                if bad.any() or werr > 1e-9:
                    first = (t, int(bad.sum()), np.nonzero(bad)[0][:5].tolist(), werr, serr, np.abs(o.weights()).max())
                    break
            print(f"n={n} dense={dense}: first divergence {first}", flush=True)
