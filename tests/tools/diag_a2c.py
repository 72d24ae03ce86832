"""Diagnostic (GPU box): first step / env where the PER_ENV A2C engine leaves the f64 oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from rsrl_b200 import abi
from rsrl_b200.engine import Engine
from oracle import pyoracle as O
kw = dict(algo=abi.A2C, policy=abi.SOFTMAX, epsilon=0.5, basis_order=5, lr=0.05, alpha=0.05, gamma=0.99)
cfg = abi.default_config(n_envs=300, dtype=abi.F64, init_mode=abi.INIT_UNIFORM, init_lo=[-0.6, 0.0], init_hi=[-0.4, 0.0],
                         max_episode_steps=120, seed=11, record_td_error=1, weight_mode=abi.PER_ENV, **kw)
with Engine(cfg) as e:
    o = O.Engine(cfg)
    for t in range(260):
        se, so = e.states().copy(), o.states().copy()
        e.step(1); o.step(1); e.sync()
        da = np.nonzero(e.actions() != o.actions())[0]
        ds = np.abs(e.states() - o.states()).max(axis=1)
        dw = np.abs(e.weights() - o.weights()).reshape(300, -1).max(axis=1)
        dp = np.abs(e.aux_weights() - o.aux_weights()).reshape(300, -1).max(axis=1)
        dtd = np.abs(e.td_errors() - o.td_errors())
        if len(da) or ds.max() > 1e-12 or dw.max() > 1e-9 or dp.max() > 1e-9:
            i = int(da[0]) if len(da) else int(np.argmax(np.maximum(ds, np.maximum(dw, dp))))
            print(f"step {t}: env {i}: action {e.actions()[i]} vs {o.actions()[i]}  ds {ds[i]:.3e} dW {dw[i]:.3e} dtheta {dp[i]:.3e} dtd {dtd[i]:.3e}")
            print("  state before", se[i], so[i], " after", e.states()[i], o.states()[i])
            print("  |theta| max", np.abs(o.aux_weights()[i]).max(), " |W| max", np.abs(o.weights()[i]).max(), "td", e.td_errors()[i], o.td_errors()[i])
            print("  envs differing: actions", len(da), " states", int((ds > 1e-12).sum()), " W", int((dw > 1e-9).sum()), " theta", int((dp > 1e-9).sum()))
            break
    else:
        print("no divergence in 260 steps; max dW", dw.max(), "max dtheta", dp.max())
