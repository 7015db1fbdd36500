"""Where the end-to-end step (host observations -> loss.backward()) spends its wall time: cProfile of bench.py's e2e step."""
import cProfile
import dataclasses
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from molgym_b200 import ppo, synth  # noqa: E402
from molgym_b200.agents.covariant.agent import CovariantAC  # noqa: E402
from molgym_b200.spaces import ActionSpace, ObservationSpace  # noqa: E402

cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else 'C2']
dev = torch.device('cuda:0')
torch.manual_seed(0)
agent = CovariantAC(ObservationSpace(cfg.canvas_size, cfg.zs), ActionSpace(cfg.zs), device=dev, **cfg.agent_kwargs())
obs, n = synth.make_observations(cfg)
act = synth.make_actions(cfg, obs, n)
with torch.no_grad():
    logp0 = agent.step(obs, act)['logp'].cpu().numpy()
old_logp, adv, ret = synth.make_ppo_targets(cfg, logp0)
data = dict(obs=obs, act=act, logp=old_logp, adv=adv, ret=ret)


def step():
    agent.zero_grad()
    loss, info = ppo.compute_loss(agent, data, 0.2, 0.5, 0.01)
    loss.backward()


for _ in range(20):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    step()
torch.cuda.synchronize()
print('wall ms/step', (time.perf_counter() - t0) / 200 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats('cumulative').print_stats(35)
