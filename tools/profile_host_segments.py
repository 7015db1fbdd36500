import sys, time, collections
sys.path.insert(0, '/root/repo')
import torch, numpy as np
from molgym_b200 import ppo, synth
from molgym_b200.agents.covariant import agent as A
from molgym_b200.spaces import ActionSpace, ObservationSpace
cfg = synth.CONFIGS['C2']
dev = torch.device('cuda:0')
torch.manual_seed(0)
ag = A.CovariantAC(ObservationSpace(cfg.canvas_size, cfg.zs), ActionSpace(cfg.zs), device=dev, **cfg.agent_kwargs())
obs, n = synth.make_observations(cfg); act = synth.make_actions(cfg, obs, n)
with torch.no_grad(): logp0 = ag.step(obs, act)['logp'].cpu().numpy()
old, adv, ret = synth.make_ppo_targets(cfg, logp0)
data = dict(obs=obs, act=act, logp=old, adv=adv, ret=ret)
seg = collections.defaultdict(float)
orig_replay = torch.cuda.CUDAGraph.replay
def timed_replay(self):
    t = time.perf_counter(); orig_replay(self); seg['replay'] += time.perf_counter() - t
torch.cuda.CUDAGraph.replay = timed_replay
orig_pack = A.pack_observations
def timed_pack(*a, **k):
    t = time.perf_counter(); r = orig_pack(*a, **k); seg['pack'] += time.perf_counter() - t; return r
A.pack_observations = timed_pack
orig_sync = torch.cuda.Event.synchronize
def timed_sync(self):
    t = time.perf_counter(); orig_sync(self); seg['event_sync'] += time.perf_counter() - t
torch.cuda.Event.synchronize = timed_sync
def step(i):
    if i % 4 == 0: ag.zero_grad()
    t = time.perf_counter(); loss, info = ppo.compute_loss(ag, data, 0.2, 0.5, 0.01); seg['compute_loss'] += time.perf_counter() - t
    t = time.perf_counter(); loss.backward(); seg['backward'] += time.perf_counter() - t
for i in range(20): step(i)
torch.cuda.synchronize(); seg.clear()
K = 200
t0 = time.perf_counter()
for i in range(K): step(i)
torch.cuda.synchronize()
tot = time.perf_counter() - t0
print('wall ms/step', tot / K * 1e3)
for k, v in seg.items(): print('%-14s %.1f us/step' % (k, v / K * 1e6))
