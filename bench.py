"""bench.py — PPO-minibatch forward+backward throughput of the covariant agent (canvases / s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2] [--impl ours|reference] [--no-per-config]
    (MOLGYM_B200_NO_GRAPH=1: eager launches only)

A "step" is one pass of the hot path over one PPO minibatch of synthetic canvases: forward (Cormorant body + heads) ->
PPO-clip loss -> backward -> (N > 1) the exchange (all-reduce of the 8-double loss info and of the flat gradient).

Multi-GPU: the minibatch is SHARDED — every rank builds the same global minibatch and the data-parallel agent evaluates its
contiguous shard (molgym_b200/parallel.py).  The headline workload (C2, which BASELINE.json quotes on one GPU) is scaled weakly:
global minibatch = 140 x N, 140 canvases per rank.  The configurations BASELINE.json names for 8 GPUs (C3 1024, C4 4096,
C5 8192 canvases) are reported in `per_config` with their FIXED global minibatch sharded over the N ranks (strong scaling;
at N = 1 C5 runs its one-GPU shard of 1024 canvases).

  value            : inputs resident in HBM, direct C-ABI calls (mgb_cov_forward, mgb_ppo_loss, mgb_cov_backward) captured as one CUDA
                     graph per slot; the better of (a) sequential replays, L2 flushed between steps, an event pair per step and
                     (b) K steps replayed round-robin over independent slots on two streams, one event pair around all of them, the
                     rotation's workspaces larger than L2 (both reported: sequential_ms_per_step / pipelined_ms_per_step); max over ranks.
  e2e              : ppo.train (ppo.py:99-160) through the public API on HOST observation tuples, TRAIN_ITERS optimizer steps per call
                     (the reference default): per optimizer step zero_grad, EPOCH_LEN x (compute_loss -> fused CUDA-graph step,
                     loss.backward()), KL check, gradient norm, clipping and optimizer.step(); wall clock around the whole loop, L2
                     flush (once per call) counted inside.
  e2e_unchanged_ppo: the same loop with the arithmetic of the reference's own compute_loss (agent.step + torch ops + autograd).
  roofline / cpu_baseline / per_config : see DESIGN.md.
"""
import argparse
import ctypes
import dataclasses
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CLIP, VF, ENT = 0.2, 0.5, 0.01   # arg_parser.py:84-86
LR, GRAD_CLIP = 3e-4, 0.5         # arg_parser.py:80,88
METRIC = 'ppo_minibatch_fwd_bwd_canvases_per_sec'
EPOCH_LEN = 4                     # minibatches per optimizer step in the e2e loop
TRAIN_ITERS = 7                   # optimizer steps per ppo.train call: the reference's default --max_num_train_iters (tools/arg_parser.py:87)
FFMA_PEAK_TFLOPS = 72.3           # measured on this pool's B200 (tools/ffma_peak.cu, profiles/r1_ffma_peak.txt); nominal 74.4


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--workload', default='C2')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=None, help='override the minibatch size (per rank)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-per-config', action='store_true')
    ap.add_argument('--profile-kernel', default='k_atom_bwd')
    return ap.parse_args()


def workload(name, batch=None):
    from molgym_b200 import synth
    cfg = synth.CONFIGS[name]
    if batch:
        cfg = dataclasses.replace(cfg, mini_batch_size=batch)
    return cfg


def config_json(cfg, world, global_batch, scaling):
    return {'workload': f'{cfg.name}: canvas_size={cfg.canvas_size} zs={cfg.zs}, global minibatch {global_batch} sharded over {world} rank(s) '
                        f'({global_batch // world} per rank, {scaling} scaling), occupancies 0..K-1 uniform (synthetic PPO buffer)',
            'global_batch': global_batch, 'canvas_size': cfg.canvas_size,
            'hyper': {'network_width': cfg.network_width, 'maxl': cfg.maxl, 'num_cg_levels': cfg.num_cg_levels,
                      'num_channels_hidden': cfg.num_channels_hidden, 'num_channels_per_element': cfg.num_channels_per_element,
                      'num_gaussians': cfg.num_gaussians, 'beta': cfg.beta},
            'parallelism': f'dp{world} (minibatch sharded inside the agent)',
            'l2': 'sequential steps: flushed between timed steps (256 MiB write); pipelined steps: every slot of the rotation has its own workspace '
                  'and the rotation exceeds 2 x L2 (or the workspace itself exceeds L2); e2e: flush inside the timed region'}


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (CPU restatement of the reference) driven by the restated ppo loss, all host threads
# ----------------------------------------------------------------------------------------------------------------
def cpu_step_fn(cfg, sample):
    import torch
    from molgym_b200 import synth
    from oracle.molgym_oracle import CovariantOracle, ppo_loss
    torch.manual_seed(0)
    oracle = CovariantOracle(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    obs, n = synth.make_observations(cfg, batch=sample)
    act = synth.make_actions(cfg, obs, n)
    with torch.no_grad():
        logp0 = oracle.step(obs, act)['logp'].numpy()
    old_logp, adv, ret = synth.make_ppo_targets(cfg, logp0)
    chunk = 128   # the reference-style formulation materialises B*N^2*(2.5k floats): chunk + accumulate (ppo.py:122-131)

    def step():
        oracle.zero_grad()
        for lo in range(0, sample, chunk):
            hi = min(sample, lo + chunk)
            out = oracle.step(obs[lo:hi], act[lo:hi])
            loss, _ = ppo_loss(out['logp'], out['ent'], out['v'], old_logp[lo:hi], adv[lo:hi], ret[lo:hi], CLIP, VF, ENT)
            (loss * ((hi - lo) / sample)).backward()
    return step


def run_cpu(cfg, steps, warmup, budget_s=25.0):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample = min(cfg.mini_batch_size, 140)
    step = cpu_step_fn(cfg, sample)
    t0 = time.perf_counter()
    step()
    first = time.perf_counter() - t0
    warmup = max(0, min(warmup, int(budget_s / 4 / max(first, 1e-3))))
    for _ in range(warmup):
        step()
    steps = max(1, min(steps, int(budget_s / max(first, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return dict(value=sample / dt, ms_per_step=dt * 1e3, steps=steps, warmup=warmup, cores=cores,
                sample=f'{sample} canvases of {cfg.name} per step, {steps} steps, oracle fwd+loss+bwd in float32, '
                       f'torch.set_num_threads({cores})')


# ----------------------------------------------------------------------------------------------------------------
# clocks sampling (pynvml) during the timed region
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: 'gpu_idle', 0x2: 'applications_clocks_setting', 0x4: 'sw_power_cap', 0x8: 'hw_slowdown', 0x10: 'sync_boost',
               0x20: 'sw_thermal_slowdown', 0x40: 'hw_thermal_slowdown', 0x80: 'hw_power_brake_slowdown', 0x100: 'display_clock_setting'}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != 'gpu_idle':
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread:
            self._stop.set()
            self._thread.join()
        return {'sm_mhz': float(np.median(self.samples)) if self.samples else None, 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons), 'samples': len(self.samples)}


# ----------------------------------------------------------------------------------------------------------------
# algorithmic work model of the dominant kernel (DESIGN.md "Work model")
# ----------------------------------------------------------------------------------------------------------------
def cg_term_counts(lib):
    """Non-zero Clebsch-Gordan coefficients of the products the atom level evaluates, for l <= 4: (edge x A_k with the ells of
    A_k limited to nlin) and (A_k x A_k)."""
    def count(n1, n2):
        tot = 0
        for l1 in range(n1):
            for l2 in range(n2):
                for l in range(abs(l1 - l2), min(l1 + l2, 4) + 1):
                    for m1 in range(-l1, l1 + 1):
                        for m2 in range(-l2, l2 + 1):
                            if abs(m1 + m2) <= l and lib.mgb_clebsch_gordan(l1, m1, l2, m2, l, m1 + m2) != 0.0:
                                tot += 1
        return tot
    return {1: (count(5, 1), count(1, 1)), 5: (count(5, 5), count(5, 5))}


def atom_bwd_work(cfg, n_atoms, cat_sizes, terms):
    """FLOPs and algorithmic HBM bytes of ALL k_atom_bwd launches of one step (one per CG level; two half kernels per level for
    small minibatches).  Per valid atom i of a canvas with n atoms, level k with nlm2 input components (1 at level 0, else 25),
    C channels, Cout output channels:
      flops  = 2 * n * 25 * nlm2 * C * 8              row + column Kronecker passes (complex MAC = 8 flop)
             + (2 * T_ag + 2 * T_sq) * C * 4          Clebsch-Gordan scatter with the REAL term counts (row, column: T_ag non-zero
                                                      coefficients each; square: T_sq for each of its two factors), real x complex
             + nlm2 * nlm2 * C * 8                    own-atom square products
      bytes  = 25 * Cout * 8 (dA_{k+1}[i], read once) + nlm2 * C * 8 (A_k[i]) + nlm2 * C * 8 (dA_k[i] written once)
             + n * (5 * C * 8 (E_ij read) + 5 * C * 8 (dE_ij written))
               -- SURVEY.md 8d: each layer-boundary tensor once; the cat / dcat vectors, A_j re-reads and the dA_j read-modify-write
               are on-chip / cache traffic of this implementation, not algorithmic bytes."""
    C, nl = cfg.num_channels_hidden, cfg.maxl + 1
    cout_last = len(cfg.zs) * cfg.num_channels_per_element
    bytes_, flops = 0, 0
    for k in range(cfg.num_cg_levels):
        nlm2 = 1 if k == 0 else 25
        t_ag, t_sq = terms[1 if k == 0 else 5]
        cout = cout_last if k == cfg.num_cg_levels - 1 else C
        for n in n_atoms:
            n = int(n)
            flops += n * (2 * n * 25 * nlm2 * C * 8 + (2 * t_ag + 2 * t_sq) * C * 4 + nlm2 * nlm2 * C * 8)
            bytes_ += n * (25 * cout * 8 + 2 * nlm2 * C * 8 + n * (2 * 5 * C * 8))
    return bytes_, flops


# ----------------------------------------------------------------------------------------------------------------
class Case:
    """One workload on this rank: the agent, the global minibatch (identical on every rank) and this rank's device-resident shard."""

    def __init__(self, cfg, global_batch, world, rank, dev, seed_shift=0):
        import torch
        from molgym_b200 import _cabi, _lib, parallel, synth
        from molgym_b200.agents.covariant.agent import CovariantAC
        from molgym_b200.spaces import ActionSpace, ObservationSpace
        self.cfg, self.world, self.rank, self.dev, self.n_global = cfg, world, rank, dev, global_batch
        self.lib = lib = _lib.load()
        torch.manual_seed(0)
        self.agent = agent = CovariantAC(ObservationSpace(cfg.canvas_size, cfg.zs), ActionSpace(cfg.zs), device=dev, **cfg.agent_kwargs())
        obs, n_atoms = synth.make_observations(cfg, batch=global_batch, seed=cfg.seed + seed_shift)
        act = synth.make_actions(cfg, obs, n_atoms, seed=cfg.seed + seed_shift)
        lo, hi = parallel.shard_bounds(global_batch, rank, world)
        self.lo, self.hi, self.B = lo, hi, hi - lo
        # old log-probabilities: this agent's own log-probabilities of the stored actions + noise.  Each rank only ever reads its
        # shard of the buffer, so it evaluates just that part.
        with torch.no_grad():
            logp0 = np.zeros(global_batch, dtype=np.float32)
            for c0 in range(lo, hi, 1024):
                c1 = min(hi, c0 + 1024)
                logp0[c0:c1] = agent.step(obs[c0:c1], act[c0:c1])['logp'].cpu().numpy()
            agent._ws_cache.clear()
        if world > 1:
            parallel.shard_agent(agent)
        old_logp, adv, ret = synth.make_ppo_targets(cfg, logp0, seed=cfg.seed + seed_shift)
        self.data = dict(obs=obs, act=act, logp=old_logp, adv=adv, ret=ret)
        self.n_atoms_local = n_atoms[lo:hi]
        B = self.B
        parsed = agent.parse_observations(obs[lo:hi])
        self.pos, self.charges, self.bags = parsed['positions'], parsed['charges'], parsed['bags']
        self.act_d = torch.as_tensor(act[lo:hi], dtype=torch.float32, device=dev)
        self.old_d = torch.as_tensor(old_logp[lo:hi], device=dev)
        self.adv_d = torch.as_tensor(adv[lo:hi], device=dev)
        self.ret_d = torch.as_tensor(ret[lo:hi], device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        self.out = torch.empty(6, B, **f32)
        self.info = torch.zeros(8, dtype=torch.float64, device=dev)
        self.grad = torch.zeros_like(agent._flat)
        self.ws_bytes = lib.mgb_cov_workspace_bytes(agent._plan, B)
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        self.outs = _cabi.CovOutputs()
        self.outs.logp, self.outs.ent, self.outs.v = self.out[0].data_ptr(), self.out[1].data_ptr(), self.out[2].data_ptr()
        self.h2d_bytes = (self.pos.numel() + self.charges.numel() + self.bags.numel() + self.act_d.numel() + self.old_d.numel()) * 4 + \
            (self.adv_d.numel() + self.ret_d.numel()) * 8
        self.graph = None

    def launch(self, stream):
        from molgym_b200 import _cabi
        lib, a, B, o = self.lib, self.agent, self.B, self.out
        _cabi.check(lib, lib.mgb_cov_forward(a._plan, B, self.pos.data_ptr(), self.charges.data_ptr(), self.bags.data_ptr(),
                                             self.act_d.data_ptr(), a._flat.data_ptr(), self.ws.data_ptr(), self.ws.numel(),
                                             ctypes.byref(self.outs), stream))
        _cabi.check(lib, lib.mgb_ppo_loss(B, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), self.old_d.data_ptr(), self.adv_d.data_ptr(),
                                          self.ret_d.data_ptr(), CLIP, VF, ENT, 1.0 / self.n_global, self.info.data_ptr(), o[3].data_ptr(),
                                          o[4].data_ptr(), o[5].data_ptr(), stream))
        _cabi.check(lib, lib.mgb_cov_backward(a._plan, B, self.pos.data_ptr(), self.charges.data_ptr(), self.bags.data_ptr(),
                                              self.act_d.data_ptr(), a._flat.data_ptr(), self.ws.data_ptr(), self.ws.numel(),
                                              o[3].data_ptr(), o[4].data_ptr(), o[5].data_ptr(), self.grad.data_ptr(), 0, stream))

    def exchange(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(self.info, op=dist.ReduceOp.SUM)
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM)

    def eager_step(self):
        import torch
        self.launch(torch.cuda.current_stream(self.dev).cuda_stream)
        self.exchange()

    def capture(self):
        import torch
        if os.environ.get('MOLGYM_B200_NO_GRAPH'):
            return None
        try:
            g = torch.cuda.CUDAGraph()
            cap = torch.cuda.Stream(self.dev, priority=-5)   # main-chain kernels outrank the weight-gradient side streams
            cap.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(cap):
                with torch.cuda.graph(g, stream=cap):
                    self.launch(torch.cuda.current_stream(self.dev).cuda_stream)
            torch.cuda.current_stream(self.dev).wait_stream(cap)
            self.graph = g
        except Exception as exc:   # pragma: no cover
            sys.stderr.write(f'CUDA graph capture failed ({exc}); eager launches only\n')
            self.graph = None
        return self.graph

    def graph_step(self):
        self.graph.replay()
        self.exchange()

    # ---- pipelined throughput: independent slots (own workspace / outputs / gradient / graph) replayed round-robin on two streams
    def make_slots(self, n_slots):
        """Minibatches of a PPO epoch are independent (the parameters only move at optimizer.step(), ppo.py:122-146): the forward of
        one overlaps the backward of the previous one, as the fused e2e path does.  Every slot has its own workspace, so with enough
        slots the working set of the rotation exceeds L2 and a slot's buffers have left the cache when its turn comes again."""
        import torch
        from molgym_b200 import _cabi
        self.slots = []
        streams = [torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)]
        keep = (self.ws, self.out, self.info, self.grad, self.outs, self.graph)
        for k in range(n_slots):
            f32 = dict(dtype=torch.float32, device=self.dev)
            self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.dev)
            self.out = torch.empty(6, self.B, **f32)
            self.info = torch.zeros(8, dtype=torch.float64, device=self.dev)
            self.grad = torch.zeros_like(self.agent._flat)
            self.outs = _cabi.CovOutputs()
            self.outs.logp, self.outs.ent, self.outs.v = self.out[0].data_ptr(), self.out[1].data_ptr(), self.out[2].data_ptr()
            self.launch(torch.cuda.current_stream(self.dev).cuda_stream)     # warm (function attributes)
            torch.cuda.synchronize(self.dev)
            self.graph = None
            if self.capture() is None:
                self.slots = []
                break
            self.slots.append(dict(ws=self.ws, out=self.out, info=self.info, grad=self.grad, outs=self.outs, graph=self.graph, stream=streams[k % 2]))
        self.ws, self.out, self.info, self.grad, self.outs, self.graph = keep
        return len(self.slots)

    def pipelined_region(self, steps):
        """Enqueue `steps` complete steps round-robin over the slots; returns after joining the slot streams into the current one."""
        import torch
        import torch.distributed as dist
        cur = torch.cuda.current_stream(self.dev)
        used = {id(s['stream']): s['stream'] for s in self.slots}.values()
        for st in used:
            st.wait_stream(cur)
        for k in range(steps):
            s = self.slots[k % len(self.slots)]
            with torch.cuda.stream(s['stream']):
                s['graph'].replay()
                if self.world > 1:
                    dist.all_reduce(s['info'], op=dist.ReduceOp.SUM)
                    dist.all_reduce(s['grad'], op=dist.ReduceOp.SUM)
        for st in used:
            cur.wait_stream(st)


class Timer:
    def __init__(self, dev, world):
        import torch
        self.dev, self.world = dev, world
        self.flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def sync(self):
        import torch
        torch.cuda.synchronize(self.dev)
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
            torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        import torch
        import torch.distributed as dist
        t = torch.tensor([x], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def events(self, fn, steps, warmup, sampler=None):
        """Device time of `steps` calls of fn (CUDA events around each, L2 flushed between them), max over ranks."""
        import torch
        for _ in range(warmup):
            fn()
        self.sync()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        stops = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        if sampler:
            sampler.start()
        for s in range(steps):
            self.flush_buf.zero_()              # evict L2 between timed steps (outside the event pair)
            starts[s].record()
            fn()
            stops[s].record()
        self.sync()
        clocks = sampler.stop() if sampler else None
        total_ms = sum(a.elapsed_time(b) for a, b in zip(starts, stops))
        return self.max_over_ranks(total_ms), clocks

    def wall(self, fn, calls, warmup, flush=True):
        """Wall clock of `calls` calls of fn with a device synchronisation (and a barrier) on both sides, max over ranks; the L2
        flush between calls is inside the timed region."""
        for _ in range(warmup):
            fn()
        self.sync()
        t0 = time.perf_counter()
        for _ in range(calls):
            if flush:
                self.flush_buf.zero_()
            fn()
        self.sync()
        return self.max_over_ranks((time.perf_counter() - t0) * 1e3)


def e2e_epoch_fn(case, fused, tail=True, iters=1):
    """`iters` optimizer steps of the PPO update (one ppo.train call, as ppo.py:337 makes it) through the public API: molgym_b200.ppo.train (the restatement of ppo.py:99-160) over
    EPOCH_LEN minibatches of host observation tuples.  fused: compute_loss takes the fused CUDA-graph step and the optimizer is
    molgym_b200.optim.FlatAdam (gradient norm + clipping + Adam as two kernels); else the arithmetic of the reference's own
    compute_loss (agent.step + torch ops + autograd) with torch.optim.Adam, compute_gradient_norm and clip_grad_norm_ exactly as the
    unchanged ppo.py runs them.  tail=False: only the minibatch loop (compute_loss + backward)."""
    import torch
    from molgym_b200 import ppo
    from molgym_b200.optim import FlatAdam
    agent, data = case.agent, case.data
    if not hasattr(case, 'epoch_data'):
        case.epoch_data = {k: (v * EPOCH_LEN if isinstance(v, list) else np.concatenate([v] * EPOCH_LEN)) for k, v in data.items()}
        case.optimizers = {True: FlatAdam(agent, lr=LR, amsgrad=False),                                   # tools/util.py:197-205
                           False: torch.optim.Adam(torch.nn.Module.parameters(agent), lr=LR, amsgrad=False)}
    optimizer = case.optimizers[fused]
    n = case.n_global

    def epoch():
        agent.fused_ppo = fused
        np.random.seed(1234)   # get_batch_generator permutes with numpy's global generator: the same minibatches on every rank
        if tail:
            return ppo.train(agent, optimizer, case.epoch_data, mini_batch_size=n, clip_ratio=CLIP, target_kl=1e9, vf_coef=VF,
                             entropy_coef=ENT, gradient_clip=GRAD_CLIP, max_num_steps=iters)
        for _ in range(iters):
            optimizer.zero_grad()
            for idx in ppo.get_batch_generator(np.arange(len(case.epoch_data['obs'])), n):
                loss, _ = ppo.compute_loss(agent, ppo.collect_data_batch(case.epoch_data, idx), CLIP, VF, ENT)
                loss.backward()
    return epoch


def measure_case(case, timer, steps, warmup, sampler=None):
    """Device-resident ms per step of a case: (sequential ms, pipelined ms or None, mode, clocks).
    sequential: one CUDA graph (forward + loss + backward) replayed step after step, L2 flushed between steps;
    pipelined : the same graph captured on independent slots and replayed round-robin on two streams (forward of one minibatch
                beside the backward of the previous one); one CUDA-event pair around all the steps; the slots' combined workspaces
                exceed L2 (126 MB), which takes the place of the flush."""
    import torch
    for _ in range(3):
        case.eager_step()
    timer.sync()
    mode = 'eager launches'
    fn = case.eager_step
    if case.capture() is not None:
        fn, mode = case.graph_step, 'CUDA graph replay of the captured step'
    total_ms, clocks = timer.events(fn, steps, max(3, warmup), sampler)
    seq_ms = total_ms / steps
    pipe_ms = None
    if case.graph is not None and not os.environ.get('MOLGYM_B200_NO_PIPELINE'):
        free, _ = torch.cuda.mem_get_info(case.dev)
        l2 = 126 << 20
        n_slots = int(min(16, max(2, -(-2 * l2 // max(case.ws_bytes, 1)))))       # working set of the rotation >= 2 x L2
        n_slots = min(n_slots, int(0.45 * free // max(case.ws_bytes, 1)))           # the e2e path allocates its own pipeline slots later
        if n_slots >= 2 and case.make_slots(n_slots) >= 2:
            case.pipelined_region(max(2 * len(case.slots), warmup))
            timer.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            case.pipelined_region(steps)
            e1.record()
            timer.sync()
            pipe_ms = timer.max_over_ranks(e0.elapsed_time(e1)) / steps
            case.pipeline_slots = len(case.slots)
            case.slots = []
    return seq_ms, pipe_ms, mode, clocks


def run_ours(args):
    import torch
    import torch.distributed as dist
    from molgym_b200 import _lib, synth

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    lib = _lib.load()
    cfg = workload(args.workload, args.batch)
    weak = cfg.name.startswith('C2') or args.batch is not None   # the headline workload keeps its per-rank minibatch
    n_global = cfg.mini_batch_size * world if weak else cfg.mini_batch_size
    scaling = 'weak' if weak else 'strong'
    timer = Timer(dev, world)
    case = Case(cfg, n_global, world, rank, dev)
    B = case.B

    # ---- eager pass with every launch bracketed by CUDA events on its own stream (mgb_profile_kernel('k_')): the dominant
    # kernel's average launch duration and its share of the summed kernel time come from this timed region
    lib.mgb_profile_kernel(b'k_')
    for _ in range(max(3, args.warmup)):
        case.eager_step()
    torch.cuda.synchronize(dev)
    report = ctypes.create_string_buffer(16 << 20)
    lib.mgb_profile_report(report, len(report))   # drop warm-up timings
    prof_steps = min(args.steps, 50)
    launches_before = lib.mgb_launch_count()
    eager_ms, _ = timer.events(case.eager_step, prof_steps, 0)
    launches_per_step = (lib.mgb_launch_count() - launches_before) / prof_steps
    lib.mgb_profile_report(report, len(report))
    lib.mgb_profile_kernel(None)
    per_kernel = {}
    for line in report.value.decode().splitlines():
        name, ms = line.rsplit(' ', 1)
        short = name.split('<')[0].strip('( ')
        t = per_kernel.setdefault(short, [0.0, 0])
        t[0] += float(ms)
        t[1] += 1
    all_kernels_ms = sum(t[0] for t in per_kernel.values())
    k_total_ms, k_count = per_kernel.get(args.profile_kernel, [0.0, 0])

    # ---- value: CUDA-graph replay of the device-resident step
    sampler = ClockSampler(local_rank) if rank == 0 else None
    seq_ms, pipe_ms, mode, clocks = measure_case(case, timer, args.steps, args.warmup, sampler)
    if eager_ms / prof_steps < seq_ms:
        seq_ms, mode = eager_ms / prof_steps, 'eager launches'
    ms_per_step = seq_ms
    if pipe_ms is not None and pipe_ms < seq_ms:
        ms_per_step = pipe_ms
        mode = (f'CUDA graph replays of {getattr(case, "pipeline_slots", 2)} independent slots round-robin on two streams (consecutive minibatches '
                'of an epoch overlap), one event pair around all steps')
    value = n_global / (ms_per_step * 1e-3)

    # ---- e2e through the public API: wall clock over whole optimizer steps (>= 50 minibatches)
    calls = max(3, args.steps // (EPOCH_LEN * TRAIN_ITERS))
    per_call = EPOCH_LEN * TRAIN_ITERS
    e2e = {}
    for key, fused in (('e2e', True), ('e2e_unchanged_ppo', False)):
        fn = e2e_epoch_fn(case, fused, iters=TRAIN_ITERS)
        ms = timer.wall(fn, calls, 2, flush=True)
        ms_nf = timer.wall(fn, calls, 1, flush=False)
        ms_loop = timer.wall(e2e_epoch_fn(case, fused, tail=False, iters=TRAIN_ITERS), calls, 1, flush=False)
        per_mb, per_mb_nf = ms / (calls * per_call), ms_nf / (calls * per_call)
        e2e[key] = {'value': n_global / (per_mb * 1e-3), 'unit': 'canvases/s', 'ms_per_step': per_mb, 'steps': calls * per_call,
                    'no_flush_value': n_global / (per_mb_nf * 1e-3), 'no_flush_ms_per_step': per_mb_nf,
                    'minibatch_loop_only_ms_per_step': ms_loop / (calls * per_call)}
    e2e['e2e'].update({
        'h2d_bytes_per_step': int(case.h2d_bytes), 'd2h_bytes_per_step': 64,
        'timing': 'time.perf_counter() around the loop, device synchronised (+ barrier) on both sides, max over ranks; the 256 MiB L2 '
                  'flush per ppo.train call is INSIDE the timed region (no_flush_*: the same loop without it)',
        'note': f'molgym_b200.ppo.train (ppo.py:99-160 restated) on host observation tuples, called as ppo.py:337 calls it with the reference default '
                f'max_num_train_iters = {TRAIN_ITERS} optimizer steps per call; per optimizer step: zero_grad, {EPOCH_LEN} x (compute_loss -> pack into '
                'pinned staging, one H2D copy, CUDA-graph replays of forward + PPO loss and of the backward, 64-byte D2H of the loss info; '
                'loss.backward()), read of the loss infos (KL early-stop check, waits for the forward passes only), then FlatAdam: gradient norm, '
                'clipping + Adam update as two kernels (the norm is read back once per call, for the returned infos); consecutive minibatches '
                'alternate between two pipeline slots; data-parallel: one all-reduce of the pending info blocks and ONE gradient all-reduce per '
                'optimizer step; ms_per_step = time per minibatch'})
    e2e['e2e_unchanged_ppo'].update({
        'note': 'the same loop as the unchanged reference runs it: compute_loss = agent.step(obs, act) (pack, H2D, CUDA-graph replay on a '
                'persistent evaluation slot) + the loss as the torch ops of ppo.py:28-52, the six info numbers read back in one stacked copy (the reference issues six .item() calls), autograd backward (graph replay + one accumulate '
                'kernel); per optimizer step compute_gradient_norm (one torch.norm per parameter tensor, tools/util.py:61-69), '
                'clip_grad_norm_ and torch.optim.Adam; minibatch_loop_only_* leaves the optimizer tail out'})

    # ---- the configurations BASELINE.json names for 8 GPUs: fixed global minibatch sharded over the ranks
    per_config = {}
    if not args.no_per_config and args.workload == 'C2' and args.batch is None:
        del case.ws, case.graph
        case.agent._fused_cache.clear()
        case.agent._eval_cache.clear()
        torch.cuda.empty_cache()
        for name in ('C3', 'C4', 'C5'):
            c = synth.CONFIGS[name]
            g_batch = c.mini_batch_size
            note = None
            if name == 'C5' and world == 1:
                g_batch, note = 1024, 'one GPU runs the 1024-canvas shard an 8-GPU job gives it (the 8192-canvas minibatch is an 8-GPU configuration)'
            try:
                pc = Case(c, g_batch, world, rank, dev)
                n_steps = max(3, min(args.steps, 10 if name == 'C3' else 5))
                pc_seq, pc_pipe, pc_mode, _ = measure_case(pc, timer, n_steps, 3)
                ms = min(pc_seq, pc_pipe) if pc_pipe is not None else pc_seq
                entry = {'workload': c.name, 'global_batch': g_batch, 'per_rank_batch': pc.B, 'scaling': 'strong', 'n_gpus': world,
                         'ms_per_step': ms, 'value': g_batch / (ms * 1e-3), 'unit': 'canvases/s', 'steps': n_steps,
                         'sequential_ms_per_step': pc_seq, 'pipelined_ms_per_step': pc_pipe, 'launch_mode': pc_mode,
                         'workspace_gb': pc.ws_bytes / 1e9}
                # e2e for the same config: one epoch loop, fused step
                fn = e2e_epoch_fn(pc, True)
                e_epochs = 2
                e_ms = timer.wall(fn, e_epochs, 1, flush=False)
                entry['e2e'] = {'value': g_batch / (e_ms / (e_epochs * EPOCH_LEN) * 1e-3), 'unit': 'canvases/s',
                                'ms_per_step': e_ms / (e_epochs * EPOCH_LEN), 'steps': e_epochs * EPOCH_LEN,
                                'note': 'working set larger than L2: no flush needed'}
                if note:
                    entry['note'] = note
                per_config[name] = entry
                del pc
            except Exception as exc:   # pragma: no cover
                per_config[name] = {'error': repr(exc)[:300]}
            torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    n_timed = max(k_count, 1)
    k_ms = k_total_ms / n_timed
    k_launches_per_step = n_timed / prof_steps
    step_bytes, step_flops = atom_bwd_work(cfg, case.n_atoms_local, case.agent._cat_sizes, cg_term_counts(lib))
    alg_bytes = step_bytes / max(k_launches_per_step, 1.0)
    alg_flops = step_flops / max(k_launches_per_step, 1.0)
    tflops = alg_flops / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
    gbs = alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json'))).get(args.profile_kernel, {}).get(cfg.name)
    except Exception:
        pass
    top = sorted(per_kernel.items(), key=lambda kv: -kv[1][0])[:8]
    line = {
        'metric': METRIC, 'value': value, 'unit': 'canvases/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': scaling, 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'config': config_json(cfg, world, n_global, scaling), 'clocks': clocks,
        'e2e': e2e['e2e'], 'e2e_unchanged_ppo': e2e['e2e_unchanged_ppo'],
        'gpu_launches': int(round(launches_per_step * args.steps)), 'gpu_launches_per_step': launches_per_step,
        'roofline': {'bound': 'fp32', 'kernel': args.profile_kernel, 'achieved': tflops, 'peak': FFMA_PEAK_TFLOPS, 'unit': 'TFLOP/s',
                     'frac': tflops / FFMA_PEAK_TFLOPS, 'traffic': traffic,
                     'peak_source': 'FFMA-chain microbenchmark on this pool\'s B200 (tools/ffma_peak.cu -> profiles/r1_ffma_peak.txt, 72.3 TFLOP/s '
                                    'of nominal 74.4); MEASURED_PEAKS.json carries no fp32 figure (its bf16 tensor peak does not bound this kernel)',
                     'kernel_ms_per_launch': k_ms, 'kernel_launches_per_step': k_launches_per_step,
                     'kernel_share_of_step': k_total_ms / all_kernels_ms if all_kernels_ms > 0 else None,
                     'timed_in': 'eager pass of the timed region: CUDA events around every launch on its own stream; the share is the '
                                 "kernel's part of the summed kernel time (side-stream kernels overlap the main stream)",
                     'algorithmic_flops_per_launch': alg_flops, 'algorithmic_bytes_per_launch': alg_bytes,
                     'hbm': {'achieved': gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': gbs / hbm_peak,
                             'peak_source': 'MEASURED_PEAKS.json hbm_gbs (burst copy)' if peaks else 'fallback 6650 GB/s',
                             'traffic_over_algorithmic': (traffic / alg_bytes) if (traffic and alg_bytes) else None},
                     'note': 'SURVEY.md 8d: the Clebsch-Gordan atom kernels are FP32-FMA bound (arithmetic intensity >> ridge 11 flop/B); FLOPs '
                             'count the real (non-zero) Clebsch-Gordan terms, bytes count each layer-boundary tensor once (DESIGN.md section 3)'},
        'top_kernels': [{'kernel': k, 'share': v[0] / all_kernels_ms, 'ms_per_step': v[0] / prof_steps, 'launches_per_step': v[1] / prof_steps}
                        for k, v in top],
        'launch_mode': mode, 'eager_ms_per_step': eager_ms / prof_steps, 'sequential_ms_per_step': seq_ms, 'pipelined_ms_per_step': pipe_ms,
        'per_config': per_config,
    }
    if not args.no_cpu_baseline and world == 1:
        cpu = run_cpu(cfg, steps=5, warmup=1, budget_s=20.0)
        line['cpu_baseline'] = {'value': cpu['value'], 'unit': 'canvases/s', 'cores': cpu['cores'], 'kind': 'port', 'sample': cpu['sample']}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    """The reference's own CPU implementation of the path: here the oracle port (the reference's third-party dependencies are
    not installable and /root/reference does not travel to the GPU box), all host threads, rank 0 only."""
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if rank != 0:
        return
    cfg = workload(args.workload, args.batch)
    weak = cfg.name.startswith('C2') or args.batch is not None
    n_global = cfg.mini_batch_size * world if weak else cfg.mini_batch_size
    cpu = run_cpu(cfg, steps=args.steps, warmup=args.warmup, budget_s=60.0)
    line = {'impl': 'reference', 'metric': METRIC, 'value': cpu['value'], 'unit': 'canvases/s', 'n_gpus': world, 'steps': cpu['steps'],
            'warmup': cpu['warmup'], 'ms_per_step': cpu['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak' if weak else 'strong',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config_json(cfg, world, n_global, 'weak' if weak else 'strong'),
            'cpu_baseline': {'value': cpu['value'], 'unit': 'canvases/s', 'cores': cpu['cores'], 'kind': 'port', 'sample': cpu['sample']},
            'e2e': {'value': cpu['value'], 'unit': 'canvases/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


def run_internal(args):
    """--workload C1: the internal-coordinate (SchNet cfconv) agent, one GPU; same JSON shape, roofline for k_sch_bwd."""
    import torch
    from molgym_b200 import _lib, ppo, synth
    from molgym_b200.agents.internal.agent import SchNetAC
    from molgym_b200.spaces import ActionSpace, ObservationSpace
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0')))
    torch.cuda.set_device(dev)
    lib = _lib.load()
    cfg = workload('C1', args.batch)
    B = cfg.mini_batch_size
    torch.manual_seed(0)
    agent = SchNetAC(ObservationSpace(cfg.canvas_size, cfg.zs), ActionSpace(cfg.zs), device=dev, **cfg.agent_kwargs())
    obs, n = synth.make_observations(cfg, batch=B)
    act = synth.make_actions(cfg, obs, n)
    with torch.no_grad():
        logp0 = agent.step(obs, act)['logp'].cpu().numpy()
    old_logp, adv, ret = synth.make_ppo_targets(cfg, logp0)
    data = dict(obs=obs, act=act, logp=old_logp, adv=adv, ret=ret)
    timer = Timer(dev, 1)

    def step():
        agent.zero_grad()
        loss, _ = ppo.compute_loss(agent, data, CLIP, VF, ENT)
        loss.backward()

    lib.mgb_profile_kernel(b'k_')
    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize(dev)
    report = ctypes.create_string_buffer(8 << 20)
    lib.mgb_profile_report(report, len(report))
    prof_steps = min(args.steps, 50)
    before = lib.mgb_launch_count()
    total_ms, clocks = timer.events(step, prof_steps, 0, ClockSampler(dev.index))
    launches = (lib.mgb_launch_count() - before) / prof_steps
    lib.mgb_profile_report(report, len(report))
    lib.mgb_profile_kernel(None)
    per_kernel = {}
    for line in report.value.decode().splitlines():
        name, ms = line.rsplit(' ', 1)
        t = per_kernel.setdefault(name.split('<')[0].strip('( '), [0.0, 0])
        t[0] += float(ms)
        t[1] += 1
    allk = sum(t[0] for t in per_kernel.values())
    wall_ms = timer.wall(step, args.steps, 3, flush=True)
    # work model of the cfconv kernels: per molecule of n atoms and interaction: filter MLP 25->128->128 per ordered
    # pair (2 * (25*128 + 128*128) flop), in2f / f2out / dense per atom; three molecules per canvas (n, n+1, n+1 atoms); backward 2x
    flops = 0
    for k in n:
        for m in (int(k), int(k) + 1, int(k) + 1):
            flops += 3 * 3 * (m * max(m - 1, 0) * 2 * (25 * 128 + 128 * 128) + m * 2 * (64 * 128 + 128 * 64 + 64 * 64))
    sch = per_kernel.get('k_sch_bwd', [0.0, 1])
    sch_f = per_kernel.get('k_sch_fwd', [0.0, 1])
    sch_ms = (sch[0] + sch_f[0]) / prof_steps
    tflops = flops / (sch_ms * 1e-3) / 1e12 if sch_ms > 0 else 0.0
    line = {'metric': METRIC, 'value': B / (total_ms / prof_steps * 1e-3), 'unit': 'canvases/s', 'n_gpus': 1, 'steps': prof_steps, 'warmup': args.warmup,
            'ms_per_step': total_ms / prof_steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'{cfg.name}: internal-coordinate (SchNet) agent, canvas_size={cfg.canvas_size}, mini_batch_size={B}',
                       'l2': 'flushed between timed steps'},
            'clocks': clocks, 'gpu_launches': int(launches * prof_steps), 'gpu_launches_per_step': launches,
            'e2e': {'value': B / (wall_ms / args.steps * 1e-3), 'unit': 'canvases/s', 'ms_per_step': wall_ms / args.steps,
                    'h2d_bytes_per_step': int(B * cfg.canvas_size * 3 * (4 + 12) + B * 7 * 4), 'd2h_bytes_per_step': 48,
                    'note': 'agent.step(obs, act) + torch PPO loss + backward on host observation tuples (z-matrix placement on the host)'},
            'roofline': {'bound': 'fp32', 'kernel': 'k_sch_fwd + k_sch_bwd', 'achieved': tflops, 'peak': FFMA_PEAK_TFLOPS, 'unit': 'TFLOP/s',
                         'frac': tflops / FFMA_PEAK_TFLOPS, 'traffic': None, 'kernel_ms_per_step': sch_ms,
                         'kernel_share_of_step': (sch[0] + sch_f[0]) / allk if allk else None, 'algorithmic_flops_per_step': flops,
                         'note': 'C1 is the reference\'s CPU plumbing configuration: 28 canvases x 3 molecules of <= 7 atoms, one CTA per molecule; '
                                 'latency-bound at this size'},
            'top_kernels': [{'kernel': k, 'share': v[0] / allk, 'ms_per_step': v[0] / prof_steps} for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1][0])[:6]]}
    print(json.dumps(line), flush=True)


if __name__ == '__main__':
    a = parse_args()
    if a.impl == 'reference':
        run_reference(a)
    elif a.workload == 'C1':
        run_internal(a)
    else:
        run_ours(a)
