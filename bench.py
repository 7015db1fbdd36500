"""bench.py — PPO-minibatch forward+backward throughput of the covariant agent (canvases / s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2] [--impl ours|reference]
    (MOLGYM_B200_SLOTS=n: number of independent slots of the `two_slot` figure, default 2; MOLGYM_B200_NO_GRAPH=1: eager launches only)

A "step" is one pass of the hot path over one minibatch of synthetic canvases: forward (Cormorant body + heads) ->
PPO-clip loss -> backward -> (N > 1) gradient all-reduce.  Weak scaling: every rank processes its own minibatch of the
workload's size; `value` = canvases processed by all ranks / max-over-ranks device time.

  value : inputs resident in HBM, direct C-ABI calls (mgb_cov_forward, mgb_ppo_loss, mgb_cov_backward), CUDA events.
  e2e   : the reference-facing call — CovariantAC.step(list of observation tuples, actions) driven by the restated
          ppo.compute_loss, loss.backward(), and a device->host read of the loss info — host packing and H2D copies inside.
  roofline / cpu_baseline : see DESIGN.md.
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
METRIC = 'ppo_minibatch_fwd_bwd_canvases_per_sec'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--workload', default='C2')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=None, help='override the minibatch size (per rank)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--profile-kernel', default='k_atom_bwd')
    return ap.parse_args()


def workload(name, batch=None):
    from molgym_b200 import synth
    cfg = synth.CONFIGS[name]
    if batch:
        cfg = dataclasses.replace(cfg, mini_batch_size=batch)
    return cfg


def config_json(cfg, world):
    return {'workload': f'{cfg.name}: canvas_size={cfg.canvas_size} zs={cfg.zs} mini_batch_size={cfg.mini_batch_size} per rank, '
                        f'occupancies 0..K-1 uniform (synthetic PPO buffer)',
            'global_batch': cfg.mini_batch_size * world, 'canvas_size': cfg.canvas_size,
            'hyper': {'network_width': cfg.network_width, 'maxl': cfg.maxl, 'num_cg_levels': cfg.num_cg_levels,
                      'num_channels_hidden': cfg.num_channels_hidden, 'num_channels_per_element': cfg.num_channels_per_element,
                      'num_gaussians': cfg.num_gaussians, 'beta': cfg.beta},
            'parallelism': f'dp{world}', 'l2': 'flushed between timed steps (256 MiB write)'}


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
# algorithmic work model (DESIGN.md "Work model")
# ----------------------------------------------------------------------------------------------------------------
def atom_bwd_work(cfg, n_atoms, cat_sizes):
    """Algorithmic HBM bytes and FLOPs of ALL k_atom_bwd launches of one step (one per CG level; two half kernels per level
    for small minibatches), from the kernel's own decomposition.  Per valid atom i of a canvas with n atoms,
    level k with nlm2 input components (1 at level 0, else 25), C channels and a cat vector of totA_k complex entries:
      bytes = totA_k*8 (dcat slice) + nlm2*C*8 (A_i) + n * [nlm2*C*8 (A_j) + 5*C*8 (E_ij) + 5*C*8 (dE_ij) + 2*nlm2*C*8 (dA_j RMW)]
      flops = n * 2 * 25*nlm2 * C * 8 (row + column Kronecker passes, complex MAC = 8 flop)
              + 3 * n_pairs(k) * 5 * C * 4 (padded Clebsch-Gordan scatter: row, column, square; real coefficient x complex)"""
    C, nl = cfg.num_channels_hidden, cfg.maxl + 1
    bytes_, flops = 0, 0
    for k in range(cfg.num_cg_levels):
        nlm2 = 1 if k == 0 else 25
        tot_a = sum(cat_sizes[(k * nl + l) * 2 + 1] * (2 * l + 1) for l in range(nl))
        for n in n_atoms:
            n = int(n)
            bytes_ += n * (tot_a * 8 + nlm2 * C * 8 + n * (nlm2 * C * 8 + 5 * C * 8 + 5 * C * 8 + 2 * nlm2 * C * 8))
            flops += n * (n * 2 * 25 * nlm2 * C * 8 + 3 * 25 * nlm2 * 5 * C * 4)
    return bytes_, flops


# ----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from molgym_b200 import _cabi, _lib, parallel, ppo, synth
    from molgym_b200.agents.covariant.agent import CovariantAC
    from molgym_b200.spaces import ActionSpace, ObservationSpace

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    lib = _lib.load()
    cfg = workload(args.workload, args.batch)
    B = cfg.mini_batch_size

    torch.manual_seed(0)
    agent = CovariantAC(ObservationSpace(cfg.canvas_size, cfg.zs), ActionSpace(cfg.zs), device=dev, **cfg.agent_kwargs())
    if world > 1:
        parallel.shard_agent(agent)
    obs, n_atoms = synth.make_observations(cfg, batch=B, seed=cfg.seed + 17 * rank, start_index=rank)
    act = synth.make_actions(cfg, obs, n_atoms, seed=cfg.seed + 17 * rank)
    with torch.no_grad():
        logp0 = agent.step(obs, act)['logp'].cpu().numpy()
    old_logp, adv, ret = synth.make_ppo_targets(cfg, logp0, seed=cfg.seed + 17 * rank)
    data = dict(obs=obs, act=act, logp=old_logp, adv=adv, ret=ret)

    # ---- device-resident inputs for the `value` measurement
    parsed = agent.parse_observations(obs)
    pos, charges, bags = parsed['positions'], parsed['charges'], parsed['bags']
    act_d = torch.as_tensor(act, dtype=torch.float32, device=dev)
    old_d = torch.as_tensor(old_logp, device=dev)
    adv_d = torch.as_tensor(adv, device=dev)
    ret_d = torch.as_tensor(ret, device=dev)
    f32 = dict(dtype=torch.float32, device=dev)
    logp, ent, v = torch.empty(B, **f32), torch.empty(B, **f32), torch.empty(B, **f32)
    g_logp, g_ent, g_v = torch.empty(B, **f32), torch.empty(B, **f32), torch.empty(B, **f32)
    info = torch.zeros(8, dtype=torch.float64, device=dev)
    grad = torch.zeros_like(agent._flat)
    ws = torch.empty(lib.mgb_cov_workspace_bytes(agent._plan, B), dtype=torch.uint8, device=dev)
    outs = _cabi.CovOutputs()
    outs.logp, outs.ent, outs.v = logp.data_ptr(), ent.data_ptr(), v.data_ptr()
    stream = torch.cuda.current_stream(dev).cuda_stream
    inv_global = 1.0 / (B * world)

    def device_step():
        _cabi.check(lib, lib.mgb_cov_forward(agent._plan, B, pos.data_ptr(), charges.data_ptr(), bags.data_ptr(), act_d.data_ptr(),
                                             agent._flat.data_ptr(), ws.data_ptr(), ws.numel(), ctypes.byref(outs), stream))
        _cabi.check(lib, lib.mgb_ppo_loss(B, logp.data_ptr(), ent.data_ptr(), v.data_ptr(), old_d.data_ptr(), adv_d.data_ptr(),
                                          ret_d.data_ptr(), CLIP, VF, ENT, inv_global, info.data_ptr(), g_logp.data_ptr(),
                                          g_ent.data_ptr(), g_v.data_ptr(), stream))
        _cabi.check(lib, lib.mgb_cov_backward(agent._plan, B, pos.data_ptr(), charges.data_ptr(), bags.data_ptr(), act_d.data_ptr(),
                                              agent._flat.data_ptr(), ws.data_ptr(), ws.numel(), g_logp.data_ptr(),
                                              g_ent.data_ptr(), g_v.data_ptr(), grad.data_ptr(), 0, stream))
        if world > 1:
            dist.all_reduce(grad, op=dist.ReduceOp.SUM)

    # The device-resident step works on fixed buffers, so its launches + memsets can be captured once into a CUDA graph and
    # replayed (the gradient all-reduce stays outside the graph).  MOLGYM_B200_NO_GRAPH=1 keeps only the eager launches.
    def capture(ws_, outs_, logp_, ent_, v_, g3_, info_, grad_):
        g = torch.cuda.CUDAGraph()
        cap_stream = torch.cuda.Stream(dev, priority=-5)   # main-chain kernels outrank the weight-gradient side streams
        cap_stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(cap_stream):
            with torch.cuda.graph(g, stream=cap_stream):
                s_ptr = torch.cuda.current_stream(dev).cuda_stream
                _cabi.check(lib, lib.mgb_cov_forward(agent._plan, B, pos.data_ptr(), charges.data_ptr(), bags.data_ptr(),
                                                     act_d.data_ptr(), agent._flat.data_ptr(), ws_.data_ptr(), ws_.numel(),
                                                     ctypes.byref(outs_), s_ptr))
                _cabi.check(lib, lib.mgb_ppo_loss(B, logp_.data_ptr(), ent_.data_ptr(), v_.data_ptr(), old_d.data_ptr(),
                                                  adv_d.data_ptr(), ret_d.data_ptr(), CLIP, VF, ENT, inv_global, info_.data_ptr(),
                                                  g3_[0].data_ptr(), g3_[1].data_ptr(), g3_[2].data_ptr(), s_ptr))
                _cabi.check(lib, lib.mgb_cov_backward(agent._plan, B, pos.data_ptr(), charges.data_ptr(), bags.data_ptr(),
                                                      act_d.data_ptr(), agent._flat.data_ptr(), ws_.data_ptr(), ws_.numel(),
                                                      g3_[0].data_ptr(), g3_[1].data_ptr(), g3_[2].data_ptr(), grad_.data_ptr(), 0, s_ptr))
        torch.cuda.current_stream(dev).wait_stream(cap_stream)
        return g

    graph, graph2, grad2, extra = None, None, None, []
    if not os.environ.get('MOLGYM_B200_NO_GRAPH'):
        try:
            graph = capture(ws, outs, logp, ent, v, (g_logp, g_ent, g_v), info, grad)
            # more independent slots (own workspace / outputs / gradient) for the multi-slot throughput figure below
            extra = []
            n_extra = int(os.environ.get('MOLGYM_B200_SLOTS', '2')) - 1
            if ws.numel() * (n_extra + 3) > 0.5 * torch.cuda.get_device_properties(dev).total_memory:
                n_extra = 0   # the agent's own fused slots need their workspaces too
            for _ in range(n_extra):
                ws2 = torch.empty_like(ws)
                o2 = [torch.empty(B, **f32) for _ in range(6)]
                outs2 = _cabi.CovOutputs()
                outs2.logp, outs2.ent, outs2.v = o2[0].data_ptr(), o2[1].data_ptr(), o2[2].data_ptr()
                info2, grad2 = torch.zeros_like(info), torch.zeros_like(grad)
                graph2 = capture(ws2, outs2, o2[0], o2[1], o2[2], o2[3:], info2, grad2)
                extra.append((graph2, grad2, (ws2, o2, outs2, info2)))
        except Exception as exc:   # pragma: no cover
            sys.stderr.write(f'CUDA graph capture failed ({exc}); eager launches only\n')
            graph = graph2 = None

    def graph_step():
        graph.replay()
        if world > 1:
            dist.all_reduce(grad, op=dist.ReduceOp.SUM)

    # ppo.train's inner loop (ppo.py:118-131): optimizer.zero_grad() once per epoch, then every minibatch of the epoch runs
    # compute_loss + backward and the gradients accumulate; EPOCH_LEN minibatches per epoch here
    EPOCH_LEN = 4
    e2e_count = [0]

    def e2e_step():
        if e2e_count[0] % EPOCH_LEN == 0:
            agent.zero_grad()
        e2e_count[0] += 1
        loss, info_d = ppo.compute_loss(agent, data, CLIP, VF, ENT)
        (loss / world).backward()
        return info_d

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn, steps, warmup, sampler=None):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        stops = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        if sampler:
            sampler.start()
        wall0 = time.perf_counter()
        for s in range(steps):
            flush_buf.zero_()              # evict L2 between timed steps (outside the event pair)
            starts[s].record()
            fn()
            stops[s].record()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        wall = time.perf_counter() - wall0
        clocks = sampler.stop() if sampler else None
        total_ms = sum(a.elapsed_time(b) for a, b in zip(starts, stops))
        if world > 1:
            t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms, wall, clocks

    # ---- value (device-resident): eager pass with the dominant kernel timed live, then the CUDA-graph replay of the same step
    # every launch of the eager pass is bracketed by CUDA events on the stream it is launched on (mgb_profile_kernel('k_')):
    # the dominant kernel's average launch duration and its share of the summed kernel time come from this timed region
    lib.mgb_profile_kernel(b'k_')
    for _ in range(args.warmup):
        device_step()
    torch.cuda.synchronize(dev)
    report = ctypes.create_string_buffer(8 << 20)
    lib.mgb_profile_report(report, len(report))   # drop warm-up timings
    launches_before = lib.mgb_launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    eager_ms, wall, clocks = timed(device_step, args.steps, 0, sampler)
    launches = lib.mgb_launch_count() - launches_before
    lib.mgb_profile_report(report, len(report))
    lib.mgb_profile_kernel(None)
    k_total_ms, k_count, all_kernels_ms = 0.0, 0, 0.0
    for line in report.value.decode().splitlines():
        name, ms = line.rsplit(' ', 1)
        all_kernels_ms += float(ms)
        if args.profile_kernel in name:
            k_total_ms += float(ms)
            k_count += 1
    total_ms, mode = eager_ms, 'eager launches'
    graph_ms = None
    if graph is not None:
        sampler2 = ClockSampler(local_rank) if rank == 0 else None
        graph_ms, wall_g, clocks_g = timed(graph_step, args.steps, max(3, args.warmup), sampler2)
        if graph_ms < eager_ms:
            total_ms, wall, clocks, mode = graph_ms, wall_g, clocks_g, 'CUDA graph replay of the captured step'
    ms_per_step = total_ms / args.steps
    value = B * world / (ms_per_step * 1e-3)

    # ---- the same device-resident step with TWO independent slots replayed alternately on two streams (what the e2e path
    # does with consecutive minibatches of an epoch): whole region timed, no L2 flush inside it (reported beside `value`)
    two_slot_ms = None
    if extra:
        slots = [(graph, grad, torch.cuda.Stream(dev))] + [(g_, gr_, torch.cuda.Stream(dev)) for g_, gr_, _ in extra]

        def two_slot_region(n):
            cur = torch.cuda.current_stream(dev)
            for _, _, st_ in slots:
                st_.wait_stream(cur)
            for s_ in range(n):
                g_, gr_, st_ = slots[s_ % len(slots)]
                with torch.cuda.stream(st_):
                    g_.replay()
                    if world > 1:
                        dist.all_reduce(gr_, op=dist.ReduceOp.SUM)
            for _, _, st_ in slots:
                cur.wait_stream(st_)

        two_slot_region(max(4, args.warmup))
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t_host = time.perf_counter()
        two_slot_region(args.steps)
        t_host = time.perf_counter() - t_host   # host time spent enqueuing (cudaGraphLaunch): the floor of any replay-based loop
        e1.record()
        torch.cuda.synchronize(dev)
        two_slot_ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([two_slot_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            two_slot_ms = float(t.item())

    # ---- e2e through the public API (host observations, packing, H2D, torch loss, D2H of the loss info)
    e2e_steps = max(10, args.steps // 4)
    e2e_ms, e2e_wall, _ = timed(e2e_step, e2e_steps, max(3, args.warmup // 4))
    # the e2e step has host work outside the event pairs' GPU time only if the GPU idles; events bracket the call, so
    # host time shows up as the gap between start.record() and the first kernel: elapsed_time includes it.
    e2e_value = B * world / (e2e_ms / e2e_steps * 1e-3)
    h2d = pos.numel() * 4 + charges.numel() * 4 + bags.numel() * 4 + act_d.numel() * 4 + old_d.numel() * 4 + adv_d.numel() * 8 + ret_d.numel() * 8
    d2h = 8 * 8   # the loss-info block (8 doubles) read back every step

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
    launches_per_step_of_kernel = n_timed / args.steps
    # per-launch algorithmic work: the step's total over the kernel's launches in one step (levels x half kernels)
    step_bytes, step_flops = atom_bwd_work(cfg, n_atoms, agent._cat_sizes)
    alg_bytes = step_bytes / max(launches_per_step_of_kernel, 1.0)
    alg_flops = step_flops / max(launches_per_step_of_kernel, 1.0)
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json'))).get(args.profile_kernel, {}).get(cfg.name)
    except Exception:
        pass
    line = {
        'metric': METRIC, 'value': value, 'unit': 'canvases/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'config': config_json(cfg, world), 'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': 'canvases/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                'ms_per_step': e2e_ms / e2e_steps, 'steps': e2e_steps,
                'note': 'ppo.train inner loop (ppo.py:118-131): zero_grad once per epoch of 4 minibatches, then compute_loss + '
                        'loss.backward() per minibatch on host observation tuples; consecutive minibatches alternate between two '
                        'pipeline slots, so the forward of step i+1 overlaps the backward of step i (the parameters are fixed within '
                        'a PPO epoch); `value` is the strictly sequential device-resident step'},
        'gpu_launches': int(launches),
        'roofline': {'bound': 'hbm', 'kernel': args.profile_kernel, 'achieved': achieved, 'peak': hbm_peak, 'unit': 'GB/s',
                     'frac': achieved / hbm_peak, 'traffic': traffic,
                     'peak_source': 'MEASURED_PEAKS.json hbm_gbs (burst copy)' if peaks else 'fallback 6650 GB/s',
                     'kernel_ms_per_launch': k_ms, 'kernel_launches_per_step': launches_per_step_of_kernel,
                     'kernel_share_of_step': k_total_ms / all_kernels_ms if all_kernels_ms > 0 else None,
                     'timed_in': 'eager pass of the timed region: CUDA events around every launch on its own stream; the share is the '
                                 "kernel's part of the summed kernel time (side-stream kernels overlap the main stream)",
                     'algorithmic_bytes_per_launch': alg_bytes, 'algorithmic_flops_per_launch': alg_flops,
                     'fp32_achieved_tflops': alg_flops / (k_ms * 1e-3) / 1e12 if k_ms > 0 else None,
                     'fp32_peak_tflops': 72.3, 'fp32_frac': (alg_flops / (k_ms * 1e-3) / 1e12 / 72.3) if k_ms > 0 else None,
                     'note': 'the CG kernels are FP32-FMA bound (arithmetic intensity >> ridge 11 flop/B): the binding roof is the '
                             'measured FFMA peak (profiles/r1_ffma_peak.txt), reported beside the HBM figure BASELINE.json asks for; '
                             'see DESIGN.md section 3'},
        'wall_ms_per_step': wall / args.steps * 1e3, 'launch_mode': mode, 'eager_ms_per_step': eager_ms / args.steps,
        'graph_ms_per_step': graph_ms / args.steps if graph_ms is not None else None,
        'two_slot': None if two_slot_ms is None else {
            'ms_per_step': two_slot_ms / args.steps, 'value': B * world / (two_slot_ms / args.steps * 1e-3), 'unit': 'canvases/s',
            'host_enqueue_ms_per_step': t_host / args.steps * 1e3,
            'note': 'device-resident steps of two independent slots replayed alternately on two streams (forward of one beside the '
                    'backward of the other), whole region timed without L2 flushes; `value` above is the strictly sequential step'},
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
    cpu = run_cpu(cfg, steps=args.steps, warmup=args.warmup, budget_s=60.0)
    line = {'impl': 'reference', 'metric': METRIC, 'value': cpu['value'], 'unit': 'canvases/s', 'n_gpus': world, 'steps': cpu['steps'],
            'warmup': cpu['warmup'], 'ms_per_step': cpu['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic', 'config': config_json(cfg, world),
            'cpu_baseline': {'value': cpu['value'], 'unit': 'canvases/s', 'cores': cpu['cores'], 'kind': 'port', 'sample': cpu['sample']},
            'e2e': {'value': cpu['value'], 'unit': 'canvases/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


if __name__ == '__main__':
    a = parse_args()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
