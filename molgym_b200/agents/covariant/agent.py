"""CovariantAC — the reference's covariant actor-critic (molgym/agents/covariant/agent.py:20-334) behind the same
constructor and `step(observations, actions)` contract, computed by the hand-written sm_100a kernels of
molgym_b200/csrc through the C ABI in include/molgym_b200.h.

What stays Python: parameter ownership (ordinary nn.Parameters named exactly like the reference's, so state_dicts,
optimizers, clipping and whole-module pickles interchange), observation packing, autograd glue, and rollout-mode
sampling.  There is no CPU fallback: without the CUDA library / a CUDA device construction fails.
"""
import ctypes
import math
import weakref
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from molgym_b200 import _cabi
from molgym_b200.agents._flat import FlatParamMixin
from molgym_b200.agents._runtime import CudaRuntime
from molgym_b200.agents.base import AbstractActorCritic
from molgym_b200.agents.covariant import sampling
from molgym_b200.agents.covariant.packing import pack_observations

_LEBEDEV = None


def _lebedev_071():
    """quadpy.u3._lebedev.lebedev_071() (spherical_dists.py:209): 1730 points, weights summing to 1."""
    global _LEBEDEV
    if _LEBEDEV is None:
        from scipy.integrate import lebedev_rule
        pts, w = lebedev_rule(71)
        _LEBEDEV = (np.ascontiguousarray(pts.T, dtype=np.float64), np.ascontiguousarray(w / (4 * np.pi), dtype=np.float64))
    return _LEBEDEV


class _CovEvaluate(torch.autograd.Function):
    """logp, ent, v = f(parameters; canvases, actions).  Parameter gradients are written by the CUDA backward straight into
    the agent's flat gradient buffer, of which every `p.grad` is a view; autograd only routes the three cotangents."""

    @staticmethod
    def forward(ctx, anchor, agent, pos, charges, bags, actions):
        outs, ws = agent._forward_raw(pos, charges, bags, actions, want_extras=True)
        ctx.agent = agent
        ctx.saved = (pos, charges, bags, actions, ws)
        ctx.mark_non_differentiable(*outs[3:])
        return outs

    @staticmethod
    def backward(ctx, g_logp, g_ent, g_v, *unused):
        agent = ctx.agent
        pos, charges, bags, actions, ws = ctx.saved
        B = pos.shape[0]

        def prep(g):
            if g is None:
                return torch.zeros(B, dtype=torch.float32, device=pos.device)
            return g.to(torch.float32).contiguous()

        agent._backward_raw(pos, charges, bags, actions, ws, prep(g_logp), prep(g_ent), prep(g_v))
        return (None, ) * 6


class _CovEvaluateSlot(torch.autograd.Function):
    """The same function evaluated by CUDA-graph replays on the persistent buffers of an evaluation slot (the path the
    reference's unchanged ppo.compute_loss takes: agent.step(obs, act) -> torch loss -> loss.backward())."""

    @staticmethod
    def forward(ctx, anchor, agent, st):
        st.g_forward.replay()
        outs = agent._slot_outputs(st, st.out_all.clone())   # one copy: the slot's buffers are reused by the next step
        ctx.agent, ctx.st, ctx.generation = agent, st, st.generation
        st.live, st.done = weakref.ref(ctx), False   # the slot is busy until this node ran its backward or was dropped
        ctx.mark_non_differentiable(*outs[3:])
        return outs

    @staticmethod
    def backward(ctx, g_logp, g_ent, g_v, *unused):
        st, agent = ctx.st, ctx.agent
        if st.generation != ctx.generation:
            raise RuntimeError('internal: evaluation slot reused before its backward ran')
        for dst, g in zip(st.cot, (g_logp, g_ent, g_v)):
            if g is None:
                dst.zero_()
            else:
                dst.copy_(g)
        st.g_backward.replay()
        agent._accumulate_scratch(st.grad, agent._one, agent._rt.current_stream())
        st.done = True
        return None, None, None


class _FusedPPOLoss(torch.autograd.Function):
    """The scalar PPO loss of CovariantAC.fused_ppo_loss.  Its parameter gradient already sits in the step's scratch buffer
    (the backward graph was enqueued right behind the forward); backward() scales it by the incoming cotangent into `.grad`."""

    @staticmethod
    def forward(ctx, anchor, agent, state, loss):
        ctx.agent, ctx.state, ctx.generation = agent, state, state.generation
        return loss

    @staticmethod
    def backward(ctx, g):
        if ctx.state.generation != ctx.generation:
            raise RuntimeError('this PPO loss was superseded by a later compute_loss on the same minibatch size before it was '
                               'differentiated; set agent.fused_ppo = False to keep several losses alive')
        ctx.agent._fused_backward(ctx.state, g)
        return None, None, None, None


class _StepState:
    """Persistent buffers + captured CUDA graphs of one minibatch size (fused PPO step or evaluation slot)."""
    pass


_INFO_RING = 16   # loss-info blocks in flight per pipeline slot before the oldest must be read out
_INFO_KEYS = ('total_loss', 'policy_loss', 'entropy_loss', 'vf_loss', 'approx_kl', 'clip_fraction')


class LazyLossInfo(dict):
    """The loss info of ppo.compute_loss (ppo.py:54-61: policy_loss, entropy_loss, vf_loss, total_loss, approx_kl, clip_fraction)
    whose six numbers are fetched from the device when they are first READ.  ppo.train only reads them after the minibatch loop
    (compute_mean_dict, ppo.py:133), so the host does not wait for every forward pass and consecutive minibatches really overlap.
    A plain dict once resolved; every reading access resolves.

    Data-parallel agents: each rank's block holds its shard's share of the global means; reading ANY pending info sums all
    pending blocks of the agent over the ranks in one all-reduce (a collective: every rank reads its infos at the same point of
    the same loop, as ppo.train does), so the ranks meet once per optimizer step instead of once per minibatch."""

    def __init__(self, agent, host_block, event, sharded):
        super().__init__()
        self._pending = (agent, host_block, event, sharded)
        if sharded:
            agent._pending_infos.append(self)

    def _fill(self, vals):
        self._pending = None
        for i, key in enumerate(_INFO_KEYS):
            dict.__setitem__(self, key, float(vals[i]))

    def _resolve(self):
        if self._pending is None:
            return self
        agent, host_block, event, sharded = self._pending
        if not sharded:
            event.synchronize()
            self._fill(host_block.numpy())
            return self
        todo, agent._pending_infos = agent._pending_infos, []
        for info in todo:
            info._pending[2].synchronize()
        blocks = torch.stack([info._pending[1] for info in todo])            # [k, 8] float64 (host)
        if agent._rt.is_cuda:
            # on a stream of its own: the caller's stream is queued behind the backward passes (gradient accumulation), the loss
            # numbers are not — the ranks exchange them while the last backward is still running
            if getattr(agent, '_info_stream', None) is None:
                agent._info_stream = agent._rt.new_stream()
            with agent._rt.stream_ctx(agent._info_stream):
                dev = blocks.to(agent.device)
                torch.distributed.all_reduce(dev, op=torch.distributed.ReduceOp.SUM)
                vals = dev.cpu().numpy()
        else:
            dev = blocks.clone()
            torch.distributed.all_reduce(dev, op=torch.distributed.ReduceOp.SUM)
            vals = dev.numpy()
        for info, row in zip(todo, vals):
            info._fill(row)
        return self

    def __getitem__(self, key):
        return dict.__getitem__(self._resolve(), key)

    def __iter__(self):
        return dict.__iter__(self._resolve())

    def __len__(self):
        return dict.__len__(self._resolve())

    def __contains__(self, key):
        return dict.__contains__(self._resolve(), key)

    def __eq__(self, other):
        return dict.__eq__(self._resolve(), other._resolve() if isinstance(other, LazyLossInfo) else other)

    def __ne__(self, other):
        return not self.__eq__(other)

    def __repr__(self):
        return dict.__repr__(self._resolve())

    def keys(self):
        return dict.keys(self._resolve())

    def values(self):
        return dict.values(self._resolve())

    def items(self):
        return dict.items(self._resolve())

    def get(self, key, default=None):
        return dict.get(self._resolve(), key, default)

    def copy(self):
        return dict(self._resolve())

    def __reduce__(self):
        return (dict, (dict(self._resolve()), ))


def _as_numpy_actions(actions, n: int, width: int) -> np.ndarray:
    a = np.asarray(actions.detach().cpu().numpy() if torch.is_tensor(actions) else actions, dtype=np.float32)
    assert a.shape == (n, width)
    return a


class CovariantAC(FlatParamMixin, AbstractActorCritic):
    _runtime_cls = CudaRuntime   # the test-suite's emulator-backed subclass substitutes a host runtime

    def __init__(
        self,
        observation_space,
        action_space,
        min_max_distance: Tuple[float, float],
        network_width: int,
        maxl: int,
        num_cg_levels: int,
        num_channels_hidden: int,
        num_channels_per_element: int,
        num_gaussians: int,
        bag_scale: int,
        beta: Optional[float] = None,
        device=None,
    ):
        super().__init__(observation_space, action_space)
        self._rt = self._runtime_cls(device)
        self.device = self._rt.device
        self.dtype = torch.float
        self.zs = list(self.observation_space.zs)
        self.min_distance, self.max_distance = min_max_distance
        assert self.min_distance < self.max_distance
        self.beta = beta
        self.max_sh = maxl
        self.num_cg_levels = num_cg_levels
        self.num_channels_hidden = num_channels_hidden
        self.num_channels_per_element = num_channels_per_element
        self.num_gaussians = num_gaussians
        self.num_channels_out = len(self.zs) * num_channels_per_element
        self.network_width = network_width
        self.bag_scale = bag_scale
        self.canvas_size = self.observation_space.canvas_space.size
        self.data_parallel = False   # set by molgym_b200.parallel.shard_agent
        self.fused_ppo = True        # molgym_b200.ppo.compute_loss may use fused_ppo_loss (CUDA-graph replay of the whole step)
        self.fused_sync_params = False   # True: the fused step waits for the caller's stream every time (see fused_ppo_loss)
        self.graph_evaluate = True   # evaluate-mode step() under autograd replays CUDA graphs on persistent slots
        self.fused_lazy_info = True  # fused_ppo_loss returns a loss-info dict that reads the device when first accessed
        self._check_supported()
        self._init_native()
        self._init_parameters()

    # ------------------------------------------------------------------------------------------------------
    # native plan + parameter storage
    # ------------------------------------------------------------------------------------------------------
    def _config_kwargs(self):
        return dict(min_max_distance=(self.min_distance, self.max_distance), network_width=self.network_width, maxl=self.max_sh,
                    num_cg_levels=self.num_cg_levels, num_channels_hidden=self.num_channels_hidden,
                    num_channels_per_element=self.num_channels_per_element, num_gaussians=self.num_gaussians,
                    bag_scale=self.bag_scale, beta=self.beta)

    def _check_supported(self):
        """The reference accepts any hyper-parameters (tools/arg_parser.py:55-61); the sm_100a kernels are instantiated for the
        ranges below (they cover BASELINE.json's five configurations and the reference's defaults)."""
        problems = []
        if self.max_sh != 4:
            problems.append(f'maxl={self.max_sh} (supported: 4)')
        if not 1 <= self.num_cg_levels <= 4:
            problems.append(f'num_cg_levels={self.num_cg_levels} (supported: 1..4)')
        if not 1 <= self.num_channels_hidden <= 10:
            problems.append(f'num_channels_hidden={self.num_channels_hidden} (supported: 1..10)')
        if not 1 <= self.num_channels_per_element <= 4:
            problems.append(f'num_channels_per_element={self.num_channels_per_element} (supported: 1..4)')
        if self.num_channels_out > 32:
            problems.append(f'len(zs) * num_channels_per_element={self.num_channels_out} (supported: <= 32)')
        if not 1 <= self.canvas_size <= 64:
            problems.append(f'canvas_size={self.canvas_size} (supported: 1..64)')
        if not 1 <= self.num_gaussians <= 8:
            problems.append(f'num_gaussians={self.num_gaussians} (supported: 1..8)')
        if not 1 <= self.network_width <= 1024:
            problems.append(f'network_width={self.network_width} (supported: 1..1024)')
        if problems:
            raise ValueError('molgym_b200.CovariantAC: unsupported hyper-parameters: ' + '; '.join(problems))

    def _init_native(self):
        lib = self._rt.lib()
        self._cfg = _cabi.make_config(self.zs, self.canvas_size, **self._config_kwargs())
        xyz, w = _lebedev_071()
        plan = ctypes.c_void_p()
        with self._rt.device_ctx():
            _cabi.check(lib, lib.mgb_cov_plan_create(ctypes.byref(self._cfg), xyz.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                                     w.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(w), ctypes.byref(plan)))
        self._plan = plan
        n = lib.mgb_cov_param_count(plan)
        off = (ctypes.c_int64 * n)()
        num = (ctypes.c_int64 * n)()
        tot = ctypes.c_int64()
        _cabi.check(lib, lib.mgb_cov_param_layout(plan, off, num, ctypes.byref(tot)))
        cats = (ctypes.c_int32 * (self.num_cg_levels * 10 + 5))()
        _cabi.check(lib, lib.mgb_cov_cat_sizes(plan, cats))
        self._p_offsets, self._p_numels, self._p_total = list(off), list(num), tot.value
        self._p_names = _cabi.param_names(self.num_cg_levels, self.max_sh)
        assert len(self._p_names) == n
        self._cat_sizes = list(cats)
        self._ws_cache: Dict[int, torch.Tensor] = {}
        self._fused_cache: Dict[tuple, _StepState] = {}
        self._eval_cache: Dict[int, list] = {}
        self._fused_streams = None
        self._fused_turn = 0
        self._fused_param_version = None
        self._fused_acc_event = None
        self._pending_infos = []
        self._one = torch.ones(1, dtype=torch.float32, device=self.device)

    def _param_shapes(self) -> Dict[str, tuple]:
        C, Z, cpe, W, G = self.num_channels_hidden, len(self.zs), self.num_channels_per_element, self.network_width, self.num_gaussians
        nl = self.max_sh + 1
        lat, late = (self.max_sh + 2) * Z * cpe * 2, (self.max_sh + 2) * cpe * 2
        shapes = {'cg_model.input_func_atom.lin.weight': (2 * C, 4 * Z), 'cg_model.input_func_atom.lin.bias': (2 * C, )}
        for k in range(self.num_cg_levels):
            rad = f'cg_model.rad_funcs.rad_funcs.{k}'
            shapes[f'{rad}.scales'] = (1, 1, 1, 8)
            shapes[f'{rad}.phases'] = (1, 1, 1, 8)
            cout = C if k < self.num_cg_levels - 1 else Z * cpe
            for l in range(nl):
                shapes[f'{rad}.linear.{l}.weight'] = (2 * C, 32)
                shapes[f'{rad}.linear.{l}.bias'] = (2 * C, )
                shapes[f'cg_model.cormorant_cg.edge_levels.{k}.cat_mix.mix_reps.weights.{l}'] = (C, self._cat_sizes[(k * nl + l) * 2], 2)
                shapes[f'cg_model.cormorant_cg.atom_levels.{k}.cat_mix.mix_reps.weights.{l}'] = (cout, self._cat_sizes[(k * nl + l) * 2 + 1], 2)
        for l in range(nl):
            shapes[f'cg_mix.cat_mix.mix_reps.weights.{l}'] = (cpe, self._cat_sizes[self.num_cg_levels * nl * 2 + l], 2)
        for head, (i, o) in dict(phi_focus=(lat, 1), phi_element=(lat, Z), phi_d=(late, 2 * G), phi_trans=(lat, W), phi_v=(W, 1)).items():
            shapes[f'{head}.layers.0.weight'] = (W, i)
            shapes[f'{head}.layers.0.bias'] = (W, )
            shapes[f'{head}.layers.1.weight'] = (o, W)
            shapes[f'{head}.layers.1.bias'] = (o, )
        shapes['distance_log_stds'] = (G, )
        return shapes

    def _initial_values(self) -> Dict[str, torch.Tensor]:
        """Same distributions as the reference's constructors, drawn in the reference's construction order
        (covariant/modules.py:59-95 -> cormorant RadialFilters / InputLinear / CormorantCG 'rand' weights with level_gain 10,
        modules.py:30-50 orthogonal MLPs, agent.py:131-133 log-stds)."""
        C, Z, cpe = self.num_channels_hidden, len(self.zs), self.num_channels_per_element
        shapes = self._param_shapes()
        vals: Dict[str, torch.Tensor] = {}
        nl = self.max_sh + 1

        def linear(name, out_f, in_f):
            lin = nn.Linear(in_f, out_f)
            vals[f'{name}.weight'], vals[f'{name}.bias'] = lin.weight.data, lin.bias.data

        def rand_mix(name, gain=10.0):
            shape = shapes[name]
            vals[name] = (gain / max(shape)) * (2 * torch.rand(shape) - 1)

        for k in range(self.num_cg_levels):
            rad = f'cg_model.rad_funcs.rad_funcs.{k}'
            scales = torch.cat([torch.arange(4), torch.arange(4)]).view(1, 1, 1, -1).to(torch.float)
            phases = torch.cat([torch.zeros(4), math.pi / 2 * torch.ones(4)]).view(1, 1, 1, -1)
            phases[0, 0, 0, 0] = math.pi / 2
            vals[f'{rad}.scales'], vals[f'{rad}.phases'] = scales, phases
            for l in range(nl):
                linear(f'{rad}.linear.{l}', 2 * C, 32)
        linear('cg_model.input_func_atom.lin', 2 * C, 4 * Z)
        for k in range(self.num_cg_levels):
            for l in range(nl):
                rand_mix(f'cg_model.cormorant_cg.edge_levels.{k}.cat_mix.mix_reps.weights.{l}', gain=1.0)
            for l in range(nl):
                rand_mix(f'cg_model.cormorant_cg.atom_levels.{k}.cat_mix.mix_reps.weights.{l}')
        for l in range(nl):
            rand_mix(f'cg_mix.cat_mix.mix_reps.weights.{l}')
        for head in ('phi_focus', 'phi_element', 'phi_d'):
            self._init_mlp(vals, shapes, head)
        vals['distance_log_stds'] = torch.log(torch.tensor([0.1] * self.num_gaussians, dtype=torch.float))
        for head in ('phi_trans', 'phi_v'):
            self._init_mlp(vals, shapes, head)
        return vals

    @staticmethod
    def _init_mlp(vals, shapes, head):
        for layer in (0, 1):
            out_f, in_f = shapes[f'{head}.layers.{layer}.weight']
            lin = nn.Linear(in_f, out_f)
            nn.init.orthogonal_(lin.weight.data)
            nn.init.constant_(lin.bias.data, 0)
            vals[f'{head}.layers.{layer}.weight'], vals[f'{head}.layers.{layer}.bias'] = lin.weight.data, lin.bias.data

    def _init_parameters(self):
        # registration order follows the reference's named_parameters(): log-stds first (agent.py:131), then the module tree
        order = ['distance_log_stds'] + [n for n in self._p_names if n != 'distance_log_stds']
        self._init_flat(self._param_shapes(), self._initial_values(), order)

    def __getstate__(self):
        state = self.__dict__.copy()
        for k in ('_rt', '_plan', '_cfg', '_ws_cache', '_fused_cache', '_eval_cache', '_fused_streams', '_fused_acc_event', '_one', '_pending_infos', '_info_stream',
                  '_flat', '_flat_grad', '_grad_local', '_grad_pending', '_views', '_grad_views', '_param_list'):
            state.pop(k, None)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self.__dict__.setdefault('graph_evaluate', True)
        self.__dict__.setdefault('fused_lazy_info', True)
        # torch.load(map_location=...) moves the unpickled parameters: follow them when they sit on a usable device
        # (tools/model_util.py:93-117 loads whole modules), else keep the pickled device, else the current one
        where = next(iter(torch.nn.Module.parameters(self))).device
        try:
            self._rt = self._runtime_cls(where)
        except Exception:
            self._rt = self._runtime_cls(None)
        self.device = self._rt.device
        self._init_native()
        self._rebuild_flat_after_unpickle()

    def __del__(self):
        try:
            self._rt.lib().mgb_cov_plan_destroy(self._plan)
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------------------
    # raw forward / backward through the C ABI
    # ------------------------------------------------------------------------------------------------------
    def _workspace(self, B: int, fresh: bool) -> torch.Tensor:
        nbytes = self._rt.lib().mgb_cov_workspace_bytes(self._plan, B)
        if fresh:
            return torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        ws = self._ws_cache.get(B)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._ws_cache = {B: ws}
        return ws

    def _out_layout(self, B: int):
        """Segments (in floats) of the packed output block: logp, ent, v, logp_parts, focus_probs, element_probs, gmm,
        coefficients, log_z."""
        N, Z, G, cpe = self.canvas_size, len(self.zs), self.num_gaussians, self.num_channels_per_element
        shapes = [(B, ), (B, ), (B, ), (B, 4), (B, N), (B, Z), (B, 3, G), (B, 25, cpe, 2), (B, )]
        offs, o = [], 0
        for sh in shapes:
            offs.append(o)
            o += (int(np.prod(sh)) + 3) // 4 * 4   # every segment starts on a 16-byte boundary (the kernels store float2 / float4)
        return shapes, offs, o

    def _slot_outputs(self, st, block: torch.Tensor):
        sizes = getattr(st, 'out_sizes', None)
        if sizes is None:
            sizes = st.out_sizes = [int(np.prod(sh)) for sh in st.out_shapes]
        return tuple(block[o:o + k].view(sh) for sh, o, k in zip(st.out_shapes, st.out_offs, sizes))

    def _cov_outputs(self, block: torch.Tensor, B: int, want_extras: bool):
        shapes, offs, _ = self._out_layout(B)
        o = _cabi.CovOutputs()
        base = block.data_ptr()
        names = ('logp', 'ent', 'v', 'logp_parts', 'focus_probs', 'element_probs', 'gmm', 'coefficients', 'log_z')
        for name, off in list(zip(names, offs))[:9 if want_extras else 3]:
            setattr(o, name, base + 4 * off)
        return o, shapes, offs

    def _forward_raw(self, pos, charges, bags, actions, want_extras=True, policy_only_ws=None):
        lib = self._rt.lib()
        if not self._params_aliased():
            self._realias()
        B = pos.shape[0]
        shapes, offs, total = self._out_layout(B)
        block = torch.empty(total, dtype=torch.float32, device=self.device)
        o, _, _ = self._cov_outputs(block, B, want_extras)
        outs = tuple(block[off:off + int(np.prod(sh))].view(sh) for sh, off in list(zip(shapes, offs))[:9 if want_extras else 3])
        stream = self._rt.stream_ptr()
        with self._rt.device_ctx():
            if policy_only_ws is not None:
                ws = policy_only_ws
                _cabi.check(lib, lib.mgb_cov_policy(self._plan, B, bags.data_ptr(), actions.data_ptr(), self._flat.data_ptr(),
                                                    ws.data_ptr(), ws.numel(), ctypes.byref(o), stream))
            else:
                ws = self._workspace(B, fresh=torch.is_grad_enabled())
                _cabi.check(lib, lib.mgb_cov_forward(self._plan, B, pos.data_ptr(), charges.data_ptr(), bags.data_ptr(),
                                                     actions.data_ptr(), self._flat.data_ptr(), ws.data_ptr(), ws.numel(),
                                                     ctypes.byref(o), stream))
        return outs, ws

    def _rollout_raw(self, pos, charges, bags, mode: int, seed: int):
        """Body + on-device sampling of the four sub-actions + evaluation of the chosen action (mgb_cov_rollout)."""
        lib = self._rt.lib()
        if not self._params_aliased():
            self._realias()
        B = pos.shape[0]
        shapes, offs, total = self._out_layout(B)
        block = torch.empty(total, dtype=torch.float32, device=self.device)
        o, _, _ = self._cov_outputs(block, B, True)
        outs = tuple(block[off:off + int(np.prod(sh))].view(sh) for sh, off in zip(shapes, offs))
        act = torch.empty(B, 6, dtype=torch.float32, device=self.device)
        ws = self._workspace(B, fresh=False)
        with self._rt.device_ctx():
            _cabi.check(lib, lib.mgb_cov_rollout(self._plan, B, pos.data_ptr(), charges.data_ptr(), bags.data_ptr(), self._flat.data_ptr(),
                                                 ws.data_ptr(), ws.numel(), mode, seed, act.data_ptr(), ctypes.byref(o),
                                                 self._rt.stream_ptr()))
        return act, outs

    def _backward_raw(self, pos, charges, bags, actions, ws, g_logp, g_ent, g_v):
        lib = self._rt.lib()
        B = pos.shape[0]
        stream = self._rt.stream_ptr()
        keep = self._attach_grads()
        target, accumulate = self._grad_target(keep)
        with self._rt.device_ctx():
            _cabi.check(lib, lib.mgb_cov_backward(self._plan, B, pos.data_ptr(), charges.data_ptr(), bags.data_ptr(),
                                                  actions.data_ptr(), self._flat.data_ptr(), ws.data_ptr(), ws.numel(),
                                                  g_logp.data_ptr(), g_ent.data_ptr(), g_v.data_ptr(), target.data_ptr(),
                                                  accumulate, stream))

    def _accumulate_scratch(self, scratch: torch.Tensor, scale: torch.Tensor, stream, keep=None):
        """.grad (or, data-parallel, the shard-local gradient) (+)= scale * scratch: one kernel on `stream`."""
        lib = self._rt.lib()
        if keep is None:
            keep = self._attach_grads()                  # may zero the flat gradient on the caller's stream
        target, accumulate = self._grad_target(keep)
        _cabi.check(lib, lib.mgb_scale_accumulate(target.data_ptr(), scratch.data_ptr(), scale.data_ptr(),
                                                  1 if scale.dtype == torch.float64 else 0, target.numel(), accumulate,
                                                  self._rt.stream_ptr(stream)))

    def _check_actions(self, actions_np: np.ndarray):
        """Discrete sub-actions index the canvas and the species list on the device; the reference raises from to_one_hot for
        indices outside the range (modules.py:8-23, pinned by tests/test_modules.py:22-29)."""
        if actions_np.shape[0] == 0:
            return
        focus, element = np.rint(actions_np[:, 0]), np.rint(actions_np[:, 1])
        if not (np.all((focus >= 0) & (focus < self.canvas_size)) and np.all((element >= 0) & (element < len(self.zs)))):
            raise RuntimeError(f'action index out of range: focus must be in [0, {self.canvas_size}), element in [0, {len(self.zs)})')

    def _shard(self, n: int) -> Tuple[int, int]:
        """The slice of a minibatch of n canvases this rank evaluates (everything when the agent is not data-parallel)."""
        if not self._is_sharded():
            return 0, n
        from molgym_b200.parallel import shard_bounds
        lo, hi = shard_bounds(n, torch.distributed.get_rank(), torch.distributed.get_world_size())
        if hi <= lo:
            raise RuntimeError(f'a minibatch of {n} canvases cannot be sharded over {torch.distributed.get_world_size()} ranks')
        return lo, hi

    # ------------------------------------------------------------------------------------------------------
    # persistent step state: pinned + device staging, workspace, outputs, CUDA graphs
    # ------------------------------------------------------------------------------------------------------
    def _new_step_state(self, B: int, with_targets: bool) -> _StepState:
        lib, rt = self._rt.lib(), self._rt
        dev, N, Z = self.device, self.canvas_size, len(self.zs)
        st = _StepState()
        sizes = [B * N * 3 * 4, B * N * 4, B * Z * 4, B * 6 * 4]           # pos, charges, bags, act
        if with_targets:
            sizes += [B * 4, B * 8, B * 8]                                   # old_logp, adv, ret
        offs = [0]
        for sz in sizes:
            offs.append((offs[-1] + sz + 255) // 256 * 256)
        st.host = rt.pinned(offs[-1])
        st.dev = torch.empty(offs[-1], dtype=torch.uint8, device=dev)
        h = st.host.numpy()
        seg = lambda buf, i: buf[offs[i]:offs[i] + sizes[i]]
        st.h_pos = seg(h, 0).view(np.float32).reshape(B, N, 3)
        st.h_charges = seg(h, 1).view(np.int32).reshape(B, N)
        st.h_bags = seg(h, 2).view(np.float32).reshape(B, Z)
        st.h_act = seg(h, 3).view(np.float32).reshape(B, 6)
        st.pos = seg(st.dev, 0).view(torch.float32).view(B, N, 3)
        st.charges = seg(st.dev, 1).view(torch.int32).view(B, N)
        st.bags = seg(st.dev, 2).view(torch.float32).view(B, Z)
        st.act = seg(st.dev, 3).view(torch.float32).view(B, 6)
        if with_targets:
            st.h_old = seg(h, 4).view(np.float32)
            st.h_adv = seg(h, 5).view(np.float64)
            st.h_ret = seg(h, 6).view(np.float64)
            st.old = seg(st.dev, 4).view(torch.float32)
            st.adv = seg(st.dev, 5).view(torch.float64)
            st.ret = seg(st.dev, 6).view(torch.float64)
        h[...] = 0
        st.dev.copy_(st.host)
        st.h2d_bytes = int(sum(sizes))
        st.grad = torch.zeros_like(self._flat_grad)
        st.ws = torch.empty(lib.mgb_cov_workspace_bytes(self._plan, B), dtype=torch.uint8, device=dev)
        st.B = B
        st.generation = 0
        return st

    def _fused_state(self, B: int, n_global: int, clip_ratio: float, vf_coef: float, entropy_coef: float, slot: int) -> _StepState:
        key = (B, n_global, float(clip_ratio), float(vf_coef), float(entropy_coef), self._flat.data_ptr(), slot)
        st = self._fused_cache.get(key)
        if st is not None:
            return st
        lib, rt = self._rt.lib(), self._rt
        dev = self.device
        st = self._new_step_state(B, with_targets=True)
        f32 = dict(dtype=torch.float32, device=dev)
        st.out = torch.empty(6, B, **f32)   # logp, ent, v, g_logp, g_ent, g_v
        st.info = torch.zeros(8, dtype=torch.float64, device=dev)
        st.info_ring = rt.pinned(64 * _INFO_RING).view(torch.float64).view(_INFO_RING, 8)   # pinned blocks the lazy infos read from
        st.ring_infos = [None] * _INFO_RING
        st.ring_pos = 0
        st.acc_event = rt.new_event()
        st.copied = rt.new_event()
        st.copied.record(rt.current_stream())
        st.stream = self._fused_streams[slot]
        o = _cabi.CovOutputs()
        o.logp, o.ent, o.v = st.out[0].data_ptr(), st.out[1].data_ptr(), st.out[2].data_ptr()
        st.outputs = o
        plan, flat = self._plan, self._flat
        inv_global = 1.0 / n_global   # the loss is the mean over the GLOBAL minibatch (ppo.py:36-52); a shard adds its part

        def forward_and_loss(stream):
            _cabi.check(lib, lib.mgb_cov_forward(plan, B, st.pos.data_ptr(), st.charges.data_ptr(), st.bags.data_ptr(), st.act.data_ptr(),
                                                 flat.data_ptr(), st.ws.data_ptr(), st.ws.numel(), ctypes.byref(o), stream))
            _cabi.check(lib, lib.mgb_ppo_loss(B, st.out[0].data_ptr(), st.out[1].data_ptr(), st.out[2].data_ptr(), st.old.data_ptr(),
                                              st.adv.data_ptr(), st.ret.data_ptr(), clip_ratio, vf_coef, entropy_coef, inv_global,
                                              st.info.data_ptr(), st.out[3].data_ptr(), st.out[4].data_ptr(), st.out[5].data_ptr(), stream))

        def backward(stream):
            _cabi.check(lib, lib.mgb_cov_backward(plan, B, st.pos.data_ptr(), st.charges.data_ptr(), st.bags.data_ptr(), st.act.data_ptr(),
                                                  flat.data_ptr(), st.ws.data_ptr(), st.ws.numel(), st.out[3].data_ptr(),
                                                  st.out[4].data_ptr(), st.out[5].data_ptr(), st.grad.data_ptr(), 0, stream))

        with rt.device_ctx():
            st.g_forward = rt.capture(forward_and_loss)
            st.g_backward = rt.capture(backward)
        if len(self._fused_cache) >= 8:   # (minibatch size + remainder size) x two pipeline slots (+ a change of coefficients)
            self._fused_cache.pop(next(iter(self._fused_cache)))
        self._fused_cache[key] = st
        return st

    def _eval_slot(self, B: int) -> Optional[_StepState]:
        """A free evaluation slot for minibatches of B canvases, or None when both are still waiting for their backward."""
        slots = self._eval_cache.setdefault(B, [])
        for st in slots:
            if st.flat_ptr != self._flat.data_ptr():
                continue
            if st.done or st.live is None or st.live() is None:
                return st
        if len([s for s in slots if s.flat_ptr == self._flat.data_ptr()]) >= 2:
            return None
        if sum(len(v) for v in self._eval_cache.values()) >= 6:
            self._eval_cache.pop(next(iter(self._eval_cache)))
            slots = self._eval_cache.setdefault(B, [])
        lib, rt = self._rt.lib(), self._rt
        st = self._new_step_state(B, with_targets=False)
        st.flat_ptr = self._flat.data_ptr()
        st.out_shapes, st.out_offs, total = self._out_layout(B)
        st.out_all = torch.empty(total, dtype=torch.float32, device=self.device)
        st.cot = torch.zeros(3, B, dtype=torch.float32, device=self.device)
        o, _, _ = self._cov_outputs(st.out_all, B, True)
        st.outputs = o
        st.live, st.done = None, True
        st.copied = rt.new_event()
        st.copied.record(rt.current_stream())
        plan, flat = self._plan, self._flat

        def forward(stream):
            _cabi.check(lib, lib.mgb_cov_forward(plan, B, st.pos.data_ptr(), st.charges.data_ptr(), st.bags.data_ptr(), st.act.data_ptr(),
                                                 flat.data_ptr(), st.ws.data_ptr(), st.ws.numel(), ctypes.byref(o), stream))

        def backward(stream):
            _cabi.check(lib, lib.mgb_cov_backward(plan, B, st.pos.data_ptr(), st.charges.data_ptr(), st.bags.data_ptr(), st.act.data_ptr(),
                                                  flat.data_ptr(), st.ws.data_ptr(), st.ws.numel(), st.cot[0].data_ptr(),
                                                  st.cot[1].data_ptr(), st.cot[2].data_ptr(), st.grad.data_ptr(), 0, stream))

        with rt.device_ctx():
            st.g_forward = rt.capture(forward)
            st.g_backward = rt.capture(backward)
        slots.append(st)
        return st

    # ------------------------------------------------------------------------------------------------------
    # fused PPO minibatch step: pack -> one H2D copy -> CUDA-graph replay of forward + PPO-clip loss, then of the backward
    # ------------------------------------------------------------------------------------------------------
    def train_shard(self, n: int):
        """(lo, hi) of a minibatch of n canvases when the caller may hand over just this rank's slice (molgym_b200.ppo.train does:
        on a data-parallel agent with the fused step it collects data[lo:hi] only and passes n_global = n), else None."""
        if not (self._is_sharded() and self.fused_ppo):
            return None
        return self._shard(n)

    def fused_ppo_loss(self, observations: List, actions, old_logp, adv, ret, clip_ratio: float, vf_coef: float,
                       entropy_coef: float, n_global: Optional[int] = None):
        """The arithmetic of ppo.compute_loss (ppo.py:18-63) on this agent, as one pinned staging copy and two CUDA-graph
        replays: forward + PPO-clip loss (float64, k_ppo_loss), then the backward into a scratch gradient, enqueued right away
        (ppo.train always differentiates the loss it just computed, ppo.py:126-131).  Returns (loss, info) like compute_loss;
        loss.backward() scales the scratch gradient into the parameters' .grad.

        Consecutive calls alternate between two pipeline slots (own stream, staging buffers, workspace, graphs): within a PPO
        epoch the minibatches are independent — the parameters only move at optimizer.step() (ppo.py:122-146) — so the forward
        of minibatch i+1 runs beside the backward of minibatch i.  The slot's stream waits for the caller's stream only when the
        parameters changed since the last fused call (in-place version counters; `fused_sync_params = True` forces the wait
        every time, for code that edits parameters through `.data`); the caller's stream waits for the accumulated gradient in
        loss.backward().

        Data-parallel (parallel.shard_agent): every rank is handed the SAME minibatch and evaluates its contiguous shard of
        it; the loss terms are sums over the shard divided by the global minibatch size.  The returned info is global on every
        rank (identical bits: every rank takes the same early-stop branch, ppo.py:138-140): with `fused_lazy_info` the pending
        8-double blocks of an epoch are summed over ranks in ONE all-reduce when the info is first read, else each block is
        all-reduced on the stream right behind its forward pass.  The returned loss TENSOR carries the exact gradient; in the
        lazy data-parallel mode its value is this rank's share of the global loss (the shares sum to info['total_loss']).
        The shard's gradient is accumulated locally and reduced once per optimizer step (FlatParamMixin.sync_grads)."""
        n = len(observations)
        rt = self._rt
        if not self._params_aliased():
            self._realias()
        actions_np = _as_numpy_actions(actions, n, 6)
        self._check_actions(actions_np)
        sharded = self._is_sharded()
        if n_global is not None and sharded:
            lo, hi = 0, n           # the caller collected this rank's slice of a minibatch of n_global canvases (train_shard)
            n = int(n_global)
        else:
            lo, hi = self._shard(n)
        B = hi - lo
        presliced = n_global is not None and sharded
        if self._fused_streams is None:
            self._fused_streams = [rt.new_stream(), rt.new_stream()]
        slot = self._fused_turn
        # two slots double the workspace: keep one when that is too much for the device
        if 2 * rt.lib().mgb_cov_workspace_bytes(self._plan, B) > 0.25 * rt.total_memory():
            slot = 0
        else:
            self._fused_turn ^= 1
        st = self._fused_state(B, n, clip_ratio, vf_coef, entropy_coef, slot)
        st.generation += 1
        k = st.ring_pos
        st.ring_pos = (k + 1) % _INFO_RING
        if st.ring_infos[k] is not None:
            st.ring_infos[k]._resolve()   # a block is read out before it is overwritten (sixteen steps later: a no-op in practice)
        st.copied.synchronize()        # ... and its previous staging copy has left the pinned buffer (two steps ago: no wait in practice)
        pack_observations(observations[lo:hi] if (sharded and not presliced) else observations, self.zs, self.canvas_size, cfg=self._cfg,
                          out=(st.h_pos, st.h_charges, st.h_bags), lib=rt.lib())
        st.h_act[...] = actions_np[lo:hi]
        st.h_old[...] = old_logp[lo:hi]
        st.h_adv[...] = adv[lo:hi]
        st.h_ret[...] = ret[lo:hi]
        with rt.device_ctx():
            current = rt.current_stream()
            version = self._param_version()
            if self.fused_sync_params or version != self._fused_param_version:
                for stream in self._fused_streams:   # parameter updates enqueued on the caller's stream come first
                    stream.wait_stream(current)
                self._fused_param_version = version
            with rt.stream_ctx(st.stream):
                st.dev.copy_(st.host, non_blocking=True)
                st.copied.record(st.stream)
                st.g_forward.replay()
                if sharded and not self.fused_lazy_info:
                    torch.distributed.all_reduce(st.info, op=torch.distributed.ReduceOp.SUM)
                st.info_ring[k].copy_(st.info, non_blocking=True)
                loss_dev = st.info[0].clone()
                event = rt.new_event()       # per call: the lazy info may be read after the slot has moved on
                event.record(st.stream)
                st.g_backward.replay()
            current.wait_event(event)        # loss_dev is consumed on the caller's stream (device-side wait: the host goes on)
            if rt.is_cuda:
                loss_dev.record_stream(current)
        info = LazyLossInfo(self, st.info_ring[k], event, sharded and self.fused_lazy_info)
        st.ring_infos[k] = info
        if not self.fused_lazy_info:
            info._resolve()
        loss = _FusedPPOLoss.apply(self._param_list[-1], self, st, loss_dev)
        return loss, info

    def _fused_backward(self, st: _StepState, g: torch.Tensor):
        rt = self._rt
        with rt.device_ctx():
            current = rt.current_stream()
            scale = g.detach()
            if scale.dtype not in (torch.float32, torch.float64) or scale.device != self._flat.device:
                scale = scale.to(device=self.device, dtype=torch.float32)
            keep = self._attach_grads()                      # may zero the flat gradient on the caller's stream ...
            st.stream.wait_stream(current)                   # ... and the cotangent is produced there
            if self._fused_acc_event is not None:
                st.stream.wait_event(self._fused_acc_event)  # accumulations of the two slots into .grad stay ordered
            # .grad (+)= cotangent * scratch gradient: one kernel on the slot's stream, behind the backward graph
            self._accumulate_scratch(st.grad, scale, st.stream, keep=keep)
            if rt.is_cuda:
                scale.record_stream(st.stream)
            st.acc_event.record(st.stream)
            self._fused_acc_event = st.acc_event
            current.wait_event(st.acc_event)                 # whoever reads .grad next on the caller's stream sees it complete

    # ------------------------------------------------------------------------------------------------------
    # the reference surface
    # ------------------------------------------------------------------------------------------------------
    def to_action_space(self, action, observation):
        """agent.py:147-163."""
        action = np.asarray(action.detach().cpu().numpy() if torch.is_tensor(action) else action)
        assert action.shape == (6, )
        focus = int(round(action[0].item()))
        element_index = int(round(action[1].item()))
        d, so3 = action[2], action[-3:]
        null = self.zs.index(0)
        atoms = [item for item in observation[0] if item[0] != null]
        if len(atoms):
            position = tuple(np.asarray(atoms[focus][1], dtype=np.float64) + d * so3)
        else:
            position = (0.0, 0.0, 0.0)
        return element_index, position

    def parse_observations(self, observations: List, actions: Optional[np.ndarray] = None) -> Dict[str, torch.Tensor]:
        """agent.py:165-197 — device tensors for a list of observations.  Everything (positions, charges, bags and, when
        given, the actions) is packed into ONE pinned staging buffer and crosses PCIe as one non-blocking copy."""
        B, N, Z = len(observations), self.canvas_size, len(self.zs)
        sizes = [B * N * 3 * 4, B * N * 4, B * Z * 4, B * 6 * 4 if actions is not None else 0]
        offs = [0]
        for sz in sizes:
            offs.append((offs[-1] + sz + 255) // 256 * 256)
        stage = self._rt.pinned(max(offs[-1], 256))
        host = stage.numpy()
        pos = host[offs[0]:offs[0] + sizes[0]].view(np.float32).reshape(B, N, 3)
        charges = host[offs[1]:offs[1] + sizes[1]].view(np.int32).reshape(B, N)
        bags = host[offs[2]:offs[2] + sizes[2]].view(np.float32).reshape(B, Z)
        pack_observations(observations, self.zs, N, cfg=self._cfg, out=(pos, charges, bags), lib=self._rt.lib())
        if actions is not None:
            host[offs[3]:offs[3] + sizes[3]].view(np.float32).reshape(B, 6)[...] = actions
        dev = stage.to(self.device, non_blocking=True)
        data = dict(positions=dev[offs[0]:offs[0] + sizes[0]].view(torch.float32).view(B, N, 3),
                    charges=dev[offs[1]:offs[1] + sizes[1]].view(torch.int32).view(B, N),
                    bags=dev[offs[2]:offs[2] + sizes[2]].view(torch.float32).view(B, Z))
        if actions is not None:
            data['actions'] = dev[offs[3]:offs[3] + sizes[3]].view(torch.float32).view(B, 6)
        return data

    def step(self, observations: List, actions: Optional[np.ndarray] = None) -> dict:
        """agent.py:209-334.  Data-parallel agents evaluate their shard of the minibatch and all-gather logp / ent / v (the
        gather routes the cotangents back to the owning rank), so callers see the global vectors; `dists` then describes the
        local shard only."""
        response: Dict[str, Any] = {}
        if actions is not None:
            n = len(observations)
            actions_np = _as_numpy_actions(actions, n, 6)
            self._check_actions(actions_np)
            lo, hi = self._shard(n)
            sharded = self._is_sharded()
            if sharded:
                observations, actions_np = observations[lo:hi], actions_np[lo:hi]
            st = self._eval_slot(hi - lo) if (self.graph_evaluate and torch.is_grad_enabled()) else None
            if st is not None:
                outs, act, charges = self._evaluate_slot(st, observations, actions_np)
            else:
                data = self.parse_observations(observations, actions_np)
                pos, charges, bags, act = data['positions'], data['charges'], data['bags'], data['actions']
                outs = self._evaluate(pos, charges, bags, act)
            if sharded:
                from molgym_b200.parallel import gather_shards
                outs = tuple(gather_shards(o, n) for o in outs[:3]) + tuple(outs[3:])
                act = torch.as_tensor(_as_numpy_actions(actions, n, 6), device=self.device)
        else:
            data = self.parse_observations(observations)
            pos, charges, bags = data['positions'], data['charges'], data['bags']
            act, outs = sampling.rollout(self, pos, charges, bags, training=self.training)
            response['actions'] = [self.to_action_space(a, o) for a, o in zip(act.detach().cpu().numpy(), observations)]
        logp, ent, v, parts, fprobs, eprobs, gmm, coeff, log_z = outs
        response.update({
            'a': act, 'logp': logp, 'ent': ent, 'v': v,
            'dists': sampling.LazyDists(self, fprobs, eprobs, gmm, coeff, log_z, charges),
        })
        return response

    def _evaluate(self, pos, charges, bags, act):
        if torch.is_grad_enabled():
            return _CovEvaluate.apply(self._param_list[-1], self, pos, charges, bags, act)
        outs, _ = self._forward_raw(pos, charges, bags, act, want_extras=True)
        return outs

    def _evaluate_slot(self, st: _StepState, observations: List, actions_np: np.ndarray):
        """Evaluate-mode step on a persistent slot: pack into the pinned staging buffer, one H2D copy, graph replay."""
        if not self._params_aliased():
            self._realias()
        st.generation += 1
        st.copied.synchronize()   # the previous staging copy of this slot has left the pinned buffer
        pack_observations(observations, self.zs, self.canvas_size, cfg=self._cfg, out=(st.h_pos, st.h_charges, st.h_bags),
                          lib=self._rt.lib())
        st.h_act[...] = actions_np
        with self._rt.device_ctx():
            st.dev.copy_(st.host, non_blocking=True)
            st.copied.record(self._rt.current_stream())
            outs = _CovEvaluateSlot.apply(self._param_list[-1], self, st)
        # the staging buffers are rewritten by the next step on this slot: hand out copies of the small inputs the response keeps
        return outs, st.act.clone(), st.charges.clone()
