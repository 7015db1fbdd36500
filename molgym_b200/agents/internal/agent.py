"""SchNetAC — the reference's internal-coordinate actor-critic (molgym/agents/internal/agent.py:17-353) behind the same
constructor and `step(observations, actions)` contract, computed by the sm_100a kernels of molgym_b200/csrc/internal.cuh
through the C ABI (mgb_int_forward / mgb_int_backward).

The reference evaluates SchNet three times per observation at batch size one inside Python loops (agent.py:124-128,
163-177); here all canvases and both hypothetical-atom variants go through the device at once.  The z-matrix placement of
the hypothetical atom (zmat.py:99-133) is float64 host arithmetic in the reference and stays on the host.
"""
import ctypes
import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from molgym_b200 import _cabi
from molgym_b200.agents._flat import FlatParamMixin
from molgym_b200.agents._runtime import CudaRuntime
from molgym_b200.agents.base import AbstractActorCritic
from molgym_b200.agents.internal import zmat


class _IntEvaluate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, agent, numbers, positions, bags, actions):
        outs, ws = agent._forward_raw(numbers, positions, bags, actions)
        ctx.agent = agent
        ctx.saved = (numbers, positions, bags, actions, ws)
        ctx.mark_non_differentiable(*outs[3:])
        return outs

    @staticmethod
    def backward(ctx, g_logp, g_ent, g_v, *unused):
        agent = ctx.agent
        numbers, positions, bags, actions, ws = ctx.saved
        B = numbers.shape[0]

        def prep(g):
            if g is None:
                return torch.zeros(B, dtype=torch.float32, device=numbers.device)
            return g.to(torch.float32).contiguous()

        agent._backward_raw(numbers, positions, bags, actions, ws, prep(g_logp), prep(g_ent), prep(g_v))
        return (None, ) * 6


class SchNetAC(FlatParamMixin, AbstractActorCritic):
    _runtime_cls = CudaRuntime   # the test-suite's emulator-backed subclass substitutes a host runtime

    def __init__(self, observation_space, action_space, min_max_distance: Tuple[float, float], network_width: int, device=None):
        super().__init__(observation_space=observation_space, action_space=action_space)
        self._rt = self._runtime_cls(device)
        self.device = self._rt.device
        self.zs = list(self.observation_space.zs)
        self.num_atoms = self.observation_space.canvas_space.size
        self.num_zs = len(self.zs)
        self.network_width = network_width
        self.num_afeats = network_width // 2
        self.num_latent_beta = network_width // 4
        self.num_latent = self.num_afeats + self.num_latent_beta
        self.min_distance, self.max_distance = min_max_distance
        self.data_parallel = False
        self._init_native()
        order = ['log_stds'] + [n for n in self._p_names if n != 'log_stds']   # agent.py:66 registers log_stds after the MLPs;
        self._init_flat(self._param_shapes(), self._initial_values(), order)   # order only affects iteration, not names

    # ------------------------------------------------------------------------------------------------------
    def _init_native(self):
        lib = self._rt.lib()
        self._cfg = _cabi.make_int_config(self.zs, self.num_atoms, (self.min_distance, self.max_distance), self.network_width)
        plan = ctypes.c_void_p()
        with self._rt.device_ctx():
            _cabi.check(lib, lib.mgb_int_plan_create(ctypes.byref(self._cfg), ctypes.byref(plan)))
        self._plan = plan
        n = lib.mgb_int_param_count(plan)
        off = (ctypes.c_int64 * n)()
        num = (ctypes.c_int64 * n)()
        tot = ctypes.c_int64()
        _cabi.check(lib, lib.mgb_int_param_layout(plan, off, num, ctypes.byref(tot)))
        self._p_offsets, self._p_numels, self._p_total = list(off), list(num), tot.value
        self._p_names = _cabi.int_param_names()
        assert len(self._p_names) == n

    def _param_shapes(self) -> Dict[str, tuple]:
        F, W, Z, LB, lat = self.num_afeats, self.network_width, self.num_zs, self.num_latent_beta, self.num_latent
        shapes = {'embedding_fn.embedding.weight': (100, F)}
        for t in range(3):
            it = f'embedding_fn.interactions.{t}'
            shapes.update({f'{it}.filter_network.0.weight': (128, 25), f'{it}.filter_network.0.bias': (128, ),
                           f'{it}.filter_network.1.weight': (128, 128), f'{it}.filter_network.1.bias': (128, ),
                           f'{it}.cfconv.in2f.weight': (128, F), f'{it}.cfconv.f2out.weight': (F, 128), f'{it}.cfconv.f2out.bias': (F, ),
                           f'{it}.dense.weight': (F, F), f'{it}.dense.bias': (F, )})
        for head, (i, o) in dict(phi_beta=(Z, LB), phi_focus=(lat, 1), phi_element=(lat, Z), phi_continuous=(lat + Z, 3),
                                 phi_kappa=(lat, 1)).items():
            shapes.update({f'{head}.layers.0.weight': (W, i), f'{head}.layers.0.bias': (W, ), f'{head}.layers.1.weight': (o, W),
                           f'{head}.layers.1.bias': (o, )})
        shapes.update({'critic.layers.0.weight': (W, lat), 'critic.layers.0.bias': (W, ), 'critic.layers.1.weight': (W, W),
                       'critic.layers.1.bias': (W, ), 'critic.layers.2.weight': (1, W), 'critic.layers.2.bias': (1, ),
                       'log_stds': (3, )})
        return shapes

    def _initial_values(self) -> Dict[str, torch.Tensor]:
        """schnetpack 0.3 initialisation (Embedding N(0,1) with padding row 0 zeroed, Dense xavier-uniform weights / zero
        biases) and molgym's orthogonal MLPs (modules.py:30-50), log-stds as agent.py:66."""
        shapes = self._param_shapes()
        vals: Dict[str, torch.Tensor] = {}
        emb = nn.Embedding(100, self.num_afeats, padding_idx=0)
        vals['embedding_fn.embedding.weight'] = emb.weight.data
        for name, shape in shapes.items():
            if not name.startswith('embedding_fn.interactions'):
                continue
            if name.endswith('bias'):
                vals[name] = torch.zeros(shape)
            else:
                w = torch.empty(shape)
                nn.init.xavier_uniform_(w)
                vals[name] = w
        for head in ('phi_beta', 'phi_focus', 'phi_element', 'phi_continuous', 'phi_kappa', 'critic'):
            layer = 0
            while f'{head}.layers.{layer}.weight' in shapes:
                out_f, in_f = shapes[f'{head}.layers.{layer}.weight']
                lin = nn.Linear(in_f, out_f)
                nn.init.orthogonal_(lin.weight.data)
                nn.init.constant_(lin.bias.data, 0)
                vals[f'{head}.layers.{layer}.weight'], vals[f'{head}.layers.{layer}.bias'] = lin.weight.data, lin.bias.data
                layer += 1
        vals['log_stds'] = torch.log(torch.tensor([0.15, 0.25, 0.25], dtype=torch.float32))
        return vals

    def load_state_dict(self, state_dict, strict: bool = True):
        """The reference registers the filter network twice (SchNetInteraction.filter_network and .cfconv.filter_network
        share their parameters), so its state_dict carries every filter tensor under two names."""
        state_dict = {k: v for k, v in state_dict.items() if '.cfconv.filter_network.' not in k}
        return super().load_state_dict(state_dict, strict=strict)

    def __getstate__(self):
        state = self.__dict__.copy()
        for k in ('_rt', '_plan', '_cfg', '_flat', '_flat_grad', '_grad_local', '_grad_pending', '_views', '_grad_views', '_param_list'):
            state.pop(k, None)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        where = next(iter(torch.nn.Module.parameters(self))).device   # torch.load(map_location=...) moved them
        try:
            self._rt = self._runtime_cls(where)
        except Exception:
            self._rt = self._runtime_cls(None)
        self.device = self._rt.device
        self._init_native()
        self._rebuild_flat_after_unpickle()

    def __del__(self):
        try:
            self._rt.lib().mgb_int_plan_destroy(self._plan)
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------------------
    def _forward_raw(self, numbers, positions, bags, actions):
        lib = self._rt.lib()
        if not self._params_aliased():
            self._realias()
        B, dev = numbers.shape[0], self.device
        f32 = dict(dtype=torch.float32, device=dev)
        logp, ent, v = torch.empty(B, **f32), torch.empty(B, **f32), torch.empty(B, **f32)
        terms, fprobs, eprobs = torch.empty(B, 6, **f32), torch.empty(B, self.num_atoms, **f32), torch.empty(B, self.num_zs, **f32)
        means, klog = torch.empty(B, 3, **f32), torch.empty(B, 2, **f32)
        o = _cabi.IntOutputs(logp=logp.data_ptr(), ent=ent.data_ptr(), v=v.data_ptr(), logp_terms=terms.data_ptr(),
                             focus_probs=fprobs.data_ptr(), element_probs=eprobs.data_ptr(), means=means.data_ptr(),
                             kappa_logits=klog.data_ptr())
        ws = torch.empty(lib.mgb_int_workspace_bytes(self._plan, B), dtype=torch.uint8, device=dev)
        stream = self._rt.stream_ptr()
        with self._rt.device_ctx():
            _cabi.check(lib, lib.mgb_int_forward(self._plan, B, numbers.data_ptr(), positions.data_ptr(), bags.data_ptr(),
                                                 actions.data_ptr(), self._flat.data_ptr(), ws.data_ptr(), ws.numel(), ctypes.byref(o),
                                                 stream))
        return (logp, ent, v, terms, fprobs, eprobs, means, klog), ws

    def _backward_raw(self, numbers, positions, bags, actions, ws, g_logp, g_ent, g_v):
        lib = self._rt.lib()
        B = numbers.shape[0]
        stream = self._rt.stream_ptr()
        keep = self._attach_grads()
        target, accumulate = self._grad_target(keep)
        with self._rt.device_ctx():
            _cabi.check(lib, lib.mgb_int_backward(self._plan, B, numbers.data_ptr(), positions.data_ptr(), bags.data_ptr(),
                                                  actions.data_ptr(), self._flat.data_ptr(), ws.data_ptr(), ws.numel(),
                                                  g_logp.data_ptr(), g_ent.data_ptr(), g_v.data_ptr(), target.data_ptr(), accumulate,
                                                  stream))

    def _check_actions(self, actions_np: np.ndarray):
        """focus indexes the canvas, element the species list (zmat.build_molecules, the head kernels); the reference raises
        from to_one_hot / zmat for indices outside the range (modules.py:8-23, zmat.py:106-107)."""
        if actions_np.shape[0] == 0:
            return
        focus, element = np.rint(actions_np[:, 1]), np.rint(actions_np[:, 2])
        if not (np.all((focus >= 0) & (focus < self.num_atoms)) and np.all((element >= 0) & (element < self.num_zs))):
            raise RuntimeError(f'action index out of range: focus must be in [0, {self.num_atoms}), element in [0, {self.num_zs})')

    def _device_inputs(self, observations, actions_np):
        self._check_actions(actions_np)
        numbers, positions, bags = zmat.build_molecules(observations, actions_np, self.zs, self.num_atoms)
        dev = self.device
        return (torch.from_numpy(numbers).to(dev), torch.from_numpy(positions).to(dev), torch.from_numpy(bags).to(dev),
                torch.as_tensor(actions_np, dtype=torch.float32, device=dev).contiguous())

    def _evaluate(self, observations, actions_np):
        numbers, positions, bags, act = self._device_inputs(observations, actions_np)
        if torch.is_grad_enabled():
            return act, _IntEvaluate.apply(self._param_list[-1], self, numbers, positions, bags, act)
        outs, _ = self._forward_raw(numbers, positions, bags, act)
        return act, outs

    # ------------------------------------------------------------------------------------------------------
    def to_action_space(self, action, observation):
        """agent.py:91-110."""
        action = np.asarray(action.detach().cpu().numpy() if torch.is_tensor(action) else action)
        stop, focus, element, distance, angle, dihedral, kappa = action
        if stop:
            return None
        focus, element = int(round(focus)), int(round(element))
        sign = -1 if int(round(kappa)) else 1
        null = self.zs.index(0)
        positions = [np.asarray(xyz, dtype=np.float64) for label, xyz in observation[0] if label != null]
        position = zmat.position_atom_helper(positions, focus=focus, distance=distance, angle=angle, dihedral=sign * dihedral)
        atomic_number_index = list(self.action_space.zs).index(self.zs[element])
        return atomic_number_index, tuple(position)

    @torch.no_grad()
    def _rollout(self, observations):
        """agent.py:212-306 with actions=None: focus / element / distance / angle / dihedral do not depend on the hypothetical
        atom, so they are drawn from a first evaluation; kappa needs the placed atom and comes from a second one."""
        B = len(observations)
        act = np.zeros((B, 7), dtype=np.float32)
        act[:, 3] = 0.5 * (self.min_distance + self.max_distance)
        act[:, 4] = act[:, 5] = 0.5 * math.pi
        for b, (_, bag) in enumerate(observations):
            act[b, 2] = int(np.argmax(np.asarray(bag) > 0))
        _, outs = self._evaluate(observations, act)
        fprobs = outs[4]
        focus = torch.distributions.Categorical(probs=fprobs).sample() if self.training else torch.argmax(fprobs, dim=-1)
        act[:, 1] = focus.cpu().numpy()
        _, outs = self._evaluate(observations, act)
        eprobs = outs[5]
        element = torch.distributions.Categorical(probs=eprobs).sample() if self.training else torch.argmax(eprobs, dim=-1)
        act[:, 2] = element.cpu().numpy()
        _, outs = self._evaluate(observations, act)
        means = outs[6]
        if self.training:
            stds = torch.exp(1e-6 + self._views['log_stds'].detach())
            cont = torch.normal(means, stds.expand_as(means))
            cont[:, 0] = cont[:, 0].clamp(0.001)
        else:
            cont = means
        act[:, 3:6] = cont.cpu().numpy()
        _, outs = self._evaluate(observations, act)
        klog = outs[7]
        kappa = torch.distributions.Categorical(logits=klog).sample() if self.training else torch.argmax(klog, dim=-1)
        act[:, 6] = kappa.cpu().numpy()
        return self._evaluate(observations, act)

    def step(self, observations: List, actions: Optional[np.ndarray] = None) -> dict:
        if actions is not None:
            actions_np = np.asarray(actions.detach().cpu().numpy() if torch.is_tensor(actions) else actions, dtype=np.float32)
            assert actions_np.shape == (len(observations), 7)
            act, outs = self._evaluate(observations, actions_np)
        else:
            act, outs = self._rollout(observations)
        logp, ent, v = outs[0], outs[1], outs[2]
        return {
            'a': act, 'logp': logp, 'ent': ent, 'v': v,
            'actions': [self.to_action_space(a, o) for a, o in zip(act.detach().cpu().numpy(), observations)],
        }
