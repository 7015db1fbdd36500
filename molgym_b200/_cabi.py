"""ctypes declarations for the C ABI in include/molgym_b200.h (no torch types cross this boundary)."""
import ctypes
from ctypes import POINTER, c_char_p, c_double, c_float, c_int32, c_int64, c_size_t, c_void_p
from typing import List, Sequence, Tuple

MGB_MAX_SPECIES = 16
MAXL = 4


class CovConfig(ctypes.Structure):
    _fields_ = [
        ('canvas_size', c_int32),
        ('num_species', c_int32),
        ('zs', c_int32 * MGB_MAX_SPECIES),
        ('maxl', c_int32),
        ('num_cg_levels', c_int32),
        ('num_channels_hidden', c_int32),
        ('num_channels_per_element', c_int32),
        ('num_gaussians', c_int32),
        ('network_width', c_int32),
        ('min_distance', c_float),
        ('max_distance', c_float),
        ('bag_scale', c_float),
        ('has_beta', c_int32),
        ('beta', c_float),
        ('rel_sh_normalize', c_int32),
    ]


class CovOutputs(ctypes.Structure):
    _fields_ = [(name, c_void_p) for name in ('logp', 'ent', 'v', 'logp_parts', 'focus_probs', 'element_probs', 'gmm',
                                              'coefficients', 'log_z', 'covariats')]


class IntConfig(ctypes.Structure):
    _fields_ = [('canvas_size', c_int32), ('num_species', c_int32), ('zs', c_int32 * MGB_MAX_SPECIES), ('network_width', c_int32),
                ('min_distance', c_float), ('max_distance', c_float)]


class IntOutputs(ctypes.Structure):
    _fields_ = [(name, c_void_p) for name in ('logp', 'ent', 'v', 'logp_terms', 'focus_probs', 'element_probs', 'means', 'kappa_logits')]


EXPORTS = ('mgb_last_error', 'mgb_version', 'mgb_is_cuda_build', 'mgb_clebsch_gordan', 'mgb_cov_plan_create', 'mgb_cov_plan_destroy',
           'mgb_cov_param_count', 'mgb_cov_param_layout', 'mgb_cov_cat_sizes', 'mgb_cov_workspace_bytes',
           'mgb_cov_forward', 'mgb_cov_rollout', 'mgb_cov_policy', 'mgb_cov_backward', 'mgb_ppo_loss', 'mgb_pack_observations', 'mgb_scale_accumulate', 'mgb_optim_scratch_bytes', 'mgb_grad_norm', 'mgb_adam_step', 'mgb_launch_count',
           'mgb_profile_kernel', 'mgb_profile_read', 'mgb_profile_report', 'mgb_int_plan_create', 'mgb_int_plan_destroy', 'mgb_int_param_count',
           'mgb_int_param_layout', 'mgb_int_workspace_bytes', 'mgb_int_forward', 'mgb_int_backward')


def bind(lib: ctypes.CDLL) -> ctypes.CDLL:
    lib.mgb_last_error.restype = c_char_p
    lib.mgb_last_error.argtypes = []
    lib.mgb_version.restype = ctypes.c_int
    lib.mgb_is_cuda_build.restype = ctypes.c_int
    lib.mgb_clebsch_gordan.restype = c_double
    lib.mgb_clebsch_gordan.argtypes = [c_int32] * 6
    lib.mgb_cov_plan_create.restype = ctypes.c_int
    lib.mgb_cov_plan_create.argtypes = [POINTER(CovConfig), POINTER(c_double), POINTER(c_double), c_int32, POINTER(c_void_p)]
    lib.mgb_cov_plan_destroy.restype = None
    lib.mgb_cov_plan_destroy.argtypes = [c_void_p]
    lib.mgb_cov_param_count.restype = ctypes.c_int
    lib.mgb_cov_param_count.argtypes = [c_void_p]
    lib.mgb_cov_param_layout.restype = ctypes.c_int
    lib.mgb_cov_param_layout.argtypes = [c_void_p, POINTER(c_int64), POINTER(c_int64), POINTER(c_int64)]
    lib.mgb_cov_cat_sizes.restype = ctypes.c_int
    lib.mgb_cov_cat_sizes.argtypes = [c_void_p, POINTER(c_int32)]
    lib.mgb_cov_workspace_bytes.restype = c_size_t
    lib.mgb_cov_workspace_bytes.argtypes = [c_void_p, c_int32]
    lib.mgb_cov_forward.restype = ctypes.c_int
    lib.mgb_cov_forward.argtypes = [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                    POINTER(CovOutputs), c_void_p]
    lib.mgb_cov_rollout.restype = ctypes.c_int
    lib.mgb_cov_rollout.argtypes = [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int32, ctypes.c_uint64,
                                    c_void_p, POINTER(CovOutputs), c_void_p]
    lib.mgb_cov_policy.restype = ctypes.c_int
    lib.mgb_cov_policy.argtypes = [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, POINTER(CovOutputs), c_void_p]
    if hasattr(lib, 'mgb_cov_backward'):
        lib.mgb_cov_backward.restype = ctypes.c_int
        lib.mgb_cov_backward.argtypes = [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p]
    if hasattr(lib, 'mgb_ppo_loss'):
        lib.mgb_ppo_loss.restype = ctypes.c_int
        lib.mgb_ppo_loss.argtypes = [c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_double,
                                     c_double, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    if hasattr(lib, 'mgb_pack_observations'):
        lib.mgb_pack_observations.restype = ctypes.c_int
        lib.mgb_pack_observations.argtypes = [POINTER(CovConfig), c_int32, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.mgb_int_plan_create.restype = ctypes.c_int
    lib.mgb_int_plan_create.argtypes = [POINTER(IntConfig), POINTER(c_void_p)]
    lib.mgb_int_plan_destroy.restype = None
    lib.mgb_int_plan_destroy.argtypes = [c_void_p]
    lib.mgb_int_param_count.restype = ctypes.c_int
    lib.mgb_int_param_count.argtypes = [c_void_p]
    lib.mgb_int_param_layout.restype = ctypes.c_int
    lib.mgb_int_param_layout.argtypes = [c_void_p, POINTER(c_int64), POINTER(c_int64), POINTER(c_int64)]
    lib.mgb_int_workspace_bytes.restype = c_size_t
    lib.mgb_int_workspace_bytes.argtypes = [c_void_p, c_int32]
    lib.mgb_int_forward.restype = ctypes.c_int
    lib.mgb_int_forward.argtypes = [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                    POINTER(IntOutputs), c_void_p]
    lib.mgb_int_backward.restype = ctypes.c_int
    lib.mgb_int_backward.argtypes = [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p]
    lib.mgb_launch_count.restype = c_int64
    lib.mgb_launch_count.argtypes = []
    lib.mgb_profile_kernel.restype = ctypes.c_int
    lib.mgb_profile_kernel.argtypes = [c_char_p]
    lib.mgb_profile_read.restype = ctypes.c_int
    lib.mgb_profile_read.argtypes = [POINTER(c_double), POINTER(c_int64)]
    lib.mgb_scale_accumulate.restype = ctypes.c_int
    lib.mgb_scale_accumulate.argtypes = [c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_int32, c_void_p]
    lib.mgb_optim_scratch_bytes.restype = c_size_t
    lib.mgb_optim_scratch_bytes.argtypes = []
    lib.mgb_grad_norm.restype = ctypes.c_int
    lib.mgb_grad_norm.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]
    lib.mgb_adam_step.restype = ctypes.c_int
    lib.mgb_adam_step.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_double, c_double, c_double, c_double,
                                  c_double, c_int64, c_int32, c_int32, c_void_p, c_double, c_void_p]
    lib.mgb_profile_report.restype = ctypes.c_int
    lib.mgb_profile_report.argtypes = [c_char_p, c_int64]
    return lib


def make_config(zs: Sequence[int], canvas_size: int, min_max_distance: Tuple[float, float], network_width: int, maxl: int,
                num_cg_levels: int, num_channels_hidden: int, num_channels_per_element: int, num_gaussians: int,
                bag_scale: float, beta=None) -> CovConfig:
    cfg = CovConfig()
    cfg.canvas_size = canvas_size
    cfg.num_species = len(zs)
    for i, z in enumerate(zs):
        cfg.zs[i] = int(z)
    cfg.maxl = maxl
    cfg.num_cg_levels = num_cg_levels
    cfg.num_channels_hidden = num_channels_hidden
    cfg.num_channels_per_element = num_channels_per_element
    cfg.num_gaussians = num_gaussians
    cfg.network_width = network_width
    cfg.min_distance, cfg.max_distance = float(min_max_distance[0]), float(min_max_distance[1])
    cfg.bag_scale = float(bag_scale)
    cfg.has_beta = 0 if beta is None else 1
    cfg.beta = 0.0 if beta is None else float(beta)
    cfg.rel_sh_normalize = 0
    return cfg


def param_names(num_cg_levels: int, maxl: int = MAXL) -> List[str]:
    """Reference `named_parameters()` names of CovariantAC in the order of the flat buffer (include/molgym_b200.h)."""
    names = ['cg_model.input_func_atom.lin.weight', 'cg_model.input_func_atom.lin.bias']
    ells = range(maxl + 1)
    for k in range(num_cg_levels):
        rad = f'cg_model.rad_funcs.rad_funcs.{k}'
        names += [f'{rad}.scales', f'{rad}.phases']
        names += [f'{rad}.linear.{l}.weight' for l in ells]
        names += [f'{rad}.linear.{l}.bias' for l in ells]
        names += [f'cg_model.cormorant_cg.edge_levels.{k}.cat_mix.mix_reps.weights.{l}' for l in ells]
        names += [f'cg_model.cormorant_cg.atom_levels.{k}.cat_mix.mix_reps.weights.{l}' for l in ells]
    names += [f'cg_mix.cat_mix.mix_reps.weights.{l}' for l in ells]
    for head in ('phi_focus', 'phi_element', 'phi_d', 'phi_trans', 'phi_v'):
        names += [f'{head}.layers.0.weight', f'{head}.layers.0.bias', f'{head}.layers.1.weight', f'{head}.layers.1.bias']
    names.append('distance_log_stds')
    return names


def check(lib, rc: int):
    if rc != 0:
        raise RuntimeError(f'molgym_b200 C-ABI error {rc}: {lib.mgb_last_error().decode()}')


def make_int_config(zs: Sequence[int], canvas_size: int, min_max_distance: Tuple[float, float], network_width: int) -> IntConfig:
    cfg = IntConfig()
    cfg.canvas_size = canvas_size
    cfg.num_species = len(zs)
    for i, z in enumerate(zs):
        cfg.zs[i] = int(z)
    cfg.network_width = network_width
    cfg.min_distance, cfg.max_distance = float(min_max_distance[0]), float(min_max_distance[1])
    return cfg


def int_param_names() -> List[str]:
    """Reference `named_parameters()` names of SchNetAC in the order of the flat buffer (include/molgym_b200.h)."""
    names = ['embedding_fn.embedding.weight']
    for t in range(3):
        it = f'embedding_fn.interactions.{t}'
        names += [f'{it}.filter_network.0.weight', f'{it}.filter_network.0.bias', f'{it}.filter_network.1.weight',
                  f'{it}.filter_network.1.bias', f'{it}.cfconv.in2f.weight', f'{it}.cfconv.f2out.weight', f'{it}.cfconv.f2out.bias',
                  f'{it}.dense.weight', f'{it}.dense.bias']
    for head in ('phi_beta', 'phi_focus', 'phi_element', 'phi_continuous', 'phi_kappa'):
        names += [f'{head}.layers.0.weight', f'{head}.layers.0.bias', f'{head}.layers.1.weight', f'{head}.layers.1.bias']
    names += [f'critic.layers.{i}.{w}' for i in range(3) for w in ('weight', 'bias')]
    names.append('log_stds')
    return names
