"""ORACLE / TEST INFRASTRUCTURE — stand-in for `cormorant.models.cormorant_qm9.expand_var_list`
(used at molgym/agents/covariant/modules.py:6,34-40)."""


def expand_var_list(var, num_cg_levels):
    if isinstance(var, list):
        return var + (num_cg_levels - len(var)) * [var[-1]]
    if isinstance(var, (float, int)):
        return [var] * num_cg_levels
    raise ValueError('Incorrect type {}'.format(type(var)))
