from .cormorant_cg import CormorantCG  # noqa: F401
from .cormorant_qm9 import expand_var_list  # noqa: F401
