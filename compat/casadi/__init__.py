"""`casadi` look-alike for do-mpc templates: re-exports do_mpc_b200.casadi_api.
Put `compat/` on sys.path ONLY when the real CasADi is absent and the controller runs on do_mpc_b200."""
from do_mpc_b200.casadi_api import *  # noqa: F401,F403
from do_mpc_b200.casadi_api import SX, MX, DM, Function  # noqa: F401
from . import tools  # noqa: F401
