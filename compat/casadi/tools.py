"""`casadi.tools` look-alike (entry / struct_symSX / struct_SX ...)."""
from do_mpc_b200.sym.struct import entry, struct_symSX, struct_SX, struct_symMX, struct_MX, SymStruct, NumStruct  # noqa: F401
from do_mpc_b200.casadi_api import *  # noqa: F401,F403
from do_mpc_b200.differentiator import indexf  # noqa: F401,E402
