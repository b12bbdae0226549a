from kelvin_oracle.cc_equations import *  # noqa: F401,F403
from kelvin_oracle.cc_equations import (  # noqa: F401
    _Stanton, _u_Stanton, _LS_TS, _u_LS_TS, _Lambda_opt, _uccsd_Lambda_opt,
    _S_S, _S_D, _D_S, _D_D, _D_DD, _LS_LS, _LS_LD, _LD_LS, _LD_LD, _LD_LDTD)
