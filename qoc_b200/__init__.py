"""Import shim: the product package lives in `quantumoptimalcontrol.jl_b200/` (a directory name Python's import
statement cannot spell because of the dot).  `import qoc_b200` exposes it under an importable name."""
import os as _os

__path__.append(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                              "quantumoptimalcontrol.jl_b200"))
from .grape import *  # noqa: F401,F403,E402
from . import grape, _lib, build, fidelities, configs, sharding, callbacks, pulse_io  # noqa: F401,E402
from .fidelities import *  # noqa: F401,F403,E402
