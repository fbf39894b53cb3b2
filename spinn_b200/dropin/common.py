"""Flat-name shim: put this directory ahead of the reference's code/ on sys.path and the
reference's unchanged problem files (`from common import ...`) get the B200 implementation."""
from spinn_b200.common import *  # noqa: F401,F403
from spinn_b200.common import __dict__ as _d
globals().update({k: v for k, v in _d.items() if not k.startswith("__")})
