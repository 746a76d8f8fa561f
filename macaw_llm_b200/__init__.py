"""Importable alias of the `macaw-llm_b200/` package directory (a hyphen cannot appear in a Python module name).

`import macaw_llm_b200` executes macaw-llm_b200/__init__.py with this module's namespace and resolves
sub-modules (`macaw_llm_b200.ops`, `macaw_llm_b200.modeling`, ...) from that directory.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "macaw-llm_b200")
__path__ = [_real]
_init = _os.path.join(_real, "__init__.py")
exec(compile(open(_init).read(), _init, "exec"))
