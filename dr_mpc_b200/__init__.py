"""Importable alias of the `dr-mpc_b200/` package directory (a hyphen cannot appear in a Python
module name): `import dr_mpc_b200` loads the code that lives in ../dr-mpc_b200/."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "dr-mpc_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
del _f
