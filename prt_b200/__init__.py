"""Import alias for the product package.

The product lives in ``physics-based-ray-tracing_b200/`` (the directory name the build contract
asks for); a hyphen cannot appear in a Python import, so ``import prt_b200`` maps onto that
directory.
"""
import os as _os

_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
_impl = _os.path.join(_root, "physics-based-ray-tracing_b200")
__path__.insert(0, _impl)
with open(_os.path.join(_impl, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_impl, "__init__.py"), "exec"))
