"""Reference-named plugin modules (CustomIntegrator, CustomBSDF, CustomSensor, CustomEmmitter).

Importing this package puts its directory on sys.path so that the reference's own import lines
(`from CustomIntegrator import UltraIntegrator`, /root/reference/USMain.py:14-23) resolve to the
B200-backed classes."""
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
if _here not in _sys.path:
    _sys.path.insert(0, _here)
