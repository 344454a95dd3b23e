"""Stand-ins for the reference's un-installed third-party imports (mitsuba, drjit, ultraspy, matplotlib).

``install()`` puts this directory on sys.path ONLY for the packages that are not importable, so a real
Mitsuba / matplotlib installation always wins."""
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def install(force=()):
    missing = [m for m in ("mitsuba", "drjit", "ultraspy", "matplotlib") if m in force or importlib.util.find_spec(m) is None]
    if missing and _HERE not in sys.path:
        sys.path.append(_HERE)      # appended: anything genuinely installed is found first
    from .. import plugins  # noqa: F401  CustomIntegrator / CustomBSDF / CustomSensor / CustomEmmitter
    return missing
