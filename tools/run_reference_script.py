#!/usr/bin/env python3
"""Run one of the reference's own scripts (USMain.py, TestScene.py) UNCHANGED against this repo.

    python tools/run_reference_script.py /root/reference/USMain.py

sys.path is arranged so that `CustomIntegrator`, `CustomBSDF`, `CustomSensor`, `CustomEmmitter` resolve to
physics-based-ray-tracing_b200/plugins and -- only where the real packages are not installed -- `mitsuba`,
`drjit`, `ultraspy`, `matplotlib` resolve to physics-based-ray-tracing_b200/shims.  The script's own directory
is deliberately NOT put on sys.path (runpy.run_path does not add it), otherwise the reference's Custom*.py
would shadow the B200-backed modules.
"""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    script = os.path.abspath(sys.argv[1])
    from prt_b200 import shims
    missing = shims.install()
    print(f"[run_reference_script] stand-ins active for: {missing or 'nothing (real packages found)'}", file=sys.stderr)
    sys.argv = [script] + sys.argv[2:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
