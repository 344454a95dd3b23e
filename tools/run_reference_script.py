#!/usr/bin/env python3
"""Run one of the reference's own scripts (USMain.py, TestScene.py) UNCHANGED against this repo.

    python tools/run_reference_script.py /root/reference/USMain.py [--timing out.json]

--timing wraps (from OUTSIDE the script: its text is executed byte for byte) the plugin's
``UltraIntegrator.simulate_acquisition_parallel`` and the beamformer's ``beamform`` / ``compute_envelope`` with wall-clock
timers and writes {calls, per-call ms, total} to the given JSON file when the script ends.

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


def _install_timers(log):
    import time
    import CustomIntegrator
    from ultraspy.beamformers.das import DelayAndSum

    def wrap(cls, name, key):
        inner = getattr(cls, name)

        def timed(self, *a, **k):
            t0 = time.perf_counter()
            try:
                return inner(self, *a, **k)
            finally:
                log.setdefault(key, []).append(1e3 * (time.perf_counter() - t0))
        setattr(cls, name, timed)
    wrap(CustomIntegrator.UltraIntegrator, "simulate_acquisition_parallel", "simulate_acquisition_parallel_ms")
    wrap(DelayAndSum, "beamform", "beamform_ms")
    wrap(DelayAndSum, "compute_envelope", "compute_envelope_ms")


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    script = os.path.abspath(sys.argv[1])
    args = sys.argv[2:]
    timing = None
    if "--timing" in args:
        i = args.index("--timing")
        timing = args[i + 1]
        del args[i:i + 2]
    from prt_b200 import shims
    missing = shims.install()
    print(f"[run_reference_script] stand-ins active for: {missing or 'nothing (real packages found)'}", file=sys.stderr)
    log = {}
    if timing:
        _install_timers(log)
    sys.argv = [script] + args
    import time
    t0 = time.perf_counter()
    try:
        runpy.run_path(script, run_name="__main__")
    finally:
        if timing:
            import json
            out = {"script": os.path.basename(script), "wall_s": time.perf_counter() - t0}
            for k, v in log.items():
                out[k] = {"calls": len(v), "first": v[0], "median": sorted(v)[len(v) // 2], "total": sum(v)}
            with open(timing, "w") as f:
                json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
