"""Time the stages of one us_render() of the reference driver (USMain.py:92-224) at its own sizes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from prt_b200 import mi_compat as mi, scenes
from prt_b200.engine import das_beamform, pulse_shape

d = scenes.usmain_scene_dict()
scene = mi.load_dict(d)
integ = scene.integrator()
print("angles", integ.n_angles, "elements", integ.n_elements, "T", integ.time_samples, "spp", integ.samples_per_element)
lam = integ.sound_speed / integ.frequency
x = np.arange(-0.04, 0.04 + lam / 4, lam / 4)
z = np.arange(0.001, 0.05 + lam / 4, lam / 4)
for it in range(4):
    t0 = time.perf_counter()
    integ.simulate_acquisition_parallel(scene)
    t1 = time.perf_counter()
    ch = integ.channel_buf
    scene.device().ctx.profile_begin()
    rf, env = das_beamform(ch, integ.angles.numpy(), x, z, integ.fs, integ.sound_speed, integ.pitch)
    t2 = time.perf_counter()
    print("   kernel times:", scene.device().ctx.profile_read())
    db = 20 * np.log10(env + 1e-12)
    mx = db.max()
    img = (np.clip(db, mx - 60, mx) - (mx - 60)) / 60
    t3 = time.perf_counter()
    ps = pulse_shape(ch, integ.fs, integ.frequency, wave_cycles=5)
    t4 = time.perf_counter()
    print(f"iter {it}: acquire {1e3*(t1-t0):.2f} ms, DAS+envelope {1e3*(t2-t1):.2f} ms ({x.size}x{z.size} px), log-compress (numpy) {1e3*(t3-t2):.2f} ms, pulse_shape {1e3*(t4-t3):.2f} ms; lib kernel {integ.last_stats['kernel_ms']:.3f} ms")

for it in range(4):
    t0 = time.perf_counter()
    img = integ.render_bmode(scene, x, z, dynamic_range=60.0)
    t1 = time.perf_counter()
    st = integ.last_stats
    print(f"render_bmode iter {it}: wall {1e3*(t1-t0):.2f} ms (device {st['kernel_ms']:.3f} ms, with D2H {st['total_ms']:.3f} ms), image {img.shape}")
