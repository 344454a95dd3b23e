"""GPU: the reference driver's call sequence (/root/reference/USMain.py:12-24,92-224,257-289) through the
stand-in packages, and the DAS beamformer ("next" row f1) against a numpy / scipy restatement."""
import os

import numpy as np
import pytest

from conftest import ROOT
from prt_b200 import scenes

pytestmark = pytest.mark.gpu


def _das_numpy(ch, angles_deg, x, z, fs, c, pitch, t0, f_number):
    import pyref
    return pyref.das_beamform(ch, angles_deg, x, z, fs, c, pitch, t0, f_number)


def test_das_matches_numpy():
    from scipy.signal import hilbert
    from prt_b200.engine import das_beamform, envelope
    rng = np.random.default_rng(0)
    n_a, n_e, T = 3, 16, 600
    # smooth (band-limited) RF so that fp32 time-of-flight rounding cannot flip a whole sample
    tt = np.arange(T) / 50e6
    ch = np.stack([[np.sin(2 * np.pi * 3e6 * tt + 0.3 * e + a) * np.exp(-((tt - 6e-6 - 1e-7 * e) / 2e-6) ** 2)
                    for e in range(n_e)] for a in range(n_a)]).astype(np.float32)
    angles = np.array([-10.0, 0.0, 10.0])
    x = np.linspace(-0.004, 0.004, 37)
    z = np.linspace(0.002, 0.012, 101)
    fs, c, pitch = 50e6, 1540.0, 3e-4
    rf, env = das_beamform(ch, angles, x, z, fs, c, pitch, t0=0.0, f_number=0.0)
    ref = _das_numpy(ch.astype(np.float64), angles, x, z, fs, c, pitch, 0.0, 0.0)
    assert rf.shape == (37, 101)
    assert np.abs(rf - ref).max() <= 2e-3 * np.abs(ref).max()       # fp32 time-of-flight * 50 MHz interpolation
    # with an f-number the aperture edge may include / exclude one element where |dx| 2 f# == z to within an ulp
    rf2, _ = das_beamform(ch, angles, x, z, fs, c, pitch, t0=0.0, f_number=0.83)
    ref2 = _das_numpy(ch.astype(np.float64), angles, x, z, fs, c, pitch, 0.0, 0.83)
    assert (np.abs(rf2 - ref2) <= 2e-3 * np.abs(ref2).max()).mean() > 0.995
    ref_env = np.abs(hilbert(rf.astype(np.float64), axis=1))
    assert np.abs(env - ref_env).max() <= 1e-4 * ref_env.max()
    assert np.abs(envelope(rf) - env).max() <= 1e-6 * ref_env.max()


def test_pulse_shape_matches_oracle():
    """Row f4: delta echoes -> Gaussian-modulated tone bursts (RayTracingV0.py:185-204) against the float64 numpy oracle;
    ragged sizes (T not a multiple of the 1024-sample tile, echoes at both row ends), one-echo analytic check."""
    import pyref
    from prt_b200.engine import pulse_shape
    rng = np.random.default_rng(3)
    fs, fc = 50e6, 3e6
    for (rows, T, cycles) in [((3, 5), 2500, 5), ((1,), 1024, 2), ((2,), 37, 1), ((4,), 10000, 8)]:
        ch = np.zeros(rows + (T,), dtype=np.float32)
        k = max(T // 40, 3)
        flat = ch.reshape(-1, T)
        for r in range(flat.shape[0]):
            idx = rng.integers(0, T, size=k)
            flat[r, idx] = rng.normal(size=k).astype(np.float32)
            flat[r, 0], flat[r, T - 1] = 1.0, -0.5            # echoes on both edges: the halo is zero-padded
        sigma = cycles / (4 * fc)
        got = pulse_shape(ch, fs, fc, wave_cycles=cycles)
        ref = pyref.pulse_shape(ch, fs, fc, sigma)
        assert got.shape == ch.shape and got.dtype == np.float32
        assert np.abs(got - ref).max() <= 2e-5 * max(np.abs(ref).max(), 1.0), (rows, T, cycles)
    # a single unit echo at sample 500 reproduces the prototype's pulse() sampled on the grid
    ch = np.zeros((1, 1000), dtype=np.float32)
    ch[0, 500] = 2.0
    t = (np.arange(1000) - 500) / fs
    sigma = 1e-7
    expect = 2.0 * np.sin(2 * np.pi * fc * t) * np.exp(-t ** 2 / sigma ** 2)
    expect[np.abs(t) > 4 * sigma + 0.5 / fs] = 0.0
    assert np.abs(pulse_shape(ch, fs, fc, sigma_s=sigma)[0] - expect).max() < 1e-5
    # linearity: shaping is a linear map of the channel data
    a, b = rng.normal(size=(2, 3, 700)).astype(np.float32)
    lhs = pulse_shape(a + 2 * b, fs, fc, wave_cycles=3)
    rhs = pulse_shape(a, fs, fc, wave_cycles=3) + 2 * pulse_shape(b, fs, fc, wave_cycles=3)
    assert np.abs(lhs - rhs).max() <= 1e-4 * np.abs(lhs).max()


@pytest.mark.parametrize("shape_pulse", [False, True])
def test_us_render_equals_the_staged_pipeline(shape_pulse):
    """prt_us_render (acquisition -> pulse shaping -> DAS -> envelope -> log compression on the device) against the same
    stages run one by one the way the driver does (USMain.py:99-224: simulate_acquisition_parallel, DelayAndSum.beamform,
    compute_envelope, numpy log compression), same seed."""
    from prt_b200 import mi_compat as mi
    from prt_b200.engine import das_beamform
    d = scenes.usmain_scene_dict()
    d["integrator"]["angles"] = np.linspace(-15, 15, 5)
    d["integrator"]["samples_per_element"] = 32
    d["integrator"]["shape_pulse"] = shape_pulse
    scene = mi.load_dict(d)
    integ = scene.integrator()
    lam = integ.sound_speed / integ.frequency
    x = np.arange(-0.02, 0.02 + lam / 2, lam / 2)
    z = np.arange(0.001, 0.05 + lam / 4, lam / 4)              # ragged: neither a multiple of 32 nor of 256
    img = integ.render_bmode(scene, x, z, dynamic_range=60.0, f_number=1.0)
    st = integ.last_stats
    assert img.shape == (len(z), len(x)) and img.dtype == np.float32
    assert 0.0 <= img.min() and img.max() == pytest.approx(1.0, abs=1e-6)
    integ.simulate_acquisition_parallel(scene)                   # applies the pulse shaping itself when shape_pulse is set
    assert integ.last_stats["paths"] == st["paths"] and integ.last_stats["deposits"] == st["deposits"]
    rf, env = das_beamform(integ.channel_buf, integ.angles.numpy(), x, z, integ.fs, integ.sound_speed, integ.pitch, f_number=1.0)
    assert np.abs(integ.last_envelope - env).max() <= 1e-4 * env.max()
    db = 20 * np.log10(env.astype(np.float32) + np.float32(1e-12))
    mx = db.max()
    ref = ((np.clip(db, mx - 60, mx) - (mx - 60)) / 60).T
    # the display image is a log of the envelope: compare where the envelope is well above the 60 dB floor, and the
    # clipped floor itself
    assert np.abs(img - ref).max() <= 2e-3
    assert ((img == 0) == (ref == 0)).mean() > 0.999
    # the multi-GPU form of the same call: the acquisition lands in a DEVICE buffer (what the all-reduce of the sample shards
    # leaves behind) and prt_us_postprocess_dev develops the image from it
    import torch
    from prt_b200.distributed import acquire_sharded
    dev = scene.device()
    p = integ.acq_params(scene)
    buf, _, _ = acquire_sharded(dev, p, seed=integ.seed, spp_total=32, to_host=False)
    img2, env2 = dev.us_postprocess_dev(p, buf.data_ptr(), x, z, stream=torch.cuda.current_stream().cuda_stream, f_number=1.0,
                                        dynamic_range=60.0, shape_pulse=shape_pulse, wave_cycles=integ.wave_cycles)
    assert np.abs(env2 - integ.last_envelope).max() <= 1e-4 * env.max() and np.abs(img2 - img).max() <= 2e-3


def test_sharded_entry_points_single_rank():
    """distributed.acquire_sharded / render_sharded without a process group (world size 1) == the plain calls; the
    multi-rank arithmetic (shard_samples + one sum all-reduce) is covered by the gloo test and by sharded == unsharded."""
    torch = pytest.importorskip("torch")
    from prt_b200 import mi_compat as mi
    from prt_b200.distributed import acquire_sharded, render_sharded
    from prt_b200.scene import AcqParams
    desc = scenes.ultrasound_scene("Plate_Box", "intended")
    scene = mi.Scene(desc)
    p = AcqParams.from_props(desc.integrator, desc.sensor)
    dev = scene.device()
    ref, rtx, rst = dev.acquire(p, seed=3, spp=64)
    for _ in range(2):          # second call reuses the cached device tensors
        got, tx, st = acquire_sharded(dev, p, seed=3, spp_total=64)
        assert st["paths"] == rst["paths"] and st["segments"] == rst["segments"] and st["deposits"] == rst["deposits"]
        assert np.allclose(tx, rtx) and np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()
    d_buf, _, d_st = acquire_sharded(dev, p, seed=3, spp_total=64, to_host=False)
    assert d_buf.is_cuda and int(d_st[0]) == rst["paths"]
    cb = mi.Scene(scenes.cbox_scene(48, 8))
    rp = cb.integrator().render_params(cb)
    img, ist = cb.device().render_image(rp, seed=5, spp=8)
    got, gst = render_sharded(cb.device(), rp, seed=5, spp_total=8, develop=True)
    assert gst["rays"] == ist["rays"] and got.shape == (48, 48, 3)
    assert np.allclose(got, img, rtol=2e-4, atol=1e-6)


def test_usmain_call_sequence():
    """What USMain.py does, with 2 optimisation iterations instead of 25 and 64 samples per element so that the
    finite-difference loss is not pure noise."""
    from prt_b200 import shims
    shims.install()
    import drjit as dr
    import mitsuba as mi
    from ultraspy.beamformers.das import DelayAndSum
    from ultraspy.probes.factory import build_probe
    from ultraspy.scan import GridScan
    mi.set_variant("llvm_ad_mono")
    from CustomIntegrator import UltraIntegrator
    from CustomSensor import UltraSensor
    from CustomEmmitter import CustomEmitter
    from CustomBSDF import UltraBSDF
    mi.register_integrator("ultrasound_integrator", UltraIntegrator)
    mi.register_sensor("ultrasound_sensor", UltraSensor)
    mi.register_emitter('ultrasound_emitter', CustomEmitter)
    mi.register_bsdf('ultrasound_bsdf', UltraBSDF)
    d = scenes.usmain_scene_dict()
    d["integrator"]["angles"] = dr.linspace(mi.Float, -15, 15, 5)
    d["integrator"]["samples_per_element"] = 64
    scene = mi.load_dict(d)
    params = mi.traverse(scene)

    def us_render(scene):
        integrator = scene.integrator()
        assert integrator.simulate_acquisition_parallel(scene) is True
        channel_buf, delays = integrator.channel_buf, integrator.transmission_delays_buf
        n_a, n_e, T = integrator.n_angles, integrator.n_elements, integrator.time_samples
        assert channel_buf.shape == (n_a, n_e, T) and channel_buf.dtype == np.float32
        assert np.isfinite(np.sum(channel_buf)) and np.max(channel_buf) > 0
        assert np.allclose(integrator.angles.numpy(), [-15, -7.5, 0, 7.5, 15])
        channel_data = channel_buf.reshape((n_a, n_e, T))
        tx = delays.reshape((n_a, n_e))
        probe = build_probe(geometry_type='linear', nb_elements=n_e, pitch=integrator.pitch, central_freq=integrator.frequency, bandwidth=70)
        seq = {'emitted': np.tile(np.arange(n_e), (n_a, 1)), 'received': np.tile(np.arange(n_e), (n_a, 1))}
        info = {'sampling_freq': integrator.fs, 't0': 0, 'prf': None, 'signal_duration': None, 'delays': tx,
                'sound_speed': integrator.sound_speed, 'sequence_elements': seq}
        bf = DelayAndSum(on_gpu=False)
        bf.automatic_setup(info, probe)
        assert str(bf)
        lam = integrator.sound_speed / integrator.frequency
        x_scan = np.arange(-0.04, 0.04 + lam / 4, lam / 4)
        z_scan = np.arange(0.001, 0.05 + lam / 4, lam / 4)
        scan = GridScan(x_scan, z_scan)
        out = bf.beamform(channel_data[np.newaxis][0], scan)
        env = bf.compute_envelope(out, scan).astype(np.float32)
        assert env.shape == (len(x_scan), len(z_scan)) == (1040, 638)
        db = 20 * np.log10(env + 1e-12)
        mx = np.max(db)
        img = (np.clip(db, mx - 60, mx) - (mx - 60)) / 60
        return img.T

    ref = us_render(scene)
    assert ref.shape == (638, 1040) and 0.0 <= ref.min() and ref.max() == pytest.approx(1.0)
    assert (ref > 0).mean() > 0.01

    def forward(rough):
        params['shape.bsdf.roughness'] = rough            # USMain.py:264
        params.update()
        return us_render(scene)

    f0 = np.mean((forward(0.1) - ref) ** 2)
    f1 = np.mean((forward(0.7) - ref) ** 2)
    assert np.isfinite(f0) and np.isfinite(f1)
    assert f1 < f0           # same seed, same roughness as the reference render (0.7) -> identical image
    assert f1 == 0.0


def _reference_dir():
    """Where a copy of the reference's scripts can be found: the build container mounts /root/reference; on a GPU box it only
    exists if the operator shipped one (PRT_REFERENCE_DIR, or the git-ignored .scratch_ref/ next to the repo's files -- the
    reference's sources are never committed)."""
    for d in (os.environ.get("PRT_REFERENCE_DIR"), "/root/reference", os.path.join(ROOT, ".scratch_ref")):
        if d and os.path.isfile(os.path.join(d, "USMain.py")):
            return d
    return None


@pytest.mark.gpu
@pytest.mark.skipif(_reference_dir() is None, reason="no copy of the reference's USMain.py on this box")
def test_unmodified_usmain_runs_to_completion(tmp_path):
    """/root/reference/USMain.py, byte for byte, to its last line on a GPU (USMain.py:256-298): the reference image, then 25
    finite-difference iterations x 2 forwards = 51 acquisitions + 51 beamforming passes, and the printed loss trace."""
    import json
    import subprocess
    import sys
    timing = tmp_path / "timing.json"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_reference_script.py"),
                        os.path.join(_reference_dir(), "USMain.py"), "--timing", str(timing)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    iters = [l for l in r.stdout.splitlines() if l.startswith("iter ")]
    assert len(iters) == 25 and iters[0].startswith("iter 0: loss=") and "Final Roughness Value:" in r.stdout
    rough = [float(l.split("rough=")[1]) for l in iters]
    assert all(1e-4 <= x <= 1.0 for x in rough)
    t = json.load(open(timing))
    assert t["simulate_acquisition_parallel_ms"]["calls"] == 51 and t["beamform_ms"]["calls"] == 51
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "usmain_gpu.json"), "w") as f:
            json.dump({"timing": t, "loss_trace": iters, "tail": r.stdout.splitlines()[-3:]}, f, indent=1)
