"""``UltraIntegrator`` -- drop-in for /root/reference/CustomIntegrator.py:12-412 -- and the ``path``
integrator scenes/cbox.xml:5-9 names, both running on the B200 engine.

Same constructor properties and defaults (:16-42), same attributes the driver reads afterwards
(``channel_buf``, ``transmission_delays_buf``, ``n_angles`` ... USMain.py:103-121), same methods:
``sample`` (stub, :52-53), ``simulate_acquisition`` (:60-232), ``simulate_acquisition_parallel`` (:235-405),
``traverse`` (exposes ``pitch``, :408-409), ``parameters_changed`` (:411).  Where the reference spends its
time in a Python loop of width-1 Dr.Jit launches, both methods here make ONE call into libprt_b200.so
(``prt_acquire`` / ``prt_acquire_dev``).

Extensions (not in the reference; defaults reproduce it): ``samples_per_element`` (paths per (angle,
element); the reference traces exactly 1), ``seed`` (the reference is unseeded), ``quirk_flags``
(SURVEY.md Appendix A), and sample-sharding over torch.distributed ranks when a process group exists.
"""
import numpy as np

from prt_b200 import capi
from prt_b200 import mi_compat as mi
from prt_b200.scene import AcqParams


def _as_float_array(x):
    a = np.asarray(x.numpy() if hasattr(x, "numpy") and not isinstance(x, np.ndarray) else x, dtype=np.float64)
    return a.reshape(-1)


class UltraIntegrator(mi.SamplingIntegrator):
    def __init__(self, props):
        super().__init__(props)
        # scene-independent ray tracing parameters (CustomIntegrator.py:16-23)
        self.max_depth = props.get('max_depth', 2)
        self.frequency = props.get('frequency', 5e6)
        self.sound_speed = props.get('sound_speed', 1540)
        self.attenuation = props.get('attenuation', 0.5)
        self.wave_cycles = props.get('wave_cycles', 5)
        self.main_beam_angle = props.get("main_beam_angle", 10)
        self.cutoff_angle = props.get("cutoff_angle", 20)
        self.fs = props.get('sampling_rate', 50e6)
        # transducer geometry (:26-30)
        self.n_elements = props.get('n_elements', 128)
        self.pitch = props.get('pitch', 0.00035)
        self.elem_x = mi.Float(self.pitch * (np.arange(self.n_elements, dtype=np.float32) - (self.n_elements - 1) / 2))
        self.elem_pos = mi.Vector3f(self.elem_x, 0, 0)
        self.trans_norm = mi.Vector3f(0, 0, 1)
        # plane-wave transmission (:33-34)
        angles = props.get('angles', None)
        self.angles = mi.Float(np.linspace(-30, 30, 25) if angles is None else _as_float_array(angles))
        self.n_angles = len(self.angles)
        self.init_amp, self.init_atten, self.init_tof = 1.0, 1.0, 0.0
        # echo accumulation buffers (:42-46)
        self.time_samples = props.get('time_samples', 3000)
        self.channel_buf = np.zeros(self.n_angles * self.n_elements * self.time_samples, dtype=np.float32)
        self.transmission_delays_buf = np.zeros(self.n_angles * self.n_elements, dtype=np.float32)
        self.ray_count = 0
        # extensions
        self.samples_per_element = int(props.get('samples_per_element', 1))
        self.seed = int(props.get('seed', 0))
        self.quirk_flags = int(props.get('quirk_flags', 0))
        self.max_path_len = float(props.get('max_path_len', 0.2))       # hard-coded 0.2 in the reference (:141)
        # pulse shaping: convolve the delta echoes with a `wave_cycles`-long Gaussian-modulated tone
        # burst (the prototype's model, RayTracingV0.py:185-204).  Off by default: the reference never uses wave_cycles.
        self.shape_pulse = bool(props.get('shape_pulse', False))
        self.last_stats = None

    # :52-53 -- the Mitsuba entry point is a stub in the reference
    def sample(self, scene, sampler, ray, medium, active=True):
        return mi.Color1f(0.0), active, []

    def acq_params(self, scene) -> AcqParams:
        sensors = scene.sensors()
        T = sensors[0].transform.matrix if sensors else np.eye(4)           # :101 / :272
        return AcqParams(n_elements=int(self.n_elements), pitch=float(self.pitch), angles_deg=_as_float_array(self.angles),
                         time_samples=int(self.time_samples), max_depth=int(self.max_depth), fs=float(self.fs),
                         sound_speed=float(self.sound_speed), frequency=float(self.frequency),
                         attenuation=float(self.attenuation), main_beam_deg=float(self.main_beam_angle),
                         cutoff_deg=float(self.cutoff_angle), max_path_len=float(self.max_path_len),
                         sensor_to_world=np.array(T, dtype=np.float64), quirk_flags=int(self.quirk_flags))

    def _acquire(self, scene, quirk_flags):
        p = self.acq_params(scene)
        p.quirk_flags = quirk_flags
        self.n_angles = p.n_angles
        dev = scene.device()
        spp = max(int(self.samples_per_element), 1)
        world, rank = 1, 0
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                world, rank = dist.get_world_size(), dist.get_rank()
        except ImportError:
            pass
        if world == 1:
            buf, tx, st = dev.acquire(p, seed=self.seed, spp=spp)
        else:
            from prt_b200.distributed import acquire_sharded
            buf, tx, st = acquire_sharded(dev, p, seed=self.seed, spp_total=spp, to_host=True)
        self.last_stats = st
        self.ray_count += int(st["segments"])
        if self.shape_pulse:
            from prt_b200.engine import pulse_shape
            buf = pulse_shape(buf, self.fs, self.frequency, wave_cycles=self.wave_cycles, context=dev.ctx)
        return buf, tx

    def simulate_acquisition(self, scene):
        """The Dr.Jit formulation (:60-232): flat buffers; out-of-range samples are clamped (:192) and tof only
        carries the last segment (:165,226) when ``quirk_flags`` asks for the literal "D" behaviour; by default the
        canonical path is traced, as in simulate_acquisition_parallel."""
        buf, tx = self._acquire(scene, self.quirk_flags)
        self.channel_buf = buf.reshape(-1)
        self.transmission_delays_buf = tx.reshape(-1)
        print("Simulation complete. Channel buffer populated. Transmission delays stored.")
        print(f"{self.ray_count} spawned in simulation")
        return True

    def simulate_acquisition_parallel(self, scene):
        """The thread-pool formulation the driver calls (USMain.py:99; :235-405): results left on ``self`` as
        ``channel_buf [n_angles, n_elements, time_samples] float32`` (:260) and ``transmission_delays_buf``
        (flat, :257)."""
        buf, tx = self._acquire(scene, self.quirk_flags)
        self.channel_buf = buf
        self.transmission_delays_buf = tx.reshape(-1)
        print("Simulation complete - traced", self.n_angles * self.n_elements * max(self.samples_per_element, 1), "primary rays")
        print("Channel buffer shape:", self.channel_buf.shape)
        return True

    def simulate_acquisition_variants(self, scene, key, values):
        """Extension for the driver's finite-difference loop (USMain.py:262-289): what
        ``params[key] = v; params.update(); simulate_acquisition_parallel(scene)`` yields for every v in ``values``,
        traced in ONE library call with common random numbers (prt_acquire_variants).  ``key`` is a mi.traverse key
        such as 'shape.bsdf.roughness'.  Returns ``[len(values), n_angles, n_elements, time_samples]`` float32; the
        scene's own parameter is left untouched."""
        targets = mi.traverse(scene)._targets.get(key)
        if not targets:
            raise KeyError(key)
        index = targets[0][1]
        if any(i != index for _, i in targets):
            raise ValueError(f"{key}: mixed parameters")
        p = self.acq_params(scene)
        p.quirk_flags = self.quirk_flags
        bufs, tx, st = scene.device().acquire_variants(p, [m for m, _ in targets], index, values, seed=self.seed,
                                                       spp=max(int(self.samples_per_element), 1))
        self.last_stats = st
        self.transmission_delays_buf = tx.reshape(-1)
        if self.shape_pulse:
            from prt_b200.engine import pulse_shape
            bufs = pulse_shape(bufs, self.fs, self.frequency, wave_cycles=self.wave_cycles, context=scene.device().ctx)
        return bufs

    def render_bmode(self, scene, x_scan, z_scan, dynamic_range=60.0, f_number=1.0, t0=0.0):
        """Extension: everything us_render() of the driver does after ``scene.integrator()`` (USMain.py:99-224) --
        acquisition, delay-and-sum, envelope, log compression -- in ONE library call with the channel data resident on the
        device (prt_us_render).  Returns the driver's ``display_image`` ([len(z_scan), len(x_scan)], values in [0, 1]);
        the envelope is left on ``self.last_envelope``.  With an initialised process group the acquisition is
        sample-sharded over the ranks and every rank develops the image from the all-reduced device buffer."""
        p = self.acq_params(scene)
        p.quirk_flags = self.quirk_flags
        world = 1
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                world = dist.get_world_size()
        except ImportError:
            pass
        if world > 1:
            # sample shards: every rank traces its samples, the channel buffers are summed over NVLink (one all-reduce, pipelined
            # by steering angle), and every rank develops the image from the summed buffer without it leaving the device
            import torch
            from prt_b200.distributed import acquire_sharded
            dev = scene.device()
            buf, _, stats = acquire_sharded(dev, p, seed=self.seed, spp_total=max(int(self.samples_per_element), 1), to_host=False)
            stream = torch.cuda.current_stream(buf.device)
            img, env = dev.us_postprocess_dev(p, buf.data_ptr(), x_scan, z_scan, stream=stream.cuda_stream, t0=t0, f_number=f_number,
                                              dynamic_range=dynamic_range, shape_pulse=self.shape_pulse, wave_cycles=self.wave_cycles)
            hs = stats.cpu().numpy()
            st = dict(paths=int(hs[0]), segments=int(hs[1]), rays=int(hs[2]), deposits=int(hs[3]), misses=int(hs[4]))
            self.last_stats, self.last_envelope = st, env
            self.ray_count += int(st["segments"])
            return img
        img, env, st = scene.device().us_render(p, x_scan, z_scan, seed=self.seed, spp=max(int(self.samples_per_element), 1),
                                                t0=t0, f_number=f_number, dynamic_range=dynamic_range,
                                                shape_pulse=self.shape_pulse, wave_cycles=self.wave_cycles)
        self.last_stats, self.last_envelope = st, env
        self.ray_count += int(st["segments"])
        return img

    def traverse(self, callback):
        callback.put_parameter('pitch', self.pitch, mi.ParamFlags.Differentiable)

    def parameters_changed(self, keys=None):
        pass


class PathIntegrator(mi.SamplingIntegrator):
    """Mitsuba's `path` integrator as scenes/cbox.xml:5-9 configures it (max_depth, rr_depth 5); SURVEY.md C.7."""

    def __init__(self, props):
        super().__init__(props)
        self.max_depth = int(props.get("max_depth", -1))
        self.rr_depth = int(props.get("rr_depth", 5))
        self.last_stats = None

    def render_params(self, scene, sensor=0):
        s = scene.sensors()[sensor]
        w, h = s.film_size()
        rp = capi.RenderParamsC()
        m = np.array(s.transform.matrix, dtype=np.float64).reshape(16)
        for i in range(16):
            rp.to_world[i] = m[i]
        fov = s.fov
        axis = s.fov_axis
        # convert to the fov along the smaller axis, which is what the kernel takes
        if axis in ("x", "y", "larger", "diagonal"):
            aspect = w / h
            fx = fov if axis == "x" else None
            if axis == "y":
                fx = 2 * np.degrees(np.arctan(np.tan(np.radians(fov) / 2) * aspect))
            elif axis == "larger":
                fx = fov if w >= h else 2 * np.degrees(np.arctan(np.tan(np.radians(fov) / 2) * aspect))
            elif axis == "diagonal":
                diag = 2 * np.tan(np.radians(fov) / 2)
                fx = 2 * np.degrees(np.arctan(diag / (2 * np.sqrt(1 + 1 / aspect ** 2))))
            fy = 2 * np.degrees(np.arctan(np.tan(np.radians(fx) / 2) / aspect))
            fov = fx if w <= h else fy
        rp.fov_deg = float(fov)
        rp.near_clip, rp.far_clip = s.near_clip, s.far_clip
        rp.width, rp.height = w, h
        rp.max_depth = self.max_depth if self.max_depth >= 0 else 1 << 20
        rp.rr_depth = self.rr_depth
        rf = s._rfilter.plugin_name() if s._rfilter is not None else "gaussian"
        rp.rfilter = 1 if rf == "tent" else 0
        return rp

    def render(self, scene, sensor=0, seed=0, spp=0):
        rp = self.render_params(scene, sensor)
        spp = spp or scene.sensors()[sensor].sample_count()
        dev = scene.device()
        world = 1
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                world = dist.get_world_size()
        except ImportError:
            pass
        if world == 1:
            img, st = dev.render_image(rp, seed=seed, spp=spp)
        else:
            from prt_b200.distributed import render_sharded
            img, st = render_sharded(dev, rp, seed=seed, spp_total=spp, to_host=True, develop=True)
        self.last_stats = st
        return img

    def sample(self, scene, sampler, ray, medium, active=True):
        raise NotImplementedError("per-ray sample() is not exposed; use render()")
