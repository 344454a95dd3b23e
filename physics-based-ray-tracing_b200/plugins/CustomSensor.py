"""Sensors: ``UltraSensor`` (what /root/reference/USMain.py:17-18 imports and every XML names as
``ultrasound_sensor``; it only survives in the reference's stale __pycache__/CustomSensor.cpython-312.pyc --
SURVEY.md Appendix B), ``CustomSensor`` (the class the current /root/reference/CustomSensor.py:7-73 defines:
a per-element RF buffer with ``put_data``), and the ``perspective`` camera scenes/cbox.xml:11-21 names.
"""
import numpy as np

from prt_b200 import mi_compat as mi


class UltraSensor(mi.Sensor):
    """Properties / defaults per the recovered pyc (Appendix B lines 11-32).  The acquisition kernel reads
    exactly one attribute: ``transform`` (CustomIntegrator.py:101,272)."""

    def __init__(self, props):
        super().__init__(props)
        self.num_elements_lateral = props.get('num_elements_lateral', 128)
        self.element_width = props.get('elements_width', 0.003)
        self.element_height = props.get('elements_height', 0.01)
        self.pitch = props.get('pitch', 0.00035)
        self.radius = props.get('radius', float('inf'))
        self.center_frequency = props.get('center_frequency', 5e6)
        self.sound_speed = props.get('sound_speed', 1540)
        to_world_prop = props.get('to_world', mi.ScalarTransform4f())
        self.transform = mi.Transform4f(to_world_prop.matrix) if hasattr(to_world_prop, 'matrix') else to_world_prop
        self.emission_time = mi.Float(0)
        self.directivity = props.get('directivity', 1.0)

    def sample_ray(self, time, wavelength_sample, position_sample, aperture_sample, active=True):
        """Image-mode ray generation (pyc lines 37-90): random element from position_sample.x, jitter inside the
        element by aperture_sample, uniform-hemisphere direction, weight cos(2 pi f t) |d.z| directivity."""
        self.emission_time = time
        ps, ap = mi.Vector2f(position_sample), mi.Vector2f(aperture_sample)
        n = self.num_elements_lateral
        idx = np.minimum(np.floor(np.asarray(ps.x, dtype=np.float64) * n), n - 1)
        if np.isinf(self.radius):
            start_x = -((n - 1) * self.pitch) / 2
            ex, ez = start_x + idx * self.pitch, np.zeros_like(idx)
        else:
            th = (idx - n / 2) * (self.pitch / self.radius)
            ex, ez = self.radius * np.sin(th), self.radius * (1 - np.cos(th))
        ox = (np.asarray(ap.x, dtype=np.float64) - 0.5) * self.element_width
        oy = (np.asarray(ap.y, dtype=np.float64) - 0.5) * self.element_height
        origin_local = np.stack([ex + ox, oy, ez], -1)
        d_local = np.asarray(mi.warp.square_to_uniform_hemisphere(ap), dtype=np.float64)
        o_w = self.transform.transform_point(origin_local)
        d_w = self.transform.transform_vector(d_local)
        d_w = d_w / np.linalg.norm(d_w, axis=-1, keepdims=True)
        t = np.asarray(time, dtype=np.float64)
        weight = np.cos(2 * np.pi * self.center_frequency * t) * np.abs(d_local[..., 2]) * self.directivity
        return mi.Ray3f(o_w, d_w), mi.Float(weight)

    def traverse(self, callback):
        pass


class CustomSensor(mi.Sensor):
    """/root/reference/CustomSensor.py:7-73: RF buffer [number_of_elements, time_samples] filled by put_data."""

    def __init__(self, props):
        super().__init__(props)
        self.number_of_elements = props.get("number_of_elements", 128)
        self.pitch = props.get("pitch", 0.0003)
        self.element_width = props.get("element_width", 0.00027)
        self.element_height = props.get("element_height", 0.005)
        self.sample_rate = props.get("sample_rate", 50e6)
        self.speed_of_sound = props.get("speed_of_sound", 1540.0)
        self.time_samples = props.get("time_samples", 3000)
        self.channel_buffer = np.zeros((self.number_of_elements, self.time_samples), dtype=np.float32)

    def put_data(self, ray, amplitude, active=True):
        x = float(np.asarray(ray.o)[..., 0].reshape(-1)[0])
        idx = int(np.round(x / self.pitch + self.number_of_elements / 2))        # :36 (round-half-even)
        t = float(np.asarray(ray.time).reshape(-1)[0])
        index = int(np.round(t * self.sample_rate))                              # :43
        d = -np.asarray(ray.d, dtype=np.float64).reshape(-1)[:3]
        d = d / np.linalg.norm(d)
        gain = max(0.0, float(d @ np.array([0.0, 0.0, 1.0])))                    # :46-51
        if 0 <= idx < self.number_of_elements and 0 <= index < self.time_samples:   # :58
            self.channel_buffer[idx, index] += np.float32(float(np.asarray(amplitude).reshape(-1)[0]) * gain)

    def channel_data(self):
        return self.channel_buffer

    def clear(self):
        self.channel_buffer = np.zeros((self.number_of_elements, self.time_samples), dtype=np.float32)

    def traverse(self, callback):
        for k in ("number_of_elements", "pitch", "element_width", "element_height", "sample_rate", "speed_of_sound"):
            callback.put_parameter(k, getattr(self, k), mi.ParamFlags.NonDifferentiable)

    def parameters_changed(self, keys=None):
        self.clear()

    parameters = parameters_changed


class PerspectiveSensor(mi.Sensor):
    """The `perspective` camera of scenes/cbox.xml:11-21 (fov along the smaller axis, look_at pose)."""

    def __init__(self, props):
        super().__init__(props)
        self.fov = float(props.get("fov", 39.3077))
        self.fov_axis = props.get("fov_axis", "x")
        self.near_clip = float(props.get("near_clip", 1e-2))
        self.far_clip = float(props.get("far_clip", 1e4))
        tw = props.get("to_world", mi.ScalarTransform4f())
        self.transform = mi.Transform4f(tw.matrix) if hasattr(tw, "matrix") else tw
        self._film = self._sampler = self._rfilter = None

    def film_size(self):
        f = self._film
        return (int(f.get("width", 768)), int(f.get("height", 576))) if f is not None else (768, 576)

    def sample_count(self):
        return int(self._sampler.get("sample_count", 4)) if self._sampler is not None else 4
