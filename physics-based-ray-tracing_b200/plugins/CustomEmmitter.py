"""``CustomEmitter`` -- drop-in for /root/reference/CustomEmmitter.py (module name keeps the reference's
double-m).  The reference class cannot be constructed (typo at :25, float gather index at :57-61) and is
never called by the hot path; this one implements what those lines specify: linear (radius == 0) or convex
element geometry (:30-49), ``sample_position`` (:51-79), ``sample_ray`` with a uniform steering angle and the
plane-wave delay -(x sin psi)/c (:81-107), ``traverse`` keys (:114-124).  It is host-side parameter code; the
acquisition kernel generates its own primary rays (CustomIntegrator.py:97-107)."""
import numpy as np

from prt_b200 import mi_compat as mi


class CustomEmitter(mi.Emitter):
    def __init__(self, props):
        super().__init__(props)
        g = props.get
        self.number_of_elements = int(g("number_of_elements", g("num_elements_lateral", 64)))
        self.pitch = float(g("pitch", 0.0003))
        self.element_width = float(g("element_width", g("elements_width", 0.0003)))
        self.element_height = float(g("element_height", g("elements_height", 0.0005)))
        self.radius = float(g("radius", 0.0))
        self.opening_angle = float(g("opening_angle", 0.0))
        self.number_of_rays_per_element = int(g("number_of_rays_per_element", 1))
        self.number_of_total_rays = self.number_of_elements * self.number_of_rays_per_element
        self.speed_of_sound = float(g("speed_of_sound", 1540))
        self.steering_angle_min = float(g("steering_angle_min", -10.0))
        self.steering_angle_max = float(g("steering_angle_max", 10.0))
        self.element_positions, self.element_normals = self.compute_element_geometry()
        self._flags = mi.EmitterFlags.Surface | mi.EmitterFlags.SpatiallyVarying
        self._id = props.id()

    # the reference calls the misspelt name at :25; keep both
    def compute_element_geometry(self):
        n = self.number_of_elements
        if self.radius == 0.0 or not np.isfinite(self.radius) or self.opening_angle == 0.0 and self.radius > 1e3:
            x = np.linspace(-(n - 1) / 2 * self.pitch, (n - 1) / 2 * self.pitch, n)
            pos = np.stack([x, np.zeros(n), np.zeros(n)], 1)
            nrm = np.tile([0.0, 0.0, 1.0], (n, 1))
        else:
            span = np.deg2rad(self.opening_angle)
            th = np.linspace(-span / 2, span / 2, n)
            pos = np.stack([self.radius * np.sin(th), np.zeros(n), self.radius * np.cos(th)], 1)
            nrm = np.stack([np.sin(th), np.zeros(n), np.cos(th)], 1)
        nrm = nrm / np.linalg.norm(nrm, axis=1, keepdims=True)
        return mi.Point3f(pos), mi.Vector3f(nrm)

    compute_element_geoemtry = compute_element_geometry

    def sample_position(self, time, sample, active=True):
        sample1, sample2 = sample
        s2 = mi.Vector2f(sample2)
        idx = np.minimum(np.floor(np.asarray(sample1, dtype=np.float64) * self.number_of_elements),
                         self.number_of_elements - 1).astype(np.int64)
        center = np.asarray(self.element_positions)[idx]
        normal = np.asarray(self.element_normals)[idx]
        dx = (np.asarray(s2.x) - 0.5) * self.element_width
        dy = (np.asarray(s2.y) - 0.5) * self.element_height
        ps = mi.PositionSample3f()
        ps.p = mi.Point3f(center + np.stack([dx, dy, np.zeros_like(dx)], -1))
        ps.n = mi.Normal3f(normal)
        ps.time = time
        ps.delta = False
        pdf = 1 / (self.number_of_elements * self.element_width * self.element_height)
        return ps, pdf

    def sample_ray(self, time, sample1, sample2, sample3, active=True):
        ps, pdf = self.sample_position(time, (sample1, sample2), active)
        psi = np.deg2rad(self.steering_angle_min) + np.asarray(sample3, dtype=np.float64) * (
            np.deg2rad(self.steering_angle_max) - np.deg2rad(self.steering_angle_min))
        direction = mi.Vector3f(np.sin(psi), 0.0 * psi, np.cos(psi))
        time_delay = -(np.asarray(ps.p)[..., 0] * np.sin(psi)) / self.speed_of_sound
        delta_t = np.asarray(time, dtype=np.float64) + time_delay
        fd = np.maximum(0.0, np.sum(np.asarray(direction) * np.asarray(ps.n), -1))
        weight = fd / self.number_of_total_rays
        ray = mi.Ray3f(o=ps.p, d=direction, time=delta_t)
        return ray, mi.UnpolarizedSpectrum(weight)

    def sample_ray_differential(self, *args, **kwargs):
        ray, spec = self.sample_ray(*args, **kwargs)
        return ray, spec, mi.RayDifferential3f()

    def traverse(self, callback):
        for k, attr in (("number_of_elements", "number_of_elements"), ("pitch", "pitch"), ("element_width", "element_width"),
                        ("element_height", "element_height"), ("radius", "radius"), ("opening_angle", "opening_angle"),
                        ("steering_angle_min", "steering_angle_min"), ("steering_angle_max", "steering_angle_max"),
                        ("speed_of_sound", "speed_of_sound"), ("rays_per_element", "number_of_rays_per_element")):
            callback.put_parameter(k, getattr(self, attr), mi.ParamFlags.Differentiable)

    def parameters_changed(self, keys=None):
        self.element_positions, self.element_normals = self.compute_element_geometry()
        self.number_of_total_rays = self.number_of_elements * self.number_of_rays_per_element
