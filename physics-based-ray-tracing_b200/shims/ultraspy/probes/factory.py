import numpy as np


class Probe:
    def __init__(self, geometry_type, nb_elements, pitch, central_freq, bandwidth):
        self.geometry_type, self.nb_elements, self.pitch = geometry_type, int(nb_elements), float(pitch)
        self.central_freq, self.bandwidth = float(central_freq), float(bandwidth)
        x = self.pitch * (np.arange(self.nb_elements) - (self.nb_elements - 1) / 2)
        self.geometry = np.stack([x, np.zeros_like(x), np.zeros_like(x)])

    def __str__(self):
        return f"Probe({self.geometry_type}, {self.nb_elements} elements, pitch {self.pitch * 1e3:.3f} mm, {self.central_freq / 1e6:.2f} MHz)"


def build_probe(geometry_type='linear', nb_elements=128, pitch=3e-4, central_freq=5e6, bandwidth=70, **kw):
    """ultraspy.probes.factory.build_probe (USMain.py:129-135)."""
    if geometry_type != 'linear':
        raise NotImplementedError("only linear arrays are on the reference's path")
    return Probe(geometry_type, nb_elements, pitch, central_freq, bandwidth)
