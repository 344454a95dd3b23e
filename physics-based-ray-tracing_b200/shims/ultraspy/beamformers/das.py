import numpy as np


class DelayAndSum:
    """ultraspy.beamformers.das.DelayAndSum as the driver uses it (USMain.py:175-208): automatic_setup(acquisition_info,
    probe), beamform(data [n_angles, n_elements, T], scan) -> [nx, nz], compute_envelope(beamformed, scan) -> [nx, nz].
    `on_gpu=False` is accepted and ignored: the work runs on the B200 (prt_das_beamform) either way."""

    def __init__(self, on_gpu=False, **kw):
        self.on_gpu = on_gpu
        self.setups = {"f_number": 1.0}
        self._info = self._probe = None
        self._last = None

    def automatic_setup(self, acquisition_info, probe):
        self._info, self._probe = acquisition_info, probe
        self.setups.update(sampling_freq=acquisition_info["sampling_freq"], sound_speed=acquisition_info["sound_speed"],
                           t0=acquisition_info.get("t0", 0) or 0)

    def update_setup(self, name, value):
        self.setups[name] = value

    def _angles_deg(self):
        """Plane-wave steering angles recovered from the per-element transmit delays (x_e sin(theta) / c,
        CustomIntegrator.py:87)."""
        d = np.asarray(self._info["delays"], dtype=np.float64)
        x = self._probe.geometry[0]
        slope = (d[:, -1] - d[:, 0]) / (x[-1] - x[0])
        return np.degrees(np.arcsin(np.clip(slope * float(self._info["sound_speed"]), -1.0, 1.0)))

    def beamform(self, data, scan):
        from prt_b200.engine import das_beamform
        data = np.asarray(data, dtype=np.float32)
        if data.ndim == 4:
            data = data[0]
        rf, env = das_beamform(data, self._angles_deg(), scan.x_axis, scan.z_axis, self.setups["sampling_freq"],
                               self.setups["sound_speed"], self._probe.pitch, t0=self.setups["t0"],
                               f_number=self.setups.get("f_number", 1.0), tx_delays=self._info["delays"])
        self._last = (rf, env)
        return rf

    def compute_envelope(self, data, scan):
        if self._last is not None and data is self._last[0]:
            return self._last[1]
        from prt_b200.engine import envelope
        return envelope(np.asarray(data, dtype=np.float32))

    def __str__(self):
        return f"DelayAndSum(B200 plane-wave DAS, f# {self.setups.get('f_number')}, fs {self.setups.get('sampling_freq')}, c {self.setups.get('sound_speed')})"
