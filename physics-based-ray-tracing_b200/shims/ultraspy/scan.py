import numpy as np


class GridScan:
    """ultraspy.scan.GridScan(x, z): a rectilinear imaging grid (USMain.py:204)."""

    def __init__(self, x, z, *rest):
        self.x_axis = np.asarray(getattr(x, "numpy", lambda: x)(), dtype=np.float64).reshape(-1)
        self.z_axis = np.asarray(getattr(z, "numpy", lambda: z)(), dtype=np.float64).reshape(-1)

    @property
    def shape(self):
        return (self.x_axis.size, self.z_axis.size)
