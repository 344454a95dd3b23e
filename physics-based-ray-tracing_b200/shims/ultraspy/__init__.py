"""Stand-in for the `ultraspy` package (un-vendored, un-pinned dependency of /root/reference/USMain.py:8-10),
covering exactly what the driver calls: DelayAndSum, GridScan, build_probe -- backed by the B200 DAS kernel."""
__version__ = "0.0-prt-shim"
