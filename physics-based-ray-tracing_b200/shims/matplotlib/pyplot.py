"""matplotlib.pyplot stand-in: every call the reference makes is accepted and ignored."""


class _Obj:
    def __getattr__(self, name):
        return lambda *a, **k: _Obj()

    def __iter__(self):
        return iter(())


def _noop(*a, **k):
    return _Obj()


figure = imshow = xlabel = ylabel = title = colorbar = tight_layout = show = hist = plot = legend = grid = savefig = _noop
subplots = lambda *a, **k: (_Obj(), _Obj())
gca = gcf = _noop
close = clf = _noop
