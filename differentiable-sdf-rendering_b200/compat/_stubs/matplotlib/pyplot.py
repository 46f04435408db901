def _noop(*a, **kw):
    return None


figure = plot = xlabel = ylabel = title = savefig = close = _noop
