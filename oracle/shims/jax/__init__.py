import numpy as _np
from . import numpy
Array = _np.ndarray
class _Cfg:
    def update(self, *a, **k): pass
config = _Cfg()
def jit(f): return f
def grad(f, argnums=0):
    def g(*a, **k): raise NotImplementedError
    return g
def vmap(f, in_axes=0, out_axes=0):
    if not isinstance(in_axes, (tuple, list)): in_axes = (in_axes,)
    def w(*args):
        n = [len(a) for a, ax in zip(args, in_axes) if ax is not None][0]
        outs = [f(*[a[i] if ax is not None else a for a, ax in zip(args, in_axes)]) for i in range(n)]
        if isinstance(outs[0], tuple):
            return tuple(_np.array([o[j] for o in outs]) for j in range(len(outs[0])))
        return _np.array(outs)
    return w
