"""`utils/activation_func.py` of the reference on the B200 path."""
from .. import device as D
from ..device import DeviceArray, asarray


def multi_cross_entropy_error(y, t):
    """-sum(log(y[arange, t] + 1e-7)) / batch for probabilities y (reference :4-13); one-hot t is
    reduced with argmax(axis=1).  Returns a 0-d DeviceArray."""
    y = y if isinstance(y, DeviceArray) else asarray(y)
    if y.ndim == 1:
        y = y.reshape(1, y.size)
        t = t.reshape(1, t.size) if getattr(t, "ndim", 1) == 1 and getattr(t, "size", 0) == y.size else t
    if getattr(t, "size", None) == y.size:
        t = D.argmax_rows((t if isinstance(t, DeviceArray) else asarray(t)).reshape(y.shape))
    else:
        t = D.as_index(t).reshape(y.shape[0])
    lc, _ = D.xent_fwd(y, t, ignore_label=-(2 ** 62), eps=1e-7, denom_mode=1, input_is_probs=True,
                       want_probs=False)
    return lc[0]
