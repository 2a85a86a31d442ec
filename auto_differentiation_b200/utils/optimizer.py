"""`utils/optimizer.py` of the reference (Adam, :4-29) on the B200 path."""
from ..optim_host import AdamCore


class Adam(AdamCore):
    """Adam (http://arxiv.org/abs/1412.6980v8) as written in the reference: ONE fused multi-tensor
    kernel per update instead of a Python loop of three NumPy expressions per parameter."""
