"""`utils/useful_func.py` of the reference: remove_duplicate (:24-47).  The shipped version raises
NameError for same-shape square 2-D params (missing `import numpy`); this one works."""
from ..optim_host import remove_duplicate  # noqa: F401
