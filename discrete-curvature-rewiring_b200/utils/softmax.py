"""Drop-in for the reference's ``utils/softmax.py`` (utils/softmax.py:4-10), host side.

The SDRF kernel evaluates the same function on the device; this module is the host mirror used for host-side
re-decisions (a uniform within the guard band of a CDF boundary) and by callers that import it directly.
No max-shift is applied, on purpose: the reference overflows to NaN for large ``a * tau`` and ``np.random.choice``
then raises ``ValueError`` — behaviour the drop-in keeps.
"""
import numpy as np


def softmax(a, tau=1):
    if tau == float('inf'):
        one_hot = np.zeros(len(a))
        one_hot[np.argmax(a)] = 1
        return one_hot
    weights = np.exp(a * tau)
    return weights / weights.sum()
