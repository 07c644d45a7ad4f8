"""Chebyshev polynomial coefficients for the polynomial smoother (host-side setup).

Mirror of ``pyamg/relaxation/chebyshev.py:7-51`` (``chebyshev_polynomial_coefficients``): the degree-d
polynomial with C(0) = 1 of least maximum magnitude on [a, b] is the Chebyshev polynomial whose roots are
the Chebyshev nodes of [-1, 1] mapped affinely onto [a, b]; its monomial coefficients come from those roots
and are normalised by the value at zero.  Returned in descending order, like ``numpy.poly``.
"""
import numpy as np

__all__ = ["chebyshev_polynomial_coefficients"]


def chebyshev_polynomial_coefficients(a, b, degree):
    """Coefficients (descending) of the Chebyshev polynomial C on [a, b] with C(0) = 1."""
    if a >= b or a <= 0:
        raise ValueError(f"invalid interval [{a},{b}]")
    nodes = np.cos(np.pi * (np.arange(degree) + 0.5) / degree)        # roots of T_degree on [-1, 1]
    roots = 0.5 * (b - a) * (1 + nodes) + a                           # ... mapped onto [a, b]
    coeffs = np.poly(roots)                                           # monic polynomial with these roots
    coeffs /= np.polyval(coeffs, 0)                                   # C(0) = 1
    return coeffs
