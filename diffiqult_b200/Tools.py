"""Index maps of the packed ERI vector (host-side integer helpers; reference: diffiqult/Tools.py:7-51)."""
import numpy as np


def matrix_tovector(i, j, n):
    """Row-major upper-triangle index of (i<=j) in an n x n symmetric matrix (Tools.py:17-22)."""
    return int(i * n - i * (i + 1) // 2 + j)


def vec_tomatrix(x, n):
    """Inverse of matrix_tovector (Tools.py:7-14)."""
    i = int((2 * n + 1 - np.sqrt((2 * n + 1) ** 2 - 8 * x)) // 2)
    while i > 0 and matrix_tovector(i, i, n) > x:
        i -= 1
    while i + 1 < n and matrix_tovector(i + 1, i + 1, n) <= x:
        i += 1
    return i, int(x - matrix_tovector(i, i, n) + i)


def eri_index(i, j, k, l, n):
    """Canonical position of (ij|kl) in the packed vector (Tools.py:25-51)."""
    if i > j:
        i, j = j, i
    if k > l:
        k, l = l, k
    x, y = matrix_tovector(i, j, n), matrix_tovector(k, l, n)
    if x > y:
        x, y = y, x
    return matrix_tovector(x, y, n * (n + 1) // 2)


def eri_vector_size(n):
    m = n * (n + 1) // 2
    return m * (m + 1) // 2
