"""Shared test helpers: random structured matrices and the tolerance of BASELINE.json's north star."""
import numpy as np

from oracle import formats as F

EPS = {np.dtype(np.float64): 2.220446049250313e-16, np.dtype(np.float32): 1.1920929e-07}


def tol(dtype, nnz_row):
    """normwise relative error bound 64*eps(T)*(nonzeros per row) -- BASELINE.json north_star."""
    return 64.0 * EPS[np.dtype(dtype)] * max(1, nnz_row)


def relerr(got, ref):
    got = np.asarray(got, dtype=np.longdouble)
    ref = np.asarray(ref, dtype=np.longdouble)
    den = np.linalg.norm(ref.ravel())
    num = np.linalg.norm((got - ref).ravel())
    if den == 0:
        return float(num)
    return float(num / den)


def random_bbb(rng, rows, cols, l, u, lam, mu, dtype=np.float64, poison=True):
    """Random band storage with NaN in every unused slot (the reference leaves them undef)."""
    shape = F.bbb_data_shape(cols, l, u, lam, mu)
    data = np.asfortranarray(rng.standard_normal(shape).astype(dtype))
    if poison and data.size:
        mask = F.bbb_valid_mask(rows, cols, l, u, lam, mu)
        data[~mask] = np.nan
    return data


def random_sky(rng, rows, cols, l, u, dtype=np.float64, scale=1.0):
    n = F.sky_numentries(rows, cols, l, u)
    return (rng.standard_normal(n) * scale).astype(dtype)


def dense_mul_ld(Adense, X):
    return (Adense.astype(np.longdouble) @ X.astype(np.longdouble))


BBB_CASES = [
    # (rows, cols, l, u, lam, mu)  -- citations: reference tests that use the same structure
    ([100] * 5, [100] * 5, 1, 1, 1, 1),                      # cfg1-like (examples/finitedifference_2d.jl)
    (list(range(1, 11)), list(range(1, 11)), 1, 1, 1, 1),    # test/test_linalg.jl:79-84
    (list(range(1, 21)), list(range(1, 21)), 2, 2, 2, 2),    # cfg3-like
    (list(range(1, 5)), list(range(1, 5)), 2, 1, 1, 2),      # test/test_bandedblockbanded.jl:98-103
    ([1, 2, 3, 4], [2, 3, 4, 5], 1, 2, 2, 1),                # test/test_adjtransblockbanded.jl:10
    (list(range(1, 6)), list(range(1, 7)), -1, 1, 0, 1),     # test/test_bandedblockbanded.jl:134-140
    (list(range(1, 6)), list(range(1, 7)), -1, 1, -1, 1),    # test/test_bandedblockbanded.jl:285-292
    (list(range(1, 6)), list(range(1, 7)), 1, -1, 1, -1),
    (list(range(1, 6)), list(range(1, 6)), -2, 1, 1, 1),     # -l > u: empty (test/test_bandedblockbanded.jl:425-434)
    (list(range(1, 6)), list(range(1, 6)), 1, 1, -2, 1),     # -lam > mu: empty
    ([2, 6], [1, 3], 0, 0, 4, 0),                            # test/test_bandedblockbanded.jl:48-60 (band cut by r[K])
    ([4, 4, 4, 4], [4, 4, 4], 1, 1, 1, 1),                   # test/test_broadcasting.jl:193 (Nr != Nc)
    ([], [], 1, 1, 1, 1),                                    # test/test_bandedblockbanded.jl:469-471
    ([7], [7], 0, 0, 2, 3),                                  # one block (OneTo axes, :493-502)
    ([3, 0, 5, 2], [3, 0, 5, 2], 1, 1, 1, 1),                # zero-size blocks
    ([300, 17, 513, 64], [290, 40, 500, 70], 1, 2, 3, 2),    # blocks spanning several tiles, ragged
    ([5] * 40, [5] * 40, 3, 4, 0, 0),                        # wide block band, diagonal blocks
    ([33] * 6, [33] * 6, 0, 0, 8, 9),                        # wide sub-band
]
