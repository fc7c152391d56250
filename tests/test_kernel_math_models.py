"""CPU models (NumPy float32, FMA emulated in float64) of the two hand-written math routines in
quadsim_kernels.cuh, checked against NumPy -- the library the reference runs on.  The GPU parity tests check the
real kernels; these pin the algorithms and their error bounds without a GPU."""
import numpy as np

f32 = np.float32


def fma(a, b, c):
    return (np.asarray(a, np.float64) * np.float64(b) + np.asarray(c, np.float64)).astype(f32)


def sincos_fast(x):
    """qs::sincos_fast: Cody-Waite by pi/2 in three parts + Cephes minimax polynomials."""
    x = x.astype(f32)
    j = np.rint((x * f32(0.636619772367581343)).astype(f32)).astype(f32)
    q = j.astype(np.int64)
    a = fma(j, f32(-1.5707962512969970703), x)
    a = fma(j, f32(-7.5497894158615963534e-08), a)
    a = fma(j, f32(-5.3903029534742383927e-15), a)
    z = (a * a).astype(f32)
    ps = fma(z, f32(-1.9515295891e-4), f32(8.3321608736e-3))
    ps = fma(ps, z, f32(-1.6666654611e-1))
    s = fma((a * z).astype(f32), ps, a)
    pc = fma(z, f32(2.443315711809948e-5), f32(-1.388731625493765e-3))
    pc = fma(pc, z, f32(4.166664568298827e-2))
    pc = fma(pc, z, f32(-0.5))
    c = fma(pc, z, f32(1.0))
    swap = (q & 1) == 1
    ss, cc = np.where(swap, c, s), np.where(swap, s, c)
    return np.where((q & 2) != 0, -ss, ss).astype(f32), np.where(((q + 1) & 2) != 0, -cc, cc).astype(f32)


def wrap_yaw(a):
    """qs::wrap_yaw: floor-quotient + one FMA, then the two wrap branches."""
    a = a.astype(f32)
    b, pi = f32(6.283185307179586), f32(3.141592653589793)
    n = np.floor((a * f32(0.15915494309189535)).astype(f32)).astype(f32)
    m = fma(-n, b, a)
    m = np.where(m < 0, (m + b).astype(f32), m)
    m = np.where(m >= b, (m - b).astype(f32), m)
    m = np.where(m > pi, (m - b).astype(f32), m)
    m = np.where(m < -pi, (m + b).astype(f32), m)
    return m.astype(f32)


def reference_wrap(yaw):
    """`3D quad race.ipynb:393-396` verbatim semantics on a float32 array."""
    yaw = yaw.astype(f32).copy()
    yaw %= 2 * np.pi
    yaw[yaw > np.pi] -= 2 * np.pi
    yaw[yaw < -np.pi] += 2 * np.pi
    return yaw


def test_sincos_fast_error_is_in_numpys_class():
    rng = np.random.default_rng(0)
    for lim in (0.4, 3.2, 30.0, 1000.0, 9.9e4):
        x = rng.uniform(-lim, lim, 400_000).astype(f32)
        s, c = sincos_fast(x)
        rs, rc = np.sin(x.astype(np.float64)), np.cos(x.astype(np.float64))
        assert np.abs(s - rs).max() < 1.0e-7 and np.abs(c - rc).max() < 1.0e-7
        ulp = lambda got, ref: (np.abs(got - ref) / np.spacing(np.abs(ref).astype(f32))).max()
        assert ulp(s, rs) <= 1.6 and ulp(c, rc) <= 1.6
        assert ulp(np.sin(x), rs) > 0.5  # NumPy's own float32 sin is not correctly rounded either


def test_wrap_yaw_is_bit_identical_to_numpy_remainder():
    rng = np.random.default_rng(1)
    a = np.concatenate([rng.uniform(-7, 7, 500_000), rng.uniform(-2000, 2000, 500_000), rng.uniform(-9e4, 9e4, 200_000),
                        np.arange(-40, 41) * np.float64(f32(6.283185307179586)), np.arange(-40, 41) * np.pi,
                        [0.0, -0.0, 1e-30, -1e-30, 3.1415927, -3.1415927, 6.2831855, -6.2831855]]).astype(f32)
    got, want = wrap_yaw(a), reference_wrap(a)
    np.testing.assert_array_equal(got, want)
