"""GPU parity: BLAS-1 kernels and pack/unpack."""
import ctypes
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [0, 1, 2, 3, 31, 257, 4097, 1_000_003])
def test_blas1(ctx, n):
    from pmg_dolfinx_b200 import api
    rng = np.random.default_rng(n)
    xh, yh = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    x, y, r = api.Vector(ctx, n), api.Vector(ctx, n), api.Vector(ctx, n)
    x.copy_from_host(xh)
    y.copy_from_host(yh)
    api.axpy(r, 0.37, x, y)
    assert np.array_equal(r.data_copy(), xh * 0.37 + yh) or np.allclose(r.data_copy(), xh * 0.37 + yh, rtol=1e-15, atol=1e-16)
    api.pointwise_mult(r, x, y)
    assert np.array_equal(r.data_copy(), xh * yh)
    api.copy(r, x)
    assert np.array_equal(r.data_copy(), xh)
    api.scale(r, -2.5)
    assert np.array_equal(r.data_copy(), xh * -2.5)
    r.set(1.25)
    assert (r.data_copy() == 1.25).all()
    d = api.inner_product(x, y)
    assert abs(d - np.dot(xh, yh)) <= 1e-13 * max(1.0, np.abs(xh * yh).sum())
    assert abs(api.norm(x) - np.linalg.norm(xh)) <= 1e-13 * max(1.0, np.linalg.norm(xh))
    if n:
        assert api.norm(x, "linf") == np.abs(xh).max()
    # in-place axpy as the solvers use it (r aliases y)
    api.axpy(y, 2.0, x, y)
    assert np.allclose(y.data_copy(), 2.0 * xh + yh, rtol=1e-15, atol=1e-16)


def test_dot_is_deterministic(ctx):
    from pmg_dolfinx_b200 import api
    n = 3_000_001
    rng = np.random.default_rng(1)
    x, y = api.Vector(ctx, n), api.Vector(ctx, n)
    x.copy_from_host(rng.normal(size=n))
    y.copy_from_host(rng.normal(size=n))
    vals = {api.inner_product(x, y) for _ in range(5)}
    assert len(vals) == 1


def test_owned_only_semantics(ctx):
    """axpy/dot/copy touch owned entries only, set/scale also the ghost block (quirk Q7)."""
    from pmg_dolfinx_b200 import api
    a, b = api.Vector(ctx, 10, 4), api.Vector(ctx, 10, 4)
    a.set(2.0)
    b.set(3.0)
    assert api.inner_product(a, b) == 60.0
    api.axpy(a, 1.0, a, b)
    h = a.data_copy()
    assert (h[:10] == 5.0).all() and (h[10:] == 2.0).all()
    api.scale(a, 2.0)
    h = a.data_copy()
    assert (h[:10] == 10.0).all() and (h[10:] == 4.0).all()


def test_pack_unpack(ctx):
    from pmg_dolfinx_b200 import api
    from pmg_dolfinx_b200.capi import lib, check, ptr
    rng = np.random.default_rng(0)
    n, m = 1000, 357
    src = rng.normal(size=n)
    idx = rng.integers(0, n, m).astype(np.int32)
    d_src, d_idx = ctx.to_device(src), ctx.to_device(idx)
    buf = ctx.zeros(m)
    check(lib.pmgx_pack(ctx.h, m, ptr(d_idx), ptr(d_src), ptr(buf)))
    ctx.sync()
    assert np.array_equal(buf.cpu().numpy(), src[idx])
    perm = rng.permutation(n)[:m].astype(np.int32)
    out = ctx.zeros(n)
    check(lib.pmgx_unpack(ctx.h, m, ptr(ctx.to_device(perm)), ptr(buf), ptr(out)))
    ctx.sync()
    ref = np.zeros(n)
    ref[perm] = src[idx]
    assert np.array_equal(out.cpu().numpy(), ref)
    out2 = ctx.zeros(n)
    check(lib.pmgx_unpack_add(ctx.h, m, ptr(d_idx), ptr(buf), ptr(out2)))
    ctx.sync()
    ref2 = np.zeros(n)
    np.add.at(ref2, idx, src[idx])
    assert np.allclose(out2.cpu().numpy(), ref2, rtol=1e-14, atol=1e-15)
