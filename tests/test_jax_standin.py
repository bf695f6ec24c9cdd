"""Semantics of the NumPy stand-in for jax that runs the reference's projector source (CPU only; does
not need /root/reference).  Each test states the JAX rule (x64 disabled) the stand-in implements."""
import numpy as np

from oracle import jax_standin as J

jax_mods = J._build_jax()
jax, jnp = jax_mods["jax"], jax_mods["jax.numpy"]


def test_results_are_float32_and_int32():
    a = jnp.asarray(np.arange(4, dtype=np.float64))
    assert a.dtype == np.float32                                  # canonical dtype of a float64 input
    assert (a * 2.5).dtype == np.float32 and (a + a).dtype == np.float32
    assert jnp.arange(5).dtype == np.int32 and jnp.mgrid[:2, :3].dtype == np.int32
    assert (jnp.mgrid[:2, :3] + 0.5).dtype == np.float32          # int array + weak Python float -> f32
    assert (a * jnp.arange(4)).dtype == np.float32                # f32 * i32 -> f32
    assert jnp.floor(a).astype(int).dtype == np.int32 and a.astype("int32").dtype == np.int32
    assert jnp.zeros((2, 2), dtype=np.float64).dtype == np.float32


def test_weak_python_scalars_do_not_widen():
    a = jnp.asarray(np.float32(1) / np.float32(3))
    want = np.float32(np.float32(1) / np.float32(3)) - np.float32(0.25)
    assert np.asarray(a - 0.25) == want
    assert np.asarray(jnp.minimum(a, 0.5)).dtype == np.float32
    assert np.asarray(jnp.where(a > 0, a, 0.0)).dtype == np.float32


def test_scatter_add_wraps_negatives_drops_out_of_bounds_and_accumulates_duplicates():
    z = jnp.zeros((4,), dtype=np.float32)
    out = z.at[jnp.asarray(np.array([0, 0, -1, 4, 7, 2]))].add(jnp.asarray(np.array([1, 2, 4, 8, 16, 32], dtype=np.float32)))
    np.testing.assert_array_equal(np.asarray(out), [3, 0, 32, 4])  # duplicates add, -1 wraps, 4 and 7 dropped
    np.testing.assert_array_equal(np.asarray(z), [0, 0, 0, 0])      # functional update
    m = jnp.zeros((2, 3), dtype=np.float32)
    rows, cols = jnp.asarray(np.array([[0, 1], [1, 2]])), jnp.asarray(np.array([[0, 2], [3, 1]]))
    out = m.at[rows, cols].add(jnp.asarray(np.ones((2, 2), np.float32)), mode="drop")
    np.testing.assert_array_equal(np.asarray(out), [[1, 0, 0], [0, 0, 1]])  # (1, 3) and (2, 1) are out of bounds
    x = jnp.mgrid[:2, :2] + 0.5
    np.testing.assert_array_equal(np.asarray(x.at[0].add(3))[0], [[3.5, 3.5], [4.5, 4.5]])


def test_gather_clamps_out_of_bounds_indices():
    y = jnp.asarray(np.arange(12, dtype=np.float32).reshape(3, 4))
    r, c = jnp.asarray(np.array([0, 2, 5, -1])), jnp.asarray(np.array([0, 7, 1, -1]))
    np.testing.assert_array_equal(np.asarray(y[r, c]), [0, 11, 9, 11])  # (2, 7) -> (2, 3); (5, 1) -> (2, 1); -1 wraps


def test_jit_canonicalises_arguments_and_keeps_static_ones():
    seen = {}

    def f(x, n, scale=1.0):
        seen["x"], seen["n"], seen["scale"] = x, n, scale
        return x

    g = jax.jit(f, static_argnames=("n",))
    g(np.arange(3, dtype=np.float64), (4, 5), scale=0.5)
    assert seen["x"].dtype == np.float32 and seen["n"] == (4, 5)
    assert np.asarray(seen["scale"]).dtype == np.float32  # a traced Python float is committed to f32


def test_vmap_map_and_scan():
    f = jax.vmap(lambda a, b: (a + b, a * b), in_axes=(0, None))
    s, p = f(jnp.asarray(np.array([1.0, 2.0, 3.0])), jnp.asarray(np.float32(2)))
    np.testing.assert_array_equal(np.asarray(s), [3, 4, 5])
    np.testing.assert_array_equal(np.asarray(p), [2, 4, 6])
    out = jax.lax.map(lambda m: m * 2, jnp.asarray(np.ones((3, 2), np.float32)), batch_size=2)
    assert np.asarray(out).shape == (3, 2) and np.asarray(out).dtype == np.float32
    carry, _ = jax.lax.scan(lambda c, xs: (c + xs[0] * xs[1], None), jnp.zeros((), np.float32),
                            (jnp.asarray(np.array([1.0, 2.0, 3.0])), jnp.asarray(np.array([1.0, 10.0, 100.0]))))
    assert float(np.asarray(carry)) == 321.0  # in order: ((0 + 1) + 20) + 300
