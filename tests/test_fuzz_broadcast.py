"""Property-based fuzzing of the broadcasting planner + reduce-to-shape: random NumPy-broadcastable
shape pairs through the tape's binary ops, forward and both operand gradients, against NumPy in fp64.
The general reduce-to-shape must agree with the reference's check_shape wherever that heuristic is right
(the oracle asserts that separately on the golden patterns)."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from conftest import assert_close
from oracle import np_oracle as O


@st.composite
def broadcastable_pair(draw):
    rank = draw(st.integers(1, 4))
    full = [draw(st.integers(1, 6)) for _ in range(rank)]

    def operand():
        drop = draw(st.integers(0, rank - 1))
        shape = full[drop:]
        return tuple(d if draw(st.booleans()) else 1 for d in shape) or (1,)
    return operand(), operand()


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(pair=broadcastable_pair(), opname=st.sampled_from(["add", "sub", "mul", "div"]), seed=st.integers(0, 2 ** 16))
def test_binary_ops_random_broadcast_shapes(backend, pair, opname, seed):
    from auto_differentiation_b200.auto_grad import tgrad as ad
    sa, sb = pair
    rng = np.random.default_rng(seed)
    a, b = rng.uniform(0.5, 1.5, sa), rng.uniform(0.5, 1.5, sb)
    out_shape = np.broadcast_shapes(sa, sb)
    g = rng.standard_normal(out_shape)
    A, B = ad.Matrix(a), ad.Matrix(b)
    node = {"add": ad.Add, "sub": ad.sub, "mul": ad.Multi, "div": ad.Divide}[opname](A, B)
    node.gradient(g)
    ref = {"add": np.add, "sub": np.subtract, "mul": np.multiply, "div": np.divide}[opname](a, b)
    if opname == "add":
        ra, rb = g, g
    elif opname == "sub":
        ra, rb = -g, g                      # the reference's sign inversion (tgrad.py:754) is part of the contract
    elif opname == "mul":
        ra, rb = g * b, g * a
    else:
        ra, rb = g / (b + 1e-8), g * (-a / (b * b + 1e-8))
    assert_close(node.val, ref, what="%s %s %s fwd" % (opname, sa, sb))
    assert_close(A.grad, O.reduce_to_shape(ra, sa), rtol=2e-5, atol=2e-6, what="%s %s %s ga" % (opname, sa, sb))
    assert_close(B.grad, O.reduce_to_shape(rb, sb), rtol=2e-5, atol=2e-6, what="%s %s %s gb" % (opname, sa, sb))


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(m=st.integers(1, 40), n=st.integers(1, 40), k=st.integers(1, 40), ta=st.booleans(), tb=st.booleans(),
       seed=st.integers(0, 2 ** 16))
def test_matmul_random_shapes_and_layouts(backend, m, n, k, ta, tb, seed):
    from auto_differentiation_b200 import device as D
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((k, m) if ta else (m, k))
    B = rng.standard_normal((n, k) if tb else (k, n))
    ref = (A.T if ta else A).astype(np.float32).astype(np.float64) @ (B.T if tb else B).astype(np.float32).astype(np.float64)
    got = D.matmul(D.asarray(A), D.asarray(B), trans_a=ta, trans_b=tb)
    assert_close(got, ref, rtol=2e-5, atol=2e-6 * np.sqrt(k), what="matmul %dx%dx%d ta=%d tb=%d" % (m, n, k, ta, tb))
