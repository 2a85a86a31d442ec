"""Kernel-level checks through the C ABI wrappers: every code path of each kernel family (vector body,
scalar tail, unaligned base, all broadcast modes, every softmax width class, two-stage reductions, strided
GEMM operands ...) vs NumPy fp64.  `backend`: libb200ad.so on the GPU (-m gpu) / the emulator on CPU (which
validates the test logic itself)."""
import ctypes

import numpy as np
import pytest

from conftest import assert_close


def _D():
    from auto_differentiation_b200 import device
    return device


def _Lc():
    from auto_differentiation_b200 import _lib
    return _lib


def _unaligned(D, a):
    """Device copy of `a` whose base pointer is only 4-byte aligned (exercises the scalar paths)."""
    flat = D.asarray(np.concatenate([[0.0], np.asarray(a, np.float64).reshape(-1)]))
    return D.DeviceArray(flat.t[1:]).reshape(np.shape(a))


@pytest.mark.parametrize("n", [1, 3, 4, 255, 1024, 4099, 300001])
def test_unary_all_ops_vector_tail_unaligned(backend, n):
    D, L = _D(), _Lc()
    rng = np.random.default_rng(n)
    cases = {
        L.U_EXP: (rng.standard_normal(n), 0.0, np.exp), L.U_TANH: (rng.standard_normal(n) * 3, 0.0, np.tanh),
        L.U_SIGMOID: (rng.standard_normal(n) * 4, 0.0, lambda v: 1 / (1 + np.exp(-v))),
        L.U_RELU: (rng.standard_normal(n), 0.0, lambda v: np.maximum(v, 0)),
        L.U_LN: (rng.uniform(0.1, 5, n), 0.0, np.log), L.U_SQRT: (rng.uniform(0.1, 5, n), 0.0, np.sqrt),
        L.U_SQUARE: (rng.standard_normal(n), 0.0, np.square), L.U_POW: (rng.uniform(0.5, 2, n), 2.5, lambda v: v ** 2.5),
        L.U_SCALE: (rng.standard_normal(n), -1.75, lambda v: -1.75 * v),
        L.U_DIVC: (rng.standard_normal(n), 8.0, lambda v: v / 8.0), L.U_RDIVC: (rng.uniform(0.5, 2, n), 3.0, lambda v: 3.0 / v),
    }
    for op, (x, p, f) in cases.items():
        x32 = x.astype(np.float32).astype(np.float64)
        assert_close(D.unary(op, D.asarray(x), p), f(x32), what="op %d n %d" % (op, n))
        if n in (3, 4099):
            assert_close(D.unary(op, _unaligned(D, x), p), f(x32), what="op %d unaligned" % op)
    # backward spot checks (tanh uses y, ln uses x, scale uses neither)
    g = rng.standard_normal(n)
    x = rng.standard_normal(n)
    y = np.tanh(x)
    assert_close(D.unary_bwd(L.U_TANH, D.asarray(g), y=D.asarray(y)), g.astype(np.float32) * (1 - y.astype(np.float32).astype(np.float64) ** 2),
                 what="tanh bwd")
    xp = rng.uniform(0.5, 2, n)
    assert_close(D.unary_bwd(L.U_LN, D.asarray(g), x=D.asarray(xp)), g / xp, what="ln bwd")
    assert_close(D.unary_bwd(L.U_SCALE, D.asarray(g), param=0.25), g * 0.25, what="scale bwd")


@pytest.mark.parametrize("shape_a,shape_b", [((37, 64), (37, 64)), ((37, 64), (64,)), ((37, 64), (37, 1)),
                                             ((37, 64), (1, 1)), ((37, 63), (63,)), ((5, 3, 7), (3, 1)),
                                             ((4, 1, 6), (1, 5, 1)), ((130, 256), (1, 256)), ((3000, 36), (36,))])
def test_binary_modes_fwd_bwd(backend, shape_a, shape_b):
    D, L = _D(), _Lc()
    rng = np.random.default_rng(11)
    a, b = rng.uniform(0.5, 1.5, shape_a), rng.uniform(0.5, 1.5, shape_b)
    out_shape = np.broadcast_shapes(shape_a, shape_b)
    g = rng.standard_normal(out_shape)
    from oracle import np_oracle as O
    fns = {L.B_ADD: np.add, L.B_SUB: np.subtract, L.B_MUL: np.multiply, L.B_DIV: np.divide}
    for op, f in fns.items():
        for (x, y) in ((a, b), (b, a)):
            assert_close(D.binary(op, D.asarray(x), D.asarray(y)), f(x, y), what="binary %d fwd" % op)
            ga, gb = D.binary_bwd(op, D.asarray(g), D.asarray(x), D.asarray(y), True, True, True)
            if op == L.B_ADD:
                ra, rb = g, g
            elif op == L.B_SUB:
                ra, rb = -g, g
            elif op == L.B_MUL:
                ra, rb = g * y, g * x
            else:
                ra, rb = g / (y + 1e-8), g * (-x / (y * y + 1e-8))
            assert_close(ga, O.reduce_to_shape(ra, x.shape), rtol=2e-5, atol=2e-6, what="binary %d ga" % op)
            assert_close(gb, O.reduce_to_shape(rb, y.shape), rtol=2e-5, atol=2e-6, what="binary %d gb" % op)


@pytest.mark.parametrize("outer,red,inner", [(1, 7, 1), (1, 100000, 1), (3, 20000, 1), (5000, 33, 1), (70, 512, 1),
                                             (1, 5000, 10), (4, 300, 3), (1, 64, 10), (1, 4096, 256), (1, 100, 1000),
                                             (3, 17, 36), (8, 16, 32), (2, 3000, 64), (1, 1, 50), (4, 0, 6)])
def test_reduce_all_paths(backend, outer, red, inner):
    D = _D()
    rng = np.random.default_rng(outer * 131 + red * 7 + inner)
    x = rng.standard_normal((outer, red, inner))
    got = D.reduce_mid(D.asarray(x), outer, red, inner, 0.5)
    assert_close(got.reshape(outer, inner), x.astype(np.float32).astype(np.float64).sum(axis=1) * 0.5, rtol=2e-5,
                 atol=2e-6 * max(1.0, np.sqrt(max(red, 1))), what="reduce (%d,%d,%d)" % (outer, red, inner))


def test_reduce_to_shape_and_repeat(backend):
    D = _D()
    from oracle import np_oracle as O
    rng = np.random.default_rng(12)
    for gs, ts in [((8, 16, 32), (32,)), ((8, 16, 32), (1, 1, 32)), ((8, 16, 32), (8, 1, 32)), ((8, 16, 32), (1, 16, 1)),
                   ((8, 16, 32), (16, 1)), ((2, 4, 8, 8), (2, 4, 8, 1)), ((64, 10), (1, 10)), ((6, 5), ())]:
        g = rng.standard_normal(gs)
        assert_close(D.reduce_to_shape(D.asarray(g), ts), O.reduce_to_shape(g, ts), rtol=2e-5, atol=2e-6,
                     what="%s->%s" % (gs, ts))
    x = rng.standard_normal((3, 1, 5))
    assert_close(D.repeat_axis(D.asarray(x), 4, 1), np.repeat(x, 4, axis=1), what="repeat")
    assert_close(D.sum_axis(D.asarray(x), None, False), x.sum(), what="sum all")


@pytest.mark.parametrize("cols", [1, 3, 10, 31, 32, 33, 64, 100, 128, 200, 256, 500, 512, 1000, 1024, 1025, 3000])
def test_softmax_every_width_class(backend, cols):
    D = _D()
    rng = np.random.default_rng(cols)
    rows = 37
    x = rng.standard_normal((rows, cols)) * 4
    g = rng.standard_normal((rows, cols))
    x32 = x.astype(np.float32).astype(np.float64)
    e = np.exp(x32 - x32.max(axis=1, keepdims=True))
    y = e / e.sum(axis=1, keepdims=True)
    got = D.softmax_lastaxis(D.asarray(x))
    assert_close(got, y, what="softmax fwd")
    dx = y * g
    dx = dx - y * dx.sum(axis=1, keepdims=True)
    assert_close(D.softmax_bwd_lastaxis(D.asarray(g), D.asarray(y)), dx, rtol=2e-5, atol=2e-6, what="softmax bwd")
    assert_close(D.softmax_bwd_lastaxis(D.asarray(g), D.asarray(y), diag_only=True), g * y * (1 - y), rtol=2e-5,
                 atol=2e-6, what="softmax diag bwd")
    if cols in (10, 33, 1000):      # unaligned base
        assert_close(D.softmax_lastaxis(_unaligned(D, x)), y, what="softmax unaligned")
    # fused in_scale (attention) and the fused cross-entropy
    e2 = np.exp(0.125 * x32 - (0.125 * x32).max(axis=1, keepdims=True))
    assert_close(D.softmax_lastaxis(D.asarray(x), in_scale=0.125), e2 / e2.sum(axis=1, keepdims=True), what="scaled")
    t = rng.integers(0, cols, rows)
    if rows > 3:
        t[2] = -1
    mask = t != -1
    lc, probs = D.xent_fwd(D.asarray(x), D.as_index(t))
    ref_loss = -(np.log(y[np.arange(rows), t]) * mask).sum() / mask.sum()
    assert abs(float(lc[0]) - ref_loss) < 2e-5 * max(1.0, abs(ref_loss))
    assert_close(probs, y, what="xent probs")
    ref_dx = y.copy()
    ref_dx[np.arange(rows), t] -= 1
    ref_dx = ref_dx / mask.sum() * mask[:, None]
    assert_close(D.xent_bwd(probs, D.as_index(t), lc), ref_dx, what="xent bwd")


def test_argmax_bit_exact(backend):
    D = _D()
    rng = np.random.default_rng(13)
    for n in (1, 5, 1023, 1025, 250000, 3000001):
        x = rng.integers(0, 50, n).astype(np.float64)          # many ties -> first occurrence matters
        assert int(D.argext_flat(D.asarray(x), False)) == int(np.argmax(x))
        assert int(D.argext_flat(D.asarray(x), True)) == int(np.argmin(x))
    x = rng.integers(0, 4, (77, 13)).astype(np.float64)
    np.testing.assert_array_equal(np.asarray(D.argmax_rows(D.asarray(x))), x.argmax(axis=1))
    arr = D.asarray(x)
    assert arr.argmax(axis=1).get().dtype == np.int64


def test_embedding_exact_and_deterministic(backend):
    D = _D()
    rng = np.random.default_rng(14)
    for vocab, dim, n in ((9, 4, 15), (1000, 256, 1024), (50, 33, 700)):
        W = rng.standard_normal((vocab, dim))
        idx = rng.integers(0, vocab, n)
        idx[: n // 4] = idx[0]                                   # heavy collisions
        out = D.embedding_gather(D.asarray(W), D.as_index(idx))
        np.testing.assert_array_equal(np.asarray(out), W.astype(np.float32)[idx])
        d = rng.standard_normal((n, dim))
        ref = np.zeros((vocab, dim)); np.add.at(ref, idx, d.astype(np.float32).astype(np.float64))
        dw1, dw2 = D.empty((vocab, dim)), D.empty((vocab, dim))
        D.embedding_scatter_add_(D.asarray(d), D.as_index(idx), dw1)
        D.embedding_scatter_add_(D.asarray(d), D.as_index(idx), dw2)
        assert_close(dw1, ref, rtol=2e-5, atol=2e-6, what="scatter-add")
        np.testing.assert_array_equal(np.asarray(dw1), np.asarray(dw2))   # bit-reproducible
    W = D.asarray(rng.standard_normal((6, 3)))
    np.testing.assert_array_equal(np.asarray(D.embedding_gather(W, D.as_index(np.array([-1, -6])))),
                                  np.asarray(W)[[-1, -6]])      # negative indices wrap like NumPy


def test_strided_copy_and_views(backend):
    D = _D()
    rng = np.random.default_rng(15)
    x = rng.standard_normal((3, 8, 4, 16))
    a = D.asarray(x)
    for perm in ((0, 2, 1, 3), (0, 1, 3, 2), (3, 2, 1, 0), (1, 0, 2, 3), (0, 3, 1, 2)):
        np.testing.assert_array_equal(np.asarray(a.transpose(*perm).contiguous()), x.astype(np.float32).transpose(perm))
    y = rng.standard_normal((37, 53))
    np.testing.assert_array_equal(np.asarray(D.asarray(y).T.contiguous()), y.astype(np.float32).T)
    np.testing.assert_array_equal(np.asarray(D.asarray(y)[3:20, 5:50:3].contiguous()), y.astype(np.float32)[3:20, 5:50:3])
    z = D.zeros((6, 8))
    z[1:4] = D.asarray(np.ones((3, 8)))
    z[:, 7] = 5.0
    ref = np.zeros((6, 8)); ref[1:4] = 1; ref[:, 7] = 5
    np.testing.assert_array_equal(np.asarray(z), ref)
    np.testing.assert_array_equal(np.asarray(D.broadcast_to(D.asarray(y[:1]), (4, 53))), np.repeat(y[:1], 4, 0).astype(np.float32))


GEMM_CASES = [  # (M, N, K)
    (1, 1, 1), (7, 9, 5), (64, 10, 256), (64, 256, 784), (129, 130, 131), (256, 256, 64), (300, 17, 1000), (16, 1000, 40)]


@pytest.mark.parametrize("M,N,K", GEMM_CASES)
def test_gemm_fp32_all_layouts(backend, M, N, K):
    D, L = _D(), _Lc()
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A, B = rng.standard_normal((M, K)), rng.standard_normal((K, N))
    ref = A.astype(np.float32).astype(np.float64) @ B.astype(np.float32).astype(np.float64)
    tol = dict(rtol=2e-5, atol=2e-6 * np.sqrt(K))
    a, b = D.asarray(A), D.asarray(B)
    at, bt = D.asarray(A.T.copy()), D.asarray(B.T.copy())
    assert_close(D.matmul(a, b, precision="fp32"), ref, what="NN", **tol)
    assert_close(D.matmul(at, b, trans_a=True, precision="fp32"), ref, what="TN", **tol)
    assert_close(D.matmul(a, bt, trans_b=True, precision="fp32"), ref, what="NT", **tol)
    assert_close(D.matmul(at.T, bt.T, precision="fp32"), ref, what="TT via views", **tol)
    bias = rng.standard_normal(N)
    assert_close(D.matmul(a, b, bias=D.asarray(bias), act=L.ACT_TANH, precision="fp32"), np.tanh(ref + bias), what="bias+tanh", **tol)
    assert_close(D.matmul(a, b, bias=D.asarray(bias), act=L.ACT_RELU, precision="fp32"), np.maximum(ref + bias, 0), what="bias+relu", **tol)
    c0 = rng.standard_normal((M, N))
    c = D.asarray(c0)
    D.matmul(a, b, out=c, alpha=0.5, beta=2.0, precision="fp32")
    assert_close(c, 0.5 * ref + 2.0 * c0.astype(np.float32), what="alpha/beta", **tol)


def test_gemm_default_path_fuzz(backend):
    """The DEFAULT precision (fp32-faithful tensor-core path where the shape allows: generation-3 pair kernel for
    N, M > 128, 1-CTA kernel otherwise, skinny / SIMT kernels for what TMA cannot take) over seeded random shapes,
    layouts, batch axes and epilogues, vs fp64.  Covers 256-row / 256-column tile tails, split-K, K tails (K % 32),
    TMA-store clipping, strided head-split batch views."""
    D, L = _D(), _Lc()
    rng = np.random.default_rng(20250)
    cases = [(257, 259, 96), (384, 640, 512), (130, 1000, 72), (1024, 136, 200), (512, 512, 4160), (300, 300, 33 * 32 + 8),
             (129, 257, 16), (2048, 264, 64)]
    for _ in range(10):
        cases.append((int(rng.integers(1, 700)), int(rng.integers(1, 700)), int(rng.integers(1, 1500))))
    for idx, (M, N, K) in enumerate(cases):
        ta, tb = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        A = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
        B = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
        ref = (A.T if ta else A).astype(np.float64) @ (B.T if tb else B).astype(np.float64)
        kind = idx % 4
        kw, c0 = {}, None
        if kind == 1:
            bias = rng.standard_normal(N).astype(np.float32)
            kw = dict(bias=D.asarray(bias), act=L.ACT_TANH)
            ref = np.tanh(ref + bias)
        elif kind == 2:
            c0 = rng.standard_normal((M, N)).astype(np.float32)
            kw = dict(out=D.asarray(c0.copy()), alpha=0.5, beta=1.0)
            ref = 0.5 * ref + c0
        elif kind == 3:
            kw = dict(alpha=-2.0)
            ref = -2.0 * ref
        got = D.matmul(D.asarray(A), D.asarray(B), trans_a=ta, trans_b=tb, **kw)
        assert_close(got, ref, rtol=2e-5, atol=2e-6 * np.sqrt(K), what="case %d M=%d N=%d K=%d ta=%d tb=%d kind=%d" % (idx, M, N, K, ta, tb, kind))
    # batched head-split views (the MultiHeadAttention projections / unfused score products)
    Bt, h, S, dk = 3, 4, 160, 72
    x = rng.standard_normal((Bt, S, h * dk)).astype(np.float32)
    y = rng.standard_normal((Bt, S, h * dk)).astype(np.float32)
    xh = D.asarray(x).reshape(Bt, S, h, dk).transpose(0, 2, 1, 3)
    yh = D.asarray(y).reshape(Bt, S, h, dk).transpose(0, 2, 1, 3)
    got = D.matmul(xh, yh, trans_b=True)
    xr = x.reshape(Bt, S, h, dk).transpose(0, 2, 1, 3).astype(np.float64)
    yr = y.reshape(Bt, S, h, dk).transpose(0, 2, 1, 3).astype(np.float64)
    assert_close(got, xr @ yr.transpose(0, 1, 3, 2), rtol=2e-5, atol=2e-6 * np.sqrt(dk), what="batched head views")


@pytest.mark.parametrize("M,N,K,ta,tb", [(256, 512, 256, False, False), (200, 300, 1000, False, True), (64, 256, 784, True, False),
                                         (512, 512, 8192, True, True), (300, 72, 96, False, False)])
def test_gemm_bf16_variant_stated_tolerance(backend, M, N, K, ta, tb):
    """bf16 tensor-core variant (B200AD_GEMM_BF16): operands rounded to bf16, fp32 accumulation.  Stated tolerance:
    |err| <= 2^-7 * (|A| @ |B|) elementwise (each operand carries a 2^-9 relative rounding error), and the result must
    match the same product evaluated in fp64 on bf16-ROUNDED inputs to fp32-accumulation accuracy."""
    D = _D()
    rng = np.random.default_rng(M + N + K)
    A = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    B = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    got = np.asarray(D.matmul(D.asarray(A), D.asarray(B), trans_a=ta, trans_b=tb, precision="bf16", bias=D.asarray(bias)),
                     dtype=np.float64)
    Aop = (A.T if ta else A).astype(np.float64)
    Bop = (B.T if tb else B).astype(np.float64)
    exact = Aop @ Bop + bias
    bound = 2.0 ** -7 * (np.abs(Aop) @ np.abs(Bop)) + 1e-6
    assert np.all(np.abs(got - exact) <= bound), float(np.max(np.abs(got - exact) / bound))
    if backend == "cuda":
        def bf16_round(x):                       # round-to-nearest-even to 8 significant bits
            u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
            u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
            return u.astype(np.uint32).view(np.float32).astype(np.float64)
        ref = bf16_round(A.T if ta else A) @ bf16_round(B.T if tb else B) + bias
        assert_close(got, ref, rtol=2e-5, atol=2e-5 * np.sqrt(K), what="bf16 GEMM vs fp64 on bf16-rounded inputs")


F16X3_CASES = [(1, 1, 1), (7, 9, 5), (256, 512, 256), (257, 259, 96), (300, 300, 33 * 32 + 8), (130, 1000, 72), (2048, 264, 64),
               (512, 512, 8192), (640, 384, 1031)]


@pytest.mark.parametrize("M,N,K", F16X3_CASES)
def test_gemm_f16x3_all_layouts_and_epilogues(backend, M, N, K):
    """fp16 hi/lo three-product variant (B200AD_GEMM_F16X3, csrc/gemm_h16.cu) held to the SAME tolerance as the default
    fp32-faithful path: all four operand layouts (row split / column-maximum + transposing split pre-passes), tile and K
    tails, split-K (K = 8192), K % 8 != 0 (padded planes), fused epilogues."""
    D, L = _D(), _Lc()
    rng = np.random.default_rng(M * 5 + N * 3 + K)
    A, B = rng.standard_normal((M, K)).astype(np.float32), rng.standard_normal((K, N)).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64)
    tol = dict(rtol=2e-5, atol=2e-6 * np.sqrt(K))
    a, b = D.asarray(A), D.asarray(B)
    at, bt = D.asarray(A.T.copy()), D.asarray(B.T.copy())
    assert_close(D.matmul(a, b, precision="f16x3"), ref, what="NN", **tol)
    assert_close(D.matmul(at, b, trans_a=True, precision="f16x3"), ref, what="TN", **tol)
    assert_close(D.matmul(a, bt, trans_b=True, precision="f16x3"), ref, what="NT", **tol)
    assert_close(D.matmul(at, bt, trans_a=True, trans_b=True, precision="f16x3"), ref, what="TT", **tol)
    bias = rng.standard_normal(N).astype(np.float32)
    assert_close(D.matmul(a, b, bias=D.asarray(bias), act=L.ACT_TANH, precision="f16x3"), np.tanh(ref + bias), what="bias+tanh", **tol)
    c0 = rng.standard_normal((M, N)).astype(np.float32)
    c = D.asarray(c0.copy())
    D.matmul(a, b, out=c, alpha=0.5, beta=2.0, precision="f16x3")
    assert_close(c, 0.5 * ref + 2.0 * c0, what="alpha/beta", **tol)


def test_gemm_f16x3_row_scales_and_dynamic_range(backend):
    """The power-of-two row scales keep the split fp32-faithful when operand rows span ten decades (1e-6 .. 1e4) and for
    gradient-sized rows (1e-8): every output element is held to an ELEMENTWISE bound relative to sum_k |a||b| (a max-norm
    tolerance would hide a lost small row).  Also an all-zero row and a batched head-split view."""
    D = _D()
    rng = np.random.default_rng(77)
    M, N, K = 384, 320, 520
    A = rng.standard_normal((M, K)) * 10.0 ** rng.uniform(-6, 4, size=(M, 1))
    B = rng.standard_normal((K, N)) * 10.0 ** rng.uniform(-6, 4, size=(1, N))
    A[5] = 0.0
    A[7] *= 1e-8 / np.abs(A[7]).max()
    A, B = A.astype(np.float32), B.astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64)
    bound = (np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64))
    for ta, tb in ((False, False), (True, True)):
        a = D.asarray(A.T.copy() if ta else A)
        b = D.asarray(B.T.copy() if tb else B)
        got = np.asarray(D.matmul(a, b, trans_a=ta, trans_b=tb, precision="f16x3"), dtype=np.float64)
        ratio = np.abs(got - ref) / (4e-6 * bound + 1e-38)
        from conftest import record_parity
        i = np.unravel_index(np.argmax(ratio), ratio.shape)
        record_parity("f16x3 elementwise ta=%d tb=%d" % (ta, tb), abs(got[i] - ref[i]), 4e-6 * bound[i] + 1e-38, ratio.max(), 4e-6, 0.0, ratio.size)
        assert ratio.max() <= 1.0, "f16x3 elementwise error %.3g of the 4e-6 |A||B| bound (ta=%d tb=%d)" % (ratio.max(), ta, tb)
        assert np.all(got[5] == 0.0)
    Bt, h, S, dk = 3, 4, 160, 72
    x = rng.standard_normal((Bt, S, h * dk)).astype(np.float32)
    y = rng.standard_normal((Bt, S, h * dk)).astype(np.float32)
    xh = D.asarray(x).reshape(Bt, S, h, dk).transpose(0, 2, 1, 3)
    yh = D.asarray(y).reshape(Bt, S, h, dk).transpose(0, 2, 1, 3)
    got = D.matmul(xh, yh, trans_b=True, precision="f16x3")
    xr = x.reshape(Bt, S, h, dk).transpose(0, 2, 1, 3).astype(np.float64)
    yr = y.reshape(Bt, S, h, dk).transpose(0, 2, 1, 3).astype(np.float64)
    assert_close(got, xr @ yr.transpose(0, 1, 3, 2), rtol=2e-5, atol=2e-6 * np.sqrt(dk), what="f16x3 batched head views")


def test_gemm_f16x3_partial_column_tiles_repeatable(backend):
    """Partial column tiles (N % 32 != 0) make a fold warp issue a single output group per tile: consecutive stores of a warp
    must still alternate between its two staging tiles (a version that picked the tile by the group index let the next fill
    overwrite a tile the TMA engine was still reading — an intermittent wrong block).  Many row tiles per CTA, repeated."""
    D = _D()
    rng = np.random.default_rng(91)
    for (M, N, K) in ((6000, 260, 128), (5000, 530, 96)):
        A = rng.standard_normal((M, K)).astype(np.float32)
        B = rng.standard_normal((K, N)).astype(np.float32)
        ref = A.astype(np.float64) @ B.astype(np.float64)
        a, b = D.asarray(A), D.asarray(B)
        first = None
        for rep in range(6):
            got = np.asarray(D.matmul(a, b, precision="f16x3"))
            if first is None:
                first = got
                assert_close(got, ref, rtol=2e-5, atol=2e-6 * np.sqrt(K), what="M=%d N=%d K=%d" % (M, N, K))
            else:
                np.testing.assert_array_equal(got, first)


def test_gemm_shared_a_and_auto_dispatch(backend):
    """b200ad_gemm_shared_a (one operand split for several products: the Q / K / V projections) equals separate calls
    bit for bit, and the AUTO dispatch (b200ad_gemm_set_f16x3_auto: never / by the device-time model / always) stays inside
    the default tolerance with F16X3 forced on and off."""
    D = _D()
    rng = np.random.default_rng(5)
    M, N, K = 16384 + 200, 512, 512
    A = rng.standard_normal((M, K)).astype(np.float32)
    Ws = [rng.standard_normal((K, N)).astype(np.float32) / np.sqrt(K) for _ in range(3)]
    a, ws = D.asarray(A), [D.asarray(w) for w in Ws]
    refs = [A.astype(np.float64) @ w.astype(np.float64) for w in Ws]
    outs = D.matmul_shared_a(a, ws, precision="f16x3")
    for o, w, r in zip(outs, ws, refs):
        assert_close(o, r, rtol=2e-5, atol=2e-6 * np.sqrt(K), what="shared-A product")
        np.testing.assert_array_equal(np.asarray(o), np.asarray(D.matmul(a, w, precision="f16x3")))
    prev = D.set_f16x3_auto(2)
    try:
        on = [np.asarray(o) for o in D.matmul_shared_a(a, ws)] + [np.asarray(D.matmul(a, ws[0]))]
        D.set_f16x3_auto(0)
        off = [np.asarray(o) for o in D.matmul_shared_a(a, ws)] + [np.asarray(D.matmul(a, ws[0]))]
    finally:
        D.set_f16x3_auto(prev)
    for o, f, r in zip(on, off, refs + [refs[0]]):
        assert_close(o, r, rtol=2e-5, atol=2e-6 * np.sqrt(K), what="auto, f16x3 on")
        assert_close(f, r, rtol=2e-5, atol=2e-6 * np.sqrt(K), what="auto, f16x3 off")
    # different geometries fall back to a loop
    outs = D.matmul_shared_a(a[:100], [ws[0], D.asarray(Ws[1][:, :64].copy())])
    assert_close(outs[1], A[:100].astype(np.float64) @ Ws[1][:, :64].astype(np.float64), rtol=2e-5, atol=2e-6 * np.sqrt(K), what="fallback loop")


@pytest.mark.parametrize("M,N,K,tb,act,beta", [(64, 1024, 1024, False, 1, 1.0), (64, 1024, 1024, True, 0, 0.0), (17, 2048, 1024, False, 0, 1.0),
                                               (1, 256, 256, True, 2, 0.0), (64, 512, 2048, False, 3, 0.5), (64, 1024, 4096, True, 0, 0.0),
                                               (64, 264, 1024, False, 1, 1.0), (40, 328, 992, True, 0, 0.0)])
def test_gemm_skinny_recurrent_step(backend, M, N, K, tb, act, beta):
    """The M <= 64 recurrent-step kernel (fused beta * C + bias + activation epilogue), both B layouts, strided
    output rows (a time slice of an (N, T, H) tensor), vs fp64."""
    D, L = _D(), _Lc()
    rng = np.random.default_rng(M * 7 + N + K)
    A = (rng.standard_normal((M, K)) / np.sqrt(K)).astype(np.float32)
    B = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    C0 = rng.standard_normal((M, 3, N)).astype(np.float32)
    out_full = D.asarray(C0.copy())                           # (the CPU emulation would alias the host array)
    out = out_full[:, 1, :]                                   # row stride 3N
    D.matmul(D.asarray(A), D.asarray(B), trans_b=tb, out=out, beta=beta, bias=D.asarray(bias), act=act)
    pre = A.astype(np.float64) @ (B.T if tb else B).astype(np.float64) + beta * C0[:, 1, :].astype(np.float64) + bias
    ref = {0: lambda v: v, 1: np.tanh, 2: lambda v: 1 / (1 + np.exp(-v)), 3: lambda v: np.maximum(v, 0)}[act](pre)
    got = np.asarray(out_full)
    # absolute part relative to the magnitude of the accumulated pre-activation (every activation here is 1-Lipschitz):
    # the small weight matrices run on the exact-fp32 kernels, N * K > 512 Ki on the 3xTF32 tensor-core path
    scale = max(1.0, float(np.max(np.abs(pre)))) / max(1.0, float(np.max(np.abs(ref))))
    assert_close(got[:, 1, :], ref, rtol=2e-5, atol=2e-6 * scale, what="skinny GEMM")
    assert np.array_equal(got[:, 0, :], C0[:, 0, :]) and np.array_equal(got[:, 2, :], C0[:, 2, :])


@pytest.mark.parametrize("M,N,K,prec", [(300, 260, 96, None), (300, 260, 96, "bf16"), (130, 520, 40, None), (100, 100, 64, None),
                                        (33, 512, 256, None), (513, 300, 4128, None), (300, 260, 96, "f16x3"),
                                        (513, 300, 4128, "f16x3"), (2304, 264, 72, None)])
@pytest.mark.parametrize("col_off", [4, 3])
def test_gemm_output_guards(backend, M, N, K, prec, col_off):
    """C is a window of a larger buffer filled with a sentinel: rows above / below and columns left / right of the window
    must stay untouched (TMA-store clipping for 16 B-aligned windows, guarded direct stores otherwise), for the pair,
    1-CTA, skinny and bf16 kernels and for split-K."""
    D = _D()
    rng = np.random.default_rng(M + N + K + col_off)
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = rng.standard_normal((K, N)).astype(np.float32)
    pad = 8
    big = np.full((M + 2, N + 2 * pad), 7.25, dtype=np.float32)
    buf = D.asarray(big.copy())
    out = buf[1:1 + M, col_off:col_off + N]
    kw = {} if prec is None else dict(precision=prec)
    D.matmul(D.asarray(A), D.asarray(B), out=out, **kw)
    got = np.asarray(buf)
    ref = A.astype(np.float64) @ B.astype(np.float64)
    if prec == "bf16":
        assert np.all(np.abs(got[1:1 + M, col_off:col_off + N] - ref) <= 2.0 ** -7 * (np.abs(A).astype(np.float64) @ np.abs(B)) + 1e-6)
    else:
        assert_close(got[1:1 + M, col_off:col_off + N], ref, rtol=2e-5, atol=2e-6 * np.sqrt(K), what="window")
    mask = np.ones_like(big, dtype=bool)
    mask[1:1 + M, col_off:col_off + N] = False
    assert np.all(got[mask] == 7.25), "GEMM wrote outside its output window"


def test_gemm_batched_strided_views(backend):
    D = _D()
    rng = np.random.default_rng(16)
    B, h, S, dk = 2, 3, 20, 8
    Q, K, V = (rng.standard_normal((B, S, h * dk)) for _ in range(3))
    q, k, v = (D.asarray(x).reshape(B, S, h, dk).transpose(0, 2, 1, 3) for x in (Q, K, V))    # strided views
    Qh, Kh, Vh = (x.astype(np.float32).astype(np.float64).reshape(B, S, h, dk).transpose(0, 2, 1, 3) for x in (Q, K, V))
    s = D.matmul(q, k, trans_b=True, precision="fp32")
    assert_close(s, Qh @ Kh.transpose(0, 1, 3, 2), rtol=2e-5, atol=1e-5, what="QK^T on views")
    o = D.empty((B, S, h, dk))
    D.matmul(s, v, out=o.transpose(0, 2, 1, 3), precision="fp32")            # written in the merged layout
    ref = ((Qh @ Kh.transpose(0, 1, 3, 2)) @ Vh).transpose(0, 2, 1, 3)
    assert_close(o, ref, rtol=5e-5, atol=5e-5, what="PV into merged layout")
    # broadcast batch (3-D @ 2-D) and >2 batch axes
    x3, w = rng.standard_normal((4, 6, 5)), rng.standard_normal((5, 7))
    assert_close(D.matmul(D.asarray(x3), D.asarray(w), precision="fp32"), x3 @ w, rtol=2e-5, atol=1e-5, what="3D@2D")
    x5 = rng.standard_normal((2, 3, 2, 4, 5)); y5 = rng.standard_normal((2, 3, 2, 5, 3))
    assert_close(D.matmul(D.asarray(x5).transpose(1, 0, 2, 3, 4), D.asarray(y5).transpose(1, 0, 2, 3, 4), precision="fp32"),
                 (x5 @ y5).transpose(1, 0, 2, 3, 4), rtol=2e-5, atol=1e-5, what="5-D permuted batch")
    with pytest.raises(ValueError):
        D.matmul(D.asarray(x3), D.asarray(rng.standard_normal((4, 7))))


def test_cells_with_row_strides(backend):
    """lstm/rnn cell kernels address time slices of (N,T,X) tensors in place through row strides."""
    D, L = _D(), _Lc()
    from oracle import np_oracle as O
    rng = np.random.default_rng(17)
    for N, H in ((3, 6), (5, 64), (2, 7)):
        T = 4
        A = rng.standard_normal((N, T, 4 * H)); C = rng.standard_normal((N, T, H))
        a, c = D.asarray(A), D.asarray(C)
        gates, cn, hn = D.empty((N, T, 4 * H)), D.empty((N, T, H)), D.empty((N, T, H))
        lib = L.load()
        t = 2
        L.check(lib.b200ad_lstm_cell_fwd(a[:, t].ptr, T * 4 * H, c[:, t].ptr, T * H, gates[:, t].ptr, T * 4 * H,
                                         cn[:, t].ptr, T * H, hn[:, t].ptr, T * H, N, H, D.stream_ptr()))
        A32, C32 = A.astype(np.float32).astype(np.float64), C.astype(np.float32).astype(np.float64)
        sg = lambda v: 1 / (1 + np.exp(-v))
        f, g, i, o = sg(A32[:, t, :H]), np.tanh(A32[:, t, H:2 * H]), sg(A32[:, t, 2 * H:3 * H]), sg(A32[:, t, 3 * H:])
        cref = f * C32[:, t] + g * i
        assert_close(cn[:, t].contiguous(), cref, what="cell c'")
        assert_close(hn[:, t].contiguous(), o * np.tanh(cref), what="cell h'")
        dh, dc = rng.standard_normal((N, T, H)), rng.standard_normal((N, H))
        dA, dcp = D.empty((N, T, 4 * H)), D.empty((N, H))
        dh_d, dc_d = D.asarray(dh), D.asarray(dc)          # keep the buffers alive across the raw call
        L.check(lib.b200ad_lstm_cell_bwd(dh_d[:, t].ptr, T * H, None, 0, dc_d.ptr, gates[:, t].ptr,
                                         T * 4 * H, c[:, t].ptr, T * H, cn[:, t].ptr, T * H, hn[:, t].ptr, T * H,
                                         dA[:, t].ptr, T * 4 * H, dcp.ptr, N, H, D.stream_ptr()))
        cache = (cref, o * np.tanh(cref), np.zeros((N, H)), C32[:, t], np.zeros((N, 1)), f, g, i, o)
        _, _, dc_ref, _, _, _ = O.lstm_backward(cache, np.zeros((1, 4 * H)), np.zeros((H, 4 * H)),
                                                dh[:, t].astype(np.float32).astype(np.float64), dc.astype(np.float32).astype(np.float64))
        assert_close(dcp, dc_ref, rtol=2e-5, atol=2e-6, what="cell dc")


def test_multi_tensor_optimizer_odd_sizes(backend):
    D = _D()
    from auto_differentiation_b200.optim_host import AdamCore, SGDCore, clip_grads
    from oracle import np_oracle as O
    rng = np.random.default_rng(18)
    shapes = [(1,), (3,), (4096,), (4097,), (33, 77), (10000,)]
    p0 = [rng.standard_normal(s) for s in shapes]
    for Core, Ref in ((AdamCore, O.Adam), (SGDCore, O.SGD)):
        params = [D.asarray(p) for p in p0]
        ref = [p.astype(np.float32).astype(np.float64) for p in p0]
        opt, ropt = Core(), Ref()
        for step in range(4):
            gs = [rng.standard_normal(s) for s in shapes]
            opt.update(params, [D.asarray(g) for g in gs])
            ropt.update(ref, [g.astype(np.float32).astype(np.float64) for g in gs])
        for p, r in zip(params, ref):
            assert_close(p, r, rtol=2e-5, atol=2e-6, what=Core.__name__)
    gs = [rng.standard_normal(s) for s in shapes]
    dev = [D.asarray(g) for g in gs]
    clip_grads(dev, 2.0)
    ref = [g.astype(np.float32).astype(np.float64) for g in gs]
    O.clip_grads(ref, 2.0)
    for a, r in zip(dev, ref):
        assert_close(a, r, rtol=2e-5, atol=2e-6, what="clip")


@pytest.mark.gpu
def test_large_size_properties():
    """BASELINE-scale sizes (config 2: up to 1e8 elements here) through size-independent properties."""
    import torch
    from auto_differentiation_b200 import _lib as L, device as D
    n = 100_000_000 // 512 * 512
    rows = n // 256
    x = D.DeviceArray(torch.empty(n, device="cuda").uniform_(0.5, 1.5))      # RNG plumbing only
    # exp/ln round trip, sigmoid symmetry, relu idempotence
    rt = D.unary(L.U_LN, D.unary(L.U_EXP, x))
    assert float(D.sum_axis(D.unary(L.U_SQUARE, rt - x), None)) / n < 1e-12
    s = D.unary(L.U_SIGMOID, x) + D.unary(L.U_SIGMOID, D.unary(L.U_NEG, x))
    assert abs(float(D.sum_axis(s, None)) / n - 1.0) < 1e-5
    r = D.unary(L.U_RELU, D.unary(L.U_ADDC, x, -1.0))
    assert float(D.sum_axis(D.unary(L.U_SQUARE, D.unary(L.U_RELU, r) - r), None)) == 0.0
    # broadcast-add backward: column sums of g == ones^T g (checksum of checksums), fused and unfused agree
    g = x.reshape(rows, 256)
    b = D.zeros((1, 256))
    ga, gb = D.binary_bwd(L.B_MUL, g, g, D.full((1, 256), 2.0))
    total = float(D.sum_axis(gb, None))                      # sum_c sum_r g*a  == sum(g*g)
    assert abs(total - float(D.sum_axis(D.unary(L.U_SQUARE, g), None))) / total < 1e-5
    gb2 = D.reduce_mid(D.unary(L.U_SQUARE, g), 1, rows, 256)
    assert_close(gb.reshape(-1), np.asarray(gb2), rtol=1e-5, atol=1e-6, what="fused colsum == unfused")
    # softmax rows sum to one; linearity of the GEMM in its first operand
    sm = D.softmax_lastaxis(x.reshape(n // 512, 512))
    assert abs(float(D.sum_axis(sm, None)) - n // 512) < 1e-5 * (n // 512)       # fp32 sum of ~2e5 row sums


@pytest.mark.gpu
def test_cluster_kernels_bitwise_repeatable():
    """The CTA-pair kernels hand data between CTAs through mbarriers whose arrives are `.release.cta` (the
    `.release.cluster` form costs a MEMBAR.ALL.GPU per call): a missed ordering would show up as run-to-run differences.
    Thirty repetitions each of the gen3 projection / split-K weight-gradient GEMMs, the single-row-tile split-K step and
    the pair attention passes must be bit-identical, with other kernels racing on the GPU in between."""
    import torch
    D, L = _D(), _Lc()
    g = torch.Generator(device="cuda").manual_seed(7)
    A = D.DeviceArray(torch.randn((4096, 512), generator=g, device="cuda"))
    W = D.DeviceArray(torch.randn((512, 512), generator=g, device="cuda"))
    G = D.DeviceArray(torch.randn((4096, 512), generator=g, device="cuda"))
    h = D.DeviceArray(torch.randn((64, 1024), generator=g, device="cuda") / 32)
    Wh = D.DeviceArray(torch.randn((1024, 1024), generator=g, device="cuda"))
    q, k, v, do = (D.DeviceArray(torch.randn((4, 512, 8, 64), generator=g, device="cuda")).transpose(0, 2, 1, 3) for _ in range(4))
    noise = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")

    def run():
        y = D.matmul(A, W)                                   # gen3, projection shape
        dw = D.matmul(A, G, trans_a=True)                    # gen3, split-K
        s = D.matmul(h, Wh, bias=None, act=L.ACT_TANH)       # 1-CTA kernel, 64-column tiles, 8-way split-K
        o, lse = D.attn_fwd(q, k, v, 0.125)                  # pair passes (512 rows = 4 row tiles)
        dq, dk, dv = D.attn_bwd(q, k, v, o, do, lse, 0.125)
        return [np.asarray(t.contiguous()) for t in (y, dw, s, o, lse, dq, dk, dv)]

    first = run()
    for rep in range(30):
        noise.fill_(rep)                                     # unrelated traffic between repetitions
        for a, b in zip(first, run()):
            assert np.array_equal(a, b), "repetition %d differs" % rep


def test_uniform_stream(backend):
    """The dropout draws come from the library's own Philox stream (b200ad_uniform): values in [0, 1) on a 2^-24 grid,
    uniform moments, reproducible from the seed, and consecutive calls continue the stream instead of repeating it."""
    from auto_differentiation_b200 import device as D
    D.seed_uniform(42)
    a = np.asarray(D.uniform((1000, 333)))
    b = np.asarray(D.uniform((1000, 333)))
    D.seed_uniform(42)
    a2 = np.asarray(D.uniform((1000, 333)))
    c = np.asarray(D.uniform((7,)))                       # odd tail (n % 4 != 0)
    np.testing.assert_array_equal(a, a2)
    assert a.min() >= 0.0 and a.max() < 1.0 and c.shape == (7,)
    assert np.all(a * (1 << 24) == np.floor(a * (1 << 24)))
    assert abs(a.mean() - 0.5) < 2e-3 and abs(a.var() - 1.0 / 12.0) < 2e-3
    assert not np.array_equal(a, b) and abs(np.corrcoef(a.ravel(), b.ravel())[0, 1]) < 5e-3
    if backend == "cuda":                                 # known-answer: Philox4x32-10, counter 0, key (seed lo, seed hi)
        D.seed_uniform(0)
        first = np.asarray(D.uniform((4,)))
        ref = np.array([0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8], dtype=np.uint64)      # Random123 KAT, zero key / counter
        np.testing.assert_array_equal(first, ((ref >> 8).astype(np.float64) / (1 << 24)).astype(np.float32))
