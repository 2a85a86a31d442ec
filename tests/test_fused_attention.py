"""Fused attention core (b200ad_attn_fwd / b200ad_attn_bwd) through the C ABI wrappers vs an fp64 NumPy
statement of the reference math (MultiHeadAttention.forward layers.py:1351-1357 and the tape gradients of
Test/MultiHeadAttentionTest.py:70-85), and the MultiHeadAttention layer fused vs unfused.

Sizes cover: one 128-row tile / several tiles, one 64-column block / several, ragged tails in both the lane
and the column dimension (TMA zero fill + column masking), seq_q != seq_kv, more work items than SMs
(persistent loop, barrier phases across items), strided head-split views of merged (B*S, h*dk) matrices."""
import math

import numpy as np
import pytest

from conftest import assert_close


def _D():
    from auto_differentiation_b200 import device
    return device


def _ref(q, k, v, do, scale):
    q, k, v, do = (a.astype(np.float64) for a in (q, k, v, do))
    s = scale * (q @ k.transpose(0, 1, 3, 2))
    e = np.exp(s - s.max(-1, keepdims=True))
    p = e / e.sum(-1, keepdims=True)
    o = p @ v
    dv = p.transpose(0, 1, 3, 2) @ do
    dp = do @ v.transpose(0, 1, 3, 2)
    ds = p * (dp - (dp * p).sum(-1, keepdims=True)) * scale
    return o, ds @ k, ds.transpose(0, 1, 3, 2) @ q, dv


def _heads(D, a):
    """(B,h,S,dk) host array -> head-split device VIEW of a merged (B,S,h*dk) buffer (the layer's layout)."""
    B, h, S, dk = a.shape
    merged = D.asarray(np.ascontiguousarray(a.transpose(0, 2, 1, 3)).reshape(B, S, h * dk))
    return merged.reshape(B, S, h, dk).transpose(0, 2, 1, 3)


CASES = [
    # B, h, Sq, Skv
    (1, 1, 128, 64),
    (1, 2, 128, 128),
    (2, 3, 256, 192),
    (1, 2, 100, 100),      # ragged rows and columns
    (2, 2, 70, 200),       # Sq < one tile, ragged columns
    (1, 1, 333, 77),
    (3, 8, 512, 512),      # config 3/5 shape per sequence
    (40, 8, 128, 128),     # 320 work items > 148 SMs: persistent loop, odd/even block counts across items
    (25, 8, 192, 64),      # nb = 1: A tile released by the first score MMA
    (2, 2, 1, 1),          # a single query against a single key
    (1, 3, 5, 3),          # fewer rows than one warp, odd everything
    (1, 1, 129, 65),       # one row / one column past a tile / block boundary
]


@pytest.mark.parametrize("B,h,Sq,Skv", CASES)
def test_attn_fwd_bwd_vs_fp64(backend, B, h, Sq, Skv):
    D = _D()
    if backend == "emu" and B * h * Sq * Skv > 2_000_000:
        pytest.skip("CPU emulation: small cases only")
    dk = 64
    rng = np.random.default_rng(B * 1000 + Sq + Skv)
    q = rng.standard_normal((B, h, Sq, dk)).astype(np.float32)
    k = rng.standard_normal((B, h, Skv, dk)).astype(np.float32)
    v = rng.standard_normal((B, h, Skv, dk)).astype(np.float32)
    do = rng.standard_normal((B, h, Sq, dk)).astype(np.float32)
    scale = 1.0 / math.sqrt(dk)
    Q, K, V, dO = (_heads(D, a) for a in (q, k, v, do))
    assert D.attn_supported(Q, K, V)
    O, lse = D.attn_fwd(Q, K, V, scale)
    o_ref, dq_ref, dk_ref, dv_ref = _ref(q, k, v, do, scale)
    assert_close(np.asarray(O), o_ref, rtol=2e-5, atol=2e-6, what="O")
    # lse is the log2-domain log-sum-exp of the scaled scores
    s = scale * (q.astype(np.float64) @ k.astype(np.float64).transpose(0, 1, 3, 2))
    m = s.max(-1)
    lse_ref = (m + np.log(np.exp(s - m[..., None]).sum(-1))) / math.log(2.0)
    assert_close(np.asarray(lse), lse_ref, rtol=1e-5, atol=1e-6, what="lse")
    dQ, dK, dV = D.attn_bwd(Q, K, V, O, dO, lse, scale)
    assert_close(np.asarray(dV), dv_ref, rtol=2e-5, atol=2e-6, what="dV")
    assert_close(np.asarray(dQ), dq_ref, rtol=2e-5, atol=2e-6, what="dQ")
    assert_close(np.asarray(dK), dk_ref, rtol=2e-5, atol=2e-6, what="dK")


@pytest.mark.parametrize("dk", [32, 16, 48, 8, 4])
def test_attn_head_sizes_below_64(backend, dk):
    """Head sizes that are multiples of 4 up to 64 run on the fused passes (64-wide tiles, zero fill beyond d_k): both
    carriers, ragged lengths, against fp64.  (d_k = 5, the reference test's shape, is not 16 B aligned: composition path,
    tests/test_layers_parity.py.)"""
    D = _D()
    B, h, Sq, Skv = 3, 2, 200, 136
    rng = np.random.default_rng(dk)
    q = rng.standard_normal((B, h, Sq, dk)).astype(np.float32)
    k = rng.standard_normal((B, h, Skv, dk)).astype(np.float32)
    v = rng.standard_normal((B, h, Skv, dk)).astype(np.float32)
    do = rng.standard_normal((B, h, Sq, dk)).astype(np.float32)
    scale = 1.0 / math.sqrt(dk)
    Q, K, V, dO = (_heads(D, a) for a in (q, k, v, do))
    assert D.attn_supported(Q, K, V)
    ref = _ref(q, k, v, do, scale)
    variants = (-1,) if backend == "emu" else (-1, 0, 1)
    lib = None
    if backend != "emu":
        from auto_differentiation_b200 import _lib as L
        lib = L.load()
    prev = lib.b200ad_attn_set_variant(-1) if lib else None
    try:
        for variant in variants:
            if lib:
                lib.b200ad_attn_set_variant(variant)
            O, lse = D.attn_fwd(Q, K, V, scale)
            dQ, dK, dV = D.attn_bwd(Q, K, V, O, dO, lse, scale)
            for a, r, name in zip((O, dQ, dK, dV), ref, ("O", "dQ", "dK", "dV")):
                assert np.asarray(a).shape == r.shape
                assert_close(np.asarray(a), r, rtol=2e-5, atol=2e-6, what="d_k=%d variant %d %s" % (dk, variant, name))
    finally:
        if lib:
            lib.b200ad_attn_set_variant(prev)
    # head sizes the fused passes do not take
    bad = _heads(D, rng.standard_normal((1, 1, 16, 6)).astype(np.float32))
    assert not D.attn_supported(bad, bad, bad)


def test_attn_ragged_fuzz(backend):
    """Seeded random (batch, heads, seq_q, seq_kv) incl. lengths that are not multiples of the 128-row tile or the
    64-column block, odd tile counts, one-block and many-block rows."""
    D = _D()
    rng = np.random.default_rng(77)
    dk = 64
    scale = 1.0 / math.sqrt(dk)
    for _ in range(6):
        B, h = int(rng.integers(1, 4)), int(rng.integers(1, 5))
        Sq, Skv = int(rng.integers(1, 400)), int(rng.integers(1, 400))
        q = rng.standard_normal((B, h, Sq, dk)).astype(np.float32)
        k = rng.standard_normal((B, h, Skv, dk)).astype(np.float32)
        v = rng.standard_normal((B, h, Skv, dk)).astype(np.float32)
        do = rng.standard_normal((B, h, Sq, dk)).astype(np.float32)
        Q, K, V, dO = (_heads(D, a) for a in (q, k, v, do))
        O, lse = D.attn_fwd(Q, K, V, scale)
        o_ref, dq_ref, dk_ref, dv_ref = _ref(q, k, v, do, scale)
        tag = "B=%d h=%d Sq=%d Skv=%d " % (B, h, Sq, Skv)
        assert_close(np.asarray(O), o_ref, rtol=2e-5, atol=2e-6, what=tag + "O")
        dQ, dK, dV = D.attn_bwd(Q, K, V, O, dO, lse, scale)
        assert_close(np.asarray(dV), dv_ref, rtol=2e-5, atol=2e-6, what=tag + "dV")
        assert_close(np.asarray(dQ), dq_ref, rtol=2e-5, atol=2e-6, what=tag + "dQ")
        assert_close(np.asarray(dK), dk_ref, rtol=2e-5, atol=2e-6, what=tag + "dK")


@pytest.mark.parametrize("Sq,Skv", [(100, 77), (128, 64), (300, 200)])
def test_attn_output_guards(backend, Sq, Skv):
    """Outputs are windows (rows 1..S) of sentinel-filled buffers: the rows before / after every (batch) window, the lse
    neighbours and the bwd workspace neighbours must stay untouched (TMA-store clipping of ragged tiles)."""
    import ctypes
    D = _D()
    from auto_differentiation_b200 import _lib as L
    lib = L.load()
    rng = np.random.default_rng(Sq * 3 + Skv)
    B, h, dk = 2, 3, 64
    scale = 1.0 / math.sqrt(dk)
    SENT = 3.5

    def window(S, data=None):
        """(B, S+2, h, dk) buffer; returns (buffer, head-split view of rows 1..S)."""
        host = np.full((B, S + 2, h, dk), SENT, dtype=np.float32)
        if data is not None:
            host[:, 1:S + 1] = data.transpose(0, 2, 1, 3)
        buf = D.asarray(host)
        return buf, buf[:, 1:S + 1].transpose(0, 2, 1, 3)

    q = rng.standard_normal((B, h, Sq, dk)).astype(np.float32)
    k = rng.standard_normal((B, h, Skv, dk)).astype(np.float32)
    v = rng.standard_normal((B, h, Skv, dk)).astype(np.float32)
    do = rng.standard_normal((B, h, Sq, dk)).astype(np.float32)
    (_, Q), (_, K), (_, V), (_, dO) = window(Sq, q), window(Skv, k), window(Skv, v), window(Sq, do)
    obuf, O = window(Sq)
    n_lse = B * h * Sq
    lse = D.asarray(np.full(n_lse + 16, SENT, dtype=np.float32))
    d = D._attn_desc(Q, K, V, O, scale)
    s = D.stream_ptr()
    n_fws = int(lib.b200ad_attn_fwd_ws_floats(ctypes.byref(d)))
    assert n_fws == 4 * B * h
    fws = D.asarray(np.full(n_fws + 16, SENT, dtype=np.float32))
    L.check(lib.b200ad_attn_fwd(ctypes.byref(d), Q.ptr, K.ptr, V.ptr, O.ptr, lse.ptr + 32, fws.ptr + 32, s))
    fg = np.asarray(fws)
    assert np.all(fg[:8] == SENT) and np.all(fg[8 + n_fws:] == SENT), "forward workspace written out of range"
    o_ref, dq_ref, dk_ref, dv_ref = _ref(q, k, v, do, scale)
    og = np.asarray(obuf)
    assert_close(og[:, 1:Sq + 1].transpose(0, 2, 1, 3), o_ref, rtol=2e-5, atol=2e-6, what="O window")
    assert np.all(og[:, 0] == SENT) and np.all(og[:, Sq + 1] == SENT), "attention forward wrote outside its rows"
    lg = np.asarray(lse)
    assert np.all(lg[:8] == SENT) and np.all(lg[8 + n_lse:] == SENT), "lse written out of range"
    (qb, dQ), (kb, dK), (vb, dV) = window(Sq), window(Skv), window(Skv)
    n_ws = int(lib.b200ad_attn_bwd_ws_floats(ctypes.byref(d)))
    assert n_ws == B * h * (Sq + 4)
    ws = D.asarray(np.full(n_ws + 16, SENT, dtype=np.float32))
    L.check(lib.b200ad_attn_bwd(ctypes.byref(d), Q.ptr, K.ptr, V.ptr, O.ptr, dO.ptr, lse.ptr + 32, dQ.ptr, dK.ptr, dV.ptr,
                                ws.ptr + 32, s))
    for name, buf, S, ref in (("dQ", qb, Sq, dq_ref), ("dK", kb, Skv, dk_ref), ("dV", vb, Skv, dv_ref)):
        g = np.asarray(buf)
        assert_close(g[:, 1:S + 1].transpose(0, 2, 1, 3), ref, rtol=2e-5, atol=2e-6, what=name + " window")
        assert np.all(g[:, 0] == SENT) and np.all(g[:, S + 1] == SENT), name + " written outside its rows"
    wg = np.asarray(ws)
    assert np.all(wg[:8] == SENT) and np.all(wg[8 + n_ws:] == SENT), "backward workspace written out of range"


def test_attn_large_scores_stable(backend):
    """Scores of magnitude ~60 (the regime where the reference's naive exp/sum is still finite in fp64):
    the online softmax must agree with the stable fp64 evaluation."""
    D = _D()
    rng = np.random.default_rng(7)
    B, h, S, dk = 1, 2, 192, 64
    q = (rng.standard_normal((B, h, S, dk)) * 3).astype(np.float32)
    k = (rng.standard_normal((B, h, S, dk)) * 3).astype(np.float32)
    v = rng.standard_normal((B, h, S, dk)).astype(np.float32)
    do = rng.standard_normal((B, h, S, dk)).astype(np.float32)
    scale = 1.0 / math.sqrt(dk)
    Q, K, V, dO = (_heads(D, a) for a in (q, k, v, do))
    O, lse = D.attn_fwd(Q, K, V, scale)
    o_ref, dq_ref, dk_ref, dv_ref = _ref(q, k, v, do, scale)
    assert_close(np.asarray(O), o_ref, rtol=5e-5, atol=5e-6, what="O")
    dQ, dK, dV = D.attn_bwd(Q, K, V, O, dO, lse, scale)
    assert_close(np.asarray(dV), dv_ref, rtol=5e-5, atol=5e-6, what="dV")
    assert_close(np.asarray(dQ), dq_ref, rtol=5e-5, atol=5e-6, what="dQ")
    assert_close(np.asarray(dK), dk_ref, rtol=5e-5, atol=5e-6, what="dK")


def test_mha_layer_fused_equals_unfused(backend):
    """The layer's fused path and its GEMM + softmax composition give the same output and gradients."""
    D = _D()
    from auto_differentiation_b200.layers import layers as Ls
    rng = np.random.default_rng(11)
    B, S, d, h = 2, 160, 512, 8
    x = rng.standard_normal((B, S, d)).astype(np.float32)
    ws = [(rng.standard_normal((d, d)) / math.sqrt(d)).astype(np.float32) for _ in range(4)]
    dout = rng.standard_normal((B, S, d)).astype(np.float32)
    res = []
    for fused in (True, False):
        D.set_fused_attention(fused)
        try:
            m = Ls.MultiHeadAttention(d, d, h, weights=[w.copy() for w in ws])
            out = m.forward(x, x, x)
            assert (m.cache[6] is None) == fused
            dx = m.backward(dout)
            res.append([np.asarray(out)] + [np.asarray(g) for g in m.grads] + [np.asarray(g) for g in dx])
        finally:
            D.set_fused_attention(True)
    names = ["out", "dwq", "dwk", "dwv", "dwo", "dq", "dk", "dv"]
    for n, a, b in zip(names, res[0], res[1]):
        assert_close(a, b, rtol=3e-5, atol=3e-6, what=n)


@pytest.mark.gpu
def test_attn_config5_shape_properties():
    """Config-5 micro-batch shape (256 x 8 heads x 512 x 512 x 64) through size-independent properties:
    rows of V equal -> O equals that row exactly (softmax sums to one); dV column sums equal dO column sums
    (P^T 1 = column mass, sum over kv rows of dV = sum over q rows of dO); run-to-run bit reproducibility."""
    import torch
    assert torch.cuda.is_available()
    D = _D()
    rng = np.random.default_rng(5)
    B, h, S, dk = 64, 8, 512, 64
    q = rng.standard_normal((B, h, S, dk)).astype(np.float32)
    k = rng.standard_normal((B, h, S, dk)).astype(np.float32)
    vrow = rng.standard_normal((B, h, 1, dk)).astype(np.float32)
    v = np.broadcast_to(vrow, (B, h, S, dk)).copy()
    do = rng.standard_normal((B, h, S, dk)).astype(np.float32)
    scale = 1.0 / math.sqrt(dk)
    Q, K, V, dO = (_heads(D, a) for a in (q, k, v, do))
    O, lse = D.attn_fwd(Q, K, V, scale)
    o = np.asarray(O)
    assert np.allclose(o, np.broadcast_to(vrow, o.shape), rtol=2e-5, atol=2e-6)
    dQ, dK, dV = D.attn_bwd(Q, K, V, O, dO, lse, scale)
    dv = np.asarray(dV).astype(np.float64)
    assert np.allclose(dv.sum(axis=2), do.astype(np.float64).sum(axis=2), rtol=1e-4, atol=1e-3)
    # identical V rows -> dP rows are constant -> dS = 0 -> dQ = dK = 0 (up to rounding of dP - D)
    assert np.abs(np.asarray(dQ)).max() < 1e-3 and np.abs(np.asarray(dK)).max() < 1e-3
    O2, lse2 = D.attn_fwd(Q, K, V, scale)
    assert np.array_equal(np.asarray(O2), o) and np.array_equal(np.asarray(lse2), np.asarray(lse))
    dQ2, dK2, dV2 = D.attn_bwd(Q, K, V, O, dO, lse, scale)
    assert np.array_equal(np.asarray(dV2), np.asarray(dV)) and np.array_equal(np.asarray(dK2), np.asarray(dK))


@pytest.mark.gpu
@pytest.mark.parametrize("B,h,Sq,Skv", [(2, 8, 512, 512), (3, 2, 200, 136), (1, 1, 128, 64)])
def test_attn_variants_fp16_split_and_tf32_split_agree(B, h, Sq, Skv):
    """The two carriers of the fp32-faithful split — fp16 hi/lo at the kind::f16 rate (default) and tf32 hi/lo — against
    fp64 and against each other, selected through b200ad_attn_set_variant."""
    D = _D()
    from auto_differentiation_b200 import _lib as L
    lib = L.load()
    rng = np.random.default_rng(B + Sq + Skv)
    dk, scale = 64, 1.0 / 8.0
    q, k, v = (rng.standard_normal((B, h, S, dk)).astype(np.float32) for S in (Sq, Skv, Skv))
    do = rng.standard_normal((B, h, Sq, dk)).astype(np.float32)
    ref = _ref(q, k, v, do, scale)
    Q, K, V, dO = (_heads(D, a) for a in (q, k, v, do))
    got = {}
    prev = lib.b200ad_attn_set_variant(0)
    try:
        for variant in (0, 1, 2, -1):
            lib.b200ad_attn_set_variant(variant)
            O, lse = D.attn_fwd(Q, K, V, scale)
            dQ, dK, dV = D.attn_bwd(Q, K, V, O, dO, lse, scale)
            got[variant] = [np.asarray(a, dtype=np.float64) for a in (O, dQ, dK, dV)]
            for a, r, name in zip(got[variant], ref, ("O", "dQ", "dK", "dV")):
                assert_close(a, r, rtol=2e-5, atol=2e-6, what="variant %d %s" % (variant, name))
    finally:
        lib.b200ad_attn_set_variant(prev)
    for a, b, name in zip(got[0], got[1], ("O", "dQ", "dK", "dV")):
        assert_close(a, b, rtol=2e-5, atol=2e-6, what="fp16-split vs tf32-split " + name)
    # variant 0 computes dK and dV in one fused pass, variant 2 in two separate ones: same operands, same split
    for a, b, name in zip(got[0], got[2], ("O", "dQ", "dK", "dV")):
        assert_close(a, b, rtol=2e-6, atol=2e-7, what="fused dK+dV pass vs separate passes " + name)


@pytest.mark.gpu
def test_attn_fp16_split_operand_ranges():
    """The fp16 carrier rescales every (batch, head) operand slab by a power of two: gradients of 1e-8, values of 1e3,
    scores up to ~40 and slabs of very different magnitude in one launch must all come out fp32-faithful."""
    D = _D()
    from auto_differentiation_b200 import _lib as L
    prev = L.load().b200ad_attn_set_variant(0)
    try:
        _ranges_body(D)
    finally:
        L.load().b200ad_attn_set_variant(prev)


def _ranges_body(D):
    rng = np.random.default_rng(99)
    B, h, S, dk = 2, 4, 256, 64
    scale = 1.0 / 8.0
    q = (rng.standard_normal((B, h, S, dk)) * 2.5).astype(np.float32)
    k = (rng.standard_normal((B, h, S, dk)) * 2.5).astype(np.float32)
    v = rng.standard_normal((B, h, S, dk)).astype(np.float32)
    do = rng.standard_normal((B, h, S, dk)).astype(np.float32)
    slab = np.array([1e3, 1.0, 1e-3, 7.0, 1e-5, 300.0, 0.02, 1.0]).reshape(B, h, 1, 1)
    v = (v * slab).astype(np.float32)                       # per-slab magnitudes from 1e-5 to 1e3
    do = (do * 1e-8 * slab[::-1]).astype(np.float32)        # tiny upstream gradients, differently scaled per slab
    ref = _ref(q, k, v, do, scale)
    Q, K, V, dO = (_heads(D, a) for a in (q, k, v, do))
    O, lse = D.attn_fwd(Q, K, V, scale)
    dQ, dK, dV = D.attn_bwd(Q, K, V, O, dO, lse, scale)
    for a, r, name in zip((O, dQ, dK, dV), ref, ("O", "dQ", "dK", "dV")):
        a = np.asarray(a, dtype=np.float64)
        for b_ in range(B):
            for h_ in range(h):                              # per slab: the tolerance is relative to the SLAB's scale
                sc = max(float(np.abs(r[b_, h_]).max()), 1e-300)
                assert_close(a[b_, h_] / sc, r[b_, h_] / sc, rtol=5e-5, atol=5e-6, what="ranges %s slab (%d,%d)" % (name, b_, h_))


@pytest.mark.gpu
@pytest.mark.parametrize("B,h,Sq,Skv,dk", [(3, 2, 200, 136, 64), (2, 8, 512, 512, 64), (2, 3, 100, 300, 32)])
def test_attn_forward_hands_slab_maxima_to_backward(B, h, Sq, Skv, dk):
    """The forward leaves max |Q|, |K|, |V| per (batch, head) slab in its workspace; the backward of the same operands takes
    them from there instead of re-reading the three tensors.  The table must equal the slab maxima exactly, and the
    gradients must be bit-identical to a backward that computes the maxima itself."""
    D = _D()
    rng = np.random.default_rng(B + Sq + dk)
    mag = np.exp(rng.uniform(np.log(1e-3), np.log(1e2), (3, B, h, 1, 1)))          # slabs of very different magnitude
    q = (rng.standard_normal((B, h, Sq, dk)) * mag[0]).astype(np.float32)
    k = (rng.standard_normal((B, h, Skv, dk)) * mag[1]).astype(np.float32)
    v = (rng.standard_normal((B, h, Skv, dk)) * mag[2]).astype(np.float32)
    do = rng.standard_normal((B, h, Sq, dk)).astype(np.float32)
    scale = 1.0 / (math.sqrt(dk) * float(np.abs(q).max() * np.abs(k).max()) ** 0.5)   # keeps the scores moderate
    Q, K, V, dO = (_heads(D, a) for a in (q, k, v, do))
    O, lse = D.attn_fwd(Q, K, V, scale)
    table = getattr(lse, "fwd_ws", None)
    assert table is not None
    t = np.asarray(table).reshape(B * h, 4)
    for slot, a in enumerate((q, k, v)):
        assert np.array_equal(t[:, slot], np.abs(a).reshape(B * h, -1).max(axis=1)), "slab maxima, slot %d" % slot
    with_hint = D.attn_bwd(Q, K, V, O, dO, lse, scale)
    without = D.attn_bwd(Q, K, V, O, dO, D.DeviceArray(lse.t), scale)
    for a, b_, name in zip(with_hint, without, ("dQ", "dK", "dV")):
        assert np.array_equal(np.asarray(a), np.asarray(b_)), name
    ref = _ref(q, k, v, do, scale)
    for a, r, name in zip(with_hint, ref[1:], ("dQ", "dK", "dV")):
        assert_close(np.asarray(a), r, rtol=2e-5, atol=2e-6, what="maxima from the forward: " + name)


@pytest.mark.gpu
@pytest.mark.parametrize("B,h,Sq,Skv,dk", [(12, 8, 512, 512, 64), (9, 4, 333, 200, 64)])
def test_attn_passes_bit_reproducible_under_load(B, h, Sq, Skv, dk):
    """Two MMA-issuing warps, aliased score buffers and spin-wait hand-offs must not make a result depend on the schedule:
    the same inputs give bit-identical O / lse / dQ / dK / dV on every repetition, with unrelated work in between
    (profiles/attn_stress.py is the long version: 510 repetitions, profiles/r2_attn_stress.log)."""
    import torch
    D = _D()
    g = torch.Generator(device="cuda").manual_seed(B + Sq)
    bufs = [torch.randn((B, S, h * dk), device="cuda", generator=g) for S in (Sq, Skv, Skv, Sq)]
    Q, K, V, dO = (D.DeviceArray(b).reshape(B, b.shape[1], h, dk).transpose(0, 2, 1, 3) for b in bufs)
    scale = 1.0 / math.sqrt(dk)
    noise = torch.randn((2048, 2048), device="cuda", generator=g)
    ref = None
    for i in range(12):
        if i % 3 == 1:
            noise = noise @ noise.T * 1e-4
        O, lse = D.attn_fwd(Q, K, V, scale)
        cur = [t.t.clone() for t in (O, lse) + tuple(D.attn_bwd(Q, K, V, O, dO, lse, scale))]
        if ref is None:
            ref = cur
        else:
            for name, a, b_ in zip(("O", "lse", "dQ", "dK", "dV"), ref, cur):
                assert torch.equal(a, b_), "%s differs on repetition %d" % (name, i)


@pytest.mark.gpu
@pytest.mark.parametrize("B,h,Sq,Skv", [(2, 8, 512, 512), (3, 2, 200, 136)])
def test_attn_looser_tolerance_variant(B, h, Sq, Skv):
    """b200ad_attn_set_variant(3): only the hi . hi fp16 product of every contraction (operands rounded to 11 significant
    bits, fp32 accumulation) — the attention counterpart of the bf16 GEMM variant (north_star: "to a stated looser
    tolerance").  Stated tolerance: 4e-3 of the tensor scale (rtol = atol-of-scale = 4e-3); it must also be a DIFFERENT
    result from the fp32-faithful one (else the variant is not being exercised)."""
    D = _D()
    from auto_differentiation_b200 import _lib as L
    lib = L.load()
    rng = np.random.default_rng(B * 7 + Sq)
    dk, scale = 64, 1.0 / 8.0
    q, k, v = (rng.standard_normal((B, h, S, dk)).astype(np.float32) for S in (Sq, Skv, Skv))
    do = rng.standard_normal((B, h, Sq, dk)).astype(np.float32)
    ref = _ref(q, k, v, do, scale)
    Q, K, V, dO = (_heads(D, a) for a in (q, k, v, do))
    prev = lib.b200ad_attn_set_variant(3)
    try:
        O, lse = D.attn_fwd(Q, K, V, scale)
        grads = D.attn_bwd(Q, K, V, O, dO, lse, scale)
    finally:
        lib.b200ad_attn_set_variant(prev)
    got = [np.asarray(a, dtype=np.float64) for a in (O,) + tuple(grads)]
    for a, r, name in zip(got, ref, ("O", "dQ", "dK", "dV")):
        assert_close(a, r, rtol=4e-3, atol=4e-3, what="single-product variant " + name)
    worst = max(float(np.abs(a - r).max() / np.abs(r).max()) for a, r in zip(got, ref))
    assert worst > 2e-5, "variant 3 produced fp32-faithful results: the single-product path was not taken"
