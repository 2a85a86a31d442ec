"""TEST INFRASTRUCTURE — a NumPy emulation of the C ABI in include/b200ad.h, used ONLY by the
`-m "not gpu"` tests to exercise the HOST-side logic of the package (tape sweep, broadcasting planner,
strided-view GEMM planner, restructured recurrent layers, optimizer tables, data-parallel bucketing) on
machines without a GPU.  It is installed by a pytest fixture that monkey-patches
`auto_differentiation_b200._lib.load` and makes `device.py` allocate on the CPU; the product never
imports this file and has no CPU path of its own.  The real kernels are checked against the oracle by the
`-m gpu` tests, which run the very same assertions through libb200ad.so.

Each function reads/writes raw addresses exactly as the CUDA entry would (same argument order).
"""
import ctypes

import numpy as np


def _f32(ptr, n):
    if n <= 0 or not ptr:
        return np.zeros(0, np.float32)
    return np.ctypeslib.as_array((ctypes.c_float * int(n)).from_address(int(ptr)))


def _i64(ptr, n):
    if n <= 0 or not ptr:
        return np.zeros(0, np.int64)
    return np.ctypeslib.as_array((ctypes.c_int64 * int(n)).from_address(int(ptr)))


def _rows(ptr, ld, rows, cols):
    """strided (rows, cols) view with row stride ld (elements)."""
    if rows == 0 or cols == 0:
        return np.zeros((rows, cols), np.float32)
    flat = _f32(ptr, (rows - 1) * ld + cols)
    return np.lib.stride_tricks.as_strided(flat, (rows, cols), (ld * 4, 4))


def _ival(x):
    return int(x.value) if hasattr(x, "value") else int(x or 0)


_U = {
    0: np.exp, 1: np.tanh, 2: lambda x: 1 / (1 + np.exp(-x)), 3: lambda x: np.maximum(x, 0), 4: np.log,
    6: np.sqrt, 7: np.square, 9: np.abs, 10: np.negative, 11: np.sin, 12: np.cos, 13: np.tan,
    14: lambda x: 1 / np.cos(x), 15: np.sinh, 16: np.cosh, 17: np.arcsin, 18: np.arctan,
    19: lambda x: np.arctan(1 / x), 20: lambda x: np.arccos(1 / x),
    26: lambda x: 1 / np.tan(x), 27: lambda x: 1 / np.sin(x), 28: np.arccos,
}
_USES = {0: 1, 1: 1, 2: 1, 3: 1, 6: 1, 10: 2, 21: 2, 22: 2, 23: 2, 24: 2}


class Emulator:
    """Object with the same attribute names as the ctypes library handle."""

    def __init__(self):
        self.launches = 0
        self._err = b""

    # ---- runtime ---------------------------------------------------------------------------------
    def b200ad_abi_version(self):
        return 1

    def b200ad_last_error(self):
        return self._err

    def b200ad_launch_count(self):
        return self.launches

    def _fail(self, msg, code=-1):
        self._err = msg.encode()
        return code

    # ---- unary -----------------------------------------------------------------------------------
    def b200ad_unary_bwd_uses(self, op):
        return _USES.get(op, 0)

    def b200ad_unary_fwd(self, op, x, y, n, p, s):
        self.launches += 1
        xs = _f32(x, n).astype(np.float64)
        with np.errstate(all="ignore"):
            if op in _U:
                r = _U[op](xs)
            elif op == 5:
                r = np.log(xs) / np.log(p)
            elif op == 8:
                r = np.power(xs, p)
            elif op == 21:
                r = p * xs
            elif op == 22:
                r = xs + p
            elif op == 23:
                r = xs / p
            elif op == 24:
                r = p - xs
            elif op == 25:
                r = p / xs
            else:
                return self._fail("bad op")
        _f32(y, n)[:] = r
        return 0

    def b200ad_unary_bwd(self, op, g, x, y, gx, n, p, s):
        self.launches += 1
        gg = _f32(g, n).astype(np.float64)
        uses = _USES.get(op, 0)
        sv = (_f32(x, n) if uses == 0 else _f32(y, n) if uses == 1 else gg).astype(np.float64)
        with np.errstate(all="ignore"):
            r = {
                0: lambda: gg * sv, 1: lambda: gg * (1 - sv * sv), 2: lambda: gg * sv * (1 - sv),
                3: lambda: gg * (sv > 0), 4: lambda: gg / sv, 5: lambda: gg / (sv * np.log(p)),
                6: lambda: gg * 0.5 / sv, 7: lambda: gg * 2 * sv, 8: lambda: gg * p * np.power(sv, p - 1),
                9: lambda: np.where(sv < 0, -gg, gg), 10: lambda: -gg, 11: lambda: gg * np.cos(sv),
                12: lambda: -gg * np.sin(sv), 13: lambda: gg / np.cos(sv) ** 2,
                14: lambda: gg * np.tan(sv) / np.cos(sv), 15: lambda: gg * np.cosh(sv), 16: lambda: gg * np.sinh(sv),
                17: lambda: gg / np.sqrt(1 - sv * sv), 18: lambda: gg / (1 + sv * sv), 19: lambda: -gg / (1 + sv * sv),
                20: lambda: gg / (np.abs(sv) * np.sqrt(sv * sv - 1)), 21: lambda: gg * p, 22: lambda: gg,
                23: lambda: gg / (p + 1e-8), 24: lambda: -gg, 25: lambda: gg * (-p / (sv * sv + 1e-8)),
                26: lambda: gg / np.sin(sv) ** 2, 27: lambda: -gg * np.cos(sv) / np.sin(sv) ** 2,
                28: lambda: gg * (-1.0 / np.sqrt(sv + 1.0)),
            }[op]()
        _f32(gx, n)[:] = r
        return 0

    # ---- binary ----------------------------------------------------------------------------------
    @staticmethod
    def _operand(ptr, mode, n, period):
        idx = np.arange(n)
        if mode == 0:
            return _f32(ptr, n).astype(np.float64)
        if mode == 1:
            return _f32(ptr, period).astype(np.float64)[idx % period]
        if mode == 2:
            return _f32(ptr, n // period).astype(np.float64)[idx // period]
        return np.full(n, float(_f32(ptr, 1)[0]))

    def b200ad_binary_fwd(self, op, a, b, out, n, ma, mb, period, s):
        self.launches += 1
        if ma in (1, 2) or mb in (1, 2):
            if period <= 0 or n % period:
                return self._fail("period does not divide n")
        period = max(period, 1)
        x, y = self._operand(a, ma, n, period), self._operand(b, mb, n, period)
        _f32(out, n)[:] = [np.add, np.subtract, np.multiply, np.divide][op](x, y)
        return 0

    @staticmethod
    def _bwd_pair(op, g, x, y, ref_compat):
        if op == 0:
            return g, g
        if op == 1:
            return (-g, g) if ref_compat else (g, -g)
        if op == 2:
            return g * y, g * x
        return g / (y + 1e-8), g * (-x / (y * y + 1e-8))

    def b200ad_binary_bwd(self, op, g, a, b, ga, gb, n, ma, mb, period, ref_compat, s):
        self.launches += 1
        period = max(period, 1)
        gg = _f32(g, n).astype(np.float64)
        x = self._operand(a, ma, n, period) if op >= 2 else None
        y = self._operand(b, mb, n, period) if op >= 2 else None
        ra, rb = self._bwd_pair(op, gg, x, y, ref_compat)
        if ga:
            _f32(ga, n)[:] = ra
        if gb:
            _f32(gb, n)[:] = rb
        return 0

    def b200ad_binary_bwd_rowfused(self, op, g, a, b, ga, gb, rows, cols, ref_compat, ws, s):
        self.launches += 1
        n = rows * cols
        gg = _f32(g, n).astype(np.float64)
        x = _f32(a, n).astype(np.float64) if op >= 2 else None
        y = np.tile(_f32(b, cols).astype(np.float64), rows) if op >= 2 else None
        ra, rb = self._bwd_pair(op, gg, x, y, ref_compat)
        if ga:
            _f32(ga, n)[:] = ra
        _f32(gb, cols)[:] = rb.reshape(rows, cols).sum(axis=0)
        return 0

    # ---- reductions ------------------------------------------------------------------------------
    def b200ad_reduce_ws_floats(self, outer, red, inner):
        return 8

    def b200ad_colsum_ws_floats(self, rows, cols):
        return 8

    def b200ad_reduce(self, inp, out, outer, red, inner, scale, ws, s):
        self.launches += 1
        if outer * inner == 0:
            return 0
        v = _f32(inp, outer * red * inner).astype(np.float64).reshape(outer, red, inner)
        _f32(out, outer * inner)[:] = (v.sum(axis=1) * scale).reshape(-1)
        return 0

    def b200ad_repeat_axis(self, inp, out, outer, rep, inner, scale, s):
        self.launches += 1
        v = _f32(inp, outer * inner).reshape(outer, 1, inner)
        _f32(out, outer * rep * inner)[:] = (np.repeat(v, rep, axis=1) * scale).reshape(-1)
        return 0

    # ---- softmax family --------------------------------------------------------------------------
    def b200ad_softmax_fwd(self, x, y, rows, cols, in_scale, s):
        self.launches += 1
        v = _f32(x, rows * cols).astype(np.float64).reshape(rows, cols) * in_scale
        e = np.exp(v - v.max(axis=1, keepdims=True))
        _f32(y, rows * cols)[:] = (e / e.sum(axis=1, keepdims=True)).reshape(-1)
        return 0

    def b200ad_softmax_bwd(self, g, y, gx, rows, cols, out_scale, diag_only, s):
        self.launches += 1
        gg = _f32(g, rows * cols).astype(np.float64).reshape(rows, cols)
        yy = _f32(y, rows * cols).astype(np.float64).reshape(rows, cols)
        if diag_only:
            r = gg * yy * (1 - yy)
        else:
            dx = gg * yy
            r = dx - yy * dx.sum(axis=1, keepdims=True)
        _f32(gx, rows * cols)[:] = (out_scale * r).reshape(-1)
        return 0

    def b200ad_xent_fwd(self, logits, labels, probs, row_loss, lc, rows, cols, ignore, eps, denom_mode, is_probs, s):
        self.launches += 2
        if rows <= 0 or cols <= 0:
            return self._fail("xent_fwd: empty input")
        v = _f32(logits, rows * cols).astype(np.float64).reshape(rows, cols)
        t = _i64(labels, rows).copy()
        mask = t != ignore
        tw = np.where(t < 0, t + cols, t)
        if is_probs:
            p = v
        else:
            e = np.exp(v - v.max(axis=1, keepdims=True))
            p = e / e.sum(axis=1, keepdims=True)
        if probs:
            _f32(probs, rows * cols)[:] = p.reshape(-1)
        safe = np.clip(tw, 0, cols - 1)
        with np.errstate(all="ignore"):
            l = -np.log(p[np.arange(rows), safe] + eps)
        l = np.where(mask, l, 0.0)
        l = np.where(mask & ((tw < 0) | (tw >= cols)), np.nan, l)
        _f32(row_loss, rows)[:] = l
        denom = mask.sum() if denom_mode == 0 else rows
        out = _f32(lc, 2)
        with np.errstate(all="ignore"):
            out[0] = np.float64(l.sum()) / denom
        out[1] = denom
        return 0

    def b200ad_xent_bwd(self, probs, labels, lc, dx, rows, cols, ignore, dout, denom_mode, s):
        self.launches += 1
        p = _f32(probs, rows * cols).astype(np.float64).reshape(rows, cols).copy()
        t = _i64(labels, rows)
        mask = t != ignore
        tw = np.where(t < 0, t + cols, t)
        ok = mask & (tw >= 0) & (tw < cols)
        p[np.arange(rows)[ok], tw[ok]] -= 1
        p *= dout / float(_f32(lc, 2)[1])
        p *= mask[:, None]
        _f32(dx, rows * cols)[:] = p.reshape(-1)
        return 0

    # ---- index work ------------------------------------------------------------------------------
    def b200ad_argext_flat(self, x, n, is_min, out, ws, s):
        self.launches += 2
        if n <= 0:
            return self._fail("argmax: attempt to get argmax of an empty sequence")
        v = _f32(x, n)
        _i64(out, 1)[0] = np.argmin(v) if is_min else np.argmax(v)
        return 0

    def b200ad_argmax_rows(self, x, rows, cols, out, s):
        self.launches += 1
        _i64(out, rows)[:] = _f32(x, rows * cols).reshape(rows, cols).argmax(axis=1)
        return 0

    def b200ad_embedding_gather(self, w, idx, out, n_idx, vocab, dim, s):
        self.launches += 1
        W = _f32(w, vocab * dim).reshape(vocab, dim)
        _f32(out, n_idx * dim)[:] = W[_i64(idx, n_idx)].reshape(-1)
        return 0

    def b200ad_embedding_scatter_add(self, dout, idx, dw, n_idx, vocab, dim, s):
        self.launches += 1
        acc = np.zeros((vocab, dim), np.float64)
        np.add.at(acc, _i64(idx, n_idx), _f32(dout, n_idx * dim).reshape(n_idx, dim).astype(np.float64))
        _f32(dw, vocab * dim)[:] = acc.reshape(-1)
        return 0

    # ---- layout ----------------------------------------------------------------------------------
    def b200ad_strided_copy(self, inp, out, rank, shape, in_strides, out_strides, s):
        self.launches += 1
        shp = [int(shape[i]) for i in range(rank)]
        ist = [int(in_strides[i]) for i in range(rank)]
        total = int(np.prod(shp)) if rank else 1
        if total == 0:
            return 0
        span_in = 1 + sum((n - 1) * st for n, st in zip(shp, ist))
        src = np.lib.stride_tricks.as_strided(_f32(inp, span_in), shp, [st * 4 for st in ist])
        if out_strides:
            ost = [int(out_strides[i]) for i in range(rank)]
            span_out = 1 + sum((n - 1) * st for n, st in zip(shp, ost))
            dst = np.lib.stride_tricks.as_strided(_f32(out, span_out), shp, [st * 4 for st in ost])
            dst[...] = src
        else:
            _f32(out, total)[:] = np.ascontiguousarray(src).reshape(-1)
        return 0

    def b200ad_fill(self, out, n, value, s):
        self.launches += 1
        _f32(out, n)[:] = value
        return 0

    def b200ad_axpby(self, x, y, n, a, b, s):
        self.launches += 1
        yy = _f32(y, n)
        yy[:] = a * _f32(x, n).astype(np.float64) + (b * yy.astype(np.float64) if b != 0 else 0.0)
        return 0

    def b200ad_dropout_fwd(self, x, u, mask, y, n, ratio, scale, s):
        self.launches += 1
        m = np.where(_f32(u, n) > np.float32(ratio), np.float32(scale), np.float32(0.0))
        _f32(mask, n)[:] = m
        _f32(y, n)[:] = _f32(x, n) * m
        return 0

    # ---- GEMM --------------------------------------------------------------------------------------
    def b200ad_gemm_workspace_bytes(self, d):
        return 0

    def b200ad_uniform(self, out, n, seed, offset, s):
        self.launches += 1
        rng = np.random.default_rng([int(seed) & 0xFFFFFFFF, int(seed) >> 32, int(offset)])
        _f32(out, n)[...] = np.floor(rng.random(n) * (1 << 24)) / (1 << 24)
        return 0

    def b200ad_gemm_supported(self, d, a, b, c):
        return 1

    def b200ad_gemm_reload_env(self):
        return 0

    def b200ad_gemm_set_recurrent_dispatch(self, mode):
        return 0

    def b200ad_gemm_set_f16x3_auto(self, mode):
        return 1

    def b200ad_gemm_shared_a_workspace_bytes(self, d, count):
        return 0

    def b200ad_gemm_shared_a(self, dref, a, count, bs, cs, biases, ws, ws_bytes, s):
        for i in range(count):
            rc = self.b200ad_gemm(dref, a, bs[i], cs[i], biases[i] if biases else None, ws, ws_bytes, s)
            if rc:
                return rc
        return 0

    def b200ad_gemm(self, dref, a, b, c, bias, ws, ws_bytes, s):
        self.launches += 1
        d = dref._obj if hasattr(dref, "_obj") else dref
        M, N, K = d.m, d.n, d.k
        a, b, c, bias = _ival(a), _ival(b), _ival(c), _ival(bias)
        a_in = d.m if d.trans_a else d.k
        b_in = d.k if d.trans_b else d.n
        if d.lda < a_in or d.ldb < b_in or d.ldc < d.n:
            return self._fail("gemm: leading dimension smaller than the row")
        if M == 0 or N == 0:
            return 0
        for i0 in range(d.batch0):
            for i1 in range(d.batch1):
                pa = a + 4 * (i0 * d.a_stride0 + i1 * d.a_stride1)
                pb = b + 4 * (i0 * d.b_stride0 + i1 * d.b_stride1)
                pc = c + 4 * (i0 * d.c_stride0 + i1 * d.c_stride1)
                A = (_rows(pa, d.lda, K, M).T if d.trans_a else _rows(pa, d.lda, M, K)).astype(np.float64)
                B = (_rows(pb, d.ldb, N, K).T if d.trans_b else _rows(pb, d.ldb, K, N)).astype(np.float64)
                C = _rows(pc, d.ldc, M, N)
                r = d.alpha * (A @ B)
                if d.beta != 0:
                    r = r + d.beta * C.astype(np.float64)
                if bias:
                    r = r + _f32(bias, N).astype(np.float64)
                if d.act == 1:
                    r = np.tanh(r)
                elif d.act == 2:
                    r = 1 / (1 + np.exp(-r))
                elif d.act == 3:
                    r = np.maximum(r, 0)
                C[...] = r
        return 0

    # ---- fused attention ---------------------------------------------------------------------------
    @staticmethod
    def _attn_view(ptr, st, B, H, S, dk):
        n = (B - 1) * st.batch + (H - 1) * st.head + (S - 1) * st.row + dk
        flat = _f32(ptr, n)
        return np.lib.stride_tricks.as_strided(flat, (B, H, S, dk), (st.batch * 4, st.head * 4, st.row * 4, 4))

    def b200ad_attn_supported(self, dref):
        d = dref._obj if hasattr(dref, "_obj") else dref
        return 1 if 4 <= d.head_dim <= 64 and d.head_dim % 4 == 0 else 0

    def b200ad_attn_bwd_ws_floats(self, dref):
        d = dref._obj if hasattr(dref, "_obj") else dref
        return d.batch * d.heads * (d.seq_q + 4)

    def b200ad_attn_fwd_ws_floats(self, dref):
        d = dref._obj if hasattr(dref, "_obj") else dref
        return 4 * d.batch * d.heads

    def b200ad_attn_set_variant(self, variant):
        return 0

    def _attn_probs(self, d, q, k):
        Q = self._attn_view(_ival(q), d.q, d.batch, d.heads, d.seq_q, d.head_dim).astype(np.float64)
        K = self._attn_view(_ival(k), d.k, d.batch, d.heads, d.seq_kv, d.head_dim).astype(np.float64)
        S = d.scale * (Q @ K.transpose(0, 1, 3, 2))
        m = S.max(axis=-1, keepdims=True)
        e = np.exp(S - m)
        l = e.sum(axis=-1, keepdims=True)
        return Q, K, e / l, (m + np.log(l))[..., 0]

    def b200ad_attn_fwd(self, dref, q, k, v, o, lse, ws, s):
        self.launches += 1
        d = dref._obj if hasattr(dref, "_obj") else dref
        if not (4 <= d.head_dim <= 64 and d.head_dim % 4 == 0):
            return self._fail("attn_fwd: head_dim must be a multiple of 4 in 4..64", -3)
        _, _, P, ln_lse = self._attn_probs(d, q, k)
        V = self._attn_view(_ival(v), d.v, d.batch, d.heads, d.seq_kv, d.head_dim).astype(np.float64)
        self._attn_view(_ival(o), d.o, d.batch, d.heads, d.seq_q, d.head_dim)[...] = P @ V
        _f32(_ival(lse), d.batch * d.heads * d.seq_q)[:] = (ln_lse / np.log(2.0)).reshape(-1)
        return 0

    def b200ad_attn_bwd_maxima(self, dref, q, k, v, o, dout, lse, fwd_ws, dq, dk, dv, ws, s):
        return self.b200ad_attn_bwd(dref, q, k, v, o, dout, lse, dq, dk, dv, ws, s)

    def b200ad_attn_bwd(self, dref, q, k, v, o, dout, lse, dq, dk, dv, ws, s):
        self.launches += 4
        d = dref._obj if hasattr(dref, "_obj") else dref
        B, H, Sq, Skv, dkk = d.batch, d.heads, d.seq_q, d.seq_kv, d.head_dim
        Q = self._attn_view(_ival(q), d.q, B, H, Sq, dkk).astype(np.float64)
        K = self._attn_view(_ival(k), d.k, B, H, Skv, dkk).astype(np.float64)
        V = self._attn_view(_ival(v), d.v, B, H, Skv, dkk).astype(np.float64)
        O = self._attn_view(_ival(o), d.o, B, H, Sq, dkk).astype(np.float64)
        dO = self._attn_view(_ival(dout), d.o, B, H, Sq, dkk).astype(np.float64)
        lse2 = _f32(_ival(lse), B * H * Sq).astype(np.float64).reshape(B, H, Sq, 1)
        P = np.exp2(d.scale * np.log2(np.e) * (Q @ K.transpose(0, 1, 3, 2)) - lse2)      # recomputed from the LSE
        Dv = (dO * O).sum(axis=-1, keepdims=True)
        _f32(_ival(ws), B * H * Sq)[:] = Dv.reshape(-1)
        dP = dO @ V.transpose(0, 1, 3, 2)
        dS = P * (dP - Dv) * d.scale
        self._attn_view(_ival(dv), d.v, B, H, Skv, dkk)[...] = P.transpose(0, 1, 3, 2) @ dO
        self._attn_view(_ival(dk), d.k, B, H, Skv, dkk)[...] = dS.transpose(0, 1, 3, 2) @ Q
        self._attn_view(_ival(dq), d.q, B, H, Sq, dkk)[...] = dS @ K
        return 0

    # ---- cells -----------------------------------------------------------------------------------
    def b200ad_lstm_cell_fwd(self, a, lda, c_prev, ldcp, gates, ldg, c_next, ldcn, h_next, ldh, n, h, s):
        self.launches += 1
        A = _rows(a, lda, n, 4 * h).astype(np.float64)
        cp = _rows(c_prev, ldcp, n, h).astype(np.float64)
        sg = lambda v: 1 / (1 + np.exp(-v))
        f, g, i, o = sg(A[:, :h]), np.tanh(A[:, h:2 * h]), sg(A[:, 2 * h:3 * h]), sg(A[:, 3 * h:])
        cn = f * cp + g * i
        _rows(gates, ldg, n, 4 * h)[...] = np.hstack((f, g, i, o))
        _rows(c_next, ldcn, n, h)[...] = cn
        _rows(h_next, ldh, n, h)[...] = o * np.tanh(cn)
        return 0

    def b200ad_lstm_cell_bwd(self, dh_a, ld_dha, dh_b, ld_dhb, dc_next, gates, ldg, c_prev, ldcp, c_next, ldcn,
                             h_next, ldh, da, ldda, dc_prev, n, h, s):
        self.launches += 1
        d = _rows(dh_a, ld_dha, n, h).astype(np.float64)
        if dh_b:
            d = d + _rows(dh_b, ld_dhb, n, h)
        dcn = _rows(dc_next, h, n, h).astype(np.float64) if dc_next else np.zeros((n, h))
        G = _rows(gates, ldg, n, 4 * h).astype(np.float64)
        f, g, i, o = G[:, :h], G[:, h:2 * h], G[:, 2 * h:3 * h], G[:, 3 * h:]
        cp = _rows(c_prev, ldcp, n, h).astype(np.float64)
        tc = np.tanh(_rows(c_next, ldcn, n, h).astype(np.float64))
        ds = d * o * (1 - tc * tc) + dcn
        _rows(da, ldda, n, 4 * h)[...] = np.hstack((ds * f * cp * (1 - f), ds * i * (1 - g * g), ds * g * i * (1 - i),
                                                    d * (o * tc) * (1 - o)))
        _rows(dc_prev, h, n, h)[...] = ds * f
        return 0

    @staticmethod
    def _actb(g, y, act):
        if act == 1:
            return g * (1 - y * y)
        if act == 2:
            return g * y * (1 - y)
        if act == 3:
            return g * (y > 0)
        return g

    def b200ad_rnn_cell_bwd(self, dh_a, ld_dha, dh_b, ld_dhb, h_next, ldh, dt, lddt, n, h, s):
        self.launches += 1
        d = _rows(dh_a, ld_dha, n, h).astype(np.float64)
        if dh_b:
            d = d + _rows(dh_b, ld_dhb, n, h)
        _rows(dt, lddt, n, h)[...] = self._actb(d, _rows(h_next, ldh, n, h).astype(np.float64), 1)
        return 0

    def b200ad_lstm_seq_supported(self, batch, hidden):
        return 1 if (batch >= 1 and hidden >= 4 and hidden % 4 == 0 and batch * 4 * hidden * hidden >= (1 << 18)) else 0

    def b200ad_lstm_seq_workspace_bytes(self, batch, hidden):
        return 0

    def b200ad_lstm_seq_fwd(self, hs_ext, cs_ext, gates, pre, wh, T, N, H, ws, ws_bytes, s):
        w = _f32(wh, H * 4 * H).reshape(H, 4 * H).astype(np.float64)
        for t in range(T):
            a_t = _ival(pre) + 4 * t * N * 4 * H
            A = _f32(a_t, N * 4 * H).reshape(N, 4 * H)
            A[...] = A.astype(np.float64) + _f32(_ival(hs_ext) + 4 * t * N * H, N * H).reshape(N, H).astype(np.float64) @ w
            self.launches += 1
            self.b200ad_lstm_cell_fwd(a_t, 4 * H, _ival(cs_ext) + 4 * t * N * H, H, _ival(gates) + 4 * t * N * 4 * H, 4 * H,
                                      _ival(cs_ext) + 4 * (t + 1) * N * H, H, _ival(hs_ext) + 4 * (t + 1) * N * H, H, N, H, s)
        return 0

    def b200ad_lstm_seq_bwd(self, dhs, sn, st, hs_ext, cs_ext, gates, wh, dA, dh0, dc_scratch, T, N, H, ws, ws_bytes, s):
        w = _f32(wh, H * 4 * H).reshape(H, 4 * H).astype(np.float64)
        have = False
        for t in reversed(range(T)):
            dc_out = _ival(dc_scratch) + 4 * (t & 1) * N * H
            dc_in = None if t == T - 1 else _ival(dc_scratch) + 4 * ((t + 1) & 1) * N * H
            da_t = _ival(dA) + 4 * t * N * 4 * H
            self.b200ad_lstm_cell_bwd(_ival(dhs) + 4 * t * st, sn, _ival(dh0) if have else None, H, dc_in,
                                      _ival(gates) + 4 * t * N * 4 * H, 4 * H, _ival(cs_ext) + 4 * t * N * H, H,
                                      _ival(cs_ext) + 4 * (t + 1) * N * H, H, None, H, da_t, 4 * H, dc_out, N, H, s)
            _f32(dh0, N * H).reshape(N, H)[...] = _f32(da_t, N * 4 * H).reshape(N, 4 * H).astype(np.float64) @ w.T
            self.launches += 1
            have = True
        return 0

    def b200ad_rnn_seq_supported(self, batch, hidden):
        return 1 if (1 <= batch <= 64 and hidden in (256, 512, 1024)) else 0

    def b200ad_rnn_tanh_seq_fwd(self, hs_ext, pre, wh, T, N, H, s):
        self.launches += 1
        hs = _f32(hs_ext, (T + 1) * N * H).reshape(T + 1, N, H)
        pr = _f32(pre, T * N * H).reshape(T, N, H)
        w = _f32(wh, H * H).reshape(H, H).astype(np.float64)
        for t in range(T):
            hs[t + 1] = np.tanh(pr[t].astype(np.float64) + hs[t].astype(np.float64) @ w)
        return 0

    def b200ad_rnn_tanh_seq_bwd(self, dhs, sn, st, hs_ext, wh, dA, dh0, T, N, H, s):
        self.launches += 1
        hs = _f32(hs_ext, (T + 1) * N * H).reshape(T + 1, N, H)
        w = _f32(wh, H * H).reshape(H, H).astype(np.float64)
        da = _f32(dA, T * N * H).reshape(T, N, H)
        span = (N - 1) * sn + (T - 1) * st + H
        g = np.lib.stride_tricks.as_strided(_f32(dhs, span), shape=(N, T, H), strides=(4 * sn, 4 * st, 4))
        dh = np.zeros((N, H))
        for t in reversed(range(T)):
            da[t] = (g[:, t, :].astype(np.float64) + dh) * (1.0 - hs[t + 1].astype(np.float64) ** 2)
            dh = da[t].astype(np.float64) @ w.T
        if dh0:
            _f32(dh0, N * H).reshape(N, H)[...] = dh
        return 0

    def b200ad_act_bwd(self, g, ldg, y, ldy, gx, ldgx, rows, cols, act, s):
        self.launches += 1
        _rows(gx, ldgx, rows, cols)[...] = self._actb(_rows(g, ldg, rows, cols).astype(np.float64),
                                                      _rows(y, ldy, rows, cols).astype(np.float64), act)
        return 0

    def b200ad_bias_act_fwd(self, x, ldx, bias, out, ldo, rows, cols, act, s):
        self.launches += 1
        v = _rows(x, ldx, rows, cols).astype(np.float64)
        if bias:
            v = v + _f32(bias, cols)
        if act == 1:
            v = np.tanh(v)
        elif act == 2:
            v = 1 / (1 + np.exp(-v))
        elif act == 3:
            v = np.maximum(v, 0)
        _rows(out, ldo, rows, cols)[...] = v
        return 0

    # ---- optimizers ------------------------------------------------------------------------------
    @staticmethod
    def _entries(table, n):
        raw = np.ctypeslib.as_array((ctypes.c_int64 * (5 * n)).from_address(int(table))).reshape(n, 5)
        return [tuple(int(v) for v in row) for row in raw]

    def b200ad_adam_step_scalar(self, iter_dev, lr, b1, b2, lr_t_dev, s):
        self.launches += 1
        it = _i64(iter_dev, 1)
        it[0] += 1
        _f32(lr_t_dev, 1)[0] = lr * np.sqrt(1.0 - b2 ** int(it[0])) / (1.0 - b1 ** int(it[0]))
        return 0

    def b200ad_adam_multi(self, table, n, chunks, ct, ci, lr_t, b1, b2, eps, gs, lr_t_dev, s):
        self.launches += 1
        if lr_t_dev:
            lr_t = float(_f32(lr_t_dev, 1)[0])
        for p, g, m, v, numel in self._entries(table, n):
            P, G, M, V = _f32(p, numel), _f32(g, numel).astype(np.float64) * gs, _f32(m, numel), _f32(v, numel)
            m64 = M + (1 - b1) * (G - M)
            v64 = V + (1 - b2) * (G * G - V)
            M[:] = m64
            V[:] = v64
            P[:] = P - lr_t * m64 / (np.sqrt(v64) + eps)
        return 0

    def b200ad_sgd_multi(self, table, n, chunks, ct, ci, lr, gs, s):
        self.launches += 1
        for p, g, m, v, numel in self._entries(table, n):
            P = _f32(p, numel)
            P[:] = P - lr * gs * _f32(g, numel).astype(np.float64)
        return 0

    def b200ad_sumsq_multi(self, table, n, chunks, ct, ci, ws, out, s):
        self.launches += 2
        tot = 0.0
        for p, g, m, v, numel in self._entries(table, n):
            tot += float(np.sum(_f32(g, numel).astype(np.float64) ** 2))
        _f32(out, 1)[0] = tot
        return 0

    def b200ad_clip_scale_multi(self, table, n, chunks, ct, ci, norm_sq, max_norm, s):
        self.launches += 1
        rate = max_norm / (np.sqrt(float(_f32(norm_sq, 1)[0])) + 1e-6)
        if rate < 1:
            for p, g, m, v, numel in self._entries(table, n):
                G = _f32(g, numel)
                G[:] = G * rate
        return 0
