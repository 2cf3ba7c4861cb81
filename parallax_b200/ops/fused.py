"""Fused model-side ops (autograd Functions) for the LM1B hot path.

`lstm_layer`  — a whole unrolled LSTMP layer as ONE autograd node: the
  sequential part per time step is 2 small GEMMs + 1 fused cell kernel
  (forward) and 2 GEMMs + 1 fused kernel (backward); every weight gradient is
  a single GEMM batched over all time steps (the reference's TF graph issues
  one small GEMM + ~25 elementwise kernels per step and direction:
  `examples/lm1b/language_model.py:76-87`).
`sampled_softmax_loss` — logits GEMM + one fused kernel that produces the
  loss and the softmax probabilities in place (= gradient wrt logits), so the
  backward is two GEMMs and a few row scalings.

Both have a pure-PyTorch implementation (`*_reference`) which is the numerics
oracle in the tests and the path used on the host fabric.
"""
import ctypes

import torch

from . import lib as _lib, check as _check, register_signatures

_vp, _i, _f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
register_signatures({
    "px_lstm_cell_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _f, _i, _vp]),
    "px_lstm_cell_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "px_lstm_gates_tc": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp]),
    "px_sampled_softmax": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "px_sampled_softmax_dot": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i,
                                    _vp]),
    "px_ssm_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _f, _vp, _vp, _vp, _i, _vp, _i, _i, _i,
                        _vp]),
})
_DT = {torch.float32: 0, torch.bfloat16: 1}


def _p(t):
    return _vp(t.data_ptr())


def _stream():
    return _vp(torch.cuda.current_stream().cuda_stream)


def _count(n=1):
    from ..parallel import nvops
    nvops.launches["n"] += n


# ===========================================================================
# LSTM layer
# ===========================================================================
def lstm_layer_reference(x, Wx, Wh, bias, W_P, c0, h0, forget_bias=1.0):
    """x: [T, B, E] → (H [T, B, P], c_T, h_T); plain autograd-able torch."""
    T, Bsz, E = x.shape
    S = W_P.shape[0]
    xw = torch.addmm(bias, x.reshape(T * Bsz, E), Wx).view(T, Bsz, 4 * S)
    c, h, outs = c0.float(), h0, []
    for t in range(T):
        gates = torch.addmm(xw[t], h, Wh).float()
        i, j, f, o = gates.split(S, dim=1)
        c = torch.sigmoid(f + forget_bias) * c + torch.sigmoid(i) * torch.tanh(j)
        m = (torch.sigmoid(o) * torch.tanh(c)).to(x.dtype)
        h = m @ W_P
        outs.append(h)
    return torch.stack(outs), c, h


_perm_cache = {}


def _tc_forward_enabled():
    """The fused tcgen05 forward step is opt-in (`PARALLAX_LSTM_TC_FWD=1`):
    measured on B200 at the LM1B shape it is numerically equivalent but, with the
    per-step weight re-layout it needs, slower than cuBLAS + the stand-alone cell
    kernel (profiles/README.md); the tcgen05 split-K product on the backward path
    is always on."""
    import os
    return os.environ.get("PARALLAX_LSTM_TC_FWD", "0") == "1"


def gate_interleave_perm(S, device):
    """perm[n'] = n : column n' = tile·128 + g·32 + j of the gate-interleaved
    layout holds original column n = g·S + tile·32 + j (g: i, j, f, o)."""
    key = (S, str(device))
    if key not in _perm_cache:
        tiles = S // 32
        g = torch.arange(4, device=device).view(1, 4, 1)
        t = torch.arange(tiles, device=device).view(tiles, 1, 1)
        j = torch.arange(32, device=device).view(1, 1, 32)
        perm = (g * S + t * 32 + j).reshape(-1)
        inv = torch.empty_like(perm)
        inv[perm] = torch.arange(4 * S, device=device)
        _perm_cache[key] = (perm, inv)
    return _perm_cache[key]


def _bwd_fused_w():
    """Backward chain with the combined weight `Wc = W_P @ Wh` ([S, 4S]):
    ``dm_{t-1} = dH_{t-1} W_P^T + dgates_t Wc^T`` — ONE product per time step on the
    critical path instead of two (``dh = dH + dgates Wh^T`` then ``dm = dh W_P^T``); the
    per-step `dh` values (needed only for dW_P) are produced afterwards by one batched GEMM
    off the critical path.  "tc": tcgen05 split-K kernel, "cublas": torch.addmm, "0": off.
    Measured on B200 at the LM1B shape (tools/bench_lstm_gemms.py): the combined product
    (128 x 2048 x 8192, 32 MB of weights per step from L2) costs 13.5 us (cuBLAS) / 16.2 us
    (tcgen05 split-K) against 2.8 + 11.7 us for the two it replaces — no gain (1.46 / 1.52
    vs 1.43 ms per training step), so it stays OFF by default."""
    import os
    return os.environ.get("PARALLAX_LSTM_BWD_FUSEDW", "0")


def _wpt_side():
    """PARALLAX_LSTM_WPT_SIDE=0: transpose W_P at the head of the backward pass instead."""
    import os
    return os.environ.get("PARALLAX_LSTM_WPT_SIDE", "1") != "0"


def _dbias_stream():
    """PARALLAX_LSTM_DBIAS_STREAM=1: the bias gradient (a bandwidth-bound column sum over
    dgates) runs on a second side stream, next to the compute-bound weight-gradient GEMMs."""
    import os
    return os.environ.get("PARALLAX_LSTM_DBIAS_STREAM", "1") != "0"


def _wgrad_chunks(T):
    """How many pieces the weight-gradient GEMMs are cut into along time so that the
    earlier pieces run on the side stream underneath the (latency-bound) recurrent
    backward chain.  Default 1 (all after the loop, still on the side stream): measured on
    B200, 2 chunks cost 1.253 vs 1.239 ms/step — the big GEMMs slow the chain they overlap."""
    import os
    n = int(os.environ.get("PARALLAX_LSTM_WGRAD_CHUNKS", "1"))
    return max(1, min(n, T // 2 if T >= 4 else 1))


class _LSTMLayerFn(torch.autograd.Function):
    """`W` is either the stacked `[E+P, 4S]` kernel of the reference's LSTM cell
    (`Wh is None`; rows `[:E]` multiply x, rows `[E:]` multiply h — one parameter, one
    gradient written straight into its bucket sink) or just `Wx` with `Wh` separate."""

    @staticmethod
    def forward(ctx, x, W, Wh, bias, W_P, c0, h0, forget_bias):
        L = _lib()
        T, Bsz, E = x.shape
        S, P = W_P.shape
        dt = x.dtype
        dev = x.device
        x = x.contiguous()
        stacked = Wh is None
        ctx.stacked = stacked
        ctx.param_refs = (W, bias, W_P)
        if stacked:
            Wx, Wh = W.detach()[:E], W.detach()[E:]
        else:
            Wx = W
        # tcgen05 path: recurrent GEMM with the LSTM cell fused into its epilogue
        # (gate-interleaved column layout, see gemm_tc.cu)
        tc = (dt == torch.bfloat16 and Bsz % 128 == 0 and P % 64 == 0 and S % 32 == 0 and
              _tc_forward_enabled())
        if tc:
            perm, inv = gate_interleave_perm(S, dev)
            Wx_l = Wx.index_select(1, perm)
            Wh_l = Wh.index_select(1, perm)                 # [P, 4S]  (K-contiguous for bwd)
            WhT = Wh_l.t().contiguous()                     # [4S, P]  (K-contiguous for fwd)
            bias_l = bias.index_select(0, perm)
        else:
            Wx_l, Wh_l, bias_l, WhT = Wx, Wh, bias, None
        # W_P^T for the backward chain (dm_t = dh_t W_P^T as a plain NN GEMM): a 2 MB true
        # transpose, 17 us of uncoalesced copy — done here on the side stream, underneath
        # the forward chain, instead of at the head of the backward pass
        ctx.WPT = None
        if dev.type == "cuda" and any(ctx.needs_input_grad) and _wpt_side():
            from . import sinks
            cur = torch.cuda.current_stream(dev)
            ws = sinks.side_stream(dev)
            ws.wait_stream(cur)
            with torch.cuda.stream(ws):
                ctx.WPT = W_P.detach().t().contiguous()
                ctx.WPT_ev = torch.cuda.Event()
                ctx.WPT_ev.record(ws)
            ctx.WPT.record_stream(cur)
        xw = torch.addmm(bias_l, x.view(T * Bsz, E), Wx_l).view(T, Bsz, 4 * S)
        act = torch.empty(T, Bsz, 4 * S, dtype=dt, device=dev)
        c_all = torch.empty(T + 1, Bsz, S, dtype=torch.float32, device=dev)
        m_all = torch.empty(T, Bsz, S, dtype=dt, device=dev)
        h_all = torch.empty(T + 1, Bsz, P, dtype=dt, device=dev)
        c_all[0].copy_(c0)
        h_all[0].copy_(h0)
        st = _stream()
        if tc:
            for t in range(T):
                _check(L.px_lstm_gates_tc(_p(h_all[t]), _p(WhT), _p(xw[t]), _p(c_all[t]),
                                          _p(c_all[t + 1]), _p(m_all[t]), _p(act[t]), Bsz, S, P,
                                          float(forget_bias), st), "lstm_gates_tc")
                torch.mm(m_all[t], W_P, out=h_all[t + 1])
        else:
            for t in range(T):
                # accumulate straight into xw[t] (an `out=` different from the addend makes
                # torch copy the 2 MB addend first — one more launch per step on the
                # critical path)
                gpre = xw[t].addmm_(h_all[t], Wh_l)
                _check(L.px_lstm_cell_fwd(_p(gpre), _p(c_all[t]), _p(act[t]),
                                          _p(c_all[t + 1]), _p(m_all[t]), Bsz, S,
                                          float(forget_bias), _DT[dt], st), "lstm_cell_fwd")
                torch.mm(m_all[t], W_P, out=h_all[t + 1])
        _count(T)
        ctx.save_for_backward(x, Wx_l, Wh_l, W_P, act, c_all, m_all, h_all)
        ctx.dims = (T, Bsz, E, S, P)
        ctx.tc = tc
        ctx.Wc = None
        if (_bwd_fused_w() != "0" and dt == torch.bfloat16 and T > 1 and
                torch.is_grad_enabled() is False and any(ctx.needs_input_grad)):
            # combined weight for the backward chain, computed underneath the forward chain
            from . import sinks
            cur = torch.cuda.current_stream(dev)
            ws = sinks.side_stream(dev)
            ws.wait_stream(cur)
            with torch.cuda.stream(ws):
                ctx.Wc = torch.mm(W_P.detach(), Wh_l.detach())          # [S, 4S]
                ctx.Wc_ev = torch.cuda.Event()
                ctx.Wc_ev.record(ws)
            ctx.Wc.record_stream(cur)
        return h_all[1:], c_all[T].clone(), h_all[T].clone()

    @staticmethod
    def backward(ctx, dH, dcT, dhT):
        from . import sinks
        L = _lib()
        x, Wx, Wh, W_P, act, c_all, m_all, h_all = ctx.saved_tensors   # Wx/Wh: layout of `act`
        T, Bsz, E, S, P = ctx.dims
        tc, stacked = ctx.tc, ctx.stacked
        W_ref, bias_ref, WP_ref = ctx.param_refs
        dt, dev = x.dtype, x.device
        dH = dH.contiguous()
        dgates = torch.empty(T, Bsz, 4 * S, dtype=dt, device=dev)
        dh_tot = torch.empty(T, Bsz, P, dtype=dt, device=dev)
        dc = torch.zeros(Bsz, S, dtype=torch.float32, device=dev) if dcT is None \
            else dcT.float().clone()
        dh_rec = None if dhT is None else dhT.to(dt)
        dm = torch.empty(Bsz, S, dtype=dt, device=dev)
        # dm_t = dh_t @ W_P^T runs as a plain NN GEMM on a materialised W_P^T.
        if ctx.WPT is not None:
            torch.cuda.current_stream(dev).wait_event(ctx.WPT_ev)
            WPT = ctx.WPT
        else:
            WPT = W_P.t().contiguous()
        # dh_{t-1} = dH_{t-1} + dgates_t @ Wh^T : Wh [P, 4S] is already the
        # K-contiguous "B^T" operand, so this skinny product (M=B, N=P, K=4S)
        # goes to our tcgen05 split-K kernel with the +dH addend fused in.
        from . import gemm as _gemm
        use_tc = (dt == torch.bfloat16 and Bsz % 128 == 0 and P % 64 == 0 and
                  (4 * S) % 1024 == 0 and Wh.is_contiguous())
        # K-splits of the dh product: 8 CTAs per tile reducing through DSMEM in a cluster
        # (8.7 us at the LM1B shape) or 16 through the L2 workspace (11.5 us)
        ksp = 8 if _gemm.cluster_default() else 16
        WhT = None if use_tc else Wh.t().contiguous()
        st = _stream()

        # ---- weight-gradient outputs: the parameters' bucket sinks when a dense group
        # registered them (no pack copy afterwards), else fresh tensors
        def sink_of(ref, shape):
            s_ = None if tc else sinks.get(ref)
            if s_ is not None and s_.dtype == dt and tuple(s_.shape) == tuple(shape):
                return s_, True
            return torch.empty(shape, dtype=dt, device=dev), False
        if stacked:
            dW, w_sunk = sink_of(W_ref, (E + P, 4 * S))
            dWx_o, dWh_o = dW[:E], dW[E:]
        else:
            dWx_o, w_sunk = sink_of(W_ref, (E, 4 * S))
            dWh_o = torch.empty(P, 4 * S, dtype=dt, device=dev)
        dbias_o, b_sunk = sink_of(bias_ref, (4 * S,))
        dWP_o, p_sunk = sink_of(WP_ref, (S, P))
        # ---- weight-gradient GEMMs on the side stream, cut in time chunks so the earlier
        # ones run underneath the recurrent chain
        cur = torch.cuda.current_stream(dev)
        ws = sinks.side_stream(dev)
        nchunks = _wgrad_chunks(T)
        bounds = [round(i * T / nchunks) for i in range(nchunks + 1)]   # over t, ascending
        state = {"first": True}
        ws2 = sinks.side_stream(dev, 1) if (nchunks == 1 and _dbias_stream()) else None

        def wgrad(lo, hi):
            """accumulate the contribution of steps [lo, hi) (their dgates / dh_tot are final)"""
            ws.wait_stream(cur)
            with torch.cuda.stream(ws):
                dg = dgates[lo:hi].view(-1, 4 * S)
                hT = h_all[lo:hi].reshape(-1, P).t()
                xT = x[lo:hi].reshape(-1, E).t()
                mT = m_all[lo:hi].reshape(-1, S).t()
                dh2 = dh_tot[lo:hi].view(-1, P)
                if state["first"]:
                    torch.mm(hT, dg, out=dWh_o)
                    torch.mm(xT, dg, out=dWx_o)
                    if ws2 is None:
                        torch.sum(dg, 0, out=dbias_o)
                    torch.mm(mT, dh2, out=dWP_o)
                    state["first"] = False
                else:
                    dWh_o.addmm_(hT, dg)
                    dWx_o.addmm_(xT, dg)
                    dbias_o.add_(dg.sum(0))
                    dWP_o.addmm_(mT, dh2)
        pending_hi = T
        Wc = ctx.Wc
        mode = _bwd_fused_w()
        if dh_rec is None:
            dh_tot[T - 1].copy_(dH[T - 1])
        else:
            torch.add(dH[T - 1], dh_rec, out=dh_tot[T - 1])
        if Wc is not None:
            cur.wait_event(ctx.Wc_ev)
            fused_tc = (mode == "tc" and Bsz % 128 == 0 and S % 64 == 0 and (4 * S) % 512 == 0)
            WcT = None if fused_tc else Wc.t()
            WhT_v = Wh.t()
            # DMH[t] = dH[t] W_P^T for every step at once (dh_tot[T-1] carries dhT)
            DMH = torch.empty(T, Bsz, S, dtype=dt, device=dev)
            torch.mm(dH[:T - 1].reshape(-1, P), WPT, out=DMH[:T - 1].view(-1, S))
            torch.mm(dh_tot[T - 1], WPT, out=DMH[T - 1])

            def dh_chunk(lo, hi):
                """dh_tot[lo:hi] (off the critical path; steps whose dgates[t+1] is final)"""
                top = min(hi, T - 1)
                if top > lo:
                    torch.addmm(dH[lo:top].reshape(-1, P),
                                dgates[lo + 1:top + 1].view(-1, 4 * S), WhT_v,
                                out=dh_tot[lo:top].view(-1, P))
            dm_t = DMH[T - 1]
            for t in range(T - 1, -1, -1):
                _check(L.px_lstm_cell_bwd(_p(dm_t), _p(dc), _p(act[t]), _p(c_all[t]),
                                          _p(c_all[t + 1]), _p(dgates[t]), Bsz, S, _DT[dt],
                                          1 if tc else 0, st), "lstm_cell_bwd")
                if t > 0:
                    if fused_tc:
                        _gemm.gemm_tn(dgates[t], Wc, addend=DMH[t - 1], splits=4, bn=64,
                                      out=dm)
                    else:
                        torch.addmm(DMH[t - 1], dgates[t], WcT, out=dm)
                    dm_t = dm
                if t in bounds[1:-1]:
                    ws.wait_stream(cur)
                    with torch.cuda.stream(ws):
                        dh_chunk(t, pending_hi)
                    wgrad(t, pending_hi)
                    pending_hi = t
            dh_rec = _gemm.gemm_tn(dgates[0], Wh, splits=ksp, bn=64) if use_tc \
                else torch.mm(dgates[0], WhT_v)
            ws.wait_stream(cur)
            with torch.cuda.stream(ws):
                dh_chunk(0, pending_hi)
            for t_ in (DMH, dH):
                t_.record_stream(ws)
        else:
            for t in range(T - 1, -1, -1):
                torch.mm(dh_tot[t], WPT, out=dm)
                _check(L.px_lstm_cell_bwd(_p(dm), _p(dc), _p(act[t]), _p(c_all[t]),
                                          _p(c_all[t + 1]), _p(dgates[t]), Bsz, S, _DT[dt],
                                          1 if tc else 0, st), "lstm_cell_bwd")
                if t > 0:
                    if use_tc:
                        _gemm.gemm_tn(dgates[t], Wh, addend=dH[t - 1], splits=ksp, bn=64,
                                      out=dh_tot[t - 1])
                    else:
                        torch.addmm(dH[t - 1], dgates[t], WhT, out=dh_tot[t - 1])
                else:
                    dh_rec = _gemm.gemm_tn(dgates[0], Wh, splits=ksp, bn=64) if use_tc \
                        else torch.mm(dgates[0], WhT)
                if t in bounds[1:-1]:
                    wgrad(t, pending_hi)
                    pending_hi = t
        _count(T)
        dg2 = dgates.view(T * Bsz, 4 * S)
        # dx first: the embedding gradient is what the rest of backward (and the sparse
        # push) is waiting for; the last weight-gradient chunk goes to the side stream
        ev_b = None
        if ws2 is not None:
            # bandwidth-bound column sum: runs next to the compute-bound GEMMs below
            ws2.wait_stream(cur)
            with torch.cuda.stream(ws2):
                torch.sum(dg2, 0, out=dbias_o)
            ev_b = torch.cuda.Event()
            ev_b.record(ws2)
            dgates.record_stream(ws2)
            dbias_o.record_stream(ws2)
        dx = (dg2 @ Wx.t()).view(T, Bsz, E)
        wgrad(0, pending_hi)
        ev = torch.cuda.Event()
        ev.record(ws)
        if ev_b is None:
            ev_b = ev
        for t_ in (dgates, dh_tot, x, h_all, m_all, dWh_o, dWx_o, dbias_o, dWP_o):
            t_.record_stream(ws)
        outs, plain = [], False
        for ref, o, sunk, e_ in ((W_ref, dW if stacked else dWx_o, w_sunk, ev),
                                 (bias_ref, dbias_o, b_sunk, ev_b), (WP_ref, dWP_o, p_sunk, ev)):
            if sunk:
                # in the bucket already: hand it to the dense group directly (its fused
                # reduce/update kernel waits for the event), nothing goes through
                # AccumulateGrad
                sinks.deliver(ref, e_)
                outs.append(None)
            else:
                outs.append(o)
                plain = True
        if plain or not stacked:
            cur.wait_event(ev)        # plain tensors are consumed on the current stream
            if ev_b is not ev:
                cur.wait_event(ev_b)
        dW_out, dbias, dW_P = outs
        dWh = None if stacked else dWh_o
        if tc:          # back to the caller's (plain) gate-column order
            _, inv = gate_interleave_perm(S, dev)
            if stacked:
                dW_out = dW_out.index_select(1, inv)      # (tc path never uses sinks)
            else:
                dW_out, dWh = dW_out.index_select(1, inv), dWh.index_select(1, inv)
            dbias = dbias.index_select(0, inv)
        return dx, dW_out, dWh, dbias, dW_P, dc, dh_rec, None


def lstm_layer(x, Wx, Wh, bias, W_P, c0, h0, forget_bias=1.0):
    if x.is_cuda and x.dtype in _DT:
        return _LSTMLayerFn.apply(x, Wx, Wh, bias, W_P, c0, h0, forget_bias)
    return lstm_layer_reference(x, Wx, Wh, bias, W_P, c0, h0, forget_bias)


def lstm_layer_stacked(x, W, bias, W_P, c0, h0, forget_bias=1.0):
    """Same layer with the reference's stacked kernel `W = [Wx; Wh]` ([E+P, 4S],
    `examples/lm1b/language_model.py:76-87`) passed whole: one parameter, one gradient
    tensor, written by the weight-gradient GEMMs directly into the parameter's bucket."""
    E = x.shape[-1]
    if x.is_cuda and x.dtype in _DT:
        return _LSTMLayerFn.apply(x, W, None, bias, W_P, c0, h0, forget_bias)
    return lstm_layer_reference(x, W[:E], W[E:], bias, W_P, c0, h0, forget_bias)


# ===========================================================================
# sampled softmax
# ===========================================================================
def sampled_softmax_reference(inputs, true_w, samp_w, true_b, samp_b, logq_true,
                              logq_samp, targets, sampled):
    """Per-example loss [N] (tf.nn.sampled_softmax_loss semantics)."""
    true_logits = (inputs * true_w).sum(-1).float() + true_b.float() - logq_true
    samp_logits = (inputs @ samp_w.t()).float() + (samp_b.float() - logq_samp)
    hits = targets.unsqueeze(1) == sampled.unsqueeze(0)
    samp_logits = samp_logits.masked_fill(hits, float("-inf"))
    lse = torch.logsumexp(torch.cat([true_logits.unsqueeze(1), samp_logits], 1), 1)
    return lse - true_logits


class _SampledSoftmaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inputs, true_w, samp_w, true_b, samp_b, logq_true, logq_samp,
                targets, sampled):
        L = _lib()
        N, S = inputs.shape[0], samp_w.shape[0]
        dt = inputs.dtype
        logits = inputs @ samp_w.t()                               # [N, S]  (cuBLAS)
        true_dot = (inputs.float() * true_w.float()).sum(-1)
        adj_t = (true_b.float() - logq_true).contiguous()
        adj_s = (samp_b.float() - logq_samp).contiguous()
        loss = torch.empty(N, dtype=torch.float32, device=inputs.device)
        dtrue = torch.empty(N, dtype=torch.float32, device=inputs.device)
        tg = targets.to(torch.int64).contiguous()
        sm = sampled.to(torch.int64).contiguous()
        _check(L.px_sampled_softmax(_p(logits), _p(true_dot), _p(adj_t), _p(adj_s), _p(tg),
                                    _p(sm), _p(loss), _p(dtrue), N, S, _DT[dt], _stream()),
               "sampled_softmax")
        _count(1)
        ctx.save_for_backward(inputs, true_w, samp_w, logits, dtrue)
        return loss

    @staticmethod
    def backward(ctx, g):
        inputs, true_w, samp_w, probs, dtrue = ctx.saved_tensors
        dt = inputs.dtype
        g = g.float()
        gt = (g * dtrue).unsqueeze(1)                              # [N,1]
        d_inputs = ((probs @ samp_w).float() * g.unsqueeze(1) +
                    gt * true_w.float()).to(dt)
        gi = (inputs.float() * g.unsqueeze(1)).to(dt)              # diag(g)·H
        d_samp_w = probs.t() @ gi
        d_true_w = (gt * inputs.float()).to(dt)
        d_samp_b = (probs.float().t() @ g) if probs.dtype == torch.float32 else \
            (probs.t() @ g.to(dt).unsqueeze(1)).squeeze(1).float()
        d_true_b = gt.squeeze(1)
        return d_inputs, d_true_w, d_samp_w, d_true_b, d_samp_b, None, None, None, None


def sampled_softmax_loss(inputs, true_w, samp_w, true_b, samp_b, logq_true, logq_samp,
                         targets, sampled):
    if inputs.is_cuda and inputs.dtype in _DT and samp_w.shape[0] <= 256 * 64:
        return _SampledSoftmaxFn.apply(inputs, true_w.to(inputs.dtype),
                                       samp_w.to(inputs.dtype), true_b, samp_b,
                                       logq_true, logq_samp, targets, sampled)
    return sampled_softmax_reference(inputs, true_w, samp_w, true_b, samp_b, logq_true,
                                     logq_samp, targets, sampled)


# ---------------------------------------------------------------------------
# the whole loss head as one node
# ---------------------------------------------------------------------------
def _head_enabled():
    """PARALLAX_SSM_HEAD=0 falls back to `sampled_softmax_loss` + PyTorch glue."""
    import os
    return os.environ.get("PARALLAX_SSM_HEAD", "1") != "0"


def _head_forward(inputs, w_all, adj, targets, sampled):
    """-> (probs [N,S] in inputs.dtype, loss [N] fp32, dtrue [N] fp32).  `w_all` holds the
    N true-class rows followed by the S sampled rows; `adj` = bias - log Q for the same
    N + S positions."""
    N, S = inputs.shape[0], w_all.shape[0] - inputs.shape[0]
    P = inputs.shape[1]
    logits = inputs @ w_all[N:].t()                                # [N, S]
    if inputs.is_cuda:
        loss = torch.empty(N, dtype=torch.float32, device=inputs.device)
        dtrue = torch.empty(N, dtype=torch.float32, device=inputs.device)
        _check(_lib().px_sampled_softmax_dot(
            _p(logits), _p(inputs), _p(w_all), P, _p(adj), _vp(adj.data_ptr() + 4 * N),
            _p(targets), _p(sampled), _p(loss), _p(dtrue), N, S, _DT[inputs.dtype], _stream()),
            "sampled_softmax_dot")
        _count(1)
        return logits, loss, dtrue
    # same math in PyTorch (host fabric / CPU tests of the node's algebra)
    tl = (inputs.float() * w_all[:N].float()).sum(-1) + adj[:N]
    a = logits.float() + adj[N:]
    a = a.masked_fill(targets.unsqueeze(1) == sampled.unsqueeze(0), float("-inf"))
    mx = torch.maximum(a.max(1).values, tl)
    e, et = torch.exp(a - mx.unsqueeze(1)), torch.exp(tl - mx)
    denom = e.sum(1) + et
    return (e / denom.unsqueeze(1)).to(inputs.dtype), mx + torch.log(denom) - tl, et / denom - 1.0


def _head_backward(G, inputs, w_all, g, row_w, dtrue, gi, d_w_all, db, grow):
    """in place: G -> d_inputs; fills gi, d_w_all[:N], db[:N], grow (see `px_ssm_bwd`)."""
    N, P = inputs.shape
    if inputs.is_cuda:
        _check(_lib().px_ssm_bwd(
            _p(G), _p(inputs), _p(w_all), _p(g), 0 if g.numel() == 1 else 1,
            _p(row_w) if row_w is not None else None, _p(dtrue), 1.0 / N, _p(gi), _p(d_w_all),
            _p(db), 1 if db.dtype == torch.bfloat16 else 0, _p(grow), N, P, _DT[inputs.dtype],
            _stream()), "ssm_bwd")
        _count(1)
        return
    gr = g.reshape(-1).float() * (1.0 / N)
    gr = gr.expand(N) if gr.numel() == 1 else gr
    if row_w is not None:
        gr = gr * row_w
    gt = gr * dtrue
    x, wt = inputs.float(), w_all[:N].float()
    G.copy_(G.float() * gr.unsqueeze(1) + gt.unsqueeze(1) * wt)
    gi.copy_(x * gr.unsqueeze(1))
    d_w_all[:N].copy_(x * gt.unsqueeze(1))
    db[:N].copy_(gt)
    grow.copy_(gr)


class _SampledSoftmaxHeadFn(torch.autograd.Function):
    """mean (optionally row-weighted) sampled-softmax loss of `inputs` [N, P] against the
    rows `w_all` / `b_all` of (targets ++ sampled).  Forward: logits GEMM + one kernel
    (true-class dot product, bias - log Q, accidental hits, log-sum-exp, loss, softmax
    probabilities in place).  Backward: two GEMMs, one GEMV and ONE glue kernel; the
    gradients of the looked-up rows come out as single [N+S, ·] tensors in lookup order —
    exactly what the co-lookup group's push kernel consumes (no slice / cat / cast
    launches in between).  Reference: tf.nn.sampled_softmax_loss as used by
    `examples/lm1b/language_model.py:96-107`."""

    @staticmethod
    def forward(ctx, inputs, w_all, b_all, adj, targets, sampled, row_w):
        probs, loss, dtrue = _head_forward(inputs, w_all, adj, targets, sampled)
        ctx.save_for_backward(inputs, w_all, probs, dtrue, row_w)
        ctx.b_meta = (tuple(b_all.shape), b_all.dtype)
        if row_w is not None:
            loss = loss * row_w
        return loss.mean()

    @staticmethod
    def backward(ctx, g):
        inputs, w_all, probs, dtrue, row_w = ctx.saved_tensors
        N, P = inputs.shape
        S = w_all.shape[0] - N
        dt, dev = inputs.dtype, inputs.device
        b_shape, b_dt = ctx.b_meta
        g = g.float().contiguous()
        G = probs @ w_all[N:]                                      # [N, P] -> d_inputs
        d_w_all = torch.empty(N + S, P, dtype=dt, device=dev)
        db = torch.empty(N + S, dtype=b_dt if b_dt in _DT else torch.float32, device=dev)
        gi = torch.empty(N, P, dtype=dt, device=dev)
        grow = torch.empty(N, dtype=dt, device=dev)
        _head_backward(G, inputs, w_all, g, row_w, dtrue, gi, d_w_all, db, grow)
        torch.mm(probs.t(), gi, out=d_w_all[N:])                   # d w_sampled
        if db.dtype == dt:                                         # d b_sampled
            torch.mm(probs.t(), grow.view(N, 1), out=db[N:].view(S, 1))
        else:
            db[N:].copy_((probs.t() @ grow.view(N, 1)).view(S))
        db = db.view(b_shape)
        if db.dtype != b_dt:
            db = db.to(b_dt)
        return G, d_w_all, db, None, None, None, None


def sampled_softmax_head(inputs, w_all, b_all, logq, targets, sampled, row_w=None, adj=None):
    """Scalar mean sampled-softmax loss.  `w_all` [N+S, P], `b_all` [N+S] or [N+S, 1] and
    `logq` [N+S] are ordered (targets ++ sampled); `adj` may carry a precomputed
    ``b - log Q`` (fp32, no gradient — the bias gradient is produced from `b_all`)."""
    N, P = inputs.shape
    S = w_all.shape[0] - N
    dt = inputs.dtype
    vec = 4 if dt == torch.float32 else 8
    fused = (inputs.is_cuda and _head_enabled() and dt in _DT and w_all.dtype == dt and
             0 < S <= 256 * 64 and P % vec == 0 and b_all.numel() == N + S)
    if fused:
        inputs = inputs.contiguous()
        w_all = w_all.contiguous()
        fused = inputs.data_ptr() % 16 == 0 and w_all.data_ptr() % 16 == 0
    if not fused:
        b = b_all.reshape(-1)
        loss = sampled_softmax_loss(inputs, w_all[:N], w_all[N:], b[:N], b[N:], logq[:N],
                                    logq[N:], targets, sampled)
        if row_w is not None:
            loss = loss * row_w.to(loss.dtype)
        return loss.mean()
    if adj is None:
        adj = b_all.detach().reshape(-1).float() - logq
    adj = adj.float().contiguous()
    tg = targets.to(torch.int64).contiguous()
    sm = sampled.to(torch.int64).contiguous()
    rw = None if row_w is None else row_w.detach().float().contiguous()
    return _SampledSoftmaxHeadFn.apply(inputs, w_all, b_all, adj, tg, sm, rw)
