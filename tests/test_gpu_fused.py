"""Fused LSTM layer / sampled-softmax kernels vs plain PyTorch fp32."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 6e-2)])
def test_lstm_layer_matches_reference(dtype, tol):
    from parallax_b200.ops.fused import lstm_layer, lstm_layer_reference
    torch.manual_seed(0)
    T, B, E, S, P = 5, 16, 24, 64, 24
    dev = "cuda"
    mk = lambda *s: (torch.randn(*s, device=dev) * 0.3)
    x, Wx, Wh, b, WP = mk(T, B, E), mk(E, 4 * S), mk(P, 4 * S), mk(4 * S), mk(S, P)
    c0, h0 = mk(B, S), mk(B, P)
    gH, gc, gh = mk(T, B, P), mk(B, S), mk(B, P)

    def run(fn, dt):
        args = [t.clone().to(dt).requires_grad_(True) for t in (x, Wx, Wh, b, WP)]
        c = c0.clone().requires_grad_(True)
        h = h0.clone().to(dt).requires_grad_(True)
        H, cT, hT = fn(args[0], args[1], args[2], args[3], args[4], c, h, 1.0)
        loss = (H.float() * gH).sum() + (cT.float() * gc).sum() + (hT.float() * gh).sum()
        loss.backward()
        return [H.float(), cT.float(), hT.float()] + \
            [a.grad.float() for a in args] + [c.grad.float(), h.grad.float()]
    ref = run(lstm_layer_reference, torch.float32)
    got = run(lstm_layer, dtype)
    for r, g in zip(ref, got):
        scale = float(r.abs().max()) + 1e-6
        assert float((r - g).abs().max()) <= tol * scale + tol, \
            (float((r - g).abs().max()), scale)


@pytest.mark.parametrize("tc_fwd", ["0", "1"])
def test_lstm_layer_bf16_tcgen05_backward_path(tc_fwd, monkeypatch):
    """B=128, 4S multiple of 1024: the recurrent backward product runs on the
    tcgen05 split-K kernel with the fused addend; with PARALLAX_LSTM_TC_FWD=1 the
    forward step runs on the tcgen05 kernel with the LSTM cell fused in its
    epilogue (gate-interleaved layout)."""
    monkeypatch.setenv("PARALLAX_LSTM_TC_FWD", tc_fwd)
    from parallax_b200.ops.fused import lstm_layer, lstm_layer_reference
    torch.manual_seed(0)
    T, B, E, S, P = 3, 128, 64, 256, 64
    mk = lambda *s: (torch.randn(*s, device="cuda") * 0.2)
    x, Wx, Wh, b, WP = mk(T, B, E), mk(E, 4 * S), mk(P, 4 * S), mk(4 * S), mk(S, P)
    c0, h0, gH = mk(B, S), mk(B, P), mk(T, B, P)

    def run(fn, dt):
        args = [t.clone().to(dt).requires_grad_(True) for t in (x, Wx, Wh, b, WP)]
        c = c0.clone().requires_grad_(True)
        h = h0.clone().to(dt).requires_grad_(True)
        H, cT, hT = fn(args[0], args[1], args[2], args[3], args[4], c, h, 1.0)
        (H.float() * gH).sum().backward()
        return [H.float()] + [a.grad.float() for a in args] + [h.grad.float()]
    ref = run(lstm_layer_reference, torch.float32)
    got = run(lstm_layer, torch.bfloat16)
    for r, g in zip(ref, got):
        scale = float(r.detach().abs().max()) + 1e-6
        assert float((r - g).detach().abs().max()) <= 6e-2 * scale + 6e-2


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 4e-2)])
@pytest.mark.parametrize("S", [100, 1024, 8192])
def test_sampled_softmax_matches_reference(dtype, tol, S):
    from parallax_b200.ops.fused import sampled_softmax_loss, sampled_softmax_reference
    torch.manual_seed(1)
    N, P, V = 64, 32, 5000
    dev = "cuda"
    inputs = torch.randn(N, P, device=dev) * 0.5
    tw, sw = torch.randn(N, P, device=dev) * 0.5, torch.randn(S, P, device=dev) * 0.5
    tb, sb = torch.randn(N, device=dev), torch.randn(S, device=dev)
    lqt, lqs = torch.randn(N, device=dev), torch.randn(S, device=dev)
    targets = torch.randint(0, V, (N,), device=dev)
    sampled = torch.randint(0, V, (S,), device=dev)
    sampled[:5] = targets[:5]                      # accidental hits
    w = torch.rand(N, device=dev)

    def run(fn, dt):
        a = [t.clone().to(dt).requires_grad_(True) for t in (inputs, tw, sw)]
        b = [t.clone().requires_grad_(True) for t in (tb, sb)]
        loss = fn(a[0], a[1], a[2], b[0], b[1], lqt, lqs, targets, sampled)
        (loss.float() * w).sum().backward()
        return [loss.float()] + [t.grad.float() for t in a + b]
    ref = run(sampled_softmax_reference, torch.float32)
    got = run(sampled_softmax_loss, dtype)
    for r, g in zip(ref, got):
        scale = float(r.abs().max()) + 1e-6
        assert float((r - g).abs().max()) <= tol * scale + tol
