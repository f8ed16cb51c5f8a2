"""GPU: the tcgen05 3xTF32 GEMM kernels against an fp64 matmul and against the exact-fp32 CUDA-core GEMM."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    from pinns_rl_pde_b200 import build
    build.build()
    return torch.device("cuda:0")


def _err(Z, ref):
    return float((Z.double() - ref).norm() / ref.norm()), float((Z.double() - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize("M,K,N,C", [(4 * 3001, 128, 128, 4), (5 * 777, 128, 128, 5), (31, 128, 128, 1),
                                     (3 * 4096, 64, 128, 3), (4 * 2048, 128, 256, 4), (4 * 1001, 256, 256, 4),
                                     (5 * 333, 256, 128, 5)])
def test_tc_linear_fwd_matches_fp64(dev, M, K, N, C):
    from pinns_rl_pde_b200 import _lib
    g = torch.Generator(device="cpu").manual_seed(M + K)
    X = (torch.randn(M, K, generator=g) * torch.logspace(-3, 1, K)).to(dev)        # wide dynamic range per feature
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    ref = X.double() @ W.double().t()
    ref[::C] += b.double()
    z_tc = _lib.debug_linear_fwd(X, W, b, C, 1)
    z_sg = _lib.debug_linear_fwd(X, W, b, C, 0)
    torch.cuda.synchronize()
    e_tc, m_tc = _err(z_tc, ref)
    e_sg, m_sg = _err(z_sg, ref)
    print(f"M={M} K={K} N={N}: 3xTF32 rel {e_tc:.2e} (max {m_tc:.2e}); fp32 FFMA rel {e_sg:.2e} (max {m_sg:.2e})")
    assert e_sg < 1e-6
    assert e_tc < 2e-6 and m_tc < 5e-6


@pytest.mark.parametrize("M,K,N", [(4 * 3001, 128, 128), (5 * 777, 128, 128), (17, 128, 128), (4 * 1024, 256, 128),
                                   (4 * 1001, 256, 256), (6 * 500, 128, 256)])
def test_tc_linear_dgrad_matches_fp64(dev, M, K, N):
    from pinns_rl_pde_b200 import _lib
    g = torch.Generator(device="cpu").manual_seed(M + K + 1)
    dZ = (torch.randn(M, N, generator=g) * torch.logspace(-3, 1, N)).to(dev)
    W = (torch.randn(N, K, generator=g) / N ** 0.5).to(dev)
    ref = dZ.double() @ W.double()
    e_tc, m_tc = _err(_lib.debug_linear_dgrad(dZ, W, 1), ref)
    e_sg, _ = _err(_lib.debug_linear_dgrad(dZ, W, 0), ref)
    print(f"dgrad M={M} K={K} N={N}: 3xTF32 rel {e_tc:.2e} (max {m_tc:.2e}); fp32 FFMA rel {e_sg:.2e}")
    assert e_sg < 1e-6 and e_tc < 2e-6 and m_tc < 5e-6


@pytest.mark.parametrize("M,K,N,C", [(4 * 3001, 128, 128, 4), (5 * 777, 128, 128, 5), (8, 128, 128, 1),
                                     (4 * 20000, 128, 128, 4), (4 * 1024, 256, 128, 4), (4 * 1024, 128, 256, 4),
                                     (3 * 5001, 64, 128, 3), (4 * 999, 32, 256, 4)])     # input widths below 128 (Fourier features)
def test_tc_linear_wgrad_matches_fp64(dev, M, K, N, C):
    from pinns_rl_pde_b200 import _lib
    g = torch.Generator(device="cpu").manual_seed(M + K + 2)
    dZ = (torch.randn(M, N, generator=g) * torch.logspace(-2, 1, N)).to(dev)
    X = torch.tanh(torch.randn(M, K, generator=g)).to(dev)
    ref_w = dZ.double().t() @ X.double()
    ref_b = dZ.double()[::C].sum(0)
    dW, db = _lib.debug_linear_wgrad(dZ, X, C, 1)
    dW0, db0 = _lib.debug_linear_wgrad(dZ, X, C, 0)
    e_tc, m_tc = _err(dW, ref_w)
    e_sg, _ = _err(dW0, ref_w)
    print(f"wgrad M={M} K={K} N={N}: 3xTF32 rel {e_tc:.2e} (max {m_tc:.2e}); fp32 FFMA rel {e_sg:.2e}; "
          f"bias rel {_err(db, ref_b)[0]:.2e} / {_err(db0, ref_b)[0]:.2e}")
    assert e_sg < 2e-6 and e_tc < 2e-6 and m_tc < 5e-6
    assert _err(db, ref_b)[0] < 2e-6 and _err(db0, ref_b)[0] < 2e-6


def test_tc_wgrad_long_accumulation(dev):
    """1M points x 4 jet columns: thousands of K-steps per CTA.  The tensor core rounds each accumulate toward zero;
    the kernel's segment flush must keep the gradient at fp32 quality (a single TMEM accumulator drifts by ~1e-4)."""
    from pinns_rl_pde_b200 import _lib
    M, K, N, C = 4 * (1 << 20), 128, 128, 4
    g = torch.Generator(device="cuda").manual_seed(0)
    dZ = torch.randn(M, N, generator=g, device=dev) + 0.25          # non-zero mean: sums do not cancel
    X = torch.tanh(torch.randn(M, K, generator=g, device=dev)) + 0.1
    ref_w = torch.zeros(N, K, dtype=torch.float64, device=dev)
    for i in range(0, M, 1 << 18):
        ref_w += dZ[i:i + (1 << 18)].double().t() @ X[i:i + (1 << 18)].double()
    ref_b = dZ[::C].double().sum(0)
    dW, db = _lib.debug_linear_wgrad(dZ, X, C, 1)
    e_tc, m_tc = _err(dW, ref_w)
    print(f"wgrad 4M rows: 3xTF32 rel {e_tc:.2e} (max {m_tc:.2e}); bias rel {_err(db, ref_b)[0]:.2e}")
    assert e_tc < 2e-6 and m_tc < 5e-6 and _err(db, ref_b)[0] < 2e-6


def test_tc_linear_fwd_throughput(dev, capsys):
    """Not a pass/fail timing test: prints the numbers a builder wants next to the parity result."""
    from pinns_rl_pde_b200 import _lib
    M, K, N, C = 4 * (1 << 18), 128, 128, 4
    X = torch.randn(M, K, device=dev)
    W = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev)
    for mode, name in ((1, "tcgen05 3xTF32"), (0, "CUDA-core fp32")):
        for _ in range(3):
            _lib.debug_linear_fwd(X, W, b, C, mode)
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            _lib.debug_linear_fwd(X, W, b, C, mode)
        e.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(e) / 10
        with capsys.disabled():
            print(f"\n[fwd {name}] {M}x{K}x{N}: {ms:.3f} ms, {2 * M * K * N / ms / 1e9:.1f} TFLOP/s (algorithmic), "
                  f"{(M * K + M * N) * 4 / ms / 1e6:.0f} GB/s")
    dZ = torch.randn(M, N, device=dev)
    for fn, label, nbytes in ((lambda m: _lib.debug_linear_dgrad(dZ, W, m), "dgrad", (M * K + M * N) * 4),
                              (lambda m: _lib.debug_linear_wgrad(dZ, X, C, m), "wgrad", (M * K + M * N) * 4)):
        for mode, name in ((1, "tcgen05 3xTF32"), (0, "CUDA-core fp32")):
            for _ in range(3):
                fn(mode)
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                fn(mode)
            e.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(e) / 10
            with capsys.disabled():
                print(f"[{label} {name}] {M}x{K}x{N}: {ms:.3f} ms, {2 * M * K * N / ms / 1e9:.1f} TFLOP/s (algorithmic), "
                      f"{nbytes / ms / 1e6:.0f} GB/s")
