"""GPU: the tcgen05 / TMEM / TMA GEMM (pvae_gemm_tf32) for every operand-major combination the
lin|lin P-VAE needs (forward, dgrad, wgrad), ragged sizes and split-K."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from poissonvae_b200 import _lib
    return _lib.load()


def run(lib, A, B, M, N, Kd, a_k, b_k, splits=1, precise=0):
    C = torch.full((M, N), float("nan"), device="cuda")
    ws = torch.empty(max(1, lib.pvae_gemm_ws_floats(M, N, splits)), device="cuda")
    rc = lib.pvae_gemm_tf32(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, Kd, a_k, b_k, precise, splits, ws.data_ptr(),
                            torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.pvae_last_error()
    torch.cuda.synchronize()
    return C


def operands(M, N, Kd, a_k, b_k, gen, integers):
    if integers:
        A2 = torch.randint(-4, 5, (M, Kd), device="cuda", generator=gen).float()
        B2 = torch.randint(-4, 5, (N, Kd), device="cuda", generator=gen).float()
    else:
        A2 = torch.randn(M, Kd, device="cuda", generator=gen)
        B2 = torch.randn(N, Kd, device="cuda", generator=gen)
    A = A2 if a_k else A2.t().contiguous()  # stored (Kd, M) when MN-major
    B = B2 if b_k else B2.t().contiguous()
    return A2, B2, A, B


SHAPES = [(128, 128, 32), (128, 64, 64), (4096, 512, 256), (4096, 256, 512), (12, 64, 256), (520, 136, 100),
          (256, 512, 1024), (8, 128, 4)]


@pytest.mark.parametrize("majors", [(1, 1), (1, 0), (0, 0), (0, 1)])
@pytest.mark.parametrize("shape", SHAPES)
def test_exact_on_tf32_representable_inputs(lib, majors, shape):
    """Small integers are exact in TF32 and the sums stay below 2^24: any layout / descriptor mistake shows."""
    M, N, Kd = shape
    gen = torch.Generator("cuda").manual_seed(M + N + Kd)
    A2, B2, A, B = operands(M, N, Kd, *majors, gen, integers=True)
    for precise in (0, 1):
        C = run(lib, A, B, M, N, Kd, *majors, precise=precise)
        assert torch.equal(C, (A2.double() @ B2.double().t()).float()), precise


@pytest.mark.parametrize("majors", [(1, 1), (1, 0), (0, 0), (0, 1)])
@pytest.mark.parametrize("shape", [(1024, 512, 256), (256, 512, 4096), (200, 264, 520)])
def test_3xtf32_is_fp32_grade(lib, majors, shape):
    """Operand splitting (hi*hi + lo*hi + hi*lo): error at the level of an fp32 GEMM, ~1000x below single-pass TF32."""
    M, N, Kd = shape
    gen = torch.Generator("cuda").manual_seed(17)
    A2, B2, A, B = operands(M, N, Kd, *majors, gen, integers=False)
    C = run(lib, A, B, M, N, Kd, *majors, precise=1, splits=4 if Kd >= 4096 else 1)
    ref = A2.double() @ B2.double().t()
    rel = ((C.double() - ref).norm() / ref.norm()).item()
    C1 = run(lib, A, B, M, N, Kd, *majors, precise=0, splits=4 if Kd >= 4096 else 1)
    rel1 = ((C1.double() - ref).norm() / ref.norm()).item()
    fp32 = (((A2 @ B2.t()).double() - ref).norm() / ref.norm()).item()
    print(f"rel err: 3xTF32 {rel:.2e}  TF32 {rel1:.2e}  torch fp32 {fp32:.2e}")
    # products are good to ~2^-22; what remains is the tensor core's truncating fp32 accumulation
    # (one-sided, ~0.5 ulp of the running sum per K=8 step)
    assert rel < 3e-6 and rel < rel1 / 50, (rel, rel1, fp32)
    bound = 2.0 ** -22 * (A2.abs().double() @ B2.abs().double().t()) + (Kd / 8) * 2.0 ** -24 * ref.abs().max()
    assert ((C.double() - ref).abs() <= bound).all()


@pytest.mark.parametrize("majors", [(1, 1), (1, 0), (0, 0)])
def test_tf32_error_bound(lib, majors):
    M, N, Kd = 1024, 512, 256
    gen = torch.Generator("cuda").manual_seed(5)
    A2, B2, A, B = operands(M, N, Kd, *majors, gen, integers=False)
    C = run(lib, A, B, M, N, Kd, *majors)
    ref = A2.double() @ B2.double().t()
    bound = 2.0 ** -10 * (A2.abs().double() @ B2.abs().double().t())  # operands truncated to 10 mantissa bits
    assert ((C.double() - ref).abs() <= bound).all()
    rel = ((C.double() - ref).norm() / ref.norm()).item()
    assert rel < 1e-3, rel


@pytest.mark.parametrize("splits", [2, 16, 64])
def test_split_k_wgrad_shape(lib, splits):
    """wgrad: g_W (256, 512) = g_y^T z over a batch of 4096, both operands stored batch-major (MN-major)."""
    M, N, Kd = 256, 512, 4096
    gen = torch.Generator("cuda").manual_seed(splits)
    A2, B2, A, B = operands(M, N, Kd, 0, 0, gen, integers=True)
    C = run(lib, A, B, M, N, Kd, 0, 0, splits)
    assert torch.equal(C, (A2.double() @ B2.double().t()).float())
    C1 = run(lib, A, B, M, N, Kd, 0, 0, splits)
    assert torch.equal(C, C1)  # fixed-order reduction: deterministic


def test_argument_checks(lib):
    t = torch.zeros(64, 64, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    assert lib.pvae_gemm_tf32(t.data_ptr(), t.data_ptr(), t.data_ptr(), 64, 64, 6, 1, 1, 1, 1, None, s) == -2
    assert lib.pvae_gemm_tf32(None, t.data_ptr(), t.data_ptr(), 64, 64, 64, 1, 1, 1, 1, None, s) == -1
    assert lib.pvae_gemm_tf32(t.data_ptr(), t.data_ptr(), t.data_ptr(), 64, 64, 64, 1, 1, 1, 2, None, s) == -1  # no workspace


@pytest.mark.parametrize("shape", [(4096, 256, 512), (200, 256, 64), (12, 136, 100)])
def test_residual_epilogue_matches_separate_kernels(lib, shape):
    """Decoder GEMM with loss_recon fused into its epilogue == GEMM, then pvae_recon_residual."""
    M, N, Kd = shape
    gen = torch.Generator("cuda").manual_seed(3)
    z = torch.rand(M, Kd, device="cuda", generator=gen)
    w = torch.randn(N, Kd, device="cuda", generator=gen) * 0.05
    x = torch.randn(M, N, device="cuda", generator=gen)
    s = torch.cuda.current_stream().cuda_stream
    scale = 2.0 / M
    y = run(lib, z, w, M, N, Kd, 1, 1, precise=1)
    g_ref, r_ref = torch.empty_like(x), torch.empty(M, device="cuda")
    assert lib.pvae_recon_residual(y.data_ptr(), x.data_ptr(), g_ref.data_ptr(), r_ref.data_ptr(), M, N, scale, s) == 0
    P = lib.pvae_gemm_residual_parts(N)
    g, parts = torch.full_like(x, float("nan")), torch.full((M, P), float("nan"), device="cuda")
    rc = lib.pvae_gemm_tf32_residual(z.data_ptr(), w.data_ptr(), x.data_ptr(), g.data_ptr(), parts.data_ptr(), M, N, Kd,
                                     scale, 1, s)
    assert rc == 0, lib.pvae_last_error()
    torch.cuda.synchronize()
    assert torch.equal(g, g_ref)  # same accumulators, same subtraction
    assert torch.allclose(parts.sum(1), r_ref, rtol=1e-6, atol=0)
    # the loss kernel accepts the partial sums
    kl = torch.rand(M, device="cuda", generator=gen)
    out_a, out_b = torch.empty(3, device="cuda"), torch.empty(3, device="cuda")
    assert lib.pvae_loss_assemble(parts.data_ptr(), kl.data_ptr(), out_a.data_ptr(), M, 2.5, M, P, s) == 0
    assert lib.pvae_loss_assemble(r_ref.data_ptr(), kl.data_ptr(), out_b.data_ptr(), M, 2.5, M, 1, s) == 0
    assert torch.allclose(out_a, out_b, rtol=1e-6)


@pytest.mark.parametrize("majors", [(1, 1), (1, 0), (0, 0), (0, 1)])
@pytest.mark.parametrize("shape", [(4096, 512, 256), (4096, 256, 512), (256, 512, 4096), (520, 136, 100), (12, 64, 256)])
def test_operand_planes_are_bit_identical_to_the_in_kernel_split(lib, majors, shape):
    """pvae_gemm_tf32_planes (low parts supplied as their own planes, loaded by TMA) issues the same MMAs on the same
    bits as precise = 1 (low parts computed per stage in shared memory): every element equal, with and without split-K."""
    M, N, Kd = shape
    gen = torch.Generator("cuda").manual_seed(3 * M + N + Kd)
    _, _, A, B = operands(M, N, Kd, *majors, gen, integers=False)
    s = torch.cuda.current_stream().cuda_stream
    A_lo, B_lo = torch.empty_like(A), torch.empty_like(B)
    assert lib.pvae_tf32_lo(A.data_ptr(), A_lo.data_ptr(), A.numel(), s) == 0, lib.pvae_last_error()
    assert lib.pvae_tf32_lo(B.data_ptr(), B_lo.data_ptr(), B.numel(), s) == 0, lib.pvae_last_error()
    hi = (A.view(torch.int32) & -8192).view(torch.float32)  # trunc_tf32: the 13 low mantissa bits cleared
    assert torch.equal(A_lo, A - hi)
    for splits in (1, 4) if Kd >= 256 else (1,):
        want = run(lib, A, B, M, N, Kd, *majors, splits=splits, precise=1)
        C = torch.full((M, N), float("nan"), device="cuda")
        ws = torch.empty(max(1, lib.pvae_gemm_ws_floats(M, N, splits)), device="cuda")
        rc = lib.pvae_gemm_tf32_planes(A.data_ptr(), A_lo.data_ptr(), B.data_ptr(), B_lo.data_ptr(), C.data_ptr(), M, N, Kd, *majors,
                                       splits, ws.data_ptr(), s)
        assert rc == 0, lib.pvae_last_error()
        torch.cuda.synchronize()
        assert torch.equal(C, want), splits
