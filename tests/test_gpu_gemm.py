"""tcgen05/TMA GEMM mainloop (vast_gemm_nt) vs exact / torch fp32 references.  Validates the TMA
tensor maps, UMMA smem + instruction descriptors, TMEM accumulator plumbing and split-K path."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _grid(shape, gen, dtype):
    # values k/8, |k| <= 4: every product and partial sum is exact in fp32 -> bit-exact comparison
    return (torch.randint(-4, 5, shape, generator=gen, device="cuda").float() / 8).to(dtype)


@pytest.mark.parametrize("m,n,k", [(128, 256, 64), (128, 256, 256), (256, 512, 1024), (200, 300, 136),
                                   (1, 8, 8), (129, 257, 72), (512, 1024, 4096), (64, 4096, 512)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_gemm_exact_grid(m, n, k, dtype):
    from vast_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(m * 7 + n * 3 + k)
    a = _grid((m, k), g, dtype)
    b = _grid((n, k), g, dtype)
    c = ops.gemm_nt(a, b)
    ref = a.double() @ b.double().T
    torch.cuda.synchronize()
    assert torch.equal(c.double(), ref), f"max abs err {(c.double() - ref).abs().max().item()}"


@pytest.mark.parametrize("m,n,k", [(128, 256, 64), (128, 256, 128), (256, 512, 1024), (200, 304, 136), (1, 8, 8),
                                   (129, 264, 72), (512, 1024, 4096), (100, 72, 304)])
@pytest.mark.parametrize("da,db", [(torch.bfloat16, torch.bfloat16), (torch.float16, torch.float16)])
def test_gemm_nn_mn_major_b_exact_grid(m, n, k, da, db):
    """B row-major [K, N] consumed as the MN-major UMMA operand (no transpose): the dQ GEMM of the contrastive
    step multiplies the fp16 probabilities by the row-major gathered features.  (Mixed fp16 x bf16 operands
    are rejected: tcgen05.mma kind::f16 raises an illegal-instruction fault for them on sm_100a.)"""
    from vast_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(m * 5 + n * 11 + k)
    a = _grid((m, k), g, da)
    b = _grid((k, n), g, db)
    c = ops.gemm_nn(a, b)
    ref = a.double() @ b.double()
    torch.cuda.synchronize()
    assert torch.equal(c.double(), ref), f"max abs err {(c.double() - ref).abs().max().item()}"


def test_gemm_nn_strided_view_of_packed_buffer():
    from vast_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    pack = torch.randn(700, 2 * 264, generator=g, device="cuda").half()
    p = torch.rand(300, 704, generator=g, device="cuda").half()[:, :700]   # ld 704 >= 700
    for half in (0, 1):
        b = pack[:, half * 264:(half + 1) * 264]
        c = ops.gemm_nn(p, b)
        ref = p.double() @ b.double()
        assert (c.double() - ref).abs().max().item() < 2e-4 * ref.abs().max().item()


def test_gemm_random_and_alpha_and_strides():
    from vast_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    big_a = torch.randn(300, 2 * 264, generator=g, device="cuda").bfloat16()
    big_b = torch.randn(700, 2 * 264, generator=g, device="cuda").bfloat16()
    a, b = big_a[:, 264:], big_b[:, :264]  # strided views, like the packed all-gather buffer
    c = ops.gemm_nt(a, b, alpha=0.5)
    ref = 0.5 * (a.double() @ b.double().T)
    assert (c.double() - ref).abs().max().item() < 2e-4 * ref.abs().max().item()


def test_gemm_large_splitk_matches_fp64():
    from vast_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn(4096, 4096, generator=g, device="cuda").bfloat16()
    b = torch.randn(1024, 4096, generator=g, device="cuda").bfloat16()
    c = ops.gemm_nt(a, b)
    ref = a.double() @ b.double().T
    rel = ((c.double() - ref).norm() / ref.norm()).item()
    assert rel < 1e-5, rel


def test_gemm_bad_args_raise():
    from vast_b200 import ops
    a = torch.zeros(16, 12, device="cuda", dtype=torch.bfloat16)  # ld 12 not a multiple of 8
    b = torch.zeros(16, 12, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        ops.gemm_nt(a, b)
