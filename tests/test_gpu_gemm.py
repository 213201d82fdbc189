"""The contraction engines through the C ABI (aa_gemm): fp32 SIMT (exact), tcgen05 bf16 and
tcgen05 tf32, all four operand layouts, ragged shapes -- against torch float64 matmul on the
same (already rounded) inputs."""
import ctypes

import pytest
import torch

from adaptive_b200 import _lib
from adaptive_b200.functional import _ptr, _stream

pytestmark = pytest.mark.gpu

SHAPES = [(128, 128, 64), (80, 2048, 512), (1440, 49, 512), (1360, 10000, 512), (49, 512, 3920), (257, 96, 40), (4096, 520, 776),
          (2304, 4352, 1024)]      # (the last one is large enough for the 256-column bf16 tiles)


def _run(engine, M, N, K, a_k, b_k, with_c, with_bias, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    dt = torch.bfloat16 if engine == 1 else torch.float32
    A = torch.randn((M, K) if a_k else (K, M), generator=g, device="cuda").to(dt)
    B = torch.randn((N, K) if b_k else (K, N), generator=g, device="cuda").to(dt)
    C = torch.randn(M, N, generator=g, device="cuda") if with_c else None
    bias = torch.randn(N, generator=g, device="cuda") if with_bias else None
    D = torch.full((M, N), float("nan"), device="cuda")
    lib = _lib.load()
    rc = lib.aa_gemm(engine, M, N, K, _ptr(A), A.stride(0), 1 if a_k else 0, _ptr(B), B.stride(0), 1 if b_k else 0, _ptr(C), N, 0.5,
                     _ptr(bias), _ptr(D), N, _stream(D.device))
    _lib.check(rc, "aa_gemm")
    torch.cuda.synchronize()
    Ad = A.double() if a_k else A.double().t()
    Bd = B.double() if b_k else B.double().t()
    ref = Ad @ Bd.t()
    if with_c:
        ref = ref + 0.5 * C.double()
    if with_bias:
        ref = ref + bias.double()
    return D, ref


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("a_k,b_k", [(1, 1), (1, 0), (0, 0), (0, 1)])
def test_simt_fp32(M, N, K, a_k, b_k):
    D, ref = _run(0, M, N, K, a_k, b_k, True, True)
    assert torch.isfinite(D).all()
    assert float((D.double() - ref).abs().max() / ref.abs().max()) < 1e-5     # fp32 accumulation over K up to 3920


def _aligned(M, N, K, a_k, b_k, es):
    # TMA needs 16-byte row strides: the contiguous extent of each operand must be a multiple of 16/es
    q = 16 // es
    return ((K if a_k else M) % q == 0) and ((K if b_k else N) % q == 0)


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("a_k,b_k", [(1, 1), (1, 0), (0, 0), (0, 1)])
def test_tcgen05_bf16(M, N, K, a_k, b_k):
    if not _aligned(M, N, K, a_k, b_k, 2):
        pytest.skip("operand row stride not 16-byte aligned")
    D, ref = _run(1, M, N, K, a_k, b_k, True, True)
    assert torch.isfinite(D).all()
    # inputs are exact bf16; products exact in fp32; only the fp32 accumulation order differs
    assert float((D.double() - ref).abs().max() / ref.abs().max()) < 1e-5


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_tcgen05_tf32(M, N, K):
    a_k = b_k = 1      # the tf32 engine serves the forward (K-major) form only
    D, ref = _run(2, M, N, K, a_k, b_k, False, True)
    assert torch.isfinite(D).all()
    assert float((D.double() - ref).abs().max() / ref.abs().max()) < 2e-3      # tf32: 10-bit mantissa inputs


def test_tcgen05_tf32_rejects_mn_major():
    lib = _lib.load()
    A = torch.randn(64, 64, device="cuda")
    D = torch.empty(64, 64, device="cuda")
    rc = lib.aa_gemm(2, 64, 64, 64, _ptr(A), 64, 0, _ptr(A), 64, 1, None, 0, 0.0, None, _ptr(D), 64, _stream(D.device))
    assert rc == 3 and b"K-major" in lib.aa_last_error()


def test_tcgen05_rejects_misaligned():
    lib = _lib.load()
    A = torch.randn(64, 66, device="cuda").bfloat16()[:, :65]
    B = torch.randn(64, 65, device="cuda").bfloat16()
    D = torch.empty(64, 64, device="cuda")
    rc = lib.aa_gemm(1, 64, 64, 65, _ptr(A), 65, 1, _ptr(B), 65, 1, None, 0, 0.0, None, _ptr(D), 64, _stream(D.device))
    assert rc != 0 and b"16 bytes" in lib.aa_last_error()


@pytest.mark.parametrize("M,N,K", SHAPES + [(4096, 10000, 512), (300, 49, 100), (64, 2560, 776)])
def test_tcgen05_split3_is_fp32_accurate(M, N, K):
    """3xTF32 over (hi, lo)-split operands: fp32-level accuracy (the engine of the decode contractions)."""
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(M, K, generator=g, device="cuda")
    B = torch.randn(N, K, generator=g, device="cuda")
    bias = torch.randn(N, generator=g, device="cuda")
    Kp = (K + 31) // 32 * 32
    As = torch.full((M, 2 * Kp), float("nan"), device="cuda")
    Bs = torch.full((N, 2 * Kp), float("nan"), device="cuda")
    D = torch.full((M, N), float("nan"), device="cuda")
    lib = _lib.load()
    st = _stream(D.device)
    _lib.check(lib.aa_split_tf32(_ptr(A), K, M, K, _ptr(As), Kp, st), "aa_split_tf32")
    _lib.check(lib.aa_split_tf32(_ptr(B), K, N, K, _ptr(Bs), Kp, st), "aa_split_tf32")
    _lib.check(lib.aa_gemm_split3(M, N, Kp, _ptr(As), _ptr(Bs), _ptr(bias), _ptr(D), N, st), "aa_gemm_split3")
    torch.cuda.synchronize()
    # the split itself: hi has 13 zero low mantissa bits, hi + lo reproduces x to ~2^-22, pads are zero
    hi, lo = As[:, :K], As[:, Kp:Kp + K]
    assert int((hi.view(torch.int32) & 0x1FFF).abs().max()) == 0 and int((lo.view(torch.int32) & 0x1FFF).abs().max()) == 0
    assert float(((hi.double() + lo.double()) - A.double()).abs().max()) <= 2.0 ** -21 * float(A.abs().max())
    if Kp > K:
        assert float(As[:, K:Kp].abs().max()) == 0 and float(As[:, Kp + K:].abs().max()) == 0
    ref = A.double() @ B.double().t() + bias.double()
    err = float((D.double() - ref).abs().max() / ref.abs().max())
    simt = torch.empty(M, N, device="cuda")
    _lib.check(lib.aa_gemm(0, M, N, K, _ptr(A), K, 1, _ptr(B), K, 1, None, N, 0.0, _ptr(bias), _ptr(simt), N, st), "aa_gemm")
    err_simt = float((simt.double() - ref).abs().max() / ref.abs().max())
    assert torch.isfinite(D).all()
    # single-pass tf32 is ~1e-3, fp32 SIMT ~1e-6.  3xTF32 measures 3.6e-6 at K=512 and 2.9e-5 at K=3920 on B200: the tensor
    # core adds into its fp32 accumulator with truncation, so the error grows linearly with the number of accumulation
    # steps (3 per 8 elements of K) instead of with its square root.  The decode contractions have K <= 1536.
    assert err < 1e-5 * max(1.0, K / 1024), (err, err_simt)


@pytest.mark.parametrize("M,N,K", [(4096, 2048, 512), (300, 2560, 256), (64, 10000, 512), (517, 1024, 800)])
def test_cta_pair_split3_equals_single_cta(M, N, K):
    """The CTA-pair kernel (tcgen05.mma.cta_group::2, 256 x 256 tiles over two SMs) against the single-CTA kernel on the same
    pre-split operands: the same fp32-accurate result (each output element is the same sequence of k-steps into one fp32
    accumulator; the decode tests' shard == slice property must not depend on which kernel a batch size selects -- the selection
    does not look at M, and this pins that even a different kernel would not move a logit by more than the last place)."""
    g = torch.Generator(device="cuda").manual_seed(11)
    A = torch.randn(M, K, generator=g, device="cuda")
    B = torch.randn(N, K, generator=g, device="cuda")
    bias = torch.randn(N, generator=g, device="cuda")
    Kp = (K + 31) // 32 * 32
    As, Bs = torch.zeros(M, 2 * Kp, device="cuda"), torch.zeros(N, 2 * Kp, device="cuda")
    lib = _lib.load()
    st = _stream(A.device)
    _lib.check(lib.aa_split_tf32(_ptr(A), K, M, K, _ptr(As), Kp, st), "aa_split_tf32")
    _lib.check(lib.aa_split_tf32(_ptr(B), K, N, K, _ptr(Bs), Kp, st), "aa_split_tf32")
    outs = []
    try:
        for pair in (0, 1):
            lib.aa_debug_set_gemm_pair(pair)
            D = torch.full((M, N), float("nan"), device="cuda")
            _lib.check(lib.aa_gemm_split3(M, N, Kp, _ptr(As), _ptr(Bs), _ptr(bias), _ptr(D), N, st), "aa_gemm_split3")
            torch.cuda.synchronize()
            outs.append(D)
    finally:
        lib.aa_debug_set_gemm_pair(-1)
    assert torch.isfinite(outs[1]).all()
    ref = A.double() @ B.double().t() + bias.double()
    scale = float(ref.abs().max())
    assert float((outs[1].double() - ref).abs().max()) / scale < 1e-5
    # bit for bit wherever the single-CTA kernel keeps one K range (measured on B200: the first three shapes); it cuts K = 800 of the
    # last shape into two ranges summed with red.add, which moves the last places
    assert float((outs[1] - outs[0]).abs().max()) / scale < (2e-7 if K <= 512 else 1e-5)
    print("pair == single bit for bit:", bool(torch.equal(outs[0], outs[1])))
