"""GPU parity tests for the GEMM backends through ase_gemm (C ABI).
  backend 0 (SIMT fp32)    vs torch fp64 matmul: <= 1e-5 relative to max|C| (fp32 accumulation order only)
  backend 1 (tcgen05 3xTF32) vs torch fp64 matmul: <= 2e-5 relative to max|C|  -- i.e. fp32-class accuracy,
  two orders of magnitude tighter than single-pass TF32 (~2e-3) so a broken hi/lo split cannot hide.
  backend 2 (tcgen05 3xFP16, per-tensor power-of-two scaled hi/lo planes): the same bar."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(A, B, a_trans, b_trans, bias, act, mask_src, mask_mode, alpha):
    a = A.double().t() if a_trans else A.double()
    b = B.double() if b_trans else B.double().t()
    c = alpha * (a @ b)
    if bias is not None:
        c = c + bias.double()
    if act == 1:
        c = torch.relu(c)
    elif act == 2:
        c = torch.tanh(c)
    if mask_mode == 1:
        c = c * (mask_src > 0).double()
    elif mask_mode == 2:
        c = c * (1 - mask_src.double() ** 2)
    return c


def _run(M, N, K, a_trans, b_trans, backend, bias=False, act=0, mask_mode=0, accumulate=False, split_k=0, alpha=1.0, seed=0,
         lda_pad=0, tol=1e-5):
    from ase_b200 import ops
    g = torch.Generator().manual_seed(seed + M + 7 * N + 13 * K)
    A = torch.randn((K, M + lda_pad) if a_trans else (M, K + lda_pad), generator=g).cuda()
    B = torch.randn((K, N + lda_pad) if b_trans else (N, K + lda_pad), generator=g).cuda()
    Av = A[:, :M] if a_trans else A[:, :K]
    Bv = B[:, :N] if b_trans else B[:, :K]
    bias_t = torch.randn(N, generator=g).cuda() if bias else None
    mask_src = torch.randn(M, N, generator=g).cuda().clamp(-0.9, 0.9) if mask_mode else None
    out = None
    base = 0
    if accumulate:
        out = torch.randn(M, N, generator=g).cuda()
        base = out.double().clone()
    C = ops.gemm(Av, Bv, a_trans, b_trans, bias_t, act, mask_src, mask_mode, out, accumulate, split_k, alpha, backend)
    torch.cuda.synchronize()
    ref = _ref(Av, Bv, a_trans, b_trans, bias_t, act, mask_src, mask_mode, alpha) + base
    err = float((C.double() - ref).abs().max() / ref.abs().max())
    assert err < tol, (M, N, K, a_trans, b_trans, backend, err)


@pytest.mark.parametrize('a_trans,b_trans', [(False, False), (False, True), (True, False), (True, True)])
def test_simt_layouts_ragged(a_trans, b_trans):
    _run(200, 150, 77, a_trans, b_trans, 0)
    _run(128, 128, 8, a_trans, b_trans, 0)
    _run(1, 1, 1, a_trans, b_trans, 0)
    _run(259, 31, 317, a_trans, b_trans, 0, lda_pad=3)      # unaligned leading dimensions -> scalar loads


def test_simt_epilogues():
    _run(300, 200, 64, False, False, 0, bias=True, act=1)
    _run(300, 64, 256, False, False, 0, bias=True, act=2, alpha=1.0 / 16, tol=2e-5)
    _run(300, 200, 64, False, True, 0, mask_mode=1)
    _run(300, 64, 100, False, True, 0, mask_mode=2)
    _run(96, 200, 4096, True, True, 0, accumulate=True, split_k=7, tol=2e-5)
    _run(300, 1, 512, False, False, 0, bias=True)           # value / logit heads (N = 1)
    _run(4096, 512, 1, False, True, 0, mask_mode=1)         # K = 1 outer product (dV . w_value)
    _run(64, 64, 64, False, False, 0, alpha=0.37)


def test_simt_learner_shapes():
    _run(512, 1024, 1400, False, False, 0, bias=True, act=1)
    _run(512, 1400, 1024, False, True, 0)
    _run(1024, 317, 2048, True, True, 0, accumulate=True, split_k=4, tol=2e-5)


TC_BACKENDS = [1, 2]      # 1: 3xTF32 planes, 2: 3xFP16 scaled planes -- same kernels, same bar


@pytest.mark.parametrize('tc', TC_BACKENDS)
@pytest.mark.parametrize('a_trans,b_trans', [(False, False), (False, True), (True, False), (True, True)])
def test_tc_matches_fp64(a_trans, b_trans, tc):
    # a_trans / b_trans operands are consumed MN-major by the tensor core (no transposition pass)
    _run(256, 256, 64, a_trans, b_trans, tc, tol=1e-5)
    _run(384, 192, 317, a_trans, b_trans, tc, lda_pad=3, tol=1e-5)     # ragged K (zero padded to 320), N tail with BN=128
    _run(1000, 100, 96, a_trans, b_trans, tc, tol=1e-5)                # ragged M and N
    _run(200, 40, 50, a_trans, b_trans, tc, tol=1e-5)                  # BN = 64 path
    _run(31, 512, 700, a_trans, b_trans, tc, tol=1e-5)                 # M smaller than one tile (mu-head weight gradient)
    _run(300, 1, 512, a_trans, b_trans, tc, tol=1e-5)                  # N = 1 (value / logit heads)
    _run(512, 96, 1, a_trans, b_trans, tc, tol=1e-5)                   # K = 1 outer product
    _run(1, 1, 1, a_trans, b_trans, tc, tol=1e-5)


@pytest.mark.parametrize('tc', TC_BACKENDS)
@pytest.mark.parametrize('a_trans,b_trans', [(False, False), (False, True), (True, False), (True, True)])
def test_tc_wide_tiles_128x256(a_trans, b_trans, tc):
    """N >= 384 runs on the 128x256-tile kernel (8 drain warps, single-buffered main accumulator)."""
    _run(256, 512, 96, a_trans, b_trans, tc, tol=1e-5)
    _run(300, 1400, 317, a_trans, b_trans, tc, lda_pad=3, tol=1e-5)      # ragged everything, 6 N tiles with a tail
    _run(1000, 384, 1024, a_trans, b_trans, tc, bias=True, act=1, tol=1e-5)
    _run(512, 1024, 2048, a_trans, b_trans, tc, mask_mode=1, tol=1e-5)    # long K: many main/corr hand-offs


@pytest.mark.parametrize('tc', TC_BACKENDS)
def test_tc_wide_tiles_split_k_and_colsum(tc):
    from ase_b200 import ops
    _run(1024, 1024, 8192, True, True, tc, accumulate=True, split_k=3, tol=2e-5)
    g = torch.Generator().manual_seed(5)
    A = torch.randn(700, 256, generator=g).cuda(); B = torch.randn(512, 256, generator=g).cuda()
    cs = torch.zeros(512, device='cuda')
    C = ops.gemm(A, B, backend=tc, colsum_out=cs)
    ref = A.double() @ B.double().t()
    assert float((C.double() - ref).abs().max() / ref.abs().max()) < 1e-5
    assert float((cs.double() - ref.sum(0)).abs().max() / ref.sum(0).abs().max()) < 1e-5


@pytest.mark.parametrize('tc', TC_BACKENDS)
def test_tc_epilogues_and_split_k(tc):
    _run(512, 256, 512, False, False, tc, bias=True, act=1, tol=1e-5)
    _run(256, 128, 256, False, False, tc, bias=True, act=2, alpha=1.0 / 16, tol=3e-5)   # O(1) pre-activations
    _run(512, 320, 256, False, True, tc, mask_mode=1, tol=2e-5)
    _run(256, 64, 128, False, True, tc, mask_mode=2, tol=2e-5)
    _run(1024, 512, 4096, True, True, tc, accumulate=True, split_k=5, tol=3e-5)
    _run(256, 256, 2048, False, False, tc, alpha=0.25, tol=2e-5)      # > STAGES k-blocks: ring wrap-around + phase flips


@pytest.mark.parametrize('tc', TC_BACKENDS)
@pytest.mark.parametrize('M,N,K', [(300, 512, 96), (256, 96, 64), (200, 40, 50), (1000, 1024, 128)])
def test_tc_relu_activity_bits_roundtrip(tc, M, N, K):
    """relu_bits_out packs (C > 0) into 1 bit per element (word n / 32, bit n % 32); a masked GEMM given those bits instead of the
    fp32 activation must produce exactly what it produces from the fp32 mask (the backward pass reads 1 bit instead of 32)."""
    from ase_b200 import ops
    g = torch.Generator().manual_seed(M + N)
    A = torch.randn(M, K, generator=g).cuda(); W = torch.randn(N, K, generator=g).cuda(); bias = torch.randn(N, generator=g).cuda()
    ldb = (N + 31) // 32
    bits = torch.zeros(M, ldb, dtype=torch.int32, device='cuda')
    Y = ops.gemm(A, W, bias=bias, act=1, backend=tc, relu_bits_out=bits)
    torch.cuda.synchronize()
    cols = torch.arange(N, device='cuda')
    unpacked = ((bits[:, cols // 32] >> (cols % 32)) & 1).bool()
    assert torch.equal(unpacked, Y > 0)
    if N % 32:
        assert int((bits[:, -1].long() & 0xFFFFFFFF >> (N % 32) << (N % 32)).abs().max()) == 0      # padding bits stay clear
    dZ = torch.randn(M, 70, generator=torch.Generator().manual_seed(1)).cuda(); W2 = torch.randn(70, N, generator=g).cuda()
    a = ops.gemm(dZ, W2, b_trans=True, mask_src=Y, mask_mode=1, backend=tc)
    b = ops.gemm(dZ, W2, b_trans=True, mask_src=Y, mask_mode=1, backend=tc, mask_bits=bits)
    assert torch.equal(a, b)


@pytest.mark.parametrize('a_trans,b_trans', [(False, False), (False, True), (True, False), (True, True)])
def test_tc_many_tiles_per_cta(a_trans, b_trans):
    """More output tiles than resident CTAs / CTA pairs: the persistent pair kernel (M % 128 == 0 shapes) walks several work items per
    pair, alternating its two TMEM buffers across item boundaries; the ragged shapes take the one-tile-per-CTA kernels in several
    waves.  Odd and even k-block counts per tile exercise both parities of every barrier across tile boundaries."""
    from ase_b200 import ops
    _many_tile_cases(a_trans, b_trans, ops)


def _many_tile_cases(a_trans, b_trans, ops):
    _run(8192, 1024, 192, a_trans, b_trans, 2, bias=True, act=1, tol=1e-5)       # 512 tiles, 3 k-blocks each
    _run(4096, 1024, 64, a_trans, b_trans, 2, tol=1e-5)                           # 256 tiles, 1 k-block each
    _run(5000, 900, 256, a_trans, b_trans, 2, mask_mode=1, tol=1e-5)              # 320 tiles, ragged M and N tails, 4 k-blocks
    _run(3000, 1400, 130, a_trans, b_trans, 2, lda_pad=3, tol=1e-5)               # 264 tiles, ragged everything, unaligned ld
    g = torch.Generator().manual_seed(9)
    A = torch.randn(6016, 320, generator=g).cuda(); B = torch.randn(768, 320, generator=g).cuda()
    if a_trans or b_trans:
        return
    cs = torch.zeros(768, device='cuda')
    C = ops.gemm(A, B, backend=2, colsum_out=cs)                                    # 282 tiles + fused column sums
    ref = A.double() @ B.double().t()
    assert float((C.double() - ref).abs().max() / ref.abs().max()) < 1e-5
    assert float((cs.double() - ref.sum(0)).abs().max() / ref.sum(0).abs().max()) < 1e-5


def test_tc_fp16_planes_dynamic_range():
    """Backend 2 scales every tensor by a power of two before the FP16 hi/lo split: gradient-sized (1e-9) and large (1e+6)
    operands, and a tensor whose entries span 7 decades, must come out as accurately as O(1) ones."""
    from ase_b200 import ops
    g = torch.Generator().manual_seed(3)
    A = torch.randn(300, 512, generator=g); B = torch.randn(200, 512, generator=g)
    for sa, sb in ((1e-9, 1.0), (1e6, 1e-3), (3e-20, 7e12), (1e-30, 1e-5)):
        a = (A * sa).cuda(); b = (B * sb).cuda()
        C = ops.gemm(a, b, backend=2)
        ref = a.double() @ b.double().t()
        assert float((C.double() - ref).abs().max() / ref.abs().max()) < 1e-5, (sa, sb)
    # rows of very different magnitude inside one tensor: each output row is judged against ITS OWN scale down to 1e-5 of the
    # tensor max (an exactly scaled split keeps 22 significant bits for every element within 2^-26 of the max)
    rs = torch.logspace(0, -5, 300).unsqueeze(1)
    a = (A * rs).cuda(); b = B.cuda()
    C = ops.gemm(a, b, backend=2)
    ref = a.double() @ b.double().t()
    err = (C.double() - ref).abs().amax(1) / ref.abs().amax(1)
    assert float(err.max()) < 2e-5, float(err.max())
    z = ops.gemm(torch.zeros(64, 64, device='cuda'), b[:, :64].contiguous(), backend=2)       # all-zero operand: scale 1, exact zeros
    assert float(z.abs().max()) == 0.0


@pytest.mark.parametrize('tc', TC_BACKENDS)
def test_tc_accuracy_is_fp32_class_not_tf32(tc):
    """Inputs chosen so that single-pass TF32 would miss by ~1e-3: all products need the low halves."""
    from ase_b200 import ops
    g = torch.Generator().manual_seed(1)
    A = (1.0 + torch.rand(256, 1024, generator=g) * 1e-3).cuda()      # values whose information sits below TF32's 10 mantissa bits
    B = (1.0 + torch.rand(128, 1024, generator=g) * 1e-3).cuda()
    C = ops.gemm(A, B, backend=tc)
    ref = A.double() @ B.double().t()
    # single-pass TF32 rounds every operand to 10 mantissa bits: |err| ~ 0.25 on these sums of ~1025.  3xTF32 with the
    # tensor core accumulating across all of K left a one-sided 0.006 (truncating accumulator); with the k-block partials
    # accumulated outside in fp32 RN the error is a few ulps of 1025 (1.2e-4 each).
    assert float((C.double() - ref).abs().max()) < 1.5e-3


# ------------------------------------------------------------------------------------------------ persistent CTA-pair kernel (gemm_tc2.cu)
def _pair(on):
    import os
    os.environ['ASE_TC_PAIR'] = '1' if on else '0'


@pytest.mark.parametrize('a_trans,b_trans', [(False, False), (False, True), (True, False), (True, True)])
def test_tc_pair_kernel_many_items_per_pair(a_trans, b_trans):
    """Backend 2, N >= 384, M % 128 == 0 runs on the persistent cta_group::2 kernel (256 x 256 tiles, 74 CTA pairs).  More work items than
    pairs: every pair walks several tiles, so the ring / main / correction barriers cross tile boundaries with both parities (odd and
    even k-block counts), the last pair-tile of an odd 128-row tile count has an idle peer, N tails are zero-filled by TMA."""
    _pair(True)
    _run(8192, 1024, 192, a_trans, b_trans, 2, bias=True, act=1, tol=1e-5)        # 128 items, 3 k-blocks each
    _run(16384, 512, 64, a_trans, b_trans, 2, tol=1e-5)                            # 128 items, 1 k-block each
    _run(4992, 1400, 256, a_trans, b_trans, 2, bias=True, tol=1e-5)                # 39 x 128 rows (odd), 6 column tiles with a 120-column tail
    _run(2944, 904, 130, a_trans, b_trans, 2, lda_pad=3, tol=1e-5)                 # ragged K (zero padded), N % 8 == 0 tail, unaligned operand ld
    _run(128, 384, 640, a_trans, b_trans, 2, bias=True, act=1, tol=1e-5)                # a single item: peer CTA has no rows
    _run(32768, 1024, 1024, a_trans, b_trans, 2, tol=1e-5)                         # the benchmarked forward / dX shape


def test_tc_pair_kernel_agrees_with_one_tile_kernels():
    """Both kernel families form the same products; they differ only in where the tensor core's truncating adds happen (the pair kernel
    drains main + correction terms of every k-block together, the 128x256 kernel keeps the correction terms in TMEM across K): the
    results must agree to a few fp32 ulps of the output scale -- far inside the distance of either from fp64."""
    from ase_b200 import ops
    g = torch.Generator().manual_seed(11)
    for (M, N, K, bt) in ((1024, 1024, 1024, False), (768, 1400, 320, True), (2048, 512, 4096, False)):
        A = torch.randn(M, K, generator=g).cuda()
        B = (torch.randn(K, N, generator=g) if bt else torch.randn(N, K, generator=g)).cuda()
        bias = torch.randn(N, generator=g).cuda()
        _pair(True); c1 = ops.gemm(A, B, b_trans=bt, bias=bias, act=1, backend=2)
        _pair(False); c0 = ops.gemm(A, B, b_trans=bt, bias=bias, act=1, backend=2)
        _pair(True)
        torch.cuda.synchronize()
        ref = torch.relu(A.double() @ (B.double() if bt else B.double().t()) + bias.double())
        scale = float(ref.abs().max())
        assert float((c0 - c1).abs().max()) <= 2e-6 * scale, (M, N, K, bt, float((c0 - c1).abs().max()) / scale)
        e1, e0 = float((c1.double() - ref).abs().max()) / scale, float((c0.double() - ref).abs().max()) / scale
        assert e1 <= max(2.0 * e0, 2e-6), (M, N, K, bt, e1, e0)


def test_tc_pair_kernel_split_k_colsum_bits():
    from ase_b200 import ops
    _pair(True)
    _run(1024, 1024, 32768, True, True, 2, accumulate=True, split_k=9, tol=2e-5)     # the dW shape: 16 pair tiles x 9 splits on 74 pairs
    _run(512, 1400, 12288, True, True, 2, accumulate=True, split_k=4, tol=2e-5)
    _run(1024, 512, 4096, True, True, 2, accumulate=True, split_k=1, tol=2e-5)       # accumulate without split
    g = torch.Generator().manual_seed(5)
    A = torch.randn(6016, 320, generator=g).cuda(); B = torch.randn(768, 320, generator=g).cuda()
    cs = torch.zeros(768, device='cuda')
    C = ops.gemm(A, B, backend=2, colsum_out=cs)                                      # 47 x 128 rows, fused column sums (shuffle butterfly)
    ref = A.double() @ B.double().t()
    assert float((C.double() - ref).abs().max() / ref.abs().max()) < 1e-5
    assert float((cs.double() - ref.sum(0)).abs().max() / ref.sum(0).abs().max()) < 1e-5
    # activity bits out (thread = row layout: one 16-byte store of 4 words per row) and bits in
    for (M, N, K) in ((1024, 1024, 128), (640, 1400, 96), (256, 416, 64)):
        A = torch.randn(M, K, generator=g).cuda(); W = torch.randn(N, K, generator=g).cuda(); bias = torch.randn(N, generator=g).cuda()
        bits = torch.zeros(M, (N + 31) // 32, dtype=torch.int32, device='cuda')
        Y = ops.gemm(A, W, bias=bias, act=1, backend=2, relu_bits_out=bits)
        cols = torch.arange(N, device='cuda')
        unpacked = ((bits[:, cols // 32] >> (cols % 32)) & 1).bool()
        assert torch.equal(unpacked, Y > 0), (M, N, K)
        if N % 32:
            assert int((bits[:, -1].long() & 0xFFFFFFFF >> (N % 32) << (N % 32)).abs().max()) == 0
        dZ = torch.randn(M, 512, generator=g).cuda(); W2 = torch.randn(512, N, generator=g).cuda()
        cs2 = torch.zeros(N, device='cuda')
        a = ops.gemm(dZ, W2, b_trans=True, mask_src=Y, mask_mode=1, backend=2, mask_bits=bits, colsum_out=cs2)      # pair kernel (bit mask)
        ref = (dZ.double() @ W2.double()) * (Y > 0).double()
        assert float((a.double() - ref).abs().max() / ref.abs().max()) < 1e-5, (M, N, K)
        assert float((cs2.double() - ref.sum(0)).abs().max() / ref.sum(0).abs().max()) < 2e-5, (M, N, K)
