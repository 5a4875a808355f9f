"""``gtb_rows_atb_f32`` -- out += act(A[index])^T B, colsum += sum(B) -- against float64: the tensor-core kernel
(csrc/atb_tc.cu: 64 x 64 blocks over >= 4096 rows) and the CUDA-core kernel (csrc/grad.cu) it falls back to.
Tolerance: 2e-6 of the largest entry of the exact result's absolute-value sum (3xTF32 products, fp32 atomics)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _check(n, ka, nb, gather, relu, with_colsum, seed=0):
    from gnn_tracking_b200 import ops
    gen = torch.Generator().manual_seed(seed + n)
    rows_a = n + 17 if gather else n
    a = torch.randn(rows_a, ka, generator=gen)
    b = torch.randn(n, nb, generator=gen)
    idx = torch.randint(0, rows_a, (n,), generator=gen).to(torch.int32) if gather else None
    out0 = torch.randn(ka, nb, generator=gen)
    cs0 = torch.randn(nb, generator=gen)
    out, cs = out0.clone().cuda(), cs0.clone().cuda()
    ops.rows_atb(a.cuda(), b.cuda(), out, a_index=idx.cuda() if gather else None, a_relu=relu, colsum=cs if with_colsum else None)
    torch.cuda.synchronize()
    d = torch.float64
    ag = a[idx.long()] if gather else a
    ag = torch.relu(ag) if relu else ag
    ref = out0.to(d) + ag.to(d).T @ b.to(d)
    scale = float((ag.abs().to(d).T @ b.abs().to(d)).max()) + 1.0
    err = float((out.cpu().to(d) - ref).abs().max())
    assert err <= 2e-6 * scale, (n, ka, nb, gather, relu, err, scale)
    if with_colsum:
        refc = cs0.to(d) + b.to(d).sum(0)
        errc = float((cs.cpu().to(d) - refc).abs().max())
        assert errc <= 2e-6 * (float(b.abs().to(d).sum(0).max()) + 1.0), errc
    else:
        assert torch.equal(cs.cpu(), cs0)


@pytest.mark.parametrize("n", [4096, 5000, 70001, 1000003])
@pytest.mark.parametrize("gather,relu,with_colsum", [(False, False, True), (True, True, False), (False, True, True), (True, False, True)])
def test_rows_atb_tensor_core(n, gather, relu, with_colsum):
    _check(n, 64, 64, gather, relu, with_colsum)


@pytest.mark.parametrize("n,ka,nb", [(100, 64, 64), (3000, 64, 64), (5000, 14, 64), (5000, 64, 1), (5000, 4, 64), (5000, 40, 24),
                                     (70001, 64, 1), (70001, 4, 64), (70001, 3, 40), (70001, 40, 2), (500, 4, 64), (70001, 1, 1)])
def test_rows_atb_cuda_core_shapes(n, ka, nb):
    _check(n, ka, nb, gather=True, relu=True, with_colsum=True)


def test_rows_atb_paths_agree(monkeypatch):
    """The two kernels on the same operands (the environment switch is read once per process: compare against float64 instead
    and check that strided operands -- column blocks of wider tables -- are taken)."""
    from gnn_tracking_b200 import ops
    gen = torch.Generator().manual_seed(3)
    n = 20000
    wide_a, wide_b = torch.randn(n, 192, generator=gen).cuda(), torch.randn(n, 128, generator=gen).cuda()
    out = torch.zeros(64, 64, device="cuda")
    ops.rows_atb(wide_a[:, 64:128], wide_b[:, 64:], out)
    ref = wide_a[:, 64:128].double().T @ wide_b[:, 64:].double()
    assert float((out.double() - ref).abs().max()) <= 2e-6 * float((wide_a[:, 64:128].abs().double().T @ wide_b[:, 64:].abs().double()).max())
