"""The C-ABI entry points of include/gtb200.h called directly through ctypes with raw device
pointers (what a non-Python host would do): one Interaction-Network layer through
gtb_plan_build / gtb_mlp_pack / gtb_in_edge_forward_f32 / gtb_in_node_forward_f32 against the CPU
oracle of reference models/interaction_network.py:54-103.  Tolerance 1e-5 (fp32)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("impl_name", ["ffma", "tcgen05"])
@pytest.mark.parametrize("dims", [(64, 64, 64), (8, 4, 40)])
def test_in_layer_through_the_c_abi(impl_name, dims):
    from gnn_tracking_b200 import _lib
    from oracle import in_oracle as O
    L = _lib.lib()
    impl = {"ffma": _lib.IMPL_FFMA, "tcgen05": _lib.IMPL_TCGEN05}[impl_name]
    dn, de, h = dims
    gen = torch.Generator().manual_seed(7)
    n, e = 2500, 30000
    ei = torch.randint(0, n, (2, e), generator=gen)
    ei[1, :900] = 5
    x = torch.randn(n, dn, generator=gen)
    ea = torch.randn(e, de, generator=gen)
    sd = {}
    for pre, k0, out in (("relational_model.", 2 * dn + de, de), ("object_model.", dn + de, dn)):
        for li, (a, b) in enumerate(((k0, h), (h, h), (h, out))):
            sd[f"{pre}layers.{2 * li}.weight"] = torch.randn(b, a, generator=gen) / a ** 0.5
            sd[f"{pre}layers.{2 * li}.bias"] = torch.randn(b, generator=gen) * 0.1
    xt_ref, et_ref = O.interaction_network(x, ei, ea, sd, "")

    dev = torch.device("cuda")
    st = torch.cuda.current_stream().cuda_stream
    d = {k: v.to(dev).contiguous() for k, v in sd.items()}
    xd, eid, ead = x.to(dev), ei.to(dev).contiguous(), ea.to(dev)
    i32 = dict(dtype=torch.int32, device=dev)
    perm, src, dst = torch.empty(e, **i32), torch.empty(e, **i32), torch.empty(e, **i32)
    rowptr, status = torch.empty(n + 1, **i32), torch.empty(1, **i32)
    ws = torch.empty(L.gtb_plan_workspace_bytes(n, e), dtype=torch.uint8, device=dev)
    _lib.check(L.gtb_plan_build(eid.data_ptr(), n, e, perm.data_ptr(), rowptr.data_ptr(), src.data_ptr(), dst.data_ptr(),
                                status.data_ptr(), ws.data_ptr(), ws.numel(), st))

    def pack(pre, k0, out, blocks):
        dims_c = (C.c_int32 * 4)(k0, h, h, out)
        bw = (C.c_int32 * len(blocks))(*blocks)
        nb = L.gtb_mlp_packed_bytes(3, dims_c, len(blocks), bw, impl)
        assert nb > 0
        buf = torch.empty(nb, dtype=torch.uint8, device=dev)
        wp = (C.c_void_p * 3)(*[d[f"{pre}layers.{2 * i}.weight"].data_ptr() for i in range(3)])
        bp = (C.c_void_p * 3)(*[d[f"{pre}layers.{2 * i}.bias"].data_ptr() for i in range(3)])
        _lib.check(L.gtb_mlp_pack(3, dims_c, len(blocks), bw, wp, bp, impl, buf.data_ptr(), st))
        return buf

    p_rel = pack("relational_model.", 2 * dn + de, de, [dn, dn, de])
    p_obj = pack("object_model.", dn + de, dn, [dn, de])
    e_t = torch.empty(e, de, device=dev)
    aggr = torch.empty(n, de, device=dev)
    x_t = torch.empty(n, dn, device=dev)
    _lib.check(L.gtb_in_edge_forward_f32(xd.data_ptr(), dn, 0, ead.data_ptr(), de, 0, n, e, perm.data_ptr(), rowptr.data_ptr(),
                                         src.data_ptr(), dst.data_ptr(), dn, de, h, de, p_rel.data_ptr(), impl,
                                         e_t.data_ptr(), de, aggr.data_ptr(), st))
    _lib.check(L.gtb_in_node_forward_f32(xd.data_ptr(), dn, 0, aggr.data_ptr(), n, dn, de, h, dn, p_obj.data_ptr(), impl,
                                         0.0, 1.0, None, 0, x_t.data_ptr(), dn, st))
    torch.cuda.synchronize()
    for got, ref, what in ((e_t, et_ref, "e_tilde"), (x_t, xt_ref, "x_tilde")):
        scale = max(1.0, float(ref.abs().max()))
        err = float((got.cpu() - ref).abs().max())
        assert err <= 1e-5 * scale, (what, err, scale)
