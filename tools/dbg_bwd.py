import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, numpy as np
from oracle import durf_oracle as O
import test_gpu_kernels as T
from durf_b200 import ops, _lib
for topo, M in [((60, 256, 8, 4, 27, 128), 3), ((60, 256, 8, 4, 27, 128), 311), ((63, 128, 8, 4, 27, 128), 150)]:
    N = 128
    layers, x, cond = T._mlp_inputs(topo, M, N, 41)
    ot = O.MLPTopology(*topo)
    # oracle with bf16-rounded weights/inputs to separate quantisation from bugs
    params = [(torch.from_numpy(k).requires_grad_(True), torch.from_numpy(b).requires_grad_(True)) for k, b in layers]
    import durf_test_helpers as H
    rgb, den = (H.mlp_apply_bf16_emulated if os.environ.get('EMU','1')=='1' else O.mlp_apply)(params, ot, x, cond)
    g = torch.Generator().manual_seed(3)
    d_rgb = torch.randn(M, N, 3, generator=g) * 0.1
    d_den = torch.randn(M, N, generator=g) * 0.1
    (rgb * d_rgb).sum().add((den[..., 0] * d_den).sum()).backward()
    blob = T._blob(topo, layers)
    packed = ops.mlp_pack(topo, blob)
    tiles = T._tile_images(x, M, topo[0]).cuda()
    _, _, saved = ops.mlp_fwd(topo, tiles, cond.cuda(), blob, M=M, N=N, precision=_lib.PREC_BF16, packed=packed, save=True)
    d_blob = torch.zeros_like(blob)
    ops.mlp_bwd(topo, tiles, cond.cuda(), blob, saved, d_rgb.cuda(), d_den.cuda(), d_blob, M=M, N=N, precision=_lib.PREC_BF16, packed=packed)
    torch.cuda.synchronize()
    print("topo", topo, "M", M)
    for i, ((dw, db), (pk, pb)) in enumerate(zip(ops.mlp_layer_views(topo, d_blob), params)):
        for got, want, what in ((dw.cpu(), pk.grad, f"dW{i}"), (db.cpu(), pb.grad, f"db{i}")):
            cos = float((got * want).sum() / (got.norm() * want.norm() + 1e-30))
            rel = float((got - want).norm() / (want.norm() + 1e-30))
            print(f"  {what:5s} cos {cos:.5f} rel {rel:.3e} |want| {float(want.norm()):.3e}")
