"""N>1 host logic on CPU: world_size-2 gloo.  View sharding, the packed SUM / MAX all-reduce, and the
equivalence of "reduce ||means2D.grad|| per view locally, then all-reduce the accumulators" with the
reference's single-process per-view loop (geometry/gaussian_base.py:815-819, 846-851)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from b200splat import batched
from b200splat import dist as bdist


def test_shard_views_partitions():
    for B in (1, 4, 7, 32, 64):
        for W in (1, 2, 3, 4, 8):
            spans = [bdist.shard_views(B, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_packed_layout():
    pk = batched.PackedGrads(100, 16, "cpu")
    assert pk.buffer.numel() == 100 * (13 + 3 * 16)            # 4*P*(13+3M) bytes  
    assert pk.views["shs"].shape == (100, 16, 3) and pk.views["opacities"].shape == (100, 1)
    pk.views["means3D"].fill_(1.0)
    pk.views["denom"].fill_(2.0)
    assert float(pk.buffer.sum()) == 300 + 200
    pk2 = batched.PackedGrads(10, 0, "cpu", color_mode="colors_precomp")
    # P is padded to a multiple of 4 per field so that every field starts on a 16-byte boundary for any P
    assert pk2.P4 == 12 and pk2.buffer.numel() == 12 * (13 + 3) == batched.PackedGrads.floats(10, 0, "colors_precomp")
    assert all(off % 4 == 0 for _, off, _ in pk2.fields)
    assert pk2.views["rotations"].shape == (10, 4) and pk2.views["rotations"].is_contiguous()
    segs = pk2.segments(4)                      # Gaussians [4, P): the tail range covers the (zero) padding
    assert all(sg[0] % 4 == 0 and sg[1] % 4 == 0 for sg in segs) and segs[-1][:3] == (pk2.buffer.numel() + 4, 8, 1)
    assert all(sg[3] == 0 for sg in segs)       # no live map (no shared exchange storage): every range is dense
    # with exchange storage that has room for the live map, the float4-granular per-Gaussian fields become row-sparse
    n = batched.PackedGrads.floats(10, 16) + 12 + batched.PackedGrads.live_floats(10)
    pk3 = batched.PackedGrads(10, 16, "cpu", storage=torch.zeros(n))
    assert pk3.live_map is not None and pk3.live_map.numel() == 12 and pk3.live_map.dtype == torch.uint8
    assert pk3.live_offset_bytes == (batched.PackedGrads.floats(10, 16) + 12) * 4
    sp = {name: sg for (name, _, _), sg in zip(pk3.fields, pk3.segments(4))}
    assert sp["rotations"][3:] == (4, 4) and sp["shs"][3:] == (48, 4)
    assert all(sp[k][3] == 0 for k in ("means3D", "scales", "opacities", "grad_accum", "denom"))
    pk2.views["means3D"].fill_(1.0)
    assert float(pk2.buffer.sum()) == 30.0      # the padding Gaussians stay zero


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _view_data(P, v):
    g = torch.Generator().manual_seed(100 + v)
    radii = torch.randint(0, 30, (P,), generator=g, dtype=torch.int32)
    radii[torch.rand(P, generator=g) < 0.3] = 0
    m2 = torch.randn(P, 3, generator=g) * (radii > 0)[:, None]
    gx = torch.randn(P, 3, generator=g) * (radii > 0)[:, None]
    return radii, m2, gx


def _worker(rank, world, port, P, B, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pk = batched.PackedGrads(P, 1, "cpu")
    pk.zero_stats_()
    b, e = bdist.shard_views(B, rank, world)
    for i, v in enumerate(range(b, e)):
        radii, m2, gx = _view_data(P, v)
        vis = radii > 0
        # what the fused preprocess-backward epilogue does per view (csrc/preprocess_bwd.cu)
        if i == 0:
            pk.views["means3D"].copy_(gx)
        else:
            pk.views["means3D"].add_(gx)
        pk.views["grad_accum"][vis] += m2[vis, :2].norm(dim=-1)
        pk.views["denom"][vis] += 1
        torch.maximum(pk.max_radii, radii.float(), out=pk.max_radii)
    bdist.allreduce_packed(pk.buffer, pk.max_radii)
    if rank == 0:
        torch.save({"buffer": pk.buffer, "max_radii": pk.max_radii}, out)
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world2_allreduce_matches_reference_update_states(tmp_path):
    P, B, world = 257, 5, 2
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(world, _free_port(), P, B, out), nprocs=world, join=True)
    got = torch.load(out)
    # single-process restatement of the reference loop over the B views
    xyz_accum, denom, max_r, gsum = torch.zeros(P, 1), torch.zeros(P, 1), torch.zeros(P), torch.zeros(P, 3)
    for v in range(B):
        radii, m2, gx = _view_data(P, v)
        vis = radii > 0
        max_r = torch.max(max_r, radii.float())
        xyz_accum[vis] += torch.norm(m2[vis, :2], dim=-1, keepdim=True)
        denom[vis] += 1
        gsum += gx
    pk = batched.PackedGrads(P, 1, "cpu")
    pk.buffer.copy_(got["buffer"])
    assert torch.allclose(pk.views["means3D"], gsum, atol=1e-6)
    assert torch.allclose(pk.views["grad_accum"], xyz_accum[:, 0], atol=1e-6)
    assert torch.equal(pk.views["denom"], denom[:, 0])
    assert torch.equal(got["max_radii"], max_r)
    # and the helper that folds a step's reduced statistics into the persistent accumulators
    A, D, M = torch.zeros(P, 1), torch.zeros(P, 1), torch.zeros(P)
    bdist.reference_update_states(A, D, M, pk.views["grad_accum"], pk.views["denom"], got["max_radii"])
    assert torch.allclose(A, xyz_accum, atol=1e-6) and torch.equal(D, denom) and torch.equal(M, max_r)
