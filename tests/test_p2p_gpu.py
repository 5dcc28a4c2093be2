"""All-reduce over NVLink peer memory (csrc/p2p.cu) against torch sums; needs >= 2 GPUs on the box
(`gpurun --gpus 2 -- python -m pytest tests/test_p2p_gpu.py -m gpu`); skipped on a single GPU."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_sum, n_max, rounds, out, kind="p2p"):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "threestudio-3dgs_b200"))
    import torch.distributed as dist
    from b200splat import dist as bdist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    if kind == "mc":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        try:
            ar = bdist.MulticastAllReduce(n_sum, n_max, dev)
        except RuntimeError as exc:       # no NVSwitch multicast on this box: all ranks raise together
            if rank == 0:
                out.put(-1)
            dist.destroy_process_group()
            return
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
        ar = bdist.P2PAllReduce(n_sum, n_max, dev)
    ok = True
    for r in range(rounds):
        gens = [torch.Generator().manual_seed(1000 * r + k) for k in range(world)]
        parts = [torch.randn(n_sum + n_max, generator=g) for g in gens]
        for prt in parts:
            prt[n_sum:] = prt[n_sum:].abs().round()      # the MAX segment holds radii: non-negative
        want_sum = parts[0][:n_sum].clone()
        for k in range(1, world):            # the kernel sums in rank order: bit-exact expectation
            want_sum += parts[k][:n_sum]
        want_max = torch.stack([p[n_sum:] for p in parts]).max(0).values
        ar.buffer.copy_(parts[rank].to(dev))
        ar()
        torch.cuda.synchronize(dev)
        got = ar.buffer.cpu()
        if kind == "mc":     # the switch's summation order is its own: tolerance on the sum, identical bits on all ranks
            scale = float(want_sum.abs().max().clamp_min(1e-30))
            ok = ok and float((got[:n_sum] - want_sum).abs().max()) <= 1e-6 * scale and torch.equal(got[n_sum:], want_max)
            chk = ar.buffer.view(torch.int32).to(torch.int64).sum().reshape(1)
            allchk = [torch.empty_like(chk) for _ in range(world)]
            dist.all_gather(allchk, chk)
            ok = ok and all(bool(torch.equal(c, allchk[0])) for c in allchk)
        else:
            ok = ok and torch.equal(got[:n_sum], want_sum) and torch.equal(got[n_sum:], want_max)
    ok = ok and not ar.failed()
    ar.close()
    res = torch.tensor([1 if ok else 0], device=dev if kind == "mc" else "cpu")
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(int(res.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_sum,n_max", [(61 * 4096, 4096), (1 << 20, 0), (4, 4)])
def test_p2p_allreduce_matches_rank_ordered_sum(n_sum, n_max):
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    mp.spawn(_worker, args=(world, _free_port(), n_sum, n_max, 3, q), nprocs=world, join=True)
    assert q.get() == 1


@pytest.mark.parametrize("n_sum,n_max", [(61 * 4096, 4096), (1 << 20, 0), (8, 8)])
def test_multicast_allreduce_matches_sum_and_max(n_sum, n_max):
    """b200splat_mc_allreduce (multimem.ld_reduce / multimem.st through NVSwitch multicast)."""
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    mp.spawn(_worker, args=(world, _free_port(), n_sum, n_max, 3, q, "mc"), nprocs=world, join=True)
    r = q.get()
    if r == -1:
        pytest.skip("no NVSwitch multicast support on this box")
    assert r == 1


def _sparse_worker(rank, world, port, rows, out, kind):
    """Row-sparse exchange: a packed-gradient-like layout [dense 3 floats/row | rot 4/row | sh 48/row | max 1/row |
    live map 1 byte/row]; each rank has gradients for a random ~15 % of the rows (zeros elsewhere) and flags them in
    its live map; result == the dense rank-ordered sum, rows nobody flagged stay untouched (zero)."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "threestudio-3dgs_b200"))
    import torch.distributed as dist
    from b200splat import dist as bdist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    R4 = (rows + 3) // 4 * 4
    widths = [3, 4, 48]
    n_sum, n_max, n_tail = sum(widths) * R4, R4, R4 // 4
    if kind == "mc":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        try:
            ar = bdist.MulticastAllReduce(n_sum, n_max, dev, n_tail=n_tail)
        except RuntimeError:
            if rank == 0:
                out.put(-1)
            dist.destroy_process_group()
            return
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
        ar = bdist.P2PAllReduce(n_sum, n_max, dev, n_tail=n_tail)
    ar.live_offset = (n_sum + n_max) * 4
    ok = True
    for r, g0 in enumerate((0, 0, 256)):         # whole range twice, then a sub-range starting at row 256
        lives, parts = [], []
        for k in range(world):
            g = torch.Generator().manual_seed(77 * r + k)
            live = torch.rand(R4, generator=g) < 0.15
            live[rows:] = False
            buf = torch.zeros(n_sum + n_max)
            off = 0
            for w in widths:
                fld = torch.randn(R4, w, generator=g) * live[:, None]
                buf[off:off + w * R4] = fld.reshape(-1)
                off += w * R4
            buf[n_sum:] = (torch.rand(R4, generator=g) * 9).round()
            lives.append(live)
            parts.append(buf)
        want = parts[0].clone()
        for k in range(1, world):
            want[:n_sum] += parts[k][:n_sum]
        want[n_sum:] = torch.stack([p[n_sum:] for p in parts]).max(0).values
        segs, off = [], 0
        for w in widths:                          # (offset, count, op, row floats, first row)
            segs.append((off + w * g0, w * (R4 - g0), 0, w if w % 4 == 0 else 0, g0))
            off += w * R4
        segs.append((n_sum + g0, R4 - g0, 1, 0, 0))
        if g0:                                    # rows below g0 are not exchanged: they keep the local values
            off = 0
            for w in widths:
                want[off:off + w * g0] = parts[rank][off:off + w * g0]
                off += w * R4
            want[n_sum:n_sum + g0] = parts[rank][n_sum:n_sum + g0]
        ar.buffer[:n_sum + n_max].copy_(parts[rank].to(dev))
        ar.buffer[n_sum + n_max:].view(torch.uint8).copy_(lives[rank].to(torch.uint8).to(dev))
        torch.cuda.synchronize(dev)
        dist.barrier()
        ar(segs)
        torch.cuda.synchronize(dev)
        got = ar.buffer[:n_sum + n_max].cpu()
        if kind == "mc":
            scale = float(want[:n_sum].abs().max().clamp_min(1e-30))
            ok = ok and float((got[:n_sum] - want[:n_sum]).abs().max()) <= 1e-6 * scale
            ok = ok and torch.equal(got[n_sum:], want[n_sum:])
        else:
            ok = ok and torch.equal(got, want)
    ok = ok and not ar.failed()
    ar.close()
    res = torch.tensor([1 if ok else 0], device=dev if kind == "mc" else "cpu")
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(int(res.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["p2p", "mc"])
@pytest.mark.parametrize("rows", [100_003, 5000, 700])
def test_row_sparse_exchange_equals_the_dense_sum(rows, kind):
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    mp.spawn(_sparse_worker, args=(world, _free_port(), rows, q, kind), nprocs=world, join=True)
    r = q.get()
    if r == -1:
        pytest.skip("no NVSwitch multicast support on this box")
    assert r == 1
