"""All-reduce of the packed gradient buffer alone: own NVSwitch-multicast kernel vs own NVLink peer-memory kernel vs NCCL.
torchrun --nproc-per-node N scripts/p2p_bench.py [P] [M]   -> one line per method on rank 0"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "threestudio-3dgs_b200"))
import torch, torch.distributed as dist
from b200splat import dist as bdist, batched

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 16
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=dev)
n_sum, n_max = batched.PackedGrads.floats(P, M), batched.PackedGrads.padded(P)
ar = bdist.P2PAllReduce(n_sum, n_max, dev)
try:
    mc = bdist.MulticastAllReduce(n_sum, n_max, dev)
except Exception as exc:
    mc = None
    if rank == 0:
        print(json.dumps({"method": "nvswitch_multicast_kernel", "unavailable": repr(exc)}))
ref = torch.zeros(n_sum + n_max, device=dev)

def check(obj, name):
    """random per-rank data -> the kernel's result against an NCCL SUM / MAX of the same data; identical on all ranks"""
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    x = torch.randn(n_sum + n_max, generator=g).to(dev)
    x[n_sum:] = x[n_sum:].abs().round()          # radii: non-negative
    obj.buffer.copy_(x)
    want = x.clone()
    dist.all_reduce(want[:n_sum]); dist.all_reduce(want[n_sum:], op=dist.ReduceOp.MAX)
    obj(); torch.cuda.synchronize()
    err = float((obj.buffer[:n_sum] - want[:n_sum]).abs().max() / want[:n_sum].abs().max())
    max_ok = bool(torch.equal(obj.buffer[n_sum:], want[n_sum:]))
    chk = obj.buffer.view(torch.int32).to(torch.int64).sum().reshape(1)
    allchk = [torch.empty_like(chk) for _ in range(world)]
    dist.all_gather(allchk, chk)
    same = all(bool(torch.equal(c, allchk[0])) for c in allchk)
    if rank == 0:
        print(json.dumps({"check": name, "rel_err_vs_nccl": err, "max_exact": max_ok, "bit_identical_across_ranks": same}))
    assert err < 1e-5 and max_ok and same, name

check(ar, "p2p_nvlink_kernel")
if mc is not None:
    check(mc, "nvswitch_multicast_kernel")

def timed(fn, iters=20):
    for _ in range(3):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

def nccl():
    dist.all_reduce(ref[:n_sum]); dist.all_reduce(ref[n_sum:], op=dist.ReduceOp.MAX)

ms_p2p = timed(lambda: ar())
ms_nccl = timed(nccl)
ms_mc = timed(lambda: mc()) if mc is not None else None
if rank == 0:
    nbytes = (n_sum + n_max) * 4
    for name, ms in (("nvswitch_multicast_kernel", ms_mc), ("p2p_nvlink_kernel", ms_p2p), ("nccl", ms_nccl)):
        if ms is None:
            continue
        print(json.dumps({"method": name, "world": world, "bytes": nbytes, "ms": ms, "algbw_GBps": nbytes / ms / 1e6,
                          "per_gpu_each_way_GBps": nbytes * (world - 1) / world / ms / 1e6}))
assert not ar.failed()
if mc is not None:
    assert not mc.failed()
    mc.close()
ar.close()
dist.destroy_process_group()
