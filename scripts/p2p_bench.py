"""All-reduce of the packed gradient buffer alone: own NVLink peer-memory kernel vs NCCL.
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
n_sum, n_max = batched.PackedGrads.floats(P, M), P
ar = bdist.P2PAllReduce(n_sum, n_max, dev)
ref = torch.zeros(n_sum + n_max, device=dev)

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
if rank == 0:
    nbytes = (n_sum + n_max) * 4
    for name, ms in (("p2p_nvlink_kernel", ms_p2p), ("nccl", ms_nccl)):
        print(json.dumps({"method": name, "world": world, "bytes": nbytes, "ms": ms, "algbw_GBps": nbytes / ms / 1e6,
                          "per_gpu_each_way_GBps": nbytes * (world - 1) / world / ms / 1e6}))
assert not ar.failed()
ar.close()
dist.destroy_process_group()
