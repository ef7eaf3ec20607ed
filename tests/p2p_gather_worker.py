"""Worker of tests/test_parity_gpu.py::test_p2p_gather_two_gpus (one process per GPU, launched with torchrun)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from sbl_for_multilingual_lip_reading_b200 import ops, sharding

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
ops.init()
n = 32 * 29 * 512
g = sharding.P2PGather(n, dev)
ok = True
for step in range(6):
    local = torch.full((n,), float(100 * step + rank), device=dev) + torch.arange(n, device=dev).float() * 1e-3
    full = g(local)
    torch.cuda.synchronize()
    want = torch.cat([torch.full((n,), float(100 * step + r), device=dev) + torch.arange(n, device=dev).float() * 1e-3
                      for r in range(world)])
    ok = ok and torch.equal(full, want)
    ref = torch.empty_like(want)
    dist.all_gather_into_tensor(ref, local)
    ok = ok and torch.equal(full, ref)
# timing: one-shot peer-memory gather vs NCCL
def timed(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
local = torch.randn(n, device=dev)
ref = torch.empty(world * n, device=dev)
t_p2p = timed(lambda: g(local))
t_nccl = timed(lambda: dist.all_gather_into_tensor(ref, local))
g.close()
print(f"rank {rank}: p2p gather {'OK' if ok else 'MISMATCH'}; {t_p2p:.1f} us vs NCCL all_gather {t_nccl:.1f} us "
      f"({n * 4 / 1e6:.1f} MB per rank, world {world})", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
