"""Worker of tests/test_training_gpu.py::test_ddp_gradient_allreduce_two_gpus (one process per GPU, torchrun):
the stage-1 model on the drop-ins under DistributedDataParallel — bucketed NCCL all-reduce of the gradients the libsblk
backward produces block by block (BASELINE configs[3]: "NCCL grad allreduce")."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.nn.functional as F

from sbl_for_multilingual_lip_reading_b200 import ops, stage1, synth

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
ops.init()
n, t = 4, 31
torch.manual_seed(7)
model = stage1.Stage1Classifier(n_layers_enc=3, dropout=0.0).to(dev).train()
model.load_synthetic(1, 3)
x = synth.structured_clips(n, t, seed=500 + rank).to(dev)
y = torch.randint(0, 1500, (n,), generator=torch.Generator().manual_seed(rank)).to(dev)
lang = torch.randint(0, 2, (n,), generator=torch.Generator().manual_seed(10 + rank)).to(dev)


def loss_of(m):
    torch.manual_seed(99)                       # same always-on dropout mask in both passes
    v_t, v_l = m(x)
    return F.cross_entropy(v_t, y) + 0.1 * F.cross_entropy(v_l, lang)


# local gradients without DDP
model.zero_grad()
loss_of(model).backward()
local = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
bn_before = {k: b.detach().clone() for k, b in model.named_buffers() if k.endswith("running_mean")}
# the same step under DDP
ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index], broadcast_buffers=False)
model.zero_grad()
loss_of(ddp).backward()
ok = True
worst = 0.0
for k, p in model.named_parameters():
    want = local[k].clone()
    dist.all_reduce(want)
    want /= world
    err = ((p.grad - want).norm() / (want.norm() + 1e-20)).item()
    # batch-stat BN + bf16: the two passes are the same function of the same inputs -> (near) bit-identical local grads
    worst = max(worst, err)
    ok = ok and err < 1e-3
    g0 = p.grad.clone()
    dist.broadcast(g0, 0)
    ok = ok and torch.equal(g0, p.grad)       # every rank holds the same reduced gradient
print(f"rank {rank}: ddp gradient all-reduce {'OK' if ok else 'MISMATCH'} (worst rel err vs mean of local grads {worst:.2e})",
      flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
