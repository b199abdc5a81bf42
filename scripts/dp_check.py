"""Data-parallel check, run under torchrun on R GPUs of one box:
   - every rank trains K steps on its shard with one NCCL all-reduce of the [P+3] buffer per step;
   - parameters must be BIT-IDENTICAL on all ranks after K steps;
   - rank 0 compares loss curve and parameters with the float64 oracle run on the global point set.
Prints 'DP_CHECK OK ...' on success (exit code 0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from __graft_entry__ import load

P = load()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
which = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4099
K = int(sys.argv[3]) if len(sys.argv) > 3 else 10
nb = 257
prec = int(os.environ.get("PINN_DP_PRECISION", "0"))      # pinn_precision (4 = FP16X3 on the width-20 mma kernel)
cfg = P.default_config(which, n_interior=n, n_boundary=nb, n_initial=0 if which == 3 else nb, rank=rank, nranks=world, device=local,
                       precision=prec)
net = P.Network(cfg)
box = [P.Network.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
net.comm_init(box[0])
p2p = os.environ.get("PINN_P2P", "0") == "1"
if p2p:      # fused reduce + all-reduce + Adam over NVLink peer memory instead of NCCL
    hs = [None] * world
    dist.all_gather_object(hs, net.p2p_export())
    net.p2p_open(hs)
w0 = net.get_params()
losses = [net.train_step()["loss"] for _ in range(K)]
w = torch.from_numpy(net.get_params()).cuda()
allw = [torch.empty_like(w) for _ in range(world)]
dist.all_gather(allw, w)
ok = all(torch.equal(allw[0].view(torch.int32), a.view(torch.int32)) for a in allw)
# also exercise the caller-owned collective path: step_local + torch all_reduce + step_apply
ptr, nfl = net.grad_buffer()
net.step_local(); net.synchronize()
class _Buf:   # wrap the library's device buffer for torch without copying
    __cuda_array_interface__ = {"shape": (nfl,), "typestr": "<f4", "data": (ptr, False), "version": 2}
g = torch.as_tensor(_Buf(), device=f"cuda:{local}")
dist.all_reduce(g)
torch.cuda.synchronize()
o = net.step_apply()
if rank == 0:
    from oracle import oracle as O
    ocfg = O.baseline_config(which, n_interior=n, n_boundary=nb, n_initial=0 if which == 3 else nb)
    p_ref, _, _, l_ref, _ = O.train(ocfg, w0.astype(np.float64), K + 1, dtype=np.float64)
    dl = float(np.abs(np.array(losses + [o["loss"]]) / l_ref[:, 0] - 1).max())
    dp = float(np.abs(net.get_params() - p_ref).max())
    good = ok and dl < 1e-5 and dp < 1e-5
    print(f"DP_CHECK {'OK' if good else 'FAIL'} ranks={world} collective={'p2p-fused' if p2p else 'nccl'} precision={prec} identical_across_ranks={ok} loss_rel={dl:.2e} params_abs={dp:.2e}", flush=True)
    if not good:
        sys.exit(1)
net.close()
dist.destroy_process_group()
