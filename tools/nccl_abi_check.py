#!/usr/bin/env python
"""dgadj_allreduce_indicators on a raw NCCL communicator, one rank per GPU (run under torchrun):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tools/nccl_abi_check.py
Every rank marches its shard of one global batch, reduces its indicator partials on the device,
combines them through the C-ABI on an ncclComm_t created here with ncclCommInitRank (the id is
passed around with torch.distributed), and checks the result against (a) the torch.distributed
path (`sharding.allreduce_indicators`, ordered) bit for bit and (b) the unsharded batch on one GPU."""
import ctypes as C
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import dgadj_loader

pkg = dgadj_loader.load_package()
from adjoint_ode_adaptivity_b200.sharding import allreduce_indicators, allreduce_indicators_comm, shard_range

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))

nccl = C.CDLL("libnccl.so.2")      # the instance torch loaded (same soname)


class UniqueId(C.Structure):
    _fields_ = [("internal", C.c_byte * 128)]


uid = UniqueId()
if rank == 0:
    assert nccl.ncclGetUniqueId(C.byref(uid)) == 0
t = torch.tensor(list(bytes(uid)), dtype=torch.uint8, device="cuda")
dist.broadcast(t, 0)
C.memmove(C.byref(uid), bytes(t.cpu().numpy().tobytes()), 128)
comm = C.c_void_p()
nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
assert nccl.ncclCommInitRank(C.byref(comm), world, uid, rank) == 0

N, K, B, S = 4, 64, 1000, 30
s = pkg.AdvecDG1D(N, K, domain=(0.0, 2 * math.pi), alpha=0.0, bc="periodic", device=local)
rng = np.random.default_rng(7)
x = s.g.x
u0_all = np.sin(x[None] * rng.integers(1, 4, (B, 1, 1)) + rng.uniform(0, 6.28, (B, 1, 1)))
dt, _ = s.cfl_dt(2 * math.pi)
lo, hi = shard_range(B, rank, world)
out = s.fwd_adj(torch.tensor(u0_all[lo:hi], device="cuda"), 2 * math.pi, dt, S, want_uT=False)
part = s.reduce_indicators(out["eta"], out["J"])
ref = allreduce_indicators(part.clone(), ordered=True)
got = allreduce_indicators_comm(s, comm.value, part.clone())
torch.cuda.synchronize()
assert torch.equal(got, ref), (got - ref).abs().max()
full = s.fwd_adj(torch.tensor(u0_all, device="cuda"), 2 * math.pi, dt, S, want_uT=False)
whole = s.reduce_indicators(full["eta"], full["J"])
relerr = float(((got - whole).abs() / whole.abs().clamp_min(1e-300)).max())
assert relerr < 1e-12, relerr
assert int(got[:K].argmax()) == int(whole[:K].argmax())
gathered = [None] * world
dist.all_gather_object(gathered, got.cpu().numpy().tobytes())
assert all(g == gathered[0] for g in gathered)          # the same bits on every rank
# count-independent form: block partials (125 trajectories per block: whole blocks per rank for 2, 4 or 8 ranks)
# all-gathered and combined in global block order through the C-ABI on the raw communicator -- the SAME BITS as
# the unsharded batch reduced on one GPU, and as the torch.distributed path
R = 125
parts = s.reduce_indicator_blocks(out["eta"], out["J"], rows_per_block=R)
tot = torch.empty(K + 4, dtype=torch.float64, device="cuda")
rc = s.lib.dgadj_allreduce_indicator_blocks(s._h, C.c_void_p(comm.value), K, parts.shape[0], C.c_void_p(parts.data_ptr()),
                                            C.c_void_p(tot.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
assert rc == 0, s.lib.dgadj_last_error(s._h)
torch.cuda.synchronize()
one = pkg.combine_blocks(s.reduce_indicator_blocks(full["eta"], full["J"], rows_per_block=R))
assert torch.equal(tot, one), (tot - one).abs().max()
assert torch.equal(pkg.allreduce_indicator_blocks(parts), one)
import hashlib
digest = hashlib.sha256(tot.cpu().numpy().tobytes()).hexdigest()
nccl.ncclCommDestroy(comm)
if rank == 0:
    print("nccl C-ABI all-reduce ok: world %d, K+4 = %d, vs torch.distributed bit-identical, vs unsharded rel %.1e, refine element %d"
          % (world, K + 4, relerr, int(got[:K].argmax())))
    print("nccl C-ABI blocked all-reduce ok: world %d, %d blocks per rank, BIT-IDENTICAL to the unsharded batch; sha256 %s"
          % (world, parts.shape[0], digest))
dist.destroy_process_group()
