"""Slab exchange / kernel timing probe under torchrun: per-rank device times of the ghost exchange and the fine-level kernel.
usage: torchrun ... scripts/halo_probe.py MX [options]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import exsaddle_b200 as X
from oracle import oracle as O
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dist.init_process_group("nccl", device_id=torch.device("cuda", local))
mx = int(sys.argv[1]); lv = 6 if mx == 64 else 7 if mx == 128 else 5
abf = " ".join(l for l in O.ABF_OPTS.split("\n") if l.strip())
g = X.ExSaddle(abf + " -saddle_fieldsplit_u_pc_mg_levels %d -xsb_matrix_free full -model 6 -mx %d -eta1 1e6 -saddle_ksp_rtol 1e-8 %s" % (lv, mx, " ".join(sys.argv[2:])), nsd=3, device=local)
uid = [X.comm_unique_id() if rank == 0 else None]; dist.broadcast_object_list(uid, src=0); g.comm_init(uid[0], rank, world)
g.assemble().ksp_setup()
res = g.time_halo(100); res["ilu_ms"] = g.time_pc_schur(20); res["rank"] = rank; res["p2p"] = g.comm_info()["p2p"]
allr = [None] * world; dist.all_gather_object(allr, res)
if rank == 0:
    for r in allr: print("halo_probe", mx, json.dumps(r))
g.close(); dist.barrier(); dist.destroy_process_group()
