"""The whole-model Wanda 2:4 run through load_pruner().prune() under cProfile, 1 GPU or data-parallel under torchrun:
where the host time of the drop-in driver goes.  python scripts/full_model_dp_probe.py  |  torchrun ... scripts/full_model_dp_probe.py"""
import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vlm-compression_b200"))
import torch
import torch.distributed as dist
import bench
from vlmc import native
native.load()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
method = sys.argv[1] if len(sys.argv) > 1 else "wanda_nm"
for rep in range(2):
    pr = cProfile.Profile()
    if rep == 1 and rank == 0:
        pr.enable()
    r = bench.full_model_vicuna(torch, native, dev, method, rank, world)
    if rep == 1 and rank == 0:
        pr.disable()
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
        print(s.getvalue()[:9000])
    if rank == 0:
        print(f"rep {rep} world {world} {method}: {r['value']:.3f} s/model ok={r['structure_ok']}", flush=True)
if world > 1:
    dist.destroy_process_group()
