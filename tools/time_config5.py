"""Config-5 timing alone (run plain for 1 GPU or under torchrun): prints bench.config5_metrics."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import bench  # noqa: E402
import fes_b200 as F  # noqa: E402

world, rank, local = int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
out = bench.config5_metrics(F, world, rank, torch.device('cuda', local), n=int(sys.argv[1]) if len(sys.argv) > 1 else 151,
                            pcg_iters=int(sys.argv[2]) if len(sys.argv) > 2 else 300)
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
