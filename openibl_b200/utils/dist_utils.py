"""One process per GPU over torch.distributed (mirror of ibl/utils/dist_utils.py:11-76)."""
import os
import subprocess

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def init_dist(launcher, args, backend="nccl"):
    if mp.get_start_method(allow_none=True) is None:
        mp.set_start_method("spawn")
    if launcher == "pytorch":
        init_dist_pytorch(args, backend)
    elif launcher == "slurm":
        init_dist_slurm(args, backend)
    else:
        raise ValueError("Invalid launcher type: {}".format(launcher))


def init_dist_pytorch(args, backend="nccl"):
    args.rank = int(os.environ["LOCAL_RANK"])
    args.ngpus_per_node = torch.cuda.device_count()
    args.gpu = args.rank
    args.world_size = args.ngpus_per_node
    torch.cuda.set_device(args.gpu)
    dist.init_process_group(backend=backend)


def init_dist_slurm(args, backend="nccl"):
    args.rank = int(os.environ["SLURM_PROCID"])
    args.world_size = int(os.environ["SLURM_NTASKS"])
    args.ngpus_per_node = torch.cuda.device_count()
    args.gpu = args.rank % args.ngpus_per_node
    torch.cuda.set_device(args.gpu)
    addr = subprocess.getoutput("scontrol show hostname {} | head -n1".format(os.environ["SLURM_NODELIST"]))
    os.environ.update(MASTER_PORT=str(args.tcp_port), MASTER_ADDR=addr, WORLD_SIZE=str(args.world_size),
                      RANK=str(args.rank))
    dist.init_process_group(backend=backend)
    args.total_gpus = dist.get_world_size()


def convert_sync_bn(model, process_group=None, gpu=None):
    """VGG16 has no BatchNorm (netvlad_img.py:103): kept for API compatibility."""
    converted = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model, process_group)
    return converted.cuda(gpu) if gpu is not None else converted


def synchronize():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
