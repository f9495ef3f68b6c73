"""Diagnostic (multi-GPU): does the chunked all-reduce deliver the sum of the ranks' gradients?  Variants: overlap on/off,
NCCL-registered bucket on/off.  torchrun --nproc-per-node N tools/dp_diag.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bts_fully_tf_b200 import parallel, trainer  # noqa: E402
from bts_fully_tf_b200.decoder import BtsDecoder  # noqa: E402


def main():
    rank, local_rank, world = parallel.init_distributed("nccl")
    dev = torch.device("cuda", local_rank)
    torch.backends.cudnn.allow_tf32 = False
    chans, F, H, W, B = [64, 8, 8, 16, 24], 256, 64, 96, 2 * world
    torch.manual_seed(0)
    feats_all = [torch.relu(torch.randn(B, H // s, W // s, c, device=dev)) for s, c in zip((32, 2, 4, 8, 16), chans)]
    gt_all = torch.rand(B, H, W, 1, device=dev) * 10.0
    lo, hi = parallel.shard_range(B, world, rank)
    feats, gt = [f[lo:hi].contiguous() for f in feats_all], gt_all[lo:hi].contiguous()
    out = []
    for overlap in (True, False):
        for reg in (True, False):
            torch.manual_seed(0)
            dec = BtsDecoder(chans, 10.0, num_filters=F).to(dev)
            eng = trainer.DataParallelStep(dec, feats, gt, dataset="nyu", base_lr=1e-3, use_graph=False, overlap=overlap, register_nccl=reg)
            local, _ = eng.local_gradients()
            local2, _ = eng.local_gradients()
            gathered = [torch.empty_like(local) for _ in range(world)]
            dist.all_gather(gathered, local)
            ref = torch.stack(gathered).double().sum(0)
            summed, _ = eng.reduced_gradients()
            diff = (summed.double() - ref).abs()
            per_chunk = [float(diff[a:b].max() / ref.abs().max()) for a, b, _, _ in eng.flat.chunks]
            # is the result simply the local gradient (no exchange) ?
            same_as_local = float((summed.double() - local.double()).abs().max() / ref.abs().max())
            plain = local.clone()
            dist.all_reduce(plain)
            plain_err = float((plain.double() - ref).abs().max() / ref.abs().max())
            out.append({"overlap": overlap, "registered": bool(eng.registered), "rerun": float((local2 - local).abs().max() / local.abs().max()),
                        "per_chunk_err": per_chunk, "vs_local_only": same_as_local, "plain_allreduce_err": plain_err,
                        "order": list(eng.comm.launch_order), "hook_streams": sorted(eng.comm.hook_streams), "engine_stream": eng.stream.cuda_stream,
                        "default_stream": torch.cuda.default_stream(dev).cuda_stream})
            eng.close()
            del eng
    if rank == 0:
        print("DIAG " + json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
