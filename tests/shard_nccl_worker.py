"""torchrun worker of tests/test_gpu_shard.py::test_two_gpu_nccl: every rank holds one bucket-range shard of the small
golden index on its own GPU, the exchanges are NCCL all_to_all_single, and each rank's results must equal the unsharded
path's on the same reads."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    import krepp_b200
    import krepp_b200.dist as kd
    from test_gpu_parity import fastq_reads
    from test_gpu_shard import SMALL, assert_same_results
    rank, world, local = kd.env_rank_world()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    idx = os.path.join(SMALL, "index")
    _, reads = fastq_reads(os.path.join(SMALL, "reads.fq"))
    mine = reads[slice(*kd.shard_range(len(reads), rank, world))]
    me = kd.ShardRank(idx, local, rank, world, len(mine), sum(len(s) for s in mine) + 64)
    job = kd.ShardedJob([me])
    from krepp_b200.capi import pack_reads
    bases, offs = pack_reads(mine)
    pad = np.zeros(len(bases) + 64, np.uint8)
    pad[:len(bases)] = bases
    for _ in range(2):  # twice: buffers are reused
        res = job.run([(torch.from_numpy(pad).cuda(), torch.from_numpy(offs.astype(np.int64)).cuda(), len(mine))])[0]
    res = {k: (np.array(v, copy=True) if isinstance(v, np.ndarray) else v) for k, v in res.items()}
    me.close()
    os.environ["KREPP_PIPELINE"] = "sorted"
    ix = krepp_b200.Index(idx, local)
    b = krepp_b200.IBatch(ix, mine)
    b.submit()
    assert_same_results(res, b.wait(), f"nccl rank {rank}")
    assert job.bytes_exchanged > 0
    dist.barrier()
    print(f"rank {rank} ok: {len(mine)} reads, {len(res['records'])} records, {job.bytes_exchanged} bytes received", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
