"""N>1 host logic on CPU (gloo, world_size 2): the corpus sharding bench.py uses (rank r owns bytes
[r*unit, (r+1)*unit) of the seeded G2 stream), the max-over-ranks timing reduction, and the fact that
per-rank frames concatenate to the single-process result (output does not depend on the GPU count)."""
import hashlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers

ROOT = helpers.ROOT
UNIT = 3 << 20      # 3 MiB per rank: whole frames at the default frame size


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import __graft_entry__ as ge
    pkg = ge.load_package()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shard = pkg.corpus.g2(UNIT, offset=rank * UNIT, threads=2)
    comp = helpers.oracle_compress(shard.tobytes())            # stands in for the rank's GPU (bytes are identical by test_gpu_*)
    lz = helpers.oracle_lzma2_compress(shard.tobytes())[1]     # method 21: the rank's chunk stream (ends with its own 0x00)
    zp = helpers.oracle_compress(shard.tobytes(), flags=1 | 0x20)                    # the price-based parses shard the same way
    lp = helpers.oracle_lzma2_compress(shard.tobytes(), flags=1 | (2 << 8) | 0x10)[1]
    t = torch.tensor([0.25 + rank, float(len(comp))], dtype=torch.float64)
    tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    q.put((rank, hashlib.sha256(shard.tobytes()).hexdigest(), comp, float(tmax[0]), float(tsum[1]), lz, zp, lp))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding(pkg):
    world, port = 2, 29000 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole = pkg.corpus.g2(world * UNIT).tobytes()
    for r, sha, comp, tmax, tsum, _lz, _zp, _lp in res:
        assert sha == hashlib.sha256(whole[r * UNIT:(r + 1) * UNIT]).hexdigest()      # shards tile the stream
        assert tmax == 0.25 + (world - 1)                                              # max over ranks
        assert tsum == sum(len(x[2]) for x in res)
    joined = b"".join(x[2] for x in res)
    assert joined == helpers.oracle_compress(whole)                                    # invariant to the rank count
    assert helpers.oracle_decompress(joined, len(whole)) == whole
    # method 21: shards are runs of dictionary-reset blocks; dropping every rank's end marker but the last gives the one-process stream
    lzj = b"".join(x[5][:-1] for x in res) + b"\x00"
    prop, want = helpers.oracle_lzma2_compress(whole)
    assert lzj == want and helpers.oracle_lzma2_decompress(lzj, len(whole), prop)[0] == whole
    # the price-based parses (stage C + stage Z / stage P): frames and dictionary-reset blocks are still the only units
    assert b"".join(x[6] for x in res) == helpers.oracle_compress(whole, flags=1 | 0x20)
    assert b"".join(x[7][:-1] for x in res) + b"\x00" == helpers.oracle_lzma2_compress(whole, flags=1 | (2 << 8) | 0x10)[1]
