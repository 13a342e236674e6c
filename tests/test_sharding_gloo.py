"""Multi-GPU plan on CPU: world_size-2 gloo.  The hot path has no collective - requests shard by
index - so what is tested is exactly the control plane bench.py uses: contiguous request sharding,
barrier, MAX-reduce of per-rank times, SUM-reduce of per-rank bytes, and that each rank's share of a
batch (encoded here by the CPU oracle standing in for the device) reassembles to the full batch."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(HERE)]


def shard(n_requests, world, rank):
    """Contiguous block per rank (request r -> GPU r // ceil(N/G))."""
    per = -(-n_requests // world)
    return range(min(rank * per, n_requests), min((rank + 1) * per, n_requests))


def _worker(rank, world, port, n_requests, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import wire_oracle

    mine = shard(n_requests, world, rank)
    wires = []
    for r in mine:
        x = np.random.default_rng(r).standard_normal((3, 8, 8), dtype=np.float32)
        wires.append(wire_oracle.encode_predict_request("default", 1, [("image", x)]))
    dist.barrier()
    t = torch.tensor([0.010 + 0.005 * rank], dtype=torch.float64)      # pretend device time of this rank
    nbytes = torch.tensor([float(sum(len(w) for w in wires))], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(nbytes, op=dist.ReduceOp.SUM)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.array([t.item(), nbytes.item(), mine.start, mine.stop]))
    with open(os.path.join(out_dir, f"w{rank}.bin"), "wb") as fh:
        for w in wires:
            fh.write(len(w).to_bytes(4, "little") + w)
    dist.destroy_process_group()


def test_two_rank_sharding(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    n_requests, world = 11, 2
    mp.spawn(_worker, args=(world, port, n_requests, str(tmp_path)), nprocs=world, join=True)
    from oracle import wire_oracle

    stats = [np.load(tmp_path / f"r{r}.npy") for r in range(world)]
    assert all(abs(s[0] - 0.015) < 1e-12 for s in stats)                # MAX over ranks
    got = []
    for r in range(world):
        blob = (tmp_path / f"w{r}.bin").read_bytes()
        i = 0
        while i < len(blob):
            n = int.from_bytes(blob[i:i + 4], "little")
            got.append(blob[i + 4:i + 4 + n])
            i += 4 + n
    assert len(got) == n_requests
    for r, w in enumerate(got):                                         # shards reassemble to the full batch, in order
        x = np.random.default_rng(r).standard_normal((3, 8, 8), dtype=np.float32)
        assert w == wire_oracle.encode_predict_request("default", 1, [("image", x)])
    assert stats[0][1] == sum(len(w) for w in got)                      # SUM over ranks
    assert [tuple(map(int, s[2:])) for s in stats] == [(0, 6), (6, 11)]


def test_shard_covers_everything_once():
    for n in (0, 1, 7, 8192):
        for world in (1, 2, 4, 8):
            seen = [i for r in range(world) for i in shard(n, world, r)]
            assert seen == list(range(n))
