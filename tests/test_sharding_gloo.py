"""Multi-GPU plan.  The hot path has no collective - requests shard by index - so what runs on more than
one rank is: the product's block partition (``min_tfs_client.sharding.shard``), each rank's share through the
product (on CPU: the host planner of libb200tfs.so, which computes every framing byte of a request;
on the GPU box: the device codec behind ``ShardedCodec``), and the control plane ``bench.py`` uses
(barrier, MAX of the per-rank times, SUM of the per-rank bytes).  World size 2 over ``gloo`` here."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path[:0] = [os.path.join(REPO, "min-tfs-client_b200"), REPO, HERE]

from min_tfs_client import _native as N  # noqa: E402
from min_tfs_client.codec import _Prepared  # noqa: E402
from min_tfs_client.sharding import ShardedCodec, owner, per_shard, shard  # noqa: E402


def image(r):
    return np.random.default_rng(r).standard_normal((3, 8, 8), dtype=np.float32)


def planned_request_wire(model_name, model_version, inputs):
    """PredictRequest bytes from the PRODUCT's host planner (every framing byte) + the payloads laid in by numpy;
    fixed-width inputs only (no sNaNs in these inputs, so the float32 payload is the raw memory)."""
    lib = N.load()
    preps = [_Prepared(a, k.encode(), None, False, False) for k, a in inputs]
    arr = (N.Tensor * max(len(preps), 1))(*[p.struct for p in preps])
    name = model_name.encode()
    req = N.Request(model_name=name, model_name_len=len(name), has_version=int(model_version is not None), order=N.ORDER_UPB,
                    version=model_version or 0, n_inputs=len(preps), flags=0, inputs=arr)
    cap = 1 << 16
    buf = C.create_string_buffer(cap)
    flen = C.c_uint64()
    n = max(len(preps), 1)
    poff, plen, perm = (C.c_uint64 * n)(), (C.c_uint64 * n)(), (C.c_int32 * n)()
    N.check(lib.b200tfs_request_frame(C.byref(req), buf, cap, C.byref(flen), poff, plen, perm))
    frame, wire, fpos = buf.raw[: flen.value], bytearray(), 0
    for j in range(len(preps)):
        take = poff[j] - len(wire)
        wire += frame[fpos: fpos + take]
        fpos += take
        wire += preps[perm[j]].array.tobytes()
    return bytes(wire + frame[fpos:])


def _worker(rank, world, port, n_requests, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard(n_requests, world, rank)                              # the product's partition
    wires = [planned_request_wire("default", 1, [("image", image(r))]) for r in mine]   # the product's planner
    dist.barrier()
    t = torch.tensor([0.010 + 0.005 * rank], dtype=torch.float64)      # stand-in for this rank's device time
    nbytes = torch.tensor([float(sum(len(w) for w in wires))], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(nbytes, op=dist.ReduceOp.SUM)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.array([t.item(), nbytes.item(), mine.start, mine.stop]))
    with open(os.path.join(out_dir, f"w{rank}.bin"), "wb") as fh:
        for w in wires:
            fh.write(len(w).to_bytes(4, "little") + w)
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_two_rank_sharding(tmp_path):
    n_requests, world = 11, 2
    mp.spawn(_worker, args=(world, _free_port(), n_requests, str(tmp_path)), nprocs=world, join=True)
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
    for r, w in enumerate(got):                                         # the shares reassemble to the batch, in order, bit-exact
        assert w == wire_oracle.encode_predict_request("default", 1, [("image", image(r))])
    assert stats[0][1] == sum(len(w) for w in got)                      # SUM over ranks
    assert [tuple(map(int, s[2:])) for s in stats] == [(0, 6), (6, 11)]


def test_shard_covers_everything_once():
    for n in (0, 1, 7, 8192):
        for world in (1, 2, 4, 8):
            seen = [i for r in range(world) for i in shard(n, world, r)]
            assert seen == list(range(n))
            assert all(owner(i, n, world) == r for r in range(world) for i in shard(n, world, r))
    assert per_shard(8192, 8) == 1024 and list(shard(8192, 8, 3))[:1] == [3072]      # BASELINE configs[4]: request r -> GPU r // 1024
    with pytest.raises(ValueError):
        shard(4, 2, 2)


def test_sharded_codec_refuses_without_a_gpu():
    if N.device_count():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        ShardedCodec()
    sc = ShardedCodec(devices=[0, 0])       # explicit ordinals: the failure surfaces from the worker's Codec(0)
    with pytest.raises(RuntimeError):
        sc.encode_predict_requests([("m", 1, {"x": np.zeros(3, np.float32)})])
    sc.close()


@pytest.mark.gpu
def test_sharded_codec_on_the_device():
    """Two shards on the visible GPU(s) (two contexts / host threads when the box has one): a batch of requests encoded and
    the matching responses decoded through ShardedCodec equal the oracle, in request order; uneven shares and empty shares."""
    from oracle import wire_oracle

    n_gpu = N.device_count()
    devices = list(range(n_gpu)) if n_gpu >= 2 else [0, 0]
    with ShardedCodec(devices) as sc:
        assert sc.world == len(devices)
        for n in (1, 7, 64):
            reqs = [("default", 1, [("image", image(r)), ("label", np.array([r % 1000], dtype=np.int64))]) for r in range(n)]
            wires = sc.encode_predict_requests(reqs)
            assert len(wires) == n
            for r, w in enumerate(wires):
                assert w == wire_oracle.encode_predict_request("default", 1, reqs[r][2])
            resp = [wire_oracle.build_predict_response([("scores", image(10000 + r))], "default", 1, "serving_default") for r in range(n)]
            outs = sc.decode_predict_responses(resp)
            assert len(outs) == n
            for r, (arrays, spec) in enumerate(outs):
                assert arrays["scores"].tobytes() == image(10000 + r).tobytes() and spec.name == "default"
        assert sc.encode_predict_requests([]) == []


@pytest.mark.gpu
def test_sharded_codec_batch_kernel_on_every_device():
    """Shares large enough (80 x 602 KB per GPU) for the TMA-staged batch decode kernel, whose opt-in to > 48 KB of dynamic shared
    memory is per device: every GPU of the process must get it, not only the first one that launched."""
    from oracle import wire_oracle

    n_gpu = N.device_count()
    devices = list(range(n_gpu)) if n_gpu >= 2 else [0, 0]
    imgs = [image(20000 + k) for k in range(5)]
    wires = [wire_oracle.build_predict_response([("scores", imgs[k])], "default", 1, "serving_default") for k in range(5)]
    n = 80 * len(devices)
    with ShardedCodec(devices) as sc:
        for _ in range(2):
            outs = sc.decode_predict_responses([wires[r % 5] for r in range(n)])
            assert len(outs) == n
            for r, (arrays, spec) in enumerate(outs):
                assert arrays["scores"].tobytes() == imgs[r % 5].tobytes(), r
