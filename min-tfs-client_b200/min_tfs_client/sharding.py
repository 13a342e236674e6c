"""Sharding a batch of independent PredictRequests / PredictResponses over the GPUs of one box.

The reference serialises one request at a time on one core (``for k, v in input_dict.items(): ...
CopyFrom(ndarray_to_tensor_proto(v))``, requests.py:47-48, then ``SerializeToString`` at
prediction_service_pb2_grpc.py:52).  Every request is a self-contained message, so a batch cuts into
contiguous blocks - request ``r`` of ``n`` belongs to GPU ``r // ceil(n / G)`` - with
no exchange step: there is no collective on the data path, only the caller's own join.

Two ways to use it:

* one process per GPU (``torchrun``; what ``bench.py --gpus N`` does): every rank calls
  ``shard(n, world, rank)`` and feeds its block to its own ``Codec`` / C-ABI context;
* one process driving several GPUs: ``ShardedCodec`` keeps one ``Codec`` and one host thread per GPU
  (ctypes releases the GIL inside the library, so the per-GPU planners, copies and kernels overlap) and
  returns results in request order.
"""
from __future__ import annotations

import queue
import threading
from typing import Callable, List, Optional, Sequence

from . import _native as N


def per_shard(n_items: int, world: int) -> int:
    """Block size: ceil(n / G) (0 for an empty batch)."""
    if world <= 0:
        raise ValueError("world must be positive")
    return -(-int(n_items) // world) if n_items > 0 else 0


def shard(n_items: int, world: int, rank: int) -> range:
    """The contiguous block of ``range(n_items)`` that shard ``rank`` of ``world`` owns."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside [0, {world})")
    per = per_shard(n_items, world)
    return range(min(rank * per, n_items), min((rank + 1) * per, n_items))


def owner(index: int, n_items: int, world: int) -> int:
    """Which shard request ``index`` lands on: ``index // ceil(n / G)``."""
    if not 0 <= index < n_items:
        raise IndexError(index)
    return index // per_shard(n_items, world)


class _Worker(threading.Thread):
    """One host thread bound to one GPU; owns that GPU's Codec (native contexts are not re-entrant)."""

    def __init__(self, device: int):
        super().__init__(daemon=True, name=f"b200tfs-shard-gpu{device}")
        self.device = device
        self.jobs: "queue.Queue" = queue.Queue()
        self.codec = None
        self.start()

    def run(self):
        from .codec import Codec

        while True:
            job = self.jobs.get()
            if job is None:
                break
            fn, box, done = job
            try:
                if self.codec is None:
                    self.codec = Codec(self.device)   # raises without a GPU: there is no CPU path behind this
                box.append(("ok", fn(self.codec)))
            except BaseException as exc:  # noqa: BLE001 - handed to the caller's thread
                box.append(("err", exc))
            done.set()
        if self.codec is not None:
            self.codec.close()

    def submit(self, fn: Callable):
        box, done = [], threading.Event()
        self.jobs.put((fn, box, done))
        return box, done


class ShardedCodec:
    """A batch API over every GPU of the box: same calls as ``Codec``, results in request order.

    ``devices`` lists the CUDA ordinals to use (default: all visible); an ordinal may repeat, which gives
    that GPU two independent contexts / host threads.
    """

    def __init__(self, devices: Optional[Sequence[int]] = None):
        if devices is None:
            n = N.device_count()
            if n == 0:
                raise RuntimeError("no CUDA device: min_tfs_client has no CPU codec to fall back to")
            devices = list(range(n))
        if not devices:
            raise ValueError("devices is empty")
        self.devices = [int(d) for d in devices]
        self._workers = [_Worker(d) for d in self.devices]

    @property
    def world(self) -> int:
        return len(self._workers)

    def close(self):
        for w in self._workers:
            w.jobs.put(None)
        for w in self._workers:
            w.join(timeout=30)
        self._workers = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _scatter(self, items: Sequence, call: Callable) -> List:
        """Run ``call(codec, block)`` on every shard's block, concatenate in order."""
        n = len(items)
        pending = []
        for rank, w in enumerate(self._workers):
            block = shard(n, self.world, rank)
            if len(block) == 0:
                continue
            part = items[block.start: block.stop]
            pending.append(w.submit(lambda codec, part=part: call(codec, part)))
        out: List = []
        error = None
        for box, done in pending:
            done.wait()
            kind, val = box[0]
            if kind == "err":
                error = error or val
            else:
                out.extend(val)
        if error is not None:
            raise error
        return out

    # same signatures as Codec's batch entry points
    def encode_predict_requests(self, requests, **kw) -> List[bytes]:
        reqs = list(requests)
        return self._scatter(reqs, lambda codec, part: codec.encode_predict_requests(part, **kw))

    def decode_predict_responses(self, wires: Sequence[bytes], **kw):
        return self._scatter(list(wires), lambda codec, part: codec.decode_predict_responses(part, **kw))

    def encode_tensor_protos(self, arrays: Sequence, **kw) -> List[bytes]:
        return self._scatter(list(arrays), lambda codec, part: codec.encode_tensor_protos(part, **kw))

    def decode_tensor_protos(self, wires: Sequence[bytes], **kw):
        return self._scatter(list(wires), lambda codec, part: codec.decode_tensor_protos(part, **kw))
