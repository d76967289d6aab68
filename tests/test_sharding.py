"""The N>1 host logic (shot sharding + the single all-reduce of the counters) on CPU: world_size 2 over gloo.
The shards are decoded by the CPU checker here (no GPU in this container); tests/test_gpu_parity.py covers the
decoder itself."""
import socket

import numpy as np
import pytest

import workloads as W
from tesseract_decoder_b200 import capi, sharding


def test_shard_chunks_partition_the_shots():
    for n, world, chunk in [(0, 2, 4), (10, 3, 4), (4096 * 5 + 17, 8, 4096), (7, 8, 1)]:
        seen = []
        for r in range(world):
            seen += [i for b, e in sharding.shard_chunks(n, r, world, chunk) for i in range(b, e)]
        assert sorted(seen) == list(range(n))
    assert sharding.shard_chunks(10, 1, 2, 4) == [(4, 8)]
    with pytest.raises(ValueError):
        sharding.shard_chunks(10, 2, 2)


def test_count_errors_follows_the_cli_accounting():  # tesseract_main.cc:597-601
    pred = np.array([[1], [0], [1], [0]], dtype=np.uint8)
    true = np.array([[1], [1], [0], [0]], dtype=np.uint8)
    low = np.array([0, 0, 1, 0], dtype=np.uint8)
    assert sharding.count_errors(pred, true, low) == (4, 1, 1)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from oracle.oracle import OracleDecoder, available
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        dem, kw = W.decoder_kwargs(W.CONFIGS["C1"], capi.build_det_orders)
        dets, obs = capi.sample_dem_b8(dem, 4321, 600, 120, 1)
        dec = OracleDecoder(dem, kind="reference" if available("reference") else "port", **kw)
        totals, out = sharding.decode_sharded(dec.decode_batch_b8, dets, obs, rank, world, chunk=64)
        q.put((rank, totals, len(out["obs"])))
    finally:
        dist.destroy_process_group()


def test_two_ranks_over_gloo_agree_with_one_rank():
    import torch.multiprocessing as mp
    from oracle.oracle import OracleDecoder, available
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    dem, kw = W.decoder_kwargs(W.CONFIGS["C1"], capi.build_det_orders)
    dets, obs = capi.sample_dem_b8(dem, 4321, 600, 120, 1)
    ref = OracleDecoder(dem, kind="reference" if available("reference") else "port", **kw).decode_batch_b8(dets)
    expect = sharding.count_errors(ref["obs"], obs, ref["low_confidence"])
    assert all(t == expect for _, t, _ in got)          # every rank holds the whole-job totals
    assert sorted(n for _, _, n in got) == [280, 320]   # chunks of 64 interleaved: 5 vs 5 chunks (last one short)
