"""Multi-GPU entry points of the C ABI (csrc/group.cu): tsr_group_* (one process, several GPUs) and tsr_comm_* (one
process per GPU), both summing {shots, logical errors, low confidence} with one ncclAllReduce.  The single-GPU cases run
on any GPU box; the 2-GPU cases need `gpurun --gpus 2` (skipped otherwise).  tests/test_sharding.py covers the host-side
partition logic over gloo on CPU."""

import numpy as np
import pytest

import workloads as W
from tesseract_decoder_b200 import capi, sharding

pytestmark = pytest.mark.gpu


def device_count() -> int:
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 1


def check_group(devices, shots=5000, chunk=512):
    dem, kw = W.decoder_kwargs(W.CONFIGS["C1"], capi.build_det_orders)
    single = capi.Decoder(dem, device=devices[0], **kw)
    dets, obs = capi.sample_dem_b8(dem, 909, shots, single.num_detectors, single.num_observables)
    ref = single.decode_batch_b8(dets)
    g = capi.Group(dem, devices, **kw)
    assert g.size == len(devices) and all(m.comm_world == len(devices) for m in g.members)
    out = g.decode_batch_b8(dets, obs_true=obs, chunk_shots=chunk, want_error_use=True)
    for k in ("obs", "cost", "low_confidence", "err_offsets", "err_idx"):
        assert np.array_equal(out[k], ref[k]), k
    n, wrong, low = sharding.count_errors(ref["obs"], obs, ref["low_confidence"])
    assert out["tally"] == {"shots": n, "logical_errors": wrong, "low_confidence": low}
    use = np.zeros(single.num_dem_errors, dtype=np.uint64)
    lists = W.split_lists(ref["err_offsets"], ref["err_idx"])
    good = ~(ref["obs"] != obs).any(axis=1) & (ref["low_confidence"] == 0)
    for s in np.nonzero(good)[0]:
        for e in lists[s]:
            use[e] += 1
    assert np.array_equal(out["error_use"], use)
    # without truth every shot "matches"; ragged tail chunk; chunk larger than the batch
    out2 = g.decode_batch_b8(dets[:777], chunk_shots=4096)
    assert np.array_equal(out2["obs"], ref["obs"][:777]) and out2["tally"]["shots"] == 777 and out2["tally"]["logical_errors"] == 0
    assert g.decode_batch_b8(dets[:0])["tally"]["shots"] == 0
    g.close()


def test_group_of_one_device_matches_the_plain_decoder():
    check_group([0])


@pytest.mark.skipif(device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_group_of_two_devices_matches_the_plain_decoder():
    check_group([0, 1])
    check_group([1, 0], shots=3001, chunk=100)


def test_comm_of_one_rank_and_errors():
    dem, kw = W.decoder_kwargs(W.CONFIGS["C1"], capi.build_det_orders)
    dec = capi.Decoder(dem, device=0, **kw)
    assert dec.comm_world == 0
    with pytest.raises(ValueError, match="no communicator"):
        dec.allreduce_sum([1, 2, 3])
    dec.comm_init_rank(capi.comm_unique_id(), 0, 1)
    assert dec.comm_world == 1
    assert dec.allreduce_sum([5, 0, 2 ** 40]).tolist() == [5, 0, 2 ** 40]
    with pytest.raises(ValueError):
        dec.comm_init_rank(capi.comm_unique_id(), 1, 1)
    assert capi.lib().tsr_nccl_version() >= 22000
    with pytest.raises(ValueError, match="does not exist"):
        capi.Group(dem, [0, 99], **kw)
    with pytest.raises(ValueError, match="only once"):
        capi.Group(dem, [0, 0], **kw)


def _rank_worker(rank, world, uid, q):
    dem, kw = W.decoder_kwargs(W.CONFIGS["C1"], capi.build_det_orders)
    dec = capi.Decoder(dem, device=rank, **kw)
    dec.comm_init_rank(uid, rank, world)
    dets, obs = capi.sample_dem_b8(dem, 4321, 6000, 120, 1)
    mine, truth = sharding.shard_rows(dets, rank, world, 256), sharding.shard_rows(obs, rank, world, 256)
    out = dec.decode_batch_b8(mine, want_errors=False)
    local = sharding.count_errors(out["obs"], truth, out["low_confidence"])
    q.put((rank, dec.allreduce_sum(list(local)).tolist()))


@pytest.mark.skipif(device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_ranks_one_process_each_sum_their_counters_over_nccl():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    uid = capi.comm_unique_id()
    procs = [ctx.Process(target=_rank_worker, args=(r, 2, uid, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    dem, kw = W.decoder_kwargs(W.CONFIGS["C1"], capi.build_det_orders)
    dets, obs = capi.sample_dem_b8(dem, 4321, 6000, 120, 1)
    ref = capi.Decoder(dem, device=0, **kw).decode_batch_b8(dets, want_errors=False)
    expect = list(sharding.count_errors(ref["obs"], obs, ref["low_confidence"]))
    assert got[0] == expect and got[1] == expect
