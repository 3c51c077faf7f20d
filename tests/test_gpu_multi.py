"""Multi-GPU parity: the sharded paths reproduce the single-GPU result.  The tests that need
two GPUs skip on a one-GPU box; the single-GPU tests cover the exchange format and the empty-shard case."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def test_device_exchange_equals_host_fuse(seqj, handle):
    """seqj_sequence_shard + seqj_fuse_dev (group results stay on the device) against the one-call sequence():
    two 'virtual ranks' on one GPU pack their groups into one exchange buffer; order, eta and D are identical."""
    import torch
    from sequencerj_b200 import api, sharding
    N, L = 700, 384
    scales, metrics = (1, 2, 4, 8), seqj.ALL_METRICS
    At = np.maximum(np.random.default_rng(21).random((N, L)), 2.3e-16)
    ref = api._sequence(At, scales, metrics, handle)
    dev = torch.from_numpy(At).cuda()
    world = 2
    owner = sharding.plan_groups(L, metrics, scales, world)
    cap = int(max((owner == r).sum() for r in range(world)))
    allp = torch.zeros((world * cap, 2 * N), dtype=torch.float64, device="cuda")
    assert handle.lib.seqj_pack_stride(N) == 2 * N
    torch.cuda.synchronize()
    parts = []
    for r in range(world):
        mask = (owner == r).astype(np.uint8)
        parts.append(api._sequence(At, scales, metrics, handle, group_mask=mask, device_ptr=dev.data_ptr(),
                                   pack=allp[r * cap:(r + 1) * cap]))
    fin = api.fuse_dev(N, owner.size, allp, sharding.slots_of_groups(owner, cap), handle=handle)
    assert np.array_equal(fin.order, ref.order) and fin.eta == ref.eta and fin.root == ref.root
    def canon(t):                     # the triplets are a set: first-occurrence order depends on the group order
        i, j, v = t
        o = np.lexsort((j, i))
        return i[o], j[o], v[o]
    (fi, fj, fv), (ri, rj, rv) = canon(fin.D_triplets), canon(ref.D_triplets)
    assert np.array_equal(fi, ri) and np.array_equal(fj, rj)
    # values to rounding only: a rank that owns the coarse part of a cut chain computes its first scale by tiles where
    # the single call derived it (~1e-13 relative in the matrix entries, hence in the MST weights behind D)
    assert np.allclose(fv, rv, rtol=1e-9, atol=0.0)
    for p in parts:
        for key, (eta, tree) in p.EOAlgScale.items():
            assert eta == ref.EOAlgScale[key][0] and np.array_equal(tree.parents, ref.EOAlgScale[key][1].parents)


def test_empty_shard_is_a_valid_call(seqj, handle):
    """A rank that was assigned nothing (chunk-major with fewer tasks than ranks; an all-zero group
    mask) gets an empty result, not SEQJ_E_OOM, and bad input is still reported."""
    import torch
    from sequencerj_b200 import api
    N, L = 300, 64
    At = np.maximum(np.random.default_rng(22).random((N, L)), 2.3e-16)
    dev = torch.from_numpy(At).cuda()
    S = torch.zeros((1, N * (-(-N // 16) * 16)), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    G = 4
    z = np.zeros(G, dtype=np.int32)
    r = api._sequence_partial((N, L), (1,), seqj.ALL_METRICS, z, z, S.data_ptr(), S.shape[1], dev.data_ptr(), handle)
    assert r.EOSeg == {} and r.EOAlgScale == {}
    r = api._sequence(At, (1,), seqj.ALL_METRICS, handle, group_mask=np.zeros(G, dtype=np.uint8), device_ptr=dev.data_ptr())
    assert r.EOSeg == {} and r.EOAlgScale == {}
    bad = torch.from_numpy(-At).cuda()
    torch.cuda.synchronize()
    with pytest.raises(AssertionError):
        api._sequence(At, (1,), seqj.ALL_METRICS, handle, group_mask=np.zeros(G, dtype=np.uint8), device_ptr=bad.data_ptr())


@pytest.mark.parametrize("shape", [(600, 512, "1,2,4,8"), (1500, 1024, "1,2,4,8,16")])
def test_torchrun_sharded_paths_match_single_gpu(shape):
    """Two ranks (one process per GPU, NCCL): group-major with the device-side exchange (resident input and sharded
    upload) and chunk-major with the N x N reduce both reproduce the single-GPU call on every rank."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    N, L, scales = shape
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "_mgpu_worker.py"), str(N), str(L), scales]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-4000:]
    assert res.stdout.count("match=True") == 6, res.stdout[-4000:]


def test_single_process_multi_device_handle(seqj):
    """seqj_create(devices, ndev = 2): one process (the Julia host's situation) drives two GPUs; results equal the
    single-device handle's bit for bit, including per-chunk and per-group details."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    A = np.asfortranarray(np.random.default_rng(5).random((512, 900)))
    scales = (1, 2, 4, 8)
    h1, hm = seqj.Handle(0), seqj.Handle([0, 1])
    try:
        r1 = seqj.sequence(A, scales=scales, metrics=seqj.ALL_METRICS, handle=h1)
        rm = seqj.sequence(A, scales=scales, metrics=seqj.ALL_METRICS, handle=hm)
        assert np.array_equal(r1.order, rm.order) and r1.eta == rm.eta
        assert r1.EOAlgScale.keys() == rm.EOAlgScale.keys()
        for key in r1.EOAlgScale:
            assert r1.EOAlgScale[key][0] == rm.EOAlgScale[key][0]
            assert np.array_equal(r1.EOAlgScale[key][1].parents, rm.EOAlgScale[key][1].parents)
            assert r1.EOSeg[key][0] == rm.EOSeg[key][0]
    finally:
        hm.close(); h1.close()
