"""Host logic of the multi-GPU path on CPU: assignment properties and a world_size-2 gloo run."""
import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_assignment_covers_every_group_once_and_balances(seqj):
    from sequencerj_b200 import sharding
    metrics, scales = seqj.ALL_METRICS, (1, 2, 4, 8, 16)
    costs = sharding.group_costs(4096, metrics, scales)
    assert costs.shape == (4, 5) and np.all(costs > 0)
    assert costs[1, 0] > costs[0, 0]                       # EMD costs twice the FP64 issue slots of a contraction
    for world in (1, 2, 4, 8):
        owner = sharding.assign_groups(costs, world)
        assert owner.shape == costs.shape and owner.min() >= 0 and owner.max() < world
        load = np.array([costs[owner == r].sum() for r in range(world)])
        assert load.min() > 0
        assert load.max() <= costs.sum() / world + costs.max()          # LPT bound
    assert sharding.nchunks(10, 6) == 5 and sharding.nchunks(4096, 16) == 16 and sharding.nchunks(50, 4) == 4
    # nested scales: every contraction metric stays on one rank (its coarse scales are derived there)
    assert sharding.scales_chain_nested(4096, scales) and not sharding.scales_chain_nested(50, (1, 2, 4))
    for world in (2, 4, 8):
        owner = sharding.plan_groups(4096, metrics, scales, world)
        for k, m in enumerate(metrics):
            if m.id != 1:
                assert len(set(owner[k].tolist())) == 1
        assert len(set(owner.reshape(-1).tolist())) == world


def test_merge_is_metric_major(seqj):
    from sequencerj_b200 import sharding
    metrics, scales = (seqj.L2, seqj.KLD), (1, 2)
    parts = [{(0, 2): (2.0, [0], [1], [0.5]), (2, 1): (3.0, [0], [1], [0.7])},
             {(0, 1): (1.0, [0], [1], [0.4]), (2, 2): (4.0, [0], [1], [0.9])}]
    etas, msts = sharding.merge_group_results(parts, metrics, scales)
    assert etas == [1.0, 2.0, 3.0, 4.0]                    # (L2,1), (L2,2), (KLD,1), (KLD,2): sequencer.jl:194-196
    assert [t.weight[0] for t in msts] == [0.4, 0.5, 0.7, 0.9]


def test_world_size_2_gloo(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "result.txt"
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   SEQJ_TEST_OUT=str(out), OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_gloo_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=300)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert out.read_text() == "OK"


def test_chunk_partition_covers_every_chunk_once_even_with_idle_ranks(seqj):
    """Chunk-major split: every (group, chunk) task belongs to exactly one rank, ranges are contiguous per rank, and a
    rank may legitimately end up with nothing (fewer tasks than ranks): the library then returns an empty result
    (seqj_sequence_partial with empty ranges) so that the rank still joins every collective."""
    from sequencerj_b200 import sharding
    for M, scales, world in ((4096, (1, 2, 4, 8, 16), 8), (50, (1,), 8), (53, (3, 6, 7), 5), (2048, (1, 2, 4, 8, 16, 32), 8)):
        metrics = seqj.ALL_METRICS
        lo, hi = sharding.chunk_partition(M, metrics, scales, world)
        G = len(metrics) * len(scales)
        assert lo.shape == hi.shape == (world, G)
        for g in range(G):
            nch = sharding.nchunks(M, scales[g % len(scales)])
            cover = np.zeros(nch, dtype=int)
            for r in range(world):
                assert 0 <= lo[r, g] <= hi[r, g] <= nch or lo[r, g] >= hi[r, g]
                if lo[r, g] < hi[r, g]:
                    cover[lo[r, g]:hi[r, g]] += 1
            assert np.all(cover == 1), (M, scales, world, g, cover)
    lo, hi = sharding.chunk_partition(50, seqj.ALL_METRICS, (1,), 8)
    assert any(not (lo[r] < hi[r]).any() for r in range(8))          # the case ADVICE r01 flagged: idle ranks exist


def test_shard_plan_cuts_a_chain_only_when_it_balances(seqj):
    """seqj_plan_shards (csrc/host_shard.cuh): config 3 keeps every contraction metric whole at 2 / 4 / 8 ranks (eight
    units of ~equal cost); config 4 (EMD + Energy, scales 1..32) on 8 ranks cuts the Energy chain into consecutive
    sub-chains so that no rank carries all 63 Energy matrices; config 5 (two chains) uses 4 ranks by cutting both."""
    from sequencerj_b200 import sharding
    for world in (2, 4, 8):
        owner, load = sharding.plan_groups(4096, seqj.ALL_METRICS, (1, 2, 4, 8, 16), world, with_load=True)
        assert load.min() > 0 and load.sum() / (world * load.max()) > 0.9
    scales = (1, 2, 4, 8, 16, 32)
    owner, load = sharding.plan_groups(2048, (seqj.WASS1D, seqj.ENERGY), scales, 8, with_load=True)
    assert owner.shape == (2, 6) and owner.min() == 0 and owner.max() == 7
    assert len(set(owner[0].tolist())) == 6                       # every EMD group on its own rank
    en = owner[1].tolist()                                        # scales 1..32; the chain runs 32 -> 1
    assert len(set(en)) >= 2
    runs = [en[i] for i in range(len(en)) if i == 0 or en[i] != en[i - 1]]
    assert len(runs) == len(set(runs)), "a rank's Energy scales must be consecutive (a sub-chain)"
    single = sharding.plan_groups(2048, (seqj.WASS1D, seqj.ENERGY), scales, 1, with_load=True)[1].max()
    assert single / load.max() > 5.0                              # modelled speed-up on 8 GPUs
    owner, load = sharding.plan_groups(512, (seqj.L2, seqj.KLD), (1, 2, 4), 4, with_load=True)
    assert len(set(owner.reshape(-1).tolist())) == 4
    owner1 = sharding.plan_groups(512, (seqj.L2, seqj.KLD), (1, 2, 4), 1)
    assert np.all(owner1 == 0)
