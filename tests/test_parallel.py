"""CPU suite: the N>1 host logic (frame sharding + counter all-reduce) with world_size 2 on gloo."""
import os
import subprocess
import sys
import textwrap

from conftest import ROOT
from informationbottleneckdecodingldpc_b200.parallel import allreduce_counters, shard_frames


def test_shard_frames_partition():
    for total, world in ((16384, 8), (100, 3), (5, 8), (0, 2)):
        spans = [shard_frames(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def test_allreduce_is_identity_without_a_group():
    assert allreduce_counters([3, 1, 10, 500]) == [3, 1, 10, 500]


def test_counter_allreduce_world2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r})
        from informationbottleneckdecodingldpc_b200.parallel import init_distributed, allreduce_counters, shard_frames
        rank, world, local = init_distributed(2, backend="gloo")
        lo, hi = shard_frames(1001, rank, world)
        # every rank contributes its own shard's counters; all ranks must see the same totals
        tot = allreduce_counters([rank + 1, 10 * (rank + 1), hi - lo, 50 * (hi - lo)], world)
        assert tot == [3, 30, 1001, 50 * 1001], tot
        # the BER loop's stop decision is therefore identical on every rank
        errors, rounds = 0, 0
        while errors < 7:
            errors += allreduce_counters([rank + 1])[0]
            rounds += 1
        assert (errors, rounds) == (9, 3)
        print("rank", rank, "ok")
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_cpulist_parser_and_numa_binding_without_topology(tmp_path):
    """bind_to_gpu_numa_node must be a no-op (None) when sysfs does not expose the GPU; the cpulist parser
    handles ranges and singletons."""
    from informationbottleneckdecodingldpc_b200 import parallel
    assert parallel._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert parallel._parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    assert parallel.bind_to_gpu_numa_node(0, sysfs=str(tmp_path)) is None
    assert os.sched_getaffinity(0) == before
