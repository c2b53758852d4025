"""CPU, world_size 2 over gloo: the N>1 host path (column partition, block-maxima all-reduce, tie
all-gather, owner column broadcast, Gini min all-reduce) of kover_b200.sharding, with checker-backed
shards standing in for the GPU shards, against the golden vectors minted from the real reference."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_learners_over_gloo(tmp_path, world):
    port = free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), KV_TEST_OUT=str(tmp_path), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_gloo_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(o.decode(errors="replace"))
    for rank, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (rank, o[-3000:])
    checks = [json.load(open(os.path.join(str(tmp_path), "rank%d.json" % r)))["checks"] for r in range(world)]
    assert len(set(checks)) == 1 and checks[0] > 40
