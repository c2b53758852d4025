"""GPU, needs >= 2 GPUs (skipped otherwise): one process per GPU over NCCL, column shards, both exchange
paths (peer mailbox over NVLink, NCCL all-gather) against the CPU checker.  Run with
`gpurun --gpus 2 -- python -m pytest tests/test_multigpu_nccl.py -m gpu`."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def n_gpus():
    try:
        from kover_b200 import _native
        return _native.device_count()
    except Exception:  # noqa: BLE001
        return 0


@pytest.mark.skipif(n_gpus() < 2, reason="needs at least 2 GPUs")
def test_sharded_learners_over_nccl():
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_nccl_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(o.decode(errors="replace"))
    for rank, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (rank, o[-3000:])
    assert "NCCL_WORKER_OK" in outs[0]
    print(outs[0][-300:])
