"""Multi-rank tests: the same SPMD worker runs on CPU (gloo, world_size 2 - portable
partition/all-to-all path) and on GPUs (NCCL bootstrap + peer-memory kernels)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(nproc, port, env_extra, script="dist_worker.py", args=(), token="DIST_WORKER_OK"):
    env = dict(os.environ)
    env.update(env_extra)
    env["PYTHONPATH"] = ROOT
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", script)] + list(args)
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    out = p.stdout + p.stderr
    assert p.returncode == 0 and token in out, out[-4000:]


def test_two_ranks_cpu_gloo():
    _run(2, 29641, {"GLB_TEST_DEVICE": "cpu", "CUDA_VISIBLE_DEVICES": ""})


def test_two_ranks_cpu_gloo_feature_cache():
    _run(2, 29644, {"GLB_TEST_DEVICE": "cpu", "CUDA_VISIBLE_DEVICES": "", "GLB_TEST_CACHE": "1000"})


def test_public_api_two_ranks_cpu_gloo(tmp_path):
    from tests import fixtures as fx
    d = fx.write_graph(str(tmp_path))
    _run(2, 29647, {"GLB_TEST_DEVICE": "cpu", "CUDA_VISIBLE_DEVICES": ""}, "dist_api_worker.py", [d], "DIST_API_OK")


def test_sharded_embedding_two_ranks_cpu_gloo():
    _run(2, 29650, {"CUDA_VISIBLE_DEVICES": ""}, "dist_embedding_worker.py", [], "EMB_ALL_OK")


def test_worker_mode_with_tracker_directory_no_torchrun(tmp_path):
    """The reference's worker-mode launch: N PLAIN processes, each calling g.init(task_index=i, task_count=N, tracker=dir) - the
    tracker directory is the rendezvous (parallel/runtime.py:bootstrap_cluster), no torchrun, no RANK / WORLD_SIZE variables."""
    from tests import fixtures as fx
    d = fx.write_graph(str(tmp_path / "g"))
    tracker = str(tmp_path / "tracker")
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT")}
    env.update(PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES="")
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "tracker_worker.py"), d, tracker, str(i), "2"], env=env,
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for i in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for i, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and "TRACKER_WORKER_OK %d" % i in o, o[-3000:]


def test_server_mode_two_independent_servers_with_tracker(tmp_path):
    """Server mode launched the reference's way: 2 server + 2 client PLAIN processes, cluster = {server_count, client_count,
    tracker}.  The servers publish their endpoints under the tracker directory, hold the whole graph and traverse their hash
    share of the ids: the union of what the two clients see is every user / every edge exactly once; look-ups of any id and
    full-neighbour (sparse) values work on either server."""
    import json
    from tests import fixtures as fx
    d = fx.write_graph(str(tmp_path / "g"))
    tracker = str(tmp_path / "tracker")
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT")}
    env.update(PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES="")
    procs = []
    for job, i in (("server", 0), ("server", 1), ("client", 0), ("client", 1)):
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "server_mode_proc.py"), d, tracker, job, str(i),
                                       str(tmp_path / ("out_%s_%d.json" % (job, i)))], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0 and "SERVER_MODE_PROC_OK" in o, o[-3000:]
    res = [json.load(open(str(tmp_path / ("out_client_%d.json" % i)))) for i in range(2)]
    users = res[0]["users"] + res[1]["users"]
    assert sorted(users) == list(range(fx.N_USER))                                     # every user exactly once over the clients
    assert set(res[0]["users"]) == {u for u in range(fx.N_USER) if u % 2 == 0}        # client 0 <-> server 0 <-> even ids
    adj = fx.u2i_adj()
    want_edges = sorted((u, i) for u in adj for i, _ in adj[u])
    assert sorted(map(tuple, res[0]["edges"] + res[1]["edges"])) == want_edges
    assert res[0]["full"] == [len(adj[u]) for u in res[0]["users"][:5]]
    assert res[0]["stats"]["user"] == [fx.N_USER]


def test_nn_utils_two_ranks_cpu_gloo():
    _run(2, 29653, {"CUDA_VISIBLE_DEVICES": ""}, "dist_nn_utils_worker.py", [], "NN_UTILS_OK")


def test_data_parallel_example_two_ranks_cpu_gloo():
    """examples/train_gcn_sparse.py under torchrun: masks + full neighbourhoods + generic Trainer (flat gradients,
    all-reduce, lock-step epochs) - replicas must end with identical parameters."""
    env = dict(os.environ, GLB_TEST_DEVICE="cpu", CUDA_VISIBLE_DEVICES="", PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29649", os.path.join(ROOT, "examples", "train_gcn_sparse.py"), "--device", "cpu",
           "--epochs", "2", "--nodes", "500"]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600, cwd=os.path.join(ROOT, "examples"))
    assert p.returncode == 0 and "replicas in sync on 2 ranks" in p.stdout, (p.stdout + p.stderr)[-3000:]


@pytest.mark.gpu
@pytest.mark.multigpu
def test_public_api_two_ranks_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from tests import fixtures as fx
    d = fx.write_graph(str(tmp_path))
    _run(2, 29648, {}, "dist_api_worker.py", [d], "DIST_API_OK")


@pytest.mark.gpu
@pytest.mark.multigpu
def test_two_ranks_gpu_peer():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    _run(2, 29642, {})


@pytest.mark.gpu
@pytest.mark.multigpu
def test_two_ranks_gpu_peer_feature_cache():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    _run(2, 29645, {"GLB_TEST_CACHE": "1000"})           # partial cache: vid -> slot map inside the kernels
    _run(2, 29646, {"GLB_TEST_CACHE": "100000"})         # full replica: peer slots point at local copies


@pytest.mark.gpu
@pytest.mark.multigpu
def test_eight_ranks_gpu_peer():
    if torch.cuda.device_count() < 8:
        pytest.skip("needs 8 GPUs")
    _run(8, 29643, {})
