"""Conformance run: the REFERENCE's own Python unit tests (graphlearn/python/{tests,sampler/tests,gsl/tests}/test_*.py) executed
against this package.  Nothing is copied: the test files and their helper (python/tests/utils.py) are loaded from the reference
checkout at run time, with ``graphlearn`` aliased to ``graphlearn_b200`` in ``sys.modules`` - exactly what a user does who
switches the import.

    python tools/run_reference_pytests.py [--ref /root/reference] [--pattern test_node] [-v]

Every test file runs in its own process (the reference's tests share global flags and tracker directories), on the CPU.
Prints one line per file and a summary; exit code 0 when everything that ran passed.
"""
import argparse
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

RUNNER = r'''
import importlib.util, os, sys, types, unittest
sys.path.insert(0, {root!r})
import graphlearn_b200 as gl
import graphlearn_b200.python as glp
import graphlearn_b200.python.nn.tf, graphlearn_b200.python.nn.pytorch
# alias every loaded module of the package: 'graphlearn.x.y' must be THE SAME module object as 'graphlearn_b200.x.y' (a second
# copy would break isinstance checks)
for name, mod in list(sys.modules.items()):
    if name == "graphlearn_b200" or name.startswith("graphlearn_b200."):
        sys.modules["graphlearn" + name[len("graphlearn_b200"):]] = mod
for name, mod in (("errors", "errors"), ("utils", "utils"), ("config", "config"), ("data", "data"), ("sampler", "sampler"), ("gsl", "gsl")):
    sys.modules["graphlearn.python." + name] = importlib.import_module("graphlearn_b200." + mod)
    setattr(glp, name, sys.modules["graphlearn.python." + name])
ref = {ref!r}
def attach(modname, m):
    sys.modules[modname] = m
    parent, _, leaf = modname.rpartition(".")
    if parent in sys.modules:
        try:
            setattr(sys.modules[parent], leaf, m)
        except Exception:
            pass
def load(modname, path):
    spec = importlib.util.spec_from_file_location(modname, path)
    m = importlib.util.module_from_spec(spec)
    attach(modname, m)
    spec.loader.exec_module(m)
    return m
for name in ("graphlearn.python.tests", "graphlearn.python.sampler.tests", "graphlearn.python.gsl.tests"):
    pkg = types.ModuleType(name); pkg.__path__ = []
    attach(name, pkg)
load("graphlearn.python.tests.utils", os.path.join(ref, "graphlearn/python/tests/utils.py"))
# other test files import their base classes as graphlearn.python.<dir>.tests.<module>: resolve those names to the reference files
import importlib.abc, importlib.machinery
class RefTests(importlib.abc.MetaPathFinder):
    DIRS = {{"graphlearn.python.tests": "graphlearn/python/tests", "graphlearn.python.sampler.tests": "graphlearn/python/sampler/tests",
            "graphlearn.python.gsl.tests": "graphlearn/python/gsl/tests"}}
    def find_spec(self, fullname, path, target=None):
        parent, _, leaf = fullname.rpartition(".")
        d = self.DIRS.get(parent)
        if d is None:
            return None
        f = os.path.join(ref, d, leaf + ".py")
        return importlib.util.spec_from_file_location(fullname, f) if os.path.exists(f) else None
sys.meta_path.insert(0, RefTests())
os.chdir({cwd!r})
gl.set_default_neighbor_id(0) if False else None
m = load("ref_test_module", {test!r})
suite = unittest.defaultTestLoader.loadTestsFromModule(m)
res = unittest.TextTestRunner(verbosity={verbosity}, stream=sys.stdout).run(suite)
print("RESULT run=%d failures=%d errors=%d skipped=%d" % (res.testsRun, len(res.failures), len(res.errors), len(res.skipped)))
'''


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--pattern", default="")
    ap.add_argument("-v", action="store_true")
    ap.add_argument("--timeout", type=int, default=300)
    ap.add_argument("--jobs", type=int, default=4, help="test files run concurrently (each in its own process)")
    ap.add_argument("--quick", action="store_true", help="a representative third of the files (used by the pytest wrapper)")
    a = ap.parse_args()
    py = os.path.join(a.ref, "graphlearn", "python")
    if not os.path.isdir(py):
        print("no reference checkout at", a.ref)
        return 0
    files = sorted(glob.glob(py + "/tests/test_*.py") + glob.glob(py + "/sampler/tests/test_*.py") + glob.glob(py + "/gsl/tests/test_*.py") +
                   glob.glob(py + "/nn/pytorch/data/test/test_*.py"))        # nn/tf tests need TensorFlow
    files = [f for f in files if a.pattern in os.path.basename(f)]
    QUICK = ("test_node_weighted_labeled_attributed", "test_edge_weighted_labeled_attributed", "test_node_iterate_gsl", "test_edge_shuffle_gsl",
             "test_node_query_attribute", "test_gsl_sampling", "test_gsl_traverse", "test_gsl_mask", "test_gsl_random_walk",
             "test_edge_weight_neighbor_sampling", "test_full_neighbor_sampling", "test_conditional_negative_sampling",
             "test_subgraph_sampling", "test_in_degree_neighbor_sampling", "test_dataset")
    if a.quick:
        files = [f for f in files if os.path.basename(f)[:-3] in QUICK]
    import tempfile
    tot = {"run": 0, "failures": 0, "errors": 0, "skipped": 0}
    bad_files = []
    def run_file(f):
        cwd = tempfile.mkdtemp(prefix="glb_reftest_")
        code = RUNNER.format(root=ROOT, ref=a.ref, cwd=cwd, test=f, verbosity=2 if a.v else 0)
        env = dict(os.environ, CUDA_VISIBLE_DEVICES="", GLB_TEST_DEVICE="cpu")
        try:
            p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=a.timeout, env=env)
            return p.stdout + p.stderr
        except subprocess.TimeoutExpired:
            return "TIMEOUT"
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=max(1, a.jobs)) as pool:      # every file has its own process and scratch directory
        outputs = list(pool.map(run_file, files))
    for f, out in zip(files, outputs):
        line = [l for l in out.splitlines() if l.startswith("RESULT")]
        rel = os.path.relpath(f, py)
        if line:
            kv = dict(x.split("=") for x in line[-1].split()[1:])
            for k in tot:
                tot[k] += int(kv[k])
            ok = int(kv["failures"]) == 0 and int(kv["errors"]) == 0
            print("%-55s %s  %s" % (rel, "ok  " if ok else "FAIL", line[-1][7:]))
            if not ok:
                bad_files.append(rel)
                if a.v:
                    print(out[-3000:])
        else:
            print("%-55s CRASH %s" % (rel, out.strip().splitlines()[-1][:160] if out.strip() else ""))
            bad_files.append(rel)
            if a.v:
                print(out[-3000:])
    # the reference's examples/basic scripts (gen_test_data.py + test_local.py: node / edge iteration over 2-hop queries, truncated
    # full sampling, conditional negatives, stats; test_subgraph.py: SEAL-style sub-graph sampler + GSL SubGraph;
    # test_local_temporal_{sampler,loader}.py on examples/data/gen_temporal_data.py; test_actor_local.py) - executed from a scratch
    # copy because they write next to themselves
    basic = os.path.join(a.ref, "graphlearn", "examples", "basic")
    scripts = [("test_local.py", "gen_test_data.py"), ("test_subgraph.py", "gen_test_data.py"), ("test_actor_local.py", "gen_test_data.py"),
               ("test_local_temporal_sampler.py", "gen_temporal_data.py"), ("test_local_temporal_loader.py", "gen_temporal_data.py")]
    if a.quick:
        scripts = scripts[:2]
    for script, gen in scripts:
        label = "examples/basic/" + script
        if a.pattern and a.pattern not in label.replace("/", "_"):
            continue
        import shutil
        d = tempfile.mkdtemp(prefix="glb_refbasic_")
        w = os.path.join(d, "basic")
        shutil.copytree(basic, w)
        shutil.copy(os.path.join(a.ref, "graphlearn", "examples", "data", "gen_temporal_data.py"), w)
        os.makedirs(os.path.join(w, "data"), exist_ok=True)
        env = dict(os.environ, CUDA_VISIBLE_DEVICES="", GLB_TEST_DEVICE="cpu")
        code = ("import sys, os, runpy\nsys.path.insert(0, %r)\nimport graphlearn_b200 as gl, graphlearn_b200.python as glp\n"
                "import graphlearn_b200.python.nn.tf, graphlearn_b200.python.nn.pytorch\n"
                "for n, m in list(sys.modules.items()):\n    if n == 'graphlearn_b200' or n.startswith('graphlearn_b200.'): sys.modules['graphlearn' + n[15:]] = m\n"
                "sys.path.insert(0, %r); os.chdir(%r)\n"
                "runpy.run_path(%r, run_name='__main__')\nsys.argv = [%r]\n"
                "runpy.run_path(%r, run_name='__main__')\nprint('BASIC_OK')\n" % (ROOT, w, w, gen, script, script))
        try:
            p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=a.timeout, env=env)
            ok = p.returncode == 0 and "BASIC_OK" in p.stdout
            out = p.stdout + p.stderr
        except subprocess.TimeoutExpired:
            ok, out = False, "TIMEOUT"
        print("%-55s %s" % (label, "ok" if ok else "FAIL"))
        if not ok:
            bad_files.append(label)
            if a.v:
                print(out[-3000:])
    # the reference's DISTRIBUTED worker-mode launch (run_dist_worker_mode_{fs,rpc}_tracker.sh): two plain processes, no torchrun -
    # each calls g.init(task_index=i, task_count=2, tracker=dir) or g.init(task_index=i, hosts="h:p,h:p")
    import socket

    def free_port():
        sk = socket.socket()
        sk.bind(("127.0.0.1", 0))
        pt = sk.getsockname()[1]
        sk.close()
        return pt
    for script, mk_args in (("test_dist_worker_mode_fs_tracker.py", lambda i, w, hp: ["--task_index=%d" % i, "--task_count=2", "--tracker=" + os.path.join(w, "tracker")]),
                            ("test_dist_worker_mode_rpc_tracker.py", lambda i, w, hp: ["--task_index=%d" % i, "--hosts=" + hp])):
        label = "examples/basic/" + script + " (2 workers)"
        if (a.pattern and a.pattern not in label.replace("/", "_")) or (a.quick and "rpc" in script):
            continue
        import shutil
        d = tempfile.mkdtemp(prefix="glb_refdist_")
        w = os.path.join(d, "basic")
        shutil.copytree(basic, w)
        os.makedirs(os.path.join(w, "data"), exist_ok=True)
        os.makedirs(os.path.join(w, "tracker"), exist_ok=True)
        env = dict(os.environ, CUDA_VISIBLE_DEVICES="", GLB_TEST_DEVICE="cpu")
        for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
            env.pop(k, None)
        pre = ("import sys, os, runpy\nsys.path.insert(0, %r)\nimport graphlearn_b200 as gl, graphlearn_b200.python as glp\n"
               "import graphlearn_b200.python.nn.tf, graphlearn_b200.python.nn.pytorch\n"
               "for n, m in list(sys.modules.items()):\n    if n == 'graphlearn_b200' or n.startswith('graphlearn_b200.'): sys.modules['graphlearn' + n[15:]] = m\n"
               "sys.path.insert(0, %r); os.chdir(%r)\n" % (ROOT, w, w))
        subprocess.run([sys.executable, "-c", pre + "runpy.run_path('gen_test_data.py', run_name='__main__')\n"], capture_output=True, env=env, timeout=a.timeout)
        hp = "127.0.0.1:%d,127.0.0.1:%d" % (free_port(), free_port())
        procs = []
        for i in range(2):
            code = pre + "sys.argv = %r\nrunpy.run_path(%r, run_name='__main__')\nprint('WORKER_OK')\n" % ([script] + mk_args(i, w, hp), script)
            procs.append(subprocess.Popen([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env))
        outs = []
        ok = True
        for pr in procs:
            try:
                o = pr.communicate(timeout=a.timeout)[0]
            except subprocess.TimeoutExpired:
                pr.kill()
                o = "TIMEOUT"
            outs.append(o)
            ok = ok and pr.returncode == 0 and "WORKER_OK" in o
        print("%-55s %s" % (label, "ok" if ok else "FAIL"))
        if not ok:
            bad_files.append(label)
            if a.v:
                print("\n".join(x[-2000:] for x in outs))
    # the reference's SERVER-mode launch (run_dist_server_mode_{fs,rpc}_tracker.sh): 2 graph servers + 3 clients as plain processes
    for script, mk_args in (("test_dist_server_mode_fs_tracker.py",
                             lambda job, i, w, hp: ["--server_count=2", "--client_count=3", "--tracker=" + os.path.join(w, "tracker"),
                                                    "--job_name=" + job, "--task_index=%d" % i]),
                            ("test_dist_server_mode_rpc_tracker.py",
                             lambda job, i, w, hp: ["--server=" + hp, "--client_count=3", "--job_name=" + job, "--task_index=%d" % i])):
        label = "examples/basic/" + script + " (2 servers + 3 clients)"
        if (a.pattern and a.pattern not in label.replace("/", "_")) or (a.quick and "rpc" in script):
            continue
        import shutil
        d = tempfile.mkdtemp(prefix="glb_refsrv_")
        w = os.path.join(d, "basic")
        shutil.copytree(basic, w)
        os.makedirs(os.path.join(w, "data"), exist_ok=True)
        os.makedirs(os.path.join(w, "tracker"), exist_ok=True)
        env = dict(os.environ, CUDA_VISIBLE_DEVICES="", GLB_TEST_DEVICE="cpu")
        for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
            env.pop(k, None)
        pre = ("import sys, os, runpy\nsys.path.insert(0, %r)\nimport graphlearn_b200 as gl, graphlearn_b200.python as glp\n"
               "import graphlearn_b200.python.nn.tf, graphlearn_b200.python.nn.pytorch\n"
               "for n, m in list(sys.modules.items()):\n    if n == 'graphlearn_b200' or n.startswith('graphlearn_b200.'): sys.modules['graphlearn' + n[15:]] = m\n"
               "sys.path.insert(0, %r); os.chdir(%r)\n" % (ROOT, w, w))
        subprocess.run([sys.executable, "-c", pre + "runpy.run_path('gen_test_data.py', run_name='__main__')\n"], capture_output=True, env=env, timeout=a.timeout)
        hp = "127.0.0.1:%d,127.0.0.1:%d" % (free_port(), free_port())
        procs = []
        for job, i in (("server", 0), ("server", 1), ("client", 0), ("client", 1), ("client", 2)):
            code = pre + "sys.argv = %r\nrunpy.run_path(%r, run_name='__main__')\nprint('PROC_OK')\n" % ([script] + mk_args(job, i, w, hp), script)
            procs.append(subprocess.Popen([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env))
            if job == "server":
                import time as _t
                _t.sleep(0.5)
        outs, ok = [], True
        for pr in procs:
            try:
                o = pr.communicate(timeout=a.timeout)[0]
            except subprocess.TimeoutExpired:
                pr.kill()
                o = "TIMEOUT"
            outs.append(o)
            ok = ok and pr.returncode == 0 and "PROC_OK" in o
        print("%-55s %s" % (label, "ok" if ok else "FAIL"))
        if not ok:
            bad_files.append(label)
            if a.v:
                print("\n".join(x[-1500:] for x in outs))
    # the reference's DGS tutorial data: its u2i generator (python/data/u2i/u2i_generator.py -> /tmp/u2i_gen) + its conf/u2i schema and
    # install-query files drive THIS service through the pattern-file loader (native parser and Python loader must agree)
    label = "dynamic_graph_service u2i generator + conf -> service"
    gen = os.path.join(a.ref, "dynamic_graph_service", "python", "data", "u2i", "u2i_generator.py")
    if os.path.exists(gen) and (not a.pattern or a.pattern in "dgs_u2i") and not a.quick:
        code = ("import json, sys, runpy\nsys.path.insert(0, %r)\nrunpy.run_path(%r, run_name='__main__')\n"
                "from graphlearn_b200.dgs import Schema, QueryPlan, DynamicGraphService, FileLoader\nbase = %r\n"
                "sch = Schema.from_json(base + '/schema.u2i.json'); iq = json.load(open(base + '/install_query.u2i.json'))\nres = []\n"
                "for native in (True, False):\n"
                "    svc = DynamicGraphService(sch.to_service_schema(capacity=64, feat_dims={'user': 10, 'item': 10}), device='cpu')\n"
                "    svc.install_query(0, QueryPlan.from_json(iq, sch))\n"
                "    n = FileLoader('/tmp/u2i_gen/streaming/u2i.pattern', sch, native=native).load('/tmp/u2i_gen/streaming/u2i.streaming', svc)\n"
                "    r = svc.run_query(0, [0, 1, 2])\n    res.append((n, r['hops'][0]['ids'].tolist(), r['hops'][1]['ids'][:3].tolist()))\n"
                "assert res[0] == res[1] and res[0][0] > 10000 and all(x >= 0 for x in res[0][1][0])\nprint('DGS_U2I_OK')\n"
                % (ROOT, gen, os.path.join(a.ref, "dynamic_graph_service", "conf", "u2i")))
        env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
        try:
            p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=a.timeout, env=env, cwd=tempfile.mkdtemp())
            ok, out = p.returncode == 0 and "DGS_U2I_OK" in p.stdout, p.stdout + p.stderr
        except subprocess.TimeoutExpired:
            ok, out = False, "TIMEOUT"
        print("%-55s %s" % (label, "ok" if ok else "FAIL"))
        if not ok:
            bad_files.append(label)
            if a.v:
                print(out[-2000:])
    print("TOTAL files=%d %s  not-clean: %s" % (len(files), tot, bad_files))
    return 0 if not bad_files else 1


if __name__ == "__main__":
    sys.exit(main())
