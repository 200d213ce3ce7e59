"""Host logic of the product: structure build, landmark sharding, symbolic factorisation and the
level schedule the CUDA solver walks — exercised on the CPU by tests/cpp/test_structure.cpp, which also
interprets the PACKED device programs (the subtree-per-CTA program of k_tree_solve with a race check per barrier
interval; the level program of k_reduced_solve: rounds per CTA, slots, reader masks, look-ahead) against a dense solve."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_structure_builder_cpp():
    exe = "/tmp/ssba_test_structure"
    cuda_inc = "/usr/local/cuda/include"
    subprocess.run(["g++", "-O2", "-std=c++17", "-I" + cuda_inc, "-I" + os.path.join(ROOT, "include"),
                    "-I" + os.path.join(ROOT, "ssvio_b200", "csrc"),
                    os.path.join(ROOT, "tests", "cpp", "test_structure.cpp"),
                    os.path.join(ROOT, "ssvio_b200", "csrc", "ssba_structure.cpp"),
                    os.path.join(ROOT, "ssvio_b200", "csrc", "ssba_tree_program.cpp"), "-o", exe, "-lpthread"], check=True)
    # every cluster size the device solver can be built for (the packed program differs: rounds per CTA,
    # reader masks, look-ahead placement)
    # ... for both solver programs: the subtree-per-CTA program of k_tree_solve (default; with the 8- and the
    # 16-CTA cluster cap) and the level program of k_reduced_solve (SSBA_SOLVER=level, also the fallback)
    for solver, cap in (("", "8"), ("", "16"), ("level", "8")):
        for cluster in ("", "1", "2", "4", "8"):
            env = dict(os.environ)
            env["SSBA_TREE_CLUSTER_CAP"] = cap
            if solver:
                env["SSBA_SOLVER"] = solver
            if cluster:
                env["SSBA_SOLVE_CLUSTER"] = cluster
            r = subprocess.run([exe], capture_output=True, text=True, env=env)
            assert r.returncode == 0, f"solver={solver or 'tree'} cap={cap} cluster={cluster or 'default'}\n" + r.stdout + r.stderr
            assert r.stdout.strip().endswith("OK")


def test_structure_build_is_independent_of_the_thread_count():
    """The parallel passes of build_structure (forced on for the small test graphs with SSBA_HOST_PAR_MIN) give the
    same structures, programs and statistics as the single-thread build: the report of the C++ test is identical."""
    exe = "/tmp/ssba_test_structure"
    if not os.path.exists(exe):
        test_structure_builder_cpp()
    outs = []
    for threads, par_min in (("1", ""), ("4", "64"), ("7", "64")):
        env = dict(os.environ)
        env["SSBA_HOST_THREADS"] = threads
        if par_min:
            env["SSBA_HOST_PAR_MIN"] = par_min
        r = subprocess.run([exe], capture_output=True, text=True, env=env)
        assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout[-2000:] + r.stderr[-2000:]
        outs.append(r.stdout)
    assert outs[0] == outs[1] == outs[2]
