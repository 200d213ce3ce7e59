"""Multi-GPU parity worker, launched one rank per GPU by tests/test_gpu_multi.py:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P tests/mgpu_worker.py

Every rank hands the FULL graph to libssba; the library shards the landmarks (and their edges)
by rank and all-reduces the reduced pose system over NCCL each LM trial.  Checked against the
golden fixtures of the compiled reference and against a single-GPU run on rank 0.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from common import CHI2_RTOL, golden_case, golden_scalars, rel  # noqa: E402
from ssvio_b200 import ba, synth  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    idt = torch.zeros(ba.SSBA_NCCL_ID_BYTES, dtype=torch.uint8, device=dev)
    if rank == 0:
        idt = torch.tensor(list(ba.nccl_unique_id()), dtype=torch.uint8, device=dev)
    dist.broadcast(idt, 0)
    nccl_id = bytes(idt.cpu().tolist())
    opt = ba.BundleAdjuster(device_id=local, rank=rank, world_size=world, nccl_id=nccl_id)
    gold = golden_scalars()
    cases = [("small_fixed", golden_case("small_fixed")[0]), ("cfg1", golden_case("cfg1")[0]),
             ("cfg2", synth.make_config("cfg2")), ("cfg3", synth.make_config("cfg3"))]
    for name, g in cases:
        opt.set_graph(g)
        opt.initialize_optimization()
        rep = opt.optimize(g.iters)
        poses, points, errs = opt.poses(), opt.points(), opt.edge_errors()
        plain, robust = opt.chi2()
        nout, nin = opt.count_outliers(5.891)
        assert rep.iterations == gold[name]["numeric"]["iterations"], (name, rep.iterations)
        assert rel(rep.chi2_robust, gold[name]["numeric"]["chi2_robust"]) < CHI2_RTOL, name
        assert rel(rep.chi2_robust, gold[name]["analytic"]["chi2_robust"]) < 1e-9, name
        assert rel(robust, rep.chi2_robust) < 1e-12 and rel(plain, rep.chi2_plain) < 1e-12
        e2 = (errs ** 2).sum(1)
        assert rel(e2.sum(), rep.chi2_plain) < 1e-11, name
        assert (nout, nin) == (int((e2 > 5.891).sum()), int((e2 <= 5.891).sum())), name
        # every rank holds the same poses and the same gathered points
        t = torch.from_numpy(np.concatenate([poses.ravel(), points.ravel()])).to(dev)
        t0 = t.clone()
        dist.broadcast(t0, 0)
        assert torch.equal(t, t0), f"{name}: ranks disagree on the estimates"
        if rank == 0:
            with ba.BundleAdjuster(device_id=local) as single:
                single.set_graph(g)
                r1 = single.optimize(g.iters)
                assert rel(rep.chi2_robust, r1.chi2_robust) < 1e-11
                np.testing.assert_allclose(poses, single.poses(), atol=1e-9)
                np.testing.assert_allclose(points, single.points(), atol=1e-8)
            print(f"[mgpu x{world}] {name}: chi2={rep.chi2_robust:.8f} rel_vs_ref={rel(rep.chi2_robust, gold[name]['numeric']['chi2_robust']):.2e} OK",
                  flush=True)
        dist.barrier()
    opt.close()

    # ---- pre-sharded input (ssba_options.presharded): every rank hands over only the edges of its own landmarks;
    # the ranks agree on the active sets and the co-visibility pattern inside ssba_initialize
    idt = torch.zeros(ba.SSBA_NCCL_ID_BYTES, dtype=torch.uint8, device=dev)
    if rank == 0:
        idt = torch.tensor(list(ba.nccl_unique_id()), dtype=torch.uint8, device=dev)
    dist.broadcast(idt, 0)
    opt = ba.BundleAdjuster(device_id=local, rank=rank, world_size=world, nccl_id=bytes(idt.cpu().tolist()), presharded=True)
    for name, g in cases:
        mine = g.shard(rank, world)
        for attempt in range(2):  # the second pass re-uses the resident structure (same indices): values only
            opt.set_graph(mine)
            rep = opt.optimize(g.iters)
        poses, points, errs = opt.poses(), opt.points(), opt.edge_errors()
        info = opt.problem_info()
        assert info.n_structure_reuses >= 1, name
        assert rep.iterations == gold[name]["numeric"]["iterations"], (name, rep.iterations)
        assert rel(rep.chi2_robust, gold[name]["numeric"]["chi2_robust"]) < CHI2_RTOL, name
        assert rel(rep.chi2_robust, gold[name]["analytic"]["chi2_robust"]) < 1e-9, name
        nout, nin = opt.count_outliers(5.891)
        assert nout + nin == g.n_edges, (name, nout, nin, g.n_edges)
        assert errs.shape[0] == mine.n_edges
        t = torch.from_numpy(np.concatenate([poses.ravel(), points.ravel()])).to(dev)
        t0 = t.clone()
        dist.broadcast(t0, 0)
        assert torch.equal(t, t0), f"{name}: ranks disagree on the estimates (pre-sharded)"
        with ba.BundleAdjuster(device_id=local) as single:
            single.set_graph(g)
            r1 = single.optimize(g.iters)
            assert rel(rep.chi2_robust, r1.chi2_robust) < 1e-11
            np.testing.assert_allclose(poses, single.poses(), atol=1e-9)
            np.testing.assert_allclose(points, single.points(), atol=1e-8)
            # this rank's edges, in its own order, against the same edges of the single-GPU run
            sel = np.nonzero(np.isin(g.point_idx, np.unique(mine.point_idx)))[0]
            np.testing.assert_allclose(errs, single.edge_errors()[sel], atol=1e-8)
        if rank == 0:
            print(f"[mgpu x{world}, pre-sharded] {name}: chi2={rep.chi2_robust:.8f} rel_vs_ref={rel(rep.chi2_robust, gold[name]['numeric']['chi2_robust']):.2e} OK", flush=True)
        dist.barrier()
    opt.close()
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_OK", flush=True)


if __name__ == "__main__":
    main()
