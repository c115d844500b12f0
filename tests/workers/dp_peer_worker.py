"""Data-parallel train step with the one-shot peer all-reduce (qfa_peer_allreduce) on >= 2 GPUs, one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
        tests/workers/dp_peer_worker.py
Checks (every rank asserts, rank 0 prints DP-PEER-OK):
  * enable_data_parallel() really selected the peer kernel (symmetric memory came up on every rank);
  * the summed accumulator is bit-identical on all ranks, equals the rank-ordered sum of the gathered local accumulators
    bit for bit, and agrees with ncclAllReduce to float round-off;
  * the captured train-step graph (gather -> accumulate -> peer all-reduce -> Adam + clip) leaves bit-identical parameters on
    all ranks after two epochs and tracks a second model that runs the same steps with NCCL."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from qfa_b200 import QFA, Adam, DeviceDataloader, step_scheduler, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import datetime
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("QFA_WORKER_WATCHDOG", "240")), exit=True)   # a hang ends with a traceback
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    log = (lambda *a: print("[rank %d]" % rank, *a, flush=True))
    for grid_name, Nh, prec in (("sdss", 8, "mixed"), ("l32", 32, "mixed"), ("sdss", 8, "fp64")):
        grid = synth.GRIDS[grid_name]
        P, mu = synth.smooth_random_params(grid, Nh, seed=31)
        Pn = {k: v.numpy() for k, v in P.items()}
        B = 1536 if prec != "fp64" else 96
        d = synth.make_spectra(P, mu, grid, B, seed=500 + rank, device=dev)
        ins = [d["delta"], d["error"], d["zabs"], d["mask"].view(torch.uint8)]
        m = QFA(grid.Nb, grid.Nr, Nh, dev, model_params=Pn, precision=prec)
        m.enable_data_parallel()
        log(grid_name, prec, "allreduce:", m.allreduce_kind)
        assert m.allreduce_kind.startswith("qfa_peer_allreduce"), m.allreduce_kind
        for step in range(5):
            sl = [t[step * 7:] for t in ins]                    # a different batch every step
            loc = m.accumulate(*sl, zero=True).clone()
            nccl = loc.clone()
            dist.all_reduce(nccl)
            mine = loc.clone()
            m._allreduce(mine)
            parts = [torch.empty_like(loc) for _ in range(world)]
            dist.all_gather(parts, loc)
            tot = parts[0].clone()
            for q in range(1, world):
                tot += parts[q]
            assert torch.equal(mine, tot), (grid_name, prec, step, "rank-ordered sum")
            got = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(got, mine)
            assert all(torch.equal(g, got[0]) for g in got), (grid_name, prec, step, "ranks differ")
            scale = nccl.abs().max()
            assert float((mine - nccl).abs().max() / scale) < (1e-6 if prec != "fp64" else 1e-14), (grid_name, prec, step)
        log(grid_name, prec, "sums ok")
    # ---- the captured train step
    grid = synth.GRIDS["sdss"]
    P, mu = synth.smooth_random_params(grid, 8, seed=32)
    Pn = {k: v.numpy() for k, v in P.items()}
    rows, bs = 4096, 1024
    d = synth.make_spectra(P, mu, grid, rows, seed=900 + rank, device=dev)
    final = {}
    for peer in (True, False):
        m = QFA(grid.Nb, grid.Nr, 8, dev, model_params=Pn, precision="mixed")
        m.mu = mu
        m.enable_data_parallel(peer_allreduce=peer)
        assert m.allreduce_kind.startswith("qfa_peer_allreduce") == peer, m.allreduce_kind
        dl = DeviceDataloader(d["flux"], d["error"], d["zqso"], d["mask"], grid.wav(), batch_size=bs, device=dev,
                              shuffle=True, seed=5)
        opt = Adam(params=m.parameters, device=dev, scheduler=step_scheduler(0.9, 10), learning_rate=1e-3, weight_decay=0.1)
        losses = []
        for _ in range(2):
            dl.rewind()
            losses.append(m._graphed_epoch(opt, dl, rows // bs))
            opt.step()
        log("graphed epochs done, peer =", peer, losses)
        p = m._params.clone()
        got = [torch.empty_like(p) for _ in range(world)]
        dist.all_gather(got, p)
        assert all(torch.equal(g, got[0]) for g in got), ("parameters differ between ranks", peer)
        assert all(l == l for l in losses), losses
        final[peer] = (p, losses)
    dp = float((final[True][0] - final[False][0]).abs().max())
    assert dp < 2e-4, dp                                         # 8 Adam steps of lr 1e-3: same trajectory
    assert abs(final[True][1][-1] - final[False][1][-1]) < 1e-3 * abs(final[False][1][-1]) + 1e-3, final
    dist.barrier()
    if rank == 0:
        print("DP-PEER-OK world=%d max|dparam| peer vs nccl = %.2e losses %s" % (world, dp, final[True][1]), flush=True)
    # ncclCommDestroy waits for every CUDA graph that captured one of its collectives (the NCCL arm above): drop them first
    import gc
    del m, opt, dl
    final.clear()
    gc.collect()
    torch.cuda.synchronize()
    faulthandler.cancel_dump_traceback_later()
    faulthandler.dump_traceback_later(30, exit=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
