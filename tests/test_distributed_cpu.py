"""CPU, world_size 2 over gloo: the two collectives the hot path adds to data-parallel training —
the packed statistics all-reduce inside ActivityMonitor.step() and the GroupNorm-gamma broadcast that keeps
replicas identical after a rank-0-only nudge (train.py:244-246,315-319)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import vcd_b200
        vcd_b200.add_src_to_path()
        from tracking.monitor import ActivityMonitor, _Target
        ops = vcd_b200.ops
        torch.manual_seed(0)
        vae = vcd_b200.B200AutoencoderKL()
        # 1) packed all-reduce of the running accumulators: every rank reports the mean over ALL ranks' forwards
        mon = ActivityMonitor(vae, {"enabled": True, "track_interval": 1, "target_layers": []})
        tgt = _Target("encoder.conv_norm_out.output", ["mean_abs_activation_per_channel", "mean_activation"])
        tgt.slot = ops.TrackSlot(4, "cpu")
        # rank r saw (r + 1) forwards whose per-forward mean|x| vectors sum to base * (r + 1)
        base = torch.tensor([1.0, 2.0, 3.0, 4.0])
        tgt.slot.run.view(5, 4)[0] = base * (rank + 1)
        tgt.slot.run.view(5, 4)[3] = base * (rank + 1)     # running max|x|: NOT additive across ranks
        tgt.slot.scal[:] = torch.tensor([0.5 * (rank + 1), 0.0, rank + 1.0], dtype=torch.float64)
        mon._targets[tgt.identifier] = tgt
        mon._fired.append(tgt.identifier)
        wb = mon.step(1)
        data = mon.get_data_for_step(1)[tgt.identifier]
        total_fwd = sum(r + 1 for r in range(world))
        want = (base * total_fwd / total_fwd).numpy()
        ok_stats = bool(abs(data["mean_abs_activation_per_channel"] - want).max() < 1e-6) and \
            abs(float(data["mean_activation"]) - 0.5) < 1e-9 and \
            f"tracking/{tgt.identifier}/mean_abs_activation_per_channel_overall_mean" in wb
        ext = mon.get_extended_stats_for_step(1)[tgt.identifier]
        ok_stats = ok_stats and bool(abs(ext["max_abs_per_channel"] - (base * world).numpy()).max() < 1e-6)
        # 2) gamma broadcast: rank 0 "nudges", step() marked the sync, next encode-side hook re-synchronises
        g = vae.decoder.conv_norm_out.weight
        if rank == 0:
            with torch.no_grad():
                g[::8] *= 1.2
        assert getattr(vae, "_gamma_sync_pending", False)
        vae._sync_gamma_if_pending()
        ref = torch.ones(128)
        ref[::8] = 1.2
        ok_gamma = bool(torch.allclose(g.detach(), ref)) and not vae._gamma_sync_pending
        q.put((rank, ok_stats, ok_gamma))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world_size_2_stats_allreduce_and_gamma_broadcast():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(60)
    assert res == [(0, True, True), (1, True, True)], res
