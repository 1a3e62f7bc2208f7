"""World-size-2 test of the only collective on the path (the partition-mask all-gather) on the gloo
backend, plus the planner's rank-independence: every rank must take identical FULL/REGION/SKIP decisions, otherwise
the all-gather at step warmup-1 would not line up."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank: int, world: int, port: int, out_dir: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from regione_b200 import flux_kontext as fk
        from regione_b200 import params
        from regione_b200.manager import RegionManager, plan_steps
        from standins.diffusers_like import FlowMatchEulerDiscreteScheduler
        import numpy as np

        L = 4096
        g = torch.Generator().manual_seed(100 + rank)
        raw = (torch.rand(L, generator=g) < 0.25 + 0.25 * rank).to(torch.uint8)
        # disabled: stays local
        fk.enable_mask_allgather(False)
        assert fk.allgather_masks(raw).shape == (1, L)
        fk.enable_mask_allgather(True)
        allm = fk.allgather_masks(raw)
        assert allm.shape == (world, L) and torch.equal(allm[rank], raw)
        for r in range(world):
            gr = torch.Generator().manual_seed(100 + r)
            expect = (torch.rand(L, generator=gr) < 0.25 + 0.25 * r).to(torch.uint8)
            assert torch.equal(allm[r], expect), f"rank {rank}: row {r} of the gathered mask is wrong"
        # the asynchronous form the scheduler step uses (side stream on CUDA; async work on gloo) + the manager's view
        pend = fk.allgather_masks_async(raw)
        mm = RegionManager()
        mm.batch_masks_pending = pend
        assert torch.equal(mm.batch_masks, allm) and mm.batch_masks is mm.batch_masks
        assert mm.batch_edited_counts() == [int(allm[r].sum()) for r in range(world)]
        # planner decisions depend only on the schedule -> identical on every rank
        m = RegionManager()
        m.set_parameters(dict(params.DEFAULTS["FluxKontextPipeline"], threshold=0.88))
        s = FlowMatchEulerDiscreteScheduler()
        s.set_timesteps(sigmas=np.linspace(1.0, 1 / 28, 28), mu=fk.calculate_shift(L))
        plan = torch.tensor([int(skip) for skip, _ in plan_steps(s.timesteps, params.GAMMA["FluxKontextPipeline"], m)])
        plans = [torch.zeros_like(plan) for _ in range(world)]
        dist.all_gather(plans, plan)
        assert all(torch.equal(p, plan) for p in plans)
        assert int(plan.sum()) == 14
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_mask_allgather_world2_gloo(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(2))
