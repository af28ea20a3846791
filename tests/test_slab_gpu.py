"""2-GPU parity of the slab-decomposed fused substep against the 1-GPU plan (needs >= 2 GPUs;
run with `gpurun --gpus 2`).  The reference asserts the same property for its FFT_SLAB mode
(parallel result == serial result, test/tests/gradient/tests:11-30)."""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, n, nsub, mode, q, chunks=1, sync="barrier", inv_ctas=0):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["MRL_SLAB_CHUNKS"] = str(chunks)   # read at plan creation: forward phase in y-chunks
    os.environ["MRL_SLAB_SYNC"] = sync            # barrier between the phases / arrival counters per column block
    os.environ["MRL_SLAB_INV_CTAS"] = str(inv_ctas)
    # "peer": exchanges fused into the passes as bulk stores; "copy": staged layouts + copy-engine exchanges (the default)
    os.environ["MRL_SLAB_EXCHANGE"] = "copy" if mode == "copy" else "store"
    if mode == "copy":
        mode = "peer"
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from marlin_b200 import capi, slab
    from marlin_b200.capi import AB_BETA
    try:
        L = n * 8 * math.pi / 200
        ctx = slab.SlabContext(rank, capi.F64)
        ctx.use_torch_stream()
        ctx.domain_set_slab((n, n, n), (0,) * 3, (L,) * 3, rank, world)
        torch.manual_seed(0)
        full = torch.rand((n, n, n), dtype=torch.float64) * 0.12 + 0.44
        nyl, y0 = ctx.shape[1], ctx.rbegin[1]
        c = full[:, y0:y0 + nyl, :].contiguous().cuda()
        plan = slab.SlabPlan(ctx, (0.1, 0.0, 1.0), 0.2, -0.001, history=1, mode=mode)
        dt = 1e-3
        plan.substep(c, dt, AB_BETA[0], 0)
        plan.advance_state()
        for _ in range(nsub - 1):
            plan.substep(c, dt, AB_BETA[1], 1)
            plan.advance_state()
        torch.cuda.synchronize()
        err = None
        if rank == 0:
            # serial plan on the same device
            sctx = capi.Context(0, capi.F64)
            sctx.use_torch_stream()
            sctx.domain_set(3, (n, n, n), (0,) * 3, (L,) * 3)
            cs = full.cuda()
            sp = sctx.split_plan(double_well=(0.1, 0.0, 1.0), M_factor=0.2, L_factor=-0.001, history=1)
            sp.substep(cs, dt, AB_BETA[0], 0)
            sp.advance_state()
            for _ in range(nsub - 1):
                sp.substep(cs, dt, AB_BETA[1], 1)
                sp.advance_state()
            torch.cuda.synchronize()
            ref = cs[:, y0:y0 + nyl, :]
            err = float(torch.linalg.norm((c - ref).reshape(-1)) / torch.linalg.norm(ref.reshape(-1)))
            sp.close()
            sctx.close()
        plan.close()
        ctx.close()
        dist.barrier()
        q.put((rank, err))
    finally:
        dist.destroy_process_group()


CASES = [(128, "nccl", 1, "barrier", 0), (128, "peer", 1, "barrier", 0), (256, "peer", 1, "barrier", 0), (128, "peer", 2, "barrier", 0),
         (256, "peer", 4, "barrier", 0), (128, "peer", 1, "flags", 0), (256, "peer", 4, "flags", 0), (256, "peer", 1, "flags", 40),
         (512, "peer", 4, "flags", 48), (128, "copy", 1, "barrier", 0), (256, "copy", 4, "barrier", 0), (256, "copy", 2, "barrier", 0),
         (512, "copy", 4, "barrier", 0)]


@pytest.mark.parametrize("n,mode,chunks,sync,inv_ctas", CASES)
def test_slab_matches_single_gpu(n, mode, chunks, sync, inv_ctas):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8 if n >= 256 else 4)
    if world == 3 or 4 < world < 8:
        world = 2 if world == 3 else 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 1000) + (7 if mode == "peer" else 3 if mode == "copy" else 0) + n // 128 + 11 * chunks + (23 if sync == "flags" else 0) + inv_ctas
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, 6, mode, q, chunks, sync, inv_ctas)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res[0] is not None and res[0] < 1e-12, res
