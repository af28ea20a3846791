"""Generic decomposed transforms (mrl_dist_*: DomainAction::partitionSlabs / fftSlab / ifftSlab and partitionPencils /
fftPencil / ifftPencil for any grid size, unequal parts) against torch.fft on the CPU: the local reciprocal block must
equal the rank's slice of the serial transform, the round trip the local real part (the reference asserts parallel ==
serial for its FFT_SLAB and FFT_PENCIL modes, test/tests/gradient/tests:11-30).

One process per rank; the ranks are spread over the GPUs of the box and SHARE devices when there are fewer GPUs than
ranks (CUDA IPC works between processes on one device; the device-side barriers then wait for the driver's time
slicing), so the 2-, 3- and 4-rank cases also run on a one-GPU box.  The IPC records travel over gloo."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

SLAB_SHAPES = [(20, 20), (9, 7), (16, 12, 10), (9, 7, 6), (128, 64, 32), (40, 40, 40)]
PENCIL_SHAPES = [(8, 8, 8), (9, 7, 6), (40, 40, 40), (64, 32, 16)]


def _worker(rank, world, port, q, f32=False):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = rank % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from marlin_b200 import capi
    errs = {}
    try:
        cases = [("slab", s, w) for s in SLAB_SHAPES for w in (None, [1 + (r % 2) for r in range(world)])]
        if world >= 4:
            cases += [("pencil", s, None) for s in PENCIL_SHAPES]
        if f32:
            cases = [("slab", (20, 20), None), ("slab", (16, 12, 10), None), ("slab", (128, 64, 32), None)] + \
                    ([("pencil", (8, 8, 8), None), ("pencil", (64, 32, 16), None)] if world >= 4 else [])
        for mode, shape, weights in cases:
            dim = len(shape)
            ctx = capi.DistContext(dev, capi.F32 if f32 else capi.F64)
            ctx.use_torch_stream()
            lens = (1.0, 2.0, 3.0)
            if mode == "slab":
                ctx.domain_set_dist(dim, shape, (0,) * 3, lens, rank, world, weights)
            else:
                ctx.domain_set_pencil(shape, (0,) * 3, lens, rank, world)
            d = capi.Dist(ctx)
            torch.manual_seed(5)
            full = torch.rand((2,) + tuple(shape), dtype=torch.float64)            # batch of two fields
            sp = list(range(1, dim + 1))
            if mode == "slab":
                ref = torch.fft.rfftn(full, dim=sp)                                # half spectrum on the last axis
            else:
                ref = torch.fft.fftn(torch.fft.rfft(full, dim=1), dim=(2, 3))      # half spectrum on x (fftPencil)
            rsl = (slice(None),) + tuple(slice(b, b + n) for b, n in zip(ctx.rbegin, ctx.shape))
            ksl = (slice(None),) + tuple(slice(b, b + n) for b, n in zip(ctx.kbegin, ctx.rshape))
            if mode == "slab":
                # slab sizes follow partitionHepler
                cnt = (capi.C.c_int64 * world)()
                w = (capi.C.c_double * world)(*[float(v) for v in weights]) if weights else None
                capi._ck(capi.lib().mrl_partition(capi.C.c_int64(shape[1]), world, w, cnt))
                assert ctx.shape[1] == cnt[rank] and ctx.rbegin[1] == sum(cnt[:rank])
            loc = full[rsl].contiguous().cuda().to(ctx.rdtype)
            spec = d.rfftn(loc)
            want = ref[ksl]
            assert tuple(spec.shape) == tuple(want.shape), (mode, shape, spec.shape, want.shape)
            e1 = float((spec.cpu().to(torch.complex128) - want).abs().max() / ref.abs().max())
            back = d.irfftn(spec)
            e2 = float((back - loc).abs().max())
            # local axes are the slices of the global ones
            for a in range(dim if not f32 else 0):
                g = capi.axis_values(shape[a], 0.0, lens[a])
                assert torch.equal(ctx.axis(a, False), torch.tensor(g[ctx.rbegin[a]:ctx.rbegin[a] + ctx.shape[a]], dtype=torch.float64))
                half = (a == dim - 1) if mode == "slab" else (a == 0)
                gk = capi.axis_values(shape[a], 0.0, lens[a], reciprocal=True, half=half)
                assert torch.equal(ctx.axis(a, True), torch.tensor(gk[ctx.kbegin[a]:ctx.kbegin[a] + ctx.rshape[a]], dtype=torch.float64))
            errs[(mode, shape, bool(weights))] = (e1, e2)
            d.close()
            ctx.close()
        q.put((rank, errs))
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("world,f32", [(1, False), (2, False), (3, False), (4, False), (4, True)])
def test_dist_transforms_match_serial(world, f32):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() * 5 + world + (17 if f32 else 0)) % 1000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, f32)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, errs in res.items():
        for key, (e1, e2) in errs.items():
            tol = 2e-6 if f32 else 1e-13
            assert e1 < tol and e2 < tol, (rank, key, e1, e2)
