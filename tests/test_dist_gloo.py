"""world_size 2-4 gloo tests (CPU) of the decomposition rules behind [Domain] parallel_mode = FFT_SLAB / FFT_PENCIL.

The CUDA kernels cannot run here; torch.fft stands in for the local passes (test code, not product code).  Under test is
the block arithmetic marlin_b200/csrc/mrl_dist.cu implements - who owns which rows, where a block lands - with the
partition rules taken from the library itself (mrl_partition = partitionHepler, mrl_pencil_factors = the Py x Pz choice of
partitionPencils; both host-only entry points):

  slab    real [nx][ny_r](,[nz]) -> local passes on z (r2c) and x -> x-block s to rank s, landing at y = ybegin[me]..
          -> y pass -> reciprocal [nx_r][ny](,[nz/2+1]) = the x-slice of the serial rfftn (unequal parts, weights, 2-D)
  pencil  real [nx][ny_a][nz_b], rank = b Py + a -> rfft along x -> kx-block a' to rank (b, a') at y = ybegin[a]..
          -> fft along y -> ky-block b' to rank (b', a) at z = zbegin[b].. -> fft along z
          -> reciprocal [(nx/2+1)_a][ny_b][nz]: fftPencil's layout (src/actions/DomainAction.C:1022-1034, :1106-1256)

Composing these steps over gloo must give every rank its slice of the serial transform, and the way back the local real
part (the reference asserts parallel == serial for both modes, test/tests/gradient/tests:11-30)."""
import ctypes as C
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _partition(total, parts, weights=None):
    from marlin_b200 import capi
    cnt = (C.c_int64 * parts)()
    w = (C.c_double * parts)(*[float(v) for v in weights]) if weights else None
    capi._ck(capi.lib().mrl_partition(C.c_int64(total), parts, w, cnt))
    cnt = list(cnt)
    return cnt, [sum(cnt[:i]) for i in range(parts)]


def _send_blocks(blocks, group_ranks, me):
    """shape + payload to every other member (pairs with the receives of _exchange on their side)"""
    for i, dst in enumerate(group_ranks):
        if dst == me:
            continue
        shape = torch.full((4,), -1, dtype=torch.int64)
        shape[:blocks[i].dim()] = torch.tensor(blocks[i].shape)
        dist.send(shape, dst)
        dist.send(blocks[i].contiguous(), dst)


def _all_to_all(blocks, group_ranks, me):
    """unequal-block all-to-all inside a group: rank order fixed, sends before receives per source to stay deadlock free"""
    got = [None] * len(group_ranks)
    for k, src in enumerate(group_ranks):
        if src == me:
            _send_blocks(blocks, group_ranks, me)
            got[k] = blocks[group_ranks.index(me)].clone()
        else:
            shape = torch.empty(4, dtype=torch.int64)
            dist.recv(shape, src)
            t = torch.empty(tuple(int(v) for v in shape if v >= 0), dtype=torch.complex128)
            dist.recv(t, src)
            got[k] = t
    return got


def _slab_worker(rank, world, port, shape, weights, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dim = len(shape)
        nx, ny = shape[0], shape[1]
        ycnt, ybeg = _partition(ny, world, weights)
        xcnt, xbeg = _partition(nx, world, weights)
        torch.manual_seed(3)
        full = torch.rand(shape, dtype=torch.float64)
        ref = torch.fft.rfftn(full)
        loc = full[:, ybeg[rank]:ybeg[rank] + ycnt[rank]]
        ranks = list(range(world))
        # forward: local z (r2c) and x passes, x-block s to rank s, concatenated along y in rank order, y pass
        a = torch.fft.fft(torch.fft.rfft(loc, dim=2) if dim == 3 else loc.to(torch.complex128), dim=0)
        got = _all_to_all([a[xbeg[s]:xbeg[s] + xcnt[s]] for s in ranks], ranks, rank)
        spec = torch.fft.fft(torch.cat(got, dim=1), dim=1)
        if dim == 2:
            spec = spec[:, :ny // 2 + 1]                       # the half ky <= ny/2: x-slice of the serial rfft2 layout
        err_f = float((spec - ref[xbeg[rank]:xbeg[rank] + xcnt[rank]]).abs().max() / ref.abs().max())
        # inverse: (2-D: one-sided form, the conjugate half lives on other ranks) y pass, y-block s back to rank s at
        # x = xbeg[me].., x pass, z c2r / real part
        if dim == 2:
            full_ky = torch.zeros((xcnt[rank], ny), dtype=torch.complex128)
            full_ky[:, :ny // 2 + 1] = spec * 2
            full_ky[:, 0] = spec[:, 0]
            if ny % 2 == 0:
                full_ky[:, ny // 2] = spec[:, ny // 2]
            b = torch.fft.ifft(full_ky, dim=1)
        else:
            b = torch.fft.ifft(spec, dim=1)
        got = _all_to_all([b[:, ybeg[s]:ybeg[s] + ycnt[s]] for s in ranks], ranks, rank)
        c = torch.fft.ifft(torch.cat(got, dim=0), dim=0)
        out = torch.fft.irfft(c, n=shape[2], dim=2) if dim == 3 else c.real
        err_b = float((out - loc).abs().max())
        q.put((rank, err_f, err_b))
    finally:
        dist.barrier()
        dist.destroy_process_group()


def _pencil_worker(rank, world, port, shape, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from marlin_b200 import capi
    try:
        nx, ny, nz = shape
        nxc = nx // 2 + 1
        py, pz = C.c_int(), C.c_int()
        capi._ck(capi.lib().mrl_pencil_factors(world, (C.c_int64 * 3)(*shape), C.byref(py), C.byref(pz)))
        Py, Pz = py.value, pz.value
        a_, b_ = rank % Py, rank // Py
        ycnt, ybeg = _partition(ny, Py)
        zcnt, zbeg = _partition(nz, Pz)
        xcnt, xbeg = _partition(nxc, Py)
        y2cnt, y2beg = _partition(ny, Pz)
        torch.manual_seed(3)
        full = torch.rand(shape, dtype=torch.float64)
        ref = torch.fft.fftn(torch.fft.rfft(full, dim=0), dim=(1, 2))
        loc = full[:, ybeg[a_]:ybeg[a_] + ycnt[a_], zbeg[b_]:zbeg[b_] + zcnt[b_]]
        zgroup = [b_ * Py + p for p in range(Py)]            # same z part: exchange x <-> y
        xgroup = [q_ * Py + a_ for q_ in range(Pz)]          # same kx part: exchange y <-> z
        s1 = torch.fft.rfft(loc, dim=0)
        got = _all_to_all([s1[xbeg[p]:xbeg[p] + xcnt[p]] for p in range(Py)], zgroup, rank)
        s2 = torch.fft.fft(torch.cat(got, dim=1), dim=1)
        got = _all_to_all([s2[:, y2beg[q_]:y2beg[q_] + y2cnt[q_]] for q_ in range(Pz)], xgroup, rank)
        spec = torch.fft.fft(torch.cat(got, dim=2), dim=2)
        want = ref[xbeg[a_]:xbeg[a_] + xcnt[a_], y2beg[b_]:y2beg[b_] + y2cnt[b_]]
        err_f = float((spec - want).abs().max() / ref.abs().max())
        # inverse: the stages in reverse; the real transform along x completes the half spectrum locally
        t2 = torch.fft.ifft(spec, dim=2)
        got = _all_to_all([t2[:, :, zbeg[q_]:zbeg[q_] + zcnt[q_]] for q_ in range(Pz)], xgroup, rank)
        t1 = torch.fft.ifft(torch.cat(got, dim=1), dim=1)
        got = _all_to_all([t1[:, ybeg[p]:ybeg[p] + ycnt[p]] for p in range(Py)], zgroup, rank)
        out = torch.fft.irfft(torch.cat(got, dim=0), n=nx, dim=0)
        err_b = float((out - loc).abs().max())
        q.put((rank, err_f, err_b))
    finally:
        dist.barrier()
        dist.destroy_process_group()


def _run(target, world, args):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29200 + (os.getpid() * 3 + world + len(str(args))) % 150
    procs = [ctx.Process(target=target, args=(r, world, port) + args + (q,)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ef, eb in res:
        assert ef < 1e-13 and eb < 1e-13, (rank, ef, eb)


@pytest.mark.parametrize("world,shape,weights", [(2, (20, 20), None), (3, (9, 7), None), (3, (10, 8, 6), None), (2, (16, 12, 10), [3, 1]),
                                                 (4, (9, 7, 6), None)])
def test_slab_rules(world, shape, weights):
    _run(_slab_worker, world, (shape, weights))


@pytest.mark.parametrize("world,shape", [(4, (8, 8, 8)), (4, (9, 7, 6)), (6, (10, 9, 7))])
def test_pencil_rules(world, shape):
    _run(_pencil_worker, world, (shape,))
