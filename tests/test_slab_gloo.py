"""world_size-2/4 gloo tests (CPU) of the slab exchange contract used by the multi-GPU path.

The CUDA kernels cannot run here, so the local passes are stood in by torch.fft on the CPU
(test code, not product code); what is under test is marlin_b200.slab's exchange layout:
  send buffer  [nx][ny/P][pitch]           -> chunk s = x-block of rank s (contiguous)
  recv buffer  [P][nx/P][ny/P][pitch]      -> chunk s = what rank s sent
  and back, landing as [nx][ny/P][pitch].
Composing local z,x transforms + exchange + y transform on the staged layout must equal the
x-slice of the serial rfftn (the reference asserts the same for fftSlab: parallel == serial,
test/tests/gradient/tests:11-30)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, n, pitch, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from marlin_b200 import slab
    try:
        nx, ny, nz = n
        nzc = nz // 2 + 1
        nyl, nxl = ny // world, nx // world
        torch.manual_seed(1)
        full = torch.rand(n, dtype=torch.float64)
        ref = torch.fft.rfftn(full)                                   # serial result
        loc = full[:, rank * nyl:(rank + 1) * nyl, :]                 # this rank's real slab
        # phase 1 stand-in: z r2c + x forward on the local slab, written with the padded pitch
        a = torch.fft.fft(torch.fft.rfft(loc, dim=2), dim=0)          # [nx][nyl][nzc]
        send = torch.zeros((nx, nyl, pitch), dtype=torch.complex128)
        send[:, :, :nzc] = a
        recv = torch.zeros((world, nxl, nyl, pitch), dtype=torch.complex128)
        chunk = nxl * nyl * pitch
        slab.exchange_forward(torch.view_as_real(recv).view(world, 2 * chunk),
                              torch.view_as_real(send).view(world, 2 * chunk))
        # phase 2 stand-in: y forward on the staged layout [P][nxl][nyl] -> y = s*nyl + yl
        staged = recv.permute(1, 0, 2, 3).reshape(nxl, ny, pitch)
        spec = torch.fft.fft(staged, dim=1)[:, :, :nzc]
        err_f = float((spec - ref[rank * nxl:(rank + 1) * nxl]).abs().max() / ref.abs().max())
        # way back: y inverse, staged layout out, exchange, x inverse, z c2r
        back = torch.zeros((nxl, ny, pitch), dtype=torch.complex128)
        back[:, :, :nzc] = torch.fft.ifft(spec, dim=1)
        sb = back.reshape(nxl, world, nyl, pitch).permute(1, 0, 2, 3).contiguous()
        ret = torch.zeros((nx, nyl, pitch), dtype=torch.complex128)
        slab.exchange_backward(torch.view_as_real(ret).view(world, 2 * chunk),
                               torch.view_as_real(sb).view(world, 2 * chunk))
        out = torch.fft.irfft(torch.fft.ifft(ret[:, :, :nzc], dim=0), n=nz, dim=2)
        err_b = float((out - loc).abs().max())
        q.put((rank, err_f, err_b))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,pitch", [(2, (8, 8, 10), 8), (2, (16, 8, 6), 4), (4, (8, 16, 12), 8)])
def test_slab_exchange_layout(world, n, pitch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, pitch, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ef, eb in res:
        assert ef < 1e-13 and eb < 1e-13, (rank, ef, eb)
