import os, sys, traceback
LOG = os.open("gpurun_out/debug1.log", os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
def log(*a):
    os.write(LOG, (" ".join(str(x) for x in a) + "\n").encode()); os.fsync(LOG)
try:
    log("start", sys.executable)
    import torch
    log("torch", torch.__version__, torch.cuda.is_available())
    sys.path.insert(0, "."); sys.path.insert(0, "tests")
    from marlin_b200 import capi
    log("lib", capi.lib().mrl_version())
    ctx = capi.Context(0, capi.F64)
    log("ctx ok")
    for shape in [(16,), (8, 9), (16, 16), (64, 64), (16, 16, 16), (20, 20, 20)]:
        ctx.domain_set(len(shape), shape)
        log("domain", shape, ctx.rshape)
        a = torch.rand(shape, dtype=torch.float64)
        g = a.cuda()
        log("  calling rfftn")
        out = ctx.rfftn(g)
        log("  launched")
        torch.cuda.synchronize()
        log("  synced")
        ref = torch.fft.rfftn(a, dim=list(range(len(shape))))
        err = (out.cpu() - ref).abs().max().item()
        log("  err", err)
        back = ctx.irfftn(out); torch.cuda.synchronize()
        log("  roundtrip", (back.cpu() - a).abs().max().item())
    log("done")
except BaseException as e:
    log("EXC", repr(e)); log(traceback.format_exc())
    raise
