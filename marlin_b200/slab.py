"""Multi-GPU slab decomposition of the fused Cahn-Hilliard substep.

Mirrors DomainAction::partitionSlabs / fftSlab / ifftSlab (src/actions/DomainAction.C:511-566,
:870-1019): real space split along y, reciprocal space along x, z never split, one process
per GPU.  The three compute phases are C-ABI calls (mrl_slab_forward / _update / _inverse);
the exchange between them is an all-to-all whose chunks are contiguous on both sides, issued
here through torch.distributed (NCCL over NVLink on the GPU box; any backend for the layout
tests).  The half spectrum travels, not the reference's full c2c spectrum.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import capi
from .capi import AB_BETA, _ck, _p, lib


def exchange_forward(recv, send, group=None):
    """recv[s] <- chunk this rank's peer s cut out for us (send[s] goes to rank s).
    `send`/`recv`: real views [P, chunk] of the spectra."""
    dist.all_to_all_single(recv, send, group=group)


def exchange_backward(recv, send, group=None):
    dist.all_to_all_single(recv, send, group=group)


class SlabContext(capi.Context):
    def domain_set_slab(self, n, mins, maxs, rank, nranks):
        n3 = (C.c_int64 * 3)(*[int(v) for v in n])
        mn = (C.c_double * 3)(*[float(v) for v in mins])
        mx = (C.c_double * 3)(*[float(v) for v in maxs])
        _ck(lib().mrl_domain_set_slab(self.h, 3, n3, mn, mx, int(rank), int(nranks)))
        self.dim = 3
        rs, rb, ks, kb = [(C.c_int64 * 3)() for _ in range(4)]
        _ck(lib().mrl_domain_local(self.h, rs, rb, ks, kb))
        self.shape, self.rbegin = list(rs), list(rb)      # local real shape / first global index
        self.rshape, self.kbegin = list(ks), list(kb)     # local reciprocal shape / first global index
        self.rank, self.nranks = rank, nranks
        f, c, pitch = C.c_int64(), C.c_int64(), C.c_int()
        _ck(lib().mrl_slab_sizes(self.h, C.byref(f), C.byref(c), C.byref(pitch)))
        self.field_elems, self.chunk_elems, self.pitch = f.value, c.value, pitch.value


class SlabPlan:
    """Fused semi-implicit substep on one slab (the per-rank part of AdamsBashforthMoulton::substep
    with the FFTs of DomainAction::fftSlab / ifftSlab)."""

    def __init__(self, ctx, double_well, M_factor, L_factor=None, history=1, group=None, mode="peer"):
        """mode "peer": the all-to-all is fused into the passes (bulk cp.async stores from shared memory into
        NVLink-mapped peer memory, CUDA IPC; blocked staging layouts, marlin_b200/csrc/mrl_passes_slab.cuh).
        MRL_SLAB_SYNC=flags replaces the two barriers of a substep by per-column-block arrival counters (the
        barrier calls below become no-ops), MRL_SLAB_INV_CTAS=n lets the x inverse pass trail the fused y pass
        on n SMs.  mode "nccl": the three phases with torch.distributed all-to-all calls between them (the
        baseline the fused mode is measured against)."""
        self.ctx, self.group, self.mode = ctx, group, mode
        # barrier between the phases in peer mode: "dev" = flags in peer memory (mrl_slab_barrier),
        # "nccl" = a 1-element all-reduce
        import os
        self.barrier_kind = os.environ.get("MRL_SLAB_BARRIER", "dev")
        self.sync = os.environ.get("MRL_SLAB_SYNC", "barrier")
        dev = ctx.device
        f = ctx.field_elems
        d = capi.SplitDesc()
        d.nonlin_kind = capi.NONLIN_DOUBLE_WELL
        A, a, b = double_well
        d.nonlin_params = (C.c_double * 4)(A, a, b, 0.0)
        d.M_closed_form, d.M_factor = 1, float(M_factor)
        d.has_L, d.L_closed_form, d.L_factor = int(L_factor is not None), 1, float(L_factor or 0.0)
        d.history = history
        self.h = C.c_void_p()
        if mode == "peer":
            _ck(lib().mrl_slab_plan_create_peer(ctx.h, C.byref(d), C.byref(self.h)))
            mine = (C.c_ubyte * 128)()
            _ck(lib().mrl_slab_ipc_export(self.h, mine))
            t = torch.tensor(list(mine), dtype=torch.uint8, device=dev)
            alls = [torch.empty_like(t) for _ in range(ctx.nranks)]
            dist.all_gather(alls, t, group=group)
            blob = bytes(torch.cat(alls).cpu().tolist())
            _ck(lib().mrl_slab_ipc_import(self.h, blob))
            self._flag = torch.zeros(1, dtype=torch.float32, device=dev)
            dist.all_reduce(self._flag, group=group)  # every rank has mapped every buffer
            torch.cuda.synchronize()
            return
        self.send_fwd = torch.zeros(2 * f, dtype=ctx.cdtype, device=dev)
        self.recv_fwd = torch.zeros(2 * f, dtype=ctx.cdtype, device=dev)
        self.send_bwd = torch.zeros(f, dtype=ctx.cdtype, device=dev)
        _ck(lib().mrl_slab_plan_create(ctx.h, C.byref(d), _p(self.send_fwd), _p(self.recv_fwd), _p(self.send_bwd),
                                       C.byref(self.h)))
        P, ch = ctx.nranks, ctx.chunk_elems
        r = torch.view_as_real
        # real views [P, 2*chunk] used by the exchanges
        self._sf = [r(self.send_fwd[i * f:(i + 1) * f]).view(P, 2 * ch) for i in range(2)]
        self._rf = [r(self.recv_fwd[i * f:(i + 1) * f]).view(P, 2 * ch) for i in range(2)]
        self._sb = r(self.send_bwd).view(P, 2 * ch)

    def substep(self, c, dt, beta, nold):
        b = (C.c_double * 5)(*(list(beta) + [0.0] * 5)[:5])
        if self.mode == "peer":
            # stream-ordered cross-rank barriers (a 1-element all-reduce) separate the phases: a
            # rank's pass may only read what every peer's previous pass has finished storing
            _ck(lib().mrl_slab_forward(self.h, _p(c)))
            self._barrier()
            _ck(lib().mrl_slab_update(self.h, C.c_double(dt), b, int(nold)))
            self._barrier()
            _ck(lib().mrl_slab_inverse(self.h, _p(c)))
            return
        _ck(lib().mrl_slab_forward(self.h, _p(c)))
        for i in (1, 0):
            exchange_forward(self._rf[i], self._sf[i], self.group)
        _ck(lib().mrl_slab_update(self.h, C.c_double(dt), b, int(nold)))
        exchange_backward(self._sf[0], self._sb, self.group)
        _ck(lib().mrl_slab_inverse(self.h, _p(c)))

    def _barrier(self):
        if self.barrier_kind == "dev":
            _ck(lib().mrl_slab_barrier(self.h))
        else:
            dist.all_reduce(self._flag, group=self.group)

    def substep_timed(self, c, dt, beta, nold):
        """substep() with CUDA events between the phases (on the launching stream); returns the five
        intervals in ms: forward passes, exchange/barrier, fused update pass, exchange/barrier,
        inverse passes."""
        b = (C.c_double * 5)(*(list(beta) + [0.0] * 5)[:5])
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        ev[0].record()
        _ck(lib().mrl_slab_forward(self.h, _p(c)))
        ev[1].record()
        if self.mode == "peer":
            self._barrier()
        else:
            for i in (1, 0):
                exchange_forward(self._rf[i], self._sf[i], self.group)
        ev[2].record()
        _ck(lib().mrl_slab_update(self.h, C.c_double(dt), b, int(nold)))
        ev[3].record()
        if self.mode == "peer":
            self._barrier()
        else:
            exchange_backward(self._sf[0], self._sb, self.group)
        ev[4].record()
        _ck(lib().mrl_slab_inverse(self.h, _p(c)))
        ev[5].record()
        torch.cuda.synchronize()
        return [ev[i].elapsed_time(ev[i + 1]) for i in range(5)]

    def advance_state(self):
        v = C.c_int()
        _ck(lib().mrl_slab_advance_state(self.h, C.byref(v)))
        return v.value

    def close(self):
        if self.h:
            lib().mrl_slab_plan_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def bind_host_side(local, world):
    """Pin this rank's host threads (and, by first touch, the pinned buffers it allocates afterwards) to cores of the NUMA
    node its GPU hangs off, a distinct core range per rank: eight ranks sharing one node's cores and memory is what kept
    the end-to-end figure from scaling.  Returns a description for the bench line."""
    import os
    info = {"numa_node": None, "cpus": None}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        node = -1
        for cand in (bus.lower(), bus[4:].lower()):
            path = f"/sys/bus/pci/devices/{cand}/numa_node"
            if os.path.exists(path):
                node = int(open(path).read().strip())
                break
        nodes = sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
        if node < 0:
            node = nodes[local * len(nodes) // max(world, 1)] if len(nodes) > 1 else (nodes[0] if nodes else 0)
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0)) or sorted(os.sched_getaffinity(0))
        # the ranks whose GPUs share this node split its cores
        per = max(1, len(allowed) // max(1, world))
        mine = allowed[(local * per) % len(allowed):][:per] or allowed
        os.sched_setaffinity(0, mine)
        info = {"numa_node": node, "cpus": f"{mine[0]}-{mine[-1]}", "nodes": len(nodes)}
    except Exception as exn:  # binding is an optimisation: never fail the run over it
        info["error"] = str(exn)[:120]
    return info


def bench(args, rank, world, metric):
    """bench.py leg for N > 1: CH-3D-n slab-decomposed over `world` GPUs (strong scaling)."""
    import json
    import math
    import os
    import sys
    import time

    from bench import ClockSampler, algorithmic_bytes, workload  # noqa: E402 (bench.py is on sys.path)

    local = int(os.environ.get("LOCAL_RANK", rank))
    host_binding = bind_host_side(local, world)   # before any pinned allocation
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.n
    L = n * 8 * math.pi / 200
    ctx = SlabContext(local, capi.F64)
    ctx.use_torch_stream()
    ctx.domain_set_slab((n, n, n), (0,) * 3, (L,) * 3, rank, world)
    # the same global initial condition as the 1-GPU bench; every rank takes its y-slab
    torch.manual_seed(0)
    nyl, y0 = ctx.shape[1], ctx.rbegin[1]
    full = torch.rand((n, n, n), dtype=torch.float64) * 0.12 + 0.44
    host_c = full[:, y0:y0 + nyl, :].contiguous().pin_memory()
    del full
    c = host_c.cuda()
    mode = os.environ.get("MRL_SLAB_MODE", "peer")
    plan = SlabPlan(ctx, (0.1, 0.0, 1.0), 0.2, -0.001, history=1, mode=mode)
    dt = 1e-3
    plan.substep(c, dt, AB_BETA[0], 0)
    plan.advance_state()

    # parity inside this very run: every rank repeats the first 6 substeps (1 x AB1, 5 x AB2) of the GLOBAL field with the
    # single-GPU fused plan on its own device (the plan tests/test_gpu_parity.py::test_ch3d_512_matches_oracle and the
    # 1-GPU bench line's `parity` object pin to the oracle at this size) and compares its y-slab; max over ranks
    par_sub = 6
    for _ in range(par_sub - 1):
        plan.substep(c, dt, AB_BETA[1], 1)
        plan.advance_state()
    parity = None
    try:
        ctx1 = capi.Context(local, capi.F64)
        ctx1.use_torch_stream()
        ctx1.domain_set(3, (n, n, n), (0,) * 3, (L,) * 3)
        torch.manual_seed(0)
        cfull = (torch.rand((n, n, n), dtype=torch.float64) * 0.12 + 0.44).cuda()
        p1 = ctx1.split_plan(double_well=(0.1, 0.0, 1.0), M_factor=0.2, L_factor=-0.001, history=1)
        p1.substep(cfull, dt, AB_BETA[0], 0)
        p1.advance_state()
        for _ in range(par_sub - 1):
            p1.substep(cfull, dt, AB_BETA[1], 1)
            p1.advance_state()
        ref = cfull[:, y0:y0 + nyl, :]
        num = (c - ref).double().pow(2).sum().reshape(1)
        den = ref.double().pow(2).sum().reshape(1)
        mx = (c - ref).abs().max().reshape(1)
        dist.all_reduce(num)
        dist.all_reduce(den)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        rel = float((num / den).sqrt().item())
        parity = {"status": "green" if rel <= 1e-10 else "RED", "rel_l2_c_vs_single_gpu_plan": rel, "max_abs": float(mx.item()),
                  "tolerance": 1e-10, "substeps": par_sub, "ranks": world,
                  "what": "slab-decomposed substeps on all ranks vs the single-GPU fused plan (itself pinned to the oracle at "
                          "this size) on the same global seed-0 initial condition; L2 over the whole grid, max over ranks"}
        p1.close()
        ctx1.close()
        del cfull, ref, p1, ctx1
        torch.cuda.empty_cache()
    except Exception as exn:  # e.g. a grid that does not fit one GPU (1024^3 does: 27 GB)
        parity = {"status": "not run", "why": str(exn)[:200]}

    def step():
        plan.substep(c, dt, AB_BETA[1], 1)
        plan.advance_state()

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    launches = ctx.launch_count() - l0

    # per-phase device times (CUDA events between the phases), max over ranks
    reps, acc = 10, None
    for _ in range(reps):
        dist.barrier()
        tph = plan.substep_timed(c, dt, AB_BETA[1], 1)
        plan.advance_state()
        acc = tph if acc is None else [a + b for a, b in zip(acc, tph)]
    tph = torch.tensor([a / reps for a in acc], dtype=torch.float64, device="cuda")
    dist.all_reduce(tph, op=dist.ReduceOp.MAX)
    phases_ms = [round(float(v), 4) for v in tph.tolist()]

    # end to end: H2D of the local slab, substep, D2H of the local slab, every step, through the staged
    # transfer entry points on alternating buffers (uploads / downloads of neighbouring steps overlap)
    nbytes = host_c.numel() * 8
    e2e_steps = max(4, min(args.steps, 10))
    cbuf = [c, torch.empty_like(c)]
    out_host = [torch.empty_like(host_c).pin_memory() for _ in range(2)]

    def e2e_run(nsteps):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(nsteps):
            b = k & 1
            ctx.upload_staged(cbuf[b], host_c)
            plan.substep(cbuf[b], dt, AB_BETA[1], 1)
            plan.advance_state()
            ctx.download_staged(out_host[b], cbuf[b])
        ctx.staged_wait()
        ctx.synchronize()
        return (time.perf_counter() - t0) * 1e3

    e2e_run(2)
    t = torch.tensor([e2e_run(e2e_steps)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item()) / e2e_steps
    clocks = sampler.stop() if sampler else None

    # the same workload the way a reference input reaches it: every rank runs the host driver on the verbatim
    # examples/cahn_hilliard/cahnhilliard2.i with [Domain] parallel_mode = FFT_SLAB (the ranks of the driver find each
    # other through the torchrun environment, host/shim/comm.h); its AdamsBashforthMoulton builds the fused slab plan with
    # the file's ParsedCompute nonlinearity compiled into the first pass.  Device-synchronised solve time of the last
    # step per substep, max over ranks.
    host_driver = None
    if not getattr(args, "fast", False) and os.environ.get("MRL_BENCH_HOST_DRIVER", "1") != "0":
        torch.cuda.synchronize()
        dist.barrier()          # no peer is still pushing rows into this rank's staging buffers
        plan.close()
        del c, cbuf
        torch.cuda.empty_cache()
        try:
            from bench import host_driver_bench
            hd = host_driver_bench(n, substeps=50, steps=3, extra_args=["Domain/parallel_mode=FFT_SLAB"], timeout=150)
            t = torch.tensor([hd["ms_per_substep_last_step"]], dtype=torch.float64, device="cuda")
            ok = torch.tensor([1.0 if hd["fused_plan"] else 0.0], dtype=torch.float64, device="cuda")
        except Exception as exn:
            hd = {"error": str(exn)[-300:]}
            t = torch.tensor([float("nan")], dtype=torch.float64, device="cuda")
            ok = torch.tensor([0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if "error" not in hd:
            hd["ms_per_substep_last_step"] = round(float(t.item()), 4)
            hd["fused_plan"] = bool(ok.item())
            hd["command"] = f"{world} x (" + hd["command"] + " Domain/parallel_mode=FFT_SLAB), one process per GPU"
            hd["ratio_to_harness"] = round(float(t.item()) / (total_ms / args.steps), 4)
        host_driver = hd

    if rank == 0:
        ms = total_ms / args.steps
        s_r, s_c, _ = algorithmic_bytes(n, 1)
        b_alg = 2 * s_r + 14 * s_c
        a2a = 3 * (world - 1) / world ** 2 * s_c
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        line = {
            "metric": metric, "value": 1e3 / ms, "unit": "substeps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload(n),
                       "impl_detail": f"slab-decomposed (y real / x reciprocal) over {world} GPUs, half spectrum on the wire, "
                                      + (("exchanges by the copy engines (peer-to-peer copies over NVLink, y-chunk by y-chunk behind the passes), "
                                          if os.environ.get("MRL_SLAB_EXCHANGE", "store") == "copy" else
                                          "all-to-all fused into the passes (bulk stores from shared memory into peer HBM over NVLink), ")
                                         + (f"per-column-block arrival counters, inverse x pass on {os.environ.get('MRL_SLAB_INV_CTAS', '0')} SMs beside the fused y pass"
                                            if plan.sync == "flags" else f"{plan.barrier_kind} barrier between the phases") if mode == "peer"
                                         else "NCCL all-to-all between the phases"),
                       "l2": "inputs larger than L2" if s_r / world > 126e6 else "per-GPU slab comparable to L2",
                       "parallelism": f"slab{world}"},
            "clocks": clocks,
            "e2e": {"value": 1e3 / e2e_ms, "unit": "substeps/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": nbytes * world, "d2h_bytes_per_step": nbytes * world,
                    "host_binding_rank0": host_binding},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": round(b_alg / world / 1e9 / (ms / 1e3), 1), "peak": peak,
                         "unit": "GB/s", "frac": round(b_alg / world / 1e9 / (ms / 1e3) / peak, 4), "traffic": None,
                         "note": "per-GPU algorithmic HBM bytes / step time; the step also moves "
                                 f"{a2a / 1e9:.3f} GB per GPU over NVLink ({a2a / 1e9 / (ms / 1e3):.0f} GB/s achieved "
                                 "if it were the only cost; 770 GB/s per direction measured peer copy)"},
            "phases_ms": dict(zip(["forward (z r2c, x fwd + peer stores)", "barrier 1", "fused y pass (+ peer stores)",
                                   "barrier 2", "inverse (x inv, z c2r)"], phases_ms)),
            "forward_chunks": os.environ.get("MRL_SLAB_CHUNKS", "4 (default)"),
            "cpu_baseline": None,
            "parity": parity,
            "host_driver": host_driver,
        }
        print(json.dumps(line), flush=True)
    plan.close()
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()
