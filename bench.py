#!/usr/bin/env python
"""bench.py -- GNNML3 training throughput (graphs/s) on synthetic dataset-shaped graphs, one process per GPU.

    python bench.py --gpus 1 --steps K --warmup W                      (this framework, default)
    python bench.py --impl reference --gpus N --steps K --warmup W     (the reference's CPU path: oracle port)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one optimisation step (CSR build for the new batch, forward, SUM loss, backward, gradient
all-reduce when N > 1, Adam) of the reference's GNNML3 model for the workload (default: the Zinc12k.py model on
ZINC-shaped graphs, BASELINE.json configs[1]) on one batch of `--batch` graphs per GPU.  Prints ONE JSON line.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (pool kind, model config, loss, default per-GPU batch, reference script)
    "zinc": ("zinc", "zinc", "l1", 8192, "Zinc12k.py GNNML3 (4 x ML3Layer 30||2, K=8, add-pool, L1-sum, Adam 1e-3)"),
    "counting": ("counting", "counting", "mse", 8192, "counting.py GNNML3 (5 x ML3Layer 16||16, K=12, add-pool, MSE-sum, Adam 1e-3)"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="zinc", choices=sorted(WORKLOADS) + ["spectconv_sweep"],
                    help="zinc (default, BASELINE.json configs[1]) / counting: GNNML3 training; spectconv_sweep: one SpectConv "
                         "layer fwd+bwd on a 1M-node batch (BASELINE.json configs[4]), reported as HBM GB/s")
    ap.add_argument("--sweep-nodes", type=int, default=1000000)
    ap.add_argument("--sweep-f", type=int, default=64)
    ap.add_argument("--batch", type=int, default=0, help="graphs per GPU per step (0 = workload default)")
    ap.add_argument("--pool", type=int, default=2048, help="distinct synthetic graphs in the pool")
    ap.add_argument("--ring", type=int, default=6, help="distinct resident batches cycled through (> L2 in total)")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "tf32"])
    ap.add_argument("--cpu-graphs", type=int, default=2048, help="graphs per step of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-breakdown", action="store_true")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def spmm_bytes(shapes):
    """Algorithmic bytes of one stand-alone gnnml3_spmm_k launch: 4*[N*F + E*(K+1) + (N+1) + N*K*F].
    shapes = (rowptr, col, [eperm], ea, x, [out])"""
    two_d = [s for s in shapes if len(s) == 2]
    (E, K), (N, F) = two_d[0], two_d[1]
    return 4.0 * (N * F + E * (K + 1) + (N + 1) + N * K * F)


def run_reference(args, rank, world):
    """The reference's CPU implementation of the path (oracle port of libs/spect_conv.py + the script's model
    and training step), all host threads, on a bounded sample of the same workload."""
    if rank != 0:
        return
    from oracle import gnnml3_oracle as O
    from gnn_matlang_b200.synthetic import GraphPool
    from gnn_matlang_b200.train import loss_fn
    kind, cfg, loss, _, desc = WORKLOADS[args.workload]
    torch.set_num_threads(os.cpu_count() or 1)
    pool = GraphPool(kind, min(args.pool, 512), seed=0)
    rng = np.random.default_rng(0)
    B = args.cpu_graphs
    torch.manual_seed(0)
    model = O.OracleGNNML3(cfg, pool.K, pool.F)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    batches = [pool.draw(rng, B) for _ in range(2)]

    def step(hb):
        b = dict(x=hb.x, edge_index2=hb.edge_index2, edge_attr2=hb.edge_attr2, batch=hb.batch, num_graphs=hb.num_graphs)
        opt.zero_grad()
        l = loss_fn(loss, model(b), hb.y)
        l.backward()
        opt.step()
        return float(l)

    for i in range(args.warmup):
        step(batches[i % 2])
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(batches[i % 2])
    dt = time.perf_counter() - t0
    v = B * args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": "GNNML3 train graphs/s", "value": v, "unit": "graphs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "reference_script": desc, "graphs_per_step": B},
        "cpu_baseline": {"value": v, "unit": "graphs/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": "%d steps of %d %s-shaped graphs, oracle port of the reference's PyG path on host CPU" % (args.steps, B, args.workload)},
        "e2e": {"value": v, "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def run_sweep(args, local):
    """Second headline metric of BASELINE.json: SpectConv fwd+bwd achieved HBM GB/s on a 1M-node batch of 30-100-node
    graphs (K = 10 supports, F -> F).  achieved = SURVEY.md 8d algorithmic bytes (fwd + bwd, every operand once, the
    [N, K*F] aggregate never credited) / device time of one forward + backward."""
    from gnn_matlang_b200 import _lib
    from gnn_matlang_b200.graph import get_plan, set_range_check
    from gnn_matlang_b200.libs.spect_conv import SpectConv
    from gnn_matlang_b200.synthetic import GraphPool
    _lib.load()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    set_range_check(False)
    F, K = args.sweep_f, 10
    pool = GraphPool("sweep", 2048, seed=1, K=K, nfeat=F)
    rng = np.random.default_rng(3)
    B = int(args.sweep_nodes / float(pool.n.mean()))
    hb = pool.draw(rng, B)
    N, E = hb.x.shape[0], hb.edge_index2.shape[1]
    torch.manual_seed(0)
    layer = SpectConv(F, F, K, selfconn=False).to(dev)
    ei = hb.edge_index2.to(dev)
    ea = hb.edge_attr2.to(dev).requires_grad_(True)
    xs = [torch.randn(N, F, device=dev).requires_grad_(True) for _ in range(3)]       # 3 x 256 MB inputs > L2
    gout = torch.randn(N, F, device=dev)
    get_plan(ei, N)

    def step(i):
        x = xs[i % len(xs)]
        out = layer(x, ei, ea)
        out.backward(gout)
        x.grad = None
        ea.grad = None
        layer.zero_grad(set_to_none=True)

    for i in range(max(args.warmup, 3)):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    from gnn_matlang_b200 import ops
    ops.profile_start()
    for i in range(2):
        step(i)
    agg = {}
    for n_, m_, _ in ops.profile_stop():
        agg[n_] = agg.get(n_, 0.0) + m_ / 2
    breakdown = {k: round(v, 4) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])}
    fwd = 4.0 * (N * F + E * K + E + (N + 1) + K * F * F + F + N * F)
    bwd = 4.0 * (N * F + N * F + E * K + 2 * E + (N + 1) + N * F + E * K + 2 * K * F * F + F)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    ach = (fwd + bwd) / (ms * 1e-3) / 1e9
    print(json.dumps({
        "metric": "SpectConv fwd+bwd HBM GB/s", "value": ach, "unit": "GB/s", "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "spectconv_sweep", "nodes": int(N), "support_entries": int(E), "graphs": int(B), "K": K, "F_in": F,
                   "F_out": F, "mask": "1-hop", "l2_policy": "3 distinct %d MB inputs cycled" % (N * F * 4 // 1000000)},
        "roofline": {"kernel": "SpectConv layer fwd+bwd (all launches)", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                     "frac": ach / peak, "traffic": None, "algorithmic_bytes_fwd": fwd, "algorithmic_bytes_bwd": bwd},
        "graphs_per_s": B / (ms * 1e-3), "kernel_ms_per_step": breakdown}), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload == "spectconv_sweep":
        if rank == 0:
            run_sweep(args, local)
        return

    import torch.distributed as dist
    from gnn_matlang_b200 import _lib, ops
    from gnn_matlang_b200.graph import set_range_check
    from gnn_matlang_b200.models import GNNML3
    from gnn_matlang_b200.synthetic import GraphPool
    from gnn_matlang_b200.train import HostFeeder, Trainer

    _lib.load()                      # fails loudly if the CUDA library is missing
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"      # the NCCL version banner goes to stdout; stdout carries ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    kind, cfg, loss, defB, desc = WORKLOADS[args.workload]
    B = args.batch or defB
    set_range_check(False)

    # ---- data: pool of synthetic graphs -> ring of distinct batches (pinned host + HBM-resident copies)
    pool = GraphPool(kind, args.pool, seed=1000 + rank)
    rng = np.random.default_rng(7 + rank)
    host_ring = [pool.draw(rng, B).pin_memory() for _ in range(args.ring)]
    dev_ring = [hb.to(dev, non_blocking=False) for hb in host_ring]
    batch_bytes = host_ring[0].nbytes()
    N0, E0 = host_ring[0].x.shape[0], host_ring[0].edge_index2.shape[1]

    torch.manual_seed(0)             # identical replicas
    model = GNNML3(cfg, pool.K, pool.F, precision=args.precision).to(dev)
    trainer = Trainer(model, loss=loss, lr=1e-3, distributed=world > 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up
    for i in range(max(args.warmup, 3)):
        trainer.step(dev_ring[i % args.ring].fresh())
    barrier()

    # ---- timed region: inputs resident in HBM; successive steps use different batches (ring > L2)
    sampler = ClockSampler(local)
    sampler.start()
    lib = _lib.load()
    lib.gnnml3_fused_profile(1)          # CUDA events around every fused layer-kernel launch, on the launching stream
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    th0 = time.perf_counter()
    for i in range(args.steps):
        loss_t = trainer.step(dev_ring[i % args.ring].fresh())
    host_ms = (time.perf_counter() - th0) * 1e3 / args.steps      # CPU time to ENQUEUE a step (no sync inside the loop)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - n0
    import ctypes
    pbuf = (ctypes.c_double * (10 * 4096))()
    nrec = lib.gnnml3_fused_profile_fetch(pbuf, 4096)
    lib.gnnml3_fused_profile(0)
    per_step = nrec // max(args.steps, 1)
    recs = []
    for j in range(nrec):
        msj, Nn, Kk, Ff, Ncc, Fss, smode, Nss, Gg, hp = [pbuf[j * 10 + t] for t in range(10)]
        Ee = float(host_ring[(j // max(per_step, 1)) % args.ring].edge_index2.shape[1])
        # algorithmic HBM bytes: every operand and result once, the [N, K*F] aggregate never credited
        nbytes = 4.0 * (Nn * Ff + (Nn * Fss if smode == 2 else 0) + Ee * Kk + Ee * (2 if hp else 1) + (Nn + 1) + Kk * Ff * Ncc
                        + Fss * (Nss if smode == 1 else (Ncc if smode == 2 else 0)) + Ncc + Nn * (Ncc + (Gg if smode == 1 else 0))
                        + (Nn * 2 * Gg if smode == 1 else 0))
        recs.append(("fused_agg_proj", msj, ("algorithmic_bytes", nbytes)))
    sampler.stop_flag = True
    sampler.join()
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (the fused aggregate+project layer kernel: forward launches and the transposed
    #      dx launches of the backward), from the per-launch CUDA events of the timed region.  achieved = algorithmic bytes
    #      (every operand and result once; the [N, K*F] aggregate is never credited) / launch time.
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    recs = [r for r in recs if r[0] == "fused_agg_proj"]
    sp_ms = sum(r[1] for r in recs)
    sp_bytes = sum(r[2][1] for r in recs)
    roofline = None
    if recs and sp_ms > 0:
        ach = sp_bytes / (sp_ms * 1e-3) / 1e9
        traffic = None
        try:      # dram bytes per launch from the committed ncu --set full capture of the same kernel (profiles/)
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_fused_kernel_ncu_summary.json")))["dram_bytes_per_launch"]
        except Exception:
            pass
        roofline = {"kernel": "k_fused_agg_proj (gnnml3_fused_agg_proj: SpectConv aggregate + projection + gates, fwd and dx launches)",
                    "bound": "hbm", "achieved": ach, "peak": peak,
                    "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650",
                    "launches": len(recs), "avg_launch_ms": sp_ms / len(recs), "share_of_step": sp_ms / ms,
                    "algorithmic_bytes_per_launch": sp_bytes / len(recs),
                    "note": "latency/issue-bound gather + fixed-cost tcgen05.mma issue, see DESIGN.md section 4"}

    # ---- kernel breakdown pass (separate, untimed): share of every library call
    breakdown = None
    if not args.no_breakdown:
        # per-kernel view: compose the layers from the single-kernel entry points for this (untimed) pass -- the timed
        # region above goes through the whole-layer entry points, which the Python-side event wrappers cannot look into
        from gnn_matlang_b200.libs import spect_conv as _sc
        _sc.USE_LAYER_API = False
        ops.profile_start()
        for i in range(2):
            trainer.step(dev_ring[i % args.ring].fresh())
        recs_all = ops.profile_stop()
        _sc.USE_LAYER_API = True
        agg = {}
        for n, m, _ in recs_all:
            agg[n] = agg.get(n, 0.0) + m / 2
        breakdown = {k: round(v, 4) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])}

    # ---- end to end: host (pinned) batches through Trainer; H2D of every step's inputs + D2H of the loss inside
    e2e = None
    if not args.no_e2e:
        feeder = HostFeeder(dev)
        feeder.prefetch(host_ring[0])
        for i in range(3):
            b = feeder.get()
            feeder.prefetch(host_ring[(i + 1) % args.ring])
            float(trainer.step(b).item())
        barrier()
        t0 = time.perf_counter()
        e0.record()
        # every step's loss is read back to the host: the device -> host copy is enqueued right behind the step (pinned
        # buffer, non-blocking) and harvested while the next step is being enqueued, so the read never stalls the pipeline
        loss_host = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
        loss_ev = [torch.cuda.Event(), torch.cuda.Event()]
        lv = float("nan")
        for i in range(args.steps):
            b = feeder.get()
            feeder.prefetch(host_ring[(i + 4) % args.ring])
            lt = trainer.step(b)
            loss_host[i & 1].copy_(lt.reshape(1), non_blocking=True)      # device -> host read of the step's loss
            loss_ev[i & 1].record()
            if i > 0:
                loss_ev[(i - 1) & 1].synchronize()
                lv = float(loss_host[(i - 1) & 1][0])
        loss_ev[(args.steps - 1) & 1].synchronize()
        lv = float(loss_host[(args.steps - 1) & 1][0])
        e1.record()
        barrier()
        ms2 = e0.elapsed_time(e1)
        t2 = torch.tensor([ms2], device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        ms2 = float(t2.item())
        # this box's pinned host -> device rate for one batch (explains e2e when the link, not the GPU, is the bound:
        # a step needs h2d_bytes_per_step / ms_per_step of it)
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for i in range(3):
            host_ring[i % args.ring].to(dev, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        h2d_gbs = 3 * batch_bytes / (c0.elapsed_time(c1) * 1e-3) / 1e9
        e2e = {"value": world * B * args.steps / (ms2 * 1e-3), "unit": "graphs/s", "h2d_bytes_per_step": int(batch_bytes),
               "d2h_bytes_per_step": 4, "ms_per_step": ms2 / args.steps, "last_loss": lv,
               "h2d_link_gbs_measured": round(h2d_gbs, 2)}

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only): oracle port, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import gnnml3_oracle as O
        from gnn_matlang_b200.train import loss_fn
        torch.set_num_threads(os.cpu_count() or 1)
        Bc = min(args.cpu_graphs, B)
        hb = pool.collate(np.arange(Bc) % len(pool.n))
        ob = dict(x=hb.x, edge_index2=hb.edge_index2, edge_attr2=hb.edge_attr2, batch=hb.batch, num_graphs=hb.num_graphs)
        cm = O.OracleGNNML3(cfg, pool.K, pool.F)
        copt = torch.optim.Adam(cm.parameters(), lr=1e-3)

        def cstep():
            copt.zero_grad()
            l = loss_fn(loss, cm(ob), hb.y)
            l.backward()
            copt.step()

        cstep()
        t0 = time.perf_counter()
        nrep = 0
        while nrep < 3 or time.perf_counter() - t0 < 8.0:
            cstep()
            nrep += 1
        dt = time.perf_counter() - t0
        cpu = {"value": Bc * nrep / dt, "unit": "graphs/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": "%d training steps of %d %s-shaped graphs (oracle port of the reference's PyG path, torch CPU)" % (nrep, Bc, args.workload)}

    if rank == 0:
        print(json.dumps({
            "metric": "GNNML3 train graphs/s", "value": value, "unit": "graphs/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "tf32", "data": "synthetic",
            "config": {"workload": args.workload, "reference_script": desc, "graphs_per_gpu_per_step": B, "graphs_per_step_all_gpus": B * world,
                       "nodes_per_step_per_gpu": int(N0), "support_entries_per_step_per_gpu": int(E0), "K": pool.K,
                       "edge_attr": pool.supports, "gemm_arithmetic": "3xTF32 (FP32-grade)" if args.precision == "fp32" else "TF32",
                       "sharding": "whole graphs over %d rank(s), one gradient SUM all-reduce per step" % world, "l2_policy": "ring of %d distinct resident batches, %.0f MB in total (> 126 MB L2)" % (args.ring, args.ring * batch_bytes / 1e6),
                       "timed_step": "csr_build + fwd + loss + bwd + (allreduce) + Adam"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": sampler.summary(), "host_enqueue_ms_per_step": host_ms, "kernel_ms_per_step": breakdown, "final_loss": float(loss_t.item())}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
