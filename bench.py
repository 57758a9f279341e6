#!/usr/bin/env python
"""bench.py -- image-text pairs/sec through the ViLBERT VQA head (36 regions x 30 tokens), BASELINE.json's metric.

    python bench.py --gpus N --steps K --warmup W                 # this repo's sm_100a engine
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference's CPU PyTorch path (oracle port)

A "step" is one forward of one batch of `--batch` pairs (default 64 = BASELINE.json configs[1]) per GPU;
batches shard over ranks with no collective -> weak scaling.  Rank 0 prints ONE JSON line.

value      : whole-job pairs/s with inputs resident in HBM (device-pointer C-ABI call, CUDA-event timed, max over ranks)
e2e        : the same through the host-buffer C-ABI call (pinned host inputs -> H2D -> forward -> D2H logits)
roofline   : tensor-pipe roofline of the dominant kernel family (the tcgen05 GEMMs, 97 % of the FLOPs)
cpu_baseline: the fp32 PyTorch oracle (port of the reference's eager forward, all heads as the reference runs
             them) timed on this box's host cores, rank 0 at N=1 only, bounded sample
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "image-text pairs/sec (VQA head, 36 regions x 30 tok)"
UNIT = "pairs/s"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=200)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--batch", type=int, default=64, help="pairs per GPU per step")
    p.add_argument("--n-tokens", type=int, default=30)
    p.add_argument("--n-regions", type=int, default=36)
    p.add_argument("--rotate", type=int, default=8, help="distinct resident input batches cycled through (L2 defeat)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-graph", action="store_true")
    p.add_argument("--pdl", choices=["default", "on", "off"], default="default",
                   help="programmatic dependent launch: engine default (every kernel), forced on, or off")
    p.add_argument("--all-heads", action="store_true", help="compute the seven task heads instead of VQA only")
    p.add_argument("--dtype", default="fp16", choices=["fp16", "bf16"],
                   help="16-bit format of the tensor-core operands (same tcgen05 rate; fp16 is the engine default, see DESIGN.md 2)")
    p.add_argument("--inflight", type=int, default=2,
                   help="batches in flight per GPU: steps alternate over this many CUDA streams / engine workspace slots")
    p.add_argument("--ops-table", default="", help="write the per-shape kernel time table (isolated graph replays) to this file")
    p.add_argument("--fused-ln", action="store_true", help="cluster-LayerNorm GEMM epilogue instead of GEMM + row LayerNorm")
    return p.parse_args()


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return {"bf16_burst": d["bf16_tflops"], "bf16_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "hbm": d["hbm_gbs"], "src": "measured"}
    except Exception:
        return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "src": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz, self.err = index, False, [], set(), None, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                     "hw_power_brake": 0x80, "sw_power_cap": 0x4, "sync_boost": 0x10}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
                time.sleep(0.003)
        except Exception as e:          # NVML missing: report it, never fail the bench
            self.err = repr(e)

    def result(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "error": self.err}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ reference arm
def oracle_model(sd, cfg_dict, num_labels):
    """The ONLY place bench.py touches oracle/: the CPU baseline legs."""
    import torch
    from oracle import vilbert_ref as R
    m = R.VILBertForVLTasks(R.RefConfig(**{k: v for k, v in cfg_dict.items() if k in R.DEFAULT_CONFIG}), num_labels=num_labels)
    m.load_state_dict(sd, strict=True)
    return m.eval()


CPU_THREADS = min(16, os.cpu_count() or 1)   # measured on the 128-vCPU GPU box: 8 -> 56, 16 -> 78, 32 -> 44, 64 -> 23, 128 -> 0.2 pairs/s


def time_oracle(model, req, steps, warmup):
    import torch
    torch.set_num_threads(CPU_THREADS)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            model(*req, output_all_attention_masks=True, compute_pretraining_heads=True)   # as the reference runs it
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return times


def run_reference(args):
    rank, _, world = env_rank()
    if rank != 0:
        return
    import torch
    import vilbert_b200 as vb
    from vilbert_b200 import synthetic as S
    cfg = vb.BertConfig(task_specific_tokens=True, visualization=True)
    sd = S.synthetic_state_dict(cfg, seed=42)
    model = oracle_model(sd, cfg.to_dict(), 3129)
    # bounded sample: ~2400 pairs in total (about two minutes of host work), never more than the real batch
    ref_batch = max(4, min(args.batch, 2400 // max(1, args.steps)))
    req = S.synthetic_request(ref_batch, args.n_tokens, args.n_regions, seed=1234)
    steps = args.steps
    times = time_oracle(model, req, steps, max(1, min(args.warmup, 2)))
    total = sum(times)
    value = ref_batch * len(times) / total
    cores = CPU_THREADS
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"batch={args.batch} VQA forward, {args.n_regions} regions x {args.n_tokens} tokens "
                                   "(BASELINE.json configs[1]); reference arm = fp32 PyTorch oracle port on host cores, "
                                   "all heads + pre-training heads as the reference executes them",
                       "global_batch": args.batch, "parallelism": "cpu"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{len(times)} forwards of {ref_batch} pairs each (bounded sample of the batch-{args.batch} "
                                       f"workload; torch fp32, {torch.get_num_threads()} threads)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import vilbert_b200 as vb
    from vilbert_b200 import synthetic as S
    from vilbert_b200 import _lib as L

    rank, local_rank, world = env_rank()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the vilbert_b200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B, Tin, V = args.batch, args.n_tokens, args.n_regions
    cfg = vb.BertConfig(task_specific_tokens=True, visualization=True)       # worker.py:509-522
    sd = S.synthetic_state_dict(cfg, seed=42)
    model = vb.VILBertForVLTasks.from_pretrained(sd, config=cfg, num_labels=3129, use_cuda_graph=not args.no_graph,
                                                 use_pdl={"default": None, "on": True, "off": False}[args.pdl], compute_dtype=args.dtype,
                                                 fused_layernorm=args.fused_ln).eval().cuda(local_rank)
    select = L.OUT_TASK_HEADS if args.all_heads else L.OUT_VIL_PREDICTION
    n_launch, flops = model.plan_info(B, Tin, V, select)

    # resident inputs: `rotate` distinct batches (8 x 19 MB > 126 MB L2 together with 466 MB of weights)
    reqs = [S.synthetic_request(B, Tin, V, seed=1234 + rank * 1000 + i) for i in range(args.rotate)]
    dreqs = [[t.to(dev) for t in r] for r in reqs]
    in_bytes = sum(t.numel() * t.element_size() for i, t in enumerate(reqs[0]) if i != 6)

    nfl = max(1, args.inflight if not args.no_graph else 1)
    streams = [torch.cuda.Stream(device=dev) for _ in range(nfl)] if nfl > 1 else [torch.cuda.current_stream(dev)]

    def step(i):
        # step i = one forward of one batch; consecutive steps go to alternating streams / workspace slots so that the
        # kernels of one batch fill SMs the other leaves idle (every step still runs start to finish inside the timed region)
        if nfl == 1:
            return model(*dreqs[i % args.rotate], select=select)
        with torch.cuda.stream(streams[i % nfl]):
            return model(*dreqs[i % args.rotate], select=select, slot=i % nfl)

    def fork():
        if nfl > 1:
            ev = torch.cuda.Event()
            ev.record()
            for s_ in streams:
                s_.wait_event(ev)

    def join():
        if nfl > 1:
            for s_ in streams:
                ev = torch.cuda.Event()
                ev.record(s_)
                torch.cuda.current_stream(dev).wait_event(ev)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    fork()
    for i in range(max(args.warmup, 3) * nfl):
        out = step(i)
    join()
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fork()
    for i in range(args.steps):
        out = step(i)
    join()
    e1.record()
    sync_all()
    sampler.stop_flag = True
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    sampler.join(timeout=1.0)
    assert torch.isfinite(out[0]).all(), "non-finite logits"
    value = B * world * args.steps / (ms * 1e-3)

    # ---- e2e: host-buffer C-ABI call, pinned inputs, H2D + forward + D2H of the logits inside the timed region
    hreqs = [[t.pin_memory() for i, t in enumerate(r) if i != 6] for r in reqs]
    houts = [{"vil_prediction": torch.empty(B, 3129, dtype=torch.float32).pin_memory()} for _ in range(nfl)]
    out_bytes = houts[0]["vil_prediction"].numel() * 4

    def estep(i):
        # host-buffer C-ABI call on slot/stream i % nfl: H2D of this step's inputs, forward, D2H of its logits -- all inside
        # the timed region.  A slot is re-used only after its previous step has completed (its pinned logits are final then),
        # so with nfl slots one batch's copies overlap another batch's kernels.
        q, f, s, seg, im, vm, tk = hreqs[i % args.rotate]
        j = i % nfl
        if nfl == 1:
            return model.forward_host(q, f, s, seg, im, vm, tk, houts[0], select=L.OUT_VIL_PREDICTION)
        streams[j].synchronize()
        with torch.cuda.stream(streams[j]):
            model.forward_host(q, f, s, seg, im, vm, tk, houts[j], select=L.OUT_VIL_PREDICTION, slot=j, synchronize=False)

    for i in range(max(args.warmup, 3) * nfl):
        estep(i)
    sync_all()
    t0 = time.perf_counter()
    for i in range(args.steps):
        estep(i)
    sync_all()                         # every stream drained: the logits of all steps are in host memory
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = B * world * args.steps / e2e_s
    hout = houts[(args.steps - 1) % nfl]
    # sanity: host path and device path agree
    chk = model(*dreqs[(args.steps - 1) % args.rotate], select=L.OUT_VIL_PREDICTION)[0].cpu()
    assert torch.allclose(chk, hout["vil_prediction"], atol=1e-5), "host/device C-ABI paths disagree"

    pk = peaks()
    tflops = flops * args.steps / (ms * 1e-3) / 1e12           # per GPU (ms is the max over ranks)
    # ---- per-kernel times, live: every kernel of the step replayed from its own CUDA graph between two CUDA events
    ops = model.profile_ops(B, Tin, V, select, iters=5)
    if args.ops_table:
        agg = {}
        for o in ops:
            k = (o["kind"],) + tuple(o["dims"])
            a_ = agg.setdefault(k, [0, 0.0, 0.0])
            a_[0] += 1; a_[1] += o["ms"]; a_[2] += o["flops"]
        rows = [dict(kind=k[0], dims=list(k[1:]), launches=v[0], total_us=round(v[1] * 1e3, 1), us=round(v[1] * 1e3 / v[0], 2),
                     tflops=round(v[2] / (v[1] * 1e-3) / 1e12, 1) if v[2] else None) for k, v in agg.items()]
        rows.sort(key=lambda r: -r["total_us"])
        with open(args.ops_table, "w") as f:
            for r in rows:
                f.write(json.dumps(r) + "\n")
    fam = {}
    for o in ops:
        f = fam.setdefault(o["kind"], {"launches": 0, "ms": 0.0, "flops": 0.0})
        f["launches"] += 1; f["ms"] += o["ms"]; f["flops"] += o["flops"]
    g = fam["gemm"]
    gemm_tflops = g["flops"] / (g["ms"] * 1e-3) / 1e12
    serial_ms = sum(f["ms"] for f in fam.values())
    top = max((o for o in ops if o["kind"] == "gemm"), key=lambda o: o["flops"])
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            traffic = json.load(f).get("gemm_dram_bytes_per_launch")
    except Exception:
        pass
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"batch={B} per GPU, VQA head, {V} regions x {Tin} tokens (BASELINE.json configs[1]), "
                                   f"{args.dtype} tensor-core operands / fp32 accumulate (tcgen05 kind::f16, same rate as bf16); "
                                   f"random-init 268M-param ViLBERT (seed 42)",
                       "global_batch": B * world, "per_gpu_batch": B, "n_tokens": Tin, "n_regions": V,
                       "parallelism": f"dp{world} (batch sharding, no collective)",
                       "l2": f"inputs rotate over {args.rotate} resident batches ({args.rotate * in_bytes / 1e6:.0f} MB) "
                             f"+ {model._dims['weight_bytes'] / 1e6:.0f} MB of weights > 126 MB L2",
                       "batches_in_flight": nfl,
                       "cuda_graph": not args.no_graph, "pdl": "every kernel" if args.pdl != "off" else "off", "layernorm": "fused" if args.fused_ln else "split",
                       "heads": "task heads" if args.all_heads else "vil_prediction"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
                    "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": int(n_launch) * args.steps,
            "launches_per_step": int(n_launch),
            "roofline": {"bound": "tensor", "achieved": gemm_tflops, "peak": pk["bf16_burst"], "unit": "TFLOP/s",
                         "frac": gemm_tflops / pk["bf16_burst"], "traffic": traffic,
                         "kernel": "gemm_persistent_kernel (tcgen05): all %d GEMM launches of one step, algorithmic 2MNK FLOPs / "
                                   "CUDA-event time per launch" % g["launches"],
                         "avg_launch_us": 1e3 * g["ms"] / g["launches"],
                         "largest_gemm": {"M": top["dims"][0], "N": top["dims"][1], "K": top["dims"][2], "us": 1e3 * top["ms"],
                                          "tflops": top["flops"] / (top["ms"] * 1e-3) / 1e12},
                         "share_of_step": g["ms"] / serial_ms,
                         "families_ms": {k: round(v["ms"], 4) for k, v in fam.items()},
                         "whole_step_tflops": tflops, "whole_step_frac": tflops / pk["bf16_sustained"],
                         "flops_per_step": flops,
                         "peak_source": pk["src"] + ": burst cuBLAS bf16 for the kernels timed alone (peak), sustained %.0f for the whole step"
                                        % pk["bf16_sustained"]},
            "clocks": sampler.result()}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        model_cpu = oracle_model(sd, cfg.to_dict(), 3129)
        nb = 3
        times = time_oracle(model_cpu, reqs[0], nb, 1)
        v = B * len(times) / sum(times)
        import torch as _t
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": CPU_THREADS, "kind": "port",
                                "sample": f"{nb} forwards of batch {B} after 1 warm-up (torch fp32 oracle, "
                                          f"{_t.get_num_threads()} threads of {os.cpu_count()} vCPUs, all heads + pre-training heads)"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    model.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
