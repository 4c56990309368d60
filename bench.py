#!/usr/bin/env python
"""bench.py -- frames/sec of the DEFT per-frame dense hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (1 process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU (oracle port)

Workload (BASELINE config C2, "MOT17-shape"): synthetic 1088x608 RGB stream, DLA-34+DCNv2 forward, sigmoid,
decode K=100, AFE extractor for the 100 detections, affinity against a 5-frame x 100 = 500-embedding track memory.
A step = one frame.  N GPUs = N independent video streams (weak scaling, no collective on the data path).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, K_DETS, MEM_FRAMES, MAX_OBJECT = 608, 1088, 100, 5, 100
P4 = (H // 4) * (W // 4)
DCN_BYTES_PER_FRAME = 5018.7 * P4 + 13.12e6      # algorithmic HBM bytes of the 16 DCN layers
METRIC = "frames/sec @1088x608, 100 dets x 500 tracks"
WORKLOAD = "C2 MOT17-shape 1088x608 stream, DLA-34+DCNv2, K=100 dets, 5x100 track-memory affinity"


def ncu_traffic():
    """dram__bytes_read+write per deft_dcnv2_forward launch from the committed ncu --set full capture (profiles/)."""
    path = os.path.join(ROOT, "profiles", "r1_dcn_ncu_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
        return {"traffic": t["dram_bytes_per_launch_avg"], "traffic_source": "profiles/r1_dcn_ncu_traffic.json (ncu --set full, gather-GEMM kernel; "
                "outputs stay in L2 under ncu's replay, so this is the cold read side)"}
    except (OSError, KeyError, ValueError):
        return {"traffic": None}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
                time.sleep(0.02)
        except Exception as e:  # NVML missing: report it, never fake numbers
            self.reasons.add("nvml_unavailable:" + type(e).__name__)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm
def oracle_frame(sd, heads, x, memory, restate, torch):
    """One frame of the reference algorithm on CPU (oracle/restate.py == the reference's own PyTorch ops)."""
    out, fm = restate.dlaseg_forward(sd, x, heads)
    sig = restate.sigmoid_output(out[0])
    dec = restate.generic_decode(sig, K=K_DETS)
    oh, ow = sig["hm"].shape[2:]
    centres = restate.convert_detection(dec["bboxes"][0].numpy(), oh, ow)
    emb = restate.afe_extract(sd, fm, centres)
    affs = [restate.afe_affinity(sd, memory[f:f + 1], emb, MAX_OBJECT, True) for f in range(MEM_FRAMES)]   # tracker.py:76-90
    return emb, affs


def run_cpu(steps, warmup, budget_s, state_dict=None):
    import torch
    from oracle import restate, spec
    heads = spec.MOT_HEADS
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = state_dict if state_dict is not None else spec.make_state_dict(spec.dlaseg_spec(heads))
    x = spec.make_input((1, 3, H, W), tag="bench")
    memory = spec.make_input((MEM_FRAMES, MAX_OBJECT, 416), tag="benchmem").abs()
    times = []
    t_start = time.time()
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.time()
            emb, _ = oracle_frame(sd, heads, x, memory, restate, torch)
            dt = time.time() - t0
            if i >= warmup:
                times.append(dt)
            if time.time() - t_start > budget_s and len(times) >= 1:
                break
    n = len(times)
    total = sum(times)
    return {"fps": n / total, "steps": n, "ms_per_step": 1000 * total / n, "cores": torch.get_num_threads()}


def reference_arm(args, rank):
    if rank != 0:
        return
    r = run_cpu(args.steps, min(args.warmup, 1), budget_s=150.0)
    line = {"impl": "reference", "metric": METRIC, "value": r["fps"], "unit": "frames/s", "n_gpus": args.gpus, "steps": r["steps"],
            "warmup": min(args.warmup, 1), "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "reference algorithm on host CPU cores (oracle port of the reference's PyTorch modules; "
                                                     "DCN = torchvision.ops.deform_conv2d since upstream DCNv2 is un-vendored); "
                                                     "timed steps bounded to ~150 s"},
            "cpu_baseline": {"value": r["fps"], "unit": "frames/s", "cores": r["cores"], "kind": "port", "sample": f"{r['steps']} full C2 frames"},
            "e2e": {"value": r["fps"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def gpu_arm(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from deft_b200.model.model import create_model
    from deft_b200.opts import Opt
    from deft_b200.parallel import aggregate_throughput
    from deft_b200.pipeline import FramePipeline
    from deft_b200.synthetic import randomize_

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    opt = Opt(K=K_DETS, max_object=MAX_OBJECT, tensor_core=not args.simt, tf32x1=args.tf32x1)
    heads = opt.heads
    model = create_model("dla_34", dict(heads), {h: [256] for h in heads}, opt).eval()
    randomize_(model, seed=317 + rank)
    model = model.to(dev)
    pipe = FramePipeline(model, opt, H, W, mem_frames=MEM_FRAMES, device=dev, use_cuda_graph=not args.no_graph, depth=args.depth, dense_streams=args.dense_streams)
    g = torch.Generator(); g.manual_seed(1234 + rank)
    frames_host = [torch.randn((1, 3, H, W), generator=g).pin_memory() for _ in range(4)]    # distinct frames, pinned
    frames_dev = [f.to(dev) for f in frames_host]
    pipe.x.copy_(frames_dev[0])
    pipe.warmup()
    out_host = None
    for i in range(max(args.warmup, 3)):
        out_host = pipe.step(frames_host[i % 4], out_host)
    torch.cuda.synchronize(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(step_fn):
        """K back-to-back steps between one CUDA-event pair on the launching stream (the pipeline's second stream is
        joined into it before the closing event), barrier + device synchronize on both sides, max over ranks.
        No L2 flush: one frame touches ~1 GB of unique bytes (0.33 GB packed weights + tile images, 0.8 GB
        activation arena, two arenas alternating) >> 126 MB L2, and the input frames rotate."""
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            step_fn(i)
        pipe.join()
        e1.record()
        barrier()
        sampler.stop_flag = True
        sampler.join()
        ms = e0.elapsed_time(e1)
        _, ms_max = aggregate_throughput(float(args.steps), ms, device=dev)      # max over ranks of the device time
        return ms_max, sampler.summary()

    # (1) device-resident value: frames already in HBM
    ms_dev, clocks = timed(lambda i: pipe.step_device(frames_dev[i % 4]))
    # (2) end to end through the public API: pinned host frame in, host results out, every copy inside the timed
    #     region; each frame arrives cold from the host and a frame's 1.4 GB working set >> L2, so no extra flush
    ms_e2e, clocks_e2e = timed(lambda i: pipe.step(frames_host[i % 4], out_host))
    # (2b) latency of ONE isolated frame (submit, then wait for its results): what a depth-1, one-frame-at-a-time caller sees
    lat = []
    for i in range(5):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        pipe.step(frames_host[i % 4], out_host)
        pipe.synchronize()
        lat.append((time.perf_counter() - t0) * 1e3)
    latency_ms = sorted(lat)[len(lat) // 2]
    h2d = frames_host[0].numel() * 4
    d2h = sum(v.numel() * v.element_size() for v in out_host.values())

    # (3) per-kernel DCN timing for the roofline: CUDA events around every deft_dcnv2_forward of the eager launch list
    dcn_ms = []
    for _ in range(3):
        prof = pipe.plan.launch_profiled()
        dcn_ms.append(sum(ms for name, _, ms in prof if name == "deft_dcnv2_forward"))
    dcn_ms = sorted(dcn_ms)[1]
    peak, peak_src = peaks()
    achieved = DCN_BYTES_PER_FRAME / (dcn_ms * 1e-3) / 1e9
    n_dcn = sum(1 for name, _, _ in prof if name == "deft_dcnv2_forward")
    launches = pipe.count_launches()

    line = None
    if rank == 0:
        fps = world * args.steps / (ms_dev * 1e-3)
        fps_e2e = world * args.steps / (ms_e2e * 1e-3)
        line = {"metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if args.simt else ("tf32" if args.tf32x1 else "f32 (3xTF32 tensor-core where covered, fp32 SIMT elsewhere)"),
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "streams_per_gpu": 1, "parallelism": f"stream-parallel x{world} (no collective)",
                           "cuda_graph": not args.no_graph, "frame_pipelining": f"{pipe.depth} frames of the stream in flight: {pipe.S} dense stream(s) + 1 tail stream (results bit-identical to depth 1, delivered in frame order)",
                           "l2": "no flush: inputs larger than L2 (~1 GB of unique weights + activations per frame vs 126 MB L2, rotating input frames)",
                           "engine": "simt-fp32" if args.simt else "tcgen05"},
                "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
                "frame_latency_ms": latency_ms,
                "gpu_launches": launches * args.steps,
                "gpu_launches_per_step": launches,
                "clocks": clocks, "clocks_e2e": clocks_e2e,
                "roofline": {"bound": "hbm", "kernel": "deft_dcnv2_forward (offset/mask conv + fused gather-GEMM), 16 layers/frame",
                             "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                             "algorithmic_bytes_per_frame": DCN_BYTES_PER_FRAME, "dcn_ms_per_frame": dcn_ms, "dcn_layers": n_dcn,
                             "algorithmic_bytes_per_launch": DCN_BYTES_PER_FRAME / max(n_dcn, 1), "launch_ms_avg": dcn_ms / max(n_dcn, 1),
                             **ncu_traffic()}}
        if world == 1 and not args.no_cpu_baseline:
            sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
            r = run_cpu(2, 0, budget_s=40.0, state_dict=sd)
            line["cpu_baseline"] = {"value": r["fps"], "unit": "frames/s", "cores": r["cores"], "kind": "port",
                                    "sample": f"{r['steps']} full C2 frames of the oracle (reference modules' PyTorch-CPU arithmetic, DCN=torchvision deform_conv2d)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)      # >= 200 timed frames (0.4 s of GPU time)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="deft_b200", choices=["deft_b200", "reference"])
    ap.add_argument("--simt", action="store_true", help="fp32 SIMT anchor engine only")
    ap.add_argument("--tf32x1", action="store_true", help="single-pass TF32 (outside the parity budget)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--depth", type=int, default=4, choices=[1, 2, 3, 4], help="FramePipeline: activation arenas = frames of the stream in flight")
    ap.add_argument("--dense-streams", type=int, default=None, help="FramePipeline: CUDA streams the dense parts rotate over (default depth - 1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        sys.exit("launch multi-GPU runs with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 "
                 "--master-port P bench.py --gpus N ...")
    gpu_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
