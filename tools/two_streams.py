"""Experiment (GPU box; not a pytest file): aggregate frame rate of S independent FramePipelines on S CUDA streams of one GPU."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tools/ sits next to deft_b200/
sys.path.insert(0, ROOT)
import torch
import bench
from deft_b200.model.model import create_model
from deft_b200.opts import Opt
from deft_b200.pipeline import FramePipeline
from deft_b200.synthetic import randomize_

dev = torch.device("cuda:0")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 2
H, W = 608, 1088
heads = {"hm": 1, "reg": 2, "wh": 2, "ltrb_amodal": 4, "tracking": 2}
pipes, streams, frames = [], [], []
for s in range(S):
    opt = Opt(dataset="mot", K=100)
    model = create_model("dla_34", dict(heads), {h: [256] for h in heads}, opt).eval()
    randomize_(model, seed=317 + s)
    model = model.to(dev)
    st = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(st):
        p = FramePipeline(model, opt, H, W, mem_frames=5, device=dev, depth=int(os.environ.get("DEPTH", "1")))
        f = torch.randn((1, 3, H, W), device=dev)
        p.x.copy_(f)
        p.warmup()
        for _ in range(3):
            p.step_device()
    st.synchronize()
    pipes.append(p); streams.append(st); frames.append(f)
torch.cuda.synchronize()
steps = 30
for rep in range(3):
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(S)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(S)]
    torch.cuda.synchronize()
    for s in range(S):
        e0[s].record(streams[s])
    for i in range(steps):
        for s in range(S):
            with torch.cuda.stream(streams[s]):
                pipes[s].step_device()
    for s in range(S):
        e1[s].record(streams[s])
    torch.cuda.synchronize()
    ms = max(e0[0].elapsed_time(e1[s]) for s in range(S))
    print(f"streams={S} frames={S * steps} ms={ms:.2f} aggregate fps={S * steps / ms * 1e3:.1f}", flush=True)
