"""Host-side profile of one DeepIC forward (run_example.py's tracker, one 120x160 pair): where the 7 ms go."""
import copy, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from baseline import reference as REF
from deep_prob_feature_track_b200 import algorithms as A

dev = torch.device("cuda:0")
flags = ["--encoder_name", "ConvRGBD2", "--mestimator", "MultiScale2w", "--solver", "Direct-ResVol", "--uncertainty", "None"]
net = A.patch_tracker(REF.make_tracker(flags, seed=0).to(dev)).eval()
batch = REF.synthetic_rgbd(1, 120, 160, seed=3, device=dev)
with torch.no_grad():
    for _ in range(5):
        net(*batch)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(20):
        net(*batch)
    torch.cuda.synchronize()
    print(f"forward: {(time.perf_counter() - t) / 20 * 1e3:.2f} ms")
    t = time.perf_counter()
    for _ in range(20):
        net._preprocess(*batch)
    torch.cuda.synchronize()
    print(f"_preprocess alone: {(time.perf_counter() - t) / 20 * 1e3:.2f} ms")
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            net(*batch)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=60))
    ev = prof.key_averages()
    print("cuda kernels per forward:", sum(e.count for e in ev if e.device_type.name == "CUDA") / 5)
