"""Kernel timeline of graph-replayed training steps (CUPTI through torch.profiler): per-stream busy time, idle gaps of the
chain stream, and the kernels on it.  usage: python scripts/trace_step.py [out.json]   (B200 box)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import seqvae_b200 as S

B = int(os.environ.get("TRACE_B", "100"))
ds = S.SyntheticDataset("celebA", B, seed=1)
model = S.SequentialVAE(ds, B, "c_inhomog", operand_dtype="bf16", restore=False, seed=0)
st = torch.cuda.Stream(priority=-1)
model.use_torch_stream(st)
x = torch.from_numpy(ds.next_batch(B)).cuda()
for i in range(5):
    model.train_async(x, x)
model.sync(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(2):
        model.train_async(x, x)
    model.sync(); torch.cuda.synchronize()
ev = []
for e in prof.events():
    if e.device_type is not None and "cuda" in str(e.device_type).lower():
        ev.append((e.time_range.start, e.time_range.end, e.name))
# torch's FunctionEvent does not expose the stream: read the chrome trace instead
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/trace_step.json"
prof.export_chrome_trace(out + ".full")
tr = json.load(open(out + ".full"))
ks = [e for e in tr["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
ks.sort(key=lambda e: e["ts"])
t0 = ks[0]["ts"]
small = [dict(n=e["name"][:60], s=e["args"].get("stream"), t=round(e["ts"] - t0, 2), d=round(e["dur"], 2), g=e["args"].get("grid"), c=e.get("cat")) for e in ks]
json.dump(small, open(out, "w"))
os.remove(out + ".full")
print("kernels traced:", len(small), "span %.1f us" % (small[-1]["t"] + small[-1]["d"]))
