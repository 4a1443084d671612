"""Per-kernel-class table of the generation-mode chain (CelebA-64, B = 4096, all 8 steps): python scripts/gen_profile.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import seqvae_b200 as S
GB = int(os.environ.get("GEN_B", "4096"))
ds = S.SyntheticDataset("celebA", GB, seed=1)
m = S.SequentialVAE(ds, GB, "c_inhomog", operand_dtype="bf16", restore=False, seed=0, train=False)
st = torch.cuda.Stream(); m.use_torch_stream(st)
T, (H, W, C) = m.mc_steps, m.data_dims
with torch.cuda.stream(st):
    out = torch.empty([T, GB, H, W, C], device="cuda", dtype=torch.float32)
for i in range(2): m.generate_async(GB, out, None, seed=i)
m.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for i in range(3): m.generate_async(GB, out, None, seed=5 + i)
e1.record(st); m.sync(); torch.cuda.synchronize()
print("generation: %.2f ms per chain, %.0f img/s" % (e0.elapsed_time(e1) / 3, GB / (e0.elapsed_time(e1) / 3e3)))
m.profile(True)
m.generate_async(GB, out, None, seed=9); m.sync()
p = m.profile_read(); m.profile(False)
tot = sum(v["ms"] for v in p.values())
for k, v in sorted(p.items(), key=lambda kv: -kv[1]["ms"]):
    print("  %-22s n=%4d ms=%8.3f share=%5.1f%%  %8.1f TF/s %8.1f GB/s" % (k, v["launches"], v["ms"], 100 * v["ms"] / tot, v["flops"] / (v["ms"] * 1e9) if v["ms"] else 0, v["bytes"] / (v["ms"] * 1e6) if v["ms"] else 0))
print("serialised sum %.2f ms" % tot)
