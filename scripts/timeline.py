"""Per-stream phase timeline of one training step (SVAE_TIMELINE=1: eager forked execution, events at phase boundaries)."""
import os, sys
os.environ["SVAE_TIMELINE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import seqvae_b200 as S
B = 100
ds = S.SyntheticDataset("celebA", B, seed=1)
model = S.SequentialVAE(ds, B, "c_inhomog", operand_dtype="bf16", restore=False, seed=0)
st = torch.cuda.Stream(priority=-1)
model.use_torch_stream(st)
x = torch.from_numpy(ds.next_batch(B)).cuda()
for i in range(4):
    model.train_async(x, x)
    if i < 3:
        torch.cuda.synchronize()
        import ctypes
model.sync()
