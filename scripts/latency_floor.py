"""Step time of the CelebA chain as a function of the batch size: at B = 2 the kernels do almost no work, what remains is the
dispatch / dependency latency of the ~1170-node graph.  usage: python scripts/latency_floor.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import seqvae_b200 as S
for B in [int(v) for v in os.environ.get("FLOOR_BS", "2,8,25,50,100").split(",")]:
    ds = S.SyntheticDataset("celebA", B, seed=1)
    model = S.SequentialVAE(ds, B, "c_inhomog", operand_dtype="bf16", restore=False, seed=0)
    st = torch.cuda.Stream(priority=-1)
    model.use_torch_stream(st)
    x = torch.from_numpy(ds.next_batch(B)).cuda()
    for i in range(5):
        model.train_async(x, x)
    model.sync(); torch.cuda.synchronize()
    n0 = model.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record(st)
        for i in range(10):
            model.train_async(x, x)
        e1.record(st)
    model.sync(); torch.cuda.synchronize()
    print("B=%3d  ms/step %.3f  launches/step %d" % (B, e0.elapsed_time(e1) / 10, (model.launch_count - n0) // 10), flush=True)
    model.close()
