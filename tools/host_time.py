import sys, os, time, numpy as np
sys.path.insert(0, os.getcwd())
import bench
from catears_b200 import api, synth
import torch
conf = os.path.join(bench.model_dir(), "tdnn.conf")
pcm, off = synth.synth_batch(512, 160000)
m = api.AcousticModelGpu(config=conf, precision="int8")
d_pcm = torch.from_numpy(pcm).cuda()
frames = int(api.frame_offsets(off)[-1])
d_ll = torch.empty((frames, m.num_pdfs), dtype=torch.float32, device="cuda")
d_am = torch.empty(frames, dtype=torch.int32, device="cuda")
s = torch.cuda.current_stream()
for i in range(3):
    m.forward(d_pcm, off, loglik=d_ll, argmax=d_am, stream=s)
torch.cuda.synchronize()
host = []
t_all = time.perf_counter()
for i in range(10):
    t0 = time.perf_counter(); m.forward(d_pcm, off, loglik=d_ll, argmax=d_am, stream=s); host.append(1e3 * (time.perf_counter() - t0))
torch.cuda.synchronize()
print("host enqueue ms per step:", [round(h, 2) for h in host], "total per step", round(1e3 * (time.perf_counter() - t_all) / 10, 2))
