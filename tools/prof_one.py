import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
import bench
from catears_b200 import api, synth
import torch
conf = os.path.join(bench.model_dir(), "tdnn.conf")
n = int(os.environ.get("N_UTTS", "128"))
pcm, off = synth.synth_batch(n, 160000)
m = api.AcousticModelGpu(config=conf, precision=os.environ.get("PRECISION", "int8"))
d_pcm = torch.from_numpy(pcm).cuda()
frames = int(api.frame_offsets(off)[-1])
d_ll = torch.empty((frames, m.num_pdfs), dtype=torch.float32, device="cuda")
d_am = torch.empty(frames, dtype=torch.int32, device="cuda")
for i in range(int(os.environ.get("PASSES", "3"))):
    if i == 2: sys.stderr.write("---- measured pass\n")
    m.forward(d_pcm, off, loglik=d_ll, argmax=d_am)
    torch.cuda.synchronize()
