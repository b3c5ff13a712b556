#!/usr/bin/env python
"""Per-call latency of the streaming form on the bench model (int8): N live streams, each getting
`ms` milliseconds of new audio per call; host-resident state (ce_host::StreamBatch) against
device-resident state (ce_gpu_streams_* / ce_host::DeviceStreamBatch)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    d = bench.model_dir()
    exe = "/tmp/ce_stream_probe"
    subprocess.check_call(["g++", "-std=c++11", "-O2", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cc"), "-o", exe,
                           "-L" + os.path.join(ROOT, "catears_b200"), "-lce_gpu",
                           "-Wl,-rpath," + os.path.join(ROOT, "catears_b200")])
    # rows as the 64 best (loglik, pdf) pairs: dense rows would make this a measurement of moving
    # 12 KB a frame through host memory, not of the streaming machinery
    env = dict(os.environ, HOST_MIRROR_SELECT=os.environ.get("HOST_MIRROR_SELECT", "topk:64"))
    for n, ms in ((64, 100), (512, 100), (512, 500), (2048, 100)):
        subprocess.check_call([exe, "streambench", os.path.join(d, "tdnn.conf"), "0", os.path.join(d, "tdnn.cmvn"),
                               str(n), str(16 * ms), "21"], stdin=subprocess.DEVNULL, timeout=240, env=env)


if __name__ == "__main__":
    main()
