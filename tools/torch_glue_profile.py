"""torch.profiler view of one PEMS07 training step: which aten ops (the glue around the library's kernels --
gradient adds, fills, copies, cat / pad, loss, Adam) launch how many CUDA kernels and cost how much device time.

    python tools/torch_glue_profile.py [workload] > gpurun_out/torch_glue.txt
"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from dstagnn_drought_b200 import make_model  # noqa: E402
from dstagnn_drought_b200.parallel import GradBuckets  # noqa: E402


def main():
    w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "pems07"]
    dev = torch.device("cuda:0")
    tmd, pa = bench.synthetic_graph(w["N"], 0)
    torch.manual_seed(1)
    net = make_model(dev, 1, w["nb"], 1, w["K"], 32, 32, 1, torch.from_numpy(tmd), torch.from_numpy(pa),
                     torch.from_numpy(tmd), w["P"], w["T"], w["N"], w["D"], 32, 32, w["H"])
    net.train()
    buckets = GradBuckets(net, overlap=True)
    opt = torch.optim.Adam([p for p in net.parameters() if p.grad is not None], lr=1e-4, fused=True)
    crit = torch.nn.SmoothL1Loss()
    x = torch.randn(w["B"], w["N"], 1, w["T"], device=dev)
    y = torch.randn(w["B"], w["N"], w["P"], device=dev)

    def step():
        buckets.zero()
        loss = crit(net(x), y)
        loss.backward()
        buckets.finish()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=False) as prof:
        step()
        torch.cuda.synchronize()
    print(prof.key_averages(group_by_input_shape=False).table(sort_by="self_cuda_time_total", row_limit=45,
                                                              max_name_column_width=40, max_shapes_column_width=70))


if __name__ == "__main__":
    main()
