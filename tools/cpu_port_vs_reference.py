"""Build container only: time the UNMODIFIED reference (oracle/ref_harness.py) and the CPU oracle port
(oracle/sdnet_oracle.py) on the same bounded sample of the cfg-3 shape, same host, same thread count, so that
the bias of the port that bench.py times on the GPU box (`cpu_baseline.kind = "port"`) is visible
(VERDICT r1 weak #12).  Writes profiles/r02_cpu_port_vs_reference.json."""
import copy
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import ref_harness, sdnet_oracle  # noqa: E402
from ruart_b200 import synth  # noqa: E402

n_q = int(sys.argv[1]) if len(sys.argv) > 1 else 16
threads = os.cpu_count()
torch.set_num_threads(threads)
cfg = bench.sample_cfg("cfg3", n_q)
opt = synth.make_opt(cfg)
batch = synth.make_batch(cfg, seed=2003)
net = ref_harness.build_reference(opt, seed=1033)
ref_harness.run_reference(net, batch)             # warm-up
t0 = time.perf_counter()
p_ref, _, _ = ref_harness.run_reference(net, batch)
t_ref = time.perf_counter() - t0
sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
with torch.no_grad():
    sdnet_oracle.sdnet_forward(sd, opt, *copy.deepcopy(batch))
    t0 = time.perf_counter()
    p_port, _, _ = sdnet_oracle.sdnet_forward(sd, opt, *copy.deepcopy(batch))
    t_port = time.perf_counter() - t0
res = {"sample": "%d questions of the cfg3 shape, fp32, torch %s CPU" % (n_q, torch.__version__), "cores": threads,
       "reference_questions_per_s": n_q / t_ref, "port_questions_per_s": n_q / t_port,
       "port_over_reference": t_ref / t_port, "max_abs_dprob": float((p_ref - p_port).abs().max()),
       "where": "build container (the reference tree does not travel to the GPU box)"}
with open(os.path.join(ROOT, "profiles", "r02_cpu_port_vs_reference.json"), "w") as f:
    json.dump(res, f, indent=1)
print(json.dumps(res))
