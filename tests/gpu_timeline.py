"""Diagnostics: kernel timeline of one bench step (K concurrent solves) through torch.profiler (CUPTI).

    python tests/gpu_timeline.py [workload] [K]   -> gpurun_out/timeline_<wl>_<K>.csv + a summary on stdout
"""
import os
import sys
import json
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
from qiskit_addon_sqd_b200 import fermion  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda", 0)
norb, nelec, h, g, batches = bench.make_batches(wl, 0, K)
opts = fermion._solver_options({})
ints = fermion._DeviceIntegrals(torch, h, g, dev)
strs_dev = [(torch.from_numpy(a.astype(np.uint64).view(np.int64)).to(dev),
             torch.from_numpy(b.astype(np.uint64).view(np.int64)).to(dev)) for a, b in batches]
streams = [torch.cuda.Stream(device=dev) for _ in range(K)]
pool = ThreadPoolExecutor(max_workers=K)


def step():
    main = torch.cuda.current_stream()

    def work(k):
        with torch.cuda.device(dev), torch.cuda.stream(streams[k]):
            streams[k].wait_stream(main)
            return fermion._solve_on_device(batches[k][0], batches[k][1], norb, ints, None, 0.2, opts,
                                            want_spin=False, want_rdm=False, strs_dev=strs_dev[k],
                                            download=False,
                                            throughput=fermion._throughput_mode(
                                                K, sum(len(a) * len(b) for a, b in batches)))

    res = list(pool.map(work, range(K)))
    for s in streams:
        main.wait_stream(s)
    torch.cuda.synchronize()
    return res


if len(sys.argv) > 3 and sys.argv[3] == "e2e":
    def step():  # noqa: F811  -- the public plugin call with host inputs and outputs (RDMs included)
        out = fermion.solve_sci_batch(batches, h, g, norb, nelec)
        torch.cuda.synchronize()
        return out

for _ in range(3):
    step()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
os.makedirs("gpurun_out", exist_ok=True)
trace = f"gpurun_out/trace_{wl}_{K}.json"
prof.export_chrome_trace(trace)
ev = json.load(open(trace))["traceEvents"]
kern = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
kern.sort(key=lambda e: e["ts"])
t0 = kern[0]["ts"]
t1 = max(e["ts"] + e["dur"] for e in kern)
span = t1 - t0
busy_sum = sum(e["dur"] for e in kern)
# union of busy intervals
union, cur_s, cur_e = 0.0, None, None
for e in kern:
    s, en = e["ts"], e["ts"] + e["dur"]
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            union += cur_e - cur_s
        cur_s, cur_e = s, en
    else:
        cur_e = max(cur_e, en)
union += cur_e - cur_s
print(f"workload {wl} K={K}: span {span/1e3:.2f} ms, kernels {len(kern)}, sum of durations {busy_sum/1e3:.2f} ms, "
      f"device busy (union) {union/1e3:.2f} ms ({100*union/span:.0f}% of span), mean concurrency {busy_sum/union:.2f}")
by = {}
for e in kern:
    n = e["name"].split("(")[0].replace("void ", "").replace("sqd::", "")[:40]
    d = by.setdefault(n, [0, 0.0])
    d[0] += 1
    d[1] += e["dur"]
for n, (c, d) in sorted(by.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f"  {n:42s} {c:6d} launches {d/1e3:8.2f} ms  avg {d/c:7.1f} us  {100*d/busy_sum:5.1f}%")
# per-stream: busy vs idle inside its own active window
streams_seen = {}
for e in kern:
    streams_seen.setdefault(e["args"].get("stream"), []).append(e)
for sid, es in sorted(streams_seen.items(), key=lambda kv: -len(kv[1]))[:K]:
    a, b = es[0]["ts"], max(x["ts"] + x["dur"] for x in es)
    busy = sum(x["dur"] for x in es)
    print(f"  stream {sid}: {len(es)} kernels, window {(b-a)/1e3:.2f} ms, busy {busy/1e3:.2f} ms ({100*busy/(b-a):.0f}%)")
with open(f"gpurun_out/timeline_{wl}_{K}.csv", "w") as f:
    f.write("ts_us,dur_us,stream,name\n")
    for e in kern:
        f.write(f"{e['ts']-t0:.3f},{e['dur']:.3f},{e['args'].get('stream')},{e['name'].split('(')[0][:60]}\n")
# CPU-side: cuda runtime calls
rt = [e for e in ev if e.get("cat") == "cuda_runtime" and "dur" in e]
# Was the device waiting for the host, or the host-issued kernel waiting for the device?  For every
# kernel: ready = max(end of its launch call, end of the previous kernel of its stream);
# delay = start - ready.  "late launch" = the previous kernel had already finished when the launch returned.
launch_end = {}
for e in rt:
    cid = e.get("args", {}).get("correlation")
    if cid is not None:
        launch_end[cid] = e["ts"] + e["dur"]
prev_end = {}
n_late = 0
late_wait = dev_wait = 0.0
delays = []
for e in kern:
    sid = e["args"].get("stream")
    cid = e["args"].get("correlation")
    le = launch_end.get(cid)
    pe = prev_end.get(sid)
    if le is not None and pe is not None:
        if le > pe:
            n_late += 1
            late_wait += le - pe           # stream idle because the host had not launched yet
            dev_wait += e["ts"] - le        # launch -> start latency
        else:
            dev_wait += e["ts"] - pe        # kernel was queued; device took this long to start it
        delays.append(e["ts"] - max(le, pe))
    prev_end[sid] = e["ts"] + e["dur"]
if delays:
    d = np.array(delays)
    print(f"  launches that arrived after the previous kernel of the stream had finished: {n_late} of {len(d)}; "
          f"stream-idle time waiting for the host {late_wait/1e3:.2f} ms; start delay after ready: "
          f"total {d.sum()/1e3:.2f} ms, median {np.median(d):.1f} us, p90 {np.percentile(d, 90):.1f} us")
byrt = {}
for e in rt:
    d = byrt.setdefault(e["name"], [0, 0.0])
    d[0] += 1
    d[1] += e["dur"]
for n, (c, d) in sorted(byrt.items(), key=lambda kv: -kv[1][1])[:8]:
    print(f"  runtime {n:32s} {c:6d} calls {d/1e3:8.2f} ms  avg {d/c:7.1f} us")
os.remove(trace)
