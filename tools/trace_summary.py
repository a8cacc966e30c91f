#!/usr/bin/env python
"""Timeline of the LAST step in a chrome trace written by `bench.py --trace`: every kernel / memcpy with start offset,
duration and stream, so that exposed tails and serialisation are visible.  usage: trace_summary.py trace.json [N]"""
import json
import sys

ev = [e for e in json.load(open(sys.argv[1]))["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
# one step = the events between two consecutive launches of a marker kernel (default: the loss kernel)
marker = sys.argv[2] if len(sys.argv) > 2 else "softmax_xent"
idx = [i for i, e in enumerate(ev) if marker in e["name"]]
step = ev[idx[-2] + 1: idx[-1] + 1] if len(idx) >= 2 else ev
t0 = step[0]["ts"]
print(f"{len(step)} events, step span {max(e['ts'] + e['dur'] for e in step) - t0:.1f} us")
for e in step:
    name = e["name"].replace("ed2::(anonymous namespace)::", "").replace("void ", "")[:70]
    print(f"{e['ts'] - t0:9.1f} +{e['dur']:8.1f}  s{e['args'].get('stream', '?'):<4} {name}")
