#!/bin/bash
# tools/ab_bench.sh NAME...  -> phases of the mid workload for each variant library (and the default build first)
cd "$(dirname "$0")/.."
run() {
  timeout 300 python bench.py --workload mid --no-e2e --no-cpu > gpurun_out/ab_$1.json 2> gpurun_out/ab_$1.err
  python - "$1" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/ab_{n}.json"))
    p = d["phases_ms"]
    print(f"{n:10s} value {d['value']/1e9:.3f} G/s  attn {p['attn_fwd']:.2f} spmm {p['spmm_fwd']:.2f} pairf {p['pair_fwd']:.2f} pairb {p['pair_bwd']:.2f} gath {p['bwd_gather']:.2f} edges {p['bwd_edges']:.2f}  loss {d['loss']:.6f}")
except Exception as e:
    print(n, "FAILED", e)
PY
}
unset DL_LIB_PATH
run default
for v in "$@"; do
  export DL_LIB_PATH=$PWD/disenlink_b200/_variants/lib_$v.so
  run $v
done
