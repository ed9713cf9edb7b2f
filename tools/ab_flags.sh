#!/bin/bash
# tools/ab_flags.sh WORKLOAD FLAGSET...  -> phases of the workload for each DL_FLAGS setting ("-" = default)
cd "$(dirname "$0")/.."
wl=$1; shift
for f in "$@"; do
  name=$(echo "$f" | tr ',' '_')
  if [ "$f" = "-" ]; then unset DL_FLAGS; name=default; else export DL_FLAGS=$f; fi
  timeout 600 python bench.py --workload $wl --no-e2e --no-cpu > gpurun_out/ab_${wl}_$name.json 2> gpurun_out/ab_${wl}_$name.err
  python - "$wl" "$name" <<'PY'
import json, sys
wl, n = sys.argv[1:3]
try:
    d = json.load(open(f"gpurun_out/ab_{wl}_{n}.json"))
    p = d["phases_ms"]
    print(f"{wl} {n:22s} value {d['value']/1e9:.3f} G/s  attn {p['attn_fwd']:.2f} spmm {p['spmm_fwd']:.2f} pairf {p['pair_fwd']:.2f} pairb {p['pair_bwd']:.2f} gath {p['bwd_gather']:.2f} edges {p['bwd_edges']:.2f}  loss {d['loss']:.7f} clk {d['clocks']['sm_mhz']}")
except Exception as e:
    print(wl, n, "FAILED", e)
PY
done
