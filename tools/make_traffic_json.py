"""profiles/traffic.json from an ncu summary written by tools/ncu_summary.py.
Usage: python tools/make_traffic_json.py profiles/r01_ncu_full_mid.txt NNZ P > profiles/traffic.json"""
import json
import re
import sys


def main(path, nnz, pairs):
    kernels, cur = {}, None
    for line in open(path):
        m = re.match(r"----- (.*)", line)          # every launch header, namespaced or not
        if m:
            name = re.sub(r"^void ", "", m.group(1).strip())
            name = re.sub(r"^(<unnamed>|\(anonymous namespace\))::", "", name)
            cur = kernels.setdefault(name, {})
            continue
        m = re.match(r"\s+(\S+)\s+([\d.]+) (\S+)", line)
        if m and cur is not None:
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(m.group(3), 1.0)
            if m.group(1) == "dram__bytes_read.sum":
                cur["dram_bytes_read"] = float(m.group(2)) * scale
            elif m.group(1) == "dram__bytes_write.sum":
                cur["dram_bytes_write"] = float(m.group(2)) * scale
            elif m.group(1) == "gpu__time_duration.sum":
                cur["duration_ms_under_ncu"] = float(m.group(2))
    for k in kernels.values():
        k["traffic"] = k.get("dram_bytes_read", 0.0) + k.get("dram_bytes_write", 0.0)
    # phase -> the kernels that make it up (the one-sided attention and pass 2 are two launches each)
    phase_of = [("bwd_edges", ("k_bwd_edges", "k_bwd_sym_lower"), nnz), ("attn_fwd", ("k_attn_", "k_sym_expand"), nnz),
                ("spmm_fwd", ("k_gather_stream<DlMap<8, 16>, 0>", "k_scale_rows"), nnz),
                ("bwd_gather", ("k_gather_stream<DlMap<8, 16>, 1>",), nnz),
                ("pair_fwd", ("k_pair_score_fwd",), pairs), ("pair_bwd", ("k_pair_bwd_stream",), pairs)]
    out = {"workload": f"mid (nnz={nnz}, K=8, d=16, P={pairs}) -- ncu --set full is too slow at c5; bytes scale "
                       "with nnz / P", "source": path, "nnz": nnz, "pairs": pairs, "kernels": kernels}
    for phase, pats, units in phase_of:
        names = [n for n in kernels if any(n.startswith(p) for p in pats)]
        if names:
            tr = sum(kernels[n]["traffic"] for n in names)
            out[phase] = {"kernels": names, "traffic_bytes_per_launch": tr, "per_entry": tr / units}
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]))
