"""SASS digest of the hot kernels of libdisenlink_b200.so: per kernel, registers / shared memory / spills as
ptxas reported them and the static instruction-class counts that prove the memory path (LDGSTS = cp.async,
LDGSTS...LTC64B = L2::64B prefetch size, REDG = fire-and-forget reduction, no ATOMG on floats) plus the probe's
UTMALDG.2D.GATHER4.   python tools/sass_digest.py > profiles/r02_sass_digest.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "disenlink_b200", "libdisenlink_b200.so")
HOT = ["k_attn_fl<8, 16>", "k_sym_expand<8>", "k_gather_stream<DlMap<8, 16>, 0>", "k_gather_stream<DlMap<8, 16>, 1>",
       "k_bwd_edges_fl<8, 16, 2>", "k_bwd_edges_fl<8, 16, 1>", "k_bwd_sym_lower<8, 16>", "k_gather_chain",
       "k_attn_fl<5, 32>", "k_bwd_edges_fl<5, 32, 2>", "k_bwd_sym_lower<5, 32>",
       "k_pair_score_fwd<DlMap<8, 16>>", "k_pair_bwd_stream<DlMap<8, 16>>",
       "k_scale_rows", "k_push_rows", "k_need_masks"]
CLASSES = ["LDGSTS", "LDG", "STG", "LDS", "STS", "REDG", "ATOMG", "ATOMS", "RED", "SHFL", "FFMA", "FADD", "FMUL", "MUFU",
           "IMAD", "BRA", "LDL", "STL", "UTMALDG", "UBLKCP", "HMMA", "UTCHMMA"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def digest(path, wanted):
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    blocks = re.split(r"\n\s*Function : ", sass)[1:]
    names = [b.split("\n", 1)[0].strip() for b in blocks]
    dm = demangle(names)
    for name, b in zip(names, blocks):
        pretty = re.sub(r"\(anonymous namespace\)::", "", dm.get(name, name))
        pretty = re.sub(r"^void ", "", pretty)
        short = re.sub(r"\(.*", "", pretty)
        if not any(short == w or short.startswith(w) for w in wanted):
            continue
        ops = collections.Counter()
        flavours = collections.Counter()
        for line in b.split("\n"):
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                op = m.group(1)
                ops[op.split(".")[0]] += 1
                if op.startswith(("LDGSTS", "REDG", "UTMALDG", "ATOMG")):
                    flavours[op] += 1
        total = sum(ops.values())
        print(f"{short}\n    {total} SASS instructions; " + ", ".join(f"{c} {ops[c]}" for c in CLASSES if ops.get(c)))
        if flavours:
            print("    " + ", ".join(f"{k} x{v}" for k, v in sorted(flavours.items())))


if __name__ == "__main__":
    print("# SASS digest, sm_100a --", os.path.relpath(LIB, ROOT))
    digest(LIB, HOT)
    probe = os.path.join(ROOT, "tools", "_bin", "gather_probe")
    if os.path.exists(probe):
        print("\n# tools/gather_probe (random-gather ceiling probe)")
        digest(probe, ["k_gather4<64", "k_cpasync<64, 4, 4>", "k_ldg<64"])
