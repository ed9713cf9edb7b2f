"""Summarise the -Xptxas -v logs left by disenlink_b200.build (registers, stack, spills per kernel)."""
import glob
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    for f in sorted(glob.glob(os.path.join(HERE, "_obj", "*.ptxas.log"))):
        txt = open(f).read()
        print("==", os.path.basename(f))
        pat = re.compile(r"Function properties for (\S+)\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, "
                         r"(\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers")
        names = [m.group(1) for m in pat.finditer(txt)]
        dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
        for m, dn in zip(pat.finditer(txt), dem):
            dn = re.sub(r"^void \(anonymous namespace\)::", "", dn)
            dn = re.sub(r"\(.*", "", dn)
            print(f"{m.group(5):>4} regs  stack {m.group(2):>5}  spill {m.group(3)}/{m.group(4)}  {dn}")


if __name__ == "__main__":
    main()
