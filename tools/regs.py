"""Print registers / spills per kernel from the ptxas logs the build writes (build/*.ptxas.log)."""
import re, sys, glob, os, subprocess
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "neighbour_feature_pooling_b200", "build")
pat = sys.argv[1] if len(sys.argv) > 1 else "stream"
for log in sorted(glob.glob(os.path.join(root, f"*{pat}*.ptxas.log"))):
    txt = open(log).read()
    for m in re.finditer(r"Compiling entry function '(\S+)'.*?\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", txt):
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"nfp::stream::|\(int\)|void |\(nfp.*", "", name)
        print(f"{os.path.basename(log)[:18]:18s} regs {m.group(5):>3s} stack {m.group(2):>4s} spill st/ld {m.group(3)}/{m.group(4)}  {name}")
