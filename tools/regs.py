"""Print registers / spills per kernel from the ptxas logs of the last build."""
import glob, os, re, subprocess, sys
pat = sys.argv[1] if len(sys.argv) > 1 else ""
for f in sorted(glob.glob(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gdn_b200/build/*.ptxas.log"))):
    txt = open(f).read()
    ents = re.findall(r"Compiling entry function '([^']+)' for 'sm_100a'.*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", txt, re.S)
    for name, stack, ss, sl, regs in ents:
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"\(.*", "", dem)
        if pat in dem:
            print(f"{dem[:64]:64s} regs={regs:>3s} spill={ss}/{sl}")
