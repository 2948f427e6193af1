"""Prints, for every failing CASE of tools/dbg_local_shard.py output, the host trace lines that precede it (relative ms)."""
import sys

buf = []
for line in open(sys.argv[1]):
    if line.startswith("CASE"):
        if "FAIL" in line:
            ts = [float(l.split()[-2]) for l in buf if l.startswith("[b200 trace]")]
            t0 = min(ts) if ts else 0.0
            for l in buf:
                if l.startswith("[b200 trace]"):
                    f = l.split()
                    print("   ", " ".join(f[2:-2]), f"{float(f[-2]) - t0:9.3f}")
                else:
                    print("   ", l.rstrip()[:200])
            print(line[:120])
        buf = []
    else:
        buf.append(line)
