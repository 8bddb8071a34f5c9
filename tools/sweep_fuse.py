import os, sys, subprocess, json
for v in ("1000000", "888", "444", "240", "100", "30"):
    env = dict(os.environ, MIPM_FRONT_FUSE_MIN=v)
    out = subprocess.run([sys.executable, "tools/profile_factor.py"], env=env, capture_output=True, text=True).stdout
    line = [l for l in out.splitlines() if l.startswith("factor ms")]
    print("C2 fuse_min", v, line)
    out = subprocess.run([sys.executable, "tools/profile_factor_k2.py"], env=env, capture_output=True, text=True).stdout
    line = [l for l in out.splitlines() if l.startswith("factor ms") or l.startswith("solve ms")]
    print("C3 fuse_min", v, line, flush=True)
