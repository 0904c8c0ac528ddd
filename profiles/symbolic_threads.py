"""Host symbolic analysis against the number of nested-dissection threads (B200_ND_THREADS), configs 2 and 3."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
print("host cores:", os.cpu_count(), " usable:", len(os.sched_getaffinity(0)))
for t in (1, 2, 4, 8, 16):
    env = dict(os.environ, B200_ND_THREADS=str(t), B200_SYM_TRACE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "symbolic_timing.py")], env=env, capture_output=True, text=True)
    lines = [l.strip() for l in (out.stdout + out.stderr).splitlines() if "nested dissection" in l or "analyze" in l]
    print(f"threads {t:2d}: " + " | ".join(lines))
