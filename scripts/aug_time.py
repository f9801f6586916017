"""GPU timing aid: the bench's augment / mel / embed stage times only (no e2e, no CPU baseline)."""
import subprocess, sys, json, os
env = dict(os.environ)
out = subprocess.run([sys.executable, "bench.py", "--no-cpu-baseline", "--steps", "12", "--warmup", "3"], capture_output=True, text=True, env=env).stdout
d = json.loads(out.strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"], 3), {k: round(v["ms_per_step"], 3) for k, v in d["stages"].items()}, "e2e", round(d["e2e"]["value"]))
