"""Runs tools/tune.py's timing for several library builds (RT_B200_LIBNAME) in subprocesses."""
import subprocess, sys, json, os
specs = [a.split(":") for a in sys.argv[1:]]  # libtag:blocks_per_sm[:variant]
for sp in specs:
    tag, bps = sp[0], int(sp[1])
    variant = int(sp[2]) if len(sp) > 2 else 1
    env = dict(os.environ, RT_B200_LIBNAME=f"librt_b200_{tag}.so" if tag != "default" else "librt_b200.so")
    opts = json.dumps({"trace_mode": 1, "traversal_variant": variant, "blocks_per_sm": bps})
    out = subprocess.run([sys.executable, "tools/tune.py", opts], env=env, capture_output=True, text=True)
    print(tag, bps, out.stdout.strip()[-900:] if out.returncode == 0 else out.stderr[-500:], flush=True)
