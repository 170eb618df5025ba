import sys, os, subprocess, json
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def run(batch, env, steps=10):
    e = dict(os.environ, **env)
    r = subprocess.run([sys.executable, "bench.py", "--no-cpu", "--batch", str(batch), "--steps", str(steps), "--warmup", "3"], capture_output=True, text=True, env=e, cwd=root)
    for l in r.stdout.splitlines():
        if l.startswith("{"):
            d = json.loads(l)
            km = d["kernel_ms"]
            print("batch", batch, env, "fps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 2),
                  {k: km[k] for k in ("k_lsd_band_rounds", "k_lsd_spec", "k_lsd_commit") if k in km}, flush=True)
            return
    print("batch", batch, env, "FAILED", r.stderr[-400:], flush=True)
for batch in (32, 64, 128, 256, 512):
    run(batch, {"PLVI_LSD_BR_MAX": "0"})
    run(batch, {"PLVI_LSD_BR_MAX": "1024"})
run(1024, {"PLVI_LSD_BR_MAX": "0"}, 5)
run(1024, {"PLVI_LSD_BR_MAX": "1024"}, 5)
