import json, sys
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    print("value %.0f %s  ms/step %.3f  e2e %s" % (d["value"], d["unit"], d["ms_per_step"], json.dumps(d.get("e2e"))))
    print("hbm_frac_step", d.get("hbm_frac_step"), "excl_nlm", d.get("hbm_frac_excl_nlm"), "clocks", d.get("clocks"))
    for k in d.get("kernels", []):
        print("  ", k)
    if "cpu_baseline" in d:
        print("cpu", d["cpu_baseline"])
    if "extras" in d:
        print("extras", json.dumps(d["extras"], indent=1))
