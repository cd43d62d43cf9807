import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): 
        print(line[:300]); continue
    d=json.loads(line)
    print("it/s %.1f  e2e %.1f  ms/step %.4f  roofline frac %.3f  launches %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"] or 0, d["gpu_launches"]))
    print("   ", {k:round(v["ms"]*1e3,1) for k,v in d["kernels"].items()}, "us")
    if "cpu_baseline" in d: print("   cpu", d["cpu_baseline"])
