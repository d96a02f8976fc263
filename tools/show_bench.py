import json, sys
d = json.load(open(sys.argv[1]))
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d.get("gpu_launches"))
tot = 0
for k, v in d.get("stages", {}).items():
    tot += v["ms"]
    print(f"  {k:16s} {v['ms']*1000:8.1f} us  {v.get('achieved_gbs','')!s:>8} GB/s {v.get('achieved_tflops','')!s:>8} TF")
print(f"  {'sum':16s} {tot*1000:8.1f} us")
