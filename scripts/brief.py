import json,sys
for ln in sys.stdin:
    ln=ln.strip()
    if not ln.startswith('{'): continue
    d=json.loads(ln)
    print("ms/step", round(d["ms_per_step"],2), "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],2), "roof", d["roofline"]["kernel"], round(d["roofline"]["frac"],3))
    for k,v in d["detail"]["kernels"].items(): print("   ", k, round(v["ms_per_step"],3), v["GBps"] and round(v["GBps"]))
    print("    glue", round(d["detail"]["torch_glue_ms_per_step"],2))
