import json,sys
d=json.load(open(sys.argv[1]))
print(d["value"], d["stage_ms"], d["e2e"]["value"], d["e2e"]["blocking"]["value"], d["roofline"]["frac"], d["roofline_src"]["frac"], d["roofline_pcm"]["frac"])
f=d.get("fast_mode")
if f: print(f["value"], f["stage_ms"], f["e2e"]["value"], f["e2e"]["blocking"]["value"])
