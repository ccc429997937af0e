import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]) if "e2e" in d else None, "ms/step", round(d["ms_per_step"],3), {k: round(v,4) for k,v in d["roofline"]["kernels_ms_per_step"].items()}, d["icp"])
