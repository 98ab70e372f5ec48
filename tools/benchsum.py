import json,sys
lines=[l for l in open(sys.argv[1]) if l.startswith("{")]
d=json.loads(lines[-1]);print("n_gpus",d["n_gpus"],"ms/step",d["ms_per_step"],"it/s",d["value"],"e2e",d["e2e"]["value"])
for k in d["roofline"]["kernels"]: print("  ",k["group"],round(k["ms"]*1e3,1),"us",round(k["gbs"]),"GB/s")
