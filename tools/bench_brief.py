import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "value %.3f M hyp/s  ms/step %.3f  e2e %.3f M"%(d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6), {k:(round(v["score_kernel_ms"],3),round(v["fit_ms"],3)) for k,v in d["per_primitive"].items()})
