import json,sys
t=json.load(open(sys.argv[1]))
print({k:t[k] for k in t if k!="phases"})
for k,v in t["phases"].items(): print("  %-14s %6.1f%% %8.0f"%(k,100*v["share"],v["per_iteration"]))
