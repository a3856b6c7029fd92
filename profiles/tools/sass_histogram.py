"""Per-kernel SASS opcode histogram of the shipped library (cuobjdump -sass), with the
Blackwell / Hopper-class data-movement opcodes called out.
usage: python profiles/tools/sass_histogram.py [lib.so] > profiles/rNN_sass_opcodes.json"""
import collections, json, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "speedy_b200/libspeedy_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn = None
hist = collections.defaultdict(collections.Counter)
for line in txt.split("\n"):
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r"\(speedy::K\dParams.*", "", fn).replace("void speedy::", "")
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and fn:
        hist[fn][m.group(1)] += 1
special = ["UBLKCP", "UBLKPF", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "REDUX", "CREDUX", "VABSDIFF", "IDP"]
out = {"library": lib, "kernels": {}, "totals": collections.Counter()}
for k, c in sorted(hist.items()):
    out["kernels"][k] = {"instructions": sum(c.values()), "top": dict(c.most_common(12)),
                         "async_and_special": {s: c[s] for s in special if c[s]}}
    out["totals"].update(c)
out["totals"] = {s: out["totals"][s] for s in special if out["totals"][s]}
print(json.dumps(out, indent=1))
