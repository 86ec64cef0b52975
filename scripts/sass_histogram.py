"""Per-kernel SASS opcode histogram of the shipped library (tcgen05 / TMEM / TMA evidence; runs without a GPU).
usage: python scripts/sass_histogram.py [lib.so] > profiles/r02_sass_histogram.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "ucf_vit_b200/lib/libucfvit_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "SYNCS", "MUFU.EX2", "FFMA2", "HMMA",
         "LDL", "STL"]
cur, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        hist[cur]["_total"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                hist[cur][w] += 1
        if op.startswith("UTCHMMA") and ".2CTA" in op:
            hist[cur]["UTCHMMA.2CTA"] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.splitlines()
print(f"{lib}: {len(hist)} kernels (sm_100a SASS, cuobjdump {subprocess.run(['cuobjdump','--version'],capture_output=True,text=True).stdout.strip().splitlines()[-1]})")
print(f"{'kernel':84s} {'instr':>6s} " + " ".join(f"{w[:8]:>8s}" for w in WATCH))
tot = collections.Counter()
for (k, h), name in zip(hist.items(), demangle):
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("ucf::", "")[:84]
    if not any(h[w] for w in WATCH[:7]) and h["_total"] < 400:
        continue
    print(f"{name:84s} {h['_total']:6d} " + " ".join(f"{h[w]:8d}" for w in WATCH))
    tot.update(h)
print(f"{'TOTAL (all kernels)':84s} {sum(h['_total'] for h in hist.values()):6d} " + " ".join(f"{sum(h[w] for h in hist.values()):8d}" for w in WATCH))
