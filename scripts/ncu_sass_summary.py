"""Summarise `ncu --page source --csv --print-source sass` output: instruction mix + stall samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
tot, stot, samples = {}, {}, 0
hot = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        ix = {h: i for i, h in enumerate(hdr)}
        stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].startswith("0x"):
        continue
    toks = [o for o in r[ix["Source"]].split() if not o.startswith("@")]
    op = toks[0].split(".")[0]
    ex = int(r[ix["Instructions Executed"]])
    tot[op] = tot.get(op, 0) + ex
    s = int(r[ix["# Samples"]])
    samples += s
    hot.append((s, r[ix["Source"]].strip(), {h: int(r[ix[h]]) for h in stall_cols if int(r[ix[h]])}))
    for h in stall_cols:
        stot[h] = stot.get(h, 0) + int(r[ix[h]])
print("instruction mix (warp-level executed):")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:16]:
    print("  %-10s %d" % (k, v))
print("stall samples: total", samples)
for k, v in sorted(stot.items(), key=lambda kv: -kv[1])[:10]:
    print("  %-24s %8d  %.1f%%" % (k, v, 100.0 * v / max(1, samples)))
print("hottest instructions:")
for s, src, st in sorted(hot, key=lambda t: -t[0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    print("  %6d  %-60s %s" % (s, src[:60], dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])))
