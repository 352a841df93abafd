"""Warp-stall samples of an .ncu-rep summed between marker instructions (TMA / MMA / TMEM / global
loads and stores / barrier waits), in program order: shows which warp ROLE of a warp-specialised
kernel the time goes to.  usage: python tools/ncu_regions.py report.ncu-rep [regex-of-markers]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else r'UTCHMMA|UTMALDG|LDTM|STG|SYNCS|UTCBAR|BAR\.SYNC|EXIT|LDG|FENCE|STS\.128|LDS\.128')
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
isrc, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
tot = sum(int(r[isamp] or 0) for r in data)
print('total samples', tot, 'instructions', len(data))
cur = 0
last = None
for k, r in enumerate(data):
    cur += int(r[isamp] or 0)
    if pat.search(r[isrc]):
        key = re.sub(r'\s+', ' ', r[isrc])[:70]
        if cur >= max(20, tot // 400) or key != last:
            print(f'{k:5d} {cur:6d} {r[iex]:>9s}  {key}')
        cur = 0
        last = key
